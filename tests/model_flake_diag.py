"""Diagnostic (GPU): runs the golden model window N times and reports every run whose frequency-axis onsets differ from the
first run (the model is deterministic: any difference is a race), with the 128-token decoder tiles that differ."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_diag import make_extractor
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
ex, _ = make_extractor(max_windows=4)
z = np.load(os.path.join(ROOT, "tests", "golden", "model_window.npz"))
x = torch.from_numpy(z["input_spec"]).cuda()
ref = None
bad_runs = 0
for rep in range(n):
    o = ex.model(x)
    on_f, on_t = o[0].cpu().numpy()[0], o[5].cpu().numpy()[0]
    if ref is None:
        ref = (on_f, on_t)
        print("run 0 vs golden: onset_f max err", float(np.abs(on_f - z["onset_f"][0]).max()))
        continue
    if not (np.array_equal(on_f, ref[0]) and np.array_equal(on_t, ref[1])):
        bad_runs += 1
        d = np.abs(on_f - ref[0]).reshape(-1)            # token = frame * 88 + note
        tiles = sorted(set((np.nonzero(d > 0)[0] // 128).tolist()))
        print(f"run {rep}: differs; onset_f max diff {d.max():.4f}; decoder 128-token tiles that differ: {tiles[:24]} ({len(tiles)} tiles); "
              f"time-axis max diff {np.abs(on_t - ref[1]).max():.4f}")
print(f"{bad_runs} of {n - 1} runs differ from run 0")
