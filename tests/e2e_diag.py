"""Where does the end-to-end leg lose time against the device-resident leg?  (diagnostic, run on the GPU box)

Times extract_many over S synthetic songs (a) as shipped, (b) with the note stage stubbed out, (c) with the host staging
copy stubbed out (waves already in the pinned buffer), next to the device-resident transcribe_device."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from etude_b200 import AMTAPC_Extractor, ExtractorConfig, synth  # noqa: E402
from etude_b200.weights import default_state_dict  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 16
torch.save(default_state_dict(seed=0), "/tmp/e2e_diag_sd.pth")
ex = AMTAPC_Extractor(ExtractorConfig(), "/tmp/e2e_diag_sd.pth", device="cuda:0", max_windows=32)
waves = [synth.noise(16000 * 240, 1234 + i) for i in range(S)]
n = [len(w) for w in waves]
off = np.concatenate([[0], np.cumsum(n)]).astype(np.int64)
pinned = torch.empty(int(off[-1]), dtype=torch.float32, pin_memory=True)
pinned.numpy()[:] = np.concatenate(waves)
dev = pinned.to("cuda:0")
secs = sum(n) / 16000


def timed(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps


t_dev = timed(lambda: ex.transcribe_device(dev, off[:-1], n))
print(f"device-resident: {1e3 * t_dev:8.1f} ms  {secs / t_dev:8.0f} audio-s/s")
t_e2e = timed(lambda: ex.extract_many(waves, as_dicts=False, pinned=pinned))
print(f"e2e as shipped : {1e3 * t_e2e:8.1f} ms  {secs / t_e2e:8.0f} audio-s/s")
real_notes = ex.engine.notes
ex.engine.notes = lambda on, of, mp, ve, ro, rows, *a, **k: [np.zeros(0, dtype=[("pitch", np.int32)]) for _ in rows]
t_nonotes = timed(lambda: ex.extract_many(waves, as_dicts=False, pinned=pinned))
print(f"e2e, no notes  : {1e3 * t_nonotes:8.1f} ms  {secs / t_nonotes:8.0f} audio-s/s")
ex.engine.notes = real_notes
views = [pinned[o:o + k].numpy() for o, k in zip(off[:-1], n)]   # staging becomes a self-copy of pinned memory
t_nostage = timed(lambda: ex.extract_many(views, as_dicts=False, pinned=pinned))
print(f"e2e, waves already pinned (self-copy): {1e3 * t_nostage:8.1f} ms  {secs / t_nostage:8.0f} audio-s/s")
t_onegroup = timed(lambda: ex.extract_many(waves, as_dicts=False, pinned=pinned, group_songs=None))
print(f"e2e, one group (no pipeline): {1e3 * t_onegroup:8.1f} ms  {secs / t_onegroup:8.0f} audio-s/s")
for g, nb in ((4, 1), (4, 4), (4, 8), (4, 16), (2, 4)):
    t_g = timed(lambda: ex.extract_many(waves, as_dicts=False, pinned=pinned, group_songs=g, notes_batch=nb))
    print(f"e2e, group_songs={g} notes_batch={nb}: {1e3 * t_g:8.1f} ms  {secs / t_g:8.0f} audio-s/s")

# ---- per-group GPU spans on the launching stream during a shipped e2e run
import etude_b200.extractor as X  # noqa: E402
spans = []
orig = ex.transcribe_device


def traced(wave_dev, wave_off, n_samples):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = orig(wave_dev, wave_off, n_samples)
    e1.record()
    spans.append((len(n_samples), e0, e1, time.perf_counter()))
    return r


ex.transcribe_device = traced
ex.extract_many(waves, as_dicts=False, pinned=pinned)
spans.clear()
torch.cuda.synchronize()
t0 = time.perf_counter()
ex.extract_many(waves, as_dicts=False, pinned=pinned)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
first = spans[0][1]
print(f"e2e wall {1e3 * wall:.1f} ms; per group: songs, GPU span ms (ms per song), start / end offsets on the stream, host enqueue-done time")
tot = 0.0
for k, e0, e1, th in spans:
    ms = e0.elapsed_time(e1)
    tot += ms
    print(f"  {k} songs: {ms:7.1f} ms ({ms / k:5.1f}/song)  [{first.elapsed_time(e0):7.1f} .. {first.elapsed_time(e1):7.1f}]  host {1e3 * (th - t0):7.1f}")
print(f"sum of model-group spans {tot:.1f} ms vs device-resident {1e3 * t_dev:.1f} ms")
