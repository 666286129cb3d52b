"""Diagnostic (GPU): where one config-3 step (one 4-min song, 30 windows) spends its time -- host-side stage timestamps,
the per-launch device timeline of a profiled pass with the idle gaps between consecutive launches, and the unprofiled
step time next to the sum of kernel durations.  Run: python tests/config3_timeline.py > gpurun_out/config3_timeline.txt"""
import os, sys, time, tempfile
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from etude_b200 import AMTAPC_Extractor, ExtractorConfig
from etude_b200.weights import default_state_dict
from etude_b200.synth import noise

dev = torch.device("cuda", 0)
ckpt = os.path.join(tempfile.gettempdir(), "etude_t3.pth")
torch.save(default_state_dict(seed=0), ckpt)
ex = AMTAPC_Extractor(ExtractorConfig(), ckpt, device=dev, max_windows=64)
eng = ex.engine
n = 3_840_000
wave = torch.from_numpy(noise(n, 1234)).to(dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

def step():
    return ex.extract_many([n], as_dicts=False, wave_dev=wave)

for _ in range(3):
    step()
torch.cuda.synchronize()
ts = []
for _ in range(10):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    h0 = time.perf_counter()
    a.record(); step(); b.record()
    torch.cuda.synchronize()
    ts.append((a.elapsed_time(b), (time.perf_counter() - h0) * 1e3))
print("unprofiled steps (device ms, host ms):", [(round(x, 2), round(y, 2)) for x, y in ts])
eng.profile_reset(timing=True)
flush.fill_(1); torch.cuda.synchronize()
step()
tl = eng.profile_timeline()
eng.profile_reset(timing=False)
print(f"profiled pass: {len(tl)} launches, span {tl[-1][1] + tl[-1][2]:.2f} ms, sum of kernels {sum(d for _, _, d in tl):.2f} ms")
prev_end = 0.0
for i, (c, s, d) in enumerate(tl):
    gap = s - prev_end
    print(f"{i:3d} {c:16s} start {s:8.3f} dur {d:7.3f} gap {gap:7.3f}" + ("   <<<" if gap > 0.05 else ""))
    prev_end = max(prev_end, s + d)

# ---- host-side cost of each engine call of one step (no synchronisation added: a call that waits for the GPU shows it)
import functools
host = []
def wrap(name):
    f = getattr(eng, name)       # bound (or static) attribute, shadowed on the instance
    @functools.wraps(f)
    def g(*a, **k):
        t0 = time.perf_counter()
        r = f(*a, **k)
        host.append((name, t0, time.perf_counter()))
        return r
    setattr(eng, name, g)
for name in ("alloc_rolls", "notes_reserve", "logmel", "forward_windows", "notes"):
    wrap(name)
for rep in range(3):
    host.clear()
    flush.fill_(1); torch.cuda.synchronize()
    h0 = time.perf_counter()
    step()
    h1 = time.perf_counter()
    torch.cuda.synchronize()
    h2 = time.perf_counter()
    print(f"-- rep {rep}: extract_many host {1e3 * (h1 - h0):.2f} ms (+ {1e3 * (h2 - h1):.2f} ms to idle)")
    for name, a, b in host:
        print(f"   {name:16s} enter +{1e3 * (a - h0):7.3f} ms   host time {1e3 * (b - a):7.3f} ms")
