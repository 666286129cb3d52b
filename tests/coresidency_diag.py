"""Do the note-decoding kernels co-reside with the persistent model kernels?  (diagnostic, GPU box)

Times one chain / attention / projection launch alone, then again while a notes call for 4 songs runs on another stream."""
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_diag as D  # noqa: E402
from etude_b200 import _lib  # noqa: E402

ex, sd = D.make_extractor(max_windows=32)
eng = ex.engine
lib = eng.lib
P, stream = D.P, D.stream
torch.manual_seed(0)
bf = lambda t: t.to(torch.bfloat16)
M = 128 * 148 * 16
ctx = bf(torch.randn(M, 256, device="cuda") * 0.7)
wo, w1, w2 = bf(torch.randn(256, 256, device="cuda") / 16), bf(torch.randn(512, 256, device="cuda") / 16), bf(torch.randn(256, 512, device="cuda") / 22)
bo, b1, b2 = (0.1 * torch.randn(k, device="cuda") for k in (256, 512, 256))
gamma, beta = 1 + 0.1 * torch.randn(256, device="cuda"), 0.1 * torch.randn(256, device="cuda")
x = bf(torch.randn(M, 256, device="cuda"))
S, L = 2048, 256
qkv = torch.randn(S * L, 768, device="cuda").to(torch.bfloat16)
aout = torch.zeros((S * L, 256), dtype=torch.bfloat16, device="cuda")
wq = bf(torch.randn(768, 256, device="cuda") / 16)
bq = 0.1 * torch.randn(768, device="cuda")
gout = torch.zeros((M, 768), dtype=torch.bfloat16, device="cuda")


def chain():
    _lib.check(lib.etude_k_chain(P(ctx), P(wo), P(bo), P(w1), P(b1), P(w2), P(b2), P(gamma), P(beta), P(x), 0, M, P(x), M, stream()), "chain")


def attn():
    _lib.check(lib.etude_k_attention(P(qkv), S * L, 768, 0, L, P(qkv), 768, 256, 512, S, L, L, P(aout), None, stream()), "attn")


def gemm():
    _lib.check(lib.etude_k_gemm(P(ctx), P(wq), P(bq), M, 768, 256, 0, P(gout), None, 0, None, None, None, stream()), "gemm")


# rolls for 4 songs at random-init-like density
T = 15360
n_songs = 4
rng = np.random.default_rng(0)
sig = lambda a: (1.0 / (1.0 + np.exp(-a))).astype(np.float32)
on = torch.from_numpy(sig(rng.normal(1.0, 1.0, (T * n_songs, 88)))).cuda()
off = torch.from_numpy(sig(rng.normal(12.0, 6.0, (T * n_songs, 88)))).cuda()
mpe = torch.from_numpy(sig(rng.normal(0.5, 1.5, (T * n_songs, 88)))).cuda()
vel = torch.from_numpy(rng.integers(0, 128, (T * n_songs, 88)).astype(np.int8)).cuda()
row_off = [i * T for i in range(n_songs)]
rows = [T] * n_songs
side = torch.cuda.Stream()


def notes():
    with torch.cuda.stream(side):
        return eng.notes(on, off, mpe, vel, row_off, rows, 0.5, 1.0, 0.5)


notes()
torch.cuda.synchronize()
t = time.perf_counter(); notes(); torch.cuda.synchronize()
t_notes = time.perf_counter() - t
print(f"notes alone (4 songs): {1e3 * t_notes:.1f} ms")
main2 = torch.cuda.Stream()
if os.environ.get("MODEL_ON_SIDE_STREAM"):
    print("model launches on a non-default stream")
    torch.cuda.set_stream(main2)
for name, fn, reps in (("chain", chain, 12), ("attention", attn, 12), ("projection", gemm, 12)):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    alone = e0.elapsed_time(e1)
    th = threading.Thread(target=notes)
    th.start()
    time.sleep(0.004)   # let the first notes kernels get onto the SMs
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    th.join(); torch.cuda.synchronize()
    both = e0.elapsed_time(e1)
    print(f"{name:10s}: {reps} launches alone {alone:7.2f} ms, beside a notes call {both:7.2f} ms  (+{both - alone:6.2f} ms)")
