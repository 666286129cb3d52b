"""Runs the fused attention kernel at one shape a few times (target of `ncu -k regex:attention`): encoder self-attention by default."""
import ctypes
import sys

import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from etude_b200 import _lib  # noqa: E402

S, L = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (8192, 256)
lib = _lib.load()
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
qkv = torch.randn(S * L, 768, device="cuda").to(torch.bfloat16)
out = torch.zeros((S * L, 256), dtype=torch.bfloat16, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(4):
    _lib.check(lib.etude_k_attention(P(qkv), S * L, 768, 0, L, P(qkv), 768, 256, 512, S, L, L, P(out), None, st), "attn")
torch.cuda.synchronize()
print("ok", S, L)
