"""Measured precision of tensor-core DFT formulations for the log-mel front-end: the 2048-point real DFT
as a 32 x 64 two-stage matrix DFT with the operand splits a tcgen05 kernel could use, emulated in numpy (every product of
one MMA is the exact product of the rounded operands, accumulation in fp32 like TMEM), against the float64 restatement.
Inputs: the reference-generated goldens (tests/golden/logmel.npz); metric: max |log-mel - reference| (tolerance 1e-3).

    python tools/frontend_precision_table.py > profiles/r2_frontend_precision_table.txt
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import logmel as ologmel  # noqa: E402  (tool, not product)


def round_to(x, mant_bits):
    """Round fp32 values to `mant_bits` explicit mantissa bits (bf16: 7, fp16: 10 (ignoring its range), tf32: 10)."""
    x = np.asarray(x, np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    drop = 23 - mant_bits
    u = (u + (1 << (drop - 1)) - 1 + ((u >> drop) & 1)) >> drop << drop       # round to nearest even
    return u.astype(np.uint32).view(np.float32)


def split(x, mant_bits, terms):
    parts, r = [], np.asarray(x, np.float32)
    for _ in range(terms):
        p = round_to(r, mant_bits)
        parts.append(p)
        r = (r - p).astype(np.float32)
    return parts


def mm_split(a, b, mant_bits, a_terms, b_terms, max_order):
    """sum over (i, j) with i + j <= max_order of A_i @ B_j, fp32 accumulate (float64 here: the MMA accumulator is not the issue)."""
    A, B = split(a, mant_bits, a_terms), split(b, mant_bits, b_terms)
    out = np.zeros((a.shape[0], b.shape[1]), np.float64)
    n = 0
    for i, ai in enumerate(A):
        for j, bj in enumerate(B):
            if i + j <= max_order:
                out += ai.astype(np.float64) @ bj.astype(np.float64)
                n += 1
    return out.astype(np.float32), n


def dft2048_two_stage(frames, mm):
    """frames [F, 2048] real (windowed).  n = 64 n1 + n2, k = k1 + 32 k2: X[k1 + 32 k2] = sum_n2 W2048^(n2 k1) W64^(n2 k2) sum_n1 W32^(n1 k1) x.
    Real / imaginary parts carried as separate real matrices (what an MMA would do).  Returns (re, im) [F, 1025]."""
    F = frames.shape[0]
    x = frames.reshape(F, 32, 64)                                   # [F, n1, n2]
    n1 = np.arange(32)
    w32 = np.exp(-2j * np.pi * np.outer(n1, n1) / 32)               # [n1, k1]
    a = x.transpose(0, 2, 1).reshape(F * 64, 32)                    # rows (F, n2), K = n1
    re1, c1 = mm(a, w32.real.astype(np.float32))
    im1, _ = mm(a, w32.imag.astype(np.float32))
    y = (re1.astype(np.float64) + 1j * im1.astype(np.float64)).reshape(F, 64, 32)      # [F, n2, k1]
    tw = np.exp(-2j * np.pi * np.outer(np.arange(64), n1) / 2048)   # [n2, k1]
    y = (y * tw).astype(np.complex64)                               # twiddles in fp32 on the CUDA cores
    n2 = np.arange(64)
    w64 = np.exp(-2j * np.pi * np.outer(n2, n2) / 64)               # [n2, k2]
    b = y.transpose(0, 2, 1).reshape(F * 32, 64)                    # rows (F, k1), K = n2
    rr, c2 = mm(b.real.astype(np.float32), w64.real.astype(np.float32))
    ii, _ = mm(b.imag.astype(np.float32), w64.imag.astype(np.float32))
    ri, _ = mm(b.real.astype(np.float32), w64.imag.astype(np.float32))
    ir, _ = mm(b.imag.astype(np.float32), w64.real.astype(np.float32))
    z = ((rr.astype(np.float64) - ii) + 1j * (ri.astype(np.float64) + ir)).reshape(F, 32, 64)   # [F, k1, k2]
    X = z.transpose(0, 2, 1).reshape(F, 2048)                       # k = k1 + 32 k2
    return X[:, :1025], c1


def logmel_with(wave, mm):
    n = len(wave)
    pad = np.pad(np.asarray(wave, np.float64), 1024, mode="reflect")
    T = 1 + n // 256
    idx = np.arange(T)[:, None] * 256 + np.arange(2048)[None]
    win = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(2048) / 2048)
    frames = (pad[idx] * win).astype(np.float32)
    X, per_mm = dft2048_two_stage(frames, mm)
    power = (X.real.astype(np.float64) ** 2 + X.imag.astype(np.float64) ** 2).astype(np.float32)
    fb = ologmel.mel_filterbank()                                   # [1025, 256] float64
    return np.log(power.astype(np.float64) @ fb + 1e-8).astype(np.float32), per_mm


def main():
    z = np.load(os.path.join(ROOT, "tests", "golden", "logmel.npz"))
    schemes = [
        ("fp32 operands (reference point: the two-stage DFT itself)", lambda a, b: (a.astype(np.float64) @ b.astype(np.float64), 1)),
        ("bf16 x bf16 (1 MMA)", lambda a, b: mm_split(a, b, 7, 1, 1, 0)),
        ("bf16 hi/lo both operands, 3 MMAs (hh + hl + lh)", lambda a, b: mm_split(a, b, 7, 2, 2, 1)),
        ("bf16 3-term both operands, 6 MMAs (order <= 2)", lambda a, b: mm_split(a, b, 7, 3, 3, 2)),
        ("fp16 hi/lo both operands, 3 MMAs", lambda a, b: mm_split(a, b, 10, 2, 2, 1)),
        ("tf32 hi/lo both operands, 3 MMAs (3xTF32)", lambda a, b: mm_split(a, b, 10, 2, 2, 1)),
        ("tf32 hi/lo, 4 MMAs (all products)", lambda a, b: mm_split(a, b, 10, 2, 2, 2)),
    ]
    print("max |log-mel - reference| on the reference-generated goldens (tolerance 1e-3); two-stage 32 x 64 matrix DFT, numpy emulation")
    print(f"{'scheme':62s} {'MMAs':>4s} " + " ".join(f"{c:>13s}" for c in ("noise_1s", "tones_2s", "noise_ragged")))
    for name, mm in schemes:
        errs, cnt = [], 1
        for case in ("noise_1s", "tones_2s", "noise_ragged"):
            got, cnt = logmel_with(z[case + "_wave"], mm)
            errs.append(float(np.abs(got - z[case + "_feat"]).max()))
        print(f"{name:62s} {cnt:4d} " + " ".join(f"{e:13.3e}" for e in errs))


if __name__ == "__main__":
    main()
