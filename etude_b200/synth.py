"""Seeded synthetic 16 kHz mono audio: the inputs of tests and bench.

There is no network and no dataset on the GPU box, so every measurement and parity test uses these
generators; the same tensors are fed to the oracle and to the CUDA path.
"""
import numpy as np
import torch

SR = 16000


def noise(n_samples, seed=1234):
    """Uniform noise in [-0.5, 0.5): ``(rand(N)*2-1)*0.5`` with torch.Generator(seed)."""
    g = torch.Generator().manual_seed(int(seed))
    return ((torch.rand(int(n_samples), generator=g) * 2 - 1) * 0.5).numpy()


def tones(n_samples, seed=4321):
    """Piano-like decaying harmonic notes over a 1e-3 noise floor (exercises log-mel dynamic range).

    Six notes start per second; fundamentals from MIDI 21..108, 8 harmonics with 1/h amplitude,
    exponential decay tau = 0.5 s, peak 0.1..0.3; clipped to [-1, 1].
    """
    rng = np.random.default_rng(int(seed))
    n = int(n_samples)
    out = rng.normal(0.0, 1e-3, n)
    n_notes = max(1, int(6 * n / SR))
    for _ in range(n_notes):
        start = int(rng.integers(0, max(1, n - 1)))
        midi = int(rng.integers(21, 109))
        f0 = 440.0 * 2.0 ** ((midi - 69) / 12.0)
        peak = float(rng.uniform(0.1, 0.3))
        length = min(n - start, 2 * SR)
        t = np.arange(length) / SR
        env = peak * np.exp(-t / 0.5)
        sig = np.zeros(length)
        for h in range(1, 9):
            if f0 * h < SR / 2:
                sig += np.sin(2 * np.pi * f0 * h * t) / h
        out[start : start + length] += env * sig
    return np.clip(out, -1.0, 1.0).astype(np.float32)
