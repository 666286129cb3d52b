"""CPU restatement of the audio ingest in front of the log-mel (reference etude/data/extractor.py:180-184).

TEST INFRASTRUCTURE (see oracle/__init__.py): only tests, smoke() and bench.py's cpu_baseline may import this.

The reference calls ``torch.mean(wave, dim=0)`` and ``torchaudio.transforms.Resample(sr, 16000)`` with torchaudio's
defaults (sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99).  torchaudio is a third-party dependency, not part of
/root/reference; its published algorithm (torchaudio/functional/functional.py, ``_get_sinc_resample_kernel`` /
``_apply_sinc_resample_kernel``, pinned 2.6.0 by the reference, 2.11.0 installed here) is restated below and pinned
against the installed torchaudio in tests/test_oracle_golden.py::test_resample_oracle_matches_torchaudio.
"""
import math

import numpy as np

LOWPASS_FILTER_WIDTH = 6
ROLLOFF = 0.99


def sinc_kernel(orig_freq, new_freq):
    """[new/g, 2 width + orig/g] float32 polyphase kernel and `width`, exactly as torchaudio builds it: indices in
    float64, the phase term -p / new_freq rounded to float32 first (an int64 tensor divided by an int is float32)."""
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base_freq = min(orig, new) * ROLLOFF
    width = math.ceil(LOWPASS_FILTER_WIDTH * orig / base_freq)
    idx = np.arange(-width, width + orig, dtype=np.float64)[None, :] / orig
    phase = (np.arange(0, -new, -1).astype(np.float32) / np.float32(new)).astype(np.float64)[:, None]
    t = (phase + idx) * base_freq
    t = np.clip(t, -LOWPASS_FILTER_WIDTH, LOWPASS_FILTER_WIDTH)
    window = np.cos(t * math.pi / LOWPASS_FILTER_WIDTH / 2) ** 2
    t = t * math.pi
    scale = base_freq / orig
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(t == 0, 1.0, np.sin(t) / t)
    return (k * window * scale).astype(np.float32), width, orig, new


def resample(wave, orig_freq, new_freq=16000):
    """wave: float32 [C, N] (or [N]) -> mono float32 [ceil(new N / orig)]: channel mean, then sinc resampling."""
    x = np.asarray(wave, dtype=np.float32)
    if x.ndim == 2:
        x = (x.sum(axis=0, dtype=np.float32) / np.float32(x.shape[0])).astype(np.float32) if x.shape[0] > 1 else x[0]
    if int(orig_freq) == int(new_freq):
        return x
    kern, width, orig, new = sinc_kernel(orig_freq, new_freq)
    n = x.shape[0]
    pad = np.zeros(n + 2 * width + orig, np.float32)
    pad[width : width + n] = x
    n_win = n // orig + 1
    K = kern.shape[1]
    frames = np.lib.stride_tricks.as_strided(pad, shape=(n_win, K), strides=(pad.strides[0] * orig, pad.strides[0]))
    out = (frames.astype(np.float64) @ kern.astype(np.float64).T).astype(np.float32).reshape(-1)   # [n_win, new] row-major
    target = int(math.ceil(new * n / orig))
    return out[:target]
