"""Oracle: log-mel front-end (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates what ``AMTAPC_Extractor._wav2feature`` computes for an already-mono,
already-16 kHz waveform (reference: etude/data/extractor.py:178-197):

    torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=2048,
        win_length=2048, hop_length=256, n_mels=256, norm="slaney")
    -> log(mel + 1e-8) -> transpose -> [T, 256]

with torchaudio's defaults center=True, pad_mode="reflect", periodic Hann,
power=2, onesided, mel_scale="htk", f_min=0, f_max=sr/2.  The arithmetic is
third-party (torchaudio ``functional.melscale_fbanks`` / ``spectrogram``;
pinned 2.6.0 by the reference, 2.11.0 installed) and is restated from its
published formulae in float64, then rounded to float32.
"""
import math

import numpy as np

SR = 16000
N_FFT = 2048
HOP = 256
N_MELS = 256
N_FREQ = N_FFT // 2 + 1
LOG_OFFSET = 1e-8


def hz_to_mel_htk(f):
    return 2595.0 * np.log10(1.0 + np.asarray(f, dtype=np.float64) / 700.0)


def mel_to_hz_htk(m):
    return 700.0 * (10.0 ** (np.asarray(m, dtype=np.float64) / 2595.0) - 1.0)


def mel_filterbank(n_freqs=N_FREQ, f_min=0.0, f_max=SR / 2.0, n_mels=N_MELS, sr=SR):
    """torchaudio.functional.melscale_fbanks(norm="slaney", mel_scale="htk").

    Returns fb[n_freqs, n_mels] float64.  Triangles on linspace(0, sr//2,
    n_freqs) with corner points equally spaced on the HTK mel scale, each
    filter scaled by 2 / (f_hi - f_lo) (slaney area normalisation).
    """
    all_freqs = np.linspace(0.0, sr // 2, n_freqs)
    m_pts = np.linspace(hz_to_mel_htk(f_min), hz_to_mel_htk(f_max), n_mels + 2)
    f_pts = mel_to_hz_htk(m_pts)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    enorm = 2.0 / (f_pts[2 : n_mels + 2] - f_pts[:n_mels])
    return fb * enorm[None, :]


def num_frames(n_samples):
    """center=True STFT frame count: 1 + floor(N / hop)."""
    return 1 + n_samples // HOP


def logmel(wave, dtype=np.float64):
    """wave[N] -> log-mel feature [T, 256] (float32), T = 1 + N // 256.

    Frame t covers reflect-padded samples [t*256, t*256 + 2048), i.e. original
    samples centred on t*256 (torch.stft center=True, pad_mode="reflect").
    """
    x = np.asarray(wave, dtype=dtype).reshape(-1)
    n = x.shape[0]
    pad = N_FFT // 2
    xp = np.pad(x, (pad, pad), mode="reflect")
    t = num_frames(n)
    idx = np.arange(t)[:, None] * HOP + np.arange(N_FFT)[None, :]
    k = np.arange(N_FFT, dtype=np.float64)
    window = (0.5 - 0.5 * np.cos(2.0 * math.pi * k / N_FFT)).astype(dtype)  # periodic Hann
    fb = mel_filterbank().astype(dtype)
    out = np.empty((t, N_MELS), dtype=np.float32)
    step = 4096
    for s in range(0, t, step):
        fr = xp[idx[s : s + step]] * window[None, :]
        spec = np.fft.rfft(fr, axis=1)
        power = spec.real.astype(dtype) ** 2 + spec.imag.astype(dtype) ** 2
        mel = power @ fb
        out[s : s + step] = np.log(mel + dtype(LOG_OFFSET)).astype(np.float32)
    return out
