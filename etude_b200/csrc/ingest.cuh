// Audio ingest in front of the log-mel: channel mean + polyphase sinc resampling to 16 kHz on the device
// (reference etude/data/extractor.py:180-184: torch.mean(wave, dim=0) -> torchaudio.transforms.Resample(sr, 16000) with
// torchaudio's defaults: sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99).
//
//     y[i * new + p] = sum_k kern[p][k] * xm[i * orig + k - width],   xm = channel mean (zero outside the clip)
//
// orig / new are the rates divided by their gcd (441 / 160 for 44.1 kHz), K = 2 width + orig taps per phase (475), the
// table is built on the host exactly like torchaudio's _get_sinc_resample_kernel.  15 MFLOP per audio-second (the model is
// 125 GFLOP): one thread per output sample, neighbouring threads read overlapping input (L1), the table sits in L2.
#pragma once
#include "common.cuh"

namespace etude {

struct IngestParams {
    const float* pcm;     // [channels][n_in] planar fp32
    const float* kern;    // [new][K]
    float* out;           // [n_out]
    int64_t n_in, n_out;
    int channels, orig, nw, width, K;
};

__global__ void __launch_bounds__(256) ingest_kernel(const IngestParams p) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p.n_out) return;
    const float inv_c = 1.f / (float)p.channels;
    if (p.kern == nullptr) {  // same rate: channel mean only
        float s = 0.f;
        for (int c = 0; c < p.channels; ++c) s += p.pcm[(int64_t)c * p.n_in + j];
        p.out[j] = p.channels > 1 ? s / (float)p.channels : s;
        return;
    }
    const int64_t i = j / p.nw;
    const int ph = (int)(j - i * p.nw);
    const int64_t base = i * p.orig - p.width;
    const float* __restrict__ kr = p.kern + (size_t)ph * p.K;
    const int k0 = base < 0 ? (int)(-base) : 0;
    const int k1 = (int)min((int64_t)p.K, p.n_in - base);
    float acc = 0.f;
    if (p.channels == 1) {
        for (int k = k0; k < k1; ++k) acc = fmaf(__ldg(kr + k), __ldg(p.pcm + base + k), acc);
    } else {
        for (int k = k0; k < k1; ++k) {
            float s = 0.f;
            for (int c = 0; c < p.channels; ++c) s += __ldg(p.pcm + (int64_t)c * p.n_in + base + k);
            acc = fmaf(__ldg(kr + k), s / (float)p.channels, acc);
        }
    }
    (void)inv_c;
    p.out[j] = acc;
}

}  // namespace etude
