// Fused log-mel front-end: framing (center=True, reflect pad) + periodic Hann + 2048-point real DFT + |.|^2 +
// sparse slaney/htk mel filterbank + log(x + 1e-8), written straight into the -18-padded per-song feature
// buffer that the window loop reads.  No [1025, T] spectrum ever reaches HBM.
//
// Replaces torchaudio MelSpectrogram -> log -> .T (reference etude/data/extractor.py:186-197) and the numpy
// padding of _transcript (extractor.py:210-213).
//
// The 2048-point real DFT is a 1024-point complex Stockham radix-4 FFT (5 stages, one butterfly per thread per
// stage, fp32, table twiddles computed in double) followed by the even/odd split.
#pragma once
#include "common.cuh"

namespace etude {

constexpr int kNfft = 2048;
constexpr int kHop = 256;
constexpr int kFreqs = 1025;
constexpr int kMelMaxNnz = 2304;  // packed filter weights (2036 used with the default config)
constexpr int kLogmelThreads = 256;
constexpr int kLogmelRowsPerCta = 8;

struct LogmelTables {
    const float2* tw1024;   // exp(-2 pi i n / 1024), n < 1024
    const float2* tw2048;   // exp(-2 pi i k / 2048), k <= 1024
    const float* window;    // periodic Hann, 2048
    const int* mel_start;   // [256] first FFT bin of filter m
    const int* mel_count;   // [256] number of bins
    const int* mel_offset;  // [256] offset into mel_weight
    const float* mel_weight;
};

struct LogmelSong {
    int64_t wave_off;  // first sample of this song in the wave buffer
    int64_t n_samples;
    int64_t row_off;   // first row of this song's padded feature block
    int64_t n_rows;    // 32 + T_pad + 32
    int64_t n_frames;  // T = 1 + n_samples / 256
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

__global__ void __launch_bounds__(kLogmelThreads)
logmel_kernel(const float* __restrict__ wave, const LogmelSong* __restrict__ songs, LogmelTables tab, float* __restrict__ feat,
              float min_value, float log_offset) {
    __shared__ float2 buf0[1024];
    __shared__ float2 buf1[1024];
    __shared__ float power[kFreqs + 3];
    __shared__ float2 s_tw[1024];

    const LogmelSong song = songs[blockIdx.y];
    const int64_t r_begin = (int64_t)blockIdx.x * kLogmelRowsPerCta;
    if (r_begin >= song.n_rows) return;
    const int tid = threadIdx.x;
    for (int i = tid; i < 1024; i += kLogmelThreads) s_tw[i] = tab.tw1024[i];
    const int m_start = tab.mel_start[tid], m_count = tab.mel_count[tid], m_off = tab.mel_offset[tid];
    const float* __restrict__ x = wave + song.wave_off;
    const int64_t n = song.n_samples;

    for (int rr = 0; rr < kLogmelRowsPerCta; ++rr) {
        const int64_t r = r_begin + rr;
        if (r >= song.n_rows) break;
        float* out = feat + (song.row_off + r) * kBins;
        const int64_t t = r - kMargin;
        if (t < 0 || t >= song.n_frames) {  // the -18 rows of _transcript's padding
            out[tid] = min_value;
            continue;
        }
        // windowed frame -> packed complex z[i] = (x[2i] w[2i], x[2i+1] w[2i+1])
        const int64_t base = t * kHop - kNfft / 2;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = tid + q * kLogmelThreads;
            int64_t j0 = base + 2 * i, j1 = j0 + 1;
            if (j0 < 0) j0 = -j0;
            if (j0 >= n) j0 = 2 * (n - 1) - j0;
            if (j1 < 0) j1 = -j1;
            if (j1 >= n) j1 = 2 * (n - 1) - j1;
            buf0[i] = make_float2(x[j0] * tab.window[2 * i], x[j1] * tab.window[2 * i + 1]);
        }
        __syncthreads();
        // 5 Stockham radix-4 stages, ping-pong buf0 -> buf1 -> ... -> result in buf1
        float2* src = buf0;
        float2* dst = buf1;
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            const int ns = 1 << (2 * s);
            const int k = tid & (ns - 1);
            const int tstep = 256 >> (2 * s);  // N / (4 ns)
            float2 v0 = src[tid], v1 = src[tid + 256], v2 = src[tid + 512], v3 = src[tid + 768];
            if (s > 0) {
                v1 = cmul(v1, s_tw[k * tstep]);
                v2 = cmul(v2, s_tw[2 * k * tstep]);
                v3 = cmul(v3, s_tw[3 * k * tstep]);
            }
            const float2 t0 = make_float2(v0.x + v2.x, v0.y + v2.y);
            const float2 t1 = make_float2(v0.x - v2.x, v0.y - v2.y);
            const float2 t2 = make_float2(v1.x + v3.x, v1.y + v3.y);
            const float2 t3 = make_float2(v1.y - v3.y, -(v1.x - v3.x));  // (v1 - v3) * (-i)
            const int j0 = ((tid >> (2 * s)) << (2 * s + 2)) + k;
            dst[j0] = make_float2(t0.x + t2.x, t0.y + t2.y);
            dst[j0 + ns] = make_float2(t1.x + t3.x, t1.y + t3.y);
            dst[j0 + 2 * ns] = make_float2(t0.x - t2.x, t0.y - t2.y);
            dst[j0 + 3 * ns] = make_float2(t1.x - t3.x, t1.y - t3.y);
            __syncthreads();
            float2* tmp = src;
            src = dst;
            dst = tmp;
        }
        // src now holds Z (natural order).  X[k] = E[k] + w^k O[k], k = 0..1024
        for (int k = tid; k <= 1024; k += kLogmelThreads) {
            const float2 zk = src[k & 1023];
            const float2 zc = src[(1024 - k) & 1023];  // conj taken below
            const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y));
            const float2 o = make_float2(0.5f * (zk.y + zc.y), -0.5f * (zk.x - zc.x));  // (zk - conj(zc)) / (2i)
            const float2 wo = cmul(tab.tw2048[k], o);
            const float re = e.x + wo.x, im = e.y + wo.y;
            power[k] = re * re + im * im;
        }
        __syncthreads();
        float acc = 0.f;
        for (int i = 0; i < m_count; ++i) acc += tab.mel_weight[m_off + i] * power[m_start + i];
        out[tid] = logf(acc + log_offset);
        __syncthreads();
    }
}

}  // namespace etude
