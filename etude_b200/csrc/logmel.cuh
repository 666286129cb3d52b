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
    int64_t n_rows;    // rows of the block: front_rows + T + the trailing pad rows (32 + T_pad + 32 for _transcript)
    int64_t n_frames;  // T = 1 + n_samples / 256
    int64_t front_rows;  // pad rows before frame 0 (32 = margin_b for _transcript; 64 = margin_b + n_offset for _transcript_stride)
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

}  // namespace etude
