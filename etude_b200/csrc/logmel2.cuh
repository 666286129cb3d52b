// Fused log-mel front-end (second generation; the first, one CTA per frame with five smem passes, has been removed):
// framing (center=True, reflect pad) + periodic Hann + 2048-point real DFT + |.|^2 + sparse slaney/htk mel filterbank +
// log(x + 1e-8), written straight into the -18-padded per-song feature block (reference etude/data/extractor.py:186-197,
// 210-213).  Same arithmetic contract as logmel.cuh (fp32, table twiddles computed in double); the mel sums use four partial sums.
//
// What changed (profiles/r1c_ncu_front_details.txt): the first kernel spent a CTA of 256 threads on one frame at a time,
// five radix-4 passes through shared memory with a block barrier after each.  Here ONE WARP owns a frame: the 1024-point
// complex FFT of the packed frame is 32 x 32 -- a radix-32 DFT entirely in registers (each lane holds 32 points), one
// twiddle + transpose through a padded smem tile, a second radix-32 DFT in registers -- so a frame costs two smem round
// trips and only __syncwarp.  12 warps per CTA walk neighbouring frames (their 87.5 % overlapping samples hit L1: DRAM reads
// the samples once).  Staging the CTA's span with a bulk copy instead was built and measured: see ETUDE_LOGMEL_STAGED below.
#pragma once
#include "logmel.cuh"

namespace etude {

// ETUDE_LOGMEL_STAGED = 0 (shipped): two CTAs of 12 warps per SM reading their frames through L1;
// 1: one CTA of 16 warps per SM, the CTA's sample span staged in smem by one bulk copy (the TMA engine).  The kernel is
// issue-bound (DRAM 1 % busy), so what decides between them is resident warps, not memory traffic: the per-warp 8.4 KB FFT
// tile leaves room for the 73 KB span only at 16 warps per SM.  Measured on the 10 h microbench (profiles/r2j_frontend_*.json):
// staged 314.7 GB/s, unstaged 375.0 GB/s -- so the staged build is a measured experiment, not the product.
#ifndef ETUDE_LOGMEL_STAGED
#define ETUDE_LOGMEL_STAGED 0
#endif
constexpr bool kL2Staged = ETUDE_LOGMEL_STAGED != 0;
constexpr int kL2Warps = kL2Staged ? 16 : 12;
constexpr int kL2Threads = kL2Warps * 32;
constexpr int kL2RowsPerWarp = kL2Staged ? 4 : 8;
constexpr int kL2RowsPerCta = kL2Warps * kL2RowsPerWarp;   // 64 (staged) / 96 consecutive rows of the feature block
constexpr int kL2BufFloat2 = 32 * 33;   // transpose tile, row stride 33 (also holds Z in natural order, then the power spectrum)
// The samples of the CTA's 64 frames overlap by 87.5 %: their union, (64 - 1) * 256 + 2048 = 18 176 contiguous samples, is
// staged ONCE into smem with one bulk copy (cp.async.bulk global -> shared, completion on an mbarrier) and every frame is
// read from there.  Frames that reach past the ends of the song (reflect / constant padding) take the per-sample path.
constexpr int kL2SpanFloats = kL2Staged ? (kL2RowsPerCta - 1) * kHop + kNfft : 4;
constexpr size_t kLogmel2SmemBytes = (size_t)kL2Warps * kL2BufFloat2 * 8 + 32 * 32 * 8 + (size_t)kL2SpanFloats * 4 + 16;

// W_32^m = cos(2 pi m / 32) - i sin(2 pi m / 32), m < 16
__device__ __forceinline__ float2 tw32_mul(float2 d, int m) {
    constexpr float C[16] = {1.000000000e+00f, 9.807852804e-01f, 9.238795325e-01f, 8.314696123e-01f, 7.071067812e-01f, 5.555702330e-01f,
                             3.826834324e-01f, 1.950903220e-01f, 0.f, -1.950903220e-01f, -3.826834324e-01f, -5.555702330e-01f,
                             -7.071067812e-01f, -8.314696123e-01f, -9.238795325e-01f, -9.807852804e-01f};
    constexpr float S[16] = {0.f, 1.950903220e-01f, 3.826834324e-01f, 5.555702330e-01f, 7.071067812e-01f, 8.314696123e-01f,
                             9.238795325e-01f, 9.807852804e-01f, 1.000000000e+00f, 9.807852804e-01f, 9.238795325e-01f, 8.314696123e-01f,
                             7.071067812e-01f, 5.555702330e-01f, 3.826834324e-01f, 1.950903220e-01f};
    if (m == 0) return d;
    if (m == 8) return make_float2(d.y, -d.x);   // * (-i)
    const float c = C[m], s = S[m];
    return make_float2(fmaf(d.x, c, d.y * s), fmaf(d.y, c, -d.x * s));
}

// In-register 32-point DFT, decimation in frequency: natural order in, a[bitrev5(k)] = A[k] out.  Fully unrolled: every
// index and twiddle is a compile-time constant.
__device__ __forceinline__ void fft32(float2 (&a)[32]) {
#pragma unroll
    for (int len = 32; len >= 2; len >>= 1) {
        const int half = len >> 1, step = 32 / len;
#pragma unroll
        for (int blk = 0; blk < 32; blk += len) {
#pragma unroll
            for (int j = 0; j < half; ++j) {
                const float2 u = a[blk + j], v = a[blk + j + half];
                a[blk + j] = make_float2(u.x + v.x, u.y + v.y);
                a[blk + j + half] = tw32_mul(make_float2(u.x - v.x, u.y - v.y), j * step);
            }
        }
    }
}
__host__ __device__ constexpr int bitrev5(int k) { return ((k & 1) << 4) | ((k & 2) << 2) | (k & 4) | ((k & 8) >> 2) | ((k & 16) >> 4); }

__global__ void __launch_bounds__(kL2Threads, kL2Staged ? 1 : 2)
logmel2_kernel(const float* __restrict__ wave, const LogmelSong* __restrict__ songs, LogmelTables tab, const float2* __restrict__ tw32x32,
               float* __restrict__ feat, float min_value, float log_offset, int reflect) {
    extern __shared__ float2 l2_smem[];
    float2* s_tw = l2_smem;                                           // [32 k1][32 n2] = W_1024^(n2 k1)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2* buf = l2_smem + 32 * 32 + warp * kL2BufFloat2;

    float* s_span = reinterpret_cast<float*>(l2_smem + 32 * 32 + kL2Warps * kL2BufFloat2);
    uint64_t* span_bar = reinterpret_cast<uint64_t*>(s_span + kL2SpanFloats);

    const LogmelSong song = songs[blockIdx.y];
    const int64_t r_cta = (int64_t)blockIdx.x * kL2RowsPerCta;
    if (r_cta >= song.n_rows) return;
    const float* __restrict__ x = wave + song.wave_off;
    const int64_t n = song.n_samples;
    // ---- stage the contiguous sample span of this CTA's frames: samples [span0, span1) of the song, clipped to the song and
    // to 16-byte alignment of the global address (bulk copies move multiples of 16 B between 16-B aligned addresses)
    const int64_t t_first = r_cta - song.front_rows;                     // frame of the CTA's first row (may be negative: pad rows)
    int64_t span0 = t_first * kHop - kNfft / 2, span1 = (t_first + kL2RowsPerCta - 1) * kHop + kNfft / 2;
    if (span0 < 0) span0 = 0;
    if (span1 > n) span1 = n;
    {
        const int64_t mis = ((reinterpret_cast<uintptr_t>(x + span0) & 15) / 4);   // floats past the previous 16-B boundary
        if (mis) span0 += 4 - mis;
        span1 = span0 + ((span1 - span0) & ~int64_t(3));
    }
    const bool staged = kL2Staged && span1 > span0;
    if (threadIdx.x == 0) {
        mbar_init(span_bar, 1);
        mbar_fence_init();
        if (staged) {
            const uint32_t bytes = (uint32_t)(span1 - span0) * 4;
            mbar_expect_tx(span_bar, bytes);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(s_span)),
                         "l"(reinterpret_cast<uint64_t>(x + span0)), "r"(bytes), "r"(smem_u32(span_bar))
                         : "memory");
        }
    }
    for (int i = threadIdx.x; i < 32 * 32; i += kL2Threads) s_tw[i] = tw32x32[i];
    __syncthreads();
    if (staged) mbar_wait(span_bar, 0);
    const float2* __restrict__ win2 = reinterpret_cast<const float2*>(tab.window);

    // neighbouring frames go to neighbouring warps: row = r_cta + it * 12 + warp
    for (int it = 0; it < kL2RowsPerWarp; ++it) {
        const int64_t r = r_cta + (int64_t)it * kL2Warps + warp;
        if (r >= song.n_rows) break;
        float* out = feat + (song.row_off + r) * kBins;
        const int64_t t = r - song.front_rows;
        if (t < 0 || t >= song.n_frames) {  // the min_value rows of _transcript's padding
#pragma unroll
            for (int i = 0; i < 8; ++i) out[lane + 32 * i] = min_value;
            continue;
        }
        // ---- windowed frame, packed complex z[m] = (x[2m] w[2m], x[2m+1] w[2m+1]); lane n2 holds z[32 n1 + n2], n1 = 0..31
        float2 a[32];
        const int64_t base = t * kHop - kNfft / 2;
        if (staged && base >= span0 && base + kNfft <= span1) {   // the whole frame lies in the staged span
            const float* __restrict__ xs = s_span + (base - span0);
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) {
                const int m = 32 * n1 + lane;
                const float2 wv = __ldg(win2 + m);
                a[n1] = make_float2(xs[2 * m] * wv.x, xs[2 * m + 1] * wv.y);   // 8-byte stride across lanes: conflict-free
            }
        } else if (!kL2Staged && ((reinterpret_cast<uintptr_t>(x) & 7) == 0) && base >= 0 && base + kNfft <= n) {
            const float2* __restrict__ x2 = reinterpret_cast<const float2*>(x + base);
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) {
                const float2 xv = __ldg(x2 + 32 * n1 + lane), wv = __ldg(win2 + 32 * n1 + lane);
                a[n1] = make_float2(xv.x * wv.x, xv.y * wv.y);
            }
        } else {
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) {
                const int m = 32 * n1 + lane;
                int64_t j0 = base + 2 * m, j1 = j0 + 1;
                const bool in0 = j0 >= 0 && j0 < n, in1 = j1 >= 0 && j1 < n;
                if (j0 < 0) j0 = -j0;                    // center=True, pad_mode="reflect" (extractor.py:186-193: torchaudio default)
                if (j0 >= n) j0 = 2 * (n - 1) - j0;
                if (j1 < 0) j1 = -j1;
                if (j1 >= n) j1 = 2 * (n - 1) - j1;
                const float2 wv = __ldg(win2 + m);
                // pad_mode="constant" (hft_transformer.py:131): samples outside the wave are zero
                const float x0 = (reflect || in0) ? x[j0] : 0.f, x1 = (reflect || in1) ? x[j1] : 0.f;
                a[n1] = make_float2(x0 * wv.x, x1 * wv.y);
            }
        }
        // ---- stage 1: DFT over n1 (stride 32), twiddle W_1024^(n2 k1), transpose through smem
        fft32(a);
#pragma unroll
        for (int pos = 0; pos < 32; ++pos) {
            const int k1 = bitrev5(pos);
            float2 v = a[pos];
            if (k1 != 0) v = cmul(v, s_tw[k1 * 32 + lane]);
            buf[k1 * 33 + lane] = v;
        }
        __syncwarp();
        // ---- stage 2: lane = k1 gathers its 32 values over n2, DFT over n2 -> Z[k1 + 32 k2]
#pragma unroll
        for (int n2 = 0; n2 < 32; ++n2) a[n2] = buf[lane * 33 + n2];
        __syncwarp();
        fft32(a);
#pragma unroll
        for (int pos = 0; pos < 32; ++pos) buf[lane + 32 * bitrev5(pos)] = a[pos];   // Z in natural order
        __syncwarp();
        // ---- real split: X[k] = E[k] + w^k O[k], X[1024 - k] = conj(E[k] - w^k O[k]); powers written in place over Z
#pragma unroll 8
        for (int m = 0; m < 16; ++m) {
            const int k = lane + 32 * m;   // 0 .. 511
            if (k == 0) {
                const float2 z0 = buf[0];
                const float p0 = (z0.x + z0.y) * (z0.x + z0.y), pn = (z0.x - z0.y) * (z0.x - z0.y);
                const float2 zm = buf[512];
                buf[0].x = p0;
                buf[1024].x = pn;                       // slot 1024 lies in the row padding of the tile
                buf[512].x = zm.x * zm.x + zm.y * zm.y;  // X[512] = conj(Z[512])
            } else {
                const float2 zk = buf[k], zc = buf[1024 - k];
                const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y));
                const float2 o = make_float2(0.5f * (zk.y + zc.y), -0.5f * (zk.x - zc.x));  // (zk - conj(zc)) / (2i)
                const float2 wo = cmul(__ldg(tab.tw2048 + k), o);
                const float re0 = e.x + wo.x, im0 = e.y + wo.y, re1 = e.x - wo.x, im1 = e.y - wo.y;
                buf[k].x = re0 * re0 + im0 * im0;
                buf[1024 - k].x = re1 * re1 + im1 * im1;
            }
        }
        __syncwarp();
        // ---- mel filterbank (lane owns filters lane, lane + 32, ...: neighbouring filters have similar widths) + log
#pragma unroll 2
        for (int i = 0; i < 8; ++i) {
            const int mel = lane + 32 * i;
            const int m_start = __ldg(tab.mel_start + mel), m_count = __ldg(tab.mel_count + mel), m_off = __ldg(tab.mel_offset + mel);
            // four independent partial sums (loads of a group in flight together); filters are 1 .. 22 bins wide
            float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
            int j = 0;
            for (; j + 4 <= m_count; j += 4) {
                const float w0 = __ldg(tab.mel_weight + m_off + j), w1 = __ldg(tab.mel_weight + m_off + j + 1);
                const float w2 = __ldg(tab.mel_weight + m_off + j + 2), w3 = __ldg(tab.mel_weight + m_off + j + 3);
                acc0 = fmaf(w0, buf[m_start + j].x, acc0);
                acc1 = fmaf(w1, buf[m_start + j + 1].x, acc1);
                acc2 = fmaf(w2, buf[m_start + j + 2].x, acc2);
                acc3 = fmaf(w3, buf[m_start + j + 3].x, acc3);
            }
            for (; j < m_count; ++j) acc0 = fmaf(__ldg(tab.mel_weight + m_off + j), buf[m_start + j].x, acc0);
            out[mel] = logf((acc0 + acc1) + (acc2 + acc3) + log_offset);
        }
        __syncwarp();
    }
}

}  // namespace etude
