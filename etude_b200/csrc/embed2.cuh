// Token embedding of the frequency encoder on the tensor cores:  unfold(65) -> Conv2d(1,4,(1,5)) -> Linear(244,256) -> *16 + pos[bin]
// (reference amt_apc.py:79-109), folded at load time into one 65-tap, 1 -> 256 channel filter along time per mel bin:
//     x[(w,f,b), h] = sum_t W16[h][t] * feat_w[f + t][b] + posb[b][h]
//
// GEMM formulation: one tile = 128 consecutive frames of ONE bin; A[i][t] = feat[f0 + i + t][b] is a Hankel matrix the
// producer warps build in shared memory from a staged [192 rows x 32 bins] slab of the feature block.  The fp32 features
// (log-mel, -18 ... +5) keep their precision through a bf16 hi/lo split  x = hi + lo  (|lo| <= 2^-9 |x|):
//     D[128 x 256] = A_hi[128 x 64] W^T + A_lo[128 x 64] W^T,   W = bf16(W16[:, 0:64])      (8 MMAs M128 N256 K16)
// and tap 64 (K = 65 does not fill a 16-wide MMA step) is one fp32 FMA per output in the epilogue together with the
// position / bias row:  y = D + W16[h][64] * feat[f0 + i + 64][b] + posb[b][h]  -> bf16 -> swizzled smem box -> TMA store
// straight into the (window, frame, bin)-major activation tensor (3-D tensor map: h, bin, frame).
// Two TMEM accumulators and two A tiles: building / multiplying tile n + 1 overlaps the epilogue of tile n.
//
// Warps: 0-3 producers (thread = A row), 4-11 epilogue (TMEM lane quarter = warp & 3, channel half = (warp - 4) >> 2: two
// epilogue warps per SM sub-partition hide each other's tcgen05.ld / fence latencies), 12 MMA issue + TMEM allocation.
#pragma once
#include "common.cuh"

namespace etude {

constexpr int kE2Threads = 13 * 32;
constexpr int kE2BinsPerJob = 32;
constexpr int kE2SlabRows = 128 + 64;                 // frames f0 .. f0 + 127 and their 64 rows of right context
constexpr int kE2SlabStride = kE2BinsPerJob + 1;      // padded: thread i reads row i + t, conflict-free across i
constexpr int kE2WBytes = 256 * 64 * 2;               // W (bf16) [256 h][64 taps], K-major SW128
constexpr int kE2ABytes = 2 * 128 * 64 * 2;           // A_hi | A_lo of one tile
constexpr int kE2StageBytes = 8 * 2 * 4096;           // per epilogue warp: its two [32 rows x 64 h] bf16 boxes of one tile
constexpr size_t kEmbed2SmemBytes = 1024 + kE2WBytes + 2 * kE2ABytes + kE2StageBytes + kE2SlabRows * kE2SlabStride * 4 + 4 * 128 * 4 + 4 * 256 * 4 + 256 * 4 + 256;

struct Embed2Params {
    const float* feat;         // padded feature blocks [rows][256]
    const int64_t* win_row0;   // [n_windows] first padded row of every window
    const float* w64;          // [256] W16[h][64]
    const float* posb;         // [256 bins][256 h]
    int n_windows;
    int fblocks;               // 128-frame blocks per window: 4 (512-frame windows) or 1 (128-frame windows of HFT_Transformer)
    int debug_no_store;        // diagnostic bit mask: 1 skip the TMA stores, 2 skip the epilogue arithmetic, 4 skip the A-tile build
};

__device__ __forceinline__ void named_bar_sync(int id, int n_threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory"); }

__global__ void __launch_bounds__(kE2Threads, 1)
embed2_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_out, const Embed2Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sW = smem;
    uint8_t* sA = sW + kE2WBytes;                       // [2][A_hi 16 KB | A_lo 16 KB]
    uint8_t* sStage = sA + 2 * kE2ABytes;
    float* sSlab = reinterpret_cast<float*>(sStage + kE2StageBytes);
    // [4][128] feat[f0 + i + 64][b] of tile n in slot n & 3: slot reuse (tile n + 4) is ordered behind a_empty of tile n + 2,
    // i.e. behind the MMA of tile n + 2, which waited for d_empty of tile n, which the epilogue arrives on after reading xi
    float* sXlast = sSlab + kE2SlabRows * kE2SlabStride;
    float* sPb = sXlast + 4 * 128;                         // [4][256] posb[bin][:] of tile n in slot n & 3 (same ordering argument)
    float* sW64 = sPb + 4 * 256;                           // [256]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sW64 + 256);
    uint64_t* w_full = bars;          // [1]
    uint64_t* a_full = bars + 1;      // [2] producers (4 warps) -> MMA
    uint64_t* a_empty = bars + 3;     // [2] MMA commit -> producers
    uint64_t* d_full = bars + 5;      // [2] MMA commit -> epilogue
    uint64_t* d_empty = bars + 7;     // [2] epilogue (8 warps) -> MMA
    uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(bars + 9);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int n_jobs = p.n_windows * p.fblocks * (kBins / kE2BinsPerJob);   // (window, 128-frame block, 32-bin group)
    const int my_jobs = ((int)blockIdx.x < n_jobs) ? (n_jobs - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int n_tiles = my_jobs * kE2BinsPerJob;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_out);
        mbar_init(w_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_full[i], 4);
            mbar_init(&a_empty[i], 1);
            mbar_init(&d_full[i], 1);
            mbar_init(&d_empty[i], 8);
        }
        mbar_fence_init();
    }
    if (warp == 12) tmem_alloc(tmem_base_ptr, 512);
    for (int i = threadIdx.x; i < 256; i += kE2Threads) sW64[i] = p.w64[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_ptr;

    if (warp < 4) {
        // ===================================================== producers: slab staging + Hankel A tiles (hi / lo split)
        const int i = threadIdx.x;  // A row = frame f0 + i
        for (int jl = 0; jl < my_jobs; ++jl) {
            const int job = blockIdx.x + jl * gridDim.x;
            const int bg = job & 7, fb = (job >> 3) % p.fblocks, w = (job >> 3) / p.fblocks;
            const float* src = p.feat + (p.win_row0[w] + fb * 128) * kBins + bg * kE2BinsPerJob;
            named_bar_sync(1, 128);  // every producer is done with the previous slab
#pragma unroll 4
            for (int idx = i; idx < kE2SlabRows * kE2BinsPerJob; idx += 128) {
                const int r = idx >> 5, c = idx & 31;
                sSlab[r * kE2SlabStride + c] = __ldg(src + (size_t)r * kBins + c);
            }
            named_bar_sync(1, 128);
            for (int bl = 0; bl < kE2BinsPerJob; ++bl) {
                const int n = jl * kE2BinsPerJob + bl;  // tile counter of this CTA
                const int buf = n & 1;
                mbar_wait(&a_empty[buf], ((n >> 1) & 1) ^ 1);
                if (p.debug_no_store & 4) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&a_full[buf]);
                    continue;
                }
                const float2 pbv = __ldg(reinterpret_cast<const float2*>(p.posb + (size_t)(bg * kE2BinsPerJob + bl) * kHid) + i);
                uint8_t* a_hi = sA + buf * kE2ABytes + i * 128;
                uint8_t* a_lo = a_hi + 128 * 128;
                const float* col = sSlab + i * kE2SlabStride + bl;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float x0 = col[(8 * c + 2 * e) * kE2SlabStride], x1 = col[(8 * c + 2 * e + 1) * kE2SlabStride];
                        const __nv_bfloat16 h0 = __float2bfloat16(x0), h1 = __float2bfloat16(x1);
                        hi[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                        lo[e] = pack_bf16x2(x0 - __bfloat162float(h0), x1 - __bfloat162float(h1));
                    }
                    const int pos = (c ^ (i & 7)) << 4;
                    *reinterpret_cast<uint4*>(a_hi + pos) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4*>(a_lo + pos) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
                sXlast[(n & 3) * 128 + i] = col[64 * kE2SlabStride];
                reinterpret_cast<float2*>(sPb + (n & 3) * 256)[i] = pbv;
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_full[buf]);
            }
        }
    } else if (warp < 12) {
        // ===================================================== epilogue: D + tap 64 + position row -> bf16 -> TMA store
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int ch = (warp - 4) >> 2;                 // channel half: boxes 2 ch, 2 ch + 1
        uint8_t* stage = sStage + (warp - 4) * 8192;    // two boxes = 128 channels of this warp's 32 rows
        const bool leader = elect_one();
        for (int n = 0; n < n_tiles; ++n) {
            const int jl = n / kE2BinsPerJob, bl = n % kE2BinsPerJob;
            const int job = blockIdx.x + jl * gridDim.x;
            const int bg = job & 7, fb = (job >> 3) % p.fblocks, w = (job >> 3) / p.fblocks;
            const int bin = bg * kE2BinsPerJob + bl;
            const int buf = n & 1;
            mbar_wait(&d_full[buf], (n >> 1) & 1);
            __syncwarp();
            tc_fence_after();
            if (p.debug_no_store & 2) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&d_empty[buf]);
                continue;
            }
            const float xi = sXlast[(n & 3) * 128 + row];
            const float* pb = sPb + (n & 3) * 256;
            const uint32_t tmem_d = tmem_base + buf * 256 + lane_off;
            if (leader) tma_store_wait_read<0>();   // the previous tile's stores (issued a tile period ago) have read the boxes
            __syncwarp();
            // one fence / one store group per tile: a per-box fence + wait chain left the single epilogue warp of a
            // sub-partition latency-bound (~2500 clk per box)
#pragma unroll 1
            for (int hc = 2 * ch; hc < 2 * ch + 2; ++hc) {   // 64 output channels per box
                uint8_t* box = stage + (hc & 1) * 4096;
                float v[64];
                tmem_ld32(tmem_d + hc * 64, v);
                tmem_ld32(tmem_d + hc * 64 + 32, v + 32);
                tc_wait_ld();
#pragma unroll
                for (int chunk = 0; chunk < 8; ++chunk) {   // 8 channels -> one 16-byte chunk
                    const int h0 = hc * 64 + chunk * 8;
                    const float4 w0 = *reinterpret_cast<const float4*>(sW64 + h0), w1 = *reinterpret_cast<const float4*>(sW64 + h0 + 4);
                    const float4 p0 = *reinterpret_cast<const float4*>(pb + h0), p1 = *reinterpret_cast<const float4*>(pb + h0 + 4);
                    uint4 pk;
                    pk.x = pack_bf16x2(v[chunk * 8 + 0] + fmaf(w0.x, xi, p0.x), v[chunk * 8 + 1] + fmaf(w0.y, xi, p0.y));
                    pk.y = pack_bf16x2(v[chunk * 8 + 2] + fmaf(w0.z, xi, p0.z), v[chunk * 8 + 3] + fmaf(w0.w, xi, p0.w));
                    pk.z = pack_bf16x2(v[chunk * 8 + 4] + fmaf(w1.x, xi, p1.x), v[chunk * 8 + 5] + fmaf(w1.y, xi, p1.y));
                    pk.w = pack_bf16x2(v[chunk * 8 + 6] + fmaf(w1.z, xi, p1.z), v[chunk * 8 + 7] + fmaf(w1.w, xi, p1.w));
                    *reinterpret_cast<uint4*>(box + lane * 128 + ((chunk ^ (lane & 7)) << 4)) = pk;
                }
            }
            tc_fence_before();   // the accumulator has been read: hand the TMEM buffer back to the MMA warp
            __syncwarp();
            if (lane == 0) mbar_arrive(&d_empty[buf]);
            fence_async_smem();
            __syncwarp();
            if (leader && !(p.debug_no_store & 1)) {
#pragma unroll
                for (int k = 0; k < 2; ++k) tma_store_3d(&tmap_out, stage + k * 4096, (2 * ch + k) * 64, bin, (w * p.fblocks + fb) * 128 + q * 32);
                tma_store_commit();
            }
        }
        if (leader) tma_store_wait_all<0>();
    } else {
        // ===================================================== MMA issuer (warp-uniform loop, one elected lane issues)
        const bool leader = elect_one();
        if (leader) {
            mbar_expect_tx(w_full, kE2WBytes);
            tma_load_2d(sW, &tmap_w, w_full, 0, 0);
        }
        __syncwarp();
        mbar_wait(w_full, 0);
        const uint32_t idesc = make_idesc_bf16(128, 256, 0, 0);
        const uint64_t w_desc = make_sw128_desc(smem_u32(sW));
        for (int n = 0; n < n_tiles; ++n) {
            const int buf = n & 1;
            mbar_wait(&a_full[buf], (n >> 1) & 1);
            mbar_wait(&d_empty[buf], ((n >> 1) & 1) ^ 1);
            tc_fence_after();
            if (leader) {
                const uint64_t a_desc = make_sw128_desc(smem_u32(sA + buf * kE2ABytes));
                const uint32_t tmem_d = tmem_base + buf * 256;
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_d, a_desc + 2 * k, w_desc + 2 * k, idesc, k != 0);              // A_hi
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_d, a_desc + 1024 + 2 * k, w_desc + 2 * k, idesc, 1u);            // A_lo (+16 KB)
                tc_commit(&a_empty[buf]);
                tc_commit(&d_full[buf]);
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 12) tmem_dealloc(tmem_base, 512);
}

}  // namespace etude
