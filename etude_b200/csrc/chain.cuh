// Fused token-local chain of one transformer sub-block on tcgen05 (A-stationary, weights streamed):
//
//     y   = LayerNorm(ctx @ Wo^T + bo + x)                 attention output projection + residual + LN
//     out = LayerNorm(y + relu(y @ W1^T + b1) @ W2^T + b2) position-wise FFN + residual + LN (same LN module)
//
// i.e. everything of EncoderLayer / DecoderLayer(_Zero).forward that follows the attention core (reference
// amt_apc.py:250-258, 276-284, 310-318 with fc_o of amt_apc.py:371 and the FFN of 383-392).  One persistent CTA per
// SM walks 128-token tiles; per tile nothing but ctx (bf16), the residual x (bf16) and out (bf16) touch HBM: the
// pre-norm rows, y (fp32) and the FFN accumulator live in TMEM, y (bf16) and the ReLU'd hidden chunk live in smem
// as MMA A operands.  FFN2 accumulates ON TOP of (y + b2) parked in TMEM, so the second residual add is free and
// is exact fp32.
//
// TMEM (512 columns) = two 256-column regions R0/R1 whose roles swap every tile (parity p = tile & 1):
//     D1   = R[p]   : ctx Wo^T accumulator -> pre-norm row (parked) -> y + b2 -> + FFN2 accumulation -> LN2 input
//     ACC2 = R[p^1] : two 128-column FFN1 chunk accumulators (chunk j -> half j & 1)
// so the next tile's first GEMM runs while this tile's LN2 epilogue drains D1.
//
// Warps (11): 0 = TMA weight/ctx ring producer, 1 = MMA issuer + TMEM allocator, 2..9 = epilogue (two warps per TMEM
// lane quarter, each thread owns half of the columns of one row), 10 = TMA residual producer.
// smem: ring 6 x 16 KB ([128 x 64] bf16 boxes, SW128) | Y 64 KB (residual in -> y bf16 in place, FFN1 A operand)
//       | H 32 KB (hidden chunk, FFN2 A operand; LN pair statistics alias it) | out staging 2 x 16 KB.
#pragma once
#include "common.cuh"

namespace etude {

constexpr int kChainThreads = 11 * 32;
constexpr int kChStages = 6;
constexpr int kChStageBytes = 128 * 64 * 2;  // 16 KB
constexpr int kChYBytes = 4 * kChStageBytes;
constexpr int kChHBytes = 2 * kChStageBytes;
constexpr int kChOutBytes = 2 * kChStageBytes;
constexpr size_t kChainSmemBytes = 1024 + kChStages * kChStageBytes + kChYBytes + kChHBytes + kChOutBytes + 256;

struct ChainParams {
    int M;
    int num_tiles;
    int resid_mod;  // 0: residual row == row;  >0: residual row == row % resid_mod (wrapped bf16 table, see api.cu)
    const float* bo;     // [256]
    const float* b1;     // [512]
    const float* b2;     // [256]
    const float* gamma;  // [256]
    const float* beta;   // [256]
    long long* trace;    // debug timeline (clock64 stamps of CTA 0), or nullptr
};

// Debug timeline: role r (0 MMA thread, 1 epilogue warp 2 lane 0, 2 ring producer) appends (event id, clock64) pairs.
constexpr int kChTraceSlots = 512;
#define CH_TRACE(role, id)                                                                         \
    do {                                                                                           \
        if (p.trace != nullptr && blockIdx.x == 0 && tr_n < kChTraceSlots) {                        \
            p.trace[((role) * kChTraceSlots + tr_n) * 2] = (id);                                    \
            p.trace[((role) * kChTraceSlots + tr_n) * 2 + 1] = clock64();                           \
            ++tr_n;                                                                                \
        }                                                                                          \
    } while (0)

__device__ __forceinline__ void unpack_bf16x8(const uint4& u, float* f) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
}

// FFN = true : the whole chain above.   FFN = false : y only (out = y), the "fc_o + residual + LN" step of a decoder
// layer's self-attention, which is followed by the cross-attention rather than by the FFN (amt_apc.py:304).
template <bool FFN>
__global__ void __launch_bounds__(kChainThreads, 1)
chain_kernel(const __grid_constant__ CUtensorMap tmap_ctx, const __grid_constant__ CUtensorMap tmap_wo,
             const __grid_constant__ CUtensorMap tmap_w1, const __grid_constant__ CUtensorMap tmap_w2,
             const __grid_constant__ CUtensorMap tmap_resid, const __grid_constant__ CUtensorMap tmap_out, const ChainParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sRing = smem;
    uint8_t* sY = sRing + kChStages * kChStageBytes;
    uint8_t* sH = sY + kChYBytes;
    uint8_t* sOut = sH + kChHBytes;
    float2* s_stat = reinterpret_cast<float2*>(sH);  // [2 halves][128 rows] (mean, M2): aliases H, which is idle during LN
    uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + kChOutBytes);
    uint64_t* full = bars;               // [6] ring: TMA -> MMA
    uint64_t* empty = full + kChStages;  // [6] ring: MMA -> TMA
    uint64_t* y_full = empty + kChStages;  // residual tile landed in Y            (TMA -> epilogue)
    uint64_t* y_free = y_full + 1;         // FFN1 finished reading Y              (MMA commit -> residual producer)
    uint64_t* g1_full = y_free + 1;        // [2] ctx Wo^T accumulator complete    (MMA commit -> epilogue), per region
    uint64_t* e1_done = g1_full + 2;       // y in smem, y + b2 in TMEM            (8 epilogue warps -> MMA)
    uint64_t* f1_full = e1_done + 1;       // [2] FFN1 chunk accumulator complete  (MMA commit -> epilogue)
    uint64_t* h_full = f1_full + 2;        // hidden chunk in smem                 (8 epilogue warps -> MMA)
    uint64_t* h_free = h_full + 1;         // FFN2 partial finished reading H      (MMA commit -> epilogue)
    uint64_t* f2_full = h_free + 1;        // FFN2 accumulation complete           (MMA commit -> epilogue)
    uint64_t* qfree = f2_full + 1;         // [4] 128-column TMEM quarter drained  (8 epilogue warps -> MMA)
    uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(qfree + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int my_tiles = ((int)blockIdx.x < p.num_tiles) ? (p.num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_ctx); tma_prefetch_desc(&tmap_wo); tma_prefetch_desc(&tmap_w1);
        tma_prefetch_desc(&tmap_w2); tma_prefetch_desc(&tmap_resid); tma_prefetch_desc(&tmap_out);
        for (int s = 0; s < kChStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(y_full, 1); mbar_init(y_free, FFN ? 1 : 8); mbar_init(&g1_full[0], 1); mbar_init(&g1_full[1], 1); mbar_init(e1_done, 8);
        mbar_init(&f1_full[0], 1); mbar_init(&f1_full[1], 1);
        mbar_init(h_full, 8); mbar_init(h_free, 1); mbar_init(f2_full, 1);
        for (int q = 0; q < 4; ++q) mbar_init(&qfree[q], 8);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_base_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_ptr;

    if (warp == 0) {
        // ===================================================== ring producer: ctx + weight boxes in consumption order
        if (lane == 0) {
            uint32_t c = 0;
            int tr_n = 0;
            auto load = [&](const CUtensorMap* m, int col, int row) {
                const uint32_t s = c % kChStages;
                mbar_wait(&empty[s], ((c / kChStages) & 1) ^ 1);
                mbar_expect_tx(&full[s], kChStageBytes);
                tma_load_2d(sRing + s * kChStageBytes, m, &full[s], col, row);
                ++c;
            };
            auto load_w1 = [&](int j) { for (int kb = 0; kb < 4; ++kb) load(&tmap_w1, kb * 64, j * 128); };
            auto load_w2 = [&](int j) {
                for (int kk = 0; kk < 2; ++kk) { load(&tmap_w2, j * 128 + kk * 64, 0); load(&tmap_w2, j * 128 + kk * 64, 128); }
            };
            for (int it = 0; it < my_tiles; ++it) {
                const int row0 = (blockIdx.x + it * gridDim.x) * 128;
                CH_TRACE(2, it * 100);
                for (int kb = 0; kb < 4; ++kb) {  // first GEMM: A = ctx k-block, B = Wo k-block (two 128-row halves)
                    load(&tmap_ctx, kb * 64, row0);
                    load(&tmap_wo, kb * 64, 0);
                    load(&tmap_wo, kb * 64, 128);
                }
                // FFN, in the MMA thread's software-pipelined order F1(0) F1(1) F2(0) F1(2) F2(1) F1(3) F2(2) F2(3)
                CH_TRACE(2, it * 100 + 1);
                if constexpr (FFN) { load_w1(0); load_w1(1); load_w2(0); load_w1(2); load_w2(1); load_w1(3); load_w2(2); load_w2(3); }
                CH_TRACE(2, it * 100 + 2);
            }
        }
    } else if (warp == 10) {
        // ===================================================== residual producer: x tile -> Y (4 boxes [128 x 64])
        if (lane == 0) {
            for (int it = 0; it < my_tiles; ++it) {
                int row0 = (blockIdx.x + it * gridDim.x) * 128;
                if (p.resid_mod) row0 %= p.resid_mod;
                mbar_wait(y_free, (it & 1) ^ 1);
                mbar_expect_tx(y_full, kChYBytes);
                for (int kb = 0; kb < 4; ++kb) tma_load_2d(sY + kb * kChStageBytes, &tmap_resid, y_full, kb * 64, row0);
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer (one thread)
        if (lane == 0) {
            constexpr uint32_t idesc128 = make_idesc_bf16(128, 128, 0, 0);
            uint32_t c = 0;            // ring consumption counter
            uint32_t prod[4] = {0, 0, 0, 0};  // productions into each TMEM quarter so far
            uint32_t n_h = 0;          // hidden chunks consumed so far (h_full phase)
            int tr_n = 0;
            auto acquire = [&]() -> uint32_t {
                const uint32_t s = c % kChStages;
                mbar_wait(&full[s], (c / kChStages) & 1);
                ++c;
                return s;
            };
            auto wait_quarter = [&](int q) { mbar_wait(&qfree[q], (prod[q] & 1) ^ 1); ++prod[q]; };
            // D[128 x 128] (+)= A[128 x 64] B[128 x 64]^T : one smem k-block = 4 MMAs of K 16
            auto mma_kblock = [&](uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, bool acc_first) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16_ss(tmem_d, make_sw128_desc(a_addr + k * 32), make_sw128_desc(b_addr + k * 32), idesc128, (acc_first || k) ? 1u : 0u);
            };
            auto ring_addr = [&](uint32_t s) { return smem_u32(sRing + s * kChStageBytes); };
            const uint32_t y_addr = smem_u32(sY), h_addr = smem_u32(sH);

            for (int it = 0; it < my_tiles; ++it) {
                const int par = it & 1;
                const uint32_t d1 = tmem_base + par * 256;         // region R[par]
                const uint32_t a2 = tmem_base + (par ^ 1) * 256;   // region R[par ^ 1]
                const int qd = par * 2, qa = (par ^ 1) * 2;        // first quarter index of each region
                // ---- G1: D1 = ctx Wo^T
                CH_TRACE(0, it * 100);
                wait_quarter(qd);
                wait_quarter(qd + 1);
                tc_fence_after();
                CH_TRACE(0, it * 100 + 1);
                for (int kb = 0; kb < 4; ++kb) {
                    const uint32_t sa = acquire(), sb0 = acquire(), sb1 = acquire();
                    tc_fence_after();
                    mma_kblock(d1, ring_addr(sa), ring_addr(sb0), kb != 0);
                    mma_kblock(d1 + 128, ring_addr(sa), ring_addr(sb1), kb != 0);
                    tc_commit(&empty[sa]); tc_commit(&empty[sb0]); tc_commit(&empty[sb1]);
                }
                tc_commit(&g1_full[par]);
                CH_TRACE(0, it * 100 + 2);
                if constexpr (!FFN) continue;
                // ---- FFN, software pipelined
                auto f1 = [&](int j) {  // ACC2[j & 1] = y W1_j^T
                    wait_quarter(qa + (j & 1));
                    tc_fence_after();
                    CH_TRACE(0, it * 100 + 10 + j);
                    for (int kb = 0; kb < 4; ++kb) {
                        const uint32_t s = acquire();
                        tc_fence_after();
                        mma_kblock(a2 + (j & 1) * 128, y_addr + kb * kChStageBytes, ring_addr(s), kb != 0);
                        tc_commit(&empty[s]);
                    }
                    tc_commit(&f1_full[j & 1]);
                    CH_TRACE(0, it * 100 + 20 + j);
                };
                auto f2 = [&](int j) {  // D1 += h_j W2[:, 128 j ..]^T   (D1 already holds y + b2)
                    mbar_wait(h_full, n_h & 1);
                    ++n_h;
                    tc_fence_after();
                    CH_TRACE(0, it * 100 + 30 + j);
                    for (int kk = 0; kk < 2; ++kk) {
                        const uint32_t s0 = acquire(), s1 = acquire();
                        tc_fence_after();
                        mma_kblock(d1, h_addr + kk * kChStageBytes, ring_addr(s0), true);
                        mma_kblock(d1 + 128, h_addr + kk * kChStageBytes, ring_addr(s1), true);
                        tc_commit(&empty[s0]); tc_commit(&empty[s1]);
                    }
                    tc_commit(h_free);
                    CH_TRACE(0, it * 100 + 40 + j);
                };
                mbar_wait(e1_done, it & 1);
                tc_fence_after();
                CH_TRACE(0, it * 100 + 3);
                f1(0); f1(1); f2(0); f1(2); f2(1); f1(3);
                tc_commit(y_free);  // every FFN1 MMA (the readers of Y) has been issued
                f2(2); f2(3);
                tc_commit(f2_full);
            }
        }
    } else {
        // ===================================================== epilogue warps (2..9)
        const int q = warp & 3;            // TMEM lane quarter
        const int hf = (warp - 2) >> 2;    // column half owned by this thread
        const int row = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int pair_bar = 1 + q;        // named barrier of the two warps sharing a lane quarter (64 threads)
        const int half_bar = 5 + hf;       // named barrier of the four warps of one column half (128 threads)
        const int sw = row & 7;            // 128B-swizzle phase of this row
        uint32_t n_f1[2] = {0, 0};         // FFN1 chunks seen per ACC2 half
        uint32_t n_hfree = 0;              // waits on h_free so far
        (void)n_f1; (void)n_hfree;
        float v[32];
        int tr_n = (warp == 2 && lane == 0) ? 0 : kChTraceSlots;

        // combines this thread's (mean, M2) over 128 columns with its partner's -> mean, rstd over 256 columns
        auto pair_stats = [&](float mean_a, float m2_a, float& mean, float& rstd) {
            s_stat[hf * 128 + row] = make_float2(mean_a, m2_a);
            asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
            const float2 o = s_stat[(hf ^ 1) * 128 + row];
            asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");  // both reads done before the slot is reused
            const float d = mean_a - o.x;
            mean = 0.5f * (mean_a + o.x);
            const float var = (m2_a + o.y + d * d * 64.f) * (1.f / 256.f);
            rstd = rsqrtf(fmaxf(var, 0.f) + 1e-5f);
        };

        for (int it = 0; it < my_tiles; ++it) {
            const int par = it & 1;
            const int row0 = (blockIdx.x + it * gridDim.x) * 128;
            const uint32_t d1 = tmem_base + par * 256 + lane_off + hf * 128;
            const uint32_t a2 = tmem_base + (par ^ 1) * 256 + lane_off;
            const int qd = par * 2, qa = (par ^ 1) * 2;

            // ---------------- E1: pre = acc + bo + x ; y = LN(pre) ; D1 <- y + b2 ; Y <- bf16(y)
            CH_TRACE(1, it * 100);
            mbar_wait(y_full, it & 1);
            CH_TRACE(1, it * 100 + 1);
            mbar_wait(&g1_full[par], (it >> 1) & 1);
            __syncwarp();
            tc_fence_after();
            CH_TRACE(1, it * 100 + 2);
            float s1 = 0.f, s2 = 0.f, pivot = 0.f;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                tmem_ld32(d1 + c * 32, v);
                tc_wait_ld();
                const int col = hf * 128 + c * 32;
                const uint8_t* yrow = sY + (col >> 6) * kChStageBytes + row * 128;
                const int ch0 = (col & 63) >> 3;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    float r[8];
                    unpack_bf16x8(*reinterpret_cast<const uint4*>(yrow + (((ch0 + g) ^ sw) << 4)), r);
                    const float4 ba = __ldg(reinterpret_cast<const float4*>(p.bo + col + 8 * g));
                    const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bo + col + 8 * g + 4));
                    v[8 * g + 0] += ba.x + r[0]; v[8 * g + 1] += ba.y + r[1]; v[8 * g + 2] += ba.z + r[2]; v[8 * g + 3] += ba.w + r[3];
                    v[8 * g + 4] += bb.x + r[4]; v[8 * g + 5] += bb.y + r[5]; v[8 * g + 6] += bb.z + r[6]; v[8 * g + 7] += bb.w + r[7];
                }
                if (c == 0) pivot = v[0];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float d = v[i] - pivot;
                    s1 += d;
                    s2 = fmaf(d, d, s2);
                }
                tmem_st32(d1 + c * 32, v);
            }
            tc_wait_st();
            float mean, rstd;
            CH_TRACE(1, it * 100 + 3);
            {
                const float m1 = s1 * (1.f / 128.f);
                pair_stats(pivot + m1, fmaxf(s2 - s1 * m1, 0.f), mean, rstd);
            }
            CH_TRACE(1, it * 100 + 4);
            if constexpr (!FFN) {
                // the residual tile has been consumed; the parked pre-norm rows go straight to the store path below
                __syncwarp();
                if (lane == 0) mbar_arrive(y_free);
            } else {
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                tmem_ld32(d1 + c * 32, v);
                tc_wait_ld();
                const int col = hf * 128 + c * 32;
                uint8_t* yrow = sY + (col >> 6) * kChStageBytes + row * 128;
                const int ch0 = (col & 63) >> 3;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    float y[8];
#pragma unroll
                    for (int h2 = 0; h2 < 2; ++h2) {
                        const float4 ga = __ldg(reinterpret_cast<const float4*>(p.gamma + col + 8 * g + 4 * h2));
                        const float4 be = __ldg(reinterpret_cast<const float4*>(p.beta + col + 8 * g + 4 * h2));
                        const float4 b2 = __ldg(reinterpret_cast<const float4*>(p.b2 + col + 8 * g + 4 * h2));
                        const int o = 8 * g + 4 * h2;
                        y[4 * h2 + 0] = (v[o + 0] - mean) * rstd * ga.x + be.x; v[o + 0] = y[4 * h2 + 0] + b2.x;
                        y[4 * h2 + 1] = (v[o + 1] - mean) * rstd * ga.y + be.y; v[o + 1] = y[4 * h2 + 1] + b2.y;
                        y[4 * h2 + 2] = (v[o + 2] - mean) * rstd * ga.z + be.z; v[o + 2] = y[4 * h2 + 2] + b2.z;
                        y[4 * h2 + 3] = (v[o + 3] - mean) * rstd * ga.w + be.w; v[o + 3] = y[4 * h2 + 3] + b2.w;
                    }
                    uint4 pk;
                    pk.x = pack_bf16x2(y[0], y[1]); pk.y = pack_bf16x2(y[2], y[3]);
                    pk.z = pack_bf16x2(y[4], y[5]); pk.w = pack_bf16x2(y[6], y[7]);
                    *reinterpret_cast<uint4*>(yrow + (((ch0 + g) ^ sw) << 4)) = pk;
                }
                tmem_st32(d1 + c * 32, v);
            }
            tc_wait_st();
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(e1_done);
            CH_TRACE(1, it * 100 + 5);

            // ---------------- E2(j): h_j = relu(acc2 + b1) -> H (bf16, two [128 x 64] k-blocks; this thread's 64 cols = k-block hf)
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                const int hb = j & 1;
                mbar_wait(&f1_full[hb], n_f1[hb] & 1);
                ++n_f1[hb];
                CH_TRACE(1, it * 100 + 10 + j);
                mbar_wait(h_free, (n_hfree & 1) ^ 1);  // FFN2 partial j-1 (or the previous tile's last) no longer reads H
                ++n_hfree;
                __syncwarp();
                tc_fence_after();
                CH_TRACE(1, it * 100 + 20 + j);
                uint8_t* hrow = sH + hf * kChStageBytes + row * 128;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    tmem_ld32(a2 + hb * 128 + hf * 64 + c * 32, v);
                    tc_wait_ld();
                    const int hcol = j * 128 + hf * 64 + c * 32;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const float4 ba = __ldg(reinterpret_cast<const float4*>(p.b1 + hcol + 8 * g));
                        const float4 bb = __ldg(reinterpret_cast<const float4*>(p.b1 + hcol + 8 * g + 4));
                        uint4 pk;
                        pk.x = pack_bf16x2(fmaxf(v[8 * g + 0] + ba.x, 0.f), fmaxf(v[8 * g + 1] + ba.y, 0.f));
                        pk.y = pack_bf16x2(fmaxf(v[8 * g + 2] + ba.z, 0.f), fmaxf(v[8 * g + 3] + ba.w, 0.f));
                        pk.z = pack_bf16x2(fmaxf(v[8 * g + 4] + bb.x, 0.f), fmaxf(v[8 * g + 5] + bb.y, 0.f));
                        pk.w = pack_bf16x2(fmaxf(v[8 * g + 6] + bb.z, 0.f), fmaxf(v[8 * g + 7] + bb.w, 0.f));
                        *reinterpret_cast<uint4*>(hrow + (((c * 4 + g) ^ sw) << 4)) = pk;
                    }
                }
                fence_async_smem();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&qfree[qa + hb]);
                    mbar_arrive(h_full);
                }
                CH_TRACE(1, it * 100 + 30 + j);
            }

            // ---------------- LN2: out = LN(D1) -> bf16 -> staging -> TMA store
            mbar_wait(f2_full, it & 1);
            __syncwarp();
            tc_fence_after();
            CH_TRACE(1, it * 100 + 40);
            s1 = 0.f; s2 = 0.f;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                tmem_ld32(d1 + c * 32, v);
                tc_wait_ld();
                if (c == 0) pivot = v[0];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float d = v[i] - pivot;
                    s1 += d;
                    s2 = fmaf(d, d, s2);
                }
            }
            {
                const float m1 = s1 * (1.f / 128.f);
                pair_stats(pivot + m1, fmaxf(s2 - s1 * m1, 0.f), mean, rstd);
            }
            CH_TRACE(1, it * 100 + 41);
            }  // FFN
            // ---------------- store path: out = LN(rows parked in D1) -> bf16 -> staging -> TMA store
            uint8_t* obuf = sOut + hf * kChStageBytes;  // one [128 x 64] staging box per column half, used twice per tile
#pragma unroll 1
            for (int bx = 0; bx < 2; ++bx) {
                // the TMA store that last read this staging box has finished reading it
                if (warp == 2 + 4 * hf && lane == 0) tma_store_wait_read<0>();
                asm volatile("bar.sync %0, 128;" ::"r"(half_bar) : "memory");
                uint8_t* orow = obuf + row * 128;
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int c = bx * 2 + cc;
                    tmem_ld32(d1 + c * 32, v);
                    tc_wait_ld();
                    const int col = hf * 128 + c * 32;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        float y[8];
#pragma unroll
                        for (int h2 = 0; h2 < 2; ++h2) {
                            const float4 ga = __ldg(reinterpret_cast<const float4*>(p.gamma + col + 8 * g + 4 * h2));
                            const float4 be = __ldg(reinterpret_cast<const float4*>(p.beta + col + 8 * g + 4 * h2));
                            const int o = 8 * g + 4 * h2;
                            y[4 * h2 + 0] = (v[o + 0] - mean) * rstd * ga.x + be.x;
                            y[4 * h2 + 1] = (v[o + 1] - mean) * rstd * ga.y + be.y;
                            y[4 * h2 + 2] = (v[o + 2] - mean) * rstd * ga.z + be.z;
                            y[4 * h2 + 3] = (v[o + 3] - mean) * rstd * ga.w + be.w;
                        }
                        uint4 pk;
                        pk.x = pack_bf16x2(y[0], y[1]); pk.y = pack_bf16x2(y[2], y[3]);
                        pk.z = pack_bf16x2(y[4], y[5]); pk.w = pack_bf16x2(y[6], y[7]);
                        *reinterpret_cast<uint4*>(orow + (((cc * 4 + g) ^ sw) << 4)) = pk;
                    }
                }
                fence_async_smem();
                asm volatile("bar.sync %0, 128;" ::"r"(half_bar) : "memory");
                if (warp == 2 + 4 * hf && lane == 0) {
                    tma_store_2d(&tmap_out, obuf, hf * 128 + bx * 64, row0);
                    tma_store_commit();
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&qfree[qd]);
                mbar_arrive(&qfree[qd + 1]);
            }
            CH_TRACE(1, it * 100 + 42);
        }
        if ((warp == 2 || warp == 6) && lane == 0) tma_store_wait_all<0>();  // smem must outlive the last stores
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace etude
