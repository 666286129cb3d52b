// Persistent warp-specialised tcgen05 GEMM:  C[M,N] = A[M,K] * W[N,K]^T  (both bf16, K contiguous) with fp32
// accumulators in TMEM and fused epilogues (bias / ReLU / residual+LayerNorm / sigmoid+argmax heads).
//
// Replaces, on the reference path, every nn.Linear call site (amt_apc.py:342-344,371,386-389,186-189,217-220),
// the residual adds + nn.LayerNorm that follow fc_o / fc_2 (amt_apc.py:250,256,276,282,304,310,316) and the
// sigmoid / argmax heads (amt_apc.py:186-189,217-220; extractor.py:239-248).
//
// Roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue.  Pipelines: smem ring full/empty (TMA <-> MMA) and 2 TMEM accumulators full/empty
// (MMA <-> epilogue), so the epilogue of tile i overlaps the mainloop of tile i+1.
//
// Epilogue data movement is all TMA: each epilogue warp owns 32 accumulator rows (tcgen05.ld 32x32b: one thread
// per row), stages its output in a private 128B/64B-swizzled smem box with conflict-free 16-byte stores and lets
// the TMA engine write the box to HBM (cp.async.bulk.tensor store, double buffered with bulk groups); the fp32
// residual of the LayerNorm epilogue arrives the same way (TMA load into a private double-buffered box, one
// mbarrier per buffer, prefetched two chunks ahead).  Per-row LayerNorm statistics are thread-local; the
// pre-norm row is parked in TMEM (tcgen05.st) between the statistics pass and the normalise pass.
#pragma once
#include "common.cuh"

namespace etude {

enum GemmEpilogue : int {
    EPI_BIAS = 0,       // out_bf16 = acc + bias
    EPI_BIAS_RELU = 1,  // out_bf16 = relu(acc + bias)
    EPI_RESID_LN = 2,   // y = LN(acc + bias + resid) -> out_f32 and out_bf16
    EPI_HEADS = 3,      // cols 0..2 -> sigmoid rolls, cols 3..130 -> argmax velocity (N tile = 144)
};

struct GemmParams {
    int M, N, K;
    int num_m_tiles, num_n_tiles;
    const float* bias;  // [N] (padded to the tile width)
    // EPI_RESID_LN  (N == 256)
    int resid_mod;  // 0: resid row == row;  >0: resid row == row % resid_mod (wrapped embedding table, see api.cu)
    const float* ln_gamma;
    const float* ln_beta;
    // EPI_HEADS
    int heads_time_major;       // 0: rows are (gframe, note); 1: rows are (window, note, frame)
    const int64_t* heads_row0;  // per window: first output row of this window in the rolls
    float* roll_onset;
    float* roll_offset;
    float* roll_mpe;
    int8_t* roll_velocity;
    float* vel_logits;  // optional fp32 [rows,128] in (window, frame, note) order
    int frames;         // frames per window (512: AMT-APC extractor, 128: HFT_Transformer)
    int keep0, keepn;   // only window frames [keep0, keep0 + keepn) reach the rolls, at row heads_row0[w] + frame - keep0
                        // (everything for _transcript; the centre half for _transcript_stride, hft_transformer.py:352-441)
};

constexpr int kGemmThreads = 192;
constexpr int kBlockM = 128;
constexpr int kBlockK = 64;

template <int EPI>
__host__ __device__ constexpr int gemm_stages() { return EPI == EPI_RESID_LN ? 2 : 4; }   // the LN GEMMs are HBM-bound 3-8x over: 2 stages feed them
template <int EPI>
__host__ __device__ constexpr int gemm_epi_bytes_per_warp() {
    return EPI == EPI_RESID_LN ? (2 * 4096 /*resid*/ + 2 * 4096 /*out f32*/ + 2 * 2048 /*out bf16*/)
                               : (EPI == EPI_HEADS ? 0 : 2 * 4096 /*out bf16, 64-col boxes*/);
}
template <int BLOCK_N, int EPI>
constexpr size_t gemm_smem_bytes() {
    return 1024 /*align slack*/ + (size_t)gemm_stages<EPI>() * (kBlockM * kBlockK * 2 + BLOCK_N * kBlockK * 2) +
           4 * gemm_epi_bytes_per_warp<EPI>() + (EPI == EPI_RESID_LN ? 3 : 1) * 1024 /*bias (, gamma, beta)*/ + 256 /*barriers*/;
}

template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_out_bf16, const __grid_constant__ CUtensorMap tmap_out_f32,
                    const __grid_constant__ CUtensorMap tmap_resid, const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int STAGES = gemm_stages<EPI>();
    constexpr int A_BYTES = kBlockM * kBlockK * 2;
    constexpr int B_BYTES = BLOCK_N * kBlockK * 2;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr int EPI_BYTES = gemm_epi_bytes_per_warp<EPI>();
    uint8_t* epi_all = smem + STAGES * STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_all + 4 * EPI_BYTES);  // 256 B
    float* s_bias = reinterpret_cast<float*>(epi_all + 4 * EPI_BYTES + 256);  // [256]
    float* s_gamma = s_bias + 256;                                            // LN only
    float* s_beta = s_gamma + 256;                                            // LN only
    uint64_t* full_bar = bars;                     // [STAGES]
    uint64_t* empty_bar = bars + STAGES;           // [STAGES]
    uint64_t* tmem_full = bars + 2 * STAGES;       // [2]
    uint64_t* tmem_empty = bars + 2 * STAGES + 2;  // [2]
    uint64_t* resid_bar = bars + 2 * STAGES + 4;   // [4 warps][2 buffers]
    uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 12);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform (see chain3.cuh on why it matters)
    const int lane = threadIdx.x & 31;
    const int num_tiles = p.num_m_tiles * p.num_n_tiles;
    const int num_kb = p.K / kBlockK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        if (EPI != EPI_HEADS) tma_prefetch_desc(&tmap_out_bf16);
        if (EPI == EPI_RESID_LN) {
            tma_prefetch_desc(&tmap_out_f32);
            tma_prefetch_desc(&tmap_resid);
        }
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], 4);
        }
        for (int s = 0; s < 8; ++s) mbar_init(&resid_bar[s], 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_base_ptr, 512);
    if (EPI == EPI_RESID_LN && warp >= 2) {  // LayerNorm affine parameters and the (single n-tile) bias: once per CTA
        for (int i = threadIdx.x - 64; i < 256; i += 128) {
            s_bias[i] = p.bias[i];
            s_gamma[i] = p.ln_gamma[i];
            s_beta[i] = p.ln_beta[i];
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_ptr;

    if (warp == 0) {
        // ===== TMA producer (warp-uniform loop, one elected lane issues) =====
        const bool leader = elect_one();
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_blk = tile / p.num_n_tiles, n_blk = tile % p.num_n_tiles;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (leader) {
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                    tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * kBlockK, m_blk * kBlockM);
                    tma_load_2d(sa + A_BYTES, &tmap_b, &full_bar[stage], kb * kBlockK, n_blk * BLOCK_N);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (warp-uniform loop, one elected lane issues) =====
        const bool leader = elect_one();
        constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BLOCK_N, 0, 0);
        const uint64_t a_desc0 = make_sw128_desc(smem_u32(smem));
        int stage = 0;
        uint32_t phase = 0;
        int as = 0;
        uint32_t aphase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            mbar_wait(&tmem_empty[as], aphase ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + as * 256;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (leader) {
                    const uint64_t ad = a_desc0 + (uint64_t)(stage * (STAGE_BYTES >> 4)), bd = ad + (uint64_t)(A_BYTES >> 4);
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) umma_bf16_ss(tmem_d, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
                    tc_commit(&empty_bar[stage]);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (leader) tc_commit(&tmem_full[as]);
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
    } else {
        // ===== epilogue warps =====
        const int q = warp & 3;   // TMEM lane quarter this warp may access
        const int ew = warp - 2;  // private smem slice
        uint8_t* epi = epi_all + ew * EPI_BYTES;
        int as = 0;
        uint32_t aphase = 0;
        uint32_t n_store = 0;  // output chunks staged so far (selects the double buffer)
        uint32_t n_resid = 0;  // residual chunks consumed so far
        uint64_t* rbar = resid_bar + ew * 2;
        (void)rbar; (void)n_resid; (void)n_store; (void)epi;

        // residual chunk c (32 fp32 columns x this warp's 32 rows) of the tile with first row `r0` -> buffer (seq & 1)
        auto issue_resid = [&](int tile, int c, uint32_t seq) {
            if (lane == 0) {
                const int m_blk = tile / p.num_n_tiles;
                int r0 = m_blk * kBlockM + q * 32;
                if (p.resid_mod) r0 %= p.resid_mod;
                uint64_t* bar = &rbar[seq & 1];
                mbar_expect_tx(bar, 4096);
                tma_load_2d(epi + (seq & 1) * 4096, &tmap_resid, bar, c * 32, r0);
            }
        };
        if constexpr (EPI == EPI_RESID_LN) {
            if ((int)blockIdx.x < num_tiles) {
                issue_resid(blockIdx.x, 0, 0);
                issue_resid(blockIdx.x, 1, 1);
            }
        }

        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_blk = tile / p.num_n_tiles, n_blk = tile % p.num_n_tiles;
            const int row0 = m_blk * kBlockM + q * 32;  // first row of this warp
            const int col0 = n_blk * BLOCK_N;
            float v[32];

            if constexpr (EPI == EPI_BIAS || EPI == EPI_BIAS_RELU) {
                // bias slice of this n-tile -> smem (all 4 epilogue warps, named barrier 1)
                asm volatile("bar.sync 1, 128;" ::: "memory");   // previous tile's readers are done
                for (int i = threadIdx.x - 64; i < BLOCK_N; i += 128) s_bias[i] = p.bias[col0 + i];
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            mbar_wait(&tmem_full[as], aphase);
            __syncwarp();
            tc_fence_after();
            const uint32_t tbase = tmem_base + as * 256 + ((uint32_t)(q * 32) << 16);

            if constexpr (EPI == EPI_BIAS || EPI == EPI_BIAS_RELU) {
#pragma unroll 1
                for (int c = 0; c < BLOCK_N / 64; ++c) {  // 64-column output boxes
                    uint8_t* obuf = epi + (n_store & 1) * 4096;
                    if (lane == 0) tma_store_wait_read<1>();  // the store that last read this buffer has drained
                    __syncwarp();
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        tmem_ld32(tbase + c * 64 + hh * 32, v);
                        tc_wait_ld();
                        const float4* b4 = reinterpret_cast<const float4*>(s_bias + c * 64 + hh * 32);
#pragma unroll
                        for (int g = 0; g < 4; ++g) {  // 8 columns -> one 16-byte chunk
                            const float4 ba = b4[2 * g], bb = b4[2 * g + 1];
                            float x0 = v[8 * g + 0] + ba.x, x1 = v[8 * g + 1] + ba.y, x2 = v[8 * g + 2] + ba.z, x3 = v[8 * g + 3] + ba.w;
                            float x4 = v[8 * g + 4] + bb.x, x5 = v[8 * g + 5] + bb.y, x6 = v[8 * g + 6] + bb.z, x7 = v[8 * g + 7] + bb.w;
                            if (EPI == EPI_BIAS_RELU) {
                                x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); x2 = fmaxf(x2, 0.f); x3 = fmaxf(x3, 0.f);
                                x4 = fmaxf(x4, 0.f); x5 = fmaxf(x5, 0.f); x6 = fmaxf(x6, 0.f); x7 = fmaxf(x7, 0.f);
                            }
                            uint4 pk;
                            pk.x = pack_bf16x2(x0, x1); pk.y = pack_bf16x2(x2, x3); pk.z = pack_bf16x2(x4, x5); pk.w = pack_bf16x2(x6, x7);
                            const int chunk = hh * 4 + g;
                            *reinterpret_cast<uint4*>(obuf + lane * 128 + ((chunk ^ (lane & 7)) << 4)) = pk;
                        }
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmap_out_bf16, obuf, col0 + c * 64, row0);
                        tma_store_commit();
                    }
                    ++n_store;
                }
            } else if constexpr (EPI == EPI_RESID_LN) {
                static_assert(EPI != EPI_RESID_LN || BLOCK_N == 256, "LayerNorm epilogue needs the whole row in one tile");
                // ---- pass 1: x = acc + bias + resid -> TMEM; shifted one-pass statistics (pivot = first element)
                float pivot = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 1
                for (int c = 0; c < 8; ++c) {
                    const uint32_t seq = n_resid;
                    mbar_wait(&rbar[seq & 1], (seq >> 1) & 1);
                    __syncwarp();
                    const uint8_t* rbuf = epi + (seq & 1) * 4096;
                    tmem_ld32(tbase + c * 32, v);
                    tc_wait_ld();
                    const float4* b4 = reinterpret_cast<const float4*>(s_bias + c * 32);
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        const float4 r4 = *reinterpret_cast<const float4*>(rbuf + lane * 128 + ((g ^ (lane & 7)) << 4));
                        const float4 bb = b4[g];
                        v[4 * g + 0] += bb.x + r4.x;
                        v[4 * g + 1] += bb.y + r4.y;
                        v[4 * g + 2] += bb.z + r4.z;
                        v[4 * g + 3] += bb.w + r4.w;
                    }
                    if (c == 0) pivot = v[0];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float d = v[j] - pivot;
                        s1 += d;
                        s2 = fmaf(d, d, s2);
                    }
                    tmem_st32(tbase + c * 32, v);
                    __syncwarp();  // every lane has finished reading rbuf
                    ++n_resid;
                    // prefetch: chunk c+2 of this tile, or the first two chunks of this CTA's next tile
                    if (c + 2 < 8) issue_resid(tile, c + 2, seq + 2);
                    else if (tile + (int)gridDim.x < num_tiles) issue_resid(tile + gridDim.x, c + 2 - 8, seq + 2);
                }
                tc_wait_st();
                const float m1 = s1 * (1.f / 256.f);
                const float mean = pivot + m1;
                const float var = fmaxf(s2 * (1.f / 256.f) - m1 * m1, 0.f);
                const float rstd = rsqrtf(var + 1e-5f);
                // ---- pass 2: normalise; fp32 + bf16 boxes of 32 columns -> TMA stores
#pragma unroll 1
                for (int c = 0; c < 8; ++c) {
                    uint8_t* of32 = epi + 8192 + (n_store & 1) * 4096;
                    uint8_t* ob16 = epi + 16384 + (n_store & 1) * 2048;
                    if (lane == 0) tma_store_wait_read<1>();
                    __syncwarp();
                    tmem_ld32(tbase + c * 32, v);
                    tc_wait_ld();
                    const float4* g4 = reinterpret_cast<const float4*>(s_gamma + c * 32);
                    const float4* be4 = reinterpret_cast<const float4*>(s_beta + c * 32);
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        const float4 ga = g4[g], be = be4[g];
                        float4 y;
                        y.x = (v[4 * g + 0] - mean) * rstd * ga.x + be.x;
                        y.y = (v[4 * g + 1] - mean) * rstd * ga.y + be.y;
                        y.z = (v[4 * g + 2] - mean) * rstd * ga.z + be.z;
                        y.w = (v[4 * g + 3] - mean) * rstd * ga.w + be.w;
                        *reinterpret_cast<float4*>(of32 + lane * 128 + ((g ^ (lane & 7)) << 4)) = y;
                        v[4 * g + 0] = y.x; v[4 * g + 1] = y.y; v[4 * g + 2] = y.z; v[4 * g + 3] = y.w;
                    }
#pragma unroll
                    for (int g = 0; g < 4; ++g) {  // 32 bf16 = 64-byte rows, 64B swizzle
                        uint4 pk;
                        pk.x = pack_bf16x2(v[8 * g + 0], v[8 * g + 1]); pk.y = pack_bf16x2(v[8 * g + 2], v[8 * g + 3]);
                        pk.z = pack_bf16x2(v[8 * g + 4], v[8 * g + 5]); pk.w = pack_bf16x2(v[8 * g + 6], v[8 * g + 7]);
                        *reinterpret_cast<uint4*>(ob16 + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4)) = pk;
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmap_out_f32, of32, c * 32, row0);
                        tma_store_2d(&tmap_out_bf16, ob16, c * 32, row0);
                        tma_store_commit();
                    }
                    ++n_store;
                }
            } else {  // EPI_HEADS, BLOCK_N == 144: thread == row, tiny output (0.3 % of the path's FLOPs)
                const int row = row0 + lane;
                float best = -INFINITY;
                int best_i = 0;
                float head3[3] = {0.f, 0.f, 0.f};
                size_t out_idx = 0, logit_row = 0;
                bool kept = false;
                if (row < p.M) {
                    int w, f, n;
                    if (p.heads_time_major) {
                        f = row % p.frames;
                        const int wn = row / p.frames;
                        n = wn % kNotes;
                        w = wn / kNotes;
                    } else {
                        n = row % kNotes;
                        const int wf = row / kNotes;
                        f = wf % p.frames;
                        w = wf / p.frames;
                    }
                    kept = f >= p.keep0 && f < p.keep0 + p.keepn;
                    out_idx = (size_t)(p.heads_row0[w] + f - p.keep0) * kNotes + n;
                    logit_row = ((size_t)w * p.frames + f) * kNotes + n;
                }
#pragma unroll 1
                for (int c = 0; c < 5; ++c) {  // 5 chunks of 32 columns cover 160 >= 131 (accumulator stride is 256)
                    tmem_ld32(tbase + c * 32, v);
                    tc_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int col = c * 32 + j;
                        if (col < 3 + kVel) {
                            const float x = v[j] + __ldg(&p.bias[col]);
                            if (col < 3) {
                                if (c == 0 && j < 3) head3[j < 3 ? j : 0] = x;
                            } else {
                                if (x > best) { best = x; best_i = col - 3; }  // first maximum wins, like torch.argmax
                                if (p.vel_logits && row < p.M) p.vel_logits[logit_row * kVel + (col - 3)] = x;
                            }
                        }
                    }
                }
                if (kept) {
                    p.roll_onset[out_idx] = 1.f / (1.f + expf(-head3[0]));
                    p.roll_offset[out_idx] = 1.f / (1.f + expf(-head3[1]));
                    p.roll_mpe[out_idx] = 1.f / (1.f + expf(-head3[2]));
                    p.roll_velocity[out_idx] = (int8_t)best_i;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[as]);
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
        if (EPI != EPI_HEADS && lane == 0) tma_store_wait_all<0>();  // smem must outlive the last stores
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// B-stationary variant for the K = 256 projections (Q|K|V, the packed cross-attention K|V, fc_q of the decoder):
// out_bf16[M, N] = A[M, 256] W[N, 256]^T + bias.  Every CTA keeps ONE 256-column slice of W (128 KB) resident in smem
// for its whole life and streams only A tiles, so the L2 -> SM traffic per 128 x 256 output tile drops from 192 KB
// (A + the re-streamed W slice) to 64 KB: the generic kernel above runs these shapes at the L2 fabric limit
// (profiles/r1c: 79 % of the ~6300 B/clk LTS cap), this one is bounded by the HBM write of the output instead.
// CTA b owns n-slice b % num_n_tiles and walks m-tiles b / num_n_tiles + k * (gridDim.x / num_n_tiles), so the CTAs
// that share an A tile read it at about the same time (one HBM read, the rest L2 hits).
constexpr int kBsStages = 4;
constexpr int kBsThreads = 10 * 32;                   // TMA, MMA, 8 epilogue warps
constexpr int kBsABytes = kBlockM * kBlockK * 2;      // 16 KB per k-block of A
constexpr int kBsBBytes = 256 * kBlockK * 2;          // 32 KB per k-block of the W slice
// no alignment slack (window declared 1024-aligned and checked): leaves room for a one-warp block of another stream on the SM
constexpr size_t kGemmBsSmemBytes = 4 * kBsBBytes + kBsStages * kBsABytes + 8 * 4096 /*staging*/ + 1024 /*bias*/ + 256;

__global__ void __launch_bounds__(kBsThreads, 1)
gemm_bstat_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ CUtensorMap tmap_out, const GemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
    uint8_t* sB = smem;                                  // 4 k-blocks [256 x 64] bf16, SW128
    uint8_t* sA = sB + 4 * kBsBBytes;                    // ring of [128 x 64] k-blocks
    uint8_t* epi_all = sA + kBsStages * kBsABytes;       // 8 warps x 4 KB staging
    float* s_bias = reinterpret_cast<float*>(epi_all + 8 * 4096);
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_all + 8 * 4096 + 1024);
    uint64_t* full_bar = bars;                           // [kBsStages]
    uint64_t* empty_bar = bars + kBsStages;              // [kBsStages]
    uint64_t* tmem_full = bars + 2 * kBsStages;          // [2]
    uint64_t* tmem_empty = bars + 2 * kBsStages + 2;     // [2]
    uint64_t* b_full = bars + 2 * kBsStages + 4;
    uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kBsStages + 5);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform
    const int lane = threadIdx.x & 31;
    const int groups = gridDim.x / p.num_n_tiles;        // CTAs per n-slice
    const int n_blk = blockIdx.x % p.num_n_tiles;
    const int m_first = blockIdx.x / p.num_n_tiles;
    const int col0 = n_blk * 256;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a); tma_prefetch_desc(&tmap_b); tma_prefetch_desc(&tmap_out);
        for (int s = 0; s < kBsStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 8); }
        mbar_init(b_full, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_base_ptr, 512);
    if (warp >= 2)
        for (int i = threadIdx.x - 64; i < 256; i += 256) s_bias[i] = p.bias[col0 + i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_ptr;

    if (warp == 0) {
        const bool leader = elect_one();
        if (m_first < p.num_m_tiles) {
            if (leader) {
                mbar_expect_tx(b_full, 4 * kBsBBytes);
                for (int kb = 0; kb < 4; ++kb) tma_load_2d(sB + kb * kBsBBytes, &tmap_b, b_full, kb * kBlockK, col0);
            }
            int stage = 0;
            uint32_t phase = 0;
            for (int m_blk = m_first; m_blk < p.num_m_tiles; m_blk += groups)
                for (int kb = 0; kb < 4; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (leader) {
                        mbar_expect_tx(&full_bar[stage], kBsABytes);
                        tma_load_2d(sA + stage * kBsABytes, &tmap_a, &full_bar[stage], kb * kBlockK, m_blk * kBlockM);
                    }
                    __syncwarp();
                    if (++stage == kBsStages) { stage = 0; phase ^= 1; }
                }
        }
    } else if (warp == 1) {
        const bool leader = elect_one();
        if (m_first < p.num_m_tiles) {
            constexpr uint32_t idesc = make_idesc_bf16(kBlockM, 256, 0, 0);
            const uint64_t a_desc0 = make_sw128_desc(smem_u32(sA)), b_desc0 = make_sw128_desc(smem_u32(sB));
            int stage = 0, as = 0;
            uint32_t phase = 0, aphase = 0;
            mbar_wait(b_full, 0);
            for (int m_blk = m_first; m_blk < p.num_m_tiles; m_blk += groups) {
                mbar_wait(&tmem_empty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * 256;
#pragma unroll
                for (int kb = 0; kb < 4; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (leader) {
                        const uint64_t ad = a_desc0 + (uint64_t)(stage * (kBsABytes >> 4)), bd = b_desc0 + (uint64_t)(kb * (kBsBBytes >> 4));
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) umma_bf16_ss(tmem_d, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
                        tc_commit(&empty_bar[stage]);
                    }
                    __syncwarp();
                    if (++stage == kBsStages) { stage = 0; phase ^= 1; }
                }
                if (leader) tc_commit(&tmem_full[as]);
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else {
        // Eight epilogue warps: TMEM lane quarter warp & 3, column half (warp - 2) >> 2 (two 64-column boxes each).  With
        // four warps and one staging box per warp the epilogue was a latency chain (tcgen05.ld -> wait for the previous
        // box's TMA store to release the box -> pack -> fence -> TMA store, four times per tile, ~4 400 clk) that paced
        // the kernel; two warps per sub-partition overlap each other's waits and the box wait sits behind the arithmetic.
        const int q = warp & 3, ew = warp - 2, half = ew >> 2;
        uint8_t* obuf = epi_all + ew * 4096;
        int as = 0;
        uint32_t aphase = 0;
        for (int m_blk = m_first; m_blk < p.num_m_tiles; m_blk += groups) {
            const int row0 = m_blk * kBlockM + q * 32;
            mbar_wait(&tmem_full[as], aphase);
            __syncwarp();
            tc_fence_after();
            const uint32_t tbase = tmem_base + as * 256 + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int c = 2 * half; c < 2 * half + 2; ++c) {  // 64-column output boxes
                float v[64];
                tmem_ld32(tbase + c * 64, v);
                tmem_ld32(tbase + c * 64 + 32, v + 32);
                tc_wait_ld();
                uint4 pk[8];
                const float4* b4 = reinterpret_cast<const float4*>(s_bias + c * 64);
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const float4 ba = b4[2 * g], bb = b4[2 * g + 1];
                    pk[g].x = pack_bf16x2(v[8 * g + 0] + ba.x, v[8 * g + 1] + ba.y); pk[g].y = pack_bf16x2(v[8 * g + 2] + ba.z, v[8 * g + 3] + ba.w);
                    pk[g].z = pack_bf16x2(v[8 * g + 4] + bb.x, v[8 * g + 5] + bb.y); pk[g].w = pack_bf16x2(v[8 * g + 6] + bb.z, v[8 * g + 7] + bb.w);
                }
                if (c == 2 * half + 1) {  // this warp's part of the accumulator is in registers: hand the buffer back early
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[as]);
                }
                if (lane == 0) tma_store_wait_read<0>();  // the TMA store that last read the staging box has drained it
                __syncwarp();
#pragma unroll
                for (int g = 0; g < 8; ++g) *reinterpret_cast<uint4*>(obuf + lane * 128 + ((g ^ (lane & 7)) << 4)) = pk[g];
                fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tmap_out, obuf, col0 + c * 64, row0);
                    tma_store_commit();
                }
            }
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
        if (lane == 0) tma_store_wait_all<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// (window, frame, note)-major rows -> (window, note, frame)-major, * scale + pos_time[frame]: the single global
// transpose of the decoder (amt_apc.py:203-205).  One warp per 512-byte bf16 row; HBM-bound.
__global__ void __launch_bounds__(256)
transpose_time_kernel(const __nv_bfloat16* __restrict__ in, const float* __restrict__ pos_time, float scale, int n_rows, int frames,
                      __nv_bfloat16* __restrict__ out) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int lane = threadIdx.x & 31;
    const int n = row % kNotes, wf = row / kNotes;
    const int f = wf % frames, w = wf / frames;
    const size_t prow = ((size_t)w * kNotes + n) * frames + f;
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(in + (size_t)row * kHid) + lane);
    const float4 p0 = __ldg(reinterpret_cast<const float4*>(pos_time + (size_t)f * kHid) + 2 * lane);
    const float4 p1 = __ldg(reinterpret_cast<const float4*>(pos_time + (size_t)f * kHid) + 2 * lane + 1);
    const uint32_t wv[4] = {a.x, a.y, a.z, a.w};
    float x[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        x[2 * i] = __uint_as_float(wv[i] << 16);
        x[2 * i + 1] = __uint_as_float(wv[i] & 0xFFFF0000u);
    }
    uint4 o;
    o.x = pack_bf16x2(x[0] * scale + p0.x, x[1] * scale + p0.y);
    o.y = pack_bf16x2(x[2] * scale + p0.z, x[3] * scale + p0.w);
    o.z = pack_bf16x2(x[4] * scale + p1.x, x[5] * scale + p1.y);
    o.w = pack_bf16x2(x[6] * scale + p1.z, x[7] * scale + p1.w);
    reinterpret_cast<uint4*>(out + prow * kHid)[lane] = o;
}

// fp32 <-> bf16 row copies at the encode / decode boundary of _Spec2MIDI (extractor.py:58-75): 8 elements per thread.
__global__ void __launch_bounds__(256) cvt_bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, size_t n8) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n8) return;
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(in) + i);
    const uint32_t wv[4] = {a.x, a.y, a.z, a.w};
    float4 lo, hi;
    lo.x = __uint_as_float(wv[0] << 16); lo.y = __uint_as_float(wv[0] & 0xFFFF0000u);
    lo.z = __uint_as_float(wv[1] << 16); lo.w = __uint_as_float(wv[1] & 0xFFFF0000u);
    hi.x = __uint_as_float(wv[2] << 16); hi.y = __uint_as_float(wv[2] & 0xFFFF0000u);
    hi.z = __uint_as_float(wv[3] << 16); hi.w = __uint_as_float(wv[3] & 0xFFFF0000u);
    reinterpret_cast<float4*>(out)[2 * i] = lo;
    reinterpret_cast<float4*>(out)[2 * i + 1] = hi;
}
__global__ void __launch_bounds__(256) cvt_f32_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t n8) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n8) return;
    const float4 lo = __ldg(reinterpret_cast<const float4*>(in) + 2 * i), hi = __ldg(reinterpret_cast<const float4*>(in) + 2 * i + 1);
    uint4 o;
    o.x = pack_bf16x2(lo.x, lo.y); o.y = pack_bf16x2(lo.z, lo.w); o.z = pack_bf16x2(hi.x, hi.y); o.w = pack_bf16x2(hi.z, hi.w);
    reinterpret_cast<uint4*>(out)[i] = o;
}

}  // namespace etude
