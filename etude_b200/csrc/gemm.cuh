// Persistent warp-specialised tcgen05 GEMM:  C[M,N] = A[M,K] * W[N,K]^T  (both bf16, K contiguous) with fp32
// accumulators in TMEM and fused epilogues (bias / ReLU / residual+LayerNorm / sigmoid+argmax heads).
//
// Replaces, on the reference path, every nn.Linear call site (amt_apc.py:342-344,371,386-389,186-189,217-220),
// the residual adds + nn.LayerNorm that follow fc_o / fc_2 (amt_apc.py:250,256,276,282,304,310,316) and the
// sigmoid / argmax heads (amt_apc.py:186-189,217-220; extractor.py:239-248).
//
// Roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> smem transpose -> coalesced global stores).
// Pipelines: smem ring full/empty (TMA <-> MMA) and 2 TMEM accumulators full/empty (MMA <-> epilogue), so the
// epilogue of tile i overlaps the mainloop of tile i+1.
#pragma once
#include "common.cuh"

namespace etude {

enum GemmEpilogue : int {
    EPI_BIAS = 0,       // out_bf16 = acc + bias
    EPI_BIAS_RELU = 1,  // out_bf16 = relu(acc + bias)
    EPI_RESID_LN = 2,   // y = LN(acc + bias + resid) -> out_f32 / out_bf16 (+ optional permuted, scaled copy)
    EPI_HEADS = 3,      // cols 0..2 -> sigmoid rolls, cols 3..130 -> argmax velocity (N tile = 144)
};

struct GemmParams {
    int M, N, K;
    int num_m_tiles, num_n_tiles;
    const float* bias;  // [N] (padded to the tile width)
    // EPI_BIAS / EPI_BIAS_RELU
    __nv_bfloat16* out_bf16;
    int ld_out;  // elements per output row
    // EPI_RESID_LN  (N == 256)
    const float* resid;  // fp32 [*,256]
    int resid_mod;       // 0: resid row == row;  >0: resid row == row % resid_mod (broadcast embedding)
    const float* ln_gamma;
    const float* ln_beta;
    float* out_f32;  // may be null
    // optional second output: rows (w, f, n) -> (w, n, f), value * perm_scale + perm_pos[f]   (amt_apc.py:203-205)
    float* perm_f32;
    __nv_bfloat16* perm_bf16;
    const float* perm_pos;  // [512,256]
    float perm_scale;
    // EPI_HEADS
    int heads_time_major;     // 0: rows are (gframe, note); 1: rows are (window, note, frame)
    const int64_t* heads_row0;  // per window: first output row of this window in the rolls
    float* roll_onset;
    float* roll_offset;
    float* roll_mpe;
    int8_t* roll_velocity;
    float* vel_logits;  // optional fp32 [rows,128] in (window, frame, note) order
};

constexpr int kGemmThreads = 192;
constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kStages = 4;
constexpr int kScratchPerWarp = 32 * 33;  // floats

template <int BLOCK_N>
constexpr size_t gemm_smem_bytes() {
    return 1024 /*align slack*/ + (size_t)kStages * (kBlockM * kBlockK * 2 + BLOCK_N * kBlockK * 2) + 4 * kScratchPerWarp * 4 + 256;
}

template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int A_BYTES = kBlockM * kBlockK * 2;
    constexpr int B_BYTES = BLOCK_N * kBlockK * 2;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    float* scratch_all = reinterpret_cast<float*>(smem + kStages * STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * STAGE_BYTES + 4 * kScratchPerWarp * 4);
    uint64_t* full_bar = bars;                    // [kStages]
    uint64_t* empty_bar = bars + kStages;         // [kStages]
    uint64_t* tmem_full = bars + 2 * kStages;     // [2]
    uint64_t* tmem_empty = bars + 2 * kStages + 2;  // [2]
    uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = p.num_m_tiles * p.num_n_tiles;
    const int num_kb = p.K / kBlockK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], 4);
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_base_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_ptr;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_blk = tile / p.num_n_tiles, n_blk = tile % p.num_n_tiles;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                    tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * kBlockK, m_blk * kBlockM);
                    tma_load_2d(sa + A_BYTES, &tmap_b, &full_bar[stage], kb * kBlockK, n_blk * BLOCK_N);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BLOCK_N, 0, 0);
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
                    const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
                    const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) {
                        umma_bf16_ss(tmem_d, make_sw128_desc(a_addr + k * 32), make_sw128_desc(b_addr + k * 32), idesc,
                                     (kb | k) != 0);
                    }
                    tc_commit(&empty_bar[stage]);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                tc_commit(&tmem_full[as]);
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else {
        // ===== epilogue warps =====
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        float* scratch = scratch_all + (warp - 2) * kScratchPerWarp;
        int as = 0;
        uint32_t aphase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_blk = tile / p.num_n_tiles, n_blk = tile % p.num_n_tiles;
            mbar_wait(&tmem_full[as], aphase);
            __syncwarp();
            tc_fence_after();
            const uint32_t tbase = tmem_base + as * 256 + ((uint32_t)(q * 32) << 16);
            const int row0 = m_blk * kBlockM + q * 32;  // first row of this warp
            const int col0 = n_blk * BLOCK_N;
            float v[32];

            if constexpr (EPI == EPI_BIAS || EPI == EPI_BIAS_RELU) {
#pragma unroll 1
                for (int c = 0; c < BLOCK_N / 32; ++c) {
                    tmem_ld32(tbase + c * 32, v);
                    tc_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float x = v[j] + __ldg(&p.bias[col0 + c * 32 + j]);
                        if (EPI == EPI_BIAS_RELU) x = fmaxf(x, 0.f);
                        scratch[lane * 33 + j] = x;
                    }
                    __syncwarp();
                    // two rows per iteration: lanes 0..15 -> row rr, lanes 16..31 -> row rr+1; 2 columns per lane
                    const int half = lane >> 4, l2 = (lane & 15) * 2;
#pragma unroll 4
                    for (int rr = 0; rr < 32; rr += 2) {
                        const int r = rr + half;
                        const int row = row0 + r;
                        if (row < p.M) {
                            uint32_t pk = pack_bf16x2(scratch[r * 33 + l2], scratch[r * 33 + l2 + 1]);
                            *reinterpret_cast<uint32_t*>(p.out_bf16 + (size_t)row * p.ld_out + col0 + c * 32 + l2) = pk;
                        }
                    }
                    __syncwarp();
                }
            } else if constexpr (EPI == EPI_RESID_LN) {
                static_assert(EPI != EPI_RESID_LN || BLOCK_N == 256, "LayerNorm epilogue needs the whole row in one tile");
                // pass 1: v = acc + bias + resid, kept in TMEM; row sum
                float sum = 0.f;
#pragma unroll 1
                for (int c = 0; c < 8; ++c) {
#pragma unroll 4
                    for (int rr = 0; rr < 32; ++rr) {
                        const int row = row0 + rr;
                        float x = 0.f;
                        if (row < p.M) {
                            const size_t rrow = p.resid_mod ? (size_t)(row % p.resid_mod) : (size_t)row;
                            x = __ldg(&p.resid[rrow * 256 + c * 32 + lane]);
                        }
                        scratch[rr * 33 + lane] = x;
                    }
                    __syncwarp();
                    tmem_ld32(tbase + c * 32, v);
                    tc_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        v[j] = v[j] + __ldg(&p.bias[c * 32 + j]) + scratch[lane * 33 + j];
                        sum += v[j];
                    }
                    __syncwarp();
                    tmem_st32(tbase + c * 32, v);
                }
                tc_wait_st();
                const float mean = sum * (1.f / 256.f);
                // pass 2: centred second moment
                float sq = 0.f;
#pragma unroll 1
                for (int c = 0; c < 8; ++c) {
                    tmem_ld32(tbase + c * 32, v);
                    tc_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float d = v[j] - mean;
                        sq += d * d;
                    }
                }
                const float rstd = rsqrtf(sq * (1.f / 256.f) + 1e-5f);
                // pass 3: normalise, transpose through smem, coalesced row stores
                const int my_row = row0 + lane;
                (void)my_row;
#pragma unroll 1
                for (int c = 0; c < 8; ++c) {
                    tmem_ld32(tbase + c * 32, v);
                    tc_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int col = c * 32 + j;
                        scratch[lane * 33 + j] = (v[j] - mean) * rstd * __ldg(&p.ln_gamma[col]) + __ldg(&p.ln_beta[col]);
                    }
                    __syncwarp();
                    const int col = c * 32 + lane;
#pragma unroll 4
                    for (int rr = 0; rr < 32; ++rr) {
                        const int row = row0 + rr;
                        if (row < p.M) {
                            const float y = scratch[rr * 33 + lane];
                            if (p.out_f32) p.out_f32[(size_t)row * 256 + col] = y;
                            if (p.out_bf16) p.out_bf16[(size_t)row * 256 + col] = __float2bfloat16(y);
                            if (p.perm_f32) {
                                // row = (w*512 + f)*88 + n  ->  (w*88 + n)*512 + f
                                const int n = row % kNotes, wf = row / kNotes;
                                const int f = wf % kFrames, w = wf / kFrames;
                                const size_t prow = ((size_t)w * kNotes + n) * kFrames + f;
                                const float z = y * p.perm_scale + __ldg(&p.perm_pos[f * 256 + col]);
                                p.perm_f32[prow * 256 + col] = z;
                                p.perm_bf16[prow * 256 + col] = __float2bfloat16(z);
                            }
                        }
                    }
                    __syncwarp();
                }
            } else {  // EPI_HEADS, BLOCK_N == 144: thread == row
                const int row = row0 + lane;
                float best = -INFINITY;
                int best_i = 0;
                float head3[3] = {0.f, 0.f, 0.f};
                size_t out_idx = 0, logit_row = 0;
                if (row < p.M) {
                    int w, f, n;
                    if (p.heads_time_major) {
                        f = row % kFrames;
                        const int wn = row / kFrames;
                        n = wn % kNotes;
                        w = wn / kNotes;
                    } else {
                        n = row % kNotes;
                        const int wf = row / kNotes;
                        f = wf % kFrames;
                        w = wf / kFrames;
                    }
                    out_idx = (size_t)(p.heads_row0[w] + f) * kNotes + n;
                    logit_row = ((size_t)w * kFrames + f) * kNotes + n;
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
                if (row < p.M) {
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
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace etude
