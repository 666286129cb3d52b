// Fused multi-head attention on tcgen05: S = Q K^T into TMEM, fp32 softmax in registers (one thread per query
// row), P (bf16) written back over S in TMEM and fed to the second MMA as its TMEM A operand, O = P V into TMEM
// columns S no longer needs; scores and probabilities never touch HBM or shared memory.  256 TMEM columns and
// 80 KB of smem per CTA let two CTAs share an SM, so one CTA's softmax overlaps the other's loads and MMAs.
// (P_TMEM = false keeps P in swizzled smem instead: the cross-check variant used by the kernel tests.)
//
// Replaces MultiHeadAttentionLayer.forward's energy / softmax / matmul (reference amt_apc.py:349-368) for all
// four shapes on the path: encoder self (256x256), decoder cross (88 <- 256), decoder self (88x88),
// time-axis self (512x512).  head_dim = 64, 4 heads, no mask in the reference; keys past Lk (tile padding of
// the 88-key case) are masked here.
#pragma once
#include "common.cuh"

namespace etude {

struct AttnParams {
    int Lq, Lk;        // queries / keys per sequence
    int n_seq;
    int q_seq_stride;  // rows between consecutive sequences in the Q source (0 = every sequence shares one Q)
    int q_tiles;       // ceil(Lq / 128)
    int kb_rows;       // keys per KV block (box rows of the KV tensor map): 96 or 256
    int n_kv_blocks;   // ceil(Lk / kb_rows)
    int q_col0;        // column of head 0 of Q in the Q source matrix
    int k_col0, v_col0;  // columns of head 0 of K / V in the KV source matrix
    __nv_bfloat16* out;  // [n_seq * Lq, 256]: head h -> columns [64h, 64h+64)
    float scale_log2e;   // log2(e) / sqrt(head_dim)
    float* probs;        // optional fp32 [n_seq, 4, Lq, Lk] softmax probabilities (single KV block only)
};

constexpr int kAttnThreads = 128;
constexpr int kAttnQBytes = 128 * 64 * 2;
constexpr int kAttnKVBytes = 256 * 64 * 2;
template <bool P_TMEM>
__host__ __device__ constexpr size_t attn_smem_bytes() {
    return 1024 + kAttnQBytes + 2 * kAttnKVBytes + (P_TMEM ? 0 : 4 * kAttnQBytes) + 64;
}

template <bool P_TMEM>
__global__ void __launch_bounds__(kAttnThreads, P_TMEM ? 2 : 1)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv, const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + kAttnQBytes;
    uint8_t* sV = sK + kAttnKVBytes;
    uint8_t* sP = sV + kAttnKVBytes;  // 4 k-blocks of [128 rows x 64 keys] bf16, each K-major SW128
    uint64_t* bars = reinterpret_cast<uint64_t*>(P_TMEM ? sP : sP + 4 * kAttnQBytes);
    uint64_t* bar_load = bars;
    uint64_t* bar_s = bars + 1;
    uint64_t* bar_o = bars + 2;
    uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(bars + 3);

    const int tid = threadIdx.x, warp = tid >> 5;
    int bid = blockIdx.x;
    const int qt = bid % p.q_tiles;
    bid /= p.q_tiles;
    const int head = bid % kHeads;
    const int seq = bid / kHeads;

    if (tid == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_kv);
        mbar_init(bar_load, 1);
        mbar_init(bar_s, 1);
        mbar_init(bar_o, 1);
        mbar_fence_init();
    }
    constexpr uint32_t kTmemCols = P_TMEM ? 256 : 512;
    if (warp == 0) tmem_alloc(tmem_base_ptr, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_ptr;
    const uint32_t tmem_s = tmem_base;                          // S: [128 lanes x <=256 cols]; P (packed bf16) over cols [0,128)
    const uint32_t tmem_o = tmem_base + (P_TMEM ? 128 : 256);   // O: [128 lanes x 64 cols] (S cols 128.. are dead once P is written)
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;

    const int kv_bytes = p.kb_rows * 128;
    const int q_row0 = seq * p.q_seq_stride + qt * 128;
    const int kv_row0 = seq * p.Lk;
    const uint32_t idesc_s = make_idesc_bf16(128, p.kb_rows, 0, 0);
    const uint32_t idesc_o = make_idesc_bf16(128, kHeadDim, 0, 1);  // B = V is MN-major (d contiguous)

    float acc[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) acc[j] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    uint32_t ph_load = 0, ph_s = 0, ph_o = 0;
    float v[32];

    for (int blk = 0; blk < p.n_kv_blocks; ++blk) {
        if (tid == 0) {
            mbar_expect_tx(bar_load, (blk == 0 ? kAttnQBytes : 0) + 2 * kv_bytes);
            if (blk == 0) tma_load_2d(sQ, &tmap_q, bar_load, p.q_col0 + head * kHeadDim, q_row0);
            tma_load_2d(sK, &tmap_kv, bar_load, p.k_col0 + head * kHeadDim, kv_row0 + blk * p.kb_rows);
            tma_load_2d(sV, &tmap_kv, bar_load, p.v_col0 + head * kHeadDim, kv_row0 + blk * p.kb_rows);
            mbar_wait(bar_load, ph_load);
            tc_fence_after();
            // S = Q K^T : M = 128 queries, N = kb_rows keys, K = 64 (4 x UMMA_K 16)
            const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_bf16_ss(tmem_s, make_sw128_desc(qa + k * 32), make_sw128_desc(ka + k * 32), idesc_s, k != 0);
            tc_commit(bar_s);
        }
        ph_load ^= 1;
        mbar_wait(bar_s, ph_s);
        ph_s ^= 1;
        __syncwarp();
        tc_fence_after();

        const int keys_here = min(p.kb_rows, p.Lk - blk * p.kb_rows);  // valid keys in this block
        const int n_chunks = p.kb_rows / 32;                             // 3 or 8
        // pass 1: row maximum
        float m_blk = -INFINITY;
        for (int c = 0; c < n_chunks; ++c) {
            tmem_ld32(tmem_s + lane_off + c * 32, v);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (c * 32 + j < keys_here) m_blk = fmaxf(m_blk, v[j]);
        }
        const float m_new = fmaxf(m_run, m_blk);
        const float alpha = exp2f((m_run - m_new) * p.scale_log2e);  // 0 on the first block (m_run = -inf)
        const float m_sc = m_new * p.scale_log2e;
        // pass 2: p = exp2(s*scale - m*scale), row sum, bf16 P into swizzled smem (A operand of P V)
        float l_blk = 0.f;
        for (int c = 0; c < n_chunks; ++c) {
            tmem_ld32(tmem_s + lane_off + c * 32, v);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float e = (c * 32 + j < keys_here) ? exp2f(v[j] * p.scale_log2e - m_sc) : 0.f;
                v[j] = e;
                l_blk += e;
            }
            if constexpr (P_TMEM) {
                // keys [32c, 32c+32) of this row -> 16 packed columns [16c, 16c+16) (already-consumed S columns)
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
                tmem_st16(tmem_s + lane_off + c * 16, pk);
            } else {
                // keys [32c, 32c+32) of row tid: k-block c/2, 16-byte chunks (c%2)*4 .. +3
                uint8_t* prow = sP + (c >> 1) * kAttnQBytes + tid * 128;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int chunk = (c & 1) * 4 + g;
                    uint4 pk;
                    pk.x = pack_bf16x2(v[g * 8 + 0], v[g * 8 + 1]);
                    pk.y = pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]);
                    pk.z = pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]);
                    pk.w = pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]);
                    *reinterpret_cast<uint4*>(prow + ((chunk ^ (tid & 7)) << 4)) = pk;
                }
            }
            if (p.probs != nullptr && p.n_kv_blocks == 1) {
                // un-normalised here; normalised in place below once the row sum is known
                const int qrow = qt * 128 + tid;
                if (qrow < p.Lq) {
                    float* dst = p.probs + (((size_t)seq * kHeads + head) * p.Lq + qrow) * p.Lk + c * 32;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c * 32 + j < p.Lk) dst[j] = v[j];
                }
            }
        }
        l_run = l_run * alpha + l_blk;
        m_run = m_new;
        if constexpr (P_TMEM) tc_wait_st();
        else fence_async_smem();  // generic-proxy writes of P -> visible to the tensor core (async proxy)
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            // O_blk = P V : M = 128, N = 64, K = kb_rows keys (16 per MMA = 8 packed TMEM columns of P)
            const uint32_t pa = smem_u32(sP), va = smem_u32(sV);
            const int ksteps = p.kb_rows / 16;
            for (int j = 0; j < ksteps; ++j) {
                if constexpr (P_TMEM)
                    umma_bf16_ts(tmem_o, tmem_s + j * 8, make_sw128_desc(va + j * 2048, 8192), idesc_o, j != 0);
                else
                    umma_bf16_ss(tmem_o, make_sw128_desc(pa + (j >> 2) * kAttnQBytes + (j & 3) * 32),
                                 make_sw128_desc(va + j * 2048, 8192), idesc_o, j != 0);
            }
            tc_commit(bar_o);
        }
        mbar_wait(bar_o, ph_o);
        ph_o ^= 1;
        __syncwarp();
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            tmem_ld32(tmem_o + lane_off + c * 32, v);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[c * 32 + j] = acc[c * 32 + j] * alpha + v[j];
        }
        tc_fence_before();
        __syncthreads();  // S / P / K / V may be overwritten by the next block
    }

    const int qrow = qt * 128 + tid;
    if (qrow < p.Lq) {
        const float inv = 1.f / l_run;
        __nv_bfloat16* dst = p.out + (size_t)(seq * p.Lq + qrow) * kHid + head * kHeadDim;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            uint4 pk;
            pk.x = pack_bf16x2(acc[g * 8 + 0] * inv, acc[g * 8 + 1] * inv);
            pk.y = pack_bf16x2(acc[g * 8 + 2] * inv, acc[g * 8 + 3] * inv);
            pk.z = pack_bf16x2(acc[g * 8 + 4] * inv, acc[g * 8 + 5] * inv);
            pk.w = pack_bf16x2(acc[g * 8 + 6] * inv, acc[g * 8 + 7] * inv);
            *reinterpret_cast<uint4*>(dst + g * 8) = pk;
        }
        if (p.probs != nullptr && p.n_kv_blocks == 1) {
            float* dst_p = p.probs + (((size_t)seq * kHeads + head) * p.Lq + qrow) * p.Lk;
            for (int j = 0; j < p.Lk; ++j) dst_p[j] *= inv;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace etude
