// Persistent, warp-specialised fused multi-head attention on tcgen05 (second generation; attention.cuh is the
// first-generation one-tile-per-CTA kernel, since removed).
//
// Replaces MultiHeadAttentionLayer.forward's energy / softmax / matmul (reference amt_apc.py:349-368) for the four
// shapes on the path: encoder self (256x256), decoder cross (88 <- 256), decoder self (88x88), time-axis self
// (512x512).  head_dim 64, 4 heads, no mask in the reference (keys past Lk in the 96-row box of the 88-key case are
// masked here).
//
// One CTA per SM walks work items (sequence, head).  Per item the TMA warp loads K/V blocks and Q tiles once into
// smem rings; the MMA thread streams "S tiles" (128 queries x KB keys) through TWO 256-column TMEM buffers:
//     S = Q K^T (SS MMA)  ->  softmax warps: fp32 max / exp2 / sum, P (bf16) written back over S in TMEM
//                         ->  O = P V (TS MMA, A operand = P in TMEM) into free columns of the same buffer
//                         ->  drain warps: O -> registers (online-softmax rescale across KV blocks) -> bf16 -> HBM
// so the softmax of tile g overlaps S of tile g+1 and P V / drain of tile g-1; scores never leave the SM.
//
// Warp roles (14 warps): 0 = TMA producer, 1 = MMA issuer + TMEM allocator, 2..5 = drain (one per TMEM lane
// quarter), 6..13 = softmax (two warps per lane quarter: each thread owns half of the keys of one query row).
#pragma once
#include "common.cuh"

namespace etude {

constexpr int kAttn2Threads = 14 * 32;
constexpr int kA2KvSlots = 5;
constexpr int kA2QSlots = 3;
constexpr int kA2KvSlotBytes = 256 * 64 * 2;  // one K or V block of up to 256 keys
constexpr int kA2QSlotBytes = 128 * 64 * 2;
constexpr int kA2StatsBytes = (2 * 2 * 128 + 2 * 128 + 2 * 2 * 128) * 4;  // xm[2][2][128], alpha[2][128], l[2][2][128]
constexpr size_t kAttn2SmemBytes = 1024 + kA2KvSlots * kA2KvSlotBytes + kA2QSlots * kA2QSlotBytes + kA2StatsBytes + 256;

struct Attn2Params {
    int Lq, Lk;
    int n_items;       // n_seq * 4 heads
    int q_seq_stride;  // rows between consecutive sequences in the Q source (0 = every sequence shares one Q)
    int QT;            // query tiles per item: ceil(Lq / 128)
    int NKV;           // KV blocks per item: ceil(Lk / KB)
    int q_col0, k_col0, v_col0;
    __nv_bfloat16* out;  // [n_seq * Lq, 256]
    float scale_log2e;
    float* probs;        // optional fp32 [n_seq, 4, Lq, Lk] (NKV == 1 only)
    long long* trace;    // optional debug timeline (attention4: 8 roles x 192 tiles x (id, clock64)), CTA 0 only
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 32 lanes x 16 columns of fp32
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// KB = keys per KV block (box rows of the KV tensor map): 256 (Lk = 256 / 512) or 96 (Lk = 88).
template <int KB>
__global__ void __launch_bounds__(kAttn2Threads, 1)
attention2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv, const Attn2Params p) {
    constexpr int HALF = KB / 2;               // keys per softmax thread
    constexpr int CH = (KB == 256) ? 32 : 16;  // columns per tcgen05.ld
    constexpr int NCH = HALF / CH;             // 4 or 3 chunks per half
    constexpr int KSTEPS = KB / 16;            // UMMA_K steps of P V
    constexpr int O_COL = (KB == 256) ? 64 : 96;  // O accumulator columns inside the buffer (see header comment)
    constexpr uint32_t KV_BYTES = KB * 128;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sKV = smem;
    uint8_t* sQ = sKV + kA2KvSlots * kA2KvSlotBytes;
    float* s_xm = reinterpret_cast<float*>(sQ + kA2QSlots * kA2QSlotBytes);  // [2 buf][2 half][128]
    float* s_alpha = s_xm + 2 * 2 * 128;                                     // [2 buf][128]
    float* s_l = s_alpha + 2 * 128;                                          // [2 buf][2 half][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_l + 2 * 2 * 128);
    uint64_t* kv_full = bars;                         // [5]
    uint64_t* kv_empty = kv_full + kA2KvSlots;        // [5]
    uint64_t* q_full = kv_empty + kA2KvSlots;         // [3]
    uint64_t* q_empty = q_full + kA2QSlots;           // [3]
    uint64_t* s_full = q_empty + kA2QSlots;           // [2]  MMA -> softmax
    uint64_t* p_full = s_full + 2;                    // [2]  softmax (8 warps) -> MMA, drain
    uint64_t* o_full = p_full + 2;                    // [2]  MMA -> drain
    uint64_t* buf_free = o_full + 2;                  // [2]  drain (4 warps) -> MMA
    uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(buf_free + 2);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform (see chain3.cuh on why it matters)
    const int lane = threadIdx.x & 31;
    const int NT = p.QT * p.NKV;  // S tiles per item
    const int my_items = ((int)blockIdx.x < p.n_items) ? (p.n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int G = my_items * NT;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_kv);
        for (int i = 0; i < kA2KvSlots; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
        for (int i = 0; i < kA2QSlots; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], 8);
            mbar_init(&o_full[i], 1);
            mbar_init(&buf_free[i], 4);
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_base_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_ptr;

    if (warp == 0) {
        // ===================================================== TMA producer (warp-uniform loop, one elected lane issues)
        const bool leader = elect_one();
        uint32_t kvc = 0, qc = 0;  // ring counters
        auto load_kv = [&](int col, int row) {
            const uint32_t s = kvc % kA2KvSlots, r = kvc / kA2KvSlots;
            mbar_wait(&kv_empty[s], (r & 1) ^ 1);
            if (leader) {
                mbar_expect_tx(&kv_full[s], KV_BYTES);
                tma_load_2d(sKV + s * kA2KvSlotBytes, &tmap_kv, &kv_full[s], col, row);
            }
            ++kvc;
        };
        auto load_q = [&](int col, int row) {
            const uint32_t s = qc % kA2QSlots, r = qc / kA2QSlots;
            mbar_wait(&q_empty[s], (r & 1) ^ 1);
            if (leader) {
                mbar_expect_tx(&q_full[s], kA2QSlotBytes);
                tma_load_2d(sQ + s * kA2QSlotBytes, &tmap_q, &q_full[s], col, row);
            }
            ++qc;
        };
        for (int il = 0; il < my_items; ++il) {
            const int item = blockIdx.x + il * gridDim.x;
            const int head = item & 3, seq = item >> 2;
            const int q_row0 = seq * p.q_seq_stride, kv_row0 = seq * p.Lk;
            const int qcol = p.q_col0 + head * kHeadDim, kcol = p.k_col0 + head * kHeadDim, vcol = p.v_col0 + head * kHeadDim;
            // issue order = first-use order of the MMA warp's (t, j) t-major schedule
            load_kv(kcol, kv_row0);
            load_q(qcol, q_row0);
            load_kv(vcol, kv_row0);
            for (int j = 1; j < p.NKV; ++j) {
                load_kv(kcol, kv_row0 + j * KB);
                load_kv(vcol, kv_row0 + j * KB);
            }
            for (int t = 1; t < p.QT; ++t) load_q(qcol, q_row0 + t * 128);
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer (warp-uniform loop, one elected lane issues)
        const bool leader = elect_one();
        const uint32_t idesc_s = make_idesc_bf16(128, KB, 0, 0);
        const uint32_t idesc_o = make_idesc_bf16(128, kHeadDim, 0, 1);  // B = V is MN-major (d contiguous)
        // descriptors = base + adds: Q slots 16 KB (+1024), K/V slots 32 KB (+2048), K = 16 step: +2 (K-major K) / +128 (MN-major V, 2048 B)
        const uint64_t q_desc0 = make_sw128_desc(smem_u32(sQ));
        const uint64_t k_desc0 = make_sw128_desc(smem_u32(sKV));
        const uint64_t v_desc0 = make_sw128_desc(smem_u32(sKV), 8192);
        auto issue_s = [&](int g) {
            const int il = g / NT, n = g % NT, t = n / p.NKV, j = n % p.NKV;
            const uint32_t kc = (uint32_t)(il * p.NKV + j) * 2, qc = (uint32_t)(il * p.QT + t);
            const uint32_t ks = kc % kA2KvSlots, qs = qc % kA2QSlots;
            const int b = g & 1;
            mbar_wait(&kv_full[ks], (kc / kA2KvSlots) & 1);
            mbar_wait(&q_full[qs], (qc / kA2QSlots) & 1);
            mbar_wait(&buf_free[b], ((g >> 1) & 1) ^ 1);
            tc_fence_after();
            if (leader) {
                const uint64_t qd = q_desc0 + (uint64_t)(qs * (kA2QSlotBytes >> 4)), kd = k_desc0 + (uint64_t)(ks * (kA2KvSlotBytes >> 4));
                const uint32_t tmem_s = tmem_base + b * 256;
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_s, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
                tc_commit(&s_full[b]);
                if (j == p.NKV - 1) tc_commit(&q_empty[qs]);  // last S that reads this Q tile
            }
            __syncwarp();
        };
        auto issue_pv = [&](int g) {
            const int il = g / NT, n = g % NT, t = n / p.NKV, j = n % p.NKV;
            const uint32_t kc = (uint32_t)(il * p.NKV + j) * 2, vc = kc + 1;
            const uint32_t ks = kc % kA2KvSlots, vs = vc % kA2KvSlots;
            const int b = g & 1;
            mbar_wait(&kv_full[vs], (vc / kA2KvSlots) & 1);
            mbar_wait(&p_full[b], (g >> 1) & 1);
            tc_fence_after();
            if (leader) {
                const uint64_t vd = v_desc0 + (uint64_t)(vs * (kA2KvSlotBytes >> 4));
                const uint32_t tmem_buf = tmem_base + b * 256;
#pragma unroll
                for (int s = 0; s < KSTEPS; ++s) {
                    const uint32_t pcol = (s < KSTEPS / 2) ? s * 8 : HALF + (s - KSTEPS / 2) * 8;  // P of half 0 / half 1
                    umma_bf16_ts(tmem_buf + O_COL, tmem_buf + pcol, vd + (uint64_t)(s * 128), idesc_o, s != 0);
                }
                tc_commit(&o_full[b]);
                if (t == p.QT - 1) {  // last use of this K / V block
                    tc_commit(&kv_empty[ks]);
                    tc_commit(&kv_empty[vs]);
                }
            }
            __syncwarp();
        };
        if (G > 0) issue_s(0);
        for (int g = 0; g < G; ++g) {
            if (g + 1 < G) issue_s(g + 1);
            issue_pv(g);
        }
    } else if (warp < 6) {
        // ===================================================== drain warps: O -> registers -> bf16 -> HBM
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        float acc[64];
        float v[32];
        for (int g = 0; g < G; ++g) {
            const int il = g / NT, n = g % NT, t = n / p.NKV, j = n % p.NKV;
            const int b = g & 1;
            const uint32_t ph = (g >> 1) & 1;
            mbar_wait(&p_full[b], ph);  // softmax statistics of this tile are visible
            mbar_wait(&o_full[b], ph);
            __syncwarp();
            tc_fence_after();
            const float alpha = s_alpha[b * 128 + row];
            const uint32_t tmem_o = tmem_base + b * 256 + O_COL + lane_off;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                tmem_ld32(tmem_o + c * 32, v);
                tc_wait_ld();
                if (j == 0) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[c * 32 + i] = v[i];
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[c * 32 + i] = fmaf(acc[c * 32 + i], alpha, v[i]);
                }
            }
            float inv = 0.f;
            if (j == p.NKV - 1) inv = 1.f / (s_l[(b * 2 + 0) * 128 + row] + s_l[(b * 2 + 1) * 128 + row]);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&buf_free[b]);
            if (j == p.NKV - 1) {
                const int item = blockIdx.x + il * gridDim.x;
                const int head = item & 3, seq = item >> 2;
                const int qrow = t * 128 + row;
                if (qrow < p.Lq) {
                    __nv_bfloat16* dst = p.out + (size_t)(seq * p.Lq + qrow) * kHid + head * kHeadDim;
#pragma unroll
                    for (int gq = 0; gq < 8; ++gq) {
                        uint4 pk;
                        pk.x = pack_bf16x2(acc[gq * 8 + 0] * inv, acc[gq * 8 + 1] * inv);
                        pk.y = pack_bf16x2(acc[gq * 8 + 2] * inv, acc[gq * 8 + 3] * inv);
                        pk.z = pack_bf16x2(acc[gq * 8 + 4] * inv, acc[gq * 8 + 5] * inv);
                        pk.w = pack_bf16x2(acc[gq * 8 + 6] * inv, acc[gq * 8 + 7] * inv);
                        *reinterpret_cast<uint4*>(dst + gq * 8) = pk;
                    }
                }
            }
        }
    } else {
        // ===================================================== softmax warps
        const int sw = warp - 6;
        const int q = warp & 3;   // TMEM lane quarter of this warp
        const int hf = sw >> 2;   // which half of the keys this thread owns
        const int row = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int bar_id = 1 + q;  // named barrier shared by the two warps of a lane quarter
        float m_run = -INFINITY, l_run = 0.f;
        float v[CH];
        for (int g = 0; g < G; ++g) {
            const int il = g / NT, n = g % NT, t = n / p.NKV, j = n % p.NKV;
            const int b = g & 1;
            mbar_wait(&s_full[b], (g >> 1) & 1);
            __syncwarp();
            tc_fence_after();
            if (j == 0) { m_run = -INFINITY; l_run = 0.f; }
            const int keys_here = min(KB, p.Lk - j * KB);
            const uint32_t tmem_s = tmem_base + b * 256 + lane_off + hf * HALF;
            // ---- pass 1: maximum over this thread's half of the keys
            float m_part = -INFINITY;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                if constexpr (CH == 32) tmem_ld32(tmem_s + c * CH, v); else tmem_ld16(tmem_s + c * CH, v);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < CH; ++i)
                    if (KB == 256 || hf * HALF + c * CH + i < keys_here) m_part = fmaxf(m_part, v[i]);
            }
            s_xm[(b * 2 + hf) * 128 + row] = m_part;
            asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
            const float m_blk = fmaxf(m_part, s_xm[(b * 2 + (hf ^ 1)) * 128 + row]);
            const float m_new = fmaxf(m_run, m_blk);
            const float alpha = ex2_approx((m_run - m_new) * p.scale_log2e);  // 0 on the first block (m_run = -inf)
            const float m_sc = m_new * p.scale_log2e;
            m_run = m_new;
            // ---- pass 2: p = exp2(s*scale - m*scale) -> bf16 P over the already-consumed S columns of this half
            float l_blk = 0.f;
            float* probs_row = nullptr;
            if (p.probs != nullptr) {
                const int item = blockIdx.x + il * gridDim.x;
                const int qrow = t * 128 + row;
                if (qrow < p.Lq) probs_row = p.probs + (((size_t)(item >> 2) * kHeads + (item & 3)) * p.Lq + qrow) * p.Lk;
            }
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                if constexpr (CH == 32) tmem_ld32(tmem_s + c * CH, v); else tmem_ld16(tmem_s + c * CH, v);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    float e = ex2_approx(fmaf(v[i], p.scale_log2e, -m_sc));
                    if (KB != 256 && hf * HALF + c * CH + i >= keys_here) e = 0.f;
                    v[i] = e;
                    l_blk += e;
                }
                uint32_t pk[CH / 2];
#pragma unroll
                for (int i = 0; i < CH / 2; ++i) pk[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
                if constexpr (CH == 32) tmem_st16(tmem_s + c * (CH / 2), pk); else tmem_st8(tmem_s + c * (CH / 2), pk);
                if (probs_row != nullptr) {  // un-normalised here; normalised in place below
#pragma unroll
                    for (int i = 0; i < CH; ++i) {
                        const int key = hf * HALF + c * CH + i;
                        if (key < p.Lk) probs_row[key] = v[i];
                    }
                }
            }
            l_run = l_run * alpha + l_blk;
            if (hf == 0) s_alpha[b * 128 + row] = alpha;
            if (j == p.NKV - 1) s_l[(b * 2 + hf) * 128 + row] = l_run;
            if (p.probs != nullptr) {  // 9-tuple attention output (single KV block): needs the full row sum now
                asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
                if (probs_row != nullptr) {
                    const float inv = 1.f / (l_run + s_l[(b * 2 + (hf ^ 1)) * 128 + row]);
                    for (int i = 0; i < HALF; ++i) {
                        const int key = hf * HALF + i;
                        if (key < p.Lk) probs_row[key] *= inv;
                    }
                }
            }
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[b]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace etude
