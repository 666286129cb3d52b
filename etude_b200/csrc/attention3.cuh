// Fused multi-head attention, third generation (attention2.cuh is the second; ETUDE_ATTN_V2=1 selects it as the
// cross-check variant).  Same math and interface: MultiHeadAttentionLayer.forward's energy / softmax / matmul
// (reference amt_apc.py:349-368) for the four shapes on the path, head_dim 64, 4 heads, no mask in the reference.
//
// What changed against attention2.cuh (profiles/r1c_ncu_attn_*): there the 8 softmax warps were two per query row
// (half the keys each), exchanged their maxima through smem + a named barrier, and all worked on the SAME S tile, so on
// every SM sub-partition both softmax warps sat in the same phase: the MUFU pipe (ex2) idled through every max pass,
// every tcgen05.ld wait and every barrier (41 % busy), and the kernel ran at ~5500 clk per 128 x 256 tile against a
// 2048 clk MUFU floor.  Here
//   * a softmax thread owns a whole query row (no exchange, no barrier) and the 8 warps form TWO GROUPS that work on
//     different tiles (group = TMEM buffer = tile parity), so the two warps of a sub-partition are out of phase: one
//     is in its max pass / waiting on TMEM while the other feeds the MUFU pipe;
//   * softmax is stateless per KV block: every (query tile, KV block) is normalised against its own maximum and the
//     drain warps combine the (at most two) blocks of a row exactly, flash-decoding style
//         O = (O0 a0 + O1 a1) / (l0 a0 + l1 a1),  a_j = 2^(m_j - max(m0, m1));
//   * tcgen05.ld is software pipelined against the arithmetic (two register chunks in flight);
//   * P (bf16) is written back contiguously over the S columns the thread has already consumed (columns [0, KB/2)),
//     the O accumulator sits in columns [KB/2, KB/2 + 64) of the same buffer.
//
// One CTA per SM walks work items (sequence, head); TMA warp, MMA warp (warp-uniform issue), 4 drain warps (one per
// TMEM lane quarter, warps 4-7), 8 softmax warps (warps 8-15: group g = (warp - 8) >> 2, lane quarter warp & 3).
#pragma once
#include "attention2.cuh"
#include "common.cuh"

namespace etude {

constexpr int kAttn3Threads = 16 * 32;  // 4 warpgroups: {TMA, MMA, 2 idle}, drain, softmax group 0, softmax group 1
constexpr int kA3StatsBytes = (2 * 128 + 2 * 128) * 4;  // m[2][128] (scaled, log2 domain), l[2][128]
constexpr size_t kAttn3SmemBytes = 1024 + kA2KvSlots * kA2KvSlotBytes + kA2QSlots * kA2QSlotBytes + kA3StatsBytes + 256;

__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// Register budget (64 K per SM, 16 warps): the kernel starts at 128 per thread; the TMA/MMA warpgroup gives back down to 40,
// the drain warpgroup takes 136 and each softmax warpgroup 168 (40 + 136 + 2 x 168 = 512 = 4 x 128) with setmaxnreg, so a softmax thread keeps
// two 32-column S chunks, the packed P chunk and its running statistics in registers without spilling.  Each role branch
// issues its own setmaxnreg: code reachable from a .dec is compiled against the reduced budget.
template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

template <int KB>
__global__ void __launch_bounds__(kAttn3Threads, 1)
attention3_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv, const Attn2Params p) {
    constexpr int NCH = KB / 32;                  // 32-column chunks of an S row: 8 or 3
    constexpr int KSTEPS = KB / 16;               // UMMA_K steps of P V
    constexpr int O_COL = (KB == 256) ? 128 : 64;   // O accumulator columns inside the buffer (P occupies [0, KB/2))
    constexpr uint32_t KV_BYTES = KB * 128;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sKV = smem;
    uint8_t* sQ = sKV + kA2KvSlots * kA2KvSlotBytes;
    float* s_m = reinterpret_cast<float*>(sQ + kA2QSlots * kA2QSlotBytes);  // [2 buf][128] block maximum * scale * log2(e)
    float* s_l = s_m + 2 * 128;                                             // [2 buf][128] block sum of 2^(s - m)
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_l + 2 * 128);
    uint64_t* kv_full = bars;                         // [5]
    uint64_t* kv_empty = kv_full + kA2KvSlots;        // [5]
    uint64_t* q_full = kv_empty + kA2KvSlots;         // [3]
    uint64_t* q_empty = q_full + kA2QSlots;           // [3]
    uint64_t* s_full = q_empty + kA2QSlots;           // [2]  MMA -> softmax group b
    uint64_t* p_full = s_full + 2;                    // [2]  softmax group b (4 warps) -> MMA, drain
    uint64_t* o_full = p_full + 2;                    // [2]  MMA -> drain
    uint64_t* buf_free = o_full + 2;                  // [2]  drain (4 warps) -> MMA
    uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(buf_free + 2);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform
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
            mbar_init(&p_full[i], 4);
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

    if (warp < 4) {
      reg_dec<56>();  // one setmaxnreg for the whole TMA / MMA warpgroup (warps 2 and 3 only take part in the barriers)
      if (warp == 0) {
        // ===================================================== TMA producer (warp-uniform loop, one elected lane issues)
        const bool leader = elect_one();
        uint32_t kvc = 0, qc = 0;  // ring counters
        auto load_kv = [&](int col, int row) {
            const uint32_t s = kvc % kA2KvSlots, r = kvc / kA2KvSlots;
            mbar_wait_inl(&kv_empty[s], (r & 1) ^ 1);
            if (leader) {
                mbar_expect_tx(&kv_full[s], KV_BYTES);
                tma_load_2d(sKV + s * kA2KvSlotBytes, &tmap_kv, &kv_full[s], col, row);
            }
            ++kvc;
        };
        auto load_q = [&](int col, int row) {
            const uint32_t s = qc % kA2QSlots, r = qc / kA2QSlots;
            mbar_wait_inl(&q_empty[s], (r & 1) ^ 1);
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
        const uint64_t q_desc0 = make_sw128_desc(smem_u32(sQ));
        const uint64_t k_desc0 = make_sw128_desc(smem_u32(sKV));
        const uint64_t v_desc0 = make_sw128_desc(smem_u32(sKV), 8192);
        auto issue_s = [&](int g) {
            const int il = g / NT, n = g % NT, t = n / p.NKV, j = n % p.NKV;
            const uint32_t kc = (uint32_t)(il * p.NKV + j) * 2, qc = (uint32_t)(il * p.QT + t);
            const uint32_t ks = kc % kA2KvSlots, qs = qc % kA2QSlots;
            const int b = g & 1;
            mbar_wait_inl(&kv_full[ks], (kc / kA2KvSlots) & 1);
            mbar_wait_inl(&q_full[qs], (qc / kA2QSlots) & 1);
            mbar_wait_inl(&buf_free[b], ((g >> 1) & 1) ^ 1);
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
            mbar_wait_inl(&kv_full[vs], (vc / kA2KvSlots) & 1);
            mbar_wait_inl(&p_full[b], (g >> 1) & 1);
            tc_fence_after();
            if (leader) {
                const uint64_t vd = v_desc0 + (uint64_t)(vs * (kA2KvSlotBytes >> 4));
                const uint32_t tmem_buf = tmem_base + b * 256;
#pragma unroll
                for (int s = 0; s < KSTEPS; ++s)  // P: bf16 pairs, 8 columns per K = 16; V: 16 keys = 2048 B further
                    umma_bf16_ts(tmem_buf + O_COL, tmem_buf + s * 8, vd + (uint64_t)(s * 128), idesc_o, s != 0);
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
      }
    } else if (warp < 8) {
        // ===================================================== drain warps: O -> registers -> combine KV blocks -> bf16 -> HBM
        reg_inc<136>();
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        float acc[64];
        float v[32];
        float m0 = 0.f, l0 = 0.f;
        for (int g = 0; g < G; ++g) {
            const int il = g / NT, n = g % NT, t = n / p.NKV, j = n % p.NKV;
            const int b = g & 1;
            const uint32_t ph = (g >> 1) & 1;
            mbar_wait_inl(&p_full[b], ph);  // softmax statistics of this tile are visible
            mbar_wait_inl(&o_full[b], ph);
            __syncwarp();
            tc_fence_after();
            const float mj = s_m[b * 128 + row], lj = s_l[b * 128 + row];
            const uint32_t tmem_o = tmem_base + b * 256 + O_COL + lane_off;
            float inv;
            if (j == 0) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    tmem_ld32(tmem_o + c * 32, v);
                    tc_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[c * 32 + i] = v[i];
                }
                m0 = mj; l0 = lj;
                inv = 1.f / lj;
            } else {  // second KV block of this query row: exact combination of two independently normalised blocks
                const float mm = fmaxf(m0, mj);
                const float a0 = ex2_approx(m0 - mm), a1 = ex2_approx(mj - mm);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    tmem_ld32(tmem_o + c * 32, v);
                    tc_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[c * 32 + i] = fmaf(acc[c * 32 + i], a0, v[i] * a1);
                }
                inv = 1.f / fmaf(l0, a0, lj * a1);
            }
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
        // ===================================================== softmax warps: group = buffer = tile parity, one thread per query row
        reg_inc<160>();
        const int grp = (warp - 8) >> 2;
        const int q = warp & 3;   // TMEM lane quarter of this warp
        const int row = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const uint32_t tmem_s = tmem_base + grp * 256 + lane_off;
        const float scale = p.scale_log2e;
        float va[32], vb[32];
        for (int g = grp; g < G; g += 2) {
            const int il = g / NT, n = g % NT, t = n / p.NKV, j = n % p.NKV;
            mbar_wait_inl(&s_full[grp], (g >> 1) & 1);
            __syncwarp();
            tc_fence_after();
            const int keys_here = min(KB, p.Lk - j * KB);
            // ---- pass 1: row maximum (tcgen05.ld of chunk c + 1 in flight while chunk c is reduced)
            float m0 = -INFINITY, m1 = -INFINITY;
            tmem_ld32(tmem_s, va);
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                float* cur = (c & 1) ? vb : va;
                float* nxt = (c & 1) ? va : vb;
                tc_wait_ld();
                if (c + 1 < NCH) tmem_ld32(tmem_s + (c + 1) * 32, nxt);
                if (KB == 256 || (c + 1) * 32 <= keys_here) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        m0 = fmax3(m0, cur[i], cur[i + 1]);
                        m1 = fmax3(m1, cur[i + 2], cur[i + 3]);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (c * 32 + i < keys_here) m0 = fmaxf(m0, cur[i]);
                }
            }
            const float m_sc = fmaxf(m0, m1) * scale;
            // ---- pass 2: p = 2^(s * scale - m * scale) -> bf16 P over the S columns already consumed; row sum
            float l0 = 0.f, l1 = 0.f;
            float* probs_row = nullptr;
            if (p.probs != nullptr) {
                const int item = blockIdx.x + il * gridDim.x;
                const int qrow = t * 128 + row;
                if (qrow < p.Lq) probs_row = p.probs + (((size_t)(item >> 2) * kHeads + (item & 3)) * p.Lq + qrow) * p.Lk;
            }
            tmem_ld32(tmem_s, va);
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                float* cur = (c & 1) ? vb : va;
                float* nxt = (c & 1) ? va : vb;
                tc_wait_ld();
                if (c + 1 < NCH) tmem_ld32(tmem_s + (c + 1) * 32, nxt);
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float e0 = ex2_approx(fmaf(cur[i], scale, -m_sc));
                    float e1 = ex2_approx(fmaf(cur[i + 1], scale, -m_sc));
                    if (KB != 256) {
                        if (c * 32 + i >= keys_here) e0 = 0.f;
                        if (c * 32 + i + 1 >= keys_here) e1 = 0.f;
                    }
                    cur[i] = e0; cur[i + 1] = e1;
                    l0 += e0; l1 += e1;
                }
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(cur[2 * i], cur[2 * i + 1]);
                // P chunk c -> columns [16 c, 16 c + 16): S columns below 32 (c + 1) are consumed; the prefetched chunk c + 1
                // (columns [32 (c + 1), 32 (c + 2))) is already in registers or in flight from columns >= 32 > 16 c + 16 for c >= 1,
                // and for c == 0 the write covers [0, 16) while chunk 1 reads [32, 64)
                tmem_st16(tmem_s + c * 16, pk);
                if (probs_row != nullptr) {  // un-normalised here; normalised in place below
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int key = c * 32 + i;
                        if (key < p.Lk) probs_row[key] = cur[i];
                    }
                }
            }
            const float l = l0 + l1;
            s_m[grp * 128 + row] = m_sc;
            s_l[grp * 128 + row] = l;
            if (probs_row != nullptr) {  // 9-tuple attention output (single KV block)
                const float inv = 1.f / l;
                for (int key = 0; key < p.Lk; ++key) probs_row[key] *= inv;
            }
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[grp]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace etude
