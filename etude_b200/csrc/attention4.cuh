// Fused multi-head attention, fourth generation: attention3.cuh with KV blocks of 128 keys, FOUR TMEM buffers, three
// softmax groups and two drain groups.
//
// Why (profiles/r1e_ncu_attn3_enc_*): the MUFU pipe is the floor of this kernel (one ex2 per score: 2048 clk per
// 128 x 256 scores per SM) and was 47 % busy at a tile period of 4900 clk.  A softmax warp spends its time in phases that do
// not touch the MUFU (waiting for S through P V -> drain -> next S: 30 %; the maximum pass; exposed tcgen05.ld latency), and
// with two softmax warps per SM sub-partition both are often in such a phase at once.  Here the S tile is 128 x 128
// (KB = 128; 96 for the 88-key shapes), a buffer is 128 TMEM columns (S [0, KB) -> P bf16 [0, KB/2), O [64, 128)), there
// are four buffers, three softmax groups (tile g -> group g % 3) so that every sub-partition hosts three softmax warps in
// different phases (a fourth group measured no gain), S and P V are issued by separate warps (no head-of-line blocking
// between the two barriers), and two drain warpgroups.  Softmax stays stateless per KV block; the drain warps fold any number of
// blocks of a query tile exactly (running maximum m, sum l, accumulator O: flash-decoding combination).
// Same math and interface as attention2/3: MultiHeadAttentionLayer.forward's energy / softmax / matmul (reference
// amt_apc.py:349-368), head_dim 64, 4 heads, no mask.  The probabilities output (9-tuple API only) stays with attention2.
//
// One CTA per SM walks work items (sequence, head); warp 0 TMA, warp 1 S-MMA issue, warp 2 PV-MMA issue, warps 4-11 drain (one
// per TMEM lane quarter) x 2 groups alternating query tiles, warps 12-23 softmax (3 groups, tile g -> group g % 3, lane
// quarter warp & 3, one thread per query row).  Registers (768 threads start at 80, setmaxnreg): TMA/MMA warpgroup 40,
// drain 2 x 112, softmax 3 x 72.
#pragma once
#include "attention2.cuh"
#include "common.cuh"

namespace etude {

constexpr int kAttn4Threads = 24 * 32;  // 6 warpgroups: {TMA, S issue, PV issue, idle}, drain x 2, softmax x 3
constexpr int kA4Bufs = 4;               // TMEM buffers of 128 columns
constexpr int kA4SoftmaxGroups = 3;      // softmax warpgroups (tile g -> group g % 3, buffer g & 3)
constexpr int kA4KvSlots = 10;           // K / V blocks of up to 128 keys
constexpr int kA4QSlots = 3;
constexpr int kA4StageBytes = 0;
constexpr int kA4KvSlotBytes = 128 * 64 * 2;
constexpr int kA4StatsBytes = 2 * kA4Bufs * 128 * 4;  // m[4][128] (scaled, log2 domain), l[4][128]
constexpr size_t kAttn4SmemBytes = 1024 + kA4KvSlots * kA4KvSlotBytes + kA4QSlots * kA2QSlotBytes + kA4StageBytes + kA4StatsBytes + 512;

// The drain warps share their sub-partition's MUFU pipe with four softmax warps that keep it saturated (by design), so a
// MUFU instruction on the drain's dependency chain waits behind a few hundred clocks of queued ex2.  The drain therefore
// uses none: block references are INTEGERS in the log2 domain (softmax subtracts ceil(max)), so the combination factors
// 2^(r_a - r_b) are exact exponent-field constructions, and 1 / l is three Newton steps from a bit-trick seed (FMA pipe).
__device__ __forceinline__ float pow2_int(float d) {  // 2^d for an integer-valued d <= 0
    const int e = (int)d;
    return e < -126 ? 0.f : __int_as_float((e + 127) << 23);
}
__device__ __forceinline__ float rcp_fma(float x) {   // 1 / x for a positive normal x, relative error < 1e-7
    float y = __int_as_float(0x7EF311C7 - __float_as_int(x));
    y = y * fmaf(-x, y, 2.f);
    y = y * fmaf(-x, y, 2.f);
    y = y * fmaf(-x, y, 2.f);
    return y;
}

#define A4_TRACE(role, g)                                                                         \
    do {                                                                                          \
        if (p.trace != nullptr && blockIdx.x == 0 && lane == 0 && (g) < 192) {                    \
            p.trace[((role) * 192 + (g)) * 2] = (g);                                              \
            p.trace[((role) * 192 + (g)) * 2 + 1] = clock64();                                    \
        }                                                                                         \
    } while (0)

template <int KB>
__global__ void __launch_bounds__(kAttn4Threads, 1)
attention4_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv, const Attn2Params p) {
    static_assert(KB == 128 || KB == 96, "KV block of 128 or 96 keys");
    constexpr int NCH = KB / 16;                  // 16-column chunks of an S row: 8 or 6
    constexpr int KSTEPS = KB / 16;               // UMMA_K steps of P V
    constexpr int O_COL = 64;                     // O accumulator columns inside the buffer (P occupies [0, KB/2))
    constexpr int BUF_COLS = 128;
    constexpr uint32_t KV_BYTES = KB * 128;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sKV = smem;
    uint8_t* sQ = sKV + kA4KvSlots * kA4KvSlotBytes;
    float* s_m = reinterpret_cast<float*>(sQ + kA4QSlots * kA2QSlotBytes);            // [4 buf][128] block reference ceil(max * scale * log2(e))
    float* s_l = s_m + kA4Bufs * 128;                                       // [4 buf][128] block sum of 2^(s - m)
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_l + kA4Bufs * 128);
    uint64_t* kv_full = bars;                         // [10]
    uint64_t* kv_empty = kv_full + kA4KvSlots;        // [10]
    uint64_t* q_full = kv_empty + kA4KvSlots;         // [3]
    uint64_t* q_empty = q_full + kA4QSlots;           // [3]
    uint64_t* s_full = q_empty + kA4QSlots;           // [4]  MMA -> softmax (buffer b)
    uint64_t* p_full = s_full + kA4Bufs;              // [4]  softmax (4 warps) -> MMA, drain
    uint64_t* o_full = p_full + kA4Bufs;              // [4]  MMA -> drain
    uint64_t* buf_free = o_full + kA4Bufs;            // [4]  drain (4 warps) -> MMA
    uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(buf_free + kA4Bufs);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform
    const int lane = threadIdx.x & 31;
    const int NT = p.QT * p.NKV;  // S tiles per item
    const int my_items = ((int)blockIdx.x < p.n_items) ? (p.n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int G = my_items * NT;
    // 512-key shape (4 KV blocks x 4 query tiles): two query tiles are interleaved per KV block -- tile order (t0,j0) (t1,j0)
    // (t0,j1) (t1,j1) ... (t0,j3) (t1,j3) (t2,j0) (t3,j0) ... -- so that a query tile only ever uses buffers of one parity
    // (t0: 0,2,0,2; t1: 1,3,1,3) and the two drain groups can share the work there too (ownership by buffer parity).
    const bool inter = (p.NKV == 4 && p.QT == 4);
    auto tile_of = [&](int g, int& il, int& t, int& j) {
        if (inter) {
            il = g >> 4;
            const int n = g & 15;
            j = (n & 7) >> 1;
            t = ((n >> 3) << 1) | (n & 1);
        } else {
            il = g / NT;
            const int n = g % NT;
            t = n / p.NKV;
            j = n % p.NKV;
        }
    };

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_kv);
        for (int i = 0; i < kA4KvSlots; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
        for (int i = 0; i < kA4QSlots; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
        for (int i = 0; i < kA4Bufs; ++i) {
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
      reg_dec<40>();  // one setmaxnreg for the whole TMA / MMA warpgroup (warp 3 only takes part in the barriers)
      if (warp == 0) {
        // ===================================================== TMA producer (warp-uniform loop, one elected lane issues)
        const bool leader = elect_one();
        uint32_t kvc = 0, qc = 0;  // ring counters
        auto load_kv = [&](int col, int row) {
            const uint32_t s = kvc % kA4KvSlots, r = kvc / kA4KvSlots;
            mbar_wait_inl(&kv_empty[s], (r & 1) ^ 1);
            if (leader) {
                mbar_expect_tx(&kv_full[s], KV_BYTES);
                tma_load_2d(sKV + s * kA4KvSlotBytes, &tmap_kv, &kv_full[s], col, row);
            }
            ++kvc;
        };
        auto load_q = [&](int col, int row) {
            const uint32_t s = qc % kA4QSlots, r = qc / kA4QSlots;
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
            if (inter) load_q(qcol, q_row0 + 128);   // the second query tile of the pair is used by the very next S
            load_kv(vcol, kv_row0);
            for (int j = 1; j < p.NKV; ++j) {
                load_kv(kcol, kv_row0 + j * KB);
                load_kv(vcol, kv_row0 + j * KB);
            }
            for (int t = inter ? 2 : 1; t < p.QT; ++t) load_q(qcol, q_row0 + t * 128);
        }
      } else if (warp == 1) {
        // ===================================================== S = Q K^T issuer (warp-uniform loop, one elected lane issues).
        // S and P V are issued by different warps so that neither waits behind the other's barrier: a buffer's next S
        // goes out as soon as the drain warps free it, a tile's P V as soon as its softmax group is done.
        const bool leader = elect_one();
        const uint32_t idesc_s = make_idesc_bf16(128, KB, 0, 0);
        const uint64_t q_desc0 = make_sw128_desc(smem_u32(sQ));
        const uint64_t k_desc0 = make_sw128_desc(smem_u32(sKV));
        for (int g = 0; g < G; ++g) {
            int il, t, j;
            tile_of(g, il, t, j);
            const uint32_t kc = (uint32_t)(il * p.NKV + j) * 2, qc = (uint32_t)(il * p.QT + t);
            const uint32_t ks = kc % kA4KvSlots, qs = qc % kA4QSlots;
            const int b = g & 3;
            mbar_wait_inl(&kv_full[ks], (kc / kA4KvSlots) & 1);
            mbar_wait_inl(&q_full[qs], (qc / kA4QSlots) & 1);
            mbar_wait_inl(&buf_free[b], ((g >> 2) & 1) ^ 1);
            tc_fence_after();
            if (leader) {
                const uint64_t qd = q_desc0 + (uint64_t)(qs * (kA2QSlotBytes >> 4)), kd = k_desc0 + (uint64_t)(ks * (kA4KvSlotBytes >> 4));
                const uint32_t tmem_s = tmem_base + b * BUF_COLS;
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_s, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
                tc_commit(&s_full[b]);
                if (j == p.NKV - 1) tc_commit(&q_empty[qs]);  // last S that reads this Q tile
            }
            A4_TRACE(0, g);
            __syncwarp();
        }
      } else if (warp == 2) {
        // ===================================================== O = P V issuer
        const bool leader = elect_one();
        const uint32_t idesc_o = make_idesc_bf16(128, kHeadDim, 0, 1);  // B = V is MN-major (d contiguous)
        const uint64_t v_desc0 = make_sw128_desc(smem_u32(sKV), 8192);
        for (int g = 0; g < G; ++g) {
            int il, t, j;
            tile_of(g, il, t, j);
            const uint32_t kc = (uint32_t)(il * p.NKV + j) * 2, vc = kc + 1;
            const uint32_t ks = kc % kA4KvSlots, vs = vc % kA4KvSlots;
            const int b = g & 3;
            mbar_wait_inl(&kv_full[vs], (vc / kA4KvSlots) & 1);
            mbar_wait_inl(&p_full[b], (g >> 2) & 1);
            tc_fence_after();
            if (leader) {
                const uint64_t vd = v_desc0 + (uint64_t)(vs * (kA4KvSlotBytes >> 4));
                const uint32_t tmem_buf = tmem_base + b * BUF_COLS;
#pragma unroll
                for (int s = 0; s < KSTEPS; ++s)  // P: bf16 pairs, 8 columns per K = 16; V: 16 keys = 2048 B further
                    umma_bf16_ts(tmem_buf + O_COL, tmem_buf + s * 8, vd + (uint64_t)(s * 128), idesc_o, s != 0);
                tc_commit(&o_full[b]);
                if (t == p.QT - 1) {  // last use of this K / V block: its S tiles completed before their softmax, hence before this P V
                    tc_commit(&kv_empty[ks]);
                    tc_commit(&kv_empty[vs]);
                }
            }
            A4_TRACE(1, g);
            if (p.trace != nullptr) {  // diagnostic only (serialises P V issue): when does this warp see its own commit land?
                mbar_wait_inl(&o_full[b], (g >> 2) & 1);
                A4_TRACE(7, g);
            }
            __syncwarp();
        }
      }
    } else if (warp < 12) {
        // ===================================================== drain warps: O -> registers -> combine KV blocks -> bf16 -> HBM.
        // TWO drain warpgroups: the drain of a tile is a latency chain (two barrier waits, tcgen05.ld round trips, the
        // store) that one warp per sub-partition cannot hide, and the timeline showed it pacing the kernel.  Query tiles
        // alternate between the groups (all KV blocks of a query tile stay with one group: it carries the running O, m, l).
        reg_inc<112>();
        const int dgrp = (warp - 4) >> 2;
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        float acc[64];
        float v[16];
        float m0 = 0.f, l0 = 0.f;
        int il = 0, t = 0, j = 0;   // tile g = (item il, query tile t, KV block j), advanced incrementally (no divisions in the loop)
        int qt = 0;                 // running query-tile counter: this group owns the tiles with (qt & 1) == dgrp
        for (int g = 0; g < G; ++g) {
            // A group may only wait on a buffer whose previous phase it has itself seen complete (mbarrier parity waits alias
            // two phases back).  With NKV <= 2 the alternation gives each group its own buffers ({0,1} / {2,3} or {0,2} / {1,3});
            // with NKV = 4 and t-major order the tiles of one query tile span all four buffers, so group 0 drains everything
            // (the 4 x 4 shape of this model takes the interleaved order above instead).
            if (inter) {
                if ((g & 1) != dgrp) continue;
                tile_of(g, il, t, j);
            } else {
                const bool mine = (p.NKV > 2) ? (dgrp == 0) : ((qt & 1) == dgrp);
                if (!mine) {
                    if (++j == p.NKV) {
                        j = 0; ++qt;
                        if (++t == p.QT) { t = 0; ++il; }
                    }
                    continue;
                }
            }
            const int b = g & 3;
            const uint32_t ph = (g >> 2) & 1;
            if (q == 0) A4_TRACE(2, g);
            mbar_wait_inl(&p_full[b], ph);  // softmax statistics of this tile are visible
            mbar_wait_inl(&o_full[b], ph);
            __syncwarp();
            tc_fence_after();
            if (q == 0) A4_TRACE(4, g);
            const float mj = s_m[b * 128 + row], lj = s_l[b * 128 + row];
            const uint32_t tmem_o = tmem_base + b * BUF_COLS + O_COL + lane_off;
            float inv;
            if (j == 0) {  // straight into the accumulator registers, all four loads in flight
#pragma unroll
                for (int c = 0; c < 4; ++c) tmem_ld16(tmem_o + c * 16, acc + c * 16);
                tc_wait_ld();
                m0 = mj; l0 = lj;
                inv = rcp_fma(lj);
            } else {  // further KV block of this query tile: exact combination of independently normalised blocks
                const float mm = fmaxf(m0, mj);
                const float a0 = pow2_int(m0 - mm), a1 = pow2_int(mj - mm);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    tmem_ld16(tmem_o + c * 16, v);
                    tc_wait_ld();
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc[c * 16 + i] = fmaf(acc[c * 16 + i], a0, v[i] * a1);
                }
                m0 = mm;
                l0 = fmaf(l0, a0, lj * a1);
                inv = rcp_fma(l0);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&buf_free[b]);
            if (q == 0) A4_TRACE(5, g);
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
            if (q == 0) A4_TRACE(6, g);
            if (!inter && ++j == p.NKV) {
                j = 0; ++qt;
                if (++t == p.QT) { t = 0; ++il; }
            }
        }
    } else {
        // ===================================================== softmax warps: three groups, tile g -> group g % 3 (buffer g & 3), one
        // thread per query row.  Three warps per SM sub-partition in different phases keep the MUFU pipe fed through each
        // other's maximum pass, tcgen05.ld latencies and waits for S; 16-column chunks keep a thread under 72 registers.
        reg_dec<72>();
        const int grp = (warp - 12) >> 2;   // 0..2
        const int q = warp & 3;   // TMEM lane quarter of this warp
        const int row = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const float scale = p.scale_log2e;
        float v[16];
        int j = grp % p.NKV;   // KV block of tile g (NT is a multiple of NKV, so j = g mod NKV), advanced incrementally
        const int j_step = kA4SoftmaxGroups % p.NKV;
        for (int g = grp; g < G; g += kA4SoftmaxGroups, j = (j + j_step >= p.NKV) ? j + j_step - p.NKV : j + j_step) {
            const int b = g & 3;
            const uint32_t tmem_s = tmem_base + b * BUF_COLS + lane_off;
            mbar_wait_inl(&s_full[b], (g >> 2) & 1);
            __syncwarp();
            tc_fence_after();
            const int keys_here = min(KB, p.Lk - j * KB);
            // ---- pass 1: row maximum
            float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                tmem_ld16(tmem_s + c * 16, v);
                tc_wait_ld();
                if (KB == 128 || (c + 1) * 16 <= keys_here) {
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        m0 = fmax3(m0, v[i], v[i + 1]);
                        m1 = fmax3(m1, v[i + 2], v[i + 3]);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (c * 16 + i < keys_here) m0 = fmaxf(m0, v[i]);
                }
            }
            const float m_sc = ceilf(fmaxf(m0, m1) * scale);  // integer reference >= the row maximum (log2 domain): see pow2_int
            // ---- pass 2: p = 2^(s * scale - m * scale) -> bf16 P over the S columns already consumed; row sum
            float l0 = 0.f, l1 = 0.f;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                tmem_ld16(tmem_s + c * 16, v);
                tc_wait_ld();
                uint32_t pk[8];
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    float e0 = ex2_approx(fmaf(v[i], scale, -m_sc));
                    float e1 = ex2_approx(fmaf(v[i + 1], scale, -m_sc));
                    if (KB != 128) {
                        if (c * 16 + i >= keys_here) e0 = 0.f;
                        if (c * 16 + i + 1 >= keys_here) e1 = 0.f;
                    }
                    l0 += e0; l1 += e1;
                    pk[i >> 1] = pack_bf16x2(e0, e1);
                }
                // P chunk c (16 keys) -> columns [8 c, 8 c + 8): below the S columns [16 (c + 1), KB) still to be read
                tmem_st8(tmem_s + c * 8, pk);
            }
            s_m[b * 128 + row] = m_sc;
            s_l[b * 128 + row] = l0 + l1;
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
