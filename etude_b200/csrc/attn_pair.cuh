// Self-attention of a frequency-axis EncoderLayer with the Q|K|V projection fused in, on CTA-PAIR MMAs (cta_group::2).
// Same math and interface as attn_qkv.cuh (reference amt_apc.py:342-368; sequences of 256 tokens, 4 heads x 64): one cluster
// of two CTAs per sequence, CTA r owns tokens [128 r, 128 r + 128).
//
// Why a second structure: attn_qkv.cuh is latency-bound (ncu, profiles/r2y_ncu_attn_qkv_details.txt: tensor pipe 36 %, MUFU
// 29 %, issue 30 %; timeline profiles/r2u_*): every head's K and V halves have to cross the cluster (a 16 KB DSMEM bulk
// copy each, ~3 000 clk) before S / P V can be issued, the two key blocks share one softmax reference and therefore move
// in lock step, and smem is full, so nothing can be double-buffered.  With pair MMAs the operands never move:
//   * projection  ACC[256 x 192] = x[256 x 256] W_h^T : A = each CTA's resident x half, B = W_h split along N -- each CTA
//     loads only ITS 96 rows of a weight box (12 KB stages instead of 24 KB, no multicast);
//   * S_j[256 x 128] = Q K_j^T for the two key blocks : B = K rows split along N -- CTA r supplies the K rows of ITS tokens
//     [64 j, 64 j + 64) straight from where its projection epilogue wrote them.  No K exchange;
//   * O[256 x 64] = P_0 V_0 + P_1 V_1 : A = P in each CTA's TMEM, B = V^T split along N -- CTA r supplies output dims
//     [32 r, 32 r + 32) for all 256 keys.  The projection epilogue writes V transposed (K-major V^T tiles); the half of its
//     tokens' V that belongs to the peer's dims (8 KB) goes through a staging buffer and a DSMEM bulk copy -- a quarter of
//     the old exchange, and off the critical path: V^T and the staging are double-buffered and V is needed only after the
//     softmax.
// The key blocks are DECOUPLED: block 0 takes its own integer reference m_0 = ceil(max_0); block 1 uses r_1 = max(m_0,
// ceil(max_1) - 100).  If r_1 = m_0 both blocks share one reference (exact; P_1 <= 2^100 fits bf16 and the fp32 sums); if
// not, block 0's true weight relative to block 1 is below 2^-100 and it is computed as exactly 2^-100: both vanish at fp32
// precision, so one O accumulator and l = l_0 + l_1 stay correct without any rescaling.  Block 1 only needs block 0's
// maximum (ready long before its own S), so the two softmax groups run out of phase and keep the MUFU pipe busy while the
// other block is in its P V -> next S window.
//
// Only the leader CTA (rank 0) issues MMAs; its barriers collect both CTAs' producers (counts doubled, remote mbarrier
// arrives / TMA complete_tx into the leader's barrier), and every completion is multicast to both CTAs.
//
// TMEM (512 columns, allocated pair-wise): [0, 192) projection accumulator Q|K|V, [192, 256) O, [256, 384) / [384, 512) S_j / P_j.
// smem: x 64 KB | W ring 4 x 12 KB | Q 16 KB | K 16 KB | V^T 2 x 16 KB | V staging for the peer 2 x 8 KB | statistics 4 KB |
//       bias 3 KB | barriers.
// Warps (24), in the order of the warp scheduler's preference (it favours HIGHER warp ids): 20 TMA producer, 21 projection
// issue (leader) + TMEM alloc, 22 S issue (leader), 23 P V issue (leader; in the peer: forwards "my V^T is complete" to the
// leader) -- a handful of instructions per head, but every clock they wait for an issue slot is a clock of the head's critical
// cycle; 12-19 projection epilogue and 8-11 drain (short, latency-critical: they publish K / V^T and free O); 0-3 / 4-7
// softmax of block 0 / 1 (long, MUFU-bound: they fill whatever the others leave).  With the epilogue below the softmax warps a
// head's V^T was published 3 900 clk late; with the issue warps below them every hand-off to an MMA took ~1 000 clk.
#pragma once
#include "attn_qkv.cuh"
#include "pairmma.cuh"

namespace etude {

// Debug timeline of cluster 0, BOTH ranks (dev build sets p.trace): role r < 8 of rank k goes to role slot 8 k + r; clocks are
// relative to each thread's clock64 right after the start-up cluster barrier, which puts the two SMs on one axis to within
// the barrier's release skew (a few hundred clocks).
#define AP_TRACE(role, n, e)                                                                                    \
    do {                                                                                                        \
        if (p.trace != nullptr && cid == 0 && lane == 0 && (n) < 64)                                            \
            p.trace[((((int)rank * 8 + (role)) * 64 + (n)) << 3) + (e)] = clock64() - t_origin + 1;              \
    } while (0)

// Wait that does not poll: try_wait with a suspend-time hint parks the thread until the phase completes (or the hint
// expires).  The warp scheduler prefers higher warp ids, so the short roles placed above the softmax warps must not spin on
// their barriers for the thousands of clocks they wait per head -- a polling warp takes issue slots from the MUFU-bound ones.
__device__ __forceinline__ void mbar_wait_park(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    auto try_once = [&]() -> bool {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity), "r"(20000u)
            : "memory");
        return ok != 0;
    };
    if (try_once()) return;
    const unsigned long long t0 = globaltimer_ns();
#pragma unroll 1
    for (;;) {
#pragma unroll 1
        for (uint32_t i = 0; i < 1024u; ++i)
            if (try_once()) return;
        if (globaltimer_ns() - t0 > kWaitBudgetNs) break;
    }
    __trap();
}

// 2^x for a pair of x <= 0 on the FMA and integer pipes instead of the MUFU (16 ex2 per clock and SM: with both key blocks
// in their exponential pass the MUFU pipe is what the two softmax warps of a sub-partition queue on, while two thirds of its
// issue slots are idle).  Round-to-nearest split x = j + f with the 1.5 * 2^23 trick (the low mantissa bits of
// t = x + 1.5 * 2^23 hold j in two's complement), degree-3 minimax polynomial of 2^f on [-0.5, 0.5] (relative error 7.6e-5,
// far below the bf16 rounding of P), then j goes into the exponent field with one shift-add.  x is clamped at -125 so that the
// exponent cannot wrap.  AP_POLY_PAIRS of the 8 pairs of every 16-column chunk take this path.
__device__ __forceinline__ float2 ap_exp2_poly2(float2 x) {
    const float kMagic = 12582912.f;   // 1.5 * 2^23
    x.x = fmaxf(x.x, -125.f);
    x.y = fmaxf(x.y, -125.f);
    const float2 t = f2add(x, make_float2(kMagic, kMagic));
    const float2 fl = f2add(t, make_float2(-kMagic, -kMagic));
    const float2 f = f2fma(fl, make_float2(-1.f, -1.f), x);
    float2 pl = f2fma(f, make_float2(0.05520550534129143f, 0.05520550534129143f), make_float2(0.24261397123336792f, 0.24261397123336792f));
    pl = f2fma(pl, f, make_float2(0.6932547688484192f, 0.6932547688484192f));
    pl = f2fma(pl, f, make_float2(0.9999276995658875f, 0.9999276995658875f));
    float2 r;
    r.x = __int_as_float((__float_as_int(t.x) << 23) + __float_as_int(pl.x));
    r.y = __int_as_float((__float_as_int(t.y) << 23) + __float_as_int(pl.y));
    return r;
}
#ifndef AP_POLY_PAIRS
#define AP_POLY_PAIRS 3   // measured in the 32-song step (profiles/r3f_ab_notes.txt): 0 -> 323 ms, 2 -> 329, 3 -> 304, 4 -> 317, 5 -> 333
#endif
#ifndef AP_S_N256
#define AP_S_N256 0   // 1: both key blocks from one N = 256 MMA per K step (saves 188 clk of tensor time per head, but ties the blocks together): measured 306 -> 331 ms per 32-song step, not used
#endif
#ifndef AP_SKEW_NS
#define AP_SKEW_NS 0   // measured in the step: 0 -> 323 ms, 900 -> 328 ms per 32 songs (the dependencies through Q / K and O pull the blocks back together)
#endif
constexpr int kApWStages = 4;
constexpr int kApWStageBytes = 96 * 64 * 2;           // this CTA's 96 rows of a [192 x 64] weight box
constexpr int kApVtBytes = 4 * 32 * 64 * 2;           // V^T of this CTA's 32 dims: 4 chunks of 64 keys x [32 x 64] bf16
constexpr int kApVxBytes = 2 * 32 * 64 * 2;           // staging: the peer's 32 dims of this CTA's 128 tokens (2 chunks)
constexpr size_t kAttnPairSmemBytes = kAqXBytes + kApWStages * kApWStageBytes + 2 * kAqQBytes + 2 * kApVtBytes + 2 * kApVxBytes +
                                      kAqStatBytes + 768 * 4 + 512;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kAqThreads, 1)
attn_pair_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const AttnQkvParams p) {
    constexpr int O_COL = 192, BUF0_COL = 256, BUF_COLS = 128;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
    uint8_t* sX = smem;
    uint8_t* sW = sX + kAqXBytes;
    uint8_t* sQ = sW + kApWStages * kApWStageBytes;
    uint8_t* sK = sQ + kAqQBytes;
    uint8_t* sVT = sK + kAqQBytes;              // [2 buffers][4 chunks][32 x 64]
    uint8_t* sVX = sVT + 2 * kApVtBytes;        // [2 buffers][2 chunks][32 x 64]
    float* s_mx = reinterpret_cast<float*>(sVX + 2 * kApVxBytes);  // [n & 1][128] block 0's integer reference (log2 domain)
    float* s_l = s_mx + 2 * 2 * 128;                                // [n & 1][2 blocks][128] block sums
    float* s_bias = s_l + 2 * 2 * 128;                              // [4][192]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + 768);
    uint64_t* w_full = bars;                    // [4] leader: 1 arrive + 24 KB (both CTAs' halves)
    uint64_t* w_empty = w_full + kApWStages;    // [4] each CTA: projection MMAs on the slot complete (multicast commit)
    uint64_t* x_full = w_empty + kApWStages;    // leader: 1 arrive + 128 KB
    uint64_t* x_free = x_full + 1;              // each CTA (multicast commit)
    uint64_t* acc_full = x_free + 1;            // each CTA (multicast commit)
    uint64_t* acc_free = acc_full + 1;          // leader: 16 epilogue warps (8 local + 8 remote)
    uint64_t* qk_ready = acc_free + 1;          // leader: 2 (one thread per CTA once Q and K are written)
    uint64_t* qk_free = qk_ready + 1;           // each CTA (multicast commit after the S MMAs)
    uint64_t* vt_full = qk_free + 1;            // [2] each CTA: 1 local arrive + 8 KB from the peer's bulk copies
    uint64_t* v_peer = vt_full + 2;             // [2] leader: the peer's vt_full has completed (remote arrive)
    uint64_t* v_free = v_peer + 2;              // [2] each CTA (multicast commit after the P V MMAs)
    uint64_t* s_full = v_free + 2;              // [2] each CTA (multicast commit)
    uint64_t* p_full = s_full + 2;              // [2 blocks][2 (n & 1)] each CTA: 4 softmax warps (for the drain: row sums visible)
    uint64_t* p_ready = p_full + 4;             // [2 blocks][2 (n & 1)] leader: 8 softmax warps (4 local + 4 remote)
    uint64_t* m0_ready = p_ready + 4;           // [4 lane quarters][2 (n & 1)] each CTA: block 0's reference published
    uint64_t* buf_free = m0_ready + 8;          // [2] leader: P_j V_j complete (S / P buffer j reusable in both CTAs)
    uint64_t* o_full = buf_free + 2;            // each CTA (multicast commit)
    uint64_t* o_free = o_full + 1;              // leader: 8 drain warps
    uint64_t* s_issued = o_free + 1;            // [2 (n & 1)] leader: the S MMAs of head n are in the tensor queue (S issue -> projection issue)
    uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(s_issued + 2);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank(), peer = rank ^ 1u;
    const bool lead_cta = rank == 0;
    const int cid = (int)cluster_id_x(), ncl = (int)cluster_nctaid_x();
    const int my_items = (cid < p.n_seq) ? (p.n_seq - 1 - cid) / ncl + 1 : 0;
    const int N = my_items * 4;   // (sequence, head) iterations
    constexpr uint16_t kBoth = 3;
    // arrive on a barrier of the leader CTA from either CTA
    auto arrive_leader = [&](uint64_t* bar) {
        if (lead_cta) mbar_arrive(bar);
        else mbar_arrive_cluster(mapa_u32(smem_u32(bar), 0));
    };
    // the same for hand-offs of TMEM regions only (no ordinary memory to publish: see mbar_arrive_cluster_relaxed)
    auto arrive_leader_tmem = [&](uint64_t* bar) {
        if (lead_cta) mbar_arrive(bar);
        else mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(bar), 0));
    };

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_w);
        for (int s = 0; s < kApWStages; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        mbar_init(x_full, 1); mbar_init(x_free, 1);
        mbar_init(acc_full, 1); mbar_init(acc_free, 16);
        mbar_init(qk_ready, 2); mbar_init(qk_free, 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(&vt_full[b], 1); mbar_init(&v_peer[b], 1); mbar_init(&v_free[b], 1);
            mbar_init(&s_full[b], 1); mbar_init(&buf_free[b], 1);
        }
        for (int b = 0; b < 4; ++b) { mbar_init(&p_full[b], 4); mbar_init(&p_ready[b], 8); }
        for (int b = 0; b < 8; ++b) mbar_init(&m0_ready[b], 1);
        mbar_init(o_full, 1); mbar_init(o_free, 8);
        mbar_init(&s_issued[0], 1); mbar_init(&s_issued[1], 1);
        mbar_fence_init();
    }
    if (warp == 21) tmem_alloc2(tmem_base_ptr, 512);
    for (int i = threadIdx.x; i < 768; i += kAqThreads) s_bias[i] = p.bias[i];
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // the peer's barriers exist before anything is signalled in this CTA
    tc_fence_after();
    const long long t_origin = clock64();
    const uint32_t tmem_base = *tmem_base_ptr;

    if (warp >= 20) {
      reg_dec<40>();
      if (warp == 20) {
        // ===================================================== TMA producer (both CTAs): own x half, own 96 rows of every W box;
        // the bytes are counted on the LEADER's barriers
        const bool leader = elect_one();
        const uint32_t x_full_l = mapa_u32(smem_u32(x_full), 0);
        uint32_t c = 0;   // ring counter
        for (int il = 0; il < my_items; ++il) {
            const int seq = cid + il * ncl;
            const int row0 = seq * 256 + (int)rank * 128;
            mbar_wait_cl(x_free, (il & 1) ^ 1);
            AP_TRACE(0, il * 4, 0);
            if (leader) {
                if (lead_cta) mbar_expect_tx(x_full, 2 * kAqXBytes);
#pragma unroll
                for (int kc = 0; kc < 4; ++kc) tma_load_2d_pair(sX + kc * 16384, &tmap_x, x_full_l, kc * 64, row0);
            }
            for (int hk = 0; hk < 16; ++hk, ++c) {   // (head, K-chunk) boxes in consumption order
                const uint32_t s = c % kApWStages;
                mbar_wait_cl(&w_empty[s], ((c / kApWStages) & 1) ^ 1);
                AP_TRACE(0, il * 4 + (hk >> 2), 1 + (hk & 3));
                if (leader) {
                    if (lead_cta) mbar_expect_tx(&w_full[s], 2 * kApWStageBytes);
                    tma_load_2d_pair(sW + s * kApWStageBytes, &tmap_w, mapa_u32(smem_u32(&w_full[s]), 0), (hk & 3) * 64,
                                     (hk >> 2) * 192 + (int)rank * 96);
                }
            }
            __syncwarp();
        }
      } else if (warp == 21) {
        // ===================================================== projection issue (leader CTA): ACC[256 x 192] = x W_h^T
        if (lead_cta) {
            const bool leader = elect_one();
            constexpr uint32_t idesc_p = make_idesc_bf16(256, 192, 0, 0);
            const uint64_t x_desc0 = make_sw128_desc(smem_u32(sX));
            const uint64_t w_desc0 = make_sw128_desc(smem_u32(sW));
            uint32_t c = 0;
            for (int n = 0; n < N; ++n) {
                const int h = n & 3;
                if (h == 0) mbar_wait_cl(x_full, (n >> 2) & 1);
                mbar_wait_cl(acc_free, (n & 1) ^ 1);
                // The tensor pipe runs its queue in order: a head's 16 projection MMAs (1 500 clk) issued just before an S or
                // P V batch delay that batch -- and S -> softmax -> P V -> next S is the cycle that sets the head period.  The
                // projection of head n therefore enters the queue right AFTER the S MMAs of head n - 2 (it then runs under
                // that head's softmax); the accumulator is free long before.
                mbar_wait_inl(&s_issued[n & 1], ((n >> 1) & 1) ^ 1);
                tc_fence_after();
                AP_TRACE(1, n, 0);
#pragma unroll 1
                for (int kc = 0; kc < 4; ++kc, ++c) {
                    const uint32_t s = c % kApWStages;
                    mbar_wait_cl(&w_full[s], (c / kApWStages) & 1);
                    tc_fence_after();
                    AP_TRACE(1, n, 1 + kc);
                    if (leader) {
                        const uint64_t ad = x_desc0 + (uint64_t)(kc * (16384 >> 4)), bd = w_desc0 + (uint64_t)(s * (kApWStageBytes >> 4));
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma2_bf16_ss(tmem_base, ad + 2 * k, bd + 2 * k, idesc_p, (kc | k) ? 1u : 0u);
                        tc_commit2_mc(&w_empty[s], kBoth);
                    }
                    __syncwarp();
                }
                if (leader) {
                    tc_commit2_mc(acc_full, kBoth);
                    if (h == 3) tc_commit2_mc(x_free, kBoth);
                }
                __syncwarp();
            }
        }
      } else if (warp == 22) {
        // ===================================================== S_j = Q K_j^T issue (leader CTA): block j = keys [64 j, 64 j + 64) of each CTA
        if (lead_cta) {
            const bool leader = elect_one();
            constexpr uint32_t idesc_s = make_idesc_bf16(256, AP_S_N256 ? 256 : 128, 0, 0);
            const uint64_t q_desc = make_sw128_desc(smem_u32(sQ));
            const uint64_t k_desc0 = make_sw128_desc(smem_u32(sK));
            for (int n = 0; n < N; ++n) {
                mbar_wait_cl(qk_ready, n & 1);
                AP_TRACE(2, n, 0);
#if AP_S_N256
                // One N = 256 MMA per K step for both key blocks: an SS MMA costs 43 + N / 2 clk, so 4 x 171 clk
                // instead of 8 x 107.  N is split over the CTAs as rows [0, 128) of each sK, so column c < 128 of S is key c of
                // CTA 0 and column 128 + c key c of CTA 1: block j = the keys of CTA j (the V^T chunk order follows, see PV).
                mbar_wait_inl(&buf_free[0], (n & 1) ^ 1);
                mbar_wait_inl(&buf_free[1], (n & 1) ^ 1);
                tc_fence_after();
                AP_TRACE(2, n, 1);
                if (leader) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma2_bf16_ss(tmem_base + BUF0_COL, q_desc + 2 * k, k_desc0 + 2 * k, idesc_s, k != 0);
                    tc_commit2_mc(&s_full[0], kBoth);
                    tc_commit2_mc(&s_full[1], kBoth);
                    tc_commit2_mc(qk_free, kBoth);
                    mbar_arrive(&s_issued[n & 1]);
                }
                __syncwarp();
#else
#pragma unroll 1
                for (int j = 0; j < 2; ++j) {
                    mbar_wait_inl(&buf_free[j], (n & 1) ^ 1);
                    tc_fence_after();
                    AP_TRACE(2, n, 1 + j);
                    if (leader) {
                        const uint64_t kd = k_desc0 + (uint64_t)(j * (8192 >> 4));
                        const uint32_t tmem_s = tmem_base + BUF0_COL + j * BUF_COLS;
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma2_bf16_ss(tmem_s, q_desc + 2 * k, kd + 2 * k, idesc_s, k != 0);
                        tc_commit2_mc(&s_full[j], kBoth);
                        if (j == 1) {
                            tc_commit2_mc(qk_free, kBoth);
                            mbar_arrive(&s_issued[n & 1]);
                        }
                    }
                    __syncwarp();
                }
#endif
            }
        }
      } else {
        // ===================================================== O = P_0 V_0 + P_1 V_1 issue (leader CTA); in the peer this warp tells the
        // leader when the peer's V^T buffer is complete (its own writes + the leader's 8 KB)
        if (lead_cta) {
            const bool leader = elect_one();
            constexpr uint32_t idesc_o = make_idesc_bf16(256, kHeadDim, 0, 0);
            const uint64_t vt_desc0 = make_sw128_desc(smem_u32(sVT));
            for (int n = 0; n < N; ++n) {
                const int buf = n & 1, k2 = (n >> 1) & 1;
                mbar_wait_cl(&vt_full[buf], k2);
                mbar_wait_cl(&v_peer[buf], k2);
                mbar_wait_cl(o_free, (n & 1) ^ 1);
                AP_TRACE(3, n, 0);
#pragma unroll 1
                for (int j = 0; j < 2; ++j) {
                    mbar_wait_cl(&p_ready[j * 2 + (n & 1)], k2);
                    tc_fence_after();
                    AP_TRACE(3, n, 1 + j);
                    if (leader) {
                        const uint32_t tmem_p = tmem_base + BUF0_COL + j * BUF_COLS;
                        const uint64_t vd = vt_desc0 + (uint64_t)((buf * kApVtBytes + j * 8192) >> 4);
#pragma unroll
                        for (int s = 0; s < 8; ++s)   // 16 keys per step: chunk 2 j + (s >> 2) (4 KB each), 32 B per step inside
                            umma2_bf16_ts(tmem_base + O_COL, tmem_p + s * 8, vd + (uint64_t)((s >> 2) * (4096 >> 4) + (s & 3) * 2), idesc_o,
                                          (j | s) ? 1u : 0u);
                        tc_commit2(&buf_free[j]);
                        if (j == 1) {
                            tc_commit2_mc(o_full, kBoth);
                            tc_commit2_mc(&v_free[buf], kBoth);
                        }
                    }
                    __syncwarp();
                }
            }
        } else {
            const uint32_t v_peer_l0 = mapa_u32(smem_u32(&v_peer[0]), 0), v_peer_l1 = mapa_u32(smem_u32(&v_peer[1]), 0);
            for (int n = 0; n < N; ++n) {
                mbar_wait_cl(&vt_full[n & 1], (n >> 1) & 1);
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster((n & 1) ? v_peer_l1 : v_peer_l0);
            }
        }
      }
    } else if (warp >= 12) {
        // ===================================================== projection epilogue (8 warps, both CTAs): ACC + bias -> bf16 Q, K rows
        // (K-major, where the S MMAs read them) and V TRANSPOSED: dims [32 rank, + 32) into this CTA's V^T buffer, the other 32
        // dims into the staging buffer that one thread bulk-copies into the peer's V^T buffer
        const int q = warp & 3, half = (warp - 12) >> 2;   // TMEM lane quarter, 32-column half of each of Q / K / V
        const int row = q * 32 + lane;                     // token of this CTA's half = TMEM lane
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int sw = row & 7;
        const uint32_t q_row = smem_u32(sQ) + row * 128, k_row = smem_u32(sK) + row * 128;
        const bool own_dims = (uint32_t)half == rank;      // this warp's V dims stay in this CTA
        const int jj = row >> 6, kk = row & 63;            // key block and key inside the block
        // byte offset of (dim row 0, key kk) inside a [32 x 64] V^T chunk, minus the swizzle term that depends on the dim row
        const uint32_t vt_col = (uint32_t)((kk & 7) << 1), vt_c16 = (uint32_t)(kk >> 3);
        const bool copier = (warp == 12) && elect_one();
        float v[32];
        auto load_pack = [&](int c, const float* bias, uint4 (&pk)[4]) {   // ACC columns [32 c, 32 c + 32) + bias -> 32 bf16
            tmem_ld32(tmem_base + lane_off + c * 32, v);
            tc_wait_ld();
            const float4* b4 = reinterpret_cast<const float4*>(bias + c * 32);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const float4 ba = b4[2 * g], bb = b4[2 * g + 1];
                pk[g].x = pack_bf16x2(v[8 * g + 0] + ba.x, v[8 * g + 1] + ba.y); pk[g].y = pack_bf16x2(v[8 * g + 2] + ba.z, v[8 * g + 3] + ba.w);
                pk[g].z = pack_bf16x2(v[8 * g + 4] + bb.x, v[8 * g + 5] + bb.y); pk[g].w = pack_bf16x2(v[8 * g + 6] + bb.z, v[8 * g + 7] + bb.w);
            }
        };
        auto st_row = [&](uint32_t row_addr, const uint4 (&pk)[4]) {   // this warp's half of a 128-byte row: chunks [4 half, 4 half + 4)
#pragma unroll
            for (int g = 0; g < 4; ++g)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_addr + (uint32_t)(((half * 4 + g) ^ sw) << 4)), "r"(pk[g].x),
                             "r"(pk[g].y), "r"(pk[g].z), "r"(pk[g].w) : "memory");
        };
        for (int n = 0; n < N; ++n) {
            const float* bias = s_bias + (n & 3) * 192;
            const int buf = n & 1;
            uint4 pq[4], pk[4];
            mbar_wait_park(acc_full, n & 1);
            __syncwarp();
            tc_fence_after();
            if (warp == 12) AP_TRACE(4, n, 0);
            // The accumulator goes back to the projection issue warp as soon as its three 32-column slices are in registers
            // (the next head's projection needs BOTH CTAs' epilogues to have let go of it: the round trip through the peer is
            // on the projection loop's cycle), before any smem store or wait of this head
            load_pack(half, bias, pq);          // Q columns [32 half, + 32)
            load_pack(2 + half, bias, pk);      // K
            tmem_ld32(tmem_base + lane_off + (4 + half) * 32, v);   // V: 32 dims of this thread's token
            tc_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_leader_tmem(acc_free);
            if (warp == 12) AP_TRACE(4, n, 1);
            mbar_wait_park(qk_free, (n & 1) ^ 1);   // the S MMAs of the previous head have read Q / K in both CTAs
            st_row(q_row, pq);
            st_row(k_row, pk);
            fence_async_smem();                 // generic-proxy writes -> visible to the pair MMAs (async proxy)
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (copier) arrive_leader(qk_ready);
            if (warp == 12) AP_TRACE(4, n, 2);
            // V as bf16, scattered into a K-major V^T tile (one 2-byte store per dim; the 32 lanes of a warp write 64
            // contiguous bytes of one row)
            mbar_wait_park(&v_free[buf], ((n >> 1) & 1) ^ 1);   // the P V MMAs of head n - 2 have read this V^T buffer (and the
                                                              // copies out of this staging buffer have landed)
            if (warp == 12) AP_TRACE(4, n, 3);
            {
                const uint32_t base = own_dims ? smem_u32(sVT) + buf * kApVtBytes + (AP_S_N256 ? 2 * (int)rank + jj : 2 * jj + (int)rank) * 4096
                                               : smem_u32(sVX) + buf * kApVxBytes + jj * 4096;
                const float4* bv4 = reinterpret_cast<const float4*>(bias + (4 + half) * 32);
                uint32_t hv[16];   // bf16 pairs (dims 2 i, 2 i + 1): all values first, then 32 independent stores
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const float4 b = bv4[g];
                    hv[2 * g] = pack_bf16x2(v[4 * g] + b.x, v[4 * g + 1] + b.y);
                    hv[2 * g + 1] = pack_bf16x2(v[4 * g + 2] + b.z, v[4 * g + 3] + b.w);
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const uint32_t addr = base + (uint32_t)(i * 128) + (((vt_c16 ^ (uint32_t)(i & 7)) << 4) | vt_col);
                    const uint16_t h16 = (uint16_t)((i & 1) ? (hv[i >> 1] >> 16) : (hv[i >> 1] & 0xFFFFu));
                    asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(h16));
                }
            }
            fence_async_smem();
            asm volatile("bar.sync 2, 256;" ::: "memory");
            if (copier) {
                mbar_expect_tx(&vt_full[buf], 2 * 4096);   // arrive (own writes done) + the peer's two chunks on their way
                const uint32_t bar_peer = mapa_u32(smem_u32(&vt_full[buf]), peer);
#pragma unroll
                for (int c = 0; c < 2; ++c)   // staging chunk c = my tokens [64 c, 64 c + 64) -> the peer's chunk 2 c + rank
                    dsmem_bulk_copy(mapa_u32(smem_u32(sVT) + buf * kApVtBytes + (AP_S_N256 ? 2 * (int)rank + c : 2 * c + (int)rank) * 4096, peer),
                                    smem_u32(sVX) + buf * kApVxBytes + c * 4096, 4096, bar_peer);
            }
            if (warp == 12) AP_TRACE(4, n, 4);
        }
    } else if (warp >= 8) {
        // ===================================================== drain: O / (l_0 + l_1) -> bf16 context rows -> HBM
        reg_inc<96>();
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        float acc[64];
        for (int n = 0; n < N; ++n) {
            const uint32_t ph = n & 1;
            mbar_wait_park(&p_full[0 + ph], (n >> 1) & 1);   // the row sums of both blocks are visible
            mbar_wait_park(&p_full[2 + ph], (n >> 1) & 1);
            mbar_wait_park(o_full, ph);
            __syncwarp();
            tc_fence_after();
            if (q == 0) AP_TRACE(5, n, 0);
            const uint32_t tmem_o = tmem_base + O_COL + lane_off;
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_ld16(tmem_o + c * 16, acc + c * 16);   // all four loads in flight
            const float inv = rcp_fma(s_l[(ph * 2 + 0) * 128 + row] + s_l[(ph * 2 + 1) * 128 + row]);   // before o_free: the slot is rewritten at n + 2
            tc_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_leader_tmem(o_free);
            if (q == 0) AP_TRACE(5, n, 1);
            const int seq = cid + (n >> 2) * ncl, head = n & 3;
            __nv_bfloat16* dst = p.out + (size_t)(seq * 256 + (int)rank * 128 + row) * kHid + head * kHeadDim;
#pragma unroll
            for (int gq = 0; gq < 8; ++gq) {
                uint4 pk;
                pk.x = pack_bf16x2(acc[gq * 8 + 0] * inv, acc[gq * 8 + 1] * inv);
                pk.y = pack_bf16x2(acc[gq * 8 + 2] * inv, acc[gq * 8 + 3] * inv);
                pk.z = pack_bf16x2(acc[gq * 8 + 4] * inv, acc[gq * 8 + 5] * inv);
                pk.w = pack_bf16x2(acc[gq * 8 + 6] * inv, acc[gq * 8 + 7] * inv);
                *reinterpret_cast<uint4*>(dst + gq * 8) = pk;
            }
            if (q == 0) AP_TRACE(5, n, 2);
        }
    } else {
        // ===================================================== softmax of key block j (warps 0-3: j = 0, 4-7: j = 1): one thread per
        // (query row, block).  Block 0 publishes its integer reference; block 1 adopts it unless its own maximum is more than
        // 2^100 above (see the header) -- the two groups never wait for each other otherwise.
        reg_inc<88>();
        const int j = warp >> 2;
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const float scale = p.scale_log2e;
        const uint32_t tmem_s = tmem_base + BUF0_COL + j * BUF_COLS + lane_off;
        float v[16], vb[16];
        for (int n = 0; n < N; ++n) {
            const int par = n & 1;
            mbar_wait_cl(&s_full[j], par);
#if AP_SKEW_NS
            // Start-up skew: the key blocks are independent pipelines (S_j -> softmax_j -> P_j V_j -> next S_j).  Started together
            // they stay in phase: both softmax groups fight for the MUFU pipe, then both idle through P V and the next S.  Held
            // back once by half a period, block 1 keeps that lag (nothing re-aligns the two cycles), and one group's MUFU
            // phase runs under the other's MMA phase.
            if (j == 1 && n == 0) __nanosleep(AP_SKEW_NS);
#endif
            __syncwarp();
            tc_fence_after();
            if (q == 0) AP_TRACE(6 + j, n, 0);
            // ---- pass 1: row maximum of this block (TMEM loads software-pipelined over two 16-column register buffers)
            float m0 = -INFINITY, m1 = -INFINITY;
            auto max_chunk = [&](const float* w) {
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    m0 = fmax3(m0, w[i], w[i + 1]);
                    m1 = fmax3(m1, w[i + 2], w[i + 3]);
                }
            };
            tmem_ld16(tmem_s, v);
#pragma unroll
            for (int c = 0; c < 8; c += 2) {
                tc_wait_ld();
                tmem_ld16(tmem_s + (c + 1) * 16, vb);
                max_chunk(v);
                tc_wait_ld();
                tmem_ld16(tmem_s + ((c + 2) & 7) * 16, v);   // after the last chunk: chunk 0 again, for pass 2
                max_chunk(vb);
            }
            float m_sc = ceilf(fmaxf(m0, m1) * scale);   // integer reference >= the block's row maximum (log2 domain)
            if (j == 0) {
                s_mx[par * 128 + row] = m_sc;
                __syncwarp();
                if (lane == 0) mbar_arrive(&m0_ready[q * 2 + par]);
            } else {
                mbar_wait_inl(&m0_ready[q * 2 + par], (n >> 1) & 1);
                m_sc = fmaxf(s_mx[par * 128 + row], m_sc - 100.f);
            }
            if (q == 0) AP_TRACE(6 + j, n, 1);
            // ---- pass 2: p = 2^(s * scale - m) -> bf16 P over the S columns already consumed; block row sum
            float2 l2 = make_float2(0.f, 0.f);
            const float2 sc2 = make_float2(scale, scale), nm2 = make_float2(-m_sc, -m_sc);
            auto exp_chunk = [&](const float* w, int c) {
                uint32_t pk[8];
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const float2 x = f2fma(make_float2(w[i], w[i + 1]), sc2, nm2);
                    const float2 e = (i >> 1) < AP_POLY_PAIRS ? ap_exp2_poly2(x) : make_float2(ex2_approx(x.x), ex2_approx(x.y));
                    l2 = f2add(l2, e);
                    pk[i >> 1] = pack_bf16x2(e.x, e.y);
                }
                tmem_st8(tmem_s + c * 8, pk);   // P chunk c (16 keys) -> columns [8 c, 8 c + 8): below the S columns still to be read
            };
#pragma unroll
            for (int c = 0; c < 8; c += 2) {
                tc_wait_ld();
                tmem_ld16(tmem_s + (c + 1) * 16, vb);
                exp_chunk(v, c);
                tc_wait_ld();
                if (c + 2 < 8) tmem_ld16(tmem_s + (c + 2) * 16, v);
                exp_chunk(vb, c + 1);
            }
            s_l[(par * 2 + j) * 128 + row] = l2.x + l2.y;
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&p_full[j * 2 + par]);
                arrive_leader_tmem(&p_ready[j * 2 + par]);
            }
            if (q == 0) AP_TRACE(6 + j, n, 2);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // no CTA leaves while its peer may still copy into its smem, signal its barriers or run pair MMAs on its TMEM
    if (warp == 21) tmem_dealloc2(tmem_base, 512);
}

}  // namespace etude
