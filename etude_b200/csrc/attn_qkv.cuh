// Self-attention of a frequency-axis EncoderLayer with the Q|K|V projection FUSED in: x -> softmax(Q K^T / 8) V per head,
// where Q, K, V = x Wq^T + bq, ... never touch HBM (reference amt_apc.py:342-368: fc_q / fc_k / fc_v, the head split, energy,
// softmax, matmul; sequences of L = 256 tokens = the 256 mel bins of one frame, 4 heads x 64).
//
// Why: as separate kernels the projection is HBM-bound (512 B in + 1536 B of Q|K|V out per token, read again by the
// attention kernel: 4.4 TB/s, 27 % of the step) and the attention kernel MUFU-bound with an idle tensor pipe; fused, the
// projection MMAs of head h + 1 run under the softmax of head h and per token only x (512 B) is read and the context
// (512 B) written.
//
// One CLUSTER OF TWO CTAs per sequence.  CTA r owns tokens [128 r, 128 r + 128): it keeps that half of x resident in smem
// (4 K-chunks [128 x 64] bf16, SW128), projects Q_h | K_h | V_h for its tokens head by head against the head-major weight
// stream (boxes [192 x 64], each CTA loads half of a box and TMA-multicasts it to both: one L2 read per cluster),
// accumulating [128 x 192] in TMEM.  The projection epilogue warps add the bias and write bf16 operands into this CTA's
// smem: Q_h, and the K_h / V_h rows of its 128 tokens, which one thread then copies into the PEER's smem as well with a
// 16 KB shared-to-shared bulk copy (cp.async.bulk.shared::cluster.shared::cta: the DSMEM copy engine, completion counted in
// bytes on the peer's mbarrier).  So each CTA holds K_h, V_h of all 256 keys and runs the attention of its 128 queries:
// S_j = Q K_j^T for the two 128-key blocks into two 128-column TMEM buffers, softmax with one thread per (row, block) against
// a reference COMMON to both blocks (the two threads of a row exchange their block maxima through smem), P bf16 back into
// TMEM over S, and O = P_0 V_0 + P_1 V_1 accumulated by the tensor core in ONE 64-column accumulator; the drain warps scale
// by 1 / (l_0 + l_1) and write bf16 context rows straight to HBM.
//
// What the timelines said (profiles/r2c, r2d, r2e, r2f *_timeline.txt; tests/gpu_diag.py attn_qkv_trace):
//   * per-thread st.shared::cluster stores of the K / V rows cost ~350 clk each and made the epilogue the bottleneck
//     (9 600 clk per head); pulling with ld.shared::cluster is as slow; the bulk copy takes 1 700 - 3 500 clk for 16 KB but
//     runs beside everything else;
//   * gating the projection of head n + 1 behind S(n) serialises projection -> epilogue -> exchange (7 400 clk per head);
//     issued as soon as the accumulator is free it runs under the previous head's softmax by itself;
//   * with the O accumulators inside the two S / P buffers (attention4.cuh's layout) a buffer stays occupied from S until
//     the drain has read O: 5 600 clk per tile, 6 400 clk per head.  Here a buffer is released by the commit of its P V MMAs.
//
// TMEM (512 columns): [0, 192) projection accumulator Q|K|V, [192, 256) O, [256, 384) / [384, 512) the two S / P buffers.
// smem: x 64 KB | W ring 3 x 24 KB | K 32 KB | V 32 KB | Q 16 KB | softmax statistics 4 KB | bias 3 KB | barriers.
// Warps (24): 0 TMA producer (x, W ring), 1 projection MMA issue (+ TMEM alloc), 2 S = Q K^T issue, 3 P V issue,
// 4-11 projection epilogue (lane quarter warp & 3, column half (warp - 4) >> 2), 12-15 drain, 16-19 / 20-23 softmax of KV
// block 0 / 1.  Registers: 768 threads start at 80; setmaxnreg draws from the CTA's own pool (6 x 80 x 128), so the budgets
// sum to <= 480: TMA / MMA 40, epilogue 2 x 80, drain 96, softmax 2 x 88.
//
// Cross-CTA protocol (every barrier completes exactly once per (sequence, head) iteration n; waits use parity n & 1):
//   qk_ready  one local arrive (Q and the own K block written) + 16 KB from the peer's bulk copy   -> S issue
//   v_ready   one local arrive + 16 KB from the peer's bulk copy                                   -> P V issue
//   qk_free   S MMAs of BOTH CTAs complete (tcgen05.commit multicast)   -> both epilogues may overwrite Q / K
//   v_free    P V MMAs of BOTH CTAs complete                            -> both epilogues may overwrite V
//   w_empty   both CTAs' projection MMAs on a ring slot complete        -> both producers may multicast into it
#pragma once
#include "attention4.cuh"
#include "cluster.cuh"
#include "common.cuh"

namespace etude {

constexpr int kAqThreads = 24 * 32;
constexpr int kAqWStages = 3;
constexpr int kAqXBytes = 4 * 128 * 64 * 2;          // 64 KB
constexpr int kAqWStageBytes = 192 * 64 * 2;         // 24 KB: [Q_h | K_h | V_h rows] x 64 input channels
constexpr int kAqWHalfBytes = kAqWStageBytes / 2;    // rows loaded (and multicast) by one CTA
constexpr int kAqKVBytes = 2 * 128 * 64 * 2;         // K (or V) of all 256 keys: two blocks of 128 keys
constexpr int kAqQBytes = 128 * 64 * 2;
constexpr int kAqStatBytes = 2 * 2 * 2 * 128 * 4;    // block maxima and block sums: [2 (n & 1)][2 blocks][128 rows] each
constexpr size_t kAttnQkvSmemBytes = kAqXBytes + kAqWStages * kAqWStageBytes + 2 * kAqKVBytes + kAqQBytes + kAqStatBytes + 768 * 4 + 512;

// Experiments (profiles/r2l_*): AQ_W_MULTICAST = 0 lets every CTA load whole weight boxes itself (no cluster multicast);
// AQ_COPY_SPLIT = n issues the 16 KB K / V exchange as n bulk copies.
#ifndef AQ_W_MULTICAST
#define AQ_W_MULTICAST 1
#endif
#ifndef AQ_COPY_SPLIT
#define AQ_COPY_SPLIT 1
#endif

struct AttnQkvParams {
    int n_seq;                  // sequences of 256 tokens
    const float* bias;          // [4 heads][192]: bq_h | bk_h | bv_h
    __nv_bfloat16* out;         // context [n_seq * 256, 256]
    float scale_log2e;
    long long* trace;           // dev build: clock64 timeline of cluster 0 / rank 0 (8 roles x 64 iterations x 8 events), or nullptr
};

// Debug timeline (libetude_b200_dev.so only sets p.trace): role r, iteration n < 64, event e < 8
#define AQ_TRACE(role, n, e)                                                                        \
    do {                                                                                            \
        if (p.trace != nullptr && blockIdx.x == 0 && lane == 0 && (n) < 64)                         \
            p.trace[(((role) * 64 + (n)) << 3) + (e)] = clock64();                                  \
    } while (0)

// 16-byte-granular shared -> peer-shared bulk copy; the peer's mbarrier receives complete_tx(bytes)
__device__ __forceinline__ void dsmem_bulk_copy(uint32_t dst_cluster_addr, uint32_t src_cta_addr, uint32_t bytes, uint32_t bar_cluster_addr) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster_addr),
                 "r"(src_cta_addr), "r"(bytes), "r"(bar_cluster_addr)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// wait with cluster-scope acquire: the barrier guards data written (or smem reads completed) by the peer CTA
__device__ __forceinline__ void mbar_wait_cl(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    auto try_once = [&]() -> bool {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        return ok != 0;
    };
    if (try_once()) return;
    const unsigned long long t0 = globaltimer_ns();
#pragma unroll 1
    for (;;) {
#pragma unroll 1
        for (uint32_t i = 0; i < 4096u; ++i)
            if (try_once()) return;
        if (globaltimer_ns() - t0 > kWaitBudgetNs) break;
    }
    __trap();
}

// The product runs attn_pair.cuh; this kernel (an independent implementation of the same operator) is compiled into the
// test-only library, where tests/gpu_diag.py holds the two against each other, and into -DETUDE_ATTN_PAIR=0 A/B builds.
#ifndef ETUDE_ATTN_PAIR
#define ETUDE_ATTN_PAIR 1
#endif
#if defined(ETUDE_DEV_BUILD) || !ETUDE_ATTN_PAIR
#define ETUDE_HAVE_ATTN_QKV_CTA1 1
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kAqThreads, 1)
attn_qkv_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const AttnQkvParams p) {
    constexpr int O_COL = 192, BUF0_COL = 256, BUF_COLS = 128;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
    uint8_t* sX = smem;
    uint8_t* sW = sX + kAqXBytes;
    uint8_t* sK = sW + kAqWStages * kAqWStageBytes;
    uint8_t* sV = sK + kAqKVBytes;
    uint8_t* sQ = sV + kAqKVBytes;
    float* s_mx = reinterpret_cast<float*>(sQ + kAqQBytes);  // [n & 1][2 blocks][128] raw row maximum of the block
    float* s_l = s_mx + 2 * 2 * 128;                          // [n & 1][2 blocks][128] block sum of 2^(s * scale - m)
    float* s_bias = s_l + 2 * 2 * 128;                        // [4][192]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + 768);
    uint64_t* w_full = bars;                    // [3]
    uint64_t* w_empty = w_full + kAqWStages;    // [3] count 2 (one projection-MMA commit per CTA)
    uint64_t* x_full = w_empty + kAqWStages;
    uint64_t* x_free = x_full + 1;
    uint64_t* acc_full = x_free + 1;            // projection accumulator complete (MMA commit -> epilogue)
    uint64_t* acc_free = acc_full + 1;          // count 8 (epilogue warps -> projection MMA)
    uint64_t* qk_ready = acc_free + 1;          // count 1 + 16 KB of transaction bytes
    uint64_t* qk_free = qk_ready + 1;           // count 2
    uint64_t* v_ready = qk_free + 1;            // count 1 + 16 KB of transaction bytes
    uint64_t* v_free = v_ready + 1;             // count 2
    uint64_t* s_full = v_free + 1;              // [2] S_j complete                  (MMA commit -> softmax group j)
    // [2 blocks][2 (n & 1)] count 4: P_j and its row sum written (softmax warps -> P V issue, drain).  One barrier per
    // iteration parity: the drain is not on the path that produces the next P (S(n + 1) only needs P V(n) to complete), so
    // with a single barrier per block P(n + 1) could complete before the drain has observed P(n) and its parity wait would
    // alias; a barrier that completes every second iteration cannot get two completions ahead (P(n + 2) needs o_free(n)).
    uint64_t* p_full = s_full + 2;
    uint64_t* buf_free = p_full + 4;            // [2] P_j V_j complete: buffer j reusable    (MMA commit -> S issue)
    uint64_t* o_full = buf_free + 2;            // O = P_0 V_0 + P_1 V_1 complete             (MMA commit -> drain)
    uint64_t* o_free = o_full + 1;              // count 4: O read                            (drain warps -> P V issue)
    uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(o_free + 1);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank(), peer = rank ^ 1u;
    const int cid = (int)cluster_id_x(), ncl = (int)cluster_nctaid_x();
    const int my_items = (cid < p.n_seq) ? (p.n_seq - 1 - cid) / ncl + 1 : 0;
    const int N = my_items * 4;   // (sequence, head) iterations
    constexpr uint16_t kBoth = 3;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_w);
        for (int s = 0; s < kAqWStages; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], AQ_W_MULTICAST ? 2 : 1); }
        mbar_init(x_full, 1); mbar_init(x_free, 1);
        mbar_init(acc_full, 1); mbar_init(acc_free, 8);
        mbar_init(qk_ready, 1); mbar_init(qk_free, 2);
        mbar_init(v_ready, 1); mbar_init(v_free, 2);
        for (int b = 0; b < 2; ++b) { mbar_init(&s_full[b], 1); mbar_init(&buf_free[b], 1); }
        for (int b = 0; b < 4; ++b) mbar_init(&p_full[b], 4);
        mbar_init(o_full, 1); mbar_init(o_free, 4);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_base_ptr, 512);
    for (int i = threadIdx.x; i < 768; i += kAqThreads) s_bias[i] = p.bias[i];
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // the peer's barriers exist before anything is multicast into / signalled in this CTA
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_ptr;

    if (warp < 4) {
      reg_dec<40>();
      if (warp == 0) {
        // ===================================================== TMA producer: x half-tile per sequence, head-major W stream
        const bool leader = elect_one();
        uint32_t c = 0;   // ring counter
        for (int il = 0; il < my_items; ++il) {
            const int seq = cid + il * ncl;
            const int row0 = seq * 256 + (int)rank * 128;
            mbar_wait_inl(x_free, (il & 1) ^ 1);
            AQ_TRACE(0, il * 4, 0);
            if (leader) {
                mbar_expect_tx(x_full, kAqXBytes);
#pragma unroll
                for (int kc = 0; kc < 4; ++kc) tma_load_2d(sX + kc * 16384, &tmap_x, x_full, kc * 64, row0);
            }
            for (int hk = 0; hk < 16; ++hk, ++c) {   // (head, K-chunk) boxes in consumption order
                const uint32_t s = c % kAqWStages;
                mbar_wait_cl(&w_empty[s], ((c / kAqWStages) & 1) ^ 1);
                AQ_TRACE(0, il * 4 + (hk >> 2), 1 + (hk & 3));   // W box (head, chunk) issued
                if (leader) {
                    mbar_expect_tx(&w_full[s], kAqWStageBytes);
#if AQ_W_MULTICAST
                    tma_load_2d_mc(sW + s * kAqWStageBytes + rank * kAqWHalfBytes, &tmap_w, &w_full[s], (hk & 3) * 64,
                                   (hk >> 2) * 192 + (int)rank * 96, kBoth);
#else
                    tma_load_2d(sW + s * kAqWStageBytes, &tmap_w, &w_full[s], (hk & 3) * 64, (hk >> 2) * 192);
                    tma_load_2d(sW + s * kAqWStageBytes + kAqWHalfBytes, &tmap_w, &w_full[s], (hk & 3) * 64, (hk >> 2) * 192 + 96);
#endif
                }
            }
            __syncwarp();
        }
      } else if (warp == 1) {
        // ===================================================== projection MMA issue: ACC[128 x 192] = x_half W_h^T
        const bool leader = elect_one();
        constexpr uint32_t idesc_p = make_idesc_bf16(128, 192, 0, 0);
        const uint64_t x_desc0 = make_sw128_desc(smem_u32(sX));
        const uint64_t w_desc0 = make_sw128_desc(smem_u32(sW));
        uint32_t c = 0;
        for (int n = 0; n < N; ++n) {
            const int h = n & 3;
            if (h == 0) mbar_wait_inl(x_full, (n >> 2) & 1);
            mbar_wait_inl(acc_free, (n & 1) ^ 1);
            tc_fence_after();
            AQ_TRACE(1, n, 0);
#pragma unroll 1
            for (int kc = 0; kc < 4; ++kc, ++c) {
                const uint32_t s = c % kAqWStages;
                mbar_wait_inl(&w_full[s], (c / kAqWStages) & 1);
                tc_fence_after();
                AQ_TRACE(1, n, 1 + kc);   // W chunk kc landed
                if (leader) {
                    const uint64_t ad = x_desc0 + (uint64_t)(kc * (16384 >> 4)), bd = w_desc0 + (uint64_t)(s * (kAqWStageBytes >> 4));
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base, ad + 2 * k, bd + 2 * k, idesc_p, (kc | k) ? 1u : 0u);
#if AQ_W_MULTICAST
                    tc_commit_mc(&w_empty[s], kBoth);
#else
                    tc_commit(&w_empty[s]);
#endif
                }
                __syncwarp();
            }
            if (leader) {
                tc_commit(acc_full);
                if (h == 3) tc_commit(x_free);
            }
            if (p.trace != nullptr) {   // diagnostic only (serialises the issue loop): when does the accumulator complete?
                mbar_wait_inl(acc_full, n & 1);
                AQ_TRACE(1, n, 5);
            }
            __syncwarp();
        }
      } else if (warp == 2) {
        // ===================================================== S_j = Q K_j^T issue (two 128-key blocks -> buffers 0 / 1)
        const bool leader = elect_one();
        constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
        const uint64_t q_desc = make_sw128_desc(smem_u32(sQ));
        const uint64_t k_desc0 = make_sw128_desc(smem_u32(sK));
        for (int n = 0; n < N; ++n) {
            mbar_wait_cl(qk_ready, n & 1);
            fence_proxy_async_all();
            AQ_TRACE(2, n, 0);
#pragma unroll 1
            for (int j = 0; j < 2; ++j) {
                mbar_wait_inl(&buf_free[j], (n & 1) ^ 1);
                tc_fence_after();
                AQ_TRACE(2, n, 1 + j);
                if (leader) {
                    const uint64_t kd = k_desc0 + (uint64_t)(j * (16384 >> 4));
                    const uint32_t tmem_s = tmem_base + BUF0_COL + j * BUF_COLS;
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_s, q_desc + 2 * k, kd + 2 * k, idesc_s, k != 0);
                    tc_commit(&s_full[j]);
                    if (j == 1) tc_commit_mc(qk_free, kBoth);   // every S MMA of this head has been issued before this commit
                }
                __syncwarp();
            }
        }
      } else {
        // ===================================================== O = P_0 V_0 + P_1 V_1 issue (one accumulator)
        const bool leader = elect_one();
        constexpr uint32_t idesc_o = make_idesc_bf16(128, kHeadDim, 0, 1);  // B = V is MN-major (d contiguous)
        const uint64_t v_desc0 = make_sw128_desc(smem_u32(sV), 8192);
        for (int n = 0; n < N; ++n) {
            mbar_wait_cl(v_ready, n & 1);
            fence_proxy_async_all();
            mbar_wait_inl(o_free, (n & 1) ^ 1);   // the drain warps have read the previous head's O
            AQ_TRACE(3, n, 0);
#pragma unroll 1
            for (int j = 0; j < 2; ++j) {
                mbar_wait_inl(&p_full[j * 2 + (n & 1)], (n >> 1) & 1);
                tc_fence_after();
                AQ_TRACE(3, n, 1 + j);
                if (leader) {
                    const uint64_t vd = v_desc0 + (uint64_t)(j * (16384 >> 4));
                    const uint32_t tmem_p = tmem_base + BUF0_COL + j * BUF_COLS;
#pragma unroll
                    for (int s = 0; s < 8; ++s)  // P: bf16 pairs, 8 columns per K = 16; V: 16 keys = 2048 B further
                        umma_bf16_ts(tmem_base + O_COL, tmem_p + s * 8, vd + (uint64_t)(s * 128), idesc_o, (j | s) ? 1u : 0u);
                    tc_commit(&buf_free[j]);
                    if (j == 1) {
                        tc_commit(o_full);
                        tc_commit_mc(v_free, kBoth);
                    }
                }
                __syncwarp();
            }
        }
      }
    } else if (warp < 12) {
        // ===================================================== projection epilogue (8 warps): ACC + bias -> bf16 Q, K, V rows in smem;
        // the K / V blocks of this CTA's 128 tokens are then bulk-copied into the peer's smem by one thread
        const int q = warp & 3, half = (warp - 4) >> 2;   // TMEM lane quarter, 32-column half of each of Q / K / V
        const int row = q * 32 + lane;                     // token of this CTA's half = TMEM lane
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int sw = row & 7;
        const uint32_t q_row = smem_u32(sQ) + row * 128;
        const uint32_t k_blk = smem_u32(sK) + rank * 16384, v_blk = smem_u32(sV) + rank * 16384;   // block `rank` = this CTA's keys
        const uint32_t k_row = k_blk + row * 128, v_row = v_blk + row * 128;
        const uint32_t k_blk_peer = mapa_u32(k_blk, peer), v_blk_peer = mapa_u32(v_blk, peer);
        const uint32_t qk_ready_peer = mapa_u32(smem_u32(qk_ready), peer), v_ready_peer = mapa_u32(smem_u32(v_ready), peer);
        const bool copier = (warp == 4) && elect_one();
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
            uint4 pq[4], pk[4];
            mbar_wait_inl(acc_full, n & 1);
            __syncwarp();
            tc_fence_after();
            if (warp == 4) AQ_TRACE(4, n, 0);
            // Q and K are packed into registers BEFORE the wait for the previous head's S: that wait ends the head's
            // critical path (S(n-1) done -> K(n) published -> exchange -> S(n)), so only the stores remain behind it
            load_pack(half, bias, pq);          // Q columns [32 half, + 32)
            load_pack(2 + half, bias, pk);      // K
            mbar_wait_cl(qk_free, (n & 1) ^ 1);   // the S MMAs of the previous head (both CTAs) have read Q / K
            if (warp == 4) AQ_TRACE(4, n, 1);
            st_row(q_row, pq);
            st_row(k_row, pk);
            fence_async_smem();                 // generic-proxy writes -> visible to the MMAs and to the bulk copy (async proxy)
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (copier) {
                mbar_expect_tx(qk_ready, 16384);                       // arrive + the peer's K block on its way into this CTA
#pragma unroll
                for (int i = 0; i < AQ_COPY_SPLIT; ++i)
                    dsmem_bulk_copy(k_blk_peer + i * (16384 / AQ_COPY_SPLIT), k_blk + i * (16384 / AQ_COPY_SPLIT), 16384 / AQ_COPY_SPLIT, qk_ready_peer);
            }
            if (warp == 4) AQ_TRACE(4, n, 2);
            load_pack(4 + half, bias, pk);      // V
            tc_fence_before();                  // the accumulator has been read: hand it back to the projection MMA warp
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_free);
            mbar_wait_cl(v_free, (n & 1) ^ 1);    // the P V MMAs of the previous head (both CTAs) have read V
            if (warp == 4) AQ_TRACE(4, n, 3);
            st_row(v_row, pk);
            fence_async_smem();
            asm volatile("bar.sync 2, 256;" ::: "memory");
            if (copier) {
                mbar_expect_tx(v_ready, 16384);
#pragma unroll
                for (int i = 0; i < AQ_COPY_SPLIT; ++i)
                    dsmem_bulk_copy(v_blk_peer + i * (16384 / AQ_COPY_SPLIT), v_blk + i * (16384 / AQ_COPY_SPLIT), 16384 / AQ_COPY_SPLIT, v_ready_peer);
            }
            if (warp == 4) AQ_TRACE(4, n, 4);
        }
    } else if (warp < 16) {
        // ===================================================== drain: O / (l_0 + l_1) -> bf16 context rows -> HBM
        reg_inc<96>();
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        float acc[64];
        for (int n = 0; n < N; ++n) {
            const uint32_t ph = n & 1;
            mbar_wait_inl(&p_full[0 + ph], (n >> 1) & 1);   // the row sums of both blocks are visible
            mbar_wait_inl(&p_full[2 + ph], (n >> 1) & 1);
            mbar_wait_inl(o_full, ph);
            __syncwarp();
            tc_fence_after();
            if (q == 0) AQ_TRACE(5, n, 0);
            const uint32_t tmem_o = tmem_base + O_COL + lane_off;
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_ld16(tmem_o + c * 16, acc + c * 16);   // all four loads in flight
            const float inv = rcp_fma(s_l[(ph * 2 + 0) * 128 + row] + s_l[(ph * 2 + 1) * 128 + row]);   // before o_free: the slot is rewritten at n + 2
            tc_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(o_free);
            if (q == 0) AQ_TRACE(5, n, 1);
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
            if (q == 0) AQ_TRACE(5, n, 2);
        }
    } else {
        // ===================================================== softmax of KV block j (warps 16-19: j = 0, 20-23: j = 1): one thread per
        // (query row, block).  Both blocks use ONE reference -- the two threads of a row swap their block maxima through smem
        // behind a 64-thread named barrier -- so that P_0 V_0 and P_1 V_1 can accumulate into a single O.
        reg_inc<88>();
        const int j = (warp - 16) >> 2;
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const float scale = p.scale_log2e;
        const uint32_t tmem_s = tmem_base + BUF0_COL + j * BUF_COLS + lane_off;
        const int pair_bar = 3 + q;   // named barrier of the two warps that own the same 32 rows
        float v[16], vb[16];
        for (int n = 0; n < N; ++n) {
            const int par = n & 1;
            mbar_wait_inl(&s_full[j], par);
            __syncwarp();
            tc_fence_after();
            if (q == 0) AQ_TRACE(6 + j, n, 0);
            // Both passes are software pipelined over two 16-column register buffers: the tcgen05.ld of chunk c + 1 is in
            // flight while chunk c is processed (tcgen05.wait::ld covers every outstanding load of the thread, so it sits
            // right before the next load is issued).  S(n) -> softmax -> P V(n) -> S(n + 1) is the critical cycle of a head
            // (the S buffers are single), so the exposed TMEM round trips of this warp were head-period clocks.
            // ---- pass 1: row maximum of this block
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
            const float mx = fmaxf(m0, m1);
            s_mx[(par * 2 + j) * 128 + row] = mx;
            asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
            const float m_sc = ceilf(fmaxf(mx, s_mx[(par * 2 + (j ^ 1)) * 128 + row]) * scale);   // integer reference >= the row maximum
            if (q == 0) AQ_TRACE(6 + j, n, 1);
            // ---- pass 2: p = 2^(s * scale - m) -> bf16 P over the S columns already consumed; block row sum
            float2 l2 = make_float2(0.f, 0.f);
            const float2 sc2 = make_float2(scale, scale), nm2 = make_float2(-m_sc, -m_sc);
            auto exp_chunk = [&](const float* w, int c) {
                uint32_t pk[8];
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const float2 x = f2fma(make_float2(w[i], w[i + 1]), sc2, nm2);
                    const float2 e = make_float2(ex2_approx(x.x), ex2_approx(x.y));
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
            if (lane == 0) mbar_arrive(&p_full[j * 2 + par]);
            if (q == 0) AQ_TRACE(6 + j, n, 2);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // no CTA leaves while its peer may still write into its smem or signal its barriers
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}
#endif  // ETUDE_DEV_BUILD || !ETUDE_ATTN_PAIR

}  // namespace etude
