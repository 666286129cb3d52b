// Fused token-local chain of one transformer sub-block on tcgen05 (third generation):
//
//     y   = LayerNorm(ctx @ Wo^T + bo + x)                 attention output projection + residual + LN
//     out = LayerNorm(y + relu(y @ W1^T + b1) @ W2^T + b2) position-wise FFN + residual + LN (same LN module)
//
// (reference amt_apc.py:250-258, 276-284, 310-318 with fc_o of amt_apc.py:371 and the FFN of 383-392).
//
// Structure kept from the second generation: clusters of TWO CTAs that fetch every weight box from L2 once (each CTA loads
// half of the box and TMA-multicasts it into both rings), N = 256 MMAs against adjacent ring slot pairs, the ReLU'd hidden
// chunk parked in TMEM over its own accumulator and read back by FFN2 as a TMEM A operand, 16 epilogue warps (TMEM lane
// quarter x column quarter), cta_group::1 MMAs with CTA-local MMA <-> epilogue handshakes.
//
// What changed, and why (timeline profiles/r1f_chain2_timeline.txt: a 27.4 k clk tile period = G1 || LN2 6 k + E1 7 k exposed
// + FFN 13 k; the FFN window was paced by the 8-slot weight ring, 32 boxes at ~400 clk each, and G1 by the four boxes
// that did not fit ahead of it):
//   * the ring has TEN slots: the 32 KB output staging buffer is gone (the normalised rows are staged in Y, which is
//     dead between the last FFN1 MMA and the next tile's y), and the residual tile no longer has a buffer of its own: its
//     four boxes travel through the ring right behind the G1 operands, E1 reads them from their slots and releases the
//     slots in both CTAs (remote mbarrier arrive).  (Reading the residual rows with per-thread global loads was measured
//     first: 128 B per thread and row is 32 L1 wavefronts per load instruction, 7.7 k clk per tile.);
//   * E1 is ONE pass over TMEM: the pre-norm row stays in registers across the statistics exchange (setmaxnreg gives the
//     epilogue warps 112 registers), so the park-and-reload round trip through TMEM is gone;
//   * the elementwise math runs as packed fp32 pairs (FFMA2 / FADD2 / FMUL2): half the issue slots per element, and ReLU
//     is folded into the bf16 conversion (cvt.rn.relu.bf16x2.f32);
//   * the statistics exchange of a lane quarter lives in the 1 KB of Y that the same quarter overwrites next.
//
// TMEM (512 columns) = two 256-column regions R0/R1 whose roles swap every tile (parity p = tile & 1):
//     D1   = R[p]   : ctx Wo^T accumulator -> y + b2 -> + FFN2 accumulation -> LN2 input
//     ACC2 = R[p^1] : two 128-column FFN1 chunk accumulators (chunk j -> half j & 1)
// Warps (20): 0 = ring producer (ctx + weights + residual), 1 = MMA issuer + TMEM allocator, 2 = frees the residual slots
// once the epilogue warps have read them, 3 = issues the TMA output store once they have staged the rows (the epilogue
// warps never meet at a 512-thread barrier: they hand over through mbarriers and run on),
// 4..19 = epilogue (warp 4 + e: TMEM lane quarter e & 3, column quarter e >> 2).  Registers: the CTA launches at 96 per
// thread; setmaxnreg moves them inside the CTA's own pool (640 x 96 = 128 x 32 + 512 x 112).
// smem: ring 10 x 16 KB ([128 x 64] bf16 boxes, SW128; weight boxes that form one N = 256 operand sit in an even/odd
//       slot pair) | Y 64 KB (y bf16: FFN1 A operand; then the staging of the TMA output store; its first 1 KB per lane
//       quarter doubles as the LayerNorm statistics exchange) | barriers | LayerNorm gamma, beta (2 KB).
#pragma once
#include "cluster.cuh"
#include "common.cuh"

namespace etude {

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

// Debug timeline: role r (0 MMA thread, 1 epilogue warp 4 lane 0, 2 ring producer) appends (event id, clock64) pairs.
constexpr int kChTraceSlots = 512;
#define CH_TRACE(role, id)                                                                         \
    do {                                                                                           \
        if (p.trace != nullptr && blockIdx.x == 0 && tr_n < kChTraceSlots) {                        \
            p.trace[((role) * kChTraceSlots + tr_n) * 2] = (id);                                    \
            p.trace[((role) * kChTraceSlots + tr_n) * 2 + 1] = clock64();                           \
            ++tr_n;                                                                                \
        }                                                                                          \
    } while (0)

constexpr int kC3Cluster = 2;
constexpr int kC3Threads = 20 * 32;
constexpr int kC3Stages = 10;
constexpr int kC3StageBytes = 128 * 64 * 2;            // 16 KB
constexpr int kC3PartBytes = kC3StageBytes / kC3Cluster;  // rows of a weight box loaded (and multicast) by one CTA
constexpr int kC3PartRows = 128 / kC3Cluster;
constexpr int kC3YBytes = 4 * kC3StageBytes;
// No 1 KB alignment slack: the dynamic window is declared 1024-aligned (checked at kernel entry).
// + LayerNorm gamma | beta (2 KB): as L1-resident global loads they sat on the dependency chain of both normalisation
// passes (887 vs 806 TFLOP/s stand-alone with the two vectors in shared memory, profiles/r2q_*).  The CTA now leaves no
// room for a second block on its SM; the end-to-end rate stayed at 99.5 % of the device-resident one.
constexpr size_t kChain3SmemBytes = kC3Stages * kC3StageBytes + kC3YBytes + 512 + 2048;

// (hi, lo) -> packed bf16 pair with ReLU folded into the conversion
__device__ __forceinline__ uint32_t pack_bf16x2_relu(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}

template <bool FFN>
__global__ void __cluster_dims__(kC3Cluster, 1, 1) __launch_bounds__(kC3Threads, 1)
chain3_kernel(const __grid_constant__ CUtensorMap tmap_ctx, const __grid_constant__ CUtensorMap tmap_wo,
              const __grid_constant__ CUtensorMap tmap_w1, const __grid_constant__ CUtensorMap tmap_w2,
              const __grid_constant__ CUtensorMap tmap_resid, const __grid_constant__ CUtensorMap tmap_out, const ChainParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();  // SW128 boxes and the multicast ring need the 1 KB alignment
    uint8_t* sRing = smem;
    uint8_t* sY = sRing + kC3Stages * kC3StageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sY + kC3YBytes);
    uint64_t* full = bars;                 // [10] ring: TMA (own part + the peers' multicast parts) -> MMA
    uint64_t* empty = full + kC3Stages;    // [10] ring: one MMA commit per CTA of the cluster -> producer
    uint64_t* g1_full = empty + kC3Stages; // [2] ctx Wo^T accumulator complete    (MMA commit -> epilogue), per region
    uint64_t* e1_done = g1_full + 2;       // y in smem, y + b2 in TMEM            (16 epilogue warps -> MMA)
    uint64_t* f1_full = e1_done + 1;       // [2] FFN1 chunk accumulator complete  (MMA commit -> epilogue)
    // [2] hidden chunk parked in TMEM, per ACC2 half (16 epilogue warps -> MMA).  One barrier per half, NOT one shared
    // barrier: E2(j+1) does not depend on FFN2(j), so with a single barrier its arrivals could complete a second
    // phase before the MMA warp has observed the first one and the parity wait would miss it (seen as a hang).
    uint64_t* h_full = f1_full + 2;
    uint64_t* f2_full = h_full + 2;        // FFN2 accumulation complete           (MMA commit -> epilogue)
    uint64_t* qfree = f2_full + 1;         // [4] 128-column TMEM quarter drained  (16 epilogue warps -> MMA)
    uint64_t* d1_done = qfree + 4;         // y + b2 in TMEM                       (16 epilogue warps -> MMA, before FFN2(0))
    uint64_t* resid_read = d1_done + 1;    // residual boxes consumed              (16 epilogue warps -> warp 2, which frees the slots)
    uint64_t* staged = resid_read + 1;     // output rows staged in Y              (16 epilogue warps -> warp 3, which stores them)
    uint64_t* y_free = staged + 1;         // the TMA store has finished reading Y (warp 3 -> epilogue warps)
    uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(y_free + 1);
    float* s_gamma = reinterpret_cast<float*>(sY + kC3YBytes + 512);
    float* s_beta = s_gamma + 256;
    if (threadIdx.x < 256) { s_gamma[threadIdx.x] = p.gamma[threadIdx.x]; s_beta[threadIdx.x] = p.beta[threadIdx.x]; }

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform: role branches stay uniform
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cid = (int)cluster_id_x(), ncl = (int)cluster_nctaid_x();
    // tile pairs are dealt to clusters round-robin; rank r of the cluster takes tile 2 * pair + r.  Both CTAs of a
    // cluster run the same number of iterations (the weight ring is shared); a tile index past the end is a dummy
    // whose loads are zero-filled and whose stores are masked.
    const int n_pairs = (p.num_tiles + kC3Cluster - 1) / kC3Cluster;
    const int my_iters = (cid < n_pairs) ? (n_pairs - 1 - cid) / ncl + 1 : 0;
    constexpr uint16_t kAllMask = (1u << kC3Cluster) - 1;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_ctx); tma_prefetch_desc(&tmap_wo); tma_prefetch_desc(&tmap_w1);
        tma_prefetch_desc(&tmap_w2); tma_prefetch_desc(&tmap_resid); tma_prefetch_desc(&tmap_out);
        for (int s = 0; s < kC3Stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kC3Cluster); }
        mbar_init(&g1_full[0], 1); mbar_init(&g1_full[1], 1); mbar_init(e1_done, 16);
        mbar_init(&f1_full[0], 1); mbar_init(&f1_full[1], 1);
        mbar_init(&h_full[0], 16); mbar_init(&h_full[1], 16); mbar_init(f2_full, 1);
        for (int q = 0; q < 4; ++q) mbar_init(&qfree[q], 16);
        mbar_init(d1_done, 16); mbar_init(resid_read, 16); mbar_init(staged, 16); mbar_init(y_free, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_base_ptr, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // the peer's barriers exist before anything is multicast into this CTA
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_ptr;

    // Producer and MMA roles run WARP-UNIFORM: every lane executes the loops and the waits, one elected lane issues.
    // With the loop confined to `if (lane == 0)` the compiler cannot keep descriptors / addresses in uniform registers
    // and spends ~20 instructions + 5 R2UR moves per tcgen05.mma (measured 178 clk per MMA whatever its shape,
    // tests/gpu_diag.py mma_bench); uniform code issues UTCHMMAs back to back from descriptors that differ by an add.
    if (warp < 4) {
      reg_dec<32>();
      if (warp == 0) {
        // ===================================================== ring producer: ctx (own) + weight boxes (shared) in consumption order
        const bool leader = elect_one();
        uint32_t c = 0;
        int tr_n = leader ? 0 : kChTraceSlots;
        auto slot_acquire = [&]() -> uint32_t {
            const uint32_t s = c % kC3Stages;
            mbar_wait_inl(&empty[s], ((c / kC3Stages) & 1) ^ 1);
            if (leader) mbar_expect_tx(&full[s], kC3StageBytes);
            ++c;
            return s;
        };
        auto load_a = [&](int col, int row) {  // this CTA's ctx k-block
            const uint32_t s = slot_acquire();
            if (leader) tma_load_2d(sRing + s * kC3StageBytes, &tmap_ctx, &full[s], col, row);
        };
        auto load_w = [&](const CUtensorMap* m, int col, int row) {  // my rows of a weight box, to every CTA of the cluster
            const uint32_t s = slot_acquire();
            if (leader)
                tma_load_2d_mc(sRing + s * kC3StageBytes + rank * kC3PartBytes, m, &full[s], col, row + (int)rank * kC3PartRows, kAllMask);
        };
        auto load_r = [&](int col, int row) {  // this CTA's residual box (consumed and released by the epilogue warps)
            const uint32_t s = slot_acquire();
            if (leader) tma_load_2d(sRing + s * kC3StageBytes, &tmap_resid, &full[s], col, row);
        };
        auto load_w1 = [&](int j) { for (int kb = 0; kb < 4; ++kb) load_w(&tmap_w1, kb * 64, j * 128); };
        auto load_w2 = [&](int j) {
            for (int kk = 0; kk < 2; ++kk) { load_w(&tmap_w2, j * 128 + kk * 64, 0); load_w(&tmap_w2, j * 128 + kk * 64, 128); }
        };
        for (int it = 0; it < my_iters; ++it) {
            const int row0 = ((cid + it * ncl) * kC3Cluster + (int)rank) * 128;
            CH_TRACE(2, it * 100);
            for (int h = 0; h < 2; ++h) {  // first GEMM: two ctx k-blocks, then their Wo k-blocks as (rows 0-127, rows 128-255) slot pairs
                load_a((2 * h) * 64, row0);
                load_a((2 * h + 1) * 64, row0);
                for (int kb = 2 * h; kb < 2 * h + 2; ++kb) { load_w(&tmap_wo, kb * 64, 0); load_w(&tmap_wo, kb * 64, 128); }
            }
            {
                const int rrow0 = p.resid_mod ? row0 % p.resid_mod : row0;
                for (int kb = 0; kb < 4; ++kb) load_r(kb * 64, rrow0);
            }
            CH_TRACE(2, it * 100 + 1);
            // FFN, in the MMA warp's software-pipelined order F1(0) F1(1) F2(0) F1(2) F2(1) F1(3) F2(2) F2(3)
            if constexpr (FFN) { load_w1(0); load_w1(1); load_w2(0); load_w1(2); load_w2(1); load_w1(3); load_w2(2); load_w2(3); }
            CH_TRACE(2, it * 100 + 2);
        }
      } else if (warp == 1) {
        // ===================================================== MMA issuer (warp-uniform, one elected lane issues)
        const bool leader = elect_one();
        constexpr uint32_t idesc128 = make_idesc_bf16(128, 128, 0, 0);
        constexpr uint32_t idesc256 = make_idesc_bf16(128, 256, 0, 0);
        // descriptors differ from these bases by adds: + 1024 per 16 KB ring slot / Y box, + 2 per K = 16 step (32 B)
        const uint64_t ring_desc0 = make_sw128_desc(smem_u32(sRing));
        const uint64_t y_desc0 = make_sw128_desc(smem_u32(sY));
        uint32_t c = 0;          // ring consumption counter
        uint32_t prod_par = 0;   // bit q: parity of the productions into TMEM quarter q so far
        uint32_t n_h = 0;        // hidden chunks consumed so far (h_full phase)
        int tr_n = leader ? 0 : kChTraceSlots;
        auto acquire = [&]() -> uint32_t {
            const uint32_t s = c % kC3Stages;
            mbar_wait_inl(&full[s], (c / kC3Stages) & 1);
            ++c;
            return s;
        };
        auto slot_desc = [&](uint32_t s) -> uint64_t { return ring_desc0 + (uint64_t)(s * (kC3StageBytes >> 4)); };
        auto wait_quarter = [&](int q) {
            mbar_wait_inl(&qfree[q], ((prod_par >> q) & 1) ^ 1);
            prod_par ^= 1u << q;
        };

        for (int it = 0; it < my_iters; ++it) {
            const int par = it & 1;
            const uint32_t d1 = tmem_base + par * 256;         // region R[par]
            const uint32_t a2 = tmem_base + (par ^ 1) * 256;   // region R[par ^ 1]
            const int qd = par * 2, qa = (par ^ 1) * 2;        // first quarter index of each region
            // ---- G1: D1 = ctx Wo^T  (N = 256 MMAs: B = an even/odd slot pair)
            CH_TRACE(0, it * 100);
            wait_quarter(qd);
            wait_quarter(qd + 1);
            tc_fence_after();
            CH_TRACE(0, it * 100 + 1);
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                const uint32_t sa0 = acquire(), sa1 = acquire();
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    const uint32_t sa = kk ? sa1 : sa0;
                    const uint32_t sb0 = acquire(), sb1 = acquire();
                    tc_fence_after();
                    if (leader) {
                        const uint64_t ad = slot_desc(sa), bd = slot_desc(sb0);
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_bf16_ss(d1, ad + 2 * k, bd + 2 * k, idesc256, (h | kk | k) ? 1u : 0u);
                        tc_commit_mc(&empty[sa], kAllMask); tc_commit_mc(&empty[sb0], kAllMask); tc_commit_mc(&empty[sb1], kAllMask);
                    }
                    __syncwarp();
                }
            }
            if (leader) tc_commit(&g1_full[par]);
            // The four residual boxes are consumed by the epilogue warps, but this warp still has to SEE them land before it
            // moves on: ten positions later the same slots hold G1 operands of the next tile, and a parity wait on full[s] may
            // only target phase n once phase n - 1 is known complete.  Skipping them (c += 4) let this warp -- which runs up to
            // two tiles ahead of the epilogue when there is no FFN -- wait for the next tile's operands while the residual box
            // was still in flight (ctx fresh in L2 from the attention kernel, the residual rows not): the wait then returned at
            // once on the stale phase, the MMAs read a slot that had not been written and released it early.  Seen as one
            // CTA's second and third tiles wrong in ~1 of 60 forward passes (tests/race_stress_diag.py).
            for (int r = 0; r < 4; ++r) (void)acquire();
            CH_TRACE(0, it * 100 + 2);
            if constexpr (!FFN) continue;
            // ---- FFN, software pipelined
            auto f1 = [&](int j) {  // ACC2[j & 1] = y W1_j^T
                wait_quarter(qa + (j & 1));
                tc_fence_after();
                CH_TRACE(0, it * 100 + 10 + j);
#pragma unroll
                for (int kb = 0; kb < 4; ++kb) {
                    const uint32_t s = acquire();
                    tc_fence_after();
                    if (leader) {
                        const uint64_t ad = y_desc0 + (uint64_t)(kb * (kC3StageBytes >> 4)), bd = slot_desc(s);
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_bf16_ss(a2 + (j & 1) * 128, ad + 2 * k, bd + 2 * k, idesc128, (kb | k) ? 1u : 0u);
                        tc_commit_mc(&empty[s], kAllMask);
                    }
                    __syncwarp();
                }
                if (leader) tc_commit(&f1_full[j & 1]);
                CH_TRACE(0, it * 100 + 20 + j);
            };
            auto f2 = [&](int j) {  // D1 += h_j W2[:, 128 j ..]^T   (D1 already holds y + b2; h_j sits in TMEM)
                mbar_wait_inl(&h_full[n_h & 1], (n_h >> 1) & 1);  // chunk j = n_h & 3 lives in half j & 1
                ++n_h;
                tc_fence_after();
                CH_TRACE(0, it * 100 + 30 + j);
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    const uint32_t s0 = acquire(), s1 = acquire();  // slot pair = W2[:, 128 j + 64 kk ..] for all 256 outputs
                    tc_fence_after();
                    if (leader) {
                        const uint64_t bd = slot_desc(s0);
                        const uint32_t at = a2 + (j & 1) * 128 + kk * 32;  // bf16 pairs: 8 columns per K = 16
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_bf16_ts(d1, at + k * 8, bd + 2 * k, idesc256, 1u);
                        tc_commit_mc(&empty[s0], kAllMask); tc_commit_mc(&empty[s1], kAllMask);
                    }
                    __syncwarp();
                }
                CH_TRACE(0, it * 100 + 40 + j);
            };
            mbar_wait_inl(e1_done, it & 1);
            tc_fence_after();
            CH_TRACE(0, it * 100 + 3);
            f1(0); f1(1);
            mbar_wait_inl(d1_done, it & 1);   // FFN2 accumulates onto y + b2
            f2(0); f1(2); f2(1); f1(3); f2(2); f2(3);
            if (leader) tc_commit(f2_full);
        }
      } else if (warp == 2) {
        // ===================================================== slot releaser: once the 16 epilogue warps have read the residual boxes
        // of a tile, their ring slots are free again -- in this CTA and in the peer (whose multicasts land in them, too)
        constexpr int kBoxes = FFN ? 48 : 16;
        const uint32_t empty_peer0 = mapa_u32(smem_u32(&empty[0]), rank ^ 1u);
        for (int it = 0; it < my_iters; ++it) {
            mbar_wait_inl(resid_read, it & 1);
            if (lane < 4) {
                const uint32_t s_r = (uint32_t)(it * kBoxes + 12 + lane) % kC3Stages;
                mbar_arrive(&empty[s_r]);
                mbar_arrive_cluster(empty_peer0 + s_r * 8);
            }
            __syncwarp();
        }
      } else {
        // ===================================================== output storer: Y (the staged bf16 rows) -> global, four [128 x 64] boxes
        for (int it = 0; it < my_iters; ++it) {
            const int row0 = ((cid + it * ncl) * kC3Cluster + (int)rank) * 128;
            mbar_wait_inl(staged, it & 1);
            if (lane == 0) {
#pragma unroll
                for (int kb = 0; kb < 4; ++kb) tma_store_2d(&tmap_out, sY + kb * kC3StageBytes, kb * 64, row0);
                tma_store_commit();
                tma_store_wait_read<0>();
                mbar_arrive(y_free);
            }
            __syncwarp();
        }
        if (lane == 0) tma_store_wait_all<0>();  // smem must outlive the last stores
      }
    } else {
        // ===================================================== epilogue warps (4..19)
        reg_inc<112>();
        const int e = warp - 4;
        const int q = e & 3;               // TMEM lane quarter (== warp & 3)
        const int cq = e >> 2;             // column quarter: columns [64 cq, 64 cq + 64) of a 256-wide row
        const int row = q * 32 + lane;
        const int col0 = cq * 64;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int quad_bar = 1 + q;        // named barrier of the four warps that share a lane quarter (128 threads)
        const int sw = row & 7;            // 128B-swizzle phase of this row
        uint8_t* yrow = sY + cq * kC3StageBytes + row * 128;  // this thread's 64 columns = row `row` of box cq of Y
        // statistics exchange of this lane quarter: [4 column quarters][32 lanes] (mean, M2) = the first 1 KB of the Y rows
        // that warp (q, cq = 0) overwrites next (always after the second quarter barrier below)
        float2* s_stat = reinterpret_cast<float2*>(sY + q * 4096);
        int tr_n = (warp == 4 && lane == 0) ? 0 : kChTraceSlots;

        // combines this thread's (mean, M2) over its 64 columns with the three other quarters of the row
        auto quad_stats = [&](float mean_a, float m2_a, float& mean, float& rstd) {
            s_stat[cq * 32 + lane] = make_float2(mean_a, m2_a);
            asm volatile("bar.sync %0, 128;" ::"r"(quad_bar) : "memory");
            const float2 a0 = s_stat[lane], a1 = s_stat[32 + lane], a2 = s_stat[64 + lane], a3 = s_stat[96 + lane];
            asm volatile("bar.sync %0, 128;" ::"r"(quad_bar) : "memory");  // all reads done before the rows are overwritten
            mean = 0.25f * ((a0.x + a1.x) + (a2.x + a3.x));
            const float d0 = a0.x - mean, d1 = a1.x - mean, d2 = a2.x - mean, d3 = a3.x - mean;
            const float m2 = ((a0.y + a1.y) + (a2.y + a3.y)) + 64.f * ((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3));
            rstd = rsqrtf(fmaxf(m2 * (1.f / 256.f), 0.f) + 1e-5f);
        };
        // sum and sum of squares of 32 values as packed pairs.  (No pivot shift: the rows are LayerNorm inputs -- a
        // normalised row plus a bounded update -- so |mean| is of the order of the standard deviation and the one-sweep
        // M2 = s2 - s1 * mean loses nothing in fp32; the quarters are then combined by their own means.)
        auto accum_stats = [&](const float* w, float pivot, float2& s1, float2& s2) {
            const float2 np = make_float2(-pivot, -pivot);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float2 d = f2add(make_float2(w[2 * i], w[2 * i + 1]), np);
                s1 = f2add(s1, d);
                s2 = f2fma(d, d, s2);
            }
        };
        constexpr int kBoxesPerTile = FFN ? 48 : 16;   // ring positions per tile: 12 G1 operands, 4 residual boxes, 32 FFN weights
        // y = ((v - mean) * rstd) * gamma + beta for 4 columns starting at column c: v <- y, and the packed bf16 pairs
        auto norm4 = [&](float* w, int c, float2 rs2, float2 nmr2, uint32_t& p0, uint32_t& p1) {
            const float4 ga = *reinterpret_cast<const float4*>(s_gamma + c);
            const float4 be = *reinterpret_cast<const float4*>(s_beta + c);
            const float2 n0 = f2fma(make_float2(w[0], w[1]), rs2, nmr2), n1 = f2fma(make_float2(w[2], w[3]), rs2, nmr2);
            const float2 y0 = f2fma(n0, make_float2(ga.x, ga.y), make_float2(be.x, be.y));
            const float2 y1 = f2fma(n1, make_float2(ga.z, ga.w), make_float2(be.z, be.w));
            p0 = pack_bf16x2(y0.x, y0.y);
            p1 = pack_bf16x2(y1.x, y1.y);
            w[0] = y0.x; w[1] = y0.y; w[2] = y1.x; w[3] = y1.y;
        };
        // normalises this thread's 64 columns (v <- y) and writes them, bf16, into its row of Y
        auto norm_row_to_y = [&](float* v, float mean, float rstd) {
            const float2 rs2 = make_float2(rstd, rstd), nmr2 = make_float2(-mean * rstd, -mean * rstd);
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                uint4 pk;
                norm4(v + 8 * g, col0 + 8 * g, rs2, nmr2, pk.x, pk.y);
                norm4(v + 8 * g + 4, col0 + 8 * g + 4, rs2, nmr2, pk.z, pk.w);
                *reinterpret_cast<uint4*>(yrow + ((g ^ sw) << 4)) = pk;
            }
        };
        // this warp's rows are staged in Y (visible to the async proxy): warp 3 stores the tile once all 16 warps are here
        auto stage_done = [&]() {
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(staged);
        };

        for (int it = 0; it < my_iters; ++it) {
            const int par = it & 1;
            const uint32_t d1 = tmem_base + par * 256 + lane_off + col0;
            const uint32_t a2 = tmem_base + (par ^ 1) * 256 + lane_off;
            const int qd = par * 2, qa = (par ^ 1) * 2;
            (void)qa; (void)a2;

            // ---------------- E1: pre = acc + bo + x (kept in registers) ; y = LN(pre) ; D1 <- y + b2 ; Y <- bf16(y)
            CH_TRACE(1, it * 100);
            const uint32_t c_r = (uint32_t)(it * kBoxesPerTile + 12 + cq);   // ring position of residual box cq of this tile
            const uint32_t s_r = c_r % kC3Stages;
            const uint8_t* rrow = sRing + s_r * kC3StageBytes + row * 128;
            // g1_full FIRST: a parity wait may only target phase n once phase n - 1 is known complete, and the slot's previous
            // occupant (ring position c_r - 10, a G1 operand of this very tile) has certainly landed once G1 has completed
            mbar_wait_inl(&g1_full[par], (it >> 1) & 1);
            mbar_wait_inl(&full[s_r], (c_r / kC3Stages) & 1);
            __syncwarp();
            tc_fence_after();
            CH_TRACE(1, it * 100 + 2);
            float v[64];
            float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
            float pivot = 0.f;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                tmem_ld32(d1 + h * 32, v + h * 32);
                tc_wait_ld();
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const uint4 u = *reinterpret_cast<const uint4*>(rrow + (((h * 4 + g) ^ sw) << 4));
                    const float4 ba = __ldg(reinterpret_cast<const float4*>(p.bo + col0 + h * 32 + 8 * g));
                    const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bo + col0 + h * 32 + 8 * g + 4));
                    float* w = v + h * 32 + 8 * g;
                    const float2 t0 = f2add(f2add(make_float2(w[0], w[1]), make_float2(ba.x, ba.y)), bf16x2_to_f2(u.x));
                    const float2 t1 = f2add(f2add(make_float2(w[2], w[3]), make_float2(ba.z, ba.w)), bf16x2_to_f2(u.y));
                    const float2 t2 = f2add(f2add(make_float2(w[4], w[5]), make_float2(bb.x, bb.y)), bf16x2_to_f2(u.z));
                    const float2 t3 = f2add(f2add(make_float2(w[6], w[7]), make_float2(bb.z, bb.w)), bf16x2_to_f2(u.w));
                    w[0] = t0.x; w[1] = t0.y; w[2] = t1.x; w[3] = t1.y; w[4] = t2.x; w[5] = t2.y; w[6] = t3.x; w[7] = t3.y;
                }
                if (h == 0) pivot = v[0];
                accum_stats(v + h * 32, pivot, s1, s2);
            }
            if constexpr (!FFN) {  // the accumulator has been read: G1 of the tile after next may overwrite the region
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&qfree[qd]);
                    mbar_arrive(&qfree[qd + 1]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(resid_read);   // this warp is done with the residual boxes (warp 2 frees the slots)
            CH_TRACE(1, it * 100 + 3);
            // Y (statistics exchange, then y / the staged output rows) is about to be overwritten: the previous tile's TMA
            // store must have finished reading it
            if (it > 0) mbar_wait_inl(y_free, (it - 1) & 1);
            float mean, rstd;
            {
                const float t1 = s1.x + s1.y, t2 = s2.x + s2.y;
                const float m1 = t1 * (1.f / 64.f);
                quad_stats(pivot + m1, fmaxf(t2 - t1 * m1, 0.f), mean, rstd);
            }
            CH_TRACE(1, it * 100 + 4);
            norm_row_to_y(v, mean, rstd);
            if constexpr (!FFN) {
                stage_done();
                CH_TRACE(1, it * 100 + 42);
            } else {
            // y is in smem: FFN1 can start; the FFN2 accumulator initialisation y + b2 follows off the critical path
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(e1_done);
            CH_TRACE(1, it * 100 + 5);
#pragma unroll
            for (int g = 0; g < 16; ++g) {
                const float4 bz = __ldg(reinterpret_cast<const float4*>(p.b2 + col0 + 4 * g));
                const float2 t0 = f2add(make_float2(v[4 * g], v[4 * g + 1]), make_float2(bz.x, bz.y));
                const float2 t1 = f2add(make_float2(v[4 * g + 2], v[4 * g + 3]), make_float2(bz.z, bz.w));
                v[4 * g] = t0.x; v[4 * g + 1] = t0.y; v[4 * g + 2] = t1.x; v[4 * g + 3] = t1.y;
            }
            tmem_st32(d1, v);
            tmem_st32(d1 + 32, v + 32);
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(d1_done);
            CH_TRACE(1, it * 100 + 6);

            // ---------------- E2(j): h_j = relu(acc2 + b1) -> bf16, written back over the first 64 columns of its accumulator
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                const int hb = j & 1;
                mbar_wait_inl(&f1_full[hb], (j >> 1) & 1);   // two chunks per half and tile: the phase parity is the chunk's turn
                __syncwarp();
                tc_fence_after();
                CH_TRACE(1, it * 100 + 20 + j);
                tmem_ld32(a2 + hb * 128 + cq * 32, v);
                tc_wait_ld();
                const int hcol = j * 128 + cq * 32;
                uint32_t pk[16];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const float4 ba = __ldg(reinterpret_cast<const float4*>(p.b1 + hcol + 8 * g));
                    const float4 bb = __ldg(reinterpret_cast<const float4*>(p.b1 + hcol + 8 * g + 4));
                    const float2 t0 = f2add(make_float2(v[8 * g + 0], v[8 * g + 1]), make_float2(ba.x, ba.y));
                    const float2 t1 = f2add(make_float2(v[8 * g + 2], v[8 * g + 3]), make_float2(ba.z, ba.w));
                    const float2 t2 = f2add(make_float2(v[8 * g + 4], v[8 * g + 5]), make_float2(bb.x, bb.y));
                    const float2 t3 = f2add(make_float2(v[8 * g + 6], v[8 * g + 7]), make_float2(bb.z, bb.w));
                    pk[4 * g + 0] = pack_bf16x2_relu(t0.x, t0.y);
                    pk[4 * g + 1] = pack_bf16x2_relu(t1.x, t1.y);
                    pk[4 * g + 2] = pack_bf16x2_relu(t2.x, t2.y);
                    pk[4 * g + 3] = pack_bf16x2_relu(t3.x, t3.y);
                }
                // the packed columns [16 cq, 16 cq + 16) overlap fp32 columns that another column quarter of this lane
                // quarter may still be reading: all four warps have their values in registers first
                asm volatile("bar.sync %0, 128;" ::"r"(quad_bar) : "memory");
                tmem_st16(a2 + hb * 128 + cq * 16, pk);
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&qfree[qa + hb]);
                    mbar_arrive(&h_full[hb]);
                }
                CH_TRACE(1, it * 100 + 30 + j);
            }

            // ---------------- LN2: statistics over D1 (every MMA of the tile has completed: Y is dead, too)
            mbar_wait_inl(f2_full, it & 1);
            __syncwarp();
            tc_fence_after();
            CH_TRACE(1, it * 100 + 40);
            s1 = make_float2(0.f, 0.f); s2 = make_float2(0.f, 0.f);
            tmem_ld32(d1, v);
            tmem_ld32(d1 + 32, v + 32);
            tc_wait_ld();
            pivot = v[0];
            accum_stats(v, pivot, s1, s2);
            accum_stats(v + 32, pivot, s1, s2);
            tc_fence_before();   // D1 is in registers: G1 of the tile after next may overwrite the region
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&qfree[qd]);
                mbar_arrive(&qfree[qd + 1]);
            }
            {
                const float t1 = s1.x + s1.y, t2 = s2.x + s2.y;
                const float m1 = t1 * (1.f / 64.f);
                quad_stats(pivot + m1, fmaxf(t2 - t1 * m1, 0.f), mean, rstd);
            }
            CH_TRACE(1, it * 100 + 41);
            norm_row_to_y(v, mean, rstd);
            stage_done();
            CH_TRACE(1, it * 100 + 42);
            }  // FFN
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // no CTA leaves while a peer may still multicast into its ring or signal its barriers
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace etude
