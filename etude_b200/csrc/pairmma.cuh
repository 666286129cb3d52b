// cta_group::2 ("CTA pair") tcgen05 primitives: two CTAs of a cluster execute ONE MMA with M = 256 (128 accumulator rows in
// each CTA's TMEM).  A comes from each CTA's own smem (or TMEM) at the same offset; B is split along N: the CTA of rank r
// supplies rows [r N/2, (r + 1) N/2) from ITS smem at the same offset, and the hardware reads both halves for both CTAs.
// Only the leader CTA (rank 0) issues; completion is multicast to mbarriers at the same offset in both CTAs.  Every
// tcgen05 alloc / mma / commit of a kernel must use the same cta_group, so a kernel uses either these or common.cuh's.
//
// The dev build adds a self-test (pairmma_test_kernel) that pins down what attn_pair.cuh relies on: the N-split of B in
// an SS MMA, an A operand in TMEM (bf16 pairs written by tcgen05.st in both CTAs) in a TS MMA, and the commit multicast.
#pragma once
#include "cluster.cuh"
#include "common.cuh"

namespace etude {

// Issued by one warp (same warp index) in EACH CTA of the pair, same smem offset for the result.
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
// arrives (once) on this CTA's mbarrier when every pair MMA issued so far by this thread has completed
__device__ __forceinline__ void tc_commit2(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// ... on the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void tc_commit2_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
// D[tmem of both CTAs] (+)= A[smem desc, each CTA's own 128 rows] * B[smem desc, N/2 rows from each CTA]
__device__ __forceinline__ void umma2_bf16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// ... with A in TMEM (bf16 pairs packed per 32-bit column, each CTA's own 128 lanes)
__device__ __forceinline__ void umma2_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// TMA load into THIS CTA's smem whose completion bytes go to an mbarrier that may live in the peer CTA of the pair
// (bar_cluster_addr from mapa_u32): the leader's barrier collects both CTAs' halves of an operand.
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
// byte offset of element (row, col) of a K-major [rows x 64] bf16 tile in the 128-byte-swizzle layout (rows of 128 B)
__device__ __forceinline__ uint32_t sw128_off(int row, int col) {
    return (uint32_t)(row * 128 + ((((col >> 3) ^ (row & 7)) << 4) | ((col & 7) << 1)));
}

#ifdef ETUDE_DEV_BUILD
// Self-test: D[256 x 128] = A[256 x 64] B[128 x 64]^T (SS pair MMA, B split 64 + 64 rows), P = bf16(D) written back into
// TMEM over D, O[256 x 64] = P[256 x 128] VT[64 x 128]^T (TS pair MMA, VT split 32 + 32 rows, K = 128 as two 64-key chunks).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pairmma_test_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B, const __nv_bfloat16* __restrict__ VT,
                    float* __restrict__ D, float* __restrict__ O) {
    extern __shared__ __align__(1024) uint8_t pm_smem[];
    uint8_t* sA = pm_smem;                  // [128 x 64] 16 KB
    uint8_t* sB = sA + 16384;               // [64 x 64]   8 KB
    uint8_t* sVT = sB + 8192;               // 2 chunks x [32 x 64] 4 KB
    uint64_t* done = reinterpret_cast<uint64_t*>(sVT + 8192);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done + 2);
    const uint32_t rank = cluster_ctarank();
    const int tid = threadIdx.x, warp = tid >> 5;
    if ((smem_u32(pm_smem) & 1023u) != 0) __trap();
    if (tid == 0) {
        mbar_init(&done[0], 1);
        mbar_init(&done[1], 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc2(tmem_ptr, 512);
    auto put = [](uint8_t* tile, int row, int col, __nv_bfloat16 v) { *reinterpret_cast<__nv_bfloat16*>(tile + sw128_off(row, col)) = v; };
    for (int i = tid; i < 128 * 64; i += 128) put(sA, i >> 6, i & 63, A[((int)rank * 128 + (i >> 6)) * 64 + (i & 63)]);
    for (int i = tid; i < 64 * 64; i += 128) put(sB, i >> 6, i & 63, B[((int)rank * 64 + (i >> 6)) * 64 + (i & 63)]);
    for (int i = tid; i < 32 * 128; i += 128) {
        const int d = i >> 7, key = i & 127;
        put(sVT + (key >> 6) * 4096, d, key & 63, VT[((int)rank * 32 + d) * 128 + key]);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    if (rank == 0 && tid == 0) {
        constexpr uint32_t idesc = make_idesc_bf16(256, 128, 0, 0);
        const uint64_t ad = make_sw128_desc(smem_u32(sA)), bd = make_sw128_desc(smem_u32(sB));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma2_bf16_ss(tmem_base, ad + 2 * k, bd + 2 * k, idesc, k != 0);
        tc_commit2_mc(&done[0], 3);
    }
    mbar_wait(&done[0], 0);
    tc_fence_after();
    const int row = tid;   // TMEM lane
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    float v[32];
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
        tmem_ld32(tmem_base + lane_off + c * 32, v);
        tc_wait_ld();
        uint32_t pk[16];
        for (int i = 0; i < 32; ++i) D[(size_t)((int)rank * 128 + row) * 128 + c * 32 + i] = v[i];
        for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
        tmem_st16(tmem_base + lane_off + c * 16, pk);   // P chunk c -> columns [16 c, 16 c + 16): below the D columns still to be read
    }
    tc_wait_st();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    if (rank == 0 && tid == 0) {
        constexpr uint32_t idesc = make_idesc_bf16(256, 64, 0, 0);
        const uint64_t vd = make_sw128_desc(smem_u32(sVT));
#pragma unroll
        for (int s = 0; s < 8; ++s)
            umma2_bf16_ts(tmem_base + 128, tmem_base + s * 8, vd + (uint64_t)((s >> 2) * (4096 >> 4) + (s & 3) * 2), idesc, s != 0);
        tc_commit2_mc(&done[1], 3);
    }
    mbar_wait(&done[1], 0);
    tc_fence_after();
    for (int c = 0; c < 2; ++c) {
        tmem_ld32(tmem_base + 128 + lane_off + c * 32, v);
        tc_wait_ld();
        for (int i = 0; i < 32; ++i) O[(size_t)((int)rank * 128 + row) * 64 + c * 32 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc2(tmem_base, 512);
}

// Rate of cta_group::2 MMAs (M = 256 over the pair, N = n, K = 16): `iters` MMAs issued by the leader CTA's elected thread from
// uniform code, A in smem (ts = 0) or in TMEM (ts = 1), B = n / 2 rows per CTA.  out[0] = issue clocks, out[1] = clocks until
// completion (cluster 0).  Compare with mmabench.cuh's cta_group::1 numbers: with both operands in smem those run at
// 60 % (N = 128) / 75 % (N = 256) of the tensor floor -- bound by the ~75 B/clk of operand reads per SM; a pair MMA reads only
// half of B from each SM's smem.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) pairmma_bench_kernel(int ts, int n, int iters, int alt, long long* out) {
    extern __shared__ __align__(1024) uint8_t pb_smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const uint32_t rank = cluster_ctarank();
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc2(&tmem_ptr, 512);
    for (int i = threadIdx.x; i < 192 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(pb_smem)[i] = make_uint4(0, 0, 0, 0);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = tmem_ptr;
    if (warp == 1 && rank == 0) {
        const uint32_t idesc = make_idesc_bf16(256, n, 0, 0);
        const uint64_t a_desc0 = make_sw128_desc(smem_u32(pb_smem));
        const uint64_t b_desc0 = make_sw128_desc(smem_u32(pb_smem) + 64 * 1024);
        const bool leader = elect_one();
        const long long t0 = clock64();
        for (int i = 0; i < iters; i += 4) {
            const uint32_t buf = (uint32_t)(i >> 2) & 3u;
            const uint64_t ad = a_desc0 + (uint64_t)(buf * 1024u), bd = b_desc0 + (uint64_t)(buf * 1024u);   // 16 KB apart
            if (leader) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    // alt = 1: consecutive MMAs alternate between two accumulators; 2: ... and share the A k-slice (two N-chunks
                    // of one GEMM issued k-step by k-step)
                    const uint32_t d = tmem + 256 + (alt ? (uint32_t)(k & 1) * 128u : 0u);
                    const uint64_t a_k = ad + 2 * (alt == 2 ? (k >> 1) : k);
                    if (ts) umma2_bf16_ts(d, tmem + k * 8, bd + 2 * k, idesc, 1u);
                    else umma2_bf16_ss(d, a_k, bd + 2 * k, idesc, 1u);
                }
            }
            __syncwarp();
        }
        if (leader) tc_commit2(&bar);
        const long long t1 = clock64();
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0 && leader) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc2(tmem, 512);
}
#endif  // ETUDE_DEV_BUILD

}  // namespace etude
