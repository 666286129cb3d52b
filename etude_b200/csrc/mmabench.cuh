// Micro-benchmark of tcgen05.mma issue/execute rates (debug entry etude_debug_mma_bench): one CTA per SM issues `iters`
// back-to-back MMAs of one shape from one thread against resident (uninitialised) smem / TMEM operands and reports the
// clock64 span until the final commit lands.  mode 0: SS, mode 1: TS (A in TMEM).
#pragma once
#include "common.cuh"

namespace etude {

__global__ void __launch_bounds__(128, 1) mma_bench_kernel(int mode, int n, int iters, int n_bufs, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tmem_ptr, 512);
    // zero the operands (denormal / NaN patterns could change power, not timing; keep it clean anyway)
    for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_ptr;
    if (mode >= 2) {
        // warp-uniform issue loop: every lane runs the loop (addresses stay in uniform registers), one elected lane issues
        const int uwarp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
        if (uwarp == 1) {
            const uint32_t idesc = make_idesc_bf16(128, n, 0, 0);
            const uint64_t a_desc0 = make_sw128_desc(smem_u32(smem));
            const uint64_t b_desc0 = make_sw128_desc(smem_u32(smem) + 64 * 1024);
            const bool leader = elect_one();
            const long long t0 = clock64();
            for (int i = 0; i < iters; i += 4) {
                const uint32_t buf = (uint32_t)(i >> 2) & 3u;
                const uint64_t ad = a_desc0 + (uint64_t)(buf * 1024u), bd = b_desc0 + (uint64_t)(buf * 2048u);
                if (leader) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (mode == 2) umma_bf16_ss(tmem + 256, ad + 2 * k, bd + 2 * k, idesc, 1u);
                        else umma_bf16_ts(tmem + 256, tmem + k * 8, bd + 2 * k, idesc, 1u);
                    }
                }
                __syncwarp();
            }
            if (leader) tc_commit(&bar);
            const long long t1 = clock64();
            mbar_wait(&bar, 0);
            const long long t2 = clock64();
            if (blockIdx.x == 0 && leader) { out[0] = t1 - t0; out[1] = t2 - t0; }
        }
    } else if (warp == 1 && lane == 0) {
        const uint32_t idesc = make_idesc_bf16(128, n, 0, 0);
        const uint32_t a_base = smem_u32(smem);                 // A: [128 x 64] boxes, 16 KB each
        const uint32_t b_base = smem_u32(smem) + 64 * 1024;     // B: [n x 64] boxes, up to 32 KB each
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const int buf = i % n_bufs, k = i & 3;
            const uint32_t a_addr = a_base + (buf & 3) * 16384 + k * 32;
            const uint32_t b_addr = b_base + (buf % 4) * 32768 + k * 32;
            if (mode == 0) umma_bf16_ss(tmem + 256, make_sw128_desc(a_addr), make_sw128_desc(b_addr), idesc, 1u);
            else umma_bf16_ts(tmem + 256, tmem + (i & 7) * 8, make_sw128_desc(b_addr), idesc, 1u);
        }
        tc_commit(&bar);
        const long long t1 = clock64();
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace etude
