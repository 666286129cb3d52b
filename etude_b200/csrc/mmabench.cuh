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
            // modes 4 / 5: B is MN-major (the P V product of the attention kernels: V tile [keys][64 d], d contiguous)
            const bool mn = mode >= 4;
            const uint32_t idesc = make_idesc_bf16(128, n, 0, mn ? 1 : 0);
            const uint64_t a_desc0 = make_sw128_desc(smem_u32(smem));
            const uint64_t b_desc0 = mn ? make_sw128_desc(smem_u32(smem) + 64 * 1024, 8192) : make_sw128_desc(smem_u32(smem) + 64 * 1024);
            const uint32_t bstep = mn ? 128u : 2u;
            const bool leader = elect_one();
            const long long t0 = clock64();
            for (int i = 0; i < iters; i += 4) {
                const uint32_t buf = (uint32_t)(i >> 2) & 3u;
                const uint64_t ad = a_desc0 + (uint64_t)(buf * 1024u), bd = b_desc0 + (uint64_t)(buf * 2048u);
                if (leader) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (mode == 2 || mode == 5) umma_bf16_ss(tmem + 256, ad + 2 * k, bd + bstep * k, idesc, 1u);
                        else umma_bf16_ts(tmem + 256, tmem + k * 8, bd + bstep * k, idesc, 1u);
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


// Does TMEM load/store traffic from other warps slow tcgen05.mma down?  Warp 1 issues `iters` MMAs (ts = 1: A from TMEM,
// M128 N64 K16, the P V shape; ts = 0: both operands from smem, M128 N128 K16, the Q K^T shape) while `n_ld` warps
// (warps 4...) loop tcgen05.ld x16 + tcgen05.st x8 on other TMEM columns until the MMAs are done.  Reports clk per MMA.
__global__ void __launch_bounds__(768, 1) mma_mix_kernel(int ts, int iters, int n_ld, int st_too, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    __shared__ volatile int done;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); done = 0; }
    if (warp == 0) tmem_alloc(&tmem_ptr, 512);
    for (int i = threadIdx.x; i < 128 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_ptr;
    if (warp == 1) {
        const int n = ts ? 64 : 128;
        const uint32_t idesc = make_idesc_bf16(128, n, 0, ts ? 1 : 0);
        const uint64_t a_desc0 = make_sw128_desc(smem_u32(smem));
        const uint64_t b_desc0 = ts ? make_sw128_desc(smem_u32(smem) + 64 * 1024, 8192) : make_sw128_desc(smem_u32(smem) + 64 * 1024);
        const uint32_t bstep = ts ? 128u : 2u;
        const bool leader = elect_one();
        const long long t0 = clock64();
        for (int i = 0; i < iters; i += 4) {
            if (leader) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (ts) umma_bf16_ts(tmem + 64, tmem + k * 8, b_desc0 + bstep * k, idesc, 1u);
                    else umma_bf16_ss(tmem, a_desc0 + 2 * k, b_desc0 + bstep * k, idesc, 1u);
                }
            }
            __syncwarp();
        }
        if (leader) tc_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0 && leader) out[0] = t2 - t0;
        done = 1;
    } else if (warp >= 4 && warp < 4 + n_ld) {
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256 + (uint32_t)((warp >> 2) & 3) * 32;
        float v[16];
        uint32_t pk[8];
        float acc = 0.f;
        long long n_iter = 0;
        while (!done) {
            tmem_ld16(taddr, v);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 8; ++i) pk[i] = __float_as_uint(v[2 * i] + v[2 * i + 1]);
            if (st_too) {
                tmem_st8(taddr + 16, pk);
                tc_wait_st();
            } else {
                acc += __uint_as_float(pk[0]);
            }
            ++n_iter;
        }
        if (acc == 1.2345f) out[2] = 1;
        if (blockIdx.x == 0 && warp == 4 && (threadIdx.x & 31) == 0) out[1] = n_iter;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// Micro-benchmark of the softmax-side resources (debug entry etude_debug_tmem_bench): `blockDim.x / 32` warps per CTA, one
// CTA per SM, every warp loops `iters` times over
//   mode 0: tcgen05.ld 32x32b.x32 + wait          (TMEM read bandwidth: 4 KB per warp-instruction)
//   mode 1: 32 x ex2.approx per lane              (MUFU rate)
//   mode 2: 16 x cvt.rn.bf16x2.f32 per lane       (pack rate)
//   mode 3: ld x32 + wait, 32 x (ffma, ex2, fadd), 16 packs, tcgen05.st x16 + wait   (the softmax pass-2 body)
//   mode 4: 32 x 3-input fmax on loaded registers (max pass ALU work, no TMEM)
// and warp 0 of CTA 0 reports its clock64 span.
__global__ void __launch_bounds__(512, 1) tmem_bench_kernel(int mode, int iters, long long* out, float* sink) {
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(&tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t taddr = tmem_ptr + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) & 3) * 64;
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = (float)(threadIdx.x + i) * 1e-3f;
    float acc = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    if (mode == 0) {
        for (int it = 0; it < iters; ++it) {
            tmem_ld32(taddr, v);
            tc_wait_ld();
            acc += v[0] + v[31];  // consume: the wait is a scoreboard on the destination registers
        }
    } else if (mode == 1) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                float y;
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v[i]));
                v[i] = y;
            }
        }
    } else if (mode == 2) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                uint32_t pk;
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(v[2 * i + 1]), "f"(v[2 * i]));
                v[2 * i] = __uint_as_float(pk);
            }
        }
    } else if (mode == 3) {
        for (int it = 0; it < iters; ++it) {
            tmem_ld32(taddr, v);
            tc_wait_ld();
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                float e0, e1;
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmaf(v[i], 0.18f, -1.f)));
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(v[i + 1], 0.18f, -1.f)));
                acc += e0 + e1;
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[i >> 1]) : "f"(e1), "f"(e0));
            }
            tmem_st16(taddr, pk);
            tc_wait_st();
        }
    } else if (mode == 5) {  // tcgen05.st x16 (2 KB) + wait
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = __float_as_uint(v[i]);
        for (int it = 0; it < iters; ++it) {
            tmem_st16(taddr, pk);
            tc_wait_st();
        }
    } else if (mode == 6) {  // two tcgen05.ld x32 in flight before the wait (8 KB)
        float w[32];
        for (int it = 0; it < iters; ++it) {
            tmem_ld32(taddr, v);
            tmem_ld32(taddr + 32, w);
            tc_wait_ld();
            acc += v[0] + w[31];
        }
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                float d;
                asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(acc), "f"(v[i]), "f"(v[i + 1]));
                acc = d;
                v[i] += 1.f;
            }
        }
    }
    const long long t1 = clock64();
#pragma unroll
    for (int i = 0; i < 32; ++i) acc += v[i];
    if (acc == 123.456f) sink[0] = acc;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_ptr, 512);
}

}  // namespace etude
