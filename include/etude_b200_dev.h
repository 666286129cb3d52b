/* etude_b200_dev.h -- test-only entry points of libetude_b200_dev.so (the product source compiled with
 * -DETUDE_DEV_BUILD): tcgen05 / TMEM micro-benchmarks, clock64 kernel timelines and the generic tile GEMM epilogues.
 * None of this is in libetude_b200.so; an integrator never needs it.  The dev library also exports everything
 * etude_b200.h and etude_b200_kernels.h declare (same code), so the kernel tests can run entirely against it. */
#ifndef ETUDE_B200_DEV_H
#define ETUDE_B200_DEV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Debug: tcgen05.mma rate micro-benchmark (mmabench.cuh): `iters` M128 x N x K16 bf16 MMAs from one thread per CTA,
 * mode 0 = both operands in smem, 1 = A in TMEM; host_out[0] = issue clocks, host_out[1] = clocks until completion. */
int etude_debug_mma_bench(int mode, int n, int iters, int n_bufs, int grid, int64_t* host_out);
/* TMEM read / MUFU / pack micro-benchmark (mmabench.cuh): clock span of `iters` loop bodies with n_warps warps per CTA. */
int etude_debug_tmem_bench(int mode, int n_warps, int iters, int grid, int64_t* host_out);
/* tcgen05.mma rate under concurrent tcgen05.ld/st traffic from n_ld other warps (mmabench.cuh). host_out: {clk, ld iterations}. */
int etude_debug_mma_mix(int ts, int iters, int n_ld, int st_too, int grid, int64_t* host_out);

/* Debug: clock64 timeline of CTA 0 of the next etude_k_chain launches.  enable != 0 allocates / clears the device
 * buffer, 0 frees it; host_out (optional) first receives the current buffer: 3 roles (MMA thread, one epilogue
 * thread, ring producer) x 512 (event id, clock) int64 pairs. */
int etude_debug_chain_trace(int enable, int64_t* host_out, int n_values);

/* etude_k_attn_qkv's operator (fused Q|K|V projection + self-attention of 256-token sequences) through the cta_group::1
 * kernel of attn_qkv.cuh: an independent implementation that the tests hold against the product's CTA-pair kernel. */
int etude_debug_attn_qkv_cta1(const void* x_dev, const void* w_hm_dev, const float* bias_hm_dev, int n_seq, void* out_dev, void* stream);
/* cta_group::2 (CTA pair) tcgen05 self-test (pairmma.cuh): a bf16 [256, 64], b bf16 [128, 64], vt bf16 [64, 128] on the device ->
 * d fp32 [256, 128] = a b^T (SS pair MMA, B split 64 + 64 rows over the two CTAs), o fp32 [256, 64] = bf16(d) vt^T (TS pair MMA,
 * A in TMEM, vt split 32 + 32 rows).  Synchronous. */
int etude_debug_pairmma(const void* a, const void* b, const void* vt, float* d, float* o);
/* Rate of cta_group::2 MMAs (M = 256 over a CTA pair, N = n, K = 16; ts: A in TMEM): host_out = {issue clocks, clocks to completion}. */
int etude_debug_pairmma_bench(int ts, int n, int iters, int alt, int grid, int64_t* host_out);   /* alt: 1 alternate two accumulators, 2 also share the A slice */

#ifdef __cplusplus
}
#endif
#endif
