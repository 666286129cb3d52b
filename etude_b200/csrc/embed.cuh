// Token embedding of the frequency encoder:  unfold(65) -> Conv2d(1,4,(1,5)) -> Linear(244,256) -> *16 + pos[bin]
// (reference amt_apc.py:79-109).  There is no non-linearity between the conv and the linear layer, so the two
// are folded at load time into one 65-tap, 1 -> 256 channel filter along time per mel bin:
//     x[(w,f,b), h] = sum_t W16[h][t] * feat_w[f + t][b] + posb[b][h]
// with W16 = 16 * W_eff and posb = 16 * b_eff + pos_embedding_freq.  fp32 on CUDA cores: 0.4 % of the path's FLOPs.
#pragma once
#include "common.cuh"

namespace etude {

constexpr int kEmbedFrames = 4;   // frames per CTA
constexpr int kEmbedRows = kEmbedFrames + kProc - 1;  // 68 feature rows staged in smem
constexpr int kEmbedThreads = 256;                    // one thread per hidden channel

__global__ void __launch_bounds__(kEmbedThreads)
embed_kernel(const float* __restrict__ feat, const int64_t* __restrict__ win_row0, const float* __restrict__ w16 /*[256][65]*/,
             const float* __restrict__ posb /*[256 bins][256]*/, __nv_bfloat16* __restrict__ out_bf16) {
    extern __shared__ float s_feat[];  // [68][256]
    const int w = blockIdx.y;
    const int f0 = blockIdx.x * kEmbedFrames;
    const int h = threadIdx.x;
    const float* src = feat + (win_row0[w] + f0) * kBins;  // window w covers padded rows [row0, row0 + 576)
    for (int i = threadIdx.x; i < kEmbedRows * kBins / 4; i += kEmbedThreads)
        reinterpret_cast<float4*>(s_feat)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
    float wr[kProc];
#pragma unroll
    for (int t = 0; t < kProc; ++t) wr[t] = __ldg(&w16[h * kProc + t]);
    __syncthreads();

    const size_t tok0 = ((size_t)w * kFrames + f0) * kBins;
#pragma unroll 1
    for (int b = 0; b < kBins; b += 4) {
        float acc[kEmbedFrames][4];
#pragma unroll
        for (int f = 0; f < kEmbedFrames; ++f)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[f][j] = 0.f;
#pragma unroll
        for (int r = 0; r < kEmbedRows; ++r) {
            const float4 x = *reinterpret_cast<const float4*>(&s_feat[r * kBins + b]);  // broadcast read
#pragma unroll
            for (int f = 0; f < kEmbedFrames; ++f) {
                const int t = r - f;
                if (t >= 0 && t < kProc) {
                    acc[f][0] = fmaf(wr[t], x.x, acc[f][0]);
                    acc[f][1] = fmaf(wr[t], x.y, acc[f][1]);
                    acc[f][2] = fmaf(wr[t], x.z, acc[f][2]);
                    acc[f][3] = fmaf(wr[t], x.w, acc[f][3]);
                }
            }
        }
#pragma unroll
        for (int f = 0; f < kEmbedFrames; ++f)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float y = acc[f][j] + __ldg(&posb[(b + j) * kHid + h]);
                const size_t idx = (tok0 + (size_t)f * kBins + b + j) * kHid + h;
                out_bf16[idx] = __float2bfloat16(y);
            }
    }
}

}  // namespace etude
