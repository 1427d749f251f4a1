// First layer of the Wav2Vec2 convolutional feature encoder in channels-last layout, the part of the audio front-end
// that is not a GEMM: Conv1d(1, C, k, stride s) over the raw waveform, GroupNorm(C groups = per-channel statistics
// over time) and GELU (HF Wav2Vec2GroupNormConvLayer, called from the reference at models/tav.py:352 through
// wav2vec2.feature_extractor and at :476 inside wav2vec2(...)), plus their backward (GroupNorm affine and conv
// weight gradients; the waveform itself needs no gradient).
//
// Layout contract shared with the GEMM layers that follow (frontends.py): activations are bf16 [B, R, C] with R >= T
// padded rows per sample chosen so that R_{l-1} = stride_l * R_l, which turns every later strided Conv1d into one
// plain tcgen05 GEMM over the whole batch (A row pitch = stride*C, row length k*C: overlapping TMA rows).  Rows
// t >= T of a sample are padding: written as zeros here, and they never feed a valid output downstream.
//
// All three kernels are HBM/L2-bound streaming passes (C_in = 1: 2*k FLOP per output element).
#include "../../include/tavk.h"
#include "common.cuh"

namespace tavk {

constexpr int kConv0MaxK = 16;

// u[b, t, c] = sum_j wav[b, s*t + j] * w[c, j] (+ bias[c]); grid (ceil(R/64), B), 256 threads, 2 channels per thread
// per pass over C; the block's s*64+k samples are staged in shared memory.
__global__ void __launch_bounds__(256)
conv0_fwd_kernel(const float* __restrict__ wav, const float* __restrict__ w, const float* __restrict__ bias,
                 __nv_bfloat16* __restrict__ u, int L, int R, int T, int C, int k, int s) {
    extern __shared__ float sw[];
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * 64;
    const int nsamp = 63 * s + k;
    const float* wb = wav + (size_t)b * L;
    for (int i = threadIdx.x; i < nsamp; i += blockDim.x) {
        const int idx = t0 * s + i;
        sw[i] = idx < L ? wb[idx] : 0.f;
    }
    __syncthreads();
    for (int c = threadIdx.x * 2; c < C; c += blockDim.x * 2) {
        float w0[kConv0MaxK], w1[kConv0MaxK];
#pragma unroll
        for (int j = 0; j < kConv0MaxK; ++j) {
            w0[j] = j < k ? __ldg(w + (size_t)c * k + j) : 0.f;
            w1[j] = j < k ? __ldg(w + (size_t)(c + 1) * k + j) : 0.f;
        }
        const float b0 = bias ? bias[c] : 0.f, b1 = bias ? bias[c + 1] : 0.f;
        for (int r = 0; r < 64; ++r) {
            const int t = t0 + r;
            if (t >= R) break;
            float a0 = b0, a1 = b1;
#pragma unroll
            for (int j = 0; j < kConv0MaxK; ++j) {
                if (j < k) {
                    const float x = sw[r * s + j];
                    a0 = fmaf(x, w0[j], a0);
                    a1 = fmaf(x, w1[j], a1);
                }
            }
            if (t >= T) a0 = a1 = 0.f;
            *reinterpret_cast<uint32_t*>(u + ((size_t)b * R + t) * C + c) = pack_bf16x2(a0, a1);
        }
    }
}

// One block per (sample b, 16 channels): thread = (row lane 0..31, channel pair 0..7).
// Pass 1 mean, pass 2 centred variance (two-pass: no E[x^2]-E[x]^2 cancellation), pass 3 z = gamma*xhat + beta (bf16,
// saved for GELU') and a = GELU(z) (bf16, the next layer's GEMM operand).  The 16-channel slab of one sample
// (R x 32 B) stays in L2 between the passes.
constexpr int kGnCh = 16;
constexpr int kGnLanes = 32;

TAVK_DEVINL float2 gn_block_reduce2(float2 v, float2* red, int lane, int cp) {
    red[lane * (kGnCh / 2) + cp] = v;
    __syncthreads();
    if (lane == 0) {
        float2 a = make_float2(0.f, 0.f);
        for (int i = 0; i < kGnLanes; ++i) {
            const float2 q = red[i * (kGnCh / 2) + cp];
            a.x += q.x; a.y += q.y;
        }
        red[cp] = a;
    }
    __syncthreads();
    const float2 r = red[cp];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(256)
groupnorm_gelu_fwd_kernel(const __nv_bfloat16* __restrict__ u, const float* __restrict__ gamma,
                          const float* __restrict__ beta, __nv_bfloat16* __restrict__ z, __nv_bfloat16* __restrict__ a,
                          float* __restrict__ mean, float* __restrict__ rstd, int R, int T, int C, float eps) {
    __shared__ float2 red[kGnLanes * (kGnCh / 2)];
    const int b = blockIdx.y;
    const int cp = threadIdx.x & 7, lane = threadIdx.x >> 3;
    const int c = blockIdx.x * kGnCh + cp * 2;
    const size_t base = (size_t)b * R * C + c;
    float2 s = make_float2(0.f, 0.f);
    for (int t = lane; t < T; t += kGnLanes) {
        const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(u + base + (size_t)t * C));
        s.x += v.x; s.y += v.y;
    }
    s = gn_block_reduce2(s, red, lane, cp);
    const float m0 = s.x / T, m1 = s.y / T;
    float2 q = make_float2(0.f, 0.f);
    for (int t = lane; t < T; t += kGnLanes) {
        const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(u + base + (size_t)t * C));
        q.x += (v.x - m0) * (v.x - m0); q.y += (v.y - m1) * (v.y - m1);
    }
    q = gn_block_reduce2(q, red, lane, cp);
    const float r0 = rsqrtf(q.x / T + eps), r1 = rsqrtf(q.y / T + eps);
    if (lane == 0) {
        mean[(size_t)b * C + c] = m0; mean[(size_t)b * C + c + 1] = m1;
        rstd[(size_t)b * C + c] = r0; rstd[(size_t)b * C + c + 1] = r1;
    }
    const float g0 = gamma[c] * r0, g1 = gamma[c + 1] * r1;
    const float o0 = beta[c] - m0 * g0, o1 = beta[c + 1] - m1 * g1;
    for (int t = lane; t < R; t += kGnLanes) {
        float z0 = 0.f, z1 = 0.f, a0 = 0.f, a1 = 0.f;
        if (t < T) {
            const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(u + base + (size_t)t * C));
            z0 = fmaf(v.x, g0, o0); z1 = fmaf(v.y, g1, o1);
            gelu_fast2(z0, z1, a0, a1);
        }
        *reinterpret_cast<uint32_t*>(z + base + (size_t)t * C) = pack_bf16x2(z0, z1);
        *reinterpret_cast<uint32_t*>(a + base + (size_t)t * C) = pack_bf16x2(a0, a1);
    }
}

// Backward of GroupNorm + conv0 for dz (gradient w.r.t. the GroupNorm output, i.e. after GELU' was applied by the
// producing dgrad GEMM epilogue).  Same block shape as the forward.  Pass 1: s1 = sum_t dz, s2 = sum_t dz*xhat
// (-> dbeta, dgamma).  Pass 2: du = gamma*rstd*(dz - s1/T - xhat*s2/T) and dW[c, j] += sum_t du[t, c]*wav[s*t + j]
// (accumulated in registers, block-reduced, one atomic per (channel, tap) per block); du itself is never stored.
__global__ void __launch_bounds__(256)
groupnorm_conv0_bwd_kernel(const __nv_bfloat16* __restrict__ dz, const __nv_bfloat16* __restrict__ u,
                           const float* __restrict__ mean, const float* __restrict__ rstd,
                           const float* __restrict__ gamma, const float* __restrict__ wav, float* __restrict__ dgamma,
                           float* __restrict__ dbeta, float* __restrict__ dw, float* __restrict__ dbias, int L, int R,
                           int T, int C, int k, int s) {
    __shared__ float2 red[kGnLanes * (kGnCh / 2)];
    const int b = blockIdx.y;
    const int cp = threadIdx.x & 7, lane = threadIdx.x >> 3;
    const int c = blockIdx.x * kGnCh + cp * 2;
    const size_t base = (size_t)b * R * C + c;
    const float m0 = mean[(size_t)b * C + c], m1 = mean[(size_t)b * C + c + 1];
    const float r0 = rstd[(size_t)b * C + c], r1 = rstd[(size_t)b * C + c + 1];
    float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
    for (int t = lane; t < T; t += kGnLanes) {
        const float2 g = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dz + base + (size_t)t * C));
        const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(u + base + (size_t)t * C));
        s1.x += g.x; s1.y += g.y;
        s2.x += g.x * (v.x - m0) * r0; s2.y += g.y * (v.y - m1) * r1;
    }
    s1 = gn_block_reduce2(s1, red, lane, cp);
    s2 = gn_block_reduce2(s2, red, lane, cp);
    if (lane == 0) {
        atomicAdd(dbeta + c, s1.x);  atomicAdd(dbeta + c + 1, s1.y);
        atomicAdd(dgamma + c, s2.x); atomicAdd(dgamma + c + 1, s2.y);
    }
    const float k0 = gamma[c] * r0, k1 = gamma[c + 1] * r1;
    const float a0 = s1.x / T, a1 = s1.y / T, b0 = s2.x / T, b1 = s2.y / T;
    float acc0[kConv0MaxK], acc1[kConv0MaxK];
#pragma unroll
    for (int j = 0; j < kConv0MaxK; ++j) acc0[j] = acc1[j] = 0.f;
    float sb0 = 0.f, sb1 = 0.f;
    const float* wb = wav + (size_t)b * L;
    for (int t = lane; t < T; t += kGnLanes) {
        const float2 g = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dz + base + (size_t)t * C));
        const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(u + base + (size_t)t * C));
        const float du0 = k0 * (g.x - a0 - (v.x - m0) * r0 * b0);
        const float du1 = k1 * (g.y - a1 - (v.y - m1) * r1 * b1);
        sb0 += du0; sb1 += du1;
#pragma unroll
        for (int j = 0; j < kConv0MaxK; ++j) {
            if (j < k) {
                const int idx = t * s + j;
                const float x = idx < L ? __ldg(wb + idx) : 0.f;
                acc0[j] = fmaf(du0, x, acc0[j]);
                acc1[j] = fmaf(du1, x, acc1[j]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < kConv0MaxK; ++j) {
        if (j < k) {   // uniform
            const float2 tot = gn_block_reduce2(make_float2(acc0[j], acc1[j]), red, lane, cp);
            if (lane == 0) {
                atomicAdd(dw + (size_t)c * k + j, tot.x);
                atomicAdd(dw + (size_t)(c + 1) * k + j, tot.y);
            }
        }
    }
    if (dbias != nullptr) {
        const float2 tot = gn_block_reduce2(make_float2(sb0, sb1), red, lane, cp);
        if (lane == 0) {
            atomicAdd(dbias + c, tot.x);
            atomicAdd(dbias + c + 1, tot.y);
        }
    }
}

}  // namespace tavk

using namespace tavk;

extern "C" int tavk_conv0_fwd(const float* wav, const float* w, const float* bias, void* u, int B, int L, int R, int T,
                              int C, int k, int s, void* stream) {
    TAVK_CHECK(wav && w && u, 1, "tavk_conv0_fwd: null pointer");
    TAVK_CHECK(B > 0 && L > 0 && R >= T && T > 0, 1, "tavk_conv0_fwd: bad sizes B=%d L=%d R=%d T=%d", B, L, R, T);
    TAVK_CHECK(C % 2 == 0 && k >= 1 && k <= kConv0MaxK && s >= 1, 1, "tavk_conv0_fwd: C even, 1<=k<=%d (C=%d k=%d)",
               kConv0MaxK, C, k);
    TAVK_CHECK((T - 1) * (long long)s + k <= L, 1, "tavk_conv0_fwd: T=%d frames need more than L=%d samples", T, L);
    dim3 grid((R + 63) / 64, B);
    const size_t smem = (size_t)(63 * s + k) * sizeof(float);
    conv0_fwd_kernel<<<grid, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
        wav, w, bias, reinterpret_cast<__nv_bfloat16*>(u), L, R, T, C, k, s);
    TAVK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tavk_groupnorm_gelu_fwd(const void* u, const float* gamma, const float* beta, void* z, void* a,
                                       float* mean, float* rstd, int B, int R, int T, int C, float eps, void* stream) {
    TAVK_CHECK(u && gamma && beta && z && a && mean && rstd, 1, "tavk_groupnorm_gelu_fwd: null pointer");
    TAVK_CHECK(C % kGnCh == 0 && R >= T && T > 0 && B > 0, 1, "tavk_groupnorm_gelu_fwd: C %% %d == 0 required (C=%d)", kGnCh, C);
    dim3 grid(C / kGnCh, B);
    groupnorm_gelu_fwd_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(u), gamma, beta, reinterpret_cast<__nv_bfloat16*>(z),
        reinterpret_cast<__nv_bfloat16*>(a), mean, rstd, R, T, C, eps);
    TAVK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tavk_groupnorm_conv0_bwd(const void* dz, const void* u, const float* mean, const float* rstd,
                                        const float* gamma, const float* wav, float* dgamma, float* dbeta, float* dw,
                                        float* dbias, int B, int L, int R, int T, int C, int k, int s, void* stream) {
    TAVK_CHECK(dz && u && mean && rstd && gamma && wav && dgamma && dbeta && dw, 1, "tavk_groupnorm_conv0_bwd: null pointer");
    TAVK_CHECK(C % kGnCh == 0 && R >= T && T > 0 && B > 0 && k >= 1 && k <= kConv0MaxK, 1,
               "tavk_groupnorm_conv0_bwd: bad sizes C=%d k=%d", C, k);
    dim3 grid(C / kGnCh, B);
    groupnorm_conv0_bwd_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(dz), reinterpret_cast<const __nv_bfloat16*>(u), mean, rstd, gamma, wav,
        dgamma, dbeta, dw, dbias, L, R, T, C, k, s);
    TAVK_CUDA(cudaGetLastError());
    return 0;
}
