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
// All kernels here are HBM/L2-bound streaming passes (C_in = 1: 2*k FLOP per output element).
#include "../../include/tavk.h"
#include "common.cuh"

namespace tavk {

constexpr int kConv0MaxK = 16;

// u[b, t, c] = sum_j wav[b, s*t + j] * w[c, j] (+ bias[c]); grid (ceil(R/64), B), 256 threads, 2 channels per thread
// per pass over C; the block's s*64+k samples are staged in shared memory.
__global__ void __launch_bounds__(256)
conv0_fwd_kernel(const float* __restrict__ wav, const float* __restrict__ w, const float* __restrict__ bias,
                 __nv_bfloat16* __restrict__ u, int L, int R, int T, int C, int k, int s) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
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

// ---- GroupNorm (C groups = per-channel statistics over time) + GELU, and its backward, as streaming passes --------
// Common shape: grid (C/64, row chunks, B), 256 threads = 32 row lanes x 8 channel octets; every access is a 16-byte
// (8 x bf16) load/store, 128 contiguous bytes per row per block.  Per-(sample, channel) sums are reduced by warp
// shuffles, shared memory across the 8 warps, then one fp32 atomic per channel per block.
constexpr int kGnRowsPerBlock = 512;

struct bf16x8 {
    float v[8];
};
TAVK_DEVINL bf16x8 ld_bf16x8(const __nv_bfloat16* p) {
    const uint4 q = *reinterpret_cast<const uint4*>(p);
    bf16x8 r;
    float2 t;
    t = unpack_bf16x2(q.x); r.v[0] = t.x; r.v[1] = t.y;
    t = unpack_bf16x2(q.y); r.v[2] = t.x; r.v[3] = t.y;
    t = unpack_bf16x2(q.z); r.v[4] = t.x; r.v[5] = t.y;
    t = unpack_bf16x2(q.w); r.v[6] = t.x; r.v[7] = t.y;
    return r;
}
TAVK_DEVINL void st_bf16x8(__nv_bfloat16* p, const float (&v)[8]) {
    *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                              pack_bf16x2(v[6], v[7]));
}
// acc[2][8] per thread -> out0/out1[c0 + 8*cg + i] += block totals
TAVK_DEVINL void gn_block_accumulate(float (&a0)[8], float (&a1)[8], float* red, float* out0, float* out1, int c0) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, cg = threadIdx.x & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a0[i] += __shfl_xor_sync(0xffffffffu, a0[i], 8);  a1[i] += __shfl_xor_sync(0xffffffffu, a1[i], 8);
        a0[i] += __shfl_xor_sync(0xffffffffu, a0[i], 16); a1[i] += __shfl_xor_sync(0xffffffffu, a1[i], 16);
    }
    if (lane < 8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            red[(warp * 64 + cg * 8 + i) * 2] = a0[i];
            red[(warp * 64 + cg * 8 + i) * 2 + 1] = a1[i];
        }
    }
    __syncthreads();
    if (threadIdx.x < 64) {
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            s0 += red[(w * 64 + threadIdx.x) * 2];
            s1 += red[(w * 64 + threadIdx.x) * 2 + 1];
        }
        atomicAdd(out0 + c0 + threadIdx.x, s0);
        atomicAdd(out1 + c0 + threadIdx.x, s1);
    }
}

// sums[b][0][c] += sum_t (u - k), sums[b][1][c] += sum_t (u - k)^2 with the shift k = u[b, 0, c] (removes the mean's
// bulk, so the one-pass variance does not cancel)
__global__ void __launch_bounds__(256)
groupnorm_stats_kernel(const __nv_bfloat16* __restrict__ u, float* __restrict__ sums, int R, int T, int C) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    __shared__ float red[8 * 64 * 2];
    const int b = blockIdx.z, c0 = blockIdx.x * 64, cg = threadIdx.x & 7, rl = threadIdx.x >> 3;
    const __nv_bfloat16* base = u + (size_t)b * R * C + c0 + cg * 8;
    const bf16x8 k = ld_bf16x8(base);
    const int t0 = blockIdx.y * kGnRowsPerBlock, t1 = min(t0 + kGnRowsPerBlock, T);
    float s[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
#pragma unroll 4
    for (int t = t0 + rl; t < t1; t += 32) {
        const bf16x8 x = ld_bf16x8(base + (size_t)t * C);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float d = x.v[i] - k.v[i];
            s[i] += d;
            q[i] = fmaf(d, d, q[i]);
        }
    }
    gn_block_accumulate(s, q, red, sums + (size_t)b * 2 * C, sums + (size_t)b * 2 * C + C, c0);
}

// mean/rstd from the shifted sums, z = gamma*xhat + beta (bf16, kept for GELU'), a = GELU(z) (bf16, next GEMM operand);
// padding rows t >= T are written as zeros
__global__ void __launch_bounds__(256)
groupnorm_gelu_apply_kernel(const __nv_bfloat16* __restrict__ u, const float* __restrict__ sums,
                            const float* __restrict__ gamma, const float* __restrict__ beta,
                            __nv_bfloat16* __restrict__ z, __nv_bfloat16* __restrict__ a, float* __restrict__ mean,
                            float* __restrict__ rstd, int R, int T, int C, float eps) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    const int b = blockIdx.z, c0 = blockIdx.x * 64, cg = threadIdx.x & 7, rl = threadIdx.x >> 3;
    const int c = c0 + cg * 8;
    const size_t base = (size_t)b * R * C + c;
    const bf16x8 k = ld_bf16x8(u + base);
    float g[8], o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float sd = sums[(size_t)b * 2 * C + c + i] / T;
        const float var = fmaxf(sums[(size_t)b * 2 * C + C + c + i] / T - sd * sd, 0.f);
        const float m = k.v[i] + sd, r = rsqrtf(var + eps);
        if (blockIdx.y == 0 && rl == 0) { mean[(size_t)b * C + c + i] = m; rstd[(size_t)b * C + c + i] = r; }
        g[i] = gamma[c + i] * r;
        o[i] = beta[c + i] - m * g[i];
    }
    const int t0 = blockIdx.y * kGnRowsPerBlock, t1 = min(t0 + kGnRowsPerBlock, R);
#pragma unroll 2
    for (int t = t0 + rl; t < t1; t += 32) {
        float zz[8], aa[8];
        if (t < T) {
            const bf16x8 x = ld_bf16x8(u + base + (size_t)t * C);
#pragma unroll
            for (int i = 0; i < 8; ++i) zz[i] = fmaf(x.v[i], g[i], o[i]);
#pragma unroll
            for (int i = 0; i < 8; i += 2) gelu_fast2(zz[i], zz[i + 1], aa[i], aa[i + 1]);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) zz[i] = aa[i] = 0.f;
        }
        st_bf16x8(z + base + (size_t)t * C, zz);
        st_bf16x8(a + base + (size_t)t * C, aa);
    }
}

// backward pass 1: sums[b][0][c] += sum_t dz, sums[b][1][c] += sum_t dz * xhat
__global__ void __launch_bounds__(256)
groupnorm_bwd_stats_kernel(const __nv_bfloat16* __restrict__ dz, const __nv_bfloat16* __restrict__ u,
                           const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ sums,
                           int R, int T, int C) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    __shared__ float red[8 * 64 * 2];
    const int b = blockIdx.z, c0 = blockIdx.x * 64, cg = threadIdx.x & 7, rl = threadIdx.x >> 3;
    const int c = c0 + cg * 8;
    const size_t base = (size_t)b * R * C + c;
    float m[8], r[8], s1[8], s2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        m[i] = mean[(size_t)b * C + c + i];
        r[i] = rstd[(size_t)b * C + c + i];
        s1[i] = s2[i] = 0.f;
    }
    const int t0 = blockIdx.y * kGnRowsPerBlock, t1 = min(t0 + kGnRowsPerBlock, T);
#pragma unroll 4
    for (int t = t0 + rl; t < t1; t += 32) {
        const bf16x8 g = ld_bf16x8(dz + base + (size_t)t * C);
        const bf16x8 x = ld_bf16x8(u + base + (size_t)t * C);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s1[i] += g.v[i];
            s2[i] = fmaf(g.v[i], (x.v[i] - m[i]) * r[i], s2[i]);
        }
    }
    gn_block_accumulate(s1, s2, red, sums + (size_t)b * 2 * C, sums + (size_t)b * 2 * C + C, c0);
}

// backward pass 2: du = gamma*rstd*(dz - s1/T - xhat*s2/T) (bf16; zero in the padding rows): the conv0 weight gradient is
// then the tcgen05 wgrad GEMM du^T x [waveform windows]
__global__ void __launch_bounds__(256)
groupnorm_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dz, const __nv_bfloat16* __restrict__ u,
                           const float* __restrict__ mean, const float* __restrict__ rstd,
                           const float* __restrict__ gamma, const float* __restrict__ sums,
                           __nv_bfloat16* __restrict__ du, int R, int T, int C) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    const int b = blockIdx.z, c0 = blockIdx.x * 64, cg = threadIdx.x & 7, rl = threadIdx.x >> 3;
    const int c = c0 + cg * 8;
    const size_t base = (size_t)b * R * C + c;
    float m[8], r[8], kk[8], a1[8], a2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        m[i] = mean[(size_t)b * C + c + i];
        r[i] = rstd[(size_t)b * C + c + i];
        kk[i] = gamma[c + i] * r[i];
        a1[i] = sums[(size_t)b * 2 * C + c + i] / T;
        a2[i] = sums[(size_t)b * 2 * C + C + c + i] / T;
    }
    const int t0 = blockIdx.y * kGnRowsPerBlock, t1 = min(t0 + kGnRowsPerBlock, R);
#pragma unroll 2
    for (int t = t0 + rl; t < t1; t += 32) {
        float d[8];
        if (t < T) {
            const bf16x8 g = ld_bf16x8(dz + base + (size_t)t * C);
            const bf16x8 x = ld_bf16x8(u + base + (size_t)t * C);
#pragma unroll
            for (int i = 0; i < 8; ++i) d[i] = kk[i] * (g.v[i] - a1[i] - (x.v[i] - m[i]) * r[i] * a2[i]);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) d[i] = 0.f;
        }
        st_bf16x8(du + base + (size_t)t * C, d);
    }
}

// ---- layer-norm family (HF Wav2Vec2LayerNormConvLayer: Conv1d + bias -> LayerNorm over the C channels -> GELU; the
// wav2vec2-large feature encoder the reference actually loads, models/tav.py:257,455) ------------------------------
// One warp per (b, t) row of the channels-last activations [B, R, C], C % 256 == 0, C <= 1024: lane holds C/32 values as
// 16-byte groups of 8 bf16; fp32 statistics (two-pass in registers); rows t >= T are padding (written as zeros).
constexpr int kLnGMaxGroups = 4;
__global__ void __launch_bounds__(256)
chan_ln_gelu_fwd_kernel(const __nv_bfloat16* __restrict__ u, const float* __restrict__ gamma, const float* __restrict__ beta,
                        __nv_bfloat16* __restrict__ a, float* __restrict__ mean, float* __restrict__ rstd, long long rows,
                        int R, int T, int C, float eps) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    const int lane = threadIdx.x & 31;
    const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (row >= rows) return;
    const int ng = C >> 8;
    const bool pad = (int)(row % R) >= T;
    float x[kLnGMaxGroups][8];
    float sum = 0.f;
#pragma unroll
    for (int g = 0; g < kLnGMaxGroups; ++g) {
        if (g < ng) {
            const uint4 q = *reinterpret_cast<const uint4*>(u + row * C + (g * 32 + lane) * 8);
            const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = unpack_bf16x2(w4[i]);
                x[g][2 * i] = f.x; x[g][2 * i + 1] = f.y;
                sum += f.x + f.y;
            }
        }
    }
    const float mu = warp_sum(sum) / (float)C;
    float var = 0.f;
#pragma unroll
    for (int g = 0; g < kLnGMaxGroups; ++g)
        if (g < ng)
#pragma unroll
            for (int i = 0; i < 8; ++i) { const float d = x[g][i] - mu; var += d * d; }
    const float rs = rsqrtf(warp_sum(var) / (float)C + eps);
    if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
#pragma unroll
    for (int g = 0; g < kLnGMaxGroups; ++g) {
        if (g < ng) {
            const int c0 = (g * 32 + lane) * 8;
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float z0 = (x[g][2 * i] - mu) * rs * __ldg(gamma + c0 + 2 * i) + __ldg(beta + c0 + 2 * i);
                const float z1 = (x[g][2 * i + 1] - mu) * rs * __ldg(gamma + c0 + 2 * i + 1) + __ldg(beta + c0 + 2 * i + 1);
                o[i] = pad ? 0u : pack_bf16x2(gelu_erf(z0), gelu_erf(z1));
            }
            *reinterpret_cast<uint4*>(a + row * C + c0) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}
// du = LN'(gelu'(z) * da) with z recomputed from (u, mean, rstd); dgamma += sum dz*xhat, dbeta += sum dz (block-level
// reduction, one atomic per channel per block); du is zero in the padding rows
__global__ void __launch_bounds__(256)
chan_ln_gelu_bwd_kernel(const __nv_bfloat16* __restrict__ da, const __nv_bfloat16* __restrict__ u,
                        const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                        const float* __restrict__ beta, __nv_bfloat16* __restrict__ du, float* __restrict__ dgamma,
                        float* __restrict__ dbeta, long long rows, int R, int T, int C, int rows_per_warp) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    extern __shared__ float s_red[];        // [2][C] block partial sums
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int ng = C >> 8;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_red[i] = 0.f;
    __syncthreads();
    float dg[kLnGMaxGroups][8], db[kLnGMaxGroups][8];
#pragma unroll
    for (int g = 0; g < kLnGMaxGroups; ++g)
#pragma unroll
        for (int i = 0; i < 8; ++i) { dg[g][i] = 0.f; db[g][i] = 0.f; }
    const long long row0 = ((long long)blockIdx.x * nwarps + warp) * rows_per_warp;
    for (int rr = 0; rr < rows_per_warp; ++rr) {
        const long long row = row0 + rr;
        if (row >= rows) break;
        const bool pad = (int)(row % R) >= T;
        if (pad) {
#pragma unroll
            for (int g = 0; g < kLnGMaxGroups; ++g)
                if (g < ng) *reinterpret_cast<uint4*>(du + row * C + (g * 32 + lane) * 8) = make_uint4(0u, 0u, 0u, 0u);
            continue;
        }
        const float mu = mean[row], rs = rstd[row];
        float xh[kLnGMaxGroups][8], dzg[kLnGMaxGroups][8];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int g = 0; g < kLnGMaxGroups; ++g) {
            if (g < ng) {
                const int c0 = (g * 32 + lane) * 8;
                const uint4 qu = *reinterpret_cast<const uint4*>(u + row * C + c0);
                const uint4 qa = *reinterpret_cast<const uint4*>(da + row * C + c0);
                const uint32_t wu[4] = {qu.x, qu.y, qu.z, qu.w}, wa[4] = {qa.x, qa.y, qa.z, qa.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 fu = unpack_bf16x2(wu[i]), fa = unpack_bf16x2(wa[i]);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int e = 2 * i + h;
                        const float gm = __ldg(gamma + c0 + e);
                        const float xhat = ((h ? fu.y : fu.x) - mu) * rs;
                        const float z = xhat * gm + __ldg(beta + c0 + e);
                        const float dz = (h ? fa.y : fa.x) * gelu_erf_grad(z);
                        xh[g][e] = xhat;
                        dzg[g][e] = dz * gm;
                        dg[g][e] += dz * xhat;
                        db[g][e] += dz;
                        s1 += dzg[g][e];
                        s2 += dzg[g][e] * xhat;
                    }
                }
            }
        }
        s1 = warp_sum(s1) / (float)C;
        s2 = warp_sum(s2) / (float)C;
#pragma unroll
        for (int g = 0; g < kLnGMaxGroups; ++g) {
            if (g < ng) {
                uint32_t o[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    o[i] = pack_bf16x2(rs * (dzg[g][2 * i] - s1 - xh[g][2 * i] * s2), rs * (dzg[g][2 * i + 1] - s1 - xh[g][2 * i + 1] * s2));
                *reinterpret_cast<uint4*>(du + row * C + (g * 32 + lane) * 8) = make_uint4(o[0], o[1], o[2], o[3]);
            }
        }
    }
#pragma unroll
    for (int g = 0; g < kLnGMaxGroups; ++g)
        if (g < ng)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                atomicAdd(&s_red[(g * 32 + lane) * 8 + i], dg[g][i]);
                atomicAdd(&s_red[C + (g * 32 + lane) * 8 + i], db[g][i]);
            }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        atomicAdd(dgamma + i, s_red[i]);
        atomicAdd(dbeta + i, s_red[C + i]);
    }
}

// win[b*R + t, j] = bf16(wav[b, s*t + j]) for j < k (0 for k <= j < 16 and for rows t >= T): the B operand of the conv0
// weight-gradient GEMM
__global__ void wave_windows_kernel(const float* __restrict__ wav, __nv_bfloat16* __restrict__ win, int L, int R, int T,
                                    int k, int s, long long total) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(i & 15);
        const long long row = i >> 4;
        const int t = (int)(row % R);
        const long long b = row / R;
        const int idx = t * s + j;
        float v = 0.f;
        if (j < k && t < T && idx < L) v = wav[b * L + idx];
        win[i] = __float2bfloat16(v);
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
    TAVK_CUDA(launch_kernel(conv0_fwd_kernel, dim3(grid), dim3(256), (size_t)(smem), reinterpret_cast<cudaStream_t>(stream), 
        wav, w, bias, reinterpret_cast<__nv_bfloat16*>(u), L, R, T, C, k, s));
    TAVK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tavk_groupnorm_gelu_fwd(const void* u, const float* gamma, const float* beta, void* z, void* a,
                                       float* mean, float* rstd, float* sums_ws, int B, int R, int T, int C, float eps,
                                       void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    TAVK_CHECK(u && gamma && beta && z && a && mean && rstd && sums_ws, 1, "tavk_groupnorm_gelu_fwd: null pointer");
    TAVK_CHECK(C % 64 == 0 && R >= T && T > 0 && B > 0, 1, "tavk_groupnorm_gelu_fwd: C %% 64 == 0 required (C=%d)", C);
    TAVK_CUDA(cudaMemsetAsync(sums_ws, 0, (size_t)B * 2 * C * sizeof(float), stream));
    dim3 g1(C / 64, (T + kGnRowsPerBlock - 1) / kGnRowsPerBlock, B), g2(C / 64, (R + kGnRowsPerBlock - 1) / kGnRowsPerBlock, B);
    TAVK_CUDA(launch_kernel(groupnorm_stats_kernel, dim3(g1), dim3(256), (size_t)(0), stream, reinterpret_cast<const __nv_bfloat16*>(u), sums_ws, R, T, C));
    TAVK_CUDA(launch_kernel(groupnorm_gelu_apply_kernel, dim3(g2), dim3(256), (size_t)(0), stream, reinterpret_cast<const __nv_bfloat16*>(u), sums_ws, gamma, beta,
                                                        reinterpret_cast<__nv_bfloat16*>(z),
                                                        reinterpret_cast<__nv_bfloat16*>(a), mean, rstd, R, T, C, eps));
    TAVK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tavk_groupnorm_bwd(const void* dz, const void* u, const float* mean, const float* rstd,
                                  const float* gamma, void* du, float* sums_ws, int B, int R, int T, int C,
                                  void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    TAVK_CHECK(dz && u && mean && rstd && gamma && du && sums_ws, 1, "tavk_groupnorm_bwd: null pointer");
    TAVK_CHECK(C % 64 == 0 && R >= T && T > 0 && B > 0, 1, "tavk_groupnorm_bwd: C %% 64 == 0 required (C=%d)", C);
    TAVK_CUDA(cudaMemsetAsync(sums_ws, 0, (size_t)B * 2 * C * sizeof(float), stream));
    dim3 g1(C / 64, (T + kGnRowsPerBlock - 1) / kGnRowsPerBlock, B), g2(C / 64, (R + kGnRowsPerBlock - 1) / kGnRowsPerBlock, B);
    TAVK_CUDA(launch_kernel(groupnorm_bwd_stats_kernel, dim3(g1), dim3(256), (size_t)(0), stream, reinterpret_cast<const __nv_bfloat16*>(dz),
                                                       reinterpret_cast<const __nv_bfloat16*>(u), mean, rstd, sums_ws, R, T, C));
    TAVK_CUDA(launch_kernel(groupnorm_bwd_apply_kernel, dim3(g2), dim3(256), (size_t)(0), stream, reinterpret_cast<const __nv_bfloat16*>(dz),
                                                       reinterpret_cast<const __nv_bfloat16*>(u), mean, rstd, gamma, sums_ws,
                                                       reinterpret_cast<__nv_bfloat16*>(du), R, T, C));
    TAVK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tavk_chan_ln_gelu_fwd(const void* u, const float* gamma, const float* beta, void* a, float* mean, float* rstd,
                                     int B, int R, int T, int C, float eps, void* stream) {
    TAVK_CHECK(u && gamma && beta && a && mean && rstd, 1, "tavk_chan_ln_gelu_fwd: null pointer");
    TAVK_CHECK(C % 256 == 0 && C <= 256 * kLnGMaxGroups && R >= T && T >= 0, 2, "tavk_chan_ln_gelu_fwd: unsupported shape C=%d R=%d T=%d", C, R, T);
    const long long rows = (long long)B * R;
    if (rows <= 0) return 0;
    const long long blocks = (rows * 32 + 255) / 256;
    TAVK_CUDA(launch_kernel(chan_ln_gelu_fwd_kernel, dim3((unsigned)blocks), dim3(256), (size_t)0, reinterpret_cast<cudaStream_t>(stream),
                            reinterpret_cast<const __nv_bfloat16*>(u), gamma, beta, reinterpret_cast<__nv_bfloat16*>(a), mean, rstd,
                            rows, R, T, C, eps));
    TAVK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tavk_chan_ln_gelu_bwd(const void* da, const void* u, const float* mean, const float* rstd, const float* gamma,
                                     const float* beta, void* du, float* dgamma, float* dbeta, int B, int R, int T, int C,
                                     void* stream) {
    TAVK_CHECK(da && u && mean && rstd && gamma && beta && du && dgamma && dbeta, 1, "tavk_chan_ln_gelu_bwd: null pointer");
    TAVK_CHECK(C % 256 == 0 && C <= 256 * kLnGMaxGroups && R >= T && T >= 0, 2, "tavk_chan_ln_gelu_bwd: unsupported shape C=%d R=%d T=%d", C, R, T);
    const long long rows = (long long)B * R;
    if (rows <= 0) return 0;
    const int rows_per_warp = 8;
    const long long warps = (rows + rows_per_warp - 1) / rows_per_warp;
    const long long blocks = (warps + 7) / 8;
    TAVK_CUDA(launch_kernel(chan_ln_gelu_bwd_kernel, dim3((unsigned)blocks), dim3(256), (size_t)(2 * C * sizeof(float)),
                            reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<const __nv_bfloat16*>(da),
                            reinterpret_cast<const __nv_bfloat16*>(u), mean, rstd, gamma, beta, reinterpret_cast<__nv_bfloat16*>(du),
                            dgamma, dbeta, rows, R, T, C, rows_per_warp));
    TAVK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tavk_wave_windows(const float* wav, void* win, int B, int L, int R, int T, int k, int s, void* stream) {
    TAVK_CHECK(wav && win, 1, "tavk_wave_windows: null pointer");
    TAVK_CHECK(k >= 1 && k <= 16 && s >= 1 && B > 0 && R >= T, 1, "tavk_wave_windows: 1 <= k <= 16 (k=%d)", k);
    const long long total = (long long)B * R * 16;
    const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    TAVK_CUDA(launch_kernel(wave_windows_kernel, dim3(blocks), dim3(256), (size_t)(0), reinterpret_cast<cudaStream_t>(stream), 
        wav, reinterpret_cast<__nv_bfloat16*>(win), L, R, T, k, s, total));
    TAVK_CUDA(cudaGetLastError());
    return 0;
}
