// LayerNorm forward / backward over the last dim (HBM-bound, one warp per row, float4 accesses).
// Replaces nn.LayerNorm at reference utils/TAVFormer.py:237,239 (eps 1e-12), :108,:118 (eps 1e-5) and
// models/tav.py:439,443,445,447, plus the LayerNorms inside the HF RoBERTa / Wav2Vec2 / VideoMAE layers.
// Statistics are fp32 and two-pass (mean, then centred variance) because the reference's post-softmax mask quirk
// drives the residual stream to 1e6..4e7 where E[x^2]-mu^2 cancels catastrophically.
#include <stdlib.h>

#include "../../include/tavk.h"
#include "common.cuh"

namespace tavk {

constexpr int kLnMaxVec = 8;  // float4 per lane -> H <= 1024

// ---------------------------------------------------------------- forward
template <int VEC>
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     __nv_bfloat16* __restrict__ y_bf16, float* __restrict__ y_f32, float* __restrict__ mean_out,
                     float* __restrict__ rstd_out, int M, float eps) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    constexpr int H = VEC * 128;
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    if (row >= M) return;
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * H);
    float4 v[VEC];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        v[i] = xr[lane + 32 * i];
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mu = warp_sum(s) * (1.0f / H);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const float a = v[i].x - mu, b = v[i].y - mu, c = v[i].z - mu, d = v[i].w - mu;
        q += (a * a + b * b) + (c * c + d * d);
    }
    const float var = warp_sum(q) * (1.0f / H);
    const float rs = rsqrtf(var + eps);
    if (lane == 0) {
        if (mean_out) mean_out[row] = mu;
        if (rstd_out) rstd_out[row] = rs;
    }
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const float4 g = __ldg(g4 + lane + 32 * i);
        const float4 b = __ldg(b4 + lane + 32 * i);
        float4 o;
        o.x = (v[i].x - mu) * rs * g.x + b.x;
        o.y = (v[i].y - mu) * rs * g.y + b.y;
        o.z = (v[i].z - mu) * rs * g.z + b.z;
        o.w = (v[i].w - mu) * rs * g.w + b.w;
        if (y_f32) reinterpret_cast<float4*>(y_f32 + (size_t)row * H)[lane + 32 * i] = o;
        if (y_bf16) {
            uint2 p;
            p.x = pack_bf16x2(o.x, o.y);
            p.y = pack_bf16x2(o.z, o.w);
            reinterpret_cast<uint2*>(y_bf16 + (size_t)row * H)[lane + 32 * i] = p;
        }
    }
}

// ---------------------------------------------------------------- backward
// dx = (resid) + rstd * (g - mean(g) - xhat * mean(g*xhat)),  g = dy * gamma
// dgamma += sum_rows dy * xhat, dbeta += sum_rows dy   (per-warp register partials -> smem -> one atomic per column
// per block)
template <int VEC>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                     const float* __restrict__ rstd, const float* __restrict__ gamma,
                     const float* __restrict__ dx_resid, float* __restrict__ dx_f32,
                     __nv_bfloat16* __restrict__ dx_bf16, float* __restrict__ dgamma, float* __restrict__ dbeta,
                     float* __restrict__ dx_colsum, int M) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    constexpr int H = VEC * 128;
    extern __shared__ float red[];  // [warps][H]
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int warps_per_block = blockDim.x >> 5;
    float4 g4[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) g4[i] = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
    float4 dg[VEC], db[VEC], dc[VEC];   // dc: column sums of dx (the bias gradient of the Linear that produced x)
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        dc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int row = blockIdx.x * warps_per_block + warp; row < M; row += gridDim.x * warps_per_block) {
        const float mu = mean[row], rs = rstd[row];
        const float4* dyr = reinterpret_cast<const float4*>(dy + (size_t)row * H);
        const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * H);
        float4 d[VEC], xh[VEC], rr[VEC];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) d[i] = dyr[lane + 32 * i];
        if (dx_resid) {     // requested with dy and x: one more operand row in flight per warp instead of a second round trip
#pragma unroll
            for (int i = 0; i < VEC; ++i) rr[i] = reinterpret_cast<const float4*>(dx_resid + (size_t)row * H)[lane + 32 * i];
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float4 xv = xr[lane + 32 * i];
            xh[i].x = (xv.x - mu) * rs; xh[i].y = (xv.y - mu) * rs;
            xh[i].z = (xv.z - mu) * rs; xh[i].w = (xv.w - mu) * rs;
            dg[i].x += d[i].x * xh[i].x; dg[i].y += d[i].y * xh[i].y;
            dg[i].z += d[i].z * xh[i].z; dg[i].w += d[i].w * xh[i].w;
            db[i].x += d[i].x; db[i].y += d[i].y; db[i].z += d[i].z; db[i].w += d[i].w;
            // g = dy * gamma (reuse d)
            d[i].x *= g4[i].x; d[i].y *= g4[i].y; d[i].z *= g4[i].z; d[i].w *= g4[i].w;
            s1 += (d[i].x + d[i].y) + (d[i].z + d[i].w);
            s2 += (d[i].x * xh[i].x + d[i].y * xh[i].y) + (d[i].z * xh[i].z + d[i].w * xh[i].w);
        }
        const float c2 = warp_sum(s1) * (1.0f / H);
        const float c1 = warp_sum(s2) * (1.0f / H);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            float4 o;
            o.x = rs * (d[i].x - c2 - xh[i].x * c1);
            o.y = rs * (d[i].y - c2 - xh[i].y * c1);
            o.z = rs * (d[i].z - c2 - xh[i].z * c1);
            o.w = rs * (d[i].w - c2 - xh[i].w * c1);
            if (dx_resid) { o.x += rr[i].x; o.y += rr[i].y; o.z += rr[i].z; o.w += rr[i].w; }
            dc[i].x += o.x; dc[i].y += o.y; dc[i].z += o.z; dc[i].w += o.w;
            if (dx_f32) reinterpret_cast<float4*>(dx_f32 + (size_t)row * H)[lane + 32 * i] = o;
            if (dx_bf16) {
                uint2 p;
                p.x = pack_bf16x2(o.x, o.y);
                p.y = pack_bf16x2(o.z, o.w);
                reinterpret_cast<uint2*>(dx_bf16 + (size_t)row * H)[lane + 32 * i] = p;
            }
        }
    }
    // block reduction of the parameter gradients, dgamma, dbeta, then the dx column sums through the same smem buffer
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
        float* out = pass == 0 ? dgamma : (pass == 1 ? dbeta : dx_colsum);
        if (out == nullptr) continue;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < VEC; ++i)
            reinterpret_cast<float4*>(red + warp * H)[lane + 32 * i] = pass == 0 ? dg[i] : (pass == 1 ? db[i] : dc[i]);
        __syncthreads();
        for (int c = threadIdx.x; c < H; c += blockDim.x) {
            float s = 0.f;
            for (int w = 0; w < warps_per_block; ++w) s += red[w * H + c];
            atomicAdd(out + c, s);
        }
    }
}

}  // namespace tavk

using namespace tavk;

extern "C" int tavk_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y_bf16, float* y_f32,
                                  float* mean, float* rstd, int M, int H, float eps, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    TAVK_CHECK(x && gamma && beta && (y_bf16 || y_f32), 1, "tavk_layernorm_fwd: null pointer");
    TAVK_CHECK(M >= 0 && H > 0, 1, "tavk_layernorm_fwd: bad shape M=%d H=%d", M, H);
    TAVK_CHECK(H % 128 == 0 && H / 128 <= kLnMaxVec, 2, "tavk_layernorm_fwd: H=%d unsupported (multiple of 128, <=1024)",
               H);
    if (M == 0) return 0;
    const int warps = 8;
    const int grid = (M + warps - 1) / warps;
    __nv_bfloat16* yb = reinterpret_cast<__nv_bfloat16*>(y_bf16);
#define LN_FWD(V)                                                                                              \
    case V:                                                                                                    \
        TAVK_CUDA(launch_kernel(layernorm_fwd_kernel<V>, dim3(grid), dim3(warps * 32), (size_t)(0), stream, x, gamma, beta, yb, y_f32, mean, rstd, M, eps)); \
        break;
    switch (H / 128) {
        LN_FWD(1) LN_FWD(2) LN_FWD(3) LN_FWD(4) LN_FWD(5) LN_FWD(6) LN_FWD(7) LN_FWD(8)
    }
#undef LN_FWD
    TAVK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tavk_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd,
                                  const float* gamma, const float* dx_resid, float* dx_f32, void* dx_bf16,
                                  float* dgamma, float* dbeta, float* dx_colsum, int M, int H, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    TAVK_CHECK(dy && x && mean && rstd && gamma, 1, "tavk_layernorm_bwd: null pointer");
    TAVK_CHECK(M >= 0 && H > 0, 1, "tavk_layernorm_bwd: bad shape M=%d H=%d", M, H);
    TAVK_CHECK(H % 128 == 0 && H / 128 <= kLnMaxVec, 2, "tavk_layernorm_bwd: H=%d unsupported (multiple of 128, <=1024)",
               H);
    if (M == 0) return 0;
    const int warps = 8;
    int grid = (M + warps - 1) / warps;
    // one RESIDENT wave: the kernel keeps ~180 registers per thread (three column accumulators), so one 256-thread block fits
    // per SM; blocks beyond that only queue behind it and each pays the gamma prologue and three block-reduction passes
    // (TAVK_LN_BWD_WAVES: measurement knob, blocks per SM)
    static const int waves = getenv("TAVK_LN_BWD_WAVES") ? atoi(getenv("TAVK_LN_BWD_WAVES")) : 1;
    const int cap = sm_count() * (waves > 0 ? waves : 1);
    if (grid > cap) grid = cap;
    const size_t smem = (size_t)warps * H * sizeof(float);
    __nv_bfloat16* db16 = reinterpret_cast<__nv_bfloat16*>(dx_bf16);
#define LN_BWD(V)                                                                                                  \
    case V:                                                                                                        \
        TAVK_CUDA(launch_kernel(layernorm_bwd_kernel<V>, dim3(grid), dim3(warps * 32), (size_t)(smem), stream, dy, x, mean, rstd, gamma, dx_resid, dx_f32, db16, \
                                                                   dgamma, dbeta, dx_colsum, M));                   \
        break;
    switch (H / 128) {
        LN_BWD(1) LN_BWD(2) LN_BWD(3) LN_BWD(4) LN_BWD(5) LN_BWD(6) LN_BWD(7) LN_BWD(8)
    }
#undef LN_BWD
    TAVK_CUDA(cudaGetLastError());
    return 0;
}
