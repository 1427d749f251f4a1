// Weighted softmax cross-entropy (reference utils/global_functions.py:63-64,76,83) and the optimiser step
// (reference train_model/tav_train.py:61-62: clip_grad_norm_ followed by torch.optim.AdamW.step) as HBM-bound
// kernels over flat fp32 buffers.
#include "../../include/tavk.h"
#include "common.cuh"

namespace tavk {

// ---------------------------------------------------------------- softmax CE
// One block; thread per row (B is a per-GPU batch, <= a few thousand).  Deterministic block reduction.
__global__ void __launch_bounds__(256)
softmax_ce_fwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target,
                      const float* __restrict__ cw, float* __restrict__ probs, float* __restrict__ loss_num,
                      float* __restrict__ loss_den, int B, int C) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    __shared__ float s_num[8], s_den[8];
    float num = 0.f, den = 0.f;
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        const float* z = logits + (size_t)i * C;
        float mx = -INFINITY;
        for (int c = 0; c < C; ++c) mx = fmaxf(mx, z[c]);
        float se = 0.f;
        for (int c = 0; c < C; ++c) se += expf(z[c] - mx);
        const float lse = mx + logf(se);
        const int y = (int)target[i];
        const bool valid = (y >= 0 && y < C);
        const float w = valid ? (cw ? cw[y] : 1.0f) : 0.f;
        if (probs)
            for (int c = 0; c < C; ++c) probs[(size_t)i * C + c] = expf(z[c] - lse);
        if (valid) {
            num += w * (lse - z[y]);
            den += w;
        }
    }
    num = warp_sum(num);
    den = warp_sum(den);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_num[warp] = num; s_den[warp] = den; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += s_num[w]; b += s_den[w]; }
        *loss_num = a;
        *loss_den = b;
    }
}

__global__ void softmax_ce_bwd_kernel(const float* __restrict__ probs, const int64_t* __restrict__ target,
                                      const float* __restrict__ cw, const float* __restrict__ gscale,
                                      float* __restrict__ dlogits, int B, int C) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * C) return;
    const int r = i / C, c = i - r * C;
    const int y = (int)target[r];
    const bool valid = (y >= 0 && y < C);
    const float w = valid ? (cw ? cw[y] : 1.0f) : 0.f;
    dlogits[i] = (*gscale) * w * (probs[i] - (c == y ? 1.0f : 0.f));
}

// ---------------------------------------------------------------- grad norm
__global__ void __launch_bounds__(256)
grad_sqnorm_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    __shared__ float s_part[8];
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    float acc = 0.f;
    for (long long i = t; i < n4; i += stride) {
        const float4 v = reinterpret_cast<const float4*>(g)[i];
        acc += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
    for (long long i = (n4 << 2) + t; i < n; i += stride) acc += g[i] * g[i];
    acc = warp_sum(acc);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) s_part[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f;
        for (int w = 0; w < 8; ++w) a += s_part[w];
        atomicAdd(out, a);
    }
}

// ---------------------------------------------------------------- AdamW
// Same update order as torch.optim.AdamW (single-tensor path): decay, moments, bias-corrected step.
struct AdamArgs {
    float lr, beta1, beta2, omb1, omb2, eps, wd, bc1, bc2_sqrt, max_norm, prescale;   // omb = 1 - beta, rounded from double
    int zero_grad;
};

TAVK_DEVINL float adam_one(float& p, float& m, float& v, float g, const AdamArgs& a) {
    p *= (1.0f - a.lr * a.wd);
    m = a.beta1 * m + a.omb1 * g;
    v = a.beta2 * v + a.omb2 * g * g;
    const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
    p -= (a.lr / a.bc1) * (m / denom);
    return p;
}

// Device-resident optimiser clock (tavk_adamw_prep): the step counter, the learning rate and the two bias corrections
// live in device memory so that a captured CUDA graph advances them on every replay (by-value kernel arguments would
// freeze the capture-time step and learning rate).  hyper = {lr, 1 - beta1^step, sqrt(1 - beta2^step), unused}.
__global__ void adamw_prep_kernel(int* __restrict__ step, float* __restrict__ hyper, float* __restrict__ sqnorm,
                                  double beta1, double beta2) {
    pdl_wait();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const int s = *step + 1;
        *step = s;
        hyper[1] = (float)(1.0 - pow(beta1, (double)s));
        hyper[2] = (float)sqrt(1.0 - pow(beta2, (double)s));
        if (sqnorm != nullptr) *sqnorm = 0.f;
    }
}

__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, float* __restrict__ g,
             __nv_bfloat16* __restrict__ p_bf16, long long n, const float* __restrict__ sqnorm,
             const float* __restrict__ hyper, AdamArgs a) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    if (hyper != nullptr) {
        a.lr = hyper[0];
        a.bc1 = hyper[1];
        a.bc2_sqrt = hyper[2];
    }
    float gs = a.prescale;
    if (sqnorm != nullptr && a.max_norm > 0.f) {
        // clip_grad_norm_: the norm is taken over the (pre-scaled) gradient
        const float norm = sqrtf(*sqnorm) * fabsf(a.prescale);
        const float clip = a.max_norm / (norm + 1e-6f);
        gs *= fminf(clip, 1.0f);
    }
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    // Two float4 per array per thread and iteration, all eight loads issued before the first use; m, v and g are touched
    // once per step (streaming loads / stores keep them from displacing the parameters and the bf16 mirror, which the next
    // forward reads first, in L2).
    auto ld = [](const float* q, long long i) { return __ldcs(reinterpret_cast<const float4*>(q) + i); };
    auto update = [&](long long i, float4 pv, float4 mv, float4 vv, const float4 gv) {
        adam_one(pv.x, mv.x, vv.x, gv.x * gs, a);
        adam_one(pv.y, mv.y, vv.y, gv.y * gs, a);
        adam_one(pv.z, mv.z, vv.z, gv.z * gs, a);
        adam_one(pv.w, mv.w, vv.w, gv.w * gs, a);
        reinterpret_cast<float4*>(p)[i] = pv;
        __stcs(reinterpret_cast<float4*>(m) + i, mv);
        __stcs(reinterpret_cast<float4*>(v) + i, vv);
        if (p_bf16) reinterpret_cast<uint2*>(p_bf16)[i] = make_uint2(pack_bf16x2(pv.x, pv.y), pack_bf16x2(pv.z, pv.w));
        if (a.zero_grad) __stcs(reinterpret_cast<float4*>(g) + i, make_float4(0.f, 0.f, 0.f, 0.f));
    };
    long long i = t;
    for (; i + stride < n4; i += 2 * stride) {
        const long long j = i + stride;
        const float4 p0 = reinterpret_cast<float4*>(p)[i], p1 = reinterpret_cast<float4*>(p)[j];
        const float4 m0 = ld(m, i), m1 = ld(m, j), v0 = ld(v, i), v1 = ld(v, j), g0 = ld(g, i), g1 = ld(g, j);
        update(i, p0, m0, v0, g0);
        update(j, p1, m1, v1, g1);
    }
    if (i < n4) update(i, reinterpret_cast<float4*>(p)[i], ld(m, i), ld(v, i), ld(g, i));
    for (long long i = (n4 << 2) + t; i < n; i += stride) {
        float pv = p[i], mv = m[i], vv = v[i];
        adam_one(pv, mv, vv, g[i] * gs, a);
        p[i] = pv; m[i] = mv; v[i] = vv;
        if (p_bf16) p_bf16[i] = __float2bfloat16_rn(pv);
        if (a.zero_grad) g[i] = 0.f;
    }
}

}  // namespace tavk

using namespace tavk;
#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int tavk_softmax_ce_fwd(const float* logits, const int64_t* target, const float* class_weight, float* probs,
                                   float* loss_num, float* loss_den, int B, int C, void* stream) {
    TAVK_CHECK(logits && target && loss_num && loss_den, 1, "tavk_softmax_ce_fwd: null pointer");
    TAVK_CHECK(B >= 0 && C >= 1, 1, "tavk_softmax_ce_fwd: bad shape B=%d C=%d", B, C);
    TAVK_CUDA(launch_kernel(softmax_ce_fwd_kernel, dim3(1), dim3(256), (size_t)(0), STREAM(stream), logits, target, class_weight, probs, loss_num, loss_den, B, C));
    TAVK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tavk_softmax_ce_bwd(const float* probs, const int64_t* target, const float* class_weight,
                                   const float* gscale_dev, float* dlogits, int B, int C, void* stream) {
    TAVK_CHECK(probs && target && gscale_dev && dlogits, 1, "tavk_softmax_ce_bwd: null pointer");
    if (B <= 0) return 0;
    const int total = B * C;
    TAVK_CUDA(launch_kernel(softmax_ce_bwd_kernel, dim3((total + 127) / 128), dim3(128), (size_t)(0), STREAM(stream), probs, target, class_weight, gscale_dev,
                                                                          dlogits, B, C));
    TAVK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tavk_grad_sqnorm(const float* g, int64_t n, float* out, void* stream) {
    TAVK_CHECK(g && out, 1, "tavk_grad_sqnorm: null pointer");
    TAVK_CHECK((reinterpret_cast<uintptr_t>(g) & 15) == 0, 1, "tavk_grad_sqnorm: g must be 16-byte aligned");
    if (n <= 0) return 0;
    long long blocks = ((n + 3) / 4 + 255) / 256;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    TAVK_CUDA(launch_kernel(grad_sqnorm_kernel, dim3((int)blocks), dim3(256), (size_t)(0), STREAM(stream), g, n, out));
    TAVK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tavk_adamw(float* p, float* m, float* v, float* g, void* p_bf16, int64_t n, float lr, float beta1,
                          float beta2, float eps, float weight_decay, int step, const float* sqnorm_dev, float max_norm,
                          float grad_prescale, int zero_grad, void* stream) {
    TAVK_CHECK(p && m && v && g, 1, "tavk_adamw: null pointer");
    TAVK_CHECK(step >= 1, 1, "tavk_adamw: step must be >= 1 (got %d)", step);
    TAVK_CHECK(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v) |
                 reinterpret_cast<uintptr_t>(g)) & 15) == 0,
               1, "tavk_adamw: buffers must be 16-byte aligned");
    TAVK_CHECK(p_bf16 == nullptr || (reinterpret_cast<uintptr_t>(p_bf16) & 7) == 0, 1,
               "tavk_adamw: bf16 shadow must be 8-byte aligned");
    if (n <= 0) return 0;
    AdamArgs a;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.wd = weight_decay;
    a.omb1 = 1.0f - beta1; a.omb2 = 1.0f - beta2;
    a.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
    a.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    a.max_norm = max_norm; a.prescale = grad_prescale; a.zero_grad = zero_grad;
    long long blocks = ((n + 3) / 4 + 255) / 256;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    TAVK_CUDA(launch_kernel(adamw_kernel, dim3((int)blocks), dim3(256), (size_t)(0), STREAM(stream), p, m, v, g, reinterpret_cast<__nv_bfloat16*>(p_bf16), n,
                                                          sqnorm_dev, (const float*)nullptr, a));
    TAVK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tavk_adamw_prep(int* step_dev, float* hyper_dev, float* sqnorm_dev, double beta1, double beta2, void* stream) {
    TAVK_CHECK(step_dev && hyper_dev, 1, "tavk_adamw_prep: null pointer");
    TAVK_CHECK(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f, 1, "tavk_adamw_prep: betas must be in [0,1)");
    TAVK_CUDA(launch_kernel(adamw_prep_kernel, dim3(1), dim3(32), (size_t)(0), STREAM(stream), step_dev, hyper_dev, sqnorm_dev, beta1, beta2));
    TAVK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tavk_adamw_dev(float* p, float* m, float* v, float* g, void* p_bf16, int64_t n, const float* hyper_dev,
                              double beta1, double beta2, float eps, float weight_decay, const float* sqnorm_dev,
                              float max_norm, float grad_prescale, int zero_grad, void* stream) {
    TAVK_CHECK(p && m && v && g && hyper_dev, 1, "tavk_adamw_dev: null pointer");
    TAVK_CHECK(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v) |
                 reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(hyper_dev)) & 15) == 0,
               1, "tavk_adamw_dev: buffers must be 16-byte aligned");
    TAVK_CHECK(p_bf16 == nullptr || (reinterpret_cast<uintptr_t>(p_bf16) & 7) == 0, 1,
               "tavk_adamw_dev: bf16 shadow must be 8-byte aligned");
    if (n <= 0) return 0;
    AdamArgs a;
    // betas arrive as doubles so that 1 - beta is rounded once, like torch's Python-scalar arithmetic (1 - 0.999 taken in
    // float is 1.3e-5 off in relative terms, which shows in exp_avg_sq)
    a.lr = 0.f; a.beta1 = (float)beta1; a.beta2 = (float)beta2; a.eps = eps; a.wd = weight_decay;
    a.omb1 = (float)(1.0 - beta1); a.omb2 = (float)(1.0 - beta2);
    a.bc1 = 1.f; a.bc2_sqrt = 1.f;      // replaced by hyper_dev[0..2] on the device
    a.max_norm = max_norm; a.prescale = grad_prescale; a.zero_grad = zero_grad;
    long long blocks = ((n + 3) / 4 + 255) / 256;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    TAVK_CUDA(launch_kernel(adamw_kernel, dim3((int)blocks), dim3(256), (size_t)(0), STREAM(stream), p, m, v, g, reinterpret_cast<__nv_bfloat16*>(p_bf16), n,
                                                          sqnorm_dev, hyper_dev, a));
    TAVK_CUDA(cudaGetLastError());
    return 0;
}
