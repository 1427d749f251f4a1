// Shared device/host helpers for the tavk kernel library (sm_100a only).
// PTX wrappers: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld), ldmatrix,
// mma.sync, cp.async.  No CUTLASS dependency: everything the kernels need is spelled out here.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define TAVK_DEVINL __device__ __forceinline__

namespace tavk {

// ---------------------------------------------------------------- error plumbing (host)
void set_error(const char* fmt, ...);
#define TAVK_CHECK(cond, code, ...)        \
    do {                                   \
        if (!(cond)) {                     \
            ::tavk::set_error(__VA_ARGS__); \
            return (code);                 \
        }                                  \
    } while (0)
#define TAVK_CUDA(expr)                                                              \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess) {                                                     \
            ::tavk::set_error("%s failed: %s", #expr, cudaGetErrorString(_e));       \
            return 3;                                                                \
        }                                                                            \
    } while (0)

int sm_count();
bool pdl_enabled();   // programmatic dependent launch: opt-in with TAVK_PDL=1 (see api.cu for the measurement)

// Launch with the programmatic-stream-serialization attribute: the grid may be scheduled while its predecessor on the
// stream is still draining, so its launch latency and its prologue (barrier init, TMEM allocation, descriptor
// prefetch) overlap the predecessor's tail.  Every kernel launched this way executes pdl_wait() before it touches
// global memory its predecessor may have written; the dependency is captured as a programmatic edge in CUDA graphs.
template <typename... KArgs, typename... Args>
cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                          Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Same for a kernel that runs as clusters of `cluster_x` CTAs along x (grid.x must be a multiple of it).
template <typename... KArgs, typename... Args>
cudaError_t launch_kernel_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                  unsigned cluster_x, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster_x;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------- programmatic dependent launch (device side)
// Blocks until the predecessor grid has completed and its memory is visible (no-op without the launch attribute).
TAVK_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Lets the successor grid start being scheduled (its blocks still stop at their own pdl_wait()).  Used ONLY by kernels
// whose whole grid is resident at once (the persistent one-CTA-per-SM GEMM): a multi-wave grid that triggers early lets
// waiting successor blocks take the SMs its own unscheduled blocks still need (observed as a hang).
TAVK_DEVINL void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------- small math
TAVK_DEVINL float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Column sums of a 32x32 block held as "lane = row, v[c] = column c": recursive halving, 31 shuffles; lane c returns
// the sum of column c (v is destroyed).
TAVK_DEVINL float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const bool upper = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < o; ++i) {
            const float send = upper ? v[i] : v[i + o];
            const float keep = upper ? v[i + o] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    return v[0];
}
TAVK_DEVINL float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// exact-erf GELU and its derivative (reference: nn.GELU() default, utils/TAVFormer.py:398,405)
TAVK_DEVINL float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
TAVK_DEVINL float gelu_erf_grad(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
    const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
    return cdf + x * pdf;
}

// ---------------------------------------------------------------- packed fp32x2 math (Blackwell FFMA2 / FMUL2 / FADD2)
// The GEMM epilogues are issue-slot bound (K = 768: ~24 slots per 32 outputs before the tensor pipe waits), so the
// polynomial work runs two outputs per instruction.
struct f32x2 {
    unsigned long long u;
};
TAVK_DEVINL f32x2 pk(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.u) : "f"(lo), "f"(hi));
    return r;
}
TAVK_DEVINL void unpk(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v.u)); }
TAVK_DEVINL f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.u) : "l"(a.u), "l"(b.u), "l"(c.u));
    return r;
}
TAVK_DEVINL f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.u) : "l"(a.u), "l"(b.u));
    return r;
}
TAVK_DEVINL f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.u) : "l"(a.u), "l"(b.u));
    return r;
}
TAVK_DEVINL float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
TAVK_DEVINL float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// erf via Abramowitz & Stegun 7.1.28: erf(z) = 1 - (1 + a1 z + ... + a6 z^6)^-16, |err| <= 3e-7 for z >= 0.  The
// coefficients below are pre-scaled by 2^(-k/2) so the polynomial takes |x| directly (z = |x|/sqrt 2).  Returns
// r = (...)^-16 = 1 - erf(|x|/sqrt 2) for two inputs; measured |gelu - exact| < 1e-6, |gelu' - exact| < 1e-6 over
// [-12, 12] in fp32 (three orders below the bf16 rounding of the stored result).
TAVK_DEVINL f32x2 erfc_abs2(f32x2 ax) {
    f32x2 p = fma2(pk(5.38297500e-06f, 5.38297500e-06f), ax, pk(4.88906356e-05f, 4.88906356e-05f));
    p = fma2(p, ax, pk(3.80035750e-05f, 3.80035750e-05f));
    p = fma2(p, ax, pk(3.27762632e-03f, 3.27762632e-03f));
    p = fma2(p, ax, pk(2.11410062e-02f, 2.11410062e-02f));
    p = fma2(p, ax, pk(4.98673470e-02f, 4.98673470e-02f));
    p = fma2(p, ax, pk(1.0f, 1.0f));
    p = mul2(p, p);
    p = mul2(p, p);
    p = mul2(p, p);
    p = mul2(p, p);
    float lo, hi;
    unpk(p, lo, hi);
    return pk(rcp_approx(lo), rcp_approx(hi));
}
// gelu(x) = 0.5 (x + |x| erf(|x|/sqrt 2)) = 0.5 (x + |x| - |x| r)
TAVK_DEVINL void gelu_fast2(float x0, float x1, float& g0, float& g1) {
    const f32x2 x = pk(x0, x1), ax = pk(fabsf(x0), fabsf(x1));
    const f32x2 r = erfc_abs2(ax);
    const f32x2 t = fma2(mul2(ax, r), pk(-1.0f, -1.0f), ax);
    const f32x2 h = pk(0.5f, 0.5f);
    unpk(fma2(x, h, mul2(t, h)), g0, g1);
}
// v *= gelu'(x) = Phi(x) + x phi(x).  Here the pdf's exp(-x^2/2) is needed anyway, so erfc comes from A&S 7.1.26
// (t = 1/(1 + p z), erfc(z) = t (a1 + t (a2 + t (a3 + t (a4 + t a5)))) exp(-z^2), |err| <= 1.5e-7) which shares it:
// Phi(x) = 0.5 + copysign(0.5 - 0.5 erfc(|x|/sqrt 2), x).  Measured |gelu' - exact| < 4e-7 over [-12, 12].
TAVK_DEVINL void gelu_grad_mul2(float x0, float x1, float& v0, float& v1) {
    const f32x2 x = pk(x0, x1), ax = pk(fabsf(x0), fabsf(x1));
    float d0, d1;
    unpk(fma2(ax, pk(2.316418883e-01f, 2.316418883e-01f), pk(1.0f, 1.0f)), d0, d1);
    const f32x2 t = pk(rcp_approx(d0), rcp_approx(d1));
    f32x2 q = fma2(t, pk(1.061405429f, 1.061405429f), pk(-1.453152027f, -1.453152027f));
    q = fma2(q, t, pk(1.421413741f, 1.421413741f));
    q = fma2(q, t, pk(-0.284496736f, -0.284496736f));
    q = fma2(q, t, pk(0.254829592f, 0.254829592f));
    q = mul2(q, t);
    float e0, e1;
    unpk(mul2(mul2(x, x), pk(-0.72134752f, -0.72134752f)), e0, e1);   // -x^2/2 * log2(e)
    const f32x2 e = pk(ex2_approx(e0), ex2_approx(e1));
    float h0, h1;
    unpk(fma2(mul2(q, e), pk(-0.5f, -0.5f), pk(0.5f, 0.5f)), h0, h1);
    const f32x2 cdf = add2(pk(copysignf(h0, x0), copysignf(h1, x1)), pk(0.5f, 0.5f));
    const f32x2 g = fma2(mul2(x, pk(0.3989422804f, 0.3989422804f)), e, cdf);
    unpk(mul2(pk(v0, v1), g), v0, v1);
}
// gelu(x) AND gelu'(x) from one evaluation of the A&S 7.1.26 form above: g = x Phi(x), d = Phi(x) + x phi(x).  The FFN-up
// forward stores d (bf16) instead of the pre-activation, which turns the backward's GELU' epilogue into one multiply.
TAVK_DEVINL void gelu_and_grad2(float x0, float x1, float& g0, float& g1, float& d0, float& d1) {
    const f32x2 x = pk(x0, x1), ax = pk(fabsf(x0), fabsf(x1));
    float t0, t1;
    unpk(fma2(ax, pk(2.316418883e-01f, 2.316418883e-01f), pk(1.0f, 1.0f)), t0, t1);
    const f32x2 t = pk(rcp_approx(t0), rcp_approx(t1));
    f32x2 q = fma2(t, pk(1.061405429f, 1.061405429f), pk(-1.453152027f, -1.453152027f));
    q = fma2(q, t, pk(1.421413741f, 1.421413741f));
    q = fma2(q, t, pk(-0.284496736f, -0.284496736f));
    q = fma2(q, t, pk(0.254829592f, 0.254829592f));
    q = mul2(q, t);
    float e0, e1;
    unpk(mul2(mul2(x, x), pk(-0.72134752f, -0.72134752f)), e0, e1);   // -x^2/2 * log2(e)
    const f32x2 e = pk(ex2_approx(e0), ex2_approx(e1));
    float h0, h1;
    unpk(fma2(mul2(q, e), pk(-0.5f, -0.5f), pk(0.5f, 0.5f)), h0, h1);
    const f32x2 cdf = add2(pk(copysignf(h0, x0), copysignf(h1, x1)), pk(0.5f, 0.5f));
    unpk(mul2(x, cdf), g0, g1);
    unpk(fma2(mul2(x, pk(0.3989422804f, 0.3989422804f)), e, cdf), d0, d1);
}
TAVK_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
TAVK_DEVINL float2 unpack_bf16x2(uint32_t u) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(v);
}

// ---------------------------------------------------------------- shared-memory address / elect
TAVK_DEVINL uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
TAVK_DEVINL bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "elect.sync _|P, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, P;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

TAVK_DEVINL void st_shared_v4(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
TAVK_DEVINL float4 ld_shared_v4(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
    return v;
}

// Read-once global data that another pointer of the same kernel may legally alias in the caller's eyes (residual vs
// output): the non-coherent path lets the compiler hoist these loads above earlier stores.
TAVK_DEVINL float4 ld_global_nc_v4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

TAVK_DEVINL uint2 ld_global_nc_v2(const void* p) {
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}

// ---------------------------------------------------------------- mbarrier
TAVK_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
TAVK_DEVINL void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
TAVK_DEVINL void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
TAVK_DEVINL void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
TAVK_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Non-blocking, warp-uniform test (for a warp that polls several barriers and serves whichever completes first).
TAVK_DEVINL bool mbar_test_warp(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return __all_sync(0xffffffffu, ok != 0);
}
// Bounded wait: a protocol bug traps (launch fails with an error) instead of hanging the GPU.  The report names the
// source line of the wait and the full block index (one line per warp).
TAVK_DEVINL void mbar_timeout_report(int line) {
    if ((threadIdx.x & 31) == 0)
        printf("tavk: mbarrier wait timed out (source line %d, block %d,%d,%d warp %d)\n", line, (int)blockIdx.x,
               (int)blockIdx.y, (int)blockIdx.z, (int)(threadIdx.x >> 5));
    __trap();
}
TAVK_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity, int line = __builtin_LINE()) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 22)) mbar_timeout_report(line);
    }
}

// Same, for the single-thread producer / issuer warps: they share schedulers with the compute warps, and a tight
// try_wait + branch loop was taking ~10x more issue slots than the warp's real work (ncu: 2.3 M executed spin
// instructions against 0.2 M per compute instruction in the attention forward) — sleep between polls instead.
template <int kSleepNs>
TAVK_DEVINL void mbar_wait_backoff(uint64_t* bar, uint32_t parity, int line = __builtin_LINE()) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(kSleepNs);
        if (++spins > (1u << 22)) mbar_timeout_report(line);
    }
}

// ---------------------------------------------------------------- CTA pair (cluster of two, tcgen05 cta_group::2)
TAVK_DEVINL uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// every thread of both CTAs
TAVK_DEVINL void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
TAVK_DEVINL uint32_t mapa_shared(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
TAVK_DEVINL void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait on a barrier of this CTA whose arrivals come from the peer CTA too
TAVK_DEVINL void mbar_wait_cluster(uint64_t* bar, uint32_t parity, int line = __builtin_LINE()) {
    uint32_t spins = 0, ok = 0;
    while (true) {
        asm volatile(
            "{\n"
            ".reg .pred P;\n"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n"
            "selp.u32 %0, 1, 0, P;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) break;
        if (++spins > (1u << 22)) mbar_timeout_report(line);
    }
}

// ---------------------------------------------------------------- TMA
TAVK_DEVINL void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
TAVK_DEVINL void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// Same from a CTA of a pair (tcgen05 cta_group::2): the bytes land in THIS CTA's shared memory, the complete_tx goes to the
// barrier at shared::cluster address `bar_cluster` — the leader CTA's full barrier, which its MMA issuer waits on.
TAVK_DEVINL void tma_load_2d_pair(void* smem_dst, const void* tmap, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster), "r"(c0), "r"(c1)
        : "memory");
}
TAVK_DEVINL void tma_load_3d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "r"(c2)
        : "memory");
}

// TMA store of one box from shared memory (bulk async-group completion: commit, then wait_group[.read]).
TAVK_DEVINL void tma_store_2d(const void* tmap, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(tmap)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
TAVK_DEVINL void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
TAVK_DEVINL void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
TAVK_DEVINL void tma_store_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
TAVK_DEVINL void fence_proxy_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------- tcgen05 / TMEM
TAVK_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
TAVK_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
template <uint32_t kCols>
TAVK_DEVINL void tmem_alloc(uint32_t* smem_result) {  // whole warp, .sync.aligned
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
TAVK_DEVINL void tmem_dealloc(uint32_t addr) {  // whole warp, same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(kCols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T ; one thread issues.
TAVK_DEVINL void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T : the A operand (M = 128 rows = TMEM lanes, K-major, two bf16 per 32-bit column, so
// one K = 16 step is 8 columns) is read from tensor memory instead of shared memory.
TAVK_DEVINL void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrives on the mbarrier once all previously issued tcgen05.mma of this thread have completed.
TAVK_DEVINL void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// ---- CTA-pair forms (cta_group::2): one warp of EACH CTA of the pair allocates / frees; only the leader CTA (cluster rank
// 0) issues MMAs — D is 256 rows (128 TMEM lanes in each CTA) x N columns, A = each CTA's own 128-row tile, B = N/2 rows
// from each CTA's shared memory at the same offset — and its commits arrive on the barrier at the same offset in both CTAs.
template <uint32_t kCols>
TAVK_DEVINL void tmem_alloc_pair(uint32_t* smem_result) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
TAVK_DEVINL void tmem_dealloc_pair(uint32_t addr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(kCols) : "memory");
}
TAVK_DEVINL void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
TAVK_DEVINL void umma_commit_pair(uint64_t* bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i = lane i of the warp's quarter).
TAVK_DEVINL void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
TAVK_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory operand descriptor (SWIZZLE_128B, Blackwell version field = 1).
//   bits [0,14)  start address >> 4      bits [16,30) leading-dim byte offset >> 4
//   bits [32,46) stride-dim byte offset >> 4   bits [46,48) version = 1   bits [61,64) layout = 2 (SW128)
TAVK_DEVINL uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Same for a K-major operand slice of 16 bf16 (32-byte rows) in the SWIZZLE_32B canonical layout: 8-row atoms of 256
// bytes, `sbo_bytes` between atoms (256 when packed), layout = 6 (SW32); the tile base must be 256-byte aligned.
TAVK_DEVINL uint64_t umma_smem_desc_sw32(uint32_t saddr, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)6 << 61;
    return d;
}
// UMMA instruction descriptor for kind::f16, bf16 x bf16 -> fp32.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
    return (1u << 4)                          // c_format = F32
           | (1u << 7)                        // a_format = BF16
           | (1u << 10)                       // b_format = BF16
           | ((a_mn_major ? 1u : 0u) << 15)   // a_major
           | ((b_mn_major ? 1u : 0u) << 16)   // b_major
           | ((uint32_t)(N >> 3) << 17)       // n_dim
           | ((uint32_t)(M >> 4) << 24);      // m_dim
}

// ---------------------------------------------------------------- legacy warp MMA path (attention v1)
TAVK_DEVINL void ldmatrix_x4(uint32_t (&r)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(saddr));
}
TAVK_DEVINL void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(saddr));
}
// D(16x8,f32) += A(16x16,bf16) * B(16x8,bf16)
TAVK_DEVINL void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
TAVK_DEVINL void cp_async_16(uint32_t saddr, const void* gptr, bool pred) {
    const int sz = pred ? 16 : 0;  // src-size 0 => zero-fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(gptr), "r"(sz) : "memory");
}
TAVK_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
TAVK_DEVINL void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

}  // namespace tavk
