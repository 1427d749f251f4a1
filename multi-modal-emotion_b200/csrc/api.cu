// Library-level entry points of libtavk.so: error string, version, device probe.
// There is no CPU fallback anywhere in this library: tavk_device_check() is what the Python host calls at import
// time on a GPU box, and every compute entry point enqueues sm_100a-only kernels.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/tavk.h"
#include "common.cuh"

namespace tavk {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        // Off unless TAVK_PDL=1.  Measured on the full step (r1): with the early trigger only in the single-wave GEMM the
        // step time is unchanged (46.9 vs 46.6 ms) — GEMM->GEMM chains, where it saves 0.9 us per launch, are rare — and
        // an early trigger in multi-wave kernels hung (see pdl_trigger in common.cuh).
        const char* e = getenv("TAVK_PDL");
        v = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    return v != 0;
}

}  // namespace tavk

extern "C" const char* tavk_last_error(void) { return tavk::g_err; }

extern "C" int tavk_version(void) { return TAVK_VERSION; }

extern "C" int tavk_sm_count(void) { return tavk::sm_count(); }

// Workspace sizes of the entry points that need caller-owned scratch (everything else needs none).
extern "C" int64_t tavk_workspace_bytes_attn_bwd(int B, int S, int nh) {
    if (B <= 0 || S <= 0 || nh <= 0) return 0;
    return (int64_t)B * nh * S * (int64_t)sizeof(float);            // tavk_attn_bwd_args.delta: f32 [B, nh, S]
}
extern "C" int64_t tavk_workspace_bytes_groupnorm(int B, int C) {
    if (B <= 0 || C <= 0) return 0;
    return (int64_t)B * 2 * C * (int64_t)sizeof(float);             // sums_ws: f32 [B, 2, C]
}
extern "C" int64_t tavk_workspace_bytes_gemm(const tavk_gemm_args* a) {
    (void)a;
    return 0;   // accumulators live in tensor memory, split-K partial sums go straight to the output (red.global.add)
}

extern "C" int tavk_device_check(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        tavk::set_error("tavk_device_check: no CUDA device (%s)", cudaGetErrorString(e));
        return 3;
    }
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (major != 10) {
        tavk::set_error("tavk_device_check: device %d is sm_%d%d; this library is built for sm_100a only", dev, major,
                        minor);
        return 3;
    }
    return 0;
}
