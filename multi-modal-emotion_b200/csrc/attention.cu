// Fused flash-style attention for head_dim 64: forward, and backward as two kernels (dQ; dK+dV) that recompute the
// probabilities from the saved log-sum-exp, so no [S,S] tensor ever reaches HBM.
//
// Replaces (reference): utils/TAVFormer.py:357-387 (VideoMAESelfAttention: QK^T/8, softmax, PV),
// utils/TAVFormer.py:57-86 (MultiHeadAttention with the pre-softmax additive mask) and the eager/SDPA attention of
// the HF RoBERTa / Wav2Vec2 / VideoMAE layers.  The fusion encoder's POST-softmax mask add (utils/TAVFormer.py:372-375)
// is a rank-1 term outside the softmax; it is applied by the caller through tavk_masked_colsum + the out-projection
// row-bias, and only its dV contribution (dv_rowscale x dv_rank1) is folded into the dK/dV kernel epilogue here.
//
// v1 data path: 64x64 tiles, 4 warps per CTA, bf16 mma.sync m16n8k16 with fp32 accumulators held in registers, online
// softmax in registers with quad shuffles, K/V (or Q/dO) tiles double-buffered in XOR-swizzled shared memory via
// cp.async.  Ragged S is handled by zero-filled loads + -inf key masking + predicated stores.
#include <stdlib.h>
#include <string.h>

#include "../../include/tavk.h"
#include "common.cuh"

namespace tavk {

constexpr int kTile = 64;        // queries per CTA, keys per inner tile
constexpr int kHeadDim = 64;
constexpr int kTileBytes = kTile * kHeadDim * 2;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// 64x64 bf16 tile, 128-byte rows, 16-byte chunks XOR-swizzled with the row index (bank-conflict-free ldmatrix)
TAVK_DEVINL uint32_t tile_addr(uint32_t base, int row, int chunk) {
    return base + row * 128 + ((chunk ^ (row & 7)) << 4);
}

// global [rows, ld] (head slice already applied to gptr) -> smem tile; rows >= rows_valid are zero-filled
TAVK_DEVINL void load_tile_async(uint32_t smem_base, const __nv_bfloat16* gptr, long long ld, int row0, int rows_total) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int id = threadIdx.x + 128 * i;
        const int r = id >> 3, c = id & 7;
        const bool ok = (row0 + r) < rows_total;
        const __nv_bfloat16* src = gptr + (long long)(ok ? row0 + r : 0) * ld + c * 8;
        cp_async_16(tile_addr(smem_base, r, c), src, ok);
    }
}

// A fragments (16 rows x 64 k) of rows [row0, row0+16) -> 4 k-steps x 4 regs
TAVK_DEVINL void load_a_frags(uint32_t (&f)[4][4], uint32_t smem_base, int row0, int lane) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) ldmatrix_x4(f[kk], tile_addr(smem_base, row0 + (lane & 15), 2 * kk + (lane >> 4)));
}

// acc[16 x 64] += A(16 x 64, regs) * T^T where the smem tile T is stored [n][k]  (n = 64 tile rows, k = 64 cols)
TAVK_DEVINL void mma_a_tileT(float (&acc)[8][4], const uint32_t (&a)[4][4], uint32_t tile, int lane) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {
            uint32_t b[4];
            ldmatrix_x4(b, tile_addr(tile, np * 16 + (lane & 7) + ((lane >> 4) << 3), 2 * kk + ((lane >> 3) & 1)));
            mma_bf16_16816(acc[2 * np], a[kk], b[0], b[1]);
            mma_bf16_16816(acc[2 * np + 1], a[kk], b[2], b[3]);
        }
    }
}

// acc[16 x 64] += P(16 x 64, fp32 regs converted to bf16 A fragments) * T where the smem tile T is stored [k][n]
TAVK_DEVINL void mma_p_tile(float (&acc)[8][4], const float (&p)[8][4], uint32_t tile, int lane) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        uint32_t a[4];
        a[0] = pack_bf16x2(p[2 * kk][0], p[2 * kk][1]);
        a[1] = pack_bf16x2(p[2 * kk][2], p[2 * kk][3]);
        a[2] = pack_bf16x2(p[2 * kk + 1][0], p[2 * kk + 1][1]);
        a[3] = pack_bf16x2(p[2 * kk + 1][2], p[2 * kk + 1][3]);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
            uint32_t b[4];
            ldmatrix_x4_trans(b, tile_addr(tile, kk * 16 + (lane & 15), 2 * np + (lane >> 4)));
            mma_bf16_16816(acc[2 * np], a, b[0], b[1]);
            mma_bf16_16816(acc[2 * np + 1], a, b[2], b[3]);
        }
    }
}

// Each warp stages its 16 x 64 fp32 accumulator block as bf16 into its own 16 rows of `tile`, then stores it with
// coalesced 16-byte writes to global rows [row0 + warp*16, ...) (predicated on rows_total).
TAVK_DEVINL void store_acc_tile(const float (&acc)[8][4], uint32_t tile, __nv_bfloat16* gptr, long long ld, int row0,
                                int rows_total, int warp, int lane) {
    const int r = warp * 16 + (lane >> 2);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t lo = pack_bf16x2(acc[j][0], acc[j][1]);
        const uint32_t hi = pack_bf16x2(acc[j][2], acc[j][3]);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(tile_addr(tile, r, j) + (lane & 3) * 4), "r"(lo) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(tile_addr(tile, r + 8, j) + (lane & 3) * 4), "r"(hi) : "memory");
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int id = lane + 32 * i;
        const int rr = warp * 16 + (id >> 3), c = id & 7;
        uint4 v;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "r"(tile_addr(tile, rr, c)));
        if (row0 + rr < rows_total) *reinterpret_cast<uint4*>(gptr + (long long)(row0 + rr) * ld + c * 8) = v;
    }
}

TAVK_DEVINL float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
TAVK_DEVINL float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// ---------------------------------------------------------------- forward
struct AttnFwdDev {
    const __nv_bfloat16 *q, *k, *v;
    long long ld_qkv;
    __nv_bfloat16* o;
    long long ld_o;
    float* lse;
    const float* key_bias;
    int B, S, nh;
    float scale_log2;
};

__global__ void __launch_bounds__(128) attn_fwd_kernel(const AttnFwdDev p) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    __shared__ __align__(128) uint8_t smem[5 * kTileBytes];  // Q | K0 K1 | V0 V1
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * kTile, h = blockIdx.y, b = blockIdx.z;
    const uint32_t sQ = smem_u32(smem), sK = sQ + kTileBytes, sV = sQ + 3 * kTileBytes;
    const long long head_off = (long long)b * p.S * p.ld_qkv + h * kHeadDim;
    const __nv_bfloat16* gq = p.q + head_off;
    const __nv_bfloat16* gk = p.k + head_off;
    const __nv_bfloat16* gv = p.v + head_off;
    const float* kb = p.key_bias ? p.key_bias + (long long)b * p.S : nullptr;
    const int nkt = (p.S + kTile - 1) / kTile;

    load_tile_async(sQ, gq, p.ld_qkv, q0, p.S);
    load_tile_async(sK, gk, p.ld_qkv, 0, p.S);
    load_tile_async(sV, gv, p.ld_qkv, 0, p.S);
    cp_async_commit();

    uint32_t qf[4][4];
    float o[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

    for (int kt = 0; kt < nkt; ++kt) {
        cp_async_wait<0>();
        __syncthreads();
        if (kt == 0) load_a_frags(qf, sQ, warp * 16, lane);
        if (kt + 1 < nkt) {
            load_tile_async(sK + ((kt + 1) & 1) * kTileBytes, gk, p.ld_qkv, (kt + 1) * kTile, p.S);
            load_tile_async(sV + ((kt + 1) & 1) * kTileBytes, gv, p.ld_qkv, (kt + 1) * kTile, p.S);
            cp_async_commit();
        }
        const uint32_t tK = sK + (kt & 1) * kTileBytes, tV = sV + (kt & 1) * kTileBytes;
        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
        mma_a_tileT(s, qf, tK, lane);
        // scale, additive key bias, key validity
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int key = kt * kTile + j * 8 + (lane & 3) * 2;
            const bool v0 = key < p.S, v1 = key + 1 < p.S;
            const float b0 = (kb && v0) ? kb[key] * kLog2e : 0.f;
            const float b1 = (kb && v1) ? kb[key + 1] * kLog2e : 0.f;
            s[j][0] = v0 ? fmaf(s[j][0], p.scale_log2, b0) : -INFINITY;
            s[j][1] = v1 ? fmaf(s[j][1], p.scale_log2, b1) : -INFINITY;
            s[j][2] = v0 ? fmaf(s[j][2], p.scale_log2, b0) : -INFINITY;
            s[j][3] = v1 ? fmaf(s[j][3], p.scale_log2, b1) : -INFINITY;
            mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
            mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
        }
        mx0 = quad_max(mx0);
        mx1 = quad_max(mx1);
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
        const float ms0 = (mn0 == -INFINITY) ? 0.f : mn0, ms1 = (mn1 == -INFINITY) ? 0.f : mn1;
        const float c0 = exp2f(m0 - ms0), c1 = exp2f(m1 - ms1);  // m = -inf -> 0
        m0 = mn0; m1 = mn1;
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s[j][0] = exp2f(s[j][0] - ms0);
            s[j][1] = exp2f(s[j][1] - ms0);
            s[j][2] = exp2f(s[j][2] - ms1);
            s[j][3] = exp2f(s[j][3] - ms1);
            rs0 += s[j][0] + s[j][1];
            rs1 += s[j][2] + s[j][3];
            o[j][0] *= c0; o[j][1] *= c0; o[j][2] *= c1; o[j][3] *= c1;
        }
        l0 = l0 * c0 + rs0;
        l1 = l1 * c1 + rs1;
        mma_p_tile(o, s, tV, lane);
    }
    l0 = quad_sum(l0);
    l1 = quad_sum(l1);
    const float i0 = l0 > 0.f ? 1.0f / l0 : 0.f, i1 = l1 > 0.f ? 1.0f / l1 : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { o[j][0] *= i0; o[j][1] *= i0; o[j][2] *= i1; o[j][3] *= i1; }
    if ((lane & 3) == 0 && p.lse) {
        const int r0 = q0 + warp * 16 + (lane >> 2);
        float* lse = p.lse + ((long long)b * p.nh + h) * p.S;
        if (r0 < p.S) lse[r0] = l0 > 0.f ? m0 * kLn2 + logf(l0) : -INFINITY;
        if (r0 + 8 < p.S) lse[r0 + 8] = l1 > 0.f ? m1 * kLn2 + logf(l1) : -INFINITY;
    }
    // the Q tile rows of this warp are dead after the fragment load: reuse them to coalesce the O store
    store_acc_tile(o, sQ, p.o + (long long)b * p.S * p.ld_o + h * kHeadDim, p.ld_o, q0, p.S, warp, lane);
}

// ---------------------------------------------------------------- backward: delta = rowsum(dO * O)
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o,
                                  long long ld_o, float* __restrict__ delta, int B, int S, int nh) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    // 8 lanes per (row, head): 8 x 16-byte loads cover the 64-wide head slice
    const long long gid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 3;
    const int sub = threadIdx.x & 7;
    const long long total = (long long)B * S * nh;
    float acc = 0.f;
    long long row = 0;
    int h = 0;
    const bool ok = gid < total;
    if (ok) {
        row = gid / nh;
        h = (int)(gid - row * nh);
        const uint4 a = *reinterpret_cast<const uint4*>(o + row * ld_o + h * kHeadDim + sub * 8);
        const uint4 g = *reinterpret_cast<const uint4*>(d_o + row * ld_o + h * kHeadDim + sub * 8);
        const uint32_t av[4] = {a.x, a.y, a.z, a.w}, gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 x = unpack_bf16x2(av[i]), y = unpack_bf16x2(gv[i]);
            acc += x.x * y.x + x.y * y.y;
        }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (ok && sub == 0) {
        const long long bb = row / S, s = row - bb * S;
        delta[(bb * nh + h) * S + s] = acc;
    }
}

struct AttnBwdDev {
    const __nv_bfloat16 *q, *k, *v;
    long long ld_qkv;
    const __nv_bfloat16* d_o;
    long long ld_o;
    const float *lse, *delta, *key_bias;
    __nv_bfloat16 *dq, *dk, *dv;
    long long ld_dqkv;
    const float *dv_rowscale, *dv_rank1;
    int B, S, nh;
    float scale, scale_log2;
};

// ---------------------------------------------------------------- backward: dQ (CTA = 64 queries, loops over keys)
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(const AttnBwdDev p) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    extern __shared__ __align__(128) uint8_t smem[];  // Q | dO | K0 K1 | V0 V1
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * kTile, h = blockIdx.y, b = blockIdx.z;
    const uint32_t sQ = smem_u32(smem), sDO = sQ + kTileBytes, sK = sQ + 2 * kTileBytes, sV = sQ + 4 * kTileBytes;
    const long long head_off = (long long)b * p.S * p.ld_qkv + h * kHeadDim;
    const __nv_bfloat16* gk = p.k + head_off;
    const __nv_bfloat16* gv = p.v + head_off;
    const float* kb = p.key_bias ? p.key_bias + (long long)b * p.S : nullptr;
    const int nkt = (p.S + kTile - 1) / kTile;

    load_tile_async(sQ, p.q + head_off, p.ld_qkv, q0, p.S);
    load_tile_async(sDO, p.d_o + (long long)b * p.S * p.ld_o + h * kHeadDim, p.ld_o, q0, p.S);
    load_tile_async(sK, gk, p.ld_qkv, 0, p.S);
    load_tile_async(sV, gv, p.ld_qkv, 0, p.S);
    cp_async_commit();

    const int r0 = q0 + warp * 16 + (lane >> 2);
    const long long stat_off = ((long long)b * p.nh + h) * p.S;
    const float lse0 = r0 < p.S ? p.lse[stat_off + r0] * kLog2e : 0.f;
    const float lse1 = r0 + 8 < p.S ? p.lse[stat_off + r0 + 8] * kLog2e : 0.f;
    const float dl0 = r0 < p.S ? p.delta[stat_off + r0] : 0.f;
    const float dl1 = r0 + 8 < p.S ? p.delta[stat_off + r0 + 8] : 0.f;

    uint32_t qf[4][4], dof[4][4];
    float dq[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f;

    for (int kt = 0; kt < nkt; ++kt) {
        cp_async_wait<0>();
        __syncthreads();
        if (kt == 0) {
            load_a_frags(qf, sQ, warp * 16, lane);
            load_a_frags(dof, sDO, warp * 16, lane);
        }
        if (kt + 1 < nkt) {
            load_tile_async(sK + ((kt + 1) & 1) * kTileBytes, gk, p.ld_qkv, (kt + 1) * kTile, p.S);
            load_tile_async(sV + ((kt + 1) & 1) * kTileBytes, gv, p.ld_qkv, (kt + 1) * kTile, p.S);
            cp_async_commit();
        }
        const uint32_t tK = sK + (kt & 1) * kTileBytes, tV = sV + (kt & 1) * kTileBytes;
        float s[8][4], dp[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
            dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
        }
        mma_a_tileT(s, qf, tK, lane);
        mma_a_tileT(dp, dof, tV, lane);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int key = kt * kTile + j * 8 + (lane & 3) * 2;
            const bool v0 = key < p.S, v1 = key + 1 < p.S;
            const float b0 = (kb && v0) ? kb[key] * kLog2e : 0.f;
            const float b1 = (kb && v1) ? kb[key + 1] * kLog2e : 0.f;
            const float p00 = v0 ? exp2f(fmaf(s[j][0], p.scale_log2, b0) - lse0) : 0.f;
            const float p01 = v1 ? exp2f(fmaf(s[j][1], p.scale_log2, b1) - lse0) : 0.f;
            const float p10 = v0 ? exp2f(fmaf(s[j][2], p.scale_log2, b0) - lse1) : 0.f;
            const float p11 = v1 ? exp2f(fmaf(s[j][3], p.scale_log2, b1) - lse1) : 0.f;
            s[j][0] = p00 * (dp[j][0] - dl0);
            s[j][1] = p01 * (dp[j][1] - dl0);
            s[j][2] = p10 * (dp[j][2] - dl1);
            s[j][3] = p11 * (dp[j][3] - dl1);
        }
        mma_p_tile(dq, s, tK, lane);  // dQ += dS * K   (K tile stored [key][d] = [k][n])
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { dq[j][0] *= p.scale; dq[j][1] *= p.scale; dq[j][2] *= p.scale; dq[j][3] *= p.scale; }
    store_acc_tile(dq, sQ, p.dq + (long long)b * p.S * p.ld_dqkv + h * kHeadDim, p.ld_dqkv, q0, p.S, warp, lane);
}

// ---------------------------------------------------------------- backward: dK, dV (CTA = 64 keys, loops over queries)
__global__ void __launch_bounds__(128) attn_bwd_dkv_kernel(const AttnBwdDev p) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    extern __shared__ __align__(128) uint8_t smem[];  // K | V | Q0 Q1 | dO0 dO1 | lse[2][64] delta[2][64]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k0 = blockIdx.x * kTile, h = blockIdx.y, b = blockIdx.z;
    const uint32_t sK = smem_u32(smem), sV = sK + kTileBytes, sQ = sK + 2 * kTileBytes, sDO = sK + 4 * kTileBytes;
    float* s_lse = reinterpret_cast<float*>(smem + 6 * kTileBytes);  // [2][64]
    float* s_del = s_lse + 2 * kTile;                                // [2][64]
    const long long head_off = (long long)b * p.S * p.ld_qkv + h * kHeadDim;
    const __nv_bfloat16* gq = p.q + head_off;
    const __nv_bfloat16* gdo = p.d_o + (long long)b * p.S * p.ld_o + h * kHeadDim;
    const long long stat_off = ((long long)b * p.nh + h) * p.S;
    const int nqt = (p.S + kTile - 1) / kTile;

    auto load_stats = [&](int qt, int buf) {
        const int t = threadIdx.x;
        if (t < kTile) {
            const int r = qt * kTile + t;
            s_lse[buf * kTile + t] = r < p.S ? p.lse[stat_off + r] * kLog2e : 0.f;
        } else {
            const int r = qt * kTile + t - kTile;
            s_del[buf * kTile + t - kTile] = r < p.S ? p.delta[stat_off + r] : 0.f;
        }
    };

    load_tile_async(sK, p.k + head_off, p.ld_qkv, k0, p.S);
    load_tile_async(sV, p.v + head_off, p.ld_qkv, k0, p.S);
    load_tile_async(sQ, gq, p.ld_qkv, 0, p.S);
    load_tile_async(sDO, gdo, p.ld_o, 0, p.S);
    cp_async_commit();
    load_stats(0, 0);

    // this thread's two key rows
    const int kr0 = k0 + warp * 16 + (lane >> 2), kr1 = kr0 + 8;
    const bool kv0 = kr0 < p.S, kv1 = kr1 < p.S;
    const float* kb = p.key_bias ? p.key_bias + (long long)b * p.S : nullptr;
    const float kb0 = (kb && kv0) ? kb[kr0] * kLog2e : 0.f;
    const float kb1 = (kb && kv1) ? kb[kr1] * kLog2e : 0.f;

    uint32_t kf[4][4], vf[4][4];
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = 0.f;
        dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.f;
    }

    for (int qt = 0; qt < nqt; ++qt) {
        cp_async_wait<0>();
        __syncthreads();
        if (qt == 0) {
            load_a_frags(kf, sK, warp * 16, lane);
            load_a_frags(vf, sV, warp * 16, lane);
        }
        if (qt + 1 < nqt) {
            load_tile_async(sQ + ((qt + 1) & 1) * kTileBytes, gq, p.ld_qkv, (qt + 1) * kTile, p.S);
            load_tile_async(sDO + ((qt + 1) & 1) * kTileBytes, gdo, p.ld_o, (qt + 1) * kTile, p.S);
            cp_async_commit();
            load_stats(qt + 1, (qt + 1) & 1);
        }
        const uint32_t tQ = sQ + (qt & 1) * kTileBytes, tDO = sDO + (qt & 1) * kTileBytes;
        const float* lse_t = s_lse + (qt & 1) * kTile;
        const float* del_t = s_del + (qt & 1) * kTile;
        float st[8][4], dpt[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            st[j][0] = st[j][1] = st[j][2] = st[j][3] = 0.f;
            dpt[j][0] = dpt[j][1] = dpt[j][2] = dpt[j][3] = 0.f;
        }
        mma_a_tileT(st, kf, tQ, lane);    // S^T[key, query]
        mma_a_tileT(dpt, vf, tDO, lane);  // dP^T[key, query] = V dO^T
        float pt[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int qc = j * 8 + (lane & 3) * 2;  // query column inside the tile
            const int qg = qt * kTile + qc;
            const bool q0v = qg < p.S, q1v = qg + 1 < p.S;
            const float ls0 = lse_t[qc], ls1 = lse_t[qc + 1];
            const float d0 = del_t[qc], d1 = del_t[qc + 1];
            pt[j][0] = (kv0 && q0v) ? exp2f(fmaf(st[j][0], p.scale_log2, kb0) - ls0) : 0.f;
            pt[j][1] = (kv0 && q1v) ? exp2f(fmaf(st[j][1], p.scale_log2, kb0) - ls1) : 0.f;
            pt[j][2] = (kv1 && q0v) ? exp2f(fmaf(st[j][2], p.scale_log2, kb1) - ls0) : 0.f;
            pt[j][3] = (kv1 && q1v) ? exp2f(fmaf(st[j][3], p.scale_log2, kb1) - ls1) : 0.f;
            st[j][0] = pt[j][0] * (dpt[j][0] - d0);
            st[j][1] = pt[j][1] * (dpt[j][1] - d1);
            st[j][2] = pt[j][2] * (dpt[j][2] - d0);
            st[j][3] = pt[j][3] * (dpt[j][3] - d1);
        }
        mma_p_tile(dv, pt, tDO, lane);  // dV += P^T dO     (dO tile stored [query][d] = [k][n])
        mma_p_tile(dk, st, tQ, lane);   // dK += dS^T Q
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { dk[j][0] *= p.scale; dk[j][1] *= p.scale; dk[j][2] *= p.scale; dk[j][3] *= p.scale; }
    if (p.dv_rowscale != nullptr && p.dv_rank1 != nullptr) {
        // rank-1 term of the post-softmax mask add: dV[b,k,h,:] += m[b,k] * dc[b,h,:]
        const float w0 = kv0 ? p.dv_rowscale[(long long)b * p.S + kr0] : 0.f;
        const float w1 = kv1 ? p.dv_rowscale[(long long)b * p.S + kr1] : 0.f;
        const float* dc = p.dv_rank1 + ((long long)b * p.nh + h) * kHeadDim;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float2 c = *reinterpret_cast<const float2*>(dc + j * 8 + (lane & 3) * 2);
            dv[j][0] += w0 * c.x; dv[j][1] += w0 * c.y;
            dv[j][2] += w1 * c.x; dv[j][3] += w1 * c.y;
        }
    }
    const long long out_off = (long long)b * p.S * p.ld_dqkv + h * kHeadDim;
    store_acc_tile(dk, sK, p.dk + out_off, p.ld_dqkv, k0, p.S, warp, lane);
    store_acc_tile(dv, sV, p.dv + out_off, p.ld_dqkv, k0, p.S, warp, lane);
}

static int check_common(const void* q, const void* k, const void* v, long long ld_qkv, int B, int S, int nh, int mode,
                        const float* key_bias, const char* who) {
    TAVK_CHECK(q && k && v, 1, "%s: null q/k/v", who);
    TAVK_CHECK(B > 0 && S > 0 && nh > 0, 1, "%s: bad shape B=%d S=%d nh=%d", who, B, S, nh);
    TAVK_CHECK(ld_qkv % 8 == 0 && ld_qkv >= (long long)nh * kHeadDim, 1, "%s: ld_qkv=%lld must be a multiple of 8 and >= nh*64",
               who, ld_qkv);
    TAVK_CHECK(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) & 15) == 0,
               1, "%s: q/k/v must be 16-byte aligned", who);
    TAVK_CHECK(mode == TAVK_ATTN_NONE || mode == TAVK_ATTN_KEY_BIAS, 1, "%s: bad mode %d", who, mode);
    TAVK_CHECK(mode != TAVK_ATTN_KEY_BIAS || key_bias != nullptr, 1, "%s: TAVK_ATTN_KEY_BIAS needs key_bias", who);
    TAVK_CHECK(nh <= 65535 && B <= 65535, 2, "%s: nh/B exceed the grid limits", who);
    return 0;
}

// attention_tc.cu: tcgen05/TMEM kernels (mask-free mode)
int attn_fwd_tc_launch(const tavk_attn_args* a, cudaStream_t stream);
int attn_bwd_tc_launch(const tavk_attn_bwd_args* a, cudaStream_t stream);
int attn_bias_grads_by_colsum(const tavk_attn_bwd_args* a, cudaStream_t stream);

// TAVK_ATTN_IMPL=legacy forces the mma.sync kernels everywhere (A/B testing); default: tcgen05 where it applies
static bool use_tc_path() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("TAVK_ATTN_IMPL");
        v = (e != nullptr && strcmp(e, "legacy") == 0) ? 0 : 1;
    }
    return v == 1;
}

}  // namespace tavk

using namespace tavk;

extern "C" int tavk_attn_fwd(const tavk_attn_args* a, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    TAVK_CHECK(a != nullptr, 1, "tavk_attn_fwd: null args");
    int rc = check_common(a->q, a->k, a->v, a->ld_qkv, a->B, a->S, a->nh, a->mode, a->key_bias, "tavk_attn_fwd");
    if (rc) return rc;
    TAVK_CHECK(a->o != nullptr && a->ld_o % 8 == 0 && (reinterpret_cast<uintptr_t>(a->o) & 15) == 0, 1,
               "tavk_attn_fwd: bad output");
    if (a->mode == TAVK_ATTN_NONE && a->S >= 128 && use_tc_path()) return attn_fwd_tc_launch(a, stream);
    AttnFwdDev d;
    d.q = reinterpret_cast<const __nv_bfloat16*>(a->q);
    d.k = reinterpret_cast<const __nv_bfloat16*>(a->k);
    d.v = reinterpret_cast<const __nv_bfloat16*>(a->v);
    d.ld_qkv = a->ld_qkv;
    d.o = reinterpret_cast<__nv_bfloat16*>(a->o);
    d.ld_o = a->ld_o;
    d.lse = a->lse;
    d.key_bias = a->mode == TAVK_ATTN_KEY_BIAS ? a->key_bias : nullptr;
    d.B = a->B; d.S = a->S; d.nh = a->nh;
    d.scale_log2 = a->scale * kLog2e;
    dim3 grid((a->S + kTile - 1) / kTile, a->nh, a->B);
    TAVK_CUDA(launch_kernel(attn_fwd_kernel, dim3(grid), dim3(128), (size_t)(0), stream, d));
    TAVK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tavk_attn_bwd(const tavk_attn_bwd_args* a, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    TAVK_CHECK(a != nullptr, 1, "tavk_attn_bwd: null args");
    int rc = check_common(a->q, a->k, a->v, a->ld_qkv, a->B, a->S, a->nh, a->mode, a->key_bias, "tavk_attn_bwd");
    if (rc) return rc;
    TAVK_CHECK(a->o && a->d_o && a->lse && a->delta && a->dq && a->dk && a->dv, 1, "tavk_attn_bwd: null pointer");
    TAVK_CHECK(a->ld_o % 8 == 0 && a->ld_dqkv % 8 == 0, 1, "tavk_attn_bwd: strides must be multiples of 8");
    TAVK_CHECK(((reinterpret_cast<uintptr_t>(a->o) | reinterpret_cast<uintptr_t>(a->d_o) | reinterpret_cast<uintptr_t>(a->dq) |
                 reinterpret_cast<uintptr_t>(a->dk) | reinterpret_cast<uintptr_t>(a->dv)) & 15) == 0,
               1, "tavk_attn_bwd: buffers must be 16-byte aligned");
    TAVK_CHECK((a->dv_rowscale == nullptr) == (a->dv_rank1 == nullptr), 1,
               "tavk_attn_bwd: dv_rowscale and dv_rank1 must be given together");
    TAVK_CHECK(a->dv_rank1 == nullptr || (reinterpret_cast<uintptr_t>(a->dv_rank1) & 7) == 0, 1,
               "tavk_attn_bwd: dv_rank1 must be 8-byte aligned");
    AttnBwdDev d;
    d.q = reinterpret_cast<const __nv_bfloat16*>(a->q);
    d.k = reinterpret_cast<const __nv_bfloat16*>(a->k);
    d.v = reinterpret_cast<const __nv_bfloat16*>(a->v);
    d.ld_qkv = a->ld_qkv;
    d.d_o = reinterpret_cast<const __nv_bfloat16*>(a->d_o);
    d.ld_o = a->ld_o;
    d.lse = a->lse; d.delta = a->delta;
    d.key_bias = a->mode == TAVK_ATTN_KEY_BIAS ? a->key_bias : nullptr;
    d.dq = reinterpret_cast<__nv_bfloat16*>(a->dq);
    d.dk = reinterpret_cast<__nv_bfloat16*>(a->dk);
    d.dv = reinterpret_cast<__nv_bfloat16*>(a->dv);
    d.ld_dqkv = a->ld_dqkv;
    d.dv_rowscale = a->dv_rowscale; d.dv_rank1 = a->dv_rank1;
    d.B = a->B; d.S = a->S; d.nh = a->nh;
    d.scale = a->scale; d.scale_log2 = a->scale * kLog2e;

    const long long pairs = (long long)a->B * a->S * a->nh;
    TAVK_CUDA(launch_kernel(attn_delta_kernel, dim3((int)((pairs * 8 + 255) / 256)), dim3(256), (size_t)(0), stream, 
        reinterpret_cast<const __nv_bfloat16*>(a->o), d.d_o, a->ld_o, a->delta, a->B, a->S, a->nh));
    TAVK_CUDA(cudaGetLastError());

    if (a->mode == TAVK_ATTN_NONE && a->S >= 128 && use_tc_path()) return attn_bwd_tc_launch(a, stream);

    constexpr int kSmemDq = 6 * kTileBytes;
    constexpr int kSmemDkv = 6 * kTileBytes + 4 * kTile * (int)sizeof(float);
    static bool attr_done = false;
    if (!attr_done) {
        TAVK_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemDq));
        TAVK_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemDkv));
        attr_done = true;
    }
    dim3 grid((a->S + kTile - 1) / kTile, a->nh, a->B);
    TAVK_CUDA(launch_kernel(attn_bwd_dkv_kernel, dim3(grid), dim3(128), (size_t)(kSmemDkv), stream, d));
    TAVK_CUDA(cudaGetLastError());
    TAVK_CUDA(launch_kernel(attn_bwd_dq_kernel, dim3(grid), dim3(128), (size_t)(kSmemDq), stream, d));
    TAVK_CUDA(cudaGetLastError());
    return attn_bias_grads_by_colsum(a, stream);
}
