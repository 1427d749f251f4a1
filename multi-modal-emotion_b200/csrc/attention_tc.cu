// tcgen05 / TMEM flash attention for head_dim 64, forward and backward (attention.cu keeps the mma.sync v1 that still
// serves key-bias masks and sequences shorter than 128).
//
// Replaces the same reference op sites as attention.cu: utils/TAVFormer.py:357-387 (QK^T/8, softmax, PV) and the HF
// encoders' attention.  Motivation (profiles/r1_*): at the VideoMAE shape (S=1464, 192 (b,h) pairs) the mma.sync
// kernels were 24 ms of a 90 ms step; the legacy tensor path peaks at a quarter of tcgen05's rate.
//
// Common shape of the three kernels: one CTA = 128 rows (queries, or keys for dK/dV) of one (batch, head); 192 threads:
//   warp 0      TMA producer (3-D tensor maps over [B][S][row], SWIZZLE_128B, rows >= S zero-filled by the hardware)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (accumulators S / dP / O / dQ / dK / dV in TMEM)
//   warps 2..5  elementwise: one thread per tile row (TMEM lane), tcgen05.ld -> ex2 / dS in registers -> bf16 P / dS
//               handed to the next MMA as its A operand: through TENSOR memory (tcgen05.st into the 64 spare columns,
//               tcgen05.mma with A in TMEM) in the forward and dQ kernels, through shared memory (UMMA K-major SW128
//               layout) in the dK/dV kernel, whose 256 columns are all accumulators
// Every kernel keeps within 113 KB of shared memory and 256 TMEM columns so that TWO CTAs share an SM: with d = 64 the
// elementwise stage, not the tensor pipe, sets the pace (16 ex2/clk/SM against 8192 tensor FLOP/clk/SM), and two
// independent CTAs keep each other's idle phases busy.  An earlier one-CTA-per-SM version (8 elementwise warps, 128-key
// tiles, 2 threads per row) measured 286 / 386 / 294 us (fwd / dK,dV / dQ) at B=16, S=1464; with P / dS staged in shared
// memory 191 / 332 / 207 us; these 180 / 321 / 197 us.  What the shared-memory hand-off cost: with N = 64 an MMA's
// operand fetch (6 KB at 128 B/clk) takes longer than its math (32 clk), and the P / dS stores plus the proxy fence
// share that pipe (profiles/r1_ncu_attention_smem_pipe.txt).
#include "../../include/tavk.h"
#include <stdlib.h>

#include "common.cuh"

namespace tavk {

constexpr int kTcQ = 128, kTcKV = 128, kTcD = 64;
constexpr int kTcTile = kTcKV * kTcD * 2;            // 16 KB  (K, V, Q tiles)
constexpr float kTcLog2e = 1.4426950408889634f, kTcLn2 = 0.6931471805599453f;
constexpr float kRescaleThreshold = 8.0f;

// The two-CTA-per-SM kernels use every byte of their 113 KB share, so they cannot afford 1 KB of alignment slack: the
// dynamic shared-memory window of a CTA starts 1024-byte aligned when the kernel has no static shared memory (checked).
TAVK_DEVINL uint8_t* smem_base_1024(uint8_t* raw) {
    if ((smem_u32(raw) & 1023u) != 0u) {
        if (threadIdx.x == 0) printf("tavk: dynamic shared memory is not 1024-byte aligned (%u)\n", smem_u32(raw));
        __trap();
    }
    return raw;
}
TAVK_DEVINL void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
TAVK_DEVINL void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
TAVK_DEVINL void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
TAVK_DEVINL void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 fp32 accumulator columns of this thread's row -> bf16 -> 64 contiguous bytes in global
TAVK_DEVINL void store_row32_bf16(__nv_bfloat16* dst, const float (&v)[32]) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int i = 0; i < 4; ++i)
        d4[i] = make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                           pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
}
struct AttnTcDev {
    __nv_bfloat16* o;
    long long ld_o;
    float* lse;
    int B, S, nh;
    float scale_log2;
};

// =====================================================================================================
// Forward.  KV tile = 64 keys, 4-stage (K_j, V_j) ring; TMEM: S double-buffered (2 x 64 columns), O (64 columns), bf16 P
// double-buffered (2 x 32 columns, two keys per column): the softmax threads write P back with tcgen05.st and PV reads
// its A operand from there — no P in shared memory, no proxy fence.  QK of tile j+1 is issued before PV of tile j, so
// the tensor pipe works while tile j's softmax runs.  Online softmax without any shuffle (thread = row); lazy rescaling: the running max only moves
// (and O in TMEM is only rescaled, tcgen05.ld/st) when it grows by more than 8 in the log2 domain.
// Epilogue: O / l -> bf16 rows, lse = (m + log2 l) ln2.
// =====================================================================================================
constexpr int kF3KV = 64;
constexpr int kF3Stages = 4;
constexpr int kF3KTile = kF3KV * kTcD * 2;            // 8 KB (K or V tile)
constexpr int kF3Smem = kTcTile + kF3Stages * 2 * kF3KTile + 256;   // 80.25 KB (two CTAs per SM)
constexpr int kF3Threads = 6 * 32;
constexpr uint32_t kF3TmemCols = 256;

__global__ void __launch_bounds__(kF3Threads, 2)
attn_fwd_tc64_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_v, const AttnTcDev p) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_base_1024(smem_raw);
    uint8_t* sQ = smem;
    uint8_t* sK = smem + kTcTile;
    uint8_t* sV = sK + kF3Stages * kF3KTile;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kF3Stages * kF3KTile);
    uint64_t* q_full = bars;               // 1
    uint64_t* kv_full = bars + 1;          // kF3Stages
    uint64_t* kv_empty = kv_full + kF3Stages;
    uint64_t* s_full = kv_empty + kF3Stages;  // 2
    uint64_t* p_full = s_full + 2;            // 2
    uint64_t* p_empty = p_full + 2;           // 2
    uint64_t* pv_done = p_empty + 2;          // 1
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(pv_done + 1);

    const int warp_idx = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // warp-uniform
    const int q0 = blockIdx.x * kTcQ, h = blockIdx.y, b = blockIdx.z;
    const int n_tiles = (p.S + kF3KV - 1) / kF3KV;

    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_k);
        tma_prefetch_desc(&tmap_v);
        mbar_init(q_full, 1);
        for (int i = 0; i < kF3Stages; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); mbar_init(&p_empty[i], 1); }
        mbar_init(pv_done, 1);
        mbar_fence_init();
    }
    if (warp_idx == 1) tmem_alloc<kF3TmemCols>(tmem_ptr_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);
    const uint32_t tmem_o = tmem_base + 2 * kF3KV;
    const uint32_t tmem_p = tmem_base + 3 * kF3KV;     // two bf16 P buffers of 32 columns each

    if (warp_idx == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, kTcTile);
            tma_load_3d(sQ, &tmap_q, q_full, h * kTcD, q0, b);
            for (int j = 0; j < n_tiles; ++j) {
                const int st = j % kF3Stages;
                mbar_wait_backoff<128>(&kv_empty[st], ((j / kF3Stages) & 1) ^ 1);
                mbar_arrive_expect_tx(&kv_full[st], 2 * kF3KTile);
                tma_load_3d(sK + st * kF3KTile, &tmap_k, &kv_full[st], h * kTcD, j * kF3KV, b);
                tma_load_3d(sV + st * kF3KTile, &tmap_v, &kv_full[st], h * kTcD, j * kF3KV, b);
            }
        }
    } else if (warp_idx == 1) {
        const bool leader = elect_one();
        constexpr uint32_t idesc_qk = umma_idesc_bf16(kTcQ, kF3KV, false, false);
        constexpr uint32_t idesc_pv = umma_idesc_bf16(kTcQ, kTcD, false, true);
        const uint64_t dq0 = umma_smem_desc(smem_u32(sQ), 16, 1024);
        const uint64_t dk0 = umma_smem_desc(smem_u32(sK), 16, 1024);
        const uint64_t dv0 = umma_smem_desc(smem_u32(sV), 8192, 1024);
        auto issue_qk = [&](int j) {
            const int st = j % kF3Stages;
            mbar_wait_backoff<32>(&kv_full[st], (j / kF3Stages) & 1);
            tc_fence_after();
            if (leader) {
                const uint64_t dk = dk0 + (uint64_t)(st * (kF3KTile >> 4));
                const uint32_t tmem_s = tmem_base + (j & 1) * kF3KV;
#pragma unroll
                for (int k = 0; k < kTcD / 16; ++k) umma_bf16(tmem_s, dq0 + 2 * k, dk + 2 * k, idesc_qk, k > 0 ? 1u : 0u);
                umma_commit(&s_full[j & 1]);
            }
            __syncwarp();
        };
        mbar_wait(q_full, 0);
        tc_fence_after();
        issue_qk(0);
        for (int j = 0; j < n_tiles; ++j) {
            if (j + 1 < n_tiles) issue_qk(j + 1);
            mbar_wait_backoff<32>(&p_full[j & 1], (j >> 1) & 1);
            tc_fence_after();
            if (leader) {
                const int st = j % kF3Stages;
                const uint64_t dv = dv0 + (uint64_t)(st * (kF3KTile >> 4));
#pragma unroll
                for (int k = 0; k < kF3KV / 16; ++k)   // A = P from tensor memory: 16 keys = 8 columns of bf16 pairs
                    umma_bf16_ts(tmem_o, tmem_p + (uint32_t)((j & 1) * 32 + k * 8), dv + (uint64_t)(k * (2048 >> 4)),
                                 idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
                umma_commit(&kv_empty[st]);
                umma_commit(&p_empty[j & 1]);
                umma_commit(pv_done);
            }
            __syncwarp();
        }
    } else {
        // ===================== softmax / epilogue: one thread per query row, all 64 key columns of the tile ============
        const int quarter = warp_idx & 3;
        const int row = quarter * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
        float m_used = -INFINITY, l = 0.f;
        for (int j = 0; j < n_tiles; ++j) {
            mbar_wait(&s_full[j & 1], (j >> 1) & 1);
            tc_fence_after();
            uint32_t sr[64];
            const uint32_t ts = tmem_base + lane_sel + (uint32_t)((j & 1) * kF3KV);
            tmem_ld_32x32(ts + 0, reinterpret_cast<uint32_t(&)[32]>(sr[0]));
            tmem_ld_32x32(ts + 32, reinterpret_cast<uint32_t(&)[32]>(sr[32]));
            tmem_ld_wait();
            const int valid = p.S - j * kF3KV;   // key columns of this tile that exist
            if (valid < 64) {
#pragma unroll
                for (int c = 0; c < 64; ++c)
                    if (c >= valid) sr[c] = 0xff800000u;  // -inf
            }
            // four independent running maxima (a single 64-long dependent chain would cost ~250 cycles of latency)
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
            for (int c = 0; c < 64; c += 8) {
                mx0 = fmaxf(mx0, fmaxf(__uint_as_float(sr[c]), __uint_as_float(sr[c + 1])));
                mx1 = fmaxf(mx1, fmaxf(__uint_as_float(sr[c + 2]), __uint_as_float(sr[c + 3])));
                mx2 = fmaxf(mx2, fmaxf(__uint_as_float(sr[c + 4]), __uint_as_float(sr[c + 5])));
                mx3 = fmaxf(mx3, fmaxf(__uint_as_float(sr[c + 6]), __uint_as_float(sr[c + 7])));
            }
            float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * p.scale_log2;
            float factor = 1.0f;
            const bool grow = mx > m_used + kRescaleThreshold;   // always true on the first tile (m_used = -inf)
            if (grow) {
                factor = ex2_approx(m_used - mx);                 // 0 on the first tile
                l *= factor;
                m_used = mx;
            }
            float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
            const float neg_m = -m_used;
#pragma unroll
            for (int c = 0; c < 64; c += 8) {
                float e[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) e[i] = ex2_approx(fmaf(__uint_as_float(sr[c + i]), p.scale_log2, neg_m));
                sum0 += e[0] + e[1]; sum1 += e[2] + e[3]; sum2 += e[4] + e[5]; sum3 += e[6] + e[7];
#pragma unroll
                for (int i = 0; i < 4; ++i) sr[(c >> 1) + i] = pack_bf16x2(e[2 * i], e[2 * i + 1]);
            }
            l += (sum0 + sum1) + (sum2 + sum3);
            // P[j&1] must no longer be read by PV(j-2)
            if (j >= 2) mbar_wait(&p_empty[j & 1], ((j >> 1) - 1) & 1);
            tmem_st_32x32(tmem_p + lane_sel + (uint32_t)((j & 1) * 32), reinterpret_cast<uint32_t(&)[32]>(sr[0]));
            // Every warp observes EVERY completion of pv_done (one phase per PV(j)), in order.  A parity wait can only tell
            // "odd or even number of completions": a warp that skipped this wait and later found the barrier two phases
            // ahead would see the parity it is waiting for as "still pending" and block forever (observed: with other
            // kernels co-resident, or under CUPTI, PV(n-2) and PV(n-1) both retired before a warp's first epilogue poll).
            // PV(j-1) was issued when the slowest warp finished softmax(j-1) and has retired long before softmax(j) ends.
            if (j > 0) {
                mbar_wait(pv_done, (j - 1) & 1);
                // rare: the running max moved -> rescale this warp's 32 rows of O (needs PV(j-1) retired)
                if (__any_sync(0xffffffffu, grow)) {
                    tc_fence_after();
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        uint32_t orow[32];
                        const uint32_t to = tmem_o + lane_sel + hh * 32;
                        tmem_ld_32x32(to, orow);
                        tmem_ld_wait();
#pragma unroll
                        for (int c = 0; c < 32; ++c) orow[c] = __float_as_uint(__uint_as_float(orow[c]) * factor);
                        tmem_st_32x32(to, orow);
                    }
                }
            }
            tmem_st_wait();            // P (and a rescaled O) are in tensor memory
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[j & 1]);
        }
        // ---- epilogue: normalise and store this thread's 64-column output row (128 contiguous bytes)
        // this warp has seen PV(0..n-2) retire inside the loop: exactly one completion, PV(n-1), is outstanding
        mbar_wait(pv_done, (n_tiles - 1) & 1);
        tc_fence_after();
        const int qrow = q0 + row;
        const float inv = l > 0.f ? 1.0f / l : 0.f;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            uint32_t orow[32];
            tmem_ld_32x32(tmem_o + lane_sel + hh * 32, orow);
            tmem_ld_wait();
            if (qrow < p.S) {
                float v[32];
#pragma unroll
                for (int c = 0; c < 32; ++c) v[c] = __uint_as_float(orow[c]) * inv;
                store_row32_bf16(p.o + ((long long)b * p.S + qrow) * p.ld_o + h * kTcD + hh * 32, v);
            }
        }
        if (qrow < p.S && p.lse)
            p.lse[((long long)b * p.nh + h) * p.S + qrow] = l > 0.f ? (m_used + log2f(l)) * kTcLn2 : -INFINITY;
    }
    tc_fence_before();
    __syncthreads();
    if (warp_idx == 1) {
        tc_fence_after();
        tmem_dealloc<kF3TmemCols>(tmem_base);
    }
}

// =====================================================================================================
// Backward (mask-free mode), two kernels that recompute P from the saved log-sum-exp:
//   dK/dV kernel: CTA = 128 keys of one (b,h); loops over 64-query steps.
//       S^T = K Q_i^T - lse_q/c, dP^T = V dO_i^T - delta_q   (UMMA 128x64x16 x4 each + one statistics k-step -> TMEM)
//       P^T = ex2(S^T c), dS^T = P^T dP^T                     (thread = key row)
//       dV += P^T dO_i, dK += dS^T Q_i               (A = bf16 P^T / dS^T from smem, B = dO_i / Q_i MN-major)
//   dQ kernel:    CTA = 128 queries; loops over 64-key steps.
//       S = Q K_j^T, dP = dO V_j^T -> dS = ex2(S c - lse_row)(dP - delta_row) -> dQ += dS K_j  (A = bf16 dS from TMEM)
// No atomics in the data path, deterministic.  Queries >= S are neutralised with lse = +inf (P = 0), keys >= S by
// zeroing P.  Optional: the q/k/v projection bias gradients (column sums of dQ / dK / dV) from the epilogues.
// =====================================================================================================
constexpr int kBwStep = 64;                              // inner-loop tile (queries for dK/dV, keys for dQ)
constexpr int kBwSmall = kBwStep * kTcD * 2;             // 8 KB
constexpr int kBwPBytes = 128 * kBwStep * 2;             // 16 KB: [128 rows][64 k] bf16, one SW128 atom column

struct AttnTcBwdDev {
    const float *lse, *delta;
    __nv_bfloat16 *dq, *dk, *dv;
    long long ld_dqkv;
    const float *dv_rowscale, *dv_rank1;
    int B, S, nh;
    float scale, scale_log2, inv_scale;
    float *dbq, *dbk, *dbv;      // optional bias-gradient accumulators (column sums of dq / dk / dv), f32 [nh*64]
};

// a -> {hi, lo, lo2, 0, 0, 0, 0, 0} bf16 with hi + lo + lo2 = a to 24 bits: one 16-byte chunk of a statistics row
TAVK_DEVINL uint4 split3_bf16(float a) {
    const __nv_bfloat16 h = __float2bfloat16_rn(a);
    const float r = a - __bfloat162float(h);
    const __nv_bfloat16 l = __float2bfloat16_rn(r);
    const __nv_bfloat16 l2 = __float2bfloat16_rn(r - __bfloat162float(l));
    return make_uint4((uint32_t)__bfloat16_as_ushort(h) | ((uint32_t)__bfloat16_as_ushort(l) << 16),
                      (uint32_t)__bfloat16_as_ushort(l2), 0u, 0u);
}

// 32 packed-pair registers (= 64 bf16... here 16 regs = 32 bf16) -> this row's 4 swizzled 16-byte chunks
TAVK_DEVINL void st_row_chunks(uint32_t tile_base, int row, int half, const uint32_t (&pk)[16]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const uint32_t addr = tile_base + row * 128 + ((((half << 2) + c) ^ (row & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * c]), "r"(pk[4 * c + 1]),
                     "r"(pk[4 * c + 2]), "r"(pk[4 * c + 3])
                     : "memory");
    }
}

// S/dP (TMEM) and P/dS (smem) are single buffers — the overlap comes from the co-resident CTA.
//   s_free / st_free: the elementwise warps have copied S/dP of the step out of TMEM -> the issuer may overwrite them.
constexpr int kB3Stages = 2;
constexpr int kB3Threads = 6 * 32;
// dK/dV kernel: two ring stages (a third does not fit next to K, V, P^T and dS^T in 113 KB).
// Per-query statistics enter the dK/dV kernel through the tensor core: S^T and dP^T get one extra K = 16 step
//   S^T += 1 * (-lse_q / scale),   dP^T += 1 * (-delta_q)
// with A = a [128][16] tile of ones and B = a [64 queries][16] tile whose row q holds the statistic split into three
// bf16 terms (hi + lo + lo2 carries 24 mantissa bits; the ones make the sum independent of WHERE in the row they sit, so
// the 16-byte chunk swizzle of the layout does not matter).  Why: thread = key row, so every thread needed all 64 lse
// and 64 delta values of a step, and a broadcast shared-memory load costs one wavefront per 4 bytes — ncu counted 598
// load + 319 store + 768 tensor-operand wavefronts per step = the kernel's whole duration (shared-memory pipe 100 %
// busy, profiles/r1_ncu_attention_smem_pipe.txt).  The extra k-steps read 96 wavefronts instead of the 598.  On its own
// this moved the bound to the issuer's program order (see the ready-first loop below): 332 -> 355 us; with it 321 us.
constexpr int kAugRow = 32;                              // bytes per row of a 16 x bf16 operand slice (SWIZZLE_32B)
constexpr int kAugOnes = 128 * kAugRow;                  // 4 KB
constexpr int kAugX = kBwStep * kAugRow;                 // 2 KB per statistic per stage
constexpr int kDkv3Smem = 2 * kTcTile + kB3Stages * 2 * kBwSmall + 2 * kBwPBytes + kAugOnes + kB3Stages * 2 * kAugX + 1024 + 256;
constexpr int kDq3Stages = 4;             // K_j/V_j ring of the dQ kernel: three steps of TMA look-ahead (2 stages: 258 us,
                                          // 3 stages: 208 us at B=16 S=1464 — the ring depth, not the math, set the pace)
constexpr int kDq3Smem = 2 * kTcTile + kDq3Stages * 2 * kBwSmall + 256;

// Development-only event trace (tools/attn_trace.py builds a private copy of the library with -DTAVK_ATTN_TRACE): the
// chosen CTAs write clock64() at each hand-off of every step; compiled out of libtavk.so.
#ifdef TAVK_ATTN_TRACE
__device__ long long* g_attn_trace = nullptr;
constexpr int kTraceSlots = 16, kTraceCtas = 8, kTraceFirst = 1000;
#define TAVK_TRACE(step, slot)                                                                                      \
    do {                                                                                                            \
        if (trace_cta >= 0 && g_attn_trace != nullptr && (threadIdx.x & 31) == 0)                                   \
            g_attn_trace[((long long)trace_cta * 64 + (step)) * kTraceSlots + (slot)] = clock64();                  \
    } while (0)
#else
#define TAVK_TRACE(step, slot) do { } while (0)
#endif

__global__ void __launch_bounds__(kB3Threads, 2)
attn_bwd_dkv_tc2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                        const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                        const AttnTcBwdDev p) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sK = smem;
    uint8_t* sV = smem + kTcTile;
    uint8_t* sQ = sV + kTcTile;                       // ring: [stage] Q_i
    uint8_t* sDO = sQ + kB3Stages * kBwSmall;         // ring: [stage] dO_i
    uint8_t* sP = sDO + kB3Stages * kBwSmall;         // P^T
    uint8_t* sDS = sP + kBwPBytes;                    // dS^T
    uint8_t* sOnes = sDS + kBwPBytes;                 // [128][16] bf16 ones (A operand of the statistics k-step)
    uint8_t* sXL = sOnes + kAugOnes;                  // ring: [stage][64 queries][16] bf16, row q = split(-lse_q / scale)
    uint8_t* sXD = sXL + kB3Stages * kAugX;           // ring: [stage][64 queries][16] bf16, row q = split(-delta_q)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sXD + kB3Stages * kAugX);
    uint64_t* kv_full = bars;
    uint64_t* qdo_full = bars + 1;
    uint64_t* qdo_empty = qdo_full + kB3Stages;
    uint64_t* st_full = qdo_empty + kB3Stages;
    uint64_t* st_free = st_full + 1;
    uint64_t* pds_full = st_free + 1;
    uint64_t* pds_empty = pds_full + 1;
    uint64_t* done = pds_empty + 1;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(done + 1);

    const int warp_idx = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const int k0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
    const int n_steps = (p.S + kBwStep - 1) / kBwStep;
#ifdef TAVK_ATTN_TRACE
    const int lin_cta = (int)(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z));
    const int trace_cta = (lin_cta >= kTraceFirst && lin_cta < kTraceFirst + kTraceCtas) ? lin_cta - kTraceFirst : -1;
    if (trace_cta >= 0 && threadIdx.x == 0 && g_attn_trace != nullptr) {
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g_attn_trace[((long long)trace_cta * 64 + 63) * kTraceSlots] = smid;
        g_attn_trace[((long long)trace_cta * 64 + 63) * kTraceSlots + 1] = clock64();
    }
#endif

    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v); tma_prefetch_desc(&tmap_do);
        mbar_init(kv_full, 1);
        for (int i = 0; i < kB3Stages; ++i) { mbar_init(&qdo_full[i], 2); mbar_init(&qdo_empty[i], 1); }  // TMA + stats
        mbar_init(st_full, 1); mbar_init(st_free, 4); mbar_init(pds_full, 4); mbar_init(pds_empty, 1);
        mbar_init(done, 1);
        mbar_fence_init();
    }
    if (warp_idx == 1) tmem_alloc<256>(tmem_ptr_smem);
    for (int c = threadIdx.x; c < kAugOnes / 16; c += kB3Threads)
        reinterpret_cast<uint4*>(sOnes)[c] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);
    const uint32_t tmem_dv = tmem_base + 128, tmem_dk = tmem_base + 192;

    if (warp_idx == 0) {
        // producer: lane 0 drives TMA; all lanes stage the per-query statistics of the step (lse*log2e, delta)
        const long long stat_off = ((long long)b * p.nh + h) * p.S;
        if (lane == 0) {
            mbar_arrive_expect_tx(kv_full, 2 * kTcTile);
            tma_load_3d(sK, &tmap_k, kv_full, h * kTcD, k0, b);
            tma_load_3d(sV, &tmap_v, kv_full, h * kTcD, k0, b);
        }
        for (int i = 0; i < n_steps; ++i) {
            const int st = i % kB3Stages;
            mbar_wait_backoff<128>(&qdo_empty[st], ((i / kB3Stages) & 1) ^ 1);
            TAVK_TRACE(i, 0);
            if (lane == 0) {
                mbar_arrive_expect_tx(&qdo_full[st], 2 * kBwSmall);
                tma_load_3d(sQ + st * kBwSmall, &tmap_q, &qdo_full[st], h * kTcD, i * kBwStep, b);
                tma_load_3d(sDO + st * kBwSmall, &tmap_do, &qdo_full[st], h * kTcD, i * kBwStep, b);
            }
#pragma unroll
            for (int t = lane; t < kBwStep; t += 32) {
                const int qi = i * kBwStep + t;
                const bool ok = qi < p.S;
                // (S^T - lse/scale) * scale*log2e = S^T*scale*log2e - lse*log2e; queries >= S: -1e30 -> P = 0
                uint4* xl = reinterpret_cast<uint4*>(sXL + st * kAugX + t * kAugRow);
                uint4* xd = reinterpret_cast<uint4*>(sXD + st * kAugX + t * kAugRow);
                xl[0] = split3_bf16(ok ? -__ldg(p.lse + stat_off + qi) * p.inv_scale : -1.0e30f);
                xl[1] = make_uint4(0u, 0u, 0u, 0u);
                xd[0] = split3_bf16(ok ? -__ldg(p.delta + stat_off + qi) : 0.f);
                xd[1] = make_uint4(0u, 0u, 0u, 0u);
            }
            fence_proxy_async_smem();   // the statistics rows are read by the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(&qdo_full[st]);
            TAVK_TRACE(i, 1);
        }
    } else if (warp_idx == 1) {
        const bool leader = elect_one();
        constexpr uint32_t idesc_kk = umma_idesc_bf16(128, kBwStep, false, false);  // S^T, dP^T
        constexpr uint32_t idesc_mn = umma_idesc_bf16(128, kTcD, false, true);      // dV, dK (B MN-major)
        const uint64_t dK = umma_smem_desc(smem_u32(sK), 16, 1024), dV = umma_smem_desc(smem_u32(sV), 16, 1024);
        const uint64_t dQk0 = umma_smem_desc(smem_u32(sQ), 16, 1024), dDOk0 = umma_smem_desc(smem_u32(sDO), 16, 1024);
        const uint64_t dQm0 = umma_smem_desc(smem_u32(sQ), 8192, 1024), dDOm0 = umma_smem_desc(smem_u32(sDO), 8192, 1024);
        const uint64_t dP0 = umma_smem_desc(smem_u32(sP), 16, 1024), dDS0 = umma_smem_desc(smem_u32(sDS), 16, 1024);
        const uint64_t dOnes = umma_smem_desc_sw32(smem_u32(sOnes), 256);
        const uint64_t dXL0 = umma_smem_desc_sw32(smem_u32(sXL), 256), dXD0 = umma_smem_desc_sw32(smem_u32(sXD), 256);
        mbar_wait(kv_full, 0);
        tc_fence_after();
        // S^T / dP^T of step i1 (needs the step's Q/dO stage and the elementwise warps' copy of step i1-1 out of TMEM)
        auto issue_scores = [&](int i) {
            const int st = i % kB3Stages;
            const uint64_t so = (uint64_t)(st * (kBwSmall >> 4)), sx = (uint64_t)(st * (kAugX >> 4));
            const uint32_t t_st = tmem_base, t_dp = tmem_base + 64;
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(t_st, dK + 2 * k, dQk0 + so + 2 * k, idesc_kk, k > 0 ? 1u : 0u);
            umma_bf16(t_st, dOnes, dXL0 + sx, idesc_kk, 1u);    // S^T -= lse_q / scale
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(t_dp, dV + 2 * k, dDOk0 + so + 2 * k, idesc_kk, k > 0 ? 1u : 0u);
            umma_bf16(t_dp, dOnes, dXD0 + sx, idesc_kk, 1u);    // dP^T -= delta_q
            umma_commit(st_full);
        };
        // dV / dK of step `kstep` (needs its P^T / dS^T staged); frees the step's Q/dO stage and the P^T/dS^T buffers
        auto issue_grads = [&](int kstep) {
            const int st = kstep % kB3Stages;
            const uint64_t so = (uint64_t)(st * (kBwSmall >> 4));
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_bf16(tmem_dv, dP0 + 2 * k, dDOm0 + so + (uint64_t)(k * 128), idesc_mn, (kstep > 0 || k > 0) ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_bf16(tmem_dk, dDS0 + 2 * k, dQm0 + so + (uint64_t)(k * 128), idesc_mn, (kstep > 0 || k > 0) ? 1u : 0u);
            umma_commit(&qdo_empty[st]);
            umma_commit(pds_empty);
        };
        {
            // Whichever is ready first.  With a two-stage Q/dO ring the reload of a stage (TMA round trip + statistics,
            // ~1400 cycles measured) starts when dV/dK of step i-1 complete; in program order S^T(i+1) -> dV/dK(i) the
            // issuer sat in the wait for that reload while P^T/dS^T of step i were already staged, which delayed the
            // NEXT reload in turn (tools/attn_trace.py timeline: the whole period was this chain).
            int i1 = 0, i2 = 0;
            uint32_t idle = 0;
            while (i2 < n_steps) {
                bool did = false;
                if (i2 < i1 && mbar_test_warp(pds_full, i2 & 1)) {
                    TAVK_TRACE(i2, 5);
                    tc_fence_after();
                    if (leader) issue_grads(i2);
                    __syncwarp();
                    TAVK_TRACE(i2, 6);
                    ++i2;
                    did = true;
                }
                if (i1 < n_steps && mbar_test_warp(&qdo_full[i1 % kB3Stages], (i1 / kB3Stages) & 1) &&
                    (i1 == 0 || mbar_test_warp(st_free, (i1 - 1) & 1))) {
                    TAVK_TRACE(i1, 3);
                    tc_fence_after();
                    if (leader) issue_scores(i1);
                    __syncwarp();
                    TAVK_TRACE(i1, 4);
                    ++i1;
                    did = true;
                }
                if (!did) {
                    __nanosleep(20);
                    if (++idle > (1u << 26)) {
                        if (lane == 0) printf("tavk: dK/dV issuer stalled (block %d, i1 %d, i2 %d)\n", (int)blockIdx.x, i1, i2);
                        __trap();
                    }
                }
            }
        }
        if (leader) umma_commit(done);
        __syncwarp();
    } else {
        const int quarter = warp_idx & 3;
        const int row = quarter * 32 + lane;                       // key row inside the tile
        const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
        for (int i = 0; i < n_steps; ++i) {
            if (quarter == 0) TAVK_TRACE(i, 7);
            mbar_wait(st_full, i & 1);
            if (quarter == 0) TAVK_TRACE(i, 8);
            tc_fence_after();
            {
                // all 128 accumulator words of the row first, so the issuer can overwrite S^T / dP^T (step i+1) while
                // this step's ex2 / products / stores run
                uint32_t s0[32], s1[32], d0[32], d1[32];
                tmem_ld_32x32(tmem_base + lane_sel, s0);
                tmem_ld_32x32(tmem_base + lane_sel + 32, s1);
                tmem_ld_32x32(tmem_base + lane_sel + 64, d0);
                tmem_ld_32x32(tmem_base + lane_sel + 96, d1);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(st_free);
                if (quarter == 0) TAVK_TRACE(i, 9);
                uint32_t pk[16], dsk[16];
#pragma unroll
                for (int c = 0; c < 32; c += 2) {
                    const float p0 = ex2_approx(__uint_as_float(s0[c]) * p.scale_log2);
                    const float p1 = ex2_approx(__uint_as_float(s0[c + 1]) * p.scale_log2);
                    pk[c >> 1] = pack_bf16x2(p0, p1);
                    dsk[c >> 1] = pack_bf16x2(p0 * __uint_as_float(d0[c]), p1 * __uint_as_float(d0[c + 1]));
                }
                if (quarter == 0) TAVK_TRACE(i, 10);
                if (i >= 1) mbar_wait(pds_empty, (i - 1) & 1);   // dV/dK MMAs of step i-1 have read P^T/dS^T
                if (quarter == 0) TAVK_TRACE(i, 11);
                st_row_chunks(smem_u32(sP), row, 0, pk);
                st_row_chunks(smem_u32(sDS), row, 0, dsk);
#pragma unroll
                for (int c = 0; c < 32; c += 2) {
                    const float p0 = ex2_approx(__uint_as_float(s1[c]) * p.scale_log2);
                    const float p1 = ex2_approx(__uint_as_float(s1[c + 1]) * p.scale_log2);
                    pk[c >> 1] = pack_bf16x2(p0, p1);
                    dsk[c >> 1] = pack_bf16x2(p0 * __uint_as_float(d1[c]), p1 * __uint_as_float(d1[c + 1]));
                }
                st_row_chunks(smem_u32(sP), row, 1, pk);
                st_row_chunks(smem_u32(sDS), row, 1, dsk);
            }
            if (quarter == 0) TAVK_TRACE(i, 12);
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(pds_full);
            if (quarter == 0) TAVK_TRACE(i, 13);
        }
        mbar_wait(done, 0);
        tc_fence_after();
        const int key = k0 + row;
        float w = 0.f;
        const bool rank1 = p.dv_rowscale != nullptr && key < p.S;
        if (rank1) w = p.dv_rowscale[(long long)b * p.S + key];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t r[32];
            float v[32];
            const long long out_off = ((long long)b * p.S + key) * p.ld_dqkv + h * kTcD + half * 32;
            tmem_ld_32x32(tmem_dk + lane_sel + half * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 32; ++c) v[c] = key < p.S ? __uint_as_float(r[c]) * p.scale : 0.f;
            if (key < p.S) store_row32_bf16(p.dk + out_off, v);
            if (p.dbk != nullptr) atomicAdd(p.dbk + h * kTcD + half * 32 + lane, warp_colsum32(v, lane));
            // dV (+ rank-1 term of the post-softmax mask: dV[b,k,h,:] += m[b,k] * dc[b,h,:])
            tmem_ld_32x32(tmem_dv + lane_sel + half * 32, r);
            tmem_ld_wait();
            const float* dc = rank1 ? p.dv_rank1 + ((long long)b * p.nh + h) * kTcD + half * 32 : nullptr;
#pragma unroll
            for (int c = 0; c < 32; ++c) v[c] = key < p.S ? __uint_as_float(r[c]) + (dc ? w * __ldg(dc + c) : 0.f) : 0.f;
            if (key < p.S) store_row32_bf16(p.dv + out_off, v);
            if (p.dbv != nullptr) atomicAdd(p.dbv + h * kTcD + half * 32 + lane, warp_colsum32(v, lane));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp_idx == 1) {
        tc_fence_after();
        tmem_dealloc<256>(tmem_base);
    }
}

__global__ void __launch_bounds__(kB3Threads, 2)
attn_bwd_dq_tc2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                       const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                       const AttnTcBwdDev p) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_base_1024(smem_raw);
    uint8_t* sQ = smem;
    uint8_t* sDO = smem + kTcTile;
    uint8_t* sK = sDO + kTcTile;                      // ring
    uint8_t* sV = sK + kDq3Stages * kBwSmall;          // ring
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kDq3Stages * kBwSmall);
    uint64_t* qdo_full = bars;
    uint64_t* kv_full = bars + 1;
    uint64_t* kv_empty = kv_full + kDq3Stages;
    uint64_t* s_full = kv_empty + kDq3Stages;
    uint64_t* s_free = s_full + 1;
    uint64_t* ds_full = s_free + 1;
    uint64_t* ds_empty = ds_full + 1;
    uint64_t* done = ds_empty + 1;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(done + 1);

    const int warp_idx = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
    const int n_steps = (p.S + kBwStep - 1) / kBwStep;

    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v); tma_prefetch_desc(&tmap_do);
        mbar_init(qdo_full, 1);
        for (int i = 0; i < kDq3Stages; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
        mbar_init(s_full, 1); mbar_init(s_free, 4); mbar_init(ds_full, 4); mbar_init(ds_empty, 1);
        mbar_init(done, 1);
        mbar_fence_init();
    }
    if (warp_idx == 1) tmem_alloc<256>(tmem_ptr_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);
    const uint32_t tmem_dq = tmem_base + 128;
    const uint32_t tmem_ds = tmem_base + 192;          // bf16 dS of the step, 32 columns

    if (warp_idx == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(qdo_full, 2 * kTcTile);
            tma_load_3d(sQ, &tmap_q, qdo_full, h * kTcD, q0, b);
            tma_load_3d(sDO, &tmap_do, qdo_full, h * kTcD, q0, b);
            for (int j = 0; j < n_steps; ++j) {
                const int st = j % kDq3Stages;
                mbar_wait_backoff<128>(&kv_empty[st], ((j / kDq3Stages) & 1) ^ 1);
                mbar_arrive_expect_tx(&kv_full[st], 2 * kBwSmall);
                tma_load_3d(sK + st * kBwSmall, &tmap_k, &kv_full[st], h * kTcD, j * kBwStep, b);
                tma_load_3d(sV + st * kBwSmall, &tmap_v, &kv_full[st], h * kTcD, j * kBwStep, b);
            }
        }
    } else if (warp_idx == 1) {
        const bool leader = elect_one();
        constexpr uint32_t idesc_kk = umma_idesc_bf16(128, kBwStep, false, false);  // S, dP
        constexpr uint32_t idesc_mn = umma_idesc_bf16(128, kTcD, false, true);      // dQ (B = K_j MN-major)
        const uint64_t dQ = umma_smem_desc(smem_u32(sQ), 16, 1024), dDO = umma_smem_desc(smem_u32(sDO), 16, 1024);
        const uint64_t dKk0 = umma_smem_desc(smem_u32(sK), 16, 1024), dVk0 = umma_smem_desc(smem_u32(sV), 16, 1024);
        const uint64_t dKm0 = umma_smem_desc(smem_u32(sK), 8192, 1024);
        mbar_wait(qdo_full, 0);
        tc_fence_after();
        for (int j = 0; j <= n_steps; ++j) {
            if (j < n_steps) {
                const int st = j % kDq3Stages;
                mbar_wait_backoff<32>(&kv_full[st], (j / kDq3Stages) & 1);
                if (j >= 1) mbar_wait_backoff<32>(s_free, (j - 1) & 1);
                tc_fence_after();
                if (leader) {
                    const uint64_t so = (uint64_t)(st * (kBwSmall >> 4));
                    const uint32_t t_s = tmem_base, t_dp = tmem_base + 64;
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(t_s, dQ + 2 * k, dKk0 + so + 2 * k, idesc_kk, k > 0 ? 1u : 0u);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(t_dp, dDO + 2 * k, dVk0 + so + 2 * k, idesc_kk, k > 0 ? 1u : 0u);
                    umma_commit(s_full);
                }
                __syncwarp();
            }
            if (j >= 1) {
                const int kstep = j - 1, st = kstep % kDq3Stages;
                mbar_wait_backoff<32>(ds_full, kstep & 1);
                tc_fence_after();
                if (leader) {
                    const uint64_t so = (uint64_t)(st * (kBwSmall >> 4));
#pragma unroll
                    for (int k = 0; k < 4; ++k)   // A = dS from tensor memory
                        umma_bf16_ts(tmem_dq, tmem_ds + (uint32_t)(k * 8), dKm0 + so + (uint64_t)(k * 128), idesc_mn, (kstep > 0 || k > 0) ? 1u : 0u);
                    umma_commit(&kv_empty[st]);
                    umma_commit(ds_empty);
                }
                __syncwarp();
            }
        }
        if (leader) umma_commit(done);
        __syncwarp();
    } else {
        const int quarter = warp_idx & 3;
        const int row = quarter * 32 + lane;                       // query row inside the tile
        const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
        const int qrow = q0 + row;
        const long long stat_off = ((long long)b * p.nh + h) * p.S;
        const float lse2 = qrow < p.S ? p.lse[stat_off + qrow] * kTcLog2e : INFINITY;
        const float dlt = qrow < p.S ? p.delta[stat_off + qrow] : 0.f;
        for (int j = 0; j < n_steps; ++j) {
            mbar_wait(s_full, j & 1);
            tc_fence_after();
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t s[32], dp[32];
                tmem_ld_32x32(tmem_base + lane_sel + half * 32, s);
                tmem_ld_32x32(tmem_base + lane_sel + 64 + half * 32, dp);
                tmem_ld_wait();
                if (half == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(s_free);
                }
                const int valid = p.S - (j * kBwStep + half * 32);     // key columns of this slab that exist
                uint32_t dsk[16];
                if (valid >= 32) {                                     // uniform: only the last step has a ragged edge
#pragma unroll
                    for (int c = 0; c < 32; c += 2) {
                        const float p0 = ex2_approx(fmaf(__uint_as_float(s[c]), p.scale_log2, -lse2));
                        const float p1 = ex2_approx(fmaf(__uint_as_float(s[c + 1]), p.scale_log2, -lse2));
                        dsk[c >> 1] = pack_bf16x2(p0 * (__uint_as_float(dp[c]) - dlt), p1 * (__uint_as_float(dp[c + 1]) - dlt));
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 32; c += 2) {
                        float p0 = ex2_approx(fmaf(__uint_as_float(s[c]), p.scale_log2, -lse2));
                        float p1 = ex2_approx(fmaf(__uint_as_float(s[c + 1]), p.scale_log2, -lse2));
                        if (c >= valid) p0 = 0.f;
                        if (c + 1 >= valid) p1 = 0.f;
                        dsk[c >> 1] = pack_bf16x2(p0 * (__uint_as_float(dp[c]) - dlt), p1 * (__uint_as_float(dp[c + 1]) - dlt));
                    }
                }
                if (half == 0 && j >= 1) mbar_wait(ds_empty, (j - 1) & 1);   // the dQ MMAs of step j-1 have read dS
                tmem_st_32x16(tmem_ds + lane_sel + (uint32_t)(half * 16), dsk);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(ds_full);
        }
        mbar_wait(done, 0);
        tc_fence_after();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t r[32];
            float v[32];
            tmem_ld_32x32(tmem_dq + lane_sel + half * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 32; ++c) v[c] = qrow < p.S ? __uint_as_float(r[c]) * p.scale : 0.f;
            if (qrow < p.S) store_row32_bf16(p.dq + ((long long)b * p.S + qrow) * p.ld_dqkv + h * kTcD + half * 32, v);
            if (p.dbq != nullptr) atomicAdd(p.dbq + h * kTcD + half * 32 + lane, warp_colsum32(v, lane));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp_idx == 1) {
        tc_fence_after();
        tmem_dealloc<256>(tmem_base);
    }
}

// ---------------------------------------------------------------- host
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (fn == nullptr) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    }
    return fn;
}

// bf16 [B][S][cols] with row pitch ld elements; box = {64 cols, box_rows, 1}
int make_tmap_bsd(CUtensorMap* map, const void* base, int B, int S, int cols, long long ld, int box_rows) {
    PFN_encodeTiled enc = encode_fn();
    TAVK_CHECK(enc != nullptr, 3, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)S, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)S * (cuuint64_t)ld * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TAVK_CHECK(r == CUDA_SUCCESS, 3, "cuTensorMapEncodeTiled(3d) failed (%d) B=%d S=%d cols=%d ld=%lld", (int)r, B, S, cols, ld);
    return 0;
}

int attn_fwd_tc_launch(const tavk_attn_args* a, cudaStream_t stream) {
    CUtensorMap tq, tk64, tv64;
    const int cols = a->nh * kTcD;
    int rc = make_tmap_bsd(&tq, a->q, a->B, a->S, cols, a->ld_qkv, kTcQ);
    if (rc) return rc;
    if ((rc = make_tmap_bsd(&tk64, a->k, a->B, a->S, cols, a->ld_qkv, kF3KV))) return rc;
    if ((rc = make_tmap_bsd(&tv64, a->v, a->B, a->S, cols, a->ld_qkv, kF3KV))) return rc;
    AttnTcDev d;
    d.o = reinterpret_cast<__nv_bfloat16*>(a->o);
    d.ld_o = a->ld_o;
    d.lse = a->lse;
    d.B = a->B; d.S = a->S; d.nh = a->nh;
    d.scale_log2 = a->scale * kTcLog2e;
    static bool attr_done = false;
    if (!attr_done) {
        TAVK_CUDA(cudaFuncSetAttribute(attn_fwd_tc64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kF3Smem));
        attr_done = true;
    }
    dim3 grid((a->S + kTcQ - 1) / kTcQ, a->nh, a->B);
    TAVK_CUDA(launch_kernel(attn_fwd_tc64_kernel, dim3(grid), dim3(kF3Threads), (size_t)(kF3Smem), stream, tq, tk64, tv64, d));
    TAVK_CUDA(cudaGetLastError());
    return 0;
}

// Bias gradients for the kernels that do not accumulate them in their epilogue (v1 mma.sync path, v2 A/B kernels): one
// column-sum pass per requested vector over the dq / dk / dv just written.
int attn_bias_grads_by_colsum(const tavk_attn_bwd_args* a, cudaStream_t stream) {
    const int M = a->B * a->S, N = a->nh * kTcD;
    const void* src[3] = {a->dq, a->dk, a->dv};
    float* dst[3] = {a->dbq, a->dbk, a->dbv};
    for (int i = 0; i < 3; ++i) {
        if (dst[i] == nullptr) continue;
        const int rc = tavk_colsum(src[i], TAVK_BF16, a->ld_dqkv, dst[i], M, N, 1, stream);
        if (rc) return rc;
    }
    return 0;
}

#ifdef TAVK_ATTN_TRACE
extern "C" int tavk_debug_set_attn_trace(void* buf) {
    long long* pbuf = reinterpret_cast<long long*>(buf);
    return cudaMemcpyToSymbol(g_attn_trace, &pbuf, sizeof(pbuf)) == cudaSuccess ? 0 : 1;
}
#endif

// delta must already hold rowsum(dO * O) (attention.cu: attn_delta_kernel)
int attn_bwd_tc_launch(const tavk_attn_bwd_args* a, cudaStream_t stream) {
    CUtensorMap tq64, tk64, tv64, tdo64, tq128, tk128, tv128, tdo128;
    const int cols = a->nh * kTcD;
    int rc = 0;
    if ((rc = make_tmap_bsd(&tq64, a->q, a->B, a->S, cols, a->ld_qkv, kBwStep))) return rc;
    if ((rc = make_tmap_bsd(&tk64, a->k, a->B, a->S, cols, a->ld_qkv, kBwStep))) return rc;
    if ((rc = make_tmap_bsd(&tv64, a->v, a->B, a->S, cols, a->ld_qkv, kBwStep))) return rc;
    if ((rc = make_tmap_bsd(&tdo64, a->d_o, a->B, a->S, cols, a->ld_o, kBwStep))) return rc;
    if ((rc = make_tmap_bsd(&tq128, a->q, a->B, a->S, cols, a->ld_qkv, 128))) return rc;
    if ((rc = make_tmap_bsd(&tk128, a->k, a->B, a->S, cols, a->ld_qkv, 128))) return rc;
    if ((rc = make_tmap_bsd(&tv128, a->v, a->B, a->S, cols, a->ld_qkv, 128))) return rc;
    if ((rc = make_tmap_bsd(&tdo128, a->d_o, a->B, a->S, cols, a->ld_o, 128))) return rc;
    AttnTcBwdDev d;
    d.lse = a->lse; d.delta = a->delta;
    d.dq = reinterpret_cast<__nv_bfloat16*>(a->dq);
    d.dk = reinterpret_cast<__nv_bfloat16*>(a->dk);
    d.dv = reinterpret_cast<__nv_bfloat16*>(a->dv);
    d.ld_dqkv = a->ld_dqkv;
    d.dv_rowscale = a->dv_rowscale; d.dv_rank1 = a->dv_rank1;
    d.B = a->B; d.S = a->S; d.nh = a->nh;
    d.scale = a->scale; d.scale_log2 = a->scale * kTcLog2e; d.inv_scale = 1.0f / a->scale;
    d.dbq = a->dbq; d.dbk = a->dbk; d.dbv = a->dbv;
    static bool attr_done = false;
    if (!attr_done) {
        TAVK_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDkv3Smem));
        TAVK_CUDA(cudaFuncSetAttribute(attn_bwd_dq_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDq3Smem));
        attr_done = true;
    }
    dim3 grid((a->S + 127) / 128, a->nh, a->B);
    TAVK_CUDA(launch_kernel(attn_bwd_dkv_tc2_kernel, dim3(grid), dim3(kB3Threads), (size_t)(kDkv3Smem), stream, tq64, tk128, tv128, tdo64, d));
    TAVK_CUDA(launch_kernel(attn_bwd_dq_tc2_kernel, dim3(grid), dim3(kB3Threads), (size_t)(kDq3Smem), stream, tq128, tk64, tv64, tdo128, d));
    TAVK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace tavk
