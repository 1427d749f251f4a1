// tavk_gemm_bf16: C[M,N] = epilogue(alpha * A[M,K] · B[N,K]^T), bf16 operands, fp32 accumulation in TMEM.
//
// Replaces every nn.Linear / F.linear contraction on the TAV hot path (reference utils/TAVFormer.py:348-350 QKV,
// :422 attention out-proj, :404 FFN up, :434 FFN down; models/tav.py:264,457 audio projection) and, through the
// operand-major flags, their autograd backward (dgrad: A = dY K-major, B = W MN-major; wgrad: both MN-major).
//
// B200 design (one CTA per SM, persistent over output tiles, warp-specialised):
//   warp 0      : TMA producer  (cp.async.bulk.tensor.2d, SWIZZLE_128B boxes, mbarrier complete_tx)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BLOCK_N x 16, kind::f16)
//   warps 2..9  : epilogue (tcgen05.ld 32x32b.x32 -> smem transpose patch -> coalesced bias / GELU / residual /
//                 row-bias / column-sum -> global)
//   smem ring of kStages {A tile 128x64, B tile BLOCK_Nx64}; TMEM double-buffered accumulator so the
//   epilogue of tile i overlaps the MMA main loop of tile i+1.
// CTA-pair mode (template CTAS = 2, tcgen05 cta_group::2): the grid is launched as clusters of two CTAs (one TPC); a pair
//   owns a 256 x BLOCK_N output tile.  Each CTA loads its own 128 rows of A and HALF of the B tile (BLOCK_N/2 rows), the
//   leader CTA's warp 1 issues UMMA 256 x BLOCK_N x 16 which reads B from both CTAs' shared memory, and each CTA's TMEM
//   receives its 128 rows.  Per CTA and k-block that is 32 KB instead of 48 KB of TMA traffic and shared-memory fill, and
//   the ring is 6 deep instead of 4 (1.5x more k in flight per SM), which is what the single-CTA main loop was short of.
//   Barriers: both CTAs' TMA loads complete_tx on the LEADER's full barrier (armed with the bytes of both), the MMA
//   commits multicast to the empty / tmem_full barriers of both CTAs, and the peer's epilogue warps release the
//   accumulator with a remote arrive on the leader's tmem_empty barrier.  Static tile order (no dynamic scheduler).
// K-major operand tile  : one TMA box {64 k, rows}; UMMA desc SBO = 1024 B, K advance = +32 B per UMMA_K.
// MN-major operand tile : rows/64 TMA boxes {64 mn, 64 k}; UMMA desc LBO = 8192 B (next 64-wide MN group),
//                         SBO = 1024 B (next 8 k rows), K advance = +2048 B per UMMA_K.
#include <stdlib.h>

#include "common.cuh"
#include "../../include/tavk.h"

namespace tavk {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
constexpr int kNumEpiWarps = 8;
constexpr int kGemmThreads = (2 + kNumEpiWarps) * 32;

template <int BLOCK_N, int CTAS>
struct GemmCfg {
    static_assert(CTAS == 1 || (CTAS == 2 && BLOCK_N >= 128), "CTA pairs: BLOCK_N 128 or 256");
    static constexpr int kBRows = BLOCK_N / CTAS;           // rows of the B tile THIS CTA loads
    static constexpr int kStages = CTAS == 2 ? (BLOCK_N == 256 ? 6 : 8) : ((BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 6 : 8));
    static constexpr int kABytes = kBlockM * kBlockK * 2;
    static constexpr int kBBytes = kBRows * kBlockK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kTmemCols = 2 * BLOCK_N;  // double-buffered accumulator (128, 256 or 512 columns)
    static constexpr int kStagingBytes = kNumEpiWarps * 4096;   // one 32x32 fp32 transpose patch per epilogue warp
    static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 /*align slack*/ + 512 /*barriers*/;
};

struct GemmDev {
    int M, N, K;
    int m_tile;                     // output rows per tile: 128, or 256 for a CTA pair (each CTA owns 128 of them)
    int num_m_blocks, num_n_blocks, k_splits, kb_per_split, num_kb;
    void* out;
    long long ldo;
    int out_bf16;
    __nv_bfloat16* out2;
    long long ldo2;
    const float* bias;
    const float* resid;
    long long ldr;
    const float* rowbias;
    int rows_per_group;
    const __nv_bfloat16* aux;
    long long ldaux;
    float* colsum;
    // grouped / convolution-walk generalisation (defaults: groups = 1, a_kstep = 64, offsets 0, b_box_mn_step = 64)
    int groups, tiles_per_group;
    int a_kstep;                    // K-major A: k-coordinate advance per k-block (a conv tap walk uses the row pitch)
    int a_g_mn, a_g_k, b_g_mn, b_g_k;   // per-group offsets of the TMA coordinates
    int b_box_mn_step, b_box_k_shift;   // MN-major B: per 64-wide box j of a tile: mn += step, k += shift * (global box)
    int out_g_row, out_g_col;       // per-group offset of the output block (also bias / colsum / resid / aux columns)
    int epilogue;
    int accumulate;
    float alpha;
    int* sched;                     // dynamic tile scheduler workspace {next tile, finished CTAs} (NULL = static round robin)
    int tma_epi;                    // bf16 outputs leave through TMA stores (epilogue_loop_tma); see the host-side conditions
    int debug;                      // TAVK_GEMM_DEBUG (measurement only): 1 = drain TMEM and drop the tile, 2 = no global stores,
                                    // 3 = stores folded onto 128 rows (no DRAM write-back)
};

// ---------------------------------------------------------------------------------------------- epilogue
// One epilogue warp owns 32 accumulator rows (its TMEM lane quarter) x every second 32-column chunk of each tile the
// CTA visits.  Per chunk: TMEM -> registers (lane = row) -> XOR-swizzled 4 KB smem patch -> registers in the COALESCED
// layout (8 lanes x 4 columns cover one row's chunk, 4 rows per instruction), where bias / row-bias / residual /
// GELU / GELU' / column sums are applied and the result stored: each global access of the warp touches 4 full
// 128-byte (fp32) or 64-byte (bf16) row segments instead of 32 rows x 16 B.
// The chunk's global operands (residual + row-bias, or the saved pre-activation) are software-pipelined one chunk
// ahead (across tile boundaries too) in two register buffers, and the per-row work is branch-free (predicated loads
// and stores only) so the 16 independent GELU polynomial chains of a chunk interleave.
// Where a CTA's tiles come from.  Static: tile i of this CTA = blockIdx.x + i * gridDim.x.  Dynamic (GemmDev::sched != NULL):
// the producer lane claims tiles from a global counter (atomicAdd) and publishes them, one tile ahead of the one it is
// loading, through a 4-deep shared-memory ring that the MMA warp and the 8 epilogue warps read.  With one persistent CTA
// per SM and a static split, a CTA that starts late — its SM was still busy with an NCCL all-reduce kernel or with another
// stream's kernel (tav.branch_streams) — finishes late and the whole grid waits for it; with the counter it simply claims
// fewer tiles.  A negative tile ends the sequence.
constexpr int kSchedDepth = 4;
struct TileSrc {
    int* sched;
    int* ring;
    uint64_t* full;
    uint64_t* empty;
    int num_tiles, lane;
    int first, stride;                      // static order: this CTA's (pair's) first tile and the number of CTAs (pairs)
    TAVK_DEVINL int get(int i) const {      // whole warp
        if (sched == nullptr) {
            const long long t = (long long)first + (long long)i * stride;
            return t < num_tiles ? (int)t : -1;
        }
        const int slot = i & (kSchedDepth - 1);
        mbar_wait(&full[slot], (uint32_t)((i / kSchedDepth) & 1));
        const int t = *reinterpret_cast<volatile int*>(ring + slot);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);
        return t;
    }
};

struct EpiItem {
    int row_base, col;      // first of this warp's 32 rows; this lane's first of 4 columns (global output coordinates)
    int rows_valid;         // how many of the 32 rows lie inside the (group's) M
    bool col_ok, lead;      // column group inside the (group's) N; first K-split (applies bias / row-bias / residual)
    bool full;              // WARP-UNIFORM: all 32 rows and all 32 columns of the chunk are inside the output
    int rb_group, rb_next;  // row-bias group of row_base and the first row of the next group (one division per TILE)
};

enum { OUT_F32 = 0, OUT_BF16 = 1, OUT_RED = 2 };   // plain fp32 store | bf16 store | fp32 red.global.add (split-K / +=)

template <int MODE>
struct EpiOperands {        // what one chunk needs from global memory besides the accumulator
    float4 b4;
    float4 res[MODE == TAVK_EPI_LINEAR ? 8 : 1];
    uint2 aux[(MODE == TAVK_EPI_GELU_BWD || MODE == TAVK_EPI_MUL) ? 8 : 1];
};

// FULL: the chunk lies entirely inside the output (32 valid rows, valid column group) — the common case; every
// predicate folds away at compile time (the predicated version spent ~100 ISETP and as many address selects per chunk).
template <int MODE, bool FULL>
TAVK_DEVINL void epi_issue_loads(const GemmDev& p, const EpiItem& w, int rsub, EpiOperands<MODE>& o) {
    o.b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias != nullptr && w.lead && (FULL || w.col_ok)) o.b4 = __ldg(reinterpret_cast<const float4*>(p.bias + w.col));
    if (MODE == TAVK_EPI_LINEAR) {
        const bool any = FULL || w.rows_valid > 0;      // a CTA pair's lower half may lie entirely past M
        const bool use_res = p.resid != nullptr && w.lead && any;
        const bool use_rb = p.rowbias != nullptr && w.lead && any;
        if (use_res) {
            const float* rp = p.resid + (long long)(w.row_base + rsub) * p.ldr + w.col;
            const long long rstep = 4 * p.ldr;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                o.res[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (FULL || ((i * 4 + rsub) < w.rows_valid && w.col_ok)) o.res[i] = ld_global_nc_v4(rp);
                rp += rstep;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) o.res[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (use_rb) {
            // the chunk's 32 rows touch the group of its first row and, past rb_next, the following one(s)
            const float* g0 = p.rowbias + (long long)w.rb_group * p.N + w.col;
            if (w.row_base + 32 <= w.rb_next) {            // uniform: one group for the whole chunk
                if (FULL || w.col_ok) {
                    const float4 q4 = __ldg(reinterpret_cast<const float4*>(g0));
#pragma unroll
                    for (int i = 0; i < 8; ++i) { o.res[i].x += q4.x; o.res[i].y += q4.y; o.res[i].z += q4.z; o.res[i].w += q4.w; }
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int row = w.row_base + i * 4 + rsub;
                    if (FULL || ((i * 4 + rsub) < w.rows_valid && w.col_ok)) {
                        const int gi = row < w.rb_next ? w.rb_group : row / p.rows_per_group;
                        const float4 q4 = __ldg(reinterpret_cast<const float4*>(p.rowbias + (long long)gi * p.N + w.col));
                        o.res[i].x += q4.x; o.res[i].y += q4.y; o.res[i].z += q4.z; o.res[i].w += q4.w;
                    }
                }
            }
        }
    }
    if (MODE == TAVK_EPI_GELU_BWD || MODE == TAVK_EPI_MUL) {
        const __nv_bfloat16* ap = p.aux + (long long)(w.row_base + rsub) * p.ldaux + w.col;
        const long long astep = 4 * p.ldaux;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            o.aux[i] = make_uint2(0u, 0u);
            if (FULL || ((i * 4 + rsub) < w.rows_valid && w.col_ok)) o.aux[i] = ld_global_nc_v2(ap);
            ap += astep;
        }
    }
}

template <int MODE, int OUT, bool FULL>
TAVK_DEVINL void epi_process(const GemmDev& p, const EpiItem& w, const EpiOperands<MODE>& o, uint32_t taddr, uint32_t stg,
                             int lane, uint64_t* release_bar, uint32_t release_remote) {
    const int cc = lane & 7, rsub = lane >> 3;
    uint32_t r[32];
    tmem_ld_32x32(taddr, r);
    tmem_ld_wait();
    if (release_bar != nullptr) {
        // last chunk of the tile for this warp: hand the TMEM buffer back to the MMA issuer before the global stores
        // (release_remote: this is the peer CTA of a pair — the issuer's barrier lives in the leader's shared memory)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
            if (release_remote != 0) mbar_arrive_cluster(release_remote);
            else mbar_arrive(release_bar);
        }
    }
    if (p.debug == 1) return;       // measurement: main loop + TMEM drain only
    const uint32_t st_addr = stg + lane * 128;
#pragma unroll
    for (int j = 0; j < 8; ++j)
        st_shared_v4(st_addr + ((j ^ (lane & 7)) << 4), r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
    __syncwarp();
    // this lane's first output element; rows advance by 4 per step (pointer increments instead of 64-bit multiplies)
    constexpr int kEsz = (OUT == OUT_BF16) ? 2 : 4;
    // debug 3 (measurement only): same stores, but every tile lands on the first 128 rows (L2-resident: no DRAM write-back)
    const int orow = (p.debug == 3 ? (w.row_base & 127) : w.row_base) + rsub;
    char* op = reinterpret_cast<char*>(p.out) + ((long long)orow * p.ldo + w.col) * kEsz;
    const long long ostep = 4 * p.ldo * kEsz;
    char* op2 = nullptr;
    long long ostep2 = 0;
    constexpr bool kGelu = (MODE == TAVK_EPI_GELU || MODE == TAVK_EPI_GELU_GRAD);
    if (kGelu) {
        op2 = reinterpret_cast<char*>(p.out2) + ((long long)orow * p.ldo2 + w.col) * 2;
        ostep2 = 4 * p.ldo2 * 2;
    }
    const f32x2 alpha2 = pk(p.alpha, p.alpha), blo = pk(o.b4.x, o.b4.y), bhi = pk(o.b4.z, o.b4.w);
    f32x2 cslo = pk(0.f, 0.f), cshi = pk(0.f, 0.f);
    const bool want_cs = !kGelu && p.colsum != nullptr;     // uniform
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int rl = i * 4 + rsub;
        const bool ok = (FULL || (rl < w.rows_valid && w.col_ok)) && p.debug != 2;
        // row rl, 16-byte slot (cc ^ (rl & 7)): rl & 7 = rsub | ((i & 1) << 2)  (rsub < 4)
        float4 v = ld_shared_v4(stg + rl * 128 + ((cc ^ (rsub | ((i & 1) << 2))) << 4));
        unpk(fma2(pk(v.x, v.y), alpha2, blo), v.x, v.y);
        unpk(fma2(pk(v.z, v.w), alpha2, bhi), v.z, v.w);
        if (kGelu) {
            // out = pre-activation (GELU) or gelu'(pre) (GELU_GRAD), out2 = gelu(pre); all bf16
            float4 g;
            if (MODE == TAVK_EPI_GELU) {
                gelu_fast2(v.x, v.y, g.x, g.y);
                gelu_fast2(v.z, v.w, g.z, g.w);
            } else {
                float4 dgl;
                gelu_and_grad2(v.x, v.y, g.x, g.y, dgl.x, dgl.y);
                gelu_and_grad2(v.z, v.w, g.z, g.w, dgl.z, dgl.w);
                v = dgl;
            }
            uint2 a, gg;
            a.x = pack_bf16x2(v.x, v.y);  a.y = pack_bf16x2(v.z, v.w);
            gg.x = pack_bf16x2(g.x, g.y); gg.y = pack_bf16x2(g.z, g.w);
            if (ok) {
                *reinterpret_cast<uint2*>(op) = a;
                *reinterpret_cast<uint2*>(op2) = gg;
            }
            op2 += ostep2;
        } else {
            if (MODE == TAVK_EPI_GELU_BWD) {
                // out = acc * gelu'(aux)
                const float2 a0 = unpack_bf16x2(o.aux[i].x), a1 = unpack_bf16x2(o.aux[i].y);
                gelu_grad_mul2(a0.x, a0.y, v.x, v.y);
                gelu_grad_mul2(a1.x, a1.y, v.z, v.w);
            } else if (MODE == TAVK_EPI_MUL) {
                const float2 a0 = unpack_bf16x2(o.aux[i].x), a1 = unpack_bf16x2(o.aux[i].y);
                v.x *= a0.x; v.y *= a0.y; v.z *= a1.x; v.w *= a1.y;
            } else {
                v.x += o.res[i].x; v.y += o.res[i].y; v.z += o.res[i].z; v.w += o.res[i].w;
            }
            if (ok) {
                if (OUT == OUT_BF16) {
                    uint2 a;
                    a.x = pack_bf16x2(v.x, v.y); a.y = pack_bf16x2(v.z, v.w);
                    *reinterpret_cast<uint2*>(op) = a;
                } else if (OUT == OUT_RED) {
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(op), "f"(v.x), "f"(v.y), "f"(v.z),
                                 "f"(v.w)
                                 : "memory");
                } else {
                    *reinterpret_cast<float4*>(op) = v;
                }
            }
            if (want_cs) {      // rows past M contribute nothing
                cslo = add2(cslo, pk(ok ? v.x : 0.f, ok ? v.y : 0.f));
                cshi = add2(cshi, pk(ok ? v.z : 0.f, ok ? v.w : 0.f));
            }
        }
        op += ostep;
    }
    if (want_cs) {
        // column sums of the stored values (bias gradient of the Linear whose output gradient this GEMM produces)
        float cs0, cs1, cs2, cs3;
        unpk(cslo, cs0, cs1);
        unpk(cshi, cs2, cs3);
        cs0 += __shfl_xor_sync(0xffffffffu, cs0, 8);  cs1 += __shfl_xor_sync(0xffffffffu, cs1, 8);
        cs2 += __shfl_xor_sync(0xffffffffu, cs2, 8);  cs3 += __shfl_xor_sync(0xffffffffu, cs3, 8);
        cs0 += __shfl_xor_sync(0xffffffffu, cs0, 16); cs1 += __shfl_xor_sync(0xffffffffu, cs1, 16);
        cs2 += __shfl_xor_sync(0xffffffffu, cs2, 16); cs3 += __shfl_xor_sync(0xffffffffu, cs3, 16);
        if (lane < 8 && (FULL || w.col_ok))
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.colsum + w.col), "f"(cs0), "f"(cs1),
                         "f"(cs2), "f"(cs3)
                         : "memory");
    }
    __syncwarp();   // the staging patch is rewritten by the next chunk
}

template <int BLOCK_N, int MODE, int OUT>
TAVK_DEVINL void epilogue_loop(const GemmDev& p, uint32_t tmem_base, uint32_t stg, int lane, int quarter, int half,
                               const TileSrc& src, uint64_t* tmem_full_bar, uint64_t* tmem_empty_bar, int row_off,
                               uint32_t empty_remote) {
    constexpr int kPer = BLOCK_N / 64;      // chunks per warp per tile (1, 2 or 4)
    const int cc = lane & 7, rsub = lane >> 3;
    // per TILE (integer divisions live here, not in the per-chunk path): everything but the chunk's column
    struct TileInfo { int row_base, rows_valid, col0, cols_left, rb_group, rb_next; bool lead; };
    auto tile_info = [&](int tile) {
        TileInfo t;
        const int g = tile / p.tiles_per_group;
        const int tg = tile - g * p.tiles_per_group;
        const int mn = tg / p.k_splits;
        const int m_blk = mn / p.num_n_blocks;
        const int row_local = m_blk * p.m_tile + row_off + quarter * 32;     // row_off: 128 in the peer CTA of a pair
        const int col_chunk = (mn - m_blk * p.num_n_blocks) * BLOCK_N + half * 32;      // warp-uniform
        t.lead = (tg - mn * p.k_splits) == 0;
        t.rows_valid = p.M - row_local;
        t.row_base = g * p.out_g_row + row_local;
        t.col0 = g * p.out_g_col + col_chunk + cc * 4;
        t.cols_left = p.N - col_chunk;       // columns left from the first chunk's first column (warp-uniform)
        t.rb_group = 0;
        t.rb_next = 0x7fffffff;
        if (MODE == TAVK_EPI_LINEAR && p.rowbias != nullptr) {
            t.rb_group = t.row_base / p.rows_per_group;
            t.rb_next = (t.rb_group + 1) * p.rows_per_group;
        }
        return t;
    };
    auto item = [&](const TileInfo& t, int ci) {
        EpiItem w;
        w.row_base = t.row_base; w.rows_valid = t.rows_valid; w.lead = t.lead;
        w.col = t.col0 + ci * 64;
        w.col_ok = ci * 64 + cc * 4 < t.cols_left;       // N % 8 == 0: a 4-column group is in or out as a whole
        // the fast path contains warp-synchronous instructions (tcgen05.ld, __syncwarp): its condition must not depend
        // on the lane
        w.full = t.rows_valid >= 32 && ci * 64 + 32 <= t.cols_left;
        w.rb_group = t.rb_group; w.rb_next = t.rb_next;
        return w;
    };
    // The work of this warp is the flat sequence of (tile, chunk) items; the global operands of item i+1 are
    // requested before item i is processed (one chunk ahead, across tile boundaries too).
    int i_tile = 0, ci = 0, it = 0;
    int tile = src.get(0);
    if (tile < 0) return;
    int tile_nxt = src.get(1);              // one tile of lookahead (published before the current tile's loads start)
    TileInfo tcur = tile_info(tile);
    EpiItem w_cur = item(tcur, 0), w_nxt = w_cur;
    EpiOperands<MODE> op_cur, op_nxt;
    if (w_cur.full) epi_issue_loads<MODE, true>(p, w_cur, rsub, op_cur);
    else epi_issue_loads<MODE, false>(p, w_cur, rsub, op_cur);
#pragma unroll 1
    while (true) {
        int ntile = tile, nci = ci + 1;
        if (nci == kPer) { nci = 0; ntile = tile_nxt; }
        const bool more = ntile >= 0;
        if (more) {
            if (nci == 0) tcur = tile_info(ntile);
            w_nxt = item(tcur, nci);
            if (w_nxt.full) epi_issue_loads<MODE, true>(p, w_nxt, rsub, op_nxt);
            else epi_issue_loads<MODE, false>(p, w_nxt, rsub, op_nxt);
        }
        const int acc = it & 1;
        if (ci == 0) {
            mbar_wait(&tmem_full_bar[acc], (uint32_t)((it >> 1) & 1));
            tc_fence_after();
        }
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BLOCK_N + (half + 2 * ci) * 32);
        const bool last = (ci == kPer - 1);
        uint64_t* rel = last ? &tmem_empty_bar[acc] : nullptr;
        const uint32_t rel_remote = empty_remote != 0 ? empty_remote + 8u * (uint32_t)acc : 0u;
        if (w_cur.full) epi_process<MODE, OUT, true>(p, w_cur, op_cur, taddr, stg, lane, rel, rel_remote);
        else epi_process<MODE, OUT, false>(p, w_cur, op_cur, taddr, stg, lane, rel, rel_remote);
        if (last) ++it;
        if (!more) break;
        if (nci == 0) {                     // advanced to the next tile: fetch the one after it
            ++i_tile;
            tile_nxt = src.get(i_tile + 1);
        }
        tile = ntile; ci = nci;
        w_cur = w_nxt;
        op_cur = op_nxt;
    }
}

// ---------------------------------------------------------------------------------------------- TMA-store epilogue
// bf16 outputs without residual / row-bias (QKV, FFN-up + GELU, the FFN-up dgrad, attention-output dgrad, the conv
// feature-encoder layers).  Measured on the coalesced-store epilogue above (TAVK_GEMM_DEBUG, profiles/r2_gemm_epilogue_*):
// at M = 23424, N = 3072, K = 768 the main loop + TMEM drain takes 70 us, staging + GELU math 11 us more, and the st.global
// instructions another 40 us even when they never reach DRAM — 8-byte-per-lane bf16 stores move 64-byte row pieces, and the
// LSU path, not HBM, paces the epilogue.  Here the warps never issue a global store: everything elementwise happens in the
// accumulator layout (lane = row, 32 consecutive columns in registers; bias arrives as warp-uniform loads), the bf16
// rows go to a 2 KB SWIZZLE_64B patch (half the shared-memory traffic of the fp32 transpose patch) and one lane hands the
// patch to the TMA unit, which writes full lines asynchronously and clips the M / N edges itself.  The GELU' multiplier
// of TAVK_EPI_MUL / GELU_BWD comes in the same way (TMA load of the 32x32 bf16 box, one chunk ahead).
TAVK_DEVINL uint32_t sw64_off(int row, int chunk16) {      // byte offset of 16-byte chunk `chunk16` of 64-byte row `row`
    return (uint32_t)((row >> 3) * 512 + (row & 7) * 64 + ((chunk16 ^ ((row >> 1) & 3)) << 4));
}

template <int BLOCK_N, int MODE>
TAVK_DEVINL void epilogue_loop_tma(const GemmDev& p, const CUtensorMap* tmap_out, const CUtensorMap* tmap_out2,
                                   const CUtensorMap* tmap_aux, uint32_t tmem_base, uint8_t* stg, uint64_t* aux_bar, int lane,
                                   int quarter, int half, const TileSrc& src, uint64_t* tmem_full_bar, uint64_t* tmem_empty_bar,
                                   int row_off, uint32_t empty_remote) {
    constexpr int kPer = BLOCK_N / 64;      // chunks per warp per tile (1, 2 or 4)
    constexpr bool kGelu = (MODE == TAVK_EPI_GELU || MODE == TAVK_EPI_GELU_GRAD);
    constexpr bool kAux = (MODE == TAVK_EPI_GELU_BWD || MODE == TAVK_EPI_MUL);
    uint8_t* buf0 = stg;                    // out
    uint8_t* buf1 = stg + 2048;             // out2 (GELU modes) | aux (MUL / GELU_BWD) | second out buffer (LINEAR)
    auto coords = [&](int tile, int ci, int& row0, int& col0) {
        const int mn = tile / p.k_splits;                    // k_splits == 1 on this path, groups == 1
        const int m_blk = mn / p.num_n_blocks;
        row0 = m_blk * p.m_tile + row_off + quarter * 32;
        col0 = (mn - m_blk * p.num_n_blocks) * BLOCK_N + (half + 2 * ci) * 32;
    };
    int i_tile = 0, ci = 0, it = 0, n_item = 0;
    int tile = src.get(0);
    if (tile < 0) return;
    int tile_nxt = src.get(1);
    int row0, col0;
    coords(tile, 0, row0, col0);
    uint32_t aux_phase = 0;
    if (kAux && lane == 0) {
        mbar_arrive_expect_tx(aux_bar, 2048);
        tma_load_2d(buf1, tmap_aux, aux_bar, col0, row0);
    }
#pragma unroll 1
    while (true) {
        int ntile = tile, nci = ci + 1;
        if (nci == kPer) { nci = 0; ntile = tile_nxt; }
        const bool more = ntile >= 0;
        int nrow0 = 0, ncol0 = 0;
        if (more) coords(ntile, nci, nrow0, ncol0);
        // bias of the chunk's 32 columns: the same addresses in every lane (broadcast loads), issued before the waits
        float bias[32];
        if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (col0 + 4 * j < p.N) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j);   // N % 8 == 0
                bias[4 * j] = b4.x; bias[4 * j + 1] = b4.y; bias[4 * j + 2] = b4.z; bias[4 * j + 3] = b4.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) bias[j] = 0.f;
        }
        const int acc = it & 1;
        if (ci == 0) {
            mbar_wait(&tmem_full_bar[acc], (uint32_t)((it >> 1) & 1));
            tc_fence_after();
        }
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BLOCK_N + (half + 2 * ci) * 32), r);
        tmem_ld_wait();
        if (ci == kPer - 1) {
            // last chunk of the tile for this warp: hand the TMEM buffer back to the MMA issuer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (empty_remote != 0) mbar_arrive_cluster(empty_remote + 8u * (uint32_t)acc);
                else mbar_arrive(&tmem_empty_bar[acc]);
            }
            ++it;
        }
        if (p.debug == 1) {
            if (!more) break;
            if (nci == 0) { ++i_tile; tile_nxt = src.get(i_tile + 1); }
            tile = ntile; ci = nci; row0 = nrow0; col0 = ncol0;
            continue;
        }
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaf(__uint_as_float(r[j]), p.alpha, bias[j]);
        uint32_t o[16], o2[16];
        if (kAux) {
            // this chunk's multiplier box has landed in buf1: read my row, then let the next box overwrite it
            mbar_wait(aux_bar, aux_phase);
            aux_phase ^= 1;
            uint32_t a[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float4 t = ld_shared_v4(smem_u32(buf1) + sw64_off(lane, c));
                a[4 * c] = __float_as_uint(t.x); a[4 * c + 1] = __float_as_uint(t.y);
                a[4 * c + 2] = __float_as_uint(t.z); a[4 * c + 3] = __float_as_uint(t.w);
            }
            fence_proxy_async_shared();     // generic reads of buf1 before the async-proxy write of the next box
            __syncwarp();
            if (more && lane == 0) {
                mbar_arrive_expect_tx(aux_bar, 2048);
                tma_load_2d(buf1, tmap_aux, aux_bar, ncol0, nrow0);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float2 m2 = unpack_bf16x2(a[j]);
                if (MODE == TAVK_EPI_MUL) {
                    v[2 * j] *= m2.x; v[2 * j + 1] *= m2.y;
                } else {
                    gelu_grad_mul2(m2.x, m2.y, v[2 * j], v[2 * j + 1]);
                }
            }
        }
        if (kGelu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float g0, g1;
                if (MODE == TAVK_EPI_GELU) {
                    gelu_fast2(v[2 * j], v[2 * j + 1], g0, g1);
                    o[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
                } else {
                    float d0, d1;
                    gelu_and_grad2(v[2 * j], v[2 * j + 1], g0, g1, d0, d1);
                    o[j] = pack_bf16x2(d0, d1);
                }
                o2[j] = pack_bf16x2(g0, g1);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
        }
        // staging buffers: LINEAR alternates buf0 / buf1 (one store may still be reading the other); the two-output and
        // aux modes own one buffer per role and wait for the previous chunk's store to have read it
        uint8_t* ob = (!kGelu && !kAux && (n_item & 1)) ? buf1 : buf0;
        if (lane == 0) {
            if (!kGelu && !kAux) tma_store_wait_read<1>();
            else tma_store_wait_read<0>();
        }
        __syncwarp();
        if (p.debug != 2) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
                st_shared_v4(smem_u32(ob) + sw64_off(lane, c), o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
            if (kGelu) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    st_shared_v4(smem_u32(buf1) + sw64_off(lane, c), o2[4 * c], o2[4 * c + 1], o2[4 * c + 2], o2[4 * c + 3]);
            }
            fence_proxy_async_shared();
            __syncwarp();
            if (lane == 0 && row0 < p.M && col0 < p.N) {
                tma_store_2d(tmap_out, ob, col0, row0);
                if (kGelu) tma_store_2d(tmap_out2, buf1, col0, row0);
                tma_store_commit();
            }
        }
        if (!kGelu && p.colsum != nullptr) {
            // column sums of the stored values (bias gradient of the Linear whose output gradient this GEMM produces)
            if (row0 + lane >= p.M) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0.f;
            }
            const float cs = warp_colsum32(v, lane);
            if (col0 + lane < p.N) atomicAdd(p.colsum + col0 + lane, cs);
        }
        ++n_item;
        if (!more) break;
        if (nci == 0) { ++i_tile; tile_nxt = src.get(i_tile + 1); }
        tile = ntile; ci = nci; row0 = nrow0; col0 = ncol0;
    }
    if (lane == 0) tma_store_wait<0>();     // shared memory must outlive the last stores' reads; make them complete
    __syncwarp();
}

template <int BLOCK_N, bool A_MN, bool B_MN, int CTAS>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_out2,
                         const __grid_constant__ CUtensorMap tmap_aux, const GemmDev p) {
    using Cfg = GemmCfg<BLOCK_N, CTAS>;
    constexpr bool kPair = CTAS == 2;
    const uint32_t cta_rank = kPair ? cluster_ctarank() : 0u;      // 0 = leader (issues the MMAs of the pair)
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment.
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + Cfg::kStages * Cfg::kABytes;
    uint8_t* smem_stage = smem + Cfg::kStages * Cfg::kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stage + Cfg::kStagingBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + Cfg::kStages;
    uint64_t* tmem_full_bar = bars + 2 * Cfg::kStages;
    uint64_t* tmem_empty_bar = bars + 2 * Cfg::kStages + 2;
    uint64_t* aux_bar = bars + 2 * Cfg::kStages + 4;          // one per epilogue warp (TMA-store epilogue, aux boxes)
    uint64_t* sched_full = aux_bar + kNumEpiWarps;            // dynamic tile scheduler ring (TileSrc)
    uint64_t* sched_empty = sched_full + kSchedDepth;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sched_empty + kSchedDepth);
    int* sched_ring = reinterpret_cast<int*>(tmem_ptr_smem + 4);

    const int warp_idx = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform for the compiler
    const int lane = threadIdx.x & 31;

    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int i = 0; i < Cfg::kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], kNumEpiWarps * CTAS);     // pair: the peer's epilogue warps arrive remotely
        }
        for (int i = 0; i < kNumEpiWarps; ++i) mbar_init(&aux_bar[i], 1);
        for (int i = 0; i < kSchedDepth; ++i) {
            mbar_init(&sched_full[i], 1);
            mbar_init(&sched_empty[i], 1 + kNumEpiWarps);       // the MMA warp and every epilogue warp read each entry
        }
        if (p.tma_epi) {
            tma_prefetch_desc(&tmap_out);
            if (p.out2 != nullptr) tma_prefetch_desc(&tmap_out2);
            if (p.aux != nullptr) tma_prefetch_desc(&tmap_aux);
        }
        mbar_fence_init();
    }
    if (warp_idx == 1) {
        if constexpr (kPair) tmem_alloc_pair<Cfg::kTmemCols>(tmem_ptr_smem);
        else tmem_alloc<Cfg::kTmemCols>(tmem_ptr_smem);
    }
    tc_fence_before();
    if constexpr (kPair) cluster_sync_all();    // the peer's barriers exist before anything signals them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);

    const int num_tiles = p.groups * p.tiles_per_group;
    const TileSrc src{kPair ? nullptr : p.sched, sched_ring, sched_full, sched_empty, num_tiles, lane,
                      kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x};
    const int row_off = (int)cta_rank * kBlockM;
    pdl_trigger();   // the next kernel on the stream may start its own prologue ...
    pdl_wait();      // ... and this one touches global memory only after its predecessor has completed

    if (warp_idx == 0) {
        // ===================== TMA producer (one thread) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            // tile i of this CTA: static round robin, or claimed from the global counter and published one tile ahead
            auto claim = [&](int i) -> int {
                if (src.sched == nullptr) {
                    const long long t = (long long)src.first + (long long)i * src.stride;
                    return t < num_tiles ? (int)t : -1;
                }
                const int c = atomicAdd(p.sched, 1);
                const int t = c < num_tiles ? c : -1;
                const int slot = i & (kSchedDepth - 1);
                mbar_wait(&sched_empty[slot], (uint32_t)(((i / kSchedDepth) & 1) ^ 1));
                *reinterpret_cast<volatile int*>(sched_ring + slot) = t;
                mbar_arrive(&sched_full[slot]);
                return t;
            };
            int tile = claim(0);
            for (int i_tile = 0; tile >= 0; ++i_tile) {
                const int tile_after = claim(i_tile + 1);
                const int g = tile / p.tiles_per_group;
                const int tg = tile - g * p.tiles_per_group;
                const int split = tg % p.k_splits;
                const int mn = tg / p.k_splits;
                const int n_blk = mn % p.num_n_blocks;
                const int m_blk = mn / p.num_n_blocks;
                const int kb0 = split * p.kb_per_split;
                const int kb1 = min(kb0 + p.kb_per_split, p.num_kb);
                const int a_mn0 = m_blk * p.m_tile + row_off + g * p.a_g_mn, a_k0 = g * p.a_g_k;
                const int b_k0 = g * p.b_g_k;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem_a + stage * Cfg::kABytes;
                    uint8_t* sb = smem_b + stage * Cfg::kBBytes;
                    if constexpr (kPair) {
                        // both CTAs' loads of this stage complete on the leader's barrier, armed with the bytes of both
                        if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
                        const uint32_t fb = mapa_shared(smem_u32(&full_bar[stage]), 0);
                        if constexpr (A_MN) {
#pragma unroll
                            for (int j = 0; j < kBlockM / 64; ++j)
                                tma_load_2d_pair(sa + j * 8192, &tmap_a, fb, a_mn0 + j * 64, a_k0 + kb * kBlockK);
                        } else {
                            tma_load_2d_pair(sa, &tmap_a, fb, a_k0 + kb * kBlockK, a_mn0);
                        }
                        const int b_mn0 = n_blk * BLOCK_N + (int)cta_rank * Cfg::kBRows;    // this CTA's half of the B tile
                        if constexpr (B_MN) {
#pragma unroll
                            for (int j = 0; j < Cfg::kBRows / 64; ++j)
                                tma_load_2d_pair(sb + j * 8192, &tmap_b, fb, b_mn0 + j * 64, b_k0 + kb * kBlockK);
                        } else {
                            tma_load_2d_pair(sb, &tmap_b, fb, b_k0 + kb * kBlockK, b_mn0);
                        }
                        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                    if constexpr (A_MN) {
#pragma unroll
                        for (int j = 0; j < kBlockM / 64; ++j)
                            tma_load_2d(sa + j * 8192, &tmap_a, &full_bar[stage], a_mn0 + j * 64, a_k0 + kb * kBlockK);
                    } else {
                        tma_load_2d(sa, &tmap_a, &full_bar[stage], a_k0 + kb * p.a_kstep, a_mn0);
                    }
                    if constexpr (B_MN) {
#pragma unroll
                        for (int j = 0; j < BLOCK_N / 64; ++j) {
                            const int box = n_blk * (BLOCK_N / 64) + j;
                            tma_load_2d(sb + j * 8192, &tmap_b, &full_bar[stage], box * p.b_box_mn_step + g * p.b_g_mn,
                                        b_k0 + kb * kBlockK + box * p.b_box_k_shift);
                        }
                    } else {
                        tma_load_2d(sb, &tmap_b, &full_bar[stage], b_k0 + kb * kBlockK, n_blk * BLOCK_N + g * p.b_g_mn);
                    }
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                }
                tile = tile_after;
            }
            if (src.sched != nullptr) {
                // last CTA to finish claiming resets the workspace for the next launch that uses it
                __threadfence();
                if (atomicAdd(p.sched + 1, 1) == (int)gridDim.x - 1) {
                    p.sched[0] = 0;
                    p.sched[1] = 0;
                    __threadfence();
                }
            }
        }
    } else if (warp_idx == 1 && cta_rank == 0) {
        // ===================== MMA issuer: the whole warp walks the (warp-uniform) loop, one elected lane issues ====
        // (of a CTA pair only the leader's warp 1 comes here; the peer's only allocates and frees its tensor memory)
        const bool leader = elect_one();
        constexpr uint32_t idesc = umma_idesc_bf16(kBlockM * CTAS, BLOCK_N, A_MN, B_MN);
        const uint64_t da0 = A_MN ? umma_smem_desc(smem_u32(smem_a), 8192, 1024) : umma_smem_desc(smem_u32(smem_a), 16, 1024);
        const uint64_t db0 = B_MN ? umma_smem_desc(smem_u32(smem_b), 8192, 1024) : umma_smem_desc(smem_u32(smem_b), 16, 1024);
        constexpr uint64_t kAStep = A_MN ? (2048 >> 4) : (32 >> 4);   // descriptor address units (16 B) per UMMA_K
        constexpr uint64_t kBStep = B_MN ? (2048 >> 4) : (32 >> 4);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int tile = src.get(0); tile >= 0; tile = src.get(++it)) {
            const int split = (tile % p.tiles_per_group) % p.k_splits;
            const int kb0 = split * p.kb_per_split;
            const int kb1 = min(kb0 + p.kb_per_split, p.num_kb);
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            if constexpr (kPair) mbar_wait_cluster(&tmem_empty_bar[acc], acc_phase ^ 1);
            else mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (leader) {
                    const uint64_t da = da0 + (uint64_t)(stage * (Cfg::kABytes >> 4));
                    const uint64_t db = db0 + (uint64_t)(stage * (Cfg::kBBytes >> 4));
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                        if constexpr (kPair) umma_bf16_pair(tmem_d, da + k * kAStep, db + k * kBStep, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                        else umma_bf16(tmem_d, da + k * kAStep, db + k * kBStep, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    // frees the smem slot (in both CTAs of a pair) when these MMAs retire
                    if constexpr (kPair) umma_commit_pair(&empty_bar[stage]);
                    else umma_commit(&empty_bar[stage]);
                }
                __syncwarp();
                if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
            }
            if (leader) {   // accumulator complete -> epilogue (of both CTAs)
                if constexpr (kPair) umma_commit_pair(&tmem_full_bar[acc]);
                else umma_commit(&tmem_full_bar[acc]);
            }
            __syncwarp();
        }
    } else if (warp_idx >= 2) {
        // ===================== epilogue warps =====================
        const int ew = warp_idx - 2;            // 0..7
        const int quarter = warp_idx & 3;       // TMEM lane quarter this warp may access
        const int half = ew >> 2;               // which half of the column chunks
        const uint32_t stg = smem_u32(smem_stage) + ew * 4096;
        // peer CTA of a pair: the accumulator is released on the leader's tmem_empty barriers
        const uint32_t empty_remote = (kPair && cta_rank != 0) ? mapa_shared(smem_u32(&tmem_empty_bar[0]), 0) : 0u;
#define TAVK_EPI_TMA(MODE)                                                                                             \
    epilogue_loop_tma<BLOCK_N, MODE>(p, &tmap_out, &tmap_out2, &tmap_aux, tmem_base, smem_stage + ew * 4096, &aux_bar[ew], lane, \
                                     quarter, half, src, tmem_full_bar, tmem_empty_bar, row_off, empty_remote)
#define TAVK_EPI(MODE, OUT) \
    epilogue_loop<BLOCK_N, MODE, OUT>(p, tmem_base, stg, lane, quarter, half, src, tmem_full_bar, tmem_empty_bar, row_off, empty_remote)
        if (p.tma_epi) {
            if (p.epilogue == TAVK_EPI_GELU) TAVK_EPI_TMA(TAVK_EPI_GELU);
            else if (p.epilogue == TAVK_EPI_GELU_GRAD) TAVK_EPI_TMA(TAVK_EPI_GELU_GRAD);
            else if (p.epilogue == TAVK_EPI_MUL) TAVK_EPI_TMA(TAVK_EPI_MUL);
            else if (p.epilogue == TAVK_EPI_GELU_BWD) TAVK_EPI_TMA(TAVK_EPI_GELU_BWD);
            else TAVK_EPI_TMA(TAVK_EPI_LINEAR);
        } else if (p.epilogue == TAVK_EPI_GELU) TAVK_EPI(TAVK_EPI_GELU, OUT_BF16);
        else if (p.epilogue == TAVK_EPI_GELU_GRAD) TAVK_EPI(TAVK_EPI_GELU_GRAD, OUT_BF16);
        else if (p.epilogue == TAVK_EPI_MUL) TAVK_EPI(TAVK_EPI_MUL, OUT_BF16);
        else if (p.epilogue == TAVK_EPI_GELU_BWD) TAVK_EPI(TAVK_EPI_GELU_BWD, OUT_BF16);
        else if (p.out_bf16) TAVK_EPI(TAVK_EPI_LINEAR, OUT_BF16);
        else if (p.accumulate) TAVK_EPI(TAVK_EPI_LINEAR, OUT_RED);
        else TAVK_EPI(TAVK_EPI_LINEAR, OUT_F32);
#undef TAVK_EPI
#undef TAVK_EPI_TMA
    }

    tc_fence_before();
    // pair: neither CTA may leave (or free tensor memory) while the other can still signal its barriers, read its shared
    // memory through the MMA, or drain its half of the accumulator
    if constexpr (kPair) cluster_sync_all();
    else __syncthreads();
    if (warp_idx == 1) {
        __syncwarp();
        tc_fence_after();
        if constexpr (kPair) tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base);
        else tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
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

// 2-D bf16 tensor map over a row-major matrix [rows, cols] (cols contiguous, row pitch ld elements).
static int make_tmap_bf16(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_cols,
                          int box_rows, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    PFN_encodeTiled enc = get_encode_fn();
    TAVK_CHECK(enc != nullptr, 3, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TAVK_CHECK(r == CUDA_SUCCESS, 3, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld box=%dx%d base=%p",
               (int)r, rows, cols, ld, box_cols, box_rows, base);
    return 0;
}

template <int BLOCK_N, bool A_MN, bool B_MN, int CTAS>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& to2,
                       const CUtensorMap& tx, const GemmDev& dev, int grid, cudaStream_t stream) {
    using Cfg = GemmCfg<BLOCK_N, CTAS>;
    auto kern = gemm_bf16_tcgen05_kernel<BLOCK_N, A_MN, B_MN, CTAS>;
    static bool attr_done = false;
    if (!attr_done) {
        TAVK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
        attr_done = true;
    }
    if (CTAS == 2)
        TAVK_CUDA(launch_kernel_cluster(kern, dim3(grid), dim3(kGemmThreads), (size_t)Cfg::kSmemBytes, stream, 2u, ta, tb, to, to2, tx, dev));
    else
        TAVK_CUDA(launch_kernel(kern, dim3(grid), dim3(kGemmThreads), (size_t)Cfg::kSmemBytes, stream, ta, tb, to, to2, tx, dev));
    return 0;
}

}  // namespace tavk

using namespace tavk;

extern "C" int tavk_gemm_bf16(const tavk_gemm_args* a, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    TAVK_CHECK(a != nullptr, 1, "tavk_gemm_bf16: null args");
    TAVK_CHECK(a->M > 0 && a->N > 0 && a->K > 0, 1, "tavk_gemm_bf16: bad shape M=%d N=%d K=%d", a->M, a->N, a->K);
    TAVK_CHECK(a->N % 8 == 0, 1, "tavk_gemm_bf16: N=%d must be a multiple of 8", a->N);
    TAVK_CHECK(a->lda % 8 == 0 && a->ldb % 8 == 0, 1, "tavk_gemm_bf16: lda/ldb must be multiples of 8 elements");
    TAVK_CHECK((reinterpret_cast<uintptr_t>(a->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->B) & 15) == 0, 1,
               "tavk_gemm_bf16: operands must be 16-byte aligned");
    TAVK_CHECK(a->out != nullptr, 1, "tavk_gemm_bf16: null output");
    TAVK_CHECK(a->out_dtype == TAVK_F32 || a->out_dtype == TAVK_BF16, 1, "tavk_gemm_bf16: bad out_dtype");
    const int k_splits = a->k_splits < 1 ? 1 : a->k_splits;
    TAVK_CHECK(!(a->accumulate || k_splits > 1) || a->out_dtype == TAVK_F32, 1,
               "tavk_gemm_bf16: accumulate / split-K need an f32 output");
    TAVK_CHECK(k_splits == 1 || a->accumulate, 1, "tavk_gemm_bf16: split-K requires accumulate=1 (atomic adds)");
    TAVK_CHECK(a->epilogue >= TAVK_EPI_LINEAR && a->epilogue <= TAVK_EPI_MUL, 1, "tavk_gemm_bf16: bad epilogue %d", a->epilogue);
    const bool epi_gelu = a->epilogue == TAVK_EPI_GELU || a->epilogue == TAVK_EPI_GELU_GRAD;
    const bool epi_aux = a->epilogue == TAVK_EPI_GELU_BWD || a->epilogue == TAVK_EPI_MUL;
    TAVK_CHECK(!epi_gelu || (a->out_dtype == TAVK_BF16 && a->out2 != nullptr && k_splits == 1), 1,
               "tavk_gemm_bf16: GELU / GELU_GRAD epilogues need bf16 out + out2 and no split-K");
    TAVK_CHECK(!epi_aux || (a->aux != nullptr && k_splits == 1 && a->out_dtype == TAVK_BF16), 1,
               "tavk_gemm_bf16: GELU_BWD / MUL epilogues need aux, a bf16 output and no split-K");
    TAVK_CHECK(a->rowbias == nullptr || a->rows_per_group > 0, 1, "tavk_gemm_bf16: rowbias needs rows_per_group");
    const int vec = (a->out_dtype == TAVK_BF16) ? 8 : 4;
    TAVK_CHECK(a->ldo % vec == 0, 1, "tavk_gemm_bf16: ldo must be a multiple of %d", vec);

    const int groups = a->groups > 1 ? a->groups : 1;
    TAVK_CHECK(a->b_box_k_shift == 0 || a->b_mn_major, 1, "tavk_gemm_bf16: b_box_k_shift needs an MN-major B operand");
    TAVK_CHECK(a->a_kstep == 0 || (!a->a_mn_major && a->a_kstep % 8 == 0), 1,
               "tavk_gemm_bf16: a_kstep needs a K-major A operand and a multiple of 8 elements");

    // tile-shape heuristic: the widest tile whose wave quantisation is not noticeably worse than a narrower one's
    // per-call SM budget (a data-parallel host keeps a few SMs free for concurrently running NCCL kernels)
    int sms = (a->max_ctas > 0 && a->max_ctas < sm_count()) ? a->max_ctas : sm_count();
    const int mblocks = (a->M + kBlockM - 1) / kBlockM;
    auto waves_eff = [&](int bn) {
        const long long tiles = (long long)groups * mblocks * ((a->N + bn - 1) / bn) * k_splits;
        const long long waves = (tiles + sms - 1) / sms;
        return (double)tiles / (double)(waves * sms);
    };
    int block_n = 256;
    if (a->block_n == 64 || a->block_n == 128 || a->block_n == 256) block_n = a->block_n;
    else if (a->N <= 64) block_n = 64;
    // (a 128-wide tile runs at ~0.8x the rate of a 256-wide one — twice the A traffic per FLOP — so it has to win more than
    // that in wave quantisation: M = 5168, N = 2304 measured 24.0 us with 128-wide tiles, 20.9 us with 256-wide ones)
    else if (a->N <= 128 || waves_eff(128) > waves_eff(256) * 1.3) block_n = 128;
    // CTA pairs (cta_group::2, 256-row tiles): plain ungrouped problems with at least two row blocks.  The CTA count and
    // its wave quantisation are those of the single-CTA tiling (a pair = two CTAs = two 128-row blocks), except for the
    // idle lower half of the last pair when the number of row blocks is odd.  a->cta_pair: 0 heuristic, 1 never, 2 force.
    static const int pair_env = getenv("TAVK_GEMM_PAIR") ? atoi(getenv("TAVK_GEMM_PAIR")) : -1;     // 0 = off (measurement)
    const bool pair_ok = groups == 1 && a->a_kstep == 0 && a->b_box_k_shift == 0 && block_n >= 128 && mblocks >= 2 && sms >= 2;
    const int pair_req = pair_env >= 0 ? (pair_env ? 0 : 1) : a->cta_pair;
    // heuristic (tools/gemm_pair_probe.py, profiles/r2_gemm_pair_probe.txt): pairs win 5-15 % once the problem has at least
    // a wave of single-CTA tiles (and already at half a wave for the split-K wgrad shapes); below that the launch is one
    // partial wave either way and the pair's cluster scheduling + second barrier hop cost ~0.5 us
    const long long tiles1 = (long long)mblocks * ((a->N + block_n - 1) / block_n) * k_splits;
    const bool pair_auto = tiles1 >= sms || (a->a_mn_major && a->b_mn_major && 2 * tiles1 >= sms);
    const bool pair = pair_ok && pair_req != 1 && (pair_req == 2 || pair_auto);
    const int ctas = pair ? 2 : 1;
    if (pair) sms &= ~1;

    GemmDev d;
    d.M = a->M; d.N = a->N; d.K = a->K;
    d.m_tile = kBlockM * ctas;
    d.num_m_blocks = (a->M + d.m_tile - 1) / d.m_tile;
    d.num_n_blocks = (a->N + block_n - 1) / block_n;
    d.num_kb = (a->K + kBlockK - 1) / kBlockK;
    d.k_splits = k_splits > d.num_kb ? d.num_kb : k_splits;
    d.kb_per_split = (d.num_kb + d.k_splits - 1) / d.k_splits;
    d.k_splits = (d.num_kb + d.kb_per_split - 1) / d.kb_per_split;  // no empty splits
    d.out = a->out; d.ldo = a->ldo; d.out_bf16 = (a->out_dtype == TAVK_BF16);
    d.out2 = reinterpret_cast<__nv_bfloat16*>(a->out2); d.ldo2 = a->ldo2;
    d.bias = a->bias; d.resid = a->resid; d.ldr = a->ldr;
    d.rowbias = a->rowbias; d.rows_per_group = a->rows_per_group;
    d.aux = reinterpret_cast<const __nv_bfloat16*>(a->aux); d.ldaux = a->ldaux;
    d.colsum = a->colsum;
    d.epilogue = a->epilogue; d.accumulate = a->accumulate; d.alpha = a->alpha;
    static const int dbg = getenv("TAVK_GEMM_DEBUG") ? atoi(getenv("TAVK_GEMM_DEBUG")) : 0;
    d.debug = dbg;
    d.sched = pair ? nullptr : reinterpret_cast<int*>(a->sched_workspace);
    d.groups = groups;
    d.tiles_per_group = d.num_m_blocks * d.num_n_blocks * d.k_splits;
    d.a_kstep = a->a_kstep > 0 ? a->a_kstep : kBlockK;
    d.a_g_mn = a->a_g_mn; d.a_g_k = a->a_g_k; d.b_g_mn = a->b_g_mn; d.b_g_k = a->b_g_k;
    d.b_box_k_shift = a->b_box_k_shift;
    d.b_box_mn_step = a->b_box_k_shift != 0 ? 0 : 64;
    d.out_g_row = a->out_g_row; d.out_g_col = a->out_g_col;

    // tensor maps over the row-major matrices in memory ([rows, cols], pitch ld); explicit extents let a caller
    // describe overlapping rows (cols > ld: strided / sliding-window convolutions) and grouped problems
    CUtensorMap ta, tb;
    int rc;
    const long long a_rows = a->a_rows > 0 ? a->a_rows : (a->a_mn_major ? a->K : a->M);
    const long long a_cols = a->a_cols > 0 ? a->a_cols : (a->a_mn_major ? a->M : a->K);
    const long long b_rows = a->b_rows > 0 ? a->b_rows : (a->b_mn_major ? a->K : a->N);
    const long long b_cols = a->b_cols > 0 ? a->b_cols : (a->b_mn_major ? a->N : a->K);
    if (a->a_mn_major) rc = make_tmap_bf16(&ta, a->A, a_rows, a_cols, a->lda, 64, kBlockK);
    else               rc = make_tmap_bf16(&ta, a->A, a_rows, a_cols, a->lda, kBlockK, kBlockM);
    if (rc) return rc;
    if (a->b_mn_major) rc = make_tmap_bf16(&tb, a->B, b_rows, b_cols, a->ldb, 64, kBlockK);
    else               rc = make_tmap_bf16(&tb, a->B, b_rows, b_cols, a->ldb, kBlockK, block_n / ctas);
    if (rc) return rc;

    // TMA-store epilogue (see epilogue_loop_tma): plain bf16 outputs of one ungrouped problem
    static const int tma_off = getenv("TAVK_GEMM_TMA_EPI") ? (atoi(getenv("TAVK_GEMM_TMA_EPI")) == 0) : 0;
    d.tma_epi = (!tma_off && d.out_bf16 && d.groups == 1 && d.k_splits == 1 && !a->accumulate && a->resid == nullptr &&
                 a->rowbias == nullptr && (reinterpret_cast<uintptr_t>(a->out) & 15) == 0 && a->ldo % 8 == 0 &&
                 (a->out2 == nullptr || ((reinterpret_cast<uintptr_t>(a->out2) & 15) == 0 && a->ldo2 % 8 == 0)) &&
                 (a->aux == nullptr || ((reinterpret_cast<uintptr_t>(a->aux) & 15) == 0 && a->ldaux % 8 == 0)) &&
                 (a->bias == nullptr || (reinterpret_cast<uintptr_t>(a->bias) & 15) == 0))
                    ? 1 : 0;
    CUtensorMap to = ta, to2 = ta, tx = ta;      // placeholders when unused (never dereferenced)
    if (d.tma_epi) {
        rc = make_tmap_bf16(&to, a->out, a->M, a->N, a->ldo, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
        if (rc) return rc;
        if (a->out2 != nullptr) {
            rc = make_tmap_bf16(&to2, a->out2, a->M, a->N, a->ldo2, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
            if (rc) return rc;
        }
        if (epi_aux) {
            rc = make_tmap_bf16(&tx, a->aux, a->M, a->N, a->ldaux, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
            if (rc) return rc;
        }
    }

    const long long tiles = (long long)d.groups * d.tiles_per_group;        // pair mode: 256-row tiles, two CTAs each
    const int grid = (int)(tiles * ctas < sms ? tiles * ctas : sms);
#define TAVK_GEMM_DISPATCH(BN, CT)                                                                   \
    if (a->a_mn_major && a->b_mn_major) return launch_gemm<BN, true, true, CT>(ta, tb, to, to2, tx, d, grid, stream);   \
    if (a->a_mn_major && !a->b_mn_major) return launch_gemm<BN, true, false, CT>(ta, tb, to, to2, tx, d, grid, stream); \
    if (!a->a_mn_major && a->b_mn_major) return launch_gemm<BN, false, true, CT>(ta, tb, to, to2, tx, d, grid, stream); \
    return launch_gemm<BN, false, false, CT>(ta, tb, to, to2, tx, d, grid, stream);
    if (pair) {
        if (block_n == 256) { TAVK_GEMM_DISPATCH(256, 2) }
        TAVK_GEMM_DISPATCH(128, 2)
    }
    if (block_n == 256) { TAVK_GEMM_DISPATCH(256, 1) }
    if (block_n == 128) { TAVK_GEMM_DISPATCH(128, 1) }
    TAVK_GEMM_DISPATCH(64, 1)
#undef TAVK_GEMM_DISPATCH
}
