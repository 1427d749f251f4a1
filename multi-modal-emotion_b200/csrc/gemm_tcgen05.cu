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
// K-major operand tile  : one TMA box {64 k, rows}; UMMA desc SBO = 1024 B, K advance = +32 B per UMMA_K.
// MN-major operand tile : rows/64 TMA boxes {64 mn, 64 k}; UMMA desc LBO = 8192 B (next 64-wide MN group),
//                         SBO = 1024 B (next 8 k rows), K advance = +2048 B per UMMA_K.
#include "common.cuh"
#include "../../include/tavk.h"

namespace tavk {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
constexpr int kNumEpiWarps = 8;
constexpr int kGemmThreads = (2 + kNumEpiWarps) * 32;

template <int BLOCK_N>
struct GemmCfg {
    static constexpr int kStages = (BLOCK_N == 256) ? 4 : 6;
    static constexpr int kABytes = kBlockM * kBlockK * 2;
    static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kTmemCols = 2 * BLOCK_N;  // double-buffered accumulator (256 or 512 columns)
    static constexpr int kStagingBytes = kNumEpiWarps * 4096;   // one 32x32 fp32 transpose patch per epilogue warp
    static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct GemmDev {
    int M, N, K;
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
    int epilogue;
    int accumulate;
    float alpha;
};

struct EpiCtx {
    uint32_t tmem_acc;      // TMEM address of this warp's lane quarter, column 0 of the tile's accumulator
    uint32_t stg;           // this warp's 4 KB staging patch (shared-memory address)
    int lane, half, row_base, col_base;
    bool use_bias, use_rb, use_res;
    uint64_t* tmem_empty;
};

// One epilogue warp's share of one output tile: 32 accumulator rows x every second 32-column chunk.
// TMEM -> registers (lane = row) -> XOR-swizzled smem patch -> registers in the COALESCED layout (8 lanes x 4 columns
// cover one row's chunk, 4 rows per instruction), where bias / row-bias / residual / GELU / GELU' / column sums are
// applied and the result stored: each global access of the warp touches 4 full 128-byte (fp32) or 64-byte (bf16) row
// segments instead of 32 rows x 16 B.  The chunk's global operands (residual, pre-activation) are requested BEFORE
// the TMEM read so their latency hides behind it.
template <int BLOCK_N, int MODE>
TAVK_DEVINL void epilogue_tile(const GemmDev& p, const EpiCtx& cx) {
    constexpr int kChunks = BLOCK_N / 32;
    const int lane = cx.lane;
    const int cc = lane & 7;                // 4-column group inside the chunk
    const int rsub = lane >> 3;             // row inside each group of 4
    const uint32_t st_addr = cx.stg + lane * 128;
#pragma unroll 1
    for (int c = cx.half; c < kChunks; c += 2) {
        const int col = cx.col_base + c * 32 + cc * 4;
        const bool col_ok = col < p.N;          // N % 8 == 0: a 4-column group is in or out as a whole
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cx.use_bias && col_ok) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
        float4 res[8];
        uint2 ax[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = cx.row_base + i * 4 + rsub;
            const bool ok = row < p.M && col_ok;
            if (MODE == TAVK_EPI_GELU_BWD) {
                ax[i] = make_uint2(0u, 0u);
                if (ok) ax[i] = __ldg(reinterpret_cast<const uint2*>(p.aux + (long long)row * p.ldaux + col));
            }
            if (MODE == TAVK_EPI_LINEAR) {
                res[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (cx.use_res && ok) res[i] = ld_global_nc_v4(p.resid + (long long)row * p.ldr + col);
                if (cx.use_rb && ok) {
                    const float4 q4 = __ldg(reinterpret_cast<const float4*>(
                        p.rowbias + (long long)(row / p.rows_per_group) * p.N + col));
                    res[i].x += q4.x; res[i].y += q4.y; res[i].z += q4.z; res[i].w += q4.w;
                }
            }
        }
        uint32_t r[32];
        tmem_ld_32x32(cx.tmem_acc + (uint32_t)(c * 32), r);
        tmem_ld_wait();
        if (c + 2 >= kChunks) {
            // this warp has read its whole share of the accumulator: hand the TMEM buffer back to the MMA issuer
            // before the global-memory part of the last chunk
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(cx.tmem_empty);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
            st_shared_v4(st_addr + ((j ^ (lane & 7)) << 4), r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        __syncwarp();
        float cs0 = 0.f, cs1 = 0.f, cs2 = 0.f, cs3 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int rl = i * 4 + rsub;
            const int row = cx.row_base + rl;
            float4 v = ld_shared_v4(cx.stg + rl * 128 + ((cc ^ (rl & 7)) << 4));
            if (row < p.M && col_ok) {
                v.x = fmaf(v.x, p.alpha, b4.x); v.y = fmaf(v.y, p.alpha, b4.y);
                v.z = fmaf(v.z, p.alpha, b4.z); v.w = fmaf(v.w, p.alpha, b4.w);
                if (MODE == TAVK_EPI_GELU) {
                    // out = pre-activation (bf16), out2 = GELU(pre) (bf16)
                    float4 g;
                    gelu_fast2(v.x, v.y, g.x, g.y);
                    gelu_fast2(v.z, v.w, g.z, g.w);
                    uint2 a, gg;
                    a.x = pack_bf16x2(v.x, v.y);  a.y = pack_bf16x2(v.z, v.w);
                    gg.x = pack_bf16x2(g.x, g.y); gg.y = pack_bf16x2(g.z, g.w);
                    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)row * p.ldo + col) = a;
                    *reinterpret_cast<uint2*>(p.out2 + (long long)row * p.ldo2 + col) = gg;
                    v = g;
                } else {
                    if (MODE == TAVK_EPI_GELU_BWD) {
                        // out = acc * gelu'(aux)
                        const float2 a0 = unpack_bf16x2(ax[i].x), a1 = unpack_bf16x2(ax[i].y);
                        gelu_grad_mul2(a0.x, a0.y, v.x, v.y);
                        gelu_grad_mul2(a1.x, a1.y, v.z, v.w);
                    } else {
                        v.x += res[i].x; v.y += res[i].y; v.z += res[i].z; v.w += res[i].w;
                    }
                    if (p.out_bf16) {
                        uint2 a;
                        a.x = pack_bf16x2(v.x, v.y); a.y = pack_bf16x2(v.z, v.w);
                        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)row * p.ldo + col) = a;
                    } else {
                        float* o = reinterpret_cast<float*>(p.out) + (long long)row * p.ldo + col;
                        if (p.accumulate) {
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(v.x), "f"(v.y),
                                         "f"(v.z), "f"(v.w)
                                         : "memory");
                        } else {
                            *reinterpret_cast<float4*>(o) = v;
                        }
                    }
                }
                cs0 += v.x; cs1 += v.y; cs2 += v.z; cs3 += v.w;
            }
        }
        if (p.colsum != nullptr) {
            // column sums of the stored values (bias gradient of the Linear whose output gradient this GEMM produces)
            cs0 += __shfl_xor_sync(0xffffffffu, cs0, 8);  cs1 += __shfl_xor_sync(0xffffffffu, cs1, 8);
            cs2 += __shfl_xor_sync(0xffffffffu, cs2, 8);  cs3 += __shfl_xor_sync(0xffffffffu, cs3, 8);
            cs0 += __shfl_xor_sync(0xffffffffu, cs0, 16); cs1 += __shfl_xor_sync(0xffffffffu, cs1, 16);
            cs2 += __shfl_xor_sync(0xffffffffu, cs2, 16); cs3 += __shfl_xor_sync(0xffffffffu, cs3, 16);
            if (lane < 8 && col_ok)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.colsum + col), "f"(cs0), "f"(cs1),
                             "f"(cs2), "f"(cs3)
                             : "memory");
        }
        __syncwarp();   // the staging patch is rewritten by the next chunk
    }
}

template <int BLOCK_N, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const GemmDev p) {
    using Cfg = GemmCfg<BLOCK_N>;
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
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::kStages + 4);

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
            mbar_init(&tmem_empty_bar[i], kNumEpiWarps);
        }
        mbar_fence_init();
    }
    if (warp_idx == 1) tmem_alloc<Cfg::kTmemCols>(tmem_ptr_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);

    const int num_tiles = p.num_m_blocks * p.num_n_blocks * p.k_splits;

    if (warp_idx == 0) {
        // ===================== TMA producer (one thread) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int split = tile % p.k_splits;
                const int mn = tile / p.k_splits;
                const int n_blk = mn % p.num_n_blocks;
                const int m_blk = mn / p.num_n_blocks;
                const int kb0 = split * p.kb_per_split;
                const int kb1 = min(kb0 + p.kb_per_split, p.num_kb);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                    uint8_t* sa = smem_a + stage * Cfg::kABytes;
                    uint8_t* sb = smem_b + stage * Cfg::kBBytes;
                    if constexpr (A_MN) {
#pragma unroll
                        for (int j = 0; j < kBlockM / 64; ++j)
                            tma_load_2d(sa + j * 8192, &tmap_a, &full_bar[stage], m_blk * kBlockM + j * 64,
                                        kb * kBlockK);
                    } else {
                        tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * kBlockK, m_blk * kBlockM);
                    }
                    if constexpr (B_MN) {
#pragma unroll
                        for (int j = 0; j < BLOCK_N / 64; ++j)
                            tma_load_2d(sb + j * 8192, &tmap_b, &full_bar[stage], n_blk * BLOCK_N + j * 64,
                                        kb * kBlockK);
                    } else {
                        tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * kBlockK, n_blk * BLOCK_N);
                    }
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp_idx == 1) {
        // ===================== MMA issuer: the whole warp walks the (warp-uniform) loop, one elected lane issues ====
        const bool leader = elect_one();
        constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BLOCK_N, A_MN, B_MN);
        const uint64_t da0 = A_MN ? umma_smem_desc(smem_u32(smem_a), 8192, 1024) : umma_smem_desc(smem_u32(smem_a), 16, 1024);
        const uint64_t db0 = B_MN ? umma_smem_desc(smem_u32(smem_b), 8192, 1024) : umma_smem_desc(smem_u32(smem_b), 16, 1024);
        constexpr uint64_t kAStep = A_MN ? (2048 >> 4) : (32 >> 4);   // descriptor address units (16 B) per UMMA_K
        constexpr uint64_t kBStep = B_MN ? (2048 >> 4) : (32 >> 4);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int split = tile % p.k_splits;
            const int kb0 = split * p.kb_per_split;
            const int kb1 = min(kb0 + p.kb_per_split, p.num_kb);
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (leader) {
                    const uint64_t da = da0 + (uint64_t)(stage * (Cfg::kABytes >> 4));
                    const uint64_t db = db0 + (uint64_t)(stage * (Cfg::kBBytes >> 4));
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k)
                        umma_bf16(tmem_d, da + k * kAStep, db + k * kBStep, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
                }
                __syncwarp();
                if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
            }
            if (leader) umma_commit(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
            __syncwarp();
        }
    } else {
        // ===================== epilogue warps =====================
        const int ew = warp_idx - 2;            // 0..7
        const int quarter = warp_idx & 3;       // TMEM lane quarter this warp may access
        const int half = ew >> 2;               // which half of the column chunks
        const uint32_t stg = smem_u32(smem_stage) + ew * 4096;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int split = tile % p.k_splits;
            const int mn = tile / p.k_splits;
            const int n_blk = mn % p.num_n_blocks;
            const int m_blk = mn / p.num_n_blocks;
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait(&tmem_full_bar[acc], acc_phase);
            tc_fence_after();
            const int row_base = m_blk * kBlockM + quarter * 32;
            const bool lead_split = (split == 0);
            const bool use_bias = p.bias != nullptr && lead_split;
            const bool use_rb = p.rowbias != nullptr && lead_split;
            const bool use_res = p.resid != nullptr && lead_split;
            EpiCtx cx;
            cx.tmem_acc = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BLOCK_N);
            cx.stg = stg; cx.lane = lane; cx.half = half; cx.row_base = row_base; cx.col_base = n_blk * BLOCK_N;
            cx.use_bias = use_bias; cx.use_rb = use_rb; cx.use_res = use_res;
            cx.tmem_empty = &tmem_empty_bar[acc];
            if (p.epilogue == TAVK_EPI_GELU) epilogue_tile<BLOCK_N, TAVK_EPI_GELU>(p, cx);
            else if (p.epilogue == TAVK_EPI_GELU_BWD) epilogue_tile<BLOCK_N, TAVK_EPI_GELU_BWD>(p, cx);
            else epilogue_tile<BLOCK_N, TAVK_EPI_LINEAR>(p, cx);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp_idx == 1) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc<Cfg::kTmemCols>(tmem_base);
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
                          int box_rows) {
    PFN_encodeTiled enc = get_encode_fn();
    TAVK_CHECK(enc != nullptr, 3, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TAVK_CHECK(r == CUDA_SUCCESS, 3, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld box=%dx%d base=%p",
               (int)r, rows, cols, ld, box_cols, box_rows, base);
    return 0;
}

template <int BLOCK_N, bool A_MN, bool B_MN>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmDev& dev, int grid, cudaStream_t stream) {
    using Cfg = GemmCfg<BLOCK_N>;
    auto kern = gemm_bf16_tcgen05_kernel<BLOCK_N, A_MN, B_MN>;
    static bool attr_done = false;
    if (!attr_done) {
        TAVK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
        attr_done = true;
    }
    kern<<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ta, tb, dev);
    TAVK_CUDA(cudaGetLastError());
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
    TAVK_CHECK(a->epilogue != TAVK_EPI_GELU || (a->out_dtype == TAVK_BF16 && a->out2 != nullptr && k_splits == 1), 1,
               "tavk_gemm_bf16: GELU epilogue needs bf16 out + out2 and no split-K");
    TAVK_CHECK(a->epilogue != TAVK_EPI_GELU_BWD || (a->aux != nullptr && k_splits == 1), 1,
               "tavk_gemm_bf16: GELU_BWD epilogue needs aux and no split-K");
    TAVK_CHECK(a->rowbias == nullptr || a->rows_per_group > 0, 1, "tavk_gemm_bf16: rowbias needs rows_per_group");
    const int vec = (a->out_dtype == TAVK_BF16) ? 8 : 4;
    TAVK_CHECK(a->ldo % vec == 0, 1, "tavk_gemm_bf16: ldo must be a multiple of %d", vec);

    // tile-shape heuristic: prefer 128x256 unless 128x128 fills the SMs noticeably better
    const int sms = sm_count();
    const int mblocks = (a->M + kBlockM - 1) / kBlockM;
    auto waves_eff = [&](int bn) {
        const long long tiles = (long long)mblocks * ((a->N + bn - 1) / bn) * k_splits;
        const long long waves = (tiles + sms - 1) / sms;
        return (double)tiles / (double)(waves * sms);
    };
    int block_n = 256;
    if (a->block_n == 128 || a->block_n == 256) block_n = a->block_n;
    else if (a->N <= 128 || waves_eff(128) > waves_eff(256) * 1.15) block_n = 128;

    GemmDev d;
    d.M = a->M; d.N = a->N; d.K = a->K;
    d.num_m_blocks = mblocks;
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

    CUtensorMap ta, tb;
    int rc;
    if (a->a_mn_major) rc = make_tmap_bf16(&ta, a->A, a->K, a->M, a->lda, 64, kBlockK);
    else               rc = make_tmap_bf16(&ta, a->A, a->M, a->K, a->lda, kBlockK, kBlockM);
    if (rc) return rc;
    if (a->b_mn_major) rc = make_tmap_bf16(&tb, a->B, a->K, a->N, a->ldb, 64, kBlockK);
    else               rc = make_tmap_bf16(&tb, a->B, a->N, a->K, a->ldb, kBlockK, block_n);
    if (rc) return rc;

    const long long tiles = (long long)d.num_m_blocks * d.num_n_blocks * d.k_splits;
    const int grid = (int)(tiles < sms ? tiles : sms);
#define TAVK_GEMM_DISPATCH(BN)                                                                   \
    if (a->a_mn_major && a->b_mn_major) return launch_gemm<BN, true, true>(ta, tb, d, grid, stream);   \
    if (a->a_mn_major && !a->b_mn_major) return launch_gemm<BN, true, false>(ta, tb, d, grid, stream); \
    if (!a->a_mn_major && a->b_mn_major) return launch_gemm<BN, false, true>(ta, tb, d, grid, stream); \
    return launch_gemm<BN, false, false>(ta, tb, d, grid, stream);
    if (block_n == 256) { TAVK_GEMM_DISPATCH(256) }
    TAVK_GEMM_DISPATCH(128)
#undef TAVK_GEMM_DISPATCH
}
