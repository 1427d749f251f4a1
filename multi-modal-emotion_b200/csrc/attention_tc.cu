// tcgen05 / TMEM flash attention for head_dim 64 (v2 of the attention path; attention.cu keeps the mma.sync v1 that
// still serves key-bias masks and very short sequences).
//
// Replaces the same reference op sites as attention.cu: utils/TAVFormer.py:357-387 (QK^T/8, softmax, PV) and the HF
// encoders' attention.  Motivation (profiles/r1_*): at the VideoMAE shape (S=1464, 192 (b,h) pairs) the mma.sync
// kernels were 24 ms of a 90 ms step; the legacy tensor path peaks at a quarter of tcgen05's rate.
//
// Forward, one CTA = 128 queries of one (batch, head):
//   warp 0      TMA producer: Q once, then a 3-stage ring of (K_j, V_j) 128x64 bf16 tiles (3-D tensor maps over
//               [B][S][row], SWIZZLE_128B, rows >= S zero-filled by the hardware)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer:
//                 S_j = Q K_j^T   (UMMA 128x128x16 x4, both operands K-major)        -> TMEM S[j&1] (128 fp32 columns)
//                 O  += P_j V_j   (UMMA 128x64x16 x8, A = P from smem, B = V MN-major) -> TMEM O (64 columns)
//               QK of tile j+1 is issued before PV of tile j, so the tensor pipe works while tile j's softmax runs
//   warps 2..5  softmax: one thread per query row (TMEM lane).  tcgen05.ld the 128 scores of the row, row max / sum
//               without any shuffle, p = ex2(s*scale*log2e - m), bf16 P written to smem in the UMMA K-major SW128
//               layout.  Lazy rescaling: the running max only moves (and O in TMEM is only rescaled, tcgen05.ld/st)
//               when it grows by more than 8 in the log2 domain, so the common path never touches O.
// Epilogue: O / l -> bf16 rows, lse = (m + log2 l) ln2.
#include "../../include/tavk.h"
#include "common.cuh"

namespace tavk {

constexpr int kTcQ = 128, kTcKV = 128, kTcD = 64;
constexpr int kTcStages = 3;
constexpr int kTcTile = kTcKV * kTcD * 2;            // 16 KB  (K, V, Q tiles)
constexpr int kTcPBytes = kTcQ * kTcKV * 2;          // 32 KB  (P tile, two 64-key SW128 atoms)
constexpr int kTcSmem = kTcTile * (1 + 2 * kTcStages) + 2 * kTcPBytes + 1024 + 256;
constexpr int kTcThreads = 6 * 32;
constexpr uint32_t kTcTmemCols = 512;                // S[2] 256 + O 64 -> next power of two
constexpr float kTcLog2e = 1.4426950408889634f, kTcLn2 = 0.6931471805599453f;
constexpr float kRescaleThreshold = 8.0f;

TAVK_DEVINL float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
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
TAVK_DEVINL void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

struct AttnTcDev {
    __nv_bfloat16* o;
    long long ld_o;
    float* lse;
    int B, S, nh;
    float scale_log2;
};

__global__ void __launch_bounds__(kTcThreads, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                   const __grid_constant__ CUtensorMap tmap_v, const AttnTcDev p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = smem + kTcTile;
    uint8_t* sV = sK + kTcStages * kTcTile;
    uint8_t* sP = sV + kTcStages * kTcTile;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kTcPBytes);
    uint64_t* q_full = bars;               // 1
    uint64_t* kv_full = bars + 1;          // kTcStages
    uint64_t* kv_empty = kv_full + kTcStages;
    uint64_t* s_full = kv_empty + kTcStages;  // 2
    uint64_t* p_full = s_full + 2;            // 2
    uint64_t* p_empty = p_full + 2;           // 2
    uint64_t* pv_done = p_empty + 2;          // 1
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(pv_done + 1);

    const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * kTcQ, h = blockIdx.y, b = blockIdx.z;
    const int n_tiles = (p.S + kTcKV - 1) / kTcKV;

    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_k);
        tma_prefetch_desc(&tmap_v);
        mbar_init(q_full, 1);
        for (int i = 0; i < kTcStages; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); mbar_init(&p_empty[i], 1); }
        mbar_init(pv_done, 1);
        mbar_fence_init();
    }
    if (warp_idx == 1) tmem_alloc<kTcTmemCols>(tmem_ptr_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t tmem_o = tmem_base + 256;

    if (warp_idx == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, kTcTile);
            tma_load_3d(sQ, &tmap_q, q_full, h * kTcD, q0, b);
            for (int j = 0; j < n_tiles; ++j) {
                const int st = j % kTcStages;
                const uint32_t ph = (j / kTcStages) & 1;
                mbar_wait(&kv_empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&kv_full[st], 2 * kTcTile);
                tma_load_3d(sK + st * kTcTile, &tmap_k, &kv_full[st], h * kTcD, j * kTcKV, b);
                tma_load_3d(sV + st * kTcTile, &tmap_v, &kv_full[st], h * kTcD, j * kTcKV, b);
            }
        }
    } else if (warp_idx == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_qk = umma_idesc_bf16(kTcQ, kTcKV, false, false);
            constexpr uint32_t idesc_pv = umma_idesc_bf16(kTcQ, kTcD, false, true);
            const uint32_t q_addr = smem_u32(sQ);
            auto issue_qk = [&](int j) {
                const int st = j % kTcStages;
                mbar_wait(&kv_full[st], (j / kTcStages) & 1);
                tc_fence_after();
                const uint32_t k_addr = smem_u32(sK + st * kTcTile);
                const uint32_t tmem_s = tmem_base + (j & 1) * kTcKV;
#pragma unroll
                for (int k = 0; k < kTcD / 16; ++k)
                    umma_bf16(tmem_s, umma_smem_desc(q_addr + k * 32, 16, 1024), umma_smem_desc(k_addr + k * 32, 16, 1024),
                              idesc_qk, k > 0 ? 1u : 0u);
                umma_commit(&s_full[j & 1]);
            };
            mbar_wait(q_full, 0);
            tc_fence_after();
            issue_qk(0);
            for (int j = 0; j < n_tiles; ++j) {
                if (j + 1 < n_tiles) issue_qk(j + 1);
                mbar_wait(&p_full[j & 1], (j >> 1) & 1);
                tc_fence_after();
                const int st = j % kTcStages;
                const uint32_t p_addr = smem_u32(sP + (j & 1) * kTcPBytes);
                const uint32_t v_addr = smem_u32(sV + st * kTcTile);
#pragma unroll
                for (int k = 0; k < kTcKV / 16; ++k) {
                    const uint64_t da = umma_smem_desc(p_addr + (k >> 2) * (kTcQ * 128) + (k & 3) * 32, 16, 1024);
                    const uint64_t db = umma_smem_desc(v_addr + k * 2048, 8192, 1024);
                    umma_bf16(tmem_o, da, db, idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&kv_empty[st]);
                umma_commit(&p_empty[j & 1]);
                umma_commit(pv_done);
            }
        }
    } else {
        // ===================== softmax / epilogue: thread = one query row = one TMEM lane =====================
        const int quarter = warp_idx & 3;
        const int row = quarter * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
        float m_used = -INFINITY, l = 0.f;
        for (int j = 0; j < n_tiles; ++j) {
            mbar_wait(&s_full[j & 1], (j >> 1) & 1);
            tc_fence_after();
            uint32_t sr[128];
            const uint32_t ts = tmem_base + lane_sel + (uint32_t)((j & 1) * kTcKV);
            tmem_ld_32x32(ts + 0, reinterpret_cast<uint32_t(&)[32]>(sr[0]));
            tmem_ld_32x32(ts + 32, reinterpret_cast<uint32_t(&)[32]>(sr[32]));
            tmem_ld_32x32(ts + 64, reinterpret_cast<uint32_t(&)[32]>(sr[64]));
            tmem_ld_32x32(ts + 96, reinterpret_cast<uint32_t(&)[32]>(sr[96]));
            tmem_ld_wait();
            const int valid = min(kTcKV, p.S - j * kTcKV);   // keys of this tile that exist
            float mx = -INFINITY;
            if (valid == kTcKV) {
#pragma unroll
                for (int c = 0; c < 128; ++c) mx = fmaxf(mx, __uint_as_float(sr[c]));
            } else {
#pragma unroll
                for (int c = 0; c < 128; ++c) {
                    if (c >= valid) sr[c] = 0xff800000u;  // -inf
                    mx = fmaxf(mx, __uint_as_float(sr[c]));
                }
            }
            mx *= p.scale_log2;
            float factor = 1.0f;
            const bool grow = mx > m_used + kRescaleThreshold;   // always true on the first tile (m_used = -inf)
            if (grow) {
                factor = ex2_approx(m_used - mx);                 // 0 on the first tile
                l *= factor;
                m_used = mx;
            }
            float sum = 0.f;
            const float neg_m = -m_used;
#pragma unroll
            for (int c = 0; c < 128; c += 2) {
                const float p0 = ex2_approx(fmaf(__uint_as_float(sr[c]), p.scale_log2, neg_m));
                const float p1 = ex2_approx(fmaf(__uint_as_float(sr[c + 1]), p.scale_log2, neg_m));
                sum += p0 + p1;
                sr[c >> 1] = pack_bf16x2(p0, p1);
            }
            l += sum;
            // P[j&1] must no longer be read by PV(j-2)
            if (j >= 2) mbar_wait(&p_empty[j & 1], ((j >> 1) - 1) & 1);
            const uint32_t pbase = smem_u32(sP + (j & 1) * kTcPBytes) + row * 128;
#pragma unroll
            for (int ch = 0; ch < 16; ++ch) {
                const uint32_t addr = pbase + (ch >> 3) * (kTcQ * 128) + (((ch & 7) ^ (row & 7)) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(sr[4 * ch]), "r"(sr[4 * ch + 1]),
                             "r"(sr[4 * ch + 2]), "r"(sr[4 * ch + 3])
                             : "memory");
            }
            // rare: the running max moved -> rescale this warp's 32 rows of O (needs PV(j-1) retired)
            if (j > 0 && __any_sync(0xffffffffu, grow)) {
                mbar_wait(pv_done, (j - 1) & 1);
                tc_fence_after();
                uint32_t orow[32];
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const uint32_t to = tmem_o + lane_sel + half * 32;
                    tmem_ld_32x32(to, orow);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 32; ++c) orow[c] = __float_as_uint(__uint_as_float(orow[c]) * factor);
                    tmem_st_32x32(to, orow);
                }
                tmem_st_wait();
            }
            fence_proxy_async_smem();   // generic-proxy smem writes of P -> visible to the tensor core (async proxy)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[j & 1]);
        }
        // ---- epilogue
        mbar_wait(pv_done, (n_tiles - 1) & 1);
        tc_fence_after();
        const int qrow = q0 + row;
        const float inv = l > 0.f ? 1.0f / l : 0.f;
        uint32_t packed[32];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t orow[32];
            tmem_ld_32x32(tmem_o + lane_sel + half * 32, orow);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 32; c += 2)
                packed[half * 16 + (c >> 1)] = pack_bf16x2(__uint_as_float(orow[c]) * inv, __uint_as_float(orow[c + 1]) * inv);
        }
        if (qrow < p.S) {
            uint4* dst = reinterpret_cast<uint4*>(p.o + ((long long)b * p.S + qrow) * p.ld_o + h * kTcD);
#pragma unroll
            for (int i = 0; i < 8; ++i) dst[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
            if (p.lse) p.lse[((long long)b * p.nh + h) * p.S + qrow] = l > 0.f ? (m_used + log2f(l)) * kTcLn2 : -INFINITY;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp_idx == 1) {
        tc_fence_after();
        tmem_dealloc<kTcTmemCols>(tmem_base);
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
    CUtensorMap tq, tk, tv;
    const int cols = a->nh * kTcD;
    int rc = make_tmap_bsd(&tq, a->q, a->B, a->S, cols, a->ld_qkv, kTcQ);
    if (rc) return rc;
    rc = make_tmap_bsd(&tk, a->k, a->B, a->S, cols, a->ld_qkv, kTcKV);
    if (rc) return rc;
    rc = make_tmap_bsd(&tv, a->v, a->B, a->S, cols, a->ld_qkv, kTcKV);
    if (rc) return rc;
    AttnTcDev d;
    d.o = reinterpret_cast<__nv_bfloat16*>(a->o);
    d.ld_o = a->ld_o;
    d.lse = a->lse;
    d.B = a->B; d.S = a->S; d.nh = a->nh;
    d.scale_log2 = a->scale * kTcLog2e;
    static bool attr_done = false;
    if (!attr_done) {
        TAVK_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem));
        attr_done = true;
    }
    dim3 grid((a->S + kTcQ - 1) / kTcQ, a->nh, a->B);
    attn_fwd_tc_kernel<<<grid, kTcThreads, kTcSmem, stream>>>(tq, tk, tv, d);
    TAVK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace tavk
