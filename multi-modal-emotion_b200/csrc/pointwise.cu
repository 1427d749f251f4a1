// HBM-bound helper kernels of the TAV path: modality-embedding add, sequence mean-pool, (weighted) column sums,
// launch-bound fp32 small linears (classifier head, rank-1 attention term), casts, scaling and head dropout.
// All are coalesced float4 / 8-byte bf16 accesses; grids are sized from the SM count.
#include "../../include/tavk.h"
#include "common.cuh"

namespace tavk {

// ---------------------------------------------------------------- embed add (models/tav.py:474)
__global__ void embed_add_fwd_kernel(const float4* __restrict__ x, const int64_t* __restrict__ idx,
                                     const float4* __restrict__ table, float4* __restrict__ y, int rows, int H4,
                                     int n_embed) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    const long long total = (long long)rows * H4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(i / H4);
        const int c = (int)(i - (long long)row * H4);
        long long j = idx[row];
        j = j < 0 ? 0 : (j >= n_embed ? n_embed - 1 : j);
        const float4 a = x[i];
        const float4 t = __ldg(table + j * H4 + c);
        y[i] = make_float4(a.x + t.x, a.y + t.y, a.z + t.z, a.w + t.w);
    }
}

// dtable[j, c] += sum over rows with idx == j; block = 256 columns-quads slab, rows chunked over blockIdx.y
__global__ void embed_add_bwd_kernel(const float* __restrict__ dy, const int64_t* __restrict__ idx,
                                     float* __restrict__ dtable, int rows, int H, int n_embed, int rows_per_chunk) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= H) return;
    const int r0 = blockIdx.y * rows_per_chunk;
    const int r1 = min(r0 + rows_per_chunk, rows);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int r = r0; r < r1; ++r) {
        const int j = (int)idx[r];
        const float v = dy[(size_t)r * H + c];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += (k == j) ? v : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (k < n_embed && acc[k] != 0.f) atomicAdd(dtable + (size_t)k * H + c, acc[k]);
}

// ---------------------------------------------------------------- RoBERTa input embeddings (models/tav.py:349,485)
// HF RobertaEmbeddings.forward up to (not including) its LayerNorm: y[b,t,:] = word[ids[b,t]] + pos[p[b,t]] + type[0],
// p = padding_idx + (running count of non-pad tokens) for non-pad tokens, padding_idx for pad tokens
// (create_position_ids_from_input_ids).  One block per sequence: thread 0 scans the T ids into shared memory.
__global__ void __launch_bounds__(256)
roberta_embed_fwd_kernel(const int64_t* __restrict__ ids, const float4* __restrict__ word, const float4* __restrict__ pos,
                         const float4* __restrict__ type0, float4* __restrict__ y, int64_t* __restrict__ pos_ids, int T,
                         int H4, int V, int P, int pad) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    extern __shared__ int s_pos[];
    const int b = blockIdx.x;
    const int64_t* row = ids + (size_t)b * T;
    if (threadIdx.x == 0) {
        int run = 0;
        for (int t = 0; t < T; ++t) {
            const bool tok = row[t] != pad;
            run += tok ? 1 : 0;
            int pi = tok ? run + pad : pad;
            pi = pi < 0 ? 0 : (pi >= P ? P - 1 : pi);
            s_pos[t] = pi;
            pos_ids[(size_t)b * T + t] = pi;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < T * H4; i += blockDim.x) {
        const int t = i / H4, c = i - t * H4;
        long long w = row[t];
        w = w < 0 ? 0 : (w >= V ? V - 1 : w);
        const float4 a = __ldg(word + (size_t)w * H4 + c), p4 = __ldg(pos + (size_t)s_pos[t] * H4 + c), t4 = __ldg(type0 + c);
        y[((size_t)b * T + t) * H4 + c] = make_float4((a.x + t4.x) + p4.x, (a.y + t4.y) + p4.y, (a.z + t4.z) + p4.z, (a.w + t4.w) + p4.w);   // HF order
    }
}
// dtable[idx[r], :] += dy[r, :]  (row scatter with vector reductions; the table gradient is never materialised densely
// by this call: it accumulates into whatever buffer the caller owns, e.g. the parameter's slice of the flat gradient)
__global__ void embedding_scatter_add_kernel(const float* __restrict__ dy, const int64_t* __restrict__ idx,
                                             float* __restrict__ dtable, int rows, int H4, int n_embed, int skip_idx) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    const long long total = (long long)rows * H4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / H4), c = (int)(i - (long long)r * H4);
        const long long j = idx[r];
        if (j < 0 || j >= n_embed || j == skip_idx) continue;     // nn.Embedding(padding_idx=...): that row gets no gradient
        const float4 v = reinterpret_cast<const float4*>(dy)[i];
        float* o = dtable + ((size_t)j * H4 + c) * 4;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    }
}

// ---------------------------------------------------------------- mean pool (models/tav.py:478,481,488)
// `lengths` (int32 [B] or NULL): masked mean over the first lengths[b] rows of sample b (divisor lengths[b]); NULL = all S
// rows (the reference pools unmasked; the masked form is the boundary's optional argument).
__global__ void mean_pool_fwd_kernel(const float4* __restrict__ x, float* __restrict__ y, int S, int H4,
                                     int rows_per_chunk, float inv_s, const int* __restrict__ lengths) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= H4) return;
    const int b = blockIdx.z;
    int len = S;
    if (lengths != nullptr) {
        len = min(max(lengths[b], 0), S);
        inv_s = len > 0 ? 1.0f / (float)len : 0.f;
    }
    const int s0 = blockIdx.y * rows_per_chunk;
    const int s1 = min(s0 + rows_per_chunk, len);
    if (s0 >= s1) return;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* xb = x + (size_t)b * S * H4;
    for (int s = s0; s < s1; ++s) {
        const float4 v = xb[(size_t)s * H4 + c];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    float* o = y + ((size_t)b * H4 + c) * 4;
    atomicAdd(o + 0, acc.x * inv_s);
    atomicAdd(o + 1, acc.y * inv_s);
    atomicAdd(o + 2, acc.z * inv_s);
    atomicAdd(o + 3, acc.w * inv_s);
}

__global__ void mean_pool_bwd_kernel(const float4* __restrict__ dy, float4* __restrict__ dx,
                                     uint2* __restrict__ dx_bf16, int S, int H4, long long total, float inv_s,
                                     const int* __restrict__ lengths) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / H4;
        const int c = (int)(i - row * H4);
        const int b = (int)(row / S);
        float sc = inv_s;
        if (lengths != nullptr) {
            const int len = min(max(lengths[b], 0), S);
            sc = ((int)(row - (long long)b * S) < len) ? 1.0f / (float)len : 0.f;
        }
        float4 v = __ldg(dy + (size_t)b * H4 + c);
        v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
        if (dx) dx[i] = v;
        if (dx_bf16) dx_bf16[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
}

// ---------------------------------------------------------------- (weighted) column sums
// out[b, n] (+)= sum_{s in chunk} w[b,s] * x[b,s,n]; each thread owns 4 consecutive columns.
template <bool BF16>
__global__ void colsum_kernel(const void* __restrict__ x_, long long ld, const float* __restrict__ w,
                              float* __restrict__ out, int S, int N, int rows_per_chunk) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    const int c4 = blockIdx.x * blockDim.x + threadIdx.x;
    if (c4 * 4 >= N) return;
    const int b = blockIdx.z;
    const int s0 = blockIdx.y * rows_per_chunk;
    const int s1 = min(s0 + rows_per_chunk, S);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* wb = w ? w + (size_t)b * S : nullptr;
    auto load = [&](int s) -> float4 {
        if (BF16) {
            const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(x_) +
                                                           ((size_t)b * S + s) * ld + c4 * 4);
            const float2 a = unpack_bf16x2(u.x), c = unpack_bf16x2(u.y);
            return make_float4(a.x, a.y, c.x, c.y);
        }
        return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x_) + ((size_t)b * S + s) * ld + c4 * 4);
    };
    int s = s0;
    for (; s + 8 <= s1; s += 8) {   // eight independent row loads in flight per thread
        float4 v[8];
        float ws[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { v[j] = load(s + j); ws[j] = wb ? wb[s + j] : 1.0f; }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            acc.x += ws[j] * v[j].x; acc.y += ws[j] * v[j].y; acc.z += ws[j] * v[j].z; acc.w += ws[j] * v[j].w;
        }
    }
    for (; s < s1; ++s) {
        const float4 v = load(s);
        const float ws = wb ? wb[s] : 1.0f;
        acc.x += ws * v.x; acc.y += ws * v.y; acc.z += ws * v.z; acc.w += ws * v.w;
    }
    float* o = out + (size_t)b * N + c4 * 4;
    atomicAdd(o + 0, acc.x);
    atomicAdd(o + 1, acc.y);
    atomicAdd(o + 2, acc.z);
    atomicAdd(o + 3, acc.w);
}

static int launch_colsum(const void* x, int x_dtype, long long ld, const float* w, float* out, int B, int S, int N,
                         int accumulate, cudaStream_t stream) {
    TAVK_CHECK(x && out, 1, "colsum: null pointer");
    TAVK_CHECK(x_dtype == TAVK_F32 || x_dtype == TAVK_BF16, 1, "colsum: bad dtype");
    TAVK_CHECK(N % 4 == 0 && ld % 4 == 0, 1, "colsum: N and ld must be multiples of 4 (N=%d ld=%lld)", N, ld);
    if (B <= 0 || N <= 0) return 0;
    if (!accumulate) TAVK_CUDA(cudaMemsetAsync(out, 0, (size_t)B * N * sizeof(float), stream));
    if (S <= 0) return 0;
    // one block spans the columns when they fit 256 threads (N = 768: 192 threads, no half-empty second block)
    const int threads = (N / 4 <= 256) ? ((N / 4 + 31) / 32) * 32 : 128;
    const int gx = (N / 4 + threads - 1) / threads;
    // enough row chunks for ~4 blocks per SM: with eight 8/16-byte loads in flight per thread that is what it takes to
    // cover HBM latency (the previous 2 blocks per SM x 4 loads read the fusion block's dx1 at 1.1 TB/s)
    int chunks = (4 * sm_count() + gx * B - 1) / (gx * B);
    if (chunks < 1) chunks = 1;
    int rows_per_chunk = (S + chunks - 1) / chunks;
    if (rows_per_chunk < 16) rows_per_chunk = 16;
    chunks = (S + rows_per_chunk - 1) / rows_per_chunk;
    dim3 grid(gx, chunks, B);
    if (x_dtype == TAVK_BF16) TAVK_CUDA(launch_kernel(colsum_kernel<true>, dim3(grid), dim3(threads), (size_t)(0), stream, x, ld, w, out, S, N, rows_per_chunk));
    else                      TAVK_CUDA(launch_kernel(colsum_kernel<false>, dim3(grid), dim3(threads), (size_t)(0), stream, x, ld, w, out, S, N, rows_per_chunk));
    TAVK_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------- small fp32 linears
// fp32 end to end (the rank-1 term of the fusion attention reaches 1e7) and tiny: M = per-GPU batch rows.
// y[m,n] = sum_k x[m,k] w[n,k] + b[n]: one warp per output column n and per tile of kSlRows rows, so a weight row is
// read once per 8 rows (the previous warp-per-output kernel re-read W for every row: 34 us at M = 128).
constexpr int kSlRows = 8;
__global__ void __launch_bounds__(256)
small_linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                        float* __restrict__ y, int M, int N, int K) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    const int lane = threadIdx.x & 31;
    const int n = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
    const int m0 = blockIdx.y * kSlRows;
    if (n >= N) return;
    const float* wr = w + (size_t)n * K;
    float acc[kSlRows];
#pragma unroll
    for (int j = 0; j < kSlRows; ++j) acc[j] = 0.f;
    if ((K & 3) == 0) {
        for (int k = lane * 4; k < K; k += 128) {
            const float4 c = __ldg(reinterpret_cast<const float4*>(wr + k));
#pragma unroll
            for (int j = 0; j < kSlRows; ++j) {
                if (m0 + j < M) {
                    const float4 a = *reinterpret_cast<const float4*>(x + (size_t)(m0 + j) * K + k);
                    acc[j] += (a.x * c.x + a.y * c.y) + (a.z * c.z + a.w * c.w);
                }
            }
        }
    } else {
        for (int k = lane; k < K; k += 32) {
            const float c = wr[k];
#pragma unroll
            for (int j = 0; j < kSlRows; ++j)
                if (m0 + j < M) acc[j] += x[(size_t)(m0 + j) * K + k] * c;
        }
    }
    const float bn = b ? b[n] : 0.f;
#pragma unroll
    for (int j = 0; j < kSlRows; ++j) {
        const float s = warp_sum(acc[j]);
        if (lane == 0 && m0 + j < M) y[(size_t)(m0 + j) * N + n] = s + bn;
    }
}
// dx[m,k] (+)= sum_n dy[m,n] w[n,k]: thread = one k column, block = 128 columns x one chunk of n (blockIdx.y) x one
// tile of kSlRows rows (blockIdx.z); the dy tile sits in shared memory (broadcast reads), W[n, k..k+127] is a coalesced
// row segment read once per row tile; partial sums are combined with one atomic per (row, column, chunk).
constexpr int kSlChunk = 64;
__global__ void __launch_bounds__(128)
small_linear_bwd_x_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx, int M,
                          int N, int K) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    __shared__ float s_dy[kSlRows][kSlChunk];
    const int k = blockIdx.x * 128 + threadIdx.x;
    const int n0 = blockIdx.y * kSlChunk, m0 = blockIdx.z * kSlRows;
    const int nn = min(kSlChunk, N - n0);
    for (int i = threadIdx.x; i < kSlRows * kSlChunk; i += 128) {
        const int j = i / kSlChunk, c = i - j * kSlChunk;
        s_dy[j][c] = (m0 + j < M && c < nn) ? dy[(size_t)(m0 + j) * N + n0 + c] : 0.f;
    }
    __syncthreads();
    if (k >= K) return;
    float acc[kSlRows];
#pragma unroll
    for (int j = 0; j < kSlRows; ++j) acc[j] = 0.f;
    const float* wp = w + (size_t)n0 * K + k;
    int c = 0;
    for (; c + 4 <= nn; c += 4) {
        float a[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) a[u] = __ldg(wp + (size_t)(c + u) * K);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int j = 0; j < kSlRows; ++j) acc[j] += s_dy[j][c + u] * a[u];
    }
    for (; c < nn; ++c) {
        const float a = __ldg(wp + (size_t)c * K);
#pragma unroll
        for (int j = 0; j < kSlRows; ++j) acc[j] += s_dy[j][c] * a;
    }
#pragma unroll
    for (int j = 0; j < kSlRows; ++j)
        if (m0 + j < M) atomicAdd(dx + (size_t)(m0 + j) * K + k, acc[j]);
}
// dw[n,k] += sum_m dy[m,n] x[m,k]; db[n] += sum_m dy[m,n]; one thread per (n, 4 consecutive k) when K % 4 == 0.
// The accumulation into dw / db is atomic (red.global.add): the layer engine runs this next to the tensor-core wgrad GEMM
// that adds the main term into the same gradient buffer from another stream.
__global__ void __launch_bounds__(128)
small_linear_bwd_w_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dw,
                          float* __restrict__ db, int M, int N, int K) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    const int kv = (K & 3) == 0 ? 4 : 1;
    const int kq = K / kv;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)N * kq) return;
    const int n = (int)(i / kq), k = (int)(i - (long long)n * kq) * kv;
    float sb = 0.f;
    if (kv == 4) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int m = 0; m < M; ++m) {
            const float d = __ldg(dy + (size_t)m * N + n);
            const float4 a = __ldg(reinterpret_cast<const float4*>(x + (size_t)m * K + k));
            s.x += d * a.x; s.y += d * a.y; s.z += d * a.z; s.w += d * a.w;
            sb += d;
        }
        if (dw)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dw + (size_t)n * K + k), "f"(s.x), "f"(s.y),
                         "f"(s.z), "f"(s.w)
                         : "memory");
    } else {
        float s = 0.f;
        for (int m = 0; m < M; ++m) {
            const float d = dy[(size_t)m * N + n];
            s += d * x[(size_t)m * K + k];
            sb += d;
        }
        if (dw) atomicAdd(dw + (size_t)n * K + k, s);
    }
    if (db && k == 0) atomicAdd(db + n, sb);
}


// ---------------------------------------------------------------- tiled fp32 GEMM for the aligned small linears
// C[M,N] (=|+=) sum_k A(m,k) B(k,n) (+ bias[n]) with A(m,k) = TA ? a[k*lda + m] : a[m*lda + k] and
// B(k,n) = TB ? b[n*ldb + k] : b[k*ldb + n]; everything a multiple of 4 and 16-byte aligned.  The three small linears of
// the fusion attention's rank-1 path ([B,768] x [768,768], forward / dgrad / wgrad) are 19-150 MFLOP each; the
// warp-per-output kernels above take 6-35 us for them because every warp walks K serially behind global-memory latency.
// Here a 256-thread block owns a 32 x 64 tile of C and one K-slice (blockIdx.z): operands pass through shared memory
// with 16-byte global loads, 2 x 4 outputs per thread, and enough (tile, K-slice) blocks to cover the SMs about twice;
// K-slices combine with red.global.add into a zeroed (or accumulating) C.  rowsum (TA only, blockIdx.x == 0):
// rowsum[m] += sum_k A(m,k) — the bias gradient that rides on the weight-gradient form.
constexpr int kSgBM = 32, kSgBN = 64, kSgBK = 32;
template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
small_gemm_f32_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb, float* __restrict__ c, int ldc,
                      const float* __restrict__ bias, float* __restrict__ rowsum, int M, int N, int K, int k_per_split,
                      int atomic) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    __shared__ __align__(16) float As[kSgBK][kSgBM + 4];
    __shared__ __align__(16) float Bs[kSgBK][kSgBN + 4];
    const int t = threadIdx.x;
    const int n0 = blockIdx.x * kSgBN, m0 = blockIdx.y * kSgBM;
    const int k_begin = blockIdx.z * k_per_split, k_end = min(K, k_begin + k_per_split);
    const int tx = t & 15, ty = t >> 4;          // 4 columns (tx*4 ..) x 2 rows (ty*2 ..) per thread
    float acc[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float rs = 0.f;
    for (int k0 = k_begin; k0 < k_end; k0 += kSgBK) {
        // ---- A tile -> As[k][m]
        {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (TA) {                                   // contiguous along m: thread = (k = t/8, 4 m)
                const int k = k0 + (t >> 3), m = m0 + (t & 7) * 4;
                if (k < k_end && m < M) v = __ldg(reinterpret_cast<const float4*>(a + (size_t)k * lda + m));
                *reinterpret_cast<float4*>(&As[t >> 3][(t & 7) * 4]) = v;
            } else {                                    // contiguous along k: thread = (m = t/8, 4 k), stored transposed
                const int m = m0 + (t >> 3), k = k0 + (t & 7) * 4;
                if (m < M && k < k_end) v = __ldg(reinterpret_cast<const float4*>(a + (size_t)m * lda + k));
                const int kk = (t & 7) * 4, mm = t >> 3;
                As[kk][mm] = v.x; As[kk + 1][mm] = v.y; As[kk + 2][mm] = v.z; As[kk + 3][mm] = v.w;
            }
        }
        // ---- B tile -> Bs[k][n]  (2048 elements: two float4 per thread)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (TB) {                                   // contiguous along k: thread = (n = t/8 + 32h, 4 k), stored transposed
                const int nn = (t >> 3) + 32 * h, kk = (t & 7) * 4;
                if (n0 + nn < N && k0 + kk < k_end) v = __ldg(reinterpret_cast<const float4*>(b + (size_t)(n0 + nn) * ldb + k0 + kk));
                Bs[kk][nn] = v.x; Bs[kk + 1][nn] = v.y; Bs[kk + 2][nn] = v.z; Bs[kk + 3][nn] = v.w;
            } else {                                    // contiguous along n: thread = (k = t/16 + 16h, 4 n)
                const int kk = (t >> 4) + 16 * h, nn = (t & 15) * 4;
                if (k0 + kk < k_end && n0 + nn < N) v = __ldg(reinterpret_cast<const float4*>(b + (size_t)(k0 + kk) * ldb + n0 + nn));
                *reinterpret_cast<float4*>(&Bs[kk][nn]) = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kSgBK; ++k) {
            const float a0 = As[k][ty * 2], a1 = As[k][ty * 2 + 1];
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            acc[0][0] += a0 * b4.x; acc[0][1] += a0 * b4.y; acc[0][2] += a0 * b4.z; acc[0][3] += a0 * b4.w;
            acc[1][0] += a1 * b4.x; acc[1][1] += a1 * b4.y; acc[1][2] += a1 * b4.z; acc[1][3] += a1 * b4.w;
        }
        if (rowsum != nullptr && blockIdx.x == 0 && t < kSgBM) {
#pragma unroll
            for (int k = 0; k < kSgBK; ++k) rs += As[k][t];
        }
        __syncthreads();
    }
    if (rowsum != nullptr && blockIdx.x == 0 && t < kSgBM && m0 + t < M) atomicAdd(rowsum + m0 + t, rs);
    const int n = n0 + tx * 4;
    if (n >= N) return;
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias != nullptr && blockIdx.z == 0) b4 = __ldg(reinterpret_cast<const float4*>(bias + n));
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int m = m0 + ty * 2 + i;
        if (m >= M) continue;
        float* o = c + (size_t)m * ldc + n;
        const float v0 = acc[i][0] + b4.x, v1 = acc[i][1] + b4.y, v2 = acc[i][2] + b4.z, v3 = acc[i][3] + b4.w;
        if (atomic) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(v0), "f"(v1), "f"(v2), "f"(v3) : "memory");
        else *reinterpret_cast<float4*>(o) = make_float4(v0, v1, v2, v3);
    }
}

static bool small_gemm_ok(const void* p0, const void* p1, const void* p2, const void* p3, int M, int N, int K) {
    const uintptr_t bits = reinterpret_cast<uintptr_t>(p0) | reinterpret_cast<uintptr_t>(p1) | reinterpret_cast<uintptr_t>(p2) |
                           reinterpret_cast<uintptr_t>(p3);
    // worth it from a few MFLOP up (the classifier head and other slivers stay on the warp-per-output kernels)
    return (bits & 15) == 0 && (M & 3) == 0 && (N & 3) == 0 && (K & 3) == 0 && (long long)M * N * K >= (1ll << 21) && N >= 64;
}

// C (=|+=) A B as described above.  `overwrite`: C is replaced (zeroed first when K is split), else accumulated into.
template <bool TA, bool TB>
static int launch_small_gemm(const float* a, int lda, const float* b, int ldb, float* c, int ldc, const float* bias, float* rowsum,
                             int M, int N, int K, bool overwrite, cudaStream_t stream) {
    const int gx = (N + kSgBN - 1) / kSgBN, gy = (M + kSgBM - 1) / kSgBM;
    const int kb = (K + kSgBK - 1) / kSgBK;
    int splits = (2 * sm_count() + gx * gy - 1) / (gx * gy);
    if (splits > kb) splits = kb;
    if (splits < 1) splits = 1;
    const int k_per_split = ((kb + splits - 1) / splits) * kSgBK;
    splits = (K + k_per_split - 1) / k_per_split;
    TAVK_CHECK(gy <= 65535 && splits <= 65535, 2, "small gemm: shape too large for this kernel (M=%d K=%d)", M, K);
    const int atomic = (!overwrite || splits > 1) ? 1 : 0;
    if (overwrite && splits > 1) TAVK_CUDA(cudaMemsetAsync(c, 0, (size_t)M * ldc * sizeof(float), stream));
    TAVK_CUDA(launch_kernel(small_gemm_f32_kernel<TA, TB>, dim3(gx, gy, splits), dim3(256), (size_t)(0), stream, a, lda, b, ldb,
                            c, ldc, bias, rowsum, M, N, K, k_per_split, atomic));
    TAVK_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------- casts / scaling / dropout
__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    for (long long i = t; i < n4; i += stride) {
        const float4 v = reinterpret_cast<const float4*>(x)[i];
        reinterpret_cast<uint2*>(y)[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
    for (long long i = (n4 << 2) + t; i < n; i += stride) y[i] = __float2bfloat16_rn(x[i]);
}
__global__ void scale_f32_kernel(const float* __restrict__ x, float* __restrict__ y, float scale, long long n) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) y[i] = x[i] * scale;
}
TAVK_DEVINL float uniform01(uint64_t seed, uint64_t ctr) {
    // splitmix64 finaliser over a Weyl sequence: counter-based, reproducible for (seed, offset + index)
    uint64_t z = seed + (ctr + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}
__global__ void dropout_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, uint8_t* __restrict__ keep,
                                   long long n, float p, float inv_keep, uint64_t seed, uint64_t offset,
                                   const uint64_t* __restrict__ offset_dev) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    if (offset_dev != nullptr) offset += (*offset_dev) << 32;  // device-side step counter (CUDA-graph replays)
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        const bool k = uniform01(seed, offset + (uint64_t)i) >= p;
        keep[i] = k ? 1 : 0;
        y[i] = k ? x[i] * inv_keep : 0.f;
    }
}
__global__ void dropout_bwd_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ keep,
                                   float* __restrict__ dx, long long n, float inv_keep) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride)
        dx[i] = keep[i] ? dy[i] * inv_keep : 0.f;
}

__global__ void dropout_bwd_add_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ keep,
                                       const float* __restrict__ resid, float* __restrict__ dx, long long n,
                                       float inv_keep) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride)
        dx[i] = resid[i] + (keep[i] ? dy[i] * inv_keep : 0.f);
}

// [B,S,nh,d] -> [B,nh,d,S] (the reference MultiHeadAttention "concat" quirk, utils/TAVFormer.py:86)
// via a 32x32 smem transpose of the (S, d) plane of every (b, h); `inverse` maps back (used in backward).
__global__ void permute_bshd_bhds_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int S,
                                         int nh, int d, int inverse) {
    pdl_wait();   // programmatic dependent launch: see common.cuh
    __shared__ __nv_bfloat16 tile[32][33];
    const int bh = blockIdx.z;
    const int b = bh / nh, h = bh % nh;
    const int s0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
    if (!inverse) {
        for (int j = ty; j < 32; j += 8) {
            const int s = s0 + j, dd = d0 + tx;
            if (s < S && dd < d) tile[j][tx] = in[(((size_t)b * S + s) * nh + h) * d + dd];
        }
        __syncthreads();
        for (int j = ty; j < 32; j += 8) {
            const int dd = d0 + j, s = s0 + tx;
            if (s < S && dd < d) out[(((size_t)b * nh + h) * d + dd) * S + s] = tile[tx][j];
        }
    } else {
        for (int j = ty; j < 32; j += 8) {
            const int dd = d0 + j, s = s0 + tx;
            if (s < S && dd < d) tile[tx][j] = in[(((size_t)b * nh + h) * d + dd) * S + s];
        }
        __syncthreads();
        for (int j = ty; j < 32; j += 8) {
            const int s = s0 + j, dd = d0 + tx;
            if (s < S && dd < d) out[(((size_t)b * S + s) * nh + h) * d + dd] = tile[j][tx];
        }
    }
}

static inline int grid_for(long long n, int threads) {
    long long g = (n + threads - 1) / threads;
    const long long cap = (long long)sm_count() * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace tavk

using namespace tavk;
#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int tavk_embed_add_fwd(const float* x, const int64_t* idx, const float* table, float* y, int rows, int H,
                                  int n_embed, void* stream) {
    TAVK_CHECK(x && idx && table && y, 1, "tavk_embed_add_fwd: null pointer");
    TAVK_CHECK(H % 4 == 0 && n_embed >= 1, 1, "tavk_embed_add_fwd: H=%d must be a multiple of 4", H);
    if (rows <= 0) return 0;
    const long long total = (long long)rows * (H / 4);
    TAVK_CUDA(launch_kernel(embed_add_fwd_kernel, dim3(grid_for(total, 256)), dim3(256), (size_t)(0), STREAM(stream), 
        reinterpret_cast<const float4*>(x), idx, reinterpret_cast<const float4*>(table), reinterpret_cast<float4*>(y),
        rows, H / 4, n_embed));
    TAVK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tavk_embed_add_bwd(const float* dy, const int64_t* idx, float* dtable, int rows, int H, int n_embed,
                                  void* stream) {
    TAVK_CHECK(dy && idx && dtable, 1, "tavk_embed_add_bwd: null pointer");
    TAVK_CHECK(n_embed >= 1 && n_embed <= 8, 2, "tavk_embed_add_bwd: n_embed=%d unsupported (1..8)", n_embed);
    if (rows <= 0) return 0;
    const int threads = 128;
    const int gx = (H + threads - 1) / threads;
    int chunks = (2 * sm_count() + gx - 1) / gx;
    int rpc = (rows + chunks - 1) / chunks;
    if (rpc < 16) rpc = 16;
    chunks = (rows + rpc - 1) / rpc;
    TAVK_CUDA(launch_kernel(embed_add_bwd_kernel, dim3(dim3(gx, chunks)), dim3(threads), (size_t)(0), STREAM(stream), dy, idx, dtable, rows, H, n_embed, rpc));
    TAVK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tavk_masked_mean_pool_fwd(const float* x, const int* lengths, float* y, int B, int S, int H, void* stream);
extern "C" int tavk_roberta_embed_fwd(const int64_t* ids, const float* word, const float* pos, const float* type0,
                                      float* y, int64_t* pos_ids, int B, int T, int H, int vocab, int n_pos, int pad_id,
                                      void* stream) {
    TAVK_CHECK(ids && word && pos && type0 && y && pos_ids, 1, "tavk_roberta_embed_fwd: null pointer");
    TAVK_CHECK(H % 4 == 0 && T >= 1 && T <= 8192, 1, "tavk_roberta_embed_fwd: bad shape T=%d H=%d", T, H);
    TAVK_CHECK(((reinterpret_cast<uintptr_t>(word) | reinterpret_cast<uintptr_t>(pos) | reinterpret_cast<uintptr_t>(type0) |
                 reinterpret_cast<uintptr_t>(y)) & 15) == 0, 1, "tavk_roberta_embed_fwd: tables and output must be 16-byte aligned");
    if (B <= 0) return 0;
    TAVK_CUDA(launch_kernel(roberta_embed_fwd_kernel, dim3(B), dim3(256), (size_t)T * sizeof(int), STREAM(stream), ids,
                            reinterpret_cast<const float4*>(word), reinterpret_cast<const float4*>(pos),
                            reinterpret_cast<const float4*>(type0), reinterpret_cast<float4*>(y), pos_ids, T, H / 4, vocab,
                            n_pos, pad_id));
    TAVK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tavk_embedding_scatter_add(const float* dy, const int64_t* idx, float* dtable, int rows, int H, int n_embed,
                                          int skip_idx, void* stream) {
    TAVK_CHECK(dy && idx && dtable, 1, "tavk_embedding_scatter_add: null pointer");
    TAVK_CHECK(H % 4 == 0, 1, "tavk_embedding_scatter_add: H=%d must be a multiple of 4", H);
    TAVK_CHECK(((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dtable)) & 15) == 0, 1,
               "tavk_embedding_scatter_add: buffers must be 16-byte aligned");
    if (rows <= 0) return 0;
    const long long total = (long long)rows * (H / 4);
    TAVK_CUDA(launch_kernel(embedding_scatter_add_kernel, dim3(grid_for(total, 256)), dim3(256), (size_t)(0), STREAM(stream), dy, idx, dtable,
                            rows, H / 4, n_embed, skip_idx));
    TAVK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tavk_mean_pool_fwd(const float* x, float* y, int B, int S, int H, void* stream) {
    return tavk_masked_mean_pool_fwd(x, nullptr, y, B, S, H, stream);
}
extern "C" int tavk_masked_mean_pool_fwd(const float* x, const int* lengths, float* y, int B, int S, int H, void* stream) {
    TAVK_CHECK(x && y, 1, "tavk_mean_pool_fwd: null pointer");
    TAVK_CHECK(H % 4 == 0, 1, "tavk_mean_pool_fwd: H=%d must be a multiple of 4", H);
    if (B <= 0) return 0;
    TAVK_CUDA(cudaMemsetAsync(y, 0, (size_t)B * H * sizeof(float), STREAM(stream)));
    if (S <= 0) return 0;
    const int threads = 64;
    const int gx = (H / 4 + threads - 1) / threads;
    int chunks = (2 * sm_count() + gx * B - 1) / (gx * B);
    if (chunks < 1) chunks = 1;
    int rpc = (S + chunks - 1) / chunks;
    if (rpc < 8) rpc = 8;
    chunks = (S + rpc - 1) / rpc;
    TAVK_CUDA(launch_kernel(mean_pool_fwd_kernel, dim3(dim3(gx, chunks, B)), dim3(threads), (size_t)(0), STREAM(stream), reinterpret_cast<const float4*>(x), y, S,
                                                                            H / 4, rpc, 1.0f / (float)S, lengths));
    TAVK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tavk_masked_mean_pool_bwd(const float* dy, const int* lengths, float* dx, void* dx_bf16, int B, int S, int H,
                                         void* stream);
extern "C" int tavk_mean_pool_bwd(const float* dy, float* dx, void* dx_bf16, int B, int S, int H, void* stream) {
    return tavk_masked_mean_pool_bwd(dy, nullptr, dx, dx_bf16, B, S, H, stream);
}
extern "C" int tavk_masked_mean_pool_bwd(const float* dy, const int* lengths, float* dx, void* dx_bf16, int B, int S, int H,
                                         void* stream) {
    TAVK_CHECK(dy && (dx || dx_bf16), 1, "tavk_mean_pool_bwd: null pointer");
    TAVK_CHECK(H % 4 == 0, 1, "tavk_mean_pool_bwd: H=%d must be a multiple of 4", H);
    if (B <= 0 || S <= 0) return 0;
    const long long total = (long long)B * S * (H / 4);
    TAVK_CUDA(launch_kernel(mean_pool_bwd_kernel, dim3(grid_for(total, 256)), dim3(256), (size_t)(0), STREAM(stream), 
        reinterpret_cast<const float4*>(dy), reinterpret_cast<float4*>(dx), reinterpret_cast<uint2*>(dx_bf16), S, H / 4,
        total, 1.0f / (float)S, lengths));
    TAVK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tavk_colsum(const void* x, int x_dtype, int64_t ld, float* out, int M, int N, int accumulate,
                           void* stream) {
    return launch_colsum(x, x_dtype, ld, nullptr, out, 1, M, N, accumulate, STREAM(stream));
}

extern "C" int tavk_masked_colsum(const void* x, int x_dtype, int64_t ld, const float* w, float* out, int B, int S,
                                  int N, void* stream) {
    return launch_colsum(x, x_dtype, ld, w, out, B, S, N, 0, STREAM(stream));
}

extern "C" int tavk_small_linear_fwd(const float* x, const float* w, const float* b, float* y, int M, int N, int K,
                                     void* stream) {
    TAVK_CHECK(x && w && y, 1, "tavk_small_linear_fwd: null pointer");
    TAVK_CHECK((K & 3) != 0 || ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w)) & 15) == 0, 1,
               "tavk_small_linear_fwd: x and w must be 16-byte aligned when K %% 4 == 0");
    if (M <= 0 || N <= 0) return 0;
    if (small_gemm_ok(x, w, y, b, M, N, K))           // y = x w^T + b as a tiled GEMM (A = x, B(k,n) = w[n,k])
        return launch_small_gemm<false, true>(x, K, w, K, y, N, b, nullptr, M, N, K, true, STREAM(stream));
    const int gx = (N + 7) / 8;                       // 8 warps (output columns) per block
    const int gy = (M + kSlRows - 1) / kSlRows;
    TAVK_CHECK(gy <= 65535, 2, "tavk_small_linear_fwd: M=%d too large for this kernel", M);
    TAVK_CUDA(launch_kernel(small_linear_fwd_kernel, dim3(gx, gy), dim3(256), (size_t)(0), STREAM(stream), x, w, b, y, M, N, K));
    TAVK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tavk_small_linear_bwd_x(const float* dy, const float* w, float* dx, int M, int N, int K, int accumulate,
                                       void* stream) {
    TAVK_CHECK(dy && w && dx, 1, "tavk_small_linear_bwd_x: null pointer");
    if (M <= 0 || K <= 0) return 0;
    if (N > 0 && small_gemm_ok(dy, w, dx, nullptr, M, K, N))   // dx[M,K] (+)= dy[M,N] w[N,K]
        return launch_small_gemm<false, false>(dy, N, w, K, dx, K, nullptr, nullptr, M, K, N, !accumulate, STREAM(stream));
    const long long total = (long long)M * K;
    if (!accumulate) TAVK_CUDA(cudaMemsetAsync(dx, 0, (size_t)total * sizeof(float), STREAM(stream)));
    if (N <= 0) return 0;
    const int gz = (M + kSlRows - 1) / kSlRows;
    TAVK_CHECK(gz <= 65535, 2, "tavk_small_linear_bwd_x: M=%d too large for this kernel", M);
    TAVK_CUDA(launch_kernel(small_linear_bwd_x_kernel, dim3((K + 127) / 128, (N + kSlChunk - 1) / kSlChunk, gz), dim3(128),
                            (size_t)(0), STREAM(stream), dy, w, dx, M, N, K));
    TAVK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tavk_small_linear_bwd_w(const float* dy, const float* x, float* dw, float* db, int M, int N, int K,
                                       void* stream) {
    TAVK_CHECK(dy && x && (dw || db), 1, "tavk_small_linear_bwd_w: null pointer");
    TAVK_CHECK((K & 3) != 0 || ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dw)) & 15) == 0, 1,
               "tavk_small_linear_bwd_w: x and dw must be 16-byte aligned when K %% 4 == 0");
    if (N <= 0 || K <= 0) return 0;
    if (dw != nullptr && M > 0 && small_gemm_ok(dy, x, dw, nullptr, N, K, M))    // dw[N,K] += dy^T[N,M] x[M,K]; db[N] += row sums of dy^T
        return launch_small_gemm<true, false>(dy, N, x, K, dw, K, nullptr, db, N, K, M, false, STREAM(stream));
    const long long total = (long long)N * ((K & 3) == 0 ? K / 4 : K);
    TAVK_CUDA(launch_kernel(small_linear_bwd_w_kernel, dim3((int)((total + 127) / 128)), dim3(128), (size_t)(0), STREAM(stream), dy, x, dw, db, M, N, K));
    TAVK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tavk_cast_f32_bf16(const float* x, void* y, int64_t n, void* stream) {
    TAVK_CHECK(x && y, 1, "tavk_cast_f32_bf16: null pointer");
    TAVK_CHECK((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 7) == 0, 1,
               "tavk_cast_f32_bf16: buffers must be 16/8-byte aligned");
    if (n <= 0) return 0;
    TAVK_CUDA(launch_kernel(cast_f32_bf16_kernel, dim3(grid_for((n + 3) / 4, 256)), dim3(256), (size_t)(0), STREAM(stream), x, reinterpret_cast<__nv_bfloat16*>(y),
                                                                                 n));
    TAVK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tavk_scale_f32(const float* x, float* y, float scale, int64_t n, void* stream) {
    TAVK_CHECK(x && y, 1, "tavk_scale_f32: null pointer");
    if (n <= 0) return 0;
    TAVK_CUDA(launch_kernel(scale_f32_kernel, dim3(grid_for(n, 256)), dim3(256), (size_t)(0), STREAM(stream), x, y, scale, n));
    TAVK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tavk_dropout(const float* x, float* y, uint8_t* keep_mask, int64_t n, float p, uint64_t seed,
                            uint64_t offset, const uint64_t* offset_dev, void* stream) {
    TAVK_CHECK(x && y && keep_mask, 1, "tavk_dropout: null pointer");
    TAVK_CHECK(p >= 0.f && p < 1.f, 1, "tavk_dropout: p=%f out of [0,1)", (double)p);
    if (n <= 0) return 0;
    TAVK_CUDA(launch_kernel(dropout_fwd_kernel, dim3(grid_for(n, 256)), dim3(256), (size_t)(0), STREAM(stream), x, y, keep_mask, n, p, 1.0f / (1.0f - p), seed,
                                                                     offset, offset_dev));
    TAVK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tavk_dropout_bwd(const float* dy, const uint8_t* keep_mask, float* dx, int64_t n, float p,
                                void* stream) {
    TAVK_CHECK(dy && dx && keep_mask, 1, "tavk_dropout_bwd: null pointer");
    if (n <= 0) return 0;
    TAVK_CUDA(launch_kernel(dropout_bwd_kernel, dim3(grid_for(n, 256)), dim3(256), (size_t)(0), STREAM(stream), dy, keep_mask, dx, n, 1.0f / (1.0f - p)));
    TAVK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tavk_dropout_bwd_add(const float* dy, const uint8_t* keep_mask, const float* resid, float* dx, int64_t n,
                                    float p, void* stream) {
    TAVK_CHECK(dy && dx && keep_mask && resid, 1, "tavk_dropout_bwd_add: null pointer");
    TAVK_CHECK(p >= 0.f && p < 1.f, 1, "tavk_dropout_bwd_add: p=%f out of [0,1)", (double)p);
    if (n <= 0) return 0;
    TAVK_CUDA(launch_kernel(dropout_bwd_add_kernel, dim3(grid_for(n, 256)), dim3(256), (size_t)(0), STREAM(stream), dy, keep_mask, resid, dx, n,
                            1.0f / (1.0f - p)));
    TAVK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tavk_permute_bshd_bhds(const void* in, void* out, int B, int S, int nh, int d, int inverse,
                                      void* stream) {
    TAVK_CHECK(in && out, 1, "tavk_permute_bshd_bhds: null pointer");
    if (B <= 0 || S <= 0) return 0;
    dim3 grid((S + 31) / 32, (d + 31) / 32, B * nh);
    TAVK_CUDA(launch_kernel(permute_bshd_bhds_kernel, dim3(grid), dim3(dim3(32, 8)), (size_t)(0), STREAM(stream), reinterpret_cast<const __nv_bfloat16*>(in),
                                                                       reinterpret_cast<__nv_bfloat16*>(out), S, nh, d,
                                                                       inverse));
    TAVK_CUDA(cudaGetLastError());
    return 0;
}
