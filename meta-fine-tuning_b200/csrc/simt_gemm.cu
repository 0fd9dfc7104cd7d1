// Non-template parts of the CUDA-core GEMM toolkit (see simt_gemm.cuh).
#include "simt_gemm.cuh"

namespace mft {

constexpr int BG_T = 32;

__global__ void __launch_bounds__(256)
bgemm_kernel(BView A, BView Bm, float* __restrict__ C, long scb, int ldc, int M, int N, int K, float beta) {
    // 32x32 output tile, K chunks of 32, next chunk prefetched into registers while the current one
    // is multiplied (these products are small and latency bound: 105 <= K <= 1680)
    __shared__ float As[2][BG_T][BG_T + 1];   // [m][k]
    __shared__ float Bs[2][BG_T][BG_T + 1];   // [k][n]
    const int b = blockIdx.z;
    const int m0 = blockIdx.y * BG_T, n0 = blockIdx.x * BG_T;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const float* Ab = A.p + (size_t)b * A.sb;
    const float* Bb = Bm.p + (size_t)b * Bm.sb;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    float ra[4], rb[4];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int r = ty + 8 * q;
            int m = m0 + r, k = k0 + tx;
            ra[q] = (m < M && k < K) ? __ldg(Ab + (size_t)m * A.s0 + (size_t)k * A.s1) : 0.f;
            int kk = k0 + r, n = n0 + tx;
            rb[q] = (kk < K && n < N) ? __ldg(Bb + (size_t)kk * Bm.s0 + (size_t)n * Bm.s1) : 0.f;
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            As[buf][ty + 8 * q][tx] = ra[q];
            Bs[buf][ty + 8 * q][tx] = rb[q];
        }
    };
    fetch(0);
    stash(0);
    __syncthreads();
    int buf = 0;
    for (int k0 = 0; k0 < K; k0 += BG_T) {
        const bool more = k0 + BG_T < K;
        if (more) fetch(k0 + BG_T);
#pragma unroll
        for (int k = 0; k < BG_T; ++k) {
            float bv = Bs[buf][k][tx];
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] = fmaf(As[buf][ty + 8 * q][k], bv, acc[q]);
        }
        if (more) {
            stash(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        int m = m0 + ty + 8 * q, n = n0 + tx;
        if (m < M && n < N) {
            float* c = C + (size_t)b * scb + (size_t)m * ldc + n;
            *c = (beta == 0.f) ? acc[q] : fmaf(beta, *c, acc[q]);
        }
    }
}

// 64x64 output tile, 4x4 outputs per thread, K chunks of 16 (two LDS.128 per 16 FMAs instead of five
// LDS per four): used whenever the problem has at least one full-ish tile in both dimensions.
constexpr int BG2_T = 64;
constexpr int BG2_K = 16;

__global__ void __launch_bounds__(256)
bgemm64_kernel(BView A, BView Bm, float* __restrict__ C, long scb, int ldc, int M, int N, int K, float beta) {
    __shared__ __align__(16) float As[2][BG2_K][BG2_T + 4];   // [k][m]
    __shared__ __align__(16) float Bs[2][BG2_K][BG2_T + 4];   // [k][n]
    const int b = blockIdx.z;
    const int m0 = blockIdx.y * BG2_T, n0 = blockIdx.x * BG2_T;
    const int t = threadIdx.x;
    const int tx = t & 15, ty = t >> 4;          // 16 x 16 threads, each a 4x4 block
    const float* Ab = A.p + (size_t)b * A.sb;
    const float* Bb = Bm.p + (size_t)b * Bm.sb;
    // loader mapping: 64 x 16 elements per operand per chunk = 4 per thread.  Pick the thread
    // layout that walks the operand's unit-stride index fastest.
    const bool a_k_fast = A.s1 == 1;             // A(m,k): k contiguous
    const bool b_n_fast = Bm.s1 == 1;            // B(k,n): n contiguous
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float ra[4], rb[4];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int idx = t + 256 * q;                              // 0..1023
            int am = a_k_fast ? (idx >> 4) : (idx & 63);
            int ak = a_k_fast ? (idx & 15) : (idx >> 6);
            int m = m0 + am, k = k0 + ak;
            ra[q] = (m < M && k < K) ? __ldg(Ab + (size_t)m * A.s0 + (size_t)k * A.s1) : 0.f;
            int bn = b_n_fast ? (idx & 63) : (idx >> 4);
            int bk = b_n_fast ? (idx >> 6) : (idx & 15);
            int n = n0 + bn, kk = k0 + bk;
            rb[q] = (kk < K && n < N) ? __ldg(Bb + (size_t)kk * Bm.s0 + (size_t)n * Bm.s1) : 0.f;
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int idx = t + 256 * q;
            int am = a_k_fast ? (idx >> 4) : (idx & 63);
            int ak = a_k_fast ? (idx & 15) : (idx >> 6);
            As[buf][ak][am] = ra[q];
            int bn = b_n_fast ? (idx & 63) : (idx >> 4);
            int bk = b_n_fast ? (idx >> 6) : (idx & 15);
            Bs[buf][bk][bn] = rb[q];
        }
    };
    fetch(0);
    stash(0);
    __syncthreads();
    int buf = 0;
    for (int k0 = 0; k0 < K; k0 += BG2_K) {
        const bool more = k0 + BG2_K < K;
        if (more) fetch(k0 + BG2_K);
#pragma unroll
        for (int k = 0; k < BG2_K; ++k) {
            float4 a = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            float4 bb = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (more) {
            stash(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n < N) {
                float* c = C + (size_t)b * scb + (size_t)m * ldc + n;
                *c = (beta == 0.f) ? acc[i][j] : fmaf(beta, *c, acc[i][j]);
            }
        }
    }
}

// One-shot variant for K <= 512 (every Gconv product except the weight gradients, and GnnNet's fc): a 32x32 output
// tile whose complete A and B panels are brought into shared memory by cp.async in ONE round of
// loads (the looped kernels above pay one L2 round trip per 16-32 columns of K with nothing to
// overlap it: 27-CTA launches that took 12-17 us), then multiplied from shared memory.
constexpr int OS_T = 32;
constexpr int OS_MAXK = 512;                        // 2 x 32 x 513 floats = 131 KB of panels at most

__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src, bool valid) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst_smem));
    const int sz = valid ? 4 : 0;      // src-size 0: the destination is zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}

// Panels of one (A, B) pair into shared memory, columns [kofs, kofs + K) of the concatenated K axis.
__device__ __forceinline__ void os_fill(const BView& A, const BView& Bm, int b, int m0, int n0, int M, int N, int K,
                                        int kofs, int lda, float* As, float* Bs, bool b_n_fast) {
    const int t = threadIdx.x;
    const float* Ab = A.p + (size_t)b * A.sb;
    const float* Bb = Bm.p + (size_t)b * Bm.sb;
    const int total = OS_T * K;
    const int lane = t & 31, wrp = t >> 5;
    if (A.s1 == 1) {                             // k contiguous in memory: a warp walks along a row (no division)
        for (int m = wrp; m < OS_T; m += 8) {
            const bool ok = m0 + m < M;
            const float* src = Ab + (size_t)(ok ? m0 + m : 0) * A.s0;
            for (int k = lane; k < K; k += 32) cp_async4(As + m * lda + kofs + k, src + k, ok);
        }
    } else {                                     // m contiguous (or general strides)
        for (int idx = t; idx < total; idx += 256) {
            const int k = idx >> 5, m = idx & 31;
            const bool ok = m0 + m < M;
            cp_async4(As + m * lda + kofs + k, Ab + (size_t)(ok ? m0 + m : 0) * A.s0 + (size_t)k * A.s1, ok);
        }
    }
    if (b_n_fast) {                              // n contiguous
        for (int idx = t; idx < total; idx += 256) {
            const int k = idx >> 5, n = idx & 31;
            const bool ok = n0 + n < N;
            cp_async4(Bs + (kofs + k) * OS_T + n, Bb + (size_t)k * Bm.s0 + (size_t)(ok ? n0 + n : 0) * Bm.s1, ok);
        }
    } else {                                     // k contiguous (or general strides): a warp walks along a column;
        for (int n = wrp; n < OS_T; n += 8) {    // the panel is kept [n][k] (odd stride) so that neither these
            const bool ok = n0 + n < N;          // writes nor the reads below conflict on banks
            const float* src = Bb + (size_t)(ok ? n0 + n : 0) * Bm.s1;
            for (int k = lane; k < K; k += 32) cp_async4(Bs + n * lda + kofs + k, src + (size_t)k * Bm.s0, ok);
        }
    }
}

// C = beta C + A B (+ A2 B2) (+ addend): the second pair extends the K axis (K2 = 0: none), so a sum of two
// products costs one launch and one pass over C.
__global__ void __launch_bounds__(256)
bgemm_oneshot_kernel(BView A, BView Bm, float* __restrict__ C, long scb, int ldc, int M, int N, int K, float beta,
                     const float* __restrict__ addend, long sadd, int ldadd, BView A2, BView B2, int K2) {
    extern __shared__ float os_smem[];
    const int KT = K + K2;
    const int lda = KT | 1;                      // odd row stride: conflict-free column walks
    float* As = os_smem;                         // [32][lda]  (m, k)
    float* Bs = os_smem + OS_T * lda;            // [KT][32]   (k, n)   or [32][lda] (n, k)
    const int b = blockIdx.z;
    const int m0 = blockIdx.y * OS_T, n0 = blockIdx.x * OS_T;
    const int t = threadIdx.x;
    const bool b_n_fast = Bm.s1 == 1 && (K2 == 0 || B2.s1 == 1);
    os_fill(A, Bm, b, m0, n0, M, N, K, 0, lda, As, Bs, b_n_fast);
    if (K2 > 0) os_fill(A2, B2, b, m0, n0, M, N, K2, K, lda, As, Bs, b_n_fast);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const int tx = t & 15, ty = t >> 4;          // thread -> outputs (2*ty + {0,1}, 2*tx + {0,1})
    const float* a0 = As + (2 * ty) * lda;
    const float* a1 = a0 + lda;
    float c00 = 0.f, c01 = 0.f, c10 = 0.f, c11 = 0.f;
    if (b_n_fast) {                              // B panel [k][n]
        const float* bp = Bs + 2 * tx;
#pragma unroll 8
        for (int k = 0; k < KT; ++k) {
            const float x0 = a0[k], x1 = a1[k];
            const float2 y = *reinterpret_cast<const float2*>(bp + k * OS_T);
            c00 = fmaf(x0, y.x, c00); c01 = fmaf(x0, y.y, c01);
            c10 = fmaf(x1, y.x, c10); c11 = fmaf(x1, y.y, c11);
        }
    } else {                                     // B panel [n][k]
        const float* b0 = Bs + (2 * tx) * lda;
        const float* b1 = b0 + lda;
#pragma unroll 8
        for (int k = 0; k < KT; ++k) {
            const float x0 = a0[k], x1 = a1[k];
            const float y0 = b0[k], y1 = b1[k];
            c00 = fmaf(x0, y0, c00); c01 = fmaf(x0, y1, c01);
            c10 = fmaf(x1, y0, c10); c11 = fmaf(x1, y1, c11);
        }
    }
    const float acc[2][2] = {{c00, c01}, {c10, c11}};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int m = m0 + 2 * ty + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int n = n0 + 2 * tx + j;
            if (n < N) {
                float* c = C + (size_t)b * scb + (size_t)m * ldc + n;
                float v = acc[i][j];
                if (addend) v += addend[(size_t)b * sadd + (size_t)m * ldadd + n];
                *c = (beta == 0.f) ? v : fmaf(beta, *c, v);
            }
        }
    }
}

bool bgemm_takes_addend(int K) { return K >= 1 && K <= OS_MAXK; }

static cudaError_t launch_oneshot(BView A, BView Bm, float* C, long scb, int ldc, int batch, int M, int N, int K,
                                  float beta, cudaStream_t st, const float* addend, long sadd, int ldadd, BView A2,
                                  BView B2, int K2) {
    static bool attr_set = false;
    const size_t max_smem = (size_t)(2 * OS_T * (OS_MAXK | 1)) * sizeof(float);
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(bgemm_oneshot_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)max_smem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const size_t smem = (size_t)(2 * OS_T * ((K + K2) | 1)) * sizeof(float);
    dim3 grid(cdiv(N, OS_T), cdiv(M, OS_T), batch);
    bgemm_oneshot_kernel<<<grid, 256, smem, st>>>(A, Bm, C, scb, ldc, M, N, K, beta, addend, sadd, ldadd, A2, B2, K2);
    return cudaGetLastError();
}

// C[b] = beta * C[b] + A[b] B[b] + A2[b] B2[b]: both products in one launch (K + K2 <= 512)
cudaError_t launch_bgemm2(BView A, BView Bm, int K, BView A2, BView B2, int K2, float* C, long scb, int ldc,
                          int batch, int M, int N, float beta, cudaStream_t st) {
    if (batch <= 0 || M <= 0 || N <= 0) return cudaSuccess;
    if (K < 1 || K2 < 1 || K + K2 > OS_MAXK) return cudaErrorInvalidValue;
    return launch_oneshot(A, Bm, C, scb, ldc, batch, M, N, K, beta, st, nullptr, 0, 0, A2, B2, K2);
}
bool bgemm2_supported(int K, int K2) { return K >= 1 && K2 >= 1 && K + K2 <= OS_MAXK; }

// C[b] = beta * C[b] + A[b] B[b] (+ addend[b], one-shot path only: see bgemm_takes_addend)
cudaError_t launch_bgemm(BView A, BView Bm, float* C, long scb, int ldc, int batch, int M, int N, int K,
                         float beta, cudaStream_t st, const float* addend, long sadd, int ldadd) {
    if (batch <= 0 || M <= 0 || N <= 0) return cudaSuccess;
    if (addend && !bgemm_takes_addend(K)) return cudaErrorInvalidValue;
    if (K >= 1 && K <= OS_MAXK)
        return launch_oneshot(A, Bm, C, scb, ldc, batch, M, N, K, beta, st, addend, sadd, ldadd, BView{}, BView{}, 0);
    if (M >= 48 && N >= 40) {
        dim3 grid(cdiv(N, BG2_T), cdiv(M, BG2_T), batch);
        bgemm64_kernel<<<grid, 256, 0, st>>>(A, Bm, C, scb, ldc, M, N, K, beta);
        return cudaGetLastError();
    }
    dim3 grid(cdiv(N, BG_T), cdiv(M, BG_T), batch);
    bgemm_kernel<<<grid, 256, 0, st>>>(A, Bm, C, scb, ldc, M, N, K, beta);
    return cudaGetLastError();
}

}  // namespace mft
