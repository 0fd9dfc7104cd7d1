// CUDA-core fp32 GEMM templates of the exact-fp32 path (MFT_PREC_FP32) and of
// every small GEMM around the edge MLP.  Three shapes:
//
//   gemm_rows_kernel : C[r, n] = sum_k A(r,k) * W(k,n)      r over pair rows / nodes
//                      (forward layers and dgrad).  A is produced element-wise by a
//                      functor (abs-diff of node features, BN+LeakyReLU of a stored
//                      pre-BN tensor, plain load), the epilogue is a functor too
//                      (store + batch statistics, dy + statistics, dx scatter ...).
//   gemm_tn_kernel   : C[m, n] += sum_r P(r,m) * Q(r,n)      (wgrad; split over rows,
//                      fp32 atomics into a zeroed destination).
//   bgemm_kernel     : small batched strided GEMM for the per-graph products
//                      (A x, A^T dU, dU x^T).
//
// Tile: 128 rows x 96 columns x 16 deep, 256 threads, 8x6 accumulators per thread,
// register double buffering.  None of this is meant to approach the tensor-core
// path; it is the fp32 yardstick and the fallback-free path for odd shapes.
#pragma once

#include "common.cuh"

namespace mft {

constexpr int GM_BM = 128;
constexpr int GM_BN = 96;
constexpr int GM_BK = 16;
constexpr int GM_THREADS = 256;

// Weight operand view: elem(k, n) = w[(n % nmod) * sn + (n / nmod) * soff + k * sk].
// nmod/soff express the [n_out, 2F] -> [2 n_out, F] split of the Gconv fc weight.
struct WView {
    const float* w;
    int sn, sk, nmod, soff;
    __device__ __forceinline__ float at(int k, int n) const {
        int q = n / nmod;
        int m = n - q * nmod;
        return __ldg(w + (size_t)m * sn + (size_t)q * soff + (size_t)k * sk);
    }
};
inline WView wview_nt(const float* w, int ld) { return WView{w, ld, 1, 1 << 30, 0}; }   // W[n][k]
inline WView wview_nn(const float* w, int ld) { return WView{w, 1, ld, 1 << 30, 0}; }   // W[k][n]

struct GemmSmem {
    float As[2][GM_BK][GM_BM + 4];
    float Bs[2][GM_BK][GM_BN + 2];
    float aux_a[4 * kMaxC];
    float aux_e[4 * kMaxC];
};

template <bool KContig, class AOp, class Epi>
__global__ void __launch_bounds__(GM_THREADS)
gemm_rows_kernel(AOp aop, WView bv, Epi epi, int M, int N, int K) {
    __shared__ GemmSmem sm;
    const int t = threadIdx.x;
    const int tx = t & 15, ty = t >> 4;
    const int row_base = blockIdx.x * GM_BM;
    const int col_base = blockIdx.y * GM_BN;

    aop.init(sm.aux_a);
    epi.init(sm.aux_e);
    __syncthreads();

    const int a_row = t & (GM_BM - 1);
    const int a_k0 = (t >> 7) * 8;
    const bool a_valid = (row_base + a_row) < M;
    typename AOp::Ctx actx = aop.row(a_valid ? row_base + a_row : 0);

    float ra[8], rb[6];
    float acc[8][6];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) acc[i][j] = 0.f;

    auto fetch = [&](int kt) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            int k = kt + a_k0 + q;
            ra[q] = (a_valid && k < K) ? aop.at(actx, k, sm.aux_a) : 0.f;
        }
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            int idx = t + GM_THREADS * c;
            int kk = KContig ? (idx & (GM_BK - 1)) : (idx / GM_BN);
            int n = KContig ? (idx >> 4) : (idx % GM_BN);
            rb[c] = (kt + kk < K && col_base + n < N) ? bv.at(kt + kk, col_base + n) : 0.f;
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int q = 0; q < 8; ++q) sm.As[buf][a_k0 + q][a_row] = ra[q];
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            int idx = t + GM_THREADS * c;
            int kk = KContig ? (idx & (GM_BK - 1)) : (idx / GM_BN);
            int n = KContig ? (idx >> 4) : (idx % GM_BN);
            sm.Bs[buf][kk][n] = rb[c];
        }
    };

    fetch(0);
    stash(0);
    __syncthreads();
    int buf = 0;
    for (int kt = 0; kt < K; kt += GM_BK) {
        const bool more = kt + GM_BK < K;
        if (more) fetch(kt + GM_BK);
#pragma unroll
        for (int kk = 0; kk < GM_BK; ++kk) {
            float a[8], b[6];
            float4 a0 = *reinterpret_cast<const float4*>(&sm.As[buf][kk][ty * 8]);
            float4 a1 = *reinterpret_cast<const float4*>(&sm.As[buf][kk][ty * 8 + 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
            a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
            for (int c2 = 0; c2 < 3; ++c2) {
                float2 bb = *reinterpret_cast<const float2*>(&sm.Bs[buf][kk][tx * 2 + 32 * c2]);
                b[2 * c2] = bb.x;
                b[2 * c2 + 1] = bb.y;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 6; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (more) {
            stash(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }

    // ---- epilogue: thread owns rows row_base + ty*8 + i, columns col_base + tx*2 + 32*c2 + e
    float s0[6], s1[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) { s0[j] = 0.f; s1[j] = 0.f; }
    epi.tile(row_base + ty * 8, col_base + tx * 2, M, N, acc, s0, s1, sm.aux_e);

    if (Epi::kStats) {
        __syncthreads();   // everyone is done with As/Bs
        float* red0 = &sm.As[0][0][0];                 // [16][96]
        float* red1 = red0 + 16 * GM_BN;               // [16][96]  (2*16*96 <= 2*16*132)
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            int cl = tx * 2 + 32 * (j >> 1) + (j & 1);
            red0[ty * GM_BN + cl] = s0[j];
            red1[ty * GM_BN + cl] = s1[j];
        }
        __syncthreads();
        if (t < GM_BN && col_base + t < N) {
            float v0 = 0.f, v1 = 0.f;
#pragma unroll
            for (int y = 0; y < 16; ++y) {
                v0 += red0[y * GM_BN + t];
                v1 += red1[y * GM_BN + t];
            }
            epi.commit(col_base + t, v0, v1);
        }
    }
}

template <bool KContig, class AOp, class Epi>
inline cudaError_t launch_gemm_rows(const AOp& aop, const WView& bv, const Epi& epi, int M, int N, int K,
                                    cudaStream_t st) {
    if (M <= 0 || N <= 0) return cudaSuccess;
    dim3 grid(cdiv(M, GM_BM), cdiv(N, GM_BN));
    gemm_rows_kernel<KContig, AOp, Epi><<<grid, GM_THREADS, 0, st>>>(aop, bv, epi, M, N, K);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// wgrad: out[m, n] += sum_r P(r, m) * Q(r, n), r in this CTA's row chunk.
// ---------------------------------------------------------------------------
struct TnSmem {
    float As[2][GM_BK][GM_BM + 4];
    float Bs[2][GM_BK][GM_BN + 2];
    float aux_p[4 * kMaxC];
    float aux_q[4 * kMaxC];
};

template <class POp, class QOp>
__global__ void __launch_bounds__(GM_THREADS)
gemm_tn_kernel(POp pop, QOp qop, float* __restrict__ out, int ldo, int M, int N, int R, int rows_per_cta) {
    __shared__ TnSmem sm;
    const int t = threadIdx.x;
    const int tx = t & 15, ty = t >> 4;
    const int m_base = blockIdx.y * GM_BM;
    const int n_base = blockIdx.x * GM_BN;
    const int r_begin = blockIdx.z * rows_per_cta;
    const int r_end = min(R, r_begin + rows_per_cta);

    pop.init(sm.aux_p);
    qop.init(sm.aux_q);
    __syncthreads();

    float ra[8], rb[6];
    float acc[8][6];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) acc[i][j] = 0.f;

    auto fetch = [&](int r0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            int idx = t + GM_THREADS * q;
            int m = idx & (GM_BM - 1);
            int r = r0 + (idx >> 7);
            ra[q] = (r < r_end && m_base + m < M) ? pop.at(r, m_base + m, sm.aux_p) : 0.f;
        }
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            int idx = t + GM_THREADS * c;
            int n = idx % GM_BN;
            int r = r0 + idx / GM_BN;
            rb[c] = (r < r_end && n_base + n < N) ? qop.at(r, n_base + n, sm.aux_q) : 0.f;
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            int idx = t + GM_THREADS * q;
            sm.As[buf][idx >> 7][idx & (GM_BM - 1)] = ra[q];
        }
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            int idx = t + GM_THREADS * c;
            sm.Bs[buf][idx / GM_BN][idx % GM_BN] = rb[c];
        }
    };

    if (r_begin >= r_end) return;
    fetch(r_begin);
    stash(0);
    __syncthreads();
    int buf = 0;
    for (int r0 = r_begin; r0 < r_end; r0 += GM_BK) {
        const bool more = r0 + GM_BK < r_end;
        if (more) fetch(r0 + GM_BK);
#pragma unroll
        for (int kk = 0; kk < GM_BK; ++kk) {
            float a[8], b[6];
            float4 a0 = *reinterpret_cast<const float4*>(&sm.As[buf][kk][ty * 8]);
            float4 a1 = *reinterpret_cast<const float4*>(&sm.As[buf][kk][ty * 8 + 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
            a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
            for (int c2 = 0; c2 < 3; ++c2) {
                float2 bb = *reinterpret_cast<const float2*>(&sm.Bs[buf][kk][tx * 2 + 32 * c2]);
                b[2 * c2] = bb.x;
                b[2 * c2 + 1] = bb.y;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 6; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (more) {
            stash(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int m = m_base + ty * 8 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            int n = n_base + tx * 2 + 32 * (j >> 1) + (j & 1);
            if (n < N) atomicAdd(out + (size_t)m * ldo + n, acc[i][j]);
        }
    }
}

template <class POp, class QOp>
inline cudaError_t launch_gemm_tn(const POp& pop, const QOp& qop, float* out, int ldo, int M, int N, int R,
                                  cudaStream_t st) {
    if (M <= 0 || N <= 0 || R <= 0) return cudaSuccess;
    // enough row chunks to fill the chip a few times, each a multiple of the K tile
    int tiles = cdiv(M, GM_BM) * cdiv(N, GM_BN);
    int want = max(1, (148 * 3) / tiles);
    int rows_per_cta = max(GM_BK * 4, cdiv(cdiv(R, want), GM_BK) * GM_BK);
    dim3 grid(cdiv(N, GM_BN), cdiv(M, GM_BM), cdiv(R, rows_per_cta));
    gemm_tn_kernel<POp, QOp><<<grid, GM_THREADS, 0, st>>>(pop, qop, out, ldo, M, N, R, rows_per_cta);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Small batched strided GEMM: C[b](m,n) = beta*C + sum_k A[b](m,k) * B[b](k,n).
// ---------------------------------------------------------------------------
struct BView {
    const float* p;
    long sb, s0, s1;   // batch stride, stride of the first index, of the second
};

__global__ void bgemm_kernel(BView A, BView Bm, float* __restrict__ C, long scb, int ldc, int M, int N, int K,
                             float beta);

// C[b] = beta * C[b] + A[b] B[b] + addend[b]; the addend (batch stride sadd, row stride ldadd) is only
// available where bgemm_takes_addend(K) (the one-shot kernel).
bool bgemm_takes_addend(int K);
cudaError_t launch_bgemm(BView A, BView Bm, float* C, long scb, int ldc, int batch, int M, int N, int K,
                         float beta, cudaStream_t st, const float* addend = nullptr, long sadd = 0, int ldadd = 0);
// C[b] = beta * C[b] + A[b] B[b] + A2[b] B2[b] in one launch, where bgemm2_supported(K, K2)
bool bgemm2_supported(int K, int K2);
cudaError_t launch_bgemm2(BView A, BView Bm, int K, BView A2, BView B2, int K2, float* C, long scb, int ldc,
                          int batch, int M, int N, float beta, cudaStream_t st);

}  // namespace mft
