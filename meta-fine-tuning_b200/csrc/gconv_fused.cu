// One-launch Gconv forward (gnn.py:16-56: gmul with J = 2, Linear, BatchNorm1d, LeakyReLU).
//
// The Gconv arithmetic is small (a 1680 x 229 x 96 product, a batched 105 x 105 x 48 product, two statistics
// over 1680 rows); what it cost was its LAUNCH CHAIN: six dependent launches of 5-10 us on 16-60 CTAs each, on
// the critical path between two edge MLPs -- a quarter of the 5w20s step (profiles/r02_summary.md, main-stream
// spans).  Here one kernel does all of it: CTA (slab, graph) owns 32 nodes of one graph,
//
//   phase 1   [U | V] = x_slab [Wa | Wb]^T                  (32 x 2 n_out outputs, K = F through shared memory)
//             V goes to global memory; a per-GRAPH counter tells when all slabs of the graph have written theirs
//   phase 2   Y = U + A_slab V_graph + b ;  per-column sum / sum of squares -> fp64 atomics (kStatCopies copies)
//             a GRID-wide counter tells when every CTA has contributed
//   phase 3   out = LeakyReLU(BN(Y)) with the batch statistics over all B*N rows (skipped without BatchNorm)
//
// The counters are spin barriers in global memory: every CTA of the launch must be resident at once (the host
// checks grid <= SMs x occupancy) and every wait is bounded (trap after ~2 s), as for the mbarrier waits of the
// tensor-core kernels.  Pre-BN Y and the statistics slot are saved in the layout the backward expects.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "wcompute.cuh"
#include "prof.cuh"

namespace mft {

constexpr int GF_ROWS = 32;          // nodes per CTA
constexpr int GF_THREADS = 256;
constexpr int GF_KT = 32;            // K tile of phase 1
constexpr int GF_MAX_OUT = 64;       // n_out <= 64 (the reference: 48 and n_way)
constexpr int GF_MAX_N = 192;        // nodes per graph

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// All threads call it; returns when `*cnt` has reached `target` (every participating CTA has arrived).
__device__ __forceinline__ void spin_barrier(int* cnt, int target) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicAdd(cnt, 1);
        long long t0 = clock64();
        while (ld_acquire_gpu(cnt) < target) {
            __nanosleep(40);
            if (clock64() - t0 > 4000000000LL) {
                printf("mft: gconv spin barrier timed out (block %d,%d target %d)\n", (int)blockIdx.x, (int)blockIdx.y, target);
                __trap();
            }
        }
    }
    __syncthreads();
}

struct GconvFusedArgs {
    const float* adj;      // [B, N, N]
    const float* x;        // [B*N, ldx]
    const float* fc_w;     // [n_out, 2F]
    const float* fc_b;     // [n_out]
    const float* bn_g;     // [n_out] or null (no BatchNorm: layer_last)
    const float* bn_b;
    float* out;            // [B*N, ldo]
    float* Y;              // saved pre-BN output [B*N, n_out] (BatchNorm only)
    float* V;              // workspace [B*N, n_out]
    double* fsums;         // statistics slot (zeroed before the launch)
    int* sync;             // [B + 1] counters (zeroed before the launch)
    int ldx, ldo, B, N, F, n_out, lrelu_on;
};

__global__ void __launch_bounds__(GF_THREADS)
gconv_fused_fwd_kernel(const GconvFusedArgs a) {
    extern __shared__ float sm[];
    const int tid = threadIdx.x;
    const int tx = tid & 31, ty = tid >> 5;                    // ty: 4 rows each
    const int b = blockIdx.y, S = gridDim.x;
    const int n0 = blockIdx.x * GF_ROWS;
    const int nrows = min(GF_ROWS, a.N - n0);
    const int n_out = a.n_out, F = a.F, N = a.N;
    const int ncol2 = 2 * n_out;
    // shared memory: phase 1: xs[32][KT+1], ws[2 n_out][KT+1]; phase 2: us[32][n_out+1], vs[N][n_out], as[32][N+1]
    float* us = sm;                                            // [GF_ROWS][GF_MAX_OUT + 1]
    float* un = us + GF_ROWS * (GF_MAX_OUT + 1);               // union region
    float* xs = un;                                            // [GF_ROWS][GF_KT + 1]
    float* ws = xs + GF_ROWS * (GF_KT + 1);                    // [2 * GF_MAX_OUT][GF_KT + 1]
    float* vs = un;                                            // [N][n_out]
    float* as = vs + (size_t)N * n_out;                        // [GF_ROWS][N + 1]

    // ---------------- phase 1: [U | V] = x_slab [Wa | Wb]^T
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const float* xrow0 = a.x + (size_t)(b * N + n0) * a.ldx;
    for (int k0 = 0; k0 < F; k0 += GF_KT) {
        for (int idx = tid; idx < GF_ROWS * GF_KT; idx += GF_THREADS) {
            const int r = idx >> 5, k = idx & 31;
            xs[r * (GF_KT + 1) + k] = (r < nrows && k0 + k < F) ? __ldg(xrow0 + (size_t)r * a.ldx + k0 + k) : 0.f;
        }
        for (int idx = tid; idx < ncol2 * GF_KT; idx += GF_THREADS) {
            const int c = idx >> 5, k = idx & 31;
            // column c < n_out: Wa[c][k] = fc_w[c][k]; c >= n_out: Wb[c - n_out][k] = fc_w[c - n_out][F + k]
            const float* wrow = c < n_out ? a.fc_w + (size_t)c * 2 * F : a.fc_w + (size_t)(c - n_out) * 2 * F + F;
            ws[c * (GF_KT + 1) + k] = (k0 + k < F) ? __ldg(wrow + k0 + k) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < GF_KT; ++k) {
            float xv[4], wv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) xv[i] = xs[(ty * 4 + i) * (GF_KT + 1) + k];
#pragma unroll
            for (int j = 0; j < 4; ++j) wv[j] = (tx + 32 * j < ncol2) ? ws[(tx + 32 * j) * (GF_KT + 1) + k] : 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xv[i], wv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = ty * 4 + i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = tx + 32 * j;
            if (c < n_out) us[r * (GF_MAX_OUT + 1) + c] = acc[i][j];
            else if (c < ncol2 && r < nrows) a.V[(size_t)(b * N + n0 + r) * n_out + (c - n_out)] = acc[i][j];
        }
    }
    spin_barrier(a.sync + b, S);                               // every slab of this graph has written its V rows

    // ---------------- phase 2: Y = U + A_slab V_graph + bias
    for (int idx = tid; idx < N * n_out; idx += GF_THREADS) vs[idx] = a.V[(size_t)b * N * n_out + idx];
    for (int idx = tid; idx < GF_ROWS * N; idx += GF_THREADS) {
        const int r = idx / N, j = idx - r * N;
        as[r * (N + 1) + j] = r < nrows ? __ldg(a.adj + ((size_t)b * N + n0 + r) * N + j) : 0.f;
    }
    __syncthreads();
    float y[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) { y[i][0] = 0.f; y[i][1] = 0.f; }
    const bool c1 = tx + 32 < n_out, c0 = tx < n_out;
    for (int j = 0; j < N; ++j) {
        const float v0 = c0 ? vs[j * n_out + tx] : 0.f;
        const float v1 = c1 ? vs[j * n_out + tx + 32] : 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float av = as[(ty * 4 + i) * (N + 1) + j];
            y[i][0] = fmaf(av, v0, y[i][0]);
            y[i][1] = fmaf(av, v1, y[i][1]);
        }
    }
    const float b0 = c0 ? __ldg(a.fc_b + tx) : 0.f, b1 = c1 ? __ldg(a.fc_b + tx + 32) : 0.f;
    float p0[2] = {0.f, 0.f}, p1[2] = {0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = ty * 4 + i;
        if (c0) y[i][0] += us[r * (GF_MAX_OUT + 1) + tx] + b0;
        if (c1) y[i][1] += us[r * (GF_MAX_OUT + 1) + tx + 32] + b1;
        if (r < nrows) {
            p0[0] += y[i][0]; p1[0] += y[i][0] * y[i][0];
            p0[1] += y[i][1]; p1[1] += y[i][1] * y[i][1];
        }
    }
    const bool has_bn = a.bn_g != nullptr;
    if (!has_bn) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = ty * 4 + i;
            if (r < nrows) {
                float* o = a.out + (size_t)(b * N + n0 + r) * a.ldo;
                if (c0) o[tx] = y[i][0];
                if (c1) o[tx + 32] = y[i][1];
            }
        }
        return;
    }
    // column statistics: 8 row groups -> shared memory -> one fp64 atomic per column per CTA (kStatCopies copies)
    __syncthreads();                                           // vs / as no longer needed: reuse the union region
    float* red = un;                                           // [8][2][GF_MAX_OUT]
    if (c0) { red[(ty * 2 + 0) * GF_MAX_OUT + tx] = p0[0]; red[(ty * 2 + 1) * GF_MAX_OUT + tx] = p1[0]; }
    if (c1) { red[(ty * 2 + 0) * GF_MAX_OUT + tx + 32] = p0[1]; red[(ty * 2 + 1) * GF_MAX_OUT + tx + 32] = p1[1]; }
    __syncthreads();
    const int copy = (blockIdx.y * gridDim.x + blockIdx.x) % kStatCopies;
    if (tid < 2 * n_out) {
        const int which = tid / n_out, c = tid - which * n_out;
        float v = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) v += red[(q * 2 + which) * GF_MAX_OUT + c];
        atomicAdd(a.fsums + copy * kStatCopyStride + which * n_out + c, (double)v);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {                              // pre-BN output, saved for the backward
        const int r = ty * 4 + i;
        if (r < nrows) {
            float* o = a.Y + (size_t)(b * N + n0 + r) * n_out;
            if (c0) o[tx] = y[i][0];
            if (c1) o[tx + 32] = y[i][1];
        }
    }
    spin_barrier(a.sync + a.B, S * a.B);                       // every CTA has added its partial statistics

    // ---------------- phase 3: BatchNorm (batch statistics over all B*N rows) + LeakyReLU
    float* bn = un + 8 * 2 * GF_MAX_OUT;                       // [2][GF_MAX_OUT]: scale, shift
    if (tid < n_out) {
        float m, r;
        bn_mean_rstd(a.fsums, n_out, tid, 1.0 / ((double)a.B * (double)N), m, r);
        const float sc = __ldg(a.bn_g + tid) * r;
        bn[tid] = sc;
        bn[GF_MAX_OUT + tid] = __ldg(a.bn_b + tid) - m * sc;
    }
    __syncthreads();
    const float sc0 = c0 ? bn[tx] : 0.f, sh0 = c0 ? bn[GF_MAX_OUT + tx] : 0.f;
    const float sc1 = c1 ? bn[tx + 32] : 0.f, sh1 = c1 ? bn[GF_MAX_OUT + tx + 32] : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = ty * 4 + i;
        if (r < nrows) {
            float* o = a.out + (size_t)(b * N + n0 + r) * a.ldo;
            float z0 = fmaf(y[i][0], sc0, sh0), z1 = fmaf(y[i][1], sc1, sh1);
            if (a.lrelu_on) { z0 = lrelu(z0); z1 = lrelu(z1); }
            if (c0) o[tx] = z0;
            if (c1) o[tx + 32] = z1;
        }
    }
}

static size_t gconv_fused_smem(int N, int n_out) {
    const size_t p1 = (size_t)GF_ROWS * (GF_KT + 1) + (size_t)2 * GF_MAX_OUT * (GF_KT + 1);
    const size_t p2 = (size_t)N * n_out + (size_t)GF_ROWS * (N + 1);
    const size_t p3 = (size_t)8 * 2 * GF_MAX_OUT + 2 * GF_MAX_OUT;
    size_t un = p1 > p2 ? p1 : p2;
    if (p3 > un) un = p3;
    return ((size_t)GF_ROWS * (GF_MAX_OUT + 1) + un) * sizeof(float);
}

// 1 = the fused kernel may run for this shape on this device (every CTA resident at once), 0 = use the launch chain
int gconv_fused_supported(int B, int N, int F, int n_out) {
    // Opt-in (MFT_GCONV_FUSED=1): correct (same parity tests), but on B200 the single launch takes ~40 us against
    // the ~30 us critical path of the launch chain it replaces (5w20s step 1.37 vs 1.32 ms, profiles/r02_summary.md):
    // phase 1 runs on 64 CTAs with single-buffered K tiles.  Kept as the starting point of a faster one.
    static const int env = [] { const char* e = getenv("MFT_GCONV_FUSED"); return e ? atoi(e) : 0; }();
    if (!env || n_out > GF_MAX_OUT || N > GF_MAX_N || F < 1) return 0;
    const size_t smem = gconv_fused_smem(N, n_out);
    if (smem > 200 * 1024) return 0;
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    int per_sm = 0;
    if (cudaFuncSetAttribute(gconv_fused_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gconv_fused_fwd_kernel, GF_THREADS, smem) != cudaSuccess)
        return 0;
    return cdiv(N, GF_ROWS) * B <= per_sm * sms ? 1 : 0;
}

int gconv_fused_fwd(const float* adj, const float* x, int ldx, int B, int N, int F, int n_out, const mft_gconv_params* p,
                    int lrelu_on, float* out, int ldo, float* Y, float* V, double* fsums, int* sync, cudaStream_t st) {
    GconvFusedArgs a;
    a.adj = adj; a.x = x; a.fc_w = p->fc_w; a.fc_b = p->fc_b; a.bn_g = p->bn_g; a.bn_b = p->bn_b;
    a.out = out; a.Y = Y; a.V = V; a.fsums = fsums; a.sync = sync;
    a.ldx = ldx; a.ldo = ldo; a.B = B; a.N = N; a.F = F; a.n_out = n_out; a.lrelu_on = lrelu_on;
    const size_t smem = gconv_fused_smem(N, n_out);
    MFT_CHECK_CUDA(cudaFuncSetAttribute(gconv_fused_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope ps(PC_GCONV_FWD, st);
    gconv_fused_fwd_kernel<<<dim3(cdiv(N, GF_ROWS), B), GF_THREADS, smem, st>>>(a);
    MFT_CHECK_LAUNCH();
    return MFT_OK;
}

}  // namespace mft
