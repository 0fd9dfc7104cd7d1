// Gconv (J=2 aggregation + Linear + BatchNorm1d) forward and backward.
// Replaces gmul + Gconv.forward of the reference (methods/gnn.py:16-28, 43-56).
// The identity operator of gmul is never multiplied out, and A (x Wb^T) is
// evaluated as written there instead of (A x) Wb^T: same result, the N x N
// product runs on n_out (48 or n_way) columns instead of F (133..229).  The backward
// uses the same association (see gconv_bwd), which is why x Wb^T is part of the saved state.
#include "common.cuh"
#include "simt_gemm.cuh"
#include "wcompute.cuh"
#include "prof.cuh"

namespace mft {

namespace {

struct PlainOp {
    const float* p;
    int ld;
    __device__ __forceinline__ void init(float*) const {}
    __device__ __forceinline__ float at(int r, int k, const float*) const { return p[(size_t)r * ld + k]; }
};

constexpr int kGcRows = 4;    // rows (nodes) per CTA in the per-node kernels (8 / 16 measured: within noise)
constexpr int kGcCols = 64;   // threads along the output channel

// Y[r, c] += bias[c]; optional BatchNorm1d batch statistics (sum y, sum y^2) over the B*N rows.
__global__ void __launch_bounds__(kGcRows * kGcCols)
gconv_bias_stats_kernel(float* __restrict__ Y, int ldy, const float* __restrict__ bias, int rows, int n_out,
                        double* sums) {
    __shared__ float red[2][kGcRows][kMaxC];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int row = blockIdx.x * kGcRows + ty;
    const bool live = row < rows;
    for (int c = tx; c < n_out; c += kGcCols) {
        float y = 0.f;
        if (live) {
            y = Y[(size_t)row * ldy + c] + bias[c];
            Y[(size_t)row * ldy + c] = y;
        }
        red[0][ty][c] = y;
        red[1][ty][c] = y * y;
    }
    if (sums == nullptr) return;
    __syncthreads();
    if (ty == 0) {
        for (int c = tx; c < n_out; c += kGcCols) {
            float v0 = 0.f, v1 = 0.f;
#pragma unroll
            for (int q = 0; q < kGcRows; ++q) { v0 += red[0][q][c]; v1 += red[1][q][c]; }
            stat_add(sums, n_out, c, 0, v0);
            stat_add(sums, n_out, c, 1, v1);
        }
    }
}

// out = act(BN1d(Y))
__global__ void gconv_apply_kernel(const float* __restrict__ Y, int rows, int n_out, const double* sums,
                                   const float* gamma, const float* beta, int lrelu_on, float* __restrict__ out,
                                   int ldo) {
    __shared__ float aux[4 * kMaxC];
    bn_smem_fill(bn_smem_at(aux), sums, gamma, beta, n_out, 1.0 / (double)rows);
    __syncthreads();
    BnSmem s = bn_smem_at(aux);
    int total = rows * n_out;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        int r = idx / n_out, c = idx - r * n_out;
        float hh = (Y[idx] - s.mean[c]) * s.rstd[c];
        float z = fmaf(hh, s.gamma[c], s.beta[c]);
        out[(size_t)r * ldo + c] = lrelu_on ? lrelu(z) : z;
    }
}

// dz = d_out * lrelu'(z) -> dY buffer, with column sums of dz and dz*hhat
__global__ void __launch_bounds__(kGcRows * kGcCols)
gconv_dz_kernel(const float* __restrict__ d_out, int ldo, const float* __restrict__ Y, int rows, int n_out,
                const double* fsums, const float* gamma, const float* beta, int has_bn, int lrelu_on,
                float* __restrict__ dY, double* bsums) {
    __shared__ float aux[4 * kMaxC];
    __shared__ float red[2][kGcRows][kMaxC];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int t = ty * kGcCols + tx;
    if (has_bn) {
        for (int c = t; c < n_out; c += kGcRows * kGcCols) {
            float m, r;
            bn_mean_rstd(fsums, n_out, c, 1.0 / (double)rows, m, r);
            aux[c] = m; aux[kMaxC + c] = r; aux[2 * kMaxC + c] = gamma[c]; aux[3 * kMaxC + c] = beta[c];
        }
    }
    __syncthreads();
    BnSmem s = bn_smem_at(aux);
    const int row = blockIdx.x * kGcRows + ty;
    const bool live = row < rows;
    for (int c = tx; c < n_out; c += kGcCols) {
        float d = 0.f, dh = 0.f;
        if (live) {
            d = d_out[(size_t)row * ldo + c];
            if (has_bn) {
                float hh = (Y[(size_t)row * n_out + c] - s.mean[c]) * s.rstd[c];
                float z = fmaf(hh, s.gamma[c], s.beta[c]);
                if (lrelu_on) d *= dlrelu(z);
                dh = d * hh;
            }
            dY[(size_t)row * n_out + c] = d;
        }
        red[0][ty][c] = d;
        red[1][ty][c] = dh;
    }
    __syncthreads();
    if (ty == 0) {
        for (int c = tx; c < n_out; c += kGcCols) {
            float v0 = 0.f, v1 = 0.f;
#pragma unroll
            for (int q = 0; q < kGcRows; ++q) { v0 += red[0][q][c]; v1 += red[1][q][c]; }
            stat_add(bsums, n_out, c, 0, v0);
            stat_add(bsums, n_out, c, 1, v1);
        }
    }
}

// BatchNorm1d backward in place on dY, plus the small parameter gradients.
__global__ void gconv_dy_kernel(float* __restrict__ dY, const float* __restrict__ Y, int rows, int n_out,
                                const double* fsums, const float* gamma, const double* bsums) {
    __shared__ float aux[4 * kMaxC];
    __shared__ float m1[kMaxC], m2[kMaxC];
    bn_smem_fill(bn_smem_at(aux), fsums, gamma, nullptr, n_out, 1.0 / (double)rows);
    for (int c = threadIdx.x; c < n_out; c += blockDim.x) {
        m1[c] = (float)(stat_get(bsums, n_out, c, 0) / (double)rows);
        m2[c] = (float)(stat_get(bsums, n_out, c, 1) / (double)rows);
    }
    __syncthreads();
    BnSmem s = bn_smem_at(aux);
    int total = rows * n_out;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        int c = idx % n_out;
        float hh = (Y[idx] - s.mean[c]) * s.rstd[c];
        dY[idx] = s.gamma[c] * s.rstd[c] * (dY[idx] - m1[c] - hh * m2[c]);
    }
}

__global__ void gconv_small_grads_kernel(const double* bsums, int n_out, int has_bn, float* fc_b, float* bn_g,
                                         float* bn_b) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_out) return;
    if (has_bn) {
        if (bn_b) bn_b[c] = (float)stat_get(bsums, n_out, c, 0);
        if (bn_g) bn_g[c] = (float)stat_get(bsums, n_out, c, 1);
        if (fc_b) fc_b[c] = 0.f;                   // BatchNorm1d removes the mean: exactly zero
    } else if (fc_b) {
        fc_b[c] = (float)stat_get(bsums, n_out, c, 0);
    }
}

// ---------------------------------------------------------------------------------------------
// Pre-head of GnnNet: fc = Linear(feat_dim -> D) + BatchNorm1d over the n_way*(n_support+n_query)
// episode rows (gnnnet.py:30, 71-78), then the assembly of the n_query graphs (gnnnet.py:79-83: every
// graph = all supports + the q-th query of each class) with the support one-hot labels appended
// (gnnnet.py:35-38, 212).  The assembly is pure data movement; here BN-apply writes straight into the
// [n_query, n_way*(n_support+1), D+n_way] node tensor, and the backward gathers the node gradients
// back onto the episode rows (sum over the graphs for a support row) in the kernel that also forms
// the BatchNorm-backward reductions.
// ---------------------------------------------------------------------------------------------
__global__ void head_assemble_kernel(const float* __restrict__ Y, int D, const double* sums, const float* gamma,
                                     const float* beta, int n_way, int n_support, int n_query,
                                     float* __restrict__ nodes) {
    __shared__ float aux[4 * kMaxC];
    const int rows = n_way * (n_support + n_query);
    bn_smem_fill(bn_smem_at(aux), sums, gamma, beta, D, 1.0 / (double)rows);
    __syncthreads();
    BnSmem s = bn_smem_at(aux);
    const int npg = n_way * (n_support + 1), ldn = D + n_way;
    const int total = n_query * npg;
    for (int node = blockIdx.x; node < total; node += gridDim.x) {
        const int q = node / npg, n = node - q * npg;
        const int c = n / (n_support + 1), sidx = n - c * (n_support + 1);
        const bool is_query = sidx == n_support;
        const int src = c * (n_support + n_query) + (is_query ? n_support + q : sidx);
        float* dst = nodes + (size_t)node * ldn;
        for (int f = threadIdx.x; f < ldn; f += blockDim.x) {
            float v;
            if (f < D) v = fmaf((Y[(size_t)src * D + f] - s.mean[f]) * s.rstd[f], s.gamma[f], s.beta[f]);
            else v = (!is_query && f - D == c) ? 1.f : 0.f;
            dst[f] = v;
        }
    }
}

// dz[row, f] = sum over the graphs that hold the row; column sums of dz and dz*zhat for BatchNorm backward
__global__ void __launch_bounds__(kGcRows * kGcCols)
head_dz_kernel(const float* __restrict__ d_nodes, const float* __restrict__ Y, int D, const double* fsums,
               int n_way, int n_support, int n_query, float* __restrict__ dY, double* bsums) {
    __shared__ float aux[2 * kMaxC];
    __shared__ float red[2][kGcRows][kMaxC];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int rows = n_way * (n_support + n_query);
    for (int c = ty * kGcCols + tx; c < D; c += kGcRows * kGcCols) {
        float m, r;
        bn_mean_rstd(fsums, D, c, 1.0 / (double)rows, m, r);
        aux[c] = m; aux[kMaxC + c] = r;
    }
    __syncthreads();
    const int npg = n_way * (n_support + 1), ldn = D + n_way;
    const int row = blockIdx.x * kGcRows + ty;
    const bool live = row < rows;
    const int c = live ? row / (n_support + n_query) : 0, sidx = live ? row - c * (n_support + n_query) : 0;
    for (int f = tx; f < D; f += kGcCols) {
        float d = 0.f, dh = 0.f;
        if (live) {
            if (sidx < n_support) {
                const int n = c * (n_support + 1) + sidx;
                for (int q = 0; q < n_query; ++q) d += d_nodes[((size_t)q * npg + n) * ldn + f];
            } else {
                const int q = sidx - n_support, n = c * (n_support + 1) + n_support;
                d = d_nodes[((size_t)q * npg + n) * ldn + f];
            }
            dh = d * ((Y[(size_t)row * D + f] - aux[f]) * aux[kMaxC + f]);
            dY[(size_t)row * D + f] = d;
        }
        red[0][ty][f] = d;
        red[1][ty][f] = dh;
    }
    __syncthreads();
    if (ty == 0) {
        for (int f = tx; f < D; f += kGcCols) {
            float v0 = 0.f, v1 = 0.f;
#pragma unroll
            for (int q = 0; q < kGcRows; ++q) { v0 += red[0][q][f]; v1 += red[1][q][f]; }
            stat_add(bsums, D, f, 0, v0);
            stat_add(bsums, D, f, 1, v1);
        }
    }
}

}  // namespace

struct HeadLayout {
    float* Y;        // saved: pre-BN fc output [rows, D]
    double* fsums;   // saved: BatchNorm1d statistics
    float* dY;       // workspace bwd: [rows, D]
    double* bsums;   // workspace bwd
    size_t saved_bytes, workspace_bytes;
};

static HeadLayout head_layout(int rows, int D, void* saved, void* workspace) {
    HeadLayout L;
    Carver sv(saved);
    L.Y = sv.take<float>((size_t)rows * D);
    L.fsums = sv.take<double>(kStatSlot);
    L.saved_bytes = sv.used();
    Carver ws(workspace);
    L.dY = ws.take<float>((size_t)rows * D);
    L.bsums = ws.take<double>(kStatSlot);
    L.workspace_bytes = ws.used();
    return L;
}

size_t head_saved_bytes(int rows, int D) { return head_layout(rows, D, nullptr, nullptr).saved_bytes; }
size_t head_workspace_bytes(int rows, int D) { return head_layout(rows, D, nullptr, nullptr).workspace_bytes; }

int head_fwd(const float* feat, int feat_dim, int n_way, int n_support, int n_query, int D,
             const mft_gconv_params* fc, float* nodes, void* saved, void* workspace, cudaStream_t st) {
    MFT_REQUIRE(n_way > 0 && n_support > 0 && n_query > 0 && feat_dim > 0 && D > 0, "head_fwd: bad shape");
    MFT_REQUIRE(D <= kMaxC, "head_fwd: D=%d exceeds %d", D, kMaxC);
    MFT_REQUIRE(fc->bn_g && fc->bn_b, "head_fwd: the fc pre-head has a BatchNorm1d (gnnnet.py:30)");
    const int rows = n_way * (n_support + n_query);
    HeadLayout L = head_layout(rows, D, saved, workspace);
    MFT_CHECK_CUDA(cudaMemsetAsync(L.fsums, 0, sizeof(double) * kStatSlot, st));
    {
        BView X{feat, 0, feat_dim, 1};               // (m = row, k)
        BView W{fc->fc_w, 0, 1, feat_dim};           // (k, n = c) -> fc_w[c*feat_dim + k]
        ProfScope ps(PC_GCONV_FWD, st);
        MFT_CHECK_CUDA(launch_bgemm(X, W, L.Y, 0, D, 1, rows, D, feat_dim, 0.f, st));
    }
    {
        ProfScope ps(PC_GCONV_FWD, st);
        gconv_bias_stats_kernel<<<cdiv(rows, kGcRows), dim3(kGcCols, kGcRows), 0, st>>>(L.Y, D, fc->fc_b, rows, D, L.fsums);
        MFT_CHECK_LAUNCH();
    }
    {
        ProfScope ps(PC_GCONV_FWD, st);
        const int total = n_query * n_way * (n_support + 1);
        head_assemble_kernel<<<min(total, 148 * 8), 160, 0, st>>>(L.Y, D, L.fsums, fc->bn_g, fc->bn_b, n_way, n_support,
                                                                   n_query, nodes);
        MFT_CHECK_LAUNCH();
    }
    return MFT_OK;
}

// Cross-entropy of the query nodes and its gradient in ONE launch (GnnNet.forward_gnn's score selection +
// set_forward_loss, gnnnet.py:216-224): graph q holds the q-th query of class c at node c*(n_support+1) +
// n_support, its label is c, the loss is the mean over the n_way*n_query query nodes.  d_out gets
// (softmax - onehot) / count on the query nodes and exact zeros on the support nodes.  One CTA, fixed
// summation order (a tree of fixed shape).  (The torch sequence -- permute/contiguous, log_softmax, nll_loss and their backwards,
// zero fill, index_put -- is nine launches in the middle of a 1.5 ms step.)
constexpr int kCeThreads = 1024;
__global__ void __launch_bounds__(kCeThreads)
query_ce_kernel(const float* __restrict__ out, int n_way, int n_support, int n_query, float* __restrict__ loss,
                float* __restrict__ d_out) {
    __shared__ float part[kCeThreads];
    const int npg = n_way * (n_support + 1);
    const int total = n_query * npg;
    const float inv = 1.f / (float)(n_way * n_query);
    float acc = 0.f;
    for (int node = threadIdx.x; node < total; node += blockDim.x) {
        const int n = node % npg;
        const int c = n / (n_support + 1), sidx = n - c * (n_support + 1);
        const float* z = out + (size_t)node * n_way;
        float* d = d_out + (size_t)node * n_way;
        if (sidx != n_support) {
            for (int k = 0; k < n_way; ++k) d[k] = 0.f;
            continue;
        }
        float m = z[0];
        for (int k = 1; k < n_way; ++k) m = fmaxf(m, z[k]);
        float se = 0.f;
        for (int k = 0; k < n_way; ++k) se += expf(z[k] - m);
        const float lse = m + logf(se);
        acc += lse - z[c];
        const float r = inv / se;
        for (int k = 0; k < n_way; ++k) d[k] = expf(z[k] - m) * r - (k == c ? inv : 0.f);
    }
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int half = kCeThreads / 2; half > 0; half >>= 1) {   // fixed-shape tree: same order every run
        if ((int)threadIdx.x < half) part[threadIdx.x] += part[threadIdx.x + half];
        __syncthreads();
    }
    if (threadIdx.x == 0) *loss = part[0] * inv;
}

int query_ce(const float* out, int n_way, int n_support, int n_query, float* loss, float* d_out, cudaStream_t st) {
    MFT_REQUIRE(n_way > 0 && n_support > 0 && n_query > 0, "query_ce: bad shape");
    ProfScope ps(PC_MISC, st);
    query_ce_kernel<<<1, kCeThreads, 0, st>>>(out, n_way, n_support, n_query, loss, d_out);
    MFT_CHECK_LAUNCH();
    return MFT_OK;
}

int head_bwd(const float* feat, int feat_dim, int n_way, int n_support, int n_query, int D,
             const mft_gconv_params* fc, const float* d_nodes, float* d_feat, const mft_gconv_grads* g, void* saved,
             void* workspace, cudaStream_t st) {
    MFT_REQUIRE(n_way > 0 && n_support > 0 && n_query > 0 && feat_dim > 0 && D > 0, "head_bwd: bad shape");
    MFT_REQUIRE(D <= kMaxC, "head_bwd: D=%d exceeds %d", D, kMaxC);
    const int rows = n_way * (n_support + n_query);
    HeadLayout L = head_layout(rows, D, saved, workspace);
    MFT_CHECK_CUDA(cudaMemsetAsync(L.bsums, 0, sizeof(double) * kStatSlot, st));
    MFT_CHECK_CUDA(cudaMemsetAsync(g->fc_w, 0, sizeof(float) * (size_t)D * feat_dim, st));
    {
        ProfScope ps(PC_GCONV_BWD, st);
        head_dz_kernel<<<cdiv(rows, kGcRows), dim3(kGcCols, kGcRows), 0, st>>>(d_nodes, L.Y, D, L.fsums, n_way, n_support,
                                                                               n_query, L.dY, L.bsums);
        MFT_CHECK_LAUNCH();
    }
    {
        const int total = rows * D;
        ProfScope ps(PC_GCONV_BWD, st);
        gconv_dy_kernel<<<min(cdiv(total, 256), 148 * 4), 256, 0, st>>>(L.dY, L.Y, rows, D, L.fsums, fc->bn_g, L.bsums);
        MFT_CHECK_LAUNCH();
    }
    {
        ProfScope ps(PC_GCONV_BWD, st);
        gconv_small_grads_kernel<<<cdiv(D, 128), 128, 0, st>>>(L.bsums, D, 1, g->fc_b, g->bn_g, g->bn_b);
        MFT_CHECK_LAUNCH();
    }
    {
        PlainOp dy{L.dY, D};
        PlainOp qx{feat, feat_dim};
        ProfScope ps(PC_GCONV_BWD, st);
        MFT_CHECK_CUDA((launch_gemm_tn(dy, qx, g->fc_w, feat_dim, D, feat_dim, rows, st)));
    }
    if (d_feat) {   // only when the features carry a gradient (meta-training through the backbone)
        BView Dy{L.dY, 0, D, 1};                     // (m = row, k = c)
        BView Wf{fc->fc_w, 0, feat_dim, 1};          // (k = c, n = f)
        ProfScope ps(PC_GCONV_BWD, st);
        MFT_CHECK_CUDA(launch_bgemm(Dy, Wf, d_feat, 0, feat_dim, 1, rows, feat_dim, D, 0.f, st));
    }
    return MFT_OK;
}

namespace {
}  // namespace

GcLayout gc_layout(int B, int N, int F, int n_out, void* saved, void* workspace) {
    GcLayout L;
    size_t rows = (size_t)B * N;
    Carver sv(saved);
    L.Y = sv.take<float>(rows * n_out);
    L.fsums = sv.take<double>(kStatSlot + 64);
    L.sync = reinterpret_cast<int*>(L.fsums + kStatSlot);
    L.XWb = sv.take<float>(rows * n_out);
    L.saved_bytes = sv.used();
    Carver ws(workspace);
    L.dY = ws.take<float>(rows * n_out);
    L.T = ws.take<float>(rows * n_out);
    L.bsums = ws.take<double>(kStatSlot);
    (void)F;
    L.workspace_bytes = ws.used();
    return L;
}

// Everything of the forward that needs no adjacency: x Wa^T -> Y (or `out` for the BN-less last layer),
// x Wb^T -> XWb, statistics slot cleared.  gnn_fwd runs it on a side stream beside the score / softmax kernels of
// the Wcompute that produces this layer's adjacency.
int gconv_fwd_products(const float* x, int ldx, int B, int N, int F, int n_out, const mft_gconv_params* p,
                       float* out, int ldo, void* saved, cudaStream_t st) {
    GcLayout L = gc_layout(B, N, F, n_out, saved, nullptr);
    const int rows = B * N;
    const bool has_bn = p->bn_g != nullptr;
    BView X{x, 0, ldx, 1};                           // (m = row, k = f)
    if (has_bn) {
        MFT_CHECK_CUDA(cudaMemsetAsync(L.fsums, 0, sizeof(double) * kStatSlot, st));
        // both products in one launch: "batch" h selects the half of fc.weight and the output buffer
        BView W{p->fc_w, (long)F, 1, 2 * F};         // (k = f, n = c) -> fc_w[c*2F + h*F + f]
        ProfScope ps(PC_GCONV_FWD, st);
        MFT_CHECK_CUDA(launch_bgemm(X, W, L.Y, (long)(L.XWb - L.Y), n_out, 2, rows, n_out, F, 0.f, st));
    } else {
        BView Wa{p->fc_w, 0, 1, 2 * F};
        BView Wb{p->fc_w + F, 0, 1, 2 * F};
        { ProfScope ps(PC_GCONV_FWD, st); MFT_CHECK_CUDA(launch_bgemm(X, Wa, out, 0, ldo, 1, rows, n_out, F, 0.f, st)); }
        { ProfScope ps(PC_GCONV_FWD, st); MFT_CHECK_CUDA(launch_bgemm(X, Wb, L.XWb, 0, n_out, 1, rows, n_out, F, 0.f, st)); }
    }
    return MFT_OK;
}

// Y += adj (x Wb^T); one pass adds the bias and accumulates the BatchNorm1d statistics, one applies BN + LeakyReLU.
int gconv_fwd_finish(const float* adj, int B, int N, int F, int n_out, const mft_gconv_params* p, int lrelu_on,
                     float* out, int ldo, void* saved, cudaStream_t st) {
    GcLayout L = gc_layout(B, N, F, n_out, saved, nullptr);
    const int rows = B * N;
    const bool has_bn = p->bn_g != nullptr;
    const bool direct = !has_bn;                     // no BN: the result goes straight to `out`
    float* Y = direct ? out : L.Y;
    const int ldy = direct ? ldo : n_out;
    {
        BView Am{adj, (long)N * N, N, 1};            // (m = i, k = j)
        BView Um{L.XWb, (long)N * n_out, n_out, 1};  // (k = j, n = c)
        ProfScope ps(PC_GCONV_FWD, st);
        MFT_CHECK_CUDA(launch_bgemm(Am, Um, Y, (long)N * ldy, ldy, B, N, n_out, N, 1.f, st));
    }
    {
        ProfScope ps(PC_GCONV_FWD, st);
        gconv_bias_stats_kernel<<<cdiv(rows, kGcRows), dim3(kGcCols, kGcRows), 0, st>>>(Y, ldy, p->fc_b, rows, n_out,
                                                                                      has_bn ? L.fsums : nullptr);
        MFT_CHECK_LAUNCH();
    }
    if (has_bn) {
        int total = rows * n_out;
        ProfScope ps(PC_GCONV_FWD, st);
        gconv_apply_kernel<<<min(cdiv(total, 256), 148 * 4), 256, 0, st>>>(L.Y, rows, n_out, L.fsums, p->bn_g,
                                                                           p->bn_b, lrelu_on, out, ldo);
        MFT_CHECK_LAUNCH();
    }
    return MFT_OK;
}

int gconv_fwd_check(int B, int N, int F, int n_out, int ldx, int ldo, const mft_gconv_params* p, int lrelu_on) {
    MFT_REQUIRE(B > 0 && N > 0 && F > 0 && n_out > 0, "gconv_fwd: bad shape");
    MFT_REQUIRE(n_out <= kMaxC, "gconv_fwd: n_out=%d exceeds %d", n_out, kMaxC);
    MFT_REQUIRE(ldx >= F && ldo >= n_out, "gconv_fwd: bad leading dimensions");
    MFT_REQUIRE(p->bn_g != nullptr || !lrelu_on, "gconv_fwd: LeakyReLU without BatchNorm is not a reference configuration");
    return MFT_OK;
}

int gconv_fwd(const float* adj, const float* x, int ldx, int B, int N, int F, int n_out, const mft_gconv_params* p,
              int lrelu_on, float* out, int ldo, void* saved, void* workspace, cudaStream_t st) {
    int rc = gconv_fwd_check(B, N, F, n_out, ldx, ldo, p, lrelu_on);
    if (rc != MFT_OK) return rc;
    if (gconv_fused_supported(B, N, F, n_out)) {
        // one launch (gconv_fused.cu); statistics slot and barrier counters cleared by one memset
        GcLayout L = gc_layout(B, N, F, n_out, saved, workspace);
        MFT_CHECK_CUDA(cudaMemsetAsync(L.fsums, 0, sizeof(double) * (kStatSlot + 64), st));
        return gconv_fused_fwd(adj, x, ldx, B, N, F, n_out, p, lrelu_on, out, ldo, L.Y, L.XWb, L.fsums, L.sync, st);
    }
    rc = gconv_fwd_products(x, ldx, B, N, F, n_out, p, out, ldo, saved, st);
    if (rc != MFT_OK) return rc;
    return gconv_fwd_finish(adj, B, N, F, n_out, p, lrelu_on, out, ldo, saved, st);
}

int gconv_bwd(const float* adj, const float* x, int ldx, int B, int N, int F, int n_out, const mft_gconv_params* p,
              int lrelu_on, const float* d_out, int ldo, float* dx, float* d_adj, const mft_gconv_grads* g,
              void* saved, void* workspace, cudaStream_t st, Branches* late, int late_slot) {
    MFT_REQUIRE(B > 0 && N > 0 && F > 0 && n_out > 0, "gconv_bwd: bad shape");
    MFT_REQUIRE(n_out <= kMaxC, "gconv_bwd: n_out=%d exceeds %d", n_out, kMaxC);
    GcLayout L = gc_layout(B, N, F, n_out, saved, workspace);
    const int rows = B * N;
    const int has_bn = p->bn_g != nullptr;

    // (d fc.weight is accumulated by split-K atomics: cleared on the stream its products run on)
    if (!late) MFT_CHECK_CUDA(cudaMemsetAsync(g->fc_w, 0, sizeof(float) * (size_t)n_out * 2 * F, st));
    // With Y = x Wa^T + adj (x Wb^T) + b every product of the backward runs on n_out (48 or n_way) columns or
    // over K = n_out, never on F x N:
    //   T     = adj^T dY                         [B][N, n_out]   (K = N)
    //   dx   += dY Wa + T Wb                     one launch, K = 2 n_out
    //   d_adj = dY (x Wb^T)^T                    [B][N, N]       (K = n_out; x Wb^T saved by the forward)
    //   dWa   = dY^T x,  dWb = (adj x)^T dY = T^T x
    //   main  :  dZ -> dY -> T -> dx
    //   side 1:  [after dY] d_adj ; small grads    late (or side 0 / 1): [after T] dWb ; dWa
    MFT_CHECK_CUDA(cudaMemsetAsync(L.bsums, 0, sizeof(double) * kStatSlot, st));
    dim3 blk(kGcCols, kGcRows);
    { ProfScope ps(PC_GCONV_BWD, st); gconv_dz_kernel<<<cdiv(rows, kGcRows), blk, 0, st>>>(d_out, ldo, L.Y, rows, n_out, L.fsums, p->bn_g, p->bn_b,
                                                          has_bn, lrelu_on, L.dY, L.bsums);
    MFT_CHECK_LAUNCH(); }
    if (has_bn) {
        int total = rows * n_out;
        { ProfScope ps(PC_GCONV_BWD, st); gconv_dy_kernel<<<min(cdiv(total, 256), 148 * 4), 256, 0, st>>>(L.dY, L.Y, rows, n_out, L.fsums, p->bn_g,
                                                                        L.bsums);
        MFT_CHECK_LAUNCH(); }
    }
    Branches br(st);
    cudaStream_t s1 = br.fork(1);                    // dY and the reductions exist
    PlainOp dy{L.dY, n_out};
    PlainOp qx{x, ldx};
    // the main chain is enqueued first: the replayed graph dispatches in this order, and a side product put
    // ahead of `dx` held it back until the side product had drained (measured: 12 us)
    {   // T[b,j,c] = sum_i adj[b,i,j] dY[b,i,c]
        BView At{adj, (long)N * N, 1, N};                       // (m = j, k = i) -> adj[b, i, j]
        BView Dy{L.dY, (long)N * n_out, n_out, 1};              // (k = i, n = c)
        ProfScope ps(PC_GCONV_BWD, st);
        MFT_CHECK_CUDA(launch_bgemm(At, Dy, L.T, (long)N * n_out, n_out, B, N, n_out, N, 0.f, st));
    }
    // the two weight-gradient products: nothing downstream of this call reads them
    cudaStream_t sw_a = late ? late->fork(late_slot) : br.fork(0);       // T and dY exist
    if (late) MFT_CHECK_CUDA(cudaMemsetAsync(g->fc_w, 0, sizeof(float) * (size_t)n_out * 2 * F, sw_a));
    {   // dx[r, f] += sum_c dY[r,c] fc_w[c, f] + sum_c T[r,c] fc_w[c, F + f]
        BView Dy{L.dY, 0, n_out, 1};                 // (m = row, k = c)
        BView Tt{L.T, 0, n_out, 1};
        BView Wa{p->fc_w, 0, 2 * F, 1};              // (k = c, n = f)
        BView Wb{p->fc_w + F, 0, 2 * F, 1};
        if (bgemm2_supported(n_out, n_out)) {
            ProfScope ps(PC_GCONV_BWD, st);
            MFT_CHECK_CUDA(launch_bgemm2(Dy, Wa, n_out, Tt, Wb, n_out, dx, 0, ldx, 1, rows, F, 1.f, st));
        } else {
            { ProfScope ps(PC_GCONV_BWD, st); MFT_CHECK_CUDA(launch_bgemm(Dy, Wa, dx, 0, ldx, 1, rows, F, n_out, 1.f, st)); }
            { ProfScope ps(PC_GCONV_BWD, st); MFT_CHECK_CUDA(launch_bgemm(Tt, Wb, dx, 0, ldx, 1, rows, F, n_out, 1.f, st)); }
        }
    }
    {   // d_adj[b,i,j] = sum_c dY[b,i,c] (x Wb^T)[b,j,c]
        BView Dy{L.dY, (long)N * n_out, n_out, 1};              // (m = i, k = c)
        BView Ut{L.XWb, (long)N * n_out, 1, n_out};             // (k = c, n = j) -> XWb[b, j, c]
        { ProfScope ps(PC_GCONV_BWD, s1); MFT_CHECK_CUDA(launch_bgemm(Dy, Ut, d_adj, (long)N * N, N, B, N, N, n_out, 0.f, s1)); }
    }
    { ProfScope ps(PC_GCONV_BWD, s1); gconv_small_grads_kernel<<<cdiv(n_out, 128), 128, 0, s1>>>(L.bsums, n_out, has_bn, g->fc_b, g->bn_g, g->bn_b);
    MFT_CHECK_LAUNCH(); }
    {
        PlainOp tt{L.T, n_out};
        cudaStream_t sw_b = late ? sw_a : s1;                             // without a late stream: one product per side stream
        { ProfScope ps(PC_GCONV_BWD, sw_a); MFT_CHECK_CUDA((launch_gemm_tn(tt, qx, g->fc_w + F, 2 * F, n_out, F, rows, sw_a))); }
        { ProfScope ps(PC_GCONV_BWD, sw_b); MFT_CHECK_CUDA((launch_gemm_tn(dy, qx, g->fc_w, 2 * F, n_out, F, rows, sw_b))); }
    }
    br.join(0);
    br.join(1);
    MFT_REQUIRE(br.ok(), "gconv_bwd: stream fork/join failed: %s", cudaGetErrorString(cudaGetLastError()));
    return MFT_OK;
}

}  // namespace mft
