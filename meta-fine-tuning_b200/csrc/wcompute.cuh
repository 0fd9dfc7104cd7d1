// Internal interfaces between the translation units of libmft_gnn.
#pragma once

#include "common.cuh"

namespace mft {

// Carved views of the Wcompute `saved` / `workspace` blobs.
// The tcgen05 wgrad CTAs each store their partial dW into a private copy (plain stores; atomics from
// 148 CTAs onto one 147 KB matrix cost more than the GEMM); finalize_grads_kernel sums the copies.
constexpr int kWgMaxCopies = 160;   // >= number of SMs

struct WcLayout {
    int C[5];            // channel widths: F, 2nf, 2nf, nf, nf
    float* H[4];         // saved: pre-BN activations of the four conv layers, [R, C[k+1]]
    double* fsums;       // saved: forward batch statistics, 4 x [2*kMaxC]
    float* tscale;       // saved (tcgen05 path): power-of-two tape scale s_k of the four conv layers (umma_layers.cu)
    int* tri;            // workspace: unordered-pair table [Rg]
    int* inv;            // workspace: (i, j) -> table index [N*N]
    float* roww;         // workspace: per-row multiplicity [R + 1] (PairGeom::roww)
    int2* rowij;         // workspace: per-row node-matrix rows [R + 1] (PairGeom::rowij)
    float* S;            // workspace: scores / dS, [B,N,N]
    float* dyA;          // workspace: ping-pong gradient buffers [R, 2nf]
    float* dyB;
    float* dyC;          // third buffer: lets wgrad_k (side stream) read dy_k while dgrad_{k-1} writes dy_{k-2}
    double* bsums;       // workspace: backward reductions, 5 x [2*kMaxC] (last = d conv2d_last.weight)
    float* wimg;         // workspace: swizzled TF32 weight image of the tcgen05 path
    float* dD;           // workspace (tcgen05 path): dL/d|x_i-x_j| per unordered pair, [R, roundup4(F)]
    float* wgpart;       // workspace (tcgen05 path): per-CTA partial copies of the four conv-weight gradients
    size_t wgpart_off[4];  // float offset of layer k's copies inside wgpart (copy stride = C[k+1]*roundup4(C[k]))
    size_t wgpart_floats;
    size_t saved_bytes, workspace_bytes;
};

size_t umma_workspace_floats(int F, int nf);
// Upper bound on the CTAs of the next tcgen05 GEMM launches (0 = all SMs): wcompute_bwd runs the wgrad of a
// layer beside its dgrad, each on half of the SMs (persistent kernels loop over their tiles whatever the grid).
void umma_set_grid_limit(int ctas);
int umma_num_sms();
bool umma_shape_supported(int F, int nf);
int umma_debug_gemm(const float* A, int lda, const float* W, int ldw, int transpose_w, float* C, int ldc, int M,
                    int N, int K, float* wimg, cudaStream_t st);
size_t umma_wimg_floats(int N, int K);
int umma_debug_wgrad(const float* P, int ldp, const float* Q, int ldq, float* dW, int ldw, int R, int Cout,
                     int Cin, cudaStream_t st);

// precision: MFT_PREC_TF32 sizes the activation tape H_1..H_4 for fp16 elements, anything else for fp32
WcLayout wc_layout(int B, int N, int F, int nf, void* saved, void* workspace, int precision = MFT_PREC_FP32);

// wcompute_fwd_prepare: everything of the forward that depends on the parameters and the shape only
// (statistics slots zeroed, pair tables, the four tcgen05 weight images) -- gnn_fwd issues it for layer
// l+1 on a side branch while the Gconv of layer l runs; wcompute_fwd(prepared = true) then skips it.
int wcompute_fwd_prepare(int B, int N, int F, int nf, const mft_wcompute_params* p, void* saved, void* workspace,
                         int precision, const unsigned char* shared_nodes, cudaStream_t st);
int wcompute_fwd(const float* x, int ldx, int B, int N, int F, int nf, const mft_wcompute_params* p, float* adj,
                 void* saved, void* workspace, int precision, const unsigned char* shared_nodes, cudaStream_t st,
                 bool prepared = false, class Branches* mid = nullptr, int mid_slot = 0);
// mid: side stream `mid_slot` of *mid is forked once the four conv layers are enqueued, i.e. beside the score and
// softmax kernels, which leave most of the chip idle (gnn_fwd puts the next Gconv's x W products there).
// tail_st: stream for the kernels that only finish parameter gradients (finalize_grads, wgrad_reduce);
// when it differs from st the caller has ordered it after st's work so far and joins it before the
// workspace is reused (gnn_bwd runs them under the next layer's Gconv backward).
int wcompute_bwd(const float* x, int ldx, int B, int N, int F, int nf, const mft_wcompute_params* p,
                 const float* adj, const float* d_adj, float* dx, const mft_wcompute_grads* gr, void* saved,
                 void* workspace, int precision, const unsigned char* shared_nodes, cudaStream_t st,
                 class Branches* tail = nullptr, int tail_slot = 0, bool prepared = false);
// prepared: wcompute_bwd_prepare has run (and been joined) for this layer since the workspace's tables / images
// were last overwritten.
int wcompute_bwd_prepare(int B, int N, int F, int nf, const mft_wcompute_params* p, void* saved, void* workspace,
                         int precision, const unsigned char* shared_nodes, cudaStream_t st);

// tcgen05 (MFT_PREC_TF32) replacements for the four forward layer GEMMs and for the
// dgrad + wgrad pair of one backward layer; same buffers in and out as the fp32 path.
int wcompute_fwd_layers_tf32(const float* x, int ldx, int F, int nf, const mft_wcompute_params* p,
                             const WcLayout& L, const PairGeom& g, cudaStream_t st);
int wcompute_fwd_prepare_tf32(const mft_wcompute_params* p, const WcLayout& L, int F, int nf, double count,
                              cudaStream_t st);
int wcompute_bwd_layer_tf32(int k, float* dh, float* dy_next, const float* x, int ldx, float* dx, int F, int nf,
                            const mft_wcompute_params* p, const mft_wcompute_grads* gr, const WcLayout& L,
                            const PairGeom& g, cudaStream_t st);

int wcompute_bwd_prepare_tf32(const mft_wcompute_params* p, const WcLayout& L, int F, int nf, cudaStream_t st);
int wcompute_wgrad_layer_tf32(int k, const float* dh, const float* x, int ldx, int F,
                              const mft_wcompute_params* p, const mft_wcompute_grads* gr, const WcLayout& L,
                              const PairGeom& g, int* copies, cudaStream_t st);

struct GcLayout {
    float* Y;            // saved: pre-BN Gconv output [B*N, n_out]
    double* fsums;       // saved: BN1d statistics [2*kMaxC]
    int* sync;           // saved (right behind fsums: one memset clears both): barrier counters of the fused kernels
    float* XWb;          // saved: x Wb^T [B*N, n_out], the operand of the adjacency product (d_adj = dY (x Wb^T)^T)
    float* dY;           // workspace bwd: [B*N, n_out]
    float* T;            // workspace bwd: adj^T dY  [B*N, n_out]
    double* bsums;       // workspace bwd: [2*kMaxC] + fc bias sums [kMaxC]
    size_t saved_bytes, workspace_bytes;
};

GcLayout gc_layout(int B, int N, int F, int n_out, void* saved, void* workspace);

int gconv_fwd(const float* adj, const float* x, int ldx, int B, int N, int F, int n_out, const mft_gconv_params* p,
              int lrelu, float* out, int ldo, void* saved, void* workspace, cudaStream_t st);
// the two halves of gconv_fwd (not for the one-launch variant): what needs no adjacency, and the rest
int gconv_fwd_check(int B, int N, int F, int n_out, int ldx, int ldo, const mft_gconv_params* p, int lrelu);
int gconv_fwd_products(const float* x, int ldx, int B, int N, int F, int n_out, const mft_gconv_params* p,
                       float* out, int ldo, void* saved, cudaStream_t st);
int gconv_fwd_finish(const float* adj, int B, int N, int F, int n_out, const mft_gconv_params* p, int lrelu,
                     float* out, int ldo, void* saved, cudaStream_t st);
// late / late_slot: side stream of the caller's Branches for the two fc.weight gradient products, which nothing
// downstream of the call needs; the caller joins it before the Gconv workspace is reused (null: joined inside).
int gconv_bwd(const float* adj, const float* x, int ldx, int B, int N, int F, int n_out, const mft_gconv_params* p,
              int lrelu, const float* d_out, int ldo, float* dx, float* d_adj, const mft_gconv_grads* g,
              void* saved, void* workspace, cudaStream_t st, class Branches* late = nullptr, int late_slot = 0);

// gconv_fused.cu: the whole Gconv forward in one launch (spin barriers between its phases)
int gconv_fused_supported(int B, int N, int F, int n_out);
int gconv_fused_fwd(const float* adj, const float* x, int ldx, int B, int N, int F, int n_out, const mft_gconv_params* p,
                    int lrelu_on, float* out, int ldo, float* Y, float* V, double* fsums, int* sync, cudaStream_t st);

int query_ce(const float* out, int n_way, int n_support, int n_query, float* loss, float* d_out, cudaStream_t st);

size_t head_saved_bytes(int rows, int D);
size_t head_workspace_bytes(int rows, int D);
int head_fwd(const float* feat, int feat_dim, int n_way, int n_support, int n_query, int D,
             const mft_gconv_params* fc, float* nodes, void* saved, void* workspace, cudaStream_t st);
int head_bwd(const float* feat, int feat_dim, int n_way, int n_support, int n_query, int D,
             const mft_gconv_params* fc, const float* d_nodes, float* d_feat, const mft_gconv_grads* g, void* saved,
             void* workspace, cudaStream_t st);

}  // namespace mft
