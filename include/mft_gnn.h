/*
 * mft_gnn.h -- C ABI of the B200-native GNN few-shot head (libmft_gnn.so).
 *
 * The reference (johncai117/Meta-Fine-Tuning) has no FFI layer: its boundary for
 * this path is the Python nn.Module surface of methods/gnn.py.  The host side of
 * this repo (meta-fine-tuning_b200/gnn.py) mirrors that surface and binds the
 * entry points below through ctypes; each entry point names the reference
 * function it replaces.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller (allocated by
 *    PyTorch), contiguous float32 unless stated, and must stay alive until the
 *    stream has executed the call;
 *  - all work is enqueued on `stream` (a cudaStream_t passed as void*); no
 *    entry point synchronises the device or allocates device memory;
 *  - return value 0 = success; non-zero = failure, message via mft_last_error()
 *    (thread-local).  There is no CPU fallback: on a device that is not
 *    sm_100 the compute entry points fail with MFT_ERR_DEVICE;
 *  - `precision`: MFT_PREC_FP32 = CUDA-core fp32 path; MFT_PREC_TF32 = tcgen05
 *    kind::tf32 tensor-core path for the edge-MLP GEMMs (fp32 accumulate in
 *    TMEM, fp32 statistics / softmax / Gconv).
 *
 * Node features live in a strided matrix x[(b*N + n) * ldx + f] so that the
 * dense concatenation of GNN_nl (gnn.py:161) is a column range of one buffer.
 */
#ifndef MFT_GNN_H
#define MFT_GNN_H

#include <stddef.h>
#include <stdint.h>

/* Threading / streams.  Calls are ordered on the stream passed in and never synchronise the device.  Inside a call
 * independent kernels run on library-owned side streams (three per DEVICE, created on first use) that are forked
 * from and joined back into the caller's stream with events before the call returns.  The side streams and their
 * events are shared by every caller of a device: drive a device from ONE host thread and ONE caller stream at a
 * time (the one-process-per-GPU model of this library); two calls in flight on different caller streams of the
 * same device would re-record the same events.  mft_set_*, the grid-limit used by the backward and the profile
 * switches are process-global for the same reason. */
#ifdef __cplusplus
extern "C" {
#endif

#define MFT_PREC_FP32 0
#define MFT_PREC_TF32 1

#define MFT_OK            0
#define MFT_ERR_ARG       1
#define MFT_ERR_DEVICE    2
#define MFT_ERR_CUDA      3
#define MFT_ERR_UNSUPPORTED 4

#define MFT_MAX_LAYERS 3   /* Wcompute/Gconv pairs in one GNN_nl (num_layers 2 + last) */

/* Parameters of one Wcompute (gnn.py:60-76): conv2d_k.weight is [C_k, C_{k-1}]
 * row-major (the [out,in,1,1] tensor as stored), C = {F, 2nf, 2nf, nf, nf}.
 * conv2d_{1..4}.bias is not read: batch-statistic BatchNorm cancels it exactly. */
typedef struct {
    const float* conv_w[4];
    const float* bn_g[4];
    const float* bn_b[4];
    const float* last_w;   /* conv2d_last.weight [nf] */
    const float* last_b;   /* conv2d_last.bias   [1]  */
} mft_wcompute_params;

/* Gradient destinations for one Wcompute; the library overwrites them. */
typedef struct {
    float* conv_w[4];
    float* conv_b[4];      /* analytically zero -> written as 0 */
    float* bn_g[4];
    float* bn_b[4];
    float* last_w;
    float* last_b;         /* analytically zero -> written as 0 */
} mft_wcompute_grads;

/* Parameters of one Gconv (gnn.py:32-41): fc.weight [n_out, 2F], fc.bias [n_out],
 * bn.weight/bias [n_out] or NULL when bn_bool is False. */
typedef struct {
    const float* fc_w;
    const float* fc_b;
    const float* bn_g;
    const float* bn_b;
} mft_gconv_params;

typedef struct {
    float* fc_w;
    float* fc_b;
    float* bn_g;
    float* bn_b;
} mft_gconv_grads;

/* Whole GNN_nl (gnn.py:134-152): layer_w0, layer_l0, layer_w1, layer_l1,
 * w_comp_last, layer_last. */
typedef struct {
    mft_wcompute_params w[MFT_MAX_LAYERS];
    mft_gconv_params    l[MFT_MAX_LAYERS];
} mft_gnn_params;

typedef struct {
    mft_wcompute_grads w[MFT_MAX_LAYERS];
    mft_gconv_grads    l[MFT_MAX_LAYERS];
} mft_gnn_grads;

/* ---- library / device -------------------------------------------------- */

/* Message of the last failure on this thread ("" if none). */
const char* mft_last_error(void);
/* ABI version of this header. */
int mft_version(void);
/* 0 when `device` is an sm_100 part this library can run on. */
int mft_device_check(int device);
/* 1 when the tcgen05 TF32 path supports this Wcompute shape, else 0. */
int mft_tf32_supported(int F, int nf);

/* ---- Wcompute: replaces Wcompute.forward (gnn.py:78-132) ------------------ */

/* Bytes of the activation tape kept from forward to backward / of scratch. */
size_t mft_wcompute_saved_bytes(int B, int N, int F, int nf);
/* The same for a known precision (MFT_PREC_TF32 keeps an fp16 activation tape: half the bytes). */
size_t mft_wcompute_saved_bytes_for(int B, int N, int F, int nf, int precision);
size_t mft_wcompute_workspace_bytes(int B, int N, int F, int nf);

/* x [B*N, ldx] (first F columns used) -> adj [B,N,N]: adj[b,i,:] =
 * softmax_j(edge_mlp(|x_i - x_j|) - 1e8*[i==j]).  The reference's output
 * [B,N,N,2] is stack(identity, adj); the identity half is built by the host. */
int mft_wcompute_fwd(const float* x, int ldx, int B, int N, int F, int nf,
                     const mft_wcompute_params* p, float* adj,
                     void* saved, void* workspace, int precision,
                     const unsigned char* shared_nodes, void* stream);

/* Backward: d_adj [B,N,N] -> dx (ACCUMULATED into dx[(b*N+n)*ldx + f], f < F)
 * and parameter gradients (overwritten). */
int mft_wcompute_bwd(const float* x, int ldx, int B, int N, int F, int nf,
                     const mft_wcompute_params* p, const float* adj, const float* d_adj,
                     float* dx, const mft_wcompute_grads* g,
                     void* saved, void* workspace, int precision,
                     const unsigned char* shared_nodes, void* stream);

/* ---- Gconv: replaces gmul + Gconv.forward (gnn.py:16-28, 43-56) ------------- */

size_t mft_gconv_saved_bytes(int B, int N, int F, int n_out);
size_t mft_gconv_workspace_bytes(int B, int N, int F, int n_out);

/* out[(b*N+n)*ldo + c] = act(BN1d(x Wa^T + adj (x Wb^T) + b)), c < n_out.
 * bn: p->bn_g != NULL; lrelu != 0 applies the LeakyReLU of gnn.py:160. */
int mft_gconv_fwd(const float* adj, const float* x, int ldx, int B, int N, int F, int n_out,
                  const mft_gconv_params* p, int lrelu, float* out, int ldo,
                  void* saved, void* workspace, void* stream);

/* d_out [B*N, ldo] -> dx (ACCUMULATED, ldx), d_adj [B,N,N] (overwritten),
 * parameter gradients (overwritten). */
int mft_gconv_bwd(const float* adj, const float* x, int ldx, int B, int N, int F, int n_out,
                  const mft_gconv_params* p, int lrelu, const float* d_out, int ldo,
                  float* dx, float* d_adj, const mft_gconv_grads* g,
                  void* saved, void* workspace, void* stream);

/* ---- GNN_nl: replaces GNN_nl.forward (gnn.py:154-166) in one call ---------- */

size_t mft_gnn_saved_bytes(int B, int N, int F0, int nf, int n_way);
size_t mft_gnn_saved_bytes_for(int B, int N, int F0, int nf, int n_way, int precision);
size_t mft_gnn_workspace_bytes(int B, int N, int F0, int nf, int n_way);

/* x [B,N,F0] contiguous -> out [B,N,n_way] contiguous. */
int mft_gnn_fwd(const float* x, int B, int N, int F0, int nf, int n_way,
                const mft_gnn_params* p, float* out,
                void* saved, void* workspace, int precision,
                const unsigned char* shared_nodes, void* stream);

/* d_out [B,N,n_way] -> dx [B,N,F0] (overwritten) + every parameter gradient. */
int mft_gnn_bwd(const float* d_out, int B, int N, int F0, int nf, int n_way,
                const mft_gnn_params* p, float* dx, const mft_gnn_grads* g,
                void* saved, void* workspace, int precision,
                const unsigned char* shared_nodes, void* stream);

/* ---- GnnNet pre-head: fc + graph assembly (gnnnet.py:30, 71-83, 35-38, 212) ----------------
 * feat [n_way, n_support+n_query, feat_dim] contiguous (backbone features of one episode)
 *   z = BatchNorm1d(Linear(feat))            (fc->fc_w [D, feat_dim], fc_b, bn_g, bn_b [D]; batch statistics)
 *   nodes [n_query, n_way*(n_support+1), D+n_way]: graph q = per class its n_support support rows
 *   then its q-th query row; columns D.. hold the one-hot class of a support node, zeros for a query.
 * The backward sums the node gradients of a support row over the n_query graphs, runs BatchNorm1d
 * backward and returns the fc / bn gradients (overwritten) and, if d_feat != NULL, d_feat. */
size_t mft_head_saved_bytes(int n_way, int n_support, int n_query, int D);
size_t mft_head_workspace_bytes(int n_way, int n_support, int n_query, int D);
int mft_head_fwd(const float* feat, int feat_dim, int n_way, int n_support, int n_query, int D,
                 const mft_gconv_params* fc, float* nodes, void* saved, void* workspace, void* stream);
int mft_head_bwd(const float* feat, int feat_dim, int n_way, int n_support, int n_query, int D,
                 const mft_gconv_params* fc, const float* d_nodes, float* d_feat,
                 const mft_gconv_grads* g, void* saved, void* workspace, void* stream);

/* Cross-entropy of the query nodes and its gradient in one launch (GnnNet.forward_gnn's score selection,
 * gnnnet.py:216, + set_forward_loss, gnnnet.py:219-224).  out [n_query, n_way*(n_support+1), n_way] is
 * GNN_nl's output for the graphs of one episode; the query of class c sits at node c*(n_support+1)+n_support
 * and has label c.  *loss = mean over the n_way*n_query query nodes; d_out (same shape as out) = d loss / d out,
 * exact zeros on the support nodes. */
int mft_query_ce(const float* out, int n_way, int n_support, int n_query, float* loss, float* d_out,
                 void* stream);

/* ---- measurement hooks (new; the reference has no profiler) ---- */

/* Kernels launched by this library in this process so far (bench.py: gpu_launches). */
unsigned long long mft_launch_count(void);
/* Bracket every kernel launch with CUDA events on its stream (on != 0) or stop (0). */
int mft_prof_enable(int on);
/* Launch overlap (new): 0 = plain stream order; 1 = the consecutive tcgen05 GEMM launches of an
 * edge MLP use programmatic dependent launch (prologue of launch n+1 -- barrier/TMEM set-up and the
 * resident weight image -- runs under the tail of launch n); 2 = also the row kernels around them
 * (default; environment MFT_PDL).  Worth about 1 % at 5-way 20-shot and 4 % at 5-way 5-shot on B200: every
 * layer is a grid-wide BatchNorm dependency, so only launch latency and set-up overlap.  Results are
 * identical at every level.  Returns the old level. */
int mft_set_pdl(int level);
int mft_prof_categories(void);
const char* mft_prof_name(int cat);
/* Synchronise, then ms[c] / counts[c] = summed device time and launches of category c
 * since the last collect/enable; n = capacity of both arrays. */
int mft_prof_collect(float* ms, int* counts, int n);

/* ---- test entry: one tcgen05 "rows x weights" GEMM with plain operands ------------ */

/* C[M,N] = A[M,K] * op(W)^T on the tensor-core path; W is [N,K] (transpose_w = 0) or
 * [K,N] (transpose_w = 1).  K <= 256, any N (split into passes).  Used by the tests to
 * pin descriptors / swizzle / pipeline independently of the edge-MLP fusion. */
size_t mft_debug_umma_gemm_workspace_bytes(int N, int K);
int mft_debug_umma_gemm(const float* A, int lda, const float* W, int ldw, int transpose_w,
                        float* C, int ldc, int M, int N, int K, void* workspace, void* stream);
/* dW[Cout,Cin] += P[R,Cout]^T * Q[R,Cin] on the tensor-core wgrad kernel (Cout <= 192, Cin <= 256). */
/* Tests: byte offsets of the activation tape inside a Wcompute `saved` blob (H_1..H_4, forward statistics,
 * tape scales; out[6], out[7] = doubles per statistics slot / per copy).  out: size_t[8]. */
int mft_debug_wcompute_saved_offsets(int B, int N, int F, int nf, int precision, size_t* out);

/* Per-CTA clock64 timeline of ONE rows-GEMM launch, the (skip+1)-th from now, into buf [grid][16]. */
int mft_debug_set_timeline(void* buf, int skip);
int mft_debug_umma_wgrad(const float* P, int ldp, const float* Q, int ldq, float* dW, int ldw,
                         int R, int Cout, int Cin, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MFT_GNN_H */
