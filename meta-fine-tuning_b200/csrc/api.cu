// extern "C" surface of libmft_gnn.so (declared in include/mft_gnn.h).
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "common.cuh"
#include "wcompute.cuh"

namespace mft {

static thread_local char g_err[512] = "";
static thread_local int g_code = 0;

void set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    g_code = code;
}
int last_code() { return g_code; }

size_t gnn_saved_bytes(int B, int N, int F0, int nf, int n_way, int precision);
size_t gnn_workspace_bytes(int B, int N, int F0, int nf, int n_way);
int gnn_fwd(const float* x, int B, int N, int F0, int nf, int n_way, const mft_gnn_params* p, float* out,
            void* saved, void* workspace, int precision, const unsigned char* shared_nodes, cudaStream_t st);
int gnn_bwd(const float* d_out, int B, int N, int F0, int nf, int n_way, const mft_gnn_params* p, float* dx,
            const mft_gnn_grads* g, void* saved, void* workspace, int precision, const unsigned char* shared_nodes,
            cudaStream_t st);

// The kernels are compiled for sm_100a only; refuse anything else up front instead of
// failing with "no kernel image" somewhere in the middle of a call.
static int require_sm100() {
    static int cached[64] = {0};   // 0 unknown, 1 ok, 2 bad
    int dev = 0;
    MFT_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return MFT_OK;
    if (cached[dev] == 0) {
        int major = 0;
        MFT_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
        cached[dev] = (major == 10) ? 1 : 2;
    }
    if (cached[dev] != 1) {
        set_error(MFT_ERR_DEVICE, "device %d is not an sm_100 (Blackwell B200) part; libmft_gnn has no other path", dev);
        return MFT_ERR_DEVICE;
    }
    return MFT_OK;
}

}  // namespace mft

using namespace mft;
namespace mft { extern long long* g_umma_dbg; extern int g_umma_dbg_skip; }

#define MFT_ENTER()                         \
    do {                                    \
        int _rc = require_sm100();          \
        if (_rc != MFT_OK) return _rc;      \
    } while (0)

extern "C" {

const char* mft_last_error(void) { return g_err; }
int mft_version(void) { return 1; }

int mft_device_check(int device) {
    int major = 0;
    cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    if (e != cudaSuccess) {
        set_error(MFT_ERR_CUDA, "cudaDeviceGetAttribute failed: %s", cudaGetErrorString(e));
        return MFT_ERR_CUDA;
    }
    if (major != 10) {
        set_error(MFT_ERR_DEVICE, "device %d has compute capability major %d, need 10 (sm_100)", device, major);
        return MFT_ERR_DEVICE;
    }
    return MFT_OK;
}

int mft_tf32_supported(int F, int nf) { return umma_shape_supported(F, nf) ? 1 : 0; }

size_t mft_wcompute_saved_bytes(int B, int N, int F, int nf) {
    return wc_layout(B, N, F, nf, nullptr, nullptr).saved_bytes;      // fp32 tape: enough for either precision
}
size_t mft_wcompute_saved_bytes_for(int B, int N, int F, int nf, int precision) {
    return wc_layout(B, N, F, nf, nullptr, nullptr, precision).saved_bytes;
}
size_t mft_wcompute_workspace_bytes(int B, int N, int F, int nf) {
    return wc_layout(B, N, F, nf, nullptr, nullptr).workspace_bytes;
}

int mft_wcompute_fwd(const float* x, int ldx, int B, int N, int F, int nf, const mft_wcompute_params* p,
                     float* adj, void* saved, void* workspace, int precision, const unsigned char* shared_nodes,
                     void* stream) {
    MFT_ENTER();
    MFT_REQUIRE(x && p && adj && saved && workspace, "mft_wcompute_fwd: null pointer");
    return wcompute_fwd(x, ldx, B, N, F, nf, p, adj, saved, workspace, precision, shared_nodes,
                        (cudaStream_t)stream);
}

int mft_wcompute_bwd(const float* x, int ldx, int B, int N, int F, int nf, const mft_wcompute_params* p,
                     const float* adj, const float* d_adj, float* dx, const mft_wcompute_grads* g, void* saved,
                     void* workspace, int precision, const unsigned char* shared_nodes, void* stream) {
    MFT_ENTER();
    MFT_REQUIRE(x && p && adj && d_adj && dx && g && saved && workspace, "mft_wcompute_bwd: null pointer");
    return wcompute_bwd(x, ldx, B, N, F, nf, p, adj, d_adj, dx, g, saved, workspace, precision, shared_nodes,
                        (cudaStream_t)stream);
}

size_t mft_gconv_saved_bytes(int B, int N, int F, int n_out) {
    return gc_layout(B, N, F, n_out, nullptr, nullptr).saved_bytes;
}
size_t mft_gconv_workspace_bytes(int B, int N, int F, int n_out) {
    return gc_layout(B, N, F, n_out, nullptr, nullptr).workspace_bytes;
}

int mft_gconv_fwd(const float* adj, const float* x, int ldx, int B, int N, int F, int n_out,
                  const mft_gconv_params* p, int lrelu, float* out, int ldo, void* saved, void* workspace,
                  void* stream) {
    MFT_ENTER();
    MFT_REQUIRE(adj && x && p && out && saved && workspace, "mft_gconv_fwd: null pointer");
    return gconv_fwd(adj, x, ldx, B, N, F, n_out, p, lrelu, out, ldo, saved, workspace, (cudaStream_t)stream);
}

int mft_gconv_bwd(const float* adj, const float* x, int ldx, int B, int N, int F, int n_out,
                  const mft_gconv_params* p, int lrelu, const float* d_out, int ldo, float* dx, float* d_adj,
                  const mft_gconv_grads* g, void* saved, void* workspace, void* stream) {
    MFT_ENTER();
    MFT_REQUIRE(adj && x && p && d_out && dx && d_adj && g && saved && workspace, "mft_gconv_bwd: null pointer");
    return gconv_bwd(adj, x, ldx, B, N, F, n_out, p, lrelu, d_out, ldo, dx, d_adj, g, saved, workspace,
                     (cudaStream_t)stream);
}

size_t mft_gnn_saved_bytes(int B, int N, int F0, int nf, int n_way) { return gnn_saved_bytes(B, N, F0, nf, n_way, MFT_PREC_FP32); }
size_t mft_gnn_saved_bytes_for(int B, int N, int F0, int nf, int n_way, int precision) {
    return gnn_saved_bytes(B, N, F0, nf, n_way, precision);
}
size_t mft_gnn_workspace_bytes(int B, int N, int F0, int nf, int n_way) {
    return gnn_workspace_bytes(B, N, F0, nf, n_way);
}

int mft_gnn_fwd(const float* x, int B, int N, int F0, int nf, int n_way, const mft_gnn_params* p, float* out,
                void* saved, void* workspace, int precision, const unsigned char* shared_nodes, void* stream) {
    MFT_ENTER();
    MFT_REQUIRE(x && p && out && saved && workspace, "mft_gnn_fwd: null pointer");
    return gnn_fwd(x, B, N, F0, nf, n_way, p, out, saved, workspace, precision, shared_nodes,
                   (cudaStream_t)stream);
}

int mft_gnn_bwd(const float* d_out, int B, int N, int F0, int nf, int n_way, const mft_gnn_params* p, float* dx,
                const mft_gnn_grads* g, void* saved, void* workspace, int precision,
                const unsigned char* shared_nodes, void* stream) {
    MFT_ENTER();
    MFT_REQUIRE(d_out && p && dx && g && saved && workspace, "mft_gnn_bwd: null pointer");
    return gnn_bwd(d_out, B, N, F0, nf, n_way, p, dx, g, saved, workspace, precision, shared_nodes,
                   (cudaStream_t)stream);
}

int mft_query_ce(const float* out, int n_way, int n_support, int n_query, float* loss, float* d_out, void* stream) {
    MFT_ENTER();
    MFT_REQUIRE(out && loss && d_out, "mft_query_ce: null pointer");
    return query_ce(out, n_way, n_support, n_query, loss, d_out, (cudaStream_t)stream);
}

size_t mft_head_saved_bytes(int n_way, int n_support, int n_query, int D) {
    return head_saved_bytes(n_way * (n_support + n_query), D);
}
size_t mft_head_workspace_bytes(int n_way, int n_support, int n_query, int D) {
    return head_workspace_bytes(n_way * (n_support + n_query), D);
}

int mft_head_fwd(const float* feat, int feat_dim, int n_way, int n_support, int n_query, int D,
                 const mft_gconv_params* fc, float* nodes, void* saved, void* workspace, void* stream) {
    MFT_ENTER();
    MFT_REQUIRE(feat && fc && nodes && saved && workspace, "mft_head_fwd: null pointer");
    return head_fwd(feat, feat_dim, n_way, n_support, n_query, D, fc, nodes, saved, workspace, (cudaStream_t)stream);
}

int mft_head_bwd(const float* feat, int feat_dim, int n_way, int n_support, int n_query, int D,
                 const mft_gconv_params* fc, const float* d_nodes, float* d_feat, const mft_gconv_grads* g,
                 void* saved, void* workspace, void* stream) {
    MFT_ENTER();
    MFT_REQUIRE(feat && fc && d_nodes && g && saved && workspace, "mft_head_bwd: null pointer");
    return head_bwd(feat, feat_dim, n_way, n_support, n_query, D, fc, d_nodes, d_feat, g, saved, workspace,
                    (cudaStream_t)stream);
}

size_t mft_debug_umma_gemm_workspace_bytes(int N, int K) { return (umma_wimg_floats(N, K) + 64) * sizeof(float); }

int mft_debug_umma_gemm(const float* A, int lda, const float* W, int ldw, int transpose_w, float* C, int ldc,
                        int M, int N, int K, void* workspace, void* stream) {
    MFT_ENTER();
    MFT_REQUIRE(A && W && C && workspace, "mft_debug_umma_gemm: null pointer");
    MFT_REQUIRE(umma_wimg_floats(N, K) > 0, "mft_debug_umma_gemm: unsupported N=%d K=%d", N, K);
    return umma_debug_gemm(A, lda, W, ldw, transpose_w, C, ldc, M, N, K, static_cast<float*>(workspace),
                           (cudaStream_t)stream);
}

int mft_debug_umma_wgrad(const float* P, int ldp, const float* Q, int ldq, float* dW, int ldw, int R, int Cout,
                         int Cin, void* stream) {
    MFT_ENTER();
    MFT_REQUIRE(P && Q && dW, "mft_debug_umma_wgrad: null pointer");
    return umma_debug_wgrad(P, ldp, Q, ldq, dW, ldw, R, Cout, Cin, (cudaStream_t)stream);
}

/* Debug / tests: where the activation tape lives inside a Wcompute `saved` blob.  out[0..3] = byte offsets of the
 * pre-BatchNorm activations H_1..H_4 ([R, C_k] row-major, R pair rows in the order of the row table: fp16 on the
 * tensor-core path, fp32 on the fp32 path), out[4] = forward statistics (4 slots of kStatCopies x [2][kMaxC]
 * doubles), out[5] = the four tape scales (floats), out[6] = doubles per statistics slot, out[7] = doubles per copy. */
int mft_debug_wcompute_saved_offsets(int B, int N, int F, int nf, int precision, size_t* out) {
    MFT_REQUIRE(out && B > 0 && N > 0 && F > 0 && nf > 0, "mft_debug_wcompute_saved_offsets: bad argument");
    char* base = nullptr;
    WcLayout L = wc_layout(B, N, F, nf, base, base, precision);
    for (int k = 0; k < 4; ++k) out[k] = (size_t)(reinterpret_cast<char*>(L.H[k]) - base);
    out[4] = (size_t)(reinterpret_cast<char*>(L.fsums) - base);
    out[5] = (size_t)(reinterpret_cast<char*>(L.tscale) - base);
    out[6] = (size_t)kStatSlot;
    out[7] = (size_t)kStatCopyStride;
    return MFT_OK;
}

/* Debug: have every following tcgen05 rows-GEMM launch write a per-CTA clock64 timeline
 * ([grid][16] long long) into `buf` (device memory) for ONE launch: the (skip+1)-th from now. */
int mft_debug_set_timeline(void* buf, int skip) {
    g_umma_dbg = static_cast<long long*>(buf);
    g_umma_dbg_skip = skip;
    return MFT_OK;
}

}  // extern "C"
