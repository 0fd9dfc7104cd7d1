// GNN_nl forward / backward in one library call each.
// Replaces GNN_nl.forward (methods/gnn.py:154-166): two (Wcompute, Gconv, LeakyReLU,
// concat) blocks and the final (Wcompute, Gconv).  The dense concatenation is a
// column range of one wide node buffer `xcat` [B*N, LDX]; layer l reads columns
// [0, F0 + l*nf/2) and writes its 48 new features right behind them, so no copy
// is ever made.  The backward mirrors it with `dxcat`.
#include "common.cuh"
#include "wcompute.cuh"
#include "prof.cuh"

namespace mft {

struct GnnLayout {
    int L;                 // Wcompute/Gconv pairs (num_layers + 1)
    int F[MFT_MAX_LAYERS]; // input width of pair l
    int nout[MFT_MAX_LAYERS];
    int ldx;
    float* xcat;           // saved
    float* adj[MFT_MAX_LAYERS];
    void* wc_saved[MFT_MAX_LAYERS];
    void* gc_saved[MFT_MAX_LAYERS];
    float* dxcat;          // workspace
    float* d_adj;
    void* wc_ws;           // shared by every Wcompute call (stream-ordered)
    void* gc_ws;           // shared by every Gconv call; separate from wc_ws because a Gconv runs beside the
                           // preparation (forward) / gradient finalisation (backward) of a neighbouring Wcompute
    size_t saved_bytes, workspace_bytes;
};

static GnnLayout gnn_layout(int B, int N, int F0, int nf, int n_way, void* saved, void* workspace,
                            int precision = MFT_PREC_FP32) {
    GnnLayout G;
    G.L = MFT_MAX_LAYERS;
    const int half = nf / 2;
    for (int l = 0; l < G.L; ++l) {
        G.F[l] = F0 + half * l;
        G.nout[l] = (l == G.L - 1) ? n_way : half;
    }
    int ftot = G.F[G.L - 1];
    G.ldx = (ftot + 3) & ~3;
    size_t rows = (size_t)B * N;
    Carver sv(saved);
    G.xcat = sv.take<float>(rows * G.ldx);
    size_t wc_ws = 0, gc_ws = 0;
    for (int l = 0; l < G.L; ++l) {
        G.adj[l] = sv.take<float>((size_t)B * N * N);
        WcLayout w = wc_layout(B, N, G.F[l], nf, nullptr, nullptr, precision);
        GcLayout c = gc_layout(B, N, G.F[l], G.nout[l], nullptr, nullptr);
        G.wc_saved[l] = sv.take<char>(w.saved_bytes);
        G.gc_saved[l] = sv.take<char>(c.saved_bytes);
        wc_ws = wc_ws > w.workspace_bytes ? wc_ws : w.workspace_bytes;
        gc_ws = gc_ws > c.workspace_bytes ? gc_ws : c.workspace_bytes;
    }
    G.saved_bytes = sv.used();
    Carver ws(workspace);
    G.dxcat = ws.take<float>(rows * G.ldx);
    G.d_adj = ws.take<float>((size_t)B * N * N);
    G.wc_ws = ws.take<char>(wc_ws);
    G.gc_ws = ws.take<char>(gc_ws);
    G.workspace_bytes = ws.used();
    return G;
}

size_t gnn_saved_bytes(int B, int N, int F0, int nf, int n_way, int precision) {
    return gnn_layout(B, N, F0, nf, n_way, nullptr, nullptr, precision).saved_bytes;
}
size_t gnn_workspace_bytes(int B, int N, int F0, int nf, int n_way) {
    return gnn_layout(B, N, F0, nf, n_way, nullptr, nullptr).workspace_bytes;
}

static bool gconv_hoist() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MFT_GCONV_HOIST");
        v = e ? (atoi(e) != 0) : 1;
    }
    return v != 0;
}

int gnn_fwd(const float* x, int B, int N, int F0, int nf, int n_way, const mft_gnn_params* p, float* out,
            void* saved, void* workspace, int precision, const unsigned char* shared_nodes, cudaStream_t st) {
    MFT_REQUIRE(B > 0 && N > 0 && F0 > 0 && nf >= 2 && (nf % 2) == 0 && n_way > 0, "gnn_fwd: bad shape");
    GnnLayout G = gnn_layout(B, N, F0, nf, n_way, saved, workspace, precision);
    const int rows = B * N;
    // (the node buffer's not-yet-written columns are read -- and masked -- by the 16-byte loads of the |x_i - x_j|
    // producers at the K tail: give them defined values)
    MFT_CHECK_CUDA(cudaMemsetAsync(G.xcat, 0, sizeof(float) * (size_t)rows * G.ldx, st));
    MFT_CHECK_CUDA(cudaMemcpy2DAsync(G.xcat, sizeof(float) * G.ldx, x, sizeof(float) * F0, sizeof(float) * F0,
                                     rows, cudaMemcpyDeviceToDevice, st));
    // later layers see per-graph features: only layer 0 may share support pairs
    auto prepare = [&](int l, cudaStream_t s) {
        return wcompute_fwd_prepare(B, N, G.F[l], nf, &p->w[l], G.wc_saved[l], G.wc_ws, precision,
                                    l == 0 ? shared_nodes : nullptr, s);
    };
    int rc = prepare(0, st);
    if (rc != MFT_OK) return rc;
    Branches br(st);
    for (int l = 0; l < G.L; ++l) {
        const bool last = (l == G.L - 1);
        float* dst = last ? out : G.xcat + G.F[l];
        int ldo = last ? n_way : G.ldx;
        // The Gconv's two products with x need no adjacency: they run on side stream 3, forked inside wcompute_fwd
        // behind the fourth conv layer, i.e. beside the score / softmax kernels (not for the one-launch Gconv).
        const bool hoist = gconv_hoist() && !gconv_fused_supported(B, N, G.F[l], G.nout[l]);
        {
            ProfScope span(PC_SPAN_WC_FWD, st, false);
            rc = wcompute_fwd(G.xcat, G.ldx, B, N, G.F[l], nf, &p->w[l], G.adj[l], G.wc_saved[l], G.wc_ws,
                              precision, l == 0 ? shared_nodes : nullptr, st, true, hoist ? &br : nullptr, 3);
        }
        if (rc != MFT_OK) return rc;
        if (hoist) {
            rc = gconv_fwd_check(B, N, G.F[l], G.nout[l], G.ldx, ldo, &p->l[l], last ? 0 : 1);
            if (rc == MFT_OK)
                rc = gconv_fwd_products(G.xcat, G.ldx, B, N, G.F[l], G.nout[l], &p->l[l], dst, ldo, G.gc_saved[l],
                                        br.stream(3));
            if (rc != MFT_OK) return rc;
        }
        if (!last) {
            // tables and weight images of the next Wcompute (parameters only) beside this layer's Gconv
            rc = prepare(l + 1, br.fork(2));
            if (rc != MFT_OK) return rc;
        }
        {
            ProfScope span(PC_SPAN_GC_FWD, st, false);
            if (hoist) {
                br.join(3);
                rc = gconv_fwd_finish(G.adj[l], B, N, G.F[l], G.nout[l], &p->l[l], last ? 0 : 1, dst, ldo,
                                      G.gc_saved[l], st);
            } else {
                rc = gconv_fwd(G.adj[l], G.xcat, G.ldx, B, N, G.F[l], G.nout[l], &p->l[l], last ? 0 : 1, dst, ldo,
                               G.gc_saved[l], G.gc_ws, st);
            }
            if (rc != MFT_OK) return rc;
            br.join(2);
        }
    }
    MFT_REQUIRE(br.ok(), "gnn_fwd: stream fork/join failed: %s", cudaGetErrorString(cudaGetLastError()));
    return MFT_OK;
}

int gnn_bwd(const float* d_out, int B, int N, int F0, int nf, int n_way, const mft_gnn_params* p, float* dx,
            const mft_gnn_grads* g, void* saved, void* workspace, int precision, const unsigned char* shared_nodes,
            cudaStream_t st) {
    MFT_REQUIRE(B > 0 && N > 0 && F0 > 0 && nf >= 2 && (nf % 2) == 0 && n_way > 0, "gnn_bwd: bad shape");
    GnnLayout G = gnn_layout(B, N, F0, nf, n_way, saved, workspace, precision);
    const int rows = B * N;
    MFT_CHECK_CUDA(cudaMemsetAsync(G.dxcat, 0, sizeof(float) * (size_t)rows * G.ldx, st));
    // slot 2: parameter-gradient finalisation of layer l beside the Gconv backward of layer l-1
    // slot 3: the Gconv's fc.weight gradient products beside the Wcompute backward of the same layer
    // slot 4: tables and dgrad weight images of Wcompute l (parameters only) beside the Gconv backward of layer l
    Branches tail(st);
    for (int l = G.L - 1; l >= 0; --l) {
        const bool last = (l == G.L - 1);
        // upstream of this Gconv: d_out for the last one, else the columns it produced in xcat
        // (every later layer has already accumulated its dx there)
        const float* up = last ? d_out : G.dxcat + G.F[l];
        int ldu = last ? n_way : G.ldx;
        int rc;
        const unsigned char* sh = l == 0 ? shared_nodes : nullptr;
        rc = wcompute_bwd_prepare(B, N, G.F[l], nf, &p->w[l], G.wc_saved[l], G.wc_ws, precision, sh, tail.fork(4));
        if (rc != MFT_OK) return rc;
        {
            ProfScope span(PC_SPAN_GC_BWD, st, false);
            tail.join(3);      // the Gconv workspace (dY, T) is about to be rewritten
            rc = gconv_bwd(G.adj[l], G.xcat, G.ldx, B, N, G.F[l], G.nout[l], &p->l[l], last ? 0 : 1, up, ldu,
                           G.dxcat, G.d_adj, &g->l[l], G.gc_saved[l], G.gc_ws, st, &tail, 3);
            if (rc != MFT_OK) return rc;
            tail.join(2);      // the Wcompute workspace (partial dW copies, reductions) is about to be reused
            tail.join(4);
        }
        {
            ProfScope span(PC_SPAN_WC_BWD, st, false);
            rc = wcompute_bwd(G.xcat, G.ldx, B, N, G.F[l], nf, &p->w[l], G.adj[l], G.d_adj, G.dxcat, &g->w[l],
                              G.wc_saved[l], G.wc_ws, precision, sh, st, &tail, 2, true);
        }
        if (rc != MFT_OK) return rc;
    }
    MFT_CHECK_CUDA(cudaMemcpy2DAsync(dx, sizeof(float) * F0, G.dxcat, sizeof(float) * G.ldx, sizeof(float) * F0,
                                     rows, cudaMemcpyDeviceToDevice, st));
    tail.join(2);
    tail.join(3);
    MFT_REQUIRE(tail.ok(), "gnn_bwd: stream fork/join failed: %s", cudaGetErrorString(cudaGetLastError()));
    return MFT_OK;
}

}  // namespace mft
