// Wcompute (edge MLP + adjacency softmax) forward and backward.
// Replaces Wcompute.forward of the reference (methods/gnn.py:78-132) and the
// autograd graph PyTorch builds behind it.  Algorithm: tests/kernel_model.py
// (unordered pairs with multiplicities, BN-cancelled conv biases dropped,
// closed-form backward -- SURVEY.md Appendix A).
#include <cstdlib>
#include "common.cuh"
#include "simt_gemm.cuh"
#include "wcompute.cuh"
#include "prof.cuh"

namespace mft {

// =========================== operand functors ================================

// A(r, k) = |x[b,i,k] - x[b,j,k]|   (gnn.py:79-81)
struct AbsDiffA {
    const float* x;
    int ldx;
    PairGeom g;
    struct Ctx { const float* xi; const float* xj; };
    __device__ __forceinline__ void init(float*) const {}
    __device__ __forceinline__ Ctx row(int r) const {
        PairRow p = decode_row(r, g);
        Ctx c;
        c.xi = x + (size_t)(p.b * g.N + p.i) * ldx;
        c.xj = x + (size_t)(p.b * g.N + p.j) * ldx;
        return c;
    }
    __device__ __forceinline__ float at(const Ctx& c, int k, const float*) const {
        return fabsf(__ldg(c.xi + k) - __ldg(c.xj + k));
    }
    // (row, col) form for the wgrad kernel
    __device__ __forceinline__ float at(int r, int k, const float*) const {
        PairRow p = decode_row(r, g);
        const float* xi = x + (size_t)(p.b * g.N + p.i) * ldx;
        const float* xj = x + (size_t)(p.b * g.N + p.j) * ldx;
        return fabsf(__ldg(xi + k) - __ldg(xj + k));
    }
};

// A(r, k) = LeakyReLU(BN(H[r,k]))  with batch statistics from `sums` (gnn.py:85-86)
struct BnActA {
    const float* H;
    int C;
    const double* sums;
    const float* gamma;
    const float* beta;
    double inv_count;
    struct Ctx { const float* row; };
    __device__ __forceinline__ void init(float* aux) const {
        bn_smem_fill(bn_smem_at(aux), sums, gamma, beta, C, inv_count);
    }
    __device__ __forceinline__ Ctx row(int r) const { return Ctx{H + (size_t)r * C}; }
    __device__ __forceinline__ float at(const Ctx& c, int k, const float* aux) const {
        BnSmem s = bn_smem_at(const_cast<float*>(aux));
        float hh = (c.row[k] - s.mean[k]) * s.rstd[k];
        return lrelu(fmaf(hh, s.gamma[k], s.beta[k]));
    }
    __device__ __forceinline__ float at(int r, int k, const float* aux) const {
        return at(Ctx{H + (size_t)r * C}, k, aux);
    }
};

struct PlainA {
    const float* p;
    int ld;
    struct Ctx { const float* row; };
    __device__ __forceinline__ void init(float*) const {}
    __device__ __forceinline__ Ctx row(int r) const { return Ctx{p + (size_t)r * ld}; }
    __device__ __forceinline__ float at(const Ctx& c, int k, const float*) const { return c.row[k]; }
    __device__ __forceinline__ float at(int r, int k, const float*) const { return p[(size_t)r * ld + k]; }
};

// =========================== epilogue functors ===============================

#define MFT_EPI_COL(j) (c0 + 32 * ((j) >> 1) + ((j) & 1))

// store pre-BN H and accumulate the weighted batch statistics
struct EpiFwdStats {
    static constexpr int kStats = 1;
    float* H;
    int C;
    double* sums;
    PairGeom g;
    __device__ __forceinline__ void init(float*) const {}
    __device__ __forceinline__ void tile(int r0, int c0, int M, int N, float (&acc)[8][6], float (&s0)[6],
                                         float (&s1)[6], const float*) const {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int r = r0 + i;
            if (r >= M) continue;
            float w = decode_row(r, g).w;
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                int c = MFT_EPI_COL(j);
                if (c < N) {
                    float v = acc[i][j];
                    H[(size_t)r * C + c] = v;
                    s0[j] = fmaf(w, v, s0[j]);
                    s1[j] = fmaf(w * v, v, s1[j]);
                }
            }
        }
    }
    __device__ __forceinline__ void commit(int c, float v0, float v1) const {
        stat_add(sums, C, c, 0, v0);
        stat_add(sums, C, c, 1, v1);
    }
};

// dgrad epilogue: acc = dL/d a_{k-1}; dy = acc * lrelu'(BN(H_{k-1})); store dy and the two
// BN-backward reductions sum(dy), sum(dy * hhat)
struct EpiDy {
    static constexpr int kStats = 1;
    const float* H;      // pre-BN activations of layer k-1
    float* dy;
    int C;
    const double* fsums; // forward statistics of layer k-1
    const float* gamma;
    const float* beta;
    double inv_count;
    double* bsums;
    __device__ __forceinline__ void init(float* aux) const {
        bn_smem_fill(bn_smem_at(aux), fsums, gamma, beta, C, inv_count);
    }
    __device__ __forceinline__ void tile(int r0, int c0, int M, int N, float (&acc)[8][6], float (&s0)[6],
                                         float (&s1)[6], const float* aux) const {
        BnSmem s = bn_smem_at(const_cast<float*>(aux));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int r = r0 + i;
            if (r >= M) continue;
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                int c = MFT_EPI_COL(j);
                if (c < N) {
                    float hh = (H[(size_t)r * C + c] - s.mean[c]) * s.rstd[c];
                    float y = fmaf(hh, s.gamma[c], s.beta[c]);
                    float d = acc[i][j] * dlrelu(y);
                    dy[(size_t)r * C + c] = d;
                    s0[j] += d;
                    s1[j] = fmaf(d, hh, s1[j]);
                }
            }
        }
    }
    __device__ __forceinline__ void commit(int c, float v0, float v1) const {
        stat_add(bsums, C, c, 0, v0);
        stat_add(bsums, C, c, 1, v1);
    }
};

// layer-1 dgrad epilogue: acc = dL/dD (summed over the ordered twins); scatter
// dx_i += sign(x_i - x_j) * acc, dx_j -= the same (abs + broadcast sub, gnn.py:79-81)
struct EpiDx {
    static constexpr int kStats = 0;
    const float* x;
    float* dx;
    int ldx;
    PairGeom g;
    __device__ __forceinline__ void init(float*) const {}
    __device__ __forceinline__ void tile(int r0, int c0, int M, int N, float (&acc)[8][6], float (&)[6],
                                         float (&)[6], const float*) const {
        float run[6];
        int run_node = -1;
#pragma unroll
        for (int j = 0; j < 6; ++j) run[j] = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int r = r0 + i;
            if (r >= M) break;
            PairRow p = decode_row(r, g);
            int ni = p.b * g.N + p.i, nj = p.b * g.N + p.j;
            if (ni != run_node) {
                if (run_node >= 0) flush(run_node, c0, N, run);
                run_node = ni;
            }
            if (ni == nj) continue;   // sign(0) = 0 on the diagonal
            const float* xi = x + (size_t)ni * ldx;
            const float* xj = x + (size_t)nj * ldx;
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                int c = MFT_EPI_COL(j);
                if (c < N) {
                    float df = __ldg(xi + c) - __ldg(xj + c);
                    float sg = (df > 0.f) ? 1.f : ((df < 0.f) ? -1.f : 0.f);
                    float v = sg * acc[i][j];
                    run[j] += v;
                    if (v != 0.f) atomicAdd(dx + (size_t)nj * ldx + c, -v);
                }
            }
        }
        if (run_node >= 0) flush(run_node, c0, N, run);
    }
    __device__ __forceinline__ void flush(int node, int c0, int N, float (&run)[6]) const {
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            int c = MFT_EPI_COL(j);
            if (c < N && run[j] != 0.f) atomicAdd(dx + (size_t)node * ldx + c, run[j]);
            run[j] = 0.f;
        }
    }
    __device__ __forceinline__ void commit(int, float, float) const {}
};

// =========================== small kernels ===================================

// Pair tables (common.cuh): tri lists the Rs shared pairs first, then the Rq per-graph pairs, each
// in row-major (i, j >= i) order; inv maps (i, j) back.  With no shared node this is the plain
// triangular enumeration.  The rank of a pair among the shared (or the other) pairs is a closed form
// of s(n) = number of shared nodes below n.
__device__ __forceinline__ int shared_below(const NodeMask& m, int n) {
    int c = 0;
    for (int q = 0; q < (n >> 5); ++q) c += __popc(m.w[q]);
    return c + __popc(m.w[n >> 5] & ((1u << (n & 31)) - 1u));
}

__global__ void tri_table_kernel(int* tri, int* inv, float* roww, int2* rowij, int B, int N, int Rg, int Rs,
                                 int n_shared, NodeMask mask) {
    int rl = blockIdx.x * blockDim.x + threadIdx.x;
    const int Rq = Rg - Rs, R = Rs + B * Rq;
    if (rl < 64) roww[R + rl] = 0.f;     // "past the end" entries (the wgrad kernel copies whole 32-row slabs)
    if (rl == 0) rowij[R] = make_int2(0, 0);
    if (rl >= Rg) return;
    int i, j;
    decode_local(rl, N, i, j);
    const int packed = (j << 16) | i;
    int q;                               // index inside the per-graph part, or -1 for a shared pair
    if (Rs == 0) {
        tri[rl] = packed;
        inv[i * N + j] = rl;
        q = rl;
    } else {
        const bool mi = (mask.w[i >> 5] >> (i & 31)) & 1u, mj = (mask.w[j >> 5] >> (j & 31)) & 1u;
        const int si = shared_below(mask, i), sj = shared_below(mask, j);
        const int before = si * n_shared - (si * (si - 1)) / 2 + (mi ? (sj - si) : 0);   // shared pairs ahead of (i,j)
        if (mi && mj) {
            tri[before] = packed;
            inv[i * N + j] = -1 - before;
            roww[before] = (float)((i == j) ? B : 2 * B);
            rowij[before] = make_int2(i, j);
            q = -1;
        } else {
            tri[Rs + rl - before] = packed;
            inv[i * N + j] = rl - before;
            q = rl - before;
        }
    }
    if (q >= 0) {
        const float w = (i == j) ? 1.f : 2.f;
        for (int b = 0; b < B; ++b) {    // consecutive threads -> consecutive rows of graph b
            roww[Rs + b * Rq + q] = w;
            rowij[Rs + b * Rq + q] = make_int2(b * N + i, b * N + j);
        }
    }
}

constexpr int kRowWarps = 8;   // warps per CTA in the warp-per-row kernels
constexpr int kChPerLane = kMaxC / 32;

// S[b,i,j] = S[b,j,i] = conv2d_last(LeakyReLU(BN4(H4[r])))   (gnn.py:99-103)
template <bool HalfTape, bool Shared>
__global__ void __launch_bounds__(kRowWarps * 32)
score_kernel(const void* __restrict__ H4, int C, const double* sums, const float* gamma, const float* beta,
             const float* last_w, const float* last_b, PairGeom g, float* __restrict__ S) {
    __shared__ float aux[4 * kMaxC];
    __shared__ float wl[kMaxC];
    bn_smem_fill(bn_smem_at(aux), sums, gamma, beta, C, g.inv_pairs);
    for (int c = threadIdx.x; c < C; c += blockDim.x) wl[c] = last_w[c];
    __syncthreads();
    BnSmem s = bn_smem_at(aux);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float bias = last_b[0];
    // two rows per warp iteration: twice the loads in flight (the kernel is a pure HBM stream of H4)
    const int stride = gridDim.x * kRowWarps;
    for (int r = blockIdx.x * kRowWarps + warp; r < g.R; r += 2 * stride) {
        const int r2 = r + stride;
        const bool has2 = r2 < g.R;
        const size_t o1 = (size_t)r * C, o2 = (size_t)(has2 ? r2 : r) * C;
        float acc = 0.f, acc2 = 0.f;
        for (int c = lane; c < C; c += 32) {
            float h1 = TapeH<HalfTape>::ld(H4, o1 + c), h2 = TapeH<HalfTape>::ld(H4, o2 + c);
            float hh = (h1 - s.mean[c]) * s.rstd[c];
            float hh2 = (h2 - s.mean[c]) * s.rstd[c];
            acc = fmaf(lrelu(fmaf(hh, s.gamma[c], s.beta[c])), wl[c], acc);
            acc2 = fmaf(lrelu(fmaf(hh2, s.gamma[c], s.beta[c])), wl[c], acc2);
        }
        acc = warp_sum(acc);
        acc2 = warp_sum(acc2);
        if (Shared || lane == 0) {   // a shared pair's score goes to every graph (lanes over the graphs)
            PairRow p = decode_row(r, g);
            float v = acc + bias;
            for (int b = p.b + lane; b < p.b + p.nb; b += 32) {
                size_t base = (size_t)b * g.N * g.N;
                S[base + (size_t)p.i * g.N + p.j] = v;
                S[base + (size_t)p.j * g.N + p.i] = v;
            }
            if (has2) {
                PairRow q = decode_row(r2, g);
                float v2 = acc2 + bias;
                for (int b = q.b + lane; b < q.b + q.nb; b += 32) {
                    size_t base2 = (size_t)b * g.N * g.N;
                    S[base2 + (size_t)q.i * g.N + q.j] = v2;
                    S[base2 + (size_t)q.j * g.N + q.i] = v2;
                }
            }
        }
    }
}

// ---- vector forms of score / dy4 (C % 4 == 0): lane l owns channels 4l..4l+3 (+128 per extra group), its
// BN constants live in registers, H4 is read with one 8-byte (fp16 tape) or 16-byte (fp32) load per
// group and dy4 written with one 16-byte store.  The scalar kernels above/below stay as the fallback.
template <bool HalfTape>
__device__ __forceinline__ float4 tape_ld4(const void* p, size_t i) {
    if constexpr (HalfTape) return unpack_half4(__ldg(reinterpret_cast<const uint2*>(static_cast<const __half*>(p) + i)));
    else return __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(p) + i));
}

template <int NG>
struct LaneConsts {
    float4 sc[NG], sh[NG], wl[NG], mean[NG], rstd[NG];
    bool on[NG];
};

template <int NG>
__device__ __forceinline__ void lane_consts(LaneConsts<NG>& k, const float* aux, const float* wl, int C, int lane) {
    BnSmem s = bn_smem_at(const_cast<float*>(aux));
#pragma unroll
    for (int q = 0; q < NG; ++q) {
        const int c = 4 * (lane + 32 * q);
        k.on[q] = c < C;
        const int cc = k.on[q] ? c : 0;
        float sc[4], sh[4], w[4], m[4], r[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            sc[e] = s.gamma[cc + e] * s.rstd[cc + e];
            sh[e] = s.beta[cc + e] - s.mean[cc + e] * sc[e];
            w[e] = wl[cc + e];
            m[e] = s.mean[cc + e];
            r[e] = s.rstd[cc + e];
        }
        k.sc[q] = make_float4(sc[0], sc[1], sc[2], sc[3]);
        k.sh[q] = make_float4(sh[0], sh[1], sh[2], sh[3]);
        k.wl[q] = make_float4(w[0], w[1], w[2], w[3]);
        k.mean[q] = make_float4(m[0], m[1], m[2], m[3]);
        k.rstd[q] = make_float4(r[0], r[1], r[2], r[3]);
    }
}

template <bool HalfTape, bool Shared, int NG>
__global__ void __launch_bounds__(kRowWarps * 32)
score_vec_kernel(const void* __restrict__ H4, int C, const double* sums, const float* gamma, const float* beta,
                 const float* last_w, const float* last_b, PairGeom g, float* __restrict__ S) {
    __shared__ float aux[4 * kMaxC];
    __shared__ float wls[kMaxC];
    bn_smem_fill(bn_smem_at(aux), sums, gamma, beta, C, g.inv_pairs);
    for (int c = threadIdx.x; c < C; c += blockDim.x) wls[c] = last_w[c];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    LaneConsts<NG> k;
    lane_consts<NG>(k, aux, wls, C, lane);
    const float bias = last_b[0];
    constexpr int RU = 4;                      // rows in flight per warp
    const int stride = gridDim.x * kRowWarps;
    for (int r0 = blockIdx.x * kRowWarps + warp; r0 < g.R; r0 += RU * stride) {
        float4 h[RU][NG];
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            const int r = min(r0 + u * stride, g.R - 1);
#pragma unroll
            for (int q = 0; q < NG; ++q)
                h[u][q] = k.on[q] ? tape_ld4<HalfTape>(H4, (size_t)r * C + 4 * (lane + 32 * q))
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            const int r = r0 + u * stride;
            if (r >= g.R) break;
            float acc = 0.f;
#pragma unroll
            for (int q = 0; q < NG; ++q) {
                if (!k.on[q]) continue;
                acc = fmaf(lrelu(fmaf(h[u][q].x, k.sc[q].x, k.sh[q].x)), k.wl[q].x, acc);
                acc = fmaf(lrelu(fmaf(h[u][q].y, k.sc[q].y, k.sh[q].y)), k.wl[q].y, acc);
                acc = fmaf(lrelu(fmaf(h[u][q].z, k.sc[q].z, k.sh[q].z)), k.wl[q].z, acc);
                acc = fmaf(lrelu(fmaf(h[u][q].w, k.sc[q].w, k.sh[q].w)), k.wl[q].w, acc);
            }
            acc = warp_sum(acc);
            if (Shared || lane == 0) {
                PairRow p = decode_row(r, g);
                const float v = acc + bias;
                for (int b = p.b + lane; b < p.b + p.nb; b += 32) {
                    size_t base = (size_t)b * g.N * g.N;
                    S[base + (size_t)p.i * g.N + p.j] = v;
                    S[base + (size_t)p.j * g.N + p.i] = v;
                }
            }
        }
    }
}

template <bool HalfTape, bool Shared, int NG>
__global__ void __launch_bounds__(kRowWarps * 32)
dy4_vec_kernel(const float* __restrict__ dS, const void* __restrict__ H4, int C, const double* fsums,
               const float* gamma, const float* beta, const float* last_w, PairGeom g, float* __restrict__ dy4,
               double* bsums, double* lastsum) {
    __shared__ float aux[4 * kMaxC];
    __shared__ float wls[kMaxC];
    __shared__ float red[3][kRowWarps][kMaxC];
    bn_smem_fill(bn_smem_at(aux), fsums, gamma, beta, C, g.inv_pairs);
    for (int c = threadIdx.x; c < C; c += blockDim.x) wls[c] = last_w[c];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    LaneConsts<NG> k;
    lane_consts<NG>(k, aux, wls, C, lane);
    float4 p0[NG], p1[NG], p2[NG];             // sum d, sum d*h (turned into sum d*hhat at the end), sum G*a4
#pragma unroll
    for (int q = 0; q < NG; ++q) p0[q] = p1[q] = p2[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    constexpr int RU = 4;
    const int stride = gridDim.x * kRowWarps;
    for (int r0 = blockIdx.x * kRowWarps + warp; r0 < g.R; r0 += RU * stride) {
        float G[RU];
        float4 h[RU][NG];
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            const int r = min(r0 + u * stride, g.R - 1);
            PairRow p = decode_row(r, g);
            if (!Shared || p.nb == 1) {
                size_t base = (size_t)p.b * g.N * g.N;
                float a = __ldg(dS + base + (size_t)p.i * g.N + p.j);
                float b2 = __ldg(dS + base + (size_t)p.j * g.N + p.i);
                G[u] = (p.i != p.j) ? a + b2 : a;
            } else {
                float a = 0.f;
                for (int b = lane; b < p.nb; b += 32) {
                    size_t base = (size_t)b * g.N * g.N;
                    a += __ldg(dS + base + (size_t)p.i * g.N + p.j);
                    if (p.i != p.j) a += __ldg(dS + base + (size_t)p.j * g.N + p.i);
                }
                G[u] = warp_sum(a);
            }
#pragma unroll
            for (int q = 0; q < NG; ++q)
                h[u][q] = k.on[q] ? tape_ld4<HalfTape>(H4, (size_t)r * C + 4 * (lane + 32 * q))
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            const int r = r0 + u * stride;
            if (r >= g.R) break;
#pragma unroll
            for (int q = 0; q < NG; ++q) {
                if (!k.on[q]) continue;
                const float4 hv = h[u][q];
                float4 y, d;
                y.x = fmaf(hv.x, k.sc[q].x, k.sh[q].x); y.y = fmaf(hv.y, k.sc[q].y, k.sh[q].y);
                y.z = fmaf(hv.z, k.sc[q].z, k.sh[q].z); y.w = fmaf(hv.w, k.sc[q].w, k.sh[q].w);
                d.x = G[u] * k.wl[q].x * dlrelu(y.x); d.y = G[u] * k.wl[q].y * dlrelu(y.y);
                d.z = G[u] * k.wl[q].z * dlrelu(y.z); d.w = G[u] * k.wl[q].w * dlrelu(y.w);
                *reinterpret_cast<float4*>(dy4 + (size_t)r * C + 4 * (lane + 32 * q)) = d;
                p0[q].x += d.x; p0[q].y += d.y; p0[q].z += d.z; p0[q].w += d.w;
                p1[q].x = fmaf(d.x, hv.x, p1[q].x); p1[q].y = fmaf(d.y, hv.y, p1[q].y);
                p1[q].z = fmaf(d.z, hv.z, p1[q].z); p1[q].w = fmaf(d.w, hv.w, p1[q].w);
                p2[q].x = fmaf(G[u], lrelu(y.x), p2[q].x); p2[q].y = fmaf(G[u], lrelu(y.y), p2[q].y);
                p2[q].z = fmaf(G[u], lrelu(y.z), p2[q].z); p2[q].w = fmaf(G[u], lrelu(y.w), p2[q].w);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < NG; ++q) {
        if (!k.on[q]) continue;
        const int c = 4 * (lane + 32 * q);
        // sum d*hhat = rstd * (sum d*h - mean * sum d)
        const float4 hh = make_float4(k.rstd[q].x * (p1[q].x - k.mean[q].x * p0[q].x),
                                      k.rstd[q].y * (p1[q].y - k.mean[q].y * p0[q].y),
                                      k.rstd[q].z * (p1[q].z - k.mean[q].z * p0[q].z),
                                      k.rstd[q].w * (p1[q].w - k.mean[q].w * p0[q].w));
        *reinterpret_cast<float4*>(&red[0][warp][c]) = p0[q];
        *reinterpret_cast<float4*>(&red[1][warp][c]) = hh;
        *reinterpret_cast<float4*>(&red[2][warp][c]) = p2[q];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float v0 = 0.f, v1 = 0.f, v2 = 0.f;
#pragma unroll
        for (int w = 0; w < kRowWarps; ++w) { v0 += red[0][w][c]; v1 += red[1][w][c]; v2 += red[2][w][c]; }
        stat_add(bsums, C, c, 0, v0);
        stat_add(bsums, C, c, 1, v1);
        stat_add(lastsum, C, c, 0, v2);
    }
}

// ---- eight lanes per row (C % 4 == 0, C <= 32*NGL): a warp works on four rows at a time, lane (l & 7)
// owns the 4-channel groups (l & 7) + 8q.  Compared with one warp per row this divides the per-row
// overhead (row decode, reduction shuffles, address arithmetic, scatter) by four, which is what
// bounds these kernels once the loads are vectorised (ncu: issue-bound at ~1 TB/s).
// Row quads a warp keeps in flight (measured on B200 with tools/ab_build2.sh, 5w20s step): score8 1 / 2 / 3 the
// same; dy8 1: 1.2735 ms, 2: 1.2773, 3: 1.2855 (spills) -- its three reductions leave few registers for loads.
#ifndef MFT_SCORE_RU
#define MFT_SCORE_RU 2
#endif
#ifndef MFT_DY_RU
#define MFT_DY_RU 1
#endif
template <int NGL>
struct Lane8Consts {
    float4 sc[NGL], sh[NGL], wl[NGL];
    bool on[NGL];
};
template <int NGL>
__device__ __forceinline__ void lane8_consts(Lane8Consts<NGL>& k, const float* aux, const float* wl, int C, int sl) {
    BnSmem s = bn_smem_at(const_cast<float*>(aux));
#pragma unroll
    for (int q = 0; q < NGL; ++q) {
        const int c = 4 * (sl + 8 * q);
        k.on[q] = c < C;
        const int cc = k.on[q] ? c : 0;
        float sc[4], sh[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            sc[e] = s.gamma[cc + e] * s.rstd[cc + e];
            sh[e] = s.beta[cc + e] - s.mean[cc + e] * sc[e];
        }
        k.sc[q] = make_float4(sc[0], sc[1], sc[2], sc[3]);
        k.sh[q] = make_float4(sh[0], sh[1], sh[2], sh[3]);
        k.wl[q] = make_float4(wl[cc], wl[cc + 1], wl[cc + 2], wl[cc + 3]);
    }
}
__device__ __forceinline__ float sum8(float v) {   // over the 8 lanes of a row group
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}

template <bool HalfTape, bool Shared, int NGL>
__global__ void __launch_bounds__(kRowWarps * 32)
score8_kernel(const void* __restrict__ H4, int C, const double* sums, const float* gamma, const float* beta,
              const float* last_w, const float* last_b, PairGeom g, float* __restrict__ S) {
    __shared__ float aux[4 * kMaxC];
    __shared__ float wls[kMaxC];
    for (int c = threadIdx.x; c < C; c += blockDim.x) wls[c] = last_w[c];   // a parameter: not upstream data
    pdl_enter();
    bn_smem_fill(bn_smem_at(aux), sums, gamma, beta, C, g.inv_pairs);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sl = lane & 7, rg = lane >> 3;
    Lane8Consts<NGL> k;
    lane8_consts<NGL>(k, aux, wls, C, sl);
    const float bias = last_b[0];
    constexpr int RU = MFT_SCORE_RU;               // row quads in flight per warp
    const int stride = gridDim.x * kRowWarps * 4;  // rows per grid sweep
    for (int rb = (blockIdx.x * kRowWarps + warp) * 4; rb < g.R; rb += RU * stride) {   // warp-uniform trip count
        const int r0 = rb + rg;
        float4 h[RU][NGL];
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            const int r = min(r0 + u * stride, g.R - 1);
#pragma unroll
            for (int q = 0; q < NGL; ++q)
                h[u][q] = k.on[q] ? tape_ld4<HalfTape>(H4, (size_t)r * C + 4 * (sl + 8 * q))
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            const int r = r0 + u * stride;
            float acc = 0.f;
#pragma unroll
            for (int q = 0; q < NGL; ++q) {
                if (!k.on[q]) continue;
                acc = fmaf(lrelu(fmaf(h[u][q].x, k.sc[q].x, k.sh[q].x)), k.wl[q].x, acc);
                acc = fmaf(lrelu(fmaf(h[u][q].y, k.sc[q].y, k.sh[q].y)), k.wl[q].y, acc);
                acc = fmaf(lrelu(fmaf(h[u][q].z, k.sc[q].z, k.sh[q].z)), k.wl[q].z, acc);
                acc = fmaf(lrelu(fmaf(h[u][q].w, k.sc[q].w, k.sh[q].w)), k.wl[q].w, acc);
            }
            acc = sum8(acc);                       // (all lanes take part; rows past the end are dropped below)
            if (r < g.R && (Shared || sl == 0)) {
                PairRow p = decode_row(r, g);
                const float v = acc + bias;
                for (int b = p.b + sl; b < p.b + p.nb; b += 8) {
                    size_t base = (size_t)b * g.N * g.N;
                    S[base + (size_t)p.i * g.N + p.j] = v;
                    S[base + (size_t)p.j * g.N + p.i] = v;
                }
            }
        }
    }
}

template <bool HalfTape, bool Shared, int NGL>
__global__ void __launch_bounds__(kRowWarps * 32, 2)
dy8_kernel(const float* __restrict__ dS, const void* __restrict__ H4, int C, const double* fsums,
           const float* gamma, const float* beta, const float* last_w, PairGeom g, float* __restrict__ dy4,
           double* bsums, double* lastsum) {
    __shared__ float aux[4 * kMaxC];
    __shared__ float wls[kMaxC];
    __shared__ float red[3][kRowWarps][32 * NGL];
    for (int c = threadIdx.x; c < C; c += blockDim.x) wls[c] = last_w[c];
    pdl_enter();
    bn_smem_fill(bn_smem_at(aux), fsums, gamma, beta, C, g.inv_pairs);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sl = lane & 7, rg = lane >> 3;
    Lane8Consts<NGL> k;
    lane8_consts<NGL>(k, aux, wls, C, sl);
    float4 p0[NGL], p1[NGL], p2[NGL];
#pragma unroll
    for (int q = 0; q < NGL; ++q) p0[q] = p1[q] = p2[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    constexpr int RU = MFT_DY_RU;
    const int stride = gridDim.x * kRowWarps * 4;
    for (int rb = (blockIdx.x * kRowWarps + warp) * 4; rb < g.R; rb += RU * stride) {   // warp-uniform trip count
        const int r0 = rb + rg;
        float G[RU];
        float4 h[RU][NGL];
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            const int r = min(r0 + u * stride, g.R - 1);
            PairRow p = decode_row(r, g);
            float a = 0.f;
            if (!Shared || p.nb == 1) {
                if (sl == 0) {
                    size_t base = (size_t)p.b * g.N * g.N;
                    a = __ldg(dS + base + (size_t)p.i * g.N + p.j);
                    if (p.i != p.j) a += __ldg(dS + base + (size_t)p.j * g.N + p.i);
                }
            } else {   // shared pair: sum over the graphs it stands for (8 lanes over b)
                for (int b = sl; b < p.nb; b += 8) {
                    size_t base = (size_t)b * g.N * g.N;
                    a += __ldg(dS + base + (size_t)p.i * g.N + p.j);
                    if (p.i != p.j) a += __ldg(dS + base + (size_t)p.j * g.N + p.i);
                }
            }
            G[u] = a;
#pragma unroll
            for (int q = 0; q < NGL; ++q)
                h[u][q] = k.on[q] ? tape_ld4<HalfTape>(H4, (size_t)r * C + 4 * (sl + 8 * q))
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            const int r = r0 + u * stride;
            const float Gr = sum8(G[u]);
            if (r >= g.R) continue;
#pragma unroll
            for (int q = 0; q < NGL; ++q) {
                if (!k.on[q]) continue;
                const float4 hv = h[u][q];
                float4 y, d;
                y.x = fmaf(hv.x, k.sc[q].x, k.sh[q].x); y.y = fmaf(hv.y, k.sc[q].y, k.sh[q].y);
                y.z = fmaf(hv.z, k.sc[q].z, k.sh[q].z); y.w = fmaf(hv.w, k.sc[q].w, k.sh[q].w);
                d.x = Gr * k.wl[q].x * dlrelu(y.x); d.y = Gr * k.wl[q].y * dlrelu(y.y);
                d.z = Gr * k.wl[q].z * dlrelu(y.z); d.w = Gr * k.wl[q].w * dlrelu(y.w);
                *reinterpret_cast<float4*>(dy4 + (size_t)r * C + 4 * (sl + 8 * q)) = d;
                p0[q].x += d.x; p0[q].y += d.y; p0[q].z += d.z; p0[q].w += d.w;
                p1[q].x = fmaf(d.x, hv.x, p1[q].x); p1[q].y = fmaf(d.y, hv.y, p1[q].y);
                p1[q].z = fmaf(d.z, hv.z, p1[q].z); p1[q].w = fmaf(d.w, hv.w, p1[q].w);
                p2[q].x = fmaf(Gr, lrelu(y.x), p2[q].x); p2[q].y = fmaf(Gr, lrelu(y.y), p2[q].y);
                p2[q].z = fmaf(Gr, lrelu(y.z), p2[q].z); p2[q].w = fmaf(Gr, lrelu(y.w), p2[q].w);
            }
        }
    }
    // combine the four row groups of the warp (lanes with equal sl), then the warps through shared memory
#pragma unroll
    for (int q = 0; q < NGL; ++q) {
        float* ps[3] = {&p0[q].x, &p1[q].x, &p2[q].x};
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float v = ps[a][e];
                v += __shfl_xor_sync(0xffffffffu, v, 8);
                v += __shfl_xor_sync(0xffffffffu, v, 16);
                ps[a][e] = v;
            }
        if (rg == 0) {
            const int c = 4 * (sl + 8 * q);
            const float* mean = aux, *rstd = aux + kMaxC;           // (bn_smem_at layout)
            const float4 hh = make_float4(rstd[c] * (p1[q].x - mean[c] * p0[q].x),
                                          rstd[c + 1] * (p1[q].y - mean[c + 1] * p0[q].y),
                                          rstd[c + 2] * (p1[q].z - mean[c + 2] * p0[q].z),
                                          rstd[c + 3] * (p1[q].w - mean[c + 3] * p0[q].w));
            *reinterpret_cast<float4*>(&red[0][warp][c]) = p0[q];
            *reinterpret_cast<float4*>(&red[1][warp][c]) = hh;      // sum d*hhat = rstd*(sum d*h - mean*sum d)
            *reinterpret_cast<float4*>(&red[2][warp][c]) = p2[q];
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float v0 = 0.f, v1 = 0.f, v2 = 0.f;
#pragma unroll
        for (int w = 0; w < kRowWarps; ++w) { v0 += red[0][w][c]; v1 += red[1][w][c]; v2 += red[2][w][c]; }
        stat_add(bsums, C, c, 0, v0);
        stat_add(bsums, C, c, 1, v1);
        stat_add(lastsum, C, c, 0, v2);
    }
}

// adj[b,i,:] = softmax_j(S[b,i,j] - 1e8 [i==j])   (gnn.py:105-115), one warp per row
__global__ void __launch_bounds__(kRowWarps * 32)
softmax_rows_kernel(const float* __restrict__ S, float* __restrict__ adj, int rows, int N) {
    pdl_enter();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int row = blockIdx.x * kRowWarps + warp;
    if (row >= rows) return;
    int i = row % N;
    const float* s = S + (size_t)row * N;
    float* a = adj + (size_t)row * N;
    float mx = -INFINITY;
    for (int j = lane; j < N; j += 32) {
        float v = s[j] - (j == i ? kDiagMask : 0.f);
        mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < N; j += 32) {
        float v = s[j] - (j == i ? kDiagMask : 0.f);
        sum += expf(v - mx);
    }
    sum = warp_sum(sum);
    float inv = 1.f / sum;
    for (int j = lane; j < N; j += 32) {
        float v = s[j] - (j == i ? kDiagMask : 0.f);
        a[j] = expf(v - mx) * inv;
    }
}

// dS[b,i,j] = A_ij (dA_ij - sum_k A_ik dA_ik), one warp per row
__global__ void __launch_bounds__(kRowWarps * 32)
softmax_bwd_kernel(const float* __restrict__ adj, const float* __restrict__ d_adj, float* __restrict__ dS,
                   int rows, int N) {
    pdl_enter();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int row = blockIdx.x * kRowWarps + warp;
    if (row >= rows) return;
    const float* a = adj + (size_t)row * N;
    const float* d = d_adj + (size_t)row * N;
    float dot = 0.f;
    for (int j = lane; j < N; j += 32) dot = fmaf(a[j], d[j], dot);
    dot = warp_sum(dot);
    for (int j = lane; j < N; j += 32) dS[(size_t)row * N + j] = a[j] * (d[j] - dot);
}

// Backward through conv2d_last and the layer-4 LeakyReLU:
//   G_r = dS_ij + dS_ji, dy4 = G_r * w_last * lrelu'(y4); reductions sum dy4, sum dy4*hhat4,
//   sum G_r * a4 (= d conv2d_last.weight).  One warp per row, lanes over channels.
template <bool HalfTape, bool Shared>
__global__ void __launch_bounds__(kRowWarps * 32)
dy4_kernel(const float* __restrict__ dS, const void* __restrict__ H4, int C, const double* fsums,
           const float* gamma, const float* beta, const float* last_w, PairGeom g, void* __restrict__ dy4,
           double* bsums, double* lastsum) {
    __shared__ float aux[4 * kMaxC];
    __shared__ float wl[kMaxC];
    __shared__ float red[3][kRowWarps][kMaxC];
    bn_smem_fill(bn_smem_at(aux), fsums, gamma, beta, C, g.inv_pairs);
    for (int c = threadIdx.x; c < C; c += blockDim.x) wl[c] = last_w[c];
    __syncthreads();
    BnSmem s = bn_smem_at(aux);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float p0[kChPerLane], p1[kChPerLane], p2[kChPerLane];
#pragma unroll
    for (int q = 0; q < kChPerLane; ++q) { p0[q] = 0.f; p1[q] = 0.f; p2[q] = 0.f; }
    // four rows per warp iteration: all their H4 / dS loads are issued before any is consumed
    constexpr int RU = 4;
    const int stride = gridDim.x * kRowWarps;
    for (int r0 = blockIdx.x * kRowWarps + warp; r0 < g.R; r0 += RU * stride) {
        float G[RU];
        float hraw[RU][kChPerLane];
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            const int r = min(r0 + u * stride, g.R - 1);
            PairRow p = decode_row(r, g);
            if (!Shared || p.nb == 1) {   // (Shared is a template switch: the branch would serialise the loads)
                size_t base = (size_t)p.b * g.N * g.N;
                float a = __ldg(dS + base + (size_t)p.i * g.N + p.j);
                float b2 = __ldg(dS + base + (size_t)p.j * g.N + p.i);
                G[u] = (p.i != p.j) ? a + b2 : a;
            } else {   // shared pair: the row's gradient is the sum over the graphs it stands for
                float a = 0.f;
                for (int b = lane; b < p.nb; b += 32) {
                    size_t base = (size_t)b * g.N * g.N;
                    a += __ldg(dS + base + (size_t)p.i * g.N + p.j);
                    if (p.i != p.j) a += __ldg(dS + base + (size_t)p.j * g.N + p.i);
                }
                G[u] = warp_sum(a);
            }
#pragma unroll
            for (int q = 0; q < kChPerLane; ++q) {
                int c = lane + 32 * q;
                hraw[u][q] = (c < C) ? TapeH<HalfTape>::ld(H4, (size_t)r * C + c) : 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            const int r = r0 + u * stride;
            if (r >= g.R) break;
#pragma unroll
            for (int q = 0; q < kChPerLane; ++q) {
                int c = lane + 32 * q;
                if (c < C) {
                    float hh = (hraw[u][q] - s.mean[c]) * s.rstd[c];
                    float y = fmaf(hh, s.gamma[c], s.beta[c]);
                    float d = G[u] * wl[c] * dlrelu(y);
                    TapeD<false>::st(dy4, (size_t)r * C + c, d);   // gradients stay fp32 on both paths
                    p0[q] += d;
                    p1[q] = fmaf(d, hh, p1[q]);
                    p2[q] = fmaf(G[u], lrelu(y), p2[q]);
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < kChPerLane; ++q) {
        red[0][warp][lane + 32 * q] = p0[q];
        red[1][warp][lane + 32 * q] = p1[q];
        red[2][warp][lane + 32 * q] = p2[q];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float v0 = 0.f, v1 = 0.f, v2 = 0.f;
#pragma unroll
        for (int w = 0; w < kRowWarps; ++w) { v0 += red[0][w][c]; v1 += red[1][w][c]; v2 += red[2][w][c]; }
        stat_add(bsums, C, c, 0, v0);
        stat_add(bsums, C, c, 1, v1);
        stat_add(lastsum, C, c, 0, v2);
    }
}

// BatchNorm backward applied in place: dH = gamma*rstd*(dy - w*m1 - w*hhat*m2), the twin-summed
// form of the dense formula (w = row multiplicity; m1, m2 = means over all B*N*N ordered pairs).
__global__ void __launch_bounds__(kRowWarps * 32)
dh_kernel(float* __restrict__ dy, const float* __restrict__ H, int C, const double* fsums, const float* gamma,
          const double* bsums, PairGeom g) {
    __shared__ float aux[4 * kMaxC];
    __shared__ float m1[kMaxC], m2[kMaxC];
    bn_smem_fill(bn_smem_at(aux), fsums, gamma, nullptr, C, g.inv_pairs);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        m1[c] = (float)(stat_get(bsums, C, c, 0) * g.inv_pairs);
        m2[c] = (float)(stat_get(bsums, C, c, 1) * g.inv_pairs);
    }
    __syncthreads();
    BnSmem s = bn_smem_at(aux);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = blockIdx.x * kRowWarps + warp; r < g.R; r += gridDim.x * kRowWarps) {
        float w = decode_row(r, g).w;
        const float* hrow = H + (size_t)r * C;
        float* drow = dy + (size_t)r * C;
        for (int c = lane; c < C; c += 32) {
            float hh = (hrow[c] - s.mean[c]) * s.rstd[c];
            drow[c] = s.gamma[c] * s.rstd[c] * (drow[c] - w * m1[c] - w * hh * m2[c]);
        }
    }
}

struct FinalizeArgs {
    const float* wgpart[4];   // tcgen05 path: ncopies[k] partial weight gradients per layer (else null)
    int ncopies[4];
    float* conv_w[4];
    int wsize[4];
    int cin[4];
    const double* bsums[4];
    const double* lastsum;
    float* bn_g[4];
    float* bn_b[4];
    float* conv_b[4];
    float* last_w;
    float* last_b;
    int C[4];
    int nf;
    const float* tscale;      // tcgen05 path: the layers' tape scales s_k (dW = s_k * dW'), else null
};

__global__ void finalize_grads_kernel(FinalizeArgs a) {   // grid = 5: one CTA per BN layer + one for conv2d_last
    pdl_enter();
    const int t = threadIdx.x, k = blockIdx.x;
    if (k < 4) {
        if (t < a.C[k]) {
            if (a.bn_b[k]) a.bn_b[k][t] = (float)stat_get(a.bsums[k], a.C[k], t, 0);
            if (a.bn_g[k]) a.bn_g[k][t] = (float)stat_get(a.bsums[k], a.C[k], t, 1);
            if (a.conv_b[k]) a.conv_b[k][t] = 0.f;   // BN removes the mean: exactly zero
        }
        return;
    }
    if (t < a.nf && a.last_w) a.last_w[t] = (float)stat_get(a.lastsum, a.nf, t, 0);
    if (t == 0 && a.last_b) a.last_b[0] = 0.f;         // softmax shift invariance: exactly zero
}

// Conv weight gradients of the tcgen05 path: add up the partial copies the wgrad CTAs stored.  One CTA
// per (layer, output channel) row; thread (cx, cy) sums copies cy, cy+S, ... of float4 column cx, the
// S slices are combined through shared memory in a fixed order (bitwise reproducible result).
constexpr int kRedThreads = 256;
__global__ void __launch_bounds__(kRedThreads)
wgrad_reduce_kernel(FinalizeArgs a) {
    __shared__ float4 part[kRedThreads];
    pdl_enter();
    int k = 0, co = blockIdx.x;
    while (k < 4 && co >= a.wsize[k] / a.cin[k]) { co -= a.wsize[k] / a.cin[k]; ++k; }
    if (k == 4 || a.wgpart[k] == nullptr) return;
    const int cin = a.cin[k], ldp = (cin + 3) & ~3, cout = a.wsize[k] / cin, nc = a.ncopies[k];
    const int ncol4 = ldp / 4;                       // <= 64
    const int slices = kRedThreads / ncol4;
    const int cx = threadIdx.x % ncol4, cy = threadIdx.x / ncol4;
    const size_t cstride = (size_t)cout * ldp;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (cy < slices) {
        const float* src = a.wgpart[k] + (size_t)co * ldp + cx * 4;
        int q = cy;
        for (; q + 3 * slices < nc; q += 4 * slices) {   // four independent 16-byte loads in flight
            float4 v0 = __ldcs(reinterpret_cast<const float4*>(src + (size_t)q * cstride));
            float4 v1 = __ldcs(reinterpret_cast<const float4*>(src + (size_t)(q + slices) * cstride));
            float4 v2 = __ldcs(reinterpret_cast<const float4*>(src + (size_t)(q + 2 * slices) * cstride));
            float4 v3 = __ldcs(reinterpret_cast<const float4*>(src + (size_t)(q + 3 * slices) * cstride));
            acc.x += (v0.x + v1.x) + (v2.x + v3.x);
            acc.y += (v0.y + v1.y) + (v2.y + v3.y);
            acc.z += (v0.z + v1.z) + (v2.z + v3.z);
            acc.w += (v0.w + v1.w) + (v2.w + v3.w);
        }
        for (; q < nc; q += slices) {
            float4 v = __ldcs(reinterpret_cast<const float4*>(src + (size_t)q * cstride));
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    part[threadIdx.x] = acc;
    __syncthreads();
    if (cy == 0) {
        for (int sl = 1; sl < slices; ++sl) {
            float4 v = part[sl * ncol4 + cx];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        float* dst = a.conv_w[k] + (size_t)co * cin + cx * 4;
        // the GEMMs ran on W' = s W (exact power of two): dL/dW = s * dL/dW'
        const float sc = a.tscale ? a.tscale[k] : 1.f;
        const float o[4] = {acc.x * sc, acc.y * sc, acc.z * sc, acc.w * sc};
        for (int e = 0; e < 4; ++e)
            if (cx * 4 + e < cin) dst[e] = o[e];
    }
}

// =========================== host orchestration ==============================

static inline int row_grid(int R) { return min(cdiv(R, kRowWarps), 148 * 6); }
static inline int row_grid8(int R) { return min(cdiv(R, kRowWarps * 4), 148 * 4); }   // four rows per warp

WcLayout wc_layout(int B, int N, int F, int nf, void* saved, void* workspace, int precision) {
    WcLayout L;
    L.C[0] = F; L.C[1] = 2 * nf; L.C[2] = 2 * nf; L.C[3] = nf; L.C[4] = nf;
    int Rg = N * (N + 1) / 2;
    size_t R = (size_t)B * Rg;
    Carver sv(saved);
    const size_t esz = precision == MFT_PREC_TF32 ? sizeof(__half) : sizeof(float);   // the tensor-core path keeps an fp16 tape
    for (int k = 0; k < 4; ++k) L.H[k] = reinterpret_cast<float*>(sv.take<char>(R * L.C[k + 1] * esz));
    L.fsums = sv.take<double>(4 * kStatSlot);
    L.tscale = sv.take<float>(64);
    L.saved_bytes = sv.used();
    Carver ws(workspace);
    L.tri = ws.take<int>(Rg);
    L.inv = ws.take<int>((size_t)N * N);
    L.roww = ws.take<float>(R + 64);
    L.rowij = ws.take<int2>(R + 4);
    L.S = ws.take<float>((size_t)B * N * N);
    L.dyA = ws.take<float>(R * 2 * nf);
    L.dyB = ws.take<float>(R * 2 * nf);
    L.dyC = ws.take<float>(umma_shape_supported(F, nf) ? R * 2 * nf : 0);
    L.bsums = ws.take<double>(5 * kStatSlot);
    L.wimg = ws.take<float>(umma_workspace_floats(F, nf) + 64);
    L.dD = ws.take<float>(umma_shape_supported(F, nf) ? R * (size_t)((F + 3) & ~3) : 0);
    {
        size_t tot = 0;
        for (int k = 0; k < 4; ++k) {
            L.wgpart_off[k] = tot;
            tot += (size_t)kWgMaxCopies * L.C[k + 1] * ((L.C[k] + 3) & ~3);   // rows padded to 4 floats
        }
        L.wgpart = ws.take<float>(umma_shape_supported(F, nf) ? tot : 0);
        L.wgpart_floats = tot;
    }
    L.workspace_bytes = ws.used();
    return L;
}

int wcompute_fwd_prepare(int B, int N, int F, int nf, const mft_wcompute_params* p, void* saved, void* workspace,
                         int precision, const unsigned char* shared_nodes, cudaStream_t st) {
    MFT_REQUIRE(B > 0 && N > 0 && F > 0 && nf > 0, "wcompute_fwd: bad shape B=%d N=%d F=%d nf=%d", B, N, F, nf);
    MFT_REQUIRE(2 * nf <= kMaxC, "wcompute_fwd: nf=%d exceeds the supported maximum %d", nf, kMaxC / 2);
    MFT_REQUIRE(N < 32768, "wcompute_fwd: N=%d too large", N);
    WcLayout L = wc_layout(B, N, F, nf, saved, workspace, precision);
    NodeMask mask;
    const int n_shared = mask_from_host(shared_nodes, B, N, mask);
    PairGeom g = make_geom(B, N, L.tri, L.inv, n_shared, L.roww, L.rowij);
    MFT_CHECK_CUDA(cudaMemsetAsync(L.fsums, 0, sizeof(double) * 4 * kStatSlot, st));
    {
        ProfScope ps(PC_PREP, st);
        tri_table_kernel<<<cdiv(g.Rg, 256), 256, 0, st>>>(L.tri, L.inv, L.roww, L.rowij, B, N, g.Rg, g.Rs, n_shared, mask);
        MFT_CHECK_LAUNCH();
    }
    if (precision == MFT_PREC_TF32) return wcompute_fwd_prepare_tf32(p, L, F, nf, 1.0 / g.inv_pairs, st);
    return MFT_OK;
}

int wcompute_fwd(const float* x, int ldx, int B, int N, int F, int nf, const mft_wcompute_params* p, float* adj,
                 void* saved, void* workspace, int precision, const unsigned char* shared_nodes, cudaStream_t st,
                 bool prepared, Branches* mid, int mid_slot) {
    MFT_REQUIRE(B > 0 && N > 0 && F > 0 && nf > 0, "wcompute_fwd: bad shape B=%d N=%d F=%d nf=%d", B, N, F, nf);
    MFT_REQUIRE(2 * nf <= kMaxC, "wcompute_fwd: nf=%d exceeds the supported maximum %d", nf, kMaxC / 2);
    MFT_REQUIRE(N < 32768, "wcompute_fwd: N=%d too large", N);
    MFT_REQUIRE(ldx >= F, "wcompute_fwd: ldx=%d < F=%d", ldx, F);
    WcLayout L = wc_layout(B, N, F, nf, saved, workspace, precision);
    NodeMask mask;
    const int n_shared = mask_from_host(shared_nodes, B, N, mask);
    PairGeom g = make_geom(B, N, L.tri, L.inv, n_shared, L.roww, L.rowij);

    if (!prepared) {
        int rc = wcompute_fwd_prepare(B, N, F, nf, p, saved, workspace, precision, shared_nodes, st);
        if (rc != MFT_OK) return rc;
    }

    if (precision == MFT_PREC_TF32) {
        int rc = wcompute_fwd_layers_tf32(x, ldx, F, nf, p, L, g, st);
        if (rc != MFT_OK) return rc;
    } else {
        for (int k = 0; k < 4; ++k) {
            double* sums = L.fsums + (size_t)k * kStatSlot;
            EpiFwdStats epi{L.H[k], L.C[k + 1], sums, g};
            WView wv = wview_nt(p->conv_w[k], L.C[k]);
            ProfScope ps(PC_FWD_L1 + k, st);
            if (k == 0) {
                AbsDiffA a{x, ldx, g};
                MFT_CHECK_CUDA((launch_gemm_rows<true>(a, wv, epi, g.R, L.C[1], F, st)));
            } else {
                const double* ps = L.fsums + (size_t)(k - 1) * kStatSlot;
                BnActA a{L.H[k - 1], L.C[k], ps, p->bn_g[k - 1], p->bn_b[k - 1], g.inv_pairs};
                MFT_CHECK_CUDA((launch_gemm_rows<true>(a, wv, epi, g.R, L.C[k + 1], L.C[k], st)));
            }
        }
    }
    if (mid) mid->fork(mid_slot);
    {
        ProfScope ps(PC_SCORE, st);
#define MFT_SCORE(HALF, SHARED)                                                                                   \
    do {                                                                                                          \
        if (nf % 4 == 0 && nf <= 96)                                                                              \
            MFT_CHECK_CUDA(launch_kernel(score8_kernel<HALF, SHARED, 3>, dim3(row_grid8(g.R)),                    \
                dim3(kRowWarps * 32), 0, st, row_pdl, (const void*)L.H[3], nf,                                    \
                (const double*)(L.fsums + 3 * kStatSlot), p->bn_g[3], p->bn_b[3], p->last_w, p->last_b, g, L.S));  \
        else if (nf % 4 == 0 && nf <= 128)                                                                        \
            score_vec_kernel<HALF, SHARED, 1><<<row_grid(g.R), kRowWarps * 32, 0, st>>>(                           \
                L.H[3], nf, L.fsums + 3 * kStatSlot, p->bn_g[3], p->bn_b[3], p->last_w, p->last_b, g, L.S);        \
        else                                                                                                      \
            score_kernel<HALF, SHARED><<<row_grid(g.R), kRowWarps * 32, 0, st>>>(                                  \
                L.H[3], nf, L.fsums + 3 * kStatSlot, p->bn_g[3], p->bn_b[3], p->last_w, p->last_b, g, L.S);        \
    } while (0)
        // follows the layer-4 GEMM launch on the tensor-core path (wait-then-release convention)
        const bool row_pdl = precision == MFT_PREC_TF32 && pdl_level() >= 2;
        if (precision == MFT_PREC_TF32) {
            if (g.Rs > 0) MFT_SCORE(true, true); else MFT_SCORE(true, false);
        } else {
            if (g.Rs > 0) MFT_SCORE(false, true); else MFT_SCORE(false, false);
        }
#undef MFT_SCORE
        MFT_CHECK_LAUNCH();
    }
    {
        ProfScope ps(PC_SOFTMAX, st);
        // only score8_kernel is known to release its dependents after its own wait
        const bool pdl = precision == MFT_PREC_TF32 && pdl_level() >= 2 && nf % 4 == 0 && nf <= 96;
        MFT_CHECK_CUDA(launch_kernel(softmax_rows_kernel, dim3(cdiv(B * N, kRowWarps)), dim3(kRowWarps * 32), 0, st,
                                     pdl, (const float*)L.S, adj, B * N, N));
    }
    return MFT_OK;
}

// Parameter-only part of the backward: pair tables and (tensor-core path) the four dgrad weight images.  Touches
// the tables and the image region of the workspace only (not the reductions / partial copies a previous layer's
// gradient finalisation may still be reading).
int wcompute_bwd_prepare(int B, int N, int F, int nf, const mft_wcompute_params* p, void* saved, void* workspace,
                         int precision, const unsigned char* shared_nodes, cudaStream_t st) {
    WcLayout L = wc_layout(B, N, F, nf, saved, workspace, precision);
    NodeMask mask;
    const int n_shared = mask_from_host(shared_nodes, B, N, mask);
    PairGeom g = make_geom(B, N, L.tri, L.inv, n_shared, L.roww, L.rowij);
    {
        ProfScope ps(PC_PREP, st);
        tri_table_kernel<<<cdiv(g.Rg, 256), 256, 0, st>>>(L.tri, L.inv, L.roww, L.rowij, B, N, g.Rg, g.Rs, n_shared, mask);
        MFT_CHECK_LAUNCH();
    }
    if (precision == MFT_PREC_TF32) return wcompute_bwd_prepare_tf32(p, L, F, nf, st);   // four images, one launch
    return MFT_OK;
}

int wcompute_bwd(const float* x, int ldx, int B, int N, int F, int nf, const mft_wcompute_params* p,
                 const float* adj, const float* d_adj, float* dx, const mft_wcompute_grads* gr, void* saved,
                 void* workspace, int precision, const unsigned char* shared_nodes, cudaStream_t st,
                 Branches* tail, int tail_slot, bool prepared) {
    MFT_REQUIRE(B > 0 && N > 0 && F > 0 && nf > 0, "wcompute_bwd: bad shape");
    MFT_REQUIRE(2 * nf <= kMaxC, "wcompute_bwd: nf=%d exceeds the supported maximum %d", nf, kMaxC / 2);
    WcLayout L = wc_layout(B, N, F, nf, saved, workspace, precision);
    NodeMask mask;
    const int n_shared = mask_from_host(shared_nodes, B, N, mask);
    PairGeom g = make_geom(B, N, L.tri, L.inv, n_shared, L.roww, L.rowij);

    int wg_copies[4] = {0, 0, 0, 0};
    if (precision != MFT_PREC_TF32) {
        for (int k = 0; k < 4; ++k)
            MFT_CHECK_CUDA(cudaMemsetAsync(gr->conv_w[k], 0, sizeof(float) * (size_t)L.C[k + 1] * L.C[k], st));
    }
    {
        // the pair tables and (tensor-core path) the four dgrad weight images do not depend on the
        // upstream gradient: build them on a side branch while the softmax backward runs -- unless the caller
        // has already had them built (wcompute_bwd_prepare; gnn_bwd does it beside the Gconv backward)
        Branches br(st);
        cudaStream_t s0 = br.fork(0);
        // (the reductions of the backward are first touched by the dy4 kernel: cleared beside the softmax backward)
        MFT_CHECK_CUDA(cudaMemsetAsync(L.bsums, 0, sizeof(double) * 5 * kStatSlot, s0));
        if (!prepared) {
            int rc = wcompute_bwd_prepare(B, N, F, nf, p, saved, workspace, precision, shared_nodes, s0);
            if (rc != MFT_OK) return rc;
        }
        {
            ProfScope ps(PC_SOFTMAX_BWD, st);
            softmax_bwd_kernel<<<cdiv(B * N, kRowWarps), kRowWarps * 32, 0, st>>>(adj, d_adj, L.S, B * N, N);
            MFT_CHECK_LAUNCH();
        }
        br.join(0);
        MFT_REQUIRE(br.ok(), "wcompute_bwd: stream fork/join failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    double* lastsum = L.bsums + 4 * kStatSlot;
    {
        ProfScope ps(PC_DY4, st);
#define MFT_DY4(HALF, SHARED)                                                                                     \
    do {                                                                                                          \
        if (nf % 4 == 0 && nf <= 96)                                                                              \
            dy8_kernel<HALF, SHARED, 3><<<row_grid8(g.R), kRowWarps * 32, 0, st>>>(                                \
                L.S, L.H[3], nf, L.fsums + 3 * kStatSlot, p->bn_g[3], p->bn_b[3], p->last_w, g, L.dyA,             \
                L.bsums + 3 * kStatSlot, lastsum);                                                                 \
        else if (nf % 4 == 0 && nf <= 128)                                                                        \
            dy4_vec_kernel<HALF, SHARED, 1><<<row_grid(g.R), kRowWarps * 32, 0, st>>>(                             \
                L.S, L.H[3], nf, L.fsums + 3 * kStatSlot, p->bn_g[3], p->bn_b[3], p->last_w, g, L.dyA,             \
                L.bsums + 3 * kStatSlot, lastsum);                                                                 \
        else                                                                                                      \
            dy4_kernel<HALF, SHARED><<<row_grid(g.R), kRowWarps * 32, 0, st>>>(                                    \
                L.S, L.H[3], nf, L.fsums + 3 * kStatSlot, p->bn_g[3], p->bn_b[3], p->last_w, g, L.dyA,             \
                L.bsums + 3 * kStatSlot, lastsum);                                                                 \
    } while (0)
        if (precision == MFT_PREC_TF32) {
            if (g.Rs > 0) MFT_DY4(true, true); else MFT_DY4(true, false);
        } else {
            if (g.Rs > 0) MFT_DY4(false, true); else MFT_DY4(false, false);
        }
#undef MFT_DY4
        MFT_CHECK_LAUNCH();
    }

    // Tensor-core path: wgrad_k and dgrad_k both consume dy_k and nothing of each other, and a persistent
    // GEMM pays a fixed fill + drain (one tile's producer time before, one tile's epilogue after: a third of
    // a 4.7-tiles-per-CTA launch).  So the wgrad of a layer runs on a side stream beside the dgrad chain, each
    // kernel on part of the SMs: twice the tiles per CTA amortise the fixed part, one kernel's fill hides
    // under the other's steady state, and both read dy_k / H_k at about the same time (L2).  Three dy
    // buffers rotate so that dgrad_{k-1} never writes what wgrad_k is still reading; wgrads alternate between
    // two side streams and dgrad_{k-2} joins the one wgrad_k ran on.  MFT_BWD_SPLIT=0 restores the serial order.
    static const int split_env = [] { const char* e = getenv("MFT_BWD_SPLIT"); return e ? atoi(e) : 1; }();
    static const int small_env = [] { const char* e = getenv("MFT_BWD_SPLIT_SMALL"); return e ? atoi(e) : 1; }();
    const int sms = umma_num_sms();
    const int tiles = (g.R + 127) / 128;             // CTAs a dgrad launch can use
    const bool one_wave = small_env != 0 && tiles <= sms - 40;
    const bool split = precision == MFT_PREC_TF32 && split_env != 0 && (g.R >= 16 * 128 * 4 || one_wave);
    // SMs of the wgrad kernels (the dgrad chain gets the rest): MFT_BWD_SPLIT=<count> overrides.  Many tiles
    // (measured on B200, 5w20s / 5w50c): 40 % of the SMs for the wgrads is the optimum -- the dgrad chain is the
    // critical path, the wgrads only have to keep up (profiles/r02_summary.md).  A dgrad launch that fits in one wave
    // (5w5s: 58 tiles; the shared-support layer_w0 of 5w20s: 104) keeps one SM per tile and the wgrads get the rest.
    const int wg_sms = split_env >= 8 ? min(split_env, sms - 8) : (one_wave ? sms - tiles : (sms * 2) / 5);
    ProfScope* region = precision == MFT_PREC_TF32 ? new ProfScope(PC_BWD_REGION, st, false) : nullptr;
    Branches wb(st);
    float* bufs[3] = {L.dyA, L.dyB, L.dyC};
    int bi = 0;
    float* cur = L.dyA;
    float* nxt = L.dyB;
    for (int k = 3; k >= 0; --k) {   // layer k+1 of the reference (conv2d_{k+1}, bn_{k+1})
        const int Cout = L.C[k + 1], Cin = L.C[k];
        const double* fs = L.fsums + (size_t)k * kStatSlot;
        const double* bs = L.bsums + (size_t)k * kStatSlot;
        if (precision != MFT_PREC_TF32) {   // the tensor-core path applies BN-backward inside its operand producers
            ProfScope ps(PC_DH, st);
            dh_kernel<<<row_grid(g.R), kRowWarps * 32, 0, st>>>(cur, L.H[k], Cout, fs, p->bn_g[k], bs, g);
            MFT_CHECK_LAUNCH();
        }
        {
            PlainA dh{cur, Cout};
            // wgrad: d conv2d_{k+1}.weight [Cout, Cin] = dH^T a_k
            if (precision == MFT_PREC_TF32) {
                cudaStream_t wst = st;
                if (split) {
                    const int slot = k & 1;
                    wb.join(slot);                    // dy buffer (k+2) % 3 .. is about to be rewritten by this level's dgrad
                    wst = wb.fork(slot);              // ordered after the kernel that produced dy_k
                    umma_set_grid_limit(wg_sms);
                }
                int rc = wcompute_wgrad_layer_tf32(k, cur, x, ldx, F, p, gr, L, g, &wg_copies[k], wst);
                if (split) umma_set_grid_limit(umma_num_sms() - wg_sms);
                if (rc != MFT_OK) { umma_set_grid_limit(0); delete region; return rc; }
            } else if (k == 0) {
                ProfScope ps(PC_WGRAD_L1, st);
                AbsDiffA q{x, ldx, g};
                MFT_CHECK_CUDA((launch_gemm_tn(dh, q, gr->conv_w[0], Cin, Cout, Cin, g.R, st)));
            } else {
                const double* ps = L.fsums + (size_t)(k - 1) * kStatSlot;
                BnActA q{L.H[k - 1], Cin, ps, p->bn_g[k - 1], p->bn_b[k - 1], g.inv_pairs};
                ProfScope psc(PC_WGRAD_L1 + k, st);
                MFT_CHECK_CUDA((launch_gemm_tn(dh, q, gr->conv_w[k], Cin, Cout, Cin, g.R, st)));
            }
            // dgrad: dL/d a_k = dH W
            if (precision == MFT_PREC_TF32) {
                if (split) nxt = bufs[(bi + 1) % 3];
                int rc = wcompute_bwd_layer_tf32(k, cur, nxt, x, ldx, dx, F, nf, p, gr, L, g, st);
                umma_set_grid_limit(0);
                if (rc != MFT_OK) { delete region; return rc; }
            } else {
                WView wv = wview_nn(p->conv_w[k], Cin);
                ProfScope psd(PC_DGRAD_L1 + k, st);
                if (k == 0) {
                    EpiDx epi{x, dx, ldx, g};
                    MFT_CHECK_CUDA((launch_gemm_rows<false>(dh, wv, epi, g.R, Cin, Cout, st)));
                } else {
                    const double* ps = L.fsums + (size_t)(k - 1) * kStatSlot;
                    double* pbs = L.bsums + (size_t)(k - 1) * kStatSlot;
                    EpiDy epi{L.H[k - 1], nxt, Cin, ps, p->bn_g[k - 1], p->bn_b[k - 1], g.inv_pairs, pbs};
                    MFT_CHECK_CUDA((launch_gemm_rows<false>(dh, wv, epi, g.R, Cin, Cout, st)));
                }
            }
        }
        if (split) { bi = (bi + 1) % 3; cur = bufs[bi]; }
        else { float* tmp = cur; cur = nxt; nxt = tmp; }
    }
    wb.join(0);
    wb.join(1);
    delete region;
    MFT_REQUIRE(wb.ok(), "wcompute_bwd: stream fork/join failed: %s", cudaGetErrorString(cudaGetLastError()));

    FinalizeArgs fa;
    for (int k = 0; k < 4; ++k) {
        fa.bsums[k] = L.bsums + (size_t)k * kStatSlot;
        fa.bn_g[k] = gr->bn_g[k];
        fa.bn_b[k] = gr->bn_b[k];
        fa.conv_b[k] = gr->conv_b[k];
        fa.C[k] = L.C[k + 1];
        fa.wgpart[k] = precision == MFT_PREC_TF32 ? L.wgpart + L.wgpart_off[k] : nullptr;
        fa.conv_w[k] = gr->conv_w[k];
        fa.wsize[k] = L.C[k + 1] * L.C[k];
        fa.cin[k] = L.C[k];
        fa.ncopies[k] = wg_copies[k];
    }
    fa.lastsum = lastsum;
    fa.last_w = gr->last_w;
    fa.last_b = gr->last_b;
    fa.nf = nf;
    fa.tscale = precision == MFT_PREC_TF32 ? L.tscale : nullptr;
    {
        // These only finish parameter gradients: on the caller's tail branch (if any) they run beside
        // whatever the main stream does next.  On the main stream of the tensor-core path they follow the
        // dx gather launched by wcompute_bwd_layer_tf32(0, ...) (programmatic launch allowed).
        cudaStream_t fs = tail ? tail->fork(tail_slot) : st;
        const bool pdl = fs == st && precision == MFT_PREC_TF32 && pdl_level() >= 2;
        ProfScope ps(PC_FINALIZE, fs);
        MFT_CHECK_CUDA(launch_kernel(finalize_grads_kernel, dim3(5), dim3(kMaxC), 0, fs, pdl, fa));
        if (precision == MFT_PREC_TF32) {
            MFT_CHECK_CUDA(launch_kernel(wgrad_reduce_kernel, dim3(L.C[1] + L.C[2] + L.C[3] + L.C[4]),
                                         dim3(kRedThreads), 0, fs, pdl, fa));
        }
    }
    return MFT_OK;
}

}  // namespace mft
