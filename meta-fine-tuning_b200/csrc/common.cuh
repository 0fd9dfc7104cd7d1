// Shared definitions of libmft_gnn: error plumbing, the unordered-pair row
// geometry, batch-statistic helpers.  See DESIGN.md for the data layout.
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stddef.h>

#include "mft_gnn.h"

namespace mft {

constexpr float kBnEps = 1e-5f;     // reference gnn.py:65 (torch BatchNorm default)
constexpr float kSlope = 0.01f;     // F.leaky_relu default, gnn.py:86
constexpr float kDiagMask = 1e8f;   // gnn.py:106
constexpr int kMaxC = 256;          // widest channel count a BN'd layer may have (2*nf)

void set_error(int code, const char* fmt, ...);
int last_code();

#define MFT_CHECK_CUDA(expr)                                                          \
    do {                                                                               \
        cudaError_t _e = (expr);                                                       \
        if (_e != cudaSuccess) {                                                       \
            mft::set_error(MFT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,               \
                           cudaGetErrorString(_e), __FILE__, __LINE__);                \
            return MFT_ERR_CUDA;                                                       \
        }                                                                              \
    } while (0)

#define MFT_CHECK_LAUNCH()  MFT_CHECK_CUDA(cudaGetLastError())

#define MFT_REQUIRE(cond, ...)                                                        \
    do {                                                                               \
        if (!(cond)) {                                                                 \
            mft::set_error(MFT_ERR_ARG, __VA_ARGS__);                                  \
            return MFT_ERR_ARG;                                                        \
        }                                                                              \
    } while (0)

// ---------------------------------------------------------------------------
// Programmatic dependent launch (griddepcontrol): a kernel launched with
// cudaLaunchAttributeProgrammaticStreamSerialization may start while its
// predecessor on the stream is still draining; pdl_wait() blocks until that
// predecessor has completed and its writes are visible, pdl_launch_dependents()
// lets the NEXT launch of the stream start early in turn.  Both are no-ops in a
// kernel launched the ordinary way.  Convention of this library: wait first,
// release dependents second, touch upstream data third -- so completion of a
// kernel always implies completion of everything before it on the stream.
// ---------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_wait(); pdl_launch_dependents(); }
#endif

// 0 = plain stream order, 1 = programmatic dependent launch inside the edge-MLP GEMM chains,
// 2 = also for the row kernels around them (default; MFT_PDL in the environment, mft_set_pdl()).
// Measured on B200 (profiles/r01_summary.md): about 1 % at 5w20s and 4 % at 5w5s -- every layer is a
// grid-wide BatchNorm dependency and the persistent CTAs own whole SMs, so only the launch latency and
// the set-up of the next kernel overlap.
int pdl_level();
bool prof_enabled();

// Kernel launch with the programmatic-serialization attribute when `pdl` (and never while
// per-launch event timing is on: the events would sit between the two kernels anyway).
template <class... KArgs, class... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                 bool pdl, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (pdl && !prof_enabled()) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}

// ---------------------------------------------------------------------------
// Pair-row geometry.  One graph has N nodes; the edge MLP runs on the Rg =
// N(N+1)/2 unordered pairs (i<=j) of each of the B graphs, row-major in i:
//   r_local = i*N - i*(i-1)/2 + (j-i),   r = b*Rg + r_local.
// Off-diagonal rows stand for the two ordered pairs (i,j),(j,i) of the
// reference's dense [B,N,N] tensor and carry multiplicity 2 into the batch
// statistics; `pairs` = B*N*N is the reference's BatchNorm population.
// ---------------------------------------------------------------------------
struct PairGeom {
    int B, N, Rg, R;
    int Rs, Rq;       // shared rows (one for all B graphs) / per-graph rows; Rs = 0, Rq = Rg without sharing
    double inv_pairs;
    const int* tri;   // [Rg] packed (j << 16) | i: the Rs shared pairs first, then the Rq per-graph ones
    const int* inv;   // [N*N] (i <= j): index into the per-graph part, or -1 - index into the shared part
    // Per-ROW tables (tri_table_kernel) for the tcgen05 kernels, whose roles must not stall on a
    // division plus a dependent table load per row: a pure load each, consumed a tile later.
    const float* roww;  // [R + 64] multiplicity w of row r (see PairRow); roww[R...] = 0 stands for "past the end"
    const int2* rowij;  // [R + 1] node-matrix rows (b*N + i, b*N + j) of row r (b = 0 for a shared row)
};

// Shared nodes (include/mft_gnn.h, `shared_nodes`): node n is "shared" when x[b, n, :] is the same
// row for every graph b -- the support nodes of GnnNet's graphs (gnnnet.py:79-80 replicates them
// into each query's graph).  A pair of two shared nodes has the same edge features in all B graphs,
// so it gets ONE row standing for B (diagonal) or 2B (off-diagonal) ordered pairs of the reference's
// dense tensor.  The mask travels by value (kernel parameter), one bit per node.
constexpr int kMaxMaskNodes = 1024;
struct NodeMask {
    uint32_t w[kMaxMaskNodes / 32];
};

inline int mask_from_host(const unsigned char* shared_nodes, int B, int N, NodeMask& m) {
    for (int q = 0; q < kMaxMaskNodes / 32; ++q) m.w[q] = 0u;
    if (shared_nodes == nullptr || B < 2 || N > kMaxMaskNodes) return 0;
    int S = 0;
    for (int n = 0; n < N; ++n)
        if (shared_nodes[n]) {
            m.w[n >> 5] |= 1u << (n & 31);
            ++S;
        }
    return S;
}

inline PairGeom make_geom(int B, int N, const int* tri, const int* inv = nullptr, int n_shared = 0,
                          const float* roww = nullptr, const int2* rowij = nullptr) {
    PairGeom g;
    g.B = B;
    g.N = N;
    g.Rg = N * (N + 1) / 2;
    g.Rs = n_shared * (n_shared + 1) / 2;
    g.Rq = g.Rg - g.Rs;
    g.R = g.Rs + B * g.Rq;
    g.inv_pairs = 1.0 / ((double)B * (double)N * (double)N);
    g.tri = tri;
    g.inv = inv;
    g.roww = roww;
    g.rowij = rowij;
    return g;
}

__host__ __device__ inline int tri_start(int i, int N) { return i * N - (i * (i - 1)) / 2; }

__device__ __forceinline__ void decode_local(int rl, int N, int& i, int& j) {
    float fn = (float)(2 * N + 1);
    int ii = (int)((fn - sqrtf(fn * fn - 8.0f * (float)rl)) * 0.5f);
    ii = max(0, min(ii, N - 1));
    while (tri_start(ii, N) > rl) --ii;
    while (ii + 1 < N && tri_start(ii + 1, N) <= rl) ++ii;
    i = ii;
    j = ii + (rl - tri_start(ii, N));
}

struct PairRow {
    int b, i, j;
    int nb;    // graphs the row stands for: 1, or B for a shared pair (then b = 0)
    float w;   // multiplicity in the reference's dense tensor: nb (diagonal) or 2*nb; 0 past the end
};

__device__ __forceinline__ PairRow decode_row(int r, const PairGeom& g) {
    PairRow p;
    if (r >= g.R) {
        p.b = 0; p.i = 0; p.j = 0; p.nb = 0; p.w = 0.f;
        return p;
    }
    int t;
    if (r < g.Rs) {
        p.b = 0; p.nb = g.B; t = r;
    } else {
        int rr = r - g.Rs;
        p.b = rr / g.Rq; p.nb = 1; t = g.Rs + (rr - p.b * g.Rq);
    }
    int packed = __ldg(g.tri + t);
    p.i = packed & 0xffff;
    p.j = packed >> 16;
    p.w = (float)((p.i == p.j) ? p.nb : 2 * p.nb);
    return p;
}

// Row of the unordered pair {i <= j} of graph b; `shared` tells whether it is a shared row.
__device__ __forceinline__ int pair_row(const PairGeom& g, int b, int i, int j, bool& shared) {
    if (g.Rs == 0) {
        shared = false;
        return b * g.Rg + tri_start(i, g.N) + (j - i);
    }
    int q = __ldg(g.inv + i * g.N + j);
    shared = q < 0;
    return shared ? (-1 - q) : (g.Rs + b * g.Rq + q);
}

__device__ __forceinline__ float lrelu(float y) { return y > 0.f ? y : y * kSlope; }
__device__ __forceinline__ float dlrelu(float y) { return y > 0.f ? 1.f : kSlope; }

// Batch statistics of one BN layer live in a "slot": kStatCopies copies of [2][C] doubles
// (which = 0: sum w*h or sum dy, which = 1: sum w*h^2 or sum dy*hhat).  CTAs accumulate with fp64
// atomics into copy (blockIdx.x % kStatCopies): same-address atomics serialise in L2 (~200 cycles
// each, measured: 148 CTAs on one copy cost 17 us per layer), a few copies remove that.  Consumers add
// the copies up (stat_get) and turn them into mean / rstd themselves -- in EVERY kernel's prologue, on the critical
// path of every launch, which is why more is not better (5w20s step on B200, tools/ab_build2.sh: 2 / 4 / 6 copies
// 1.251 ms, 8: 1.255, 16: 1.272, 32: 1.321).
#ifndef MFT_STAT_COPIES
#define MFT_STAT_COPIES 4
#endif
constexpr int kStatCopies = MFT_STAT_COPIES;
constexpr int kStatCopyStride = 2 * kMaxC;
constexpr int kStatSlot = kStatCopies * kStatCopyStride;   // doubles per slot

__device__ __forceinline__ void stat_add(double* slot, int C, int c, int which, float v) {
    atomicAdd(slot + (blockIdx.x % kStatCopies) * kStatCopyStride + which * C + c, (double)v);
}
__device__ __forceinline__ double stat_get(const double* slot, int C, int c, int which) {
    double a = 0.0;
#pragma unroll
    for (int q = 0; q < kStatCopies; ++q) a += slot[q * kStatCopyStride + which * C + c];
    return a;
}

__device__ __forceinline__ void bn_mean_rstd(const double* sums, int C, int c, double inv_count,
                                             float& mean, float& rstd) {
    double m = stat_get(sums, C, c, 0) * inv_count;
    // The second-moment sum of a tensor-core-path slot carries count * (s^2 - 1) * eps (s = the layer's
    // power-of-two tape scale, umma_layer_scales_kernel), so that "variance + eps" below is
    // var(s h) + s^2 eps and the normalisation is exactly BatchNorm(h; eps) whatever s is.  Hence the
    // clamp applies to the sum, not to the variance alone.
    double v = stat_get(sums, C, c, 1) * inv_count - m * m + (double)kBnEps;
    if (v < 1e-300) v = 1e-300;
    mean = (float)m;
    rstd = (float)(1.0 / sqrt(v));
}

// Per-channel constants staged in shared memory by every kernel that applies a BN.
struct BnSmem {
    float* mean;
    float* rstd;
    float* gamma;
    float* beta;
};

__device__ __forceinline__ BnSmem bn_smem_at(float* base) {
    BnSmem s;
    s.mean = base;
    s.rstd = base + kMaxC;
    s.gamma = base + 2 * kMaxC;
    s.beta = base + 3 * kMaxC;
    return s;
}

__device__ __forceinline__ void bn_smem_fill(BnSmem s, const double* sums, const float* gamma,
                                             const float* beta, int C, double inv_count) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float m, r;
        bn_mean_rstd(sums, C, c, inv_count, m, r);
        s.mean[c] = m;
        s.rstd[c] = r;
        s.gamma[c] = gamma ? gamma[c] : 1.f;
        s.beta[c] = beta ? beta[c] : 0.f;
    }
}

// ---------------------------------------------------------------------------
// Activation tape element types.  The fp32 path keeps everything in float.  The tensor-core path
// stores the pre-BN activations H as fp16 (10-bit mantissa: the precision the TF32 operand
// rounding applies to them anyway; values are clamped to the fp16 range on store) and the
// gradients dy as bf16 (fp32 range -- gradients span many decades -- 8-bit mantissa).  Halving the
// tape halves the HBM traffic of kernels that are all HBM bound.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float4 unpack_half4(uint2 u) {
    float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
    float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ uint2 pack_half4(float4 v) {
    uint2 u;   // round to nearest, saturate to +-65504 instead of overflowing to inf
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(u.x) : "f"(v.y), "f"(v.x));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(u.y) : "f"(v.w), "f"(v.z));
    return u;
}
__device__ __forceinline__ float4 unpack_bf4(uint2 u) {
    float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ uint2 pack_bf4(float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
    __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    return u;
}
// scalar tape accessors for the row-per-warp kernels shared by both paths
template <bool Half> struct TapeH;
template <> struct TapeH<false> {
    typedef float T;
    static __device__ __forceinline__ float ld(const void* p, size_t i) { return __ldg(static_cast<const float*>(p) + i); }
};
template <> struct TapeH<true> {
    typedef __half T;
    static __device__ __forceinline__ float ld(const void* p, size_t i) {
        return __half2float(__ldg(static_cast<const __half*>(p) + i));
    }
};
template <bool Half> struct TapeD;
template <> struct TapeD<false> {
    static __device__ __forceinline__ void st(void* p, size_t i, float v) { static_cast<float*>(p)[i] = v; }
};
template <> struct TapeD<true> {
    static __device__ __forceinline__ void st(void* p, size_t i, float v) {
        static_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
    }
};

// Packed fp32 pairs (Blackwell FFMA2 / FMUL2 / FADD2): one issue slot for two lanes of arithmetic.
// The hot loops of the tensor-core kernels are bound by instruction issue of one warp per scheduler.
struct f2 { unsigned long long v; };
__device__ __forceinline__ f2 pack2(float lo, float hi) {
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f2 a, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
    f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
    f2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
    f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Bump carving of the caller-provided blobs (saved / workspace).
struct Carver {
    char* base;
    size_t off;
    explicit Carver(void* p) : base(static_cast<char*>(p)), off(0) {}
    template <typename T>
    T* take(size_t count) {
        off = (off + 255) & ~size_t(255);
        T* r = reinterpret_cast<T*>(base + off);
        off += count * sizeof(T);
        return r;
    }
    size_t used() const { return (off + 255) & ~size_t(255); }
};

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

}  // namespace mft
