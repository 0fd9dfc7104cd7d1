// tcgen05 (MFT_PREC_TF32) edge-MLP layers.
//
// One persistent, warp-specialised kernel template covers every "pair rows x weights" GEMM of
// the edge MLP (the four forward 1x1-conv layers and the four dgrad products):
//
//     C[r, n] = sum_k A(r, k) * W(n, k)        r: 128-row tiles of unordered pairs
//
//   warps 0-7   producers : build the A tile -- |x_i - x_j| from the L2-resident node matrix for layer 1,
//                           LeakyReLU(BN(H)) from the fp16 tape (landed by TMA in a raw ring) for later
//                           layers, the fused BatchNorm-backward dH (in place on the landed dy block) for
//                           dgrad -- round it to TF32 and write it K-major / SWIZZLE_128B into a ring of
//                           16 KB K blocks, so neither the N^2 x C pair tensor nor any post-BN activation
//                           ever exists in HBM;
//   warps 8-11  epilogue  : tcgen05.ld 32 columns at a time -> padded smem staging -> coalesced row
//                           segments to global, fused with the per-channel batch statistics (forward) or
//                           with the LeakyReLU'/BN-backward reductions (dgrad);
//   warp  12    MMA issuer: one thread; weights (pre-swizzled TF32 image) land once per CTA by bulk async
//                           copy and stay resident; tcgen05.mma kind::tf32, M=128, N = tile width,
//                           accumulators double-buffered in TMEM;
//   warp  13    TMA loader: one thread; tensor-map tile loads of the tape / gradient operands.
//
// Pipelines: rawfull/rawempty mbarriers per raw slot (TMA <-> producers; a slot is released only when the
// loads from it have COMPLETED, see the producers), full/empty per A stage (producers <-> MMA),
// tmem_full/tmem_empty per accumulator (MMA <-> epilogue).  Every wait is bounded (umma::mbar_wait traps
// on timeout).
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cuda.h>   // CUtensorMap types only; the encoder is fetched through cudaGetDriverEntryPoint

#include "common.cuh"
#include "wcompute.cuh"
#include "umma.cuh"
#include "prof.cuh"

namespace mft {

using namespace umma;

constexpr int UM_ROWS = 128;
constexpr int UM_KB = 32;                 // tf32 elements per K block (128 bytes)
constexpr int UM_BLOCK_FLOATS = UM_ROWS * UM_KB;   // one A stage = 4096 floats = 16 KB
constexpr int UM_PROD_WARPS = 8;
constexpr int UM_PROD_THREADS = UM_PROD_WARPS * 32;
constexpr int UM_EPI_WARP0 = UM_PROD_WARPS;          // epilogue warps 8..11 (warp % 4 = TMEM lane quadrant)
constexpr int UM_MMA_WARP = UM_PROD_WARPS + 4;
constexpr int UM_TMA_WARP = UM_MMA_WARP + 1;           // issues the tensor-map tile loads of kTma operands
constexpr int UM_THREADS = (UM_PROD_WARPS + 6) * 32;  // 8 producer + 4 epilogue + MMA + TMA warp = 448
constexpr int UM_STAGE_LD = 36;           // padded row length of the epilogue staging tile
constexpr int UM_ACC_STRIDE = 256;        // TMEM columns between the two accumulators
constexpr int UM_TMEM_COLS = 512;
constexpr int UM_MAX_STAGES = 4;
constexpr int UM_MAX_PSTAGES = 8;          // A ring depth of the CTA-pair kernel
constexpr int UM_MAX_RAW = 8;              // raw fp16 K blocks a TMA operand may have in flight
constexpr int UM_RAW_BYTES = UM_ROWS * UM_KB * 2;   // [128 rows x 32 ch] fp16 = 8 KB
constexpr unsigned WG_ROWS_C = 32;        // rows per wgrad K chunk (declared early for the TMA functors)
constexpr int UM_MAX_CHUNKS = 8;          // 32-column epilogue chunks (N_TILE <= 256)
constexpr int UM_STAT_CHUNKS = 6;         // chunks that can carry column statistics (N_TILE <= 192)
constexpr int UM_MAX_NTILE = 240;
constexpr int UM_MAX_KC = 8;

struct UmmaShape {
    int R;        // valid rows
    int N;        // valid output columns (global)
    int n0;       // first output column of this pass
    int N_TILE;   // MMA N of this pass (multiple of 16)
    int K;        // valid K
    int KC;       // K blocks (ceil(K/32))
    int stages;   // A ring depth
    int raw_stages;   // fp16 raw-block ring depth (TMA operands only, else 0)
    long long* dbg;   // optional [gridDim.x][16] clock64 timeline (debug builds of the tests only)
    int reverse;      // walk the row tiles from the last to the first (see next_direction())
    unsigned zero;    // always 0; unknown to the compiler (see the raw-ring release in the producers)
};

// CTA-pair kernel: half of the weight rows per CTA, a longer barrier block
static inline size_t umma_smem_bytes_pair(const UmmaShape& s) {
    return 1024 + (size_t)s.KC * (s.N_TILE / 2) * 128 + (size_t)s.stages * UM_BLOCK_FLOATS * 4 +
           (size_t)UM_ROWS * UM_STAGE_LD * 4 + (3 + 4) * kMaxC * 4 + (size_t)s.raw_stages * UM_RAW_BYTES + 48 * 8 + 16;
}

static inline size_t umma_smem_bytes(const UmmaShape& s) {
    return 1024 + (size_t)s.KC * s.N_TILE * 128 + (size_t)s.stages * UM_BLOCK_FLOATS * 4 +
           (size_t)UM_ROWS * UM_STAGE_LD * 4 + (3 + 4) * kMaxC * 4 + (size_t)s.raw_stages * UM_RAW_BYTES + 32 * 8 + 16;
}

// ------------------------------------------------------------------ weight image
// img[kc][n][32] (floats), K-major SWIZZLE_128B, TF32-rounded, zero padded:
// element (n, k) of the operand = transpose ? W[k*ldw + n0+n] : W[(n0+n)*ldw + k].
__global__ void umma_weight_image_kernel(const float* __restrict__ W, int ldw, int transpose, int N, int K,
                                         int n0, int N_TILE, int KC, float* __restrict__ img) {
    int total = KC * N_TILE * UM_KB;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        int kc = idx / (N_TILE * UM_KB);
        int rem = idx - kc * N_TILE * UM_KB;
        int n = rem / UM_KB, kk = rem - n * UM_KB;
        int k = kc * UM_KB + kk;
        float v = 0.f;
        if (n0 + n < N && k < K) v = transpose ? W[(size_t)k * ldw + n0 + n] : W[(size_t)(n0 + n) * ldw + k];
        img[(size_t)kc * N_TILE * UM_KB + sw128_offset(n, kk)] = to_tf32(v);
    }
}

// All weight images of one Wcompute direction in ONE launch (blockIdx.y = job).
struct ImgJob { const float* W; float* img; const float* scale; int ldw, transpose, N, K, n0, N_TILE, KC; };
struct ImgJobs { int n; ImgJob j[12]; };

__global__ void umma_weight_images_kernel(ImgJobs jobs) {
    const ImgJob& jb = jobs.j[blockIdx.y];
    const int total = jb.KC * jb.N_TILE * UM_KB;
    const float sc = jb.scale ? __ldg(jb.scale) : 1.f;      // the layer's tape scale (a power of two: exact)
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        int kc = idx / (jb.N_TILE * UM_KB);
        int rem = idx - kc * jb.N_TILE * UM_KB;
        int n = rem / UM_KB, kk = rem - n * UM_KB;
        int k = kc * UM_KB + kk;
        float v = 0.f;
        if (jb.n0 + n < jb.N && k < jb.K)
            v = jb.transpose ? jb.W[(size_t)k * jb.ldw + jb.n0 + n] : jb.W[(size_t)(jb.n0 + n) * jb.ldw + k];
        jb.img[(size_t)kc * jb.N_TILE * UM_KB + sw128_offset(n, kk)] = to_tf32(v * sc);
    }
}

// ------------------------------------------------------------------ tape scales
// The fp16 tape holds the PRE-BatchNorm activations, whose scale is whatever the conv weights make it,
// while BatchNorm makes the function itself invariant to that scale.  To keep the tape inside fp16's
// range (no saturation at 65504, no flush into subnormals) for any weights, layer k runs on
// W'_k = s_k W_k with s_k = 2^-(e_w + e_a): e_w the binary exponent of max|W_k|, e_a that of
// max(|gamma_{k-1}|, |beta_{k-1}|) (the scale of the layer's input; 0 for layer 1, whose input is
// |x_i - x_j| of BatchNorm'd node features).  Exponents as frexp gives them, each and their sum clamped
// to [-60, 60]; both maxima are exact in any evaluation order, so oracle/gnn_oracle.py tape_scale()
// reproduces s_k bit for bit.  Consequences, all exact because s_k is a power of two:
//   tape and statistics hold h' = s h;  BN(h'; s^2 eps) = BN(h; eps): the slot's second-moment sum is
//   pre-loaded with count (s^2 - 1) eps (see bn_mean_rstd);  dgrad uses W' as is (h' = a W'^T);
//   dL/dW = s dL/dW' (wgrad_reduce_kernel).
struct ScaleJobs {
    const float* W[4];
    int wn[4];              // elements of W_k
    const float* g[4];      // gamma / beta of the PREVIOUS BatchNorm (null for layer 1)
    const float* b[4];
    int gc[4];              // channels of the previous BatchNorm
    int C[4];               // output channels of layer k
    float* tscale;          // [4]
    double* fsums;          // 4 statistics slots (zeroed on this stream before this launch)
    double count;           // BatchNorm population B*N*N
};

__device__ __forceinline__ int frexp_exponent(float v) {
    if (!(v > 0.f) || !isfinite(v)) return 0;
    int e;
    frexpf(v, &e);
    return max(-60, min(60, e));
}

__global__ void __launch_bounds__(1024) umma_layer_scales_kernel(ScaleJobs j) {   // grid = 4: one CTA per conv layer
    __shared__ float red[2][32];
    const int k = blockIdx.x, t = threadIdx.x;
    // max |W_k|: a latency-bound sweep of <= 45k floats -- 1024 threads, four independent 16-byte loads each
    float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f, ma = 0.f;
    const float* W = j.W[k];
    const int n = j.wn[k];
    if ((reinterpret_cast<uintptr_t>(W) & 15) == 0) {
        const float4* W4 = reinterpret_cast<const float4*>(W);
        const int n4 = n >> 2;
        auto amax4 = [](float4 v) { return fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))); };
        int i = t;
        for (; i + 3 * 1024 < n4; i += 4 * 1024) {
            const float4 a = __ldg(W4 + i), b = __ldg(W4 + i + 1024), c = __ldg(W4 + i + 2048), d = __ldg(W4 + i + 3072);
            m0 = fmaxf(m0, amax4(a)); m1 = fmaxf(m1, amax4(b)); m2 = fmaxf(m2, amax4(c)); m3 = fmaxf(m3, amax4(d));
        }
        for (; i < n4; i += 1024) m0 = fmaxf(m0, amax4(__ldg(W4 + i)));
        for (int q = (n4 << 2) + t; q < n; q += 1024) m1 = fmaxf(m1, fabsf(__ldg(W + q)));
    } else {
        for (int i = t; i < n; i += 1024) m0 = fmaxf(m0, fabsf(__ldg(W + i)));
    }
    float mw = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
    if (j.g[k])
        for (int c = t; c < j.gc[k]; c += 1024) ma = fmaxf(ma, fmaxf(fabsf(__ldg(j.g[k] + c)), fabsf(__ldg(j.b[k] + c))));
    mw = warp_max(mw);
    ma = warp_max(ma);
    if ((t & 31) == 0) { red[0][t >> 5] = mw; red[1][t >> 5] = ma; }
    __syncthreads();
    mw = red[0][0]; ma = red[1][0];
#pragma unroll
    for (int q = 1; q < 32; ++q) { mw = fmaxf(mw, red[0][q]); ma = fmaxf(ma, red[1][q]); }
    int e = frexp_exponent(mw) + (j.g[k] ? frexp_exponent(ma) : 0);
    e = max(-60, min(60, e));
    const float sc = ldexpf(1.f, -e);
    if (t == 0) j.tscale[k] = sc;
    const double corr = j.count * ((double)sc * (double)sc - 1.0) * (double)kBnEps;
    double* slot = j.fsums + (size_t)k * kStatSlot;              // copy 0, second moments
    for (int c = t; c < j.C[k]; c += 1024) slot[j.C[k] + c] = corr;
}

// ------------------------------------------------------------------ producer functors
// fetch(row, k, raw): issue the (read-only, non-coherent) global loads of four consecutive K
// elements starting at k (k % 4 == 0); finish(raw, k, aux): turn them into operand values, zero
// beyond K.  The split lets a producer thread keep all eight rows of a K block (and, for
// kDouble operands, the next block as well) in flight before it touches any of them.

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// NOTE on fetch(): it must be nothing but loads into fresh registers.  Zero-initialising a Raw and
// then conditionally loading into it makes ptxas protect the write-after-write with a scoreboard
// wait, and with six scoreboard slots for a dozen loads that serialises every load of a K block
// behind the previous one (seen in ncu as long_scoreboard stalls on CS2R).  So addresses are
// clamped into bounds instead, and finish() masks what lies beyond K.

struct AbsDiffU {
    static constexpr bool kTma = false;
    static constexpr int kAhead = 0;         // x is small and L2 resident; 8 loads in flight suffice
    const float* x;
    int ldx, F;
    PairGeom g;
    int vec_ok;                              // rows 16-byte aligned and ldx >= roundup4(F)
    struct Row { const float* xi; const float* xj; };
    struct Raw { float4 a, b; };
    __device__ __forceinline__ void init(float*, int, int) const {}
    __device__ __forceinline__ Row row(int r) const {
        const int2 n = __ldg(g.rowij + r);             // one load, no division (tri_table_kernel)
        return Row{x + (size_t)n.x * ldx, x + (size_t)n.y * ldx};
    }
    __device__ __forceinline__ void fetch(const Row& rw, int k, Raw& o) const {
        if (vec_ok) {
            const int kk = min(k, ((F + 3) & ~3) - 4);
            o.a = ldg4(rw.xi + kk);
            o.b = ldg4(rw.xj + kk);
        } else {
            const int k0 = min(k, F - 1), k1 = min(k + 1, F - 1), k2 = min(k + 2, F - 1), k3 = min(k + 3, F - 1);
            o.a = make_float4(__ldg(rw.xi + k0), __ldg(rw.xi + k1), __ldg(rw.xi + k2), __ldg(rw.xi + k3));
            o.b = make_float4(__ldg(rw.xj + k0), __ldg(rw.xj + k1), __ldg(rw.xj + k2), __ldg(rw.xj + k3));
        }
    }
    __device__ __forceinline__ float4 finish(const Raw& r, const Row&, int k, const float*) const {
        float4 v = make_float4(fabsf(r.a.x - r.b.x), fabsf(r.a.y - r.b.y), fabsf(r.a.z - r.b.z),
                               fabsf(r.a.w - r.b.w));
        if (k + 0 >= F) v.x = 0.f;
        if (k + 1 >= F) v.y = 0.f;
        if (k + 2 >= F) v.z = 0.f;
        if (k + 3 >= F) v.w = 0.f;
        return v;
    }
};

__device__ __forceinline__ uint2 ldg8(const void* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }

struct PlainU {
    static constexpr bool kTma = false;
    static constexpr int kAhead = 2;
    const float* p;
    int ld, K;
    int vec_ok;                              // rows 16-byte aligned and K % 4 == 0
    struct Row { const float* q; };
    struct Raw { float4 v; };
    __device__ __forceinline__ void init(float*, int, int) const {}
    __device__ __forceinline__ Row row(int r) const { return Row{p + (size_t)r * ld}; }
    __device__ __forceinline__ void fetch(const Row& rw, int k, Raw& o) const {
        if (vec_ok) {
            o.v = ldg4(rw.q + min(k, K - 4));
        } else {
            const int k0 = min(k, K - 1), k1 = min(k + 1, K - 1), k2 = min(k + 2, K - 1), k3 = min(k + 3, K - 1);
            o.v = make_float4(__ldg(rw.q + k0), __ldg(rw.q + k1), __ldg(rw.q + k2), __ldg(rw.q + k3));
        }
    }
    __device__ __forceinline__ float4 finish(const Raw& r, const Row&, int k, const float*) const {
        float4 v = r.v;
        if (k + 0 >= K) v.x = 0.f;
        if (k + 1 >= K) v.y = 0.f;
        if (k + 2 >= K) v.z = 0.f;
        if (k + 3 >= K) v.w = 0.f;
        return v;
    }
};

// dH = gamma*rstd*(dy - w*m1 - w*hhat*m2): BatchNorm backward of the twin-summed gradient, fused
// into the operand build so that dH never exists in HBM.  With P = gamma*rstd, S = P*rstd*m2,
// Q = P*m1 - S*mean this is  dH = P*dy - w*(Q + S*h)  (C % 4 == 0): DhInPlaceT (rows kernel) and DhT (wgrad).

// a = LeakyReLU(scale*h + shift), scale = gamma*rstd, shift = beta - mean*scale (C % 4 == 0), tensor-map fed: the raw [128 rows x 32 ch] block of H lands in
// the ring stage by TMA, already in the K-major SWIZZLE_128B arrangement the MMA wants (the TMA and
// UMMA 128-byte swizzles are the same function), and the producer warps apply BN + LeakyReLU + TF32
// rounding IN PLACE.  No register staging, no scoreboard-limited prefetch: every free stage has a
// 16 KB load in flight.
struct BnActT {
    static constexpr bool kTma = true;
    static constexpr bool kInPlace = false;
    static constexpr int kAhead = 0;
    alignas(64) CUtensorMap tmap;
    int C;
    const double* sums;
    const float* gamma;
    const float* beta;
    double inv_count;
    struct Row { int dummy; };
    struct Raw { int dummy; };
    __device__ __forceinline__ void init(float* aux, int tid, int nthreads) const {
        for (int c = tid; c < C; c += nthreads) {
            float m, r;
            bn_mean_rstd(sums, C, c, inv_count, m, r);
            float sc = gamma[c] * r;
            aux[c] = sc;
            aux[kMaxC + c] = beta[c] - m * sc;
        }
    }
    __device__ __forceinline__ Row row(int) const { return Row{0}; }
    __device__ __forceinline__ void fetch(const Row&, int, Raw&) const {}
    __device__ __forceinline__ float4 finish(const Raw&, const Row&, int, const float*) const {
        return make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // per-K-block constants of the thread's four channels: loaded once, used for all its rows
    struct Consts { float4 sc, sh; bool on; };
    __device__ __forceinline__ Consts consts(int k, const float* aux) const {
        Consts c;
        c.on = k < C;
        const int kk = c.on ? k : 0;
        c.sc = *reinterpret_cast<const float4*>(aux + kk);
        c.sh = *reinterpret_cast<const float4*>(aux + kMaxC + kk);
        return c;
    }
    __device__ __forceinline__ float4 transform(uint2 raw, const Consts& c) const {
        if (!c.on) return make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 h = unpack_half4(raw);
        const f2 ylo = fma2(pack2(h.x, h.y), pack2(c.sc.x, c.sc.y), pack2(c.sh.x, c.sh.y));
        const f2 yhi = fma2(pack2(h.z, h.w), pack2(c.sc.z, c.sc.w), pack2(c.sh.z, c.sh.w));
        const f2 sl = pack2(kSlope, kSlope);
        float4 y, z;
        unpack2(ylo, y.x, y.y); unpack2(yhi, y.z, y.w);
        unpack2(mul2(ylo, sl), z.x, z.y); unpack2(mul2(yhi, sl), z.z, z.w);
        return make_float4(fmaxf(y.x, z.x), fmaxf(y.y, z.y), fmaxf(y.z, z.z), fmaxf(y.w, z.w));
    }
};

// dH for the rows kernel (dgrad A operand), tensor-map fed.  The fp32 [128 rows x 32 ch] block
// of dy lands by TMA (SWIZZLE_128B) DIRECTLY in the ring stage -- the K-major swizzled arrangement
// the MMA reads -- and the matching fp16 block of H in a raw slot of the same index; the producer
// threads then turn dy into dH = P*dy - w*(Q + S*h) in place.  No register staging: every stage the
// MMA has released has 24 KB of loads in flight.
struct DhInPlaceT {
    static constexpr bool kTma = true;
    static constexpr bool kInPlace = true;
    static constexpr int kAhead = 0;
    alignas(64) CUtensorMap tmap_dy;         // fp32 [R, C], box {32, 128}, SWIZZLE_128B
    alignas(64) CUtensorMap tmap;            // fp16 [R, C], box {32, 128}, unswizzled (64-byte rows)
    int C;
    const double* fsums;
    const float* gamma;
    const double* bsums;
    double inv_count;
    PairGeom g;
    struct Row { int dummy; };
    struct Raw { int dummy; };
    __device__ __forceinline__ void init(float* aux, int tid, int nthreads) const {
        for (int c = tid; c < C; c += nthreads) {
            float m, r;
            bn_mean_rstd(fsums, C, c, inv_count, m, r);
            float P = gamma[c] * r;
            float S = P * r * (float)(stat_get(bsums, C, c, 1) * inv_count);
            aux[c] = P;
            aux[kMaxC + c] = P * (float)(stat_get(bsums, C, c, 0) * inv_count) - S * m;
            aux[2 * kMaxC + c] = S;
        }
    }
    __device__ __forceinline__ Row row(int) const { return Row{0}; }
    __device__ __forceinline__ void fetch(const Row&, int, Raw&) const {}
    __device__ __forceinline__ float4 finish(const Raw&, const Row&, int, const float*) const {
        return make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __device__ __forceinline__ float row_weight(int r) const { return __ldg(g.roww + r); }   // r <= R
    struct Consts { float4 P, Q, S; bool on; };
    __device__ __forceinline__ Consts consts(int k, const float* aux) const {
        Consts c;
        c.on = k < C;
        const int kk = c.on ? k : 0;
        c.P = *reinterpret_cast<const float4*>(aux + kk);
        c.Q = *reinterpret_cast<const float4*>(aux + kMaxC + kk);
        c.S = *reinterpret_cast<const float4*>(aux + 2 * kMaxC + kk);
        return c;
    }
    __device__ __forceinline__ float4 transform2(float4 d, uint2 hraw, float w, const Consts& c) const {
        if (!c.on) return make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 h = unpack_half4(hraw);
        const f2 nw = pack2(-w, -w);
        const f2 lo = fma2(nw, fma2(pack2(c.S.x, c.S.y), pack2(h.x, h.y), pack2(c.Q.x, c.Q.y)),
                           mul2(pack2(c.P.x, c.P.y), pack2(d.x, d.y)));
        const f2 hi = fma2(nw, fma2(pack2(c.S.z, c.S.w), pack2(h.z, h.w), pack2(c.Q.z, c.Q.w)),
                           mul2(pack2(c.P.z, c.P.w), pack2(d.z, d.w)));
        float4 r;
        unpack2(lo, r.x, r.y); unpack2(hi, r.z, r.w);
        return r;
    }
};

// Tensor-map operands of the wgrad kernel: whole [32 rows x C] row-major slabs land in a raw ring
// (dy fp32, H fp16), producers convert them into the MN-major operand blocks.
struct DhT {                                     // P = dH_k from (dy_k, H_k)
    static constexpr bool kTma = true;
    static constexpr int kAhead = 0;
    alignas(64) CUtensorMap tmap_dy;             // fp32 [R, C], box {C, 32}
    alignas(64) CUtensorMap tmap_h;              // fp16 [R, C], box {C, 32}
    int C;
    const double* fsums;
    const float* gamma;
    const double* bsums;
    double inv_count;
    PairGeom g;
    struct Row { float w; };
    struct Raw { int dummy; };
    __device__ __forceinline__ void init(float* aux, int tid, int nthreads) const {
        for (int c = tid; c < C; c += nthreads) {
            float m, r;
            bn_mean_rstd(fsums, C, c, inv_count, m, r);
            float P = gamma[c] * r;
            float S = P * r * (float)(stat_get(bsums, C, c, 1) * inv_count);
            aux[c] = P;
            aux[kMaxC + c] = P * (float)(stat_get(bsums, C, c, 0) * inv_count) - S * m;
            aux[2 * kMaxC + c] = S;
        }
    }
    __device__ __forceinline__ Row row(int r) const { return Row{__ldg(g.roww + r)}; }       // r <= R
    __device__ __forceinline__ void fetch(const Row&, int, Raw&) const {}
    __device__ __forceinline__ float4 finish(const Raw&, const Row&, int, const float*) const {
        return make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __device__ __forceinline__ uint32_t raw_bytes() const { return WG_ROWS_C * C * 6u; }   // fp32 + fp16 slab
    // The per-channel constants of a thread's 4 channels never change inside a launch: the wgrad producers
    // hold them in registers (kload once) instead of three LDS.128 per block per chunk -- those loads
    // sat between the ring stores (possible aliasing kept the compiler from batching them) and were half
    // of the role's stall samples.
    struct K { float4 P, Q, S; };
    __device__ __forceinline__ K kload(int k, const float* aux) const {
        const int kk = k < C ? k : 0;
        return K{*reinterpret_cast<const float4*>(aux + kk), *reinterpret_cast<const float4*>(aux + kMaxC + kk),
                 *reinterpret_cast<const float4*>(aux + 2 * kMaxC + kk)};
    }
    __device__ __forceinline__ float4 transform_k(float4 d, uint2 hraw, float w, int k, const K& c) const {
        if (k >= C) return make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 h = unpack_half4(hraw);
        const float4 P = c.P, Q = c.Q, S = c.S;
        const f2 nw = pack2(-w, -w);
        const f2 lo = fma2(nw, fma2(pack2(S.x, S.y), pack2(h.x, h.y), pack2(Q.x, Q.y)), mul2(pack2(P.x, P.y), pack2(d.x, d.y)));
        const f2 hi = fma2(nw, fma2(pack2(S.z, S.w), pack2(h.z, h.w), pack2(Q.z, Q.w)), mul2(pack2(P.z, P.w), pack2(d.z, d.w)));
        float4 r;
        unpack2(lo, r.x, r.y); unpack2(hi, r.z, r.w);
        return r;
    }
    __device__ __forceinline__ float4 transform(float4 d, uint2 hraw, float w, int k, const float* aux) const {
        if (k >= C) return make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 h = unpack_half4(hraw);
        float4 P = *reinterpret_cast<const float4*>(aux + k);
        float4 Q = *reinterpret_cast<const float4*>(aux + kMaxC + k);
        float4 S = *reinterpret_cast<const float4*>(aux + 2 * kMaxC + k);
        const f2 nw = pack2(-w, -w);
        const f2 lo = fma2(nw, fma2(pack2(S.x, S.y), pack2(h.x, h.y), pack2(Q.x, Q.y)), mul2(pack2(P.x, P.y), pack2(d.x, d.y)));
        const f2 hi = fma2(nw, fma2(pack2(S.z, S.w), pack2(h.z, h.w), pack2(Q.z, Q.w)), mul2(pack2(P.z, P.w), pack2(d.z, d.w)));
        float4 r;
        unpack2(lo, r.x, r.y); unpack2(hi, r.z, r.w);
        return r;
    }
};

struct BnActQT {                                 // Q = LeakyReLU(BN(H_{k-1}))
    static constexpr bool kTma = true;
    static constexpr int kAhead = 1;             // (not a late-fetched operand)
    alignas(64) CUtensorMap tmap_h;              // fp16 [R, C], box {C, 32}
    int C;
    const double* sums;
    const float* gamma;
    const float* beta;
    double inv_count;
    struct Row { int dummy; };
    struct Raw { int dummy; };
    __device__ __forceinline__ void init(float* aux, int tid, int nthreads) const {
        for (int c = tid; c < C; c += nthreads) {
            float m, r;
            bn_mean_rstd(sums, C, c, inv_count, m, r);
            float sc = gamma[c] * r;
            aux[c] = sc;
            aux[kMaxC + c] = beta[c] - m * sc;
        }
    }
    __device__ __forceinline__ Row row(int) const { return Row{0}; }
    __device__ __forceinline__ void fetch(const Row&, int, Raw&) const {}
    __device__ __forceinline__ float4 finish(const Raw&, const Row&, int, const float*) const {
        return make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __device__ __forceinline__ uint32_t raw_bytes() const { return WG_ROWS_C * C * 2u; }
    struct K { float4 sc, sh; };                 // see DhT::K
    __device__ __forceinline__ K kload(int k, const float* aux) const {
        const int kk = k < C ? k : 0;
        return K{*reinterpret_cast<const float4*>(aux + kk), *reinterpret_cast<const float4*>(aux + kMaxC + kk)};
    }
    __device__ __forceinline__ float4 transform_k(uint2 hraw, int k, const K& c) const {
        if (k >= C) return make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 h = unpack_half4(hraw);
        const float4 sc = c.sc, sh = c.sh;
        const f2 ylo = fma2(pack2(h.x, h.y), pack2(sc.x, sc.y), pack2(sh.x, sh.y));
        const f2 yhi = fma2(pack2(h.z, h.w), pack2(sc.z, sc.w), pack2(sh.z, sh.w));
        const f2 sl = pack2(kSlope, kSlope);
        float4 y, z;
        unpack2(ylo, y.x, y.y); unpack2(yhi, y.z, y.w);
        unpack2(mul2(ylo, sl), z.x, z.y); unpack2(mul2(yhi, sl), z.z, z.w);
        return make_float4(fmaxf(y.x, z.x), fmaxf(y.y, z.y), fmaxf(y.z, z.z), fmaxf(y.w, z.w));
    }
    __device__ __forceinline__ float4 transform(uint2 hraw, int k, const float* aux) const {
        if (k >= C) return make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 h = unpack_half4(hraw);
        float4 sc = *reinterpret_cast<const float4*>(aux + k);
        float4 sh = *reinterpret_cast<const float4*>(aux + kMaxC + k);
        const f2 ylo = fma2(pack2(h.x, h.y), pack2(sc.x, sc.y), pack2(sh.x, sh.y));
        const f2 yhi = fma2(pack2(h.z, h.w), pack2(sc.z, sc.w), pack2(sh.z, sh.w));
        const f2 sl = pack2(kSlope, kSlope);
        float4 y, z;
        unpack2(ylo, y.x, y.y); unpack2(yhi, y.z, y.w);
        unpack2(mul2(ylo, sl), z.x, z.y); unpack2(mul2(yhi, sl), z.z, z.w);
        return make_float4(fmaxf(y.x, z.x), fmaxf(y.y, z.y), fmaxf(y.z, z.z), fmaxf(y.w, z.w));
    }
};

// ------------------------------------------------------------------ epilogue functors
// apply(): four consecutive output columns col..col+3 (col % 4 == 0) of global row r;
// nvalid = how many of them exist.  s0/s1: the thread's running column statistics.

struct EpiStoreU {
    static constexpr bool kPrefetch = false;
    static constexpr bool kStats = false;
    static constexpr bool kRowWeight = false;
    float* out;
    int ld;
    int vec_ok;
    __device__ __forceinline__ void init(float*, int, int) const {}
    __device__ __forceinline__ float row_weight(int) const { return 1.f; }
    struct Consts {};
    __device__ __forceinline__ int row_stride() const { return ld; }
    __device__ __forceinline__ Consts consts(int, const float*) const { return Consts{}; }
    __device__ __forceinline__ uint2 prefetch(unsigned, int) const { return make_uint2(0u, 0u); }
    __device__ __forceinline__ void apply(unsigned roff, int chcol, bool ok, float, int, float4 v, uint2, int nvalid,
                                          float*, float*, const Consts&) const {
        float* o = (out + roff) + chcol;
        if (!ok) return;
        if (vec_ok && nvalid == 4) {
            *reinterpret_cast<float4*>(o) = v;
        } else {
            o[0] = v.x;
            if (nvalid > 1) o[1] = v.y;
            if (nvalid > 2) o[2] = v.z;
            if (nvalid > 3) o[3] = v.w;
        }
    }
    __device__ __forceinline__ void commit(int, float, float, const float*) const {}
};

// dD (gradient of |x_i - x_j|, consumed only by the dx gather) is stored as bf16: fp32 range, and its
// 2^-9 rounding is far below what TF32 leaves in the gradients anyway (ld % 4 == 0, padded columns
// hold exact zeros because the padded weight-image rows are zero).
struct EpiStoreBf16U {
    static constexpr bool kPrefetch = false;
    static constexpr bool kStats = false;
    static constexpr bool kRowWeight = false;
    __nv_bfloat16* out;
    int ld;
    __device__ __forceinline__ void init(float*, int, int) const {}
    __device__ __forceinline__ float row_weight(int) const { return 1.f; }
    struct Consts {};
    __device__ __forceinline__ int row_stride() const { return ld; }
    __device__ __forceinline__ Consts consts(int, const float*) const { return Consts{}; }
    __device__ __forceinline__ uint2 prefetch(unsigned, int) const { return make_uint2(0u, 0u); }
    __device__ __forceinline__ void apply(unsigned roff, int chcol, bool ok, float, int col, float4 v, uint2, int,
                                          float*, float*, const Consts&) const {
        uint2* o = reinterpret_cast<uint2*>((out + roff) + chcol);
        const uint2 pk = pack_bf4(v);
        if (ok && col < ld) *o = pk;
    }
    __device__ __forceinline__ void commit(int, float, float, const float*) const {}
};

// forward: store pre-BN H, accumulate sum w*h and sum w*h^2 (C % 4 == 0)
struct EpiFwdStatsU {
    static constexpr bool kPrefetch = false;
    static constexpr bool kStats = true;
    static constexpr bool kRowWeight = true;
    __half* H;                               // fp16 tape
    int C;
    double* sums;
    PairGeom g;
    __device__ __forceinline__ void init(float*, int, int) const {}
    __device__ __forceinline__ float row_weight(int r) const { return __ldg(g.roww + r); }   // r <= R
    struct Consts {};
    __device__ __forceinline__ int row_stride() const { return C; }
    __device__ __forceinline__ Consts consts(int, const float*) const { return Consts{}; }
    __device__ __forceinline__ uint2 prefetch(unsigned, int) const { return make_uint2(0u, 0u); }
    __device__ __forceinline__ void apply(unsigned roff, int chcol, bool ok, float w, int, float4 v, uint2, int,
                                          float* s0, float* s1, const Consts&) const {
        const uint2 packed = pack_half4(v);
        uint2* o = reinterpret_cast<uint2*>((H + roff) + chcol);
        if (ok) *o = packed;
        // statistics of exactly what the next layer will read (taking them from the fp32 accumulators
        // saves four instructions per access but moves enough LeakyReLU kinks on the 84-row fixture to
        // break its gradient tolerance, and bought nothing measurable: the role is not issue-bound)
        v = unpack_half4(packed);
        const f2 ww = pack2(w, w), vlo = pack2(v.x, v.y), vhi = pack2(v.z, v.w);
        unpack2(fma2(ww, vlo, pack2(s0[0], s0[1])), s0[0], s0[1]);
        unpack2(fma2(ww, vhi, pack2(s0[2], s0[3])), s0[2], s0[3]);
        unpack2(fma2(mul2(ww, vlo), vlo, pack2(s1[0], s1[1])), s1[0], s1[1]);
        unpack2(fma2(mul2(ww, vhi), vhi, pack2(s1[2], s1[3])), s1[2], s1[3]);
    }
    __device__ __forceinline__ void commit(int c, float v0, float v1, const float*) const {
        stat_add(sums, C, c, 0, v0);
        stat_add(sums, C, c, 1, v1);
    }
};

// dgrad: acc = dL/d a_{k-1}; dy = acc * lrelu'(BN(H_{k-1})), stored; reductions sum dy and
// sum dy*hhat (accumulated as sum dy*h and corrected per CTA: hhat = (h - mean) * rstd)
struct EpiDyU {
    static constexpr bool kPrefetch = true;    // needs H_{k-1}(r, col): loaded one chunk ahead
    static constexpr bool kStats = true;
    static constexpr bool kRowWeight = false;
    const __half* H;             // pre-BN activations of layer k-1 (fp16 tape)
    float* dy;                   // gradient tape (fp32)
    int C;
    const double* fsums;
    const float* gamma;
    const float* beta;
    double inv_count;
    double* bsums;
    __device__ __forceinline__ void init(float* aux, int tid, int nthreads) const {
        for (int c = tid; c < C; c += nthreads) {
            float m, r;
            bn_mean_rstd(fsums, C, c, inv_count, m, r);
            float sc = gamma[c] * r;
            aux[c] = sc;
            aux[kMaxC + c] = beta[c] - m * sc;
            aux[2 * kMaxC + c] = m;
            aux[3 * kMaxC + c] = r;
        }
    }
    __device__ __forceinline__ float row_weight(int) const { return 1.f; }
    struct Consts { float4 sc, sh; };
    __device__ __forceinline__ int row_stride() const { return C; }
    __device__ __forceinline__ Consts consts(int col, const float* aux) const {   // once per chunk, not per row
        Consts c;
        c.sc = *reinterpret_cast<const float4*>(aux + col);
        c.sh = *reinterpret_cast<const float4*>(aux + kMaxC + col);
        return c;
    }
    __device__ __forceinline__ uint2 prefetch(unsigned roff, int chcol) const { return ldg8((H + roff) + chcol); }
    __device__ __forceinline__ void apply(unsigned roff, int chcol, bool ok, float, int, float4 v, uint2 hraw, int,
                                          float* s0, float* s1, const Consts& k) const {
        const float4 h = unpack_half4(hraw);
        const f2 hlo = pack2(h.x, h.y), hhi = pack2(h.z, h.w);
        float4 y;
        unpack2(fma2(hlo, pack2(k.sc.x, k.sc.y), pack2(k.sh.x, k.sh.y)), y.x, y.y);
        unpack2(fma2(hhi, pack2(k.sc.z, k.sc.w), pack2(k.sh.z, k.sh.w)), y.z, y.w);
        const f2 dlo = mul2(pack2(v.x, v.y), pack2(y.x > 0.f ? 1.f : kSlope, y.y > 0.f ? 1.f : kSlope));
        const f2 dhi = mul2(pack2(v.z, v.w), pack2(y.z > 0.f ? 1.f : kSlope, y.w > 0.f ? 1.f : kSlope));
        float4 d;
        unpack2(dlo, d.x, d.y); unpack2(dhi, d.z, d.w);
        float4* o = reinterpret_cast<float4*>((dy + roff) + chcol);
        if (ok) *o = d;
        unpack2(add2(pack2(s0[0], s0[1]), dlo), s0[0], s0[1]);
        unpack2(add2(pack2(s0[2], s0[3]), dhi), s0[2], s0[3]);
        unpack2(fma2(dlo, hlo, pack2(s1[0], s1[1])), s1[0], s1[1]);
        unpack2(fma2(dhi, hhi, pack2(s1[2], s1[3])), s1[2], s1[3]);
    }
    __device__ __forceinline__ void commit(int c, float v0, float v1, const float* aux) const {
        float m = aux[2 * kMaxC + c], r = aux[3 * kMaxC + c];
        stat_add(bsums, C, c, 0, v0);
        stat_add(bsums, C, c, 1, r * (v1 - m * v0));
    }
};

// ------------------------------------------------------------------ the kernel
template <class AOp> __host__ __device__ constexpr bool tma_in_place() {
    if constexpr (AOp::kTma) return AOp::kInPlace; else return false;
}

template <class AOp, class Epi>
__global__ void __launch_bounds__(UM_THREADS, 1)
umma_rows_kernel(const __grid_constant__ AOp aop, const __grid_constant__ Epi epi, const float* __restrict__ wimg,
                 const __grid_constant__ UmmaShape s) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by OFFSETTING the shared array (an integer round trip would turn every
    // shared-memory pointer below into a generic one: LD/ST instead of LDS/STS throughout the kernel)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t w_bytes = (uint32_t)s.KC * s.N_TILE * 128;
    float* Wsm = reinterpret_cast<float*>(smem);
    float* Asm = reinterpret_cast<float*>(smem + w_bytes);
    float* stage = Asm + (size_t)s.stages * UM_BLOCK_FLOATS;
    float* aux_a = stage + UM_ROWS * UM_STAGE_LD;
    float* aux_e = aux_a + 3 * kMaxC;
    uint8_t* rawring = reinterpret_cast<uint8_t*>(aux_e + 4 * kMaxC);   // [raw_stages][UM_RAW_BYTES], kTma only
    uint64_t* bars = reinterpret_cast<uint64_t*>(rawring + (size_t)s.raw_stages * UM_RAW_BYTES);
    uint64_t* full = bars;          // [UM_MAX_STAGES]
    uint64_t* empty = bars + 4;     // [UM_MAX_STAGES]
    uint64_t* tfull = bars + 8;     // [2]
    uint64_t* tempty = bars + 10;   // [2]
    uint64_t* wbar = bars + 12;
    uint64_t* rawfull = bars + 13;   // [UM_MAX_RAW]  TMA operands: raw fp16 block landed
    uint64_t* rawempty = bars + 21;  // [UM_MAX_RAW]  producers are done reading it
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 30);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int ntiles = (s.R + UM_ROWS - 1) / UM_ROWS;
    // physical row tile of logical tile t: consecutive launches alternate direction so that a
    // kernel starts on the rows its predecessor touched last (still L2 resident)
    auto phys = [&](int t) { return s.reverse ? ntiles - 1 - t : t; };

    if (tid == 0) {
        for (int i = 0; i < UM_MAX_STAGES; ++i) {
            mbar_init(&full[i], UM_PROD_WARPS);     // one elected arrival per producer warp
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < UM_MAX_RAW; ++i) {
            mbar_init(&rawfull[i], 1);
            mbar_init(&rawempty[i], UM_PROD_WARPS);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], 4);               // one per epilogue warp
        }
        mbar_init(wbar, 1);
        fence_mbar_init();
    }
    if (warp == UM_MMA_WARP) tmem_alloc(tmem_slot, UM_TMEM_COLS);
    __syncthreads();                                   // barrier inits visible to the MMA thread
    if (warp == UM_MMA_WARP && lane == 0) {
        // The weight image was complete before the PREVIOUS launch of this stream started (host
        // side: pdl only between consecutive GEMM launches of one call), so under programmatic
        // dependent launch it may land while that launch is still draining.
        mbar_arrive_expect_tx(wbar, w_bytes);
        const uint32_t blk = (uint32_t)s.N_TILE * 128;
        for (int kc = 0; kc < s.KC; ++kc)
            bulk_g2s(reinterpret_cast<uint8_t*>(Wsm) + (size_t)kc * blk,
                     reinterpret_cast<const uint8_t*>(wimg) + (size_t)kc * blk, blk, wbar);
    }
    // Everything above overlaps the tail of the preceding kernel when this one was launched with
    // programmatic stream serialization; everything below reads what that kernel wrote.  (Both
    // instructions are no-ops for an ordinary launch.)  Dependents are released only AFTER the
    // wait, so a kernel's prologue never runs beside anything older than its direct predecessor.
    pdl_wait();
    pdl_launch_dependents();
    aop.init(aux_a, tid, UM_THREADS);
    epi.init(aux_e, tid, UM_THREADS);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    long long* dbg = s.dbg ? s.dbg + (size_t)blockIdx.x * 16 : nullptr;
#define MFT_MARK(slot) do { if (dbg && lane == 0) dbg[slot] = clock64(); } while (0)
    if (dbg && tid == 0) dbg[0] = clock64();

    if (warp < UM_PROD_WARPS) {
        // ===================== producers =====================
        // Thread t owns 16-byte column c16 = t%8 of rows q*32 + t/8 (q = 0..3) of every K block, so
        // its per-channel constants never change within a block and a warp's loads cover four
        // full 128-byte row segments.  The fetch iterator runs AOp::kAhead K blocks ahead of the
        // write iterator so that global latency overlaps the transform + smem stores.
        reg_dec<96>();   // 8 warps x 32 regs released ...
        constexpr int RQ = UM_ROWS * 8 / UM_PROD_THREADS;      // rows per thread per K block (4)
        constexpr int RSTEP = UM_PROD_THREADS / 8;              // 32
        if constexpr (tma_in_place<AOp>()) {
            // In-place operands (DhInPlaceT): stage st holds the fp32 block the TMA warp landed (already
            // swizzled), raw slot st the fp16 companion; transform own elements and hand over to the MMA.
            const int rsub = tid >> 3, c16 = tid & 7;
            const int sw = rsub & 7;
            int st = 0;
            uint32_t ph = 0;
            // row multiplicities come from the pair table (a dependent global load): fetched one tile ahead
            float wq[RQ], wq_next[RQ];
            auto load_weights = [&](int tile, float (&dst)[RQ]) {
#pragma unroll
                for (int q = 0; q < RQ; ++q) {
                    // pure loads (entry R of the row table is 0): nothing here may wait on a result
                    const int r = phys(min(tile, ntiles - 1)) * UM_ROWS + q * RSTEP + rsub;
                    dst[q] = aop.row_weight(tile < ntiles ? min(r, s.R) : s.R);
                }
            };
            load_weights(blockIdx.x, wq_next);
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int row0 = phys(tile) * UM_ROWS;
                const bool partial = row0 + UM_ROWS > s.R;       // only the last tile has rows past the end
#pragma unroll
                for (int q = 0; q < RQ; ++q) wq[q] = wq_next[q];
                load_weights(tile + gridDim.x, wq_next);
                for (int kc = 0; kc < s.KC; ++kc) {
                    mbar_wait(&rawfull[st], ph);
                    const uint8_t* rawb = rawring + (size_t)st * UM_RAW_BYTES;
                    float* blk = Asm + (size_t)st * UM_BLOCK_FLOATS;
                    const int k = kc * UM_KB + c16 * 4;
                    uint2 hraw[RQ];
                    float4 d[RQ];
                    const typename AOp::Consts kc4 = aop.consts(k, aux_a);
#pragma unroll
                    for (int q = 0; q < RQ; ++q) {
                        const int rl = q * RSTEP + rsub;
                        hraw[q] = *reinterpret_cast<const uint2*>(rawb + rl * 64 + c16 * 8);
                        d[q] = *reinterpret_cast<const float4*>(blk + (rl >> 3) * 256 + sw * 32 + ((c16 ^ sw) << 2));
                    }
#pragma unroll
                    for (int q = 0; q < RQ; ++q) {
                        const int rl = q * RSTEP + rsub;
                        float4 v = aop.transform2(d[q], hraw[q], wq[q], kc4);
                        if (partial && row0 + rl >= s.R) v = make_float4(0.f, 0.f, 0.f, 0.f);
                        v.x = to_tf32_fast(v.x); v.y = to_tf32_fast(v.y); v.z = to_tf32_fast(v.z); v.w = to_tf32_fast(v.w);
                        *reinterpret_cast<float4*>(blk + (rl >> 3) * 256 + sw * 32 + ((c16 ^ sw) << 2)) = v;
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full[st]);
                    if (++st == s.stages) { st = 0; ph ^= 1; }
                }
                if (warp == 0 && tile == (int)blockIdx.x) MFT_MARK(2);
            }
        } else if constexpr (AOp::kTma) {
            // The TMA warp lands raw fp16 blocks [128 rows x 32 ch] (row = 64 bytes, unswizzled) in the
            // raw ring; each producer thread converts its (row, 4-channel) pieces, applies BN +
            // LeakyReLU + TF32 rounding and writes the fp32 K block of the A ring, swizzled.
            const int rsub = tid >> 3, c16 = tid & 7;
            const int sw = rsub & 7;
            int st = 0, rs = 0;
            uint32_t ph = 0, rph = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int row0 = phys(tile) * UM_ROWS;
                const bool partial = row0 + UM_ROWS > s.R;       // only the last tile has rows past the end
                for (int kc = 0; kc < s.KC; ++kc) {
                    mbar_wait(&rawfull[rs], rph);
                    const uint8_t* rawb = rawring + (size_t)rs * UM_RAW_BYTES;
                    uint2 raw[RQ];
#pragma unroll
                    for (int q = 0; q < RQ; ++q)
                        raw[q] = *reinterpret_cast<const uint2*>(rawb + (q * RSTEP + rsub) * 64 + c16 * 8);
                    // Release the slot only once the loads have COMPLETED in every lane, not merely been
                    // issued: the TMA thread refills a released slot at once, and an LDS can sit in the memory
                    // pipeline behind the other warps' stores and the MMA's operand reads (the wgrad kernel
                    // showed exactly that race on B200).  The arrival count is made to depend on every
                    // lane's loaded registers through a warp-wide OR with a run-time zero, so neither the
                    // compiler nor the hardware can let the arrive overtake the loads.
                    {
                        unsigned t = 0;
#pragma unroll
                        for (int q = 0; q < RQ; ++q) t |= raw[q].x | raw[q].y;
                        const unsigned z = __reduce_or_sync(0xffffffffu, t & s.zero);
                        if (lane == 0) mbar_arrive_n(&rawempty[rs], 1u + z);
                    }
                    if (++rs == s.raw_stages) { rs = 0; rph ^= 1; }
                    mbar_wait(&empty[st], ph ^ 1);
                    float* blk = Asm + (size_t)st * UM_BLOCK_FLOATS;
                    const int k = kc * UM_KB + c16 * 4;
                    const typename AOp::Consts kc4 = aop.consts(k, aux_a);
#pragma unroll
                    for (int q = 0; q < RQ; ++q) {
                        const int rl = q * RSTEP + rsub;
                        float4 v = aop.transform(raw[q], kc4);
                        if (partial && row0 + rl >= s.R) v = make_float4(0.f, 0.f, 0.f, 0.f);
                        v.x = to_tf32_fast(v.x); v.y = to_tf32_fast(v.y); v.z = to_tf32_fast(v.z); v.w = to_tf32_fast(v.w);
                        *reinterpret_cast<float4*>(blk + (rl >> 3) * 256 + sw * 32 + ((c16 ^ sw) << 2)) = v;
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full[st]);
                    if (++st == s.stages) { st = 0; ph ^= 1; }
                }
                if (warp == 0 && tile == (int)blockIdx.x) MFT_MARK(2);
            }
        } else {
        constexpr int NBUF = AOp::kAhead + 1;
        const int rsub = tid >> 3, c16 = tid & 7;
        const int sw = rsub & 7;
        int st = 0;
        uint32_t ph = 0;
        typename AOp::Row rc[RQ];               // rows of the tile the fetch iterator is in
        typename AOp::Raw raw[NBUF][RQ];
        typename AOp::Row rrow[NBUF][RQ];       // rows each in-flight buffer belongs to
        uint32_t vmask[NBUF];
        uint32_t vrows = 0;
        int f_tile = blockIdx.x, f_kc = 0;
        auto load_rows = [&](int tile) {
            vrows = 0;
#pragma unroll
            for (int q = 0; q < RQ; ++q) {
                int r = phys(tile) * UM_ROWS + q * RSTEP + rsub;
                bool ok = r < s.R;
                vrows |= (ok ? 1u : 0u) << q;
                rc[q] = aop.row(ok ? r : 0);           // row 0 stands in for rows past the end
            }
        };
        // issue the loads of the block the fetch iterator points at into raw[slot], then advance it
        auto fetch_next = [&](typename AOp::Raw (&dst)[RQ], typename AOp::Row (&drow)[RQ], uint32_t& vm) {
            if (f_tile >= ntiles) return;
            if (f_kc == 0) load_rows(f_tile);
            const int k = f_kc * UM_KB + c16 * 4;
            vm = vrows;
#pragma unroll
            for (int q = 0; q < RQ; ++q) {
                aop.fetch(rc[q], k, dst[q]);
                drow[q] = rc[q];
            }
            if (++f_kc == s.KC) { f_kc = 0; f_tile += gridDim.x; }
        };
#pragma unroll
        for (int b = 0; b < NBUF; ++b) fetch_next(raw[b], rrow[b], vmask[b]);
        int w_tile = blockIdx.x, w_kc = 0;
        // One step = write the K block held in buffer b, then refill b with the block NBUF ahead.
        // Buffers are addressed statically (the loop is unrolled NBUF times): copying a register
        // that an in-flight load still targets would wait for that load and flatten the pipeline.
        auto step = [&](typename AOp::Raw (&buf)[RQ], typename AOp::Row (&brow)[RQ], uint32_t& vm) -> bool {
            if (w_tile >= ntiles) return false;
            mbar_wait(&empty[st], ph ^ 1);
            float* dst = Asm + (size_t)st * UM_BLOCK_FLOATS;
            const int k = w_kc * UM_KB + c16 * 4;
#pragma unroll
            for (int q = 0; q < RQ; ++q) {
                const int rl = q * RSTEP + rsub;
                float4 v = aop.finish(buf[q], brow[q], k, aux_a);
                if (!((vm >> q) & 1u)) v = make_float4(0.f, 0.f, 0.f, 0.f);
                v.x = to_tf32_fast(v.x); v.y = to_tf32_fast(v.y); v.z = to_tf32_fast(v.z); v.w = to_tf32_fast(v.w);
                *reinterpret_cast<float4*>(dst + (rl >> 3) * 256 + sw * 32 + ((c16 ^ sw) << 2)) = v;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[st]);
            if (++st == s.stages) { st = 0; ph ^= 1; }
            fetch_next(buf, brow, vm);
            if (++w_kc == s.KC) {
                w_kc = 0;
                if (warp == 0 && w_tile == (int)blockIdx.x) MFT_MARK(2);   // first tile written
                w_tile += gridDim.x;
            }
            return true;
        };
        for (;;) {
            bool go = true;
#pragma unroll
            for (int b = 0; b < NBUF; ++b) {
                if (go) go = step(raw[b], rrow[b], vmask[b]);
            }
            if (!go) break;
        }
        }   // !kTma
        if (warp == 0) MFT_MARK(3);                                        // producers done
    } else if (warp == UM_MMA_WARP) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            const uint32_t blk = (uint32_t)s.N_TILE * 128;
            mbar_wait(wbar, 0);                                            // issued in the prologue
            MFT_MARK(1);                                                   // weights resident
            const uint32_t idesc = make_idesc_tf32(UM_ROWS, s.N_TILE);
            const int ksteps = (s.K + 7) / 8;
            const uint32_t a0 = smem_u32(Asm), b0 = smem_u32(Wsm);
            int st = 0;
            uint32_t ph = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t aph = (it >> 1) & 1;
                mbar_wait(&tempty[acc], aph ^ 1);
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + acc * UM_ACC_STRIDE;
                for (int kc = 0; kc < s.KC; ++kc) {
                    mbar_wait(&full[st], ph);
                    tc_fence_after_sync();
                    const uint32_t ab = a0 + (uint32_t)st * (UM_BLOCK_FLOATS * 4);
                    const uint32_t bb = b0 + (uint32_t)kc * blk;
                    const int nks = min(4, ksteps - kc * 4);
                    for (int ks = 0; ks < nks; ++ks) {
                        mma_tf32_ss(d_tmem, make_desc_sw128(ab + ks * 32, 1024, 16),
                                    make_desc_sw128(bb + ks * 32, 1024, 16), idesc, (kc | ks) != 0 ? 1u : 0u);
                    }
                    mma_commit(&empty[st]);
                    if (++st == s.stages) { st = 0; ph ^= 1; }
                }
                mma_commit(&tfull[acc]);
                if (it == 0) MFT_MARK(4);                                  // first tile issued
            }
            MFT_MARK(5);                                                   // all MMAs issued
        }
        __syncwarp();
    } else if (warp == UM_TMA_WARP) {
        // ===================== TMA loader (one thread; kTma operands only) =====================
        if constexpr (tma_in_place<AOp>()) {
            if (lane == 0) {
                tma_prefetch_desc(&aop.tmap);
                tma_prefetch_desc(&aop.tmap_dy);
                int st = 0;
                uint32_t ph = 0;
                for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                    const int row0 = phys(tile) * UM_ROWS;
                    for (int kc = 0; kc < s.KC; ++kc) {
                        mbar_wait(&empty[st], ph ^ 1);           // the MMA has consumed the stage
                        mbar_arrive_expect_tx(&rawfull[st], UM_RAW_BYTES + UM_BLOCK_FLOATS * 4);
                        tma_load_2d(Asm + (size_t)st * UM_BLOCK_FLOATS, &aop.tmap_dy, kc * UM_KB, row0, &rawfull[st]);
                        tma_load_2d(rawring + (size_t)st * UM_RAW_BYTES, &aop.tmap, kc * UM_KB, row0, &rawfull[st]);
                        if (++st == s.stages) { st = 0; ph ^= 1; }
                    }
                }
            }
        } else if constexpr (AOp::kTma) {
            if (lane == 0) {
                tma_prefetch_desc(&aop.tmap);
                int rs = 0;
                uint32_t rph = 0;
                for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                    const int row0 = phys(tile) * UM_ROWS;
                    for (int kc = 0; kc < s.KC; ++kc) {
                        mbar_wait(&rawempty[rs], rph ^ 1);
                        mbar_arrive_expect_tx(&rawfull[rs], UM_RAW_BYTES);
                        tma_load_2d(rawring + (size_t)rs * UM_RAW_BYTES, &aop.tmap, kc * UM_KB, row0, &rawfull[rs]);
                        if (++rs == s.raw_stages) { rs = 0; rph ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue =====================
        reg_inc<192>();  // ... cover the 4 epilogue warps x 64 (the pool is per CTA)
        // Warp w drains TMEM lanes 32w..32w+31 = rows 32w.. of the tile and owns a private
        // [32][36] staging slab: TMEM -> registers (one row per lane) -> slab -> registers in
        // (row = i*4 + lane/8, 16-byte column = lane%8) order, so that every global access is four
        // full 128-byte row segments per instruction.  Only __syncwarp is needed.
        const int ew = warp & 3;
        float* slab = stage + ew * 32 * UM_STAGE_LD;
        const int rsub = lane >> 3, c4 = (lane & 7) * 4;
        const int nchunks = (s.N_TILE + 31) / 32;
        float s0[UM_STAT_CHUNKS][4], s1[UM_STAT_CHUNKS][4];
#pragma unroll
        for (int ch = 0; ch < UM_STAT_CHUNKS; ++ch)
#pragma unroll
            for (int e = 0; e < 4; ++e) { s0[ch][e] = 0.f; s1[ch][e] = 0.f; }
        int it = 0;
        float wq[8], wq_next[8];                              // row weights, fetched one tile ahead
        auto load_weights = [&](int tile, float (&dst)[8]) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int r = phys(min(tile, ntiles - 1)) * UM_ROWS + ew * 32 + q * 4 + rsub;
                dst[q] = Epi::kRowWeight ? epi.row_weight(tile < ntiles ? min(r, s.R) : s.R) : 0.f;   // pure loads
            }
        };
        load_weights(blockIdx.x, wq_next);
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            const int row0 = phys(tile) * UM_ROWS + ew * 32;  // first row of this warp's slab
#pragma unroll
            for (int q = 0; q < 8; ++q) wq[q] = wq_next[q];
            if (Epi::kRowWeight) load_weights(tile + gridDim.x, wq_next);
            // element offset of this thread's first column (chunk 0) in each of its rows (row clamped);
            // pinned in a register: left alone, the compiler rebuilds row * stride + column and reloads
            // the base pointer for every 16-byte access (9 of 23 instructions per access in the profile)
            unsigned roff[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                roff[q] = (unsigned)(min(row0 + q * 4 + rsub, s.R - 1) * epi.row_stride() + s.n0 + c4);
                asm volatile("" : "+r"(roff[q]));
            }
            mbar_wait(&tfull[acc], aph);
            tc_fence_after_sync();
            if (warp == UM_EPI_WARP0 && it == 0) MFT_MARK(12);             // first accumulator ready
            // Software pipeline over the 32-column chunks: the TMEM load of chunk ch+1 (and, for
            // epilogues that need a global operand, its loads) are issued as soon as chunk ch has
            // been copied to the slab, so they are in flight while chunk ch is written out.
            const uint32_t tbase = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * UM_ACC_STRIDE;
            uint32_t v[32];
            uint2 pre[Epi::kPrefetch ? 2 : 1][8];
            auto chunk_live = [&](int ch) { return ch * 32 + c4 < s.N_TILE && s.n0 + ch * 32 + c4 < s.N; };
            auto load_pre = [&](int ch, uint2 (&dst)[8]) {
                if (Epi::kPrefetch && chunk_live(ch)) {
#pragma unroll
                    for (int q = 0; q < 8; ++q)       // pure loads, row clamped in bounds
                        dst[q] = epi.prefetch(roff[q], ch * 32);
                }
            };
            tmem_ld_32x32(tbase, v);
            load_pre(0, pre[0]);
            long long t_ld = 0, t_sts = 0, t_apply = 0, tA = 0, tB = 0;
#ifdef MFT_UMMA_TIMING
            const bool timing = dbg != nullptr && warp == UM_EPI_WARP0 && lane == 0;
#else
            constexpr bool timing = false;   // phase counters (slots 13-15) only in -DMFT_UMMA_TIMING builds
#endif
#pragma unroll
            for (int ch = 0; ch < UM_MAX_CHUNKS; ++ch) {
                if (ch < nchunks) {
                    if (timing) tA = clock64();
                    tmem_ld_wait();
                    if (timing) { tB = clock64(); t_ld += tB - tA; }
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        *reinterpret_cast<float4*>(slab + lane * UM_STAGE_LD + q * 4) =
                            make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                        __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
                    __syncwarp();
                    if (timing) { tA = clock64(); t_sts += tA - tB; }
                    if (ch + 1 < nchunks) {
                        tmem_ld_32x32(tbase + (ch + 1) * 32, v);
                        load_pre(ch + 1, pre[Epi::kPrefetch ? ((ch + 1) & 1) : 0]);
                    }
                    const int cl = ch * 32 + c4;          // column inside this pass
                    const int col = s.n0 + cl;            // global output column
                    if (cl < s.N_TILE && col < s.N) {
                        const int nvalid = min(4, s.N - col);
                        const typename Epi::Consts ec = epi.consts(col, aux_e);
                        // straight-line over the 8 rows (no branch per row: rows past the end hold
                        // exact zeros in TMEM and only their store is predicated off), so the eight
                        // LDS / convert / store chains interleave
                        float4 a[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            a[q] = *reinterpret_cast<const float4*>(slab + (q * 4 + rsub) * UM_STAGE_LD + c4);
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const int r = row0 + q * 4 + rsub;
                            epi.apply(roff[q], ch * 32, r < s.R, wq[q], col, a[q], pre[Epi::kPrefetch ? (ch & 1) : 0][q],
                                      nvalid, s0[ch < UM_STAT_CHUNKS ? ch : 0], s1[ch < UM_STAT_CHUNKS ? ch : 0],
                                      ec);
                        }
                    }
                    __syncwarp();
                    if (timing) { tB = clock64(); t_apply += tB - tA; }
                }
            }
            if (timing) { dbg[13] += t_ld; dbg[14] += t_sts; dbg[15] += t_apply; }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (warp == UM_EPI_WARP0 && it < 4) MFT_MARK(8 + it);          // epilogue finished tile `it`
        }
        if (warp == UM_EPI_WARP0) MFT_MARK(6);                             // epilogue tiles done
        if (Epi::kStats) {
            // lanes sharing a column (lane, lane^8, lane^16, lane^24) combine by shuffle; each warp then
            // owns one row of the [4][256] partials (red0/red1 alias the idle slab) -- no smem atomics
            float* part0 = stage;                       // [4][256]
            float* part1 = stage + 4 * 256;             // [4][256]
            named_bar_sync(1, 128);                     // every epilogue warp is done with its slab
#pragma unroll
            for (int ch = 0; ch < UM_STAT_CHUNKS; ++ch) {
                if (ch < nchunks) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float a0 = s0[ch][e], a1 = s1[ch][e];
                        a0 += __shfl_xor_sync(0xffffffffu, a0, 8);
                        a1 += __shfl_xor_sync(0xffffffffu, a1, 8);
                        a0 += __shfl_xor_sync(0xffffffffu, a0, 16);
                        a1 += __shfl_xor_sync(0xffffffffu, a1, 16);
                        const int cl = ch * 32 + c4 + e;
                        if (lane < 8 && cl < 256) {
                            part0[ew * 256 + cl] = a0;
                            part1[ew * 256 + cl] = a1;
                        }
                    }
                }
            }
            named_bar_sync(1, 128);
            for (int cl = tid - UM_EPI_WARP0 * 32; cl < s.N_TILE; cl += 128)
                if (s.n0 + cl < s.N)
                    epi.commit(s.n0 + cl, part0[cl] + part0[256 + cl] + part0[512 + cl] + part0[768 + cl],
                               part1[cl] + part1[256 + cl] + part1[512 + cl] + part1[768 + cl], aux_e);
        }
    }

    if (warp == UM_EPI_WARP0) MFT_MARK(7);                                 // statistics committed
#undef MFT_MARK
    tc_fence_before_sync();
    __syncthreads();
    if (warp == UM_MMA_WARP) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, UM_TMEM_COLS);
    }
}

// ------------------------------------------------------------------ the rows kernel with A in TENSOR MEMORY
// What paces umma_rows_kernel is shared-memory bandwidth (128 B / clock / SM), not the roles' instruction
// streams: one M=128, N=192, K=8 kind::tf32 MMA reads 4 KB of A and 6 KB of B from shared memory in its 96
// cycles, i.e. 107 B / clock by itself; per 128 x 192 x 192 tile the MMA reads 245 KB, the producers write the
// 98 KB A tile (after reading 48 KB of raw tape that the TMA wrote), and the epilogue pushes 196 KB through its
// transposing slab: 635 KB = 5.0k cycles of a 7.3k-cycle tile (profiles/r02_summary.md; neither deeper rings,
// nor producer groups, nor CTA pairs moved it).  Here the A operand never touches shared memory: a producer
// thread owns one ROW of the tile, builds its 32 K-values of a block in registers and writes them to tensor
// memory (tcgen05.st.32x32b), and the MMA takes A from there ([a_tmem] operand form).  Saves the 98 KB of A
// stores and the 98 KB of A operand reads per tile (31 % of the traffic) and the A ring's 32-64 KB of shared memory.
//
// Tensor memory (512 columns): accumulator a at column 256 a (N_TILE <= 224 columns), A slots of 32 columns
// (one K block of 32 tf32 words for all 128 lanes) packed downwards from the top of each half:
// NS = 2 * floor((256 - N_TILE) / 32) <= 8 slots (4 at N = 192).  Producer warps 0-3 and 4-7 form two groups
// (warp % 4 = tensor-memory lane quadrant = rows 32 (warp % 4) ..); group g builds K blocks g, g + 2, ...
// Barriers: rawfull / rawempty per raw slot (TMA <-> producers), full / empty per A slot (producers <-> MMA),
// tfull / tempty per accumulator (MMA <-> epilogue).
constexpr int UM_TS_MAX_SLOTS = 8;
constexpr int UM_TS_MAX_NTILE = 224;

__host__ __device__ inline int ts_slots(int N_TILE) {
    const int n = 2 * ((256 - N_TILE) / 32);
    return n > UM_TS_MAX_SLOTS ? UM_TS_MAX_SLOTS : n;
}
__device__ __forceinline__ uint32_t ts_slot_col(int slot) {      // slot j: half j & 1, (j >> 1)-th from the top
    return (uint32_t)((slot & 1) * 256 + 256 - 32 * ((slot >> 1) + 1));
}

template <class AOp> __host__ __device__ constexpr int ts_raw_bytes() {
    if constexpr (tma_in_place<AOp>()) return UM_BLOCK_FLOATS * 4 + UM_RAW_BYTES;   // dy fp32 block + H fp16 block
    else if constexpr (AOp::kTma) return UM_RAW_BYTES;
    else return 0;
}

template <class AOp>
static inline size_t umma_smem_bytes_ts(const UmmaShape& s, int epi_warps = 4) {
    return 1024 + (size_t)s.KC * s.N_TILE * 128 + (size_t)epi_warps * 32 * UM_STAGE_LD * 4 + (3 + 4) * kMaxC * 4 +
           (size_t)s.raw_stages * ts_raw_bytes<AOp>() + 48 * 8 + 16;
}

// EW = 4 or 8 epilogue warps.  With eight, the two warps of a lane quadrant take alternate 32-column chunks of
// the accumulator (each has its own staging slab): the source-level profile shows the epilogue warp alone on
// its scheduler, stalled on fixed-latency dependencies a third of the time and issuing a fifth of it, and the
// whole kernel paced by it -- a second warp per scheduler fills those slots.
template <class AOp, class Epi, int EW>
__global__ void __launch_bounds__((UM_PROD_WARPS + EW + 2) * 32, 1)
umma_rows_ts_kernel(const __grid_constant__ AOp aop, const __grid_constant__ Epi epi, const float* __restrict__ wimg,
                    const __grid_constant__ UmmaShape s) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t w_bytes = (uint32_t)s.KC * s.N_TILE * 128;
    constexpr int kRawBytes = ts_raw_bytes<AOp>();
    constexpr int kThreads = (UM_PROD_WARPS + EW + 2) * 32;
    constexpr int kMmaWarp = UM_PROD_WARPS + EW, kTmaWarp = kMmaWarp + 1;
    float* Wsm = reinterpret_cast<float*>(smem);
    float* stage = reinterpret_cast<float*>(smem + w_bytes);
    float* aux_a = stage + EW * 32 * UM_STAGE_LD;
    float* aux_e = aux_a + 3 * kMaxC;
    uint8_t* rawring = reinterpret_cast<uint8_t*>(aux_e + 4 * kMaxC);
    uint64_t* bars = reinterpret_cast<uint64_t*>(rawring + (size_t)s.raw_stages * kRawBytes);
    uint64_t* full = bars;            // [8] A slot written
    uint64_t* empty = bars + 8;       // [8] A slot consumed by the MMA
    uint64_t* tfull = bars + 16;      // [2]
    uint64_t* tempty = bars + 18;     // [2]
    uint64_t* wbar = bars + 20;
    uint64_t* rawfull = bars + 22;    // [8]
    uint64_t* rawempty = bars + 30;   // [8]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 38);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int ntiles = (s.R + UM_ROWS - 1) / UM_ROWS;
    auto phys = [&](int t) { return s.reverse ? ntiles - 1 - t : t; };
    const int my_tiles = (int)blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int NS = s.stages;          // A slots in tensor memory

    if (tid == 0) {
        for (int i = 0; i < UM_TS_MAX_SLOTS; ++i) {
            mbar_init(&full[i], 4);                 // the four warps (lane quadrants) of a producer group
            mbar_init(&empty[i], 1);
            mbar_init(&rawfull[i], 1);
            mbar_init(&rawempty[i], 4);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], EW);
        }
        mbar_init(wbar, 1);
        fence_mbar_init();
    }
    if (warp == kMmaWarp) tmem_alloc(tmem_slot, UM_TMEM_COLS);
    __syncthreads();
    if (warp == kMmaWarp && lane == 0) {
        mbar_arrive_expect_tx(wbar, w_bytes);
        const uint32_t blk = (uint32_t)s.N_TILE * 128;
        for (int kc = 0; kc < s.KC; ++kc)
            bulk_g2s(reinterpret_cast<uint8_t*>(Wsm) + (size_t)kc * blk,
                     reinterpret_cast<const uint8_t*>(wimg) + (size_t)kc * blk, blk, wbar);
    }
    pdl_wait();
    pdl_launch_dependents();
    aop.init(aux_a, tid, kThreads);
    epi.init(aux_e, tid, kThreads);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    long long* dbg = s.dbg ? s.dbg + (size_t)blockIdx.x * 16 : nullptr;   // clock64 timeline (tools/umma_timeline.py)
#define MFT_MARK(slot) do { if (dbg && lane == 0) dbg[slot] = clock64(); } while (0)
    if (dbg && tid == 0) dbg[0] = clock64();

    if (warp < UM_PROD_WARPS) {
        // ===================== producers: one row per thread, 32 K-values per block, into tensor memory
        // Register pool of the CTA = threads x the kernel's register count: 448 x 128 with four epilogue warps
        // (producers 96, epilogue 192), 576 x 96 with eight (producers 72, epilogue 120, the MMA / TMA warps keep
        // their 96 -- they are half a warpgroup and setmaxnreg is a warpgroup instruction: 18432 + 30720 + 6144 =
        // 55296; setmaxnreg.inc blocks until the CTA's own warps have released enough).
        if constexpr (EW == 8) reg_dec<72>(); else reg_dec<96>();
        const int grp = warp >> 2, quad = warp & 3;
        const int rl = quad * 32 + lane;                       // row of the tile = tensor-memory lane
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        const int nblocks = my_tiles * s.KC;
        int cur_ti = -1;
        bool ok = false;
        [[maybe_unused]] float wrow = 0.f;
        [[maybe_unused]] typename AOp::Row rrow{};
        for (int b = grp; b < nblocks; b += 2) {
            const int ti = b / s.KC, kc = b - ti * s.KC;
            const int row0 = phys((int)blockIdx.x + ti * (int)gridDim.x) * UM_ROWS;
            if (ti != cur_ti) {
                cur_ti = ti;
                ok = row0 + rl < s.R;
                if constexpr (tma_in_place<AOp>()) wrow = aop.row_weight(min(row0 + rl, s.R));
                if constexpr (!AOp::kTma) rrow = aop.row(ok ? row0 + rl : 0);
            }
            const int slot = b % NS;
            const uint32_t ph = (uint32_t)(b / NS) & 1u;
            const int k0 = kc * UM_KB;
            uint32_t v[32];
            if constexpr (tma_in_place<AOp>()) {
                const int rs = b % s.raw_stages;
                const uint32_t rph = (uint32_t)(b / s.raw_stages) & 1u;
                mbar_wait(&rawfull[rs], rph);
                const uint8_t* rb = rawring + (size_t)rs * kRawBytes;
                // dy: [128 rows x 32 fp32], SWIZZLE_128B as the TMA wrote it; H: [128 rows x 32 fp16], plain
                const float* dyrow = reinterpret_cast<const float*>(rb) + (rl >> 3) * 256 + (rl & 7) * 32;
                const uint4* hrow = reinterpret_cast<const uint4*>(rb + UM_BLOCK_FLOATS * 4 + rl * 64);
                float4 d[8];
                uint4 h[4];
#pragma unroll
                for (int j = 0; j < 8; ++j) d[j] = *reinterpret_cast<const float4*>(dyrow + ((j ^ (rl & 7)) << 2));
#pragma unroll
                for (int j = 0; j < 4; ++j) h[j] = hrow[j];
                {   // release the raw slot on COMPLETION of the loads (see umma_rows_kernel)
                    unsigned t = 0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) t |= __float_as_uint(d[j].x) | __float_as_uint(d[j].w);
#pragma unroll
                    for (int j = 0; j < 4; ++j) t |= h[j].x | h[j].w;
                    const unsigned z = __reduce_or_sync(0xffffffffu, t & s.zero);
                    if (lane == 0) mbar_arrive_n(&rawempty[rs], 1u + z);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const typename AOp::Consts c = aop.consts(k0 + 4 * j, aux_a);
                    const uint2 hp = (j & 1) ? make_uint2(h[j >> 1].z, h[j >> 1].w) : make_uint2(h[j >> 1].x, h[j >> 1].y);
                    float4 o = aop.transform2(d[j], hp, wrow, c);
                    if (!ok) o = make_float4(0.f, 0.f, 0.f, 0.f);
                    v[4 * j] = __float_as_uint(to_tf32_fast(o.x)); v[4 * j + 1] = __float_as_uint(to_tf32_fast(o.y));
                    v[4 * j + 2] = __float_as_uint(to_tf32_fast(o.z)); v[4 * j + 3] = __float_as_uint(to_tf32_fast(o.w));
                }
            } else if constexpr (AOp::kTma) {
                const int rs = b % s.raw_stages;
                const uint32_t rph = (uint32_t)(b / s.raw_stages) & 1u;
                mbar_wait(&rawfull[rs], rph);
                const uint4* hrow = reinterpret_cast<const uint4*>(rawring + (size_t)rs * kRawBytes + rl * 64);
                uint4 h[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) h[j] = hrow[j];
                {
                    unsigned t = 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) t |= h[j].x | h[j].y | h[j].z | h[j].w;
                    const unsigned z = __reduce_or_sync(0xffffffffu, t & s.zero);
                    if (lane == 0) mbar_arrive_n(&rawempty[rs], 1u + z);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const typename AOp::Consts c = aop.consts(k0 + 4 * j, aux_a);
                    const uint2 hp = (j & 1) ? make_uint2(h[j >> 1].z, h[j >> 1].w) : make_uint2(h[j >> 1].x, h[j >> 1].y);
                    float4 o = aop.transform(hp, c);
                    if (!ok) o = make_float4(0.f, 0.f, 0.f, 0.f);
                    v[4 * j] = __float_as_uint(to_tf32_fast(o.x)); v[4 * j + 1] = __float_as_uint(to_tf32_fast(o.y));
                    v[4 * j + 2] = __float_as_uint(to_tf32_fast(o.z)); v[4 * j + 3] = __float_as_uint(to_tf32_fast(o.w));
                }
            } else {
#pragma unroll
                for (int hq = 0; hq < 2; ++hq) {                // two batches of loads: 32 registers in flight
                    typename AOp::Raw raw[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) aop.fetch(rrow, k0 + 4 * (4 * hq + j), raw[j]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int jj = 4 * hq + j;
                        float4 o = aop.finish(raw[j], rrow, k0 + 4 * jj, aux_a);
                        if (!ok) o = make_float4(0.f, 0.f, 0.f, 0.f);
                        v[4 * jj] = __float_as_uint(to_tf32_fast(o.x)); v[4 * jj + 1] = __float_as_uint(to_tf32_fast(o.y));
                        v[4 * jj + 2] = __float_as_uint(to_tf32_fast(o.z)); v[4 * jj + 3] = __float_as_uint(to_tf32_fast(o.w));
                    }
                }
            }
            mbar_wait(&empty[slot], ph ^ 1);                    // the MMAs that read this slot last have completed
            tc_fence_after_sync();
            tmem_st_32x32(tmem_base + lane_base + ts_slot_col(slot), v);
            tmem_st_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[slot]);
            if (warp == 0 && ti == 0 && kc + 2 >= s.KC) MFT_MARK(2);          // first tile written
        }
        if (warp == 0) MFT_MARK(3);                                           // producers done
    } else if (warp == kMmaWarp) {
        // ===================== MMA issuer (one thread): A from tensor memory, B = resident weight image
        if (lane == 0) {
            const uint32_t blk = (uint32_t)s.N_TILE * 128;
            mbar_wait(wbar, 0);
            MFT_MARK(1);                                                   // weights resident
            const uint32_t idesc = make_idesc_tf32(UM_ROWS, s.N_TILE);
            const int ksteps = (s.K + 7) / 8;
            const uint32_t b0 = smem_u32(Wsm);
            int b = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const int acc = it & 1;
                const uint32_t aph = (it >> 1) & 1;
                mbar_wait(&tempty[acc], aph ^ 1);
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + acc * UM_ACC_STRIDE;
                for (int kc = 0; kc < s.KC; ++kc, ++b) {
                    const int slot = b % NS;
                    const uint32_t ph = (uint32_t)(b / NS) & 1u;
                    mbar_wait(&full[slot], ph);
                    tc_fence_after_sync();
                    const uint32_t a_tmem = tmem_base + ts_slot_col(slot);
                    const uint32_t bb = b0 + (uint32_t)kc * blk;
                    const int nks = min(4, ksteps - kc * 4);
                    for (int ks = 0; ks < nks; ++ks)
                        mma_tf32_ts(d_tmem, a_tmem + ks * 8, make_desc_sw128(bb + ks * 32, 1024, 16), idesc,
                                    (kc | ks) != 0 ? 1u : 0u);
                    mma_commit(&empty[slot]);
                }
                mma_commit(&tfull[acc]);
                if (it == 0) MFT_MARK(4);                                  // first tile issued
            }
            MFT_MARK(5);                                                   // all MMAs issued
        }
        __syncwarp();
    } else if (warp == kTmaWarp) {
        // ===================== TMA loader (one thread): raw blocks of the tape / gradient operands
        if constexpr (AOp::kTma) {
            if (lane == 0) {
                tma_prefetch_desc(&aop.tmap);
                if constexpr (tma_in_place<AOp>()) tma_prefetch_desc(&aop.tmap_dy);
                int b = 0;
                for (int it = 0; it < my_tiles; ++it) {
                    const int row0 = phys((int)blockIdx.x + it * (int)gridDim.x) * UM_ROWS;
                    for (int kc = 0; kc < s.KC; ++kc, ++b) {
                        const int rs = b % s.raw_stages;
                        const uint32_t rph = (uint32_t)(b / s.raw_stages) & 1u;
                        mbar_wait(&rawempty[rs], rph ^ 1);
                        uint8_t* rb = rawring + (size_t)rs * kRawBytes;
                        mbar_arrive_expect_tx(&rawfull[rs], kRawBytes);
                        if constexpr (tma_in_place<AOp>()) {
                            tma_load_2d(rb, &aop.tmap_dy, kc * UM_KB, row0, &rawfull[rs]);
                            tma_load_2d(rb + UM_BLOCK_FLOATS * 4, &aop.tmap, kc * UM_KB, row0, &rawfull[rs]);
                        } else {
                            tma_load_2d(rb, &aop.tmap, kc * UM_KB, row0, &rawfull[rs]);
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue: EW warps; warp (quadrant ew, half hsel) owns chunks ch % (EW/4) == hsel
        if constexpr (EW == 8) reg_inc<120>(); else reg_inc<192>();
        constexpr int kSplit = EW / 4;                            // warps per lane quadrant
        constexpr int kOwn = (UM_STAT_CHUNKS + kSplit) / kSplit;  // chunks a warp may own (N_TILE <= 224: 7 chunks)
        const int ewi = warp - UM_EPI_WARP0;
        const int ew = ewi & 3, hsel = ewi >> 2;
        float* slab = stage + ewi * 32 * UM_STAGE_LD;
        const int rsub = lane >> 3, c4 = (lane & 7) * 4;
        const int nchunks = (s.N_TILE + 31) / 32;
        float s0[kOwn][4], s1[kOwn][4];
#pragma unroll
        for (int o = 0; o < kOwn; ++o)
#pragma unroll
            for (int e = 0; e < 4; ++e) { s0[o][e] = 0.f; s1[o][e] = 0.f; }
        float wq[8], wq_next[8];
        auto load_weights = [&](int it, float (&dst)[8]) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int tile = (int)blockIdx.x + min(it, max(my_tiles - 1, 0)) * (int)gridDim.x;
                const int r = phys(min(tile, ntiles - 1)) * UM_ROWS + ew * 32 + q * 4 + rsub;
                dst[q] = Epi::kRowWeight ? epi.row_weight(it < my_tiles ? min(r, s.R) : s.R) : 0.f;
            }
        };
        load_weights(0, wq_next);
        for (int it = 0; it < my_tiles; ++it) {
            const int acc = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            const int row0 = phys((int)blockIdx.x + it * (int)gridDim.x) * UM_ROWS + ew * 32;
#pragma unroll
            for (int q = 0; q < 8; ++q) wq[q] = wq_next[q];
            if (Epi::kRowWeight) load_weights(it + 1, wq_next);
            unsigned roff[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                roff[q] = (unsigned)(min(row0 + q * 4 + rsub, s.R - 1) * epi.row_stride() + s.n0 + c4);
                asm volatile("" : "+r"(roff[q]));
            }
            mbar_wait(&tfull[acc], aph);
            tc_fence_after_sync();
            if (ewi == 0 && it == 0) MFT_MARK(12);                         // first accumulator ready
            const uint32_t tbase = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * UM_ACC_STRIDE;
            uint32_t v[32];
            uint2 pre[Epi::kPrefetch ? 2 : 1][8];
            auto chunk_live = [&](int ch) { return ch * 32 + c4 < s.N_TILE && s.n0 + ch * 32 + c4 < s.N; };
            auto load_pre = [&](int ch, uint2 (&dst)[8]) {
                if (Epi::kPrefetch && chunk_live(ch)) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) dst[q] = epi.prefetch(roff[q], ch * 32);
                }
            };
            if (hsel < nchunks) {
                tmem_ld_32x32(tbase + hsel * 32, v);
                load_pre(hsel, pre[0]);
            }
#pragma unroll
            for (int o = 0; o < kOwn; ++o) {
                const int ch = o * kSplit + hsel;                 // this warp's o-th chunk
                if (ch < nchunks) {
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        *reinterpret_cast<float4*>(slab + lane * UM_STAGE_LD + q * 4) =
                            make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                        __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
                    __syncwarp();
                    if (ch + kSplit < nchunks) {
                        tmem_ld_32x32(tbase + (ch + kSplit) * 32, v);
                        load_pre(ch + kSplit, pre[Epi::kPrefetch ? ((o + 1) & 1) : 0]);
                    }
                    const int cl = ch * 32 + c4;
                    const int col = s.n0 + cl;
                    if (cl < s.N_TILE && col < s.N) {
                        const int nvalid = min(4, s.N - col);
                        const typename Epi::Consts ec = epi.consts(col, aux_e);
                        float4 a[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            a[q] = *reinterpret_cast<const float4*>(slab + (q * 4 + rsub) * UM_STAGE_LD + c4);
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const int r = row0 + q * 4 + rsub;
                            epi.apply(roff[q], ch * 32, r < s.R, wq[q], col, a[q], pre[Epi::kPrefetch ? (o & 1) : 0][q],
                                      nvalid, s0[o], s1[o], ec);
                        }
                    }
                    __syncwarp();
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (ewi == 0 && it < 4) MFT_MARK(8 + it);                      // epilogue finished tile `it`
        }
        if (ewi == 0) MFT_MARK(6);
        if (Epi::kStats) {
            // every (quadrant, column) pair is owned by exactly one warp: [4][256] partials, no atomics
            float* part0 = stage;
            float* part1 = stage + 4 * 256;
            named_bar_sync(1, EW * 32);                  // every epilogue warp is done with its slab
#pragma unroll
            for (int o = 0; o < kOwn; ++o) {
                const int ch = o * kSplit + hsel;
                if (ch < nchunks) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float a0 = s0[o][e], a1 = s1[o][e];
                        a0 += __shfl_xor_sync(0xffffffffu, a0, 8);
                        a1 += __shfl_xor_sync(0xffffffffu, a1, 8);
                        a0 += __shfl_xor_sync(0xffffffffu, a0, 16);
                        a1 += __shfl_xor_sync(0xffffffffu, a1, 16);
                        const int cl = ch * 32 + c4 + e;
                        if (lane < 8 && cl < 256) {
                            part0[ew * 256 + cl] = a0;
                            part1[ew * 256 + cl] = a1;
                        }
                    }
                }
            }
            named_bar_sync(1, EW * 32);
            for (int cl = tid - UM_EPI_WARP0 * 32; cl < s.N_TILE; cl += EW * 32)
                if (s.n0 + cl < s.N)
                    epi.commit(s.n0 + cl, part0[cl] + part0[256 + cl] + part0[512 + cl] + part0[768 + cl],
                               part1[cl] + part1[256 + cl] + part1[512 + cl] + part1[768 + cl], aux_e);
        }
    }

    if (warp == UM_EPI_WARP0) MFT_MARK(7);
#undef MFT_MARK
    tc_fence_before_sync();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, UM_TMEM_COLS);
    }
}

// ------------------------------------------------------------------ the rows kernel on CTA pairs
// Same GEMM, same roles, same operand / epilogue functors as umma_rows_kernel, but the two CTAs of a cluster
// share every tcgen05.mma (cta_group::2, M = 256: 128 rows of A per CTA, HALF of the weight rows per CTA).
// What that buys is shared memory: the resident weight image -- 147 KB of the 227 KB for the 192 x 192 layers,
// the reason those kernels had two A stages and the in-place dgrad 48 KB of loads in flight per SM -- halves, and
// the freed 74 KB go into deeper rings (dgrad: five 24 KB stages in flight instead of two: that kernel is bound
// by bytes in flight x HBM latency, profiles/r02_summary.md) and into producer GROUPS: the eight producer warps
// form PG groups, group g builds K blocks g, g + PG, ... so that PG of the serial wait / LDS / transform / STS /
// fence / arrive chains overlap (PG <= ring depth: a group may not run two barrier phases ahead).
//
// Protocol differences: `full` and `tempty` live in the LEADER's shared memory and count arrivals of both CTAs
// (the peer arrives through the cluster window, release / acquire at cluster scope); `empty` and `tfull` are
// per CTA and are signalled by the multicast commit; each CTA loads its weight half itself, the peer reports it
// on `wpeer`.  Both CTAs walk the same number of 256-row super-tiles; a 128-row half past the end is produced
// as zeros and stores nothing.
template <class AOp, class Epi, int PG>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(UM_THREADS, 1)
umma_rows_pair_kernel(const __grid_constant__ AOp aop, const __grid_constant__ Epi epi, const float* __restrict__ wimg,
                      const __grid_constant__ UmmaShape s) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int NH = s.N_TILE / 2;                              // weight rows held by this CTA
    const uint32_t w_bytes = (uint32_t)s.KC * NH * 128;
    float* Wsm = reinterpret_cast<float*>(smem);
    float* Asm = reinterpret_cast<float*>(smem + w_bytes);
    float* stage = Asm + (size_t)s.stages * UM_BLOCK_FLOATS;
    float* aux_a = stage + UM_ROWS * UM_STAGE_LD;
    float* aux_e = aux_a + 3 * kMaxC;
    uint8_t* rawring = reinterpret_cast<uint8_t*>(aux_e + 4 * kMaxC);
    uint64_t* bars = reinterpret_cast<uint64_t*>(rawring + (size_t)s.raw_stages * UM_RAW_BYTES);
    uint64_t* full = bars;            // [UM_MAX_PSTAGES]  leader's copy is the live one
    uint64_t* empty = bars + 8;       // [UM_MAX_PSTAGES]  per CTA
    uint64_t* tfull = bars + 16;      // [2] per CTA
    uint64_t* tempty = bars + 18;     // [2] leader's copy is the live one
    uint64_t* wbar = bars + 20;       // this CTA's weight half
    uint64_t* wpeer = bars + 21;      // leader: the peer's weight half has landed
    uint64_t* rawfull = bars + 22;    // [UM_MAX_RAW]
    uint64_t* rawempty = bars + 30;   // [UM_MAX_RAW]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 38);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int nst = (s.R + 2 * UM_ROWS - 1) / (2 * UM_ROWS);          // 256-row super-tiles
    const int my_tiles = pair < nst ? (nst - 1 - pair) / npairs + 1 : 0;
    // first row of this CTA's half of the pair's i-th super-tile (serpentine direction as in umma_rows_kernel)
    auto tile_row0 = [&](int i) {
        const int t = 2 * (pair + i * npairs) + (int)rank;
        return (s.reverse ? 2 * nst - 1 - t : t) * UM_ROWS;
    };
    constexpr int GW = UM_PROD_WARPS / PG;

    if (tid == 0) {
        for (int i = 0; i < UM_MAX_PSTAGES; ++i) {
            mbar_init(&full[i], 2 * GW);            // one elected arrival per producer warp of the group, both CTAs
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < UM_MAX_RAW; ++i) {
            mbar_init(&rawfull[i], 1);
            mbar_init(&rawempty[i], GW);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], 8);               // four epilogue warps in each CTA
        }
        mbar_init(wbar, 1);
        mbar_init(wpeer, 1);
        fence_mbar_init();
    }
    if (warp == UM_MMA_WARP) tmem_alloc2(tmem_slot, UM_TMEM_COLS);
    tc_fence_before_sync();
    cluster_sync_all();                                 // barrier inits of BOTH CTAs visible before any remote arrive
    tc_fence_after_sync();
    if (warp == UM_MMA_WARP && lane == 0) {
        mbar_arrive_expect_tx(wbar, w_bytes);
        const uint32_t blk = (uint32_t)NH * 128;         // this CTA's rows of one K block: a contiguous range of the image
        for (int kc = 0; kc < s.KC; ++kc)
            bulk_g2s(reinterpret_cast<uint8_t*>(Wsm) + (size_t)kc * blk,
                     reinterpret_cast<const uint8_t*>(wimg) + ((size_t)kc * s.N_TILE + (size_t)rank * NH) * 128, blk, wbar);
    }
    pdl_wait();
    pdl_launch_dependents();
    aop.init(aux_a, tid, UM_THREADS);
    epi.init(aux_e, tid, UM_THREADS);
    __syncthreads();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < UM_PROD_WARPS) {
        // ===================== producers (PG groups, see the header comment) =====================
        reg_dec<96>();
        constexpr int GT = UM_PROD_THREADS / PG;
        constexpr int RQ = UM_ROWS * 8 / GT;
        constexpr int RSTEP = GT / 8;
        const int grp = tid / GT, gtid = tid - grp * GT;
        const int rsub = gtid >> 3, c16 = gtid & 7;
        const int sw = rsub & 7;
        const int sw_off = sw * 32 + ((c16 ^ sw) << 2);
        const int nblocks = my_tiles * s.KC;
        const uint32_t full0 = mapa_shared(full, 0);         // the leader's `full` array in the cluster window
        if constexpr (tma_in_place<AOp>()) {
            constexpr int QB = RQ < 8 ? RQ : 8;
            float wq[RQ];
            int cur_ti = -1;
            for (int b = grp; b < nblocks; b += PG) {
                const int ti = b / s.KC, kc = b - ti * s.KC;
                const int row0 = tile_row0(ti);
                const bool partial = row0 + UM_ROWS > s.R;
                if (ti != cur_ti) {
                    cur_ti = ti;
#pragma unroll
                    for (int q = 0; q < RQ; ++q) wq[q] = aop.row_weight(min(row0 + q * RSTEP + rsub, s.R));
                }
                const int st = b % s.stages;
                const uint32_t ph = (uint32_t)(b / s.stages) & 1u;
                mbar_wait(&rawfull[st], ph);
                const uint8_t* rawb = rawring + (size_t)st * UM_RAW_BYTES;
                float* blk = Asm + (size_t)st * UM_BLOCK_FLOATS;
                const int k = kc * UM_KB + c16 * 4;
                const typename AOp::Consts kc4 = aop.consts(k, aux_a);
#pragma unroll
                for (int q0 = 0; q0 < RQ; q0 += QB) {
                    uint2 hraw[QB];
                    float4 d[QB];
#pragma unroll
                    for (int q = 0; q < QB; ++q) {
                        const int rl = (q0 + q) * RSTEP + rsub;
                        hraw[q] = *reinterpret_cast<const uint2*>(rawb + rl * 64 + c16 * 8);
                        d[q] = *reinterpret_cast<const float4*>(blk + (rl >> 3) * 256 + sw_off);
                    }
#pragma unroll
                    for (int q = 0; q < QB; ++q) {
                        const int rl = (q0 + q) * RSTEP + rsub;
                        float4 v = aop.transform2(d[q], hraw[q], wq[q0 + q], kc4);
                        if (partial && row0 + rl >= s.R) v = make_float4(0.f, 0.f, 0.f, 0.f);
                        v.x = to_tf32_fast(v.x); v.y = to_tf32_fast(v.y); v.z = to_tf32_fast(v.z); v.w = to_tf32_fast(v.w);
                        *reinterpret_cast<float4*>(blk + (rl >> 3) * 256 + sw_off) = v;
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(full0 + (uint32_t)st * 8u);
            }
        } else if constexpr (AOp::kTma) {
            for (int b = grp; b < nblocks; b += PG) {
                const int ti = b / s.KC, kc = b - ti * s.KC;
                const int row0 = tile_row0(ti);
                const bool partial = row0 + UM_ROWS > s.R;
                const int rs = b % s.raw_stages;
                const uint32_t rph = (uint32_t)(b / s.raw_stages) & 1u;
                const int st = b % s.stages;
                const uint32_t ph = (uint32_t)(b / s.stages) & 1u;
                mbar_wait(&rawfull[rs], rph);
                const uint8_t* rawb = rawring + (size_t)rs * UM_RAW_BYTES;
                uint2 raw[RQ];
#pragma unroll
                for (int q = 0; q < RQ; ++q)
                    raw[q] = *reinterpret_cast<const uint2*>(rawb + (q * RSTEP + rsub) * 64 + c16 * 8);
                {   // release the raw slot on COMPLETION of the loads (see umma_rows_kernel)
                    unsigned t = 0;
#pragma unroll
                    for (int q = 0; q < RQ; ++q) t |= raw[q].x | raw[q].y;
                    const unsigned z = __reduce_or_sync(0xffffffffu, t & s.zero);
                    if (lane == 0) mbar_arrive_n(&rawempty[rs], 1u + z);
                }
                mbar_wait(&empty[st], ph ^ 1);
                float* blk = Asm + (size_t)st * UM_BLOCK_FLOATS;
                const int k = kc * UM_KB + c16 * 4;
                const typename AOp::Consts kc4 = aop.consts(k, aux_a);
#pragma unroll
                for (int q = 0; q < RQ; ++q) {
                    const int rl = q * RSTEP + rsub;
                    float4 v = aop.transform(raw[q], kc4);
                    if (partial && row0 + rl >= s.R) v = make_float4(0.f, 0.f, 0.f, 0.f);
                    v.x = to_tf32_fast(v.x); v.y = to_tf32_fast(v.y); v.z = to_tf32_fast(v.z); v.w = to_tf32_fast(v.w);
                    *reinterpret_cast<float4*>(blk + (rl >> 3) * 256 + sw_off) = v;
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(full0 + (uint32_t)st * 8u);
            }
        } else {
            constexpr int QB = RQ <= 4 ? RQ : 2;
            constexpr int NB = RQ / QB;
            typename AOp::Row rc[RQ];
            uint32_t vrows = 0;
            int cur_ti = -1;
            for (int b = grp; b < nblocks; b += PG) {
                const int ti = b / s.KC, kc = b - ti * s.KC;
                const int row0 = tile_row0(ti);
                if (ti != cur_ti) {
                    cur_ti = ti;
                    vrows = 0;
#pragma unroll
                    for (int q = 0; q < RQ; ++q) {
                        const int r = row0 + q * RSTEP + rsub;
                        const bool ok = r < s.R;
                        vrows |= (ok ? 1u : 0u) << q;
                        rc[q] = aop.row(ok ? r : 0);
                    }
                }
                const int st = b % s.stages;
                const uint32_t ph = (uint32_t)(b / s.stages) & 1u;
                const int k = kc * UM_KB + c16 * 4;
                typename AOp::Raw raw[2][QB];
#pragma unroll
                for (int q = 0; q < QB; ++q) aop.fetch(rc[q], k, raw[0][q]);
                mbar_wait(&empty[st], ph ^ 1);
                float* blk = Asm + (size_t)st * UM_BLOCK_FLOATS;
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    if (i + 1 < NB) {
#pragma unroll
                        for (int q = 0; q < QB; ++q) aop.fetch(rc[(i + 1) * QB + q], k, raw[(i + 1) & 1][q]);
                    }
#pragma unroll
                    for (int q = 0; q < QB; ++q) {
                        const int qq = i * QB + q;
                        const int rl = qq * RSTEP + rsub;
                        float4 v = aop.finish(raw[i & 1][q], rc[qq], k, aux_a);
                        if (!((vrows >> qq) & 1u)) v = make_float4(0.f, 0.f, 0.f, 0.f);
                        v.x = to_tf32_fast(v.x); v.y = to_tf32_fast(v.y); v.z = to_tf32_fast(v.z); v.w = to_tf32_fast(v.w);
                        *reinterpret_cast<float4*>(blk + (rl >> 3) * 256 + sw_off) = v;
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(full0 + (uint32_t)st * 8u);
            }
        }
    } else if (warp == UM_MMA_WARP) {
        // ===================== MMA issuer: one thread of the LEADER =====================
        if (lane == 0) {
            mbar_wait(wbar, 0);                                            // own weight half
            if (rank != 0) {
                mbar_arrive_cluster(mapa_shared(wpeer, 0));                // tell the leader
            } else {
                mbar_wait_cluster(wpeer, 0);
                const uint32_t blk = (uint32_t)NH * 128;
                const uint32_t idesc = make_idesc_tf32(2 * UM_ROWS, s.N_TILE);
                const int ksteps = (s.K + 7) / 8;
                const uint32_t a0 = smem_u32(Asm), b0 = smem_u32(Wsm);
                int b = 0;
                for (int it = 0; it < my_tiles; ++it) {
                    const int acc = it & 1;
                    const uint32_t aph = (it >> 1) & 1;
                    mbar_wait_cluster(&tempty[acc], aph ^ 1);
                    tc_fence_after_sync();
                    const uint32_t d_tmem = tmem_base + acc * UM_ACC_STRIDE;
                    for (int kc = 0; kc < s.KC; ++kc, ++b) {
                        const int st = b % s.stages;
                        const uint32_t ph = (uint32_t)(b / s.stages) & 1u;
                        mbar_wait_cluster(&full[st], ph);
                        tc_fence_after_sync();
                        const uint32_t ab = a0 + (uint32_t)st * (UM_BLOCK_FLOATS * 4);
                        const uint32_t bb = b0 + (uint32_t)kc * blk;
                        const int nks = min(4, ksteps - kc * 4);
                        for (int ks = 0; ks < nks; ++ks)
                            mma_tf32_ss2(d_tmem, make_desc_sw128(ab + ks * 32, 1024, 16),
                                         make_desc_sw128(bb + ks * 32, 1024, 16), idesc, (kc | ks) != 0 ? 1u : 0u);
                        mma_commit2(&empty[st]);
                    }
                    mma_commit2(&tfull[acc]);
                }
            }
        }
        __syncwarp();
    } else if (warp == UM_TMA_WARP) {
        // ===================== TMA loader (one thread per CTA, its own rows) =====================
        if constexpr (tma_in_place<AOp>()) {
            if (lane == 0) {
                tma_prefetch_desc(&aop.tmap);
                tma_prefetch_desc(&aop.tmap_dy);
                int b = 0;
                for (int it = 0; it < my_tiles; ++it) {
                    const int row0 = tile_row0(it);
                    for (int kc = 0; kc < s.KC; ++kc, ++b) {
                        const int st = b % s.stages;
                        const uint32_t ph = (uint32_t)(b / s.stages) & 1u;
                        mbar_wait(&empty[st], ph ^ 1);
                        mbar_arrive_expect_tx(&rawfull[st], UM_RAW_BYTES + UM_BLOCK_FLOATS * 4);
                        tma_load_2d(Asm + (size_t)st * UM_BLOCK_FLOATS, &aop.tmap_dy, kc * UM_KB, row0, &rawfull[st]);
                        tma_load_2d(rawring + (size_t)st * UM_RAW_BYTES, &aop.tmap, kc * UM_KB, row0, &rawfull[st]);
                    }
                }
            }
        } else if constexpr (AOp::kTma) {
            if (lane == 0) {
                tma_prefetch_desc(&aop.tmap);
                int b = 0;
                for (int it = 0; it < my_tiles; ++it) {
                    const int row0 = tile_row0(it);
                    for (int kc = 0; kc < s.KC; ++kc, ++b) {
                        const int rs = b % s.raw_stages;
                        const uint32_t rph = (uint32_t)(b / s.raw_stages) & 1u;
                        mbar_wait(&rawempty[rs], rph ^ 1);
                        mbar_arrive_expect_tx(&rawfull[rs], UM_RAW_BYTES);
                        tma_load_2d(rawring + (size_t)rs * UM_RAW_BYTES, &aop.tmap, kc * UM_KB, row0, &rawfull[rs]);
                    }
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue (as in umma_rows_kernel; tempty is the leader's) =====================
        reg_inc<192>();
        const int ew = warp & 3;
        float* slab = stage + ew * 32 * UM_STAGE_LD;
        const int rsub = lane >> 3, c4 = (lane & 7) * 4;
        const int nchunks = (s.N_TILE + 31) / 32;
        const uint32_t tempty0 = mapa_shared(tempty, 0);
        float s0[UM_STAT_CHUNKS][4], s1[UM_STAT_CHUNKS][4];
#pragma unroll
        for (int ch = 0; ch < UM_STAT_CHUNKS; ++ch)
#pragma unroll
            for (int e = 0; e < 4; ++e) { s0[ch][e] = 0.f; s1[ch][e] = 0.f; }
        float wq[8], wq_next[8];
        auto load_weights = [&](int it, float (&dst)[8]) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int r = tile_row0(min(it, max(my_tiles - 1, 0))) + ew * 32 + q * 4 + rsub;
                dst[q] = Epi::kRowWeight ? epi.row_weight(it < my_tiles ? min(r, s.R) : s.R) : 0.f;
            }
        };
        load_weights(0, wq_next);
        for (int it = 0; it < my_tiles; ++it) {
            const int acc = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            const int row0 = tile_row0(it) + ew * 32;
#pragma unroll
            for (int q = 0; q < 8; ++q) wq[q] = wq_next[q];
            if (Epi::kRowWeight) load_weights(it + 1, wq_next);
            unsigned roff[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                roff[q] = (unsigned)(min(row0 + q * 4 + rsub, s.R - 1) * epi.row_stride() + s.n0 + c4);
                asm volatile("" : "+r"(roff[q]));
            }
            mbar_wait(&tfull[acc], aph);
            tc_fence_after_sync();
            const uint32_t tbase = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * UM_ACC_STRIDE;
            uint32_t v[32];
            uint2 pre[Epi::kPrefetch ? 2 : 1][8];
            auto chunk_live = [&](int ch) { return ch * 32 + c4 < s.N_TILE && s.n0 + ch * 32 + c4 < s.N; };
            auto load_pre = [&](int ch, uint2 (&dst)[8]) {
                if (Epi::kPrefetch && chunk_live(ch)) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) dst[q] = epi.prefetch(roff[q], ch * 32);
                }
            };
            tmem_ld_32x32(tbase, v);
            load_pre(0, pre[0]);
#pragma unroll
            for (int ch = 0; ch < UM_MAX_CHUNKS; ++ch) {
                if (ch < nchunks) {
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        *reinterpret_cast<float4*>(slab + lane * UM_STAGE_LD + q * 4) =
                            make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                        __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
                    __syncwarp();
                    if (ch + 1 < nchunks) {
                        tmem_ld_32x32(tbase + (ch + 1) * 32, v);
                        load_pre(ch + 1, pre[Epi::kPrefetch ? ((ch + 1) & 1) : 0]);
                    }
                    const int cl = ch * 32 + c4;
                    const int col = s.n0 + cl;
                    if (cl < s.N_TILE && col < s.N) {
                        const int nvalid = min(4, s.N - col);
                        const typename Epi::Consts ec = epi.consts(col, aux_e);
                        float4 a[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            a[q] = *reinterpret_cast<const float4*>(slab + (q * 4 + rsub) * UM_STAGE_LD + c4);
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const int r = row0 + q * 4 + rsub;
                            epi.apply(roff[q], ch * 32, r < s.R, wq[q], col, a[q], pre[Epi::kPrefetch ? (ch & 1) : 0][q],
                                      nvalid, s0[ch < UM_STAT_CHUNKS ? ch : 0], s1[ch < UM_STAT_CHUNKS ? ch : 0], ec);
                        }
                    }
                    __syncwarp();
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty0 + (uint32_t)acc * 8u);
        }
        if (Epi::kStats) {
            float* part0 = stage;
            float* part1 = stage + 4 * 256;
            named_bar_sync(1, 128);
#pragma unroll
            for (int ch = 0; ch < UM_STAT_CHUNKS; ++ch) {
                if (ch < nchunks) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float a0 = s0[ch][e], a1 = s1[ch][e];
                        a0 += __shfl_xor_sync(0xffffffffu, a0, 8);
                        a1 += __shfl_xor_sync(0xffffffffu, a1, 8);
                        a0 += __shfl_xor_sync(0xffffffffu, a0, 16);
                        a1 += __shfl_xor_sync(0xffffffffu, a1, 16);
                        const int cl = ch * 32 + c4 + e;
                        if (lane < 8 && cl < 256) {
                            part0[ew * 256 + cl] = a0;
                            part1[ew * 256 + cl] = a1;
                        }
                    }
                }
            }
            named_bar_sync(1, 128);
            for (int cl = tid - UM_EPI_WARP0 * 32; cl < s.N_TILE; cl += 128)
                if (s.n0 + cl < s.N)
                    epi.commit(s.n0 + cl, part0[cl] + part0[256 + cl] + part0[512 + cl] + part0[768 + cl],
                               part1[cl] + part1[256 + cl] + part1[512 + cl] + part1[768 + cl], aux_e);
        }
    }

    // neither CTA may leave (or free its TMEM) while the other can still reach into it
    tc_fence_before_sync();
    cluster_sync_all();
    if (warp == UM_MMA_WARP) {
        tc_fence_after_sync();
        tmem_dealloc2(tmem_base, UM_TMEM_COLS);
    }
}

// ------------------------------------------------------------------ wgrad kernel
// dW[co, ci] = sum_r P(r, co) * Q(r, ci) over the CTA's slice of pair rows, accumulated in TMEM
// for the whole kernel and added into global dW once at the end.
//
// The reduction index is the ROW, so both operands are MN-major: P^T as A (M = co) and Q^T as B
// (N = ci).  For 32-bit (tf32) MN-major operands the hardware accepts one swizzled layout only,
// SWIZZLE_128B_BASE32B: atoms of 4 k-rows x 128 bytes (32 consecutive channels), the four
// 32-byte chunks of row k stored at chunk index (c ^ (k % 4)).  Producers store [32 rows x 32
// channels] blocks in that form; LBO = byte stride between 32-channel blocks, SBO = stride
// between 4-row atoms (512, dense).  One K chunk = 32 rows = four tcgen05.mma K steps of 8 rows.
//
// M = co is covered by two M=128 instructions: rows [0,128) and, when Cout > 128, rows
// [Cout-128, Cout) (overlap recomputed -- keeps the plain 128-lane TMEM layout).  For Cout < 128
// the operand read runs past the last channel block into the neighbouring buffer: those
// accumulator rows are garbage and never leave TMEM.
constexpr int WG_ROWS = 32;                       // rows per K chunk
constexpr int WG_BLOCK_FLOATS = WG_ROWS * UM_KB;  // [32 rows x 32 ch] = 1024 floats = 4 KB
constexpr int WG_PROD_THREADS = 256;
constexpr int WG_THREADS = WG_PROD_THREADS + 64;   // 8 producer warps (0-3 also drain TMEM at the end) + MMA + TMA warp
constexpr int WG_MAX_PB = 6;                      // Cout <= 192
constexpr int WG_MAX_QB = 8;                      // Cin  <= 256

struct WgradShape {
    int R;
    int Cout, Cin;     // valid channels of P and Q
    int PB, QB;        // 32-channel blocks of P and Q
    int N_TILE;        // MMA N = roundup16(Cin)
    int stages;
    int chunks_per_cta;
    int reverse;
    int raw_stages;    // raw slab ring depth (TMA operands), 0 = register path
    int raw_bytes;     // bytes of one raw stage: [dy fp32 | H fp16 | Hq fp16] slabs of 32 rows
    int copies;        // > 1: every CTA stores its partial dW into its own copy, `copy_stride` floats apart
    int copy_stride;   //      (dW 16-byte aligned, ldw % 4 == 0); 1: atomic accumulation into dW
    unsigned zero;     // always 0; unknown to the compiler (raw-ring release of the producers)
};

static inline size_t wgrad_smem_bytes(const WgradShape& s) {
    return 1024 + (size_t)s.stages * (s.PB + s.QB) * WG_BLOCK_FLOATS * 4 + (size_t)UM_ROWS * UM_STAGE_LD * 4 +
           (3 + 3) * kMaxC * 4 + (size_t)s.raw_stages * s.raw_bytes + 32 * 8 + 16;
}

// PBc / QBc: compile-time operand block counts (0 = take them from the shape at run time; the run-time
// guards cost more instructions per chunk than the arithmetic of the small layers)
template <class POp, class QOp, int PBc, int QBc>
__global__ void __launch_bounds__(WG_THREADS, 1)   // 10 warps = 3 on one SM sub-partition: 16384 / 96 -> at most 168 registers
umma_wgrad_kernel(const __grid_constant__ POp pop, const __grid_constant__ QOp qop, float* __restrict__ dW, int ldw,
                  const __grid_constant__ WgradShape s) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by OFFSETTING the shared array (an integer round trip would turn every
    // shared-memory pointer below into a generic one: LD/ST instead of LDS/STS throughout the kernel)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int nPB = PBc ? PBc : s.PB, nQB = QBc ? QBc : s.QB;
    constexpr int kPBmax = PBc ? PBc : WG_MAX_PB, kQBmax = QBc ? QBc : WG_MAX_QB;
    const int stage_floats = (nPB + nQB) * WG_BLOCK_FLOATS;
    float* ring = reinterpret_cast<float*>(smem);
    float* stage = ring + (size_t)s.stages * stage_floats;
    float* aux_p = stage + UM_ROWS * UM_STAGE_LD;
    float* aux_q = aux_p + 3 * kMaxC;
    uint8_t* rawring = reinterpret_cast<uint8_t*>(aux_q + 3 * kMaxC);   // [raw_stages][raw_bytes]
    uint64_t* bars = reinterpret_cast<uint64_t*>(rawring + (size_t)s.raw_stages * s.raw_bytes);
    uint64_t* full = bars;        // [UM_MAX_STAGES]
    uint64_t* empty = bars + 4;   // [UM_MAX_STAGES]
    uint64_t* done = bars + 8;
    uint64_t* rawfull = bars + 9;     // [4]
    uint64_t* rawempty = bars + 13;   // [4]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int nchunks_total = (s.R + WG_ROWS - 1) / WG_ROWS;
    // CTA b owns a contiguous chunk range; with `reverse` the ranges are handed out from the end so
    // that the kernel starts on the rows the previous launch touched last (L2 residency)
    const int owner = s.reverse ? (int)gridDim.x - 1 - (int)blockIdx.x : (int)blockIdx.x;
    const int c_begin = owner * s.chunks_per_cta;
    const int c_end = min(nchunks_total, c_begin + s.chunks_per_cta);
    const int my_chunks = max(0, c_end - c_begin);

    if (tid == 0) {
        for (int i = 0; i < UM_MAX_STAGES; ++i) {
            mbar_init(&full[i], WG_PROD_THREADS / 32);   // one elected arrival per producer warp
            mbar_init(&empty[i], 1);
        }
        mbar_init(done, 1);
        for (int i = 0; i < 4; ++i) {
            mbar_init(&rawfull[i], 1);
            mbar_init(&rawempty[i], WG_PROD_THREADS / 32);
        }
        fence_mbar_init();
    }
    if (warp == 8) tmem_alloc(tmem_slot, UM_TMEM_COLS);
    pdl_wait();                 // see umma_rows_kernel: no-ops unless launched with programmatic serialization
    pdl_launch_dependents();
    pop.init(aux_p, tid, WG_THREADS);
    qop.init(aux_q, tid, WG_THREADS);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const bool two_blocks = s.Cout > 128;

    if (warp < 8) {
        // ===================== producers: thread t -> row t/8 of the chunk, 16-byte column t%8
        const int rl = tid >> 3, c16 = tid & 7;
        // floats: 4-row atom (rl/4) * 128 + row-in-atom * 32 + swizzled 32-byte chunk * 8 + 16-byte half * 4
        const int off = (rl >> 2) * 128 + (rl & 3) * 32 + ((((c16 >> 1) ^ (rl & 3))) << 3) + ((c16 & 1) << 2);
        int st = 0;
        uint32_t ph = 0;
        if constexpr (POp::kTma) {
            // raw stage: [dy: 32 x Cout fp32][H_k: 32 x Cout fp16][H_{k-1}: 32 x Cin fp16], row-major
            int rs = 0;
            uint32_t rph = 0;
            const uint32_t off_h = WG_ROWS * s.Cout * 4, off_q = off_h + WG_ROWS * s.Cout * 2;
            typename QOp::Row qr;
            typename QOp::Raw qraw[QOp::kTma ? 1 : kQBmax];
            // This thread's channel constants of P stay in registers for the whole launch where the budget
            // (168) allows: next to a TMA-fed Q.  With the |x_i - x_j| operand (48 registers of loads in
            // flight) they are reloaded per chunk, three blocks per batch of loads.
            constexpr bool kPersist = QOp::kTma;
            constexpr int kBatch = kPersist ? kPBmax : (kPBmax + 1) / 2;
            typename POp::K pk[kPersist ? kPBmax : kBatch];
            if constexpr (kPersist) {
#pragma unroll
                for (int b = 0; b < kPBmax; ++b)
                    if (PBc != 0 || b < s.PB) pk[b] = pop.kload(b * UM_KB + c16 * 4, aux_p);
            }
            // the row multiplicity: a pure load from the per-row table (entry R = 0), fetched one chunk ahead
            float w_next = pop.row(c_begin < c_end ? min(c_begin * WG_ROWS + rl, s.R) : s.R).w;
            for (int c = c_begin; c < c_end; ++c) {
                const int r = c * WG_ROWS + rl;
                const bool ok = r < s.R;
                const float w = w_next;
                {
                    const int rn = r + WG_ROWS;
                    w_next = pop.row(c + 1 < c_end ? min(rn, s.R) : s.R).w;
                }
                mbar_wait(&rawfull[rs], rph);
                const uint8_t* rb = rawring + (size_t)rs * s.raw_bytes;
                float4 dv[kPBmax];
                uint2 hv[kPBmax];
                uint2 qv[kQBmax];
#pragma unroll
                for (int b = 0; b < kPBmax; ++b) {
                    if (PBc != 0 || b < s.PB) {
                        const int k = min(b * UM_KB + c16 * 4, s.Cout - 4);
                        dv[b] = *reinterpret_cast<const float4*>(rb + ((size_t)rl * s.Cout + k) * 4);
                        hv[b] = *reinterpret_cast<const uint2*>(rb + off_h + ((size_t)rl * s.Cout + k) * 2);
                    }
                }
                if constexpr (QOp::kTma) {
#pragma unroll
                    for (int b = 0; b < kQBmax; ++b)
                        if (QBc != 0 || b < s.QB) {
                            const int k = min(b * UM_KB + c16 * 4, s.Cin - 4);
                            qv[b] = *reinterpret_cast<const uint2*>(rb + off_q + ((size_t)rl * s.Cin + k) * 2);
                        }
                }
                // Release the slab only once its loads have COMPLETED in every lane -- they have merely been
                // ISSUED here, and the TMA thread refills a released slot at once, under LDS that still sit in
                // the memory pipeline behind the other warps' stores and the MMA's operand reads (seen on B200
                // as a 10 % run-to-run spread of d conv2d_1.weight once this role stopped stalling on the pair
                // table).  The arrival count is tied to every lane's loaded registers through a warp-wide OR
                // with a run-time zero, so neither compiler nor hardware can let the arrive overtake the loads.
                {
                    unsigned t = 0;
#pragma unroll
                    for (int b = 0; b < kPBmax; ++b)
                        if (PBc != 0 || b < s.PB)
                            t |= __float_as_uint(dv[b].x) | __float_as_uint(dv[b].y) | __float_as_uint(dv[b].z) |
                                 __float_as_uint(dv[b].w) | hv[b].x | hv[b].y;
                    if constexpr (QOp::kTma) {
#pragma unroll
                        for (int b = 0; b < kQBmax; ++b)
                            if (QBc != 0 || b < s.QB) t |= qv[b].x | qv[b].y;
                    }
                    const unsigned z = __reduce_or_sync(0xffffffffu, t & s.zero);
                    if (lane == 0) mbar_arrive_n(&rawempty[rs], 1u + z);
                }
                if (++rs == s.raw_stages) { rs = 0; rph ^= 1; }
                if constexpr (!QOp::kTma) {                       // |x_i - x_j| from the L2-resident node matrix
                    qr = qop.row(ok ? r : 0);
#pragma unroll
                    for (int b = 0; b < kQBmax; ++b)
                        if (QBc != 0 || b < s.QB) qop.fetch(qr, b * UM_KB + c16 * 4, qraw[b]);
                }
                mbar_wait(&empty[st], ph ^ 1);
                float* dst = ring + (size_t)st * stage_floats + off;
                const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int b0 = 0; b0 < kPBmax; b0 += kBatch) {
                    if constexpr (!kPersist) {
#pragma unroll
                        for (int b = b0; b < b0 + kBatch && b < kPBmax; ++b)
                            if (PBc != 0 || b < s.PB) pk[b - b0] = pop.kload(b * UM_KB + c16 * 4, aux_p);
                    }
#pragma unroll
                    for (int b = b0; b < b0 + kBatch && b < kPBmax; ++b) {
                        if (PBc != 0 || b < s.PB) {
                            float4 v = ok ? pop.transform_k(dv[b], hv[b], w, b * UM_KB + c16 * 4, pk[kPersist ? b : b - b0]) : zero;
                            v.x = to_tf32_fast(v.x); v.y = to_tf32_fast(v.y); v.z = to_tf32_fast(v.z); v.w = to_tf32_fast(v.w);
                            *reinterpret_cast<float4*>(dst + b * WG_BLOCK_FLOATS) = v;
                        }
                    }
                }
                if constexpr (QOp::kTma) {                        // all of Q's constants in one batch of loads
                    typename QOp::K qk[kQBmax];
#pragma unroll
                    for (int b = 0; b < kQBmax; ++b)
                        if (QBc != 0 || b < s.QB) qk[b] = qop.kload(b * UM_KB + c16 * 4, aux_q);
#pragma unroll
                    for (int b = 0; b < kQBmax; ++b) {
                        if (QBc != 0 || b < s.QB) {
                            float4 v = ok ? qop.transform_k(qv[b], b * UM_KB + c16 * 4, qk[b]) : zero;
                            v.x = to_tf32_fast(v.x); v.y = to_tf32_fast(v.y); v.z = to_tf32_fast(v.z); v.w = to_tf32_fast(v.w);
                            *reinterpret_cast<float4*>(dst + (nPB + b) * WG_BLOCK_FLOATS) = v;
                        }
                    }
                } else {
#pragma unroll
                for (int b = 0; b < kQBmax; ++b) {
                    if (QBc != 0 || b < s.QB) {
                        float4 v;
                        v = ok ? qop.finish(qraw[QOp::kTma ? 0 : b], qr, b * UM_KB + c16 * 4, aux_q) : zero;
                        v.x = to_tf32_fast(v.x); v.y = to_tf32_fast(v.y); v.z = to_tf32_fast(v.z); v.w = to_tf32_fast(v.w);
                        *reinterpret_cast<float4*>(dst + (nPB + b) * WG_BLOCK_FLOATS) = v;
                    }
                }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[st]);
                if (++st == s.stages) { st = 0; ph ^= 1; }
            }
        } else {
        typename POp::Raw praw[kPBmax];
        typename QOp::Raw qraw[kQBmax];
        bool ok = false;
        typename POp::Row pr;
        typename QOp::Row qr;
        // One chunk of loads in flight per thread (issued right after the previous chunk was handed
        // to the MMA).  Deeper register prefetch was tried (two static buffers) and bought nothing:
        // with six scoreboard slots the extra loads alias the slots of the ones being consumed.
        // Q operands with kAhead == 0 (|x_i - x_j| from the small L2-resident node matrix) are
        // fetched and consumed inside the step, after the P blocks have been stored, so that their
        // registers do not add to P's in-flight set.
        constexpr bool kLateQ = (QOp::kAhead == 0);
        auto fetch_p = [&](int chunk) {
            const int r = chunk * WG_ROWS + rl;
            ok = r < s.R;
            pr = pop.row(ok ? r : 0);
#pragma unroll
            for (int b = 0; b < kPBmax; ++b)
                if (PBc != 0 || b < s.PB) pop.fetch(pr, b * UM_KB + c16 * 4, praw[b]);
        };
        auto fetch_q = [&](int chunk) {
            const int r = chunk * WG_ROWS + rl;
            qr = qop.row(r < s.R ? r : 0);
#pragma unroll
            for (int b = 0; b < kQBmax; ++b)
                if (QBc != 0 || b < s.QB) qop.fetch(qr, b * UM_KB + c16 * 4, qraw[b]);
        };
        if (my_chunks > 0) {
            fetch_p(c_begin);
            if (!kLateQ) fetch_q(c_begin);
        }
        for (int c = c_begin; c < c_end; ++c) {
            mbar_wait(&empty[st], ph ^ 1);
            float* dst = ring + (size_t)st * stage_floats + off;
            const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int b = 0; b < kPBmax; ++b) {
                if (PBc != 0 || b < s.PB) {
                    float4 v = ok ? pop.finish(praw[b], pr, b * UM_KB + c16 * 4, aux_p) : zero;
                    v.x = to_tf32_fast(v.x); v.y = to_tf32_fast(v.y); v.z = to_tf32_fast(v.z); v.w = to_tf32_fast(v.w);
                    *reinterpret_cast<float4*>(dst + b * WG_BLOCK_FLOATS) = v;
                }
            }
            if (kLateQ) fetch_q(c);
#pragma unroll
            for (int b = 0; b < kQBmax; ++b) {
                if (QBc != 0 || b < s.QB) {
                    float4 v = ok ? qop.finish(qraw[b], qr, b * UM_KB + c16 * 4, aux_q) : zero;
                    v.x = to_tf32_fast(v.x); v.y = to_tf32_fast(v.y); v.z = to_tf32_fast(v.z); v.w = to_tf32_fast(v.w);
                    *reinterpret_cast<float4*>(dst + (nPB + b) * WG_BLOCK_FLOATS) = v;
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[st]);
            if (++st == s.stages) { st = 0; ph ^= 1; }
            if (c + 1 < c_end) {
                fetch_p(c + 1);
                if (!kLateQ) fetch_q(c + 1);
            }
        }
        }   // register path
    } else if (warp == 9) {
        // ===================== TMA loader (one thread): slabs of 32 rows, all operands of a chunk
        if constexpr (POp::kTma) {
            if (lane == 0 && my_chunks > 0) {
                tma_prefetch_desc(&pop.tmap_dy);
                tma_prefetch_desc(&pop.tmap_h);
                const uint32_t off_h = WG_ROWS * s.Cout * 4, off_q = off_h + WG_ROWS * s.Cout * 2;
                uint32_t bytes = WG_ROWS * s.Cout * 6;
                if constexpr (QOp::kTma) bytes += WG_ROWS * s.Cin * 2;
                int rs = 0;
                uint32_t rph = 0;
                for (int c = c_begin; c < c_end; ++c) {
                    mbar_wait(&rawempty[rs], rph ^ 1);
                    uint8_t* rb = rawring + (size_t)rs * s.raw_bytes;
                    mbar_arrive_expect_tx(&rawfull[rs], bytes);
                    tma_load_2d(rb, &pop.tmap_dy, 0, c * WG_ROWS, &rawfull[rs]);
                    tma_load_2d(rb + off_h, &pop.tmap_h, 0, c * WG_ROWS, &rawfull[rs]);
                    if constexpr (QOp::kTma) tma_load_2d(rb + off_q, &qop.tmap_h, 0, c * WG_ROWS, &rawfull[rs]);
                    if (++rs == s.raw_stages) { rs = 0; rph ^= 1; }
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== MMA issuer
        if (lane == 0 && my_chunks > 0) {
            const uint32_t idesc = make_idesc_tf32(128, s.N_TILE, true, true);
            const uint32_t ring0 = smem_u32(ring);
            const uint32_t lbo = WG_BLOCK_FLOATS * 4;       // next 32-channel block
            const uint32_t sbo = 512;                       // next 4-row atom
            const uint32_t kstep = 1024;                    // 8 rows per MMA
            const uint32_t m2_off = two_blocks ? (uint32_t)((s.Cout - 128) / UM_KB) * lbo : 0;
            int st = 0;
            uint32_t ph = 0;
            for (int c = c_begin; c < c_end; ++c) {
                mbar_wait(&full[st], ph);
                tc_fence_after_sync();
                const uint32_t pb = ring0 + (uint32_t)st * stage_floats * 4;
                const uint32_t qb = pb + (uint32_t)nPB * lbo;
                for (int ks = 0; ks < WG_ROWS / 8; ++ks) {
                    const uint32_t acc_on = (c > c_begin || ks > 0) ? 1u : 0u;
                    const uint64_t bdesc = make_desc_sw128(qb + ks * kstep, sbo, lbo, 1);
                    mma_tf32_ss(tmem_base, make_desc_sw128(pb + ks * kstep, sbo, lbo, 1), bdesc, idesc, acc_on);
                    if (two_blocks)
                        mma_tf32_ss(tmem_base + UM_ACC_STRIDE,
                                    make_desc_sw128(pb + m2_off + ks * kstep, sbo, lbo, 1), bdesc, idesc, acc_on);
                }
                mma_commit(&empty[st]);
                if (++st == s.stages) { st = 0; ph ^= 1; }
            }
            mma_commit(done);
        }
        __syncwarp();
    }

    // ===================== dump: warps 0-3 drain TMEM and add into global dW
    if (warp < 4 && my_chunks > 0) {
        mbar_wait(done, 0);
        tc_fence_after_sync();
        const int et = tid;                 // 0..127 = TMEM lane
        const int rsub = et >> 3, c4 = (et & 7) * 4;
        const int nchunks = (s.N_TILE + 31) / 32;
        const int nblk = two_blocks ? 2 : 1;
        for (int blk = 0; blk < nblk; ++blk) {
            // block 0: lane l <-> co l; block 1: lane l <-> co (Cout-128)+l, only the rows block 0 lacks
            const int co_base = blk == 0 ? 0 : s.Cout - 128;
            const int lane_lo = blk == 0 ? 0 : 128 - (s.Cout - 128);
            for (int ch = 0; ch < nchunks; ++ch) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + blk * UM_ACC_STRIDE + ch * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    *reinterpret_cast<float4*>(stage + et * UM_STAGE_LD + q * 4) =
                        make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                    __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
                named_bar_sync(1, 128);
                const int col = ch * 32 + c4;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int l = q * 16 + rsub;
                    const int co = co_base + l;
                    if (l >= lane_lo && co < s.Cout && col < s.Cin) {
                        const float* sp = stage + l * UM_STAGE_LD + c4;
                        if (s.copies > 1) {
                            // private partial per CTA, plain 16-byte stores (rows padded to 4 floats);
                            // finalize_grads_kernel adds the copies up in a fixed order
                            float* gp = dW + (size_t)blockIdx.x * s.copy_stride + (size_t)co * ldw + col;
                            *reinterpret_cast<float4*>(gp) = *reinterpret_cast<const float4*>(sp);
                        } else {
                            float* gp = dW + (size_t)co * ldw + col;
                            const int nv = min(4, s.Cin - col);
                            for (int e = 0; e < nv; ++e) atomicAdd(gp + e, sp[e]);
                        }
                    }
                }
                named_bar_sync(1, 128);
            }
        }
        tc_fence_before_sync();
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, UM_TMEM_COLS);
    }
}

// ------------------------------------------------------------------ host side
// ---- tensor maps ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// Row-major fp32 matrix [rows, cols] with leading dimension ld (elements): boxes of box_cols x box_rows.
static int make_tmap_2d(CUtensorMap* out, const void* base, CUtensorMapDataType dtype, int esize, int rows, int cols,
                        int ld, int box_cols, int box_rows, CUtensorMapSwizzle swz) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) {
        set_error(MFT_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
        return MFT_ERR_CUDA;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * esize};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, dtype, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error(MFT_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d ld=%d", (int)r, rows, cols, ld);
        return MFT_ERR_CUDA;
    }
    return MFT_OK;
}

static inline int plain_vec_ok(const float* p, int ld, int K) {
    return (ld % 4 == 0 && K % 4 == 0 && K >= 4 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) ? 1 : 0;
}
static inline int absdiff_vec_ok(const float* x, int ldx, int F) {
    return (ldx % 4 == 0 && ldx >= ((F + 3) & ~3) && F >= 4 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) ? 1 : 0;
}

long long* g_umma_dbg = nullptr;   // set by mft_debug_set_timeline()
int g_umma_dbg_skip = 0;           // rows-kernel launches to let pass before recording one

// Serpentine schedule: every pair-row kernel of the tensor-core path flips the direction in which
// it walks the rows, so the tail of what one launch wrote is the head of what the next one reads.
static int g_direction = 0;
static int next_direction() { return (g_direction ^= 1); }
static int g_num_sms = 0;
static int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

static int g_grid_limit = 0;
void umma_set_grid_limit(int ctas) { g_grid_limit = ctas; }
int umma_num_sms() { return num_sms(); }
static int grid_cap() { return g_grid_limit > 0 ? min(g_grid_limit, num_sms()) : num_sms(); }

constexpr size_t kSmemLimit = 232448;   // 227 KB opt-in maximum per CTA on sm_100

// Which rows kernel runs (MFT_ROWS_KERNEL): "base" (default) = umma_rows_kernel, "ts" = A operand in tensor memory
// (umma_rows_ts_kernel), "pair" = CTA pairs (umma_rows_pair_kernel).  All three pass the same parity tests; on
// B200 the base kernel is the fastest by 3-5 % (profiles/r02_summary.md: the three share their bottleneck, the
// issue rate of the CUDA-core work around the MMA, ~23k warp instructions per 128-row tile).  Read once: the pass split of a layer
// (and with it the layout of its weight image) depends on it.  A pass the chosen kernel cannot take (ts: more
// than 224 output columns) is never planned: plan_pass() refuses it and the layer is split further.
enum { ROWS_BASE = 0, ROWS_PAIR = 1, ROWS_TS = 2 };
static int g_rows_mode = -1;
static int rows_mode() {
    if (g_rows_mode < 0) {
        const char* e = getenv("MFT_ROWS_KERNEL");
        g_rows_mode = ROWS_BASE;
        if (e && !strcmp(e, "pair")) g_rows_mode = ROWS_PAIR;
        if (e && !strcmp(e, "ts")) g_rows_mode = ROWS_TS;
    }
    return g_rows_mode;
}
static bool pair_mode() { return rows_mode() == ROWS_PAIR; }
static inline size_t rows_smem_bytes(const UmmaShape& s) { return pair_mode() ? umma_smem_bytes_pair(s) : umma_smem_bytes(s); }

static bool plan_pass(int N_TILE, int K, UmmaShape& s) {
    s.N_TILE = N_TILE;
    s.KC = (K + UM_KB - 1) / UM_KB;
    if (s.KC > UM_MAX_KC || N_TILE > UM_MAX_NTILE || (N_TILE % 16) != 0) return false;
    s.raw_stages = 0;
    if (rows_mode() == ROWS_TS) {
        // resident weights + two raw slots of the widest operand (dy fp32 + H fp16 blocks of the dgrad)
        if (N_TILE > UM_TS_MAX_NTILE) return false;
        s.stages = ts_slots(N_TILE);
        s.raw_stages = 2;
        return umma_smem_bytes_ts<DhInPlaceT>(s) <= kSmemLimit;
    }
    for (int st = UM_MAX_STAGES; st >= 2; --st) {
        s.stages = st;
        // (the in-place dgrad operand needs a raw slot per stage: plan for at least two such pairs)
        UmmaShape t = s;
        if (pair_mode()) { t.stages = 2; t.raw_stages = 2; if (rows_smem_bytes(t) > kSmemLimit) return false; }
        if (rows_smem_bytes(s) <= kSmemLimit) return true;
    }
    return false;
}

// Split N output columns into passes whose resident weight image fits shared memory.
static int plan_passes(int N, int K, int* n0s, int* ntiles) {
    int padded = (N + 15) & ~15;
    for (int passes = 1; passes <= 4; ++passes) {
        int per = (((padded + passes - 1) / passes) + 15) & ~15;
        UmmaShape s{};
        if (per <= UM_MAX_NTILE && plan_pass(per, K, s)) {
            for (int p = 0; p < passes; ++p) {
                n0s[p] = p * per;
                ntiles[p] = per;
            }
            return passes;
        }
    }
    return 0;
}

size_t umma_wimg_floats(int N, int K) {
    int n0s[4], nt[4];
    int passes = plan_passes(N, K, n0s, nt);
    if (passes == 0) return 0;
    int KC = (K + UM_KB - 1) / UM_KB;
    return (size_t)passes * nt[0] * KC * UM_KB;
}

template <class AOp, class Epi>
static int umma_rows_gemm(const AOp& aop, const Epi& epi, const float* W, int ldw, int transpose, int R, int N,
                          int K, float* wimg, cudaStream_t st, int cat, bool prebuilt = false, bool pdl = false) {
    // pdl: the launch directly before this one on `st` is a kernel of this library that follows the
    // wait-then-release convention (common.cuh) AND started after the weight images were complete.
    int n0s[4], nts[4];
    int passes = plan_passes(N, K, n0s, nts);
    if (passes == 0) {
        set_error(MFT_ERR_UNSUPPORTED, "umma_rows_gemm: no shared-memory plan for N=%d K=%d", N, K);
        return MFT_ERR_UNSUPPORTED;
    }
    if ((size_t)R * (size_t)max(max(N, K), 256) >= ((size_t)1 << 31)) {   // the kernels index rows with 32-bit element offsets
        set_error(MFT_ERR_UNSUPPORTED, "umma_rows_gemm: R=%d too large for 32-bit row offsets", R);
        return MFT_ERR_UNSUPPORTED;
    }
    const int ntiles = cdiv(R, UM_ROWS);
    const int grid = min(ntiles, grid_cap());
    for (int p = 0; p < passes; ++p) {
        const int dir = next_direction();
        UmmaShape s{};
        plan_pass(nts[p], K, s);
        if (rows_mode() == ROWS_TS) {
            s.R = R; s.N = N; s.n0 = n0s[p]; s.K = K; s.reverse = dir; s.zero = 0u; s.dbg = nullptr;
            if (g_umma_dbg && g_umma_dbg_skip-- == 0) { s.dbg = g_umma_dbg; g_umma_dbg = nullptr; }
            s.stages = ts_slots(s.N_TILE);
            s.raw_stages = 0;
            // eight epilogue warps where their second set of staging slabs still leaves room for the raw ring
            // (the in-place dgrad operand: two 24 KB slots; the fp16 tape: four 8 KB slots) and the producers'
            // register budget allows it (MFT_EPI_WARPS=4 forces four)
            static const int ew_env = [] { const char* e = getenv("MFT_EPI_WARPS"); return e ? atoi(e) : 8; }();
            int ew = (ew_env == 4 || tma_in_place<AOp>()) ? 4 : 8;
            if (AOp::kTma) {
                for (;;) {
                    s.raw_stages = 0;
                    for (int rs = UM_MAX_RAW; rs >= (ew == 8 ? 4 : 2); --rs) {
                        s.raw_stages = rs;
                        if (umma_smem_bytes_ts<AOp>(s, ew) <= kSmemLimit) break;
                        s.raw_stages = 0;
                    }
                    if (s.raw_stages != 0 || ew == 4) break;
                    ew = 4;
                }
                if (s.raw_stages == 0) {
                    set_error(MFT_ERR_UNSUPPORTED, "umma_rows_gemm: no room for the raw ring (N=%d K=%d)", N, K);
                    return MFT_ERR_UNSUPPORTED;
                }
            } else if (umma_smem_bytes_ts<AOp>(s, ew) > kSmemLimit) {
                ew = 4;
            }
            float* img = wimg + (size_t)p * s.N_TILE * s.KC * UM_KB;
            if (!prebuilt) {
                ProfScope ps(PC_PREP, st);
                int total = s.KC * s.N_TILE * UM_KB;
                umma_weight_image_kernel<<<cdiv(total, 256), 256, 0, st>>>(W, ldw, transpose, N, K, s.n0, s.N_TILE,
                                                                           s.KC, img);
                MFT_CHECK_LAUNCH();
            }
            const size_t smem = umma_smem_bytes_ts<AOp>(s, ew);
            ProfScope ps(cat, st);
            const bool use_pdl = prebuilt && pdl_level() >= 1 && (pdl || p > 0);
            if (ew == 8) {
                MFT_CHECK_CUDA(cudaFuncSetAttribute(umma_rows_ts_kernel<AOp, Epi, 8>,
                                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                MFT_CHECK_CUDA(launch_kernel(umma_rows_ts_kernel<AOp, Epi, 8>, dim3(grid), dim3((UM_PROD_WARPS + 10) * 32),
                                             smem, st, use_pdl, aop, epi, (const float*)img, s));
            } else {
                MFT_CHECK_CUDA(cudaFuncSetAttribute(umma_rows_ts_kernel<AOp, Epi, 4>,
                                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                MFT_CHECK_CUDA(launch_kernel(umma_rows_ts_kernel<AOp, Epi, 4>, dim3(grid), dim3(UM_THREADS), smem, st,
                                             use_pdl, aop, epi, (const float*)img, s));
            }
            continue;
        }
        if (tma_in_place<AOp>()) {
            // stage st and raw slot st travel together: as many as fit
            bool fit = false;
            for (int stg = pair_mode() ? UM_MAX_PSTAGES : UM_MAX_STAGES; stg >= 2 && !fit; --stg) {
                s.stages = stg;
                s.raw_stages = stg;
                fit = rows_smem_bytes(s) <= kSmemLimit;
            }
            if (!fit) {
                set_error(MFT_ERR_UNSUPPORTED, "umma_rows_gemm: no room for in-place TMA stages (N=%d K=%d)", N, K);
                return MFT_ERR_UNSUPPORTED;
            }
        } else if (AOp::kTma) {
            // two fp32 A stages are enough once the loads run ahead in the fp16 raw ring: give the
            // rest of shared memory to raw blocks in flight
            s.stages = 2;
            s.raw_stages = 0;
            if (pair_mode()) {                       // four A stages (one per producer group) when four raw slots still fit
                s.stages = 4;
                s.raw_stages = 4;
                if (rows_smem_bytes(s) > kSmemLimit) s.stages = 2;
                s.raw_stages = 0;
            }
            for (int rs = UM_MAX_RAW; rs >= 2; --rs) {
                s.raw_stages = rs;
                if (rows_smem_bytes(s) <= kSmemLimit) break;
                s.raw_stages = 0;
            }
            if (s.raw_stages == 0) {
                set_error(MFT_ERR_UNSUPPORTED, "umma_rows_gemm: no room for the raw ring (N=%d K=%d)", N, K);
                return MFT_ERR_UNSUPPORTED;
            }
        }
        s.R = R; s.N = N; s.n0 = n0s[p]; s.K = K; s.reverse = dir; s.zero = 0u;
        s.dbg = nullptr;
        if (g_umma_dbg && g_umma_dbg_skip-- == 0) { s.dbg = g_umma_dbg; g_umma_dbg = nullptr; }
        float* img = wimg + (size_t)p * s.N_TILE * s.KC * UM_KB;
        if (!prebuilt) {
            ProfScope ps(PC_PREP, st);
            int total = s.KC * s.N_TILE * UM_KB;
            umma_weight_image_kernel<<<cdiv(total, 256), 256, 0, st>>>(W, ldw, transpose, N, K, s.n0, s.N_TILE,
                                                                       s.KC, img);
            MFT_CHECK_LAUNCH();
        }
        size_t smem = rows_smem_bytes(s);
        static_assert(sizeof(AOp) + sizeof(Epi) < 3500, "kernel parameter space");
        ProfScope ps(cat, st);
        // a second pass follows the first pass of the same layer: same images, same convention
        const bool use_pdl = prebuilt && pdl_level() >= 1 && (pdl || p > 0);
        if (pair_mode()) {
            const int nst = cdiv(R, 2 * UM_ROWS);
            const int npairs = max(1, min(nst, grid_cap() / 2));
            // producer groups: as many as every ring the role uses is deep (a group may not run two phases ahead)
            int pg = 4;
            if (!AOp::kTma) pg = 2;                   // register-fed operands: 16 rows of addresses per thread would spill
            while (pg > 1 && (pg > s.stages || (AOp::kTma && !tma_in_place<AOp>() && pg > s.raw_stages))) pg >>= 1;
#define MFT_PAIR_LAUNCH(PGV)                                                                                         \
    do {                                                                                                             \
        MFT_CHECK_CUDA(cudaFuncSetAttribute(umma_rows_pair_kernel<AOp, Epi, PGV>,                                    \
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                 \
        MFT_CHECK_CUDA(launch_kernel(umma_rows_pair_kernel<AOp, Epi, PGV>, dim3(2 * npairs), dim3(UM_THREADS), smem, \
                                     st, use_pdl, aop, epi, (const float*)img, s));                                  \
    } while (0)
            if (pg == 4) MFT_PAIR_LAUNCH(4);
            else if (pg == 2) MFT_PAIR_LAUNCH(2);
            else MFT_PAIR_LAUNCH(1);
#undef MFT_PAIR_LAUNCH
        } else {
            MFT_CHECK_CUDA(cudaFuncSetAttribute(umma_rows_kernel<AOp, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)smem));
            MFT_CHECK_CUDA(launch_kernel(umma_rows_kernel<AOp, Epi>, dim3(grid), dim3(UM_THREADS), smem, st, use_pdl,
                                         aop, epi, (const float*)img, s));
        }
    }
    return MFT_OK;
}

template <class POp, class QOp, bool kSpecialize = true>
static int umma_wgrad(const POp& pop, const QOp& qop, float* dW, int ldw, int R, int Cout, int Cin,
                      cudaStream_t st, int cat, int* copies_out = nullptr, bool pdl = false) {
    WgradShape s{};
    s.R = R; s.Cout = Cout; s.Cin = Cin;
    s.copy_stride = Cout * ldw;
    if (copies_out && (ldw % 4 != 0 || (reinterpret_cast<uintptr_t>(dW) & 15) != 0)) {
        set_error(MFT_ERR_ARG, "umma_wgrad: partial copies need 16-byte aligned rows (ldw=%d)", ldw);
        return MFT_ERR_ARG;
    }
    s.PB = cdiv(Cout, UM_KB);
    s.QB = cdiv(Cin, UM_KB);
    s.N_TILE = (Cin + 15) & ~15;
    if (s.PB > WG_MAX_PB || s.QB > WG_MAX_QB || s.N_TILE > 256 || Cout > 192 || (Cout > 128 && Cout % UM_KB != 0)) {
        set_error(MFT_ERR_UNSUPPORTED, "umma_wgrad: unsupported Cout=%d Cin=%d", Cout, Cin);
        return MFT_ERR_UNSUPPORTED;
    }
    s.stages = 0;
    s.raw_stages = 0;
    s.raw_bytes = 0;
    if (POp::kTma) {
        // raw slabs of one chunk: dy fp32 + H_k fp16 (+ H_{k-1} fp16); two operand stages, and as
        // many raw stages in flight as shared memory allows (at least two)
        s.raw_bytes = WG_ROWS * Cout * 6 + (QOp::kTma ? WG_ROWS * Cin * 2 : 0);
        s.raw_bytes = (s.raw_bytes + 127) & ~127;
        s.stages = 2;
        for (int rs = 4; rs >= 2; --rs) {
            s.raw_stages = rs;
            if (wgrad_smem_bytes(s) <= kSmemLimit) break;
            s.raw_stages = 0;
        }
        if (s.raw_stages == 0) s.stages = 0;
    } else
    for (int stg = UM_MAX_STAGES; stg >= 2; --stg) {
        s.stages = stg;
        if (wgrad_smem_bytes(s) <= kSmemLimit) break;
        s.stages = 0;
    }
    if (s.stages == 0) {
        set_error(MFT_ERR_UNSUPPORTED, "umma_wgrad: no shared-memory plan for Cout=%d Cin=%d", Cout, Cin);
        return MFT_ERR_UNSUPPORTED;
    }
    const int nchunks = cdiv(R, WG_ROWS);
    const int grid = min(nchunks, min(grid_cap(), kWgMaxCopies));
    s.chunks_per_cta = cdiv(nchunks, grid);
    s.copies = copies_out ? max(2, cdiv(nchunks, s.chunks_per_cta)) : 1;   // (a one-CTA launch still stores)
    if (copies_out) *copies_out = cdiv(nchunks, s.chunks_per_cta);
    s.reverse = next_direction();
    s.zero = 0u;
    size_t smem = wgrad_smem_bytes(s);
    const int grid_x = cdiv(nchunks, s.chunks_per_cta);
#define MFT_WG_LAUNCH(PBC, QBC)                                                                                      \
    do {                                                                                                             \
        MFT_CHECK_CUDA(cudaFuncSetAttribute(umma_wgrad_kernel<POp, QOp, PBC, QBC>,                                    \
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                 \
        ProfScope ps(cat, st);                                                                                       \
        MFT_CHECK_CUDA(launch_kernel(umma_wgrad_kernel<POp, QOp, PBC, QBC>, dim3(grid_x), dim3(WG_THREADS), smem, st, \
                                     pdl, pop, qop, dW, ldw, s));                                                    \
        return MFT_OK;                                                                                               \
    } while (0)
    // the block counts of the reference's layer widths (nf = 96: 192/192/96/96, F = 133/181/229) are compiled in
    if (kSpecialize) {
        if (s.PB == 6 && s.QB == 6) MFT_WG_LAUNCH(6, 6);
        if (s.PB == 3 && s.QB == 6) MFT_WG_LAUNCH(3, 6);
        if (s.PB == 3 && s.QB == 3) MFT_WG_LAUNCH(3, 3);
        if (s.PB == 6 && s.QB == 5) MFT_WG_LAUNCH(6, 5);
        if (s.PB == 6 && s.QB == 8) MFT_WG_LAUNCH(6, 8);
    }
    MFT_WG_LAUNCH(0, 0);
#undef MFT_WG_LAUNCH
}

// Test entry: dW[Cout, Cin] += P[R, Cout]^T Q[R, Cin] with plain operands.
int umma_debug_wgrad(const float* P, int ldp, const float* Q, int ldq, float* dW, int ldw, int R, int Cout,
                     int Cin, cudaStream_t st) {
    PlainU p{P, ldp, Cout, plain_vec_ok(P, ldp, Cout)};
    PlainU q{Q, ldq, Cin, plain_vec_ok(Q, ldq, Cin)};
    return umma_wgrad<PlainU, PlainU, false>(p, q, dW, ldw, R, Cout, Cin, st, PC_MISC);
}

bool umma_shape_supported(int F, int nf) {
    if (nf % 16 != 0 || 2 * nf > kMaxC || 2 * nf > UM_MAX_NTILE) return false;
    int n0s[4], nt[4];
    if (F > UM_MAX_KC * UM_KB) return false;
    return plan_passes(2 * nf, F, n0s, nt) > 0 && plan_passes(2 * nf, 2 * nf, n0s, nt) > 0 &&
           plan_passes((F + 15) & ~15, 2 * nf, n0s, nt) > 0;
}

// Region of layer k's weight image inside the workspace: all four layers of one direction are
// resident at once (built by one launch); forward and backward reuse the same regions.
static size_t img_region_floats(int F, int nf, int k) {
    int Cs[5] = {F, 2 * nf, 2 * nf, nf, nf};
    size_t a = umma_wimg_floats(Cs[k + 1], Cs[k]);      // forward: N = C_out, K = C_in
    size_t b = umma_wimg_floats(Cs[k], Cs[k + 1]);      // dgrad:   N = C_in,  K = C_out
    return ((a > b ? a : b) + 255) & ~size_t(255);
}
static size_t img_offset(int F, int nf, int k) {
    size_t off = 0;
    for (int q = 0; q < k; ++q) off += img_region_floats(F, nf, q);
    return off;
}
size_t umma_workspace_floats(int F, int nf) { return img_offset(F, nf, 4); }

// Build the images of all four layers (forward: W[n][k]; backward: W^T for dgrad) with one launch.
static int build_images(const mft_wcompute_params* p, float* wimg, const float* tscale, int F, int nf, bool backward,
                        cudaStream_t st) {
    int Cs[5] = {F, 2 * nf, 2 * nf, nf, nf};
    ImgJobs jobs{};
    int max_total = 0;
    for (int k = 0; k < 4; ++k) {
        const int N = backward ? Cs[k] : Cs[k + 1];
        const int K = backward ? Cs[k + 1] : Cs[k];
        int n0s[4], nts[4];
        int passes = plan_passes(N, K, n0s, nts);
        if (passes == 0) {
            set_error(MFT_ERR_UNSUPPORTED, "tf32 path: no plan for N=%d K=%d", N, K);
            return MFT_ERR_UNSUPPORTED;
        }
        const int KC = cdiv(K, UM_KB);
        for (int q = 0; q < passes; ++q) {
            ImgJob& jb = jobs.j[jobs.n++];
            jb.W = p->conv_w[k];
            jb.scale = tscale + k;
            jb.ldw = Cs[k];                               // conv2d_{k+1}.weight is [Cout, Cin]
            jb.transpose = backward ? 1 : 0;
            jb.N = N; jb.K = K; jb.n0 = n0s[q]; jb.N_TILE = nts[q]; jb.KC = KC;
            jb.img = wimg + img_offset(F, nf, k) + (size_t)q * nts[q] * KC * UM_KB;
            max_total = max(max_total, KC * nts[q] * UM_KB);
        }
    }
    ProfScope ps(PC_PREP, st);
    umma_weight_images_kernel<<<dim3(cdiv(max_total, 256), jobs.n), 256, 0, st>>>(jobs);
    MFT_CHECK_LAUNCH();
    return MFT_OK;
}

int wcompute_bwd_prepare_tf32(const mft_wcompute_params* p, const WcLayout& L, int F, int nf, cudaStream_t st) {
    return build_images(p, L.wimg, L.tscale, F, nf, true, st);   // the scales were saved by the forward
}

int wcompute_fwd_prepare_tf32(const mft_wcompute_params* p, const WcLayout& L, int F, int nf, double count,
                              cudaStream_t st) {
    if (!umma_shape_supported(F, nf)) {
        set_error(MFT_ERR_UNSUPPORTED, "tf32 path: unsupported shape F=%d nf=%d", F, nf);
        return MFT_ERR_UNSUPPORTED;
    }
    {
        ScaleJobs sj{};
        for (int k = 0; k < 4; ++k) {
            sj.W[k] = p->conv_w[k];
            sj.wn[k] = L.C[k + 1] * L.C[k];
            sj.g[k] = k > 0 ? p->bn_g[k - 1] : nullptr;
            sj.b[k] = k > 0 ? p->bn_b[k - 1] : nullptr;
            sj.gc[k] = L.C[k];
            sj.C[k] = L.C[k + 1];
        }
        sj.tscale = L.tscale;
        sj.fsums = L.fsums;
        sj.count = count;
        ProfScope ps(PC_PREP, st);
        umma_layer_scales_kernel<<<4, 1024, 0, st>>>(sj);          // after the memset of the slots on `st`
        MFT_CHECK_LAUNCH();
    }
    return build_images(p, L.wimg, L.tscale, F, nf, false, st);
}

// (the four forward weight images exist: wcompute_fwd_prepare_tf32)
int wcompute_fwd_layers_tf32(const float* x, int ldx, int F, int nf, const mft_wcompute_params* p,
                             const WcLayout& L, const PairGeom& g, cudaStream_t st) {
    if (!umma_shape_supported(F, nf)) {
        set_error(MFT_ERR_UNSUPPORTED, "tf32 path: unsupported shape F=%d nf=%d", F, nf);
        return MFT_ERR_UNSUPPORTED;
    }
    for (int k = 0; k < 4; ++k) {
        double* sums = L.fsums + (size_t)k * kStatSlot;
        EpiFwdStatsU epi{reinterpret_cast<__half*>(L.H[k]), L.C[k + 1], sums, g};
        float* img = L.wimg + img_offset(F, nf, k);
        int rc;
        if (k == 0) {
            AbsDiffU a{x, ldx, F, g, absdiff_vec_ok(x, ldx, F)};
            rc = umma_rows_gemm(a, epi, p->conv_w[0], F, 0, g.R, L.C[1], F, img, st, PC_FWD_L1, true, false);
        } else {
            const double* ps = L.fsums + (size_t)(k - 1) * kStatSlot;
            BnActT a{};
            rc = make_tmap_2d(&a.tmap, L.H[k - 1], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, g.R, L.C[k], L.C[k], UM_KB,
                              UM_ROWS, CU_TENSOR_MAP_SWIZZLE_NONE);
            if (rc != MFT_OK) return rc;
            a.C = L.C[k]; a.sums = ps; a.gamma = p->bn_g[k - 1]; a.beta = p->bn_b[k - 1]; a.inv_count = g.inv_pairs;
            rc = umma_rows_gemm(a, epi, p->conv_w[k], L.C[k], 0, g.R, L.C[k + 1], L.C[k], img, st,
                                PC_FWD_L1 + k, true, true);      // follows the previous layer's launch
        }
        if (rc != MFT_OK) return rc;
    }
    return MFT_OK;
}

// dx[b,n,:] += sum_{m != n} sign(x_n - x_m) * dD[pair{n,m}]  -- the backward of abs() and of the
// broadcast subtraction (gnn.py:79-81) as a gather: one CTA per node, no atomics.  dD holds the
// twin-summed gradient of each unordered pair, so the same formula serves both ends of a pair.
__global__ void __launch_bounds__(256)
dx_gather_kernel(const __nv_bfloat16* __restrict__ dD, int ldd, const float* __restrict__ x, float* __restrict__ dx,
                 int ldx, int F, PairGeom g) {
    pdl_enter();
    const int N = g.N;
    const int node = blockIdx.x;            // b*N + n
    const int b = node / N, n = node - b * N;
    const float* xn = x + (size_t)node * ldx;
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        const float xv = xn[f];
        float acc = 0.f;
        for (int m = 0; m < N; ++m) {
            if (m == n) continue;
            bool shared;
            const int r = pair_row(g, b, min(n, m), max(n, m), shared);
            if (shared && b != 0) continue;   // a shared pair's (graph-summed) gradient lands in graph 0's copy
            const float df = xv - __ldg(x + (size_t)(b * N + m) * ldx + f);
            const float d = __bfloat162float(__ldg(dD + (size_t)r * ldd + f));
            acc += (df > 0.f) ? d : ((df < 0.f) ? -d : 0.f);
        }
        dx[(size_t)node * ldx + f] += acc;
    }
}

// Vector form for 16-byte aligned rows: QF = ceil(F/4) lanes x float4 over the features, 256/QF
// slices over the partner nodes m (independent load streams, combined through shared memory).
__global__ void __launch_bounds__(256)
dx_gather_vec_kernel(const __nv_bfloat16* __restrict__ dD, int ldd, const float* __restrict__ x, float* __restrict__ dx,
                     int ldx, int F, PairGeom g) {
    __shared__ float4 part[256];
    pdl_enter();
    const int N = g.N;
    const int node = blockIdx.x;
    const int b = node / N, n = node - b * N;
    const int QF = (F + 3) >> 2;                 // <= 64
    const int slices = 256 / QF;
    const int tx = threadIdx.x % QF, ty = threadIdx.x / QF;
    const int f = tx * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ty < slices) {
        const float4 xv = ldg4(x + (size_t)node * ldx + f);
        const __nv_bfloat16* Df = dD + f;
        const float* xb = x + (size_t)b * N * ldx + f;
#pragma unroll 4
        for (int m = ty; m < N; m += slices) {
            if (m == n) continue;
            bool shared;
            const int r = pair_row(g, b, min(n, m), max(n, m), shared);
            if (shared && b != 0) continue;
            const float4 xm = ldg4(xb + (size_t)m * ldx);
            const float4 d = unpack_bf4(ldg8(Df + (size_t)r * ldd));
            acc.x += (xv.x > xm.x) ? d.x : ((xv.x < xm.x) ? -d.x : 0.f);
            acc.y += (xv.y > xm.y) ? d.y : ((xv.y < xm.y) ? -d.y : 0.f);
            acc.z += (xv.z > xm.z) ? d.z : ((xv.z < xm.z) ? -d.z : 0.f);
            acc.w += (xv.w > xm.w) ? d.w : ((xv.w < xm.w) ? -d.w : 0.f);
        }
    }
    part[threadIdx.x] = acc;
    __syncthreads();
    if (ty == 0) {
        for (int sl = 1; sl < slices; ++sl) {
            const float4 v = part[sl * QF + tx];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        float* o = dx + (size_t)node * ldx + f;
        const float v[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (f + e < F) o[e] += v[e];
    }
}

// dgrad of conv layer k (0-based): dL/d a_k = dH_k W_k on tensor cores; the epilogue turns it into
// dy_{k-1} (+ BN-backward reductions) or, for layer 0, into dD followed by the dx gather.
int wcompute_bwd_layer_tf32(int k, float* dh, float* dy_next, const float* x, int ldx, float* dx, int F, int nf,
                            const mft_wcompute_params* p, const mft_wcompute_grads* gr, const WcLayout& L,
                            const PairGeom& g, cudaStream_t st) {
    (void)gr;
    const int Cout = L.C[k + 1], Cin = L.C[k];
    DhInPlaceT a{};
    {
        int rc = make_tmap_2d(&a.tmap_dy, dh, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, g.R, Cout, Cout, UM_KB, UM_ROWS,
                              CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc != MFT_OK) return rc;
        rc = make_tmap_2d(&a.tmap, L.H[k], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, g.R, Cout, Cout, UM_KB, UM_ROWS,
                          CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc != MFT_OK) return rc;
        a.C = Cout; a.fsums = L.fsums + (size_t)k * kStatSlot; a.gamma = p->bn_g[k];
        a.bsums = L.bsums + (size_t)k * kStatSlot; a.inv_count = g.inv_pairs; a.g = g;
    }
    if (k == 0) {
        const int ldd = (F + 3) & ~3;
        EpiStoreBf16U e{reinterpret_cast<__nv_bfloat16*>(L.dD), ldd};
        int rc = umma_rows_gemm(a, e, p->conv_w[0], Cin, 1, g.R, Cin, Cout, L.wimg + img_offset(F, nf, 0), st,
                                PC_DGRAD_L1, true, true);        // follows this layer's wgrad launch
        if (rc != MFT_OK) return rc;
        ProfScope ps(PC_DGRAD_L1, st);
        const bool gpdl = pdl_level() >= 2;
        // x rows 16-byte aligned and padded to a multiple of 4 floats (always true for the xcat of gnn_fwd):
        // beyond-F lanes of the last float4 read padding that is masked on the way out
        if (absdiff_vec_ok(x, ldx, F) && F <= 256)
            MFT_CHECK_CUDA(launch_kernel(dx_gather_vec_kernel, dim3(g.B * g.N), dim3(256), 0, st, gpdl,
                                         reinterpret_cast<const __nv_bfloat16*>(L.dD), ldd, x, dx, ldx, F, g));
        else
            MFT_CHECK_CUDA(launch_kernel(dx_gather_kernel, dim3(g.B * g.N), dim3(256), 0, st, gpdl,
                                         reinterpret_cast<const __nv_bfloat16*>(L.dD), ldd, x, dx, ldx, F, g));
        return MFT_OK;
    }
    const double* ps = L.fsums + (size_t)(k - 1) * kStatSlot;
    double* pbs = L.bsums + (size_t)(k - 1) * kStatSlot;
    EpiDyU e{reinterpret_cast<const __half*>(L.H[k - 1]), dy_next, Cin, ps, p->bn_g[k - 1], p->bn_b[k - 1], g.inv_pairs, pbs};
    return umma_rows_gemm(a, e, p->conv_w[k], Cin, 1, g.R, Cin, Cout, L.wimg + img_offset(F, nf, k), st,
                          PC_DGRAD_L1 + k, true, true);          // follows this layer's wgrad launch
}

// wgrad of conv layer k: d conv2d_{k+1}.weight [Cout, Cin] += dH_k^T a_k (a_0 = |x_i - x_j|).
int wcompute_wgrad_layer_tf32(int k, const float* dh, const float* x, int ldx, int F,
                              const mft_wcompute_params* p, const mft_wcompute_grads* gr, const WcLayout& L,
                              const PairGeom& g, int* copies, cudaStream_t st) {
    const int Cout = L.C[k + 1], Cin = L.C[k];
    // layer 3 follows the dy8 row kernel, the others the dgrad launch of the layer above
    const bool pdl = pdl_level() >= (k == 3 ? 2 : 1);
    DhT P{};
    int rc = make_tmap_2d(&P.tmap_dy, dh, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, g.R, Cout, Cout, Cout, WG_ROWS,
                          CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc != MFT_OK) return rc;
    rc = make_tmap_2d(&P.tmap_h, L.H[k], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, g.R, Cout, Cout, Cout, WG_ROWS,
                      CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc != MFT_OK) return rc;
    P.C = Cout; P.fsums = L.fsums + (size_t)k * kStatSlot; P.gamma = p->bn_g[k];
    P.bsums = L.bsums + (size_t)k * kStatSlot; P.inv_count = g.inv_pairs; P.g = g;
    if (k == 0) {
        AbsDiffU Q{x, ldx, F, g, absdiff_vec_ok(x, ldx, F)};
        return umma_wgrad(P, Q, L.wgpart + L.wgpart_off[0], (Cin + 3) & ~3, g.R, Cout, Cin, st, PC_WGRAD_L1, copies, pdl);
    }
    BnActQT Q{};
    rc = make_tmap_2d(&Q.tmap_h, L.H[k - 1], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, g.R, Cin, Cin, Cin, WG_ROWS,
                      CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc != MFT_OK) return rc;
    Q.C = Cin; Q.sums = L.fsums + (size_t)(k - 1) * kStatSlot; Q.gamma = p->bn_g[k - 1]; Q.beta = p->bn_b[k - 1];
    Q.inv_count = g.inv_pairs;
    return umma_wgrad(P, Q, L.wgpart + L.wgpart_off[k], (Cin + 3) & ~3, g.R, Cout, Cin, st, PC_WGRAD_L1 + k, copies, pdl);
}

// Debug / test entry: C[M, N] = A[M, K] * op(W)^T through the tcgen05 rows kernel with plain
// operands (tests/test_gpu_umma.py checks it against an fp32 product).
int umma_debug_gemm(const float* A, int lda, const float* W, int ldw, int transpose_w, float* C, int ldc, int M,
                    int N, int K, float* wimg, cudaStream_t st) {
    PlainU a{A, lda, K, plain_vec_ok(A, lda, K)};
    EpiStoreU e{C, ldc, (ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0) ? 1 : 0};
    return umma_rows_gemm(a, e, W, ldw, transpose_w, M, N, K, wimg, st, PC_MISC);
}

}  // namespace mft
