// tcgen05 (MFT_PREC_TF32) edge-MLP layers.
//
// One persistent, warp-specialised kernel template covers every "pair rows x weights" GEMM of
// the edge MLP (the four forward 1x1-conv layers and the four dgrad products):
//
//     C[r, n] = sum_k A(r, k) * W(n, k)        r: 128-row tiles of unordered pairs
//
//   warps 0-3  producers : build the A tile from global memory -- |x_i - x_j| for layer 1,
//                          LeakyReLU(BN(H)) for later layers, plain loads for dgrad -- round it to
//                          TF32 and write it K-major / SWIZZLE_128B into a ring of 16 KB K blocks,
//                          so neither the N^2 x C pair tensor nor any post-BN activation ever
//                          exists in HBM;
//   warp  8    MMA issuer: one thread; weights (pre-swizzled TF32 image) land once per CTA by
//                          bulk async copy and stay resident; tcgen05.mma kind::tf32, M=128,
//                          N = tile width, accumulators double-buffered in TMEM;
//   warps 4-7  epilogue  : tcgen05.ld 32 columns at a time -> padded smem staging -> coalesced
//                          128-byte row segments to global, fused with the per-channel batch
//                          statistics (forward) or with the LeakyReLU'/BN-backward reductions (dgrad).
//
// Pipelines: full/empty mbarriers per A stage (producers <-> MMA), tmem_full/tmem_empty per
// accumulator (MMA <-> epilogue).  Every wait is bounded (umma::mbar_wait traps on timeout).
#include <cstdio>

#include "common.cuh"
#include "wcompute.cuh"
#include "umma.cuh"
#include "prof.cuh"

namespace mft {

using namespace umma;

constexpr int UM_ROWS = 128;
constexpr int UM_KB = 32;                 // tf32 elements per K block (128 bytes)
constexpr int UM_BLOCK_FLOATS = UM_ROWS * UM_KB;   // one A stage = 4096 floats = 16 KB
constexpr int UM_THREADS = 288;           // 4 producer + 4 epilogue + 1 MMA warp
constexpr int UM_STAGE_LD = 36;           // padded row length of the epilogue staging tile
constexpr int UM_ACC_STRIDE = 256;        // TMEM columns between the two accumulators
constexpr int UM_TMEM_COLS = 512;
constexpr int UM_MAX_STAGES = 4;
constexpr int UM_MAX_CHUNKS = 8;          // 32-column epilogue chunks (N_TILE <= 256)
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
};

static inline size_t umma_smem_bytes(const UmmaShape& s) {
    return 1024 + (size_t)s.KC * s.N_TILE * 128 + (size_t)s.stages * UM_BLOCK_FLOATS * 4 +
           (size_t)UM_ROWS * UM_STAGE_LD * 4 + (3 + 4) * kMaxC * 4 + UM_ROWS * 4 + 2 * 256 * 4 + 16 * 8 + 16;
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

// ------------------------------------------------------------------ producer functors
// load4(row, k): four consecutive K elements starting at k (k % 4 == 0), zero beyond K.

struct AbsDiffU {
    const float* x;
    int ldx, F;
    PairGeom g;
    int vec_ok;
    struct Row { const float* xi; const float* xj; };
    __device__ __forceinline__ void init(float*, int, int) const {}
    __device__ __forceinline__ Row row(int r) const {
        PairRow p = decode_row(r, g);
        return Row{x + (size_t)(p.b * g.N + p.i) * ldx, x + (size_t)(p.b * g.N + p.j) * ldx};
    }
    __device__ __forceinline__ float4 load4(const Row& rw, int k, const float*) const {
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k >= F) return o;
        if (vec_ok && k + 3 < F) {
            float4 a = __ldg(reinterpret_cast<const float4*>(rw.xi + k));
            float4 b = __ldg(reinterpret_cast<const float4*>(rw.xj + k));
            return make_float4(fabsf(a.x - b.x), fabsf(a.y - b.y), fabsf(a.z - b.z), fabsf(a.w - b.w));
        }
        o.x = fabsf(__ldg(rw.xi + k) - __ldg(rw.xj + k));
        if (k + 1 < F) o.y = fabsf(__ldg(rw.xi + k + 1) - __ldg(rw.xj + k + 1));
        if (k + 2 < F) o.z = fabsf(__ldg(rw.xi + k + 2) - __ldg(rw.xj + k + 2));
        if (k + 3 < F) o.w = fabsf(__ldg(rw.xi + k + 3) - __ldg(rw.xj + k + 3));
        return o;
    }
};

// a = LeakyReLU(scale*h + shift), scale = gamma*rstd, shift = beta - mean*scale (C % 4 == 0)
struct BnActU {
    const float* H;
    int C;
    const double* sums;
    const float* gamma;
    const float* beta;
    double inv_count;
    struct Row { const float* h; };
    __device__ __forceinline__ void init(float* aux, int tid, int nthreads) const {
        for (int c = tid; c < C; c += nthreads) {
            float m, r;
            bn_mean_rstd(sums, C, c, inv_count, m, r);
            float sc = gamma[c] * r;
            aux[c] = sc;
            aux[kMaxC + c] = beta[c] - m * sc;
        }
    }
    __device__ __forceinline__ Row row(int r) const { return Row{H + (size_t)r * C}; }
    __device__ __forceinline__ float4 load4(const Row& rw, int k, const float* aux) const {
        if (k >= C) return make_float4(0.f, 0.f, 0.f, 0.f);
        float4 h = *reinterpret_cast<const float4*>(rw.h + k);
        float4 sc = *reinterpret_cast<const float4*>(aux + k);
        float4 sh = *reinterpret_cast<const float4*>(aux + kMaxC + k);
        float4 y;
        y.x = fmaf(h.x, sc.x, sh.x); y.y = fmaf(h.y, sc.y, sh.y);
        y.z = fmaf(h.z, sc.z, sh.z); y.w = fmaf(h.w, sc.w, sh.w);
        return make_float4(fmaxf(y.x, kSlope * y.x), fmaxf(y.y, kSlope * y.y), fmaxf(y.z, kSlope * y.z),
                           fmaxf(y.w, kSlope * y.w));
    }
};

struct PlainU {
    const float* p;
    int ld, K;
    int vec_ok;
    struct Row { const float* q; };
    __device__ __forceinline__ void init(float*, int, int) const {}
    __device__ __forceinline__ Row row(int r) const { return Row{p + (size_t)r * ld}; }
    __device__ __forceinline__ float4 load4(const Row& rw, int k, const float*) const {
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k >= K) return o;
        if (vec_ok && k + 3 < K) return *reinterpret_cast<const float4*>(rw.q + k);
        o.x = rw.q[k];
        if (k + 1 < K) o.y = rw.q[k + 1];
        if (k + 2 < K) o.z = rw.q[k + 2];
        if (k + 3 < K) o.w = rw.q[k + 3];
        return o;
    }
};

// ------------------------------------------------------------------ epilogue functors
// apply(): four consecutive output columns col..col+3 (col % 4 == 0) of global row r;
// nvalid = how many of them exist.  s0/s1: the thread's running column statistics.

struct EpiStoreU {
    static constexpr bool kStats = false;
    static constexpr bool kRowWeight = false;
    float* out;
    int ld;
    int vec_ok;
    __device__ __forceinline__ void init(float*, int, int) const {}
    __device__ __forceinline__ float row_weight(int) const { return 1.f; }
    __device__ __forceinline__ void apply(int r, float, int col, float4 v, int nvalid, float*, float*,
                                          const float*) const {
        float* o = out + (size_t)r * ld + col;
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

// forward: store pre-BN H, accumulate sum w*h and sum w*h^2 (C % 4 == 0)
struct EpiFwdStatsU {
    static constexpr bool kStats = true;
    static constexpr bool kRowWeight = true;
    float* H;
    int C;
    double* sums;
    PairGeom g;
    __device__ __forceinline__ void init(float*, int, int) const {}
    __device__ __forceinline__ float row_weight(int r) const { return decode_row(r, g).w; }
    __device__ __forceinline__ void apply(int r, float w, int col, float4 v, int, float* s0, float* s1,
                                          const float*) const {
        *reinterpret_cast<float4*>(H + (size_t)r * C + col) = v;
        s0[0] = fmaf(w, v.x, s0[0]); s1[0] = fmaf(w * v.x, v.x, s1[0]);
        s0[1] = fmaf(w, v.y, s0[1]); s1[1] = fmaf(w * v.y, v.y, s1[1]);
        s0[2] = fmaf(w, v.z, s0[2]); s1[2] = fmaf(w * v.z, v.z, s1[2]);
        s0[3] = fmaf(w, v.w, s0[3]); s1[3] = fmaf(w * v.w, v.w, s1[3]);
    }
    __device__ __forceinline__ void commit(int c, float v0, float v1, const float*) const {
        atomicAdd(sums + c, (double)v0);
        atomicAdd(sums + C + c, (double)v1);
    }
};

// dgrad: acc = dL/d a_{k-1}; dy = acc * lrelu'(BN(H_{k-1})), stored; reductions sum dy and
// sum dy*hhat (accumulated as sum dy*h and corrected per CTA: hhat = (h - mean) * rstd)
struct EpiDyU {
    static constexpr bool kStats = true;
    static constexpr bool kRowWeight = false;
    const float* H;
    float* dy;
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
    __device__ __forceinline__ void apply(int r, float, int col, float4 v, int, float* s0, float* s1,
                                          const float* aux) const {
        float4 h = *reinterpret_cast<const float4*>(H + (size_t)r * C + col);
        float4 sc = *reinterpret_cast<const float4*>(aux + col);
        float4 sh = *reinterpret_cast<const float4*>(aux + kMaxC + col);
        float4 d;
        d.x = v.x * (fmaf(h.x, sc.x, sh.x) > 0.f ? 1.f : kSlope);
        d.y = v.y * (fmaf(h.y, sc.y, sh.y) > 0.f ? 1.f : kSlope);
        d.z = v.z * (fmaf(h.z, sc.z, sh.z) > 0.f ? 1.f : kSlope);
        d.w = v.w * (fmaf(h.w, sc.w, sh.w) > 0.f ? 1.f : kSlope);
        *reinterpret_cast<float4*>(dy + (size_t)r * C + col) = d;
        s0[0] += d.x; s1[0] = fmaf(d.x, h.x, s1[0]);
        s0[1] += d.y; s1[1] = fmaf(d.y, h.y, s1[1]);
        s0[2] += d.z; s1[2] = fmaf(d.z, h.z, s1[2]);
        s0[3] += d.w; s1[3] = fmaf(d.w, h.w, s1[3]);
    }
    __device__ __forceinline__ void commit(int c, float v0, float v1, const float* aux) const {
        float m = aux[2 * kMaxC + c], r = aux[3 * kMaxC + c];
        atomicAdd(bsums + c, (double)v0);
        atomicAdd(bsums + C + c, (double)(r * (v1 - m * v0)));
    }
};

// ------------------------------------------------------------------ the kernel
template <class AOp, class Epi>
__global__ void __launch_bounds__(UM_THREADS, 1)
umma_rows_kernel(AOp aop, Epi epi, const float* __restrict__ wimg, UmmaShape s) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t w_bytes = (uint32_t)s.KC * s.N_TILE * 128;
    float* Wsm = reinterpret_cast<float*>(smem);
    float* Asm = reinterpret_cast<float*>(smem + w_bytes);
    float* stage = Asm + (size_t)s.stages * UM_BLOCK_FLOATS;
    float* aux_a = stage + UM_ROWS * UM_STAGE_LD;
    float* aux_e = aux_a + 3 * kMaxC;
    float* wrow = aux_e + 4 * kMaxC;
    float* red0 = wrow + UM_ROWS;
    float* red1 = red0 + 256;
    uint64_t* bars = reinterpret_cast<uint64_t*>(red1 + 256);
    uint64_t* full = bars;          // [UM_MAX_STAGES]
    uint64_t* empty = bars + 4;     // [UM_MAX_STAGES]
    uint64_t* tfull = bars + 8;     // [2]
    uint64_t* tempty = bars + 10;   // [2]
    uint64_t* wbar = bars + 12;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int ntiles = (s.R + UM_ROWS - 1) / UM_ROWS;

    if (tid == 0) {
        for (int i = 0; i < UM_MAX_STAGES; ++i) {
            mbar_init(&full[i], 128);
            mbar_init(&empty[i], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], 128);
        }
        mbar_init(wbar, 1);
        fence_mbar_init();
    }
    if (warp == 8) tmem_alloc(tmem_slot, UM_TMEM_COLS);
    aop.init(aux_a, tid, UM_THREADS);
    epi.init(aux_e, tid, UM_THREADS);
    for (int c = tid; c < 512; c += UM_THREADS) red0[c] = 0.f;   // red0 and red1 are contiguous
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // ===================== producers =====================
        const int rsub = tid >> 3, c16 = tid & 7;
        const int sw = rsub & 7;
        int st = 0;
        uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int row0 = tile * UM_ROWS;
            typename AOp::Row rc[8];
            bool valid[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                int r = row0 + q * 16 + rsub;
                valid[q] = r < s.R;
                rc[q] = aop.row(valid[q] ? r : 0);
            }
            for (int kc = 0; kc < s.KC; ++kc) {
                mbar_wait(&empty[st], ph ^ 1);
                float* dst = Asm + (size_t)st * UM_BLOCK_FLOATS;
                const int k = kc * UM_KB + c16 * 4;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int rl = q * 16 + rsub;
                    float4 v = valid[q] ? aop.load4(rc[q], k, aux_a) : make_float4(0.f, 0.f, 0.f, 0.f);
                    v.x = to_tf32(v.x); v.y = to_tf32(v.y); v.z = to_tf32(v.z); v.w = to_tf32(v.w);
                    *reinterpret_cast<float4*>(dst + (rl >> 3) * 256 + sw * 32 + ((c16 ^ sw) << 2)) = v;
                }
                fence_proxy_async_smem();
                mbar_arrive(&full[st]);
                if (++st == s.stages) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 8) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            mbar_arrive_expect_tx(wbar, w_bytes);
            const uint32_t blk = (uint32_t)s.N_TILE * 128;
            for (int kc = 0; kc < s.KC; ++kc)
                bulk_g2s(reinterpret_cast<uint8_t*>(Wsm) + (size_t)kc * blk,
                         reinterpret_cast<const uint8_t*>(wimg) + (size_t)kc * blk, blk, wbar);
            mbar_wait(wbar, 0);
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
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue =====================
        const int et = tid - 128;          // 0..127 = row of the tile this thread drains from TMEM
        const int ew = warp & 3;           // TMEM lane window of this warp
        const int rsub = et >> 3, c4 = (et & 7) * 4;
        const int nchunks = (s.N_TILE + 31) / 32;
        float s0[UM_MAX_CHUNKS][4], s1[UM_MAX_CHUNKS][4];
#pragma unroll
        for (int ch = 0; ch < UM_MAX_CHUNKS; ++ch)
#pragma unroll
            for (int e = 0; e < 4; ++e) { s0[ch][e] = 0.f; s1[ch][e] = 0.f; }
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            const int row0 = tile * UM_ROWS;
            if (Epi::kRowWeight) wrow[et] = (row0 + et < s.R) ? epi.row_weight(row0 + et) : 0.f;
            mbar_wait(&tfull[acc], aph);
            tc_fence_after_sync();
#pragma unroll
            for (int ch = 0; ch < UM_MAX_CHUNKS; ++ch) {
                if (ch < nchunks) {
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + acc * UM_ACC_STRIDE + ch * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        *reinterpret_cast<float4*>(stage + et * UM_STAGE_LD + q * 4) =
                            make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                        __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
                    named_bar_sync(1, 128);
                    const int cl = ch * 32 + c4;          // column inside this pass
                    const int col = s.n0 + cl;            // global output column
                    if (cl < s.N_TILE && col < s.N) {
                        const int nvalid = min(4, s.N - col);
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const int rl = q * 16 + rsub;
                            const int r = row0 + rl;
                            if (r < s.R) {
                                float4 a = *reinterpret_cast<const float4*>(stage + rl * UM_STAGE_LD + c4);
                                epi.apply(r, Epi::kRowWeight ? wrow[rl] : 1.f, col, a, nvalid, s0[ch], s1[ch],
                                          aux_e);
                            }
                        }
                    }
                    named_bar_sync(1, 128);
                }
            }
            tc_fence_before_sync();
            mbar_arrive(&tempty[acc]);
        }
        if (Epi::kStats) {
#pragma unroll
            for (int ch = 0; ch < UM_MAX_CHUNKS; ++ch) {
                if (ch < nchunks) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        int cl = ch * 32 + c4 + e;
                        if (cl < 256) {
                            atomicAdd(&red0[cl], s0[ch][e]);
                            atomicAdd(&red1[cl], s1[ch][e]);
                        }
                    }
                }
            }
            named_bar_sync(1, 128);
            for (int cl = et; cl < s.N_TILE; cl += 128)
                if (s.n0 + cl < s.N) epi.commit(s.n0 + cl, red0[cl], red1[cl], aux_e);
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, UM_TMEM_COLS);
    }
}

// ------------------------------------------------------------------ host side
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

constexpr size_t kSmemLimit = 232448;   // 227 KB opt-in maximum per CTA on sm_100

static bool plan_pass(int N_TILE, int K, UmmaShape& s) {
    s.N_TILE = N_TILE;
    s.KC = (K + UM_KB - 1) / UM_KB;
    if (s.KC > UM_MAX_KC || N_TILE > UM_MAX_NTILE || (N_TILE % 16) != 0) return false;
    for (int st = UM_MAX_STAGES; st >= 2; --st) {
        s.stages = st;
        if (umma_smem_bytes(s) <= kSmemLimit) return true;
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
                          int K, float* wimg, cudaStream_t st, int cat) {
    int n0s[4], nts[4];
    int passes = plan_passes(N, K, n0s, nts);
    if (passes == 0) {
        set_error(MFT_ERR_UNSUPPORTED, "umma_rows_gemm: no shared-memory plan for N=%d K=%d", N, K);
        return MFT_ERR_UNSUPPORTED;
    }
    const int ntiles = cdiv(R, UM_ROWS);
    const int grid = min(ntiles, num_sms());
    for (int p = 0; p < passes; ++p) {
        UmmaShape s{};
        plan_pass(nts[p], K, s);
        s.R = R; s.N = N; s.n0 = n0s[p]; s.K = K;
        float* img = wimg + (size_t)p * s.N_TILE * s.KC * UM_KB;
        {
            ProfScope ps(PC_PREP, st);
            int total = s.KC * s.N_TILE * UM_KB;
            umma_weight_image_kernel<<<cdiv(total, 256), 256, 0, st>>>(W, ldw, transpose, N, K, s.n0, s.N_TILE,
                                                                       s.KC, img);
            MFT_CHECK_LAUNCH();
        }
        size_t smem = umma_smem_bytes(s);
        static_assert(sizeof(AOp) + sizeof(Epi) < 3500, "kernel parameter space");
        MFT_CHECK_CUDA(cudaFuncSetAttribute(umma_rows_kernel<AOp, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)smem));
        ProfScope ps(cat, st);
        umma_rows_kernel<AOp, Epi><<<grid, UM_THREADS, smem, st>>>(aop, epi, img, s);
        MFT_CHECK_LAUNCH();
    }
    return MFT_OK;
}

bool umma_shape_supported(int F, int nf) {
    if (nf % 16 != 0 || 2 * nf > kMaxC || 2 * nf > UM_MAX_NTILE) return false;
    int n0s[4], nt[4];
    if (F > UM_MAX_KC * UM_KB) return false;
    return plan_passes(2 * nf, F, n0s, nt) > 0 && plan_passes(2 * nf, 2 * nf, n0s, nt) > 0 &&
           plan_passes((F + 15) & ~15, 2 * nf, n0s, nt) > 0;
}

size_t umma_workspace_floats(int F, int nf) {
    // one weight image at a time (stream ordered): the largest of the forward / dgrad operands
    size_t m = 0;
    int Cs[5] = {F, 2 * nf, 2 * nf, nf, nf};
    for (int k = 0; k < 4; ++k) {
        size_t a = umma_wimg_floats(Cs[k + 1], Cs[k]);      // forward: N = C_out, K = C_in
        size_t b = umma_wimg_floats(Cs[k], Cs[k + 1]);      // dgrad:   N = C_in,  K = C_out
        m = m > a ? m : a;
        m = m > b ? m : b;
    }
    return m;
}

int wcompute_fwd_layers_tf32(const float* x, int ldx, int F, int nf, const mft_wcompute_params* p,
                             const WcLayout& L, const PairGeom& g, cudaStream_t st) {
    if (!umma_shape_supported(F, nf)) {
        set_error(MFT_ERR_UNSUPPORTED, "tf32 path: unsupported shape F=%d nf=%d", F, nf);
        return MFT_ERR_UNSUPPORTED;
    }
    for (int k = 0; k < 4; ++k) {
        double* sums = L.fsums + (size_t)k * 2 * kMaxC;
        EpiFwdStatsU epi{L.H[k], L.C[k + 1], sums, g};
        int rc;
        if (k == 0) {
            AbsDiffU a{x, ldx, F, g, (ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) ? 1 : 0};
            rc = umma_rows_gemm(a, epi, p->conv_w[0], F, 0, g.R, L.C[1], F, L.wimg, st, PC_FWD_L1);
        } else {
            const double* ps = L.fsums + (size_t)(k - 1) * 2 * kMaxC;
            BnActU a{L.H[k - 1], L.C[k], ps, p->bn_g[k - 1], p->bn_b[k - 1], g.inv_pairs};
            rc = umma_rows_gemm(a, epi, p->conv_w[k], L.C[k], 0, g.R, L.C[k + 1], L.C[k], L.wimg, st,
                                PC_FWD_L1 + k);
        }
        if (rc != MFT_OK) return rc;
    }
    return MFT_OK;
}

// dx[b,n,:] += sum_{m != n} sign(x_n - x_m) * dD[pair{n,m}]  -- the backward of abs() and of the
// broadcast subtraction (gnn.py:79-81) as a gather: one CTA per node, no atomics.  dD holds the
// twin-summed gradient of each unordered pair, so the same formula serves both ends of a pair.
__global__ void __launch_bounds__(256)
dx_gather_kernel(const float* __restrict__ dD, int ldd, const float* __restrict__ x, float* __restrict__ dx,
                 int ldx, int F, int N, int Rg) {
    const int node = blockIdx.x;            // b*N + n
    const int b = node / N, n = node - b * N;
    const float* xn = x + (size_t)node * ldx;
    const float* Db = dD + (size_t)b * Rg * ldd;
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        const float xv = xn[f];
        float acc = 0.f;
        for (int m = 0; m < N; ++m) {
            if (m == n) continue;
            const int i = min(n, m), j = max(n, m);
            const int r = tri_start(i, N) + (j - i);
            const float df = xv - __ldg(x + (size_t)(b * N + m) * ldx + f);
            const float d = __ldg(Db + (size_t)r * ldd + f);
            acc += (df > 0.f) ? d : ((df < 0.f) ? -d : 0.f);
        }
        dx[(size_t)node * ldx + f] += acc;
    }
}

// dgrad of conv layer k (0-based): dL/d a_k = dH_k W_k on tensor cores; the epilogue turns it into
// dy_{k-1} (+ BN-backward reductions) or, for layer 0, into dD followed by the dx gather.
int wcompute_bwd_layer_tf32(int k, float* dh, float* dy_next, const float* x, int ldx, float* dx, int F, int nf,
                            const mft_wcompute_params* p, const mft_wcompute_grads* gr, const WcLayout& L,
                            const PairGeom& g, cudaStream_t st) {
    (void)gr; (void)nf;
    const int Cout = L.C[k + 1], Cin = L.C[k];
    PlainU a{dh, Cout, Cout, 1};
    if (k == 0) {
        const int ldd = (F + 3) & ~3;
        EpiStoreU e{L.dD, ldd, 1};
        int rc = umma_rows_gemm(a, e, p->conv_w[0], Cin, 1, g.R, Cin, Cout, L.wimg, st, PC_DGRAD_L1);
        if (rc != MFT_OK) return rc;
        ProfScope ps(PC_DGRAD_L1, st);
        dx_gather_kernel<<<g.B * g.N, 256, 0, st>>>(L.dD, ldd, x, dx, ldx, F, g.N, g.Rg);
        MFT_CHECK_LAUNCH();
        return MFT_OK;
    }
    const double* ps = L.fsums + (size_t)(k - 1) * 2 * kMaxC;
    double* pbs = L.bsums + (size_t)(k - 1) * 2 * kMaxC;
    EpiDyU e{L.H[k - 1], dy_next, Cin, ps, p->bn_g[k - 1], p->bn_b[k - 1], g.inv_pairs, pbs};
    return umma_rows_gemm(a, e, p->conv_w[k], Cin, 1, g.R, Cin, Cout, L.wimg, st, PC_DGRAD_L1 + k);
}

// Debug / test entry: C[M, N] = A[M, K] * op(W)^T through the tcgen05 rows kernel with plain
// operands (tests/test_gpu_umma.py checks it against an fp32 product).
int umma_debug_gemm(const float* A, int lda, const float* W, int ldw, int transpose_w, float* C, int ldc, int M,
                    int N, int K, float* wimg, cudaStream_t st) {
    PlainU a{A, lda, K, (lda % 4 == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0) ? 1 : 0};
    EpiStoreU e{C, ldc, (ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0) ? 1 : 0};
    return umma_rows_gemm(a, e, W, ldw, transpose_w, M, N, K, wimg, st, PC_MISC);
}

}  // namespace mft
