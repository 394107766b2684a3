// screen.cu -- tcgen05 distance screen + exact fp32 verify (K2 / K2v of SURVEY 2.2).
//
// Replaces the N x k distance evaluations of deeptime assign_chunk_to_centers / kmeans.cluster
// (call sites pyemma/coordinates/clustering/interface.py:164-165, kmeans.py:254-258) when k*d is
// large.  The reference's argmin must be reproduced bit for bit, and a direct evaluation in the
// reference's operation order costs 3 CUDA-core instructions per pair-dimension.  So:
//
//  1. screen   acc[i][j] = x~_i . c~_j - |c~_j|^2/2   (maximising acc == minimising the distance)
//              as ONE fp16 GEMM on the 5th-gen tensor cores: tcgen05.mma, operands staged by TMA
//              (SWIZZLE_128B), fp32 accumulators in TMEM.  x~ = (x-mu)*sigma is centred/scaled (power of
//              two, max |.| <= 2048) so it fits fp16; precision comes from splitting each fp32 value into
//              hi+lo fp16 parts and concatenating along K (terms=3; terms=1 keeps only the first segment):
//                 A' = [ x_hi*2^-5 | x_lo*2^5  | x_hi | 2^8  2^-3  2^-14 ]
//                 B' = [ c_lo*2^5  | c_hi*2^-5 | c_hi | -b1  -b2   -b3   ],  b = |c~|^2/2 = b1*2^8+b2*2^-3+b3*2^-14
//              (small cross terms first: while they accumulate the partial sums stay ~2^-11 of the final
//              magnitude, so only the K=16 steps that touch the hi.hi segment or the bias contribute to the
//              accumulation-error bound)
//              The power-of-two factors keep the lo parts out of the fp16 subnormal range, and every value
//              below 2^-14 is flushed to zero HERE (measured: tcgen05.mma kind::f16 does not honour fp16
//              subnormals), so the flush error is part of the margin instead of a silent loss.
//              The N x k score matrix is never written: 8 epilogue warps read the accumulators with
//              tcgen05.ld (software-pipelined), keep a running row maximum (FMNMX3) and remember every
//              32-column chunk -- with a 4-bit mask of its 8-column groups -- whose maximum is within a
//              rigorous error margin of it.
//  2. verify   the (few) surviving chunks are re-evaluated per frame with the exact reference-order
//              fp32 kernel arithmetic (common.cuh) -> the argmin matches the reference exactly.
//              Frames whose candidate list overflows fall back to a full exact scan.
//
// Error model (DESIGN.md "Screen margin"): with X=|x~|, C=max|c~|, u=2^-24, h=2^-11
//   d1 (centring in fp32)        <= 2.1 u (X+C)^2
//   d2 (operand split + bias + tensor-core accumulation; assumes each K=16 MMA step loses at most
//       17*2^-23 of its largest addend)
//   rho (the reference's own rounding + sqrt merging) ~ 2(d/4+9)u + 2^-21
//   a frame's reference label j* satisfies  acc[j*] >= max_j acc[j] - T,
//   T = (2 d2 + d1) + rho/2 * (|x~|^2 - 2 max acc + 2 d2 + d1).
#include "common.cuh"
#include "kernels.h"
#include <cuda.h>
#include <cuda_fp16.h>

namespace b2k {

static constexpr int TILE_M = 128;    // frames per CTA tile (UMMA M)
static constexpr int TILE_N = 256;    // centers per accumulator stage (UMMA N)
static constexpr int BLOCK_K = 64;    // fp16 elements per k-block (one 128-byte swizzle row)
static constexpr int STAGES = 4;
static constexpr int CHUNK = 32;      // columns per candidate chunk (one tcgen05.ld.32x32b.x32)
static constexpr int GROUP = 8;       // center rows per block of the verify kernels' tables / staged slabs
// A candidate entry is (chunk id | group mask << id bits): the chunk's 32 columns are split into groups of CG = 8, 4
// or 2 centers (4-, 8- or 16-bit mask).  Finer groups cost the epilogue a few more compares per chunk and save the
// verify kernels most of their center traffic (they re-evaluate whole groups).
static constexpr int LIST_CAP = 16;   // running candidate chunks remembered per frame and column half
static constexpr int CAND_CAP = 8;    // candidate chunks handed to the verify kernel per frame
__host__ __device__ constexpr int cand_id_bits(int cg) { return 32 - CHUNK / cg; }
__host__ __device__ constexpr uint32_t cand_id_mask(int cg) { return (1u << cand_id_bits(cg)) - 1u; }
__device__ __forceinline__ int nth_set_bit(uint32_t m, int n) {  // position of the n-th (0-based) set bit
    for (; n > 0; --n) m &= m - 1;
    return __ffs(m) - 1;
}
static constexpr int A_BYTES = TILE_M * BLOCK_K * 2;
static constexpr int B_BYTES = TILE_N * BLOCK_K * 2;
static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;

struct ScreenParams {  // device resident; written by the prep kernels, read by everything else
    float sigma;       // power-of-two scale
    float cmax;        // upper bound of max_j |c~_j|
    float xmax2_raw;   // max_i |x_i - mu|^2 over the prepared frames
    float cmax2_raw;   // max_j |c_j - mu|^2 at prepare time
    float cmax2_now;   // max_j |c~_j|^2 of the current B' operand
    float cl2_now;     // max_j |c~_j - hi(c~_j)|^2 of the current B' operand (what a hi-only center operand drops)
    int valid;         // 0: operands unusable -> every frame takes the exact fallback
    unsigned long long cand_chunks, fallback_frames;  // statistics of the last verify
    unsigned int fb_count, fb_pad;                    // frames handed to the exact fallback kernel
    float cl;          // upper bound of max_j |c~_j - hi(c~_j)|
};

struct ScreenPlan {
    b2k_ctx* ctx = nullptr;
    int64_t n_cap = 0, n_pad = 0;
    int d = 0, k = 0, k_pad = 0, terms = 3, Kc = 0, Kp = 0, nk16 = 0;
    int cg = 4;                // centers per candidate group (8, 4 or 2)
    __half* A = nullptr;       // [n_pad][Kp]
    __half* B = nullptr;       // [k_pad][Kp]
    float* X2 = nullptr;       // [n_pad] |x~|^2
    float* XL = nullptr;       // [n_pad] upper bound of |x~ - hi(x~)| (terms < 3: what a hi-only frame operand drops)
    float* mu = nullptr;       // [d]
    ScreenParams* params = nullptr;
    uint32_t* cand = nullptr;  // [n_pad][CAND_CAP]  chunk id | group mask << 28
    uint8_t* ncand = nullptr;  // [n_pad]  (255: overflow -> exact fallback)
    uint32_t* fb_list = nullptr;  // [n_pad] frames the verify kernels hand to the fallback kernel
    CUtensorMap tmA, tmB, tmBh;  // tmBh: center operand in 128-row boxes (one CTA's half of a center tile, cluster mode)
    CUtensorMap tmBg;            // center operand in single-row boxes: the source of tile::gather4 loads (listed screen)
    int k_rows = 0;              // rows of B: k_pad + 8, the rows from k on carry a -inf bias (never candidates); row k_pad
                                 // is the padding id of the per-tile center lists
    int64_t prepared_n = -1;
};

// ---- driver entry point for tensor-map encoding (no -lcuda link: resolved at run time) ------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}
static int make_tmap(CUtensorMap* tm, void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    // (box_rows = 1: the descriptor of tile::gather4 loads -- one row of the gathered dimension, BLOCK_K contiguous columns)
    PFN_encodeTiled enc = get_encode();
    if (!enc) return set_error(B2K_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(B2K_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return B2K_OK;
}

// ---- prep kernels ---------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t cdiv_dev(int64_t a, int64_t b) { return (a + b - 1) / b; }
// mu[dim] = mean of the centers (fp64, fixed tree order): one block per dimension
__global__ void __launch_bounds__(256) screen_mu_kernel(const float* __restrict__ C, int k, int d, float* __restrict__ mu) {
    __shared__ double part[256];
    const int dim = blockIdx.x;
    double s = 0;
    for (int j = threadIdx.x; j < k; j += 256) s += (double)C[(int64_t)j * d + dim];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) mu[dim] = (float)(part[0] / k);
}

// sum_e ((row[e] - mu[e]) * scale)^2 with LPR lanes per row (coalesced for wide rows); any summation order:
// these sums only pick the scale and feed the margin, which carries its own slack for their rounding.
template <int LPR>
__device__ __forceinline__ float row_sqnorm(const float* __restrict__ row, const float* __restrict__ mu, int d,
                                            float scale, int sub) {
    float s = 0.f;
    for (int e = sub; e < d; e += LPR) { const float t = (__ldg(row + e) - __ldg(mu + e)) * scale; s += t * t; }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}

// max_i |row_i - mu|^2 -> *out (float bits, atomicMax)
template <int LPR>
__global__ void __launch_bounds__(256) screen_maxnorm_kernel(const float* __restrict__ X, int64_t n, int d,
                                                             const float* __restrict__ mu, float* out) {
    float m = 0.f;
    const int sub = threadIdx.x % LPR;
    const int64_t rows_per_pass = (int64_t)gridDim.x * (256 / LPR);
    const int64_t n_round = cdiv_dev(n, rows_per_pass) * rows_per_pass;  // keep whole warps in the shuffles
    for (int64_t i = (int64_t)blockIdx.x * (256 / LPR) + threadIdx.x / LPR; i < n_round; i += rows_per_pass) {
        const float s = row_sqnorm<LPR>(X + (i < n ? i : 0) * d, mu, d, 1.f, sub);
        if (i < n) m = fmaxf(m, s);  // NaN rows are ignored here and flagged in the operand kernel
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax((int*)out, __float_as_int(m));
}

// fp16 image of v with everything below the smallest NORMAL fp16 flushed to zero
__device__ __forceinline__ float h16_ftz(float v) {
    const float f = __half2float(__float2half_rn(v));
    return fabsf(f) < 6.103515625e-5f ? 0.f : f;
}
static constexpr float SCALE_TARGET = 2048.f;  // max |x~|, |c~| after scaling (power of two keeps hi/lo splits exact)
static constexpr float CMAX_LIMIT = 4096.f;    // centers may drift this far before the operands are declared unusable

__global__ void screen_sigma_kernel(ScreenParams* p) {
    const float mx = sqrtf(fmaxf(p->xmax2_raw, p->cmax2_raw));
    float sigma = 1.f;
    int valid = 1;
    if (!(mx < 3.0e38f)) valid = 0;
    else if (mx > 0.f) {
        // largest power of two with mx*sigma <= SCALE_TARGET
        int e;
        frexpf(SCALE_TARGET / fmaxf(mx, 1e-30f), &e);  // target/mx = f*2^e, f in [0.5,1)  -> 2^(e-1) <= target/mx
        e -= 1;
        if (e > 100) e = 100;
        if (e < -100) { e = -100; valid = 0; }
        sigma = ldexpf(1.f, e);
    }
    p->sigma = sigma;
    p->valid = valid;
}

// A' rows (layout in the file header); rows >= n are zero.  One thread owns one 16-byte piece (8 fp16) of
// the row and keeps that piece for every row it visits, so the column -> (segment, dimension) decode and the
// mu values are loop invariant; stores are fully coalesced, the fp32 frame values come through L1.
__global__ void __launch_bounds__(256) screen_frames_kernel(const float* __restrict__ X, int64_t n, int64_t n_pad,
                                                            int d, int terms, int Kp,
                                                            const float* __restrict__ mu,
                                                            const ScreenParams* __restrict__ prm,
                                                            uint4* __restrict__ A) {
    const int pieces = Kp >> 3;
    const int rows_per_pass = 256 / pieces;  // pieces <= 256 (Kp <= 2048) is checked by the plan
    const int piece = threadIdx.x % pieces, rloc = threadIdx.x / pieces;
    if (rloc >= rows_per_pass) return;
    const float sigma = prm->sigma;
    const int ones0 = terms * d;
    int seg[8], dim[8];
    float muv[8], cst[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int c = piece * 8 + q;
        seg[q] = -1;
        dim[q] = 0;
        muv[q] = 0.f;
        cst[q] = 0.f;
        if (c < ones0) {
            const int ps = (c >= 2 * d) ? 2 : (c >= d ? 1 : 0);  // physical segment; the hi.hi products come LAST
            seg[q] = (terms == 3) ? (ps == 2 ? 0 : ps + 1) : (terms == 2 ? (ps == 0 ? 2 : 0) : 0);
            dim[q] = c - ps * d;
            muv[q] = __ldg(mu + dim[q]);
        } else if (c < ones0 + 3) {
            cst[q] = (c == ones0) ? 256.f : (c == ones0 + 1 ? 0.125f : 6.103515625e-5f);
        }
    }
    for (int64_t i = (int64_t)blockIdx.x * rows_per_pass + rloc; i < n_pad; i += (int64_t)gridDim.x * rows_per_pass) {
        float v[8];
        const bool live = i < n;
        const float* xrow = X + i * d;
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = (live && seg[q] >= 0) ? __ldg(xrow + dim[q]) : 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            float out = live ? cst[q] : 0.f;
            if (seg[q] >= 0) {
                const float xt = __fmul_rn(__fsub_rn(v[q], muv[q]), sigma);
                const float hi = h16_ftz(xt);
                out = seg[q] == 0 ? hi : (seg[q] == 1 ? h16_ftz(hi * 0.03125f) : h16_ftz(__fsub_rn(xt, hi) * 32.f));
                if (!live) out = 0.f;
            }
            v[q] = out;
        }
        __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
        __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
        uint4 o;
        o.x = *reinterpret_cast<uint32_t*>(&h0);
        o.y = *reinterpret_cast<uint32_t*>(&h1);
        o.z = *reinterpret_cast<uint32_t*>(&h2);
        o.w = *reinterpret_cast<uint32_t*>(&h3);
        A[i * pieces + piece] = o;
    }
}

// Same operand, built per INPUT element: a tile of R frame rows is read coalesced, every value is centred, scaled and
// split ONCE (the kernel above recomputes it for each of its three output columns and decodes the column layout per
// element: ncu showed 41 instructions per output element, ALU pipe 67 %, DRAM 22 %), the fp16 results go to a
// shared-memory image of the tile's A' rows and leave with coalesced 16-byte stores.  Identical arithmetic.
__global__ void __launch_bounds__(256) screen_frames_tile_kernel(const float* __restrict__ X, int64_t n, int64_t n_pad,
                                                                 int d, int terms, int Kp, int R,
                                                                 const float* __restrict__ mu,
                                                                 const ScreenParams* __restrict__ prm,
                                                                 uint4* __restrict__ A) {
    extern __shared__ __align__(16) __half atile[];  // [R][Kp]
    const float sigma = prm->sigma;
    const int ones0 = terms * d;
    const int64_t n_tiles = (n_pad + R - 1) / R;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * R;
        const int rows = (int)min((int64_t)R, n_pad - row0);
        __syncthreads();  // previous tile copied out
        // zero image (padding columns, rows beyond n), then the data and constant columns of the live rows
        {
            uint4* z = reinterpret_cast<uint4*>(atile);
            const int zp = rows * (Kp >> 3);
            for (int t = threadIdx.x; t < zp; t += 256) z[t] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncthreads();
        const int live_rows = (int)max((int64_t)0, min((int64_t)rows, n - row0));
        const int total = live_rows * d;
        const float* src = X + row0 * d;
        // (r, e) of element t = r*d + e, advanced incrementally: no integer division per element
        int r = threadIdx.x / d, e = threadIdx.x - r * d;
        const int dr = 256 / d, de = 256 - dr * d;
        for (int t = threadIdx.x; t < total; t += 256) {
            __half* out = atile + (size_t)r * Kp;
            const float xt = __fmul_rn(__fsub_rn(__ldg(src + t), __ldg(mu + e)), sigma);
            const float hi = h16_ftz(xt);
            if (terms == 3) {
                out[e] = __float2half_rn(h16_ftz(hi * 0.03125f));
                out[d + e] = __float2half_rn(h16_ftz(__fsub_rn(xt, hi) * 32.f));
                out[2 * d + e] = __float2half_rn(hi);
            } else if (terms == 2) {
                out[e] = __float2half_rn(h16_ftz(__fsub_rn(xt, hi) * 32.f));
                out[d + e] = __float2half_rn(hi);
            } else {
                out[e] = __float2half_rn(hi);
            }
            r += dr;
            e += de;
            if (e >= d) { e -= d; ++r; }
        }
        for (int rr = threadIdx.x; rr < live_rows; rr += 256) {
            __half* out = atile + (size_t)rr * Kp + ones0;
            out[0] = __float2half_rn(256.f);
            out[1] = __float2half_rn(0.125f);
            out[2] = __float2half_rn(6.103515625e-5f);
        }
        __syncthreads();
        const int pieces = rows * (Kp >> 3);
        const uint4* st = reinterpret_cast<const uint4*>(atile);
        uint4* dst = A + row0 * (Kp >> 3);
        for (int t = threadIdx.x; t < pieces; t += 256) dst[t] = st[t];
    }
}

template <int LPR>
__global__ void __launch_bounds__(256) screen_x2_kernel(const float* __restrict__ X, int64_t n, int64_t n_pad, int d,
                                                        const float* __restrict__ mu,
                                                        const ScreenParams* __restrict__ prm,
                                                        float* __restrict__ X2, float* __restrict__ XL) {
    const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) / LPR;  // n_pad is a multiple of 128: whole warps
    if (i >= n_pad) return;
    const float* row = X + (i < n ? i : 0) * d;
    const int sub = threadIdx.x % LPR;
    float s = row_sqnorm<LPR>(row, mu, d, prm->sigma, sub);
    if (i >= n) s = 0.f;
    else if (!(s < 3.0e38f)) s = __int_as_float(0x7f800000);  // NaN/inf frame -> flagged by the epilogue
    if (sub == 0) X2[i] = s;
    if (XL) {
        // |x~ - hi(x~)|: the part of the frame a hi-only operand does not carry (flushed values included), rounded up
        const float sigma = prm->sigma;
        float l = 0.f;
        for (int e = sub; e < d; e += LPR) {
            const float xt = __fmul_rn(__fsub_rn(__ldg(row + e), __ldg(mu + e)), sigma);
            const float t = __fsub_rn(xt, h16_ftz(xt));
            l += t * t;
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
        if (sub == 0) XL[i] = (i < n && l < 3.0e38f) ? sqrtf(l * (1.f + (d + 4) * 1.2e-7f)) * 1.000001f : 0.f;
    }
}

// B' rows (one thread per center): [c_hi | (c_lo | c_hi) | -b1 -b2 -b3 | 0..]; rows >= k: bias -inf
__global__ void __launch_bounds__(128) screen_centers_kernel(const float* __restrict__ C, int k, int k_pad, int d,
                                                             int terms, int Kp, const float* __restrict__ mu,
                                                             ScreenParams* prm, __half* __restrict__ B) {
    const int j = blockIdx.x * 128 + threadIdx.x;
    if (j >= k_pad) return;
    __half* row = B + (int64_t)j * Kp;
    const int ones0 = terms * d;
    if (j >= k) {
        for (int c = 0; c < Kp; ++c) row[c] = __float2half_rn(0.f);
        row[ones0] = __ushort_as_half((unsigned short)0xFC00);  // -inf: never a candidate
        return;
    }
    const float sigma = prm->sigma;
    float s = 0.f, sl = 0.f;
    for (int e = 0; e < d; ++e) {
        const float ct = __fmul_rn(__fsub_rn(C[(int64_t)j * d + e], mu[e]), sigma);
        s += ct * ct;
        const float hi = h16_ftz(ct);
        const float lo = __fsub_rn(ct, hi);
        sl += lo * lo;
        if (terms == 3) {
            row[e] = __float2half_rn(h16_ftz(lo * 32.f));
            row[d + e] = __float2half_rn(h16_ftz(hi * 0.03125f));
            row[2 * d + e] = __float2half_rn(hi);
        } else if (terms == 2) {
            row[e] = __float2half_rn(h16_ftz(hi * 0.03125f));
            row[d + e] = __float2half_rn(hi);
        } else {
            row[e] = __float2half_rn(hi);
        }
    }
    // b = b1*2^8 + b2*2^-3 + b3*2^-14 (+ residual <= 2^-28 + 2^-33 b); every step below is exact in fp32
    const float b = 0.5f * s;
    const float b1 = h16_ftz(b * 0.00390625f);
    const float r1 = __fsub_rn(b, b1 * 256.f);
    const float b2 = h16_ftz(r1 * 8.f);
    const float r2 = __fsub_rn(r1, b2 * 0.125f);
    const float b3 = h16_ftz(r2 * 16384.f);
    row[ones0] = __float2half_rn(-b1);
    row[ones0 + 1] = __float2half_rn(-b2);
    row[ones0 + 2] = __float2half_rn(-b3);
    for (int c = ones0 + 3; c < Kp; ++c) row[c] = __float2half_rn(0.f);
    atomicMax((int*)&prm->cmax2_now, __float_as_int(s));
    atomicMax((int*)&prm->cl2_now, __float_as_int(sl));
}

__global__ void screen_finish_centers_kernel(ScreenParams* p, int d) {
    // upper bound of max |c~| (the fp32 evaluation above is within (d+2) ulp)
    const float c2 = p->cmax2_now * (1.f + (d + 4) * 1.2e-7f);
    p->cmax = sqrtf(c2) * 1.000001f;
    p->cl = sqrtf(p->cl2_now * (1.f + (d + 4) * 1.2e-7f)) * 1.000001f;
    if (!(p->cmax <= CMAX_LIMIT)) p->valid = 0;  // centers moved outside the scaled range
}

// ---- PTX wrappers -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
// one non-blocking probe of a phase
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// ---- 2-CTA cluster helpers (streaming mode with shared center tiles) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA tile load delivered to the same shared-memory offset (and mbarrier) of every CTA in `mask`
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%4, %5}], [%2], %3;" ::"r"(smem_u32(dst)),
        "l"(tm), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1)
        : "memory");
}
// commit that arrives on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
// ---- CTA-pair helpers (cta_group::2: one M=256 MMA over the two SMs of a pair, each SM holding its own 128 frame rows
// and HALF of the center tile; the leader CTA -- rank 0 -- issues every MMA and owns the full / accumulator-empty barriers)
// TMA tile load into MY shared memory whose completion is signalled on the LEADER's barrier at the same offset
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
        "l"(tm), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}
// arrive on the barrier at this offset in the leader CTA's shared memory
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {  // arrives on the barrier at this offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, SWIZZLE_128B operand tile whose 8-row groups are 1024 bytes apart
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t desc = 0;
    desc |= (uint64_t)((saddr & 0x3FFFF) >> 4);   // start address
    desc |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major)
    desc |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset: 8 rows * 128 B
    desc |= (uint64_t)1 << 46;                    // descriptor version (sm_100)
    desc |= (uint64_t)2 << 61;                    // SWIZZLE_128B
    return desc;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// the registers of an earlier tcgen05.ld are defined only after the wait: tie every later use to this point
__device__ __forceinline__ void tmem_ld_fence(float (&v)[32]) {
#pragma unroll
    for (int e = 0; e < 32; e += 8)
        asm volatile("" : "+f"(v[e]), "+f"(v[e + 1]), "+f"(v[e + 2]), "+f"(v[e + 3]), "+f"(v[e + 4]), "+f"(v[e + 5]),
                          "+f"(v[e + 6]), "+f"(v[e + 7])::"memory");
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- margin ----------------------------------------------------------------------------------------------------
struct Margin {
    float a, r, x2;
    // xl, CL: upper bounds of |x~ - hi(x~)| (this frame) and max_j |c~_j - hi(c~_j)| -- the parts a hi-only operand does
    // not carry, MEASURED by the operand builders (flushed values included), so the Cauchy-Schwarz bounds
    //   terms=1:  |x.c - x_hi.c_hi| <= |x_lo| |c| + |x_hi| |c_lo|      terms=2:  |x.c - x.c_hi| <= |x| |c_lo|
    // use the data's actual rounding residues (~2^-12.3 of the norms) instead of the worst case 2^-11 per side.
    __device__ __forceinline__ void init(float x2_, float xl, float C, float CL, int d, int nk16, int terms) {
        const float u = 5.9604645e-8f;
        const float gam = (0.25f * d + 9.f) * u * 1.01f;
        const float rho = 2.f * gam + 4.9e-7f;
        const float X = sqrtf(x2_) * (1.f + 2.f * gam);
        const float R = X + C;
        const float d1 = 2.1f * u * R * R;
        // rounding of the scaled lo segments (terms=3: x_lo, c_lo and the dropped lo.lo; terms=2: x_lo)
        const float erep = (terms == 3) ? 3.01f * 2.3841858e-7f : (terms == 2 ? 1.01f * 2.3841858e-7f : 0.f);
        const float edrop = (terms == 3) ? 0.f : (terms == 2 ? 1.001f * X * CL : 1.001f * (xl * C + 1.0005f * X * CL));
        // K=16 steps whose addends/partial sums are of full magnitude: from the step that holds the first
        // hi.hi column (2d for terms=3, d for terms=2, 0 for terms=1) to the end; the earlier ones see sums <= 2^-10 X C
        const int nk_lo = (terms == 3) ? (2 * d) / 16 : (terms == 2 ? d / 16 : 0);
        const float eacc = ((float)((nk16 - nk_lo) * 17) + (float)(nk_lo * 17) * 9.8e-4f) * 1.1920929e-7f;
        // values flushed to zero by the operand builder: per element <= (2^-19+2^-20)(|x_e|+|c_e|) with the
        // scaled lo segments (terms=3, 2); with the hi segment alone (terms=1) the flushed values are part of xl / CL
        const float eflush = (terms == 1) ? 0.f : 2.9e-6f;
        const float d2 = erep * X * C + edrop + eflush * sqrtf((float)d) * R + eacc * (1.01f * X * C + 0.5f * C * C) +
                         (gam + 1.2e-10f) * 0.5f * C * C + 4e-9f;
        a = (2.f * d2 + d1) * 1.01f;
        r = 0.5f * rho * 1.01f;
        x2 = x2_;
    }
    __device__ __forceinline__ float threshold(float m) const {
        return m - (a + r * (fmaxf(x2 - 2.f * m, 0.f) + a));
    }
};

struct GemmArgs {
    int64_t n;          // valid frames
    int n_tiles;        // frame tiles
    int n_ntiles;       // center tiles (k_pad / 256)
    int n_kblocks;      // Kp / 64
    int nk16;           // K=16 MMA steps that carry data
    int d, terms;
    int resident;       // 1: the whole center operand B' stays in shared memory, only frame tiles stream
                        // 2: the FRAME tile (all k-blocks) stays while its center tiles stream: A is read once per
                        //    frame tile instead of once per center tile (L2->SM traffic, the limiter of mid-size rows)
    int n_stages;       // pipeline stages
    int cluster2;       // launched as clusters of 2 CTAs (streaming mode).  1: each CTA fetches half of every center
                        //    k-block and multicasts it to both (halves the center traffic out of L2).  2: CTA-pair
                        //    MMAs (cta_group::2, M=256): each SM keeps only ITS half of the center k-block, so an SM
                        //    ingests and reads 2/3 of the shared-memory bytes per MMA
    int stage_bytes;    // resident: n_kblocks*A_BYTES (a frame tile, full K); streaming: A_BYTES+B_BYTES (one k-block)
    int bres_bytes;     // resident: bytes of B' in shared memory
    const float* X2;
    const float* XL;    // terms < 3
    const ScreenParams* prm;
    uint32_t* cand;
    uint8_t* ncand;
};

static constexpr int MAX_STAGES = 8;
static constexpr int MAX_A_KBLOCKS = 8;
static constexpr int EPI_WARPS = 8;                    // 2 column halves x 4 TMEM lane quarters
static constexpr int GEMM_THREADS = 64 + EPI_WARPS * 32;  // warp0 TMA, warp1 MMA + TMEM alloc, warps 2-9 epilogue
static constexpr int HALF_CHUNKS = TILE_N / CHUNK / 2;  // chunks of one accumulator stage handled by one epilogue warp

// small shared-memory block behind the operand tiles
struct GemmSmemTail {
    uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], tfull_bar[2], tempty_bar[2], bfull_bar;
    uint64_t afull_bar[MAX_A_KBLOCKS], aempty_bar[MAX_A_KBLOCKS];  // resident-A mode: one pair per k-block of the frame tile
    uint32_t tmem_slot, pad[3];
    uint32_t list_id[2][LIST_CAP][TILE_M];
    float list_v[2][LIST_CAP][TILE_M];
    float half_m[2][TILE_M];
    uint32_t out_n[2][TILE_M];              // kept entries of each half (CAND_CAP+1: too many)
    uint32_t out_id[2][CAND_CAP][TILE_M];
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// running candidate bookkeeping of one epilogue thread (one frame, one column half)
struct RowScan {
    float m, thr;
    int cnt;
    bool overflow;
    // TRACK (listed screen): the column that holds the running maximum and the second largest score of ITS chunk.  If at
    // the end only that chunk is a candidate and its second score is below the final threshold, exactly one center can
    // be the reference's argmin: the frame is DECIDED by the screen and needs no exact evaluation (NCAND_DECIDED).
    float best_v2;
    uint32_t best_pos;
    __device__ __forceinline__ void init() {
        m = __int_as_float(0xff800000);
        thr = m;
        cnt = 0;
        overflow = false;
        best_v2 = m;
        best_pos = 0;
    }
};
static constexpr int NCAND_DECIDED = 254;  // ncand value: cand[0] is the list position of the only possible center

template <int CG, bool TRACK = false>
__device__ __forceinline__ void scan_chunk(const float (&v)[32], uint32_t chunk_id, RowScan& rs, const Margin& mg,
                                           uint32_t* lid, float* lv /* this thread's column of the list arrays */) {
    constexpr int NG = CHUNK / CG;
    float gm[NG];
#pragma unroll
    for (int q = 0; q < NG; ++q) {
        const float* w = v + q * CG;
        if (CG == 8) gm[q] = fmaxf(fmaxf(fmaxf(fmaxf(w[0], w[1]), w[2]), fmaxf(fmaxf(w[3], w[4]), w[5])), fmaxf(w[6], w[7]));
        else if (CG == 4) gm[q] = fmaxf(fmaxf(fmaxf(w[0], w[1]), w[2]), w[3]);
        else gm[q] = fmaxf(w[0], w[1]);
    }
    float t3[(NG + 2) / 3 + 1];
    int nt = 0;
#pragma unroll
    for (int q = 0; q + 2 < NG; q += 3) t3[nt++] = fmaxf(fmaxf(gm[q], gm[q + 1]), gm[q + 2]);
    float cm = (NG % 3 == 1) ? gm[NG - 1] : fmaxf(gm[NG - 2], gm[NG - 1]);
#pragma unroll
    for (int q = 0; q < nt; ++q) cm = fmaxf(cm, t3[q]);
    if (cm > rs.m) {
        rs.m = cm;
        rs.thr = mg.threshold(cm);
        if (TRACK) {  // a few times per frame: argmax column and runner-up of the chunk that now holds the maximum
            float m1 = v[0], m2 = __int_as_float(0xff800000);
            uint32_t c1 = 0;
#pragma unroll
            for (int e = 1; e < CHUNK; ++e) {
                const bool gt = v[e] > m1;
                m2 = gt ? m1 : fmaxf(m2, v[e]);  // an equal score counts as a second candidate
                c1 = gt ? (uint32_t)e : c1;
                m1 = gt ? v[e] : m1;
            }
            rs.best_v2 = m2;
            rs.best_pos = chunk_id * CHUNK + c1;
        }
    }
    if (cm >= rs.thr) {
        // groups below the CURRENT threshold can never pass the final (higher) one
        uint32_t mask = 0;
#pragma unroll
        for (int q = 0; q < NG; ++q) mask |= (gm[q] >= rs.thr ? 1u : 0u) << q;
        if (rs.cnt == LIST_CAP) {  // compact: drop entries the risen threshold has already excluded
            int w = 0;
            for (int t = 0; t < LIST_CAP; ++t) {
                const float tv = lv[t * TILE_M];
                if (tv >= rs.thr) { lv[w * TILE_M] = tv; lid[w * TILE_M] = lid[t * TILE_M]; ++w; }
            }
            rs.cnt = w;
        }
        if (rs.cnt < LIST_CAP) {
            lid[rs.cnt * TILE_M] = chunk_id | (mask << cand_id_bits(CG));
            lv[rs.cnt * TILE_M] = cm;
            ++rs.cnt;
        } else {
            rs.overflow = true;
        }
    }
}

// ---- the screen kernel ---------------------------------------------------------------------------------------------
// EXP: the experimental operand modes are compiled into separate instantiations so that the production paths (EXP=0:
// resident center operand / plain streaming) carry none of their branches.  EXP=1: resident frame tile, 2-CTA multicast;
// EXP=2: CTA-pair MMAs (a kernel containing cta_group::2 instructions can only be launched as clusters).
template <int CG, int EXP>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
screen_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmBh /* center operand, 128-row boxes */, GemmArgs g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    if (!g.prm->valid) return;  // operands unusable: the exact tile kernel takes the whole call (screen_finish_assign)
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* bres = smem;                       // resident B' (bres_bytes, multiple of 1024)
    uint8_t* tiles = smem + g.bres_bytes;       // n_stages * stage_bytes
    GemmSmemTail* T = reinterpret_cast<GemmSmemTail*>(tiles + (size_t)g.n_stages * g.stage_bytes);

    // only EXP=1 is ever launched with the resident-frame-tile mode (2): elsewhere `resident` is a plain flag
    const int resident = EXP == 1 ? g.resident : (g.resident != 0 ? 1 : 0);
    const bool cluster2 = EXP != 0 && g.cluster2 != 0;
    const bool pair = EXP == 2 && g.cluster2 == 2;  // cta_group::2 MMAs
    const bool mcast = EXP == 1 && cluster2;        // multicast center tiles, cta_group::1 MMAs
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // frame tiles of this CTA: tile = t_first, t_first + t_step, ... (< t_count); in cluster mode the two CTAs of a
    // pair walk tiles 2*tt and 2*tt+1 in lockstep (a tile index beyond n_tiles is a dummy: zero operand, no output)
    const uint32_t crank = cluster2 ? cluster_ctarank() : 0u;
    // (evaluated at each use: hoisting blockIdx.x into one register shared by all warp roles costs the producer and
    // MMA-issue threads their uniform-datapath code)
#define B2K_T_FIRST (cluster2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x)
#define B2K_T_STEP (cluster2 ? (int)(gridDim.x >> 1) : (int)gridDim.x)
#define B2K_T_COUNT (cluster2 ? (g.n_tiles + 1) >> 1 : g.n_tiles)

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int s = 0; s < g.n_stages; ++s) {
            mbar_init(&T->full_bar[s], 1);
            mbar_init(&T->empty_bar[s], mcast ? 2 : 1);  // multicast mode: the MMA warps of both CTAs release a slot
        }
        // pair mode: the leader's MMA thread waits for the epilogue warps of BOTH CTAs
        for (int s = 0; s < 2; ++s) { mbar_init(&T->tfull_bar[s], 1); mbar_init(&T->tempty_bar[s], pair ? 2 * EPI_WARPS : EPI_WARPS); }
        mbar_init(&T->bfull_bar, 1);
        if constexpr (EXP == 1)
            for (int kb = 0; kb < MAX_A_KBLOCKS; ++kb) { mbar_init(&T->afull_bar[kb], 1); mbar_init(&T->aempty_bar[kb], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (EXP == 2 && pair) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&T->tmem_slot)),
                         "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&T->tmem_slot)),
                         "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (cluster2) cluster_sync_all();  // the peer's barriers are initialised before anything is sent to them
    tc_fence_after();
    const uint32_t tmem_base = T->tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            if (EXP == 1 && resident == 2) {
                // resident-A: two independent load streams polled by this one thread.  A: k-block kb of the frame
                // tile is refilled as soon as the last center tile of the previous frame tile has consumed it (while
                // that tile's remaining MMAs still run).  B: center k-blocks through the ring, running ahead into the
                // next frame tile -- neither stream ever waits for the other.
                int a_tile = blockIdx.x, a_kb = 0;
                uint32_t a_phase = 0;
                int b_tile = blockIdx.x, b_nt = 0, b_kb = 0;
                bool a_done = a_tile >= g.n_tiles, b_done = b_tile >= g.n_tiles;
                while (!a_done || !b_done) {
                    if (!a_done && mbar_try(&T->aempty_bar[a_kb], a_phase ^ 1)) {
                        mbar_expect_tx(&T->afull_bar[a_kb], A_BYTES);
                        tma_load_2d(bres + (size_t)a_kb * A_BYTES, &tmA, &T->afull_bar[a_kb], a_kb * BLOCK_K, a_tile * TILE_M);
                        if (++a_kb == g.n_kblocks) {
                            a_kb = 0;
                            a_phase ^= 1;
                            a_tile += gridDim.x;
                            a_done = a_tile >= g.n_tiles;
                        }
                    }
                    if (!b_done && mbar_try(&T->empty_bar[stage], phase ^ 1)) {
                        mbar_expect_tx(&T->full_bar[stage], B_BYTES);
                        tma_load_2d(tiles + (size_t)stage * g.stage_bytes, &tmB, &T->full_bar[stage], b_kb * BLOCK_K,
                                    b_nt * TILE_N);
                        if (++stage == g.n_stages) { stage = 0; phase ^= 1; }
                        if (++b_kb == g.n_kblocks) {
                            b_kb = 0;
                            if (++b_nt == g.n_ntiles) {
                                b_nt = 0;
                                b_tile += gridDim.x;
                                b_done = b_tile >= g.n_tiles;
                            }
                        }
                    }
                }
            } else if (resident) {
                mbar_expect_tx(&T->bfull_bar, (uint32_t)g.bres_bytes);
                for (int nt = 0; nt < g.n_ntiles; ++nt)
                    for (int kb = 0; kb < g.n_kblocks; ++kb)
                        tma_load_2d(bres + (size_t)(nt * g.n_kblocks + kb) * B_BYTES, &tmB, &T->bfull_bar, kb * BLOCK_K,
                                    nt * TILE_N);
                for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
                    mbar_wait(&T->empty_bar[stage], phase ^ 1);
                    uint8_t* sa = tiles + (size_t)stage * g.stage_bytes;
                    mbar_expect_tx(&T->full_bar[stage], (uint32_t)g.stage_bytes);
                    for (int kb = 0; kb < g.n_kblocks; ++kb)
                        tma_load_2d(sa + (size_t)kb * A_BYTES, &tmA, &T->full_bar[stage], kb * BLOCK_K, tile * TILE_M);
                    if (++stage == g.n_stages) { stage = 0; phase ^= 1; }
                }
            } else {
                for (int tt = B2K_T_FIRST; tt < B2K_T_COUNT; tt += B2K_T_STEP) {
                    const int tile = cluster2 ? 2 * tt + (int)crank : tt;
                    for (int nt = 0; nt < g.n_ntiles; ++nt) {
                        for (int kb = 0; kb < g.n_kblocks; ++kb) {
                            mbar_wait(&T->empty_bar[stage], phase ^ 1);
                            uint8_t* sa = tiles + (size_t)stage * g.stage_bytes;
                            if (EXP == 2 && pair) {
                                // my frame rows + my half of the center k-block into my slot; both CTAs' bytes are
                                // counted on the leader's barrier (its MMA thread is the only consumer)
                                if (crank == 0) mbar_expect_tx(&T->full_bar[stage], 2 * (A_BYTES + B_BYTES / 2));
                                tma_load_2d_pair(sa, &tmA, &T->full_bar[stage], kb * BLOCK_K, tile * TILE_M);
                                tma_load_2d_pair(sa + A_BYTES, &tmBh, &T->full_bar[stage], kb * BLOCK_K,
                                                 nt * TILE_N + (int)crank * (TILE_N / 2));
                                if (++stage == g.n_stages) { stage = 0; phase ^= 1; }
                                continue;
                            }
                            mbar_expect_tx(&T->full_bar[stage], STAGE_BYTES);
                            tma_load_2d(sa, &tmA, &T->full_bar[stage], kb * BLOCK_K, tile * TILE_M);
                            if (mcast)  // my half of the center k-block, delivered to both CTAs of the pair
                                tma_load_2d_mc(sa + A_BYTES + crank * (B_BYTES / 2), &tmBh, &T->full_bar[stage], kb * BLOCK_K,
                                               nt * TILE_N + (int)crank * (TILE_N / 2), (uint16_t)3);
                            else
                                tma_load_2d(sa + A_BYTES, &tmB, &T->full_bar[stage], kb * BLOCK_K, nt * TILE_N);
                            if (++stage == g.n_stages) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0 && !(EXP == 2 && pair && crank != 0)) {  // pair mode: the leader CTA issues for both
            // instruction descriptor: D=f32, A=B=f16, K-major both, N=256, M=128 (pair: M=256 over the two CTAs)
            const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(TILE_N >> 3) << 17) |
                                   ((uint32_t)((pair ? 2 * TILE_M : TILE_M) >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t it = 0;
            uint32_t tphase = 0;
            if (resident == 1) {
                mbar_wait(&T->bfull_bar, 0);
                tc_fence_after();
            }
            for (int tt = B2K_T_FIRST; tt < B2K_T_COUNT; tt += B2K_T_STEP) {
                if (resident == 1) {
                    mbar_wait(&T->full_bar[stage], phase);
                    tc_fence_after();
                }
                for (int nt = 0; nt < g.n_ntiles; ++nt, ++it) {
                    const uint32_t acc = it & 1u, accphase = (it >> 1) & 1u;
                    mbar_wait(&T->tempty_bar[acc], accphase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * TILE_N;
                    for (int kb = 0; kb < g.n_kblocks; ++kb) {
                        uint32_t sa, sb;
                        if (EXP == 1 && resident == 2) {
                            if (nt == 0) mbar_wait(&T->afull_bar[kb], tphase);
                            mbar_wait(&T->full_bar[stage], phase);
                            tc_fence_after();
                            sa = smem_u32(bres + (size_t)kb * A_BYTES);
                            sb = smem_u32(tiles + (size_t)stage * g.stage_bytes);
                        } else if (resident) {
                            sa = smem_u32(tiles + (size_t)stage * g.stage_bytes + (size_t)kb * A_BYTES);
                            sb = smem_u32(bres + (size_t)(nt * g.n_kblocks + kb) * B_BYTES);
                        } else {
                            mbar_wait(&T->full_bar[stage], phase);
                            tc_fence_after();
                            sa = smem_u32(tiles + (size_t)stage * g.stage_bytes);
                            sb = sa + A_BYTES;
                        }
                        const uint64_t adesc = make_smem_desc(sa);
                        const uint64_t bdesc = make_smem_desc(sb);
                        const int ksteps = min(BLOCK_K / 16, g.nk16 - kb * (BLOCK_K / 16));
                        if (EXP == 2 && pair) {
                            for (int ks = 0; ks < ksteps; ++ks)
                                tc_mma_f16_pair(d_tmem, adesc + (uint64_t)(ks * 2), bdesc + (uint64_t)(ks * 2), idesc,
                                                (kb | ks) != 0 ? 1u : 0u);
                            tc_commit_pair(&T->empty_bar[stage]);  // both CTAs' slots are free once these retire
                            if (++stage == g.n_stages) { stage = 0; phase ^= 1; }
                            continue;
                        }
                        for (int ks = 0; ks < ksteps; ++ks) {
                            // +32 bytes per K=16 step inside the 128-byte swizzle row (address field is >>4)
                            tc_mma_f16(d_tmem, adesc + (uint64_t)(ks * 2), bdesc + (uint64_t)(ks * 2), idesc,
                                       (kb | ks) != 0 ? 1u : 0u);
                        }
                        if (resident != 1) {
                            // smem slot free once these MMAs retire (cluster: tell both CTAs, either may refill it)
                            if (mcast) tc_commit_mc(&T->empty_bar[stage], (uint16_t)3);
                            else tc_commit(&T->empty_bar[stage]);
                            if (++stage == g.n_stages) { stage = 0; phase ^= 1; }
                        }
                        // resident-A: this k-block of the frame tile has met its last center tile
                        if (EXP == 1 && resident == 2 && nt == g.n_ntiles - 1) tc_commit(&T->aempty_bar[kb]);
                    }
                    if (EXP == 2 && pair) tc_commit_pair(&T->tfull_bar[acc]);  // each CTA's epilogue reads its own TMEM half
                    else tc_commit(&T->tfull_bar[acc]);             // accumulator stage complete
                }
                if (resident == 1) {
                    tc_commit(&T->empty_bar[stage]);  // frame tile consumed by every center tile
                    if (++stage == g.n_stages) { stage = 0; phase ^= 1; }
                }
                tphase ^= 1;
            }
        }
    } else {
        // ===== epilogue: 8 warps = 2 column halves x 4 TMEM lane quarters, one frame (TMEM lane) per thread =====
        const int q = warp & 3;              // TMEM lane quarter this warp may access
        const int h = (warp - 2) >> 2;       // column half of every accumulator stage
        const int row = q * 32 + lane;       // row inside the frame tile
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * (TILE_N / 2));
        uint32_t* lid = &T->list_id[h][0][row];
        float* lv = &T->list_v[h][0][row];
        uint32_t it = 0;
        const float C = g.prm->cmax, CL = g.prm->cl;
        const int valid_ops = g.prm->valid;
        for (int tt = B2K_T_FIRST; tt < B2K_T_COUNT; tt += B2K_T_STEP) {
            const int tile = cluster2 ? 2 * tt + (int)crank : tt;
            const int64_t grow = (int64_t)tile * TILE_M + row;
            const float x2 = (grow < g.n) ? g.X2[grow] : 0.f;
            const float xl = (g.terms != 3 && grow < g.n) ? g.XL[grow] : 0.f;
            Margin mg;
            mg.init(x2, xl, C, CL, g.d, g.nk16, g.terms);
            RowScan rs;
            rs.init();
            float va[32], vb[32];
            {
                const uint32_t acc = it & 1u, accphase = (it >> 1) & 1u;
                mbar_wait(&T->tfull_bar[acc], accphase);
                tc_fence_after();
                tmem_ld32(lane_addr + acc * TILE_N, va);
            }
            for (int nt = 0; nt < g.n_ntiles; ++nt, ++it) {
                const uint32_t acc = it & 1u;
                const uint32_t taddr = lane_addr + acc * TILE_N;
                const uint32_t cbase = (uint32_t)(nt * (TILE_N / CHUNK) + h * HALF_CHUNKS);
#pragma unroll
                for (int c = 0; c < HALF_CHUNKS; ++c) {
                    tmem_ld_wait();
                    if (c & 1) tmem_ld_fence(vb);
                    else tmem_ld_fence(va);
                    if (c + 1 < HALF_CHUNKS) {
                        if (c & 1) tmem_ld32(taddr + (c + 1) * CHUNK, va);
                        else tmem_ld32(taddr + (c + 1) * CHUNK, vb);
                    } else {
                        // every load of this accumulator stage has landed: hand it back to the MMA warp,
                        // then start on the next stage while the last chunk is being scanned
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (EXP == 2 && pair) mbar_arrive_leader(&T->tempty_bar[acc]);
                            else mbar_arrive(&T->tempty_bar[acc]);
                        }
                        if (nt + 1 < g.n_ntiles) {
                            const uint32_t nacc = (it + 1) & 1u, nphase = ((it + 1) >> 1) & 1u;
                            mbar_wait(&T->tfull_bar[nacc], nphase);
                            tc_fence_after();
                            tmem_ld32(lane_addr + nacc * TILE_N, va);  // HALF_CHUNKS is even: chunk 0 always lands in va
                        }
                    }
                    if (c & 1) scan_chunk<CG>(vb, cbase + c, rs, mg, lid, lv);
                    else scan_chunk<CG>(va, cbase + c, rs, mg, lid, lv);
                }
            }
            // ---- merge the two column halves of this frame ----
            T->half_m[h][row] = rs.m;
            named_bar_sync(1, EPI_WARPS * 32);
            const float m = fmaxf(T->half_m[0][row], T->half_m[1][row]);
            const float thr = mg.threshold(m);
            {
                int kept = 0;
                for (int t = 0; t < rs.cnt; ++t) {
                    if (lv[t * TILE_M] >= thr) {
                        if (kept < CAND_CAP) T->out_id[h][kept][row] = lid[t * TILE_M];
                        ++kept;
                    }
                }
                if (rs.overflow || kept > CAND_CAP) kept = CAND_CAP + 1;
                T->out_n[h][row] = (uint32_t)kept;
            }
            named_bar_sync(2, EPI_WARPS * 32);
            if (h == 0 && grow < g.n) {
                const int n0 = (int)T->out_n[0][row], n1 = (int)T->out_n[1][row];
                uint32_t ids[CAND_CAP] = {0, 0, 0, 0, 0, 0, 0, 0};
                int kept = n0 + n1;
                bool overflow = n0 > CAND_CAP || n1 > CAND_CAP || kept > CAND_CAP || kept == 0 || !(m > -3.0e38f) ||
                                !valid_ops || !(x2 < 3.0e38f);
                if (!overflow) {
                    // each half's list is ascending; the verify scan needs ONE ascending list (lowest index wins ties)
                    int a = 0, b = 0;
                    for (int w = 0; w < kept; ++w) {
                        const uint32_t ia = a < n0 ? T->out_id[0][a][row] : 0xffffffffu;
                        const uint32_t ib = b < n1 ? T->out_id[1][b][row] : 0xffffffffu;
                        if (a < n0 && (b >= n1 || (ia & cand_id_mask(CG)) < (ib & cand_id_mask(CG)))) { ids[w] = ia; ++a; }
                        else { ids[w] = ib; ++b; }
                    }
                }
                uint4* cout = reinterpret_cast<uint4*>(g.cand + grow * CAND_CAP);
                cout[0] = make_uint4(ids[0], ids[1], ids[2], ids[3]);
                if (kept > 4 && !overflow) cout[1] = make_uint4(ids[4], ids[5], ids[6], ids[7]);
                g.ncand[grow] = overflow ? (uint8_t)255 : (uint8_t)kept;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (cluster2) cluster_sync_all();  // no CTA leaves while its peer may still write to its shared memory / barriers
#undef B2K_T_FIRST
#undef B2K_T_STEP
#undef B2K_T_COUNT
    if (warp == 1) {
        tc_fence_after();
        if (EXP == 2 && pair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}


// ---- the screen kernel over per-tile center lists (prune.cu) --------------------------------------------------------
// Same roles as above in streaming mode, but frame tile t meets only the centers of ITS list: per pass of <= 256 list
// entries the producer thread fetches the frame k-block with one tiled TMA load, four GATHER warps copy the listed rows
// of the center operand (L2 resident) into the stage with 16-byte cp.async, writing the SWIZZLE_128B K-major layout a
// tiled TMA load would produce (chunk c of row r at c ^ (r & 7)) and publishing it to the async proxy
// (fence.proxy.async) before they arrive on the stage's barrier; the MMA thread issues M=128 x N=rows instructions, and
// two epilogue teams (four warps each, one accumulator stage each) take alternate tiles: one thread scans every chunk of
// its frame (list lengths are multiples of 32).  A candidate entry's chunk id counts 32-entry chunks of the LIST; the listed verify kernels map list
// positions back to center indices.  (gather = 1 fetches the rows with TMA tile::gather4 instead -- four rows per
// instruction: correct, but measured 1.43 ms against the cp.async gather at 1e7 x 10, k=1000, ~160 listed centers per
// tile: one gather4 costs the TMA unit ~130 cycles, 40 of them per tile are the whole kernel.)
struct ListArgs {
    const uint16_t* tlist;   // [n_units][lcap] center ids, ascending, padded with the id of a -inf row
    const uint32_t* tcount;  // [n_units] padded list lengths (multiples of 32, <= lcap)
    const __half* B;         // center operand [k_rows][Kp]
    int lcap, Kp;
    int ushift;              // list unit of frame tile t: t >> ushift
    int decide;              // 1: frames with a single possible center are settled by the screen (NCAND_DECIDED)
    int gather;              // 0: cp.async gather warps, 1: TMA tile::gather4
};
static constexpr int GATHER_WARPS = 4;
static constexpr int LISTED_THREADS = GEMM_THREADS + GATHER_WARPS * 32;

__device__ __forceinline__ void tma_gather4(void* dst, const CUtensorMap* tm, uint64_t* bar, int col, int r0, int r1, int r2,
                                            int r3) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, "
        "%7}], [%2];" ::"r"(smem_u32(dst)),
        "l"(tm), "r"(smem_u32(bar)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
        : "memory");
}

template <int CG, bool DECIDE>
__global__ void __launch_bounds__(LISTED_THREADS, 1)
screen_gemm_listed_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBg, GemmArgs g,
                          ListArgs la) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    if (!g.prm->valid) return;
    uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    GemmSmemTail* T = reinterpret_cast<GemmSmemTail*>(tiles + (size_t)g.n_stages * STAGE_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBg) : "memory");
        // a stage is full when the frame tile has landed (the producer's expect_tx arrival + TMA bytes) and every gather
        // warp has published its rows
        for (int s = 0; s < g.n_stages; ++s) {
            mbar_init(&T->full_bar[s], la.gather == 0 ? 1 + GATHER_WARPS : 1);
            mbar_init(&T->empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) { mbar_init(&T->tfull_bar[s], 1); mbar_init(&T->tempty_bar[s], EPI_WARPS / 2); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&T->tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = T->tmem_slot;

    if (warp >= GEMM_THREADS / 32) {
        // ===== gather warps: listed rows of B' -> stage, 16 bytes per cp.async, one pass behind in signalling so that two
        // stages' copies are in flight =====
        if (la.gather == 0) {
            const int gt = (int)threadIdx.x - GEMM_THREADS;  // 0 .. 127
            const int chunk = gt & 7;                        // 16-byte chunk of a 128-byte row piece
            const int r0 = (gt >> 3) * 16;                   // this thread's 16 consecutive list rows of a pass
            const size_t row_bytes = (size_t)la.Kp * 2;
            const uint8_t* Bb = reinterpret_cast<const uint8_t*>(la.B) + (size_t)(chunk << 4);
            int stage = 0, prev_stage = -1;
            uint32_t phase = 0;
            // the per-tile metadata (list length, this thread's 16 ids of the first pass) is requested one tile ahead:
            // a dependent global load per tile on this path would cost more than the tile's whole copy
            int tile = blockIdx.x;
            int cnt = tile < g.n_tiles ? (int)__ldg(la.tcount + (tile >> la.ushift)) : 0;
            uint4 ia = make_uint4(0u, 0u, 0u, 0u), ib = ia;
            if (tile < g.n_tiles && r0 < la.lcap) {  // (r0 + 16 <= lcap: both are multiples of 16)
                const uint4* src = reinterpret_cast<const uint4*>(la.tlist + (size_t)(tile >> la.ushift) * la.lcap + r0);
                ia = __ldg(src);
                ib = __ldg(src + 1);
            }
            for (; tile < g.n_tiles; tile += gridDim.x) {
                const int ntile = tile + (int)gridDim.x;
                int ncnt = 0;
                uint4 na = make_uint4(0u, 0u, 0u, 0u), nb = na;
                if (ntile < g.n_tiles) {
                    ncnt = (int)__ldg(la.tcount + (ntile >> la.ushift));
                    if (r0 < la.lcap) {
                        const uint4* src = reinterpret_cast<const uint4*>(la.tlist + (size_t)(ntile >> la.ushift) * la.lcap + r0);
                        na = __ldg(src);
                        nb = __ldg(src + 1);
                    }
                }
                const uint16_t* tl = la.tlist + (size_t)(tile >> la.ushift) * la.lcap;
                for (int p0 = 0; p0 < cnt; p0 += TILE_N) {
                    const int rows = min(TILE_N, cnt - p0);
                    if (p0 > 0 && p0 + r0 < la.lcap) {  // later passes of a long list: fetched here
                        const uint4* src = reinterpret_cast<const uint4*>(tl + p0 + r0);
                        ia = __ldg(src);
                        ib = __ldg(src + 1);
                    }
                    const uint32_t w[8] = {ia.x, ia.y, ia.z, ia.w, ib.x, ib.y, ib.z, ib.w};
                    for (int kb = 0; kb < g.n_kblocks; ++kb) {
                        if (lane == 0) mbar_wait(&T->empty_bar[stage], phase ^ 1);
                        __syncwarp();
                        uint8_t* sb = tiles + (size_t)stage * STAGE_BYTES + A_BYTES;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int r = r0 + i;
                            if (r < rows) {
                                const uint32_t j = (i & 1) ? (w[i >> 1] >> 16) : (w[i >> 1] & 0xffffu);
                                cp_async16(sb + (size_t)r * 128 + (size_t)((chunk ^ (r & 7)) << 4),
                                           Bb + (size_t)j * row_bytes + (size_t)kb * 128);
                            }
                        }
                        cp_async_commit();
                        if (prev_stage >= 0) {
                            cp_async_wait<1>();
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&T->full_bar[prev_stage]);
                        }
                        prev_stage = stage;
                        if (++stage == g.n_stages) { stage = 0; phase ^= 1; }
                    }
                }
                cnt = ncnt;
                ia = na;
                ib = nb;
            }
            if (prev_stage >= 0) {
                cp_async_wait<0>();
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&T->full_bar[prev_stage]);
            }
        }
    } else if (warp == 0 && la.gather == 0) {
        // ===== producer (cp.async gather mode): the frame k-blocks =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int tile = blockIdx.x;
            int cnt = tile < g.n_tiles ? (int)__ldg(la.tcount + (tile >> la.ushift)) : 0;
            for (; tile < g.n_tiles; tile += gridDim.x) {
                const int ntile = tile + (int)gridDim.x;
                const int ncnt = ntile < g.n_tiles ? (int)__ldg(la.tcount + (ntile >> la.ushift)) : 0;  // one tile ahead
                for (int p0 = 0; p0 < cnt; p0 += TILE_N) {
                    for (int kb = 0; kb < g.n_kblocks; ++kb) {
                        mbar_wait(&T->empty_bar[stage], phase ^ 1);
                        mbar_expect_tx(&T->full_bar[stage], (uint32_t)A_BYTES);
                        tma_load_2d(tiles + (size_t)stage * STAGE_BYTES, &tmA, &T->full_bar[stage], kb * BLOCK_K, tile * TILE_M);
                        if (++stage == g.n_stages) { stage = 0; phase ^= 1; }
                    }
                }
                cnt = ncnt;
            }
        }
    } else if (warp == 0) {
        // ===== producer (gather4 mode): the whole warp issues (lane l owns list rows 8l .. 8l+7 of a pass) =====
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
            const int cnt = (int)la.tcount[tile >> la.ushift];
            const uint16_t* tl = la.tlist + (size_t)(tile >> la.ushift) * la.lcap;
            for (int p0 = 0; p0 < cnt; p0 += TILE_N) {
                const int rows = min(TILE_N, cnt - p0);
                const bool mine = 8 * lane < rows;
                uint4 iv = make_uint4(0u, 0u, 0u, 0u);
                if (mine) iv = __ldg(reinterpret_cast<const uint4*>(tl + p0 + 8 * lane));
                for (int kb = 0; kb < g.n_kblocks; ++kb) {
                    uint8_t* sa = tiles + (size_t)stage * STAGE_BYTES;
                    if (lane == 0) {
                        mbar_wait(&T->empty_bar[stage], phase ^ 1);
                        mbar_expect_tx(&T->full_bar[stage], (uint32_t)(A_BYTES + rows * (BLOCK_K * 2)));
                        tma_load_2d(sa, &tmA, &T->full_bar[stage], kb * BLOCK_K, tile * TILE_M);
                    }
                    __syncwarp();
                    if (mine) {
                        uint8_t* sb = sa + A_BYTES + (size_t)(8 * lane) * (BLOCK_K * 2);
                        tma_gather4(sb, &tmBg, &T->full_bar[stage], kb * BLOCK_K, (int)(iv.x & 0xffffu), (int)(iv.x >> 16),
                                    (int)(iv.y & 0xffffu), (int)(iv.y >> 16));
                        tma_gather4(sb + 4 * (BLOCK_K * 2), &tmBg, &T->full_bar[stage], kb * BLOCK_K, (int)(iv.z & 0xffffu),
                                    (int)(iv.z >> 16), (int)(iv.w & 0xffffu), (int)(iv.w >> 16));
                    }
                    if (++stage == g.n_stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            uint32_t uses[2] = {0u, 0u};  // passes each accumulator stage has carried
            uint32_t seq = 0;             // tile sequence number of this CTA: stage (and epilogue team) seq & 1
            int tile = blockIdx.x;
            int cnt = tile < g.n_tiles ? (int)__ldg(la.tcount + (tile >> la.ushift)) : 0;
            int ncnt = 0;
            for (; tile < g.n_tiles; tile += gridDim.x, cnt = ncnt, ++seq) {
                const int ntile = tile + (int)gridDim.x;
                ncnt = ntile < g.n_tiles ? (int)__ldg(la.tcount + (ntile >> la.ushift)) : 0;  // one tile ahead
                const uint32_t acc = seq & 1u;
                for (int p0 = 0; p0 < cnt; p0 += TILE_N) {
                    const int rows = min(TILE_N, cnt - p0);
                    // instruction descriptor: D=f32, A=B=f16, K-major both, N=rows, M=128
                    const uint32_t idesc = (1u << 4) | ((uint32_t)(rows >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
                    const uint32_t accphase = (acc ? uses[1] : uses[0]) & 1u;
                    if (acc) ++uses[1]; else ++uses[0];
                    mbar_wait(&T->tempty_bar[acc], accphase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * TILE_N;
                    for (int kb = 0; kb < g.n_kblocks; ++kb) {
                        mbar_wait(&T->full_bar[stage], phase);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(tiles + (size_t)stage * STAGE_BYTES);
                        const uint64_t adesc = make_smem_desc(sa);
                        const uint64_t bdesc = make_smem_desc(sa + A_BYTES);
                        const int ksteps = min(BLOCK_K / 16, g.nk16 - kb * (BLOCK_K / 16));
                        for (int ks = 0; ks < ksteps; ++ks)
                            tc_mma_f16(d_tmem, adesc + (uint64_t)(ks * 2), bdesc + (uint64_t)(ks * 2), idesc,
                                       (kb | ks) != 0 ? 1u : 0u);
                        tc_commit(&T->empty_bar[stage]);
                        if (++stage == g.n_stages) { stage = 0; phase ^= 1; }
                    }
                    tc_commit(&T->tfull_bar[acc]);
                }
            }
        }
    } else {
        // ===== epilogue: two TEAMS of four warps (one per TMEM lane quarter).  Team t owns accumulator stage t and
        // every second tile of this CTA: one thread scans ALL chunks of its frame, so there are no column halves to
        // merge and no barrier between the eight warps; while one team finishes a tile (threshold, candidate entries,
        // stores) and waits for its stage to be refilled, the other one keeps the TMEM read port busy. =====
        const int q = warp & 3;
        const int team = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(team * TILE_N);
        uint32_t* lid = &T->list_id[team][0][row];
        float* lv = &T->list_v[team][0][row];
        const float C = g.prm->cmax, CL = g.prm->cl;
        const int valid_ops = g.prm->valid;
        uint32_t uses = 0;  // passes this team's accumulator stage has carried
        float va[32], vb[32];
        bool par = false;   // the buffer that holds (or receives) the current chunk: false = va
        const int tstep = 2 * (int)gridDim.x;
        int tile = blockIdx.x + team * (int)gridDim.x;
        // per-tile metadata one tile ahead (see the gather warps)
        int cnt = 0, ncnt = 0;
        float x2 = 0.f, xl = 0.f, nx2 = 0.f, nxl = 0.f;
        if (tile < g.n_tiles) {
            cnt = (int)__ldg(la.tcount + (tile >> la.ushift));
            const int64_t gr = (int64_t)tile * TILE_M + row;
            if (gr < g.n) { x2 = __ldg(g.X2 + gr); if (g.terms != 3) xl = __ldg(g.XL + gr); }
        }
        for (; tile < g.n_tiles; tile += tstep, cnt = ncnt, x2 = nx2, xl = nxl) {
            const int64_t grow = (int64_t)tile * TILE_M + row;
            const int next_tile = tile + tstep;
            ncnt = 0; nx2 = 0.f; nxl = 0.f;
            if (next_tile < g.n_tiles) {
                ncnt = (int)__ldg(la.tcount + (next_tile >> la.ushift));
                const int64_t gr = (int64_t)next_tile * TILE_M + row;
                if (gr < g.n) { nx2 = __ldg(g.X2 + gr); if (g.terms != 3) nxl = __ldg(g.XL + gr); }
            }
            Margin mg;
            mg.init(x2, xl, C, CL, g.d, g.nk16, g.terms);
            RowScan rs;
            rs.init();
            for (int p0 = 0; p0 < cnt; p0 += TILE_N, ++uses) {
                const int nch = min(TILE_N, cnt - p0) / CHUNK;   // >= 1
                const uint32_t cbase = (uint32_t)(p0 / CHUNK);
                mbar_wait(&T->tfull_bar[team], uses & 1u);
                tc_fence_after();
                if (par) tmem_ld32(taddr0, vb); else tmem_ld32(taddr0, va);
                for (int c = 0; c < nch; ++c) {
                    tmem_ld_wait();
                    if (par) tmem_ld_fence(vb); else tmem_ld_fence(va);
                    if (c + 1 < nch) {
                        if (par) tmem_ld32(taddr0 + (uint32_t)((c + 1) * CHUNK), va);
                        else tmem_ld32(taddr0 + (uint32_t)((c + 1) * CHUNK), vb);
                    } else {
                        // the last load of this pass has landed: the MMA thread may refill the stage
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&T->tempty_bar[team]);
                    }
                    if (par) scan_chunk<CG, DECIDE>(vb, cbase + (uint32_t)c, rs, mg, lid, lv);
                    else scan_chunk<CG, DECIDE>(va, cbase + (uint32_t)c, rs, mg, lid, lv);
                    par = !par;
                }
            }
            if (grow < g.n) {
                // entries still above the final threshold, in scan (= ascending chunk) order
                const float thr = rs.thr;
                uint32_t ids[CAND_CAP] = {0, 0, 0, 0, 0, 0, 0, 0};
                int kept = 0;
                for (int t = 0; t < rs.cnt; ++t) {
                    if (lv[t * TILE_M] >= thr) {
                        const uint32_t e = lid[t * TILE_M];
#pragma unroll
                        for (int w = 0; w < CAND_CAP; ++w)
                            if (w == kept) ids[w] = e;
                        ++kept;
                    }
                }
                const bool overflow = rs.overflow || kept > CAND_CAP || kept == 0 || !(rs.m > -3.0e38f) || !valid_ops ||
                                      !(x2 < 3.0e38f);
                // one candidate chunk (the one that holds the maximum) and no second score of it within the margin:
                // only one center can be the argmin
                const bool decided = DECIDE && !overflow && kept == 1 && rs.best_v2 < thr;
                if (decided) ids[0] = rs.best_pos;
                uint4* cout = reinterpret_cast<uint4*>(g.cand + grow * CAND_CAP);
                cout[0] = make_uint4(ids[0], ids[1], ids[2], ids[3]);
                if (kept > 4 && !overflow) cout[1] = make_uint4(ids[4], ids[5], ids[6], ids[7]);
                g.ncand[grow] = overflow ? (uint8_t)255 : (decided ? (uint8_t)NCAND_DECIDED : (uint8_t)kept);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ---- verify: exact reference-order evaluation of the surviving 8-center groups --------------------------------------
// Every distance is ONE thread's sequential 4-lane sum in the reference order (common.cuh), so the work is
// spread over (frame, center) pairs, never over the dimensions of one pair.

// a frame the screen could not bound (candidate list overflow, non-finite data): queue it for the
// CTA-per-frame exact scan below instead of stalling one lane group of a verify warp on k distances
__device__ __forceinline__ void fallback_push(ScreenParams* prm, uint32_t* fb_list, int64_t frame) {
    const unsigned int slot = atomicAdd(&prm->fb_count, 1u);
    fb_list[slot] = (uint32_t)frame;
}

// exact scan of every center for the queued frames: one CTA per frame, thread t takes centers t, t+256, ...
// (ascending per thread), then an order-free (sqrt(s), j) merge
__global__ void __launch_bounds__(256) screen_fallback_kernel(const float* __restrict__ X, int d,
                                                              const float* __restrict__ Cn, int k,
                                                              const uint32_t* __restrict__ fb_list,
                                                              const ScreenParams* __restrict__ prm,
                                                              int32_t* __restrict__ labels, float* __restrict__ mind,
                                                              int lloyd, unsigned int max_count) {
    extern __shared__ __align__(16) float fsm[];
    float* xs = fsm;                                   // [d]
    float* red_s = fsm + ((d + 3) & ~3);               // [256]
    int32_t* red_j = reinterpret_cast<int32_t*>(red_s + 256);
    if (!prm->valid) return;  // the whole call went to the exact tile kernel instead
    const unsigned int count = prm->fb_count;
    if (count >= max_count) return;  // a long queue: the register-tiled kernel takes it
    for (unsigned int b = blockIdx.x; b < count; b += gridDim.x) {
        const int64_t i = fb_list[b];
        __syncthreads();
        for (int e = threadIdx.x; e < d; e += 256) xs[e] = X[i * d + e];
        __syncthreads();
        ArgMin am;
        am.init();
        for (int j = threadIdx.x; j < k; j += 256) am.offer(euclid_sq_exact(xs, Cn + (int64_t)j * d, d), j);
        red_s[threadIdx.x] = am.s;
        red_j[threadIdx.x] = am.j;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (threadIdx.x < o) {
                am.merge(red_s[threadIdx.x + o], red_j[threadIdx.x + o]);
                red_s[threadIdx.x] = am.s;
                red_j[threadIdx.x] = am.j;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            labels[i] = (lloyd && am.j < 0) ? 0 : am.j;
            if (mind) mind[i] = am.j >= 0 ? __fsqrt_rn(am.s) : 3.402823466e+38f;
        }
    }
}

__device__ __forceinline__ void verify_stats(unsigned long long groups, unsigned long long fb, ScreenParams* prm) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        groups += __shfl_xor_sync(0xffffffffu, groups, o);
        fb += __shfl_xor_sync(0xffffffffu, fb, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (groups) atomicAdd(&prm->cand_chunks, groups);
        if (fb) atomicAdd(&prm->fallback_frames, fb);
    }
}

// d <= 16 and a center table that fits shared memory: thread per frame, frame in registers, the 8 centers of a
// candidate group read with 16-byte loads from the table.  Table layout: row stride ds (= d rounded up to 4) plus
// 4 floats of skew per 8-row group, so group g starts at bank (ds*8+4)*g mod 32 -- for ds = 4, 12 (and 8, 16 with
// the extra skew below) consecutive groups rotate through all eight 16-byte bank groups and the lanes of a warp,
// which look at unrelated groups, spread over the banks instead of piling onto one.
// CGT = 8: groups of 8 centers get a branch-free, fully unrolled evaluation (the frame and the table are zero padded to
// DREG columns; a padded column adds (0-0)^2 = +0 to lane 0, which leaves the sum's bits alone), 0: any group size.
template <int DREG, int CGT>
__global__ void __launch_bounds__(256, 4) screen_verify_table_kernel(const float* __restrict__ X, int64_t n, int d,
                                                                  const float* __restrict__ Cn, int k,
                                                                  const uint32_t* __restrict__ cand,
                                                                  const uint8_t* __restrict__ ncand,
                                                                  int32_t* __restrict__ labels,
                                                                  float* __restrict__ mind, int lloyd,
                                                                  ScreenParams* prm, uint32_t* __restrict__ fb_list,
                                                                  int gstride /* floats per 8-row group */, int cg,
                                                                  int vec /* row load width */) {
    extern __shared__ __align__(16) float ctab[];
    const int idb = cand_id_bits(cg);
    const uint32_t idm = cand_id_mask(cg);
    if (!prm->valid) return;
    constexpr int DS = DREG;  // row stride inside a group
    const int n_groups = (k + GROUP - 1) / GROUP;
    for (int t = threadIdx.x; t < n_groups * GROUP * DS; t += 256) {
        const int r = t / DS, c = t - r * DS;
        ctab[(r >> 3) * gstride + (r & 7) * DS + c] = (r < k && c < d) ? __ldg(Cn + (int64_t)r * d + c) : 0.f;
    }
    __syncthreads();
    const int d4 = d & ~3;
    const bool full_last = d == DREG;  // d is in (DREG-4, DREG]: the last 4-column block is a full block or a tail
    unsigned long long my_groups = 0, my_fb = 0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        float xr[DREG];
        load_row_padded<DREG>(X, i, d, vec, xr);
        const int nc = ncand[i];
        if (nc == 255) {
            fallback_push(prm, fb_list, i);
            my_fb += 1;
            continue;
        }
        const uint4* cp = reinterpret_cast<const uint4*>(cand + i * CAND_CAP);
        const uint4 p0 = cp[0];
        uint4 p1 = make_uint4(0, 0, 0, 0);
        if (nc > 4) p1 = cp[1];
        const uint32_t ent[CAND_CAP] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
        ArgMin am;
        am.init();
#pragma unroll
        for (int t = 0; t < CAND_CAP; ++t) {
            if (t < nc) {
                const int j0 = (int)(ent[t] & idm) * CHUNK;
                uint32_t mask = ent[t] >> idb;
                while (mask) {
                    const int q = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const int jb = j0 + q * cg;  // first center of the group; a group never crosses an 8-row block
                    const float4* c4 = reinterpret_cast<const float4*>(ctab + (size_t)(jb >> 3) * gstride + (jb & 7) * DS);
                    my_groups += 1;
                    if (CGT == 8 && jb + 8 <= k) {
                        float s8[8];
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            Lanes4 L;
                            L.init();
#pragma unroll
                            for (int e = 0; e < DREG - 4; e += 4) {
                                const float4 cv = c4[(c * DS + e) >> 2];
                                L.add4(xr[e], xr[e + 1], xr[e + 2], xr[e + 3], cv.x, cv.y, cv.z, cv.w);
                            }
                            const float4 cv = c4[(c * DS + DREG - 4) >> 2];
                            if (full_last) {
                                L.add4(xr[DREG - 4], xr[DREG - 3], xr[DREG - 2], xr[DREG - 1], cv.x, cv.y, cv.z, cv.w);
                            } else {  // 1..3 tail columns into lane 0, in order; the padded ones add +0
                                L.tail(xr[DREG - 4], cv.x);
                                L.tail(xr[DREG - 3], cv.y);
                                L.tail(xr[DREG - 2], cv.z);
                            }
                            s8[c] = L.result();
                        }
#pragma unroll
                        for (int c = 0; c < 8; ++c) am.offer(s8[c], jb + c);
                        continue;
                    }
                    const int jn = min(cg, k - jb);
#pragma unroll 2
                    for (int c = 0; c < cg; ++c) {
                        if (c < jn) {
                            Lanes4 L;
                            L.init();
#pragma unroll
                            for (int e = 0; e < DREG; e += 4) {
                                if (e < d4) {
                                    const float4 cv = c4[(c * DS + e) >> 2];
                                    L.add4(xr[e], xr[e + 1], xr[e + 2], xr[e + 3], cv.x, cv.y, cv.z, cv.w);
                                }
                            }
                            if (d4 < d) {
                                const float4 cv = c4[(c * DS + d4) >> 2];
                                const float ct[3] = {cv.x, cv.y, cv.z};
#pragma unroll
                                for (int e = 0; e < DREG; ++e)
                                    if (e >= d4 && e < d) L.tail(xr[e], ct[e & 3]);
                            }
                            am.offer(L.result(), jb + c);
                        }
                    }
                }
            }
        }
        labels[i] = (lloyd && am.j < 0) ? 0 : am.j;
        if (mind) mind[i] = am.j >= 0 ? __fsqrt_rn(am.s) : 3.402823466e+38f;
    }
    verify_stats(my_groups, my_fb, prm);
}


// the same for the listed screen (prune.cu): a candidate group is `cg` consecutive POSITIONS of the frame tile's center
// list; the ids are ascending along the list, so the scan order (lowest index wins ties) is unchanged.  Consecutive
// frames belong to the same tile and mostly share candidate groups, so the lanes of a warp read the same table rows
// (shared-memory broadcasts) where the unsorted kernel above serialises on bank conflicts.
template <int DREG>
__global__ void __launch_bounds__(256, 4) screen_verify_table_listed_kernel(
    const float* __restrict__ X, int64_t n, int d, const float* __restrict__ Cn, int k, const uint32_t* __restrict__ cand,
    const uint8_t* __restrict__ ncand, const uint16_t* __restrict__ tlist, int lcap, int unit_frames,
    int32_t* __restrict__ labels, int lloyd, ScreenParams* prm, uint32_t* __restrict__ fb_list, int gstride, int cg, int vec) {
    extern __shared__ __align__(16) float ctab[];
    const int idb = cand_id_bits(cg);
    const uint32_t idm = cand_id_mask(cg);
    if (!prm->valid) return;
    constexpr int DS = DREG;
    const int n_groups = (k + GROUP - 1) / GROUP;
    for (int t = threadIdx.x; t < n_groups * GROUP * DS; t += 256) {
        const int r = t / DS, c = t - r * DS;
        ctab[(r >> 3) * gstride + (r & 7) * DS + c] = (r < k && c < d) ? __ldg(Cn + (int64_t)r * d + c) : 0.f;
    }
    __syncthreads();
    const bool full_last = d == DREG;
    unsigned long long my_groups = 0, my_fb = 0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const int nc = ncand[i];
        if (nc == 255) {
            fallback_push(prm, fb_list, i);
            my_fb += 1;
            continue;
        }
        const uint16_t* tl = tlist + (size_t)(i / unit_frames) * lcap;
        if (nc == NCAND_DECIDED) {  // settled by the screen: no distance to evaluate, the frame row is not even read
            const int j = (int)__ldg(tl + cand[i * CAND_CAP]);
            if (j < k) labels[i] = j;
            else { fallback_push(prm, fb_list, i); my_fb += 1; }
            continue;
        }
        float xr[DREG];
        load_row_padded<DREG>(X, i, d, vec, xr);
        const uint4* cp = reinterpret_cast<const uint4*>(cand + i * CAND_CAP);
        const uint4 p0 = cp[0];
        uint4 p1 = make_uint4(0, 0, 0, 0);
        if (nc > 4) p1 = cp[1];
        const uint32_t ent[CAND_CAP] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
        ArgMin am;
        am.init();
#pragma unroll
        for (int t = 0; t < CAND_CAP; ++t) {
            if (t < nc) {
                const int pos0 = (int)(ent[t] & idm) * CHUNK;
                uint32_t mask = ent[t] >> idb;
                while (mask) {
                    const int q = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const int pb = pos0 + q * cg;
                    my_groups += 1;
                    if (cg == 8) {
                        // 8 ids with one 16-byte load; branch-free unrolled evaluation (a padding id reads row k-1 and is
                        // dropped before the argmin)
                        const uint4 iv = __ldg(reinterpret_cast<const uint4*>(tl + pb));
                        const int ids[8] = {(int)(iv.x & 0xffffu), (int)(iv.x >> 16), (int)(iv.y & 0xffffu), (int)(iv.y >> 16),
                                            (int)(iv.z & 0xffffu), (int)(iv.z >> 16), (int)(iv.w & 0xffffu), (int)(iv.w >> 16)};
                        float s8[8];
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const int j = min(ids[c], k - 1);
                            const float4* c4 = reinterpret_cast<const float4*>(ctab + (size_t)(j >> 3) * gstride + (j & 7) * DS);
                            Lanes4 L;
                            L.init();
#pragma unroll
                            for (int e = 0; e < DREG - 4; e += 4) {
                                const float4 cv = c4[e >> 2];
                                L.add4(xr[e], xr[e + 1], xr[e + 2], xr[e + 3], cv.x, cv.y, cv.z, cv.w);
                            }
                            const float4 cv = c4[(DREG - 4) >> 2];
                            if (full_last) {
                                L.add4(xr[DREG - 4], xr[DREG - 3], xr[DREG - 2], xr[DREG - 1], cv.x, cv.y, cv.z, cv.w);
                            } else {
                                L.tail(xr[DREG - 4], cv.x);
                                L.tail(xr[DREG - 3], cv.y);
                                L.tail(xr[DREG - 2], cv.z);
                            }
                            s8[c] = L.result();
                        }
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            if (ids[c] < k) am.offer(s8[c], ids[c]);
                        continue;
                    }
                    for (int c = 0; c < cg; ++c) {
                        const int j = (int)__ldg(tl + pb + c);
                        if (j >= k) continue;  // list padding
                        const float4* c4 = reinterpret_cast<const float4*>(ctab + (size_t)(j >> 3) * gstride + (j & 7) * DS);
                        Lanes4 L;
                        L.init();
#pragma unroll
                        for (int e = 0; e < DREG - 4; e += 4) {
                            const float4 cv = c4[e >> 2];
                            L.add4(xr[e], xr[e + 1], xr[e + 2], xr[e + 3], cv.x, cv.y, cv.z, cv.w);
                        }
                        const float4 cv = c4[(DREG - 4) >> 2];
                        if (full_last) {
                            L.add4(xr[DREG - 4], xr[DREG - 3], xr[DREG - 2], xr[DREG - 1], cv.x, cv.y, cv.z, cv.w);
                        } else {  // 1..3 tail columns into lane 0, in order; the padded ones add +0
                            L.tail(xr[DREG - 4], cv.x);
                            L.tail(xr[DREG - 3], cv.y);
                            L.tail(xr[DREG - 2], cv.z);
                        }
                        am.offer(L.result(), j);
                    }
                }
            }
        }
        labels[i] = (lloyd && am.j < 0) ? 0 : am.j;
    }
    verify_stats(my_groups, my_fb, prm);
}


// ---- listed verify for wide rows (d > 16, d % 4 == 0): one CTA per 128-frame tile ---------------------------------------
// The frames of a tile are neighbours in space (sorted by label), so their candidate centers are a handful of list
// positions that most of the 128 frames share.  The CTA marks the positions its frames need in a shared bitmap, copies
// exactly those center rows into shared memory once (a few KB from L2 per tile instead of 4*d bytes per frame and
// candidate), and every thread then walks ITS frame row once per four candidates: the row streams through registers
// (16-byte loads), the center rows come out of shared memory (lanes of a warp mostly read the same row: broadcast), and
// every (frame, center) sum stays one thread's sequential Lanes4 sum in the reference order.
// (Measured at 1.25e7 x 64, k=2000, Lloyd step: 7.3 ms with this kernel, 7.8 ms with the 8-lanes-per-frame direct kernel,
//  9.6 ms with a variant that streamed the tile's frame rows through shared memory by cp.async one slab ahead at 2 CTAs
//  per SM -- every one of them is bound by the chain of dependent loads per tile / frame group (candidate entries -> list
//  ids -> rows), not by bandwidth: 3.6 GB at 1.2 TB/s.  profiles/r02_notes.md)
#ifndef B2K_VT_MINBLOCKS
#define B2K_VT_MINBLOCKS 8  // register cap: measured Lloyd step at cfg3 7.9 / 7.3 / 7.1 ms for 4 / 6 / 8 CTAs per SM
#endif
static constexpr int VT_CMAX = 16;        // candidate centers a thread keeps; frames with more take the exact fallback
static constexpr int VT_WORDS = 8192 / 32;  // bitmap over list positions (lcap <= 8192)

__global__ void __launch_bounds__(TILE_M, B2K_VT_MINBLOCKS) screen_verify_tile_listed_kernel(
    const float* __restrict__ X, int64_t n, int d, const float* __restrict__ Cn, int k, const uint32_t* __restrict__ cand,
    const uint8_t* __restrict__ ncand, const uint16_t* __restrict__ tlist, int lcap, int ushift, int32_t* __restrict__ labels,
    int lloyd, ScreenParams* prm, uint32_t* __restrict__ fb_list, int cg, int batch /* center rows per shared batch */) {
    extern __shared__ __align__(16) float vts[];  // [batch][d + 4] center rows
    __shared__ uint32_t bits[VT_WORDS];
    __shared__ uint16_t pre[VT_WORDS];            // set bits below word w
    __shared__ uint16_t slot_pos[1024];           // list position of slot s (first 1024 slots; more -> fallback below)
    __shared__ int s_total;
    if (!prm->valid) return;
    const int idb = cand_id_bits(cg);
    const uint32_t idm = cand_id_mask(cg);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rs = d + 4;
    const int nwords = (lcap + 31) >> 5;
    const int64_t n_tiles = (n + TILE_M - 1) / TILE_M;
    unsigned long long my_groups = 0, my_fb = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t i = tile * TILE_M + tid;
        const bool live = i < n;
        const uint16_t* tl = tlist + (size_t)(tile >> ushift) * lcap;
        for (int w = tid; w < nwords; w += TILE_M) bits[w] = 0u;
        __syncthreads();
        // ---- my candidate list positions (ascending) ----
        int nc = live ? (int)ncand[i] : 0;
        uint16_t cpos[VT_CMAX];
        int ncs = 0;
        bool fb = live && nc == 255;
        int decided_j = -1;
        if (live && nc == NCAND_DECIDED) {  // settled by the screen
            decided_j = (int)__ldg(tl + cand[i * CAND_CAP]);
            if (decided_j >= k) { decided_j = -1; fb = true; }
            nc = 255;  // nothing to evaluate
        }
        if (live && nc != 255) {
            const uint4* cp = reinterpret_cast<const uint4*>(cand + i * CAND_CAP);
            const uint4 p0 = cp[0];
            uint4 p1 = make_uint4(0, 0, 0, 0);
            if (nc > 4) p1 = cp[1];
            const uint32_t ent[CAND_CAP] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
            for (int t = 0; t < CAND_CAP; ++t) {
                if (t < nc) {
                    const int pos0 = (int)(ent[t] & idm) * CHUNK;
                    uint32_t mask = ent[t] >> idb;
                    while (mask) {
                        const int q = __ffs(mask) - 1;
                        mask &= mask - 1;
                        my_groups += 1;
                        for (int c = 0; c < cg; ++c) {
                            const int pos = pos0 + q * cg + c;
                            if (ncs < VT_CMAX) {
#pragma unroll
                                for (int u = 0; u < VT_CMAX; ++u)
                                    if (u == ncs) cpos[u] = (uint16_t)pos;
                            }
                            ++ncs;
                        }
                    }
                }
            }
            if (ncs > VT_CMAX) { fb = true; ncs = 0; }
        }
#pragma unroll
        for (int u = 0; u < VT_CMAX; ++u)
            if (u < ncs) atomicOr(&bits[cpos[u] >> 5], 1u << (cpos[u] & 31));
        __syncthreads();
        // ---- prefix over the bitmap words (warp 0), slot -> position table ----
        if (warp == 0) {
            int run = 0;
            for (int w0 = 0; w0 < nwords; w0 += 32) {
                const int w = w0 + lane;
                const int c = w < nwords ? __popc(bits[w]) : 0;
                int inc = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += v;
                }
                if (w < nwords) pre[w] = (uint16_t)(run + inc - c);
                run += __shfl_sync(0xffffffffu, inc, 31);
            }
            if (lane == 0) s_total = run;
        }
        __syncthreads();
        const int total = s_total;
        for (int w = tid; w < nwords; w += TILE_M) {
            uint32_t b = bits[w];
            int sidx = pre[w];
            while (b) {
                const int bit = __ffs(b) - 1;
                b &= b - 1;
                if (sidx < 1024) slot_pos[sidx] = (uint16_t)(w * 32 + bit);
                ++sidx;
            }
        }
        if (total > 1024) { if (!fb && live && nc != 255) fb = true; ncs = 0; }  // (never seen: 128 frames x 16 centers = 2048 at most)
        __syncthreads();
        // my candidates as (slot, center id)
        uint16_t cslot[VT_CMAX];
#pragma unroll
        for (int u = 0; u < VT_CMAX; ++u) {
            cslot[u] = 0;
            if (u < ncs) {
                const int pos = cpos[u];
                cslot[u] = (uint16_t)(pre[pos >> 5] + __popc(bits[pos >> 5] & ((1u << (pos & 31)) - 1u)));
            }
        }
        ArgMin am;
        am.init();
        const float4* x4 = reinterpret_cast<const float4*>(X + (live ? i : 0) * d);
        const int nv = d >> 2;
        int cur = 0;  // my next candidate (slots ascend with u)
        for (int b0 = 0; b0 < min(total, 1024); b0 += batch) {
            const int bn = min(batch, min(total, 1024) - b0);
            __syncthreads();  // the previous batch has been consumed
            // copy the batch's center rows (a padding id -- >= k -- leaves its row untouched: it is never offered)
            for (int t = tid; t < bn * nv; t += TILE_M) {
                const int r = t / nv, c = t - r * nv;
                const int j = (int)tl[slot_pos[b0 + r]];
                if (j < k) *reinterpret_cast<float4*>(vts + (size_t)r * rs + c * 4) =
                    __ldg(reinterpret_cast<const float4*>(Cn + (int64_t)j * d) + c);
            }
            __syncthreads();
            auto slot_at = [&](int idx) {  // cslot[idx] without dynamic register indexing
                int v = 0x7fffffff;
#pragma unroll
                for (int u = 0; u < VT_CMAX; ++u)
                    if (u == idx) v = (int)cslot[u];
                return v;
            };
            while (cur < ncs && slot_at(cur) < b0 + bn) {
                // up to four candidates of this batch at a time: the frame row is read once for all of them
                int cnt4 = 0;
                int sl[4] = {0, 0, 0, 0};
#pragma unroll
                for (int u = 0; u < VT_CMAX; ++u) {
                    if (u >= cur && u < ncs && cnt4 < 4 && (int)cslot[u] < b0 + bn) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (q == cnt4) sl[q] = (int)cslot[u] - b0;
                        ++cnt4;
                    }
                }
                Lanes4 L0, L1, L2, L3;
                L0.init(); L1.init(); L2.init(); L3.init();
                const float4* c0 = reinterpret_cast<const float4*>(vts + (size_t)sl[0] * rs);
                const float4* c1 = reinterpret_cast<const float4*>(vts + (size_t)sl[1] * rs);
                const float4* c2 = reinterpret_cast<const float4*>(vts + (size_t)sl[2] * rs);
                const float4* c3 = reinterpret_cast<const float4*>(vts + (size_t)sl[3] * rs);
                int t = 0;
                for (; t + 4 <= nv; t += 4) {  // four 16-byte loads of the frame row in flight per thread
                    float4 xv[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) xv[u] = __ldg(x4 + t + u);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float4 a = c0[t + u];
                        L0.add4(xv[u].x, xv[u].y, xv[u].z, xv[u].w, a.x, a.y, a.z, a.w);
                        if (cnt4 > 1) { const float4 b = c1[t + u]; L1.add4(xv[u].x, xv[u].y, xv[u].z, xv[u].w, b.x, b.y, b.z, b.w); }
                        if (cnt4 > 2) { const float4 c = c2[t + u]; L2.add4(xv[u].x, xv[u].y, xv[u].z, xv[u].w, c.x, c.y, c.z, c.w); }
                        if (cnt4 > 3) { const float4 e = c3[t + u]; L3.add4(xv[u].x, xv[u].y, xv[u].z, xv[u].w, e.x, e.y, e.z, e.w); }
                    }
                }
                for (; t < nv; ++t) {
                    const float4 xv = __ldg(x4 + t);
                    const float4 a = c0[t];
                    L0.add4(xv.x, xv.y, xv.z, xv.w, a.x, a.y, a.z, a.w);
                    if (cnt4 > 1) { const float4 b = c1[t]; L1.add4(xv.x, xv.y, xv.z, xv.w, b.x, b.y, b.z, b.w); }
                    if (cnt4 > 2) { const float4 c = c2[t]; L2.add4(xv.x, xv.y, xv.z, xv.w, c.x, c.y, c.z, c.w); }
                    if (cnt4 > 3) { const float4 e = c3[t]; L3.add4(xv.x, xv.y, xv.z, xv.w, e.x, e.y, e.z, e.w); }
                }
                const float res[4] = {L0.result(), L1.result(), L2.result(), L3.result()};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (q < cnt4) {
                        const int j = (int)tl[slot_pos[b0 + sl[q]]];
                        if (j < k) am.offer(res[q], j);
                    }
                }
                cur += cnt4;
            }
        }
        if (live) {
            if (fb) {
                fallback_push(prm, fb_list, i);
                my_fb += 1;
            } else if (decided_j >= 0) {
                labels[i] = decided_j;
            } else {
                labels[i] = (lloyd && am.j < 0) ? 0 : am.j;
            }
        }
        __syncthreads();  // bits / slot tables are rewritten by the next tile
    }
    verify_stats(my_groups, my_fb, prm);
}


// ---- listed verify for wide rows, one THREAD per frame (d % 4 == 0) -------------------------------------------------
// The 8-lanes-per-frame kernels keep 4 frames per warp in flight and, with ~2.6 candidate centers per frame, 5 of 8 lanes
// idle; the tile kernel serialises a tile behind half a dozen barriers.  Sorted frames make the plain mapping the best
// one: a lane owns a frame, streams its row once per four candidates (32 bytes per load step: whole sectors) and reads the
// candidates' center rows through L1 -- the 32 lanes of a warp are neighbours in space and mostly name the same handful
// of centers, so those loads are broadcasts that hit L1.  32 frames per warp in flight, no shared memory, no barrier; every
// (frame, center) sum is still one thread's sequential Lanes4 sum in the reference order.
#ifndef B2K_VF_MINBLOCKS
#define B2K_VF_MINBLOCKS 4  // measured Lloyd step at cfg3: 6.27 / 6.16 / 7.24 / 7.96 ms for 3 / 4 / 5 / 6 CTAs of 256 threads per SM
#endif
__global__ void __launch_bounds__(256, B2K_VF_MINBLOCKS) screen_verify_frame_listed_kernel(
    const float* __restrict__ X, int64_t n, int d, const float* __restrict__ Cn, int k, const uint32_t* __restrict__ cand,
    const uint8_t* __restrict__ ncand, const uint16_t* __restrict__ tlist, int lcap, int unit_frames,
    int32_t* __restrict__ labels, int lloyd, ScreenParams* prm, uint32_t* __restrict__ fb_list, int cg) {
    if (!prm->valid) return;
    const int idb = cand_id_bits(cg);
    const uint32_t idm = cand_id_mask(cg);
    const int nv = d >> 2;
    unsigned long long my_groups = 0, my_fb = 0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const int nc = ncand[i];
        if (nc == 255) {
            fallback_push(prm, fb_list, i);
            my_fb += 1;
            continue;
        }
        const uint16_t* tl = tlist + (size_t)(i / unit_frames) * lcap;
        if (nc == NCAND_DECIDED) {
            const int j = (int)__ldg(tl + cand[i * CAND_CAP]);
            if (j < k) labels[i] = j;
            else { fallback_push(prm, fb_list, i); my_fb += 1; }
            continue;
        }
        const uint4* cp = reinterpret_cast<const uint4*>(cand + i * CAND_CAP);
        const uint4 p0 = cp[0];
        uint4 p1 = make_uint4(0, 0, 0, 0);
        if (nc > 4) p1 = cp[1];
        const uint32_t ent[CAND_CAP] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
        const float4* x4 = reinterpret_cast<const float4*>(X + i * d);
        ArgMin am;
        am.init();
        // walk the candidate centers in ascending list order, four at a time
        int t = 0;            // entry
        uint32_t mask = nc > 0 ? ent[0] >> idb : 0u;
        int sub = 0;          // center inside the current group
        int q = mask ? __ffs(mask) - 1 : 0;
        bool more = nc > 0 && mask != 0;
        while (more) {
            int js[4] = {0, 0, 0, 0};
            int cnt4 = 0;
            while (more && cnt4 < 4) {
                uint32_t e = ent[0];
#pragma unroll
                for (int u = 1; u < CAND_CAP; ++u)
                    if (u == t) e = ent[u];
                const int pos = (int)(e & idm) * CHUNK + q * cg + sub;
                const int j = (int)__ldg(tl + pos);
                if (j < k) {
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (u == cnt4) js[u] = j;
                    ++cnt4;
                }
                // advance: next center of the group, next group of the entry, next entry
                if (++sub == cg) {
                    sub = 0;
                    my_groups += 1;
                    mask &= mask - 1;
                    if (mask) {
                        q = __ffs(mask) - 1;
                    } else {
                        ++t;
                        more = false;
                        while (t < nc) {
                            uint32_t e2 = ent[0];
#pragma unroll
                            for (int u = 1; u < CAND_CAP; ++u)
                                if (u == t) e2 = ent[u];
                            mask = e2 >> idb;
                            if (mask) { q = __ffs(mask) - 1; more = true; break; }
                            ++t;
                        }
                    }
                }
            }
            if (cnt4 == 0) break;
            const float4* c0 = reinterpret_cast<const float4*>(Cn + (int64_t)js[0] * d);
            const float4* c1 = reinterpret_cast<const float4*>(Cn + (int64_t)js[1] * d);
            const float4* c2 = reinterpret_cast<const float4*>(Cn + (int64_t)js[2] * d);
            const float4* c3 = reinterpret_cast<const float4*>(Cn + (int64_t)js[3] * d);
            Lanes4 L0, L1, L2, L3;
            L0.init(); L1.init(); L2.init(); L3.init();
            int v = 0;
            for (; v + 2 <= nv; v += 2) {
                const float4 xa = __ldg(x4 + v), xb = __ldg(x4 + v + 1);
                {
                    const float4 a = __ldg(c0 + v), b = __ldg(c0 + v + 1);
                    L0.add4(xa.x, xa.y, xa.z, xa.w, a.x, a.y, a.z, a.w);
                    L0.add4(xb.x, xb.y, xb.z, xb.w, b.x, b.y, b.z, b.w);
                }
                if (cnt4 > 1) {
                    const float4 a = __ldg(c1 + v), b = __ldg(c1 + v + 1);
                    L1.add4(xa.x, xa.y, xa.z, xa.w, a.x, a.y, a.z, a.w);
                    L1.add4(xb.x, xb.y, xb.z, xb.w, b.x, b.y, b.z, b.w);
                }
                if (cnt4 > 2) {
                    const float4 a = __ldg(c2 + v), b = __ldg(c2 + v + 1);
                    L2.add4(xa.x, xa.y, xa.z, xa.w, a.x, a.y, a.z, a.w);
                    L2.add4(xb.x, xb.y, xb.z, xb.w, b.x, b.y, b.z, b.w);
                }
                if (cnt4 > 3) {
                    const float4 a = __ldg(c3 + v), b = __ldg(c3 + v + 1);
                    L3.add4(xa.x, xa.y, xa.z, xa.w, a.x, a.y, a.z, a.w);
                    L3.add4(xb.x, xb.y, xb.z, xb.w, b.x, b.y, b.z, b.w);
                }
            }
            for (; v < nv; ++v) {
                const float4 xa = __ldg(x4 + v);
                { const float4 a = __ldg(c0 + v); L0.add4(xa.x, xa.y, xa.z, xa.w, a.x, a.y, a.z, a.w); }
                if (cnt4 > 1) { const float4 a = __ldg(c1 + v); L1.add4(xa.x, xa.y, xa.z, xa.w, a.x, a.y, a.z, a.w); }
                if (cnt4 > 2) { const float4 a = __ldg(c2 + v); L2.add4(xa.x, xa.y, xa.z, xa.w, a.x, a.y, a.z, a.w); }
                if (cnt4 > 3) { const float4 a = __ldg(c3 + v); L3.add4(xa.x, xa.y, xa.z, xa.w, a.x, a.y, a.z, a.w); }
            }
            am.offer(L0.result(), js[0]);
            if (cnt4 > 1) am.offer(L1.result(), js[1]);
            if (cnt4 > 2) am.offer(L2.result(), js[2]);
            if (cnt4 > 3) am.offer(L3.result(), js[3]);
        }
        labels[i] = (lloyd && am.j < 0) ? 0 : am.j;
    }
    verify_stats(my_groups, my_fb, prm);
}

// d <= 16, table too large for shared memory: 8 lanes per frame (4 frames per warp), lane `sub` evaluates center `sub` of every candidate group
// with the frame in registers; the center table sits in shared memory when it fits (row stride rs = 4 mod 8
// floats: the 8 rows of a group then cover all 32 banks, so a 16-byte load per lane is conflict free).
template <int DREG>
__global__ void __launch_bounds__(256) screen_verify_small_kernel(const float* __restrict__ X, int64_t n, int d,
                                                                  const float* __restrict__ Cn, int k,
                                                                  const uint32_t* __restrict__ cand,
                                                                  const uint8_t* __restrict__ ncand,
                                                                  int32_t* __restrict__ labels,
                                                                  float* __restrict__ mind, int lloyd,
                                                                  ScreenParams* prm, uint32_t* __restrict__ fb_list,
                                                                  int use_smem, int rs, int cg) {
    extern __shared__ __align__(16) float ctab[];
    const int idb = cand_id_bits(cg);
    const uint32_t idm = cand_id_mask(cg);
    if (!prm->valid) return;
    if (use_smem) {
        for (int t = threadIdx.x; t < k * rs; t += 256) {
            const int r = t / rs, c = t - r * rs;
            ctab[t] = c < d ? __ldg(Cn + (int64_t)r * d + c) : 0.f;
        }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, sub = lane & 7, slot = lane >> 3;
    const int d4 = d & ~3;
    unsigned long long my_groups = 0, my_fb = 0;
    const int64_t warp_global = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * 256) >> 5;
    for (int64_t base = warp_global * 4; base < n; base += n_warps * 4) {
        const int64_t i = base + slot;
        const bool live = i < n;
        float xr[DREG];
#pragma unroll
        for (int e = 0; e < DREG; ++e) xr[e] = (live && e < d) ? __ldg(X + i * d + e) : 0.f;
        const int nc = live ? (int)ncand[i] : 0;
        uint32_t ent[CAND_CAP];
        {
            const uint4* cp = reinterpret_cast<const uint4*>(cand + (live ? i : 0) * CAND_CAP);
            const uint4 p0 = cp[0];
            uint4 p1 = make_uint4(0, 0, 0, 0);
            if (nc > 4 && nc != 255) p1 = cp[1];
            ent[0] = p0.x; ent[1] = p0.y; ent[2] = p0.z; ent[3] = p0.w;
            ent[4] = p1.x; ent[5] = p1.y; ent[6] = p1.z; ent[7] = p1.w;
        }
        ArgMin am;
        am.init();
        auto eval_group = [&](int jb) {  // lanes sub < cg take the centers of the group that starts at jb
            const int j = jb + sub;
            if (sub < cg && j < k) {
                Lanes4 L;
                L.init();
                if (use_smem) {
                    const float4* c4 = reinterpret_cast<const float4*>(ctab + (size_t)j * rs);
#pragma unroll
                    for (int e = 0; e < DREG; e += 4) {
                        if (e < d4) {
                            const float4 cv = c4[e >> 2];
                            L.add4(xr[e], xr[e + 1], xr[e + 2], xr[e + 3], cv.x, cv.y, cv.z, cv.w);
                        }
                    }
                    if (d4 < d) {
                        const float4 cv = c4[d4 >> 2];
                        const float ct[3] = {cv.x, cv.y, cv.z};
#pragma unroll
                        for (int e = 0; e < DREG; ++e)
                            if (e >= d4 && e < d) L.tail(xr[e], ct[e & 3]);
                    }
                } else {
                    const float* c = Cn + (int64_t)j * d;
#pragma unroll
                    for (int e = 0; e < DREG; e += 4)
                        if (e < d4)
                            L.add4(xr[e], xr[e + 1], xr[e + 2], xr[e + 3], __ldg(c + e), __ldg(c + e + 1), __ldg(c + e + 2),
                                   __ldg(c + e + 3));
#pragma unroll
                    for (int e = 0; e < DREG; ++e)
                        if (e >= d4 && e < d) L.tail(xr[e], __ldg(c + e));
                }
                am.offer(L.result(), j);  // a lane meets its centers in ascending order
            }
        };
        if (nc == 255) {  // the screen could not bound this frame
            if (sub == 0) { fallback_push(prm, fb_list, i); my_fb += 1; }
        } else {
#pragma unroll
            for (int t = 0; t < CAND_CAP; ++t) {
                if (t < nc) {
                    const int j0 = (int)(ent[t] & idm) * CHUNK;
                    uint32_t mask = ent[t] >> idb;
                    while (mask) {
                        const int q = __ffs(mask) - 1;
                        mask &= mask - 1;
                        eval_group(j0 + q * cg);
                        if (sub == 0) my_groups += 1;
                    }
                }
            }
        }
        // combine the 8 lanes of the frame: lexicographic (sqrt(s), j), order free
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const float so = __shfl_xor_sync(0xffffffffu, am.s, o);
            const int32_t jo = __shfl_xor_sync(0xffffffffu, am.j, o);
            am.merge(so, jo);
        }
        if (live && sub == 0 && nc != 255) {
            labels[i] = (lloyd && am.j < 0) ? 0 : am.j;
            if (mind) mind[i] = am.j >= 0 ? __fsqrt_rn(am.s) : 3.402823466e+38f;
        }
    }
    verify_stats(my_groups, my_fb, prm);
}

// d > 16, d % 4 == 0: 8 lanes per frame (4 frames per warp, lane `sub` <-> center `sub` of the candidate group).
// The dimension is walked in slabs of 32 columns; per slab the warp copies, with 16-byte cp.async (LDGSTS, no
// register staging), the 4x32 slab of its frames and the 8x32 slab of each frame's candidate group (8 consecutive
// center rows) into a double-buffered shared-memory stage, then every lane advances ITS OWN four interleaved
// partial sums (Lanes4) over the slab: the reference's summation order is untouched.  Row stride 36 floats
// (odd multiple of 4): the 16-byte reads of the 8 lanes of a frame hit 8 different bank groups.
// (A TMA-box variant -- 8 rows x 128 B, SWIZZLE_128B -- ran at the same speed: the limiter is the L2->SM
// traffic of 8 center rows per candidate group, not the copy mechanism; profiles/r01_notes.md.)
static constexpr int VC_ROW = 36;                              // floats per staged center row
static constexpr int VC_FRAME = GROUP * VC_ROW;                // floats per frame's group slab
static constexpr int VC_STAGE = 4 * VC_FRAME + 4 * 32;         // + the 4 frame slabs
static constexpr int VC_WARP_FLOATS = 2 * VC_STAGE;


__global__ void __launch_bounds__(256) screen_verify_stream_kernel(const float* __restrict__ X, int64_t n, int d,
                                                                   const float* __restrict__ Cn, int k,
                                                                   const uint32_t* __restrict__ cand,
                                                                   const uint8_t* __restrict__ ncand,
                                                                   int32_t* __restrict__ labels,
                                                                   float* __restrict__ mind, int lloyd,
                                                                   ScreenParams* prm, uint32_t* __restrict__ fb_list,
                                                                   int cg) {
    extern __shared__ __align__(16) float vsm[];
    if (!prm->valid) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane & 7, slot = lane >> 3;
    // a round evaluates 8 centers per frame = S candidate groups of cg centers: lane sub <-> (group slot, center)
    const int idb = cand_id_bits(cg), ng = CHUNK / cg, S = GROUP / cg;
    const uint32_t idm = cand_id_mask(cg);
    const int gslot = sub / cg, within = sub - gslot * cg;
    float* wb = vsm + (size_t)warp * VC_WARP_FLOATS;
    const int nslab = (d + 31) >> 5;
    unsigned long long my_groups = 0, my_fb = 0;
    const int64_t n_warps = ((int64_t)gridDim.x * 256) >> 5;
    const int64_t warp_global = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    const unsigned seg_mask = 0xffu << (slot * 8);

    // candidate metadata of the NEXT quad is requested one quad ahead
    int nc_n = 0;
    uint32_t ent_n = 0;
    auto fetch_meta = [&](int64_t b) {
        const int64_t fi = b + slot;
        nc_n = fi < n ? (int)ncand[fi] : 0;
        ent_n = (fi < n && sub < CAND_CAP) ? cand[fi * CAND_CAP + sub] : 0u;
    };
    int64_t base = warp_global * 4;
    if (base < n) fetch_meta(base);
    for (; base < n; base += n_warps * 4) {
        const int64_t i = base + slot;
        const bool live = i < n;
        const int nc = nc_n;
        uint32_t ent = (nc != 255 && sub < nc) ? ent_n : 0u;
        if (base + n_warps * 4 < n) fetch_meta(base + n_warps * 4);
        const int pc = __popc(ent >> idb);
        int off = pc;  // inclusive scan over the 8 lanes of the frame
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, off, o, 8);
            if (sub >= o) off += v;
        }
        int total = __shfl_sync(0xffffffffu, off, 7, 8);
        off -= pc;
        if (nc == 255) total = 0;  // queued for the fallback kernel below
        int rounds = (total + S - 1) / S;
#pragma unroll
        for (int o = 8; o < 32; o <<= 1) rounds = max(rounds, __shfl_xor_sync(0xffffffffu, rounds, o));

        // candidate group (index into the k/cg groups) this lane works on in round r: the (r*S + gslot)-th group of
        // its frame's flattened candidate list (-1: none); warp-collective
        auto group_of_round = [&](int r) -> int {
            const int gi = r * S + gslot;
            int g = -1;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const uint32_t et = __shfl_sync(0xffffffffu, ent, t, 8);
                const int ot = __shfl_sync(0xffffffffu, off, t, 8);
                const int pt = __shfl_sync(0xffffffffu, pc, t, 8);
                if (gi >= ot && gi < ot + pt) g = (int)(et & idm) * ng + nth_set_bit(et >> idb, gi - ot);
            }
            return gi < total ? g : -1;
        };
        // start the copies of slab `sl` for round-group g into stage `stg`; warp-collective
        auto issue = [&](int g, int sl, int stg) {
            float* st = wb + stg * VC_STAGE;
            const int col0 = sl * 32;
            __syncwarp();  // every lane is done reading stage `stg`
            {   // frame slabs: lane (slot, sub) copies 16 bytes of frame `slot`
                const int col = col0 + sub * 4;
                if (live && col < d) cp_async16(st + 4 * VC_FRAME + slot * 32 + sub * 4, X + i * d + col);
            }
#pragma unroll
            for (int f = 0; f < 4; ++f) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    // staged row `row` of frame f belongs to the lane (f, sub = row): its group and center
                    const int t = lane + 32 * h, row = t >> 3, c4 = t & 7;
                    const int gr = __shfl_sync(0xffffffffu, g, f * 8 + row);
                    const int64_t j = (int64_t)gr * cg + (row % cg);
                    if (gr >= 0 && j < k && col0 + c4 * 4 < d)
                        cp_async16(st + f * VC_FRAME + row * VC_ROW + c4 * 4, Cn + j * d + col0 + c4 * 4);
                }
            }
            cp_async_commit();
        };

        ArgMin am;
        am.init();
        const int steps = rounds * nslab;
        int g_cur = -1, g_nxt = -1;
        if (steps > 0) {
            g_cur = group_of_round(0);
            issue(g_cur, 0, 0);
        }
        Lanes4 L;
        L.init();
        int r = 0, sl = 0;
        for (int step = 0; step < steps; ++step) {
            const int stg = step & 1;
            int nr = r, nsl = sl + 1;
            if (nsl == nslab) { nsl = 0; nr = r + 1; }
            if (step + 1 < steps) {
                g_nxt = (nsl == 0) ? group_of_round(nr) : g_cur;
                issue(g_nxt, nsl, stg ^ 1);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncwarp();  // every lane's copies of this stage have landed
            const int j = g_cur >= 0 ? g_cur * cg + within : -1;
            if (j >= 0 && j < k) {
                const float* st = wb + stg * VC_STAGE;
                const float* cr = st + slot * VC_FRAME + sub * VC_ROW;
                const float* xt = st + 4 * VC_FRAME + slot * 32;
                const int w = min(32, d - sl * 32);  // multiple of 4 (d % 4 == 0)
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (c * 4 < w) {
                        const float4 xv = *reinterpret_cast<const float4*>(xt + c * 4);
                        const float4 cv = *reinterpret_cast<const float4*>(cr + c * 4);
                        L.add4(xv.x, xv.y, xv.z, xv.w, cv.x, cv.y, cv.z, cv.w);
                    }
                }
            }
            if (sl == nslab - 1) {  // distance complete
                if (j >= 0 && j < k) am.offer(L.result(), j);
                if (g_cur >= 0 && within == 0 && nc != 255) my_groups += 1;
                L.init();
            }
            r = nr;
            sl = nsl;
            g_cur = (sl == 0) ? g_nxt : g_cur;
        }
        if (live && nc == 255 && sub == 0) { fallback_push(prm, fb_list, i); my_fb += 1; }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const float so = __shfl_xor_sync(0xffffffffu, am.s, o);
            const int32_t jo = __shfl_xor_sync(0xffffffffu, am.j, o);
            am.merge(so, jo);
        }
        if (live && sub == 0 && nc != 255) {
            labels[i] = (lloyd && am.j < 0) ? 0 : am.j;
            if (mind) mind[i] = am.j >= 0 ? __fsqrt_rn(am.s) : 3.402823466e+38f;
        }
    }
    verify_stats(my_groups, my_fb, prm);
}

// d > 16, any d: 8 lanes per frame again (4 frames per warp, lane `sub` <-> center `sub` of the group), but the rows are
// too wide for registers / a resident table, so the dimension is walked in slabs of SLAB columns: per round the
// warp copies, coalesced, the slab of its 4 frames and of the 4 candidate groups (8 contiguous center rows each)
// into shared memory and every lane advances ITS OWN four interleaved partial sums (Lanes4) over the slab -- the
// sums never leave the thread, so the reference's summation order is untouched.  Row stride SLAB+4 floats
// (odd multiple of 4): the 8 rows of a group cover all banks, 16-byte loads are conflict free.
static constexpr int SLAB = 32;
static constexpr int SLAB_RS = SLAB + 4;
static constexpr int VW_FLOATS = 4 * (GROUP * SLAB_RS + SLAB);  // shared floats per warp

__global__ void __launch_bounds__(256) screen_verify_wide_kernel(const float* __restrict__ X, int64_t n, int d,
                                                                 const float* __restrict__ Cn, int k,
                                                                 const uint32_t* __restrict__ cand,
                                                                 const uint8_t* __restrict__ ncand,
                                                                 int32_t* __restrict__ labels,
                                                                 float* __restrict__ mind, int lloyd,
                                                                 ScreenParams* prm, uint32_t* __restrict__ fb_list,
                                                                 int cg) {
    extern __shared__ __align__(16) float vsm[];
    if (!prm->valid) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane & 7, slot = lane >> 3;
    const int idb = cand_id_bits(cg), ng = CHUNK / cg, S = GROUP / cg;
    const uint32_t idm = cand_id_mask(cg);
    const int gslot = sub / cg, within = sub - gslot * cg;
    float* wbuf = vsm + (size_t)warp * VW_FLOATS;
    float* cb = wbuf + (size_t)slot * (GROUP * SLAB_RS);   // this frame's group rows
    float* xb = wbuf + 4 * (GROUP * SLAB_RS) + slot * SLAB;  // this frame's slab
    const int d4 = d & ~3;
    unsigned long long my_groups = 0, my_fb = 0;
    const int64_t n_warps = ((int64_t)gridDim.x * 256) >> 5;
    const int64_t warp_global = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    const unsigned seg_mask = 0xffu << (slot * 8);
    for (int64_t base = warp_global * 4; base < n; base += n_warps * 4) {
        const int64_t i = base + slot;
        const bool live = i < n;
        const int nc = live ? (int)ncand[i] : 0;
        // lane (slot, sub) holds candidate entry `sub` of its frame; prefix-count the groups of the frame
        uint32_t ent = 0;
        if (live && nc != 255 && sub < nc) ent = cand[i * CAND_CAP + sub];
        const int pc = __popc(ent >> idb);
        int off = pc;  // inclusive scan over the 8 lanes of the frame
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, off, o, 8);
            if (sub >= o) off += v;
        }
        int total = __shfl_sync(0xffffffffu, off, 7, 8);
        off -= pc;
        if (nc == 255) total = 0;  // queued for the fallback kernel below
        int rounds = (total + S - 1) / S;
#pragma unroll
        for (int o = 8; o < 32; o <<= 1) rounds = max(rounds, __shfl_xor_sync(0xffffffffu, rounds, o));
        ArgMin am;
        am.init();
        for (int r = 0; r < rounds; ++r) {
            // candidate group of this lane in round r: the (r*S + gslot)-th of its frame's flattened list (-1: none)
            int g = -1;
            {
                const int gi = r * S + gslot;
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const uint32_t et = __shfl_sync(0xffffffffu, ent, t, 8);
                    const int ot = __shfl_sync(0xffffffffu, off, t, 8);
                    const int pt = __shfl_sync(0xffffffffu, pc, t, 8);
                    if (gi >= ot && gi < ot + pt) g = (int)(et & idm) * ng + nth_set_bit(et >> idb, gi - ot);
                }
                if (gi >= total) g = -1;
            }
            const int j = g >= 0 ? g * cg + within : -1;
            Lanes4 L;
            L.init();
            for (int e0 = 0; e0 < d; e0 += SLAB) {
                const int w = min(SLAB, d - e0);  // slab width
                __syncwarp();
                // ---- stage: 4 frame slabs + 4 x 8 center-row slabs (each row piece contiguous in global memory)
                for (int t = lane; t < 4 * w; t += 32) {
                    const int f = t / w, c = t - f * w;
                    const int64_t fi = base + f;
                    wbuf[4 * (GROUP * SLAB_RS) + f * SLAB + c] = fi < n ? __ldg(X + fi * d + e0 + c) : 0.f;
                }
#pragma unroll
                for (int f = 0; f < 4; ++f) {
                    float* dst = wbuf + (size_t)f * (GROUP * SLAB_RS);
                    for (int t = lane; t < ((GROUP * w + 31) & ~31); t += 32) {  // whole-warp trips: the shuffle below
                        const int rr = min(t / w, GROUP - 1), c = t - (t / w) * w;
                        const int gr = __shfl_sync(0xffffffffu, g, f * 8 + rr);  // group of the lane that owns row rr
                        const int64_t jr = (int64_t)gr * cg + (rr % cg);
                        if (t < GROUP * w && gr >= 0 && jr < k) dst[rr * SLAB_RS + c] = __ldg(Cn + jr * d + e0 + c);
                    }
                }
                __syncwarp();
                // ---- advance the partial sums over this slab
                if (j >= 0 && j < k) {
                    const float* cr = cb + sub * SLAB_RS;
                    const int w4 = min(w, d4 - e0) & ~3;  // full 4-blocks of the reference's main loop inside this slab
                    for (int e = 0; e < w4; e += 4) {
                        const float4 xv = *reinterpret_cast<const float4*>(xb + e);
                        const float4 cv = *reinterpret_cast<const float4*>(cr + e);
                        L.add4(xv.x, xv.y, xv.z, xv.w, cv.x, cv.y, cv.z, cv.w);
                    }
                    for (int e = max(w4, 0); e < w; ++e)
                        if (e0 + e >= d4) L.tail(xb[e], cr[e]);
                }
            }
            if (j >= 0 && j < k) am.offer(L.result(), j);
            if (g >= 0 && within == 0 && nc != 255) my_groups += 1;
        }
        if (live && nc == 255 && sub == 0) { fallback_push(prm, fb_list, i); my_fb += 1; }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const float so = __shfl_xor_sync(0xffffffffu, am.s, o);
            const int32_t jo = __shfl_xor_sync(0xffffffffu, am.j, o);
            am.merge(so, jo);
        }
        if (live && sub == 0 && nc != 255) {
            labels[i] = (lloyd && am.j < 0) ? 0 : am.j;
            if (mind) mind[i] = am.j >= 0 ? __fsqrt_rn(am.s) : 3.402823466e+38f;
        }
    }
    verify_stats(my_groups, my_fb, prm);
}

// d > 16, no staging: 8 lanes per frame (4 frames per warp); lane (group slot, center) streams ITS center row and the
// frame row straight from global memory / L2 with 16-byte loads (the frame row is the same address for the lanes
// of a frame -> one broadcast transaction) and advances its own Lanes4 sums.  A frame typically has one or two
// candidate groups, so this kernel is bound by load latency, not by bandwidth or arithmetic: no shared memory and
// ~64 registers keep 32+ warps per SM in flight, which is what the staged variants above lacked (2 CTAs of 8 warps
// per SM: measured 4.2 ms for 1.25e7 x 64 against ~1 ms of L2 traffic).
template <bool VEC>
__global__ void __launch_bounds__(256) screen_verify_direct_kernel(const float* __restrict__ X, int64_t n, int d,
                                                                   const float* __restrict__ Cn, int k,
                                                                   const uint32_t* __restrict__ cand,
                                                                   const uint8_t* __restrict__ ncand,
                                                                   int32_t* __restrict__ labels,
                                                                   float* __restrict__ mind, int lloyd,
                                                                   ScreenParams* prm, uint32_t* __restrict__ fb_list,
                                                                   int cg, const uint16_t* __restrict__ tlist /* per-tile
                                                                   center lists (candidate groups count list positions) or
                                                                   null */, int lcap, int unit_frames /* frames per list */) {
    if (!prm->valid) return;
    const int lane = threadIdx.x & 31;
    const int sub = lane & 7, slot = lane >> 3;
    const int idb = cand_id_bits(cg), ng = CHUNK / cg, S = GROUP / cg;
    const uint32_t idm = cand_id_mask(cg);
    const int gslot = sub / cg, within = sub - gslot * cg;
    const int d4 = d & ~3;
    unsigned long long my_groups = 0, my_fb = 0;
    const int64_t n_warps = ((int64_t)gridDim.x * 256) >> 5;
    const int64_t warp_global = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    for (int64_t base = warp_global * 4; base < n; base += n_warps * 4) {
        const int64_t i = base + slot;
        const bool live = i < n;
        int nc = live ? (int)ncand[i] : 0;
        const uint16_t* tl = tlist ? tlist + (size_t)((live ? i : 0) / unit_frames) * lcap : nullptr;
        if (nc == NCAND_DECIDED) {  // settled by the listed screen (tl is set): no evaluation
            const int j = (int)__ldg(tl + cand[i * CAND_CAP]);
            if (sub == 0) {
                if (j < k) labels[i] = j;
                else { fallback_push(prm, fb_list, i); my_fb += 1; }
            }
            nc = 0;
        }
        const bool skip_out = live && ncand[i] == NCAND_DECIDED;
        uint32_t ent = 0;
        if (live && nc != 255 && sub < nc) ent = cand[i * CAND_CAP + sub];
        const int pc = __popc(ent >> idb);
        int off = pc;  // inclusive scan over the 8 lanes of the frame
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, off, o, 8);
            if (sub >= o) off += v;
        }
        int total = __shfl_sync(0xffffffffu, off, 7, 8);
        off -= pc;
        if (nc == 255) total = 0;  // queued for the fallback kernel below
        int rounds = (total + S - 1) / S;
#pragma unroll
        for (int o = 8; o < 32; o <<= 1) rounds = max(rounds, __shfl_xor_sync(0xffffffffu, rounds, o));
        ArgMin am;
        am.init();
        const float* xr = X + (live ? i : 0) * d;
        for (int r = 0; r < rounds; ++r) {
            int g = -1;
            {
                const int gi = r * S + gslot;
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const uint32_t et = __shfl_sync(0xffffffffu, ent, t, 8);
                    const int ot = __shfl_sync(0xffffffffu, off, t, 8);
                    const int pt = __shfl_sync(0xffffffffu, pc, t, 8);
                    if (gi >= ot && gi < ot + pt) g = (int)(et & idm) * ng + nth_set_bit(et >> idb, gi - ot);
                }
                if (gi >= total) g = -1;
            }
            int j = g >= 0 ? g * cg + within : -1;
            if (tl && j >= 0) j = (int)__ldg(tl + j);  // list position -> center id (the padding id is >= k)
            if (j >= 0 && j < k) {
                const float* cr = Cn + (int64_t)j * d;
                Lanes4 L;
                L.init();
                if (VEC) {
                    const float4* x4 = reinterpret_cast<const float4*>(xr);
                    const float4* c4 = reinterpret_cast<const float4*>(cr);
                    const int nv = d >> 2;
                    int t = 0;
                    for (; t + 8 <= nv; t += 8) {  // 16 loads of 16 bytes in flight per lane: the kernel is L2-latency bound
                        float4 xv[8], cv[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) { xv[u] = __ldg(x4 + t + u); cv[u] = __ldg(c4 + t + u); }
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            L.add4(xv[u].x, xv[u].y, xv[u].z, xv[u].w, cv[u].x, cv[u].y, cv[u].z, cv[u].w);
                    }
                    for (; t + 4 <= nv; t += 4) {
                        float4 xv[4], cv[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) { xv[u] = __ldg(x4 + t + u); cv[u] = __ldg(c4 + t + u); }
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            L.add4(xv[u].x, xv[u].y, xv[u].z, xv[u].w, cv[u].x, cv[u].y, cv[u].z, cv[u].w);
                    }
                    for (; t < nv; ++t) {
                        const float4 xv = __ldg(x4 + t), cv = __ldg(c4 + t);
                        L.add4(xv.x, xv.y, xv.z, xv.w, cv.x, cv.y, cv.z, cv.w);
                    }
                } else {
                    for (int e = 0; e < d4; e += 4)
                        L.add4(__ldg(xr + e), __ldg(xr + e + 1), __ldg(xr + e + 2), __ldg(xr + e + 3), __ldg(cr + e),
                               __ldg(cr + e + 1), __ldg(cr + e + 2), __ldg(cr + e + 3));
                    for (int e = d4; e < d; ++e) L.tail(__ldg(xr + e), __ldg(cr + e));
                }
                am.offer(L.result(), j);
            }
            if (g >= 0 && within == 0 && nc != 255) my_groups += 1;
        }
        if (live && nc == 255 && sub == 0) { fallback_push(prm, fb_list, i); my_fb += 1; }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const float so = __shfl_xor_sync(0xffffffffu, am.s, o);
            const int32_t jo = __shfl_xor_sync(0xffffffffu, am.j, o);
            am.merge(so, jo);
        }
        if (live && sub == 0 && nc != 255 && !skip_out) {
            labels[i] = (lloyd && am.j < 0) ? 0 : am.j;
            if (mind) mind[i] = am.j >= 0 ? __fsqrt_rn(am.s) : 3.402823466e+38f;
        }
    }
    verify_stats(my_groups, my_fb, prm);
}

// ---- plan -------------------------------------------------------------------------------------------------------------
// shared-memory plan of the screen kernel for a (k_pad, Kp) operand pair
struct GemmSmemPlan {
    int resident, n_stages, stage_bytes, bres_bytes;
    size_t total;
};
static GemmSmemPlan gemm_smem_plan(int k_pad, int Kp, size_t smem_optin, int allow_resident_a = 1) {
    GemmSmemPlan sp;
    const size_t tail = sizeof(GemmSmemTail) + 1024 /* alignment slack */;
    const size_t b_all = (size_t)k_pad * Kp * 2;            // multiple of B_BYTES
    const size_t a_full = (size_t)(Kp / BLOCK_K) * A_BYTES;  // one frame tile, full K
    if (b_all + 2 * a_full + tail <= smem_optin) {
        sp.resident = 1;
        sp.bres_bytes = (int)b_all;
        sp.stage_bytes = (int)a_full;
        sp.n_stages = (int)std::min<size_t>(MAX_STAGES, (smem_optin - tail - b_all) / a_full);
        if (sp.n_stages > 4) sp.n_stages = 4;
    } else if (allow_resident_a && Kp / BLOCK_K <= MAX_A_KBLOCKS && k_pad / TILE_N >= 2 &&
               a_full + 2 * (size_t)B_BYTES + tail <= smem_optin) {
        sp.resident = 2;  // frame tile resident, center k-blocks stream
        sp.bres_bytes = (int)a_full;
        sp.stage_bytes = B_BYTES;
        sp.n_stages = (int)std::min<size_t>(MAX_STAGES, (smem_optin - tail - a_full) / B_BYTES);
    } else {
        sp.resident = 0;
        sp.bres_bytes = 0;
        sp.stage_bytes = STAGE_BYTES;
        sp.n_stages = (int)std::min<size_t>(MAX_STAGES, (smem_optin - tail) / STAGE_BYTES);
    }
    sp.total = tail + (size_t)sp.bres_bytes + (size_t)sp.n_stages * sp.stage_bytes;
    return sp;
}

bool screen_supported(const b2k_ctx* ctx, int d, int k, int64_t n) {
    if (ctx->engine == B2K_ENGINE_DIRECT) return false;
    if (d < 1 || 3 * d + 3 > 2048 || k < 2 || k > (1 << 27)) return false;
    if (sizeof(GemmSmemTail) + 1024 + 2 * (size_t)STAGE_BYTES > ctx->smem_optin) return false;
    if (ctx->engine == B2K_ENGINE_SCREEN) return true;
    // auto: the screen pays off once a frame meets enough center coordinates
    return k >= 128 && (int64_t)k * d >= 2048 && n >= 4096;
}

// operand terms (MMA K = terms*d + 3): 3 = hi/lo fp16 split of frames and centers, 2 = split frames against hi-only
// centers, 1 = hi-only both.  Fewer terms issue fewer MMA flops and leave a wider margin, i.e. more candidates for the exact
// verify; option screen_terms forces 1/2/3, 0 lets screen_choose_terms measure the candidate counts on a sample.
static int screen_default_terms(const b2k_ctx* ctx) {
    return (ctx->screen_terms >= 1 && ctx->screen_terms <= 3) ? ctx->screen_terms : 3;
}

void screen_plan_destroy(ScreenPlan* p) {
    if (!p) return;
    cudaStreamSynchronize(p->ctx->stream);
    dev_free(p->A); dev_free(p->B); dev_free(p->X2); dev_free(p->XL); dev_free(p->mu); dev_free(p->params); dev_free(p->cand);
    dev_free(p->ncand); dev_free(p->fb_list);
    delete p;
}

int screen_plan_create(b2k_ctx* ctx, int64_t n_cap, int d, int k, int terms, ScreenPlan** out) {
    ScreenPlan* p = new ScreenPlan();
    p->ctx = ctx;
    p->n_cap = n_cap;
    p->n_pad = cdiv(n_cap, TILE_M) * TILE_M;
    p->d = d;
    p->k = k;
    p->k_pad = (int)(cdiv(k, TILE_N) * TILE_N);
    p->terms = (terms >= 1 && terms <= 3) ? terms : screen_default_terms(ctx);
    // candidate group size: narrow rows (d <= 16) leave the TMEM-read-bound epilogue little slack (measured at 1e7 x 10,
    // k=1000, all centers: screen kernel 2.32 / 2.44 / 2.90 ms for groups of 8 / 4 / 2, step time 3.40 / 3.36 / 3.75 ms;
    // pruned session: screen 0.752 / 0.764 / 0.792, verify 0.342 / 0.283 / 0.252, step 1.398 / 1.351 / 1.355 ms) -> 4; wide
    // rows are MMA bound and their verify pays 4*d bytes of L2 traffic per candidate center -> 2 (16-bit chunk ids)
    p->cg = (ctx->screen_group == 8 || ctx->screen_group == 4 || ctx->screen_group == 2) ? ctx->screen_group
                                                                                          : (d > 16 ? 2 : 4);
    if (p->cg == 2 && p->k_pad / CHUNK > (1 << 16)) p->cg = 4;
    p->Kc = p->terms * d + 3;
    p->Kp = (int)(cdiv(p->Kc, BLOCK_K) * BLOCK_K);
    p->nk16 = (int)cdiv(p->Kc, 16);
    cudaError_t e = dev_alloc(&p->A, (size_t)p->n_pad * p->Kp * 2);
    p->k_rows = p->k_pad + 8;
    if (e == cudaSuccess) e = dev_alloc(&p->B, (size_t)p->k_rows * p->Kp * 2);
    if (e == cudaSuccess) e = dev_alloc(&p->X2, (size_t)p->n_pad * 4);
    if (e == cudaSuccess && p->terms != 3) e = dev_alloc(&p->XL, (size_t)p->n_pad * 4);
    if (e == cudaSuccess) e = dev_alloc(&p->mu, (size_t)d * 4);
    if (e == cudaSuccess) e = dev_alloc(&p->params, sizeof(ScreenParams));
    if (e == cudaSuccess) e = dev_alloc(&p->cand, (size_t)p->n_pad * CAND_CAP * 4);
    if (e == cudaSuccess) e = dev_alloc(&p->ncand, (size_t)p->n_pad);
    if (e == cudaSuccess) e = dev_alloc(&p->fb_list, (size_t)p->n_pad * 4);
    if (e != cudaSuccess) {
        cudaGetLastError();
        screen_plan_destroy(p);
        return set_error(B2K_ERR_NOMEM, "screen plan: %s", cudaGetErrorString(e));
    }
    int rc = make_tmap(&p->tmA, p->A, (uint64_t)p->n_pad, (uint64_t)p->Kp, TILE_M);
    if (rc == B2K_OK) rc = make_tmap(&p->tmB, p->B, (uint64_t)p->k_rows, (uint64_t)p->Kp, TILE_N);
    if (rc == B2K_OK) rc = make_tmap(&p->tmBh, p->B, (uint64_t)p->k_rows, (uint64_t)p->Kp, TILE_N / 2);
    if (rc == B2K_OK) rc = make_tmap(&p->tmBg, p->B, (uint64_t)p->k_rows, (uint64_t)p->Kp, 1);
    if (rc != B2K_OK) { screen_plan_destroy(p); return rc; }
    static PerDeviceOnce attr_set;
    if (attr_set.need(ctx->device)) {
        cudaError_t ae = cudaFuncSetAttribute(screen_gemm_kernel<8, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)ctx->smem_optin);
        if (ae == cudaSuccess)
            ae = cudaFuncSetAttribute(screen_gemm_kernel<4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin);
        if (ae == cudaSuccess)
            ae = cudaFuncSetAttribute(screen_gemm_kernel<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin);
        if (ae == cudaSuccess)
            ae = cudaFuncSetAttribute(screen_gemm_kernel<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin);
        if (ae == cudaSuccess)
            ae = cudaFuncSetAttribute(screen_gemm_kernel<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin);
        if (ae == cudaSuccess)
            ae = cudaFuncSetAttribute(screen_gemm_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin);
        if (ae == cudaSuccess)
            ae = cudaFuncSetAttribute(screen_gemm_kernel<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin);
        if (ae == cudaSuccess)
            ae = cudaFuncSetAttribute(screen_gemm_kernel<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin);
        if (ae == cudaSuccess)
            ae = cudaFuncSetAttribute(screen_gemm_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin);
        if (ae != cudaSuccess) { screen_plan_destroy(p); return set_error(B2K_ERR_CUDA, "screen smem attr: %s", cudaGetErrorString(ae)); }
        attr_set.done(ctx->device);
    }
    *out = p;
    return B2K_OK;
}

int screen_plan_acquire(b2k_ctx* ctx, int64_t n, int d, int k, int terms, ScreenPlan** out) {
    ScreenPlan* c = static_cast<ScreenPlan*>(ctx->assign_plan);
    if (terms < 1 || terms > 3) terms = screen_default_terms(ctx);
    const int want_cg = (ctx->screen_group == 8 || ctx->screen_group == 4 || ctx->screen_group == 2) ? ctx->screen_group : 0;
    if (c && c->d == d && c->k == k && c->terms == terms && n <= c->n_cap && (want_cg == 0 || want_cg == c->cg)) {
        c->prepared_n = -1;
        *out = c;
        return B2K_OK;
    }
    screen_plan_release_cached(ctx);
    B2K_TRY(screen_plan_create(ctx, n, d, k, terms, &c));
    ctx->assign_plan = c;
    *out = c;
    return B2K_OK;
}

void screen_plan_release_cached(b2k_ctx* ctx) {
    if (ctx->assign_plan) screen_plan_destroy(static_cast<ScreenPlan*>(ctx->assign_plan));
    ctx->assign_plan = nullptr;
}

static unsigned capped_grid(b2k_ctx* ctx, int64_t items, int per_block) {
    int64_t b = cdiv(items, per_block);
    const int64_t cap = (int64_t)ctx->sm_count * 16;
    return (unsigned)std::max<int64_t>(1, std::min(b, cap));
}

// (re)build the frame operand: mu from the given centers, scale from the data, A' and |x~|^2
int screen_prepare_frames_with_centers(ScreenPlan* p, const float* dX, int64_t n, const float* dC) {
    b2k_ctx* ctx = p->ctx;
    cudaStream_t st = ctx->stream;
    if (n > p->n_cap) return set_error(B2K_ERR_INVALID_ARG, "screen plan too small");
    CUDA_TRY(cudaMemsetAsync(p->params, 0, sizeof(ScreenParams), st));
    screen_mu_kernel<<<(unsigned)p->d, 256, 0, st>>>(dC, p->k, p->d, p->mu);
    LAUNCH_CHECK();
    if (p->d <= 16) {
        screen_maxnorm_kernel<1><<<capped_grid(ctx, n, 256), 256, 0, st>>>(dX, n, p->d, p->mu, &p->params->xmax2_raw);
        LAUNCH_CHECK();
        screen_maxnorm_kernel<1><<<capped_grid(ctx, p->k, 256), 256, 0, st>>>(dC, p->k, p->d, p->mu, &p->params->cmax2_raw);
        LAUNCH_CHECK();
    } else if (p->d <= 128) {
        screen_maxnorm_kernel<8><<<capped_grid(ctx, n, 32), 256, 0, st>>>(dX, n, p->d, p->mu, &p->params->xmax2_raw);
        LAUNCH_CHECK();
        screen_maxnorm_kernel<8><<<capped_grid(ctx, p->k, 32), 256, 0, st>>>(dC, p->k, p->d, p->mu, &p->params->cmax2_raw);
        LAUNCH_CHECK();
    } else {
        screen_maxnorm_kernel<32><<<capped_grid(ctx, n, 8), 256, 0, st>>>(dX, n, p->d, p->mu, &p->params->xmax2_raw);
        LAUNCH_CHECK();
        screen_maxnorm_kernel<32><<<capped_grid(ctx, p->k, 8), 256, 0, st>>>(dC, p->k, p->d, p->mu, &p->params->cmax2_raw);
        LAUNCH_CHECK();
    }
    screen_sigma_kernel<<<1, 1, 0, st>>>(p->params);
    LAUNCH_CHECK();
    const int64_t n_pad_now = cdiv(n, TILE_M) * TILE_M;
    if (ctx->operand_kernel == 1) {
        screen_frames_kernel<<<capped_grid(ctx, n_pad_now, 256 / (p->Kp / 8)), 256, 0, st>>>(
            dX, n, n_pad_now, p->d, p->terms, p->Kp, p->mu, p->params, reinterpret_cast<uint4*>(p->A));
    } else {
        // rows per tile: 32 KB of shared memory per CTA -> several CTAs per SM
        int R = (int)std::max<int64_t>(1, std::min<int64_t>(256, (32 * 1024) / ((int64_t)p->Kp * 2)));
        const size_t tsm = (size_t)R * p->Kp * 2;
        const unsigned tg = (unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv(n_pad_now, R), (int64_t)ctx->sm_count * 6));
        screen_frames_tile_kernel<<<tg, 256, tsm, st>>>(dX, n, n_pad_now, p->d, p->terms, p->Kp, R, p->mu, p->params,
                                                        reinterpret_cast<uint4*>(p->A));
    }
    LAUNCH_CHECK();
    if (p->d <= 16)
        screen_x2_kernel<1><<<(unsigned)cdiv(n_pad_now, 256), 256, 0, st>>>(dX, n, n_pad_now, p->d, p->mu, p->params, p->X2, p->XL);
    else if (p->d <= 128)
        screen_x2_kernel<8><<<(unsigned)cdiv(n_pad_now, 32), 256, 0, st>>>(dX, n, n_pad_now, p->d, p->mu, p->params, p->X2, p->XL);
    else
        screen_x2_kernel<32><<<(unsigned)cdiv(n_pad_now, 8), 256, 0, st>>>(dX, n, n_pad_now, p->d, p->mu, p->params, p->X2, p->XL);
    LAUNCH_CHECK();
    p->prepared_n = n;
    return B2K_OK;
}

int screen_prepare_frames(ScreenPlan* p, const float* dX, int64_t n) {
    // the centers are not known yet: defer to the first screen_assign (which passes them)
    (void)dX;
    if (n > p->n_cap) return set_error(B2K_ERR_INVALID_ARG, "screen plan too small");
    p->prepared_n = -1;
    return B2K_OK;
}

// tail of every screen_assign: queued frames -> CTA-per-frame exact scan; unusable operands (valid == 0: the
// screen and verify kernels returned at once) -> the exact tile kernel over all frames, gated on the device flag
static int screen_finish_assign(ScreenPlan* p, const float* dX, int64_t n, const float* dC, int32_t* labels,
                                float* mind, int lloyd) {
    b2k_ctx* ctx = p->ctx;
    // queued frames: a handful -> CTA per frame, all 256 threads over the centers (the register-tiled kernel would put one
    // CTA on a 128-frame tile that holds one frame: 1.1 ms per step at k=5000, d=256); a long queue -> register-tiled exact
    // tile kernel (the CTA-per-frame scan needed 36.7 ms for 8.4e3 frames at cfg4).  Both are launched, the device-side
    // count picks one.
    const unsigned int split = ctx->fallback_mode == 1 ? 0xffffffffu : (ctx->fallback_mode == 2 ? 0u : 256u);
    if (split > 0) {
        const size_t fsmem = ((size_t)((p->d + 3) & ~3) + 512) * 4;
        screen_fallback_kernel<<<ctx->sm_count * 2, 256, fsmem, ctx->stream>>>(dX, p->d, dC, p->k, p->fb_list, p->params,
                                                                              labels, mind, lloyd, split);
        LAUNCH_CHECK();
    }
    if (split != 0xffffffffu)
        B2K_TRY(launch_tile_indexed(ctx, dX, p->d, dC, p->k, p->fb_list, &p->params->fb_count, &p->params->valid, labels,
                                    mind, lloyd, split));
    return launch_assign_exact_if(ctx, dX, n, p->d, dC, p->k, labels, mind, lloyd, &p->params->valid);
}

int screen_assign(ScreenPlan* p, const float* dX, int64_t n, const float* dC, int32_t* labels, float* mind,
                  int lloyd) {
    b2k_ctx* ctx = p->ctx;
    cudaStream_t st = ctx->stream;
    if (p->prepared_n != n) B2K_TRY(screen_prepare_frames_with_centers(p, dX, n, dC));
    // center operand for the current centers
    CUDA_TRY(cudaMemsetAsync(&p->params->cmax2_now, 0, 8, st));  // cmax2_now, cl2_now
    CUDA_TRY(cudaMemsetAsync(&p->params->cand_chunks, 0, 24, st));
    screen_centers_kernel<<<(unsigned)cdiv(p->k_rows, 128), 128, 0, st>>>(dC, p->k, p->k_rows, p->d, p->terms, p->Kp,
                                                                          p->mu, p->params, p->B);
    LAUNCH_CHECK();
    screen_finish_centers_kernel<<<1, 1, 0, st>>>(p->params, p->d);
    LAUNCH_CHECK();
    GemmArgs g;
    g.n = n;
    g.n_tiles = (int)cdiv(n, TILE_M);
    g.n_ntiles = p->k_pad / TILE_N;
    g.n_kblocks = p->Kp / BLOCK_K;
    g.nk16 = p->nk16;
    g.d = p->d;
    g.terms = p->terms;
    g.X2 = p->X2;
    g.XL = p->XL;
    g.prm = p->params;
    g.cand = p->cand;
    g.ncand = p->ncand;
    GemmSmemPlan sp = gemm_smem_plan(p->k_pad, p->Kp, ctx->smem_optin, ctx->screen_resident_a);
    g.cluster2 = 0;
    if (sp.resident == 0 && g.n_tiles >= 2 && ctx->sm_count >= 2) {
        if (ctx->screen_cluster == 2) g.cluster2 = 1;
        else if (ctx->screen_cluster == 3) {
            // CTA-pair MMAs: a ring slot holds my frame k-block and my half of the center k-block
            g.cluster2 = 2;
            const size_t tail = sizeof(GemmSmemTail) + 1024;
            sp.stage_bytes = A_BYTES + B_BYTES / 2;
            sp.n_stages = (int)std::min<size_t>(MAX_STAGES, (ctx->smem_optin - tail) / sp.stage_bytes);
            sp.total = tail + (size_t)sp.n_stages * sp.stage_bytes;
        }
    }
    g.resident = sp.resident;
    g.n_stages = sp.n_stages;
    g.stage_bytes = sp.stage_bytes;
    g.bres_bytes = sp.bres_bytes;
    const unsigned grid = (unsigned)std::min<int64_t>(g.n_tiles, ctx->sm_count);
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (ctx->profile) {
        CUDA_TRY(cudaEventCreate(&ev0));
        CUDA_TRY(cudaEventCreate(&ev1));
        CUDA_TRY(cudaEventRecord(ev0, st));
    }
    if (g.cluster2) {
        // pairs of CTAs share every center k-block (each fetches half and multicasts it): an even grid of 2-CTA clusters
        cudaLaunchConfig_t cfg = {};
        const unsigned pairs = (unsigned)std::min<int64_t>((g.n_tiles + 1) / 2, ctx->sm_count / 2);
        cfg.gridDim = dim3(2 * pairs);
        cfg.blockDim = dim3(GEMM_THREADS);
        cfg.dynamicSmemBytes = sp.total;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t le;
        if (g.cluster2 == 2) {
            if (p->cg == 8) le = cudaLaunchKernelEx(&cfg, screen_gemm_kernel<8, 2>, p->tmA, p->tmB, p->tmBh, g);
            else if (p->cg == 4) le = cudaLaunchKernelEx(&cfg, screen_gemm_kernel<4, 2>, p->tmA, p->tmB, p->tmBh, g);
            else le = cudaLaunchKernelEx(&cfg, screen_gemm_kernel<2, 2>, p->tmA, p->tmB, p->tmBh, g);
        } else if (p->cg == 8) le = cudaLaunchKernelEx(&cfg, screen_gemm_kernel<8, 1>, p->tmA, p->tmB, p->tmBh, g);
        else if (p->cg == 4) le = cudaLaunchKernelEx(&cfg, screen_gemm_kernel<4, 1>, p->tmA, p->tmB, p->tmBh, g);
        else le = cudaLaunchKernelEx(&cfg, screen_gemm_kernel<2, 1>, p->tmA, p->tmB, p->tmBh, g);
        CUDA_TRY(le);
    } else if (sp.resident == 2) {
        if (p->cg == 8) screen_gemm_kernel<8, 1><<<grid, GEMM_THREADS, sp.total, st>>>(p->tmA, p->tmB, p->tmBh, g);
        else if (p->cg == 4) screen_gemm_kernel<4, 1><<<grid, GEMM_THREADS, sp.total, st>>>(p->tmA, p->tmB, p->tmBh, g);
        else screen_gemm_kernel<2, 1><<<grid, GEMM_THREADS, sp.total, st>>>(p->tmA, p->tmB, p->tmBh, g);
    } else if (p->cg == 8) screen_gemm_kernel<8, 0><<<grid, GEMM_THREADS, sp.total, st>>>(p->tmA, p->tmB, p->tmBh, g);
    else if (p->cg == 4) screen_gemm_kernel<4, 0><<<grid, GEMM_THREADS, sp.total, st>>>(p->tmA, p->tmB, p->tmBh, g);
    else screen_gemm_kernel<2, 0><<<grid, GEMM_THREADS, sp.total, st>>>(p->tmA, p->tmB, p->tmBh, g);
    LAUNCH_CHECK();
    if (ctx->profile) {
        CUDA_TRY(cudaEventRecord(ev1, st));
        ctx->prof_events.push_back(ev0);
        ctx->prof_events.push_back(ev1);
    }
    // verify (persistent grids)
    ProfScope prof_verify(ctx, b2k_ctx::PROF_VERIFY);
    if (p->d <= 16) {
        const int ds = (p->d + 3) & ~3;
        {   // table kernel when the skewed table fits shared memory twice per SM
            const int gstride = GROUP * ds + (((GROUP * ds / 4) & 1) ? 0 : 4);  // odd number of 16-byte units
            const size_t tbytes = (size_t)cdiv(p->k, GROUP) * gstride * 4;
            if (tbytes <= 100 * 1024 && ctx->verify_mode != 3) {
                const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (200 * 1024) / tbytes));
                const unsigned tgrid =
                    (unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv(n, 256), (int64_t)ctx->sm_count * per_sm));
#define B2K_VTABLE(DR)                                                                                               \
    do {                                                                                                             \
        static PerDeviceOnce vattr;                                                                                   \
        if (vattr.need(ctx->device)) {                                                                                                \
            CUDA_TRY(cudaFuncSetAttribute(screen_verify_table_kernel<DR, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          100 * 1024));                                                              \
            CUDA_TRY(cudaFuncSetAttribute(screen_verify_table_kernel<DR, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          100 * 1024));                                                              \
            vattr.done(ctx->device);                                                                                            \
        }                                                                                                            \
        if (p->cg == 8 && ctx->verify_mode != 2)                                                                     \
            screen_verify_table_kernel<DR, 8><<<tgrid, 256, tbytes, st>>>(dX, n, p->d, dC, p->k, p->cand, p->ncand, labels, \
                                                                   mind, lloyd, p->params, p->fb_list, gstride,     \
                                                                   p->cg, row_load_width(dX, p->d, ctx->row_vec_max));                                           \
        else                                                                                                         \
            screen_verify_table_kernel<DR, 0><<<tgrid, 256, tbytes, st>>>(dX, n, p->d, dC, p->k, p->cand, p->ncand, labels, \
                                                                   mind, lloyd, p->params, p->fb_list, gstride,     \
                                                                   p->cg, row_load_width(dX, p->d, ctx->row_vec_max));                                           \
    } while (0)
                if (ds == 4) B2K_VTABLE(4);
                else if (ds == 8) B2K_VTABLE(8);
                else if (ds == 12) B2K_VTABLE(12);
                else B2K_VTABLE(16);
#undef B2K_VTABLE
                LAUNCH_CHECK();
                return screen_finish_assign(p, dX, n, dC, labels, mind, lloyd);
            }
        }
        const int rs = (ds % 8 == 4) ? ds : ds + 4;
        const size_t tab_bytes = (size_t)p->k * rs * 4;
        const int use_smem = (ctx->verify_mode == 3 && tab_bytes <= 100 * 1024) ? 1 : 0;  // 8 lanes per frame, table in smem
        const size_t vsmem = use_smem ? tab_bytes : 0;
        const int per_sm = use_smem ? (int)std::max<size_t>(1, std::min<size_t>(4, (200 * 1024) / std::max<size_t>(tab_bytes, 1))) : 4;
        const unsigned vgrid =
            (unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv(n, 32), (int64_t)ctx->sm_count * per_sm));
#define B2K_VERIFY(DR)                                                                                               \
    do {                                                                                                             \
        static PerDeviceOnce vattr;                                                                                   \
        if (vattr.need(ctx->device)) {                                                                                                \
            CUDA_TRY(cudaFuncSetAttribute(screen_verify_small_kernel<DR>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          100 * 1024));                                                              \
            vattr.done(ctx->device);                                                                                            \
        }                                                                                                            \
        screen_verify_small_kernel<DR><<<vgrid, 256, vsmem, st>>>(dX, n, p->d, dC, p->k, p->cand, p->ncand, labels,  \
                                                                  mind, lloyd, p->params, p->fb_list, use_smem, rs, \
                                                                  p->cg);                                            \
    } while (0)
        if (p->d <= 4) B2K_VERIFY(4);
        else if (p->d <= 8) B2K_VERIFY(8);
        else if (p->d <= 12) B2K_VERIFY(12);
        else B2K_VERIFY(16);
#undef B2K_VERIFY
    } else if (ctx->verify_mode != 1) {
        const bool vec = p->d % 4 == 0 && (((uintptr_t)dX | (uintptr_t)dC) & 15) == 0;
        const unsigned dgrid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv(n, 32), (int64_t)ctx->sm_count * 8));
        if (vec)
            screen_verify_direct_kernel<true><<<dgrid, 256, 0, st>>>(dX, n, p->d, dC, p->k, p->cand, p->ncand, labels, mind,
                                                                     lloyd, p->params, p->fb_list, p->cg, nullptr, 0, TILE_M);
        else
            screen_verify_direct_kernel<false><<<dgrid, 256, 0, st>>>(dX, n, p->d, dC, p->k, p->cand, p->ncand, labels, mind,
                                                                      lloyd, p->params, p->fb_list, p->cg, nullptr, 0, TILE_M);
    } else {
        if (p->d % 4 == 0 && (((uintptr_t)dX | (uintptr_t)dC) & 15) == 0) {
            const size_t tsmem = (size_t)8 * VC_WARP_FLOATS * 4;
            static PerDeviceOnce tattr;
            if (tattr.need(ctx->device)) {
                CUDA_TRY(cudaFuncSetAttribute(screen_verify_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
                tattr.done(ctx->device);
            }
            const int per_sm = (int)std::max<size_t>(1, (220 * 1024) / (tsmem + 1024));
            const unsigned tgrid =
                (unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv(n, 32), (int64_t)ctx->sm_count * per_sm));
            screen_verify_stream_kernel<<<tgrid, 256, tsmem, st>>>(dX, n, p->d, dC, p->k, p->cand, p->ncand, labels, mind,
                                                                   lloyd, p->params, p->fb_list, p->cg);
            LAUNCH_CHECK();
            return screen_finish_assign(p, dX, n, dC, labels, mind, lloyd);
        }
        const size_t vsmem = (size_t)8 * VW_FLOATS * 4;
        const int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / vsmem));
        static PerDeviceOnce vattr;
        if (vattr.need(ctx->device)) {
            CUDA_TRY(cudaFuncSetAttribute(screen_verify_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            vattr.done(ctx->device);
        }
        const unsigned vgrid =
            (unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv(n, 32), (int64_t)ctx->sm_count * ctas_per_sm));
        screen_verify_wide_kernel<<<vgrid, 256, vsmem, st>>>(dX, n, p->d, dC, p->k, p->cand, p->ncand, labels, mind, lloyd,
                                                             p->params, p->fb_list, p->cg);
    }
    LAUNCH_CHECK();
    return screen_finish_assign(p, dX, n, dC, labels, mind, lloyd);
}

// Term count for a data set: forced by option screen_terms, 3 for narrow rows / small jobs (their screen kernel is not
// MMA bound), otherwise MEASURED: the screen + verify run on a sample of the frames (four contiguous blocks spread over the
// array) with 1 and then 2 terms, and the first count whose candidate lists stay short is taken.  Labels are exact with any
// count; only the split of the work between the tensor pipe and the exact verify changes.
int screen_choose_terms(b2k_ctx* ctx, const float* dX, int64_t n, int d, const float* dC, int k, int* terms_out) {
    *terms_out = screen_default_terms(ctx);
    if (ctx->screen_terms >= 1 && ctx->screen_terms <= 3) return B2K_OK;
    // the probe costs a few milliseconds: only jobs whose 3-term screen takes ~10 ms or more are worth it
    if (d < 32 || n < 65536 || 2.0 * (double)n * k * d < 1e9 * (double)ctx->probe_min_gflop || !screen_supported(ctx, d, k, n))
        return B2K_OK;
    const int64_t blk = 8192, nblk = 4, m = blk * nblk;
    DevMem xs, ls;
    if (xs.alloc((size_t)m * d * 4) != B2K_OK || ls.alloc((size_t)m * 4) != B2K_OK) { cudaGetLastError(); return B2K_OK; }
    for (int64_t b = 0; b < nblk; ++b) {
        const int64_t off = (n - blk) * b / (nblk - 1);
        CUDA_TRY(cudaMemcpyAsync(xs.as<float>() + b * blk * d, dX + off * d, (size_t)blk * d * 4, cudaMemcpyDeviceToDevice,
                                 ctx->stream));
    }
    const int cg = (ctx->screen_group == 8 || ctx->screen_group == 4 || ctx->screen_group == 2) ? ctx->screen_group : 2;
    for (int t = 1; t <= 2; ++t) {
        ScreenPlan* tp = nullptr;
        if (screen_plan_create(ctx, m, d, k, t, &tp) != B2K_OK) { cudaGetLastError(); break; }
        int rc = screen_assign(tp, xs.as<float>(), m, dC, ls.as<int32_t>(), nullptr, 0);
        double groups = 0, fb = 0;
        if (rc == B2K_OK) rc = screen_read_stats(tp, &groups, &fb);
        screen_plan_destroy(tp);
        if (rc != B2K_OK) return rc;
        ctx->stat_probe_centers[t] = groups * cg / (double)m;
        ctx->stat_probe_fallback[t] = fb / (double)m;
        // a candidate center costs the verify ~4d bytes of L2 traffic, a dropped term saves k*d MMA flops per frame
        if (fb <= 0.002 * m && groups * cg <= (double)ctx->probe_max_centers * m) { *terms_out = t; break; }
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return B2K_OK;
}


// Listed screen (prune.cu): the plan's frames are the session's SORTED frames; tile t meets the centers tlist[t][0 .. tcount[t]).
// Labels come out in the frames' (sorted) order.
int screen_assign_listed(ScreenPlan* p, const float* dX, int64_t n, const float* dC, const uint16_t* tlist,
                         const uint32_t* tcount, int lcap, int unit_shift, int32_t* labels, int lloyd) {
    b2k_ctx* ctx = p->ctx;
    cudaStream_t st = ctx->stream;
    if (p->prepared_n != n) B2K_TRY(screen_prepare_frames_with_centers(p, dX, n, dC));
    CUDA_TRY(cudaMemsetAsync(&p->params->cmax2_now, 0, 8, st));
    CUDA_TRY(cudaMemsetAsync(&p->params->cand_chunks, 0, 24, st));
    screen_centers_kernel<<<(unsigned)cdiv(p->k_rows, 128), 128, 0, st>>>(dC, p->k, p->k_rows, p->d, p->terms, p->Kp,
                                                                          p->mu, p->params, p->B);
    LAUNCH_CHECK();
    screen_finish_centers_kernel<<<1, 1, 0, st>>>(p->params, p->d);
    LAUNCH_CHECK();
    GemmArgs g;
    g.n = n;
    g.n_tiles = (int)cdiv(n, TILE_M);
    g.n_ntiles = 0;
    g.n_kblocks = p->Kp / BLOCK_K;
    g.nk16 = p->nk16;
    g.d = p->d;
    g.terms = p->terms;
    g.X2 = p->X2;
    g.XL = p->XL;
    g.prm = p->params;
    g.cand = p->cand;
    g.ncand = p->ncand;
    g.resident = 0;
    g.cluster2 = 0;
    g.bres_bytes = 0;
    g.stage_bytes = STAGE_BYTES;
    const size_t tail = sizeof(GemmSmemTail) + 1024;
    g.n_stages = (int)std::min<size_t>(MAX_STAGES, (ctx->smem_optin - tail) / STAGE_BYTES);
    const size_t smem = tail + (size_t)g.n_stages * STAGE_BYTES;
    ListArgs la;
    la.tlist = tlist;
    la.tcount = tcount;
    la.lcap = lcap;
    la.B = p->B;
    la.Kp = p->Kp;
    la.ushift = unit_shift;
    la.decide = ctx->screen_decide;
    la.gather = ctx->screen_gather;
    static PerDeviceOnce attr_set;
    if (attr_set.need(ctx->device)) {
        CUDA_TRY(cudaFuncSetAttribute(screen_gemm_listed_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
        CUDA_TRY(cudaFuncSetAttribute(screen_gemm_listed_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
        CUDA_TRY(cudaFuncSetAttribute(screen_gemm_listed_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
        CUDA_TRY(cudaFuncSetAttribute(screen_gemm_listed_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
        CUDA_TRY(cudaFuncSetAttribute(screen_gemm_listed_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
        CUDA_TRY(cudaFuncSetAttribute(screen_gemm_listed_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
        attr_set.done(ctx->device);
    }
    const unsigned grid = (unsigned)std::min<int64_t>(g.n_tiles, ctx->sm_count);
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (ctx->profile) {
        CUDA_TRY(cudaEventCreate(&ev0));
        CUDA_TRY(cudaEventCreate(&ev1));
        CUDA_TRY(cudaEventRecord(ev0, st));
    }
    if (ctx->screen_decide) {
        if (p->cg == 8) screen_gemm_listed_kernel<8, true><<<grid, LISTED_THREADS, smem, st>>>(p->tmA, p->tmBg, g, la);
        else if (p->cg == 4) screen_gemm_listed_kernel<4, true><<<grid, LISTED_THREADS, smem, st>>>(p->tmA, p->tmBg, g, la);
        else screen_gemm_listed_kernel<2, true><<<grid, LISTED_THREADS, smem, st>>>(p->tmA, p->tmBg, g, la);
    } else if (p->cg == 8) screen_gemm_listed_kernel<8, false><<<grid, LISTED_THREADS, smem, st>>>(p->tmA, p->tmBg, g, la);
    else if (p->cg == 4) screen_gemm_listed_kernel<4, false><<<grid, LISTED_THREADS, smem, st>>>(p->tmA, p->tmBg, g, la);
    else screen_gemm_listed_kernel<2, false><<<grid, LISTED_THREADS, smem, st>>>(p->tmA, p->tmBg, g, la);
    LAUNCH_CHECK();
    if (ctx->profile) {
        CUDA_TRY(cudaEventRecord(ev1, st));
        ctx->prof_events.push_back(ev0);
        ctx->prof_events.push_back(ev1);
    }
    ProfScope prof_verify(ctx, b2k_ctx::PROF_VERIFY);
    const int ds = (p->d + 3) & ~3;
    const int gstride = GROUP * ds + (((GROUP * ds / 4) & 1) ? 0 : 4);
    const size_t tbytes = (size_t)cdiv(p->k, GROUP) * gstride * 4;
    if (p->d <= 16 && tbytes <= 100 * 1024) {
        const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (200 * 1024) / tbytes));
        const unsigned tgrid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv(n, 256), (int64_t)ctx->sm_count * per_sm));
#define B2K_VTL(DR)                                                                                                          \
    do {                                                                                                                     \
        static PerDeviceOnce vattr;                                                                                          \
        if (vattr.need(ctx->device)) {                                                                                       \
            CUDA_TRY(cudaFuncSetAttribute(screen_verify_table_listed_kernel<DR>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          100 * 1024));                                                                      \
            vattr.done(ctx->device);                                                                                         \
        }                                                                                                                    \
        screen_verify_table_listed_kernel<DR><<<tgrid, 256, tbytes, st>>>(dX, n, p->d, dC, p->k, p->cand, p->ncand, tlist,   \
                                                                          lcap, TILE_M << unit_shift, labels, lloyd, p->params, p->fb_list, \
                                                                          gstride, p->cg,                                    \
                                                                          row_load_width(dX, p->d, ctx->row_vec_max));       \
    } while (0)
        if (ds == 4) B2K_VTL(4);
        else if (ds == 8) B2K_VTL(8);
        else if (ds == 12) B2K_VTL(12);
        else B2K_VTL(16);
#undef B2K_VTL
    } else if (ctx->verify_mode == 0 && p->d % 4 == 0 && (((uintptr_t)dX | (uintptr_t)dC) & 15) == 0) {
        // one thread per frame (the default for sorted frames)
        const unsigned fgrid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv(n, 256), (int64_t)ctx->sm_count * 8));
        screen_verify_frame_listed_kernel<<<fgrid, 256, 0, st>>>(dX, n, p->d, dC, p->k, p->cand, p->ncand, tlist, lcap,
                                                                 TILE_M << unit_shift, labels, lloyd, p->params, p->fb_list, p->cg);
    } else if (ctx->verify_mode == 3 && p->d % 4 == 0 && (((uintptr_t)dX | (uintptr_t)dC) & 15) == 0 && lcap <= 8192 &&
               (size_t)(p->d + 4) * 4 * 8 <= 64 * 1024) {
        // tile verify: the center rows a tile needs are staged once in shared memory
        const int rs = p->d + 4;
        // a small batch (a tile rarely needs more than a dozen rows): ~9 KB of shared memory per CTA keeps 12+ CTAs per SM
        const int batch = (int)std::max<size_t>(8, std::min<size_t>(32, (9 * 1024) / ((size_t)rs * 4)));
        const size_t vsm = (size_t)batch * rs * 4;
        static PerDeviceOnce vattr;
        if (vattr.need(ctx->device)) {
            CUDA_TRY(cudaFuncSetAttribute(screen_verify_tile_listed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            vattr.done(ctx->device);
        }
        const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(16, (200 * 1024) / (vsm + 4096)));
        const unsigned tgrid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv(n, TILE_M), (int64_t)ctx->sm_count * per_sm));
        screen_verify_tile_listed_kernel<<<tgrid, TILE_M, vsm, st>>>(dX, n, p->d, dC, p->k, p->cand, p->ncand, tlist, lcap,
                                                                     unit_shift, labels, lloyd, p->params, p->fb_list, p->cg, batch);
    } else {
        const bool vec = p->d % 4 == 0 && (((uintptr_t)dX | (uintptr_t)dC) & 15) == 0;
        const unsigned dgrid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv(n, 32), (int64_t)ctx->sm_count * 8));
        if (vec)
            screen_verify_direct_kernel<true><<<dgrid, 256, 0, st>>>(dX, n, p->d, dC, p->k, p->cand, p->ncand, labels, nullptr,
                                                                     lloyd, p->params, p->fb_list, p->cg, tlist, lcap, TILE_M << unit_shift);
        else
            screen_verify_direct_kernel<false><<<dgrid, 256, 0, st>>>(dX, n, p->d, dC, p->k, p->cand, p->ncand, labels, nullptr,
                                                                      lloyd, p->params, p->fb_list, p->cg, tlist, lcap, TILE_M << unit_shift);
    }
    LAUNCH_CHECK();
    return screen_finish_assign(p, dX, n, dC, labels, nullptr, lloyd);
}

void screen_plan_invalidate_frames(ScreenPlan* p) { if (p) p->prepared_n = -1; }

int screen_plan_terms(const ScreenPlan* p) { return p ? p->terms : 0; }

int screen_read_stats(ScreenPlan* p, double* cand_chunks, double* fallback_frames) {
    ScreenParams h;
    CUDA_TRY(cudaMemcpyAsync(&h, p->params, sizeof(h), cudaMemcpyDeviceToHost, p->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(p->ctx->stream));
    *cand_chunks = (double)h.cand_chunks;
    *fallback_frames = (double)h.fallback_frames;
    return B2K_OK;
}

}  // namespace b2k
