// screen.cu -- tcgen05 distance screen + exact fp32 verify (K2 / K2v of SURVEY 2.2).
//
// Replaces the N x k distance evaluations of deeptime assign_chunk_to_centers / kmeans.cluster
// (call sites pyemma/coordinates/clustering/interface.py:164-165, kmeans.py:254-258) when k*d is
// large.  The reference's argmin must be reproduced bit for bit, and a direct evaluation in the
// reference's operation order costs 3 CUDA-core instructions per pair-dimension.  So:
//
//  1. screen   acc[i][j] = x~_i . c~_j - |c~_j|^2/2   (maximising acc == minimising the distance)
//              as ONE fp16 GEMM on the 5th-gen tensor cores: tcgen05.mma, operands staged by TMA
//              (SWIZZLE_128B), fp32 accumulators in TMEM.  x~ = (x-mu)*sigma is centred/scaled so it
//              fits fp16; precision comes from splitting each fp32 value into hi+lo fp16 parts and
//              concatenating along K:   A' = [x_hi | x_hi | x_lo | 1 1 1],  B' = [c_hi | c_lo | c_hi |
//              -b1 -b2 -b3]  (terms=3; terms=1 keeps only the hi parts), b = |c~|^2/2 split in 3.
//              The N x k score matrix is never written: 4 epilogue warps read the accumulators with
//              tcgen05.ld, keep a running row maximum (FMNMX3) and remember every 32-column chunk
//              whose maximum is within a rigorous error margin of it.
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
static constexpr int CHUNK = 32;      // columns per candidate chunk
static constexpr int LIST_CAP = 16;   // running candidate chunks remembered per frame
static constexpr int CAND_CAP = 4;    // candidate chunks handed to the verify kernel per frame
static constexpr int A_BYTES = TILE_M * BLOCK_K * 2;
static constexpr int B_BYTES = TILE_N * BLOCK_K * 2;
static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
static constexpr int GEMM_THREADS = 192;  // warp0 TMA, warp1 MMA + TMEM alloc, warps 2-5 epilogue

struct ScreenParams {  // device resident; written by the prep kernels, read by everything else
    float sigma;       // power-of-two scale
    float cmax;        // upper bound of max_j |c~_j|
    float xmax2_raw;   // max_i |x_i - mu|^2 over the prepared frames
    float cmax2_raw;   // max_j |c_j - mu|^2 at prepare time
    float cmax2_now;   // max_j |c~_j|^2 of the current B' operand
    int valid;         // 0: operands unusable -> every frame takes the exact fallback
    unsigned long long cand_chunks, fallback_frames;  // statistics of the last verify
};

struct ScreenPlan {
    b2k_ctx* ctx = nullptr;
    int64_t n_cap = 0, n_pad = 0;
    int d = 0, k = 0, k_pad = 0, terms = 3, Kc = 0, Kp = 0, nk16 = 0;
    __half* A = nullptr;       // [n_pad][Kp]
    __half* B = nullptr;       // [k_pad][Kp]
    float* X2 = nullptr;       // [n_pad] |x~|^2
    float* mu = nullptr;       // [d]
    ScreenParams* params = nullptr;
    uint16_t* cand = nullptr;  // [n_pad][CAND_CAP]
    uint8_t* ncand = nullptr;  // [n_pad]  (255: overflow -> exact fallback)
    CUtensorMap tmA, tmB;
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
__global__ void screen_mu_kernel(const float* __restrict__ C, int k, int d, float* __restrict__ mu) {
    const int dim = blockIdx.x * blockDim.x + threadIdx.x;
    if (dim >= d) return;
    double s = 0;
    for (int j = 0; j < k; ++j) s += (double)C[(int64_t)j * d + dim];
    mu[dim] = (float)(s / k);
}

// max_i |row_i - mu|^2 (fp32, any order; only used to pick the scale) -> *out (float bits, atomicMax)
__global__ void __launch_bounds__(256) screen_maxnorm_kernel(const float* __restrict__ X, int64_t n, int d,
                                                             const float* __restrict__ mu, float* out) {
    float m = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const float* p = X + i * d;
        float s = 0.f;
        for (int e = 0; e < d; ++e) { const float t = p[e] - mu[e]; s += t * t; }
        m = fmaxf(m, s);  // NaN rows are ignored here and flagged in the operand kernel
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax((int*)out, __float_as_int(m));
}

__global__ void screen_sigma_kernel(ScreenParams* p) {
    const float mx = sqrtf(fmaxf(p->xmax2_raw, p->cmax2_raw));
    float sigma = 1.f;
    int valid = 1;
    if (!(mx < 3.0e38f)) valid = 0;
    else if (mx > 0.f) {
        // largest power of two with mx*sigma <= 200  (=> |x~|,|c~| <= 200 < 256, |c~|^2/2 <= 2e4 < 65504)
        int e;
        frexpf(200.f / fmaxf(mx, 1e-30f), &e);  // 200/mx = f*2^e, f in [0.5,1)  -> 2^(e-1) <= 200/mx
        e -= 1;
        if (e > 100) e = 100;
        if (e < -100) { e = -100; valid = 0; }
        sigma = ldexpf(1.f, e);
    }
    p->sigma = sigma;
    p->valid = valid;
}

// A' rows: [x_hi | (x_hi | x_lo) | 1 1 1 | 0..]; X2[i] = |x~_i|^2; rows >= n are zero
__global__ void __launch_bounds__(256) screen_frames_kernel(const float* __restrict__ X, int64_t n, int64_t n_pad,
                                                            int d, int terms, int Kp,
                                                            const float* __restrict__ mu,
                                                            const ScreenParams* __restrict__ prm,
                                                            __half2* __restrict__ A) {
    const int half_cols = Kp >> 1;
    const int64_t total = n_pad * half_cols;
    const float sigma = prm->sigma;
    const int ones0 = terms * d;
    for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
        const int64_t i = t / half_cols;
        const int c0 = (int)(t - i * half_cols) * 2;
        float v[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int c = c0 + q;
            float out = 0.f;
            if (i < n) {
                if (c < ones0) {
                    const int seg = c / d, e = c - seg * d;
                    const float xt = __fmul_rn(__fsub_rn(X[i * d + e], mu[e]), sigma);
                    const float hi = __half2float(__float2half_rn(xt));
                    out = (seg == 2) ? __fsub_rn(xt, hi) : hi;  // rounded to fp16 below
                } else if (c < ones0 + 3) {
                    out = 1.f;
                }
            }
            v[q] = out;
        }
        A[t] = __floats2half2_rn(v[0], v[1]);
    }
}

__global__ void __launch_bounds__(256) screen_x2_kernel(const float* __restrict__ X, int64_t n, int64_t n_pad, int d,
                                                        const float* __restrict__ mu,
                                                        const ScreenParams* __restrict__ prm,
                                                        float* __restrict__ X2) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n_pad) return;
    float s = 0.f;
    if (i < n) {
        const float sigma = prm->sigma;
        const float* p = X + i * d;
        for (int e = 0; e < d; ++e) { const float t = (p[e] - mu[e]) * sigma; s += t * t; }
        if (!(s < 3.0e38f)) s = __int_as_float(0x7f800000);  // NaN/inf frame -> flagged by the epilogue
    }
    X2[i] = s;
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
    float s = 0.f;
    for (int e = 0; e < d; ++e) {
        const float ct = __fmul_rn(__fsub_rn(C[(int64_t)j * d + e], mu[e]), sigma);
        s += ct * ct;
        const __half hi = __float2half_rn(ct);
        const __half lo = __float2half_rn(__fsub_rn(ct, __half2float(hi)));
        row[e] = hi;
        if (terms == 3) { row[d + e] = lo; row[2 * d + e] = hi; }
    }
    const float b = 0.5f * s;
    const __half b1 = __float2half_rn(b);
    const float r1 = __fsub_rn(b, __half2float(b1));
    const __half b2 = __float2half_rn(r1);
    const float r2 = __fsub_rn(r1, __half2float(b2));
    const __half b3 = __float2half_rn(r2);
    row[ones0] = __hneg(b1);
    row[ones0 + 1] = __hneg(b2);
    row[ones0 + 2] = __hneg(b3);
    for (int c = ones0 + 3; c < Kp; ++c) row[c] = __float2half_rn(0.f);
    atomicMax((int*)&prm->cmax2_now, __float_as_int(s));
}

__global__ void screen_finish_centers_kernel(ScreenParams* p, int d) {
    // upper bound of max |c~| (the fp32 evaluation above is within (d+2) ulp)
    const float c2 = p->cmax2_now * (1.f + (d + 4) * 1.2e-7f);
    p->cmax = sqrtf(c2) * 1.000001f;
    if (!(p->cmax <= 256.f)) p->valid = 0;  // centers moved outside the scaled range
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
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
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

// ---- margin ----------------------------------------------------------------------------------------------------
struct Margin {
    float a, r, x2;
    __device__ __forceinline__ void init(float x2_, float C, int d, int nk16, int terms) {
        const float u = 5.9604645e-8f;
        const float gam = (0.25f * d + 9.f) * u * 1.01f;
        const float rho = 2.f * gam + 4.9e-7f;
        const float X = sqrtf(x2_) * (1.f + 2.f * gam);
        const float R = X + C;
        const float d1 = 2.1f * u * R * R;
        const float erep = (terms == 3) ? 3.01f * 2.3841858e-7f : (2.f * 4.8828125e-4f + 2.4e-7f) * 1.01f;
        const float eacc = (float)(nk16 * 17) * 1.1920929e-7f;
        const float d2 = erep * X * C + 3.01e-8f * sqrtf((float)d) * R + eacc * (1.01f * X * C + 0.5f * C * C) +
                         (gam + 1.2e-10f) * 0.5f * C * C;
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
    const float* X2;
    const ScreenParams* prm;
    uint16_t* cand;
    uint8_t* ncand;
};

// ---- the screen kernel ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEMM_THREADS, 1)
screen_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmArgs g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* tiles = smem;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    uint32_t* list_id = tmem_slot + 4;                                   // [LIST_CAP][128]
    float* list_v = reinterpret_cast<float*>(list_id + LIST_CAP * 128);  // [LIST_CAP][128]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
                for (int nt = 0; nt < g.n_ntiles; ++nt) {
                    for (int kb = 0; kb < g.n_kblocks; ++kb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* sa = tiles + stage * STAGE_BYTES;
                        mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                        tma_load_2d(sa, &tmA, &full_bar[stage], kb * BLOCK_K, tile * TILE_M);
                        tma_load_2d(sa + A_BYTES, &tmB, &full_bar[stage], kb * BLOCK_K, nt * TILE_N);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            // instruction descriptor: D=f32, A=B=f16, K-major both, N=256, M=128
            const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(TILE_N >> 3) << 17) |
                                   ((uint32_t)(TILE_M >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
                for (int nt = 0; nt < g.n_ntiles; ++nt, ++it) {
                    const uint32_t acc = it & 1u, accphase = (it >> 1) & 1u;
                    mbar_wait(&tempty_bar[acc], accphase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * TILE_N;
                    for (int kb = 0; kb < g.n_kblocks; ++kb) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(tiles + stage * STAGE_BYTES);
                        const uint64_t adesc = make_smem_desc(sa);
                        const uint64_t bdesc = make_smem_desc(sa + A_BYTES);
                        const int ksteps = min(BLOCK_K / 16, g.nk16 - kb * (BLOCK_K / 16));
                        for (int ks = 0; ks < ksteps; ++ks) {
                            // +32 bytes per K=16 step inside the 128-byte swizzle row (address field is >>4)
                            tc_mma_f16(d_tmem, adesc + (uint64_t)(ks * 2), bdesc + (uint64_t)(ks * 2), idesc,
                                       (kb | ks) != 0 ? 1u : 0u);
                        }
                        tc_commit(&empty_bar[stage]);  // smem slot free once these MMAs retire
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    tc_commit(&tfull_bar[acc]);  // accumulator stage complete
                }
            }
        }
    } else {
        // ===== epilogue: 4 warps, one frame (TMEM lane) per thread =====
        const int q = warp & 3;              // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;       // row inside the frame tile
        const int et = row;                  // list slot
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t it = 0;
        const float C = g.prm->cmax;
        const int valid_ops = g.prm->valid;
        for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
            const int64_t grow = (int64_t)tile * TILE_M + row;
            const float x2 = (grow < g.n) ? g.X2[grow] : 0.f;
            Margin mg;
            mg.init(x2, C, g.d, g.nk16, g.terms);
            float m = __int_as_float(0xff800000);  // -inf
            float thr = m;
            int cnt = 0;
            for (int nt = 0; nt < g.n_ntiles; ++nt, ++it) {
                const uint32_t acc = it & 1u, accphase = (it >> 1) & 1u;
                mbar_wait(&tfull_bar[acc], accphase);
                tc_fence_after();
                const uint32_t taddr = lane_addr + acc * TILE_N;
#pragma unroll 1
                for (int c = 0; c < TILE_N / CHUNK; ++c) {
                    float v[32];
                    tmem_ld32(taddr + c * CHUNK, v);
                    tmem_ld_wait();
                    float cm = fmaxf(fmaxf(v[0], v[1]), v[2]);
#pragma unroll
                    for (int e = 3; e + 1 < 32; e += 2) cm = fmaxf(fmaxf(cm, v[e]), v[e + 1]);
                    cm = fmaxf(cm, v[31]);
                    if (cm > m) { m = cm; thr = mg.threshold(m); }
                    if (cm >= thr) {
                        if (cnt < LIST_CAP) {
                            list_id[cnt * 128 + et] = (uint32_t)(nt * (TILE_N / CHUNK) + c);
                            list_v[cnt * 128 + et] = cm;
                        }
                        ++cnt;
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            }
            // final filter against the final maximum
            if (grow < g.n) {
                uint32_t ids[CAND_CAP] = {0, 0, 0, 0};
                int kept = 0;
                bool overflow = (cnt > LIST_CAP) || !(m > -3.0e38f) || !valid_ops || !(x2 < 3.0e38f);
                const int lim = cnt < LIST_CAP ? cnt : LIST_CAP;
                for (int t = 0; t < lim; ++t) {
                    if (list_v[t * 128 + et] >= thr) {
                        if (kept < CAND_CAP) ids[kept] = list_id[t * 128 + et];
                        ++kept;
                    }
                }
                if (kept > CAND_CAP || kept == 0) overflow = true;
                uint2 packed;
                packed.x = (ids[0] & 0xffffu) | (ids[1] << 16);
                packed.y = (ids[2] & 0xffffu) | (ids[3] << 16);
                *reinterpret_cast<uint2*>(g.cand + grow * CAND_CAP) = packed;
                g.ncand[grow] = overflow ? (uint8_t)255 : (uint8_t)kept;
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

// ---- verify: exact reference-order evaluation of the surviving chunks ----------------------------------------------
template <int DREG>
__global__ void __launch_bounds__(128) screen_verify_kernel(const float* __restrict__ X, int64_t n, int d,
                                                            const float* __restrict__ Cn, int k,
                                                            const uint16_t* __restrict__ cand,
                                                            const uint8_t* __restrict__ ncand,
                                                            int32_t* __restrict__ labels, float* __restrict__ mind,
                                                            int lloyd, ScreenParams* prm) {
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    unsigned long long my_chunks = 0, my_fb = 0;
    if (i < n) {
        const float* xg = X + i * d;
        float xr[DREG > 0 ? DREG : 1];
        if (DREG > 0) {
#pragma unroll
            for (int e = 0; e < DREG; ++e) xr[e] = e < d ? xg[e] : 0.f;
        }
        ArgMin am;
        am.init();
        const int nc = ncand[i];
        const int n_chunks_all = (k + CHUNK - 1) / CHUNK;
        const int loops = nc == 255 ? n_chunks_all : nc;
        const uint2 pk = *reinterpret_cast<const uint2*>(cand + i * CAND_CAP);
        for (int t = 0; t < loops; ++t) {
            int ch;
            if (nc == 255) ch = t;
            else ch = (t == 0) ? (pk.x & 0xffff) : (t == 1) ? (pk.x >> 16) : (t == 2) ? (pk.y & 0xffff) : (pk.y >> 16);
            const int j0 = ch * CHUNK, j1 = min(j0 + CHUNK, k);
            for (int j = j0; j < j1; ++j) {
                const float* c = Cn + (int64_t)j * d;
                float s;
                if (DREG > 0) {
                    Lanes4 L;
                    L.init();
                    const int d4 = d & ~3;
#pragma unroll
                    for (int e = 0; e < DREG; e += 4)
                        if (e < d4) L.add4(xr[e], xr[e + 1], xr[e + 2], xr[e + 3], c[e], c[e + 1], c[e + 2], c[e + 3]);
#pragma unroll
                    for (int e = 0; e < DREG; ++e)
                        if (e >= d4 && e < d) L.tail(xr[e], c[e]);
                    s = L.result();
                } else {
                    s = euclid_sq_exact(xg, c, d);
                }
                am.offer(s, j);
            }
        }
        labels[i] = (lloyd && am.j < 0) ? 0 : am.j;
        if (mind) mind[i] = am.j >= 0 ? __fsqrt_rn(am.s) : 3.402823466e+38f;
        my_chunks = (nc == 255) ? 0 : nc;
        my_fb = (nc == 255) ? 1 : 0;
    }
    // statistics (warp aggregated)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_chunks += __shfl_xor_sync(0xffffffffu, my_chunks, o);
        my_fb += __shfl_xor_sync(0xffffffffu, my_fb, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (my_chunks) atomicAdd(&prm->cand_chunks, my_chunks);
        if (my_fb) atomicAdd(&prm->fallback_frames, my_fb);
    }
}

// ---- plan -------------------------------------------------------------------------------------------------------------
static size_t gemm_smem_bytes() {
    return 1024 + (size_t)STAGES * STAGE_BYTES + (2 * STAGES + 4) * 8 + 16 + (size_t)LIST_CAP * 128 * 8;
}

bool screen_supported(const b2k_ctx* ctx, int d, int k, int64_t n) {
    if (ctx->engine == B2K_ENGINE_DIRECT) return false;
    if (d < 1 || d > 1300 || k < 2 || k > 65535 * CHUNK) return false;
    if (gemm_smem_bytes() > ctx->smem_optin) return false;
    if (ctx->engine == B2K_ENGINE_SCREEN) return true;
    // auto: the screen pays off once a frame meets enough center coordinates
    return k >= 128 && (int64_t)k * d >= 2048 && n >= 4096;
}

void screen_plan_destroy(ScreenPlan* p) {
    if (!p) return;
    cudaStreamSynchronize(p->ctx->stream);
    cudaFree(p->A); cudaFree(p->B); cudaFree(p->X2); cudaFree(p->mu); cudaFree(p->params); cudaFree(p->cand);
    cudaFree(p->ncand);
    delete p;
}

int screen_plan_create(b2k_ctx* ctx, int64_t n_cap, int d, int k, ScreenPlan** out) {
    ScreenPlan* p = new ScreenPlan();
    p->ctx = ctx;
    p->n_cap = n_cap;
    p->n_pad = cdiv(n_cap, TILE_M) * TILE_M;
    p->d = d;
    p->k = k;
    p->k_pad = (int)(cdiv(k, TILE_N) * TILE_N);
    p->terms = ctx->screen_terms == 1 ? 1 : (ctx->screen_terms == 3 ? 3 : (d <= 16 ? 1 : 3));
    p->Kc = p->terms * d + 3;
    p->Kp = (int)(cdiv(p->Kc, BLOCK_K) * BLOCK_K);
    p->nk16 = (int)cdiv(p->Kc, 16);
    cudaError_t e = cudaMalloc(&p->A, (size_t)p->n_pad * p->Kp * 2);
    if (e == cudaSuccess) e = cudaMalloc(&p->B, (size_t)p->k_pad * p->Kp * 2);
    if (e == cudaSuccess) e = cudaMalloc(&p->X2, (size_t)p->n_pad * 4);
    if (e == cudaSuccess) e = cudaMalloc(&p->mu, (size_t)d * 4);
    if (e == cudaSuccess) e = cudaMalloc(&p->params, sizeof(ScreenParams));
    if (e == cudaSuccess) e = cudaMalloc(&p->cand, (size_t)p->n_pad * CAND_CAP * 2);
    if (e == cudaSuccess) e = cudaMalloc(&p->ncand, (size_t)p->n_pad);
    if (e != cudaSuccess) {
        cudaGetLastError();
        screen_plan_destroy(p);
        return set_error(B2K_ERR_NOMEM, "screen plan: %s", cudaGetErrorString(e));
    }
    int rc = make_tmap(&p->tmA, p->A, (uint64_t)p->n_pad, (uint64_t)p->Kp, TILE_M);
    if (rc == B2K_OK) rc = make_tmap(&p->tmB, p->B, (uint64_t)p->k_pad, (uint64_t)p->Kp, TILE_N);
    if (rc != B2K_OK) { screen_plan_destroy(p); return rc; }
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t ae = cudaFuncSetAttribute(screen_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)gemm_smem_bytes());
        if (ae != cudaSuccess) { screen_plan_destroy(p); return set_error(B2K_ERR_CUDA, "screen smem attr: %s", cudaGetErrorString(ae)); }
        attr_set = true;
    }
    *out = p;
    return B2K_OK;
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
    screen_mu_kernel<<<(unsigned)cdiv(p->d, 128), 128, 0, st>>>(dC, p->k, p->d, p->mu);
    LAUNCH_CHECK();
    screen_maxnorm_kernel<<<capped_grid(ctx, n, 256), 256, 0, st>>>(dX, n, p->d, p->mu, &p->params->xmax2_raw);
    LAUNCH_CHECK();
    screen_maxnorm_kernel<<<capped_grid(ctx, p->k, 256), 256, 0, st>>>(dC, p->k, p->d, p->mu, &p->params->cmax2_raw);
    LAUNCH_CHECK();
    screen_sigma_kernel<<<1, 1, 0, st>>>(p->params);
    LAUNCH_CHECK();
    const int64_t n_pad_now = cdiv(n, TILE_M) * TILE_M;
    screen_frames_kernel<<<capped_grid(ctx, n_pad_now * (p->Kp / 2), 256), 256, 0, st>>>(
        dX, n, n_pad_now, p->d, p->terms, p->Kp, p->mu, p->params, reinterpret_cast<__half2*>(p->A));
    LAUNCH_CHECK();
    screen_x2_kernel<<<(unsigned)cdiv(n_pad_now, 256), 256, 0, st>>>(dX, n, n_pad_now, p->d, p->mu, p->params, p->X2);
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

int screen_assign(ScreenPlan* p, const float* dX, int64_t n, const float* dC, int32_t* labels, float* mind,
                  int lloyd) {
    b2k_ctx* ctx = p->ctx;
    cudaStream_t st = ctx->stream;
    if (p->prepared_n != n) B2K_TRY(screen_prepare_frames_with_centers(p, dX, n, dC));
    // center operand for the current centers
    CUDA_TRY(cudaMemsetAsync(&p->params->cmax2_now, 0, 4, st));
    CUDA_TRY(cudaMemsetAsync(&p->params->cand_chunks, 0, 16, st));
    screen_centers_kernel<<<(unsigned)cdiv(p->k_pad, 128), 128, 0, st>>>(dC, p->k, p->k_pad, p->d, p->terms, p->Kp,
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
    g.prm = p->params;
    g.cand = p->cand;
    g.ncand = p->ncand;
    const unsigned grid = (unsigned)std::min<int64_t>(g.n_tiles, ctx->sm_count);
    screen_gemm_kernel<<<grid, GEMM_THREADS, gemm_smem_bytes(), st>>>(p->tmA, p->tmB, g);
    LAUNCH_CHECK();
    const unsigned vgrid = (unsigned)cdiv(n, 128);
    if (p->d <= 4)
        screen_verify_kernel<4><<<vgrid, 128, 0, st>>>(dX, n, p->d, dC, p->k, p->cand, p->ncand, labels, mind, lloyd, p->params);
    else if (p->d <= 8)
        screen_verify_kernel<8><<<vgrid, 128, 0, st>>>(dX, n, p->d, dC, p->k, p->cand, p->ncand, labels, mind, lloyd, p->params);
    else if (p->d <= 12)
        screen_verify_kernel<12><<<vgrid, 128, 0, st>>>(dX, n, p->d, dC, p->k, p->cand, p->ncand, labels, mind, lloyd, p->params);
    else if (p->d <= 16)
        screen_verify_kernel<16><<<vgrid, 128, 0, st>>>(dX, n, p->d, dC, p->k, p->cand, p->ncand, labels, mind, lloyd, p->params);
    else
        screen_verify_kernel<0><<<vgrid, 128, 0, st>>>(dX, n, p->d, dC, p->k, p->cand, p->ncand, labels, mind, lloyd, p->params);
    LAUNCH_CHECK();
    return B2K_OK;
}

int screen_read_stats(ScreenPlan* p, double* cand_chunks, double* fallback_frames) {
    ScreenParams h;
    CUDA_TRY(cudaMemcpyAsync(&h, p->params, sizeof(h), cudaMemcpyDeviceToHost, p->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(p->ctx->stream));
    *cand_chunks = (double)h.cand_chunks;
    *fallback_frames = (double)h.fallback_frames;
    return B2K_OK;
}

}  // namespace b2k
