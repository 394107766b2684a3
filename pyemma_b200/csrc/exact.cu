// exact.cu -- exact fp32 CUDA-core kernels of the Euclidean path (K1 of SURVEY 2.2).
//
// Every distance here is evaluated in the pinned operation order of the reference build
// (common.cuh Lanes4 / SURVEY Appendix B.1) so that the argmin is bit-identical to the
// reference's `argmin_j sqrt(sum)` with lowest-index tie break (deeptime
// assign_chunk_to_centers, call site pyemma/coordinates/clustering/interface.py:164-165).
//
//   assign_small_kernel<D>  d <= 16: frame in registers, centers broadcast from smem.
//   tile_kernel<MODE>       any d: a tile of FB frames is staged coalesced into padded
//                           smem, 128 threads = FB frames x G center groups, 4 centers
//                           register-tiled per thread.  MODE_ARGMIN -> labels (+min dist),
//                           MODE_ALL -> all distances out[j][i] (k-means++ candidates,
//                           regspace steps).
//   labeled_dist_kernel     l_i = sqrt(dist2(x_i, C[label_i])) (cost function).
#include "common.cuh"
#include "kernels.h"

namespace b2k {

// -------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) assign_small_kernel(const float* __restrict__ X, int64_t n,
                                                           const float* __restrict__ C, int k, int kt,
                                                           int32_t* __restrict__ labels, float* __restrict__ mind,
                                                           int lloyd, const int* __restrict__ run_if_zero) {
    extern __shared__ __align__(16) float cs[];
    if (run_if_zero && *run_if_zero != 0) return;
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const bool valid = i < n;
    float x[D];
    if (valid) {
#pragma unroll
        for (int e = 0; e < D; ++e) x[e] = __ldg(X + i * D + e);
    }
    ArgMin am;
    am.init();
    for (int j0 = 0; j0 < k; j0 += kt) {
        const int kk = min(kt, k - j0);
        __syncthreads();
        for (int t = threadIdx.x; t < kk * D; t += 256) cs[t] = __ldg(C + (int64_t)j0 * D + t);
        __syncthreads();
        if (valid) {
#pragma unroll 4
            for (int j = 0; j < kk; ++j) {
                const float* c = cs + j * D;
                Lanes4 L;
                L.init();
#pragma unroll
                for (int e = 0; e + 3 < D; e += 4)
                    L.add4(x[e], x[e + 1], x[e + 2], x[e + 3], c[e], c[e + 1], c[e + 2], c[e + 3]);
#pragma unroll
                for (int e = D & ~3; e < D; ++e) L.tail(x[e], c[e]);
                am.offer(D < 4 ? L.a0 : L.result(), j0 + j);
            }
        }
    }
    if (valid) {
        labels[i] = (lloyd && am.j < 0) ? 0 : am.j;
        if (mind) mind[i] = am.j >= 0 ? __fsqrt_rn(am.s) : 3.402823466e+38f;
    }
}

// -------------------------------------------------------------------------------------------
struct TileCfg {
    int FB, G, KT, ds, xstride;
    size_t smem;
};

static TileCfg tile_cfg(int d, int k, size_t smem_budget) {
    TileCfg c;
    c.ds = (d + 3) & ~3;
    c.xstride = ((c.ds / 4) % 2 == 0) ? c.ds + 4 : c.ds + 8;
    int FB = 128;
    // X tile may use at most ~60% of the budget
    while (FB > 4 && (size_t)FB * c.xstride * 4 > smem_budget * 6 / 10) FB >>= 1;
    c.FB = FB;
    c.G = 128 / FB;
    size_t left = smem_budget - (size_t)FB * c.xstride * 4 - 128 * 8;
    int KT = (int)(left / ((size_t)c.ds * 4));
    KT = (KT / (4 * c.G)) * (4 * c.G);
    int kmax = (int)cdiv(k, 4 * c.G) * 4 * c.G;
    if (KT > kmax) KT = kmax;
    if (KT < 4 * c.G) KT = 4 * c.G;
    c.KT = KT;
    c.smem = (size_t)FB * c.xstride * 4 + (size_t)KT * c.ds * 4 + 128 * 8;
    return c;
}

template <int MODE>
__global__ void __launch_bounds__(128) tile_kernel(const float* __restrict__ X, int64_t n, int d,
                                                   const float* __restrict__ C, int k, TileCfg cfg,
                                                   int32_t* __restrict__ labels, float* __restrict__ out,
                                                   int lloyd, const int* __restrict__ run_if_zero) {
    extern __shared__ __align__(16) float sm[];
    if (run_if_zero && *run_if_zero != 0) return;
    float* xs = sm;
    float* cs = sm + (size_t)cfg.FB * cfg.xstride;
    float* red_s = cs + (size_t)cfg.KT * cfg.ds;
    int32_t* red_j = (int32_t*)(red_s + 128);
    const int tid = threadIdx.x;
    const int f = tid % cfg.FB, g = tid / cfg.FB;
    const int64_t base = (int64_t)blockIdx.x * cfg.FB;
    const int nf = (int)min((int64_t)cfg.FB, n - base);
    const int d4 = d & ~3;

    // stage the frame tile (coalesced over the flattened tile)
    {
        const float* src = X + base * d;
        const int total = nf * d;
        for (int t = tid; t < total; t += 128) {
            const int r = t / d, c = t - r * d;
            xs[(size_t)r * cfg.xstride + c] = __ldg(src + t);
        }
    }
    ArgMin am;
    am.init();
    const float* xrow = xs + (size_t)f * cfg.xstride;
    const bool valid = f < nf;

    for (int j0 = 0; j0 < k; j0 += cfg.KT) {
        const int kk = min(cfg.KT, k - j0);
        __syncthreads();
        {
            const float* src = C + (int64_t)j0 * d;
            const int total = kk * d;
            for (int t = tid; t < total; t += 128) {
                const int r = t / d, c = t - r * d;
                cs[(size_t)r * cfg.ds + c] = __ldg(src + t);
            }
        }
        __syncthreads();
        if (!valid) continue;
        for (int jj = g; jj < kk; jj += 4 * cfg.G) {
            // 4 centers jj, jj+G, jj+2G, jj+3G (clamped duplicates are discarded below)
            const float* cr[4];
            int jx[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                jx[c] = jj + c * cfg.G;
                cr[c] = cs + (size_t)min(jx[c], kk - 1) * cfg.ds;
            }
            Lanes4 L[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) L[c].init();
            for (int e = 0; e < d4; e += 4) {
                const float4 xv = *reinterpret_cast<const float4*>(xrow + e);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float4 cv = *reinterpret_cast<const float4*>(cr[c] + e);
                    L[c].add4(xv.x, xv.y, xv.z, xv.w, cv.x, cv.y, cv.z, cv.w);
                }
            }
            for (int e = d4; e < d; ++e) {
                const float xv = xrow[e];
#pragma unroll
                for (int c = 0; c < 4; ++c) L[c].tail(xv, cr[c][e]);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (jx[c] < kk) {
                    const float s = L[c].result();
                    if (MODE == MODE_ARGMIN) am.offer(s, j0 + jx[c]);
                    else out[(int64_t)(j0 + jx[c]) * n + base + f] = __fsqrt_rn(s);
                }
            }
        }
    }
    if (MODE == MODE_ARGMIN) {
        if (cfg.G > 1) {
            __syncthreads();
            red_s[tid] = am.s;
            red_j[tid] = am.j;
            __syncthreads();
            if (g == 0) {
                for (int gg = 1; gg < cfg.G; ++gg) am.merge(red_s[gg * cfg.FB + f], red_j[gg * cfg.FB + f]);
            }
        }
        if (g == 0 && valid) {
            labels[base + f] = (lloyd && am.j < 0) ? 0 : am.j;
            if (out) out[base + f] = am.j >= 0 ? __fsqrt_rn(am.s) : 3.402823466e+38f;
        }
    }
}

// -------------------------------------------------------------------------------------------
// The tile kernel over an INDEX LIST of frames whose length lives on the device: the exact scan of the frames the
// tcgen05 screen could not bound (screen.cu fb_list / fb_count).  Persistent grid, so a zero count costs one
// empty launch; rows are gathered while staging, labels scattered through the same list.
__global__ void __launch_bounds__(128) tile_indexed_kernel(const float* __restrict__ X, int d,
                                                           const float* __restrict__ C, int k, TileCfg cfg,
                                                           const uint32_t* __restrict__ row_index,
                                                           const unsigned int* __restrict__ count_dev,
                                                           const int* __restrict__ run_if_nonzero,
                                                           int32_t* __restrict__ labels, float* __restrict__ mind,
                                                           int lloyd, unsigned int min_count) {
    extern __shared__ __align__(16) float sm[];
    if (run_if_nonzero && *run_if_nonzero == 0) return;
    const unsigned int count = *count_dev;
    if (count < min_count) return;  // a handful of frames: the CTA-per-frame scan takes them (screen.cu)
    float* xs = sm;
    float* cs = sm + (size_t)cfg.FB * cfg.xstride;
    float* red_s = cs + (size_t)cfg.KT * cfg.ds;
    int32_t* red_j = (int32_t*)(red_s + 128);
    const int tid = threadIdx.x;
    const int f = tid % cfg.FB, g = tid / cfg.FB;
    const int d4 = d & ~3;
    for (unsigned int base = blockIdx.x * (unsigned)cfg.FB; base < count; base += gridDim.x * (unsigned)cfg.FB) {
        const int nf = (int)min((unsigned)cfg.FB, count - base);
        __syncthreads();  // previous tile fully consumed
        for (int r = 0; r < nf; ++r) {
            const float* src = X + (int64_t)row_index[base + r] * d;
            for (int c = tid; c < d; c += 128) xs[(size_t)r * cfg.xstride + c] = __ldg(src + c);
        }
        ArgMin am;
        am.init();
        const float* xrow = xs + (size_t)f * cfg.xstride;
        const bool valid = f < nf;
        for (int j0 = 0; j0 < k; j0 += cfg.KT) {
            const int kk = min(cfg.KT, k - j0);
            __syncthreads();
            {
                const float* src = C + (int64_t)j0 * d;
                const int total = kk * d;
                for (int t = tid; t < total; t += 128) {
                    const int r = t / d, c = t - r * d;
                    cs[(size_t)r * cfg.ds + c] = __ldg(src + t);
                }
            }
            __syncthreads();
            if (!valid) continue;
            for (int jj = g; jj < kk; jj += 4 * cfg.G) {
                const float* cr[4];
                int jx[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    jx[c] = jj + c * cfg.G;
                    cr[c] = cs + (size_t)min(jx[c], kk - 1) * cfg.ds;
                }
                Lanes4 L[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) L[c].init();
                for (int e = 0; e < d4; e += 4) {
                    const float4 xv = *reinterpret_cast<const float4*>(xrow + e);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float4 cv = *reinterpret_cast<const float4*>(cr[c] + e);
                        L[c].add4(xv.x, xv.y, xv.z, xv.w, cv.x, cv.y, cv.z, cv.w);
                    }
                }
                for (int e = d4; e < d; ++e) {
                    const float xv = xrow[e];
#pragma unroll
                    for (int c = 0; c < 4; ++c) L[c].tail(xv, cr[c][e]);
                }
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (jx[c] < kk) am.offer(L[c].result(), j0 + jx[c]);
            }
        }
        if (cfg.G > 1) {
            __syncthreads();
            red_s[tid] = am.s;
            red_j[tid] = am.j;
            __syncthreads();
            if (g == 0)
                for (int gg = 1; gg < cfg.G; ++gg) am.merge(red_s[gg * cfg.FB + f], red_j[gg * cfg.FB + f]);
        }
        if (g == 0 && valid) {
            const int64_t i = row_index[base + f];
            labels[i] = (lloyd && am.j < 0) ? 0 : am.j;
            if (mind) mind[i] = am.j >= 0 ? __fsqrt_rn(am.s) : 3.402823466e+38f;
        }
    }
}

// -------------------------------------------------------------------------------------------
// l_i = sqrt(dist2(x_i, C[label_i]))  -- thread per frame, rows streamed from global/L1.
__global__ void __launch_bounds__(256) labeled_dist_kernel(const float* __restrict__ X, int64_t n, int d,
                                                           const float* __restrict__ C,
                                                           const int32_t* __restrict__ labels,
                                                           float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int32_t a = labels[i];
    out[i] = __fsqrt_rn(euclid_sq_exact(X + i * d, C + (int64_t)a * d, d));
}

// wide rows (d > 16): a warp takes 8 frames at a time.  Their rows and the rows of their centers are copied
// coalesced into shared memory (row stride rsb, rsb/4 odd -> conflict free); lane (f = lane/4, l = lane%4) then
// owns accumulator lane l of frame f -- elements l, l+4, ... in order, the d%4 tail into lane 0 -- i.e. exactly
// the reference's four interleaved partial sums, closed as ((a0+a1)+a2)+a3 by lane (f,0).
__global__ void __launch_bounds__(256) labeled_dist_wide_kernel(const float* __restrict__ X, int64_t n, int d,
                                                                const float* __restrict__ C,
                                                                const int32_t* __restrict__ labels,
                                                                float* __restrict__ out, int rsb, int warps_per_cta) {
    extern __shared__ __align__(16) float lsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp >= warps_per_cta) return;
    float* xb = lsm + (size_t)warp * 16 * rsb;
    float* cb = xb + (size_t)8 * rsb;
    const int f = lane >> 2, l = lane & 3;
    const int d4 = d & ~3;
    const bool vec4 = (d & 3) == 0;
    const int64_t stride = (int64_t)gridDim.x * warps_per_cta * 8;
    for (int64_t base = ((int64_t)blockIdx.x * warps_per_cta + warp) * 8; base < n; base += stride) {
        const int rows = (int)min((int64_t)8, n - base);
        const int32_t my_label = (lane < rows) ? labels[base + lane] : 0;
        __syncwarp();
        for (int r = 0; r < rows; ++r) {
            const int32_t a = __shfl_sync(0xffffffffu, my_label, r);
            const float* xs = X + (base + r) * d;
            const float* cs = C + (int64_t)a * d;
            if (vec4) {
                for (int t = lane; t < (d >> 2); t += 32) {
                    reinterpret_cast<float4*>(xb + (size_t)r * rsb)[t] = __ldg(reinterpret_cast<const float4*>(xs) + t);
                    reinterpret_cast<float4*>(cb + (size_t)r * rsb)[t] = __ldg(reinterpret_cast<const float4*>(cs) + t);
                }
            } else {
                for (int t = lane; t < d; t += 32) {
                    xb[(size_t)r * rsb + t] = __ldg(xs + t);
                    cb[(size_t)r * rsb + t] = __ldg(cs + t);
                }
            }
        }
        __syncwarp();
        float acc = 0.f;
        if (f < rows) {
            const float* xr = xb + (size_t)f * rsb;
            const float* cr = cb + (size_t)f * rsb;
#pragma unroll 4
            for (int e = l; e < d4; e += 4) {
                const float t = __fsub_rn(xr[e], cr[e]);
                acc = __fadd_rn(acc, __fmul_rn(t, t));
            }
            if (l == 0) {
                for (int e = d4; e < d; ++e) {
                    const float t = __fsub_rn(xr[e], cr[e]);
                    acc = __fadd_rn(acc, __fmul_rn(t, t));
                }
            }
        }
        const float a1 = __shfl_down_sync(0xffffffffu, acc, 1);
        const float a2 = __shfl_down_sync(0xffffffffu, acc, 2);
        const float a3 = __shfl_down_sync(0xffffffffu, acc, 3);
        if (l == 0 && f < rows) out[base + f] = __fsqrt_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc, a1), a2), a3));
    }
}

// wide rows (d > 16): FOUR lanes per frame.  Lane l of the quad owns accumulator lane l of the reference's sum --
// elements l, l+4, ... in order, the d%4 tail into lane 0 -- and reads exactly those elements straight from
// global memory (a warp instruction touches 8 rows x 16 contiguous bytes; the other half of every 32-byte sector
// is consumed by the next iteration out of L1).  No staging, no barriers: enough loads are in flight (8 per lane)
// to cover the HBM latency.  Closed as ((a0+a1)+a2)+a3 by lane 0 of the quad.
__global__ void __launch_bounds__(256) labeled_dist_quad_kernel(const float* __restrict__ X, int64_t n, int d,
                                                                const float* __restrict__ C,
                                                                const int32_t* __restrict__ labels,
                                                                float* __restrict__ out) {
    const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 2;
    const int l = threadIdx.x & 3;
    const bool live = i < n;
    const int64_t ii = live ? i : 0;
    const float* xr = X + ii * d;
    const float* cr = C + (int64_t)labels[ii] * d;
    const int d4 = d & ~3;
    float acc = 0.f;
    int e = l;
    for (; e + 28 < d4; e += 32) {
        float xv[8], cv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { xv[u] = __ldg(xr + e + 4 * u); cv[u] = __ldg(cr + e + 4 * u); }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float t = __fsub_rn(xv[u], cv[u]);
            acc = __fadd_rn(acc, __fmul_rn(t, t));
        }
    }
    for (; e < d4; e += 4) {
        const float t = __fsub_rn(__ldg(xr + e), __ldg(cr + e));
        acc = __fadd_rn(acc, __fmul_rn(t, t));
    }
    if (l == 0) {
        for (int q = d4; q < d; ++q) {
            const float t = __fsub_rn(__ldg(xr + q), __ldg(cr + q));
            acc = __fadd_rn(acc, __fmul_rn(t, t));
        }
    }
    const float a1 = __shfl_down_sync(0xffffffffu, acc, 1);
    const float a2 = __shfl_down_sync(0xffffffffu, acc, 2);
    const float a3 = __shfl_down_sync(0xffffffffu, acc, 3);
    if (l == 0 && live) out[i] = __fsqrt_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc, a1), a2), a3));
}

// 4x4 transpose inside a quad of lanes: lane l holds v[c] = element 4l+c of a 16-element chunk (one coalesced 16-byte
// load per lane); afterwards lane l holds v[q] = element 4q+l, i.e. the next four addends of accumulator lane l in
// order.  Two butterfly stages, 4 SHFL + 8 SEL.
__device__ __forceinline__ void quad_transpose(float (&v)[4], int l, unsigned mask = 0xffffffffu) {
    const bool b0 = l & 1, b1 = l & 2;
    float ta = __shfl_xor_sync(mask, b0 ? v[0] : v[1], 1);
    float tb = __shfl_xor_sync(mask, b0 ? v[2] : v[3], 1);
    if (b0) { v[0] = ta; v[2] = tb; } else { v[1] = ta; v[3] = tb; }
    ta = __shfl_xor_sync(mask, b1 ? v[0] : v[2], 2);
    tb = __shfl_xor_sync(mask, b1 ? v[1] : v[3], 2);
    if (b1) { v[0] = ta; v[1] = tb; } else { v[2] = ta; v[3] = tb; }
}

// d % 4 == 0 and 16-byte aligned rows: the quad reads 64 contiguous bytes per step (full sectors, no reliance on
// L1), squares the differences where they were loaded -- (x-c)^2 is elementwise, only the ORDER of the additions is
// pinned -- and transposes the squares so that every accumulator lane receives its addends in reference order.
__global__ void __launch_bounds__(256) labeled_dist_quad_vec_kernel(const float* __restrict__ X, int64_t n, int d,
                                                                    const float* __restrict__ C,
                                                                    const int32_t* __restrict__ labels,
                                                                    float* __restrict__ out) {
    const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 2;
    const int l = threadIdx.x & 3;
    const bool live = i < n;
    const int64_t ii = live ? i : 0;
    const float4* xr = reinterpret_cast<const float4*>(X + ii * d);
    const float4* cr = reinterpret_cast<const float4*>(C + (int64_t)labels[ii] * d);
    const int nv = d >> 2;  // float4 per row
    float acc = 0.f;
    for (int t0 = 0; t0 < nv; t0 += 16) {  // 4 chunks of 16 elements per trip: 8 loads in flight per lane
        float4 xv[4], cv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int t = t0 + 4 * u + l;
            xv[u] = t < nv ? __ldg(xr + t) : make_float4(0.f, 0.f, 0.f, 0.f);
            cv[u] = t < nv ? __ldg(cr + t) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (t0 + 4 * u < nv) {  // uniform over the quad
                float sq[4];
                float t = __fsub_rn(xv[u].x, cv[u].x); sq[0] = __fmul_rn(t, t);
                t = __fsub_rn(xv[u].y, cv[u].y); sq[1] = __fmul_rn(t, t);
                t = __fsub_rn(xv[u].z, cv[u].z); sq[2] = __fmul_rn(t, t);
                t = __fsub_rn(xv[u].w, cv[u].w); sq[3] = __fmul_rn(t, t);
                quad_transpose(sq, l);
                acc = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc, sq[0]), sq[1]), sq[2]), sq[3]);
            }
        }
    }
    const float a1 = __shfl_down_sync(0xffffffffu, acc, 1);
    const float a2 = __shfl_down_sync(0xffffffffu, acc, 2);
    const float a3 = __shfl_down_sync(0xffffffffu, acc, 3);
    if (l == 0 && live) out[i] = __fsqrt_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc, a1), a2), a3));
}

// out[j][i] = sqrt(dist2(x_i, rows_j)) for a FEW rows (k-means++ candidates: m = 2 + floor(ln k)).  Quad layout as
// above, TWO frames per quad.  The m rows sit in shared memory re-ordered per accumulator lane
// (cs[j][l][s] = row_j[4s+l], lane stride padded by 16 bytes so the four lanes of a quad hit four different bank
// groups): one 16-byte shared load feeds four steps of a lane's sum for both frames, i.e. 24 fp32 instructions per
// shared-memory wavefront group -- CUDA-core bound (3 non-fusable instructions per frame, row and dimension), not
// LDS bound.  VEC (d % 4 == 0, aligned): frames are read with coalesced 16-byte loads and transposed in registers;
// otherwise lane l reads its own elements and lane 0 adds the d%4 tail.
// PRUNE (k-means++ only, FULL kernels): a (frame, candidate) pair whose distance provably cannot undercut the frame's
// current D2 is not evaluated -- out = +inf, which the contribution min(D2, dist^2) turns into D2, exactly what the
// evaluation would have given.  Proof obligation: with a = the chosen center that realises D2_i, r = |x_i - c_a|,
// R = |c_a - cand|: |x_i - cand| >= R - r, and the fp32 reference-order value dd satisfies dd >= |x_i-cand|^2 (1-g),
// sqrt(D2_i) >= r (1-g), g <= (d/4+10) 2^-24; so R >= (2 + slack) sqrt(D2_i) implies dd >= D2_i.  R is computed in
// fp64 and rounded down (kmpp.cu), slack = 1e-4 + 4e-7 d >> g.
// Two phases so that the survivors are dense: kmpp_prune_scan_kernel (thread per frame) decides every pair, records
// the frame's candidate mask and appends frames with at least one live pair to a list; the quad kernel then walks
// that list (two listed frames per quad) instead of all frames, so no warp idles on pruned neighbours, and writes
// distances for live pairs only -- consumers (kmpp.cu potentials / D2 update) read the mask first.
struct DistRowsPrune {
    const float* D;                // [n] current D2
    const int32_t* assigned;       // [n] index (into the chosen centers) of the center realising D2
    const unsigned char* taken;    // [n]
    const float* Rc;               // [found][16] lower bounds of |chosen center a - candidate j| (j < m <= 14)
    int rc_stride;                 // = 16
    float factor;                  // 2 + slack
    uint32_t* list;                // [n] frames with live pairs
    uint32_t* masks;               // [n] their candidate masks (by list position)
    unsigned int* count;           // device counter (zeroed by the launcher)
    uint16_t* framemask;           // [n] candidate mask of every frame (0: all pairs pruned)
};

__global__ void __launch_bounds__(256) kmpp_prune_scan_kernel(int64_t n, int m, DistRowsPrune pr) {
    const int lane = threadIdx.x & 31;
    const int64_t n_round = (n + 31) & ~(int64_t)31;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n_round; i += (int64_t)gridDim.x * 256) {
        uint32_t need = 0u;
        if (i < n) {
            if (!pr.taken[i]) {
                const float thr = sqrtf(pr.D[i]) * pr.factor;
                const float4* rc = reinterpret_cast<const float4*>(pr.Rc + (size_t)pr.assigned[i] * 16);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (4 * q < m) {
                        const float4 r = __ldg(rc + q);
                        need |= (!(r.x >= thr) ? 1u : 0u) << (4 * q);
                        need |= (!(r.y >= thr) ? 1u : 0u) << (4 * q + 1);
                        need |= (!(r.z >= thr) ? 1u : 0u) << (4 * q + 2);
                        need |= (!(r.w >= thr) ? 1u : 0u) << (4 * q + 3);
                    }
                }
                need &= (1u << m) - 1u;
            }
            pr.framemask[i] = (uint16_t)need;
        }
        const unsigned live = __ballot_sync(0xffffffffu, need != 0u);
        if (live) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(pr.count, (unsigned)__popc(live));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (need) {
                const unsigned pos = base + __popc(live & ((1u << lane) - 1u));
                pr.list[pos] = (uint32_t)i;
                pr.masks[pos] = need;
            }
        }
    }
}

template <int MR, bool VEC, bool FULL, bool PRUNE>
__global__ void __launch_bounds__(256) dist_rows_quad_kernel(const float* __restrict__ X, int64_t n, int d,
                                                             const float* __restrict__ rows, int m,
                                                             float* __restrict__ out, int T, DistRowsPrune pr) {
    // FULL: m == MR, one pass over the rows without per-row bounds checks (the instantiations 2..14 cover every
    // k-means++ trial count up to k = 1.6e5); otherwise the rows are walked in groups of MR.
    extern __shared__ __align__(16) float rsm[];
    if (PRUNE && (int64_t)blockIdx.x * 128 >= (int64_t)*pr.count) return;  // beyond the survivor list
    const int Tp = T + 4;
    float* cs = rsm;                          // [m][4][Tp]
    float* tl = rsm + (size_t)m * 4 * Tp;     // [m][4] tail elements
    const int d4 = d & ~3;
    for (int idx = threadIdx.x; idx < m * 4 * T; idx += 256) {
        const int j = idx / (4 * T), r = idx - j * 4 * T, l = r / T, s = r - l * T;
        const int e = 4 * s + l;
        cs[((size_t)j * 4 + l) * Tp + s] = e < d4 ? __ldg(rows + (int64_t)j * d + e) : 0.f;
    }
    for (int idx = threadIdx.x; idx < m * 4; idx += 256) {
        const int j = idx >> 2, e = d4 + (idx & 3);
        tl[idx] = e < d ? __ldg(rows + (int64_t)j * d + e) : 0.f;
    }
    __syncthreads();
    // frames of this quad: i0, i0+1 -- or, PRUNE, two consecutive entries of the survivor list
    const int64_t q0 = (((int64_t)blockIdx.x * 256 + threadIdx.x) >> 2) * 2;
    const int l = threadIdx.x & 3;
    int64_t i0 = q0, i1 = q0 + 1;
    bool live0 = i0 < n, live1 = i1 < n;
    unsigned need0 = 0xffffffffu, need1 = 0xffffffffu;
    if (PRUNE) {
        const unsigned int count = *pr.count;
        live0 = q0 < count;
        live1 = q0 + 1 < count;
        i0 = live0 ? pr.list[q0] : 0;
        i1 = live1 ? pr.list[q0 + 1] : 0;
        need0 = live0 ? pr.masks[q0] : 0u;
        need1 = live1 ? pr.masks[q0 + 1] : 0u;
    }
    const float* xr0 = X + (live0 ? i0 : 0) * d;
    const float* xr1 = X + (live1 ? i1 : 0) * d;
    const int nv = d >> 2;
    const unsigned qmask = PRUNE ? (0xFu << (threadIdx.x & 28)) : 0xffffffffu;  // shuffles stay inside the quad
    const unsigned needq = need0 | need1;
    const bool load0 = live0 && need0 != 0u, load1 = live1 && need1 != 0u;
    auto fetch = [&](const float* xr, bool live, int s0, float (&x)[4]) {
        if (VEC) {
            const float4 v = (live && s0 + l < nv) ? __ldg(reinterpret_cast<const float4*>(xr) + s0 + l)
                                                   : make_float4(0.f, 0.f, 0.f, 0.f);
            x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) { const int e = 4 * (s0 + q) + l; x[q] = (live && e < d4) ? __ldg(xr + e) : 0.f; }
        }
    };
    for (int j0 = 0; j0 < m; j0 += MR) {
        float acc0[MR], acc1[MR];
#pragma unroll
        for (int jj = 0; jj < MR; ++jj) { acc0[jj] = 0.f; acc1[jj] = 0.f; }
        const float* crow = cs + ((size_t)j0 * 4 + l) * Tp;  // row jj of this lane: crow + jj*4*Tp
        const int rstride = 4 * Tp;
        auto advance = [&](float (&x0)[4], float (&x1)[4], int s0) {  // four steps of every row sum, both frames
            if (VEC) { quad_transpose(x0, l, qmask); quad_transpose(x1, l, qmask); }
#pragma unroll
            for (int jj = 0; jj < MR; ++jj) {
                if ((FULL || j0 + jj < m) && (!PRUNE || ((needq >> jj) & 1u))) {
                    const float4 c = *reinterpret_cast<const float4*>(crow + jj * rstride + s0);
                    float t = __fsub_rn(x0[0], c.x), u = __fsub_rn(x1[0], c.x);
                    float a = __fadd_rn(acc0[jj], __fmul_rn(t, t)), b = __fadd_rn(acc1[jj], __fmul_rn(u, u));
                    t = __fsub_rn(x0[1], c.y); u = __fsub_rn(x1[1], c.y);
                    a = __fadd_rn(a, __fmul_rn(t, t)); b = __fadd_rn(b, __fmul_rn(u, u));
                    t = __fsub_rn(x0[2], c.z); u = __fsub_rn(x1[2], c.z);
                    a = __fadd_rn(a, __fmul_rn(t, t)); b = __fadd_rn(b, __fmul_rn(u, u));
                    t = __fsub_rn(x0[3], c.w); u = __fsub_rn(x1[3], c.w);
                    acc0[jj] = __fadd_rn(a, __fmul_rn(t, t));
                    acc1[jj] = __fadd_rn(b, __fmul_rn(u, u));
                }
            }
        };
        // ping-pong prefetch: buffer A holds chunk s0, buffer B chunk s0+4 (no register rotation)
        float a0[4], a1[4], b0[4], b1[4];
        if (!PRUNE || needq != 0u) {
            fetch(xr0, load0, 0, a0);
            fetch(xr1, load1, 0, a1);
            for (int s0 = 0; s0 < T; s0 += 8) {
                if (s0 + 4 < T) { fetch(xr0, load0, s0 + 4, b0); fetch(xr1, load1, s0 + 4, b1); }
                advance(a0, a1, s0);
                if (s0 + 4 < T) {
                    if (s0 + 8 < T) { fetch(xr0, load0, s0 + 8, a0); fetch(xr1, load1, s0 + 8, a1); }
                    advance(b0, b1, s0 + 4);
                }
            }
        }
        if (!VEC && l == 0 && (!PRUNE || needq != 0u)) {
            for (int q = 0; q < d - d4; ++q) {
                const float xv0 = load0 ? __ldg(xr0 + d4 + q) : 0.f, xv1 = load1 ? __ldg(xr1 + d4 + q) : 0.f;
#pragma unroll
                for (int jj = 0; jj < MR; ++jj) {
                    if (FULL || j0 + jj < m) {
                        const float c = tl[(j0 + jj) * 4 + q];
                        const float t = __fsub_rn(xv0, c), u = __fsub_rn(xv1, c);
                        acc0[jj] = __fadd_rn(acc0[jj], __fmul_rn(t, t));
                        acc1[jj] = __fadd_rn(acc1[jj], __fmul_rn(u, u));
                    }
                }
            }
        }
#pragma unroll
        for (int jj = 0; jj < MR; ++jj) {
            const float a1s = __shfl_down_sync(qmask, acc0[jj], 1), b1s = __shfl_down_sync(qmask, acc1[jj], 1);
            const float a2s = __shfl_down_sync(qmask, acc0[jj], 2), b2s = __shfl_down_sync(qmask, acc1[jj], 2);
            const float a3s = __shfl_down_sync(qmask, acc0[jj], 3), b3s = __shfl_down_sync(qmask, acc1[jj], 3);
            if (l == 0 && (FULL || j0 + jj < m)) {
                float* o = out + (int64_t)(j0 + jj) * n;
                if (live0 && (!PRUNE || ((need0 >> jj) & 1u)))
                    o[i0] = __fsqrt_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc0[jj], a1s), a2s), a3s));
                if (live1 && (!PRUNE || ((need1 >> jj) & 1u)))
                    o[i1] = __fsqrt_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc1[jj], b1s), b2s), b3s));
            }
        }
    }
}

// -------------------------------------------------------------------------------------------
template <int D>
static int launch_small(b2k_ctx* ctx, const float* X, int64_t n, const float* C, int k, int32_t* labels, float* mind,
                        int lloyd, const int* run_if_zero) {
    const size_t budget = 96 * 1024;
    int kt = (int)std::min<int64_t>(k, budget / (D * 4));
    const size_t smem = (size_t)kt * D * 4;
    static PerDeviceOnce attr_set;
    if (attr_set.need(ctx->device)) {
        CUDA_TRY(cudaFuncSetAttribute(assign_small_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)budget));
        attr_set.done(ctx->device);
    }
    const int64_t blocks = cdiv(n, 256);
    assign_small_kernel<D><<<(unsigned)blocks, 256, smem, ctx->stream>>>(X, n, C, k, kt, labels, mind, lloyd, run_if_zero);
    LAUNCH_CHECK();
    return B2K_OK;
}

static int launch_tile_gated(b2k_ctx* ctx, const float* X, int64_t n, int d, const float* C, int k, int32_t* labels,
                             float* out, int lloyd, int mode, const int* run_if_zero);

// run_if_zero (device pointer, may be null): the kernel returns at once unless *run_if_zero == 0
int launch_assign_exact_if(b2k_ctx* ctx, const float* X, int64_t n, int d, const float* C, int k, int32_t* labels,
                           float* mind, int lloyd, const int* run_if_zero) {
    if (n <= 0) return B2K_OK;
    if (k <= 0) return set_error(B2K_ERR_INVALID_ARG, "assign: no centers");
    switch (d) {
#define CASE(D) case D: return launch_small<D>(ctx, X, n, C, k, labels, mind, lloyd, run_if_zero);
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
        CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14) CASE(15) CASE(16)
#undef CASE
        default: break;
    }
    return launch_tile_gated(ctx, X, n, d, C, k, labels, mind, lloyd, MODE_ARGMIN, run_if_zero);
}

int launch_assign_exact(b2k_ctx* ctx, const float* X, int64_t n, int d, const float* C, int k, int32_t* labels,
                        float* mind, int lloyd) {
    return launch_assign_exact_if(ctx, X, n, d, C, k, labels, mind, lloyd, nullptr);
}

int launch_tile(b2k_ctx* ctx, const float* X, int64_t n, int d, const float* C, int k, int32_t* labels, float* out,
                int lloyd, int mode) {
    return launch_tile_gated(ctx, X, n, d, C, k, labels, out, lloyd, mode, nullptr);
}

static int launch_tile_gated(b2k_ctx* ctx, const float* X, int64_t n, int d, const float* C, int k, int32_t* labels,
                             float* out, int lloyd, int mode, const int* run_if_zero) {
    if (n <= 0 || k <= 0) return B2K_OK;
    const size_t budget = std::min<size_t>(ctx->smem_optin, 200 * 1024);
    TileCfg cfg = tile_cfg(d, k, budget);
    if (cfg.smem > ctx->smem_optin)
        return set_error(B2K_ERR_INVALID_ARG, "dimension %d too large for the exact tile kernel", d);
    static PerDeviceOnce attr_set;
    if (attr_set.need(ctx->device)) {
        CUDA_TRY(cudaFuncSetAttribute(tile_kernel<MODE_ARGMIN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)ctx->smem_optin));
        CUDA_TRY(cudaFuncSetAttribute(tile_kernel<MODE_ALL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)ctx->smem_optin));
        attr_set.done(ctx->device);
    }
    const int64_t blocks = cdiv(n, cfg.FB);
    if (mode == MODE_ARGMIN)
        tile_kernel<MODE_ARGMIN><<<(unsigned)blocks, 128, cfg.smem, ctx->stream>>>(X, n, d, C, k, cfg, labels, out,
                                                                                    lloyd, run_if_zero);
    else
        tile_kernel<MODE_ALL><<<(unsigned)blocks, 128, cfg.smem, ctx->stream>>>(X, n, d, C, k, cfg, labels, out,
                                                                                 lloyd, run_if_zero);
    LAUNCH_CHECK();
    return B2K_OK;
}

int launch_tile_indexed(b2k_ctx* ctx, const float* X, int d, const float* C, int k, const uint32_t* row_index,
                        const unsigned int* count_dev, const int* run_if_nonzero, int32_t* labels, float* mind,
                        int lloyd, unsigned int min_count) {
    if (k <= 0) return B2K_OK;
    const size_t budget = std::min<size_t>(ctx->smem_optin, 100 * 1024);  // two CTAs per SM
    TileCfg cfg = tile_cfg(d, k, budget);
    if (cfg.smem > ctx->smem_optin)
        return set_error(B2K_ERR_INVALID_ARG, "dimension %d too large for the exact tile kernel", d);
    static PerDeviceOnce attr_set;
    if (attr_set.need(ctx->device)) {
        CUDA_TRY(cudaFuncSetAttribute(tile_indexed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)ctx->smem_optin));
        attr_set.done(ctx->device);
    }
    tile_indexed_kernel<<<ctx->sm_count * 2, 128, cfg.smem, ctx->stream>>>(X, d, C, k, cfg, row_index, count_dev,
                                                                          run_if_nonzero, labels, mind, lloyd, min_count);
    LAUNCH_CHECK();
    return B2K_OK;
}

template <int MR, bool FULL>
static int launch_dist_rows_quad(b2k_ctx* ctx, const float* X, int64_t n, int d, const float* rows, int m, float* out,
                                 int T, size_t smem, bool vec, const DistRowsPrune* prune) {
    static PerDeviceOnce attr_set;
    if (attr_set.need(ctx->device)) {
        CUDA_TRY(cudaFuncSetAttribute(dist_rows_quad_kernel<MR, true, FULL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        CUDA_TRY(cudaFuncSetAttribute(dist_rows_quad_kernel<MR, false, FULL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        if (FULL) {
            CUDA_TRY(cudaFuncSetAttribute(dist_rows_quad_kernel<MR, true, FULL, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
            CUDA_TRY(cudaFuncSetAttribute(dist_rows_quad_kernel<MR, false, FULL, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        }
        attr_set.done(ctx->device);
    }
    const unsigned grid = (unsigned)cdiv(cdiv(n, 2) * 4, 256);
    DistRowsPrune none = {};
    if (FULL && prune) {
        CUDA_TRY(cudaMemsetAsync(prune->count, 0, 4, ctx->stream));
        const unsigned sg = (unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv(n, 256), (int64_t)ctx->sm_count * 16));
        kmpp_prune_scan_kernel<<<sg, 256, 0, ctx->stream>>>(n, m, *prune);
        LAUNCH_CHECK();
        if (vec) dist_rows_quad_kernel<MR, true, FULL, FULL><<<grid, 256, smem, ctx->stream>>>(X, n, d, rows, m, out, T, *prune);
        else dist_rows_quad_kernel<MR, false, FULL, FULL><<<grid, 256, smem, ctx->stream>>>(X, n, d, rows, m, out, T, *prune);
    } else {
        if (vec) dist_rows_quad_kernel<MR, true, FULL, false><<<grid, 256, smem, ctx->stream>>>(X, n, d, rows, m, out, T, none);
        else dist_rows_quad_kernel<MR, false, FULL, false><<<grid, 256, smem, ctx->stream>>>(X, n, d, rows, m, out, T, none);
    }
    LAUNCH_CHECK();
    return B2K_OK;
}

// prune (optional, k-means++): see DistRowsPrune; ignored when the quad kernel does not apply
int launch_dist_rows_pruned(b2k_ctx* ctx, const float* X, int64_t n, int d, const float* rows, int m, float* out,
                            const float* D, const int32_t* assigned, const unsigned char* taken, const float* Rc,
                            int rc_stride, uint32_t* list, uint32_t* masks, unsigned int* count, uint16_t* framemask) {
    if (n <= 0 || m <= 0) return B2K_OK;
    const int T = std::max(4, (int)cdiv(d / 4, 4) * 4);  // steps per accumulator lane, padded to whole 16-byte loads
    const size_t smem = ((size_t)m * 4 * (T + 4) + (size_t)m * 4) * 4;
    const bool want_prune = D != nullptr && n < (int64_t(1) << 32);
    if (m <= 32 && smem <= 160 * 1024 && (n >= 64 || (want_prune && m <= 14))) {
        const bool vec = (d % 4 == 0) && (((uintptr_t)X) & 15) == 0;
        DistRowsPrune pr = {D, assigned, taken, Rc, rc_stride, 2.0f * (1.0f + 1e-4f + 4e-7f * (float)d), list, masks, count,
                            framemask};
        const DistRowsPrune* pp = want_prune ? &pr : nullptr;
        if (!pp && framemask) CUDA_TRY(cudaMemsetAsync(framemask, 0xFF, (size_t)n * 2, ctx->stream));
        switch (m) {
#define B2K_DR(M) case M: return launch_dist_rows_quad<M, true>(ctx, X, n, d, rows, m, out, T, smem, vec, pp);
            B2K_DR(1) B2K_DR(2) B2K_DR(3) B2K_DR(4) B2K_DR(5) B2K_DR(6) B2K_DR(7) B2K_DR(8) B2K_DR(9) B2K_DR(10)
            B2K_DR(11) B2K_DR(12) B2K_DR(13) B2K_DR(14)
#undef B2K_DR
            default:
                // every pair is evaluated: consumers of the frame masks must see them all live
                if (framemask) CUDA_TRY(cudaMemsetAsync(framemask, 0xFF, (size_t)n * 2, ctx->stream));
                return launch_dist_rows_quad<8, false>(ctx, X, n, d, rows, m, out, T, smem, vec, nullptr);
        }
    }
    if (framemask) CUDA_TRY(cudaMemsetAsync(framemask, 0xFF, (size_t)n * 2, ctx->stream));
    return launch_tile(ctx, X, n, d, rows, m, nullptr, out, 0, MODE_ALL);
}

int launch_dist_rows(b2k_ctx* ctx, const float* X, int64_t n, int d, const float* rows, int m, float* out) {
    return launch_dist_rows_pruned(ctx, X, n, d, rows, m, out, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr,
                                   nullptr, nullptr);
}

int launch_labeled_dist(b2k_ctx* ctx, const float* X, int64_t n, int d, const float* C, const int32_t* labels,
                        float* out) {
    if (n <= 0) return B2K_OK;
    if (d > 16 && ctx->cost_kernel != 1 && d % 4 == 0 && ((((uintptr_t)X) | ((uintptr_t)C)) & 15) == 0) {
        labeled_dist_quad_vec_kernel<<<(unsigned)cdiv(n * 4, 256), 256, 0, ctx->stream>>>(X, n, d, C, labels, out);
        LAUNCH_CHECK();
        return B2K_OK;
    }
    if (d > 16 && ctx->cost_kernel != 1) {  // (d = 10: the thread-per-frame kernel is faster, 0.16 vs 0.26 ms per 1e7)
        labeled_dist_quad_kernel<<<(unsigned)cdiv(n * 4, 256), 256, 0, ctx->stream>>>(X, n, d, C, labels, out);
        LAUNCH_CHECK();
        return B2K_OK;
    }
    if (d > 16) {
        const int dpad = (d + 3) & ~3;
        const int rsb = ((dpad / 4) % 2 == 1) ? dpad : dpad + 4;
        const size_t per_warp = (size_t)16 * rsb * 4;
        const int wpc = (int)std::min<size_t>(8, (96 * 1024) / per_warp);
        if (wpc >= 1) {
            static PerDeviceOnce attr_set;
            if (attr_set.need(ctx->device)) {
                CUDA_TRY(cudaFuncSetAttribute(labeled_dist_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              100 * 1024));
                attr_set.done(ctx->device);
            }
            const size_t smem = per_warp * wpc;
            const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / smem));
            const unsigned grid =
                (unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv(n, 8 * wpc), (int64_t)ctx->sm_count * per_sm));
            labeled_dist_wide_kernel<<<grid, 256, smem, ctx->stream>>>(X, n, d, C, labels, out, rsb, wpc);
            LAUNCH_CHECK();
            return B2K_OK;
        }
    }
    labeled_dist_kernel<<<(unsigned)cdiv(n, 256), 256, 0, ctx->stream>>>(X, n, d, C, labels, out);
    LAUNCH_CHECK();
    return B2K_OK;
}

}  // namespace b2k
