// lloyd.cu -- centroid update, cost and data statistics (K3 of SURVEY 2.2).
//
// Replaces the accumulate half of deeptime kmeans.cluster and costAssignFunction (call site
// pyemma/coordinates/clustering/kmeans.py:254-258).  The reference sums members in fp32 in
// frame order (serial) or in arbitrary order inside `omp critical` (threaded, non-deterministic,
// tests/test_kmeans.py:104-110).  Here every sum is an exact integer sum of fixed-point values
// (x * 2^q rounded to int64): integer addition is associative, so the result does not depend on
// atomic ordering, on the launch geometry or on how many GPUs the frames are sharded over, and
// the int64 buffer is what the ranks all-reduce.
#include "common.cuh"
#include "kernels.h"

namespace b2k {

__global__ void __launch_bounds__(256) accumulate_kernel(const float* __restrict__ X, int64_t n, int d, int k,
                                                         const int32_t* __restrict__ labels, double scale,
                                                         unsigned long long* __restrict__ acc) {
    const int64_t total = n * d;
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (int64_t)gridDim.x * 256) {
        const int64_t i = e / d;
        const int dim = (int)(e - i * d);
        const int32_t a = labels[i];
        if (a < 0 || a >= k) continue;
        const long long v = __double2ll_rn((double)X[e] * scale);
        atomicAdd(acc + (int64_t)a * d + dim, (unsigned long long)v);
        if (dim == 0) atomicAdd(acc + (int64_t)k * d + a, 1ull);
    }
}

// 32-bit index variant (n*d < 2^31): cheaper integer division
__global__ void __launch_bounds__(256) accumulate_kernel32(const float* __restrict__ X, uint32_t total, uint32_t d,
                                                           int k, const int32_t* __restrict__ labels, double scale,
                                                           unsigned long long* __restrict__ acc) {
    for (uint32_t e = blockIdx.x * 256u + threadIdx.x; e < total; e += gridDim.x * 256u) {
        const uint32_t i = e / d;
        const uint32_t dim = e - i * d;
        const int32_t a = labels[i];
        if (a < 0 || a >= k) continue;
        const long long v = __double2ll_rn((double)X[e] * scale);
        atomicAdd(acc + (size_t)a * d + dim, (unsigned long long)v);
        if (dim == 0) atomicAdd(acc + (size_t)k * d + a, 1ull);
    }
}

// ---- shared-memory member sums (small k*d) -----------------------------------------------------------------------
// When the whole [k*d sums | k counts] table fits shared memory, every CTA keeps a private copy, streams its share
// of X once (coalesced, no gather) and flushes one 64-bit RED per non-zero entry at the end.  Shared memory has no
// native 64-bit add (it compiles to a CAS loop), so a sum is kept as (lo: u32, hi: i32) with the carry of the low
// word detected from the value the 32-bit atomic returns -- still an exact integer sum, order independent.
__global__ void __launch_bounds__(1024, 1) accumulate_smem_kernel(const float* __restrict__ X, int64_t n, int d, int k,
                                                                  const int32_t* __restrict__ labels, double scale,
                                                                  unsigned long long* __restrict__ acc) {
    extern __shared__ uint32_t ash[];
    const int kd = k * d;
    uint32_t* slo = ash;
    uint32_t* shi = ash + kd;
    uint32_t* scnt = ash + 2 * kd;
    for (int t = threadIdx.x; t < 2 * kd + k; t += blockDim.x) ash[t] = 0u;
    __syncthreads();
    const int64_t total = n * d;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / d;
        const int dim = (int)(e - i * d);
        const int32_t a = labels[i];
        if (a < 0 || a >= k) continue;
        const long long v = __double2ll_rn((double)X[e] * scale);
        const uint32_t lo = (uint32_t)(unsigned long long)v;
        uint32_t hi = (uint32_t)(int32_t)(v >> 32);
        const uint32_t old = atomicAdd(&slo[a * d + dim], lo);
        if (old + lo < old) hi += 1u;  // the low word wrapped
        if (hi) atomicAdd(&shi[a * d + dim], hi);
        if (dim == 0) atomicAdd(&scnt[a], 1u);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < kd; t += blockDim.x) {
        const long long sv = ((long long)(int32_t)shi[t] << 32) + (long long)slo[t];
        if (sv != 0) atomicAdd(acc + t, (unsigned long long)sv);
    }
    for (int t = threadIdx.x; t < k; t += blockDim.x)
        if (scnt[t]) atomicAdd(acc + (size_t)kd + t, (unsigned long long)scnt[t]);
}

// ---- segmented member sums ------------------------------------------------------------------------------
// One 64-bit RED per frame element (above) is bound by the L2 atomic units (measured 1.5e11 RED/s: 0.67 ms for
// 1e7 x 10, 5 ms for 1.25e7 x 64), far above the HBM time of the same pass.  Instead the frames are bucketed by
// label (counting sort of the frame indices: histogram -> scan -> scatter) and every warp sums a run of the
// sorted order in registers, issuing one RED per (label run, dimension).  All sums are exact integers, so the
// result is still independent of the scatter order, of the launch geometry and of the number of ranks.
static constexpr int SEG_TABLE_MAX = 49152;   // labels whose cursor table fits shared memory (192 KB)
static constexpr int SEG_POS_PER_WARP = 256;  // sorted positions summed by one warp

__global__ void __launch_bounds__(256) seg_hist_kernel(const int32_t* __restrict__ labels, int64_t n, int k,
                                                       uint32_t* __restrict__ hist, int use_smem) {
    extern __shared__ uint32_t sh[];
    if (use_smem) {
        for (int j = threadIdx.x; j < k; j += 256) sh[j] = 0;
        __syncthreads();
    }
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const int32_t a = labels[i];
        if (a < 0 || a >= k) continue;
        if (use_smem) atomicAdd(&sh[a], 1u);
        else atomicAdd(&hist[a], 1u);
    }
    if (use_smem) {
        __syncthreads();
        for (int j = threadIdx.x; j < k; j += 256) {
            const uint32_t c = sh[j];
            if (c) atomicAdd(&hist[j], c);
        }
    }
}

// seg[0..k] = exclusive scan of hist, cursor = seg, member counts added to the exchange buffer (one block)
__global__ void __launch_bounds__(1024) seg_scan_kernel(const uint32_t* __restrict__ hist, int k, uint32_t* __restrict__ seg,
                                                        uint32_t* __restrict__ cursor,
                                                        unsigned long long* __restrict__ acc_counts) {
    __shared__ uint32_t part[1024];
    const int per = (k + 1023) / 1024;
    const int j0 = threadIdx.x * per, j1 = min(k, j0 + per);
    uint32_t s = 0;
    for (int j = j0; j < j1; ++j) s += hist[j];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan of the per-thread totals
        const uint32_t v = threadIdx.x >= o ? part[threadIdx.x - o] : 0u;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = part[threadIdx.x] - s;
    for (int j = j0; j < j1; ++j) {
        const uint32_t c = hist[j];
        seg[j] = run;
        cursor[j] = run;
        if (c && acc_counts) atomicAdd(acc_counts + j, (unsigned long long)c);
        run += c;
    }
    if (threadIdx.x == 1023) seg[k] = part[1023];
}

__global__ void __launch_bounds__(256) seg_scatter_kernel(const int32_t* __restrict__ labels, int64_t n, int k,
                                                          uint32_t* __restrict__ cursor, uint32_t* __restrict__ perm) {
    const int lane = threadIdx.x & 31;
    const int64_t n_round = (n + 31) & ~(int64_t)31;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n_round; i += (int64_t)gridDim.x * 256) {
        const int32_t a = i < n ? labels[i] : -1;
        const bool ok = a >= 0 && a < k;
        // frames of a trajectory are time-correlated: neighbours often share a label -> one atomic per distinct label
        const unsigned peers = __match_any_sync(0xffffffffu, ok ? a : -1 - lane);
        if (!ok) continue;
        const int leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(&cursor[a], (uint32_t)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        perm[base + __popc(peers & ((1u << lane) - 1u))] = (uint32_t)i;
    }
}

// Two-level counting sort without global atomics (k labels fit a shared-memory table): every CTA owns a contiguous
// range of frames.  pass 1: per-CTA label histogram -> hist[cta][k]; column scan over the CTAs + label scan give
// every (CTA, label) its first output position; pass 2: the CTA re-reads its labels and hands out positions from
// a shared-memory cursor table (shared atomics only).
__global__ void __launch_bounds__(256) seg_hist2_kernel(const int32_t* __restrict__ labels, int64_t n, int k,
                                                        int64_t per_cta, uint32_t* __restrict__ hist) {
    extern __shared__ uint32_t sh[];
    for (int j = threadIdx.x; j < k; j += 256) sh[j] = 0;
    __syncthreads();
    const int64_t c0 = (int64_t)blockIdx.x * per_cta, c1 = min(n, c0 + per_cta);
    for (int64_t i = c0 + threadIdx.x; i < c1; i += 256) {
        const int32_t a = labels[i];
        if (a >= 0 && a < k) atomicAdd(&sh[a], 1u);
    }
    __syncthreads();
    uint32_t* row = hist + (size_t)blockIdx.x * k;
    for (int j = threadIdx.x; j < k; j += 256) row[j] = sh[j];
}

// thread per label: exclusive scan down the CTA column (in place), column total -> total[j]
__global__ void __launch_bounds__(256) seg_colscan_kernel(uint32_t* __restrict__ hist, int n_cta, int k,
                                                          uint32_t* __restrict__ total) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= k) return;
    uint32_t run = 0;
    for (int c0 = 0; c0 < n_cta; c0 += 16) {  // 16 independent loads in flight, then the short serial prefix
        uint32_t v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = (c0 + u < n_cta) ? hist[(size_t)(c0 + u) * k + j] : 0u;
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (c0 + u < n_cta) hist[(size_t)(c0 + u) * k + j] = run;
            run += v[u];
        }
    }
    total[j] = run;
}

__global__ void __launch_bounds__(256) seg_scatter2_kernel(const int32_t* __restrict__ labels, int64_t n, int k,
                                                           int64_t per_cta, const uint32_t* __restrict__ hist,
                                                           const uint32_t* __restrict__ seg, uint32_t* __restrict__ perm) {
    extern __shared__ uint32_t sh[];
    const uint32_t* row = hist + (size_t)blockIdx.x * k;
    for (int j = threadIdx.x; j < k; j += 256) sh[j] = seg[j] + row[j];
    __syncthreads();
    const int64_t c0 = (int64_t)blockIdx.x * per_cta, c1 = min(n, c0 + per_cta);
    for (int64_t i = c0 + threadIdx.x; i < c1; i += 256) {
        const int32_t a = labels[i];
        if (a >= 0 && a < k) perm[atomicAdd(&sh[a], 1u)] = (uint32_t)i;
    }
}

// LPF lanes share one frame (lane dl owns dimensions dl, dl+LPF, ... of a block of LPF*NACC dimensions),
// 32/LPF frames per warp step.
template <int LPF, int NACC>
__global__ void __launch_bounds__(256) seg_sum_kernel(const float* __restrict__ X, int d, int k,
                                                      const uint32_t* __restrict__ seg, const uint32_t* __restrict__ perm,
                                                      double scale, unsigned long long* __restrict__ acc) {
    constexpr int FPW = 32 / LPF;
    constexpr int UNR = 4;
    const int lane = threadIdx.x & 31;
    const int dl = lane % LPF, f = lane / LPF;
    const uint32_t n_valid = seg[k];
    const int64_t warp = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * 256) >> 5;
    for (int64_t w = warp; w * SEG_POS_PER_WARP < n_valid; w += n_warps) {
        const uint32_t p0 = (uint32_t)(w * SEG_POS_PER_WARP);
        const uint32_t p1 = (uint32_t)min((int64_t)n_valid, (int64_t)p0 + SEG_POS_PER_WARP);
        // label whose segment holds p0: largest a with seg[a] <= p0
        int lo = 0, hi = k;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (seg[mid] <= p0) lo = mid; else hi = mid;
        }
        for (int dim0 = 0; dim0 < d; dim0 += LPF * NACC) {
            int a = lo;
            uint32_t p = p0;
            while (p < p1) {
                while (seg[a + 1] <= p) ++a;
                const uint32_t r1 = min(seg[a + 1], p1);
                long long s[NACC];
#pragma unroll
                for (int j = 0; j < NACC; ++j) s[j] = 0;
                for (uint32_t q = p + f; q < r1; q += FPW * UNR) {
                    uint32_t idx[UNR];
#pragma unroll
                    for (int u = 0; u < UNR; ++u) idx[u] = (q + u * FPW < r1) ? perm[q + u * FPW] : 0xffffffffu;
                    float v[UNR][NACC];
#pragma unroll
                    for (int u = 0; u < UNR; ++u) {
#pragma unroll
                        for (int j = 0; j < NACC; ++j) {
                            const int dim = dim0 + j * LPF + dl;
                            v[u][j] = (idx[u] != 0xffffffffu && dim < d) ? __ldg(X + (int64_t)idx[u] * d + dim) : 0.f;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < UNR; ++u)
#pragma unroll
                        for (int j = 0; j < NACC; ++j) s[j] += __double2ll_rn((double)v[u][j] * scale);
                }
#pragma unroll
                for (int j = 0; j < NACC; ++j) {
#pragma unroll
                    for (int o = LPF; o < 32; o <<= 1) s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
                    const int dim = dim0 + j * LPF + dl;
                    if (f == 0 && dim < d && s[j] != 0)
                        atomicAdd(acc + (int64_t)a * d + dim, (unsigned long long)s[j]);
                }
                p = r1;
            }
        }
    }
}

// ---- tile-sorted member sums (narrow rows) ---------------------------------------------------------------------
// For rows shorter than two DRAM sectors the global gather above re-fetches every sector about three times (ncu at
// 1e7 x 10: 1.38 GB read for 0.44 GB of frames).  Here a CTA sorts the labels of one TILE of consecutive frames in
// shared memory (histogram, scan, scatter of 16-bit local indices) and sums the label runs of that tile only: the
// tile's rows (a few hundred KB) are read once from DRAM and re-used out of L1/L2, at the price of one RED per
// (label present in the tile, dimension) instead of one per (label run of the whole shard, dimension).
static constexpr int ST_TILE = 8192;      // frames per tile (16-bit local indices)
static constexpr int ST_KMAX = 8192;      // labels whose per-tile tables fit shared memory

template <int LPF>
__global__ void __launch_bounds__(256) seg_tile_kernel(const float* __restrict__ X, int64_t n, int d, int k,
                                                       const int32_t* __restrict__ labels, double scale,
                                                       unsigned long long* __restrict__ acc) {
    extern __shared__ uint32_t tsh[];
    uint32_t* seg = tsh;                 // [k+1] first sorted position of every label in this tile
    uint32_t* cursor = tsh + (k + 1);    // [k]   histogram, then scatter cursor
    uint16_t* sidx = reinterpret_cast<uint16_t*>(cursor + k);  // [ST_TILE] local frame index by sorted position
    __shared__ uint32_t wsum[8];
    constexpr int FPW = 32 / LPF;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int dl = lane % LPF, f = lane / LPF;
    const int per = (k + 255) / 256;  // labels scanned by one thread
    for (int64_t base = (int64_t)blockIdx.x * ST_TILE; base < n; base += (int64_t)gridDim.x * ST_TILE) {
        const int nt = (int)min((int64_t)ST_TILE, n - base);
        __syncthreads();  // previous tile done with seg / sidx
        for (int j = threadIdx.x; j < k; j += 256) cursor[j] = 0u;
        __syncthreads();
        for (int i = threadIdx.x; i < nt; i += 256) {
            const int32_t a = labels[base + i];
            if (a >= 0 && a < k) atomicAdd(&cursor[a], 1u);
        }
        __syncthreads();
        {   // exclusive scan of the histogram: thread t owns labels [t*per, (t+1)*per)
            const int j0 = threadIdx.x * per, j1 = min(k, j0 + per);
            uint32_t s = 0;
            for (int j = j0; j < j1; ++j) s += cursor[j];
            uint32_t inc = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            if (lane == 31) wsum[warp] = inc;
            __syncthreads();
            uint32_t wbase = 0;
            for (int w = 0; w < warp; ++w) wbase += wsum[w];
            uint32_t run = wbase + inc - s;
            for (int j = j0; j < j1; ++j) {
                const uint32_t c = cursor[j];
                seg[j] = run;
                cursor[j] = run;
                if (c) atomicAdd(acc + (size_t)k * d + j, (unsigned long long)c);
                run += c;
            }
            if (threadIdx.x == 255) seg[k] = run;  // labels beyond 256*per do not exist: thread 255 ends at k
        }
        __syncthreads();
        for (int i = threadIdx.x; i < nt; i += 256) {
            const int32_t a = labels[base + i];
            if (a >= 0 && a < k) sidx[atomicAdd(&cursor[a], 1u)] = (uint16_t)i;
        }
        __syncthreads();
        const uint32_t n_valid = seg[k];
        // run sums: warp w takes chunks of 256 sorted positions
        for (uint32_t p0 = warp * 256u; p0 < n_valid; p0 += 8u * 256u) {
            const uint32_t p1 = min(n_valid, p0 + 256u);
            int lo = 0, hi = k;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (seg[mid] <= p0) lo = mid; else hi = mid;
            }
            int a = lo;
            uint32_t p = p0;
            while (p < p1) {
                while (seg[a + 1] <= p) ++a;
                const uint32_t r1 = min(seg[a + 1], p1);
                long long s = 0;
                for (uint32_t q = p + f; q < r1; q += FPW * 4) {
                    float v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint32_t qq = q + u * FPW;
                        v[u] = (qq < r1 && dl < d) ? __ldg(X + (base + sidx[qq]) * d + dl) : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) s += __double2ll_rn((double)v[u] * scale);
                }
#pragma unroll
                for (int o = LPF; o < 32; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (f == 0 && dl < d && s != 0) atomicAdd(acc + (int64_t)a * d + dl, (unsigned long long)s);
                p = r1;
            }
        }
    }
}

__global__ void finalize_kernel(const long long* __restrict__ acc, int k, int d, double inv_scale,
                                const float* __restrict__ old_c, float* __restrict__ new_c) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= k * d) return;
    const int c = e / d;
    const long long cnt = acc[(size_t)k * d + c];
    if (cnt == 0) new_c[e] = old_c[e];  // empty cluster keeps its old center (deeptime kmeans.cluster)
    else new_c[e] = (float)(((double)acc[e] * inv_scale) / (double)cnt);
}

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// acc_slot += sum_i round((l_i*l_i) * scale)   (l*l in fp32 like the reference's `value += l*l`)
__global__ void __launch_bounds__(256) cost_reduce_kernel(const float* __restrict__ l, int64_t n, double scale,
                                                          unsigned long long* __restrict__ acc_slot) {
    __shared__ long long part[8];
    long long s = 0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const float li = l[i];
        s += __double2ll_rn((double)__fmul_rn(li, li) * scale);
    }
    s = warp_sum_ll(s);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        long long t = threadIdx.x < 8 ? part[threadIdx.x] : 0;
        t = warp_sum_ll(t);
        if (threadIdx.x == 0) atomicAdd(acc_slot, (unsigned long long)t);
    }
}

// Narrow rows (d <= 16, center table in shared memory): distance of every frame to ITS center and the fixed-point cost sum in
// one pass -- the arithmetic of labeled_dist_kernel + cost_reduce_kernel (sqrt, then l*l in fp32, then the exact integer
// sum), without the round trip of the per-frame distances through HBM.  Frame and table rows are zero padded to DREG
// columns; a padded column adds (0-0)^2 = +0 to lane 0 of the reference's 4-lane sum, which leaves its bits alone.
template <int DREG>
__global__ void __launch_bounds__(256) cost_fused_small_kernel(const float* __restrict__ X, int64_t n, int d,
                                                               const float* __restrict__ Cn, int k,
                                                               const int32_t* __restrict__ labels, double scale,
                                                               unsigned long long* __restrict__ acc_slot, int vec) {
    extern __shared__ __align__(16) float ctab[];  // k rows, stride RS floats = an odd number of 16-byte units
    __shared__ long long part[8];
    constexpr int RS = ((DREG / 4) & 1) ? DREG : DREG + 4;
    for (int t = threadIdx.x; t < k * DREG; t += 256) {
        const int r = t / DREG, c = t - r * DREG;
        ctab[r * RS + c] = c < d ? __ldg(Cn + (int64_t)r * d + c) : 0.f;
    }
    __syncthreads();
    const bool full_last = d == DREG;  // d is in (DREG-4, DREG]
    long long s = 0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        float xr[DREG];
        load_row_padded<DREG>(X, i, d, vec, xr);
        const int32_t a = labels[i];
        if ((uint32_t)a >= (uint32_t)k) continue;  // caller-supplied labels outside [0, k): no center, no contribution
        const float4* c4 = reinterpret_cast<const float4*>(ctab + (size_t)a * RS);
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
        const float li = __fsqrt_rn(L.result());
        s += __double2ll_rn((double)__fmul_rn(li, li) * scale);
    }
    s = warp_sum_ll(s);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        long long t = threadIdx.x < 8 ? part[threadIdx.x] : 0;
        t = warp_sum_ll(t);
        if (threadIdx.x == 0) atomicAdd(acc_slot, (unsigned long long)t);
    }
}

__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ X, int64_t count, int* out_bits) {
    float m = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < count; i += (int64_t)gridDim.x * 256) {
        const float v = fabsf(X[i]);
        if (v > m) m = v;  // NaN never wins; +inf does
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_int(m));
}

__global__ void __launch_bounds__(256) all_finite_kernel(const float* __restrict__ X, int64_t count, int* flag) {
    bool ok = true;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < count; i += (int64_t)gridDim.x * 256) {
        const uint32_t b = __float_as_uint(X[i]);
        if ((b & 0x7f800000u) == 0x7f800000u) ok = false;
    }
    if (!__all_sync(0xffffffffu, ok) && (threadIdx.x & 31) == 0) *flag = 0;
}

// fp32 CUDA-core issue rate ("the builder must measure it"): 16 independent chains of the path's
// own instruction mix -- FMUL and FADD that may not fuse -- per thread, all SMs resident
__global__ void __launch_bounds__(256) fp32_rate_kernel(float* out, int iters, float c, float e) {
    float a[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) a[q] = (float)(threadIdx.x + q) * 1e-3f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 16; ++q) a[q] = __fadd_rn(__fmul_rn(a[q], c), e);
    }
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 16; ++q) s += a[q];
    if (s == 12345.678f) out[0] = s;  // never true: keeps the chains alive
}

int measure_fp32_rate(b2k_ctx* ctx, double* lane_instr_per_s) {
    B2K_TRY(ctx->ensure_scratch(64));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    const int iters = 20000, blocks = ctx->sm_count * 8;
    fp32_rate_kernel<<<blocks, 256, 0, ctx->stream>>>((float*)ctx->scratch, 200, 0.999f, 1e-3f);  // warm-up
    CUDA_TRY(cudaEventRecord(e0, ctx->stream));
    fp32_rate_kernel<<<blocks, 256, 0, ctx->stream>>>((float*)ctx->scratch, iters, 0.999f, 1e-3f);
    CUDA_TRY(cudaEventRecord(e1, ctx->stream));
    LAUNCH_CHECK();
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *lane_instr_per_s = (double)blocks * 256.0 * iters * 32.0 / (ms * 1e-3);
    return B2K_OK;
}

static unsigned grid_for(b2k_ctx* ctx, int64_t work_items, int per_block) {
    int64_t b = cdiv(work_items, per_block);
    const int64_t cap = (int64_t)ctx->sm_count * 8;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

int launch_accumulate(b2k_ctx* ctx, const float* X, int64_t n, int d, int k, const int32_t* labels, double scale,
                      int64_t* acc) {
    if (n <= 0) return B2K_OK;
    ProfScope prof(ctx, b2k_ctx::PROF_SUMS);
    const int64_t total = n * d;
    const size_t table_bytes = ((size_t)2 * k * d + k) * 4;
    // automatic: the table kernel only in the launch-latency regime (one kernel instead of four); measured at
    // 1e7 x 10, k=1000 the shared atomics (2 per element) cost 0.66 ms against 0.41 ms for the segmented path
    const bool small_job = total <= (int64_t(1) << 22);
    if (((ctx->accumulate_mode == 0 && small_job) || ctx->accumulate_mode == 3) && total >= (int64_t(1) << 16) &&
        table_bytes <= (size_t)200 * 1024 && (int64_t)k * d < (int64_t(1) << 24)) {
        static PerDeviceOnce attr_set;
        if (attr_set.need(ctx->device)) {
            CUDA_TRY(cudaFuncSetAttribute(accumulate_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_set.done(ctx->device);
        }
        const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv(total, 1024 * 8), ctx->sm_count));
        accumulate_smem_kernel<<<grid, 1024, table_bytes, ctx->stream>>>(X, n, d, k, labels, scale, (unsigned long long*)acc);
        LAUNCH_CHECK();
        return B2K_OK;
    }
    // explicit only: measured at 1e7 x 10, k=1000 the per-tile sort phases (shared atomics, five barriers per tile)
    // cost 0.50 ms against 0.41 ms for the global segmented path, even though the DRAM over-fetch is gone
    if (ctx->accumulate_mode == 4 && total >= (int64_t(1) << 16) && d <= 16 && k <= ST_KMAX) {
        const size_t smem = ((size_t)2 * k + 1) * 4 + (size_t)ST_TILE * 2 + 16;
        const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv(n, ST_TILE), (int64_t)ctx->sm_count * 4));
        unsigned long long* a64 = (unsigned long long*)acc;
#define B2K_ST(LPF)                                                                                          \
    do {                                                                                                     \
        static PerDeviceOnce at;                                                                              \
        if (at.need(ctx->device)) {                                                                                           \
            CUDA_TRY(cudaFuncSetAttribute(seg_tile_kernel<LPF>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          (2 * ST_KMAX + 1) * 4 + ST_TILE * 2 + 16));                        \
            at.done(ctx->device);                                                                                       \
        }                                                                                                    \
        seg_tile_kernel<LPF><<<grid, 256, smem, ctx->stream>>>(X, n, d, k, labels, scale, a64);              \
    } while (0)
        if (d <= 1) B2K_ST(1);
        else if (d <= 2) B2K_ST(2);
        else if (d <= 4) B2K_ST(4);
        else if (d <= 8) B2K_ST(8);
        else B2K_ST(16);
#undef B2K_ST
        LAUNCH_CHECK();
        return B2K_OK;
    }
    if (ctx->accumulate_mode != 1 && total >= (int64_t(1) << 16) && n < (int64_t(1) << 32) - 1 && k < (1 << 30)) {
        // segmented path: scratch = hist[k] | seg[k+1] | cursor[k] | perm[n]
        const size_t words = (size_t)3 * k + 4 + (size_t)n;
        B2K_TRY(ctx->ensure_scratch(words * 4 + 64));
        uint32_t* hist = reinterpret_cast<uint32_t*>((char*)ctx->scratch + 64);
        uint32_t* seg = hist + k;
        uint32_t* cursor = seg + k + 1;
        uint32_t* perm = cursor + k + 2;
        cudaStream_t st = ctx->stream;
        if (k <= SEG_TABLE_MAX) {
            int n_cta = (int)std::min<int64_t>(std::min<int64_t>(cdiv(n, 4096), (int64_t)ctx->sm_count * 4),
                                               std::max<int64_t>(1, (int64_t(8) << 20) / k));
            if (n_cta < 1) n_cta = 1;
            const int64_t per_cta = cdiv(cdiv(n, n_cta), 256) * 256;
            n_cta = (int)cdiv(n, per_cta);
            B2K_TRY(ctx->ensure_scratch2((size_t)n_cta * k * 4));
            uint32_t* hist2 = reinterpret_cast<uint32_t*>(ctx->scratch2);
            static PerDeviceOnce attr_set;
            if (attr_set.need(ctx->device)) {
                CUDA_TRY(cudaFuncSetAttribute(seg_hist2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SEG_TABLE_MAX * 4));
                CUDA_TRY(cudaFuncSetAttribute(seg_scatter2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SEG_TABLE_MAX * 4));
                attr_set.done(ctx->device);
            }
            seg_hist2_kernel<<<n_cta, 256, (size_t)k * 4, st>>>(labels, n, k, per_cta, hist2);
            LAUNCH_CHECK();
            seg_colscan_kernel<<<(unsigned)cdiv(k, 256), 256, 0, st>>>(hist2, n_cta, k, hist);
            LAUNCH_CHECK();
            seg_scan_kernel<<<1, 1024, 0, st>>>(hist, k, seg, cursor, (unsigned long long*)acc + (int64_t)k * d);
            LAUNCH_CHECK();
            seg_scatter2_kernel<<<n_cta, 256, (size_t)k * 4, st>>>(labels, n, k, per_cta, hist2, seg, perm);
            LAUNCH_CHECK();
        } else {
            CUDA_TRY(cudaMemsetAsync(hist, 0, (size_t)k * 4, st));
            seg_hist_kernel<<<grid_for(ctx, n, 2048), 256, 0, st>>>(labels, n, k, hist, 0);
            LAUNCH_CHECK();
            seg_scan_kernel<<<1, 1024, 0, st>>>(hist, k, seg, cursor, (unsigned long long*)acc + (int64_t)k * d);
            LAUNCH_CHECK();
            seg_scatter_kernel<<<grid_for(ctx, n, 1024), 256, 0, st>>>(labels, n, k, cursor, perm);
            LAUNCH_CHECK();
        }
        const unsigned sgrid = (unsigned)std::max<int64_t>(
            1, std::min<int64_t>(cdiv(n, (int64_t)SEG_POS_PER_WARP * 8), (int64_t)ctx->sm_count * 8));
        unsigned long long* a64 = (unsigned long long*)acc;
#define B2K_SEG(LPF, NACC) seg_sum_kernel<LPF, NACC><<<sgrid, 256, 0, st>>>(X, d, k, seg, perm, scale, a64)
        if (d <= 1) B2K_SEG(1, 1);
        else if (d <= 2) B2K_SEG(2, 1);
        else if (d <= 4) B2K_SEG(4, 1);
        else if (d <= 8) B2K_SEG(8, 1);
        else if (d <= 16) B2K_SEG(16, 1);
        else if (d <= 32) B2K_SEG(32, 1);
        else if (d <= 64) B2K_SEG(32, 2);
        else if (d <= 128) B2K_SEG(32, 4);
        else B2K_SEG(32, 8);
#undef B2K_SEG
        LAUNCH_CHECK();
        return B2K_OK;
    }
    if (total < (int64_t(1) << 31))
        accumulate_kernel32<<<grid_for(ctx, total, 256), 256, 0, ctx->stream>>>(
            X, (uint32_t)total, (uint32_t)d, k, labels, scale, (unsigned long long*)acc);
    else
        accumulate_kernel<<<grid_for(ctx, total, 256), 256, 0, ctx->stream>>>(X, n, d, k, labels, scale,
                                                                             (unsigned long long*)acc);
    LAUNCH_CHECK();
    return B2K_OK;
}

// ---- incremental member sums ------------------------------------------------------------------------------------
// The sums are exact integers (fixed point), so they can be UPDATED instead of recomputed: a frame whose label did not
// change contributes the same addends to the same center as in the previous iteration.  After the first iterations of a
// Lloyd run a few per cent of the frames change their label; only those rows are read: -q(x) from the old center's
// sums, +q(x) to the new one's, counts alike.  The result is the same integer a full pass produces, bit for bit.
// One warp per 32 frames: lanes compare old and new label, then the warp walks the changed frames, lanes over dimensions.
__global__ void __launch_bounds__(256) accumulate_delta_kernel(const float* __restrict__ X, int64_t n, int d, int k,
                                                               const int32_t* __restrict__ old_l,
                                                               const int32_t* __restrict__ new_l, double scale,
                                                               unsigned long long* __restrict__ acc,
                                                               unsigned long long* __restrict__ changed) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * 256) >> 5;
    unsigned long long mine = 0;
    for (int64_t base = warp * 32; base < n; base += n_warps * 32) {
        const int64_t i = base + lane;
        int32_t o = -1, w = -1;
        if (i < n) { o = old_l[i]; w = new_l[i]; }
        unsigned m = __ballot_sync(0xffffffffu, o != w);
        if (lane == 0) mine += __popc(m);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const int32_t oo = __shfl_sync(0xffffffffu, o, src), ww = __shfl_sync(0xffffffffu, w, src);
            const float* row = X + (base + src) * d;
            const bool ok_o = oo >= 0 && oo < k, ok_w = ww >= 0 && ww < k;
            for (int e = lane; e < d; e += 32) {
                const long long q = __double2ll_rn((double)__ldg(row + e) * scale);
                if (q != 0) {
                    if (ok_w) atomicAdd(acc + (int64_t)ww * d + e, (unsigned long long)q);
                    if (ok_o) atomicAdd(acc + (int64_t)oo * d + e, (unsigned long long)(-q));
                }
            }
            if (lane == 0) {
                if (ok_w) atomicAdd(acc + (int64_t)k * d + ww, 1ull);
                if (ok_o) atomicAdd(acc + (int64_t)k * d + oo, (unsigned long long)(-1ll));
            }
        }
    }
    if (lane == 0 && mine) atomicAdd(changed, mine);
}

__global__ void __launch_bounds__(256) count_changed_kernel(const int32_t* __restrict__ old_l, const int32_t* __restrict__ new_l,
                                                            int64_t n, unsigned long long* __restrict__ changed) {
    unsigned long long mine = 0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) mine += old_l[i] != new_l[i];
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(changed, mine);
}

// acc (sums of the frames under old_l) becomes the sums under new_l; *changed += frames whose label differs
int launch_accumulate_delta(b2k_ctx* ctx, const float* X, int64_t n, int d, int k, const int32_t* old_l, const int32_t* new_l,
                            double scale, int64_t* acc, unsigned long long* changed) {
    if (n <= 0) return B2K_OK;
    ProfScope prof(ctx, b2k_ctx::PROF_SUMS);
    accumulate_delta_kernel<<<grid_for(ctx, n, 2048), 256, 0, ctx->stream>>>(X, n, d, k, old_l, new_l, scale,
                                                                            (unsigned long long*)acc, changed);
    LAUNCH_CHECK();
    return B2K_OK;
}

int launch_count_changed(b2k_ctx* ctx, const int32_t* old_l, const int32_t* new_l, int64_t n, unsigned long long* changed) {
    if (n <= 0) return B2K_OK;
    ProfScope prof(ctx, b2k_ctx::PROF_SUMS);
    count_changed_kernel<<<grid_for(ctx, n, 4096), 256, 0, ctx->stream>>>(old_l, new_l, n, changed);
    LAUNCH_CHECK();
    return B2K_OK;
}

// stable counting sort of the frame indices by label: perm[seg[a] .. seg[a+1]) = frames with label a, ascending;
// frames whose label is outside [0, k) are left out (seg[k] = number of sorted frames)
int launch_label_sort(b2k_ctx* ctx, const int32_t* labels, int64_t n, int k, uint32_t* seg, uint32_t* perm) {
    if (n <= 0 || n >= (int64_t(1) << 32) - 1 || k > SEG_TABLE_MAX)
        return set_error(B2K_ERR_INVALID_ARG, "label sort: unsupported size");
    cudaStream_t st = ctx->stream;
    B2K_TRY(ctx->ensure_scratch((size_t)(2 * k + 8) * 4 + 64));
    uint32_t* hist = reinterpret_cast<uint32_t*>((char*)ctx->scratch + 64);
    uint32_t* cursor = hist + k;
    int n_cta = (int)std::min<int64_t>(std::min<int64_t>(cdiv(n, 4096), (int64_t)ctx->sm_count * 4),
                                       std::max<int64_t>(1, (int64_t(8) << 20) / k));
    if (n_cta < 1) n_cta = 1;
    const int64_t per_cta = cdiv(cdiv(n, n_cta), 256) * 256;
    n_cta = (int)cdiv(n, per_cta);
    B2K_TRY(ctx->ensure_scratch2((size_t)n_cta * k * 4));
    uint32_t* hist2 = reinterpret_cast<uint32_t*>(ctx->scratch2);
    static PerDeviceOnce attr_set;
    if (attr_set.need(ctx->device)) {
        CUDA_TRY(cudaFuncSetAttribute(seg_hist2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SEG_TABLE_MAX * 4));
        CUDA_TRY(cudaFuncSetAttribute(seg_scatter2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SEG_TABLE_MAX * 4));
        attr_set.done(ctx->device);
    }
    seg_hist2_kernel<<<n_cta, 256, (size_t)k * 4, st>>>(labels, n, k, per_cta, hist2);
    LAUNCH_CHECK();
    seg_colscan_kernel<<<(unsigned)cdiv(k, 256), 256, 0, st>>>(hist2, n_cta, k, hist);
    LAUNCH_CHECK();
    seg_scan_kernel<<<1, 1024, 0, st>>>(hist, k, seg, cursor, nullptr);
    LAUNCH_CHECK();
    seg_scatter2_kernel<<<n_cta, 256, (size_t)k * 4, st>>>(labels, n, k, per_cta, hist2, seg, perm);
    LAUNCH_CHECK();
    return B2K_OK;
}

int launch_finalize(b2k_ctx* ctx, const int64_t* acc, int k, int d, double inv_scale, const float* old_centers,
                    float* new_centers) {
    const int total = k * d;
    finalize_kernel<<<(unsigned)cdiv(total, 256), 256, 0, ctx->stream>>>((const long long*)acc, k, d, inv_scale,
                                                                        old_centers, new_centers);
    LAUNCH_CHECK();
    return B2K_OK;
}

int launch_cost_reduce(b2k_ctx* ctx, const float* l, int64_t n, double scale, int64_t* acc_slot) {
    if (n <= 0) return B2K_OK;
    cost_reduce_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(l, n, scale, (unsigned long long*)acc_slot);
    LAUNCH_CHECK();
    return B2K_OK;
}

// returns B2K_OK with *done = 1 when the fused narrow-row kernel took the job, *done = 0 when the caller has to run the
// two-pass path (wide rows, table too large for shared memory, labels possibly out of range)
int launch_cost_fused(b2k_ctx* ctx, const float* X, int64_t n, int d, const float* C, int k, const int32_t* labels,
                      double scale, int64_t* acc_slot, int* done) {
    *done = 0;
    if (n <= 0 || d > 16 || ctx->cost_kernel == 2) return B2K_OK;
    const int ds = (d + 3) & ~3;
    const size_t tbytes = (size_t)k * (((ds / 4) & 1) ? ds : ds + 4) * 4;
    if (tbytes > 96 * 1024) return B2K_OK;
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (200 * 1024) / tbytes));
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * per_sm));
#define B2K_COSTF(DR)                                                                                                  \
    do {                                                                                                               \
        static PerDeviceOnce cattr;                                                                                    \
        if (cattr.need(ctx->device)) {                                                                                 \
            CUDA_TRY(cudaFuncSetAttribute(cost_fused_small_kernel<DR>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                          96 * 1024));                                                                 \
            cattr.done(ctx->device);                                                                                   \
        }                                                                                                              \
        cost_fused_small_kernel<DR><<<grid, 256, tbytes, ctx->stream>>>(X, n, d, C, k, labels, scale,                  \
                                                                        (unsigned long long*)acc_slot,                \
                                                                        row_load_width(X, d, ctx->row_vec_max));      \
    } while (0)
    if (ds == 4) B2K_COSTF(4);
    else if (ds == 8) B2K_COSTF(8);
    else if (ds == 12) B2K_COSTF(12);
    else B2K_COSTF(16);
#undef B2K_COSTF
    LAUNCH_CHECK();
    *done = 1;
    return B2K_OK;
}

int launch_absmax(b2k_ctx* ctx, const float* X, int64_t count, float* d_out) {
    if (count <= 0) return B2K_OK;
    absmax_kernel<<<grid_for(ctx, count, 1024), 256, 0, ctx->stream>>>(X, count, (int*)d_out);
    LAUNCH_CHECK();
    return B2K_OK;
}

int launch_all_finite(b2k_ctx* ctx, const float* X, int64_t count, int* d_flag) {
    if (count <= 0) return B2K_OK;
    all_finite_kernel<<<grid_for(ctx, count, 1024), 256, 0, ctx->stream>>>(X, count, d_flag);
    LAUNCH_CHECK();
    return B2K_OK;
}

}  // namespace b2k
