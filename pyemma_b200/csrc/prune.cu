// prune.cu -- exact center pruning for the Lloyd iterations of a session (K2p).
//
// deeptime's kmeans.cluster (pyemma/coordinates/clustering/kmeans.py:254-258) evaluates all N x k distances in every
// iteration.  Most of them cannot matter: after the first iteration the frames are kept in HBM SORTED BY LABEL, so a
// tile of 128 consecutive frames lives in a small ball B(p_t, R_t) (p_t = tile mean, R_t = max distance to it, both
// computed once per sort).  For a center c_j and the center c_b nearest to p_t the triangle inequality gives, for
// every frame x of the tile,
//        |x - c_j| >= |c_j - p_t| - R_t          |x - c_b| <= |c_b - p_t| + R_t
// so j cannot be the frame's nearest center when |c_j - p_t| > |c_b - p_t| + 2 R_t.  Per iteration one small kernel
// writes the list of surviving centers of every tile (ascending center index, padded with a dummy row); the tensor-
// core screen (screen.cu: screen_gemm_listed_kernel) gathers exactly those rows of the center operand with TMA
// tile::gather4 and the exact verify resolves list positions back to center indices.  The bound carries a relative
// slack of 1e-4 -- three orders of magnitude above every fp32 rounding on the path (the reference's own distance has
// a relative error of (d/4+9) 2^-24) -- so a pruned center is STRICTLY farther than c_b in the reference's arithmetic
// as well: labels, member sums and costs are bit-identical to the unpruned iteration (tests/test_gpu_prune.py).
//
// Wide rows (the k x d center table does not fit shared memory) use the same bound through the center-center
// distances: with a = the label of the tile's first frame, |c_j - p_t| >= |c_j - c_a| - |c_a - p_t|, so j is pruned
// when |c_j - c_a| > 2 (|c_a - p_t| + R_t): O(k^2 d + n_tiles (d + k)) work instead of O(n_tiles k d).
#include "common.cuh"
#include "kernels.h"
#include <algorithm>

namespace b2k {

static constexpr int PT = 128;          // frames per tile (= the screen kernel's TILE_M)
static constexpr float PRUNE_SLACK = 1e-4f;

struct PruneStats {  // device
    unsigned long long total;  // sum of the padded list lengths
    unsigned int max_count;    // longest list
    unsigned int overflow;     // tiles whose list did not fit lcap
};

// Xs[p][:] = X[perm[p]][:]
template <typename V>
__global__ void __launch_bounds__(256) gather_rows_kernel(const V* __restrict__ X, const uint32_t* __restrict__ perm,
                                                          int64_t n, int dv, V* __restrict__ Xs) {
    const int64_t total = n * dv;
    for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
        const int64_t p = t / dv;
        const int e = (int)(t - p * dv);
        Xs[t] = __ldg(X + (int64_t)perm[p] * dv + e);
    }
}

__global__ void __launch_bounds__(256) compose_perm_kernel(const uint32_t* __restrict__ perm_old,
                                                           const uint32_t* __restrict__ sigma, int64_t n,
                                                           uint32_t* __restrict__ perm_new) {
    for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < n; p += (int64_t)gridDim.x * 256)
        perm_new[p] = perm_old[sigma[p]];
}

template <typename T>
__global__ void __launch_bounds__(256) gather_values_kernel(const T* __restrict__ src, const uint32_t* __restrict__ idx,
                                                            int64_t n, T* __restrict__ dst) {
    for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < n; p += (int64_t)gridDim.x * 256) dst[p] = src[idx[p]];
}

__global__ void __launch_bounds__(256) scatter_labels_kernel(const int32_t* __restrict__ labels_s,
                                                             const uint32_t* __restrict__ perm, int64_t n,
                                                             int32_t* __restrict__ out) {
    for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < n; p += (int64_t)gridDim.x * 256)
        out[perm[p]] = labels_s[p];
}

// key[i] = rank[label[i]]: the frames are sorted by the RANK of their label in a kd-tree order of the centers, so that
// labels next to each other in the sorted array are neighbours in space and a tile that straddles two labels stays compact
__global__ void __launch_bounds__(256) label_rank_kernel(const int32_t* __restrict__ labels, const int32_t* __restrict__ rank,
                                                         int64_t n, int k, int32_t* __restrict__ keys) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const int32_t a = labels[i];
        keys[i] = (a >= 0 && a < k) ? __ldg(rank + a) : 0;
    }
}

// kd-tree leaf order of the centers (host; k is a few thousand): split the widest dimension at its median, recurse
static void kd_order(const float* C, int d, int* idx, int lo, int hi) {
    if (hi - lo <= 1) return;
    int best = 0;
    float spread = -1.f;
    for (int e = 0; e < d; ++e) {
        float mn = C[(size_t)idx[lo] * d + e], mx = mn;
        for (int t = lo + 1; t < hi; ++t) {
            const float v = C[(size_t)idx[t] * d + e];
            mn = std::min(mn, v);
            mx = std::max(mx, v);
        }
        if (mx - mn > spread) { spread = mx - mn; best = e; }
    }
    if (!(spread > 0.f)) return;  // identical (or non-finite) centers: any order
    const int mid = (lo + hi) / 2;
    std::nth_element(idx + lo, idx + mid, idx + hi, [&](int a, int b) {
        const float va = C[(size_t)a * d + best], vb = C[(size_t)b * d + best];
        return va < vb || (va == vb && a < b);
    });
    kd_order(C, d, idx, lo, mid);
    kd_order(C, d, idx, mid, hi);
}

// one CTA (128 threads) per LIST UNIT (S = 1 << sshift consecutive tiles share one center list): mean (fixed summation
// order) and radius of the unit's frames
__global__ void __launch_bounds__(PT) tile_meta_kernel(const float* __restrict__ Xs, int64_t n, int d, int sshift,
                                                       float* __restrict__ tmean, float* __restrict__ trad) {
    extern __shared__ __align__(16) float sm[];  // [4][d] partial sums, then [d] mean
    float* part = sm;
    float* mean = sm + 4 * d;
    __shared__ float red[PT / 32];
    const int64_t unit = blockIdx.x;
    const int urows = PT << sshift;
    const int64_t row0 = unit * urows;
    const int rows = (int)min((int64_t)urows, n - row0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per_warp = urows / 4;
    for (int e = lane; e < d; e += 32) {
        float s = 0.f;
        const int r0 = warp * per_warp, r1 = min(rows, r0 + per_warp);
        int r = r0;
        for (; r + 8 <= r1; r += 8) {  // eight independent loads in flight, summed in row order
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldg(Xs + (row0 + r + u) * d + e);
#pragma unroll
            for (int u = 0; u < 8; ++u) s += v[u];
        }
        for (; r < r1; ++r) s += __ldg(Xs + (row0 + r) * d + e);
        part[warp * d + e] = s;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < d; e += PT) {
        const float m = (((part[e] + part[d + e]) + part[2 * d + e]) + part[3 * d + e]) / (float)rows;
        mean[e] = m;
        tmean[unit * d + e] = m;
    }
    __syncthreads();
    float r2 = 0.f;
    const bool vec4 = (d & 3) == 0 && (((uintptr_t)Xs) & 15) == 0;
    for (int r = threadIdx.x; r < rows; r += PT) {
        const float* x = Xs + (row0 + r) * d;
        float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
        if (vec4) {
            const float4* x4 = reinterpret_cast<const float4*>(x);
            const float4* m4 = reinterpret_cast<const float4*>(mean);
            int t = 0;
            for (; t + 4 <= (d >> 2); t += 4) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = __ldg(x4 + t + u);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float4 mm = m4[t + u];
                    float a = v[u].x - mm.x; q0 = fmaf(a, a, q0);
                    a = v[u].y - mm.y; q1 = fmaf(a, a, q1);
                    a = v[u].z - mm.z; q2 = fmaf(a, a, q2);
                    a = v[u].w - mm.w; q3 = fmaf(a, a, q3);
                }
            }
            for (; t < (d >> 2); ++t) {
                const float4 vv = __ldg(x4 + t), mm = m4[t];
                float a = vv.x - mm.x; q0 = fmaf(a, a, q0);
                a = vv.y - mm.y; q1 = fmaf(a, a, q1);
                a = vv.z - mm.z; q2 = fmaf(a, a, q2);
                a = vv.w - mm.w; q3 = fmaf(a, a, q3);
            }
        } else {
            for (int e = 0; e < d; ++e) { const float t = __ldg(x + e) - mean[e]; q0 = fmaf(t, t, q0); }
        }
        const float q = (q0 + q1) + (q2 + q3);
        r2 = fmaxf(r2, q);
        if (!(q == q)) r2 = q;  // NaN frame: keep every center (comparisons with NaN are false)
    }
    const bool bad = !(r2 == r2);
    const unsigned anybad = __ballot_sync(0xffffffffu, bad);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r2 = fmaxf(r2, __shfl_xor_sync(0xffffffffu, r2, o));
    if (lane == 0) red[warp] = anybad ? __int_as_float(0x7fc00000) : r2;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m2 = 0.f;
        bool nan = false;
        for (int w = 0; w < PT / 32; ++w) { if (!(red[w] == red[w])) nan = true; m2 = fmaxf(m2, red[w]); }
        trad[unit] = nan ? __int_as_float(0x7fc00000) : sqrtf(m2) * (1.f + 1e-5f);
    }
}

__device__ __forceinline__ void warp_stats(PruneStats* st, unsigned int padded, bool overflow) {
    atomicAdd(&st->total, (unsigned long long)padded);
    atomicMax(&st->max_count, padded);
    if (overflow) atomicAdd(&st->overflow, 1u);
}

// narrow rows: the center table (row stride ds floats, 16-byte aligned) sits in shared memory; one warp per tile.
// pass 1: dmin = min_j |c_j - p_t|; pass 2: keep j unless |c_j - p_t| > dmin + 2 R_t (+ slack), ascending j.
template <int DS>
__global__ void __launch_bounds__(256, 4) tile_lists_direct_kernel(const float* __restrict__ C, int k, int d,
                                                                const float* __restrict__ tmean,
                                                                const float* __restrict__ trad, int n_tiles, int lcap,
                                                                int pad_to, uint16_t dummy, uint16_t* __restrict__ tlist,
                                                                uint32_t* __restrict__ tcount, PruneStats* st, float margin) {
    extern __shared__ __align__(16) float ctab[];  // [k][DS]
    for (int t = threadIdx.x; t < k * DS; t += 256) {
        const int r = t / DS, c = t - r * DS;
        ctab[t] = c < d ? __ldg(C + (int64_t)r * d + c) : 0.f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp_global = (blockIdx.x * 256 + threadIdx.x) >> 5, n_warps = (gridDim.x * 256) >> 5;
    for (int tile = warp_global; tile < n_tiles; tile += n_warps) {
        float m[DS];
#pragma unroll
        for (int e = 0; e < DS; ++e) m[e] = e < d ? __ldg(tmean + (int64_t)tile * d + e) : 0.f;
        const float R = __ldg(trad + tile);
        auto dist2 = [&](int j) {
            const float4* row = reinterpret_cast<const float4*>(ctab + (size_t)j * DS);
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;  // four independent chains
#pragma unroll
            for (int e = 0; e < DS; e += 4) {
                const float4 c = row[e >> 2];
                float t = c.x - m[e]; s0 = fmaf(t, t, s0);
                t = c.y - m[e + 1]; s1 = fmaf(t, t, s1);
                t = c.z - m[e + 2]; s2 = fmaf(t, t, s2);
                t = c.w - m[e + 3]; s3 = fmaf(t, t, s3);
            }
            return (s0 + s1) + (s2 + s3);
        };
        // k <= 1024: every lane keeps its (up to 32) squared distances between the two passes
        constexpr int RMAX = 32;
        const bool cached = k <= 32 * RMAX;
        float dd[RMAX];
        float best = __int_as_float(0x7f800000);
        if (cached) {
#pragma unroll
            for (int r = 0; r < RMAX; ++r) {
                const int j = r * 32 + lane;
                dd[r] = (r * 32 < k) ? dist2(min(j, k - 1)) : __int_as_float(0x7f800000);
                if (j < k) best = fminf(best, dd[r]);
            }
        } else {
            for (int j = lane; j < k; j += 32) best = fminf(best, dist2(j));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, o));
        // keep unless D_j > thr  <=>  keep unless D_j^2 > thr^2 (both sides >= 0); NaN anywhere keeps the center
        // (+ 2 margin: the list stays valid while no center has moved farther than `margin` from where it is now)
        const float thr = (sqrtf(best) + 2.f * R + 2.f * margin) * (1.f + PRUNE_SLACK) + 1e-30f;
        const float thr2 = thr * thr;
        uint16_t* out = tlist + (size_t)tile * lcap;
        int count = 0;
        if (cached) {
#pragma unroll
            for (int r = 0; r < RMAX; ++r) {
                if (r * 32 < k) {
                    const int j = r * 32 + lane;
                    const bool keep = j < k && !(dd[r] > thr2);
                    const unsigned b = __ballot_sync(0xffffffffu, keep);
                    const int pos = count + __popc(b & ((1u << lane) - 1u));
                    if (keep && pos < lcap) out[pos] = (uint16_t)j;
                    count += __popc(b);
                }
            }
        } else {
            for (int j0 = 0; j0 < k; j0 += 32) {
                const int j = j0 + lane;
                const bool keep = j < k && !(dist2(min(j, k - 1)) > thr2);
                const unsigned b = __ballot_sync(0xffffffffu, keep);
                const int pos = count + __popc(b & ((1u << lane) - 1u));
                if (keep && pos < lcap) out[pos] = (uint16_t)j;
                count += __popc(b);
            }
        }
        const bool overflow = count > lcap;
        int padded = overflow ? lcap : (count + pad_to - 1) / pad_to * pad_to;
        if (padded > lcap) padded = lcap;
        for (int pos = count + lane; pos < padded; pos += 32) out[pos] = dummy;
        if (lane == 0) {
            tcount[tile] = overflow ? 0xffffffffu : (uint32_t)padded;
            warp_stats(st, (unsigned int)padded, overflow);
        }
    }
}

// cc[a][j] = |c_a - c_j| (fp32 differences, any order: the consumer's slack covers it).  32 x 32 output blocks of the upper
// triangle, the d-range walked in shared-memory slabs of 32 columns; 256 threads, 2 x 2 outputs each; the mirror block is
// written from the same sums (round 2's first version -- one warp per row and 32 columns, rows straight from L2 -- took
// 3.0 ms per step at k=5000, d=256)
__global__ void __launch_bounds__(256) center_dist_kernel(const float* __restrict__ C, int k, int d, float* __restrict__ cc) {
    __shared__ float As[32][33], Bs[32][33];
    // linear block index -> (bi <= bj)
    const int nb = (k + 31) / 32;
    int bi = 0, rem = blockIdx.x;
    while (rem >= nb - bi) { rem -= nb - bi; ++bi; }
    const int bj = bi + rem;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float s00 = 0.f, s01 = 0.f, s10 = 0.f, s11 = 0.f;
    for (int e0 = 0; e0 < d; e0 += 32) {
        for (int t = threadIdx.x; t < 32 * 32; t += 256) {
            const int r = t >> 5, c = t & 31;
            const int ra = bi * 32 + r, rb = bj * 32 + r;
            As[r][c] = (ra < k && e0 + c < d) ? __ldg(C + (int64_t)ra * d + e0 + c) : 0.f;
            Bs[r][c] = (rb < k && e0 + c < d) ? __ldg(C + (int64_t)rb * d + e0 + c) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int c = 0; c < 32; ++c) {
            const float a0 = As[2 * ty][c], a1 = As[2 * ty + 1][c];
            const float b0 = Bs[2 * tx][c], b1 = Bs[2 * tx + 1][c];
            float t = a0 - b0; s00 = fmaf(t, t, s00);
            t = a0 - b1; s01 = fmaf(t, t, s01);
            t = a1 - b0; s10 = fmaf(t, t, s10);
            t = a1 - b1; s11 = fmaf(t, t, s11);
        }
        __syncthreads();
    }
    const int a0 = bi * 32 + 2 * ty, j0 = bj * 32 + 2 * tx;
    const float v[2][2] = {{sqrtf(s00), sqrtf(s01)}, {sqrtf(s10), sqrtf(s11)}};
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            const int a = a0 + u, j = j0 + w;
            if (a < k && j < k) {
                cc[(int64_t)a * k + j] = v[u][w];
                cc[(int64_t)j * k + a] = v[u][w];
            }
        }
}

// wide rows: one warp per tile; a = label of the tile's first frame, delta = |c_a - p_t|;
// keep j unless cc[a][j] > 2 (delta + R_t) (+ slack)
__global__ void __launch_bounds__(256) tile_lists_cc_kernel(const float* __restrict__ C, int k, int d,
                                                            const float* __restrict__ cc,
                                                            const int32_t* __restrict__ labels_s, int64_t n, int sshift,
                                                            const float* __restrict__ tmean,
                                                            const float* __restrict__ trad, int n_tiles, int lcap,
                                                            int pad_to, uint16_t dummy, uint16_t* __restrict__ tlist,
                                                            uint32_t* __restrict__ tcount, PruneStats* st, float margin) {
    const int lane = threadIdx.x & 31;
    const int warp_global = (blockIdx.x * 256 + threadIdx.x) >> 5, n_warps = (gridDim.x * 256) >> 5;
    for (int tile = warp_global; tile < n_tiles; tile += n_warps) {
        int a = __ldg(labels_s + ((int64_t)tile * PT << sshift));
        if (a < 0 || a >= k) a = 0;
        const float* ca = C + (int64_t)a * d;
        const float* m = tmean + (int64_t)tile * d;
        float s = 0.f;
        for (int e = lane; e < d; e += 32) { const float t = __ldg(ca + e) - __ldg(m + e); s = fmaf(t, t, s); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float delta = sqrtf(s) * (1.f + 1e-5f);
        // (+ 4 margin: c_j and c_a may each move by `margin`, against each other and away from p_t, before the list is stale)
        const float thr = (2.f * (delta + __ldg(trad + tile)) + 4.f * margin) * (1.f + PRUNE_SLACK) + 1e-30f;
        const float* row = cc + (int64_t)a * k;
        uint16_t* out = tlist + (size_t)tile * lcap;
        int count = 0;
        for (int j0 = 0; j0 < k; j0 += 32) {
            const int j = j0 + lane;
            const bool keep = j < k && !(__ldg(row + min(j, k - 1)) > thr);
            const unsigned b = __ballot_sync(0xffffffffu, keep);
            const int pos = count + __popc(b & ((1u << lane) - 1u));
            if (keep && pos < lcap) out[pos] = (uint16_t)j;
            count += __popc(b);
        }
        const bool overflow = count > lcap;
        int padded = overflow ? lcap : (count + pad_to - 1) / pad_to * pad_to;
        if (padded > lcap) padded = lcap;
        for (int pos = count + lane; pos < padded; pos += 32) out[pos] = dummy;
        if (lane == 0) {
            tcount[tile] = overflow ? 0xffffffffu : (uint32_t)padded;
            warp_stats(st, (unsigned int)padded, overflow);
        }
    }
}

// max_j |c_j - c0_j| (rounded up), as float bits in *out (non-negative floats order like their bit patterns); NaN -> +inf
__global__ void __launch_bounds__(256) center_move_kernel(const float* __restrict__ C, const float* __restrict__ C0, int k, int d,
                                                          unsigned int* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int warp_global = (blockIdx.x * 256 + threadIdx.x) >> 5, n_warps = (gridDim.x * 256) >> 5;
    float worst = 0.f;
    for (int j = warp_global; j < k; j += n_warps) {
        float s = 0.f;
        for (int e = lane; e < d; e += 32) {
            const float t = __ldg(C + (int64_t)j * d + e) - __ldg(C0 + (int64_t)j * d + e);
            s = fmaf(t, t, s);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        float m = sqrtf(s) * (1.f + 1e-5f);
        if (!(m >= 0.f)) m = __int_as_float(0x7f800000);
        worst = fmaxf(worst, m);
    }
    if (lane == 0 && worst > 0.f) atomicMax(out, __float_as_uint(worst));
}

// ---- host side -------------------------------------------------------------------------------------------------
struct PruneState {
    b2k_ctx* ctx = nullptr;
    int64_t n = 0;
    int d = 0, k = 0, n_tiles = 0, lcap = 0, pad_to = 32;
    int sshift = 0, n_units = 0;  // 1 << sshift tiles share one center list
    DevMem Xs, perm, perm2, sigma, seg, labels_s, labels_t, tmean, trad, tlist, tcount, stats, cc, rank;
    std::vector<float> hC;
    std::vector<int> hidx, hrank;
    bool sorted = false;
    // list reuse: the centers the current lists were built for, the movement they tolerate, the stats of that build
    DevMem Clist, dmove;
    float* h_move = nullptr;   // pinned
    float margin = 0.f;
    double rmean = 0;          // mean tile radius of the current sort
    bool lists_valid = false;
    double c_mean = 0;
    int c_max = 0, c_ov = 0;
    ~PruneState() { if (h_move) cudaFreeHost(h_move); }
};

static unsigned grid_cap(b2k_ctx* ctx, int64_t items, int per_block, int per_sm = 8) {
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv(items, per_block), (int64_t)ctx->sm_count * per_sm));
}

bool prune_supported(const b2k_ctx* ctx, int64_t n, int d, int k) {
    if (ctx->prune_mode == 0) return false;
    // 16-bit center ids in the lists, the counting sort's shared-memory label table, at least a few tiles per SM
    if (k < 64 || k > 49152 || n >= (int64_t(1) << 32) - 1) return false;
    if (ctx->prune_mode >= 2) return n >= 2 * PT;  // tests: small jobs too (3: listed screen even when the lists are full)
    return n >= (int64_t)ctx->sm_count * PT * 8 && k >= 256;
}

int prune_create(b2k_ctx* ctx, int64_t n, int d, int k, PruneState** out) {
    PruneState* p = new PruneState();
    p->ctx = ctx; p->n = n; p->d = d; p->k = k;
    p->n_tiles = (int)cdiv(n, PT);
    // one list per tile: since the lists are kept over most iterations (prune_lists) their build cost no longer argues for
    // sharing one between two tiles of narrow rows (cfg2: mean list 189 -> 153, step 1.45 -> 1.40 ms)
    p->sshift = ctx->prune_unit_shift < 0 ? 0 : std::min(4, ctx->prune_unit_shift);
    p->n_units = (int)cdiv(n, (int64_t)PT << p->sshift);
    // room for every center (k <= 8192): a unit the bound cannot help simply lists them all, no special case downstream
    p->lcap = (int)std::min<int64_t>(8192, cdiv(k, 64) * 64);
    const int64_t n_pad = (int64_t)p->n_tiles * PT;
    int rc = p->Xs.alloc((size_t)n_pad * d * 4);
    if (rc == B2K_OK) rc = p->perm.alloc((size_t)n * 4);
    if (rc == B2K_OK) rc = p->perm2.alloc((size_t)n * 4);
    if (rc == B2K_OK) rc = p->sigma.alloc((size_t)n * 4);
    if (rc == B2K_OK) rc = p->seg.alloc((size_t)(k + 2) * 4);
    if (rc == B2K_OK) rc = p->labels_s.alloc((size_t)n_pad * 4);
    if (rc == B2K_OK) rc = p->labels_t.alloc((size_t)n_pad * 4);
    if (rc == B2K_OK) rc = p->tmean.alloc((size_t)p->n_units * d * 4);
    if (rc == B2K_OK) rc = p->trad.alloc((size_t)p->n_units * 4);
    if (rc == B2K_OK) rc = p->tlist.alloc((size_t)p->n_units * p->lcap * 2);
    if (rc == B2K_OK) rc = p->tcount.alloc((size_t)p->n_units * 4);
    if (rc == B2K_OK) rc = p->stats.alloc(sizeof(PruneStats));
    if (rc == B2K_OK) rc = p->rank.alloc((size_t)k * 4);
    if (rc != B2K_OK) { delete p; return rc; }
    *out = p;
    return B2K_OK;
}

void prune_destroy(PruneState* p) {
    if (!p) return;
    cudaStreamSynchronize(p->ctx->stream);
    delete p;
}

const float* prune_frames(const PruneState* p) { return p->Xs.as<float>(); }
int32_t* prune_labels(PruneState* p) { return p->labels_s.as<int32_t>(); }
const uint16_t* prune_tlist(const PruneState* p) { return p->tlist.as<uint16_t>(); }
const uint32_t* prune_tcount(const PruneState* p) { return p->tcount.as<uint32_t>(); }
int prune_lcap(const PruneState* p) { return p->lcap; }
int prune_unit_shift(const PruneState* p) { return p->sshift; }
bool prune_sorted(const PruneState* p) { return p->sorted; }

// (re)sort: `labels` are in the CURRENT order of the session's frames (original order before the first sort, sorted
// order afterwards); X is always the caller's original array
int prune_sort(PruneState* p, const float* X, const int32_t* labels, const float* dC) {
    b2k_ctx* ctx = p->ctx;
    cudaStream_t st = ctx->stream;
    const int64_t n = p->n;
    uint32_t* sigma = p->sigma.as<uint32_t>();
    // rank of every label in the kd order of the current centers (host: k*d floats down, k ints up)
    p->hC.resize((size_t)p->k * p->d);
    p->hidx.resize(p->k);
    p->hrank.resize(p->k);
    CUDA_TRY(cudaMemcpyAsync(p->hC.data(), dC, (size_t)p->k * p->d * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (int j = 0; j < p->k; ++j) p->hidx[j] = j;
    kd_order(p->hC.data(), p->d, p->hidx.data(), 0, p->k);
    for (int r = 0; r < p->k; ++r) p->hrank[p->hidx[r]] = r;
    CUDA_TRY(cudaMemcpyAsync(p->rank.p, p->hrank.data(), (size_t)p->k * 4, cudaMemcpyHostToDevice, st));
    int32_t* keys = p->labels_t.as<int32_t>();  // free until the labels are gathered below
    label_rank_kernel<<<grid_cap(ctx, n, 256), 256, 0, st>>>(labels, p->rank.as<int32_t>(), n, p->k, keys);
    LAUNCH_CHECK();
    // (a frame without a label in [0, k) would drop out of the sort: Lloyd assigns always label every frame)
    B2K_TRY(launch_label_sort(ctx, keys, n, p->k, p->seg.as<uint32_t>(), sigma));
    CUDA_TRY(cudaStreamSynchronize(st));  // hrank is reused by the next sort
    if (p->sorted) {
        compose_perm_kernel<<<grid_cap(ctx, n, 256), 256, 0, st>>>(p->perm.as<uint32_t>(), sigma, n, p->perm2.as<uint32_t>());
        LAUNCH_CHECK();
        std::swap(p->perm.p, p->perm2.p);
        std::swap(p->perm.cap, p->perm2.cap);
    } else {
        CUDA_TRY(cudaMemcpyAsync(p->perm.p, sigma, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    }
    // labels in the new order (the first frame of a tile names the tile's reference center)
    gather_values_kernel<int32_t><<<grid_cap(ctx, n, 256), 256, 0, st>>>(labels, sigma, n, p->labels_t.as<int32_t>());
    LAUNCH_CHECK();
    std::swap(p->labels_s.p, p->labels_t.p);
    std::swap(p->labels_s.cap, p->labels_t.cap);
    const uint32_t* perm = p->perm.as<uint32_t>();
    float* Xs = p->Xs.as<float>();
    if (p->d % 4 == 0 && ((uintptr_t)X & 15) == 0)
        gather_rows_kernel<float4><<<grid_cap(ctx, n * (p->d / 4), 256, 16), 256, 0, st>>>(
            reinterpret_cast<const float4*>(X), perm, n, p->d / 4, reinterpret_cast<float4*>(Xs));
    else if (p->d % 2 == 0 && ((uintptr_t)X & 7) == 0)
        gather_rows_kernel<float2><<<grid_cap(ctx, n * (p->d / 2), 256, 16), 256, 0, st>>>(
            reinterpret_cast<const float2*>(X), perm, n, p->d / 2, reinterpret_cast<float2*>(Xs));
    else
        gather_rows_kernel<float><<<grid_cap(ctx, n * p->d, 256, 16), 256, 0, st>>>(X, perm, n, p->d, Xs);
    LAUNCH_CHECK();
    tile_meta_kernel<<<(unsigned)p->n_units, PT, (size_t)5 * p->d * 4, st>>>(Xs, n, p->d, p->sshift, p->tmean.as<float>(),
                                                                          p->trad.as<float>());
    LAUNCH_CHECK();
    p->sorted = true;
    p->lists_valid = false;  // new tiles
    p->rmean = 0;
    if (ctx->prune_list_margin > 0) {  // mean tile radius: the scale of the movement margin of the lists
        p->hC.resize((size_t)p->n_units);
        CUDA_TRY(cudaMemcpyAsync(p->hC.data(), p->trad.p, (size_t)p->n_units * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        double acc = 0;
        for (int t = 0; t < p->n_units; ++t) acc += std::isfinite(p->hC[t]) ? (double)p->hC[t] : 0.0;
        p->rmean = acc / std::max(p->n_units, 1);
    }
    return B2K_OK;
}

// per-iteration center lists; stats come back to the host (one 16-byte read + sync): the caller decides whether the
// pruned screen is worth running this iteration
int prune_lists(PruneState* p, const float* dC, double* mean_count, int* max_count, int* overflow_tiles) {
    b2k_ctx* ctx = p->ctx;
    cudaStream_t st = ctx->stream;
    // List reuse.  A list built for centers c0 with `margin` in its bound excludes j only if j stays strictly farther
    // than the tile's reference center for EVERY set of centers with |c_j - c0_j| <= margin (triangle inequality, see the
    // kernels); late Lloyd iterations move the centers by 1e-3 of a tile radius and less, so instead of rebuilding the
    // lists every iteration (cfg2 0.15 ms, cfg4 1.5 ms) one small kernel measures the movement since the build.
    if (ctx->prune_list_margin > 0 && (!p->h_move || !p->Clist.p)) {
        if (p->Clist.alloc((size_t)p->k * p->d * 4) != B2K_OK || p->dmove.alloc(4) != B2K_OK ||
            (!p->h_move && cudaHostAlloc((void**)&p->h_move, 64, cudaHostAllocDefault) != cudaSuccess)) {
            cudaGetLastError();
            p->lists_valid = false;
            if (p->h_move) { cudaFreeHost(p->h_move); p->h_move = nullptr; }
        }
    }
    const bool can_reuse = ctx->prune_list_margin > 0 && p->h_move && p->Clist.p;
    if (can_reuse && p->lists_valid && p->margin > 0.f) {
        ProfScope prof_move(ctx, b2k_ctx::PROF_LISTS);
        CUDA_TRY(cudaMemsetAsync(p->dmove.p, 0, 4, st));
        center_move_kernel<<<grid_cap(ctx, p->k, 8), 256, 0, st>>>(dC, p->Clist.as<float>(), p->k, p->d, p->dmove.as<unsigned int>());
        LAUNCH_CHECK();
        CUDA_TRY(cudaMemcpyAsync(p->h_move, p->dmove.p, 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if ((double)*p->h_move * 1.001 <= (double)p->margin) {
            *mean_count = p->c_mean;
            *max_count = p->c_max;
            *overflow_tiles = p->c_ov;
            ctx->stat_list_reuse += 1;
            return B2K_OK;
        }
    }
    p->margin = can_reuse ? (float)(p->rmean * 1e-3 * ctx->prune_list_margin) : 0.f;
    if (!(p->margin >= 0.f) || !std::isfinite(p->margin)) p->margin = 0.f;
    const float margin = p->margin;
    if (can_reuse) CUDA_TRY(cudaMemcpyAsync(p->Clist.p, dC, (size_t)p->k * p->d * 4, cudaMemcpyDeviceToDevice, st));
    PruneStats* ds = p->stats.as<PruneStats>();
    CUDA_TRY(cudaMemsetAsync(ds, 0, sizeof(PruneStats), st));
    ProfScope* prof = new ProfScope(ctx, b2k_ctx::PROF_LISTS);
    const int k_pad = (int)(cdiv(p->k, 256) * 256);
    const uint16_t dummy = (uint16_t)k_pad;  // a -inf row behind the center operand (screen_centers_kernel)
    const int ds4 = (p->d + 3) & ~3;
    const size_t tab = (size_t)p->k * ds4 * 4;
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (200 * 1024) / std::max<size_t>(tab, 1)));
    const unsigned grid = grid_cap(ctx, p->n_units, 8, per_sm);
    if (p->d <= 16 && tab <= 96 * 1024) {
#define B2K_TL(DS)                                                                                                    \
    do {                                                                                                              \
        static PerDeviceOnce at;                                                                                      \
        if (at.need(ctx->device)) {                                                                                   \
            CUDA_TRY(cudaFuncSetAttribute(tile_lists_direct_kernel<DS>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                          96 * 1024));                                                                \
            at.done(ctx->device);                                                                                     \
        }                                                                                                             \
        tile_lists_direct_kernel<DS><<<grid, 256, tab, st>>>(dC, p->k, p->d, p->tmean.as<float>(), p->trad.as<float>(), \
                                                             p->n_units, p->lcap, p->pad_to, dummy,                   \
                                                             p->tlist.as<uint16_t>(), p->tcount.as<uint32_t>(), ds,   \
                                                             margin);                                                 \
    } while (0)
        if (ds4 == 4) B2K_TL(4);
        else if (ds4 == 8) B2K_TL(8);
        else if (ds4 == 12) B2K_TL(12);
        else B2K_TL(16);
#undef B2K_TL
        LAUNCH_CHECK();
    } else {
        B2K_TRY(p->cc.alloc((size_t)p->k * p->k * 4));
        const int64_t nb = cdiv(p->k, 32);
        center_dist_kernel<<<(unsigned)(nb * (nb + 1) / 2), 256, 0, st>>>(dC, p->k, p->d, p->cc.as<float>());
        LAUNCH_CHECK();
        tile_lists_cc_kernel<<<grid_cap(ctx, p->n_units, 8, 8), 256, 0, st>>>(
            dC, p->k, p->d, p->cc.as<float>(), p->labels_s.as<int32_t>(), p->n, p->sshift, p->tmean.as<float>(),
            p->trad.as<float>(), p->n_units, p->lcap, p->pad_to, dummy, p->tlist.as<uint16_t>(), p->tcount.as<uint32_t>(), ds,
            margin);
        LAUNCH_CHECK();
    }
    delete prof;
    PruneStats h;
    CUDA_TRY(cudaMemcpyAsync(&h, ds, sizeof(h), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *mean_count = p->c_mean = (double)h.total / (double)std::max(p->n_units, 1);
    *max_count = p->c_max = (int)h.max_count;
    *overflow_tiles = p->c_ov = (int)h.overflow;
    p->lists_valid = can_reuse;
    return B2K_OK;
}

int prune_scatter_labels(PruneState* p, int32_t* out) {
    scatter_labels_kernel<<<grid_cap(p->ctx, p->n, 256, 16), 256, 0, p->ctx->stream>>>(
        p->labels_s.as<int32_t>(), p->perm.as<uint32_t>(), p->n, out);
    LAUNCH_CHECK();
    return B2K_OK;
}

}  // namespace b2k
