// kmpp.cu -- k-means++ seeding (K4 of SURVEY 2.2).
//
// Replaces deeptime kmeans.init_centers_kmpp(data, k, random_seed, n_threads, callback)
// (signature evidenced by pyemma/coordinates/clustering/tests/test_kmeans.py:304; invoked
// from KMeans.fit, call site pyemma/coordinates/clustering/kmeans.py:254-255).
//
// Algorithm: first center uniform; D2[i] = compute(x_i,c0)^2; every round
// draws n_trials = 2 + floor(ln k) thresholds r_j = dist_sum * u_j, picks candidate_j as the
// first frame whose running D2 prefix reaches r_j, evaluates each candidate's potential
// sum_i min(D2[i], compute(x_i, cand_j)^2), keeps the best, updates D2.
//
// The RNG stream does not depend on the data, so the host precomputes every u_j
// (own mt19937 + the libstdc++ uniform_int mapping, restated below) and the device never waits
// for a random number.  Distances are the exact fp32 reference-order kernels (exact.cu / rmsd.cu).
//
// Ordered sums.  The reference's prefix scan, potentials and dist_sum bookkeeping are fp32 sums
// in frame order, so their bits depend on that order:
//   B2K_KMPP_SERIAL  reproduces that order exactly (one warp walks the array; bit-faithful to the
//                    reference at n_jobs=1; inherently latency-bound: 1 dependent FADD per frame)
//   B2K_KMPP_BLOCKED every sum is the balanced binary tree over aligned power-of-two index blocks,
//                    dist_sum is the tree root, and candidates are found by descending the tree
//                    (left if r <= sum(left) else r -= sum(left)).  Deterministic, parallel, and
//                    defined identically in oracle/oracle.cpp so parity stays bit-exact.
#include "common.cuh"
#include "kernels.h"
#include <random>

namespace b2k {

// ---- host RNG: mt19937 + libstdc++ uniform_int_distribution (Lemire) + (T)g()/(T)g.max() -----
struct MT19937 {
    uint32_t mt[624];
    int idx;
    explicit MT19937(uint32_t seed) {
        mt[0] = seed;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        idx = 624;
    }
    uint32_t next() {
        if (idx >= 624) {
            for (int i = 0; i < 624; ++i) {
                const uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
                mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    // uniform integer in [0, range) for range <= 2^32 (libstdc++ _S_nd<uint64_t> on a 32-bit URBG)
    uint64_t below(uint64_t range) {
        if (range == (uint64_t(1) << 32)) return next();
        const uint32_t r32 = (uint32_t)range;
        uint64_t product = (uint64_t)next() * r32;
        uint32_t low = (uint32_t)product;
        if (low < r32) {
            const uint32_t threshold = (0u - r32) % r32;
            while (low < threshold) {
                product = (uint64_t)next() * r32;
                low = (uint32_t)product;
            }
        }
        return product >> 32;
    }
    float unit() { return (float)next() / (float)0xffffffffu; }
};

// ---- device state --------------------------------------------------------------------------
#define KMPP_MAX_TRIALS 24
struct KmppState {
    float dist_sum;
    int n_cand;
    long long cand[KMPP_MAX_TRIALS];  // -1 = none
    float rands[KMPP_MAX_TRIALS];
    float pot[KMPP_MAX_TRIALS];
    long long best;  // frame index of the accepted center
    int jbest;       // which candidate (-1: fallback "first non-taken frame")
    float delta_sum;
    int round;       // asynchronous rounds: centers found so far (the device-side loop counter a CUDA graph replays on)
};

__global__ void kmpp_square_init_kernel(const float* __restrict__ v, int64_t n, int64_t first, float* __restrict__ D,
                                        unsigned char* __restrict__ taken) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = v[i];
    D[i] = (i == first) ? 0.f : __fmul_rn(x, x);
    taken[i] = (i == first) ? 1 : 0;
}

// cd[j][i] <- contribution of frame i to candidate j's potential (0 for taken / self / no candidate)
__global__ void kmpp_contrib_kernel(float* __restrict__ cd, int64_t n, int m, const float* __restrict__ D,
                                    const unsigned char* __restrict__ taken, const KmppState* __restrict__ st) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool tk = taken[i] != 0;
    const float di = D[i];
    for (int j = 0; j < m; ++j) {
        const long long c = st->cand[j];
        float out = 0.f;
        if (!tk && c >= 0 && c != i) {
            const float v = cd[(int64_t)j * n + i];
            const float dd = __fmul_rn(v, v);
            out = (dd < di) ? dd : di;
        }
        cd[(int64_t)j * n + i] = out;
    }
}

// gather candidate rows (missing candidates -> row of frame 0, ignored later)
__global__ void kmpp_gather_kernel(const float* __restrict__ X, int d, const KmppState* __restrict__ st, int m,
                                   float* __restrict__ rows) {
    const int j = blockIdx.x;
    long long c = st->cand[j];
    if (c < 0) c = 0;
    for (int e = threadIdx.x; e < d; e += blockDim.x) rows[(int64_t)j * d + e] = X[c * d + e];
}

// D[i] <- min(D[i], d2(x_i,best)) using the stored contributions; delta[i] <- change (serial mode)
__global__ void kmpp_update_kernel(float* __restrict__ D, const unsigned char* __restrict__ taken, int64_t n,
                                   const float* __restrict__ src /* cd[jbest] or fresh distances (sqrt'd) */,
                                   int src_is_sqrt, float* __restrict__ delta, int32_t* __restrict__ assigned = nullptr,
                                   int32_t center_index = 0, const uint16_t* __restrict__ framemask = nullptr,
                                   int jbit = 0) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float dl = 0.f;
    // framemask: src holds a distance only where the pair was not pruned (a pruned pair cannot lower D2)
    if (!taken[i] && (!framemask || ((framemask[i] >> jbit) & 1u))) {
        float dd = src[i];
        if (src_is_sqrt) dd = __fmul_rn(dd, dd);
        const float old = D[i];
        if (dd < old) {
            dl = __fsub_rn(dd, old);
            D[i] = dd;
            if (assigned) assigned[i] = center_index;  // the chosen center that realises D2 (pruning, exact.cu)
        }
    }
    if (delta) delta[i] = dl;
}

// The same update with the accepted candidate read from the device state: the host does not wait for the round's result
// (kmpp_run_blocked, asynchronous rounds).  *fail != 0: an earlier round found no candidate -> the run is repeated with
// the synchronous loop, whatever happens here is discarded.  It shares its launch with the tree_up_kernel pass over the
// updated D2 (heights 5 and 10) the NEXT round starts from: a thread updates its frame and feeds the value straight into
// the butterflies (same statements, same order).
__global__ void __launch_bounds__(1024) kmpp_update_tree_kernel(float* __restrict__ D, const unsigned char* __restrict__ taken,
                                                                int64_t n, const float* __restrict__ cd,
                                                                const KmppState* __restrict__ st, int src_is_sqrt,
                                                                int32_t* __restrict__ assigned,
                                                                const uint16_t* __restrict__ framemask,
                                                                const int* __restrict__ fail, float* __restrict__ out5,
                                                                float* __restrict__ out10) {
    __shared__ float ws[32];
    const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    float v = 0.f;
    if (i < n) {
        const bool tk = taken[i] != 0;
        float cur = D[i];
        const int jbit = st->jbest;
        if (!*fail && jbit >= 0 && !tk && (!framemask || ((framemask[i] >> jbit) & 1u))) {
            float dd = cd[(size_t)jbit * n + i];
            if (src_is_sqrt) dd = __fmul_rn(dd, dd);
            if (dd < cur) {
                D[i] = dd;
                cur = dd;
                if (assigned) assigned[i] = st->round - 1;
            }
        }
        v = tk ? 0.f : cur;
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        ws[w] = v;
        const int64_t t5 = (int64_t)blockIdx.x * 32 + w;
        if (t5 * 32 < n) out5[t5] = v;
    }
    __syncthreads();
    if (w == 0) {
        float s = ws[threadIdx.x];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
        if (threadIdx.x == 0) out10[blockIdx.x] = s;
    }
}

__global__ void kmpp_commit_dev_kernel(KmppState* st, long long lo, int64_t n_local, const float* __restrict__ rows, int d,
                                       unsigned char* __restrict__ taken, float* __restrict__ centers,
                                       long long* __restrict__ chosen, int* fail, int k) {
    const long long best = st->best;
    const int jbest = st->jbest;
    const int found = st->round;  // this round's center index
    __syncthreads();              // every thread has read the counter before thread 0 advances it
    if (found >= k) return;       // (a graph replayed once too often: nothing left to pick)
    if (best < 0 || jbest < 0) {
        if (threadIdx.x == 0) { *fail = 1; st->round = found + 1; }
        return;
    }
    const float* row = rows + (size_t)jbest * d;
    float* center_out = centers + (size_t)found * d;
    for (int e = threadIdx.x; e < d; e += blockDim.x) center_out[e] = row[e];
    if (threadIdx.x == 0) {
        chosen[found] = best;
        const long long b = best - lo;
        if (b >= 0 && b < n_local) taken[b] = 1;
        st->round = found + 1;
    }
}

__global__ void kmpp_set_round_kernel(KmppState* st, int round) { st->round = round; }

// Rc[j][a] = lower bound of |center_a - candidate_j| (fp64 sum, rounded down): one warp per (a, j)
__global__ void __launch_bounds__(256) kmpp_center_cand_dist_kernel(const float* __restrict__ centers, int found,
                                                                    const float* __restrict__ rows, int m, int d,
                                                                    float* __restrict__ Rc, int rc_stride,
                                                                    const KmppState* __restrict__ st = nullptr) {
    if (st) found = st->round;  // asynchronous rounds: a fixed grid strides over the found x m pairs
    const int lane = threadIdx.x & 31;
    const int n_w = (gridDim.x * 256) >> 5;
    for (int w = (blockIdx.x * 256 + threadIdx.x) >> 5; w < found * m; w += n_w) {
    const int a = w / m, j = w - a * m;
    const float* c = centers + (size_t)a * d;
    const float* r = rows + (size_t)j * d;
    double s = 0.0;
    for (int e = lane; e < d; e += 32) {
        const double t = (double)c[e] - (double)r[e];
        s += t * t;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        const double R = sqrt(s) * (1.0 - 1e-6);
        float f = (float)R;
        if ((double)f > R) f = nextafterf(f, 0.f);  // never above the true distance
        Rc[(size_t)a * rc_stride + j] = (s == s) ? f : 0.f;  // [center][candidate]; NaN data: no pruning
    }
    }
}

// ---- SERIAL ordered sums: one warp, values fetched coalesced, summed in frame order ------------
// prefix scan + candidate pick (reference: `sum += D2[i]; if (sum >= r_j && cand_j == none) cand_j = i`)
__global__ void __launch_bounds__(32) kmpp_serial_scan_kernel(const float* __restrict__ D,
                                                              const unsigned char* __restrict__ taken, int64_t n,
                                                              KmppState* st, const float* __restrict__ u, int m) {
    const int lane = threadIdx.x;
    __shared__ float r[KMPP_MAX_TRIALS];
    __shared__ long long cand[KMPP_MAX_TRIALS];
    if (lane < m) {
        r[lane] = __fmul_rn(st->dist_sum, u[lane]);
        cand[lane] = -1;
    }
    __syncwarp();
    float pending = 3.402823466e+38f;  // smallest unmet threshold
    int unmet = m;
    unsigned found_mask = 0;  // warp-uniform: which thresholds already have their candidate
    for (int j = 0; j < m; ++j) pending = fminf(pending, r[j]);
    float sum = 0.f;
    for (int64_t i0 = 0; i0 < n && unmet > 0; i0 += 32) {
        const int64_t i = i0 + lane;
        float v = 0.f;
        bool act = false;
        if (i < n && !taken[i]) { v = D[i]; act = true; }
        const unsigned actmask = __ballot_sync(0xffffffffu, act);
        for (int t = 0; t < 32; ++t) {
            const float vt = __shfl_sync(0xffffffffu, v, t);
            if (!((actmask >> t) & 1u)) continue;
            sum = __fadd_rn(sum, vt);
            if (sum >= pending) {
                pending = 3.402823466e+38f;
                unmet = 0;
                for (int j = 0; j < m; ++j) {
                    if (!((found_mask >> j) & 1u)) {
                        if (sum >= r[j]) { found_mask |= 1u << j; if (lane == 0) cand[j] = i0 + t; }
                        else { pending = fminf(pending, r[j]); ++unmet; }
                    }
                }
            }
        }
    }
    __syncwarp();
    if (lane < m) {
        st->cand[lane] = cand[lane];
        st->rands[lane] = r[lane];
    }
    if (lane == 0) st->n_cand = m;
}

// out[j] = ((..(0 + a[j][0]) + a[j][1]) + ...) fp32 in index order.  The block stages tiles of
// all m arrays coalesced into smem; thread j then walks row j sequentially (one dependent FADD
// per element -- the reference's order allows nothing faster).
#define KSUM_TI 256
__global__ void __launch_bounds__(256) kmpp_serial_sum_kernel(const float* __restrict__ a, int64_t n, int m,
                                                              float* __restrict__ out) {
    __shared__ float tile[KMPP_MAX_TRIALS][KSUM_TI + 1];
    const int tid = threadIdx.x;
    float s = 0.f;
    for (int64_t i0 = 0; i0 < n; i0 += KSUM_TI) {
        const int len = (int)min((int64_t)KSUM_TI, n - i0);
        for (int j = 0; j < m; ++j)
            if (tid < len) tile[j][tid] = a[(int64_t)j * n + i0 + tid];
        __syncthreads();
        if (tid < m) {
#pragma unroll 8
            for (int t = 0; t < len; ++t) s = __fadd_rn(s, tile[tid][t]);
        }
        __syncthreads();
    }
    if (tid < m) out[tid] = s;
}

// initial dist_sum: ordered sum of D over non-taken frames (D[first]==0 so no mask needed)
__global__ void kmpp_set_dist_sum_kernel(KmppState* st, const float* __restrict__ s) { st->dist_sum = s[0]; }

// ---- BLOCKED ordered sums: balanced tree, levels stored every 5 heights ----------------------
// in: len values (level h0; masked by `taken` when given).  out5[t] = tree sum of 32, out10[b] of 1024.
__global__ void __launch_bounds__(1024) tree_up_kernel(const float* __restrict__ in, int64_t len, int64_t in_stride,
                                                       const unsigned char* __restrict__ taken,
                                                       float* __restrict__ out5, int64_t out5_stride,
                                                       float* __restrict__ out10, int64_t out10_stride) {
    __shared__ float ws[32];
    const int j = blockIdx.y;
    const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    float v = 0.f;
    if (i < len) {
        v = in[(int64_t)j * in_stride + i];
        if (taken && taken[i]) v = 0.f;
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        ws[w] = v;
        const int64_t t5 = (int64_t)blockIdx.x * 32 + w;
        if (out5 && t5 * 32 < len) out5[(int64_t)j * out5_stride + t5] = v;
    }
    __syncthreads();
    if (w == 0) {
        float s = ws[threadIdx.x];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
        if (threadIdx.x == 0) out10[(int64_t)j * out10_stride + blockIdx.x] = s;
    }
}

struct TreeLevels {
    const float* lv[7];  // heights 0,5,...,30
    long long len[7];
    int H;  // height of the root: 2^H = pow2_ceil(n)
};

__device__ float tree_node_sum(const TreeLevels& T, const unsigned char* taken, int h, long long t) {
    const int b = h / 5, r = h - 5 * b;
    const int cnt = 1 << r;
    float vals[16];
    for (int q = 0; q < cnt; ++q) {
        const long long idx = t * cnt + q;
        float v = 0.f;
        if (idx < T.len[b]) {
            v = T.lv[b][idx];
            if (b == 0 && taken && taken[idx]) v = 0.f;
        }
        vals[q] = v;
    }
    for (int s = 1; s < cnt; s <<= 1)
        for (int q = 0; q < cnt; q += 2 * s) vals[q] = __fadd_rn(vals[q], vals[q + s]);
    return vals[0];
}

__global__ void kmpp_tree_pick_kernel(TreeLevels T, const unsigned char* __restrict__ taken, int64_t n,
                                      KmppState* st, const float* __restrict__ u, int m) {
    __shared__ float root;
    if (threadIdx.x == 0) {
        root = tree_node_sum(T, taken, T.H, 0);
        st->dist_sum = root;
        st->n_cand = m;
    }
    __syncthreads();
    const int j = threadIdx.x;
    if (j >= m) return;
    float r = __fmul_rn(root, u[j]);
    st->rands[j] = r;
    long long node = 0;
    for (int h = T.H; h > 0; --h) {
        const float left = tree_node_sum(T, taken, h - 1, 2 * node);
        if (r <= left) node = 2 * node;
        else { r = __fsub_rn(r, left); node = 2 * node + 1; }
    }
    st->cand[j] = (node < n && !taken[node]) ? node : -1;
}

// Sharded form of the pick: the levels at height >= 10 are global (replicated on every rank, because every
// height-10 node lies inside one shard), the levels below are local.  Phase A (every rank, identical result)
// descends from max(H,10) to the height-10 node -- starting above H only meets zero right siblings, and
// r = root*u <= root always goes left there, so the walk equals the one from H; phase B (the owner of that node)
// finishes the descent in its shard.
__global__ void kmpp_pick_top_kernel(TreeLevels T, int Hs, KmppState* st, const float* __restrict__ u, int m,
                                     long long* __restrict__ node10, float* __restrict__ resid, int by_round = 0) {
    __shared__ float root;
    if (by_round) u += (size_t)(st->round - 1) * m;  // this round's uniforms (the host does not pass a per-round pointer)
    if (threadIdx.x == 0) {
        root = tree_node_sum(T, nullptr, Hs, 0);
        st->dist_sum = root;
        st->n_cand = m;
    }
    __syncthreads();
    const int j = threadIdx.x;
    if (j >= m) return;
    float r = __fmul_rn(root, u[j]);
    st->rands[j] = r;
    long long node = 0;
    for (int h = Hs; h > 10; --h) {
        const float left = tree_node_sum(T, nullptr, h - 1, 2 * node);
        if (r <= left) node = 2 * node;
        else { r = __fsub_rn(r, left); node = 2 * node + 1; }
    }
    node10[j] = node;
    resid[j] = r;
}

// local levels: T.lv[0] = D (n_local), T.lv[1] = L5 (local); node10_lo = index of this shard's first height-10 node
__global__ void kmpp_pick_leaf_kernel(TreeLevels T, const unsigned char* __restrict__ taken, int64_t n_local,
                                      long long node10_lo, long long lo, const long long* __restrict__ node10,
                                      const float* __restrict__ resid, int m, long long* __restrict__ cand_out) {
    const int j = threadIdx.x;
    if (j >= m) return;
    long long node = node10[j] - node10_lo;
    long long c = -1;
    if (node >= 0 && node * 1024 < n_local) {
        float r = resid[j];
        for (int h = 10; h > 0; --h) {
            const float left = tree_node_sum(T, taken, h - 1, 2 * node);
            if (r <= left) node = 2 * node;
            else { r = __fsub_rn(r, left); node = 2 * node + 1; }
        }
        if (node < n_local && !taken[node]) c = lo + node;
    }
    cand_out[j] = c;
}

__global__ void kmpp_set_cands_kernel(KmppState* st, const long long* __restrict__ cand, int m) {
    if (threadIdx.x < m) st->cand[threadIdx.x] = cand[threadIdx.x];
}

// rows[j] = X[cand_j - lo] when this shard owns the candidate, else zeros (the ranks' rows are summed)
__global__ void kmpp_gather_sharded_kernel(const float* __restrict__ X, int d, const long long* __restrict__ cand,
                                           long long lo, int64_t n_local, float* __restrict__ rows) {
    const int j = blockIdx.x;
    const long long c = cand[j] - lo;
    const bool mine = cand[j] >= 0 && c >= 0 && c < n_local;
    for (int e = threadIdx.x; e < d; e += blockDim.x) rows[(int64_t)j * d + e] = mine ? X[c * d + e] : 0.f;
}

// Asynchronous single-GPU rounds: candidate pick (upper tree levels, leaf descent), candidate registration and the gather
// of the m candidate rows in ONE launch -- the same statements as kmpp_pick_top / pick_leaf / set_cands / gather_sharded,
// four graph nodes (and their launch latencies on a 100 kB problem) fewer per round.
// One WARP per candidate.  A thread walking the tree alone pays one dependent round trip to L2 per level (and
// tree_node_sum's 2^r loads per call): 17.8 us per launch at H = 17, the longest node of a cfg1 round.  A stored level
// holds the 32 descendants of a node five levels up, so the warp loads them with one coalesced request, builds every
// pairwise partial sum of that block with five shuffle stages -- S[s+1][q] = S[s][q] + S[s][q + 2^s], the very additions
// tree_node_sum performs, in its order -- and walks five levels out of registers: one round trip per five levels.
// tree_up_kernel's two upper levels (heights 15 and 20) of ONE tree with at most 1024 height-10 sums, by the calling CTA:
// the same butterflies over the same aligned blocks of 32, so the same bits.  l15: room for ceil(n10/32) <= 32 sums
// (global or shared memory).
__device__ __forceinline__ void cta_tree_top(const float* l10, long long n10, float* l15, float* l20) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int n15 = (int)((n10 + 31) / 32);
    for (int t5 = w; t5 < n15; t5 += nw) {  // (n15 <= 32)
        const long long i = (long long)t5 * 32 + lane;
        float v = i < n10 ? l10[i] : 0.f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (lane == 0) l15[t5] = v;
    }
    __syncthreads();
    if (w == 0) {
        float sum = lane < n15 ? l15[lane] : 0.f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) sum = __fadd_rn(sum, __shfl_xor_sync(0xffffffffu, sum, o));
        if (lane == 0) l20[0] = sum;
    }
    __syncthreads();
}

__device__ __forceinline__ float warp_level_entry(const TreeLevels& T, const unsigned char* taken, int b, long long idx) {
    float v = 0.f;
    if (idx < T.len[b]) {
        v = T.lv[b][idx];
        if (b == 0 && taken && taken[idx]) v = 0.f;
    }
    return v;
}

// tree_node_sum(T, taken, h, t), computed by the whole warp (same additions); every lane returns the sum
__device__ __forceinline__ float warp_node_sum(const TreeLevels& T, const unsigned char* taken, int h, long long t) {
    const int lane = threadIdx.x & 31;
    const int b = h / 5, r = h - 5 * b;
    float v = lane < (1 << r) ? warp_level_entry(T, taken, b, (t << r) + lane) : 0.f;
    for (int s = 0; s < r; ++s) v = __fadd_rn(v, __shfl_down_sync(0xffffffffu, v, 1 << s));
    return __shfl_sync(0xffffffffu, v, 0);
}

// walk from `node` at height h_top down to height h_stop (a multiple of 5 below h_top); r is the residual
__device__ __forceinline__ long long warp_descend(const TreeLevels& T, const unsigned char* taken, int h_top, int h_stop,
                                                  long long node, float& r) {
    const int lane = threadIdx.x & 31;
    int h = h_top;
    while (h > h_stop) {
        const int b = (h - 1) / 5, rr = h - 5 * b;  // rr levels inside this block, children stored at level b
        const long long base = node << rr;
        float S[6];
        S[0] = lane < (1 << rr) ? warp_level_entry(T, taken, b, base + lane) : 0.f;
#pragma unroll
        for (int sgl = 0; sgl < 5; ++sgl) S[sgl + 1] = __fadd_rn(S[sgl], __shfl_down_sync(0xffffffffu, S[sgl], 1 << sgl));
        int off = 0;
#pragma unroll
        for (int lev = 4; lev >= 0; --lev) {
            const float left = __shfl_sync(0xffffffffu, S[lev], off);  // the 2^lev entries from `off` on: the left child
            if (lev < rr) {
                if (!(r <= left)) { r = __fsub_rn(r, left); off += 1 << lev; }
            }
        }
        node = base + off;
        h = 5 * b;
    }
    return node;
}

__global__ void __launch_bounds__(32 * KMPP_MAX_TRIALS) kmpp_pick_fused_kernel(TreeLevels T, int Hs, KmppState* st, const float* __restrict__ U,
                                                              int m, const unsigned char* __restrict__ taken, int64_t n_local,
                                                              const float* __restrict__ X, int d, long long* __restrict__ cand_out,
                                                              float* __restrict__ rows, float* top15, float* top20) {
    __shared__ long long cand_s[KMPP_MAX_TRIALS];
    // small trees (at most 1024 height-10 sums): the upper levels of the D^2 tree are built here instead of by a launch of
    // their own (top15 = T.lv[3] has room for 32 sums, top20 = T.lv[4])
    if (top15) cta_tree_top(T.lv[2], T.len[2], top15, top20);
    const float* u = U + (size_t)(st->round - 1) * m;
    const int j = threadIdx.x >> 5, lane = threadIdx.x & 31;  // blockDim.x = 32 * m
    const float root = warp_node_sum(T, nullptr, Hs, 0);
    float r = __fmul_rn(root, u[j]);
    const float r0 = r;
    long long node = 0;
    if (Hs > 10) node = warp_descend(T, nullptr, Hs, 10, 0, r);
    long long c = -1;
    if (node >= 0 && node * 1024 < n_local) {
        node = warp_descend(T, taken, 10, 0, node, r);
        if (node < n_local && !taken[node]) c = node;
    }
    if (lane == 0) {
        if (j == 0) {
            st->dist_sum = root;
            st->n_cand = m;
        }
        st->rands[j] = r0;
        cand_out[j] = c;
        st->cand[j] = c;
        cand_s[j] = c;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < m * d; t += blockDim.x) {
        const int jj = t / d, e = t - jj * d;
        const long long cc = cand_s[jj];
        rows[t] = (cc >= 0 && cc < n_local) ? X[cc * d + e] : 0.f;
    }
}

// ... and potentials -> selection -> commit in one launch (kmpp_tree_roots / select / commit_dev)
__global__ void __launch_bounds__(128) kmpp_select_commit_kernel(KmppState* st, const float* __restrict__ lv, int64_t stride, int m,
                                                                 float* __restrict__ pots, int64_t n_local,
                                                                 const float* __restrict__ rows, int d,
                                                                 unsigned char* __restrict__ taken, float* __restrict__ centers,
                                                                 long long* __restrict__ chosen, int* fail, int k,
                                                                 const float* pot10, long long n10, float* pot20) {
    __shared__ long long s_best;
    __shared__ int s_jbest, s_found;
    __shared__ float s15[32];
    // small trees: the roots of the m potential trees (height 20) from their height-10 sums, here instead of in a launch of
    // their own; lv == pot20, stride 1
    if (pot10)
        for (int j = 0; j < m; ++j) cta_tree_top(pot10 + (long long)j * n10, n10, s15, pot20 + j);
    if (threadIdx.x == 0) {
        long long best = -1;
        int jbest = -1;
        float bp = 3.402823466e+38f;
        for (int j = 0; j < m; ++j) {
            const float p = lv[(int64_t)j * stride];
            pots[j] = p;
            st->pot[j] = p;
            if (st->cand[j] >= 0 && p < bp) { bp = p; best = st->cand[j]; jbest = j; }
        }
        st->best = best;
        st->jbest = jbest;
        s_best = best;
        s_jbest = jbest;
        s_found = st->round;
    }
    __syncthreads();
    const long long best = s_best;
    const int jbest = s_jbest, found = s_found;
    if (found >= k) return;
    if (best < 0 || jbest < 0) {
        if (threadIdx.x == 0) { *fail = 1; st->round = found + 1; }
        return;
    }
    const float* row = rows + (size_t)jbest * d;
    float* center_out = centers + (size_t)found * d;
    for (int e = threadIdx.x; e < d; e += blockDim.x) center_out[e] = row[e];
    if (threadIdx.x == 0) {
        chosen[found] = best;
        if (best < n_local) taken[best] = 1;
        st->round = found + 1;
    }
}

// contribution of local frame i to candidate j's potential; candidates are GLOBAL indices
__global__ void kmpp_contrib_sharded_kernel(float* __restrict__ cd, int64_t n, int m, const float* __restrict__ D,
                                            const unsigned char* __restrict__ taken,
                                            const long long* __restrict__ cand, long long lo) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool tk = taken[i] != 0;
    const float di = D[i];
    for (int j = 0; j < m; ++j) {
        const long long c = cand[j];
        float out = 0.f;
        if (!tk && c >= 0 && c != lo + i) {
            const float v = cd[(int64_t)j * n + i];
            const float dd = __fmul_rn(v, v);
            out = (dd < di) ? dd : di;
        }
        cd[(int64_t)j * n + i] = out;
    }
}

__global__ void kmpp_commit_sharded_kernel(long long best, long long lo, int64_t n_local, const float* __restrict__ row,
                                           int d, unsigned char* __restrict__ taken, float* __restrict__ center_out) {
    for (int e = threadIdx.x; e < d; e += blockDim.x) center_out[e] = row[e];
    const long long b = best - lo;
    if (threadIdx.x == 0 && b >= 0 && b < n_local) taken[b] = 1;
}

__global__ void kmpp_first_free_sharded_kernel(const unsigned char* __restrict__ taken, int64_t n, long long lo,
                                               long long* out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !taken[i]) atomicMin((unsigned long long*)out, (unsigned long long)(lo + i));
}

// Height-10 sums of the m contribution trees straight from the pruned distance rows: frame i contributes to tree j
//   0 (taken / the candidate itself / no candidate), D2_i (pair pruned), min(D2_i, dist^2) otherwise
// -- the values kmpp_contrib_sharded_kernel writes -- and the sums run in tree_up_kernel's order (xor butterfly over
// the warp = balanced tree of 32, then over the 32 warp sums), so the potentials are bit-identical; the m x n
// contribution arrays are neither written nor re-read.
__global__ void __launch_bounds__(1024) kmpp_pot_tree_kernel(const float* __restrict__ cd, int64_t n, int m,
                                                             const float* __restrict__ D,
                                                             const unsigned char* __restrict__ taken,
                                                             const uint16_t* __restrict__ framemask,
                                                             const long long* __restrict__ cand, long long lo,
                                                             float* __restrict__ out10, int64_t out10_stride) {
    __shared__ float ws[4][32];
    const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool in = i < n;
    const bool tk = in ? taken[i] != 0 : true;
    const float di = in ? D[i] : 0.f;
    const unsigned mask = in ? (framemask ? (unsigned)framemask[i] : 0xffffu) : 0u;  // no mask: every pair is evaluated
    for (int j0 = 0; j0 < m; j0 += 4) {
        float v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int j = j0 + q;
            float c = 0.f;
            if (j < m && !tk) {
                const long long cj = cand[j];
                if (cj >= 0 && cj != lo + i) {
                    c = di;
                    if ((mask >> j) & 1u) {
                        const float dv = cd[(int64_t)j * n + i];
                        const float dd = __fmul_rn(dv, dv);
                        c = (dd < di) ? dd : di;
                    }
                }
            }
            v[q] = c;
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = __fadd_rn(v[q], __shfl_xor_sync(0xffffffffu, v[q], o));
        }
        __syncthreads();  // ws of the previous group consumed
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q) ws[q][w] = v[q];
        }
        __syncthreads();
        if (w < 4 && j0 + w < m) {
            float s = ws[w][lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
            if (lane == 0) out10[(int64_t)(j0 + w) * out10_stride + blockIdx.x] = s;
        }
    }
}

// roots of m trees: the first stored level (height 10, 20 or 30) that has a single entry IS the
// root (the levels above the true height only add zeros, and v + 0 == v exactly)
__global__ void kmpp_tree_roots_kernel(const float* __restrict__ lv, int64_t stride, int m, float* __restrict__ out) {
    const int j = threadIdx.x;
    if (j < m) out[j] = lv[(int64_t)j * stride];
}

// ---- selection -----------------------------------------------------------------------------
__global__ void kmpp_select_kernel(KmppState* st, const float* __restrict__ pots, int m,
                                   const unsigned char* __restrict__ taken, int64_t n) {
    if (threadIdx.x != 0) return;
    long long best = -1;
    int jbest = -1;
    float bp = 3.402823466e+38f;
    for (int j = 0; j < m; ++j) {
        st->pot[j] = pots[j];
        if (st->cand[j] >= 0 && pots[j] < bp) { bp = pots[j]; best = st->cand[j]; jbest = j; }
    }
    st->best = best;
    st->jbest = jbest;
}

__global__ void kmpp_first_free_kernel(const unsigned char* __restrict__ taken, int64_t n, long long* out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !taken[i]) atomicMin((unsigned long long*)out, (unsigned long long)i);
}

__global__ void kmpp_commit_kernel(KmppState* st, long long best, const float* __restrict__ X, int d,
                                   const float* __restrict__ D, unsigned char* __restrict__ taken,
                                   float* __restrict__ center_out) {
    for (int e = threadIdx.x; e < d; e += blockDim.x) center_out[e] = X[best * d + e];
    if (threadIdx.x == 0) {
        taken[best] = 1;
        st->dist_sum = __fsub_rn(st->dist_sum, D[best]);
    }
}

__global__ void kmpp_add_delta_kernel(KmppState* st, const float* __restrict__ s) {
    st->dist_sum = __fadd_rn(st->dist_sum, s[0]);
}

// ---- driver ----------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) dev_free(p); }
    int alloc(size_t bytes) {
        if (dev_alloc(&p, bytes ? bytes : 16) != cudaSuccess) {
            p = nullptr;
            cudaGetLastError();
            return set_error(B2K_ERR_NOMEM, "dev_alloc(%zu) failed", bytes);
        }
        return B2K_OK;
    }
    template <class T> T* as() { return (T*)p; }
};

static int pow2_height(int64_t n) { int h = 0; while ((int64_t(1) << h) < n) ++h; return h; }

int kmpp_run_blocked(b2k_ctx* ctx, const float* dX, int64_t n, int d, int k, int metric, int64_t seed, int64_t lo,
                     int64_t n_total, float* xf, int64_t xf_len, int64_t* xi, b2k_exchange_fn ex, void* exuser,
                     b2k_callback cb, void* user, float* dcenters_out, int64_t* chosen_host);

int kmpp_run(b2k_ctx* ctx, const float* dX, int64_t n, int d, int k, int metric, int64_t seed, int scan_mode,
             b2k_callback cb, void* user, float* dcenters_out, int64_t* chosen_host) {
    if (k < 1 || k > n) return set_error(B2K_ERR_INVALID_ARG, "k-means++: need 1 <= k <= n (k=%d, n=%lld)", k, (long long)n);
    if (metric == B2K_METRIC_MINRMSD && d % 3) return set_error(B2K_ERR_DIM_NOT_MULT3, "RMSDMetric is only implemented for input data with a dimension divisible by 3.");
    if (scan_mode == B2K_KMPP_BLOCKED)
        return kmpp_run_blocked(ctx, dX, n, d, k, metric, seed, 0, n, nullptr, 0, nullptr, nullptr, nullptr, cb, user,
                                dcenters_out, chosen_host);
    const int m = 2 + (int)std::log((double)k);
    if (m > KMPP_MAX_TRIALS) return set_error(B2K_ERR_INVALID_ARG, "k too large");
    cudaStream_t st = ctx->stream;

    // RNG stream (data independent)
    uint32_t s32;
    if (seed < 0) { std::random_device rd; s32 = rd(); } else s32 = (uint32_t)seed;
    MT19937 gen(s32);
    const int64_t first = (int64_t)gen.below((uint64_t)n);
    std::vector<float> u((size_t)(k > 1 ? (k - 1) : 1) * m);
    for (size_t t = 0; t < (size_t)(k - 1) * m; ++t) u[t] = gen.unit();

    DevBuf bD, bTaken, bCd, bDelta, bRows, bRowsC, bGb, bGa, bU, bState, bPots, bFree, bL5, bL10, bL15, bL20, bL25, bL30,
        bP5, bP10, bP15, bP20, bP25, bP30;
    B2K_TRY(bD.alloc(n * 4));
    B2K_TRY(bTaken.alloc(n));
    B2K_TRY(bCd.alloc((size_t)m * n * 4));
    B2K_TRY(bRows.alloc((size_t)m * d * 4));
    B2K_TRY(bU.alloc(u.size() * 4));
    B2K_TRY(bState.alloc(sizeof(KmppState)));
    B2K_TRY(bPots.alloc(KMPP_MAX_TRIALS * 4));
    B2K_TRY(bFree.alloc(8));
    float* D = bD.as<float>();
    unsigned char* taken = bTaken.as<unsigned char>();
    float* cd = bCd.as<float>();
    float* rows = bRows.as<float>();
    KmppState* S = bState.as<KmppState>();
    float* pots = bPots.as<float>();
    CUDA_TRY(cudaMemcpyAsync(bU.p, u.data(), u.size() * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(S, 0, sizeof(KmppState), st));

    float* Ga = nullptr;
    if (metric == B2K_METRIC_MINRMSD) {
        B2K_TRY(bGa.alloc(n * 4));
        B2K_TRY(bRowsC.alloc((size_t)m * d * 4));
        B2K_TRY(bGb.alloc(m * 4));
        Ga = bGa.as<float>();
        B2K_TRY(launch_rmsd_center(ctx, dX, n, d, nullptr, Ga));
    }
    auto dist_rows = [&](const float* R, int mm, float* out) -> int {
        if (metric == B2K_METRIC_MINRMSD) {
            B2K_TRY(launch_rmsd_center(ctx, R, mm, d, bRowsC.as<float>(), bGb.as<float>()));
            return launch_rmsd_dist_rows(ctx, dX, Ga, n, d, bRowsC.as<float>(), bGb.as<float>(), mm, out);
        }
        return launch_dist_rows(ctx, dX, n, d, R, mm, out);
    };

    // tree levels (blocked mode)
    const int H = pow2_height(n);
    const int64_t n5 = cdiv(n, 32), n10 = cdiv(n, 1024), n15 = cdiv(n10, 32), n20 = cdiv(n10, 1024),
                  n25 = cdiv(n20, 32), n30 = cdiv(n20, 1024);
    if (scan_mode == B2K_KMPP_BLOCKED) {
        B2K_TRY(bL5.alloc(n5 * 4)); B2K_TRY(bL10.alloc(n10 * 4)); B2K_TRY(bL15.alloc(n15 * 4));
        B2K_TRY(bL20.alloc(n20 * 4)); B2K_TRY(bL25.alloc(n25 * 4)); B2K_TRY(bL30.alloc(n30 * 4));
        B2K_TRY(bP10.alloc((size_t)m * n10 * 4)); B2K_TRY(bP20.alloc((size_t)m * n20 * 4));
        B2K_TRY(bP30.alloc((size_t)m * n30 * 4));
    } else {
        B2K_TRY(bDelta.alloc(n * 4));
    }
    // build the stored levels of mm trees over `in` (stride n), return top stored level info
    auto tree_build = [&](const float* in, int mm, const unsigned char* mask, float* l5, float* l10, float* l15,
                          float* l20, float* l25, float* l30) -> int {
        tree_up_kernel<<<dim3((unsigned)n10, mm), 1024, 0, st>>>(in, n, n, mask, l5, n5, l10, n10);
        LAUNCH_CHECK();
        if (H > 10) {
            tree_up_kernel<<<dim3((unsigned)n20, mm), 1024, 0, st>>>(l10, n10, n10, nullptr, l15, n15, l20, n20);
            LAUNCH_CHECK();
        }
        if (H > 20) {
            tree_up_kernel<<<dim3((unsigned)n30, mm), 1024, 0, st>>>(l20, n20, n20, nullptr, l25, n25, l30, n30);
            LAUNCH_CHECK();
        }
        return B2K_OK;
    };
    auto tree_roots = [&](float* l10, float* l20, float* l30, int mm, float* out) -> int {
        if (H > 30) return set_error(B2K_ERR_INVALID_ARG, "n too large");
        const float* lv = H <= 10 ? l10 : (H <= 20 ? l20 : l30);
        const int64_t stride = H <= 10 ? n10 : (H <= 20 ? n20 : n30);
        kmpp_tree_roots_kernel<<<1, 32, 0, st>>>(lv, stride, mm, out);
        LAUNCH_CHECK();
        return B2K_OK;
    };

    // ---- first center ----
    std::vector<int64_t> chosen((size_t)k, -1);
    chosen[0] = first;
    CUDA_TRY(cudaMemcpyAsync(dcenters_out, dX + first * d, (size_t)d * 4, cudaMemcpyDeviceToDevice, st));
    if (cb) { CUDA_TRY(cudaStreamSynchronize(st)); cb(user); }
    if (k == 1) {
        CUDA_TRY(cudaStreamSynchronize(st));
        if (chosen_host) chosen_host[0] = first;
        return B2K_OK;
    }
    B2K_TRY(dist_rows(dcenters_out, 1, cd));
    kmpp_square_init_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(cd, n, first, D, taken);
    LAUNCH_CHECK();
    if (scan_mode == B2K_KMPP_SERIAL) {
        kmpp_serial_sum_kernel<<<1, 256, 0, st>>>(D, n, 1, pots);
        LAUNCH_CHECK();
        kmpp_set_dist_sum_kernel<<<1, 1, 0, st>>>(S, pots);
        LAUNCH_CHECK();
    }

    TreeLevels T;
    T.H = H;
    T.lv[0] = D; T.len[0] = n;
    T.lv[1] = bL5.as<float>(); T.len[1] = n5;
    T.lv[2] = bL10.as<float>(); T.len[2] = n10;
    T.lv[3] = bL15.as<float>(); T.len[3] = n15;
    T.lv[4] = bL20.as<float>(); T.len[4] = n20;
    T.lv[5] = bL25.as<float>(); T.len[5] = n25;
    T.lv[6] = bL30.as<float>(); T.len[6] = n30;

    KmppState hs;
    for (int found = 1; found < k; ++found) {
        const float* ur = bU.as<float>() + (size_t)(found - 1) * m;
        // candidates
        if (scan_mode == B2K_KMPP_SERIAL) {
            kmpp_serial_scan_kernel<<<1, 32, 0, st>>>(D, taken, n, S, ur, m);
            LAUNCH_CHECK();
        } else {
            B2K_TRY(tree_build(D, 1, taken, bL5.as<float>(), bL10.as<float>(), bL15.as<float>(), bL20.as<float>(),
                               bL25.as<float>(), bL30.as<float>()));
            kmpp_tree_pick_kernel<<<1, 32, 0, st>>>(T, taken, n, S, ur, m);
            LAUNCH_CHECK();
        }
        kmpp_gather_kernel<<<m, 128, 0, st>>>(dX, d, S, m, rows);
        LAUNCH_CHECK();
        B2K_TRY(dist_rows(rows, m, cd));
        kmpp_contrib_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(cd, n, m, D, taken, S);
        LAUNCH_CHECK();
        // potentials
        if (scan_mode == B2K_KMPP_SERIAL) {
            kmpp_serial_sum_kernel<<<1, 256, 0, st>>>(cd, n, m, pots);
            LAUNCH_CHECK();
        } else {
            B2K_TRY(tree_build(cd, m, nullptr, nullptr, bP10.as<float>(), nullptr, bP20.as<float>(), nullptr,
                               bP30.as<float>()));
            B2K_TRY(tree_roots(bP10.as<float>(), bP20.as<float>(), bP30.as<float>(), m, pots));
        }
        kmpp_select_kernel<<<1, 32, 0, st>>>(S, pots, m, taken, n);
        LAUNCH_CHECK();
        CUDA_TRY(cudaMemcpyAsync(&hs, S, sizeof(KmppState), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        long long best = hs.best;
        int jbest = hs.jbest;
        if (best < 0) {  // "if for some reason we did not find a best candidate, take the next available point"
            long long init = 0x7fffffffffffffffll;
            CUDA_TRY(cudaMemcpyAsync(bFree.p, &init, 8, cudaMemcpyHostToDevice, st));
            kmpp_first_free_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(taken, n, bFree.as<long long>());
            LAUNCH_CHECK();
            CUDA_TRY(cudaMemcpyAsync(&best, bFree.p, 8, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            if (best == init) break;
            jbest = -1;
        }
        chosen[found] = best;
        kmpp_commit_kernel<<<1, 128, 0, st>>>(S, best, dX, d, D, taken, dcenters_out + (size_t)found * d);
        LAUNCH_CHECK();
        if (cb) cb(user);
        if (found + 1 < k) {
            float* delta = scan_mode == B2K_KMPP_SERIAL ? bDelta.as<float>() : nullptr;
            if (jbest >= 0) {
                kmpp_update_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(D, taken, n, cd + (size_t)jbest * n, 0, delta);
                LAUNCH_CHECK();
            } else {
                B2K_TRY(dist_rows(dcenters_out + (size_t)found * d, 1, cd));
                kmpp_update_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(D, taken, n, cd, 1, delta);
                LAUNCH_CHECK();
            }
            if (scan_mode == B2K_KMPP_SERIAL) {
                kmpp_serial_sum_kernel<<<1, 256, 0, st>>>(delta, n, 1, pots);
                LAUNCH_CHECK();
                kmpp_add_delta_kernel<<<1, 1, 0, st>>>(S, pots);
                LAUNCH_CHECK();
            }
        }
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    if (chosen_host) std::memcpy(chosen_host, chosen.data(), sizeof(int64_t) * k);
    for (int i = 0; i < k; ++i)
        if (chosen[i] < 0) return set_error(B2K_ERR_INVALID_ARG, "k-means++ could not find %d centers", k);
    return B2K_OK;
}

// ---- BLOCKED mode driver, optionally over a shard of the frames ---------------------------------------------------
// Single GPU: lo = 0, n_total = n_local, ex = null.  Sharded: every rank calls this with its shard (lo a multiple
// of 1024 so that no height-10 tree node straddles two shards), an exchange buffer pair the caller owns and the
// callback that all-reduces it.  Per round: [sum] the height-10 sums of D, [max] the candidates' frame indices,
// [sum] the candidate rows, [sum] the height-10 sums of the m contribution arrays.  Every exchanged sum only ever
// adds zeros to the one owner's value, so all ranks hold bit-identical trees and take identical decisions.
int kmpp_run_blocked(b2k_ctx* ctx, const float* dX, int64_t n, int d, int k, int metric, int64_t seed, int64_t lo,
                     int64_t n_total, float* xf, int64_t xf_len, int64_t* xi, b2k_exchange_fn ex, void* exuser,
                     b2k_callback cb, void* user, float* dcenters_out, int64_t* chosen_host) {
    if (k < 1 || k > n_total)
        return set_error(B2K_ERR_INVALID_ARG, "k-means++: need 1 <= k <= n (k=%d, n=%lld)", k, (long long)n_total);
    if (metric == B2K_METRIC_MINRMSD && d % 3) return set_error(B2K_ERR_DIM_NOT_MULT3, "RMSDMetric is only implemented for input data with a dimension divisible by 3.");
    if (n > 0 && lo % 1024)  // an empty shard owns no tree node: its offset does not matter
        return set_error(B2K_ERR_INVALID_ARG, "k-means++: shard offset must be a multiple of 1024");
    const int m = 2 + (int)std::log((double)k);
    if (m > KMPP_MAX_TRIALS) return set_error(B2K_ERR_INVALID_ARG, "k too large");
    cudaStream_t st = ctx->stream;
    const int H = pow2_height(n_total);
    if (H > 30) return set_error(B2K_ERR_INVALID_ARG, "n too large");
    const int Hs = std::max(H, 10);
    const int64_t n10g = cdiv(n_total, 1024), n15g = cdiv(n10g, 32), n20g = cdiv(n10g, 1024), n25g = cdiv(n20g, 32),
                  n30g = cdiv(n20g, 1024);
    const int64_t n5 = cdiv(std::max<int64_t>(n, 1), 32), n10 = cdiv(n, 1024);  // local
    const int64_t node_lo = lo / 1024;
    const int64_t need_f = std::max<int64_t>((int64_t)m * n10g, (int64_t)m * d);
    DevBuf bXf, bXi;
    if (!ex) {  // single GPU: the "exchange" buffers are private
        B2K_TRY(bXf.alloc((size_t)need_f * 4));
        B2K_TRY(bXi.alloc(64 * 8));
        xf = bXf.as<float>();
        xi = bXi.as<int64_t>();
    } else if (!xf || !xi || xf_len < need_f) {
        return set_error(B2K_ERR_INVALID_ARG, "k-means++: exchange buffer too small (need %lld floats)", (long long)need_f);
    }
    auto exchange = [&](int which, int64_t count, int op) -> int {
        if (!ex) return B2K_OK;
        CUDA_TRY(cudaStreamSynchronize(st));
        if (ex(exuser, which, count, op) != 0) return set_error(B2K_ERR_CUDA, "k-means++: exchange callback failed");
        return B2K_OK;
    };

    // RNG stream (data independent, identical on every rank)
    uint32_t s32;
    if (seed < 0) { std::random_device rd; s32 = rd(); } else s32 = (uint32_t)seed;
    MT19937 gen(s32);
    const int64_t first = (int64_t)gen.below((uint64_t)n_total);
    std::vector<float> u((size_t)(k > 1 ? (k - 1) : 1) * m);
    for (size_t t = 0; t < (size_t)(k - 1) * m; ++t) u[t] = gen.unit();

    DevBuf bD, bTaken, bCd, bRows, bRowsC, bGb, bGa, bU, bState, bPots, bL5, bL10g, bL15, bL20, bL25, bL30, bP20, bP30,
        bNode, bResid, bAssigned, bRc, bList, bMasks, bCount, bFrameMask;
    // (1 = automatic: a job whose frames sit in L2 -- up to 4M floats -- gains nothing from skipped reads and pays three
    // launches per round instead of one: cfg1 fit 6.0 -> 5.5 ms without; 2 = always, 0 = never; picks identical)
    const bool prune = ctx->kmpp_prune != 0 && metric == B2K_METRIC_EUCLIDEAN && m <= 14 &&
                       (ctx->kmpp_prune == 2 || n * (int64_t)d > (int64_t(1) << 22));
    const int rc_stride = 16;  // Rc is [center][16]
    const int64_t nn = std::max<int64_t>(n, 1);
    B2K_TRY(bD.alloc(nn * 4));
    B2K_TRY(bTaken.alloc(nn));
    B2K_TRY(bCd.alloc((size_t)m * nn * 4));
    B2K_TRY(bRows.alloc((size_t)m * d * 4));
    B2K_TRY(bU.alloc(u.size() * 4));
    B2K_TRY(bState.alloc(sizeof(KmppState)));
    B2K_TRY(bPots.alloc(KMPP_MAX_TRIALS * 4));
    B2K_TRY(bL5.alloc(n5 * 4));
    B2K_TRY(bL10g.alloc(n10g * 4)); B2K_TRY(bL15.alloc(n15g * 4)); B2K_TRY(bL20.alloc(n20g * 4));
    B2K_TRY(bL25.alloc(n25g * 4)); B2K_TRY(bL30.alloc(n30g * 4));
    B2K_TRY(bP20.alloc((size_t)m * n20g * 4)); B2K_TRY(bP30.alloc((size_t)m * n30g * 4));
    B2K_TRY(bNode.alloc(KMPP_MAX_TRIALS * 8)); B2K_TRY(bResid.alloc(KMPP_MAX_TRIALS * 4));
    if (prune) {
        B2K_TRY(bAssigned.alloc(nn * 4));
        B2K_TRY(bRc.alloc((size_t)k * rc_stride * 4));
        CUDA_TRY(cudaMemsetAsync(bRc.p, 0, (size_t)k * rc_stride * 4, st));
        B2K_TRY(bFrameMask.alloc(nn * 2));
        B2K_TRY(bList.alloc(nn * 4));
        B2K_TRY(bMasks.alloc(nn * 4));
        B2K_TRY(bCount.alloc(16));
        CUDA_TRY(cudaMemsetAsync(bAssigned.p, 0, nn * 4, st));  // D2 starts as the distance to center 0
    }
    float* D = bD.as<float>();
    unsigned char* taken = bTaken.as<unsigned char>();
    float* cd = bCd.as<float>();
    float* rows = bRows.as<float>();
    KmppState* S = bState.as<KmppState>();
    float* pots = bPots.as<float>();
    long long* xil = reinterpret_cast<long long*>(xi);
    CUDA_TRY(cudaMemcpyAsync(bU.p, u.data(), u.size() * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(S, 0, sizeof(KmppState), st));

    float* Ga = nullptr;
    if (metric == B2K_METRIC_MINRMSD) {
        B2K_TRY(bGa.alloc(nn * 4));
        B2K_TRY(bRowsC.alloc((size_t)m * d * 4));
        B2K_TRY(bGb.alloc(m * 4));
        Ga = bGa.as<float>();
        B2K_TRY(launch_rmsd_center(ctx, dX, n, d, nullptr, Ga));
    }
    auto dist_rows = [&](const float* R, int mm, float* out) -> int {
        if (n <= 0) return B2K_OK;
        if (metric == B2K_METRIC_MINRMSD) {
            B2K_TRY(launch_rmsd_center(ctx, R, mm, d, bRowsC.as<float>(), bGb.as<float>()));
            return launch_rmsd_dist_rows(ctx, dX, Ga, n, d, bRowsC.as<float>(), bGb.as<float>(), mm, out);
        }
        return launch_dist_rows(ctx, dX, n, d, R, mm, out);
    };
    // rows of the frames with the global indices xil[0..mm) -> `rows` on every rank
    auto fetch_rows = [&](int mm) -> int {
        if (!ex) {
            kmpp_gather_sharded_kernel<<<mm, 128, 0, st>>>(dX, d, xil, lo, n, rows);
            LAUNCH_CHECK();
            return B2K_OK;
        }
        kmpp_gather_sharded_kernel<<<mm, 128, 0, st>>>(dX, d, xil, lo, n, xf);
        LAUNCH_CHECK();
        B2K_TRY(exchange(0, (int64_t)mm * d, 0));
        CUDA_TRY(cudaMemcpyAsync(rows, xf, (size_t)mm * d * 4, cudaMemcpyDeviceToDevice, st));
        return B2K_OK;
    };
    // upper levels (heights 15..30) of mm trees whose height-10 sums are l10 (stride n10g)
    auto tree_top = [&](const float* l10, int mm, float* l15, float* l20, float* l25, float* l30) -> int {
        if (Hs > 10) {
            tree_up_kernel<<<dim3((unsigned)n20g, mm), 1024, 0, st>>>(l10, n10g, n10g, nullptr, l15, n15g, l20, n20g);
            LAUNCH_CHECK();
        }
        if (Hs > 20) {
            tree_up_kernel<<<dim3((unsigned)n30g, mm), 1024, 0, st>>>(l20, n20g, n20g, nullptr, l25, n25g, l30, n30g);
            LAUNCH_CHECK();
        }
        return B2K_OK;
    };

    // ---- first center ----
    std::vector<int64_t> chosen((size_t)k, -1);
    chosen[0] = first;
    CUDA_TRY(cudaMemcpyAsync(xi, &first, 8, cudaMemcpyHostToDevice, st));
    B2K_TRY(fetch_rows(1));
    CUDA_TRY(cudaMemcpyAsync(dcenters_out, rows, (size_t)d * 4, cudaMemcpyDeviceToDevice, st));
    if (cb) { CUDA_TRY(cudaStreamSynchronize(st)); cb(user); }
    if (k == 1) {
        CUDA_TRY(cudaStreamSynchronize(st));
        if (chosen_host) chosen_host[0] = first;
        return B2K_OK;
    }
    B2K_TRY(dist_rows(dcenters_out, 1, cd));
    if (n > 0) {
        kmpp_square_init_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(cd, n, first - lo, D, taken);
        LAUNCH_CHECK();
    }

    TreeLevels T;
    T.H = H;
    T.lv[0] = D; T.len[0] = n;
    T.lv[1] = bL5.as<float>(); T.len[1] = n5;
    T.lv[2] = bL10g.as<float>(); T.len[2] = n10g;
    T.lv[3] = bL15.as<float>(); T.len[3] = n15g;
    T.lv[4] = bL20.as<float>(); T.len[4] = n20g;
    T.lv[5] = bL25.as<float>(); T.len[5] = n25g;
    T.lv[6] = bL30.as<float>(); T.len[6] = n30g;

    // Asynchronous rounds (single GPU, no progress callback): the accepted candidate stays on the device -- commit and D2
    // update read it from the state -- so the host queues round after round without waiting; the picks come back once at
    // the end.  A round without a usable candidate (the reference's "take the next available point") raises a device flag
    // and the whole seeding is repeated with the synchronous loop below.
    const bool async_rounds = !ex && !cb && ctx->kmpp_async != 0 && n > 0;
    DevBuf bChosen, bFail;
    if (async_rounds) {
        B2K_TRY(bChosen.alloc((size_t)k * 8));
        B2K_TRY(bFail.alloc(16));
        CUDA_TRY(cudaMemsetAsync(bFail.p, 0, 16, st));
    }
    if (async_rounds) {
        kmpp_set_round_kernel<<<1, 1, 0, st>>>(S, 1);
        LAUNCH_CHECK();
    }
    // Asynchronous rounds are identical launches (every per-round quantity lives in the device state), so from the third
    // round on they are replayed from ONE captured CUDA graph: ~15 kernel nodes per graph launch instead of ~15 launches.
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    bool capturing = false;
    KmppState hs;
    for (int found = 1; found < k; ++found) {
        if (async_rounds && gexec) {
            CUDA_TRY(cudaGraphLaunch(gexec, st));
            g_launches.fetch_add(1);
            continue;
        }
        if (async_rounds && ctx->kmpp_async == 2 && found == 2 && k > 3) {
            capturing = cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) == cudaSuccess;
            if (!capturing) cudaGetLastError();
        }
        const float* ur = async_rounds ? bU.as<float>() : bU.as<float>() + (size_t)(found - 1) * m;
        // ---- tree over D: local heights 5 and 10, [sum] height 10, replicated upper levels ----
        float* l10 = ex ? xf : bL10g.as<float>();
        if (ex) CUDA_TRY(cudaMemsetAsync(xf, 0, (size_t)n10g * 4, st));
        if (n > 0 && !(async_rounds && found > 1)) {  // (asynchronous rounds: the previous round's update built it already)
            tree_up_kernel<<<dim3((unsigned)n10, 1), 1024, 0, st>>>(D, n, n, taken, bL5.as<float>(), n5, l10 + node_lo, n10g);
            LAUNCH_CHECK();
        }
        if (ex) {
            B2K_TRY(exchange(0, n10g, 0));
            CUDA_TRY(cudaMemcpyAsync(bL10g.p, xf, (size_t)n10g * 4, cudaMemcpyDeviceToDevice, st));
        }
        // (small trees on the asynchronous path: built inside the pick / select kernels, two launches per round fewer)
        const bool fuse_top = async_rounds && Hs > 10 && Hs <= 20 && n10g <= 1024 && n20g == 1;
        if (!fuse_top)
            B2K_TRY(tree_top(bL10g.as<float>(), 1, bL15.as<float>(), bL20.as<float>(), bL25.as<float>(), bL30.as<float>()));
        // ---- candidates ----
        if (async_rounds) {  // (lo = 0, one shard: pick, registration and row gather in one launch)
            kmpp_pick_fused_kernel<<<1, 32 * m, 0, st>>>(T, Hs, S, bU.as<float>(), m, taken, n, dX, d, xil, rows,
                                                         fuse_top ? bL15.as<float>() : nullptr, bL20.as<float>());
            LAUNCH_CHECK();
        } else {
            kmpp_pick_top_kernel<<<1, 32, 0, st>>>(T, Hs, S, ur, m, bNode.as<long long>(), bResid.as<float>(), 0);
            LAUNCH_CHECK();
            kmpp_pick_leaf_kernel<<<1, 32, 0, st>>>(T, taken, n, node_lo, lo, bNode.as<long long>(), bResid.as<float>(), m, xil);
            LAUNCH_CHECK();
            B2K_TRY(exchange(1, m, 1));
            kmpp_set_cands_kernel<<<1, 32, 0, st>>>(S, xil, m);
            LAUNCH_CHECK();
            B2K_TRY(fetch_rows(m));  // reads xil; `rows` valid on every rank afterwards
        }
        // ---- potentials ----
        if (prune && n > 0) {
            // distances of the m candidates to the `found` centers chosen so far, then the pruned distance rows
            if (async_rounds)  // fixed grid, `found` read from the device state
                kmpp_center_cand_dist_kernel<<<(unsigned)std::min<int64_t>(cdiv((int64_t)k * m * 32, 256), ctx->sm_count * 8), 256, 0,
                                               st>>>(dcenters_out, found, rows, m, d, bRc.as<float>(), rc_stride, S);
            else
                kmpp_center_cand_dist_kernel<<<(unsigned)cdiv((int64_t)found * m * 32, 256), 256, 0, st>>>(
                    dcenters_out, found, rows, m, d, bRc.as<float>(), rc_stride);
            LAUNCH_CHECK();
            B2K_TRY(launch_dist_rows_pruned(ctx, dX, n, d, rows, m, cd, D, bAssigned.as<int32_t>(), taken,
                                            bRc.as<float>(), rc_stride, bList.as<uint32_t>(), bMasks.as<uint32_t>(),
                                            bCount.as<unsigned int>(), bFrameMask.as<uint16_t>()));
        } else {
            B2K_TRY(dist_rows(rows, m, cd));
        }
        if (ex) CUDA_TRY(cudaMemsetAsync(xf, 0, (size_t)m * n10g * 4, st));
        // (asynchronous unpruned rounds use the same kernel without a mask: the contributions are never written, one launch
        // instead of two; the D2 update then squares the distance row itself)
        const bool pot_direct = prune || (async_rounds && m <= 16);
        if (n > 0 && pot_direct) {
            kmpp_pot_tree_kernel<<<(unsigned)n10, 1024, 0, st>>>(cd, n, m, D, taken, prune ? bFrameMask.as<uint16_t>() : nullptr,
                                                                 S->cand, lo, xf + node_lo, n10g);
            LAUNCH_CHECK();
        } else if (n > 0) {
            kmpp_contrib_sharded_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(cd, n, m, D, taken, S->cand, lo);
            LAUNCH_CHECK();
            tree_up_kernel<<<dim3((unsigned)n10, m), 1024, 0, st>>>(cd, n, n, nullptr, nullptr, 0, xf + node_lo, n10g);
            LAUNCH_CHECK();
        }
        B2K_TRY(exchange(0, (int64_t)m * n10g, 0));
        if (!fuse_top) B2K_TRY(tree_top(xf, m, nullptr, bP20.as<float>(), nullptr, bP30.as<float>()));
        {
            const float* lv = Hs <= 10 ? xf : (Hs <= 20 ? bP20.as<float>() : bP30.as<float>());
            const int64_t stride = Hs <= 10 ? n10g : (Hs <= 20 ? n20g : n30g);
            if (async_rounds) {
                kmpp_select_commit_kernel<<<1, 128, 0, st>>>(S, lv, stride, m, pots, n, rows, d, taken, dcenters_out,
                                                             bChosen.as<long long>(), bFail.as<int>(), k,
                                                             fuse_top ? xf : nullptr, n10g, bP20.as<float>());
            } else {
                kmpp_tree_roots_kernel<<<1, 32, 0, st>>>(lv, stride, m, pots);
                LAUNCH_CHECK();
                kmpp_select_kernel<<<1, 32, 0, st>>>(S, pots, m, taken, n);
            }
            LAUNCH_CHECK();
        }
        if (async_rounds) {
            // (also after the last pick: the D2 update is then unused, but every round stays the same launch sequence)
            // ... fused with the next round's tree over the updated D2 (heights 5 and 10)
            kmpp_update_tree_kernel<<<(unsigned)n10, 1024, 0, st>>>(
                D, taken, n, cd, S, pot_direct ? 1 : 0, prune ? bAssigned.as<int32_t>() : nullptr,
                prune ? bFrameMask.as<uint16_t>() : nullptr, bFail.as<int>(), bL5.as<float>(), bL10g.as<float>());
            LAUNCH_CHECK();
            if (capturing) {
                capturing = false;
                if (cudaStreamEndCapture(st, &graph) == cudaSuccess && graph &&
                    cudaGraphInstantiate(&gexec, graph, 0) == cudaSuccess) {
                    CUDA_TRY(cudaGraphLaunch(gexec, st));  // the captured round itself has not run yet
                    g_launches.fetch_add(1);
                } else {
                    cudaGetLastError();
                    if (graph) cudaGraphDestroy(graph);
                    graph = nullptr;
                    gexec = nullptr;
                    return set_error(B2K_ERR_CUDA, "k-means++: graph capture of a round failed");
                }
            }
            continue;
        }
        CUDA_TRY(cudaMemcpyAsync(&hs, S, sizeof(KmppState), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        long long best = hs.best;
        int jbest = hs.jbest;
        const float* best_row = rows + (size_t)(jbest >= 0 ? jbest : 0) * d;
        if (best < 0) {  // "if for some reason we did not find a best candidate, take the next available point"
            long long init = 0x7fffffffffffffffll;
            CUDA_TRY(cudaMemcpyAsync(xi, &init, 8, cudaMemcpyHostToDevice, st));
            if (n > 0) {
                kmpp_first_free_sharded_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(taken, n, lo, xil);
                LAUNCH_CHECK();
            }
            B2K_TRY(exchange(1, 1, 2));
            CUDA_TRY(cudaMemcpyAsync(&best, xi, 8, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            if (best == init) break;
            jbest = -1;
            B2K_TRY(fetch_rows(1));
            best_row = rows;
        }
        chosen[found] = best;
        kmpp_commit_sharded_kernel<<<1, 128, 0, st>>>(best, lo, n, best_row, d, taken, dcenters_out + (size_t)found * d);
        LAUNCH_CHECK();
        if (cb) cb(user);
        if (found + 1 < k && n > 0) {
            if (jbest >= 0) {
                // pruned mode: cd holds raw distances of the live pairs only (no contribution pass rewrote it)
                kmpp_update_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(D, taken, n, cd + (size_t)jbest * n, prune ? 1 : 0,
                                                                           nullptr, prune ? bAssigned.as<int32_t>() : nullptr,
                                                                           found, prune ? bFrameMask.as<uint16_t>() : nullptr,
                                                                           jbest);
                LAUNCH_CHECK();
            } else {
                B2K_TRY(dist_rows(dcenters_out + (size_t)found * d, 1, cd));
                kmpp_update_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(D, taken, n, cd, 1, nullptr,
                                                                           prune ? bAssigned.as<int32_t>() : nullptr, found);
                LAUNCH_CHECK();
            }
        }
    }
    if (async_rounds) {
        int fail = 0;
        if (gexec) {
            CUDA_TRY(cudaStreamSynchronize(st));
            cudaGraphExecDestroy(gexec);
            cudaGraphDestroy(graph);
        }
        CUDA_TRY(cudaMemcpyAsync(&fail, bFail.p, 4, cudaMemcpyDeviceToHost, st));
        if (k > 1)
            CUDA_TRY(cudaMemcpyAsync(chosen.data() + 1, bChosen.as<long long>() + 1, (size_t)(k - 1) * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (fail) {  // rare: repeat with the synchronous loop, which handles the fallback pick on the host
            ctx->stat_kmpp_async_fallbacks += 1;
            const int keep = ctx->kmpp_async;
            ctx->kmpp_async = 0;
            const int rc = kmpp_run_blocked(ctx, dX, n, d, k, metric, seed, lo, n_total, xf, xf_len, xi, ex, exuser, cb, user,
                                            dcenters_out, chosen_host);
            ctx->kmpp_async = keep;
            return rc;
        }
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    if (chosen_host) std::memcpy(chosen_host, chosen.data(), sizeof(int64_t) * k);
    for (int i = 0; i < k; ++i)
        if (chosen[i] < 0) return set_error(B2K_ERR_INVALID_ARG, "k-means++ could not find %d centers", k);
    return B2K_OK;
}

}  // namespace b2k
