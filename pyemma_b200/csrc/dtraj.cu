// dtraj.cu -- consumers of the discrete trajectories: state histogram and lagged transition
// count matrix, computed where the labels already are (HBM) instead of copying them to the host.
//
// Replaces pyemma/util/discrete_trajectories.py:146-181 (count_states: np.bincount per trajectory) and the
// count matrix the MSM estimators build from dtrajs (pyemma/msm/estimators/_msm_estimator_base.py:3,229,332:
// deeptime.markov.tools.estimation.count_matrix(dtrajs, lag, sliding) -- C[i,j] = #{t : s_t = i, s_{t+lag} = j},
// t over all frames (sliding) or over multiples of lag (sample); pairs never span two trajectories).
// Integer counts, exact; HBM-bound (8 bytes of labels per pair) up to the L2 atomic rate for the scattered adds:
// lanes of a warp that hit the same cell (time-correlated trajectories stay in a state for many frames) are
// aggregated with match_any so that one RED carries the whole group.
#include "common.cuh"
#include "kernels.h"

namespace b2k {

// per-CTA histogram in shared memory when the states fit (ns <= 12288), flushed with one RED per visited state;
// otherwise warp-aggregated REDs straight to global memory
__global__ void __launch_bounds__(256) count_states_kernel(const int32_t* __restrict__ l, int64_t n, int ns,
                                                           unsigned long long* __restrict__ counts, int* bad,
                                                           int use_smem) {
    extern __shared__ uint32_t hsm[];
    const int lane = threadIdx.x & 31;
    if (use_smem) {
        for (int j = threadIdx.x; j < ns; j += 256) hsm[j] = 0u;
        __syncthreads();
    }
    const int64_t n_round = (n + 31) & ~(int64_t)31;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n_round; i += (int64_t)gridDim.x * 256) {
        const int32_t a = i < n ? l[i] : -1;
        const bool ok = a >= 0 && a < ns;
        if (i < n && a >= ns) *bad = 1;
        const unsigned peers = __match_any_sync(0xffffffffu, ok ? a : -1 - lane);
        if (ok && lane == __ffs(peers) - 1) {
            if (use_smem) atomicAdd(&hsm[a], (uint32_t)__popc(peers));
            else atomicAdd(counts + a, (unsigned long long)__popc(peers));
        }
    }
    if (use_smem) {
        __syncthreads();
        for (int j = threadIdx.x; j < ns; j += 256)
            if (hsm[j]) atomicAdd(counts + j, (unsigned long long)hsm[j]);
    }
}

__global__ void __launch_bounds__(256) count_matrix_kernel(const int32_t* __restrict__ l, int64_t n_pairs, int64_t lag,
                                                           int64_t step, int ns, unsigned long long* __restrict__ Cm,
                                                           int* bad) {
    const int lane = threadIdx.x & 31;
    const int64_t n_round = (n_pairs + 31) & ~(int64_t)31;
    for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < n_round; p += (int64_t)gridDim.x * 256) {
        int32_t a = -1, b = -1;
        if (p < n_pairs) { a = l[p * step]; b = l[p * step + lag]; }
        const bool ok = a >= 0 && b >= 0 && a < ns && b < ns;
        if (p < n_pairs && (a >= ns || b >= ns)) *bad = 1;
        const unsigned long long key = ok ? (unsigned long long)a * (unsigned)ns + (unsigned)b
                                          : ~0ull - (unsigned long long)lane;
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        if (ok && lane == __ffs(peers) - 1) atomicAdd(Cm + key, (unsigned long long)__popc(peers));
    }
}

}  // namespace b2k

using namespace b2k;

static unsigned dgrid(b2k_ctx* ctx, int64_t items) {
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv(items, 1024), (int64_t)ctx->sm_count * 8));
}

static int read_bad(b2k_ctx* ctx, const char* what) {
    int bad = 0;
    CUDA_TRY(cudaMemcpyAsync(&bad, ctx->flags, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (bad) return set_error(B2K_ERR_INVALID_ARG, "%s: a state index is >= nstates", what);
    return B2K_OK;
}

B2K_API int b2k_dev_count_states(b2k_ctx* ctx, const int32_t* dlabels, int64_t n, int32_t nstates, int64_t* dcounts) {
    if (!ctx || n < 0 || nstates < 1 || !dcounts || (n > 0 && !dlabels))
        return set_error(B2K_ERR_INVALID_ARG, "count_states: bad arguments");
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (n == 0) return B2K_OK;
    CUDA_TRY(cudaMemsetAsync(ctx->flags, 0, 4, ctx->stream));
    const int use_smem = nstates <= 12288;
    const unsigned grid = use_smem ? (unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv(n, 8192), (int64_t)ctx->sm_count * 4))
                                   : dgrid(ctx, n);
    count_states_kernel<<<grid, 256, use_smem ? (size_t)nstates * 4 : 0, ctx->stream>>>(
        dlabels, n, nstates, (unsigned long long*)dcounts, ctx->flags, use_smem);
    LAUNCH_CHECK();
    return read_bad(ctx, "count_states");
}

B2K_API int b2k_dev_count_matrix(b2k_ctx* ctx, const int32_t* dlabels, int64_t n, int32_t nstates, int64_t lag,
                                 int sliding, int64_t* dC) {
    if (!ctx || n < 0 || nstates < 1 || lag < 1 || !dC || (n > 0 && !dlabels))
        return set_error(B2K_ERR_INVALID_ARG, "count_matrix: bad arguments");
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (n <= lag) return B2K_OK;  // a trajectory not longer than the lag contributes nothing
    const int64_t step = sliding ? 1 : lag;
    const int64_t n_pairs = (n - lag - 1) / step + 1;  // t = 0, step, 2 step, ... with t + lag <= n - 1
    CUDA_TRY(cudaMemsetAsync(ctx->flags, 0, 4, ctx->stream));
    count_matrix_kernel<<<dgrid(ctx, n_pairs), 256, 0, ctx->stream>>>(dlabels, n_pairs, lag, step, nstates,
                                                                      (unsigned long long*)dC, ctx->flags);
    LAUNCH_CHECK();
    return read_bad(ctx, "count_matrix");
}
