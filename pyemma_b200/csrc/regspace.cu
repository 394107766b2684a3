// regspace.cu -- regular-space (leader) clustering (K6 of SURVEY 2.2).
//
// Replaces deeptime regspace.cluster(chunk, centers, dmin, max_centers, n_threads) as driven by
// RegularSpaceClustering._estimate (pyemma/coordinates/clustering/regspace.py:144-151): frames
// are visited in order; a frame whose distance to EVERY center found so far is > dmin becomes a
// center (copy of the frame); trying to add center max_centers+1 raises
// MaxCentersReachedException (regspace.py:153-163) with the first max_centers centers kept.
//
// The dependence is sequential, but only through the centers discovered inside the current chunk:
//   pass 1 (parallel)  min distance of every chunk frame to the centers known before the chunk;
//                      frames within dmin can never become centers -> dead.
//   steps  (ordered)   the FIRST surviving frame is a new center by construction (everything before
//                      it is dead and centers only ever get added); every later survivor within
//                      dmin of it dies; repeat.  One step = 2 small kernels, no host round trip:
//                      the index of the next first-survivor is produced by atomicMin inside the
//                      kill kernel.  The host only polls a status word every STEP_BATCH steps.
// The predicate is exactly the reference's (`min_j compute(x_i,c_j) > dmin`, dmin in fp32), the
// distances are the exact reference-order kernels, so centers (order and bits) match.
#include "common.cuh"
#include "kernels.h"

namespace b2k {

// rmsd.cu device helpers are file-local there; the single-row variants needed here are small
// enough to restate via the launchers: the kill kernel below therefore works on a distance
// array produced by launch_dist_rows / launch_rmsd_dist_rows for ONE row whose device address is
// fixed (cur_row), which the append kernel refreshes.

struct RegState {
    long long n_centers;
    long long max_centers;
    int status;          // 0 running, 1 chunk exhausted, 4 max centers reached
    int pad;
    long long first[2];  // [cur, next] first surviving frame index (LLONG_MAX: none)
};

#define REG_INF 0x7fffffffffffffffll

__global__ void reg_init_alive_kernel(const float* __restrict__ mind, int64_t n, float dmin, int has_centers,
                                      unsigned char* __restrict__ alive, RegState* st) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float m = has_centers ? mind[i] : 3.402823466e+38f;
    const bool a = m > dmin;
    alive[i] = a ? 1 : 0;
    if (a) atomicMin((unsigned long long*)&st->first[0], (unsigned long long)i);
}

// if there is a first survivor: append it as a center (or flag max-centers) and publish its row
__global__ void reg_append_kernel(const float* __restrict__ X, int d, RegState* st, float* __restrict__ centers,
                                  float* __restrict__ cur_row, long long* __restrict__ frame_idx,
                                  long long chunk_offset) {
    __shared__ long long f;
    __shared__ int go;
    if (threadIdx.x == 0) {
        go = 0;
        f = st->first[0];
        if (st->status == 0) {
            if (f == REG_INF) st->status = 1;
            else if (st->n_centers + 1 > st->max_centers) st->status = 4;
            else go = 1;
        }
    }
    __syncthreads();
    if (!go) return;
    const long long nc = st->n_centers;
    for (int e = threadIdx.x; e < d; e += blockDim.x) {
        const float v = X[f * d + e];
        centers[nc * d + e] = v;
        cur_row[e] = v;
    }
    if (threadIdx.x == 0) {
        if (frame_idx) frame_idx[nc] = chunk_offset + f;
        st->n_centers = nc + 1;
        st->first[1] = REG_INF;
    }
}

// survivors after `cur` that are within dmin of the new center die; the rest bid for "next first"
__global__ void reg_kill_kernel(const float* __restrict__ dist, int64_t n, float dmin,
                                unsigned char* __restrict__ alive, RegState* st) {
    if (st->status != 0) return;
    const long long cur = st->first[0];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || i <= cur || !alive[i]) return;
    if (dist[i] <= dmin) alive[i] = 0;  // NaN distances never kill (the reference's `dj < mind` is false)
    else atomicMin((unsigned long long*)&st->first[1], (unsigned long long)i);
}

__global__ void reg_advance_kernel(RegState* st) {
    if (st->status != 0) return;
    st->first[0] = st->first[1];
}

}  // namespace b2k

using namespace b2k;

struct b2k_regspace {
    b2k_ctx* ctx = nullptr;
    int d = 0, metric = 0;
    float dmin = 0.f;
    int64_t max_centers = 0;
    int64_t n_centers = 0;  // host mirror
    bool full = false;
    int64_t frames_seen = 0;
    float* centers = nullptr;    // [max_centers][d]
    float* centers_c = nullptr;  // minRMSD: centered copies
    float* Gb = nullptr;         // minRMSD: traces of centered centers
    float* cur_row = nullptr;    // [d] the center of the current step (+ centered copy, trace)
    float* cur_row_c = nullptr;
    float* cur_g = nullptr;
    RegState* st = nullptr;
    long long* frame_idx = nullptr;
    // per-chunk scratch
    float* mind = nullptr;
    float* Ga = nullptr;
    unsigned char* alive = nullptr;
    int32_t* labels = nullptr;
    int64_t cap = 0;
};

static int reg_ensure_chunk(b2k_regspace* r, int64_t n) {
    if (n <= r->cap) return B2K_OK;
    dev_free(r->mind); dev_free(r->Ga); dev_free(r->alive); dev_free(r->labels);
    r->mind = r->Ga = nullptr; r->alive = nullptr; r->labels = nullptr; r->cap = 0;
    CUDA_TRY(dev_alloc(&r->mind, n * 4));
    CUDA_TRY(dev_alloc(&r->Ga, n * 4));
    CUDA_TRY(dev_alloc(&r->alive, n));
    CUDA_TRY(dev_alloc(&r->labels, n * 4));
    r->cap = n;
    return B2K_OK;
}

B2K_API int b2k_regspace_create(b2k_ctx* ctx, int32_t d, float dmin, int64_t max_centers, int metric,
                                b2k_regspace** out) {
    if (!ctx || !out || d < 1 || max_centers < 0 || !(dmin >= 0.f))
        return set_error(B2K_ERR_INVALID_ARG, "regspace_create: bad arguments");
    if (metric == B2K_METRIC_MINRMSD && d % 3)
        return set_error(B2K_ERR_DIM_NOT_MULT3, "RMSDMetric is only implemented for input data with a dimension divisible by 3.");
    b2k_regspace* r = new b2k_regspace();
    r->ctx = ctx; r->d = d; r->metric = metric; r->dmin = dmin; r->max_centers = max_centers;
    const size_t cbytes = (size_t)std::max<int64_t>(max_centers, 1) * d * 4;
    cudaError_t e = dev_alloc(&r->centers, cbytes);
    if (e == cudaSuccess) e = dev_alloc(&r->cur_row, (size_t)d * 4);
    if (e == cudaSuccess) e = dev_alloc(&r->st, sizeof(RegState));
    if (e == cudaSuccess) e = dev_alloc(&r->frame_idx, (size_t)std::max<int64_t>(max_centers, 1) * 8);
    if (e == cudaSuccess && metric == B2K_METRIC_MINRMSD) {
        e = dev_alloc(&r->centers_c, cbytes);
        if (e == cudaSuccess) e = dev_alloc(&r->Gb, (size_t)std::max<int64_t>(max_centers, 1) * 4);
        if (e == cudaSuccess) e = dev_alloc(&r->cur_row_c, (size_t)d * 4);
        if (e == cudaSuccess) e = dev_alloc(&r->cur_g, 4);
    }
    if (e != cudaSuccess) {
        b2k_regspace_destroy(r);
        return set_error(B2K_ERR_NOMEM, "regspace_create: %s", cudaGetErrorString(e));
    }
    RegState hs;
    hs.n_centers = 0; hs.max_centers = max_centers; hs.status = 0; hs.pad = 0; hs.first[0] = hs.first[1] = REG_INF;
    CUDA_TRY(cudaMemcpy(r->st, &hs, sizeof(hs), cudaMemcpyHostToDevice));
    *out = r;
    return B2K_OK;
}

B2K_API int b2k_regspace_destroy(b2k_regspace* r) {
    if (!r) return B2K_OK;
    dev_free(r->centers); dev_free(r->centers_c); dev_free(r->Gb); dev_free(r->cur_row); dev_free(r->cur_row_c);
    dev_free(r->cur_g); dev_free(r->st); dev_free(r->frame_idx); dev_free(r->mind); dev_free(r->Ga);
    dev_free(r->alive); dev_free(r->labels);
    delete r;
    return B2K_OK;
}

B2K_API int64_t b2k_regspace_n_centers(const b2k_regspace* r) { return r ? r->n_centers : 0; }

B2K_API int b2k_regspace_get_centers(b2k_regspace* r, float* centers_out) {
    if (!r) return set_error(B2K_ERR_INVALID_ARG, "null handle");
    CUDA_TRY(cudaStreamSynchronize(r->ctx->stream));
    CUDA_TRY(cudaMemcpy(centers_out, r->centers, (size_t)r->n_centers * r->d * 4, cudaMemcpyDeviceToHost));
    return B2K_OK;
}

// one sub-chunk of frames (the algorithm is streaming: feeding a chunk in pieces gives the same centers)
static int reg_fit_piece(b2k_regspace* r, const float* dX, int64_t n) {
    b2k_ctx* ctx = r->ctx;
    cudaStream_t st = ctx->stream;
    const int d = r->d;
    B2K_TRY(reg_ensure_chunk(r, n));
    const bool rmsd = r->metric == B2K_METRIC_MINRMSD;
    if (rmsd) B2K_TRY(launch_rmsd_center(ctx, dX, n, d, nullptr, r->Ga));

    // pass 1: distance to the centers known so far
    if (r->n_centers > 0) {
        if (rmsd)
            B2K_TRY(launch_rmsd_assign(ctx, dX, r->Ga, n, d, r->centers_c, r->Gb, (int)r->n_centers, r->labels,
                                       r->mind, 0));
        else
            B2K_TRY(launch_assign_exact(ctx, dX, n, d, r->centers, (int)r->n_centers, r->labels, r->mind, 0));
    }
    RegState hs;
    hs.n_centers = r->n_centers; hs.max_centers = r->max_centers; hs.status = 0; hs.pad = 0;
    hs.first[0] = hs.first[1] = REG_INF;
    CUDA_TRY(cudaMemcpyAsync(r->st, &hs, sizeof(hs), cudaMemcpyHostToDevice, st));
    reg_init_alive_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(r->mind, n, r->dmin, r->n_centers > 0 ? 1 : 0,
                                                                  r->alive, r->st);
    LAUNCH_CHECK();

    const int STEP_BATCH = 16;
    for (;;) {
        for (int s = 0; s < STEP_BATCH; ++s) {
            reg_append_kernel<<<1, 256, 0, st>>>(dX, d, r->st, r->centers, r->cur_row, r->frame_idx, r->frames_seen);
            LAUNCH_CHECK();
            // distances of all chunk frames to the step's center.  (Launched unconditionally: when the
            // step is a no-op the kill kernel ignores the result.)
            if (rmsd) {
                B2K_TRY(launch_rmsd_center(ctx, r->cur_row, 1, d, r->cur_row_c, r->cur_g));
                B2K_TRY(launch_rmsd_dist_rows(ctx, dX, r->Ga, n, d, r->cur_row_c, r->cur_g, 1, r->mind));
            } else {
                B2K_TRY(launch_dist_rows(ctx, dX, n, d, r->cur_row, 1, r->mind));
            }
            reg_kill_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(r->mind, n, r->dmin, r->alive, r->st);
            LAUNCH_CHECK();
            reg_advance_kernel<<<1, 1, 0, st>>>(r->st);
            LAUNCH_CHECK();
        }
        CUDA_TRY(cudaMemcpyAsync(&hs, r->st, sizeof(hs), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (hs.status != 0) break;
    }
    const int64_t added = hs.n_centers - r->n_centers;
    if (rmsd && added > 0)
        B2K_TRY(launch_rmsd_center(ctx, r->centers + r->n_centers * d, added, d, r->centers_c + r->n_centers * d,
                                   r->Gb + r->n_centers));
    r->n_centers = hs.n_centers;
    r->frames_seen += n;
    if (hs.status == 4) {
        r->full = true;
        return set_error(B2K_ERR_MAX_CENTERS, "Maximum number of cluster centers reached (%lld).", (long long)r->max_centers);
    }
    return B2K_OK;
}

B2K_API int b2k_dev_regspace_partial_fit(b2k_regspace* r, const float* dX, int64_t n) {
    if (!r) return set_error(B2K_ERR_INVALID_ARG, "null handle");
    if (r->full) return set_error(B2K_ERR_MAX_CENTERS, "Maximum number of cluster centers reached (%lld).", (long long)r->max_centers);
    if (n <= 0) return B2K_OK;
    CUDA_TRY(cudaSetDevice(r->ctx->device));
    // Every ordered step evaluates the distance of ALL frames of the piece to the new center, so the piece is kept
    // to ~64 MB of frames: a step then costs a few launches, and a max_centers stop does not pay for the frames
    // behind it.  Frames that die against the known centers (pass 1) are the bulk of the work either way.
    const int64_t piece = std::max<int64_t>(16384, std::min<int64_t>(int64_t(1) << 22, (int64_t(64) << 20) / ((int64_t)r->d * 4)));
    for (int64_t off = 0; off < n; off += piece) {
        const int rc = reg_fit_piece(r, dX + off * r->d, std::min(piece, n - off));
        if (rc != B2K_OK) return rc;
    }
    return B2K_OK;
}

// ---- host-pointer wrappers ------------------------------------------------------------------------
B2K_API int b2k_regspace_partial_fit(b2k_regspace* r, const float* X, int64_t n) {
    if (!r) return set_error(B2K_ERR_INVALID_ARG, "null handle");
    if (n <= 0) return B2K_OK;
    if (!X) return set_error(B2K_ERR_INVALID_ARG, "null frames");
    CUDA_TRY(cudaSetDevice(r->ctx->device));
    float* dX = nullptr;
    const size_t bytes = (size_t)n * r->d * 4;
    if (dev_alloc(&dX, bytes) != cudaSuccess) {
        cudaGetLastError();
        return set_error(B2K_ERR_NOMEM, "dev_alloc(%zu bytes) failed", bytes);
    }
    int rc = upload_host(r->ctx, X, dX, bytes);
    if (rc == B2K_OK) rc = b2k_dev_regspace_partial_fit(r, dX, n);
    cudaStreamSynchronize(r->ctx->stream);
    dev_free(dX);
    return rc;
}

B2K_API int b2k_regspace_cluster(b2k_ctx* ctx, const float* X, int64_t n, int32_t d, float* centers_io,
                                 int64_t* n_centers_io, float dmin, int64_t max_centers, int metric) {
    if (!ctx || !centers_io || !n_centers_io) return set_error(B2K_ERR_INVALID_ARG, "regspace_cluster: null argument");
    if (*n_centers_io != 0)
        return set_error(B2K_ERR_INVALID_ARG, "regspace_cluster: resume with existing centers goes through the handle API");
    b2k_regspace* r = nullptr;
    B2K_TRY(b2k_regspace_create(ctx, d, dmin, max_centers, metric, &r));
    int rc = b2k_regspace_partial_fit(r, X, n);
    if (rc == B2K_OK || rc == B2K_ERR_MAX_CENTERS) {
        *n_centers_io = b2k_regspace_n_centers(r);
        const int rc2 = b2k_regspace_get_centers(r, centers_io);
        if (rc2 != B2K_OK) rc = rc2;
    }
    b2k_regspace_destroy(r);
    return rc;
}
