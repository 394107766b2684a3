/* b2k.h -- C ABI of libb2k.so: the B200-native replacement for the numeric backend behind
 * PyEMMA's k-means / assign / regspace / minRMSD hot path.
 *
 * What it replaces (reference @ 3327f28, paths relative to /root/reference):
 *   the per-metric pybind11 module that
 *     pyemma/coordinates/clustering/src/clustering_module.cpp:38-43
 *   builds with deeptime::clustering::registerClusteringImplementation<Metric>(m)
 *   (functions: assign, kmeans.cluster, kmeans.cluster_loop, kmeans.cost_function,
 *   kmeans.init_centers_kmpp, regspace.cluster, compute_metric) and that is called from
 *     clustering/kmeans.py:254-258      (KMeans.fit -> init_centers_kmpp + cluster_loop)
 *     clustering/interface.py:164-165   (ClusterModel.transform -> assign)
 *     clustering/regspace.py:150        (RegularSpace.partial_fit -> regspace.cluster)
 *     clustering/tests/test_kmeans.py:250,304 (compute_metric, init_centers_kmpp)
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / numpy / C++ types cross this boundary.
 *   - frames X: float32, C-contiguous (n, d); centers: float32 (k, d); labels: int32.
 *   - the caller allocates every input and output buffer; the library never frees caller
 *     memory.  Opaque handles own streams and scratch; explicit create / destroy.
 *   - every function returns an int status (B2K_OK == 0); b2k_last_error() gives the text.
 *   - b2k_*      : HOST pointers (pageable or pinned); the library stages chunks through its
 *                  own pinned buffers onto CUDA streams (the reference's chunk hand-off,
 *                  coordinates/data/_base/datasource.py:405-410, kmeans.py:326-338).
 *   - b2k_dev_*  : DEVICE pointers (cudaMalloc / torch.Tensor.data_ptr()); work is enqueued
 *                  on the context stream; results are complete after b2k_ctx_sync() unless a
 *                  host out-parameter is written (then the call synchronises itself).
 *   - there is NO CPU fallback: every entry point fails with B2K_ERR_CUDA when no sm_100
 *     device is usable.
 */
#ifndef B2K_H
#define B2K_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes -> Python exceptions raised by the reference for the same condition */
#define B2K_OK 0
#define B2K_ERR_INVALID_ARG 2   /* std::invalid_argument -> ValueError (dim mismatch, k>n; tests/test_assign.py:183-197) */
#define B2K_ERR_DIM_NOT_MULT3 3 /* std::range_error (clustering_module.cpp:12-14) */
#define B2K_ERR_MAX_CENTERS 4   /* MaxCentersReachedException (regspace.py:153-163); centers so far are valid */
#define B2K_ERR_CUDA 5          /* CUDA runtime/driver error or no usable device */
#define B2K_ERR_NOMEM 6         /* device or pinned allocation failed (kmeans.py:187-192 MemoryError) */
#define B2K_ERR_NONFINITE 7     /* NaN/inf in input (InvalidDataInStreamException, datasource.py:1067-1075) */

#define B2K_METRIC_EUCLIDEAN 0
#define B2K_METRIC_MINRMSD 1

/* ordered-sum mode of k-means++ (DESIGN.md "k-means++"): */
#define B2K_KMPP_SERIAL 0  /* fp32 sums in frame order: bit-faithful to the reference at n_jobs=1 */
#define B2K_KMPP_BLOCKED 1 /* balanced-tree sums + tree descent: parallel, deterministic */

/* assignment engine selection (b2k_ctx_set_option "assign_engine") */
#define B2K_ENGINE_AUTO 0
#define B2K_ENGINE_DIRECT 1 /* exact fp32 CUDA-core kernel only */
#define B2K_ENGINE_SCREEN 2 /* tcgen05 screen + exact verify (euclidean only) */

typedef struct b2k_ctx b2k_ctx;
typedef struct b2k_lloyd b2k_lloyd;
typedef struct b2k_regspace b2k_regspace;
typedef void (*b2k_callback)(void* user);
/* all-reduce callback of the sharded k-means++: reduce the first `count` elements of exchange buffer `buffer_id`
 * (0: the float buffer, 1: the int64 buffer) over all ranks, op 0 = sum(f32), 1 = max(i64), 2 = min(i64); the
 * library has synchronised its stream before the call and reads the buffer on its stream afterwards; return 0 */
typedef int (*b2k_exchange_fn)(void* user, int buffer_id, int64_t count, int op);

const char* b2k_last_error(void);
int b2k_version(void);
/* number of kernels launched by this library in this process so far (bench.py "gpu_launches") */
int64_t b2k_launch_count(void);

/* ---- context ---------------------------------------------------------------------------- */
int b2k_ctx_create(int device, b2k_ctx** out);
int b2k_ctx_destroy(b2k_ctx* ctx);
/* run on an external cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream).  The handle is used
 * as is: NULL is CUDA's legacy default stream.  b2k_ctx_set_option(ctx, "own_stream", 1) goes back to a
 * private non-blocking stream (the state after b2k_ctx_create). */
int b2k_ctx_set_stream(b2k_ctx* ctx, void* cuda_stream);
int b2k_ctx_sync(b2k_ctx* ctx);
/* options: "assign_engine" (B2K_ENGINE_*), "screen_terms" (fp16 operand terms of the tensor-core screen: 3 = hi+lo split of frames and
 * centers, 2 = split frames x hi-only centers, 1 = hi-only; 0 (default) = measured on a sample per data set for wide rows),
 * "screen_group" (centers per candidate group the screen hands to the exact verify: 0 automatic, 8, 4 or 2),
 * "stage_bytes" (pinned staging buffer size per slot), "check_finite" (1: the host-pointer
 * assign / stage entry points check every staged chunk on the device and return B2K_ERR_NONFINITE for NaN/inf frames,
 * the guard of datasource.py:1067-1075; default 1), "host_copy_threads" (threads of the
 * pageable -> pinned bounce copy, default 8), "accumulate_mode" (member sums: 0 automatic, 1 one 64-bit
 * RED per frame element, 2 segmented = counting sort by label + warp run sums, 3 per-CTA shared-memory table,
 * 4 tile-sorted = per-tile shared-memory sort + run sums, narrow rows), "cost_kernel" (Lloyd cost pass: 0 automatic --
 * one-pass kernel for d <= 16, 4-lanes-per-frame kernel above; 1 shared-memory staged wide-row kernel; 2 always two passes),
 * "row_vec_max" (widest frame-row load of the narrow-row verify / cost kernels: 4, 2 or 1 floats; default 4),
 * "delta_sums" (Lloyd sessions with center pruning: 1 = the exact integer member sums are UPDATED from the frames whose
 * label changed once at most an eighth of them did in the previous iteration -- same integers as a full pass; 0 = always a
 * full pass; 2 = always incremental; default 1),
 * "prune_list_margin" (per mille of the mean tile radius, default 50: the per-tile center lists of a pruned session are
 * built with that much room for center movement and reused until some center has moved farther from where it was at the
 * build -- measured by one small kernel per iteration; 0 = lists rebuilt every iteration; exact either way),
 * "cache_mb" (bound, in MiB, of the device block cache: working buffers the library frees -- the fp16 screen operand,
 * the sorted frame copy of a Lloyd session -- stay mapped for the next call of a similar size, because cudaFree /
 * cudaMalloc of GB-sized blocks cost ~0.1 s per GB; -1 = default, half of the device memory; 0 = off),
 * "cache_release" (any value: every cached block goes back to the driver now; the library does this by itself when one
 * of its own allocations fails and in b2k_ctx_destroy),
 * "own_stream", "profile" (1: time every launch of the
 * tcgen05 screen kernel with CUDA events on the context stream; setting it again clears the record).
 * Experimental operand modes of the streaming screen kernel (results identical): "screen_resident_a"
 * (1: frame tile resident in shared memory), "screen_cluster" (2: 2-CTA clusters with TMA-multicast center tiles,
 * 3: CTA-pair MMAs, tcgen05 cta_group::2) */
int b2k_ctx_set_option(b2k_ctx* ctx, const char* name, int64_t value);
/* stats: "screen_frames", "screen_cand_chunks" (8-center groups re-evaluated exactly), "screen_fallback_frames"
 * of the last screened assign (host-pointer calls: of its last chunk); "screen_gemm_ms_total" and
 * "screen_gemm_launches" since "profile" was set; "sm_count"; "cache_bytes" (device block cache);
 * "labels_changed" (frames whose label changed in the last counted pruned step), "delta_steps" (incremental-sum steps so far),
 * "list_reuse_steps" (pruned steps that reused the center lists of an earlier step); "fp32_lane_instr_per_s" (measures the fp32 CUDA-core
 * issue rate of non-fusable FMUL/FADD chains on the spot: the denominator quoted for the CUDA-core bound kernels) */
int b2k_ctx_get_stat(b2k_ctx* ctx, const char* name, double* value);

/* the chunk hand-off on its own: host bytes -> device array through the pinned staging slots (pageable sources are
 * bounced with several copy threads); synchronous */
int b2k_upload(b2k_ctx* ctx, const void* src_host, void* dst_dev, int64_t bytes);

/* ---- compute_metric  (clustering_module.cpp:41-43) -------------------------------------- */
int b2k_compute_metric(b2k_ctx* ctx, const float* x, const float* y, int64_t d, int metric, float* out);

/* ---- assign  (deeptime assign_chunk_to_centers; interface.py:164-165) ------------------- */
/* labels[i] = first argmin_j sqrt(dist2(x_i, c_j)); -1 when no distance is < FLT_MAX */
int b2k_assign(b2k_ctx* ctx, const float* X, int64_t n, int32_t d, const float* centers, int32_t k, int metric,
               int32_t* labels);
int b2k_dev_assign(b2k_ctx* ctx, const float* dX, int64_t n, int32_t d, const float* dcenters, int32_t k,
                   int metric, int32_t* dlabels, float* dmindist_or_null);

/* the gather pass of KmeansClustering._estimate (kmeans.py:220-230, 326-338) fused with the first assignment:
 * host frames are staged chunk by chunk through pinned memory into the caller's DEVICE array dX_out (n*d floats)
 * and every chunk is assigned against dcenters while the next one is on the bus; labels go to dlabels_out (device,
 * n) and, if labels_host_or_null is given, to the host as well.  lloyd != 0: kmeans.cluster label semantics. */
int b2k_stage_assign(b2k_ctx* ctx, const float* X, int64_t n, int32_t d, const float* dcenters, int32_t k, int metric,
                     int lloyd, float* dX_out, int32_t* dlabels_out, int32_t* labels_host_or_null);

/* ---- k-means  (deeptime kmeans.cluster / cost_function / cluster_loop; kmeans.py:254-258) */
/* one Lloyd step: labels from `centers`, new centers = member mean (empty cluster keeps the old) */
int b2k_kmeans_cluster(b2k_ctx* ctx, const float* X, int64_t n, int32_t d, const float* centers, int32_t k,
                       int metric, float* new_centers, int32_t* labels);
int b2k_kmeans_cost(b2k_ctx* ctx, const float* X, int64_t n, int32_t d, const float* centers, int32_t k,
                    const int32_t* labels, int metric, float* cost);
/* do { step; cost; rel=|cost-prev|/cost; converged if rel<=tol else callback } while (it<max_iter && !conv)
 * code: 0 converged, 1 not; inertias[0..iters) */
int b2k_kmeans_cluster_loop(b2k_ctx* ctx, const float* X, int64_t n, int32_t d, float* centers_io, int32_t k,
                            int metric, int32_t max_iter, float tolerance, b2k_callback cb, void* user, int* code,
                            int* iters, float* inertias, int32_t inertias_cap);
/* k-means++ (test_kmeans.py:304: init_centers_kmpp(data,k,random_seed,n_threads,callback)); seed<0 = entropy */
int b2k_kmeans_init_centers_kmpp(b2k_ctx* ctx, const float* X, int64_t n, int32_t d, int32_t k, int metric,
                                 int64_t seed, int scan_mode, b2k_callback cb, void* user, float* centers_out,
                                 int64_t* chosen_or_null);

/* device-resident Lloyd session: the frames of ONE shard stay in HBM across iterations.
 * Multi-GPU: every rank owns a session over its shard; `acc` is the exchange buffer that the
 * caller all-reduces (sum, int64) between accumulate and finalize -- layout
 *   acc[0 .. k*d)   fixed-point coordinate sums (value * 2^qbits, exact integer adds)
 *   acc[k*d .. k*d+k) member counts
 *   acc[k*d+k]      fixed-point cost
 * so results are independent of the number of ranks and of any reduction order. */
int b2k_dev_lloyd_create(b2k_ctx* ctx, const float* dX, int64_t n_local, int32_t d, int32_t k, int metric,
                         int64_t n_total, float absmax_global, b2k_lloyd** out);
int b2k_dev_lloyd_destroy(b2k_lloyd* s);
int64_t b2k_dev_lloyd_acc_len(const b2k_lloyd* s);
/* labels (n_local) <- argmin vs dcenters (Lloyd tie/NaN semantics); acc <- local sums+counts (cost slot zeroed).
 * dlabels may be NULL: the loop itself never needs the labels in the caller's frame order (deeptime's cluster_loop returns
 * centers only, kmeans.py:254-258), the session then keeps them (after its first iteration it works on a copy of the frames
 * sorted by label) and b2k_dev_lloyd_get_labels hands them out on request. */
int b2k_dev_lloyd_assign_accumulate(b2k_lloyd* s, const float* dcenters, int32_t* dlabels_or_null, int64_t* dacc);
int b2k_dev_lloyd_get_labels(b2k_lloyd* s, int32_t* dlabels_out);
/* the two calls above for frames that are still on the HOST: the shard's frames X (n_local x d, host) are staged chunk
 * by chunk into the session's device array dX_out (the pointer given to b2k_dev_lloyd_create), every chunk is assigned
 * and its member sums are added to acc while the next chunk is on the bus; labels also go to labels_host if given */
int b2k_stage_lloyd_assign_accumulate(b2k_lloyd* s, const float* X, const float* dcenters, float* dX_out,
                                      int32_t* dlabels_out, int32_t* labels_host_or_null, int64_t* dacc);
/* Out-of-core tier (the reference spills to a host memmap when the array does not fit, kmeans.py:181-200): create the
 * session with dX = NULL, keep the frames in (pinned) host memory and run ONE pass per iteration.  Per chunk: if
 * have_prev, the cost of the previous labels (dlabels_io on entry) against dcenters -- the cost of the iteration that
 * produced dcenters -- is added to the cost slot; the chunk is assigned (dlabels_io on exit, labels_host if given); its
 * member sums and counts are added to acc.  acc and cost are bit-identical to the resident calls'. */
int b2k_stage_lloyd_pass(b2k_lloyd* s, const float* X, const float* dcenters, int32_t* dlabels_io, int have_prev,
                         int32_t* labels_host_or_null, int64_t* dacc);
/* acc <- local sums+counts for GIVEN labels (e.g. those b2k_stage_assign produced); cost slot zeroed */
int b2k_dev_lloyd_accumulate(b2k_lloyd* s, const int32_t* dlabels, int64_t* dacc);
/* new centers from (all-reduced) acc; count==0 keeps old */
int b2k_dev_lloyd_finalize(b2k_lloyd* s, const int64_t* dacc, const float* dcenters_old, float* dcenters_new);
/* acc[k*d+k] <- local fixed-point sum of compute(x_i, new_centers[label_i])^2, label_i = the labels of the session's last
 * b2k_dev_lloyd_assign_accumulate (dlabels: that call's label array, or NULL if it was called with NULL) */
int b2k_dev_lloyd_cost(b2k_lloyd* s, const float* dcenters_new, const int32_t* dlabels_or_null, int64_t* dacc);
/* host-side decode of the (all-reduced) cost slot */
double b2k_dev_lloyd_decode_cost(const b2k_lloyd* s, int64_t cost_fixed);
/* max |x| over a device array (for absmax_global; all-reduce(max) it across ranks) */
int b2k_dev_absmax(b2k_ctx* ctx, const float* dX, int64_t count, float* out_host);
/* 1 if every value is finite */
int b2k_dev_all_finite(b2k_ctx* ctx, const float* dX, int64_t count, int* out_host);

/* cluster_loop over device-resident frames of ONE GPU (no exchange); dlabels_or_null receives the labels of the last step */
int b2k_dev_kmeans_cluster_loop(b2k_ctx* ctx, const float* dX, int64_t n, int32_t d, float* dcenters_io, int32_t k,
                                int metric, int32_t max_iter, float tolerance, b2k_callback cb, void* user, int* code,
                                int* iters, float* inertias, int32_t inertias_cap, int32_t* dlabels_or_null);

int b2k_dev_kmeans_init_centers_kmpp(b2k_ctx* ctx, const float* dX, int64_t n, int32_t d, int32_t k, int metric,
                                     int64_t seed, int scan_mode, b2k_callback cb, void* user, float* dcenters_out,
                                     int64_t* chosen_host_or_null);

/* k-means++ (blocked scan mode) over frames SHARDED across ranks: every rank passes its shard (global index of its
 * first frame `global_lo`, a multiple of 1024; n_local may be 0), two caller-owned DEVICE exchange buffers
 * (xchg_f32 with at least b2k_kmpp_exchange_floats(n_total, d, k) floats, xchg_i64 with 32 int64) and the
 * all-reduce callback.  Every rank receives the same k centers; picks are bit-identical to the single-GPU call. */
int64_t b2k_kmpp_exchange_floats(int64_t n_total, int32_t d, int32_t k);
int b2k_dev_kmeans_init_centers_kmpp_sharded(b2k_ctx* ctx, const float* dX, int64_t n_local, int32_t d, int32_t k,
                                             int metric, int64_t seed, int64_t global_lo, int64_t n_total,
                                             float* xchg_f32, int64_t xchg_f32_len, int64_t* xchg_i64,
                                             b2k_exchange_fn exchange, void* exchange_user, b2k_callback cb,
                                             void* user, float* dcenters_out, int64_t* chosen_host_or_null);

/* ---- regspace  (deeptime regspace.cluster; regspace.py:144-151) -------------------------- */
int b2k_regspace_create(b2k_ctx* ctx, int32_t d, float dmin, int64_t max_centers, int metric, b2k_regspace** out);
int b2k_regspace_destroy(b2k_regspace* r);
/* feed the next chunk (frames in order).  B2K_ERR_MAX_CENTERS when a frame would be center
 * max_centers+1: the handle then holds exactly max_centers centers and ignores further chunks. */
int b2k_regspace_partial_fit(b2k_regspace* r, const float* X, int64_t n);
int b2k_dev_regspace_partial_fit(b2k_regspace* r, const float* dX, int64_t n);
int64_t b2k_regspace_n_centers(const b2k_regspace* r);
int b2k_regspace_get_centers(b2k_regspace* r, float* centers_out /* n_centers*d */);
/* one-shot convenience with the reference's signature shape */
int b2k_regspace_cluster(b2k_ctx* ctx, const float* X, int64_t n, int32_t d, float* centers_io,
                         int64_t* n_centers_io, float dmin, int64_t max_centers, int metric);

/* ---- dtraj consumers: pyemma/util/discrete_trajectories.py:146-181 (count_states) and the lagged
 * transition count matrix the MSM estimators take from the dtrajs (msm/estimators/_msm_estimator_base.py:3,229) --- */
/* dcounts[s] += #{t : labels[t] == s}; negative labels are skipped, a label >= nstates is B2K_ERR_INVALID_ARG */
int b2k_dev_count_states(b2k_ctx* ctx, const int32_t* dlabels, int64_t n, int32_t nstates, int64_t* dcounts);
/* dC[i*nstates+j] += #{t : labels[t] == i, labels[t+lag] == j} over ONE trajectory of n frames; sliding != 0: every t,
 * else t = 0, lag, 2 lag, ...; pairs with a negative label are skipped.  The caller zeroes dC and calls once per
 * trajectory (pairs never span trajectories). */
int b2k_dev_count_matrix(b2k_ctx* ctx, const int32_t* dlabels, int64_t n, int32_t nstates, int64_t lag, int sliding,
                         int64_t* dC);

/* ---- the step before the path: linear projection of TICA / PCA fused into the chunk hand-off,
 * Y = (X - mean) . W[:, :dout]  (pyemma/coordinates/transform/_tica_base.py:130-133, pca.py:263-265); mean (din) and
 * W (din x ldw, row-major, ldw >= dout) are fp64 DEVICE arrays like the reference's model, Y is fp32 (n x dout) */
int b2k_dev_project(b2k_ctx* ctx, const float* dX, int64_t n, int32_t din, const double* dmean_or_null,
                    const double* dW, int32_t ldw, int32_t dout, float* dY);
/* same with HOST frames X (n x din): staged chunk by chunk, only the projected rows stay in HBM (dY, device) */
int b2k_stage_project(b2k_ctx* ctx, const float* X, int64_t n, int32_t din, const double* dmean_or_null,
                      const double* dW, int32_t ldw, int32_t dout, float* dY);

#ifdef __cplusplus
}
#endif
#endif /* B2K_H */
