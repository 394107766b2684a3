// kernels.h -- internal launcher prototypes shared by the libb2k translation units.
#pragma once
#include "common.cuh"

namespace b2k {

enum { MODE_ARGMIN = 0, MODE_ALL = 1 };

// ---- exact.cu (Euclidean, exact fp32 reference order) -------------------------------------
int launch_assign_exact(b2k_ctx* ctx, const float* X, int64_t n, int d, const float* C, int k, int32_t* labels,
                        float* mind, int lloyd);
// same, but the kernel returns at once unless *run_if_zero == 0 (device flag; null = always run)
int launch_assign_exact_if(b2k_ctx* ctx, const float* X, int64_t n, int d, const float* C, int k, int32_t* labels,
                           float* mind, int lloyd, const int* run_if_zero);
int launch_tile(b2k_ctx* ctx, const float* X, int64_t n, int d, const float* C, int k, int32_t* labels, float* out,
                int lloyd, int mode);
// exact argmin for the frames row_index[0 .. *count_dev) (device-side count; skipped when *run_if_nonzero == 0)
int launch_tile_indexed(b2k_ctx* ctx, const float* X, int d, const float* C, int k, const uint32_t* row_index,
                        const unsigned int* count_dev, const int* run_if_nonzero, int32_t* labels, float* mind,
                        int lloyd);
// out[j][i] = sqrt(dist2(x_i, rows_j)), j < m
int launch_dist_rows(b2k_ctx* ctx, const float* X, int64_t n, int d, const float* rows, int m, float* out);
// same with the k-means++ triangle-inequality pruning (exact.cu DistRowsPrune); D == null: no pruning
int launch_dist_rows_pruned(b2k_ctx* ctx, const float* X, int64_t n, int d, const float* rows, int m, float* out,
                            const float* D, const int32_t* assigned, const unsigned char* taken, const float* Rc,
                            int rc_stride, uint32_t* list, uint32_t* masks, unsigned int* count, uint16_t* framemask);
int launch_labeled_dist(b2k_ctx* ctx, const float* X, int64_t n, int d, const float* C, const int32_t* labels,
                        float* out);

// ---- rmsd.cu (minRMSD / QCP) ----------------------------------------------------------------
// centered copies + traces of m structures (rows of `src`, optionally gathered by idx)
int launch_rmsd_center(b2k_ctx* ctx, const float* src, int64_t m, int d, float* centered_or_null, float* traces);
int launch_rmsd_assign(b2k_ctx* ctx, const float* X, const float* Ga, int64_t n, int d, const float* Cc,
                       const float* Gb, int k, int32_t* labels, float* mind, int lloyd);
int launch_rmsd_dist_rows(b2k_ctx* ctx, const float* X, const float* Ga, int64_t n, int d, const float* Rc,
                          const float* Gb, int m, float* out);
int launch_rmsd_labeled_dist(b2k_ctx* ctx, const float* X, const float* Ga, int64_t n, int d, const float* Cc,
                             const float* Gb, const int32_t* labels, float* out);

// ---- metric-generic front (api.cu) ----------------------------------------------------------
struct MetricData {  // per-dataset auxiliary data of a metric (minRMSD: traces of the centered frames)
    int metric = 0;
    float* Ga = nullptr;  // device, n floats (minRMSD only)
};

// ---- api.cu ----------------------------------------------------------------------------------
// host array -> device (pageable sources bounce through the pinned staging slots)
int upload_host(b2k_ctx* ctx, const void* src, void* dst, size_t bytes);

// ---- lloyd.cu -------------------------------------------------------------------------------
int launch_accumulate(b2k_ctx* ctx, const float* X, int64_t n, int d, int k, const int32_t* labels, double scale,
                      int64_t* acc);
int launch_finalize(b2k_ctx* ctx, const int64_t* acc, int k, int d, double inv_scale, const float* old_centers,
                    float* new_centers);
int launch_cost_reduce(b2k_ctx* ctx, const float* l, int64_t n, double scale, int64_t* acc_slot);
int launch_cost_fused(b2k_ctx* ctx, const float* X, int64_t n, int d, const float* C, int k, const int32_t* labels,
                      double scale, int64_t* acc_slot, int* done);
int measure_fp32_rate(b2k_ctx* ctx, double* lane_instr_per_s);
int launch_absmax(b2k_ctx* ctx, const float* X, int64_t count, float* d_out /* device, 1 float, pre-zeroed */);
int launch_all_finite(b2k_ctx* ctx, const float* X, int64_t count, int* d_flag /* device, pre-set to 1 */);

// ---- screen.cu (tcgen05 distance screen + exact verify) -------------------------------------
struct ScreenPlan;
int screen_plan_create(b2k_ctx* ctx, int64_t n_cap, int d, int k, ScreenPlan** out);
void screen_plan_destroy(ScreenPlan* p);
// plan cached in the context for one-shot / chunked assignment (reused while d, k match and n fits)
int screen_plan_acquire(b2k_ctx* ctx, int64_t n, int d, int k, ScreenPlan** out);
void screen_plan_release_cached(b2k_ctx* ctx);
// (re)build the frame operand for n frames at dX (done once per dataset / chunk)
int screen_prepare_frames(ScreenPlan* p, const float* dX, int64_t n);
// labels for the prepared frames against dcenters
int screen_assign(ScreenPlan* p, const float* dX, int64_t n, const float* dcenters, int32_t* labels, float* mind,
                  int lloyd);
int screen_read_stats(ScreenPlan* p, double* cand_chunks, double* fallback_frames);
bool screen_supported(const b2k_ctx* ctx, int d, int k, int64_t n);

}  // namespace b2k
