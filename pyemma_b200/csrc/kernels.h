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
                        int lloyd, unsigned int min_count = 0 /* the kernel returns at once for shorter lists */);
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
// counting sort of the frame indices by label (labels outside [0, k) are left out; seg[k] = frames sorted)
// incremental member sums (exact integers): acc under old_l -> acc under new_l, reading only the frames whose label changed
int launch_accumulate_delta(b2k_ctx* ctx, const float* X, int64_t n, int d, int k, const int32_t* old_l, const int32_t* new_l,
                            double scale, int64_t* acc, unsigned long long* changed);
int launch_count_changed(b2k_ctx* ctx, const int32_t* old_l, const int32_t* new_l, int64_t n, unsigned long long* changed);
int launch_label_sort(b2k_ctx* ctx, const int32_t* labels, int64_t n, int k, uint32_t* seg /* k+1 */, uint32_t* perm /* n */);
int measure_fp32_rate(b2k_ctx* ctx, double* lane_instr_per_s);
int launch_absmax(b2k_ctx* ctx, const float* X, int64_t count, float* d_out /* device, 1 float, pre-zeroed */);
int launch_all_finite(b2k_ctx* ctx, const float* X, int64_t count, int* d_flag /* device, pre-set to 1 */);

// ---- screen.cu (tcgen05 distance screen + exact verify) -------------------------------------
struct ScreenPlan;
// terms: 1..3, 0 = the context's default (option screen_terms, else 3)
int screen_plan_create(b2k_ctx* ctx, int64_t n_cap, int d, int k, int terms, ScreenPlan** out);
// operand term count for this data set and these centers (measured on a sample when option screen_terms is 0)
int screen_choose_terms(b2k_ctx* ctx, const float* dX, int64_t n, int d, const float* dC, int k, int* terms_out);
int screen_plan_terms(const ScreenPlan* p);
void screen_plan_invalidate_frames(ScreenPlan* p);  // the frame array changed (re-sorted): rebuild the operand at the next assign
void screen_plan_destroy(ScreenPlan* p);
// plan cached in the context for one-shot / chunked assignment (reused while d, k match and n fits)
int screen_plan_acquire(b2k_ctx* ctx, int64_t n, int d, int k, int terms, ScreenPlan** out);
void screen_plan_release_cached(b2k_ctx* ctx);
// (re)build the frame operand for n frames at dX (done once per dataset / chunk)
int screen_prepare_frames(ScreenPlan* p, const float* dX, int64_t n);
// labels for the prepared frames against dcenters
int screen_assign(ScreenPlan* p, const float* dX, int64_t n, const float* dcenters, int32_t* labels, float* mind,
                  int lloyd);
int screen_read_stats(ScreenPlan* p, double* cand_chunks, double* fallback_frames);
bool screen_supported(const b2k_ctx* ctx, int d, int k, int64_t n);

// ---- prune.cu (frames sorted by label, per-tile center lists: exact triangle-inequality pruning) --------------
struct PruneState;
bool prune_supported(const b2k_ctx* ctx, int64_t n, int d, int k);
int prune_create(b2k_ctx* ctx, int64_t n, int d, int k, PruneState** out);
void prune_destroy(PruneState* p);
int prune_sort(PruneState* p, const float* X, const int32_t* labels_current_order, const float* dC);
int prune_lists(PruneState* p, const float* dC, double* mean_count, int* max_count, int* overflow_tiles);
int prune_scatter_labels(PruneState* p, int32_t* out_original_order);
const float* prune_frames(const PruneState* p);   // frames in sorted order
int32_t* prune_labels(PruneState* p);             // labels in sorted order (written by the assign, read by the sort)
const uint16_t* prune_tlist(const PruneState* p);
const uint32_t* prune_tcount(const PruneState* p);
int prune_lcap(const PruneState* p);
int prune_unit_shift(const PruneState* p);  // 1 << shift consecutive 128-frame tiles share one list
bool prune_sorted(const PruneState* p);
// screen + verify over per-tile center lists (frames = the plan's prepared frames, in sorted order)
int screen_assign_listed(ScreenPlan* p, const float* dX, int64_t n, const float* dcenters, const uint16_t* tlist,
                         const uint32_t* tcount, int lcap, int unit_shift, int32_t* labels, int lloyd);

}  // namespace b2k
