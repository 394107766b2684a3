// common.cuh -- shared host/device helpers of libb2k (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <atomic>
#include "../../include/b2k.h"

#define B2K_API extern "C" __attribute__((visibility("default")))

namespace b2k {

// ---- error plumbing ------------------------------------------------------------------------
extern thread_local std::string g_last_error;
extern std::atomic<long long> g_launches;
int set_error(int code, const char* fmt, ...);

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return b2k::set_error(_e == cudaErrorMemoryAllocation ? B2K_ERR_NOMEM : B2K_ERR_CUDA, \
                                  "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,        \
                                  cudaGetErrorString(_e));                                      \
    } while (0)
#define B2K_TRY(expr)                  \
    do {                               \
        int _rc = (expr);              \
        if (_rc != B2K_OK) return _rc; \
    } while (0)
#define LAUNCH_CHECK()                 \
    do {                               \
        b2k::g_launches.fetch_add(1);  \
        CUDA_TRY(cudaGetLastError());  \
    } while (0)

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// cudaMalloc / cudaFree through the library's device block cache (api.cu): same synchronisation as cudaFree, but the
// block stays mapped for the next request of a similar size
cudaError_t dev_alloc_raw(void** p, size_t bytes);
template <class T> static inline cudaError_t dev_alloc(T** p, size_t bytes) { return dev_alloc_raw((void**)p, bytes); }
void dev_free(void* p);
void dev_cache_release(int dev);
size_t dev_cache_bytes(int dev);
void dev_cache_set_limit(int dev, long long bytes);

// owning device allocation (freed with the scope)
struct DevMem {
    void* p = nullptr;
    size_t cap = 0;
    ~DevMem() { if (p) dev_free(p); }
    int alloc(size_t bytes) {
        if (p && bytes <= cap) return B2K_OK;
        if (p) { dev_free(p); p = nullptr; cap = 0; }
        cap = bytes ? bytes : 16;
        if (dev_alloc(&p, cap) != cudaSuccess) {
            p = nullptr;
            cap = 0;
            cudaGetLastError();
            return set_error(B2K_ERR_NOMEM, "cudaMalloc(%zu bytes) failed", bytes);
        }
        return B2K_OK;
    }
    template <class T> T* as() const { return (T*)p; }
};

// cudaFuncSetAttribute is per device: one-time kernel attribute setup is tracked per device ordinal
struct PerDeviceOnce {
    std::atomic<unsigned long long> mask{0};
    bool need(int dev) const { return !((mask.load() >> (dev & 63)) & 1ull); }
    void done(int dev) { mask.fetch_or(1ull << (dev & 63)); }
};

}  // namespace b2k

// ---- context -------------------------------------------------------------------------------
struct b2k_ctx {
    int device = 0;
    int sm_count = 148;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_stream[2] = {nullptr, nullptr};
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    // pinned staging slots for the host-pointer entry points
    void* pinned[2] = {nullptr, nullptr};
    void* pinned_out[2] = {nullptr, nullptr};
    size_t stage_bytes = size_t(64) << 20;
    size_t pinned_cap = 0, pinned_out_cap = 0;
    // optional CUDA-event timing of the screen kernel launches (option "profile", stats "screen_gemm_ms_*")
    int profile = 0;
    std::vector<cudaEvent_t> prof_events;  // start/stop pairs on `stream`
    // the other kernels of a Lloyd step, timed the same way (ProfScope below; stats "prof_ms_<class>", "prof_n_<class>")
    enum { PROF_VERIFY = 0, PROF_SUMS, PROF_COST, PROF_LISTS, PROF_N };
    std::vector<cudaEvent_t> prof_class[PROF_N];
    // options
    int engine = B2K_ENGINE_AUTO;
    int screen_terms = 0;       // 1..3: operand terms of the screen forced, 0: measured per data set (screen_choose_terms)
    int probe_min_gflop = 1000; // ... and only for jobs of at least this many algorithmic GFLOP (2 n k d) per pass
    int probe_max_centers = 12; // the term probe accepts a count that leaves at most this many candidate centers per frame
    double stat_probe_centers[4] = {0, 0, 0, 0}, stat_probe_fallback[4] = {0, 0, 0, 0};  // last probe, per term count
    double stat_screen_terms = 0;  // term count of the last screened call
    int check_finite = 1;       // host-pointer assign entry points reject NaN/inf frames (B2K_ERR_NONFINITE)
    int host_copy_threads = 8;  // threads of the pageable -> pinned bounce copy (1e7 x 10 frames: 32.7 ms with 1, 18.2 ms with 8)
    double stat_kmpp_async_fallbacks = 0;
    int kmpp_async = 2;       // k-means++ (blocked, one GPU, no callback): 1 rounds are queued without a host round trip each,
                              // 2 (default) additionally replayed from one captured CUDA graph, 0 synchronous loop
    int kmpp_prune = 1;       // k-means++ (blocked, euclidean): skip candidate distances the triangle inequality decides
                              // (1: unless the frames fit L2 -- n*d <= 4M floats; 2: always; 0: never)
    int operand_kernel = 0;   // frame operand builder: 0 per-input-element tile kernel, 1 per-output-piece kernel
    int fallback_mode = 0;    // frames the screen cannot bound: 0 by queue length (< 256: CTA-per-frame scan, else the indexed
                              // exact tile kernel), 1 always CTA per frame, 2 always the tile kernel
    int verify_mode = 0;      // wide rows: 0 direct (no staging) verify kernel, 1 shared-memory staged variants
    int screen_cluster = 0;     // 2: streaming-mode screen kernel as 2-CTA clusters sharing the center tiles (TMA multicast)
    int screen_resident_a = 0;  // screen kernel: keep the frame tile in shared memory when the center operand does not fit.
                                // Off: measured at cfg3 it cuts the L2->SM traffic by 27 % but leaves room for only 3 ring
                                // stages of center k-blocks -- 13.0 ms against 11.0 ms for the 4-stage streaming mode
    int prune_mode = 1;       // Lloyd sessions: 1 sort the frames by label after the first iteration and screen every tile
                              // against its own center list (exact), 0 never, 2 also for small jobs, 3 listed screen even
                              // when the lists exclude nothing (tests)
    int screen_decide = 0;    // listed screen: frames with a single possible center (one candidate chunk, its runner-up below the
                              // threshold) skip the exact verify.  Off: 99.4 % of the cfg2 frames and 58 % of the cfg3 frames are
                              // decided, but finding the runner-up costs the instruction-bound epilogue more (screen kernel 0.79 ->
                              // 1.10 ms at cfg2) than the verify saves (0.36 -> ~0.1 ms)
    int screen_gather = 0;    // listed screen: 0 cp.async gather warps, 1 TMA tile::gather4
    int prune_unit_shift = -1; // 1 << shift consecutive 128-frame tiles share one center list (list kernel cost against list
                              // length: measured at 1e7 x 10, k=1000 step 1.93 / 1.87 / 1.88 / 2.01 ms for shift 0..3 when the lists were rebuilt every
                              // iteration; with kept lists 1.40 / 1.45 for shift 0 / 1); -1: 0
    int prune_list_margin = 50;  // per mille of the mean tile radius: the center lists of a pruned session are built with this
                              // much room for center movement and kept until a center has moved farther (0: rebuilt every
                              // iteration)
    double stat_list_reuse = 0;  // pruned steps that reused the lists of an earlier step
    int delta_sums = 1;       // pruned Lloyd sessions: 1 update the exact integer member sums from the frames whose label changed
                              // once at most an eighth of them did in the previous iteration, 0 always a full pass,
                              // 2 always incremental (tests)
    double stat_changed = -1, stat_delta_steps = 0;  // frames whose label changed in the last counted step; incremental steps
    int prune_resort = 0;     // re-sort schedule: 0 at iterations 1, 2, 4, 8, ... ; n > 0 every n iterations
    double stat_prune_mean = 0, stat_prune_steps = 0, stat_prune_sorts = 0;  // mean list length of the last pruned step
    int screen_group = 0;     // centers per candidate group of the screen (0: automatic; 8, 4, 2)
    int rmsd_abandon = 1;     // minRMSD argmin: abandon pairs whose msd lower bound already exceeds the frame's best (exact)
    int rmsd_kernel = 0;      // 0: slab-streaming QCP kernel, 1: whole-row tile kernel
    int row_vec_max = 4;      // widest row load of the narrow-row verify / cost kernels (4, 2 or 1 floats)
    int cost_kernel = 0;      // 0: quad kernel for wide rows / fused one-pass kernel for narrow rows, 1: the shared-memory
                              // staged variant (wide rows), 2: always the two-pass path (per-frame distances, then the sum)
    int accumulate_mode = 0;  // 0: automatic (shared-memory table when it fits, else segmented), 1: one RED per
                              // element, 2: segmented (counting sort by label), 3: shared-memory table, 4: tile-sorted
    // stats of the last screen call
    double stat_cand_chunks = 0, stat_fallback_frames = 0, stat_screen_frames = 0;
    bool stat_pending = false;
    // screen plan of the last b2k_assign / b2k_dev_assign call, kept so that chunked assignment does not
    // reallocate the operand buffers for every chunk (owned here, freed by b2k_ctx_destroy)
    void* assign_plan = nullptr;
    void* stat_plan = nullptr;  // plan the pending statistics belong to (a Lloyd session's, else assign_plan)
    // grow-only device buffers reused by the host-pointer entry points (frames, labels, ...): a 400 MB
    // cudaMalloc/cudaFree pair per call costs milliseconds and serialises the device
    enum { SLOT_FRAMES = 0, SLOT_LABELS, SLOT_CENTERS, SLOT_CENTERS2, SLOT_ACC, SLOT_CHUNK_X0, SLOT_CHUNK_X1,
           SLOT_CHUNK_L0, SLOT_CHUNK_L1, SLOT_CHUNK_G0, SLOT_CHUNK_G1, N_SLOTS };
    void* slot_ptr[N_SLOTS] = {};
    size_t slot_cap[N_SLOTS] = {};
    int slot(int which, size_t bytes, void** out);
    // 256 bytes of device flag words, allocated once by b2k_ctx_create and NEVER reallocated (the NaN/inf flag of a
    // streamed assign lives here across calls that may grow `scratch`): [0] absmax / all_finite / dtraj range flag,
    // [8] the non-finite flag of stream_assign
    int* flags = nullptr;
    // generic device scratch (grown on demand)
    void* scratch = nullptr;
    size_t scratch_cap = 0;
    int ensure_scratch(size_t bytes);
    void* scratch2 = nullptr;  // second scratch (per-CTA label histograms of the segmented member sums)
    size_t scratch2_cap = 0;
    int ensure_scratch2(size_t bytes);
};

namespace b2k {
// CUDA-event pair around a group of launches on ctx->stream when option "profile" is on (no cost otherwise)
struct ProfScope {
    b2k_ctx* ctx;
    int cls;
    cudaEvent_t e1 = nullptr;
    ProfScope(b2k_ctx* c, int k) : ctx(c), cls(k) {
        if (!ctx->profile) return;
        cudaEvent_t e0 = nullptr;
        if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) { e1 = nullptr; return; }
        cudaEventRecord(e0, ctx->stream);
        ctx->prof_class[cls].push_back(e0);
    }
    ~ProfScope() {
        if (!e1) return;
        cudaEventRecord(e1, ctx->stream);
        ctx->prof_class[cls].push_back(e1);
    }
};
}  // namespace b2k


// ---- device helpers: the exact fp32 arithmetic of the reference path -------------------------
namespace b2k {

// Euclidean squared distance in the pinned reference order:
// 4 interleaved lane accumulators over i<4*(d/4), the d%4 tail into lane 0, ((a0+a1)+a2)+a3.
// __f*_rn intrinsics are never contracted into FMA by nvcc.
struct Lanes4 {
    float a0, a1, a2, a3;
    __device__ __forceinline__ void init() { a0 = a1 = a2 = a3 = 0.f; }
    __device__ __forceinline__ void add4(float x0, float x1, float x2, float x3, float c0, float c1, float c2,
                                          float c3) {
        float t0 = __fsub_rn(x0, c0), t1 = __fsub_rn(x1, c1), t2 = __fsub_rn(x2, c2), t3 = __fsub_rn(x3, c3);
        a0 = __fadd_rn(a0, __fmul_rn(t0, t0));
        a1 = __fadd_rn(a1, __fmul_rn(t1, t1));
        a2 = __fadd_rn(a2, __fmul_rn(t2, t2));
        a3 = __fadd_rn(a3, __fmul_rn(t3, t3));
    }
    __device__ __forceinline__ void tail(float x, float c) {
        float t = __fsub_rn(x, c);
        a0 = __fadd_rn(a0, __fmul_rn(t, t));
    }
    __device__ __forceinline__ float result() const {
        return __fadd_rn(__fadd_rn(__fadd_rn(a0, a1), a2), a3);  // (0+a0)==a0 exactly
    }
};

// one frame row (d <= DREG floats) into registers, zero padded to DREG, with the widest loads the row pitch and the
// base alignment allow (vec = 4: d % 4 == 0 and 16-byte aligned base, 2: d even and 8-byte aligned, else 1).
// A warp reading 32 consecutive rows touches the same cache lines with every load instruction, so fewer, wider loads
// cut the L1 wavefronts proportionally.
__host__ __device__ __forceinline__ int row_load_width(const float* X, int d, int vmax = 4) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(X);
    if (vmax >= 4 && (d & 3) == 0 && (a & 15) == 0) return 4;
    if (vmax >= 2 && (d & 1) == 0 && (a & 7) == 0) return 2;
    return 1;
}
template <int DREG>
__device__ __forceinline__ void load_row_padded(const float* __restrict__ X, int64_t i, int d, int vec, float (&xr)[DREG]) {
    const float* row = X + i * d;
    if (vec == 4) {
#pragma unroll
        for (int e = 0; e < DREG; e += 4) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e < d) v = __ldg(reinterpret_cast<const float4*>(row + e));
            xr[e] = v.x; xr[e + 1] = v.y; xr[e + 2] = v.z; xr[e + 3] = v.w;
        }
    } else if (vec == 2) {
#pragma unroll
        for (int e = 0; e < DREG; e += 2) {
            float2 v = make_float2(0.f, 0.f);
            if (e < d) v = __ldg(reinterpret_cast<const float2*>(row + e));
            xr[e] = v.x; xr[e + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int e = 0; e < DREG; ++e) xr[e] = e < d ? __ldg(row + e) : 0.f;
    }
}

// generic-pointer version (global or shared), any d
__device__ __forceinline__ float euclid_sq_exact(const float* __restrict__ x, const float* __restrict__ c, int d) {
    Lanes4 L;
    L.init();
    const int d4 = d & ~3;
    for (int i = 0; i < d4; i += 4) L.add4(x[i], x[i + 1], x[i + 2], x[i + 3], c[i], c[i + 1], c[i + 2], c[i + 3]);
    for (int i = d4; i < d; ++i) L.tail(x[i], c[i]);
    return L.result();
}

// "first minimum AFTER the sqrt" bookkeeping on squared values.
// The reference compares dj = sqrt(s_j) with strict '<' scanning j upwards.  sqrt is monotone,
// so s_c >= best_s can never win; s_c < best_s wins unless both round to the same sqrt, which
// needs s_c >= best_s*(1-2^-20) (two floats closer than that can share a correctly rounded sqrt).
struct ArgMin {
    float s;    // squared distance of the current winner (+inf: none yet)
    int32_t j;  // its index (-1: none)
    __device__ __forceinline__ void init() { s = __int_as_float(0x7f800000); j = -1; }
    // candidate with index larger than every index seen so far by THIS scanner
    __device__ __forceinline__ void offer(float sc, int32_t jc) {
        if (sc < s) {
            if (sc < s * 0.99999905f || __fsqrt_rn(sc) < __fsqrt_rn(s)) { s = sc; j = jc; }
        }
    }
    // order-free merge of two partial winners: lexicographic (sqrt(s), j)
    __device__ __forceinline__ void merge(float so, int32_t jo) {
        if (jo < 0) return;
        if (j < 0) { s = so; j = jo; return; }
        if (so < s * 0.99999905f) { s = so; j = jo; return; }  // far enough apart: the sqrt cannot merge them
        if (s < so * 0.99999905f) return;
        const float ra = __fsqrt_rn(s), rb = __fsqrt_rn(so);
        if (rb < ra || (rb == ra && jo < j)) { s = so; j = jo; }
    }
};

}  // namespace b2k
