// api.cu -- extern "C" entry points of libb2k (see include/b2k.h for what each one replaces).
#include "common.cuh"
#include "kernels.h"
#include <cstdarg>
#include <algorithm>
#include <thread>
#include <map>
#include <mutex>
#include <unordered_map>

namespace b2k {

thread_local std::string g_last_error;
std::atomic<long long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

// ---- device block cache ----------------------------------------------------------------------
// cudaFree of the GB-sized working buffers of a Lloyd session (fp16 screen operand, sorted copy of the frames, candidate
// lists) costs ~0.12 s per GB on these boxes and the next cudaMalloc pays again: a 10-iteration fit of 1e7 x 10 frames
// spent 270-340 ms in b2k_dev_lloyd_destroy and up to 300 ms in the next session's first step, against 45 ms of
// kernels.  Freed blocks therefore stay in a per-device free list (same device synchronisation as cudaFree, so a block
// is never handed out while work that uses it is in flight) and are handed back to the next request of a similar
// size.  The list is bounded (option "cache_mb", default half of the device memory), released oldest first, released
// entirely when a cudaMalloc fails, on b2k_ctx_destroy, and on request (option "cache_release") -- e.g. before a
// caller that allocates with another allocator (torch) needs the room.
namespace {
struct CacheBlock { size_t size; int dev; unsigned long long seq; };
std::mutex g_cache_mu;
std::unordered_map<void*, CacheBlock> g_live;                       // blocks handed out
std::map<int, std::multimap<size_t, std::pair<void*, unsigned long long>>> g_free;  // per device: size -> (ptr, age)
std::map<int, size_t> g_free_bytes;
std::map<int, long long> g_cache_limit;                             // bytes; absent: default
unsigned long long g_cache_seq = 0;

size_t cache_round(size_t n) {
    const size_t g = n < (size_t(1) << 20) ? 4096 : (size_t(2) << 20);
    return (std::max<size_t>(n, 1) + g - 1) / g * g;
}

long long cache_limit_locked(int dev) {
    auto it = g_cache_limit.find(dev);
    if (it != g_cache_limit.end() && it->second >= 0) return it->second;
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) { cudaGetLastError(); tot = 0; }
    const long long lim = (long long)(tot / 2);
    g_cache_limit[dev] = lim;
    return lim;
}

// cudaFree of cached blocks of `dev` (-1: every device), oldest first, until at most `keep` bytes stay
void cache_shrink_locked(int dev, size_t keep) {
    for (auto& kv : g_free) {
        if (dev >= 0 && kv.first != dev) continue;
        auto& fl = kv.second;
        size_t& bytes = g_free_bytes[kv.first];
        while (bytes > keep && !fl.empty()) {
            auto oldest = fl.begin();
            for (auto it = fl.begin(); it != fl.end(); ++it)
                if (it->second.second < oldest->second.second) oldest = it;
            cudaFree(oldest->second.first);
            bytes -= oldest->first;
            fl.erase(oldest);
        }
    }
    cudaGetLastError();
}
}  // namespace

cudaError_t dev_alloc_raw(void** p, size_t bytes) {
    *p = nullptr;
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t want = cache_round(bytes);
    std::lock_guard<std::mutex> lk(g_cache_mu);
    auto& fl = g_free[dev];
    auto it = fl.lower_bound(want);
    if (it != fl.end() && it->first <= want + want / 4) {
        *p = it->second.first;
        g_live[*p] = CacheBlock{it->first, dev, 0};
        g_free_bytes[dev] -= it->first;
        fl.erase(it);
        return cudaSuccess;
    }
    cudaError_t e = cudaMalloc(p, want);
    if (e != cudaSuccess) {  // make room: everything this library keeps for later goes back to the driver first
        cudaGetLastError();
        cache_shrink_locked(dev, 0);
        e = cudaMalloc(p, want);
    }
    if (e != cudaSuccess) { *p = nullptr; return e; }
    g_live[*p] = CacheBlock{want, dev, 0};
    return cudaSuccess;
}

void dev_free(void* p) {
    if (!p) return;
    CacheBlock b;
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        auto it = g_live.find(p);
        if (it == g_live.end()) { cudaFree(p); return; }  // not ours (never happens inside the library)
        b = it->second;
        g_live.erase(it);
    }
    // what cudaFree guarantees: nothing submitted so far still touches the block
    int cur = 0;
    cudaGetDevice(&cur);
    if (cur != b.dev) cudaSetDevice(b.dev);
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(g_cache_mu);
    const long long lim = cache_limit_locked(b.dev);
    if ((long long)b.size > lim) {
        cudaFree(p);
    } else {
        g_free[b.dev].emplace(b.size, std::make_pair(p, ++g_cache_seq));
        g_free_bytes[b.dev] += b.size;
        if ((long long)g_free_bytes[b.dev] > lim) cache_shrink_locked(b.dev, (size_t)lim);
    }
    if (cur != b.dev) cudaSetDevice(cur);
}

void dev_cache_release(int dev) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    cache_shrink_locked(dev, 0);
}

size_t dev_cache_bytes(int dev) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    auto it = g_free_bytes.find(dev);
    return it == g_free_bytes.end() ? 0 : it->second;
}

void dev_cache_set_limit(int dev, long long bytes) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    if (bytes < 0) g_cache_limit.erase(dev);
    else {
        g_cache_limit[dev] = bytes;
        cache_shrink_locked(dev, (size_t)bytes);
    }
}

int kmpp_run(b2k_ctx* ctx, const float* dX, int64_t n, int d, int k, int metric, int64_t seed, int scan_mode,
             b2k_callback cb, void* user, float* dcenters_out, int64_t* chosen_host);

int kmpp_run_blocked(b2k_ctx* ctx, const float* dX, int64_t n, int d, int k, int metric, int64_t seed, int64_t lo,
                     int64_t n_total, float* xf, int64_t xf_len, int64_t* xi, b2k_exchange_fn ex, void* exuser,
                     b2k_callback cb, void* user, float* dcenters_out, int64_t* chosen_host);


// pageable host memory -> pinned staging slot.  One memcpy thread moves ~10 GB/s, PCIe Gen5 takes 55: large copies
// are split over a few threads so that the bounce stays off the critical path (option "host_copy_threads").
static void host_copy(const b2k_ctx* ctx, void* dst, const void* src, size_t bytes) {
    const int nt = (int)std::min<size_t>((size_t)std::max(ctx->host_copy_threads, 1), bytes / (size_t(4) << 20));
    if (nt <= 1) { std::memcpy(dst, src, bytes); return; }
    std::vector<std::thread> th;
    const size_t part = ((bytes / nt) + 4095) & ~size_t(4095);
    for (int t = 1; t < nt; ++t) {
        const size_t off = (size_t)t * part;
        if (off >= bytes) break;
        th.emplace_back([=] { std::memcpy((char*)dst + off, (const char*)src + off, std::min(part, bytes - off)); });
    }
    std::memcpy(dst, src, std::min(part, bytes));
    for (auto& x : th) x.join();
}

static bool host_ptr_is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

static int ensure_pinned(b2k_ctx* ctx, size_t in_bytes, size_t out_bytes) {
    if (in_bytes > ctx->pinned_cap) {
        for (int s = 0; s < 2; ++s) {
            if (ctx->pinned[s]) cudaFreeHost(ctx->pinned[s]);
            ctx->pinned[s] = nullptr;
            CUDA_TRY(cudaMallocHost(&ctx->pinned[s], in_bytes));
        }
        ctx->pinned_cap = in_bytes;
    }
    if (out_bytes > ctx->pinned_out_cap) {
        for (int s = 0; s < 2; ++s) {
            if (ctx->pinned_out[s]) cudaFreeHost(ctx->pinned_out[s]);
            ctx->pinned_out[s] = nullptr;
            CUDA_TRY(cudaMallocHost(&ctx->pinned_out[s], out_bytes));
        }
        ctx->pinned_out_cap = out_bytes;
    }
    return B2K_OK;
}

// Copy a host array to the device in chunks: pageable sources bounce through the two pinned
// staging slots (memcpy of chunk c+1 overlaps the DMA of chunk c); pinned sources DMA directly.
int upload_host(b2k_ctx* ctx, const void* src, void* dst, size_t bytes) {
    if (bytes == 0) return B2K_OK;
    if (host_ptr_is_pinned(src)) {
        CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
        return B2K_OK;
    }
    const size_t chunk = std::min(bytes, ctx->stage_bytes);
    B2K_TRY(ensure_pinned(ctx, chunk, 0));
    size_t off = 0;
    int c = 0;
    while (off < bytes) {
        const int s = c & 1;
        const size_t len = std::min(chunk, bytes - off);
        CUDA_TRY(cudaEventSynchronize(ctx->ev_done[s]));
        host_copy(ctx, ctx->pinned[s], (const char*)src + off, len);
        CUDA_TRY(cudaMemcpyAsync((char*)dst + off, ctx->pinned[s], len, cudaMemcpyHostToDevice, ctx->copy_stream[s]));
        CUDA_TRY(cudaEventRecord(ctx->ev_done[s], ctx->copy_stream[s]));
        off += len;
        ++c;
    }
    for (int s = 0; s < 2; ++s) CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->ev_done[s], 0));
    return B2K_OK;
}

static int check_metric_dim(int metric, int d) {
    if (metric != B2K_METRIC_EUCLIDEAN && metric != B2K_METRIC_MINRMSD)
        return set_error(B2K_ERR_INVALID_ARG, "unknown metric id %d", metric);
    if (metric == B2K_METRIC_MINRMSD && d % 3)
        return set_error(B2K_ERR_DIM_NOT_MULT3,
                         "RMSDMetric is only implemented for input data with a dimension divisible by 3.");
    return B2K_OK;
}

// ---- metric-generic primitives -----------------------------------------------------------------
// centers prepared for a metric (minRMSD: centered copies + traces, owned here)
struct PreparedCenters {
    const float* C = nullptr;  // what the kernels read (euclid: the caller's; rmsd: centered copy)
    float* Gb = nullptr;
    DevMem mem_c, mem_g;
    int prepare(b2k_ctx* ctx, const float* dC, int k, int d, int metric) {
        if (metric == B2K_METRIC_MINRMSD) {
            B2K_TRY(mem_c.alloc((size_t)k * d * 4));
            B2K_TRY(mem_g.alloc((size_t)k * 4));
            B2K_TRY(launch_rmsd_center(ctx, dC, k, d, mem_c.as<float>(), mem_g.as<float>()));
            C = mem_c.as<float>();
            Gb = mem_g.as<float>();
        } else {
            C = dC;
        }
        return B2K_OK;
    }
};

static int assign_any(b2k_ctx* ctx, const float* dX, const float* Ga, int64_t n, int d, const PreparedCenters& pc,
                      int k, int metric, int32_t* labels, float* mind, int lloyd) {
    if (metric == B2K_METRIC_MINRMSD) return launch_rmsd_assign(ctx, dX, Ga, n, d, pc.C, pc.Gb, k, labels, mind, lloyd);
    return launch_assign_exact(ctx, dX, n, d, pc.C, k, labels, mind, lloyd);
}

}  // namespace b2k

using namespace b2k;

int b2k_ctx::slot(int which, size_t bytes, void** out) {
    if (bytes > slot_cap[which]) {
        if (slot_ptr[which]) {
            cudaStreamSynchronize(stream);
            b2k::dev_free(slot_ptr[which]);
        }
        slot_ptr[which] = nullptr;
        slot_cap[which] = 0;
        const size_t want = bytes + bytes / 8;  // a little headroom for the next, slightly larger call
        if (b2k::dev_alloc(&slot_ptr[which], want) != cudaSuccess) {
            cudaGetLastError();
            if (b2k::dev_alloc(&slot_ptr[which], bytes) != cudaSuccess) {
                cudaGetLastError();
                slot_ptr[which] = nullptr;
                return b2k::set_error(B2K_ERR_NOMEM, "cudaMalloc(%zu bytes) failed", bytes);
            }
            slot_cap[which] = bytes;
        } else {
            slot_cap[which] = want;
        }
    }
    *out = slot_ptr[which];
    return B2K_OK;
}

int b2k_ctx::ensure_scratch(size_t bytes) {
    if (bytes <= scratch_cap) return B2K_OK;
    if (scratch) b2k::dev_free(scratch);
    scratch = nullptr;
    scratch_cap = 0;
    CUDA_TRY(b2k::dev_alloc(&scratch, bytes));
    scratch_cap = bytes;
    return B2K_OK;
}

int b2k_ctx::ensure_scratch2(size_t bytes) {
    if (bytes <= scratch2_cap) return B2K_OK;
    if (scratch2) b2k::dev_free(scratch2);
    scratch2 = nullptr;
    scratch2_cap = 0;
    CUDA_TRY(b2k::dev_alloc(&scratch2, bytes));
    scratch2_cap = bytes;
    return B2K_OK;
}

// =================================================================================================
B2K_API const char* b2k_last_error(void) { return g_last_error.c_str(); }
B2K_API int b2k_version(void) { return 100; }
B2K_API int64_t b2k_launch_count(void) { return g_launches.load(); }

B2K_API int b2k_ctx_create(int device, b2k_ctx** out) {
    if (!out) return set_error(B2K_ERR_INVALID_ARG, "ctx_create: null out");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return set_error(B2K_ERR_CUDA, "no CUDA device available: libb2k has no CPU fallback");
    }
    if (device < 0 || device >= count) return set_error(B2K_ERR_INVALID_ARG, "device %d out of range", device);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return set_error(B2K_ERR_CUDA, "device %d is sm_%d%d; libb2k is built for sm_100a only", device, prop.major,
                         prop.minor);
    b2k_ctx* c = new b2k_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
    CUDA_TRY(cudaMalloc(&c->flags, 256));
    CUDA_TRY(cudaMemset(c->flags, 0, 256));
    for (int s = 0; s < 2; ++s) {
        CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_stream[s], cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&c->ev_h2d[s], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&c->ev_done[s], cudaEventDisableTiming));
    }
    *out = c;
    return B2K_OK;
}

B2K_API int b2k_ctx_destroy(b2k_ctx* c) {
    if (!c) return B2K_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (int s = 0; s < 2; ++s) {
        if (c->pinned[s]) cudaFreeHost(c->pinned[s]);
        if (c->pinned_out[s]) cudaFreeHost(c->pinned_out[s]);
        if (c->copy_stream[s]) cudaStreamDestroy(c->copy_stream[s]);
        if (c->ev_h2d[s]) cudaEventDestroy(c->ev_h2d[s]);
        if (c->ev_done[s]) cudaEventDestroy(c->ev_done[s]);
    }
    screen_plan_release_cached(c);
    for (cudaEvent_t e : c->prof_events) cudaEventDestroy(e);
    for (auto& v : c->prof_class) for (cudaEvent_t e : v) cudaEventDestroy(e);
    for (int i = 0; i < b2k_ctx::N_SLOTS; ++i)
        if (c->slot_ptr[i]) dev_free(c->slot_ptr[i]);
    if (c->flags) cudaFree(c->flags);
    if (c->scratch) dev_free(c->scratch);
    if (c->scratch2) dev_free(c->scratch2);
    dev_cache_release(c->device);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return B2K_OK;
}

B2K_API int b2k_ctx_set_stream(b2k_ctx* c, void* cuda_stream) {
    if (!c) return set_error(B2K_ERR_INVALID_ARG, "null ctx");
    // the handle is used as is: NULL is CUDA's legacy default stream (what torch uses unless told otherwise)
    if (c->own_stream && c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    c->stream = (cudaStream_t)cuda_stream;
    c->own_stream = false;
    return B2K_OK;
}

B2K_API int b2k_ctx_sync(b2k_ctx* c) {
    if (!c) return set_error(B2K_ERR_INVALID_ARG, "null ctx");
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return B2K_OK;
}

B2K_API int b2k_ctx_set_option(b2k_ctx* c, const char* name, int64_t value) {
    if (!c || !name) return set_error(B2K_ERR_INVALID_ARG, "null argument");
    if (!strcmp(name, "assign_engine")) c->engine = (int)value;
    else if (!strcmp(name, "screen_terms")) c->screen_terms = (int)value;
    else if (!strcmp(name, "probe_max_centers")) c->probe_max_centers = (int)value;
    else if (!strcmp(name, "probe_min_gflop")) c->probe_min_gflop = (int)value;
    else if (!strcmp(name, "screen_gather")) c->screen_gather = (int)value;
    else if (!strcmp(name, "screen_decide")) c->screen_decide = (int)value;
    else if (!strcmp(name, "rmsd_abandon")) c->rmsd_abandon = (int)value;
    else if (!strcmp(name, "kmpp_async")) c->kmpp_async = (int)value;
    else if (!strcmp(name, "prune_mode")) c->prune_mode = (int)value;
    else if (!strcmp(name, "prune_resort")) c->prune_resort = (int)value;
    else if (!strcmp(name, "delta_sums")) c->delta_sums = (int)value;
    else if (!strcmp(name, "prune_list_margin")) c->prune_list_margin = (int)std::max<int64_t>(0, value);
    else if (!strcmp(name, "prune_unit_shift")) c->prune_unit_shift = (int)value;
    else if (!strcmp(name, "screen_group")) c->screen_group = (int)value;
    else if (!strcmp(name, "screen_resident_a")) c->screen_resident_a = (int)value;
    else if (!strcmp(name, "screen_cluster")) c->screen_cluster = (int)value;
    else if (!strcmp(name, "verify_mode")) c->verify_mode = (int)value;
    else if (!strcmp(name, "fallback_mode")) c->fallback_mode = (int)value;
    else if (!strcmp(name, "operand_kernel")) c->operand_kernel = (int)value;
    else if (!strcmp(name, "kmpp_prune")) c->kmpp_prune = (int)value;
    else if (!strcmp(name, "check_finite")) c->check_finite = value != 0;
    else if (!strcmp(name, "host_copy_threads")) c->host_copy_threads = (int)std::max<int64_t>(1, std::min<int64_t>(value, 32));
    else if (!strcmp(name, "accumulate_mode")) c->accumulate_mode = (int)value;
    else if (!strcmp(name, "cost_kernel")) c->cost_kernel = (int)value;
    else if (!strcmp(name, "row_vec_max")) c->row_vec_max = (int)value;
    else if (!strcmp(name, "rmsd_kernel")) c->rmsd_kernel = (int)value;
    else if (!strcmp(name, "profile")) {  // (re)start event timing of the screen kernel launches
        for (cudaEvent_t e : c->prof_events) cudaEventDestroy(e);
        c->prof_events.clear();
        for (auto& v : c->prof_class) { for (cudaEvent_t e : v) cudaEventDestroy(e); v.clear(); }
        c->profile = value != 0;
    }
    else if (!strcmp(name, "cache_mb")) {  // bound of the device block cache (-1: default, half of the device memory; 0: off)
        CUDA_TRY(cudaSetDevice(c->device));
        dev_cache_set_limit(c->device, value < 0 ? -1 : (long long)value << 20);
    }
    else if (!strcmp(name, "cache_release")) dev_cache_release(c->device);  // give every cached block back to the driver
    else if (!strcmp(name, "stage_bytes")) c->stage_bytes = (size_t)std::max<int64_t>(value, 1 << 16);
    else if (!strcmp(name, "own_stream")) {
        if (!c->own_stream) {
            CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
            c->own_stream = true;
        }
    }
    else return set_error(B2K_ERR_INVALID_ARG, "unknown option '%s'", name);
    return B2K_OK;
}

B2K_API int b2k_ctx_get_stat(b2k_ctx* c, const char* name, double* value) {
    if (!c || !name || !value) return set_error(B2K_ERR_INVALID_ARG, "null argument");
    if (c->stat_pending && !strncmp(name, "screen_", 7)) {  // stats of the last screened assign, read lazily
        void* plan = c->stat_plan ? c->stat_plan : c->assign_plan;
        if (plan) B2K_TRY(screen_read_stats(static_cast<ScreenPlan*>(plan), &c->stat_cand_chunks, &c->stat_fallback_frames));
        c->stat_pending = false;
    }
    if (!strcmp(name, "screen_gemm_ms_total") || !strcmp(name, "screen_gemm_launches")) {
        double total = 0;
        for (size_t i = 0; i + 1 < c->prof_events.size(); i += 2) {
            CUDA_TRY(cudaEventSynchronize(c->prof_events[i + 1]));
            float ms = 0.f;
            CUDA_TRY(cudaEventElapsedTime(&ms, c->prof_events[i], c->prof_events[i + 1]));
            total += ms;
        }
        *value = !strcmp(name, "screen_gemm_ms_total") ? total : (double)(c->prof_events.size() / 2);
        return B2K_OK;
    }
    if (!strncmp(name, "prof_ms_", 8) || !strncmp(name, "prof_n_", 7)) {
        const bool want_ms = name[5] == 'm';
        const char* cls = name + (want_ms ? 8 : 7);
        static const char* names[b2k_ctx::PROF_N] = {"verify", "sums", "cost", "lists"};
        for (int k = 0; k < b2k_ctx::PROF_N; ++k) {
            if (strcmp(cls, names[k])) continue;
            auto& v = c->prof_class[k];
            double total = 0;
            for (size_t i = 0; i + 1 < v.size(); i += 2) {
                CUDA_TRY(cudaEventSynchronize(v[i + 1]));
                float ms = 0.f;
                CUDA_TRY(cudaEventElapsedTime(&ms, v[i], v[i + 1]));
                total += ms;
            }
            *value = want_ms ? total : (double)(v.size() / 2);
            return B2K_OK;
        }
        return set_error(B2K_ERR_INVALID_ARG, "unknown stat '%s'", name);
    }
    if (!strcmp(name, "screen_cand_chunks")) *value = c->stat_cand_chunks;
    else if (!strcmp(name, "screen_fallback_frames")) *value = c->stat_fallback_frames;
    else if (!strcmp(name, "screen_frames")) *value = c->stat_screen_frames;
    else if (!strcmp(name, "screen_terms_used")) *value = c->stat_screen_terms;
    else if (!strcmp(name, "kmpp_async_fallbacks")) *value = c->stat_kmpp_async_fallbacks;
    else if (!strcmp(name, "prune_mean_list")) *value = c->stat_prune_mean;
    else if (!strcmp(name, "prune_steps")) *value = c->stat_prune_steps;
    else if (!strcmp(name, "prune_sorts")) *value = c->stat_prune_sorts;
    else if (!strcmp(name, "labels_changed")) *value = c->stat_changed;
    else if (!strcmp(name, "delta_steps")) *value = c->stat_delta_steps;
    else if (!strcmp(name, "list_reuse_steps")) *value = c->stat_list_reuse;
    else if (!strncmp(name, "probe_centers_", 14) && name[14] >= '1' && name[14] <= '3') *value = c->stat_probe_centers[name[14] - '0'];
    else if (!strncmp(name, "probe_fallback_", 15) && name[15] >= '1' && name[15] <= '3') *value = c->stat_probe_fallback[name[15] - '0'];
    else if (!strcmp(name, "sm_count")) *value = c->sm_count;
    else if (!strcmp(name, "cache_bytes")) *value = (double)dev_cache_bytes(c->device);
    else if (!strcmp(name, "fp32_lane_instr_per_s")) {  // measured now: non-fusable FMUL+FADD chains on every SM
        CUDA_TRY(cudaSetDevice(c->device));
        return measure_fp32_rate(c, value);
    }
    else return set_error(B2K_ERR_INVALID_ARG, "unknown stat '%s'", name);
    return B2K_OK;
}

// host array -> device array through the context's pinned staging (multi-threaded bounce for pageable memory);
// returns when the data is in place
B2K_API int b2k_upload(b2k_ctx* ctx, const void* src_host, void* dst_dev, int64_t bytes) {
    if (!ctx || bytes < 0 || (bytes > 0 && (!src_host || !dst_dev))) return set_error(B2K_ERR_INVALID_ARG, "upload: bad arguments");
    CUDA_TRY(cudaSetDevice(ctx->device));
    B2K_TRY(upload_host(ctx, src_host, dst_dev, (size_t)bytes));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return B2K_OK;
}

// ---- compute_metric ---------------------------------------------------------------------------
B2K_API int b2k_compute_metric(b2k_ctx* ctx, const float* x, const float* y, int64_t d, int metric, float* out) {
    if (!ctx || !x || !y || !out || d < 1) return set_error(B2K_ERR_INVALID_ARG, "compute_metric: bad arguments");
    B2K_TRY(check_metric_dim(metric, (int)d));
    CUDA_TRY(cudaSetDevice(ctx->device));
    DevMem buf;
    B2K_TRY(buf.alloc((size_t)(2 * d + 8) * 4 + 64));
    float* dx = buf.as<float>();
    float* dy = dx + d;
    float* dres = dy + d;
    CUDA_TRY(cudaMemcpyAsync(dx, x, d * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(dy, y, d * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (metric == B2K_METRIC_MINRMSD) {
        DevMem t;
        B2K_TRY(t.alloc((size_t)(d + 8) * 4));
        float* yc = t.as<float>();
        float* g = dres + 1;
        B2K_TRY(launch_rmsd_center(ctx, dx, 1, (int)d, nullptr, g));          // Ga
        B2K_TRY(launch_rmsd_center(ctx, dy, 1, (int)d, yc, g + 1));           // centered y, Gb
        B2K_TRY(launch_rmsd_dist_rows(ctx, dx, g, 1, (int)d, yc, g + 1, 1, dres));
        CUDA_TRY(cudaMemcpyAsync(out, dres, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return B2K_OK;
    }
    B2K_TRY(launch_dist_rows(ctx, dx, 1, (int)d, dy, 1, dres));
    CUDA_TRY(cudaMemcpyAsync(out, dres, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return B2K_OK;
}

// ---- assign -----------------------------------------------------------------------------------
B2K_API int b2k_dev_assign(b2k_ctx* ctx, const float* dX, int64_t n, int32_t d, const float* dC, int32_t k,
                           int metric, int32_t* dlabels, float* dmind) {
    if (!ctx || n < 0 || d < 1 || k < 1 || (n > 0 && (!dX || !dlabels)) || !dC)
        return set_error(B2K_ERR_INVALID_ARG, "assign: bad arguments (n=%lld d=%d k=%d)", (long long)n, d, k);
    B2K_TRY(check_metric_dim(metric, d));
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (n == 0) return B2K_OK;
    PreparedCenters pc;
    B2K_TRY(pc.prepare(ctx, dC, k, d, metric));
    DevMem ga;
    if (metric == B2K_METRIC_MINRMSD) {
        // traces of the centred frames: a grow-only slot of the context (no allocation per call)
        B2K_TRY(ctx->slot(b2k_ctx::SLOT_CHUNK_G0, (size_t)n * 4, &ga.p));
        float* ga_slot = (float*)ga.p;
        ga.p = nullptr;  // not owned: ~DevMem must not free a slot
        B2K_TRY(launch_rmsd_center(ctx, dX, n, d, nullptr, ga_slot));
        B2K_TRY(assign_any(ctx, dX, ga_slot, n, d, pc, k, metric, dlabels, dmind, 0));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // the centred centers (pc) die with this frame
        return B2K_OK;
    } else if (ctx->engine != B2K_ENGINE_DIRECT && screen_supported(ctx, d, k, n)) {
        // the plan (fp16 operand + candidate lists, Kp*2 + ~41 bytes per frame) is sized to what is free: frames beyond
        // its capacity are assigned piece by piece through the same plan
        ScreenPlan* plan = nullptr;
        int64_t cap = n;
        int terms = 0;
        B2K_TRY(screen_choose_terms(ctx, dX, n, d, dC, k, &terms));
        int rc = screen_plan_acquire(ctx, cap, d, k, terms, &plan);
        while (rc == B2K_ERR_NOMEM && cap > (int64_t(1) << 16)) {
            cap = (cap + 1) / 2;
            rc = screen_plan_acquire(ctx, cap, d, k, terms, &plan);
        }
        if (rc == B2K_OK) {
            for (int64_t off = 0; off < n && rc == B2K_OK; off += cap) {
                const int64_t len = std::min(cap, n - off);
                rc = screen_prepare_frames(plan, dX + off * d, len);
                if (rc == B2K_OK)
                    rc = screen_assign(plan, dX + off * d, len, dC, dlabels + off, dmind ? dmind + off : nullptr, 0);
            }
            ctx->stat_screen_frames = (double)std::min(cap, n);
            ctx->stat_screen_terms = (double)terms;
            ctx->stat_plan = nullptr;
            ctx->stat_pending = rc == B2K_OK;
            return rc;
        }
        if (rc != B2K_ERR_NOMEM) return rc;
        // not even a 64k-frame plan fits: exact engine below
    }
    B2K_TRY(assign_any(ctx, dX, ga.as<float>(), n, d, pc, k, metric, dlabels, dmind, 0));
    if (pc.mem_c.p || ga.p) CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // temporaries die with this frame
    return B2K_OK;
}

// Host frames are streamed chunk by chunk: H2D of chunk c+1 (copy stream) overlaps the kernels of
// chunk c (compute stream) and the D2H of chunk c-1's labels.  With dX_keep / dL_keep the chunks (and their
// labels) additionally stay resident in one device array, which is how b2k_kmeans_cluster gets its frames
// into HBM while the assignment of the first chunks is already running.
// dprev / cost_slot (out-of-core Lloyd pass): before a chunk is assigned, the squared distances of its frames to the
// centers their PREVIOUS labels (dprev, device, n ints; may alias dL_keep) name are added to *cost_slot (exact integer
// sum, cost_scale) -- the cost of the previous iteration, taken while the frames are on the device anyway.
static int stream_assign(b2k_ctx* ctx, const float* X, int64_t n, int32_t d, const float* dC, int32_t k, int metric,
                         int32_t* labels, int lloyd, float* dX_keep, int32_t* dL_keep, double acc_scale = 0.0,
                         int64_t* dacc = nullptr, const int32_t* dprev = nullptr, double cost_scale = 0.0,
                         int64_t* cost_slot = nullptr) {
    cudaStream_t st = ctx->stream;
    PreparedCenters pc;
    B2K_TRY(pc.prepare(ctx, dC, k, d, metric));
    const int64_t row_bytes = (int64_t)d * 4;
    int64_t cf = std::max<int64_t>(1, (int64_t)ctx->stage_bytes / row_bytes);
    cf = std::min(cf, n);
    const bool want_host = labels != nullptr;
    const bool in_pinned = host_ptr_is_pinned(X), out_pinned = !want_host || host_ptr_is_pinned(labels);
    B2K_TRY(ensure_pinned(ctx, in_pinned ? 0 : (size_t)cf * row_bytes, out_pinned ? 0 : (size_t)cf * 4));
    float* dX[2] = {nullptr, nullptr};
    int32_t* dL[2] = {nullptr, nullptr};
    float* dG[2] = {nullptr, nullptr};
    for (int s = 0; s < 2; ++s) {
        if (!dX_keep) B2K_TRY(ctx->slot(b2k_ctx::SLOT_CHUNK_X0 + s, (size_t)cf * row_bytes, (void**)&dX[s]));
        if (!dL_keep) B2K_TRY(ctx->slot(b2k_ctx::SLOT_CHUNK_L0 + s, (size_t)cf * 4, (void**)&dL[s]));
        if (metric == B2K_METRIC_MINRMSD) B2K_TRY(ctx->slot(b2k_ctx::SLOT_CHUNK_G0 + s, (size_t)cf * 4, (void**)&dG[s]));
    }
    float* dist_buf = nullptr;  // per-frame distances of the cost pass (wide rows / minRMSD)
    if (cost_slot) B2K_TRY(ctx->slot(b2k_ctx::SLOT_LABELS, (size_t)cf * 4, (void**)&dist_buf));
    const bool use_screen = metric == B2K_METRIC_EUCLIDEAN && ctx->engine != B2K_ENGINE_DIRECT &&
                            screen_supported(ctx, d, k, cf);
    ScreenPlan* plan = nullptr;
    if (use_screen) B2K_TRY(screen_plan_acquire(ctx, cf, d, k, 0, &plan));
    cudaEvent_t ev_k[2];
    for (int s = 0; s < 2; ++s) CUDA_TRY(cudaEventCreateWithFlags(&ev_k[s], cudaEventDisableTiming));
    // NaN / inf guard of the reference's chunk iterator (datasource.py:1067-1075), on the device while the chunk is
    // there anyway: one flag for the whole call, read at the end
    int* d_finite = nullptr;
    if (ctx->check_finite) {
        d_finite = ctx->flags + 8;  // dedicated allocation: launch_accumulate may regrow ctx->scratch under us
        const int one = 1;
        CUDA_TRY(cudaMemcpyAsync(d_finite, &one, 4, cudaMemcpyHostToDevice, st));
    }
    int64_t pend_off[2] = {-1, -1}, pend_len[2] = {0, 0};
    int rc = B2K_OK;
    int c = 0;
    for (int64_t off = 0; off < n && rc == B2K_OK; off += cf, ++c) {
        const int s = c & 1;
        const int64_t len = std::min(cf, n - off);
        // slot free? (its previous D2H finished) -> hand the previous labels of this slot to the caller
        cudaEventSynchronize(ctx->ev_done[s]);
        if (pend_off[s] >= 0 && !out_pinned)
            std::memcpy(labels + pend_off[s], ctx->pinned_out[s], (size_t)pend_len[s] * 4);
        pend_off[s] = -1;
        const void* src = X + off * d;
        if (!in_pinned) { host_copy(ctx, ctx->pinned[s], src, (size_t)len * row_bytes); src = ctx->pinned[s]; }
        float* dx = dX_keep ? dX_keep + off * d : dX[s];
        int32_t* dl = dL_keep ? dL_keep + off : dL[s];
        cudaMemcpyAsync(dx, src, (size_t)len * row_bytes, cudaMemcpyHostToDevice, ctx->copy_stream[s]);
        cudaEventRecord(ctx->ev_h2d[s], ctx->copy_stream[s]);
        cudaStreamWaitEvent(st, ctx->ev_h2d[s], 0);
        if (d_finite) rc = launch_all_finite(ctx, dx, len * d, d_finite);
        if (rc != B2K_OK) break;
        if (metric == B2K_METRIC_MINRMSD) rc = launch_rmsd_center(ctx, dx, len, d, nullptr, dG[s]);
        if (rc == B2K_OK && cost_slot) {  // cost of the previous labels against these centers, before dl is overwritten
            const int32_t* pl = dprev + off;
            if (metric == B2K_METRIC_MINRMSD) {
                rc = launch_rmsd_labeled_dist(ctx, dx, dG[s], len, d, pc.C, pc.Gb, pl, dist_buf);
                if (rc == B2K_OK) rc = launch_cost_reduce(ctx, dist_buf, len, cost_scale, cost_slot);
            } else {
                int fused = 0;
                rc = launch_cost_fused(ctx, dx, len, d, dC, k, pl, cost_scale, cost_slot, &fused);
                if (rc == B2K_OK && !fused) {
                    rc = launch_labeled_dist(ctx, dx, len, d, dC, pl, dist_buf);
                    if (rc == B2K_OK) rc = launch_cost_reduce(ctx, dist_buf, len, cost_scale, cost_slot);
                }
            }
        }
        if (rc != B2K_OK) break;
        if (use_screen) {
            rc = screen_prepare_frames(plan, dx, len);
            if (rc == B2K_OK) rc = screen_assign(plan, dx, len, dC, dl, nullptr, lloyd);
        } else {
            rc = assign_any(ctx, dx, dG[s], len, d, pc, k, metric, dl, nullptr, lloyd);
        }
        // member sums of this chunk while the next one is on the bus (exact integer sums: any chunking gives the same bits)
        if (rc == B2K_OK && dacc) rc = launch_accumulate(ctx, dx, len, d, k, dl, acc_scale, dacc);
        cudaEventRecord(ev_k[s], st);
        cudaStreamWaitEvent(ctx->copy_stream[s], ev_k[s], 0);
        if (want_host) {
            void* dst = out_pinned ? (void*)(labels + off) : ctx->pinned_out[s];
            cudaMemcpyAsync(dst, dl, (size_t)len * 4, cudaMemcpyDeviceToHost, ctx->copy_stream[s]);
        }
        cudaEventRecord(ctx->ev_done[s], ctx->copy_stream[s]);
        pend_off[s] = off;
        pend_len[s] = len;
    }
    for (int s = 0; s < 2; ++s) {
        cudaEventSynchronize(ctx->ev_done[s]);
        if (pend_off[s] >= 0 && !out_pinned)
            std::memcpy(labels + pend_off[s], ctx->pinned_out[s], (size_t)pend_len[s] * 4);
        cudaEventDestroy(ev_k[s]);
    }
    cudaStreamSynchronize(st);
    if (use_screen) {  // candidate statistics of the LAST chunk are readable through b2k_ctx_get_stat
        ctx->stat_screen_frames = (double)(n - (int64_t)(c - 1) * cf);
        ctx->stat_plan = nullptr;
        ctx->stat_pending = rc == B2K_OK;
    }
    if (rc != B2K_OK) return rc;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(B2K_ERR_CUDA, "assign: %s", cudaGetErrorString(e));
    if (d_finite) {
        int ok = 1;
        CUDA_TRY(cudaMemcpy(&ok, d_finite, 4, cudaMemcpyDeviceToHost));
        if (!ok) return set_error(B2K_ERR_NONFINITE, "Found invalid values (NaN/inf) in the input frames");
    }
    return B2K_OK;
}

B2K_API int b2k_assign(b2k_ctx* ctx, const float* X, int64_t n, int32_t d, const float* centers, int32_t k,
                       int metric, int32_t* labels) {
    if (!ctx || n < 0 || d < 1 || k < 1 || (n > 0 && (!X || !labels)) || !centers)
        return set_error(B2K_ERR_INVALID_ARG, "assign: bad arguments (n=%lld d=%d k=%d)", (long long)n, d, k);
    B2K_TRY(check_metric_dim(metric, d));
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (n == 0) return B2K_OK;
    float* dC;
    B2K_TRY(ctx->slot(b2k_ctx::SLOT_CENTERS, (size_t)k * d * 4, (void**)&dC));
    CUDA_TRY(cudaMemcpyAsync(dC, centers, (size_t)k * d * 4, cudaMemcpyHostToDevice, ctx->stream));
    return stream_assign(ctx, X, n, d, dC, k, metric, labels, 0, nullptr, nullptr);
}

B2K_API int b2k_stage_assign(b2k_ctx* ctx, const float* X, int64_t n, int32_t d, const float* dcenters, int32_t k,
                             int metric, int lloyd, float* dX_out, int32_t* dlabels_out, int32_t* labels_host) {
    if (!ctx || n < 0 || d < 1 || k < 1 || !dcenters || (n > 0 && (!X || !dX_out || !dlabels_out)))
        return set_error(B2K_ERR_INVALID_ARG, "stage_assign: bad arguments (n=%lld d=%d k=%d)", (long long)n, d, k);
    B2K_TRY(check_metric_dim(metric, d));
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (n == 0) return B2K_OK;
    return stream_assign(ctx, X, n, d, dcenters, k, metric, labels_host, lloyd, dX_out, dlabels_out);
}

// ---- Lloyd session ----------------------------------------------------------------------------
struct b2k_lloyd;
static double lloyd_scale_sum(const b2k_lloyd* s);

struct b2k_lloyd {
    b2k_ctx* ctx = nullptr;
    const float* dX = nullptr;
    int64_t n = 0, n_total = 0;
    int d = 0, k = 0, metric = 0;
    int q_sum = 0, q_cost = 0;
    double scale_sum = 1, scale_cost = 1;
    DevMem l, Ga;
    PreparedCenters pc;
    ScreenPlan* plan = nullptr;
    bool plan_pending = false;  // the screen plan is built at the first step: its operand term count is chosen with the centers
    // exact center pruning (prune.cu): after the first step the session works on a copy of the frames sorted by label
    PruneState* prune = nullptr;
    bool prune_wanted = false;
    bool have_labels = false;   // prune_labels() holds the labels of the last step (in the current frame order)
    int64_t steps = 0, next_sort = 1;
    int64_t last_sort_step = 0;
    double mean_after_sort = 0, mean_last = 0;  // list length right after the last sort / at the last step
    // labels of the last step when the caller did not ask for them (dlabels == NULL): the loop itself never needs them in
    // the caller's frame order (deeptime's cluster_loop returns centers only); b2k_dev_lloyd_get_labels hands them out
    DevMem own_labels;
    bool labels_in_own = false, labels_in_prune = false;
    // incremental member sums of the pruned steps (lloyd.cu: accumulate_delta_kernel): the labels of the previous step in
    // the current frame order, the exact integer sums/counts that belong to them, and the number of frames whose label
    // changed in the last step (device word copied to a pinned host word; read after the next step's first sync)
    DevMem prev_labels, acc_state, d_changed;
    unsigned long long* h_changed = nullptr;  // pinned
    bool acc_valid = false, changed_known = false;
    ~b2k_lloyd() { if (h_changed) cudaFreeHost(h_changed); }
};

static int ceil_log2_d(double v) {
    if (!(v > 0)) return 0;
    int e;
    const double m = std::frexp(v, &e);  // v = m * 2^e, m in [0.5,1)
    return (m == 0.5) ? e - 1 : e;
}

// fixed-point exponents of the exchange buffer: |x| <= 2^eM, n_total <= 2^eN  =>  |sum| * 2^q < 2^62
static void lloyd_scales(float absmax_global, int64_t n_total, int d, int* q_sum, int* q_cost) {
    const int eM = ceil_log2_d((double)absmax_global);
    const int eN = ceil_log2_d((double)std::max<int64_t>(n_total, 1)) + 1;
    *q_sum = 62 - eN - eM;
    // cost terms l^2 <= d * (2M)^2 (1+eps)
    *q_cost = 61 - eN - (2 * (eM + 1) + ceil_log2_d((double)d));
}

B2K_API int b2k_dev_lloyd_create(b2k_ctx* ctx, const float* dX, int64_t n_local, int32_t d, int32_t k, int metric,
                                 int64_t n_total, float absmax_global, b2k_lloyd** out) {
    if (!ctx || !out || n_local < 0 || d < 1 || k < 1 || n_total < n_local)
        return set_error(B2K_ERR_INVALID_ARG, "lloyd_create: bad arguments");
    const bool staged_only = n_local > 0 && !dX;  // out-of-core session: only b2k_stage_lloyd_pass / finalize / decode_cost
    if (!std::isfinite(absmax_global))
        return set_error(B2K_ERR_NONFINITE, "lloyd_create: data contains NaN or inf");
    B2K_TRY(check_metric_dim(metric, d));
    CUDA_TRY(cudaSetDevice(ctx->device));
    b2k_lloyd* s = new b2k_lloyd();
    s->ctx = ctx; s->dX = dX; s->n = n_local; s->n_total = n_total; s->d = d; s->k = k; s->metric = metric;
    lloyd_scales(absmax_global, n_total, d, &s->q_sum, &s->q_cost);
    s->scale_sum = std::ldexp(1.0, s->q_sum);
    s->scale_cost = std::ldexp(1.0, s->q_cost);
    int rc = staged_only ? B2K_OK : s->l.alloc((size_t)std::max<int64_t>(n_local, 1) * 4);
    if (rc == B2K_OK && metric == B2K_METRIC_MINRMSD && !staged_only) {
        rc = s->Ga.alloc((size_t)std::max<int64_t>(n_local, 1) * 4);
        if (rc == B2K_OK) rc = launch_rmsd_center(ctx, dX, n_local, d, nullptr, s->Ga.as<float>());
    }
    if (rc == B2K_OK && !staged_only && metric == B2K_METRIC_EUCLIDEAN && ctx->engine != B2K_ENGINE_DIRECT &&
        screen_supported(ctx, d, k, n_local)) {
        s->plan_pending = true;
        s->prune_wanted = prune_supported(ctx, n_local, d, k);
    }
    if (rc != B2K_OK) { b2k_dev_lloyd_destroy(s); return rc; }
    *out = s;
    return B2K_OK;
}

B2K_API int b2k_dev_lloyd_destroy(b2k_lloyd* s) {
    if (!s) return B2K_OK;
    cudaStreamSynchronize(s->ctx->stream);
    if (s->plan && s->ctx->stat_plan == s->plan) {  // keep the numbers of the last step readable
        if (s->ctx->stat_pending)
            screen_read_stats(s->plan, &s->ctx->stat_cand_chunks, &s->ctx->stat_fallback_frames);
        s->ctx->stat_pending = false;
        s->ctx->stat_plan = nullptr;
    }
    if (s->plan) screen_plan_destroy(s->plan);
    if (s->prune) prune_destroy(s->prune);
    delete s;
    return B2K_OK;
}

static double lloyd_scale_sum(const b2k_lloyd* s) { return s->scale_sum; }

B2K_API int b2k_stage_lloyd_assign_accumulate(b2k_lloyd* s, const float* X, const float* dcenters, float* dX_out,
                                              int32_t* dlabels_out, int32_t* labels_host, int64_t* dacc) {
    if (!s || !dcenters || !dacc || (s->n > 0 && (!X || !dX_out || !dlabels_out)))
        return set_error(B2K_ERR_INVALID_ARG, "stage_lloyd_assign_accumulate: null argument");
    if (dX_out != s->dX) return set_error(B2K_ERR_INVALID_ARG, "stage_lloyd_assign_accumulate: dX_out must be the session's frame array");
    b2k_ctx* ctx = s->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaMemsetAsync(dacc, 0, (size_t)((int64_t)s->k * s->d + s->k + 1) * 8, ctx->stream));
    if (s->n == 0) return B2K_OK;
    // the frame array is rewritten: whatever the session derived from its old content (sorted copy, fp16 operand) is stale
    if (s->prune) { prune_destroy(s->prune); s->prune = nullptr; }
    s->acc_valid = s->changed_known = false;
    s->have_labels = false;
    s->labels_in_own = s->labels_in_prune = false;
    s->steps = 0;
    s->next_sort = 1;
    if (s->plan) screen_plan_invalidate_frames(s->plan);
    return stream_assign(ctx, X, s->n, s->d, dcenters, s->k, s->metric, labels_host, 1, dX_out, dlabels_out,
                         lloyd_scale_sum(s), dacc);
}

// Out-of-core Lloyd pass (the tier below HBM; the reference spills to a host memmap, kmeans.py:181-200): the session was
// created with dX = NULL, the frames live in (pinned) host memory and only pass through the two chunk slots.  One pass
// per iteration: for every chunk (a) if have_prev, the cost of the PREVIOUS labels (dlabels_io on entry) against
// dcenters -- i.e. the cost of the iteration that produced dcenters -- goes to the cost slot, (b) the chunk is assigned
// (dlabels_io on exit, labels_host if given), (c) its member sums and counts are added to acc.  All sums are exact
// integers: acc and cost are bit-identical to the resident session's.
B2K_API int b2k_stage_lloyd_pass(b2k_lloyd* s, const float* X, const float* dcenters, int32_t* dlabels_io, int have_prev,
                                 int32_t* labels_host, int64_t* dacc) {
    if (!s || !dcenters || !dacc || (s->n > 0 && (!X || !dlabels_io)))
        return set_error(B2K_ERR_INVALID_ARG, "stage_lloyd_pass: null argument");
    b2k_ctx* ctx = s->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int64_t len = (int64_t)s->k * s->d + s->k + 1;
    CUDA_TRY(cudaMemsetAsync(dacc, 0, (size_t)len * 8, ctx->stream));
    if (s->n == 0) return B2K_OK;
    return stream_assign(ctx, X, s->n, s->d, dcenters, s->k, s->metric, labels_host, 1, nullptr, dlabels_io, s->scale_sum,
                         dacc, have_prev ? dlabels_io : nullptr, s->scale_cost, have_prev ? dacc + (len - 1) : nullptr);
}

B2K_API int64_t b2k_dev_lloyd_acc_len(const b2k_lloyd* s) { return s ? (int64_t)s->k * s->d + s->k + 1 : 0; }

B2K_API int b2k_dev_lloyd_assign_accumulate(b2k_lloyd* s, const float* dC, int32_t* dlabels, int64_t* dacc) {
    if (!s || !dC || !dacc) return set_error(B2K_ERR_INVALID_ARG, "lloyd step: null argument");
    const bool want_labels = dlabels != nullptr;
    s->labels_in_own = s->labels_in_prune = false;
    if (!dlabels && s->n > 0) {  // the caller does not want the labels: keep them here
        B2K_TRY(s->own_labels.alloc((size_t)s->n * 4));
        dlabels = s->own_labels.as<int32_t>();
    }
    if (s->n > 0 && !s->dX) return set_error(B2K_ERR_INVALID_ARG, "lloyd step: out-of-core session (use b2k_stage_lloyd_pass)");
    b2k_ctx* ctx = s->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaMemsetAsync(dacc, 0, (size_t)b2k_dev_lloyd_acc_len(s) * 8, ctx->stream));
    if (s->n == 0) return B2K_OK;
    if (s->plan_pending) {
        s->plan_pending = false;
        int terms = 0;
        B2K_TRY(screen_choose_terms(ctx, s->dX, s->n, s->d, dC, s->k, &terms));
        int rc = screen_plan_create(ctx, s->n, s->d, s->k, terms, &s->plan);
        if (rc == B2K_ERR_NOMEM) {
            // the fp16 screen operand (Kp*2 + ~41 bytes per frame) does not fit next to the frames: the session runs
            // on the exact CUDA-core engine instead of failing (the reference would still run, kmeans.py:181-200)
            s->plan = nullptr;
            rc = B2K_OK;
        }
        if (rc == B2K_OK && s->plan) rc = screen_prepare_frames(s->plan, s->dX, s->n);
        if (rc != B2K_OK) return rc;
    }
    if (s->plan && s->prune_wanted) {
        // (re)sort by the labels of the previous step when the schedule says so: steps 1, 2, 4, 8, ... or every prune_resort
        // Re-sort policy (prune_resort = 0): after iterations 1 and 2 (the labels still change a lot), then whenever the lists
        // have grown by a quarter since the last sort -- a sort (label sort, row gather, operand rebuild, tile balls) costs
        // about one and a half iterations, lists 25 % longer cost every iteration about 15 % -- and at iterations 4, 16, 64...
        // at the latest.  prune_resort = n > 0: every n iterations.
        bool resort = s->prune && s->have_labels && s->steps >= s->next_sort;
        if (s->prune && s->have_labels && !resort && ctx->prune_resort == 0 && prune_sorted(s->prune) &&
            s->steps >= s->last_sort_step + 2 && s->mean_after_sort > 0 && s->mean_last > 1.25 * s->mean_after_sort)
            resort = true;
        if (resort) {
            B2K_TRY(prune_sort(s->prune, s->dX, prune_labels(s->prune), dC));
            screen_plan_invalidate_frames(s->plan);
            s->next_sort = ctx->prune_resort > 0 ? s->steps + ctx->prune_resort : (s->steps < 4 ? s->steps * 2 : s->steps * 4);
            s->last_sort_step = s->steps;
            s->mean_after_sort = 0;
            ctx->stat_prune_sorts += 1;
        }
        if (s->prune && prune_sorted(s->prune)) {
            PruneState* pr = s->prune;
            double mean = 0;
            int mx = 0, ov = 0;
            B2K_TRY(prune_lists(pr, dC, &mean, &mx, &ov));
            ctx->stat_prune_mean = mean;
            if (s->mean_after_sort == 0) s->mean_after_sort = mean;
            s->mean_last = mean;
            // (prune_lists synchronised the stream: the changed-label count of the previous step has arrived)
            bool keep_prev = ctx->delta_sums != 0, use_delta = false;
            if (keep_prev) {
                if (s->prev_labels.alloc((size_t)s->n * 4) != B2K_OK || s->acc_state.alloc((size_t)b2k_dev_lloyd_acc_len(s) * 8) != B2K_OK ||
                    s->d_changed.alloc(8) != B2K_OK ||
                    (!s->h_changed && cudaHostAlloc((void**)&s->h_changed, 64, cudaHostAllocDefault) != cudaSuccess)) {
                    cudaGetLastError();
                    keep_prev = false;  // no room: full passes
                    s->acc_valid = false;
                }
            }
            if (keep_prev) {
                if (s->changed_known) ctx->stat_changed = (double)*s->h_changed;
                use_delta = s->acc_valid && (ctx->delta_sums == 2 || (s->changed_known && *s->h_changed * 8 <= (unsigned long long)s->n));
                // the labels the kept sums belong to, in the current (possibly just re-sorted) frame order
                CUDA_TRY(cudaMemcpyAsync(s->prev_labels.p, prune_labels(pr), (size_t)s->n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
            } else {
                s->acc_valid = false;
            }
            // the listed screen drains mean (padded) columns per frame, the full one k rounded up to 256
            if (ov == 0 && (mean <= 0.6 * (double)(cdiv(s->k, 256) * 256) || ctx->prune_mode == 3)) {
                B2K_TRY(screen_assign_listed(s->plan, prune_frames(pr), s->n, dC, prune_tlist(pr), prune_tcount(pr),
                                             prune_lcap(pr), prune_unit_shift(pr), prune_labels(pr), 1));
                ctx->stat_prune_steps += 1;
            } else {  // the lists would not pay: every center for every tile, still on the sorted frames
                B2K_TRY(screen_assign(s->plan, prune_frames(pr), s->n, dC, prune_labels(pr), nullptr, 1));
            }
            ctx->stat_screen_frames = (double)s->n;
            ctx->stat_screen_terms = (double)screen_plan_terms(s->plan);
            ctx->stat_plan = s->plan;
            ctx->stat_pending = true;
            s->have_labels = true;
            s->steps += 1;
            s->labels_in_prune = true;
            if (want_labels) B2K_TRY(prune_scatter_labels(pr, dlabels));  // back to the caller's frame order
            if (!use_delta && !keep_prev)
                return launch_accumulate(ctx, prune_frames(pr), s->n, s->d, s->k, prune_labels(pr), s->scale_sum, dacc);
            // member sums: incremental when few labels changed last time, else a full pass whose result is kept
            const size_t sums_bytes = (size_t)((int64_t)s->k * s->d + s->k) * 8;
            unsigned long long* dch = s->d_changed.as<unsigned long long>();
            CUDA_TRY(cudaMemsetAsync(dch, 0, 8, ctx->stream));
            if (use_delta) {
                B2K_TRY(launch_accumulate_delta(ctx, prune_frames(pr), s->n, s->d, s->k, s->prev_labels.as<int32_t>(),
                                                prune_labels(pr), s->scale_sum, s->acc_state.as<int64_t>(), dch));
                CUDA_TRY(cudaMemcpyAsync(dacc, s->acc_state.p, sums_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
                ctx->stat_delta_steps += 1;
            } else {
                B2K_TRY(launch_accumulate(ctx, prune_frames(pr), s->n, s->d, s->k, prune_labels(pr), s->scale_sum, dacc));
                CUDA_TRY(cudaMemcpyAsync(s->acc_state.p, dacc, sums_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
                B2K_TRY(launch_count_changed(ctx, s->prev_labels.as<int32_t>(), prune_labels(pr), s->n, dch));
                s->acc_valid = true;
            }
            CUDA_TRY(cudaMemcpyAsync(s->h_changed, dch, 8, cudaMemcpyDeviceToHost, ctx->stream));
            s->changed_known = true;
            return B2K_OK;
        }
    }
    if (s->plan) {
        B2K_TRY(screen_assign(s->plan, s->dX, s->n, dC, dlabels, nullptr, 1));
        s->steps += 1;
        if (s->prune_wanted && !s->prune) {
            // first step done: keep its labels; the frames are sorted by them at the start of the next step (a session
            // that stops after one iteration never pays for the sort)
            if (prune_create(ctx, s->n, s->d, s->k, &s->prune) != B2K_OK) {
                cudaGetLastError();
                s->prune = nullptr;
                s->prune_wanted = false;  // no room for the sorted copy: the session stays on the unsorted path
            }
        }
        if (s->prune && !prune_sorted(s->prune)) {
            CUDA_TRY(cudaMemcpyAsync(prune_labels(s->prune), dlabels, (size_t)s->n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
            s->have_labels = true;
        }
        ctx->stat_screen_frames = (double)s->n;
        ctx->stat_screen_terms = (double)screen_plan_terms(s->plan);
        ctx->stat_plan = s->plan;
        ctx->stat_pending = true;
    } else {
        B2K_TRY(s->pc.prepare(ctx, dC, s->k, s->d, s->metric));
        B2K_TRY(assign_any(ctx, s->dX, s->Ga.as<float>(), s->n, s->d, s->pc, s->k, s->metric, dlabels, nullptr, 1));
    }
    s->labels_in_own = !want_labels;
    return launch_accumulate(ctx, s->dX, s->n, s->d, s->k, dlabels, s->scale_sum, dacc);
}

// labels of the session's last step in the caller's frame order (for callers that passed dlabels = NULL to the step)
B2K_API int b2k_dev_lloyd_get_labels(b2k_lloyd* s, int32_t* dlabels_out) {
    if (!s || (s->n > 0 && !dlabels_out)) return set_error(B2K_ERR_INVALID_ARG, "lloyd get_labels: null argument");
    if (s->n == 0) return B2K_OK;
    CUDA_TRY(cudaSetDevice(s->ctx->device));
    if (s->labels_in_prune && s->prune) return prune_scatter_labels(s->prune, dlabels_out);
    if (s->labels_in_own) {
        CUDA_TRY(cudaMemcpyAsync(dlabels_out, s->own_labels.p, (size_t)s->n * 4, cudaMemcpyDeviceToDevice, s->ctx->stream));
        return B2K_OK;
    }
    return set_error(B2K_ERR_INVALID_ARG, "lloyd get_labels: the last step wrote its labels to the caller's array");
}

B2K_API int b2k_dev_lloyd_accumulate(b2k_lloyd* s, const int32_t* dlabels, int64_t* dacc) {
    if (!s || !dacc || (s->n > 0 && !dlabels)) return set_error(B2K_ERR_INVALID_ARG, "lloyd accumulate: null argument");
    if (s->n > 0 && !s->dX) return set_error(B2K_ERR_INVALID_ARG, "lloyd accumulate: out-of-core session");
    b2k_ctx* ctx = s->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaMemsetAsync(dacc, 0, (size_t)b2k_dev_lloyd_acc_len(s) * 8, ctx->stream));
    if (s->n == 0) return B2K_OK;
    return launch_accumulate(ctx, s->dX, s->n, s->d, s->k, dlabels, s->scale_sum, dacc);
}

B2K_API int b2k_dev_lloyd_finalize(b2k_lloyd* s, const int64_t* dacc, const float* dC_old, float* dC_new) {
    if (!s || !dacc || !dC_old || !dC_new) return set_error(B2K_ERR_INVALID_ARG, "lloyd finalize: null argument");
    CUDA_TRY(cudaSetDevice(s->ctx->device));
    return launch_finalize(s->ctx, dacc, s->k, s->d, std::ldexp(1.0, -s->q_sum), dC_old, dC_new);
}

B2K_API int b2k_dev_lloyd_cost(b2k_lloyd* s, const float* dC_new, const int32_t* dlabels, int64_t* dacc) {
    if (!s || !dC_new || !dacc) return set_error(B2K_ERR_INVALID_ARG, "lloyd cost: null argument");
    if (s->n > 0 && !s->dX) return set_error(B2K_ERR_INVALID_ARG, "lloyd cost: out-of-core session (b2k_stage_lloyd_pass measures it)");
    b2k_ctx* ctx = s->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    int64_t* slot = dacc + (int64_t)s->k * s->d + s->k;
    CUDA_TRY(cudaMemsetAsync(slot, 0, 8, ctx->stream));
    if (s->n == 0) return B2K_OK;
    ProfScope prof(ctx, b2k_ctx::PROF_COST);
    const float* fX = s->dX;
    if (!dlabels && s->labels_in_own) dlabels = s->own_labels.as<int32_t>();
    if (!dlabels && !(s->prune && prune_sorted(s->prune) && s->labels_in_prune))
        return set_error(B2K_ERR_INVALID_ARG, "lloyd cost: no labels (pass them, or run b2k_dev_lloyd_assign_accumulate first)");
    if (s->prune && prune_sorted(s->prune) && s->have_labels) {
        // the session works on its sorted copy of the frames: the labels of the last step are there in the same order
        // (the cost is an exact integer sum, so the order of the frames does not change a bit of it)
        fX = prune_frames(s->prune);
        dlabels = prune_labels(s->prune);
    }
    if (s->metric == B2K_METRIC_MINRMSD) {
        B2K_TRY(s->pc.prepare(ctx, dC_new, s->k, s->d, s->metric));
        B2K_TRY(launch_rmsd_labeled_dist(ctx, s->dX, s->Ga.as<float>(), s->n, s->d, s->pc.C, s->pc.Gb, dlabels,
                                         s->l.as<float>()));
    } else {
        int fused = 0;  // narrow rows: distances and the integer cost sum in one pass
        B2K_TRY(launch_cost_fused(ctx, fX, s->n, s->d, dC_new, s->k, dlabels, s->scale_cost, slot, &fused));
        if (fused) return B2K_OK;
        B2K_TRY(launch_labeled_dist(ctx, fX, s->n, s->d, dC_new, dlabels, s->l.as<float>()));
    }
    return launch_cost_reduce(ctx, s->l.as<float>(), s->n, s->scale_cost, slot);
}

B2K_API double b2k_dev_lloyd_decode_cost(const b2k_lloyd* s, int64_t cost_fixed) {
    return s ? std::ldexp((double)cost_fixed, -s->q_cost) : 0.0;
}

B2K_API int b2k_dev_absmax(b2k_ctx* ctx, const float* dX, int64_t count, float* out_host) {
    if (!ctx || !out_host) return set_error(B2K_ERR_INVALID_ARG, "absmax: null argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaMemsetAsync(ctx->flags, 0, 4, ctx->stream));
    B2K_TRY(launch_absmax(ctx, dX, count, (float*)ctx->flags));
    CUDA_TRY(cudaMemcpyAsync(out_host, ctx->flags, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return B2K_OK;
}

B2K_API int b2k_dev_all_finite(b2k_ctx* ctx, const float* dX, int64_t count, int* out_host) {
    if (!ctx || !out_host) return set_error(B2K_ERR_INVALID_ARG, "all_finite: null argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int one = 1;
    CUDA_TRY(cudaMemcpyAsync(ctx->flags, &one, 4, cudaMemcpyHostToDevice, ctx->stream));
    B2K_TRY(launch_all_finite(ctx, dX, count, ctx->flags));
    CUDA_TRY(cudaMemcpyAsync(out_host, ctx->flags, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return B2K_OK;
}

// single-GPU cluster_loop over device-resident frames (the per-iteration exchange is a no-op)
static int dev_cluster_loop(b2k_ctx* ctx, const float* dX, int64_t n, int d, float* dC_io, int k, int metric,
                            int max_iter, float tol, b2k_callback cb, void* user, int* code, int* iters,
                            float* inertias, int cap, int32_t* dlabels_opt) {
    float absmax = 0.f;
    B2K_TRY(b2k_dev_absmax(ctx, dX, n * d, &absmax));
    b2k_lloyd* s = nullptr;
    B2K_TRY(b2k_dev_lloyd_create(ctx, dX, n, d, k, metric, n, absmax, &s));
    DevMem acc, cnew, lab;
    int rc = acc.alloc((size_t)b2k_dev_lloyd_acc_len(s) * 8);
    if (rc == B2K_OK) rc = cnew.alloc((size_t)k * d * 4);
    if (rc == B2K_OK && !dlabels_opt) rc = lab.alloc((size_t)std::max<int64_t>(n, 1) * 4);
    int32_t* dl = dlabels_opt ? dlabels_opt : lab.as<int32_t>();
    float* cur = dC_io;
    float* nxt = cnew.as<float>();
    int it = 0;
    bool converged = false;
    float prev = 0.f;
    while (rc == B2K_OK) {
        rc = b2k_dev_lloyd_assign_accumulate(s, cur, dl, acc.as<int64_t>());
        if (rc == B2K_OK) rc = b2k_dev_lloyd_finalize(s, acc.as<int64_t>(), cur, nxt);
        if (rc == B2K_OK) rc = b2k_dev_lloyd_cost(s, nxt, dl, acc.as<int64_t>());
        if (rc != B2K_OK) break;
        int64_t cf = 0;
        cudaMemcpyAsync(&cf, acc.as<int64_t>() + (int64_t)k * d + k, 8, cudaMemcpyDeviceToHost, ctx->stream);
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { rc = set_error(B2K_ERR_CUDA, "cluster_loop: %s", cudaGetErrorString(e)); break; }
        std::swap(cur, nxt);
        const float cost = (float)b2k_dev_lloyd_decode_cost(s, cf);
        if (it < cap && inertias) inertias[it] = cost;
        const float rel = (cost != 0.0f) ? std::fabs(cost - prev) / cost : 0.f;
        prev = cost;
        if (rel <= tol) converged = true;
        else if (cb) cb(user);
        it += 1;
        if (!(it < max_iter && !converged)) break;
    }
    if (rc == B2K_OK && cur != dC_io) {
        cudaMemcpyAsync(dC_io, cur, (size_t)k * d * 4, cudaMemcpyDeviceToDevice, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
    }
    b2k_dev_lloyd_destroy(s);
    if (rc != B2K_OK) return rc;
    *code = converged ? 0 : 1;
    *iters = it;
    return B2K_OK;
}

B2K_API int b2k_dev_kmeans_cluster_loop(b2k_ctx* ctx, const float* dX, int64_t n, int32_t d, float* dC_io, int32_t k,
                                        int metric, int32_t max_iter, float tol, b2k_callback cb, void* user,
                                        int* code, int* iters, float* inertias, int32_t cap, int32_t* dlabels_opt) {
    if (!ctx || !dX || !dC_io || !code || !iters || n < 1 || d < 1 || k < 1)
        return set_error(B2K_ERR_INVALID_ARG, "cluster_loop: bad arguments");
    B2K_TRY(check_metric_dim(metric, d));
    CUDA_TRY(cudaSetDevice(ctx->device));
    return dev_cluster_loop(ctx, dX, n, d, dC_io, k, metric, max_iter, tol, cb, user, code, iters, inertias, cap,
                            dlabels_opt);
}

// ---- host-pointer k-means entry points ----------------------------------------------------------
struct HostFrames {  // frames of a host array made resident in HBM
    DevMem mem;
    int load(b2k_ctx* ctx, const float* X, int64_t n, int d) {
        B2K_TRY(mem.alloc((size_t)n * d * 4));
        return upload_host(ctx, X, mem.p, (size_t)n * d * 4);
    }
};

// One Lloyd step from host frames.  The frames stream into HBM chunk by chunk and every chunk is assigned
// while the next one is still on the bus (stream_assign); the member sums only need the labels and one exact
// data-range bound, so they run once over the resident array at the end.
B2K_API int b2k_kmeans_cluster(b2k_ctx* ctx, const float* X, int64_t n, int32_t d, const float* centers, int32_t k,
                               int metric, float* new_centers, int32_t* labels) {
    if (!ctx || !X || !centers || !new_centers || !labels || n < 1 || d < 1 || k < 1)
        return set_error(B2K_ERR_INVALID_ARG, "kmeans_cluster: bad arguments");
    B2K_TRY(check_metric_dim(metric, d));
    CUDA_TRY(cudaSetDevice(ctx->device));
    float *dXf, *dC, *dN;
    int32_t* dL;
    int64_t* acc;
    B2K_TRY(ctx->slot(b2k_ctx::SLOT_FRAMES, (size_t)n * d * 4, (void**)&dXf));
    B2K_TRY(ctx->slot(b2k_ctx::SLOT_CENTERS, (size_t)k * d * 4, (void**)&dC));
    B2K_TRY(ctx->slot(b2k_ctx::SLOT_CENTERS2, (size_t)k * d * 4, (void**)&dN));
    B2K_TRY(ctx->slot(b2k_ctx::SLOT_LABELS, (size_t)n * 4, (void**)&dL));
    B2K_TRY(ctx->slot(b2k_ctx::SLOT_ACC, ((size_t)k * d + k + 1) * 8, (void**)&acc));
    CUDA_TRY(cudaMemcpyAsync(dC, centers, (size_t)k * d * 4, cudaMemcpyHostToDevice, ctx->stream));
    B2K_TRY(stream_assign(ctx, X, n, d, dC, k, metric, labels, 1, dXf, dL));
    float absmax = 0.f;
    B2K_TRY(b2k_dev_absmax(ctx, dXf, n * d, &absmax));
    if (!std::isfinite(absmax)) return set_error(B2K_ERR_NONFINITE, "kmeans_cluster: data contains NaN or inf");
    int q_sum = 0, q_cost = 0;
    lloyd_scales(absmax, n, d, &q_sum, &q_cost);
    const int64_t acc_len = (int64_t)k * d + k + 1;
    CUDA_TRY(cudaMemsetAsync(acc, 0, (size_t)acc_len * 8, ctx->stream));
    B2K_TRY(launch_accumulate(ctx, dXf, n, d, k, dL, std::ldexp(1.0, q_sum), acc));
    B2K_TRY(launch_finalize(ctx, acc, k, d, std::ldexp(1.0, -q_sum), dC, dN));
    CUDA_TRY(cudaMemcpyAsync(new_centers, dN, (size_t)k * d * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return B2K_OK;
}

B2K_API int b2k_kmeans_cost(b2k_ctx* ctx, const float* X, int64_t n, int32_t d, const float* centers, int32_t k,
                            const int32_t* labels, int metric, float* cost) {
    if (!ctx || !X || !centers || !labels || !cost || n < 1 || d < 1 || k < 1)
        return set_error(B2K_ERR_INVALID_ARG, "kmeans_cost: bad arguments");
    B2K_TRY(check_metric_dim(metric, d));
    CUDA_TRY(cudaSetDevice(ctx->device));
    HostFrames F;
    B2K_TRY(F.load(ctx, X, n, d));
    DevMem dC, dL, acc;
    B2K_TRY(dC.alloc((size_t)k * d * 4));
    B2K_TRY(dL.alloc((size_t)n * 4));
    CUDA_TRY(cudaMemcpyAsync(dC.p, centers, (size_t)k * d * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(dL.p, labels, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    float absmax = 0.f;
    B2K_TRY(b2k_dev_absmax(ctx, F.mem.as<float>(), n * d, &absmax));
    // the cost fixed-point scale must also cover the centers
    float cmax = 0.f;
    B2K_TRY(b2k_dev_absmax(ctx, dC.as<float>(), (int64_t)k * d, &cmax));
    b2k_lloyd* s = nullptr;
    const int saved_engine = ctx->engine;
    ctx->engine = B2K_ENGINE_DIRECT;  // no screen operands needed for a cost evaluation
    int rc = b2k_dev_lloyd_create(ctx, F.mem.as<float>(), n, d, k, metric, n, std::max(absmax, cmax), &s);
    ctx->engine = saved_engine;
    if (rc != B2K_OK) return rc;
    rc = acc.alloc((size_t)b2k_dev_lloyd_acc_len(s) * 8);
    if (rc == B2K_OK) rc = b2k_dev_lloyd_cost(s, dC.as<float>(), dL.as<int32_t>(), acc.as<int64_t>());
    if (rc == B2K_OK) {
        int64_t cf = 0;
        cudaMemcpyAsync(&cf, acc.as<int64_t>() + (int64_t)k * d + k, 8, cudaMemcpyDeviceToHost, ctx->stream);
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = set_error(B2K_ERR_CUDA, "kmeans_cost: %s", cudaGetErrorString(e));
        else *cost = (float)b2k_dev_lloyd_decode_cost(s, cf);
    }
    b2k_dev_lloyd_destroy(s);
    return rc;
}

B2K_API int b2k_kmeans_cluster_loop(b2k_ctx* ctx, const float* X, int64_t n, int32_t d, float* centers_io, int32_t k,
                                    int metric, int32_t max_iter, float tolerance, b2k_callback cb, void* user,
                                    int* code, int* iters, float* inertias, int32_t inertias_cap) {
    if (!ctx || !X || !centers_io || !code || !iters || n < 1 || d < 1 || k < 1)
        return set_error(B2K_ERR_INVALID_ARG, "cluster_loop: bad arguments");
    B2K_TRY(check_metric_dim(metric, d));
    CUDA_TRY(cudaSetDevice(ctx->device));
    HostFrames F;
    B2K_TRY(F.load(ctx, X, n, d));
    DevMem dC;
    B2K_TRY(dC.alloc((size_t)k * d * 4));
    CUDA_TRY(cudaMemcpyAsync(dC.p, centers_io, (size_t)k * d * 4, cudaMemcpyHostToDevice, ctx->stream));
    B2K_TRY(dev_cluster_loop(ctx, F.mem.as<float>(), n, d, dC.as<float>(), k, metric, max_iter, tolerance, cb, user,
                             code, iters, inertias, inertias_cap, nullptr));
    CUDA_TRY(cudaMemcpyAsync(centers_io, dC.p, (size_t)k * d * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return B2K_OK;
}

B2K_API int b2k_dev_kmeans_init_centers_kmpp(b2k_ctx* ctx, const float* dX, int64_t n, int32_t d, int32_t k,
                                             int metric, int64_t seed, int scan_mode, b2k_callback cb, void* user,
                                             float* dcenters_out, int64_t* chosen_host) {
    if (!ctx || !dX || !dcenters_out || n < 1 || d < 1)
        return set_error(B2K_ERR_INVALID_ARG, "init_centers_kmpp: bad arguments");
    B2K_TRY(check_metric_dim(metric, d));
    CUDA_TRY(cudaSetDevice(ctx->device));
    return kmpp_run(ctx, dX, n, d, k, metric, seed, scan_mode, cb, user, dcenters_out, chosen_host);
}

B2K_API int64_t b2k_kmpp_exchange_floats(int64_t n_total, int32_t d, int32_t k) {
    if (n_total < 1 || d < 1 || k < 1) return 0;
    const int64_t m = 2 + (int64_t)std::log((double)k);
    return std::max<int64_t>(m * cdiv(n_total, 1024), m * (int64_t)d);
}

B2K_API int b2k_dev_kmeans_init_centers_kmpp_sharded(b2k_ctx* ctx, const float* dX, int64_t n_local, int32_t d,
                                                     int32_t k, int metric, int64_t seed, int64_t global_lo,
                                                     int64_t n_total, float* xchg_f32, int64_t xchg_f32_len,
                                                     int64_t* xchg_i64, b2k_exchange_fn exchange, void* exchange_user,
                                                     b2k_callback cb, void* user, float* dcenters_out,
                                                     int64_t* chosen_host) {
    if (!ctx || !dcenters_out || n_local < 0 || d < 1 || n_total < 1 || global_lo < 0 || global_lo + n_local > n_total ||
        (n_local > 0 && !dX) || !exchange || !xchg_f32 || !xchg_i64)
        return set_error(B2K_ERR_INVALID_ARG, "init_centers_kmpp_sharded: bad arguments");
    B2K_TRY(check_metric_dim(metric, d));
    CUDA_TRY(cudaSetDevice(ctx->device));
    return kmpp_run_blocked(ctx, dX, n_local, d, k, metric, seed, global_lo, n_total, xchg_f32, xchg_f32_len, xchg_i64,
                            exchange, exchange_user, cb, user, dcenters_out, chosen_host);
}

B2K_API int b2k_kmeans_init_centers_kmpp(b2k_ctx* ctx, const float* X, int64_t n, int32_t d, int32_t k, int metric,
                                         int64_t seed, int scan_mode, b2k_callback cb, void* user,
                                         float* centers_out, int64_t* chosen_or_null) {
    if (!ctx || !X || !centers_out || n < 1 || d < 1)
        return set_error(B2K_ERR_INVALID_ARG, "init_centers_kmpp: bad arguments");
    if (k < 1 || k > n) return set_error(B2K_ERR_INVALID_ARG, "k-means++: need 1 <= k <= n (k=%d, n=%lld)", k, (long long)n);
    B2K_TRY(check_metric_dim(metric, d));
    CUDA_TRY(cudaSetDevice(ctx->device));
    HostFrames F;
    B2K_TRY(F.load(ctx, X, n, d));
    DevMem dC;
    B2K_TRY(dC.alloc((size_t)k * d * 4));
    B2K_TRY(kmpp_run(ctx, F.mem.as<float>(), n, d, k, metric, seed, scan_mode, cb, user, dC.as<float>(), chosen_or_null));
    CUDA_TRY(cudaMemcpyAsync(centers_out, dC.p, (size_t)k * d * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return B2K_OK;
}

