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
    const int64_t total = n * d;
    if (total < (int64_t(1) << 31))
        accumulate_kernel32<<<grid_for(ctx, total, 256), 256, 0, ctx->stream>>>(
            X, (uint32_t)total, (uint32_t)d, k, labels, scale, (unsigned long long*)acc);
    else
        accumulate_kernel<<<grid_for(ctx, total, 256), 256, 0, ctx->stream>>>(X, n, d, k, labels, scale,
                                                                             (unsigned long long*)acc);
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
