// project.cu -- the step BEFORE the clustering path: the linear projection of TICA / PCA,
//   Y = (X - mean) . W[:, :dout]       pyemma/coordinates/transform/_tica_base.py:130-133, pca.py:263-265
// fused into the chunk hand-off so that the raw features never become resident: a host chunk is staged through
// pinned memory, projected on the device and only the (n, dout) result stays in HBM (it is what k-means gathers,
// kmeans.py:326-338).  The reference works in fp64 (mean and eigenvectors are fp64) and casts the result to fp32
// (`Y.astype(self.output_type())`); here every output is one thread's fp64 FMA chain over the input dimensions in
// index order, rounded once to fp32 -- within 1 ulp(fp32) of the reference (tests: 1e-6 relative to numpy fp64).
// HBM/PCIe bound: 4*din bytes in, 4*dout bytes out per frame; din*dout fp64 FMAs per frame.
#include "common.cuh"
#include "kernels.h"

namespace b2k {

// CTA = 256 threads: a tile of 256/DOUT_SLOTS frames x DOUT_SLOTS output columns; W (double) in shared memory,
// the frame tile is staged coalesced as floats (row stride din+1: conflict-free column walks).
__global__ void __launch_bounds__(256) project_kernel(const float* __restrict__ X, int64_t n, int din,
                                                      const double* __restrict__ mean, const double* __restrict__ W,
                                                      int ldw, int dout, float* __restrict__ Y, int frames_per_cta) {
    extern __shared__ __align__(16) unsigned char psm[];
    double* ws = reinterpret_cast<double*>(psm);                 // [din][dout]
    double* ms = ws + (size_t)din * dout;                        // [din]
    float* xs = reinterpret_cast<float*>(ms + din);              // [frames_per_cta][din+1]
    for (int t = threadIdx.x; t < din * dout; t += 256) {
        const int e = t / dout, j = t - e * dout;
        ws[t] = W[(size_t)e * ldw + j];
    }
    for (int t = threadIdx.x; t < din; t += 256) ms[t] = mean ? mean[t] : 0.0;
    const int xs_stride = din + 1;
    for (int64_t base = (int64_t)blockIdx.x * frames_per_cta; base < n; base += (int64_t)gridDim.x * frames_per_cta) {
        const int nf = (int)min((int64_t)frames_per_cta, n - base);
        __syncthreads();
        const float* src = X + base * din;
        for (int t = threadIdx.x; t < nf * din; t += 256) {
            const int r = t / din, c = t - r * din;
            xs[(size_t)r * xs_stride + c] = __ldg(src + t);
        }
        __syncthreads();
        for (int t = threadIdx.x; t < nf * dout; t += 256) {
            const int r = t / dout, j = t - r * dout;
            const float* xr = xs + (size_t)r * xs_stride;
            double acc = 0.0;
            for (int e = 0; e < din; ++e) acc = fma((double)xr[e] - ms[e], ws[(size_t)e * dout + j], acc);
            Y[(base + r) * dout + j] = (float)acc;
        }
    }
}

int launch_project(b2k_ctx* ctx, const float* X, int64_t n, int din, const double* mean, const double* W, int ldw,
                   int dout, float* Y) {
    if (n <= 0) return B2K_OK;
    const size_t wbytes = ((size_t)din * dout + din) * 8;
    if (wbytes > 120 * 1024) return set_error(B2K_ERR_INVALID_ARG, "project: din*dout too large for the shared-memory table");
    int fpc = (int)std::min<size_t>(256, (200 * 1024 - wbytes) / ((size_t)(din + 1) * 4));
    if (fpc < 1) return set_error(B2K_ERR_INVALID_ARG, "project: input dimension too large");
    const size_t smem = wbytes + (size_t)fpc * (din + 1) * 4;
    static PerDeviceOnce attr_set;
    if (attr_set.need(ctx->device)) {
        CUDA_TRY(cudaFuncSetAttribute(project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set.done(ctx->device);
    }
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv(n, fpc), (int64_t)ctx->sm_count * 2));
    project_kernel<<<grid, 256, smem, ctx->stream>>>(X, n, din, mean, W, ldw, dout, Y, fpc);
    LAUNCH_CHECK();
    return B2K_OK;
}

}  // namespace b2k

using namespace b2k;

B2K_API int b2k_dev_project(b2k_ctx* ctx, const float* dX, int64_t n, int32_t din, const double* dmean_or_null,
                            const double* dW, int32_t ldw, int32_t dout, float* dY) {
    if (!ctx || n < 0 || din < 1 || dout < 1 || ldw < dout || !dW || (n > 0 && (!dX || !dY)))
        return set_error(B2K_ERR_INVALID_ARG, "project: bad arguments");
    CUDA_TRY(cudaSetDevice(ctx->device));
    return launch_project(ctx, dX, n, din, dmean_or_null, dW, ldw, dout, dY);
}

// host frames -> pinned staging -> device chunk -> projected rows appended to dY (n x dout); the raw chunk buffers
// are the context's grow-only slots, two of them, so the copy of chunk c+1 overlaps the projection of chunk c
B2K_API int b2k_stage_project(b2k_ctx* ctx, const float* X, int64_t n, int32_t din, const double* dmean_or_null,
                              const double* dW, int32_t ldw, int32_t dout, float* dY) {
    if (!ctx || n < 0 || din < 1 || dout < 1 || ldw < dout || !dW || (n > 0 && (!X || !dY)))
        return set_error(B2K_ERR_INVALID_ARG, "stage_project: bad arguments");
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (n == 0) return B2K_OK;
    const int64_t row_bytes = (int64_t)din * 4;
    const int64_t cf = std::min<int64_t>(n, std::max<int64_t>(1, (int64_t)ctx->stage_bytes / row_bytes));
    float* dchunk[2];
    for (int s = 0; s < 2; ++s) B2K_TRY(ctx->slot(b2k_ctx::SLOT_CHUNK_X0 + s, (size_t)cf * row_bytes, (void**)&dchunk[s]));
    cudaEvent_t ev_k[2];
    for (int s = 0; s < 2; ++s) CUDA_TRY(cudaEventCreateWithFlags(&ev_k[s], cudaEventDisableTiming));
    int rc = B2K_OK;
    int c = 0;
    for (int64_t off = 0; off < n && rc == B2K_OK; off += cf, ++c) {
        const int s = c & 1;
        const int64_t len = std::min(cf, n - off);
        // the slot is free once the projection that read it two chunks ago has finished
        for (int t = 0; t < 2; ++t) cudaStreamWaitEvent(ctx->copy_stream[t], ev_k[s], 0);
        rc = upload_host(ctx, X + off * din, dchunk[s], (size_t)len * row_bytes);  // waits are enqueued on ctx->stream
        if (rc == B2K_OK) rc = launch_project(ctx, dchunk[s], len, din, dmean_or_null, dW, ldw, dout, dY + off * dout);
        cudaEventRecord(ev_k[s], ctx->stream);
    }
    cudaStreamSynchronize(ctx->stream);
    for (int s = 0; s < 2; ++s) cudaEventDestroy(ev_k[s]);
    return rc;
}
