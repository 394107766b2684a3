// rmsd.cu -- minRMSD metric: per-pair QCP kernel (K5 of SURVEY 2.2).
//
// Replaces RMSDMetric::compute_squared / compute
// (pyemma/coordinates/clustering/src/clustering_module.cpp:9-36), which per PAIR heap-allocates
// two copies, centers both with mdtraj's inplace_center_and_trace_atom_major and calls
// msd_atom_major (Theobald QCP).  Here centers are centered ONCE per call (bitwise the same
// buffer_b the reference recomputes per pair), frame traces G_a once per dataset, and -- exactly
// like clustering_module.cpp:28 -- the UNCENTERED frame is the `a` operand of the 3x3
// cross-covariance while G_a comes from its centered copy.
//
// Arithmetic contract (shared with oracle/oracle.cpp, "parity unpinned upstream"):
//   centroid, centering, traces: fp64, sequential over atoms; stored fp32
//   M = sum_atoms a b^T: fp32, atom t -> lane t%4, mul then add (no FMA), (l0+l1)+(l2+l3)
//   quartic coefficients + Newton (<=50 its, rel 1e-11) + msd: fp64, no FMA contraction
//   msd clamped at 0, cast to fp32, then sqrt (clustering_module.cpp:34)
#include "common.cuh"
#include "kernels.h"

namespace b2k {

#define DM(a, b) __dmul_rn((a), (b))
#define DA(a, b) __dadd_rn((a), (b))
#define DS(a, b) __dsub_rn((a), (b))

// `bound` (argmin callers; +inf: none): the frame's best squared distance so far.  A pair that provably cannot beat it is
// abandoned and reported as +inf -- it would lose the argmin anyway, and the winner's (and every near-tie's) arithmetic is
// untouched, so labels and distances keep their bits.  Two exits, both LOWER bounds of the pair's msd:
//  (1) before the quartic: lambda_max = max over rotations of tr(R^T M) <= s1+s2+s3 <= sqrt(3) |M|_F
//  (2) inside the Newton loop: started at (Ga+Gb)/2 >= lambda_max, the iterates of this convex, increasing branch of the
//      quartic decrease monotonically towards lambda_max, so (Ga+Gb-2 lam_t)/N never exceeds the final msd.
// The slack (1e-5 relative + 1e-6 (Ga+Gb)/N absolute) is orders of magnitude above the fp32 / fp64 rounding of either side.
__device__ __forceinline__ float qcp_msd(const float* Mf, float Ga, float Gb, int n_atoms,
                                         float bound = 3.402823466e+38f) {
    const bool bounded = bound < 3.0e38f;
    float cut = 0.f;
    if (bounded) {
        const float gs = Ga + Gb;
        cut = bound * 1.00001f + 1e-6f * gs / (float)n_atoms;
        float f = 0.f;
#pragma unroll
        for (int e = 0; e < 9; ++e) f = fmaf(Mf[e], Mf[e], f);
        const float r = sqrtf(3.f * f) * 1.000001f;
        if ((gs - 2.f * r) * (1.f - 1e-6f) / (float)n_atoms > cut) return 3.402823466e+38f;
    }
    const double Sxx = Mf[0], Sxy = Mf[1], Sxz = Mf[2];
    const double Syx = Mf[3], Syy = Mf[4], Syz = Mf[5];
    const double Szx = Mf[6], Szy = Mf[7], Szz = Mf[8];
    const double Sxx2 = DM(Sxx, Sxx), Syy2 = DM(Syy, Syy), Szz2 = DM(Szz, Szz);
    const double Sxy2 = DM(Sxy, Sxy), Syz2 = DM(Syz, Syz), Sxz2 = DM(Sxz, Sxz);
    const double Syx2 = DM(Syx, Syx), Szy2 = DM(Szy, Szy), Szx2 = DM(Szx, Szx);
    const double SyzSzymSyySzz2 = DM(2.0, DS(DM(Syz, Szy), DM(Syy, Szz)));
    const double Sxx2Syy2Szz2Syz2Szy2 = DA(DA(DS(DA(Syy2, Szz2), Sxx2), Syz2), Szy2);
    const double C2 =
        DM(-2.0, DA(DA(DA(DA(DA(DA(DA(DA(Sxx2, Syy2), Szz2), Sxy2), Syx2), Sxz2), Szx2), Syz2), Szy2));
    const double C1 = DM(
        8.0, DS(DS(DS(DA(DA(DM(DM(Sxx, Syz), Szy), DM(DM(Syy, Szx), Sxz)), DM(DM(Szz, Sxy), Syx)),
                      DM(DM(Sxx, Syy), Szz)),
                   DM(DM(Syz, Szx), Sxy)),
                DM(DM(Szy, Syx), Sxz)));
    const double SxzpSzx = DA(Sxz, Szx), SyzpSzy = DA(Syz, Szy), SxypSyx = DA(Sxy, Syx);
    const double SyzmSzy = DS(Syz, Szy), SxzmSzx = DS(Sxz, Szx), SxymSyx = DS(Sxy, Syx);
    const double SxxpSyy = DA(Sxx, Syy), SxxmSyy = DS(Sxx, Syy);
    const double Sxy2Sxz2Syx2Szx2 = DS(DS(DA(Sxy2, Sxz2), Syx2), Szx2);
    const double t0 = DM(Sxy2Sxz2Syx2Szx2, Sxy2Sxz2Syx2Szx2);
    const double t1 = DM(DA(Sxx2Syy2Szz2Syz2Szy2, SyzSzymSyySzz2), DS(Sxx2Syy2Szz2Syz2Szy2, SyzSzymSyySzz2));
    const double t2 = DM(DA(DM(-SxzpSzx, SyzmSzy), DM(SxymSyx, DS(SxxmSyy, Szz))),
                         DA(DM(-SxzmSzx, SyzpSzy), DM(SxymSyx, DA(SxxmSyy, Szz))));
    const double t3 = DM(DS(DM(-SxzpSzx, SyzpSzy), DM(SxypSyx, DS(SxxpSyy, Szz))),
                         DS(DM(-SxzmSzx, SyzmSzy), DM(SxypSyx, DA(SxxpSyy, Szz))));
    const double t4 = DM(DA(DM(SxypSyx, SyzpSzy), DM(SxzpSzx, DA(SxxmSyy, Szz))),
                         DA(DM(-SxymSyx, SyzmSzy), DM(SxzpSzx, DA(SxxpSyy, Szz))));
    const double t5 = DM(DA(DM(SxypSyx, SyzmSzy), DM(SxzmSzx, DS(SxxmSyy, Szz))),
                         DA(DM(-SxymSyx, SyzpSzy), DM(SxzmSzx, DS(SxxpSyy, Szz))));
    const double C0 = DA(DA(DA(DA(DA(t0, t1), t2), t3), t4), t5);

    const double Gs = DA((double)Ga, (double)Gb);
    double lam = DM(Gs, 0.5);
    for (int it = 0; it < 50; ++it) {
        const double old = lam;
        const double x2 = DM(lam, lam);
        const double b = DM(DA(x2, C2), lam);
        const double a = DA(b, C1);
        const double num = DA(DM(a, lam), C0);
        const double den = DA(DA(DM(DM(2.0, x2), lam), b), a);
        if (den == 0.0) break;
        lam = DS(lam, __ddiv_rn(num, den));
        if (fabs(DS(lam, old)) < fabs(DM(1e-11, lam))) break;
        if (bounded && (Gs - 2.0 * lam) > (double)cut * (double)n_atoms) return 3.402823466e+38f;
    }
    double msd = __ddiv_rn(DS(Gs, DM(2.0, lam)), (double)n_atoms);
    if (!(msd > 0.0)) msd = 0.0;
    return __double2float_rn(msd);
}

struct Cov36 {
    float a[9][4];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int e = 0; e < 9; ++e) a[e][0] = a[e][1] = a[e][2] = a[e][3] = 0.f;
    }
    template <int L>
    __device__ __forceinline__ void atom(float ax, float ay, float az, float bx, float by, float bz) {
        a[0][L] = __fadd_rn(a[0][L], __fmul_rn(ax, bx));
        a[1][L] = __fadd_rn(a[1][L], __fmul_rn(ax, by));
        a[2][L] = __fadd_rn(a[2][L], __fmul_rn(ax, bz));
        a[3][L] = __fadd_rn(a[3][L], __fmul_rn(ay, bx));
        a[4][L] = __fadd_rn(a[4][L], __fmul_rn(ay, by));
        a[5][L] = __fadd_rn(a[5][L], __fmul_rn(ay, bz));
        a[6][L] = __fadd_rn(a[6][L], __fmul_rn(az, bx));
        a[7][L] = __fadd_rn(a[7][L], __fmul_rn(az, by));
        a[8][L] = __fadd_rn(a[8][L], __fmul_rn(az, bz));
    }
    __device__ __forceinline__ void finish(float M[9]) const {
#pragma unroll
        for (int e = 0; e < 9; ++e) M[e] = __fadd_rn(__fadd_rn(a[e][0], a[e][1]), __fadd_rn(a[e][2], a[e][3]));
    }
};

// squared minRMSD of (uncentered frame row `x`, its trace Ga) vs (centered center row `c`, Gb);
// rows are 16-byte aligned (global with d%4==0, or padded smem)
template <bool ALIGNED>
__device__ __forceinline__ float rmsd_sq_pair(const float* __restrict__ x, float Ga, const float* __restrict__ c,
                                              float Gb, int n_atoms) {
    Cov36 cov;
    cov.init();
    const int n4 = n_atoms & ~3;
    int t = 0;
    if (ALIGNED) {
        for (; t < n4; t += 4) {
            const float4 x0 = *reinterpret_cast<const float4*>(x + 3 * t);
            const float4 x1 = *reinterpret_cast<const float4*>(x + 3 * t + 4);
            const float4 x2 = *reinterpret_cast<const float4*>(x + 3 * t + 8);
            const float4 c0 = *reinterpret_cast<const float4*>(c + 3 * t);
            const float4 c1 = *reinterpret_cast<const float4*>(c + 3 * t + 4);
            const float4 c2 = *reinterpret_cast<const float4*>(c + 3 * t + 8);
            cov.atom<0>(x0.x, x0.y, x0.z, c0.x, c0.y, c0.z);
            cov.atom<1>(x0.w, x1.x, x1.y, c0.w, c1.x, c1.y);
            cov.atom<2>(x1.z, x1.w, x2.x, c1.z, c1.w, c2.x);
            cov.atom<3>(x2.y, x2.z, x2.w, c2.y, c2.z, c2.w);
        }
    } else {
        for (; t < n4; t += 4) {
            const float* xp = x + 3 * t;
            const float* cp = c + 3 * t;
            cov.atom<0>(xp[0], xp[1], xp[2], cp[0], cp[1], cp[2]);
            cov.atom<1>(xp[3], xp[4], xp[5], cp[3], cp[4], cp[5]);
            cov.atom<2>(xp[6], xp[7], xp[8], cp[6], cp[7], cp[8]);
            cov.atom<3>(xp[9], xp[10], xp[11], cp[9], cp[10], cp[11]);
        }
    }
    if (t < n_atoms) { cov.atom<0>(x[3 * t], x[3 * t + 1], x[3 * t + 2], c[3 * t], c[3 * t + 1], c[3 * t + 2]); ++t; }
    if (t < n_atoms) { cov.atom<1>(x[3 * t], x[3 * t + 1], x[3 * t + 2], c[3 * t], c[3 * t + 1], c[3 * t + 2]); ++t; }
    if (t < n_atoms) { cov.atom<2>(x[3 * t], x[3 * t + 1], x[3 * t + 2], c[3 * t], c[3 * t + 1], c[3 * t + 2]); ++t; }
    float M[9];
    cov.finish(M);
    return qcp_msd(M, Ga, Gb, n_atoms);
}

// ---- centering ----------------------------------------------------------------------------
// one thread per structure; sequential fp64 sums in atom order (oracle center_and_trace)
__global__ void __launch_bounds__(128) rmsd_center_kernel(const float* __restrict__ src, int64_t m, int d,
                                                          float* __restrict__ centered, float* __restrict__ traces) {
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (i >= m) return;
    const int n_atoms = d / 3;
    const float* p = src + i * d;
    double sx = 0, sy = 0, sz = 0;
    for (int t = 0; t < n_atoms; ++t) {
        sx = DA(sx, (double)p[3 * t]);
        sy = DA(sy, (double)p[3 * t + 1]);
        sz = DA(sz, (double)p[3 * t + 2]);
    }
    sx = __ddiv_rn(sx, (double)n_atoms);
    sy = __ddiv_rn(sy, (double)n_atoms);
    sz = __ddiv_rn(sz, (double)n_atoms);
    double g = 0;
    float* q = centered ? centered + i * d : nullptr;
    for (int t = 0; t < n_atoms; ++t) {
        const float x = __double2float_rn(DS((double)p[3 * t], sx));
        const float y = __double2float_rn(DS((double)p[3 * t + 1], sy));
        const float z = __double2float_rn(DS((double)p[3 * t + 2], sz));
        if (q) { q[3 * t] = x; q[3 * t + 1] = y; q[3 * t + 2] = z; }
        g = DA(g, DM((double)x, (double)x));
        g = DA(g, DM((double)y, (double)y));
        g = DA(g, DM((double)z, (double)z));
    }
    traces[i] = __double2float_rn(g);
}

// ---- tile kernel: FB frames x G center groups, frames and centers staged in padded smem ------
struct RTileCfg {
    int FB, G, KT, xstride;
    size_t smem;
};

static RTileCfg rtile_cfg(int d, int k, size_t budget) {
    RTileCfg c;
    const int ds = (d + 3) & ~3;
    c.xstride = ((ds / 4) % 2 == 0) ? ds + 4 : ds + 8;
    int FB = 128;
    while (FB > 4 && (size_t)FB * c.xstride * 4 > budget / 2) FB >>= 1;
    c.FB = FB;
    c.G = 128 / FB;
    const size_t left = budget - (size_t)FB * c.xstride * 4 - 128 * 8 - 1024;
    int KT = (int)(left / ((size_t)ds * 4 + 4));
    KT = (KT / c.G) * c.G;
    const int kmax = (int)cdiv(k, c.G) * c.G;
    if (KT > kmax) KT = kmax;
    if (KT < c.G) KT = c.G;
    c.KT = KT;
    c.smem = (size_t)FB * c.xstride * 4 + (size_t)KT * ds * 4 + (size_t)KT * 4 + 128 * 8 + 64;
    return c;
}

template <int MODE>
__global__ void __launch_bounds__(128) rmsd_tile_kernel(const float* __restrict__ X, const float* __restrict__ Ga,
                                                        int64_t n, int d, const float* __restrict__ Cc,
                                                        const float* __restrict__ Gb, int k, RTileCfg cfg,
                                                        int32_t* __restrict__ labels, float* __restrict__ out,
                                                        int lloyd) {
    extern __shared__ __align__(16) float sm[];
    const int ds = (d + 3) & ~3;
    float* xs = sm;
    float* cs = xs + (size_t)cfg.FB * cfg.xstride;
    float* gb = cs + (size_t)cfg.KT * ds;
    float* red_s = gb + ((cfg.KT + 3) & ~3);
    int32_t* red_j = (int32_t*)(red_s + 128);
    const int tid = threadIdx.x;
    const int f = tid % cfg.FB, g = tid / cfg.FB;
    const int64_t base = (int64_t)blockIdx.x * cfg.FB;
    const int nf = (int)min((int64_t)cfg.FB, n - base);
    const int n_atoms = d / 3;
    {
        const float* src = X + base * d;
        const int total = nf * d;
        for (int t = tid; t < total; t += 128) {
            const int r = t / d, c = t - r * d;
            xs[(size_t)r * cfg.xstride + c] = __ldg(src + t);
        }
    }
    const bool valid = f < nf;
    const float ga = valid ? Ga[base + f] : 0.f;
    const float* xrow = xs + (size_t)f * cfg.xstride;
    ArgMin am;
    am.init();
    for (int j0 = 0; j0 < k; j0 += cfg.KT) {
        const int kk = min(cfg.KT, k - j0);
        __syncthreads();
        {
            const float* src = Cc + (int64_t)j0 * d;
            const int total = kk * d;
            for (int t = tid; t < total; t += 128) {
                const int r = t / d, c = t - r * d;
                cs[(size_t)r * ds + c] = __ldg(src + t);
            }
            for (int t = tid; t < kk; t += 128) gb[t] = Gb[j0 + t];
        }
        __syncthreads();
        if (!valid) continue;
        for (int jj = g; jj < kk; jj += cfg.G) {
            const float s = rmsd_sq_pair<true>(xrow, ga, cs + (size_t)jj * ds, gb[jj], n_atoms);
            if (MODE == MODE_ARGMIN) am.offer(s, j0 + jj);
            else out[(int64_t)(j0 + jj) * n + base + f] = __fsqrt_rn(s);
        }
    }
    if (MODE == MODE_ARGMIN) {
        if (cfg.G > 1) {
            __syncthreads();
            red_s[tid] = am.s;
            red_j[tid] = am.j;
            __syncthreads();
            if (g == 0)
                for (int gg = 1; gg < cfg.G; ++gg) am.merge(red_s[gg * cfg.FB + f], red_j[gg * cfg.FB + f]);
        }
        if (g == 0 && valid) {
            labels[base + f] = (lloyd && am.j < 0) ? 0 : am.j;
            if (out) out[base + f] = am.j >= 0 ? __fsqrt_rn(am.s) : 3.402823466e+38f;
        }
    }
}

// ---- slab kernel: 32 frames x 8*CPT centers per CTA, atoms streamed in slabs of 16 -----------------------------
// The tile kernel above stages whole rows (3.6 KB for 300 atoms), which leaves one 128-thread CTA per SM and a
// barrier between every load and compute phase.  Here a CTA of 256 threads (lane = frame, warp = center slot, CPT
// centers per thread) walks the atoms in slabs of 32 (384 bytes per row): the slab of step t+1 is copied with
// cp.async while step t is being summed, the 36 lane sums of every pair stay in registers across slabs, and 38 KB
// of shared memory per CTA lets two CTAs share an SM.  Per 4 atoms a warp issues 3 conflict-free 16-byte loads
// of its frames, 3*CPT broadcast loads of its centers and 72*CPT fp32 instructions: CUDA-core bound.
// The per-pair arithmetic (Cov36 lane order, qcp_msd) is the one of the tile kernel, so results are identical.
static constexpr int RS_ATOMS = 32;               // atoms per slab
static constexpr int RS_ROW = RS_ATOMS * 3;       // floats per slab row
static constexpr int RS_XROW = RS_ROW + 4;        // padded frame row: 25 x 16 B (odd) -> conflict-free 16-byte loads

__device__ __forceinline__ void rs_cp16(float* dst, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src)
                 : "memory");
}

template <int MODE, int CPT, bool ALIGNED, bool ABANDON = true>
__global__ void __launch_bounds__(256, 2) rmsd_slab_kernel(const float* __restrict__ X, const float* __restrict__ Ga,
                                                           int64_t n, int d, const float* __restrict__ Cc,
                                                           const float* __restrict__ Gb, int k,
                                                           int32_t* __restrict__ labels, float* __restrict__ out,
                                                           int lloyd) {
    constexpr int KT = 8 * CPT;
    constexpr int STAGE = 32 * RS_XROW + KT * RS_ROW;
    __shared__ __align__(16) float sm[2 * STAGE];
    __shared__ float red_s[256];
    __shared__ int32_t red_j[256];
    __shared__ int best_bits[32];  // per frame: float bits of the best squared distance any warp has seen (monotone)
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid < 32) best_bits[tid] = 0x7f800000;
    const int64_t base = (int64_t)blockIdx.x * 32;
    const int n_atoms = d / 3;
    const int n_slabs = (n_atoms + RS_ATOMS - 1) / RS_ATOMS;
    const int n_ctiles = (k + KT - 1) / KT;
    const int64_t fi = base + lane;
    const bool fvalid = fi < n;
    const float ga = fvalid ? Ga[fi] : 0.f;

    // copy the slab of step (center tile jt, slab sl) into stage buffer `buf`.  Which 16-byte pieces a thread copies
    // never changes (piece t -> row t / 24, column group t % 24), so rows, offsets and source row pointers are set
    // up once; a stage only adds the slab's column offset.
    constexpr int PPR = RS_ROW / 4;                         // 16-byte pieces per slab row
    constexpr int XP = (32 * PPR + 255) / 256;              // frame pieces per thread
    constexpr int CP = (KT * PPR + 255) / 256;              // center pieces per thread
    int x_dst[XP], x_c4[XP];
    const float* x_src[XP];
    bool x_ok[XP];
#pragma unroll
    for (int u = 0; u < XP; ++u) {
        const int t = tid + 256 * u, r = t / PPR, c4 = t - r * PPR;
        x_ok[u] = t < 32 * PPR && base + r < n;
        x_dst[u] = t < 32 * PPR ? r * RS_XROW + c4 * 4 : -1;
        x_c4[u] = c4 * 4;
        x_src[u] = X + (x_ok[u] ? base + r : 0) * d + c4 * 4;
    }
    int c_dst[CP], c_c4[CP], c_r[CP];
#pragma unroll
    for (int u = 0; u < CP; ++u) {
        const int t = tid + 256 * u, r = t / PPR, c4 = t - r * PPR;
        c_dst[u] = t < KT * PPR ? 32 * RS_XROW + r * RS_ROW + c4 * 4 : -1;
        c_c4[u] = c4 * 4;
        c_r[u] = r;
    }
    auto copy_piece = [&](float* dst, const float* src, bool row_ok, int col) {
        if (ALIGNED && row_ok && col + 4 <= d) rs_cp16(dst, src);
        else {
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = (row_ok && col + e < d) ? __ldg(src + e) : 0.f;
            *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
        }
    };
    auto stage = [&](int jt, int sl, int buf) {
        float* sb = sm + buf * STAGE;
        const int col0 = sl * RS_ROW;
#pragma unroll
        for (int u = 0; u < XP; ++u)
            if (x_dst[u] >= 0) copy_piece(sb + x_dst[u], x_src[u] + col0, x_ok[u], col0 + x_c4[u]);
#pragma unroll
        for (int u = 0; u < CP; ++u) {
            if (c_dst[u] >= 0) {
                const int j = jt * KT + c_r[u];
                copy_piece(sb + c_dst[u], Cc + (int64_t)(j < k ? j : 0) * d + col0 + c_c4[u], j < k, col0 + c_c4[u]);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    ArgMin am;
    am.init();
    const int steps = n_ctiles * n_slabs;
    stage(0, 0, 0);
    Cov36 cov[CPT];
    int jt = 0, sl = 0;
    for (int step = 0; step < steps; ++step) {
        const int buf = step & 1;
        int njt = jt, nsl = sl + 1;
        if (nsl == n_slabs) { nsl = 0; njt = jt + 1; }
        if (step + 1 < steps) {
            stage(njt, nsl, buf ^ 1);  // the other buffer was released by the barrier that ended step-1
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();  // this step's slab is visible to every thread
        if (sl == 0) {
#pragma unroll
            for (int c = 0; c < CPT; ++c) cov[c].init();
        }
        const float* xr = sm + buf * STAGE + lane * RS_XROW;
        const float* cr = sm + buf * STAGE + 32 * RS_XROW + (w * CPT) * RS_ROW;
#pragma unroll
        for (int q = 0; q < RS_ATOMS / 4; ++q) {
            const float4 x0 = *reinterpret_cast<const float4*>(xr + 12 * q);
            const float4 x1 = *reinterpret_cast<const float4*>(xr + 12 * q + 4);
            const float4 x2 = *reinterpret_cast<const float4*>(xr + 12 * q + 8);
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const float4 c0 = *reinterpret_cast<const float4*>(cr + c * RS_ROW + 12 * q);
                const float4 c1 = *reinterpret_cast<const float4*>(cr + c * RS_ROW + 12 * q + 4);
                const float4 c2 = *reinterpret_cast<const float4*>(cr + c * RS_ROW + 12 * q + 8);
                cov[c].template atom<0>(x0.x, x0.y, x0.z, c0.x, c0.y, c0.z);
                cov[c].template atom<1>(x0.w, x1.x, x1.y, c0.w, c1.x, c1.y);
                cov[c].template atom<2>(x1.z, x1.w, x2.x, c1.z, c1.w, c2.x);
                cov[c].template atom<3>(x2.y, x2.z, x2.w, c2.y, c2.z, c2.w);
            }
        }
        if (sl == n_slabs - 1) {  // the pairs of this center tile are complete
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const int j = jt * KT + w * CPT + c;
                if (j < k && fvalid) {
                    float M[9];
                    cov[c].finish(M);
                    if (MODE == MODE_ARGMIN) {
                        // (a stale bound only abandons less; the first barrier of the loop ordered the initialisation)
                        const float bound = ABANDON ? __int_as_float(*(volatile int*)&best_bits[lane]) : 3.402823466e+38f;
                        const float s = qcp_msd(M, ga, Gb[j], n_atoms, bound);
                        am.offer(s, j);
                        if (ABANDON && s < bound) atomicMin(&best_bits[lane], __float_as_int(s));  // s >= 0: int order = float order
                    } else {
                        out[(int64_t)j * n + fi] = __fsqrt_rn(qcp_msd(M, ga, Gb[j], n_atoms));
                    }
                }
            }
        }
        __syncthreads();  // everyone is done reading `buf` before the next iteration refills it
        jt = njt;
        sl = nsl;
    }
    if (MODE == MODE_ARGMIN) {
        red_s[tid] = am.s;
        red_j[tid] = am.j;
        __syncthreads();
        if (w == 0) {
            for (int g = 1; g < 8; ++g) am.merge(red_s[g * 32 + lane], red_j[g * 32 + lane]);
            if (fvalid) {
                labels[fi] = (lloyd && am.j < 0) ? 0 : am.j;
                if (out) out[fi] = am.j >= 0 ? __fsqrt_rn(am.s) : 3.402823466e+38f;
            }
        }
    }
}

__global__ void __launch_bounds__(128) rmsd_labeled_kernel(const float* __restrict__ X,
                                                           const float* __restrict__ Ga, int64_t n, int d,
                                                           const float* __restrict__ Cc,
                                                           const float* __restrict__ Gb,
                                                           const int32_t* __restrict__ labels,
                                                           float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (i >= n) return;
    const int32_t a = labels[i];
    out[i] = __fsqrt_rn(rmsd_sq_pair<false>(X + i * d, Ga[i], Cc + (int64_t)a * d, Gb[a], d / 3));
}

int launch_rmsd_center(b2k_ctx* ctx, const float* src, int64_t m, int d, float* centered_or_null, float* traces) {
    if (m <= 0) return B2K_OK;
    rmsd_center_kernel<<<(unsigned)cdiv(m, 128), 128, 0, ctx->stream>>>(src, m, d, centered_or_null, traces);
    LAUNCH_CHECK();
    return B2K_OK;
}

static int launch_rtile(b2k_ctx* ctx, const float* X, const float* Ga, int64_t n, int d, const float* Cc,
                        const float* Gb, int k, int32_t* labels, float* out, int lloyd, int mode) {
    if (n <= 0 || k <= 0) return B2K_OK;
    if (ctx->rmsd_kernel != 1 && n >= 32) {
        const unsigned grid = (unsigned)cdiv(n, 32);
        const bool al = (d % 4 == 0) && ((((uintptr_t)X) | ((uintptr_t)Cc)) & 15) == 0;
// (two CTAs of ~40 KB static shared memory per SM: ask for the maximum shared-memory carve-out instead of leaving the
//  L1 / shared split of each launch to the driver's heuristic)
#define B2K_RS(MODE_, CPT_, AL_)                                                                                          \
    do {                                                                                                                  \
        static PerDeviceOnce carve;                                                                                       \
        if (carve.need(ctx->device)) {                                                                                    \
            CUDA_TRY(cudaFuncSetAttribute(rmsd_slab_kernel<MODE_, CPT_, AL_, true>,                                       \
                                          cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)); \
            CUDA_TRY(cudaFuncSetAttribute(rmsd_slab_kernel<MODE_, CPT_, AL_, false>,                                      \
                                          cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)); \
            carve.done(ctx->device);                                                                                      \
        }                                                                                                                 \
        if (ctx->rmsd_abandon)                                                                                            \
            rmsd_slab_kernel<MODE_, CPT_, AL_, true><<<grid, 256, 0, ctx->stream>>>(X, Ga, n, d, Cc, Gb, k, labels, out, lloyd); \
        else                                                                                                              \
            rmsd_slab_kernel<MODE_, CPT_, AL_, false><<<grid, 256, 0, ctx->stream>>>(X, Ga, n, d, Cc, Gb, k, labels, out, lloyd); \
    } while (0)
        if (mode == MODE_ARGMIN) {
            if (k > 8) { if (al) B2K_RS(MODE_ARGMIN, 2, true); else B2K_RS(MODE_ARGMIN, 2, false); }
            else { if (al) B2K_RS(MODE_ARGMIN, 1, true); else B2K_RS(MODE_ARGMIN, 1, false); }
        } else {
            if (k > 8) { if (al) B2K_RS(MODE_ALL, 2, true); else B2K_RS(MODE_ALL, 2, false); }
            else { if (al) B2K_RS(MODE_ALL, 1, true); else B2K_RS(MODE_ALL, 1, false); }
        }
#undef B2K_RS
        LAUNCH_CHECK();
        return B2K_OK;
    }
    const size_t budget = std::min<size_t>(ctx->smem_optin, 200 * 1024);
    RTileCfg cfg = rtile_cfg(d, k, budget);
    if (cfg.smem > ctx->smem_optin)
        return set_error(B2K_ERR_INVALID_ARG, "dimension %d too large for the minRMSD tile kernel", d);
    static PerDeviceOnce attr_set;
    if (attr_set.need(ctx->device)) {
        CUDA_TRY(cudaFuncSetAttribute(rmsd_tile_kernel<MODE_ARGMIN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)ctx->smem_optin));
        CUDA_TRY(cudaFuncSetAttribute(rmsd_tile_kernel<MODE_ALL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)ctx->smem_optin));
        attr_set.done(ctx->device);
    }
    const int64_t blocks = cdiv(n, cfg.FB);
    if (mode == MODE_ARGMIN)
        rmsd_tile_kernel<MODE_ARGMIN><<<(unsigned)blocks, 128, cfg.smem, ctx->stream>>>(X, Ga, n, d, Cc, Gb, k, cfg,
                                                                                         labels, out, lloyd);
    else
        rmsd_tile_kernel<MODE_ALL><<<(unsigned)blocks, 128, cfg.smem, ctx->stream>>>(X, Ga, n, d, Cc, Gb, k, cfg,
                                                                                      labels, out, lloyd);
    LAUNCH_CHECK();
    return B2K_OK;
}

int launch_rmsd_assign(b2k_ctx* ctx, const float* X, const float* Ga, int64_t n, int d, const float* Cc,
                       const float* Gb, int k, int32_t* labels, float* mind, int lloyd) {
    return launch_rtile(ctx, X, Ga, n, d, Cc, Gb, k, labels, mind, lloyd, MODE_ARGMIN);
}

int launch_rmsd_dist_rows(b2k_ctx* ctx, const float* X, const float* Ga, int64_t n, int d, const float* Rc,
                          const float* Gb, int m, float* out) {
    return launch_rtile(ctx, X, Ga, n, d, Rc, Gb, m, nullptr, out, 0, MODE_ALL);
}

int launch_rmsd_labeled_dist(b2k_ctx* ctx, const float* X, const float* Ga, int64_t n, int d, const float* Cc,
                             const float* Gb, const int32_t* labels, float* out) {
    if (n <= 0) return B2K_OK;
    rmsd_labeled_kernel<<<(unsigned)cdiv(n, 128), 128, 0, ctx->stream>>>(X, Ga, n, d, Cc, Gb, labels, out);
    LAUNCH_CHECK();
    return B2K_OK;
}

}  // namespace b2k
