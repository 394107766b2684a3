// oracle/oracle.cpp -- TEST INFRASTRUCTURE ONLY (parity oracle + CPU baseline).
//
// CPU restatement of the arithmetic behind PyEMMA's k-means / assign / regspace /
// minRMSD hot path.  Nothing in pyemma_b200/ may import, link or call this file;
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs do.
//
// PARITY STATUS: "parity unpinned" for everything the reference's own tests do not
// pin.  At reference commit 3327f28 the arithmetic lives in two third-party packages
// that are absent from /root/reference and not installable offline:
//   * deeptime (setup.py:307 `deeptime>=0.4.2`)  -- Euclidean metric, assign, Lloyd,
//     cost, k-means++, regspace (call sites: clustering/kmeans.py:211-214,254-258;
//     clustering/regspace.py:144,150; clustering/interface.py:164-165)
//   * mdtraj  (setup.py:300 `mdtraj>=1.9.2`, libtheobald) -- centering + QCP msd
//     (call sites: clustering/src/clustering_module.cpp:21-28)
// The functions below restate the published algorithms of those packages as recorded
// in SURVEY.md Appendix A/B, and are pinned against every known-answer test the
// reference holds for the path and independent fp64 numpy/Kabsch implementations
// (tests/test_oracle_kats.py), and against scikit-learn's Lloyd / k-means++ and scipy's
// Rotation.align_vectors as third-party implementations (tests/test_oracle_independent.py).
//
// Build flags are part of the oracle's definition (oracle/Makefile):
//   g++ -O3 -fopenmp -ffp-contract=off   (no -march; x86-64 baseline, no FMA)
//
// All data fp32 C-contiguous (n,d); labels int32; centers fp32 (k,d).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <random>
#include <vector>
#include <omp.h>

#define ORC_API extern "C" __attribute__((visibility("default")))

enum { ORC_EUCLIDEAN = 0, ORC_MINRMSD = 1 };

// ---------------------------------------------------------------------------------
// A.1 Euclidean metric  (deeptime metric.h EuclideanMetric::compute_squared/compute;
// SURVEY Appendix A.1).  The upstream loop carries `#pragma omp simd reduction(+:sum)`;
// with gcc -fopenmp on x86-64 baseline that becomes 4 interleaved lane accumulators,
// the d%4 tail added into lane 0, then ((0+a0)+a1)+a2)+a3  (SURVEY Appendix B.1).
// This explicit form IS the oracle's definition; orc_euclid_sq_pragma below is the
// literal pragma loop, kept only so a test can check the two agree bit-for-bit under
// the pinned build flags.
// ---------------------------------------------------------------------------------
static inline float euclid_sq(const float* x, const float* y, int64_t d) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const int64_t d4 = d & ~int64_t(3);
    for (int64_t i = 0; i < d4; i += 4) {
        float t0 = x[i] - y[i], t1 = x[i + 1] - y[i + 1];
        float t2 = x[i + 2] - y[i + 2], t3 = x[i + 3] - y[i + 3];
        a0 = a0 + t0 * t0; a1 = a1 + t1 * t1; a2 = a2 + t2 * t2; a3 = a3 + t3 * t3;
    }
    for (int64_t i = d4; i < d; ++i) { float t = x[i] - y[i]; a0 = a0 + t * t; }
    return (((0.0f + a0) + a1) + a2) + a3;
}

ORC_API float orc_euclid_sq(const float* x, const float* y, int64_t d) { return euclid_sq(x, y, d); }

ORC_API float orc_euclid_sq_pragma(const float* x, const float* y, int64_t d) {
    float sum = 0.0f;
#pragma omp simd reduction(+ : sum)
    for (int64_t i = 0; i < d; ++i) { float t = x[i] - y[i]; sum += t * t; }
    return sum;
}

ORC_API float orc_euclid_sq_seq(const float* x, const float* y, int64_t d) {
    volatile float sum = 0.0f;  // strict left-to-right (what gcc emits WITHOUT -fopenmp)
    for (int64_t i = 0; i < d; ++i) { float t = x[i] - y[i]; sum = sum + t * t; }
    return sum;
}

// ---------------------------------------------------------------------------------
// A.6 minRMSD  (pyemma/coordinates/clustering/src/clustering_module.cpp:9-36 wrapping
// mdtraj center.h inplace_center_and_trace_atom_major + theobald_rmsd msd_atom_major).
// Restated from the published QCP algorithm (Theobald 2005; Liu, Agrafiotis, Theobald
// 2010).  libtheobald's SIMD summation order is not available offline: the order used
// here (4 atom lanes, (l0+l1)+(l2+l3) horizontal add, no FMA) is this oracle's
// definition -- "parity unpinned upstream".
// ---------------------------------------------------------------------------------
static void center_and_trace(float* c, float* trace, int n_atoms) {
    // centroid accumulated in double, subtracted in double, stored as float;
    // trace G = sum(x^2+y^2+z^2) of the centered floats accumulated in double.
    double sx = 0, sy = 0, sz = 0;
    for (int i = 0; i < n_atoms; ++i) { sx += c[3 * i]; sy += c[3 * i + 1]; sz += c[3 * i + 2]; }
    sx /= n_atoms; sy /= n_atoms; sz /= n_atoms;
    double g = 0;
    for (int i = 0; i < n_atoms; ++i) {
        float x = (float)((double)c[3 * i] - sx);
        float y = (float)((double)c[3 * i + 1] - sy);
        float z = (float)((double)c[3 * i + 2] - sz);
        c[3 * i] = x; c[3 * i + 1] = y; c[3 * i + 2] = z;
        g += (double)x * (double)x; g += (double)y * (double)y; g += (double)z * (double)z;
    }
    *trace = (float)g;
}

// cross-covariance M[3*p+q] = sum_atoms a_p * b_q, fp32, 4 atom lanes (atom t -> lane t%4),
// mul then add (SSE2: no FMA), horizontal add (l0+l1)+(l2+l3).
static void cross_cov(const float* a, const float* b, int n_atoms, float M[9]) {
    float acc[9][4];
    for (int e = 0; e < 9; ++e) acc[e][0] = acc[e][1] = acc[e][2] = acc[e][3] = 0.f;
    for (int t = 0; t < n_atoms; ++t) {
        const int l = t & 3;
        const float ax = a[3 * t], ay = a[3 * t + 1], az = a[3 * t + 2];
        const float bx = b[3 * t], by = b[3 * t + 1], bz = b[3 * t + 2];
        acc[0][l] = acc[0][l] + ax * bx; acc[1][l] = acc[1][l] + ax * by; acc[2][l] = acc[2][l] + ax * bz;
        acc[3][l] = acc[3][l] + ay * bx; acc[4][l] = acc[4][l] + ay * by; acc[5][l] = acc[5][l] + ay * bz;
        acc[6][l] = acc[6][l] + az * bx; acc[7][l] = acc[7][l] + az * by; acc[8][l] = acc[8][l] + az * bz;
    }
    for (int e = 0; e < 9; ++e) M[e] = (acc[e][0] + acc[e][1]) + (acc[e][2] + acc[e][3]);
}

// QCP: largest root of P(l) = l^4 + C2 l^2 + C1 l + C0 by Newton from (Ga+Gb)/2, in double.
static float msd_from_M_and_G(const float Mf[9], float Ga, float Gb, int n_atoms) {
    const double Sxx = Mf[0], Sxy = Mf[1], Sxz = Mf[2];
    const double Syx = Mf[3], Syy = Mf[4], Syz = Mf[5];
    const double Szx = Mf[6], Szy = Mf[7], Szz = Mf[8];
    const double Sxx2 = Sxx * Sxx, Syy2 = Syy * Syy, Szz2 = Szz * Szz;
    const double Sxy2 = Sxy * Sxy, Syz2 = Syz * Syz, Sxz2 = Sxz * Sxz;
    const double Syx2 = Syx * Syx, Szy2 = Szy * Szy, Szx2 = Szx * Szx;
    const double SyzSzymSyySzz2 = 2.0 * (Syz * Szy - Syy * Szz);
    const double Sxx2Syy2Szz2Syz2Szy2 = (((Syy2 + Szz2) - Sxx2) + Syz2) + Szy2;
    const double C2 = -2.0 * ((((((((Sxx2 + Syy2) + Szz2) + Sxy2) + Syx2) + Sxz2) + Szx2) + Syz2) + Szy2);
    const double C1 = 8.0 * ((((((Sxx * Syz) * Szy + (Syy * Szx) * Sxz) + (Szz * Sxy) * Syx) - (Sxx * Syy) * Szz) -
                              (Syz * Szx) * Sxy) - (Szy * Syx) * Sxz);
    const double SxzpSzx = Sxz + Szx, SyzpSzy = Syz + Szy, SxypSyx = Sxy + Syx;
    const double SyzmSzy = Syz - Szy, SxzmSzx = Sxz - Szx, SxymSyx = Sxy - Syx;
    const double SxxpSyy = Sxx + Syy, SxxmSyy = Sxx - Syy;
    const double Sxy2Sxz2Syx2Szx2 = ((Sxy2 + Sxz2) - Syx2) - Szx2;
    const double t0 = Sxy2Sxz2Syx2Szx2 * Sxy2Sxz2Syx2Szx2;
    const double t1 = (Sxx2Syy2Szz2Syz2Szy2 + SyzSzymSyySzz2) * (Sxx2Syy2Szz2Syz2Szy2 - SyzSzymSyySzz2);
    const double t2 = ((-SxzpSzx) * SyzmSzy + SxymSyx * (SxxmSyy - Szz)) *
                      ((-SxzmSzx) * SyzpSzy + SxymSyx * (SxxmSyy + Szz));
    const double t3 = ((-SxzpSzx) * SyzpSzy - SxypSyx * (SxxpSyy - Szz)) *
                      ((-SxzmSzx) * SyzmSzy - SxypSyx * (SxxpSyy + Szz));
    const double t4 = (SxypSyx * SyzpSzy + SxzpSzx * (SxxmSyy + Szz)) *
                      ((-SxymSyx) * SyzmSzy + SxzpSzx * (SxxpSyy + Szz));
    const double t5 = (SxypSyx * SyzmSzy + SxzmSzx * (SxxmSyy - Szz)) *
                      ((-SxymSyx) * SyzpSzy + SxzmSzx * (SxxpSyy - Szz));
    const double C0 = ((((t0 + t1) + t2) + t3) + t4) + t5;

    const double E0 = ((double)Ga + (double)Gb) * 0.5;
    double lam = E0;
    for (int it = 0; it < 50; ++it) {
        const double old = lam;
        const double x2 = lam * lam;
        const double b = (x2 + C2) * lam;
        const double a = b + C1;
        const double num = a * lam + C0;
        const double den = (2.0 * x2) * lam + b + a;
        if (den == 0.0) break;
        lam = lam - num / den;
        if (std::fabs(lam - old) < std::fabs(1e-11 * lam)) break;
    }
    double msd = (((double)Ga + (double)Gb) - 2.0 * lam) / (double)n_atoms;
    if (!(msd > 0.0)) msd = 0.0;  // clamp (also maps NaN -> 0)
    return (float)msd;
}

// clustering_module.cpp:10-30: buffer_a/buffer_b are centered copies; traces come from
// them; `a` handed to msd_atom_major is the ORIGINAL (uncentered) xs (float case :28).
// `yc`/`Gb` may be passed pre-centered (centering is idempotent in value only up to
// rounding, so the oracle re-centers ys every call exactly like the reference does).
static float rmsd_sq(const float* xs, const float* ys, int64_t dim, std::vector<float>& ba, std::vector<float>& bb) {
    const int n_atoms = (int)(dim / 3);
    ba.assign(xs, xs + dim);
    bb.assign(ys, ys + dim);
    float Ga, Gb;
    center_and_trace(ba.data(), &Ga, n_atoms);
    center_and_trace(bb.data(), &Gb, n_atoms);
    float M[9];
    cross_cov(xs, bb.data(), n_atoms, M);
    return msd_from_M_and_G(M, Ga, Gb, n_atoms);
}

struct Scratch { std::vector<float> a, b; };

template <int METRIC>
static inline float compute_sq(const float* x, const float* y, int64_t d, Scratch& s) {
    if (METRIC == ORC_EUCLIDEAN) return euclid_sq(x, y, d);
    return rmsd_sq(x, y, d, s.a, s.b);
}
template <int METRIC>
static inline float compute(const float* x, const float* y, int64_t d, Scratch& s) {
    return std::sqrt(compute_sq<METRIC>(x, y, d, s));
}

// compute_metric (clustering_module.cpp:41-43) and its squared form.
// returns NaN-free float; sets *err=1 when metric==minRMSD and d%3!=0 (std::range_error :12-14)
ORC_API float orc_compute_metric(const float* x, const float* y, int64_t d, int metric, int* err) {
    Scratch s;
    if (err) *err = 0;
    if (metric == ORC_MINRMSD) {
        if (d % 3 != 0) { if (err) *err = 1; return 0.f; }
        return compute<ORC_MINRMSD>(x, y, d, s);
    }
    return compute<ORC_EUCLIDEAN>(x, y, d, s);
}
ORC_API void orc_center_and_trace(float* c, int n_atoms, float* trace) { center_and_trace(c, trace, n_atoms); }

// ---------------------------------------------------------------------------------
// A.2 assign  (deeptime assign_chunk_to_centers; call site interface.py:164-165).
// argmin_j compute(x_i, c_j): strict '<' scan from j=0 starting at FLT_MAX/-1
// => lowest index wins ties; comparison happens AFTER the sqrt.
// ---------------------------------------------------------------------------------
template <int METRIC>
static void assign_t(const float* X, int64_t n, int64_t d, const float* C, int64_t k, int n_threads, int32_t* out) {
#pragma omp parallel num_threads(n_threads)
    {
        Scratch s;
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            float mind = std::numeric_limits<float>::max();
            int32_t arg = -1;
            for (int64_t j = 0; j < k; ++j) {
                const float dj = compute<METRIC>(X + i * d, C + j * d, d, s);
                if (dj < mind) { mind = dj; arg = (int32_t)j; }
            }
            out[i] = arg;
        }
    }
}

ORC_API int orc_assign(const float* X, int64_t n, int64_t d, const float* C, int64_t k, int metric, int n_threads,
                       int32_t* out) {
    if (n_threads < 1) n_threads = 1;
    if (metric == ORC_MINRMSD) {
        if (d % 3) return 3;
        assign_t<ORC_MINRMSD>(X, n, d, C, k, n_threads, out);
    } else {
        assign_t<ORC_EUCLIDEAN>(X, n, d, C, k, n_threads, out);
    }
    return 0;
}

// ---------------------------------------------------------------------------------
// A.3 Lloyd step  (deeptime kmeans.cluster).  labels: minDist starts at center 0
// (so a NaN frame gets label 0, unlike assign's -1); sums fp32 in frame order
// (acc_mode 0, the reference's serial branch) or fp64 (acc_mode 1, diagnostic);
// count==0 keeps the old center.  With n_threads>1 labels are computed in parallel
// and the accumulation stays in frame order (deterministic; the reference's threaded
// branch accumulates inside `omp critical` in arbitrary order -- slower and
// non-deterministic, tests/test_kmeans.py:104-110).
// ---------------------------------------------------------------------------------
template <int METRIC>
static void lloyd_labels(const float* X, int64_t n, int64_t d, const float* C, int64_t k, int n_threads,
                         int32_t* labels) {
#pragma omp parallel num_threads(n_threads)
    {
        Scratch s;
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            int32_t arg = 0;
            float mind = compute<METRIC>(X + i * d, C, d, s);
            for (int64_t j = 1; j < k; ++j) {
                const float dj = compute<METRIC>(X + i * d, C + j * d, d, s);
                if (dj < mind) { mind = dj; arg = (int32_t)j; }
            }
            labels[i] = arg;
        }
    }
}

static void lloyd_update(const float* X, int64_t n, int64_t d, const float* C, int64_t k, const int32_t* labels,
                         int acc_mode, float* newC) {
    std::vector<uint64_t> cnt(k, 0);
    if (acc_mode == 0) {
        std::fill(newC, newC + k * d, 0.f);
        for (int64_t i = 0; i < n; ++i) {
            const int32_t a = labels[i];
            cnt[a]++;
            float* dst = newC + (int64_t)a * d;
            const float* src = X + i * d;
            for (int64_t j = 0; j < d; ++j) dst[j] = dst[j] + src[j];
        }
        for (int64_t c = 0; c < k; ++c) {
            if (cnt[c] == 0) std::memcpy(newC + c * d, C + c * d, sizeof(float) * d);
            else for (int64_t j = 0; j < d; ++j) newC[c * d + j] = newC[c * d + j] / (float)cnt[c];
        }
    } else {
        std::vector<double> acc((size_t)(k * d), 0.0);
        for (int64_t i = 0; i < n; ++i) {
            const int32_t a = labels[i];
            cnt[a]++;
            for (int64_t j = 0; j < d; ++j) acc[(int64_t)a * d + j] += (double)X[i * d + j];
        }
        for (int64_t c = 0; c < k; ++c) {
            if (cnt[c] == 0) std::memcpy(newC + c * d, C + c * d, sizeof(float) * d);
            else for (int64_t j = 0; j < d; ++j) newC[c * d + j] = (float)(acc[c * d + j] / (double)cnt[c]);
        }
    }
}

ORC_API int orc_kmeans_cluster(const float* X, int64_t n, int64_t d, const float* C, int64_t k, int metric,
                               int n_threads, int acc_mode, float* newC, int32_t* labels) {
    if (n_threads < 1) n_threads = 1;
    if (metric == ORC_MINRMSD) {
        if (d % 3) return 3;
        lloyd_labels<ORC_MINRMSD>(X, n, d, C, k, n_threads, labels);
    } else {
        lloyd_labels<ORC_EUCLIDEAN>(X, n, d, C, k, n_threads, labels);
    }
    lloyd_update(X, n, d, C, k, labels, acc_mode, newC);
    return 0;
}

// costAssignFunction (deeptime): value += l*l with l = compute(x_i, c[label_i]) -- the
// sqrt-then-square form, summed in T.  acc_mode 0: fp32 in frame order (1 thread) or
// per-thread fp32 partials combined in thread order; acc_mode 1: fp64 sum of the same
// fp32 l*l terms, rounded to fp32 at the end.
template <int METRIC>
static float cost_t(const float* X, int64_t n, int64_t d, const float* C, const int32_t* labels, int n_threads,
                    int acc_mode) {
    if (acc_mode == 1) {
        double v = 0;
#pragma omp parallel num_threads(n_threads)
        {
            Scratch s;
            double p = 0;
#pragma omp for schedule(static)
            for (int64_t i = 0; i < n; ++i) {
                const float l = compute<METRIC>(X + i * d, C + (int64_t)labels[i] * d, d, s);
                p += (double)(l * l);
            }
#pragma omp critical
            v += p;
        }
        return (float)v;
    }
    std::vector<float> part((size_t)n_threads, 0.f);
#pragma omp parallel num_threads(n_threads)
    {
        Scratch s;
        float p = 0.f;
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            const float l = compute<METRIC>(X + i * d, C + (int64_t)labels[i] * d, d, s);
            p = p + l * l;
        }
        part[omp_get_thread_num()] = p;
    }
    float v = 0.f;
    for (int t = 0; t < n_threads; ++t) v = v + part[t];
    return v;
}

ORC_API float orc_cost(const float* X, int64_t n, int64_t d, const float* C, int64_t k, const int32_t* labels,
                       int metric, int n_threads, int acc_mode) {
    (void)k;
    if (n_threads < 1) n_threads = 1;
    if (metric == ORC_MINRMSD) return cost_t<ORC_MINRMSD>(X, n, d, C, labels, n_threads, acc_mode);
    return cost_t<ORC_EUCLIDEAN>(X, n, d, C, labels, n_threads, acc_mode);
}

// cluster_loop (deeptime kmeans.cluster_loop; call site kmeans.py:254-258):
// do { step; cost with NEW centers + labels from the OLD centers; rel=|cost-prev|/cost
// (0 if cost==0); converged if rel<=tol else callback; it++ } while (it<max_iter && !conv).
// max_iter==0 still executes one step (do-while).  code 0 = converged, 1 = not.
// centers_hist (optional, max_iter_cap*k*d floats) receives the centers after every step.
typedef void (*orc_callback)(void*);
ORC_API int orc_cluster_loop(const float* X, int64_t n, int64_t d, float* C_io, int64_t k, int metric, int n_threads,
                             int max_iter, float tol, int acc_mode, orc_callback cb, void* user, int* code,
                             int* iters, float* inertias, int inertias_cap, float* centers_hist, int32_t* labels_out) {
    if (n_threads < 1) n_threads = 1;
    if (metric == ORC_MINRMSD && d % 3) return 3;
    std::vector<float> cur(C_io, C_io + k * d), nxt((size_t)(k * d));
    std::vector<int32_t> labels((size_t)n);
    int it = 0;
    bool converged = false;
    float prev = 0.f;
    do {
        orc_kmeans_cluster(X, n, d, cur.data(), k, metric, n_threads, acc_mode, nxt.data(), labels.data());
        cur.swap(nxt);
        const float cost = orc_cost(X, n, d, cur.data(), k, labels.data(), metric, n_threads, acc_mode);
        if (it < inertias_cap) inertias[it] = cost;
        if (centers_hist && it < inertias_cap) std::memcpy(centers_hist + (int64_t)it * k * d, cur.data(), sizeof(float) * k * d);
        const float rel = (cost != 0.0f) ? std::fabs(cost - prev) / cost : 0.f;
        prev = cost;
        if (rel <= tol) converged = true;
        else if (cb) cb(user);
        it += 1;
    } while (it < max_iter && !converged);
    std::memcpy(C_io, cur.data(), sizeof(float) * k * d);
    if (labels_out) std::memcpy(labels_out, labels.data(), sizeof(int32_t) * n);
    *code = converged ? 0 : 1;
    *iters = it;
    return 0;
}

// ---------------------------------------------------------------------------------
// A.4 k-means++  (deeptime kmeans.init_centers_kmpp(data,k,random_seed,n_threads,cb);
// signature evidenced by tests/test_kmeans.py:304).
//
// RNG: std::mt19937 seeded with (uint32)seed (seed<0: std::random_device);
// first center: std::uniform_int_distribution<size_t>(0,n-1) (libstdc++ = Lemire);
// per round n_trials = 2+(size_t)log(k) thresholds r_j = dist_sum * ((T)gen()/(T)gen.max()).
//
// scan_mode 0 ("serial", reference-faithful at n_jobs=1): running fp32 prefix in frame
// order, candidate_j = first non-taken i with sum >= r_j; potentials and dist_sum are
// fp32 sums in frame order (OpenMP reduction semantics with one thread: a private
// accumulator starting at 0 that is added to the shared value once at the end).
// scan_mode 1 ("blocked", for large n; a DEFINED deviation shared with the GPU path):
// every ordered sum is the balanced binary tree over aligned power-of-two index blocks
// (taken frames weigh 0); dist_sum is recomputed as the tree root each round;
// candidate_j is found by descending the tree: left if r <= sum(left) else r -= sum(left).
// ---------------------------------------------------------------------------------
static float tree_sum(const float* v, int64_t lo, int64_t hi, int64_t n) {
    // balanced tree over [lo,hi), hi-lo a power of two; out-of-range leaves are 0.
    if (lo >= n) return 0.f;
    if (hi - lo == 1) return v[lo];
    const int64_t mid = lo + (hi - lo) / 2;
    return tree_sum(v, lo, mid, n) + tree_sum(v, mid, hi, n);
}
static int64_t pow2_ceil(int64_t n) { int64_t p = 1; while (p < n) p <<= 1; return p; }

// Build all levels of the balanced tree bottom-up. level[0] = leaves (size P), level[h] = root (size 1).
static void tree_build(const std::vector<float>& leaves, std::vector<std::vector<float>>& levels) {
    levels.clear();
    levels.push_back(leaves);
    while (levels.back().size() > 1) {
        const std::vector<float>& lo = levels.back();
        std::vector<float> up(lo.size() / 2);
        for (size_t i = 0; i < up.size(); ++i) up[i] = lo[2 * i] + lo[2 * i + 1];
        levels.push_back(std::move(up));
    }
}
static int64_t tree_descend(const std::vector<std::vector<float>>& levels, float r) {
    int64_t node = 0;
    for (int h = (int)levels.size() - 1; h > 0; --h) {
        const float left = levels[h - 1][2 * node];
        if (r <= left) node = 2 * node;
        else { r = r - left; node = 2 * node + 1; }
    }
    return node;
}

template <int METRIC>
static int kmpp_t(const float* X, int64_t n, int64_t d, int64_t k, int64_t seed, int n_threads, int scan_mode,
                  float* centers, int64_t* chosen, orc_callback cb, void* user) {
    const size_t NONE = std::numeric_limits<size_t>::max();
    const size_t n_trials = 2 + (size_t)std::log((double)k);
    std::vector<char> taken((size_t)n, 0);
    std::vector<float> sq((size_t)n, 0.f);
    std::vector<size_t> cand(n_trials);
    std::vector<float> rands(n_trials), pot(n_trials);
    std::mt19937 gen;
    if (seed < 0) { std::random_device rd; gen.seed(rd()); } else gen.seed((uint32_t)seed);
    std::uniform_int_distribution<size_t> uni(0, (size_t)n - 1);
    const size_t first = uni(gen);
    taken[first] = 1;
    std::memcpy(centers, X + (int64_t)first * d, sizeof(float) * d);
    if (chosen) chosen[0] = (int64_t)first;
    int64_t found = 1;
    if (cb) cb(user);

    const int64_t P = pow2_ceil(n);
    std::vector<float> leaves;
    std::vector<std::vector<float>> levels;

    float dist_sum = 0.f;
    {
        // D2[i] = compute(x_i, c0)^2 ; dist_sum = ordered sum
#pragma omp parallel num_threads(n_threads)
        {
            Scratch s;
#pragma omp for schedule(static)
            for (int64_t i = 0; i < n; ++i) {
                if ((size_t)i != first) {
                    float v = compute<METRIC>(X + i * d, X + (int64_t)first * d, d, s);
                    sq[i] = v * v;
                }
            }
        }
        if (scan_mode == 0) {
            float p = 0.f;
            for (int64_t i = 0; i < n; ++i) if ((size_t)i != first) p = p + sq[i];
            dist_sum = dist_sum + p;
        }
    }

    std::vector<float> cd;  // candidate distances [n][n_trials] (min(D2, d2))
    cd.resize((size_t)n * n_trials);
    while (found < k) {
        if (scan_mode == 1) {
            leaves.assign((size_t)P, 0.f);
            for (int64_t i = 0; i < n; ++i) leaves[i] = taken[i] ? 0.f : sq[i];
            tree_build(leaves, levels);
            dist_sum = levels.back()[0];
        }
        for (size_t j = 0; j < n_trials; ++j) {
            cand[j] = NONE;
            rands[j] = dist_sum * ((float)gen() / (float)gen.max());
            pot[j] = 0.f;
        }
        if (scan_mode == 0) {
            float sum = 0.f;
            for (int64_t i = 0; i < n; ++i) {
                if (taken[i]) continue;
                sum = sum + sq[i];
                for (size_t j = 0; j < n_trials; ++j)
                    if (sum >= rands[j] && cand[j] == NONE) cand[j] = (size_t)i;
            }
        } else {
            for (size_t j = 0; j < n_trials; ++j) {
                const int64_t leaf = tree_descend(levels, rands[j]);
                cand[j] = (leaf < n && !taken[leaf]) ? (size_t)leaf : NONE;
            }
        }
        // potentials: p_j = ordered sum over non-taken i of min(D2[i], compute(x_i, cand_j)^2)
#pragma omp parallel num_threads(n_threads)
        {
            Scratch s;
#pragma omp for schedule(static)
            for (int64_t i = 0; i < n; ++i) {
                for (size_t j = 0; j < n_trials; ++j) {
                    float contrib = 0.f;
                    if (!taken[i] && cand[j] != NONE && cand[j] != (size_t)i) {
                        float v = compute<METRIC>(X + i * d, X + (int64_t)cand[j] * d, d, s);
                        float dd = v * v;
                        contrib = (dd < sq[i]) ? dd : sq[i];
                    }
                    cd[(size_t)i * n_trials + j] = contrib;
                }
            }
        }
        for (size_t j = 0; j < n_trials; ++j) {
            if (cand[j] == NONE) continue;
            if (scan_mode == 0) {
                float p = 0.f;
                for (int64_t i = 0; i < n; ++i)
                    if (!taken[i] && cand[j] != (size_t)i) p = p + cd[(size_t)i * n_trials + j];
                pot[j] = pot[j] + p;
            } else {
                leaves.assign((size_t)P, 0.f);
                for (int64_t i = 0; i < n; ++i) leaves[i] = cd[(size_t)i * n_trials + j];
                std::vector<std::vector<float>> lv;
                tree_build(leaves, lv);
                pot[j] = lv.back()[0];
            }
        }
        int64_t best = -1;
        float best_pot = std::numeric_limits<float>::max();
        for (size_t j = 0; j < n_trials; ++j)
            if (cand[j] != NONE && pot[j] < best_pot) { best_pot = pot[j]; best = (int64_t)cand[j]; }
        if (best == -1)
            for (int64_t i = 0; i < n; ++i) if (!taken[i]) { best = i; break; }
        if (best < 0) break;
        std::memcpy(centers + found * d, X + best * d, sizeof(float) * d);
        if (chosen) chosen[found] = best;
        found++;
        if (cb) cb(user);
        taken[best] = 1;
        dist_sum = dist_sum - sq[best];
        if (found < k) {
            std::vector<float> delta((size_t)n, 0.f);
#pragma omp parallel num_threads(n_threads)
            {
                Scratch s;
#pragma omp for schedule(static)
                for (int64_t i = 0; i < n; ++i) {
                    if (taken[i]) continue;
                    float v = compute<METRIC>(X + i * d, X + best * d, d, s);
                    float dd = v * v;
                    if (dd < sq[i]) { delta[i] = dd - sq[i]; sq[i] = dd; }
                }
            }
            if (scan_mode == 0) {
                float p = 0.f;
                for (int64_t i = 0; i < n; ++i) if (delta[i] != 0.f) p = p + delta[i];
                dist_sum = dist_sum + p;
            }
        }
    }
    return found == k ? 0 : 2;
}

ORC_API int orc_kmpp_init(const float* X, int64_t n, int64_t d, int64_t k, int metric, int64_t seed, int n_threads,
                          int scan_mode, float* centers, int64_t* chosen, orc_callback cb, void* user) {
    if (k > n || k < 1) return 2;  // std::invalid_argument upstream
    if (n_threads < 1) n_threads = 1;
    if (metric == ORC_MINRMSD) {
        if (d % 3) return 3;
        return kmpp_t<ORC_MINRMSD>(X, n, d, k, seed, n_threads, scan_mode, centers, chosen, cb, user);
    }
    return kmpp_t<ORC_EUCLIDEAN>(X, n, d, k, seed, n_threads, scan_mode, centers, chosen, cb, user);
}

// the raw RNG stream the product's own mt19937 restatement is checked against
ORC_API void orc_rng_stream(int64_t seed, int64_t n, uint64_t* first_index, float* unit_floats, int count) {
    std::mt19937 gen((uint32_t)seed);
    std::uniform_int_distribution<size_t> uni(0, (size_t)n - 1);
    *first_index = uni(gen);
    for (int i = 0; i < count; ++i) unit_floats[i] = (float)gen() / (float)gen.max();
}

// ---------------------------------------------------------------------------------
// A.5 regspace  (deeptime regspace.cluster(chunk, centers, dmin, max_centers, n_threads);
// call site regspace.py:150).  Frames in order; mindist = min_j compute(x_i,c_j) (+max
// when there are no centers); strictly `mindist > dmin` (dmin held in T=float) appends a
// copy of the frame; appending beyond max_centers raises MaxCentersReachedException
// (return code 4; centers found so far are kept, regspace.py:153-171).
// ---------------------------------------------------------------------------------
template <int METRIC>
static int regspace_t(const float* X, int64_t n, int64_t d, float* centers, int64_t* n_centers, float dmin,
                      int64_t max_centers, int n_threads, int64_t* frame_idx) {
    std::vector<Scratch> scr((size_t)n_threads);
    int64_t nc = *n_centers;
    for (int64_t i = 0; i < n; ++i) {
        float mind = std::numeric_limits<float>::max();
        if (n_threads > 1 && nc >= 256) {
#pragma omp parallel for reduction(min : mind) num_threads(n_threads)
            for (int64_t j = 0; j < nc; ++j) {
                const float dj = compute<METRIC>(X + i * d, centers + j * d, d, scr[omp_get_thread_num()]);
                if (dj < mind) mind = dj;
            }
        } else {
            for (int64_t j = 0; j < nc; ++j) {
                const float dj = compute<METRIC>(X + i * d, centers + j * d, d, scr[0]);
                if (dj < mind) mind = dj;
            }
        }
        if (mind > dmin) {
            if (nc + 1 > max_centers) { *n_centers = nc; return 4; }
            std::memcpy(centers + nc * d, X + i * d, sizeof(float) * d);
            if (frame_idx) frame_idx[nc] = i;
            nc++;
        }
    }
    *n_centers = nc;
    return 0;
}

ORC_API int orc_regspace(const float* X, int64_t n, int64_t d, float* centers, int64_t* n_centers, float dmin,
                         int64_t max_centers, int metric, int n_threads, int64_t* frame_idx) {
    if (n_threads < 1) n_threads = 1;
    if (metric == ORC_MINRMSD) {
        if (d % 3) return 3;
        return regspace_t<ORC_MINRMSD>(X, n, d, centers, n_centers, dmin, max_centers, n_threads, frame_idx);
    }
    return regspace_t<ORC_EUCLIDEAN>(X, n, d, centers, n_centers, dmin, max_centers, n_threads, frame_idx);
}

// all pair distances x_i vs c_j (used by tests for the "manual argmin" check, test_kmeans.py:246-252)
ORC_API int orc_pairwise(const float* X, int64_t n, int64_t d, const float* C, int64_t k, int metric, float* out) {
    Scratch s;
    if (metric == ORC_MINRMSD && d % 3) return 3;
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = 0; j < k; ++j)
            out[i * k + j] = (metric == ORC_MINRMSD) ? compute<ORC_MINRMSD>(X + i * d, C + j * d, d, s)
                                                     : compute<ORC_EUCLIDEAN>(X + i * d, C + j * d, d, s);
    return 0;
}

ORC_API const char* orc_build_info() {
    return "oracle restatement of deeptime/mdtraj path; g++ " __VERSION__
           " -O3 -fopenmp -ffp-contract=off; order=omp4; parity unpinned upstream";
}
