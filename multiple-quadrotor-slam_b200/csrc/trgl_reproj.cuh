// Fused reprojection-error kernels (grid-stride, deterministic two-stage reduction: warp shuffles -> per-block
// partial sums in global memory -> ordered sum on the host).
//   k_reproj_error : cv2.projectPoints + squared residuals      (calibration_tools.py:116-124, :89-113)
//   k_pair_reproj  : both normalised cameras + good-point mask  (triangulation_comparison.py:190-217,242-260;
//                    slam2.py:556,589 status filters)
#pragma once
#include "trgl_device.cuh"
#include "trgl_eval.cuh"
#include <cmath>

namespace trgl {

constexpr int kReduceBlocks = 148 * 4;

struct ProjParams {
    double R[9], t[3];
    double fx, fy, cx, cy;
    double k1, k2, p1, p2, k3;
};

// cv2.Rodrigues(rvec) on the host (one 3x3 per call)
inline ProjParams make_proj_params(const double* K, const double* dist, const double* rvec, const double* tvec) {
    ProjParams p;
    const double th = std::sqrt(rvec[0] * rvec[0] + rvec[1] * rvec[1] + rvec[2] * rvec[2]);
    if (th < 2.220446049250313e-16) {
        for (int i = 0; i < 9; ++i) p.R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    } else {
        const double kx = rvec[0] / th, ky = rvec[1] / th, kz = rvec[2] / th;
        const double c = std::cos(th), s = std::sin(th), c1 = 1.0 - c;
        p.R[0] = c + c1 * kx * kx;      p.R[1] = c1 * kx * ky - s * kz; p.R[2] = c1 * kx * kz + s * ky;
        p.R[3] = c1 * kx * ky + s * kz; p.R[4] = c + c1 * ky * ky;      p.R[5] = c1 * ky * kz - s * kx;
        p.R[6] = c1 * kx * kz - s * ky; p.R[7] = c1 * ky * kz + s * kx; p.R[8] = c + c1 * kz * kz;
    }
    for (int i = 0; i < 3; ++i) p.t[i] = tvec[i];
    p.fx = K[0]; p.fy = K[4]; p.cx = K[2]; p.cy = K[5];
    p.k1 = dist ? dist[0] : 0.0; p.k2 = dist ? dist[1] : 0.0; p.p1 = dist ? dist[2] : 0.0;
    p.p2 = dist ? dist[3] : 0.0; p.k3 = dist ? dist[4] : 0.0;
    return p;
}

template <typename TX, typename TP>
__global__ void __launch_bounds__(kThreads)
k_reproj_error(const TX* __restrict__ x, const TP* __restrict__ imgp, const __grid_constant__ ProjParams pp, TP* __restrict__ proj,
               double* __restrict__ partials, const int64_t n) {
    double acc[5] = {0, 0, 0, 0, 0};          // sum dx^2, sum dy^2, finite count, sum |dx|, sum |dy|
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * kThreads) {
        const double X = static_cast<double>(x[3 * i + 0]), Y = static_cast<double>(x[3 * i + 1]),
                     Z = static_cast<double>(x[3 * i + 2]);
        const double xc = fma(pp.R[0], X, fma(pp.R[1], Y, fma(pp.R[2], Z, pp.t[0])));
        const double yc = fma(pp.R[3], X, fma(pp.R[4], Y, fma(pp.R[5], Z, pp.t[1])));
        const double zc = fma(pp.R[6], X, fma(pp.R[7], Y, fma(pp.R[8], Z, pp.t[2])));
        const double iz = 1.0 / zc;
        const double a = xc * iz, b = yc * iz;
        const double r2 = fma(a, a, b * b);
        const double rad = fma(r2, fma(r2, fma(r2, pp.k3, pp.k2), pp.k1), 1.0);
        const double xd = fma(a, rad, fma(2.0 * pp.p1 * a, b, pp.p2 * fma(2.0 * a, a, r2)));
        const double yd = fma(b, rad, fma(pp.p1, fma(2.0 * b, b, r2), 2.0 * pp.p2 * a * b));
        const TP u = static_cast<TP>(fma(pp.fx, xd, pp.cx)), v = static_cast<TP>(fma(pp.fy, yd, pp.cy));
        if (proj) { proj[2 * i + 0] = u; proj[2 * i + 1] = v; }
        // like the reference, the residual is formed from the projected points in their storage type
        const double dx = static_cast<double>(u) - static_cast<double>(imgp[2 * i + 0]);
        const double dy = static_cast<double>(v) - static_cast<double>(imgp[2 * i + 1]);
        acc[0] = fma(dx, dx, acc[0]);
        acc[1] = fma(dy, dy, acc[1]);
        acc[2] += (fabs(dx) <= DBL_MAX && fabs(dy) <= DBL_MAX) ? 1.0 : 0.0;
        acc[3] += fabs(dx);
        acc[4] += fabs(dy);
    }
    block_reduce_store<5>(acc, partials);
}

template <typename TI, typename TO, typename TS>
__global__ void __launch_bounds__(kThreads)
k_pair_reproj(const TO* __restrict__ x, const TI* __restrict__ u1, const TI* __restrict__ u2, const __grid_constant__ Cams<double> cams,
              const TS* __restrict__ status, const int min_status, const double max_sq_err, TO* __restrict__ err1,
              TO* __restrict__ err2, uint8_t* __restrict__ good, double* __restrict__ partials, const int64_t n,
              unsigned int* __restrict__ counter, double* __restrict__ final_out) {
    double acc[4] = {0, 0, 0, 0};             // sum err1 (good), sum err2 (good), #good, #status > min_status
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * kThreads) {
        const double X = static_cast<double>(x[3 * i + 0]), Y = static_cast<double>(x[3 * i + 1]),
                     Z = static_cast<double>(x[3 * i + 2]);
        double a1, b1, a2, b2;
        load_uv<double>(u1, i, a1, b1);
        load_uv<double>(u2, i, a2, b2);
        eval_point<TO>(cams, min_status, max_sq_err, a1, b1, a2, b2, X, Y, Z, static_cast<int>(status[i]), i, err1, err2, good, acc);
    }
    block_reduce_finalize<4>(acc, partials, counter, final_out);
}

// ---- harness statistics on the device (triangulation_comparison.py:179-188, 205-217, 242-260) ---------------------
// Squared 3-D error per point against the exact cloud, with the reductions error_rms and robustness_stat need:
// partials per block = sum of errors, #NaN errors, #false positives (error > thresh_max and status > 0),
// #false negatives (error <= thresh_min and not status > 0).
template <typename TO, typename TS>
__global__ void __launch_bounds__(kThreads)
k_sq_errors_3d(const TO* __restrict__ x, const double* __restrict__ exact, const int exact_stride,
               const TS* __restrict__ status, const double thresh_max, const double thresh_min,
               double* __restrict__ errors, double* __restrict__ partials, const int64_t n) {
    double acc[4] = {0, 0, 0, 0};
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * kThreads) {
        // error_vectors_3D: points_3D_calc - points_3D_exact[:, 0:3]; error_rms: sum(v**2, axis=1) = (dx^2 + dy^2) + dz^2
        const double dx = static_cast<double>(x[3 * i + 0]) - exact[exact_stride * i + 0];
        const double dy = static_cast<double>(x[3 * i + 1]) - exact[exact_stride * i + 1];
        const double dz = static_cast<double>(x[3 * i + 2]) - exact[exact_stride * i + 2];
        const double e = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        if (errors) errors[i] = e;
        acc[0] += e;
        if (e != e) acc[1] += 1.0;
        if (status) {
            const bool est = static_cast<int>(status[i]) > 0;
            if (!(e <= thresh_max) && est) acc[2] += 1.0;          // (errors <= max) == False  and  positives_est
            if ((e <= thresh_min) && !est) acc[3] += 1.0;
        }
    }
    block_reduce_store<4>(acc, partials);
}

// vector_stat of the comparison harness (triangulation_comparison.py:219-240): per point, the mean vector and the
// (population) covariance matrix of its 3-D error vectors x[t, i, :] - exact[i, 0:3] over the `trials` repetitions.
// One thread per point, two passes over the trials like the reference (means first, then the deviations from them);
// consecutive threads read consecutive points of one trial, so every pass is a coalesced stream over x (trials, n, 3).
// means: (n, 3), covars: (n, 3, 3) row-major doubles.
template <typename TO>
__global__ void __launch_bounds__(kThreads)
k_vector_stat(const TO* __restrict__ x, const double* __restrict__ exact, const int exact_stride, const int trials,
              double* __restrict__ means, double* __restrict__ covars, const int64_t n) {
    const double inv = 1.0 / static_cast<double>(trials);
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * kThreads) {
        const double ex = exact[exact_stride * i + 0], ey = exact[exact_stride * i + 1], ez = exact[exact_stride * i + 2];
        double mx = 0, my = 0, mz = 0;
        for (int t = 0; t < trials; ++t) {
            const TO* p = x + (static_cast<int64_t>(t) * n + i) * 3;
            mx += static_cast<double>(p[0]) - ex; my += static_cast<double>(p[1]) - ey; mz += static_cast<double>(p[2]) - ez;
        }
        mx *= inv; my *= inv; mz *= inv;
        double c[6] = {0, 0, 0, 0, 0, 0};              // xx xy xz yy yz zz
        for (int t = 0; t < trials; ++t) {
            const TO* p = x + (static_cast<int64_t>(t) * n + i) * 3;
            const double dx = (static_cast<double>(p[0]) - ex) - mx, dy = (static_cast<double>(p[1]) - ey) - my,
                         dz = (static_cast<double>(p[2]) - ez) - mz;
            c[0] = fma(dx, dx, c[0]); c[1] = fma(dx, dy, c[1]); c[2] = fma(dx, dz, c[2]);
            c[3] = fma(dy, dy, c[3]); c[4] = fma(dy, dz, c[4]); c[5] = fma(dz, dz, c[5]);
        }
        means[3 * i + 0] = mx; means[3 * i + 1] = my; means[3 * i + 2] = mz;
        double* o = covars + 9 * i;
        o[0] = c[0] * inv; o[1] = c[1] * inv; o[2] = c[2] * inv;
        o[3] = c[1] * inv; o[4] = c[3] * inv; o[5] = c[4] * inv;
        o[6] = c[2] * inv; o[7] = c[4] * inv; o[8] = c[5] * inv;
    }
}

// Squared norm of (n,2) residual vectors proj - exact (error_rms on error_vectors_2D), same partials[0..1].
template <typename TP>
__global__ void __launch_bounds__(kThreads)
k_sq_errors_2d(const TP* __restrict__ proj, const double* __restrict__ exact, double* __restrict__ errors,
               double* __restrict__ partials, const int64_t n) {
    double acc[4] = {0, 0, 0, 0};
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * kThreads) {
        const double dx = static_cast<double>(proj[2 * i + 0]) - exact[2 * i + 0];
        const double dy = static_cast<double>(proj[2 * i + 1]) - exact[2 * i + 1];
        const double e = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        if (errors) errors[i] = e;
        acc[0] += e;
        if (e != e) acc[1] += 1.0;
    }
    block_reduce_store<4>(acc, partials);
}

// Exact order statistics of non-negative doubles (np.median semantics) by most-significant-digit radix selection: the
// IEEE-754 bit pattern of a non-negative double is monotone as an unsigned integer.  One pass histograms the 8-bit digit
// at `shift` of every key whose higher digits equal `prefix` (shared-memory privatised, one global atomic per bin/block).
__global__ void __launch_bounds__(kThreads)
k_radix_hist(const double* __restrict__ v, const int64_t n, const unsigned long long prefix, const unsigned long long mask,
             const int shift, unsigned long long* __restrict__ hist /* 256 bins + [256] = NaN count (first pass) */) {
    __shared__ unsigned int h[256];
    h[threadIdx.x] = 0;                       // kThreads == 256
    __syncthreads();
    unsigned int nans = 0;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * kThreads) {
        const double val = v[i];
        const unsigned long long key = static_cast<unsigned long long>(__double_as_longlong(val));
        if ((key & mask) == prefix) atomicAdd(&h[(key >> shift) & 255ull], 1u);
        if (val != val) ++nans;
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], static_cast<unsigned long long>(h[threadIdx.x]));
    if (mask == 0ull && nans) atomicAdd(&hist[256], static_cast<unsigned long long>(nans));
}

// Smallest key strictly greater than `key0` (for the upper middle element of an even-sized sample).
__global__ void __launch_bounds__(kThreads)
k_min_above(const double* __restrict__ v, const int64_t n, const unsigned long long key0, unsigned long long* __restrict__ out) {
    unsigned long long best = ~0ull;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * kThreads) {
        const unsigned long long key = static_cast<unsigned long long>(__double_as_longlong(v[i]));
        if (key > key0 && key < best) best = key;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_down_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
    }
    if ((threadIdx.x & 31) == 0 && best != ~0ull) atomicMin(out, best);
}

// ---- whole-batch reductions for the normalised 8-point fundamental matrix (triangulation.py:228) ---------------
// stage 0: sum x1,y1,x2,y2   stage 1: sum |p1-m1|, |p2-m2|   stage 2: the 45 unique entries of A^T A, where the row of
// A for one match is (x2x1, x2y1, x2, y2x1, y2y1, y2, x1, y1, 1) in normalised coordinates.
struct F8Params { double m1[2], m2[2], s1, s2; };

template <typename TI, int STAGE>
__global__ void __launch_bounds__(kThreads)
k_f8_reduce(const TI* __restrict__ u1, const TI* __restrict__ u2, const __grid_constant__ F8Params fp, double* __restrict__ partials,
            const int64_t n) {
    constexpr int NV = STAGE == 0 ? 4 : (STAGE == 1 ? 2 : 45);
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = 0.0;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * kThreads) {
        double x1, y1, x2, y2;
        load_uv<double>(u1, i, x1, y1);
        load_uv<double>(u2, i, x2, y2);
        if constexpr (STAGE == 0) {
            acc[0] += x1; acc[1] += y1; acc[2] += x2; acc[3] += y2;
        } else if constexpr (STAGE == 1) {
            const double a = x1 - fp.m1[0], b = y1 - fp.m1[1], c = x2 - fp.m2[0], d = y2 - fp.m2[1];
            acc[0] += sqrt(fma(a, a, b * b));
            acc[1] += sqrt(fma(c, c, d * d));
        } else {
            const double a = (x1 - fp.m1[0]) * fp.s1, b = (y1 - fp.m1[1]) * fp.s1;
            const double c = (x2 - fp.m2[0]) * fp.s2, d = (y2 - fp.m2[1]) * fp.s2;
            const double r[9] = {c * a, c * b, c, d * a, d * b, d, a, b, 1.0};
            int k = 0;
#pragma unroll
            for (int p = 0; p < 9; ++p)
#pragma unroll
                for (int q = p; q < 9; ++q) { acc[k] = fma(r[p], r[q], acc[k]); ++k; }
        }
    }
    block_reduce_store<NV>(acc, partials);
}

}  // namespace trgl
