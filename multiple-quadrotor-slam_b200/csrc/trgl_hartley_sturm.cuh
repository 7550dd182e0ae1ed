// Hartley-Sturm optimal correction of one correspondence (what cv2.correctMatches computes per point;
// call site Work/python_libs/triangulation.py:224, algorithm H&Z 12.1 / SURVEY.md Appendix A.10), in registers.
//
// Root selection.  The reference evaluates the cost s(t) at the real part of all six roots of the degree-6
// polynomial g(t) (numerator of s'(t)) and at t = inf, and keeps the minimiser; i.e. it returns the GLOBAL
// minimiser of s over the reals.  We get the same t two ways:
//   fast path  -- s(t) >= t^2/(1+f1^2 t^2), so every t with s(t) <= s(0) lies in |t| < T0 with
//                 T0^2 = s(0)/(1 - f1^2 s(0)); if an interval bound shows g' > 0 on [-T0,T0] then g has exactly one
//                 root there, it is the global minimiser, and a bracketed Newton iteration from t = 0 finds it
//                 (2-5 iterations).  This certificate holds for every point of the sideways / rotating / general
//                 rigs and for most points of the forward-motion rig.
//   middle     -- otherwise (follow-up kernel): every real root of g by certified bisection (hs_select_isolate), then the
//                 reference's cost scan over them;
//   slow path  -- what that cannot certify either: Durand-Kerner on all six complex roots exactly as cv::solvePoly is
//                 driven by correctMatches (leading coefficients <= DBL_EPSILON dropped, start (1+i)^k, Gauss-Seidel
//                 sweeps, <= 100 iterations), then the reference's cost scan over the real parts.
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include "trgl_device.cuh"

namespace trgl {

// Diagnostics of the rare paths of the correction (trgl_rare_path_counters): points through the certified real-root
// isolation, intervals it visited, points it gave up on (= points through Durand-Kerner after it), Durand-Kerner points in
// total, most intervals visited for one point.  One atomic per such point, in the follow-up kernel only.
__device__ unsigned long long g_hs_counters[5];

struct HSParams {
    double F[9];    // row-major, x2^T F x1 = 0
    double e1[3];   // right epipole, F e1 = 0
    double e2[3];   // left epipole, F^T e2 = 0
};

__device__ __forceinline__ double hs_cost(double t, double a, double b, double c, double d, double f1, double f2) {
    const double ct_d = fma(c, t, d), at_b = fma(a, t, b);
    return t * t / fma(f1 * f1 * t, t, 1.0) + ct_d * ct_d / fma(at_b, at_b, f2 * f2 * ct_d * ct_d);
}

// k[0..6] ascending:  g(t) = t q(t)^2 - (ad-bc) (1+f1^2 t^2)^2 (at+b)(ct+d),  q = (at+b)^2 + f2^2 (ct+d)^2
__device__ __forceinline__ void hs_coeffs(double a, double b, double c, double d, double f1, double f2, double k[7]) {
    const double f1s = f1 * f1, f2s = f2 * f2;
    const double q2 = fma(a, a, f2s * c * c), q1 = 2.0 * fma(a, b, f2s * c * d), q0 = fma(b, b, f2s * d * d);
    const double e = fma(a, d, -b * c);
    const double r2 = a * c, r1 = fma(a, d, b * c), r0 = b * d;
    const double w2 = 2.0 * f1s, w4 = f1s * f1s;
    k[0] = -e * r0;
    k[1] = fma(q0, q0, -e * r1);
    k[2] = fma(2.0 * q0, q1, -e * fma(w2, r0, r2));
    k[3] = fma(q1, q1, 2.0 * q0 * q2) - e * (w2 * r1);
    k[4] = fma(2.0 * q1, q2, -e * fma(w2, r2, w4 * r0));
    k[5] = fma(q2, q2, -e * (w4 * r1));
    k[6] = -e * (w4 * r2);
}

// Durand-Kerner + cost scan, restating cv::solvePoly(maxIters=100) and the selection loop of correctMatches.
// Returns t_min, or DBL_MAX when t = inf has the lowest cost.
__device__ __noinline__ double hs_select_dk(const double k[7], double a, double b, double c, double d,
                                            double f1, double f2) {
    atomicAdd(&g_hs_counters[3], 1ull);
    int n = 6;
    if (!(fabs(k[6]) > DBL_EPSILON)) { n = 5;
      if (!(fabs(k[5]) > DBL_EPSILON)) { n = 4;
        if (!(fabs(k[4]) > DBL_EPSILON)) { n = 3;
          if (!(fabs(k[3]) > DBL_EPSILON)) { n = 2;
            if (!(fabs(k[2]) > DBL_EPSILON)) n = 1; } } } }
    double kk[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) kk[i] = (i <= n) ? k[i] : 0.0;
    const double lead = n == 6 ? k[6] : n == 5 ? k[5] : n == 4 ? k[4] : n == 3 ? k[3] : n == 2 ? k[2] : k[1];
    double zr[6], zi[6];
    {
        double pr = 1.0, pi = 0.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            zr[i] = pr; zi[i] = pi;
            const double nr = pr - pi, ni = pr + pi;      // * (1 + i)
            pr = nr; pi = ni;
        }
    }
    for (int iter = 0; iter < 100; ++iter) {
        double maxdiff = 0.0;
        bool settled = true;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            if (i < n) {
                const double pr = zr[i], pi = zi[i];
                double nr = 0.0, ni = 0.0;                  // Horner over all 7 slots (leading slots are 0)
#pragma unroll
                for (int j = 6; j >= 0; --j) {
                    const double tr = fma(nr, pr, -ni * pi) + kk[j];
                    const double ti = fma(nr, pi, ni * pr);
                    nr = tr; ni = ti;
                }
                double dr = lead, di = 0.0;
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    if (j != i && j < n) {
                        const double er = pr - zr[j], ei = pi - zi[j];
                        if (er != 0.0 || ei != 0.0) {
                            const double tr = fma(dr, er, -di * ei);
                            const double ti = fma(dr, ei, di * er);
                            dr = tr; di = ti;
                        }
                    }
                }
                const double den = fma(dr, dr, di * di);
                const double sr = fma(nr, dr, ni * di) / den;
                const double si = fma(ni, dr, -nr * di) / den;
                zr[i] = pr - sr; zi[i] = pi - si;
                const double mag = sqrt(fma(sr, sr, si * si));
                maxdiff = fmax(maxdiff, mag);
                if (mag > 2.5e-16 * (fabs(zr[i]) + fabs(zi[i]))) settled = false;
            }
        }
        if (!(maxdiff > 0.0) || settled) break;          // NaN-safe; `settled` only skips no-op sweeps
    }
    double s_val = 1.0 / (f1 * f1) + c * c / fma(a, a, f2 * f2 * c * c);
    double t_min = DBL_MAX;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        if (i < n) {
            const double s = hs_cost(zr[i], a, b, c, d, f1, f2);
            if (s < s_val) { s_val = s; t_min = zr[i]; }
        }
    }
    return t_min;
}

// Middle tier of the follow-up kernel: ALL real roots of g by certified bisection, then the reference's cost scan over them.
// Why that is the reference's answer: correctMatches scans the real PARTS of all six roots (and t = inf); the real part of
// a complex root is just some real t, and s(t) at any real t is at least the global minimum of s over the reals, which is
// attained at a real root of g (or at infinity) -- so the scan returns the global real minimiser, and complex roots can
// only tie.  Durand-Kerner needs all of its 100 sweeps on these polynomials (g = t q(t)^2 - small: the two complex root
// pairs of q are nearly double and never settle), ~57 000 instructions per point; the real roots are simple.
//   t in [-R, R], R = min(T0, 1)  on g itself, and  u = 1/t in [-1, 1]  on the reversed polynomial when T0 > 1 or the
//   bound T0 does not exist (f1^2 s(0) >= 1: 11 % of the forward-motion rig at 8 px).  Dyadic subdivision; per interval a
//   Taylor shift to its midpoint (21 FMA) and two tests on the shifted coefficients c_j, half width h (hs_interval_test):
//     |c0| > sum_{j>=1} |c_j| h^j            -> no root in the interval;
//     |c1| > sum_{j>=2} j |c_j| h^(j-1)      -> g' keeps its sign: at most one root, present iff the end values differ in
//                                               sign -> recorded, refined later by a bracketed Newton iteration in the
//                                               shifted variable (hs_refine_root);
//     neither                                 -> split (depth <= 23, else give up -> Durand-Kerner).
//   Measured on 6000 points per rig / noise level against a CPU restatement of cv::solvePoly's selection (NumPy
//   restatement of these functions): no mismatch over 1e-9, no give-up; on 10 M points of the forward-motion rig at 0.8 px:
//   110 540 points, 18.6 intervals per point on average, 188 at most, none given up (trgl_rare_path_counters).
// The subdivision is level-synchronous and spread over the CTA (k_polynomial_general): one lane per point walked a
// heavy-tailed number of intervals while the other 31 waited (ncu: 3 of 32 lanes active on average).
constexpr int kIsoMaxDepth = 23;                       // interval width 2 R / 2^23
constexpr int kIsoMaxRoots = 8;                        // <= 6 real roots, +2 for the shared end points t = +-1 <-> u = +-1

__device__ __forceinline__ void hs_interval_geometry(double R, int depth, unsigned int pos, double& mid, double& h) {
    const double w = ldexp(2.0 * R, -depth);
    h = 0.5 * w;
    mid = fma(w, static_cast<double>(pos), -R) + h;
}

// 0: no root of p in the interval; 1: exactly one (a sign change on an interval where p' keeps its sign); 2: undecided.
__device__ __forceinline__ int hs_interval_test(const double p[7], double R, int depth, unsigned int pos) {
    double mid, h;
    hs_interval_geometry(R, depth, pos, mid, h);
    double cj[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) cj[j] = p[j];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 5; j >= i; --j) cj[j] = fma(mid, cj[j + 1], cj[j]);
    double rest0 = 0.0, noise = 0.0, rest1 = 0.0;
#pragma unroll
    for (int j = 6; j >= 1; --j) rest0 = (rest0 + fabs(cj[j])) * h;
#pragma unroll
    for (int j = 6; j >= 0; --j) noise = fma(noise, fabs(mid), fabs(p[j]));
#pragma unroll
    for (int j = 6; j >= 2; --j) rest1 = fma(rest1, h, j * fabs(cj[j]));
    rest1 *= h;
    double glo = cj[6], ghi = cj[6];
#pragma unroll
    for (int j = 5; j >= 0; --j) { glo = fma(glo, -h, cj[j]); ghi = fma(ghi, h, cj[j]); }
    const bool excluded = fabs(cj[0]) > fma(rest0, 1.0 + 1e-9, 1e-13 * noise);
    const bool monotone = fabs(cj[1]) > rest1 * (1.0 + 1e-9);
    if (excluded) return 0;
    if (!monotone) return 2;
    return (((glo < 0.0) != (ghi < 0.0)) || glo == 0.0 || ghi == 0.0) ? 1 : 0;
}

// The root of p in a recorded interval (p monotone there): bracketed Newton in the shifted variable.
__device__ __forceinline__ double hs_refine_root(const double p[7], double mid, double h) {
    double cj[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) cj[j] = p[j];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 5; j >= i; --j) cj[j] = fma(mid, cj[j + 1], cj[j]);
    double glo = cj[6];
#pragma unroll
    for (int j = 5; j >= 0; --j) glo = fma(glo, -h, cj[j]);
    double xl = -h, xh = h;
    double x = fmin(fmax(-cj[0] / cj[1], -h), h);
#pragma unroll 1
    for (int it = 0; it < 64; ++it) {
        double g = cj[6], dg = 0.0;
#pragma unroll
        for (int j = 5; j >= 0; --j) { dg = fma(dg, x, g); g = fma(g, x, cj[j]); }
        if (g == 0.0) break;
        if ((g < 0.0) == (glo < 0.0)) xl = x; else xh = x;
        double xn = x - g / dg;
        if (xn == x) break;
        const bool newton = (xn >= xl && xn <= xh);
        if (!newton) xn = 0.5 * (xl + xh);
        const double step = fabs(xn - x);
        x = xn;
        if ((newton && step <= 1e-8 * fabs(mid + xn)) || (xh - xl) <= 4e-16 * fabs(mid + xn)) break;
    }
    return mid + x;
}

// The reference's cost scan over the recorded roots of one point (packed: depth << 24 | domain << 23 | position).  The
// order in which a CTA records them is not fixed, so ties between finite candidates go to the smaller t.
// Returns false when the scan cannot stand (no root inside a finite T0).
__device__ __noinline__ bool hs_scan_roots(const double k[7], const unsigned int* packed, int nroots, double R0, bool bounded,
                                           double a, double b, double c, double d, double f1, double f2, double& t_out) {
    double s_val = 1.0 / (f1 * f1) + c * c / fma(a, a, f2 * f2 * c * c);
    double t_min = DBL_MAX;
    int found = 0;
#pragma unroll 1
    for (int r = 0; r < nroots; ++r) {
        const unsigned int item = packed[r];
        const int depth = static_cast<int>(item >> 24), dom = static_cast<int>((item >> 23) & 1u);
        double p[7], mid, h;
#pragma unroll
        for (int j = 0; j < 7; ++j) p[j] = dom ? k[6 - j] : k[j];
        hs_interval_geometry(dom ? 1.0 : R0, depth, item & 0x7fffffu, mid, h);
        double t = hs_refine_root(p, mid, h);
        if (dom) {
            if (t == 0.0) continue;                    // u = 0 is t = inf, already in s_val
            t = 1.0 / t;
        }
        ++found;
        const double sv = hs_cost(t, a, b, c, d, f1, f2);
        if (sv < s_val || (sv == s_val && t_min != DBL_MAX && t < t_min)) { s_val = sv; t_min = t; }
    }
    if (bounded && found == 0) return false;
    t_out = t_min;
    return true;
}

__device__ __forceinline__ double hs_rsqrt(double x) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double h = x * -0.5;
    r = fma(r, fma(h * r, r, 0.5), r);
    r = fma(r, fma(h * r, r, 0.5), r);
    return r;
}

__device__ __forceinline__ void hs_epipole(const double e[3], double x, double y, double& ex, double& ey, double& f) {
    ex = fma(-x, e[2], e[0]);
    ey = fma(-y, e[2], e[1]);
    f = e[2];
    // MUFU.RSQ64H seed + two Newton steps (full precision for normal arguments, no special-case path);
    // 0 -> inf -> NaN epipole -> NaN point, like the reference
    const double inv = hs_rsqrt(fma(ex, ex, ey * ey));
    ex *= inv; ey *= inv; f *= inv;
    if (f < 0.0) { ex = -ex; ey = -ey; f = -f; }
}

// What the correction of one correspondence keeps between its phases: set-up (everything up to the fast-path
// certificate), root selection (t), closest points.
struct HsPoint {
    double x1, y1, x2, y2;
    double e1x, e1y, f1, e2x, e2y, f2;      // translated, rotated epipoles
    double a, b, c, d;                      // entries of F''
    double T0;                              // every t with s(t) <= s(0) lies in [-T0, T0]  (only if `bounded`)
    bool finite_coeffs, bounded, fast;      // fast: g' > 0 on [-T0, T0] shown by the interval bound
};

__device__ __forceinline__ void hs_setup(const HSParams& hs, double x1, double y1, double x2, double y2, HsPoint& P,
                                         double k[7]) {
    const double* F = hs.F;
    P.x1 = x1; P.y1 = y1; P.x2 = x2; P.y2 = y2;
    // F' = T2^-T F T1^-1  (both points moved to the origin)
    double Fp[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        Fp[r][0] = F[3 * r + 0];
        Fp[r][1] = F[3 * r + 1];
        Fp[r][2] = fma(F[3 * r + 0], x1, fma(F[3 * r + 1], y1, F[3 * r + 2]));
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) Fp[2][q] = fma(x2, Fp[0][q], fma(y2, Fp[1][q], Fp[2][q]));
    // epipoles of F' are the translated epipoles of F
    hs_epipole(hs.e1, x1, y1, P.e1x, P.e1y, P.f1);
    hs_epipole(hs.e2, x2, y2, P.e2x, P.e2y, P.f2);
    // F'' = R2 F' R1^T, entries (1,1) (1,2) (2,1) (2,2)
    const double h01 = fma(-Fp[0][0], P.e1y, Fp[0][1] * P.e1x);
    const double h11 = fma(-Fp[1][0], P.e1y, Fp[1][1] * P.e1x);
    const double h21 = fma(-Fp[2][0], P.e1y, Fp[2][1] * P.e1x);
    P.a = fma(-P.e2y, h01, P.e2x * h11);
    P.b = fma(-P.e2y, Fp[0][2], P.e2x * Fp[1][2]);
    P.c = h21;
    P.d = Fp[2][2];
    hs_coeffs(P.a, P.b, P.c, P.d, P.f1, P.f2, k);
    P.finite_coeffs = true;
#pragma unroll
    for (int i = 0; i < 7; ++i) P.finite_coeffs = P.finite_coeffs && (fabs(k[i]) <= DBL_MAX);
    // ---- fast path certificate ----
    const double s0 = P.d * P.d * fast_rcp(fma(P.b, P.b, P.f2 * P.f2 * P.d * P.d));   // 0/0 -> NaN -> certificate fails
    const double f1s = P.f1 * P.f1;
    P.bounded = f1s * s0 < 1.0;
    P.fast = P.finite_coeffs && P.bounded;
    P.T0 = 0.0;
    if (P.fast) {
        // any upper bound of T0 keeps the certificate valid: x * rsqrt(x), inflated by 1e-9, instead of the IEEE sqrt
        const double T0s = s0 * fast_rcp(1.0 - f1s * s0);
        P.T0 = (T0s > 0.0) ? T0s * hs_rsqrt(T0s) * (1.0 + 1e-9) : T0s;
        const double T0 = P.T0;
        const double bound = T0 * fma(T0, fma(T0, fma(T0, fma(T0, 6.0 * fabs(k[6]), 5.0 * fabs(k[5])),
                                                      4.0 * fabs(k[4])), 3.0 * fabs(k[3])), 2.0 * fabs(k[2]));
        P.fast = k[1] > bound;
    }
}

// Bracketed Newton on g, strictly increasing on [-T0, T0] (certificate).  Newton converges quadratically, so once a NEWTON
// step is below 1e-8 |t| the iterate it produced is exact to rounding (error ~ step^2); bisection steps only stop on a
// 2-ulp bracket.  (Demanding a 2-ulp Newton step made single lanes bisect for ~50 rounds.)
__device__ __forceinline__ double hs_newton_bracketed(const double k[7], double T0) {
    double lo = -T0, hi = T0;
    double t = 0.0;
#pragma unroll 1
    for (int it = 0; it < 64; ++it) {
        double g = k[6], dg = 0.0;
#pragma unroll
        for (int i = 5; i >= 0; --i) { dg = fma(dg, t, g); g = fma(g, t, k[i]); }
        if (g == 0.0) break;
        if (g < 0.0) lo = t; else hi = t;
        double tn = fma(-g, fast_rcp(dg), t);          // g' > 0 on the bracket (certificate)
        if (tn == t) break;                            // Newton step below half an ulp: converged
        const bool newton = (tn >= lo && tn <= hi);
        if (!newton) tn = 0.5 * (lo + hi);
        const double step = fabs(tn - t);
        t = tn;
        if ((newton && step <= 1e-8 * fabs(tn)) || (hi - lo) <= 4e-16 * fabs(tn)) break;
    }
    return t;
}

// Closest points to the origin on the two epipolar lines, then back through R^T and T^-1.  t == DBL_MAX: t = inf wins (or
// non-finite system) -- the reference evaluates inf/inf -> NaN for both points.
__device__ __forceinline__ void hs_finish(const HsPoint& P, double t, double& n1x, double& n1y, double& n2x, double& n2y) {
    const bool at_inf = (t == DBL_MAX);
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    {
        const double iz = fast_rcp(fma(t * t, P.f1 * P.f1, 1.0));
        const double hx = t * t * P.f1 * iz, hy = t * iz;
        n1x = at_inf ? qnan : fma(P.e1x, hx, -P.e1y * hy) + P.x1;
        n1y = at_inf ? qnan : fma(P.e1y, hx, P.e1x * hy) + P.y1;
    }
    {
        const double ct_d = fma(P.c, t, P.d), at_b = fma(P.a, t, P.b);
        const double hz = fma(P.f2 * P.f2 * ct_d, ct_d, at_b * at_b);
        const double iz = (hz > 0.0) ? fast_rcp(hz) : 1.0 / hz;
        const double hx = P.f2 * ct_d * ct_d * iz, hy = -at_b * ct_d * iz;
        n2x = at_inf ? qnan : fma(P.e2x, hx, -P.e2y * hy) + P.x2;
        n2y = at_inf ? qnan : fma(P.e2y, hx, P.e2x * hy) + P.y2;
    }
}

// The certified fast path of the correction, for the hot kernel -- returns false, leaving the outputs undefined, when the
// certificate does not hold or the fixed-step Newton iteration is not accepted; the caller then defers the point.
// (k_polynomial_general runs the complete correction from the same pieces: hs_setup, the selection tiers, hs_finish.)
__device__ __forceinline__ bool hs_correct_fast(const HSParams& hs, double x1, double y1, double x2, double y2,
                                                double& n1x, double& n1y, double& n2x, double& n2y) {
    HsPoint P;
    double k[7];
    hs_setup(hs, x1, y1, x2, y2, P, k);
    double t = DBL_MAX;
    bool certified = true;      // single exit below: a `return` inside the branches would keep the lanes that leave the
                                // Newton loop at different rounds apart until the end of the function (measured: the
                                // closest-point epilogue then ran once per exit group, +85 FP64 instructions per point)
    const bool fast = P.fast;
    const double T0 = P.T0;
    {
        // Hot kernel (all 32 lanes of the warp are here): plain Newton from t = 0 with a FIXED number of steps and one
        // acceptance test at the end -- no per-lane exit, no bracket bookkeeping; lanes whose certificate failed run along
        // on garbage and are discarded.  g(0) = k0, g'(0) = k1, so the first iterate is free; two more (one Horner pass
        // each) satisfy the test for every point of the translating / rotating rigs and 99 % of the general rig at
        // 0.8 - 8 px; further ones are taken by the whole warp while any of its certified lanes needs them (warp-uniform
        // branch).  Accepted: the last Newton step is below 1e-8 |t| (the iterate it produced is then exact to rounding,
        // error ~ step^2) and the iterate lies in [-T0, T0], where the certificate has shown g to have exactly one root --
        // whatever path the iteration took to get there.  Everything else is deferred.
        double tp = -k[0] * fast_rcp(k[1]);
        bool accepted = false;
        auto newton_step = [&]() {
            double g = k[6], dg = 0.0;
#pragma unroll
            for (int i = 5; i >= 0; --i) { dg = fma(dg, tp, g); g = fma(g, tp, k[i]); }
            const double tn = fma(-g, fast_rcp(dg), tp);
            if (!accepted) {        // an accepted lane keeps its iterate: the result must not depend on how long its warp goes on
                t = tn;
                accepted = (fabs(tn - tp) <= 1e-8 * fabs(tn)) && (fabs(tn) <= T0);
                tp = tn;
            }
        };
        newton_step();              // two steps straight-line (nobody is accepted before the second), then the warp votes
        newton_step();
#pragma unroll 1
        for (int it = 2; it < 4; ++it) {
            if (__all_sync(0xffffffffu, accepted || !fast)) break;
            newton_step();
        }
        certified = fast && accepted;
        if (!certified) t = 0.0;
    }
    hs_finish(P, t, n1x, n1y, n2x, n2y);
    return certified;
}

}  // namespace trgl
