// __global__ kernels of libtriangl_cuda: one thread per correspondence, PPT correspondences per thread issued
// back to back so that several 16-byte loads are in flight per thread (memory-level parallelism on HBM3e).
#pragma once
#include "trgl_device.cuh"
#include "trgl_hartley_sturm.cuh"
#include "trgl_tma.cuh"
#include "trgl_eval.cuh"

#ifndef TRGL_EIGEN_MINB
#define TRGL_EIGEN_MINB 3          // resident CTAs per SM the register allocation is sized for (76-80 registers)
#endif
#ifndef TRGL_EVAL_MINB
#define TRGL_EVAL_MINB 2           // linear_eigen / polynomial with the evaluation epilogue: measured faster with 128 registers
#endif
#ifndef TRGL_POLY_MINB
#define TRGL_POLY_MINB 3
#endif
#ifndef TRGL_LS_MINB
#define TRGL_LS_MINB 3             // linear_LS hot kernel without the in-line SVD tier: 3 CTAs/SM of loads in flight
#endif
#ifndef TRGL_LS_EVAL_MINB
#define TRGL_LS_EVAL_MINB 2        // linear_LS with the evaluation epilogue, 4 points per thread: 80 registers spill (184 B)
#endif
#ifndef TRGL_ITER_MINB
#define TRGL_ITER_MINB 3          // 80 registers with the evaluation epilogue, no spills
#endif

namespace trgl {

// ---- two-ray closed form of the re-weighted solve --------------------------------------------------------------
// The system of one correspondence is rows (a0, a1) of view 1 scaled by w1 and rows (c0, c1) of view 2 scaled by w2
// (triangulation.c:30-40,143-146).  Each view's two planes meet in its viewing ray  C_k + t n_k  (n1 = a0 x a1,
// C_k = camera centre, the null vector of P_k), so M_k = A_k^T A_k has rank 2, adj(M_k) = n_k n_k^T, and by
// Cauchy-Binet the normal-equation solution of the weighted system is, exactly,
//     x(kappa) = (X1 + kappa X2) / (1 + kappa),      kappa = (w2/w1)^2 * B/A,
// where X1 = C1 + (t1/A) n1 is the point of ray 1 that minimises the view-2 residuals, X2 the point of ray 2 that
// minimises the view-1 residuals, A = (c0.n1)^2 + (c1.n1)^2, B = (a0.n2)^2 + (a1.n2)^2, t1 = -sum_j (c_j.C1 - b_j)(c_j.n1).
// Only the weight RATIO moves during the iteration, and the depths are the same convex combination of the depths of
// X1 and X2, so one re-weighting round costs ~24 FP64 instructions instead of a fresh 3x3 solve (~80), with error
// ~ kappa(A) eps (no squaring: agrees with the SVD solve to < 1e-10 up to kappa^2 = 1e12, see tests).
template <typename T> struct RayGeom { T C1[3], C2[3], E1[3], E2[3]; int ok; int pad_; };   // E1 = P1 [C2;1], E2 = P2 [C1;1]
template <typename T> struct TwoRay { T X1[3], X2[3], kap, d11, d12, d21, d22; };
constexpr int kTwoRayState = 13;            // kap d1 d2 d11 d12 d21 d22 X1[3] X2[3]

// kappa in [2^-27, 2^27]: the squared ratio of the weighted row norms of the two views.  Outside (or NaN) the reference's
// SVD is ill-conditioned by the imbalance itself and the point goes to the general path.
__device__ __forceinline__ bool kappa_in_range(double kap) {
    const unsigned e = (static_cast<unsigned>(__double2hiint(kap)) >> 20) & 0xfffu;     // sign + exponent
    return (e - (1023u - 27u)) <= 54u;
}


template <typename T>
__device__ __forceinline__ bool tworay_setup(const Cams<T>& cams, const RayGeom<T>& g, T u1x, T u1y, T u2x, T u2y,
                                             TwoRay<T>& R) {
    T a0[3], a1[3], c0[3], c1[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        a0[k] = tfma(u1x, cams.P1[8 + k], -cams.P1[k]);  a1[k] = tfma(u1y, cams.P1[8 + k], -cams.P1[4 + k]);
        c0[k] = tfma(u2x, cams.P2[8 + k], -cams.P2[k]);  c1[k] = tfma(u2y, cams.P2[8 + k], -cams.P2[4 + k]);
    }
    const T n1[3] = {tfma(a0[1], a1[2], -a0[2] * a1[1]), tfma(a0[2], a1[0], -a0[0] * a1[2]), tfma(a0[0], a1[1], -a0[1] * a1[0])};
    const T n2[3] = {tfma(c0[1], c1[2], -c0[2] * c1[1]), tfma(c0[2], c1[0], -c0[0] * c1[2]), tfma(c0[0], c1[1], -c0[1] * c1[0])};
    const T s0 = tfma(c0[0], n1[0], tfma(c0[1], n1[1], c0[2] * n1[2])), s1 = tfma(c1[0], n1[0], tfma(c1[1], n1[1], c1[2] * n1[2]));
    const T r0 = tfma(a0[0], n2[0], tfma(a0[1], n2[1], a0[2] * n2[2])), r1 = tfma(a1[0], n2[0], tfma(a1[1], n2[1], a1[2] * n2[2]));
    const T A = tfma(s0, s0, s1 * s1), B = tfma(r0, r0, r1 * r1);
    // residuals of the other view's planes at this view's camera centre: c_j.C1 - b_j = u2 * E2[2] - E2[j]
    const T e0 = tfma(u2x, g.E2[2], -g.E2[0]), e1 = tfma(u2y, g.E2[2], -g.E2[1]);
    const T f0 = tfma(u1x, g.E1[2], -g.E1[0]), f1 = tfma(u1y, g.E1[2], -g.E1[1]);
    const T t1 = -tfma(e0, s0, e1 * s1), t2 = -tfma(f0, r0, f1 * r1);
    T tr = a0[0] * a0[0];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (k) tr = tfma(a0[k], a0[k], tr);
        tr = tfma(a1[k], a1[k], tr); tr = tfma(c0[k], c0[k], tr); tr = tfma(c1[k], c1[k], tr);
    }
    // det(M1 + M2) = A + B; the same kappa^2 bound as the normal-equation tiers, at the tier-2 limit
    const bool well = (tr * tr * tr < Tiers<T>::t2() * (A + B)) && (A > T(0)) && (B > T(0));
    const T iA = fast_rcp(A), iB = fast_rcp(B);
    const T tau1 = t1 * iA, tau2 = t2 * iB;
#pragma unroll
    for (int k = 0; k < 3; ++k) { R.X1[k] = tfma(tau1, n1[k], g.C1[k]); R.X2[k] = tfma(tau2, n2[k], g.C2[k]); }
    R.kap = B * iA;
    R.d11 = tfma(cams.P1[8], R.X1[0], tfma(cams.P1[9], R.X1[1], tfma(cams.P1[10], R.X1[2], cams.P1[11])));
    R.d12 = tfma(cams.P1[8], R.X2[0], tfma(cams.P1[9], R.X2[1], tfma(cams.P1[10], R.X2[2], cams.P1[11])));
    R.d21 = tfma(cams.P2[8], R.X1[0], tfma(cams.P2[9], R.X1[1], tfma(cams.P2[10], R.X1[2], cams.P2[11])));
    R.d22 = tfma(cams.P2[8], R.X2[0], tfma(cams.P2[9], R.X2[1], tfma(cams.P2[10], R.X2[2], cams.P2[11])));
    return well && kappa_in_range(static_cast<double>(R.kap));
}

// Intersection of the two viewing rays, for matches that satisfy the epipolar constraint (the output of the Hartley-Sturm
// correction): the DLT matrix then has an exact null vector, and the smallest singular vector cv2.triangulatePoints
// returns, dehomogenised, IS the least-squares point (X1 A + X2 B) / (A + B) of the two-ray form (they differ by
// O(residual^2 / gap); agrees with the SVD solve to < 2e-10 on all rigs, see tests).  Returns false -- the caller then runs
// the eigen solver -- unless (i) the system is well conditioned (same kappa^2 bound as above) and (ii) each ray passes
// through the other view's planes: squared residual of view 2 at X1 = ee - t1^2/A <= res_tol * ee, likewise for view 1.
template <typename T>
__device__ __forceinline__ bool tworay_intersection(const Cams<T>& cams, const RayGeom<T>& g, T u1x, T u1y, T u2x, T u2y,
                                                    T res_tol, T x[3]) {
    T a0[3], a1[3], c0[3], c1[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        a0[k] = tfma(u1x, cams.P1[8 + k], -cams.P1[k]);  a1[k] = tfma(u1y, cams.P1[8 + k], -cams.P1[4 + k]);
        c0[k] = tfma(u2x, cams.P2[8 + k], -cams.P2[k]);  c1[k] = tfma(u2y, cams.P2[8 + k], -cams.P2[4 + k]);
    }
    const T n1[3] = {tfma(a0[1], a1[2], -a0[2] * a1[1]), tfma(a0[2], a1[0], -a0[0] * a1[2]), tfma(a0[0], a1[1], -a0[1] * a1[0])};
    const T n2[3] = {tfma(c0[1], c1[2], -c0[2] * c1[1]), tfma(c0[2], c1[0], -c0[0] * c1[2]), tfma(c0[0], c1[1], -c0[1] * c1[0])};
    const T s0 = tfma(c0[0], n1[0], tfma(c0[1], n1[1], c0[2] * n1[2])), s1 = tfma(c1[0], n1[0], tfma(c1[1], n1[1], c1[2] * n1[2]));
    const T r0 = tfma(a0[0], n2[0], tfma(a0[1], n2[1], a0[2] * n2[2])), r1 = tfma(a1[0], n2[0], tfma(a1[1], n2[1], a1[2] * n2[2]));
    const T A = tfma(s0, s0, s1 * s1), B = tfma(r0, r0, r1 * r1);
    const T e0 = tfma(u2x, g.E2[2], -g.E2[0]), e1 = tfma(u2y, g.E2[2], -g.E2[1]);
    const T f0 = tfma(u1x, g.E1[2], -g.E1[0]), f1 = tfma(u1y, g.E1[2], -g.E1[1]);
    const T t1 = -tfma(e0, s0, e1 * s1), t2 = -tfma(f0, r0, f1 * r1);
    T tr = a0[0] * a0[0];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (k) tr = tfma(a0[k], a0[k], tr);
        tr = tfma(a1[k], a1[k], tr); tr = tfma(c0[k], c0[k], tr); tr = tfma(c1[k], c1[k], tr);
    }
    const T D = A + B;
    const T eeA = tfma(e0, e0, e1 * e1) * A, ffB = tfma(f0, f0, f1 * f1) * B;
    const bool ok = (tr * tr * tr < Tiers<T>::t2() * D) &&
                    (tfma(-t1, t1, eeA) <= res_tol * eeA) && (tfma(-t2, t2, ffB) <= res_tol * ffB);
    const T inv = fast_rcp(D);
#pragma unroll
    for (int k = 0; k < 3; ++k) x[k] = tfma(A, g.C1[k], tfma(t1, n1[k], tfma(B, g.C2[k], t2 * n2[k]))) * inv;
    return ok;
}

// One re-weighting round in closed form (triangulation.c:128-148).  Returns 1 when the reference's loop breaks here,
// 0 to continue, -1 when the point has to be redone by the general path (exact-zero depth under the Python control
// flow, weight ratio out of range).  `kap_used`, `inv` = kappa and 1 / (1 + kappa) of the round that was evaluated: the
// solve of that round is x = (X1 + kap_used X2) * inv.
template <typename T>
__device__ __forceinline__ int tworay_step(const T d11, const T d12, const T d21, const T d22, T& kap, T& d1, T& d2,
                                           T& d1n, T& d2n, T& inv, T& kap_used, const T tolerance, const int py_semantics) {
    kap_used = kap;
    inv = fast_rcp(T(1) + kap);
    d1n = tfma(kap, d12, d11) * inv;                                                                  // triangulation.c:133
    d2n = tfma(kap, d22, d21) * inv;
    const bool conv = (tabs(d1n - d1) <= tolerance) && (tabs(d2n - d2) <= tolerance);
    const bool zero = (d1n == T(0)) || (d2n == T(0));                                                 // triangulation.c:138
    if (conv || (zero && !py_semantics)) return 1;
    if (zero) return -1;
    const T r = d1n * fast_rcp(d2n);                // cumulative re-weighting: (w2/w1)^2 *= (d1n/d2n)^2, triangulation.c:143-146
    kap *= r * r;
    d1 = d1n; d2 = d2n;
    return kappa_in_range(static_cast<double>(kap)) ? 0 : -1;
}

// Uncertified points leave the hot kernels through a deferred list in global memory and are solved by a small follow-up
// kernel (k_iterative_general / k_linear_eigen_general) with the reference's arithmetic as written.  Keeping that
// ~120-register path -- and the subroutine calls of its SVD tier -- out of the hot kernels is what lets them run without
// a single spill and at 3 CTAs per SM: with the call inside, ptxas kept the loop state of the persistent loop in local
// memory (profiles/r01f).
// First instruction of every follow-up kernel: wait until the grid this one depends on (the hot kernel, launched just before
// on the same stream) has completed and flushed its writes.  A no-op when the kernel was launched without the
// programmatic-dependent-launch attribute.
__device__ __forceinline__ void wait_for_hot_kernel() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

struct Deferred {
    int64_t* idx;                 // [cap] indices of the points to redo
    unsigned int* ctl;            // ctl[0] = number of deferred points (may exceed cap), ctl[1] = ticket of the follow-up kernel,
                                  // ctl[2..3] = 64-bit running total of deferred points (diagnostics)
    unsigned int cap;
    unsigned int* hint;           // page-locked host word (may be null): the follow-up kernel leaves ctl[0] there, the host sizes
                                  // the NEXT follow-up grid of this solver with it
};
__device__ __forceinline__ void defer_point(const Deferred& df, int64_t i) {
    const unsigned int k = atomicAdd(&df.ctl[0], 1u);
    if (k < df.cap) df.idx[k] = i;            // on overflow the follow-up kernel redoes every point instead
}
// Warp-aggregated append (callable from divergent code): one atomic for the lanes that are here together.  Which of the
// two forms a kernel uses was decided by measurement: linear_LS runs at 0.93 of the copy peak with this one and at 0.84
// with the plain atomic (same registers, same occupancy -- the plain read-modify-write sits between the four
// interleaved points of a thread), the FP64-bound kernels are 2-3 % slower with it.
__device__ __forceinline__ void defer_point_warp(const Deferred& df, int64_t i) {
    // Once the list has overflowed the follow-up kernel redoes every point anyway: no more appends.  (Same-address atomics
    // run at ~1.3 G/s: a batch that defers all of its 10^8 points spent 2.4 ms on 3 M warp-level appends.)
    if (*reinterpret_cast<volatile unsigned int*>(df.ctl) > df.cap) return;
    const unsigned mask = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    unsigned int base = 0;
    if (lane == leader) base = atomicAdd(&df.ctl[0], static_cast<unsigned int>(__popc(mask)));
    base = __shfl_sync(mask, base, leader);
    const unsigned int k = base + __popc(mask & ((1u << lane) - 1u));
    if (k < df.cap) df.idx[k] = i;
}


// ---- linear_LS_triangulation (triangulation.c:65-83) -----------------------------------------------------------
// Per point: straight-line fast solve, (rare) careful redo, immediate coalesced store.  Measured on B200: storing each
// point as soon as it is solved beats "solve all PPT points, then store" by 0.81 vs 0.65 of the HBM peak, and keeping
// the 4x4 row block alive for an inline refinement path costs 2x (0.39).
// DEFER: a point beyond tier 1 is not redone in line (an out-of-line call in the hot kernel, run by one or two lanes of
// the warp) but appended to the deferred list; k_linear_ls_general redoes it with the same solve_point_careful.
template <typename TI, typename TC, typename TO, int PPT, class PRE, bool EVAL, class MIR, bool DEFER>
__device__ __forceinline__ void ls_tile(const TI* __restrict__ u1, const TI* __restrict__ u2, const Cams<TC>& cams,
                                        TO* __restrict__ x, uint8_t* __restrict__ status, const int64_t n, const PRE& pre,
                                        const MIR& mir, const EvalArg<EVAL>& ev, const Deferred& df, const int64_t block_base,
                                        TO* __restrict__ stage_warp, double (&acc)[4]) {
    const int warp = threadIdx.x >> 5;
    TC in[PPT][4];
#pragma unroll
    for (int p = 0; p < PPT; ++p) {
        const int64_t i = block_base + p * kThreads + threadIdx.x;
        if (i < n) {
            load_uv<TC>(u1, i, in[p][0], in[p][1]);
            load_uv<TC>(u2, i, in[p][2], in[p][3]);
        } else {
            in[p][0] = in[p][1] = in[p][2] = in[p][3] = TC(0);
        }
    }
#pragma unroll
    for (int p = 0; p < PPT; ++p) {
        const int64_t i = block_base + p * kThreads + threadIdx.x;
        TC xs[3];
        if constexpr (PRE::kActive) pre.template apply<TI, TC>(in[p][0], in[p][1], in[p][2], in[p][3]);
        bool done = true;
        if (!ls_point_fast<TC>(cams, in[p][0], in[p][1], in[p][2], in[p][3], xs)) {
            if constexpr (DEFER) { done = false; if (i < n) defer_point_warp(df, i); }
            else solve_point_careful<TC>(cams, in[p][0], in[p][1], in[p][2], in[p][3], TC(1), TC(1), xs);
        }
        store_x_warp(x, block_base + p * kThreads + warp * 32, n, static_cast<TO>(xs[0]),
                         static_cast<TO>(xs[1]), static_cast<TO>(xs[2]), stage_warp, mir);
        if (i < n) store_status(status, mir, i, static_cast<uint8_t>(1));
        fused_eval_point<EVAL, TO, TC>(ev, (i < n) && done, i, in[p][0], in[p][1], in[p][2], in[p][3], xs, 1, acc);
    }
}

template <typename TI, typename TC, typename TO, int PPT, class PRE = PreNone, bool EVAL = false, class MIR = Mirrors,
          bool DEFER = false>
__global__ void __launch_bounds__(kThreads, DEFER ? (EVAL ? TRGL_LS_EVAL_MINB : TRGL_LS_MINB) : 1)
k_linear_ls(const TI* __restrict__ u1, const TI* __restrict__ u2, const __grid_constant__ Cams<TC> cams,
            TO* __restrict__ x, uint8_t* __restrict__ status, const int64_t n, const __grid_constant__ PRE pre,
            const __grid_constant__ MIR mir, const __grid_constant__ EvalArg<EVAL> ev, const __grid_constant__ Deferred df) {
    __shared__ TO stage[kWarps][96];
    double acc[4] = {0, 0, 0, 0};
    if constexpr (EVAL) {
        // grid-stride loop over tiles (capped grid): the evaluation sums are reduced per CTA at the end of the kernel
        const int64_t stride = static_cast<int64_t>(gridDim.x) * (kThreads * PPT);
        for (int64_t base = static_cast<int64_t>(blockIdx.x) * (kThreads * PPT); base < n; base += stride)
            ls_tile<TI, TC, TO, PPT, PRE, EVAL, MIR, DEFER>(u1, u2, cams, x, status, n, pre, mir, ev, df, base, stage[threadIdx.x >> 5], acc);
        fused_eval_finish<EVAL>(ev, acc, DEFER ? df.ctl : nullptr, df.cap);
    } else {
        // one tile per CTA
        ls_tile<TI, TC, TO, PPT, PRE, EVAL, MIR, DEFER>(u1, u2, cams, x, status, n, pre, mir, ev, df,
                                                        static_cast<int64_t>(blockIdx.x) * (kThreads * PPT), stage[threadIdx.x >> 5], acc);
    }
}

// ---- linear_LS, FP32 mode: four CONSECUTIVE points per thread, 128-bit loads and stores --------------------------------
// At 29 bytes per point the FP32 mode is bound by instruction issue, not by HBM, unless the per-point overhead goes: here a
// thread loads its four (x,y) pairs of each view as two float4, solves them with float32 normal equations (no refinement:
// tier-1 points only -- kappa^2 bound below 300; the rest is deferred to k_linear_ls_general<float, double, float>, which
// redoes it in float64 on the float32 inputs: e.g. every point of the low-parallax forward-motion rig), and writes its 12 result floats as three float4
// and its four status bytes as one 32-bit word -- no shared-memory transposition, no per-point address arithmetic.
// Needs 16-byte aligned u1 / u2 / x and 4-byte aligned status (the launcher checks); the last < 4 points take scalar accesses.
#ifndef TRGL_F32_REDO
#define TRGL_F32_REDO 0
#endif
// Cold path of the FP32 mode (TRGL_F32_REDO): a point beyond float32 tier 1 redone with float64 normal equations in place.
__device__ __noinline__ bool ls_point_f32_redo_in_double(const Cams<double>& camsd, float a, float b, float c, float d, float* x) {
    double xd[3];
    const bool ok = ls_point_fast<double>(camsd, static_cast<double>(a), static_cast<double>(b), static_cast<double>(c),
                                          static_cast<double>(d), xd);
    x[0] = static_cast<float>(xd[0]); x[1] = static_cast<float>(xd[1]); x[2] = static_cast<float>(xd[2]);
    return ok;
}

__global__ void __launch_bounds__(kThreads, 4)
k_linear_ls_f32x4(const float* __restrict__ u1, const float* __restrict__ u2, const __grid_constant__ Cams<float> cams,
                  const __grid_constant__ Cams<double> camsd, float* __restrict__ x, uint8_t* __restrict__ status,
                  const int64_t n, const __grid_constant__ Deferred df) {
    const int64_t i0 = (static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x) * 4;
    if (i0 >= n) return;
    float in1[8], in2[8], xs[4][3];
    const bool full = i0 + 4 <= n;
    if (full) {
        const float4 a0 = __ldcs(reinterpret_cast<const float4*>(u1 + 2 * i0)), a1 = __ldcs(reinterpret_cast<const float4*>(u1 + 2 * i0) + 1);
        const float4 b0 = __ldcs(reinterpret_cast<const float4*>(u2 + 2 * i0)), b1 = __ldcs(reinterpret_cast<const float4*>(u2 + 2 * i0) + 1);
        in1[0] = a0.x; in1[1] = a0.y; in1[2] = a0.z; in1[3] = a0.w; in1[4] = a1.x; in1[5] = a1.y; in1[6] = a1.z; in1[7] = a1.w;
        in2[0] = b0.x; in2[1] = b0.y; in2[2] = b0.z; in2[3] = b0.w; in2[4] = b1.x; in2[5] = b1.y; in2[6] = b1.z; in2[7] = b1.w;
    } else {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const bool in = i0 + p < n;
            in1[2 * p] = in ? u1[2 * (i0 + p)] : 0.f; in1[2 * p + 1] = in ? u1[2 * (i0 + p) + 1] : 0.f;
            in2[2 * p] = in ? u2[2 * (i0 + p)] : 0.f; in2[2 * p + 1] = in ? u2[2 * (i0 + p) + 1] : 0.f;
        }
    }
    bool ok[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) ok[p] = ls_point_plain_f32(cams, in1[2 * p], in1[2 * p + 1], in2[2 * p], in2[2 * p + 1], xs[p]);
#if TRGL_F32_REDO
    if (!(ok[0] && ok[1] && ok[2] && ok[3])) {
#pragma unroll
        for (int p = 0; p < 4; ++p)
            if (!ok[p]) ok[p] = ls_point_f32_redo_in_double(camsd, in1[2 * p], in1[2 * p + 1], in2[2 * p], in2[2 * p + 1], xs[p]);
    }
#endif
    if (full) {
        float4* dst = reinterpret_cast<float4*>(x + 3 * i0);
        __stcs(dst + 0, make_float4(xs[0][0], xs[0][1], xs[0][2], xs[1][0]));
        __stcs(dst + 1, make_float4(xs[1][1], xs[1][2], xs[2][0], xs[2][1]));
        __stcs(dst + 2, make_float4(xs[2][2], xs[3][0], xs[3][1], xs[3][2]));
        *reinterpret_cast<uint32_t*>(status + i0) = 0x01010101u;
    } else {
#pragma unroll
        for (int p = 0; p < 4; ++p)
            if (i0 + p < n) {
                x[3 * (i0 + p)] = xs[p][0]; x[3 * (i0 + p) + 1] = xs[p][1]; x[3 * (i0 + p) + 2] = xs[p][2];
                status[i0 + p] = 1;
            }
    }
    // No more appends once the list has overflowed (the follow-up kernel then redoes every point anyway): same-address
    // atomics run at ~1.3 G/s, and a rig that defers every point (translating, forward motion) spent 2.4 of this kernel's
    // 2.9 ms per 100 M points on them.  (Aggregating the four appends of a thread into one per warp, in line or out of
    // line, cost the rigs that defer nothing 6-9 %: measured, profiles/README.md.)
    if (!(ok[0] && ok[1] && ok[2] && ok[3]) && *reinterpret_cast<volatile unsigned int*>(df.ctl) <= df.cap) {
#pragma unroll
        for (int p = 0; p < 4; ++p)
            if (!ok[p] && i0 + p < n) defer_point(df, i0 + p);
    }
}

// ---- linear_LS, per-thread cp.async ring ---------------------------------------------------------------------------
// Persistent CTAs; every thread keeps DEPTH of its own future (x,y) pairs in flight with cp.async (LDGSTS) into its
// private shared-memory slots, one commit group per tile.  No block barrier anywhere (a thread only reads what it
// copied), no registers spent on lookahead, and the bytes in flight per SM are DEPTH x 8 KB x resident CTAs regardless
// of where the warps are in their solve.
template <typename TI, typename TC, typename TO, int PPT, int DEPTH, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
k_linear_ls_ring(const TI* __restrict__ u1, const TI* __restrict__ u2, const __grid_constant__ Cams<TC> cams,
                 TO* __restrict__ x, uint8_t* __restrict__ status, const int64_t n, const __grid_constant__ Mirrors mir) {
    constexpr int TILE = kThreads * PPT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TI* ring = reinterpret_cast<TI*>(smem_raw);                                   // [DEPTH][2 views][TILE*2]
    TO* stage = reinterpret_cast<TO*>(smem_raw + static_cast<size_t>(DEPTH) * 2 * TILE * 2 * sizeof(TI));
    const int warp = threadIdx.x >> 5;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * TILE;
    int64_t tile = static_cast<int64_t>(blockIdx.x) * TILE;
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) {
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            const int j = p * kThreads + threadIdx.x;
            const int64_t i = tile + d * stride + j;
            if (i < n) {
                cp_async_pair(ring + ((d * 2 + 0) * TILE + j) * 2, u1 + 2 * i);
                cp_async_pair(ring + ((d * 2 + 1) * TILE + j) * 2, u2 + 2 * i);
            }
        }
        cp_async_commit();
    }
    int slot = 0;
    for (; tile < n; tile += stride) {
        cp_async_wait_group<DEPTH - 1>();                                        // the oldest group has landed
        TC in[PPT][4];
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            const int j = p * kThreads + threadIdx.x;
            const int64_t i = tile + j;
            TI* s1 = ring + ((slot * 2 + 0) * TILE + j) * 2;
            TI* s2 = ring + ((slot * 2 + 1) * TILE + j) * 2;
            in[p][0] = in[p][1] = in[p][2] = in[p][3] = TC(0);
            if (i < n) {
                in[p][0] = static_cast<TC>(s1[0]); in[p][1] = static_cast<TC>(s1[1]);
                in[p][2] = static_cast<TC>(s2[0]); in[p][3] = static_cast<TC>(s2[1]);
            }
            const int64_t inext = i + DEPTH * stride;                            // refill the slot just consumed
            if (inext < n) { cp_async_pair(s1, u1 + 2 * inext); cp_async_pair(s2, u2 + 2 * inext); }
        }
        cp_async_commit();
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            const int64_t i = tile + p * kThreads + threadIdx.x;
            TC xs[3];
            if (!ls_point_fast<TC>(cams, in[p][0], in[p][1], in[p][2], in[p][3], xs))
                solve_point_careful<TC>(cams, in[p][0], in[p][1], in[p][2], in[p][3], TC(1), TC(1), xs);
            store_x_warp(x, tile + p * kThreads + warp * 32, n, static_cast<TO>(xs[0]), static_cast<TO>(xs[1]),
                             static_cast<TO>(xs[2]), stage + warp * 96, mir);
            if (i < n) store_status(status, mir, i, static_cast<uint8_t>(1));
        }
        if (++slot == DEPTH) slot = 0;
    }
    cp_async_wait_all();
}

// ---- linear_LS, bulk-async pipelined variant ------------------------------------------------------------------------
// Persistent CTAs; tile = kThreads*PPT correspondences; ring of STAGES shared-memory stages filled by cp.async.bulk.
// Dynamic shared memory layout: [STAGES][2][TILE*2] TI  |  kWarps*96 TO (store staging)  |  STAGES mbarriers.
template <typename TI, typename TC, typename TO, int PPT, int STAGES, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
k_linear_ls_tma(const TI* __restrict__ u1, const TI* __restrict__ u2, const __grid_constant__ Cams<TC> cams,
                TO* __restrict__ x, uint8_t* __restrict__ status, const int64_t n, const __grid_constant__ Mirrors mir) {
    constexpr int TILE = kThreads * PPT;
    constexpr uint32_t kTileBytes = TILE * 2 * sizeof(TI);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TI* in = reinterpret_cast<TI*>(smem_raw);                                  // [STAGES][2][TILE*2]
    TO* stage_out = reinterpret_cast<TO*>(smem_raw + static_cast<size_t>(STAGES) * 2 * kTileBytes);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + static_cast<size_t>(STAGES) * 2 * kTileBytes
                                                 + kWarps * 96 * sizeof(TO));
    const int warp = threadIdx.x >> 5;
    const int64_t ntiles = (n + TILE - 1) / TILE;
    const int64_t nfull = n / TILE;                                            // tiles [0, nfull) are complete
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            const int64_t t = static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(s) * gridDim.x;
            if (t < nfull) {
                mbar_expect_tx(&full[s], 2 * kTileBytes);
                bulk_g2s(in + (s * 2 + 0) * TILE * 2, u1 + t * TILE * 2, kTileBytes, &full[s]);
                bulk_g2s(in + (s * 2 + 1) * TILE * 2, u2 + t * TILE * 2, kTileBytes, &full[s]);
            }
        }
    }
    int s = 0;
    uint32_t parity = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t tile_base = t * TILE;
        const bool is_full = t < nfull;
        if (is_full) mbar_wait(&full[s], parity);
        const TI* s1 = in + (s * 2 + 0) * TILE * 2;
        const TI* s2 = in + (s * 2 + 1) * TILE * 2;
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            const int j = p * kThreads + threadIdx.x;
            const int64_t i = tile_base + j;
            TC a = 0, b = 0, c = 0, d = 0;
            if (is_full) {
                if constexpr (sizeof(TI) == 8) {
                    const double2 v1 = reinterpret_cast<const double2*>(s1)[j], v2 = reinterpret_cast<const double2*>(s2)[j];
                    a = static_cast<TC>(v1.x); b = static_cast<TC>(v1.y); c = static_cast<TC>(v2.x); d = static_cast<TC>(v2.y);
                } else {
                    const float2 v1 = reinterpret_cast<const float2*>(s1)[j], v2 = reinterpret_cast<const float2*>(s2)[j];
                    a = static_cast<TC>(v1.x); b = static_cast<TC>(v1.y); c = static_cast<TC>(v2.x); d = static_cast<TC>(v2.y);
                }
            } else if (i < n) {                 // ragged last tile: plain loads
                load_uv<TC>(u1, i, a, b); load_uv<TC>(u2, i, c, d);
            }
            TC xs[3];
            if (!ls_point_fast<TC>(cams, a, b, c, d, xs))
                solve_point_careful<TC>(cams, a, b, c, d, TC(1), TC(1), xs);
            store_x_warp(x, tile_base + p * kThreads + warp * 32, n, static_cast<TO>(xs[0]),
                             static_cast<TO>(xs[1]), static_cast<TO>(xs[2]), stage_out + warp * 96, mir);
            if (i < n) store_status(status, mir, i, static_cast<uint8_t>(1));
        }
        __syncthreads();                        // every thread has consumed stage s
        if (threadIdx.x == 0) {
            const int64_t tn = t + static_cast<int64_t>(STAGES) * gridDim.x;
            if (tn < nfull) {
                mbar_expect_tx(&full[s], 2 * kTileBytes);
                bulk_g2s(in + (s * 2 + 0) * TILE * 2, u1 + tn * TILE * 2, kTileBytes, &full[s]);
                bulk_g2s(in + (s * 2 + 1) * TILE * 2, u2 + tn * TILE * 2, kTileBytes, &full[s]);
            }
        }
        if (++s == STAGES) { s = 0; parity ^= 1; }
    }
}

// ---- iterative_LS_triangulation (triangulation.c:104-161 / triangulation.py:100-195) ---------------------------
// One re-weighted solve + depth test.  Returns true when the reference's loop breaks at this iteration.
template <typename TC>
__device__ __forceinline__ bool iter_ls_step(const Cams<TC>& cams, TC a, TC b, TC c, TC d, const TC M1[6], const TC v1[3],
                                             const TC M2[6], const TC v2[3], TC& w1, TC& w2, TC& d1, TC& d2, TC& d1n,
                                             TC& d2n, TC xs[3], TC tolerance, int py_semantics) {
    solve_blocks<TC>(cams, a, b, c, d, M1, v1, M2, v2, w1, w2, xs);
    d1n = tfma(cams.P1[8], xs[0], tfma(cams.P1[9], xs[1], tfma(cams.P1[10], xs[2], cams.P1[11])));    // triangulation.c:133
    d2n = tfma(cams.P2[8], xs[0], tfma(cams.P2[9], xs[1], tfma(cams.P2[10], xs[2], cams.P2[11])));
    const bool conv = (tabs(d1n - d1) <= tolerance) && (tabs(d2n - d2) <= tolerance);
    const bool zero = !py_semantics && ((d1n == TC(0)) || (d2n == TC(0)));                             // triangulation.c:138
    if (conv || zero) return true;
    // cumulative re-weighting, triangulation.c:143-146 (1/0 keeps its IEEE meaning for the Python control flow)
    w1 *= (d1n != TC(0)) ? fast_rcp(d1n) : TC(1) / d1n;
    w2 *= (d2n != TC(0)) ? fast_rcp(d2n) : TC(1) / d2n;
    d1 = d1n; d2 = d2n;
    return false;
}

__device__ __forceinline__ int iter_ls_status(int it, double d1n, double d2n) {       // triangulation.c:154-159
    int st = (it < 10) && (d1n > 0.0) && (d2n > 0.0);
    if (d1n <= 0.0) st -= 1;
    if (d2n <= 0.0) st -= 2;
    return st;
}

// General path of one correspondence: the reference's loop as written (re-weighted normal equations with the
// conditioning tiers of solve_blocks).  Taken by the points the closed form below does not certify (low parallax,
// cameras without a finite centre, non-finite input, exact-zero depths under the Python control flow); instantiated at
// one place only, in the queue drain, where almost nothing else is live.
template <typename TC>
__device__ __forceinline__ int iter_point_general(const Cams<TC>& cams, TC a, TC b, TC c, TC d, TC tolerance, int py_semantics,
                                               TC xs[3], TC& d1n, TC& d2n) {
    TC M1[6], v1[3], M2[6], v2[3];
    point_blocks<TC>(cams, a, b, c, d, M1, v1, M2, v2);
    TC w1 = 1, w2 = 1, d1 = 1, d2 = 1;
    d1n = 1; d2n = 1;
    int it = py_semantics ? 9 : 10;                 // value of the loop variable after a loop that never breaks
#pragma unroll 1
    for (int k = 0; k < 10; ++k)
        if (iter_ls_step<TC>(cams, a, b, c, d, M1, v1, M2, v2, w1, w2, d1, d2, d1n, d2n, xs, tolerance, py_semantics)) {
            it = k; break;
        }
    return it;
}

// Persistent CTAs (one grid-stride loop over 256-point tiles, the next tile's four input scalars prefetched with cp.async
// while the current tile is solved); inside a CTA every WARP is autonomous -- it owns the 32-point slice of each tile, a
// FIFO work queue in shared memory for its divergent tail, and its own store staging, so the kernel has no block barrier
// and no atomic:
//   phase 1: every lane sets up the two-ray form of its point and runs the first kPhase1 rounds (on translating rigs
//            every point converges at the second solve; on rotating rigs ~70 % do).  Finished points leave through the
//            coalesced store.
//   queue  : lanes still running push their closed-form state (13 scalars) to the warp's circular queue (slot = ballot
//            prefix; head / count are warp-uniform registers).
//   phase 2: whenever the queue holds >= 32 entries the warp runs the remaining rounds on the OLDEST 32, so the warp is
//            dense instead of idling on the ~30 % of lanes that need all 10 solves; the remainder is flushed after the
//            last tile.
//   mirrors: with result mirrors (multi-GPU gather) a 32-point slice is copied to the peers, fully coalesced, as soon as
//            no point of it is queued any more -- FIFO order makes that "every slice before the one of the oldest queued
//            entry" -- instead of repeating phase 2's scattered 8-byte stores over NVLink.
#ifndef TRGL_ITER_PHASE1
#define TRGL_ITER_PHASE1 2
#endif
constexpr int kPhase1 = TRGL_ITER_PHASE1;
constexpr int kWarpQueueCap = 64;                        // per warp: <= 31 left over + 32 pushed; power of two

// Shared memory of one k_iterative_ls CTA (dynamic: above the static limit in the all-double mode).
template <typename TI, typename TC, typename TO>
struct IterSmem {
    PairPrefetch<TI, kThreads> pre;                      // cp.async landing zone of the next tile
    TC q_state[kWarps][kTwoRayState][kWarpQueueCap];     // per-warp queue: closed-form state of the point
    int64_t q_idx[kWarps][kWarpQueueCap];                //                 global point index
    TO stage[kWarps][96];                                // coalesced (n,3) store staging
};

// The four input scalars of point i again (phase 2 needs them for the evaluation epilogue and for the general path):
// they were read a few tiles ago, so this is an L2 hit.
template <typename TI, typename TC, class PRE>
__device__ __forceinline__ void reload_inputs(const TI* __restrict__ u1, const TI* __restrict__ u2, const PRE& pre_stage,
                                              int64_t i, TC& a, TC& b, TC& c, TC& d) {
    load_uv<TC>(u1, i, a, b);
    load_uv<TC>(u2, i, c, d);
    if constexpr (PRE::kActive) pre_stage.template apply<TI, TC>(a, b, c, d);
}

template <typename TI, typename TC, typename TO, class PRE, bool EVAL>
__device__ __forceinline__ void iter_ls_phase2(const TI* __restrict__ u1, const TI* __restrict__ u2, const PRE& pre_stage,
                                               const Cams<TC>& cams, const TC (*q_state)[kWarpQueueCap], const int64_t* q_idx,
                                               int slot, TO* __restrict__ x, int32_t* __restrict__ status,
                                               TC tolerance, int py_semantics, const EvalArg<EVAL>& ev, const Deferred& df,
                                               double (&acc)[4]) {
    TC kap = q_state[0][slot], d1 = q_state[1][slot], d2 = q_state[2][slot];
    const int64_t dst = q_idx[slot];
    int r = -1;
    {
        TC a = 0, b = 0, c = 0, d = 0;
        if constexpr (EVAL) reload_inputs<TI, TC, PRE>(u1, u2, pre_stage, dst, a, b, c, d);     // in flight during the rounds
        const TC d11 = q_state[3][slot], d12 = q_state[4][slot], d21 = q_state[5][slot], d22 = q_state[6][slot];
        TC d1n = d1, d2n = d2, inv = 1, kap_used = kap;
        int it = py_semantics ? 9 : 10;             // value of the loop variable after a loop that never breaks
#pragma unroll 1
        for (int k = kPhase1; k < 10; ++k) {
            r = tworay_step<TC>(d11, d12, d21, d22, kap, d1, d2, d1n, d2n, inv, kap_used, tolerance, py_semantics);
            if (r) { if (r > 0) it = k; break; }
        }
        if (r >= 0) {                               // the last solve that was evaluated is the result (triangulation.c:130,150)
            TC xs[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) xs[k] = tfma(kap_used, q_state[10 + k][slot], q_state[7 + k][slot]) * inv;
            x[3 * dst + 0] = static_cast<TO>(xs[0]);
            x[3 * dst + 1] = static_cast<TO>(xs[1]);
            x[3 * dst + 2] = static_cast<TO>(xs[2]);
            const int st = iter_ls_status(it, static_cast<double>(d1n), static_cast<double>(d2n));
            status[dst] = st;
            fused_eval_point<EVAL, TO, TC>(ev, true, dst, a, b, c, d, xs, st, acc);
        }
    }
    if (r < 0) defer_point(df, dst);                // weight ratio left its range / exact-zero depth: the follow-up kernel redoes it
}

// Copy one finished 32-point slice of x / status from local HBM to every mirror (one warp, coalesced rows).
template <typename TO>
__device__ __forceinline__ void mirror_slice(const TO* __restrict__ x, const int32_t* __restrict__ status, const Mirrors& mir,
                                             int64_t base, int64_t n) {
    if (base >= n) return;
    const int lane = threadIdx.x & 31;
    const int cnt = (n - base) >= 32 ? 32 : static_cast<int>(n - base);
    TO v[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) v[k] = (k * 32 + lane < cnt * 3) ? __ldcg(x + base * 3 + k * 32 + lane) : TO(0);
    const int32_t sv = lane < cnt ? __ldcg(status + base + lane) : 0;
    int r = mirror_rotation(mir.count);
    for (int m = 0; m < mir.count; ++m) {
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (k * 32 + lane < cnt * 3) mirror_put<TO>(mir, r, base * 3 + k * 32 + lane, v[k]);
        if (lane < cnt) static_cast<int32_t*>(mir.status[r])[base + lane] = sv;
        if (++r == mir.count) r = 0;
    }
}

template <typename TI, typename TC, typename TO, class PRE = PreNone, bool EVAL = false>
__global__ void __launch_bounds__(kThreads, TRGL_ITER_MINB)
k_iterative_ls(const TI* __restrict__ u1, const TI* __restrict__ u2, const __grid_constant__ Cams<TC> cams,
               const __grid_constant__ RayGeom<TC> geom, TO* __restrict__ x, int32_t* __restrict__ status, const int64_t n,
               const TC tolerance, const int py_semantics, const __grid_constant__ PRE pre_stage,
               const __grid_constant__ Mirrors mir, const __grid_constant__ EvalArg<EVAL> ev,
               const __grid_constant__ Deferred df) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    IterSmem<TI, TC, TO>& sm = *reinterpret_cast<IterSmem<TI, TC, TO>*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto& pre = sm.pre; auto& q_state = sm.q_state[warp]; auto& q_idx = sm.q_idx[warp]; auto& stage = sm.stage[warp];
    const Mirrors local_only = {0, 0, {nullptr}, {nullptr}};       // this kernel mirrors whole slices, see above
    const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;
    const int64_t first = static_cast<int64_t>(blockIdx.x) * kThreads;
    // loop-carried state is four 32-bit registers: pass number, mirrored passes, queue tail and length (warp-uniform)
    int mirrored = 0;                                    // passes of this CTA whose slice this warp has copied to the mirrors
    unsigned int tail = 0u;                              // queue position of the next push (oldest entry = tail - count)
    int count = 0;                                       // queue length
    double acc[4] = {0, 0, 0, 0};                        // fused evaluation sums of this thread
    pre.issue(u1, u2, first + threadIdx.x, n);
    for (int pass = 0;; ++pass) {
        const int64_t tile = first + pass * stride;
        const bool have_tile = tile < n;                 // block-uniform; the pass after the last tile only flushes the queue
        if (have_tile) {
            const int64_t i = tile + threadIdx.x;
            TC a, b, c, d;
            pre.take(i, n, a, b, c, d);
            pre.issue(u1, u2, i + stride, n);            // next tile of this CTA lands while this one is solved
            if constexpr (PRE::kActive) pre_stage.template apply<TI, TC>(a, b, c, d);
            TwoRay<TC> R;
            TC xs[3] = {0, 0, 0};
            TC kap = 0, d1 = 1, d2 = 1, d1n = 1, d2n = 1, inv = 1, kap_used = 1;
            int it = -1, r = -1;
            if (tworay_setup<TC>(cams, geom, a, b, c, d, R)) {
                kap = R.kap;
#pragma unroll                                           // straight-line: no loop-carried register moves (72 instead of 78
                                                         // registers, 0.263 instead of 0.269 ms per 10 M points)
                for (int k = 0; k < kPhase1; ++k) {
                    r = tworay_step<TC>(R.d11, R.d12, R.d21, R.d22, kap, d1, d2, d1n, d2n, inv, kap_used, tolerance, py_semantics);
                    if (r) { if (r > 0) it = k; break; }
                }
            }
            if (r > 0) {
#pragma unroll
                for (int k = 0; k < 3; ++k) xs[k] = tfma(kap_used, R.X2[k], R.X1[k]) * inv;
            }
            // r == 0: still iterating -> queued; r < 0: not certified -> deferred to the follow-up kernel
            const bool pending = (r == 0) && (i < n);
            if (r < 0 && i < n) defer_point(df, i);
            const unsigned ball = __ballot_sync(0xffffffffu, pending);
            if (pending) {
                const int slot = static_cast<int>((tail + __popc(ball & ((1u << lane) - 1u))) & (kWarpQueueCap - 1));
                q_state[0][slot] = kap;
                q_state[1][slot] = d1; q_state[2][slot] = d2;
                q_state[3][slot] = R.d11; q_state[4][slot] = R.d12; q_state[5][slot] = R.d21; q_state[6][slot] = R.d22;
#pragma unroll
                for (int k = 0; k < 3; ++k) { q_state[7 + k][slot] = R.X1[k]; q_state[10 + k][slot] = R.X2[k]; }
                q_idx[slot] = i;
            }
            tail += __popc(ball); count += __popc(ball);
            // finished points leave through the coalesced path (pending lanes write a placeholder that phase 2 overwrites);
            // the __syncwarp inside also publishes the queue entries to the warp
            store_x_warp(x, tile + warp * 32, n, static_cast<TO>(xs[0]), static_cast<TO>(xs[1]),
                             static_cast<TO>(xs[2]), stage, local_only);
            if (i < n && r > 0) {
                const int st = iter_ls_status(it, static_cast<double>(d1n), static_cast<double>(d2n));
                status[i] = st;
                fused_eval_point<EVAL, TO, TC>(ev, true, i, a, b, c, d, xs, st, acc);
            }
        }
        if (count >= 32 || (!have_tile && count > 0)) {   // count < kWarpQueueCap: <= 31 left over + 32 pushed
            const int take = count < 32 ? count : 32;
            __syncwarp();
            if (lane < take)
                iter_ls_phase2<TI, TC, TO, PRE, EVAL>(u1, u2, pre_stage, cams, q_state, q_idx,
                                                      static_cast<int>((tail - count + lane) & (kWarpQueueCap - 1)), x, status,
                                                      tolerance, py_semantics, ev, df, acc);
            count -= take;
            __syncwarp();                                // the drained slots may be overwritten, the results are visible
        }
        if (mir.count) {
            // slices of the passes before the one that holds the oldest queued point are final (pushes are in pass order)
            int safe = pass + 1;
            if (count > 0) safe = static_cast<int>((q_idx[(tail - count) & (kWarpQueueCap - 1)] - first) / stride);
            __syncwarp();
            for (; mirrored < safe; ++mirrored) mirror_slice<TO>(x, status, mir, first + mirrored * stride + warp * 32, n);
        }
        if (!have_tile) break;
    }
    fused_eval_finish<EVAL>(ev, acc, df.ctl, df.cap);
}

// Tail of the follow-up kernels: add this thread's evaluation sums to the sums the hot kernel has already written
// (warp shuffle + one atomic per warp and sum), then the last block to finish re-arms the list for the next call.
template <bool EVAL>
__device__ __forceinline__ void followup_finish(const EvalArg<EVAL>& ev, const Deferred& df, double (&acc)[4], bool any,
                                                bool rearm) {
    if constexpr (EVAL) {
        if (any) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                double v = acc[q];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(&ev.e.sums_out[q], v);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && rearm) {
        __threadfence();
        if (atomicAdd(&df.ctl[1], 1u) == gridDim.x - 1) {
            // ctl[2..3]: running total of deferred points of this (device, stream), for diagnostics (trgl_deferred_total)
            *reinterpret_cast<unsigned long long*>(df.ctl + 2) += df.ctl[0];
            if (df.hint) *reinterpret_cast<volatile unsigned int*>(df.hint) = df.ctl[0];
            df.ctl[0] = 0u; df.ctl[1] = 0u; __threadfence();
        }
    }
}

// Follow-up kernel of k_iterative_ls: the reference's loop as written for the deferred points (all == 0: the first
// ctl[0] entries of the list, or every point if the list overflowed) or for every point (all != 0: cameras without a
// finite centre / two-ray forms switched off -- then k_iterative_ls is not launched at all).  Results go to x / status
// and every mirror point by point; evaluation sums are added to the sums the hot kernel has already written.
template <typename TI, typename TC, typename TO, class PRE = PreNone, bool EVAL = false>
__global__ void __launch_bounds__(kThreads)
k_iterative_general(const TI* __restrict__ u1, const TI* __restrict__ u2, const __grid_constant__ Cams<TC> cams,
                    TO* __restrict__ x, int32_t* __restrict__ status, const int64_t n, const TC tolerance,
                    const int py_semantics, const __grid_constant__ PRE pre_stage, const __grid_constant__ Mirrors mir,
                    const __grid_constant__ EvalArg<EVAL> ev, const __grid_constant__ Deferred df, const int all) {
    wait_for_hot_kernel();
    const unsigned int listed = all ? 0u : df.ctl[0];
    const bool everything = all || listed > df.cap;
    const int64_t total = everything ? n : static_cast<int64_t>(listed);
    double acc[4] = {0, 0, 0, 0};
    for (int64_t k = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; k < total;
         k += static_cast<int64_t>(gridDim.x) * kThreads) {
        const int64_t i = everything ? k : df.idx[k];
        TC a, b, c, d, xs[3], d1n, d2n;
        reload_inputs<TI, TC, PRE>(u1, u2, pre_stage, i, a, b, c, d);
        const int it = iter_point_general<TC>(cams, a, b, c, d, tolerance, py_semantics, xs, d1n, d2n);
        const int st = iter_ls_status(it, static_cast<double>(d1n), static_cast<double>(d2n));
#pragma unroll
        for (int q = 0; q < 3; ++q) x[3 * i + q] = static_cast<TO>(xs[q]);
        status[i] = st;
        for (int r = 0; r < mir.count; ++r) {
#pragma unroll
            for (int q = 0; q < 3; ++q) mirror_put<TO>(mir, r, 3 * i + q, static_cast<TO>(xs[q]));
            static_cast<int32_t*>(mir.status[r])[i] = st;
        }
        fused_eval_point<EVAL, TO, TC>(ev, true, i, a, b, c, d, xs, st, acc);
    }
    followup_finish<EVAL>(ev, df, acc, total > 0, !all);
}

// Follow-up kernel of k_linear_ls<..., DEFER>: the deferred points (or every point if the list overflowed) through the
// conditioning tiers of solve_point_careful -- refinement step, Jacobi SVD with OpenCV's rank rule (triangulation.c:81).
template <typename TI, typename TC, typename TO, class PRE = PreNone, bool EVAL = false>
__global__ void __launch_bounds__(kThreads)
k_linear_ls_general(const TI* __restrict__ u1, const TI* __restrict__ u2, const __grid_constant__ Cams<TC> cams,
                    TO* __restrict__ x, const int64_t n, const __grid_constant__ PRE pre_stage,
                    const __grid_constant__ Mirrors mir, const __grid_constant__ EvalArg<EVAL> ev,
                    const __grid_constant__ Deferred df) {
    wait_for_hot_kernel();
    const unsigned int listed = df.ctl[0];
    const bool everything = listed > df.cap;
    const int64_t total = everything ? n : static_cast<int64_t>(listed);
    double acc[4] = {0, 0, 0, 0};
    // 4 points per thread and pass: the list entries, then the 8 input loads, are issued together -- a rig can defer
    // millions of points (forward motion: 40 % are beyond tier 1), and the gathers are what this kernel waits for
    constexpr int kBatch = 4;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;
    for (int64_t k0 = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; k0 < total; k0 += kBatch * stride) {
        int64_t idx[kBatch];
        TC in[kBatch][4];
#pragma unroll
        for (int p = 0; p < kBatch; ++p) {
            const int64_t k = k0 + p * stride;
            idx[p] = k < total ? (everything ? k : df.idx[k]) : -1;
        }
#pragma unroll
        for (int p = 0; p < kBatch; ++p) {
            in[p][0] = in[p][1] = in[p][2] = in[p][3] = TC(0);
            if (idx[p] >= 0) reload_inputs<TI, TC, PRE>(u1, u2, pre_stage, idx[p], in[p][0], in[p][1], in[p][2], in[p][3]);
        }
#pragma unroll
        for (int p = 0; p < kBatch; ++p) {       // unrolled: a run-time p would put idx[] / in[][] into local memory
            const int64_t i = idx[p];
            if (i < 0) continue;
            TC xs[3];
            // same two-step as the in-line path: points within tier 1 keep the adjugate solve (only reached on overflow)
            if (!ls_point_fast<TC>(cams, in[p][0], in[p][1], in[p][2], in[p][3], xs)) {
                // tier 2 in line (the bulk of what a low-parallax rig defers), the SVD tier out of line -- the same
                // arithmetic as solve_point_careful, which the non-deferring variants call
                bool done = false;
                if constexpr (sizeof(TC) == 8) {
                    double rows[4][4];
                    weighted_rows<double>(cams, in[p][0], in[p][1], in[p][2], in[p][3], 1.0, 1.0, rows);
                    done = solve4x3_tier2_f64(rows, xs);
                }
                if (!done) solve_point_careful<TC>(cams, in[p][0], in[p][1], in[p][2], in[p][3], TC(1), TC(1), xs);
            }
#pragma unroll
            for (int q = 0; q < 3; ++q) x[3 * i + q] = static_cast<TO>(xs[q]);
            for (int r = 0; r < mir.count; ++r) {
#pragma unroll
                for (int q = 0; q < 3; ++q) mirror_put<TO>(mir, r, 3 * i + q, static_cast<TO>(xs[q]));
            }
            fused_eval_point<EVAL, TO, TC>(ev, true, i, in[p][0], in[p][1], in[p][2], in[p][3], xs, 1, acc);
        }
    }
    followup_finish<EVAL>(ev, df, acc, total > 0, true);
}

// ---- linear_eigen_triangulation (triangulation.py:6-25) --------------------------------------------------------
// Smallest right singular vector of the ROWS x 4 DLT matrix, dehomogenised, with the finite-coordinates mask.
template <typename TC, int ROWS>
__device__ __forceinline__ void dlt_matrix(const Cams<TC>& cams, TC u1x, TC u1y, TC u2x, TC u2y, TC B[ROWS][4]) {
    constexpr int per = ROWS / 2;
    dlt_rows<TC>(cams.P1, u1x, u1y, B[0], B[1]);
    dlt_rows<TC>(cams.P2, u2x, u2y, B[per], B[per + 1]);
    if constexpr (ROWS == 6) {       // OpenCV 2.4 adds x*P[1,:] - y*P[0,:] per view
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            B[2][k] = tfma(u1x, cams.P1[4 + k], -u1y * cams.P1[k]);
            B[5][k] = tfma(u2x, cams.P2[4 + k], -u2y * cams.P2[k]);
        }
    }
}

// Reference path: one-sided Jacobi SVD (the method cv::SVD runs on a small matrix), preconditioned the Drmac-Veselic way.
// Out of line: used only when the fast path cannot certify its answer (no gap at the level of G, i.e. low parallax -- 10 %
// of the forward-motion rig --, points at infinity / on the baseline, NaN systems).
//   B = Q R by Householder reflections (R keeps B's singular values and right singular vectors; backward stable), then
//   the rotations orthogonalise the ROWS of R, i.e. the columns of R^T:  R^T V = U' Sigma  =>  R = V Sigma U'^T, so the
//   right singular vectors of B are the normalised rows themselves and no V has to be accumulated; the row of smallest
//   norm is the answer.  On the triangular factor the sweeps converge in 3 instead of 5 (12 instead of 21 rotations per
//   point on the forward rig), and accuracy is governed by the relative gap of the singular values of B, not of G = B^T B.
// Rotation parameters from two rsqrt (approximation + two Newton steps, full_rsqrt) instead of three IEEE sqrt and two divisions.
template <typename TC, int ROWS>
__device__ __noinline__ void eigen_point_jacobi(const Cams<TC>& cams, TC u1x, TC u1y, TC u2x, TC u2y, TC X[4]) {
    TC B[ROWS][4], R[4][4];
    dlt_matrix<TC, ROWS>(cams, u1x, u1y, u2x, u2y, B);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        TC nn = 0;
#pragma unroll
        for (int r = k; r < ROWS; ++r) nn = tfma(B[r][k], B[r][k], nn);
        const TC nrm = tsqrt(nn);
        const TC alpha = B[k][k] > TC(0) ? -nrm : nrm;
        const TC v0 = B[k][k] - alpha;                       // same sign: no cancellation
        const TC half_vtv = tfma(tabs(B[k][k]), nrm, nn);    // v.v / 2
        const TC ib = half_vtv > TC(0) ? TC(1) / half_vtv : TC(0);
#pragma unroll
        for (int j = 0; j < k; ++j) R[k][j] = TC(0);
        R[k][k] = alpha;
#pragma unroll
        for (int j = k + 1; j < 4; ++j) {
            TC dot = v0 * B[k][j];
#pragma unroll
            for (int r = k + 1; r < ROWS; ++r) dot = tfma(B[r][k], B[r][j], dot);
            const TC sc = dot * ib;
            R[k][j] = tfma(-sc, v0, B[k][j]);
#pragma unroll
            for (int r = k + 1; r < ROWS; ++r) B[r][j] = tfma(-sc, B[r][k], B[r][j]);
        }
    }
    const TC eps2 = Num<TC>::eps() * Num<TC>::eps();
#pragma unroll 1
    for (int sweep = 0; sweep < 30; ++sweep) {
        bool changed = false;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
#pragma unroll
            for (int j = i + 1; j < 4; ++j) {
                TC a = 0, b = 0, p = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    a = tfma(R[i][k], R[i][k], a);
                    b = tfma(R[j][k], R[j][k], b);
                    p = tfma(R[i][k], R[j][k], p);
                }
                if (!(p * p > eps2 * a * b)) continue;                        // also skips NaN
                const TC p2 = p + p, beta = a - b;
                const TC q = tfma(p2, p2, beta * beta);
                if (!(q > Num<TC>::tiny()) || !(q < Num<TC>::big())) continue;         // rotation not representable: leave the pair
                changed = true;
                const TC ig = full_rsqrt(q);                                  // 1 / gamma
                const TC h = tfma(TC(0.5) * tabs(beta), ig, TC(0.5));         // (gamma + |beta|) / (2 gamma)  in [1/2, 1]
                const TC ih = full_rsqrt(h);
                const TC big = h * ih, small = TC(0.5) * p2 * ig * ih;        // sqrt(h),  p / (gamma sqrt(h))
                const TC c = beta < TC(0) ? small : big, sn = beta < TC(0) ? big : small;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const TC t0 = tfma(c, R[i][k], sn * R[j][k]);
                    const TC t1 = tfma(c, R[j][k], -sn * R[i][k]);
                    R[i][k] = t0; R[j][k] = t1;
                }
            }
        }
        if (!changed) break;
    }
    // The row of smallest norm belongs to sigma4 -- but when sigma4 is at rounding level (exact correspondences) its
    // DIRECTION is rounding noise too, so the answer is taken as the vector orthogonal to the other three rows (4-D cross
    // product; their mutual orthogonality makes the minors well conditioned): error ~ eps sigma1 / sigma3 at worst.
    TC nsq[4];
    bool nan = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        nsq[j] = tfma(R[j][0], R[j][0], tfma(R[j][1], R[j][1], tfma(R[j][2], R[j][2], R[j][3] * R[j][3])));
        nan |= !(nsq[j] == nsq[j]);
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {                             // move the smallest row to position 3
        if (nsq[j] < nsq[3]) {
            const TC t = nsq[j]; nsq[j] = nsq[3]; nsq[3] = t;
#pragma unroll
            for (int k = 0; k < 4; ++k) { const TC r = R[j][k]; R[j][k] = R[3][k]; R[3][k] = r; }
        }
    }
    const TC m01 = tfma(R[1][0], R[2][1], -R[1][1] * R[2][0]), m02 = tfma(R[1][0], R[2][2], -R[1][2] * R[2][0]);
    const TC m03 = tfma(R[1][0], R[2][3], -R[1][3] * R[2][0]), m12 = tfma(R[1][1], R[2][2], -R[1][2] * R[2][1]);
    const TC m13 = tfma(R[1][1], R[2][3], -R[1][3] * R[2][1]), m23 = tfma(R[1][2], R[2][3], -R[1][3] * R[2][2]);
    X[0] = -tfma(R[0][1], m23, tfma(-R[0][2], m13, R[0][3] * m12));
    X[1] = tfma(R[0][0], m23, tfma(-R[0][2], m03, R[0][3] * m02));
    X[2] = -tfma(R[0][0], m13, tfma(-R[0][1], m03, R[0][3] * m01));
    X[3] = tfma(R[0][0], m12, tfma(-R[0][1], m02, R[0][2] * m01));
    const TC xx = tfma(X[0], X[0], tfma(X[1], X[1], tfma(X[2], X[2], X[3] * X[3])));
    // a unit vector, as the callers' tests assume; rank <= 2 (no cross product) falls back to the smallest row, and a
    // NaN system yields a NaN point
    const bool usable = xx > Num<TC>::tiny() && xx < Num<TC>::big();
    const TC nrm = usable ? TC(1) / tsqrt(xx) : (nsq[3] > TC(0) ? TC(1) / tsqrt(nsq[3]) : TC(1));
#pragma unroll
    for (int k = 0; k < 4; ++k) X[k] = nan ? nsq[0] + nsq[1] + nsq[2] + nsq[3] : (usable ? X[k] : R[3][k]) * nrm;
}

// LDL^T (no pivoting) of the symmetric 4x4 S (order 00 01 02 03 11 12 13 22 23 33); optionally solves S y = b.
// Returns the number of negative pivots (the inertia, by Sylvester's law) or -1 if a pivot is not usable.
template <typename T, bool SOLVE>
__device__ __forceinline__ int ldl4(const T S[10], const T b[4], T y[4]) {
    const T d0 = S[0];
    const T i0 = fast_rcp(d0);
    const T l10 = S[1] * i0, l20 = S[2] * i0, l30 = S[3] * i0;
    const T d1 = tfma(-l10, S[1], S[4]);
    const T i1 = fast_rcp(d1);
    const T t21 = tfma(-l20, S[1], S[5]), t31 = tfma(-l30, S[1], S[6]);
    const T l21 = t21 * i1, l31 = t31 * i1;
    const T d2 = tfma(-l21, t21, tfma(-l20, S[2], S[7]));
    const T i2 = fast_rcp(d2);
    const T t32 = tfma(-l31, t21, tfma(-l30, S[2], S[8]));
    const T l32 = t32 * i2;
    T d3 = tfma(-l32, t32, tfma(-l31, t31, tfma(-l30, S[3], S[9])));
    if (!(d0 != T(0) && d1 != T(0) && d2 != T(0)) || !(d3 == d3)) return -1;
    const int neg = (d0 < T(0)) + (d1 < T(0)) + (d2 < T(0)) + (d3 < T(0));
    if constexpr (SOLVE) {
        if (d3 == T(0)) d3 = Num<T>::tiny();          // exactly singular shift: any huge multiple of the null vector
        T z0 = b[0];
        T z1 = tfma(-l10, z0, b[1]);
        T z2 = tfma(-l21, z1, tfma(-l20, z0, b[2]));
        T z3 = tfma(-l32, z2, tfma(-l31, z1, tfma(-l30, z0, b[3])));
        z0 *= i0; z1 *= i1; z2 *= i2; z3 = z3 / d3;
        y[3] = z3;
        y[2] = tfma(-l32, y[3], z2);
        y[1] = tfma(-l31, y[3], tfma(-l21, y[2], z1));
        y[0] = tfma(-l30, y[3], tfma(-l20, y[2], tfma(-l10, y[1], z0)));
    }
    return neg;
}

// Fast path: Rayleigh-quotient iteration on G = B^T B started from the least-squares point (which is within noise
// of the answer), with a certificate: (i) ||G X - lam X|| <= 4 eps tr(G) and (ii) exactly one eigenvalue of G lies below
// lam + gap*tr(G) (inertia of the shifted matrix), gap = 5e-6 / |w|.  (i)+(ii) prove X is the eigenvector of the smallest eigenvalue and that
// it is separated well enough for the G-based computation to be accurate to ~1e-12; otherwise the caller falls back to
// the Jacobi SVD.  ~450 FP64 instructions instead of ~4800.
// SYNC: the caller guarantees that all 32 lanes of the warp are here (the hot kernel); the warp is then re-converged
// right behind the iteration loop, so that the lanes leaving it at different rounds run the certificate and everything
// the caller does next ONCE -- without it ptxas placed the reconvergence point behind the evaluation epilogue, which
// then ran 2.3 times per warp (profiles/README.md, r01f).
template <typename TC, int ROWS, bool SYNC = false>
__device__ __forceinline__ bool eigen_point_fast(const Cams<TC>& cams, TC u1x, TC u1y, TC u2x, TC u2y, TC X[4]) {
    TC G[10];
    {
        TC B[ROWS][4];
        dlt_matrix<TC, ROWS>(cams, u1x, u1y, u2x, u2y, B);
        int k = 0;
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int q = p; q < 4; ++q) {
                TC s = B[0][p] * B[0][q];
#pragma unroll
                for (int r = 1; r < ROWS; ++r) s = tfma(B[r][p], B[r][q], s);
                G[k++] = s;
            }
    }
    const TC tr = G[0] + G[4] + G[7] + G[9];
    {   // least-squares start: G[0:3,0:3] x = -G[0:3,3]
        const TC M[6] = {G[0], G[1], G[2], G[4], G[5], G[7]};
        const TC v[3] = {-G[3], -G[6], -G[8]};
        TC C[6], x0[3];
        const TC det = sym3_cofactors(M, C);
        sym3_apply(C, v, fast_rcp(det), x0);
        const TC nrm = trsqrt(tfma(x0[0], x0[0], tfma(x0[1], x0[1], tfma(x0[2], x0[2], TC(1)))));
        X[0] = x0[0] * nrm; X[1] = x0[1] * nrm; X[2] = x0[2] * nrm; X[3] = nrm;
    }
    const TC tol = TC(4) * Num<TC>::eps() * tr;
    TC lam = 0;
    bool conv = false;
#pragma unroll 1
    for (int it = 0; it < 5; ++it) {
        TC y[4];
        y[0] = tfma(G[0], X[0], tfma(G[1], X[1], tfma(G[2], X[2], G[3] * X[3])));
        y[1] = tfma(G[1], X[0], tfma(G[4], X[1], tfma(G[5], X[2], G[6] * X[3])));
        y[2] = tfma(G[2], X[0], tfma(G[5], X[1], tfma(G[7], X[2], G[8] * X[3])));
        y[3] = tfma(G[3], X[0], tfma(G[6], X[1], tfma(G[8], X[2], G[9] * X[3])));
        lam = tfma(X[0], y[0], tfma(X[1], y[1], tfma(X[2], y[2], X[3] * y[3])));
        TC rn = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) { const TC r = tfma(-lam, X[k], y[k]); rn = tfma(r, r, rn); }
        if (rn <= tol * tol) { conv = true; break; }
        if (it == 4) break;
        TC S[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) S[k] = G[k];
        S[0] -= lam; S[4] -= lam; S[7] -= lam; S[9] -= lam;
        if (ldl4<TC, true>(S, X, y) < 0) break;            // unusable pivot: not converged, single loop exit
        const TC nrm = trsqrt(tfma(y[0], y[0], tfma(y[1], y[1], tfma(y[2], y[2], y[3] * y[3]))));
#pragma unroll
        for (int k = 0; k < 4; ++k) X[k] = y[k] * nrm;
    }
    if constexpr (SYNC) __syncwarp();
    TC S[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) S[k] = G[k];
    // Required separation of the two smallest eigenvalues, relative to tr(G).  The entries of G carry ~4 eps tr(G) of
    // rounding, which moves the eigenvector by 4 eps tr / (lam3 - lam4) and the DEHOMOGENISED point by 1/|w| times that
    // (w = X[3] of the unit vector): measured on the forward-motion rig, 4 eps / (gap |w|) to within 10 % on the 51 of
    // 10 M points that exceeded 1e-9 with a fixed gap of 2e-5 (|w| 1e-4 .. 0.04: points far behind the scene).  So the
    // required gap is 5e-6 / |w| (error bound 4 eps / 5e-6 = 1.8e-10); below it the point takes the Jacobi SVD.
    const TC gap_rel = sizeof(TC) == 8 ? TC(5e-6) * fast_rcp(tabs(X[3])) : TC(2e-3);
    const TC shift = tfma(gap_rel, tr, lam);
    S[0] -= shift; S[4] -= shift; S[7] -= shift; S[9] -= shift;
    return (ldl4<TC, false>(S, X, X) == 1) && conv && (gap_rel == gap_rel);
}

// ---- warp-uniform variant of the fast path for the hot kernel -----------------------------------------------------------
// The loop above leaves at a per-lane round and the warp pays for its slowest lane; how many rounds a point needs depends on
// the rig and the noise (measured, residual <= 4 eps tr(G) after 2 / 3 rounds: rotating 0.8 px 98.7 % / 99.998 %,
// translating 0.8 px 85 % / 99.3 %, forward motion 48 % / 84 %, any rig at 8 px 30-45 % / 75-90 %).  Here the whole warp
// runs the same number of rounds and decides together after each one (from the second on): all lanes converged -> done;
// at most kEigenStragglers lanes left -> those are handed to the follow-up kernel (which runs the loop above, then the
// Jacobi SVD) and the warp is done; otherwise everybody runs another round (a converged lane just stays converged), up to
// kEigenMaxRounds.  One more round costs the warp 32 x ~95 instructions, a deferred point ~700 plus its share of a
// follow-up kernel that runs at low occupancy: measured per 10 M points with 3 / 2 / 1 / 0 stragglers allowed -- rotating rig
// 0.385 / 0.384 / 0.384 / 0.390 ms, translating 0.478 / 0.464 / 0.464 / 0.468, forward motion 0.841 / 0.850 / 0.855 / 0.857,
// general 0.359 / 0.359 / 0.358 / 0.355: flat, 2 is the best compromise.
//   * the solve returns the DIRECTION d3 * (G - rho I)^-1 x: the last pivot d3 -> 0 as rho -> lam4 (that is the point of the
//     iteration), so nothing is divided by it -- no IEEE division, no special-case path, no overflow;
//   * normalisation uses MUFU.RSQ64H + one Newton step (1e-12): the iteration is self-correcting, and the Rayleigh
//     quotient / residual test divide by the exact X.X.
constexpr int kEigenMaxRounds = 4;
#ifndef TRGL_EIGEN_STRAGGLERS
#define TRGL_EIGEN_STRAGGLERS 2
#endif
constexpr int kEigenStragglers = TRGL_EIGEN_STRAGGLERS;

__device__ __forceinline__ double fast_rsqrt(double x) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double h = (x * r) * -0.5;
    return fma(r, fma(h, r, 0.5), r);
}
__device__ __forceinline__ float fast_rsqrt(float x) { return rsqrtf(x); }

// y = d3 * S^-1 b by LDL^T without pivoting (S order 00 01 02 03 11 12 13 22 23 33); false if a leading pivot is unusable.
template <typename T>
__device__ __forceinline__ bool ldl4_direction(const T S[10], const T b[4], T y[4]) {
    const T d0 = S[0];
    const T i0 = fast_rcp(d0);
    const T l10 = S[1] * i0, l20 = S[2] * i0, l30 = S[3] * i0;
    const T d1 = tfma(-l10, S[1], S[4]);
    const T i1 = fast_rcp(d1);
    const T t21 = tfma(-l20, S[1], S[5]), t31 = tfma(-l30, S[1], S[6]);
    const T l21 = t21 * i1, l31 = t31 * i1;
    const T d2 = tfma(-l21, t21, tfma(-l20, S[2], S[7]));
    const T i2 = fast_rcp(d2);
    const T t32 = tfma(-l31, t21, tfma(-l30, S[2], S[8]));
    const T l32 = t32 * i2;
    const T d3 = tfma(-l32, t32, tfma(-l31, t31, tfma(-l30, S[3], S[9])));
    const T z0 = b[0];
    const T z1 = tfma(-l10, z0, b[1]);
    const T z2 = tfma(-l21, z1, tfma(-l20, z0, b[2]));
    const T z3 = tfma(-l32, z2, tfma(-l31, z1, tfma(-l30, z0, b[3])));
    y[3] = z3;
    y[2] = tfma(-l32, z3, (z2 * i2) * d3);
    y[1] = tfma(-l31, z3, tfma(-l21, y[2], (z1 * i1) * d3));
    y[0] = tfma(-l30, z3, tfma(-l20, y[2], tfma(-l10, y[1], (z0 * i0) * d3)));
    return (d0 != T(0)) && (d1 != T(0)) && (d2 != T(0)) && (d3 == d3);
}

// WARP = true (hot kernel): all 32 lanes of the warp call this together and decide together when to stop.  WARP = false
// (follow-up kernel): the same arithmetic for one point on its own, up to kEigenLoneRounds rounds.  A lane FREEZES its
// iterate at the first round whose residual passes, so the result of a point is the same bits wherever it is computed --
// in a warp that stopped early, in one that went on for its neighbours, or in the follow-up kernel: results do not depend
// on how a batch is cut into shards or where a point sits in it.  Returns true when X is certified.
constexpr int kEigenLoneRounds = 8;
template <typename TC, int ROWS, bool WARP>
__device__ __forceinline__ bool eigen_point_iter(const Cams<TC>& cams, TC u1x, TC u1y, TC u2x, TC u2y, TC X[4]) {
    TC G[10];
    {
        TC B[ROWS][4];
        dlt_matrix<TC, ROWS>(cams, u1x, u1y, u2x, u2y, B);
        int k = 0;
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = a; b < 4; ++b) {
                TC s = B[0][a] * B[0][b];
#pragma unroll
                for (int r = 1; r < ROWS; ++r) s = tfma(B[r][a], B[r][b], s);
                G[k++] = s;
            }
    }
    const TC tr = G[0] + G[4] + G[7] + G[9];
    {   // least-squares start: G[0:3,0:3] x = -G[0:3,3]
        const TC M[6] = {G[0], G[1], G[2], G[4], G[5], G[7]};
        const TC v[3] = {-G[3], -G[6], -G[8]};
        TC C[6], x0[3];
        const TC det = sym3_cofactors(M, C);
        sym3_apply(C, v, fast_rcp(det), x0);
        const TC nrm = fast_rsqrt(tfma(x0[0], x0[0], tfma(x0[1], x0[1], tfma(x0[2], x0[2], TC(1)))));
        X[0] = x0[0] * nrm; X[1] = x0[1] * nrm; X[2] = x0[2] * nrm; X[3] = nrm;
    }
    const TC tol = TC(4) * Num<TC>::eps() * tr;
    TC lam = 0;
    bool alive = true, conv = false;
    // one round: Rayleigh quotient of the (unit, to 1e-12) iterate, shifted solve, normalisation; a converged lane keeps X
    auto rqi_round = [&](const TC (&y)[4]) {
        const TC rho = conv ? lam : tfma(X[0], y[0], tfma(X[1], y[1], tfma(X[2], y[2], X[3] * y[3])));
        TC S[10], z[4];
#pragma unroll
        for (int k = 0; k < 10; ++k) S[k] = G[k];
        S[0] -= rho; S[4] -= rho; S[7] -= rho; S[9] -= rho;
        const bool usable = ldl4_direction<TC>(S, X, z);
        const TC nrm = fast_rsqrt(tfma(z[0], z[0], tfma(z[1], z[1], tfma(z[2], z[2], z[3] * z[3]))));
        if (!conv) {
            alive = alive && usable;
#pragma unroll
            for (int k = 0; k < 4; ++k) X[k] = z[k] * nrm;
        }
    };
    auto matvec = [&](TC (&y)[4]) {
        y[0] = tfma(G[0], X[0], tfma(G[1], X[1], tfma(G[2], X[2], G[3] * X[3])));
        y[1] = tfma(G[1], X[0], tfma(G[4], X[1], tfma(G[5], X[2], G[6] * X[3])));
        y[2] = tfma(G[2], X[0], tfma(G[5], X[1], tfma(G[7], X[2], G[8] * X[3])));
        y[3] = tfma(G[3], X[0], tfma(G[6], X[1], tfma(G[8], X[2], G[9] * X[3])));
    };
    TC y[4];
    // the LS start and the first iterate are never converged: two rounds straight-line, then test / vote / continue
    matvec(y); rqi_round(y);
    matvec(y); rqi_round(y);
#pragma unroll 1
    for (int round = 2;; ++round) {
        matvec(y);
        if (!conv) {
            const TC xx = tfma(X[0], X[0], tfma(X[1], X[1], tfma(X[2], X[2], X[3] * X[3])));
            lam = tfma(X[0], y[0], tfma(X[1], y[1], tfma(X[2], y[2], X[3] * y[3]))) * fast_rcp(xx);     // exact Rayleigh quotient
            TC rn = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) { const TC r = tfma(-lam, X[k], y[k]); rn = tfma(r, r, rn); }
            conv = rn <= tol * tol * xx;
        }
        if constexpr (WARP) {
            const unsigned open = __ballot_sync(0xffffffffu, alive && !conv);
            if (__popc(open) <= kEigenStragglers || round == kEigenMaxRounds) break;           // warp-uniform
        } else {
            if (conv || !alive || round == kEigenLoneRounds) break;
        }
        rqi_round(y);
    }
    // certificate: exactly one eigenvalue of G below lam + gap tr(G)   (see eigen_point_fast)
    const TC gap_rel = sizeof(TC) == 8 ? TC(5e-6) * fast_rcp(tabs(X[3])) : TC(2e-3);
    const TC shift = tfma(gap_rel, tr, lam);
    TC S[10], dummy[4];
#pragma unroll
    for (int k = 0; k < 10; ++k) S[k] = G[k];
    S[0] -= shift; S[4] -= shift; S[7] -= shift; S[9] -= shift;
    return alive && conv && (ldl4<TC, false>(S, X, dummy) == 1) && (gap_rel == gap_rel);
}

// Dehomogenise a CERTIFIED vector (w is bounded away from 0 by the certificate): 2-ulp reciprocal instead of the IEEE
// division and its special-case path.
template <typename TC>
__device__ __forceinline__ void eigen_finish_fast(const TC X[4], TC max_coord, TC xs[3], bool& good) {
    const TC inv = fast_rcp(X[3]);
    xs[0] = X[0] * inv; xs[1] = X[1] * inv; xs[2] = X[2] * inv;
    const TC m = tmax(tmax(tabs(xs[0]), tabs(xs[1])), tabs(xs[2]));
    good = (xs[0] == xs[0]) && (xs[1] == xs[1]) && (xs[2] == xs[2]) && (m <= max_coord);
}

// Dehomogenise + finite-coordinates mask (triangulation.py:22-23).
template <typename TC>
__device__ __forceinline__ void eigen_finish(const TC X[4], TC max_coord, TC xs[3], bool& good) {
    const TC inv = TC(1) / X[3];
    xs[0] = X[0] * inv; xs[1] = X[1] * inv; xs[2] = X[2] * inv;        // Inf/NaN when w == 0
    const TC m = tmax(tmax(tabs(xs[0]), tabs(xs[1])), tabs(xs[2]));
    good = (xs[0] == xs[0]) && (xs[1] == xs[1]) && (xs[2] == xs[2]) && (m <= max_coord);   // NaN -> False
}

template <typename TC, int ROWS>
__device__ __forceinline__ void eigen_point(const Cams<TC>& cams, TC u1x, TC u1y, TC u2x, TC u2y,
                                            TC max_coord, TC xs[3], bool& good) {
    TC X[4];
    if (!eigen_point_fast<TC, ROWS>(cams, u1x, u1y, u2x, u2y, X))
        eigen_point_jacobi<TC, ROWS>(cams, u1x, u1y, u2x, u2y, X);
    eigen_finish<TC>(X, max_coord, xs, good);
}

// Persistent CTAs: grid-stride loop over 256-point tiles, the next tile's inputs prefetched with cp.async while the
// current tile is solved.  The hot kernel runs the certified Rayleigh-quotient iteration only and has no subroutine call;
// what it does not certify (breakdown, no convergence in 5 solves, eigenvalue gap below the certificate: points at
// infinity / on the baseline) is deferred to k_linear_eigen_general, the one-sided Jacobi SVD cv::SVD would run.  With
// the Jacobi path inside, its ~4800 instructions ran on one or two lanes while the rest of the warp waited: 0.69 ms per
// 10 M points on the translating rig (0.7 % of the points), 3.1 ms on the forward-motion rig.
// (A per-warp queue for the lanes that need a third solve, as in k_iterative_ls, was measured and dropped: the compiled
// resumable iteration cost more per solve than the divergence it removed, profiles/README.md.)
template <typename TI, typename TC, typename TO, int ROWS, class PRE = PreNone, bool EVAL = false>
__global__ void __launch_bounds__(kThreads, EVAL ? TRGL_EVAL_MINB : TRGL_EIGEN_MINB)
k_linear_eigen(const TI* __restrict__ u1, const TI* __restrict__ u2, const __grid_constant__ Cams<TC> cams,
               TO* __restrict__ x, uint8_t* __restrict__ status, const int64_t n, const TC max_coord,
               const __grid_constant__ PRE pre_stage, const __grid_constant__ Mirrors mir,
               const __grid_constant__ EvalArg<EVAL> ev, const __grid_constant__ Deferred df) {
    double acc[4] = {0, 0, 0, 0};
    __shared__ TO stage[kWarps][96];
    __shared__ __align__(16) PairPrefetch<TI, kThreads> pre;
    const int warp = threadIdx.x >> 5;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;
    int64_t tile = static_cast<int64_t>(blockIdx.x) * kThreads;
    pre.issue(u1, u2, tile + threadIdx.x, n);
    for (; tile < n; tile += stride) {
        const int64_t i = tile + threadIdx.x;
        TC a, b, c, d;
        pre.take(i, n, a, b, c, d);
        pre.issue(u1, u2, i + stride, n);
        if constexpr (PRE::kActive) pre_stage.template apply<TI, TC>(a, b, c, d);
        TC X[4], xs[3] = {0, 0, 0};
        bool good = false;
        const bool certified = eigen_point_iter<TC, ROWS, true>(cams, a, b, c, d, X);      // all 32 lanes present (uniform trip count)
        if (certified) eigen_finish_fast<TC>(X, max_coord, xs, good);
        else if (i < n) defer_point(df, i);
        store_x_warp(x, tile + warp * 32, n, static_cast<TO>(xs[0]), static_cast<TO>(xs[1]),
                         static_cast<TO>(xs[2]), stage[warp], mir);
        if (i < n && certified) {
            store_status(status, mir, i, static_cast<uint8_t>(good ? 1 : 0));
            fused_eval_point<EVAL, TO, TC>(ev, true, i, a, b, c, d, xs, good ? 1 : 0, acc);
        }
    }
    fused_eval_finish<EVAL>(ev, acc, df.ctl, df.cap);
}

// Follow-up kernel of k_linear_eigen: the deferred points (or every point if the list overflowed) through the one-sided
// Jacobi SVD, the path cv2.triangulatePoints itself takes.
template <typename TI, typename TC, typename TO, int ROWS, class PRE = PreNone, bool EVAL = false>
__global__ void __launch_bounds__(kThreads)
k_linear_eigen_general(const TI* __restrict__ u1, const TI* __restrict__ u2, const __grid_constant__ Cams<TC> cams,
                       TO* __restrict__ x, uint8_t* __restrict__ status, const int64_t n, const TC max_coord,
                       const __grid_constant__ PRE pre_stage, const __grid_constant__ Mirrors mir,
                       const __grid_constant__ EvalArg<EVAL> ev, const __grid_constant__ Deferred df) {
    wait_for_hot_kernel();
    const unsigned int listed = df.ctl[0];
    const bool everything = listed > df.cap;
    const int64_t total = everything ? n : static_cast<int64_t>(listed);
    double acc[4] = {0, 0, 0, 0};
    for (int64_t k = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; k < total;
         k += static_cast<int64_t>(gridDim.x) * kThreads) {
        const int64_t i = everything ? k : df.idx[k];
        TC a, b, c, d, X[4], xs[3];
        bool good;
        reload_inputs<TI, TC, PRE>(u1, u2, pre_stage, i, a, b, c, d);
        // the stragglers of the hot kernel's warps get the same iteration on their own (up to kEigenLoneRounds rounds); what
        // that does not certify either (no gap at the level of G -- low parallax --, breakdown, NaN) takes the one-sided
        // Jacobi SVD, the path cv::SVD itself runs
        if (eigen_point_iter<TC, ROWS, false>(cams, a, b, c, d, X)) {
            eigen_finish_fast<TC>(X, max_coord, xs, good);          // same arithmetic, hence same bits, as in the hot kernel
        } else {
            eigen_point_jacobi<TC, ROWS>(cams, a, b, c, d, X);
            eigen_finish<TC>(X, max_coord, xs, good);
        }
#pragma unroll
        for (int q = 0; q < 3; ++q) x[3 * i + q] = static_cast<TO>(xs[q]);
        for (int m = 0; m < mir.count; ++m) {
#pragma unroll
            for (int q = 0; q < 3; ++q) mirror_put<TO>(mir, m, 3 * i + q, static_cast<TO>(xs[q]));
        }
        store_status(status, mir, i, static_cast<uint8_t>(good ? 1 : 0));
        fused_eval_point<EVAL, TO, TC>(ev, true, i, a, b, c, d, xs, good ? 1 : 0, acc);
    }
    followup_finish<EVAL>(ev, df, acc, total > 0, true);
}

// ---- polynomial_triangulation (triangulation.py:198-232) -------------------------------------------------------
// Hartley-Sturm correction of the match (cv2.correctMatches) followed by the triangulation of the corrected match, fused.
// Hot kernel: the certified fast path of the correction, then the certified intersection of the two viewing rays (the
// corrected match satisfies the epipolar constraint, so the rays meet and the smallest singular vector of the DLT system
// is their intersection).  Whatever either certificate does not cover -- a non-monotone g (root isolation), a correction that is
// not exact after rounding to float32 storage, ill-conditioned or centre-less cameras -- is deferred to
// k_polynomial_general, which runs the complete correction (CTA-cooperative) and the ray intersection / eigen solver per
// point; the hot kernel has no subroutine call.  (On the forward-motion rig, where 1 % - 67 % of the points lack the certificate, the slow lanes used to
// hold their warps for 100 Durand-Kerner sweeps: 21 ms per 10 M points before the split, 0.54 ms now.)
template <typename TI, typename TC, typename TO, int ROWS, class PRE = PreNone, bool EVAL = false>
__global__ void __launch_bounds__(kThreads, EVAL ? TRGL_EVAL_MINB : TRGL_POLY_MINB)
k_polynomial(const TI* __restrict__ u1, const TI* __restrict__ u2, const __grid_constant__ Cams<TC> cams,
             const __grid_constant__ RayGeom<TC> geom, const __grid_constant__ HSParams hs,
             TO* __restrict__ x, uint8_t* __restrict__ status, TI* __restrict__ u1c, TI* __restrict__ u2c,
             unsigned int* __restrict__ not_nan_count, const int64_t n, const TC max_coord,
             const __grid_constant__ PRE pre_stage, const __grid_constant__ Mirrors mir,
             const __grid_constant__ EvalArg<EVAL> ev, const __grid_constant__ Deferred df) {
    double acc[4] = {0, 0, 0, 0};
    __shared__ TO stage[kWarps][96];
    __shared__ __align__(16) PairPrefetch<TI, kThreads> pre;
    const int warp = threadIdx.x >> 5;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;
    int64_t tile = static_cast<int64_t>(blockIdx.x) * kThreads;
    bool any1 = false, any2 = false;           // "not all NaN" bookkeeping for the fallback test, per thread
    pre.issue(u1, u2, tile + threadIdx.x, n);
    for (; tile < n; tile += stride) {
        const int64_t i = tile + threadIdx.x;
        TC a, b, c, d;
        pre.take(i, n, a, b, c, d);
        pre.issue(u1, u2, i + stride, n);
        if constexpr (PRE::kActive) pre_stage.template apply<TI, TC>(a, b, c, d);
        // The correction always runs in double: the degree-6 coefficients span many orders of magnitude.
        double n1x, n1y, n2x, n2y;
        bool certified = hs_correct_fast(hs, static_cast<double>(a), static_cast<double>(b), static_cast<double>(c),
                                           static_cast<double>(d), n1x, n1y, n2x, n2y);
        // cv2.correctMatches returns the dtype of its input, so the corrected points are rounded to TI
        // before the triangulation (triangulation.py:224,232).
        const TI r1x = static_cast<TI>(n1x), r1y = static_cast<TI>(n1y), r2x = static_cast<TI>(n2x), r2y = static_cast<TI>(n2y);
        TC xs[3] = {0, 0, 0};
        bool good = false;
        if (certified) {
            const bool nan_match = (n1x != n1x) || (n1y != n1y) || (n2x != n2x) || (n2y != n2y);
            if (nan_match) {
                // t = inf won / non-finite system: cv2.triangulatePoints of a NaN match is a NaN point, status False
                xs[0] = xs[1] = xs[2] = static_cast<TC>(n1x + n1y + n2x + n2y);
            } else {
                const TC res_tol = sizeof(TI) == 8 ? TC(1e-11) : TC(1e-7);
                certified = geom.ok && tworay_intersection<TC>(cams, geom, static_cast<TC>(r1x), static_cast<TC>(r1y),
                                                               static_cast<TC>(r2x), static_cast<TC>(r2y), res_tol, xs);
                good = tmax(tmax(tabs(xs[0]), tabs(xs[1])), tabs(xs[2])) <= max_coord;     // triangulation.py:23
            }
        }
        if (!certified && i < n) defer_point(df, i);
        store_x_warp(x, tile + warp * 32, n, static_cast<TO>(xs[0]), static_cast<TO>(xs[1]),
                         static_cast<TO>(xs[2]), stage[warp], mir);
        if (i < n && certified) {
            if (u1c) store_uv(u1c, i, r1x, r1y);
            if (u2c) store_uv(u2c, i, r2x, r2y);
            any1 = any1 || !(n1x != n1x) || !(n1y != n1y);
            any2 = any2 || !(n2x != n2x) || !(n2y != n2y);
            store_status(status, mir, i, static_cast<uint8_t>(good ? 1 : 0));
            // the harness / SLAM evaluate against the ORIGINAL observations, not the corrected ones
            fused_eval_point<EVAL, TO, TC>(ev, true, i, a, b, c, d, xs, good ? 1 : 0, acc);
        }
    }
    // one flag update per CTA (two words shared by the whole grid: per-warp atomics would all hit the same L2 line)
    const int f1 = __syncthreads_or(any1), f2 = __syncthreads_or(any2);
    if (threadIdx.x == 0) {
        if (f1) atomicOr(&not_nan_count[0], 1u);
        if (f2) atomicOr(&not_nan_count[1], 1u);
    }
    fused_eval_finish<EVAL>(ev, acc, df.ctl, df.cap);
}

#ifndef TRGL_POLY_GENERAL_MINB
#define TRGL_POLY_GENERAL_MINB 2
#endif
// Shared memory of one k_polynomial_general CTA: the subdivision of hs_interval_test for the CTA's 256 points, one dyadic
// level at a time.  Work items are intervals, not points: a point whose roots need 188 intervals (the worst of 10 M on the
// forward-motion rig; 18.6 on average) occupies many lanes for a few levels instead of one lane for 188 rounds.
constexpr int kIsoQueueCap = 2048;
struct IsoSmem {
    double p[8][kThreads];                          // k0..k6 of each thread's point; [7] = radius of its t-domain
    unsigned int queue[2][kIsoQueueCap];            // intervals of this / the next level: point << 24 | domain << 23 | position
    unsigned int roots[kThreads][kIsoMaxRoots];     // recorded single-root intervals: depth << 24 | domain << 23 | position
    unsigned int nroots[kThreads], giveup[kThreads], visited[kThreads];
    unsigned int count[2];
};

// Follow-up kernel of k_polynomial: the complete correction -- bracketed Newton under the monotonicity certificate, else
// certified isolation of all real roots (level-synchronous over the CTA), else Durand-Kerner as cv::solvePoly runs it --
// and the triangulation of the corrected match (certified ray intersection, else the eigen solver) for the deferred points,
// or for every point if the list overflowed / the two-ray forms are switched off (all != 0: k_polynomial is then not
// launched; no ray intersection).  A CTA takes 256 points at a time: phase A one point per thread (set-up, fast path),
// phase B the subdivision of the points that need it, phase C one point per thread again (root refinement, cost scan,
// closest points, triangulation, evaluation).
template <typename TI, typename TC, typename TO, int ROWS, class PRE = PreNone, bool EVAL = false>
__global__ void __launch_bounds__(kThreads, TRGL_POLY_GENERAL_MINB)
k_polynomial_general(const TI* __restrict__ u1, const TI* __restrict__ u2, const __grid_constant__ Cams<TC> cams,
                     const __grid_constant__ RayGeom<TC> geom, const __grid_constant__ HSParams hs, TO* __restrict__ x,
                     uint8_t* __restrict__ status, TI* __restrict__ u1c, TI* __restrict__ u2c,
                     unsigned int* __restrict__ not_nan_count, const int64_t n, const TC max_coord,
                     const __grid_constant__ PRE pre_stage, const __grid_constant__ Mirrors mir,
                     const __grid_constant__ EvalArg<EVAL> ev, const __grid_constant__ Deferred df, const int all) {
    wait_for_hot_kernel();
    __shared__ IsoSmem sm;
    const unsigned int tid = threadIdx.x;
    const unsigned int listed = all ? 0u : df.ctl[0];
    const bool everything = all || listed > df.cap;
    const int64_t total = everything ? n : static_cast<int64_t>(listed);
    const int64_t nchunks = (total + kThreads - 1) / kThreads;
    double acc[4] = {0, 0, 0, 0};
    bool any1 = false, any2 = false;
    for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const int64_t kidx = chunk * kThreads + tid;
        const bool have = kidx < total;
        const int64_t i = have ? (everything ? kidx : df.idx[kidx]) : 0;
        // ---- phase A: set-up and fast path, one point per thread ----
        if (tid < 2) sm.count[tid] = 0u;
        sm.nroots[tid] = 0u; sm.giveup[tid] = 0u; sm.visited[tid] = 0u;
        __syncthreads();
        TC a = 0, b = 0, c = 0, d = 0;
        HsPoint P;
        double kk[7];
        double t = DBL_MAX;
        bool subdivide = false;
        if (have) {
            reload_inputs<TI, TC, PRE>(u1, u2, pre_stage, i, a, b, c, d);
            hs_setup(hs, static_cast<double>(a), static_cast<double>(b), static_cast<double>(c), static_cast<double>(d), P, kk);
            if (P.fast) {
                t = hs_newton_bracketed(kk, P.T0);
            } else if (P.finite_coeffs) {
                // certificate failed (g not monotone on [-T0, T0], or no finite T0): every real root, t in [-R0, R0] on g
                // and, unless T0 <= 1, u = 1/t in [-1, 1] on the reversed polynomial
                subdivide = true;
#pragma unroll
                for (int j = 0; j < 7; ++j) sm.p[j][tid] = kk[j];
                sm.p[7][tid] = P.bounded ? fmin(P.T0, 1.0) : 1.0;
                const bool second = !(P.bounded && P.T0 <= 1.0);
                const unsigned int at = atomicAdd(&sm.count[0], second ? 2u : 1u);
                sm.queue[0][at] = tid << 24;
                if (second) sm.queue[0][at + 1] = (tid << 24) | (1u << 23);
            }
        }
        __syncthreads();
        // ---- phase B: level-synchronous subdivision, work items = intervals ----
        int cur = 0;
#pragma unroll 1
        for (int depth = 0; depth <= kIsoMaxDepth; ++depth) {
            const unsigned int na = sm.count[cur];
            if (na == 0u) break;                                           // block-uniform
            for (unsigned int q = tid; q < na; q += kThreads) {
                const unsigned int item = sm.queue[cur][q];
                const unsigned int prob = item >> 24, dom = (item >> 23) & 1u, pos = item & 0x7fffffu;
                double p[7];
#pragma unroll
                for (int j = 0; j < 7; ++j) p[j] = sm.p[dom ? 6 - j : j][prob];
                const int verdict = hs_interval_test(p, dom ? 1.0 : sm.p[7][prob], depth, pos);
                atomicAdd(&sm.visited[prob], 1u);
                if (verdict == 2) {
                    if (depth == kIsoMaxDepth) {
                        sm.giveup[prob] = 1u;
                    } else {
                        const unsigned int at = atomicAdd(&sm.count[cur ^ 1], 2u);       // pairs: never half a slot at the cap
                        if (at + 2u <= kIsoQueueCap) {
                            const unsigned int child = (item & 0xff800000u) | (pos << 1);
                            sm.queue[cur ^ 1][at] = child;
                            sm.queue[cur ^ 1][at + 1] = child | 1u;
                        } else {
                            sm.giveup[prob] = 1u;
                        }
                    }
                } else if (verdict == 1) {
                    const unsigned int r = atomicAdd(&sm.nroots[prob], 1u);
                    if (r < kIsoMaxRoots) sm.roots[prob][r] = (static_cast<unsigned int>(depth) << 24) | (item & 0xffffffu);
                    else sm.giveup[prob] = 1u;
                }
            }
            __syncthreads();
            if (tid == 0) {
                if (sm.count[cur ^ 1] > kIsoQueueCap) sm.count[cur ^ 1] = kIsoQueueCap;
                sm.count[cur] = 0u;
            }
            cur ^= 1;
            __syncthreads();
        }
        // ---- phase C: one point per thread again ----
        if (have) {
            if (subdivide) {
                const unsigned int nr = sm.nroots[tid];
                bool ok = sm.giveup[tid] == 0u && nr <= kIsoMaxRoots;
                if (ok) ok = hs_scan_roots(kk, sm.roots[tid], static_cast<int>(nr), sm.p[7][tid], P.bounded, P.a, P.b, P.c, P.d,
                                           P.f1, P.f2, t);
                atomicAdd(&g_hs_counters[0], 1ull);
                atomicAdd(&g_hs_counters[1], static_cast<unsigned long long>(sm.visited[tid]));
                atomicMax(&g_hs_counters[4], static_cast<unsigned long long>(sm.visited[tid]));
                if (!ok) {
                    atomicAdd(&g_hs_counters[2], 1ull);
                    t = hs_select_dk(kk, P.a, P.b, P.c, P.d, P.f1, P.f2);
                }
            }
            double n1x, n1y, n2x, n2y;
            hs_finish(P, t, n1x, n1y, n2x, n2y);
            const TI r1x = static_cast<TI>(n1x), r1y = static_cast<TI>(n1y), r2x = static_cast<TI>(n2x), r2y = static_cast<TI>(n2y);
            if (u1c) store_uv(u1c, i, r1x, r1y);
            if (u2c) store_uv(u2c, i, r2x, r2y);
            any1 = any1 || !(n1x != n1x) || !(n1y != n1y);
            any2 = any2 || !(n2x != n2x) || !(n2y != n2y);
            TC xs[3]; bool good;
            // as in the hot kernel: the corrected match satisfies the epipolar constraint, so the two viewing rays meet and
            // their certified intersection IS the smallest singular vector; the eigen solver (Rayleigh-quotient iteration,
            // Jacobi SVD for the low-parallax points of exactly the rigs that defer here) only for what that does not certify
            bool met = false;
            if (!all && geom.ok) {
                const TC res_tol = sizeof(TI) == 8 ? TC(1e-11) : TC(1e-7);
                met = tworay_intersection<TC>(cams, geom, static_cast<TC>(r1x), static_cast<TC>(r1y), static_cast<TC>(r2x),
                                              static_cast<TC>(r2y), res_tol, xs);
                good = tmax(tmax(tabs(xs[0]), tabs(xs[1])), tabs(xs[2])) <= max_coord;     // triangulation.py:23
            }
            if (!met)
                eigen_point<TC, ROWS>(cams, static_cast<TC>(r1x), static_cast<TC>(r1y), static_cast<TC>(r2x),
                                      static_cast<TC>(r2y), max_coord, xs, good);
#pragma unroll
            for (int q = 0; q < 3; ++q) x[3 * i + q] = static_cast<TO>(xs[q]);
            for (int m = 0; m < mir.count; ++m) {
#pragma unroll
                for (int q = 0; q < 3; ++q) mirror_put<TO>(mir, m, 3 * i + q, static_cast<TO>(xs[q]));
            }
            store_status(status, mir, i, static_cast<uint8_t>(good ? 1 : 0));
            fused_eval_point<EVAL, TO, TC>(ev, true, i, a, b, c, d, xs, good ? 1 : 0, acc);
        }
        __syncthreads();                                   // the next chunk reuses the shared arrays
    }
    const int f1 = __syncthreads_or(any1), f2 = __syncthreads_or(any2);
    if (threadIdx.x == 0) {
        if (f1) atomicOr(&not_nan_count[0], 1u);
        if (f2) atomicOr(&not_nan_count[1], 1u);
    }
    followup_finish<EVAL>(ev, df, acc, total > 0, !all);
}

// ---- cv2.undistortPoints as a standalone kernel (slam2.py:551-552; output dtype = input dtype) ---------------------
template <typename TI>
__global__ void __launch_bounds__(kThreads)
k_undistort_points(const TI* __restrict__ src, TI* __restrict__ dst, const __grid_constant__ Undist U, const int64_t n) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * kThreads) {
        double u, v, x, y;
        load_uv<double>(src, i, u, v);
        undistort_pair(U, u, v, x, y);
        store_uv(dst, i, x, y);
    }
}

}  // namespace trgl
