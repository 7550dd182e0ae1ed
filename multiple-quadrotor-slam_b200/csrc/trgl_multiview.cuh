// Multi-view (m >= 2 cameras) linear least-squares triangulation: SURVEY.md section 8f rank 4, the generalisation of
// linear_LS_triangulation (triangulation.c:65-83) from the 4x3 system of two views to the 2m x 3 system of m views
//     rows (u_v.x P_v[2,:] - P_v[0,:],  u_v.y P_v[2,:] - P_v[1,:])  for every view v that observes the point,
// solved in the same minimum-norm least-squares sense as cvSolve(DECOMP_SVD) (singular values <= 2 eps sum(w) dropped).
// The reference has no such function (every call is two-view); for m = 2 this kernel returns linear_LS's answer.
//
// One thread per point (two for m <= 4).  The 2m x 3 system is never materialised: views are streamed through registers
// in groups (8 vector loads in flight per thread) and folded into the 3x3 normal equations; well-conditioned points
// (kappa^2 bound < 1e6, Tiers<double>::t1, as in the two-view solver) finish with the adjugate solve.  The rest is deferred to a follow-up
// kernel (as in k_iterative_ls: no subroutine call in the hot kernel) that runs a streaming Givens QR -- each row is rotated into a 3x4 triangular factor [R | Q^T b], whose singular values are those
// of the full system -- followed by the 3x3 Jacobi SVD with OpenCV's rank rule, so rank-deficient and ill-conditioned
// points behave like the two-view careful path.
#pragma once
#include "trgl_device.cuh"
#include "trgl_kernels.cuh"      // Deferred / defer_point

namespace trgl {

constexpr int kMaxViews = 16;

template <typename TI, typename TC>
struct MultiViewArgs {
    const TI* u[kMaxViews];            // per view: (n,2) row-major observations (normalised coordinates)
    const uint8_t* valid[kMaxViews];   // per view: (n,) 1 = the view observes the point; NULL = every point
    TC P[kMaxViews][12];               // rows 0-2 of the camera matrices
    int m;
    int min_views;                     // status = (#valid views >= min_views)
};

// Fold one row [a | b] (a.x = b) into the upper-triangular 3x4 factor T = [R | c] with three Givens rotations.
template <typename T>
__device__ __forceinline__ void givens_fold(T (&Tm)[3][4], T r0, T r1, T r2, T rb) {
    T row[4] = {r0, r1, r2, rb};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const T p = Tm[k][k], q = row[k];
        const T h2 = tfma(p, p, q * q);
        if (!(h2 == T(0))) {                              // NaN rows must poison the factor, like the SVD of a NaN system
            const T ih = trsqrt(h2);
            const T c = p * ih, s = q * ih;
#pragma unroll
            for (int j = k; j < 4; ++j) {
                const T tk = Tm[k][j], tr = row[j];
                Tm[k][j] = tfma(c, tk, s * tr);
                row[j] = tfma(c, tr, -s * tk);
            }
        }
    }
}

template <typename TI, typename TC>
__device__ __forceinline__ void multiview_point_careful(const MultiViewArgs<TI, TC>& args, int64_t i, TC x[3]) {
    double Tm[3][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
    for (int v = 0; v < args.m; ++v) {
        if (args.valid[v] && !args.valid[v][i]) continue;
        double ux, uy;
        load_uv<double>(args.u[v], i, ux, uy);
        double r0[4], r1[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            r0[k] = fma(ux, static_cast<double>(args.P[v][8 + k]), -static_cast<double>(args.P[v][k]));
            r1[k] = fma(uy, static_cast<double>(args.P[v][8 + k]), -static_cast<double>(args.P[v][4 + k]));
        }
        givens_fold<double>(Tm, r0[0], r0[1], r0[2], -r0[3]);
        givens_fold<double>(Tm, r1[0], r1[1], r1[2], -r1[3]);
    }
    // min-norm solution of R x = c with the SVD rule: the singular values of R are those of the stacked system
    double rows[4][4];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int k = 0; k < 4; ++k) rows[r][k] = (k < 3) ? Tm[r][k] : -Tm[r][3];
#pragma unroll
    for (int k = 0; k < 4; ++k) rows[3][k] = 0.0;
    double xd[3];
    solve4x3_svd<double>(rows, xd);
#pragma unroll
    for (int k = 0; k < 3; ++k) x[k] = static_cast<TC>(xd[k]);
}

// PPT points per thread, views streamed in groups of GROUP: PPT * GROUP 16-byte loads (plus the mask bytes, which are
// independent loads: an observation is read whether or not its view is marked valid, a whole warp reads whole sectors
// anyway) are in flight per thread before the first row is built.  The launcher picks (2, 4) for m <= 4 and (1, 8) above.
// NGROUPS = ceil(m / GROUP) is a template parameter so that every view index is a compile-time constant: the camera
// matrices then are constant-bank operands of the FMA instructions instead of ~100 indexed constant loads per point.
// MASKED = false (no visibility mask given): no mask loads.  MASKED = true: valid[v] is non-NULL for every view.
template <typename TI, typename TC, typename TO, int PPT, int GROUP, int NGROUPS, bool MASKED>
__global__ void __launch_bounds__(kThreads)
k_multiview_ls(const __grid_constant__ MultiViewArgs<TI, TC> args, TO* __restrict__ x, uint8_t* __restrict__ status,
               const int64_t n, const __grid_constant__ Deferred df) {
    __shared__ TO stage[kWarps][96];
    const int warp = threadIdx.x >> 5;
    const NoMirrors nomir{};
    constexpr int TILE = kThreads * PPT;
    for (int64_t base = static_cast<int64_t>(blockIdx.x) * TILE; base < n; base += static_cast<int64_t>(gridDim.x) * TILE) {
        TC M[PPT][6], v3[PPT][3];
        int nviews[PPT];
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            nviews[p] = 0;
#pragma unroll
            for (int k = 0; k < 6; ++k) M[p][k] = TC(0);
#pragma unroll
            for (int k = 0; k < 3; ++k) v3[p][k] = TC(0);
        }
        // one group (m <= GROUP): every view index is a compile-time constant; more groups: a run-time loop over the groups
        // (unrolling two groups of eight views keeps 2 x 64 camera constants live: spills, 0.45 of the copy peak at m = 16)
#pragma unroll 1
        for (int g = 0; g < NGROUPS; ++g) {
            const int v0 = NGROUPS == 1 ? 0 : g * GROUP;
            TC in[PPT][GROUP][2];
            uint8_t ok[PPT][GROUP];
            // Straight-line loads: a view / point outside the batch reads element 0 of view 0 instead (and is switched off
            // below), so there is no branch between the loads and all PPT * GROUP observation loads (+ mask bytes) are in
            // flight together.  (With an `if (live)` block per view the compiler kept every mask load and the compare that
            // consumes it inside that block: eight dependent memory latencies per tile, 0.28 of the copy peak.)
#pragma unroll
            for (int p = 0; p < PPT; ++p) {
                const int64_t i = base + p * kThreads + threadIdx.x;
#pragma unroll
                for (int j = 0; j < GROUP; ++j) {
                    const int v = v0 + j;
                    const bool live = i < n && v < args.m;
                    const int64_t ii = live ? i : 0;
                    const int vv = live ? v : 0;
                    load_uv<TC>(args.u[vv], ii, in[p][j][0], in[p][j][1]);
                    uint8_t mk = 1;
                    if constexpr (MASKED) mk = __ldcs(args.valid[vv] + ii);
                    ok[p][j] = live ? mk : uint8_t(0);
                }
            }
#pragma unroll
            for (int p = 0; p < PPT; ++p) {
#pragma unroll
                for (int j = 0; j < GROUP; ++j) {
                    if (ok[p][j]) {
                        TC r0[4], r1[4];
                        dlt_rows<TC>(args.P[v0 + j], in[p][j][0], in[p][j][1], r0, r1);
                        normal_add2<TC>(r0, r1, M[p], v3[p]);
                        ++nviews[p];
                    }
                }
            }
        }
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            const int64_t i = base + p * kThreads + threadIdx.x;
            TC C[6], xs[3];
            const TC det = sym3_cofactors(M[p], C);
            const TC tr = M[p][0] + M[p][3] + M[p][5];
            sym3_apply(C, v3[p], fast_rcp(det), xs);
            if (nviews[p] <= 1) {
                // one view: M = A^T A has rank 2 and v = A^T b lies in its range, so the minimum-norm solution is
                // M^+ v = (tr(M) v - M v) / (sum of the principal 2x2 minors)   (Cayley-Hamilton on the range);
                // no view: x = 0.  Both are what the SVD rule returns; neither needs the follow-up kernel.
                const TC e2 = C[0] + C[3] + C[5];
                const TC inv = nviews[p] == 1 ? fast_rcp(e2) : TC(0);
                const TC Mv0 = tfma(M[p][0], v3[p][0], tfma(M[p][1], v3[p][1], M[p][2] * v3[p][2]));
                const TC Mv1 = tfma(M[p][1], v3[p][0], tfma(M[p][3], v3[p][1], M[p][4] * v3[p][2]));
                const TC Mv2 = tfma(M[p][2], v3[p][0], tfma(M[p][4], v3[p][1], M[p][5] * v3[p][2]));
                xs[0] = nviews[p] == 1 ? tfma(tr, v3[p][0], -Mv0) * inv : TC(0);
                xs[1] = nviews[p] == 1 ? tfma(tr, v3[p][1], -Mv1) * inv : TC(0);
                xs[2] = nviews[p] == 1 ? tfma(tr, v3[p][2], -Mv2) * inv : TC(0);
            } else if (i < n && !(tr * tr * tr < Tiers<TC>::t1() * det)) {
                // not tier 1 (ill-conditioned, rank-deficient, NaN): x comes from the follow-up kernel
                defer_point(df, i);
            }
            store_x_warp(x, base + p * kThreads + warp * 32, n, static_cast<TO>(xs[0]), static_cast<TO>(xs[1]),
                         static_cast<TO>(xs[2]), stage[warp], nomir);
            if (i < n) status[i] = static_cast<uint8_t>(nviews[p] >= args.min_views ? 1 : 0);
        }
    }
}

// Follow-up kernel: the deferred points (or every point if the list overflowed) through the streaming Givens QR + SVD.
template <typename TI, typename TC, typename TO>
__global__ void __launch_bounds__(kThreads)
k_multiview_general(const __grid_constant__ MultiViewArgs<TI, TC> args, TO* __restrict__ x, const int64_t n,
                    const __grid_constant__ Deferred df) {
    wait_for_hot_kernel();
    const unsigned int listed = df.ctl[0];
    const bool everything = listed > df.cap;
    const int64_t total = everything ? n : static_cast<int64_t>(listed);
    for (int64_t k = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; k < total;
         k += static_cast<int64_t>(gridDim.x) * kThreads) {
        const int64_t i = everything ? k : df.idx[k];
        TC xs[3];
        multiview_point_careful<TI, TC>(args, i, xs);
#pragma unroll
        for (int q = 0; q < 3; ++q) x[3 * i + q] = static_cast<TO>(xs[q]);
    }
    double none[4] = {0, 0, 0, 0};
    followup_finish<false>(EvalArg<false>{}, df, none, false, true);
}

}  // namespace trgl
