// Two-view evaluation of one triangulated point -- re-projection into both normalised cameras, squared errors, the
// good-point mask of the harness / SLAM keyframe step (triangulation_comparison.py:190-217,242-260; slam2.py:556,589) --
// and the block / grid reductions of its sums.  Shared by the stand-alone pass (k_pair_reproj) and by the solver kernels'
// fused epilogue, so both produce the same bits per point.
#pragma once
#include "trgl_device.cuh"

namespace trgl {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

template <int NV>
__device__ __forceinline__ void block_reduce_store(double (&v)[NV], double* __restrict__ partials) {
    __shared__ double sm[NV][kThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const double s = warp_sum(v[k]);
        if (lane == 0) sm[k][warp] = s;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) s += sm[threadIdx.x][w];
        partials[blockIdx.x * NV + threadIdx.x] = s;
    }
}

// Same, followed by the grid-level sum inside the kernel when `final_out` is given: the last CTA to arrive (ticket on
// `counter`, which it resets for the next launch) adds the per-block partials in block order -- the order the host uses --
// and writes NV doubles to device memory, so the caller needs no synchronisation to own the result in stream order.
// `deferred_ctl` / `deferred_cap` (hot solver kernels only): when more points were deferred than the list holds, the
// follow-up kernel redoes EVERY point and accumulates every point's evaluation, so the hot kernel's own sums must not
// be counted a second time -- the last CTA then writes zeros.  (Every CTA's list appends precede its ticket, so the
// last CTA sees the final count.)
template <int NV>
__device__ __forceinline__ void block_reduce_finalize(double (&v)[NV], double* __restrict__ partials,
                                                      unsigned int* __restrict__ counter, double* __restrict__ final_out,
                                                      const unsigned int* deferred_ctl = nullptr, unsigned int deferred_cap = 0u) {
    block_reduce_store<NV>(v, partials);
    if (final_out == nullptr) return;
    __shared__ bool is_last;
    __threadfence();                                   // the partials of this CTA are visible device-wide ...
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);     // ... before its ticket is
    __syncthreads();
    if (is_last) {
        if (threadIdx.x < NV) {
            double s = 0.0;
            for (unsigned int b = 0; b < gridDim.x; ++b) s += __ldcg(&partials[b * NV + threadIdx.x]);
            if (deferred_ctl != nullptr && __ldcg(deferred_ctl) > deferred_cap) s = 0.0;
            final_out[threadIdx.x] = s;
        }
        if (threadIdx.x == 0) *counter = 0u;
    }
}

// One point.  X,Y,Z are the STORED coordinates (already rounded to the output dtype) widened to double, a1..b2 the
// normalised observations widened to double: exactly what a separate pass would read back from HBM.
// acc: sum err1 (good), sum err2 (good), #good, #status > min_status.
template <typename TO>
__device__ __forceinline__ void eval_point(const Cams<double>& cams, int min_status, double max_sq_err,
                                           double a1, double b1, double a2, double b2, double X, double Y, double Z,
                                           int status, int64_t i, TO* __restrict__ err1, TO* __restrict__ err2,
                                           uint8_t* __restrict__ good, double (&acc)[4]) {
    double e[2], depth[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const double* P = c == 0 ? cams.P1 : cams.P2;
        const double px = fma(P[0], X, fma(P[1], Y, fma(P[2], Z, P[3])));
        const double py = fma(P[4], X, fma(P[5], Y, fma(P[6], Z, P[7])));
        const double pz = fma(P[8], X, fma(P[9], Y, fma(P[10], Z, P[11])));
        // MUFU.RCP64H seed + 2 Newton steps (<= 2 ulp) instead of the ~20-instruction IEEE division; pz == 0 or NaN gives a
        // non-finite error either way, hence good = false
        const double iz = fast_rcp(pz);
        const double dx = fma(px, iz, -(c == 0 ? a1 : a2)), dy = fma(py, iz, -(c == 0 ? b1 : b2));
        e[c] = fma(dx, dx, dy * dy);
        depth[c] = pz;
    }
    const bool st_ok = status > min_status;
    const bool g = st_ok && (e[0] <= max_sq_err) && (e[1] <= max_sq_err) && (depth[0] > 0.0) && (depth[1] > 0.0);
    if (err1) err1[i] = static_cast<TO>(e[0]);
    if (err2) err2[i] = static_cast<TO>(e[1]);
    if (good) good[i] = g ? 1 : 0;
    if (g) { acc[0] += e[0]; acc[1] += e[1]; acc[2] += 1.0; }
    if (st_ok) acc[3] += 1.0;
}

// Fused epilogue of the solver kernels (trgl_set_fused_eval): the evaluation runs on the registers the solve just
// produced, so the second pass over x, u1, u2 and status (57-60 B/point) disappears.
struct FusedEval {
    Cams<double> cams;             // the evaluation always projects in double, like the stand-alone pass
    double max_sq_err;
    int min_status;
    int pad_;
    void* err1; void* err2;        // (n,) of x's dtype, may be NULL
    uint8_t* good;                 // (n,), may be NULL
    double* partials;              // gridDim.x * 4 doubles of scratch
    unsigned int* counter;         // last-block ticket
    double* sums_out;              // 4 doubles, device
};

// Kernel parameter of the epilogue: empty when the kernel is instantiated without it.
template <bool EVAL> struct EvalArg;
template <> struct EvalArg<false> {};
template <> struct EvalArg<true> { FusedEval e; };

// Evaluate the point a solver thread has just produced (valid lanes only).
template <bool EVAL, typename TO, typename TC>
__device__ __forceinline__ void fused_eval_point(const EvalArg<EVAL>& ev, bool valid, int64_t i, TC a, TC b, TC c, TC d,
                                                 const TC xs[3], int status, double (&acc)[4]) {
    if constexpr (EVAL) {
        if (valid)
            eval_point<TO>(ev.e.cams, ev.e.min_status, ev.e.max_sq_err, static_cast<double>(a), static_cast<double>(b),
                           static_cast<double>(c), static_cast<double>(d),
                           static_cast<double>(static_cast<TO>(xs[0])), static_cast<double>(static_cast<TO>(xs[1])),
                           static_cast<double>(static_cast<TO>(xs[2])), status, i, static_cast<TO*>(ev.e.err1),
                           static_cast<TO*>(ev.e.err2), ev.e.good, acc);
    }
}
template <bool EVAL>
__device__ __forceinline__ void fused_eval_finish(const EvalArg<EVAL>& ev, double (&acc)[4],
                                                  const unsigned int* deferred_ctl = nullptr, unsigned int deferred_cap = 0u) {
    if constexpr (EVAL) block_reduce_finalize<4>(acc, ev.e.partials, ev.e.counter, ev.e.sums_out, deferred_ctl, deferred_cap);
}

}  // namespace trgl
