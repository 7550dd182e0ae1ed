// libtriangl_cuda: C ABI (include/triangl_cuda.h) over the sm_100a solver kernels.
// Host side only: argument checks, camera-matrix preparation, launch geometry, and the chunked
// H2D -> kernel -> D2H pipeline used when the caller hands in host buffers.  No CPU compute path exists.
#include "../../include/triangl_cuda.h"
#include "trgl_kernels.cuh"
#include "trgl_multiview.cuh"
#include "trgl_reproj.cuh"
#include "trgl_probe.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <utility>

using namespace trgl;

namespace {

thread_local std::string g_err = "";
std::atomic<int64_t> g_launches{0};
std::atomic<int> g_ppt{4};

// Result mirrors attached to the NEXT device-mode solver call of this host thread (trgl_set_result_mirrors).
thread_local Mirrors g_next_mirrors = {0, 0, {nullptr}, {nullptr}};
Mirrors take_mirrors() {
    Mirrors m = g_next_mirrors;
    g_next_mirrors.count = 0;
    return m;
}
const Mirrors kNoMirrors = {0, 0, {nullptr}, {nullptr}};

// Input retention requested for the NEXT host-mode solver call of this host thread (trgl_set_input_retention): the call
// uploads u1 / u2 into these device buffers instead of its pipeline scratch, so later calls can read them in place
// (TRGL_MEM_DEVICE_IN) -- "upload once, solve many".
struct PendingRetain { void* u1; void* u2; };
thread_local PendingRetain g_next_retain = {nullptr, nullptr};

// Fused evaluation requested for the NEXT device-mode solver call of this host thread (trgl_set_fused_eval).
struct PendingEval { bool armed; int min_status; double max_sq_err; void* err1; void* err2; uint8_t* good; double* sums; };
thread_local PendingEval g_next_eval = {false, 0, 0.0, nullptr, nullptr, nullptr, nullptr};

int fail(int code, const char* what) {
    g_err = what;
    return code;
}
int cuda_fail(cudaError_t e, const char* where) {
    g_err = std::string(where) + ": " + cudaGetErrorString(e);
    return static_cast<int>(e) > 0 ? static_cast<int>(e) : TRGL_E_NODEVICE;
}
#define CK(call)                                                    \
    do {                                                            \
        cudaError_t e__ = (call);                                   \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call);       \
    } while (0)

bool have_device() {
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess && n > 0;
}

struct ModeInfo { int in_bytes, out_bytes; };
bool mode_info(int mode, ModeInfo& m) {
    switch (mode) {
        case TRGL_F64: m = {8, 8}; return true;
        case TRGL_F32IO: m = {4, 4}; return true;
        case TRGL_F32: m = {4, 4}; return true;
        case TRGL_F64_OUT32: m = {8, 4}; return true;
        case TRGL_F32_OUT64: m = {4, 8}; return true;
    }
    return false;
}

template <typename T>
Cams<T> make_cams(const double* P1, const double* P2) {
    Cams<T> c;
    for (int i = 0; i < 12; ++i) { c.P1[i] = static_cast<T>(P1[i]); c.P2[i] = static_cast<T>(P2[i]); }
    return c;
}

// Camera centres (null vectors of P1, P2) and the cross terms E1 = P1 [C2;1], E2 = P2 [C1;1] of the two-ray closed form
// (trgl_kernels.cuh).  ok = 0 when a camera has no finite, well-defined centre (left 3x3 block singular to ~1e-12:
// affine / degenerate matrices): the kernel then runs the general path for every point.
template <typename T>
RayGeom<T> make_ray_geom(const double* P1, const double* P2) {
    RayGeom<T> g{};
    double C[2][3];
    bool ok = true;
    for (int cam = 0; cam < 2; ++cam) {
        const double* P = cam == 0 ? P1 : P2;
        const double m[9] = {P[0], P[1], P[2], P[4], P[5], P[6], P[8], P[9], P[10]};
        const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
        const double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
        double fro = 0;
        for (double v : m) fro += v * v;
        if (!(std::fabs(det) > 1e-12 * fro * std::sqrt(fro)) || !std::isfinite(det)) { ok = false; break; }
        const double inv[9] = {c00, m[2] * m[7] - m[1] * m[8], m[1] * m[5] - m[2] * m[4],
                               c01, m[0] * m[8] - m[2] * m[6], m[2] * m[3] - m[0] * m[5],
                               c02, m[1] * m[6] - m[0] * m[7], m[0] * m[4] - m[1] * m[3]};
        for (int r = 0; r < 3; ++r)
            C[cam][r] = -(inv[3 * r + 0] * P[3] + inv[3 * r + 1] * P[7] + inv[3 * r + 2] * P[11]) / det;
        // one refinement step of M C = -p4 (the cofactor inverse is only kappa(M) eps accurate)
        double res[3];
        for (int r = 0; r < 3; ++r)
            res[r] = -(P[4 * r + 3] + m[3 * r + 0] * C[cam][0] + m[3 * r + 1] * C[cam][1] + m[3 * r + 2] * C[cam][2]);
        for (int r = 0; r < 3; ++r)
            C[cam][r] += (inv[3 * r + 0] * res[0] + inv[3 * r + 1] * res[1] + inv[3 * r + 2] * res[2]) / det;
        for (int r = 0; r < 3; ++r) ok = ok && std::isfinite(C[cam][r]);
    }
    if (ok) {
        for (int r = 0; r < 3; ++r) {
            g.C1[r] = static_cast<T>(C[0][r]); g.C2[r] = static_cast<T>(C[1][r]);
            g.E1[r] = static_cast<T>(P1[4 * r + 0] * C[1][0] + P1[4 * r + 1] * C[1][1] + P1[4 * r + 2] * C[1][2] + P1[4 * r + 3]);
            g.E2[r] = static_cast<T>(P2[4 * r + 0] * C[0][0] + P2[4 * r + 1] * C[0][1] + P2[4 * r + 2] * C[0][2] + P2[4 * r + 3]);
        }
    }
    g.ok = ok ? 1 : 0;
    return g;
}

// Follow-up kernels are launched with programmatic dependent launch (PDL): the launch itself -- ~4 us of latency, as long as
// the hot kernel of a SLAM-sized batch runs -- overlaps the hot kernel, and the follow-up kernel waits at its first
// instruction (griddepcontrol.wait) until the hot kernel has completed and its writes (results, deferred list) are visible.
template <typename... KArgs, typename... Args>
void launch_followup(void (*kern)(KArgs...), unsigned grid, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);
}

// Grid cap of the follow-up kernels, in CTAs per SM: they loop over the deferred list with a grid stride.  Two per SM is what
// is resident anyway at their 100-150 registers per thread, and an EMPTY follow-up launch costs 4 us with 2 x 148 CTAs
// against 8-10 us with 8 x 148 (ncu launch list, profiles/r02p_launches.md) -- the common case is a list with no entry.
constexpr int kFollowupCtasPerSm = 2;
inline unsigned grid_for(int64_t n, int per_block) { return static_cast<unsigned>((n + per_block - 1) / per_block); }

// Reduction scratch (per-block partial sums, the polynomial NaN flags, the last-block ticket) is owned per
// (device, stream): calls enqueued on one stream reuse it in stream order, calls on different streams never share it,
// so the entry points are re-entrant per stream like the reference's stack-local C code (triangulation.c:67,106).
struct Scratch {
    double* partials = nullptr;       // kPartialDoubles
    unsigned int* flags = nullptr;    // 2 words per pipeline slot + 2 for device-mode calls (polynomial all-NaN test)
    unsigned int* counter = nullptr;  // ticket of the in-kernel final reduction (self-resetting)
    int64_t* deferred = nullptr;      // deferred_cap indices: points a hot kernel hands to its follow-up kernel
    unsigned int deferred_cap = 0;    // grows with the batch size: one slot per point up to kDeferredMax
    unsigned int* deferred_ctl = nullptr;   // {count, ticket, 64-bit running total}, re-armed by the follow-up kernel
    unsigned int* hint = nullptr;     // page-locked, device-visible: deferred count of the LAST call per solver kind (kHint*)
};
// How many points the last call of each solver deferred on this (device, stream): written by the follow-up kernel's last
// block into page-locked memory, read -- without any synchronisation, it is only a hint -- when the next follow-up kernel
// of that solver is sized.  Without it the follow-up grid is 2 CTAs per SM (an empty launch costs 4 us, 8-10 us at 8 per
// SM); a rig that defers 10^5-10^6 points (forward motion) then runs them in 1.5-2 rounds per thread instead of one wave.
enum { kHintLs = 0, kHintIter, kHintEigen, kHintPoly, kHintMultiview, kHintKinds = 8 };
constexpr unsigned int kDeferredMin = 1u << 20, kDeferredMax = 1u << 26;
std::atomic<unsigned int> g_deferred_limit{kDeferredMax};     // trgl_set_deferred_capacity (test knob)
int scratch_for(cudaStream_t s, Scratch& out, int64_t deferred_points = 0);
Deferred make_deferred(const Scratch& sc, int kind);
unsigned followup_grid(const Scratch& sc, int kind, int64_t tiles);

// ------------------------------------------------------------------------------------------------------------
// Device-pointer launchers, one per solver.  MODE_SWITCH instantiates the five precision modes.
// ------------------------------------------------------------------------------------------------------------
#define MODE_SWITCH(mode, ...)                                                              \
    switch (mode) {                                                                         \
        case TRGL_F64: { using TI = double; using TC = double; using TO = double; __VA_ARGS__ } break;        \
        case TRGL_F32IO: { using TI = float; using TC = double; using TO = float; __VA_ARGS__ } break;        \
        case TRGL_F32: { using TI = float; using TC = float; using TO = float; __VA_ARGS__ } break;           \
        case TRGL_F64_OUT32: { using TI = double; using TC = double; using TO = float; __VA_ARGS__ } break;   \
        case TRGL_F32_OUT64: { using TI = float; using TC = double; using TO = double; __VA_ARGS__ } break;   \
        default: return fail(TRGL_E_BADARG, "unknown precision mode");                      \
    }

// Launch geometry of the bulk-async variants: persistent grid of SMs x (CTAs that fit in shared memory).
std::atomic<int> g_variant{-1};         // -1 = auto, 0 = per-thread loads, >= 1 = bulk-async pipeline (trgl_set_stream_variant)

// Launch geometry is cached per (device, kernel instantiation): cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the
// occupancy query apply to the CURRENT device only, and one host thread may drive several devices (trgl_set_device).
constexpr int kMaxDevices = 64;
std::atomic<int> g_sm_by_device[kMaxDevices];
int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
    return dev;
}
int sm_count() {
    const int dev = current_device();
    if (dev < 0 || dev >= kMaxDevices) return 148;
    int n = g_sm_by_device[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) { cudaGetLastError(); n = 148; }
        g_sm_by_device[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}
// Resident CTAs per SM of `kern` at kThreads threads and `dyn_smem` bytes of dynamic shared memory on the current device
// (raises the kernel's dynamic shared-memory limit there first when it exceeds the 48 KB default).
template <typename K>
int blocks_per_sm(K kern, size_t dyn_smem) {
    static thread_local std::map<std::pair<int, const void*>, int> cache;
    int& per_sm = cache[std::make_pair(current_device(), reinterpret_cast<const void*>(kern))];
    if (per_sm == 0) {
        if (dyn_smem > 48 * 1024)
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(dyn_smem));
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, dyn_smem) != cudaSuccess || per_sm < 1) {
            cudaGetLastError();
            per_sm = 1;
        }
    }
    return per_sm;
}
std::atomic<int> g_two_ray{1};          // 0 = no two-ray closed forms in iterative_LS / polynomial (trgl_set_two_ray)

template <typename TI, typename TC, typename TO, int PPT, int STAGES, int MINB>
int launch_ls_tma(const TI* a, const TI* b, const Cams<TC>& cams, TO* xo, uint8_t* status, int64_t n, cudaStream_t s,
                  const Mirrors& mir) {
    constexpr int TILE = kThreads * PPT;
    const size_t smem = size_t(STAGES) * 2 * TILE * 2 * sizeof(TI) + kWarps * 96 * sizeof(TO) + STAGES * sizeof(uint64_t);
    auto kern = k_linear_ls_tma<TI, TC, TO, PPT, STAGES, MINB>;
    const int64_t full = static_cast<int64_t>(sm_count()) * blocks_per_sm(kern, smem);
    const int64_t ntiles = (n + TILE - 1) / TILE;
    const int64_t grid = ntiles < full ? ntiles : full;
    kern<<<static_cast<unsigned>(grid), kThreads, smem, s>>>(a, b, cams, xo, status, n, mir);
    return TRGL_OK;
}

// Persistent grid of the FP64-bound solvers: SMs x resident CTAs (occupancy API, cached per kernel), or fewer when
// the batch has fewer 256-point tiles than that.
template <typename K>
unsigned persistent_grid(K kern, int64_t n, size_t dyn_smem = 0, int per_cta = kThreads) {
    const int64_t tiles = (n + per_cta - 1) / per_cta;
    const int64_t full = static_cast<int64_t>(sm_count()) * blocks_per_sm(kern, dyn_smem);
    return static_cast<unsigned>(tiles < full ? tiles : full);
}

template <typename TI, typename TC, typename TO, int PPT, int DEPTH, int MINB>
int launch_ls_ring(const TI* a, const TI* b, const Cams<TC>& cams, TO* xo, uint8_t* status, int64_t n, cudaStream_t s,
                   const Mirrors& mir) {
    const size_t smem = size_t(DEPTH) * 2 * kThreads * PPT * 2 * sizeof(TI) + kWarps * 96 * sizeof(TO);
    auto kern = k_linear_ls_ring<TI, TC, TO, PPT, DEPTH, MINB>;
    const int64_t tiles = (n + kThreads * PPT - 1) / (kThreads * PPT);
    const int64_t full = static_cast<int64_t>(sm_count()) * blocks_per_sm(kern, smem);
    kern<<<static_cast<unsigned>(tiles < full ? tiles : full), kThreads, smem, s>>>(a, b, cams, xo, status, n, mir);
    return TRGL_OK;
}

// ---- pre-stage / fused-evaluation dispatch -----------------------------------------------------------------------
// f(prearg) is called with PreUndistort{...} or PreNone{}; g(std::bool_constant<EVAL>, EvalArg<EVAL>) likewise.  The
// fused evaluation is instantiated for the modes whose input and output storage types agree (F64, F32IO, F32).
template <typename F>
void with_pre(const Undist2* pre, F&& f) {
    if (pre) f(PreUndistort{*pre}); else f(PreNone{});
}
template <typename F>
void with_eval(const FusedEval* ev, F&& f) {
    if (ev) f(std::true_type{}, EvalArg<true>{*ev}); else f(std::false_type{}, EvalArg<false>{});
}
template <typename TI, typename TO, bool EV>
constexpr bool eval_supported() { return !EV || std::is_same<TI, TO>::value; }
int eval_unsupported() { return fail(TRGL_E_BADARG, "the fused evaluation needs u and x of the same storage type (F64, F32IO, F32)"); }

int launch_linear_ls(const void* u1, const void* u2, const double* P1, const double* P2, void* x, uint8_t* status,
                     int64_t n, int mode, cudaStream_t s, const Undist2* pre = nullptr, const Mirrors& mir = kNoMirrors,
                     const FusedEval* ev = nullptr) {
    if (n == 0) return TRGL_OK;
    const int ppt = g_ppt.load();
    int variant = g_variant.load();
    // auto: per-thread vector loads with 4 points in flight.  Measured on B200 (profiles/): 0.84-0.86 of the measured
    // copy peak at 10 M - 100 M points vs 0.74 for the bulk-async pipeline (variants 1-6) and 0.78 for the per-thread
    // cp.async ring (7-12), both kept selectable.
    // FP32 arithmetic mode: the per-thread cp.async ring (2 points/thread, 4 stages) is 1.18x faster (0.63 vs 0.54 of the
    // copy peak at 29 B/point; the kernel is FP32-issue bound there and the ring frees the load registers).
    if (variant < 0) variant = (mode == TRGL_F32) ? 8 : 0;
    // cp.async.bulk needs 16-byte aligned global addresses; fall back to per-thread loads otherwise
    if ((reinterpret_cast<uintptr_t>(u1) | reinterpret_cast<uintptr_t>(u2)) & 15) variant = 0;
    if (pre || ev) variant = 0;                    // the pre-stage and the fused evaluation live in the direct kernel
    int rc = TRGL_OK;
    // FP32 mode on the plain path: four consecutive points per thread with 128-bit accesses, float32 normal equations for
    // tier-1 points, everything else through the follow-up kernel (refinement step / double)
    if (mode == TRGL_F32 && g_variant.load() < 0 && !pre && !ev && mir.count == 0 && g_two_ray.load() &&
        !((reinterpret_cast<uintptr_t>(u1) | reinterpret_cast<uintptr_t>(u2) | reinterpret_cast<uintptr_t>(x)) & 15) &&
        !(reinterpret_cast<uintptr_t>(status) & 3)) {
        Scratch sc;
        rc = scratch_for(s, sc, n);
        if (rc) return rc;
        // The list holds at most n / 8 points here: a batch that defers more than that (low parallax throughout, e.g. the
        // forward-motion rig: every point) overflows it on purpose, and the follow-up kernel then redoes the WHOLE batch in
        // float64 with coalesced loads (~1.4 ms per 100 M points) instead of gathering 100 M listed points (4.0 ms).
        // Which path a batch takes depends only on its data; either result is within the FP32 mode's 1e-4.
        Deferred df = make_deferred(sc, kHintLs);
        df.cap = std::min<unsigned int>(df.cap, static_cast<unsigned int>(std::max<int64_t>(n / 8, 1024)));
        const Cams<float> cams = make_cams<float>(P1, P2);
        const float* a = static_cast<const float*>(u1); const float* b = static_cast<const float*>(u2);
        k_linear_ls_f32x4<<<grid_for((n + 3) / 4, kThreads), kThreads, 0, s>>>(a, b, cams, make_cams<double>(P1, P2), static_cast<float*>(x), status, n, df);
        launch_followup(k_linear_ls_general<float, double, float, PreNone, false>, followup_grid(sc, kHintLs, (n + kThreads - 1) / kThreads), s,
                        a, b, make_cams<double>(P1, P2), static_cast<float*>(x), n, PreNone{}, kNoMirrors, EvalArg<false>{}, df);
        g_launches += 2;
        CK(cudaGetLastError());
        return TRGL_OK;
    }
    MODE_SWITCH(mode, {
        const Cams<TC> cams = make_cams<TC>(P1, P2);
        const TI* a = static_cast<const TI*>(u1); const TI* b = static_cast<const TI*>(u2);
        TO* xo = static_cast<TO*>(x);
        switch (variant) {
            case 1: rc = launch_ls_tma<TI, TC, TO, 2, 4, 3>(a, b, cams, xo, status, n, s, mir); break;
            case 2: rc = launch_ls_tma<TI, TC, TO, 4, 3, 2>(a, b, cams, xo, status, n, s, mir); break;
            case 3: rc = launch_ls_tma<TI, TC, TO, 1, 6, 3>(a, b, cams, xo, status, n, s, mir); break;
            case 4: rc = launch_ls_tma<TI, TC, TO, 2, 6, 2>(a, b, cams, xo, status, n, s, mir); break;
            case 5: rc = launch_ls_tma<TI, TC, TO, 4, 4, 1>(a, b, cams, xo, status, n, s, mir); break;
            case 6: rc = launch_ls_tma<TI, TC, TO, 1, 8, 3>(a, b, cams, xo, status, n, s, mir); break;
            case 7: rc = launch_ls_ring<TI, TC, TO, 1, 8, 2>(a, b, cams, xo, status, n, s, mir); break;
            case 8: rc = launch_ls_ring<TI, TC, TO, 2, 4, 2>(a, b, cams, xo, status, n, s, mir); break;
            case 9: rc = launch_ls_ring<TI, TC, TO, 4, 2, 2>(a, b, cams, xo, status, n, s, mir); break;
            case 10: rc = launch_ls_ring<TI, TC, TO, 4, 3, 2>(a, b, cams, xo, status, n, s, mir); break;
            case 11: rc = launch_ls_ring<TI, TC, TO, 2, 4, 3>(a, b, cams, xo, status, n, s, mir); break;
            case 12: rc = launch_ls_ring<TI, TC, TO, 2, 6, 2>(a, b, cams, xo, status, n, s, mir); break;
            default: {
                // float64 arithmetic with 4 points per thread (or pixel inputs): the hot kernel defers what is beyond tier 1
                // to a follow-up kernel instead of redoing it in line (no subroutine call in the hot kernel)
                const bool defer = sizeof(TC) == 8 && g_two_ray.load() && (pre || ev || ppt == 4);
                Deferred df = {nullptr, nullptr, 0u, nullptr};
                Scratch sc;
                if (defer) {
                    rc = scratch_for(s, sc, n);
                    if (rc) break;
                    df = make_deferred(sc, kHintLs);
                }
                with_eval(ev, [&](auto E, auto evarg) {
                    constexpr bool EV = decltype(E)::value;
                    if constexpr (!eval_supported<TI, TO, EV>()) { rc = eval_unsupported(); } else {
                        // with the evaluation epilogue the kernel is a persistent grid-stride loop (its block partials are
                        // reduced at the end): exactly SMs x resident CTAs blocks, so that no SM idles through a partial wave
                        auto go = [&](auto kern, int per_block, auto prearg, auto mirarg) {
                            const unsigned grid = EV ? persistent_grid(kern, n, 0, per_block) : grid_for(n, per_block);
                            kern<<<grid, kThreads, 0, s>>>(a, b, cams, xo, status, n, prearg, mirarg, evarg, df);
                        };
                        if (pre) {      // pixel inputs: the undistortion makes the kernel FP64-bound, one point per thread
                            if (defer) go(k_linear_ls<TI, TC, TO, 1, PreUndistort, EV, Mirrors, true>, kThreads, PreUndistort{*pre}, mir);
                            else go(k_linear_ls<TI, TC, TO, 1, PreUndistort, EV>, kThreads, PreUndistort{*pre}, mir);
                        } else if (ppt == 4 && mir.count == 0) {
                            // THE hot path: no pre-stage, no mirrors compiled in
                            if (defer) go(k_linear_ls<TI, TC, TO, 4, PreNone, EV, NoMirrors, true>, kThreads * 4, PreNone{}, NoMirrors{});
                            else go(k_linear_ls<TI, TC, TO, 4, PreNone, EV, NoMirrors>, kThreads * 4, PreNone{}, NoMirrors{});
                        } else if (EV || ppt == 4) {
                            if (defer) go(k_linear_ls<TI, TC, TO, 4, PreNone, EV, Mirrors, true>, kThreads * 4, PreNone{}, mir);
                            else go(k_linear_ls<TI, TC, TO, 4, PreNone, EV>, kThreads * 4, PreNone{}, mir);
                        } else if (ppt == 2) {
                            k_linear_ls<TI, TC, TO, 2, PreNone, false><<<grid_for(n, kThreads * 2), kThreads, 0, s>>>(a, b, cams, xo, status, n, PreNone{}, mir, EvalArg<false>{}, df);
                        } else {
                            k_linear_ls<TI, TC, TO, 1, PreNone, false><<<grid_for(n, kThreads), kThreads, 0, s>>>(a, b, cams, xo, status, n, PreNone{}, mir, EvalArg<false>{}, df);
                        }
                        if (defer) {
                            const unsigned fgrid = followup_grid(sc, kHintLs, (n + kThreads - 1) / kThreads);
                            if (pre) launch_followup(k_linear_ls_general<TI, TC, TO, PreUndistort, EV>, fgrid, s, a, b, cams, xo, n, PreUndistort{*pre}, mir, evarg, df);
                            else launch_followup(k_linear_ls_general<TI, TC, TO, PreNone, EV>, fgrid, s, a, b, cams, xo, n, PreNone{}, mir, evarg, df);
                            g_launches++;
                        }
                    }
                });
            }
        }
    })
    if (rc) return rc;
    g_launches++;
    CK(cudaGetLastError());
    return TRGL_OK;
}

int launch_iterative_ls(const void* u1, const void* u2, const double* P1, const double* P2, void* x, int32_t* status,
                        int64_t n, double tol, int semantics, int mode, cudaStream_t s, const Undist2* pre = nullptr,
                        const Mirrors& mir = kNoMirrors, const FusedEval* ev = nullptr) {
    if (n == 0) return TRGL_OK;
    if (mode == TRGL_F32) mode = TRGL_F32IO;     // float32 storage, float64 registers (see header)
    Scratch sc;
    int rc = scratch_for(s, sc, n);
    if (rc) return rc;
    const Deferred df = make_deferred(sc, kHintIter);
    const int py = semantics == TRGL_ITER_PY ? 1 : 0;
    int nlaunch = 0;
    MODE_SWITCH(mode, {
        const Cams<TC> cams = make_cams<TC>(P1, P2);
        const RayGeom<TC> geom = make_ray_geom<TC>(P1, P2);
        const bool closed_form = geom.ok && g_two_ray.load();
        constexpr size_t smem = sizeof(IterSmem<TI, TC, TO>);
        const TI* a = static_cast<const TI*>(u1); const TI* b = static_cast<const TI*>(u2);
        with_pre(pre, [&](auto prearg) {
            using PRE = decltype(prearg);
            with_eval(ev, [&](auto E, auto evarg) {
                constexpr bool EV = decltype(E)::value;
                if constexpr (!eval_supported<TI, TO, EV>()) { rc = eval_unsupported(); } else {
                    auto general = k_iterative_general<TI, TC, TO, PRE, EV>;
                    if (closed_form) {
                        // hot kernel (two-ray closed form) + follow-up over the points it deferred (usually none)
                        auto kern = k_iterative_ls<TI, TC, TO, PRE, EV>;
                        kern<<<persistent_grid(kern, n, smem), kThreads, smem, s>>>(
                            a, b, cams, geom, static_cast<TO*>(x), status, n, static_cast<TC>(tol), py, prearg, mir, evarg, df);
                        launch_followup(general, followup_grid(sc, kHintIter, (n + kThreads - 1) / kThreads), s,
                                        a, b, cams, static_cast<TO*>(x), status, n, static_cast<TC>(tol), py, prearg, mir, evarg, df, 0);
                        nlaunch = 2;
                    } else {
                        // no finite camera centre / closed forms switched off: the reference's loop for every point
                        if constexpr (EV) cudaMemsetAsync(evarg.e.sums_out, 0, 4 * sizeof(double), s);
                        general<<<persistent_grid(general, n), kThreads, 0, s>>>(
                            a, b, cams, static_cast<TO*>(x), status, n, static_cast<TC>(tol), py, prearg, mir, evarg, df, 1);
                        nlaunch = 1;
                    }
                }
            });
        });
    })
    if (rc) return rc;
    g_launches += nlaunch;
    CK(cudaGetLastError());
    return TRGL_OK;
}

int launch_linear_eigen(const void* u1, const void* u2, const double* P1, const double* P2, void* x, uint8_t* status,
                        int64_t n, double maxc, int rows, int mode, cudaStream_t s, const Undist2* pre = nullptr,
                        const Mirrors& mir = kNoMirrors, const FusedEval* ev = nullptr) {
    if (n == 0) return TRGL_OK;
    if (mode == TRGL_F32) mode = TRGL_F32IO;     // float32 storage, float64 registers (see header)
    Scratch sc;
    int rc = scratch_for(s, sc, n);
    if (rc) return rc;
    const Deferred df = make_deferred(sc, kHintEigen);
    MODE_SWITCH(mode, {
        const Cams<TC> cams = make_cams<TC>(P1, P2);
        const TI* a = static_cast<const TI*>(u1); const TI* b = static_cast<const TI*>(u2);
        const int64_t tiles = (n + kThreads - 1) / kThreads;
        with_pre(pre, [&](auto prearg) {
            using PRE = decltype(prearg);
            with_eval(ev, [&](auto E, auto evarg) {
                constexpr bool EV = decltype(E)::value;
                if constexpr (!eval_supported<TI, TO, EV>()) { rc = eval_unsupported(); } else {
                    // hot kernel (Rayleigh-quotient iteration, certified) + follow-up over the points it deferred (Jacobi SVD)
                    auto launch = [&](auto kern, auto general) {
                        kern<<<persistent_grid(kern, n), kThreads, 0, s>>>(a, b, cams, static_cast<TO*>(x), status, n, static_cast<TC>(maxc), prearg, mir, evarg, df);
                        launch_followup(general, followup_grid(sc, kHintEigen, tiles), s, a, b, cams, static_cast<TO*>(x), status, n, static_cast<TC>(maxc), prearg, mir, evarg, df);
                    };
                    if (rows == 4) launch(k_linear_eigen<TI, TC, TO, 4, PRE, EV>, k_linear_eigen_general<TI, TC, TO, 4, PRE, EV>);
                    else launch(k_linear_eigen<TI, TC, TO, 6, PRE, EV>, k_linear_eigen_general<TI, TC, TO, 6, PRE, EV>);
                }
            });
        });
    })
    if (rc) return rc;
    g_launches += 2;
    CK(cudaGetLastError());
    return TRGL_OK;
}

int launch_polynomial(const void* u1, const void* u2, const double* P1, const double* P2, const HSParams& hs, void* x,
                      uint8_t* status, void* u1c, void* u2c, unsigned int* flags, int64_t n, double maxc, int rows,
                      int mode, cudaStream_t s, const Undist2* pre = nullptr, const Mirrors& mir = kNoMirrors,
                      const FusedEval* ev = nullptr) {
    if (n == 0) return TRGL_OK;
    if (mode == TRGL_F32) mode = TRGL_F32IO;     // float32 storage, float64 registers (see header)
    Scratch sc;
    int rc = scratch_for(s, sc, n);
    if (rc) return rc;
    const Deferred df = make_deferred(sc, kHintPoly);
    int nlaunch = 0;
    MODE_SWITCH(mode, {
        const Cams<TC> cams = make_cams<TC>(P1, P2);
        const RayGeom<TC> geom = make_ray_geom<TC>(P1, P2);
        const bool closed_form = geom.ok && g_two_ray.load();
        const TI* a = static_cast<const TI*>(u1); const TI* b = static_cast<const TI*>(u2);
        const int64_t tiles = (n + kThreads - 1) / kThreads;
        with_pre(pre, [&](auto prearg) {
            using PRE = decltype(prearg);
            with_eval(ev, [&](auto E, auto evarg) {
                constexpr bool EV = decltype(E)::value;
                if constexpr (!eval_supported<TI, TO, EV>()) { rc = eval_unsupported(); } else {
                    auto launch = [&](auto kern, auto general) {
                        if (closed_form) {
                            // hot kernel (certified correction + ray intersection) + follow-up over the points it deferred
                            kern<<<persistent_grid(kern, n), kThreads, 0, s>>>(a, b, cams, geom, hs, static_cast<TO*>(x), status, static_cast<TI*>(u1c), static_cast<TI*>(u2c), flags, n, static_cast<TC>(maxc), prearg, mir, evarg, df);
                            launch_followup(general, followup_grid(sc, kHintPoly, tiles), s, a, b, cams, geom, hs, static_cast<TO*>(x), status, static_cast<TI*>(u1c), static_cast<TI*>(u2c), flags, n, static_cast<TC>(maxc), prearg, mir, evarg, df, 0);
                            nlaunch = 2;
                        } else {
                            // no finite camera centre / closed forms switched off: the complete per-point path for every point
                            if constexpr (EV) cudaMemsetAsync(evarg.e.sums_out, 0, 4 * sizeof(double), s);
                            general<<<persistent_grid(general, n), kThreads, 0, s>>>(a, b, cams, geom, hs, static_cast<TO*>(x), status, static_cast<TI*>(u1c), static_cast<TI*>(u2c), flags, n, static_cast<TC>(maxc), prearg, mir, evarg, df, 1);
                            nlaunch = 1;
                        }
                    };
                    if (rows == 4) launch(k_polynomial<TI, TC, TO, 4, PRE, EV>, k_polynomial_general<TI, TC, TO, 4, PRE, EV>);
                    else launch(k_polynomial<TI, TC, TO, 6, PRE, EV>, k_polynomial_general<TI, TC, TO, 6, PRE, EV>);
                }
            });
        });
    })
    if (rc) return rc;
    g_launches += nlaunch;
    CK(cudaGetLastError());
    return TRGL_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Host-buffer pipeline: chunks of kChunk points round-robin over kSlots streams, each slot owning a grow-only
// device scratch block.  With pinned host buffers (trgl_host_alloc) H2D, kernel and D2H of neighbouring chunks
// overlap; with pageable buffers the copies serialise but the result is the same.
// ------------------------------------------------------------------------------------------------------------
constexpr int kSlots = 4;
constexpr int64_t kChunk = 1 << 19;     // 512 Ki points = 16 MB H2D + 12.5 MB D2H per chunk (FP64): pipeline fill/drain
                                        // (first H2D + last D2H, not overlapped) costs ~0.5 ms instead of ~2 ms at 2 Mi

struct Slot {
    cudaStream_t stream = nullptr;
    char* buf = nullptr;
    size_t cap = 0;
    int device = -1;
};
std::mutex g_pipe_mutex;
Slot g_slots[kSlots];
std::mutex g_host_mutex;              // host-mode calls that keep state across host_pipeline (polynomial flags)

constexpr size_t kPartialDoubles = static_cast<size_t>(kReduceBlocks) * 48;
constexpr int kFlagWords = 2 * (kSlots + 1);
std::mutex g_scratch_mutex;
std::map<std::pair<int, cudaStream_t>, Scratch> g_scratch;

int scratch_for(cudaStream_t s, Scratch& out, int64_t deferred_points) {
    int dev = 0;
    CK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_scratch_mutex);
    Scratch& sc = g_scratch[std::make_pair(dev, s)];
    if (!sc.partials) {
        if (cudaMalloc(&sc.partials, kPartialDoubles * sizeof(double)) != cudaSuccess) {
            cudaGetLastError();
            return fail(TRGL_E_NOMEM, "cudaMalloc of reduction scratch failed");
        }
        if (cudaMalloc(&sc.flags, sizeof(unsigned int) * (kFlagWords + 8)) != cudaSuccess) {
            cudaGetLastError(); cudaFree(sc.partials); sc.partials = nullptr;
            return fail(TRGL_E_NOMEM, "cudaMalloc of reduction scratch failed");
        }
        CK(cudaMemset(sc.flags, 0, sizeof(unsigned int) * (kFlagWords + 8)));
        sc.counter = sc.flags + kFlagWords;
        sc.deferred_ctl = sc.flags + kFlagWords + 2;
        if (cudaHostAlloc(reinterpret_cast<void**>(&sc.hint), sizeof(unsigned int) * kHintKinds, cudaHostAllocPortable | cudaHostAllocMapped) == cudaSuccess) {
            std::memset(sc.hint, 0, sizeof(unsigned int) * kHintKinds);
        } else {
            cudaGetLastError(); sc.hint = nullptr;      // no hint: fixed follow-up grids
        }
    }
    if (deferred_points > 0) {
        // one slot per point of the batch (a rig can defer most of its points, e.g. forward motion under heavy noise)
        int64_t want = deferred_points < kDeferredMin ? kDeferredMin : deferred_points;
        if (want > kDeferredMax) want = kDeferredMax;
        if (sc.deferred_cap < want) {
            if (sc.deferred) { CK(cudaStreamSynchronize(s)); cudaFree(sc.deferred); sc.deferred = nullptr; sc.deferred_cap = 0; }
            if (cudaMalloc(&sc.deferred, sizeof(int64_t) * want) != cudaSuccess) {
                cudaGetLastError();
                return fail(TRGL_E_NOMEM, "cudaMalloc of the deferred-point list failed");
            }
            sc.deferred_cap = static_cast<unsigned int>(want);
        }
    }
    out = sc;
    return TRGL_OK;
}

Deferred make_deferred(const Scratch& sc, int kind) {
    return Deferred{sc.deferred, sc.deferred_ctl, std::min(sc.deferred_cap, g_deferred_limit.load()), sc.hint ? sc.hint + kind : nullptr};
}
unsigned followup_grid(const Scratch& sc, int kind, int64_t tiles) {
    const int64_t sms = sm_count();
    int64_t cap = kFollowupCtasPerSm * sms;
    // Measured (profiles/r02u_sweep_rigs_10M.jsonl): one wave instead of 1.5-2 rounds per thread pays where a deferred point
    // is expensive -- polynomial's Durand-Kerner sweeps, ~5e4 instructions: forward-motion rig 1.37 -> 1.20 ms per 10 M points
    // -- and costs 3 % where it is cheap (linear_eigen, ~1e3 instructions: the larger launch outweighs the shorter tail).
    if (sc.hint && kind == kHintPoly) {
        const unsigned int last = reinterpret_cast<volatile unsigned int*>(sc.hint)[kind];
        int64_t want = (static_cast<int64_t>(last) + kThreads - 1) / kThreads;
        want += want / 8;
        cap = std::max(cap, std::min(want, 16 * sms));
    }
    return static_cast<unsigned>(tiles < cap ? tiles : cap);
}

// Result mirrors / fused evaluation apply to ONE solver call: whatever that call's outcome (argument error, n == 0,
// host mode), the request is gone when it returns.
struct PendingClear {
    ~PendingClear() { g_next_mirrors.count = 0; g_next_eval.armed = false; g_next_retain = {nullptr, nullptr}; }
};

// Consume the pending fused-evaluation request: cameras of this call, reduction scratch of this stream.
int take_eval(const double* P1, const double* P2, cudaStream_t s, FusedEval& fe, const FusedEval*& out) {
    out = nullptr;
    if (!g_next_eval.armed) return TRGL_OK;
    const PendingEval pe = g_next_eval;
    g_next_eval.armed = false;
    Scratch sc;
    int rc = scratch_for(s, sc);
    if (rc) return rc;
    for (int i = 0; i < 12; ++i) { fe.cams.P1[i] = P1[i]; fe.cams.P2[i] = P2[i]; }
    fe.max_sq_err = pe.max_sq_err; fe.min_status = pe.min_status; fe.pad_ = 0;
    fe.err1 = pe.err1; fe.err2 = pe.err2; fe.good = pe.good;
    fe.partials = sc.partials; fe.counter = sc.counter; fe.sums_out = pe.sums;
    out = &fe;
    return TRGL_OK;
}

int ensure_slot(Slot& sl, size_t bytes) {
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (sl.device != dev) {           // device changed: drop the old scratch
        if (sl.buf) { cudaFree(sl.buf); sl.buf = nullptr; sl.cap = 0; }
        if (sl.stream) { cudaStreamDestroy(sl.stream); sl.stream = nullptr; }
        sl.device = dev;
    }
    if (!sl.stream) CK(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
    if (sl.cap < bytes) {
        if (sl.buf) { CK(cudaStreamSynchronize(sl.stream)); CK(cudaFree(sl.buf)); sl.buf = nullptr; sl.cap = 0; }
        size_t want = bytes + bytes / 8 + 4096;
        if (cudaMalloc(&sl.buf, want) != cudaSuccess) return fail(TRGL_E_NOMEM, "cudaMalloc of pipeline scratch failed");
        sl.cap = want;
    }
    return TRGL_OK;
}

inline size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

// One array of a host-mode call: `in` is copied to the device before the launch, `out` is filled from the device after
// it.  `dev` (optional): the array lives in this caller-owned device buffer instead of the pipeline scratch -- with `in`
// it is uploaded there (input retention), without `in` it is already resident (TRGL_MEM_DEVICE_IN).
struct HostArray { const void* in; void* out; size_t bytes_per_point; void* dev = nullptr; };
// u1 / u2 of a solver call in host mode, input-retention mode or device-in mode.
inline HostArray input_array(const void* u, size_t bytes_per_point, int mem, void* retain) {
    if (mem == TRGL_MEM_DEVICE_IN) return HostArray{nullptr, nullptr, bytes_per_point, const_cast<void*>(u)};
    return HostArray{u, nullptr, bytes_per_point, retain};
}
constexpr int kMaxHostArrays = 2 * kMaxViews + 2;     // multi-view: m observation arrays + m masks + x + status

// Small batches (the SLAM keyframe sizes, slam2.py:1080-1082: a few hundred points): latency is all API calls, so the
// kernel works straight on page-locked, device-mapped host memory (UVA: the cudaHostAlloc pointer is valid on the
// device).  Inputs are memcpy'd into the staging block by the CPU, ONE kernel reads them over PCIe and writes x / status
// back into the same block, one stream synchronise, CPU memcpy out: 1 launch + 1 sync instead of 4 copies + launch + sync.
constexpr int64_t kZeroCopyMax = 32768;
// trgl_set_trace(1): accumulate the host-side time of the small-batch path per phase (staging memcpy in, kernel launches,
// stream synchronise, memcpy out, number of calls) -- read and cleared by trgl_get_trace (bench.py --workload slam).
std::atomic<int> g_trace{0};
double g_trace_us[5] = {0, 0, 0, 0, 0};
char* g_zc_buf = nullptr;
size_t g_zc_cap = 0;

int ensure_zero_copy(size_t bytes) {
    if (g_zc_cap >= bytes) return TRGL_OK;
    if (g_zc_buf) { cudaFreeHost(g_zc_buf); g_zc_buf = nullptr; g_zc_cap = 0; }
    const size_t want = bytes < (size_t(1) << 20) ? (size_t(1) << 20) : bytes + bytes / 4;
    if (cudaHostAlloc(reinterpret_cast<void**>(&g_zc_buf), want, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
        cudaGetLastError();
        return fail(TRGL_E_NOMEM, "cudaHostAlloc of the zero-copy staging block failed");
    }
    g_zc_cap = want;
    return TRGL_OK;
}

// launcher(dev pointers in the order of `arrays`, chunk point count, stream)
template <typename Launch>
int host_pipeline(HostArray* arrays, int narrays, int64_t n, Launch launch) {
    std::lock_guard<std::mutex> lock(g_pipe_mutex);
    bool caller_device_arrays = false;
    for (int a = 0; a < narrays; ++a) caller_device_arrays = caller_device_arrays || arrays[a].dev != nullptr;
    if (n <= kZeroCopyMax && !caller_device_arrays) {
        const bool trace = g_trace.load(std::memory_order_relaxed) != 0;
        const auto t0 = std::chrono::steady_clock::now();
        size_t total = 0;
        for (int a = 0; a < narrays; ++a) total += align256(arrays[a].bytes_per_point * n);
        int rc = ensure_zero_copy(total);
        if (rc) return rc;
        rc = ensure_slot(g_slots[0], 0);
        if (rc) return rc;
        void* dptr[kMaxHostArrays];
        size_t pos = 0;
        for (int a = 0; a < narrays; ++a) {
            dptr[a] = g_zc_buf + pos;
            pos += align256(arrays[a].bytes_per_point * n);
            if (arrays[a].in) std::memcpy(dptr[a], arrays[a].in, arrays[a].bytes_per_point * n);
        }
        const auto t1 = std::chrono::steady_clock::now();
        rc = launch(dptr, n, g_slots[0].stream, 0);
        if (rc) return rc;
        const auto t2 = std::chrono::steady_clock::now();
        CK(cudaStreamSynchronize(g_slots[0].stream));
        const auto t3 = std::chrono::steady_clock::now();
        for (int a = 0; a < narrays; ++a)
            if (arrays[a].out) std::memcpy(arrays[a].out, dptr[a], arrays[a].bytes_per_point * n);
        if (trace) {
            const auto t4 = std::chrono::steady_clock::now();
            auto us = [](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
            g_trace_us[0] += us(t0, t1); g_trace_us[1] += us(t1, t2); g_trace_us[2] += us(t2, t3); g_trace_us[3] += us(t3, t4);
            g_trace_us[4] += 1.0;
        }
        return TRGL_OK;
    }
    const int64_t chunk = n < kChunk ? n : kChunk;
    size_t per_chunk = 0;
    for (int a = 0; a < narrays; ++a) per_chunk += align256(arrays[a].bytes_per_point * chunk);
    const int nslots = n > chunk ? kSlots : 1;
    for (int s = 0; s < nslots; ++s) { int rc = ensure_slot(g_slots[s], per_chunk); if (rc) return rc; }
    // Any failure inside the loop still drains every slot before returning: earlier chunks have asynchronous copies in
    // flight into the CALLER's buffers, which the caller may free as soon as it sees the error code.
    int rc = TRGL_OK;
    int idx = 0;
    for (int64_t off = 0; off < n && rc == TRGL_OK; off += chunk, ++idx) {
        Slot& sl = g_slots[idx % nslots];
        const int64_t m = (n - off) < chunk ? (n - off) : chunk;
        void* dptr[kMaxHostArrays];
        size_t pos = 0;
        for (int a = 0; a < narrays && rc == TRGL_OK; ++a) {
            dptr[a] = arrays[a].dev ? static_cast<char*>(arrays[a].dev) + off * arrays[a].bytes_per_point : sl.buf + pos;
            pos += align256(arrays[a].bytes_per_point * chunk);
            if (arrays[a].in) {
                const cudaError_t e = cudaMemcpyAsync(dptr[a], static_cast<const char*>(arrays[a].in) + off * arrays[a].bytes_per_point,
                                                      arrays[a].bytes_per_point * m, cudaMemcpyHostToDevice, sl.stream);
                if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemcpyAsync (host pipeline, H2D)");
            }
        }
        if (rc == TRGL_OK) rc = launch(dptr, m, sl.stream, idx % nslots);
        for (int a = 0; a < narrays && rc == TRGL_OK; ++a)
            if (arrays[a].out) {
                const cudaError_t e = cudaMemcpyAsync(static_cast<char*>(arrays[a].out) + off * arrays[a].bytes_per_point, dptr[a],
                                                      arrays[a].bytes_per_point * m, cudaMemcpyDeviceToHost, sl.stream);
                if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemcpyAsync (host pipeline, D2H)");
            }
    }
    for (int s = 0; s < nslots; ++s) {
        const cudaError_t e = cudaStreamSynchronize(g_slots[s].stream);
        if (e != cudaSuccess && rc == TRGL_OK) rc = cuda_fail(e, "cudaStreamSynchronize (host pipeline)");
    }
    return rc;
}

int check_common(const void* u1, const void* u2, const double* P1, const double* P2, const void* x, const void* status,
                 int64_t n, int mode, int mem, void* stream) {
    ModeInfo mi;
    if (n < 0) return fail(TRGL_E_BADARG, "negative point count");
    if (!mode_info(mode, mi)) return fail(TRGL_E_BADARG, "unknown precision mode");
    if (mem != TRGL_MEM_HOST && mem != TRGL_MEM_DEVICE && mem != TRGL_MEM_DEVICE_IN) return fail(TRGL_E_BADARG, "unknown memory space");
    if (!P1 || !P2) return fail(TRGL_E_BADARG, "camera matrix pointer is NULL");
    if (n > 0 && (!u1 || !u2 || !x || !status)) return fail(TRGL_E_BADARG, "NULL array pointer with n > 0");
    if (g_next_retain.u1 && mem != TRGL_MEM_HOST) {
        g_next_retain = {nullptr, nullptr};
        return fail(TRGL_E_BADARG, "input retention applies to host-mode calls (TRGL_MEM_HOST)");
    }
    if (g_next_retain.u1 && n > 0 &&
        ((reinterpret_cast<uintptr_t>(g_next_retain.u1) | reinterpret_cast<uintptr_t>(g_next_retain.u2)) & static_cast<uintptr_t>(2 * mi.in_bytes - 1))) {
        g_next_retain = {nullptr, nullptr};
        return fail(TRGL_E_BADARG, "retention buffers must be aligned to one (x,y) pair");
    }
    // the kernels read one (x,y) pair per vector load / cp.async: device buffers must be aligned to a pair
    if (mem != TRGL_MEM_HOST && n > 0 &&
        ((reinterpret_cast<uintptr_t>(u1) | reinterpret_cast<uintptr_t>(u2)) & static_cast<uintptr_t>(2 * mi.in_bytes - 1)))
        return fail(TRGL_E_BADARG, "device u1/u2 must be aligned to one (x,y) pair (16 bytes float64, 8 bytes float32)");
    if (!have_device()) return fail(TRGL_E_NODEVICE, "no CUDA device available (libtriangl_cuda has no CPU fallback)");
    if (g_next_mirrors.count && (mem != TRGL_MEM_DEVICE || n == 0)) {
        g_next_mirrors.count = 0;
        if (mem != TRGL_MEM_DEVICE) return fail(TRGL_E_BADARG, "result mirrors need device buffers (TRGL_MEM_DEVICE)");
    }
    if (g_next_eval.armed && (mem != TRGL_MEM_DEVICE || n == 0)) {
        const PendingEval pe = g_next_eval;
        g_next_eval.armed = false;
        if (mem != TRGL_MEM_DEVICE) return fail(TRGL_E_BADARG, "the fused evaluation needs device buffers (TRGL_MEM_DEVICE)");
        cudaMemsetAsync(pe.sums, 0, 4 * sizeof(double), static_cast<cudaStream_t>(stream));   // n == 0: all sums are zero, in the call's stream order
    }
    return TRGL_OK;
}

// Parameters of cv2.undistortPoints: K (3x3 row-major) and (k1,k2,p1,p2,k3) or NULL.  OpenCV multiplies by 1/fx.
bool make_undist(const double* K, const double* dist, Undist& U) {
    if (!K || K[0] == 0.0 || K[4] == 0.0) return false;
    U.ifx = 1.0 / K[0]; U.ify = 1.0 / K[4]; U.cx = K[2]; U.cy = K[5];
    U.k1 = dist ? dist[0] : 0.0; U.k2 = dist ? dist[1] : 0.0; U.p1 = dist ? dist[2] : 0.0; U.p2 = dist ? dist[3] : 0.0;
    U.k3 = dist ? dist[4] : 0.0;
    // all-zero coefficients: the fixed-point loop is the identity (icdist = 1, delta = 0), skip it
    U.has_dist = (U.k1 != 0.0 || U.k2 != 0.0 || U.p1 != 0.0 || U.p2 != 0.0 || U.k3 != 0.0) ? 1 : 0;
    U.tangential = (U.p1 != 0.0 || U.p2 != 0.0) ? 1 : 0;
    return true;
}
bool make_undist2(const double* K1, const double* d1, const double* K2, const double* d2, Undist2& p) {
    return make_undist(K1, d1, p.cam[0]) && make_undist(K2, d2, p.cam[1]);
}

// Right and left epipoles of a rank-2 F as the largest cross product of two rows / columns.
void null_vector3(const double M[9], bool transpose, double e[3]) {
    double r[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r[i][j] = transpose ? M[3 * j + i] : M[3 * i + j];
    double best = -1.0;
    e[0] = e[1] = e[2] = 0.0;
    for (int i = 0; i < 3; ++i)
        for (int j = i + 1; j < 3; ++j) {
            const double c[3] = {r[i][1] * r[j][2] - r[i][2] * r[j][1], r[i][2] * r[j][0] - r[i][0] * r[j][2],
                                 r[i][0] * r[j][1] - r[i][1] * r[j][0]};
            const double nn = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
            if (nn > best) { best = nn; e[0] = c[0]; e[1] = c[1]; e[2] = c[2]; }
        }
    const double nrm = std::sqrt(best);
    if (nrm > 0) { e[0] /= nrm; e[1] /= nrm; e[2] /= nrm; }
}

HSParams make_hs(const double* F) {
    HSParams hs;
    for (int i = 0; i < 9; ++i) hs.F[i] = F[i];
    null_vector3(F, false, hs.e1);
    null_vector3(F, true, hs.e2);
    return hs;
}

// F = [t]x R of P_canon = P2_full * inv(P1_full)   (triangulation.py:211-216)
bool fundamental_from_P(const double* P1, const double* P2, double F[9]) {
    // inverse of [A1 | t1; 0 0 0 1]: [A1^-1 | -A1^-1 t1]
    const double* A = P1;
    const double a00 = A[0], a01 = A[1], a02 = A[2], a10 = A[4], a11 = A[5], a12 = A[6], a20 = A[8], a21 = A[9], a22 = A[10];
    const double c00 = a11 * a22 - a12 * a21, c01 = a12 * a20 - a10 * a22, c02 = a10 * a21 - a11 * a20;
    const double det = a00 * c00 + a01 * c01 + a02 * c02;
    double inv[3][3];
    inv[0][0] = c00 / det; inv[0][1] = (a02 * a21 - a01 * a22) / det; inv[0][2] = (a01 * a12 - a02 * a11) / det;
    inv[1][0] = c01 / det; inv[1][1] = (a00 * a22 - a02 * a20) / det; inv[1][2] = (a02 * a10 - a00 * a12) / det;
    inv[2][0] = c02 / det; inv[2][1] = (a01 * a20 - a00 * a21) / det; inv[2][2] = (a00 * a11 - a01 * a10) / det;
    const double t1[3] = {P1[3], P1[7], P1[11]};
    double it[3];
    for (int i = 0; i < 3; ++i) it[i] = -(inv[i][0] * t1[0] + inv[i][1] * t1[1] + inv[i][2] * t1[2]);
    double R[3][3], t[3];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j)
            R[i][j] = P2[4 * i + 0] * inv[0][j] + P2[4 * i + 1] * inv[1][j] + P2[4 * i + 2] * inv[2][j];
        t[i] = P2[4 * i + 0] * it[0] + P2[4 * i + 1] * it[1] + P2[4 * i + 2] * it[2] + P2[4 * i + 3];
    }
    // F[:, j] = t x R[:, j]
    for (int j = 0; j < 3; ++j) {
        F[0 + j] = t[1] * R[2][j] - t[2] * R[1][j];
        F[3 + j] = t[2] * R[0][j] - t[0] * R[2][j];
        F[6 + j] = t[0] * R[1][j] - t[1] * R[0][j];
    }
    return true;
}

// Host one-sided Jacobi SVD of a small dense matrix (row-major m x n, n <= 9); V (n x n) gets the right singular
// vectors, w the singular values (unsorted).  Only used by the rare 8-point fallback.
void host_jacobi_svd(double* A, int m, int n, double* V, double* w) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) V[i * n + j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        bool changed = false;
        for (int i = 0; i < n - 1; ++i)
            for (int j = i + 1; j < n; ++j) {
                double a = 0, b = 0, p = 0;
                for (int k = 0; k < m; ++k) { a += A[k * n + i] * A[k * n + i]; b += A[k * n + j] * A[k * n + j]; p += A[k * n + i] * A[k * n + j]; }
                if (!(std::fabs(p) > 2.220446049250313e-16 * std::sqrt(a * b))) continue;
                changed = true;
                p *= 2;
                const double beta = a - b, gamma = std::hypot(p, beta);
                double c, s;
                if (beta < 0) { const double delta = (gamma - beta) * 0.5; s = std::sqrt(delta / gamma); c = p / (gamma * s * 2); }
                else { c = std::sqrt((gamma + beta) / (gamma * 2)); s = p / (gamma * c * 2); }
                for (int k = 0; k < m; ++k) {
                    const double t0 = c * A[k * n + i] + s * A[k * n + j], t1 = -s * A[k * n + i] + c * A[k * n + j];
                    A[k * n + i] = t0; A[k * n + j] = t1;
                }
                for (int k = 0; k < n; ++k) {
                    const double t0 = c * V[k * n + i] + s * V[k * n + j], t1 = -s * V[k * n + i] + c * V[k * n + j];
                    V[k * n + i] = t0; V[k * n + j] = t1;
                }
            }
        if (!changed) break;
    }
    for (int j = 0; j < n; ++j) {
        double s = 0;
        for (int k = 0; k < m; ++k) s += A[k * n + j] * A[k * n + j];
        w[j] = std::sqrt(s);
    }
}

}  // namespace

// ================================================================================================================
extern "C" {

int trgl_version(void) { return TRGL_VERSION; }
const char* trgl_last_error_string(void) { return g_err.c_str(); }
int trgl_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
int trgl_set_device(int device) { CK(cudaSetDevice(device)); return TRGL_OK; }
int trgl_device_synchronize(void) { CK(cudaDeviceSynchronize()); return TRGL_OK; }

int trgl_device_alloc(void** ptr, size_t bytes) {
    if (!ptr) return fail(TRGL_E_BADARG, "ptr is NULL");
    *ptr = nullptr;
    if (bytes == 0) return TRGL_OK;
    CK(cudaMalloc(ptr, bytes));
    return TRGL_OK;
}
int trgl_device_free(void* ptr) { if (ptr) CK(cudaFree(ptr)); return TRGL_OK; }
int trgl_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) return fail(TRGL_E_BADARG, "ptr is NULL");
    *ptr = nullptr;
    if (bytes == 0) return TRGL_OK;
    CK(cudaHostAlloc(ptr, bytes, cudaHostAllocDefault));
    return TRGL_OK;
}
int trgl_host_free(void* ptr) { if (ptr) CK(cudaFreeHost(ptr)); return TRGL_OK; }
int trgl_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream) {
    if (bytes) CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
    return TRGL_OK;
}
int trgl_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream) {
    if (bytes) CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
    return TRGL_OK;
}
int trgl_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream) {
    if (bytes) CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    return TRGL_OK;
}
int trgl_memset_d(void* dst, int value, size_t bytes, void* stream) {
    if (bytes) CK(cudaMemsetAsync(dst, value, bytes, static_cast<cudaStream_t>(stream)));
    return TRGL_OK;
}
int trgl_stream_create(void** stream) {
    if (!stream) return fail(TRGL_E_BADARG, "stream is NULL");
    cudaStream_t s;
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = s;
    return TRGL_OK;
}
int trgl_stream_destroy(void* stream) { if (stream) CK(cudaStreamDestroy(static_cast<cudaStream_t>(stream))); return TRGL_OK; }
int trgl_stream_synchronize(void* stream) { CK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream))); return TRGL_OK; }
int trgl_event_create(void** event) {
    if (!event) return fail(TRGL_E_BADARG, "event is NULL");
    cudaEvent_t e;
    CK(cudaEventCreate(&e));
    *event = e;
    return TRGL_OK;
}
int trgl_event_destroy(void* event) { if (event) CK(cudaEventDestroy(static_cast<cudaEvent_t>(event))); return TRGL_OK; }
int trgl_event_record(void* event, void* stream) {
    CK(cudaEventRecord(static_cast<cudaEvent_t>(event), static_cast<cudaStream_t>(stream)));
    return TRGL_OK;
}
int trgl_event_synchronize(void* event) {
    CK(cudaEventSynchronize(static_cast<cudaEvent_t>(event)));
    return TRGL_OK;
}
int trgl_event_elapsed_ms(void* start, void* stop, float* ms) {
    CK(cudaEventSynchronize(static_cast<cudaEvent_t>(stop)));
    CK(cudaEventElapsedTime(ms, static_cast<cudaEvent_t>(start), static_cast<cudaEvent_t>(stop)));
    return TRGL_OK;
}

// ---- result mirrors / CUDA IPC (multi-GPU gather fused into the solver stores) ----------------------------------------
static int set_mirrors(void* const* x_mirrors, void* const* status_mirrors, int count, int x_f32) {
    if (count < 0 || count > kMaxMirrors) return fail(TRGL_E_BADARG, "at most 8 result mirrors");
    if (count > 0 && (!x_mirrors || !status_mirrors)) return fail(TRGL_E_BADARG, "NULL mirror table");
    g_next_mirrors.count = count;
    g_next_mirrors.x_f32 = x_f32 ? 1 : 0;
    for (int r = 0; r < count; ++r) {
        if (!x_mirrors[r] || !status_mirrors[r]) { g_next_mirrors.count = 0; return fail(TRGL_E_BADARG, "NULL mirror pointer"); }
        g_next_mirrors.x[r] = x_mirrors[r]; g_next_mirrors.status[r] = status_mirrors[r];
    }
    return TRGL_OK;
}
int trgl_set_result_mirrors(void* const* x_mirrors, void* const* status_mirrors, int count) {
    return set_mirrors(x_mirrors, status_mirrors, count, 0);
}
int trgl_set_result_mirrors_f32(void* const* x_mirrors, void* const* status_mirrors, int count) {
    return set_mirrors(x_mirrors, status_mirrors, count, 1);
}
int trgl_set_fused_eval(int min_status, double max_sq_err, void* err1, void* err2, uint8_t* good, double* sums_device) {
    if (!sums_device) return fail(TRGL_E_BADARG, "sums_device is NULL");
    g_next_eval.armed = true;
    g_next_eval.min_status = min_status; g_next_eval.max_sq_err = max_sq_err;
    g_next_eval.err1 = err1; g_next_eval.err2 = err2; g_next_eval.good = good; g_next_eval.sums = sums_device;
    return TRGL_OK;
}
int trgl_set_input_retention(void* u1_device, void* u2_device) {
    if (!u1_device || !u2_device) { g_next_retain = {nullptr, nullptr}; return fail(TRGL_E_BADARG, "NULL retention buffer"); }
    g_next_retain = {u1_device, u2_device};
    return TRGL_OK;
}
int trgl_ipc_export(void* device_ptr, void* handle64) {
    if (!device_ptr || !handle64) return fail(TRGL_E_BADARG, "NULL pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, device_ptr));
    std::memcpy(handle64, &h, sizeof(h));
    return TRGL_OK;
}
int trgl_ipc_import(const void* handle64, void** device_ptr) {
    if (!device_ptr || !handle64) return fail(TRGL_E_BADARG, "NULL pointer");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, sizeof(h));
    CK(cudaIpcOpenMemHandle(device_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return TRGL_OK;
}
int trgl_ipc_close(void* device_ptr) {
    if (device_ptr) CK(cudaIpcCloseMemHandle(device_ptr));
    return TRGL_OK;
}

int64_t trgl_launch_count(void) { return g_launches.load(); }
int trgl_set_stream_variant(int variant) {
    const int old = g_variant.load();
    if (variant >= -1 && variant <= 12) g_variant.store(variant);
    return old;
}
int64_t trgl_set_deferred_capacity(int64_t max_points) {
    const int64_t old = g_deferred_limit.load();
    if (max_points >= 1 && max_points <= static_cast<int64_t>(kDeferredMax)) g_deferred_limit.store(static_cast<unsigned int>(max_points));
    return old;
}
int trgl_set_trace(int enabled) {
    const int old = g_trace.exchange(enabled ? 1 : 0);
    return old;
}
int trgl_get_trace(double* out5) {
    if (!out5) return fail(TRGL_E_BADARG, "out5 is NULL");
    std::lock_guard<std::mutex> lock(g_pipe_mutex);
    for (int k = 0; k < 5; ++k) { out5[k] = g_trace_us[k]; g_trace_us[k] = 0.0; }
    return TRGL_OK;
}
int trgl_rare_path_counters(unsigned long long* out5, int reset) {
    unsigned long long* out4 = out5;
    if (!out4) return fail(TRGL_E_BADARG, "out5 is NULL");
    if (!have_device()) return fail(TRGL_E_NODEVICE, "no CUDA device available (libtriangl_cuda has no CPU fallback)");
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpyFromSymbol(out4, g_hs_counters, 5 * sizeof(unsigned long long)));
    if (reset) {
        const unsigned long long zero[5] = {0, 0, 0, 0, 0};
        CK(cudaMemcpyToSymbol(g_hs_counters, zero, sizeof(zero)));
    }
    return TRGL_OK;
}
int trgl_fp64_fma_rate(int operands, int chains, int ctas_per_sm, double* warp_fma_per_second) {
    if (!warp_fma_per_second) return fail(TRGL_E_BADARG, "warp_fma_per_second is NULL");
    *warp_fma_per_second = 0.0;
    if ((operands != 2 && operands != 3) || (chains != 1 && chains != 2 && chains != 8) || ctas_per_sm < 1 || ctas_per_sm > 8)
        return fail(TRGL_E_BADARG, "operands must be 2 or 3, chains 1, 2 or 8, ctas_per_sm 1..8");
    if (!have_device()) return fail(TRGL_E_NODEVICE, "no CUDA device available (libtriangl_cuda has no CPU fallback)");
    const int grid = sm_count() * ctas_per_sm, iters = 4000;
    double* out = nullptr;
    CK(cudaMalloc(reinterpret_cast<void**>(&out), sizeof(double) * size_t(grid) * 256));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](int it) {
        const double a = 1.0000001;
#define TRGL_PROBE(O, C) k_fp64_fma_rate<O, C><<<grid, 256>>>(out, it, a)
        if (operands == 2) { if (chains == 1) TRGL_PROBE(2, 1); else if (chains == 2) TRGL_PROBE(2, 2); else TRGL_PROBE(2, 8); }
        else { if (chains == 1) TRGL_PROBE(3, 1); else if (chains == 2) TRGL_PROBE(3, 2); else TRGL_PROBE(3, 8); }
#undef TRGL_PROBE
    };
    run(100);                                   // warm-up (clocks, instruction cache)
    float ms_short = 0.f, ms_long = 0.f;
    cudaEventRecord(e0); run(iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms_short, e0, e1);
    cudaEventRecord(e0); run(3 * iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms_long, e0, e1);
    g_launches += 3;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    CK(cudaGetLastError());
    // the difference of the two runs removes launch overhead, ramp and tail
    const double warp_fma = 2.0 * iters * 8.0 * chains * 8.0 * grid;          // 8 warps per CTA
    const double sec = (double(ms_long) - double(ms_short)) * 1e-3;
    if (!(sec > 0)) return fail(TRGL_E_BADARG, "fp64 probe: non-positive time difference");
    *warp_fma_per_second = warp_fma / sec;
    return TRGL_OK;
}
int trgl_deferred_total(void* stream, int64_t* total) {
    if (!total) return fail(TRGL_E_BADARG, "total is NULL");
    *total = 0;
    if (!have_device()) return fail(TRGL_E_NODEVICE, "no CUDA device available (libtriangl_cuda has no CPU fallback)");
    Scratch sc;
    int rc = scratch_for(static_cast<cudaStream_t>(stream), sc);
    if (rc) return rc;
    unsigned long long v = 0;
    CK(cudaMemcpyAsync(&v, sc.deferred_ctl + 2, sizeof(v), cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
    CK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    *total = static_cast<int64_t>(v);
    return TRGL_OK;
}
int trgl_set_two_ray(int enabled) {
    const int old = g_two_ray.load();
    g_two_ray.store(enabled ? 1 : 0);
    return old;
}
int trgl_set_points_per_thread(int ppt) {
    const int old = g_ppt.load();
    if (ppt == 1 || ppt == 2 || ppt == 4) g_ppt.store(ppt);
    return old;
}

// ---------------------------------------------------------------------------------------------------------------
static int impl_linear_ls(const void* u1, const void* u2, const double* P1, const double* P2, void* x, uint8_t* status,
                          int64_t n, int mode, int mem, void* stream, const Undist2* pre) {
    int rc = check_common(u1, u2, P1, P2, x, status, n, mode, mem, stream);
    if (rc || n == 0) return rc;
    if (mem == TRGL_MEM_DEVICE) {
        FusedEval fe; const FusedEval* evp;
        rc = take_eval(P1, P2, static_cast<cudaStream_t>(stream), fe, evp);
        if (rc) return rc;
        return launch_linear_ls(u1, u2, P1, P2, x, status, n, mode, static_cast<cudaStream_t>(stream), pre, take_mirrors(), evp);
    }
    ModeInfo mi; mode_info(mode, mi);
    HostArray arr[4] = {input_array(u1, size_t(2 * mi.in_bytes), mem, g_next_retain.u1),
                        input_array(u2, size_t(2 * mi.in_bytes), mem, g_next_retain.u2),
                        {nullptr, x, size_t(3 * mi.out_bytes)}, {nullptr, status, 1}};
    return host_pipeline(arr, 4, n, [&](void** d, int64_t m, cudaStream_t s, int) {
        return launch_linear_ls(d[0], d[1], P1, P2, d[2], static_cast<uint8_t*>(d[3]), m, mode, s, pre);
    });
}
int trgl_linear_ls(const void* u1, const void* u2, const double* P1, const double* P2, void* x, uint8_t* status,
                   int64_t n, int mode, int mem, void* stream) {
    PendingClear pending_clear;
    return impl_linear_ls(u1, u2, P1, P2, x, status, n, mode, mem, stream, nullptr);
}
int trgl_linear_ls_px(const void* px1, const void* px2, const double* K1, const double* dist1, const double* K2,
                      const double* dist2, const double* P1, const double* P2, void* x, uint8_t* status, int64_t n,
                      int mode, int mem, void* stream) {
    PendingClear pending_clear;
    Undist2 pre;
    if (!make_undist2(K1, dist1, K2, dist2, pre)) return fail(TRGL_E_BADARG, "camera matrix K is NULL or has a zero focal length");
    return impl_linear_ls(px1, px2, P1, P2, x, status, n, mode, mem, stream, &pre);
}

// ---- multi-view linear LS -----------------------------------------------------------------------
static int launch_multiview_ls(void* const* u, void* const* valid, const double* P, int m, int min_views, void* x,
                               uint8_t* status, int64_t n, int mode, cudaStream_t s) {
    if (n == 0) return TRGL_OK;
    if (mode == TRGL_F32) mode = TRGL_F32IO;     // float32 storage, float64 registers
    Scratch sc;
    int rc = scratch_for(s, sc, n);
    if (rc) return rc;
    const Deferred df = make_deferred(sc, kHintMultiview);
    MODE_SWITCH(mode, {
        if constexpr (sizeof(TC) == 8) {
            MultiViewArgs<TI, TC> args;
            std::memset(&args, 0, sizeof(args));
            for (int v = 0; v < m; ++v) {
                args.u[v] = static_cast<const TI*>(u[v]);
                args.valid[v] = valid ? static_cast<const uint8_t*>(valid[v]) : nullptr;
                for (int k = 0; k < 12; ++k) args.P[v][k] = P[12 * v + k];
            }
            args.m = m; args.min_views = min_views;
            const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
            auto go = [&](auto kern, int per_block) {
                const int64_t tiles = (n + per_block - 1) / per_block;
                kern<<<static_cast<unsigned>(tiles < cap ? tiles : cap), kThreads, 0, s>>>(args, static_cast<TO*>(x), status, n, df);
            };
            const bool masked = valid != nullptr;
            if (m <= 4) {
                if (masked) go(k_multiview_ls<TI, TC, TO, 2, 4, 1, true>, 2 * kThreads);
                else go(k_multiview_ls<TI, TC, TO, 2, 4, 1, false>, 2 * kThreads);
            } else if (m <= 8) {
                if (masked) go(k_multiview_ls<TI, TC, TO, 1, 8, 1, true>, kThreads);
                else go(k_multiview_ls<TI, TC, TO, 1, 8, 1, false>, kThreads);
            } else {
                if (masked) go(k_multiview_ls<TI, TC, TO, 1, 8, 2, true>, kThreads);
                else go(k_multiview_ls<TI, TC, TO, 1, 8, 2, false>, kThreads);
            }
            launch_followup(k_multiview_general<TI, TC, TO>, followup_grid(sc, kHintMultiview, (n + kThreads - 1) / kThreads), s,
                            args, static_cast<TO*>(x), n, df);
        } else {
            rc = fail(TRGL_E_BADARG, "multi-view triangulation computes in float64");
        }
    })
    if (rc) return rc;
    g_launches += 2;
    CK(cudaGetLastError());
    return TRGL_OK;
}

int trgl_multiview_ls(const void* u, const uint8_t* valid, const double* P, int m, void* x, uint8_t* status, int64_t n,
                      int min_views, int mode, int mem, void* stream) {
    PendingClear pending_clear;       // a mirror / evaluation request armed before this call does not leak into the next one
    ModeInfo mi;
    if (n < 0) return fail(TRGL_E_BADARG, "negative point count");
    if (!mode_info(mode, mi)) return fail(TRGL_E_BADARG, "unknown precision mode");
    if (mem != TRGL_MEM_HOST && mem != TRGL_MEM_DEVICE) return fail(TRGL_E_BADARG, "unknown memory space");
    if (m < 1 || m > kMaxViews) return fail(TRGL_E_BADARG, "number of views must be 1..16");
    if (!P) return fail(TRGL_E_BADARG, "camera matrix pointer is NULL");
    if (n > 0 && (!u || !x || !status)) return fail(TRGL_E_BADARG, "NULL array pointer with n > 0");
    if (mem == TRGL_MEM_DEVICE && n > 0 && (reinterpret_cast<uintptr_t>(u) & static_cast<uintptr_t>(2 * mi.in_bytes - 1)))
        return fail(TRGL_E_BADARG, "device u must be aligned to one (x,y) pair (16 bytes float64, 8 bytes float32)");
    if (!have_device()) return fail(TRGL_E_NODEVICE, "no CUDA device available (libtriangl_cuda has no CPU fallback)");
    if (n == 0) return TRGL_OK;
    const size_t view_bytes = static_cast<size_t>(n) * 2 * mi.in_bytes;
    if (mem == TRGL_MEM_DEVICE) {
        void* up[kMaxViews]; void* vp[kMaxViews];
        for (int v = 0; v < m; ++v) {
            up[v] = const_cast<char*>(static_cast<const char*>(u)) + v * view_bytes;
            vp[v] = valid ? const_cast<uint8_t*>(valid) + static_cast<size_t>(v) * n : nullptr;
        }
        return launch_multiview_ls(up, valid ? vp : nullptr, P, m, min_views, x, status, n, mode, static_cast<cudaStream_t>(stream));
    }
    HostArray arr[kMaxHostArrays];
    int na = 0;
    for (int v = 0; v < m; ++v) arr[na++] = {static_cast<const char*>(u) + v * view_bytes, nullptr, size_t(2 * mi.in_bytes)};
    if (valid) for (int v = 0; v < m; ++v) arr[na++] = {valid + static_cast<size_t>(v) * n, nullptr, 1};
    const int ix = na;
    arr[na++] = {nullptr, x, size_t(3 * mi.out_bytes)};
    arr[na++] = {nullptr, status, 1};
    return host_pipeline(arr, na, n, [&](void** d, int64_t cnt, cudaStream_t s, int) {
        return launch_multiview_ls(d, valid ? d + m : nullptr, P, m, min_views, d[ix], static_cast<uint8_t*>(d[ix + 1]), cnt, mode, s);
    });
}

static int impl_iterative_ls(const void* u1, const void* u2, const double* P1, const double* P2, void* x, int32_t* status,
                             int64_t n, double tolerance, int semantics, int mode, int mem, void* stream, const Undist2* pre) {
    int rc = check_common(u1, u2, P1, P2, x, status, n, mode, mem, stream);
    if (rc) return rc;
    if (semantics != TRGL_ITER_C && semantics != TRGL_ITER_PY) return fail(TRGL_E_BADARG, "unknown iterative semantics");
    if (n == 0) return TRGL_OK;
    if (mem == TRGL_MEM_DEVICE) {
        FusedEval fe; const FusedEval* evp;
        rc = take_eval(P1, P2, static_cast<cudaStream_t>(stream), fe, evp);
        if (rc) return rc;
        return launch_iterative_ls(u1, u2, P1, P2, x, status, n, tolerance, semantics, mode, static_cast<cudaStream_t>(stream), pre,
                                   take_mirrors(), evp);
    }
    ModeInfo mi; mode_info(mode, mi);
    HostArray arr[4] = {input_array(u1, size_t(2 * mi.in_bytes), mem, g_next_retain.u1),
                        input_array(u2, size_t(2 * mi.in_bytes), mem, g_next_retain.u2),
                        {nullptr, x, size_t(3 * mi.out_bytes)}, {nullptr, status, 4}};
    return host_pipeline(arr, 4, n, [&](void** d, int64_t m, cudaStream_t s, int) {
        return launch_iterative_ls(d[0], d[1], P1, P2, d[2], static_cast<int32_t*>(d[3]), m, tolerance, semantics, mode, s, pre);
    });
}
int trgl_iterative_ls(const void* u1, const void* u2, const double* P1, const double* P2, void* x, int32_t* status,
                      int64_t n, double tolerance, int semantics, int mode, int mem, void* stream) {
    PendingClear pending_clear;
    return impl_iterative_ls(u1, u2, P1, P2, x, status, n, tolerance, semantics, mode, mem, stream, nullptr);
}
int trgl_iterative_ls_px(const void* px1, const void* px2, const double* K1, const double* dist1, const double* K2,
                         const double* dist2, const double* P1, const double* P2, void* x, int32_t* status, int64_t n,
                         double tolerance, int semantics, int mode, int mem, void* stream) {
    PendingClear pending_clear;
    Undist2 pre;
    if (!make_undist2(K1, dist1, K2, dist2, pre)) return fail(TRGL_E_BADARG, "camera matrix K is NULL or has a zero focal length");
    return impl_iterative_ls(px1, px2, P1, P2, x, status, n, tolerance, semantics, mode, mem, stream, &pre);
}

static int impl_linear_eigen(const void* u1, const void* u2, const double* P1, const double* P2, void* x, uint8_t* status,
                             int64_t n, double max_coordinate_value, int rows, int mode, int mem, void* stream,
                             const Undist2* pre) {
    int rc = check_common(u1, u2, P1, P2, x, status, n, mode, mem, stream);
    if (rc) return rc;
    if (rows != 4 && rows != 6) return fail(TRGL_E_BADARG, "rows must be 4 or 6");
    if (n == 0) return TRGL_OK;
    if (mem == TRGL_MEM_DEVICE) {
        FusedEval fe; const FusedEval* evp;
        rc = take_eval(P1, P2, static_cast<cudaStream_t>(stream), fe, evp);
        if (rc) return rc;
        return launch_linear_eigen(u1, u2, P1, P2, x, status, n, max_coordinate_value, rows, mode, static_cast<cudaStream_t>(stream), pre,
                                   take_mirrors(), evp);
    }
    ModeInfo mi; mode_info(mode, mi);
    HostArray arr[4] = {input_array(u1, size_t(2 * mi.in_bytes), mem, g_next_retain.u1),
                        input_array(u2, size_t(2 * mi.in_bytes), mem, g_next_retain.u2),
                        {nullptr, x, size_t(3 * mi.out_bytes)}, {nullptr, status, 1}};
    return host_pipeline(arr, 4, n, [&](void** d, int64_t m, cudaStream_t s, int) {
        return launch_linear_eigen(d[0], d[1], P1, P2, d[2], static_cast<uint8_t*>(d[3]), m, max_coordinate_value, rows, mode, s, pre);
    });
}
int trgl_linear_eigen(const void* u1, const void* u2, const double* P1, const double* P2, void* x, uint8_t* status,
                      int64_t n, double max_coordinate_value, int rows, int mode, int mem, void* stream) {
    PendingClear pending_clear;
    return impl_linear_eigen(u1, u2, P1, P2, x, status, n, max_coordinate_value, rows, mode, mem, stream, nullptr);
}
int trgl_linear_eigen_px(const void* px1, const void* px2, const double* K1, const double* dist1, const double* K2,
                         const double* dist2, const double* P1, const double* P2, void* x, uint8_t* status, int64_t n,
                         double max_coordinate_value, int rows, int mode, int mem, void* stream) {
    PendingClear pending_clear;
    Undist2 pre;
    if (!make_undist2(K1, dist1, K2, dist2, pre)) return fail(TRGL_E_BADARG, "camera matrix K is NULL or has a zero focal length");
    return impl_linear_eigen(px1, px2, P1, P2, x, status, n, max_coordinate_value, rows, mode, mem, stream, &pre);
}

static int impl_polynomial_F(const void* u1, const void* u2, const double* P1, const double* P2, const double* F, void* x,
                             uint8_t* status, void* u1_corr, void* u2_corr, int* all_nan, int64_t n,
                             double max_coordinate_value, int rows, int mode, int mem, void* stream, const Undist2* pre) {
    int rc = check_common(u1, u2, P1, P2, x, status, n, mode, mem, stream);
    if (rc) return rc;
    if (!F) return fail(TRGL_E_BADARG, "F is NULL");
    if (rows != 4 && rows != 6) return fail(TRGL_E_BADARG, "rows must be 4 or 6");
    if (all_nan) *all_nan = 0;
    if (n == 0) return TRGL_OK;
    const HSParams hs = make_hs(F);
    ModeInfo mi; mode_info(mode, mi);
    unsigned int hflags[2 * (kSlots + 1)] = {0};
    if (mem == TRGL_MEM_DEVICE) {
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        Scratch sc;
        rc = scratch_for(s, sc);
        if (rc) return rc;
        unsigned int* fl = sc.flags + 2 * kSlots;
        CK(cudaMemsetAsync(fl, 0, 2 * sizeof(unsigned int), s));
        FusedEval fe; const FusedEval* evp;
        rc = take_eval(P1, P2, s, fe, evp);
        if (rc) return rc;
        rc = launch_polynomial(u1, u2, P1, P2, hs, x, status, u1_corr, u2_corr, fl, n, max_coordinate_value, rows, mode, s, pre,
                               take_mirrors(), evp);
        if (rc) return rc;
        if (all_nan) {
            CK(cudaMemcpyAsync(hflags, fl, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            *all_nan = (hflags[0] == 0 || hflags[1] == 0) ? 1 : 0;
        }
        return TRGL_OK;
    }
    std::lock_guard<std::mutex> host_lock(g_host_mutex);          // the slot flags live across host_pipeline
    Scratch sc;
    rc = scratch_for(nullptr, sc);
    if (rc) return rc;
    // the pipeline slots run on non-blocking streams, which are not ordered behind the legacy default stream: the
    // clear must have completed before any slot's kernel can set a flag
    CK(cudaMemset(sc.flags, 0, sizeof(unsigned int) * 2 * kSlots));
    CK(cudaStreamSynchronize(nullptr));
    HostArray arr[6] = {input_array(u1, size_t(2 * mi.in_bytes), mem, g_next_retain.u1),
                        input_array(u2, size_t(2 * mi.in_bytes), mem, g_next_retain.u2),
                        {nullptr, x, size_t(3 * mi.out_bytes)}, {nullptr, status, 1},
                        {nullptr, u1_corr, size_t(2 * mi.in_bytes)}, {nullptr, u2_corr, size_t(2 * mi.in_bytes)}};
    rc = host_pipeline(arr, 6, n, [&](void** d, int64_t m, cudaStream_t s, int slot) {
        return launch_polynomial(d[0], d[1], P1, P2, hs, d[2], static_cast<uint8_t*>(d[3]), u1_corr ? d[4] : nullptr,
                                 u2_corr ? d[5] : nullptr, sc.flags + 2 * slot, m, max_coordinate_value, rows, mode, s, pre);
    });
    if (rc) return rc;
    if (all_nan) {
        CK(cudaMemcpy(hflags, sc.flags, sizeof(unsigned int) * 2 * kSlots, cudaMemcpyDeviceToHost));
        unsigned a = 0, b = 0;
        for (int s = 0; s < kSlots; ++s) { a |= hflags[2 * s]; b |= hflags[2 * s + 1]; }
        *all_nan = (a == 0 || b == 0) ? 1 : 0;
    }
    return TRGL_OK;
}

int trgl_polynomial_F(const void* u1, const void* u2, const double* P1, const double* P2, const double* F, void* x,
                      uint8_t* status, void* u1_corr, void* u2_corr, int* all_nan, int64_t n,
                      double max_coordinate_value, int rows, int mode, int mem, void* stream) {
    PendingClear pending_clear;
    return impl_polynomial_F(u1, u2, P1, P2, F, x, status, u1_corr, u2_corr, all_nan, n, max_coordinate_value, rows, mode,
                             mem, stream, nullptr);
}

int trgl_polynomial_flags_async(unsigned int* host_flags2, void* stream) {
    if (!host_flags2) return fail(TRGL_E_BADARG, "host_flags2 is NULL");
    if (!have_device()) return fail(TRGL_E_NODEVICE, "no CUDA device available (libtriangl_cuda has no CPU fallback)");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    Scratch sc;
    int rc = scratch_for(s, sc);
    if (rc) return rc;
    CK(cudaMemcpyAsync(host_flags2, sc.flags + 2 * kSlots, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
    return TRGL_OK;
}

int trgl_polynomial(const void* u1, const void* u2, const double* P1, const double* P2, void* x, uint8_t* status,
                    void* u1_corr, void* u2_corr, int* all_nan, int64_t n, double max_coordinate_value, int rows,
                    int mode, int mem, void* stream) {
    PendingClear pending_clear;
    if (!P1 || !P2) return fail(TRGL_E_BADARG, "camera matrix pointer is NULL");
    double F[9];
    fundamental_from_P(P1, P2, F);
    return trgl_polynomial_F(u1, u2, P1, P2, F, x, status, u1_corr, u2_corr, all_nan, n, max_coordinate_value, rows,
                             mode, mem, stream);
}

int trgl_polynomial_px(const void* px1, const void* px2, const double* K1, const double* dist1, const double* K2,
                       const double* dist2, const double* P1, const double* P2, void* x, uint8_t* status,
                       void* u1_corr, void* u2_corr, int* all_nan, int64_t n, double max_coordinate_value, int rows,
                       int mode, int mem, void* stream) {
    PendingClear pending_clear;
    if (!P1 || !P2) return fail(TRGL_E_BADARG, "camera matrix pointer is NULL");
    Undist2 pre;
    if (!make_undist2(K1, dist1, K2, dist2, pre)) return fail(TRGL_E_BADARG, "camera matrix K is NULL or has a zero focal length");
    double F[9];
    fundamental_from_P(P1, P2, F);
    return impl_polynomial_F(px1, px2, P1, P2, F, x, status, u1_corr, u2_corr, all_nan, n, max_coordinate_value, rows,
                             mode, mem, stream, &pre);
}

// cv2.undistortPoints(src, K, dist) -> normalised coordinates in the dtype of src (slam2.py:551-552)
int trgl_undistort_points(const void* src, void* dst, const double* K, const double* dist, int64_t n, int in_is_f32,
                          int mem, void* stream) {
    if (n < 0) return fail(TRGL_E_BADARG, "negative point count");
    if (mem != TRGL_MEM_HOST && mem != TRGL_MEM_DEVICE) return fail(TRGL_E_BADARG, "unknown memory space");
    if (n > 0 && (!src || !dst)) return fail(TRGL_E_BADARG, "NULL array pointer with n > 0");
    Undist U;
    if (!make_undist(K, dist, U)) return fail(TRGL_E_BADARG, "camera matrix K is NULL or has a zero focal length");
    if (!have_device()) return fail(TRGL_E_NODEVICE, "no CUDA device available (libtriangl_cuda has no CPU fallback)");
    if (n == 0) return TRGL_OK;
    const size_t pb = in_is_f32 ? 8 : 16;
    auto run = [&](const void* s_, void* d_, int64_t m, cudaStream_t st) -> int {
        const int64_t tiles = (m + kThreads - 1) / kThreads;
        const int64_t gcap = static_cast<int64_t>(sm_count()) * 16;
        const unsigned grid = static_cast<unsigned>(tiles < gcap ? tiles : gcap);
        if (in_is_f32) k_undistort_points<float><<<grid, kThreads, 0, st>>>(static_cast<const float*>(s_), static_cast<float*>(d_), U, m);
        else k_undistort_points<double><<<grid, kThreads, 0, st>>>(static_cast<const double*>(s_), static_cast<double*>(d_), U, m);
        g_launches++;
        CK(cudaGetLastError());
        return TRGL_OK;
    };
    if (mem == TRGL_MEM_DEVICE) {
        if ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & (pb - 1))
            return fail(TRGL_E_BADARG, "device src/dst must be aligned to one (x,y) pair");
        return run(src, dst, n, static_cast<cudaStream_t>(stream));
    }
    HostArray arr[2] = {{src, nullptr, pb}, {nullptr, dst, pb}};
    return host_pipeline(arr, 2, n, [&](void** d, int64_t m, cudaStream_t st, int) { return run(d[0], d[1], m, st); });
}

// ---------------------------------------------------------------------------------------------------------------
int trgl_fundamental_8point(const void* u1, const void* u2, int64_t n, int mode, int mem, double* F, void* stream) {
    ModeInfo mi;
    if (!mode_info(mode, mi)) return fail(TRGL_E_BADARG, "unknown precision mode");
    if (mem != TRGL_MEM_HOST && mem != TRGL_MEM_DEVICE) return fail(TRGL_E_BADARG, "unknown memory space");
    if (!F || n < 8 || !u1 || !u2) return fail(TRGL_E_BADARG, "need F and at least 8 matches");
    if (!have_device()) return fail(TRGL_E_NODEVICE, "no CUDA device available (libtriangl_cuda has no CPU fallback)");
    std::unique_lock<std::mutex> lock(g_pipe_mutex, std::defer_lock);
    if (mem == TRGL_MEM_HOST) lock.lock();                  // the pipeline slots are shared by host-mode calls
    int rc;
    const size_t ub = 2 * mi.in_bytes;
    const int64_t chunk = (mem == TRGL_MEM_HOST) ? (n < kChunk ? n : kChunk) : n;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    char* d1 = nullptr; char* d2 = nullptr;
    if (mem == TRGL_MEM_HOST) {
        rc = ensure_slot(g_slots[0], 2 * align256(ub * chunk));
        if (rc) return rc;
        s = g_slots[0].stream; d1 = g_slots[0].buf; d2 = d1 + align256(ub * chunk);
    }
    Scratch sc;
    rc = scratch_for(s, sc);
    if (rc) return rc;
    static thread_local double hpart[kReduceBlocks * 45];
    F8Params fp = {{0, 0}, {0, 0}, 1, 1};
    double acc[45];
    for (int stage = 0; stage < 3; ++stage) {
        const int nv = stage == 0 ? 4 : (stage == 1 ? 2 : 45);
        for (int k = 0; k < nv; ++k) acc[k] = 0.0;
        for (int64_t off = 0; off < n; off += chunk) {
            const int64_t m = (n - off) < chunk ? (n - off) : chunk;
            const void* a = static_cast<const char*>(u1) + off * ub; const void* b = static_cast<const char*>(u2) + off * ub;
            if (mem == TRGL_MEM_HOST) {
                CK(cudaMemcpyAsync(d1, a, ub * m, cudaMemcpyHostToDevice, s));
                CK(cudaMemcpyAsync(d2, b, ub * m, cudaMemcpyHostToDevice, s));
                a = d1; b = d2;
            }
#define F8_LAUNCH(TI, ST) k_f8_reduce<TI, ST><<<kReduceBlocks, kThreads, 0, s>>>(static_cast<const TI*>(a), static_cast<const TI*>(b), fp, sc.partials, m)
            if (mi.in_bytes == 8) { if (stage == 0) F8_LAUNCH(double, 0); else if (stage == 1) F8_LAUNCH(double, 1); else F8_LAUNCH(double, 2); }
            else { if (stage == 0) F8_LAUNCH(float, 0); else if (stage == 1) F8_LAUNCH(float, 1); else F8_LAUNCH(float, 2); }
#undef F8_LAUNCH
            g_launches++;
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(hpart, sc.partials, sizeof(double) * kReduceBlocks * nv, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            for (int b2 = 0; b2 < kReduceBlocks; ++b2)
                for (int k = 0; k < nv; ++k) acc[k] += hpart[b2 * nv + k];
        }
        if (stage == 0) { fp.m1[0] = acc[0] / n; fp.m1[1] = acc[1] / n; fp.m2[0] = acc[2] / n; fp.m2[1] = acc[3] / n; }
        if (stage == 1) { fp.s1 = std::sqrt(2.0) / (acc[0] / n); fp.s2 = std::sqrt(2.0) / (acc[1] / n); }
    }
    double S[81], V[81], w[9];
    int k = 0;
    for (int p = 0; p < 9; ++p)
        for (int q = p; q < 9; ++q) { S[p * 9 + q] = S[q * 9 + p] = acc[k]; ++k; }
    host_jacobi_svd(S, 9, 9, V, w);
    int jmin = 0;
    for (int j = 1; j < 9; ++j) if (w[j] < w[jmin]) jmin = j;
    double F0[9], V3[9], w3[3];
    for (int i = 0; i < 9; ++i) F0[i] = V[i * 9 + jmin];
    // rank-2 projection: F0 = U diag(w) V^T, zero the smallest singular value
    double U3[9];
    for (int i = 0; i < 9; ++i) U3[i] = F0[i];
    host_jacobi_svd(U3, 3, 3, V3, w3);           // columns of U3 are now u_j * w_j
    int j3 = 0;
    for (int j = 1; j < 3; ++j) if (w3[j] < w3[j3]) j3 = j;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            double v = 0;
            for (int j = 0; j < 3; ++j) if (j != j3) v += U3[r * 3 + j] * V3[c * 3 + j];
            F0[r * 3 + c] = v;
        }
    // F = T2^T F0 T1 with T = [[s,0,-s mx],[0,s,-s my],[0,0,1]]
    const double T1[9] = {fp.s1, 0, -fp.s1 * fp.m1[0], 0, fp.s1, -fp.s1 * fp.m1[1], 0, 0, 1};
    const double T2[9] = {fp.s2, 0, -fp.s2 * fp.m2[0], 0, fp.s2, -fp.s2 * fp.m2[1], 0, 0, 1};
    double tmp[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) { double v = 0; for (int j = 0; j < 3; ++j) v += T2[j * 3 + r] * F0[j * 3 + c]; tmp[r * 3 + c] = v; }
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) { double v = 0; for (int j = 0; j < 3; ++j) v += tmp[r * 3 + j] * T1[j * 3 + c]; F[r * 3 + c] = v; }
    if (std::fabs(F[8]) > 1.1920929e-07) { const double inv = 1.0 / F[8]; for (int i = 0; i < 9; ++i) F[i] *= inv; }
    return TRGL_OK;
}

// ---------------------------------------------------------------------------------------------------------------
int trgl_reproj_error(const void* x, const void* imgp, const double* K, const double* dist, const double* rvec,
                      const double* tvec, void* proj, double* sums, double* abs_sums, int64_t n, int x_is_f32,
                      int img_is_f32, int mem, void* stream) {
    if (n < 0) return fail(TRGL_E_BADARG, "negative point count");
    if (!K || !rvec || !tvec || !sums) return fail(TRGL_E_BADARG, "NULL parameter pointer");
    if (mem != TRGL_MEM_HOST && mem != TRGL_MEM_DEVICE) return fail(TRGL_E_BADARG, "unknown memory space");
    if (n > 0 && (!x || !imgp)) return fail(TRGL_E_BADARG, "NULL array pointer with n > 0");
    if (!have_device()) return fail(TRGL_E_NODEVICE, "no CUDA device available (libtriangl_cuda has no CPU fallback)");
    sums[0] = sums[1] = sums[2] = 0.0;
    if (abs_sums) abs_sums[0] = abs_sums[1] = 0.0;
    if (n == 0) return TRGL_OK;
    const ProjParams pp = make_proj_params(K, dist, rvec, tvec);
    const size_t xb = x_is_f32 ? 4 : 8, ib = img_is_f32 ? 4 : 8;
    double acc[5] = {0, 0, 0, 0, 0};
    auto run = [&](const void* dx, const void* di, void* dp, int64_t m, cudaStream_t s) -> int {
        Scratch sc;
        int rcs = scratch_for(s, sc);
        if (rcs) return rcs;
        double* part = sc.partials;
        const int blocks = kReduceBlocks;
        if (x_is_f32) {
            if (img_is_f32) k_reproj_error<float, float><<<blocks, kThreads, 0, s>>>(static_cast<const float*>(dx), static_cast<const float*>(di), pp, static_cast<float*>(dp), part, m);
            else k_reproj_error<float, double><<<blocks, kThreads, 0, s>>>(static_cast<const float*>(dx), static_cast<const double*>(di), pp, static_cast<double*>(dp), part, m);
        } else {
            if (img_is_f32) k_reproj_error<double, float><<<blocks, kThreads, 0, s>>>(static_cast<const double*>(dx), static_cast<const float*>(di), pp, static_cast<float*>(dp), part, m);
            else k_reproj_error<double, double><<<blocks, kThreads, 0, s>>>(static_cast<const double*>(dx), static_cast<const double*>(di), pp, static_cast<double*>(dp), part, m);
        }
        g_launches++;
        CK(cudaGetLastError());
        static thread_local double hpart[kReduceBlocks * 5];
        CK(cudaMemcpyAsync(hpart, part, sizeof(double) * blocks * 5, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        for (int b = 0; b < blocks; ++b)
            for (int k = 0; k < 5; ++k) acc[k] += hpart[b * 5 + k];
        return TRGL_OK;
    };
    int rc;
    if (mem == TRGL_MEM_DEVICE) {
        rc = run(x, imgp, proj, n, static_cast<cudaStream_t>(stream));
        if (rc) return rc;
    } else {
        std::lock_guard<std::mutex> lock(g_pipe_mutex);
        const int64_t chunk = n < kChunk ? n : kChunk;
        const size_t need = align256(3 * xb * chunk) + 2 * align256(2 * ib * chunk);
        rc = ensure_slot(g_slots[0], need);
        if (rc) return rc;
        Slot& sl = g_slots[0];
        for (int64_t off = 0; off < n; off += chunk) {
            const int64_t m = (n - off) < chunk ? (n - off) : chunk;
            char* dx = sl.buf;
            char* di = dx + align256(3 * xb * chunk);
            char* dp = di + align256(2 * ib * chunk);
            CK(cudaMemcpyAsync(dx, static_cast<const char*>(x) + off * 3 * xb, 3 * xb * m, cudaMemcpyHostToDevice, sl.stream));
            CK(cudaMemcpyAsync(di, static_cast<const char*>(imgp) + off * 2 * ib, 2 * ib * m, cudaMemcpyHostToDevice, sl.stream));
            rc = run(dx, di, proj ? dp : nullptr, m, sl.stream);
            if (rc) return rc;
            if (proj) CK(cudaMemcpyAsync(static_cast<char*>(proj) + off * 2 * ib, dp, 2 * ib * m, cudaMemcpyDeviceToHost, sl.stream));
        }
        CK(cudaStreamSynchronize(sl.stream));
    }
    sums[0] = acc[0]; sums[1] = acc[1]; sums[2] = acc[2];
    if (abs_sums) { abs_sums[0] = acc[3]; abs_sums[1] = acc[4]; }
    return TRGL_OK;
}

static int impl_pair_reproj(const void* x, const void* u1, const void* u2, const double* P1, const double* P2,
                            const void* status, int status_is_i32, int min_status, double max_sq_err, void* err1, void* err2,
                            uint8_t* good, double* sums, double* sums_dev, int64_t n, int mode, int mem, void* stream) {
    ModeInfo mi;
    if (n < 0) return fail(TRGL_E_BADARG, "negative point count");
    if (!mode_info(mode, mi)) return fail(TRGL_E_BADARG, "unknown precision mode");
    if (mem != TRGL_MEM_HOST && mem != TRGL_MEM_DEVICE) return fail(TRGL_E_BADARG, "unknown memory space");
    if (!P1 || !P2 || (!sums && !sums_dev)) return fail(TRGL_E_BADARG, "NULL parameter pointer");
    if (sums_dev && mem != TRGL_MEM_DEVICE) return fail(TRGL_E_BADARG, "the asynchronous variant needs device buffers");
    if (n > 0 && (!x || !u1 || !u2 || !status)) return fail(TRGL_E_BADARG, "NULL array pointer with n > 0");
    if (!have_device()) return fail(TRGL_E_NODEVICE, "no CUDA device available (libtriangl_cuda has no CPU fallback)");
    if (sums) sums[0] = sums[1] = sums[2] = sums[3] = 0.0;
    if (n == 0) {
        if (sums_dev) CK(cudaMemsetAsync(sums_dev, 0, 4 * sizeof(double), static_cast<cudaStream_t>(stream)));
        return TRGL_OK;
    }
    double acc[4] = {0, 0, 0, 0};
    const Cams<double> cams = make_cams<double>(P1, P2);
    auto run = [&](const void* dx, const void* d1, const void* d2, const void* dst, void* de1, void* de2, uint8_t* dg,
                   int64_t m, cudaStream_t s) -> int {
        const int blocks = kReduceBlocks;
        Scratch sc;
        int rcs = scratch_for(s, sc);
        if (rcs) return rcs;
#define PAIR_LAUNCH(TI, TO)                                                                                      \
        if (status_is_i32)                                                                                       \
            k_pair_reproj<TI, TO, int32_t><<<blocks, kThreads, 0, s>>>(static_cast<const TO*>(dx), static_cast<const TI*>(d1), static_cast<const TI*>(d2), cams, static_cast<const int32_t*>(dst), min_status, max_sq_err, static_cast<TO*>(de1), static_cast<TO*>(de2), dg, sc.partials, m, sc.counter, sums_dev); \
        else                                                                                                     \
            k_pair_reproj<TI, TO, uint8_t><<<blocks, kThreads, 0, s>>>(static_cast<const TO*>(dx), static_cast<const TI*>(d1), static_cast<const TI*>(d2), cams, static_cast<const uint8_t*>(dst), min_status, max_sq_err, static_cast<TO*>(de1), static_cast<TO*>(de2), dg, sc.partials, m, sc.counter, sums_dev);
        if (mi.in_bytes == 8 && mi.out_bytes == 8) { PAIR_LAUNCH(double, double) }
        else if (mi.in_bytes == 4 && mi.out_bytes == 4) { PAIR_LAUNCH(float, float) }
        else if (mi.in_bytes == 8 && mi.out_bytes == 4) { PAIR_LAUNCH(double, float) }
        else { PAIR_LAUNCH(float, double) }
#undef PAIR_LAUNCH
        g_launches++;
        CK(cudaGetLastError());
        if (sums_dev) return TRGL_OK;                  // asynchronous variant: the result stays on the device
        static thread_local double hpart[kReduceBlocks * 4];
        CK(cudaMemcpyAsync(hpart, sc.partials, sizeof(double) * blocks * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        for (int b = 0; b < blocks; ++b)
            for (int k = 0; k < 4; ++k) acc[k] += hpart[b * 4 + k];
        return TRGL_OK;
    };
    int rc;
    if (mem == TRGL_MEM_DEVICE) {
        rc = run(x, u1, u2, status, err1, err2, good, n, static_cast<cudaStream_t>(stream));
        if (rc) return rc;
    } else {
        std::lock_guard<std::mutex> lock(g_pipe_mutex);
        const int64_t chunk = n < kChunk ? n : kChunk;
        const size_t sb = status_is_i32 ? 4 : 1;
        const size_t bx = align256(3 * mi.out_bytes * chunk), bu = align256(2 * mi.in_bytes * chunk),
                     bs = align256(sb * chunk), be = align256(mi.out_bytes * chunk), bg = align256(chunk);
        rc = ensure_slot(g_slots[0], bx + 2 * bu + bs + 2 * be + bg);
        if (rc) return rc;
        Slot& sl = g_slots[0];
        char* dx = sl.buf; char* d1 = dx + bx; char* d2 = d1 + bu; char* ds = d2 + bu;
        char* de1 = ds + bs; char* de2 = de1 + be; char* dg = de2 + be;
        for (int64_t off = 0; off < n; off += chunk) {
            const int64_t m = (n - off) < chunk ? (n - off) : chunk;
            CK(cudaMemcpyAsync(dx, static_cast<const char*>(x) + off * 3 * mi.out_bytes, 3 * mi.out_bytes * m, cudaMemcpyHostToDevice, sl.stream));
            CK(cudaMemcpyAsync(d1, static_cast<const char*>(u1) + off * 2 * mi.in_bytes, 2 * mi.in_bytes * m, cudaMemcpyHostToDevice, sl.stream));
            CK(cudaMemcpyAsync(d2, static_cast<const char*>(u2) + off * 2 * mi.in_bytes, 2 * mi.in_bytes * m, cudaMemcpyHostToDevice, sl.stream));
            CK(cudaMemcpyAsync(ds, static_cast<const char*>(status) + off * sb, sb * m, cudaMemcpyHostToDevice, sl.stream));
            rc = run(dx, d1, d2, ds, err1 ? de1 : nullptr, err2 ? de2 : nullptr, good ? reinterpret_cast<uint8_t*>(dg) : nullptr, m, sl.stream);
            if (rc) return rc;
            if (err1) CK(cudaMemcpyAsync(static_cast<char*>(err1) + off * mi.out_bytes, de1, mi.out_bytes * m, cudaMemcpyDeviceToHost, sl.stream));
            if (err2) CK(cudaMemcpyAsync(static_cast<char*>(err2) + off * mi.out_bytes, de2, mi.out_bytes * m, cudaMemcpyDeviceToHost, sl.stream));
            if (good) CK(cudaMemcpyAsync(good + off, dg, m, cudaMemcpyDeviceToHost, sl.stream));
        }
        CK(cudaStreamSynchronize(sl.stream));
    }
    if (sums)
        for (int k = 0; k < 4; ++k) sums[k] = acc[k];
    return TRGL_OK;
}

int trgl_pair_reproj(const void* x, const void* u1, const void* u2, const double* P1, const double* P2,
                     const void* status, int status_is_i32, int min_status, double max_sq_err, void* err1, void* err2,
                     uint8_t* good, double* sums, int64_t n, int mode, int mem, void* stream) {
    if (!sums) return fail(TRGL_E_BADARG, "NULL parameter pointer");
    return impl_pair_reproj(x, u1, u2, P1, P2, status, status_is_i32, min_status, max_sq_err, err1, err2, good, sums, nullptr,
                            n, mode, mem, stream);
}

int trgl_pair_reproj_async(const void* x, const void* u1, const void* u2, const double* P1, const double* P2,
                           const void* status, int status_is_i32, int min_status, double max_sq_err, void* err1, void* err2,
                           uint8_t* good, double* sums_device, int64_t n, int mode, void* stream) {
    if (!sums_device) return fail(TRGL_E_BADARG, "NULL parameter pointer");
    return impl_pair_reproj(x, u1, u2, P1, P2, status, status_is_i32, min_status, max_sq_err, err1, err2, good, nullptr,
                            sums_device, n, mode, TRGL_MEM_DEVICE, stream);
}

// ---------------------------------------------------------------------------------------------------------------
// Harness statistics on the device (triangulation_comparison.py:179-260): per-point squared errors, their sum, the
// false-positive / false-negative counts of robustness_stat, and the exact median (radix selection).
static int eval_errors_common(bool three_d, const void* a, const double* exact, int exact_stride, const void* status,
                              int status_is_i32, double thresh_max, double thresh_min, double* errors, double* stats,
                              int64_t n, int a_is_f32, int mem, void* stream) {
    if (n < 0) return fail(TRGL_E_BADARG, "negative point count");
    if (mem != TRGL_MEM_HOST && mem != TRGL_MEM_DEVICE) return fail(TRGL_E_BADARG, "unknown memory space");
    if (!stats) return fail(TRGL_E_BADARG, "stats is NULL");
    if (n > 0 && (!a || !exact)) return fail(TRGL_E_BADARG, "NULL array pointer with n > 0");
    if (three_d && exact_stride < 3) return fail(TRGL_E_BADARG, "exact_stride must be >= 3");
    if (!have_device()) return fail(TRGL_E_NODEVICE, "no CUDA device available (libtriangl_cuda has no CPU fallback)");
    stats[0] = stats[1] = stats[2] = stats[3] = 0.0;
    if (n == 0) return TRGL_OK;
    std::unique_lock<std::mutex> lock(g_pipe_mutex, std::defer_lock);
    if (mem == TRGL_MEM_HOST) lock.lock();
    int rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t ab = (a_is_f32 ? 4 : 8) * (three_d ? 3 : 2), eb = 8 * (three_d ? exact_stride : 2), sb = status_is_i32 ? 4 : 1;
    const void* da = a; const double* de = exact; const void* dst = status; double* derr = errors;
    if (mem == TRGL_MEM_HOST) {
        const size_t oa = 0, oe = align256(ab * n), os = oe + align256(eb * n), oerr = os + align256(sb * n);
        rc = ensure_slot(g_slots[0], oerr + align256(8 * n));
        if (rc) return rc;
        Slot& sl = g_slots[0];
        s = sl.stream;
        CK(cudaMemcpyAsync(sl.buf + oa, a, ab * n, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(sl.buf + oe, exact, eb * n, cudaMemcpyHostToDevice, s));
        if (status) CK(cudaMemcpyAsync(sl.buf + os, status, sb * n, cudaMemcpyHostToDevice, s));
        da = sl.buf + oa; de = reinterpret_cast<const double*>(sl.buf + oe); dst = status ? sl.buf + os : nullptr;
        derr = errors ? reinterpret_cast<double*>(sl.buf + oerr) : nullptr;
    }
    Scratch sc;
    rc = scratch_for(s, sc);
    if (rc) return rc;
    const int blocks = kReduceBlocks;
    if (three_d) {
#define E3(TO, TS) k_sq_errors_3d<TO, TS><<<blocks, kThreads, 0, s>>>(static_cast<const TO*>(da), de, exact_stride, static_cast<const TS*>(dst), thresh_max, thresh_min, derr, sc.partials, n)
        if (a_is_f32) { if (status_is_i32) E3(float, int32_t); else E3(float, uint8_t); }
        else { if (status_is_i32) E3(double, int32_t); else E3(double, uint8_t); }
#undef E3
    } else {
        if (a_is_f32) k_sq_errors_2d<float><<<blocks, kThreads, 0, s>>>(static_cast<const float*>(da), de, derr, sc.partials, n);
        else k_sq_errors_2d<double><<<blocks, kThreads, 0, s>>>(static_cast<const double*>(da), de, derr, sc.partials, n);
    }
    g_launches++;
    CK(cudaGetLastError());
    static thread_local double hpart[kReduceBlocks * 4];
    CK(cudaMemcpyAsync(hpart, sc.partials, sizeof(double) * blocks * 4, cudaMemcpyDeviceToHost, s));
    if (mem == TRGL_MEM_HOST && errors) CK(cudaMemcpyAsync(errors, derr, 8 * n, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (int b = 0; b < blocks; ++b)
        for (int k = 0; k < 4; ++k) stats[k] += hpart[b * 4 + k];
    return TRGL_OK;
}

int trgl_eval_errors_3d(const void* x, const double* exact, int exact_stride, const void* status, int status_is_i32,
                        double thresh_max, double thresh_min, double* errors, double* stats, int64_t n, int x_is_f32,
                        int mem, void* stream) {
    return eval_errors_common(true, x, exact, exact_stride, status, status_is_i32, thresh_max, thresh_min, errors, stats, n,
                              x_is_f32, mem, stream);
}

int trgl_eval_errors_2d(const void* proj, const double* exact, double* errors, double* stats, int64_t n, int proj_is_f32,
                        int mem, void* stream) {
    return eval_errors_common(false, proj, exact, 2, nullptr, 0, 0.0, 0.0, errors, stats, n, proj_is_f32, mem, stream);
}

int trgl_vector_stat(const void* x_trials, const double* exact, int exact_stride, int trials, double* means, double* covars,
                     int64_t n, int x_is_f32, int mem, void* stream) {
    if (n < 0 || trials < 1) return fail(TRGL_E_BADARG, "need n >= 0 and trials >= 1");
    if (mem != TRGL_MEM_HOST && mem != TRGL_MEM_DEVICE) return fail(TRGL_E_BADARG, "unknown memory space");
    if (!means || !covars || (n > 0 && (!x_trials || !exact))) return fail(TRGL_E_BADARG, "NULL pointer");
    if (exact_stride < 3) return fail(TRGL_E_BADARG, "exact_stride must be >= 3");
    if (!have_device()) return fail(TRGL_E_NODEVICE, "no CUDA device available (libtriangl_cuda has no CPU fallback)");
    if (n == 0) return TRGL_OK;
    std::unique_lock<std::mutex> lock(g_pipe_mutex, std::defer_lock);
    if (mem == TRGL_MEM_HOST) lock.lock();
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t xb = static_cast<size_t>(x_is_f32 ? 4 : 8) * 3 * n * trials, eb = sizeof(double) * exact_stride * n;
    const void* dx = x_trials; const double* de = exact; double* dm = means; double* dc = covars;
    if (mem == TRGL_MEM_HOST) {
        const size_t oe = align256(xb), om = oe + align256(eb), oc = om + align256(24 * n);
        int rc = ensure_slot(g_slots[0], oc + align256(72 * n));
        if (rc) return rc;
        Slot& sl = g_slots[0];
        s = sl.stream;
        CK(cudaMemcpyAsync(sl.buf, x_trials, xb, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(sl.buf + oe, exact, eb, cudaMemcpyHostToDevice, s));
        dx = sl.buf; de = reinterpret_cast<const double*>(sl.buf + oe);
        dm = reinterpret_cast<double*>(sl.buf + om); dc = reinterpret_cast<double*>(sl.buf + oc);
    }
    const int64_t tiles = (n + kThreads - 1) / kThreads;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
    const unsigned grid = static_cast<unsigned>(tiles < cap ? tiles : cap);
    if (x_is_f32) k_vector_stat<float><<<grid, kThreads, 0, s>>>(static_cast<const float*>(dx), de, exact_stride, trials, dm, dc, n);
    else k_vector_stat<double><<<grid, kThreads, 0, s>>>(static_cast<const double*>(dx), de, exact_stride, trials, dm, dc, n);
    g_launches++;
    CK(cudaGetLastError());
    if (mem == TRGL_MEM_HOST) {
        CK(cudaMemcpyAsync(means, dm, 24 * n, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(covars, dc, 72 * n, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    }
    return TRGL_OK;
}

int trgl_median(const double* values, int64_t n, int mem, double* median, void* stream) {
    if (n < 0) return fail(TRGL_E_BADARG, "negative count");
    if (mem != TRGL_MEM_HOST && mem != TRGL_MEM_DEVICE) return fail(TRGL_E_BADARG, "unknown memory space");
    if (!median || (n > 0 && !values)) return fail(TRGL_E_BADARG, "NULL pointer");
    if (!have_device()) return fail(TRGL_E_NODEVICE, "no CUDA device available (libtriangl_cuda has no CPU fallback)");
    *median = std::nan("");
    if (n == 0) return TRGL_OK;                       // np.median of an empty array is NaN
    std::unique_lock<std::mutex> lock(g_pipe_mutex, std::defer_lock);
    if (mem == TRGL_MEM_HOST) lock.lock();
    int rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const double* dv = values;
    if (mem == TRGL_MEM_HOST) {
        rc = ensure_slot(g_slots[0], 8 * n);
        if (rc) return rc;
        s = g_slots[0].stream;
        CK(cudaMemcpyAsync(g_slots[0].buf, values, 8 * n, cudaMemcpyHostToDevice, s));
        dv = reinterpret_cast<const double*>(g_slots[0].buf);
    }
    Scratch sc;
    rc = scratch_for(s, sc);
    if (rc) return rc;
    unsigned long long* dh = reinterpret_cast<unsigned long long*>(sc.partials);   // the partials block doubles as the histogram
    const int64_t tiles = (n + kThreads - 1) / kThreads;
    const unsigned blocks = static_cast<unsigned>(tiles < kReduceBlocks ? tiles : kReduceBlocks);
    unsigned long long hist[257];
    // rank of the lower middle element (0-based); keys: non-negative doubles sort like their bit patterns, NaNs
    // (0x7ff8...) sort above +inf, negative values are not supported (the inputs are squared errors)
    int64_t rank = (n - 1) / 2;
    unsigned long long prefix = 0, mask = 0;
    int64_t below = 0, eq = 0;
    for (int pass = 7; pass >= 0; --pass) {
        const int shift = 8 * pass;
        CK(cudaMemsetAsync(dh, 0, sizeof(hist), s));
        k_radix_hist<<<blocks, kThreads, 0, s>>>(dv, n, prefix, mask, shift, dh);
        g_launches++;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(hist, dh, sizeof(hist), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (pass == 7) {
            if (hist[256]) return TRGL_OK;            // np.median: any NaN in the sample gives NaN
            for (int b2 = 128; b2 < 256; ++b2)
                if (hist[b2]) return fail(TRGL_E_BADARG, "median: negative values are not supported (inputs are squared errors)");
        }
        int64_t cum = 0;
        int b = 0;
        for (; b < 256; ++b) {
            if (cum + static_cast<int64_t>(hist[b]) > rank) break;
            cum += static_cast<int64_t>(hist[b]);
        }
        if (b == 256) return fail(TRGL_E_BADARG, "median: inconsistent histogram (negative values?)");
        rank -= cum; below += cum; eq = static_cast<int64_t>(hist[b]);
        prefix |= static_cast<unsigned long long>(b) << shift;
        mask |= 255ull << shift;
    }
    unsigned long long key_lo = prefix, key_hi = prefix;
    if (n % 2 == 0 && below + eq <= n / 2) {          // the upper middle element is the next larger key
        unsigned long long init = ~0ull;
        CK(cudaMemcpyAsync(dh, &init, sizeof(init), cudaMemcpyHostToDevice, s));
        k_min_above<<<blocks, kThreads, 0, s>>>(dv, n, prefix, dh);
        g_launches++;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(&key_hi, dh, sizeof(key_hi), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    }
    double lo, hi;
    std::memcpy(&lo, &key_lo, 8); std::memcpy(&hi, &key_hi, 8);
    *median = (n % 2) ? lo : 0.5 * (lo + hi);
    return TRGL_OK;
}

}  // extern "C"
