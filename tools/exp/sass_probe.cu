// Static inspection TU: instantiates the FP64 hot kernels only, so that ptxas -v and cuobjdump -sass answer in seconds
// (tools/sass_count.sh).
#include "../../multiple-quadrotor-slam_b200/csrc/trgl_kernels.cuh"
#include "../../multiple-quadrotor-slam_b200/csrc/trgl_multiview.cuh"
using namespace trgl;
#ifndef PROBE_EVAL
#define PROBE_EVAL true
#endif
const void* probe_table[] = {
    reinterpret_cast<const void*>(&k_linear_eigen<double, double, double, 4, PreNone, PROBE_EVAL>),
    reinterpret_cast<const void*>(&k_polynomial<double, double, double, 4, PreNone, PROBE_EVAL>),
    reinterpret_cast<const void*>(&k_iterative_ls<double, double, double, PreNone, PROBE_EVAL>),
    reinterpret_cast<const void*>(&k_linear_ls<double, double, double, 4, PreNone, PROBE_EVAL, Mirrors, true>),
    reinterpret_cast<const void*>(&k_linear_ls<double, double, double, 4, PreNone, false, NoMirrors, true>),
    reinterpret_cast<const void*>(&k_linear_eigen_general<double, double, double, 4, PreNone, PROBE_EVAL>),
    reinterpret_cast<const void*>(&k_polynomial_general<double, double, double, 4, PreNone, PROBE_EVAL>),
    reinterpret_cast<const void*>(&k_linear_ls_f32x4),
    reinterpret_cast<const void*>(&k_multiview_ls<double, double, double, 1, 8, 1, true>),
    reinterpret_cast<const void*>(&k_multiview_ls<double, double, double, 1, 8, 1, false>),
    reinterpret_cast<const void*>(&k_linear_ls_general<float, double, float, PreNone, false>),
    reinterpret_cast<const void*>(&k_iterative_general<double, double, double, PreNone, PROBE_EVAL>),
    reinterpret_cast<const void*>(&k_linear_ls<double, double, double, 4, PreNone, PROBE_EVAL, NoMirrors, true>),
};
