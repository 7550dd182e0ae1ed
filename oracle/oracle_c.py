"""
ctypes loader of the C restatement of the oracle (oracle/c/triangl_oracle.c) -- TEST INFRASTRUCTURE ONLY.
Same call conventions as oracle/triangulation_oracle.py.  OpenMP over points, like the reference's C extension
(Work/python_libs/triangulation_c/triangulation.c:70,109).
"""
import ctypes
import os
import subprocess

import numpy as np

from oracle import triangulation_oracle as orc

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "c", "libtriangl_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB):
            subprocess.check_call(["make", "-C", os.path.join(_HERE, "c")])
        _lib = ctypes.CDLL(LIB)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _prep(u1, P1, u2, P2):
    u1 = np.ascontiguousarray(u1, dtype=np.float64).reshape(-1, 2)
    u2 = np.ascontiguousarray(u2, dtype=np.float64).reshape(-1, 2)
    P1 = np.ascontiguousarray(np.asarray(P1, dtype=np.float64)[0:3, :]); P2 = np.ascontiguousarray(np.asarray(P2, dtype=np.float64)[0:3, :])
    return u1, P1, u2, P2, len(u1)


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n):
    lib().orc_set_num_threads(ctypes.c_int(int(n)))


def linear_LS_triangulation(u1, P1, u2, P2):
    u1, P1, u2, P2, n = _prep(u1, P1, u2, P2)
    x = np.empty((n, 3)); st = np.empty(n, dtype=np.uint8)
    lib().orc_linear_ls(_p(u1), _p(u2), _p(P1), _p(P2), _p(x), _p(st), ctypes.c_int64(n))
    return x, st.view(np.bool_)


def ls_condition(u1, P1, u2, P2):
    """s_max / s_min of the unweighted 4x3 system of every point (parity-test diagnostics)."""
    u1, P1, u2, P2, n = _prep(u1, P1, u2, P2)
    cond = np.empty(n)
    lib().orc_ls_condition(_p(u1), _p(u2), _p(P1), _p(P2), _p(cond), ctypes.c_int64(n))
    return cond


def iterative_LS_triangulation(u1, P1, u2, P2, tolerance=3.e-5, semantics='c', return_nsolves=False, return_margin=False):
    """return_margin: also the knife-edge margin of the deciding convergence test (oracle.iterative_LS_core)."""
    u1, P1, u2, P2, n = _prep(u1, P1, u2, P2)
    x = np.empty((n, 3)); st = np.empty(n, dtype=np.int32); ns = np.empty(n, dtype=np.int32)
    mg = np.empty(n) if return_margin else None
    lib().orc_iterative_ls(_p(u1), _p(u2), _p(P1), _p(P2), _p(x), _p(st), _p(ns), _p(mg) if return_margin else None,
                           ctypes.c_int64(n), ctypes.c_double(tolerance), ctypes.c_int(1 if semantics == 'py' else 0))
    out = (x, st)
    if return_nsolves:
        out += (ns,)
    if return_margin:
        out += (mg,)
    return out


def linear_eigen_triangulation(u1, P1, u2, P2, max_coordinate_value=1.e16, rows=4, return_amp=False):
    """return_amp: also s1 / ((s3 - s4) |w|), the error amplification of the dehomogenised singular vector."""
    u1, P1, u2, P2, n = _prep(u1, P1, u2, P2)
    x = np.empty((n, 3)); st = np.empty(n, dtype=np.uint8)
    amp = np.empty(n) if return_amp else None
    lib().orc_linear_eigen(_p(u1), _p(u2), _p(P1), _p(P2), _p(x), _p(st), _p(amp) if return_amp else None, ctypes.c_int64(n),
                           ctypes.c_double(max_coordinate_value), ctypes.c_int(rows))
    return (x, st.view(np.bool_), amp) if return_amp else (x, st.view(np.bool_))


def correct_matches(F, u1, u2):
    u1 = np.ascontiguousarray(u1, dtype=np.float64).reshape(-1, 2); u2 = np.ascontiguousarray(u2, dtype=np.float64).reshape(-1, 2)
    F = np.ascontiguousarray(F, dtype=np.float64)
    n1 = np.empty_like(u1); n2 = np.empty_like(u2)
    lib().orc_correct_matches(_p(F), _p(u1), _p(u2), _p(n1), _p(n2), ctypes.c_int64(len(u1)))
    return n1, n2


def polynomial_triangulation(u1, P1, u2, P2, rows=4, return_amp=False):
    u1, P1, u2, P2, n = _prep(u1, P1, u2, P2)
    F = np.ascontiguousarray(orc.fundamental_from_P(P1, P2))
    x = np.empty((n, 3)); st = np.empty(n, dtype=np.uint8)
    amp = np.empty(n) if return_amp else None
    lib().orc_polynomial(_p(F), _p(u1), _p(u2), _p(P1), _p(P2), _p(x), _p(st), _p(amp) if return_amp else None,
                         ctypes.c_int64(n), ctypes.c_double(1.e16), ctypes.c_int(rows))
    return (x, st.view(np.bool_), amp) if return_amp else (x, st.view(np.bool_))


SOLVERS = {
    'linear_eigen': linear_eigen_triangulation,
    'linear_LS': linear_LS_triangulation,
    'iterative_LS': iterative_LS_triangulation,
    'polynomial': polynomial_triangulation,
}
