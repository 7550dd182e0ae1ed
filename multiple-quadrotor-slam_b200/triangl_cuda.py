"""
ctypes binding of libtriangl_cuda.so (C ABI: include/triangl_cuda.h).

This is the thin host layer the reference keeps in Work/python_libs/triangulation_c/__init__.py:18-86
(coerce inputs, allocate outputs, call the native function in place) -- with the weave/CPython-2 extension
replaced by a plain shared library.  There is no CPU fallback: if the library or a CUDA device is missing,
importing succeeds (so the symbols can be inspected) but every compute call raises TrianglCudaError.

Besides NumPy arrays, every solver accepts device-resident buffers (`DeviceArray`, or any object with
`data_ptr()` such as a CUDA torch.Tensor) for roofline runs without PCIe traffic.
"""
import ctypes
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# TRGL_CUDA_LIB: another build of the same library (tools/ab_variants.sh compares compile-time variants on the GPU box)
LIB_PATH = os.environ.get("TRGL_CUDA_LIB") or os.path.join(_HERE, "libtriangl_cuda.so")

F64, F32IO, F32, F64_OUT32, F32_OUT64 = 0, 1, 2, 3, 4
MEM_HOST, MEM_DEVICE, MEM_DEVICE_IN = 0, 1, 2
PINNED_OUTPUT_MIN_POINTS = 1 << 16      # host-mode outputs at least this long are allocated page-locked
ITER_C, ITER_PY = 0, 1

EXPORTS = [
    "trgl_version", "trgl_last_error_string", "trgl_device_count", "trgl_set_device", "trgl_device_synchronize",
    "trgl_device_alloc", "trgl_device_free", "trgl_host_alloc", "trgl_host_free", "trgl_memcpy_h2d",
    "trgl_memcpy_d2h", "trgl_memcpy_d2d", "trgl_memset_d", "trgl_stream_create", "trgl_stream_destroy", "trgl_stream_synchronize",
    "trgl_event_create", "trgl_event_destroy", "trgl_event_record", "trgl_event_synchronize", "trgl_event_elapsed_ms",
    "trgl_polynomial_flags_async",
    "trgl_linear_ls", "trgl_iterative_ls", "trgl_linear_eigen", "trgl_polynomial", "trgl_polynomial_F",
    "trgl_fundamental_8point", "trgl_reproj_error", "trgl_pair_reproj", "trgl_launch_count",
    "trgl_set_points_per_thread", "trgl_set_stream_variant", "trgl_set_two_ray", "trgl_multiview_ls", "trgl_set_deferred_capacity",
    "trgl_eval_errors_3d", "trgl_eval_errors_2d", "trgl_median", "trgl_pair_reproj_async",
    "trgl_set_fused_eval", "trgl_set_result_mirrors", "trgl_ipc_export", "trgl_ipc_import", "trgl_ipc_close",
    "trgl_set_result_mirrors_f32", "trgl_set_input_retention", "trgl_deferred_total", "trgl_vector_stat", "trgl_set_trace", "trgl_get_trace",
    "trgl_fp64_fma_rate", "trgl_rare_path_counters",
    "trgl_undistort_points", "trgl_linear_ls_px", "trgl_iterative_ls_px", "trgl_linear_eigen_px", "trgl_polynomial_px",
]


class TrianglCudaError(RuntimeError):
    pass


_lib = None


def lib():
    """Load the shared library once; fail loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise TrianglCudaError(
            "libtriangl_cuda.so is missing (%s): build it with `make -C %s` or __graft_entry__.build(); "
            "there is no CPU fallback" % (LIB_PATH, os.path.join(_HERE, "csrc")))
    L = ctypes.CDLL(LIB_PATH)
    vp, i64, dbl, cint = ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_int
    dp = vp           # double* parameters take plain integer addresses: `a.ctypes.data` costs 1 us, `data_as(POINTER)` 2.5 us
    L.trgl_version.restype = cint
    L.trgl_last_error_string.restype = ctypes.c_char_p
    L.trgl_launch_count.restype = i64
    L.trgl_device_alloc.argtypes = [ctypes.POINTER(vp), ctypes.c_size_t]
    L.trgl_device_free.argtypes = [vp]
    L.trgl_host_alloc.argtypes = [ctypes.POINTER(vp), ctypes.c_size_t]
    L.trgl_host_free.argtypes = [vp]
    L.trgl_memcpy_h2d.argtypes = [vp, vp, ctypes.c_size_t, vp]
    L.trgl_memcpy_d2h.argtypes = [vp, vp, ctypes.c_size_t, vp]
    L.trgl_memcpy_d2d.argtypes = [vp, vp, ctypes.c_size_t, vp]
    L.trgl_memset_d.argtypes = [vp, cint, ctypes.c_size_t, vp]
    L.trgl_stream_create.argtypes = [ctypes.POINTER(vp)]
    L.trgl_stream_destroy.argtypes = [vp]
    L.trgl_stream_synchronize.argtypes = [vp]
    L.trgl_event_create.argtypes = [ctypes.POINTER(vp)]
    L.trgl_event_destroy.argtypes = [vp]
    L.trgl_event_record.argtypes = [vp, vp]
    L.trgl_event_synchronize.argtypes = [vp]
    L.trgl_polynomial_flags_async.argtypes = [vp, vp]
    L.trgl_event_elapsed_ms.argtypes = [vp, vp, ctypes.POINTER(ctypes.c_float)]
    L.trgl_linear_ls.argtypes = [vp, vp, dp, dp, vp, vp, i64, cint, cint, vp]
    L.trgl_iterative_ls.argtypes = [vp, vp, dp, dp, vp, vp, i64, dbl, cint, cint, cint, vp]
    L.trgl_set_deferred_capacity.argtypes = [i64]
    L.trgl_set_deferred_capacity.restype = i64
    L.trgl_multiview_ls.argtypes = [vp, vp, dp, cint, vp, vp, i64, cint, cint, cint, vp]
    L.trgl_linear_eigen.argtypes = [vp, vp, dp, dp, vp, vp, i64, dbl, cint, cint, cint, vp]
    L.trgl_polynomial.argtypes = [vp, vp, dp, dp, vp, vp, vp, vp, ctypes.POINTER(cint), i64, dbl, cint, cint, cint, vp]
    L.trgl_polynomial_F.argtypes = [vp, vp, dp, dp, dp, vp, vp, vp, vp, ctypes.POINTER(cint), i64, dbl, cint, cint,
                                    cint, vp]
    L.trgl_fundamental_8point.argtypes = [vp, vp, i64, cint, cint, dp, vp]
    L.trgl_reproj_error.argtypes = [vp, vp, dp, dp, dp, dp, vp, dp, dp, i64, cint, cint, cint, vp]
    L.trgl_pair_reproj.argtypes = [vp, vp, vp, dp, dp, vp, cint, cint, dbl, vp, vp, vp, dp, i64, cint, cint, vp]
    L.trgl_pair_reproj_async.argtypes = [vp, vp, vp, dp, dp, vp, cint, cint, dbl, vp, vp, vp, vp, i64, cint, vp]
    L.trgl_set_fused_eval.argtypes = [cint, dbl, vp, vp, vp, vp]
    L.trgl_set_result_mirrors.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(vp), cint]
    L.trgl_set_result_mirrors_f32.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(vp), cint]
    L.trgl_set_input_retention.argtypes = [vp, vp]
    L.trgl_deferred_total.argtypes = [vp, ctypes.POINTER(i64)]
    L.trgl_ipc_export.argtypes = [vp, vp]
    L.trgl_ipc_import.argtypes = [vp, ctypes.POINTER(vp)]
    L.trgl_ipc_close.argtypes = [vp]
    L.trgl_eval_errors_3d.argtypes = [vp, vp, cint, vp, cint, dbl, dbl, vp, dp, i64, cint, cint, vp]
    L.trgl_eval_errors_2d.argtypes = [vp, vp, vp, dp, i64, cint, cint, vp]
    L.trgl_median.argtypes = [vp, i64, cint, dp, vp]
    L.trgl_vector_stat.argtypes = [vp, vp, cint, cint, vp, vp, i64, cint, cint, vp]
    L.trgl_undistort_points.argtypes = [vp, vp, dp, dp, i64, cint, cint, vp]
    L.trgl_linear_ls_px.argtypes = [vp, vp, dp, dp, dp, dp, dp, dp, vp, vp, i64, cint, cint, vp]
    L.trgl_iterative_ls_px.argtypes = [vp, vp, dp, dp, dp, dp, dp, dp, vp, vp, i64, dbl, cint, cint, cint, vp]
    L.trgl_linear_eigen_px.argtypes = [vp, vp, dp, dp, dp, dp, dp, dp, vp, vp, i64, dbl, cint, cint, cint, vp]
    L.trgl_polynomial_px.argtypes = [vp, vp, dp, dp, dp, dp, dp, dp, vp, vp, vp, vp, ctypes.POINTER(cint), i64, dbl, cint,
                                     cint, cint, vp]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise TrianglCudaError("libtriangl_cuda error %d: %s" % (rc, lib().trgl_last_error_string().decode()))


def device_count():
    return lib().trgl_device_count()


def require_device():
    if device_count() <= 0:
        raise TrianglCudaError("no CUDA device available: libtriangl_cuda has no CPU fallback")


# ---- buffers -----------------------------------------------------------------------------------------------------
class DeviceArray:
    """A typed block of HBM owned by the library allocator (freed on garbage collection)."""

    def __init__(self, shape, dtype):
        self.shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        p = ctypes.c_void_p()
        check(lib().trgl_device_alloc(ctypes.byref(p), self.nbytes))
        self.ptr = p.value or 0

    def data_ptr(self):
        return self.ptr

    def __len__(self):
        return self.shape[0]

    def copy_from_host(self, a, stream=None):
        a = np.ascontiguousarray(a, dtype=self.dtype)
        assert a.nbytes == self.nbytes, (a.shape, self.shape)
        check(lib().trgl_memcpy_h2d(self.ptr, a.ctypes.data, self.nbytes, stream))
        check(lib().trgl_stream_synchronize(stream))
        return self

    def to_host(self, out=None, stream=None, sync=True):
        """sync=False: enqueue the copy only (out should be page-locked: pinned_empty); the caller synchronises."""
        if out is None:
            out = np.empty(self.shape, dtype=self.dtype)
        check(lib().trgl_memcpy_d2h(out.ctypes.data, self.ptr, self.nbytes, stream))
        if sync:
            check(lib().trgl_stream_synchronize(stream))
        return out

    def view(self, offset, shape):
        """Non-owning view of `shape` starting `offset` elements into this buffer (keeps the parent alive)."""
        v = DeviceArray.__new__(DeviceArray)
        v.shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        v.dtype = self.dtype
        v.nbytes = int(np.prod(v.shape, dtype=np.int64)) * self.dtype.itemsize
        if offset < 0 or offset * self.dtype.itemsize + v.nbytes > self.nbytes:
            raise ValueError("view outside the buffer")
        v.ptr = self.ptr + int(offset) * self.dtype.itemsize
        v._parent = self
        return v

    def __del__(self):
        try:
            if getattr(self, "ptr", 0) and getattr(self, "_parent", None) is None:
                _lib.trgl_device_free(self.ptr)
            self.ptr = 0
        except Exception:
            pass


def to_device(a):
    a = np.ascontiguousarray(a)
    return DeviceArray(a.shape, a.dtype).copy_from_host(a)


# Page-locked host memory comes from a small caching allocator: cudaHostAlloc costs ~0.3 ms per MB, far more than the
# PCIe transfer it enables, so blocks are recycled when the last NumPy view of them is garbage collected (the same idea
# as a framework's pinned-memory caching allocator).  Large solver outputs are allocated here so that the D2H copy of
# the chunked pipeline runs asynchronously at full PCIe speed.
_PIN_GRAIN = 1 << 21
_PIN_CACHE_LIMIT = 16 << 30
_pin_lock = threading.Lock()
_pin_free = {}          # rounded size -> [ptr, ...]
_pin_cached_bytes = 0


class _PinnedOwner:
    """Returns one cudaHostAlloc block to the cache when the last NumPy view of it is gone."""

    def __init__(self, ptr, size):
        self.ptr = ptr
        self.size = size

    def __del__(self):
        global _pin_cached_bytes
        try:
            with _pin_lock:
                if _pin_cached_bytes + self.size <= _PIN_CACHE_LIMIT:
                    _pin_free.setdefault(self.size, []).append(self.ptr)
                    _pin_cached_bytes += self.size
                    return
            _lib.trgl_host_free(self.ptr)
        except Exception:
            pass


def pinned_empty(shape, dtype):
    """NumPy array over page-locked host memory (full-speed asynchronous PCIe copies in host mode)."""
    global _pin_cached_bytes
    dtype = np.dtype(dtype)
    n = int(np.prod(shape, dtype=np.int64))
    nbytes = max(n * dtype.itemsize, 1)
    size = (nbytes + _PIN_GRAIN - 1) // _PIN_GRAIN * _PIN_GRAIN
    ptr = None
    with _pin_lock:
        lst = _pin_free.get(size)
        if lst:
            ptr = lst.pop()
            _pin_cached_bytes -= size
    if ptr is None:
        p = ctypes.c_void_p()
        check(lib().trgl_host_alloc(ctypes.byref(p), size))
        ptr = p.value
    buf = (ctypes.c_char * nbytes).from_address(ptr)
    buf._trgl_owner = _PinnedOwner(ptr, size)       # arr.base -> buf -> owner: recycled with the last view
    return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)


def pinned_cache_clear():
    global _pin_cached_bytes
    with _pin_lock:
        for lst in _pin_free.values():
            for ptr in lst:
                _lib.trgl_host_free(ptr)
        _pin_free.clear()
        _pin_cached_bytes = 0


def pinned_copy(a):
    out = pinned_empty(a.shape, a.dtype)
    out[...] = a
    return out


_FLOAT_DTYPES = (np.dtype(np.float32), np.dtype(np.float64))


def _is_device(a):
    return hasattr(a, "data_ptr")


def _host_if_cpu_tensor(a):
    """A CPU torch tensor is a HOST array (its data_ptr() is a host address): hand NumPy its memory instead."""
    if hasattr(a, "data_ptr") and hasattr(a, "is_cuda") and not a.is_cuda:
        return a.detach().numpy()
    return a


def _np_dtype(a):
    return np.dtype(str(a.dtype).replace("torch.", ""))


def _check_device_array(a, cols, name, dtypes=_FLOAT_DTYPES):
    """Device buffers are read as packed row-major (n, cols) arrays: reject anything a kernel would misread."""
    shape = tuple(int(v) for v in a.shape)
    if cols:
        if len(shape) != 2 or shape[1] != cols:
            raise ValueError("%s: device buffer must have shape (n, %d), got %r" % (name, cols, shape))
    elif len(shape) != 1:
        raise ValueError("%s: device buffer must have shape (n,), got %r" % (name, shape))
    if hasattr(a, "is_contiguous") and not a.is_contiguous():
        raise ValueError("%s: device tensor must be contiguous (call .contiguous()); a strided view would be read as packed" % name)
    if dtypes is not None and _np_dtype(a) not in dtypes:
        raise ValueError("%s: unsupported dtype %s" % (name, a.dtype))


def _ptr(a):
    """Address of a host array / device buffer as a plain integer (None -> NULL); every pointer parameter is a c_void_p."""
    if a is None:
        return None
    if _is_device(a):
        return a.data_ptr() or None
    return a.ctypes.data or None


# Device buffers of the resident pairs come from a small size-keyed pool (cudaMalloc + cudaFree of two 160 MB blocks per
# handle cost 2-4 ms, as much as one solver call on 10 M points).
_DEV_POOL_LIMIT = 8 << 30
_dev_pool = {}          # nbytes -> [ptr, ...]
_dev_pool_bytes = 0


class _PooledDeviceArray(DeviceArray):
    def __init__(self, shape, dtype):
        global _dev_pool_bytes
        self.shape = tuple(int(v) for v in shape)
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        with _pin_lock:
            lst = _dev_pool.get(self.nbytes)
            ptr = lst.pop() if lst else None
            if ptr is not None:
                _dev_pool_bytes -= self.nbytes
        if ptr is None:
            p = ctypes.c_void_p()
            check(lib().trgl_device_alloc(ctypes.byref(p), self.nbytes))
            ptr = p.value or 0
        self.ptr = ptr

    def __del__(self):
        global _dev_pool_bytes
        try:
            if getattr(self, "ptr", 0):
                with _pin_lock:
                    if _dev_pool_bytes + self.nbytes <= _DEV_POOL_LIMIT:
                        _dev_pool.setdefault(self.nbytes, []).append(self.ptr)
                        _dev_pool_bytes += self.nbytes
                        self.ptr = 0
                        return
                _lib.trgl_device_free(self.ptr)
            self.ptr = 0
        except Exception:
            pass


def device_pool_clear():
    global _dev_pool_bytes
    with _pin_lock:
        for lst in _dev_pool.values():
            for ptr in lst:
                _lib.trgl_device_free(ptr)
        _dev_pool.clear()
        _dev_pool_bytes = 0


class _ResidentPair:
    def __init__(self, u1, u2):
        u1 = np.asarray(_host_if_cpu_tensor(u1)); u2 = np.asarray(_host_if_cpu_tensor(u2))
        if u1.dtype != np.float32 or u2.dtype != np.float32:
            u1 = u1.astype(np.float64, copy=False); u2 = u2.astype(np.float64, copy=False)
        self.host = (np.ascontiguousarray(u1.reshape(-1, 2)), np.ascontiguousarray(u2.reshape(-1, 2)))
        if len(self.host[0]) != len(self.host[1]):
            raise ValueError("u1 and u2 must hold the same number of points")
        self.dev = (_PooledDeviceArray(self.host[0].shape, self.host[0].dtype), _PooledDeviceArray(self.host[1].shape, self.host[1].dtype))
        self.uploaded = False


class ResidentPoints:
    """
    One of the two observation arrays of an "upload once, solve many" pair (see `resident`).  Passed as u1 / u2 to any
    solver, it behaves like the host array it was made from -- results come back as host arrays -- but the first solver
    call leaves the uploaded observations in HBM (trgl_set_input_retention) and every later call reads them there
    (TRGL_MEM_DEVICE_IN) instead of uploading them again.
    """

    def __init__(self, pair, which):
        self._pair, self._which = pair, which
        self.shape, self.dtype = pair.host[which].shape, pair.host[which].dtype

    def __len__(self):
        return self.shape[0]

    def __array__(self, dtype=None, copy=None):
        a = self._pair.host[self._which]
        return a if dtype is None else a.astype(dtype, copy=False)

    @property
    def device(self):
        """The device-resident copy (valid after the first solver call)."""
        return self._pair.dev[self._which]


def resident(u1, u2):
    """(u1, u2) host arrays -> a pair of ResidentPoints handles for the four solvers' u1 / u2 arguments."""
    pair = _ResidentPair(u1, u2)
    return ResidentPoints(pair, 0), ResidentPoints(pair, 1)


def _P12(P):
    """Rows 0..2 of a 3x4 / 4x4 camera matrix as 12 contiguous doubles (triangulation.c:24-25 reads P[4*k+l])."""
    P = np.asarray(P)
    if P.ndim != 2 or P.shape[1] != 4 or P.shape[0] not in (3, 4):
        raise ValueError("camera matrix must be 3x4 or 4x4, got %r" % (P.shape,))
    return np.ascontiguousarray(P[0:3, :], dtype=np.float64)


def _dp(a):
    return a.ctypes.data


_MODE = {(8, 8, 8): F64, (4, 8, 4): F32IO, (4, 4, 4): F32, (8, 8, 4): F64_OUT32, (4, 8, 8): F32_OUT64}


def mode_for(in_dtype, compute_dtype, out_dtype):
    key = (np.dtype(in_dtype).itemsize, np.dtype(compute_dtype).itemsize, np.dtype(out_dtype).itemsize)
    if key not in _MODE:
        raise ValueError("unsupported precision combination in=%s compute=%s out=%s" % (in_dtype, compute_dtype, out_dtype))
    return _MODE[key]


def _prep(u1, u2, compute_dtype, out_dtype, allow_resident=True):
    """Input coercion of the reference wrapper (triangulation_c/__init__.py:32-39), without the forced up-cast:
    float32 inputs stay float32 in HBM and are widened in registers.
    Returns u1, u2, mem (MEM_HOST / MEM_DEVICE / MEM_DEVICE_IN), n, precision mode, resident pair to commit (or None)."""
    if type(u1) is np.ndarray and type(u2) is np.ndarray and u1.dtype == u2.dtype and u1.ndim == 2 and u2.ndim == 2 \
            and u1.shape[1] == 2 and u1.shape == u2.shape and u1.flags.c_contiguous and u2.flags.c_contiguous \
            and u1.dtype in _FLOAT_DTYPES:
        # fast path of the small-batch (SLAM keyframe) calls: nothing to coerce
        in_dtype = u1.dtype
        if np.dtype(compute_dtype) == np.float32 and (in_dtype != np.float32 or np.dtype(out_dtype) != np.float32):
            compute_dtype = np.float64
        return u1, u2, MEM_HOST, len(u1), mode_for(in_dtype, compute_dtype, out_dtype), None
    commit = None
    mem = None
    if isinstance(u1, ResidentPoints) or isinstance(u2, ResidentPoints):
        if not (isinstance(u1, ResidentPoints) and isinstance(u2, ResidentPoints) and u1._pair is u2._pair
                and (u1._which, u2._which) == (0, 1)):
            raise ValueError("resident handles must be passed as the (u1, u2) pair `resident` returned")
        pair = u1._pair
        if pair.uploaded:
            u1, u2 = pair.dev
            mem = MEM_DEVICE_IN if allow_resident else MEM_DEVICE
        else:
            u1, u2 = pair.host
            mem = MEM_HOST
            if allow_resident:
                commit = pair               # the caller arms the retention right before its library call (_retain)
    else:
        u1 = _host_if_cpu_tensor(u1); u2 = _host_if_cpu_tensor(u2)
        dev = _is_device(u1)
        if dev != _is_device(u2):
            raise ValueError("u1 and u2 must both be host arrays or both be device buffers")
        mem = MEM_DEVICE if dev else MEM_HOST
        if not dev:
            u1 = np.asarray(u1); u2 = np.asarray(u2)
            if u1.dtype != np.float32 or u2.dtype != np.float32:
                u1 = u1.astype(np.float64, copy=False); u2 = u2.astype(np.float64, copy=False)
            u1 = np.ascontiguousarray(u1.reshape(-1, 2)); u2 = np.ascontiguousarray(u2.reshape(-1, 2))
        else:
            _check_device_array(u1, 2, "u1"); _check_device_array(u2, 2, "u2")
    if len(u1) != len(u2):
        raise ValueError("u1 and u2 must hold the same number of points")
    in_dtype = _np_dtype(u1)
    if in_dtype != _np_dtype(u2):
        raise ValueError("u1 and u2 must have the same dtype")
    if np.dtype(compute_dtype) == np.float32 and (in_dtype != np.float32 or np.dtype(out_dtype) != np.float32):
        compute_dtype = np.float64          # FP32 arithmetic is only defined for float32 in/out
    return u1, u2, mem, len(u1), mode_for(in_dtype, compute_dtype, out_dtype), commit


def _out(dev, n, cols, dtype, given):
    if given is not None:
        return given
    shape = (n, cols) if cols else (n,)
    if dev:
        return DeviceArray(shape, dtype)
    if n >= PINNED_OUTPUT_MIN_POINTS:
        return pinned_empty(shape, dtype)
    return np.empty(shape, dtype=dtype)


class Intrinsics:
    """(K1, dist1, K2, dist2) of the two views for the pixel-input (`*_px`) entry points; dist may be None, 4 or 5
    coefficients (k1,k2,p1,p2[,k3]).  One camera for both views: Intrinsics(K, dist)."""

    def __init__(self, K1, dist1=None, K2=None, dist2=None):
        self.K1 = _K9(K1); self.d1 = _dist5(dist1)
        self.K2 = self.K1 if K2 is None else _K9(K2)
        self.d2 = self.d1 if K2 is None and dist2 is None else _dist5(dist2)

    def args(self):
        return (_dp(self.K1), None if self.d1 is None else _dp(self.d1),
                _dp(self.K2), None if self.d2 is None else _dp(self.d2))


def _K9(K):
    K = np.ascontiguousarray(K, dtype=np.float64)
    if K.shape != (3, 3):
        raise ValueError("camera matrix K must be 3x3, got %r" % (K.shape,))
    return K


def _dist5(dist):
    if dist is None:
        return None
    dd = np.asarray(dist, dtype=np.float64).ravel()
    if len(dd) not in (4, 5):
        raise ValueError("distortion model must be (k1,k2,p1,p2[,k3]); got %d coefficients" % len(dd))
    d = np.zeros(5)
    d[:len(dd)] = dd
    return d


def undistort_points(src, K, dist=None, dst=None, stream=None):
    """cv2.undistortPoints(src, K, dist) -> (N,2) normalised coordinates in the dtype of src (float32 / float64);
    any (…,2) input shape is accepted (the reference passes (1,N,2), slam2.py:551-552)."""
    dev = _is_device(src)
    if not dev:
        src = np.asarray(src)
        if src.dtype != np.float32:
            src = src.astype(np.float64, copy=False)
        src = np.ascontiguousarray(src.reshape(-1, 2))
    n = len(src)
    dt = np.dtype(str(src.dtype).replace("torch.", ""))
    K = _K9(K); d = _dist5(dist)
    dst = _out(dev, n, 2, dt, dst)
    check(lib().trgl_undistort_points(_ptr(src), _ptr(dst), _dp(K), None if d is None else _dp(d), n,
                                      int(dt == np.float32), MEM_DEVICE if dev else MEM_HOST, stream))
    return dst


def linear_ls(u1, P1, u2, P2, out_dtype=np.float64, compute_dtype=np.float64, x=None, status=None, stream=None,
              pixel=None, evaluate=None):
    """pixel: an Intrinsics -> u1,u2 are pixel coordinates, undistorted in registers in front of the solve."""
    u1, u2, mem, n, mode, commit = _prep(u1, u2, compute_dtype, out_dtype)
    dev = mem == MEM_DEVICE
    P1 = _P12(P1); P2 = _P12(P2)
    x = _out(dev, n, 3, out_dtype, x); status = _out(dev, n, 0, np.bool_, status)
    _arm(evaluate, dev)
    _retain(commit)
    if pixel is None:
        check(lib().trgl_linear_ls(_ptr(u1), _ptr(u2), _dp(P1), _dp(P2), _ptr(x), _ptr(status), n, mode, mem, stream))
    else:
        check(lib().trgl_linear_ls_px(_ptr(u1), _ptr(u2), *pixel.args(), _dp(P1), _dp(P2), _ptr(x), _ptr(status), n,
                                      mode, mem, stream))
    _commit(commit)
    return x, status


def multiview_ls(us, Ps, valid=None, min_views=2, out_dtype=np.float64, x=None, status=None, stream=None):
    """us (m,n,2) host array or DeviceArray, Ps (m,3|4,4), valid (m,n) bool / uint8 or None -> x (n,3), status (n,) bool."""
    us = _host_if_cpu_tensor(us); valid = _host_if_cpu_tensor(valid)
    dev = _is_device(us)
    if not dev:
        us = np.asarray(us)
        if us.dtype != np.float32:
            us = us.astype(np.float64, copy=False)
        us = np.ascontiguousarray(us)
    if len(us.shape) != 3 or us.shape[2] != 2:
        raise ValueError("us must have shape (m, n, 2)")
    if dev:
        if hasattr(us, "is_contiguous") and not us.is_contiguous():
            raise ValueError("us: device tensor must be contiguous")
        if _np_dtype(us) not in _FLOAT_DTYPES:
            raise ValueError("us: unsupported dtype %s" % us.dtype)
    m, n = int(us.shape[0]), int(us.shape[1])
    in_dtype = np.dtype(str(us.dtype).replace("torch.", ""))
    Pm = np.ascontiguousarray(np.stack([_P12(P) for P in Ps]).reshape(-1))
    if len(Pm) != 12 * m:
        raise ValueError("need one camera matrix per view")
    if valid is not None:
        if dev != _is_device(valid):
            raise ValueError("us and valid must both be host arrays or both be device buffers")
        if not dev:
            valid = np.ascontiguousarray(np.asarray(valid).astype(np.uint8, copy=False))
        if tuple(valid.shape) != (m, n):
            raise ValueError("valid must have shape (m, n)")
        if dev and ((hasattr(valid, "is_contiguous") and not valid.is_contiguous()) or _np_dtype(valid).itemsize != 1):
            raise ValueError("valid: device mask must be a contiguous (m, n) array of 1-byte elements")
    x = _out(dev, n, 3, out_dtype, x); status = _out(dev, n, 0, np.bool_, status)
    check(lib().trgl_multiview_ls(_ptr(us), _ptr(valid) if valid is not None else None, _dp(Pm), m, _ptr(x), _ptr(status),
                                  n, int(min_views), mode_for(in_dtype, np.float64, out_dtype),
                                  MEM_DEVICE if dev else MEM_HOST, stream))
    return x, status


def iterative_ls(u1, P1, u2, P2, tolerance=3.e-5, semantics=ITER_C, out_dtype=np.float64, compute_dtype=np.float64,
                 x=None, status=None, stream=None, pixel=None, evaluate=None):
    u1, u2, mem, n, mode, commit = _prep(u1, u2, compute_dtype, out_dtype)
    dev = mem == MEM_DEVICE
    P1 = _P12(P1); P2 = _P12(P2)
    x = _out(dev, n, 3, out_dtype, x); status = _out(dev, n, 0, np.int32, status)
    _arm(evaluate, dev)
    _retain(commit)
    if pixel is None:
        check(lib().trgl_iterative_ls(_ptr(u1), _ptr(u2), _dp(P1), _dp(P2), _ptr(x), _ptr(status), n, float(tolerance),
                                      semantics, mode, mem, stream))
    else:
        check(lib().trgl_iterative_ls_px(_ptr(u1), _ptr(u2), *pixel.args(), _dp(P1), _dp(P2), _ptr(x), _ptr(status), n,
                                         float(tolerance), semantics, mode, mem, stream))
    _commit(commit)
    return x, status


def linear_eigen(u1, P1, u2, P2, max_coordinate_value=1.e16, rows=4, out_dtype=np.float64, compute_dtype=np.float64,
                 x=None, status=None, stream=None, pixel=None, evaluate=None):
    u1, u2, mem, n, mode, commit = _prep(u1, u2, compute_dtype, out_dtype)
    dev = mem == MEM_DEVICE
    P1 = _P12(P1); P2 = _P12(P2)
    x = _out(dev, n, 3, out_dtype, x); status = _out(dev, n, 0, np.bool_, status)
    _arm(evaluate, dev)
    _retain(commit)
    if pixel is None:
        check(lib().trgl_linear_eigen(_ptr(u1), _ptr(u2), _dp(P1), _dp(P2), _ptr(x), _ptr(status), n,
                                      float(max_coordinate_value), rows, mode, mem, stream))
    else:
        check(lib().trgl_linear_eigen_px(_ptr(u1), _ptr(u2), *pixel.args(), _dp(P1), _dp(P2), _ptr(x), _ptr(status), n,
                                         float(max_coordinate_value), rows, mode, mem, stream))
    _commit(commit)
    return x, status


def polynomial(u1, P1, u2, P2, F=None, max_coordinate_value=1.e16, rows=4, out_dtype=np.float64,
               compute_dtype=np.float64, x=None, status=None, want_corrected=False, check_all_nan=True, stream=None,
               pixel=None, evaluate=None):
    """Returns x, status, all_nan[, u1_corr, u2_corr]."""
    u1, u2, mem, n, mode, commit = _prep(u1, u2, compute_dtype, out_dtype)
    dev = mem == MEM_DEVICE
    P1 = _P12(P1); P2 = _P12(P2)
    x = _out(dev, n, 3, out_dtype, x); status = _out(dev, n, 0, np.bool_, status)
    in_dtype = np.float32 if mode in (F32IO, F32, F32_OUT64) else np.float64
    c1 = _out(dev, n, 2, in_dtype, None) if want_corrected else None
    c2 = _out(dev, n, 2, in_dtype, None) if want_corrected else None
    flag = ctypes.c_int(0)
    flag_p = ctypes.byref(flag) if check_all_nan else None
    _arm(evaluate, dev)
    _retain(commit)
    if pixel is not None:
        if F is not None:
            raise ValueError("pixel inputs and an explicit F cannot be combined")
        check(lib().trgl_polynomial_px(_ptr(u1), _ptr(u2), *pixel.args(), _dp(P1), _dp(P2), _ptr(x), _ptr(status),
                                       _ptr(c1), _ptr(c2), flag_p, n, float(max_coordinate_value), rows, mode, mem,
                                       stream))
    elif F is None:
        check(lib().trgl_polynomial(_ptr(u1), _ptr(u2), _dp(P1), _dp(P2), _ptr(x), _ptr(status), _ptr(c1), _ptr(c2),
                                    flag_p, n, float(max_coordinate_value), rows, mode, mem, stream))
    else:
        F = np.ascontiguousarray(F, dtype=np.float64).reshape(3, 3)
        check(lib().trgl_polynomial_F(_ptr(u1), _ptr(u2), _dp(P1), _dp(P2), _dp(F), _ptr(x), _ptr(status), _ptr(c1),
                                      _ptr(c2), flag_p, n, float(max_coordinate_value), rows, mode, mem, stream))
    _commit(commit)
    if want_corrected:
        return x, status, bool(flag.value), c1, c2
    return x, status, bool(flag.value)


def polynomial_flags_async(host_flags2, stream=None):
    """Enqueue the copy of the last device-mode polynomial call's two not-all-NaN words into a pinned (2,) uint32 array."""
    check(lib().trgl_polynomial_flags_async(host_flags2.ctypes.data, stream))


def fundamental_8point(u1, u2, compute_dtype=np.float64, stream=None):
    u1, u2, mem, n, mode, _ = _prep(u1, u2, compute_dtype, np.float32 if np.dtype(compute_dtype) == np.float32 else
                                    (np.float32 if getattr(u1, "dtype", None) == np.float32 else np.float64),
                                    allow_resident=False)
    F = np.zeros((3, 3))
    check(lib().trgl_fundamental_8point(_ptr(u1), _ptr(u2), n, mode, mem, _dp(F), stream))
    return F


def reproj_error(x, imgp, K, dist, rvec, tvec, want_proj=True, stream=None):
    """Returns (sum dx^2, sum dy^2, finite count, sum|dx|, sum|dy|), proj or None."""
    dev = _is_device(x)
    if not dev:
        x = np.asarray(x); imgp = np.asarray(imgp)
        if x.dtype != np.float32:
            x = x.astype(np.float64, copy=False)
        if imgp.dtype != np.float32:
            imgp = imgp.astype(np.float64, copy=False)
        x = np.ascontiguousarray(x.reshape(-1, 3)); imgp = np.ascontiguousarray(imgp.reshape(-1, 2))
    n = len(x)
    if len(imgp) != n:
        raise ValueError("objp and imgp must hold the same number of points")
    x32 = int(np.dtype(str(x.dtype).replace("torch.", "")) == np.float32)
    i32 = int(np.dtype(str(imgp.dtype).replace("torch.", "")) == np.float32)
    K = np.ascontiguousarray(K, dtype=np.float64).reshape(3, 3)
    d = np.zeros(5)
    if dist is not None:
        dd = np.asarray(dist, dtype=np.float64).ravel()
        d[:min(5, len(dd))] = dd[:5]
    rvec = np.ascontiguousarray(rvec, dtype=np.float64).ravel(); tvec = np.ascontiguousarray(tvec, dtype=np.float64).ravel()
    proj = _out(dev, n, 2, np.float32 if i32 else np.float64, None) if want_proj else None
    sums = np.zeros(3); abs_sums = np.zeros(2)
    check(lib().trgl_reproj_error(_ptr(x), _ptr(imgp), _dp(K), _dp(d), _dp(rvec), _dp(tvec), _ptr(proj), _dp(sums),
                                  _dp(abs_sums), n, x32, i32, MEM_DEVICE if dev else MEM_HOST, stream))
    return np.concatenate([sums, abs_sums]), proj


def pair_reproj(x, u1, P1, u2, P2, status, min_status=0, max_sq_err=np.inf, want_errors=True, want_good=True,
                stream=None, sums_device=None):
    """Two-view reprojection errors + good mask right after a solver call. Returns err1, err2, good, sums(4).
    sums_device (a 4-double device buffer, device inputs only): asynchronous variant -- the sums are finished inside
    the kernel and stay on the device, the call does not synchronise and returns sums_device in place of the host sums."""
    x, u1, u2, status = (_host_if_cpu_tensor(a) for a in (x, u1, u2, status))
    dev = _is_device(x)
    if dev:
        if not (_is_device(u1) and _is_device(u2) and _is_device(status)):
            raise ValueError("x, u1, u2 and status must all be host arrays or all be device buffers")
        _check_device_array(x, 3, "x"); _check_device_array(u1, 2, "u1"); _check_device_array(u2, 2, "u2")
        _check_device_array(status, 0, "status", None)
        if not (len(x) == len(u1) == len(u2) == len(status)):
            raise ValueError("x, u1, u2 and status must hold the same number of points")
        if _np_dtype(status).itemsize not in (1, 4):
            raise ValueError("status: device buffer must be bool / uint8 or int32")
    if not dev:
        x = np.asarray(x); u1 = np.asarray(u1); u2 = np.asarray(u2); status = np.asarray(status)
        if x.dtype != np.float32:
            x = x.astype(np.float64, copy=False)
        if u1.dtype != np.float32 or u2.dtype != np.float32:
            u1 = u1.astype(np.float64, copy=False); u2 = u2.astype(np.float64, copy=False)
        x = np.ascontiguousarray(x.reshape(-1, 3))
        u1 = np.ascontiguousarray(u1.reshape(-1, 2)); u2 = np.ascontiguousarray(u2.reshape(-1, 2))
        if status.dtype == np.bool_:
            status = np.ascontiguousarray(status).view(np.uint8)
        elif status.dtype != np.uint8:
            status = np.ascontiguousarray(status, dtype=np.int32)
    n = len(x)
    xdt = np.dtype(str(x.dtype).replace("torch.", "")); udt = np.dtype(str(u1.dtype).replace("torch.", ""))
    sdt = np.dtype(str(status.dtype).replace("torch.", ""))
    mode = mode_for(udt, np.float64, xdt)
    P1 = _P12(P1); P2 = _P12(P2)
    e1 = _out(dev, n, 0, xdt, None) if want_errors else None
    e2 = _out(dev, n, 0, xdt, None) if want_errors else None
    # want_good may be a caller-owned (n,) bool / uint8 buffer to fill
    good = want_good if (want_good is not None and not isinstance(want_good, bool)) else \
        (_out(dev, n, 0, np.bool_, None) if want_good else None)
    if sums_device is not None:
        if not dev:
            raise ValueError("the asynchronous variant needs device buffers")
        check(lib().trgl_pair_reproj_async(_ptr(x), _ptr(u1), _ptr(u2), _dp(P1), _dp(P2), _ptr(status),
                                           int(sdt.itemsize == 4), int(min_status), float(max_sq_err), _ptr(e1), _ptr(e2),
                                           _ptr(good), _ptr(sums_device), n, mode, stream))
        return e1, e2, good, sums_device
    sums = np.zeros(4)
    check(lib().trgl_pair_reproj(_ptr(x), _ptr(u1), _ptr(u2), _dp(P1), _dp(P2), _ptr(status), int(sdt.itemsize == 4),
                                 int(min_status), float(max_sq_err), _ptr(e1), _ptr(e2), _ptr(good), _dp(sums), n,
                                 mode, MEM_DEVICE if dev else MEM_HOST, stream))
    return e1, e2, good, sums


def eval_errors_3d(x, exact, status=None, thresh_max=1.0, thresh_min=1.0, want_errors=True, stream=None):
    """Squared 3-D errors against the exact cloud + the sums error_rms / robustness_stat need.
    Returns errors (or None), stats = (sum, #NaN, #false positives, #false negatives)."""
    dev = _is_device(x)
    if not dev:
        x = np.asarray(x)
        if x.dtype != np.float32:
            x = x.astype(np.float64, copy=False)
        x = np.ascontiguousarray(x.reshape(-1, 3))
        exact = np.ascontiguousarray(exact, dtype=np.float64)
        if status is not None:
            status = np.asarray(status)
            status = np.ascontiguousarray(status).view(np.uint8) if status.dtype == np.bool_ else \
                np.ascontiguousarray(status, dtype=np.uint8 if status.dtype == np.uint8 else np.int32)
    n = len(x)
    stride = int(exact.shape[1])
    x32 = int(np.dtype(str(x.dtype).replace("torch.", "")) == np.float32)
    s_i32 = int(status is not None and np.dtype(str(status.dtype).replace("torch.", "")).itemsize == 4)
    errors = _out(dev, n, 0, np.float64, None) if want_errors else None
    stats = np.zeros(4)
    check(lib().trgl_eval_errors_3d(_ptr(x), _ptr(exact), stride, _ptr(status), s_i32, float(thresh_max), float(thresh_min),
                                    _ptr(errors), _dp(stats), n, x32, MEM_DEVICE if dev else MEM_HOST, stream))
    return errors, stats


def eval_errors_2d(proj, exact, want_errors=True, stream=None):
    """Squared 2-D errors |proj - exact|^2 and their sum.  Returns errors (or None), stats = (sum, #NaN, 0, 0)."""
    dev = _is_device(proj)
    if not dev:
        proj = np.asarray(proj)
        if proj.dtype != np.float32:
            proj = proj.astype(np.float64, copy=False)
        proj = np.ascontiguousarray(proj.reshape(-1, 2))
        exact = np.ascontiguousarray(np.asarray(exact, dtype=np.float64).reshape(-1, 2))
    n = len(proj)
    p32 = int(np.dtype(str(proj.dtype).replace("torch.", "")) == np.float32)
    errors = _out(dev, n, 0, np.float64, None) if want_errors else None
    stats = np.zeros(4)
    check(lib().trgl_eval_errors_2d(_ptr(proj), _ptr(exact), _ptr(errors), _dp(stats), n, p32,
                                    MEM_DEVICE if dev else MEM_HOST, stream))
    return errors, stats


def vector_stat(x_trials, exact, stream=None):
    """vector_stat (triangulation_comparison.py:219-240) of the error vectors x_trials[t] - exact[:, 0:3]:
    x_trials (trials, n, 3) host array or device buffer, exact (n, 3|4) in the same memory space.
    Returns means (n,3), covars (n,3,3) -- host arrays for host input, DeviceArrays for device input."""
    x_trials = _host_if_cpu_tensor(x_trials); exact = _host_if_cpu_tensor(exact)
    dev = _is_device(x_trials)
    if dev != _is_device(exact):
        raise ValueError("x_trials and exact must both be host arrays or both be device buffers")
    if not dev:
        x_trials = np.asarray(x_trials)
        if x_trials.dtype != np.float32:
            x_trials = x_trials.astype(np.float64, copy=False)
        x_trials = np.ascontiguousarray(x_trials)
        exact = np.ascontiguousarray(exact, dtype=np.float64)
    shape = tuple(int(v) for v in x_trials.shape)
    if len(shape) != 3 or shape[2] != 3 or len(exact.shape) != 2 or int(exact.shape[0]) != shape[1] or int(exact.shape[1]) < 3:
        raise ValueError("x_trials must be (trials, n, 3) and exact (n, >= 3)")
    if dev and ((hasattr(x_trials, "is_contiguous") and not x_trials.is_contiguous()) or _np_dtype(exact) != np.float64):
        raise ValueError("device x_trials must be contiguous and exact float64")
    trials, n = shape[0], shape[1]
    means = _out(dev, n, 3, np.float64, None)
    covars = DeviceArray((n, 3, 3), np.float64) if dev else np.empty((n, 3, 3))
    check(lib().trgl_vector_stat(_ptr(x_trials), _ptr(exact), int(exact.shape[1]), trials, _ptr(means), _ptr(covars), n,
                                 int(_np_dtype(x_trials) == np.float32), MEM_DEVICE if dev else MEM_HOST, stream))
    return means, covars


def median(values, stream=None):
    """np.median of non-negative float64 values (host array or device buffer), exact."""
    dev = _is_device(values)
    if not dev:
        values = np.ascontiguousarray(values, dtype=np.float64).ravel()
    n = int(np.prod(values.shape))
    out = np.zeros(1)
    check(lib().trgl_median(_ptr(values), n, MEM_DEVICE if dev else MEM_HOST, _dp(out), stream))
    return float(out[0])


class FusedEval:
    """Two-view evaluation (reprojection errors, good mask, sums) fused into the NEXT device-mode solver call of this
    thread: pass an instance as `evaluate=` to linear_ls / iterative_ls / linear_eigen / polynomial.  Outputs are device
    buffers: .sums (4,) = sum err1, sum err2 over good points, #good, #status > min_status; .good (n,) bool;
    .err1 / .err2 (n,) when want_errors."""

    def __init__(self, n, x_dtype=np.float64, min_status=0, max_sq_err=np.inf, want_errors=False, want_good=True, sums=None):
        self.min_status, self.max_sq_err = int(min_status), float(max_sq_err)
        self.sums = sums if sums is not None else DeviceArray((4,), np.float64)
        self.good = DeviceArray((n,), np.bool_) if want_good else None
        self.err1 = DeviceArray((n,), x_dtype) if want_errors else None
        self.err2 = DeviceArray((n,), x_dtype) if want_errors else None

    def arm(self):
        check(lib().trgl_set_fused_eval(self.min_status, self.max_sq_err, _ptr(self.err1), _ptr(self.err2), _ptr(self.good),
                                        _ptr(self.sums)))


def _retain(pair):
    """Ask the next host-mode solver call of this thread to leave its uploaded u1 / u2 in the pair's device buffers."""
    if pair is not None:
        check(lib().trgl_set_input_retention(_ptr(pair.dev[0]), _ptr(pair.dev[1])))


def _commit(pair):
    """The host-mode call that was asked to retain its inputs has succeeded: later calls read them in HBM."""
    if pair is not None:
        pair.uploaded = True


def _arm(evaluate, dev):
    if evaluate is not None:
        if not dev:
            raise ValueError("the fused evaluation needs device-resident inputs")
        evaluate.arm()


def set_result_mirrors(mirrors, x_f32=False):
    """mirrors: list of (x_address, status_address) in peer GPUs' memory for the NEXT device-mode solver call of this
    thread (see include/triangl_cuda.h, "result gather fused into the solver's stores").  x_f32: the mirrors hold float32
    rows whatever the dtype of the call's own x (trgl_set_result_mirrors_f32)."""
    k = len(mirrors)
    xs = (ctypes.c_void_p * max(k, 1))(*[int(m[0]) for m in mirrors])
    ss = (ctypes.c_void_p * max(k, 1))(*[int(m[1]) for m in mirrors])
    check((lib().trgl_set_result_mirrors_f32 if x_f32 else lib().trgl_set_result_mirrors)(xs, ss, k))


def ipc_export(dev):
    """64-byte CUDA IPC handle of a DeviceArray (or raw device address from trgl_device_alloc)."""
    h = ctypes.create_string_buffer(64)
    check(lib().trgl_ipc_export(ctypes.c_void_p(dev.data_ptr() if hasattr(dev, "data_ptr") else int(dev)), h))
    return h.raw


def ipc_import(handle):
    """Map a peer process's exported buffer; returns the device address valid in this process."""
    p = ctypes.c_void_p()
    check(lib().trgl_ipc_import(ctypes.create_string_buffer(handle, 64), ctypes.byref(p)))
    return p.value


def ipc_close(ptr):
    check(lib().trgl_ipc_close(ctypes.c_void_p(ptr)))


def launch_count():
    return int(lib().trgl_launch_count())


def set_points_per_thread(ppt):
    return lib().trgl_set_points_per_thread(int(ppt))


def set_stream_variant(v):
    return lib().trgl_set_stream_variant(int(v))


def set_deferred_capacity(max_points):
    """Test knob: limit of the deferred-point list of the hot kernels (default 2**26); returns the old limit."""
    return lib().trgl_set_deferred_capacity(int(max_points))


def set_trace(enabled):
    return lib().trgl_set_trace(int(enabled))


def get_trace():
    """Per-phase host microseconds of the small-batch path since the last call: dict(stage_in, launch, sync, stage_out, calls)."""
    out = (ctypes.c_double * 5)()
    check(lib().trgl_get_trace(out))
    return dict(zip(("stage_in_us", "launch_us", "sync_us", "stage_out_us", "calls"), [float(v) for v in out]))


def rare_path_counters(reset=False):
    """dict(isolated, intervals, not_certified, durand_kerner, max_intervals): points through the rare paths of polynomial's correction."""
    out = (ctypes.c_ulonglong * 5)()
    check(lib().trgl_rare_path_counters(out, 1 if reset else 0))
    return dict(zip(("isolated", "intervals", "not_certified", "durand_kerner", "max_intervals"), [int(v) for v in out]))


def fp64_fma_rate(operands=2, chains=8, ctas_per_sm=4):
    """Measured FP64 FMA rate of the current device in warp instructions per second (x 64 = flop/s): `chains` independent
    chains per thread, two (one source from the constant bank) or three distinct register sources per instruction."""
    v = ctypes.c_double(0.0)
    check(lib().trgl_fp64_fma_rate(int(operands), int(chains), int(ctas_per_sm), ctypes.byref(v)))
    return float(v.value)


def deferred_total(stream=None):
    """Running total of the points the hot kernels have deferred to their follow-up kernels on this device / stream."""
    v = ctypes.c_int64(0)
    check(lib().trgl_deferred_total(stream, ctypes.byref(v)))
    return int(v.value)


def set_two_ray(enabled):
    """1 = two-ray closed forms where certified (iterative_LS, polynomial; default), 0 = the reference's arithmetic for
    every point; returns the old value."""
    return lib().trgl_set_two_ray(int(enabled))


def synchronize():
    check(lib().trgl_device_synchronize())


class Event:
    def __init__(self):
        p = ctypes.c_void_p()
        check(lib().trgl_event_create(ctypes.byref(p)))
        self.ptr = p

    def record(self, stream=None):
        check(lib().trgl_event_record(self.ptr, stream))

    def synchronize(self):
        check(lib().trgl_event_synchronize(self.ptr))

    def elapsed_ms(self, stop):
        ms = ctypes.c_float(0)
        check(lib().trgl_event_elapsed_ms(self.ptr, stop.ptr, ctypes.byref(ms)))
        return float(ms.value)
