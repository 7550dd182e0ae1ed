"""
The drop-in boundary binds: the `ctypes` stub INTEGRATION.md section 2 tells a maintainer to install as
Work/python_libs/triangulation_c/__init__.py is extracted from the document, exec'd, and
  * (CPU, this container) placed under the reference's own override shim -- Work/python_libs/triangulation.py:236-253,
    read from /root/reference at test time, never copied -- to show the shim picks the two wrappers up and that a call
    travels through ctypes into the C ABI (without a GPU it comes back as the library's "no CUDA device" error, which
    only libtriangl_cuda.so can produce);
  * (GPU) called directly and compared with the oracle.
"""
import os
import re
import sys
import types

import numpy as np
import pytest

import synthetic_rig as rig
from oracle import triangulation_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TRIANGULATION = "/root/reference/Work/python_libs/triangulation.py"


def _stub_module():
    """The first python block of INTEGRATION.md section 2, as a module named `triangulation_c`."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    sec = text[text.index("## 2. Replacing the weave extension"):]
    code = re.search(r"```python\n(.*?)```", sec, re.S).group(1)
    assert "Work/python_libs/triangulation_c/__init__.py" in code
    code = code.replace("/path/to/repo", ROOT)
    mod = types.ModuleType("triangulation_c")
    exec(compile(code, "INTEGRATION.md#2", "exec"), mod.__dict__)
    return mod


def test_stub_declares_the_reference_wrapper_signatures():
    import inspect
    stub = _stub_module()
    # Work/python_libs/triangulation_c/__init__.py:18,51
    assert list(inspect.signature(stub.linear_LS_triangulation).parameters) == ["u1", "P1", "u2", "P2"]
    sig = inspect.signature(stub.iterative_LS_triangulation)
    assert list(sig.parameters) == ["u1", "P1", "u2", "P2", "tolerance"] and sig.parameters["tolerance"].default == 3.e-5
    assert isinstance(stub.loaded, bool)


@pytest.mark.skipif(not os.path.isfile(REF_TRIANGULATION), reason="the reference checkout is only present in the build container")
def test_stub_binds_under_the_reference_override_shim():
    stub = _stub_module()
    src = open(REF_TRIANGULATION).read().split("\n")
    # lines 1-233: the four pure-Python solvers; 236-253: the override shim's `import` + `if loaded:` branch (the `else`
    # branch, 254-256, is two Python-2 print statements); 259-267: output_dtype
    shim = "\n".join(src[235:253])
    assert "\nimport triangulation_c\nif triangulation_c.loaded:" in "\n" + shim and "linear_LS_triangulation_c(*args)" in shim
    body = "\n".join(src[:233]) + "\n" + shim + "\n" + "\n".join(src[258:])
    stub.loaded = True                        # take the shim's optimised branch even though this box has no GPU
    saved = sys.modules.get("triangulation_c")
    sys.modules["triangulation_c"] = stub
    try:
        ref = types.ModuleType("reference_triangulation")
        exec(compile(body, "reference triangulation.py (not copied)", "exec"), ref.__dict__)
    finally:
        if saved is None:
            del sys.modules["triangulation_c"]
        else:
            sys.modules["triangulation_c"] = saved
    # the shim replaced the pure-Python functions by wrappers around the stub's
    assert ref.linear_LS_triangulation_c is stub.linear_LS_triangulation
    assert ref.iterative_LS_triangulation_c is stub.iterative_LS_triangulation
    assert ref.linear_LS_triangulation.__code__.co_varnames[:1] == ("args",)
    u1, P1, u2, P2, _ = rig.make_correspondences(257, "rotating", 0.8)
    import triangl_cuda
    if triangl_cuda.device_count() > 0:
        x, st = ref.linear_LS_triangulation(u1, P1, u2, P2)
        xo, so = orc.linear_LS_triangulation(u1, P1, u2, P2)
        assert np.array_equal(st, so) and np.allclose(x, xo, rtol=1e-9, atol=0)
        ref.set_triangl_output_dtype(np.float32)
        x32, _ = ref.iterative_LS_triangulation(u1, P1, u2, P2, tolerance=3e-5)
        assert x32.dtype == np.float32
    else:
        # the call crosses ctypes into the C ABI and comes back with the library's own error
        for call in (lambda: ref.linear_LS_triangulation(u1, P1, u2, P2),
                     lambda: ref.iterative_LS_triangulation(u1, P1, u2, P2, tolerance=1e-4)):
            with pytest.raises(RuntimeError, match="no CUDA device available"):
                call()


@pytest.mark.gpu
def test_stub_results_match_the_oracle_on_gpu():
    stub = _stub_module()
    assert stub.loaded
    u1, P1, u2, P2, _ = rig.make_correspondences(30011, "rotating", 0.8)
    P1f = np.eye(4); P1f[:3] = P1
    P2f = np.eye(4); P2f[:3] = P2                        # 4x4 matrices work unchanged (triangulation.c:24-25)
    x, st = stub.linear_LS_triangulation(u1.astype(np.float32), P1f, u2.astype(np.float32), P2f)     # up-cast like :32-33
    xo, so = orc.linear_LS_triangulation(u1.astype(np.float32).astype(np.float64), P1, u2.astype(np.float32).astype(np.float64), P2)
    assert st.dtype == np.bool_ and st.all() and np.max(np.abs(x - xo) / np.max(np.abs(xo), axis=1, keepdims=True)) < 1e-9
    x, st = stub.iterative_LS_triangulation(u1, P1, u2, P2, tolerance=3e-5)
    xo, so, _, margin = orc.iterative_LS_core(u1, P1, u2, P2, 3e-5, 'c')
    keep = margin > 1e-9
    assert st.dtype == np.int32 and np.array_equal(st[keep], so[keep])
    assert np.max((np.abs(x - xo) / np.max(np.abs(xo), axis=1, keepdims=True))[keep]) < 1e-9
