"""
Load the UNMODIFIED reference solvers from /root/reference for fixture generation (build container only).

Work/python_libs/triangulation.py does not import on Python 3 (print statements at :255-256 and the
weave-only `import triangulation_c` at :237), but lines 1-233 (the four pure-Python solvers) and
259-267 (output dtype switch) exec cleanly under py3 + cv2 4.x.  Nothing is copied: the source is read
from where it lies and exec'd into a fresh module object.  /root/reference does not exist on the GPU
box, so only oracle/make_golden.py (run here, outputs committed under tests/golden/) may call this.
"""
import os
import types

REFERENCE_ROOT = os.environ.get("TRGL_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "Work/python_libs/triangulation.py"))


def load_reference_triangulation():
    path = os.path.join(REFERENCE_ROOT, "Work/python_libs/triangulation.py")
    src = open(path).read().split("\n")
    mod = types.ModuleType("triangulation_reference")
    exec(compile("\n".join(src[:233] + src[258:]), path, "exec"), mod.__dict__)
    return mod
