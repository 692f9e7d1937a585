"""Compile CUDA C++ emitted by the front-end with g++ and evaluate it on the CPU (tests only).

This checks the front-end + emitter without a GPU: the emitted text only uses s2m_vec.h /
s2m_math.h / s2m_sdf3d_lib.h, which are host/device headers.  It is NOT a product path.
"""
import ctypes
import hashlib
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CSRC = os.path.join(ROOT, "sdf2mesh_b200", "csrc")
_CACHE = {}

WRAP = r'''
#include "s2m_sdf3d_lib.h"
namespace s2m_user {
using namespace s2m;
%s
}
extern "C" float host_eval1(float x, float y, float z) { return s2m_user::sdf3d(s2m::mk3(x, y, z)); }
extern "C" void host_eval(const float* pts, float* out, unsigned long long n) {
  for (unsigned long long i = 0; i < n; ++i)
    out[i] = s2m_user::sdf3d(s2m::mk3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]));
}
'''


def compile_host(cuda_body: str):
    key = hashlib.sha1(cuda_body.encode()).hexdigest()
    if key in _CACHE:
        return _CACHE[key]
    d = tempfile.mkdtemp(prefix="s2m_host_")
    src = os.path.join(d, "sdf.cpp")
    so = os.path.join(d, "sdf.so")
    with open(src, "w") as f:
        f.write(WRAP % cuda_body)
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-mfma", "-std=c++17", "-fPIC", "-shared",
                           "-I", CSRC, src, "-o", so])
    lib = ctypes.CDLL(so)
    lib.host_eval.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64]
    _CACHE[key] = lib
    return lib


def scalar_function(cuda_body: str):
    """the compiled `float sdf3d(x, y, z)` as a ctypes function (for oracle.set_plugin); the library
    stays loaded for the life of the process"""
    return compile_host(cuda_body).host_eval1


def eval_points(cuda_body: str, pts) -> np.ndarray:
    lib = compile_host(cuda_body)
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
    out = np.empty(pts.shape[0], np.float32)
    lib.host_eval(pts.ctypes.data, out.ctypes.data, pts.shape[0])
    return out
