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


WRAP_PACKED = r'''
#include "s2m_pvec.h"
namespace s2m_user_p {
using namespace s2m;
%s
}
// lane lo = point a[i], lane hi = point b[i]
extern "C" void host_eval2(const float* a, const float* b, float* out_a, float* out_b, unsigned char* dv, unsigned long long n) {
  for (unsigned long long i = 0; i < n; ++i) {
    bool d = false;
    const s2m::pf r = s2m_user_p::sdf3d2(s2m::pmk3(s2m::pf(a[3 * i], b[3 * i]), s2m::pf(a[3 * i + 1], b[3 * i + 1]), s2m::pf(a[3 * i + 2], b[3 * i + 2])), &d);
    out_a[i] = r.lo; out_b[i] = r.hi; dv[i] = d ? 1 : 0;
  }
}
'''


def compile_host_packed(packed_body: str):
    """the packed (f32x2) form of a shader, compiled for the host: pairs are two scalar operations there"""
    key = "p" + hashlib.sha1(packed_body.encode()).hexdigest()
    if key in _CACHE:
        return _CACHE[key]
    d = tempfile.mkdtemp(prefix="s2m_hostp_")
    src = os.path.join(d, "sdfp.cpp")
    so = os.path.join(d, "sdfp.so")
    with open(src, "w") as f:
        f.write(WRAP_PACKED % packed_body)
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-mfma", "-std=c++17", "-fPIC", "-shared",
                           "-I", CSRC, src, "-o", so])
    lib = ctypes.CDLL(so)
    lib.host_eval2.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_uint64]
    _CACHE[key] = lib
    return lib


def eval_pairs(packed_body: str, pts_a, pts_b):
    """-> (values lane lo, values lane hi, lanes-disagreed flags)"""
    lib = compile_host_packed(packed_body)
    a = np.ascontiguousarray(pts_a, np.float32).reshape(-1, 3)
    b = np.ascontiguousarray(pts_b, np.float32).reshape(-1, 3)
    assert a.shape == b.shape
    oa = np.empty(a.shape[0], np.float32)
    ob = np.empty(a.shape[0], np.float32)
    dv = np.empty(a.shape[0], np.uint8)
    lib.host_eval2(a.ctypes.data, b.ctypes.data, oa.ctypes.data, ob.ctypes.data, dv.ctypes.data, a.shape[0])
    return oa, ob, dv.astype(bool)


def scalar_function(cuda_body: str):
    """the compiled `float sdf3d(x, y, z)` as a ctypes function (for oracle.set_plugin); the library
    stays loaded for the life of the process"""
    return compile_host(cuda_body).host_eval1


# scalar body -> packed body of the same shader (tests/conftest.py registers every shader a test lowers);
# eval_points() then also runs the packed form on the same points and checks it lane by lane
_PACKED = {}
PACKED_STATS = {"shaders": set(), "pairs": 0, "disagreed": 0, "unpackable": set()}


def register_packed(cuda_body: str, packed_body: str) -> None:
    if packed_body:
        _PACKED[cuda_body] = packed_body
    else:
        PACKED_STATS["unpackable"].add(hashlib.sha1(cuda_body.encode()).hexdigest())


def _same(a, b):
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))


def check_packed(cuda_body: str, packed_body: str, pts, scalar_out) -> None:
    """The packed (f32x2) form against the scalar form: lane lo must ALWAYS carry the scalar value of
    its point; lane hi must carry it whenever the lanes-disagreed flag stayed clear.  Pairs: each point
    with (i) its +x neighbour one 2048^3-voxel away, (ii) an unrelated point."""
    lib = compile_host(cuda_body)
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)[:20000]
    scalar_out = scalar_out[:pts.shape[0]]
    near = pts.copy()
    near[:, 0] += np.float32(5.0 / 2047.0)
    far = np.roll(pts, 1, axis=0)
    for other in (near, far):
        other = np.ascontiguousarray(other)
        want = np.empty(other.shape[0], np.float32)
        lib.host_eval(other.ctypes.data, want.ctypes.data, other.shape[0])
        lo, hi, dv = eval_pairs(packed_body, pts, other)
        assert _same(lo, scalar_out).all(), "packed form: lane lo differs from the scalar evaluation"
        assert _same(hi, want)[~dv].all(), "packed form: lane hi differs from the scalar evaluation although the lanes agreed"
        PACKED_STATS["pairs"] += int(pts.shape[0])
        PACKED_STATS["disagreed"] += int(dv.sum())
    PACKED_STATS["shaders"].add(hashlib.sha1(cuda_body.encode()).hexdigest())


def eval_points(cuda_body: str, pts) -> np.ndarray:
    lib = compile_host(cuda_body)
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
    out = np.empty(pts.shape[0], np.float32)
    lib.host_eval(pts.ctypes.data, out.ctypes.data, pts.shape[0])
    packed = _PACKED.get(cuda_body)
    if packed and pts.shape[0] > 0:
        check_packed(cuda_body, packed, pts, out)
    return out
