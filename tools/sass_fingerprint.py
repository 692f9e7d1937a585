#!/usr/bin/env python3
"""Fingerprint of the device code of the example SDFs, without a GPU: SHA-1 of the emitted CUDA C++ (scalar / packed),
and per NVRTC program (K1 | K4a | diagnostics) the number of SASS instruction lines and their SHA-1 (cuobjdump -sass,
instruction lines only, so line tables do not matter).  Equal fingerprints before and after a host-side change mean the
kernels a GPU would run are the same ones that were validated.

usage: python tools/sass_fingerprint.py [repo root]   > fingerprint.txt
"""
import hashlib
import os
import subprocess
import sys
import tempfile

root = os.path.abspath(sys.argv[1]) if len(sys.argv) > 1 else os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
import sdf2mesh_b200 as s2m  # noqa: E402


def sha(b):
    return hashlib.sha1(b if isinstance(b, bytes) else b.encode()).hexdigest()[:12]


for f, glsl in (("torus.sdf3d", False), ("martin_cube.sdf3d", False), ("p_key.sdf3d", False), ("mandelmesh.frag", True)):
    path = os.path.join(root, "examples", f)
    sh = s2m.Sdf3DShader.from_glsl_fragment_shader(path, "sdf") if glsl else s2m.Sdf3DShader.from_path(path)
    m = sh.create_shader_module(None)
    parts = []
    for c in m.cubins():
        with tempfile.NamedTemporaryFile(suffix=".cubin", delete=False) as t:
            t.write(c)
        sass = subprocess.run(["cuobjdump", "-sass", t.name], capture_output=True, text=True).stdout
        os.unlink(t.name)
        lines = [l for l in sass.splitlines() if "/*" in l and not l.strip().startswith("//")]
        parts.append(f"{len(lines)}:{sha(chr(10).join(lines))}")
    print(f"{f:20s} cuda {sha(sh.lower_to_cuda())} packed {sha(sh.lower_to_cuda_packed())} unit {sha(m.cuda_source)} packed_k1 {int(m.packed)} sass {' '.join(parts)}")
