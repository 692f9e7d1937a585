#!/usr/bin/env python3
"""NVRTC-compile an SDF (no GPU needed) and dump the SASS of its kernels with cuobjdump.

usage: python tools/dump_sass.py examples/mandelmesh.frag [kernel-name-substring] > out.sass
"""
import ctypes
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sdf2mesh_b200 as s2m  # noqa: E402
from sdf2mesh_b200._capi import lib, check  # noqa: E402


def main():
    path = sys.argv[1]
    sh = s2m.Sdf3DShader.from_glsl_fragment_shader(path, "sdf") if path.endswith((".frag", ".glsl")) else s2m.Sdf3DShader.from_path(path)
    m = sh.create_shader_module(None)
    data, size = ctypes.c_void_p(), ctypes.c_size_t()
    check(lib().s2m_module_cubin(m._h, ctypes.byref(data), ctypes.byref(size)))
    with tempfile.NamedTemporaryFile(suffix=".cubin", delete=False) as f:
        f.write(ctypes.string_at(data, size.value))
    args = ["cuobjdump", "-sass", f.name]
    if len(sys.argv) > 2:
        args += ["-fun", sys.argv[2]]
    sys.stdout.write(subprocess.run(args, capture_output=True, text=True).stdout)
    os.unlink(f.name)


if __name__ == "__main__":
    main()
