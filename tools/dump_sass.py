#!/usr/bin/env python3
"""NVRTC-compile an SDF (no GPU needed) and dump the SASS of its kernels with cuobjdump.

usage: python tools/dump_sass.py examples/mandelmesh.frag [kernel-name-substring] > out.sass
"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sdf2mesh_b200 as s2m  # noqa: E402


def main():
    path = sys.argv[1]
    sh = s2m.Sdf3DShader.from_glsl_fragment_shader(path, "sdf") if path.endswith((".frag", ".glsl")) else s2m.Sdf3DShader.from_path(path)
    m = sh.create_shader_module(None)
    for part in m.cubins():  # K1 | K4a | diagnostic kernels
        with tempfile.NamedTemporaryFile(suffix=".cubin", delete=False) as f:
            f.write(part)
        args = ["cuobjdump", "-sass", f.name]
        if len(sys.argv) > 2:
            args += ["-fun", sys.argv[2]]
        sys.stdout.write(subprocess.run(args, capture_output=True, text=True).stdout)
        os.unlink(f.name)


if __name__ == "__main__":
    main()
