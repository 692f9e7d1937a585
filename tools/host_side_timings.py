#!/usr/bin/env python3
"""Host-side costs around the meshing path, one JSON line each (no GPU needed for the first two):

  jit      NVRTC time per example SDF, one program vs three concurrent programs (S2M_JIT_SPLIT)
  writer   STL / PLY writer throughput on a synthetic mesh (oracle torus 512^3 replicated to ~6 M triangles)
  cli      wall time of the sdf2mesh executable end to end, with its --stats lines (needs a GPU)

usage: python tools/host_side_timings.py [jit] [writer] [cli]      (default: jit writer)
"""
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import sdf2mesh_b200 as s2m  # noqa: E402

EX = os.path.join(ROOT, "examples")
SHADERS = [("torus.sdf3d", 0), ("martin_cube.sdf3d", 0), ("p_key.sdf3d", 0), ("mandelmesh.frag", 1)]


def load(f, kind):
    return s2m.Sdf3DShader.from_glsl_fragment_shader(os.path.join(EX, f), "sdf") if kind else s2m.Sdf3DShader.from_path(os.path.join(EX, f))


def jit(reps=5):
    load(*SHADERS[0]).create_shader_module(None)   # NVRTC's own first-call cost is not the subject
    for f, kind in SHADERS:
        out = {"what": "jit", "sdf": f, "cores": os.cpu_count()}
        for split in ("0", "1"):
            os.environ["S2M_JIT_SPLIT"] = split
            ms = [load(f, kind).create_shader_module(None).compile_ms()[1] for _ in range(reps)]
            out["one_program_ms" if split == "0" else "three_programs_ms"] = {"median": round(statistics.median(ms), 1), "min": round(min(ms), 1)}
        os.environ.pop("S2M_JIT_SPLIT")
        print(json.dumps(out), flush=True)


def writer(reps=3, copies=8):
    import oracle
    o = oracle.mesh_run("torus", 512, 2.0)
    nv = len(o.positions)
    pos, nrm = np.concatenate([o.positions] * copies), np.concatenate([o.normals] * copies)
    quads = np.concatenate([o.quads + np.uint64(i * nv) for i in range(copies)])
    d = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    for name, path, binary in (("stl_ascii", d + "/s2m_t.stl", False), ("ply_ascii", d + "/s2m_t.ply", False), ("stl_binary", d + "/s2m_t.stl", True)):
        secs = []
        for _ in range(reps):
            t = time.perf_counter()
            s2m.write_mesh_arrays([(pos, nrm, quads)], path, binary_stl=binary)
            secs.append(time.perf_counter() - t)
        size = os.path.getsize(path)
        os.unlink(path)
        best = min(secs)
        print(json.dumps({"what": "writer", "format": name, "triangles": 2 * len(quads), "bytes": size, "cores": os.cpu_count(), "seconds_min": round(best, 4),
                          "Mtriangles_per_s": round(2 * len(quads) / best / 1e6, 2), "MB_per_s": round(size / best / 1e6, 1), "to": d}), flush=True)


def cli():
    """every configuration twice in fresh processes with a private cubin cache: the first run compiles with NVRTC and
    meshes on a cold context (no pinned pages, no device buffers), the second finds the cubins on disk"""
    exe = os.path.join(ROOT, "sdf2mesh_b200", "sdf2mesh")
    d = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    cache = tempfile.mkdtemp(prefix="s2m_cache_")
    env = {**os.environ, "S2M_CACHE_DIR": cache}
    for args, label in ((["--glsl", os.path.join(EX, "mandelmesh.frag"), "-r", "2048", "-b", "5", "--binary-stl"], "mandelmesh 2048^3 binary STL"),
                        (["--glsl", os.path.join(EX, "mandelmesh.frag"), "-r", "1024", "-b", "5"], "mandelmesh 1024^3 ASCII STL"),
                        (["--sdf", os.path.join(EX, "torus.sdf3d"), "-r", "128", "-b", "2"], "torus 128^3 ASCII STL")):
        for attempt in ("cold: NVRTC", "cubins from the disk cache"):
            out = d + "/s2m_cli.stl"
            t = time.perf_counter()
            p = subprocess.run([exe, *args, "--mesh", out, "--stats"], capture_output=True, text=True, env=env)
            wall = time.perf_counter() - t
            print(json.dumps({"what": "cli", "run": label, "jit": attempt, "rc": p.returncode, "wall_s": round(wall, 3), "bytes": os.path.getsize(out) if os.path.exists(out) else 0,
                              "stats": [l for l in p.stderr.splitlines() if l.startswith("stats:")]}), flush=True)
            if os.path.exists(out):
                os.unlink(out)
    import shutil
    shutil.rmtree(cache, ignore_errors=True)


if __name__ == "__main__":
    for w in (sys.argv[1:] or ["jit", "writer"]):
        {"jit": jit, "writer": writer, "cli": cli}[w]()
