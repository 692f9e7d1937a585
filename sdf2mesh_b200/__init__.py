"""sdf2mesh_b200 -- B200-native SDF -> dual-contoured quad mesh engine.

Drop-in for the meshing path of WilstonOreo/sdf2mesh (the reference): same `.sdf3d` / GLSL
inputs, same `Sdf3DShader` surface, same vertex / quad order; CUDA for sm_100a underneath.
"""
from ._capi import (COMPILE_ALLOW_FMA, MESH_ALL_SLICES, MESH_CLASSIFY_FROM_SLAB, MESH_KEEP_INVALID, MESH_CONSISTENT_CORNERS, MESH_QUADS_U32, MESH_NO_SLAB, MESH_RELATIVE_QUADS, MESH_TIMINGS, MESH_EXACT_DENSE, MESH_KEEP_CANDIDATES, MESH_NO_NORMALS,
                    SRC_CUDA, SRC_GLSL_FRAGMENT, SRC_SDF3D, SRC_WGSL, MeshParams, S2mError)
from .engine import (Context, MeshData, MeshResult, Module, cost_probe, debug_slab_plane, make_params, mesh_begin,
                     mesh_run, params_from_cli, write_mesh_arrays, write_mesh_parts)
from .multi import MultiContext
from .shader import Sdf3DShader, WgslShaderCode, convert_glsl_to_wgsl

__all__ = [
    "Context", "MultiContext", "Module", "MeshParams", "MeshResult", "MeshData", "Sdf3DShader", "WgslShaderCode", "S2mError",
    "convert_glsl_to_wgsl", "make_params", "params_from_cli", "mesh_begin", "mesh_run", "write_mesh_parts", "write_mesh_arrays", "debug_slab_plane", "cost_probe",
    "MESH_ALL_SLICES", "MESH_CLASSIFY_FROM_SLAB", "MESH_KEEP_INVALID", "MESH_CONSISTENT_CORNERS", "MESH_QUADS_U32", "MESH_NO_SLAB", "MESH_RELATIVE_QUADS", "MESH_TIMINGS", "MESH_NO_NORMALS", "MESH_EXACT_DENSE", "MESH_KEEP_CANDIDATES", "COMPILE_ALLOW_FMA",
    "SRC_SDF3D", "SRC_GLSL_FRAGMENT", "SRC_WGSL", "SRC_CUDA",
]
