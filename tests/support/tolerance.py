"""Tolerance-mode parity (BASELINE.json north_star): "identical set of active cells and identical quad
connectivity (bit-exact, excluding and listing corners with |sdf| < 1e-6 voxel), and vertex
positions within 1e-4 voxel" -- a mesh (from the GPU, or from oracle.cpp) against the INDEPENDENT
evaluation in oracle/indep.cpp (f64 + libm, or f32 + libm; no s2m_math.h anywhere).

`compare` does the set logic and returns a report; the callers assert on it.  Nothing here is
product code.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

NORTH_STAR_EXCLUDE_VOXELS = 1e-6
NORTH_STAR_POSITION_VOXELS = 1e-4


def _neighbours(keys: np.ndarray) -> Dict[str, np.ndarray]:
    """keys of the cells a vertex's three quads name (mesh.rs:286-320), as label keys"""
    x = keys & np.uint64(0xFFFF)
    y = (keys >> np.uint64(16)) & np.uint64(0xFFFF)
    z = keys >> np.uint64(32)
    def k(dx, dy, dz):
        return ((x - np.uint64(dx)) & np.uint64(0xFFFF)) | (((y - np.uint64(dy)) & np.uint64(0xFFFF)) << np.uint64(16)) | ((z - np.uint64(dz)) << np.uint64(32))
    return {"x": x, "y": y, "z": z, "k": k}


def key_quads(keys: np.ndarray, nibbles: np.ndarray, present: np.ndarray) -> np.ndarray:
    """Quads as 4-tuples of CELL KEYS (after Quad::swap), for the vertices `keys` (ascending, unique)
    with sign nibbles `nibbles`; a quad is produced only if its four cells are all in `present`
    (sorted key array).  mesh.rs:286-320, lib.rs:190-197.  Returned sorted lexicographically."""
    keys = np.asarray(keys, np.uint64)
    nb = _neighbours(keys)
    x, y, z, k = nb["x"], nb["y"], nb["z"], nb["k"]
    s100 = (nibbles & 1) != 0; s010 = (nibbles & 2) != 0; s001 = (nibbles & 4) != 0; s000 = (nibbles & 8) != 0
    out = []
    def add(cond, a, b, c, d, swap):
        q = np.stack([a, b, c, d], axis=1)[cond]
        sw = swap[cond]
        q[sw] = q[sw][:, ::-1]
        out.append(q)
    add((s100 != s000) & (y > 0) & (z > 0), k(0, 1, 1), k(0, 0, 1), keys, k(0, 1, 0), s100)
    add((s010 != s000) & (x > 0) & (z > 0), k(1, 0, 1), k(0, 0, 1), keys, k(1, 0, 0), ~s010)
    add((s001 != s000) & (x > 0) & (y > 0), k(1, 1, 0), k(0, 1, 0), keys, k(1, 0, 0), s001)
    q = np.concatenate(out) if out else np.zeros((0, 4), np.uint64)
    if len(q):
        pos = np.searchsorted(present, q.reshape(-1))
        pos = np.minimum(pos, max(len(present) - 1, 0))
        ok = (present[pos] == q.reshape(-1)).reshape(-1, 4).all(axis=1) if len(present) else np.zeros(len(q), bool)
        q = q[ok]
    if len(q):
        q = q[np.lexsort((q[:, 3], q[:, 2], q[:, 1], q[:, 0]))]
    return q


def index_quads_to_keys(quads: np.ndarray, keys: np.ndarray, index_base: int = 0) -> np.ndarray:
    """a mesh's index quads -> key quads; quads naming a vertex outside [index_base, index_base+len(keys)) are dropped"""
    q = np.asarray(quads, np.int64).reshape(-1, 4) - int(index_base)
    ok = ((q >= 0) & (q < len(keys))).all(axis=1)
    q = q[ok]
    kq = np.asarray(keys, np.uint64)[q] if len(q) else np.zeros((0, 4), np.uint64)
    if len(kq):
        kq = kq[np.lexsort((kq[:, 3], kq[:, 2], kq[:, 1], kq[:, 0]))]
    return kq


@dataclass
class Report:
    n_mesh: int = 0                 # vertices of the mesh under test
    n_indep_active: int = 0         # active cells of the independent evaluation
    n_excluded_cells: int = 0       # cells touching a corner with |sdf| < exclude_voxels
    n_excluded_corner_values: int = 0   # such corner values, counted per cell (a grid corner is shared by up to 8 cells)
    exclude_voxels: float = NORTH_STAR_EXCLUDE_VOXELS
    residual_keys: List[int] = field(default_factory=list)   # active-set differences that survive the exclusion
    residual_min_abs_voxels: List[float] = field(default_factory=list)  # their closest corner, in voxels
    identity_threshold_voxels: float = 0.0   # smallest exclusion threshold at which the active sets are identical
    nibble_mismatches: int = 0      # on common, non-excluded cells
    max_position_error_voxels: float = 0.0   # on common, non-excluded cells
    p999_position_error_voxels: float = 0.0
    n_position_over: int = 0        # positions off by more than position_voxels
    quads_compared: int = 0
    quad_mismatches: int = 0        # symmetric difference of key quads whose 4 cells are common and non-excluded
    seconds_indep: float = 0.0

    def summary(self) -> dict:
        d = dict(self.__dict__)
        d["residual_keys"] = len(self.residual_keys)
        d["residual_min_abs_voxels"] = sorted(self.residual_min_abs_voxels)[-5:]
        return d


def compare(mesh_keys, mesh_positions, mesh_nibbles, indep, mesh_quads: Optional[np.ndarray] = None, quad_index_base: int = 0,
            exclude_voxels: float = NORTH_STAR_EXCLUDE_VOXELS, position_voxels: float = NORTH_STAR_POSITION_VOXELS) -> Report:
    """mesh_* : vertices of the z-range `indep` covers (label keys ascending).  indep: oracle.IndepCells."""
    mk = np.asarray(mesh_keys, np.uint64)
    mp = np.asarray(mesh_positions, np.float64).reshape(-1, 3)
    mn = np.asarray(mesh_nibbles, np.uint8)
    rep = Report(n_mesh=len(mk), exclude_voxels=exclude_voxels, seconds_indep=indep.seconds)
    vox = indep.voxel
    ik, ia = indep.keys, indep.active
    rep.n_indep_active = int(ia.sum())
    excl = indep.min_abs < exclude_voxels * vox
    rep.n_excluded_cells = int(excl.sum())
    rep.n_excluded_corner_values = int((np.abs(indep.corners[excl]) < exclude_voxels * vox).sum()) if excl.any() else 0
    excl_keys = ik[excl]
    act_keys = ik[ia]
    # active-set difference (every cell the independent evaluation did not list is inactive there and far from the surface)
    only_mesh = np.setdiff1d(mk, act_keys, assume_unique=True)
    only_indep = np.setdiff1d(act_keys, mk, assume_unique=True)
    diff = np.concatenate([only_mesh, only_indep])
    # how close to the surface is each differing cell?  (cells the independent run did not list at all have
    # min_abs >= its list threshold: report them as +inf so that they can never be "explained")
    pos_in = np.searchsorted(ik, diff)
    pos_in = np.minimum(pos_in, max(len(ik) - 1, 0))
    listed = (ik[pos_in] == diff) if len(ik) else np.zeros(len(diff), bool)
    mabs = np.where(listed, indep.min_abs[pos_in] / vox, np.inf) if len(diff) else np.zeros(0)
    rep.identity_threshold_voxels = float(mabs.max()) if len(diff) else 0.0
    resid = mabs >= exclude_voxels
    rep.residual_keys = [int(k) for k in diff[resid]]
    rep.residual_min_abs_voxels = [float(v) for v in mabs[resid]]
    # common, non-excluded cells: nibbles and positions
    common = np.intersect1d(mk, act_keys, assume_unique=True)
    common = np.setdiff1d(common, excl_keys, assume_unique=True)
    im = np.searchsorted(mk, common); ii = np.searchsorted(ik, common)
    rep.nibble_mismatches = int((mn[im] != indep.nibbles[ii]).sum())
    if len(common):
        err = np.abs(mp[im] - indep.positions[ii]).max(axis=1) / vox
        rep.max_position_error_voxels = float(err.max())
        rep.p999_position_error_voxels = float(np.quantile(err, 0.999))
        rep.n_position_over = int((err > position_voxels).sum())
    # quads over common non-excluded cells
    if mesh_quads is not None:
        present = common
        q_mesh = index_quads_to_keys(mesh_quads, mk, quad_index_base)
        if len(q_mesh):
            p = np.searchsorted(present, q_mesh.reshape(-1)); p = np.minimum(p, max(len(present) - 1, 0))
            ok = (present[p] == q_mesh.reshape(-1)).reshape(-1, 4).all(axis=1) if len(present) else np.zeros(len(q_mesh), bool)
            q_mesh = q_mesh[ok]
        # the independent side's quads: emitted by its own active non-excluded vertices with ITS nibbles
        q_ind = key_quads(common, indep.nibbles[ii], present)
        rep.quads_compared = int(len(q_ind))
        a = {tuple(r) for r in q_mesh.tolist()}
        b = {tuple(r) for r in q_ind.tolist()}
        rep.quad_mismatches = len(a ^ b)
    return rep
