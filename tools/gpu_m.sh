set -x
python - <<'PY'
import os, sys
sys.path.insert(0, ".")
import numpy as np
import sdf2mesh_b200 as s2m
ctx = s2m.Context(0)
m = s2m.Sdf3DShader.from_glsl_fragment_shader("examples/mandelmesh.frag", "sdf").create_shader_module(ctx)
R = 2048; size = np.float32(5.0) / np.float32(R - 1)
for zi in (100, 700, 1023, 1100):
    xs = (np.float32(-2.5) + size * np.arange(0, R, dtype=np.float32))
    X, Y = np.meshgrid(xs[0::2], xs, indexing="xy")
    a = np.stack([X.ravel(), Y.ravel(), np.full(X.size, np.float32(-2.5) + size * np.float32(zi), np.float32)], 1).astype(np.float32)
    b = a.copy(); b[:, 0] += size
    lo, hi, fl = m.eval_pairs(a, b, raw_flags=True)
    wa, wb = m.eval_points(a), m.eval_points(b)
    n = len(fl)
    ok = ((lo.view(np.uint32) == wa.view(np.uint32)) | (fl > 1)) & ((hi.view(np.uint32) == wb.view(np.uint32)) | (fl > 0))
    print("z", zi, "pairs", n, "dv %.4f%%" % (100 * np.mean(fl & 1 != 0)), "sl(2) %.4f%%" % (100 * np.mean(fl & 2 != 0)), "b1(4) %.4f%%" % (100 * np.mean(fl & 4 != 0)),
          "b2(8) %.4f%%" % (100 * np.mean(fl & 8 != 0)), "any-slow %.4f%%" % (100 * np.mean(fl > 1)), "exact", bool(ok.all()))
    inside = (np.sqrt((a.astype(np.float64) ** 2).sum(1)) < 2.0)
    print("   inside r<2: %.1f%% of pairs; slow among inside: %.4f%%" % (100 * inside.mean(), 100 * np.mean((fl > 1)[inside]) if inside.any() else 0))
    bad = np.nonzero(fl & 4)[0][:5]
    print("   examples b1:", a[bad], wa[bad])
PY
