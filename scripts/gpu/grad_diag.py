"""Three-way comparison of the raster gradients on the headline scene: ours / the compiled reference (float atomics) /
the CPU oracle (float64 accumulation).  Prints, per gradient array, the worst violation of |a-b| <= rtol|b| + atol max|b|
for each pair, so that one can tell reference noise from a defect of ours."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
from r2_gaussian_b200 import scene

P = int(os.environ.get("DIAG_P", "100000")); n = int(os.environ.get("DIAG_DET", "512")); kind = os.environ.get("DIAG_KIND", "init")
cloud = scene.make_cloud(P, kind=kind, seed=0)
view = scene.make_view(scene.cone_beam_scanner(n, 256), float(os.environ.get("DIAG_ANGLE", "0.0")))
dL = np.random.RandomState(5).randn(n, n).astype(np.float32)
ref = util.run_ref_raster(cloud, view, dL)
ref2 = util.run_ref_raster(cloud, view, dL)          # the reference against itself: its run-to-run atomics noise
ours = util.ours_raster_forward(cloud, view, export=False)
g = util.ours_raster_backward(cloud, view, ours, dL)
orc = util.oracle_raster_forward(cloud, view)
go = util.oracle_raster_backward(cloud, view, orc, dL)
out = {"P": P, "det": n, "kind": kind, "bwd_exact": os.environ.get("R2X_BWD_EXACT", "0")}
for k in ["dL_dmean2D", "dL_dopacity", "dL_dmean3D", "dL_dcov3D", "dL_dscale", "dL_drot"]:
    sc = None
    if k == "dL_drot":
        sc = max(float(np.abs(go[k]).max()), float(np.abs(go["dL_dscale"]).max()))
    row = {}
    for name, a, b in (("ours_vs_ref", g[k], ref["grads"][k]), ("ours_vs_oracle", g[k], go[k]), ("ref_vs_oracle", ref["grads"][k], go[k]),
                       ("ref_vs_ref", ref["grads"][k], ref2["grads"][k])):
        row[name] = round(util.grad_mismatch(a, b, 5e-4, 5e-5, sc), 3)
    d = np.abs(g[k].astype(np.float64) - go[k]); i = np.unravel_index(np.argmax(d), d.shape)
    row["worst_ours_vs_oracle"] = {"index": [int(x) for x in i], "ours": float(g[k][i]), "oracle": float(go[k][i]), "ref": float(ref["grads"][k][i]),
                                   "radius": int(orc["radii"][i[0]]), "max_abs": float(np.abs(go[k]).max())}
    a64 = g[k].astype(np.float64); b64 = ref["grads"][k].astype(np.float64)
    scale = sc if sc is not None else float(np.abs(b64).max())
    ratio = np.abs(a64 - b64) / (5e-4 * np.abs(b64) + 5e-5 * scale + 1e-30)
    j = np.unravel_index(np.argmax(ratio), ratio.shape)
    gi = int(j[0])
    co = orc["conic_opacity"][gi]; mu = float(orc["mu"][gi]); w = float(co[3]) * mu
    L = 1.4426950408889634
    row["worst_ours_vs_ref"] = {"index": [int(x) for x in j], "ratio": float(ratio[j]), "ours": float(g[k][j]), "ref": float(ref["grads"][k][j]),
                                "ref2": float(ref2["grads"][k][j]), "oracle": float(go[k][j]), "radius": int(orc["radii"][gi]), "conic": [float(x) for x in co[:3]],
                                "w": w, "lw": float(np.log2(max(w, 1e-300))), "A2": float(co[0]) * 0.5 * L, "xy": [float(x) for x in orc["xy"][gi]],
                                "n_over_tol": int((ratio > 1).sum())}
    out[k] = row
out["angle"] = float(os.environ.get("DIAG_ANGLE", "0.0"))
print(json.dumps(out))
