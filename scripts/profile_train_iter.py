"""Where one training iteration spends its time (host-side sections with a device sync after each; 100k Gaussians,
512^2 detector, 32^3 TV crop)."""
import json, os, sys, time, types
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from r2_gaussian_b200 import losses, scene
from r2_gaussian_b200.gaussian_model import GaussianModel
from r2_gaussian_b200.render_query import query, render

dev = "cuda"
scanner = scene.cone_beam_scanner(512)
cams = [scene.camera_from_view(v) for v in scene.make_views(scanner, 8)]
cloud = scene.make_cloud(100000, seed=0)
opt = types.SimpleNamespace(position_lr_init=2e-4, position_lr_final=2e-5, position_lr_max_steps=30000,
    density_lr_init=1e-2, density_lr_final=1e-3, density_lr_max_steps=30000, scaling_lr_init=5e-3, scaling_lr_final=5e-4,
    scaling_lr_max_steps=30000, rotation_lr_init=1e-3, rotation_lr_final=1e-4, rotation_lr_max_steps=30000)
gm = GaussianModel((0.001, 1.0)); gm.create_from_pcd(cloud.means, np.maximum(cloud.density, 1e-3), 1.0); gm.training_setup(opt)
pipe = types.SimpleNamespace(compute_cov3D_python=False, debug=False)
with torch.no_grad():
    gts = [render(c, gm, pipe)["render"] * 0.9 for c in cams]
acc = {}
def sec(name, t0):
    torch.cuda.synchronize(); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0; return time.perf_counter()
N = 60
for i in range(N + 10):
    if i == 10: acc.clear()
    t = time.perf_counter()
    gm.update_learning_rate(i + 1); t = sec("lr", t)
    pkg = render(cams[i % 8], gm, pipe); t = sec("render_fwd(+activations)", t)
    loss = losses.image_loss(pkg["render"], gts[i % 8], 0.25)["total"]; t = sec("image_loss", t)
    vol = query(gm, [0.1, 0.0, -0.1], [32] * 3, [0.25] * 3, pipe)["vol"]; t = sec("query_fwd", t)
    loss = loss + 0.05 * losses.tv_3d_loss(vol, "mean"); t = sec("tv_loss", t)
    loss.backward(); t = sec("backward", t)
    with torch.no_grad():
        vis = pkg["visibility_filter"]
        gm.update_max_radii(pkg["radii"], vis)
        gm.add_densification_stats(pkg["viewspace_points"], vis); t = sec("densify_stats", t)
    gm.optimizer.step(); gm.optimizer.zero_grad(set_to_none=True); t = sec("adam", t)
out = {k: v / N * 1e3 for k, v in acc.items()}
out["sum_ms"] = sum(out.values())
print(json.dumps(out))
