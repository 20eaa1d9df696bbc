"""Times the raster render kernel alone (r2x_raster_render_only) on the headline scene; the variant is
chosen through R2X_RENDER_MINB (8, 6, 5 resident CTAs per SM) -- one process per variant."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
class A: gaussians=100000; detector=512; views=50; cloud="init"
sc, views, cloud = bench.build_scene(A)
dev = torch.device("cuda")
from r2_gaussian_b200.engine import RasterEngine
m = torch.tensor(cloud.means, device=dev); s = torch.tensor(cloud.scales, device=dev); r = torch.tensor(cloud.rotations, device=dev); d = torch.tensor(cloud.density, device=dev)
dv = bench.device_views(views, dev)
eng = RasterEngine(cloud.P, 512, 512, dev, capacity=1400000)
v = dv[7]
eng.forward(m, d, s, r, v["view"], v["proj"], v["campos"], v["tx"], v["ty"], v["mode"]); torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ms = bench.timed_steps(lambda i: eng.render_only(), 50, 5, flush, torch.cuda.synchronize)
print(json.dumps({"variant": os.environ.get("R2X_RENDER_MINB", "8"), "render_ms_mean": float(np.mean(ms)), "min": float(np.min(ms)), "R": eng.num_rendered()}))
