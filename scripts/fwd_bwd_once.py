"""A few forward+backward calls of the headline scene (for ncu launch lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from r2_gaussian_b200 import _C
class A: gaussians=100000; detector=512; views=50; cloud="init"
sc, views, cloud = bench.build_scene(A)
dev = torch.device("cuda"); E = torch.Tensor([])
m = torch.tensor(cloud.means, device=dev); s = torch.tensor(cloud.scales, device=dev); r = torch.tensor(cloud.rotations, device=dev); d = torch.tensor(cloud.density, device=dev)
dv = bench.device_views(views, dev); dL = torch.randn(1, 512, 512, device=dev)
for i in range(6):
    v = dv[i]
    R, img, radii, geom, binning, imgb = _C.rasterize_gaussians(m, d, s, r, 1.0, E, v["view"], v["proj"], v["tx"], v["ty"], 512, 512, v["campos"], False, v["mode"], False)
    _C.rasterize_gaussians_backward(m, radii, s, r, 1.0, E, v["view"], v["proj"], v["tx"], v["ty"], dL, v["campos"], geom, R, binning, imgb, v["mode"], False)
    R2, vol, rx, ry, rz, g2, b2, i2 = _C.voxelize_gaussians(m, d, s, r, 1.0, E, 32, 32, 32, 0.25, 0.25, 0.25, 0.3, -0.4, 0.1, False, False)
    _C.voxelize_gaussians_backward(m, rx, ry, rz, s, r, 1.0, E, torch.randn(32, 32, 32, device=dev), g2, R2, b2, i2, 32, 32, 32, 0.25, 0.25, 0.25, 0.3, -0.4, 0.1, False)
torch.cuda.synchronize()
