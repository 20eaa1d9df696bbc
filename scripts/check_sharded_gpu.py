"""torchrun --nproc-per-node N scripts/check_sharded_gpu.py : Gaussian-sharded projection + volume equal the
unsharded ones (1e-5), and sharded gradients equal the corresponding slices of the unsharded gradients."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch, torch.distributed as dist
from types import SimpleNamespace
import util
from r2_gaussian_b200 import scene
from r2_gaussian_b200.render_query import render, query
from r2_gaussian_b200 import sharded
from r2_gaussian_b200.sharded import shard_bounds, enable_peer_exchange

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
sharded.enable()                   # Gaussian sharding is an explicit opt-in of render() / query()
if "--p2p" in sys.argv:
    enable_peer_exchange(True)     # NVLink peer-memory sum instead of NCCL for the image / volume exchange
cloud, view = util.case("cone_trained_mid")
dev = torch.device("cuda", lr)
def model(c):
    t = {k: torch.tensor(v, device=dev, requires_grad=True) for k, v in dict(xyz=c.means, den=c.density, sc=c.scales, rot=c.rotations).items()}
    return SimpleNamespace(get_xyz=t["xyz"], get_density=t["den"], get_scaling=t["sc"], get_rotation=t["rot"]), t
cam = SimpleNamespace(image_height=view.image_height, image_width=view.image_width, FoVx=view.FoVx, FoVy=view.FoVy, mode=view.mode,
                      world_view_transform=torch.tensor(view.viewmatrix, device=dev), full_proj_transform=torch.tensor(view.projmatrix, device=dev),
                      camera_center=torch.tensor(view.campos, device=dev))
pipe = SimpleNamespace(debug=False, compute_cov3D_python=False)
lo, hi = shard_bounds(cloud.P, rank, world)
shard = scene.Cloud(cloud.means[lo:hi], cloud.scales[lo:hi], cloud.rotations[lo:hi], cloud.density[lo:hi])
pc, t = model(shard)
out = render(cam, pc, pipe)
dL = torch.randn(out["render"].shape, device=dev, generator=torch.Generator(dev).manual_seed(1))
(out["render"] * dL).sum().backward()
vol = query(pc, [0, 0, 0], [32, 32, 32], [2.0, 2.0, 2.0], pipe)["vol"]
# unsharded truth on every rank (process group untouched: world>1 would all-reduce -> use the extension directly)
from r2_gaussian_b200.rasterization import GaussianRasterizationSettings, GaussianRasterizer
from r2_gaussian_b200.voxelization import GaussianVoxelizationSettings, GaussianVoxelizer
import math
full = {k: torch.tensor(v, device=dev, requires_grad=True) for k, v in dict(xyz=cloud.means, den=cloud.density, sc=cloud.scales, rot=cloud.rotations).items()}
s = GaussianRasterizationSettings(view.image_height, view.image_width, view.tanfovx, view.tanfovy, 1.0, cam.world_view_transform, cam.full_proj_transform, cam.camera_center, False, view.mode, False)
img_full, _ = GaussianRasterizer(s)(full["xyz"], torch.zeros_like(full["xyz"]), full["den"], full["sc"], full["rot"])
(img_full * dL).sum().backward()
vol_full, _ = GaussianVoxelizer(GaussianVoxelizationSettings(1.0, 32, 32, 32, 2.0, 2.0, 2.0, 0.0, 0.0, 0.0, False, False))(full["xyz"].detach(), full["den"].detach(), full["sc"].detach(), full["rot"].detach())
e_img = ((out["render"] - img_full).abs().max() / img_full.abs().max()).item()
e_vol = ((vol - vol_full).abs().max() / vol_full.abs().max()).item()
e_g = max(((t[k].grad - full[k].grad[lo:hi]).abs().max() / (full[k].grad.abs().max() + 1e-30)).item() for k in t)
ok = e_img <= 1e-5 and e_vol <= 1e-5 and e_g <= 1e-5
print(f"rank {rank}/{world}: image err {e_img:.2e}, volume err {e_vol:.2e}, shard-gradient err {e_g:.2e} -> {'OK' if ok else 'FAIL'}", flush=True)
enable_peer_exchange(False)
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok else 1)
