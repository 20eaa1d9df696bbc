"""Golden vectors for the Python-side pieces, produced by IMPORTING THE REFERENCE'S OWN MODULES in the build
container (CPU; /root/reference is not available on the GPU box, so the outputs are committed):

    python tests/golden/make_golden_host.py          # writes tests/golden/host_golden.npz

Covers `utils/loss_utils.py` (l1_loss, ssim, tv_3d_loss + autograd gradients), `utils/image_utils.py` (psnr,
metric_vol, metric_proj), `utils/gaussian_utils.py` (inverse activations, lr schedule), `utils/graphics_utils.py`
(getWorld2View2, getProjectionMatrix) and `dataset/dataset_readers.py` (angle2pose, readCTameras on a scene written by
r2_gaussian_b200.dataset.write_blender).  `plyfile` (absent from this image, only needed by unrelated helpers of
those modules) is stubbed for the import.
"""
import math
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
stub = types.ModuleType("plyfile")
stub.PlyData = stub.PlyElement = object
sys.modules.setdefault("plyfile", stub)

from r2_gaussian.utils import gaussian_utils as ref_gu  # noqa: E402
from r2_gaussian.utils import graphics_utils as ref_gr  # noqa: E402
from r2_gaussian.utils import image_utils as ref_iu  # noqa: E402
from r2_gaussian.utils import loss_utils as ref_lu  # noqa: E402
from r2_gaussian.dataset import dataset_readers as ref_dr  # noqa: E402

from r2_gaussian_b200 import dataset  # noqa: E402

out = {}
g = torch.Generator().manual_seed(20240924)


def images(H, W):
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, W), indexing="ij")
    base = torch.exp(-(xx ** 2 + yy ** 2) * 3) * 2.0
    a = (base + 0.15 * torch.rand(H, W, generator=g)).unsqueeze(0)
    b = (base * 0.9 + 0.1 * torch.rand(H, W, generator=g)).unsqueeze(0)
    return a, b


# ---- losses: float64 evaluation of the reference code = truth; float32 evaluation = what the reference itself gets
for tag, (H, W) in {"a": (50, 77), "b": (64, 64), "c": (11, 5)}.items():
    a, b = images(H, W)
    out[f"loss_{tag}_img"], out[f"loss_{tag}_gt"] = a.numpy(), b.numpy()
    for dt, name in ((torch.float64, "f64"), (torch.float32, "f32")):
        x = a.to(dt).requires_grad_(True)
        y = b.to(dt)
        l1 = ref_lu.l1_loss(x, y)
        ss = ref_lu.ssim(x, y)
        total = l1 + 0.25 * (1.0 - ss)
        total.backward()
        out[f"loss_{tag}_{name}"] = np.array([l1.item(), ss.item(), total.item()], np.float64)
        out[f"loss_{tag}_grad_{name}"] = x.grad.double().numpy()

for tag, shape in {"a": (7, 9, 5), "b": (16, 16, 16)}.items():
    v = torch.rand(*shape, generator=g)
    v[v < 0.2] = 0.0
    out[f"tv_{tag}_vol"] = v.numpy()
    for red in ("sum", "mean"):
        x = v.double().requires_grad_(True)
        t = ref_lu.tv_3d_loss(x, reduction=red)
        t.backward()
        out[f"tv_{tag}_{red}"] = np.array([t.item()])
        out[f"tv_{tag}_{red}_grad"] = x.grad.numpy()

# ---- metrics
gt = torch.rand(12, 14, 10, generator=g)
gt[:, :, 3] = 0
pred = (gt + 0.05 * torch.randn(12, 14, 10, generator=g)).clamp_min(0)
out["metric_gt"], out["metric_pred"] = gt.numpy(), pred.numpy()
out["metric_vol_psnr"] = np.array([ref_iu.metric_vol(gt, pred, "psnr")[0]])
out["metric_vol_psnr_max"] = np.array([ref_iu.metric_vol(gt, pred, "psnr", pixel_max=None)[0]])
sv, per_axis = ref_iu.metric_vol(gt, pred, "ssim")
out["metric_vol_ssim"] = np.array([sv] + list(per_axis))
pp, pper = ref_iu.metric_proj(gt, pred, "psnr")
out["metric_proj_psnr"] = np.array([pp] + [float(v) for v in pper])
ps, sper = ref_iu.metric_proj(gt, pred, "ssim", axis=0)
out["metric_proj_ssim_axis0"] = np.array([ps] + [float(v) for v in sper])
batch = torch.rand(3, 1, 8, 9, generator=g)
out["psnr_batch_in"] = batch.numpy()
out["psnr_batch"] = ref_iu.psnr(batch, batch * 0.9 + 0.01).numpy()

# ---- small helpers
x = torch.rand(64, generator=g) * 3 + 1e-3
out["act_in"] = x.numpy()
out["inverse_softplus"] = ref_gu.inverse_softplus(x).numpy()
p = torch.rand(64, generator=g) * 0.98 + 0.01
out["sig_in"] = p.numpy()
out["inverse_sigmoid"] = ref_gu.inverse_sigmoid(p).numpy()
steps = np.array([0, 1, 10, 499, 500, 15000, 29999, 30000, 40000])
f = ref_gu.get_expon_lr_func(lr_init=2e-4, lr_final=2e-5, max_steps=30000)
out["lr_steps"] = steps
out["lr_values"] = np.array([f(int(s)) for s in steps])
f2 = ref_gu.get_expon_lr_func(lr_init=1e-2, lr_final=1e-3, lr_delay_steps=1000, lr_delay_mult=0.01, max_steps=30000)
out["lr_values_delay"] = np.array([f2(int(s)) for s in steps])

# ---- geometry
angles = np.array([0.0, 0.37, 1.5, math.pi, 4.4, 6.0])
out["angles"] = angles
out["angle2pose"] = np.stack([ref_dr.angle2pose(5.0, a) for a in angles])
cfg = {"sVoxel": [2.0, 2.0, 2.0], "DSO": 5.0}
fov = 2 * math.atan2(2.0, 7.0)
out["fov"] = np.array([fov])
out["proj_cone"] = ref_gr.getProjectionMatrix(fov, fov * 0.9, 1, cfg).numpy()
out["proj_parallel"] = ref_gr.getProjectionMatrix(fov, fov, 0, cfg).numpy()
w2c = np.linalg.inv(ref_dr.angle2pose(5.0, 0.37))
out["world2view2"] = ref_gr.getWorld2View2(np.transpose(w2c[:3, :3]), w2c[:3, 3])

# ---- scene reader: a scene written by OUR writer, read by THEIR camera reader (after the reference's rescaling)
rng = np.random.default_rng(5)
scanner = {"mode": "cone", "DSD": 14.0, "DSO": 10.0, "nDetector": [24, 32], "sDetector": [6.0, 8.0],
           "nVoxel": [8, 8, 8], "sVoxel": [4.0, 4.0, 4.0], "offOrigin": [0.0, 0.0, 0.0], "offDetector": [0.0, 0.0],
           "accuracy": 0.5, "totalAngle": 360.0, "startAngle": 0.0, "filter": None}
train = [(float(a), rng.random((24, 32)).astype(np.float32)) for a in (0.1, 1.3, 2.9)]
test = [(0.7, rng.random((24, 32)).astype(np.float32))]
vol = rng.random((8, 8, 8)).astype(np.float32)
with tempfile.TemporaryDirectory() as tmp:
    case = os.path.join(tmp, "case")
    dataset.write_blender(case, scanner, train, test, vol)
    import json
    with open(os.path.join(case, "meta_data.json")) as fh:
        meta = json.load(fh)
    # readBlenderInfo's rescaling (dataset_readers.py:50-77), without its final .cuda()
    sc = meta["scanner"]
    sc["dVoxel"] = list(np.array(sc["sVoxel"]) / np.array(sc["nVoxel"]))
    sc["dDetector"] = list(np.array(sc["sDetector"]) / np.array(sc["nDetector"]))
    scale = 2 / max(sc["sVoxel"])
    for k in ["dVoxel", "sVoxel", "sDetector", "dDetector", "offOrigin", "offDetector", "DSD", "DSO"]:
        sc[k] = (np.array(sc[k]) * scale).tolist()
    cams = ref_dr.readCTameras(meta, case, True, scale)
    print()
out["reader_scanner_json"] = np.frombuffer(json.dumps(scanner).encode(), dtype=np.uint8)
for i, (a, p_) in enumerate(train):
    out[f"reader_train_{i}"] = p_
out["reader_train_angles"] = np.array([a for a, _ in train])
out["reader_test_0"], out["reader_test_angle"], out["reader_vol"] = test[0][1], np.array([test[0][0]]), vol
for split in ("train", "test"):
    for i, c in enumerate(cams[split]):
        out[f"reader_{split}_{i}_R"], out[f"reader_{split}_{i}_T"] = c.R, c.T
        out[f"reader_{split}_{i}_fov"] = np.array([c.FovX, c.FovY])
        out[f"reader_{split}_{i}_image"] = c.image
        out[f"reader_{split}_{i}_meta"] = np.array([c.uid, c.width, c.height, c.mode])
out["reader_scale"] = np.array([scale])

np.savez_compressed(os.path.join(HERE, "host_golden.npz"), **out)
print("wrote", os.path.join(HERE, "host_golden.npz"), len(out), "arrays")
