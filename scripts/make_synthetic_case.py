"""Write a synthetic scene in the reference's directory format (meta_data.json + proj_*/ + vol_gt.npy + init_*.npy)
from the Gaussian-mixture phantom of scripts/train_compare.py, projected with this repository's rasterizer
(no TIGRE in this image).  The scanner is written in "physical" units with sVoxel = 4, so the reader's rescaling
to the [-1,1]^3 cube (scene_scale = 0.5) is exercised.

    python scripts/make_synthetic_case.py <out_dir> [--det 128] [--vox 64] [--train 25] [--test 5] [--init 5000]
"""
import argparse
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from r2_gaussian_b200 import dataset, scene  # noqa: E402
from r2_gaussian_b200.render_query import query, render  # noqa: E402
from train_compare import Fixed, phantom  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--det", type=int, default=128)
    ap.add_argument("--vox", type=int, default=64)
    ap.add_argument("--train", type=int, default=25)
    ap.add_argument("--test", type=int, default=5)
    ap.add_argument("--init", type=int, default=5000)
    a = ap.parse_args()
    pipe = types.SimpleNamespace(compute_cov3D_python=False, debug=False)
    truth = Fixed(phantom())
    unit = scene.cone_beam_scanner(a.det, a.vox)                 # the normalised geometry (sVoxel = 2)
    k = 2.0                                                       # physical = normalised * k  ->  scene_scale = 1/k
    scanner = {"mode": "cone", "DSD": unit["DSD"] * k, "DSO": unit["DSO"] * k, "nDetector": [a.det, a.det],
               "sDetector": [s * k for s in unit["sDetector"]], "nVoxel": [a.vox] * 3, "sVoxel": [2.0 * k] * 3,
               "offOrigin": [0.0, 0.0, 0.0], "offDetector": [0.0, 0.0], "accuracy": 0.5, "totalAngle": 360.0,
               "startAngle": 0.0, "filter": None}
    angles = np.linspace(0, 2 * np.pi, a.train + a.test + 1)[:-1]
    test_idx = set(np.linspace(1, len(angles) - 2, a.test).astype(int).tolist())
    frames = {"train": [], "test": []}
    with torch.no_grad():
        for i, ang in enumerate(angles):
            cam = scene.camera_from_view(scene.make_view(unit, float(ang)))
            proj = render(cam, truth, pipe)["render"][0].cpu().numpy() * k      # reader multiplies by 1/k again
            frames["test" if i in test_idx else "train"].append((float(ang), proj))
        vol = query(truth, [0, 0, 0], [a.vox] * 3, [2.0] * 3, pipe)["vol"].cpu().numpy()
    dataset.write_blender(a.out, scanner, frames["train"], frames["test"], vol)
    # initial cloud from a noisy copy of the volume (stands in for the FDK reconstruction of initialize_pcd.py)
    info = dataset.read_blender(a.out, eval=False)
    rng = np.random.RandomState(0)
    recon = np.clip(vol + 0.05 * vol.max() * rng.randn(*vol.shape), 0, None).astype(np.float32)
    pts = dataset.init_point_cloud(info.scanner_cfg, a.init, recon=recon, density_thresh=0.05 * float(recon.max()), rng=rng)
    name = os.path.basename(a.out.rstrip("/"))
    np.save(os.path.join(a.out, f"init_{name}.npy"), pts)
    print(f"wrote {a.out}: {len(frames['train'])} train / {len(frames['test'])} test views, volume {vol.shape}, "
          f"{pts.shape[0]} initial points")


if __name__ == "__main__":
    main()
