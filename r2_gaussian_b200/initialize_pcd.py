"""Initial point cloud for a scene -- the CPU-only plumbing of the reference's `initialize_pcd.py:26-172`
(BASELINE.json configs[0]: `--recon_method random --n_points 1000` on a synthetic phantom).

    python -m r2_gaussian_b200.initialize_pcd --data <scene dir | NAF pickle> [--output init.npy]
        [--recon_method random|volume] [--recon recon.npy] [--n_points 50000] [--density_thresh 0.05]
        [--density_rescale 0.15] [--random_density_max 1.0]

`random` draws positions uniformly in the volume and densities in [0, random_density_max) with numpy's global
generator seeded with 0, exactly like the reference.  The reference's other mode reconstructs a volume with TIGRE's
FDK first; TIGRE is not available here, so `volume` takes that reconstruction from `--recon` (an .npy in the scene's
voxel grid, e.g. produced elsewhere) and then samples it the reference's way.  Writes [n_points, 4] = (x, y, z,
density) in the scene's normalised [-1,1]^3 coordinates to `<scene>/init_<name>.npy` unless `--output` is given.
"""
from __future__ import annotations

import argparse
import os

import numpy as np

from .dataset import init_point_cloud, read_scene
from .trainer import default_init_path


def main(argv=None) -> str:
    ap = argparse.ArgumentParser(description="Generate initialization parameters")
    ap.add_argument("--data", required=True, help="Path to data.")
    ap.add_argument("--output", default=None, help="Path to output.")
    ap.add_argument("--recon_method", default="random", choices=["random", "volume", "fdk"])
    ap.add_argument("--recon", default=None, help="reconstruction volume (.npy) for --recon_method volume")
    ap.add_argument("--n_points", type=int, default=50000)
    ap.add_argument("--density_thresh", type=float, default=0.05)
    ap.add_argument("--density_rescale", type=float, default=0.15)
    ap.add_argument("--random_density_max", type=float, default=1.0)
    a = ap.parse_args(argv)
    if a.recon_method == "fdk":
        raise SystemExit("--recon_method fdk needs TIGRE (absent here): reconstruct elsewhere and pass "
                         "--recon_method volume --recon <vol.npy>")
    np.random.seed(0)                                    # initialize_pcd.py:23
    info = read_scene(os.path.abspath(a.data), eval=False)
    recon = None
    if a.recon_method == "volume":
        if not a.recon:
            raise SystemExit("--recon_method volume needs --recon <vol.npy>")
        recon = np.load(a.recon)
    out = a.output or default_init_path(os.path.abspath(a.data))
    if os.path.exists(out):
        raise SystemExit(f"Initialization file {out} exists! Delete it first.")
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    pts = init_point_cloud(info.scanner_cfg, a.n_points, recon=recon, density_thresh=a.density_thresh,
                           density_rescale=a.density_rescale, random_density_max=a.random_density_max)
    np.save(out, pts)
    print(f"Initialization saved in {out}.")
    return out


if __name__ == "__main__":
    main()
