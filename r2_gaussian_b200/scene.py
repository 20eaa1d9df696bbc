"""Synthetic CT scenes: scanner geometry, cameras and Gaussian clouds (NumPy, host side).

Restates the reference's geometry conventions so tests and bench.py can build inputs without the
reference or any dataset:
  * scanner = data_generator/synthetic_dataset/scanner/cone_beam.yml:2-33 (DSD 7, DSO 5, 512^2
    detector of size 4x4, volume 2^3 at 256^3); scene scale 2/max(sVoxel) = 1
    (r2_gaussian/dataset/dataset_readers.py:63).
  * camera pose  = angle2pose (dataset_readers.py:156-191), R/T split (:120-127),
    getWorld2View2 + getProjectionMatrix (utils/graphics_utils.py:81-139), matrices stored transposed
    and multiplied as in dataset/cameras.py:66-84 (so the flat arrays are column-major).
  * Gaussians "init-like" = initialize_pcd.py:50-58 (uniform positions / densities, seed 0) with
    isotropic scales from the mean squared distance to the 3 nearest neighbours
    (gaussian/gaussian_model.py:145-156); "trained-like" perturbs scales and rotations.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

MODE_PARALLEL = 0
MODE_CONE = 1


def cone_beam_scanner(n_detector: int = 512, n_voxel: int = 256) -> dict:
    return {
        "mode": "cone", "DSD": 7.0, "DSO": 5.0,
        "nDetector": [n_detector, n_detector], "sDetector": [4.0, 4.0],
        "nVoxel": [n_voxel, n_voxel, n_voxel], "sVoxel": [2.0, 2.0, 2.0],
        "offOrigin": [0.0, 0.0, 0.0], "offDetector": [0.0, 0.0],
    }


def parallel_beam_scanner(n_detector: int = 512, n_voxel: int = 256) -> dict:
    s = cone_beam_scanner(n_detector, n_voxel)
    s["mode"] = "parallel"
    s["sDetector"] = [2.0, 2.0]
    return s


def angle2pose(DSO: float, angle: float) -> np.ndarray:
    """Camera-to-world of the source at `angle` on a circle of radius DSO around z."""
    c1, s1 = math.cos(-math.pi / 2), math.sin(-math.pi / 2)
    R1 = np.array([[1.0, 0.0, 0.0], [0.0, c1, -s1], [0.0, s1, c1]])
    c2, s2 = math.cos(math.pi / 2), math.sin(math.pi / 2)
    R2 = np.array([[c2, -s2, 0.0], [s2, c2, 0.0], [0.0, 0.0, 1.0]])
    ca, sa = math.cos(angle), math.sin(angle)
    R3 = np.array([[ca, -sa, 0.0], [sa, ca, 0.0], [0.0, 0.0, 1.0]])
    T = np.eye(4)
    T[:3, :3] = R3 @ R2 @ R1
    T[:3, 3] = [DSO * ca, DSO * sa, 0.0]
    return T


def projection_matrix(fovx: float, fovy: float, mode: int) -> np.ndarray:
    if mode == MODE_PARALLEL:
        return np.eye(4, dtype=np.float32)
    znear, zfar = 0.01, 100.0
    top = math.tan(fovy / 2) * znear
    right = math.tan(fovx / 2) * znear
    P = np.zeros((4, 4), dtype=np.float32)
    P[0, 0] = 2.0 * znear / (2 * right)
    P[1, 1] = 2.0 * znear / (2 * top)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


@dataclass
class View:
    """One projection geometry in the layout the rasterizer expects."""
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    viewmatrix: np.ndarray   # [4,4] float32 = world->view, TRANSPOSED (flat == column-major)
    projmatrix: np.ndarray   # [4,4] float32 = view @ proj in the transposed convention
    campos: np.ndarray       # [3]
    mode: int
    angle: float = 0.0
    FoVx: float = 0.0
    FoVy: float = 0.0


def make_view(scanner: dict, angle: float) -> View:
    mode = MODE_CONE if scanner["mode"] == "cone" else MODE_PARALLEL
    c2w = angle2pose(scanner["DSO"], angle)
    w2c = np.linalg.inv(c2w)
    R = w2c[:3, :3].T  # stored transposed, dataset_readers.py:123-125
    T = w2c[:3, 3]
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = R.T
    Rt[:3, 3] = T
    Rt[3, 3] = 1.0
    Rt = np.float32(np.linalg.inv(np.linalg.inv(Rt)))  # getWorld2View2 with zero translate, unit scale
    fovx = math.atan2(scanner["sDetector"][1] / 2, scanner["DSD"]) * 2
    fovy = math.atan2(scanner["sDetector"][0] / 2, scanner["DSD"]) * 2
    view_t = np.ascontiguousarray(Rt.T.astype(np.float32))
    proj_t = np.ascontiguousarray(projection_matrix(fovx, fovy, mode).T.astype(np.float32))
    full = (view_t.astype(np.float32) @ proj_t.astype(np.float32)).astype(np.float32)
    campos = np.linalg.inv(view_t.astype(np.float64))[3, :3].astype(np.float32)
    if mode == MODE_PARALLEL:
        tx = ty = 1.0
    else:
        tx, ty = math.tan(fovx * 0.5), math.tan(fovy * 0.5)
    return View(int(scanner["nDetector"][0]), int(scanner["nDetector"][1]), tx, ty, view_t,
                np.ascontiguousarray(full), campos, mode, angle, fovx, fovy)


def camera_from_view(view: View, device="cuda"):
    """The attributes render() reads from the reference's `Camera` (`r2_gaussian/dataset/cameras.py:20-70`:
    world_view_transform, full_proj_transform, camera_center, FoVx/FoVy, image size, mode) as device tensors."""
    import types

    import torch
    return types.SimpleNamespace(
        image_height=view.image_height, image_width=view.image_width, FoVx=view.FoVx, FoVy=view.FoVy, mode=view.mode,
        world_view_transform=torch.tensor(view.viewmatrix, device=device),
        full_proj_transform=torch.tensor(view.projmatrix, device=device),
        camera_center=torch.tensor(view.campos, device=device), angle=view.angle)


def make_views(scanner: dict, n_views: int = 50) -> list[View]:
    """Angles linspace(0, 2pi, n+1)[:-1] (data_generator/synthetic_dataset/generate_data.py:47-50)."""
    angles = np.linspace(0.0, 2.0 * math.pi, n_views + 1)[:-1]
    return [make_view(scanner, float(a)) for a in angles]


@dataclass
class Cloud:
    """Activated Gaussian parameters (what render()/query() hand to the extension)."""
    means: np.ndarray      # [P,3] float32
    scales: np.ndarray     # [P,3] float32
    rotations: np.ndarray  # [P,4] float32 (r,x,y,z), normalised
    density: np.ndarray    # [P,1] float32
    meta: dict = field(default_factory=dict)

    @property
    def P(self) -> int:
        return int(self.means.shape[0])


def knn3_mean_sq_dist(xyz: np.ndarray) -> np.ndarray:
    """simple_knn.distCUDA2 semantics: mean squared distance to the 3 nearest neighbours."""
    from scipy.spatial import cKDTree

    tree = cKDTree(xyz.astype(np.float64))
    k = min(4, xyz.shape[0])
    d, _ = tree.query(xyz.astype(np.float64), k=k, workers=-1)
    d = np.atleast_2d(d)
    if k < 2:
        return np.zeros(xyz.shape[0], dtype=np.float32)
    return (d[:, 1:] ** 2).mean(axis=1).astype(np.float32)


def make_cloud(P: int, kind: str = "init", seed: int = 0, s_voxel=(2.0, 2.0, 2.0), density_scale: float = 1.0,
               scale_bound=(0.001, 1.0)) -> Cloud:
    rng = np.random.RandomState(seed)
    s_voxel = np.asarray(s_voxel, dtype=np.float64)
    xyz = (s_voxel * (rng.rand(P, 3) - 0.5)).astype(np.float32)
    dens = (rng.rand(P, 1) * density_scale).astype(np.float32)
    dist2 = np.maximum(knn3_mean_sq_dist(xyz), 1e-6)
    iso = np.sqrt(dist2).astype(np.float32)
    # scale_bound = [scale_min, scale_max] * max(sVoxel) = [0.001, 1.0] for the 2^3 volume (train.py:59-61);
    # create_from_pcd clamps to [lo + EPS, hi - EPS] (gaussian_model.py:152-155)
    lo, hi = scale_bound[0] + 1e-5, scale_bound[1] - 1e-5
    iso = np.clip(iso, lo, hi)
    scales = np.repeat(iso[:, None], 3, axis=1).astype(np.float32)
    rots = np.zeros((P, 4), dtype=np.float32)
    rots[:, 0] = 1.0
    if kind == "trained":
        scales = (scales * rng.uniform(0.5, 1.5, size=(P, 3))).astype(np.float32)
        scales = np.clip(scales, lo, hi).astype(np.float32)
        q = rng.randn(P, 4)
        q /= np.linalg.norm(q, axis=1, keepdims=True)
        rots = q.astype(np.float32)
    elif kind != "init":
        raise ValueError(f"unknown cloud kind {kind!r}")
    return Cloud(xyz, scales, rots, dens, {"kind": kind, "seed": seed})


def shard_cloud(cloud: Cloud, rank: int, world: int) -> Cloud:
    """Contiguous index partition of the Gaussians: rank r owns [r*P/world, (r+1)*P/world)."""
    P = cloud.P
    lo = (P * rank) // world
    hi = (P * (rank + 1)) // world
    return Cloud(cloud.means[lo:hi].copy(), cloud.scales[lo:hi].copy(), cloud.rotations[lo:hi].copy(),
                 cloud.density[lo:hi].copy(), dict(cloud.meta, shard=(rank, world)))
