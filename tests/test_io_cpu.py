"""Scene I/O and evaluation metrics on the CPU: on-disk formats round-trip, geometry derived
from them equals scene.make_view, metrics follow their definitions."""
import math
import pickle

import numpy as np
import pytest
import torch

from r2_gaussian_b200 import dataset, metrics, scene


def _scanner(det=32, vox=16, size=4.0):
    # sVoxel max = 4 -> scene_scale = 0.5
    return {"mode": "cone", "DSD": 14.0, "DSO": 10.0, "nDetector": [det, det], "sDetector": [8.0, 8.0],
            "nVoxel": [vox, vox, vox], "sVoxel": [size, size, size], "offOrigin": [0.0, 0.0, 0.0],
            "offDetector": [0.0, 0.0], "accuracy": 0.5, "totalAngle": 360.0, "startAngle": 0.0, "filter": None}


def test_blender_round_trip_and_geometry(tmp_path):
    rng = np.random.default_rng(0)
    sc = _scanner()
    train = [(a, rng.random((32, 32)).astype(np.float32)) for a in np.linspace(0, 2 * math.pi, 6)[:-1]]
    test = [(0.3, rng.random((32, 32)).astype(np.float32))]
    vol = rng.random((16, 16, 16)).astype(np.float32)
    dataset.write_blender(str(tmp_path / "case"), sc, train, test, vol)
    info = dataset.read_blender(str(tmp_path / "case"), eval=True)
    assert info.scene_scale == 0.5 and info.scanner_cfg["DSO"] == 5.0 and info.scanner_cfg["sVoxel"] == [2.0] * 3
    assert np.allclose(info.scanner_cfg["dVoxel"], [2.0 / 16] * 3) and np.allclose(info.scanner_cfg["dDetector"], [4 / 32] * 2)
    assert len(info.train_cameras) == 5 and len(info.test_cameras) == 1 and info.test_cameras[0].uid == 5
    assert np.allclose(info.train_cameras[2].image, train[2][1] * 0.5)          # projections scale with the scene
    assert np.array_equal(info.vol, vol)
    only_train = dataset.read_blender(str(tmp_path / "case"), eval=False)
    assert len(only_train.test_cameras) == 0
    # the scaled scanner is scene.cone_beam_scanner's geometry: cameras must agree with scene.make_view
    cam = dataset.Camera(info.train_cameras[3], device="cpu")
    ref = scene.make_view(scene.cone_beam_scanner(32, 16), float(train[3][0]))
    assert np.allclose(cam.world_view_transform.numpy(), ref.viewmatrix, atol=1e-6)
    assert np.allclose(cam.full_proj_transform.numpy(), ref.projmatrix, atol=1e-5)
    assert np.allclose(cam.camera_center.numpy(), ref.campos, atol=1e-5)
    assert cam.mode == 1 and abs(cam.FoVx - ref.FoVx) < 1e-12 and cam.image_height == 32
    assert tuple(cam.original_image.shape) == (1, 32, 32)


def test_naf_pickle(tmp_path):
    rng = np.random.default_rng(1)
    n_tr, n_va = 4, 2
    data = {"DSD": 1400.0, "DSO": 1000.0, "nVoxel": [8, 8, 8], "dVoxel": [50.0, 50.0, 50.0], "nDetector": [16, 16],
            "dDetector": [50.0, 50.0], "offOrigin": [0, 0, 0], "offDetector": [0, 0], "totalAngle": 180.0,
            "startAngle": 0.0, "accuracy": 0.5, "mode": "parallel", "numTrain": n_tr, "numVal": n_va,
            "image": rng.random((8, 8, 8)).astype(np.float32),
            "train": {"angles": np.linspace(0, math.pi, n_tr), "projections": rng.random((n_tr, 16, 16)).astype(np.float32)},
            "val": {"angles": np.linspace(0.1, 3.0, n_va), "projections": rng.random((n_va, 16, 16)).astype(np.float32)}}
    p = tmp_path / "scan.pickle"
    with open(p, "wb") as f:
        pickle.dump(data, f)
    info = dataset.read_scene(str(p), eval=True)
    # 8 voxels x 50 mm = 0.4 m -> scale 5: DSO 1 m -> 5
    assert abs(info.scene_scale - 5.0) < 1e-12 and abs(info.scanner_cfg["DSO"] - 5.0) < 1e-12
    assert np.allclose(info.scanner_cfg["sVoxel"], [2.0] * 3) and np.allclose(info.scanner_cfg["sDetector"], [4.0] * 2)
    assert [c.uid for c in info.test_cameras] == [4, 5] and info.test_cameras[0].image_name == "0004"
    assert info.train_cameras[0].mode == 0
    assert np.allclose(info.train_cameras[1].image, data["train"]["projections"][1] * 5.0)
    with pytest.raises(ValueError):
        dataset.read_scene(str(tmp_path / "nothing.txt"))


def test_scene_and_init_point_cloud(tmp_path):
    rng = np.random.default_rng(2)
    sc = _scanner()
    frames = [(a, rng.random((32, 32)).astype(np.float32)) for a in np.linspace(0, 2 * math.pi, 9)[:-1]]
    dataset.write_blender(str(tmp_path / "c"), sc, frames[:6], frames[6:], rng.random((16, 16, 16)).astype(np.float32))
    import random
    random.seed(0)
    s = dataset.Scene(str(tmp_path / "c"), str(tmp_path / "out"), eval=True, shuffle=True, device="cpu")
    assert len(s.getTrainCameras()) == 6 and len(s.getTestCameras()) == 2
    assert torch.allclose(s.bbox, torch.tensor([[-1.0, -1, -1], [1, 1, 1]]))
    assert sorted(c.colmap_id for c in s.getTrainCameras()) == list(range(6))
    assert [c.uid for c in s.getTrainCameras()] == list(range(6))              # uid = position after the shuffle
    # random initialisation = initialize_pcd.py's recipe, i.e. what scene.make_cloud(seed=0) starts from
    np.random.seed(0)
    pc = dataset.init_point_cloud(s.scanner_cfg, 1000)
    want = scene.make_cloud(1000, seed=0)
    assert np.allclose(pc[:, :3], want.means, atol=1e-7) and np.allclose(pc[:, 3:4], want.density, atol=1e-7)
    # from a reconstruction: voxels above the threshold, without replacement, density rescaled
    recon = np.zeros((16, 16, 16), np.float32)
    recon[4:12, 4:12, 4:12] = 0.8
    pc2 = dataset.init_point_cloud(s.scanner_cfg, 200, recon=recon, rng=np.random.RandomState(1))
    assert pc2.shape == (200, 4) and np.allclose(pc2[:, 3], 0.8 * 0.15)
    assert len({tuple(r) for r in pc2[:, :3].round(6)}) == 200
    lo, hi = 4 * 0.125 - 1.0, 11 * 0.125 - 1.0
    assert pc2[:, :3].min() >= lo - 1e-9 and pc2[:, :3].max() <= hi + 1e-9
    with pytest.raises(ValueError):
        dataset.init_point_cloud(s.scanner_cfg, 10 ** 5, recon=recon)


def test_metrics_follow_their_definitions():
    g = torch.Generator().manual_seed(0)
    gt = torch.rand(12, 14, 10, generator=g)
    pred = gt + 0.05 * torch.randn(12, 14, 10, generator=g)
    v, _ = metrics.metric_vol(gt, pred, "psnr")
    assert abs(v - 10 * math.log10(1.0 / ((gt - pred) ** 2).mean().item())) < 1e-4
    v2, _ = metrics.metric_vol(gt, pred, "psnr", pixel_max=None)
    assert abs(v2 - 10 * math.log10(gt.max().item() ** 2 / ((gt - pred) ** 2).mean().item())) < 1e-4
    s_same, per_axis = metrics.metric_vol(gt, gt.clone(), "ssim")
    assert abs(s_same - 1.0) < 1e-5 and len(per_axis) == 3
    s_noisy, _ = metrics.metric_vol(gt, pred, "ssim")
    assert 0.0 < s_noisy < 1.0
    # numpy inputs are accepted like in the reference
    v3, _ = metrics.metric_vol(gt.numpy(), pred.numpy(), "psnr")
    assert abs(v3 - v) < 1e-4
    # metric_proj normalises each slice by its own maximum; an empty ground-truth slice counts as 0 but is skipped in the mean
    gt2 = gt.clone()
    gt2[:, :, 3] = 0
    m, per = metrics.metric_proj(gt2, pred, "psnr", axis=2)
    assert per[3] == 0.0 and len(per) == 10
    want = []
    for i in range(10):
        if i == 3:
            continue
        a, b = gt2[:, :, i] / gt2[:, :, i].max(), pred[:, :, i] / pred[:, :, i].max()
        want.append(10 * math.log10(1.0 / ((a - b) ** 2).mean().item()))
    assert abs(m - sum(want) / 9) < 1e-3
    batch = torch.rand(3, 1, 8, 8, generator=g)
    assert metrics.psnr(batch, batch * 0.9).shape == (3, 1) and metrics.rmse(batch, batch).abs().max() == 0


def test_trainer_settings_and_cli_defaults(tmp_path):
    from r2_gaussian_b200 import trainer
    opt, model = trainer.OptimizationParams(), trainer.ModelParams()
    # the reference's defaults (arguments/__init__.py:21-71)
    assert (opt.iterations, opt.lambda_dssim, opt.lambda_tv, opt.tv_vol_size) == (30000, 0.25, 0.05, 32)
    assert (opt.densify_from_iter, opt.densify_until_iter, opt.densification_interval) == (500, 15000, 100)
    assert (model.scale_min, model.scale_max, opt.densify_grad_threshold) == (0.0005, 0.5, 5e-5)
    cfg = {"sVoxel": [2.0, 2.0, 1.0], "dVoxel": [2 / 64, 2 / 64, 1 / 32]}
    ds = trainer.derived_settings(cfg, model, opt)
    assert np.allclose(ds["scale_bound"], [0.001, 1.0]) and abs(ds["densify_scale_threshold"] - 0.2) < 1e-12
    assert ds["max_scale"] is None and ds["tv_vol_nVoxel"] == [32, 32, 32] and np.allclose(ds["tv_vol_sVoxel"], [1.0, 1.0, 1.0])
    model.scale_min = 0
    assert trainer.derived_settings(cfg, model, opt)["scale_bound"] is None
    d = tmp_path / "chest"
    d.mkdir()
    (d / "meta_data.json").write_text("{}")
    assert trainer.default_init_path(str(d)) == str(d / "init_chest.npy")
    assert trainer.default_init_path("/data/foot_50.pickle") == "/data/init_foot_50.npy"
    with pytest.raises(ValueError):
        trainer.default_init_path("/data/unknown.bin")


def test_initialize_pcd_cli_random_mode(tmp_path):
    """BASELINE configs[0]: `initialize_pcd --recon_method random --n_points 1000` on a 64^3 synthetic phantom,
    CPU only."""
    from r2_gaussian_b200 import initialize_pcd
    rng = np.random.default_rng(3)
    sc = _scanner(det=16, vox=64)
    vol = np.zeros((64, 64, 64), np.float32)
    vol[16:48, 20:44, 24:40] = 0.6
    frames = [(a, rng.random((16, 16)).astype(np.float32)) for a in np.linspace(0, 2 * math.pi, 5)[:-1]]
    case = tmp_path / "phantom"
    dataset.write_blender(str(case), sc, frames, [], vol)
    out = initialize_pcd.main(["--data", str(case), "--recon_method", "random", "--n_points", "1000"])
    assert out == str(case / "init_phantom.npy")
    pts = np.load(out)
    np.random.seed(0)
    want_xyz = 2.0 * (np.random.rand(1000, 3) - 0.5)            # offOrigin 0, scaled sVoxel 2 (initialize_pcd.py:50-58)
    want_rho = np.random.rand(1000)
    assert pts.shape == (1000, 4) and np.allclose(pts[:, :3], want_xyz) and np.allclose(pts[:, 3], want_rho)
    with pytest.raises(SystemExit):
        initialize_pcd.main(["--data", str(case), "--recon_method", "random", "--n_points", "10"])   # file exists
    np.save(tmp_path / "recon.npy", vol)
    out2 = initialize_pcd.main(["--data", str(case), "--recon_method", "volume", "--recon", str(tmp_path / "recon.npy"),
                                "--n_points", "500", "--output", str(tmp_path / "init2.npy")])
    p2 = np.load(out2)
    assert p2.shape == (500, 4) and np.allclose(p2[:, 3], 0.6 * 0.15)
    assert p2[:, 0].min() >= 16 / 32 - 1 - 1e-9 and p2[:, 0].max() <= 47 / 32 - 1 + 1e-9
    with pytest.raises(SystemExit):
        initialize_pcd.main(["--data", str(case), "--recon_method", "fdk", "--output", str(tmp_path / "x.npy")])
