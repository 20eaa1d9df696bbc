"""Python-side mirrors against golden vectors produced by the REFERENCE'S OWN modules
(tests/golden/make_golden_host.py, run in the build container where /root/reference exists)."""
import json
import math
import os

import numpy as np
import pytest
import torch

from r2_gaussian_b200 import dataset, gaussian_utils as gu, metrics, scene

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "host_golden.npz"))


def test_metrics_equal_the_reference():
    gt, pred = torch.from_numpy(G["metric_gt"]), torch.from_numpy(G["metric_pred"])
    assert abs(metrics.metric_vol(gt, pred, "psnr")[0] - G["metric_vol_psnr"][0]) < 1e-4
    assert abs(metrics.metric_vol(gt, pred, "psnr", pixel_max=None)[0] - G["metric_vol_psnr_max"][0]) < 1e-4
    v, per = metrics.metric_vol(gt, pred, "ssim")
    assert np.allclose([v] + per, G["metric_vol_ssim"], atol=2e-6)
    v, per = metrics.metric_proj(gt, pred, "psnr")
    assert np.allclose([v] + per, G["metric_proj_psnr"], rtol=1e-5, atol=1e-4)
    v, per = metrics.metric_proj(gt, pred, "ssim", axis=0)
    assert np.allclose([v] + per, G["metric_proj_ssim_axis0"], atol=2e-6)
    b = torch.from_numpy(G["psnr_batch_in"])
    assert np.allclose(metrics.psnr(b, b * 0.9 + 0.01).numpy(), G["psnr_batch"], rtol=1e-6)
    # the (CPU-capable) SSIM used by the metrics is the reference's
    for tag in "abc":
        a, g = torch.from_numpy(G[f"loss_{tag}_img"]), torch.from_numpy(G[f"loss_{tag}_gt"])
        assert abs(float(metrics.ssim(a.double(), g.double())) - G[f"loss_{tag}_f64"][1]) < 1e-9


def test_helpers_equal_the_reference():
    assert np.allclose(gu.inverse_softplus(torch.from_numpy(G["act_in"])).numpy(), G["inverse_softplus"], rtol=1e-6, atol=1e-7)
    assert np.allclose(gu.inverse_sigmoid(torch.from_numpy(G["sig_in"])).numpy(), G["inverse_sigmoid"], rtol=1e-6, atol=1e-7)
    f = gu.get_expon_lr_func(lr_init=2e-4, lr_final=2e-5, max_steps=30000)
    assert np.allclose([f(int(s)) for s in G["lr_steps"]], G["lr_values"], rtol=1e-12)
    f2 = gu.get_expon_lr_func(lr_init=1e-2, lr_final=1e-3, lr_delay_steps=1000, lr_delay_mult=0.01, max_steps=30000)
    assert np.allclose([f2(int(s)) for s in G["lr_steps"]], G["lr_values_delay"], rtol=1e-12)


def test_geometry_equals_the_reference():
    for a, want in zip(G["angles"], G["angle2pose"]):
        assert np.allclose(scene.angle2pose(5.0, float(a)), want, atol=1e-15)
    fov = float(G["fov"][0])
    assert np.allclose(scene.projection_matrix(fov, fov * 0.9, 1), G["proj_cone"], atol=1e-7)
    assert np.allclose(scene.projection_matrix(fov, fov, 0), G["proj_parallel"])
    # getWorld2View2 of the pose at angle 0.37 == the (untransposed) view matrix of scene.make_view
    v = scene.make_view(scene.cone_beam_scanner(64, 32), 0.37)
    assert np.allclose(v.viewmatrix.T, G["world2view2"], atol=1e-6)


def test_scene_reader_equals_the_reference(tmp_path):
    scanner = json.loads(bytes(G["reader_scanner_json"]).decode())
    train = [(float(a), G[f"reader_train_{i}"]) for i, a in enumerate(G["reader_train_angles"])]
    test = [(float(G["reader_test_angle"][0]), G["reader_test_0"])]
    dataset.write_blender(str(tmp_path / "case"), scanner, train, test, G["reader_vol"])
    info = dataset.read_blender(str(tmp_path / "case"), eval=True)
    assert abs(info.scene_scale - float(G["reader_scale"][0])) < 1e-15
    for split, cams in (("train", info.train_cameras), ("test", info.test_cameras)):
        for i, c in enumerate(cams):
            assert np.allclose(c.R, G[f"reader_{split}_{i}_R"], atol=1e-15) and np.allclose(c.T, G[f"reader_{split}_{i}_T"], atol=1e-15)
            assert np.allclose([c.FovX, c.FovY], G[f"reader_{split}_{i}_fov"], atol=1e-15)
            assert np.allclose(c.image, G[f"reader_{split}_{i}_image"], atol=0)
            assert [c.uid, c.width, c.height, c.mode] == G[f"reader_{split}_{i}_meta"].tolist()
    # non-square detector: width = nDetector[1], height = nDetector[0]
    cam = dataset.Camera(info.train_cameras[0], device="cpu")
    assert (cam.image_height, cam.image_width) == (24, 32)
