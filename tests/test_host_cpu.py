"""Host-side logic that needs no GPU: scene geometry, the Python surface (names / fields / argument
validation identical to the reference's), sharding arithmetic, and the no-fallback rule."""
import math
import os

import numpy as np
import pytest
import torch

from r2_gaussian_b200 import scene
from r2_gaussian_b200.rasterization import GaussianRasterizationSettings, GaussianRasterizer
from r2_gaussian_b200.sharded import shard_bounds
from r2_gaussian_b200.voxelization import GaussianVoxelizationSettings, GaussianVoxelizer


def test_settings_fields_match_reference():
    # PYX/rasterization.py:200-211 and PYX/voxelization.py:26-38
    assert GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "scale_modifier", "viewmatrix", "projmatrix", "campos",
        "prefiltered", "mode", "debug")
    assert GaussianVoxelizationSettings._fields == (
        "scale_modifier", "nVoxel_x", "nVoxel_y", "nVoxel_z", "sVoxel_x", "sVoxel_y", "sVoxel_z", "center_x",
        "center_y", "center_z", "prefiltered", "debug")


def test_drop_in_package_exports():
    import xray_gaussian_rasterization_voxelization as pkg

    for n in ["GaussianRasterizationSettings", "GaussianRasterizer", "GaussianVoxelizationSettings", "GaussianVoxelizer"]:
        assert hasattr(pkg, n)
    for n in ["rasterize_gaussians", "rasterize_gaussians_backward", "voxelize_gaussians", "voxelize_gaussians_backward",
              "mark_visible"]:
        assert callable(getattr(pkg._C, n))


def test_exactly_one_covariance_source():
    s = GaussianRasterizationSettings(16, 16, 1.0, 1.0, 1.0, torch.eye(4), torch.eye(4), torch.zeros(3), False, 1, False)
    r = GaussianRasterizer(s)
    m = torch.zeros(4, 3)
    with pytest.raises(Exception, match="exactly one"):
        r(m, m, torch.ones(4, 1))                                     # neither
    with pytest.raises(Exception, match="exactly one"):
        r(m, m, torch.ones(4, 1), scales=torch.ones(4, 3), rotations=torch.ones(4, 4), cov3D_precomp=torch.ones(4, 6))
    with pytest.raises(Exception, match="exactly one"):
        r(m, m, torch.ones(4, 1), scales=torch.ones(4, 3))            # scales without rotations
    v = GaussianVoxelizer(GaussianVoxelizationSettings(1.0, 8, 8, 8, 1.0, 1.0, 1.0, 0.0, 0.0, 0.0, False, False))
    with pytest.raises(Exception, match="exactly one"):
        v(m, torch.ones(4, 1))


def test_no_cpu_fallback():
    s = GaussianRasterizationSettings(16, 16, 1.0, 1.0, 1.0, torch.eye(4), torch.eye(4), torch.zeros(3), False, 1, False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        GaussianRasterizer(s)(torch.zeros(4, 3), torch.zeros(4, 3), torch.ones(4, 1), torch.ones(4, 3), torch.ones(4, 4))


def test_cone_beam_view_geometry():
    sc = scene.cone_beam_scanner(512, 256)
    for ang in [0.0, 0.7, 3.0]:
        v = scene.make_view(sc, ang)
        V = v.viewmatrix.T.astype(np.float64)             # proper world->view
        np.testing.assert_allclose(V[:3, :3] @ V[:3, :3].T, np.eye(3), atol=1e-6)
        assert abs(np.linalg.norm(v.campos) - sc["DSO"]) < 1e-5
        assert abs(v.tanfovx - (sc["sDetector"][1] / 2) / sc["DSD"]) < 1e-12
        # the volume centre projects to the detector centre, at depth DSO
        p = np.array([0, 0, 0, 1.0]) @ v.viewmatrix.astype(np.float64)
        assert abs(p[2] - sc["DSO"]) < 1e-5 and abs(p[0]) < 1e-5 and abs(p[1]) < 1e-5
        h = np.array([0, 0, 0, 1.0]) @ v.projmatrix.astype(np.float64)
        assert abs(h[0] / h[3]) < 1e-5 and abs(h[1] / h[3]) < 1e-5
    vp = scene.make_view(scene.parallel_beam_scanner(64, 32), 1.0)
    assert vp.mode == 0 and vp.tanfovx == 1.0
    np.testing.assert_allclose(vp.projmatrix, vp.viewmatrix, atol=1e-7)   # projection = identity for parallel beam


def test_cloud_generation_matches_reference_recipe():
    c = scene.make_cloud(2000, kind="init", seed=0)
    rng = np.random.RandomState(0)
    np.testing.assert_array_equal(c.means, (2.0 * (rng.rand(2000, 3) - 0.5)).astype(np.float32))
    assert np.all(c.rotations == np.array([1, 0, 0, 0], np.float32))
    assert np.all(c.scales[:, 0] == c.scales[:, 1]) and c.scales.min() >= 0.001 and c.scales.max() <= 1.0
    # brute-force 3-NN check on a few points
    d2 = ((c.means[:50, None, :].astype(np.float64) - c.means[None].astype(np.float64)) ** 2).sum(-1)
    d2.sort(axis=1)
    np.testing.assert_allclose(np.sqrt(np.maximum(d2[:, 1:4].mean(1), 1e-6)), c.scales[:50, 0], rtol=1e-5)
    t = scene.make_cloud(500, kind="trained", seed=3)
    np.testing.assert_allclose(np.linalg.norm(t.rotations, axis=1), 1.0, atol=1e-6)


@pytest.mark.parametrize("P,world", [(100000, 8), (7, 3), (5, 8), (0, 2)])
def test_shard_bounds_partition(P, world):
    b = [shard_bounds(P, r, world) for r in range(world)]
    assert b[0][0] == 0 and b[-1][1] == P
    assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
    sizes = [hi - lo for lo, hi in b]
    assert max(sizes) - min(sizes) <= 1


def test_render_query_surface_with_stub_model():
    """render()/query() accept the reference's duck-typed model / camera / pipe objects and reject CPU tensors
    loudly instead of falling back."""
    from types import SimpleNamespace

    from r2_gaussian_b200.render_query import query, render

    pc = SimpleNamespace(get_xyz=torch.zeros(3, 3), get_density=torch.ones(3, 1), get_scaling=torch.ones(3, 3) * 0.1,
                         get_rotation=torch.tensor([[1.0, 0, 0, 0]] * 3))
    cam = SimpleNamespace(image_height=16, image_width=16, FoVx=0.5, FoVy=0.5, mode=1, world_view_transform=torch.eye(4),
                          full_proj_transform=torch.eye(4), camera_center=torch.zeros(3))
    pipe = SimpleNamespace(debug=False, compute_cov3D_python=False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        render(cam, pc, pipe)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        query(pc, [0, 0, 0], [8, 8, 8], [1, 1, 1], pipe)
    cam.mode = 2
    with pytest.raises(ValueError):
        render(cam, pc, pipe)


def test_gaussian_utils_match_their_definitions():
    """Helpers behind GaussianModel (r2_gaussian/utils/gaussian_utils.py:5-90): pure torch, checked on the CPU."""
    import math

    from r2_gaussian_b200 import gaussian_utils as gu
    g = torch.Generator().manual_seed(0)
    q = torch.randn(64, 4, generator=g)
    R = gu.build_rotation(q)
    eye = torch.eye(3).expand(64, 3, 3)
    assert torch.allclose(R @ R.transpose(1, 2), eye, atol=1e-5) and torch.allclose(torch.linalg.det(R), torch.ones(64), atol=1e-5)
    assert torch.allclose(gu.build_rotation(3.0 * q), R, atol=1e-6)              # normalises its input
    ident = gu.build_rotation(torch.tensor([[1.0, 0, 0, 0]]))
    assert torch.equal(ident[0], torch.eye(3))
    # 90 degrees about z: x -> y
    rz = gu.build_rotation(torch.tensor([[math.cos(math.pi / 4), 0, 0, math.sin(math.pi / 4)]]))
    assert torch.allclose(rz[0] @ torch.tensor([1.0, 0, 0]), torch.tensor([0.0, 1.0, 0]), atol=1e-6)
    s = torch.rand(64, 3, generator=g) + 0.1
    L = gu.build_scaling_rotation(s, q)
    assert torch.allclose(L, R @ torch.diag_embed(s), atol=1e-6)
    cov = L @ L.transpose(1, 2)
    six = gu.strip_symmetric(cov)
    assert torch.equal(six[:, 0], cov[:, 0, 0]) and torch.equal(six[:, 4], cov[:, 1, 2]) and six.shape == (64, 6)
    x = torch.rand(100, generator=g) * 3 + 1e-3
    assert torch.allclose(torch.nn.functional.softplus(gu.inverse_softplus(x)), x, rtol=1e-5, atol=1e-6)
    p = torch.rand(100, generator=g) * 0.98 + 0.01
    assert torch.allclose(torch.sigmoid(gu.inverse_sigmoid(p)), p, atol=1e-6)
    f = gu.get_expon_lr_func(2e-4, 2e-6, max_steps=1000)
    assert abs(f(0) - 2e-4) < 1e-18 and abs(f(1000) - 2e-6) < 1e-18 and abs(f(5000) - 2e-6) < 1e-18
    assert abs(f(500) - math.sqrt(2e-4 * 2e-6)) < 1e-15 and f(-1) == 0.0
    assert gu.get_expon_lr_func(0.0, 0.0)(10) == 0.0
    d = gu.get_expon_lr_func(1e-2, 1e-2, lr_delay_steps=100, lr_delay_mult=0.1, max_steps=1000)
    assert abs(d(0) - 1e-3) < 1e-12 and abs(d(100) - 1e-2) < 1e-12 and d(50) < 1e-2


def test_bench_reference_arm_json_contract_on_cpu():
    """`bench.py --impl reference` without a GPU falls back to the CPU oracle and still prints the contract's line."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--gaussians", "3000", "--detector", "64", "--views", "2"], capture_output=True, text=True,
                       timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["metric"] == "projections_per_sec" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
