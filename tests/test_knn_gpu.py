"""distCUDA2 replacement (r2x_knn3_mean_dist2) vs the brute-force oracle: bit-exact, every cloud shape that
stresses the grid search (clusters, outliers, duplicates, flat and collinear clouds, tiny P)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import r2_oracle as orc

pytestmark = pytest.mark.gpu


def _clouds():
    rng = np.random.default_rng(7)
    out = {}
    out["uniform_20k"] = rng.uniform(-1, 1, size=(20000, 3))
    out["gauss_5k"] = rng.normal(size=(5000, 3))
    c = rng.normal(size=(6000, 3)) * 0.01 + rng.integers(0, 3, size=(6000, 1)) * 5.0
    out["clusters_far_apart"] = c
    o = rng.normal(size=(3000, 3))
    o[:5] *= 1e4                                   # a few extreme outliers blow up the bounding box
    out["outliers"] = o
    d = rng.normal(size=(2000, 3))
    d[1000:] = d[:1000]                            # every point duplicated once
    out["duplicates"] = d
    f = rng.uniform(-1, 1, size=(4000, 3))
    f[:, 2] = 0.25                                 # flat cloud: one extent is exactly zero
    out["flat"] = f
    l = np.zeros((500, 3))
    l[:, 0] = np.linspace(0, 1, 500)               # collinear
    out["line"] = l
    out["offset_far_from_origin"] = rng.normal(size=(3000, 3)) * 0.05 + 1000.0
    for n in (1, 2, 3, 4, 5, 33):
        out[f"tiny_{n}"] = rng.normal(size=(n, 3))
    out["all_same_point"] = np.ones((64, 3))
    return {k: np.ascontiguousarray(v, np.float32) for k, v in out.items()}


@pytest.mark.parametrize("name", sorted(_clouds().keys()))
def test_knn_matches_oracle(name):
    from simple_knn._C import distCUDA2
    pts = _clouds()[name]
    got = distCUDA2(torch.from_numpy(pts).cuda()).cpu().numpy()
    ref = orc.knn3_mean_dist2(pts)
    assert got.shape == ref.shape
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), (
        name, int((got != ref).sum()), float(np.abs(got - ref).max()))


def test_knn_large_and_reproducible():
    from simple_knn._C import distCUDA2
    rng = np.random.default_rng(3)
    pts = torch.from_numpy(rng.normal(size=(300000, 3)).astype(np.float32)).cuda()
    a = distCUDA2(pts)
    b = distCUDA2(pts)
    assert torch.equal(a, b)
    # size-independent property: permuting the cloud permutes the answer
    perm = torch.randperm(pts.shape[0], device="cuda")
    c = distCUDA2(pts[perm])
    assert torch.equal(c, a[perm])
    assert bool((a > 0).all()) and bool(torch.isfinite(a).all())


def test_knn_rejects_cpu_tensor():
    from simple_knn._C import distCUDA2
    with pytest.raises(RuntimeError):
        distCUDA2(torch.zeros(10, 3))
