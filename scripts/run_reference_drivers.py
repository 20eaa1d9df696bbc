#!/usr/bin/env python
"""Run the REFERENCE's own train.py and test.py, unchanged, on this repository's drop-in packages.

The reference's host code imports exactly two native packages:
    r2_gaussian/gaussian/render_query.py:14-19   from xray_gaussian_rasterization_voxelization import (...)
    r2_gaussian/gaussian/gaussian_model.py:21    from simple_knn._C import distCUDA2
Both names resolve to this repository (xray_gaussian_rasterization_voxelization/, simple_knn/) once its root is on
PYTHONPATH.  Nothing of the reference is edited: `--prepare` copies r2_gaussian/ (without its CUDA submodules),
train.py and test.py from /root/reference into the git-ignored baseline/_ref/ (it travels to the GPU box with the
gpurun snapshot); the only additions are import placeholders for third-party modules that this image does not have
(scripts/ref_shims: matplotlib, open3d, scikit-image, plyfile, SimpleITK -- plotting / .ply / .nii.gz helpers the
drivers import at module load).

Two arms run the same commands on the same synthetic case (scripts/make_synthetic_case.py):
    ours        PYTHONPATH = <repo>                      -> our CUDA kernels behind the reference's Python surface
    refkernels  PYTHONPATH = baseline/_ref/refkernels:<repo>  -> the reference's own Python package (PYX/*.py, copied
                unchanged) over its own CUDA kernels compiled into oracle/_ref/libr2ref.so (oracle/ref_C.py plays
                the pybind module); simple_knn is ours in both arms (its upstream source is an absent submodule).
Acceptance: both complete train.py and test.py, and the 3-D PSNR of the two arms agrees to 0.2 dB.

    python scripts/run_reference_drivers.py --prepare                      # build container (needs /root/reference)
    python scripts/run_reference_drivers.py --run [--iterations 1500] [--out gpurun_out/refdrv]   # GPU box
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("R2_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
SHIMS = os.path.join(ROOT, "scripts", "ref_shims")
PYX = "r2_gaussian/submodules/xray-gaussian-rasterization-voxelization/xray_gaussian_rasterization_voxelization"


def prepare():
    if not os.path.isdir(os.path.join(REF, "r2_gaussian")):
        print(f"run_reference_drivers: {REF} not present; keeping {DST} as it is", file=sys.stderr)
        return os.path.isdir(os.path.join(DST, "r2_gaussian"))
    os.makedirs(DST, exist_ok=True)
    dst_pkg = os.path.join(DST, "r2_gaussian")
    if os.path.isdir(dst_pkg):
        shutil.rmtree(dst_pkg)
    shutil.copytree(os.path.join(REF, "r2_gaussian"), dst_pkg,
                    ignore=shutil.ignore_patterns("submodules", "__pycache__", "*.pyc"))
    for f in ("train.py", "test.py"):
        shutil.copy2(os.path.join(REF, f), os.path.join(DST, f))
    # the reference's own Python package over its own kernels (second arm)
    rk = os.path.join(DST, "refkernels", "xray_gaussian_rasterization_voxelization")
    if os.path.isdir(rk):
        shutil.rmtree(rk)
    os.makedirs(rk)
    for f in ("__init__.py", "rasterization.py", "voxelization.py"):
        shutil.copy2(os.path.join(REF, PYX, f), os.path.join(rk, f))
    with open(os.path.join(rk, "_C.py"), "w") as f:
        f.write("# written by scripts/run_reference_drivers.py: the reference's pybind module, played by ctypes over\n"
                "# oracle/_ref/libr2ref.so (the reference's CUDA sources, compiled unmodified)\n"
                "from oracle.ref_C import *  # noqa: F401,F403\n")
    print(f"run_reference_drivers: copied the reference's host code into {DST}")
    return True


def _run(cmd, env, cwd, log):
    t0 = time.time()
    with open(log, "w") as f:
        r = subprocess.run(cmd, env=env, cwd=cwd, stdout=f, stderr=subprocess.STDOUT)
    return r.returncode, time.time() - t0


def _yaml(path):
    import yaml

    with open(path) as f:
        return yaml.safe_load(f)


def run(args):
    out = os.path.abspath(args.out)
    os.makedirs(out, exist_ok=True)
    if not os.path.isdir(os.path.join(DST, "r2_gaussian")):
        raise SystemExit(f"{DST} is empty: run `python scripts/run_reference_drivers.py --prepare` in the build container")
    case = os.path.join(out, "case_phantom")
    if not os.path.exists(os.path.join(case, "meta_data.json")):
        subprocess.check_call([sys.executable, os.path.join(ROOT, "scripts", "make_synthetic_case.py"), case,
                               "--det", str(args.det), "--vox", str(args.vox), "--train", str(args.train_views),
                               "--test", str(args.test_views), "--init", str(args.init_points)])
    summary = {"case": {"detector": args.det, "volume": args.vox, "train_views": args.train_views,
                        "test_views": args.test_views, "init_points": args.init_points},
               "iterations": args.iterations, "arms": {}}
    it = args.iterations
    for arm in args.arms.split(","):
        model = os.path.join(out, f"model_{arm}")
        shutil.rmtree(model, ignore_errors=True)
        pp = [ROOT, SHIMS]
        if arm == "refkernels":
            pp.insert(0, os.path.join(DST, "refkernels"))
        env = dict(os.environ, PYTHONPATH=os.pathsep.join(pp))
        train_cmd = [sys.executable, "train.py", "-s", case, "-m", model, "--iterations", str(it),
                     "--densify_from_iter", str(args.densify_from), "--densify_until_iter", str(args.densify_until),
                     "--test_iterations", str(it), "--quiet"]
        rc_t, dt_t = _run(train_cmd, env, DST, os.path.join(out, f"train_{arm}.log"))
        rec = {"train_rc": rc_t, "train_seconds": dt_t, "train_cmd": " ".join(train_cmd[1:])}
        if rc_t == 0:
            test_cmd = [sys.executable, "test.py", "-m", model, "-s", case]
            rc_e, dt_e = _run(test_cmd, env, DST, os.path.join(out, f"test_{arm}.log"))
            rec.update(test_rc=rc_e, test_seconds=dt_e, test_cmd=" ".join(test_cmd[1:]))
            ev = os.path.join(model, "eval", f"iter_{it:06d}", "eval3d.yml")
            if os.path.exists(ev):
                y = _yaml(ev)
                rec["train_eval"] = {"psnr_3d": float(y["psnr_3d"]), "ssim_3d": float(y["ssim_3d"])}
            tv = os.path.join(model, "test", f"iter_{it}", "eval3d.yml")
            if os.path.exists(tv):
                y = _yaml(tv)
                rec["test_eval"] = {"psnr_3d": float(y["psnr_3d"]), "ssim_3d": float(y["ssim_3d"])}
            t2 = os.path.join(model, "test", f"iter_{it}", "eval2d_render_test.yml")
            if os.path.exists(t2):
                y = _yaml(t2)
                rec["test_eval"]["psnr_2d_test_views"] = float(y["psnr_2d"])
            pc = os.path.join(model, "point_cloud", f"iteration_{it}", "point_cloud.pickle")
            rec["point_cloud_written"] = os.path.exists(pc)
        summary["arms"][arm] = rec
    a = summary["arms"]
    if all("test_eval" in a.get(k, {}) for k in ("ours", "refkernels")):
        summary["psnr_3d_delta_db"] = a["ours"]["test_eval"]["psnr_3d"] - a["refkernels"]["test_eval"]["psnr_3d"]
        summary["train_speedup_wall"] = a["refkernels"]["train_seconds"] / a["ours"]["train_seconds"]
    ok = all(v.get("train_rc") == 0 and v.get("test_rc") == 0 for v in a.values())
    if "psnr_3d_delta_db" in summary:
        ok = ok and abs(summary["psnr_3d_delta_db"]) <= 0.2
    summary["pass"] = bool(ok)
    with open(os.path.join(out, "summary.json"), "w") as f:
        json.dump(summary, f, indent=1)
    print(json.dumps(summary))
    return 0 if ok else 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--prepare", action="store_true")
    ap.add_argument("--run", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "refdrv"))
    ap.add_argument("--arms", default="ours,refkernels")
    ap.add_argument("--iterations", type=int, default=1500)
    ap.add_argument("--densify_from", type=int, default=300)
    ap.add_argument("--densify_until", type=int, default=1200)
    ap.add_argument("--det", type=int, default=128)
    ap.add_argument("--vox", type=int, default=64)
    ap.add_argument("--train_views", type=int, default=25)
    ap.add_argument("--test_views", type=int, default=5)
    ap.add_argument("--init_points", type=int, default=5000)
    args = ap.parse_args()
    if args.prepare:
        prepare()
    if args.run:
        return run(args)
    return 0


if __name__ == "__main__":
    sys.exit(main())
