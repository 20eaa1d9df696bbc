#!/usr/bin/env python
"""BASELINE config 4: a 300k-Gaussian cloud trained Gaussian-sharded across N GPUs (one process per GPU, per-shard Adam
and densification, one exchange of the detector image and of the TV crop per iteration), compared with the same run on
one GPU: 3-D PSNR, 2-D PSNR on held-out views, milliseconds per iteration, final number of Gaussians.

    python scripts/run_config4.py --gpus 8 [--iterations 600] [--init 300000] [--out gpurun_out/config4]

Writes <out>/summary.json; exit code 0 iff both runs complete and their 3-D PSNR agrees to 0.3 dB."""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(cmd, log, env=None):
    t0 = time.time()
    with open(log, "w") as f:
        r = subprocess.run(cmd, stdout=f, stderr=subprocess.STDOUT, cwd=ROOT, env=env)
    return r.returncode, time.time() - t0


def last_json(path):
    for line in reversed(open(path).read().strip().splitlines()):
        line = line.strip()
        if line.startswith("{") and line.endswith("}"):
            try:
                return json.loads(line)
            except Exception:
                continue
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=8)
    ap.add_argument("--iterations", type=int, default=600)
    ap.add_argument("--init", type=int, default=300000)
    ap.add_argument("--det", type=int, default=256)
    ap.add_argument("--vox", type=int, default=160)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "config4"))
    ap.add_argument("--peer_exchange", action="store_true")
    ap.add_argument("--skip_one_gpu", action="store_true", help="only the sharded run (the 1-GPU leg was measured apart)")
    ap.add_argument("--work", default="/tmp/r2x_config4", help="case data and model directories (large; not returned)")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    os.makedirs(a.work, exist_ok=True)
    case = os.path.join(a.work, "case_phantom")
    if not os.path.exists(os.path.join(case, "meta_data.json")):
        subprocess.check_call([sys.executable, os.path.join(ROOT, "scripts", "make_synthetic_case.py"), case, "--det",
                               str(a.det), "--vox", str(a.vox), "--train", "50", "--test", "10", "--init", str(a.init)])
    it = a.iterations
    common = ["-s", case, "--iterations", str(it), "--densify_from_iter", str(it // 3), "--densify_until_iter",
              str(2 * it // 3), "--test_iterations", str(it), "--max_num_gaussians", "500000"]
    env = dict(os.environ, PYTHONPATH=ROOT)
    out = {"config": {"init_points": a.init, "detector": a.det, "volume": a.vox, "iterations": it, "gpus": a.gpus}}
    rc1 = 0
    if not a.skip_one_gpu:
        rc1, dt1 = run([sys.executable, "-m", "r2_gaussian_b200.trainer", *common, "-m", os.path.join(a.work, "model_1gpu")],
                       os.path.join(a.out, "train_1gpu.log"), env)
        out["one_gpu"] = {"rc": rc1, "wall_seconds": dt1, "result": last_json(os.path.join(a.out, "train_1gpu.log"))}
    if a.gpus <= 1:
        with open(os.path.join(a.out, "summary.json"), "w") as f:
            json.dump(out, f, indent=1)
        print(json.dumps(out))
        return rc1
    cmdN = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}", "--master-addr",
            "127.0.0.1", "--master-port", "29631", "-m", "r2_gaussian_b200.trainer", *common, "-m",
            os.path.join(a.work, f"model_{a.gpus}gpu")]
    if a.peer_exchange:
        cmdN.append("--peer_exchange")
    rcN, dtN = run(cmdN, os.path.join(a.out, f"train_{a.gpus}gpu.log"), env)
    out["sharded"] = {"rc": rcN, "wall_seconds": dtN, "result": last_json(os.path.join(a.out, f"train_{a.gpus}gpu.log")),
                      "exchange": "peer-memory kernel" if a.peer_exchange else "NCCL all-reduce"}
    ok = rc1 == 0 and rcN == 0 and out["sharded"]["result"]
    if ok and "one_gpu" in out and out["one_gpu"]["result"]:
        d = out["sharded"]["result"]["psnr_3d"] - out["one_gpu"]["result"]["psnr_3d"]
        out["psnr_3d_delta_db"] = d
        out["iteration_speedup"] = out["one_gpu"]["result"]["ms_per_iteration"] / out["sharded"]["result"]["ms_per_iteration"]
        ok = abs(d) <= 0.3
    out["pass"] = bool(ok)
    with open(os.path.join(a.out, "summary.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
