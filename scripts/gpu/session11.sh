#!/bin/bash
# 1-GPU session: two-level voxel binning (parity vs radix / oracle / reference), training-iteration host profile,
# bench with stage trace.
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests/test_voxel_gpu.py tests/test_parity_baseline_gpu.py tests/test_train_gpu.py -x -q > $O/r02_s11_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $O/r02_s11_pytest.log
TRAIN_CPROFILE=1 timeout 300 python scripts/gpu/train_profile.py > $O/r02_s11_train_profile.json 2> $O/r02_s11_train_profile.err; echo "profile rc=$?"
R2X_BENCH_TRACE=1 timeout 400 python bench.py --no-cpu-baseline > $O/r02_s11_bench.json 2> $O/r02_s11_bench.err; echo "bench rc=$?"
grep "bench +" $O/r02_s11_bench.err
python - <<'PY'
import json
O="gpurun_out/"
try:
    d=json.loads(open(O+"r02_s11_bench.json").read().strip().splitlines()[-1])
    print("bench", round(d["value"]), "proj/s", round(d["ms_per_step"]*1e3,1), "us; e2e", d["e2e"]["value"], "parity", d["parity"]["max_rel_to_max"])
    for k,v in d.get("secondary",{}).items():
        print(" ", k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a not in ("workload","roofline","parity")}, v.get("parity"))
except Exception as e:
    print("bench ERR", e, open(O+"r02_s11_bench.err").read()[-1500:])
try:
    p=json.loads(open(O+"r02_s11_train_profile.json").read().strip().splitlines()[-1])
    print("train: wall", round(p["wall_ms_per_iteration"],3), "enqueue", round(p["host_enqueue_ms_per_iteration"],3), "gpu", round(p["gpu_kernel_ms_per_iteration"],3), "launches", p["gpu_launches_per_iteration"])
    for e in p.get("cprofile_top_cumulative", [])[:32]: print("   ", e["fn"][:60].ljust(60), round(e["calls_per_iteration"],1), round(e["self_us_per_iteration"],1), round(e["cum_us_per_iteration"],1))
except Exception as e:
    print("profile ERR", e, open(O+"r02_s11_train_profile.err").read()[-1500:])
PY
