#!/bin/bash
# 1-GPU session: full suite (native training step, packed backward kernel, two-level voxel binning), backward variants,
# bench with the CPU baselines in guarded children, training-iteration numbers.
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02_s12_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $O/r02_s12_pytest.log
R2X_BWD_VARIANT=1 timeout 300 python -m pytest tests/test_raster_gpu.py tests/test_parity_baseline_gpu.py -q -k "backward or grad" > $O/r02_s12_pytest_bwd1.log 2>&1; echo "pytest(bwd variant 1) rc=$?"; tail -3 $O/r02_s12_pytest_bwd1.log
timeout 300 python scripts/gpu/train_profile.py > $O/r02_s12_train_profile.json 2> $O/r02_s12_train_profile.err; echo "profile rc=$?"
R2X_BENCH_TRACE=1 timeout 600 python bench.py > $O/r02_s12_bench.json 2> $O/r02_s12_bench.err; echo "bench rc=$?"
R2X_BWD_VARIANT=1 timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-parity --steps 20 --warmup 5 > $O/r02_s12_bench_bwd1.json 2> /dev/null; echo "bench(bwd1) rc=$?"
grep "bench +" $O/r02_s12_bench.err
python - <<'PY'
import json
O="gpurun_out/"
def sec(f):
    try:
        d=json.loads(open(O+f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), "proj/s", round(d["ms_per_step"]*1e3,1), "us; e2e", d.get("e2e",{}).get("value"), "parity", d.get("parity",{}).get("max_rel_to_max"))
        print("   cpu_baseline", d.get("cpu_baseline"), d.get("cpu_baseline_torch"))
        for k,v in d.get("secondary",{}).items():
            if isinstance(v,dict): print("  ", k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a not in ("workload","roofline","parity","api","autograd_path_api")})
    except Exception as e:
        print(f, "ERR", e)
sec("r02_s12_bench.json"); sec("r02_s12_bench_bwd1.json")
try:
    p=json.loads(open(O+"r02_s12_train_profile.json").read().strip().splitlines()[-1])
    print({k:(round(v,4) if isinstance(v,float) else v) for k,v in p.items() if not k.startswith("top") and not k.startswith("cprofile")})
except Exception as e:
    print("profile ERR", e, open(O+"r02_s12_train_profile.err").read()[-1500:])
PY
