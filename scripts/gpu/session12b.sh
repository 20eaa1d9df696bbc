#!/bin/bash
# voxel work items of up to 4096 instances walked in segments: voxel / parity / training tests + the secondary block
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests/test_voxel_gpu.py tests/test_parity_baseline_gpu.py tests/test_train_gpu.py tests/test_ref_gpu.py -x -q > $O/r02_s12b_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02_s12b_pytest.log
R2X_BENCH_TRACE=1 timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 50 --warmup 5 > $O/r02_s12b_bench.json 2> $O/r02_s12b_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_s12b_bench.json").read().strip().splitlines()[-1])
print(round(d["value"]), "proj/s")
for k,v in d.get("secondary",{}).items():
    if isinstance(v,dict): print("  ", k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a not in ("workload","roofline","api","autograd_path_api")})
PY
