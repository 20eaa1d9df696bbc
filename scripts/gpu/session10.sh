#!/bin/bash
# 1-GPU session: full GPU suite on the merged tree, bench line (secondary incl. the trained-like voxel sweep),
# training-iteration breakdown, config 4 one-GPU leg.
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02_s10_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r02_s10_pytest.log
timeout 600 python bench.py > $O/r02_s10_bench.json 2> $O/r02_s10_bench.err; echo "bench rc=$?"
timeout 300 python scripts/gpu/train_profile.py > $O/r02_s10_train_profile.json 2> $O/r02_s10_train_profile.err; echo "profile rc=$?"
R2X_FUSED_ACTIVATIONS=0 timeout 300 python scripts/gpu/train_profile.py > $O/r02_s10_train_profile_unfused.json 2>> $O/r02_s10_train_profile.err
timeout 600 python scripts/run_config4.py --gpus 1 --out $O/config4_1gpu > $O/r02_s10_config4_1gpu.log 2>&1; echo "config4 rc=$?"
python - <<'PY'
import json
O="gpurun_out/"
try:
    d=json.loads(open(O+"r02_s10_bench.json").read().strip().splitlines()[-1])
    print("bench", round(d["value"]), "proj/s", round(d["ms_per_step"]*1e3,1), "us; e2e", d["e2e"]["value"], "parity", d["parity"]["max_rel_to_max"])
    s=d.get("secondary",{})
    for k,v in s.items():
        print(" ", k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a not in ("workload","roofline","parity")}, v.get("parity"))
except Exception as e:
    print("bench ERR", e, open(O+"r02_s10_bench.err").read()[-1500:])
for f in ("r02_s10_train_profile.json","r02_s10_train_profile_unfused.json"):
    try:
        p=json.loads(open(O+f).read().strip().splitlines()[-1])
        print(f, "wall", round(p["wall_ms_per_iteration"],3), "enqueue", round(p["host_enqueue_ms_per_iteration"],3), "gpu", round(p["gpu_kernel_ms_per_iteration"],3), "launches", p["gpu_launches_per_iteration"])
        for e in p["top_gpu"][:14]: print("   G", e["name"][:60].ljust(60), round(e["us_per_iteration"],1), e["calls_per_iteration"])
        for e in p["top_host"][:12]: print("   H", e["name"][:60].ljust(60), round(e["us_per_iteration"],1), e["calls_per_iteration"])
    except Exception as e:
        print(f, "ERR", e, open(O+"r02_s10_train_profile.err").read()[-1500:])
PY
tail -c 1200 $O/r02_s10_config4_1gpu.log
