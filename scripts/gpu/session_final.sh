#!/bin/bash
# Final 1-GPU session of the round: the driver's own sequence (GPU tests, smoke, both bench arms), the ncu launch list of
# the bench command, one ncu --set full pass over every kernel, the training-iteration numbers and configs[3]'s one-GPU leg.
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02_final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02_final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r02_final_smoke.log
timeout 600 python bench.py --impl reference > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err; echo "bench(reference) rc=$?"
R2X_BENCH_TRACE=1 timeout 600 python bench.py > $O/r02_bench_ours.json 2> $O/r02_bench_ours.err; echo "bench(ours) rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_forward.csv python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-e2e --no-secondary --no-parity > /dev/null 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/r02_all_kernels python scripts/profile_all_kernels.py > $O/r02_ncu_all.log 2>&1; echo "ncu all rc=$?"
python scripts/profile_all_kernels.py --summarize /tmp/r02_all_kernels.ncu-rep $O/r02_ncu_all_kernels_final.csv >> $O/r02_ncu_all.log 2>&1
ncu -i /tmp/r02_all_kernels.ncu-rep --page details --csv 2>/dev/null | grep -E "^\"ID\"|r2x::" | gzip > $O/r02_ncu_all_details_final.csv.gz
timeout 200 python scripts/gpu/train_profile.py > $O/r02_final_train_profile.json 2> $O/r02_final_train_profile.err; echo "train profile rc=$?"
timeout 300 python scripts/run_config4.py --gpus 1 --out $O/config3_1gpu > $O/r02_final_config3_1gpu.log 2>&1; echo "config3 1 GPU rc=$?"
python - <<'PY'
import json
O="gpurun_out/"
for f in ("r02_bench_reference.json","r02_bench_ours.json"):
    try:
        d=json.loads(open(O+f).read().strip().splitlines()[-1])
        print(f, round(d["value"],1), d["unit"], round(d["ms_per_step"]*1e3,1), "us; e2e", d.get("e2e",{}).get("value"), "parity", d.get("parity",{}).get("max_rel_to_max"), "roofline frac", d.get("roofline",{}).get("frac"), "clocks", d.get("clocks"))
        for k,v in d.get("secondary",{}).items():
            if isinstance(v,dict): print("  ", k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a not in ("workload","roofline","parity","api","autograd_path_api")})
        print("   cpu", d.get("cpu_baseline"))
    except Exception as e:
        print(f, "ERR", e, open(O+f.replace(".json",".err")).read()[-600:])
try:
    p=json.loads(open(O+"r02_final_train_profile.json").read().strip().splitlines()[-1])
    print({k:(round(v,4) if isinstance(v,float) else v) for k,v in p.items() if not k.startswith("top") and not k.startswith("cprofile")})
except Exception as e: print("profile ERR", e)
PY
tail -c 700 $O/r02_final_config3_1gpu.log; echo
cut -d, -f1,2,5,6,7,12,16 $O/r02_ncu_all_kernels_final.csv | head -80
grep -c "r2x::" $O/r02_launches_forward.csv
