#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
for a in 2.1 4.4; do DIAG_KIND=trained DIAG_ANGLE=$a python scripts/gpu/grad_diag.py > $O/r02_graddiag_t$a.json 2> $O/r02_graddiag_t$a.err; done
Q="--no-cpu-baseline --no-secondary --no-e2e --steps 100 --warmup 10 --no-parity"
python bench.py $Q > $O/r02_b8_pdl.json 2> $O/r02_b8_pdl.err
R2X_NO_PDL=1 python bench.py $Q > $O/r02_b8_nopdl.json 2> $O/r02_b8_nopdl.err
ncu --set full --clock-control none --import-source on --profile-from-start off -f -o $O/r02_all_kernels python scripts/profile_all_kernels.py > $O/r02_ncu_all.log 2>&1
python scripts/profile_all_kernels.py --summarize $O/r02_all_kernels.ncu-rep $O/r02_ncu_all_kernels.csv
python - <<'PY'
import json
for a in ("2.1","4.4"):
    try:
        d=json.load(open(f"gpurun_out/r02_graddiag_t{a}.json"))
        print(a, {k:{"ovr":v["ours_vs_ref"],"ovo":v["ours_vs_oracle"],"rvo":v["ref_vs_oracle"],"worst":v.get("worst_ours_vs_ref")} for k,v in d.items() if isinstance(v,dict)})
    except Exception as e: print(a,"ERR",e)
PY
for f in $O/r02_b8_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"]), "proj/s", round(d["ms_per_step"]*1e3,1), "us/step; render", round(d["roofline"]["kernel_ms"]*1e3,1), "us; warm", round(d["value_warm_l2_back_to_back"]))
except Exception as e:
    print(sys.argv[1], "ERR", e, open(sys.argv[1].replace(".json",".err")).read()[-600:])
PY
done
cut -d, -f1,2,5,6,7,16 $O/r02_ncu_all_kernels.csv | head -60
