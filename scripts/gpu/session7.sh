#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
DIAG_KIND=trained python scripts/gpu/grad_diag.py > $O/r02_graddiag3.json 2> $O/r02_graddiag3.err
python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider > $O/r02_tests7.log 2>&1
Q="--no-cpu-baseline --no-secondary --no-e2e --steps 100 --warmup 10"
R2X_RENDER_VARIANT=3 python bench.py $Q > $O/r02_b7_v3.json 2> $O/r02_b7_v3.err
for c in 128 160 192 224; do R2X_CHUNK=$c python bench.py $Q > $O/r02_b7_c$c.json 2> $O/r02_b7_c$c.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02_launches2.csv \
    python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-secondary --no-e2e --no-parity > $O/r02_ncu_l2.log 2>&1
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_graddiag3.json"))
print({k:(v if not isinstance(v,dict) else v) for k,v in d.items()})
PY
grep -E "passed|failed" $O/r02_tests7.log | tail -3; grep -E "^/|Error|error|FAILED" $O/r02_tests7.log | head -30
for f in $O/r02_b7_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"]), "proj/s", round(d["ms_per_step"]*1e3,1), "us/step; render", round(d["roofline"]["kernel_ms"]*1e3,1), "us; parity", d.get("parity",{}).get("max_rel_to_max"), d.get("parity",{}).get("pass"))
except Exception as e:
    print(sys.argv[1], "ERR", e, open(sys.argv[1].replace(".json",".err")).read()[-600:])
PY
done
