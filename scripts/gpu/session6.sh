#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
python scripts/gpu/grad_diag.py > $O/r02_graddiag2.json 2> $O/r02_graddiag2.err
python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider > $O/r02_tests6.log 2>&1
Q="--no-cpu-baseline --no-secondary --no-e2e --steps 100 --warmup 10"
for v in 3 1; do R2X_RENDER_VARIANT=$v python bench.py $Q > $O/r02_b6_v$v.json 2> $O/r02_b6_v$v.err; done
ncu --set full --clock-control none --import-source on -k regex:raster_render_ws_kernel -c 1 -s 3 -f -o $O/r02_render_ws \
    python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-secondary --no-e2e --no-parity > $O/r02_ncu_ws.log 2>&1
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_graddiag2.json"))
print({k:(v if not isinstance(v,dict) else {a:b for a,b in v.items() if a!="worst_ours_vs_oracle"}) for k,v in d.items()})
PY
grep -E "passed|failed" $O/r02_tests6.log | tail -3; grep -E "^/|Error|error" $O/r02_tests6.log | head -30
for f in $O/r02_b6_v*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"]), "proj/s", round(d["ms_per_step"]*1e3,1), "us/step; render", round(d["roofline"]["kernel_ms"]*1e3,1), "us; parity", d.get("parity",{}).get("max_rel_to_max"), d.get("parity",{}).get("pass"))
except Exception as e:
    print(sys.argv[1], "ERR", e, open(sys.argv[1].replace(".json",".err")).read()[-600:])
PY
done
