#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
python scripts/gpu/grad_diag.py > $O/r02_graddiag_fast.json 2> $O/r02_graddiag_fast.err
R2X_BWD_EXACT=1 python scripts/gpu/grad_diag.py > $O/r02_graddiag_exact.json 2> $O/r02_graddiag_exact.err
Q="--no-cpu-baseline --no-secondary --no-e2e --steps 100 --warmup 10"
for v in 3 2 1; do R2X_RENDER_VARIANT=$v python bench.py $Q > $O/r02_b5_v$v.json 2> $O/r02_b5_v$v.err; done
python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider > $O/r02_tests5.log 2>&1
cat $O/r02_graddiag_fast.json; cat $O/r02_graddiag_exact.json; tail -c 600 $O/r02_graddiag_fast.err
grep -E "passed|failed" $O/r02_tests5.log | tail -3; grep -E "^/|Error|error" $O/r02_tests5.log | head -40
for f in $O/r02_b5_v*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"]), "proj/s", round(d["ms_per_step"]*1e3,1), "us/step; render", round(d["roofline"]["kernel_ms"]*1e3,1), "us; parity", d.get("parity",{}).get("max_rel_to_max"), d.get("parity",{}).get("pass"))
except Exception as e:
    print(sys.argv[1], "ERR", e, open(sys.argv[1].replace(".json",".err")).read()[-600:])
PY
done
