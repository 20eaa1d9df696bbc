#!/bin/bash
# GPU session 2 (round 2): full GPU test suite, render-kernel variants, bench with parity + secondary, ncu launch list
# and one ncu --set full capture of the render kernel.  Outputs under gpurun_out/.
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -45 > $O/r02_tests2.log
Q="--no-cpu-baseline --no-secondary --no-e2e --steps 100 --warmup 10"
for v in 1 0; do R2X_RENDER_VARIANT=$v python bench.py $Q > $O/r02_var_v$v.json 2> $O/r02_var_v$v.err; done
for c in 64 128 192; do R2X_CHUNK=$c python bench.py $Q > $O/r02_var_c$c.json 2> $O/r02_var_c$c.err; done
python bench.py --no-cpu-baseline > $O/r02_bench1.json 2> $O/r02_bench1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02_launches1.csv \
    python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-secondary --no-e2e --no-parity > $O/r02_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:raster_render_kernel -c 1 -s 3 -f -o $O/r02_render_v1 \
    python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-secondary --no-e2e --no-parity > $O/r02_ncu_f.log 2>&1
tail -25 $O/r02_tests2.log
for f in $O/r02_var_*.json $O/r02_bench1.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], round(d["value"]), "proj/s", round(d["ms_per_step"]*1e3,1), "us/step; render", round(d["roofline"]["kernel_ms"]*1e3,1), "us; parity", d.get("parity",{}).get("max_rel_to_max"), d.get("parity",{}).get("pass"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
tail -3 $O/r02_bench1.err
