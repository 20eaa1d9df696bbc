#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider > $O/r02_tests9.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke9.log 2>&1
ncu --set full --clock-control none --profile-from-start off -f -o /tmp/r02_all_kernels python scripts/profile_all_kernels.py > $O/r02_ncu_all.log 2>&1
python scripts/profile_all_kernels.py --summarize /tmp/r02_all_kernels.ncu-rep $O/r02_ncu_all_kernels.csv >> $O/r02_ncu_all.log 2>&1
ncu -i /tmp/r02_all_kernels.ncu-rep --page details --csv 2>/dev/null | grep -E "^\"ID\"|r2x::" | gzip > $O/r02_ncu_all_details.csv.gz
ls -la $O/r02_ncu_all_details.csv.gz /tmp/r02_all_kernels.ncu-rep
grep -E "passed|failed" $O/r02_tests9.log | tail -3; grep -E "^/|Error|error|FAILED" $O/r02_tests9.log | head -20; tail -2 $O/r02_smoke9.log
cat $O/r02_ncu_all_kernels.csv | cut -d, -f1,2,3,4,5,6,7,13,16,17 | head -70
