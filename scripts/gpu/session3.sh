#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
python scripts/gpu/repro_paths.py > $O/r02_repro.log 2>&1; echo "repro rc=$?" >> $O/r02_repro.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 8 python scripts/gpu/repro_paths.py > $O/r02_sanitize.log 2>&1; echo "sanitizer rc=$?" >> $O/r02_sanitize.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -120 > $O/r02_tests3.log
tail -30 $O/r02_repro.log; grep -E "Invalid|Error|ERROR SUMMARY|at .*kernel|OK" $O/r02_sanitize.log | head -40; tail -60 $O/r02_tests3.log
