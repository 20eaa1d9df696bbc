#!/bin/bash
# Multi-GPU session: N = number of GPUs of this box (gpurun --gpus N).  Correctness of the Gaussian-sharded path, the
# scaling bench line for both exchange implementations, and (N = 8) BASELINE config 4.
N=${R2X_GPUS:-$(nvidia-smi -L | wc -l)}
mkdir -p gpurun_out; O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29611 scripts/check_sharded_gpu.py --p2p > $O/r02_mg${N}_check_p2p.log 2>&1
$TR --master-port 29612 scripts/check_sharded_gpu.py > $O/r02_mg${N}_check_nccl.log 2>&1
$TR --master-port 29613 bench.py --gpus $N --steps 200 --warmup 20 --reduce p2p > $O/r02_mg${N}_bench_p2p.json 2> $O/r02_mg${N}_bench_p2p.err
$TR --master-port 29614 bench.py --gpus $N --steps 200 --warmup 20 --reduce nccl > $O/r02_mg${N}_bench_nccl.json 2> $O/r02_mg${N}_bench_nccl.err
if [ "$N" = "8" ]; then
  python scripts/run_config4.py --gpus 8 --out $O/config4 > $O/r02_config4.log 2>&1
  python scripts/run_config4.py --gpus 8 --out $O/config4_p2p --peer_exchange > $O/r02_config4_p2p.log 2>&1
fi
grep -h "OK\|FAIL" $O/r02_mg${N}_check_p2p.log $O/r02_mg${N}_check_nccl.log | head -20
for f in $O/r02_mg${N}_bench_p2p.json $O/r02_mg${N}_bench_nccl.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "N", d["n_gpus"], round(d["value"]), "proj/s", round(d["ms_per_step"]*1e3,1), "us/step; warm", round(d["value_warm_l2_back_to_back"]), "e2e", round(d.get("e2e",{}).get("value",0)), "parity", d.get("parity",{}).get("max_rel_to_max"))
except Exception as e:
    print(sys.argv[1], "ERR", e, open(sys.argv[1].replace(".json",".err")).read()[-800:])
PY
done
if [ "$N" = "8" ]; then tail -c 1500 $O/r02_config4.log; tail -c 800 $O/r02_config4_p2p.log; fi
