#!/bin/bash
# N-GPU session (second of the round): sharded correctness, scaling bench line for both exchanges, BASELINE config 4
# trained Gaussian-sharded with the native training step (the one-GPU leg was measured apart: profiles/r02_config4_one_gpu.json).
N=${R2X_GPUS:-$(nvidia-smi -L | wc -l)}
mkdir -p gpurun_out; O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29611 scripts/check_sharded_gpu.py --p2p > $O/r02_mgb${N}_check_p2p.log 2>&1
timeout 300 $TR --master-port 29613 bench.py --gpus $N --steps 200 --warmup 20 --reduce p2p > $O/r02_mgb${N}_bench_p2p.json 2> $O/r02_mgb${N}_bench_p2p.err
timeout 300 $TR --master-port 29614 bench.py --gpus $N --steps 200 --warmup 20 --reduce nccl > $O/r02_mgb${N}_bench_nccl.json 2> $O/r02_mgb${N}_bench_nccl.err
timeout 300 python scripts/run_config4.py --gpus $N --skip_one_gpu --out $O/config4_n${N}_nccl > $O/r02_mgb${N}_config4_nccl.log 2>&1
timeout 300 python scripts/run_config4.py --gpus $N --skip_one_gpu --out $O/config4_n${N}_p2p --peer_exchange > $O/r02_mgb${N}_config4_p2p.log 2>&1
grep -h "OK\|FAIL" $O/r02_mgb${N}_check_p2p.log | head -8
for f in $O/r02_mgb${N}_bench_p2p.json $O/r02_mgb${N}_bench_nccl.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "N", d["n_gpus"], round(d["value"]), "proj/s", round(d["ms_per_step"]*1e3,1), "us/step; e2e", round(d.get("e2e",{}).get("value",0)), "parity", d.get("parity",{}).get("max_rel_to_max"))
except Exception as e:
    print(sys.argv[1], "ERR", e, open(sys.argv[1].replace(".json",".err")).read()[-800:])
PY
done
tail -c 900 $O/r02_mgb${N}_config4_nccl.log; echo; tail -c 900 $O/r02_mgb${N}_config4_p2p.log; echo
tail -5 $O/config4_n${N}_nccl/train_${N}gpu.log | cut -c1-300
du -sh $O
