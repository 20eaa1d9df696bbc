#!/bin/bash
# N-GPU session: configs[3] (300k Gaussians sharded, native step, both exchanges; set-up outside the clock) and the
# voxelizer sweep of configs[4] at N ranks.
N=${R2X_GPUS:-$(nvidia-smi -L | wc -l)}
mkdir -p gpurun_out; O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 python scripts/run_config4.py --gpus $N --skip_one_gpu --out $O/config3_n${N}_nccl > $O/r02_mgc${N}_config3_nccl.log 2>&1
timeout 200 python scripts/run_config4.py --gpus $N --skip_one_gpu --out $O/config3_n${N}_p2p --peer_exchange > $O/r02_mgc${N}_config3_p2p.log 2>&1
timeout 200 $TR --master-port 29641 scripts/bench_voxel_sharded.py > $O/r02_mgc${N}_voxel_sweep.json 2> $O/r02_mgc${N}_voxel_sweep.err
tail -c 1100 $O/r02_mgc${N}_config3_nccl.log; echo; tail -c 1100 $O/r02_mgc${N}_config3_p2p.log; echo
tail -c 1500 $O/r02_mgc${N}_voxel_sweep.json; tail -3 $O/r02_mgc${N}_voxel_sweep.err
