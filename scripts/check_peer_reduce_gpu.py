"""torchrun --nproc-per-node N scripts/check_peer_reduce_gpu.py : the NVLink peer-memory sum equals the rank-ordered
sum bit for bit over many epochs (double buffering, flag protocol), and is timed against NCCL's all_reduce."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from r2_gaussian_b200.peer import PeerReducer

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = 512 * 512
red = PeerReducer(n, dev)
out = torch.empty(n, device=dev)
ok = True
for ep in range(40):
    g = torch.Generator(dev).manual_seed(1000 * ep + rank)
    mine = torch.randn(n, device=dev, generator=g) * (1 + rank)
    red.partial().copy_(mine)
    red.reduce(out)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    want = parts[0].clone()
    for p in range(1, world):
        want += parts[p]
    ok = ok and torch.equal(out, want)
ok = ok and red.ok()


def timeit(fn, k=200):
    for _ in range(10):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / k], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


buf = torch.randn(n, device=dev)
t_nccl = timeit(lambda: dist.all_reduce(buf))
t_peer = timeit(lambda: red.reduce(out))
ok = ok and red.ok()
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"world": world, "bitwise_equal_to_rank_ordered_sum": bool(flag.item()), "n_floats": n,
                      "nccl_all_reduce_us": t_nccl * 1e3, "peer_allreduce_us": t_peer * 1e3}), flush=True)
red.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
