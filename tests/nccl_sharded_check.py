"""Multi-GPU parity check of the column-sharded path over NCCL / peer memory (run under torchrun on a box with >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/nccl_sharded_check.py

The cases live in tests/sharded_parity.py (rank-1 and look-ahead loops, Bland and Dantzig, NCCL and peer-memory
exchange, ragged shards).  tests/test_gpu_multi.py launches this script from pytest when the box has >= 2 GPUs; the
pytest suite also covers the protocol on CPU with gloo and, on one GPU, with two emulated shards."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))

import torch
import torch.distributed as dist

from tests.sharded_parity import run_cases


def main():
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    n, ok = run_cases(rank, world, local, log=lambda s: print(s, flush=True))
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        sys.exit(1)
    if rank == 0:
        print(f"NCCL sharded parity OK ({n} cases, {world} GPUs)")


if __name__ == "__main__":
    main()
