"""Launched by tests/test_gpu_multi.py under torch.distributed.run: every rank steps its z-slab
on its own GPU with NCCL halo exchange; rank 0 also runs the undivided world and compares."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alis_b200 import capi, sharding, synthetic  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_stream(torch.cuda.Stream())
    shape, steps = (150, 140, 40 * world + 3), 60
    lib = capi.load_engine()
    w = synthetic.BoxWorld(*shape)
    e = w.make_engine(lib, device=local, rank=rank, world=world)
    r = sharding.SlabRunner(e, rank, world, cuda=True, fuse=bool(os.environ.get("ALISB_TEST_FUSE")))
    r.run(steps)
    torch.cuda.synchronize()
    mine = e.read_all_fields()[..., 1:-1]
    c0, c1 = e.shard
    ok = torch.ones(1, device="cuda")
    single = w.make_engine(lib, device=local)          # every rank checks its own planes
    single.run(steps)
    ref = single.read_all_fields()[..., c0:c1 + 1]
    if not np.array_equal(mine, ref) or np.abs(ref).max() == 0:
        ok.zero_()
        print(f"rank {rank}: MISMATCH planes {c0}..{c1}: {int((mine != ref).sum())} values", flush=True)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SHARDED_NCCL_OK" if ok.item() == 1 else "SHARDED_NCCL_FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok.item() == 1 else 1)


if __name__ == "__main__":
    main()
