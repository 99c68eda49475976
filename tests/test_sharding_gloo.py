"""z-slab sharding (N>1 path) on CPU: world_size 2 and 3 over gloo, oracle engines.
The sharded run must reproduce the single-domain run bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from alis_b200 import capi, synthetic


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, shape, steps, outdir, overlap, fuse=False):
    import torch.distributed as dist
    from alis_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = capi.load_oracle()
    w = synthetic.BoxWorld(*shape)
    e = w.make_engine(lib, rank=rank, world=world)
    r = sharding.SlabRunner(e, rank, world, cuda=False, overlap=overlap, fuse=fuse)
    r.run(steps)
    c0, c1 = e.shard
    a = e.read_all_fields()                       # local (nz_local+2) planes
    np.save(os.path.join(outdir, f"r{rank}.npy"), a[..., 1:-1])
    np.save(os.path.join(outdir, f"g{rank}.npy"), a)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,shape,overlap,fuse", [(2, (40, 34, 44), True, False), (3, (36, 40, 61), True, False), (2, (40, 34, 44), False, False),
                                                     (3, (36, 40, 61), True, True), (2, (40, 34, 44), False, True)])
def test_sharded_equals_single(world, shape, overlap, fuse, tmp_path, oracle_lib):
    """fuse=True drives the runner's fused-stepping control flow (begin_pass / deferred E / flush); the
    oracle engine steps in two plain passes either way, so the result must not change"""
    steps = 40
    w = synthetic.BoxWorld(*shape)
    single = w.make_engine(oracle_lib)
    single.run(steps)
    ref = single.read_all_fields()
    assert np.abs(ref).max() > 0
    mp.spawn(_worker, args=(world, _free_port(), shape, steps, str(tmp_path), overlap, fuse), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"r{r}.npy") for r in range(world)]
    got = np.concatenate(parts, axis=-1)
    assert np.array_equal(got, ref[..., 1:-1]), "sharded interior planes differ from the single-domain run"
    # the ghost planes at the interfaces hold the neighbour's tangential components (x,y)
    for r in range(world - 1):
        lo, hi = np.load(tmp_path / f"g{r}.npy"), np.load(tmp_path / f"g{r + 1}.npy")
        assert np.array_equal(lo[0, 1, :2, :, :, -1], hi[0, 1, :2, :, :, 1])      # H ghost of lower == first H plane of upper
        assert np.array_equal(hi[0, 0, :2, :, :, 0], lo[0, 0, :2, :, :, -2])      # E ghost of upper == top E plane of lower


def test_split_pass_equals_whole_pass(oracle_lib):
    """SEND_PLANE followed by REST is the same pass.  The split is defined for the side of a shard
    that faces a neighbour (rank 0 splits its E pass, the last rank its H pass); oracle engine,
    the CUDA twin is in test_gpu_parity."""
    from tests import helpers
    w = synthetic.BoxWorld(30, 28, 34, num=4)
    for rank, split_set in ((0, capi.SET_E), (1, capi.SET_H)):
        a, b = w.make_engine(oracle_lib, rank=rank, world=2), w.make_engine(oracle_lib, rank=rank, world=2)
        helpers.seed_fields([a, b], (32, 30, a.nz + 2), seed=5)
        for _ in range(12):
            a.step_E()
            a.step_H()
            for s in (capi.SET_E, capi.SET_H):
                if s == split_set:
                    b.step_part(s, capi.PART_SEND_PLANE)
                    b.step_part(s, capi.PART_REST)
                else:
                    b.step_part(s, capi.PART_ALL)
        assert a.time() == b.time()
        helpers.assert_fields_identical(a, b, f"split pass rank {rank}")
