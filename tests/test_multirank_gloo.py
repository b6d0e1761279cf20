"""CPU, world_size 2, gloo: the multi-GPU host logic (sample / tile sharding + the single reduce of the f64
XYZA accumulators).  The per-rank render is done by the oracle here (no GPU in this container); on the GPU
box the same sharding drives the CUDA path (bench.py --gpus N)."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import parity_util as pu

sharding = importlib.import_module("simple-spectral_b200.sharding")


def test_shards_partition_the_job():
    for world in (1, 2, 3, 4, 8):
        for total in (8, 64, 65):
            r = [sharding.sample_shard(k, world, total) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1
        r = [sharding.tile_shard(k, world, 30) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == 30 and all(r[i][1] == r[i + 1][0] for i in range(world - 1))
    with pytest.raises(ValueError):
        sharding.sample_shard(2, 2, 8)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mode, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
    sys.path.insert(0, os.path.dirname(__file__))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    flat = pu.load_flat("cornell", "ours1931")
    W, H, SPP = 16, 12, 6
    opt = sharding.shard_options(lambda **kw: pu.options("ours1931", W, H, seed=3, **kw), rank, world, SPP, mode=mode, height=H, band_height=5)
    acc, _, _ = pu.oracle_render(flat, opt)
    t = torch.from_numpy(acc.reshape(-1))
    sharding.reduce_accumulators(t, dist, dst=0)
    if rank == 0:
        np.save(out_path, t.numpy().reshape(H, W, 4))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["samples", "tiles", "bands"])
def test_two_ranks_reduce_to_the_single_rank_frame(tmp_path, mode):
    out = str(tmp_path / "acc.npy")
    mp.spawn(_worker, args=(2, _free_port(), mode, out), nprocs=2, join=True)
    got = np.load(out)
    flat = pu.load_flat("cornell", "ours1931")
    full = pu.options("ours1931", 16, 12, 6, seed=3)
    want, _, _ = pu.oracle_render(flat, full)
    if mode in ("tiles", "bands"):
        assert pu.bits_equal(got, want)            # disjoint pixels: bit-identical
    else:
        assert np.allclose(got, want, rtol=1e-14, atol=0)  # same samples, f64 summation order differs
