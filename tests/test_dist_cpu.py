"""Multi-process host logic on CPU (gloo, world_size 2): index-info serialisation, piecewise buffer
broadcast, read sharding and ordered merge - the plumbing bench.py and the multi-GPU path use with
NCCL on the box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT  # noqa: F401  (sys.path)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from shark_b200 import capi, dist_index
    # 1. index info travels as bytes
    info = capi.IndexInfo(n_records=7, n_genes=6, n_set_bits=123456789012, tot_ids=99, n_windows=5, bf_bits=1 << 33,
                          device_bytes=42, build_ms=1.5, front_shift=11, front_entries=1 << 22)
    t = dist_index.pack_info(info if rank == 0 else None)
    dist.broadcast(t, src=0)
    got = dist_index.unpack_info(t)
    assert (got.n_records, got.n_genes, got.n_set_bits, got.bf_bits, got.front_shift, got.front_entries) == \
        (7, 6, 123456789012, 1 << 33, 11, 1 << 22)
    # 2. piecewise broadcast of a large byte view (piece size shrunk to force several pieces)
    dist_index.PIECE = 1000
    rng = np.random.default_rng(5)
    src = torch.from_numpy(rng.integers(0, 256, 4567, dtype=np.uint8))
    buf = src.clone() if rank == 0 else torch.zeros(4567, dtype=torch.uint8)
    dist_index.broadcast_bytes(buf, 0)
    assert torch.equal(buf, src)
    # 3. sharding: weak (per-rank block ranges) and strong (chunk j -> rank j mod world)
    first, n = dist_index.shard_blocks(rank, world, 10_000_000, 1 << 20)
    assert n == 10_000_000 and first == rank * 10 * (1 << 20)
    mine = dist_index.shard_chunks(7, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    assert sorted(sum(gathered, [])) == list(range(7))
    # 4. ordered merge of per-chunk results equals the single-rank order
    per_chunk = {j: [(j * 100 + i, j) for i in range(3)] for j in mine}
    allres = [None] * world
    dist.all_gather_object(allres, per_chunk)
    merged = {}
    for d in allres:
        merged.update(d)
    flat = [x for j in sorted(merged) for x in merged[j]]
    assert flat == [(j * 100 + i, j) for j in range(7) for i in range(3)]
    # 5. max-over-ranks timing reduction as bench.py does it
    tt = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    assert tt.item() == world
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(tmp, "ok%d" % rank), "w").write("ok")


def test_gloo_world2(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))


def test_reference_arm_exits_quietly_on_nonzero_rank(tmp_path):
    """`bench.py --impl reference` under torchrun: only rank 0 works, the others exit 0 silently."""
    import subprocess
    import sys
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 0 and p.stdout == b""
