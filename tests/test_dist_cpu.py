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


class _FakeShark:
    """Records the sharded-build protocol calls; checks the cross-rank ordering through files:
    a step may only start when every rank has finished the previous one."""

    def __init__(self, rank, world, tmp):
        self.rank, self.world, self.tmp = rank, world, tmp
        self.log = []

    def _done(self, step):
        open(os.path.join(self.tmp, "%s.%d" % (step, self.rank)), "w").write("x")

    def _need(self, step):
        for r in range(self.world):
            assert os.path.exists(os.path.join(self.tmp, "%s.%d" % (step, r))), (step, r, self.rank)

    def shard_begin(self, bases, rec_off, shard, n_shards):
        from shark_b200 import capi
        assert (shard, n_shards) == (self.rank, self.world)
        m = capi.ShardMem(pid=1000 + self.rank, device=self.rank, shard=shard, ipc_ok=1)
        m.dev_ptr[0] = 0x1000 * (self.rank + 1)
        self.log.append("begin")
        self._done("begin")
        return m

    def shard_open(self, peer):
        from shark_b200 import capi
        assert peer.shard != self.rank and peer.pid == 1000 + peer.shard
        o = capi.ShardMem.from_buffer_copy(bytes(peer))
        o.pid = -1
        self.log.append("open%d" % peer.shard)
        return o

    def shard_merge(self, phase, arr):
        self._need("begin" if phase == 1 else "merge1")
        assert [arr[s].shard for s in range(self.world)] == list(range(self.world))
        assert arr[self.rank].pid == 1000 + self.rank and all(arr[s].pid == -1 for s in range(self.world) if s != self.rank)
        assert all(arr[s].dev_ptr[0] == 0x1000 * (s + 1) for s in range(self.world))
        self.log.append("merge%d" % phase)
        self._done("merge%d" % phase)

    def shard_rank(self):
        self.log.append("rank")
        self._done("rank")

    def shard_finish(self, arr):
        from shark_b200 import capi
        self._need("rank")
        self.log.append("finish")
        self._done("finish")
        return capi.IndexInfo(n_shards=self.world)

    def shard_close(self, opened):
        self._need("finish")
        self.log.append("close%d" % opened.shard)
        self._done("close")

    def shard_end(self):
        self._need("close")
        self.log.append("end")


def _shard_worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from shark_b200 import dist_index
    sh = _FakeShark(rank, world, tmp)
    info, secs = dist_index.build_index_sharded(sh, np.zeros(4, np.uint8), np.array([0, 4], np.uint64))
    assert info.n_shards == world and secs >= 0
    other = 1 - rank
    assert sh.log == ["begin", "open%d" % other, "merge1", "merge2", "rank", "finish", "close%d" % other, "end"], sh.log
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(tmp, "ok%d" % rank), "w").write("ok")


def test_sharded_build_protocol_gloo_world2(tmp_path):
    """Host side of the one-process-per-GPU sharded index build: the handle structs travel as bytes,
    arrive in shard order, and the barriers separate the steps on every rank."""
    world = 2
    mp.spawn(_shard_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))


def test_shard_cuts_properties():
    """shk_shard_cuts (pure host): monotone cuts at record boundaries covering [0, total), balanced."""
    from shark_b200 import dist_index
    rng = np.random.default_rng(11)
    for n_rec in (0, 1, 7, 1000):
        lens = rng.integers(0, 5000, n_rec)
        rec_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        total = int(rec_off[-1])
        for n in (1, 2, 3, 8, 16):
            cuts = dist_index.shard_cuts(rec_off, n)
            assert len(cuts) == n + 1 and cuts[0] == 0 and cuts[-1] == total
            assert np.all(np.diff(cuts.astype(np.int64)) >= 0)
            assert np.all(np.isin(cuts, rec_off))
            if n_rec == 1000:
                ideal = total / n
                assert np.max(np.abs(np.diff(cuts.astype(np.int64)) - ideal)) <= 5000


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


def test_reference_arm_json_contract():
    """`bench.py --impl reference` (the reference's own CPU path, oracle/_ref/shark) on a small sample: one JSON
    line with the keys the driver reads; no GPU involved."""
    import json
    import subprocess
    import sys
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "shark")):
        pytest.skip("oracle/_ref/shark is not built here")
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--cpu-sample", "100000", "--workloads", "c2"], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    lines = [ln for ln in p.stdout.decode().splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "reads/sec" and d["unit"] == "reads/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] == d["cpu_baseline"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["config"]["workload"].startswith("C2") and d["gpu_launches"] == 0
