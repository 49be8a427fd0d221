"""Multi-GPU plumbing (one process per GPU): replicate the index built on one rank to all others
with NCCL broadcasts over NVLink, and shard reads by rank.  torch.distributed is
only the transport; the buffers are the library's own device allocations (shk_index_views_get).
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import capi

PIECE = 1 << 30  # broadcast large views in 1 GiB pieces


class _DevView:
    """Device memory owned by libshark_b200 exposed through __cuda_array_interface__."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def view_tensor(ptr, nbytes, device):
    return torch.as_tensor(_DevView(ptr, nbytes), device=device)


def pack_info(info):
    """shk_index_info -> uint8 tensor (CPU)."""
    raw = bytes(info) if info is not None else bytes(C.sizeof(capi.IndexInfo))
    return torch.frombuffer(bytearray(raw), dtype=torch.uint8).clone()


def unpack_info(t):
    return capi.IndexInfo.from_buffer_copy(bytes(t.cpu().numpy().tobytes()))


def broadcast_bytes(t, src):
    """Broadcast a 1-D uint8 tensor in pieces (works for NCCL and gloo tensors)."""
    n = t.numel()
    for a in range(0, n, PIECE):
        dist.broadcast(t[a:min(a + PIECE, n)], src=src)


def shard_blocks(rank, world, reads_per_rank, block):
    """Weak-scaling shard of the synthetic read stream: rank r owns the block range
    [r*B, (r+1)*B) with B = ceil(reads_per_rank / block).  -> (first_read, n_reads)."""
    b = (reads_per_rank + block - 1) // block
    return rank * b * block, reads_per_rank


def shard_chunks(n_chunks, rank, world):
    """Strong-scaling shard of an ordered chunk stream: chunk j goes to rank j mod world; the host
    concatenates per-chunk results in chunk order (keeps the reference's -t 1 ordering)."""
    return list(range(rank, n_chunks, world))


def broadcast_index(sh, src=0):
    """Replicates sh's index from rank `src` to every rank.  Returns the device time in ms."""
    rank = dist.get_rank()
    dev = torch.device("cuda", torch.cuda.current_device())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    info_t = pack_info(sh.info if rank == src else None).to(dev)
    dist.broadcast(info_t, src=src)
    if rank != src:
        sh.adopt_index(unpack_info(info_t))
    views = sh.index_views()
    torch.cuda.synchronize()
    e0.record()
    for i in range(len(views.bytes)):
        if views.bytes[i]:
            broadcast_bytes(view_tensor(views.dev_ptr[i], views.bytes[i], dev), src)
    e1.record()
    torch.cuda.synchronize()
    if rank != src:
        sh.finalize_index()
    return e0.elapsed_time(e1)


def shard_cuts(rec_off, n_shards):
    """Position cuts of the sharded build (shk_shard_cuts): shard s owns window ends [cuts[s], cuts[s+1])."""
    rec_off = np.ascontiguousarray(rec_off, dtype=np.uint64)
    cuts = np.zeros(n_shards + 1, np.uint64)
    rc = capi.load().shk_shard_cuts(capi.ptr(rec_off), len(rec_off) - 1, n_shards, capi.ptr(cuts))
    if rc:
        raise capi.SharkError(rc, capi.load().shk_last_error(None).decode())
    return cuts


def build_index_sharded(sh, bases, rec_off, barrier=None, all_gather=None):
    """Sharded index build, one process per GPU (SURVEY.md 8e second mode): rank r enumerates the
    k-mers of record shard r into its own filter, the filters are OR-merged by the library's P2P
    kernel over NVLink (peer buffers mapped through CUDA IPC handles that travel as bytes over
    torch.distributed), then every rank finishes the same index locally.  torch.distributed only
    carries the 240-byte handle structs and the barriers.  Returns (info, wall seconds); the host time
    of each step is left in build_index_sharded.last_steps_ms."""
    import time
    barrier = barrier or dist.barrier
    rank, world = dist.get_rank(), dist.get_world_size()
    steps = {}
    t0 = last = time.perf_counter()

    def mark(name):
        nonlocal last
        now = time.perf_counter()
        steps[name] = (now - last) * 1e3
        last = now

    mine = sh.shard_begin(bases, rec_off, rank, world)
    mark("begin")
    blobs = [None] * world
    (all_gather or dist.all_gather_object)(blobs, bytes(mine))
    mark("exchange")
    arr = (capi.ShardMem * world)()
    for s in range(world):
        peer = capi.ShardMem.from_buffer_copy(blobs[s])
        if peer.shard != s:
            raise capi.SharkError(-3, "shard structs arrived out of order")
        arr[s] = mine if s == rank else sh.shard_open(peer)
    mark("open")
    barrier()
    sh.shard_merge(1, arr)
    barrier()
    sh.shard_merge(2, arr)
    mark("merge")
    sh.shard_rank()
    barrier()
    mark("rank")
    info = sh.shard_finish(arr)
    barrier()
    mark("finish")
    for s in range(world):
        if s != rank:
            sh.shard_close(arr[s])
    barrier()  # nobody frees a buffer that a peer still has mapped
    sh.shard_end()
    mark("close")
    build_index_sharded.last_steps_ms = steps
    return info, time.perf_counter() - t0
