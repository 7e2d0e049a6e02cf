"""Frame-parallel multi-GPU plumbing (BASELINE configs[3]: 64 pairs over 8 GPUs).

Pairs are independent, so the path shards by frame with NO data-path collective
(SURVEY.md section 8e).  The only communication is the trivial scatter of the
packed input pyramids from rank 0 and the gather of the flows back, done with
torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests).  One
process per GPU; rank r owns the contiguous block shard_frames(n, world, r).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_frames(n_frames: int, world: int, rank: int):
    """Contiguous block [f0, f1) of rank `rank`; the first n % world ranks get one extra frame."""
    base, rem = divmod(n_frames, world)
    f0 = rank * base + min(rank, rem)
    return f0, f0 + base + (1 if rank < rem else 0)


def scatter_frames(packed_all, n_frames: int, frame_elems: int, device, dtype=torch.float32, group=None):
    """rank 0 passes `packed_all` (n_frames, frame_elems) on `device`; every rank returns its own
    block (f1-f0, frame_elems).  Point-to-point send/recv so blocks may have unequal sizes."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    f0, f1 = shard_frames(n_frames, world, rank)
    mine = torch.empty((f1 - f0, frame_elems), dtype=dtype, device=device)
    if rank == 0:
        reqs = []
        for r in range(1, world):
            a, b = shard_frames(n_frames, world, r)
            if b > a:
                reqs.append(dist.isend(packed_all[a:b].contiguous(), dst=r, group=group))
        mine.copy_(packed_all[f0:f1])
        for q in reqs:
            q.wait()
    elif f1 > f0:
        dist.recv(mine, src=0, group=group)
    return mine


def gather_flows(flows_local, n_frames: int, flow_elems: int, group=None):
    """Inverse of scatter_frames: rank 0 returns (n_frames, flow_elems), other ranks None."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if rank == 0:
        out = torch.empty((n_frames, flow_elems), dtype=flows_local.dtype, device=flows_local.device)
        f0, f1 = shard_frames(n_frames, world, 0)
        out[f0:f1].copy_(flows_local)
        for r in range(1, world):
            a, b = shard_frames(n_frames, world, r)
            if b > a:
                dist.recv(out[a:b], src=r, group=group)
        return out
    f0, f1 = shard_frames(n_frames, world, rank)
    if f1 > f0:
        dist.send(flows_local.contiguous(), dst=0, group=group)
    return None


def run_sharded(packed_all, n_frames, frame_elems, flow_elems, compute, device, group=None):
    """scatter -> compute(local_packed) -> gather.  `compute` maps (m, frame_elems) -> (m, flow_elems)
    on `device` (the C-ABI engine on GPUs; the tests inject the oracle)."""
    local = scatter_frames(packed_all, n_frames, frame_elems, device, group=group)
    flows = compute(local) if local.shape[0] else torch.empty((0, flow_elems), dtype=torch.float32, device=device)
    return gather_flows(flows, n_frames, flow_elems, group=group)
