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


class ShardedEngine:
    """BASELINE configs[3] as written: rank 0 owns all pairs, the pairs are scattered over the ranks'
    GPUs (NCCL), every rank runs the C-ABI engine on its block, the flows are gathered back to rank 0.

    io = "cli":     8-bit frames [pair][2][h][w][noc] in, full-resolution flow [pair][h][w][nop] out
                    (ofdis_upload_frames_u8 + ofdis_get_flow_fullres, both with device pointers):
                    what run_OF_INT computes between imread and SaveFlowFile
    io = "ofclass": un-padded float images of level sc_l in, flow of level sc_l out
                    (ofdis_upload_finest_level + ofdis_get_flow_batch): the OFClass region

    Everything is enqueued on `stream` (a torch.cuda.Stream, also the context's stream); rank 0's
    host buffers are pinned.  step() = H2D on rank 0 -> scatter -> run -> gather -> D2H on rank 0.
    """

    def __init__(self, prm, n_pairs: int, width_org: int, height_org: int, io: str, device, stream, group=None):
        from . import api  # the CUDA library; fails loudly when it is missing

        self.prm, self.n, self.io, self.group = prm, n_pairs, io, group
        self.w_org, self.h_org = width_org, height_org
        self.device, self.stream = device, stream
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.f0, self.f1 = shard_frames(n_pairs, self.world, self.rank)
        self.m = self.f1 - self.f0
        scf = 1 << prm.sc_f
        W, H = (width_org + scf - 1) // scf * scf, (height_org + scf - 1) // scf * scf
        self.ctx = api.Context(prm, W, H, prm.p_samp_s, max(self.m, 1), device=device.index or 0,
                               stream=stream.cuda_stream)
        self.ctx.set_graph_mode(True)
        if io == "cli":
            self.in_elems, self.in_dtype = 2 * height_org * width_org * prm.noc, torch.uint8
            self.out_elems = height_org * width_org * prm.nop
        else:
            self.in_elems, self.in_dtype = self.ctx.finest_level_frame_floats, torch.float32
            li = self.ctx.level_info(prm.sc_l)
            self.out_elems = li["w"] * li["h"] * prm.nop
        with torch.cuda.stream(stream):
            self.d_in = torch.empty((max(self.m, 1), self.in_elems), dtype=self.in_dtype, device=device)
            self.d_out = torch.empty((max(self.m, 1), self.out_elems), dtype=torch.float32, device=device)
            if self.rank == 0:
                self.d_all_in = torch.empty((n_pairs, self.in_elems), dtype=self.in_dtype, device=device)
                self.d_all_out = torch.empty((n_pairs, self.out_elems), dtype=torch.float32, device=device)
        self.equal = (n_pairs % self.world == 0)

    def step(self, host_in=None, host_out=None):
        """host_in / host_out: rank 0's pinned (n_pairs, in_elems) / (n_pairs, out_elems) tensors."""
        from . import api

        with torch.cuda.stream(self.stream):
            if self.rank == 0:
                self.d_all_in.copy_(host_in, non_blocking=True)
            if self.world == 1:
                local = self.d_all_in
            elif self.equal:
                chunks = list(self.d_all_in.view(self.world, self.m, self.in_elems).unbind(0)) if self.rank == 0 else None
                dist.scatter(self.d_in, chunks, src=0, group=self.group)
                local = self.d_in
            else:
                local = scatter_frames(self.d_all_in if self.rank == 0 else None, self.n, self.in_elems, self.device,
                                       dtype=self.in_dtype, group=self.group)
            if self.m:
                if self.io == "cli":
                    self.ctx.upload_frames_u8(0, self.m, local.data_ptr(), self.w_org, self.h_org, memkind=api.MEM_DEVICE)
                    self.ctx.run(self.m)
                    self.ctx.get_flow_fullres(0, self.m, self.d_out.data_ptr(), self.w_org, self.h_org, memkind=api.MEM_DEVICE)
                else:
                    self.ctx.upload_finest_level(0, self.m, local.data_ptr(), memkind=api.MEM_DEVICE)
                    self.ctx.run(self.m)
                    self.ctx.get_flow_batch(0, self.m, self.d_out.data_ptr(), memkind=api.MEM_DEVICE)
            flows = self.d_out[:self.m]
            if self.world == 1:
                out = flows
            elif self.equal:
                outs = list(self.d_all_out.view(self.world, self.m, self.out_elems).unbind(0)) if self.rank == 0 else None
                dist.gather(flows, outs, dst=0, group=self.group)
                out = self.d_all_out if self.rank == 0 else None
            else:
                out = gather_flows(flows, self.n, self.out_elems, group=self.group)
            if self.rank == 0 and host_out is not None:
                host_out.copy_(out, non_blocking=True)
        return out

    def close(self):
        self.ctx.close()
