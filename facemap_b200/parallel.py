"""Frame-range sharding of the path across the GPUs of one box (one process per GPU,
torch.distributed; NCCL on GPUs, gloo in the CPU tests).

The reference is single-process; this is the one parallelism strategy the path
admits:
  stage 1  the 10 mean segments are dealt round-robin; their INTEGER sums are all-reduced (exact,
           order independent) and the float32 accumulation is replayed in segment order on every
           rank, so avgframe/avgmotion are bit-identical to the single-GPU result;
  stage 2  the <= 15 SVD chunks are dealt in consecutive runs; every rank writes its component blocks (U S)^T at
           their FINAL rows of the stacked buffer and one in-place all-gather fills in the others' (no staging
           copies); the merge eigensolve is replicated (deterministic);
  stage 3  contiguous frame ranges aligned to the reference's 500-frame grid, one-frame halo, no
           communication; the projections / energies are all-gathered at the end.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import svd


def assign_round_robin(n_items, world):
    """item i -> rank i % world."""
    return [list(range(r, n_items, world)) for r in range(world)]


def assign_contiguous(n_items, world):
    """consecutive items per rank, ceil(n / world) each (the last ranks may hold fewer or none): rank r's items sit at
    positions [r * per, (r + 1) * per) of the gathered order, so an all-gather can run in place."""
    per = -(-n_items // world) if n_items else 0
    return [list(range(r * per, min(n_items, (r + 1) * per))) for r in range(world)]


def frame_ranges(N, world, align=500):
    """`world` contiguous [f0, f1) ranges covering [0, N), boundaries on multiples of `align`
    (the reference's projection chunk grid, process.py:394), as even as that allows."""
    nblocks = -(-N // align)
    cuts = [min(N, align * ((nblocks * r) // world)) for r in range(world)] + [N]
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


class Shard:
    def __init__(self, group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def any_rank(self, flag, device):
        """True on every rank if `flag` is set on any (one tiny all-reduce): decisions every rank must take together"""
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return bool(t.item())

    def broadcast_merge(self, Ut, s, src):
        """masks (U^T [k, p]: a view of a row-padded buffer, kernels.alloc_matrix) and singular values of one merged
        target from the rank that solved it; the padded storage travels as it is"""
        whole = Ut if Ut.is_contiguous() else Ut._base
        if whole is not None and whole.is_contiguous() and whole.data_ptr() == Ut.data_ptr():
            dist.broadcast(whole, src=src, group=self.group)
        else:                                                  # some other kind of view: through a packed copy
            tmp = Ut.contiguous()
            dist.broadcast(tmp, src=src, group=self.group)
            if self.rank != src:
                Ut.copy_(tmp)
        dist.broadcast(s, src=src, group=self.group)
        return Ut, s

    # ---- collectives on plain tensors (device agnostic: used by the gloo CPU tests too) ----
    def allreduce_segment_sums(self, local, nseg, p, device, dtype=torch.int64):
        """local: {segment index: (tsum_bin [p], tsum_adiff [p], T)} for the segments this rank computed -> the
        same dict for ALL segments.  int64 time sums (exact in any order) or, for a non-power-of-two sbin, float64
        segment means — each segment comes from exactly one rank, so the SUM all-reduce only adds zeros to it."""
        pack = torch.zeros((nseg, 2, p), dtype=dtype, device=device)
        counts = torch.zeros(nseg, dtype=torch.int64, device=device)
        for i, (tsb, tsa, T) in local.items():
            pack[i, 0] = tsb
            pack[i, 1] = tsa
            counts[i] = T
        dist.all_reduce(pack, op=dist.ReduceOp.SUM, group=self.group)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=self.group)
        counts = counts.tolist()
        return {i: (pack[i, 0], pack[i, 1], int(counts[i])) for i in range(nseg)}

    def gather_chunk_blocks(self, local, k, nsegs):
        """local: [n_local * k, p] blocks of the chunks assign_round_robin gives this rank (in that
        order) -> [nsegs * k, p] in chunk order on every rank (one all-gather, padded to equal size)."""
        p = local.shape[1]
        per = -(-nsegs // self.world)
        send = torch.zeros((per * k, p), dtype=local.dtype, device=local.device)
        send[:local.shape[0]] = local
        recv = torch.empty((self.world * per * k, p), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(recv, send, group=self.group)
        out = torch.empty((nsegs * k, p), dtype=local.dtype, device=local.device)
        owners = assign_round_robin(nsegs, self.world)
        for r, items in enumerate(owners):
            for slot, ci in enumerate(items):
                src = (r * per + slot) * k
                out[ci * k:(ci + 1) * k] = recv[src:src + k]
        return out

    def gather_rows(self, local, ranges):
        """local rows of this rank's frame range -> all rows [N, ...] on every rank."""
        longest = max(b - a for a, b in ranges)
        tail = tuple(local.shape[1:])
        send = torch.zeros((longest,) + tail, dtype=local.dtype, device=local.device)
        send[:local.shape[0]] = local
        recv = torch.empty((self.world * longest,) + tail, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(recv, send, group=self.group)
        return torch.cat([recv[r * longest:r * longest + (b - a)] for r, (a, b) in enumerate(ranges)], dim=0)

    # ---- the three stages ----
    def stage1(self, source, geo, device, timer):
        N = source.nframes
        _, tf = svd.stage1_plan(N, source.block_frames)
        mine = assign_round_robin(len(tf), self.world)[self.rank]
        local, nseg = svd.stage1_means(source, geo, device, timer, segments=mine)
        with timer("stage1.allreduce"):
            dt = torch.float64 if svd.stage1_uses_reference(geo.sbin) else torch.int64
            sums = self.allreduce_segment_sums(local, nseg, geo.p, device, dtype=dt)
        return svd.stage1_finish(sums, nseg, geo, device)

    def chunk_buffers(self, nsegs, ks, sizes, mods, fullSVD, device):
        """Per (modality, target) the stacked component buffer [world * per * k, ld] every rank allocates alike, and
        this rank's window on it: rows [rank * per * k, ...) are where its chunks' blocks belong."""
        per = -(-nsegs // self.world)
        bases, mine = {}, {}
        for m in mods:
            bases[m], mine[m] = [], []
            for idx, (k, p) in enumerate(zip(ks, sizes)):
                if idx == 0 and not fullSVD:
                    bases[m].append(None)
                    mine[m].append(None)
                    continue
                ld = max(32, svd.K.round_up(p, 32))
                base = torch.empty((self.world * per * k, ld), dtype=torch.float32, device=device)
                bases[m].append(base)
                mine[m].append(base[self.rank * per * k:(self.rank + 1) * per * k, :p])
        return bases, mine, per

    def gather_in_place(self, base, rows_per_rank):
        """all-gather of a buffer whose rows [r * rows_per_rank, (r + 1) * rows_per_rank) rank r has filled"""
        seg = base[self.rank * rows_per_rank:(self.rank + 1) * rows_per_rank]
        dist.all_gather_into_tensor(base, seg, group=self.group)
        return base

    def stage2_chunks(self, source, geo, avgframe, avgmotion, mods, fullSVD, ws, device, timer):
        nt0, nsegs, nc, _ = svd.stage2_plan(source.nframes)
        mine = assign_contiguous(nsegs, self.world)[self.rank]
        sizes = [geo.p] + geo.roi_npix
        ks = [min(nc, s) for s in sizes]
        bases, windows, per = self.chunk_buffers(nsegs, ks, sizes, mods, fullSVD, device)
        # every rank's blocks land at their final rows; the gather then needs no padding, reordering or copy
        svd.stage2_chunks(source, geo, avgframe, avgmotion, mods, fullSVD, ws, device, timer, chunks=mine, Yt_out=windows)
        full = {}
        with timer("stage2.allgather"):
            for m in mods:
                full[m] = []
                for idx, base in enumerate(bases[m]):
                    if base is None:
                        full[m].append(None)
                        continue
                    self.gather_in_place(base, per * ks[idx])
                    full[m].append(base[:nsegs * ks[idx], :sizes[idx]])
        return full, ks, nsegs

    def gather_stage3(self, V, E, E_roi, N, mods, timer):
        """All-gather the per-rank projections / energies of frame_ranges(N, world) into full arrays."""
        ranges = frame_ranges(N, self.world)
        with timer("stage3.allgather"):
            Vf = {m: [None if v is None else self.gather_rows(v, ranges) for v in V[m]] for m in mods}
            Ef = self.gather_rows(E.t().contiguous(), ranges).t().contiguous()
            Erf = self.gather_rows(E_roi.t().contiguous(), ranges).t().contiguous()
        return Vf, Ef, Erf
