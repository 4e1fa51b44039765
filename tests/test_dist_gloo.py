"""Multi-rank host logic on CPU (gloo, world_size 2, 3 and 4): work assignment, the integer (and, for
non-power-of-two bins, float64) all-reduce of stage 1, the chunk-block all-gather of stage 2 and the row gather of
stage 3 and of the ROI-processor outputs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from facemap_b200 import parallel


def test_assignments_cover_everything_once():
    for world in (1, 2, 3, 4, 8):
        for n in (1, 2, 10, 15):
            got = sorted(i for r in parallel.assign_round_robin(n, world) for i in r)
            assert got == list(range(n))
            runs = parallel.assign_contiguous(n, world)
            assert [i for r in runs for i in r] == list(range(n)) and len(runs) == world
            per = -(-n // world)
            assert all(r == list(range(q * per, min(n, (q + 1) * per))) for q, r in enumerate(runs))
        for N in (1, 499, 500, 501, 2300, 100000, 800000):
            rg = parallel.frame_ranges(N, world)
            assert rg[0][0] == 0 and rg[-1][1] == N and len(rg) == world
            for (a, b), (c, d) in zip(rg[:-1], rg[1:]):
                assert b == c and a <= b
                assert b % 500 == 0 or b == N          # boundaries sit on the reference's 500-frame grid
    assert parallel.frame_ranges(100000, 8) == [(i * 12500, (i + 1) * 12500) for i in range(8)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sh = parallel.Shard()
        p, nseg = 37, 10
        rng = np.random.RandomState(123)
        sums_all = rng.randint(0, 10 ** 9, size=(nseg, 2, p)).astype(np.int64)
        mine = parallel.assign_round_robin(nseg, world)[rank]
        local = {i: (torch.from_numpy(sums_all[i, 0]), torch.from_numpy(sums_all[i, 1]), 100) for i in mine}
        full = sh.allreduce_segment_sums(local, nseg, p, torch.device("cpu"))
        ok1 = all(np.array_equal(full[i][0].numpy(), sums_all[i, 0]) and np.array_equal(full[i][1].numpy(), sums_all[i, 1])
                  and full[i][2] == 100 for i in range(nseg))
        # stage-2 block gather: 7 chunks of k=3 rows over `world` ranks
        nsegs, k, cols = 7, 3, 11
        blocks = rng.rand(nsegs * k, cols).astype(np.float32)
        mine = parallel.assign_round_robin(nsegs, world)[rank]
        loc = torch.from_numpy(np.concatenate([blocks[c * k:(c + 1) * k] for c in mine], axis=0)) if mine \
            else torch.zeros((0, cols))
        ok2 = np.array_equal(sh.gather_chunk_blocks(loc, k, nsegs).numpy(), blocks)
        # the product's form: consecutive chunks per rank, written at their final rows, gathered in place
        ks, sizes = [k, 2], [cols, 5]
        bases, windows, per = sh.chunk_buffers(nsegs, ks, sizes, ["mot"], True, torch.device("cpu"))
        blocks_roi = rng.rand(nsegs * 2, 5).astype(np.float32)
        mine_c = parallel.assign_contiguous(nsegs, world)[rank]
        for slot, c in enumerate(mine_c):
            windows["mot"][0][slot * k:(slot + 1) * k] = torch.from_numpy(blocks[c * k:(c + 1) * k])
            windows["mot"][1][slot * 2:(slot + 1) * 2] = torch.from_numpy(blocks_roi[c * 2:(c + 1) * 2])
        g0 = sh.gather_in_place(bases["mot"][0], per * k)[:nsegs * k, :cols]
        g1 = sh.gather_in_place(bases["mot"][1], per * 2)[:nsegs * 2, :5]
        ok2 = ok2 and np.array_equal(g0.numpy(), blocks) and np.array_equal(g1.numpy(), blocks_roi) \
            and g0.stride(0) % 32 == 0
        # stage-3 row gather
        N = 2300
        rows = rng.rand(N, 5).astype(np.float32)
        ranges = parallel.frame_ranges(N, world)
        a, b = ranges[rank]
        ok3 = np.array_equal(sh.gather_rows(torch.from_numpy(rows[a:b]), ranges).numpy(), rows)
        # stage 1 of a non-power-of-two sbin: float64 segment means, each from exactly one rank (x + 0 == x)
        means_all = rng.rand(nseg, 2, p) * 255 + rng.rand(nseg, 2, p) * 1e-9
        mine = parallel.assign_round_robin(nseg, world)[rank]
        local = {i: (torch.from_numpy(means_all[i, 0]), torch.from_numpy(means_all[i, 1]), 100) for i in mine}
        full = sh.allreduce_segment_sums(local, nseg, p, torch.device("cpu"), dtype=torch.float64)
        ok4 = all(full[i][0].dtype == torch.float64 and np.array_equal(full[i][0].numpy(), means_all[i, 0])
                  and np.array_equal(full[i][1].numpy(), means_all[i, 1]) for i in range(nseg))
        # outputs of the ROI processors / fm_motion_ref: float64 [n, 9] (with NaN rows), int32 [n, 2], float64 [n]
        pup = rng.rand(N, 9)
        pup[rng.rand(N) < 0.05] = np.nan
        run = rng.randint(-40, 41, size=(N, 2)).astype(np.int32)
        eref = rng.rand(N) * 1e5
        ok5 = (np.array_equal(sh.gather_rows(torch.from_numpy(pup[a:b]), ranges).numpy(), pup, equal_nan=True)
               and np.array_equal(sh.gather_rows(torch.from_numpy(run[a:b]), ranges).numpy(), run)
               and np.array_equal(sh.gather_rows(torch.from_numpy(eref[a:b]), ranges).numpy(), eref))
        # masks of a merged target from the rank that solved it: a row-padded view (kernels.alloc_matrix layout)
        k_, p_ = 5, 37
        buf = torch.zeros((k_, 64))
        Ut = buf[:, :p_]
        sv = torch.zeros(k_)
        want_U, want_s = rng.rand(k_, p_).astype(np.float32), rng.rand(k_).astype(np.float32)
        src = world - 1
        if rank == src:
            Ut.copy_(torch.from_numpy(want_U))
            sv.copy_(torch.from_numpy(want_s))
        Ut2, sv2 = sh.broadcast_merge(Ut, sv, src)
        ok6 = np.array_equal(Ut2.numpy(), want_U) and np.array_equal(sv2.numpy(), want_s) and Ut2.data_ptr() == buf.data_ptr()
        out[rank] = (ok1, ok2, ok3, ok4, ok5, ok6)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3, 4])
def test_shard_collectives_gloo(world):
    port = _free_port()
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    out = mgr.dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(120)
        assert pr.exitcode == 0
    assert len(out) == world and all(all(v) for v in out.values()), dict(out)
