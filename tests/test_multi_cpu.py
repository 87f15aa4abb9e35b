"""world_size-2 tests of the multi-GPU host logic (vrdd_b200.dist) on CPU with gloo.

The compute of each rank is played by the oracle (decode of the rank's z-slab, rendering of the
rank's tiles); what is under test is the sharding rules and the two collectives: the in-place
all-gather of decoded z-slabs and the SUM reduction of partial frames."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, dims, img, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.vrdd_oracle import Oracle
        import vrdd_b200.dist as D
        o = Oracle(); o.set_num_threads(2)
        W, H, Dz = dims
        slice_vox = W * H
        # --- decode: my z-slab, then replicate the planes in place
        lo, hi = D.slab_range(Dz, rank, world)
        mine = o.decode_hist(o.synth_histograms(3, dims, z0=lo, nz=hi - lo))
        planes = []
        for c in range(3):
            full = torch.full((Dz * slice_vox,), float("nan"))
            full[lo * slice_vox:hi * slice_vox] = torch.from_numpy(mine[:, c].copy())
            planes.append(D.allgather_plane(full, Dz, slice_vox, rank, world))
        whole = o.decode_hist(o.synth_histograms(3, dims))
        for c in range(3):
            assert np.array_equal(planes[c].numpy(), whole[:, c]), ("plane", c)
        # --- render: my tiles of the frame, then assemble on rank 0
        vol4 = np.zeros((Dz * slice_vox, 4), np.float32)
        for c in range(3):
            vol4[:, c] = planes[c].numpy()
        view = o.view_matrix(15.0, 40.0)
        frame, _ = o.render(vol4, dims, view, image=img)
        owner = D.tile_owner(img[0], img[1], 16, 16, world).numpy()
        partial = torch.from_numpy(np.where(owner == rank, frame, 0).astype(np.uint32).view(np.int32).copy())
        scratch = torch.empty_like(partial)
        before = partial.clone()
        out = D.reduce_frame(partial, scratch, dst=0)
        assert torch.equal(partial, before)
        if rank == 0:
            assert np.array_equal(out.numpy().view(np.uint32), frame)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dims", [(12, 10, 8), (9, 7, 5)])       # even and uneven z split
def test_two_ranks_decode_allgather_and_tile_reduce(dims):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, dims, (96, 64), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def _sortlast_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import vrdd_b200.dist as D
        from oracle.vrdd_oracle import Oracle
        o = Oracle()
        grid = D.brick_grid(world)
        W, H = 40, 64
        for rot in ((0.0, 0.0), (30.0, 50.0), (-35.0, 200.0)):
            view = o.view_matrix(*rot)
            row0, rows, union = D.brick_row_windows(view, grid, H)
            # a pass-1 image that is non-zero exactly on the rows the window promises to cover
            g = torch.Generator().manual_seed(100 + rank)
            seg = torch.zeros(H, W)
            seg[row0[rank]:row0[rank] + rows] = torch.rand(rows, W, generator=g)
            buf = torch.empty(world * H * W)
            seg_rows = D.gather_row_windows(seg, buf, row0, rows, rank, world)
            full = torch.empty(world * H, W)
            dist.all_gather_into_tensor(full, seg)
            full = full.view(world, H, W)
            for b in range(world):                              # windows == the same rows of the whole images
                assert torch.equal(seg_rows[b], full[b, row0[b]:row0[b] + rows]), (rot, b)
            # increments: zero outside the union on every rank; reducing the union rows is reducing the frame
            part = torch.zeros(H, W, 4)
            part[union[0]:union[1]] = torch.rand(union[1] - union[0], W, 4, generator=g)
            ref = part.clone()
            dist.reduce(ref, dst=0, op=dist.ReduceOp.SUM)
            D.reduce_union_rows(part, union, dst=0)
            if rank == 0:
                assert torch.equal(part, ref), rot
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_two_ranks_sortlast_row_window_exchange():
    """The sort-last exchange restricted to row windows (dist.gather_row_windows / reduce_union_rows) moves the
    same data as whole-frame collectives, world size 2 (bricks 2x1x1) over gloo."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sortlast_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_sharding_rules():
    import vrdd_b200.dist as D
    for depth in (1, 7, 64, 1024):
        for world in (1, 2, 3, 4, 8):
            r = [D.slab_range(depth, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == depth and all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    assert [D.frame_size(1024, n) for n in (1, 2, 4, 8)] == [(1024, 1024), (2048, 1024), (2048, 2048), (4096, 2048)]
    for n in (1, 2, 4, 8):
        w, h = D.frame_size(1024, n)
        assert w * h == n * 1024 * 1024
        own = D.tile_owner(w, h, 64, 64, n)
        counts = torch.bincount(own.flatten().long(), minlength=n)
        assert counts.min() == counts.max() == 1024 * 1024               # every rank keeps 1024^2 pixels
    own = D.tile_owner(100, 70, 32, 16, 3)
    assert own[0, 0] == 0 and own[0, 32] == 1 and own[0, 96] == 0 and own[16, 0] == 1 and own[69, 99] == (4 * 4 + 3) % 3
