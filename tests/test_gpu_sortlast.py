"""Sort-last rendering of a brick-decomposed volume (csrc/sortlast.cu), emulated on ONE GPU: the
bricks of a 2x2x2 (and 2x1x1, 3x2x1) decomposition are rendered one after the other by separate
handles, the collectives are played by torch ops, and the assembled frame must match the
single-volume render and the oracle within +-1 LSB — including rays that terminate early."""
import numpy as np
import pytest

gpu = pytest.mark.gpu


def _lsb(a, b):
    return np.abs(np.ascontiguousarray(a).view(np.uint8).astype(np.int16) -
                  np.ascontiguousarray(b).view(np.uint8).astype(np.int16))


def _render_bricks(V, D, oracle, gdims, grid, view, img, params, seed, hist_full, layout):
    import torch
    w, h = img
    nb = grid[0] * grid[1] * grid[2]
    handles, bricks = [], []
    for b in range(nb):
        q = (b % grid[0], (b // grid[0]) % grid[1], b // (grid[0] * grid[1]))
        origin, size, lo, hi = D.brick_geometry(gdims, grid, q)
        r = V.Renderer(0)
        r.set_sampler(layout)
        r.set_volume(*size)
        # the brick's histograms: generated on the device from GLOBAL voxel coordinates
        d_hist = torch.empty(size[0] * size[1] * size[2], 32, dtype=torch.float32, device="cuda")
        r.synth_histograms_region_device(seed, gdims, origin, 0, size[2], d_hist)
        r.synchronize()
        if hist_full is not None:
            sub = hist_full.reshape(gdims[2], gdims[1], gdims[0], 32)[origin[2]:origin[2] + size[2],
                                                                     origin[1]:origin[1] + size[1],
                                                                     origin[0]:origin[0] + size[0]].reshape(-1, 32)
            assert np.array_equal(d_hist.cpu().numpy(), sub)          # brick == sub-box of the whole, bit for bit
        r.set_histograms_device(d_hist, 0, size[2])
        r.decode(V.SRC_ORIGINAL)
        r.synchronize()
        del d_hist
        r.set_view(view)
        br = V.Brick(gdims[0], gdims[1], gdims[2], origin[0], origin[1], origin[2], (V.C.c_float * 3)(*lo),
                     (V.C.c_float * 3)(*hi))
        handles.append((r, q)); bricks.append(br)
    seg_all = torch.zeros(nb, h, w, dtype=torch.float32, device="cuda")
    for b, (r, q) in enumerate(handles):
        r.render_brick_alpha(seg_all[b], w, h, params, bricks[b])            # pass 1 (+ "all-gather")
    # the row windows the multi-GPU path gathers instead of whole images: conservative (no brick has alpha outside
    # its window), and the composition from the windows is the composition from the whole images, bit for bit
    row0, rows, (u0, u1) = D.brick_row_windows(view, grid, h)
    handles[0][0].synchronize()
    seg_rows = torch.zeros(nb, rows, w, dtype=torch.float32, device="cuda")
    for b in range(nb):
        outside = seg_all[b].clone()
        outside[row0[b]:row0[b] + rows] = 0
        assert float(outside.abs().max()) == 0.0, (grid, b, row0[b], rows)
        assert u0 <= row0[b] or float(seg_all[b][:u0].abs().max()) == 0.0
        seg_rows[b].copy_(seg_all[b][row0[b]:row0[b] + rows])
    total = torch.zeros(h, w, 4, dtype=torch.float32, device="cuda")
    samples = 0
    for b, (r, q) in enumerate(handles):
        a_in = torch.empty(h, w, dtype=torch.float32, device="cuda")
        part = torch.empty(h, w, 4, dtype=torch.float32, device="cuda")
        r.synchronize()
        r.compose_alpha_in(seg_all, grid, q, a_in, w, h)
        a_win = torch.empty(h, w, dtype=torch.float32, device="cuda")
        r.compose_alpha_in_rows(seg_rows, grid, q, row0, rows, a_win, w, h)
        r.synchronize()
        assert torch.equal(a_win, a_in)
        r.count_samples(True)
        r.render_brick_color(a_in, part, w, h, params, bricks[b])            # pass 2
        r.synchronize()
        samples += r.get_sample_count()
        assert float(part[:u0].abs().max() if u0 > 0 else 0.0) == 0.0 and float(part[u1:].abs().max() if u1 < h else 0.0) == 0.0
        total += part                                                        # "reduce (SUM)" (of the union rows)
    out = torch.zeros(h, w, dtype=torch.int32, device="cuda")
    handles[0][0].pack_frame(total, out, w, h, params.brightness)
    handles[0][0].synchronize()
    # The direct-send form of the same frame (vrdd_render_brick_alpha_send / _color_send / vrdd_pack_frame_slots): every
    # "rank" stores its window rows into all ranks' tables and bumps their counters, waits on its own counter from its
    # stream, composes, stores its increments into the root's table; the root sums the slots in brick order.  Twice, to
    # exercise the counter generations.  Must be the frame of the collective form, bit for bit.
    root = handles[0][0]
    seg_tabs = [torch.zeros(nb * rows * w, dtype=torch.float32, device="cuda") for _ in range(nb)]
    flags = [torch.zeros(16, dtype=torch.int32, device="cuda") for _ in range(nb)]
    slots = torch.zeros(nb * rows * w * 4, dtype=torch.float32, device="cuda")
    root_flag = torch.zeros(16, dtype=torch.int32, device="cuda")
    # band owners (vrdd_render_brick_color_send_bands / vrdd_pack_band_slots): "rank" o owns rows [o * band, (o + 1) * band)
    band = (h + nb - 1) // nb
    owner_tabs = [torch.zeros(nb * band * w * 4, dtype=torch.float32, device="cuda") for _ in range(nb)]
    owner_flags = [torch.zeros(16, dtype=torch.int32, device="cuda") for _ in range(nb)]
    frame_flag = torch.zeros(16, dtype=torch.int32, device="cuda")
    # Generation 1 runs while samples are counted (pass 2 marches every pixel), 2 with the fused first segment (pass 1 keeps
    # the colour of the march from alpha 0 and pass 2 forwards it where the incoming alpha is 0), 3 with the fusion off.
    for gen in (1, 2, 3):
        for b, (r, q) in enumerate(handles):
            r.count_samples(gen == 1)
            r.set_variant("sortlast_fuse", "off" if gen == 3 else "on")
            r.set_view(view)
            r.render_brick_alpha_send(seg_tabs, flags, b, row0[b], rows, w, h, params, bricks[b])
        for b, (r, q) in enumerate(handles):
            a_in = torch.empty(h, w, dtype=torch.float32, device="cuda")
            r.stream_wait_flag(flags[b], nb * gen)
            r.compose_alpha_in_rows(seg_tabs[b], grid, q, row0, rows, a_in, w, h)
            r.render_brick_color_send(a_in, slots, root_flag, b, row0[b], rows, w, h, params, bricks[b])
            r.render_brick_color_send_bands(a_in, owner_tabs, owner_flags, band, b, row0[b], rows, w, h, params, bricks[b])
            r.synchronize()
        out2 = torch.zeros(h, w, dtype=torch.int32, device="cuda")
        root.stream_wait_flag(root_flag, nb * gen)
        root.pack_frame_slots(slots, nb, row0, rows, out2, w, h, params.brightness)
        root.synchronize()
        assert torch.equal(out2, out), (grid, gen, int((out2 != out).sum()))
        assert int(root_flag[0]) == nb * gen and all(int(f[0]) == nb * gen for f in flags)
        out3 = torch.full((h, w), -1, dtype=torch.int32, device="cuda")
        for o, (r, q) in enumerate(handles):                   # every owner sums + packs its band into the "root's" frame
            r.stream_wait_flag(owner_flags[o], nb * gen)
            r.pack_band_slots(owner_tabs[o], nb, row0, rows, o, band, out3, frame_flag, w, h, params.brightness)
            r.synchronize()
        root.stream_wait_flag(frame_flag, nb * gen)
        root.synchronize()
        assert torch.equal(out3, out), (grid, gen, int((out3 != out).sum()))
    for r, _ in handles:
        r.close()
    return out.cpu().numpy().view(np.uint32), samples


@gpu
@pytest.mark.parametrize("grid", [(2, 2, 2), (2, 1, 1), (3, 2, 1)])
@pytest.mark.parametrize("rot", [(0.0, 0.0), (25.0, 40.0), (-35.0, 200.0), (90.0, 0.0)])
@pytest.mark.parametrize("store", ["texture", "planes"])
def test_sortlast_matches_single_volume_and_oracle(oracle, grid, rot, store):
    import vrdd_b200 as V
    import vrdd_b200.dist as D
    gdims, img, seed = (24, 20, 16), (96, 80), 31
    hist = oracle.synth_histograms(seed, gdims)
    vol = oracle.decode_hist(hist)
    view = oracle.view_matrix(*rot)
    for over in ({}, {"density": 0.3, "opacity_threshold": 0.8, "brightness": 1.3}):     # default, and heavy early exit
        params = V.default_render_params(query_method=1, **over)
        ref, ref_s = oracle.render(vol, gdims, view, image=img, density=params.density, brightness=params.brightness,
                                   opacity_threshold=params.opacity_threshold)
        # bricks as 3-D arrays filtered by the texture unit (the default), or as planes in global memory
        # (both plane layouts) filtered by the integer restatement of the unit
        layout = V.SAMPLER_TEXTURE if store == "texture" else (V.SAMPLER_BRICKED if over else V.SAMPLER_LINEAR)
        got, s = _render_bricks(V, D, oracle, gdims, grid, view, img, params, seed, hist, layout)
        d = _lsb(got, ref)
        assert d.max() <= 1, (grid, rot, over, int(d.max()), int((d > 1).sum()))
        assert abs(s - ref_s) <= max(2, ref_s // 5000), (s, ref_s)
        # single-volume GPU render of the same thing
        r = V.Renderer(0)
        r.set_volume(*gdims); r.set_histograms_host(hist); r.decode(V.SRC_ORIGINAL); r.set_view(view)
        import torch
        one = torch.zeros(img[1], img[0], dtype=torch.int32, device="cuda")
        r.render(one, img[0], img[1], params, clear_misses=True); r.synchronize()
        assert _lsb(got, one.cpu().numpy().view(np.uint32)).max() <= 1
        r.close()


def test_brick_row_windows_are_consistent():
    """Host logic of the row windows (no GPU): one common row count, windows inside the frame and inside the union,
    the whole frame for a matrix that is not a rotation, and symmetric windows for the frontal view."""
    import vrdd_b200.dist as D
    import math
    def view(rx, ry, tz=4.0):                                 # Rx(-rx) Ry(-ry) T(0,0,tz), rows as in volumeRender.cpp:229-246
        ax, ay = math.radians(-rx), math.radians(-ry)
        cx, sx, cy, sy = math.cos(ax), math.sin(ax), math.cos(ay), math.sin(ay)
        R = [[cy, 0.0, sy], [sx * sy, cx, -sx * cy], [-cx * sy, sx, cx * cy]]
        return [R[0][0], R[0][1], R[0][2], R[0][2] * tz, R[1][0], R[1][1], R[1][2], R[1][2] * tz,
                R[2][0], R[2][1], R[2][2], R[2][2] * tz]
    H = 512
    for grid in ((2, 2, 2), (2, 1, 1), (3, 2, 1), (1, 1, 1)):
        for rot in ((0.0, 0.0), (25.0, 40.0), (-35.0, 200.0), (90.0, 0.0), (0.0, 45.0)):
            row0, rows, (u0, u1) = D.brick_row_windows(view(*rot), grid, H)
            assert len(row0) == grid[0] * grid[1] * grid[2] and 1 <= rows <= H and 0 <= u0 < u1 <= H
            assert all(0 <= y and y + rows <= H for y in row0)
            assert u1 - u0 < H                                   # the box never fills the frame from this distance
    row0, rows, (u0, u1) = D.brick_row_windows(view(0.0, 0.0), (2, 2, 2), H)
    assert abs((u0 + u1) - H) <= 2 and rows < 0.45 * H           # frontal view: centred, a brick spans about a third of the rows
    sheared = view(0.0, 0.0); sheared[1] = 0.3
    assert D.brick_row_windows(sheared, (2, 2, 2), H)[1:] == (H, (0, H))
    assert D.brick_row_windows(view(0.0, 0.0, tz=0.5), (2, 2, 2), H)[1] == H      # eye inside the box: whole frame


def test_brick_geometry_partitions_the_volume():
    import vrdd_b200.dist as D
    for gdims, grid in (((24, 20, 16), (2, 2, 2)), ((2048, 2048, 2048), (2, 2, 2)), ((50, 50, 10), (3, 2, 1))):
        owned = 0
        for b in range(grid[0] * grid[1] * grid[2]):
            q = (b % grid[0], (b // grid[0]) % grid[1], b // (grid[0] * grid[1]))
            origin, size, lo, hi = D.brick_geometry(gdims, grid, q)
            n = 1
            for ax in range(3):
                b0 = gdims[ax] * q[ax] // grid[ax]; b1 = gdims[ax] * (q[ax] + 1) // grid[ax]
                assert origin[ax] == max(b0 - 1, 0) and origin[ax] + size[ax] == min(b1 + 1, gdims[ax])
                assert (lo[ax] == float("-inf")) == (q[ax] == 0) and (hi[ax] == float("inf")) == (q[ax] == grid[ax] - 1)
                n *= b1 - b0
            owned += n
        assert owned == gdims[0] * gdims[1] * gdims[2]
    assert [D.brick_grid(n) for n in (1, 2, 4, 8)] == [(1, 1, 1), (2, 1, 1), (2, 2, 1), (2, 2, 2)]
    assert [D.brick_of_rank(r, (2, 2, 2)) for r in (0, 1, 2, 5, 7)] == [(0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 0, 1), (1, 1, 1)]
