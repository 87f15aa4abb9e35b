#!/usr/bin/env python
"""bench.py — headline benchmark of the two hot paths on B200 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload), all inputs synthetic and generated on the device (include/vrdd_synth.h):
  * decode: a VOL^3 distribution volume (default 1024^3 x 32 bins = 137 GB of raw histograms) is decoded z-slab by
    z-slab (default 256 slices = 34 GB per launch); every slab is resident in HBM before its timed region starts.
    Reported in `decode` and `roofline_decode` (GB/s against the measured HBM peak), with the fractal-code decode of
    the same volume next to it.
  * ray cast (the `metric`): a "step" renders one IMG x IMG view (default 1024^2) of the decoded volume with the
    reference's constants (tstep 0.01, 500 steps, threshold 0.95, density 0.05, rainbow transfer function,
    queryMethod 1).  The K timed views are ALWAYS spread over the whole 64-view orbit of BASELINE.json configs[1]
    (view k * 64 / K), whatever K is, and the sample counts, the device-timed loop, the kernel-only loop and the
    end-to-end loop all use exactly those views.  Gsamples/s = transfer-function lookups / time; the lookups are
    counted exactly by the kernel in an untimed pass over the same views.
  * N > 1 (one process per GPU, torchrun): every rank decodes VOL/N z-slices and the decode kernel itself stores its
    slab into every other rank's planes over NVLink (vrdd_set_peer_planes); image-space tiles (64x64, round-robin over
    ranks); the kernels store their tiles straight into rank 0's frame over NVLink and the last block of each launch
    bumps a counter next to the frame that rank 0's stream waits on — no host barrier per frame.  Two records:
    weak scaling (the frame grows with N so every GPU keeps IMG^2 pixels: the top-level line) and `strong` (BASELINE
    configs[2]: a fixed 2048 x 2048 frame of the same volume at every N).  `sortlast` (configs[4]): one VOL^3 brick per
    rank of a volume too large for one GPU.
Timing: CUDA events on the launching stream, >= 3 warm-up steps, barrier + synchronize on both sides, max over ranks.
The sampled plane (4.3 GB) and every decode slab (>= 8 GB) are far larger than the 126 MB L2, so no L2 flush is needed.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HIST_BYTES_PER_VOXEL = 128 + 12          # 32 fp32 bins read, three fp32 planes written
SAMPLE_BYTES = 32                        # 8 fp32 texels per trilinear sample
ORBIT_VIEWS = 64
HBM_FALLBACK_GBS = 6650.0                # /opt/skills/guides/B200_PROFILING.md
STRONG_IMAGE = 2048                      # BASELINE.json configs[2]
L1TEX_FALLBACK_GSAMPLES = 559.1          # profiles/l1tex_peak_r2.txt (tools/l1tex_peak.cu on a B200 of this pool), when not measured live


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return {"hbm_gbs": HBM_FALLBACK_GBS}, "fallback (B200_PROFILING.md)"


def traffic_record(key):
    """DRAM bytes per launch of a kernel / launch shape from the committed ncu captures (profiles/traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(key)
    except Exception:
        return None


def ncu_traffic(kernel, bytes_per_launch):
    t = traffic_record(kernel)
    if t and abs(t["algorithmic_bytes_per_launch"] - bytes_per_launch) < 1e-6 * bytes_per_launch:
        return t["dram_bytes_per_launch"]
    return None


def ray_roofline(samples_per_launch, kernel_ms, volume_edge, fw, fh, hbm_peak, peak_src, traffic, tex_peak):
    """which roofline bounds the ray caster depends on how dense the rays are in the volume.  Rays more than
    one voxel apart (the headline: 1024^3 at 1024^2, two voxels): every sample brings its own texels, S * 32 B / t against
    the HBM peak.  Rays one voxel apart or closer (2048^2 frames, the larger frames of N > 1): neighbouring rays share
    their texels in L1 / L2, DRAM moves a fraction of 32 B per sample, and the 32-byte figure is not a bound (it exceeds
    the HBM peak); the binding limit is the texture pipe: S / t against its measured fetch rate.  The other figure is kept
    next to it for information."""
    spacing = 2.0 * volume_edge / max(fw, fh)
    gs = samples_per_launch / (kernel_ms * 1e-3) / 1e9
    hbm = {"achieved": samples_per_launch * SAMPLE_BYTES / (kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s", "peak_source": peak_src}
    hbm["frac"] = hbm["achieved"] / hbm_peak
    live = bool(tex_peak and "tex3d_linear_l1_gsamples_per_s" in tex_peak)
    l1_peak = tex_peak["tex3d_linear_l1_gsamples_per_s"] if live else L1TEX_FALLBACK_GSAMPLES
    l1 = {"achieved": gs, "peak": l1_peak, "unit": "Gsamples/s (one trilinear fp32 fetch each)", "frac": gs / l1_peak,
          "peak_source": "tools/l1tex_peak.cu, run in this process" if live else "profiles/l1tex_peak_r2.txt (tools/l1tex_peak.cu)"}
    if live:
        l1["all"] = tex_peak
    r = {"launch_ms": kernel_ms, "bytes_per_launch": samples_per_launch * SAMPLE_BYTES, "gsamples_per_s_kernel_only": gs,
         "ray_spacing_voxels": spacing, "traffic": traffic}
    if spacing > 1.0:
        r.update({"bound": "hbm", **hbm, "l1tex": l1})
    else:
        r.update({"bound": "l1tex", **l1, "hbm_algorithmic": dict(hbm, note="not a bound at this ray density: neighbouring rays share "
                                                                            "their texels on chip (see traffic for what DRAM moved)")})
    if traffic:
        r["traffic_gbs"] = traffic / (kernel_ms * 1e-3) / 1e9
        r["traffic_frac"] = r["traffic_gbs"] / hbm_peak
    return r


def l1tex_peak():
    """The texture pipe's own peak for the ray caster's operation, measured live (tools/l1tex_peak.cu: one trilinear
    fp32 tex3D fetch per sample, nothing else in the loop): SURVEY.md §8d's L1TEX roofline."""
    exe = os.path.join(ROOT, "tools", "l1tex_peak")
    try:
        out = subprocess.run([exe, "json"], capture_output=True, text=True, timeout=120).stdout.strip().splitlines()
        return json.loads(out[0])
    except Exception as exc:
        return {"error": repr(exc)}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed regions (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons = [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); out["sm_max_mhz"] = float(f[2])
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            out["sm_mhz"] = sm[len(sm) // 2]
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def timed_views(steps):
    """The K views a run times: spread over the whole 64-view orbit whatever K is (K >= 64: the orbit, repeated)."""
    if steps >= ORBIT_VIEWS:
        return [k % ORBIT_VIEWS for k in range(steps)]
    return [(k * ORBIT_VIEWS) // steps for k in range(steps)]


def orbit_view(V, k):
    """View k of the 64-view orbit: viewRotation.y = k * 5.625 deg, translation (0,0,-4)
    (volumeRender.cpp:126, 229-246)."""
    return V.view_matrix(0.0, (k % ORBIT_VIEWS) * (360.0 / ORBIT_VIEWS), (0.0, 0.0, -4.0))


# ---------------------------------------------------------------------------------------------
# CPU legs (oracle).  bench.py may execute oracle/ only here: as the timed CPU baseline.
# ---------------------------------------------------------------------------------------------

def host_threads():
    """Threads the CPU legs use: every core this process may run on.  Set explicitly because torchrun exports
    OMP_NUM_THREADS=1 to its workers, which would silently make the reference arm single-threaded."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_decode_baseline(seed):
    """Raw-histogram decode on the host cores: a 256x256x64 slab of the same synthetic volume."""
    from oracle.vrdd_oracle import Oracle
    o = Oracle(fast=True)
    o.set_num_threads(host_threads())
    hist = o.synth_histograms(seed, (256, 256, 256), z0=96, nz=64)
    best = 1e30
    for _ in range(3):
        t0 = time.perf_counter(); o.decode_hist(hist); best = min(best, time.perf_counter() - t0)
    nvox = hist.shape[0]
    return {"value": nvox * HIST_BYTES_PER_VOXEL / best / 1e9, "unit": "GB/s", "cores": o.num_threads(),
            "kind": "port", "sample": "256x256x64 slab (4.2 M voxels) of the synthetic volume, best of 3, "
            "OpenMP oracle -O3 -march=x86-64-v3", "mvoxels_per_s": nvox / best / 1e6}


class CpuRaycaster:
    """The oracle ray caster (OpenMP port of the reference's d_render, the reference build's rounding) on the SAME
    workload as the GPU arm: the EDGE^3 synthetic volume (default 1024^3, 1.07 G voxels, 137 GB of histograms generated
    and decoded on the host 4 M voxels at a time — about half a minute on 16 cores, untimed like the GPU arm's decode);
    only the plane queryMethod 1 samples is kept (4.3 GB).  Frames are the arm's own (IMG_W x IMG_H views of the same
    orbit, the reference's constants)."""

    def __init__(self, seed, img, edge=512):
        import numpy as np
        from oracle.vrdd_oracle import Oracle
        self.o = Oracle(fast=True)
        self.o.set_reference_build(True)
        self.o.set_num_threads(host_threads())
        self.edge = edge
        self.dims = (edge, edge, edge)
        self.img = img
        plane = np.empty(edge ** 3, np.float32)
        sl = edge * edge
        step = max(1, min(edge, (1 << 22) // sl))                 # ~4 M voxels (0.5 GB of histograms) at a time
        t0 = time.perf_counter()
        for z0 in range(0, edge, step):
            nz = min(step, edge - z0)
            plane[z0 * sl:(z0 + nz) * sl] = self.o.decode_hist(self.o.synth_histograms(seed, self.dims, z0=z0, nz=nz))[:, 0]
        self.setup_s = time.perf_counter() - t0
        self.plane = plane

    def step(self, k):
        view = self.o.view_matrix(0.0, (k % ORBIT_VIEWS) * (360.0 / ORBIT_VIEWS))
        _, s = self.o.render(None, self.dims, view, image=self.img, plane=self.plane)
        return s


def run_reference(args):
    """--impl reference: the reference's CPU path.  The reference has no CPU implementation and its CUDA file does not
    compile as it stands (CUDA 12.9 removed texture references), so this is the OpenMP oracle port of
    its d_render with all host threads, on the same metric, the same volume (--ref-volume, 1024^3, decoded on the
    host), the same views and the arm's own frame size.  One step = one view."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import vrdd_b200.dist as D
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    img = D.frame_size(args.image, world)
    rc = CpuRaycaster(args.seed, img, args.ref_volume)
    steps = max(1, min(args.steps, 1024))
    views = timed_views(steps)
    warm = min(args.warmup, len(views))
    for k in views[:warm]:
        rc.step(k)
    t0 = time.perf_counter()
    samples = sum(rc.step(k) for k in views)
    dt = time.perf_counter() - t0
    val = samples / dt / 1e9
    sample = (f"{steps} {img[0]}x{img[1]} views spread over the 64-view orbit on a {args.ref_volume}^3 volume decoded on the host "
              f"({rc.setup_s:.0f} s, untimed)")
    line = {"impl": "reference", "metric": "raycast_throughput", "value": val, "unit": "Gsamples/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": dt / steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "fps": steps / dt, "samples_per_frame": samples / steps,
            "config": {"workload": f"{args.ref_volume}^3 distribution volume (32 bins) decoded on the host, ray cast {img[0]}x{img[1]} "
                                   "per step over the 64-view orbit, reference constants (tstep 0.01, 500 steps, threshold 0.95, "
                                   "density 0.05, queryMethod 1); OpenMP port of the reference's d_render",
                       "volume": [args.ref_volume] * 3, "image": list(img),
                       "views": f"{steps} timed views spread over the 64-view orbit (view k*64/K)"},
            "cpu_baseline": {"value": val, "unit": "Gsamples/s", "cores": rc.o.num_threads(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": val, "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------

class Dist:
    """Rank bookkeeping + the two collectives the harness itself needs."""

    def __init__(self, torch, dist, dev):
        self.torch, self.dist, self.dev = torch, dist, dev
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_list(self, xs):
        if self.world == 1:
            return list(xs)
        t = self.torch.tensor(list(xs), dtype=self.torch.int64, device=self.dev)
        self.dist.all_reduce(t)
        return [int(v) for v in t.tolist()]

    def all_ok(self, ok):
        if self.world == 1:
            return bool(ok)
        t = self.torch.tensor([1 if ok else 0], dtype=self.torch.int32, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return int(t.item()) == 1


class PeerFrames:
    """N > 1: two frames (double buffer) + three counters in rank 0's HBM, mapped by every rank over CUDA IPC.  The
    ray-cast kernels of all ranks store their tiles straight into the current frame; the last block of each launch bumps
    done[slot]; rank 0's STREAM waits for world * generation and then bumps `consumed`, which the other ranks' streams
    wait for before they render into that slot again.  No host barrier per frame."""

    def __init__(self, V, r, ctx, nbytes):
        self.V, self.r, self.ctx = V, r, ctx
        self.nbytes = nbytes
        dist = ctx.dist
        mine = [r.frame_alloc(nbytes), r.frame_alloc(nbytes), r.frame_alloc(256)] if ctx.rank == 0 else None
        box = [[r.frame_export(p) for p in mine]] if ctx.rank == 0 else [None]
        dist.broadcast_object_list(box, src=0)
        ok, opened = True, None
        try:
            opened = mine if ctx.rank == 0 else [r.frame_open(hb) for hb in box[0]]
        except V.VrddError:
            ok = False
        self.ok = ctx.all_ok(ok)
        self.owned, self.ptrs = mine, (opened if ok else None)
        if not self.ok:
            self.close()
            return
        self.frames = opened[:2]
        self.done = [opened[2], opened[2] + 64]
        self.consumed = opened[2] + 128
        self.frame_no = 0                                         # frames rendered so far (same on every rank)

    def render(self, fw, fh, params, part):
        """One frame: every rank renders its tiles into rank 0's current frame; returns the slot."""
        r, k = self.r, self.frame_no
        s = k & 1
        if self.ctx.rank != 0 and k >= 2:
            r.stream_wait_flag(self.consumed, k - 1)              # frame k - 2 (the slot's previous content) has been consumed
        r.set_frame_signal(self.done[s])
        r.render(self.frames[s], fw, fh, params, part=part, clear_misses=True)
        r.set_frame_signal(None)
        if self.ctx.rank == 0:
            # all ranks' tiles of frame k are in -> consumed (one stream operation)
            r.stream_wait_post_flag(self.done[s], self.ctx.world * (k // 2 + 1), self.consumed)
        self.frame_no = k + 1
        return s

    def close(self):
        """Importers unmap first, then the owner frees (freeing exported memory that is still mapped elsewhere is undefined)."""
        if self.ptrs and self.ctx.rank != 0:
            for p in self.ptrs:
                self.r.frame_close(p)
        self.ctx.barrier()
        if self.ptrs and self.ctx.rank == 0:
            for p in self.ptrs:
                self.r.frame_free(p)
        self.ptrs = None


def decode_volume(V, D, r, ctx, torch, args, peers=None):
    """P1a: raw-histogram decode of my z-range, slab by slab; returns (my_ms, n_slabs, slab, z_lo, z_hi)."""
    W = H = Dz = args.volume
    slice_vox = W * H
    z_lo, z_hi = D.slab_range(Dz, ctx.rank, ctx.world)
    # N > 1: a rank's whole z-range is one slab (N = 2: 69 GB of histograms), so the replication pass below decodes from a
    # resident buffer
    slab = (z_hi - z_lo) if ctx.world > 1 else max(1, min(args.slab_z, z_hi - z_lo))
    reps = max(1, args.decode_reps)
    hist_buf = torch.empty(slab * slice_vox * 32, dtype=torch.float32, device=ctx.dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    dec_ms = 0.0
    for z0 in range(z_lo, z_hi, slab):
        nz = min(slab, z_hi - z0)
        r.synth_histograms_device(args.seed, z0, nz, hist_buf)     # untimed: the slab is resident before timing
        r.set_histograms_device(hist_buf, z0, nz)
        r.decode(V.SRC_ORIGINAL, z0, nz)                           # warm-up (also creates the arrays)
        ctx.barrier()
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(reps):
            r.decode(V.SRC_ORIGINAL, z0, nz)
        e1.record()
        torch.cuda.synchronize()
        dec_ms += e0.elapsed_time(e1) / reps
    return dec_ms, (z_hi - z_lo + slab - 1) // slab, slab, z_lo, z_hi, hist_buf


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import vrdd_b200 as V
    import vrdd_b200.dist as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU leg)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = Dist(torch, dist, dev)
    barrier, max_over_ranks = ctx.barrier, ctx.max
    peaks, peak_src = measured_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", HBM_FALLBACK_GBS))

    W = H = Dz = args.volume
    slice_vox = W * H
    total_vox = slice_vox * Dz
    r = V.Renderer(local)
    r.set_stream(torch.cuda.current_stream().cuda_stream)
    r.set_sampler(V.SAMPLER_TEXTURE if args.sampler == "texture" else V.SAMPLER_BRICKED)
    if args.tf:
        r.set_variant("raycast_tf", args.tf)
    if args.unroll:
        r.set_variant("raycast_unroll", str(args.unroll))
    r.set_variant("raycast_layout", args.layout)
    r.set_variant("decode_hist", args.decode_variant)
    r.set_variant("decode_fractal", args.fractal_variant)
    if world > 1:
        r.keep_linear_planes(True)
    r.set_volume(W, H, Dz)

    ev = lambda: torch.cuda.Event(enable_timing=True)
    clocks = ClockSampler(local)
    clocks.start()
    launches0 = r.kernel_launches()

    # ---- P1a: raw-histogram decode of my z-range, slab by slab (no replication yet) ------------------------
    my_dec_ms, n_slabs, slab, z_lo, z_hi, hist_buf = decode_volume(V, D, r, ctx, torch, args)
    dec_ms = max_over_ranks(my_dec_ms)
    dec_gbs = total_vox * HIST_BYTES_PER_VOXEL / (dec_ms * 1e-3) / 1e9
    decode = {"kernel": "decode_hist_" + args.decode_variant + "_kernel", "voxels": total_vox, "ms": dec_ms,
              "gbs": dec_gbs, "gvoxels_per_s": total_vox / (dec_ms * 1e-3) / 1e9,
              "bytes_per_voxel": HIST_BYTES_PER_VOXEL, "slab_z": slab, "launches": n_slabs * world,
              "peak_gbs": hbm_peak * world, "frac_of_hbm_peak": dec_gbs / (hbm_peak * world)}
    launch_bytes = min(slab, z_hi - z_lo) * slice_vox * HIST_BYTES_PER_VOXEL
    launch_ms = my_dec_ms * (min(slab, z_hi - z_lo) / (z_hi - z_lo))
    roofline_decode = {"kernel": decode["kernel"], "bound": "hbm", "achieved": launch_bytes / (launch_ms * 1e-3) / 1e9,
                       "peak": hbm_peak, "unit": "GB/s", "peak_source": peak_src, "launch_ms": launch_ms,
                       "bytes_per_launch": launch_bytes, "traffic": ncu_traffic(decode["kernel"], launch_bytes)}
    roofline_decode["frac"] = roofline_decode["achieved"] / roofline_decode["peak"]

    # ---- N > 1: replication fused into the decode ---------------------------------------------------------------
    # Every rank maps the other ranks' linear planes (CUDA IPC) and decodes its slab once more with them attached: the
    # decode kernel stores every value into all N copies over NVLink.  The plane the frames sample (mean) first, then
    # the other two; each rank commits what it received into its 3-D arrays.  An in-place NCCL all-gather of one plane
    # is timed next to it for comparison.
    replicate = None
    if world > 1:
        my_planes = r.get_decoded_planes_device(V.SRC_ORIGINAL)
        handles = [None] * world
        dist.all_gather_object(handles, [r.frame_export(p) for p in my_planes])
        opened, ok = {}, True
        try:
            for q in range(world):
                if q != rank:
                    opened[q] = [r.frame_open(hb) for hb in handles[q]]
        except V.VrddError:
            ok = False
        fused = ctx.all_ok(ok)
        others = [q for q in range(world) if q != rank]
        nz_mine = z_hi - z_lo
        # the slab buffer still holds my LAST slab; regenerate per slab when there are several
        def decode_with_peers(mask):
            peers = [[opened[q][i] if (mask >> i) & 1 else None for i in range(3)] for q in others]
            r.set_peer_planes(V.SRC_ORIGINAL, peers)
            for z0 in range(z_lo, z_hi, slab):
                nz = min(slab, z_hi - z0)
                if n_slabs > 1:
                    r.synth_histograms_device(args.seed, z0, nz, hist_buf)
                    r.set_histograms_device(hist_buf, z0, nz)
                r.decode(V.SRC_ORIGINAL, z0, nz)
            r.set_peer_planes(V.SRC_ORIGINAL, [])

        def commit_received(mask):
            for q in others:
                a, b = D.slab_range(Dz, q, world)
                r.commit_planes(V.SRC_ORIGINAL, a, b - a, plane_mask=mask)

        replicate = {"fused": fused}
        if fused:
            times = {}
            for name, mask in (("all_planes", 7), ("mean", 1), ("variance_entropy", 6)):
                barrier()
                e0, e1, e2 = ev(), ev(), ev()
                e0.record()
                decode_with_peers(mask)
                e1.record()
                barrier()                                          # every rank's stores have landed
                e1b = ev(); e1b.record()
                commit_received(mask)
                e2.record()
                torch.cuda.synchronize()
                times[name] = {"decode_and_peer_stores_ms": max_over_ranks(e0.elapsed_time(e1)),
                               "commit_ms": max_over_ranks(e1b.elapsed_time(e2))}
                times[name]["total_ms"] = times[name]["decode_and_peer_stores_ms"] + times[name]["commit_ms"]
            replicate.update(times)
            replicate["note"] = ("decode of my z-slab with the other ranks' planes attached (vrdd_set_peer_planes): the decode "
                                 "kernel stores into all N copies over NVLink; then the received slabs are committed to the 3-D "
                                 "arrays.  all_planes = one pass for the three planes; mean = only what the frames sample (decode + "
                                 "replicate + commit), variance_entropy = the other two planes afterwards")
            replicate["bytes_sent_per_rank"] = {"all_planes": nz_mine * slice_vox * 12 * (world - 1), "mean": nz_mine * slice_vox * 4 * (world - 1),
                                                "variance_entropy": nz_mine * slice_vox * 8 * (world - 1)}
        # comparison / fallback: in-place NCCL all-gather of the mean plane (and of all planes when IPC is not available)
        planes_t = [V.as_torch(p, (Dz * slice_vox,), device=dev) for p in my_planes]
        barrier()
        e0, e1 = ev(), ev()
        e0.record()
        for p in (planes_t if not fused else planes_t[:1]):
            D.allgather_plane(p, Dz, slice_vox, rank, world)
        if not fused:
            r.commit_planes(V.SRC_ORIGINAL, 0, Dz)
        e1.record()
        torch.cuda.synchronize()
        replicate["nccl_allgather_ms"] = max_over_ranks(e0.elapsed_time(e1))
        replicate["nccl_allgather_what"] = "mean plane, in place, no commit (comparison)" if fused else "three planes + commit"
        # queryMethod 7 interpolates the UN-normalised block means: the same fused replication for that plane
        if fused and args.mode7:
            r.enable_interpolated_mean(True)
            raw_handles = [None] * world
            dist.all_gather_object(raw_handles, r.frame_export(r.get_mean_raw_device()))
            raw_opened = {q: r.frame_open(raw_handles[q]) for q in others}
            r.set_peer_planes(V.SRC_ORIGINAL, [[None, None, None] for _ in others])
            r.set_peer_mean_raw([raw_opened[q] for q in others])
            barrier()
            e0, e1 = ev(), ev()
            e0.record()
            r.decode(V.SRC_ORIGINAL, z_lo, z_hi - z_lo)          # n_slabs == 1 at N > 1: the slab is still attached
            r.set_peer_planes(V.SRC_ORIGINAL, [])
            barrier()
            for q in others:
                a, b = D.slab_range(Dz, q, world)
                r.commit_mean_raw(a, b - a)
            e1.record()
            torch.cuda.synchronize()
            replicate["mean_raw_ms"] = max_over_ranks(e0.elapsed_time(e1))
            barrier()
            for q in raw_opened:
                r.frame_close(raw_opened[q])
        for q in opened:
            for p in opened[q]:
                r.frame_close(p)
    del hist_buf
    torch.cuda.empty_cache()

    # ---- P1b: fractal-code decode of the same volume (compact codes) --------------------------
    reps = max(1, args.decode_reps)
    if args.fractal and world == 1:
        T, max_ne = 622, 8
        nvs = slab * slice_vox
        cb = torch.empty(nvs * 4, dtype=torch.int32, device=dev)
        er = torch.empty(nvs * max_ne * 2, dtype=torch.float32, device=dev)
        off = torch.empty((nvs + V.ERR_CHUNK - 1) // V.ERR_CHUNK + 1, dtype=torch.int64, device=dev)
        tm = torch.empty(T * 32, dtype=torch.float32, device=dev)
        fr_ms, fr_bytes = 0.0, 0
        for z0 in range(0, Dz, slab):
            nz = min(slab, Dz - z0)
            tot = r.synth_fractal_device(args.seed, T, max_ne, z0, nz, cb, er, off, tm)
            r.set_fractal_device(cb, er, off, tm, T, z0, nz)
            r.decode(V.SRC_FRACTAL, z0, nz)
            torch.cuda.synchronize()
            e0, e1 = ev(), ev()
            e0.record()
            for _ in range(reps):
                r.decode(V.SRC_FRACTAL, z0, nz)
            e1.record()
            torch.cuda.synchronize()
            fr_ms += e0.elapsed_time(e1) / reps
            fr_bytes += nz * slice_vox * (16 + 12) + tot * 8      # codebook + planes + 8 B per error
        fr_kernel = {"moments": "decode_fractal_moments_smem_kernel", "moments768": "decode_fractal_moments_smem_kernel",
                     "moments2": "decode_fractal_moments2_kernel", "moments2r": "decode_fractal_moments2_kernel",
                     "moments_global": "decode_fractal_moments_kernel", "dense": "decode_fractal_dense_kernel"}
        fr_traffic = traffic_record("decode_fractal_moments2_kernel")
        decode["fractal"] = {"kernel": fr_kernel.get(args.fractal_variant, args.fractal_variant), "ms": fr_ms,
                             "gbs": fr_bytes / (fr_ms * 1e-3) / 1e9,
                             "frac_of_hbm_peak": fr_bytes / (fr_ms * 1e-3) / 1e9 / hbm_peak,
                             "gvoxels_per_s": total_vox / (fr_ms * 1e-3) / 1e9, "bytes_per_voxel": fr_bytes / total_vox,
                             "templates": T, "mean_ne": (fr_bytes / total_vox - 28) / 8}
        if fr_traffic and args.fractal_variant == "moments2":
            decode["fractal"]["dram_bytes_over_algorithmic"] = fr_traffic["dram_bytes_per_launch"] / (fr_bytes / (Dz / slab))
        del cb, er, off, tm
        torch.cuda.empty_cache()

    # ---- P2: ray casting -----------------------------------------------------------------------
    views = timed_views(args.steps)
    warm_views = [views[k % len(views)] for k in range(args.warmup)]
    distinct = sorted(set(views))
    p2p = world > 1 and args.assemble == "p2p"

    class Frames:
        """One frame geometry (weak: grows with N; strong: fixed 2048^2): buffers, partition, how a step is rendered."""

        def __init__(self, fw, fh):
            self.fw, self.fh = fw, fh
            self.part = V.TilePartition(D.TILE, D.TILE, rank, world) if world > 1 else None
            self.img = torch.zeros(fh, fw, dtype=torch.int32, device=dev)
            self.red = torch.zeros_like(self.img) if world > 1 else None
            self.peer = None
            if p2p:
                pf = PeerFrames(V, r, ctx, fw * fh * 4)
                self.peer = pf if pf.ok else None

        def render_step(self, k, params):
            r.set_view(orbit_view(V, k))
            if self.peer:
                return self.peer.render(self.fw, self.fh, params, self.part)
            r.render(self.img, self.fw, self.fh, params, part=self.part, clear_misses=True)
            if world > 1:
                D.reduce_frame(self.img, self.red, dst=0)
            return 0

        def render_local(self, k, params):                       # the ray-cast kernel alone, into a local buffer
            r.set_view(orbit_view(V, k))
            r.render(self.img, self.fw, self.fh, params, part=self.part, clear_misses=True)

        def count_samples(self, params, view_ids):
            r.count_samples(True)
            counts = []
            for k in view_ids:
                self.render_local(k, params)
                counts.append(r.get_sample_count())
            r.count_samples(False)
            return dict(zip(view_ids, ctx.sum_list(counts)))

        def time(self, params, view_ids, warm_ids, local=False):
            step = self.render_local if local else self.render_step
            for k in warm_ids:
                step(k, params)
            barrier()
            l0 = r.kernel_launches()
            e0, e1 = ev(), ev()
            e0.record()
            for k in view_ids:
                step(k, params)
            e1.record()
            barrier()
            return max_over_ranks(e0.elapsed_time(e1)), r.kernel_launches() - l0

        def check_assembly(self, params):
            """The frame the kernels assemble in rank 0's HBM must be, bit for bit, the NCCL-reduced one."""
            if not self.peer:
                return
            self.render_local(7, params)
            D.reduce_frame(self.img, self.red, dst=0)
            s = self.render_step(7, params)
            barrier()
            if rank == 0 and not torch.equal(V.as_torch(self.peer.frames[s], (self.fh, self.fw), typestr="<i4", device=dev), self.red):
                raise SystemExit("bench.py: peer-assembled frame differs from the NCCL-reduced frame")

        def device_frame(self, slot):
            if self.peer:
                return V.as_torch(self.peer.frames[slot], (self.fh, self.fw), typestr="<i4", device=dev)
            return self.red if world > 1 else self.img

        def close(self):
            if self.peer:
                barrier()
                self.peer.close()

    params = V.default_render_params(query_method=1)
    fw, fh = D.frame_size(args.image, world)
    F = Frames(fw, fh)
    F.check_assembly(params)
    counts = F.count_samples(params, distinct)                   # also builds the layered copies the views select
    samples = sum(counts[k] for k in views)
    ray_ms, ray_launches = F.time(params, views, warm_views)
    gsamples = samples / (ray_ms * 1e-3) / 1e9
    ms_per_step = ray_ms / args.steps
    kernel_ms, _ = F.time(params, views, warm_views, local=True)  # the ray-cast kernel alone, on the SAME views
    kernel_ms /= args.steps

    tex_peak = l1tex_peak() if (rank == 0 and args.l1tex) else None   # the other ranks wait at the next barrier

    # ---- strong scaling (BASELINE.json configs[2]): a fixed 2048 x 2048 frame of the same volume at every N --------
    strong = None
    if args.strong:
        S = Frames(STRONG_IMAGE, STRONG_IMAGE) if (fw, fh) != (STRONG_IMAGE, STRONG_IMAGE) else F
        s_counts = S.count_samples(params, distinct) if S is not F else counts
        s_samples = sum(s_counts[k] for k in views)
        s_ms, s_launches = (S.time(params, views, warm_views) if S is not F else (ray_ms, ray_launches))
        sk_ms, _ = (S.time(params, views, warm_views, local=True) if S is not F else (kernel_ms * args.steps, 0))
        tk = "raycast_strong_n%d" % world
        tr = traffic_record(tk)
        strong = {"workload": f"{args.volume}^3 volume, fixed {STRONG_IMAGE}x{STRONG_IMAGE} frame, image-space tiles over {world} GPU(s) "
                              "(BASELINE.json configs[2])", "scaling": "strong", "n_gpus": world,
                  "value": s_samples / (s_ms * 1e-3) / 1e9, "unit": "Gsamples/s", "ms_per_step": s_ms / args.steps,
                  "fps": args.steps / (s_ms * 1e-3), "samples_per_frame": s_samples / args.steps,
                  "kernel_ms_per_step": sk_ms / args.steps, "gpu_launches": int(s_launches),
                  "roofline": ray_roofline(s_samples / world / args.steps, sk_ms / args.steps, args.volume, STRONG_IMAGE, STRONG_IMAGE,
                                           hbm_peak, peak_src, tr["dram_bytes_per_launch"] if (tr and args.volume == 1024) else None, tex_peak)}
        if S is not F:
            S.close()
            del S

    # resolution-matched step: tstep = 2/N so the volume is sampled once per voxel
    matched = None
    if args.matched and world == 1:
        mp_ = V.default_render_params(query_method=1, tstep=2.0 / args.volume,
                                      max_steps=int(math.ceil(math.sqrt(3.0) * args.volume)) + 1)
        mviews = timed_views(min(args.steps, 16))
        mcounts = F.count_samples(mp_, sorted(set(mviews)))
        m_ms, _ = F.time(mp_, mviews, mviews[:3])
        ms_ = sum(mcounts[k] for k in mviews)
        matched = {"tstep": 2.0 / args.volume, "max_steps": mp_.max_steps, "gsamples_per_s": ms_ / (m_ms * 1e-3) / 1e9,
                   "ms_per_step": m_ms / len(mviews), "fps": len(mviews) / (m_ms * 1e-3), "samples_per_frame": ms_ / len(mviews)}

    # ---- end to end through the C ABI with host buffers --------------------------------------
    host_img = torch.empty(fh, fw, dtype=torch.int32).pin_memory()
    e2e_sync = None
    if world == 1:
        # (i) one frame at a time: render, read back, synchronise
        for k in warm_views[:3]:
            r.set_view(orbit_view(V, k)); r.render_host(host_img, fw, fh, params)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in views:
            r.set_view(orbit_view(V, k))                         # copyInvViewMatrix: 48 B host -> kernel parameters
            r.render_host(host_img, fw, fh, params)              # render + D2H of the frame + synchronize
        dt_sync = time.perf_counter() - t0
        e2e_sync = {"value": samples / dt_sync / 1e9, "fps": args.steps / dt_sync,
                    "call": "vrdd_set_view + vrdd_render_host (render, device->pinned-host frame copy, synchronize) per frame"}
        # (ii) the orbit as a sequence: every frame still goes to host memory inside the timed region, but frame
        # k's read-back (second stream) overlaps frame k+1's rendering
        host_imgs = [host_img, torch.empty(fh, fw, dtype=torch.int32).pin_memory()]
        for i, k in enumerate(warm_views[:3]):
            r.set_view(orbit_view(V, k)); r.render_host_async(host_imgs[i & 1], fw, fh, params)
        r.render_host_wait()
        t0 = time.perf_counter()
        for i, k in enumerate(views):
            r.set_view(orbit_view(V, k))
            r.render_host_async(host_imgs[i & 1], fw, fh, params)
        r.render_host_wait()                                     # all frames are in host memory
        dt = time.perf_counter() - t0
        check = torch.empty(fh, fw, dtype=torch.int32).pin_memory()   # the pipelined frame is the synchronous frame
        r.set_view(orbit_view(V, views[-1])); r.render_host(check, fw, fh, params)
        assert torch.equal(check, host_imgs[(len(views) - 1) & 1]), "pipelined read-back differs from the synchronous frame"
        call = ("vrdd_set_view + vrdd_render_host_async per frame, vrdd_render_host_wait at the end (two frames in "
                "flight: read-back of frame k overlaps rendering of frame k+1)")
    else:
        # N ranks fill ONE frame in shared host memory: full-width 64-row bands, round-robin over ranks; every rank
        # renders its bands and copies them to their place in the frame over its own PCIe link, the copy of frame k
        # overlapping the rendering of frame k+1.  A barrier per frame (behind a device-side fence on the previous
        # frame's copy) marks frame k-1 complete on all ranks.
        from multiprocessing import shared_memory
        nbytes = fw * fh * 4
        shm, host2, shared_ok = None, None, 1
        if rank == 0:
            try:
                shm = shared_memory.SharedMemory(create=True, size=2 * nbytes)
            except Exception:
                shm = None
        box_ = [shm.name if shm is not None else None]
        dist.broadcast_object_list(box_, src=0)                   # every rank takes part, whatever happened above
        try:
            if box_[0] is None:
                raise RuntimeError("no shared memory segment")
            if rank != 0:
                shm = shared_memory.SharedMemory(name=box_[0])
                try:                                              # rank 0 owns the segment: keep this process's
                    from multiprocessing import resource_tracker  # tracker from unlinking it again at exit
                    resource_tracker.unregister(shm._name, "shared_memory")
                except Exception:
                    pass
            host2 = np.ndarray((2, fh, fw), dtype=np.int32, buffer=shm.buf)
            for y0 in range(64 * rank, fh, 64 * world):          # first touch by the rank that fills the band: its pages
                host2[:, y0:y0 + 64] = 0                          # land on that rank's NUMA node
            V.host_register(host2.ctypes.data, 2 * nbytes)
        except Exception:
            shared_ok = 0
        if ctx.all_ok(shared_ok):
            bands = V.TilePartition(fw, 64, rank, world)
            barrier()
            for i, k in enumerate(warm_views[:3]):
                r.set_view(orbit_view(V, k)); r.render_host_async(host2[i & 1], fw, fh, params, part=bands)
            r.render_host_wait()
            barrier()
            t0 = time.perf_counter()
            for i, k in enumerate(views):
                r.set_view(orbit_view(V, k))
                r.render_host_async(host2[i & 1], fw, fh, params, part=bands)
                if args.e2e_sync_every and (i % args.e2e_sync_every) == 0:
                    r.render_host_fence(1)                        # stream waits for the copy of frame k-1 ...
                    dist.barrier()                                # ... then all ranks: frame k-1 is complete in host memory
            r.render_host_wait()
            barrier()
            dt = max_over_ranks(time.perf_counter() - t0)
            # the frame in shared host memory is, bit for bit, the one the device-side assembly gives
            slot = F.render_step(views[-1], params)
            torch.cuda.synchronize()
            barrier()
            if rank == 0:
                dev_frame = F.device_frame(slot).cpu().numpy()
                if not np.array_equal(dev_frame, host2[(len(views) - 1) & 1]):
                    bad = np.argwhere(dev_frame != host2[(len(views) - 1) & 1])
                    raise SystemExit("bench.py: the frame assembled in shared host memory differs from the device-assembled frame: "
                                     f"{len(bad)} pixels, rows {bad[:, 0].min()}..{bad[:, 0].max()}, columns {bad[:, 1].min()}..{bad[:, 1].max()}, "
                                     f"first {bad[:4].tolist()}, device {[hex(int(dev_frame[y, x])) for y, x in bad[:4]]} "
                                     f"host {[hex(int(host2[(len(views) - 1) & 1][y, x])) for y, x in bad[:4]]}")
            call = ("vrdd_set_view + vrdd_render_host_async(part = my 64-row bands) per frame on every rank into one frame in "
                    "shared, page-locked host memory (N PCIe links), vrdd_render_host_fence + barrier per frame, "
                    "vrdd_render_host_wait at the end")
            V.host_unregister(host2.ctypes.data)
            del host2
            shm.close()
            barrier()
            if rank == 0:
                shm.unlink()
        else:
            host2 = None
            if shm is not None:
                shm.close()
                if rank == 0:
                    shm.unlink()
            barrier()
            t0 = time.perf_counter()
            for k in views:
                slot = F.render_step(k, params)
                if rank == 0:
                    host_img.copy_(F.device_frame(slot), non_blocking=True)
                torch.cuda.synchronize()
            barrier()
            dt = max_over_ranks(time.perf_counter() - t0)
            call = ("vrdd_set_view + vrdd_render (my tiles, stored into rank 0's frame over NVLink, frame-complete counter) + "
                    "device->pinned-host frame copy") if F.peer else \
                   "vrdd_set_view + vrdd_render (my tiles) + NCCL reduce to rank 0 + device->pinned-host frame copy"
    e2e = {"value": samples / dt / 1e9, "unit": "Gsamples/s", "h2d_bytes_per_step": 48 + 32,
           "d2h_bytes_per_step": fw * fh * 4, "fps": args.steps / dt, "call": call}
    if e2e_sync:
        e2e["one_frame_at_a_time"] = e2e_sync

    # ---- decode end to end from HOST memory (upload inside the timed region) ------------------
    if world == 1 and args.e2e_decode_z > 0:
        r2 = V.Renderer(local)
        ez = args.e2e_decode_z
        r2.set_volume(W, H, ez)
        tmp = torch.empty(ez * slice_vox * 32, dtype=torch.float32, device=dev)
        r2.synth_histograms_device(args.seed, 0, ez, tmp)
        r2.synchronize()
        h_hist = torch.empty(ez * slice_vox * 32, dtype=torch.float32).pin_memory()
        h_hist.copy_(tmp)
        del tmp
        best = 1e30
        for _ in range(3):
            t0 = time.perf_counter()
            r2.set_histograms_host(h_hist)                      # initCuda: H2D of the histograms
            r2.decode(V.SRC_ORIGINAL)                           # basicDataProcessing
            r2.synchronize()
            best = min(best, time.perf_counter() - t0)
        nv = ez * slice_vox
        decode["e2e"] = {"value": nv * HIST_BYTES_PER_VOXEL / best / 1e9, "unit": "GB/s", "h2d_bytes": nv * 128,
                         "sample": f"{W}x{H}x{ez} slab from pinned host memory: vrdd_set_histograms_host + "
                                   "vrdd_decode + synchronize, best of 3"}
        r2.close()
        del h_hist

    # ---- BASELINE.json configs[1]: 512^3 volume, 1024x1024 frames, the 64-view orbit, one GPU ----------------------
    cfg1 = None
    if world == 1 and args.config1 and args.volume != 512:
        try:
            r5 = V.Renderer(local)
            r5.set_stream(torch.cuda.current_stream().cuda_stream)
            r5.set_volume(512, 512, 512)
            hb = torch.empty(512 ** 3 * 32, dtype=torch.float32, device=dev)
            r5.synth_histograms_device(args.seed, 0, 512, hb)
            r5.set_histograms_device(hb, 0, 512)
            r5.decode(V.SRC_ORIGINAL)
            torch.cuda.synchronize()
            e0, e1 = ev(), ev()
            e0.record(); r5.decode(V.SRC_ORIGINAL); e1.record()
            torch.cuda.synchronize()
            d_ms = e0.elapsed_time(e1)
            del hb
            torch.cuda.empty_cache()
            img5 = torch.zeros(1024, 1024, dtype=torch.int32, device=dev)
            p5 = V.default_render_params(query_method=1)
            r5.count_samples(True)
            c5 = []
            for k in range(ORBIT_VIEWS):
                r5.set_view(orbit_view(V, k)); r5.render(img5, 1024, 1024, p5, clear_misses=True); c5.append(r5.get_sample_count())
            r5.count_samples(False)
            torch.cuda.synchronize()
            e0, e1 = ev(), ev()
            e0.record()
            for k in range(ORBIT_VIEWS):
                r5.set_view(orbit_view(V, k)); r5.render(img5, 1024, 1024, p5, clear_misses=True)
            e1.record()
            torch.cuda.synchronize()
            o_ms = e0.elapsed_time(e1)
            cfg1 = {"workload": "512^3 volume decoded once, 1024x1024 frames, 64-view orbit, reference constants",
                    "decode_ms": d_ms, "decode_frac_of_hbm_peak": 512 ** 3 * HIST_BYTES_PER_VOXEL / (d_ms * 1e-3) / 1e9 / hbm_peak,
                    "gsamples_per_s": sum(c5) / (o_ms * 1e-3) / 1e9, "fps": ORBIT_VIEWS / (o_ms * 1e-3),
                    "ms_per_view": o_ms / ORBIT_VIEWS, "samples_per_frame": sum(c5) / ORBIT_VIEWS}
            r5.close()
            del img5
        except Exception as exc:
            cfg1 = {"error": repr(exc)}

    # ---- queryMethod 7 (interpolated block means, point-sampled cells) on the same volume and views -------
    mode7 = None
    if args.mode7 and (world == 1 or (replicate and "mean_raw_ms" in replicate)):
        try:
            if world == 1:
                r.enable_interpolated_mean(True)                  # the block-mean plane is only kept on request
                hb = torch.empty(slab * slice_vox * 32, dtype=torch.float32, device=dev)
                for z0 in range(0, Dz, slab):
                    nz = min(slab, Dz - z0)
                    r.synth_histograms_device(args.seed, z0, nz, hb)
                    r.set_histograms_device(hb, z0, nz)
                    r.decode(V.SRC_ORIGINAL, z0, nz)
                torch.cuda.synchronize()
                del hb
                torch.cuda.empty_cache()
            p7 = V.default_render_params(query_method=7)
            v7 = timed_views(min(args.steps, 32))
            c7 = F.count_samples(p7, sorted(set(v7)))
            ms7, _ = F.time(p7, v7, v7[:3])
            s7 = sum(c7[k] for k in v7)
            mode7 = {"gsamples_per_s": s7 / (ms7 * 1e-3) / 1e9, "ms_per_step": ms7 / len(v7), "fps": len(v7) / (ms7 * 1e-3),
                     "steps": len(v7), "kernel": "raycast_mode7_kernel<gather>",
                     "note": "the 8 cell corners of the un-normalised block means come from two tld4 gathers per sample "
                             "(layered 2-D array; 8 point fetches where the extents do not allow it), the reference's cell "
                             "cache and degenerate-cell artefact kept (volumeRender_kernel.cu:395-480); the reference "
                             "reports < 5 fps for this mode on its 50x50x10 volume (ver1.9.6.txt:168)"}
        except Exception as exc:                                   # never let the side measurement void the headline
            mode7 = {"error": repr(exc)}

    # ---- the flexible-block chain on the reference's own configuration (64^3 raw volume, block size 6) ----
    flex = None
    if world == 1 and args.flex:
        try:
            from vrdd_b200 import synth_flex                    # synthetic span store (numpy; the span tables do not ship)
            tables = synth_flex.make_tables(5, 64, n_templates=469, block=6)
            r.flex_set_tables_host(tables)
            r.flex_process(6)
            torch.cuda.synchronize()
            e0, e1 = ev(), ev()
            e0.record()
            for _ in range(20):
                r.flex_process(6, want_missing=False)
            e1.record()
            torch.cuda.synchronize()
            fms = e0.elapsed_time(e1) / 20
            ref_ms = 0.678 + 194764.078 + 2.869                 # d_queryBlockNew + d_querySpanNew + d_computeBlock, ver1.9.6.txt:8-10
            flex = {"workload": "dataProcessing(): 64^3 raw volume, block size 6 -> 11^3 blocks, 10 648 corners; synthetic "
                                "lossless span store (20 851 fractal + 6 149 simple spans of the 262 144 the reference loads)",
                    "ms": fms, "kernels": "flex_corner_kernel + flex_block_kernel",
                    "reference_ms": ref_ms, "reference_hardware": "Quadro K5000, CUDA 5.0 (ver1.9.6.txt:6-10; BASELINE.md)",
                    "vs_reference": ref_ms / fms,
                    "note": "the reference scans its 131 072-entry span tables linearly per thread; here they are hash-indexed"}
        except Exception as exc:                                # the chain is not on the north-star path: never fail the bench
            flex = {"error": repr(exc)}

    clk = clocks.stop()
    total_launches = r.kernel_launches() - launches0

    # ---- CPU baselines (rank 0, N = 1) ---------------------------------------------------------
    cpu_ray = None
    if rank == 0 and world == 1 and not args.no_cpu:
        decode["cpu_baseline"] = cpu_decode_baseline(args.seed)
        rc = CpuRaycaster(args.seed, (args.image, args.image), args.ref_volume)
        cviews = timed_views(min(args.steps, 32))
        rc.step(cviews[0])
        t0 = time.perf_counter()
        s_cpu = sum(rc.step(k) for k in cviews)
        dt_c = time.perf_counter() - t0
        cpu_ray = {"value": s_cpu / dt_c / 1e9, "unit": "Gsamples/s", "cores": rc.o.num_threads(), "kind": "port",
                   "sample": f"{len(cviews)} {args.image}x{args.image} views spread over the orbit on a {args.ref_volume}^3 volume decoded "
                             f"on the host ({rc.setup_s:.0f} s, untimed; same constants), OpenMP oracle -O3 -march=x86-64-v3",
                   "fps": len(cviews) / dt_c}
        del rc

    # ---- N > 1: sort-last bricks (configs[4]) as a sub-record of the same line --------------------------------------
    F.close()
    r.close()
    sortlast = None
    if world > 1 and args.sortlast:
        torch.cuda.empty_cache()
        try:
            sortlast = sortlast_record(args, ctx, V, D, torch, dist, local, dev, hbm_peak, steps=min(args.steps, 16))
        except Exception as exc:
            sortlast = {"error": repr(exc)}

    if rank == 0:
        my_samples = samples / world / args.steps
        # the dominant kernel of the timed step is the ray caster; at 1024^3 with the reference's fixed step ncu shows it
        # DRAM-bound (profiles/README.md).  achieved = 32 algorithmic bytes per sample / launch time; what
        # DRAM really moves is whole 128-byte lines (traffic, traffic_frac): tools/probe_atom*.cu.
        tkey = "raycast_kernel" if world == 1 else "raycast_tiles_n%d" % world
        tj = traffic_record(tkey)
        ray_traffic, traffic_what = None, "no ncu capture for this configuration"
        if tj and args.volume == 1024 and args.image == 1024:
            ray_traffic = tj["dram_bytes_per_launch"]
            traffic_what = "mean DRAM bytes of the %d orbit launches under ncu (%s)" % (tj.get("launches", 1), tj.get("source", "profiles/"))
        roofline = ray_roofline(my_samples, kernel_ms, args.volume, fw, fh, hbm_peak, peak_src, ray_traffic, tex_peak)
        roofline["kernel"] = "raycast_kernel / raycast_gather_kernel (per view)"
        roofline["note"] = ("algorithmic bytes = 32 B per trilinear sample (8 fp32 texels) x samples of the average launch; "
                            "traffic = " + traffic_what + "; DRAM delivers whole 128-byte lines (8x4x1 texels), so on oblique views "
                            "every line of the region the rays cross is read once: 64 B per sample is compulsory there, 24 B on "
                            "views along an axis")
        line = {"metric": "raycast_throughput", "value": gsamples, "unit": "Gsamples/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "fps": 1e3 / ms_per_step, "samples_per_frame": samples / args.steps,
                "config": {"workload": f"{args.volume}^3 distribution volume (32 bins) decoded on device, ray cast "
                                       f"{fw}x{fh} per step over the 64-view orbit, reference constants "
                                       "(tstep 0.01, 500 steps, threshold 0.95, density 0.05, queryMethod 1)",
                           "volume": [W, H, Dz], "image": [fw, fh], "sampler": args.sampler, "layout": args.layout,
                           "views": f"{args.steps} timed views spread over the 64-view orbit (view k*64/K)",
                           "partition": "single GPU" if world == 1 else
                           f"64x64 image tiles round-robin over {world} ranks; z-slab decode with the slabs stored into every rank's "
                           "planes by the decode kernel (NVLink); " + ("tiles stored by the ray-cast kernels into rank 0's frame over NVLink "
                                                                      "(CUDA IPC), frame-complete counter instead of a barrier" if F.peer else
                                                                      "NCCL reduce of frames to rank 0"),
                           "l2": f"inputs larger than L2 ({total_vox * 4 / 1e9:.1f} GB sampled plane, "
                                 f"{launch_bytes / 1e9:.1f} GB per decode launch); no flush"},
                "parity": {"bar": "+-1 LSB per RGBA8 channel (queryMethod 1-7), fp32 decode rtol 2e-5",
                           "pinned_by": "the reference's own device code run on a B200 (tests/golden/ref_gpu_v1.npz, ref_gpu_flex_v1.npz)",
                           "fractal_decode": "matches the reference BINARY once the stores its build drops are modelled "
                                             "(fractalDecoding returns a pointer to a local array, volumeRender_kernel.cu:196-221); the "
                                             "kernels keep the source's intent, all 32 bins"},
                "decode": decode, "roofline": roofline, "roofline_decode": roofline_decode,
                "e2e": e2e, "gpu_launches": int(ray_launches), "gpu_launches_total": int(total_launches),
                "kernel_ms_per_step": kernel_ms, "clocks": clk}
        if "fractal" in decode:
            line["decode_fractal_frac_of_hbm_peak"] = decode["fractal"]["frac_of_hbm_peak"]
        if strong:
            line["strong"] = strong
        if matched:
            line["raycast_resolution_matched"] = matched
        if cfg1:
            line["config_512"] = cfg1
        if mode7:
            line["query_method_7"] = mode7
        if flex:
            line["flex_chain"] = flex
        if replicate is not None:
            line["replicate"] = replicate
        if sortlast:
            line["sortlast"] = sortlast
        if cpu_ray:
            line["cpu_baseline"] = cpu_ray
        ref_gpu = traffic_record("reference_gpu_kernel")
        if ref_gpu:
            line["reference_gpu_kernel"] = ref_gpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def sortlast_record(args, ctx, V, D, torch, dist, local, dev, hbm_peak, steps, warmup=3):
    """configs[4]: every rank owns one VOL^3 brick (plus a one-voxel ghost layer) of a volume too large for one GPU —
    8 ranks: 2x2x2 bricks of a (2*VOL)^3 volume, 1.1 TB of histograms at VOL = 1024.  Per step: alpha pre-pass -> NCCL
    all-gather of the segment alphas -> incoming alpha -> colour pass -> NCCL SUM reduction of the float4 increments ->
    pack on rank 0.  Returns the record (rank 0) or None."""
    world, rank = ctx.world, ctx.rank
    grid = D.brick_grid(world)
    q = D.brick_of_rank(rank, grid)
    E = args.volume
    gdims = (E * grid[0], E * grid[1], E * grid[2])
    origin, size, lo, hi = D.brick_geometry(gdims, grid, q)
    r = V.Renderer(local)
    r.set_stream(torch.cuda.current_stream().cuda_stream)
    r.set_sampler({"texture": V.SAMPLER_TEXTURE, "linear": V.SAMPLER_LINEAR, "bricked": V.SAMPLER_BRICKED}[args.sortlast_layout])
    r.set_volume(*size)
    r.set_variant("sortlast_fuse", args.sortlast_fuse)
    if args.sortlast_blocks >= 0:
        r.set_variant("sortlast_blocks_per_sm", str(args.sortlast_blocks))
    ev = lambda: torch.cuda.Event(enable_timing=True)
    barrier, max_over_ranks = ctx.barrier, ctx.max
    # ---- decode my brick slab by slab -----------------------------------------------------------
    slice_vox = size[0] * size[1]
    slab = max(1, min(args.slab_z, size[2]))
    buf = torch.empty(slab * slice_vox * 32, dtype=torch.float32, device=dev)
    dec_ms = 0.0
    for z0 in range(0, size[2], slab):
        nz = min(slab, size[2] - z0)
        r.synth_histograms_region_device(args.seed, gdims, origin, z0, nz, buf)
        r.set_histograms_device(buf, z0, nz)
        r.decode(V.SRC_ORIGINAL, z0, nz)
        barrier()
        e0, e1 = ev(), ev()
        e0.record(); r.decode(V.SRC_ORIGINAL, z0, nz); e1.record()
        torch.cuda.synchronize()
        dec_ms += e0.elapsed_time(e1)
    del buf
    torch.cuda.empty_cache()
    dec_ms = max_over_ranks(dec_ms)
    brick_vox = size[0] * size[1] * size[2]
    tv = torch.tensor([brick_vox], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tv)
    decoded_vox = float(tv.item())

    # ---- render ------------------------------------------------------------------------------------
    fw = fh = args.image_sortlast
    params = V.default_render_params(query_method=1)
    br = V.Brick(gdims[0], gdims[1], gdims[2], origin[0], origin[1], origin[2], (V.C.c_float * 3)(*lo), (V.C.c_float * 3)(*hi))
    seg = torch.zeros(fh, fw, dtype=torch.float32, device=dev)
    seg_all = torch.zeros(world, fh, fw, dtype=torch.float32, device=dev)
    a_in = torch.zeros(fh, fw, dtype=torch.float32, device=dev)
    part = torch.zeros(fh, fw, 4, dtype=torch.float32, device=dev)
    frame = torch.zeros(fh, fw, dtype=torch.int32, device=dev)
    windows = bool(args.sortlast_windows)
    coll_bytes = [0, 0]                                   # all-gather (received per rank), reduce (sent per rank), per frame

    # ---- direct-send exchange: tables in every rank's HBM mapped over CUDA IPC, counters instead of collectives ----
    # Per frame parity: every rank owns a table of segment alphas float[world][rows][W] + a counter; rank 0 also a table of
    # increments float4[world][rows][W] + a counter.  Pass 1 stores its window rows into slot [rank] of EVERY table, pass 2
    # into slot [rank] of rank 0's; the last block of a launch bumps the counters (red.release.sys); readers wait in-stream.
    # Reuse of a parity's tables at frame k + 2 is safe without a further signal: a rank has waited for every rank's pass 1
    # of frame k + 1, which those ranks' streams issue after they finished reading the tables of frame k.
    direct = None
    banded = args.sortlast_compose == "bands" or (args.sortlast_compose == "auto" and world >= 4)
    band_rows = (fh + world - 1) // world
    if world > 1 and windows and args.sortlast_exchange == "direct":
        seg_bytes = world * fh * fw * 4
        # [0, 1] segment-alpha tables, [2] counters (64 B apart: segment 0/1, root slots 0/1, band 0/1, frame 0/1), then either
        # the band tables [3, 4] of every rank (+ the root's two frames [5, 6]) or the root's slot tables [3, 4]
        mine = [r.frame_alloc(seg_bytes), r.frame_alloc(seg_bytes), r.frame_alloc(512)]
        if banded:
            mine += [r.frame_alloc(world * band_rows * fw * 16), r.frame_alloc(world * band_rows * fw * 16)]
            if rank == 0:
                mine += [r.frame_alloc(fh * fw * 4), r.frame_alloc(fh * fw * 4)]
        elif rank == 0:
            mine += [r.frame_alloc(world * fh * fw * 16), r.frame_alloc(world * fh * fw * 16)]
        exported = [None] * world
        dist.all_gather_object(exported, [r.frame_export(p) for p in mine])
        ok, opened = True, {}
        try:
            for qq in range(world):
                if qq != rank:
                    opened[qq] = [r.frame_open(hb) for hb in exported[qq]]
        except V.VrddError:
            ok = False
        if ctx.all_ok(ok):
            ptrs = {qq: (mine if qq == rank else opened[qq]) for qq in range(world)}
            direct = {"mine": mine, "opened": opened, "frame_no": 0,
                      "seg": [[ptrs[qq][s] for qq in range(world)] for s in (0, 1)],
                      "seg_flag": [[ptrs[qq][2] + 64 * s for qq in range(world)] for s in (0, 1)]}
            if banded:
                direct["band"] = [[ptrs[qq][3 + s] for qq in range(world)] for s in (0, 1)]
                direct["band_flag"] = [[ptrs[qq][2] + 256 + 64 * s for qq in range(world)] for s in (0, 1)]
                direct["frame"] = [ptrs[0][5 + s] for s in (0, 1)]
                direct["frame_flag"] = [ptrs[0][2] + 384 + 64 * s for s in (0, 1)]
            else:
                direct["slots"] = [ptrs[0][3 + s] for s in (0, 1)]
                direct["slot_flag"] = [ptrs[0][2] + 128 + 64 * s for s in (0, 1)]
        else:
            for qq in opened:
                for pp in opened[qq]:
                    r.frame_close(pp)
            barrier()
            for pp in mine:
                r.frame_free(pp)

    main_s = torch.cuda.current_stream()
    side_s = torch.cuda.Stream(device=dev) if direct else None
    if direct:
        direct["packed"] = [None, None]
        direct["last_packed"] = None
        direct["last_frame"] = frame

    def step_direct(k, frame):
        view = orbit_view(V, k)
        r.set_view(view)
        n = direct["frame_no"]; s = n & 1; gen = n // 2 + 1
        row0, rows, (u0, u1) = D.brick_row_windows(view, grid, fh)
        if direct["packed"][s] is not None:
            main_s.wait_event(direct["packed"][s])             # my sum of frame n - 2 is done: its slots may be overwritten (see above)
        r.render_brick_alpha_send(direct["seg"][s], direct["seg_flag"][s], rank, row0[rank], rows, fw, fh, params, br)
        r.stream_wait_flag(direct["seg_flag"][s][rank], world * gen)
        r.compose_alpha_in_rows(direct["seg"][s][rank], grid, q, row0, rows, a_in, fw, fh)
        if banded:
            r.render_brick_color_send_bands(a_in, direct["band"][s], direct["band_flag"][s], band_rows, rank, row0[rank], rows, fw, fh, params, br)
        else:
            r.render_brick_color_send(a_in, direct["slots"][s], direct["slot_flag"][s], rank, row0[rank], rows, fw, fh, params, br)
        if banded or rank == 0:
            # the wait for everybody's increments and the sum + pack run on a second stream: this rank's pass 1 of the next
            # frame (which every other rank waits for) is not held up by the slowest pass 2 of this one
            after_p2 = torch.cuda.Event(); after_p2.record(main_s)
            side_s.wait_event(after_p2)
            r.set_stream(side_s.cuda_stream)
            if banded:
                r.stream_wait_flag(direct["band_flag"][s][rank], world * gen)
                dst = direct["host_frames"][s] if direct.get("host_frames") else direct["frame"][s]
                r.pack_band_slots(direct["band"][s][rank], world, row0, rows, rank, band_rows, dst, direct["frame_flag"][s],
                                  fw, fh, params.brightness)
                if rank == 0:
                    r.stream_wait_flag(direct["frame_flag"][s], world * gen)     # every band is in: the frame is complete
                    if not direct.get("host_frames"):
                        direct["last_frame"] = V.as_torch(direct["frame"][s], (fh, fw), typestr="<i4", device=dev)
            else:
                r.stream_wait_flag(direct["slot_flag"][s], world * gen)
                r.pack_frame_slots(direct["slots"][s], world, row0, rows, frame, fw, fh, params.brightness)
                direct["last_frame"] = frame
            r.set_stream(main_s.cuda_stream)
            done = torch.cuda.Event(); done.record(side_s)
            direct["packed"][s] = direct["last_packed"] = done
        direct["frame_no"] = n + 1
        coll_bytes[0] = (world - 1) * rows * fw * 4            # peer stores sent per rank, pass 1
        coll_bytes[1] = rows * fw * 16                         # peer stores sent per rank, pass 2

    def step(k, frame=frame, collectives=False):
        if direct and not collectives:
            return step_direct(k, frame)
        view = orbit_view(V, k)
        r.set_view(view)
        r.render_brick_alpha(seg, fw, fh, params, br)
        if windows:
            # Only the image rows a brick's screen footprint can touch travel: every rank derives the same
            # table of row windows from the view matrix (dist.brick_row_windows), contributes the rows of its
            # own window (a contiguous slab of `seg`), and the composition reads brick b at row y - row0[b].
            row0, rows, (u0, u1) = D.brick_row_windows(view, grid, fh)
            seg_rows = D.gather_row_windows(seg, seg_all, row0, rows, rank, world)
            r.compose_alpha_in_rows(seg_rows, grid, q, row0, rows, a_in, fw, fh)
            coll_bytes[0] = world * rows * fw * 4
        else:
            if world > 1:
                dist.all_gather_into_tensor(seg_all, seg)
            else:
                seg_all[0].copy_(seg)
            r.compose_alpha_in(seg_all, grid, q, a_in, fw, fh)
            u0, u1 = 0, fh
            coll_bytes[0] = world * fh * fw * 4
        r.render_brick_color(a_in, part, fw, fh, params, br)
        D.reduce_union_rows(part, (u0, u1), dst=0)             # no brick has samples outside the union rows
        coll_bytes[1] = (u1 - u0) * fw * 16
        if rank == 0:
            r.pack_frame(part, frame, fw, fh, params.brightness)

    views = timed_views(steps)
    distinct = sorted(set(views))
    exchange_check = None
    if direct:
        # the frame of the direct-send exchange against the frame of the NCCL exchange: the same increments, summed in brick
        # order instead of NCCL's order (fp32 addition is not associative: a last-bit difference may flip a byte by 1)
        other = torch.zeros_like(frame)
        worst, differ = 0, 0
        for k in (7, 23):
            step(k, frame)
            step(k, other, collectives=True)
            torch.cuda.synchronize()
            if rank == 0:
                a = direct["last_frame"].view(torch.uint8).to(torch.int16); b = other.view(torch.uint8).to(torch.int16)
                worst = max(worst, int((a - b).abs().max())); differ += int((a != b).sum())
        barrier()
        if rank == 0:
            if worst > 1:
                raise SystemExit(f"bench.py: direct-send sort-last frame differs from the NCCL form by {worst} LSB")
            exchange_check = {"views": [7, 23], "max_lsb": worst, "bytes_differing": differ, "bytes": 2 * fw * fh * 4}
        del other
    r.count_samples(True)
    cnt = []
    for k in distinct:
        step(k)
        cnt.append(r.get_sample_count())
    r.count_samples(False)
    counts = dict(zip(distinct, ctx.sum_list(cnt)))
    for k in views[:warmup]:
        step(k)
    barrier()
    l0 = r.kernel_launches()
    e0, e1 = ev(), ev()
    e0.record()
    for k in views:
        step(k)
    if direct and direct["last_packed"] is not None:
        main_s.wait_event(direct["last_packed"])               # the last frame is packed inside the timed region
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = r.kernel_launches() - l0
    samples = sum(counts[k] for k in views)
    # end to end: the packed frame of step k is copied to pinned host memory by a second stream while step k+1 runs
    # (two device frames, two host frames; step k+2 waits for the copy of frame k before it packs into that buffer)
    # Band owners: no copy at all — the frame lives in POSIX shared memory page-locked and mapped by every rank, and every
    # owner's pack kernel stores its RGBA8 rows straight into it over its own PCIe link (4 B per pixel, whole lines);
    # rank 0's stream learns from the frame counter that all bands have landed.
    e2e_zero_copy = False
    shm = None
    if direct and banded:
        from multiprocessing import shared_memory
        import numpy as np
        nbytes = fw * fh * 4
        shared_ok, host2 = 1, None
        if rank == 0:
            try:
                shm = shared_memory.SharedMemory(create=True, size=2 * nbytes)
            except Exception:
                shm = None
        box_ = [shm.name if shm is not None else None]
        dist.broadcast_object_list(box_, src=0)
        try:
            if box_[0] is None:
                raise RuntimeError("no shared memory segment")
            if rank != 0:
                shm = shared_memory.SharedMemory(name=box_[0])
                try:
                    from multiprocessing import resource_tracker
                    resource_tracker.unregister(shm._name, "shared_memory")
                except Exception:
                    pass
            host2 = np.ndarray((2, fh, fw), dtype=np.int32, buffer=shm.buf)
            y0, y1 = min(rank * band_rows, fh), min((rank + 1) * band_rows, fh)
            host2[:, y0:y1] = 0                                   # first touch by the band's owner
            V.host_register(host2.ctypes.data, 2 * nbytes)
        except Exception:
            shared_ok = 0
        if ctx.all_ok(shared_ok):
            e2e_zero_copy = True
            direct["host_frames"] = [host2[0].ctypes.data, host2[1].ctypes.data]
            barrier()
            for k in views[:2]:
                step(k)
            torch.cuda.synchronize()
            barrier()
            t0 = time.perf_counter()
            for k in views:
                step(k)
            torch.cuda.synchronize()                              # rank 0: the last frame counter has been reached
            barrier()
            dt = max_over_ranks(time.perf_counter() - t0)
            # the frame in host memory is the frame the device-side assembly gives
            last_parity = (direct["frame_no"] - 1) & 1
            direct["host_frames"] = None
            step(views[-1])
            torch.cuda.synchronize()
            barrier()
            if rank == 0 and not np.array_equal(direct["last_frame"].cpu().numpy(), host2[last_parity]):
                raise SystemExit("bench.py: sort-last frame in shared host memory differs from the device-assembled frame")
            barrier()
            V.host_unregister(host2.ctypes.data)
        host2 = None
        if shm is not None:
            try:
                shm.close()
                if rank == 0:
                    shm.unlink()
            except Exception:
                pass
    host = [torch.empty(fh, fw, dtype=torch.int32).pin_memory() for _ in range(2)]
    frames2 = [frame, torch.zeros_like(frame)]
    copy_stream = torch.cuda.Stream(device=dev)
    copied = [None, None]
    main = torch.cuda.current_stream()
    barrier()
    if not e2e_zero_copy:
        t0 = time.perf_counter()
    for i, k in enumerate([] if e2e_zero_copy else views):
        if copied[i & 1] is not None:
            # root slots: only the pack (second stream) writes the frame; band owners: the other ranks' packs write it, and
            # they follow this rank's pass 2 of the same frame
            (side_s if (direct and not banded) else main).wait_event(copied[i & 1])
        step(k, frames2[i & 1])
        if rank == 0:
            if direct:
                packed = direct["last_packed"]
            else:
                packed = torch.cuda.Event(); packed.record(main)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(packed)
                host[i & 1].copy_(direct["last_frame"] if direct else frames2[i & 1], non_blocking=True)
                copied[i & 1] = torch.cuda.Event(); copied[i & 1].record(copy_stream)
    torch.cuda.synchronize()
    barrier()
    if not e2e_zero_copy:
        dt = max_over_ranks(time.perf_counter() - t0)
    if direct:
        for qq in direct["opened"]:
            for pp in direct["opened"][qq]:
                r.frame_close(pp)
        barrier()
        for pp in direct["mine"]:
            r.frame_free(pp)
    r.close()
    if rank != 0:
        return None
    how = ("alpha pre-pass storing each brick's window rows into every rank's table over NVLink, in-stream counter wait, colour pass "
           + ("storing the float4 increments into the table of the owner of each band of image rows, every owner sums its band in brick "
              "order, packs and stores the RGBA8 rows into rank 0's frame" if banded else
              "storing the float4 increments into rank 0's table, sum in brick order + pack on rank 0")
           + " (no collective, no host synchronisation in the loop)") if direct else (
           "alpha pre-pass, NCCL all-gather of segment alphas, colour pass, NCCL SUM reduction of float4 increments, pack on rank 0"
           + ("; collectives restricted to the image rows of each brick's screen footprint" if windows else ""))
    dec_gbs = decoded_vox * HIST_BYTES_PER_VOXEL / (dec_ms * 1e-3) / 1e9
    return {"workload": f"sort-last: {gdims[0]}x{gdims[1]}x{gdims[2]} distribution volume in {grid[0]}x{grid[1]}x{grid[2]} "
                        f"bricks of {E}^3 (+1 ghost) over {world} GPUs, {fw}x{fh} frames of the orbit, reference constants; " + how,
            "exchange": "direct" if direct else "nccl", "exchange_check": exchange_check,
            "fused_first_segment": bool(direct) and args.sortlast_fuse == "on", "compose": ("bands" if banded else "root") if direct else "nccl",
            "n_gpus": world, "value": samples / (ms * 1e-3) / 1e9, "unit": "Gsamples/s", "steps": len(views), "ms_per_step": ms / len(views),
            "fps": len(views) / (ms * 1e-3), "samples_per_frame": samples / len(views),
            "collective_bytes_per_frame": {("alpha_sent_per_rank" if direct else "all_gather"): coll_bytes[0],
                                           ("increments_sent_per_rank" if direct else "reduce"): coll_bytes[1],
                                           "full_frame": [world * fh * fw * 4, fh * fw * 16]},
            "volume": list(gdims), "image": [fw, fh], "histogram_bytes_total": decoded_vox * 128,
            "decode": {"kernel": "decode_hist_tma_kernel", "voxels": decoded_vox, "ms": dec_ms, "gbs": dec_gbs,
                       "frac_of_hbm_peak": dec_gbs / (hbm_peak * world), "peak_gbs": hbm_peak * world},
            "e2e": {"value": samples / dt / 1e9, "unit": "Gsamples/s", "h2d_bytes_per_step": 80, "d2h_bytes_per_step": fw * fh * 4,
                    "fps": len(views) / dt,
                    "zero_copy": e2e_zero_copy,
                    "call": ("vrdd_render_brick_alpha_send/compose/color_send_bands + vrdd_pack_band_slots: every band owner's pack kernel stores "
                             "its RGBA8 rows straight into the frame in page-locked shared host memory (checked against the device-assembled "
                             "frame); otherwise: " if e2e_zero_copy else "") +
                            ("vrdd_render_brick_alpha_send/compose/color_send + vrdd_pack_frame_slots" if direct else
                             "vrdd_render_brick_alpha/compose/color + NCCL + vrdd_pack_frame") + "; the frame of step k is copied to pinned host memory "
                            "by a second stream while step k+1 renders"},
            "gpu_launches": int(launches)}


def run_sortlast(args):
    """--workload sortlast: the sort-last record as the top-level line."""
    import torch
    import torch.distributed as dist
    import vrdd_b200 as V
    import vrdd_b200.dist as D
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = Dist(torch, dist, dev)
    peaks, _ = measured_peaks()
    clocks = ClockSampler(local)
    clocks.start()
    rec = sortlast_record(args, ctx, V, D, torch, dist, local, dev, float(peaks.get("hbm_gbs", HBM_FALLBACK_GBS)), steps=args.steps,
                          warmup=args.warmup)
    clk = clocks.stop()
    if ctx.rank == 0:
        line = {"metric": "raycast_throughput", "value": rec["value"], "unit": "Gsamples/s", "n_gpus": world, "steps": rec["steps"],
                "warmup": args.warmup, "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "fps": rec["fps"], "samples_per_frame": rec["samples_per_frame"],
                "config": {"workload": rec["workload"], "collective_bytes_per_frame": rec["collective_bytes_per_frame"],
                           "volume": rec["volume"], "image": rec["image"], "histogram_bytes_total": rec["histogram_bytes_total"],
                           "l2": "inputs larger than L2; no flush"},
                "decode": rec["decode"], "e2e": rec["e2e"], "gpu_launches": rec["gpu_launches"], "clocks": clk}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--volume", type=int, default=1024, help="distribution volume edge (voxels)")
    ap.add_argument("--image", type=int, default=1024, help="frame edge per GPU (pixels)")
    ap.add_argument("--ref-volume", type=int, default=1024,
                    help="volume edge of the CPU legs, decoded on the host (1024^3: ~35 s of untimed set-up on 16 cores)")
    ap.add_argument("--slab-z", type=int, default=256, help="z-slices decoded per launch (256 -> 34 GB of histograms)")
    ap.add_argument("--decode-reps", type=int, default=3)
    ap.add_argument("--decode-variant", default="tma", choices=["tma", "ldg"])
    ap.add_argument("--fractal-variant", default="moments2",
                    choices=["moments2", "moments2r", "moments", "moments768", "moments_global", "dense"])
    ap.add_argument("--sampler", default="texture", choices=["texture", "bricked"])
    ap.add_argument("--layout", default="auto", choices=["auto", "array", "layers_x", "layers_y"])
    ap.add_argument("--tf", default=None, choices=[None, "texture", "smem"])
    ap.add_argument("--unroll", type=int, default=0, choices=[0, 1, 2, 4, 8])
    ap.add_argument("--fractal", type=int, default=1)
    ap.add_argument("--matched", type=int, default=1)
    ap.add_argument("--strong", type=int, default=1, help="also time the fixed 2048x2048 frame (BASELINE.json configs[2])")
    ap.add_argument("--sortlast", type=int, default=1, help="N > 1: also emit the sort-last record (BASELINE.json configs[4])")
    ap.add_argument("--l1tex", type=int, default=1, help="N = 1: run tools/l1tex_peak for the L1TEX roofline")
    ap.add_argument("--e2e-sync-every", type=int, default=1,
                    help="N > 1 end-to-end path: fence + barrier between ranks every K frames (0: only at the end)")
    ap.add_argument("--config1", type=int, default=1, help="also time BASELINE.json configs[1] (512^3 volume, 64-view orbit)")
    ap.add_argument("--mode7", type=int, default=1, help="also time queryMethod 7 on the same volume and views")
    ap.add_argument("--flex", type=int, default=1, help="also time the flexible-block chain (64^3, block 6)")
    ap.add_argument("--e2e-decode-z", type=int, default=8, help="z-slices of the host-memory decode leg (0 = skip)")
    ap.add_argument("--sortlast-windows", type=int, default=1,
                    help="sort-last: restrict the collectives to each brick's screen rows (0: whole frames)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--workload", default="tiles", choices=["tiles", "sortlast"],
                    help="tiles: volume replicated, image-space tiles (default); sortlast: one VOL^3 brick per GPU")
    ap.add_argument("--image-sortlast", type=int, default=2048)
    ap.add_argument("--assemble", default="p2p", choices=["p2p", "reduce"],
                    help="N > 1 tiles: p2p = kernels store tiles into rank 0's frame over NVLink; reduce = NCCL reduce")
    ap.add_argument("--sortlast-layout", default="texture", choices=["texture", "bricked", "linear"])
    ap.add_argument("--sortlast-compose", default="auto", choices=["auto", "bands", "root"],
                    help="direct exchange: increments summed + packed by band owners (every rank 1/N of the rows) or all by rank 0; "
                         "auto = bands from 4 ranks on (N = 2: root 1581 fps, bands 1505; N = 8: root 1508, bands 1792)")
    ap.add_argument("--sortlast-blocks", type=int, default=-1, help="sort-last brick kernel: resident blocks per SM (-1: library default)")
    ap.add_argument("--sortlast-fuse", default="on", choices=["on", "off"],
                    help="direct exchange: pass 1 keeps the colour of the march from alpha 0, pass 2 marches only pixels with incoming alpha")
    ap.add_argument("--sortlast-exchange", default="direct", choices=["direct", "nccl"],
                    help="direct: the kernels store into peer tables over NVLink (CUDA IPC) and signal with counters; nccl: collectives")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        if args.workload == "sortlast":
            run_sortlast(args)
        else:
            run_ours(args)


if __name__ == "__main__":
    main()
