#!/usr/bin/env python
"""bench.py — headline benchmark of the two hot paths on B200 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload), all inputs synthetic and generated on the device
(include/vrdd_synth.h):
  * decode: a VOL^3 distribution volume (default 1024^3 x 32 bins = 137 GB of raw
    histograms) is decoded z-slab by z-slab (default 256 slices = 34 GB per launch); every
    slab is resident in HBM before its timed region starts.  Reported in `decode` and in
    `roofline` (GB/s against the measured HBM peak), with the fractal-code decode of the same
    volume next to it.
  * ray cast (the `metric`): a "step" renders one IMG x IMG view (default 1024^2) of the
    decoded volume with the reference's constants (tstep 0.01, 500 steps, threshold 0.95,
    density 0.05, rainbow transfer function, queryMethod 1), cycling through the 64-view orbit
    of BASELINE.json configs[1].  Gsamples/s = transfer-function lookups / time; the lookups
    are counted exactly by the kernel in an untimed pass over the same views.
  * N > 1: image-space tiles (64x64, round-robin over ranks) of a frame that grows with N so
    every GPU keeps IMG^2 pixels (weak scaling); every rank decodes VOL/N z-slices and the
    decoded planes are all-gathered (NCCL); partial frames are reduced to rank 0 (NCCL).
Timing: CUDA events on the launching stream, >= 3 warm-up steps, barrier + synchronize on
both sides, max over ranks.  The sampled plane (4.3 GB) and every decode slab (>= 8 GB) are
far larger than the 126 MB L2, so no L2 flush is needed between iterations.
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


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return {"hbm_gbs": HBM_FALLBACK_GBS}, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, bytes_per_launch):
    """dram bytes per launch of `kernel` from the committed ncu capture (profiles/traffic.json),
    if that capture was taken on the same launch shape; else None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f).get(kernel)
        if t and abs(t["algorithmic_bytes_per_launch"] - bytes_per_launch) < 1e-6 * bytes_per_launch:
            return t["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


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
    """The oracle ray caster on a bounded volume: 256^3 decoded on the host, IMG^2 views of the
    same orbit with the reference's constants (the sample count per view hardly depends on
    the volume resolution because tstep is fixed)."""

    def __init__(self, seed, img):
        import numpy as np
        from oracle.vrdd_oracle import Oracle
        self.o = Oracle(fast=True)
        self.o.set_num_threads(host_threads())
        self.dims = (256, 256, 256)
        self.img = img
        vol = np.empty((256 ** 3, 4), np.float32)
        sl = 256 * 256
        for z0 in range(0, 256, 32):
            vol[z0 * sl:(z0 + 32) * sl] = self.o.decode_hist(self.o.synth_histograms(seed, self.dims, z0=z0, nz=32))
        self.vol = vol

    def step(self, k):
        view = self.o.view_matrix(0.0, (k % ORBIT_VIEWS) * (360.0 / ORBIT_VIEWS))
        _, s = self.o.render(self.vol, self.dims, view, image=self.img)
        return s


def run_reference(args):
    """--impl reference: the reference's CPU path.  The reference itself cannot be compiled
    (CUDA 12.9 removed texture references), so this is the OpenMP oracle port
    with all host threads, on the same metric/config as our arm.  One step = one view."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    img = (args.image, args.image)
    rc = CpuRaycaster(args.seed, img)
    warm = min(args.warmup, ORBIT_VIEWS)                # a step takes ~60 ms on 16 cores: K and W are honoured as given
    for k in range(warm):
        rc.step(k)
    steps = max(1, min(args.steps, 1024))
    t0 = time.perf_counter()
    samples = sum(rc.step(warm + k) for k in range(steps))
    dt = time.perf_counter() - t0
    val = samples / dt / 1e9
    sample = f"{steps} {img[0]}x{img[1]} views of the 64-view orbit on a 256^3 volume decoded on the host"
    line = {"impl": "reference", "metric": "raycast_throughput", "value": val, "unit": "Gsamples/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": dt / steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "fps": steps / dt,
            "config": {"workload": f"ray cast {img[0]}x{img[1]} orbit views, reference constants (tstep 0.01, 500 steps, "
                                   "threshold 0.95), OpenMP port of the reference's d_render on a bounded 256^3 volume"},
            "cpu_baseline": {"value": val, "unit": "Gsamples/s", "cores": rc.o.num_threads(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": val, "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------

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
    r.set_variant("decode_hist", args.decode_variant)
    r.set_variant("decode_fractal", args.fractal_variant)
    if world > 1:
        r.keep_linear_planes(True)
    r.set_volume(W, H, Dz)

    ev = lambda: torch.cuda.Event(enable_timing=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    clocks = ClockSampler(local)
    clocks.start()
    launches0 = r.kernel_launches()

    # ---- P1a: raw-histogram decode of my z-range, slab by slab -------------------------------
    z_lo, z_hi = D.slab_range(Dz, rank, world)
    slab = max(1, min(args.slab_z, z_hi - z_lo))
    n_slabs = (z_hi - z_lo + slab - 1) // slab
    reps = max(1, args.decode_reps)
    hist_buf = torch.empty(slab * slice_vox * 32, dtype=torch.float32, device=dev)
    dec_ms = 0.0
    for z0 in range(z_lo, z_hi, slab):
        nz = min(slab, z_hi - z0)
        r.synth_histograms_device(args.seed, z0, nz, hist_buf)     # untimed: the slab is resident before timing
        r.set_histograms_device(hist_buf, z0, nz)
        r.decode(V.SRC_ORIGINAL, z0, nz)                           # warm-up (also creates the arrays)
        barrier()
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(reps):
            r.decode(V.SRC_ORIGINAL, z0, nz)
        e1.record()
        torch.cuda.synchronize()
        dec_ms += e0.elapsed_time(e1) / reps
    del hist_buf
    torch.cuda.empty_cache()
    my_dec_ms = dec_ms
    dec_ms = max_over_ranks(dec_ms)
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

    # ---- P1b: fractal-code decode of the same volume (compact codes) --------------------------
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
        decode["fractal"] = {"kernel": fr_kernel.get(args.fractal_variant, args.fractal_variant), "ms": fr_ms,
                             "gbs": fr_bytes / (fr_ms * 1e-3) / 1e9,
                             "frac_of_hbm_peak": fr_bytes / (fr_ms * 1e-3) / 1e9 / hbm_peak,
                             "gvoxels_per_s": total_vox / (fr_ms * 1e-3) / 1e9, "bytes_per_voxel": fr_bytes / total_vox,
                             "templates": T, "mean_ne": (fr_bytes / total_vox - 28) / 8}
        del cb, er, off, tm
        torch.cuda.empty_cache()

    # ---- N > 1: replicate the decoded planes, one in-place all-gather per plane (NCCL) --------
    gather_ms = None
    if world > 1:
        planes = r.get_decoded_planes_device(V.SRC_ORIGINAL)
        barrier()
        e0, e1 = ev(), ev()
        e0.record()
        for p in planes:
            D.allgather_plane(V.as_torch(p, (Dz * slice_vox,), device=dev), Dz, slice_vox, rank, world)
        r.commit_planes(V.SRC_ORIGINAL, 0, Dz)
        e1.record()
        torch.cuda.synchronize()
        gather_ms = max_over_ranks(e0.elapsed_time(e1))

    # ---- P2: ray casting -----------------------------------------------------------------------
    fw, fh = D.frame_size(args.image, world)
    part = V.TilePartition(D.TILE, D.TILE, rank, world) if world > 1 else None
    img = torch.zeros(fh, fw, dtype=torch.int32, device=dev)
    red = torch.zeros_like(img) if world > 1 else None
    # N > 1, --assemble p2p: rank 0 owns two frames (double buffer); every other rank maps them (CUDA IPC) and
    # its ray-cast kernel stores its tiles straight into rank 0's HBM over NVLink; one barrier orders a frame.
    p2p = world > 1 and args.assemble == "p2p"
    frames = None
    if p2p:
        nbytes = fw * fh * 4
        mine = [r.frame_alloc(nbytes), r.frame_alloc(nbytes)] if rank == 0 else None
        box = [[r.frame_export(p) for p in mine]] if rank == 0 else [None]
        dist.broadcast_object_list(box, src=0)
        ok = 1
        try:
            frames = mine if rank == 0 else [r.frame_open(hb) for hb in box[0]]
        except V.VrddError:
            ok = 0
        t_ok = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
        if int(t_ok.item()) == 0:                      # some rank cannot map the frame: NCCL reduce instead
            if rank == 0:
                for p_ in mine:
                    r.frame_free(p_)
            elif ok:
                for p_ in frames:
                    r.frame_close(p_)
            p2p, frames = False, None

    def render_step(k, params):
        r.set_view(orbit_view(V, k))
        if p2p:
            r.render(frames[k & 1], fw, fh, params, part=part, clear_misses=True)
            dist.barrier()                                    # frame k is complete on rank 0
        else:
            r.render(img, fw, fh, params, part=part, clear_misses=True)
            if world > 1:
                D.reduce_frame(img, red, dst=0)

    def count_samples(params, nviews):
        r.count_samples(True)
        counts = []
        for k in range(nviews):
            r.set_view(orbit_view(V, k))
            r.render(img, fw, fh, params, part=part, clear_misses=True)
            counts.append(r.get_sample_count())
        r.count_samples(False)
        if world > 1:
            t = torch.tensor(counts, dtype=torch.int64, device=dev)
            dist.all_reduce(t)
            counts = [int(x) for x in t.tolist()]
        return counts

    def time_render(params, steps, warmup, with_reduce=True):
        for k in range(warmup):
            render_step(k, params) if with_reduce else (r.set_view(orbit_view(V, k)),
                                                        r.render(img, fw, fh, params, part=part, clear_misses=True))
        barrier()
        l0 = r.kernel_launches()
        e0, e1 = ev(), ev()
        e0.record()
        for k in range(warmup, warmup + steps):
            if with_reduce:
                render_step(k, params)
            else:
                r.set_view(orbit_view(V, k))
                r.render(img, fw, fh, params, part=part, clear_misses=True)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)), r.kernel_launches() - l0

    params = V.default_render_params(query_method=1)
    if p2p:
        # the frame the kernels assemble in rank 0's HBM must be, bit for bit, the NCCL-reduced one
        r.set_view(orbit_view(V, 7))
        r.render(img, fw, fh, params, part=part, clear_misses=True)
        D.reduce_frame(img, red, dst=0)
        r.render(frames[0], fw, fh, params, part=part, clear_misses=True)
        barrier()
        if rank == 0 and not torch.equal(V.as_torch(frames[0], (fh, fw), typestr="<i4", device=dev), red):
            raise SystemExit("bench.py: peer-assembled frame differs from the NCCL-reduced frame")
    nviews = min(ORBIT_VIEWS, args.warmup + args.steps)
    counts = count_samples(params, nviews)
    steps_samples = lambda c, w, s: sum(c[k % len(c)] for k in range(w, w + s))
    ray_ms, ray_launches = time_render(params, args.steps, args.warmup)
    samples = steps_samples(counts, args.warmup, args.steps)
    gsamples = samples / (ray_ms * 1e-3) / 1e9
    ms_per_step = ray_ms / args.steps
    kernel_ms, _ = time_render(params, args.steps, 0, with_reduce=False)      # ray-cast kernel alone
    kernel_ms /= args.steps

    # resolution-matched step: tstep = 2/N so the volume is sampled once per voxel
    matched = None
    if args.matched and world == 1:
        mp_ = V.default_render_params(query_method=1, tstep=2.0 / args.volume,
                                      max_steps=int(math.ceil(math.sqrt(3.0) * args.volume)) + 1)
        msteps = min(args.steps, 16)
        mcounts = count_samples(mp_, min(ORBIT_VIEWS, args.warmup + msteps))
        m_ms, _ = time_render(mp_, msteps, args.warmup)
        ms_ = steps_samples(mcounts, args.warmup, msteps)
        matched = {"tstep": 2.0 / args.volume, "max_steps": mp_.max_steps, "gsamples_per_s": ms_ / (m_ms * 1e-3) / 1e9,
                   "ms_per_step": m_ms / msteps, "fps": msteps / (m_ms * 1e-3), "samples_per_frame": ms_ / msteps}

    # ---- end to end through the C ABI with host buffers --------------------------------------
    host_img = torch.empty(fh, fw, dtype=torch.int32).pin_memory()
    e2e_sync = None
    if world == 1:
        # (i) one frame at a time: render, read back, synchronise
        for k in range(3):
            r.set_view(orbit_view(V, k)); r.render_host(host_img, fw, fh, params)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(args.warmup, args.warmup + args.steps):
            r.set_view(orbit_view(V, k))                         # copyInvViewMatrix: 48 B host -> kernel parameters
            r.render_host(host_img, fw, fh, params)              # render + D2H of the frame + synchronize
        dt_sync = time.perf_counter() - t0
        e2e_sync = {"value": samples / dt_sync / 1e9, "fps": args.steps / dt_sync,
                    "call": "vrdd_set_view + vrdd_render_host (render, device->pinned-host frame copy, synchronize) per frame"}
        # (ii) the orbit as a sequence: every frame still goes to host memory inside the timed region, but frame
        # k's read-back (second stream) overlaps frame k+1's rendering
        host_imgs = [host_img, torch.empty(fh, fw, dtype=torch.int32).pin_memory()]
        for k in range(3):
            r.set_view(orbit_view(V, k)); r.render_host_async(host_imgs[k & 1], fw, fh, params)
        r.render_host_wait()
        t0 = time.perf_counter()
        for k in range(args.warmup, args.warmup + args.steps):
            r.set_view(orbit_view(V, k))
            r.render_host_async(host_imgs[k & 1], fw, fh, params)
        r.render_host_wait()                                     # all frames are in host memory
        dt = time.perf_counter() - t0
        last = args.warmup + args.steps - 1                      # the pipelined frame is the synchronous frame
        check = torch.empty(fh, fw, dtype=torch.int32).pin_memory()
        r.set_view(orbit_view(V, last)); r.render_host(check, fw, fh, params)
        assert torch.equal(check, host_imgs[last & 1]), "pipelined read-back differs from the synchronous frame"
        call = ("vrdd_set_view + vrdd_render_host_async per frame, vrdd_render_host_wait at the end (two frames in "
                "flight: read-back of frame k overlaps rendering of frame k+1)")
    else:
        # N ranks fill ONE frame in shared host memory: full-width 64-row bands, round-robin over ranks; every rank
        # renders its bands and copies them to their place in the frame over its own PCIe link, the copy of frame k
        # overlapping the rendering of frame k+1.  A barrier per frame (behind a device-side fence on the previous
        # frame's copy) marks frame k-1 complete on all ranks.
        from multiprocessing import shared_memory
        import numpy as np
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
        t_ok = torch.tensor([shared_ok], dtype=torch.int32, device=dev)
        dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
        if int(t_ok.item()) == 1:
            bands = V.TilePartition(fw, 64, rank, world)
            band_counts = None
            barrier()
            for k in range(3):
                r.set_view(orbit_view(V, k)); r.render_host_async(host2[k & 1], fw, fh, params, part=bands)
            r.render_host_wait()
            barrier()
            t0 = time.perf_counter()
            for k in range(args.warmup, args.warmup + args.steps):
                r.set_view(orbit_view(V, k))
                r.render_host_async(host2[k & 1], fw, fh, params, part=bands)
                if args.e2e_sync_every and (k % args.e2e_sync_every) == 0:
                    r.render_host_fence(1)                        # stream waits for the copy of frame k-1 ...
                    dist.barrier()                                # ... then all ranks: frame k-1 is complete in host memory
            r.render_host_wait()
            barrier()
            dt = max_over_ranks(time.perf_counter() - t0)
            # the frame in shared host memory is, bit for bit, the one the device-side assembly gives
            last = args.warmup + args.steps - 1
            render_step(last, params)
            torch.cuda.synchronize()
            barrier()
            if rank == 0:
                src = V.as_torch(frames[last & 1], (fh, fw), typestr="<i4", device=dev) if p2p else red
                if not np.array_equal(src.cpu().numpy(), host2[last & 1]):
                    raise SystemExit("bench.py: the frame assembled in shared host memory differs from the device-assembled frame")
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
            for k in range(args.warmup, args.warmup + args.steps):
                render_step(k, params)
                if rank == 0:
                    src = V.as_torch(frames[k & 1], (fh, fw), typestr="<i4", device=dev) if p2p else red
                    host_img.copy_(src, non_blocking=True)
                torch.cuda.synchronize()
            barrier()
            dt = max_over_ranks(time.perf_counter() - t0)
            call = ("vrdd_set_view + vrdd_render (my tiles, stored into rank 0's frame over NVLink) + barrier + "
                    "device->pinned-host frame copy") if p2p else \
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
    if world == 1 and args.mode7:
        try:
            r.enable_interpolated_mean(True)                      # the block-mean plane is only kept on request
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
            n7 = min(args.steps, 32)
            c7 = count_samples(p7, min(ORBIT_VIEWS, args.warmup + n7))
            ms7, _ = time_render(p7, n7, args.warmup)
            s7 = steps_samples(c7, args.warmup, n7)
            mode7 = {"gsamples_per_s": s7 / (ms7 * 1e-3) / 1e9, "ms_per_step": ms7 / n7, "fps": n7 / (ms7 * 1e-3),
                     "steps": n7, "kernel": "raycast_mode7_kernel<gather>",
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
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import flex_synth                                   # synthetic span store (test data, numpy)
            tables = flex_synth.make_tables(5, 64, n_templates=469, block=6)
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
        rc = CpuRaycaster(args.seed, (args.image, args.image))
        rc.step(0)
        n_cpu = ORBIT_VIEWS                                     # the whole orbit the GPU arm renders, once (~4 s on 16 cores)
        t0 = time.perf_counter()
        s_cpu = sum(rc.step(k) for k in range(n_cpu))
        dt = time.perf_counter() - t0
        cpu_ray = {"value": s_cpu / dt / 1e9, "unit": "Gsamples/s", "cores": rc.o.num_threads(), "kind": "port",
                   "sample": f"the {n_cpu} {args.image}x{args.image} views of the orbit on a 256^3 volume decoded on the host "
                             "(same constants), OpenMP oracle -O3 -march=x86-64-v3", "fps": n_cpu / dt}

    if rank == 0:
        my_samples = samples / world / args.steps
        # the dominant kernel of the timed step is the ray caster; at 1024^3 with the reference's fixed step ncu
        # shows it DRAM-bound (profiles/README.md), so its roofline is HBM with 32 algorithmic bytes per sample
        ray_traffic, traffic_what = None, "no ncu capture for this configuration"
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)["raycast_kernel"]
            if args.volume == 1024 and world == 1 and args.image == 1024:
                ray_traffic = tj["dram_bytes_per_launch"]
                traffic_what = ("mean DRAM bytes of the %d orbit launches under ncu (cold L2 before each)" % tj["launches"]
                                if "launches" in tj else "DRAM bytes of one ncu-captured launch (an orbit side view)")
        except Exception:
            pass
        roofline = {"kernel": "raycast_kernel", "bound": "hbm",
                    "achieved": my_samples * SAMPLE_BYTES / (kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "peak_source": peak_src, "launch_ms": kernel_ms, "bytes_per_launch": my_samples * SAMPLE_BYTES,
                    "gsamples_per_s_kernel_only": my_samples / (kernel_ms * 1e-3) / 1e9, "traffic": ray_traffic,
                    "note": "algorithmic bytes = 32 B per trilinear sample (8 fp32 texels) x samples of the average launch; "
                            "traffic = " + traffic_what + ", profiles/traffic.json; sector over-fetch, not re-reads, "
                            "separates the two"}
        roofline["frac"] = roofline["achieved"] / roofline["peak"]
        if ray_traffic:                                   # what the DRAM actually moved per launch, against the same peak
            roofline["traffic_gbs"] = ray_traffic / (kernel_ms * 1e-3) / 1e9
            roofline["traffic_frac"] = roofline["traffic_gbs"] / roofline["peak"]
        line = {"metric": "raycast_throughput", "value": gsamples, "unit": "Gsamples/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "fps": 1e3 / ms_per_step, "samples_per_frame": samples / args.steps,
                "config": {"workload": f"{args.volume}^3 distribution volume (32 bins) decoded on device, ray cast "
                                       f"{fw}x{fh} per step over the 64-view orbit, reference constants "
                                       "(tstep 0.01, 500 steps, threshold 0.95, density 0.05, queryMethod 1)",
                           "volume": [W, H, Dz], "image": [fw, fh], "sampler": args.sampler,
                           "partition": "single GPU" if world == 1 else
                           f"64x64 image tiles round-robin over {world} ranks; z-slab decode + NCCL all-gather of the "
                           "decoded planes; " + ("tiles stored by the ray-cast kernels into rank 0's frame over NVLink "
                                                 "(CUDA IPC), one barrier per frame" if p2p else
                                                 "NCCL reduce of frames to rank 0"),
                           "l2": f"inputs larger than L2 ({total_vox * 4 / 1e9:.1f} GB sampled plane, "
                                 f"{launch_bytes / 1e9:.1f} GB per decode launch); no flush"},
                "decode": decode, "roofline": roofline, "roofline_decode": roofline_decode,
                "e2e": e2e, "gpu_launches": int(ray_launches), "gpu_launches_total": int(total_launches),
                "clocks": clk}
        if matched:
            line["raycast_resolution_matched"] = matched
        if cfg1:
            line["config_512"] = cfg1
        if mode7:
            line["query_method_7"] = mode7
        if flex:
            line["flex_chain"] = flex
        if gather_ms is not None:
            line["allgather_planes_ms"] = gather_ms
        if cpu_ray:
            line["cpu_baseline"] = cpu_ray
        print(json.dumps(line), flush=True)
    if p2p:
        barrier()
        for p in frames:
            (r.frame_free if rank == 0 else r.frame_close)(p)
    r.close()
    if world > 1:
        dist.destroy_process_group()


def run_sortlast(args):
    """--workload sortlast (BASELINE.json configs[4]): every rank owns one VOL^3 brick (plus a one-voxel
    ghost layer) of a volume too large for one GPU — 8 ranks: 2x2x2 bricks of a (2*VOL)^3 volume, 1.1 TB of
    histograms at VOL = 1024.  Per step: alpha pre-pass -> NCCL all-gather of the segment alphas ->
    incoming alpha -> colour pass -> NCCL SUM reduction of the float4 increments -> pack on rank 0."""
    import torch
    import torch.distributed as dist
    import vrdd_b200 as V
    import vrdd_b200.dist as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks, peak_src = measured_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", HBM_FALLBACK_GBS))
    grid = D.brick_grid(world)
    q = D.brick_of_rank(rank, grid)
    E = args.volume
    gdims = (E * grid[0], E * grid[1], E * grid[2])
    origin, size, lo, hi = D.brick_geometry(gdims, grid, q)
    r = V.Renderer(local)
    r.set_stream(torch.cuda.current_stream().cuda_stream)
    r.set_sampler({"texture": V.SAMPLER_TEXTURE, "linear": V.SAMPLER_LINEAR, "bricked": V.SAMPLER_BRICKED}[args.sortlast_layout])
    r.set_volume(*size)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    clocks = ClockSampler(local)
    clocks.start()
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

    def step(k):
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

    nviews = min(ORBIT_VIEWS, args.warmup + args.steps)
    r.count_samples(True)
    counts = []
    for k in range(nviews):
        step(k)
        counts.append(r.get_sample_count())
    r.count_samples(False)
    if world > 1:
        t = torch.tensor(counts, dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        counts = [int(x) for x in t.tolist()]
    for k in range(args.warmup):
        step(k)
    barrier()
    l0 = r.kernel_launches()
    e0, e1 = ev(), ev()
    e0.record()
    for k in range(args.warmup, args.warmup + args.steps):
        step(k)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = r.kernel_launches() - l0
    samples = sum(counts[k % len(counts)] for k in range(args.warmup, args.warmup + args.steps))
    host = torch.empty(fh, fw, dtype=torch.int32).pin_memory()
    barrier()
    t0 = time.perf_counter()
    for k in range(args.warmup, args.warmup + args.steps):
        step(k)
        if rank == 0:
            host.copy_(frame, non_blocking=True)
        torch.cuda.synchronize()
    barrier()
    dt = max_over_ranks(time.perf_counter() - t0)
    clk = clocks.stop()
    if rank == 0:
        dec_gbs = decoded_vox * HIST_BYTES_PER_VOXEL / (dec_ms * 1e-3) / 1e9
        line = {"metric": "raycast_throughput", "value": samples / (ms * 1e-3) / 1e9, "unit": "Gsamples/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "fps": args.steps / (ms * 1e-3),
                "samples_per_frame": samples / args.steps,
                "config": {"workload": f"sort-last: {gdims[0]}x{gdims[1]}x{gdims[2]} distribution volume in {grid[0]}x{grid[1]}x{grid[2]} "
                                       f"bricks of {E}^3 (+1 ghost) over {world} GPUs, {fw}x{fh} frames of the 64-view orbit, reference "
                                       "constants; alpha pre-pass, NCCL all-gather of segment alphas, colour pass, NCCL SUM "
                                       "reduction of float4 increments, pack on rank 0"
                                       + ("; collectives restricted to the image rows of each brick's screen footprint" if windows else ""),
                           "collective_bytes_per_frame": {"all_gather": coll_bytes[0], "reduce": coll_bytes[1],
                                                          "full_frame": [world * fh * fw * 4, fh * fw * 16]},
                           "volume": list(gdims), "image": [fw, fh], "histogram_bytes_total": decoded_vox * 128,
                           "l2": "inputs larger than L2; no flush"},
                "decode": {"kernel": "decode_hist_tma_kernel", "voxels": decoded_vox, "ms": dec_ms, "gbs": dec_gbs,
                           "frac_of_hbm_peak": dec_gbs / (hbm_peak * world), "peak_gbs": hbm_peak * world},
                "e2e": {"value": samples / dt / 1e9, "unit": "Gsamples/s", "h2d_bytes_per_step": 80,
                        "d2h_bytes_per_step": fw * fh * 4, "fps": args.steps / dt,
                        "call": "vrdd_render_brick_alpha/compose/color + NCCL + vrdd_pack_frame + device->pinned-host frame copy"},
                "gpu_launches": int(launches), "clocks": clk}
        print(json.dumps(line), flush=True)
    r.close()
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
    ap.add_argument("--slab-z", type=int, default=256, help="z-slices decoded per launch (256 -> 34 GB of histograms)")
    ap.add_argument("--decode-reps", type=int, default=3)
    ap.add_argument("--decode-variant", default="tma", choices=["tma", "ldg"])
    ap.add_argument("--fractal-variant", default="moments2",
                    choices=["moments2", "moments2r", "moments", "moments768", "moments_global", "dense"])
    ap.add_argument("--sampler", default="texture", choices=["texture", "bricked"])
    ap.add_argument("--tf", default=None, choices=[None, "texture", "smem"])
    ap.add_argument("--unroll", type=int, default=0, choices=[0, 1, 2, 4, 8])
    ap.add_argument("--fractal", type=int, default=1)
    ap.add_argument("--matched", type=int, default=1)
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
