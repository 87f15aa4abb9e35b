"""Host-side multi-GPU logic: how the two hot paths shard over ranks.

The reference is single-GPU.  Both paths shard naturally:
  * decode: voxels are independent -> each rank decodes a contiguous range of z-slices; the
    decoded planes are replicated with ONE all-gather (each rank's slab is contiguous in the
    x-fastest plane, so it is gathered in place);
  * ray casting of a volume that fits one GPU: image-space tiles, round-robin over ranks
    (centre tiles cost more than border tiles, so they are interleaved); every pixel is
    computed by exactly one rank with identical math, and a partial frame is zero outside
    its rank's tiles, so a SUM reduction to rank 0 assembles the frame bit for bit.
Everything here is backend-agnostic torch.distributed (NCCL on the GPUs, gloo in the CPU
tests); there is no compute in this module.
"""
import torch
import torch.distributed as dist

TILE = 64


def slab_range(depth, rank, world):
    """z-slices [lo, hi) decoded by `rank`: contiguous, balanced, covering [0, depth)."""
    return depth * rank // world, depth * (rank + 1) // world


def frame_size(per_gpu_edge, world):
    """Weak scaling: the frame grows with the number of ranks so that every GPU keeps
    per_gpu_edge^2 pixels: 1 -> E x E, 2 -> 2E x E, 4 -> 2E x 2E, 8 -> 4E x 2E."""
    if world in (1, 2, 4, 8):
        w = per_gpu_edge * (2 if world in (2, 8) else 1) * (2 if world >= 4 else 1)
        h = per_gpu_edge * (2 if world >= 4 else 1)
        return w, h
    return per_gpu_edge * world, per_gpu_edge


def tile_owner(width, height, tile_w=TILE, tile_h=TILE, parts=1):
    """int32[height][width]: which part renders each pixel (tile index, row-major, modulo
    parts) — the rule vrdd_render applies (include/vrdd.h, vrdd_tile_partition)."""
    tiles_x = (width + tile_w - 1) // tile_w
    ys = torch.arange(height).unsqueeze(1) // tile_h
    xs = torch.arange(width).unsqueeze(0) // tile_w
    return ((ys * tiles_x + xs) % parts).to(torch.int32)


def allgather_plane(full, depth, slice_vox, rank, world, group=None):
    """Replicates a decoded plane: `full` is float[depth*slice_vox] holding this rank's
    z-slab at its final position; on return every rank holds the whole plane."""
    lo, hi = slab_range(depth, rank, world)
    mine = full[lo * slice_vox:hi * slice_vox].clone()
    if world == 1:
        return full
    if depth % world == 0:
        dist.all_gather_into_tensor(full, mine, group=group)
    else:
        # uneven slabs: all_gather needs equal sizes on every backend, so each slab is broadcast
        for q in range(world):
            a, b = slab_range(depth, q, world)
            dist.broadcast(full[a * slice_vox:b * slice_vox], src=q, group=group)
    return full


def reduce_frame(partial, scratch, dst=0, group=None):
    """Assembles the frame on `dst` from per-rank partial frames (zero outside a rank's own
    tiles).  `partial` is left untouched so the next frame can be rendered into it."""
    scratch.copy_(partial)
    dist.reduce(scratch, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return scratch
