"""Aggregate device -> host bandwidth of N GPUs copying into ONE buffer in POSIX shared memory (page-locked by every rank),
the ingest that bounds bench.py's end-to-end number at N > 1.  Every rank copies its own contiguous share
of the buffer over its own PCIe link; all ranks start on a barrier and the slowest one defines the time.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/probe_shm_n.py [MB per rank] [frame MB]
Prints one JSON line on rank 0: aggregate and per-rank GB/s for one large copy per rank and for frame-sized pieces
(the bench's granularity: a 64-row band of a 4096 x 2048 frame is 1 MB)."""
import json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
from multiprocessing import shared_memory

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vrdd_b200 as V

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
per_rank = (int(sys.argv[1]) if len(sys.argv) > 1 else 256) << 20
piece = (int(sys.argv[2]) if len(sys.argv) > 2 else 1) << 20
total = per_rank * world
shm = None
if rank == 0:
    shm = shared_memory.SharedMemory(create=True, size=total)
box = [shm.name if shm is not None else None]
if world > 1:
    dist.broadcast_object_list(box, src=0)
if rank != 0:
    shm = shared_memory.SharedMemory(name=box[0])
    try:
        from multiprocessing import resource_tracker
        resource_tracker.unregister(shm._name, "shared_memory")
    except Exception:
        pass
host = np.ndarray((total,), dtype=np.uint8, buffer=shm.buf)
mine = host[rank * per_rank:(rank + 1) * per_rank]
mine[:] = 0                                                  # first touch by the rank that fills it
V.host_register(host.ctypes.data, total)
src = torch.full((per_rank,), rank + 1, dtype=torch.uint8, device=dev)
dst = torch.from_numpy(mine)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def timed(fn, reps):
    fn(); barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    mine_s = time.perf_counter() - t0
    barrier()
    t = torch.tensor([mine_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), mine_s


def whole():
    dst.copy_(src, non_blocking=True)


def pieces():                                                # round-robin pieces like the bench's bands
    for o in range(0, per_rank, piece):
        dst[o:o + piece].copy_(src[o:o + piece], non_blocking=True)


reps = 8
t_whole, my_whole = timed(whole, reps)
t_piece, my_piece = timed(pieces, reps)
ok = int(host[rank * per_rank + 12345]) == rank + 1
rates = torch.tensor([reps * per_rank / my_whole / 1e9, reps * per_rank / my_piece / 1e9], dtype=torch.float64, device=dev)
allr = [torch.zeros_like(rates) for _ in range(world)]
if world > 1:
    dist.all_gather(allr, rates)
else:
    allr = [rates]
barrier()
V.host_unregister(host.ctypes.data)
del dst, mine, host
shm.close()
if rank == 0:
    shm.unlink()
    print(json.dumps({"n_gpus": world, "bytes_per_rank": per_rank, "piece_bytes": piece, "data_ok": ok,
                      "aggregate_gbs_whole_copies": reps * total / t_whole / 1e9,
                      "aggregate_gbs_pieces": reps * total / t_piece / 1e9,
                      "per_rank_gbs_whole": [float(a[0]) for a in allr], "per_rank_gbs_pieces": [float(a[1]) for a in allr]}), flush=True)
if world > 1:
    dist.destroy_process_group()
