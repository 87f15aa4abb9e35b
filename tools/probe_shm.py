import os, mmap, ctypes, torch, numpy as np
from multiprocessing import shared_memory
print(os.popen("df -h /dev/shm /tmp | cat").read())
rt = torch.cuda.cudart()
print("has cudaHostRegister:", hasattr(rt, "cudaHostRegister"))
torch.cuda.init(); torch.zeros(1, device="cuda")
size = 80 << 20
try:
    shm = shared_memory.SharedMemory(create=True, size=size)
    a = np.ndarray((size,), dtype=np.uint8, buffer=shm.buf)
    a[:] = 0
    r = rt.cudaHostRegister(a.ctypes.data, size, 0)
    print("shm register:", r)
    t = torch.zeros(size, dtype=torch.uint8, device="cuda") + 7
    import time
    lib = ctypes.CDLL("libcudart.so") if False else None
    # copy via torch: wrap as tensor
    ht = torch.from_numpy(a)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): ht.copy_(t, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("D2H GB/s into registered shm:", 10 * size / dt / 1e9, int(a[12345]))
    rt.cudaHostUnregister(a.ctypes.data)
    del ht, a
    shm.close(); shm.unlink()
except Exception as e:
    print("shm failed:", repr(e))
try:
    fn = "/tmp/vrdd_probe_frame.bin"
    with open(fn, "wb") as f: f.truncate(size)
    m = np.memmap(fn, dtype=np.uint8, mode="r+", shape=(size,))
    m[:] = 0
    r = rt.cudaHostRegister(m.ctypes.data, size, 0)
    print("file mmap register:", r)
    rt.cudaHostUnregister(m.ctypes.data)
    del m; os.remove(fn)
except Exception as e:
    print("mmap failed:", repr(e))
