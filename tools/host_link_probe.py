"""Host<->device link probe: pinned and pageable copy bandwidth in each direction and both at once,
host memcpy bandwidth (1..n threads), host memory and core counts.  usage: python tools/host_link_probe.py"""
import os
import threading
import time

import numpy as np
import torch

print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for ln in open("/proc/meminfo"):
    if ln.split(":")[0] in ("MemTotal", "MemAvailable", "HugePages_Total"):
        print(ln.strip())
n = 1 << 31
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
dev2 = torch.empty(n, dtype=torch.uint8, device="cuda")
t0 = time.perf_counter()
pin = torch.empty(n, dtype=torch.uint8, pin_memory=True)
print(f"pin 2 GiB: {time.perf_counter() - t0:.2f} s")
pin2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
pag = torch.empty(n, dtype=torch.uint8)
pag.zero_()


def bw(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return n * reps / (time.perf_counter() - t0) / 1e9


print(f"pinned   H2D {bw(lambda: dev.copy_(pin, non_blocking=True)):.1f} GB/s   D2H {bw(lambda: pin.copy_(dev, non_blocking=True)):.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def both():
    with torch.cuda.stream(s1):
        dev.copy_(pin, non_blocking=True)
    with torch.cuda.stream(s2):
        pin2.copy_(dev2, non_blocking=True)


print(f"pinned   H2D+D2H concurrently {2 * bw(both):.1f} GB/s total")
print(f"pageable H2D {bw(lambda: dev.copy_(pag)):.1f} GB/s   D2H {bw(lambda: pag.copy_(dev)):.1f} GB/s")
t0 = time.perf_counter()
torch.cuda.cudart().cudaHostRegister(pag.data_ptr(), n, 0)
print(f"cudaHostRegister 2 GiB: {time.perf_counter() - t0:.2f} s")
print(f"registered H2D {bw(lambda: dev.copy_(pag, non_blocking=True)):.1f} GB/s")
torch.cuda.cudart().cudaHostUnregister(pag.data_ptr())
a = np.frombuffer(pag.numpy(), dtype=np.uint8)
b = np.frombuffer(pin.numpy(), dtype=np.uint8)
for nt in (1, 2, 4, 8, 16):
    if nt > (os.cpu_count() or 1):
        break
    chunk = n // nt
    def work(k):
        np.copyto(b[k * chunk:(k + 1) * chunk], a[k * chunk:(k + 1) * chunk])
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(k,)) for k in range(nt)]
    [t.start() for t in th]
    [t.join() for t in th]
    print(f"host memcpy pageable->pinned, {nt} threads: {n / (time.perf_counter() - t0) / 1e9:.1f} GB/s")
