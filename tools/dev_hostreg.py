import time, numpy as np, torch, ctypes, os
torch.cuda.init()
rt = torch.cuda.cudart()
print("cpus", os.cpu_count())
for gb in (1, 4):
    n = gb * (1 << 30) // 8
    a = np.ones(n)                       # touched pageable memory
    t0 = time.perf_counter()
    r = rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0)
    t1 = time.perf_counter()
    d = torch.empty(n, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    d.copy_(torch.from_numpy(a), non_blocking=True); torch.cuda.synchronize()
    t3 = time.perf_counter()
    rt.cudaHostUnregister(a.ctypes.data)
    t4 = time.perf_counter()
    print(f"{gb} GB: register {1e3*(t1-t0):.1f} ms ({gb/(t1-t0):.1f} GB/s) rc={r}; copy {1e3*(t3-t2):.1f} ms ({gb/(t3-t2):.1f} GB/s); unregister {1e3*(t4-t3):.1f} ms")
    b = np.ones(n)
    t0 = time.perf_counter(); d.copy_(torch.from_numpy(b)); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"   plain pageable copy {1e3*(t1-t0):.1f} ms ({gb/(t1-t0):.1f} GB/s)")
    c = np.empty(n); t0 = time.perf_counter(); np.copyto(c, b); t1 = time.perf_counter()
    print(f"   single-thread memcpy {gb/(t1-t0):.1f} GB/s")
    del a, b, c, d
