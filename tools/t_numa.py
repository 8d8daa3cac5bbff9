"""Is pinned-memory H2D bandwidth NUMA dependent on this box? (scratch tool)"""
import os, sys, time, glob, torch
sys.path.insert(0, '/root/repo')
dev = torch.device('cuda', 0)
prop = torch.cuda.get_device_properties(0)
bus = f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0" if hasattr(prop, "pci_bus_id") else None
print("pci", bus, "cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for f in ("numa_node", "local_cpulist"):
    try: print(f, open(f"/sys/bus/pci/devices/{bus}/{f}").read().strip())
    except Exception as e: print(f, "n/a", e)
nodes = sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))
for nd in nodes:
    print(nd.split('/')[-1], open(nd + "/cpulist").read().strip())
d = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
def bw(h):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); return 3 * h.numel() / (time.perf_counter() - t0) / 1e9
all_cpus = sorted(os.sched_getaffinity(0))
for nd in nodes:
    cl = open(nd + "/cpulist").read().strip()
    cpus = set()
    for part in cl.split(','):
        if '-' in part: a, b = part.split('-'); cpus.update(range(int(a), int(b) + 1))
        elif part: cpus.add(int(part))
    cpus &= set(all_cpus)
    if not cpus: print(nd.split('/')[-1], "no allowed cpus"); continue
    os.sched_setaffinity(0, cpus)
    h = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True); h.fill_(1)
    print(nd.split('/')[-1], "pinned alloc from its cpus -> H2D GB/s", round(bw(h), 1), round(bw(h), 1))
    del h
os.sched_setaffinity(0, all_cpus)
