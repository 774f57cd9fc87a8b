"""PCIe copy latencies on this box: back-to-back cudaMemcpyAsync on one stream (pinned host memory)."""
import time, torch
dev = torch.device("cuda:0")
def bench(nbytes, direction, n=1000):
    h = torch.empty(nbytes * 64, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes * 64, dtype=torch.uint8, device=dev)
    def go(k):
        for i in range(k):
            o = (i % 64) * nbytes
            if direction == "h2d": d[o:o + nbytes].copy_(h[o:o + nbytes], non_blocking=True)
            else: h[o:o + nbytes].copy_(d[o:o + nbytes], non_blocking=True)
    go(50); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(); go(n); e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    print(f"{direction} {nbytes:8d} B: device {1e3*e0.elapsed_time(e1)/n:6.2f} us/copy  enqueue {1e6*(t1-t0)/n:6.2f} us/copy", flush=True)
for nb in (4, 64, 65536, 131072, 1 << 20):
    bench(nb, "h2d"); bench(nb, "d2h")
import subprocess
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
print(subprocess.run(["bash", "-c", "lscpu | head -20; nproc"], capture_output=True, text=True).stdout)
