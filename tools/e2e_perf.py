"""End-to-end (host buffers) throughput of dfl_b200_compress_host on the benchmark corpus (dev tool)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import deflate_compression_b200 as D
import bench
n = 256 << 20
host = bench.corpus(0, n)
h_in = torch.from_numpy(host).pin_memory()
comp = D.Compressor()
cap = comp.bound(n)
h_out = torch.empty(cap + 16, dtype=torch.uint8).pin_memory()
for _ in range(2):
    z = comp.compress_into(h_in.data_ptr(), n, h_out.data_ptr(), cap + 16)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    t0 = time.perf_counter(); z = comp.compress_into(h_in.data_ptr(), n, h_out.data_ptr(), cap + 16); ts.append(time.perf_counter() - t0)
print(f"e2e best {n/min(ts)/1e9:.2f} GB/s  median {n/sorted(ts)[2]/1e9:.2f} GB/s  ({1e3*min(ts):.2f} ms)  out {z}")
