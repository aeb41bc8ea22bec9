"""fd -> fd throughput over memfds for several piece sizes (dev tool)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import deflate_compression_b200 as D
import bench
n = 256 << 20
host = bench.corpus(0, n)
fi, fo = os.memfd_create("dfl_in"), os.memfd_create("dfl_out")
view = memoryview(host); pos = 0
while pos < len(view): pos += os.write(fi, view[pos:pos + (64 << 20)])
comp = D.Compressor()
for piece in (16, 24, 32, 48, 64, 96):
    ts = []
    for rep in range(4):
        os.lseek(fi, 0, os.SEEK_SET); os.lseek(fo, 0, os.SEEK_SET)
        t0 = time.perf_counter(); comp.compress_fd(fi, fo, piece << 20); ts.append(time.perf_counter() - t0)
    print(f"piece {piece} MiB: best {n/min(ts[1:])/1e9:.2f} GB/s ({1e3*min(ts[1:]):.2f} ms)", flush=True)
