"""Timeline of one deflate_compress(memfd, memfd, -1, 32768, 0) call on the benchmark text (dev tool): DFL_B200_FD_TRACE=1."""
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
for rep in range(3):
    os.lseek(fi, 0, os.SEEK_SET); os.lseek(fo, 0, os.SEEK_SET)
    if rep == 2: os.environ["DFL_B200_FD_TRACE"] = "1"
    t0 = time.perf_counter(); rc = D.deflate_compress(fi, fo, -1, 32768, 0); dt = time.perf_counter() - t0
    print(f"rep {rep}: rc {rc} {1e3*dt:.2f} ms {n/dt/1e9:.2f} GB/s", flush=True)
