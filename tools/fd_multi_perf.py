"""deflate_compress(memfd, memfd, -1, 32768, 0) on 1 GiB of benchmark text, one device against DFL_B200_DEVICES lists (dev tool)."""
import os, sys, time, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import deflate_compression_b200 as D
import bench
n = 1 << 30
host = np.concatenate([bench.corpus(0, 256 << 20)] * 4)
fi, fo = os.memfd_create("i"), os.memfd_create("o")
v = memoryview(host); p = 0
while p < len(v): p += os.write(fi, v[p:p + (64 << 20)])
lists = [None, "0,0"] + (["0,1"] if torch.cuda.device_count() > 1 else []) + (["0,1,2,3"] if torch.cuda.device_count() > 3 else [])
ref = None
for devs in lists:
    if devs: os.environ["DFL_B200_DEVICES"] = devs
    ts = []
    for rep in range(4):
        os.lseek(fi, 0, 0); os.lseek(fo, 0, 0)
        t0 = time.perf_counter(); rc = D.deflate_compress(fi, fo, -1, 32768, 0); ts.append(time.perf_counter() - t0)
    size = os.lseek(fo, 0, 1)
    z = os.pread(fo, size, 0)
    if ref is None: ref = z
    print(devs, rc, "best %.2f GB/s" % (n / min(ts[1:]) / 1e9), ["%.1f" % (1e3 * t) for t in ts], size, "same stream" if z == ref else "DIFFERENT", flush=True)
print("inflates:", zlib.decompress(ref) == host.tobytes())
