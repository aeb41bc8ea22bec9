"""Device-side timing of k_inflate on the synthetic corpora (dev tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import deflate_compression_b200 as D
from deflate_compression_b200 import synth

def run(name, data):
    n = data.size
    dev = torch.device("cuda:0")
    d_in = torch.from_numpy(data).to(dev)
    comp = D.Compressor()
    comp.enable_timing(True)
    cap = comp.bound(n)
    d_z = torch.empty(cap, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    out_len, adler = comp.deflate_device(d_in.data_ptr(), n, 0, True, d_z.data_ptr(), cap, st)
    nseg = (n + 32767) // 32768
    d_off = torch.from_numpy(comp.seg_index().astype(np.int64)).to(dev)
    d_back = torch.zeros(n, dtype=torch.uint8, device=dev)
    best = 1e9
    for _ in range(3):
        a2 = comp.inflate_device(d_z.data_ptr(), out_len, d_off.data_ptr(), nseg, d_back.data_ptr(), n, 0, st)
        k = comp.kernel_ms()
        tot = k["k_inflate"] + k["k_inf_resolve"] + k["k_inf_finish"]
        if tot < best:
            best, parts = tot, k
    ok = a2 == adler and bool(torch.equal(d_back, d_in))
    print(f"{name:10s} n={n>>20:5d} MiB ratio {n/out_len:7.2f}  inflate {best:8.2f} ms (decode {parts["k_inflate"]:.2f} resolve {parts["k_inf_resolve"]:.2f} finish {parts["k_inf_finish"]:.2f})  {n/best/1e6:8.2f} GB/s out  ok={ok} rounds={comp.inflate_rounds()}", flush=True)
    comp.close()

if __name__ == "__main__":
    mb = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    n = mb << 20
    run("markov", synth.markov(n, 2))
    run("log", synth.log(n, 3))
    run("random", synth.random_bytes(n, 4))
    run("png", synth.png_rows(n, 6))
