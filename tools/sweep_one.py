import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import deflate_compression_b200 as D
from deflate_compression_b200 import synth
def run(data, **prm):
    dev = torch.device("cuda:0"); d_in = torch.from_numpy(data).to(dev)
    comp = D.Compressor(**prm); comp.enable_timing(True)
    cap = comp.bound(data.size); d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
    best = 1e9
    for r in range(3):
        comp.enqueue(d_in.data_ptr(), data.size, 0, True, d_out.data_ptr(), cap, torch.cuda.current_stream().cuda_stream)
        comp.finish(torch.cuda.current_stream().cuda_stream)
        best = min(best, comp.kernel_ms()["k_match"])
    ph = comp.phase_cycles(); nseg = data.size // 32768
    return best, ph["build"] / nseg, ph["search"] / nseg, ph["plain_segments"] / nseg
n = int(os.environ.get("SWEEP_MB", "64")) << 20
res = []
for name, data, prm in (("text", synth.markov(n, 2), {}), ("text_k16", synth.markov(n, 2), dict(chain_limit=16)), ("log", synth.log(n, 3), {}), ("png", synth.png_rows(n, 6), {}), ("rand", synth.random_bytes(n, 4), {}), ("zeros", np.zeros(n, dtype=np.uint8), {})):
    ms, b, s, pl = run(data, **prm)
    res.append(f"{name}: {ms:6.2f} ms (build {b/1e3:4.0f}K search {s/1e3:5.0f}K plain {pl:.2f})")
print(" | ".join(res))
