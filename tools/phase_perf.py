"""Per-phase k_match cycles on one corpus (dev tool): python tools/phase_perf.py [markov|log|random|png] [MiB]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import deflate_compression_b200 as D
from deflate_compression_b200 import synth
kind = sys.argv[1] if len(sys.argv) > 1 else "markov"
n = (int(sys.argv[2]) if len(sys.argv) > 2 else 256) << 20
data = {"markov": lambda: synth.markov(n, 2), "log": lambda: synth.log(n, 3), "random": lambda: synth.random_bytes(n, 4), "png": lambda: synth.png_rows(n, 6)}[kind]()
dev = torch.device("cuda:0"); d_in = torch.from_numpy(data).to(dev)
for mode in (0, 1):
    comp = D.Compressor(); comp.set_match_mode(mode); comp.enable_timing(True)
    cap = comp.bound(n); d_out = torch.empty(cap, dtype=torch.uint8, device=dev); st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        comp.enqueue(d_in.data_ptr(), n, 0, True, d_out.data_ptr(), cap, st); comp.finish(st)
    ph = comp.phase_cycles(); nseg = n // 32768
    print(f"{kind} {n>>20} MiB match_mode={mode}: k_match {comp.kernel_ms()['k_match']:.2f} ms; cycles/seg: " + " ".join(f"{a}={b/nseg/1e3:.1f}K" for a, b in ph.items() if b and not a.startswith('sx') and a not in ('plain_segments','warp_end_avg','warp_end_max')))
    comp.close()
