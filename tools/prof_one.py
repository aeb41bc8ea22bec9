"""One compress call on a small corpus (target for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import deflate_compression_b200 as D
from deflate_compression_b200 import synth
kind = sys.argv[1] if len(sys.argv) > 1 else "markov"
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 16
n = mb << 20
data = {"markov": lambda: synth.markov(n, 2), "log": lambda: synth.log(n, 3), "random": lambda: synth.random_bytes(n, 4)}[kind]()
dev = torch.device("cuda:0")
d_in = torch.from_numpy(data).to(dev)
comp = D.Compressor()
cap = comp.bound(n)
d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
for _ in range(2):
    out_len, adler = comp.deflate_device(d_in.data_ptr(), n, 0, True, d_out.data_ptr(), cap, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print(kind, n, out_len)
