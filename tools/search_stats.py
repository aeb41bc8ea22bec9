"""Lane-state census of k_match's search loop (needs a library built with -DDFL_SEARCH_STATS; dev tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import deflate_compression_b200 as D
from deflate_compression_b200 import synth
n = 16 << 20
for name, data, prm in (("text", synth.markov(n, 2), {}), ("png", synth.png_rows(n, 6), {}), ("rand", synth.random_bytes(n, 4), {}), ("log", synth.log(n, 3), {})):
    dev = torch.device("cuda:0"); d_in = torch.from_numpy(data).to(dev)
    comp = D.Compressor(**prm); cap = comp.bound(n); d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    comp.enqueue(d_in.data_ptr(), n, 0, True, d_out.data_ptr(), cap, st); comp.finish(st)
    ph = comp.phase_cycles(); nseg = n // 32768; it = max(ph["sx_iter"], 1)
    print(f"{name}: search {ph['search']/nseg/1e3:.0f}K cyc/seg; per seg: iters/warp {it/nseg/32:.0f} heads {ph['sx_heads']/nseg:.0f} hops {ph['sx_hops']/nseg:.0f} seeks {ph['sx_seek']/nseg:.0f}; "
          f"per iter: walk {ph['sx_walk']/it:.1f} fin {ph['sx_fin']/it:.1f} done {ph['sx_done']/it:.1f} cand {ph['sx_cand']/it:.1f}; "
          f"warps run out of work after avg {ph['warp_end_avg']/nseg/1e3:.0f}K / last {ph['warp_end_max']/nseg/1e3:.0f}K cycles")
