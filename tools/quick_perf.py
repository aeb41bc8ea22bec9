"""Quick device-side timing of each kernel class on the synthetic corpora (dev tool)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import deflate_compression_b200 as D
from deflate_compression_b200 import synth

def run(name, data, reps=3, **prm):
    dev = torch.device("cuda:0")
    d_in = torch.from_numpy(data).to(dev)
    comp = D.Compressor(**prm)
    comp.enable_timing(True)
    cap = comp.bound(data.size)
    d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
    best = None
    for r in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st = torch.cuda.current_stream().cuda_stream
        e0.record()
        comp.enqueue(d_in.data_ptr(), data.size, 0, True, d_out.data_ptr(), cap, st)
        e1.record()
        out_len, adler = comp.finish(st)
        ms = e0.elapsed_time(e1)
        if best is None or ms < best[0]:
            best = (ms, comp.kernel_ms(), comp.phase_cycles())
        fb = comp.sort_fallbacks()
        if fb:
            print("   !! sort fallbacks:", fb)
    ms, k, ph = best
    nseg = (data.size + 32767) // 32768
    print("   per-segment k_match cycles: " + " ".join(f"{a}={b/nseg:.0f}" for a, b in ph.items() if b))
    print(f"{name:28s} n={data.size>>20:5d} MiB out={out_len/data.size:.4f} {ms:9.2f} ms {data.size/ms/1e6:8.2f} GB/s  " +
          " ".join(f"{a}={b:.2f}" for a, b in k.items()), flush=True)
    comp.close()

if __name__ == "__main__":
    mb = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    n = mb << 20
    run("markov", synth.markov(n, 2))
    run("markov chain_limit=16", synth.markov(n, 2), chain_limit=16)
    run("markov chain_limit=4", synth.markov(n, 2), chain_limit=4)
    run("markov window=4096", synth.markov(n, 2), window=4096)
    run("log", synth.log(n, 3))
    run("random", synth.random_bytes(n, 4))
    run("png", synth.png_rows(n, 6))
    run("zeros", np.zeros(n, dtype=np.uint8))
