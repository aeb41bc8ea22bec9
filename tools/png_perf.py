"""Device-side timing of k_png_filter and the PNG encode path (dev tool)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import deflate_compression_b200 as D

h, w, c = (int(sys.argv[1]) if len(sys.argv) > 1 else 21845), 4096, 3
rng = np.random.default_rng(1)
yy, xx = np.mgrid[0:64, 0:w]
tile = np.stack([(xx * (k + 1) // 16 + yy * (3 - k) // 8 + 40 * k) for k in range(c)], axis=-1)
img = np.concatenate([tile] * ((h + 63) // 64), axis=0)[:h]
img = ((img + rng.integers(-2, 3, size=img.shape)) & 255).astype(np.uint8)
dev = torch.device("cuda:0")
d_pix = torch.from_numpy(img.reshape(-1)).to(dev)
n = img.size
d_out = torch.zeros(h * (w * c + 1), dtype=torch.uint8, device=dev)
comp = D.Compressor(); comp.enable_timing(True)
st = torch.cuda.current_stream().cuda_stream
for mode in (0, 1, 5):
    best = 1e9
    for _ in range(5):
        comp.png_filter_device(d_pix.data_ptr(), w * c, h, c, d_out.data_ptr(), mode, st)
        best = min(best, comp.kernel_ms()["k_png_filter"])
    print(f"k_png_filter mode {mode}: {n>>20} MiB  {best:.3f} ms  {n/best/1e6:.1f} GB/s of pixels  ({2*n/best/1e6:.1f} GB/s read+write)", flush=True)
small = img[:2048]
t0 = time.perf_counter(); png = comp.png_encode(small); t1 = time.perf_counter()
png = comp.png_encode(small); t2 = time.perf_counter()
print(f"png_encode {small.nbytes>>20} MiB image -> {len(png)>>10} KiB in {1e3*(t2-t1):.1f} ms ({small.nbytes/(t2-t1)/1e9:.2f} GB/s end to end, host buffers)")
