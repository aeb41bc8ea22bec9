"""Small multi-case compress run for compute-sanitizer (dev tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, zlib
import deflate_compression_b200 as D
from deflate_compression_b200 import synth
comp = D.Compressor()
for data in (synth.markov(100001, 1), synth.random_bytes(40000, 4), synth.log(70000, 3), np.zeros(33000, dtype=np.uint8), synth.markov(5, 1), synth.markov(32768, 2)):
    z = comp.compress(data)
    assert zlib.decompress(z) == data.tobytes()
comp2 = D.Compressor(seg_bytes=1000, window=300, chain_limit=4)
z = comp2.compress(synth.markov(10000, 3)); assert zlib.decompress(z) == synth.markov(10000, 3).tobytes()
print("sanitize run ok")
