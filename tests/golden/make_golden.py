"""Generate tests/golden/ref_tokens.npz: token streams produced by the REAL
reference compressor (oracle/_ref/libref.so, built from /root/reference by
oracle/Makefile) on deterministic synthetic inputs.  Run in the build container:

    python tests/golden/make_golden.py

The inputs are regenerated at test time from (kind, seed, offset, n) with
deflate_compression_b200.synth, so only the reference's outputs are stored.
Each record of the reference's fd_stats stream (struct compress_stats,
src/include/deflate_ext.h:19-31) is reduced to one uint32 token: the literal
byte, or dist << 16 | len.  The last record's cost fields
(bytes, tree_bits, ll_bits, d_bits) are kept as well.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from deflate_compression_b200 import synth  # noqa: E402
from oracle import pyoracle as O  # noqa: E402

CASES = [
    # name, kind, seed, offset, n
    ("markov1_head32k", "markov", 1, 0, 32768),
    ("markov1_mid32k", "markov", 1, 500000, 32768),
    ("markov2_3000", "markov", 2, 0, 3000),
    ("markov2_32767", "markov", 2, 1000, 32767),
    ("markov2_32766", "markov", 2, 2000, 32766),
    ("markov1_100", "markov", 1, 777, 100),
    ("markov1_4", "markov", 1, 0, 4),
    ("markov1_3", "markov", 1, 0, 3),
    ("random_4096", "random", 4, 0, 4096),
    ("random_32768", "random", 4, 8192, 32768),
    ("log_32768", "log", 3, 0, 32768),
    ("log_mid_20000", "log", 3, 123456, 20000),
    ("png_32768", "png", 6, 0, 32768),
    ("zeros_32768", "zeros", 0, 0, 32768),
    ("a40", "const_a", 0, 0, 40),
    ("period7_5000", "period7", 0, 0, 5000),
]


def make_input(kind: str, seed: int, off: int, n: int) -> np.ndarray:
    if kind == "markov":
        return synth.markov(off + n, seed)[off:]
    if kind == "random":
        return synth.random_bytes(off + n, seed)[off:]
    if kind == "log":
        return synth.log(off + n, seed)[off:]
    if kind == "png":
        return synth.png_rows(off + n, seed)[off:]
    if kind == "zeros":
        return np.zeros(n, dtype=np.uint8)
    if kind == "const_a":
        return np.full(n, ord("a"), dtype=np.uint8)
    if kind == "period7":
        return np.resize(np.frombuffer(b"abcdefg", dtype=np.uint8), n).copy()
    raise ValueError(kind)


def main() -> None:
    assert O.have_ref(), "needs oracle/_ref/libref.so (build container only)"
    out = {}
    names = []
    for name, kind, seed, off, n in CASES:
        data = make_input(kind, seed, off, n)
        recs = O.ref_tokens(data)
        toks = O.ref_records_to_tokens(recs)
        assert O.tokens_expand(toks) == data.tobytes(), name  # check_lld method
        out[f"{name}__tokens"] = toks
        out[f"{name}__last"] = recs[-1, :4].astype(np.int32)
        out[f"{name}__sha"] = np.frombuffer(hashlib.sha256(data.tobytes()).digest(), dtype=np.uint8)
        out[f"{name}__spec"] = np.array([seed, off, n], dtype=np.int64)
        names.append(f"{name}:{kind}")
        print(f"{name}: {n} B -> {toks.size} tokens, last record {recs[-1].tolist()}")
    out["names"] = np.array(names)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_tokens.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
