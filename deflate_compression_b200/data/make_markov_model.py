"""Derive the order-2 character Markov model used by the synthetic text corpora
 from the reference's only sample text,
test_files/original/bee_movie_script.txt.  Run in the build container (needs
/root/reference); the output markov2.npz (trigram counts, ~60 KB) is committed so
that nothing reads /root/reference at run time.

    python deflate_compression_b200/data/make_markov_model.py
"""
import os

import numpy as np

SRC = "/root/reference/test_files/original/bee_movie_script.txt"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "markov2.npz")


def main() -> None:
    t = np.frombuffer(open(SRC, "rb").read(), dtype=np.uint8).astype(np.uint32)
    key = (t[:-2] << 16) | (t[1:-1] << 8) | t[2:]
    uniq, cnt = np.unique(key, return_counts=True)
    np.savez_compressed(
        OUT,
        ctx=(uniq >> 8).astype(np.uint16),       # c0 << 8 | c1
        nxt=(uniq & 0xFF).astype(np.uint8),
        cnt=cnt.astype(np.uint32),
        start=np.array([(int(t[0]) << 8) | int(t[1])], dtype=np.uint32),
        src_len=np.array([t.size], dtype=np.uint32),
    )
    print(f"{OUT}: {uniq.size} trigrams from {t.size} bytes")


if __name__ == "__main__":
    main()
