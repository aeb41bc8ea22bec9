import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def built():
    """Build the product library and the oracle once per session."""
    import deflate_compression_b200 as D
    from oracle import pyoracle as O

    D.build()
    O.build()
    return D


def make_input(kind: str, seed: int, off: int, n: int) -> np.ndarray:
    """Regenerate a golden input from its spec (mirrors tests/golden/make_golden.py)."""
    from deflate_compression_b200 import synth

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


def golden_cases():
    path = os.path.join(ROOT, "tests", "golden", "ref_tokens.npz")
    z = np.load(path)
    out = []
    for entry in z["names"]:
        name, kind = str(entry).split(":")
        seed, off, n = (int(v) for v in z[f"{name}__spec"])
        out.append((name, kind, seed, off, n, z[f"{name}__tokens"], z[f"{name}__last"], z[f"{name}__sha"]))
    return out


def striped_input(n: int, seed: int = 11) -> np.ndarray:
    """Data whose statistics change every few KiB (text / random bytes / a numeric table, stripes of 2-12 KiB): the
    case adaptive block splitting (block_mode 4) exists for."""
    rng = np.random.default_rng(seed)
    parts, tot, i = [], 0, 0
    while tot < n:
        m = int(rng.integers(2000, 12000))
        if i % 3 == 0:
            parts.append(make_input("markov", seed + i, 0, m))
        elif i % 3 == 1:
            parts.append(rng.integers(0, 256, m, dtype=np.uint8))
        else:
            parts.append(np.frombuffer((b"%d,%d;" % (i, i * 7)) * (m // 6 + 1), dtype=np.uint8)[:m].copy())
        tot += m
        i += 1
    return np.concatenate(parts)[:n]
