"""Synthetic corpora of BASELINE.json's configs.

Host generators live in csrc/host/synth.c; the order-2 Markov model is the
committed table data/markov2.npz (made by data/make_markov_model.py from the
reference's sample text).  Nothing here touches /root/reference."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
_MODEL = None
_LIB = None


def lib() -> C.CDLL:
    """libdfl_synth.so: the host generators (csrc/host/synth.c), a library of their own so that generating a
    corpus -- the reference arm of bench.py does -- never loads the product library."""
    global _LIB
    if _LIB is None:
        path = os.path.join(PKG, "libdfl_synth.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-C", os.path.join(PKG, "csrc"), "-s", path], check=True)
        L = C.CDLL(path)
        vp = C.c_void_p
        L.dfl_synth_random.argtypes = [vp, C.c_size_t, C.c_uint64]
        L.dfl_synth_markov.argtypes = [vp, C.c_size_t, C.c_uint64, vp, vp, vp, C.c_uint32]
        L.dfl_synth_log.argtypes = [vp, C.c_size_t, C.c_uint64, vp, C.c_size_t]
        L.dfl_synth_png.argtypes = [vp, C.c_size_t, C.c_uint64]
        _LIB = L
    return _LIB


def _model():
    global _MODEL
    if _MODEL is None:
        z = np.load(os.path.join(PKG, "data", "markov2.npz"))
        ctx = z["ctx"].astype(np.int64)
        cnt = z["cnt"].astype(np.int64)
        ctx_off = np.zeros(65537, dtype=np.uint32)
        np.add.at(ctx_off, ctx + 1, 1)
        ctx_off = np.cumsum(ctx_off).astype(np.uint32)
        csum = np.cumsum(cnt)
        start_sum = np.concatenate([[0], csum])[ctx_off[ctx]]
        cum = (csum - start_sum).astype(np.uint32)  # inclusive cumulative count inside each context
        _MODEL = (ctx_off, np.ascontiguousarray(z["nxt"]), np.ascontiguousarray(cum), int(z["start"][0]))
    return _MODEL


def random_bytes(n: int, seed: int) -> np.ndarray:
    out = np.empty(n, dtype=np.uint8)
    lib().dfl_synth_random(out.ctypes.data, n, seed)
    return out


def markov(n: int, seed: int) -> np.ndarray:
    """Order-2 Markov 'English-like' text (configs 1 and 2: seeds 1 and 2)."""
    ctx_off, nxt, cum, start = _model()
    out = np.empty(n, dtype=np.uint8)
    lib().dfl_synth_markov(out.ctypes.data, n, seed, ctx_off.ctypes.data, nxt.ctypes.data, cum.ctypes.data, start)
    return out


def log(n: int, seed: int = 3) -> np.ndarray:
    """Highly repetitive log lines (config 3)."""
    text = markov(64 * 300 + 300, seed)
    out = np.empty(n, dtype=np.uint8)
    lib().dfl_synth_log(out.ctypes.data, n, seed, text.ctypes.data, text.size)
    return out


def png_rows(n: int, seed: int = 6) -> np.ndarray:
    """PNG-style filtered RGB scanlines (config 5)."""
    out = np.empty(n, dtype=np.uint8)
    lib().dfl_synth_png(out.ctypes.data, n, seed)
    return out


def mixed(lo: int, hi: int, stripe: int = 32 << 20, seed_text: int = 5, seed_png: int = 6) -> np.ndarray:
    """Bytes [lo, hi) of config 5's corpus: alternating `stripe`-byte stripes of
    Markov text and filtered scanlines; every stripe is seeded by its index so any
    rank can generate its own range."""
    out = np.empty(hi - lo, dtype=np.uint8)
    pos = lo
    while pos < hi:
        k = pos // stripe
        s0 = k * stripe
        end = min(hi, s0 + stripe)
        gen = markov if k % 2 == 0 else png_rows
        seed = (seed_text if k % 2 == 0 else seed_png) * 1000003 + k
        full = gen(end - s0, seed)  # generated from the stripe start so content is position independent
        out[pos - lo:end - lo] = full[pos - s0:]
        pos = end
    return out
