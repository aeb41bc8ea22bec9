"""ctypes view of the CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  It wraps

  * oracle/liboracle.so      -- the plain-C restatement (oracle/deflate_oracle.c)
  * oracle/_ref/libref.so    -- the UNMODIFIED reference LZ77 path compiled from
                                /root/reference by oracle/Makefile (optional at
                                run time: the GPU box only has the prebuilt file)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_PATH = os.path.join(HERE, "_ref", "libref.so")


class OrcParams(C.Structure):
    _fields_ = [
        ("window", C.c_int32),
        ("seg_bytes", C.c_int32),
        ("chain_limit", C.c_int32),
        ("use_dict", C.c_int32),
        ("block_mode", C.c_int32),
    ]


class OrcSegStats(C.Structure):
    _fields_ = [
        ("btype", C.c_int32),
        ("ntokens", C.c_int32),
        ("block_bits", C.c_int64),
        ("out_bytes", C.c_int64),
        ("stored_bits", C.c_int64),
        ("fixed_bits", C.c_int64),
        ("dynamic_bits", C.c_int64),
        ("tree_bits", C.c_int64),
        ("ll_bits", C.c_int64),
        ("d_bits", C.c_int64),
        ("nblocks", C.c_int64),
    ]


def build(force: bool = False) -> None:
    """Compile the oracle (and, when /root/reference exists, oracle/_ref)."""
    del force  # make decides what is stale
    subprocess.run(["make", "-C", HERE, "-s"], check=True)


_lib = None
_ref = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        u8p = C.c_void_p
        L.orc_default_params.argtypes = [C.POINTER(OrcParams)]
        L.orc_hash3.restype = C.c_uint32
        L.orc_hash3.argtypes = [u8p]
        L.orc_lz77_segment.restype = C.c_long
        L.orc_lz77_segment.argtypes = [u8p, C.c_long, C.c_long, C.c_long, C.POINTER(OrcParams), u8p]
        L.orc_huff_lengths.argtypes = [u8p, C.c_int, C.c_int, u8p]
        L.orc_canonical_codes.argtypes = [u8p, C.c_int, u8p]
        L.orc_deflate_body.restype = C.c_long
        L.orc_deflate_body.argtypes = [u8p, C.c_long, C.c_long, C.POINTER(OrcParams), C.c_int, u8p,
                                       C.c_long, C.c_void_p, C.c_long]
        L.orc_num_segments.restype = C.c_long
        L.orc_num_segments.argtypes = [C.c_long, C.POINTER(OrcParams)]
        L.orc_bound.restype = C.c_long
        L.orc_bound.argtypes = [C.c_long, C.POINTER(OrcParams)]
        L.orc_zlib_compress.restype = C.c_long
        L.orc_zlib_compress.argtypes = [u8p, C.c_long, C.POINTER(OrcParams), u8p, C.c_long]
        L.orc_png_filter.argtypes = [u8p, C.c_long, C.c_long, C.c_int, C.c_int, u8p]
        L.orc_crc32.restype = C.c_uint32
        L.orc_crc32.argtypes = [u8p, C.c_size_t]
        L.orc_adler32.restype = C.c_uint32
        L.orc_adler32.argtypes = [C.c_uint32, u8p, C.c_size_t]
        L.orc_adler32_combine.restype = C.c_uint32
        L.orc_adler32_combine.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64]
        L.orc_inflate_raw.restype = C.c_long
        L.orc_inflate_raw.argtypes = [u8p, C.c_long, u8p, C.c_long, C.POINTER(C.c_long)]
        L.orc_zlib_decompress.restype = C.c_long
        L.orc_zlib_decompress.argtypes = [u8p, C.c_long, u8p, C.c_long]
        L.orc_tokens_expand.restype = C.c_long
        L.orc_tokens_expand.argtypes = [u8p, C.c_long, u8p, C.c_long, u8p, C.c_long]
        _lib = L
    return _lib


def have_ref() -> bool:
    build()
    return os.path.exists(REF_PATH)


def ref() -> C.CDLL:
    global _ref
    if _ref is None:
        build()
        R = C.CDLL(REF_PATH)
        R.ref_tokens.restype = C.c_long
        R.ref_tokens.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_long]
        R.ref_compress_slices.restype = C.c_long
        R.ref_compress_slices.argtypes = [C.c_void_p, C.c_long]
        _ref = R
    return _ref


def params(window=32768, seg_bytes=32768, chain_limit=0, use_dict=1, block_mode=0) -> OrcParams:
    return OrcParams(window, seg_bytes, chain_limit, use_dict, block_mode)


def _u8(data) -> np.ndarray:
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data, dtype=np.uint8)
    return np.frombuffer(bytes(data), dtype=np.uint8)


def lz77_segment(data, hist_lo: int, seg_lo: int, seg_hi: int, prm: OrcParams) -> np.ndarray:
    """Tokens (uint32: literal byte, or dist<<16|len) of data[seg_lo:seg_hi]."""
    a = _u8(data)
    toks = np.empty(max(seg_hi - seg_lo, 1), dtype=np.uint32)
    n = lib().orc_lz77_segment(a.ctypes.data, hist_lo, seg_lo, seg_hi, C.byref(prm), toks.ctypes.data)
    if n < 0:
        raise RuntimeError(f"orc_lz77_segment failed: {n}")
    return toks[:n].copy()


def deflate_body(data, data_lo: int, data_hi: int, prm: OrcParams, is_final: bool = True):
    """(raw deflate bytes, per-segment stats list) for data[data_lo:data_hi]."""
    a = _u8(data)
    n = data_hi - data_lo
    cap = lib().orc_bound(n, C.byref(prm))
    out = np.empty(cap, dtype=np.uint8)
    nseg = lib().orc_num_segments(n, C.byref(prm))
    stats = (OrcSegStats * max(nseg, 1))()
    r = lib().orc_deflate_body(a.ctypes.data, data_lo, data_hi, C.byref(prm), int(is_final),
                               out.ctypes.data, cap, C.cast(stats, C.c_void_p), nseg)
    if r < 0:
        raise RuntimeError(f"orc_deflate_body failed: {r}")
    return out[:r].tobytes(), list(stats)[:nseg]


def zlib_compress(data, prm: OrcParams) -> bytes:
    a = _u8(data)
    cap = lib().orc_bound(a.size, C.byref(prm)) + 16
    out = np.empty(cap, dtype=np.uint8)
    r = lib().orc_zlib_compress(a.ctypes.data, a.size, C.byref(prm), out.ctypes.data, cap)
    if r < 0:
        raise RuntimeError(f"orc_zlib_compress failed: {r}")
    return out[:r].tobytes()


def zlib_decompress(comp, cap: int) -> bytes:
    a = _u8(comp)
    out = np.empty(max(cap, 1), dtype=np.uint8)
    r = lib().orc_zlib_decompress(a.ctypes.data, a.size, out.ctypes.data, cap)
    if r < 0:
        raise RuntimeError(f"orc_zlib_decompress failed: {r}")
    return out[:r].tobytes()


def inflate_raw(comp, cap: int) -> bytes:
    a = _u8(comp)
    out = np.empty(max(cap, 1), dtype=np.uint8)
    used = C.c_long(0)
    r = lib().orc_inflate_raw(a.ctypes.data, a.size, out.ctypes.data, cap, C.byref(used))
    if r < 0:
        raise RuntimeError(f"orc_inflate_raw failed: {r}")
    return out[:r].tobytes()


def adler32(data, start: int = 1) -> int:
    a = _u8(data)
    return int(lib().orc_adler32(start, a.ctypes.data, a.size))


def adler32_combine(a1: int, a2: int, len2: int) -> int:
    return int(lib().orc_adler32_combine(a1, a2, len2))


def huff_lengths(freq, limit: int) -> np.ndarray:
    f = np.ascontiguousarray(freq, dtype=np.uint32)
    out = np.zeros(f.size, dtype=np.uint8)
    lib().orc_huff_lengths(f.ctypes.data, f.size, limit, out.ctypes.data)
    return out


def canonical_codes(lens) -> np.ndarray:
    l = np.ascontiguousarray(lens, dtype=np.uint8)
    out = np.zeros(l.size, dtype=np.uint16)
    lib().orc_canonical_codes(l.ctypes.data, l.size, out.ctypes.data)
    return out


def tokens_expand(tokens, hist=b"", cap: int | None = None) -> bytes:
    t = np.ascontiguousarray(tokens, dtype=np.uint32)
    h = _u8(hist)
    if cap is None:
        cap = int(np.where((t >> 16) == 0, 1, t & 0xFFFF).sum()) if t.size else 0
    out = np.empty(max(cap, 1), dtype=np.uint8)
    r = lib().orc_tokens_expand(t.ctypes.data, t.size, h.ctypes.data, h.size, out.ctypes.data, cap)
    if r < 0:
        raise RuntimeError(f"orc_tokens_expand failed: {r}")
    return out[:r].tobytes()


def ref_tokens(data) -> np.ndarray:
    """compress_stats records (n,6 int32) from the real reference, 3 <= len <= 32768."""
    a = _u8(data)
    recs = np.zeros((a.size + 8, 6), dtype=np.int32)
    n = ref().ref_tokens(a.ctypes.data, a.size, recs.ctypes.data, a.size + 8)
    if n < 0:
        raise RuntimeError(f"ref_tokens failed: {n}")
    return recs[:n].copy()


def ref_records_to_tokens(recs: np.ndarray) -> np.ndarray:
    """Appendix B of SURVEY.md: d == 0 -> literal ll&0xFF, else (ll, d) match."""
    ll = recs[:, 4].astype(np.uint32)
    d = recs[:, 5].astype(np.uint32)
    return np.where(d == 0, ll & 0xFF, (d << 16) | ll).astype(np.uint32)


def png_filter(img, mode: int = 0) -> np.ndarray:
    """Filtered scanlines ([type | row] per row) of an (h, w, c) or (h, row_bytes) uint8 image."""
    a = np.ascontiguousarray(img, dtype=np.uint8)
    h = a.shape[0]
    bpp = a.shape[2] if a.ndim == 3 else 1
    row = a.size // h if h else 0
    out = np.empty(h * (row + 1), dtype=np.uint8)
    lib().orc_png_filter(a.ctypes.data, row, h, bpp, mode, out.ctypes.data)
    return out


def crc32(data) -> int:
    a = _u8(data)
    return int(lib().orc_crc32(a.ctypes.data, a.size))
