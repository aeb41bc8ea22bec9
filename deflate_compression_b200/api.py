"""ctypes mirror of the C interface of libdeflate_b200.so.

Two layers, as in the C code:
  * the reference's own API (include/deflate_ext.h; reference:
    src/include/deflate_ext.h:10-17): deflate_compress(fd_in, fd_out, fd_stats,
    sw, ops), deflate_decompress(...)
  * the thin CUDA C-ABI beneath it (include/deflate_b200.h): Compressor

No CPU fallback: if the CUDA library is missing or no B200 is visible, the calls
raise DflError."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
_LIB = os.environ.get("DFL_B200_LIB") or os.path.join(PKG, "libdeflate_b200.so")
_lib = None


class DflError(RuntimeError):
    def __init__(self, code: int, what: str = ""):
        self.code = code
        msg = lib().dfl_b200_strerror(code).decode() if _lib is not None else ""
        super().__init__(f"{what}: error {code} ({msg})")


class DflParams(C.Structure):
    _fields_ = [
        ("window", C.c_int32),
        ("seg_bytes", C.c_int32),
        ("chain_limit", C.c_int32),
        ("use_dict", C.c_int32),
        ("block_mode", C.c_int32),
    ]


class DflSegInfo(C.Structure):
    _fields_ = [
        ("btype", C.c_int32),
        ("ntokens", C.c_int32),
        ("block_bits", C.c_int64),
        ("out_bytes", C.c_int64),
        ("out_offset", C.c_int64),
        ("tree_bits", C.c_int64),
        ("ll_bits", C.c_int64),
        ("d_bits", C.c_int64),
        ("nblocks", C.c_int64),
    ]


class StringLen(C.Structure):
    _fields_ = [("str", C.POINTER(C.c_ubyte)), ("len", C.c_size_t)]


def lib_path() -> str:
    return _LIB


def build(force: bool = False) -> None:
    """Compile libdeflate_b200.so + zencode in-tree (nvcc -gencode arch=compute_100a,code=sm_100a)."""
    del force  # make decides what is stale
    subprocess.run(["make", "-C", os.path.join(PKG, "csrc"), "-s"], check=True)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            raise OSError(f"{_LIB} is missing: run deflate_compression_b200.build() (there is no CPU fallback)")
        L = C.CDLL(_LIB)
        vp = C.c_void_p
        L.dfl_b200_default_params.argtypes = [C.POINTER(DflParams)]
        L.dfl_b200_strerror.restype = C.c_char_p
        L.dfl_b200_strerror.argtypes = [C.c_int]
        L.dfl_b200_create.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(DflParams)]
        L.dfl_b200_destroy.argtypes = [vp]
        L.dfl_b200_set_params.argtypes = [vp, C.POINTER(DflParams)]
        L.dfl_b200_bound.restype = C.c_size_t
        L.dfl_b200_bound.argtypes = [C.POINTER(DflParams), C.c_size_t]
        L.dfl_b200_deflate_device.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_int, vp, C.c_size_t,
                                              C.POINTER(C.c_size_t), C.POINTER(C.c_uint32), vp]
        L.dfl_b200_deflate_enqueue.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_int, vp, C.c_size_t, vp]
        L.dfl_b200_deflate_finish.argtypes = [vp, C.POINTER(C.c_size_t), C.POINTER(C.c_uint32), vp]
        L.dfl_b200_compress_host.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, C.POINTER(C.c_size_t)]
        L.dfl_b200_keep_tokens.argtypes = [vp, C.c_int]
        L.dfl_b200_set_parse_mode.argtypes = [vp, C.c_int]
        L.dfl_b200_seg_index.restype = C.c_long
        L.dfl_b200_seg_index.argtypes = [vp, vp, C.c_long]
        L.dfl_b200_seg_index_device.restype = vp
        L.dfl_b200_seg_index_device.argtypes = [vp]
        L.dfl_b200_inflate_device.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, C.c_uint32, C.c_size_t, vp, C.c_size_t,
                                              C.POINTER(C.c_uint32), vp]
        L.dfl_b200_inflate_rounds.argtypes = [vp]
        L.dfl_b200_png_filter_device.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_int, C.c_int, vp, vp]
        L.dfl_b200_png_bound.restype = C.c_size_t
        L.dfl_b200_png_bound.argtypes = [C.POINTER(DflParams), C.c_uint32, C.c_uint32, C.c_uint32]
        L.dfl_b200_png_encode_host.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, vp, C.c_size_t,
                                               C.POINTER(C.c_size_t)]
        L.dfl_b200_decompress_host.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, C.c_uint32, vp, C.c_size_t]
        L.dfl_b200_num_segments.restype = C.c_long
        L.dfl_b200_num_segments.argtypes = [vp]
        L.dfl_b200_seg_info.argtypes = [vp, C.c_long, C.POINTER(DflSegInfo)]
        L.dfl_b200_tokens.restype = C.c_long
        L.dfl_b200_tokens.argtypes = [vp, C.c_long, vp, C.c_long]
        L.dfl_b200_set_match_mode.argtypes = [vp, C.c_int]
        L.dfl_b200_compress_fd.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
        L.dfl_b200_set_fd_piece.argtypes = [vp, C.c_size_t]
        L.dfl_b200_compress_host_multi.argtypes = [C.POINTER(vp), C.c_int, vp, C.c_size_t, vp, C.c_size_t, C.POINTER(C.c_size_t)]
        L.dfl_b200_adler32_combine.restype = C.c_uint32
        L.dfl_b200_adler32_combine.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64]
        L.dfl_b200_copy_async.argtypes = [vp, vp, C.c_size_t, vp]
        L.dfl_b200_kernel_launches.restype = C.c_long
        L.dfl_b200_kernel_launches.argtypes = [vp]
        L.dfl_b200_enable_timing.argtypes = [vp, C.c_int]
        L.dfl_b200_num_kernels.restype = C.c_int
        L.dfl_b200_kernel_name.restype = C.c_char_p
        L.dfl_b200_kernel_name.argtypes = [C.c_int]
        L.dfl_b200_kernel_ms.restype = C.c_double
        L.dfl_b200_kernel_ms.argtypes = [vp, C.c_int]
        L.dfl_b200_phase_cycles.restype = C.c_uint64
        L.dfl_b200_phase_cycles.argtypes = [vp, C.c_int]
        L.dfl_b200_sort_fallbacks.restype = C.c_uint
        L.dfl_b200_sort_fallbacks.argtypes = [vp]
        L.dfl_b200_synth_random.argtypes = [vp, C.c_size_t, C.c_uint64, vp]
        # reference API (include/deflate_ext.h)
        L.deflate_compress.restype = C.c_int
        L.deflate_compress.argtypes = [C.c_int, C.c_int, C.c_int, C.c_ushort, C.c_int]
        L.deflate_decompress.restype = C.c_int
        L.deflate_decompress.argtypes = [C.POINTER(StringLen), C.POINTER(StringLen), C.c_int]
        L.spawn_deflate_compr_t.restype = vp
        L.deflate_compr_init.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_ushort]
        L.deflate_compr_deinit.argtypes = [vp]
        _lib = L
    return _lib


def _u8(data) -> np.ndarray:
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data, dtype=np.uint8)
    return np.frombuffer(bytes(data), dtype=np.uint8)


class Compressor:
    """One dfl_ctx (include/deflate_b200.h) bound to a CUDA device.  Parameters as struct dfl_params: window (the
    reference's `sw`), seg_bytes, chain_limit (0 = walk every chain to its end, as the reference does), use_dict,
    block_mode (0 = cheapest of stored / fixed / dynamic per segment, 1-3 = force one, 4 = cheapest with adaptive block
    splitting: up to 8 blocks per segment where that saves bits)."""

    def __init__(self, device: int = 0, window: int = 32768, seg_bytes: int = 32768, chain_limit: int = 0,
                 use_dict: int = 1, block_mode: int = 0):
        self.params = DflParams(window, seg_bytes, chain_limit, use_dict, block_mode)
        self._ctx = C.c_void_p()
        rc = lib().dfl_b200_create(C.byref(self._ctx), device, C.byref(self.params))
        if rc:
            self._ctx = C.c_void_p()
            raise DflError(rc, "dfl_b200_create")

    def close(self) -> None:
        if self._ctx:
            lib().dfl_b200_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def bound(self, n: int) -> int:
        return int(lib().dfl_b200_bound(C.byref(self.params), n))

    # ---- host buffers: complete zlib stream, H2D + D2H inside ----
    def compress(self, data) -> bytes:
        a = _u8(data)
        cap = self.bound(a.size) + 16
        out = np.empty(cap, dtype=np.uint8)
        n = self.compress_into(a.ctypes.data, a.size, out.ctypes.data, cap)
        return out[:n].tobytes()

    def compress_into(self, in_ptr: int, n: int, out_ptr: int, cap: int) -> int:
        out_len = C.c_size_t(0)
        rc = lib().dfl_b200_compress_host(self._ctx, in_ptr, n, out_ptr, cap, C.byref(out_len))
        if rc:
            raise DflError(rc, "dfl_b200_compress_host")
        return int(out_len.value)

    # ---- device buffers: raw deflate body ----
    def deflate_device(self, d_in: int, n: int, hist_len: int, is_final: bool, d_out: int, cap: int,
                       stream: int = 0):
        out_len = C.c_size_t(0)
        adler = C.c_uint32(0)
        rc = lib().dfl_b200_deflate_device(self._ctx, d_in, n, hist_len, int(is_final), d_out, cap,
                                           C.byref(out_len), C.byref(adler), stream)
        if rc:
            raise DflError(rc, "dfl_b200_deflate_device")
        return int(out_len.value), int(adler.value)

    def enqueue(self, d_in: int, n: int, hist_len: int, is_final: bool, d_out: int, cap: int, stream: int = 0):
        rc = lib().dfl_b200_deflate_enqueue(self._ctx, d_in, n, hist_len, int(is_final), d_out, cap, stream)
        if rc:
            raise DflError(rc, "dfl_b200_deflate_enqueue")

    def finish(self, stream: int = 0):
        out_len = C.c_size_t(0)
        adler = C.c_uint32(0)
        rc = lib().dfl_b200_deflate_finish(self._ctx, C.byref(out_len), C.byref(adler), stream)
        if rc:
            raise DflError(rc, "dfl_b200_deflate_finish")
        return int(out_len.value), int(adler.value)

    # ---- parity / stats ----
    def keep_tokens(self, enable: bool = True) -> None:
        lib().dfl_b200_keep_tokens(self._ctx, int(enable))

    # ---- segment index + GPU inflate ----
    def seg_index(self) -> np.ndarray:
        """Byte offset of every segment of the last compress call in the DEFLATE body, + the body length."""
        cnt = int(lib().dfl_b200_seg_index(self._ctx, None, 0))
        if cnt < 0:
            raise DflError(cnt, "dfl_b200_seg_index")
        off = np.zeros(cnt, dtype=np.uint64)
        rc = int(lib().dfl_b200_seg_index(self._ctx, off.ctypes.data, cnt))
        if rc < 0:
            raise DflError(rc, "dfl_b200_seg_index")
        return off

    def seg_index_device(self) -> int:
        return int(lib().dfl_b200_seg_index_device(self._ctx) or 0)

    def inflate_device(self, d_body: int, body_len: int, d_seg_off: int, nseg: int, d_out: int, n: int,
                       hist_len: int = 0, stream: int = 0) -> int:
        """Segment-parallel inflate of a body this library wrote; returns the Adler-32 of the n bytes."""
        adler = C.c_uint32(0)
        rc = lib().dfl_b200_inflate_device(self._ctx, d_body, body_len, d_seg_off, nseg, self.params.seg_bytes, hist_len,
                                           d_out, n, C.byref(adler), stream)
        if rc:
            raise DflError(rc, "dfl_b200_inflate_device")
        return int(adler.value)

    def inflate_rounds(self) -> int:
        return int(lib().dfl_b200_inflate_rounds(self._ctx))

    def decompress(self, z: bytes, seg_off: np.ndarray, n: int) -> bytes:
        """zlib stream + its segment index -> the n original bytes, inflated on the GPU (host buffers)."""
        src = np.frombuffer(z, dtype=np.uint8)
        off = np.ascontiguousarray(seg_off, dtype=np.uint64)
        out = np.empty(max(n, 1), dtype=np.uint8)
        rc = lib().dfl_b200_decompress_host(self._ctx, src.ctypes.data, src.size, off.ctypes.data, off.size - 1,
                                            self.params.seg_bytes, out.ctypes.data, n)
        if rc:
            raise DflError(rc, "dfl_b200_decompress_host")
        return out[:n].tobytes()

    # ---- PNG scanline filter + container ----
    def png_filter_device(self, d_pix: int, row_bytes: int, rows: int, bpp: int, d_out: int, mode: int = 0,
                          stream: int = 0) -> None:
        rc = lib().dfl_b200_png_filter_device(self._ctx, d_pix, row_bytes, rows, bpp, mode, d_out, stream)
        if rc:
            raise DflError(rc, "dfl_b200_png_filter_device")

    def png_encode(self, img: np.ndarray, filter_mode: int = 0) -> bytes:
        """(h, w) or (h, w, c) uint8 pixels, c in 1..4 -> PNG file bytes (filter + deflate on the GPU)."""
        a = np.ascontiguousarray(img, dtype=np.uint8)
        h, w = a.shape[0], a.shape[1]
        c = a.shape[2] if a.ndim == 3 else 1
        cap = int(lib().dfl_b200_png_bound(C.byref(self.params), w, h, c))
        out = np.empty(cap, dtype=np.uint8)
        n = C.c_size_t(0)
        rc = lib().dfl_b200_png_encode_host(self._ctx, a.ctypes.data, w, h, c, filter_mode, out.ctypes.data, cap, C.byref(n))
        if rc:
            raise DflError(rc, "dfl_b200_png_encode_host")
        return out[:n.value].tobytes()

    def set_parse_mode(self, mode: int) -> None:
        """0 = by segment count, 1 = warp per segment, 2 = CTA per segment (same tokens)."""
        rc = lib().dfl_b200_set_parse_mode(self._ctx, int(mode))
        if rc:
            raise DflError(rc, "dfl_b200_set_parse_mode")

    def compress_fd(self, fd_in: int, fd_out: int, piece_bytes: int = 0):
        """fd -> fd (the body of deflate_compress): returns (bytes read, bytes written)."""
        lib().dfl_b200_set_fd_piece(self._ctx, int(piece_bytes))
        nin, nout = C.c_ulonglong(0), C.c_ulonglong(0)
        rc = lib().dfl_b200_compress_fd(self._ctx, int(fd_in), int(fd_out), C.byref(nin), C.byref(nout))
        if rc:
            raise DflError(rc, "dfl_b200_compress_fd")
        return int(nin.value), int(nout.value)

    def set_match_mode(self, mode: int) -> None:
        """0 = k_match_run where the geometry is the default one, 1 = always k_match (same tokens)."""
        rc = lib().dfl_b200_set_match_mode(self._ctx, int(mode))
        if rc:
            raise DflError(rc, "dfl_b200_set_match_mode")

    def num_segments(self) -> int:
        return int(lib().dfl_b200_num_segments(self._ctx))

    def seg_info(self, seg: int) -> DflSegInfo:
        info = DflSegInfo()
        rc = lib().dfl_b200_seg_info(self._ctx, seg, C.byref(info))
        if rc:
            raise DflError(rc, "dfl_b200_seg_info")
        return info

    def tokens(self, seg: int) -> np.ndarray:
        nt = lib().dfl_b200_tokens(self._ctx, seg, None, 0)
        if nt < 0:
            raise DflError(int(nt), "dfl_b200_tokens")
        out = np.empty(max(nt, 1), dtype=np.uint32)
        lib().dfl_b200_tokens(self._ctx, seg, out.ctypes.data, nt)
        return out[:nt]

    # ---- measurement ----
    def enable_timing(self, enable: bool = True) -> None:
        lib().dfl_b200_enable_timing(self._ctx, int(enable))

    def kernel_ms(self) -> dict:
        L = lib()
        return {L.dfl_b200_kernel_name(k).decode(): float(L.dfl_b200_kernel_ms(self._ctx, k))
                for k in range(L.dfl_b200_num_kernels())}

    def phase_cycles(self) -> dict:
        names = ("stage_in", "build", "search", "total", "sort1", "park", "sort2", "offsets", "s1_count", "s1_scan", "s1_scatter", "d11",
                 "sx_iter", "sx_walk", "sx_fin", "sx_done", "sx_cand", "sx_heads", "sx_hops", "sx_seek", "plain_segments", "warp_end_avg", "warp_end_max")
        d = {nm: int(lib().dfl_b200_phase_cycles(self._ctx, k)) for k, nm in enumerate(names)}
        if not d["total"]:  # k_match_run accumulates the three phases only
            d["total"] = d["stage_in"] + d["build"] + d["search"]
        return d

    def sort_fallbacks(self) -> int:
        return int(lib().dfl_b200_sort_fallbacks(self._ctx))

    def kernel_launches(self) -> int:
        return int(lib().dfl_b200_kernel_launches(self._ctx))


def compress_multi(comps, in_ptr: int, n: int, out_ptr: int, cap: int) -> int:
    """dfl_b200_compress_host_multi: comps = one Compressor per GPU (same parameters); host pointers; -> stream length."""
    arr = (C.c_void_p * len(comps))(*[c._ctx.value for c in comps])
    out_len = C.c_size_t(0)
    rc = lib().dfl_b200_compress_host_multi(arr, len(comps), in_ptr, n, out_ptr, cap, C.byref(out_len))
    if rc:
        raise DflError(rc, "dfl_b200_compress_host_multi")
    return int(out_len.value)


def deflate_compress(fd_in: int, fd_out: int, fd_stats: int = -1, sw: int = 32768, ops: int = 0) -> int:
    """The reference entry point (src/include/deflate_ext.h:17): 0 or a negative error."""
    return int(lib().deflate_compress(fd_in, fd_out, fd_stats, sw, ops))


def deflate_decompress(comp, ops: int = 0) -> bytes:
    """The reference entry point (src/include/deflate_ext.h:16) over a bytes object."""
    a = _u8(comp)
    src = StringLen(a.ctypes.data_as(C.POINTER(C.c_ubyte)), a.size)
    dst = StringLen()
    rc = lib().deflate_decompress(C.byref(dst), C.byref(src), ops)
    if rc < 0:
        raise DflError(rc, "deflate_decompress")
    try:
        return C.string_at(dst.str, dst.len) if dst.len else b""
    finally:
        if dst.str:
            C.CDLL(None).free(dst.str)
