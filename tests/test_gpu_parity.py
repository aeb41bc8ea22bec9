"""GPU parity tests (-m gpu): every call goes through the C-ABI of
libdeflate_b200.so; the CPU oracle is only the checker.

Bar: bit-exact.  (a) the per-segment token stream equals the oracle's and, for
no-dictionary 32 KiB segments, the real reference's (golden fixtures made from
oracle/_ref); (b) the emitted stream equals the oracle's byte for byte;
(c) every stream inflates to the input through zlib, the oracle's repaired
deflate_decompress and the product's deflate_decompress."""
import hashlib
import os
import zlib

import numpy as np
import pytest

from conftest import golden_cases, make_input

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def D(built):
    return built


@pytest.fixture(scope="module")
def O(built):
    from oracle import pyoracle

    return pyoracle


def gpu_tokens(comp):
    return [comp.tokens(s).copy() for s in range(comp.num_segments())]


def check_stream(D, O, data: np.ndarray, **prm):
    """Compress on the GPU with token keeping; compare tokens + bytes with the oracle."""
    comp = D.Compressor(**prm)
    comp.keep_tokens(True)
    z = comp.compress(data)
    op = O.params(**prm)
    zo = O.zlib_compress(data, op)
    n = data.size
    seg = prm.get("seg_bytes", 32768)
    nseg = (n + seg - 1) // seg
    assert comp.num_segments() == nseg
    for s in range(nseg):
        lo, hi = s * seg, min(n, (s + 1) * seg)
        ot = O.lz77_segment(data, 0, lo, hi, op)
        gt = comp.tokens(s)
        assert gt.size == ot.size and np.array_equal(gt, ot), f"token mismatch in segment {s}"
    assert z == zo, "stream differs from the oracle's"
    raw = data.tobytes()
    assert zlib.decompress(z) == raw
    assert O.zlib_decompress(z, n) == raw
    assert D.deflate_decompress(z) == raw
    comp.close()
    return z


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c[0])
def test_golden_reference_tokens(D, O, case):
    """Token stream of a no-dictionary segment == the real reference's (fixture)."""
    name, kind, seed, off, n, ref_toks, last, sha = case
    data = make_input(kind, seed, off, n)
    assert hashlib.sha256(data.tobytes()).digest() == sha.tobytes()
    comp = D.Compressor(use_dict=0)
    comp.keep_tokens(True)
    z = comp.compress(data)
    assert comp.num_segments() == 1
    gt = comp.tokens(0)
    assert gt.size == ref_toks.size and np.array_equal(gt, ref_toks)
    assert zlib.decompress(z) == data.tobytes()
    comp.close()


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 31, 32, 33, 257, 258, 259, 260, 1000, 32767, 32768, 32769, 65535,
                               65536, 65537, 100000])
def test_sizes_text(D, O, n):
    check_stream(D, O, make_input("markov", 1, 12345, n))


def test_empty(D, O):
    comp = D.Compressor()
    z = comp.compress(b"")
    assert z == O.zlib_compress(b"", O.params())
    assert zlib.decompress(z) == b""


@pytest.mark.parametrize("kind,seed", [("markov", 2), ("random", 4), ("log", 3), ("png", 6), ("zeros", 0),
                                        ("period7", 0), ("const_a", 0)])
def test_kinds_multi_segment(D, O, kind, seed):
    check_stream(D, O, make_input(kind, seed, 0, 200000))


@pytest.mark.parametrize("prm", [
    dict(seg_bytes=4096), dict(seg_bytes=4096, use_dict=0), dict(seg_bytes=1000), dict(seg_bytes=32768, use_dict=0),
    dict(window=1024), dict(window=4096, seg_bytes=8192), dict(window=1), dict(window=32767),
    dict(chain_limit=1), dict(chain_limit=8), dict(chain_limit=128, seg_bytes=16384),
    dict(block_mode=1), dict(block_mode=2), dict(block_mode=3),
], ids=str)
def test_params(D, O, prm):
    for kind, seed in (("markov", 1), ("log", 3), ("random", 4)):
        check_stream(D, O, make_input(kind, seed, 0, 150001), **prm)


def test_mixed_blocks(D, O):
    """Text, random and zero stripes: stored, dynamic (and possibly fixed) blocks in one stream."""
    parts = [make_input("markov", 1, 0, 40000), make_input("random", 4, 0, 100000), make_input("zeros", 0, 0, 70000),
             make_input("markov", 2, 0, 20), make_input("random", 9, 0, 40)]
    data = np.concatenate(parts)
    comp = D.Compressor()
    comp.keep_tokens(True)
    z = comp.compress(data)
    types = {comp.seg_info(s).btype for s in range(comp.num_segments())}
    assert 0 in types and 2 in types
    assert z == O.zlib_compress(data, O.params())
    assert zlib.decompress(z) == data.tobytes()


def test_tiny_inputs_all_lengths(D, O):
    base = make_input("markov", 1, 0, 64)
    comp = D.Compressor()
    for n in range(0, 40):
        z = comp.compress(base[:n])
        assert z == O.zlib_compress(base[:n], O.params())
        assert zlib.decompress(z) == base[:n].tobytes()


def test_deep_huffman_tree(D, O):
    """Fibonacci-weighted literals force code lengths past 15: the length limiter must agree with the oracle."""
    fib = [1, 1]
    while len(fib) < 22:
        fib.append(fib[-1] + fib[-2])
    rng = np.random.default_rng(7)
    sym = np.concatenate([np.full(f, i * 7 + 3, dtype=np.uint8) for i, f in enumerate(fib)])
    rng.shuffle(sym)
    data = sym[:32768]
    check_stream(D, O, data, block_mode=3)


def test_fd_api_roundtrip(D, O, tmp_path):
    """The reference entry point: deflate_compress(fd_in, fd_out, fd_stats, sw, ops)."""
    data = make_input("markov", 2, 0, 300000)
    src, dst, st = tmp_path / "in", tmp_path / "out.z", tmp_path / "stats"
    src.write_bytes(data.tobytes())
    fi = os.open(src, os.O_RDONLY)
    fo = os.open(dst, os.O_WRONLY | os.O_CREAT)
    fs = os.open(st, os.O_WRONLY | os.O_CREAT)
    try:
        assert D.deflate_compress(fi, fo, fs, 32768, 0) == 0
    finally:
        os.close(fi), os.close(fo), os.close(fs)
    z = dst.read_bytes()
    assert z == O.zlib_compress(data, O.params())
    assert D.deflate_decompress(z) == data.tobytes()
    # fd_stats: expanding the (ll, d) records reproduces the input (tests/check_lld.c method)
    recs = np.frombuffer(st.read_bytes(), dtype=np.int32).reshape(-1, 6)
    toks = O.ref_records_to_tokens(recs)
    assert recs[0, 0] == 1
    out = bytearray()
    for ll, d in zip(recs[:, 4], recs[:, 5]):
        if d == 0:
            out.append(int(ll) & 0xFF)
        else:
            for _ in range(int(ll)):
                out.append(out[-int(d)])
    assert bytes(out) == data.tobytes()
    assert toks.size == recs.shape[0]
    # cost fields: every token carries its block's split (tree / lit-len / distance bits), == the oracle's
    body, stats = O.deflate_body(data, 0, data.size, O.params())
    k = 0
    for s_, stt in enumerate(stats):
        blk = recs[k:k + stt.ntokens]
        assert blk.shape[0] == stt.ntokens
        assert (blk[:, 1] == stt.tree_bits).all() and (blk[:, 2] == stt.ll_bits).all() and (blk[:, 3] == stt.d_bits).all(), s_
        assert int(blk[0, 1]) + int(blk[0, 2]) + int(blk[0, 3]) + 3 == stt.block_bits
        k += stt.ntokens
    assert k == recs.shape[0]


@pytest.mark.parametrize("n,piece", [(0, 0), (1, 0), (70000, 0), (3 * 65536, 65536), (3 * 65536 + 1, 65536), (5 * 65536 - 7, 65536),
                                     (1000003, 131072), (40 << 20, 8 << 20)])
def test_fd_pipeline_matches_whole_buffer(D, O, tmp_path, n, piece):
    """dfl_b200_compress_fd (reader thread, pinned pieces, three streams, writer thread) writes the stream the
    whole-buffer call writes -- pieces are whole segments primed with the 32 KiB before them -- for inputs that
    end inside a piece, exactly on a piece boundary, and for an empty input; regular files (pread by several
    threads) and pipes (read until 0)."""
    data = make_input("markov", 91, 0, n) if n else np.zeros(0, dtype=np.uint8)
    comp = D.Compressor()
    want = comp.compress(data)
    if n <= 1000003:
        assert want == O.zlib_compress(data, O.params())
    src, dst = tmp_path / "in", tmp_path / "out.z"
    src.write_bytes(data.tobytes())
    fi, fo = os.open(src, os.O_RDONLY), os.open(dst, os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
    try:
        nin, nout = comp.compress_fd(fi, fo, piece)
        assert os.lseek(fi, 0, os.SEEK_CUR) == n  # the offset is where read(2) would have left it
    finally:
        os.close(fi), os.close(fo)
    assert (nin, nout) == (n, len(want))
    assert dst.read_bytes() == want
    # the same through pipes, fed and drained by threads
    import threading

    r_in, w_in = os.pipe()
    r_out, w_out = os.pipe()
    got = bytearray()

    def feed():
        view = memoryview(data.tobytes())
        pos = 0
        while pos < len(view):
            pos += os.write(w_in, view[pos:pos + (1 << 20)])
        os.close(w_in)

    def drain():
        while True:
            b = os.read(r_out, 1 << 20)
            if not b:
                break
            got.extend(b)

    tf, td = threading.Thread(target=feed), threading.Thread(target=drain)
    tf.start(), td.start()
    try:
        nin, nout = comp.compress_fd(r_in, w_out, piece)
    finally:
        os.close(w_out)
        tf.join(), td.join()
        os.close(r_in), os.close(r_out)
    assert (nin, nout) == (n, len(want)) and bytes(got) == want
    comp.close()


def test_compress_host_multi_one_gpu_many_contexts(D, O):
    """dfl_b200_compress_host_multi with several contexts (all on GPU 0 here; one per GPU in bench.py --gpus N):
    contiguous segment-aligned shards, each primed with the 32 KiB before it, bodies in rank order, Adler-32
    combined in C -- the stream is the single-context stream."""
    for n, k in ((0, 3), (5, 2), (32768 * 3, 4), (700001, 3), (2000003, 8), (100, 5)):
        data = make_input("markov", 92, 0, n) if n else np.zeros(0, dtype=np.uint8)
        comps = [D.Compressor() for _ in range(k)]
        cap = comps[0].bound(n) + 64
        out = np.empty(cap, dtype=np.uint8)
        z_len = D.compress_multi(comps, data.ctypes.data if n else 0, n, out.ctypes.data, cap)
        z = out[:z_len].tobytes()
        assert z == O.zlib_compress(data, O.params()), (n, k)
        assert zlib.decompress(z) == data.tobytes()
        for c in comps:
            c.close()
    assert D.lib().dfl_b200_adler32_combine(O.adler32(b"abc"), O.adler32(b"defgh"), 5) == O.adler32(b"abcdefgh")


@pytest.mark.parametrize("kind,seed,n", [("markov", 2, 200000), ("log", 3, 150000), ("random", 4, 70000), ("markov", 2, 2)])
def test_check_lld_target(D, tmp_path, kind, seed, n):
    """`make check_lld` (the reference's test target, Makefile:26-27): a C program compiled against
    include/deflate_ext.h that forks, compresses in the child and, in the parent, expands the compress_stats
    records (tests/check_lld.c:20-54) and inflates the stream; exit status 0 iff both give the file back."""
    import subprocess

    exe = os.path.join(os.path.dirname(D.lib_path()), "check_lld")
    assert os.path.exists(exe)
    src = tmp_path / "in"
    src.write_bytes(make_input(kind, seed, 0, n).tobytes())
    r = subprocess.run([exe, str(src)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout.startswith("check_lld ok: %d bytes" % n)


def test_fd_api_pipe_and_errors(D):
    r, w = os.pipe()
    os.write(w, b"hello hello hello hello")
    os.close(w)
    r2, w2 = os.pipe()
    assert D.deflate_compress(r, w2, -1, 32768, 0) == 0
    os.close(w2)
    z = os.read(r2, 1 << 16)
    os.close(r), os.close(r2)
    assert zlib.decompress(z) == b"hello hello hello hello"
    r, w = os.pipe()
    os.close(w)
    assert D.deflate_compress(r, -1, -1, 0, 0) < 0  # window 0 is out of range
    os.close(r)


def test_device_api_shard_with_history(D, O):
    """dfl_b200_deflate_device on a shard that is primed with the preceding bytes (multi-GPU building block)."""
    import torch

    data = make_input("markov", 5, 0, 400000)
    cut = 3 * 32768 + 5000  # shard boundary (not a multiple of the segment size on purpose)
    dev = torch.device("cuda:0")
    d_in = torch.from_numpy(data).to(dev)
    comp = D.Compressor()
    bodies, adlers = [], []
    for lo, hi, final in ((0, cut, False), (cut, data.size, True)):
        n = hi - lo
        hist = min(lo, 32768)
        cap = comp.bound(n)
        d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
        out_len, adler = comp.deflate_device(d_in.data_ptr() + lo, n, hist, final, d_out.data_ptr(), cap)
        body = d_out[:out_len].cpu().numpy().tobytes()
        ob, _ = O.deflate_body(data, lo, hi, O.params(), is_final=final)
        assert body == ob
        assert adler == O.adler32(data[lo:hi])
        bodies.append(body)
        adlers.append((adler, n))
    a = adlers[0][0]
    a = O.adler32_combine(a, adlers[1][0], adlers[1][1])
    stream = b"\x78\x9c" + b"".join(bodies) + a.to_bytes(4, "big")
    assert zlib.decompress(stream) == data.tobytes()


def test_large_roundtrip_properties(D, O):
    """64 MiB of text: size-independent properties (round trip, segment accounting, Adler-32)."""
    n = 64 << 20
    data = make_input("markov", 2, 0, n)
    comp = D.Compressor()
    z = comp.compress(data)
    raw = zlib.decompress(z)
    assert len(raw) == n and raw == data.tobytes()
    assert int.from_bytes(z[-4:], "big") == zlib.adler32(raw)
    # first and last segments against the oracle
    comp.keep_tokens(True)
    z2 = comp.compress(data)
    assert z2 == z
    op = O.params()
    for s in (0, 1, n // 32768 - 1):
        assert np.array_equal(comp.tokens(s), O.lz77_segment(data, 0, s * 32768, (s + 1) * 32768, op))


def test_unaligned_device_pointers(D, O):
    import torch

    data = make_input("log", 3, 0, 100000)
    dev = torch.device("cuda:0")
    comp = D.Compressor()
    for shift_in, shift_out in ((1, 0), (3, 5), (7, 13), (15, 1)):
        buf = torch.zeros(data.size + 64, dtype=torch.uint8, device=dev)
        buf[shift_in:shift_in + data.size] = torch.from_numpy(data).to(dev)
        cap = comp.bound(data.size)
        d_out = torch.zeros(cap + 64, dtype=torch.uint8, device=dev)
        out_len, adler = comp.deflate_device(buf.data_ptr() + shift_in, data.size, 0, True,
                                             d_out.data_ptr() + shift_out, cap)
        body = d_out[shift_out:shift_out + out_len].cpu().numpy().tobytes()
        ob, _ = O.deflate_body(data, 0, data.size, O.params(), is_final=True)
        assert body == ob
        assert int(d_out[shift_out + out_len:].max()) == 0 and int(d_out[:shift_out].sum()) == 0


# --------------------------------------------------------------------------- BASELINE.json configs as parity cases
# (sizes scaled to what a test run can hold; the properties are size independent)

def _roundtrip_device(D, data: np.ndarray, **prm):
    """Compress device-resident data, inflate with zlib, compare; returns (stream bytes, compressor)."""
    import torch

    dev = torch.device("cuda:0")
    comp = D.Compressor(**prm)
    d_in = torch.from_numpy(data).to(dev)
    cap = comp.bound(data.size)
    d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
    out_len, adler = comp.deflate_device(d_in.data_ptr(), data.size, 0, True, d_out.data_ptr(), cap)
    body = d_out[:out_len].cpu().numpy().tobytes()
    raw = zlib.decompressobj(-15).decompress(body)
    assert len(raw) == data.size and raw == data.tobytes()
    assert adler == zlib.adler32(raw)
    assert comp.sort_fallbacks() == 0
    return body, comp


def test_config3_repetitive_log(D, O):
    """configs[2]: highly repetitive log data (long matches, deep buckets); 256 MiB here."""
    data = make_input("log", 3, 0, 256 << 20)
    body, comp = _roundtrip_device(D, data)
    assert len(body) < data.size // 20
    # first / last segments bit-exact against the oracle
    ob, _ = O.deflate_body(data[: 4 * 32768], 0, 4 * 32768, O.params(), is_final=False)
    assert body[: len(ob)] == ob


def test_config4_incompressible(D, O):
    """configs[3]: random bytes -> stored blocks, output = input + 10 bytes per 32 KiB segment; 512 MiB here."""
    import torch

    n = 512 << 20
    dev = torch.device("cuda:0")
    d_in = torch.empty(n, dtype=torch.uint8, device=dev)
    assert D.lib().dfl_b200_synth_random(d_in.data_ptr(), n, 4, None) == 0
    torch.cuda.synchronize()
    host = d_in.cpu().numpy()
    assert np.array_equal(host[: 1 << 20], make_input("random", 4, 0, 1 << 20))  # device generator == host generator
    comp = D.Compressor()
    cap = comp.bound(n)
    d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
    out_len, adler = comp.deflate_device(d_in.data_ptr(), n, 0, True, d_out.data_ptr(), cap)
    assert out_len == n + 10 * (n // 32768)  # every segment stored: 5 B header + 5 B sync marker
    body = d_out[:out_len].cpu().numpy()
    seg = body[: 32778]
    assert seg[0] == 0 and seg[1] == 0x00 and seg[2] == 0x80 and np.array_equal(seg[5:32773], host[:32768])
    assert adler == zlib.adler32(host.tobytes())
    assert zlib.decompressobj(-15).decompress(body.tobytes()) == host.tobytes()


def test_config5_mixed_corpus_sharded(D, O):
    """configs[4]: text + PNG-filtered scanline stripes, sharded over ranks (2 emulated ranks on one GPU),
    bodies concatenated by rank order inflate to the original; 128 MiB here with 8 MiB stripes."""
    import torch
    from deflate_compression_b200 import shard, synth

    n = 128 << 20
    data = synth.mixed(0, n, stripe=8 << 20)
    dev = torch.device("cuda:0")
    d_all = torch.from_numpy(data).to(dev)
    comp = D.Compressor()
    bodies, adlers, lens = [], [], []
    world = 4
    for r in range(world):
        lo, hi = shard.shard_range(n, r, world)
        hist = min(lo, 32768)
        cap = comp.bound(hi - lo)
        d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
        out_len, adler = comp.deflate_device(d_all.data_ptr() + lo, hi - lo, hist, r == world - 1, d_out.data_ptr(), cap)
        bodies.append(d_out[:out_len].cpu().numpy().tobytes())
        adlers.append(adler)
        lens.append(hi - lo)
    z = shard.stitch(bodies, adlers, lens)
    assert zlib.decompress(z) == data.tobytes()
    # identical to the single-rank stream (segments are independent of the sharding)
    assert z == comp.compress(data)
    ratio = n / len(z)
    assert 1.5 < ratio < 4.0


def test_config2_text_256mib(D, O):
    """configs[1] at full size: 256 MiB Markov text, every block dynamic, round trip, sampled token parity."""
    n = 256 << 20
    data = make_input("markov", 2, 0, n)
    comp = D.Compressor()
    comp.keep_tokens(True)
    z = comp.compress(data)
    assert zlib.decompress(z) == data.tobytes()
    nseg = comp.num_segments()
    assert nseg == n // 32768
    assert all(comp.seg_info(s).btype == 2 for s in range(0, nseg, 97))
    op = O.params()
    for s in list(range(0, nseg, 64)) + [nseg - 1]:
        assert np.array_equal(comp.tokens(s), O.lz77_segment(data, 0, s * 32768, (s + 1) * 32768, op)), s
    # vs the REAL reference semantics on no-dictionary segments: >= 64 segments
    comp0 = D.Compressor(use_dict=0)
    comp0.keep_tokens(True)
    sub = data[: 64 * 32768]
    comp0.compress(sub)
    op0 = O.params(use_dict=0)
    for s in range(64):
        assert np.array_equal(comp0.tokens(s), O.lz77_segment(sub, 0, s * 32768, (s + 1) * 32768, op0)), s


def test_error_paths(D):
    """Error behaviour of the C-ABI: small output buffers are reported, never overrun."""
    import ctypes as C

    import torch

    data = make_input("random", 4, 0, 100000)
    comp = D.Compressor()
    out = np.zeros(50000 + 64, dtype=np.uint8)
    with pytest.raises(D.DflError) as e:
        comp.compress_into(data.ctypes.data, data.size, out.ctypes.data, 50000)
    assert e.value.code == -6  # E_SZ
    assert int(out[50000:].max()) == 0  # nothing past the capacity was touched
    big = make_input("random", 9, 0, 80 << 20)  # pipelined host path
    out = np.zeros((40 << 20) + 64, dtype=np.uint8)
    with pytest.raises(D.DflError) as e:
        comp.compress_into(big.ctypes.data, big.size, out.ctypes.data, 40 << 20)
    assert e.value.code == -6
    assert int(out[40 << 20:].max()) == 0
    dev = torch.device("cuda:0")
    d_in = torch.from_numpy(data).to(dev)
    d_out = torch.zeros(60000, dtype=torch.uint8, device=dev)
    with pytest.raises(D.DflError) as e:
        comp.deflate_device(d_in.data_ptr(), data.size, 0, True, d_out.data_ptr(), 40000)
    assert e.value.code == -6
    assert int(d_out[40000:].max()) == 0
    # the context is still usable afterwards
    z = comp.compress(data)
    assert zlib.decompress(z) == data.tobytes()
    L = D.lib()
    assert L.dfl_b200_compress_host(None, None, 0, None, 0, None) == -11  # E_INVAL
    ctx = C.c_void_p()
    assert L.dfl_b200_create(C.byref(ctx), 99, None) == -10  # no such device: E_RANGE


def test_context_reuse_and_param_switch(D, O):
    """One context, many calls of different sizes and parameters (workspace reuse / regrowth)."""
    comp = D.Compressor()
    for n in (10, 70000, 3, 33 << 20, 1000, 70 << 20, 40000):
        data = make_input("markov", 7, 0, n)
        assert zlib.decompress(comp.compress(data)) == data.tobytes()
    for prm in (dict(seg_bytes=2048), dict(seg_bytes=32768, window=512), dict(chain_limit=3)):
        p = D.DflParams(prm.get("window", 32768), prm.get("seg_bytes", 32768), prm.get("chain_limit", 0), 1, 0)
        assert D.lib().dfl_b200_set_params(comp._ctx, p) == 0
        comp.params = p
        data = make_input("log", 3, 0, 200000)
        assert comp.compress(data) == O.zlib_compress(data, O.params(**prm))


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("kind,seed,n", [("markov", 11, 300001), ("random", 12, 100000), ("log", 13, 200000),
                                         ("zeros", 0, 70000), ("markov", 14, 5)])
def test_parse_kernels_agree(D, O, mode, kind, seed, n):
    """k_parse (CTA per segment, pointer doubling) and k_parse_warp (warp per segment) resolve the
    same greedy parse: tokens == oracle's with either one forced."""
    data = make_input(kind, seed, 0, n)
    comp = D.Compressor()
    comp.set_parse_mode(mode)
    comp.keep_tokens(True)
    z = comp.compress(data)
    op = O.params()
    for s in range(comp.num_segments()):
        lo, hi = s * 32768, min(n, (s + 1) * 32768)
        assert np.array_equal(comp.tokens(s), O.lz77_segment(data, 0, lo, hi, op)), f"mode {mode} segment {s}"
    assert z == O.zlib_compress(data, op)
    comp.close()


@pytest.mark.parametrize("kind,seed,n", [("markov", 15, 1200001), ("random", 16, 300000), ("log", 17, 700000), ("png", 18, 500000),
                                         ("zeros", 0, 200000), ("period7", 0, 150000), ("markov", 19, 32768 * 3 + 1),
                                         ("markov", 19, 32768 * 2 + 2), ("markov", 19, 32768 * 2), ("markov", 20, 40)])
def test_match_kernels_agree(D, O, kind, seed, n):
    """k_match_run (runs of consecutive segments per CTA, the window's sorted index inherited) and k_match (one
    CTA per segment, window and segment sorted together) find the same tokens: == the oracle's with either one,
    and the streams are identical."""
    data = make_input(kind, seed, 0, n)
    op = O.params()
    streams = []
    for mode in (0, 1):
        comp = D.Compressor()
        comp.set_match_mode(mode)
        comp.keep_tokens(True)
        streams.append(comp.compress(data))
        for s in range(comp.num_segments()):
            lo, hi = s * 32768, min(n, (s + 1) * 32768)
            assert np.array_equal(comp.tokens(s), O.lz77_segment(data, 0, lo, hi, op)), f"match mode {mode} segment {s}"
        assert comp.sort_fallbacks() == 0
        comp.close()
    assert streams[0] == streams[1] == O.zlib_compress(data, op)


def test_match_run_shard_with_partial_history(D, O):
    """k_match_run on a shard whose history is shorter than a window (and not a multiple of 16): the first
    segment's window block is partial, later ones inherit."""
    import torch

    whole = make_input("markov", 23, 0, 500000)
    for hist in (0, 1, 777, 32767, 32768):
        lo = 40000
        dev = torch.device("cuda:0")
        d_all = torch.from_numpy(whole).to(dev)
        comp = D.Compressor()
        comp.keep_tokens(True)
        n = whole.size - lo
        cap = comp.bound(n)
        d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
        comp.deflate_device(d_all.data_ptr() + lo, n, hist, True, d_out.data_ptr(), cap, torch.cuda.current_stream().cuda_stream)
        op = O.params()
        for s in range(comp.num_segments()):
            a, b = lo + s * 32768, min(whole.size, lo + (s + 1) * 32768)
            want = O.lz77_segment(whole, lo - hist, a, b, op)
            assert np.array_equal(comp.tokens(s), want), f"hist {hist} segment {s}"
        comp.close()


@pytest.mark.parametrize("prm", [dict(), dict(block_mode=1), dict(block_mode=2), dict(block_mode=3), dict(seg_bytes=5000)])
def test_seg_info_cost_split(D, O, prm):
    """dfl_seg_info's tree / lit-len / distance bits (k_huff) == the oracle's, and add up to the block."""
    data = np.concatenate([make_input("markov", 44, 0, 100000), make_input("random", 45, 0, 40000),
                           make_input("log", 46, 0, 70000), make_input("zeros", 0, 0, 40000)])
    comp = D.Compressor(**prm)
    comp.keep_tokens(True)
    comp.compress(data)
    body, stats = O.deflate_body(data, 0, data.size, O.params(**prm))
    assert comp.num_segments() == len(stats)
    for s_, stt in enumerate(stats):
        info = comp.seg_info(s_)
        assert (info.btype, info.block_bits) == (stt.btype, stt.block_bits), s_
        assert (info.tree_bits, info.ll_bits, info.d_bits) == (stt.tree_bits, stt.ll_bits, stt.d_bits), s_
        assert info.tree_bits + info.ll_bits + info.d_bits + 3 == info.block_bits
    comp.close()


# ---- segment index + GPU inflate ----

def _gpu_roundtrip(D, data, **prm):
    comp = D.Compressor(**prm)
    z = comp.compress(data)
    off = comp.seg_index()
    seg = prm.get("seg_bytes", 32768)
    nseg = (data.size + seg - 1) // seg
    assert off.size == nseg + 1 and int(off[-1]) == len(z) - 6
    assert all(int(off[i]) < int(off[i + 1]) for i in range(nseg)), "segments are non-empty and in order"
    back = comp.decompress(z, off, data.size)
    comp.close()
    return z, off, back


@pytest.mark.parametrize("kind,seed,n", [("markov", 21, 1), ("markov", 21, 2), ("markov", 21, 700), ("markov", 22, 32768),
                                         ("markov", 23, 32769), ("markov", 24, 1000003), ("random", 25, 200001),
                                         ("log", 26, 500000), ("png", 27, 400000), ("zeros", 0, 300000),
                                         ("period7", 0, 100000), ("const_a", 0, 65536)])
def test_gpu_inflate_roundtrip(D, kind, seed, n):
    """compress -> (stream, segment index) -> k_inflate gives the input back, bit for bit, and the
    stream is what zlib inflates too (so the GPU decoder reads standard DEFLATE)."""
    data = make_input(kind, seed, 0, n)
    z, off, back = _gpu_roundtrip(D, data)
    assert back == data.tobytes()
    assert zlib.decompress(z) == data.tobytes()


@pytest.mark.parametrize("prm", [dict(seg_bytes=2048), dict(seg_bytes=16384, window=4096), dict(use_dict=0),
                                 dict(block_mode=1), dict(block_mode=2), dict(block_mode=3), dict(chain_limit=4),
                                 dict(seg_bytes=1000, window=32768)])
def test_gpu_inflate_params(D, prm):
    """Small segments (matches reaching many segments back), no dictionary, forced stored / fixed /
    dynamic blocks."""
    data = np.concatenate([make_input("markov", 31, 0, 150000), make_input("random", 32, 0, 30000),
                           make_input("log", 33, 0, 120000)])
    z, off, back = _gpu_roundtrip(D, data, **prm)
    assert back == data.tobytes()


def test_gpu_inflate_empty_and_errors(D):
    comp = D.Compressor()
    z = comp.compress(np.zeros(0, dtype=np.uint8))
    off = comp.seg_index()
    assert off.size == 1
    assert comp.decompress(z, off, 0) == b""
    data = make_input("markov", 41, 0, 200000)
    z = comp.compress(data)
    off = comp.seg_index()
    assert comp.decompress(z, off, data.size) == data.tobytes()
    bad = bytearray(z)
    bad[len(bad) // 2] ^= 0x55  # corrupt the body: the decoder must fail (or the Adler-32 must), never hang
    with pytest.raises(D.DflError):
        comp.decompress(bytes(bad), off, data.size)
    bad = bytearray(z)
    bad[-1] ^= 1  # wrong Adler-32
    with pytest.raises(D.DflError):
        comp.decompress(bytes(bad), off, data.size)
    bad_off = off.copy()
    bad_off[2] += 3  # wrong index
    with pytest.raises(D.DflError):
        comp.decompress(z, bad_off, data.size)
    assert comp.decompress(z, off, data.size) == data.tobytes()  # the context survives
    comp.close()


def test_gpu_inflate_device_large(D):
    """64 MiB, device to device: inflate(compress(x)) == x and the Adler-32 values agree."""
    import torch

    n = 64 << 20
    data = np.concatenate([make_input("markov", 51, 0, n // 2), make_input("png", 52, 0, n // 4),
                           make_input("random", 53, 0, n // 8), make_input("log", 54, 0, n // 8)])
    dev = torch.device("cuda:0")
    d_in = torch.from_numpy(data).to(dev)
    comp = D.Compressor()
    cap = comp.bound(n)
    d_z = torch.empty(cap, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    out_len, adler = comp.deflate_device(d_in.data_ptr(), n, 0, True, d_z.data_ptr(), cap, st)
    nseg = n // 32768
    d_off = torch.from_numpy(comp.seg_index().astype(np.int64)).to(dev)
    d_back = torch.zeros(n, dtype=torch.uint8, device=dev)
    adler2 = comp.inflate_device(d_z.data_ptr(), out_len, d_off.data_ptr(), nseg, d_back.data_ptr(), n, 0, st)
    torch.cuda.synchronize()
    assert adler2 == adler == (zlib.adler32(data.tobytes()) & 0xFFFFFFFF)
    assert torch.equal(d_back, d_in)
    comp.close()


def _structured(rng, n):
    """Bytes with the structures an LZ77 search has to get right: runs, short periods, far repeats just
    inside / outside the window, repeats across segment boundaries, text, noise."""
    out = bytearray()
    pool = [bytes(rng.integers(0, 256, rng.integers(3, 400), dtype=np.uint8)) for _ in range(6)]
    while len(out) < n:
        r = rng.integers(0, 8)
        if r == 0:
            out += bytes([int(rng.integers(0, 256))]) * int(rng.integers(1, 700))
        elif r == 1:
            p = bytes(rng.integers(97, 123, rng.integers(1, 9), dtype=np.uint8))
            out += p * int(rng.integers(1, 200))
        elif r == 2 and len(out) > 10:
            d = int(rng.integers(1, min(len(out), 40000)))
            ln = int(rng.integers(3, 600))
            for _ in range(ln):  # overlapping copy, like an LZ77 match
                out.append(out[-d])
        elif r == 3:
            out += pool[int(rng.integers(0, len(pool)))]
        elif r == 4:
            out += bytes(rng.integers(0, 256, rng.integers(1, 3000), dtype=np.uint8))
        elif r == 5:
            out += bytes(rng.integers(0, 4, rng.integers(1, 2000), dtype=np.uint8))
        elif r == 6 and len(out) > 32768:
            d = int(rng.integers(32760, 32776))  # around the window edge
            if d <= len(out):
                out += out[-d:len(out) - d + int(rng.integers(3, 300))]
        else:
            out += bytes(rng.integers(32, 127, rng.integers(1, 500), dtype=np.uint8))
    return np.frombuffer(bytes(out[:n]), dtype=np.uint8).copy()


@pytest.mark.parametrize("seed", range(40))
def test_fuzz_structured_inputs(D, O, seed):
    """Seeded fuzz: structured inputs x random parameters; tokens, stream bytes and both inflaters
    (CPU and GPU) must agree with the oracle / the input."""
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.choice([1, 7, 300, 5000, 32768, 40000, 70001, 150000, 260000]))
    data = _structured(rng, n)
    prm = {}
    if rng.integers(0, 3) == 0:
        prm["seg_bytes"] = int(rng.choice([512, 4096, 20000, 32768]))
    if rng.integers(0, 3) == 0:
        prm["window"] = int(rng.choice([1, 300, 4096, 32768]))
    if rng.integers(0, 4) == 0:
        prm["use_dict"] = 0
    if rng.integers(0, 4) == 0:
        prm["chain_limit"] = int(rng.choice([1, 5, 64]))
    if rng.integers(0, 4) == 0:
        prm["block_mode"] = int(rng.integers(0, 4))
    z = check_stream(D, O, data, **prm)
    comp = D.Compressor(**prm)
    z2 = comp.compress(data)
    assert z2 == z
    assert comp.decompress(z2, comp.seg_index(), data.size) == data.tobytes()
    comp.close()


# ---- PNG scanline filter + container ----

def _img(h, w, c, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([(xx * (k + 1) // 3 + yy * (4 - k) // 2 + 40 * k) for k in range(c)], axis=-1)
    img = img + rng.integers(-2, 3, size=img.shape)
    img[h // 3:h // 2, w // 4:w // 2] = rng.integers(0, 256, size=img[h // 3:h // 2, w // 4:w // 2].shape)
    return (img & 255).astype(np.uint8)


@pytest.mark.parametrize("h,w,c", [(1, 1, 1), (3, 5, 3), (17, 33, 4), (300, 257, 3), (64, 4096, 3), (1000, 100, 2), (7, 5000, 1)])
@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4, 5])
def test_png_filter_matches_oracle(D, O, h, w, c, mode):
    """k_png_filter == the oracle's forward filter, byte for byte (filter choice included)."""
    import torch

    img = _img(h, w, c, 3 * h + w + c)
    dev = torch.device("cuda:0")
    d_pix = torch.from_numpy(img.reshape(-1)).to(dev)
    d_out = torch.zeros(h * (w * c + 1), dtype=torch.uint8, device=dev)
    comp = D.Compressor()
    comp.png_filter_device(d_pix.data_ptr(), w * c, h, c, d_out.data_ptr(), mode, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), O.png_filter(img, mode))
    comp.close()


def _png_chunks(png: bytes):
    assert png[:8] == b"\x89PNG\r\n\x1a\n"
    pos, chunks = 8, []
    while pos < len(png):
        n = int.from_bytes(png[pos:pos + 4], "big")
        typ, data = png[pos + 4:pos + 8], png[pos + 8:pos + 8 + n]
        assert int.from_bytes(png[pos + 8 + n:pos + 12 + n], "big") == (zlib.crc32(typ + data) & 0xFFFFFFFF), typ
        chunks.append((typ, data))
        pos += 12 + n
    assert pos == len(png)
    return chunks


@pytest.mark.parametrize("h,w,c", [(1, 1, 1), (48, 64, 3), (33, 17, 4), (20, 50, 2), (600, 800, 3)])
def test_png_encode_file(D, O, h, w, c):
    """The file has the structure src/png/png_decoder.c reads (signature, IHDR, IDAT, IEND with good
    CRCs); its IDAT is a zlib stream of exactly the oracle's filtered scanlines, and for the small images
    un-filtering them (the reference's defilter direction) gives the pixels back."""
    from test_oracle import _png_unfilter

    img = _img(h, w, c, 11 * h + w)
    comp = D.Compressor()
    png = comp.png_encode(img)
    comp.close()
    chunks = _png_chunks(png)
    assert [t for t, _ in chunks] == [b"IHDR", b"IDAT", b"IEND"]
    ihdr = chunks[0][1]
    assert int.from_bytes(ihdr[0:4], "big") == w and int.from_bytes(ihdr[4:8], "big") == h
    assert ihdr[8] == 8 and ihdr[9] == {1: 0, 2: 4, 3: 2, 4: 6}[c] and ihdr[10:13] == b"\x00\x00\x00"
    filtered = np.frombuffer(zlib.decompress(chunks[1][1]), dtype=np.uint8)
    assert np.array_equal(filtered, O.png_filter(img, 0))
    assert D.deflate_decompress(chunks[1][1]) == filtered.tobytes()
    if h * w <= 4096:
        assert np.array_equal(_png_unfilter(filtered, h, w * c, c), img.reshape(h, w * c))


def test_png_then_inflate_then_png_one_context(D, O):
    """One context: png_encode, then a decompress that has to grow the inflate workspace, then png_encode again.
    (Round 1 freed the PNG scanline buffer inside the inflate grow branch and kept the dangling pointer.)"""
    img = _img(300, 400, 3, 77)
    comp = D.Compressor()
    png1 = comp.png_encode(img)
    data = make_input("markov", 61, 0, 3 * 1000 * 1000)  # more words than anything this context inflated before
    z = comp.compress(data)
    off = comp.seg_index()
    assert comp.decompress(z, off, data.size) == data.tobytes()
    png2 = comp.png_encode(img)
    assert png1 == png2
    filtered = np.frombuffer(zlib.decompress(_png_chunks(png2)[1][1]), dtype=np.uint8)
    assert np.array_equal(filtered, O.png_filter(img, 0))
    big = _img(900, 1200, 3, 78)  # grows the scanline buffer after the inflate
    png3 = comp.png_encode(big)
    assert np.array_equal(np.frombuffer(zlib.decompress(_png_chunks(png3)[1][1]), dtype=np.uint8), O.png_filter(big, 0))
    comp.close()


@pytest.mark.parametrize("seed", range(6))
def test_gpu_inflate_corrupt_streams_fail_cleanly(D, seed):
    """Corrupt bodies and wrong indexes end in DflError (never a fault): every word a corrupt segment did not
    decode reads as a value, and the resolve rounds refuse distances that reach before the history."""
    rng = np.random.default_rng(1000 + seed)
    comp = D.Compressor()
    data = np.concatenate([make_input("markov", 70 + seed, 0, 300000), make_input("log", 80 + seed, 0, 200000)])
    z = comp.compress(data)
    off = comp.seg_index()
    for _ in range(4):
        bad = bytearray(z)
        for _k in range(int(rng.integers(1, 6))):
            bad[int(rng.integers(2, len(bad) - 4))] ^= int(rng.integers(1, 256))
        try:
            back = comp.decompress(bytes(bad), off, data.size)
            assert back != data.tobytes() or bytes(bad) == z
        except D.DflError:
            pass
        bad_off = off.copy()
        bad_off[int(rng.integers(1, off.size - 1))] += int(rng.integers(1, 40))
        with pytest.raises(D.DflError):
            comp.decompress(z, bad_off, data.size)
    assert comp.decompress(z, off, data.size) == data.tobytes()  # the context survives all of it
    comp.close()


# ---- adaptive block splitting (block_mode 4) ----

@pytest.mark.gpu
@pytest.mark.parametrize("kind,seed,n", [("markov", 2, 150001), ("log", 3, 100000), ("png", 6, 100000), ("random", 4, 70000),
                                         ("zeros", 0, 70000), ("striped", 11, 300000), ("striped", 12, 4097), ("markov", 1, 4096),
                                         ("markov", 1, 8200), ("striped", 13, 40000), ("striped", 14, 32768 * 3 + 5)])
def test_split_stream_equals_oracle(D, O, kind, seed, n):
    """block_mode 4: same tokens as the oracle, the emitted stream == the oracle's byte for byte (block boundaries,
    types, headers, the byte alignment of a stored block inside a segment), and it inflates everywhere."""
    from conftest import striped_input
    data = striped_input(n, seed) if kind == "striped" else make_input(kind, seed, 0, n)
    check_stream(D, O, data, block_mode=4)


@pytest.mark.gpu
@pytest.mark.parametrize("prm", [dict(seg_bytes=16384), dict(seg_bytes=5000), dict(window=4096), dict(use_dict=0), dict(chain_limit=8)], ids=str)
def test_split_with_other_geometries(D, O, prm):
    from conftest import striped_input
    check_stream(D, O, striped_input(120000, 21), block_mode=4, **prm)


@pytest.mark.gpu
def test_split_seg_info_matches_oracle_stats(D, O):
    """Per segment: block count, summed block bits and the cost split == orc_seg_stats; never more bits than unsplit."""
    from conftest import striped_input
    data = striped_input(1 << 18, 11)
    comp = D.Compressor(block_mode=4)
    comp.keep_tokens(True)
    z = comp.compress(data)
    _, st4 = O.deflate_body(data, 0, data.size, O.params(block_mode=4))
    _, st0 = O.deflate_body(data, 0, data.size, O.params(block_mode=0))
    assert comp.num_segments() == len(st4)
    split_any = False
    for s, (a, b) in enumerate(zip(st4, st0)):
        i = comp.seg_info(s)
        assert (i.nblocks, i.btype, i.block_bits, i.out_bytes) == (a.nblocks, a.btype, a.block_bits, a.out_bytes), s
        assert (i.tree_bits, i.ll_bits, i.d_bits) == (a.tree_bits, a.ll_bits, a.d_bits), s
        assert i.tree_bits + i.ll_bits + i.d_bits + 3 * i.nblocks == i.block_bits <= b.block_bits
        split_any |= i.nblocks > 1
    assert split_any
    assert len(z) < len(zlib.compress(data.tobytes(), 6))
    comp.close()


@pytest.mark.gpu
def test_split_gpu_inflate_and_sharded(D, O):
    """A split stream through the GPU inflate (several blocks per segment, stored blocks among them), and block_mode 4
    on a shard with history == the oracle's body for that range."""
    import torch
    from conftest import striped_input
    data = striped_input(400000, 31)
    comp = D.Compressor(block_mode=4)
    z = comp.compress(data)
    assert comp.decompress(z, comp.seg_index(), data.size) == data.tobytes()
    lo, hi = 3 * 32768, 9 * 32768
    dev = torch.device("cuda:0")
    d_in = torch.from_numpy(data).to(dev)
    cap = comp.bound(hi - lo)
    d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
    out_len, adler = comp.deflate_device(d_in.data_ptr() + lo, hi - lo, 32768, False, d_out.data_ptr(), cap)
    ob, _ = O.deflate_body(data, lo, hi, O.params(block_mode=4), is_final=False)
    assert d_out[:out_len].cpu().numpy().tobytes() == ob
    assert adler == O.adler32(data[lo:hi])
    comp.close()


@pytest.mark.gpu
def test_split_through_the_reference_entry_point(D, O, tmp_path, monkeypatch):
    """DFL_B200_SPLIT=1 switches deflate_compress(fd_in, fd_out, fd_stats, sw, ops) to block_mode 4, with and without
    a record stream; the next call without it is back to one block per segment (the context is parked in between)."""
    from conftest import striped_input
    data = striped_input(300000, 41)
    src = tmp_path / "in"
    src.write_bytes(data.tobytes())

    def run(with_stats):
        dst, st = tmp_path / "out.z", tmp_path / "stats"
        fi = os.open(src, os.O_RDONLY)
        fo = os.open(dst, os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
        fs = os.open(st, os.O_WRONLY | os.O_CREAT | os.O_TRUNC) if with_stats else -1
        try:
            assert D.deflate_compress(fi, fo, fs, 32768, 0) == 0
        finally:
            os.close(fi), os.close(fo)
            if fs >= 0:
                os.close(fs)
        return dst.read_bytes()

    z4 = O.zlib_compress(data, O.params(block_mode=4))
    z0 = O.zlib_compress(data, O.params())
    assert len(z4) < len(z0)
    monkeypatch.setenv("DFL_B200_SPLIT", "1")
    assert run(False) == z4
    assert run(True) == z4
    monkeypatch.delenv("DFL_B200_SPLIT")
    assert run(False) == z0


@pytest.mark.gpu
def test_segment_index_of_an_empty_input(D, O):
    """No segments: the index is its end entry alone, the length of the lone empty stored block."""
    comp = D.Compressor()
    z = comp.compress(np.zeros(0, dtype=np.uint8))
    assert z == O.zlib_compress(np.zeros(0, dtype=np.uint8), O.params())
    assert comp.seg_index().tolist() == [5]
    comp.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n,piece", [(5 << 20, 1 << 20), (40 << 20, 8 << 20), (3 * 65536 + 1, 65536)])
def test_fd_pipeline_mapped_output(D, O, tmp_path, n, piece):
    """An output fd that can be mapped (opened read-write: a regular file, a memfd) is filled through a shared mapping
    instead of pwrite(2): same bytes, the file offset left where write(2) would have left it, also when the output does
    not start at offset 0 and when the file already holds more bytes than the stream (nothing is truncated)."""
    data = make_input("markov", 92, 0, n)
    comp = D.Compressor()
    want = comp.compress(data)
    src = tmp_path / "in"
    src.write_bytes(data.tobytes())
    # regular file, read-write, output appended after 4099 bytes that are already there
    dst = tmp_path / "out.z"
    dst.write_bytes(b"x" * 4099)
    fi, fo = os.open(src, os.O_RDONLY), os.open(dst, os.O_RDWR)
    try:
        os.lseek(fo, 4099, os.SEEK_SET)
        nin, nout = comp.compress_fd(fi, fo, piece)
        assert (nin, nout) == (n, len(want))
        assert os.lseek(fo, 0, os.SEEK_CUR) == 4099 + len(want)
    finally:
        os.close(fi), os.close(fo)
    got = dst.read_bytes()
    assert got[:4099] == b"x" * 4099 and got[4099:] == want
    # memfd, written twice: the second stream is shorter than what the file holds
    mi, mo = os.memfd_create("in"), os.memfd_create("out")
    try:
        os.write(mi, data.tobytes())
        for m in (n, n // 2):
            os.ftruncate(mi, m)
            os.lseek(mi, 0, os.SEEK_SET), os.lseek(mo, 0, os.SEEK_SET)
            nin, nout = comp.compress_fd(mi, mo, piece)
            z = os.pread(mo, nout, 0)
            assert nin == m and zlib.decompress(z) == data[:m].tobytes()
            assert os.lseek(mo, 0, os.SEEK_CUR) == nout
    finally:
        os.close(mi), os.close(mo)
    comp.close()


@pytest.mark.gpu
def test_device_list_through_the_reference_entry_point(D, O, tmp_path, monkeypatch):
    """DFL_B200_DEVICES names several devices: deflate_compress(fd_in, fd_out, -1, sw, 0) shards the input over one
    context per entry (here all on GPU 0, plus 0,1 when there is a second GPU) and writes the stream one GPU writes;
    regular file, memfd and pipe inputs; with a record stream asked for the call stays on the first device."""
    import threading
    import torch

    data = make_input("markov", 17, 0, 1500001)
    want = O.zlib_compress(data, O.params())
    src = tmp_path / "in"
    src.write_bytes(data.tobytes())
    lists = ["0,0,0"] + (["0,1"] if torch.cuda.device_count() > 1 else [])
    for devs in lists:
        monkeypatch.setenv("DFL_B200_DEVICES", devs)
        # regular file in, read-write file out
        dst = tmp_path / "out.z"
        fi, fo = os.open(src, os.O_RDONLY), os.open(dst, os.O_RDWR | os.O_CREAT | os.O_TRUNC)
        try:
            assert D.deflate_compress(fi, fo, -1, 32768, 0) == 0
            assert os.lseek(fi, 0, os.SEEK_CUR) == data.size
        finally:
            os.close(fi), os.close(fo)
        assert dst.read_bytes() == want, devs
        # pipe in, memfd out
        r_in, w_in = os.pipe()
        mo = os.memfd_create("out")

        def feed():
            view = memoryview(data.tobytes())
            pos = 0
            while pos < len(view):
                pos += os.write(w_in, view[pos:pos + (1 << 20)])
            os.close(w_in)

        t = threading.Thread(target=feed)
        t.start()
        try:
            assert D.deflate_compress(r_in, mo, -1, 32768, 0) == 0
        finally:
            t.join()
            os.close(r_in)
        size = os.lseek(mo, 0, os.SEEK_CUR)
        assert os.pread(mo, size, 0) == want, devs
        os.close(mo)
        # a record stream keeps the call on one device; same stream
        fi, fo, fs = os.open(src, os.O_RDONLY), os.open(dst, os.O_WRONLY | os.O_TRUNC), os.open(tmp_path / "stats", os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
        try:
            assert D.deflate_compress(fi, fo, fs, 32768, 0) == 0
        finally:
            os.close(fi), os.close(fo), os.close(fs)
        assert dst.read_bytes() == want
    monkeypatch.delenv("DFL_B200_DEVICES")
    # empty input on several devices
    monkeypatch.setenv("DFL_B200_DEVICES", "0,0")
    mi, mo = os.memfd_create("in"), os.memfd_create("out")
    assert D.deflate_compress(mi, mo, -1, 32768, 0) == 0
    size = os.lseek(mo, 0, os.SEEK_CUR)
    assert os.pread(mo, size, 0) == O.zlib_compress(np.zeros(0, dtype=np.uint8), O.params())
    os.close(mi), os.close(mo)
