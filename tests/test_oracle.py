"""CPU tests of the oracle (the checker): pinned against the REAL reference's
token streams (fixtures made from oracle/_ref by tests/golden/make_golden.py,
and oracle/_ref itself when it is present), against zlib as an independent
decoder, and against RFC 1951 invariants."""
import hashlib
import zlib

import numpy as np
import pytest

from conftest import golden_cases, make_input


@pytest.fixture(scope="module")
def O(built):
    from oracle import pyoracle

    return pyoracle


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c[0])
def test_oracle_matches_reference_fixture(O, case):
    name, kind, seed, off, n, ref_toks, last, sha = case
    data = make_input(kind, seed, off, n)
    assert hashlib.sha256(data.tobytes()).digest() == sha.tobytes(), "generator drifted from the fixture"
    ot = O.lz77_segment(data, 0, 0, n, O.params(use_dict=0))
    assert ot.size == ref_toks.size and np.array_equal(ot, ref_toks)
    assert O.tokens_expand(ot) == data.tobytes()  # check_lld method


def test_oracle_matches_live_reference(O):
    if not O.have_ref():
        pytest.skip("oracle/_ref/libref.so not built (needs /root/reference)")
    rng = np.random.default_rng(11)
    for trial in range(12):
        n = int(rng.integers(3, 32769))
        kind = ["markov", "log", "png", "random"][trial % 4]
        data = make_input(kind, 100 + trial, int(rng.integers(0, 5000)), n)
        rt = O.ref_records_to_tokens(O.ref_tokens(data))
        ot = O.lz77_segment(data, 0, 0, n, O.params(use_dict=0))
        assert np.array_equal(rt, ot), (kind, n)


def test_reference_cost_fields_against_oracle(O):
    """The reference's own estimate: ll_bits / d_bits of the last record are the optimal
    unlimited Huffman cost of the token histogram plus one NYT leaf; the oracle's dynamic block for the same
    tokens must not cost more than that estimate + extra bits + the header estimate by more than a few bits
    per symbol of length limiting (none at these sizes)."""
    for case in golden_cases():
        name, kind, seed, off, n, ref_toks, last, sha = case
        if n < 3000:
            continue
        data = make_input(kind, seed, off, n)
        body, stats = O.deflate_body(data, 0, n, O.params(use_dict=0, block_mode=3))
        extra = 0
        for t in ref_toks:
            d, l = int(t) >> 16, int(t) & 0xFFFF
            if d:
                lt = l - 3
                extra += 0 if (l == 258 or lt < 8) else (lt.bit_length() - 1) - 2
                dt = d - 1
                extra += 0 if dt < 4 else (dt.bit_length() - 1) - 1
        ref_est = int(last[1]) + int(last[2]) + int(last[3]) + extra + 3
        # the reference's last record predates its final token's AHT update by design; allow 1% slack
        assert abs(stats[0].dynamic_bits - ref_est) <= 0.01 * ref_est + 64, (name, stats[0].dynamic_bits, ref_est)


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_cost_split_adds_up_and_matches_reference_fields(O, mode):
    """tree_bits + ll_bits + d_bits + 3 == block_bits for every block type (struct compress_stats' split,
    src/include/deflate_ext.h:19-31); for dynamic blocks ll_bits / d_bits less their extra bits are the Huffman
    cost of the histogram, which the reference's last record estimates (its AHT is the unlimited optimum + 1 NYT
    leaf), and tree_bits is what h_tree_d_lens sums (src/h_tree.c:75-148): both within 1 %."""
    for case in golden_cases():
        name, kind, seed, off, n, ref_toks, last, sha = case
        data = make_input(kind, seed, off, n)
        body, stats = O.deflate_body(data, 0, n, O.params(use_dict=0, block_mode=mode))
        st = stats[0]
        assert st.tree_bits + st.ll_bits + st.d_bits + 3 == st.block_bits, name
        assert st.tree_bits >= 0 and st.ll_bits >= 0 and st.d_bits >= 0
        if st.btype != 2:
            assert st.tree_bits == 0
        if st.btype == 0:
            assert st.d_bits == 0 and st.ll_bits == 8 * (n + 5 * max(1, (n + 65534) // 65535)) - 3
        if mode == 3 and n >= 3000:
            lx = dx = 0
            for t in ref_toks:
                d, l = int(t) >> 16, int(t) & 0xFFFF
                if d:
                    lt, dt = l - 3, d - 1
                    lx += 0 if (l == 258 or lt < 8) else (lt.bit_length() - 1) - 2
                    dx += 0 if dt < 4 else (dt.bit_length() - 1) - 1
            assert abs((st.ll_bits - lx) - int(last[2])) <= 0.01 * int(last[2]) + 32, (name, st.ll_bits - lx, int(last[2]))
            assert abs((st.d_bits - dx) - int(last[3])) <= 0.01 * int(last[3]) + 32, (name, st.d_bits - dx, int(last[3]))


@pytest.mark.parametrize("kind,seed", [("markov", 1), ("random", 4), ("log", 3), ("png", 6), ("zeros", 0),
                                        ("period7", 0)])
@pytest.mark.parametrize("prm", [dict(), dict(seg_bytes=4096), dict(use_dict=0), dict(window=1024),
                                 dict(chain_limit=8), dict(block_mode=1), dict(block_mode=2), dict(block_mode=3)],
                         ids=str)
def test_stream_roundtrip(O, kind, seed, prm):
    data = make_input(kind, seed, 0, 100001)
    z = O.zlib_compress(data, O.params(**prm))
    assert zlib.decompress(z) == data.tobytes()
    assert O.zlib_decompress(z, data.size) == data.tobytes()


def test_edge_sizes(O):
    base = make_input("markov", 1, 0, 70000)
    for n in list(range(0, 12)) + [257, 258, 259, 32767, 32768, 32769, 65535, 65536, 65537]:
        z = O.zlib_compress(base[:n], O.params())
        assert zlib.decompress(z) == base[:n].tobytes()
    assert O.zlib_compress(b"", O.params()) == bytes.fromhex("789c010000ffff00000001")


def test_inflate_accepts_zlib_streams(O, built):
    data = make_input("markov", 2, 0, 300000).tobytes()
    for lvl in (0, 1, 6, 9):
        z = zlib.compress(data, lvl)
        assert O.zlib_decompress(z, len(data)) == data
        assert built.deflate_decompress(z) == data


def test_huffman_lengths_are_valid_and_limited(O):
    rng = np.random.default_rng(3)
    fib = [1, 1]
    while len(fib) < 40:
        fib.append(fib[-1] + fib[-2])
    cases = [rng.integers(0, 1000, 286), rng.integers(0, 3, 286), np.array(fib[:30]), np.array(fib[:19]),
             np.zeros(30, dtype=np.int64), np.eye(1, 30, 7, dtype=np.int64)[0], np.ones(286, dtype=np.int64)]
    for freq in cases:
        for limit in (15, 7):
            if limit == 7 and freq.size > 19:
                continue
            lens = O.huff_lengths(freq, limit)
            used = lens[lens > 0]
            assert used.size >= 2 and used.max() <= limit
            assert sum(2.0 ** -int(l) for l in used) == pytest.approx(1.0)  # complete code
            assert all(lens[i] > 0 for i in range(freq.size) if freq[i] > 0)
            # canonical codes are prefix free
            codes = O.canonical_codes(lens)
            strs = sorted(format(int(f"{int(c):0{int(l)}b}"[::-1], 2), f"0{int(l)}b") for c, l in zip(codes, lens) if l)
            assert all(not b.startswith(a) for a, b in zip(strs, strs[1:]))


def test_huffman_optimal_when_unconstrained(O):
    import heapq

    rng = np.random.default_rng(5)
    for _ in range(20):
        freq = rng.integers(1, 500, int(rng.integers(2, 60)))
        lens = O.huff_lengths(freq, 15)
        h = [int(f) for f in freq]
        heapq.heapify(h)
        cost = 0
        while len(h) > 1:
            a, b = heapq.heappop(h), heapq.heappop(h)
            cost += a + b
            heapq.heappush(h, a + b)
        assert int((freq * lens).sum()) == cost


def test_adler_combine(O):
    data = make_input("png", 6, 0, 200000)
    cut = 77777
    a1, a2 = O.adler32(data[:cut]), O.adler32(data[cut:])
    assert O.adler32_combine(a1, a2, data.size - cut) == zlib.adler32(data.tobytes())
    from deflate_compression_b200 import shard

    assert shard.adler32_combine(a1, a2, data.size - cut) == zlib.adler32(data.tobytes())


def test_dictionary_priming_helps_and_is_consistent(O):
    data = make_input("log", 3, 0, 200000)
    z1 = O.zlib_compress(data, O.params(use_dict=1))
    z0 = O.zlib_compress(data, O.params(use_dict=0))
    assert len(z1) < len(z0)
    # a shard body with history equals the same range of the whole stream's segments
    lo = 3 * 32768
    body, _ = O.deflate_body(data, lo, data.size, O.params(), is_final=True)
    whole, stats = O.deflate_body(data, 0, data.size, O.params(), is_final=True)
    skip = sum(s.out_bytes for s in stats[:3])
    assert whole[skip:] == body


# ---- PNG scanline filter + CRC ----

def _png_unfilter(filtered: np.ndarray, h: int, row: int, bpp: int) -> np.ndarray:
    """The reference's defilter() direction (src/png/png_decoder.c:819-857) with the specification's Paeth:
    an independent check that the oracle's forward filter is invertible and standard."""
    out = np.zeros((h, row), dtype=np.uint8)
    for y in range(h):
        t = int(filtered[y * (row + 1)])
        f = filtered[y * (row + 1) + 1:(y + 1) * (row + 1)].astype(np.int32)
        prev = out[y - 1].astype(np.int32) if y else np.zeros(row, dtype=np.int32)
        cur = np.zeros(row, dtype=np.int32)
        for x in range(row):
            a = cur[x - bpp] if x >= bpp else 0
            b = prev[x]
            c = prev[x - bpp] if x >= bpp else 0
            if t == 0:
                p = 0
            elif t == 1:
                p = a
            elif t == 2:
                p = b
            elif t == 3:
                p = (a + b) >> 1
            else:
                pp = a + b - c
                pa, pb, pc = abs(pp - a), abs(pp - b), abs(pp - c)
                p = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
            cur[x] = (f[x] + p) & 255
        out[y] = cur.astype(np.uint8)
    return out


def _test_image(h: int, w: int, c: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([(xx * (k + 1) // 3 + yy * (4 - k) // 2 + 40 * k) for k in range(c)], axis=-1)
    img = img + rng.integers(-2, 3, size=img.shape)
    img[h // 3:h // 2, w // 4:w // 2] = rng.integers(0, 256, size=img[h // 3:h // 2, w // 4:w // 2].shape)  # a noisy patch
    return (img & 255).astype(np.uint8)


@pytest.mark.parametrize("h,w,c", [(1, 1, 1), (3, 5, 3), (17, 33, 4), (40, 64, 3), (9, 100, 2), (25, 31, 1)])
@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4, 5])
def test_png_filter_inverts(O, h, w, c, mode):
    img = _test_image(h, w, c, 7 * h + w + c)
    f = O.png_filter(img, mode)
    assert f.size == h * (w * c + 1)
    types = f[::w * c + 1]
    assert types.max() <= 4 and (mode == 0 or (types == mode - 1).all())
    assert np.array_equal(_png_unfilter(f, h, w * c, c), img.reshape(h, w * c))


def test_png_filter_picks_minimum_sum(O):
    img = _test_image(30, 50, 3, 99)
    row = 150
    f0 = O.png_filter(img, 0)
    sums = []
    for m in range(1, 6):
        fm = O.png_filter(img, m).reshape(30, row + 1)[:, 1:].astype(np.int32)
        sums.append(np.where(fm < 128, fm, 256 - fm).sum(axis=1))
    sums = np.stack(sums, axis=1)
    assert np.array_equal(f0[::row + 1], sums.argmin(axis=1).astype(np.uint8))  # argmin = first minimum


def test_crc32_matches_zlib(O):
    import zlib

    for data in (b"", b"IEND", b"123456789", bytes(range(256)) * 9):
        assert O.crc32(data) == (zlib.crc32(data) & 0xFFFFFFFF)


# ---- adaptive block splitting (block_mode 4; SURVEY.md section 8(f) N2, README.md:15-20) ----

@pytest.mark.parametrize("kind,seed,n", [("markov", 2, 150001), ("log", 3, 100000), ("png", 6, 100000), ("random", 4, 70000),
                                         ("zeros", 0, 70000), ("striped", 11, 200000), ("striped", 12, 4097), ("markov", 1, 4096),
                                         ("markov", 1, 8200), ("striped", 13, 40000)])
def test_split_roundtrip_never_worse_same_tokens(O, kind, seed, n):
    """A split stream inflates (zlib and the oracle's inflate), holds the SAME tokens, and no segment takes more bits
    than it does as one block; tree + ll + d + 3 per block == block_bits."""
    from conftest import striped_input
    data = striped_input(n, seed) if kind == "striped" else make_input(kind, seed, 0, n)
    z4 = O.zlib_compress(data, O.params(block_mode=4))
    assert zlib.decompress(z4) == data.tobytes()
    assert O.zlib_decompress(z4, n) == data.tobytes()
    _, st0 = O.deflate_body(data, 0, n, O.params(block_mode=0))
    _, st4 = O.deflate_body(data, 0, n, O.params(block_mode=4))
    assert len(st0) == len(st4)
    for a, b in zip(st0, st4):
        assert a.ntokens == b.ntokens and a.nblocks == 1 and 1 <= b.nblocks <= 8
        assert b.block_bits <= a.block_bits
        assert (b.nblocks == 1) == (b.block_bits == a.block_bits)  # a split is taken only when it saves bits
        assert b.tree_bits + b.ll_bits + b.d_bits + 3 * b.nblocks == b.block_bits
        assert (a.stored_bits, a.fixed_bits, a.dynamic_bits) == (b.stored_bits, b.fixed_bits, b.dynamic_bits)
    assert len(z4) <= len(O.zlib_compress(data, O.params(block_mode=0)))


def test_split_pays_on_striped_data_and_beats_zlib6(O):
    """Where the statistics change inside a segment the split stream is clearly smaller -- smaller than zlib level 6,
    which the unsplit stream (one block per 32 KiB) is not."""
    from conftest import striped_input
    data = striped_input(1 << 18, 11)
    z0 = O.zlib_compress(data, O.params(block_mode=0))
    z4 = O.zlib_compress(data, O.params(block_mode=4))
    z6 = zlib.compress(data.tobytes(), 6)
    assert len(z4) < 0.98 * len(z0)
    assert len(z4) < len(z6) < len(z0)
    _, st4 = O.deflate_body(data, 0, data.size, O.params(block_mode=4))
    assert max(s.nblocks for s in st4) >= 4
