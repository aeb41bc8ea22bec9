"""CPU tests of the host side: the C-ABI library loads and exports every symbol
the headers declare, fails loudly without a GPU, the CPU half of the drop-in API
(deflate_decompress) behaves like the reference's contract, the synthetic
generators are deterministic, and the multi-rank stitching works over gloo."""
import ctypes
import os
import re
import subprocess
import sys
import zlib

import numpy as np
import pytest

from conftest import ROOT, make_input


def declared_functions(header: str):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b([a-z_0-9]+)\s*\([^;{]*\)\s*;", src)) - {"defined"})


def test_library_exports_every_declared_symbol(built):
    L = ctypes.CDLL(built.lib_path())
    names = declared_functions("deflate_b200.h") + declared_functions("deflate_ext.h")
    assert "deflate_compress" in names and "dfl_b200_deflate_device" in names and len(names) >= 25
    for name in names:
        assert hasattr(L, name), f"{name} declared in include/ but not exported"


def test_product_does_not_link_or_import_the_oracle(built):
    out = subprocess.run(["ldd", built.lib_path()], capture_output=True, text=True).stdout
    assert "oracle" not in out
    pkg = os.path.join(ROOT, "deflate_compression_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in text and "liboracle" not in text, f


def test_no_cpu_fallback(built):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(built.DflError):
        built.Compressor()
    r, w = os.pipe()
    os.write(w, b"abc")
    os.close(w)
    assert built.deflate_compress(r, -1, -1, 32768, 0) < 0  # -E_GPU: no silent CPU path
    os.close(r)


def test_params_validation_happens_before_cuda(built):
    import torch

    if torch.cuda.is_available():
        for bad in (dict(window=0), dict(window=40000), dict(seg_bytes=0), dict(seg_bytes=65536), dict(block_mode=7)):
            with pytest.raises(built.DflError) as e:
                built.Compressor(**bad)
            assert e.value.code == -10  # E_RANGE


def test_deflate_decompress_contract(built):
    data = make_input("markov", 1, 0, 50000).tobytes()
    z = zlib.compress(data, 6)
    assert built.deflate_decompress(z) == data
    assert built.deflate_decompress(z, ops=1) == data + b"\0"          # DEFLATE_NULLTERM
    assert built.deflate_decompress(zlib.compress(b"abc\0"), ops=1) == b"abc\0"  # already terminated
    assert built.deflate_decompress(b"") == b""                        # "no data, skip"
    E = 1 << 24
    bad = bytearray(z)
    bad[-1] ^= 1
    with pytest.raises(built.DflError) as e:
        built.deflate_decompress(bytes(bad))
    assert e.value.code == -(E + 5)  # E_ZADL32
    with pytest.raises(built.DflError) as e:
        built.deflate_decompress(b"\x78\x9d" + z[2:])
    assert e.value.code == -(E + 7)  # E_ZFCHCK
    with pytest.raises(built.DflError) as e:
        built.deflate_decompress(b"\x79\x9b" + z[2:])  # CM != 8, FCHECK ok?
    assert e.value.code in (-(E + 7), -(E + 8))
    with pytest.raises(built.DflError):
        built.deflate_decompress(z[: len(z) // 2])
    # FDICT is rejected like the reference does (src/deflate_decompress.c:361-364)
    hdr = bytearray(b"\x78\x20")
    hdr[1] += 31 - ((hdr[0] * 256 + hdr[1]) % 31)
    with pytest.raises(built.DflError) as e:
        built.deflate_decompress(bytes(hdr) + z[2:])
    assert e.value.code == -(E + 10)  # E_ZPDICT


def test_synth_is_deterministic(built):
    from deflate_compression_b200 import synth

    a, b = synth.markov(100000, 2), synth.markov(100000, 2)
    assert np.array_equal(a, b) and not np.array_equal(a, synth.markov(100000, 3))
    assert np.array_equal(synth.markov(50000, 2), a[:50000])  # prefix stable
    assert np.array_equal(synth.random_bytes(4096, 4)[:1000], synth.random_bytes(1000, 4))
    lg = synth.log(100000)
    assert lg.tobytes().count(b"\n") > 300
    mx = synth.mixed(1000, 9000, stripe=4096)
    assert np.array_equal(mx, synth.mixed(0, 9000, stripe=4096)[1000:])
    assert 0.2 < len(zlib.compress(a.tobytes(), 6)) / a.size < 0.6  # "compressible English-like text"


def test_zencode_cli_decompress(built, tmp_path):
    exe = os.path.join(ROOT, "deflate_compression_b200", "zencode")
    assert os.path.exists(exe)
    data = make_input("log", 3, 0, 123457).tobytes()
    (tmp_path / "in.z").write_bytes(zlib.compress(data))
    r = subprocess.run([exe, "-d", str(tmp_path / "in.z"), str(tmp_path / "out")], capture_output=True)
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "out").read_bytes() == data


WORKER = r"""
import os, sys, zlib
import numpy as np
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
from deflate_compression_b200 import shard, synth
from oracle import pyoracle as O
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=int(sys.argv[3]), world_size=int(sys.argv[4]))
rank, world = dist.get_rank(), dist.get_world_size()
data = np.concatenate([synth.markov(300001, 5), synth.random_bytes(100000, 4), synth.log(250000, 3)])
prm = O.params()
def deflate_range(d, lo, hi, final):  # the oracle stands in for the GPU on the CPU box
    body, _ = O.deflate_body(d, lo, hi, prm, is_final=final)
    return body, O.adler32(d[lo:hi])
z = shard.compress_sharded(data, rank, world, deflate_range, dist=dist)
if rank == 0:
    assert zlib.decompress(z) == data.tobytes()
    assert z == O.zlib_compress(data, prm), "sharded stream differs from the single-rank stream"
    print("OK", len(z))
dist.barrier()
dist.destroy_process_group()
"""


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_stitch_over_gloo(built, tmp_path, world):
    """N > 1 host logic: contiguous shards, gather to rank 0, stitched stream == single-rank stream."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = 29500 + os.getpid() % 2000 + world
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r), str(world)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "OK" in outs[0]


def test_shard_ranges_cover_input():
    from deflate_compression_b200 import shard

    for n in (0, 1, 32768, 32769, 1000000, (1 << 28) + 5):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard.shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            assert all(lo % 32768 == 0 for lo, _ in ranges)
