"""Multi-GPU sharding of one stream.

The input shards contiguously, one range per rank; every range is compressed
independently (primed with the preceding <= 32 KiB, ended byte-aligned by its
sync marker), so the only exchange is the gather of sizes and bytes to rank 0,
which stitches header + bodies + combined Adler-32.  The helpers here are pure
host logic and work with any torch.distributed backend (NCCL on the GPUs; the
CPU tests drive them over gloo)."""
from __future__ import annotations

from typing import Callable, List, Tuple

import numpy as np

ADLER_MOD = 65521


def shard_range(n: int, rank: int, world: int, seg_bytes: int = 32768) -> Tuple[int, int]:
    """Contiguous range of rank `rank`: boundaries rounded to whole segments."""
    nseg = (n + seg_bytes - 1) // seg_bytes
    lo_seg = nseg * rank // world
    hi_seg = nseg * (rank + 1) // world
    return min(n, lo_seg * seg_bytes), min(n, hi_seg * seg_bytes)


def adler32_combine(a1: int, a2: int, len2: int) -> int:
    """Adler-32 of a concatenation from the parts' checksums (each part started from 1)."""
    s1a, s2a, s1b, s2b = a1 & 0xFFFF, a1 >> 16, a2 & 0xFFFF, a2 >> 16
    s1 = (s1a + s1b + ADLER_MOD - 1) % ADLER_MOD
    s2 = (s2a + s2b + (len2 % ADLER_MOD) * ((s1a + ADLER_MOD - 1) % ADLER_MOD)) % ADLER_MOD
    return (s2 << 16) | s1


def stitch(bodies: List[bytes], adlers: List[int], lengths: List[int]) -> bytes:
    """zlib stream from per-rank raw deflate bodies (rank order) and their Adler-32 / input lengths."""
    a = 1
    for ad, ln in zip(adlers, lengths):
        a = adler32_combine(a, ad, ln)
    return b"\x78\x9c" + b"".join(bodies) + a.to_bytes(4, "big")


def compress_sharded(data: np.ndarray, rank: int, world: int,
                     deflate_range: Callable[[np.ndarray, int, int, bool], Tuple[bytes, int]],
                     seg_bytes: int = 32768, window: int = 32768, dist=None):
    """Compress this rank's range with `deflate_range(data, lo, hi, is_final) -> (body, adler)` and gather
    to rank 0, which returns the complete zlib stream (other ranks return None)."""
    n = int(data.size)
    lo, hi = shard_range(n, rank, world, seg_bytes)
    body, adler = deflate_range(data, lo, hi, rank == world - 1)
    if dist is None or world == 1:
        return stitch([body], [adler], [hi - lo])
    import torch

    meta = torch.tensor([len(body), adler, hi - lo], dtype=torch.int64)
    metas = [torch.zeros(3, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(metas, meta)
    sizes = [int(m[0]) for m in metas]
    mine = torch.from_numpy(np.frombuffer(body, dtype=np.uint8).copy()) if body else torch.zeros(0, dtype=torch.uint8)
    if rank == 0:
        bodies = [body]
        for r in range(1, world):
            buf = torch.empty(sizes[r], dtype=torch.uint8)
            if sizes[r]:
                dist.recv(buf, src=r)
            bodies.append(buf.numpy().tobytes())
        return stitch(bodies, [int(m[1]) for m in metas], [int(m[2]) for m in metas])
    if sizes[rank]:
        dist.send(mine, dst=0)
    return None
