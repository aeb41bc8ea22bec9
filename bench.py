#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 DEFLATE compress path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): input GB/s compressed, device-timed, + ratio vs reference.
A step = one pass of the hot path (LZ77 match search + greedy parse + Huffman
construction + bit packing) over this rank's shard of synthetic order-2 Markov
text (BASELINE.json configs[1]: 256 MiB per GPU, dynamic Huffman blocks), with
the inputs already resident in HBM.  `e2e` is the same metric through the host
API (dfl_b200_compress_host: pinned host buffers, H2D + D2H inside the timed
region).  For N > 1 the input shards contiguously (weak scaling: 256 MiB per
rank, each shard primed with the preceding 32 KiB) and every step gathers sizes
and compressed bytes to rank 0 with NCCL; the transfer of step k runs while the
kernels of step k + 1 do (two output buffers), and the clock stops only after the
last step's bytes have arrived.

After the timed region the same line gets `other_workloads` (one device-timed pass each of BASELINE configs 3, 4
and 5: 1 GiB log lines, 4 GiB random bytes generated on the device, and this rank's share of the 16 GiB mixed
text / PNG-scanline corpus), `e2e_fd` (the reference entry point itself, deflate_compress(fd_in, fd_out, -1, 32768,
0) over memfds) and, for every N, `roundtrip_ok`: the bytes the timed loop wrote (gathered on rank 0 for N > 1,
stitched with the combined Adler-32) are inflated with zlib and compared with the corpus.

`--impl reference` times the reference's own CPU implementation of the path
(oracle/_ref/libref.so, the unmodified reference LZ77 loop, on independent
32 KiB slices) on all host cores; the default arm also times it single-threaded
on a bounded sample (`cpu_baseline`).  Only those two legs touch oracle/.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHARD_BYTES = 256 << 20          # per GPU (configs[1])
STRIPE = 32 << 20                # the corpus is a sequence of independently seeded 32 MiB Markov stripes
TEXT_SEED = 2                    # configs[1]: seed 2
WINDOW = 32768
CPU_SAMPLE = 16 << 20            # cpu_baseline sample (single thread)
REF_STEP_SAMPLE = 64 << 20       # --impl reference: bytes per step, all cores
LOG_BYTES = 1 << 30              # configs[2]
RANDOM_BYTES = 4 << 30           # configs[3]
MIXED_BYTES = 16 << 30           # configs[4], split over the ranks
MIXED_SEEDS = 4                  # distinct stripes per kind; the corpus cycles through them (a stripe is 1000 windows long)


def corpus(lo: int, hi: int) -> np.ndarray:
    """Bytes [lo, hi) of the benchmark corpus (order-2 Markov text)."""
    from deflate_compression_b200 import synth

    out = np.empty(hi - lo, dtype=np.uint8)
    pos = lo
    while pos < hi:
        k = pos // STRIPE
        s0 = k * STRIPE
        end = min(hi, s0 + STRIPE)
        full = synth.markov(end - s0, TEXT_SEED * 1000003 + k)
        out[pos - lo:end - lo] = full[pos - s0:]
        pos = end
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for t, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9 or not (t0 - 0.05 <= t <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak() -> tuple[float, str]:
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(n_in: int) -> float | None:
    """dram bytes (read + write) of the dominant kernel's launch, from the committed ncu capture of this command's own
    256 MiB launch (profiles/r02_k_match_traffic.json; null when the capture was taken at another launch size)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_k_match_traffic.json")) as f:
            j = json.load(f)
        return float(j["dram_bytes_per_launch"]) if int(j["launch_input_bytes"]) == n_in else None
    except Exception:
        return None


def bind_to_gpu_numa_node(index: int) -> str:
    """Run this rank (and first-touch its pinned buffers) on the CPUs next to its GPU."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus near gpu {index}"
    except Exception as e:  # no NVML, one NUMA node, or a restricted cpuset: run where we are
        return f"unbound ({type(e).__name__})"
    return "unbound"


def mixed_shard(lo: int, hi: int):
    """[lo - hist, hi) of config 5's corpus as (list of (kind, seed index, offset in stripe, length), hist): alternating 32 MiB
    stripes of Markov text and filtered scanlines; the stripes of a kind cycle through MIXED_SEEDS seeds."""
    hist = min(lo, WINDOW)
    parts, pos = [], lo - hist
    while pos < hi:
        k = pos // STRIPE
        end = min(hi, (k + 1) * STRIPE)
        parts.append((k % 2, (k // 2) % MIXED_SEEDS, pos - k * STRIPE, end - pos))
        pos = end
    return parts, hist


# --------------------------------------------------------------------------- reference arm / cpu baseline

def _ref_worker(args):
    path, lo, hi = args
    from oracle import pyoracle as O

    data = np.load(path, mmap_mode="r")[lo:hi]
    a = np.ascontiguousarray(data)
    if O.have_ref():
        return int(O.ref().ref_compress_slices(a.ctypes.data, a.size))
    prm = O.params(use_dict=0)
    return sum(O.lz77_segment(a, 0, s, min(a.size, s + 32768), prm).size for s in range(0, a.size, 32768))


def cpu_reference_throughput(sample: np.ndarray, procs: int, repeat: int = 1):
    """GB/s of the reference CPU path over `sample`, split over `procs` processes (32 KiB slices)."""
    import multiprocessing as mp
    import tempfile

    from oracle import pyoracle as O

    kind = "reference" if O.have_ref() else "port"
    if O.have_ref():
        O.ref()  # loaded here, before the pool forks: the workers inherit oracle/_ref/libref.so
    with tempfile.NamedTemporaryFile(suffix=".npy", delete=False) as f:
        np.save(f, sample)
        path = f.name
    try:
        per = ((sample.size + procs - 1) // procs + 32767) // 32768 * 32768
        jobs = [(path, lo, min(sample.size, lo + per)) for lo in range(0, sample.size, per)]
        times, tokens = [], 0
        if procs == 1:
            for _ in range(repeat):
                t0 = time.perf_counter()
                tokens = _ref_worker(jobs[0])
                times.append(time.perf_counter() - t0)
        else:
            with mp.get_context("fork").Pool(procs) as pool:
                pool.map(_ref_worker, [(path, 0, 32768)] * procs)  # start the workers
                for _ in range(repeat):
                    t0 = time.perf_counter()
                    tokens = sum(pool.map(_ref_worker, jobs))
                    times.append(time.perf_counter() - t0)
    finally:
        os.unlink(path)
    return kind, times, tokens


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = corpus(0, REF_STEP_SAMPLE)
    kind, times, tokens = cpu_reference_throughput(sample, cores, repeat=args.warmup + args.steps)
    t = times[args.warmup:]
    total = sum(t)
    value = sample.size * len(t) / total / 1e9
    line = {
        "impl": "reference", "metric": "input GB/s compressed", "value": value, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(t), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "256 MiB order-2 Markov text per GPU (BASELINE configs[1]); this arm: first 64 MiB per step",
                   "window": WINDOW, "segment_bytes": 32768, "chain_limit": 0},
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": cores, "kind": kind,
                         "sample": f"first {REF_STEP_SAMPLE >> 20} MiB of the workload as independent 32 KiB slices, "
                                   f"{cores} processes; {tokens} tokens per step"},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- GPU arm

def device_pass(comp, d_in: int, n: int, hist: int, is_final: bool, d_out, cap: int, stream: int, reps: int = 2):
    """`reps` device-timed passes of the whole path over one resident input; returns the last one's (ms, out_len, adler, kernel ms)."""
    import torch

    best = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        comp.enqueue(d_in, n, hist, is_final, d_out.data_ptr(), cap, stream)
        e1.record()
        out_len, adler = comp.finish(stream)
        torch.cuda.synchronize()
        best = (e0.elapsed_time(e1), out_len, adler, comp.kernel_ms())
    return best


def other_workloads(D, comp, dev, rank: int, world: int, stream: int, dist, peak: float) -> dict:
    """One device-timed pass (after one warm pass) each of BASELINE configs 3, 4 and 5, with the device round trip
    (segment-parallel inflate, compared on the device) of what the pass wrote."""
    import torch

    from deflate_compression_b200 import synth

    res = {}

    def roundtrip(d_buf, hist, n, d_out, out_len, adler):
        d_off = torch.from_numpy(comp.seg_index().astype(np.int64)).to(dev)
        d_back = torch.empty(hist + n, dtype=torch.uint8, device=dev)
        d_back[:hist].copy_(d_buf[:hist])
        a2 = comp.inflate_device(d_out.data_ptr(), out_len, d_off.data_ptr(), (n + 32767) // 32768, d_back.data_ptr() + hist, n, hist, stream)
        ok = bool(a2 == adler) and bool(torch.equal(d_back[hist:], d_buf[hist:hist + n]))
        del d_back, d_off
        return ok

    def run(name, d_buf, hist, n, is_final, what):
        cap = comp.bound(n)
        d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
        ms, out_len, adler, k = device_pass(comp, d_buf.data_ptr() + hist, n, hist, is_final, d_out, cap, stream)
        ok = roundtrip(d_buf, hist, n, d_out, out_len, adler)
        t = torch.tensor([ms, k["k_match"], 0.0 if ok else 1.0], dtype=torch.float64, device=dev)
        tot = torch.tensor([float(n), float(out_len)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        ms, km, bad = t.tolist()
        n_tot, out_tot = tot.tolist()
        res[name] = {"workload": what, "value": n_tot / (ms * 1e-3) / 1e9, "unit": "GB/s", "ms": ms, "input_bytes": int(n_tot),
                     "ratio": n_tot / out_tot, "kernel_ms": {a: round(b, 3) for a, b in k.items() if b},
                     "roofline": {"bound": "hbm", "kernel": "k_match", "frac": (n + out_len) / (km * 1e-3) / 1e9 / peak,
                                  "path_frac": (n_tot + out_tot) / world / (ms * 1e-3) / 1e9 / peak},
                     "roundtrip_ok": bad == 0.0}
        del d_out
        torch.cuda.empty_cache()

    # config 3: every rank takes the same 1 GiB of log lines (a replica, not a shard: the config is a 1-GPU one)
    d_log = torch.from_numpy(synth.log(LOG_BYTES, 3)).to(dev)
    run("log", d_log, 0, LOG_BYTES, True, "1 GiB repetitive log lines per GPU (BASELINE configs[2], seed 3)")
    del d_log
    # config 4: splitmix64 bytes generated on the device
    d_rnd = torch.empty(RANDOM_BYTES, dtype=torch.uint8, device=dev)
    D.lib().dfl_b200_synth_random(d_rnd.data_ptr(), RANDOM_BYTES, 4 + rank, stream)
    torch.cuda.synchronize()
    run("random", d_rnd, 0, RANDOM_BYTES, True, "4 GiB splitmix64 bytes per GPU, generated on the device (BASELINE configs[3], seed 4 + rank)")
    del d_rnd
    torch.cuda.empty_cache()
    # config 5: this rank's contiguous share of the 16 GiB mixed corpus (2 GiB of it on a single GPU)
    total = MIXED_BYTES if world > 1 else 2 << 30
    lo, hi = rank * (total // world), (rank + 1) * (total // world)
    parts, hist = mixed_shard(lo, hi)
    stripes = {}
    d_mix = torch.empty(hist + hi - lo, dtype=torch.uint8, device=dev)
    pos = 0
    for kind, sd, off, ln in parts:
        if (kind, sd) not in stripes:
            gen = synth.markov if kind == 0 else synth.png_rows
            stripes[(kind, sd)] = torch.from_numpy(gen(STRIPE, (5 if kind == 0 else 6) * 1000003 + sd)).to(dev)
        d_mix[pos:pos + ln].copy_(stripes[(kind, sd)][off:off + ln])
        pos += ln
    del stripes
    run("mixed", d_mix, hist, hi - lo, rank == world - 1,
        f"{total >> 30} GiB mixed corpus over {world} GPU(s): alternating 32 MiB stripes of Markov text and PNG-filtered RGB "
        f"scanlines cycling through {MIXED_SEEDS} seeds per kind, contiguous shards (BASELINE configs[4])")
    del d_mix
    torch.cuda.empty_cache()
    return res


def striped_corpus(n: int, seed: int = 7) -> np.ndarray:
    """Data whose statistics change every 2-12 KiB (text, random bytes, a numeric table in turn): what adaptive block
    splitting is for.  Cut from three long buffers, so generating it costs next to nothing."""
    from deflate_compression_b200 import synth

    rng = np.random.default_rng(seed)
    src = [synth.markov(n // 2 + 16384, 9), synth.random_bytes(n // 2 + 16384, 9),
           np.frombuffer(b"".join(b"%d,%d;" % (i, i * 7) for i in range(n // 20)), dtype=np.uint8)]
    pos = [0, 0, 0]
    out = np.empty(n, dtype=np.uint8)
    o = i = 0
    while o < n:
        k = i % 3
        m = min(int(rng.integers(2048, 12288)), n - o, src[k].size - pos[k])
        if m <= 0:
            pos[k] = 0
            continue
        out[o:o + m] = src[k][pos[k]:pos[k] + m]
        pos[k] += m
        o += m
        i += 1
    return out


def block_split_report(D, dev, local: int, d_buf, hist: int, n: int, stream: int, host_sample: np.ndarray, unsplit_ratio: float) -> dict:
    """Adaptive block splitting (block_mode 4) beside the headline: one device-timed pass over the
    resident benchmark text and one over 64 MiB of striped data, each with the device round trip, and the ratios next
    to zlib level 6 on a sample."""
    import torch
    import zlib

    res = {}
    comp4 = D.Compressor(device=local, block_mode=4)
    comp4.enable_timing(True)
    comp0 = D.Compressor(device=local)

    def one(name, d_in, h, nn, sample):
        cap = comp4.bound(nn)
        d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
        ms, out_len, adler, k = device_pass(comp4, d_in.data_ptr() + h, nn, h, True, d_out, cap, stream)
        d_off = torch.from_numpy(comp4.seg_index().astype(np.int64)).to(dev)
        d_back = torch.empty(h + nn, dtype=torch.uint8, device=dev)
        d_back[:h].copy_(d_in[:h])
        a2 = comp4.inflate_device(d_out.data_ptr(), out_len, d_off.data_ptr(), (nn + 32767) // 32768, d_back.data_ptr() + h, nn, h, stream)
        ok = bool(a2 == adler) and bool(torch.equal(d_back[h:], d_in[h:h + nn]))
        _, out0, _, _ = device_pass(comp0, d_in.data_ptr() + h, nn, h, True, d_out, cap, stream, reps=1)
        res[name] = {"value": nn / (ms * 1e-3) / 1e9, "unit": "GB/s", "ms": ms, "input_bytes": int(nn), "ratio_split": nn / out_len,
                     "ratio_one_block_per_segment": nn / out0, "bytes_saved_pct": 100.0 * (out0 - out_len) / out0,
                     "ratio_zlib6_on_sample": sample.size / len(zlib.compress(sample.tobytes(), 6)), "sample_bytes": int(sample.size),
                     "kernel_ms": {a: round(b, 3) for a, b in k.items() if b}, "roundtrip_ok": ok}
        del d_out, d_back, d_off

    one("text", d_buf, hist, n, host_sample)
    striped = striped_corpus(64 << 20)
    d_s = torch.from_numpy(striped).to(dev)
    one("striped", d_s, 0, striped.size, striped[:16 << 20])
    del d_s
    comp4.close()
    comp0.close()
    torch.cuda.empty_cache()
    res["what"] = ("block_mode 4: a 32 KiB segment is cut where its tokens cross multiples of 4 KiB and bisected top-down while "
                   "two halves, each its own cheapest block by exact cost, take fewer bits than one block; same tokens")
    return res


def e2e_fd_throughput(D, host_bytes: np.ndarray, reps: int = 3):
    """GB/s of the reference entry point itself: deflate_compress(fd_in, fd_out, -1, 32768, 0) over memfds."""
    fi, fo = os.memfd_create("dfl_in"), os.memfd_create("dfl_out")
    try:
        view = memoryview(host_bytes)
        pos = 0
        while pos < len(view):
            pos += os.write(fi, view[pos:pos + (64 << 20)])
        times, size = [], 0
        for _ in range(reps + 1):
            os.lseek(fi, 0, os.SEEK_SET)
            os.lseek(fo, 0, os.SEEK_SET)  # (the output memfd keeps its pages from the warm-up call: no page allocation in the timed calls)
            t0 = time.perf_counter()
            rc = D.deflate_compress(fi, fo, -1, 32768, 0)
            times.append(time.perf_counter() - t0)
            if rc:
                return None, None, rc
            size = os.lseek(fo, 0, os.SEEK_CUR)
        os.lseek(fo, 0, os.SEEK_SET)
        z = os.read(fo, size)
        while len(z) < size:
            z += os.read(fo, size - len(z))
        return host_bytes.size / min(times[1:]) / 1e9, z, 0
    finally:
        os.close(fi), os.close(fo)


def run_gpu(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    binding = bind_to_gpu_numa_node(local)  # before anything is allocated: pinned buffers are first-touched here
    import torch
    import torch.distributed as dist

    import deflate_compression_b200 as D

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime

        # (measured at N = 2: capping NCCL's send / recv at one channel makes the 125 MB per step the bottleneck, 28 instead of
        # 39 GB/s; the default fan of channels costs 4 % against compress-only.  DFL_BENCH_P2P_CHANNELS overrides.)
        if os.environ.get("DFL_BENCH_P2P_CHANNELS"):
            os.environ.setdefault("NCCL_MAX_P2P_NCHANNELS", os.environ["DFL_BENCH_P2P_CHANNELS"])
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    D.build()

    # ---- synthetic input: this rank's contiguous shard, primed with the preceding 32 KiB ----
    lo, hi = rank * SHARD_BYTES, (rank + 1) * SHARD_BYTES
    hist = min(lo, WINDOW)
    host = corpus(lo - hist, hi)
    n = hi - lo
    d_buf = torch.from_numpy(host).to(dev)              # [history | shard] resident in HBM
    d_in = d_buf.data_ptr() + hist
    comp = D.Compressor(device=local)
    comp.enable_timing(True)
    cap = comp.bound(n)
    # output buffers in rotation: the gather of step k overlaps the kernels of step k + 1 (three when there is a gather:
    # a shard's buffer is written again two steps after rank 0 was seen to have taken it)
    NB = 3 if world > 1 else 2
    d_outs = [torch.empty(cap, dtype=torch.uint8, device=dev) for _ in range(NB)]
    d_out = d_outs[0]
    stream = torch.cuda.current_stream().cuda_stream
    is_final = rank == world - 1

    gather_bufs = None
    if world > 1 and rank == 0:
        gather_bufs = [torch.empty(world * cap, dtype=torch.uint8, device=dev) for _ in range(NB)]
    pending = [None] * NB
    last_gather = {}

    # How the shards get to rank 0.  "nccl" (default): batched isend / irecv; its kernels take SMs from the match kernels
    # (N = 8: 143.6 of 160.8 GB/s compress-only).  "ipc" (DFL_BENCH_GATHER=ipc): every rank opens its output buffers to
    # rank 0 through CUDA IPC and rank 0's copy engine pulls the bytes over NVLink on a stream of its own -- no SM taken,
    # but one copy stream does not carry seven shards a step: measured 40.2 against 41.2 GB/s at N = 2 and 81 against 144
    # at N = 8, so it stays an option.  Falls back to "nccl" when the buffers cannot be opened.
    gather_mode = os.environ.get("DFL_BENCH_GATHER", "nccl") if world > 1 else "none"
    peer_bufs, copy_stream, prev_copy = None, None, [None]
    if gather_mode == "ipc":
        ok = 1
        mine = None
        try:
            from torch.multiprocessing.reductions import reduce_tensor

            mine = [reduce_tensor(t) for t in d_outs]
        except Exception:
            ok = 0
        handles = [None] * world
        dist.all_gather_object(handles, mine)
        if rank == 0 and ok:
            try:
                peer_bufs = {r: [fn(*a) for fn, a in handles[r]] for r in range(1, world)}
                copy_stream = torch.cuda.Stream(device=dev)
                torch.cuda.set_device(local)
            except Exception:
                ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            gather_mode, peer_bufs = "nccl", None

    def drain(b=None):
        """Output buffer b's transfer (or every buffer's) has landed; drain() without an argument also joins the second
        stream into the current one, where the clock's events are recorded."""
        for k in (range(NB) if b is None else (b,)):
            if pending[k] is not None:
                if gather_mode == "ipc":
                    pending[k].synchronize()
                else:
                    with torch.cuda.stream(tstreams[k & 1]):
                        for q in pending[k]:
                            q.wait()
                pending[k] = None
        if b is None:
            torch.cuda.current_stream().wait_stream(tstreams[1])
            if copy_stream is not None:
                torch.cuda.current_stream().wait_stream(copy_stream)

    # two contexts on two streams take the steps in turn (a context has one call in flight): step k + 1's kernels are
    # already enqueued while the host reads step k's size back -- and, for N > 1, exchanges the sizes and posts the transfer
    comps = [comp, D.Compressor(device=local)]
    tstreams = [torch.cuda.current_stream(), torch.cuda.Stream(device=dev)]
    for c in comps:
        c.enable_timing(True)

    def launch(i: int):
        b, ob = i & 1, i % NB
        drain(ob)
        comps[b].enqueue(d_in, n, hist, is_final, d_outs[ob].data_ptr(), cap, tstreams[b].cuda_stream)

    def collect(i: int, with_gather: bool = True):
        """Sizes of step i (its kernels are done), then the gather of sizes and bytes to rank 0; the transfer itself is
        left in flight and is waited for before its buffers are used again, or by drain()."""
        b, ob = i & 1, i % NB
        out_len, adler = comps[b].finish(tstreams[b].cuda_stream)
        if world > 1 and with_gather:
            if gather_mode == "ipc" and rank == 0 and prev_copy[0] is not None:
                # rank 0 joins the exchange of sizes only once it has taken the step before: the others read that as
                # "the buffer of two steps ago is free" (they write it again one step from now)
                prev_copy[0].synchronize()
            with torch.cuda.stream(tstreams[b]):
                sizes = torch.zeros(2 * world, dtype=torch.int64, device=dev)
                mine = torch.tensor([out_len, adler], dtype=torch.int64, device=dev)
                dist.all_gather_into_tensor(sizes, mine)
                meta = sizes.tolist()
                sz = meta[0::2]
                if rank == 0:
                    offs = np.concatenate([[0], np.cumsum(sz)])
                    gbuf = gather_bufs[ob]
                    if gather_mode == "ipc":
                        cs = copy_stream.cuda_stream
                        L = D.lib()
                        L.dfl_b200_copy_async(gbuf.data_ptr(), d_outs[ob].data_ptr(), sz[0], cs)
                        for r in range(1, world):
                            L.dfl_b200_copy_async(gbuf.data_ptr() + int(offs[r]), peer_bufs[r][ob].data_ptr(), sz[r], cs)
                        ev = torch.cuda.Event()
                        ev.record(copy_stream)
                        pending[ob] = ev
                        prev_copy[0] = ev
                    else:
                        gbuf[:sz[0]].copy_(d_outs[ob][:sz[0]])
                        ops = [dist.P2POp(dist.irecv, gbuf[int(offs[r]):int(offs[r]) + sz[r]], r) for r in range(1, world)]
                        pending[ob] = dist.batch_isend_irecv(ops)
                    last_gather.update(buf=ob, sizes=sz, adlers=meta[1::2])
                elif gather_mode != "ipc":
                    pending[ob] = dist.batch_isend_irecv([dist.P2POp(dist.isend, d_outs[ob][:out_len], 0)])
        return out_len, adler

    def step(i: int = 0, with_gather: bool = True):
        """One pass of the hot path over the shard, start to finish."""
        launch(i)
        return collect(i, with_gather)

    def steps_pipelined(k_steps: int, with_gather: bool = True, record=None):
        """k_steps passes; step k + 1 is enqueued before step k's size is read back (and exchanged, N > 1)."""
        res = None
        launch(0)
        for k in range(k_steps):
            if k + 1 < k_steps:
                launch(k + 1)
            res = collect(k, with_gather)
            if record is not None:
                record.append(comps[k & 1].kernel_ms()["k_match"])
        return res

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-timed throughput ----
    for k in range(args.warmup):
        out_len, adler = step(k)
    drain()
    launches0 = sum(c.kernel_launches() for c in comps)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k_match_ms = []
    t0 = time.perf_counter()
    e0.record()
    out_len, adler = steps_pipelined(args.steps, record=k_match_ms)
    drain()  # every step's bytes have arrived on rank 0 before the clock stops
    e1.record()
    sync()
    t1 = time.perf_counter()
    ms = e0.elapsed_time(e1)
    launches = sum(c.kernel_launches() for c in comps) - launches0
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    kernel_ms = comps[(args.steps - 1) & 1].kernel_ms()
    timed_out = d_outs[(args.steps - 1) % NB]  # the buffer the last timed step wrote

    # ---- what the timed loop produced, checked outside the timed region ----
    import hashlib
    import zlib

    digest = hashlib.blake2b(host[hist:].tobytes(), digest_size=16).digest()
    roundtrip_ok, roundtrip_what = None, None
    if world == 1:
        timed_body = timed_out[:out_len].cpu().numpy().tobytes()
        timed_stream = b"\x78\x9c" + timed_body + int(adler).to_bytes(4, "big")
        roundtrip_ok = zlib.decompress(timed_stream) == host[hist:].tobytes()
        roundtrip_what = "zlib.decompress(header + the body the last timed step left in HBM + its Adler-32) == the corpus"
    else:
        digests = [None] * world
        dist.all_gather_object(digests, digest)
        if rank == 0:
            from deflate_compression_b200 import shard

            sz, b = last_gather["sizes"], last_gather["buf"]
            body = gather_bufs[b][:sum(sz)].cpu().numpy().tobytes()
            offs = np.concatenate([[0], np.cumsum(sz)])
            z = shard.stitch([body[int(offs[r]):int(offs[r + 1])] for r in range(world)], last_gather["adlers"], [SHARD_BYTES] * world)
            del body
            dec = zlib.decompressobj()
            ok, pos, h, got = True, 0, hashlib.blake2b(digest_size=16), 0
            for c0 in range(0, len(z), 8 << 20):
                out = dec.decompress(z[c0:c0 + (8 << 20)])
                view = memoryview(out)
                while len(view):
                    take = min(len(view), SHARD_BYTES - got)
                    h.update(view[:take])
                    got += take
                    view = view[take:]
                    if got == SHARD_BYTES:
                        ok = ok and pos < world and h.digest() == digests[pos]
                        pos, h, got = pos + 1, hashlib.blake2b(digest_size=16), 0
            roundtrip_ok = bool(ok and dec.eof and pos == world and got == 0)   # (zlib has checked the combined Adler-32)
            roundtrip_what = (f"the {world} bodies gathered on rank 0 by the last timed step, stitched with the combined Adler-32, "
                              "inflate with zlib to the shards (per-shard BLAKE2b digests from their ranks)")
            del z

    # compress-only (no gather) for N > 1
    ms_nogather = ms
    drain()
    if world > 1:
        sync()
        e0.record()
        steps_pipelined(args.steps, with_gather=False)
        drain()
        e1.record()
        sync()
        ms_nogather = e0.elapsed_time(e1)

    # ---- the match kernel's own launch duration, for the roofline: in the loop above two streams take the steps in
    # turn, so the events around a launch also bracket the time its CTAs wait for the SMs the other stream's kernels
    # hold.  A few steps on one stream, nothing else in flight, give the launch itself. ----
    sync()
    k_match_solo = []
    for _ in range(max(2, min(args.steps, 5))):
        step(0, with_gather=False)
        k_match_solo.append(comps[0].kernel_ms()["k_match"])
    kernel_ms_solo = comps[0].kernel_ms()  # every kernel of one step, nothing else in flight
    drain()
    sync()

    # ---- end to end through the host API: pinned host buffers, H2D + D2H inside ----
    h_in = torch.from_numpy(host[hist:]).pin_memory()
    h_out = torch.empty(cap + 16, dtype=torch.uint8).pin_memory()
    comp_h = D.Compressor(device=local)
    z_len = comp_h.compress_into(h_in.data_ptr(), n, h_out.data_ptr(), cap + 16)  # warm-up (allocations)
    comp_h.compress_into(h_in.data_ptr(), n, h_out.data_ptr(), cap + 16)
    sync()
    e2e_steps = max(2, min(args.steps, 5))
    te0 = time.perf_counter()
    for _ in range(e2e_steps):
        z_len = comp_h.compress_into(h_in.data_ptr(), n, h_out.data_ptr(), cap + 16)
    torch.cuda.synchronize()
    te = time.perf_counter() - te0
    e2e_stream = h_out[:z_len].numpy().tobytes()
    e2e_ok = zlib.decompress(e2e_stream) == host[hist:].tobytes() if rank == 0 and world == 1 else None
    if world == 1:
        # the device-timed leg and the end-to-end leg wrote the same body
        roundtrip_ok = bool(roundtrip_ok and e2e_stream[2:-4] == timed_body)
    comp_h.close()
    del h_out

    # ---- the reference entry point itself: deflate_compress(fd_in, fd_out, -1, 32768, 0) over memfds ----
    sync()
    os.environ["DFL_B200_DEVICE"] = str(local)  # the entry point has no device argument
    fd_gbs, fd_stream, fd_rc = e2e_fd_throughput(D, host[hist:])
    fd_ok = bool(fd_stream is not None and zlib.decompress(fd_stream) == host[hist:].tobytes()) if rank == 0 else None

    # ---- the way back: segment-parallel inflate of what the timed step wrote, device to
    # device, compared with the shard on the device.  Not part of `value`; reported beside it. ----
    out_len, adler = step(0, with_gather=False)
    d_off = torch.from_numpy(comp.seg_index().astype(np.int64)).to(dev)
    d_back = torch.zeros(hist + n, dtype=torch.uint8, device=dev)
    d_back[:hist].copy_(d_buf[:hist])
    nseg = (n + 32767) // 32768
    inf_ms = []
    for _ in range(3):
        adler_back = comp.inflate_device(d_out.data_ptr(), out_len, d_off.data_ptr(), nseg, d_back.data_ptr() + hist, n,
                                         hist, stream)
        k = comp.kernel_ms()
        inf_ms.append(k["k_inflate"] + k["k_inf_resolve"] + k["k_inf_finish"])
    inflate_ok = bool(adler_back == adler) and bool(torch.equal(d_back[hist:], d_buf[hist:]))
    inflate_rounds = comp.inflate_rounds()
    del d_back, d_off

    # ---- reduce over ranks: max time ----
    t = torch.tensor([ms, ms_nogather, te * 1e3, min(inf_ms), 0.0 if inflate_ok else 1.0, n / fd_gbs / 1e6 if fd_gbs else 1e12],
                     dtype=torch.float64, device=dev)
    tot = torch.tensor([float(n), float(out_len)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms, ms_nogather, te_ms, inf_best_ms, inf_bad, fd_ms = t.tolist()
    n_total, out_total = tot.tolist()

    peak, peak_src = measured_peak()
    del d_outs, d_out, gather_bufs, timed_out
    torch.cuda.empty_cache()
    others = None
    split = None
    if not args.no_other_workloads:
        if rank == 0:
            split = block_split_report(D, dev, local, d_buf, hist, n, stream, host[hist:hist + CPU_SAMPLE], n_total / out_total)
        others = other_workloads(D, comp, dev, rank, world, stream, dist, peak)

    if rank == 0:
        km = sum(k_match_solo) / len(k_match_solo)      # ms of k_match per launch (one launch per 256 MiB chunk), nothing else on the GPU
        km_loop = sum(k_match_ms) / len(k_match_ms)     # the same events inside the two-stream loop
        algo_bytes = n + out_len                        # 1 B read + 1/ratio B written per input byte
        achieved = algo_bytes / (km * 1e-3) / 1e9
        traffic = ncu_traffic(n)
        line = {
            "metric": "input GB/s compressed", "value": n_total * args.steps / (ms * 1e-3) / 1e9, "unit": "GB/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "256 MiB order-2 Markov text per GPU (BASELINE configs[1], seed 2, 32 MiB stripes)",
                       "window": WINDOW, "segment_bytes": 32768, "chain_limit": 0, "block_mode": "cheapest (all dynamic)",
                       "l2": "inputs larger than L2 (256 MiB per step)", "steps_overlap": "two contexts on two streams take the steps in turn; step k + 1 is enqueued while step k's size is read back", "sharding": f"contiguous, {world} rank(s)",
                       "gather": {"ipc": "rank 0's copy engine pulls every shard's bytes out of buffers opened through CUDA IPC (no SMs)",
                                  "nccl": "batched NCCL isend / irecv to rank 0", "none": "single GPU"}[gather_mode],
                       "cpu_binding": binding},
            "compression_ratio": n_total / out_total,
            "roundtrip_ok": roundtrip_ok,
            "roundtrip_check": roundtrip_what,
            "clocks": clocks,
            "e2e": {"value": n_total * e2e_steps / (te_ms * 1e-3) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": int(n),
                    "d2h_bytes_per_step": int(z_len), "steps": e2e_steps, "roundtrip_ok": e2e_ok,
                    "call": "dfl_b200_compress_host: pinned host in/out, H2D + D2H inside"},
            "e2e_fd": {"value": n_total / (fd_ms * 1e-3) / 1e9 if fd_ms < 1e11 else None, "unit": "GB/s", "roundtrip_ok": fd_ok,
                       "call": "deflate_compress(memfd_in, memfd_out, -1, 32768, 0): read(2)/pread(2) + H2D + kernels + D2H + "
                               "write(2) inside, best of 3" + ("" if not fd_rc else f" (error {fd_rc})")},
            "gpu_launches": int(launches),
            "kernel_ms_per_step": {k: v for k, v in kernel_ms_solo.items()},
            "kernel_ms_per_step_what": "CUDA events around each launch, one step on one stream right after the timed region (inside the "
                                       "timed two-stream loop the events also bracket waits for SMs: kernel_ms_in_timed_loop)",
            "kernel_ms_in_timed_loop": {k: v for k, v in kernel_ms.items() if v},
            "roofline": {"bound": "hbm", "kernel": "k_match", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(algo_bytes), "launch_ms": km,
                         "launch_ms_what": "CUDA events around the launch on its stream, average of a few steps run on one stream right after the timed region",
                         "launch_ms_in_timed_loop": km_loop,
                         "launch_ms_in_timed_loop_what": "the same events inside the timed two-stream loop: they also bracket the wait for SMs the other stream's kernels hold"},
        }
        line["inflate"] = {"value": n_total / (inf_best_ms * 1e-3) / 1e9, "unit": "GB/s of output (device-timed, k_inflate + "
                           "k_inf_resolve + k_inf_finish, best of 3)", "ms": inf_best_ms, "pointer_jumping_rounds": inflate_rounds,
                           "roundtrip_ok": inf_bad == 0.0}
        if others is not None:
            line["other_workloads"] = others
        if split is not None:
            line["block_split"] = split
        if world > 1:
            line["compress_only_value"] = n_total * args.steps / (ms_nogather * 1e-3) / 1e9
        # ---- CPU baseline beside it (rank 0, N = 1 only): the reference's single-threaded compressor ----
        if world == 1 and not args.no_cpu_baseline:
            from oracle import pyoracle as O

            sample = host[hist:hist + CPU_SAMPLE]
            kind, times, tokens = cpu_reference_throughput(sample, 1)
            cb = {"value": sample.size / times[0] / 1e9, "unit": "GB/s", "cores": 1, "kind": kind,
                  "host_cores": os.cpu_count(),
                  "sample": f"first {CPU_SAMPLE >> 20} MiB of the workload as independent 32 KiB slices "
                            f"(the largest input the reference survives), stats on; {tokens} tokens"}
            # ratio vs reference on the same sample: oracle stream (reference tokens + RFC 1951 encoder) vs GPU stream
            t0 = time.perf_counter()
            ref_size = len(O.zlib_compress(sample, O.params()))
            cb["oracle_port_gbs"] = sample.size / (time.perf_counter() - t0) / 1e9
            gpu_size = len(D.Compressor(device=local).compress(sample))
            line["ratio_vs_reference"] = {"sample_bytes": int(sample.size), "reference_bytes": ref_size,
                                          "gpu_bytes": gpu_size, "gpu_over_reference": gpu_size / ref_size,
                                          "zlib6_bytes": len(zlib.compress(sample.tobytes(), 6))}
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-workloads", action="store_true", help="skip the passes over configs 3, 4 and 5")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
