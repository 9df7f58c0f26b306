#!/usr/bin/env python
"""bench.py — headline benchmark of compu-b200 (contract: see the task prompt; metric from BASELINE.json).

Workload at N=1 (BASELINE.json configs[1]): batched inflate of 65,536 independent zlib streams of 64 KiB synthetic
Markov text (compressed by madler zlib 1.3 at level 6, the "reference-encoded" input) on one B200. With N>1 each rank
inflates its own 65,536-stream shard (weak scaling, no data-path collective: streams are independent).

  value        uncompressed GB/s, inputs and outputs resident in HBM (CUDA events around K launches)
  e2e          the same metric through the host-memory C-ABI call cz_inflate_batch: pinned host buffers, H2D of the
               compressed streams and D2H of the decoded bytes inside the timed region
  roofline     (U + C) bytes per launch / average launch time, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline the oracle (compu's glue over zlib 1.3 — stand-in for compu zlib-ng + rayon) on the host cores

`--impl reference` times that CPU path alone (all host threads) and prints the same JSON line with "impl": "reference".
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STREAM_BYTES = 65536
METRIC = "inflate_uncompressed_GBps"
UNIT = "GB/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=65536, help="streams per GPU (cfg2: 65536)")
    ap.add_argument("--seed", type=int, default=20261018)
    ap.add_argument("--workload", default="inflate", choices=["inflate", "deflate"],
                    help="inflate = cfg2 (the headline, default); deflate = cfg3 (chunked deflate L6 of ONE stream, full-flush segments)")
    ap.add_argument("--mib", type=int, default=4096, help="deflate workload: MiB of input per GPU (cfg3: 4096)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


def compress_streams(plain, n, threads):
    """zlib 1.3 level 6, windowBits 15, one independent stream per 64 KiB (what compu's Encoder would emit per item)."""
    mv = memoryview(plain)

    def work(rng):
        return [zlib.compress(mv[i * STREAM_BYTES:(i + 1) * STREAM_BYTES], 6) for i in rng]

    step = max(1, n // (threads * 8))
    ranges = [range(i, min(n, i + step)) for i in range(0, n, step)]
    with ThreadPoolExecutor(max_workers=threads) as ex:
        parts = list(ex.map(work, ranges))
    streams = [s for p in parts for s in p]
    return streams


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def pin_to_gpu_numa_node(index):
    """Runs this process on the CPUs NVML reports as local to GPU `index`, so that the pinned host buffers of the end-to-end
    leg are allocated on the GPU's NUMA node (what a deployment does with numactl). Best effort; returns a short note."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w in range(words) for b in range(64) if (mask[w] >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus and len(cpus) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, cpus)
            return "pinned to %d CPUs local to GPU %d" % (len(cpus), index)
        return "all CPUs are local to GPU %d" % index
    except Exception as e:  # no NVML, no permission: run as is
        return "unchanged (%s)" % type(e).__name__


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index),
                 "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for k, nme in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class CpuInflate:
    """Oracle on the host cores (the ONLY place bench.py executes oracle/). Buffers are prepared once, run() is timed."""

    def __init__(self, streams_sample, threads):
        import oracle
        self.L = oracle.lib()
        self.threads = threads
        n = self.n = len(streams_sample)
        self.offs = np.zeros(n + 1, dtype=np.uint64)
        self.offs[1:] = np.cumsum([len(s) for s in streams_sample], dtype=np.uint64)
        self.inbuf = np.frombuffer(b"".join(streams_sample) + b"\0" * 16, dtype=np.uint8)
        self.out_off = (np.arange(n + 1, dtype=np.uint64) * STREAM_BYTES)
        self.out = np.empty(n * STREAM_BYTES + 16, dtype=np.uint8)
        self.lens = np.zeros(n, dtype=np.uint64)
        self.st = np.zeros(n, dtype=np.int32)

    def run(self):
        t0 = time.perf_counter()
        bad = self.L.oz_inflate_batch(self.n, _p(self.inbuf), _p(self.offs), _p(self.out), _p(self.out_off), _p(self.lens),
                                      _p(self.st), 15, self.threads)
        dt = time.perf_counter() - t0
        assert bad == 0
        return dt


def cpu_inflate_leg(streams_sample, threads, repeats=3):
    """Returns (GB/s uncompressed, seconds) — best of `repeats`."""
    c = CpuInflate(streams_sample, threads)
    best = min(c.run() for _ in range(repeats))
    return c.n * STREAM_BYTES / best / 1e9, best


def host_synth(n, seed):
    """Same bytes as the device generator (cz_synth_fill_host), for the --impl reference arm when no GPU work is wanted."""
    from compu_b200 import _lib
    L = _lib.lib()
    corpus = np.frombuffer(open(os.path.join(ROOT, "tests", "golden", "alice29.txt"), "rb").read(), dtype=np.uint8)
    model = np.zeros(int(L.cz_synth_model_bytes()), dtype=np.uint8)
    _lib.check(L.cz_synth_build_model(_p(corpus), len(corpus), _p(model)), "cz_synth_build_model")
    offs = np.arange(n + 1, dtype=np.uint64) * STREAM_BYTES
    out = np.empty(n * STREAM_BYTES, dtype=np.uint8)
    _lib.check(L.cz_synth_fill_host(0, seed, n, _p(out), _p(offs), _p(model)), "cz_synth_fill_host")
    return out, model



def deflate_traffic(mib):
    """DRAM bytes of the dominant kernel (match search) per launch, from the committed ncu capture (profiles/traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return int(json.load(f)["deflate_match_dram_bytes_per_gib"] * mib / 1024)
    except Exception:
        return None


def cpu_threads():
    """Host threads for the CPU arm: every CPU this process may run on. (torchrun exports OMP_NUM_THREADS=1 to its workers;
    the oracle takes its thread count as an argument, so that default does not shrink the baseline to one core.)"""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    e = os.environ.get("CZ_CPU_THREADS")
    return max(1, int(e)) if e else max(1, n)


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path for this metric (oracle port over zlib 1.3, all host threads)."""
    if rank != 0:
        return
    import oracle
    threads = cpu_threads()
    # bounded sample of the same workload: enough streams for ~1 s per step on this host
    n = min(args.streams, max(256, 512 * threads))
    plain, _ = host_synth(n, args.seed)
    streams = compress_streams(plain, n, threads)
    c = CpuInflate(streams, threads)
    for _ in range(args.warmup):
        c.run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c.run()
    dt = time.perf_counter() - t0
    val = args.steps * n * STREAM_BYTES / dt / 1e9
    sample = "%d of %d streams x 64 KiB per step (same generator, same seed)" % (n, args.streams)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "cfg2: batched inflate of independent 64 KiB zlib streams (Markov text, zlib 1.3 L6)",
                   "streams_per_step": n, "stream_bytes": STREAM_BYTES, "window_bits": 15,
                   "codec": "compu glue restated in C over madler zlib 1.3 (stand-in for compu zlib-ng + rayon)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
# cfg3: chunked deflate (level 6) of ONE synthetic stream into one valid zlib stream made of full-flush segments.
SEG_BYTES = 1 << 20
DEFLATE_METRIC = "deflate_uncompressed_GBps"


class CpuDeflate:
    """Oracle encoder on the host cores: independent 1 MiB pieces across threads (the pigz-like split of SURVEY §8d)."""

    def __init__(self, plain, threads):
        import oracle
        self.L = oracle.lib()
        self.threads = threads
        n = self.n = len(plain) // SEG_BYTES
        self.plain = plain
        self.in_off = np.arange(n + 1, dtype=np.uint64) * SEG_BYTES
        bound = SEG_BYTES + SEG_BYTES // 8 + 1024
        self.out_off = np.arange(n + 1, dtype=np.uint64) * bound
        self.out = np.empty(n * bound + 16, dtype=np.uint8)
        self.lens = np.zeros(n, dtype=np.uint64)
        self.st = np.zeros(n, dtype=np.int32)

    def run(self):
        t0 = time.perf_counter()
        bad = self.L.oz_deflate_batch(self.n, _p(self.plain), _p(self.in_off), _p(self.out), _p(self.out_off), _p(self.lens),
                                      _p(self.st), 6, -15, 8, 0, self.threads)
        dt = time.perf_counter() - t0
        assert bad == 0
        return dt


def host_synth_bytes(nbytes, seed):
    from compu_b200 import _lib
    L = _lib.lib()
    corpus = np.frombuffer(open(os.path.join(ROOT, "tests", "golden", "alice29.txt"), "rb").read(), dtype=np.uint8)
    model = np.zeros(int(L.cz_synth_model_bytes()), dtype=np.uint8)
    _lib.check(L.cz_synth_build_model(_p(corpus), len(corpus), _p(model)), "cz_synth_build_model")
    n = nbytes // STREAM_BYTES
    offs = np.arange(n + 1, dtype=np.uint64) * STREAM_BYTES
    out = np.empty(n * STREAM_BYTES + 16, dtype=np.uint8)
    _lib.check(L.cz_synth_fill_host(0, seed, n, _p(out), _p(offs), _p(model)), "cz_synth_fill_host")
    return out


def run_reference_deflate(args, rank):
    if rank != 0:
        return
    import oracle
    threads = cpu_threads()
    sample_mib = min(args.mib, max(16, 8 * threads))  # ~1-2 s of CPU work per step at level 6
    plain = host_synth_bytes(sample_mib << 20, args.seed)
    c = CpuDeflate(plain, threads)
    for _ in range(args.warmup):
        c.run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c.run()
    dt = time.perf_counter() - t0
    val = args.steps * (sample_mib << 20) / dt / 1e9
    sample = "%d MiB of the %d MiB stream per step, as independent 1 MiB pieces (same generator, same seed)" % (sample_mib, args.mib)
    print(json.dumps({
        "impl": "reference", "metric": DEFLATE_METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "cfg3: chunked deflate level 6 of one synthetic stream (Markov text), full-flush segments",
                   "segment_bytes": SEG_BYTES, "level": 6,
                   "codec": "compu glue restated in C over madler zlib 1.3 (stand-in for compu zlib-ng + rayon)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


def run_deflate(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from compu_b200 import _lib

    L = _lib.lib()
    _lib.require_device()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    U = args.mib << 20
    nseg = U // SEG_BYTES
    stream = torch.cuda.current_stream()
    sp = ctypes.c_void_p(stream.cuda_stream)
    corpus = np.frombuffer(open(os.path.join(ROOT, "tests", "golden", "alice29.txt"), "rb").read(), dtype=np.uint8)
    model = np.zeros(int(L.cz_synth_model_bytes()), dtype=np.uint8)
    _lib.check(L.cz_synth_build_model(_p(corpus), len(corpus), _p(model)), "cz_synth_build_model")
    d_model = torch.from_numpy(model).to(dev)
    d_in = torch.empty(U + 16, dtype=torch.uint8, device=dev)
    gu = U // STREAM_BYTES
    d_goff = torch.arange(gu + 1, dtype=torch.int64, device=dev) * STREAM_BYTES
    _lib.check(L.cz_synth_fill_device(sp, 0, args.seed + rank * gu, gu, d_in.data_ptr(), d_goff.data_ptr(), d_model.data_ptr()), "synth")
    d_off = torch.arange(nseg + 1, dtype=torch.int64, device=dev) * SEG_BYTES
    bound = int(L.cz_deflate_segment_bound(SEG_BYTES))
    d_out = torch.empty(nseg * bound + 16, dtype=torch.uint8, device=dev)
    d_out_off = torch.arange(nseg + 1, dtype=torch.int64, device=dev) * bound
    d_lens = torch.zeros(nseg, dtype=torch.int64, device=dev)
    d_st = torch.zeros(nseg, dtype=torch.int32, device=dev)
    d_chk = torch.zeros(2 * nseg, dtype=torch.int32, device=dev)
    ws = int(L.cz_deflate_workspace_bytes(nseg, U))
    d_ws = torch.empty(ws, dtype=torch.uint8, device=dev)

    def step():
        rc = L.cz_deflate_segments_device(sp, nseg, d_in.data_ptr(), d_off.data_ptr(), U, d_out.data_ptr(), d_out_off.data_ptr(),
                                          d_lens.data_ptr(), d_st.data_ptr(), d_chk.data_ptr(), 6, 0, d_ws.data_ptr(), ws)
        _lib.check(rc, "cz_deflate_segments_device")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # correctness gate: every segment Finished, a sample of segments inflates (zlib) to the plaintext
    step()
    torch.cuda.synchronize()
    assert bool((d_st == 2).all())
    C = int(d_lens.sum().item())
    lens = d_lens.cpu().numpy()
    for i in list(range(0, nseg, max(1, nseg // 8)))[:8]:
        piece = d_out[i * bound:i * bound + int(lens[i])].cpu().numpy().tobytes()
        assert zlib.decompressobj(-15).decompress(piece) == d_in[i * SEG_BYTES:(i + 1) * SEG_BYTES].cpu().numpy().tobytes()
    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_step = ms / args.steps
    value = world * U / (ms_step * 1e-3) / 1e9

    e2e = None
    plain = None
    if not args.no_e2e:
        cap = int(L.cz_deflate_bound(U, 15, SEG_BYTES))
        h_in_p = L.cz_host_alloc(U + 16)
        h_out_p = L.cz_host_alloc(cap + 16)
        assert h_in_p and h_out_p, _lib.last_error()
        h_in = np.ctypeslib.as_array(ctypes.cast(h_in_p, ctypes.POINTER(ctypes.c_uint8)), shape=(U + 16,))
        h_out = np.ctypeslib.as_array(ctypes.cast(h_out_p, ctypes.POINTER(ctypes.c_uint8)), shape=(cap + 16,))
        h_in[:U] = d_in[:U].cpu().numpy()
        plain = h_in
        out_len = ctypes.c_uint64(0)
        nsegs = ctypes.c_uint64(0)

        def e2e_step():
            rc = L.cz_deflate_segmented(ctypes.c_void_p(h_in_p), U, ctypes.c_void_p(h_out_p), cap, ctypes.byref(out_len), 6, 15, 0,
                                        SEG_BYTES, 1 << local_rank, None, 0, ctypes.byref(nsegs))
            _lib.check(rc, "cz_deflate_segmented")

        for _ in range(max(1, min(args.warmup, 2))):
            e2e_step()
        # ONE valid zlib stream: the host zlib inflates a prefix of it back to the plaintext
        d = zlib.decompressobj(15)
        head = d.decompress(h_out[:min(out_len.value, 8 << 20)].tobytes())
        assert head == h_in[:len(head)].tobytes() and len(head) > 0
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * U * args.steps / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(U),
               "d2h_bytes_per_step": int(out_len.value), "ms_per_step": dt / args.steps * 1e3}

    if rank == 0:
        peak, peak_src = load_peaks()
        achieved = (U + C) / (ms_step * 1e-3) / 1e9
        cpu = None
        if not args.no_cpu and world == 1:
            import oracle
            cthreads = cpu_threads()
            sample_mib = min(args.mib, max(16, 8 * cthreads))
            src = plain if plain is not None else d_in[:sample_mib << 20].cpu().numpy()
            c = CpuDeflate(np.ascontiguousarray(src[:sample_mib << 20]), cthreads)
            secs = min(c.run() for _ in range(2))
            cpu = {"value": (sample_mib << 20) / secs / 1e9, "unit": UNIT, "cores": cthreads, "kind": "port",
                   "sample": "first %d MiB of the stream as independent 1 MiB pieces, best of 2 (%.2f s each), level 6; zlib 1.3 "
                             "stands in for zlib-ng; CPU ratio %.3f" % (sample_mib, secs, (sample_mib << 20) / float(c.lens.sum()))}
        print(json.dumps({
            "metric": DEFLATE_METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": "cfg3: chunked deflate level 6 of one %d MiB synthetic stream (Markov text) into one valid zlib "
                                   "stream of full-flush segments" % args.mib, "segment_bytes": SEG_BYTES, "level": 6,
                       "ratio": U / C, "l2": "input per step (%.2f GB) exceeds the 126 MB L2; no flush needed" % (U / 1e9)},
            "e2e": e2e, "gpu_launches": 11 * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": deflate_traffic(args.mib), "peak_source": peak_src, "algorithmic_bytes_per_launch": int(U + C),
                         "note": "traffic = DRAM bytes of the dominant kernel alone (ncu); achieved = whole kernel chain of one deflate call; the dominant kernel deflate_match_sweep_kernel (53 of ~80 ms per "
                                 "GiB) is instruction-bound integer code (L1 hit rate 99 %, 79 % issue utilisation at 12.8 active "
                                 "lanes): profiles/r1_deflate_match_sweep_ncu.md, profiles/r1_deflate_4gib_launches.csv"},
            "cpu_baseline": cpu, "clocks": clocks}))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload == "deflate":
        if args.impl == "reference":
            run_reference_deflate(args, rank)
        else:
            run_deflate(args, rank, local_rank, world)
        return
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from compu_b200 import _lib

    L = _lib.lib()
    _lib.require_device()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = pin_to_gpu_numa_node(local_rank)  # before any pinned allocation: page-locked buffers land on the GPU's node
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    n = args.streams
    U = n * STREAM_BYTES
    seed = args.seed + rank * n  # every rank inflates different streams
    stream = torch.cuda.current_stream()
    sp = ctypes.c_void_p(stream.cuda_stream)

    # ---- synthetic plaintext on the device (Markov text from the alice29 model), then reference-encode on the host
    corpus = np.frombuffer(open(os.path.join(ROOT, "tests", "golden", "alice29.txt"), "rb").read(), dtype=np.uint8)
    model = np.zeros(int(L.cz_synth_model_bytes()), dtype=np.uint8)
    _lib.check(L.cz_synth_build_model(_p(corpus), len(corpus), _p(model)), "cz_synth_build_model")
    d_model = torch.from_numpy(model).to(dev)
    d_plain = torch.empty(U, dtype=torch.uint8, device=dev)
    d_out_off = (torch.arange(n + 1, dtype=torch.int64, device=dev) * STREAM_BYTES)
    t0 = time.perf_counter()
    _lib.check(L.cz_synth_fill_device(sp, 0, seed, n, d_plain.data_ptr(), d_out_off.data_ptr(), d_model.data_ptr()), "synth")
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t0
    plain = d_plain.cpu().numpy()
    threads = cpu_threads()
    t0 = time.perf_counter()
    streams = compress_streams(plain, n, max(1, threads // max(1, world)))
    t_comp = time.perf_counter() - t0
    C = sum(len(s) for s in streams)
    in_off = np.zeros(n + 1, dtype=np.int64)
    in_off[1:] = np.cumsum([len(s) for s in streams])

    # pinned host buffers (the pinned-buffer management of the boundary: cz_host_alloc)
    h_in_p = L.cz_host_alloc(C + 16)
    h_out_p = L.cz_host_alloc(U + 16)
    assert h_in_p and h_out_p, _lib.last_error()
    h_in = np.ctypeslib.as_array(ctypes.cast(h_in_p, ctypes.POINTER(ctypes.c_uint8)), shape=(C + 16,))
    h_out = np.ctypeslib.as_array(ctypes.cast(h_out_p, ctypes.POINTER(ctypes.c_uint8)), shape=(U + 16,))
    pos = 0
    for s in streams:
        h_in[pos:pos + len(s)] = np.frombuffer(s, dtype=np.uint8)
        pos += len(s)

    d_in = torch.from_numpy(h_in).to(dev)
    d_in_off = torch.from_numpy(in_off).to(dev)
    d_out = torch.empty(U + 16, dtype=torch.uint8, device=dev)
    d_lens = torch.zeros(n, dtype=torch.int64, device=dev)
    d_stat = torch.zeros(n, dtype=torch.int32, device=dev)
    ws_bytes = int(L.cz_inflate_workspace_bytes(n, U))
    d_ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)

    def step():
        rc = L.cz_inflate_batch_device(sp, n, d_in.data_ptr(), d_in_off.data_ptr(), d_out.data_ptr(), d_out_off.data_ptr(),
                                       U, d_lens.data_ptr(), d_stat.data_ptr(), None, 15, d_ws.data_ptr(), d_ws.numel())
        _lib.check(rc, "cz_inflate_batch_device")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- correctness gate before any timing: every stream Finished and bytes identical to the plaintext
    step()
    torch.cuda.synchronize()
    assert bool((d_stat == 2).all()), "not every stream finished: %s" % torch.unique(d_stat).tolist()
    assert bool((d_lens == STREAM_BYTES).all())
    assert torch.equal(d_out[:U], d_plain), "inflated bytes differ from the plaintext"

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    L.cz_profile_enable(1)  # CUDA events around each of the two kernels, on the launching stream, inside the timed region
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    L.cz_profile_enable(0)
    ka, kb = ctypes.c_double(0), ctypes.c_double(0)
    nprof = L.cz_profile_read(ctypes.byref(ka), ctypes.byref(kb))
    kernel_ms = {"inflate_tok_kernel": ka.value / max(1, nprof), "inflate_lz_kernel": kb.value / max(1, nprof)} if nprof else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_step = ms / args.steps
    value = world * U / (ms_step * 1e-3) / 1e9

    # ---- end to end through the host-memory C ABI (pinned host buffers, H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        h_in_off = in_off.astype(np.uint64)
        h_out_off = (np.arange(n + 1, dtype=np.uint64) * STREAM_BYTES)
        h_lens = np.zeros(n, dtype=np.uint64)
        h_stat = np.zeros(n, dtype=np.int32)

        def e2e_step():
            rc = L.cz_inflate_batch(n, ctypes.c_void_p(h_in_p), _p(h_in_off), ctypes.c_void_p(h_out_p), _p(h_out_off), _p(h_lens),
                                    _p(h_stat), None, 15, 1 << local_rank)
            _lib.check(rc, "cz_inflate_batch")

        for _ in range(max(1, min(args.warmup, 2))):
            e2e_step()
        assert (h_stat == 2).all()
        k = np.random.default_rng(1).integers(0, n, 64)
        for i in k:
            assert (h_out[i * STREAM_BYTES:(i + 1) * STREAM_BYTES] == plain[i * STREAM_BYTES:(i + 1) * STREAM_BYTES]).all()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * U * args.steps / dt / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": int(C + 16 * (n + 1)), "d2h_bytes_per_step": int(U + 12 * n),
               "ms_per_step": dt / args.steps * 1e3}

    if rank == 0:
        peak, peak_src = load_peaks()
        k_ms = (kernel_ms["inflate_tok_kernel"] + kernel_ms["inflate_lz_kernel"]) if kernel_ms else ms_step
        achieved = (U + C) / (k_ms * 1e-3) / 1e9
        cpu = None
        if not args.no_cpu and world == 1:
            import oracle
            cthreads = cpu_threads()
            ns = min(n, max(256, 512 * cthreads))
            v, secs = cpu_inflate_leg(streams[:ns], cthreads)
            cpu = {"value": v, "unit": UNIT, "cores": cthreads, "kind": "port",
                   "sample": "first %d of %d streams, best of 3 (%.2f s each); zlib 1.3 stands in for zlib-ng" % (ns, n, secs)}
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get("inflate_cfg2_dram_bytes_per_launch")
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": "cfg2: batched inflate of 65,536 independent 64 KiB zlib streams (Markov text, zlib 1.3 L6)",
                       "streams_per_gpu": n, "stream_bytes": STREAM_BYTES, "window_bits": 15, "ratio": U / C,
                       "l2": "inputs+outputs per step (%.2f GB) exceed the 126 MB L2; no flush needed" % ((U + C) / 1e9),
                       "inflate_cfg": os.environ.get("CZ_INFLATE_CFG", "default"), "host_numa": numa,
                       "setup_s": {"synth_gpu": round(t_gen, 3), "zlib_compress_host": round(t_comp, 2)}},
            "e2e": e2e,
            "gpu_launches": 2 * args.steps,  # per step: inflate_tok_kernel + inflate_lz_kernel (plus a 256-byte memset node)
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(U + C),
                         "launch": "one inflate = inflate_tok_kernel + inflate_lz_kernel (the dominant one); achieved = (U + C) / "
                                   "(sum of both kernels' CUDA-event durations per step)",
                         "kernel_ms": kernel_ms},
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        print(json.dumps(line))
    L.cz_host_free(ctypes.c_void_p(h_in_p))
    L.cz_host_free(ctypes.c_void_p(h_out_p))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
