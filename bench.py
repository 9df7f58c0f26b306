#!/usr/bin/env python
"""bench.py — headline benchmark of compu-b200 (contract: see the task prompt; metric from BASELINE.json).

ONE JSON line. The headline (`metric`, `value`, `e2e`, `roofline`, `cpu_baseline`) is BASELINE.json configs[1]: batched
inflate of 65,536 independent zlib streams of 64 KiB synthetic Markov text (encoded by madler zlib 1.3 at level 6, the
"reference-encoded" input) on one B200; with N>1 every rank inflates its own 65,536-stream shard (weak scaling, no data-path
collective: streams are independent). The same line carries the other halves of BASELINE.json's metric as sub-records:

  "deflate"      cfg3: chunked deflate level 6 of ONE 4 GiB stream into one valid zlib stream of full-flush segments
                 (device-resident through cz_deflate_segments_device, end to end through cz_deflate_segmented), the
                 compression ratio and the size relative to zlib 1.3 level 6 on the same bytes
  "gzip_cfg4"    cfg4: gzip encode + decode of 1 GiB mixed-entropy data, CRC-32 folded from per-segment values and checked
                 against the CRC of the whole buffer
  "partitioned"  cfg5 shape: rank 0 drives ALL N GPUs of the job from one process through the product's partitioner
                 (cz_deflate_batch / cz_inflate_batch with devices_mask = (1 << N) - 1) on streams of 4 KiB - 16 MiB

  value        uncompressed GB/s, inputs and outputs resident in HBM (CUDA events around K launches)
  e2e          the same metric through the host-memory C-ABI call: pinned host buffers, H2D and D2H inside the timed region;
               e2e.link_floor_ms = the same byte counts copied H2D + D2H concurrently with nothing else (all ranks at once)
  roofline     (U + C) bytes per launch / summed CUDA-event durations of the launch's kernels, against MEASURED_PEAKS.json
  cpu_baseline the oracle (compu's glue over zlib 1.3 — stand-in for compu zlib-ng + rayon) on the host cores

`--impl reference` times that CPU path alone (all host threads, rank 0 only) and prints the same JSON line with
"impl": "reference". It loads only oracle/ libraries (the data generator included), never the product library.
"""
import argparse
import ctypes
import glob
import hashlib
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
SEG_BYTES = 1 << 20
METRIC = "inflate_uncompressed_GBps"
UNIT = "GB/s"
CODEC_NOTE = "compu glue restated in C over madler zlib 1.3 (stand-in for compu zlib-ng + rayon; zlib-ng is not in this image)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=65536, help="streams per GPU (cfg2: 65536)")
    ap.add_argument("--seed", type=int, default=20261018)
    ap.add_argument("--workload", default="all", choices=["all", "inflate", "deflate", "gzip", "partitioned"],
                    help="all = the headline (cfg2 inflate) plus the deflate / gzip_cfg4 / partitioned sub-records (default)")
    ap.add_argument("--mib", type=int, default=4096, help="deflate sub-record: MiB of input per GPU (cfg3: 4096)")
    ap.add_argument("--gzip-mib", type=int, default=1024, help="gzip sub-record: MiB (cfg4: 1024)")
    ap.add_argument("--part-mib-per-gpu", type=int, default=2048, help="partitioned sub-record: MiB per GPU")
    ap.add_argument("--sub-steps", type=int, default=3, help="timed steps of each sub-record")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def kernel_source_hash():
    """Identifies the kernel sources a profile belongs to: sha256 over compu_b200/csrc/*.cuh (sorted) — every __global__
    function of the library lives in a .cuh; the .cu / .h files are the host side of the ABI."""
    h = hashlib.sha256()
    for f in sorted(glob.glob(os.path.join(ROOT, "compu_b200", "csrc", "*.cuh"))):
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def traffic_entry(key):
    """DRAM bytes per launch from the committed `ncu --set full` capture (profiles/traffic.json), only if that capture was
    taken on the kernel sources being benchmarked (its `src_hash` must match); otherwise null + why."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
    except Exception:
        return None, "profiles/traffic.json missing"
    if t.get("src_hash") != kernel_source_hash():
        return None, "profiles/traffic.json was captured on other kernel sources (src_hash %s != %s)" % (t.get("src_hash"), kernel_source_hash())
    src_key = {"deflate_cfg3_dram_bytes_per_launch": "deflate_cfg3_source"}.get(key, "source")
    return t.get(key), t.get(src_key, "profiles/traffic.json")


def cpu_threads():
    """Host threads for the CPU arm: every CPU this process may run on. (torchrun exports OMP_NUM_THREADS=1 to its workers;
    the oracle takes its thread count as an argument, so that default does not shrink the baseline to one core.)"""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    e = os.environ.get("CZ_CPU_THREADS")
    return max(1, int(e)) if e else max(1, n)


def mem_available_bytes():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) * 1024
    except Exception:
        pass
    return 64 << 30


def compress_streams(plain, n, threads, unit=STREAM_BYTES, level=6):
    """zlib 1.3 level 6, windowBits 15, one independent stream per unit (what compu's Encoder would emit per item)."""
    mv = memoryview(plain)

    def work(rng):
        return [zlib.compress(mv[i * unit:(i + 1) * unit], level) for i in rng]

    step = max(1, n // (threads * 8))
    ranges = [range(i, min(n, i + step)) for i in range(0, n, step)]
    with ThreadPoolExecutor(max_workers=threads) as ex:
        parts = list(ex.map(work, ranges))
    return [s for p in parts for s in p]


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


# ---------------------------------------------------------------------------------------------------------------------
# CPU legs (the ONLY places bench.py executes oracle/). Buffers are prepared once, run() is timed.
class CpuInflate:
    def __init__(self, streams_sample, threads, unit=STREAM_BYTES, window_bits=15):
        import oracle
        self.L = oracle.lib()
        self.threads = threads
        self.wb = window_bits
        n = self.n = len(streams_sample)
        self.unit = unit
        self.offs = np.zeros(n + 1, dtype=np.uint64)
        self.offs[1:] = np.cumsum([len(s) for s in streams_sample], dtype=np.uint64)
        self.inbuf = np.frombuffer(b"".join(streams_sample) + b"\0" * 16, dtype=np.uint8)
        self.out_off = (np.arange(n + 1, dtype=np.uint64) * unit)
        self.out = np.empty(n * unit + 16, dtype=np.uint8)
        self.lens = np.zeros(n, dtype=np.uint64)
        self.st = np.zeros(n, dtype=np.int32)

    def run(self):
        t0 = time.perf_counter()
        bad = self.L.oz_inflate_batch(self.n, _p(self.inbuf), _p(self.offs), _p(self.out), _p(self.out_off), _p(self.lens),
                                      _p(self.st), self.wb, self.threads)
        dt = time.perf_counter() - t0
        assert bad == 0
        return dt


class CpuDeflate:
    """Oracle encoder on the host cores: independent 1 MiB pieces across threads (the pigz-like split of SURVEY §8d)."""

    def __init__(self, plain, threads, window_bits=-15):
        import oracle
        self.L = oracle.lib()
        self.threads = threads
        self.wb = window_bits
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
                                      _p(self.st), 6, self.wb, 8, 0, self.threads)
        dt = time.perf_counter() - t0
        assert bad == 0
        return dt

    def streams(self):
        return [self.out[int(self.out_off[i]):int(self.out_off[i]) + int(self.lens[i])].tobytes() for i in range(self.n)]


def oracle_synth(kind, nbytes, seed, unit=STREAM_BYTES):
    """The generator of SURVEY §8d from oracle/libcompu_synth.so (bit-identical to the device generator; pinned by
    tests/test_synth.py), so that the reference arm does not load the product library."""
    import oracle
    S = oracle.synth()
    corpus = np.frombuffer(open(os.path.join(ROOT, "tests", "golden", "alice29.txt"), "rb").read(), dtype=np.uint8)
    model = np.zeros(int(S.oz_synth_model_bytes()), dtype=np.uint8)
    assert S.oz_synth_build_model(_p(corpus), len(corpus), _p(model)) == 0
    n = nbytes // unit
    offs = np.arange(n + 1, dtype=np.uint64) * unit
    out = np.empty(n * unit + 16, dtype=np.uint8)
    assert S.oz_synth_fill(kind, seed, n, _p(out), _p(offs), _p(model)) == 0
    return out


def run_reference(args, rank):
    """--impl reference: the reference's CPU path (oracle port over zlib 1.3, all host threads), rank 0 only. The headline
    is the arm's own workload, one GPU's shard per step: every one of the 65,536 streams, same generator, same seed."""
    if rank != 0:
        return
    threads = cpu_threads()
    n = args.streams
    t0 = time.perf_counter()
    plain = oracle_synth(0, n * STREAM_BYTES, args.seed)
    streams = compress_streams(plain, n, threads)
    t_setup = time.perf_counter() - t0
    c = CpuInflate(streams, threads)
    for _ in range(args.warmup):
        c.run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c.run()
    dt = time.perf_counter() - t0
    assert (c.out[:n * STREAM_BYTES] == plain[:n * STREAM_BYTES]).all()
    val = args.steps * n * STREAM_BYTES / dt / 1e9
    sample = "all %d streams x 64 KiB of one GPU's shard per step (same generator, same seed)" % n
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "cfg2: batched inflate of 65,536 independent 64 KiB zlib streams (Markov text, zlib 1.3 L6)",
                   "streams_per_step": n, "stream_bytes": STREAM_BYTES, "window_bits": 15, "codec": CODEC_NOTE,
                   "note": "per-byte throughput of the host: at N GPUs the GPU arm processes N such shards per step, the host "
                           "cores are the same", "setup_s": round(t_setup, 1)},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if args.workload in ("all", "deflate"):
        mib = min(args.mib, max(16, 8 * threads))
        d = CpuDeflate(np.ascontiguousarray(plain[:mib << 20]), threads)
        secs = min(d.run() for _ in range(2))
        line["deflate"] = {"metric": "deflate_uncompressed_GBps", "value": (mib << 20) / secs / 1e9, "unit": UNIT, "cores": threads,
                           "ratio": (mib << 20) / float(d.lens.sum()),
                           "sample": "%d MiB of the stream as independent 1 MiB pieces, best of 2, level 6" % mib}
    if args.workload in ("all", "gzip"):
        mib = min(args.gzip_mib, max(16, 8 * threads))
        mixed = oracle_synth(3, mib << 20, args.seed + 7)
        e = CpuDeflate(np.ascontiguousarray(mixed[:mib << 20]), threads, window_bits=31)
        t_enc = min(e.run() for _ in range(2))
        di = CpuInflate(e.streams(), threads, unit=SEG_BYTES, window_bits=31)
        t_dec = min(di.run() for _ in range(2))
        line["gzip_cfg4"] = {"metric": "gzip_roundtrip_uncompressed_GBps", "value": (mib << 20) / (t_enc + t_dec) / 1e9, "unit": UNIT,
                             "encode_GBps": (mib << 20) / t_enc / 1e9, "decode_GBps": (mib << 20) / t_dec / 1e9, "cores": threads,
                             "sample": "%d MiB mixed-entropy as independent 1 MiB gzip members, best of 2" % mib}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
class Ctx:
    """Per-process state of the GPU arm: library, device, process groups."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from compu_b200 import _lib
        self.torch, self.dist, self._lib, self.args = torch, dist, _lib, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.L = _lib.lib()
        _lib.require_device()
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.numa = pin_to_gpu_numa_node(self.local_rank)  # before any pinned allocation
        self.cpu_group = None
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
            self.cpu_group = dist.new_group(backend="gloo")  # host-side barriers that keep the GPUs idle
        self.stream = torch.cuda.current_stream()
        self.sp = ctypes.c_void_p(self.stream.cuda_stream)
        corpus = np.frombuffer(open(os.path.join(ROOT, "tests", "golden", "alice29.txt"), "rb").read(), dtype=np.uint8)
        self.model = np.zeros(int(self.L.cz_synth_model_bytes()), dtype=np.uint8)
        _lib.check(self.L.cz_synth_build_model(_p(corpus), len(corpus), _p(self.model)), "cz_synth_build_model")
        self.d_model = torch.from_numpy(self.model).to(self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def cpu_barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier(group=self.cpu_group)

    def max_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def synth_device(self, kind, nbytes, seed, dev=None):
        """nbytes of class `kind` on the device, as independently seeded 64 KiB pieces."""
        torch = self.torch
        dev = dev or self.dev
        n = nbytes // STREAM_BYTES
        d = torch.empty(n * STREAM_BYTES + 16, dtype=torch.uint8, device=dev)
        off = torch.arange(n + 1, dtype=torch.int64, device=dev) * STREAM_BYTES
        model = self.d_model if dev == self.dev else self.d_model.to(dev)
        with torch.cuda.device(dev):
            sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            self._lib.check(self.L.cz_synth_fill_device(sp, kind, seed, n, d.data_ptr(), off.data_ptr(), model.data_ptr()), "synth")
            torch.cuda.synchronize(dev)
        return d

    def host_alloc(self, nbytes):
        p = self.L.cz_host_alloc(nbytes)
        assert p, self._lib.last_error()
        return p, np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(nbytes,))

    def host_free(self, p):
        self.L.cz_host_free(ctypes.c_void_p(p))

    def link_floor_ms(self, h2d_bytes, d2h_bytes, repeats=3):
        """H2D of h2d_bytes and D2H of d2h_bytes, concurrently on two streams from pinned memory, nothing else running;
        every rank at once (the host path is shared by the GPUs of a box). Best of `repeats`, max over ranks."""
        torch = self.torch
        h_a = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
        h_b = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
        d_a = torch.empty(h2d_bytes, dtype=torch.uint8, device=self.dev)
        d_b = torch.empty(d2h_bytes, dtype=torch.uint8, device=self.dev)
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        best = None
        for it in range(repeats + 1):
            self.cpu_barrier()
            t0 = time.perf_counter()
            with torch.cuda.stream(s1):
                d_a.copy_(h_a, non_blocking=True)
            with torch.cuda.stream(s2):
                h_b.copy_(d_b, non_blocking=True)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) * 1e3
            dt = self.max_over_ranks(dt)
            if it > 0:
                best = dt if best is None else min(best, dt)
        del h_a, h_b, d_a, d_b
        return best


# ---------------------------------------------------------------------------------------------------------------------
def bench_inflate(C):
    """cfg2, the headline. Returns the fields of the JSON line that belong to it."""
    torch, L, _lib, args = C.torch, C.L, C._lib, C.args
    n = args.streams
    U = n * STREAM_BYTES
    seed = args.seed + C.rank * n  # every rank inflates different streams
    d_out_off = (torch.arange(n + 1, dtype=torch.int64, device=C.dev) * STREAM_BYTES)
    t0 = time.perf_counter()
    d_plain = C.synth_device(0, U, seed)[:U]
    t_gen = time.perf_counter() - t0
    plain = d_plain.cpu().numpy()
    threads = cpu_threads()
    t0 = time.perf_counter()
    streams = compress_streams(plain, n, max(1, threads // max(1, C.world)))
    t_comp = time.perf_counter() - t0
    Cb = sum(len(s) for s in streams)
    in_off = np.zeros(n + 1, dtype=np.int64)
    in_off[1:] = np.cumsum([len(s) for s in streams])

    # pinned host buffers (the pinned-buffer management of the boundary: cz_host_alloc)
    h_in_p, h_in = C.host_alloc(Cb + 16)
    h_out_p, h_out = C.host_alloc(U + 16)
    pos = 0
    for s in streams:
        h_in[pos:pos + len(s)] = np.frombuffer(s, dtype=np.uint8)
        pos += len(s)

    d_in = torch.from_numpy(h_in).to(C.dev)
    d_in_off = torch.from_numpy(in_off).to(C.dev)
    d_out = torch.empty(U + 16, dtype=torch.uint8, device=C.dev)
    d_lens = torch.zeros(n, dtype=torch.int64, device=C.dev)
    d_stat = torch.zeros(n, dtype=torch.int32, device=C.dev)
    ws_bytes = int(L.cz_inflate_workspace_bytes(n, U))
    d_ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=C.dev)

    def step():
        rc = L.cz_inflate_batch_device(C.sp, n, d_in.data_ptr(), d_in_off.data_ptr(), d_out.data_ptr(), d_out_off.data_ptr(),
                                       U, d_lens.data_ptr(), d_stat.data_ptr(), None, 15, d_ws.data_ptr(), d_ws.numel())
        _lib.check(rc, "cz_inflate_batch_device")

    # ---- correctness gate before any timing: every stream Finished and bytes identical to the plaintext
    step()
    torch.cuda.synchronize()
    assert bool((d_stat == 2).all()), "not every stream finished: %s" % torch.unique(d_stat).tolist()
    assert bool((d_lens == STREAM_BYTES).all())
    assert torch.equal(d_out[:U], d_plain), "inflated bytes differ from the plaintext"

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(C.local_rank)
    C.barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    L.cz_profile_enable(1)  # CUDA events around each of the two kernels, on the launching stream, inside the timed region
    launches0 = int(L.cz_launch_count())
    ev0.record(C.stream)
    for _ in range(args.steps):
        step()
    ev1.record(C.stream)
    C.barrier()
    launches = int(L.cz_launch_count()) - launches0
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    L.cz_profile_enable(0)
    ka, kb = ctypes.c_double(0), ctypes.c_double(0)
    nprof = L.cz_profile_read(ctypes.byref(ka), ctypes.byref(kb))
    kernel_ms = {"inflate_tok_kernel": ka.value / max(1, nprof), "inflate_lz_kernel": kb.value / max(1, nprof)} if nprof else None
    ms = C.max_over_ranks(ms)
    ms_step = ms / args.steps
    value = C.world * U / (ms_step * 1e-3) / 1e9

    # ---- end to end through the host-memory C ABI (pinned host buffers, H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        h_in_off = in_off.astype(np.uint64)
        h_out_off = (np.arange(n + 1, dtype=np.uint64) * STREAM_BYTES)
        h_lens = np.zeros(n, dtype=np.uint64)
        h_stat = np.zeros(n, dtype=np.int32)

        def e2e_step():
            rc = L.cz_inflate_batch(n, ctypes.c_void_p(h_in_p), _p(h_in_off), ctypes.c_void_p(h_out_p), _p(h_out_off), _p(h_lens),
                                    _p(h_stat), None, 15, 1 << C.local_rank)
            _lib.check(rc, "cz_inflate_batch")

        for _ in range(max(1, min(args.warmup, 2))):
            e2e_step()
        assert (h_stat == 2).all()
        assert (h_out[:U] == plain).all(), "end-to-end output differs from the plaintext"
        C.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        C.barrier()
        dt = C.max_over_ranks(time.perf_counter() - t0)
        h2d, d2h = int(Cb + 16 * (n + 1)), int(U + 12 * n)
        floor_ms = C.link_floor_ms(h2d, d2h)
        e2e_ms = dt / args.steps * 1e3
        e2e = {"value": C.world * U * args.steps / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": e2e_ms, "link_floor_ms": floor_ms, "frac_of_floor": floor_ms / e2e_ms if floor_ms else None,
               "link_floor": "H2D + D2H of exactly these byte counts, concurrently from pinned memory on all %d ranks at once, no "
                             "kernels; best of 3, max over ranks" % C.world}

    rec = None
    if C.rank == 0:
        peak, peak_src = load_peaks()
        k_ms = (kernel_ms["inflate_tok_kernel"] + kernel_ms["inflate_lz_kernel"]) if kernel_ms else ms_step
        achieved = (U + Cb) / (k_ms * 1e-3) / 1e9
        cpu = None
        if not args.no_cpu and C.world == 1:
            cthreads = cpu_threads()
            ns = min(n, max(256, 1024 * cthreads))
            c = CpuInflate(streams[:ns], cthreads)
            secs = min(c.run() for _ in range(3))
            cpu = {"value": ns * STREAM_BYTES / secs / 1e9, "unit": UNIT, "cores": cthreads, "kind": "port",
                   "sample": "first %d of %d streams, best of 3 (%.2f s each); zlib 1.3 stands in for zlib-ng" % (ns, n, secs)}
        traffic, traffic_src = traffic_entry("inflate_cfg2_dram_bytes_per_launch")
        dom = max(kernel_ms, key=kernel_ms.get) if kernel_ms else None
        rec = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": C.world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": "cfg2: batched inflate of 65,536 independent 64 KiB zlib streams (Markov text, zlib 1.3 L6)",
                       "streams_per_gpu": n, "stream_bytes": STREAM_BYTES, "window_bits": 15, "ratio": U / Cb,
                       "l2": "inputs+outputs per step (%.2f GB) exceed the 126 MB L2; no flush needed" % ((U + Cb) / 1e9),
                       "inflate_cfg": os.environ.get("CZ_INFLATE_CFG", "default"), "host_numa": C.numa,
                       "kernel_src_hash": kernel_source_hash(),
                       "setup_s": {"synth_gpu": round(t_gen, 3), "zlib_compress_host": round(t_comp, 2)}},
            "e2e": e2e,
            "gpu_launches": launches,  # counted by the library (cz_launch_count) over the timed region
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(U + Cb),
                         "launch": "one inflate launch = inflate_tok_kernel (Huffman decode into tokens) + inflate_lz_kernel (LZ77 "
                                   "resolution); achieved = (U + C) / (sum of both kernels' CUDA-event durations per step); the "
                                   "longer of the two is %s" % dom,
                         "kernel_ms": kernel_ms},
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
    C.host_free(h_in_p)
    C.host_free(h_out_p)
    return rec


# ---------------------------------------------------------------------------------------------------------------------
def bench_deflate(C):
    """cfg3: chunked deflate (level 6) of ONE synthetic stream into one valid zlib stream made of full-flush segments."""
    torch, L, _lib, args = C.torch, C.L, C._lib, C.args
    steps = max(1, args.sub_steps)
    U = args.mib << 20
    nseg = U // SEG_BYTES
    d_in = C.synth_device(0, U, args.seed + 1000003 * (C.rank + 1))
    d_off = torch.arange(nseg + 1, dtype=torch.int64, device=C.dev) * SEG_BYTES
    bound = int(L.cz_deflate_segment_bound(SEG_BYTES))
    d_out = torch.empty(nseg * bound + 16, dtype=torch.uint8, device=C.dev)
    d_out_off = torch.arange(nseg + 1, dtype=torch.int64, device=C.dev) * bound
    d_lens = torch.zeros(nseg, dtype=torch.int64, device=C.dev)
    d_st = torch.zeros(nseg, dtype=torch.int32, device=C.dev)
    d_chk = torch.zeros(2 * nseg, dtype=torch.int32, device=C.dev)
    ws = int(L.cz_deflate_workspace_bytes(nseg, U))
    d_ws = torch.empty(ws, dtype=torch.uint8, device=C.dev)

    def step():
        rc = L.cz_deflate_segments_device(C.sp, nseg, d_in.data_ptr(), d_off.data_ptr(), U, d_out.data_ptr(), d_out_off.data_ptr(),
                                          d_lens.data_ptr(), d_st.data_ptr(), d_chk.data_ptr(), 6, 0, d_ws.data_ptr(), ws)
        _lib.check(rc, "cz_deflate_segments_device")

    # correctness gate: every segment Finished, a sample of segments inflates (zlib) to the plaintext
    step()
    torch.cuda.synchronize()
    assert bool((d_st == 2).all())
    Cb = int(d_lens.sum().item())
    lens = d_lens.cpu().numpy()
    for i in list(range(0, nseg, max(1, nseg // 8)))[:8]:
        piece = d_out[i * bound:i * bound + int(lens[i])].cpu().numpy().tobytes()
        assert zlib.decompressobj(-15).decompress(piece) == d_in[i * SEG_BYTES:(i + 1) * SEG_BYTES].cpu().numpy().tobytes()
    for _ in range(2):
        step()
    C.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    L.cz_profile_enable(1)
    launches0 = int(L.cz_launch_count())
    ev0.record(C.stream)
    for _ in range(steps):
        step()
    ev1.record(C.stream)
    C.barrier()
    launches = int(L.cz_launch_count()) - launches0
    L.cz_profile_enable(0)
    km, kc = ctypes.c_double(0), ctypes.c_double(0)
    nprof = L.cz_profile_read_deflate(ctypes.byref(km), ctypes.byref(kc))
    ka, kb = ctypes.c_double(0), ctypes.c_double(0)
    L.cz_profile_read(ctypes.byref(ka), ctypes.byref(kb))  # (drain)
    ms_step = C.max_over_ranks(ev0.elapsed_time(ev1)) / steps
    value = C.world * U / (ms_step * 1e-3) / 1e9

    # size against zlib 1.3 level 6 on the same bytes (ONE whole zlib stream of the first 32 MiB; zlib-ng is unobtainable here)
    ratio_vs = None
    if C.rank == 0:
        sm = min(32, args.mib)
        head = d_in[:sm << 20].cpu().numpy().tobytes()
        zsize = len(zlib.compress(head, 6))
        gsize = int(lens[:sm].sum()) + 2 + 2 + 4  # + zlib header, final block, Adler-32
        ratio_vs = {"gpu_bytes_over_zlib_bytes": gsize / zsize, "sample_mib": sm, "zlib": "madler zlib %s level 6, one whole stream" % zlib.ZLIB_RUNTIME_VERSION,
                    "note": "zlib-ng (the north star's yardstick) is not in this image; zlib 1.3 L6 compresses at least as well"}

    e2e = None
    plain = None
    if not args.no_e2e:
        cap = int(L.cz_deflate_bound(U, 15, SEG_BYTES))
        h_in_p, h_in = C.host_alloc(U + 16)
        h_out_p, h_out = C.host_alloc(cap + 16)
        h_in[:U] = d_in[:U].cpu().numpy()
        plain = h_in
        out_len = ctypes.c_uint64(0)
        nsegs = ctypes.c_uint64(0)

        def e2e_step():
            rc = L.cz_deflate_segmented(ctypes.c_void_p(h_in_p), U, ctypes.c_void_p(h_out_p), cap, ctypes.byref(out_len), 6, 15, 0,
                                        SEG_BYTES, 1 << C.local_rank, None, 0, ctypes.byref(nsegs))
            _lib.check(rc, "cz_deflate_segmented")

        for _ in range(2):
            e2e_step()
        # ONE valid zlib stream: the host zlib inflates a prefix of it back to the plaintext
        d = zlib.decompressobj(15)
        head = d.decompress(h_out[:min(out_len.value, 8 << 20)].tobytes())
        assert head == h_in[:len(head)].tobytes() and len(head) > 0
        C.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            e2e_step()
        C.barrier()
        dt = C.max_over_ranks(time.perf_counter() - t0)
        floor_ms = C.link_floor_ms(int(U), int(out_len.value))
        e2e_ms = dt / steps * 1e3
        e2e = {"value": C.world * U * steps / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(U),
               "d2h_bytes_per_step": int(out_len.value), "ms_per_step": e2e_ms, "link_floor_ms": floor_ms,
               "frac_of_floor": floor_ms / e2e_ms if floor_ms else None}

    rec = None
    if C.rank == 0:
        peak, peak_src = load_peaks()
        chain_ms = kc.value / max(1, nprof) if nprof else ms_step
        achieved = (U + Cb) / (chain_ms * 1e-3) / 1e9
        cpu = None
        if not args.no_cpu and C.world == 1:
            cthreads = cpu_threads()
            sample_mib = min(args.mib, max(16, 8 * cthreads))
            src = plain if plain is not None else d_in[:sample_mib << 20].cpu().numpy()
            c = CpuDeflate(np.ascontiguousarray(src[:sample_mib << 20]), cthreads)
            secs = min(c.run() for _ in range(2))
            cpu = {"value": (sample_mib << 20) / secs / 1e9, "unit": UNIT, "cores": cthreads, "kind": "port",
                   "sample": "first %d MiB of the stream as independent 1 MiB pieces, best of 2 (%.2f s each), level 6; zlib 1.3 "
                             "stands in for zlib-ng; CPU ratio %.3f" % (sample_mib, secs, (sample_mib << 20) / float(c.lens.sum()))}
        traffic, traffic_src = traffic_entry("deflate_cfg3_dram_bytes_per_launch")
        rec = {
            "metric": "deflate_uncompressed_GBps", "value": value, "unit": UNIT, "n_gpus": C.world, "steps": steps, "ms_per_step": ms_step,
            "scaling": "weak", "dtype": "u8", "data": "synthetic",
            "config": {"workload": "cfg3: chunked deflate level 6 of one %d MiB synthetic stream (Markov text) into one valid zlib "
                                   "stream of full-flush segments" % args.mib, "segment_bytes": SEG_BYTES, "level": 6,
                       "l2": "input per step (%.2f GB) exceeds the 126 MB L2; no flush needed" % (U / 1e9)},
            "ratio": U / Cb, "ratio_vs_zlib13_l6": ratio_vs,
            "e2e": e2e, "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(U + Cb),
                         "launch": "one deflate launch = the kernel chain of cz_deflate_segments_device (checksum, chains, match search, "
                                   "parse, histogram, plan, layout, emit); achieved = (U + C) / CUDA-event duration of the chain; "
                                   "kernel_ms.match = the dominant kernel (deflate_match_sweep_kernel)",
                         "kernel_ms": {"match": km.value / max(1, nprof), "chain_total": chain_ms} if nprof else None},
            "cpu_baseline": cpu}
    if not args.no_e2e:
        C.host_free(h_in_p)
        C.host_free(h_out_p)
    return rec


# ---------------------------------------------------------------------------------------------------------------------
def bench_gzip(C):
    """cfg4: gzip encode + decode of 1 GiB mixed-entropy data (1 MiB runs of Markov text / repeated substrings / near-random),
    through the host-memory ABI (one gzip member in, one out), CRC-32 folded from per-segment values by the library."""
    torch, L, _lib, args = C.torch, C.L, C._lib, C.args
    steps = max(1, args.sub_steps)
    U = args.gzip_mib << 20
    d_plain = C.synth_device(3, U, args.seed + 7 + 104729 * C.rank)
    cap = int(L.cz_deflate_bound(U, 31, SEG_BYTES))
    h_in_p, h_in = C.host_alloc(U + 16)
    h_gz_p, h_gz = C.host_alloc(cap + 16)
    h_back_p, h_back = C.host_alloc(U + 16)
    h_in[:U] = d_plain[:U].cpu().numpy()
    del d_plain
    nseg_max = U // SEG_BYTES + 2
    idx = np.zeros(nseg_max + 1, dtype=np.uint64)
    out_len, nsegs, got = ctypes.c_uint64(0), ctypes.c_uint64(0), ctypes.c_uint64(0)
    mask = 1 << C.local_rank

    def enc():
        _lib.check(L.cz_deflate_segmented(ctypes.c_void_p(h_in_p), U, ctypes.c_void_p(h_gz_p), cap, ctypes.byref(out_len), 6, 31, 0,
                                          SEG_BYTES, mask, _p(idx), len(idx), ctypes.byref(nsegs)), "cz_deflate_segmented")

    def dec_index():
        _lib.check(L.cz_inflate_segmented(ctypes.c_void_p(h_gz_p), out_len.value, ctypes.c_void_p(h_back_p), U, ctypes.byref(got), 31,
                                          SEG_BYTES, _p(idx), nsegs.value, mask), "cz_inflate_segmented")

    one_in = np.zeros(2, dtype=np.uint64)
    one_out = np.array([0, U], dtype=np.uint64)
    one_len = np.zeros(1, dtype=np.uint64)
    one_st = np.zeros(1, dtype=np.int32)

    def dec_plain():
        one_in[1] = out_len.value
        _lib.check(L.cz_inflate_batch(1, ctypes.c_void_p(h_gz_p), _p(one_in), ctypes.c_void_p(h_back_p), _p(one_out), _p(one_len),
                                      _p(one_st), None, 31, mask), "cz_inflate_batch")

    def timed(fn):
        fn()
        C.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        C.barrier()
        return C.max_over_ranks(time.perf_counter() - t0) / steps

    launches0 = int(L.cz_launch_count())
    t_enc = timed(enc)
    Cb = int(out_len.value)
    # parity: the gzip trailer holds the CRC-32 folded from the segments; it must be the CRC-32 of the whole buffer
    crc_whole = zlib.crc32(h_in[:U]) & 0xffffffff
    trailer = h_gz[Cb - 8:Cb].tobytes()
    assert int.from_bytes(trailer[:4], "little") == crc_whole, "combined CRC-32 differs from crc32() of the whole buffer"
    assert int.from_bytes(trailer[4:], "little") == U & 0xffffffff
    d = zlib.decompressobj(31)
    head = d.decompress(h_gz[:min(Cb, 4 << 20)].tobytes())
    assert len(head) > 0 and head == h_in[:len(head)].tobytes()
    h_back[:U] = 0
    t_dec_idx = timed(dec_index)
    assert got.value == U and (h_back[:U] == h_in[:U]).all(), "segment-index decode differs"
    h_back[:U] = 0
    t_dec = timed(dec_plain)
    assert one_st[0] == 2 and one_len[0] == U and (h_back[:U] == h_in[:U]).all(), "decode of the gzip member differs"
    launches = int(L.cz_launch_count()) - launches0
    rec = None
    if C.rank == 0:
        W = C.world
        rec = {"metric": "gzip_roundtrip_uncompressed_GBps", "value": W * U / (t_enc + t_dec) / 1e9, "unit": UNIT, "n_gpus": W, "steps": steps,
               "config": {"workload": "cfg4: gzip encode + decode of %d MiB mixed-entropy data (1 MiB runs of the three classes), one "
                                      "gzip member, 1 MiB full-flush segments, level 6; host buffers (pinned), H2D + D2H timed" % args.gzip_mib},
               "encode_GBps": W * U / t_enc / 1e9, "decode_GBps": W * U / t_dec / 1e9, "decode_with_segment_index_GBps": W * U / t_dec_idx / 1e9,
               "encode_ms": t_enc * 1e3, "decode_ms": t_dec * 1e3, "decode_with_segment_index_ms": t_dec_idx * 1e3,
               "ratio": U / Cb, "crc32_combined_equals_whole": True, "gpu_launches": launches,
               "decode_path": "cz_inflate_batch on the single member: speculative split at verified full-flush points, per-piece "
                              "CRC-32 folded with crc32_combine and compared with the trailer"}
    for p in (h_in_p, h_gz_p, h_back_p):
        C.host_free(p)
    return rec


# ---------------------------------------------------------------------------------------------------------------------
def bench_partitioned(C):
    """cfg5 shape through the product's partitioner: rank 0 alone drives ALL GPUs of the job from one process
    (devices_mask = (1 << N) - 1), the other ranks wait on a host-side barrier with their GPUs idle."""
    torch, L, _lib, args = C.torch, C.L, C._lib, C.args
    torch.cuda.empty_cache()
    C.cpu_barrier()
    rec = None
    if C.rank == 0:
        N = C.world
        mask = (1 << N) - 1
        want = (args.part_mib_per_gpu << 20) * N
        budget = int(mem_available_bytes() * 0.25)  # plaintext + compressed + decoded, all pinned
        total = min(want, max(256 << 20, budget))
        rng = np.random.default_rng(args.seed)
        sizes = []
        acc = 0
        while acc < total:
            s = int(np.exp(rng.uniform(np.log(4096), np.log(16 << 20))))
            s = min(s, total - acc) if total - acc >= 4096 else total - acc
            sizes.append(s)
            acc += s
        n = len(sizes)
        in_off = np.zeros(n + 1, dtype=np.uint64)
        in_off[1:] = np.cumsum(sizes, dtype=np.uint64)
        T = int(in_off[n])
        h_plain_p, h_plain = C.host_alloc(T + 16)
        # classes round-robin per stream: class k's streams are cut from a device buffer of that class (64 KiB seeded pieces)
        t0 = time.perf_counter()
        pinned_view = torch.from_numpy(h_plain)
        for k in range(3):
            ids = list(range(k, n, 3))
            need = sum(sizes[i] for i in ids)
            if not need:
                continue
            d = C.synth_device(k, (need + STREAM_BYTES - 1) // STREAM_BYTES * STREAM_BYTES, args.seed + 31 + k)
            o = 0
            for i in ids:
                a = int(in_off[i])
                pinned_view[a:a + sizes[i]].copy_(d[o:o + sizes[i]], non_blocking=True)
                o += sizes[i]
            torch.cuda.synchronize()
            del d
        torch.cuda.empty_cache()
        t_gen = time.perf_counter() - t0
        caps = np.array([int(L.cz_deflate_bound(s, 15, SEG_BYTES)) for s in sizes], dtype=np.uint64)
        c_off = np.zeros(n + 1, dtype=np.uint64)
        c_off[1:] = np.cumsum(caps)
        h_comp_p, h_comp = C.host_alloc(int(c_off[n]) + 16)
        c_lens = np.zeros(n, dtype=np.uint64)
        c_st = np.zeros(n, dtype=np.int32)

        def enc():
            _lib.check(L.cz_deflate_batch(n, ctypes.c_void_p(h_plain_p), _p(in_off), ctypes.c_void_p(h_comp_p), _p(c_off), _p(c_lens),
                                          _p(c_st), 6, 15, 0, SEG_BYTES, mask), "cz_deflate_batch")

        def timed(fn, k=2):
            fn()
            t0 = time.perf_counter()
            for _ in range(k):
                fn()
            return (time.perf_counter() - t0) / k

        launches0 = int(L.cz_launch_count())
        t_enc = timed(enc)
        assert (c_st == 2).all()
        # pack the compressed streams (the inflate entry point takes a packed batch), gate a sample against the oracle
        p_off = np.zeros(n + 1, dtype=np.uint64)
        p_off[1:] = np.cumsum(c_lens)
        Cb = int(p_off[n])
        h_pack_p, h_pack = C.host_alloc(Cb + 16)
        for i in range(n):
            h_pack[int(p_off[i]):int(p_off[i + 1])] = h_comp[int(c_off[i]):int(c_off[i]) + int(c_lens[i])]
        C.host_free(h_comp_p)
        import oracle
        O = oracle.lib()
        for i in sorted(set(rng.integers(0, n, 24).tolist() + [int(np.argmax(sizes))])):
            sbytes = h_pack[int(p_off[i]):int(p_off[i + 1])]
            out = np.empty(sizes[i] + 16, dtype=np.uint8)
            o_off = np.array([0, sizes[i]], dtype=np.uint64)
            i_off = np.array([0, len(sbytes)], dtype=np.uint64)
            ln, st = np.zeros(1, dtype=np.uint64), np.zeros(1, dtype=np.int32)
            sb = np.ascontiguousarray(np.concatenate([sbytes, np.zeros(16, dtype=np.uint8)]))
            assert O.oz_inflate_batch(1, _p(sb), _p(i_off), _p(out), _p(o_off), _p(ln), _p(st), 15, 1) == 0, "oracle rejects stream %d" % i
            assert ln[0] == sizes[i] and (out[:sizes[i]] == h_plain[int(in_off[i]):int(in_off[i + 1])]).all(), "oracle decode of stream %d differs" % i
        h_back_p, h_back = C.host_alloc(T + 16)
        b_lens = np.zeros(n, dtype=np.uint64)
        b_st = np.zeros(n, dtype=np.int32)

        def dec():
            _lib.check(L.cz_inflate_batch(n, ctypes.c_void_p(h_pack_p), _p(p_off), ctypes.c_void_p(h_back_p), _p(in_off), _p(b_lens),
                                          _p(b_st), None, 15, mask), "cz_inflate_batch")

        t_dec = timed(dec)
        assert (b_st == 2).all() and (b_lens == np.asarray(sizes, dtype=np.uint64)).all()
        assert (h_back[:T] == h_plain[:T]).all(), "partitioned inflate output differs from the plaintext"
        launches = int(L.cz_launch_count()) - launches0
        cuts_in = np.zeros(N + 1, dtype=np.uint64)
        cuts_out = np.zeros(N + 1, dtype=np.uint64)
        L.cz_partition_by_bytes(n, _p(p_off), N, _p(cuts_in))
        L.cz_partition_by_bytes(n, _p(in_off), N, _p(cuts_out))
        per_dev_inflate = [int(p_off[int(cuts_in[k + 1])] - p_off[int(cuts_in[k])]) for k in range(N)]
        per_dev_deflate = [int(in_off[int(cuts_out[k + 1])] - in_off[int(cuts_out[k])]) for k in range(N)]
        rec = {"metric": "partitioned_uncompressed_GBps", "n_gpus": N, "devices_mask": mask, "unit": UNIT,
               "config": {"workload": "cfg5 shape: %d streams, sizes log-uniform in [4 KiB, 16 MiB], classes round-robin (Markov text / "
                                      "repeated substrings / near-random), %.2f GiB in total (%d MiB per GPU asked, host memory allows "
                                      "%.1f GiB), ONE process drives all %d GPUs through devices_mask; zlib streams made by "
                                      "cz_deflate_batch (1 MiB full-flush segments), host buffers pinned, H2D + D2H timed"
                                      % (n, T / 2**30, args.part_mib_per_gpu, budget / 2**30, N)},
               "deflate_GBps": T / t_enc / 1e9, "inflate_GBps": T / t_dec / 1e9, "deflate_ms": t_enc * 1e3, "inflate_ms": t_dec * 1e3,
               "ratio": T / Cb, "per_device_bytes": {"inflate_compressed": per_dev_inflate, "deflate_uncompressed": per_dev_deflate},
               "imbalance": {"inflate": max(per_dev_inflate) / (sum(per_dev_inflate) / N), "deflate": max(per_dev_deflate) / (sum(per_dev_deflate) / N)},
               "oracle_gate": "25 sampled streams (the largest included) inflate with the oracle to the plaintext; all %d round-trip on the GPUs" % n,
               "gpu_launches": launches, "setup_s": round(t_gen, 1)}
        for p in (h_plain_p, h_pack_p, h_back_p):
            C.host_free(p)
    C.cpu_barrier()
    return rec


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args, int(os.environ.get("RANK", "0")))
        return
    C = Ctx(args)
    torch = C.torch
    line = None
    if args.workload in ("all", "inflate"):
        line = bench_inflate(C)
        torch.cuda.empty_cache()
    subs = {}
    if args.workload in ("all", "deflate"):
        subs["deflate"] = bench_deflate(C)
        torch.cuda.empty_cache()
    if args.workload in ("all", "gzip"):
        subs["gzip_cfg4"] = bench_gzip(C)
        torch.cuda.empty_cache()
    if args.workload in ("all", "partitioned"):
        subs["partitioned"] = bench_partitioned(C)
    if C.rank == 0:
        if line is None:  # a sub-record run alone (development): promote it
            key = next(iter(subs))
            line = subs.pop(key) or {}
        line.update({k: v for k, v in subs.items()})
        print(json.dumps(line))
    if C.world > 1:
        C.dist.destroy_process_group()


if __name__ == "__main__":
    main()
