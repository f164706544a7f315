#!/usr/bin/env python
"""bench.py -- k-mers/s inserted into the de Bruijn graph (k=53) on N B200s.

Workload (default `cfg4`, BASELINE.json configs[4]): human-scale table geometry
(k=53, L=24, W=250 -> 4.19 G slots x 24 B = 100.7 GB, hash-sharded over the N
GPUs), 2x150 bp paired-end reads drawn by a counter-based generator from a
synthetic 3.1 Gbp genome.  Before the timed region the table is PRE-FILLED,
untimed, with every genomic k-mer once (error-free reads tiling the genome), so
that the timed steps run in the steady state of a 30x build: ~99 % of the
operations update an existing node and the rest insert novel error k-mers into a
~75 % full table.  One step = one batch of `--pairs-per-step` read pairs PER
GPU (weak scaling) through the whole hot path: quality/N filter -> 2-bit k-mers
-> canonical key + edge byte -> (N>1: route + all-to-all) -> find-or-insert.

  value : k-mers/s with the batch already resident in HBM (CUDA-event timed)
  e2e   : the same through the C-ABI host-buffer call (pinned host memory ->
          device copies and the statistics read-back inside the timed region)
  roofline    : dominant kernel, algorithmic bytes / CUDA-event kernel time over
                the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline: the compiled reference (oracle/_ref/libref_63.so) timed on one
                host core on a bounded sample of the same reads (N=1 only)

`--impl reference` times the reference's own single-threaded CPU loader instead.
`--workload cfg0` runs BASELINE.json configs[0] (63 Mbp, 2x100 bp, L=22 W=100)
as one complete graph build per step.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "kmers_inserted_per_s_k53"
UNIT = "k-mers/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=["cfg4", "cfg0", "cfg1", "cfg2"])
    ap.add_argument("--pairs-per-step", type=int, default=1 << 20)
    ap.add_argument("--no-prefill", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="insert each batch before the next one is partitioned (one stream)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-pairs", type=int, default=100_000)
    return ap.parse_args()


# ----------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: NVML polled every few ms from a thread
    (nvidia-smi -lms as the fallback when the NVML binding is missing)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NVML_REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
                    0x4: "sw_power_cap"}

    def __init__(self, device_index: int):
        self.idx, self.proc, self.lines, self.samples = device_index, None, [], []
        self.stop_flag = threading.Event()
        self.nvml = None

    def start(self):
        try:
            import pynvml
            try:
                pynvml.nvmlInit()
            except Exception:
                time.sleep(0.2)
                pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.idx]) if vis and vis.split(",")[self.idx].isdigit() else self.idx
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag.is_set():
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                try:
                    mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.samples.append((time.monotonic(), mhz, mask))
            except Exception:
                pass
            time.sleep(0.004)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.monotonic(), line.strip()))

    def window(self, t0, t1):
        if self.nvml:
            return [x for x in self.samples if t0 <= x[0] <= t1] or self.samples[-3:]
        got = [l for t, l in self.lines if t0 <= t <= t1] or [l for _, l in self.lines[-3:]]
        if not got:   # the sampler never produced a line (slow start): one synchronous reading instead
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.idx)], capture_output=True, text=True, timeout=10).stdout
                got = [l.strip() for l in out.splitlines() if l.strip()]
            except Exception:
                got = []
        return got

    def stop(self):
        self.stop_flag.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summarise(self, lines):
        if self.nvml:
            if not lines:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
            reasons = set()
            for _, _, mask in lines:
                for bit, name in self.NVML_REASONS.items():
                    if mask & bit:
                        reasons.add(name)
            return {"sm_mhz": statistics.median(x[1] for x in lines), "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(reasons), "samples": len(lines), "source": "nvml"}
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi"}


def bind_to_gpu_numa_node(device_index: int):
    """Run this process on the CPUs of the NUMA node the GPU hangs off, BEFORE any pinned host memory is
    allocated (first touch puts the pages on that node): at 8 ranks the host->device copies otherwise all
    read one socket's memory.  Returns the node, or None if the topology cannot be read."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(device_index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        dom, b, rest = bus.split(":")
        node = int(Path(f"/sys/bus/pci/devices/{dom[-4:]}:{b}:{rest}/numa_node").read_text())
        if node < 0:
            return None
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel_key: str):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture."""
    p = ROOT / "profiles" / "ncu_traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get(kernel_key)
        except Exception:
            return None
    return None


# --------------------------------------------------------------- reference arm
class _Silence:
    """The reference prints progress to stdout; keep our stdout one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *a):
        os.dup2(self.saved, 1)
        os.close(self.saved)
        os.close(self.null)


class ReferenceLoader:
    """The reference's own loader (load_pe_filelists_into_graph_colour,
    file_reader.c:910) from oracle/_ref/libref_63.so -- compiled, unmodified
    reference code -- on a reference HashTable; falls back to the oracle port
    (oracle/liboracle.so) only if the compiled reference is absent."""

    def __init__(self, k, L, W):
        from oracle import oracle as O
        self.O, self.k, self.L, self.W = O, k, L, W
        so = O.REF_DIR / f"libref_{O.MAXK_OF_N[O.words_per_kmer(k)]}.so"
        self.kind = "reference" if so.exists() else "port"
        self.tmp = Path(tempfile.mkdtemp(prefix="lasv_bench_ref_"))
        self.inserts_before = 0
        if self.kind == "reference":
            self.lib = C.CDLL(str(so))
            self.lib.hash_table_new.restype = C.c_void_p
            self.lib.hash_table_new.argtypes = [C.c_int, C.c_int, C.c_int, C.c_short]
            with _Silence():
                self.ht = self.lib.hash_table_new(L, W, 15, k)
            if not self.ht:
                raise MemoryError("reference hash_table_new failed")
            from lasv_b200._capi import RefHashTable
            self.view = C.cast(self.ht, C.POINTER(RefHashTable)).contents
            ull = C.POINTER(C.c_ulonglong)
            self.lib.load_pe_filelists_into_graph_colour.restype = None
            self.lib.load_pe_filelists_into_graph_colour.argtypes = [
                C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_byte, C.c_char, C.c_int, C.c_void_p, C.c_char,
                C.POINTER(C.c_uint), ull, ull, ull, ull, C.c_void_p, C.c_ulong]
        else:
            self.graph = O.OracleGraph(k, L, W)

    def _inserts(self):
        if self.kind == "reference":
            n = self.view.max_rehash_tries
            arr = C.cast(self.view.collisions, C.POINTER(C.c_longlong * n)).contents
            return int(sum(arr))
        return int(self.graph.stats.kmers_inserted)

    def load(self, seq, qual, read_len, q_thresh):
        """One sample of reads (fixed stride stream, mates interleaved).  Returns
        (k-mers inserted, seconds spent inside the reference's loader)."""
        O = self.O
        before = self._inserts()
        if self.kind == "reference":
            stride = read_len + 1
            s2, q2 = seq.reshape(-1, stride), qual.reshape(-1, stride)
            f1, f2 = self.tmp / "s_1.fq", self.tmp / "s_2.fq"
            O.write_fastq_fixed(f1, s2[0::2].reshape(-1), q2[0::2].reshape(-1), read_len)
            O.write_fastq_fixed(f2, s2[1::2].reshape(-1), q2[1::2].reshape(-1), read_len)
            (self.tmp / "l").write_text(f"{f1}\n")
            (self.tmp / "r").write_text(f"{f2}\n")
            files = C.c_uint(0)
            a, b, c, d = (C.c_ulonglong(0) for _ in range(4))
            hist = np.zeros(20001, dtype=np.uint64)
            with _Silence():
                t0 = time.perf_counter()
                self.lib.load_pe_filelists_into_graph_colour(
                    str(self.tmp / "l").encode(), str(self.tmp / "r").encode(), q_thresh, 0, 0, bytes([33]),
                    0, self.ht, b"\0", C.byref(files), C.byref(a), C.byref(b), C.byref(c), C.byref(d),
                    hist.ctypes.data, len(hist))
                dt = time.perf_counter() - t0
        else:
            from lasv_b200 import synth
            offs = synth.offsets_fixed(read_len, len(seq) // (read_len + 1))
            t0 = time.perf_counter()
            ok = self.graph.load_reads(seq, qual, offs, q_thresh + 33, paired=True)
            dt = time.perf_counter() - t0
            if not ok:
                raise RuntimeError("oracle table full")
        return self._inserts() - before, dt

    def close(self):
        shutil.rmtree(self.tmp, ignore_errors=True)


def host_cpu_name():
    try:
        for line in Path("/proc/cpuinfo").read_text().splitlines():
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def sample_table_bits(w, sample_pairs, steps):
    """Table for the CPU sample: same W, L chosen so the sample fills it about as
    much as the full job fills the full table (<= ~70 %)."""
    kmers = sample_pairs * 2 * (w.read_len - w.kmer_size + 1) * max(1, steps)
    L = 10
    while (1 << L) * w.bucket_size * 0.6 < kmers:
        L += 1
    return L


def run_reference_arm(args, w):
    from lasv_b200 import synth
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_pairs = max(1000, args.cpu_sample_pairs // 4)     # ~5 s of CPU work per step
    total = args.steps + args.warmup
    L = sample_table_bits(w, sample_pairs, total)
    ref = ReferenceLoader(w.kmer_size, L, w.bucket_size)
    p = w.params()
    kmers, secs = 0, 0.0
    t_all = []
    for s in range(total):
        seq, qual, _ = synth.reads_host(p, 2 * sample_pairs * s, 2 * sample_pairs)
        n, dt = ref.load(seq, qual, w.read_len, w.q_thresh)
        if s >= args.warmup:
            kmers += n
            secs += dt
            t_all.append(dt)
    ref.close()
    value = kmers / secs
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": workload_config(args, w, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": ref.kind,
                         "sample": f"{sample_pairs} read pairs ({2 * sample_pairs * w.read_len} bases) per step of the "
                                   f"same synthetic reads, reference table 2^{L}x{w.bucket_size}, "
                                   f"single-threaded loader (the reference has no threading), host {host_cpu_name()}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, w, n_gpus):
    if args.workload == "cfg4":
        return {"workload": "cfg4 human-scale steady state: synthetic 3.1 Gbp genome, 2x150 bp PE, k=53 "
                            "(MAXK=63), -s 5, table L=24 W=250 (100.7 GB) hash-sharded over the GPUs, "
                            "pre-filled with the genome's k-mers, then batches of random reads",
                "kmer_size": w.kmer_size, "mem_height": w.number_bits, "mem_width": w.bucket_size,
                "read_pairs_per_gpu_per_step": args.pairs_per_step, "read_len": w.read_len,
                "prefill": not args.no_prefill, "sharding": f"hash x{n_gpus}",
                "l2_policy": "inputs (0.6 GB/step) and table (>=12.6 GB/GPU) exceed the 126 MB L2"}
    return {"workload": f"{w.name}: synthetic {w.genome_len / 1e6:.0f} Mbp genome, 2x{w.read_len} bp PE at {w.coverage:.0f}x, "
                        f"k={w.kmer_size} (MAXK={32 * ((2 * w.kmer_size) // 64 + 1) - 1}), -s {w.q_thresh}, L={w.number_bits} "
                        f"W={w.bucket_size}; one complete graph build (table zeroing included) per step",
            "kmer_size": w.kmer_size, "mem_height": w.number_bits, "mem_width": w.bucket_size,
            "read_pairs_per_step": w.n_pairs, "read_len": w.read_len, "sharding": f"hash x{n_gpus}",
            "l2_policy": "inputs (GBs) and table (>= 6.7 GB) exceed the 126 MB L2"}


# ------------------------------------------------------------------- our arm
def main():
    args = parse_args()
    from lasv_b200 import synth
    w = synth.WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args, w)
        return

    import torch
    import torch.distributed as dist
    from lasv_b200 import _capi
    from lasv_b200.sharded import ShardedDBGraph

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    numa_node = bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _capi.lib()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: int) -> int:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        return int(t.item())

    K, Wm = args.steps, args.warmup
    stride = w.read_len + 1
    kpr = w.read_len - w.kmer_size + 1                       # k-mers per clean read
    cfg0 = args.workload != "cfg4"        # cfg0/1/2: one complete graph build (table zeroing included) per step
    if cfg0:
        reads_per_step = 2 * (w.n_pairs // world)            # strong split of the one build
    else:
        reads_per_step = 2 * args.pairs_per_step
    n_bytes = reads_per_step * stride
    stream = torch.cuda.current_stream().cuda_stream

    graph = ShardedDBGraph(w.kmer_size, w.number_bits, w.bucket_size, rank, world, local_rank,
                           max_kmers_per_batch=reads_per_step * kpr, max_bytes_per_batch=n_bytes,
                           max_reads_per_batch=reads_per_step, pipeline=(world > 1 and not args.no_pipeline))
    g = graph.graph
    table_bytes = g.capacity * g.element_bytes

    # ---- random-access ceiling on this GPU: 32 B read + atomic update per op over the table's own memory
    ra = None
    if rank == 0:
        n_sect = table_bytes // 32
        n_ops = 1 << 28
        tbl = lib.lasv_graph_device_table(g._h)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for it in range(2):
            e0.record()
            _capi.check(lib.lasv_random_access_probe(tbl, n_sect, n_ops, 1234 + it, stream))
            e1.record()
            torch.cuda.synchronize()
        ra_s = e0.elapsed_time(e1) * 1e-3
        ra = {"ops_per_s": n_ops / ra_s, "bytes_per_op": 64, "GBps": n_ops * 64 / ra_s / 1e9,
              "what": "uniform random 32 B sector read + 32-bit L2 atomic (32 B write-back) over the table"}
        g.clear(stream)
    barrier()

    p = w.params()
    offs_dev = torch.arange(reads_per_step + 1, dtype=torch.int64, device=dev) * stride
    n_batches = 1 if cfg0 else (K + Wm)
    seqs = [torch.empty(n_bytes + 16, dtype=torch.uint8, device=dev) for _ in range(n_batches)]
    quals = [torch.empty(n_bytes + 16, dtype=torch.uint8, device=dev) for _ in range(n_batches)]

    def gen(b, first_read, params=p, nr=reads_per_step):
        synth.reads_device(params, first_read, nr, seqs[b], quals[b], stream)

    # ---- untimed pre-fill: the genome's own k-mers, once each, through the same kernels
    prefill = {"kmers": 0, "seconds": 0.0}
    if not cfg0 and not args.no_prefill:
        pt = w.params(tiling=True)
        n_tiles = (w.genome_len + pt.tiling_step - 1) // pt.tiling_step
        per_rank = (n_tiles + world - 1) // world
        t0 = time.perf_counter()
        done = 0
        while done < per_rank:
            nr = min(reads_per_step, per_rank - done)
            first = rank * per_rank + done
            synth.reads_device(pt, first, nr, seqs[0], quals[0], stream)
            if nr < reads_per_step:                          # every rank must join every exchange
                seqs[0][nr * stride: n_bytes].fill_(10)
            graph.load_reads_device(seqs[0], quals[0], n_bytes, w.cutoff)
            done += reads_per_step
        barrier()
        prefill["seconds"] = time.perf_counter() - t0
        prefill["kmers"] = graph.stats()["kmers_inserted"]

    for b in range(n_batches):
        gen(b, (b * world + rank) * reads_per_step if not cfg0 else rank * reads_per_step)
    barrier()

    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(K)]

    def step(b, timed_idx=None):
        if cfg0:
            g.clear(stream)
        if world == 1:
            if timed_idx is not None:
                ev[timed_idx][0].record()
            g.load_reads_device(seqs[b], quals[b], n_bytes, None, 0, w.cutoff, stream)   # the build kernel
            if timed_idx is not None:
                ev[timed_idx][1].record()
            g.read_stats_device(seqs[b], quals[b], offs_dev, reads_per_step, w.cutoff, stream)
        else:
            graph.load_reads_device(seqs[b], quals[b], n_bytes, w.cutoff, offs_dev, reads_per_step)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---------------------------------------------------------------- value leg
    for s in range(Wm):
        step(0 if cfg0 else s)
    barrier()
    st0 = graph.stats()
    l0 = lib.lasv_kernel_launches()
    if world == 1:
        g.enable_kernel_timing(True)   # CUDA events around every launch of the two kernels, on their stream
    barrier()
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.monotonic()
    torch.cuda.profiler.start()        # `ncu --profile-from-start off` sees the timed steps only
    t_begin.record()
    if os.environ.get("LASV_BENCH_TRACE") and world > 1 and graph.pipeline:
        graph.trace = []
    for s in range(K):
        step(0 if cfg0 else Wm + s, s)
    graph.flush()                      # side-stream inserts of the last batches join the timed stream
    t_end.record()
    barrier()
    torch.cuda.profiler.stop()
    wall1 = time.monotonic()
    launches = lib.lasv_kernel_launches() - l0
    ktimes = g.kernel_times() if world == 1 else None
    if world == 1:
        g.enable_kernel_timing(False)
    dt = max_over_ranks(t_begin.elapsed_time(t_end) * 1e-3)
    st1 = graph.stats()
    kmers = st1["kmers_inserted"] - st0["kmers_inserted"] if not cfg0 else st1["kmers_inserted"] * K
    value = kmers / dt
    clocks = sampler.summarise(sampler.window(wall0, wall1)) if rank == 0 else None

    # ---- roofline of the dominant kernel (N=1: the fused build kernel)
    peak, peak_src = measured_peaks()
    slot = g.element_bytes
    kmers_per_launch = kmers / (K * world)
    step_alg_bytes = kmers_per_launch * 2 * slot + 2 * n_bytes       # 2*sizeof(Element) + 2 B/base
    extra = {}
    if world == 1 and ktimes and ktimes["batches"]:
        # dominant kernel: insert_bins_kernel (the find-or-insert); its algorithmic bytes are the
        # slot read + write-back, the 2 B/base of input belong to the partition kernel
        kt = ktimes["insert_ms"] / ktimes["batches"] * 1e-3
        alg_bytes = kmers_per_launch * 2 * slot
        kname = "insert_bins_kernel"
        extra = {"partition_kernel": "build_kernel<MODE_BIN>",
                 "partition_kernel_ms": ktimes["partition_ms"] / ktimes["batches"],
                 "launches_timed": ktimes["batches"],
                 "step_achieved": step_alg_bytes / (dt / K) / 1e9, "step_frac": step_alg_bytes / (dt / K) / 1e9 / peak}
    elif world == 1:
        kt = statistics.mean(e[0].elapsed_time(e[1]) for e in ev) * 1e-3
        alg_bytes = step_alg_bytes
        kname = "build_kernel<FUSED>"
    else:
        kt = dt / K                                          # step time bounds every kernel in it
        alg_bytes = kmers_per_launch * (2 * slot + 2 * graph.rec_bytes) + 2 * n_bytes
        kname = "bin+exchange+insert step"
    achieved = alg_bytes / kt / 1e9
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_src,
                "algorithmic_bytes_per_kmer": alg_bytes / kmers_per_launch,
                "kernel_ms": kt * 1e3, "traffic": ncu_traffic(kname)}
    roofline.update(extra)
    if ra:
        ra["kernel_kmers_per_s"] = kmers_per_launch / kt
        ra["frac_of_random_access_ceiling"] = (kmers_per_launch / kt) / ra["ops_per_s"]

    # ------------------------------------------------------------------ e2e leg
    e2e = None
    if not args.no_e2e:
        n_e2e = 1 if cfg0 else (K + Wm)
        host_seq = [torch.empty(n_bytes + 16, dtype=torch.uint8).pin_memory() for _ in range(n_e2e)]
        host_qual = [torch.empty(n_bytes + 16, dtype=torch.uint8).pin_memory() for _ in range(n_e2e)]
        base_batch = n_batches
        for b in range(n_e2e):
            first = ((base_batch + b) * world + rank) * reads_per_step if not cfg0 else rank * reads_per_step
            gen(0, first)
            torch.cuda.synchronize()
            host_seq[b].copy_(seqs[0])
            host_qual[b].copy_(quals[0])
        offs_host = (np.arange(reads_per_step + 1, dtype=np.uint64) * np.uint64(stride))
        barrier()

        copy_stream = torch.cuda.Stream(device=dev)
        copied = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [None, None]
        results = torch.zeros((K + Wm + 1, 9), dtype=torch.int64).pin_memory()    # one statistics row per step
        step_no = [0]

        def e2e_step(b):
            if cfg0:
                g.clear(None)
            if world == 1:
                return g.load_reads_host(host_seq[b], host_qual[b], offs_host, w.cutoff)   # C-ABI, host buffers
            # sharded: pinned host -> device on a copy stream (two device buffers), binned exchange and
            # inserts behind it, the step's statistics copied back behind the inserts -- nothing blocks the host
            i = step_no[0]
            step_no[0] += 1
            k2 = i % min(2, len(seqs))
            cur = torch.cuda.current_stream()
            with torch.cuda.stream(copy_stream):
                if consumed[k2] is not None:
                    copy_stream.wait_event(consumed[k2])
                seqs[k2].copy_(host_seq[b], non_blocking=True)
                quals[k2].copy_(host_qual[b], non_blocking=True)
                copied[k2].record(copy_stream)
            cur.wait_event(copied[k2])
            graph.load_reads_device(seqs[k2], quals[k2], n_bytes, w.cutoff, offs_dev, reads_per_step)
            consumed[k2] = torch.cuda.Event()
            consumed[k2].record(cur)
            graph.stats_async(results[i])                                                  # D2H of the counters
            return None

        for s in range(Wm):
            e2e_step(0 if cfg0 else s)
        barrier()
        sa = graph.stats()
        t0 = time.perf_counter()
        for s in range(K):
            e2e_step(0 if cfg0 else Wm + s)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        if world > 1 and int(results[step_no[0] - 1, 6]) <= int(results[Wm - 1, 6]):
            raise RuntimeError("e2e: the per-step statistics did not advance")
        dte = max_over_ranks(t1 - t0)
        barrier()
        sb = graph.stats()
        ke = sb["kmers_inserted"] - sa["kmers_inserted"] if not cfg0 else sb["kmers_inserted"] * K
        e2e = {"value": ke / dte, "unit": UNIT,
               "h2d_bytes_per_step": int(2 * n_bytes + 8 * (reads_per_step + 1)),
               "d2h_bytes_per_step": (7 * 8 + 20 * 8 + 8) if world == 1 else 9 * 8, "ms_per_step": 1e3 * dte / K,
               "api": "lasv_graph_load_reads_host (C ABI, pinned host buffers)" if world == 1 else
                      "pinned host -> device copies on a copy stream + sharded bin/all_to_all/insert + per-step statistics read-back (async)"}

    sampler.stop()

    # ------------------------------------------------------------ cpu baseline
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            sp = args.cpu_sample_pairs
            L = sample_table_bits(w, sp, 1)
            ref = ReferenceLoader(w.kmer_size, L, w.bucket_size)
            seq, qual, _ = synth.reads_host(p, 0, 2 * sp)
            n, secs = ref.load(seq, qual, w.read_len, w.q_thresh)
            ref.close()
            cpu = {"value": n / secs, "unit": UNIT, "cores": 1, "kind": ref.kind,
                   "sample": f"first {sp} read pairs of the same synthetic reads ({n} k-mers, {secs:.1f} s) into a "
                             f"2^{L}x{w.bucket_size} reference table; the reference loader is single-threaded; "
                             f"host {host_cpu_name()}, {os.cpu_count()} logical cores present"}
        except Exception as ex:  # the baseline is reported, never required
            cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": "unavailable", "sample": repr(ex)[:200]}

    if getattr(graph, "trace", None):
        tl = graph.timeline()
        print(f"[rank {rank}] stage timeline (ms): partition, exchange, insert", file=sys.stderr)
        for row in tl:
            print(f"[rank {rank}]   " + "  ".join(f"{a:7.2f}-{b:7.2f}" for a, b in zip(row[0::2], row[1::2])), file=sys.stderr)
        graph.trace = None
    summary = graph.summary()
    final = graph.stats()
    if rank == 0:
        line = {
            "metric": METRIC if w.kmer_size == 53 else f"kmers_inserted_per_s_k{w.kmer_size}", "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": 1e3 * dt / K, "higher_is_better": True, "scaling": "strong" if cfg0 else "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": workload_config(args, w, world),
            "kmers_per_step": kmers // K, "gpu_launches": int(launches),
            "roofline": roofline, "random_access": ra, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
            "prefill": prefill, "host": {"numa_node_of_rank0": numa_node, "cpus_of_rank0": len(os.sched_getaffinity(0))},
            "graph": {"unique_kmers": final["unique_kmers"], "fill": final["unique_kmers"] / (g.capacity * world),
                      "sum_covg": summary["sum_covg"], "kmers_inserted_total": final["kmers_inserted"],
                      "bad_reads": final["bad_reads"], "fingerprint": f"{summary['fingerprint']:016x}"},
        }
        print(json.dumps(line), flush=True)
    graph.free()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
