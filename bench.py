#!/usr/bin/env python
"""bench.py -- k-mers/s inserted into the de Bruijn graph (k=53) on N B200s.

Workload (default `cfg4`, BASELINE.json configs[4]): THE WHOLE JOB -- the complete 30x build of the human-scale
graph from an EMPTY table (k=53, L=24, W=250 -> 4.19 G slots x 24 B = 100.7 GB, hash-sharded over the N GPUs;
310 M pairs of 2x150 bp reads drawn by a counter-based generator from a synthetic 3.1 Gbp genome; 53 G k-mer
insertions; 296 batches of 2^20 pairs).  The job is cut into K equal steps (`--steps`); what is timed, back to back,
are WINDOWS of consecutive batches whose reads were generated on the device and are resident in HBM before the
window's timed region starts (as many batches as the pool holds, whatever K is).  A batch is partitioned into the
graph's region stripes; the stripes accumulate batches and are inserted whenever they are full and at the end of
the job (N=1: lasv_graph_set_accumulation, streamed insert; N>1: every rank bins A batches, ONE exchange and ONE
insert on the owners per A batches, two buffer sets).  The W warm-up steps run the first window's reads into the
table, which is then cleared (untimed).  The same read set is used for every N (strong scaling), so the graph's
order-independent fingerprint must be equal at N = 1, 2, 4, 8 -- and equal to CFG4_FINGERPRINT.

  value : k-mers/s of the whole job, reads already resident in HBM (CUDA events around every window, max over ranks)
  steady_state : the same after the job, on the finished table (99 % of the operations update an existing node)
  e2e   : k-mers/s through the C-ABI host-buffer call (N=1: lasv_graph_load_reads_host_async through a staging ring;
          pinned host memory -> device copies and the statistics read-back inside the timed region, which ends with
          a flush and a device synchronisation), on the finished table
  roofline    : dominant kernel of the job, algorithmic bytes / CUDA-event kernel time over the measured HBM
                copy bandwidth (MEASURED_PEAKS.json); the other kernels' shares next to it
  parity      : every 64th pair of the job's reads through the same (sharded, accumulating) path into a second small
                graph, and at N=1 also into the job's own 2^24 x 250 table: every node compared on rank 0 with the
                compiled reference's own table (oracle/_ref)
  cpu_baseline: the compiled reference (oracle/_ref/libref_63.so) timed on one host core on a bounded sample
                of the same reads (N=1 only)

`--impl reference` times the reference's own single-threaded CPU loader instead.
`--workload cfg0|cfg1|cfg2` runs the small configs as one complete graph build per step.
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
# order-independent fingerprint of the complete default cfg-4 graph (296 batches of 2^20 pairs): the same at N = 1, 2, 4, 8,
# with the random-access insert of round 1 and with the streamed insert -- any run of the default job must reproduce it
CFG4_FINGERPRINT = "b6385601b9dfaa1e"
UNIT = "k-mers/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=["cfg4", "cfg0", "cfg1", "cfg2"])
    ap.add_argument("--pairs-per-step", type=int, default=1 << 20,
                    help="cfg0/1/2 only; cfg4 cuts the whole job into --steps steps")
    ap.add_argument("--pairs-per-batch", type=int, default=1 << 20, help="cfg4: read pairs per partition launch")
    ap.add_argument("--fraction", type=float, default=1.0, help="cfg4: part of the 30x read set to load (experiments)")
    ap.add_argument("--accumulate-gb", type=float, default=-1.0,
                    help="cfg4, N=1: read stream per insert in GB (-1: all that fits next to the table, 0: one insert per batch)")
    ap.add_argument("--accumulate-batches", type=int, default=0,
                    help="cfg4, N>1: batches a rank bins before one exchange + insert (0: from the free memory)")
    ap.add_argument("--insert", choices=["auto", "random", "streamed"], default="auto")
    ap.add_argument("--steady-steps", type=int, default=3)
    ap.add_argument("--e2e-pairs", type=int, default=4 << 20, help="cfg4: read pairs per host-buffer call of the e2e leg")
    ap.add_argument("--pool-batches", type=int, default=8, help="cfg4, N=1: batches of reads resident per timed window")
    ap.add_argument("--e2e-steps", type=int, default=0, help="cfg4, N=1: host-buffer calls of the e2e leg (0: --steps)")
    ap.add_argument("--e2e-staging-gb", type=float, default=6.0, help="cfg4, N=1: device staging ring of the host-buffer calls")
    ap.add_argument("--parity-pairs", type=int, default=20_000)
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-prefill", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="insert each batch before the next one is partitioned (one stream)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--trace", action="store_true", help="cfg4, N>1: CUDA-event timeline of rank 0's rounds (bin / exchange / insert)")
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
        "higher_is_better": True, "scaling": "strong" if args.workload in ("cfg4", "cfg0") else "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args, w, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": ref.kind,
                         "sample": f"{sample_pairs} read pairs ({2 * sample_pairs * w.read_len} bases) per step of the "
                                   f"same synthetic reads, reference table 2^{L}x{w.bucket_size}, "
                                   f"single-threaded loader (the reference has no threading), host {host_cpu_name()}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def job_geometry(args, w, world):
    """cfg4: the whole read set cut into K steps of whole batches; every rank takes part in every batch round."""
    ppb = args.pairs_per_batch
    total_pairs = int(w.n_pairs * args.fraction)
    n_batches = (total_pairs + ppb - 1) // ppb
    # the same read set at N = 1, 2, 4, 8: whole batches, a multiple of 8 of them (of the world size beyond 8)
    mult = 8 if 8 % world == 0 else world
    n_batches = (n_batches + mult - 1) // mult * mult
    per_rank = n_batches // world
    K = args.steps
    # batches of rank r in step s: per_rank / K, the first per_rank % K steps one more
    counts = [per_rank // K + (1 if s_ < per_rank % K else 0) for s_ in range(K)]
    prefix = [0]
    for c in counts:
        prefix.append(prefix[-1] + c)
    return {"pairs_per_batch": ppb, "batches": n_batches, "pairs": n_batches * ppb, "per_rank": per_rank,
            "step_counts": counts, "step_prefix": prefix, "max_per_step": max(counts) if counts else 0,
            "batches_per_step": max(counts) * world if counts else 0}


def workload_config(args, w, n_gpus):
    if args.workload == "cfg4":
        geo = job_geometry(args, w, n_gpus)
        return {"workload": "cfg4 human-scale, the whole job: synthetic 3.1 Gbp genome, 30x of 2x150 bp PE reads, k=53 "
                            "(MAXK=63), -s 5, table L=24 W=250 (100.7 GB) hash-sharded over the GPUs, built from an "
                            "empty table in `steps` equal steps",
                "kmer_size": w.kmer_size, "mem_height": w.number_bits, "mem_width": w.bucket_size,
                "read_pairs_total": geo["pairs"], "read_pairs_per_step": geo["pairs"] / max(args.steps, 1),
                "read_pairs_per_batch": geo["pairs_per_batch"], "read_len": w.read_len, "fraction_of_30x": args.fraction,
                "sharding": f"hash x{n_gpus}",
                "l2_policy": "inputs (>= 0.6 GB per batch), stripes (GBs) and table (>= 13 GB/GPU) exceed the 126 MB L2"}
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
    elif args.workload == "cfg4":
        main_job(args, w)
    else:
        main_small(args, w)


def reference_records(ref, N):
    """Sorted-able records (kmer, covg, edges) of whatever a ReferenceLoader has built: the reference's own
    HashTable memory (element.h:77-82) when the compiled reference runs, else the oracle port's dump."""
    from lasv_b200.graph import element_dtype, record_dtype
    if ref.kind != "reference":
        return ref.graph.records()
    n_slots = int(ref.view.number_buckets) * int(ref.view.bucket_size)
    ed = element_dtype(N)
    raw = (C.c_ubyte * (n_slots * ed.itemsize)).from_address(ref.view.table)
    table = np.frombuffer(raw, dtype=ed)
    live = table[table["status"] != 0]
    out = np.zeros(len(live), dtype=record_dtype(N))
    out["kmer"], out["covg"], out["edges"] = live["kmer"], live["coverage"], live["edges"]
    return out


PARITY_EVERY = 64     # BASELINE.md section 3 / SURVEY 8d: a deterministic every-64th-pair subsample goes through the reference


def parity_reads(w, pairs):
    """Pairs 0, 64, 128, ... of the job's read set (host arrays in loader order: mate 1, mate 2)."""
    from lasv_b200 import synth
    p = w.params()
    seqs, quals = [], []
    for i in range(pairs):
        sq, ql, _ = synth.reads_host(p, 2 * PARITY_EVERY * i, 2)
        seqs.append(sq)
        quals.append(ql)
    return np.concatenate(seqs), np.concatenate(quals)


_REF_CACHE = {}


def reference_graph_of(w, seq, qual, L, N):
    """(records of the compiled reference's own table for these reads, k-mers it inserted, kind, seconds)."""
    key = (L, len(seq))
    if key not in _REF_CACHE:
        t0 = time.perf_counter()
        ref = ReferenceLoader(w.kmer_size, L, w.bucket_size)
        n_ref, _ = ref.load(seq, qual, w.read_len, w.q_thresh)
        theirs = reference_records(ref, N)
        kind = ref.kind
        ref.close()
        _REF_CACHE[key] = (theirs, n_ref, kind, time.perf_counter() - t0)
    return _REF_CACHE[key]


def full_geometry_leg(args, w, g, dev):
    """N = 1: the same subsample through the job's own path into THE JOB'S OWN TABLE (2^24 x 250, 107 GB of lines:
    64-bit line offsets, 50 lines per bucket), cleared first; every node read back and compared with the compiled
    reference's table for the same reads (file_reader.c:910-1077 -> hash_table.c:270-327)."""
    import torch
    from lasv_b200 import _capi
    from oracle import oracle as O
    pairs = args.parity_pairs
    seq, qual = parity_reads(w, pairs)
    stride = w.read_len + 1
    n_reads = 2 * pairs
    stream = torch.cuda.current_stream().cuda_stream
    g.set_accumulation(0)
    g.clear(stream)
    torch.cuda.synchronize()
    g.set_insert_mode(_capi.INSERT_STREAMED)       # the job's insert kind (a batch this small would choose random access)
    before = g.insert_counts()
    chunk = ((n_reads + 2) // 3 + 1) // 2 * 2
    g.set_accumulation(2 * chunk * stride)
    pad = np.full(16, 10, np.uint8)
    d_seq = torch.from_numpy(np.concatenate([seq, pad])).to(dev)
    d_qual = torch.from_numpy(np.concatenate([qual, pad])).to(dev)
    d_offs = torch.arange(chunk + 1, dtype=torch.int64, device=dev) * stride
    for c0 in range(0, n_reads, chunk):
        nr = min(chunk, n_reads - c0)
        g.load_reads_device(d_seq[c0 * stride:], d_qual[c0 * stride:], nr * stride, d_offs[: nr + 1], nr, w.cutoff, stream)
    g.flush(stream)
    torch.cuda.synchronize()
    st = g.stats()
    counts = {k_: v - before[k_] for k_, v in g.insert_counts().items()}
    ours = g.records()
    L = max(sample_table_bits(w, pairs, 1), 9)
    theirs, n_ref, kind, secs = reference_graph_of(w, seq, qual, L, g.N)
    same = len(ours) == len(theirs) and np.array_equal(O.sort_records(ours), O.sort_records(theirs))
    return {"status": "ok" if same and st["kmers_inserted"] == n_ref and counts["streamed"] == 2 and not counts["random"] else "FAILED",
            "against": "compiled reference (oracle/_ref), its own HashTable" if kind == "reference" else "oracle port",
            "table": f"the job's own 2^{w.number_bits}x{w.bucket_size} table ({g.table_bytes / 1e9:.1f} GB of lines), cleared",
            "reads": f"every {PARITY_EVERY}th pair of the job's read set, {pairs} pairs",
            "nodes": int(len(ours)), "nodes_reference": int(len(theirs)), "kmers_inserted": int(st["kmers_inserted"]),
            "kmers_reference": int(n_ref), "inserts_by_kind": counts}


def parity_leg(args, w, rank, world, local_rank, dev):
    """A fixed small read set through the same path as the job (partition -> [exchange] -> streamed insert) into a
    second, small graph; its records, gathered on rank 0, against the compiled reference's table for the same
    reads.  Collective."""
    import torch
    import torch.distributed as dist
    from lasv_b200 import synth
    from lasv_b200.sharded import ShardedDBGraph
    from oracle import oracle as O
    pairs = args.parity_pairs
    n_reads = 2 * pairs
    L = max(sample_table_bits(w, pairs, 1), 9 + int(np.log2(world)))
    stride = w.read_len + 1
    seq, qual = parity_reads(w, pairs)
    per = (n_reads // world) // 2 * 2
    lo_r, hi_r = rank * per, (n_reads if rank == world - 1 else (rank + 1) * per)
    nb = (hi_r - lo_r) * stride
    saved = {k: os.environ.get(k) for k in ("LASV_B200_INSERT", "LASV_B200_BINS")}
    os.environ["LASV_B200_INSERT"] = "streamed"        # the job's insert kind, on a table too small to choose it itself
    os.environ["LASV_B200_BINS"] = str(max(1, (1 << L) // world // 256))
    try:
        # three batches per rank, two of them accumulated per exchange / insert, pipelined like the job
        n_mine = hi_r - lo_r
        chunk = ((n_mine + 2) // 3 + 1) // 2 * 2
        small = ShardedDBGraph(w.kmer_size, L, w.bucket_size, rank, world, local_rank,
                               max_kmers_per_batch=chunk * stride, max_bytes_per_batch=chunk * stride,
                               max_reads_per_batch=chunk, pipeline=(world > 1 and not args.no_pipeline),
                               accumulate_batches=2)
        if world == 1:
            small.graph.set_accumulation(2 * chunk * stride)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    pad = np.full(16, 10, np.uint8)
    d_seq = torch.from_numpy(np.concatenate([seq[lo_r * stride: hi_r * stride], pad])).to(dev)
    d_qual = torch.from_numpy(np.concatenate([qual[lo_r * stride: hi_r * stride], pad])).to(dev)
    d_offs = torch.arange(chunk + 1, dtype=torch.int64, device=dev) * stride
    for c0 in range(0, n_mine, chunk):
        nr = min(chunk, n_mine - c0)
        small.load_reads_device(d_seq[c0 * stride:], d_qual[c0 * stride:], nr * stride, w.cutoff, d_offs[: nr + 1], nr)
    small.flush()
    torch.cuda.synchronize()
    st = small.stats()
    counts = small.graph.insert_counts()
    mine = small.graph.records()
    raw = torch.from_numpy(mine.view(np.uint8).reshape(-1).copy()).to(dev)
    if world > 1:
        sizes = torch.zeros(world, dtype=torch.int64, device=dev)
        sizes[rank] = raw.numel()
        dist.all_reduce(sizes)
        cap = int(sizes.max().item())
        padded = torch.zeros(cap, dtype=torch.uint8, device=dev)
        padded[: raw.numel()] = raw
        parts = [torch.empty(cap, dtype=torch.uint8, device=dev) for _ in range(world)]
        dist.all_gather(parts, padded)
        blobs = [parts[r][: int(sizes[r].item())].cpu().numpy() for r in range(world)] if rank == 0 else []
    else:
        blobs = [raw.cpu().numpy()]
    small.free()
    if rank != 0:
        return None
    ours = np.concatenate([b.view(mine.dtype) for b in blobs]) if blobs else mine
    theirs, n_ref, kind, ref_secs = reference_graph_of(w, seq, qual, L, small.N)
    same = len(ours) == len(theirs) and np.array_equal(O.sort_records(ours), O.sort_records(theirs))
    out = {"status": "ok" if same and st["kmers_inserted"] == n_ref else "FAILED",
           "against": "compiled reference (oracle/_ref), its own HashTable" if kind == "reference" else "oracle port",
           "read_pairs": pairs, "reads": f"every {PARITY_EVERY}th pair of the job's read set",
           "table": f"2^{L}x{w.bucket_size} over {world} GPU(s)", "nodes": int(len(ours)),
           "nodes_reference": int(len(theirs)), "kmers_inserted": int(st["kmers_inserted"]), "kmers_reference": int(n_ref),
           "inserts_by_kind_rank0": counts, "reference_seconds": ref_secs}
    return out


def main_job(args, w):
    """cfg4: the whole job in K steps (see the module docstring)."""
    import torch
    import torch.distributed as dist
    from lasv_b200 import _capi, synth
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

    def max_over_ranks(xs):
        t = torch.tensor(list(xs), dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.cpu()]

    K, Wm = args.steps, args.warmup
    geo = job_geometry(args, w, world)
    stride = w.read_len + 1
    reads_per_batch = 2 * geo["pairs_per_batch"]
    n_bytes = reads_per_batch * stride
    local_per_step = geo["max_per_step"]
    stream = torch.cuda.current_stream().cuda_stream
    p = w.params()
    batch_dev_bytes = 2 * (n_bytes + 16)

    # The reads of a WINDOW of consecutive batches are generated on the device and are resident in HBM before the
    # window's timed region starts; a window is as many batches as the pool holds, whatever K is (the K steps are
    # the job cut into equal parts, the windows are what is timed, back to back).  N = 1: 8 batches (5 GB) --
    # everything runs on one stream, so the generator of the next window simply queues behind this window's kernels
    # and no flush is needed in between.  N > 1: the exchange and the inserts run on side streams, so a timed
    # region must end with everything joined; the pool is an eighth (N = 8: a sixth, the whole job) of the memory
    # next to the shard, and the batches inside a window follow each other without a flush.
    free0, _ = torch.cuda.mem_get_info()
    shard_bytes = (1 << w.number_bits) // world * ((w.bucket_size + 4) // 5) * 128
    if world == 1:
        pool_cap = max(1, args.pool_batches)
        resident = pool_cap * batch_dev_bytes <= max(0, free0 - shard_bytes - (36 << 30))
    else:
        budget = min(48 << 30, (free0 - shard_bytes) // (8 if world <= 4 else 6))
        pool_cap = int(max(1, budget // batch_dev_bytes))
        resident = True
    pool_cap = max(1, min(pool_cap, geo["per_rank"]))
    job_items = [(s_, i) for s_ in range(K) for i in range(geo["step_counts"][s_])]
    windows = [job_items[x: x + pool_cap] for x in range(0, len(job_items), pool_cap)]
    n_pool = max(2, pool_cap) if resident else 2
    R = pool_cap / max(1, local_per_step)          # steps per window (reported)

    # N > 1: a rank bins A batches into the same stripes before ONE exchange and ONE insert on the owners (two buffer
    # sets: the next A batches are binned while the previous ones travel and are inserted).  A from the memory next
    # to the shard and the resident reads (4 x 3 GB of stripes per batch: send + receive, two sets, + the insert's
    # scratch), at least ~6 rounds per job.
    acc_batches = 1
    if world > 1 and args.accumulate_gb:
        # (LASV_B200_EXCHANGE=stores: no send stripes, half the memory per batch)
        n_sets = 2 if os.environ.get("LASV_B200_EXCHANGE") == "stores" else 4
        per_batch = n_sets * 1.04 * 16 * reads_per_batch * (w.read_len - w.kmer_size + 1) * 1.13
        room = free0 - shard_bytes - n_pool * batch_dev_bytes - (16 << 30)
        acc_batches = int(max(1, min(room / per_batch, (geo["per_rank"] + 5) // 6, 16)))
        if args.accumulate_batches > 0:
            acc_batches = args.accumulate_batches
    graph = ShardedDBGraph(w.kmer_size, w.number_bits, w.bucket_size, rank, world, local_rank,
                           max_kmers_per_batch=reads_per_batch * (w.read_len - w.kmer_size + 1),
                           max_bytes_per_batch=n_bytes, max_reads_per_batch=reads_per_batch,
                           pipeline=(world > 1 and not args.no_pipeline), accumulate_batches=acc_batches)
    g = graph.graph
    slot = g.element_bytes
    g.set_insert_mode({"auto": _capi.INSERT_AUTO, "random": _capi.INSERT_RANDOM, "streamed": _capi.INSERT_STREAMED}[args.insert])

    pool = [(torch.empty(n_bytes + 16, dtype=torch.uint8, device=dev), torch.empty(n_bytes + 16, dtype=torch.uint8, device=dev))
            for _ in range(n_pool)]
    offs_dev = torch.arange(reads_per_batch + 1, dtype=torch.int64, device=dev) * stride
    granted = 0
    if world == 1 and args.accumulate_gb:
        granted = g.set_accumulation((1 << 64) - 1 if args.accumulate_gb < 0 else int(args.accumulate_gb * 1e9))

    def step_count(step):              # batches of this rank in a step (steps past the job: the largest step)
        return geo["step_counts"][step] if step < K else local_per_step

    def first_read(step, i):           # (step, local batch i of this rank) -> first read of that batch
        if step < K:
            mine = geo["step_prefix"][step] + i
        else:
            mine = geo["per_rank"] + (step - K) * local_per_step + i
        return (mine * world + rank) * reads_per_batch

    def run_window(todo, e0=None, e1=None, flush=True):
        """The batches `todo` = [(step, i)] through the hot path.  The stripes are inserted whenever they are full
        (N=1: lasv_graph_set_accumulation; N>1: every `acc_batches` batches); at the end of the window everything
        queued on side streams is joined (batches binned but not yet exchanged stay in the stripes), and with `flush`
        (the end of the job) what the stripes still hold is inserted too."""
        if resident:
            for slot_, (s_, i) in enumerate(todo):
                synth.reads_device(p, first_read(s_, i), reads_per_batch, pool[slot_][0], pool[slot_][1], stream)
            if world > 1:
                barrier()              # (N=1: same stream as the kernels, nothing to wait for)
        l0 = lib.lasv_kernel_launches()
        if e0 is not None:
            e0.record()
        for slot_, (s_, i) in enumerate(todo):
            sq = pool[slot_ % n_pool]
            if not resident:
                synth.reads_device(p, first_read(s_, i), reads_per_batch, sq[0], sq[1], stream)
            graph.load_reads_device(sq[0], sq[1], n_bytes, w.cutoff, offs_dev, reads_per_batch)
        if flush:
            graph.flush()
        elif world > 1:
            graph.join()               # rounds already shipped complete inside this window; pending batches stay binned
        if e1 is not None:
            e1.record()
        return lib.lasv_kernel_launches() - l0

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---------------------------------------------------------------- warm-up, then an empty table again
    for _ in range(Wm):
        run_window(windows[0])
    barrier()
    graph.stats()                      # (raises if anything overflowed)
    g.clear(stream)
    barrier()

    # ---------------------------------------------------------------- the job
    if world == 1:
        g.enable_kernel_timing(True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in windows]
    launches = 0
    if args.trace and world > 1:
        graph.trace = []
    barrier()
    wall0 = time.monotonic()
    torch.cuda.profiler.start()        # `ncu --profile-from-start off` sees the timed steps only
    # the K steps, window by window; the job ends with the insert of what the stripes still hold
    for wi, items in enumerate(windows):
        launches += run_window(items, *ev[wi], flush=(wi == len(windows) - 1))
    barrier()
    torch.cuda.profiler.stop()
    wall1 = time.monotonic()
    step_ms = max_over_ranks(e0.elapsed_time(e1) for e0, e1 in ev)
    dt = sum(step_ms) * 1e-3
    timeline = None
    if args.trace and world > 1:
        timeline = [[round(x, 2) for x in row] for row in graph.timeline()]
        graph.trace = None
    kinds = g.kernel_times_by_kind() if world == 1 else None
    if world == 1:
        g.enable_kernel_timing(False)
    st_job = graph.stats()
    summary = graph.summary()
    kmers = st_job["kmers_inserted"]
    value = kmers / dt
    clocks = sampler.summarise(sampler.window(wall0, wall1)) if rank == 0 else None
    inserts_by_kind = g.insert_counts()

    # ---------------------------------------------------------------- steady state on the finished table
    steady = None
    if args.steady_steps > 0:
        st_items = [(K + s_, i) for s_ in range(args.steady_steps) for i in range(local_per_step)]
        sw = [st_items[x: x + pool_cap] for x in range(0, len(st_items), pool_cap)]
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in sw]
        for wi, items in enumerate(sw):
            run_window(items, *evs[wi], flush=(wi == len(sw) - 1))
        barrier()
        sms = max_over_ranks(e0.elapsed_time(e1) for e0, e1 in evs)
        st2 = graph.stats()
        steady = {"value": (st2["kmers_inserted"] - kmers) / (sum(sms) * 1e-3), "unit": UNIT, "steps": args.steady_steps,
                  "ms_per_step": sum(sms) / args.steady_steps, "novel_kmers": st2["unique_kmers"] - st_job["unique_kmers"],
                  "what": "further steps of fresh reads on the finished table (80 % full)"}

    # ---------------------------------------------------------------- roofline
    peak, peak_src = measured_peaks()
    bases = geo["pairs"] * 2 * w.read_len
    step_alg_bytes = kmers * 2 * slot + 2 * bases * stride / w.read_len      # 2*sizeof(Element) + 2 B/base
    rec_bytes = 16
    if world == 1 and kinds:
        alg = {"partition": 2 * bases * stride / w.read_len + kmers * rec_bytes,   # reads in, one record out
               "sort": kmers * rec_bytes * 3,                                      # records: two reads, one write
               "stream": kmers * 2 * slot,                                         # find-or-insert: slot in, slot out
               "insert": kmers * 2 * slot}
        names = {"partition": "build_kernel<MODE_BIN>", "sort": "sg_sort_kernel", "stream": "sg_insert_kernel",
                 "insert": "insert_bins_kernel"}
        cands = ["partition", "sort", "stream"] if kinds["stream"][1] else ["partition", "insert"]
        dom = max(cands, key=lambda k_: kinds[k_][0])
        if dom == "partition" and kinds["stream"][1]:
            dom = "stream"             # the find-or-insert is the kernel the roofline is asked of; the shares are below
        ms_tot, n_l = kinds[dom]
        achieved = alg[dom] / (ms_tot * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "peak_source": peak_src,
                    "algorithmic_bytes_per_kmer": alg[dom] / kmers, "kernel_ms": ms_tot / max(n_l, 1), "launches_timed": n_l,
                    "kmers_per_launch": kmers / max(n_l, 1), "traffic": ncu_traffic(names[dom]),
                    "kernels": {k_: {"ms_total": kinds[k_][0], "launches": kinds[k_][1], "share_of_job": kinds[k_][0] * 1e-3 / dt,
                                     "algorithmic_GBps": (alg[k_] / (kinds[k_][0] * 1e-3) / 1e9) if k_ in alg and kinds[k_][0] else None}
                                for k_ in kinds},
                    "step_achieved": step_alg_bytes / dt / 1e9, "step_frac": step_alg_bytes / dt / 1e9 / peak}
    else:
        alg_bytes = kmers * (2 * slot + 2 * getattr(graph, "rec_bytes", rec_bytes)) + 2 * bases * stride / w.read_len
        achieved = alg_bytes / world / dt / 1e9
        roofline = {"bound": "hbm", "kernel": "bin+exchange+insert step (per GPU)", "achieved": achieved, "peak": peak,
                    "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                    "algorithmic_bytes_per_kmer": alg_bytes / kmers, "kernel_ms": 1e3 * dt / K, "traffic": None}

    # ---------------------------------------------------------------- e2e: host buffers through the public call
    e2e = None
    if not args.no_e2e:
        if world == 1:
            ep = max(geo["pairs_per_batch"], args.e2e_pairs // geo["pairs_per_batch"] * geo["pairs_per_batch"])
            sub = ep // geo["pairs_per_batch"]
            nb_e = sub * n_bytes
            host = []
            for b_ in range(2):
                hs, hq = torch.empty(nb_e + 16, dtype=torch.uint8).pin_memory(), torch.empty(nb_e + 16, dtype=torch.uint8).pin_memory()
                for i in range(sub):
                    synth.reads_device(p, (geo["batches"] + (args.steady_steps + 1) * geo["batches_per_step"] + b_ * sub + i) * reads_per_batch,
                                       reads_per_batch, pool[0][0], pool[0][1], stream)
                    torch.cuda.synchronize()
                    hs[i * n_bytes: (i + 1) * n_bytes].copy_(pool[0][0][:n_bytes])
                    hq[i * n_bytes: (i + 1) * n_bytes].copy_(pool[0][1][:n_bytes])
                host.append((hs, hq))
            offs_host = np.arange(sub * reads_per_batch + 1, dtype=np.uint64) * np.uint64(stride)
            # Asynchronous host-buffer calls: a call returns when its copies have left the pinned buffers; the copies
            # run ahead of the kernels through a staging ring in HBM, the chunks join the graph's accumulation and
            # the stripes are inserted whenever they are full -- so the host->device traffic of the next calls
            # overlaps the insert.  Every call queues a device->host copy of the statistics behind its kernels.
            # The timed region ends with the flush of what the stripes still hold and a device synchronisation.
            g.set_accumulation(0)
            del pool[1:]
            torch.cuda.empty_cache()
            ring = g.set_staging(int(args.e2e_staging_gb * (1 << 30)))
            granted_e = g.set_accumulation((1 << 64) - 1)
            Ke = args.e2e_steps or K
            rows = torch.zeros((Ke + Wm, 9), dtype=torch.int64).pin_memory()
            torch.cuda.synchronize()
            for s_ in range(Wm):
                g.load_reads_host_async(host[s_ % 2][0][:nb_e], host[s_ % 2][1][:nb_e], offs_host, w.cutoff, rows[s_])
            g.flush(stream)
            torch.cuda.synchronize()
            sa = g.stats()
            t0 = time.perf_counter()
            for s_ in range(Ke):
                g.load_reads_host_async(host[s_ % 2][0][:nb_e], host[s_ % 2][1][:nb_e], offs_host, w.cutoff, rows[Wm + s_])
            g.flush(stream)
            torch.cuda.synchronize()
            dte = time.perf_counter() - t0
            sb = g.stats()
            ke = sb["kmers_inserted"] - sa["kmers_inserted"]
            reads_seen = int(rows[Wm + Ke - 1][0]) - int(rows[Wm - 1][0]) if Wm else int(rows[Ke - 1][0])
            if reads_seen != Ke * sub * reads_per_batch or ke <= 0:
                raise RuntimeError(f"e2e: the per-call statistics did not advance as expected ({reads_seen} reads, {ke} k-mers)")
            e2e = {"value": ke / dte, "unit": UNIT, "h2d_bytes_per_step": int(2 * nb_e + 8 * (sub * reads_per_batch + 1)),
                   "d2h_bytes_per_step": 9 * 8, "ms_per_step": 1e3 * dte / Ke, "read_pairs_per_step": ep,
                   "steps": Ke, "table": "the finished graph (steady state)",
                   "staging_ring_bytes": ring, "accumulate_stream_bytes_granted": granted_e,
                   "inserts_by_kind_total": g.insert_counts(),
                   "api": "lasv_graph_load_reads_host_async (C ABI, pinned host buffers, one call per step, statistics "
                          "copied back per call) + lasv_graph_flush and a device synchronisation before the clock stops"}
        else:
            host_seq = [torch.empty(n_bytes + 16, dtype=torch.uint8).pin_memory() for _ in range(2)]
            host_qual = [torch.empty(n_bytes + 16, dtype=torch.uint8).pin_memory() for _ in range(2)]
            for b_ in range(2):
                synth.reads_device(p, (geo["batches"] + (args.steady_steps + 1) * geo["batches_per_step"] + b_ * world + rank) * reads_per_batch,
                                   reads_per_batch, pool[0][0], pool[0][1], stream)
                torch.cuda.synchronize()
                host_seq[b_].copy_(pool[0][0])
                host_qual[b_].copy_(pool[0][1])
            barrier()
            copy_stream = torch.cuda.Stream(device=dev)
            copied = [torch.cuda.Event(), torch.cuda.Event()]
            consumed = [None, None]
            results = torch.zeros((K + Wm + 1, 9), dtype=torch.int64).pin_memory()    # one statistics row per step
            step_no = [0]

            def e2e_step(b_):
                # pinned host -> device on a copy stream (two device buffers), binned exchange and inserts behind
                # it, the step's statistics copied back behind the inserts -- nothing blocks the host
                i = step_no[0]
                step_no[0] += 1
                k2 = i % 2
                cur = torch.cuda.current_stream()
                with torch.cuda.stream(copy_stream):
                    if consumed[k2] is not None:
                        copy_stream.wait_event(consumed[k2])
                    pool[k2][0].copy_(host_seq[b_], non_blocking=True)
                    pool[k2][1].copy_(host_qual[b_], non_blocking=True)
                    copied[k2].record(copy_stream)
                cur.wait_event(copied[k2])
                graph.load_reads_device(pool[k2][0], pool[k2][1], n_bytes, w.cutoff, offs_dev, reads_per_batch)
                consumed[k2] = torch.cuda.Event()
                consumed[k2].record(cur)
                graph.stats_async(results[i])

            for s_ in range(Wm):
                e2e_step(s_ % 2)
            barrier()
            sa = graph.stats()
            t0 = time.perf_counter()
            for s_ in range(K):
                e2e_step(s_ % 2)
            graph.flush()              # what the stripes still hold is exchanged and inserted inside the timed region
            torch.cuda.synchronize()
            dte = max_over_ranks([time.perf_counter() - t0])[0]
            barrier()
            sb = graph.stats()
            e2e = {"value": (sb["kmers_inserted"] - sa["kmers_inserted"]) / dte, "unit": UNIT,
                   "h2d_bytes_per_step": int(2 * n_bytes), "d2h_bytes_per_step": 9 * 8, "ms_per_step": 1e3 * dte / K,
                   "read_pairs_per_step": geo["pairs_per_batch"] * world, "steps": K, "table": "the finished graph (steady state)",
                   "api": "pinned host -> device copies on a copy stream + sharded bin/exchange/insert + per-step statistics read-back (async)"}
    sampler.stop()
    final = graph.stats()
    capacity = g.capacity * world
    parity_full = None
    if world == 1 and not args.no_parity:
        try:
            parity_full = full_geometry_leg(args, w, g, dev)
        except Exception as ex:          # reported as a failure, never silently dropped
            parity_full = {"status": "FAILED", "error": repr(ex)[:300]}
    graph.free()
    del pool
    torch.cuda.empty_cache()

    # ---------------------------------------------------------------- parity leg, CPU baseline
    parity = None
    if not args.no_parity:
        try:
            parity = parity_leg(args, w, rank, world, local_rank, dev)
        except Exception as ex:          # reported as a failure, never silently dropped
            parity = {"status": "FAILED", "error": repr(ex)[:300]}
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

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": 1e3 * dt / K, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": workload_config(args, w, world),
            "job": {"seconds": dt, "kmers_inserted": kmers, "read_pairs": geo["pairs"], "unique_kmers": st_job["unique_kmers"],
                    "fill": st_job["unique_kmers"] / capacity, "sum_covg": summary["sum_covg"],
                    "covg_equals_inserts": summary["sum_covg"] == kmers, "bad_reads": st_job["bad_reads"],
                    "fingerprint": f"{summary['fingerprint']:016x}", "inputs_resident_before_each_step": bool(resident),
                    "accumulate_stream_bytes_granted": granted, "inserts_by_kind": inserts_by_kind,
                    "wall_seconds_incl_read_generation": wall1 - wall0, "window_ms": step_ms, "steps_per_window": R, "batches_per_window": pool_cap,
                    "batches_accumulated_per_exchange": acc_batches if world > 1 else None,
                    "exchange": getattr(graph, "exchange", None) if world > 1 else None,
                    "round_timeline_ms_rank0": timeline},
            "fingerprint": f"{summary['fingerprint']:016x}",
            "fingerprint_expected": (CFG4_FINGERPRINT if args.fraction == 1.0 and args.pairs_per_batch == 1 << 20 else None),
            "fingerprint_ok": ((f"{summary['fingerprint']:016x}" == CFG4_FINGERPRINT)
                               if args.fraction == 1.0 and args.pairs_per_batch == 1 << 20 else None),
            "parity": (("ok" if all(x["status"] == "ok" for x in (parity, parity_full) if x) else "FAILED")
                       if (parity or parity_full) else None),
            "parity_detail": parity, "parity_full_geometry": parity_full,
            "kmers_per_step": kmers // K, "gpu_launches": int(launches),
            "roofline": roofline, "steady_state": steady, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
            "host": {"numa_node_of_rank0": numa_node, "cpus_of_rank0": len(os.sched_getaffinity(0))},
            "graph": {"unique_kmers": final["unique_kmers"], "kmers_inserted_total": final["kmers_inserted"],
                      "fill": final["unique_kmers"] / capacity},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main_small(args, w):
    """cfg0 / cfg1 / cfg2: one complete graph build (table zeroing included) per step."""
    from lasv_b200 import synth

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
