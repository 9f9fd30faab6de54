#!/usr/bin/env python
"""bench.py — KrakenUniq classification hot path on B200: Mreads/s on 150 bp reads (BASELINE.json metric).

  python bench.py --gpus 1 --steps 10 --warmup 3            # our arm (libkuq.so)
  python bench.py --impl reference --gpus 1 --steps 3 --warmup 1   # the unmodified reference on the host cores
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # one rank per GPU (replicas)

Workload (config.workload): BASELINE.json configs[1] — an "8 GB" synthetic KrakenDB (k=31, m=15: 666 M records
→ 8.0 GB database.kdb + 8.6 GB database.idx, all of it resident in HBM) and 10 M synthetic 150 bp reads,
generated on the GPU from fixed seeds (krakenuniq_b200/synth_gpu.py).  One step = one pass of the hot path
over one batch of 1 M reads (150 MB of read text, > L2; the lookups touch the 16.6 GB database at random).

`value`  : device-resident inputs, CUDA events on the launching stream (the slot's stream), max over ranks.
`e2e`    : the same metric through the C ABI with HOST buffers: pinned reads → H2D → kernel → D2H of calls and
           run-length hit lists every step, pipelined over the context's batch slots.
`--mode shards` (BASELINE configs[3]): the database is cut into one minimizer range per GPU
(each rank builds only its range), every GPU scans every batch, hits go to the owner GPU over NVLink
(`--merge p2p`, fused lookup + peer scatter) or through an NCCL all-reduce (`--merge nccl`).
Multi-GPU default: the database fits one card, so ranks are replicas; reads are partitioned across
ranks (weak scaling: 1 M reads per rank per step); the only collective is the once-per-run merge of the per-taxon
state (allreduce MAX over HLL registers, SUM over counters, all-gather of the sparse-tier keys); it is not part of a
step and is reported separately as config.end_of_run_merge_ms.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import re
import shutil
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "Mreads/s (150 bp)"
READ_LEN = 150
K, NT = 31, 15


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--db-records", type=int, default=666_000_000, help="records of the synthetic database")
    ap.add_argument("--db-passes", type=int, default=1, help="build the synthetic DB in this many minimizer ranges")
    ap.add_argument("--genomes", type=int, default=2000)
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads in the pool")
    ap.add_argument("--batch-reads", type=int, default=1_000_000, help="reads per step")
    ap.add_argument("--cpu-sample-reads", type=int, default=250_000, help="reads per CPU-reference step")
    ap.add_argument("--cpu-threads", type=int, default=0, help="threads for the reference (0 = all host cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--hll-mode", type=int, default=0, help="0 preload rule (reference default), 1 chunked, 2 dense only")
    ap.add_argument("--cache-dir", default=os.environ.get("KUQ_BENCH_CACHE", "/dev/shm"))
    ap.add_argument("--mode", default="replicas", choices=["replicas", "shards"],
                    help="multi-GPU layout: replicas (DB on every GPU, reads partitioned) or minimizer-range shards")
    ap.add_argument("--merge", default="p2p", choices=["p2p", "nccl"],
                    help="shards: hits stored into the owner's buffer over NVLink (p2p) or ids all-reduced (nccl)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi DURING the timed regions
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.idx = [], None, gpu_index
        self.marks = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark(self, t0, t1):
        self.marks.append((t0, t1))

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, line in self.rows:
            if not any(a <= ts <= b for a, b in self.marks):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------------------
def cache_dir(args):
    return os.path.join(args.cache_dir, f"kuq_bench_r{args.db_records}_g{args.genomes}_k{K}m{NT}")


def write_fastq(path, bases: np.ndarray, n_reads: int):
    """n_reads x 150 ASCII bases → FASTQ with fixed-width records (vectorised)."""
    width = 10
    hdr = np.char.zfill(np.arange(n_reads).astype("U"), width)
    rec_len = 1 + 1 + width + 1 + READ_LEN + 1 + 2 + READ_LEN + 1
    out = np.empty((n_reads, rec_len), np.uint8)
    out[:, 0] = ord("@"); out[:, 1] = ord("r")
    out[:, 2:2 + width] = np.frombuffer("".join(hdr.tolist()).encode(), np.uint8).reshape(n_reads, width)
    p = 2 + width
    out[:, p] = ord("\n"); p += 1
    out[:, p:p + READ_LEN] = bases[:n_reads * READ_LEN].reshape(n_reads, READ_LEN); p += READ_LEN
    out[:, p] = ord("\n"); out[:, p + 1] = ord("+"); out[:, p + 2] = ord("\n"); p += 3
    out[:, p:p + READ_LEN] = ord("I"); p += READ_LEN
    out[:, p] = ord("\n")
    out.tofile(path)


def ensure_files(args, db, sample_bases_host):
    """database.kdb / database.idx / taxDB / sample FASTQ for the reference binary, cached in tmpfs."""
    d = cache_dir(args)
    done = os.path.join(d, "COMPLETE")
    fq = os.path.join(d, f"sample_{args.cpu_sample_reads}.fq")
    if not os.path.exists(done):
        os.makedirs(d, exist_ok=True)
        db.write_files(os.path.join(d, "database.kdb"), os.path.join(d, "database.idx"))
        db.write_taxdb(os.path.join(d, "taxDB"))
        open(done, "w").write("ok\n")
    if not os.path.exists(fq):
        write_fastq(fq, sample_bases_host, args.cpu_sample_reads)
    return d, fq


def run_reference_classify(d, fq_files, threads):
    """Run the UNMODIFIED reference (oracle/_ref/classify -M) and parse its own stats line
    (classify.cpp:361-375): returns (n_sequences, seconds)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "classify")
    cmd = [exe, "-d", os.path.join(d, "database.kdb"), "-i", os.path.join(d, "database.idx"), "-a",
           os.path.join(d, "taxDB"), "-M", "-t", str(threads), "-o", "/dev/null"] + list(fq_files)
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    m = re.search(r"(\d+) sequences \(([\d.]+) Mbp\) processed in ([\d.]+)s", r.stderr)
    if r.returncode != 0 or not m:
        raise RuntimeError(f"reference classify failed ({r.returncode}): {r.stderr[-1500:]}")
    return int(m.group(1)), float(m.group(3))


def host_threads(args=None):
    if args is not None and getattr(args, "cpu_threads", 0) > 0:
        return args.cpu_threads
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference" and rank != 0:
        return 0                                              # rank 0 alone runs the CPU reference
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (workload generation and the product path are CUDA only)")
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    dist = None
    if world > 1 and args.impl == "ours":
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device(dev))

    from krakenuniq_b200 import synth_gpu
    t_gen = time.time()
    sharded = args.mode == "shards" and world > 1
    db = synth_gpu.GpuDatabase(args.db_records, n_genomes=args.genomes, k=K, nt=NT, seed=2, device=dev,
                               passes=args.db_passes, shard=(rank, world) if sharded else None)
    # every step classifies reads no earlier step has seen (a re-classified read finds its records already flagged /
    # its sparse-tier keys already stored and would be cheaper): the pool covers all steps of the run
    steps_total = (args.warmup + args.steps) + 4 + (max(args.warmup, 3) + args.steps)
    n_pool = max(args.reads, args.batch_reads * steps_total) if args.impl == "ours" and args.mode != "shards" else max(args.reads, args.batch_reads)
    # replicas: every rank draws its own reads (seed + rank), reads are partitioned across GPUs;
    # shards: every GPU scans the same reads
    pool_bases, _ = db.sample_reads(n_pool, READ_LEN, seed=3 + (0 if sharded else 1000 * rank))
    torch.cuda.synchronize()
    gen_s = time.time() - t_gen
    total_records = db.key_ct
    if sharded:
        t = torch.tensor([db.key_ct], device=dev, dtype=torch.int64)
        dist.all_reduce(t)
        total_records = int(t.item())
    workload = (f"{'configs[1]' if args.db_records == 666_000_000 else 'scaled configs[1]'}: {total_records * 12 / 1e9:.1f} GB synthetic KrakenDB (k={K}, m={NT}, {total_records} records) "
                f"+ {8 * ((1 << (2 * NT)) + 1) / 1e9:.1f} GB index in HBM, {n_pool} x {READ_LEN} bp reads, "
                f"{args.batch_reads} reads per step")
    n_batches = n_pool // args.batch_reads
    B = args.batch_reads

    # ---------------- reference arm: the unmodified reference on the host cores ------------------------------
    if args.impl == "reference":
        threads = host_threads(args)
        sample = pool_bases[:args.cpu_sample_reads * READ_LEN].cpu().numpy()
        d, fq = ensure_files(args, db, sample)
        del db, pool_bases
        torch.cuda.empty_cache()
        if args.warmup > 0:
            run_reference_classify(d, [fq] * min(args.warmup, 1), threads)      # warms the page cache (run-fq.sh:19)
        n_seq, secs = run_reference_classify(d, [fq] * args.steps, threads)
        v = n_seq / secs / 1e6
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "Mreads/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": {"workload": workload, "sample": f"{args.cpu_sample_reads} reads per step (FASTQ in tmpfs)"},
                "cpu_baseline": {"value": v, "unit": "Mreads/s", "cores": threads, "kind": "reference",
                                 "sample": f"oracle/_ref/classify -M -t {threads} -o /dev/null, {args.steps} x "
                                           f"{args.cpu_sample_reads} reads, time from its own stats line "
                                           "(DB load excluded, FASTQ parsing + Kraken output formatting included)"},
                "e2e": {"value": v, "unit": "Mreads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ---------------- our arm ----------------------------------------------------------------------------------
    from krakenuniq_b200 import binding
    cpu_baseline = None
    if rank == 0 and not args.no_cpu_baseline and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "classify")):
        try:
            threads = host_threads(args)
            sample = pool_bases[:args.cpu_sample_reads * READ_LEN].cpu().numpy()
            d, fq = ensure_files(args, db, sample)                 # must precede attach (values still raw taxids)
            run_reference_classify(d, [fq], threads)               # warm-up (page cache, run-fq.sh:19)
            n_files = 4                                            # same protocol as the reference arm: several files per run
            n_seq, secs = run_reference_classify(d, [fq] * n_files, threads)
            cpu_baseline = {"value": n_seq / secs / 1e6, "unit": "Mreads/s", "cores": threads, "kind": "reference",
                            "sample": f"oracle/_ref/classify -M -t {threads} -o /dev/null on {n_files} x {args.cpu_sample_reads} "
                                      "reads of the same workload (FASTQ in tmpfs), after one warm-up run; time from its own "
                                      "stats line (DB load excluded, FASTQ parsing + Kraken formatting included) — the "
                                      "reference arm's protocol"}
        except Exception as e:  # noqa: BLE001
            cpu_baseline = {"value": None, "unit": "Mreads/s", "cores": host_threads(), "kind": "reference",
                            "sample": f"failed: {e}"[:300]}

    if args.mode == "shards":
        return run_shards(args, db, pool_bases, rank, world, local_rank, dev, dist, workload, gen_s)
    clf = binding.Classifier(device=local_rank, n_slots=3, max_reads=B, max_bases=B * READ_LEN + 4096,
                             hll_mode=args.hll_mode, sparse_set_slots=1 << 30)
    clf.attach_db_device(db.records.data_ptr(), db.key_ct, db.offsets.data_ptr(), K, NT, 2)
    clf.set_taxonomy(*db.parent_map())

    d_offsets = (torch.arange(B + 2, dtype=torch.int64, device=dev) * READ_LEN)   # +1: slices are copied in 16-byte units
    # work units as process_file cuts them (classify.cpp:514-520): reads join a unit until it holds >= 500000 nt;
    # every batch starts a fresh unit (a batch = one input file of the reference)
    per_unit = -(-500000 // READ_LEN)
    units_per_batch = -(-B // per_unit)
    unit_local = (torch.arange(B, dtype=torch.int64, device=dev) // per_unit).to(torch.int32)
    host_unit_local = (np.arange(B, dtype=np.int64) // per_unit).astype(np.uint32)
    unit_bufs = {}

    def d_units(step_idx):
        if args.hll_mode != 0:
            return None
        t = unit_local + step_idx * units_per_batch
        unit_bufs[step_idx % 4] = t              # keep alive until the kernels have run
        return t.data_ptr()

    def h_units(step_idx):
        return (host_unit_local + np.uint32(step_idx * units_per_batch)) if args.hll_mode == 0 else None
    host_offsets = (np.arange(B + 1, dtype=np.uint64) * READ_LEN)

    def batch_ptr(i):
        return pool_bases.data_ptr() + (i % n_batches) * B * READ_LEN        # 16-byte aligned: B*150 % 16 == 0

    assert (B * READ_LEN) % 16 == 0 or n_batches == 1
    stream = torch.cuda.ExternalStream(clf.slot_stream(0), device=dev)
    sampler = ClockSampler(local_rank)
    sampler.start()

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def merge_state():
        """end-of-run merge across replicas: NCCL allreduce MAX/SUM + sparse-tier union"""
        if dist:
            from krakenuniq_b200 import dist as kdist
            kdist.merge_classifier_state(clf, dev)

    # ---- value: device-resident inputs ----------------------------------------------------------------------------
    step = 0
    for _ in range(args.warmup):
        clf.classify_device(0, batch_ptr(step), d_offsets.data_ptr(), B, B * READ_LEN, d_units(step)); step += 1
    clf.sync(0)
    clf.finish()                                  # the warm-up's flagged records are harvested outside the timed region
    barrier()
    launches0 = clf.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    t0 = time.time()
    with torch.cuda.stream(stream):
        ev0.record(stream)
    for _ in range(args.steps):
        clf.classify_device(0, batch_ptr(step), d_offsets.data_ptr(), B, B * READ_LEN, d_units(step)); step += 1
    # end of the run: the records flagged by the K steps become sparse-tier keys (kuq_finish) — part of the job, so
    # inside the timed region
    clf.finish()
    with torch.cuda.stream(stream):
        ev1.record(stream)
    clf.sync(0)
    barrier()
    harvest_ms_value = clf.sparse_tier_info()["last_harvest_ms"]
    sampler.mark(t0, time.time())
    launches = clf.launch_count() - launches0
    dev_ms = ev0.elapsed_time(ev1)
    if dist:
        t = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
    value = world * B * args.steps / (dev_ms / 1e3) / 1e6

    # per-launch duration of the dominant kernel + algorithmic bytes (one extra, untimed, instrumented step)
    clf.classify_device(0, batch_ptr(step), d_offsets.data_ptr(), B, B * READ_LEN, d_units(step), flags=binding.F_STATS); step += 1
    clf.sync(0)
    n_lookups, sum_probes = clf.slot_stats(0)
    # B_kmer = 16 + 12*P(n_b) + 1 per non-ambiguous k-mer, B_read = L + 4 + 4*(L-30) per read
    algo_bytes = 17 * n_lookups + 12 * sum_probes + B * (READ_LEN + 4 + 4 * (READ_LEN - 30))
    k_ms, st_ms = [], []
    for _ in range(3):
        clf.classify_device(0, batch_ptr(step), d_offsets.data_ptr(), B, B * READ_LEN, d_units(step)); step += 1
        clf.sync(0)
        k_ms.append(clf.last_kernel_ms(0))
        st_ms.append(clf.last_stage_ms(0))
    st_ms = np.mean(np.array(st_ms), axis=0)
    kern_ms = float(st_ms[1])            # k_lookup: the random-HBM stage dominates the step
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = algo_bytes / (kern_ms / 1e3) / 1e9
    traffic, traffic_src = None, None
    try:    # DRAM bytes per launch of the same kernel on the same workload, from the committed ncu capture of THIS build
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic_r02.json")))
        if args.db_records == 666_000_000 and B == 1_000_000:
            e = tj["k_lookup_dense_only" if args.hll_mode == 2 else "k_lookup_exact_hll"]
            traffic = e["dram_bytes"]
            traffic_src = {"capture": e["capture"], "kernel_ms_in_capture": e["kernel_ms"], "build": tj.get("build")}
    except Exception:
        pass
    # second roofline: random 32-byte-sector gathers, measured here with k_lookup's launch shape
    gather = None
    try:
        gs, gb, gms = binding.random_gather_peak(local_rank, 16 << 30, 1 << 29, 32)
        sectors = n_lookups + sum_probes                       # 32*(1+P) bytes per looked-up k-mer
        gather = {"random_gather_peak": gb, "unit": "GB/s of 32 B sectors", "gsectors_per_s": gs,
                  "algorithmic_sectors_per_launch": int(sectors),
                  "frac_random": 32.0 * sectors / (kern_ms / 1e3) / 1e9 / gb,
                  "how": "kuq_random_gather_peak: independent loads at hashed sector addresses of a 16 GB buffer, 256 threads x "
                         "8 CTAs/SM, best of 3; frac_random = 32*(1+P)*lookups / kernel time / this peak.  It exceeds 1 when "
                         "consecutive windows share a minimizer bin (their sectors come from L1/L2, not DRAM)"}
    except Exception as e:  # noqa: BLE001
        gather = {"random_gather_peak": None, "error": str(e)[:200]}
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "kernel": "k_lookup<MODE_FUSED>", "kernel_ms": kern_ms,
                "stage_ms": {"k_scan": float(st_ms[0]), "k_lookup": float(st_ms[1]), "k_resolve": float(st_ms[2])},
                "algorithmic_bytes_per_launch": algo_bytes, "bytes_per_read": algo_bytes / B,
                "random": gather,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy bandwidth, of measured)" if peaks else "fallback 6650 GB/s"}

    # ---- e2e: host buffers through the C ABI, H2D + D2H inside the timed region ------------------------------------
    # pinned host copies of the batches the e2e steps will classify (each batch once: no step re-sees reads)
    n_e2e = max(args.warmup, 3) + args.steps
    first = step % n_batches
    n_host = min(n_batches, n_e2e)
    host_bufs = []
    for i in range(n_host):
        p = clf.L.kuq_host_alloc(B * READ_LEN + 64)
        hb = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(B * READ_LEN + 64,))
        b = (first + i) % n_batches
        hb[:B * READ_LEN] = pool_bases[b * B * READ_LEN:(b + 1) * B * READ_LEN].cpu().numpy()
        host_bufs.append((p, hb))
    n_slots = 3
    d2h_bytes = [0]
    sanity = {"reads": 0, "classified": 0}
    e2e_step = [step]

    e2e_buf = [0]

    def run_e2e(n_steps):
        inflight = []
        for s in range(n_steps):
            slot = s % n_slots
            if len(inflight) == n_slots:
                res = clf.wait(inflight.pop(0))
                d2h_bytes[0] = 4 * B * 4 + 8 * res["n_runs"] + 64
                sanity["reads"] += B; sanity["classified"] += res["n_classified"]
            clf.submit(slot, host_bufs[e2e_buf[0] % n_host][0], host_offsets, h_units(e2e_step[0])); e2e_step[0] += 1
            e2e_buf[0] += 1
            inflight.append(slot)
        for slot in inflight:
            res = clf.wait(slot)
            d2h_bytes[0] = 4 * B * 4 + 8 * res["n_runs"] + 64
            sanity["reads"] += B; sanity["classified"] += res["n_classified"]

    run_e2e(max(args.warmup, 3))
    clf.finish()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    run_e2e(args.steps)
    clf.finish()                                  # harvest of the run's flagged records: part of the job
    torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = e0.elapsed_time(e1)
    barrier()
    sampler.mark(t0, time.time())
    if dist:
        t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * B * args.steps / (e2e_ms / 1e3) / 1e6
    # end-of-run merge of the per-taxon state across ranks: once per run, not per step — timed on its own
    merge_ms = 0.0
    if dist:
        barrier()
        tm0 = torch.cuda.Event(enable_timing=True); tm1 = torch.cuda.Event(enable_timing=True)
        tm0.record(); merge_state(); tm1.record(); torch.cuda.synchronize()
        t = torch.tensor([tm0.elapsed_time(tm1)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        merge_ms = float(t.item())
    clocks = sampler.stop()
    for p, _ in host_bufs:
        clf.L.kuq_host_free(p)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "Mreads/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": {"workload": workload, "parallelism": f"replicas x{world} (reads partitioned, DB replicated)",
                           "l2": "inputs larger than L2: 150 MB of reads per step, 16.6 GB database probed at random",
                           "hll_mode": ["preload rule", "chunked rule", "dense only"][args.hll_mode],
                           "timing": "CUDA events on the slot stream around K steps + the end-of-run harvest (kuq_finish), max over "
                                     "ranks; the once-per-run NCCL merge of the per-taxon state across ranks is timed separately",
                           "end_of_run_merge_ms": merge_ms,
                           "harvest_ms_in_timed_region": harvest_ms_value,
                           "read_pool": f"{n_pool} reads: every step of the run classifies reads no earlier step saw",
                           "sparse_tier": clf.sparse_tier_info(),
                           "workload_gen_s": gen_s},
                "roofline": roofline,
                "cpu_baseline": cpu_baseline,
                "e2e": {"value": e2e_value, "unit": "Mreads/s", "h2d_bytes_per_step": B * READ_LEN + 8 * (B + 1),
                        "d2h_bytes_per_step": int(d2h_bytes[0]), "ms_per_step": e2e_ms / args.steps,
                        "how": "kuq_submit_batch/kuq_wait_batch over 3 slots, pinned host reads, results = calls + "
                               "window counts + RLE hit lists"},
                "gpu_launches": int(launches),
                "sanity": {"classified_fraction": sanity["classified"] / max(sanity["reads"], 1),
                           "expected": "about 0.80: 80 % of the reads are sampled from the database genomes (1 % substitutions)"},
                "clocks": clocks}
        print(json.dumps(line))
    if dist:
        dist.destroy_process_group()
    return 0


def run_shards(args, db, pool_bases, rank, world, local_rank, dev, dist, workload, gen_s):
    """the database is split into `world` minimizer ranges (balanced by records, like
    prepare_chunking); every GPU scans the SAME batch and looks up the k-mers whose minimizer it owns; read r is
    resolved by one owner GPU.  Merge of the per-window ids: `p2p` = each hit is stored by the lookup kernel straight
    into the owner's buffer over NVLink (fused lookup + scatter, CUDA IPC mapped peer memory); `nccl` = every rank
    writes its own buffer and the buffers are max-all-reduced.  value = reads of the job / time."""
    import torch
    from krakenuniq_b200 import binding
    from krakenuniq_b200 import dist as kdist
    B = args.batch_reads
    n_pool = max(args.reads, B)
    n_batches = n_pool // B
    n_bins = 1 << (2 * NT)
    # every rank generated only its own minimizer range (GpuDatabase(shard=...)): ranges hold about equal records
    lo_bin, hi_bin = db.bin_lo, db.bin_hi
    rec_lo, rec_hi = 0, db.key_ct
    clf = binding.Classifier(device=local_rank, n_slots=2, max_reads=B, max_bases=B * READ_LEN + 4096,
                             hll_mode=args.hll_mode, sparse_set_slots=1 << 30)
    clf.set_db_taxid_universe(np.array(db.species, np.uint32))
    clf.attach_db_device(db.records.data_ptr(), db.key_ct, db.offsets.data_ptr(), K, NT, 2, lo_bin, hi_bin)
    clf.set_taxonomy(*db.parent_map())
    per_unit = -(-500000 // READ_LEN)
    # owner shares: whole work units, contiguous
    n_units = -(-B // per_unit)
    share_units = [kdist.partition(n_units, world, r) for r in range(world)]
    shares = [(min(a * per_unit, B), min(b * per_unit, B)) for a, b in share_units]
    lo, hi = shares[rank]
    d_offsets = (torch.arange(B + 2, dtype=torch.int64, device=dev) * READ_LEN)
    unit_local = (torch.arange(B, dtype=torch.int64, device=dev) // per_unit).to(torch.int32)
    total = B * READ_LEN
    nbytes = (total + 64) * 4
    my_buf = clf.device_alloc(nbytes)
    bounds = np.array([s[0] * READ_LEN for s in shares] + [total], np.uint64)
    bounds[0] = 0
    peers = None
    if args.merge == "p2p":
        handles = [None] * world
        dist.all_gather_object(handles, clf.ipc_export(my_buf))
        peers = [my_buf if r == rank else clf.ipc_open(handles[r]) for r in range(world)]
    my_t = kdist.device_view(my_buf, nbytes, torch.int32, dev)
    stream = torch.cuda.ExternalStream(clf.slot_stream(0), device=dev)
    keep = {}

    def step(i):
        bptr = pool_bases.data_ptr() + (i % n_batches) * B * READ_LEN
        clf.device_memset(0, my_buf, 0, nbytes)
        clf.sync(0)
        dist.barrier()                                       # all buffers zeroed
        if args.merge == "p2p":
            clf.lookup_device_peers(0, bptr, d_offsets.data_ptr(), B, total, peers, bounds)
            clf.sync(0)
            dist.barrier()                                   # all hits landed
        else:
            clf.lookup_device(0, bptr, d_offsets.data_ptr(), B, total, my_buf, only_hits=1)
            clf.sync(0)
            dist.all_reduce(my_t, op=dist.ReduceOp.MAX)
            torch.cuda.synchronize()
        if hi > lo:
            u = unit_local[lo:hi] + i * n_units
            keep[i % 3] = u
            clf.resolve_device(0, bptr, d_offsets.data_ptr() + lo * 8, hi - lo, total, my_buf,
                               u.data_ptr() if args.hll_mode == 0 else None)

    assert lo % 2 == 0
    sampler = ClockSampler(local_rank)
    sampler.start()
    s = 0
    for _ in range(args.warmup):
        step(s); s += 1
    clf.sync(0); torch.cuda.synchronize(); dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    launches0 = clf.launch_count()
    with torch.cuda.stream(stream):
        ev0.record(stream)
    for _ in range(args.steps):
        step(s); s += 1
    with torch.cuda.stream(stream):
        ev1.record(stream)
    clf.sync(0); torch.cuda.synchronize(); dist.barrier()
    sampler.mark(t0, time.time())
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dev_ms = float(ms.item())
    launches = clf.launch_count() - launches0
    st = clf.last_stage_ms(0)
    tm0 = torch.cuda.Event(enable_timing=True); tm1 = torch.cuda.Event(enable_timing=True)
    tm0.record(); kdist.merge_classifier_state(clf, dev); tm1.record(); torch.cuda.synchronize()
    merge_ms = tm0.elapsed_time(tm1)
    clocks = sampler.stop()
    cnt = clf.counts()
    tot_reads = int(cnt["n_reads"].sum())
    unclassified = int(cnt["n_reads"][cnt["taxid"] == 0].sum())
    if rank == 0:
        value = B * args.steps / (dev_ms / 1e3) / 1e6
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "Mreads/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload, "parallelism": f"database sharded by minimizer range over {world} GPUs "
                       f"({(rec_hi - rec_lo) * 12 / 1e9:.1f} GB of records on rank 0), every GPU scans every batch, "
                       f"id merge = {args.merge}", "hll_mode": ["preload rule", "chunked rule", "dense only"][args.hll_mode],
                       "timing": "CUDA events on the slot stream around the K steps (host barriers between the phases "
                                 "included), max over ranks", "end_of_run_merge_ms": merge_ms, "workload_gen_s": gen_s,
                       "owner_stage_ms_last_step": {"k_scan": st[0], "k_lookup(hll from merged ids)": st[1], "k_resolve": st[2]}},
            "sanity": {"reads_counted": tot_reads, "classified_fraction": 1.0 - unclassified / max(tot_reads, 1),
                       "expected": "about 0.80 (80 % of the reads come from the database genomes)"},
            "gpu_launches": int(launches), "clocks": clocks}))
    dist.barrier()
    if peers:
        for r in range(world):
            if r != rank:
                clf.ipc_close(peers[r])
    dist.barrier()
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
