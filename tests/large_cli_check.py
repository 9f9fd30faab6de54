#!/usr/bin/env python
"""Large-scale check of the drop-in `classify` on the bench workload (run on a GPU box, not part of pytest):
   1. parity: our classify vs the unmodified reference (`-M -t 1`) on the 8 GB synthetic DB, N_PARITY reads —
      Kraken output byte-identical, report rows identical;
   2. throughput of the whole command (FASTQ file in tmpfs → Kraken output + report) on N_SPEED reads, using the
      reference's own stats line format (classification time, excludes DB staging) and the wall clock.
usage: python tests/large_cli_check.py [n_parity_reads] [n_speed_reads]"""
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from krakenuniq_b200 import build, synth_gpu  # noqa: E402


def report_rows(path):
    rows = {}
    for line in open(path):
        if line.startswith("#") or line.startswith("%"):
            continue
        f = line.rstrip("\n").split("\t")
        rows[f[6]] = f
    return rows


def main():
    n_par = int(sys.argv[1]) if len(sys.argv) > 1 else 250_000
    n_speed = int(sys.argv[2]) if len(sys.argv) > 2 else 4_000_000

    class A:
        cache_dir, db_records, genomes, cpu_sample_reads = "/dev/shm", 666_000_000, 2000, n_par
    db = synth_gpu.GpuDatabase(A.db_records, n_genomes=A.genomes, k=31, nt=15, seed=2, device="cuda:0")
    pool, _ = db.sample_reads(max(n_par, n_speed), 150, seed=3)
    host = pool[:max(n_par, n_speed) * 150].cpu().numpy()
    d, fq_par = bench.ensure_files(A, db, host)
    fq_speed = os.path.join(d, f"sample_{n_speed}.fq")
    if not os.path.exists(fq_speed):
        bench.write_fastq(fq_speed, host, n_speed)
    del db, pool
    torch.cuda.empty_cache()
    exe = build.build_classify()
    dbargs = ["-d", f"{d}/database.kdb", "-i", f"{d}/database.idx", "-a", f"{d}/taxDB"]
    env = dict(os.environ, KUQ_SPARSE_SLOTS=str(1 << 30))

    # 1. parity at scale
    t0 = time.time()
    r = subprocess.run([os.path.join(ROOT, "oracle/_ref/classify")] + dbargs + ["-M", "-t", "1", "-o", f"{d}/ref.kraken", "-r",
                       f"{d}/ref.report", fq_par], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-1000:]
    t_ref = time.time() - t0
    os.remove(f"{d}/database.kdb.counts") if os.path.exists(f"{d}/database.kdb.counts.keep") else None
    t0 = time.time()
    r2 = subprocess.run([exe] + dbargs + ["-M", "-t", "1", "-o", f"{d}/our.kraken", "-r", f"{d}/our.report", fq_par],
                        capture_output=True, text=True, env=env)
    assert r2.returncode == 0, r2.stderr[-1000:]
    t_our = time.time() - t0
    same = open(f"{d}/ref.kraken", "rb").read() == open(f"{d}/our.kraken", "rb").read()
    a, b = report_rows(f"{d}/ref.report"), report_rows(f"{d}/our.report")
    diff = [k for k in a if a[k] != b.get(k)] + [k for k in b if k not in a]
    print(f"parity on {n_par} reads vs the 8 GB DB: kraken identical={same}, report rows={len(a)} differing={len(diff)} "
          f"{diff[:5]}; wall: reference {t_ref:.1f}s, ours {t_our:.1f}s")
    for k in diff[:3]:
        print("  ref", a.get(k), "\n  our", b.get(k))

    # 2. throughput of the whole command
    t0 = time.time()
    r3 = subprocess.run([exe] + dbargs + ["-M", "-t", str(min(64, os.cpu_count() or 1)), "-o", f"{d}/speed.kraken", "-r",
                         f"{d}/speed.report", fq_speed],
                        capture_output=True, text=True, env=env)
    wall = time.time() - t0
    assert r3.returncode == 0, r3.stderr[-1000:]
    m = re.search(r"(\d+) sequences \(([\d.]+) Mbp\) processed in ([\d.]+)s", r3.stderr)
    n, secs = int(m.group(1)), float(m.group(3))
    print(f"classify CLI: {n} reads in {secs:.2f}s classification time = {n / secs / 1e6:.2f} Mreads/s "
          f"(whole command incl. DB staging + report: {wall:.1f}s)")
    return 0 if (same and not diff) else 1


if __name__ == "__main__":
    sys.exit(main())
