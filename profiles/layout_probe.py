#!/usr/bin/env python
"""Record-layout / search-shape experiment of the bin search (kuq_layout_experiment).
For every database size: generate the synthetic database in HBM, classify one batch of the bench's reads through the
product path (lookup only: KUQ_F_NO_COUNTS), then time the lookup-only variants on the windows of that batch and check
them against the product's per-window ids.
usage: python profiles/layout_probe.py [records ...]     (default: 666e6 = configs[1], 6e9 = 72 GB)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from krakenuniq_b200 import binding, synth_gpu  # noqa: E402

sizes = [int(float(a)) for a in sys.argv[1:]] or [666_000_000, 6_000_000_000]
B, L = 1_000_000, 150
dev = "cuda:0"
for records in sizes:
    t0 = time.time()
    db = synth_gpu.GpuDatabase(records, n_genomes=2000, k=31, nt=15, seed=2, device=dev,
                               passes=max(1, records // 600_000_000))
    pool, _ = db.sample_reads(2 * B, L, seed=3)
    torch.cuda.synchronize()
    gen_s = time.time() - t0
    torch.cuda.empty_cache()
    clf = binding.Classifier(device=0, n_slots=1, max_reads=B, max_bases=B * L + 4096, hll_mode=2)
    clf.attach_db_device(db.records.data_ptr(), db.key_ct, db.offsets.data_ptr(), 31, 15, 2)
    clf.set_taxonomy(*db.parent_map())
    d_off = torch.arange(B + 2, dtype=torch.int64, device=dev) * L
    for s in range(2):                                          # batch 0 = warm-up, batch 1 = the measured windows
        clf.classify_device(0, pool.data_ptr() + s * B * L, d_off.data_ptr(), B, B * L, None, flags=binding.F_NO_COUNTS)
        clf.sync(0)
    stage = clf.last_stage_ms(0)
    free_b, total_b = torch.cuda.mem_get_info()
    res = clf.layout_experiment(0, B * L, reps=5)
    res["database"] = {"records": int(db.key_ct), "record_gb_12B": db.key_ct * 12 / 1e9, "record_gb_8B": db.key_ct * 8 / 1e9,
                       "gen_s": round(gen_s, 1), "free_gb_before_transcode": round(free_b / 1e9, 1)}
    res["product_stage_ms"] = {"k_scan": stage[0], "k_lookup<MODE_FUSED>, lookup only (KUQ_F_NO_COUNTS)": stage[1],
                               "k_resolve": stage[2]}
    print(json.dumps(res), flush=True)
    base = res["variants"][0]["best_ms"]
    for v in res["variants"]:
        print(f"  {v['record_bytes']:2d} B records, {v['arity']:2d}-ary, scan <= {v['scan_window']:2d}, {v['shape']:46s}: "
              f"best {v['best_ms']:.3f} ms  mean {v['mean_ms']:.3f} ms  x{base / v['best_ms']:.2f}  mismatches {v['mismatches']}",
              flush=True)
    clf.close()
    del clf, db, pool, d_off
    torch.cuda.empty_cache()
