#!/usr/bin/env python
"""Profile aid: classify N steps of the bench workload, then time the partitioned export of the sparse tier
(kuq_sparse_export_partitioned_alloc: k_keys_parts over the set and over the flagged records) the way the multi-GPU
merge calls it.  usage: python profiles/export_probe.py [steps] [n_parts]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from krakenuniq_b200 import binding, synth_gpu  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n_parts = int(sys.argv[2]) if len(sys.argv) > 2 else 8
B, L = 1_000_000, 150
db = synth_gpu.GpuDatabase(666_000_000, n_genomes=2000, k=31, nt=15, seed=2, device="cuda:0")
pool, _ = db.sample_reads(steps * B, L, seed=3)
clf = binding.Classifier(device=0, n_slots=2, max_reads=B, max_bases=B * L + 4096, hll_mode=0, sparse_set_slots=1 << 30)
clf.attach_db_device(db.records.data_ptr(), db.key_ct, db.offsets.data_ptr(), 31, 15, 2)
clf.set_taxonomy(*db.parent_map())
d_off = torch.arange(B + 2, dtype=torch.int64, device="cuda:0") * L
per_unit = -(-500000 // L)
unit = (torch.arange(B, dtype=torch.int64, device="cuda:0") // per_unit).to(torch.int32)
for s in range(steps):
    u = unit + s * 400
    clf.classify_device(0, pool.data_ptr() + s * B * L, d_off.data_ptr(), B, B * L, u.data_ptr())
    clf.sync(0)
torch.cuda.synchronize()
t0 = time.time()
ptr, counts = clf.sparse_export_partitioned_alloc(n_parts)
torch.cuda.synchronize()
t1 = time.time()
print(f"export of {int(counts.sum())} keys into {n_parts} parts: {(t1 - t0) * 1e3:.1f} ms (8 GB of records + 8 GB set, 2 passes each)", counts.tolist())
t0 = time.time()
clf.sparse_replace(ptr, int(counts.sum()))
torch.cuda.synchronize()
print(f"replace + import: {(time.time() - t0) * 1e3:.1f} ms")
clf.device_free(ptr)
