#!/usr/bin/env python
"""Profile aid: classify N steps of the bench workload, then run the end-of-run harvest (kuq_finish) so that ncu can
capture k_harvest_* in isolation.  usage: python profiles/harvest_probe.py [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from krakenuniq_b200 import binding, synth_gpu  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
B, L = 1_000_000, 150
db = synth_gpu.GpuDatabase(666_000_000, n_genomes=2000, k=31, nt=15, seed=2, device="cuda:0")
pool, _ = db.sample_reads(steps * B, L, seed=3)
clf = binding.Classifier(device=0, n_slots=2, max_reads=B, max_bases=B * L + 4096, hll_mode=0, sparse_set_slots=1 << 31)
clf.attach_db_device(db.records.data_ptr(), db.key_ct, db.offsets.data_ptr(), 31, 15, 2)
clf.set_taxonomy(*db.parent_map())
d_off = torch.arange(B + 2, dtype=torch.int64, device="cuda:0") * L
per_unit = -(-500000 // L)
unit = (torch.arange(B, dtype=torch.int64, device="cuda:0") // per_unit).to(torch.int32)
for s in range(steps):
    u = unit + s * 400
    clf.classify_device(0, pool.data_ptr() + s * B * L, d_off.data_ptr(), B, B * L, u.data_ptr())
    clf.sync(0)
clf.finish()
print(clf.sparse_tier_info())
