#!/usr/bin/env python
"""Golden fixtures for the multi-database first-hit loop (classify.cpp:928-936) made with the UNMODIFIED reference:
   tests/golden/multidb/{db2.kdb,db2.idx,ab.kraken,ab.report.tsv,ba.kraken,ba.report.tsv}
DB1 = tests/golden/mini; DB2 = a second database (m = 5, KRAKIDX order) that shares records with DB1 under other
taxa, adds k-mers DB1 does not have (taken from the reads), and gives nonzero taxa to keys DB1 stores with taxon 0
(a stored 0 is a found key that stops the loop)."""
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from krakenuniq_b200 import synth  # noqa: E402
from oracle import oracle_py  # noqa: E402
from tests import util  # noqa: E402

MINI = os.path.join(ROOT, "tests", "golden", "mini")
OUT = os.path.join(ROOT, "tests", "golden", "multidb")


def main():
    rng = np.random.default_rng(42)
    os.makedirs(OUT, exist_ok=True)
    k, keys, taxa = synth.parse_kdb(np.fromfile(os.path.join(MINI, "database.kdb"), np.uint8))
    assert k == 31
    # 1. a third of DB1's records under genus-level taxa, 2. every DB1 key stored with taxon 0 → taxon 103,
    sel = rng.random(len(keys)) < 0.33
    k2 = [keys[sel]]
    t2 = [rng.choice(np.array([102, 103, 104, 1], np.uint32), int(sel.sum()))]
    zero = keys[(taxa == 0) & ~sel]
    k2.append(zero); t2.append(np.full(len(zero), 103, np.uint32))
    # 3. k-mers of the reads that DB1 does not hold (misses become hits), a few of them stored with taxon 0
    _, seqs = util.read_fasta(os.path.join(MINI, "reads.fa"))
    extra = []
    have = set(keys.tolist())
    for s in seqs[:400]:
        km, ok = synth.forward_kmers(synth.encode(s), 31)
        for c in synth.canonical(km[ok], 31).tolist():
            if c not in have:
                extra.append(c)
    extra = np.unique(np.array(extra, np.uint64))
    extra = extra[rng.random(len(extra)) < 0.5]
    et = rng.choice(np.array([101, 105, 0, 108], np.uint32), len(extra), p=[0.5, 0.3, 0.1, 0.1])
    k2.append(extra); t2.append(et)
    allk, first = np.unique(np.concatenate(k2), return_index=True)
    allt = np.concatenate(t2)[first]
    kdb, idx = synth.build_db_images(allk, allt, 31, 5, 1)
    kdb.tofile(os.path.join(OUT, "db2.kdb"))
    idx.tofile(os.path.join(OUT, "db2.idx"))
    a = ["-d", os.path.join(MINI, "database.kdb"), "-i", os.path.join(MINI, "database.idx")]
    b = ["-d", os.path.join(OUT, "db2.kdb"), "-i", os.path.join(OUT, "db2.idx")]
    for tag, dbs in (("ab", a + b), ("ba", b + a)):
        for f in ("db2.kdb.counts",):
            if os.path.exists(os.path.join(OUT, f)):
                os.remove(os.path.join(OUT, f))
        r = oracle_py.run_ref_tool("classify", dbs + ["-a", os.path.join(MINI, "taxDB"), "-M", "-t", 1, "-u", 20000, "-r",
                                   os.path.join(OUT, f"{tag}.report.tsv"), "-o", os.path.join(OUT, f"{tag}.kraken"),
                                   os.path.join(MINI, "reads.fa")], cwd=OUT)
        assert r.returncode == 0, r.stderr[-2000:]
    shutil.copy(os.path.join(OUT, "db2.kdb.counts"), os.path.join(OUT, "db2.kdb.counts.golden"))
    print(os.listdir(OUT))
    print(open(os.path.join(OUT, "ab.kraken")).read()[:300])


if __name__ == "__main__":
    main()
