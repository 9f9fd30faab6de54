#!/usr/bin/env python
"""Golden fixtures for `classify -q -m N` with two databases (classify.cpp:928-936 inside the quick loop :943-944), made
with the UNMODIFIED reference: tests/golden/multidb/{ab_q2,ba_q1}.{kraken,report.tsv}.  Run after make_golden_multidb.py."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_py  # noqa: E402

MINI = os.path.join(ROOT, "tests", "golden", "mini")
OUT = os.path.join(ROOT, "tests", "golden", "multidb")


def main():
    a = ["-d", os.path.join(MINI, "database.kdb"), "-i", os.path.join(MINI, "database.idx")]
    b = ["-d", os.path.join(OUT, "db2.kdb"), "-i", os.path.join(OUT, "db2.idx")]
    for tag, dbs, m in (("ab_q2", a + b, 2), ("ba_q1", b + a, 1)):
        r = oracle_py.run_ref_tool("classify", dbs + ["-a", os.path.join(MINI, "taxDB"), "-M", "-t", 1, "-u", 20000, "-q", "-m", m,
                                   "-r", os.path.join(OUT, f"{tag}.report.tsv"), "-o", os.path.join(OUT, f"{tag}.kraken"),
                                   os.path.join(MINI, "reads.fa")], cwd=OUT)
        assert r.returncode == 0, r.stderr[-2000:]
        print(tag, open(os.path.join(OUT, f"{tag}.kraken")).read()[:200])


if __name__ == "__main__":
    main()
