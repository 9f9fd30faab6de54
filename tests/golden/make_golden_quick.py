"""Golden outputs of the unmodified reference `classify` in quick mode (-q, -m N; classify.cpp:943-944,963-964,989-990).
Run here (needs /root/reference and oracle/_ref): python tests/golden/make_golden_quick.py
Reuses the mini database / reads of make_golden.py; writes tests/golden/quick/."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_py  # noqa: E402

MINI = os.path.join(ROOT, "tests", "golden", "mini")
OUT = os.path.join(ROOT, "tests", "golden", "quick")


def main():
    os.makedirs(OUT, exist_ok=True)
    db = ["-d", os.path.join(MINI, "database.kdb"), "-i", os.path.join(MINI, "database.idx"), "-a", os.path.join(MINI, "taxDB")]
    runs = [("q1", ["-M", "-q"], "reads.fa"), ("q3_u20000", ["-M", "-q", "-m", 3, "-u", 20000], "reads.fa"),
            ("q40", ["-M", "-q", "-m", 40], "reads.fa"), ("q2_crlf", ["-M", "-q", "-m", 2], "reads_crlf_60.fa"),
            ("q2_chunked", ["-x", "40K", "-q", "-m", 2], "reads.fa")]
    for tag, extra, reads in runs:
        r = oracle_py.run_ref_tool("classify", db + ["-t", 1, "-r", os.path.join(OUT, f"{tag}.report.tsv"), "-o",
                                   os.path.join(OUT, f"{tag}.kraken")] + extra + [os.path.join(MINI, reads)], cwd=OUT)
        assert r.returncode == 0, r.stderr[-2000:]
        print(tag, open(os.path.join(OUT, f"{tag}.kraken")).readline().strip())


if __name__ == "__main__":
    main()
