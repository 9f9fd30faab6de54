#!/usr/bin/env python
"""Generate the committed golden fixtures with the UNMODIFIED reference built by oracle/build_ref.sh.

Run here (where /root/reference exists):   python tests/golden/make_golden.py
Outputs (small, committed):                 tests/golden/mini/*

Pipeline = SURVEY.md §8(c): synthetic `database.jdb` → reference `db_sort` → reference `set_lcas` → reference
`classify` in several modes.  The reference ships no golden vectors of its own, so these files are
the golden vectors; the GPU box only reads them.
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from krakenuniq_b200 import synth  # noqa: E402
from oracle import oracle_py  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "mini")
K, NT = 31, 6
SEED = 1


def write_fasta(path, names, seqs, width=None):
    with open(path, "wb") as f:
        for n, s in zip(names, seqs):
            f.write(b">" + n.encode() + b"\n")
            if width:
                for i in range(0, len(s), width):
                    f.write(s[i:i + width] + b"\n")
            else:
                f.write(s + b"\n")


def run(tool, args, cwd):
    r = oracle_py.run_ref_tool(tool, args, cwd=cwd)
    if r.returncode != 0:
        raise RuntimeError(f"{tool} failed: {r.stderr[-2000:]}")
    return r


def main():
    rng = np.random.default_rng(SEED)
    if os.path.exists(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    tmp = tempfile.mkdtemp(prefix="kuq_golden_")

    # taxonomy: root 1 → 2 families → 3 genera → 6 species; one extra species (id 900) lives in the DB only
    tax = synth.make_taxonomy(6, 3, 2, first_id=100)
    sp = synth.species_ids(tax)
    genomes = synth.random_genomes(rng, 6, 1800, shared_frac=0.2)
    # a 7th genome labelled with a taxid that taxDB does not contain
    orphan = rng.integers(0, 4, 900, dtype=np.uint8)
    tax.write(os.path.join(OUT, "taxDB"))

    # library FASTA + seqid→taxid map for set_lcas
    names = [f"seq{i}" for i in range(6)] + ["orphan"]
    seqs = [synth.decode(g).tobytes() for g in genomes] + [synth.decode(orphan).tobytes()]
    write_fasta(os.path.join(tmp, "lib.fa"), names, seqs, width=70)
    with open(os.path.join(tmp, "seqid2taxid.map"), "w") as f:
        for n, t in zip(names, sp + [900]):
            f.write(f"{n}\t{t}\n")

    # unsorted Jellyfish-style dump of all distinct canonical k-mers
    ks = []
    for g in genomes + [orphan]:
        km, ok = synth.forward_kmers(g, K)
        ks.append(synth.canonical(km[ok], K))
    allk = np.unique(np.concatenate(ks))
    rng.shuffle(allk)
    synth.unsorted_jdb_image(allk, K).tofile(os.path.join(tmp, "database.jdb"))

    run("db_sort", ["-z", "-n", NT, "-d", "database.jdb", "-o", "database0.kdb", "-i", "database.idx"], tmp)
    shutil.copy(os.path.join(OUT, "taxDB"), os.path.join(tmp, "taxDB"))
    run("set_lcas", ["-x", "-d", "database0.kdb", "-o", "database.kdb", "-i", "database.idx", "-b", "taxDB",
                     "-m", "seqid2taxid.map", "-F", "lib.fa"], tmp)
    if not os.path.exists(os.path.join(tmp, "database.kdb")):   # set_lcas without -M rewrites the input in place
        shutil.copy(os.path.join(tmp, "database0.kdb"), os.path.join(tmp, "database.kdb"))
    for f in ("database.kdb", "database.idx"):
        shutil.copy(os.path.join(tmp, f), os.path.join(OUT, f))

    # reads: sampled + edge cases
    reads, offs = synth.sample_reads(rng, genomes + [orphan], 1500, 150, 0.01, 0.3, 0.2)
    rs = [reads[int(offs[i]):int(offs[i + 1])].tobytes() for i in range(1500)]
    g0, g3 = seqs[0], seqs[3]
    edge = [b"ACGT", b"A" * 30, b"A" * 31, b"N" * 150, g0[100:250].lower(), g0[10:55] + b"N" + g3[200:245],
            g0[10:55] + b"N" + g3[200:246], g0[300:400] + g3[300:400] + g0[500:571], b"", g0[0:31], g0[0:32],
            g3[17:48] + b"NN" + g3[50:90], seqs[6][100:250], g0[100:400] + b"N" + g3[100:400]]
    rs = rs[:700] + edge + rs[700:]
    rnames = [f"r{i}" for i in range(len(rs))]
    write_fasta(os.path.join(OUT, "reads.fa"), rnames, rs)
    # the same reads as FASTQ (first 300) and a CRLF / multi-line FASTA variant (first 60)
    with open(os.path.join(OUT, "reads_300.fq"), "wb") as f:
        for n, s in list(zip(rnames, rs))[:300]:
            if len(s) == 0:
                continue
            f.write(b"@" + n.encode() + b" extra words\n" + s + b"\n+\n" + b"I" * len(s) + b"\n")
    with open(os.path.join(OUT, "reads_crlf_60.fa"), "wb") as f:
        for n, s in list(zip(rnames, rs))[:60]:
            f.write(b">" + n.encode() + b"\r\n")
            for i in range(0, len(s), 60):
                f.write(s[i:i + 60] + b"\r\n")

    db = ["-d", os.path.join(OUT, "database.kdb"), "-i", os.path.join(OUT, "database.idx"), "-a",
          os.path.join(OUT, "taxDB")]

    def classify(tag, extra, reads_file="reads.fa"):
        rep = os.path.join(OUT, f"{tag}.report.tsv")
        out = os.path.join(OUT, f"{tag}.kraken")
        r = run("classify", db + ["-t", 1, "-r", rep, "-o", out] + extra + [os.path.join(OUT, reads_file)], OUT)
        with open(os.path.join(OUT, f"{tag}.stderr.txt"), "w") as f:
            f.write("\n".join(l for l in r.stderr.replace("\r", "\n").split("\n")
                              if "processed in" in l or "sequences classified" in l or "unclassified" in l))
        return r

    classify("preload", ["-M"])                          # default work unit (500000 nt)
    classify("preload_u20000", ["-M", "-u", 20000])      # small work units: different sparse/dense outcomes
    classify("chunked", ["-x", "40K"])                   # --preload-size path (global sketches)
    classify("fastq", ["-M"], "reads_300.fq")
    classify("crlf", ["-M"], "reads_crlf_60.fa")
    # exact distinct counts (classifyExact) for HLL-error reporting
    rep = os.path.join(OUT, "exact.report.tsv")
    run("classifyExact", db + ["-M", "-t", 1, "-r", rep, "-o", "off", os.path.join(OUT, "reads.fa")], OUT)
    # the side effect file the reference creates on first report (classify.cpp:263-285)
    print("golden fixtures written to", OUT)
    subprocess.run(["ls", "-la", OUT])
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
