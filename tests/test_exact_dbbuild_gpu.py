"""GPU tests of SURVEY.md §8 rows f3 (classifyExact / KUQ_HLL_EXACT) and f4 (GPU db_sort / set_lcas): the drop-in
executables and the C-ABI entry points against the oracle, whose CPU statements of these tools are pinned against the
reference executables (tests/test_oracle_db_build.py, tests/test_golden.py::test_oracle_exact_counting)."""
import os
import subprocess

import numpy as np
import pytest

from krakenuniq_b200 import build
from tests import util

pytestmark = [pytest.mark.gpu]
G = util.GOLDEN


def _report_lines(path):
    rows = {}
    for line in open(path):
        if line.startswith("#") or line.startswith("%"):
            continue
        f = line.rstrip("\n").split("\t")
        rows[f[6]] = f
    return rows


@pytest.mark.parametrize("extra,env", [(["-M"], None), (["-M", "-u", "20000"], None),
                                       (["-x", "40K"], {"KUQ_FORCE_CHUNKS": "1"})])
def test_cli_classify_exact(tmp_path, extra, env):
    """classifyExact (krakenuniq --exact): the report's kmers / dup / cov columns from exact distinct counts, equal to
    the unmodified reference's classifyExact (tests/golden/mini/exact.report.tsv); work units and chunking cannot
    change a set union, so every variant must give the same report.  Kraken lines as in the sketch build."""
    build.build_classify()
    exe = os.path.join(os.path.dirname(build.CLASSIFY), "classifyExact")
    out, rep = tmp_path / "exact.kraken", tmp_path / "exact.report.tsv"
    cmd = [exe, "-d", os.path.join(G, "database.kdb"), "-i", os.path.join(G, "database.idx"), "-a", os.path.join(G, "taxDB"),
           "-t", "1", "-r", str(rep), "-o", str(out)] + extra + [os.path.join(G, "reads.fa")]
    r = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, KUQ_SPARSE_SLOTS=str(1 << 22), **(env or {})))
    assert r.returncode == 0, r.stderr[-2000:]
    assert open(out).read() == open(os.path.join(G, "preload.kraken")).read()
    assert _report_lines(rep) == _report_lines(os.path.join(G, "exact.report.tsv"))


def test_exact_mode_matches_oracle_sets(oracle):
    """KUQ_HLL_EXACT through the C ABI on a seeded workload: per-taxon set sizes and clade unions equal the oracle's"""
    from krakenuniq_b200 import binding, synth
    rng = np.random.default_rng(11)
    tax = synth.make_taxonomy(12)
    sp = synth.species_ids(tax)
    genomes = synth.random_genomes(rng, 12, 3000, shared_frac=0.25)
    km, tx = synth.label_kmers(genomes, sp, tax, 31)
    kdb, idx = synth.build_db_images(km, tx, 31, 6, 1)
    bases, offs = synth.sample_reads(rng, genomes, 600, 150, 0.02, 0.2, 0.2)
    odb = oracle.open_db(kdb, idx)
    pm = oracle.parent_map(*tax.parent_map())
    run = oracle.run(odb, pm, 500000, 0)
    run.set_exact()
    calls, _, _ = run.classify(bases, offs, want_codes=False)
    run.finish()
    want = run.counts()
    clf = binding.Classifier(hll_mode=binding.HLL_EXACT, sparse_set_slots=1 << 20, max_reads=1 << 16, max_bases=16 << 20)
    clf.stage_db(kdb, idx)
    clf.set_taxonomy(*tax.parent_map())
    res = clf.classify(bases, offs)
    assert np.array_equal(res["call"], calls)
    clf.finish()
    got = clf.counts()
    assert np.array_equal(got["taxid"], want["taxid"])
    assert np.array_equal(got["n_kmers"], want["n_kmers"]) and np.array_equal(got["n_reads"], want["n_reads"])
    assert np.array_equal(got["unique"], want["unique"])
    members = util.clade_members(tax.rows, want["taxid"])
    for taxid, mem in members.items():
        assert clf.clade(mem) == run.clade(mem), taxid


@pytest.mark.parametrize("nt,zero", [(5, True), (7, False), (11, True)])
def test_gpu_db_sort_matches_oracle(oracle, nt, zero):
    """kuq_db_sort (CUB radix sorts + our minimizer / gather kernels) == the oracle's db_sort, which is pinned
    against the reference executable (tests/test_oracle_db_build.py)"""
    from krakenuniq_b200 import binding, synth
    rng = np.random.default_rng(nt)
    keys = np.unique(rng.integers(0, 1 << 62, 50000, dtype=np.uint64))
    rng.shuffle(keys)
    rec = np.zeros(len(keys), synth._REC)
    rec["key"], rec["taxon"] = keys, rng.integers(0, 1 << 20, len(keys))
    jdb = np.concatenate([synth.kdb_header(31, len(keys)), rec.view(np.uint8)])
    want_kdb, want_idx = oracle.db_sort(jdb, nt, zero)
    kdb, idx = binding.db_sort(jdb, nt, zero)
    assert np.array_equal(idx, want_idx)
    assert np.array_equal(kdb, want_kdb[:kdb.size])


def test_gpu_set_lcas_matches_oracle(oracle):
    """kuq_set_lcas_batch over SKIP_LEN-style pieces == the oracle's set_lcas (pinned against the reference tool)"""
    from krakenuniq_b200 import binding, synth
    rng = np.random.default_rng(5)
    tax = synth.make_taxonomy(8, 4, 2, first_id=100)
    sp = synth.species_ids(tax)
    genomes = synth.random_genomes(rng, 8, 2500, shared_frac=0.35)
    ks = []
    for g in genomes:
        km, ok = synth.forward_kmers(g, 31)
        ks.append(synth.canonical(km[ok], 31))
    allk = np.unique(np.concatenate(ks))
    drop = rng.random(len(allk)) < 0.05                      # some library k-mers are not in the database (-x)
    kdb0, idx = synth.build_db_images(allk[~drop], np.zeros(int((~drop).sum()), np.uint32), 31, 6, 2)
    seqs = [synth.decode(g).tobytes() for g in genomes]
    seqs[2] = seqs[2][:900] + b"N" + seqs[2][901:]
    ids, parents = tax.parent_map()
    pm = oracle.parent_map(ids, parents)
    odb = oracle.open_db(kdb0.copy(), idx)
    missing = sum(oracle.set_lcas_sequence(odb, pm, s, t) for s, t in zip(seqs, sp))
    # pieces of 400 bases overlapping by k-1, as process_single_file cuts them (set_lcas.cpp:363-364)
    pieces, taxids = [], []
    for s, t in zip(seqs, sp):
        for i in range(0, len(s), 400):
            pieces.append(s[i:i + 400 + 30])
            taxids.append(t)
    bases, offs = synth.pack_reads(pieces)
    clf = binding.Classifier(max_reads=1 << 16, max_bases=16 << 20)
    clf.stage_db(kdb0, idx)
    clf.set_taxonomy(ids, parents)
    got_missing = clf.set_lcas(bases, offs, taxids)
    out = kdb0.copy()
    clf.export_db_values(out)
    assert np.array_equal(out, odb.kdb)
    assert got_missing == missing


def test_cli_db_sort_and_set_lcas_rebuild_the_golden_database(tmp_path, oracle):
    """The drop-in db_sort + set_lcas executables on a seeded library: same database.kdb / .idx bytes as the oracle's
    (pinned) CPU statements of the two tools, same .counts file as count_taxons would write."""
    from krakenuniq_b200 import synth
    db_sort, set_lcas = build.build_dbtools()
    rng = np.random.default_rng(21)
    tax = synth.make_taxonomy(6, 3, 2, first_id=100)
    sp = synth.species_ids(tax)
    genomes = synth.random_genomes(rng, 6, 1500, shared_frac=0.3)
    ks = []
    for g in genomes:
        km, ok = synth.forward_kmers(g, 31)
        ks.append(synth.canonical(km[ok], 31))
    allk = np.unique(np.concatenate(ks))
    rng.shuffle(allk)
    jdb = synth.unsorted_jdb_image(allk, 31)
    jdb.tofile(tmp_path / "database.jdb")
    names = [f"seq{i}" for i in range(6)] + ["unknown_tax", "novel.2"]
    extra = rng.integers(0, 4, 700, dtype=np.uint8)
    seqs = [synth.decode(g).tobytes() for g in genomes] + [synth.decode(genomes[0][:400]).tobytes(),
                                                           synth.decode(genomes[1][:300]).tobytes() + b"N" + synth.decode(extra).tobytes()]
    taxids = list(sp) + [4242, sp[3]]
    with open(tmp_path / "lib.fa", "wb") as f:
        for n, s in zip(names, seqs):
            f.write(b">" + n.encode() + b" some description\n")
            for i in range(0, len(s), 61):
                f.write(s[i:i + 61] + b"\n")
    with open(tmp_path / "seqid2taxid.map", "w") as f:
        for n, t in zip(names, taxids):
            f.write(f"{n[:-2] if n.endswith('.2') else n}\t{t}\n")
    tax.write(str(tmp_path / "taxDB"))
    r = subprocess.run([db_sort, "-z", "-n", "6", "-d", "database.jdb", "-o", "database0.kdb", "-i", "database.idx"],
                       cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    want_kdb0, want_idx = oracle.db_sort(jdb, 6, True)
    assert np.array_equal(np.fromfile(tmp_path / "database.idx", np.uint8), want_idx)
    assert np.array_equal(np.fromfile(tmp_path / "database0.kdb", np.uint8), want_kdb0)
    r = subprocess.run([set_lcas, "-x", "-d", "database0.kdb", "-i", "database.idx", "-b", "taxDB", "-m", "seqid2taxid.map",
                        "-F", "lib.fa", "-c", "database.kdb.counts"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ids, parents = tax.parent_map()
    pm = oracle.parent_map(ids, parents)
    odb = oracle.open_db(want_kdb0.copy(), want_idx)
    known = set(int(t) for t in ids)
    for s, t in zip(seqs, taxids):
        if t in known:
            oracle.set_lcas_sequence(odb, pm, s, t)
    assert np.array_equal(np.fromfile(tmp_path / "database0.kdb", np.uint8), odb.kdb)
    # -TR then -T on a sub-library, as build_db.sh:286-296 re-labels hierarchically
    with open(tmp_path / "sub.fa", "wb") as f:
        for n, s in list(zip(names, seqs))[1:3]:
            f.write(b">" + n.encode() + b"\n" + s + b"\n")
    for fl, code in (("-TR", 3), ("-T", 1)):
        r = subprocess.run([set_lcas, "-x", "-d", "database0.kdb", "-i", "database.idx", "-b", "taxDB", "-m", "seqid2taxid.map",
                            "-F", "sub.fa", fl, "-c", "database.kdb.counts"], cwd=tmp_path, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        for s, t in list(zip(seqs, taxids))[1:3]:
            oracle.set_lcas_sequence(odb, pm, s, t, flags=code)
        assert np.array_equal(np.fromfile(tmp_path / "database0.kdb", np.uint8), odb.kdb), fl
    _, _, vals = synth.parse_kdb(odb.kdb)
    t, c = np.unique(vals, return_counts=True)
    assert open(tmp_path / "database.kdb.counts").read() == "".join(f"{a}\t{b}\n" for a, b in zip(t.tolist(), c.tolist()))
