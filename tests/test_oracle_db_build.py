"""Database build tools: the oracle's db_sort and set_lcas against the unmodified reference
executables (oracle/_ref/db_sort, oracle/_ref/set_lcas) on the same inputs, byte for byte."""
import os

import numpy as np
import pytest

from krakenuniq_b200 import synth
from oracle import oracle_py

K = 31


def _library(seed, n_genomes=6, glen=1500):
    rng = np.random.default_rng(seed)
    tax = synth.make_taxonomy(n_genomes, 3, 2, first_id=100)
    sp = synth.species_ids(tax)
    genomes = synth.random_genomes(rng, n_genomes, glen, shared_frac=0.3)
    ks = []
    for g in genomes:
        km, ok = synth.forward_kmers(g, K)
        ks.append(synth.canonical(km[ok], K))
    allk = np.unique(np.concatenate(ks))
    rng.shuffle(allk)
    return rng, tax, sp, genomes, allk


def _need_ref():
    if not oracle_py.have_reference():
        pytest.skip("oracle/_ref not built (needs /root/reference)")


@pytest.mark.parametrize("nt,zero", [(5, True), (7, False), (9, True)])
def test_db_sort_matches_reference(oracle, tmp_path, nt, zero):
    _need_ref()
    rng, tax, sp, genomes, allk = _library(nt)
    rec = np.zeros(len(allk), synth._REC)
    rec["key"] = allk
    rec["taxon"] = rng.integers(0, 1 << 20, len(allk))
    jdb = np.concatenate([synth.kdb_header(K, len(allk)), rec.view(np.uint8)])
    jdb[100:140] = rng.integers(0, 255, 40)           # header bytes the tools ignore must be copied through
    jdb.tofile(tmp_path / "in.jdb")
    args = (["-z"] if zero else []) + ["-n", nt, "-d", "in.jdb", "-o", "out.kdb", "-i", "out.idx"]
    r = oracle_py.run_ref_tool("db_sort", args, cwd=tmp_path)
    assert r.returncode == 0, r.stderr
    kdb, idx = oracle.db_sort(jdb, nt, zero)
    assert np.array_equal(np.fromfile(tmp_path / "out.kdb", np.uint8), kdb)
    assert np.array_equal(np.fromfile(tmp_path / "out.idx", np.uint8), idx)
    # and the numpy builder used for the synthetic databases agrees (zeroed values / kept values)
    _, keys, vals = synth.parse_kdb(kdb)
    assert set(keys.tolist()) == set(allk.tolist())
    if zero:
        assert not vals.any()


def test_set_lcas_matches_reference(oracle, tmp_path):
    _need_ref()
    rng, tax, sp, genomes, allk = _library(21)
    # two extra library sequences: one whose taxid the taxonomy does not know (skipped, set_lcas.cpp:336-341) and
    # one that holds k-mers the database lacks (-x: ignored, :439-447); one id only matches without its ".2" suffix
    extra = rng.integers(0, 4, 700, dtype=np.uint8)
    names = [f"seq{i}" for i in range(len(genomes))] + ["unknown_tax", "novel.2"]
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
    synth.unsorted_jdb_image(allk, K).tofile(tmp_path / "database.jdb")
    r = oracle_py.run_ref_tool("db_sort", ["-z", "-n", 6, "-d", "database.jdb", "-o", "database0.kdb", "-i", "database.idx"], cwd=tmp_path)
    assert r.returncode == 0, r.stderr
    kdb0 = np.fromfile(tmp_path / "database0.kdb", np.uint8)
    idx = np.fromfile(tmp_path / "database.idx", np.uint8)
    r = oracle_py.run_ref_tool("set_lcas", ["-x", "-d", "database0.kdb", "-i", "database.idx", "-b", "taxDB", "-m",
                                            "seqid2taxid.map", "-F", "lib.fa"], cwd=tmp_path)
    assert r.returncode == 0, r.stderr
    want = np.fromfile(tmp_path / "database0.kdb", np.uint8)      # without -M the mmap'ed input is rewritten in place
    assert not np.array_equal(want, kdb0)

    ids, parents = tax.parent_map()
    known = set(int(t) for t in ids)
    pm = oracle.parent_map(ids, parents)
    db = oracle.open_db(kdb0.copy(), idx)
    missing = 0
    for s, t in zip(seqs, taxids):
        if t not in known:
            continue
        missing += oracle.set_lcas_sequence(db, pm, s, t)
    assert missing > 0
    assert np.array_equal(db.kdb, want)
    # the values are the LCA of the owners, as the numpy labeller computes them
    km, tx = synth.label_kmers(genomes + [np.concatenate([genomes[1][:300]])], list(sp) + [sp[3]], tax, K)
    _, keys, vals = synth.parse_kdb(db.kdb)
    got = dict(zip(keys.tolist(), vals.tolist()))
    for kk, t in zip(km.tolist(), tx.tolist()):
        assert got[kk] == t


def test_set_lcas_contaminant_and_reset_flags_match_reference(oracle, tmp_path):
    """set_lcas -T (what build_db.sh always passes: 'synthetic construct' 32630 / 'artificial sequences' 81077 stick,
    set_lcas.cpp:462-474) and -TR (values of the library's k-mers reset to 0, :458-459)"""
    _need_ref()
    rng, tax, sp, genomes, allk = _library(33)
    tax = synth.Taxonomy(tax.rows + [(32630, 1, "synthetic construct", "species"), (81077, 1, "artificial sequences", "species")])
    g = [synth.decode(x).tobytes() for x in genomes]
    novel1 = synth.decode(rng.integers(0, 4, 300, dtype=np.uint8)).tobytes()
    vec1 = g[2][200:500] + novel1 + g[0][900:1100]           # contaminant 1: parts of genomes 2 and 0
    vec2 = g[2][350:650] + g[4][100:400]                     # contaminant 2: overlaps vec1 on genome 2, plus genome 4
    names = ["seq0", "seq1", "seq2", "vec1", "seq3", "vec2", "seq4", "seq5"]
    seqs = [g[0], g[1], g[2], vec1, g[3], vec2, g[4], g[5]]
    taxids = [sp[0], sp[1], sp[2], 32630, sp[3], 81077, sp[4], sp[5]]
    ks = [allk]
    for s in (vec1,):
        km, ok = synth.forward_kmers(synth.encode(s), K)
        ks.append(synth.canonical(km[ok], K))
    allk2 = np.unique(np.concatenate(ks))
    rng.shuffle(allk2)
    with open(tmp_path / "lib.fa", "wb") as f:
        for n, s in zip(names, seqs):
            f.write(b">" + n.encode() + b"\n" + s + b"\n")
    with open(tmp_path / "seqid2taxid.map", "w") as f:
        for n, t in zip(names, taxids):
            f.write(f"{n}\t{t}\n")
    tax.write(str(tmp_path / "taxDB"))
    synth.unsorted_jdb_image(allk2, K).tofile(tmp_path / "database.jdb")
    r = oracle_py.run_ref_tool("db_sort", ["-z", "-n", 6, "-d", "database.jdb", "-o", "database0.kdb", "-i", "database.idx"], cwd=tmp_path)
    assert r.returncode == 0, r.stderr
    kdb0 = np.fromfile(tmp_path / "database0.kdb", np.uint8)
    idx = np.fromfile(tmp_path / "database.idx", np.uint8)
    ids, parents = tax.parent_map()
    pm = oracle.parent_map(ids, parents)

    def ref(flags_arg, fa="lib.fa"):
        r = oracle_py.run_ref_tool("set_lcas", ["-x", "-d", "database0.kdb", "-i", "database.idx", "-b", "taxDB", "-m",
                                                "seqid2taxid.map", "-F", fa, flags_arg], cwd=tmp_path)
        assert r.returncode == 0, r.stderr
        return np.fromfile(tmp_path / "database0.kdb", np.uint8)

    want = ref("-T")
    db = oracle.open_db(kdb0.copy(), idx)
    for s, t in zip(seqs, taxids):
        oracle.set_lcas_sequence(db, pm, s, t, flags=1)
    assert np.array_equal(db.kdb, want)
    _, keys, vals = synth.parse_kdb(db.kdb)
    got = dict(zip(keys.tolist(), vals.tolist()))
    km, ok = synth.forward_kmers(synth.encode(g[2][350:500]), K)      # shared by both contaminants: the first one wins
    assert {got[int(c)] for c in synth.canonical(km[ok], K).tolist()} == {32630}
    km, ok = synth.forward_kmers(synth.encode(g[4][100:400]), K)      # contaminant 2 before its genome in the file
    assert {got[int(c)] for c in synth.canonical(km[ok], K).tolist()} == {81077}

    # the hierarchical re-labelling of build_db.sh:286-296: -TR on a sub-library, then -T on it
    with open(tmp_path / "sub.fa", "wb") as f:
        for n, s in list(zip(names, seqs))[2:5]:
            f.write(b">" + n.encode() + b"\n" + s + b"\n")
    want = ref("-TR", "sub.fa")
    for s, t in list(zip(seqs, taxids))[2:5]:
        oracle.set_lcas_sequence(db, pm, s, t, flags=3)
    assert np.array_equal(db.kdb, want)
    want = ref("-T", "sub.fa")
    for s, t in list(zip(seqs, taxids))[2:5]:
        oracle.set_lcas_sequence(db, pm, s, t, flags=1)
    assert np.array_equal(db.kdb, want)
