"""GPU parity tests proper: the CUDA path, called through the C ABI (libkuq.so), against the oracle on the same
seeded inputs and against the golden outputs of the unmodified reference.  Bit-exact: calls, per-window codes,
hit lists, per-taxon read/k-mer counters and HLL registers."""
import os

import numpy as np
import pytest

from krakenuniq_b200 import binding, synth
from tests import util

pytestmark = pytest.mark.gpu
K = 31


def _classifier(**kw):
    kw.setdefault("max_reads", 1 << 16)
    kw.setdefault("max_bases", 16 << 20)
    kw.setdefault("sparse_set_slots", 1 << 22)
    return binding.Classifier(**kw)


def _runs_to_codes(res, i):
    out = []
    for code, cnt in binding.decode_runs(res, i):
        out += [code] * cnt
    return np.array(out, np.uint32)


def _check_against_oracle(oracle, kdb, idx, tax, bases, offs, hll_mode=binding.HLL_DENSE_ONLY, unit=500000,
                          oracle_mode=0):
    db = oracle.open_db(kdb, idx)
    pm = oracle.parent_map(*tax.parent_map())
    run = oracle.run(db, pm, unit, oracle_mode)
    calls, codes, code_off = run.classify(bases, offs)
    run.finish()
    want = run.counts(want_regs=True)

    clf = _classifier(hll_mode=hll_mode, work_unit_size=unit)
    clf.stage_db(kdb, idx)
    clf.set_taxonomy(*tax.parent_map())
    res = clf.classify(bases, offs, flags=binding.F_WANT_CODES)
    clf.finish()
    n = len(offs) - 1
    assert np.array_equal(res["call"], calls)
    nwin = np.diff(code_off).astype(np.uint32)
    assert np.array_equal(res["n_windows"], nwin)
    base0 = int(offs[0])
    for i in range(n):
        w = codes[int(code_off[i]):int(code_off[i + 1])]
        o = int(offs[i]) - base0
        got = res["codes"][o:o + len(w)]
        assert np.array_equal(got, w), f"read {i}"
        assert np.array_equal(_runs_to_codes(res, i), w), f"runs of read {i}"
    assert res["n_classified"] == int((calls != 0).sum())
    got = clf.counts()
    assert np.array_equal(got["taxid"], want["taxid"])
    assert np.array_equal(got["n_reads"], want["n_reads"])
    assert np.array_equal(got["n_kmers"], want["n_kmers"])
    for j, t in enumerate(want["taxid"]):
        assert np.array_equal(clf.registers(int(t)), want["regs"][j]), f"registers of taxon {t}"
    if hll_mode != binding.HLL_DENSE_ONLY:
        # the reference's sparse→dense mode rule reproduced exactly: same tier, same estimate
        assert np.array_equal(got["sparse"], want["sparse"]), (got["sparse"], want["sparse"])
        assert np.array_equal(got["unique"], want["unique"]), (got["unique"], want["unique"])
    return clf, got, want


def _synthetic(seed, nt, idx_type, n_genomes=5, glen=2500, n_reads=800, read_len=150, n_frac=0.2):
    rng = np.random.default_rng(seed)
    tax = synth.make_taxonomy(n_genomes)
    sp = synth.species_ids(tax)
    genomes = synth.random_genomes(rng, n_genomes, glen, shared_frac=0.25)
    km, tx = synth.label_kmers(genomes, sp, tax, K)
    kdb, idx = synth.build_db_images(km, tx, K, nt, idx_type)
    bases, offs = synth.sample_reads(rng, genomes, n_reads, read_len, 0.01, n_frac, 0.2)
    return tax, genomes, kdb, idx, bases, offs


@pytest.mark.parametrize("nt,idx_type", [(8, 2), (6, 1), (11, 2), (15, 2)])
def test_synthetic_db_matches_oracle(oracle, nt, idx_type):
    tax, genomes, kdb, idx, bases, offs = _synthetic(100 + nt, nt, idx_type)
    _check_against_oracle(oracle, kdb, idx, tax, bases, offs)


def test_edge_reads_match_oracle(oracle):
    tax, genomes, kdb, idx, _, _ = _synthetic(7, 8, 2)
    g = [synth.decode(x).tobytes() for x in genomes]
    rng = np.random.default_rng(3)
    long_read = b"".join(g[int(rng.integers(0, 5))][s:s + 97] for s in rng.integers(0, 2300, 40).tolist())
    seqs = [b"", b"A", b"ACGT" * 7 + b"AC", b"ACGT" * 7 + b"ACG", g[0][:31], g[0][:32], g[0][5:36].lower(),
            b"N" * 31, b"N" * 200, g[1][100:163], g[1][100:164], g[1][100:165], g[1][0:95], g[1][0:96], g[1][0:97],
            g[2][10:130] + b"N" + g[3][10:130], g[2][10:40] + b"R" + g[2][41:200], long_read, g[4][:2500],
            g[0][50:200] + b"\r", g[0][50:110] + b"\r" + g[0][110:170] + b"\r", g[0][50:80] + b"\r",
            g[0][50:79] + b"\n\r", b"\r" * 40, g[3][7:38] + b"\r\r\r", g[0][100:131] + b"acgtnACGT" * 10]
    seqs += [g[i % 5][s:s + ln] for i, (s, ln) in enumerate(zip(rng.integers(0, 2000, 200).tolist(),
                                                                rng.integers(1, 400, 200).tolist()))]
    bases, offs = synth.pack_reads(seqs)
    _check_against_oracle(oracle, kdb, idx, tax, bases, offs)


def test_orphan_taxa_ties_and_forest(oracle):
    """taxids stored in the DB but absent from taxDB, several roots, ties → LCA"""
    rng = np.random.default_rng(11)
    rows = [(1, 1, "root", "no rank"), (10, 1, "a", "genus"), (11, 10, "a1", "species"), (12, 10, "a2", "species"),
            (20, 1, "b", "genus"), (21, 20, "b1", "species"), (500, 500, "other root", "no rank"),
            (501, 500, "c1", "species"), (600, 999, "dangling parent", "species")]
    tax = synth.Taxonomy(rows)
    labels = [11, 12, 21, 501, 600, 4242, 77]          # 4242 and 77 are not in taxDB
    genomes = [rng.integers(0, 4, 1200, dtype=np.uint8) for _ in labels]
    ks, ts = [], []
    for gseq, t in zip(genomes, labels):
        km, ok = synth.forward_kmers(gseq, K)
        c = np.unique(synth.canonical(km[ok], K))
        ks.append(c); ts.append(np.full(len(c), t, np.uint32))
    km, first = np.unique(np.concatenate(ks), return_index=True)
    tx = np.concatenate(ts)[first]
    kdb, idx = synth.build_db_images(km, tx, K, 7, 2)
    g = [synth.decode(x).tobytes() for x in genomes]
    seqs = []
    for a in range(len(g)):
        for b in range(len(g)):
            for la, lb in [(45, 45), (45, 46), (60, 40)]:
                seqs.append(g[a][100:100 + la] + b"N" + g[b][300:300 + lb])
    seqs += [g[0][0:50] + b"N" + g[1][0:50] + b"N" + g[2][0:50] + b"N" + g[3][0:50] + b"N" + g[5][0:50]]
    bases, offs = synth.pack_reads(seqs)
    _check_against_oracle(oracle, kdb, idx, tax, bases, offs)


@pytest.mark.parametrize("tag,reads", [("preload", "reads.fa"), ("fastq", "reads_300.fq"), ("crlf", "reads_crlf_60.fa")])
def test_golden_reference_outputs(oracle, tag, reads):
    """Kraken lines byte-identical to the unmodified reference's `classify -M -t 1` output."""
    kdb = np.fromfile(os.path.join(util.GOLDEN, "database.kdb"), np.uint8)
    idx = np.fromfile(os.path.join(util.GOLDEN, "database.idx"), np.uint8)
    tax = synth.Taxonomy.read(os.path.join(util.GOLDEN, "taxDB"))
    path = os.path.join(util.GOLDEN, reads)
    ids, seqs = util.read_fastq(path) if reads.endswith(".fq") else util.read_fasta(path)
    bases, offs = synth.pack_reads(seqs)
    clf = _classifier(hll_mode=binding.HLL_DENSE_ONLY)
    clf.stage_db(kdb, idx)
    clf.set_taxonomy(*tax.parent_map())
    res = clf.classify(bases, offs)
    lines = []
    for i, (rid, s) in enumerate(zip(ids, seqs)):
        runs = binding.decode_runs(res, i)
        hl = " ".join(("A" if c == binding.AMBIG else str(c)) + ":" + str(n) for c, n in runs) if runs else "0:0"
        call = int(res["call"][i])
        lines.append(f"{'C' if call else 'U'}\t{rid}\t{call}\t{len(s)}\t{hl}\n")
    assert "".join(lines) == open(os.path.join(util.GOLDEN, f"{tag}.kraken")).read()
    # per-taxon read counts = taxReads column of the reference's report
    rep = util.parse_report(os.path.join(util.GOLDEN, f"{tag}.report.tsv"))
    cnt = clf.counts()
    by_tax = dict(zip(cnt["taxid"].tolist(), cnt["n_reads"].tolist()))
    for t, row in rep.items():
        assert row["taxReads"] == by_tax.get(t, 0)
    # the database scan reproduces database.kdb.counts (KrakenDB::count_taxons)
    t, c = clf.db_taxids()
    want = [tuple(int(x) for x in l.split()) for l in open(os.path.join(util.GOLDEN, "database.kdb.counts"))]
    assert list(zip(t.tolist(), c.tolist())) == want


def test_batches_and_slots_are_equivalent(oracle):
    """Splitting the input over batches / slots changes nothing (counters are order independent)."""
    tax, genomes, kdb, idx, bases, offs = _synthetic(21, 9, 2, n_reads=1000)
    clf1, got1, _ = _check_against_oracle(oracle, kdb, idx, tax, bases, offs)
    clf = _classifier(hll_mode=binding.HLL_DENSE_ONLY, n_slots=3)
    clf.stage_db(kdb, idx)
    clf.set_taxonomy(*tax.parent_map())
    calls = []
    cuts = [0, 1, 130, 131, 700, 1000]
    pending = []
    for j in range(len(cuts) - 1):
        o = np.ascontiguousarray(offs[cuts[j]:cuts[j + 1] + 1])
        slot = j % 3
        if len(pending) == 3:
            s0, o0 = pending.pop(0)
            calls.append(clf.wait(s0, o0, copy=True)["call"])
        clf.submit(slot, bases.ctypes.data, o)
        pending.append((slot, o))
    for s0, o0 in pending:
        calls.append(clf.wait(s0, o0, copy=True)["call"])
    clf.finish()
    assert np.array_equal(np.concatenate(calls), clf1.classify(bases, offs)["call"])
    got = clf.counts()
    for key in ("taxid", "n_reads", "n_kmers", "unique"):
        assert np.array_equal(got[key], got1[key]), key


@pytest.mark.parametrize("unit,mode", [(500000, 0), (1350, 0), (1500, 0), (2000, 0), (4000, 0), (20000, 0),
                                       (500000, 1)])
def test_hll_mode_rule_matches_oracle(oracle, unit, mode):
    """uniqueKmerCount under the reference's sparse/dense rule: per-work-unit sketches (preload)
    and one global sketch (chunked).  Small units put many (unit, taxon) pairs right around the 1024 threshold."""
    tax, genomes, kdb, idx, bases, offs = _synthetic(300 + unit % 97, 8, 2, n_genomes=3, glen=1600, n_reads=1500,
                                                     n_frac=0.05)
    _check_against_oracle(oracle, kdb, idx, tax, bases, offs,
                          hll_mode=binding.HLL_PRELOAD if mode == 0 else binding.HLL_CHUNKED, unit=unit,
                          oracle_mode=mode)


def test_hll_threshold_corner(oracle):
    """D == 1024 exactly: the sketch converts only if the last insert repeats an earlier code."""
    rng = np.random.default_rng(5)
    g = rng.integers(0, 4, 1024 + 30, dtype=np.uint8)          # 1024 windows → (almost surely) 1024 distinct codes
    tax = synth.make_taxonomy(1)
    sp = synth.species_ids(tax)
    km, tx = synth.label_kmers([g], sp, tax, K)
    kdb, idx = synth.build_db_images(km, tx, K, 7, 2)
    whole = synth.decode(g).tobytes()
    for tail in (whole[:60], whole[-60:], whole[500:560]):
        # one unit: all 1024 distinct k-mers, then one more read that only repeats k-mers
        seqs = [whole, tail]
        bases, offs = synth.pack_reads(seqs)
        _check_against_oracle(oracle, kdb, idx, tax, bases, offs, hll_mode=binding.HLL_PRELOAD, unit=10 ** 9)
    # the last read introduces the 1024th distinct code as the very last insert: stays sparse
    seqs = [whole[:1024 + 29], whole[-31:]]
    bases, offs = synth.pack_reads(seqs)
    _check_against_oracle(oracle, kdb, idx, tax, bases, offs, hll_mode=binding.HLL_PRELOAD, unit=10 ** 9)


def test_sparse_tier_grows_at_the_harvest_and_survives_several_readouts(oracle):
    """The fused path only flags the records of its hits; kuq_finish / kuq_read_counts harvest the flags into the
    (taxon, code) set.  Start with a set far too small (1024 slots): the harvest must grow it, and reading the
    counts between batches (harvest, more batches, harvest again) must not change the final state."""
    rng = np.random.default_rng(21)
    tax = synth.make_taxonomy(6)
    genomes = synth.random_genomes(rng, 6, 3000, shared_frac=0.25)
    km, tx = synth.label_kmers(genomes, synth.species_ids(tax), tax, K)
    kdb, idx = synth.build_db_images(km, tx, K, 8, 2)
    # few misses (they go straight into the set, which must not saturate before the first harvest), many hits
    bases, offs = synth.sample_reads(rng, genomes, 900, 150, 0.0005, 0.2, 0.01)
    db = oracle.open_db(kdb, idx)
    pm = oracle.parent_map(*tax.parent_map())
    run = oracle.run(db, pm, 1200, 0)                  # work units of 8 reads: no (unit, taxon) sketch ever converts,
    run.classify(bases, offs, want_codes=False)        # so every taxon stays in the sparse tier
    run.finish()
    want = run.counts()
    assert want["sparse"].all()
    clf = _classifier(hll_mode=binding.HLL_PRELOAD, sparse_set_slots=1 << 13, work_unit_size=1200)
    clf.stage_db(kdb, idx)
    clf.set_taxonomy(*tax.parent_map())
    n = len(offs) - 1
    cuts = [0, 296, 296, 600, n]                                    # unit boundaries; includes an empty batch
    unit = (np.arange(n) // 8).astype(np.uint32)                    # the units the oracle cut
    for a, b in zip(cuts[:-1], cuts[1:]):
        clf.classify(bases, offs[a:b + 1], unit_id=unit[a:b])
        clf.counts()                                                # readout in the middle of the run = a harvest
    clf.finish()
    got = clf.counts()
    info = clf.sparse_tier_info()
    assert info["times_grown"] >= 1 and info["slots"] > (1 << 13) and info["keys"] <= info["slots"] * 0.7 + 1
    assert np.array_equal(got["taxid"], want["taxid"])
    assert np.array_equal(got["n_kmers"], want["n_kmers"]) and np.array_equal(got["n_reads"], want["n_reads"])
    assert np.array_equal(got["sparse"], want["sparse"])
    assert np.array_equal(got["unique"], want["unique"])
    members = util.clade_members(tax.rows, want["taxid"])
    for taxid, mem in members.items():
        assert clf.clade(mem) == run.clade(mem), taxid
    # reset wipes the record flags too: a second identical run gives the same state, not a union with stale flags
    clf.reset_counts()
    clf.classify(bases, offs, unit_id=unit)
    clf.finish()
    again = clf.counts()
    assert np.array_equal(again["unique"], want["unique"]) and np.array_equal(again["n_kmers"], want["n_kmers"])


@pytest.mark.parametrize("tag,unit,mode", [("preload", 500000, binding.HLL_PRELOAD),
                                           ("preload_u20000", 20000, binding.HLL_PRELOAD),
                                           ("chunked", 500000, binding.HLL_CHUNKED)])
def test_golden_report_kmers_column(tag, unit, mode):
    """reads / kmers / dup columns of the unmodified reference's report, for every clade, exactly."""
    kdb = np.fromfile(os.path.join(util.GOLDEN, "database.kdb"), np.uint8)
    idx = np.fromfile(os.path.join(util.GOLDEN, "database.idx"), np.uint8)
    tax = synth.Taxonomy.read(os.path.join(util.GOLDEN, "taxDB"))
    ids, seqs = util.read_fasta(os.path.join(util.GOLDEN, "reads.fa"))
    bases, offs = synth.pack_reads(seqs)
    clf = _classifier(hll_mode=mode, work_unit_size=unit)
    clf.stage_db(kdb, idx)
    clf.set_taxonomy(*tax.parent_map())
    clf.classify(bases, offs)
    clf.finish()
    cnt = clf.counts()
    rep = util.parse_report(os.path.join(util.GOLDEN, f"{tag}.report.tsv"))
    members = util.clade_members(tax.rows, cnt["taxid"])
    for taxid, row in rep.items():
        u, r, k = clf.clade(members[taxid])
        assert (row["reads"], row["kmers"]) == (r, u), (taxid, row, r, u, k)


@pytest.mark.parametrize("hll_mode,unit", [(binding.HLL_DENSE_ONLY, 500000), (binding.HLL_PRELOAD, 5000)])
def test_reads_with_more_than_32_taxa(oracle, hll_mode, unit):
    """long reads / contigs: the hit list leaves the registers for a hash table in HBM"""
    rng = np.random.default_rng(77)
    n_sp = 90
    tax = synth.make_taxonomy(n_sp, 12, 4)
    sp = synth.species_ids(tax)
    genomes = [rng.integers(0, 4, 400, dtype=np.uint8) for _ in range(n_sp)]
    km, tx = synth.label_kmers(genomes, sp, tax, K)
    kdb, idx = synth.build_db_images(km, tx, K, 7, 2)
    g = [synth.decode(x).tobytes() for x in genomes]
    seqs = []
    for n_taxa in (33, 40, 64, 90):                 # pieces of 45 bases → 15 windows per species
        order = rng.permutation(n_sp)[:n_taxa]
        seqs.append(b"".join(g[i][50:95] for i in order))
        seqs.append(b"".join(g[i][50:95 + (j % 3)] for j, i in enumerate(order)))      # unequal counts → no tie
    seqs.append(b"".join(g[i][10:390] for i in range(n_sp)))                          # a 34 kb "contig"
    seqs += [g[i][0:150] for i in range(10)]                                            # ordinary reads around them
    bases, offs = synth.pack_reads(seqs)
    _check_against_oracle(oracle, kdb, idx, tax, bases, offs, hll_mode=hll_mode, unit=unit)


def test_paired_reads_merged_with_N(oracle):
    """--paired input as read_merger.pl writes it: mate1 + 'N' + mate2 (scripts/read_merger.pl:187-191)"""
    tax, genomes, kdb, idx, _, _ = _synthetic(55, 9, 2, n_genomes=6)
    rng = np.random.default_rng(9)
    g = [synth.decode(x).tobytes() for x in genomes]
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    seqs = []
    for _ in range(400):
        i = int(rng.integers(0, 6)); s = int(rng.integers(0, 2000))
        m1 = g[i][s:s + 150]
        m2 = g[i][s + 250:s + 400].translate(comp)[::-1]
        seqs.append(m1 + b"N" + m2)
    bases, offs = synth.pack_reads(seqs)
    _check_against_oracle(oracle, kdb, idx, tax, bases, offs, hll_mode=binding.HLL_PRELOAD, unit=30000)


@pytest.mark.parametrize("n_ranges,slots", [(2, 1 << 22), (5, 1 << 10)])
def test_database_ranges_through_the_halves(oracle, n_ranges, slots):
    """chunked database (classify.cpp:566-791) through the C ABI halves: stage one minimizer range at a time,
    kuq_lookup_batch for every range, element-wise max merge, kuq_resolve_batch — same calls, hit lists and
    (chunked-rule) counters as the oracle's one-pass run."""
    tax, genomes, kdb, idx, bases, offs = _synthetic(91, 9, 2, n_genomes=5, n_reads=700)
    db = oracle.open_db(kdb, idx)
    pm = oracle.parent_map(*tax.parent_map())
    run = oracle.run(db, pm, 500000, 1)
    calls, codes, code_off = run.classify(bases, offs)
    run.finish()
    want = run.counts(want_regs=True)

    n_bins = 1 << 18
    idx_off = np.frombuffer(idx[8:].tobytes(), np.uint64)
    cuts = [int(np.searchsorted(idx_off, idx_off[-1] * r // n_ranges)) for r in range(n_ranges + 1)]
    cuts[0], cuts[-1] = 0, n_bins
    # slots = 1024: the resolve half inserts its (taxon, code) pairs directly; the set must grow before the batch
    clf = _classifier(hll_mode=binding.HLL_CHUNKED, sparse_set_slots=slots)
    clf.set_taxonomy(*tax.parent_map())
    universe = set()
    for r in range(n_ranges):                       # pass 0: the taxids of every range
        if cuts[r + 1] > cuts[r]:
            clf.stage_db(kdb, idx, cuts[r], cuts[r + 1])
            universe |= set(clf.db_taxids()[0].tolist())
    clf.set_db_taxid_universe(np.array(sorted(universe), np.uint32))
    merged = np.zeros(int(offs[-1]), np.uint32)
    for r in range(n_ranges):
        if cuts[r + 1] <= cuts[r]:
            continue
        clf.stage_db(kdb, idx, cuts[r], cuts[r + 1])
        part, nwin = clf.lookup(bases, offs)
        amb = part == binding.AMBIG
        merged = np.where(amb, merged, np.maximum(merged, part))
    res = clf.resolve(bases, offs, merged, flags=binding.F_WANT_CODES)
    clf.finish()
    assert np.array_equal(res["call"], calls)
    assert np.array_equal(res["n_windows"], np.diff(code_off).astype(np.uint32))
    for i in range(len(offs) - 1):
        w = codes[int(code_off[i]):int(code_off[i + 1])]
        assert np.array_equal(_runs_to_codes(res, i), w), f"runs of read {i}"
    got = clf.counts()
    for key in ("taxid", "n_reads", "n_kmers", "sparse", "unique"):
        assert np.array_equal(got[key], want[key]), key
    if slots < (1 << 12):
        assert clf.sparse_tier_info()["times_grown"] >= 1


@pytest.mark.parametrize("n_ranges,n_batches", [(3, 2), (6, 3)])
def test_database_streamed_through_two_buffers(oracle, n_ranges, n_batches):
    """kuq_stream_*: the database arrives range by range from pinned host memory into two device buffers while the
    lookups of the previous range run; ids merge on the device (only_hits stores into one buffer per batch); the final
    pass resolves.  Same calls, hit lists and chunked-rule counters as the oracle's one-pass run (classify -x)."""
    import ctypes
    import torch
    tax, genomes, kdb, idx, bases, offs = _synthetic(93, 9, 2, n_genomes=6, n_reads=900)
    db = oracle.open_db(kdb, idx)
    pm = oracle.parent_map(*tax.parent_map())
    run = oracle.run(db, pm, 500000, 1)
    calls, codes, code_off = run.classify(bases, offs)
    run.finish()
    want = run.counts()

    _, keys, taxa = synth.parse_kdb(kdb)
    n_bins = 1 << 18
    idx_off = np.frombuffer(idx[8:].tobytes(), np.uint64)
    cuts = [int(np.searchsorted(idx_off, idx_off[-1] * r // n_ranges)) for r in range(n_ranges + 1)]
    cuts[0], cuts[-1] = 0, n_bins
    header = kdb.size - 12 * len(keys)
    clf = _classifier(hll_mode=binding.HLL_CHUNKED, n_slots=2)
    clf.set_taxonomy(*tax.parent_map())
    clf.set_db_taxid_universe(np.unique(taxa))
    max_rec = max(int(idx_off[cuts[r + 1]] - idx_off[cuts[r]]) for r in range(n_ranges))
    max_bins = max(cuts[r + 1] - cuts[r] for r in range(n_ranges))
    clf.stream_open(31, 9, 2, max(max_rec, 1), max_bins)
    # pinned host copies of the record body and the offsets
    L = clf.L
    rec_bytes = kdb[header:]
    h_rec = L.kuq_host_alloc(rec_bytes.size + 16)
    ctypes.memmove(h_rec, rec_bytes.ctypes.data, rec_bytes.size)
    h_off = L.kuq_host_alloc(idx_off.size * 8)
    ctypes.memmove(h_off, idx_off.ctypes.data, idx_off.size * 8)

    def load(buf, r):
        a, b = int(idx_off[cuts[r]]), int(idx_off[cuts[r + 1]])
        clf.stream_load(buf, h_rec + 12 * a, b - a, h_off + 8 * cuts[r], cuts[r], cuts[r + 1])

    dev = "cuda:0"
    total = int(offs[-1])
    d_bases = torch.from_numpy(np.concatenate([bases, np.full(64, ord("N"), np.uint8)])).to(dev)
    d_offs = torch.from_numpy(np.concatenate([offs, offs[-1:]]).astype(np.int64)).to(dev)
    merged = torch.zeros(total + 64, dtype=torch.int32, device=dev)
    n = len(offs) - 1
    bcuts = [2 * (n * i // n_batches // 2) for i in range(n_batches)] + [n]     # even cuts: 16-byte aligned offset slices
    ranges = [r for r in range(n_ranges) if cuts[r + 1] > cuts[r]]
    load(0, ranges[0])
    for j, r in enumerate(ranges):
        clf.stream_use(j & 1)
        if j + 1 < len(ranges):
            load((j + 1) & 1, ranges[j + 1])
        for bi in range(n_batches):
            a, b = bcuts[bi], bcuts[bi + 1]
            clf.lookup_device(bi & 1, d_bases.data_ptr(), d_offs.data_ptr() + 8 * a, b - a, total, merged.data_ptr(), only_hits=1)
    clf.sync(0); clf.sync(1)
    clf.stream_check()
    call = np.zeros(n, np.uint32)
    for bi in range(n_batches):
        a, b = bcuts[bi], bcuts[bi + 1]
        clf.resolve_device(0, d_bases.data_ptr(), d_offs.data_ptr() + 8 * a, b - a, total, merged.data_ptr(), None)
        clf.sync(0)
        r_ = clf.device_result(0)
        from krakenuniq_b200 import dist as kdist
        call[a:b] = kdist.device_view(r_.d_call, (b - a) * 4, torch.int32, dev).cpu().numpy().view(np.uint32)
    clf.finish()
    assert np.array_equal(call, calls)
    got = clf.counts()
    for key in ("taxid", "n_reads", "n_kmers", "sparse", "unique"):
        assert np.array_equal(got[key], want[key]), key
    L.kuq_host_free(h_rec); L.kuq_host_free(h_off)


def test_two_contexts_merge_into_one(oracle):
    """kuq_merge_into: what the multi-GPU `classify` does with its per-device contexts (replicas) — here both contexts
    live on one device.  Each classifies every other work unit; the merged state equals the oracle's single run."""
    tax, genomes, kdb, idx, bases, offs = _synthetic(33, 8, 2, n_genomes=6, glen=3000, n_reads=1200)
    unit = 6000
    db = oracle.open_db(kdb, idx)
    pm = oracle.parent_map(*tax.parent_map())
    run = oracle.run(db, pm, unit, 0)
    calls, _, _ = run.classify(bases, offs, want_codes=False)
    run.finish()
    want = run.counts()
    units, _, _ = synth.work_unit_ids(offs, unit)
    n = len(offs) - 1
    a, b = _classifier(hll_mode=binding.HLL_PRELOAD, work_unit_size=unit, sparse_set_slots=1 << 14), \
        _classifier(hll_mode=binding.HLL_PRELOAD, work_unit_size=unit, sparse_set_slots=1 << 14)
    for c in (a, b):
        c.stage_db(kdb, idx)
        c.set_taxonomy(*tax.parent_map())
    got_calls = np.zeros(n, np.uint32)
    # runs of whole units alternate between the two contexts
    bounds = [0] + [i for i in range(1, n) if units[i] != units[i - 1]] + [n]
    for j, (lo, hi) in enumerate(zip(bounds[:-1], bounds[1:])):
        res = (a, b)[j & 1].classify(bases, np.ascontiguousarray(offs[lo:hi + 1]), unit_id=units[lo:hi])
        got_calls[lo:hi] = res["call"]
    assert np.array_equal(got_calls, calls)
    a.finish(); b.finish()
    a.merge_from(b)
    got = a.counts()
    for key in ("taxid", "n_reads", "n_kmers", "sparse", "unique"):
        assert np.array_equal(got[key], want[key]), key
    members = util.clade_members(tax.rows, want["taxid"])
    for taxid, mem in members.items():
        assert a.clade(mem) == run.clade(mem), taxid


def _clade_tree_equals_per_clade(clf, tax_rows):
    cnt = clf.counts()
    members = util.clade_members(tax_rows, cnt["taxid"])
    ids = np.array(sorted(members), np.uint32)
    extra = np.array([4000000000, 7], np.uint32)                    # unknown / uncounted taxids give zeros
    r, k, u = clf.clade_counts_tree(np.concatenate([ids, extra]))
    assert not r[len(ids):].any() and not u[len(ids):].any()
    kinds = set()
    for i, t in enumerate(ids.tolist()):
        assert (int(u[i]), int(r[i]), int(k[i])) == clf.clade(members[t]), (t, members[t])
        with_kmers = [m for m in members[t] if cnt["n_kmers"][list(cnt["taxid"]).index(m)]]
        if len(with_kmers) > 1:
            sp = [bool(cnt["sparse"][list(cnt["taxid"]).index(m)]) for m in with_kmers]
            kinds.add("sparse" if all(sp) else "dense")
    return kinds


@pytest.mark.parametrize("case", ["all_sparse", "mixed", "dense_only", "forest", "chunked", "collisions"])
def test_all_clades_in_one_call_equal_the_per_clade_merge(case):
    """kuq_clade_counts_tree (one sort of the sparse tier + one fold of the registers for ALL clades) against
    kuq_clade_counts clade by clade, which the tests above hold against the oracle and the reference's report."""
    rng = np.random.default_rng(91)
    if case == "forest":
        rows = [(1, 1, "root", "no rank"), (10, 1, "a", "genus"), (11, 10, "a1", "species"), (12, 10, "a2", "species"),
                (20, 1, "b", "genus"), (21, 20, "b1", "species"), (500, 500, "other root", "no rank"),
                (501, 500, "c1", "species"), (502, 500, "c2", "species"), (600, 999, "dangling parent", "species")]
        tax = synth.Taxonomy(rows)
        labels = [11, 12, 21, 501, 502, 600, 4242]
        genomes = [rng.integers(0, 4, 1500, dtype=np.uint8) for _ in labels]
        # a shared stretch puts the same k-mers' hashes under several taxa's clades only through the LCA labels;
        # equal codes under different taxa come from the reads below hitting shared prefixes
        ks, ts = [], []
        for gseq, t in zip(genomes, labels):
            km, ok = synth.forward_kmers(gseq, K)
            c = np.unique(synth.canonical(km[ok], K))
            ks.append(c); ts.append(np.full(len(c), t, np.uint32))
        km, first = np.unique(np.concatenate(ks), return_index=True)
        tx = np.concatenate(ts)[first]
        kdb, idx = synth.build_db_images(km, tx, K, 7, 2)
        g = [synth.decode(x).tobytes() for x in genomes]
        seqs = [g[i % len(g)][s:s + 150] for i, s in enumerate(rng.integers(0, 1300, 400).tolist())]
        bases, offs = synth.pack_reads(seqs)
        mode, unit, tax_rows = binding.HLL_PRELOAD, 3000, rows
    else:
        # "collisions": ~10^5 codes per taxon, so that different taxa of a clade hold EQUAL encoded hashes (a few hundred
        # per pair out of 2^25) — the case in which a clade's union is smaller than the sum of its members
        n_gen, glen, n_reads = (4, 150000, 5000) if case == "collisions" else (12, 3000, 1500)
        tax = synth.make_taxonomy(n_gen, 2, 1) if case == "collisions" else synth.make_taxonomy(n_gen, 4, 2)
        genomes = synth.random_genomes(rng, n_gen, glen, shared_frac=0.3 if case != "collisions" else 0.02)
        km, tx = synth.label_kmers(genomes, synth.species_ids(tax), tax, K)
        kdb, idx = synth.build_db_images(km, tx, K, 8, 2)
        bases, offs = synth.sample_reads(rng, genomes, n_reads, 150, 0.001, 0.1, 0.02)
        tax_rows = tax.rows
        mode = {"all_sparse": binding.HLL_PRELOAD, "mixed": binding.HLL_PRELOAD, "dense_only": binding.HLL_DENSE_ONLY,
                "chunked": binding.HLL_CHUNKED, "collisions": binding.HLL_PRELOAD}[case]
        unit = {"all_sparse": 1200, "mixed": 40000, "dense_only": 500000, "chunked": 500000, "collisions": 1200}[case]
    clf = _classifier(hll_mode=mode, work_unit_size=unit)
    clf.stage_db(kdb, idx)
    clf.set_taxonomy(*tax.parent_map())
    clf.classify(bases, offs)
    clf.finish()
    kinds = _clade_tree_equals_per_clade(clf, tax_rows)
    if case in ("all_sparse", "collisions"):
        assert kinds == {"sparse"}
    if case == "dense_only":
        assert kinds == {"dense"}
    # a second batch changes the state: the roll-up follows
    clf.classify(bases[: int(offs[200])], offs[:201])
    clf.finish()
    _clade_tree_equals_per_clade(clf, tax_rows)
    clf.close()
