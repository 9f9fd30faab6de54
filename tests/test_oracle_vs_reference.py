"""Pin the C restatement (oracle/kuq_oracle.c) against the UNMODIFIED reference classes (oracle/_ref/libkuref.so),
function by function, plus the known answers SURVEY.md §8(c) extracted from the compiled reference."""
import numpy as np
import pytest

from krakenuniq_b200 import synth
from oracle.oracle_py import RefHLL

K = 31


def _mini_db(rng, nt=8, idx_type=2, n_genomes=4, glen=1500):
    tax = synth.make_taxonomy(n_genomes)
    sp = synth.species_ids(tax)
    genomes = synth.random_genomes(rng, n_genomes, glen)
    km, tx = synth.label_kmers(genomes, sp, tax, K)
    kdb, idx = synth.build_db_images(km, tx, K, nt, idx_type)
    return tax, genomes, km, tx, kdb, idx


def test_fmix_kats(oracle, refshim):
    kats = {0: 0xb456bcfc34c2cb2c, 1: 0x3abf2a20650683e7, 2: 0x0b5181c509f8d8ce,
            0x0123456789abcdef: 0x6573ed81fcdd91d5, (1 << 62) - 1: 0x6554a70955a989a0}
    for x, h in kats.items():
        assert oracle.fmix64(x) == h
        assert refshim.L.kuref_murmur_fmix(x) == h
    rng = np.random.default_rng(1)
    for x in rng.integers(0, 1 << 63, 2000, dtype=np.uint64).tolist():
        assert oracle.fmix64(x) == refshim.L.kuref_murmur_fmix(x)


def test_kmer_kat_from_survey(oracle):
    # read ACGTACGTTAGCTAGCTAGGATCGATCGATTAGC..., window 0
    seq = b"ACGTACGTTAGCTAGCTAGGATCGATCGATTAGC"
    kmers, amb = oracle.scan(seq, K)
    assert kmers[0] == 0x06c6f272728d8d8f and amb[0] == 0
    canon = oracle.canonical(int(kmers[0]), K)
    assert canon == 0x0363635c9c9c1b1b
    assert oracle.bin_key(canon, K, 15, 2) == 3375861
    assert oracle.bin_key(canon, K, 8, 2) == 1713


@pytest.mark.parametrize("idx_type,nt", [(2, 8), (1, 6), (2, 11)])
def test_bits_and_lookup_match_reference(oracle, refshim, idx_type, nt):
    rng = np.random.default_rng(10 + nt)
    tax, genomes, km, tx, kdb, idx = _mini_db(rng, nt, idx_type)
    odb = oracle.open_db(kdb, idx)
    rdb = refshim.open_db(kdb, idx)
    assert odb.k == refshim.L.kuref_db_k(rdb.h) == K
    assert odb.nt == refshim.L.kuref_db_index_nt(rdb.h) == nt
    assert odb.idx_type == refshim.L.kuref_db_index_type(rdb.h) == idx_type
    probe = np.concatenate([km[rng.integers(0, len(km), 1500)],
                            synth.canonical(rng.integers(0, 1 << 62, 1500, dtype=np.uint64), K)])
    for x in probe.tolist():
        assert oracle.canonical(x, K) == rdb.canonical(x)
        assert oracle.revcomp(x, K) == rdb.revcomp(x, K)
        assert oracle.revcomp(x & ((1 << 30) - 1), 15) == rdb.revcomp(x & ((1 << 30) - 1), 15)
        assert oracle.bin_key(x, K, nt, idx_type) == rdb.bin_key(x)
        assert oracle.bin_key(x, K, nt, 2) == rdb.bin_key_nt(x, nt)
        assert odb.query(x) == rdb.query(x)
    # numpy generator agrees too (it builds the fixtures)
    assert np.array_equal(synth.bin_key(probe, K, nt, idx_type),
                          np.array([rdb.bin_key(x) for x in probe.tolist()], np.uint64))
    # every stored key is found with its taxon; stateful (cached-range) lookups agree with stateless ones
    sel = rng.integers(0, len(km), 500)
    for x, t in zip(km[sel].tolist(), tx[sel].tolist()):
        assert odb.query(x) == (True, t)


def test_stateful_query_equals_stateless(oracle, refshim):
    rng = np.random.default_rng(3)
    tax, genomes, km, tx, kdb, idx = _mini_db(rng, 8, 2)
    odb = oracle.open_db(kdb, idx)
    rdb = refshim.open_db(kdb, idx)
    reads, offs = synth.sample_reads(rng, genomes, 200, 150, 0.02, 0.0, 0.3)
    for i in range(200):
        seq = reads[int(offs[i]):int(offs[i + 1])].tobytes()
        kmers, amb = oracle.scan(seq, K)
        canon = synth.canonical(kmers[amb == 0], K)
        got = rdb.query_stateful(canon)
        want = np.array([odb.query(x)[1] for x in canon.tolist()], np.uint32)
        assert np.array_equal(got, want)


def test_scanner_matches_reference(oracle, refshim):
    rng = np.random.default_rng(5)
    cases = [b"", b"ACGT", b"A" * 30, b"A" * 31, b"ACGTN" * 20, b"acgtacgtacgtacgtacgtacgtacgtacgtacgt",
             b"ACGT" * 10 + b"\r", b"ACGTAC" * 10 + b"\r", b"ACGTACGTAC" * 6 + b"\r" + b"GATTACA" * 8 + b"\r",
             b"N" * 40, b"ACGT" * 8 + b"R" + b"TGCA" * 9, b"ACGU" * 12, b"ACGT" * 8 + b"\n" + b"TTGA" * 9]
    for _ in range(200):
        n = int(rng.integers(0, 200))
        alphabet = np.frombuffer(b"ACGTACGTACGTacgtNnRY", np.uint8)
        cases.append(alphabet[rng.integers(0, len(alphabet), n)].tobytes())
    for seq in cases:
        k1, a1 = oracle.scan(seq, K)
        k2, a2 = refshim.scan(seq, K)
        assert len(k1) == len(k2), seq
        assert np.array_equal(a1, a2), seq
        assert np.array_equal(k1[a1 == 0], k2[a2 == 0]), seq


def _random_forest(rng, n, with_one=True):
    ids = np.unique(rng.integers(2, 5000, n)).astype(np.uint32)
    parent = np.zeros(len(ids), np.uint32)
    for i in range(len(ids)):
        r = rng.random()
        if i == 0 or r < 0.05:
            parent[i] = 1 if with_one else 0
        elif r < 0.08:
            parent[i] = 0                      # a second root (parent NULL → 0, taxdb.hpp:389-391)
        elif r < 0.10:
            parent[i] = 77777                  # parent id that has no entry of its own
        else:
            parent[i] = ids[int(rng.integers(0, i))]
    if with_one:
        ids = np.concatenate([[1], ids]).astype(np.uint32)
        parent = np.concatenate([[0], parent]).astype(np.uint32)
    return ids, parent


@pytest.mark.parametrize("with_one", [True, False])
def test_lca_and_resolve_tree_match_reference(oracle, refshim, with_one):
    rng = np.random.default_rng(7 + with_one)
    ids, parent = _random_forest(rng, 300, with_one)
    opm = oracle.parent_map(ids, parent)
    rpm = refshim.parent_map(ids, parent)
    pool = np.concatenate([ids, [0, 1, 99999]]).astype(np.uint32)
    for _ in range(3000):
        a, b = (int(x) for x in pool[rng.integers(0, len(pool), 2)])
        assert opm.lca(a, b) == rpm.lca(a, b), (a, b)
    for _ in range(1500):
        n = int(rng.integers(0, 7))
        nz = pool[pool != 0]
        taxa = np.unique(nz[rng.integers(0, len(nz), n)]) if n else np.zeros(0, np.uint32)
        hits = {int(t): int(rng.integers(1, 4)) for t in taxa}
        assert opm.resolve_tree(hits) == rpm.resolve_tree(hits), hits


def test_hll_kats_and_state_machine(oracle, refshim):
    def seq(n, mul=0x9E3779B97F4A7C15, off=0):
        return ((np.arange(off, off + n, dtype=np.uint64)) * np.uint64(mul))
    # SURVEY §8(c) known answers of the compiled reference
    for n, want in [(1000, 1000), (1024, 1024), (1025, 1017), (2000, 2000), (10000, 9965), (100000, 100000)]:
        o = oracle.hll(); o.insert(seq(n))
        r = RefHLL(refshim); r.insert(seq(n))
        assert r.cardinality() == want
        assert o.cardinality() == want
        assert o.n_observed() == r.n_observed() == n
        assert o.is_sparse() == (n <= 1024)
    # forced dense p=12 via the registers-only estimator
    for n, want in [(1000, 994), (1024, 1015)]:
        r = RefHLL(refshim, dense_p=12); r.insert(seq(n))
        assert r.cardinality() == want
        o = oracle.hll(); o.insert(seq(n))
        assert oracle.ertl_dense(o.registers(), n) == want
    # ten sparse sketches of 1000 merged stay sparse and near exact
    og, rg = oracle.hll(), RefHLL(refshim)
    for i in range(10):
        o = oracle.hll(); o.insert(seq(1000, off=1000 * i)); og.merge(o)
        r = RefHLL(refshim); r.insert(seq(1000, off=1000 * i)); rg.merge(r, move=bool(i & 1))
    assert og.is_sparse() and og.cardinality() == rg.cardinality() == 10000


def test_hll_random_merge_sequences(oracle, refshim):
    rng = np.random.default_rng(11)
    for trial in range(12):
        og, rg = oracle.hll(), RefHLL(refshim)
        for part in range(int(rng.integers(1, 8))):
            n = int(rng.choice([0, 3, 200, 1023, 1024, 1025, 1500, 5000]))
            # duplicates on purpose: draw from a pool about as large as n
            items = rng.integers(0, max(n, 1) * int(rng.choice([1, 4, 1000])), n, dtype=np.uint64)
            o = oracle.hll(); o.insert(items)
            r = RefHLL(refshim); r.insert(items)
            assert o.cardinality() == r.cardinality()
            og.merge(o)
            rg.merge(r, move=bool(part & 1))
            assert og.cardinality() == rg.cardinality(), (trial, part)
            assert og.n_observed() == rg.n_observed()
