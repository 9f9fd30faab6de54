"""How far the unique-k-mer estimates (HyperLogLog++, p = 12, Ertl estimator — what `classify` reports) are from the
exact distinct counts (`classifyExact`), on the oracle: the reference documents ~1.6 % standard error for p = 12
(1.04 / sqrt(4096)); the sparse tier (p' = 25) is near exact.  The GPU path reproduces the oracle's estimates bit for
bit (tests/test_gpu_parity.py), so this is also its error."""
import numpy as np

from krakenuniq_b200 import synth


def test_estimates_within_hll_error_of_exact_counts(oracle):
    rng = np.random.default_rng(4)
    tax = synth.make_taxonomy(6)
    sp = synth.species_ids(tax)
    genomes = synth.random_genomes(rng, 6, 60000, shared_frac=0.1)
    km, tx = synth.label_kmers(genomes, sp, tax, 31)
    kdb, idx = synth.build_db_images(km, tx, 31, 8, 2)
    bases, offs = synth.sample_reads(rng, genomes, 6000, 150, 0.01, 0.0, 0.1)
    db = oracle.open_db(kdb, idx)
    pm = oracle.parent_map(*tax.parent_map())
    est = {}
    for unit, mode in ((500000, 0), (20000, 0), (500000, 1)):        # dense after conversion / per-unit rule / -x rule
        run = oracle.run(db, pm, unit, mode)
        run.classify(bases, offs, want_codes=False)
        run.finish()
        est[(unit, mode)] = run.counts()
    run = oracle.run(db, pm, 500000, 0)
    run.set_exact()
    run.classify(bases, offs, want_codes=False)
    run.finish()
    exact = run.counts()
    worst = 0.0
    for key, c in est.items():
        assert np.array_equal(c["taxid"], exact["taxid"]) and np.array_equal(c["n_kmers"], exact["n_kmers"])
        for t, e, x, sparse in zip(c["taxid"], c["unique"], exact["unique"], c["sparse"]):
            rel = abs(int(e) - int(x)) / max(int(x), 1)
            worst = max(worst, rel)
            # 4 sigma of the dense estimator; the sparse tier loses only to 25-bit index collisions
            assert rel < (0.005 if sparse else 0.065), (key, int(t), int(e), int(x), bool(sparse))
    assert worst > 0          # the estimates are estimates: something differs

    # few reads: every sketch stays in the sparse tier, whose estimate is exact up to 25-bit index collisions
    bases, offs = synth.sample_reads(rng, genomes, 30, 150, 0.01, 0.0, 0.1)
    run = oracle.run(db, pm, 500000, 0)
    run.classify(bases, offs, want_codes=False)
    run.finish()
    c = run.counts()
    run = oracle.run(db, pm, 500000, 0)
    run.set_exact()
    run.classify(bases, offs, want_codes=False)
    run.finish()
    x = run.counts()
    sp_rows = c["sparse"].astype(bool)
    assert sp_rows.sum() >= 5
    assert np.abs(c["unique"].astype(np.int64) - x["unique"].astype(np.int64))[sp_rows].max() <= 1
