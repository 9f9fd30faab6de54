"""k-mer lengths other than the 31 of every published KrakenUniq database: the kernels' generic (non-specialised) scan and
lookup path, k = 29 and 30 (12-byte records), against the oracle — which tests/test_oracle_live_reference.py holds
against the unmodified reference for the same k."""
import numpy as np
import pytest

from krakenuniq_b200 import binding, synth
from tests.test_gpu_parity import _check_against_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k,nt,idx_type,mode,unit", [(30, 9, 2, binding.HLL_PRELOAD, 20000), (29, 7, 1, binding.HLL_DENSE_ONLY, 500000),
                                                      (30, 15, 2, binding.HLL_CHUNKED, 500000)])
def test_other_kmer_lengths_match_oracle(oracle, k, nt, idx_type, mode, unit):
    rng = np.random.default_rng(1000 + k + nt)
    tax = synth.make_taxonomy(6)
    genomes = synth.random_genomes(rng, 6, 2500, shared_frac=0.25)
    km, tx = synth.label_kmers(genomes, synth.species_ids(tax), tax, k)
    kdb, idx = synth.build_db_images(km, tx, k, nt, idx_type)
    bases, offs = synth.sample_reads(rng, genomes, 700, 150, 0.01, 0.2, 0.2)
    g0 = synth.decode(genomes[0]).tobytes()
    extra = [b"", g0[:k - 1], g0[:k], g0[:k + 1], g0[7:7 + k] + b"N" + g0[300:300 + k], g0[50:50 + 2 * k].lower()]
    seqs = [bases[int(offs[i]):int(offs[i + 1])].tobytes() for i in range(len(offs) - 1)] + extra
    bases, offs = synth.pack_reads(seqs)
    _check_against_oracle(oracle, kdb, idx, tax, bases, offs, hll_mode=mode, unit=unit,
                          oracle_mode=1 if mode == binding.HLL_CHUNKED else 0)
