"""kuq_layout_experiment (measurement aid): every record layout / search shape it times must find exactly what the
product's k_lookup found for the same windows."""
import pytest

torch = pytest.importorskip("torch")

from krakenuniq_b200 import binding, synth_gpu  # noqa: E402

pytestmark = pytest.mark.gpu


def test_layout_variants_agree_with_the_product_lookup():
    dev = "cuda:0"
    B, L = 200_000, 150
    db = synth_gpu.GpuDatabase(200_000_000, n_genomes=50, k=31, nt=15, seed=4, device=dev)
    pool, _ = db.sample_reads(B, L, seed=5)
    torch.cuda.empty_cache()
    clf = binding.Classifier(device=0, n_slots=1, max_reads=B, max_bases=B * L + 4096, hll_mode=2)
    try:
        clf.attach_db_device(db.records.data_ptr(), db.key_ct, db.offsets.data_ptr(), 31, 15, 2)
        clf.set_taxonomy(*db.parent_map())
        d_off = torch.arange(B + 2, dtype=torch.int64, device=dev) * L
        clf.classify_device(0, pool.data_ptr(), d_off.data_ptr(), B, B * L, None, flags=binding.F_NO_COUNTS)
        clf.sync(0)
        res = clf.layout_experiment(0, B * L, reps=1)
        assert res["n_records"] == db.key_ct and res["n_windows"] > 0.9 * B * (L - 30)
        assert len(res["variants"]) == binding.LAYOUT_VARIANTS
        for v in res["variants"]:
            assert v["mismatches"] == 0, v
            assert v["best_ms"] > 0
        with pytest.raises(binding.KuqError):
            clf.layout_experiment(0, B * L + 1, reps=1)         # more positions than the slot's last batch had
    finally:
        clf.close()
