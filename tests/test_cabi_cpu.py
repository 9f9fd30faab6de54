"""CPU-side checks of the C-ABI library: it builds, loads, exports every symbol include/kuq.h declares, refuses to
run without a GPU (no CPU fallback), and its host-only estimator equals the oracle's."""
import ctypes as C

import numpy as np
import pytest

from krakenuniq_b200 import binding, build


@pytest.fixture(scope="module")
def lib():
    build.build()
    return binding.load_library()


def test_library_exports_every_declared_symbol(lib):
    names = binding.exported_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"libkuq.so does not export {n}"
    assert b"sm_100a" in lib.kuq_version()


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = lib.kuq_create(None, C.byref(h))
    assert rc == -3 and not h.value            # KUQ_E_NO_DEVICE
    assert b"no CPU path" in lib.kuq_strerror(rc)


def test_host_estimator_matches_oracle(lib, oracle):
    rng = np.random.default_rng(0)
    for n in [0, 1, 10, 500, 5000, 100000, 3000000]:
        h = oracle.hll()
        items = rng.integers(0, 1 << 62, n, dtype=np.uint64)
        h.insert(items)
        regs = h.registers()
        assert binding.ertl_dense(regs, n) == oracle.ertl_dense(regs, n)
    regs = np.full(4096, 53, np.uint8)
    assert binding.ertl_dense(regs, 1 << 60) == oracle.ertl_dense(regs, 1 << 60)
