"""CPU-side checks of the C-ABI library: it builds, loads, exports every symbol include/kuq.h declares, refuses to
run without a GPU (no CPU fallback), and its host-only estimator equals the oracle's."""
import ctypes as C
import os

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


def test_scalar_minimizer_matches_oracle(tmp_path, oracle):
    """kuq_minimizer.cuh (the per-key bin_key the GPU db_sort uses) compiled for the host == the oracle's bin_key"""
    import ctypes
    import subprocess
    src = tmp_path / "m.cpp"
    hdr = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "krakenuniq_b200", "csrc", "kuq_minimizer.cuh")
    src.write_text(f'#include "{hdr}"\nextern "C" void bins(const uint64_t *k, uint64_t n, uint32_t kk, uint32_t nt, '
                   'uint32_t *out) { for (uint64_t i = 0; i < n; i++) out[i] = kuq::bin_key_of(k[i], kk, nt); }\n')
    so = tmp_path / "m.so"
    subprocess.run(["/usr/bin/g++", "-O2", "-shared", "-fPIC", "-x", "c++", str(src), "-o", str(so)], check=True)
    lib = ctypes.CDLL(str(so))
    rng = np.random.default_rng(0)
    for k, nt in [(31, 15), (31, 7), (31, 1), (25, 12), (16, 8)]:
        keys = rng.integers(0, 1 << (2 * k), 4000, dtype=np.uint64)
        keys[:4] = [0, (1 << (2 * k)) - 1, 0x5555555555555555 & ((1 << (2 * k)) - 1), 1]
        out = np.zeros(len(keys), np.uint32)
        lib.bins(keys.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(len(keys)), k, nt, out.ctypes.data_as(ctypes.c_void_p))
        want = np.array([oracle.bin_key(int(x), k, nt, 2) for x in keys.tolist()], np.uint64)
        assert np.array_equal(out.astype(np.uint64), want), (k, nt)


def test_executables_fail_loudly_without_a_gpu(tmp_path):
    """No CPU path anywhere: on a box without an sm_100 device every drop-in executable stops with EX_UNAVAILABLE"""
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from krakenuniq_b200 import synth
    from tests import util
    build.build_classify()
    db_sort, set_lcas = build.build_dbtools()
    G = util.GOLDEN
    db = ["-d", os.path.join(G, "database.kdb"), "-i", os.path.join(G, "database.idx")]
    exe = os.path.dirname(build.CLASSIFY)
    keys = np.unique(np.random.default_rng(0).integers(0, 1 << 62, 500, dtype=np.uint64))
    synth.unsorted_jdb_image(keys, 31).tofile(tmp_path / "t.jdb")
    (tmp_path / "map").write_text("r0\t100\n")
    cmds = [[os.path.join(exe, "classify")] + db + ["-a", os.path.join(G, "taxDB"), "-M", os.path.join(G, "reads.fa")],
            [os.path.join(exe, "classifyExact")] + db + ["-a", os.path.join(G, "taxDB"), "-M", os.path.join(G, "reads.fa")],
            [db_sort, "-n", "5", "-d", str(tmp_path / "t.jdb"), "-o", str(tmp_path / "t.kdb"), "-i", str(tmp_path / "t.idx")],
            [set_lcas] + db + ["-b", os.path.join(G, "taxDB"), "-m", str(tmp_path / "map"), "-F", os.path.join(G, "reads.fa"), "-p"]]
    for cmd in cmds:
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 69, (cmd[0], r.returncode, r.stderr[-300:])
        assert "no CPU path" in r.stderr
        assert r.stdout == ""
    assert not os.path.exists(tmp_path / "t.kdb")
    # argument errors come first and are the reference's: more threads than processors (classify.cpp:1087-1088)
    r = subprocess.run(cmds[0][:1] + ["-t", "1000000"] + cmds[0][1:], capture_output=True, text=True)
    assert r.returncode == 64 and "thread count exceeds number of processors" in r.stderr
