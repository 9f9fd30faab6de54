"""ctypes binding of libkuq.so (include/kuq.h) — the host-side mirror of the reference's classify driver objects.

`Classifier` plays the role of the globals `KrakenDatabases` / `Parent_map` / `taxon_counts` of
src/classify.cpp:78-113 and of its per-read function `classify_sequence()` (:897-1012): stage a database, set
the taxonomy, push batches of reads, read per-read calls / hit lists and the per-taxon counters back.

There is no CPU path here: if libkuq.so is missing or no sm_100 device is present this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

AMBIG = 0xFFFFFFFF
HLL_PRELOAD, HLL_CHUNKED, HLL_DENSE_ONLY, HLL_EXACT = 0, 1, 2, 3
F_WANT_CODES, F_NO_RUNS, F_NO_COUNTS, F_STATS = 1, 2, 4, 8

u8p, u32p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)


class KuqError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libkuq error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("n_slots", C.c_uint32), ("max_reads_per_batch", C.c_uint32),
                ("max_bases_per_batch", C.c_uint64), ("work_unit_size", C.c_uint64), ("hll_mode", C.c_uint32),
                ("reserved0", C.c_uint32), ("sparse_set_slots", C.c_uint64)]


class Run(C.Structure):
    _fields_ = [("code", C.c_uint32), ("count", C.c_uint32)]


class BatchResult(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("reserved0", C.c_uint32), ("call", u32p), ("n_windows", u32p),
                ("run_start", u32p), ("run_count", u32p), ("runs", C.POINTER(Run)), ("n_runs", C.c_uint64),
                ("codes", u32p), ("n_classified", C.c_uint64), ("kernel_ms", C.c_double)]


class DeviceResult(C.Structure):
    _fields_ = [("d_call", C.c_void_p), ("d_n_windows", C.c_void_p), ("d_codes", C.c_void_p),
                ("d_run_start", C.c_void_p), ("d_run_count", C.c_void_p), ("d_runs", C.c_void_p),
                ("d_n_runs", C.c_void_p)]


LAYOUT_VARIANTS = 22
LAYOUT_SHAPES = ["thread per window", "product k_lookup<MODE_LOOKUP>", "index prefetch, 8 CTAs/SM", "index prefetch, 6 CTAs/SM",
                 "2 windows per thread, 8 CTAs/SM", "2 windows per thread, 6 CTAs/SM", "2 windows per thread, 4 CTAs/SM",
                 "4 windows per thread, 4 CTAs/SM", "product k_lookup<MODE_LOOKUP>, lean", "product k_lookup<MODE_FUSED>",
                 "product k_lookup<MODE_FUSED>, lean", "thread per window, if-chain narrowing",
                 "thread per window, product parameter block"]


class LayoutResult(C.Structure):                     # kuq_layout_result
    _fields_ = [("n_records", C.c_uint64), ("n_positions", C.c_uint64), ("n_windows", C.c_uint64),
                ("sum_probes", C.c_uint64), ("bin_class", C.c_uint64 * 6), ("transcode_ms", C.c_double),
                ("n_variants", C.c_uint32), ("rec_bytes", C.c_uint32 * LAYOUT_VARIANTS),
                ("arity", C.c_uint32 * LAYOUT_VARIANTS), ("window", C.c_uint32 * LAYOUT_VARIANTS),
                ("shape", C.c_uint32 * LAYOUT_VARIANTS),
                ("best_ms", C.c_double * LAYOUT_VARIANTS), ("mean_ms", C.c_double * LAYOUT_VARIANTS),
                ("mismatches", C.c_uint64 * LAYOUT_VARIANTS)]


class StatePtrs(C.Structure):
    _fields_ = [("d_regs", C.c_void_p), ("regs_bytes", C.c_uint64), ("d_n_kmers", C.c_void_p),
                ("d_n_reads", C.c_void_p), ("d_dense_flag", C.c_void_p), ("n_sketch", C.c_uint32),
                ("n_taxa", C.c_uint32)]


_LIB = None


def lib_path():
    return _build.LIB


def load_library(rebuild_if_stale: bool = False):
    """Load libkuq.so; raises if it is not there (no silent fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if rebuild_if_stale:
        _build.build()
    if not os.path.exists(_build.LIB):
        raise FileNotFoundError(f"{_build.LIB} is missing: run `python -m krakenuniq_b200.build` "
                                "(the CUDA extension is the only implementation; there is no CPU path)")
    L = C.CDLL(_build.LIB)
    vp = C.c_void_p
    sigs = {
        "kuq_config_default": (None, [C.POINTER(Config)]),
        "kuq_create": (C.c_int, [C.POINTER(Config), C.POINTER(vp)]),
        "kuq_destroy": (None, [vp]),
        "kuq_strerror": (C.c_char_p, [C.c_int]),
        "kuq_last_error": (C.c_char_p, [vp]),
        "kuq_version": (C.c_char_p, []),
        "kuq_device_count": (C.c_int, []),
        "kuq_stage_db": (C.c_int, [vp, vp, C.c_uint64, vp, C.c_uint64, C.c_uint64, C.c_uint64]),
        "kuq_attach_db_device": (C.c_int, [vp, vp, C.c_uint64, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64,
                                           C.c_uint64]),
        "kuq_db_taxids": (C.c_int, [vp, u32p, u64p, C.c_uint32, u32p]),
        "kuq_set_db_taxid_universe": (C.c_int, [vp, u32p, C.c_uint32]),
        "kuq_mark_zero_hits": (C.c_int, [vp, C.c_int]),
        "kuq_db_sort": (C.c_int, [C.c_int, vp, C.c_uint64, C.c_uint32, C.c_int, vp, vp, C.c_char_p, C.c_uint64]),
        "kuq_set_lcas_batch": (C.c_int, [vp, vp, u64p, C.c_uint32, u32p, C.c_uint32, u64p]),
        "kuq_export_db_values": (C.c_int, [vp, vp, C.c_uint64]),
        "kuq_set_quick_mode": (C.c_int, [vp, C.c_uint32, C.c_int]),
        "kuq_set_taxonomy": (C.c_int, [vp, u32p, u32p, C.c_uint32]),
        "kuq_classify_batch": (C.c_int, [vp, vp, u64p, C.c_uint32, u32p, C.c_uint32, C.POINTER(BatchResult)]),
        "kuq_submit_batch": (C.c_int, [vp, C.c_uint32, vp, u64p, C.c_uint32, u32p, C.c_uint32]),
        "kuq_wait_batch": (C.c_int, [vp, C.c_uint32, C.POINTER(BatchResult)]),
        "kuq_lookup_batch": (C.c_int, [vp, vp, u64p, C.c_uint32, u32p, u32p]),
        "kuq_resolve_batch": (C.c_int, [vp, vp, u64p, C.c_uint32, u32p, u32p, C.c_uint32, C.POINTER(BatchResult)]),
        "kuq_host_alloc": (vp, [C.c_uint64]),
        "kuq_host_free": (None, [vp]),
        "kuq_classify_device": (C.c_int, [vp, C.c_uint32, vp, vp, C.c_uint32, C.c_uint64, vp, C.c_uint32]),
        "kuq_lookup_device": (C.c_int, [vp, C.c_uint32, vp, vp, C.c_uint32, C.c_uint64, vp, C.c_uint32]),
        "kuq_resolve_device": (C.c_int, [vp, C.c_uint32, vp, vp, C.c_uint32, C.c_uint64, vp, vp, C.c_uint32]),
        "kuq_lookup_device_peers": (C.c_int, [vp, C.c_uint32, vp, vp, C.c_uint32, C.c_uint64, C.POINTER(vp), u64p,
                                             C.c_uint32]),
        "kuq_device_alloc": (vp, [vp, C.c_uint64]),
        "kuq_device_free": (None, [vp, vp]),
        "kuq_device_memset": (C.c_int, [vp, C.c_uint32, vp, C.c_int, C.c_uint64]),
        "kuq_ipc_export": (C.c_int, [vp, vp, u8p]),
        "kuq_ipc_open": (C.c_int, [vp, u8p, C.POINTER(vp)]),
        "kuq_ipc_close": (C.c_int, [vp, vp]),
        "kuq_sync_slot": (C.c_int, [vp, C.c_uint32]),
        "kuq_copy_to_device": (C.c_int, [vp, C.c_uint32, vp, vp, C.c_uint64]),
        "kuq_device_free_bytes": (C.c_uint64, [vp]),
        "kuq_collect_device_batch": (C.c_int, [vp, C.c_uint32, C.POINTER(BatchResult)]),
        "kuq_slot_device_result": (C.c_int, [vp, C.c_uint32, C.POINTER(DeviceResult)]),
        "kuq_slot_stream": (vp, [vp, C.c_uint32]),
        "kuq_slot_stats": (C.c_int, [vp, C.c_uint32, u64p, u64p]),
        "kuq_launch_count": (C.c_uint64, [vp]),
        "kuq_last_kernel_ms": (C.c_double, [vp, C.c_uint32]),
        "kuq_last_stage_ms": (C.c_int, [vp, C.c_uint32, C.POINTER(C.c_double)]),
        "kuq_finish": (C.c_int, [vp]),
        "kuq_counts_size": (C.c_int, [vp, u32p]),
        "kuq_read_counts": (C.c_int, [vp, u32p, u64p, u64p, u64p, u8p, C.c_uint32]),
        "kuq_clade_counts": (C.c_int, [vp, u32p, C.c_uint32, u64p, u64p, u64p]),
        "kuq_clade_counts_tree": (C.c_int, [vp, u32p, C.c_uint32, u64p, u64p, u64p]),
        "kuq_get_registers": (C.c_int, [vp, C.c_uint32, u8p]),
        "kuq_state_ptrs_get": (C.c_int, [vp, C.POINTER(StatePtrs)]),
        "kuq_dense_taxids": (C.c_int, [vp, u32p, C.c_uint32, u32p]),
        "kuq_sparse_export": (C.c_int, [vp, vp, C.c_uint64, u64p]),
        "kuq_sparse_import": (C.c_int, [vp, vp, C.c_uint64]),
        "kuq_reset_counts": (C.c_int, [vp]),
        "kuq_sparse_tier_info": (C.c_int, [vp, u64p, u64p, u64p, C.POINTER(C.c_double)]),
        "kuq_set_shard_counting": (C.c_int, [vp, C.c_int]),
        "kuq_set_stats": (C.c_int, [vp, C.c_int]),
        "kuq_merge_into": (C.c_int, [vp, vp]),
        "kuq_enable_peer_access": (C.c_int, [vp, vp]),
        "kuq_stream_open": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64]),
        "kuq_stream_load": (C.c_int, [vp, C.c_uint32, vp, C.c_uint64, vp, C.c_uint64, C.c_uint64]),
        "kuq_stream_use": (C.c_int, [vp, C.c_uint32]),
        "kuq_stream_check": (C.c_int, [vp]),
        "kuq_host_register": (C.c_int, [vp, C.c_uint64]),
        "kuq_host_unregister": (C.c_int, [vp]),
        "kuq_signal_peers": (C.c_int, [vp, C.c_uint32, C.POINTER(vp), C.c_uint32, C.c_uint32, C.c_uint64]),
        "kuq_wait_flags": (C.c_int, [vp, C.c_uint32, vp, C.c_uint32, C.c_uint64, C.c_uint32]),
        "kuq_sparse_export_partitioned": (C.c_int, [vp, C.c_uint32, vp, C.c_uint64, u64p]),
        "kuq_sparse_export_partitioned_alloc": (C.c_int, [vp, C.c_uint32, C.POINTER(vp), u64p]),
        "kuq_sparse_replace": (C.c_int, [vp, vp, C.c_uint64]),
        "kuq_sparse_summary": (C.c_int, [vp, vp, vp]),
        "kuq_set_sparse_summary": (C.c_int, [vp, vp, vp]),
        "kuq_clade_partial": (C.c_int, [vp, u32p, C.c_uint32, u64p, u64p, C.POINTER(C.c_int), u32p]),
        "kuq_ertl_sparse": (C.c_uint64, [u32p, C.c_uint64]),
        "kuq_ertl_dense_hist": (C.c_uint64, [u32p, C.c_uint64]),
        "kuq_scan_device": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp, C.c_uint32, C.c_uint64, vp, vp]),
        "kuq_ertl_dense": (C.c_uint64, [u8p, C.c_uint64]),
        "kuq_layout_experiment": (C.c_int, [vp, C.c_uint32, C.c_uint64, C.c_uint32, C.POINTER(LayoutResult)]),
        "kuq_random_gather_peak": (C.c_int, [C.c_int, C.c_uint64, C.c_uint64, C.c_uint32, C.POINTER(C.c_double),
                                            C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    }
    for name, (res, args) in sigs.items():
        f = getattr(L, name)            # AttributeError here = the library does not export what kuq.h declares
        f.restype, f.argtypes = res, args
    _LIB = L
    return L


def exported_symbols():
    """Names include/kuq.h declares; used by the CPU tests to check the library exports them all."""
    import re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "kuq.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(kuq_[a-z0-9_]+)\s*\(", hdr)))


def _p(a, t):
    return a.ctypes.data_as(t)


class Classifier:
    """One GPU's classification context (kuq_ctx)."""

    def __init__(self, device=0, n_slots=2, max_reads=1 << 20, max_bases=192 << 20, work_unit_size=500000,
                 hll_mode=HLL_PRELOAD, sparse_set_slots=1 << 26):
        self.L = load_library()
        cfg = Config()
        self.L.kuq_config_default(C.byref(cfg))
        cfg.device, cfg.n_slots = device, n_slots
        cfg.max_reads_per_batch, cfg.max_bases_per_batch = max_reads, max_bases
        cfg.work_unit_size, cfg.hll_mode, cfg.sparse_set_slots = work_unit_size, hll_mode, sparse_set_slots
        self.cfg = cfg
        h = C.c_void_p()
        rc = self.L.kuq_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise KuqError(rc, self.L.kuq_strerror(rc).decode())
        self.h = h
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.L.kuq_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise KuqError(rc, f"{self.L.kuq_strerror(rc).decode()}: {self.L.kuq_last_error(self.h).decode()}")

    # ---- database / taxonomy -------------------------------------------------------------------------------
    def stage_db(self, kdb: np.ndarray, idx: np.ndarray, bin_lo=0, bin_hi=0):
        kdb = np.ascontiguousarray(kdb).view(np.uint8) if not isinstance(kdb, np.memmap) else kdb
        idx = np.ascontiguousarray(idx).view(np.uint8) if not isinstance(idx, np.memmap) else idx
        self._ck(self.L.kuq_stage_db(self.h, kdb.ctypes.data, kdb.size, idx.ctypes.data, idx.size, bin_lo, bin_hi))

    def stage_db_files(self, kdb_path, idx_path, bin_lo=0, bin_hi=0):
        self.stage_db(np.memmap(kdb_path, np.uint8, "r"), np.memmap(idx_path, np.uint8, "r"), bin_lo, bin_hi)

    def attach_db_device(self, d_pairs_ptr, key_ct, d_offsets_ptr, k, nt, idx_type=2, bin_lo=0, bin_hi=0):
        self._ck(self.L.kuq_attach_db_device(self.h, d_pairs_ptr, key_ct, d_offsets_ptr, k, nt, idx_type, bin_lo,
                                             bin_hi))

    def db_taxids(self):
        n = C.c_uint32(0)
        self._ck(self.L.kuq_db_taxids(self.h, None, None, 0, C.byref(n)))
        t = np.zeros(n.value, np.uint32)
        c = np.zeros(n.value, np.uint64)
        if n.value:
            self._ck(self.L.kuq_db_taxids(self.h, _p(t, u32p), _p(c, u64p), n.value, C.byref(n)))
        return t, c

    def set_quick_mode(self, min_hits, stop_at_last_hit=True):
        """classify -q -m min_hits (0 = off); stop_at_last_hit=False gives the -x path's rule (kuq_set_quick_mode)"""
        self._ck(self.L.kuq_set_quick_mode(self.h, int(min_hits), 1 if stop_at_last_hit else 0))

    def set_lcas(self, bases: np.ndarray, piece_offsets: np.ndarray, taxids, flags=0) -> int:
        """kuq_set_lcas_batch: fold the pieces' taxids into the staged database's values; returns #k-mers not found.
        flags: 1 = set_lcas -T, 2 = -R"""
        bases = np.ascontiguousarray(bases, np.uint8)
        offs = np.ascontiguousarray(piece_offsets, np.uint64)
        t = np.ascontiguousarray(taxids, np.uint32)
        missing = C.c_uint64(0)
        buf = bases if bases.size else np.zeros(1, np.uint8)
        self._ck(self.L.kuq_set_lcas_batch(self.h, buf.ctypes.data, _p(offs, u64p), len(offs) - 1, _p(t, u32p), flags, C.byref(missing)))
        return missing.value

    def export_db_values(self, kdb: np.ndarray):
        """kuq_export_db_values: write the device record values (taxids) into the host image `kdb` in place"""
        assert kdb.dtype == np.uint8 and kdb.flags["C_CONTIGUOUS"]
        self._ck(self.L.kuq_export_db_values(self.h, kdb.ctypes.data, kdb.size))

    def mark_zero_hits(self, on=True):
        """several databases: lookups report a stored taxon 0 as CODE_FOUND_ZERO (kuq_mark_zero_hits)"""
        self._ck(self.L.kuq_mark_zero_hits(self.h, 1 if on else 0))

    def set_db_taxid_universe(self, taxids):
        t = np.ascontiguousarray(taxids, np.uint32)
        self._ck(self.L.kuq_set_db_taxid_universe(self.h, _p(t, u32p), len(t)))

    def set_taxonomy(self, taxid, parent):
        t = np.ascontiguousarray(taxid, np.uint32)
        p = np.ascontiguousarray(parent, np.uint32)
        self._ck(self.L.kuq_set_taxonomy(self.h, _p(t, u32p), _p(p, u32p), len(t)))

    # ---- host-buffer classification ---------------------------------------------------------------------------
    def _result(self, res: BatchResult, offsets, copy=True):
        n = res.n_reads
        as_np = lambda ptr, cnt, dt: (np.ctypeslib.as_array(ptr, shape=(cnt,)).view(dt) if cnt else np.zeros(0, dt))
        out = dict(call=as_np(res.call, n, np.uint32), n_windows=as_np(res.n_windows, n, np.uint32),
                   n_classified=int(res.n_classified), kernel_ms=res.kernel_ms, n_runs=int(res.n_runs))
        if res.n_runs or res.run_start:
            out["run_start"] = as_np(res.run_start, n, np.uint32)
            out["run_count"] = as_np(res.run_count, n, np.uint32)
            runs = np.ctypeslib.as_array(C.cast(res.runs, u32p), shape=(int(res.n_runs) * 2,)) if res.n_runs else \
                np.zeros(0, np.uint32)
            out["runs"] = runs.reshape(-1, 2)
        if res.codes:
            total = int(offsets[-1] - offsets[0])
            out["codes"] = as_np(res.codes, total, np.uint32)
        if copy:
            out = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in out.items()}
        return out

    def classify(self, bases: np.ndarray, offsets: np.ndarray, unit_id=None, flags=0):
        """kuq_classify_batch: returns dict(call, n_windows, run_start, run_count, runs[, codes])."""
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        n = len(offsets) - 1
        u = np.ascontiguousarray(unit_id, np.uint32) if unit_id is not None else None
        res = BatchResult()
        self._ck(self.L.kuq_classify_batch(self.h, bases.ctypes.data if bases.size else None, _p(offsets, u64p), n,
                                           _p(u, u32p) if u is not None else None, flags, C.byref(res)))
        return self._result(res, offsets)

    def lookup(self, bases: np.ndarray, offsets: np.ndarray):
        """kuq_lookup_batch: per-window dense ids of the staged range (host arrays) + window counts"""
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        n = len(offsets) - 1
        total = int(offsets[-1] - offsets[0])
        codes = np.zeros(max(total, 1), np.uint32)
        nwin = np.zeros(max(n, 1), np.uint32)
        self._ck(self.L.kuq_lookup_batch(self.h, bases.ctypes.data if bases.size else None, _p(offsets, u64p), n,
                                         _p(codes, u32p), _p(nwin, u32p)))
        return codes[:total], nwin[:n]

    def resolve(self, bases: np.ndarray, offsets: np.ndarray, codes: np.ndarray, unit_id=None, flags=0):
        """kuq_resolve_batch: calls / hit lists / counters from merged per-window dense ids"""
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        codes = np.ascontiguousarray(codes, np.uint32)
        n = len(offsets) - 1
        u = np.ascontiguousarray(unit_id, np.uint32) if unit_id is not None else None
        res = BatchResult()
        self._ck(self.L.kuq_resolve_batch(self.h, bases.ctypes.data if bases.size else None, _p(offsets, u64p), n,
                                          _p(codes, u32p), _p(u, u32p) if u is not None else None, flags, C.byref(res)))
        return self._result(res, offsets)

    def submit(self, slot, bases_ptr, offsets: np.ndarray, unit_id=None, flags=0):
        n = len(offsets) - 1
        u = np.ascontiguousarray(unit_id, np.uint32) if unit_id is not None else None
        self._ck(self.L.kuq_submit_batch(self.h, slot, bases_ptr, _p(offsets, u64p), n,
                                         _p(u, u32p) if u is not None else None, flags))

    def wait(self, slot, offsets=None, copy=False):
        res = BatchResult()
        self._ck(self.L.kuq_wait_batch(self.h, slot, C.byref(res)))
        return self._result(res, offsets if offsets is not None else np.zeros(2, np.uint64), copy=copy)

    # ---- device-buffer classification ---------------------------------------------------------------------------
    def classify_device(self, slot, d_bases, d_offsets, n_reads, total_bases, d_unit=None, flags=0):
        self._ck(self.L.kuq_classify_device(self.h, slot, d_bases, d_offsets, n_reads, total_bases, d_unit, flags))

    def lookup_device(self, slot, d_bases, d_offsets, n_reads, total_bases, d_codes_out, only_hits=0):
        self._ck(self.L.kuq_lookup_device(self.h, slot, d_bases, d_offsets, n_reads, total_bases, d_codes_out,
                                          only_hits))

    def resolve_device(self, slot, d_bases, d_offsets, n_reads, total_bases, d_codes_in, d_unit=None, flags=0):
        self._ck(self.L.kuq_resolve_device(self.h, slot, d_bases, d_offsets, n_reads, total_bases, d_codes_in, d_unit,
                                           flags))

    def lookup_device_peers(self, slot, d_bases, d_offsets, n_reads, total_bases, peer_ptrs, base_bounds):
        n = len(peer_ptrs)
        arr = (C.c_void_p * n)(*[C.c_void_p(int(x)) for x in peer_ptrs])
        bounds = np.ascontiguousarray(base_bounds, np.uint64)
        assert len(bounds) == n + 1
        self._ck(self.L.kuq_lookup_device_peers(self.h, slot, d_bases, d_offsets, n_reads, total_bases, arr,
                                                _p(bounds, u64p), n))

    def device_alloc(self, nbytes):
        p = self.L.kuq_device_alloc(self.h, nbytes)
        if not p:
            raise KuqError(-8, "device allocation failed")
        return p

    def device_free(self, p):
        self.L.kuq_device_free(self.h, p)

    def device_memset(self, slot, p, value, nbytes):
        self._ck(self.L.kuq_device_memset(self.h, slot, p, value, nbytes))

    def ipc_export(self, p) -> bytes:
        h = np.zeros(64, np.uint8)
        self._ck(self.L.kuq_ipc_export(self.h, p, _p(h, u8p)))
        return h.tobytes()

    def ipc_open(self, handle: bytes):
        h = np.frombuffer(handle, np.uint8).copy()
        out = C.c_void_p()
        self._ck(self.L.kuq_ipc_open(self.h, _p(h, u8p), C.byref(out)))
        return out.value

    def ipc_close(self, p):
        self._ck(self.L.kuq_ipc_close(self.h, p))

    def sync(self, slot):
        self._ck(self.L.kuq_sync_slot(self.h, slot))

    def device_result(self, slot):
        r = DeviceResult()
        self._ck(self.L.kuq_slot_device_result(self.h, slot, C.byref(r)))
        return r

    def slot_stats(self, slot):
        a, b = C.c_uint64(0), C.c_uint64(0)
        self._ck(self.L.kuq_slot_stats(self.h, slot, C.byref(a), C.byref(b)))
        return a.value, b.value

    def slot_stream(self, slot):
        return self.L.kuq_slot_stream(self.h, slot)

    def launch_count(self):
        return int(self.L.kuq_launch_count(self.h))

    def last_kernel_ms(self, slot):
        return float(self.L.kuq_last_kernel_ms(self.h, slot))

    def last_stage_ms(self, slot):
        a = (C.c_double * 3)()
        self._ck(self.L.kuq_last_stage_ms(self.h, slot, a))
        return [a[0], a[1], a[2]]

    # ---- results ---------------------------------------------------------------------------------------------
    def finish(self):
        self._ck(self.L.kuq_finish(self.h))

    def counts(self):
        n = C.c_uint32(0)
        self._ck(self.L.kuq_counts_size(self.h, C.byref(n)))
        m = n.value
        taxid, nr, nk, un, sp = (np.zeros(m, np.uint32), np.zeros(m, np.uint64), np.zeros(m, np.uint64),
                                 np.zeros(m, np.uint64), np.zeros(m, np.uint8))
        if m:
            self._ck(self.L.kuq_read_counts(self.h, _p(taxid, u32p), _p(nr, u64p), _p(nk, u64p), _p(un, u64p),
                                            _p(sp, u8p), m))
        return dict(taxid=taxid, n_reads=nr, n_kmers=nk, unique=un, sparse=sp)

    def clade(self, taxids):
        t = np.ascontiguousarray(taxids, np.uint32)
        r, k, u = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        self._ck(self.L.kuq_clade_counts(self.h, _p(t, u32p), len(t), C.byref(r), C.byref(k), C.byref(u)))
        return u.value, r.value, k.value

    def registers(self, taxid):
        r = np.zeros(4096, np.uint8)
        self._ck(self.L.kuq_get_registers(self.h, int(taxid), _p(r, u8p)))
        return r

    def state_ptrs(self):
        s = StatePtrs()
        self._ck(self.L.kuq_state_ptrs_get(self.h, C.byref(s)))
        return s

    def dense_taxids(self):
        n = C.c_uint32(0)
        self._ck(self.L.kuq_dense_taxids(self.h, None, 0, C.byref(n)))
        t = np.zeros(n.value, np.uint32)
        self._ck(self.L.kuq_dense_taxids(self.h, _p(t, u32p), n.value, C.byref(n)))
        return t

    def sparse_export(self, d_out=None, cap=0):
        n = C.c_uint64(0)
        self._ck(self.L.kuq_sparse_export(self.h, d_out, cap, C.byref(n)))
        return n.value

    def sparse_import(self, d_keys, n):
        self._ck(self.L.kuq_sparse_import(self.h, d_keys, n))

    def stream_open(self, k, nt, idx_type, max_records, max_bins):
        self._ck(self.L.kuq_stream_open(self.h, k, nt, idx_type, max_records, max_bins))

    def stream_load(self, buf, host_records_ptr, n_records, host_offsets_ptr, bin_lo, bin_hi):
        self._ck(self.L.kuq_stream_load(self.h, buf, host_records_ptr, n_records, host_offsets_ptr, bin_lo, bin_hi))

    def stream_use(self, buf):
        self._ck(self.L.kuq_stream_use(self.h, buf))

    def stream_check(self):
        self._ck(self.L.kuq_stream_check(self.h))

    def merge_from(self, other):
        """kuq_merge_into(self, other): fold another context's per-taxon state (another GPU of this process) into this one"""
        self._ck(self.L.kuq_merge_into(self.h, other.h))

    def set_stats(self, on=True):
        self._ck(self.L.kuq_set_stats(self.h, 1 if on else 0))

    def set_shard_counting(self, on=True):
        self._ck(self.L.kuq_set_shard_counting(self.h, 1 if on else 0))

    def signal_peers(self, slot, flag_ptrs, my_index, value):
        arr = (C.c_void_p * len(flag_ptrs))(*flag_ptrs)
        self._ck(self.L.kuq_signal_peers(self.h, slot, arr, len(flag_ptrs), my_index, value))

    def wait_flags(self, slot, d_flags, n, value, timeout_ms=0):
        self._ck(self.L.kuq_wait_flags(self.h, slot, d_flags, n, value, timeout_ms))

    def sparse_export_partitioned(self, n_parts, d_out=None, cap=0):
        counts = np.zeros(n_parts, np.uint64)
        self._ck(self.L.kuq_sparse_export_partitioned(self.h, n_parts, d_out, cap, _p(counts, u64p)))
        return counts

    def sparse_export_partitioned_alloc(self, n_parts):
        """→ (device pointer of the keys grouped by part — release with device_free —, counts per part)"""
        counts = np.zeros(n_parts, np.uint64)
        ptr = C.c_void_p()
        self._ck(self.L.kuq_sparse_export_partitioned_alloc(self.h, n_parts, C.byref(ptr), _p(counts, u64p)))
        return ptr.value, counts

    def sparse_replace(self, d_keys, n):
        self._ck(self.L.kuq_sparse_replace(self.h, d_keys, n))

    def sparse_summary(self, d_hist, d_distinct):
        self._ck(self.L.kuq_sparse_summary(self.h, d_hist, d_distinct))

    def set_sparse_summary(self, d_hist, d_distinct):
        self._ck(self.L.kuq_set_sparse_summary(self.h, d_hist, d_distinct))

    def clade_counts_tree(self, taxids):
        """kuq_clade_counts_tree: (reads, kmers, unique) arrays for clade(taxid) = the taxon and all its descendants"""
        t = np.ascontiguousarray(taxids, np.uint32)
        r, k, u = (np.zeros(len(t), np.uint64) for _ in range(3))
        self._ck(self.L.kuq_clade_counts_tree(self.h, _p(t, u32p), len(t), _p(r, u64p), _p(k, u64p), _p(u, u64p)))
        return r, k, u

    def clade_partial(self, taxids):
        """(reads, kmers, is_dense, hist64) of the listed taxa's merged sketch on THIS GPU (see kuq.h)"""
        t = np.ascontiguousarray(taxids, np.uint32)
        r, k, dflag = C.c_uint64(0), C.c_uint64(0), C.c_int(0)
        hist = np.zeros(64, np.uint32)
        self._ck(self.L.kuq_clade_partial(self.h, _p(t, u32p), len(t), C.byref(r), C.byref(k), C.byref(dflag), _p(hist, u32p)))
        return r.value, k.value, bool(dflag.value), hist

    def scan_device(self, slot, k, nt, idx_type, d_bases, d_offsets, n_reads, total_bases, d_canon, d_bins):
        self._ck(self.L.kuq_scan_device(self.h, slot, k, nt, idx_type, d_bases, d_offsets, n_reads, total_bases, d_canon, d_bins))

    def sparse_tier_info(self):
        a, b, c, d = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_double()
        self._ck(self.L.kuq_sparse_tier_info(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return {"slots": a.value, "keys": b.value, "times_grown": c.value, "last_harvest_ms": d.value}

    def layout_experiment(self, slot, n_positions, reps=5):
        """kuq_layout_experiment: record layout x search shape of the bin search, timed on the windows of the slot's
        last batch and checked against the ids the product wrote for it"""
        r = LayoutResult()
        self._ck(self.L.kuq_layout_experiment(self.h, slot, n_positions, reps, C.byref(r)))
        out = {"n_records": r.n_records, "n_positions": r.n_positions, "n_windows": r.n_windows,
               "probes_per_window": r.sum_probes / max(r.n_windows, 1), "transcode_ms": r.transcode_ms,
               "windows_by_bin_size": dict(zip(["<=8", "<=16", "<=64", "<=256", "<=1024", ">1024"], list(r.bin_class))),
               "variants": []}
        for v in range(r.n_variants):
            out["variants"].append({"record_bytes": r.rec_bytes[v], "arity": r.arity[v], "scan_window": r.window[v],
                                    "shape": LAYOUT_SHAPES[r.shape[v]],
                                    "best_ms": r.best_ms[v], "mean_ms": r.mean_ms[v], "mismatches": r.mismatches[v]})
        return out

    def reset_counts(self):
        self._ck(self.L.kuq_reset_counts(self.h))


def ertl_dense(regs: np.ndarray, n_observed: int) -> int:
    L = load_library()
    regs = np.ascontiguousarray(regs, np.uint8)
    return int(L.kuq_ertl_dense(_p(regs, u8p), n_observed))


def ertl_sparse(hist64: np.ndarray, n_observed: int) -> int:
    L = load_library()
    h = np.ascontiguousarray(hist64, np.uint32)
    return int(L.kuq_ertl_sparse(_p(h, u32p), n_observed))


def ertl_dense_hist(hist64: np.ndarray, n_observed: int) -> int:
    L = load_library()
    h = np.ascontiguousarray(hist64, np.uint32)
    return int(L.kuq_ertl_dense_hist(_p(h, u32p), n_observed))


def random_gather_peak(device=0, buffer_bytes=16 << 30, n_sectors=1 << 30, bytes_per_access=32):
    """kuq_random_gather_peak: (G sectors/s, GB/s of 32-byte sectors, kernel ms) — SURVEY.md §8(d)'s second roofline"""
    L = load_library()
    gs, gb, ms = C.c_double(), C.c_double(), C.c_double()
    rc = L.kuq_random_gather_peak(device, buffer_bytes, n_sectors, bytes_per_access, C.byref(gs), C.byref(gb), C.byref(ms))
    if rc != 0:
        raise KuqError(rc, L.kuq_strerror(rc).decode())
    return gs.value, gb.value, ms.value


def decode_runs(res, i):
    """hit list of read i as [(code, count), ...] from a classify() result"""
    s, c = int(res["run_start"][i]), int(res["run_count"][i])
    return [(int(a), int(b)) for a, b in res["runs"][s:s + c]]


def db_sort(jdb: np.ndarray, nt: int, zero_vals=False, device=0):
    """kuq_db_sort: unsorted Jellyfish-style image → (database.kdb image, KRAKIX2 index image), on the GPU"""
    L = load_library()
    jdb = np.ascontiguousarray(jdb, np.uint8)
    key_bits = int(np.frombuffer(jdb[8:16].tobytes(), np.uint64)[0])
    key_ct = int(np.frombuffer(jdb[48:56].tobytes(), np.uint64)[0])
    size = 72 + 2 * (4 + 8 * key_bits) + key_ct * ((key_bits + 7) // 8 + 4)
    kdb = np.zeros(size, np.uint8)
    idx = np.zeros(8 + 8 * (4 ** nt + 1), np.uint8)
    err = C.create_string_buffer(512)
    rc = L.kuq_db_sort(device, jdb.ctypes.data, jdb.size, nt, 1 if zero_vals else 0, kdb.ctypes.data, idx.ctypes.data, err, 512)
    if rc != 0:
        raise KuqError(rc, err.value.decode() or L.kuq_strerror(rc).decode())
    return kdb, idx
