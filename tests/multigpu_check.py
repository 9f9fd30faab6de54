#!/usr/bin/env python
"""Multi-GPU parity check (run with torchrun, one rank per GPU):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py
Replicas: every rank stages the whole database, classifies its contiguous share of the reads; after the NCCL merge
(krakenuniq_b200.dist.merge_classifier_state) every rank must hold exactly the per-taxon state of a single-GPU run
over all reads — counters, HLL registers, sparse/dense tier and estimates — which is also checked against the oracle.
Shards: each rank stages one minimizer range, all ranks look the whole batch up (`only_hits`), rank r's hits are
merged by an NCCL MAX all-reduce (a key lives in one range), and the owner resolves: same final state again."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from krakenuniq_b200 import binding, synth  # noqa: E402
from krakenuniq_b200 import dist as kdist  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    rng = np.random.default_rng(17)
    # 4 abundant species (their per-unit sketches convert to dense) + 40 rare ones (stay sparse everywhere)
    tax = synth.make_taxonomy(44)
    genomes = synth.random_genomes(rng, 44, 2000, 0.2)
    km, tx = synth.label_kmers(genomes, synth.species_ids(tax), tax, 31)
    kdb, idx = synth.build_db_images(km, tx, 31, 9, 2)
    b1, o1 = synth.sample_reads(rng, genomes[:4], 5000, 150, 0.01, 0.1, 0.2)
    b2, o2 = synth.sample_reads(rng, genomes[4:], 1000, 150, 0.01, 0.1, 0.0)
    perm = rng.permutation(6000)
    allb = np.concatenate([b1, b2]).reshape(6000, 150)[perm].reshape(-1)
    bases, offs = allb, np.arange(6001, dtype=np.uint64) * np.uint64(150)
    n = len(offs) - 1
    unit = 30000          # small work units so that some taxa convert to dense and others stay sparse

    def fresh(**kw):
        c = binding.Classifier(device=local, max_reads=1 << 14, max_bases=4 << 20, sparse_set_slots=1 << 22,
                               work_unit_size=unit, **kw)
        c.set_taxonomy(*tax.parent_map())
        return c

    # reference state: one GPU, all reads (unit ids given explicitly so that every variant cuts the same units)
    units, _, _ = synth.work_unit_ids(offs, unit)
    single = fresh()
    single.stage_db(kdb, idx)
    want_res = single.classify(bases, offs, unit_id=units)
    single.finish()
    want = single.counts()

    # ---- replicas -------------------------------------------------------------------------------------------
    lo, hi = kdist.partition(n, world, rank)
    # keep whole work units together: move the cut to the next unit boundary
    while lo > 0 and lo < n and units[lo] == units[lo - 1]:
        lo += 1
    while hi < n and hi > 0 and units[hi] == units[hi - 1]:
        hi += 1
    rep = fresh()
    rep.stage_db(kdb, idx)
    o = np.ascontiguousarray(offs[lo:hi + 1])
    got_res = rep.classify(bases, o, unit_id=units[lo:hi]) if hi > lo else None
    rep.finish()
    if got_res is not None:
        assert np.array_equal(got_res["call"], want_res["call"][lo:hi])
    kdist.merge_classifier_state_partitioned(rep, dev)
    got = rep.counts()
    for key in ("taxid", "n_reads", "n_kmers", "sparse", "unique"):
        assert np.array_equal(got[key], want[key]), (rank, "replicas", key, got[key], want[key])
    for t in want["taxid"].tolist():
        assert np.array_equal(rep.registers(t), single.registers(t))
    # clade roll-ups over the code partitions (sum of the per-GPU union histograms) == the single-GPU clade sketches
    from tests import util
    members = util.clade_members(tax.rows, want["taxid"])
    clades = [members[t] for t in sorted(members)]
    got_cl = kdist.clade_counts_distributed(rep, clades, dev)
    for c, g in zip(clades, got_cl):
        assert g == single.clade(c), (rank, "replica clades", c, g, single.clade(c))

    # ---- minimizer-range shards ----------------------------------------------------------------------------------
    n_bins = 1 << 18
    idx_off = np.frombuffer(idx[8:].tobytes(), np.uint64)
    # byte-balanced cut points like prepare_chunking (krakendb.cpp:463-522): by record count
    targets = [idx_off[-1] * r // world for r in range(world + 1)]
    cuts = [int(np.searchsorted(idx_off, t, side="left")) for t in targets]
    cuts[0], cuts[-1] = 0, n_bins
    sh = fresh()
    all_t, _ = single.db_taxids()
    sh.set_db_taxid_universe(all_t)
    sh.stage_db(kdb, idx, cuts[rank], cuts[rank + 1])
    d_bases = torch.from_numpy(np.concatenate([bases, np.full(64, ord("N"), np.uint8)])).to(dev)
    d_offs = torch.from_numpy(np.concatenate([offs, offs[-1:]]).astype(np.int64)).to(dev)
    codes = torch.zeros(int(offs[-1]) + 64, dtype=torch.int32, device=dev)
    sh.lookup_device(0, d_bases.data_ptr(), d_offs.data_ptr(), n, int(offs[-1]), codes.data_ptr(), only_hits=1)
    sh.sync(0)
    dist.all_reduce(codes, op=dist.ReduceOp.MAX)            # a key hits in at most one range (classify.cpp:447)
    # every rank resolves its own share of the reads from the merged codes
    d_units = torch.from_numpy(units.astype(np.int32)).to(dev)
    if hi > lo:
        sub_offs = d_offs[lo:hi + 2].contiguous()
        # offsets stay relative to d_bases; the slice must be 16-byte aligned for the bulk copy
        assert sub_offs.data_ptr() % 16 == 0
        sh.resolve_device(0, d_bases.data_ptr(), sub_offs.data_ptr(), hi - lo, int(offs[-1]), codes.data_ptr(),
                          d_units[lo:hi].contiguous().data_ptr())
        sh.sync(0)
    sh.finish()
    kdist.merge_classifier_state(sh, dev)
    got = sh.counts()
    for key in ("taxid", "n_reads", "n_kmers", "sparse", "unique"):
        assert np.array_equal(got[key], want[key]), (rank, "shards", key, got[key], want[key])

    # ---- shards with the fused lookup + NVLink peer scatter (no reduction of ids) -------------------------------
    def share(r):
        a, b = kdist.partition(n, world, r)
        while a > 0 and a < n and units[a] == units[a - 1]:
            a += 1
        while b < n and b > 0 and units[b] == units[b - 1]:
            b += 1
        return a, b
    shares = [share(r) for r in range(world)]
    assert shares[rank] == (lo, hi)
    sp = fresh()
    sp.set_db_taxid_universe(all_t)
    sp.stage_db(kdb, idx, cuts[rank], cuts[rank + 1])
    nbytes = (int(offs[-1]) + 64) * 4
    my_buf = sp.device_alloc(nbytes)
    sp.device_memset(0, my_buf, 0, nbytes)
    sp.sync(0)
    handles = [None] * world
    dist.all_gather_object(handles, sp.ipc_export(my_buf))
    peers = [my_buf if r == rank else sp.ipc_open(handles[r]) for r in range(world)]
    bounds = np.array([int(offs[shares[r][0]]) for r in range(world)] + [int(offs[-1])], np.uint64)
    bounds[0] = 0
    dist.barrier()                                         # every buffer is zeroed
    sp.lookup_device_peers(0, d_bases.data_ptr(), d_offs.data_ptr(), n, int(offs[-1]), peers, bounds)
    sp.sync(0)
    torch.cuda.synchronize()
    dist.barrier()                                         # every rank's hits have landed
    if hi > lo:
        sub_offs = d_offs[lo:hi + 2].contiguous()
        sp.resolve_device(0, d_bases.data_ptr(), sub_offs.data_ptr(), hi - lo, int(offs[-1]), my_buf,
                          d_units[lo:hi].contiguous().data_ptr())
        sp.sync(0)
    sp.finish()
    kdist.merge_classifier_state(sp, dev)
    got = sp.counts()
    for key in ("taxid", "n_reads", "n_kmers", "sparse", "unique"):
        if not np.array_equal(got[key], want[key]):
            bad = np.nonzero(got[key] != want[key])[0]
            raise AssertionError((rank, "peer shards", key, bad[:8].tolist(), got["taxid"][bad[:8]].tolist(),
                                  got[key][bad[:8]].tolist(), want[key][bad[:8]].tolist(), got["sparse"][bad[:8]].tolist()))
    dist.barrier()
    for r in range(world):
        if r != rank:
            sp.ipc_close(peers[r])
    dist.barrier()
    sp.device_free(my_buf)

    # ---- shards as bench.py runs them: the finder of a hit does its sketch work, device flags instead of host
    # barriers, double-buffered id buffers, several steps, code-partitioned merge at the end ------------------------
    sf = fresh()
    sf.set_db_taxid_universe(all_t)
    sf.stage_db(kdb, idx, cuts[rank], cuts[rank + 1])
    sf.set_shard_counting(True)
    n_steps = 3
    step_reads = [kdist.partition(n, n_steps, i) for i in range(n_steps)]
    # cut the steps at unit boundaries too
    def unit_cut(a):
        while 0 < a < n and units[a] == units[a - 1]:
            a += 1
        return a
    step_reads = [(unit_cut(a), unit_cut(b)) for a, b in step_reads]
    bufs = [sf.device_alloc(nbytes) for _ in range(2)]
    flg = sf.device_alloc(256)
    ft = kdist.device_view(flg, 256, torch.int64, dev)
    ft.zero_(); ft[16:24] = 2
    for b_ in bufs:
        sf.device_memset(0, b_, 0, nbytes)
    sf.sync(0); torch.cuda.synchronize()
    hs = [None] * world
    dist.all_gather_object(hs, (sf.ipc_export(bufs[0]), sf.ipc_export(bufs[1]), sf.ipc_export(flg)))
    opened, pb, pf = [], [[], []], []
    for r in range(world):
        if r == rank:
            pb[0].append(bufs[0]); pb[1].append(bufs[1]); pf.append(flg)
        else:
            m = [sf.ipc_open(h) for h in hs[r]]
            opened += m
            pb[0].append(m[0]); pb[1].append(m[1]); pf.append(m[2])
    dist.barrier()
    for i, (sa, sb) in enumerate(step_reads):
        # owners of this step's reads: contiguous shares cut at unit boundaries
        cutsr = [unit_cut(sa + (sb - sa) * r // world) for r in range(world)] + [sb]
        a, b = cutsr[rank], cutsr[rank + 1]
        bnd = np.array([int(offs[c]) for c in cutsr], np.uint64)
        bnd[0], bnd[-1] = int(offs[sa]), int(offs[sb])
        sub = d_offs[sa:sb + 2].contiguous()
        assert sub.data_ptr() % 16 == 0 or sa == sb
        par = i & 1
        sf.wait_flags(0, flg + 128, world, i + 1)
        if sb > sa:
            sf.lookup_device_peers(0, d_bases.data_ptr(), sub.data_ptr(), sb - sa, int(offs[-1]), pb[par], bnd)
        sf.signal_peers(0, pf, rank, i + 1)
        sf.wait_flags(0, flg, world, i + 1)
        if b > a:
            so = d_offs[a:b + 2].contiguous()
            assert so.data_ptr() % 16 == 0
            sf.resolve_device(0, d_bases.data_ptr(), so.data_ptr(), b - a, int(offs[-1]), bufs[par],
                              d_units[a:b].contiguous().data_ptr())
        sf.device_memset(0, bufs[par], 0, nbytes)
        sf.signal_peers(0, [p_ + 128 for p_ in pf], rank, i + 3)
        sf.sync(0)
    kdist.merge_classifier_state_partitioned(sf, dev)
    got = sf.counts()
    for key in ("taxid", "n_reads", "n_kmers", "sparse", "unique"):
        if not np.array_equal(got[key], want[key]):
            bad = np.nonzero(got[key] != want[key])[0]
            raise AssertionError((rank, "flag shards", key, bad[:8].tolist(), got["taxid"][bad[:8]].tolist(),
                                  got[key][bad[:8]].tolist(), want[key][bad[:8]].tolist()))
    for t in want["taxid"].tolist():
        assert np.array_equal(sf.registers(t), single.registers(t)), ("flag shards registers", t)
    got_cl = kdist.clade_counts_distributed(sf, clades, dev)
    for c, g in zip(clades, got_cl):
        assert g == single.clade(c), (rank, "flag shard clades", c, g, single.clade(c))
    dist.barrier()
    for p_ in opened:
        sf.ipc_close(p_)
    dist.barrier()

    # and against the oracle (rank 0)
    if rank == 0:
        from oracle.oracle_py import Oracle
        o_ = Oracle()
        run = o_.run(o_.open_db(kdb, idx), o_.parent_map(*tax.parent_map()), unit, 0)
        calls, _, _ = run.classify(bases, offs, want_codes=False)
        run.finish()
        oc = run.counts()
        assert np.array_equal(calls, want_res["call"])
        assert 5 < int(want["sparse"].sum()) < len(want["taxid"]) - 2, "the case must mix sparse and dense taxa"
        for key in ("taxid", "n_reads", "n_kmers", "sparse", "unique"):
            assert np.array_equal(oc[key], want[key]), ("oracle", key)
        print(f"multigpu_check OK on {world} GPUs: replicas (code-partitioned merge), minimizer-range shards (NCCL id merge), shards with NVLink "
              f"peer scatter, and flag-synchronised shards with finder-side counting reproduce the single-GPU state "
              f"({len(want['taxid'])} taxa, {int(want['sparse'].sum())} sparse)")
    # ---- the drop-in executable on several GPUs of one process (replicas, round-robin batches, kuq_merge_into) -----------
    if rank == 0:
        import subprocess
        import tempfile
        from krakenuniq_b200 import build
        G = util.GOLDEN
        exe = build.build_classify()
        with tempfile.TemporaryDirectory() as td:
            # replicas (database on every GPU) and shards (one minimizer range per GPU: the layout classify picks for a
            # database that only fits the cards together; its sketches follow the -x rule → the reference's -x goldens)
            for tag, extra, more_env in (("preload", ["-M"], {}), ("preload_u20000", ["-M", "-u", "20000"], {}),
                                         ("chunked", ["-M"], {"KUQ_FORCE_SHARDS": "1"})):
                out, rep = os.path.join(td, tag + ".kraken"), os.path.join(td, tag + ".report.tsv")
                cmd = [exe, "-d", os.path.join(G, "database.kdb"), "-i", os.path.join(G, "database.idx"), "-a", os.path.join(G, "taxDB"),
                       "-t", "4", "-r", rep, "-o", out] + extra + [os.path.join(G, "reads.fa")]
                env = dict(os.environ, KUQ_DEVICES=",".join(str(d) for d in range(world)), KUQ_SPARSE_SLOTS=str(1 << 22),
                           KUQ_BATCH_READS="200", **more_env)        # small batches so that every device gets work
                r = subprocess.run(cmd, capture_output=True, text=True, env=env)
                assert r.returncode == 0, r.stderr[-2000:]
                assert (f"sharded over {world} GPUs" if more_env else f"{world} GPUs") in r.stderr, r.stderr[-500:]
                assert open(out).read() == open(os.path.join(G, tag + ".kraken")).read(), tag

                def rows(path):
                    d = {}
                    for line in open(path):
                        if not line.startswith("#") and not line.startswith("%"):
                            f = line.rstrip("\n").split("\t")
                            d[f[6]] = f
                    return d
                assert rows(rep) == rows(os.path.join(G, tag + ".report.tsv")), tag
        print(f"multigpu_check OK: classify on {world} GPUs reproduces the reference's golden Kraken output and report")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
