"""Multi-process plumbing (one process per GPU, torch.distributed): how reads are partitioned over ranks and how the
per-taxon state is merged at the end of a run.

The merge is the cross-process form of `taxon_counts[t] += my_taxon_counts[t]` (classify.cpp:542-544):
  counters add                         ReadCounts::operator+=     readcounts.hpp:76-88
  dense registers: element-wise max    HLL merge dense+dense       hyperloglogplus.cpp:614-620
  sparse tier: union of the code sets  HLL merge sparse+sparse     hyperloglogplus.cpp:600-603
  a taxon is dense if it is dense anywhere (sparse+dense → dense, :604-612)
Works on CPU tensors with gloo (tests) and on HBM tensors with NCCL over NVLink (bench / production).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def partition(n_items: int, world: int, rank: int):
    """contiguous, balanced [begin, end) of rank's share of n_items (reads of a file, bins of a database)"""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def merge_state_tensors(regs: torch.Tensor, n_kmers: torch.Tensor, n_reads: torch.Tensor,
                        dense_flag: torch.Tensor | None = None, group=None):
    """in-place all-reduce of the per-taxon state"""
    dist.all_reduce(regs, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(n_kmers, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(n_reads, op=dist.ReduceOp.SUM, group=group)
    if dense_flag is not None:
        dist.all_reduce(dense_flag, op=dist.ReduceOp.MAX, group=group)


def gather_sparse_keys(keys: torch.Tensor, group=None) -> torch.Tensor:
    """all-gather variable-length int64 key arrays (0 = padding); returns the keys of all OTHER ranks"""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = torch.tensor([keys.numel()], dtype=torch.int64, device=keys.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    m = int(max(int(s.item()) for s in sizes))
    if m == 0:
        return keys.new_zeros(0)
    padded = keys.new_zeros(m)
    padded[:keys.numel()] = keys
    out = [keys.new_zeros(m) for _ in range(world)]
    dist.all_gather(out, padded, group=group)
    return torch.cat([out[r][:int(sizes[r].item())] for r in range(world) if r != rank]) if world > 1 else keys.new_zeros(0)


def device_view(ptr: int, nbytes: int, dtype: torch.dtype, device):
    """torch view of library-owned device memory (for the collectives only)"""
    class _Holder:
        pass
    h = _Holder()
    itemsize = torch.empty(0, dtype=dtype).element_size()
    typestr = {torch.uint8: "|u1", torch.int64: "<i8", torch.int32: "<i4"}[dtype]
    h.__cuda_array_interface__ = {"shape": (int(nbytes // itemsize),), "typestr": typestr, "data": (int(ptr), False),
                                  "version": 3}
    return torch.as_tensor(h, device=device)


def merge_classifier_state(clf, device, group=None):
    """End-of-run merge of a `binding.Classifier`'s state across all ranks (replicas or database shards)."""
    sp = clf.state_ptrs()
    regs = device_view(sp.d_regs, sp.regs_bytes, torch.uint8, device)
    nk = device_view(sp.d_n_kmers, sp.n_sketch * 8, torch.int64, device)
    nr = device_view(sp.d_n_reads, sp.n_taxa * 8, torch.int64, device)
    flag = device_view(sp.d_dense_flag, sp.n_sketch, torch.uint8, device)
    merge_state_tensors(regs, nk, nr, flag, group)
    # sparse tier: union of the (taxon, code) sets of the taxa that are still sparse everywhere
    n = clf.sparse_export()
    keys = torch.zeros(max(n, 1), dtype=torch.int64, device=device)
    if n:
        n2 = clf.sparse_export(keys.data_ptr(), keys.numel())
        keys = keys[:n2]
    else:
        keys = keys[:0]
    others = gather_sparse_keys(keys, group)
    if others.numel():
        clf.sparse_import(others.data_ptr(), others.numel())
    torch.cuda.synchronize()
