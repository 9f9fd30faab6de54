"""Large synthetic KrakenUniq workloads generated ON the GPU with torch (BASELINE.json configs[1..]: an "8 GB
synthetic KrakenDB (k=31)" and 150 bp reads).  Workload generation only — torch is plumbing here; nothing in this
file is on the classification path.  Layouts are the on-disk ones, validated against the
reference's db_sort at small scale by tests/test_synth_gpu.py.

Recipe: one long uniform-random sequence cut into `n_genomes` equal "genomes" under a
4-level taxonomy; every canonical k-mer of the sequence becomes one record labelled with the taxid of the genome
it starts in; records sorted by (bin_key, key); reads sampled from the sequence (substitutions, reverse
complement) plus a fraction of unrelated random reads.
"""
from __future__ import annotations

import numpy as np
import torch

INDEX2_XOR_MASK = 0xE37E28C4271B5A2D


def _lsr(x: torch.Tensor, s: int) -> torch.Tensor:
    """logical shift right on int64"""
    return (x >> s) & ((1 << (64 - s)) - 1)


def revcomp(x: torch.Tensor, n: int) -> torch.Tensor:
    """krakendb.cpp:218-225 on int64 tensors holding unsigned 2n-bit values"""
    x = ((x >> 2) & 0x3333333333333333) | ((x & 0x3333333333333333) << 2)
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0F) | ((x & 0x0F0F0F0F0F0F0F0F) << 4)
    x = ((x >> 8) & 0x00FF00FF00FF00FF) | ((x & 0x00FF00FF00FF00FF) << 8)
    x = ((x >> 16) & 0x0000FFFF0000FFFF) | ((x & 0x0000FFFF0000FFFF) << 16)
    x = ((x >> 32) & 0xFFFFFFFF) | (x << 32)
    return _lsr(~x, 64 - 2 * n)


def canonical(x: torch.Tensor, n: int) -> torch.Tensor:
    return torch.minimum(x, revcomp(x, n))


def bin_key(kmers: torch.Tensor, k: int, nt: int, idx_type: int = 2) -> torch.Tensor:
    """krakendb.cpp:200-215 (int64 in, int64 out)"""
    mask = (1 << (2 * nt)) - 1
    xor = (0 if idx_type == 1 else INDEX2_XOR_MASK) & mask
    x = kmers.clone()
    best = torch.full_like(x, (1 << 62))
    for _ in range(k - nt + 1):
        best = torch.minimum(best, canonical(x & mask, nt) ^ xor)
        x >>= 2
    return best


def forward_kmers(codes: torch.Tensor, start: int, count: int, k: int) -> torch.Tensor:
    """k-mers of windows [start, start+count) of a uint8 code tensor (first base most significant)"""
    km = torch.zeros(count, dtype=torch.int64, device=codes.device)
    for j in range(k):
        km = (km << 2) | codes[start + j:start + j + count].to(torch.int64)
    return km


def make_taxonomy_rows(n_genomes: int, first_id: int = 100):
    """root(1) → families → genera → species(genomes); returns (rows, species_taxids)"""
    n_gen = max(1, n_genomes // 20)
    n_fam = max(1, n_gen // 10)
    rows = [(1, 1, "root", "no rank")]
    nid = first_id
    fam = list(range(nid, nid + n_fam)); nid += n_fam
    gen = list(range(nid, nid + n_gen)); nid += n_gen
    sp = list(range(nid, nid + n_genomes))
    rows += [(f, 1, f"family{i}", "family") for i, f in enumerate(fam)]
    rows += [(g, fam[i % n_fam], f"genus{i}", "genus") for i, g in enumerate(gen)]
    rows += [(s, gen[i % n_gen], f"species{i}", "species") for i, s in enumerate(sp)]
    return rows, sp


class _DeviceScanner:
    """canonical k-mer + minimizer bin of every window of a code tensor through the product's own scan stage
    (kuq_scan_device → k_scan): two orders of magnitude faster than the elementwise torch loops above, which remain
    the CPU / validation path (tests/test_synth_gpu.py compares both with the reference's db_sort)."""
    PIECE = 352                                   # pseudo-read length (16-byte multiple): k_scan runs a warp per read

    def __init__(self, device: torch.device, k: int, nt: int, idx_type: int, chunk: int):
        from . import binding
        self.k, self.nt, self.idx_type, self.dev = k, nt, idx_type, device
        self.win = self.PIECE - (k - 1)
        n_pieces = -(-chunk // self.win)
        self.clf = binding.Classifier(device=device.index or 0, n_slots=1, max_reads=n_pieces + 64,
                                      max_bases=n_pieces * self.PIECE + 4096, sparse_set_slots=1024)
        self.lut = torch.tensor([ord(c) for c in "ACGT"], dtype=torch.uint8, device=device)

    def scan(self, codes: torch.Tensor, start: int, count: int):
        """windows [start, start+count) of `codes` → (canonical k-mers int64, bins int64)"""
        k, win, P = self.k, self.win, self.PIECE
        n_pieces = -(-count // win)
        idx = (torch.arange(n_pieces, device=self.dev, dtype=torch.int64)[:, None] * win + start +
               torch.arange(P, device=self.dev, dtype=torch.int64)[None, :])
        idx.clamp_(max=codes.numel() - 1)
        text = torch.empty(n_pieces * P + 64, dtype=torch.uint8, device=self.dev)
        text[:n_pieces * P] = self.lut[codes[idx.reshape(-1)].to(torch.int64)]
        text[n_pieces * P:] = ord("N")
        del idx
        offs = torch.arange(n_pieces + 2, dtype=torch.int64, device=self.dev) * P
        offs[-1] = offs[-2]
        canon = torch.empty(n_pieces * P + 64, dtype=torch.int64, device=self.dev)
        bins = torch.empty(n_pieces * P + 64, dtype=torch.int32, device=self.dev)
        torch.cuda.synchronize(self.dev)
        self.clf.scan_device(0, k, self.nt, self.idx_type, text.data_ptr(), offs.data_ptr(), n_pieces, n_pieces * P,
                             canon.data_ptr(), bins.data_ptr())
        self.clf.sync(0)
        km = canon[:n_pieces * P].view(n_pieces, P)[:, :win].reshape(-1)[:count]
        bn = bins[:n_pieces * P].view(n_pieces, P)[:, :win].reshape(-1)[:count].to(torch.int64)
        return km, bn

    def close(self):
        self.clf.close()


class GpuDatabase:
    """A synthetic database resident in HBM in the on-disk layout."""

    def __init__(self, n_records: int, n_genomes: int = 2000, k: int = 31, nt: int = 15, idx_type: int = 2,
                 seed: int = 2, device: str = "cuda:0", chunk: int = 1 << 26, passes: int = 1,
                 shard: tuple[int, int] | None = None, use_kernel_scan: bool = True, defer_build: bool = False):
        """`passes` > 1 builds the database one minimizer range at a time (temporaries of one range only), for
        databases whose sort would not fit next to the result (tens of GB and more, > 2^32 records).
        `shard` = (rank, world): keep only this rank's minimizer range of a database cut into `world` ranges with
        about equal record counts (every rank derives the same cut points from the same sample); `offsets` is then
        the local slice for bins [bin_lo, bin_hi] starting at 0."""
        self.k, self.nt, self.idx_type = k, nt, idx_type
        dev = torch.device(device)
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed)
        n_pos = n_records                       # one window per position (duplicates are removed below)
        self.genome = torch.empty(n_pos + k - 1, dtype=torch.uint8, device=dev)
        for a in range(0, n_pos + k - 1, 1 << 30):      # randint in pieces: no multi-GB int64 temporaries
            c = min(1 << 30, n_pos + k - 1 - a)
            self.genome[a:a + c] = torch.randint(0, 4, (c,), dtype=torch.uint8, device=dev, generator=gen)
        self.rows, self.species = make_taxonomy_rows(n_genomes)
        self.genome_len = (n_pos + n_genomes - 1) // n_genomes
        sp = torch.tensor(self.species, dtype=torch.int32, device=dev)
        n_bins = 1 << (2 * nt)
        # minimizer-range cut points for the shards / passes, from the bins of a sample (bins are heavily skewed)
        self.bin_lo, self.bin_hi = 0, n_bins
        n_parts = (shard[1] if shard else 1) * passes
        if n_parts > 1:
            m = min(n_pos, 1 << 24)
            sb = bin_key(canonical(forward_kmers(self.genome, 0, m, k), k), k, nt, idx_type)
            q = torch.quantile(sb.to(torch.float64)[:: max(1, m >> 20)], torch.linspace(0, 1, n_parts + 1, device=dev,
                                                                                  dtype=torch.float64))
            cuts = [0] + [int(x) for x in q[1:-1].tolist()] + [n_bins]
            del sb
            if shard:
                cuts = cuts[shard[0] * passes:(shard[0] + 1) * passes + 1]
                self.bin_lo, self.bin_hi = cuts[0], cuts[-1]
        else:
            cuts = [0, n_bins]
        self._chunk, self._n_pos, self._sp = chunk, n_pos, sp
        scanner = _DeviceScanner(dev, k, nt, idx_type, chunk) if dev.type == "cuda" and use_kernel_scan else None
        selective = n_parts > 1
        self.cuts = cuts
        if defer_build:                                    # stream_ranges() builds and hands out one range at a time
            self._scanner = scanner
            self.key_ct, self.records, self.offsets = 0, None, None
            return
        n_rows = n_pos if not shard else int(n_pos / shard[1] * 1.10) + (1 << 20)
        rec = torch.empty((n_rows, 3), dtype=torch.int32, device=dev)     # (n, 3) int32 == packed 12-byte records
        counts = torch.zeros(self.bin_hi - self.bin_lo, dtype=torch.int64, device=dev)
        out = 0
        for pi in range(len(cuts) - 1):
            lo, hi = cuts[pi], cuts[pi + 1]
            if hi <= lo:
                continue
            keys, taxa, bins = self._build_range(lo, hi, selective, scanner)
            n = keys.numel()
            assert out + n <= n_rows, "shard larger than planned: raise the row reserve"
            rec[out:out + n, 0] = (keys & 0xFFFFFFFF).to(torch.int32)      # wraps to the same 32 bits
            rec[out:out + n, 1] = (keys >> 32).to(torch.int32)
            rec[out:out + n, 2] = taxa
            out += n
            counts += torch.bincount(bins.to(torch.int64) - self.bin_lo, minlength=self.bin_hi - self.bin_lo)
            del keys, taxa, bins
            torch.cuda.empty_cache() if dev.type == "cuda" else None
        if scanner is not None:
            scanner.close()
        self.key_ct = out
        self.records = rec[:out]
        off = torch.zeros(self.bin_hi - self.bin_lo + 1, dtype=torch.int64, device=dev)
        torch.cumsum(counts, 0, out=off[1:])
        del counts
        self.offsets = off
        if dev.type == "cuda":
            torch.cuda.empty_cache()

    def _build_range(self, lo: int, hi: int, selective: bool, scanner):
        """records of minimizer range [lo, hi): (keys int64, taxa int32, bins int32) sorted by (bin, key), deduplicated"""
        k, nt, idx_type, chunk, n_pos, dev = self.k, self.nt, self.idx_type, self._chunk, self._n_pos, self.genome.device
        keys_l, bins_l, pos_l = [], [], []
        for a in range(0, n_pos, chunk):
            c = min(chunk, n_pos - a)
            if scanner is not None:
                km, bn = scanner.scan(self.genome, a, c)
            else:
                km = canonical(forward_kmers(self.genome, a, c, k), k)
                bn = bin_key(km, k, nt, idx_type)
            if selective:
                sel = (bn >= lo) & (bn < hi)
                keys_l.append(km[sel])
                bins_l.append(bn[sel].to(torch.int32))
                pos_l.append(torch.nonzero(sel).squeeze(1) + a)
                del sel
            else:
                keys_l.append(km)
                bins_l.append(bn.to(torch.int32))
            del km, bn
        keys = torch.cat(keys_l); del keys_l
        bins = torch.cat(bins_l); del bins_l
        if selective:
            pos = torch.cat(pos_l); del pos_l
        # sort by (bin, key): key sort, then a stable bin sort
        keys, order = torch.sort(keys)
        bins = bins[order]
        owner = (pos[order] if selective else order) // self.genome_len
        taxa = self._sp[owner.to(torch.int64)]
        del order, owner
        if selective:
            del pos
        # drop duplicate keys (same k-mer at two positions / palindromes): keep the first owner
        keep = torch.ones(keys.numel(), dtype=torch.bool, device=dev)
        keep[1:] = keys[1:] != keys[:-1]
        if not bool(keep.all()):
            keys, bins, taxa = keys[keep], bins[keep], taxa[keep]
        del keep
        bins, order = torch.sort(bins, stable=True)
        keys = keys[order]
        taxa = taxa[order]
        del order
        return keys, taxa, bins

    def stream_ranges(self):
        """for a database built with defer_build=True: yields (bin_lo, bin_hi, records (n, 3) int32, offsets int64
        [bin_hi - bin_lo + 1] ABSOLUTE record offsets) one minimizer range (= pass) at a time; nothing is kept"""
        base = 0
        for pi in range(len(self.cuts) - 1):
            lo, hi = self.cuts[pi], self.cuts[pi + 1]
            if hi <= lo:
                continue
            keys, taxa, bins = self._build_range(lo, hi, True, self._scanner)
            n = keys.numel()
            rec = torch.empty((n, 3), dtype=torch.int32, device=keys.device)
            rec[:, 0] = (keys & 0xFFFFFFFF).to(torch.int32)
            rec[:, 1] = (keys >> 32).to(torch.int32)
            rec[:, 2] = taxa
            off = torch.zeros(hi - lo + 1, dtype=torch.int64, device=keys.device)
            torch.cumsum(torch.bincount(bins.to(torch.int64) - lo, minlength=hi - lo), 0, out=off[1:])
            off += base
            base += n
            del keys, taxa, bins
            yield lo, hi, rec, off
            del rec, off
            if self.genome.device.type == "cuda":
                torch.cuda.empty_cache()
        self.key_ct = base
        if self._scanner is not None:
            self._scanner.close()
            self._scanner = None

    # ---- host images (for the reference binary / the oracle) --------------------------------------------------
    def kdb_header(self) -> np.ndarray:
        from . import synth
        return synth.kdb_header(self.k, self.key_ct)

    def write_files(self, kdb_path: str, idx_path: str, chunk: int = 1 << 26):
        with open(kdb_path, "wb") as f:
            f.write(self.kdb_header().tobytes())
            for a in range(0, self.key_ct, chunk):
                f.write(self.records[a:a + chunk].cpu().numpy().tobytes())
        with open(idx_path, "wb") as f:
            f.write(b"KRAKIDX" if self.idx_type == 1 else b"KRAKIX2")
            f.write(bytes([self.nt]))
            n = self.offsets.numel()
            for a in range(0, n, chunk):
                f.write(self.offsets[a:a + chunk].cpu().numpy().tobytes())

    def images(self):
        kdb = np.concatenate([self.kdb_header(), self.records.cpu().numpy().view(np.uint8).reshape(-1)])
        idx = np.concatenate([np.frombuffer(b"KRAKIDX" if self.idx_type == 1 else b"KRAKIX2", np.uint8),
                              np.array([self.nt], np.uint8), self.offsets.cpu().numpy().view(np.uint8)])
        return kdb, idx

    def parent_map(self):
        taxid = np.array([r[0] for r in self.rows], np.uint32)
        parent = np.array([0 if r[1] == r[0] else r[1] for r in self.rows], np.uint32)
        return taxid, parent

    def write_taxdb(self, path):
        with open(path, "w") as f:
            for t, p, name, rank in self.rows:
                f.write(f"{t}\t{p}\t{name}\t{rank}\n")

    # ---- reads --------------------------------------------------------------------------------------------------
    def sample_reads(self, n_reads: int, read_len: int = 150, seed: int = 3, random_frac: float = 0.2,
                     sub_rate: float = 0.01, chunk: int = 1 << 20):
        """→ (bases uint8 [n_reads*read_len + 64 slack] ASCII on the GPU, offsets int64 [n_reads+1] on the GPU)"""
        dev = self.genome.device
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed)
        out = torch.full((n_reads * read_len + 64,), ord("N"), dtype=torch.uint8, device=dev)
        lut = torch.tensor([ord(c) for c in "ACGT"], dtype=torch.uint8, device=dev)
        ar = torch.arange(read_len, device=dev)
        max_start = self.genome.numel() - read_len
        for a in range(0, n_reads, chunk):
            c = min(chunk, n_reads - a)
            start = torch.randint(0, max_start, (c,), device=dev, generator=gen)
            codes = self.genome[(start[:, None] + ar[None, :])]
            sub = torch.rand((c, read_len), device=dev, generator=gen) < sub_rate
            delta = torch.randint(1, 4, (c, read_len), device=dev, generator=gen, dtype=torch.uint8)
            codes = torch.where(sub, (codes + delta) & 3, codes)
            rc = torch.rand((c,), device=dev, generator=gen) < 0.5
            codes = torch.where(rc[:, None], (3 - codes).flip(1), codes)
            rnd = torch.rand((c,), device=dev, generator=gen) < random_frac
            noise = torch.randint(0, 4, (c, read_len), device=dev, generator=gen, dtype=torch.uint8)
            codes = torch.where(rnd[:, None], noise, codes)
            out[a * read_len:(a + c) * read_len] = lut[codes.to(torch.int64)].reshape(-1)
            del codes, sub, delta, noise
        offsets = torch.arange(n_reads + 1, dtype=torch.int64, device=dev) * read_len
        return out, offsets
