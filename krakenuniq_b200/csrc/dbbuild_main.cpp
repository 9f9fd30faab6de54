// Drop-in `db_sort` and `set_lcas` executables over libkuq's C ABI — SURVEY.md §8 row f4.
// Built twice from this file: plain → db_sort (db_sort.cpp:41-187), -DTOOL_SET_LCAS → set_lcas (set_lcas.cpp:92-600).
// Same getopt strings, same file side effects; the work itself (minimizer binning + sort, k-mer lookup + LCA update)
// runs on the GPU.  ROUND 1 STATUS: written after the round's GPU budget was spent — compiles, not yet run on
// hardware.  The oracle's CPU statements of both tools are pinned against the reference executables
// (tests/test_oracle_db_build.py).
//
// set_lcas options NOT supported (exit with a message): -I (UIDs), -a / -A (new taxonomy ids).
#include <fcntl.h>
#include <getopt.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sysexits.h>
#include <unistd.h>

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/kuq.h"

using namespace std;

[[noreturn]] static void die(int code, const string &msg, const char *tool) {
  cerr << tool << ": " << msg << endl;
  exit(code);
}

struct Mapped {
  void *p = NULL;
  size_t size = 0;
  void open(const string &path, const char *tool) {
    int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) die(EX_OSERR, "unable to open " + path, tool);
    struct stat sb;
    if (fstat(fd, &sb) < 0) die(EX_OSERR, "unable to fstat " + path, tool);
    size = sb.st_size;
    p = size ? mmap(0, size, PROT_READ, MAP_PRIVATE, fd, 0) : NULL;
    if (size && p == MAP_FAILED) die(EX_OSERR, "unable to mmap " + path, tool);
    ::close(fd);
  }
};

static void write_file(const string &path, const void *data, size_t size, const char *tool) {
  ofstream out(path.c_str(), ofstream::binary);
  if (!out.is_open()) die(EX_CANTCREAT, "can't write " + path, tool);
  out.write((const char *)data, (streamsize)size);
  out.close();
  if (!out) die(EX_IOERR, "error writing " + path, tool);
}

#ifndef TOOL_SET_LCAS
// ====================================================================================================================
// db_sort
// ====================================================================================================================
static const char *TOOL = "db_sort";
static void usage(int exit_code = EX_USAGE) {
  cerr << "Usage: db_sort [-z] [-M] [-t threads] [-n nt] <-d input db> <-o output db> <-i output idx>\n";
  exit(exit_code);
}

int main(int argc, char **argv) {
  string input, output, index;
  uint32_t nt = 15;                                     // Bin_key_nt, db_sort.cpp:27
  bool zero_vals = false;
  if (argc > 1 && strcmp(argv[1], "-h") == 0) usage(0);
  int opt;
  while ((opt = getopt(argc, argv, "n:d:o:i:t:zM")) != -1) {
    switch (opt) {
      case 'n': {
        long long sig = atoll(optarg);
        if (sig < 1 || sig > 31) die(EX_USAGE, "bin key length out of range", TOOL);
        nt = (uint32_t)sig;
        break;
      }
      case 'd': input = optarg; break;
      case 'o': output = optarg; break;
      case 'i': index = optarg; break;
      case 'M': break;                                  // the images are in memory anyway
      case 't':
        if (atoll(optarg) <= 0) die(EX_USAGE, "can't use nonpositive thread count", TOOL);
        break;
      case 'z': zero_vals = true; break;
      default: usage();
    }
  }
  if (input.empty() || output.empty() || index.empty()) usage();
  if (nt > 15) die(EX_USAGE, "bin key lengths above 15 overflow the reference's own index arithmetic (krakendb.cpp:204)", TOOL);
  cerr << "db_sort: Getting database into memory ...";
  Mapped in;
  in.open(input, TOOL);
  if (in.size < 56) die(EX_DATAERR, "input database too short", TOOL);
  uint64_t key_bits, key_ct;
  memcpy(&key_bits, (const char *)in.p + 8, 8);
  memcpy(&key_ct, (const char *)in.p + 48, 8);
  const uint64_t key_len = key_bits / 8 + !!(key_bits % 8), header = 72 + 2 * (4 + 8 * key_bits);
  vector<char> kdb(header + key_ct * (key_len + 4)), idx(8 + 8 * ((1ull << (2 * nt)) + 1));
  cerr << "db_sort: Sorting ...";
  char err[512] = "";
  int dev = getenv("KUQ_DEVICE") ? atoi(getenv("KUQ_DEVICE")) : 0;
  int rc = kuq_db_sort(dev, in.p, in.size, nt, zero_vals, kdb.data(), idx.data(), err, sizeof err);
  if (rc) die(rc == KUQ_E_NO_DEVICE ? EX_UNAVAILABLE : EX_DATAERR, err[0] ? err : kuq_strerror(rc), TOOL);
  cerr << "db_sort: Sorting complete - writing database to disk ..." << endl;
  write_file(index, idx.data(), idx.size(), TOOL);       // make_index, krakendb.cpp:141-147
  write_file(output, kdb.data(), kdb.size(), TOOL);      // db_sort.cpp:69-73
  return 0;
}

#else
// ====================================================================================================================
// set_lcas
// ====================================================================================================================
static const char *TOOL = "set_lcas";
static const size_t SKIP_LEN = 50000;                    // set_lcas.cpp:31
static bool Allow_extra_kmers = false, verbose = false, Operate_in_RAM = false, Pretend = false;
static uint32_t Minimum_sequence_size = 0;
static uint32_t Lca_flags = 0;                           // -T → KUQ_LCA_FORCE_CONTAMINANT, -R → KUQ_LCA_RESET
static string DB_filename, Index_filename, TaxDB_filename, File_to_taxon_map_filename, ID_to_taxon_map_filename,
    Multi_fasta_filename, Output_DB_filename, Kmer_count_filename;

static void usage(int exit_code = EX_USAGE) {
  cerr << "Usage: set_lcas [options]" << endl << endl
       << "Options: (*mandatory)" << endl
       << "* -d filename      Kraken DB filename" << endl
       << "* -i filename      Kraken DB index filename" << endl
       << "* -b filename      Taxonomy DB file" << endl
       << "  -t #             Number of threads" << endl
       << "  -M               Copy DB to RAM during operation" << endl
       << "  -o filename      Output database to filename, instead of overwriting the input database" << endl
       << "  -x               K-mers not found in DB do not cause errors" << endl
       << "  -f filename      File to taxon map" << endl
       << "  -F filename      Multi-FASTA file with sequence data" << endl
       << "  -m filename      Sequence ID to taxon map" << endl
       << "  -T               When a k-mer appears in a 'synthetic construct' sequence, force the taxID to be the 'synthetic construct' taxID, instead of the LCA." << endl
       << "  -R               Reset the taxIDs of the k-mers of the sequences to 0" << endl
       << "  -E #             Exclude sequences that are shorter than the threshold." << endl
       << "  -c filename      Write k-mer counts per taxon to filename" << endl
       << "  -p               Pretend - do not write database back to disk" << endl
       << "  -v               Verbose output" << endl
       << "  -h               Print this message" << endl << endl
       << "-F and -m must be specified together.  If -f is given, -F/-m are ignored." << endl
       << "(GPU build: -I, -a and -A are not supported)" << endl;
  exit(exit_code);
}

// FastaReader::next_sequence (seqreader.cpp:34-79): header = line after '>', id = its first word, sequence = the
// following lines glued together until the next '>' line
struct Fasta {
  ifstream file;
  string linebuffer;
  bool valid = false;
  explicit Fasta(const string &path) : file(path.c_str()) {
    if (file.rdstate() & ifstream::failbit) die(EX_NOINPUT, "can't open " + path, TOOL);
    valid = true;
  }
  bool next(string &id, string &header, string &seq) {
    id.clear(); header.clear(); seq.clear();
    if (!file.good()) { valid = false; return false; }
    string line;
    if (linebuffer.empty()) getline(file, line);
    else { line = linebuffer; linebuffer.clear(); }
    if (line.empty() || line[0] != '>') {
      cerr << "set_lcas: malformed fasta file - expected header char > not found" << endl;
      valid = false;
      return false;
    }
    header = line.substr(1);
    istringstream ss(header);
    ss >> id;
    while (file.good()) {
      getline(file, line);
      if (!line.empty() && line[0] == '>') { linebuffer = line; break; }
      seq += line;
    }
    return true;
  }
};

struct Batch {
  string bases;
  vector<uint64_t> offs{0};
  vector<uint32_t> taxids;
  uint32_t contaminant = 0;        // the one contaminant taxid (32630 / 81077) this batch holds, 0 = none yet
};
static kuq_ctx *Ctx = NULL;
static uint32_t K = 0;
static uint64_t Missing_total = 0;
static const uint64_t BATCH_NT = 64ull << 20;

static void flush(Batch &b) {
  if (b.taxids.empty()) return;
  uint64_t missing = 0;
  if (kuq_set_lcas_batch(Ctx, b.bases.data(), b.offs.data(), (uint32_t)b.taxids.size(), b.taxids.data(), Lca_flags, &missing))
    die(EX_SOFTWARE, kuq_last_error(Ctx), TOOL);
  if (missing && !Allow_extra_kmers) die(EX_DATAERR, "kmer found in sequence that is not in database", TOOL);   // :441-443
  Missing_total += missing;
  b.bases.clear(); b.offs.assign(1, 0); b.taxids.clear(); b.contaminant = 0;
}

// the reference's pieces: [i, i + SKIP_LEN + k - 1) for i = 0, SKIP_LEN, ... (set_lcas.cpp:363-364, 399-400)
static void add_sequence(Batch &b, const string &seq, uint32_t taxid) {
  if (taxid == 32630u || taxid == 81077u) {              // TID_CONTAMINANT1/2, set_lcas.cpp:88-89
    // under -T the first contaminant taxid to reach a k-mer stays: keep file order between the two kinds
    if (b.contaminant && b.contaminant != taxid) flush(b);
    b.contaminant = taxid;
  }
  for (size_t i = 0; i < seq.size(); i += SKIP_LEN) {
    const size_t len = min(seq.size() - i, SKIP_LEN + K - 1);
    if (b.bases.size() + len > BATCH_NT || b.taxids.size() >= (1u << 19)) flush(b);
    b.bases.append(seq, i, len);
    b.offs.push_back(b.bases.size());
    b.taxids.push_back(taxid);
  }
}

int main(int argc, char **argv) {
  if (argc > 1 && strcmp(argv[1], "-h") == 0) usage(0);
  int opt;
  while ((opt = getopt(argc, argv, "f:d:i:t:n:m:F:xMTRvb:aApI:o:Sc:E:")) != -1) {
    switch (opt) {
      case 'f': File_to_taxon_map_filename = optarg; break;
      case 'd': DB_filename = optarg; break;
      case 'i': Index_filename = optarg; break;
      case 'F': Multi_fasta_filename = optarg; break;
      case 'm': ID_to_taxon_map_filename = optarg; break;
      case 't':
        if (atoll(optarg) <= 0) die(EX_USAGE, "can't use nonpositive thread count", TOOL);
        break;
      case 'v': verbose = true; break;
      case 'x': Allow_extra_kmers = true; break;
      case 'b': TaxDB_filename = optarg; break;
      case 'c': Kmer_count_filename = optarg; break;
      case 'M': Operate_in_RAM = true; break;
      case 'o': Output_DB_filename = optarg; break;
      case 'E': Minimum_sequence_size = (uint32_t)atoi(optarg); break;
      case 'p': Pretend = true; break;
      case 'n': case 'S': break;
      case 'T': Lca_flags |= KUQ_LCA_FORCE_CONTAMINANT; break;
      case 'R': Lca_flags |= KUQ_LCA_RESET; break;
      case 'I': case 'a': case 'A':
        die(EX_USAGE, string("option -") + (char)opt + " is not supported by the GPU set_lcas", TOOL);
      default: usage();
    }
  }
  if (DB_filename.empty() || Index_filename.empty() || TaxDB_filename.empty()) usage();
  if (File_to_taxon_map_filename.empty() && (Multi_fasta_filename.empty() || ID_to_taxon_map_filename.empty())) usage();
  const bool one_fasta_file = File_to_taxon_map_filename.empty();

  // taxonomy → Parent_map (getParentMap, taxdb.hpp:383-398): taxid, parent taxid from the first two columns; a row
  // whose parent has no row of its own, or is the row itself, has parent 0
  vector<uint32_t> ids, parents;
  unordered_map<uint32_t, uint32_t> parent_of;
  {
    ifstream in(TaxDB_filename.c_str());
    if (!in.is_open()) die(EX_NOINPUT, "unable to open taxonomy index file " + TaxDB_filename, TOOL);
    string line;
    vector<pair<uint32_t, uint32_t>> rows;
    while (getline(in, line)) {
      size_t a = line.find('\t'), b = a == string::npos ? a : line.find('\t', a + 1);
      if (a == string::npos || b == string::npos) continue;
      uint32_t id = (uint32_t)strtoul(line.substr(0, a).c_str(), NULL, 10);
      uint32_t par = (uint32_t)strtoul(line.substr(a + 1, b - a - 1).c_str(), NULL, 10);
      if (parent_of.count(id)) continue;
      parent_of[id] = par;
      rows.emplace_back(id, par);
    }
    for (auto &r : rows) {
      if (r.first == 0) continue;
      ids.push_back(r.first);
      parents.push_back((r.second == r.first || !parent_of.count(r.second)) ? 0 : r.second);
    }
  }

  Mapped kdb, idx;
  kdb.open(DB_filename, TOOL);
  idx.open(Index_filename, TOOL);
  if (kdb.size < 56) die(EX_DATAERR, "database too short", TOOL);
  uint64_t key_bits;
  memcpy(&key_bits, (const char *)kdb.p + 8, 8);
  K = (uint32_t)(key_bits / 2);
  kuq_config cfg;
  kuq_config_default(&cfg);
  cfg.n_slots = 1;
  cfg.max_bases_per_batch = BATCH_NT + (1 << 20);
  cfg.hll_mode = KUQ_HLL_DENSE_ONLY;                      // no sketches are used here
  if (getenv("KUQ_DEVICE")) cfg.device = atoi(getenv("KUQ_DEVICE"));
  int rc = kuq_create(&cfg, &Ctx);
  if (rc) die(EX_UNAVAILABLE, string("libkuq: ") + kuq_strerror(rc), TOOL);
  if (kuq_stage_db(Ctx, kdb.p, kdb.size, idx.p, idx.size, 0, 0)) die(EX_DATAERR, kuq_last_error(Ctx), TOOL);
  if (kuq_set_taxonomy(Ctx, ids.data(), parents.data(), (uint32_t)ids.size())) die(EX_SOFTWARE, kuq_last_error(Ctx), TOOL);

  Batch batch;
  if (one_fasta_file) {                                   // process_single_file, set_lcas.cpp:272-386
    cerr << "Reading sequence ID to taxonomy ID mapping ... ";
    unordered_map<string, uint32_t> id_to_taxon;          // read_seqid_to_taxid_map, :200-270 (first mapping wins)
    {
      ifstream map_file(ID_to_taxon_map_filename.c_str());
      if (map_file.rdstate() & ifstream::failbit) die(EX_NOINPUT, "can't open " + ID_to_taxon_map_filename, TOOL);
      string line, seq_id;
      uint32_t taxid = 0;
      while (map_file.good()) {
        getline(map_file, line);
        if (line.empty()) break;
        istringstream iss(line);
        iss >> seq_id >> taxid;
        if (!id_to_taxon.count(seq_id)) id_to_taxon[seq_id] = taxid;
      }
      if (id_to_taxon.empty()) cerr << "Error: No ID mappings present!!" << endl;
      cerr << " got " << id_to_taxon.size() << " mappings." << endl;
    }
    Fasta reader(Multi_fasta_filename);
    const string prefix = "kraken:taxid|";
    uint32_t seqs_processed = 0, seqs_skipped = 0, seqs_no_taxid = 0;
    string id, header, seq;
    while (reader.valid) {
      if (!reader.next(id, header, seq)) break;
      if (seq.empty()) { ++seqs_skipped; continue; }
      uint32_t taxid = 0;
      auto it = id_to_taxon.find(id);
      if (it != id_to_taxon.end()) {
        taxid = it->second;
      } else {                                            // "NC_0001.2" → "NC_0001", :297-313
        size_t pos = id.find_last_of('.');
        bool num = pos != string::npos;
        for (size_t i = pos + 1; num && i < id.size(); ++i) num = isdigit((unsigned char)id[i]) != 0;
        if (num) {
          it = id_to_taxon.find(id.substr(0, pos));
          if (it != id_to_taxon.end()) taxid = it->second;
        }
      }
      if (taxid == 0 && id.size() >= prefix.size() && id.compare(0, prefix.size(), prefix) == 0) {   // :315-323
        taxid = (uint32_t)strtol(id.c_str() + prefix.size(), NULL, 10);
        if (taxid == 0) cerr << "Error: taxonomy ID is zero for sequence '" << id << "'?!" << endl;
      }
      if (taxid == 0) {
        cerr << "Error! Didn't find taxonomy ID mapping for sequence " << id << "!!" << endl;
        ++seqs_skipped;
        continue;
      }
      if (Minimum_sequence_size > 0 && seq.size() < Minimum_sequence_size) {
        cerr << "Skipping sequence " << id << " as it's too short (" << seq.size() << ")" << endl;
        ++seqs_skipped;
        continue;
      }
      if (!parent_of.count(taxid) || taxid == 0) {        // :336-341
        cerr << "Skipping sequence " << id << " since taxonomy ID " << taxid << " is not in taxonomy database!" << endl;
        ++seqs_skipped;
        continue;
      }
      add_sequence(batch, seq, taxid);
      ++seqs_processed;
      cerr << "\rProcessed " << seqs_processed << " sequences";
    }
    flush(batch);
    cerr << "\r                                                                            ";
    cerr << "\rFinished processing " << seqs_processed << " sequences (skipping " << seqs_skipped
         << " empty sequences, and " << seqs_no_taxid << " sequences with no taxonomy mapping)" << endl;
  } else {                                                // process_files, :388-411: one single-FASTA file per line
    cerr << "Processing files in " << File_to_taxon_map_filename << endl;
    ifstream map_file(File_to_taxon_map_filename.c_str());
    if (map_file.rdstate() & ifstream::failbit) die(EX_NOINPUT, "can't open " + File_to_taxon_map_filename, TOOL);
    string line;
    uint32_t seqs_processed = 0;
    while (map_file.good()) {
      getline(map_file, line);
      if (line.empty()) break;
      string filename;
      uint32_t taxid = 0;
      istringstream iss(line);
      iss >> filename;
      iss >> taxid;
      Fasta reader(filename);
      string id, header, seq;
      reader.next(id, header, seq);                       // the first sequence only, :414-421
      if (!parent_of.count(taxid) || taxid == 0)
        cerr << "Skipping " << filename << " since taxonomy ID " << taxid << " is not in taxonomy database!" << endl;
      else
        add_sequence(batch, seq, taxid);
      cerr << "\rProcessed " << ++seqs_processed << " sequences";
    }
    flush(batch);
    cerr << "\r                                                       ";
    cerr << "\rFinished processing " << seqs_processed << " sequences" << endl;
  }
  if (verbose && Missing_total) cerr << Missing_total << " k-mers of the library are not in the database" << endl;

  vector<char> out((const char *)kdb.p, (const char *)kdb.p + kdb.size);
  if (kuq_export_db_values(Ctx, out.data(), out.size())) die(EX_SOFTWARE, kuq_last_error(Ctx), TOOL);
  if (!Kmer_count_filename.empty()) {                     // count_taxons, krakendb.cpp:90-113 → :141-150
    cerr << "Writing kmer counts to " << Kmer_count_filename << "..." << endl;
    const uint64_t key_len = key_bits / 8 + !!(key_bits % 8), header = 72 + 2 * (4 + 8 * key_bits);
    uint64_t key_ct;
    memcpy(&key_ct, out.data() + 48, 8);
    map<uint32_t, uint64_t> counts;
    for (uint64_t i = 0; i < key_ct; i++) {
      uint32_t v;
      memcpy(&v, out.data() + header + i * (key_len + 4) + key_len, 4);
      ++counts[v];
    }
    ofstream ofs(Kmer_count_filename.c_str());
    for (auto &kv : counts) ofs << kv.first << '\t' << kv.second << '\n';
  }
  if (!Pretend) {
    // the reference rewrites the memory-mapped input (or the RAM copy under -M) and, with -o, copies it there
    // (:152-169); the net effect with -o is that the OUTPUT holds the new values — and, without -M, so does the input
    const string target = (!Output_DB_filename.empty() && Operate_in_RAM) ? Output_DB_filename : DB_filename;
    if (Operate_in_RAM) cerr << "Writing database from RAM back to " << target << " ..." << endl;
    write_file(target, out.data(), out.size(), TOOL);
    if (!Output_DB_filename.empty() && !Operate_in_RAM) write_file(Output_DB_filename, out.data(), out.size(), TOOL);
  }
  kuq_destroy(Ctx);
  return 0;
}
#endif
