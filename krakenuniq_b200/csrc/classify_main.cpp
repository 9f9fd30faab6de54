// classify_main.cpp — drop-in `classify` executable over libkuq.so.
//
// Same argv, file formats and text outputs as the reference driver (src/classify.cpp:150-346,1068-1189), so that
// scripts/krakenuniq:240-248 can call it unchanged; the per-read work (classify_sequence, :897-1012) happens on the
// GPU behind include/kuq.h.  Host-side pieces restated here, each citing what it mirrors:
//   FASTA/FASTQ readers            src/seqreader.cpp:34-129 (gz through zlib instead of bxzstr)
//   work-unit loop                 src/classify.cpp:506-559 (batches hold whole work units)
//   Kraken output lines            src/classify.cpp:980-1010, hitlist_string :826-861
//   stats / progress lines         src/classify.cpp:361-375, 555-558
//   taxDB reader, genome sizes     src/taxdb.hpp:563-605, 850-885
//   report (clades, sorting, cols) src/taxdb.hpp:928-1123, src/classify.cpp:286-328
// Not supported (exit with a message): -I uid mapping, -x together with several -d databases (the reference's own loop
// for that combination queries database 0 in every pass, classify.cpp:585-632).
#include <fcntl.h>
#include <getopt.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <sysexits.h>
#include <unistd.h>
#include <zlib.h>

#include <dlfcn.h>
#include <omp.h>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

#include <algorithm>
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/kuq.h"

using namespace std;

static vector<string> DB_filenames, Index_filenames;
static uint32_t Minimum_hit_count = 1;        // -m (classify.cpp:45)
static bool Quick_mode = false, Print_classified = false, Print_unclassified = false, Print_kraken = true;
static bool Populate_memory = false, Only_classified_kraken_output = false, Print_sequence = false, Map_UIDs = false;
static uint64_t Populate_memory_size = 0;
static string Classified_output_file, Unclassified_output_file, Kraken_output_file, Report_output_file, TaxDB_file;
static size_t Work_unit_size = 500000;      // DEF_WORK_UNIT_SIZE, classify.cpp:38
static int Num_threads = 1;
static int HLL_PRECISION = 14;              // readcounts.hpp:29 (only selects report columns, classify.cpp:289)
static unsigned long long total_classified = 0, total_sequences = 0, total_bases = 0;

[[noreturn]] static void die(int code, const string &msg) {
  cerr << "classify: " << msg << endl;
  exit(code);
}

static void usage(int exit_code = EX_USAGE) {
  cerr << "Usage: classify [options] <fasta/fastq file(s)>" << endl
       << endl
       << "Options: (*mandatory)" << endl
       << "* -d filename      Kraken DB filename" << endl
       << "* -i filename      Kraken DB index filename" << endl
       << "  -o filename      Output file for Kraken output" << endl
       << "  -r filename      Output file for Kraken report output" << endl
       << "  -a filename      TaxDB" << endl
       << "  -I filename      UID to TaxId map" << endl
       << "  -p #             Precision for unique k-mer counting, between 10 and 18" << endl
       << "  -t #             Number of threads" << endl
       << "  -u #             Thread work unit size (in bp)" << endl
       << "  -q               Quick operation" << endl
       << "  -m #             Minimum hit count (ignored w/o -q)" << endl
       << "  -C filename      Print classified sequences" << endl
       << "  -U filename      Print unclassified sequences" << endl
       << "  -c               Only include classified reads in output" << endl
       << "  -M               Preload database files" << endl
       << "  -x size          Preload database files using x amount of RAM (e.g. 10G)" << endl
       << "  -s               Print read sequence in Kraken output" << endl
       << "  -h               Print this message" << endl
       << endl
       << "Kraken output is to standard output by default." << endl;
  exit(exit_code);
}

// parse_human_readable_size, krakenutil.cpp:30-55
static uint64_t parse_human_readable_size(const char *s) {
  char *endp = NULL;
  errno = 0;
  uint64_t x = strtoumax(s, &endp, 10);
  if (errno || endp == s) return 0;
  int sh;
  switch (*endp) {
    case 'k': case 'K': sh = 10; break;
    case 'm': case 'M': sh = 20; break;
    case 'g': case 'G': sh = 30; break;
    case 0: sh = 0; break;
    default: return 0;
  }
  if (x > SIZE_MAX >> sh) return 0;
  return x << sh;
}

static void parse_command_line(int argc, char **argv) {
  int opt;
  long long sig;
  if (argc > 1 && strcmp(argv[1], "-h") == 0) usage(0);
  while ((opt = getopt(argc, argv, "d:i:t:u:n:m:o:qcC:U:Ma:r:sI:p:x:")) != -1) {
    switch (opt) {
      case 'd': DB_filenames.push_back(optarg); break;
      case 'i': Index_filenames.push_back(optarg); break;
      case 't':
        sig = atoll(optarg);
        if (sig <= 0) die(EX_USAGE, "can't use nonpositive thread count");
        if (sig > omp_get_num_procs()) die(EX_USAGE, "thread count exceeds number of processors");   // classify.cpp:1087-1088
        Num_threads = (int)sig;                  // host threads: parallel FASTA/FASTQ parsing and text formatting
        break;
      case 'p': HLL_PRECISION = atoi(optarg); break;
      case 'q': Quick_mode = true; break;
      case 'm':
        sig = atoll(optarg);
        if (sig <= 0) die(EX_USAGE, "can't use nonpositive minimum hit count");
        Minimum_hit_count = (uint32_t)sig;
        break;
      case 'c': Only_classified_kraken_output = true; break;
      case 'C': Print_classified = true; Classified_output_file = optarg; break;
      case 'U': Print_unclassified = true; Unclassified_output_file = optarg; break;
      case 'o': Kraken_output_file = optarg; break;
      case 'r': Report_output_file = optarg; break;
      case 's': Print_sequence = true; break;
      case 'a': TaxDB_file = optarg; break;
      case 'u':
        sig = atoll(optarg);
        if (sig <= 0) die(EX_USAGE, "can't use nonpositive work unit size");
        Work_unit_size = sig;
        break;
      case 'M': Populate_memory = true; break;
      case 'x': Populate_memory = true; Populate_memory_size = parse_human_readable_size(optarg); break;
      case 'I': Map_UIDs = true; break;
      default: usage(); break;
    }
  }
  if (DB_filenames.empty()) { cerr << "Missing mandatory option -d" << endl; usage(); }
  if (Index_filenames.empty()) { cerr << "Missing mandatory option -i" << endl; usage(); }
  if (optind == argc && !Populate_memory) cerr << "No sequence data files specified" << endl;
}

// ---- output streams: plain file, ".gz" → gzip, "-" → stdout (cout_or_file, classify.cpp:133-148) ---------------
struct OutStream {
  FILE *f = NULL;
  gzFile gz = NULL;
  bool is_stdout = false;
  void open(const string &path, bool append = false) {
    if (path == "-") { f = stdout; is_stdout = true; return; }
    if (path.size() >= 3 && path.compare(path.size() - 3, 3, ".gz") == 0) {
      gz = gzopen(path.c_str(), "wb");
      if (!gz) die(EX_CANTCREAT, "can't open " + path);
    } else {
      f = fopen(path.c_str(), append ? "ab" : "wb");
      if (!f) die(EX_CANTCREAT, "can't open " + path);
    }
  }
  void write(const char *p, size_t n) {
    if (gz) gzwrite(gz, p, (unsigned)n);
    else if (f) fwrite(p, 1, n, f);
  }
  void write(const string &s) { write(s.data(), s.size()); }
  void close() {
    if (gz) gzclose(gz);
    else if (f && !is_stdout) fclose(f);
    else if (f) fflush(f);
    f = NULL; gz = NULL;
  }
};

// ---- input: FASTA / FASTQ through zlib (plain or gzip), semantics of src/seqreader.cpp -------------------------
struct Read {
  string id, header_line, quals;
  uint64_t seq_off, seq_len;          // into the batch's bases buffer
};

// bzip2 input (the reference reads it through bxzstr when built with libbz2, seqreader.hpp:24,48).  The image has the
// runtime library but not its header, so the three entry points of the high-level API are bound with dlopen.
struct Bz2Api {
  void *(*open)(const char *, const char *) = NULL;
  int (*read)(void *, void *, int) = NULL;
  void (*close)(void *) = NULL;
  bool load() {
    if (open) return true;
    void *h = NULL;
    for (const char *name : {"libbz2.so.1.0", "libbz2.so.1", "libbz2.so"}) if ((h = dlopen(name, RTLD_NOW))) break;
    if (!h) return false;
    open = (void *(*)(const char *, const char *))dlsym(h, "BZ2_bzopen");
    read = (int (*)(void *, void *, int))dlsym(h, "BZ2_bzread");
    close = (void (*)(void *))dlsym(h, "BZ2_bzclose");
    return open && read && close;
  }
};
static Bz2Api Bz2;

static bool is_bzip2(const unsigned char *magic3) { return magic3[0] == 'B' && magic3[1] == 'Z' && magic3[2] == 'h'; }

struct LineReader {     // std::ifstream + std::getline state semantics (eofbit / failbit), over zlib / libbz2
  gzFile gz = NULL;
  void *bz = NULL;
  vector<char> buf;
  size_t pos = 0, end = 0;
  bool eofbit = false, failbit = false;
  bool open(const char *path) {
    unsigned char magic[3] = {0, 0, 0};
    if (FILE *f = fopen(path, "rb")) { size_t got = fread(magic, 1, 3, f); (void)got; fclose(f); }
    buf.resize(8 << 20);
    if (is_bzip2(magic)) {
      if (!Bz2.load()) die(EX_UNAVAILABLE, string("bzip2 input needs libbz2.so.1.0, which could not be loaded: ") + path);
      bz = Bz2.open(path, "rb");
      return bz != NULL;
    }
    gz = gzopen(path, "rb");
    if (!gz) return false;
    gzbuffer(gz, 1 << 20);
    return true;
  }
  void close() {
    if (gz) gzclose(gz);
    if (bz) Bz2.close(bz);
    gz = NULL; bz = NULL;
  }
  bool fill() {           // false at end of data
    int n = bz ? Bz2.read(bz, buf.data(), (int)buf.size()) : gzread(gz, buf.data(), (unsigned)buf.size());
    if (n <= 0) return false;
    pos = 0; end = (size_t)n;
    return true;
  }
  bool good() const { return !eofbit && !failbit; }
  // std::getline: extracts up to '\n'; hitting the end of the data sets eofbit, extracting nothing sets failbit
  bool getline(string &line) {
    line.clear();
    if (!good()) { failbit = true; return false; }
    bool extracted = false;
    for (;;) {
      if (pos == end && !fill()) {
        eofbit = true;
        if (!extracted) failbit = true;
        return extracted;
      }
      char *nl = (char *)memchr(buf.data() + pos, '\n', end - pos);
      if (nl) {
        line.append(buf.data() + pos, nl - (buf.data() + pos));
        pos = (nl - buf.data()) + 1;
        return true;
      }
      line.append(buf.data() + pos, end - pos);
      extracted = extracted || end > pos;
      pos = end;
    }
  }
  int peek() {
    if (pos == end && !fill()) return -1;
    return (unsigned char)buf[pos];
  }
};

static string first_token(const string &s) {            // istringstream >> id
  size_t a = 0;
  while (a < s.size() && isspace((unsigned char)s[a])) a++;
  size_t b = a;
  while (b < s.size() && !isspace((unsigned char)s[b])) b++;
  return s.substr(a, b - a);
}

struct SeqReader {
  LineReader in;
  bool fastq = false, valid = true;
  string linebuffer;
  bool have_linebuffer = false;
  // next_sequence(): FASTA seqreader.cpp:34-79, FASTQ :93-129.  Appends the sequence to `bases`.
  bool next(Read &r, string &bases) {
    r.quals.clear();
    if (fastq) {
      if (!valid || !in.good()) { valid = false; return false; }
      string line;
      in.getline(line);
      if (line.empty()) { valid = false; return false; }            // :101-104
      if (line[0] != '@') {
        if (line[0] != '\r') fprintf(stderr, "classify: malformed fastq file - sequence header (%s)\n", line.c_str());
        valid = false;
        return false;
      }
      r.header_line = line.substr(1);
      r.id = first_token(r.header_line);
      string seq;
      in.getline(seq);
      in.getline(line);
      if (line.empty() || line[0] != '+') {
        if (line.empty() || line[0] != '\r') fprintf(stderr, "classify: malformed fastq file - quality header (%s)\n", line.c_str());
        valid = false;
        return false;
      }
      in.getline(r.quals);
      r.seq_off = bases.size();
      r.seq_len = seq.size();
      bases += seq;
      return true;
    }
    if (!in.good()) { valid = false; return false; }                 // seqreader.cpp:37-40
    string line;
    if (have_linebuffer) { line = linebuffer; have_linebuffer = false; }
    else in.getline(line);
    if (line.empty() || line[0] != '>') {
      fprintf(stderr, "classify: malformed fasta file - expected header char > not found\n");
      valid = false;
      return false;
    }
    r.header_line = line.substr(1);
    r.id = first_token(r.header_line);
    r.seq_off = bases.size();
    while (in.good()) {
      in.getline(line);
      if (!line.empty() && line[0] == '>') { linebuffer = line; have_linebuffer = true; break; }
      bases += line;
    }
    r.seq_len = bases.size() - r.seq_off;
    return true;
  }
};

// ---- taxonomy (taxdb.hpp:563-605) + report (taxdb.hpp:928-1123) ------------------------------------------------
struct TaxEntry {
  uint32_t id = 0;
  int parent = -1;                    // index into entries, -1 = NULL
  string name, rank;
  vector<int> children;
  uint64_t genomeSize = 0, genomeSizeOfChildren = 0;
};
struct TaxDB {
  vector<TaxEntry> e;
  unordered_map<uint32_t, int> idx;
  void read(const string &path) {
    ifstream in(path);
    if (!in.is_open()) die(EX_NOINPUT, "unable to open taxonomy index file " + path);
    vector<uint32_t> parents;
    string line;
    while (getline(in, line)) {
      if (line.empty()) continue;
      size_t a = line.find('\t'), b = a == string::npos ? a : line.find('\t', a + 1);
      if (a == string::npos || b == string::npos) continue;
      size_t c = line.find('\t', b + 1);
      TaxEntry t;
      t.id = (uint32_t)strtoul(line.substr(0, a).c_str(), NULL, 10);
      uint32_t par = (uint32_t)strtoul(line.substr(a + 1, b - a - 1).c_str(), NULL, 10);
      if (t.id > 1 && t.id == par) die(1, "ERROR: the parent of " + to_string(t.id) + " is itself. Should not happend for taxa other than the root.");
      t.name = line.substr(b + 1, c == string::npos ? string::npos : c - b - 1);
      t.rank = c == string::npos ? "" : line.substr(c + 1);
      if (idx.count(t.id)) continue;                                // entries.insert keeps the first
      idx[t.id] = (int)e.size();
      e.push_back(t);
      parents.push_back(par);
    }
    if (!idx.count(0)) {                                            // taxdb.hpp:596
      TaxEntry z; z.id = 0; z.name = "unclassified"; z.rank = "no rank";
      idx[0] = (int)e.size(); e.push_back(z); parents.push_back(0);
    }
    for (size_t i = 0; i < e.size(); i++) {                          // createPointers, :411-433
      if (e[i].id == parents[i]) continue;
      auto it = idx.find(parents[i]);
      if (it == idx.end()) continue;
      e[i].parent = it->second;
      e[it->second].children.push_back((int)i);
    }
  }
  // getParentMap, :383-398
  void parent_map(vector<uint32_t> &ids, vector<uint32_t> &parents) const {
    for (auto &t : e) {
      if (t.id == 0) continue;
      ids.push_back(t.id);
      parents.push_back(t.parent < 0 ? 0 : e[t.parent].id);
    }
  }
  void set_genome_size(uint32_t taxid, uint64_t size) {               // :850-866
    auto it = idx.find(taxid);
    if (it == idx.end()) { cerr << "No taxonomy entry for " << taxid << "!!" << endl; return; }
    int i = it->second;
    e[i].genomeSize += size;
    while (e[i].parent >= 0) { i = e[i].parent; e[i].genomeSizeOfChildren += size; }
  }
};

struct Clade { uint64_t reads = 0, kmers = 0, unique = 0; };

static void print_report(kuq_ctx *ctx, TaxDB &tax, ostream &out) {
  uint32_t n = 0;
  if (kuq_counts_size(ctx, &n)) die(EX_SOFTWARE, kuq_last_error(ctx));
  vector<uint32_t> t(n);
  vector<uint64_t> nr(n), nk(n), uq(n);
  if (n && kuq_read_counts(ctx, t.data(), nr.data(), nk.data(), uq.data(), NULL, n)) die(EX_SOFTWARE, kuq_last_error(ctx));
  unordered_map<uint32_t, size_t> row;
  for (uint32_t i = 0; i < n; i++) row[t[i]] = i;
  // TaxReport ctor (:928-982): every counted taxon with an entry is a member of each ancestor's clade
  cerr << "Setting values in the taxonomy tree ...";
  map<int, vector<uint32_t>> members;
  for (uint32_t i = 0; i < n; i++) {
    auto it = tax.idx.find(t[i]);
    if (it == tax.idx.end()) { cerr << "No entry for " << t[i] << " in database!" << endl; continue; }
    for (int x = it->second; x >= 0; x = tax.e[x].parent) members[x].push_back(t[i]);
  }
  unordered_map<int, Clade> clade;
  // A handful of clades: one merge per clade.  A real taxonomy (thousands of rows): all clades in one call — the
  // library walks its own copy of the tree, which is this one as long as nothing hangs below the "unclassified" entry.
  size_t batch_min = 32;
  if (const char *v = getenv("KUQ_REPORT_BATCH_MIN")) batch_min = (size_t)strtoull(v, NULL, 10);
  bool batch = getenv("KUQ_REPORT_PER_CLADE") == NULL && members.size() >= batch_min;
  {
    auto z = tax.idx.find(0);
    if (z != tax.idx.end() && !tax.e[z->second].children.empty()) batch = false;
  }
  if (batch) {
    vector<uint32_t> ids;
    vector<int> node;
    for (auto &kv : members) { ids.push_back(tax.e[kv.first].id); node.push_back(kv.first); }
    vector<uint64_t> r(ids.size()), k(ids.size()), u(ids.size());
    if (kuq_clade_counts_tree(ctx, ids.data(), (uint32_t)ids.size(), r.data(), k.data(), u.data())) die(EX_SOFTWARE, kuq_last_error(ctx));
    for (size_t i = 0; i < ids.size(); i++) { Clade c; c.reads = r[i]; c.kmers = k[i]; c.unique = u[i]; clade[node[i]] = c; }
  } else {
    for (auto &kv : members) {
      Clade c;
      if (kuq_clade_counts(ctx, kv.second.data(), (uint32_t)kv.second.size(), &c.reads, &c.kmers, &c.unique))
        die(EX_SOFTWARE, kuq_last_error(ctx));
      clade[kv.first] = c;
    }
  }
  cerr << " done" << endl;
  cerr << "Printing classification report ... ";
  uint64_t total = 0;
  for (uint32_t root : {0u, 1u, 0xFFFFFFFFu}) {
    auto it = tax.idx.find(root);
    if (it != tax.idx.end() && clade.count(it->second)) total += clade[it->second].reads;
  }
  if (total == 0) { cerr << "total number of reads is zero - not creating a report!" << endl; return; }
  const bool hll_cols = HLL_PRECISION > 0;
  out << (hll_cols ? "%\treads\ttaxReads\tkmers\tdup\tcov\ttaxID\trank\ttaxName\n" : "%\treads\ttaxReads\ttaxID\trank\ttaxName\n");
  // printReport(tax, depth), :1040-1076 (explicit stack instead of recursion)
  struct Item { int node; unsigned depth; };
  for (uint32_t root : {0u, 1u, 0xFFFFFFFFu}) {
    auto rit = tax.idx.find(root);
    if (rit == tax.idx.end()) continue;
    vector<Item> stack{{rit->second, 0}};
    while (!stack.empty()) {
      Item it = stack.back();
      stack.pop_back();
      auto cit = clade.find(it.node);
      if (cit == clade.end() || cit->second.reads == 0) continue;
      const TaxEntry &e = tax.e[it.node];
      const Clade &c = cit->second;
      auto r = row.find(e.id);
      // printLine, :1078-1123
      out << setprecision(4) << 100.0 * c.reads / total << '\t' << c.reads << '\t' << (r != row.end() ? nr[r->second] : 0) << '\t';
      if (hll_cols) {
        double genome_size = double(e.genomeSize + e.genomeSizeOfChildren);
        out << c.unique << '\t' << setprecision(3) << (double(c.kmers) / c.unique) << '\t';
        if (genome_size == 0) out << "NA"; else out << setprecision(4) << (c.unique / genome_size);
        out << '\t';
      }
      out << (e.id == 0xFFFFFFFFu ? -1 : (int32_t)e.id) << '\t' << e.rank << '\t' << string(2 * it.depth, ' ') + e.name << '\n';
      // children that have a clade, sorted descending by (reads, kmers) — ReadCounts::operator<, readcounts.hpp:90-98
      vector<int> ch;
      for (int k : e.children) if (clade.count(k)) ch.push_back(k);
      stable_sort(ch.begin(), ch.end(), [&](int a, int b) {
        const Clade &x = clade[a], &y = clade[b];
        return y.reads < x.reads || (y.reads == x.reads && y.kmers < x.kmers);
      });
      for (auto k = ch.rbegin(); k != ch.rend(); ++k) stack.push_back({*k, it.depth + 1});
    }
  }
  cerr << " done" << endl;
}

// ---- mmap helper ----------------------------------------------------------------------------------------------
struct Mapped {
  void *p = NULL;
  size_t size = 0;
  void open(const string &path) {
    int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) die(EX_OSERR, "unable to open " + path);
    struct stat sb;
    if (fstat(fd, &sb) < 0) die(EX_OSERR, "unable to fstat " + path);
    size = sb.st_size;
    p = mmap(0, size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (p == MAP_FAILED) die(EX_OSERR, "unable to mmap " + path);
    madvise(p, size, MADV_SEQUENTIAL);
    ::close(fd);
  }
};

static double now_s() { struct timeval t; gettimeofday(&t, NULL); return t.tv_sec + t.tv_usec / 1e6; }
static const bool Timing = getenv("KUQ_TIMING") != NULL;
#define TICK(label) do { if (Timing) { double t__ = now_s(); fprintf(stderr, "\n[timing] %-28s %.3f s\n", label, t__ - T0__); T0__ = t__; } } while (0)
static double get_seconds(struct timeval a, struct timeval b) {
  return (b.tv_sec - a.tv_sec) + (b.tv_usec - a.tv_usec) / 1e6;
}

// ---- the work-unit loop ---------------------------------------------------------------------------------------
struct Batch {
  string bases;
  vector<uint64_t> offs{0};
  vector<Read> reads;
  vector<uint32_t> codes;             // chunked mode: per-window dense ids merged over the database ranges
  bool fastq = false;
  int file = 0;                       // index of the input file (work units do not span files)
  void clear() { bases.clear(); offs.assign(1, 0); reads.clear(); codes.clear(); }
};

static OutStream Kraken_out, Classified_out, Unclassified_out;
static bool Fastq_input = false;

static void print_sequence(OutStream &o, const Read &r, const string &bases) {   // classify.cpp:794-805
  string s;
  if (Fastq_input) {
    s = "@" + r.header_line + "\n";
    s.append(bases, r.seq_off, r.seq_len);
    s += "\n+\n" + r.quals + "\n";
  } else {
    s = ">" + r.header_line + "\n";
    s.append(bases, r.seq_off, r.seq_len);
    s += "\n";
  }
  o.write(s);
}

// reads [first, first + n) of the batch; the result arrays are indexed from 0 (a share of the batch resolved on one GPU)
static void emit_results(const Batch &b, const kuq_batch_result &res, size_t first = 0, size_t n = (size_t)-1) {
  string out;
  if (n == (size_t)-1) n = b.reads.size() - first;
  out.reserve(n * 64);
  char num[32];
  for (size_t i = 0; i < n; i++) {
    const Read &r = b.reads[first + i];
    uint32_t call = res.call[i];
    if (Print_unclassified && !call) print_sequence(Unclassified_out, r, b.bases);
    if (Print_classified && call) print_sequence(Classified_out, r, b.bases);
    if (!Print_kraken) continue;
    if (!call && Only_classified_kraken_output) continue;           // :986-988
    out += call ? "C\t" : "U\t";
    out += r.id;
    out += '\t';
    snprintf(num, sizeof num, "%u", call); out += num;
    out += '\t';
    snprintf(num, sizeof num, "%" PRIu64, r.seq_len); out += num;
    out += '\t';
    if (Quick_mode) { snprintf(num, sizeof num, "Q:%u", res.run_count[i]); out += num; }   // :989-990
    else if (res.run_count[i] == 0) out += "0:0";                    // :994-995
    for (uint32_t j = 0; !Quick_mode && j < res.run_count[i]; j++) {   // hitlist_string, :826-861
      const kuq_run &run = res.runs[res.run_start[i] + j];
      if (j) out += ' ';
      if (run.code == KUQ_CODE_AMBIG) snprintf(num, sizeof num, "A:%u", run.count);
      else snprintf(num, sizeof num, "%u:%u", run.code, run.count);
      out += num;
    }
    if (Print_sequence) { out += '\t'; out.append(b.bases, r.seq_off, r.seq_len); }
    out += '\n';
  }
  if (Print_kraken) Kraken_out.write(out);
}

static inline void append_u32(string &o, uint32_t v) {       // decimal without snprintf
  char buf[10];
  int n = 0;
  do { buf[n++] = (char)('0' + v % 10); v /= 10; } while (v);
  while (n) o += buf[--n];
}

// background writer: formatted text is handed over batch by batch and written in order while the next batch is
// being gathered / classified / formatted
struct OutJob { vector<string> kraken, classified, unclassified; };
struct Writer {
  std::mutex m;
  std::condition_variable cv;
  std::deque<OutJob> q;
  std::deque<OutJob> spare;              // written jobs, recycled so that their string buffers stay mapped
  bool done = false;
  std::thread th;
  void start() { th = std::thread([this] { run(); }); }
  void push(OutJob &&j) { { std::lock_guard<std::mutex> l(m); q.push_back(std::move(j)); } cv.notify_one(); }
  OutJob take(size_t parts) {
    OutJob j;
    { std::lock_guard<std::mutex> l(m); if (!spare.empty()) { j = std::move(spare.front()); spare.pop_front(); } }
    j.kraken.resize(parts); j.classified.resize(parts); j.unclassified.resize(parts);
    for (auto &x : j.kraken) x.clear();
    for (auto &x : j.classified) x.clear();
    for (auto &x : j.unclassified) x.clear();
    return j;
  }
  void finish() { { std::lock_guard<std::mutex> l(m); done = true; } cv.notify_one(); if (th.joinable()) th.join(); }
  void run();
};

void Writer::run() {
  for (;;) {
    OutJob j;
    {
      std::unique_lock<std::mutex> l(m);
      cv.wait(l, [this] { return done || !q.empty(); });
      if (q.empty()) return;
      j = std::move(q.front());
      q.pop_front();
    }
    for (auto &x : j.kraken) Kraken_out.write(x);
    for (auto &x : j.classified) Classified_out.write(x);
    for (auto &x : j.unclassified) Unclassified_out.write(x);
    { std::lock_guard<std::mutex> l(m); if (spare.size() < 3) spare.push_back(std::move(j)); }
  }
}

// ---- parallel ingest for plain (uncompressed) files --------------------------------------------------------------
// The file is mmap'ed and cut into chunks at record boundaries; every chunk is parsed by its own thread into
// zero-copy views (FASTQ) with exactly the record semantics of src/seqreader.cpp; a sequential pass over the read
// lengths then cuts work units and batches like process_file (classify.cpp:506-521); sequence text is gathered into
// pinned buffers and the Kraken lines are formatted by all threads.
struct View {
  const char *hdr; uint32_t hdr_len;      // header line without the leading '>' / '@'
  const char *seq; uint32_t seq_len;      // FASTQ / single-line FASTA: points into the file
  const char *qual; uint32_t qual_len;
  int64_t owned;                          // >= 0: index into the chunk's arena (multi-line FASTA)
};
struct ChunkParse {
  vector<View> reads;
  vector<string> arena;
  bool stopped = false;                   // parser hit the end-of-input condition of the reference reader
  string warning;
};

static inline const char *line_end(const char *p, const char *end) {
  const char *nl = (const char *)memchr(p, '\n', end - p);
  return nl ? nl : end;
}

// FASTQ records [p, end): FastqReader::next_sequence, seqreader.cpp:93-129
static void parse_fastq_chunk(const char *p, const char *end, const char *file_end, ChunkParse &out) {
  while (p < end) {
    const char *e0 = line_end(p, file_end);
    if (e0 == p) { out.stopped = true; return; }                                  // empty line (:101-104)
    if (*p != '@') {
      if (*p != '\r') out.warning = "malformed fastq file - sequence header (" + string(p, e0 - p) + ")";
      out.stopped = true;
      return;
    }
    const char *s = e0 < file_end ? e0 + 1 : file_end;
    const char *e1 = line_end(s, file_end);
    const char *pl = e1 < file_end ? e1 + 1 : file_end;
    const char *e2 = line_end(pl, file_end);
    if (e2 == pl || *pl != '+') {
      if (e2 == pl || *pl != '\r') out.warning = "malformed fastq file - quality header (" + string(pl, e2 - pl) + ")";
      out.stopped = true;
      return;
    }
    const char *q = e2 < file_end ? e2 + 1 : file_end;
    const char *e3 = line_end(q, file_end);
    View v;
    v.hdr = p + 1; v.hdr_len = (uint32_t)(e0 - p - 1);
    v.seq = s; v.seq_len = (uint32_t)(e1 - s);
    v.qual = q; v.qual_len = (uint32_t)(e3 - q);
    v.owned = -1;
    out.reads.push_back(v);
    p = e3 < file_end ? e3 + 1 : file_end;
  }
}

// FASTA records [p, end): FastaReader::next_sequence, seqreader.cpp:34-79.  `p` is at a '>' line start.
static void parse_fasta_chunk(const char *p, const char *end, const char *file_end, ChunkParse &out) {
  while (p < end) {
    const char *e0 = line_end(p, file_end);
    if (*p != '>') { out.warning = "malformed fasta file - expected header char > not found"; out.stopped = true; return; }
    // a last header line that is not newline-terminated sets eofbit while being read into the line buffer, and the
    // reader then reports !good() before using it (seqreader.cpp:37-40,62-66): that record is never produced
    if (e0 == file_end) { out.stopped = true; return; }
    View v;
    v.hdr = p + 1; v.hdr_len = (uint32_t)(e0 - p - 1);
    v.qual = NULL; v.qual_len = 0; v.owned = -1;
    const char *s = e0 < file_end ? e0 + 1 : file_end;
    // sequence lines until the next line that starts with '>'
    const char *first = s, *first_end = line_end(s, file_end);
    const char *nxt = first_end < file_end ? first_end + 1 : file_end;
    if (s >= file_end) { v.seq = s; v.seq_len = 0; out.reads.push_back(v); return; }
    if (*s == '>') { v.seq = s; v.seq_len = 0; out.reads.push_back(v); p = s; continue; }
    if (nxt >= file_end || *nxt == '>') {                                          // single-line sequence: zero copy
      v.seq = first; v.seq_len = (uint32_t)(first_end - first);
      out.reads.push_back(v);
      p = nxt;
      continue;
    }
    string acc(first, first_end - first);
    const char *q = nxt;
    while (q < file_end && *q != '>') {
      const char *qe = line_end(q, file_end);
      acc.append(q, qe - q);
      q = qe < file_end ? qe + 1 : file_end;
    }
    v.owned = (int64_t)out.arena.size();
    v.seq = NULL; v.seq_len = (uint32_t)acc.size();
    out.arena.push_back(std::move(acc));
    out.reads.push_back(v);
    p = q;
  }
}

// start of the first record at or after `p` (chunk boundary search)
static const char *next_record_start(const char *p, const char *file_begin, const char *file_end, bool fastq) {
  if (p <= file_begin) return file_begin;
  // move to the start of the next line
  const char *nl = (const char *)memchr(p - 1, '\n', file_end - (p - 1));
  if (!nl) return file_end;
  p = nl + 1;
  while (p < file_end) {
    const char *e0 = line_end(p, file_end);
    if (!fastq) { if (*p == '>') return p; }
    else if (*p == '@') {
      // a header is followed two lines later by the '+' line; a quality line that starts with '@' is not
      const char *l1 = e0 < file_end ? e0 + 1 : file_end;
      const char *e1 = line_end(l1, file_end);
      const char *l2 = e1 < file_end ? e1 + 1 : file_end;
      if (l2 < file_end && *l2 == '+' && l1 < file_end && *l1 != '@') return p;
    }
    p = e0 < file_end ? e0 + 1 : file_end;
  }
  return file_end;
}

struct Stage { char *bases = NULL; uint64_t cap = 0; uint64_t *offs = NULL; size_t offs_cap = 0; size_t begin = 0, end = 0; };
// pinned staging of one batch; prepare_stage() is also called up front, while the database is being staged
static vector<Stage> Prealloc_stages;
static void prepare_stage(Stage &s, uint64_t bases, size_t reads) {
  if (s.cap < bases + 64) {
    if (s.bases) kuq_host_free(s.bases);
    s.cap = std::max<uint64_t>(bases + 64, 160ull << 20);
    s.bases = (char *)kuq_host_alloc(s.cap);
    if (!s.bases) die(EX_OSERR, "pinned allocation failed");
  }
  if (s.offs_cap < reads + 2) {
    if (s.offs) kuq_host_free(s.offs);
    s.offs_cap = std::max<size_t>(reads + 2, (1u << 20) + 8);
    s.offs = (uint64_t *)kuq_host_alloc(s.offs_cap * 8);
    if (!s.offs) die(EX_OSERR, "pinned allocation failed");
  }
}

// Several GPUs (replicas: every device holds the database): batches go round-robin over the contexts; results come back
// in submission order, so the Kraken output stays in file order whatever the number of devices.
static vector<kuq_ctx *> All_ctx;
static bool process_file_parallel(kuq_ctx *ctx0, const char *filename) {
  const vector<kuq_ctx *> ctxs = All_ctx.empty() ? vector<kuq_ctx *>{ctx0} : All_ctx;
  const int G = (int)ctxs.size();
  int fd = ::open(filename, O_RDONLY);
  if (fd < 0) return false;
  struct stat sb;
  if (fstat(fd, &sb) < 0 || !S_ISREG(sb.st_mode) || sb.st_size < 2) { ::close(fd); return false; }
  const size_t size = sb.st_size;
  const char *base = (const char *)mmap(0, size, PROT_READ, MAP_PRIVATE, fd, 0);
  ::close(fd);
  if (base == MAP_FAILED) return false;
  if (((unsigned char)base[0] == 0x1f && (unsigned char)base[1] == 0x8b) || (size >= 3 && is_bzip2((const unsigned char *)base))) {
    munmap((void *)base, size);                                     // gzip / bzip2: the serial reader decompresses
    return false;
  }
  madvise((void *)base, size, MADV_WILLNEED);
  const char *fend = base + size;
  const bool fastq = Fastq_input = base[0] == '@';                                  // determine_input_file_type
  const int T = std::max(1, Num_threads);
  const int n_chunks = (int)std::min<size_t>((size_t)T * 4, std::max<size_t>(1, size / (1 << 20)));
  vector<const char *> cut(n_chunks + 1);
  cut[0] = base; cut[n_chunks] = fend;
  for (int c = 1; c < n_chunks; c++) cut[c] = next_record_start(base + size * c / n_chunks, base, fend, fastq);
  for (int c = 1; c <= n_chunks; c++) if (cut[c] < cut[c - 1]) cut[c] = cut[c - 1];
  double T0__ = now_s();
  vector<ChunkParse> parsed(n_chunks);
#pragma omp parallel for schedule(dynamic, 1) num_threads(T)
  for (int c = 0; c < n_chunks; c++) {
    if (cut[c] >= cut[c + 1]) continue;
    if (fastq) parse_fastq_chunk(cut[c], cut[c + 1], fend, parsed[c]);
    else parse_fasta_chunk(cut[c], cut[c + 1], fend, parsed[c]);
  }
  TICK("parse");
  // the reference stops at the first record its reader rejects
  size_t n_total = 0;
  int last_chunk = n_chunks;
  for (int c = 0; c < n_chunks; c++) {
    n_total += parsed[c].reads.size();
    if (parsed[c].stopped) {
      if (!parsed[c].warning.empty()) fprintf(stderr, "classify: %s\n", parsed[c].warning.c_str());
      last_chunk = c + 1;
      break;
    }
  }
  vector<size_t> first_of(last_chunk + 1, 0);
  for (int c = 0; c < last_chunk; c++) first_of[c + 1] = first_of[c] + parsed[c].reads.size();
  n_total = first_of[last_chunk];
  // flat views of all reads + cumulative sequence length (parallel per chunk, chunk bases from a short serial scan)
  vector<const View *> rv(n_total);
  vector<uint64_t> cum(n_total + 1);
  {
    vector<uint64_t> chunk_nt(last_chunk + 1, 0);
#pragma omp parallel for schedule(dynamic, 1) num_threads(T)
    for (int c = 0; c < last_chunk; c++) {
      uint64_t nt = 0;
      for (const View &v : parsed[c].reads) nt += v.seq_len;
      chunk_nt[c + 1] = nt;
    }
    for (int c = 0; c < last_chunk; c++) chunk_nt[c + 1] += chunk_nt[c];
#pragma omp parallel for schedule(dynamic, 1) num_threads(T)
    for (int c = 0; c < last_chunk; c++) {
      uint64_t run = chunk_nt[c];
      size_t i = first_of[c];
      for (const View &v : parsed[c].reads) { rv[i] = &v; cum[i] = run; run += v.seq_len; i++; }
    }
    cum[n_total] = chunk_nt[last_chunk];
  }
  auto chunk_of = [&](size_t i) { return (int)(std::upper_bound(first_of.begin(), first_of.end(), i) - first_of.begin()) - 1; };
  // work units and batches (classify.cpp:506-521): batches end at unit boundaries
  struct Cut { size_t begin, end; uint64_t bases; };
  vector<Cut> batches;
  {
    const uint64_t BATCH_NT = 96ull << 20;
    // KUQ_BATCH_READS: close a batch at the first unit boundary after this many reads (tests: small batches)
    const size_t batch_reads = getenv("KUQ_BATCH_READS") ? std::max<size_t>(1, strtoull(getenv("KUQ_BATCH_READS"), NULL, 10)) : (1u << 20) - 4096;
    uint64_t unit_nt = 0, batch_nt = 0;
    size_t begin = 0, unit_first = 0;
    const uint64_t SLOT_NT = 280ull << 20;            // below the slot capacity (288 MiB)
    for (size_t i = 0; i < n_total; i++) {
      const uint64_t len = cum[i + 1] - cum[i];
      if (len > SLOT_NT) die(EX_DATAERR, "a single sequence is longer than 280 Mbp: not supported");
      if (batch_nt + len > SLOT_NT && i > begin) {    // a very long sequence is coming: close the batch before it
        batches.push_back({begin, i, batch_nt});
        begin = i; batch_nt = 0;
      }
      unit_nt += len; batch_nt += len;
      bool close_unit = unit_nt >= Work_unit_size;
      if (close_unit) { unit_nt = 0; unit_first = i + 1; }
      if ((close_unit && (batch_nt >= BATCH_NT || i + 1 - begin >= std::min<size_t>(batch_reads, (1u << 20) - 4096))) ||
          (!close_unit && (i + 1 - begin >= (1u << 20) - 1 || batch_nt >= (150ull << 20)))) {
        batches.push_back({begin, i + 1, batch_nt});
        begin = i + 1; batch_nt = 0;
        if (!close_unit) unit_first = i + 1;
      }
    }
    size_t end = n_total;
    if (unit_nt == 0 && unit_first < n_total) end = unit_first;    // last unit with zero length is dropped (:523-524)
    if (end > begin) batches.push_back({begin, end, cum[end] - cum[begin]});
  }
  TICK("units+batches");
  // ---- pipeline: a filler thread gathers batch b + 2 into pinned memory and submits it while the GPUs classify batch
  // b + 1 and this thread formats batch b.  Batch b runs on device b % G in slot (b / G) % NS and is staged in
  // st[b % (NS * G)]; a slot (and its staging set) is reused only after its batch has been formatted, because the
  // results live in the slot's pinned buffers until then.
  const int NS = 3;
  // Worker teams of the pipeline: more than ~32 OpenMP threads spinning at their barriers starve the thread that talks
  // to the CUDA driver (measured with -t 64 on 8 M reads: 1.47 s; OMP_WAIT_POLICY=passive 0.58 s; -t 32 0.55 s; -t 16
  // 0.56 s — profiles/cli_variants_r02.log), and the formatting is not the bottleneck beyond that anyway.
  const int Tf = std::min(8, std::max(1, T / 4)), Te = std::min(24, std::max(1, T - Tf));
  // staging sets live for the whole process: pinned allocations cost tens of ms each, the second input file reuses them
  static vector<Stage> st;
  if (st.empty() && !Prealloc_stages.empty()) st.swap(Prealloc_stages);
  if (st.size() < (size_t)NS * G) st.resize((size_t)NS * G);
  double t_fill = 0, t_emit = 0, t_wait = 0, t_submit = 0;
  auto fill = [&](Stage &s, const Cut &b) {
    s.begin = b.begin; s.end = b.end;
    const size_t n = b.end - b.begin;
    prepare_stage(s, b.bases, n);
    const uint64_t c0 = cum[b.begin];
#pragma omp parallel for schedule(static) num_threads(Tf)
    for (size_t i = 0; i < n; i++) {
      const View &v = *rv[b.begin + i];
      s.offs[i] = cum[b.begin + i] - c0;
      const char *src = v.owned >= 0 ? parsed[chunk_of(b.begin + i)].arena[v.owned].data() : v.seq;
      memcpy(s.bases + s.offs[i], src, v.seq_len);
    }
    s.offs[n] = cum[b.end] - c0;
  };
  Writer writer;
  writer.start();
  // upper bound of the Kraken line of read i (so that a part formats into one exactly sized buffer, no growth checks)
  auto emit = [&](Stage &s, const kuq_batch_result &res) {
    const size_t n = s.end - s.begin;
    const int parts = Te;
    OutJob job = writer.take(parts);
    vector<string> &kr = job.kraken, &cl = job.classified, &un = job.unclassified;
#pragma omp parallel for schedule(static, 1) num_threads(Te)
    for (int pi = 0; pi < parts; pi++) {
      const size_t a = n * pi / parts, b = n * (pi + 1) / parts;
      string &out = kr[pi];
      if (Print_kraken) {
        size_t need = 0;
        for (size_t i = a; i < b; i++) {
          const View &v = *rv[s.begin + i];
          need += 2 + v.hdr_len + 1 + 10 + 1 + 10 + 1 + 4 + 22ull * res.run_count[i] + 1 + (Print_sequence ? v.seq_len + 1 : 0);
        }
        out.resize(need);
      }
      char *o = Print_kraken ? &out[0] : NULL;
      auto put_u32 = [&](uint32_t v) {
        char buf[10];
        int k = 0;
        do { buf[k++] = (char)('0' + v % 10); v /= 10; } while (v);
        while (k) *o++ = buf[--k];
      };
      for (size_t i = a; i < b; i++) {
        const View &v = *rv[s.begin + i];
        const uint32_t call = res.call[i];
        const char *seq = s.bases + s.offs[i];
        if ((Print_unclassified && !call) || (Print_classified && call)) {       // print_sequence, :794-805
          string &oo = call ? cl[pi] : un[pi];
          oo += fastq ? '@' : '>';
          oo.append(v.hdr, v.hdr_len); oo += '\n';
          oo.append(seq, v.seq_len); oo += '\n';
          if (fastq) { oo += "+\n"; oo.append(v.qual, v.qual_len); oo += '\n'; }
        }
        if (!Print_kraken) continue;
        if (!call && Only_classified_kraken_output) continue;
        *o++ = call ? 'C' : 'U'; *o++ = '\t';
        {   // id = first whitespace-delimited token of the header line
          uint32_t x = 0;
          while (x < v.hdr_len && isspace((unsigned char)v.hdr[x])) x++;
          uint32_t y = x;
          while (y < v.hdr_len && !isspace((unsigned char)v.hdr[y])) y++;
          memcpy(o, v.hdr + x, y - x); o += y - x;
        }
        *o++ = '\t';
        put_u32(call);
        *o++ = '\t';
        put_u32(v.seq_len);
        *o++ = '\t';
        if (Quick_mode) { *o++ = 'Q'; *o++ = ':'; put_u32(res.run_count[i]); }
        else if (res.run_count[i] == 0) { *o++ = '0'; *o++ = ':'; *o++ = '0'; }
        for (uint32_t j = 0; !Quick_mode && j < res.run_count[i]; j++) {
          const kuq_run &run = res.runs[res.run_start[i] + j];
          if (j) *o++ = ' ';
          if (run.code == KUQ_CODE_AMBIG) *o++ = 'A'; else put_u32(run.code);
          *o++ = ':';
          put_u32(run.count);
        }
        if (Print_sequence) { *o++ = '\t'; memcpy(o, seq, v.seq_len); o += v.seq_len; }
        *o++ = '\n';
      }
      if (Print_kraken) out.resize(o - &out[0]);
    }
    writer.push(std::move(job));
    total_classified += res.n_classified;
    total_sequences += n;
    total_bases += s.offs[n];
    fprintf(stderr, "\r Processed %llu sequences (%.2f%% classified)", total_sequences, total_classified * 100.0 / total_sequences);
  };
  // batch bi: device bi % G, slot (bi / G) % NS, staging set bi % (NS * G)
  std::mutex pm;
  std::condition_variable pcv;
  size_t submitted = 0, collected = 0;
  string filler_error;
  std::thread filler([&] {
    for (size_t bi = 0; bi < batches.size(); bi++) {
      {
        std::unique_lock<std::mutex> l(pm);
        pcv.wait(l, [&] { return bi - collected < (size_t)(NS * G); });        // the slot's previous batch is formatted
      }
      const int k = (int)(bi % ((size_t)NS * G));
      kuq_ctx *c = ctxs[bi % G];
      double a = now_s();
      fill(st[k], batches[bi]);
      double b = now_s();
      if (kuq_submit_batch(c, (uint32_t)((bi / G) % NS), st[k].bases, st[k].offs, (uint32_t)(st[k].end - st[k].begin), NULL, 0)) {
        std::lock_guard<std::mutex> l(pm);
        filler_error = kuq_last_error(c);
        submitted = batches.size() + 1;
        pcv.notify_all();
        return;
      }
      t_fill += b - a; t_submit += now_s() - b;
      { std::lock_guard<std::mutex> l(pm); submitted = bi + 1; }
      pcv.notify_all();
    }
  });
  for (size_t bj = 0; bj < batches.size(); bj++) {
    {
      std::unique_lock<std::mutex> l(pm);
      pcv.wait(l, [&] { return submitted > bj; });
      if (!filler_error.empty()) { l.unlock(); filler.join(); die(EX_SOFTWARE, filler_error); }
    }
    const int k = (int)(bj % ((size_t)NS * G));
    kuq_ctx *c = ctxs[bj % G];
    kuq_batch_result res;
    double c0 = now_s();
    if (kuq_wait_batch(c, (uint32_t)((bj / G) % NS), &res)) { filler.detach(); die(EX_SOFTWARE, kuq_last_error(c)); }
    double d = now_s();
    emit(st[k], res);
    t_wait += d - c0; t_emit += now_s() - d;
    { std::lock_guard<std::mutex> l(pm); collected = bj + 1; }
    pcv.notify_all();
  }
  filler.join();
  TICK("pipeline");
  { double w0 = now_s(); writer.finish(); if (Timing) fprintf(stderr, "\n[timing] writer drain %.3f s\n", now_s() - w0); }
  if (Timing) fprintf(stderr, "\n[timing] fill %.3f submit %.3f (filler thread, %d threads) | wait %.3f emit %.3f s (%d threads) over %zu batches on %d device(s)\n",
                      t_fill, t_submit, Tf, t_wait, t_emit, Te, batches.size(), G);
  for (kuq_ctx *c : ctxs) if (kuq_finish(c)) die(EX_SOFTWARE, kuq_last_error(c));
  TICK("kuq_finish (harvest)");
  // tearing down a multi-GB mapping takes ~0.1 s of page-table work nobody has to wait for
  std::thread([base, size] { munmap((void *)base, size); }).detach();
  TICK("unmap (detached)");
  (void)ctx0;
  return true;
}

static void process_file(kuq_ctx *ctx, const char *filename) {
  if (!getenv("KUQ_SERIAL_INGEST") && process_file_parallel(ctx, filename)) return;   // plain files: parallel path
  SeqReader reader;
  if (!reader.in.open(filename)) die(EX_NOINPUT, string("can't open ") + filename);
  Fastq_input = reader.fastq = reader.in.peek() == '@';            // determine_input_file_type, :377-388
  const uint64_t BATCH_NT = 96ull << 20;
  Batch batches[2];
  int cur = 0, inflight = -1;
  uint64_t unit_nt = 0;
  size_t unit_first_read = 0;       // first read of the open work unit inside the current batch
  auto wait_emit = [&](int which) {
    kuq_batch_result res;
    if (kuq_wait_batch(ctx, (uint32_t)which, &res)) die(EX_SOFTWARE, kuq_last_error(ctx));
    emit_results(batches[which], res);
    total_classified += res.n_classified;
    total_sequences += batches[which].reads.size();
    total_bases += batches[which].bases.size();
    fprintf(stderr, "\r Processed %llu sequences (%.2f%% classified)", total_sequences,
            total_classified * 100.0 / total_sequences);
    batches[which].clear();
  };
  auto submit = [&](int which) {
    Batch &b = batches[which];
    if (b.reads.empty()) return;
    if (kuq_submit_batch(ctx, (uint32_t)which, b.bases.data(), b.offs.data(), (uint32_t)b.reads.size(), NULL, 0))
      die(EX_SOFTWARE, kuq_last_error(ctx));
    if (inflight >= 0) wait_emit(inflight);
    inflight = which;
  };
  Read r;
  while (reader.valid) {
    Batch &b = batches[cur];
    if (!reader.next(r, b.bases)) break;
    if (r.seq_len > (280ull << 20)) die(EX_DATAERR, "a single sequence is longer than 280 Mbp: not supported");
    if (b.bases.size() > (280ull << 20) && !b.reads.empty()) {
      // the sequence just read does not fit next to the batch: move it to the next batch
      Batch &nb = batches[cur ^ 1];
      string seq = b.bases.substr(r.seq_off);
      b.bases.resize(r.seq_off);
      submit(cur);
      cur ^= 1;
      unit_first_read = 0;
      Batch &b2 = batches[cur];
      (void)nb;
      r.seq_off = b2.bases.size();
      b2.bases += seq;
      b2.reads.push_back(r);
      b2.offs.push_back(b2.bases.size());
      unit_nt += r.seq_len;
      if (unit_nt >= Work_unit_size) { unit_nt = 0; unit_first_read = b2.reads.size(); }
      continue;
    }
    b.reads.push_back(r);
    b.offs.push_back(b.bases.size());
    unit_nt += r.seq_len;
    if (unit_nt >= Work_unit_size) {                                 // the unit closes (classify.cpp:514-520)
      unit_nt = 0;
      unit_first_read = b.reads.size();
      if (b.bases.size() >= BATCH_NT || b.reads.size() >= (1u << 20) - 4096) {
        submit(cur);
        cur ^= 1;
        unit_first_read = 0;
      }
    } else if (b.reads.size() >= (1u << 20) - 1 || b.bases.size() >= (150ull << 20)) {
      // a single work unit larger than a slot (huge -u): cut here; the unit is then counted as two
      submit(cur);
      cur ^= 1;
      unit_first_read = 0;
    }
  }
  // end of file: a last work unit whose total length is 0 is dropped by the reference (classify.cpp:523-524)
  {
    Batch &b = batches[cur];
    if (unit_nt == 0 && unit_first_read < b.reads.size()) {
      b.reads.resize(unit_first_read);
      b.offs.resize(unit_first_read + 1);
    }
    submit(cur);
  }
  if (inflight >= 0) wait_emit(inflight);
  if (kuq_finish(ctx)) die(EX_SOFTWARE, kuq_last_error(ctx));     // work units do not span input files
  reader.in.close();
}

// ---- chunked mode: `-x size` with a database that does not fit (process_file_with_db_chunk, classify.cpp:566-791) --
// Ranges of minimizer bins whose records + index slice fit the budget (prepare_chunking / upper_bound,
// krakendb.cpp:430-522).  Any byte-balanced partition gives the same results (a key lives in one bin).
static vector<pair<uint64_t, uint64_t>> plan_ranges(const uint64_t *offsets, uint64_t n_bins, uint64_t budget) {
  vector<pair<uint64_t, uint64_t>> ranges;
  uint64_t first = 0;
  while (first < n_bins) {
    uint64_t lo = first, hi = n_bins;                 // largest `end` in (first, n_bins] that still fits
    auto fits = [&](uint64_t end) { return (end - first + 1) * 8 + (offsets[end] - offsets[first]) * 12 + 8 <= budget; };
    if (!fits(first + 1)) die(EX_SOFTWARE, "a single minimizer bin exceeds the chunk budget; raise -x");
    lo = first + 1;
    while (lo < hi) {
      uint64_t mid = lo + (hi - lo + 1) / 2;
      if (fits(mid)) lo = mid; else hi = mid - 1;
    }
    if (offsets[lo] > offsets[first]) ranges.emplace_back(first, lo);   // skip ranges without records (:496-499)
    first = lo;
  }
  return ranges;
}

// unit_aligned: batches end where a work unit of process_file ends (classify.cpp:506-521) and a last unit of total
// length 0 is dropped like there (:523-524) — what the per-work-unit sketch rule of the preloaded path needs.  The -x
// path of the reference reads sequence by sequence without work units (:566-791): unit_aligned = false.
static void load_file_batches(const char *filename, vector<Batch> &out, bool unit_aligned = false) {
  SeqReader reader;
  if (!reader.in.open(filename)) die(EX_NOINPUT, string("can't open ") + filename);
  bool fq = reader.fastq = reader.in.peek() == '@';
  const uint64_t BATCH_NT = 96ull << 20;
  out.emplace_back();
  out.back().fastq = fq;
  Read r;
  uint64_t unit_nt = 0;
  size_t unit_first_read = 0;        // first read of the open work unit inside the current batch
  while (reader.valid) {
    Batch &b = out.back();
    if (!reader.next(r, b.bases)) break;
    b.reads.push_back(r);
    b.offs.push_back(b.bases.size());
    bool cut;
    if (!unit_aligned) {
      cut = b.bases.size() >= BATCH_NT || b.reads.size() >= (1u << 20) - 1;
    } else {
      unit_nt += r.seq_len;
      const bool close_unit = unit_nt >= Work_unit_size;
      if (close_unit) { unit_nt = 0; unit_first_read = b.reads.size(); }
      cut = close_unit ? (b.bases.size() >= BATCH_NT || b.reads.size() >= (1u << 20) - 4096)
                       : (b.reads.size() >= (1u << 20) - 1 || b.bases.size() >= (150ull << 20));   // one huge unit: counted as two
    }
    if (cut) { out.emplace_back(); out.back().fastq = fq; unit_first_read = 0; }
  }
  if (unit_aligned && unit_nt == 0 && unit_first_read < out.back().reads.size()) {
    Batch &b = out.back();
    b.bases.resize(b.offs[unit_first_read]);
    b.reads.resize(unit_first_read);
    b.offs.resize(unit_first_read + 1);
  }
  if (out.back().reads.empty()) out.pop_back();
  reader.in.close();
}

static void run_chunked(kuq_ctx *ctx, const Mapped &kdb, const Mapped &idx, uint64_t budget, int argc, char **argv,
                        map<uint32_t, uint64_t> &db_counts) {
  const uint8_t *q = (const uint8_t *)idx.p;
  const uint32_t nt = q[7];
  const uint64_t n_bins = 1ull << (2 * nt);
  const uint64_t *offsets = (const uint64_t *)(q + 8);
  vector<pair<uint64_t, uint64_t>> ranges = plan_ranges(offsets, n_bins, budget);
  cerr << "Database split into " << ranges.size() << " chunks" << endl;
  // pass 0: the taxids of all records, so that every range numbers taxa identically (also database.kdb.counts)
  for (auto &rg : ranges) {
    if (kuq_stage_db(ctx, kdb.p, kdb.size, idx.p, idx.size, rg.first, rg.second)) die(EX_DATAERR, kuq_last_error(ctx));
    uint32_t m = 0;
    kuq_db_taxids(ctx, NULL, NULL, 0, &m);
    vector<uint32_t> tt(m);
    vector<uint64_t> cc(m);
    if (m) kuq_db_taxids(ctx, tt.data(), cc.data(), m, &m);
    for (uint32_t i = 0; i < m; i++) db_counts[tt[i]] += cc[i];
  }
  {
    vector<uint32_t> all;
    for (auto &kv : db_counts) all.push_back(kv.first);
    if (kuq_set_db_taxid_universe(ctx, all.data(), (uint32_t)all.size())) die(EX_SOFTWARE, kuq_last_error(ctx));
  }
  vector<Batch> batches;
  for (int i = optind; i < argc; i++) load_file_batches(argv[i], batches);
  // Preferred path: reads and per-window ids stay in HBM for the whole run and the ranges stream through two device
  // buffers from the (pinned, if the mapping can be registered) database file while the previous range is looked up
  // (kuq_stream_*).  Needs 5 bytes per base + 8 per read of device memory next to the two range buffers; a run that does
  // not fit falls back to the host-side merge below.
  if (!getenv("KUQ_CHUNKS_ON_HOST")) {
    uint64_t need = 0, max_rec = 1, max_bins = 1;
    for (auto &b : batches) need += (b.bases.size() + 64) * 5 + (b.reads.size() + 4) * 8 + 1024;
    const uint64_t key_len_rec = 12;
    for (auto &rg : ranges) {
      max_rec = std::max<uint64_t>(max_rec, offsets[rg.second] - offsets[rg.first]);
      max_bins = std::max<uint64_t>(max_bins, rg.second - rg.first);
    }
    const uint64_t range_bytes = 2 * (max_rec * key_len_rec + (max_bins + 1) * 8 + 4096);
    const uint64_t free_b = kuq_device_free_bytes(ctx);
    if (need + range_bytes + (2ull << 30) < free_b || getenv("KUQ_FORCE_CHUNKS")) {
      uint64_t key_bits;
      memcpy(&key_bits, (const uint8_t *)kdb.p + 8, 8);
      const uint64_t header = 72 + 2 * (4 + 8 * key_bits);
      const uint8_t *rec0 = (const uint8_t *)kdb.p + header;
      const uint32_t k = (uint32_t)(key_bits / 2), idx_type = memcmp(q, "KRAKIDX", 7) == 0 ? 1 : 2;
      // taxonomy must be known before the stream opens (the dense numbering is fixed there): main() sets it before us
      if (kuq_stream_open(ctx, k, nt, idx_type, max_rec, max_bins)) die(EX_DATAERR, kuq_last_error(ctx));
      const bool pinned = kuq_host_register(kdb.p, kdb.size) == KUQ_OK && kuq_host_register(idx.p, idx.size) == KUQ_OK;
      if (Timing) fprintf(stderr, "[timing] chunked: %zu ranges, database mapping %s\n", ranges.size(), pinned ? "pinned" : "pageable");
      struct DevBatch { void *bases, *offs, *codes; };
      vector<DevBatch> dev(batches.size());
      for (size_t bi = 0; bi < batches.size(); bi++) {
        Batch &b = batches[bi];
        DevBatch &d = dev[bi];
        d.bases = kuq_device_alloc(ctx, b.bases.size() + 64);
        d.offs = kuq_device_alloc(ctx, (b.reads.size() + 4) * 8);
        d.codes = kuq_device_alloc(ctx, (b.bases.size() + 64) * 4);
        if (!d.bases || !d.offs || !d.codes) die(EX_OSERR, "device allocation for the read batches failed");
        vector<uint64_t> offs(b.offs);
        offs.push_back(offs.back()); offs.push_back(offs.back());
        if (kuq_device_memset(ctx, 0, (char *)d.bases + b.bases.size(), 'N', 64) ||
            kuq_copy_to_device(ctx, 0, d.bases, b.bases.data(), b.bases.size()) ||
            kuq_copy_to_device(ctx, 0, d.offs, offs.data(), offs.size() * 8) ||
            kuq_device_memset(ctx, 0, d.codes, 0, (b.bases.size() + 64) * 4) || kuq_sync_slot(ctx, 0))
          die(EX_SOFTWARE, kuq_last_error(ctx));
      }
      auto load = [&](uint32_t buf, size_t c) {
        const uint64_t lo = ranges[c].first, hi = ranges[c].second;
        if (kuq_stream_load(ctx, buf, rec0 + offsets[lo] * key_len_rec, offsets[hi] - offsets[lo], offsets + lo, lo, hi))
          die(EX_DATAERR, kuq_last_error(ctx));
      };
      load(0, 0);
      for (size_t c = 0; c < ranges.size(); c++) {
        if (kuq_stream_use(ctx, (uint32_t)(c & 1))) die(EX_SOFTWARE, kuq_last_error(ctx));
        if (c + 1 < ranges.size()) load((uint32_t)((c + 1) & 1), c + 1);
        uint64_t seqs = 0;
        for (size_t bi = 0; bi < batches.size(); bi++) {
          Batch &b = batches[bi];
          if (kuq_lookup_device(ctx, (uint32_t)(bi & 1), (const char *)dev[bi].bases, (const uint64_t *)dev[bi].offs,
                                (uint32_t)b.reads.size(), b.bases.size(), (uint32_t *)dev[bi].codes, 1))
            die(EX_SOFTWARE, kuq_last_error(ctx));
          seqs += b.reads.size();
        }
        fprintf(stderr, "\r Processed %llu sequences (database chunk %zu of %zu)\n", (unsigned long long)seqs, c + 1, ranges.size());
      }
      if (kuq_sync_slot(ctx, 0) || kuq_sync_slot(ctx, 1) || kuq_stream_check(ctx)) die(EX_SOFTWARE, kuq_last_error(ctx));
      for (size_t bi = 0; bi < batches.size(); bi++) {              // final pass (classify.cpp:663-791)
        Batch &b = batches[bi];
        kuq_batch_result res;
        if (kuq_resolve_device(ctx, 0, (const char *)dev[bi].bases, (const uint64_t *)dev[bi].offs, (uint32_t)b.reads.size(),
                               b.bases.size(), (const uint32_t *)dev[bi].codes, NULL, 0) ||
            kuq_collect_device_batch(ctx, 0, &res))
          die(EX_SOFTWARE, kuq_last_error(ctx));
        Fastq_input = b.fastq;
        emit_results(b, res);
        total_classified += res.n_classified;
        total_sequences += b.reads.size();
        total_bases += b.bases.size();
        fprintf(stderr, "\r Processed %llu sequences (%.2f%% classified)", total_sequences, total_classified * 100.0 / total_sequences);
      }
      if (kuq_finish(ctx)) die(EX_SOFTWARE, kuq_last_error(ctx));
      for (auto &d : dev) { kuq_device_free(ctx, d.bases); kuq_device_free(ctx, d.offs); kuq_device_free(ctx, d.codes); }
      if (pinned) { kuq_host_unregister(kdb.p); kuq_host_unregister(idx.p); }
      return;
    }
    cerr << "classify: the reads (" << need / (1 << 20) << " MiB on the device) do not fit next to the database ranges: merging on the host" << endl;
  }
  {
    // host-side merge needs 5 x the input in RAM: say so instead of running into the OOM killer
    uint64_t bytes = 0;
    for (auto &b : batches) bytes += b.bases.size() * 5;
    const uint64_t avail = (uint64_t)sysconf(_SC_AVPHYS_PAGES) * (uint64_t)sysconf(_SC_PAGESIZE);
    if (bytes > avail) die(EX_OSERR, "the host-side merge of the database chunks needs about " + std::to_string(bytes >> 30) +
                                     " GiB of RAM, " + std::to_string(avail >> 30) + " GiB are available: split the input");
  }
  for (auto &b : batches) b.codes.assign(b.bases.size() + 1, 0);
  vector<uint32_t> tmp;
  for (size_t c = 0; c < ranges.size(); c++) {
    // with one range staged in pass 0 last, re-staging it is skipped
    if (!(ranges.size() == 1))
      if (kuq_stage_db(ctx, kdb.p, kdb.size, idx.p, idx.size, ranges[c].first, ranges[c].second)) die(EX_DATAERR, kuq_last_error(ctx));
    uint64_t seqs = 0;
    for (auto &b : batches) {
      tmp.assign(b.bases.size() + 1, 0);
      if (kuq_lookup_batch(ctx, b.bases.data(), b.offs.data(), (uint32_t)b.reads.size(), tmp.data(), NULL))
        die(EX_SOFTWARE, kuq_last_error(ctx));
      {
        uint32_t *dst = b.codes.data();
        const uint32_t *src = tmp.data();
        const long long n = (long long)b.bases.size();
#pragma omp parallel for schedule(static) num_threads(Num_threads)
        for (long long j = 0; j < n; j++) if (src[j] > dst[j]) dst[j] = src[j];                     // merge, :390-485
      }
      seqs += b.reads.size();
      fprintf(stderr, "\r Processed %llu sequences (database chunk %zu of %zu)", (unsigned long long)seqs, c + 1, ranges.size());
    }
    fprintf(stderr, "\r Processed %llu sequences\n", (unsigned long long)seqs);
  }
  // final pass: classify from the merged taxa (classify.cpp:663-791)
  for (auto &b : batches) {
    kuq_batch_result res;
    if (kuq_resolve_batch(ctx, b.bases.data(), b.offs.data(), (uint32_t)b.reads.size(), b.codes.data(), NULL, 0, &res))
      die(EX_SOFTWARE, kuq_last_error(ctx));
    Fastq_input = b.fastq;
    emit_results(b, res);
    total_classified += res.n_classified;
    total_sequences += b.reads.size();
    total_bases += b.bases.size();
    fprintf(stderr, "\r Processed %llu sequences (%.2f%% classified)", total_sequences, total_classified * 100.0 / total_sequences);
  }
  if (kuq_finish(ctx)) die(EX_SOFTWARE, kuq_last_error(ctx));
}

// ---- a database that fits the GPUs of the node together but not one of them: minimizer-range shards ---
// GPU g stages range g.  Every GPU scans every batch and looks up the k-mers whose minimizer it owns; each hit is stored
// straight into the id buffer of the GPU that resolves the read (peer stores over NVLink, kuq_lookup_device_peers), which
// also does the hit's sketch work; device flags order the phases (no host barrier between GPUs); every GPU resolves a
// contiguous share of the batch cut at work-unit boundaries.  One process, so peer access instead of CUDA IPC.  The
// sketches follow the chunked rule, as for any database the reference could only run with -x.
static void run_sharded(const vector<kuq_ctx *> &ctxs, const Mapped &kdb, const Mapped &idx, int argc, char **argv,
                        map<uint32_t, uint64_t> &db_counts) {
  const int G = (int)ctxs.size();
  const uint8_t *q = (const uint8_t *)idx.p;
  const uint32_t nt = q[7];
  const uint64_t n_bins = 1ull << (2 * nt);
  const uint64_t *offsets = (const uint64_t *)(q + 8);
  const uint64_t key_ct = offsets[n_bins];
  vector<uint64_t> cut(G + 1, 0);
  cut[G] = n_bins;
  for (int g = 1; g < G; g++) cut[g] = std::lower_bound(offsets, offsets + n_bins, key_ct / G * g) - offsets;
  for (int g = 1; g <= G; g++) if (cut[g] < cut[g - 1]) cut[g] = cut[g - 1];
  cerr << "Database sharded over " << G << " GPUs by minimizer range" << endl;
  {
    vector<int> rcs(G, 0);
#pragma omp parallel for num_threads(G) schedule(static, 1)
    for (int g = 0; g < G; g++)
      rcs[g] = cut[g + 1] > cut[g] ? kuq_stage_db(ctxs[g], kdb.p, kdb.size, idx.p, idx.size, cut[g], cut[g + 1]) : 0;
    for (int g = 0; g < G; g++) if (rcs[g]) die(EX_DATAERR, kuq_last_error(ctxs[g]));
  }
  for (int g = 0; g < G; g++) {
    if (cut[g + 1] <= cut[g]) die(EX_DATAERR, "more GPUs than non-empty minimizer ranges: use fewer devices (KUQ_DEVICES)");
    uint32_t m = 0;
    kuq_db_taxids(ctxs[g], NULL, NULL, 0, &m);
    vector<uint32_t> tt(m);
    vector<uint64_t> cc(m);
    if (m) kuq_db_taxids(ctxs[g], tt.data(), cc.data(), m, &m);
    for (uint32_t i = 0; i < m; i++) db_counts[tt[i]] += cc[i];
  }
  vector<uint32_t> all;
  for (auto &kv : db_counts) all.push_back(kv.first);
  for (int g = 0; g < G; g++) {
    if (kuq_set_db_taxid_universe(ctxs[g], all.data(), (uint32_t)all.size()) || kuq_set_shard_counting(ctxs[g], 1))
      die(EX_SOFTWARE, kuq_last_error(ctxs[g]));
    for (int h = 0; h < G; h++) if (h != g && kuq_enable_peer_access(ctxs[g], ctxs[h])) die(EX_SOFTWARE, kuq_last_error(ctxs[g]));
  }
  vector<Batch> batches;
  for (int i = optind; i < argc; i++) load_file_batches(argv[i], batches, /*unit_aligned=*/true);
  uint64_t max_bases = 64, max_reads = 4;
  for (auto &b : batches) { max_bases = std::max<uint64_t>(max_bases, b.bases.size()); max_reads = std::max<uint64_t>(max_reads, b.reads.size()); }
  struct Dev { void *bases[2], *offs[2], *ids[2], *flags; };
  vector<Dev> dev(G);
  const uint64_t id_bytes = (max_bases + 64) * 4;
  for (int g = 0; g < G; g++) {
    Dev &d = dev[g];
    for (int k = 0; k < 2; k++) {
      d.bases[k] = kuq_device_alloc(ctxs[g], max_bases + 64);
      d.offs[k] = kuq_device_alloc(ctxs[g], (max_reads + 4) * 8);
      d.ids[k] = kuq_device_alloc(ctxs[g], id_bytes);
      if (!d.bases[k] || !d.offs[k] || !d.ids[k]) die(EX_OSERR, "device allocation for the sharded batches failed");
      if (kuq_device_memset(ctxs[g], 0, d.ids[k], 0, id_bytes)) die(EX_SOFTWARE, kuq_last_error(ctxs[g]));
    }
    d.flags = kuq_device_alloc(ctxs[g], 256);            // [0:8) done counters, [16:24) ready counters
    if (!d.flags || kuq_device_memset(ctxs[g], 0, d.flags, 0, 256)) die(EX_OSERR, "device allocation failed");
    const uint64_t two[8] = {2, 2, 2, 2, 2, 2, 2, 2};     // both id buffers are clean for steps 0 and 1
    if (kuq_copy_to_device(ctxs[g], 0, (char *)d.flags + 128, two, 64) || kuq_sync_slot(ctxs[g], 0)) die(EX_SOFTWARE, kuq_last_error(ctxs[g]));
  }
  uint64_t step = 0;
  for (auto &b : batches) {
    const size_t n = b.reads.size();
    const int par = (int)(step & 1);
    // shares: contiguous, cut where a work unit ends (the chunked rule does not depend on units; even cuts keep the
    // offset slices 16-byte aligned)
    vector<size_t> first(G + 1, 0);
    first[G] = n;
    for (int g = 1; g < G; g++) first[g] = std::min(n, (n * g / G) & ~(size_t)1);
    vector<uint64_t> bounds(G + 1);
    for (int g = 0; g <= G; g++) bounds[g] = b.offs[first[g]];
    bounds[0] = 0;
    vector<uint64_t> offs(b.offs);
    offs.push_back(offs.back()); offs.push_back(offs.back());
    // phase 1: the batch goes to every GPU (a pageable copy may wait for the stream: everything it can wait for is queued)
    for (int g = 0; g < G; g++) {
      if (kuq_copy_to_device(ctxs[g], 0, dev[g].bases[par], b.bases.data(), b.bases.size()) ||
          kuq_device_memset(ctxs[g], 0, (char *)dev[g].bases[par] + b.bases.size(), 'N', 64) ||
          kuq_copy_to_device(ctxs[g], 0, dev[g].offs[par], offs.data(), offs.size() * 8))
        die(EX_SOFTWARE, kuq_last_error(ctxs[g]));
    }
    // phase 2: lookups with peer scatter; phase 3: owners resolve — all stream ordered, flags between the GPUs
    for (int g = 0; g < G; g++) {
      vector<uint32_t *> peer_ids(G);
      vector<uint64_t *> peer_done(G);
      for (int h = 0; h < G; h++) { peer_ids[h] = (uint32_t *)dev[h].ids[par]; peer_done[h] = (uint64_t *)dev[h].flags; }
      if (kuq_wait_flags(ctxs[g], 0, (const uint64_t *)((char *)dev[g].flags + 128), (uint32_t)G, step + 1, 0) ||
          kuq_lookup_device_peers(ctxs[g], 0, (const char *)dev[g].bases[par], (const uint64_t *)dev[g].offs[par], (uint32_t)n,
                                  b.bases.size(), peer_ids.data(), bounds.data(), (uint32_t)G) ||
          kuq_signal_peers(ctxs[g], 0, peer_done.data(), (uint32_t)G, (uint32_t)g, step + 1))
        die(EX_SOFTWARE, kuq_last_error(ctxs[g]));
    }
    for (int g = 0; g < G; g++) {
      vector<uint64_t *> peer_ready(G);
      for (int h = 0; h < G; h++) peer_ready[h] = (uint64_t *)((char *)dev[h].flags + 128);
      const size_t share = first[g + 1] - first[g];
      if (kuq_wait_flags(ctxs[g], 0, (const uint64_t *)dev[g].flags, (uint32_t)G, step + 1, 0)) die(EX_SOFTWARE, kuq_last_error(ctxs[g]));
      if (share && kuq_resolve_device(ctxs[g], 0, (const char *)dev[g].bases[par], (const uint64_t *)dev[g].offs[par] + first[g],
                                      (uint32_t)share, b.bases.size(), (const uint32_t *)dev[g].ids[par], NULL, 0))
        die(EX_SOFTWARE, kuq_last_error(ctxs[g]));
      if (kuq_device_memset(ctxs[g], 0, (char *)dev[g].ids[par] + bounds[g] * 4, 0, (bounds[g + 1] - bounds[g]) * 4) ||
          kuq_signal_peers(ctxs[g], 0, peer_ready.data(), (uint32_t)G, (uint32_t)g, step + 3))
        die(EX_SOFTWARE, kuq_last_error(ctxs[g]));
    }
    // phase 4: results of the shares, in read order
    Fastq_input = b.fastq;
    for (int g = 0; g < G; g++) {
      const size_t share = first[g + 1] - first[g];
      if (!share) { if (kuq_sync_slot(ctxs[g], 0)) die(EX_SOFTWARE, kuq_last_error(ctxs[g])); continue; }
      kuq_batch_result res;
      if (kuq_collect_device_batch(ctxs[g], 0, &res)) die(EX_SOFTWARE, kuq_last_error(ctxs[g]));
      emit_results(b, res, first[g], share);
      total_classified += res.n_classified;
    }
    total_sequences += n;
    total_bases += b.bases.size();
    fprintf(stderr, "\r Processed %llu sequences (%.2f%% classified)", total_sequences, total_classified * 100.0 / total_sequences);
    step++;
  }
  for (int g = 0; g < G; g++) {
    if (kuq_finish(ctxs[g])) die(EX_SOFTWARE, kuq_last_error(ctxs[g]));
    for (int k = 0; k < 2; k++) { kuq_device_free(ctxs[g], dev[g].bases[k]); kuq_device_free(ctxs[g], dev[g].offs[k]); kuq_device_free(ctxs[g], dev[g].ids[k]); }
    kuq_device_free(ctxs[g], dev[g].flags);
  }
}

// ---- several databases: `-d a -d b` (classify.cpp:928-936) ----------------------------------------------------
// For every k-mer the reference asks the databases in command-line order and keeps the value of the first one
// that holds the key — a stored taxon 0 included.  One database is resident in HBM at a time: the reads are
// looked up against each in turn (kuq_lookup_batch reports a stored zero as KUQ_CODE_FOUND_ZERO), the first
// non-zero code per window survives, and a final pass classifies from the merged codes with the per-work-unit
// sketch rule of the preloaded path.  The databases may differ in minimizer length and index type, not in k.
static void run_multi_db(kuq_ctx *ctx, const vector<Mapped> &kdbs, const vector<Mapped> &idxs, int argc, char **argv,
                         vector<map<uint32_t, uint64_t>> &db_counts) {
  const size_t n_db = kdbs.size();
  db_counts.assign(n_db, {});
  set<uint32_t> all;
  for (size_t d = 0; d < n_db; d++) {                      // pass 0: one numbering of the taxa for all databases
    if (kuq_stage_db(ctx, kdbs[d].p, kdbs[d].size, idxs[d].p, idxs[d].size, 0, 0)) die(EX_DATAERR, kuq_last_error(ctx));
    uint32_t m = 0;
    kuq_db_taxids(ctx, NULL, NULL, 0, &m);
    vector<uint32_t> tt(m);
    vector<uint64_t> cc(m);
    if (m) kuq_db_taxids(ctx, tt.data(), cc.data(), m, &m);
    for (uint32_t i = 0; i < m; i++) { db_counts[d][tt[i]] += cc[i]; all.insert(tt[i]); }
  }
  {
    vector<uint32_t> u(all.begin(), all.end());
    if (kuq_set_db_taxid_universe(ctx, u.data(), (uint32_t)u.size())) die(EX_SOFTWARE, kuq_last_error(ctx));
  }
  if (kuq_mark_zero_hits(ctx, 1)) die(EX_SOFTWARE, "kuq_mark_zero_hits");
  vector<Batch> batches;
  for (int i = optind; i < argc; i++) {
    size_t first = batches.size();
    load_file_batches(argv[i], batches, true);
    for (size_t j = first; j < batches.size(); j++) batches[j].file = i;
  }
  {
    // the merged per-window ids of all reads stay on the host between the databases (4 bytes per base + one batch of
    // scratch): say so instead of running into the OOM killer
    uint64_t bytes = 0, largest = 0;
    for (auto &b : batches) { bytes += b.bases.size() * 4; largest = std::max<uint64_t>(largest, b.bases.size() * 4); }
    const uint64_t avail = (uint64_t)sysconf(_SC_AVPHYS_PAGES) * (uint64_t)sysconf(_SC_PAGESIZE);
    if (bytes + largest > avail) die(EX_OSERR, "classifying against several databases needs about " + std::to_string((bytes + largest) >> 30) +
                                               " GiB of RAM for the merged k-mer ids, " + std::to_string(avail >> 30) + " GiB are available: split the input");
  }
  for (auto &b : batches) b.codes.assign(b.bases.size() + 1, 0);
  vector<uint32_t> tmp;
  for (size_t d = 0; d < n_db; d++) {
    if (kuq_stage_db(ctx, kdbs[d].p, kdbs[d].size, idxs[d].p, idxs[d].size, 0, 0)) die(EX_DATAERR, kuq_last_error(ctx));
    uint64_t seqs = 0;
    for (auto &b : batches) {
      tmp.assign(b.bases.size() + 1, 0);
      if (kuq_lookup_batch(ctx, b.bases.data(), b.offs.data(), (uint32_t)b.reads.size(), tmp.data(), NULL))
        die(EX_SOFTWARE, kuq_last_error(ctx));
      {
        uint32_t *dst = b.codes.data();
        const uint32_t *src = tmp.data();
        const long long n = (long long)b.bases.size();
#pragma omp parallel for schedule(static) num_threads(Num_threads)
        for (long long j = 0; j < n; j++) if (dst[j] == 0) dst[j] = src[j];                         // first hit wins
      }
      seqs += b.reads.size();
      fprintf(stderr, "\r Processed %llu sequences (database %zu of %zu)", (unsigned long long)seqs, d + 1, n_db);
    }
    fprintf(stderr, "\r Processed %llu sequences\n", (unsigned long long)seqs);
  }
  for (size_t bi = 0; bi < batches.size(); bi++) {
    Batch &b = batches[bi];
    kuq_batch_result res;
    if (kuq_resolve_batch(ctx, b.bases.data(), b.offs.data(), (uint32_t)b.reads.size(), b.codes.data(), NULL, 0, &res))
      die(EX_SOFTWARE, kuq_last_error(ctx));
    Fastq_input = b.fastq;
    emit_results(b, res);
    total_classified += res.n_classified;
    total_sequences += b.reads.size();
    total_bases += b.bases.size();
    fprintf(stderr, "\r Processed %llu sequences (%.2f%% classified)", total_sequences, total_classified * 100.0 / total_sequences);
    if (bi + 1 == batches.size() || batches[bi + 1].file != b.file)
      if (kuq_finish(ctx)) die(EX_SOFTWARE, kuq_last_error(ctx));      // work units do not span input files
  }
}

int main(int argc, char **argv) {
  parse_command_line(argc, argv);
  if (Map_UIDs) die(EX_USAGE, "-I (UID mapping) is not supported by the GPU classify");
  if (DB_filenames.size() != Index_filenames.size()) die(EX_USAGE, "Must specify a index file for each database file");
  const size_t n_db = DB_filenames.size();
  if (n_db > 1 && Populate_memory_size > 0) die(EX_USAGE, "-x with several databases is not supported by the GPU classify");
  if (Populate_memory && Populate_memory_size == 0) cerr << "Loading database(s)... " << endl;
  vector<Mapped> kdbs(n_db), idxs(n_db);
  for (size_t i = 0; i < n_db; i++) {
    cerr << " Database " << DB_filenames[i] << endl;
    kdbs[i].open(DB_filenames[i]);
    idxs[i].open(Index_filenames[i]);
  }
  Mapped &kdb = kdbs[0], &idx = idxs[0];
  for (size_t i = 1; i < n_db; i++) {                               // classify.cpp:203-210
    auto k_of = [](const Mapped &m) { return m.size >= 16 ? (int)(((const uint64_t *)m.p)[1] / 2) : 0; };
    if (k_of(kdbs[i]) != k_of(kdb)) {
      fprintf(stderr, "Different k-mer sizes in databases 1 and %zu: %i vs %i!\n", i + 1, k_of(kdb), k_of(kdbs[i]));
      exit(1);
    }
  }
  if (optind == argc) {                                            // `-M` without inputs: page-cache warm-up idiom
    if (Populate_memory && Populate_memory_size == 0) cerr << "\ncomplete." << endl;
    return 0;
  }
  if (TaxDB_file.empty()) { cerr << "TaxDB argument is required!" << endl; return 1; }

  // -x: the reference cuts the database into chunks of that many bytes (prepare_chunking).  HBM, not host RAM, is
  // what limits us: the database is split only when it does not fit the HBM budget (or KUQ_FORCE_CHUNKS=1 asks for
  // the reference's exact chunk size); either way -x selects the chunked HLL rule.
  uint64_t hbm_budget = 150ull << 30;
  if (getenv("KUQ_HBM_BUDGET")) hbm_budget = strtoull(getenv("KUQ_HBM_BUDGET"), NULL, 10);
  uint64_t chunk_budget = 0;
  // a database larger than one card: sharded over the visible GPUs when they hold it together (run_sharded), else
  // streamed through one GPU in ranges of at most 40 GB (two range buffers + the reads have to fit, run_chunked)
  bool want_shards = false;
  if (n_db == 1 && (kdb.size + idx.size > hbm_budget || getenv("KUQ_FORCE_SHARDS"))) {
    int n_vis = 0;
    if (const char *dl = getenv("KUQ_DEVICES")) { if (strcmp(dl, "all") == 0) n_vis = kuq_device_count(); else { n_vis = 1; for (const char *c = dl; *c; c++) n_vis += *c == ','; } }
    else if (!getenv("KUQ_DEVICE")) n_vis = kuq_device_count();
    want_shards = n_vis > 1 && (getenv("KUQ_FORCE_SHARDS") || (kdb.size + idx.size) / n_vis + (16ull << 30) < hbm_budget);
  }
  if (kdb.size + idx.size > hbm_budget && !want_shards) chunk_budget = std::min<uint64_t>(hbm_budget / 4, 40ull << 30);
  if (n_db > 1) {
    for (size_t i = 0; i < n_db; i++)
      if (kdbs[i].size + idxs[i].size > hbm_budget) die(EX_USAGE, "with several databases each one has to fit the HBM budget");
    chunk_budget = 0;
  }
  if (Populate_memory_size > 0 && getenv("KUQ_FORCE_CHUNKS")) { chunk_budget = Populate_memory_size; want_shards = false; }
  kuq_config cfg;
  kuq_config_default(&cfg);
  cfg.n_slots = 3;                             // classify b + 1 on the GPU, format b, gather b + 2 (process_file_parallel)
  cfg.max_bases_per_batch = 288ull << 20;      // one read may be a whole chromosome (the largest human one is 248 Mbp)
  cfg.work_unit_size = Work_unit_size;
  // -x (or a database that has to be split) → one global sketch per taxon (classify.cpp:719); else per work unit
  cfg.hll_mode = (Populate_memory_size > 0 || chunk_budget || want_shards) ? KUQ_HLL_CHUNKED : KUQ_HLL_PRELOAD;
#ifdef EXACT_COUNTING
  // classifyExact (classify.cpp:46-49): sets of k-mers instead of sketches; work units and chunks make no difference
  cfg.hll_mode = KUQ_HLL_EXACT;
  if (want_shards) { want_shards = false; chunk_budget = std::min<uint64_t>(hbm_budget / 4, 40ull << 30); }   // exact sets are not merged across devices
  if (Quick_mode) die(EX_USAGE, "-q is not supported by the GPU classifyExact");
#endif
  if ((chunk_budget || want_shards) && !Populate_memory_size)
    cerr << "classify: database larger than one GPU's HBM budget: " << (want_shards ? "sharding it over the GPUs" : "processing it in ranges")
         << " (unique k-mer counts follow the -x rule)" << endl;
  if (getenv("KUQ_SPARSE_SLOTS")) cfg.sparse_set_slots = strtoull(getenv("KUQ_SPARSE_SLOTS"), NULL, 10);
  if (getenv("KUQ_DEVICE")) cfg.device = atoi(getenv("KUQ_DEVICE"));
  // devices: KUQ_DEVICES=0,2,3 (or "all"); default = every visible sm_100 GPU when the database is replicated (one
  // database that fits a card), else one device (KUQ_DEVICE).  Ranges / several databases run on the first device.
  vector<int> devices;
  {
    const char *dl = getenv("KUQ_DEVICES");
    const bool replicable = !chunk_budget && n_db == 1 && cfg.hll_mode != KUQ_HLL_EXACT;   // exact k-mer sets are not merged across devices
    // (replicas when the database fits a card, shards when it only fits the cards together)
    if (dl && strcmp(dl, "all") != 0) {
      for (const char *q = dl; *q;) { devices.push_back(atoi(q)); while (*q && *q != ',') q++; if (*q) q++; }
    } else if ((dl || !getenv("KUQ_DEVICE")) && replicable) {
      for (int d = 0; d < kuq_device_count(); d++) devices.push_back(d);
    }
    if (devices.empty() || !replicable) devices.assign(1, devices.empty() ? cfg.device : devices[0]);
    cfg.device = devices[0];
  }
  kuq_ctx *ctx = NULL;
  int rc = kuq_create(&cfg, &ctx);
  if (rc) die(EX_UNAVAILABLE, string("libkuq: ") + kuq_strerror(rc));
  All_ctx.assign(1, ctx);
  for (size_t g = 1; g < devices.size(); g++) {
    kuq_config c2 = cfg;
    c2.device = devices[g];
    kuq_ctx *cx = NULL;
    int r2 = kuq_create(&c2, &cx);
    if (r2) { fprintf(stderr, "classify: device %d not usable (%s): continuing without it\n", devices[g], kuq_strerror(r2)); continue; }
    All_ctx.push_back(cx);
  }
  if (want_shards && All_ctx.size() < 2) { want_shards = false; chunk_budget = std::min<uint64_t>(hbm_budget / 4, 40ull << 30); }
  if (All_ctx.size() > 1 && !want_shards) fprintf(stderr, "classify: %zu GPUs, database replicated, batches round-robin\n", All_ctx.size());
  // -q: the preloaded path leaves a read at its -m'th hit; with -x every k-mer is counted (classify.cpp:943-944 / :701-738)
  for (kuq_ctx *c : All_ctx)
    if (Quick_mode && kuq_set_quick_mode(c, Minimum_hit_count, Populate_memory_size == 0)) die(EX_SOFTWARE, kuq_last_error(c));
  map<uint32_t, uint64_t> chunk_db_counts;
  vector<map<uint32_t, uint64_t>> multi_db_counts;
  if (!chunk_budget && !want_shards && n_db == 1) {
    double T0__ = now_s();
    // the pinned staging sets of the ingest pipeline are allocated while the database travels to the GPUs
    std::thread prealloc([] {
      const size_t want = 3 * All_ctx.size();
      vector<Stage> tmp(want);
      for (auto &s : tmp) prepare_stage(s, 150ull << 20, 1u << 20);
      Prealloc_stages.swap(tmp);
    });
    // every device stages the database from the same mapped files, in parallel
    vector<int> rcs(All_ctx.size(), 0);
#pragma omp parallel for num_threads((int)All_ctx.size()) schedule(static, 1)
    for (int g = 0; g < (int)All_ctx.size(); g++) rcs[g] = kuq_stage_db(All_ctx[g], kdb.p, kdb.size, idx.p, idx.size, 0, 0);
    for (size_t g = 0; g < All_ctx.size(); g++) if (rcs[g]) die(EX_DATAERR, kuq_last_error(All_ctx[g]));
    prealloc.join();
    TICK("stage_db");
  }
  if (Populate_memory && Populate_memory_size == 0) cerr << "\ncomplete." << endl;

  TaxDB tax;
  cerr << "Reading taxonomy index from " << TaxDB_file;
  tax.read(TaxDB_file);
  cerr << ". Done.\n";
  {
    vector<uint32_t> ids, parents;
    tax.parent_map(ids, parents);
    for (kuq_ctx *c : All_ctx)
      if (kuq_set_taxonomy(c, ids.data(), parents.data(), (uint32_t)ids.size())) die(EX_SOFTWARE, kuq_last_error(c));
  }
  if (Print_classified) Classified_out.open(Classified_output_file);
  if (Print_unclassified) Unclassified_out.open(Unclassified_output_file);
  if (!Kraken_output_file.empty()) {
    if (Kraken_output_file == "off" || Kraken_output_file == "-") Print_kraken = false;   // classify.cpp:233-235
    else { cerr << "Writing Kraken output to " << Kraken_output_file << endl; Kraken_out.open(Kraken_output_file); }
  } else {
    Kraken_out.open("-");
  }

  struct timeval tv1, tv2;
  gettimeofday(&tv1, NULL);
  if (n_db > 1) run_multi_db(ctx, kdbs, idxs, argc, argv, multi_db_counts);
  else if (want_shards) run_sharded(All_ctx, kdb, idx, argc, argv, chunk_db_counts);
  else if (chunk_budget) run_chunked(ctx, kdb, idx, chunk_budget, argc, argv, chunk_db_counts);
  else for (int i = optind; i < argc; i++) process_file(ctx, argv[i]);
  // the other devices' per-taxon state joins the first one's (classify.cpp:542-544 across GPUs)
  for (size_t g = 1; g < All_ctx.size(); g++)
    if (kuq_merge_into(ctx, All_ctx[g])) die(EX_SOFTWARE, kuq_last_error(ctx));
  gettimeofday(&tv2, NULL);
  {                                                                 // report_stats, classify.cpp:361-375
    double seconds = get_seconds(tv1, tv2);
    cerr << "\r";
    fprintf(stderr, "%llu sequences (%.2f Mbp) processed in %.3fs (%.1f Kseq/m, %.2f Mbp/m).\n", total_sequences,
            total_bases / 1.0e6, seconds, total_sequences / 1.0e3 / (seconds / 60), total_bases / 1.0e6 / (seconds / 60));
    fprintf(stderr, "  %llu sequences classified (%.2f%%)\n", total_classified, total_classified * 100.0 / total_sequences);
    fprintf(stderr, "  %llu sequences unclassified (%.2f%%)\n", total_sequences - total_classified,
            (total_sequences - total_classified) * 100.0 / total_sequences);
  }

  if (!Report_output_file.empty() && Report_output_file != "off") {
    gettimeofday(&tv1, NULL);
    cerr << "Writing report file to " << Report_output_file << "  ..\n";
    for (size_t d = 0; d < n_db; d++) {                             // classify.cpp:262-285, once per database
    const string fname = DB_filenames[d] + ".counts";
    bool counts_ok = false;
    {
      ifstream ifs(fname);
      if (ifs.good()) {
        if (ifs.peek() == ifstream::traits_type::eof()) cerr << "Kmer counts file is empty - trying to regenerate ..." << endl;
        else counts_ok = true;
      }
    }
    if (!counts_ok) {                                               // classify.cpp:275-284 via kuq_db_taxids
      cerr << "Writing kmer counts to " << fname << "... [only once for this database, may take a while] " << endl;
      ofstream ofs(fname);
      if (n_db > 1) {
        for (auto &kv : multi_db_counts[d]) ofs << kv.first << '\t' << kv.second << '\n';
      } else if (chunk_budget || want_shards) {
        for (auto &kv : chunk_db_counts) ofs << kv.first << '\t' << kv.second << '\n';
      } else {
        uint32_t m = 0;
        kuq_db_taxids(ctx, NULL, NULL, 0, &m);
        vector<uint32_t> tt(m);
        vector<uint64_t> cc(m);
        if (m) kuq_db_taxids(ctx, tt.data(), cc.data(), m, &m);
        for (uint32_t i = 0; i < m; i++) ofs << tt[i] << '\t' << cc[i] << '\n';
      }
    }
    {
      cerr << "Reading genome sizes from " << fname << " ...";
      ifstream in(fname);
      // readGenomeSizes, taxdb.hpp:868-885, loop kept literally: `while (!eof) { in >> id >> size; set(...) }` counts
      // the last line of a newline-terminated file twice (the failed extraction leaves the variables unchanged);
      // the reference's `cov` column depends on it, so the drop-in reproduces it.
      uint32_t taxid = 0;
      uint64_t size = 0;
      while (!in.eof()) {
        in >> taxid >> size;
        tax.set_genome_size(taxid, size);
      }
      cerr << " done" << endl;
    }
    }
    ostringstream rep;
    double T0__ = now_s();
    print_report(ctx, tax, rep);
    TICK("print_report");
    OutStream ro;
    ro.open(Report_output_file, true);                              // appended to the wrapper's 2 header lines
    ro.write(rep.str());
    ro.close();
    gettimeofday(&tv2, NULL);
    fprintf(stderr, "Report finished in %.3f seconds.\n", get_seconds(tv1, tv2));
  }
  cerr << "Finishing up ..." << endl;
  Kraken_out.close();
  Classified_out.close();
  Unclassified_out.close();
  for (kuq_ctx *c : All_ctx) kuq_destroy(c);
  return 0;
}
