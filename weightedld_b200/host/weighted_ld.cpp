// weighted_ld.cpp — implementation of the C++ mirror API (see weighted_ld.hpp).
#include "weighted_ld.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <charconv>
#include <chrono>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <future>
#include <ios>
#include <mutex>
#include <thread>

namespace weighted_ld {

namespace {
void check(wld_ctx* c, int rc) {
  if (rc != WLD_OK) throw WldError(rc, wld_last_error(c));
}
template <class F>
void parallel_for(size_t n, size_t min_chunk, F&& f) {
  unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  size_t nthr = std::min<size_t>(hw, std::max<size_t>(1, n / std::max<size_t>(min_chunk, 1)));
  if (nthr <= 1) {
    f(0, n);
    return;
  }
  std::vector<std::thread> th;
  for (size_t t = 0; t < nthr; ++t) th.emplace_back([&, t] { f(n * t / nthr, n * (t + 1) / nthr); });
  for (auto& x : th) x.join();
}
}  // namespace

struct MappedFile {
  const char* data = nullptr;
  size_t size = 0;
  explicit MappedFile(const std::string& path) {
    int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) throw std::ios_base::failure(std::string(std::strerror(errno)) + " (os error " + std::to_string(errno) + ")");
    struct stat st;
    if (fstat(fd, &st) != 0) { ::close(fd); throw std::ios_base::failure("fstat failed"); }
    size = (size_t)st.st_size;
    if (size) {
      void* p = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
      if (p == MAP_FAILED) { ::close(fd); throw std::ios_base::failure("mmap failed"); }
      data = (const char*)p;
      madvise(p, size, MADV_WILLNEED);
    }
    ::close(fd);
  }
  MappedFile(const MappedFile&) = delete;
  ~MappedFile() { if (data) munmap((void*)data, size); }
};

// ---------------------------------------------------------------------------------------------
// read_fasta, lib.rs:277-307.  Line-oriented: '>' lines are names; EVERY other line is one whole
// sequence INCLUDING its '\n' (lib.rs:297), so an alignment of L bases has L+1 columns, the last
// being Unknown.  A final line without '\n' is one column short -> ragged -> panic in from_multiseq.
// The file is mapped, all host threads look for the line ends of their share of it, and the sequences
// are NOT copied: MultiSequence::rows points into the mapping, from where the library's staging threads
// gather them on their way to the GPU (wld_load_alignment_rows).
// ---------------------------------------------------------------------------------------------
MultiSequence read_fasta(const std::string& path) {
  MultiSequence ms;
  ms.source = path;
  auto map = std::make_shared<MappedFile>(path);
  const char* data = map->data;
  const size_t size = map->size;
  if (size == 0) return ms;

  // line ends, found in parallel: thread t scans [size*t/T, size*(t+1)/T)
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  const size_t nthr = std::max<size_t>(1, std::min<size_t>(hw, size / (4u << 20)));
  std::vector<std::vector<size_t>> found(nthr);
  {
    std::vector<std::thread> th;
    for (size_t t = 0; t < nthr; ++t)
      th.emplace_back([&, t] {
        size_t pos = size * t / nthr;
        const size_t end = size * (t + 1) / nthr;
        auto& v = found[t];
        while (pos < end) {
          const char* nl = (const char*)memchr(data + pos, '\n', end - pos);
          if (!nl) break;
          v.push_back((size_t)(nl - data));
          pos = (size_t)(nl - data) + 1;
        }
      });
    for (auto& x : th) x.join();
  }
  size_t n_nl = 0;
  for (auto& v : found) n_nl += v.size();
  std::vector<size_t> starts;  // start offset of every line; starts[k+1] - starts[k] = its length incl. '\n'
  starts.reserve(n_nl + 2);
  starts.push_back(0);
  for (auto& v : found)
    for (size_t e : v) starts.push_back(e + 1);
  if (starts.back() != size) starts.push_back(size);  // last line has no terminator
  const size_t n_lines = starts.size() - 1;

  ms.rows.reserve(n_lines / 2 + 1);
  const char* pending = nullptr;
  size_t pending_len = 0;
  for (size_t k = 0; k < n_lines; ++k) {
    const size_t pos = starts[k], len = starts[k + 1] - pos;
    if (data[pos] == '>') {
      pending = data + pos + 1;  // lib.rs:294-295 keeps the newline in the name
      pending_len = len - 1;
    } else {
      if (ms.rows.empty()) ms.n_cols = (int64_t)len;
      else if ((int64_t)len != ms.n_cols) ms.ragged = true;
      ms.rows.push_back((const uint8_t*)data + pos);
      ms.names.emplace_back(pending ? std::string(pending, pending_len) : std::string());
      pending = nullptr;
    }
  }
  ms.n_seqs = (int64_t)ms.rows.size();
  ms.row_stride = ms.n_cols;
  ms.mapping = std::move(map);
  return ms;
}

// ---------------------------------------------------------------------------------------------
// read_fasta_python, WeightedLD.py:21-41 through Bio.AlignIO.read(.., "fasta"): a record is a '>'
// line plus every following line, concatenated without line terminators; symbols are lower-cased and
// a c g t - map to 0..4, everything else to 5.  Returns codes (MultiSequence::codes = true).
// ---------------------------------------------------------------------------------------------
MultiSequence read_fasta_python(const std::string& path) {
  MultiSequence ms;
  ms.source = path;
  ms.codes = true;
  MappedFile f(path);
  uint8_t lut[256];
  std::memset(lut, 5, sizeof lut);
  lut['a'] = lut['A'] = 0; lut['c'] = lut['C'] = 1; lut['g'] = lut['G'] = 2; lut['t'] = lut['T'] = 3; lut['-'] = 4;
  std::vector<std::vector<uint8_t>> recs;
  bool in_rec = false;
  size_t pos = 0;
  while (pos < f.size) {
    const char* nl = (const char*)memchr(f.data + pos, '\n', f.size - pos);
    size_t end = nl ? (size_t)(nl - f.data) : f.size;
    size_t a = pos, b = end;
    while (a < b && std::isspace((unsigned char)f.data[a])) ++a;
    while (b > a && std::isspace((unsigned char)f.data[b - 1])) --b;
    if (pos < end && f.data[pos] == '>') {
      recs.emplace_back();
      ms.names.emplace_back(f.data + pos + 1, end - pos - 1);
      in_rec = true;
    } else if (in_rec) {
      auto& r = recs.back();
      for (size_t i = a; i < b; ++i) r.push_back(lut[(unsigned char)f.data[i]]);
    }
    pos = end + 1;
  }
  ms.n_seqs = (int64_t)recs.size();
  if (recs.empty()) return ms;
  ms.n_cols = (int64_t)recs[0].size();
  for (auto& r : recs)
    if ((int64_t)r.size() != ms.n_cols) ms.ragged = true;
  if (ms.ragged) return ms;
  ms.row_stride = (ms.n_cols + 15) / 16 * 16;
  ms.chars.assign((size_t)ms.row_stride * recs.size(), 5);
  for (size_t i = 0; i < recs.size(); ++i) std::memcpy(ms.chars.data() + i * (size_t)ms.row_stride, recs[i].data(), recs[i].size());
  return ms;
}

// ---------------------------------------------------------------------------------------------
// read_vcf, WeightedLD.py:311-379.  The reference munges the text with regular expressions; for
// well-formed phased diploid GT-only input that amounts to: variant lines follow the first line
// containing "#CHROM"; the LAST line of the file is always dropped (:365 expects a trailing blank
// line); POS (column 2) labels the site; every sample call `a|b` yields two haplotypes with the allele
// index as the symbol code, `x/y` (unphased) and `.` become 4 (Missing); np.rot90 (:375) turns the
// site x haplotype table into haplotype x site with the haplotype order REVERSED.  Variant rows are
// parsed by all host threads.
// ---------------------------------------------------------------------------------------------
MultiSequence read_vcf(const std::string& path) {
  MultiSequence ms;
  ms.source = path;
  ms.codes = true;
  MappedFile f(path);
  std::vector<std::pair<size_t, size_t>> lines;  // [begin, end) without the newline
  bool header_seen = false;
  size_t pos = 0;
  while (pos <= f.size) {
    const char* nl = pos < f.size ? (const char*)memchr(f.data + pos, '\n', f.size - pos) : nullptr;
    const size_t end = nl ? (size_t)(nl - f.data) : f.size;
    if (header_seen) {
      lines.emplace_back(pos, end);
    } else {
      static const char key[] = "#CHROM";
      if (end - pos >= 6 && std::search(f.data + pos, f.data + end, key, key + 6) != f.data + end) header_seen = true;
    }
    if (!nl) break;
    pos = end + 1;
  }
  if (!header_seen) throw std::runtime_error("No #CHROM header block identified");            // :327-330
  auto count_fields = [&](size_t a, size_t b) { return (size_t)std::count(f.data + a, f.data + b, '\t') + 1; };
  if (lines.empty() || count_fields(lines[0].first, lines[0].second) <= 12)                    // :333-337
    throw std::runtime_error("The VCF data contains too small a population, are you sure this is a multi VCF?");
  lines.pop_back();                                                                            // :365
  const size_t n_sites = lines.size();
  if (n_sites == 0) return ms;
  std::vector<std::vector<uint8_t>> site_haps(n_sites);
  ms.site_labels.assign(n_sites, 0);
  std::vector<std::string> errs(n_sites);
  parallel_for(n_sites, 4, [&](size_t lo, size_t hi) {
    for (size_t k = lo; k < hi; ++k) {
      const char* p = f.data + lines[k].first;
      const char* e = f.data + lines[k].second;
      if (e > p && e[-1] == '\r') --e;
      int field = 0;
      const char* q = p;
      while (q < e && field < 9) {  // skip CHROM..FORMAT, remember POS
        const char* t = (const char*)memchr(q, '\t', (size_t)(e - q));
        if (!t) { q = e; break; }
        if (field == 1) ms.site_labels[k] = std::strtoll(std::string(q, t).c_str(), nullptr, 10);
        q = t + 1;
        ++field;
      }
      if (field < 9) { errs[k] = "variant row with fewer than 10 columns"; continue; }
      auto& hap = site_haps[k];
      while (q <= e) {
        const char* t = (const char*)memchr(q, '\t', (size_t)(e - q));
        const char* fe = t ? t : e;
        // one call: alleles separated by '|'; a three-character `x/y` is unphased -> both missing (:353)
        if (fe - q == 3 && q[1] == '/') {
          hap.push_back(4);
          hap.push_back(4);
        } else {
          const char* a = q;
          while (a <= fe) {
            const char* bar = (const char*)memchr(a, '|', (size_t)(fe - a));
            const char* ae = bar ? bar : fe;
            if (ae - a == 1 && *a == '.') hap.push_back(4);                                    // :356
            else {
              int v = 0;
              bool ok = ae > a;
              for (const char* c = a; c < ae; ++c) { if (*c < '0' || *c > '9') ok = false; v = v * 10 + (*c - '0'); if (v > 255) ok = false; }
              if (!ok) { errs[k] = "call '" + std::string(q, fe) + "' is not a phased GT-only genotype"; break; }
              hap.push_back((uint8_t)v);
            }
            if (!bar) break;
            a = bar + 1;
          }
        }
        if (!t) break;
        q = t + 1;
      }
    }
  });
  for (auto& er : errs)
    if (!er.empty()) throw std::runtime_error("VCF: " + er);
  const size_t n_haps = site_haps[0].size();
  for (auto& h : site_haps)
    if (h.size() != n_haps) throw std::runtime_error("VCF: variant rows with different numbers of calls");
  ms.n_seqs = (int64_t)n_haps;
  ms.n_cols = (int64_t)n_sites;
  ms.row_stride = (ms.n_cols + 15) / 16 * 16;
  ms.chars.assign((size_t)ms.row_stride * n_haps, 5);
  parallel_for(n_haps, 1024, [&](size_t lo, size_t hi) {
    for (size_t h = lo; h < hi; ++h) {  // np.rot90: output row h = input haplotype n_haps-1-h
      uint8_t* dst = ms.chars.data() + h * (size_t)ms.row_stride;
      for (size_t k = 0; k < n_sites; ++k) dst[k] = site_haps[k][n_haps - 1 - h];
    }
  });
  ms.names.assign(n_haps, std::string());
  return ms;
}

// ---------------------------------------------------------------------------------------------
struct SiteSet::Impl {
  std::vector<wld_ctx*> ctx;
  ~Impl() {
    for (auto* c : ctx) wld_destroy(c);
  }
};
SiteSet::~SiteSet() = default;
SiteSet::SiteSet(SiteSet&&) noexcept = default;

std::shared_ptr<SiteSet::Impl> SiteSet::open_devices(const std::vector<int>& devices) {
  auto impl = std::make_shared<Impl>();
  impl->ctx.assign(devices.size(), nullptr);
  std::vector<int> rcs(devices.size(), WLD_OK);
  std::vector<std::string> msgs(devices.size());
  std::vector<std::thread> th;  // one context per GPU, created concurrently
  for (size_t g = 0; g < devices.size(); ++g)
    th.emplace_back([&, g] {
      wld_ctx* c = nullptr;
      rcs[g] = wld_create(devices[g], &c);
      if (rcs[g] != WLD_OK) {
        msgs[g] = c ? wld_last_error(c) : "allocation failed";
        if (c) wld_destroy(c);
        c = nullptr;
      }
      impl->ctx[g] = c;
    });
  for (auto& t : th) t.join();
  for (size_t g = 0; g < devices.size(); ++g)
    if (rcs[g] != WLD_OK) {
      for (auto*& c : impl->ctx) {
        if (c) wld_destroy(c);
        c = nullptr;
      }
      impl->ctx.clear();
      throw WldError(rcs[g], msgs[g]);
    }
  return impl;
}

SiteSet SiteSet::from_multiseq(const MultiSequence& ms, const std::vector<int>& devices) {
  if (ms.n_seqs == 0) throw Panic("index out of bounds: the len is 0 but the index is 0");  // lib.rs:178
  if (ms.ragged) throw Panic("Not all sequences have the same number of symbols");        // lib.rs:181
  return from_multiseq(ms, open_devices(devices));
}

SiteSet SiteSet::from_multiseq(const MultiSequence& ms, std::shared_ptr<Impl> opened) {
  if (ms.n_seqs == 0) throw Panic("index out of bounds: the len is 0 but the index is 0");  // lib.rs:178
  if (ms.ragged) throw Panic("Not all sequences have the same number of symbols");        // lib.rs:181
  SiteSet s;
  s.impl = std::move(opened);
  const int n = (int)s.impl->ctx.size();
  std::vector<std::thread> th;
  std::vector<std::string> errs((size_t)n);
  for (int g = 0; g < n; ++g)
    th.emplace_back([&, g] {
      wld_ctx* c = s.impl->ctx[(size_t)g];
      int rc = wld_set_partition(c, g, n);
      // stages 1-2 are split over the GPUs: GPU g counts the rows [g*R, (g+1)*R) and later sums those sequences
      const int64_t per = (ms.n_seqs + n - 1) / n, lo = std::min<int64_t>(g * per, ms.n_seqs), hi = std::min<int64_t>(lo + per, ms.n_seqs);
      if (rc == WLD_OK) rc = wld_set_row_shard(c, lo, n > 1 ? hi : -1);
      if (rc == WLD_OK) rc = wld_set_seq_shard(c, lo, n > 1 ? hi : -1);
      if (rc == WLD_OK)
        rc = !ms.rows.empty()
                 ? wld_load_alignment_rows(c, ms.rows.data(), ms.n_seqs, ms.n_cols, ms.codes ? WLD_INPUT_CODES : WLD_INPUT_ASCII)
                 : wld_load_alignment(c, ms.chars.data(), ms.n_seqs, ms.n_cols, ms.row_stride,
                                      ms.codes ? WLD_INPUT_CODES : WLD_INPUT_ASCII);
      if (rc != WLD_OK) errs[(size_t)g] = wld_last_error(c);
    });
  for (auto& t : th) t.join();
  for (auto& e : errs)
    if (!e.empty()) throw WldError(WLD_ERR_CUDA, e);
  if (n > 1) check(s.impl->ctx[0], wld_sum_histograms(s.impl->ctx.data(), n));  // exact integer sum over the GPUs
  return s;
}

SiteSet SiteSet::filter_by(float min_acgt, float min_minor, float max_minor) const {
  for (auto* c : impl->ctx) {
    int64_t kept = 0;
    check(c, wld_filter_sites(c, min_acgt, min_minor, max_minor, &kept));
  }
  SiteSet f;
  f.impl = impl;
  f.filtered = true;
  return f;
}

SiteSet SiteSet::filter_by_python(double min_acgt, double min_variability) const {
  for (auto* c : impl->ctx) {
    int64_t kept = 0;
    check(c, wld_filter_sites_python(c, min_acgt, min_variability, &kept));
  }
  SiteSet f;
  f.impl = impl;
  f.filtered = true;
  return f;
}

SiteSet SiteSet::keep_all() const {
  for (auto* c : impl->ctx) {
    int64_t kept = 0;
    check(c, wld_keep_all_sites(c, &kept));
  }
  SiteSet f;
  f.impl = impl;
  f.filtered = true;
  return f;
}

void SiteSet::set_python_compat(bool on) const {
  for (auto* c : impl->ctx) check(c, wld_set_compat(c, on ? WLD_COMPAT_PYTHON : WLD_COMPAT_RUST));
}

int64_t SiteSet::n_sites() const { return filtered ? wld_n_kept(impl->ctx[0]) : wld_n_cols(impl->ctx[0]); }
int64_t SiteSet::n_seqs() const { return wld_n_seqs(impl->ctx[0]); }
std::vector<int64_t> SiteSet::site_map() const {
  std::vector<int64_t> m((size_t)wld_n_kept(impl->ctx[0]));
  check(impl->ctx[0], wld_get_site_map(impl->ctx[0], m.data(), (int64_t)m.size()));
  return m;
}
int64_t SiteSet::parent_site_index(int64_t idx) const { return filtered ? site_map()[(size_t)idx] : idx; }

std::vector<float> henikoff_weights(const SiteSet& data) {
  std::vector<float> w((size_t)data.n_seqs());
  auto& ctxs = data.impl->ctx;
  for (auto* c : ctxs) check(c, wld_henikoff(c));  // asynchronous: the GPUs sum their sequence shards concurrently
  if (ctxs.size() > 1) {
    check(ctxs[0], wld_share_weight_sums(ctxs.data(), (int)ctxs.size()));
    for (auto* c : ctxs) check(c, wld_henikoff_finish(c));
  }
  check(data.impl->ctx[0], wld_get_weights(data.impl->ctx[0], w.data(), (int64_t)w.size()));
  return w;
}

PairStore all_weighted_ld_pairs(const SiteSet& site_set, const std::vector<float>& weights, float thr,
                                const std::function<void(size_t)>& progress_report) {
  auto& ctxs = site_set.impl->ctx;
  const int n = (int)ctxs.size();
  if (progress_report) progress_report(0);  // lib.rs:584
  struct Shared {
    std::mutex mu;
    std::vector<uint64_t> done;
    const std::function<void(size_t)>* cb;
  } shared;
  shared.done.assign((size_t)n, 0);
  shared.cb = &progress_report;
  struct User {
    Shared* sh;
    int g;
  };
  auto tramp = [](uint64_t done, void* user) {
    User* u = (User*)user;
    std::lock_guard<std::mutex> lk(u->sh->mu);  // lib.rs:672-674: callback serialised under a mutex
    u->sh->done[(size_t)u->g] = std::max(u->sh->done[(size_t)u->g], done);
    uint64_t total = 0;
    for (auto d : u->sh->done) total += d;
    if (total > 0 && *u->sh->cb) (*u->sh->cb)((size_t)total);
  };
  std::vector<uint64_t> computed((size_t)n, 0), survivors((size_t)n, 0);
  std::vector<std::string> errs((size_t)n);
  std::vector<User> users;
  for (int g = 0; g < n; ++g) users.push_back(User{&shared, g});
  std::vector<std::thread> th;
  for (int g = 0; g < n; ++g)
    th.emplace_back([&, g] {
      wld_ctx* c = ctxs[(size_t)g];
      uint64_t ns = 0, nc = 0;
      int rc = wld_set_weights(c, weights.data(), (int64_t)weights.size());
      if (rc == WLD_OK) rc = wld_ld_pairs(c, thr, +tramp, &users[(size_t)g], &ns, &nc);
      if (rc != WLD_OK) errs[(size_t)g] = wld_last_error(c);
      computed[(size_t)g] = nc;
      survivors[(size_t)g] = ns;
    });
  for (auto& t : th) t.join();
  for (auto& e : errs)
    if (!e.empty()) throw WldError(WLD_ERR_CUDA, e);
  PairStore store;
  for (auto v : computed) store.pairs_computed += v;
  // Merge (the order-preserving rayon collect of lib.rs:635-679, across devices): the other GPUs' survivors
  // travel to GPU 0 by peer copy and are ordered there together with its own when the store is read.
  store.n_ = survivors[0];
  for (int g = 1; g < n; ++g) {
    check(ctxs[0], wld_append_pairs_from(ctxs[0], ctxs[(size_t)g]));
    store.n_ += survivors[(size_t)g];
  }
  store.root_ = ctxs[0];
  store.keep_ = site_set.impl;
  return store;
}

const PairVec& PairStore::pairs() const {
  if (!cached_) {
    cache_.resize((size_t)n_);
    uint64_t got = 0;
    if (n_) check(root_, wld_fetch_pairs(root_, cache_.data(), n_, WLD_FETCH_PARENT_INDEX, &got));
    cached_ = true;
  }
  return cache_;
}

void PairStore::for_each_chunk(size_t chunk_pairs, const std::function<void(const wld_pair*, size_t, size_t)>& fn) const {
  if (cached_ || n_ <= chunk_pairs) {  // small: one fetch
    const PairVec& all = pairs();
    if (!all.empty()) fn(all.data(), 0, all.size());
    return;
  }
  PairVec buf[2];
  buf[0].resize(chunk_pairs);
  buf[1].resize(chunk_pairs);
  auto fetch = [&](int which, size_t first) -> size_t {
    uint64_t got = 0;
    check(root_, wld_fetch_pairs_range(root_, first, chunk_pairs, buf[which].data(), WLD_FETCH_PARENT_INDEX, &got));
    return (size_t)got;
  };
  size_t first = 0, have = fetch(0, 0);
  int cur = 0;
  while (have) {
    const size_t next_first = first + have;
    size_t next_have = 0;
    std::exception_ptr err;
    std::thread prefetch;
    if (next_first < n_)
      prefetch = std::thread([&] {
        try { next_have = fetch(cur ^ 1, next_first); } catch (...) { err = std::current_exception(); }
      });
    fn(buf[cur].data(), first, have);
    if (prefetch.joinable()) prefetch.join();
    if (err) std::rethrow_exception(err);
    first = next_first;
    have = next_have;
    cur ^= 1;
  }
}

// ---------------------------------------------------------------------------------------------
// Writers, main.rs:70-119.  Rust `{:.3}`: exact value rounded half-to-even (glibc %.3f agrees for
// finite values), `NaN`, `inf`, `-inf`, negative zero keeps its sign.
// ---------------------------------------------------------------------------------------------
static inline int fmt_u64(unsigned long long v, char* buf) {
  char tmp[24];
  int n = 0;
  do {
    tmp[n++] = (char)('0' + v % 10);
    v /= 10;
  } while (v);
  for (int i = 0; i < n; ++i) buf[i] = tmp[n - 1 - i];
  return n;
}
// `{:.3}` of an f32.  x*1000 is exact in f64 (24-bit significand times a 10-bit integer), so rounding it
// to the nearest integer, ties to even, IS round-half-even of the exact binary value — what Rust's
// float formatting (and glibc's %.3f) produce.  Values too large for the integer path go through printf.
static int fmt_f3(float v, char* buf) {
  if (std::isnan(v)) return std::sprintf(buf, "NaN");
  if (std::isinf(v)) return std::sprintf(buf, v < 0 ? "-inf" : "inf");
  const double x = (double)v;
  if (std::fabs(x) >= 1e15) return std::sprintf(buf, "%.3f", x);
  // llrint: one cvtsd2si in the default rounding mode (to nearest, ties to even) instead of a libm call
  const unsigned long long m = (unsigned long long)std::llrint(std::fabs(x) * 1000.0);
  int k = 0;
  if (std::signbit(v)) buf[k++] = '-';
  k += fmt_u64(m / 1000, buf + k);
  const unsigned f = (unsigned)(m % 1000);
  buf[k++] = '.';
  buf[k++] = (char)('0' + f / 100);
  buf[k++] = (char)('0' + f / 10 % 10);
  buf[k++] = (char)('0' + f % 10);
  buf[k] = 0;
  return k;
}
std::string format_f3(float v) {
  char b[64];
  fmt_f3(v, b);
  return b;
}

void write_henikoff_weights(const std::string& path, const std::vector<float>& weights) {
  FILE* f = std::fopen(path.c_str(), "w");
  if (!f) throw std::ios_base::failure(std::string(std::strerror(errno)) + " (os error " + std::to_string(errno) + ")");
  std::fputs("Sequence_index\thk_weight\n", f);
  char b[64];
  for (size_t i = 0; i < weights.size(); ++i) {
    fmt_f3(weights[i], b);
    std::fprintf(f, "%zu\t%s\n", i, b);
  }
  std::fclose(f);
}

// repr(round(numpy.float64(v), 4)): numpy rounds as rint(v*1e4)/1e4; repr prints the shortest digits
// that round-trip, in fixed notation unless the decimal exponent is < -4 or >= 16.
std::string format_py4(double v) {
  if (std::isnan(v)) return "nan";
  if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
  const double r = std::nearbyint(v * 1e4) / 1e4;
  if (r == 0.0) return std::signbit(r) ? "-0.0" : "0.0";
  char buf[64];
  auto res = std::to_chars(buf, buf + sizeof buf, std::fabs(r), std::chars_format::scientific);
  std::string sci(buf, res.ptr);                 // d[.ddd]e[+-]XX
  const size_t epos = sci.find('e');
  std::string digits;
  for (size_t i = 0; i < epos; ++i)
    if (sci[i] != '.') digits.push_back(sci[i]);
  const int exp10 = std::atoi(sci.c_str() + epos + 1);
  const int decpt = exp10 + 1;                   // value = 0.DIGITS x 10^decpt
  std::string out = r < 0 ? "-" : "";
  if (decpt <= -4 || decpt > 16) {
    out += digits.substr(0, 1);
    if (digits.size() > 1) out += "." + digits.substr(1);
    char eb[16];
    std::snprintf(eb, sizeof eb, "e%c%02d", exp10 < 0 ? '-' : '+', std::abs(exp10));
    out += eb;
  } else if (decpt <= 0) {
    out += "0." + std::string((size_t)(-decpt), '0') + digits;
  } else if ((size_t)decpt >= digits.size()) {
    out += digits + std::string((size_t)decpt - digits.size(), '0') + ".0";
  } else {
    out += digits.substr(0, (size_t)decpt) + "." + digits.substr((size_t)decpt);
  }
  return out;
}

namespace {
// Formats `n` records with `fmt_line` on all host threads and writes the pieces IN ORDER behind the writer's file
// position: every thread formats a contiguous run into its own buffer and the run lengths give each buffer its file
// offset.  The bytes then go into the file from all threads at once — through a shared mapping of the file's new
// range (concurrent write() calls on one file serialise on its inode lock; page faults on a mapping do not), or with
// pwrite() where the output cannot be mapped (a pipe, a device) — in the BACKGROUND, while the caller fetches and
// formats the next chunk.
struct WriteTiming { double format_s = 0, write_s = 0; };
WriteTiming g_write_timing;
double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct OrderedWriter {
  int fd;
  off_t pos = 0;               // end of what has been handed to the writer so far
  bool can_map = true;         // cleared on the first failure to extend / map the file
  std::future<int> pending;    // errno of the write in flight (0 = fine)

  explicit OrderedWriter(int f) : fd(f) {}
  void wait() {
    if (pending.valid()) {
      const int e = pending.get();
      if (e) throw std::ios_base::failure(std::string(std::strerror(e)) + " (os error " + std::to_string(e) + ")");
    }
  }
  // takes ownership of the formatted pieces; returns at once
  void submit(std::vector<std::string>&& pieces) {
    wait();  // one write in flight: keeps the order of failures simple and the memory bounded
    std::vector<off_t> at(pieces.size());
    const off_t begin = pos;
    for (size_t t = 0; t < pieces.size(); ++t) {
      at[t] = pos;
      pos += (off_t)pieces[t].size();
    }
    const off_t end = pos;
    pending = std::async(std::launch::async, [this, begin, end, at = std::move(at), pieces = std::move(pieces)]() -> int {
      const double t0 = now_s();
      auto each = [&](auto&& f) {
        std::vector<std::thread> th;
        for (size_t t = 1; t < pieces.size(); ++t) th.emplace_back([&, t] { f(t); });
        f(0);
        for (auto& x : th) x.join();
      };
      int err = 0;
      bool done = false;
      if (can_map && end > begin) {
        const long page = sysconf(_SC_PAGESIZE);
        const off_t map_lo = begin / page * page;
        if (::ftruncate(fd, end) == 0) {
          void* m = ::mmap(nullptr, (size_t)(end - map_lo), PROT_READ | PROT_WRITE, MAP_SHARED, fd, map_lo);
          if (m != MAP_FAILED) {
            char* base = static_cast<char*>(m) - map_lo;  // base[file offset]
            each([&](size_t t) { std::memcpy(base + at[t], pieces[t].data(), pieces[t].size()); });
            ::munmap(m, (size_t)(end - map_lo));
            done = true;
          }
        }
        if (!done) can_map = false;
      }
      if (!done) {
        std::vector<int> errs(pieces.size(), 0);
        each([&](size_t t) {
          const char* p = pieces[t].data();
          size_t left = pieces[t].size();
          off_t o = at[t];
          while (left) {
            const ssize_t w = ::pwrite(fd, p, left, o);
            if (w < 0) { if (errno == EINTR) continue; errs[t] = errno; return; }
            p += w; left -= (size_t)w; o += w;
          }
        });
        for (int e : errs) if (e && !err) err = e;
      }
      g_write_timing.write_s += now_s() - t0;
      return err;
    });
  }
};

template <class LineFn>
void write_records_parallel(OrderedWriter& w, const wld_pair* recs, size_t n, LineFn&& fmt_line) {
  if (n == 0) return;
  const double t_begin = now_s();
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  const size_t nthr = std::max<size_t>(1, std::min<size_t>(hw, n / 4096));
  std::vector<std::string> out(nthr);
  auto format = [&](size_t t) {
    const size_t lo = n * t / nthr, hi = n * (t + 1) / nthr;
    std::string& s = out[t];  // lines are formatted in place: no per-line append
    s.resize((hi - lo) * 40 + 256);
    size_t used = 0;
    for (size_t i = lo; i < hi; ++i) {
      if (s.size() - used < 192) s.resize(s.size() + s.size() / 4 + 4096);
      used += (size_t)fmt_line(recs[i], &s[used]);
    }
    s.resize(used);
  };
  if (nthr == 1) {
    format(0);
  } else {
    std::vector<std::thread> th;
    for (size_t t = 0; t < nthr; ++t) th.emplace_back([&, t] { format(t); });
    for (auto& x : th) x.join();
  }
  g_write_timing.format_s += now_s() - t_begin;
  w.submit(std::move(out));
}

int open_out(const std::string& path) {
  const int fd = ::open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
  if (fd < 0) throw std::ios_base::failure(std::string(std::strerror(errno)) + " (os error " + std::to_string(errno) + ")");
  return fd;
}
void write_header(int fd, off_t* pos, const char* text) {
  const size_t n = std::strlen(text);
  if (::pwrite(fd, text, n, *pos) != (ssize_t)n) { const int e = errno; ::close(fd); throw std::ios_base::failure(std::string(std::strerror(e))); }
  *pos += (off_t)n;
}
}  // namespace

void write_pair_stats_python(const std::string& path, const PairStore& store, const std::vector<int64_t>& labels) {
  // WeightedLD.py:177-179 prints in plain row-major order of the upper triangle
  PairVec sorted = store.pairs();
  std::sort(sorted.begin(), sorted.end(), [](const wld_pair& x, const wld_pair& y) {
    return x.site_a != y.site_a ? x.site_a < y.site_a : x.site_b < y.site_b;
  });
  auto line = [&](const wld_pair& p, char* buf) {
    const long long la = labels.empty() ? (long long)p.site_a : (long long)labels[p.site_a];
    const long long lb = labels.empty() ? (long long)p.site_b : (long long)labels[p.site_b];
    return std::snprintf(buf, 192, "%lld\t%lld\t%s\t%s\t%s\n", la, lb, format_py4((double)p.d).c_str(),
                         format_py4((double)p.d_prime).c_str(), format_py4((double)p.r2).c_str());
  };
  if (path == "-") {
    std::fputs("posa\tposb\tD\tD'\tR2\n", stdout);
    char buf[192];
    for (const wld_pair& p : sorted) std::fwrite(buf, 1, (size_t)line(p, buf), stdout);
    std::fflush(stdout);
    return;
  }
  const int fd = open_out(path);
  OrderedWriter w(fd);
  write_header(fd, &w.pos, "posa\tposb\tD\tD'\tR2\n");
  try {
    write_records_parallel(w, sorted.data(), sorted.size(), line);
    w.wait();
  } catch (...) { ::close(fd); throw; }
  ::close(fd);
}

void write_pair_stats(const std::string& path, const PairStore& store, const std::vector<int64_t>& labels) {
  const int fd = open_out(path);
  OrderedWriter w(fd);
  write_header(fd, &w.pos, "site_a\tsite_b\td\td'\tr2\n");
  auto line = [&](const wld_pair& p, char* buf) {
    int k = 0;
    if (labels.empty()) {
      k += fmt_u64(p.site_a, buf + k);
      buf[k++] = '\t';
      k += fmt_u64(p.site_b, buf + k);
      buf[k++] = '\t';
    } else {
      k = std::sprintf(buf, "%lld\t%lld\t", (long long)labels[p.site_a], (long long)labels[p.site_b]);
    }
    k += fmt_f3(p.d, buf + k);
    buf[k++] = '\t';
    k += fmt_f3(p.d_prime, buf + k);
    buf[k++] = '\t';
    k += fmt_f3(p.r2, buf + k);
    buf[k++] = '\n';
    return k;
  };
  g_write_timing = WriteTiming{};
  const double t0 = now_s();
  try {
    // streamed: the next chunk is fetched from the device while this one is formatted and written
    store.for_each_chunk((size_t)4 << 20, [&](const wld_pair* recs, size_t, size_t count) {
      write_records_parallel(w, recs, count, line);
    });
    w.wait();
  } catch (...) { ::close(fd); throw; }
  ::close(fd);
  if (std::getenv("WLD_CLI_TIMING"))
    std::fprintf(stderr, "[weighted_ld] pair writer: %.1f ms total, %.1f ms formatting, %.1f ms writing (in the background, mapped file), rest = waiting for the device copy; %u host threads\n",
                 (now_s() - t0) * 1e3, g_write_timing.format_s * 1e3, g_write_timing.write_s * 1e3, std::thread::hardware_concurrency());
}

}  // namespace weighted_ld
