// weighted_ld.cpp — implementation of the C++ mirror API (see weighted_ld.hpp).
#include "weighted_ld.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
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

// ---------------------------------------------------------------------------------------------
// read_fasta, lib.rs:277-307.  Line-oriented: '>' lines are names; EVERY other line is one whole
// sequence INCLUDING its '\n' (lib.rs:297), so an alignment of L bases has L+1 columns, the last
// being Unknown.  A final line without '\n' is one column short -> ragged -> panic in from_multiseq.
// The file is mmap'ed and the rows are copied by all host threads into a 16-byte-pitched buffer.
// ---------------------------------------------------------------------------------------------
MultiSequence read_fasta(const std::string& path) {
  MultiSequence ms;
  ms.source = path;
  int fd = ::open(path.c_str(), O_RDONLY);
  if (fd < 0) throw std::ios_base::failure(std::string(std::strerror(errno)) + " (os error " + std::to_string(errno) + ")");
  struct stat st;
  if (fstat(fd, &st) != 0) { ::close(fd); throw std::ios_base::failure("fstat failed"); }
  const size_t size = (size_t)st.st_size;
  if (size == 0) { ::close(fd); return ms; }
  const char* data = (const char*)mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
  ::close(fd);
  if (data == MAP_FAILED) throw std::ios_base::failure("mmap failed");
  madvise((void*)data, size, MADV_SEQUENTIAL);

  std::vector<std::pair<size_t, size_t>> rows;  // (offset, length incl. newline)
  std::string pending;
  bool have_name = false;
  size_t pos = 0;
  while (pos < size) {
    const char* nl = (const char*)memchr(data + pos, '\n', size - pos);
    const size_t end = nl ? (size_t)(nl - data) + 1 : size;
    if (data[pos] == '>') {
      pending.assign(data + pos + 1, end - pos - 1);  // lib.rs:294-295 keeps the newline in the name
      have_name = true;
    } else {
      rows.emplace_back(pos, end - pos);
      ms.names.push_back(have_name ? pending : std::string());
      have_name = false;
    }
    pos = end;
  }
  ms.n_seqs = (int64_t)rows.size();
  if (!rows.empty()) {
    ms.n_cols = (int64_t)rows[0].second;
    for (auto& r : rows)
      if ((int64_t)r.second != ms.n_cols) ms.ragged = true;
    if (!ms.ragged) {
      ms.row_stride = (ms.n_cols + 15) / 16 * 16;
      ms.chars.resize((size_t)ms.row_stride * rows.size());
      parallel_for(rows.size(), 256, [&](size_t a, size_t b) {
        for (size_t i = a; i < b; ++i) {
          uint8_t* dst = ms.chars.data() + i * (size_t)ms.row_stride;
          memcpy(dst, data + rows[i].first, (size_t)ms.n_cols);
          memset(dst + ms.n_cols, 0, (size_t)(ms.row_stride - ms.n_cols));
        }
      });
    }
  }
  munmap((void*)data, size);
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

SiteSet SiteSet::from_multiseq(const MultiSequence& ms, const std::vector<int>& devices) {
  if (ms.n_seqs == 0) throw Panic("index out of bounds: the len is 0 but the index is 0");  // lib.rs:178
  if (ms.ragged) throw Panic("Not all sequences have the same number of symbols");        // lib.rs:181
  SiteSet s;
  s.impl = std::make_shared<Impl>();
  for (int dev : devices) {
    wld_ctx* c = nullptr;
    int rc = wld_create(dev, &c);
    if (rc != WLD_OK) {
      std::string msg = c ? wld_last_error(c) : "allocation failed";
      if (c) wld_destroy(c);
      throw WldError(rc, msg);
    }
    s.impl->ctx.push_back(c);
  }
  const int n = (int)s.impl->ctx.size();
  std::vector<std::thread> th;
  std::vector<std::string> errs((size_t)n);
  for (int g = 0; g < n; ++g)
    th.emplace_back([&, g] {
      wld_ctx* c = s.impl->ctx[(size_t)g];
      int rc = wld_set_partition(c, g, n);
      if (rc == WLD_OK) rc = wld_load_alignment(c, ms.chars.data(), ms.n_seqs, ms.n_cols, ms.row_stride, WLD_INPUT_ASCII);
      if (rc != WLD_OK) errs[(size_t)g] = wld_last_error(c);
    });
  for (auto& t : th) t.join();
  for (auto& e : errs)
    if (!e.empty()) throw WldError(WLD_ERR_CUDA, e);
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
  for (auto* c : data.impl->ctx) check(c, wld_henikoff(c));
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
  std::vector<std::vector<wld_pair>> parts((size_t)n);
  std::vector<uint64_t> computed((size_t)n, 0);
  std::vector<std::string> errs((size_t)n);
  std::vector<User> users;
  for (int g = 0; g < n; ++g) users.push_back(User{&shared, g});
  std::vector<std::thread> th;
  for (int g = 0; g < n; ++g)
    th.emplace_back([&, g] {
      wld_ctx* c = ctxs[(size_t)g];
      uint64_t ns = 0, nc = 0, got = 0;
      int rc = wld_set_weights(c, weights.data(), (int64_t)weights.size());
      if (rc == WLD_OK) rc = wld_ld_pairs(c, thr, +tramp, &users[(size_t)g], &ns, &nc);
      if (rc == WLD_OK) {
        parts[(size_t)g].resize((size_t)ns);
        rc = wld_fetch_pairs(c, parts[(size_t)g].data(), ns, n > 1 ? WLD_FETCH_KEPT_INDEX : WLD_FETCH_PARENT_INDEX, &got);
      }
      if (rc != WLD_OK) errs[(size_t)g] = wld_last_error(c);
      computed[(size_t)g] = nc;
    });
  for (auto& t : th) t.join();
  for (auto& e : errs)
    if (!e.empty()) throw WldError(WLD_ERR_CUDA, e);
  PairStore store;
  for (auto v : computed) store.pairs_computed += v;
  if (n == 1) {
    store.pairs.swap(parts[0]);
    return store;
  }
  // host merge of the per-GPU shards (each already in reference order) by (tile key, a, b)
  const int64_t n_kept = wld_n_kept(ctxs[0]);
  const std::vector<int64_t> smap = site_set.site_map();
  size_t total = 0;
  for (auto& p : parts) total += p.size();
  store.pairs.reserve(total);
  for (auto& p : parts) store.pairs.insert(store.pairs.end(), p.begin(), p.end());
  std::sort(store.pairs.begin(), store.pairs.end(), [&](const wld_pair& x, const wld_pair& y) {
    const uint64_t kx = wld_pair_order_key(n_kept, x.site_a, x.site_b), ky = wld_pair_order_key(n_kept, y.site_a, y.site_b);
    if (kx != ky) return kx < ky;
    if (x.site_a != y.site_a) return x.site_a < y.site_a;
    return x.site_b < y.site_b;
  });
  for (auto& p : store.pairs) {
    p.site_a = (uint32_t)smap[p.site_a];
    p.site_b = (uint32_t)smap[p.site_b];
  }
  return store;
}

// ---------------------------------------------------------------------------------------------
// Writers, main.rs:70-119.  Rust `{:.3}`: exact value rounded half-to-even (glibc %.3f agrees for
// finite values), `NaN`, `inf`, `-inf`, negative zero keeps its sign.
// ---------------------------------------------------------------------------------------------
static int fmt_f3(float v, char* buf) {
  if (std::isnan(v)) return std::sprintf(buf, "NaN");
  if (std::isinf(v)) return std::sprintf(buf, v < 0 ? "-inf" : "inf");
  return std::sprintf(buf, "%.3f", (double)v);
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

void write_pair_stats(const std::string& path, const PairStore& store) {
  FILE* f = std::fopen(path.c_str(), "w");
  if (!f) throw std::ios_base::failure(std::string(std::strerror(errno)) + " (os error " + std::to_string(errno) + ")");
  std::fputs("site_a\tsite_b\td\td'\tr2\n", f);
  // format in parallel chunks, write in order
  const size_t n = store.pairs.size(), chunk = 1 << 16;
  for (size_t base = 0; base < n; base += chunk * 64) {
    const size_t hi = std::min(n, base + chunk * 64);
    const size_t nchunks = (hi - base + chunk - 1) / chunk;
    std::vector<std::string> out(nchunks);
    parallel_for(nchunks, 1, [&](size_t a, size_t b) {
      char line[160];
      for (size_t ci = a; ci < b; ++ci) {
        std::string& s = out[ci];
        const size_t lo = base + ci * chunk, up = std::min(hi, lo + chunk);
        s.reserve((up - lo) * 40);
        for (size_t i = lo; i < up; ++i) {
          const wld_pair& p = store.pairs[i];
          int k = std::sprintf(line, "%u\t%u\t", p.site_a, p.site_b);
          k += fmt_f3(p.d, line + k);
          line[k++] = '\t';
          k += fmt_f3(p.d_prime, line + k);
          line[k++] = '\t';
          k += fmt_f3(p.r2, line + k);
          line[k++] = '\n';
          s.append(line, (size_t)k);
        }
      }
    });
    for (auto& s : out) std::fwrite(s.data(), 1, s.size(), f);
  }
  std::fclose(f);
}

}  // namespace weighted_ld
