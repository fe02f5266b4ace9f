// weighted_ld.hpp — C++ host-side mirror of the reference's Rust crate API (rust/weighted_ld/src/lib.rs)
// on top of the C ABI (include/wld.h).  Same names, argument meaning and error behaviour:
//   read_fasta                lib.rs:277-307     MultiSequence (rows keep their newline column)
//   SiteSet::from_multiseq    lib.rs:176-206     throws where the reference panics (lib.rs:180-182)
//   SiteSet::filter_by        lib.rs:230-251 + is_site_of_interest lib.rs:310-338 + main.rs:139
//   henikoff_weights          lib.rs:340-358
//   all_weighted_ld_pairs     lib.rs:578-684     -> PairStore (lib.rs:529-576), reference order
// plus the input side and dialect of the reference's Python program (WeightedLD.py), which is the only
// reference implementation that reads VCF:
//   read_vcf                  WeightedLD.py:311-379   read_fasta_python   WeightedLD.py:21-41
//   SiteSet::filter_by_python WeightedLD.py:44-98     set_python_compat   wld_set_compat (include/wld.h)
//   write_pair_stats_python   WeightedLD.py:176,283-284 (`posa posb D D' R2`, round(x, 4))
// No CPU fallback: every stage below the text parser is a CUDA kernel in libwld.so.
#pragma once
#include <cstdint>
#include <functional>
#include <memory>
#include <new>
#include <type_traits>
#include <utility>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/wld.h"

namespace weighted_ld {

struct WldError : std::runtime_error {
  int status;
  WldError(int s, const std::string& m) : std::runtime_error(m), status(s) {}
};
struct Panic : std::runtime_error {  // the reference's panic!() sites
  using std::runtime_error::runtime_error;
};

struct MappedFile;                     // read-only mmap of the input (weighted_ld.cpp)

struct MultiSequence {                 // lib.rs:153-156
  std::string source;
  int64_t n_seqs = 0, n_cols = 0, row_stride = 0;
  // The sequences, one of two ways.  read_fasta leaves them WHERE THEY ARE — rows[r] points at the r-th
  // sequence line inside the mapped file (the reference's Vec<Sequence>, lib.rs:148-156, without the copy) and the
  // library gathers them straight into its pinned staging buffers (wld_load_alignment_rows).  The VCF and
  // Python-dialect readers build a matrix: n_seqs rows, pitch row_stride (16-byte multiple).
  std::vector<const uint8_t*> rows;
  std::shared_ptr<MappedFile> mapping; // keeps rows[] valid
  std::vector<uint8_t> chars;
  const uint8_t* row(int64_t r) const { return rows.empty() ? chars.data() + (size_t)(r * row_stride) : rows[(size_t)r]; }
  std::vector<std::string> names;      // lib.rs:143-146 (empty = None)
  bool ragged = false;                 // rows of unequal length: from_multiseq panics (lib.rs:180-182)
  bool codes = false;                  // chars already hold 0..5 codes (VCF allele indices, Python FASTA reader)
  std::vector<int64_t> site_labels;    // VCF: POS of every column (WeightedLD.py:369); empty = column index
};

MultiSequence read_fasta(const std::string& path);  // throws std::ios_base::failure on I/O errors
// WeightedLD.py:21-41 (Bio.AlignIO): multi-line records, no newline column, returns 0..5 codes.
MultiSequence read_fasta_python(const std::string& path);
// WeightedLD.py:311-379: phased diploid GT-only VCF -> haplotype x site codes (allele index, '.' and
// unphased calls -> 4), haplotypes in reversed column order, last line dropped, POS as site labels.
MultiSequence read_vcf(const std::string& path);

struct LdStats { float r2, d, d_prime; };          // lib.rs:382-387

// std::allocator that leaves trivially constructible elements uninitialised on resize(): the survivor
// buffer (up to tens of GB) is filled by the device-to-host copy, zeroing it first would be a wasted pass
template <class T>
struct default_init_allocator : std::allocator<T> {
  template <class U> struct rebind { using other = default_init_allocator<U>; };
  using std::allocator<T>::allocator;
  template <class U> void construct(U* p) noexcept(std::is_nothrow_default_constructible<U>::value) { ::new (static_cast<void*>(p)) U; }
  template <class U, class... A> void construct(U* p, A&&... a) { ::new (static_cast<void*>(p)) U(std::forward<A>(a)...); }
};
using PairVec = std::vector<wld_pair, default_init_allocator<wld_pair>>;

// lib.rs:529-576.  The survivors stay ON THE DEVICE, already merged over all GPUs, until they are asked for:
// pairs() copies all of them (reference order, parent indices); for_each_chunk streams them, fetching the next
// chunk while the caller works on the current one (the TSV writers use it, so a result larger than host
// memory could still be written).
class SiteSet;
class PairStore {
 public:
  uint64_t pairs_computed = 0;
  size_t len() const { return (size_t)n_; }
  const PairVec& pairs() const;
  void for_each_chunk(size_t chunk_pairs, const std::function<void(const wld_pair*, size_t first, size_t count)>& fn) const;

 private:
  friend PairStore all_weighted_ld_pairs(const SiteSet&, const std::vector<float>&, float, const std::function<void(size_t)>&);
  std::shared_ptr<void> keep_;                      // the contexts
  wld_ctx* root_ = nullptr;                         // holds the merged survivors
  uint64_t n_ = 0;
  mutable PairVec cache_;
  mutable bool cached_ = false;
};

class SiteSet {                                     // lib.rs:158-275
 public:
  // One context per GPU; the alignment is replicated to each.
  struct Impl;
  // Creates the library contexts (CUDA initialisation takes seconds): callable from a background thread
  // while the input file is being read; pass the result to from_multiseq.
  static std::shared_ptr<Impl> open_devices(const std::vector<int>& devices = {0});
  static SiteSet from_multiseq(const MultiSequence& ms, const std::vector<int>& devices = {0});
  static SiteSet from_multiseq(const MultiSequence& ms, std::shared_ptr<Impl> opened);
  SiteSet filter_by(float min_acgt_frac, float min_minor, float max_minor) const;  // main.rs:139-143
  SiteSet filter_by_python(double min_acgt, double min_variability) const;         // WeightedLD.py:44-98
  SiteSet keep_all() const;                        // no site filter (the VCF path of WeightedLD.py:385-386)
  void set_python_compat(bool on) const;           // wld_set_compat on every context
  int64_t n_sites() const;                          // lib.rs:254
  int64_t n_seqs() const;                           // lib.rs:259
  int64_t parent_site_index(int64_t idx) const;     // lib.rs:263
  std::vector<int64_t> site_map() const;
  ~SiteSet();
  SiteSet(SiteSet&&) noexcept;
  SiteSet(const SiteSet&) = delete;

  std::shared_ptr<Impl> impl;                       // shared with the filtered view (same contexts)
  bool filtered = false;
 private:
  SiteSet() = default;
};

std::vector<float> henikoff_weights(const SiteSet& data);  // lib.rs:340
PairStore all_weighted_ld_pairs(const SiteSet& site_set, const std::vector<float>& weights, float r2_threshold,
                                const std::function<void(size_t)>& progress_report);  // lib.rs:578

// main.rs:70-119
void write_henikoff_weights(const std::string& path, const std::vector<float>& weights);
// labels (optional): VCF POS printed instead of the column index
void write_pair_stats(const std::string& path, const PairStore& pairs, const std::vector<int64_t>& labels = {});
std::string format_f3(float v);  // Rust `{:.3}`
// WeightedLD.py:176,283-284; labels[i] replaces site index i when non-empty (VCF POS)
void write_pair_stats_python(const std::string& path, const PairStore& pairs, const std::vector<int64_t>& labels);
std::string format_py4(double v);  // repr(round(numpy.float64(v), 4))

}  // namespace weighted_ld
