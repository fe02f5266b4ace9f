// main.cpp — the `weighted_ld` command-line tool on the B200 library: same long flags, defaults,
// stage order, stderr log lines and TSV outputs as the reference binary
// (rust/weighted_ld/src/main.rs:14-213).  Extensions: --gpus N (tile-partitioned pair stage);
// --vcf-input (the VCF reader of the reference's Python program, WeightedLD.py:311-379; sites are
// labelled by POS); --python-compat (the whole Python dialect: its FASTA reader, site filter with
// --min-variability, Henikoff fill, per-pair allele calls, no r2 threshold, `posa posb D D' R2` output
// rounded to 4 decimals, WeightedLD.py:382-402; VCF input is then not site-filtered, as there).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <future>
#include <iostream>
#include <string>

#include "weighted_ld.hpp"

using namespace weighted_ld;
using Clock = std::chrono::steady_clock;

namespace {
int g_level = 3;  // 0 off, 1 error, 2 warn, 3 info, 4 debug  (env_logger default "info", main.rs:122)

void init_logger() {
  const char* e = std::getenv("RUST_LOG");
  if (!e) return;
  std::string s(e);
  for (auto& ch : s) ch = (char)std::tolower(ch);
  if (s == "off") g_level = 0;
  else if (s == "error") g_level = 1;
  else if (s == "warn") g_level = 2;
  else if (s == "info") g_level = 3;
  else if (s == "debug" || s == "trace") g_level = 4;
}
void log_line(int level, const char* name, const std::string& msg) {
  if (level > g_level) return;
  char ts[32];
  std::time_t t = std::time(nullptr);
  std::strftime(ts, sizeof ts, "%Y-%m-%dT%H:%M:%SZ", std::gmtime(&t));
  std::fprintf(stderr, "[%s %-5s weighted_ld] %s\n", ts, name, msg.c_str());
}
#define INFO(msg) log_line(3, "INFO", msg)
#define DEBUG(msg) log_line(4, "DEBUG", msg)

// Rust `{:?}` of a Duration: largest unit with a non-zero integer part, fraction trimmed.
std::string fmt_duration(Clock::duration d) {
  const double ns = (double)std::chrono::duration_cast<std::chrono::nanoseconds>(d).count();
  char b[64];
  auto trim = [&](double v, const char* unit) {
    std::snprintf(b, sizeof b, "%.9f", v);
    std::string s(b);
    while (!s.empty() && s.back() == '0') s.pop_back();
    if (!s.empty() && s.back() == '.') s.pop_back();
    return s + unit;
  };
  if (ns >= 1e9) return trim(ns / 1e9, "s");
  if (ns >= 1e6) return trim(ns / 1e6, "ms");
  if (ns >= 1e3) return trim(ns / 1e3, "\xc2\xb5s");
  return trim(ns, "ns");
}
// human_format::Formatter (1.0.3) defaults: 2 decimals, separator " ", SI suffixes, optional units.
std::string human(double v, const char* units = "") {
  static const char* suf[] = {"", "k", "M", "G", "T", "P", "E", "Z", "Y"};
  int i = 0;
  while (std::fabs(v) >= 1000.0 && i < 8) {
    v /= 1000.0;
    ++i;
  }
  char b[64];
  std::snprintf(b, sizeof b, "%.2f %s%s", v, suf[i], units);
  return b;
}

struct Opt {  // main.rs:19-68
  std::string fasta_input, vcf_input, weights_output, pair_output;
  float min_acgt = 0.8f, min_minor = 0.02f, max_minor = 0.5f, r2_threshold = 0.1f;
  double min_acgt_f64 = 0.8, min_variability = 0.02;  // --python-compat parses these as Python floats
  bool unweighted = false, python_compat = false, parse_only = false;
  int gpus = 1;
};

void usage(FILE* f) {
  std::fputs(
      "weighted_ld 0.1.0\nA tool for computing sequence weighted linkage disequilibrium\n\n"
      "USAGE:\n    weighted_ld [FLAGS] [OPTIONS] --fasta-input <fasta-input> --pair-output <pair-output>\n\n"
      "FLAGS:\n    -h, --help          Prints help information\n"
      "        --unweighted    Use unit weights instead of Henikoff weights\n"
      "    -V, --version       Prints version information\n\n"
      "OPTIONS:\n        --fasta-input <fasta-input>          The source file to load\n"
      "        --max-minor <max-minor>              Maximum fraction of minor symbols for a site to be considered [default: 0.5]\n"
      "        --min-acgt <min-acgt>                Minimum fractions of ACTG for a site to be considered [default: 0.8]\n"
      "        --min-minor <min-minor>              Minimum fraction of minor symbols for a site to be considered [default: 0.02]\n"
      "        --pair-output <pair-output>          Filename to write the per-pair weighted LD figures to, in Tab Separated Value format\n"
      "        --r2-threshold <r2-threshold>        Minimum value of R2 to be included in the output [default: 0.1]\n"
      "        --weights-output <weights-output>    Filename to write the per-sequence weights to, in Tab Separated Value format\n"
      "        --gpus <gpus>                        (B200 build) number of GPUs for the pair stage [default: 1]\n"
      "        --vcf-input <vcf-input>              (B200 build) phased diploid VCF instead of --fasta-input (WeightedLD.py reader)\n"
      "        --python-compat                      (B200 build) numeric dialect and output of WeightedLD.py\n"
      "        --parse-only                         (B200 build) read the input, print its shape and an FNV-1a checksum, exit (no GPU needed)\n"
      "        --min-variability <v>                (B200 build, --python-compat) minimum non-major fraction [default: 0.02]\n",
      f);
}

bool parse(int argc, char** argv, Opt& o) {
  auto need = [&](int& i, const char* flag) -> const char* {
    const char* eq = std::strchr(argv[i], '=');
    if (eq) return eq + 1;
    if (i + 1 >= argc) {
      std::fprintf(stderr, "error: The argument '%s <value>' requires a value but none was supplied\n", flag);
      std::exit(1);
    }
    return argv[++i];
  };
  auto is = [&](const char* a, const char* flag) {
    const size_t n = std::strlen(flag);
    return std::strncmp(a, flag, n) == 0 && (a[n] == 0 || a[n] == '=');
  };
  for (int i = 1; i < argc; ++i) {
    const char* a = argv[i];
    if (!std::strcmp(a, "-h") || !std::strcmp(a, "--help")) { usage(stdout); std::exit(0); }
    if (!std::strcmp(a, "-V") || !std::strcmp(a, "--version")) { std::puts("weighted_ld 0.1.0"); std::exit(0); }
    if (is(a, "--format-py4")) {  // self-test hook (no GPU): prints repr(round(float64(v), 4)) and `{:.3}` of f32(v)
      const double v = std::strtod(need(i, "--format-py4"), nullptr);
      std::printf("%s %s\n", format_py4(v).c_str(), format_f3((float)v).c_str());
      std::exit(0);
    }
    if (!std::strcmp(a, "--unweighted")) o.unweighted = true;
    else if (!std::strcmp(a, "--python-compat")) o.python_compat = true;
    else if (!std::strcmp(a, "--parse-only")) o.parse_only = true;
    else if (is(a, "--fasta-input")) o.fasta_input = need(i, "--fasta-input");
    else if (is(a, "--vcf-input")) o.vcf_input = need(i, "--vcf-input");
    else if (is(a, "--min-variability")) o.min_variability = std::strtod(need(i, "--min-variability"), nullptr);
    else if (is(a, "--pair-output")) o.pair_output = need(i, "--pair-output");
    else if (is(a, "--weights-output")) o.weights_output = need(i, "--weights-output");
    else if (is(a, "--min-acgt")) {
      const char* v = need(i, "--min-acgt");
      o.min_acgt = std::strtof(v, nullptr);
      o.min_acgt_f64 = std::strtod(v, nullptr);
    }
    else if (is(a, "--min-minor")) o.min_minor = std::strtof(need(i, "--min-minor"), nullptr);
    else if (is(a, "--max-minor")) o.max_minor = std::strtof(need(i, "--max-minor"), nullptr);
    else if (is(a, "--r2-threshold")) o.r2_threshold = std::strtof(need(i, "--r2-threshold"), nullptr);
    else if (is(a, "--gpus")) o.gpus = std::atoi(need(i, "--gpus"));
    else {
      std::fprintf(stderr, "error: Found argument '%s' which wasn't expected, or isn't valid in this context\n\nUSAGE:\n    weighted_ld [FLAGS] [OPTIONS] --fasta-input <fasta-input> --pair-output <pair-output>\n\nFor more information try --help\n", a);
      return false;
    }
  }
  if (!o.vcf_input.empty() && o.fasta_input.empty()) o.fasta_input = o.vcf_input;  // one of the two is required
  if (o.parse_only && !o.fasta_input.empty()) return true;
  if (o.fasta_input.empty() || o.pair_output.empty()) {
    std::fprintf(stderr, "error: The following required arguments were not provided:\n%s%s\nUSAGE:\n    weighted_ld [FLAGS] [OPTIONS] --fasta-input <fasta-input> --pair-output <pair-output>\n\nFor more information try --help\n",
                 o.fasta_input.empty() ? "    --fasta-input <fasta-input>\n" : "", o.pair_output.empty() ? "    --pair-output <pair-output>\n" : "");
    return false;
  }
  return true;
}
}  // namespace

int main(int argc, char** argv) {
  init_logger();
  Opt opt;
  if (!parse(argc, argv, opt)) return 1;
  try {
    if (opt.parse_only) {  // host-side ingest only: shape + checksum of the byte matrix (row-major, without pitch padding)
      const bool is_vcf = !opt.vcf_input.empty();
      MultiSequence ms = is_vcf ? read_vcf(opt.vcf_input)
                                : opt.python_compat ? read_fasta_python(opt.fasta_input) : read_fasta(opt.fasta_input);
      if (ms.ragged) throw Panic("Not all sequences have the same number of symbols");
      uint64_t h = 1469598103934665603ull;
      for (int64_t r = 0; r < ms.n_seqs; ++r) {
        const uint8_t* row = ms.row(r);
        for (int64_t c = 0; c < ms.n_cols; ++c) h = (h ^ row[c]) * 1099511628211ull;
      }
      uint64_t hl = 1469598103934665603ull;
      for (int64_t v : ms.site_labels)
        for (int b = 0; b < 8; ++b) hl = (hl ^ (uint8_t)((uint64_t)v >> (8 * b))) * 1099511628211ull;
      std::printf("%lld %lld %016llx %s %016llx\n", (long long)ms.n_seqs, (long long)ms.n_cols, (unsigned long long)h,
                  ms.codes ? "codes" : "ascii", (unsigned long long)hl);
      return 0;
    }
    std::vector<int> devices;
    for (int g = 0; g < std::max(1, opt.gpus); ++g) devices.push_back(g);
    // The CUDA driver initialises every GPU it can see (about 0.2 s each on an 8-GPU box): show it only the ones
    // this run uses, unless the caller already chose.
    if (!std::getenv("CUDA_VISIBLE_DEVICES")) {
      std::string vis;
      for (int g : devices) vis += (vis.empty() ? "" : ",") + std::to_string(g);
      setenv("CUDA_VISIBLE_DEVICES", vis.c_str(), 0);
    }

    auto sw = Clock::now();
    // CUDA initialisation takes seconds and is independent of the input: do it while the file is read
    auto opened = std::async(std::launch::async, [&] { return SiteSet::open_devices(devices); });
    const bool vcf = !opt.vcf_input.empty();
    MultiSequence multiseq = vcf ? read_vcf(opt.vcf_input)                // WeightedLD.py:311-379
                                 : opt.python_compat ? read_fasta_python(opt.fasta_input)  // WeightedLD.py:21-41
                                                     : read_fasta(opt.fasta_input);        // main.rs:129
    DEBUG("parsed input in " + fmt_duration(Clock::now() - sw));
    auto impl = opened.get();
    DEBUG("devices ready after " + fmt_duration(Clock::now() - sw));
    SiteSet siteset = SiteSet::from_multiseq(multiseq, impl);             // main.rs:130
    INFO(std::string("Loaded ") + (vcf ? "vcf" : "fasta") + " file in " + fmt_duration(Clock::now() - sw));  // main.rs:131
    INFO("    " + std::to_string(siteset.n_seqs()) + " sequences, " + std::to_string(siteset.n_sites()) + " sites");
    siteset.set_python_compat(opt.python_compat);

    sw = Clock::now();
    SiteSet filtered = !opt.python_compat ? siteset.filter_by(opt.min_acgt, opt.min_minor, opt.max_minor)  // main.rs:139-143
                       : vcf ? siteset.keep_all()                                                          // WeightedLD.py:385-386
                             : siteset.filter_by_python(opt.min_acgt_f64, opt.min_variability);            // WeightedLD.py:298-304
    INFO("Computed + filtered sites of interest in " + fmt_duration(Clock::now() - sw));
    INFO("    Found " + std::to_string(filtered.n_sites()) + " sites of interest");

    std::vector<float> weights;
    if (opt.unweighted) {
      weights.assign((size_t)siteset.n_seqs(), 1.0f);                     // main.rs:150-153
    } else {
      sw = Clock::now();
      weights = henikoff_weights(filtered);                               // main.rs:156
      INFO("Computed Henikoff weights in " + fmt_duration(Clock::now() - sw));
    }
    if (!opt.weights_output.empty()) {                                    // main.rs:161-164
      INFO("Writing weights to \"" + opt.weights_output + "\"");
      write_henikoff_weights(opt.weights_output, weights);
    }

    INFO("Beginning pairwise weighted LD computation");                   // main.rs:166
    sw = Clock::now();
    const int64_t L = filtered.n_sites();
    const uint64_t total_pairs = (uint64_t)((L - 1) * (L - 2) / 2);       // main.rs:168 (sic)
    // the Python program prints every computed pair (no threshold, WeightedLD.py:283-284)
    const float thr = opt.python_compat ? -INFINITY : opt.r2_threshold;
    PairStore store = all_weighted_ld_pairs(filtered, weights, thr, [&](size_t computed) {
      if (g_level >= 4) DEBUG("progress " + std::to_string(computed) + "/" + std::to_string(total_pairs));
    });
    const auto dur = Clock::now() - sw;
    INFO("Finished computing pairwise weighted LD stats in " + fmt_duration(dur));
    const double secs = std::chrono::duration<double>(dur).count();
    INFO("    " + human((double)total_pairs) + " pairs computed at ~" + human((double)total_pairs / secs, "pairs/s") + ", " +
         human((double)store.len()) + " passed threshold");               // main.rs:196-205

    INFO("Writing output to \"" + opt.pair_output + "\"");                // main.rs:207
    sw = Clock::now();
    if (opt.python_compat) write_pair_stats_python(opt.pair_output, store, multiseq.site_labels);  // WeightedLD.py:176,283
    else write_pair_stats(opt.pair_output, store, multiseq.site_labels);  // main.rs:209
    INFO("Finshed writing output in " + fmt_duration(Clock::now() - sw)); // main.rs:210 (sic)
    // Everything is written and closed.  Releasing tens of GB of device memory buffer by buffer and tearing the
    // CUDA context down in order takes longer than the whole computation; the process is about to end anyway.
    std::fflush(nullptr);
    std::_Exit(0);
  } catch (const Panic& p) {
    std::fprintf(stderr, "thread 'main' panicked at '%s'\n", p.what());
    return 101;  // Rust's panic exit code
  } catch (const std::exception& e) {
    std::fprintf(stderr, "Error: %s\n", e.what());                        // main.rs:121 (io::Error via `?`)
    return 1;
  }
}
