// wld_api.cu — the C ABI declared in include/wld.h (stage orchestration, ownership, errors,
// result ordering).  No torch types, no CPU fallback: every stage is a CUDA kernel in this library
// and every failure is reported through wld_status + wld_last_error.
#include <time.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <new>
#include <thread>

#include "common.cuh"
#include "pair_epilogue.cuh"

using namespace wld;

namespace {
#define WLD_CHECK_CTX(c)        \
  if (!(c)) return WLD_ERR_INVALID; \
  cudaSetDevice((c)->device)

bool host_is_pinned(const void* p);
int ensure_stagers(wld_ctx* c);
int staged_copy(wld_ctx* c, void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t row_bytes,
                size_t rows, cudaMemcpyKind dir, const uint8_t* const* src_rows = nullptr);
constexpr size_t kStageChunk = 4u << 20;  // pinning memory costs ~0.6 ms per MB on this driver: 16 chunks = 64 MB at most
constexpr size_t kStagedMin = 16u << 20;  // below this a plain copy is as fast

}  // namespace

extern "C" {

int wld_abi_version(void) { return WLD_ABI_VERSION; }

int wld_create(int device, wld_ctx** out) {
  if (!out) return WLD_ERR_INVALID;
  *out = nullptr;
  wld_ctx* c = new (std::nothrow) wld_ctx();
  if (!c) return WLD_ERR_NOMEM;
  *out = c;  // returned even on failure so that wld_last_error works; caller destroys it
  c->device = device;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return c->fail(WLD_ERR_CUDA, "no CUDA device available (%s); libwld has no CPU fallback",
                   e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return c->fail(WLD_ERR_INVALID, "device %d out of range (0..%d)", device, ndev - 1);
  cudaGetLastError();  // do not inherit a stale error from unrelated earlier calls
  WLD_CUDA(c, cudaSetDevice(device));
  cudaDeviceProp prop;
  WLD_CUDA(c, cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return c->fail(WLD_ERR_CUDA, "device %d is sm_%d%d; libwld is built for sm_100a (B200) only", device, prop.major,
                   prop.minor);
  c->sm_count = prop.multiProcessorCount;
  if (const char* e = std::getenv("WLD_SCREEN"))  // A/B of whole programs (the CLI): same meaning as wld_set_screen
    if (e[0] >= '0' && e[0] <= '2' && e[1] == 0) c->screen_opt = e[0] - '0';
  WLD_CUDA(c, cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  c->stream = c->own_stream;
  return WLD_OK;
}

void wld_destroy(wld_ctx* c) {
  if (!c) return;
  if (!c->own_stream) {  // creation failed before any device state existed
    delete c;
    return;
  }
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  DevBuf* bufs[] = {&c->raw_own, &c->hist, &c->keep, &c->rank, &c->maj_raw, &c->min_raw, &c->site_map, &c->maj,
                    &c->mnr, &c->kept_count, &c->codes, &c->table, &c->partial, &c->w64, &c->w32, &c->scalars,
                    &c->q, &c->limbs, &c->opA, &c->opB, &c->simt_tiles, &c->pairs, &c->counters, &c->py_aux, &c->die_of_sm,
                    &c->sorted, &c->sort_keys, &c->sort_idx, &c->sort_temp, &c->gain8, &c->quant, &c->glimb, &c->opB1,
                    &c->cand, &c->plans[0].tiles, &c->plans[1].tiles, &c->plans[2].tiles, &c->cell_flags,
                    &c->sample_flags};
  for (DevBuf* b : bufs) b->release();
  if (c->quant_host) cudaFreeHost(c->quant_host);
  if (c->stage_buf[0]) cudaFreeHost(c->stage_buf[0]);  // one allocation, sliced (ensure_stagers)
  for (int i = 0; i < wld_ctx::kMaxStagers; ++i)
    if (c->stage_stream[i]) cudaStreamDestroy(c->stage_stream[i]);
  for (auto& t : c->timers) {
    if (t.beg) cudaEventDestroy(t.beg);
    if (t.end) cudaEventDestroy(t.end);
  }
  if (c->poll_stream) cudaStreamDestroy(c->poll_stream);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c;
}

const char* wld_last_error(const wld_ctx* c) { return c ? c->err.c_str() : "null context"; }

int wld_set_stream(wld_ctx* c, void* cuda_stream) {
  WLD_CHECK_CTX(c);
  if (c->stream) cudaStreamSynchronize(c->stream);
  c->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : c->own_stream;
  return WLD_OK;
}

int wld_set_partition(wld_ctx* c, int part, int nparts) {
  WLD_CHECK_CTX(c);
  if (nparts < 1 || part < 0 || part >= nparts) return c->fail(WLD_ERR_INVALID, "bad partition %d/%d", part, nparts);
  c->part = part;
  c->nparts = nparts;
  return WLD_OK;
}

int wld_set_limbs(wld_ctx* c, int n_limbs) {
  WLD_CHECK_CTX(c);
  if (n_limbs < 0 || n_limbs > 4) return c->fail(WLD_ERR_INVALID, "n_limbs must be 1..4, or 0 for automatic");
  c->n_limbs_opt = n_limbs;
  return WLD_OK;
}

int wld_set_gain_bits(wld_ctx* c, int gain_bits) {
  WLD_CHECK_CTX(c);
  if (gain_bits < -1 || gain_bits > 7) return c->fail(WLD_ERR_INVALID, "gain_bits must be 0..7, or -1 for automatic");
  c->gain_opt = gain_bits;
  return WLD_OK;
}

int wld_set_limb_bits(wld_ctx* c, int limb_bits) {
  WLD_CHECK_CTX(c);
  if (limb_bits < 0 || limb_bits > 8) return c->fail(WLD_ERR_INVALID, "limb_bits must be 1..8, or 0 for automatic");
  c->limb_bits_opt = limb_bits;
  return WLD_OK;
}

int wld_set_pair_kernel(wld_ctx* c, int kind) {
  WLD_CHECK_CTX(c);
  if (kind != WLD_PAIR_KERNEL_UMMA && kind != WLD_PAIR_KERNEL_SIMT && kind != WLD_PAIR_KERNEL_UMMA_I8) return c->fail(WLD_ERR_INVALID, "unknown pair kernel %d", kind);
  c->pair_kernel = kind;
  return WLD_OK;
}

int wld_set_compat(wld_ctx* c, int mode) {
  WLD_CHECK_CTX(c);
  if (mode != WLD_COMPAT_RUST && mode != WLD_COMPAT_PYTHON) return c->fail(WLD_ERR_INVALID, "unknown compat mode %d", mode);
  c->compat = mode;
  return WLD_OK;
}

int wld_set_cta_group(wld_ctx* c, int ctas) {
  WLD_CHECK_CTX(c);
  if (ctas != 1 && ctas != 2) return c->fail(WLD_ERR_INVALID, "cta_group must be 1 or 2");
  c->cta_group = ctas;
  return WLD_OK;
}

int wld_set_screen(wld_ctx* c, int mode) {
  WLD_CHECK_CTX(c);
  if (mode < 0 || mode > 2) return c->fail(WLD_ERR_INVALID, "screen mode must be 0 (never), 1 (automatic) or 2 (always)");
  c->screen_opt = mode;
  return WLD_OK;
}

int wld_set_pair_capacity(wld_ctx* c, uint64_t pairs) {
  WLD_CHECK_CTX(c);
  c->pair_cap_opt = pairs;
  return WLD_OK;
}

// ---- stage 1 -----------------------------------------------------------------------------------
static int load_common(wld_ctx* c, const uint8_t* data, const uint8_t* const* rows, int64_t n_seqs, int64_t n_cols,
                       int64_t row_stride, int flags) {
  if (n_seqs >= (1ll << 31) || n_cols >= (1ll << 31) - 64)
    return c->fail(WLD_ERR_UNSUPPORTED, "alignment dimensions must be below 2^31");
  c->stage = Stage::Created;
  c->n_seqs = n_seqs;
  c->n_cols = n_cols;
  c->input_flags = flags;
  c->n_kept = 0;
  {
    ScopedStageTimer tm(c, WLD_STAGE_LOAD);
    if (flags & WLD_INPUT_DEVICE) {
      c->d_raw = data;
      c->row_stride = row_stride;
    } else {
      const int64_t pitch = round_up(std::max<int64_t>(n_cols, 1), 16);
      const size_t bytes = (size_t)pitch * (size_t)std::max<int64_t>(n_seqs, 1);
      WLD_CUDA(c, c->raw_own.ensure(bytes));
      if (pitch != n_cols) WLD_CUDA(c, cudaMemsetAsync(c->raw_own.p, 0, bytes, c->stream));
      const bool big = (size_t)n_seqs * (size_t)n_cols >= kStagedMin && (size_t)n_cols <= kStageChunk;
      if (n_seqs > 0 && n_cols > 0 && (rows || (big && !host_is_pinned(data)))) {
        // rows by pointer, or a large pageable source (a Rust Vec<u8>, a numpy array, a mapped file): several host
        // threads stage it through pinned buffers, each chunk's H2D copy overlapping the others' memcpy
        if ((size_t)n_cols > kStageChunk) return c->fail(WLD_ERR_UNSUPPORTED, "rows longer than %zu bytes must be passed contiguously", kStageChunk);
        WLD_CUDA(c, cudaStreamSynchronize(c->stream));
        const auto t0 = std::chrono::steady_clock::now();
        int rc = ensure_stagers(c);
        const auto t1 = std::chrono::steady_clock::now();
        if (rc == WLD_OK)
          rc = staged_copy(c, c->raw_own.p, (size_t)pitch, data, (size_t)row_stride, (size_t)n_cols, (size_t)n_seqs,
                           cudaMemcpyHostToDevice, rows);
        if (rc != WLD_OK) return rc;
        if (std::getenv("WLD_DEBUG")) {
          const double ms0 = std::chrono::duration<double, std::milli>(t1 - t0).count();
          const double ms1 = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count();
          std::fprintf(stderr, "[libwld] staged load: %d threads, pinned buffers %.1f ms, copy of %.1f MB %.1f ms (%.1f GB/s)\n",
                       c->n_stagers, ms0, (double)n_seqs * n_cols / 1e6, ms1, (double)n_seqs * n_cols / 1e6 / ms1);
        }
      } else if (n_seqs > 0 && n_cols > 0) {
        if (row_stride == pitch)  // same pitch on both sides: one linear copy (a pageable 2-D copy goes row by row)
          WLD_CUDA(c, cudaMemcpyAsync(c->raw_own.p, data, (size_t)(n_seqs - 1) * (size_t)pitch + (size_t)n_cols,
                                      cudaMemcpyHostToDevice, c->stream));
        else
          WLD_CUDA(c, cudaMemcpy2DAsync(c->raw_own.p, (size_t)pitch, data, (size_t)row_stride, (size_t)n_cols,
                                        (size_t)n_seqs, cudaMemcpyHostToDevice, c->stream));
      }
      c->d_raw = c->raw_own.as<uint8_t>();
      c->row_stride = pitch;
    }
  }
  {
    ScopedStageTimer tm(c, WLD_STAGE_HISTOGRAM);
    int rc = run_histogram(c, tm);
    if (rc != WLD_OK) return rc;
  }
  // The host buffer may be reused by the caller as soon as we return.  (A borrowed device buffer stays the
  // caller's promise until the next load, so its kernels are left running: one pipeline drain less per step.)
  if (!(flags & WLD_INPUT_DEVICE)) WLD_CUDA(c, cudaStreamSynchronize(c->stream));
  c->stage = Stage::Loaded;
  return WLD_OK;
}

int wld_load_alignment(wld_ctx* c, const uint8_t* data, int64_t n_seqs, int64_t n_cols, int64_t row_stride, int flags) {
  WLD_CHECK_CTX(c);
  if (n_seqs < 0 || n_cols < 0 || row_stride < n_cols || (!data && n_seqs * n_cols > 0))
    return c->fail(WLD_ERR_INVALID, "bad alignment shape %lld x %lld (stride %lld)", (long long)n_seqs, (long long)n_cols,
                   (long long)row_stride);
  return load_common(c, data, nullptr, n_seqs, n_cols, row_stride, flags);
}

int wld_load_alignment_rows(wld_ctx* c, const uint8_t* const* rows, int64_t n_seqs, int64_t n_cols, int flags) {
  WLD_CHECK_CTX(c);
  if (n_seqs < 0 || n_cols < 0 || (!rows && n_seqs > 0) || (flags & WLD_INPUT_DEVICE))
    return c->fail(WLD_ERR_INVALID, "bad alignment rows (%lld x %lld); rows must be host pointers", (long long)n_seqs, (long long)n_cols);
  for (int64_t r = 0; r < n_seqs && n_cols > 0; ++r)
    if (!rows[r]) return c->fail(WLD_ERR_INVALID, "row %lld is null", (long long)r);
  return load_common(c, nullptr, rows, n_seqs, n_cols, n_cols, flags);
}

static int filter_common(wld_ctx* c, int mode, float min_acgt, float min_minor, float max_minor, double py_min_acgt,
                         double py_min_variability, int64_t* n_kept) {
  if (c->stage < Stage::Loaded) return c->fail(WLD_ERR_STATE, "wld_filter_sites before wld_load_alignment");
  {
    ScopedStageTimer tm(c, WLD_STAGE_FILTER);
    int rc = run_filter(c, mode, min_acgt, min_minor, max_minor, py_min_acgt, py_min_variability, tm);
    if (rc != WLD_OK) return rc;
  }
  // (run_filter waited for the kept count; the gather that follows it keeps running)
  if (n_kept) *n_kept = c->n_kept;
  c->stage = Stage::Filtered;
  return WLD_OK;
}

int wld_filter_sites(wld_ctx* c, float min_acgt, float min_minor, float max_minor, int64_t* n_kept) {
  WLD_CHECK_CTX(c);
  return filter_common(c, 0, min_acgt, min_minor, max_minor, 0.0, 0.0, n_kept);
}

int wld_filter_sites_python(wld_ctx* c, double min_acgt, double min_variability, int64_t* n_kept) {
  WLD_CHECK_CTX(c);
  return filter_common(c, 2, 0.f, 0.f, 0.f, min_acgt, min_variability, n_kept);
}

int wld_keep_all_sites(wld_ctx* c, int64_t* n_kept) {
  WLD_CHECK_CTX(c);
  return filter_common(c, 1, 0.f, 0.f, 0.f, 0.0, 0.0, n_kept);
}

int64_t wld_n_seqs(const wld_ctx* c) { return c ? c->n_seqs : -1; }
int64_t wld_n_cols(const wld_ctx* c) { return c ? c->n_cols : -1; }
int64_t wld_n_kept(const wld_ctx* c) { return c ? c->n_kept : -1; }

int wld_get_site_map(wld_ctx* c, int64_t* out, int64_t cap) {
  WLD_CHECK_CTX(c);
  if (c->stage < Stage::Filtered) return c->fail(WLD_ERR_STATE, "site map requested before filtering");
  if (cap < c->n_kept || (!out && c->n_kept)) return c->fail(WLD_ERR_INVALID, "site map buffer too small");
  std::vector<int32_t> tmp((size_t)c->n_kept);
  if (c->n_kept)
    WLD_CUDA(c, cudaMemcpyAsync(tmp.data(), c->site_map.p, sizeof(int32_t) * tmp.size(), cudaMemcpyDeviceToHost, c->stream));
  WLD_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int64_t k = 0; k < c->n_kept; ++k) out[k] = tmp[(size_t)k];
  return WLD_OK;
}

int wld_get_histograms(wld_ctx* c, uint32_t* out, int64_t cap_cols) {
  WLD_CHECK_CTX(c);
  // bin 5 is completed by the decision kernel, so histograms are readable after filtering
  if (c->stage < Stage::Filtered) return c->fail(WLD_ERR_STATE, "histograms requested before filtering");
  if (cap_cols < c->n_cols || (!out && c->n_cols)) return c->fail(WLD_ERR_INVALID, "histogram buffer too small");
  std::vector<uint32_t> tmp((size_t)(6 * c->cols_padded));
  if (!tmp.empty())
    WLD_CUDA(c, cudaMemcpyAsync(tmp.data(), c->hist.p, sizeof(uint32_t) * tmp.size(), cudaMemcpyDeviceToHost, c->stream));
  WLD_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int64_t col = 0; col < c->n_cols; ++col)
    for (int k = 0; k < 6; ++k) out[col * 6 + k] = tmp[(size_t)(k * c->cols_padded + col)];
  return WLD_OK;
}

int wld_get_major_minor(wld_ctx* c, int8_t* major, int8_t* minor, int64_t cap) {
  WLD_CHECK_CTX(c);
  if (c->stage < Stage::Filtered) return c->fail(WLD_ERR_STATE, "major/minor requested before filtering");
  if (cap < c->n_kept) return c->fail(WLD_ERR_INVALID, "major/minor buffer too small");
  if (c->n_kept) {
    WLD_CUDA(c, cudaMemcpyAsync(major, c->maj.p, (size_t)c->n_kept, cudaMemcpyDeviceToHost, c->stream));
    WLD_CUDA(c, cudaMemcpyAsync(minor, c->mnr.p, (size_t)c->n_kept, cudaMemcpyDeviceToHost, c->stream));
  }
  WLD_CUDA(c, cudaStreamSynchronize(c->stream));
  return WLD_OK;
}

int wld_get_codes(wld_ctx* c, uint8_t* out, int64_t cap_bytes) {
  WLD_CHECK_CTX(c);
  if (c->stage < Stage::Filtered) return c->fail(WLD_ERR_STATE, "codes requested before filtering");
  if (cap_bytes < c->n_kept * c->n_seqs) return c->fail(WLD_ERR_INVALID, "code buffer too small");
  if (c->n_kept && c->n_seqs)
    WLD_CUDA(c, cudaMemcpy2DAsync(out, (size_t)c->n_seqs, c->codes.p, (size_t)c->ldc, (size_t)c->n_seqs,
                                  (size_t)c->n_kept, cudaMemcpyDeviceToHost, c->stream));
  WLD_CUDA(c, cudaStreamSynchronize(c->stream));
  return WLD_OK;
}

// ---- stage 2 -----------------------------------------------------------------------------------
int wld_henikoff(wld_ctx* c) {
  WLD_CHECK_CTX(c);
  if (c->stage < Stage::Filtered) return c->fail(WLD_ERR_STATE, "wld_henikoff before wld_filter_sites");
  {
    ScopedStageTimer tm(c, WLD_STAGE_HENIKOFF);
    const bool whole = c->seq_lo <= 0 && (c->seq_hi < 0 || c->seq_hi >= c->n_seqs);
    int rc = run_henikoff(c, tm, whole);
    if (rc != WLD_OK) return rc;
    c->weights_partial = !whole;
  }
  // a shard: the sums of the other sequences arrive by exchange, then wld_henikoff_finish normalises
  if (!c->weights_partial) c->stage = Stage::Weighted;  // no wait: the weights are read on the same stream
  return WLD_OK;
}

int wld_henikoff_finish(wld_ctx* c) {
  WLD_CHECK_CTX(c);
  if (c->stage < Stage::Filtered || !c->weights_partial)
    return c->fail(WLD_ERR_STATE, "wld_henikoff_finish without a sharded wld_henikoff before it");
  {
    ScopedStageTimer tm(c, WLD_STAGE_HENIKOFF);  // (the timer then shows the finishing kernels only)
    int rc = run_henikoff_finish(c, tm);
    if (rc != WLD_OK) return rc;
  }
  c->weights_partial = false;
  c->stage = Stage::Weighted;
  return WLD_OK;
}

int wld_set_row_shard(wld_ctx* c, int64_t row_lo, int64_t row_hi) {
  WLD_CHECK_CTX(c);
  if (row_lo < 0 || (row_hi >= 0 && row_hi < row_lo)) return c->fail(WLD_ERR_INVALID, "bad row shard");
  c->row_lo = row_lo;
  c->row_hi = row_hi;
  return WLD_OK;
}

int wld_set_seq_shard(wld_ctx* c, int64_t seq_lo, int64_t seq_hi) {
  WLD_CHECK_CTX(c);
  if (seq_lo < 0 || (seq_hi >= 0 && seq_hi < seq_lo)) return c->fail(WLD_ERR_INVALID, "bad sequence shard");
  c->seq_lo = seq_lo;
  c->seq_hi = seq_hi;
  return WLD_OK;
}

int wld_exchange_buffer(wld_ctx* c, int which, void** device_ptr, uint64_t* bytes) {
  WLD_CHECK_CTX(c);
  if (!device_ptr || !bytes) return c->fail(WLD_ERR_INVALID, "null output");
  *device_ptr = nullptr;
  *bytes = 0;
  if (which == WLD_EXCHANGE_HISTOGRAM) {
    if (c->stage < Stage::Loaded) return c->fail(WLD_ERR_STATE, "histogram buffer before wld_load_alignment");
    *device_ptr = c->hist.p;
    *bytes = sizeof(uint32_t) * 5 * (uint64_t)c->cols_padded;  // bins A C G T -; Unknown is derived from n_seqs
  } else if (which == WLD_EXCHANGE_WEIGHT_SUMS) {
    if (c->stage < Stage::Filtered || !c->w64.p) return c->fail(WLD_ERR_STATE, "weight sums before wld_henikoff");
    *device_ptr = c->w64.p;
    *bytes = sizeof(double) * (uint64_t)c->n_seqs;
  } else {
    return c->fail(WLD_ERR_INVALID, "unknown exchange buffer %d", which);
  }
  return WLD_OK;
}

// One process driving several GPUs: the exchanges by peer copies (sizes are below a megabyte).
namespace {
__global__ void add_u32_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}
}  // namespace

int wld_sum_histograms(wld_ctx* const* ctxs, int n) {
  if (!ctxs || n < 1 || !ctxs[0]) return WLD_ERR_INVALID;
  wld_ctx* root = ctxs[0];
  if (n == 1) return WLD_OK;
  cudaSetDevice(root->device);
  const size_t words = 5 * (size_t)root->cols_padded;
  for (int g = 0; g < n; ++g)
    if (!ctxs[g] || ctxs[g]->stage < Stage::Loaded || ctxs[g]->cols_padded != root->cols_padded)
      return root->fail(WLD_ERR_STATE, "wld_sum_histograms: every context must hold the same loaded alignment");
  DevBuf tmp;
  WLD_CUDA(root, tmp.ensure(sizeof(uint32_t) * std::max<size_t>(words, 1)));
  int rc = WLD_OK;
  for (int g = 1; g < n && rc == WLD_OK; ++g) {
    cudaSetDevice(ctxs[g]->device);
    cudaError_t e = cudaStreamSynchronize(ctxs[g]->stream);
    cudaSetDevice(root->device);
    if (e == cudaSuccess) e = cudaMemcpyPeerAsync(tmp.p, root->device, ctxs[g]->hist.p, ctxs[g]->device, sizeof(uint32_t) * words, root->stream);
    if (e != cudaSuccess) { rc = root->fail(WLD_ERR_CUDA, "histogram exchange failed: %s", cudaGetErrorString(e)); break; }
    if (words) add_u32_kernel<<<(unsigned)((words + 255) / 256), 256, 0, root->stream>>>(root->hist.as<uint32_t>(), tmp.as<uint32_t>(), words);
  }
  if (rc == WLD_OK && cudaStreamSynchronize(root->stream) != cudaSuccess) rc = root->fail(WLD_ERR_CUDA, "histogram exchange failed");
  for (int g = 1; g < n && rc == WLD_OK; ++g) {
    if (cudaMemcpyPeer(ctxs[g]->hist.p, ctxs[g]->device, root->hist.p, root->device, sizeof(uint32_t) * words) != cudaSuccess)
      rc = root->fail(WLD_ERR_CUDA, "histogram broadcast failed");
  }
  tmp.release();
  return rc;
}

int wld_share_weight_sums(wld_ctx* const* ctxs, int n) {
  if (!ctxs || n < 1 || !ctxs[0]) return WLD_ERR_INVALID;
  wld_ctx* root = ctxs[0];
  for (int g = 0; g < n; ++g)
    if (!ctxs[g] || !ctxs[g]->weights_partial || ctxs[g]->n_seqs != root->n_seqs)
      return root->fail(WLD_ERR_STATE, "wld_share_weight_sums: every context needs a sharded wld_henikoff first");
  for (int g = 0; g < n; ++g) {  // every source's stream is done with its slice
    cudaSetDevice(ctxs[g]->device);
    if (cudaStreamSynchronize(ctxs[g]->stream) != cudaSuccess) return root->fail(WLD_ERR_CUDA, "weight exchange failed");
  }
  for (int src = 0; src < n; ++src) {
    const wld_ctx* s_ = ctxs[src];
    const int64_t lo = std::min(s_->seq_lo, s_->n_seqs), hi = s_->seq_hi < 0 ? s_->n_seqs : std::min(s_->seq_hi, s_->n_seqs);
    if (hi <= lo) continue;
    for (int dst = 0; dst < n; ++dst) {
      if (dst == src) continue;
      if (cudaMemcpyPeer(ctxs[dst]->w64.as<double>() + lo, ctxs[dst]->device, s_->w64.as<double>() + lo, s_->device,
                         sizeof(double) * (size_t)(hi - lo)) != cudaSuccess)
        return root->fail(WLD_ERR_CUDA, "weight exchange failed");
    }
  }
  return WLD_OK;
}

int wld_set_weights(wld_ctx* c, const float* weights, int64_t n) {
  WLD_CHECK_CTX(c);
  if (c->stage < Stage::Filtered) return c->fail(WLD_ERR_STATE, "wld_set_weights before wld_filter_sites");
  if (n != c->n_seqs || (!weights && n)) return c->fail(WLD_ERR_INVALID, "expected %lld weights", (long long)c->n_seqs);
  WLD_CUDA(c, c->w32.ensure(sizeof(float) * (size_t)std::max<int64_t>(n, 1)));
  WLD_CUDA(c, c->w64.ensure(sizeof(double) * (size_t)std::max<int64_t>(n, 1)));
  std::vector<double> w64((size_t)n);
  for (int64_t i = 0; i < n; ++i) w64[(size_t)i] = (double)weights[i];
  if (n) {
    WLD_CUDA(c, cudaMemcpyAsync(c->w32.p, weights, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    WLD_CUDA(c, cudaMemcpyAsync(c->w64.p, w64.data(), sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  }
  WLD_CUDA(c, cudaStreamSynchronize(c->stream));
  c->stage = Stage::Weighted;
  return WLD_OK;
}

int wld_get_weights(wld_ctx* c, float* out, int64_t cap) {
  WLD_CHECK_CTX(c);
  if (c->stage < Stage::Weighted) return c->fail(WLD_ERR_STATE, "weights requested before wld_henikoff / wld_set_weights");
  if (cap < c->n_seqs) return c->fail(WLD_ERR_INVALID, "weight buffer too small");
  if (c->n_seqs) WLD_CUDA(c, cudaMemcpyAsync(out, c->w32.p, sizeof(float) * (size_t)c->n_seqs, cudaMemcpyDeviceToHost, c->stream));
  WLD_CUDA(c, cudaStreamSynchronize(c->stream));
  return WLD_OK;
}

int wld_get_weights_f64(wld_ctx* c, double* out, int64_t cap) {
  WLD_CHECK_CTX(c);
  if (c->stage < Stage::Weighted) return c->fail(WLD_ERR_STATE, "weights requested before wld_henikoff / wld_set_weights");
  if (cap < c->n_seqs) return c->fail(WLD_ERR_INVALID, "weight buffer too small");
  if (c->n_seqs) WLD_CUDA(c, cudaMemcpyAsync(out, c->w64.p, sizeof(double) * (size_t)c->n_seqs, cudaMemcpyDeviceToHost, c->stream));
  WLD_CUDA(c, cudaStreamSynchronize(c->stream));
  return WLD_OK;
}

// ---- stage 3 -----------------------------------------------------------------------------------
int wld_ld_pairs(wld_ctx* c, float r2_threshold, wld_progress_fn progress, void* user, uint64_t* n_survivors,
                 uint64_t* pairs_computed) {
  WLD_CHECK_CTX(c);
  if (c->stage < Stage::Weighted) return c->fail(WLD_ERR_STATE, "wld_ld_pairs before weights are set");
  if (progress) progress(0, user);  // lib.rs:584
  c->n_survivors = 0;
  c->pairs_computed = 0;
  c->sorted_key = -1;
  c->timers[WLD_STAGE_ORDER].valid = false;
  c->timers[WLD_STAGE_ORDER].launches = 0;
  c->last_thr = r2_threshold;
  c->info = wld_pair_info{};
  const int64_t L = c->n_kept;
  for (int st : {WLD_STAGE_PAIR_SAMPLE, WLD_STAGE_PAIR_REFINE}) {
    c->timers[st].valid = false;
    c->timers[st].launches = 0;
  }
  c->timers[WLD_STAGE_PAIR].carry_ms = 0.f;
  c->timers[WLD_STAGE_PAIR].carry_launches = 0;
  if (L >= 2 && c->n_seqs > 0) {
    // Screen + refine (wld_set_screen) is considered for the u8 tensor kernel, the Rust dialect and a positive
    // threshold, when the schedule is long enough for a sample to mean something (>= 4 waves of cells).
    const int64_t cells_1d = (L + 127) / 128;
    bool try_screen = c->screen_opt != 0 && c->pair_kernel == WLD_PAIR_KERNEL_UMMA_I8 && c->compat == WLD_COMPAT_RUST &&
                      ld_thr_lo(r2_threshold) > 0.0 &&
                      (c->screen_opt == 2 || cells_1d * (cells_1d + 1) / 2 / c->nparts >= 4 * (c->sm_count / c->cta_group));
    {
      ScopedStageTimer tm(c, WLD_STAGE_PAIR_PREP);
      int rc = run_pair_prep(c, tm, try_screen);
      if (rc != WLD_OK) return rc;
    }
    bool use_screen = false;
    bool expect_cells = false;  // the screen first, then the exact kernel on the cells it flags (LD confined to a minority of cells)
    if (try_screen) {
      if (c->screen_opt == 1) {
        int rc = run_pair_umma(c, r2_threshold, 2);  // the screen over ~2 waves of tiles spread over the list: count only
        if (rc != WLD_OK) return rc;
      }
      int rc = finish_quant(c);  // one synchronisation: the quantiser's decision and the sample's counters
      if (rc != WLD_OK) return rc;
      c->info.screen_top_min = c->quant_top_min;
      c->info.sample_candidates = (int64_t)c->sample_host[0];
      c->info.sample_pairs = (int64_t)c->sample_host[1];
      // the bound needs every nonzero weight to fill its top limb reasonably (kappa = 2/top_min + ...), and a second limb to drop
      const bool valid = c->geom.n_limbs >= 2 && c->geom.limb_bits == 8 && c->quant_top_min >= 32;
      // Costs in units of the exact kernel over every pair: the screen itself 1/limbs; then EITHER every candidate is
      // recomputed on its own (pair_refine.cu: ~10 ns at 10 000 sequences against 0.022 ns per pair and limb in the tensor
      // kernel, i.e. 455/limbs per unit of candidate rate) OR the exact kernel runs on the cells that hold candidates
      // (their share of the sampled tiles, plus a third for tiles clipped to cell borders, plus fixed costs).  The cheaper is
      // planned; if neither leaves a margin of 10 %, the exact kernel runs alone.
      // (WLD_SAMPLE_BLIND=1, tests: pretend the sample saw nothing, as it may on a heterogeneous input)
      const char* blind = std::getenv("WLD_SAMPLE_BLIND");
      const double limbs = (double)std::max(c->geom.n_limbs, 1);
      const double f_c = c->sample_host[1] ? (double)c->sample_host[0] / (double)c->sample_host[1] : 0.0;
      const double f_t = c->sample_tiles > 0 ? (double)c->sample_host[2] / (double)c->sample_tiles : 1.0;
      const double cost_pairs = 455.0 / limbs * f_c;
      const double cost_cells = 1.3 * f_t + 0.02;  // + the limb operand of the whole partition, a plan, a round trip
      if (valid && c->screen_opt == 2) {
        use_screen = true;
      } else if (valid && blind && blind[0] == '1') {
        use_screen = true;
      } else if (valid && 1.0 / limbs + std::min(cost_pairs, cost_cells) < 0.9) {
        use_screen = true;
        expect_cells = cost_cells < cost_pairs;
      }
      c->info.sample_tiles = c->sample_tiles;
      c->info.sample_tiles_flagged = (int64_t)c->sample_host[2];
    }
    if (c->pair_kernel != WLD_PAIR_KERNEL_SIMT && !use_screen) {
      ScopedStageTimer tm(c, WLD_STAGE_PAIR_PREP);  // (shows the limb expansion only when the screen was tried first)
      int rc = run_expand_limbs(c, tm, false);
      if (rc != WLD_OK) return rc;
    }
    if (c->pair_kernel != WLD_PAIR_KERNEL_SIMT) {
      int rc = ensure_tile_plan(c, use_screen ? 1 : 0);
      if (rc != WLD_OK) return rc;
    }
    const uint64_t total_pairs = (uint64_t)L * (uint64_t)(L - 1) / 2;
    uint64_t cap = c->pair_cap_opt;
    if (!cap) {
      // Room for EVERY pair when that is cheap (an eighth of the free memory, at most 8 GB = 4.3e8 survivors), so
      // that high-LD inputs never take the overflow path below (which re-runs the pair kernel); 2^24 otherwise.
      // (cudaMemGetInfo takes milliseconds on this driver: asked once per context.)
      if (!c->auto_cap_budget) {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) free_b = 0;
        c->auto_cap_budget = std::max<uint64_t>(1ull << 24, std::min<uint64_t>(free_b / 8, 8ull << 30) / sizeof(wld_pair));
      }
      cap = std::min<uint64_t>(total_pairs, c->auto_cap_budget);
    }
    cap = std::max<uint64_t>(cap, 1024);
    // the capacity is what the buffer holds, never a cached number (a failed growth leaves an empty buffer)
    auto ensure_pair_cap = [&](uint64_t want) -> int {
      c->pair_cap = c->pairs.p ? c->pairs.bytes / sizeof(wld_pair) : 0;
      if (c->pair_cap < want) {
        c->pair_cap = 0;
        WLD_CUDA(c, c->pairs.ensure(sizeof(wld_pair) * (size_t)want));
        c->pair_cap = c->pairs.bytes / sizeof(wld_pair);
      }
      return WLD_OK;
    };
    if (use_screen) {
      // candidates: four times the rate the screen is chosen at (1 in 64), between 2^16 and 2^26 (512 MB); grows like the survivors
      const uint64_t want = std::min<uint64_t>(std::max<uint64_t>(1ull << 16, c->plan_pairs / 64 + 1024), 1ull << 26);
      c->cand_cap = c->cand.p ? c->cand.bytes / sizeof(uint2) : 0;
      if (c->cand_cap < want) {
        c->cand_cap = 0;
        WLD_CUDA(c, c->cand.ensure(sizeof(uint2) * (size_t)want));
        c->cand_cap = c->cand.bytes / sizeof(uint2);
      }
      // every survivor is a candidate first: no need for a survivor buffer larger than the candidate buffer (a one-shot
      // run, the CLI, would otherwise pay for allocating gigabytes it cannot fill)
      int rc = ensure_pair_cap(c->pair_cap_opt ? cap : std::min<uint64_t>(cap, c->cand_cap));
      if (rc != WLD_OK) return rc;
    } else {
      int rc = ensure_pair_cap(cap);
      if (rc != WLD_OK) return rc;
    }
    if (c->compat == WLD_COMPAT_PYTHON) {
      int rc = run_pair_python_prepare(c);
      if (rc != WLD_OK) return rc;
    }
    unsigned long long progress_last = 0;
    bool cells_only = false;            // the exact kernel over the cells the screen flagged (plans[2])
    bool refine_only = false;           // the screen has run and its candidate list is complete: refine it, whatever its length
    uint64_t screened_pairs = 0;        // pairs the screen went through before that
    for (int attempt = 0; attempt < 4; ++attempt) {
      if (refine_only) {  // keep the screen's counters (candidates, pairs); clear the survivor count and the "declined" word
        WLD_CUDA(c, cudaMemsetAsync(c->counters.as<unsigned long long>(), 0, sizeof(unsigned long long), c->stream));
        WLD_CUDA(c, cudaMemsetAsync(c->counters.as<unsigned long long>() + 6, 0, sizeof(unsigned long long), c->stream));
      } else
        WLD_CUDA(c, cudaMemsetAsync(c->counters.p, 0, sizeof(unsigned long long) * 8, c->stream));
      c->die_used = 0;
      if (use_screen && !refine_only) {
        const size_t n_screen_tiles = (size_t)std::max<int64_t>(c->plans[1].n_tiles, 1);
        WLD_CUDA(c, c->cell_flags.ensure(n_screen_tiles));
        WLD_CUDA(c, cudaMemsetAsync(c->cell_flags.p, 0, n_screen_tiles, c->stream));
      }
      {
        int rc = refine_only ? WLD_OK
                 : c->pair_kernel == WLD_PAIR_KERNEL_SIMT ? run_pair_simt(c, r2_threshold)
                                                          : run_pair_umma(c, r2_threshold, use_screen ? 1 : cells_only ? 3 : 0);
        // Automatic mode: a short candidate list (recomputing it costs below 3 % of the exact kernel: 455/limbs per unit
        // of candidate rate, see above) is refined at once; with a longer one the refinement declines on the device and
        // the host chooses below, from the exact counts, how to finish.  (expect_cells: the sample already said so.)
        if (rc == WLD_OK && use_screen) {
          const uint64_t short_list = std::max<uint64_t>(1ull << 16, (uint64_t)((double)c->plan_pairs * 0.03 * c->geom.n_limbs / 455.0));
          rc = run_pair_refine(c, r2_threshold, (c->screen_opt == 2 || refine_only) ? ~0ull : expect_cells ? 0ull : short_list);
        }
        // pairs whose per-pair allele call may differ from the per-site call (WeightedLD.py:186-211)
        if (rc == WLD_OK && c->compat == WLD_COMPAT_PYTHON) rc = run_pair_python_fixup(c, r2_threshold);
        if (rc != WLD_OK) return rc;
      }
      if (progress) {
        // all_weighted_ld_pairs reports the running pair count while tiles finish (lib.rs:670-674).  The kernel
        // publishes its count every few tiles; it is polled here, on the caller's thread, over a second stream.
        if (!c->poll_stream) WLD_CUDA(c, cudaStreamCreateWithFlags(&c->poll_stream, cudaStreamNonBlocking));
        while (cudaStreamQuery(c->stream) == cudaErrorNotReady) {
          unsigned long long v = 0;
          if (cudaMemcpyAsync(&v, c->counters.as<unsigned long long>() + 1, sizeof v, cudaMemcpyDeviceToHost,
                              c->poll_stream) != cudaSuccess || cudaStreamSynchronize(c->poll_stream) != cudaSuccess)
            break;
          if (v > progress_last) {  // never decreases, also across a re-run of the stage
            progress(v, user);
            progress_last = v;
          }
          struct timespec ts = {0, 2000000};  // 2 ms
          nanosleep(&ts, nullptr);
        }
      }
      unsigned long long cnt[7] = {0, 0, 0, 0, 0, 0, 0};
      cudaError_t e = cudaMemcpyAsync(cnt, c->counters.p, sizeof cnt, cudaMemcpyDeviceToHost, c->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
      if (e != cudaSuccess)
        return c->fail(WLD_ERR_CUDA, "pair kernel failed: %s (pipeline watchdog code %d)", cudaGetErrorString(e),
                       (int)cnt[2]);
      const char* experiment = std::getenv("WLD_EXPERIMENT_SKIP_EPILOGUE");  // timing experiments count no pairs
      const uint64_t expected_pairs = cells_only ? c->plans[2].pairs : c->plan_pairs;
      if (c->die_used && ((int)cnt[2] == 7 || (cnt[1] != expected_pairs && !(experiment && experiment[0] == '1')))) {
        // the CTA pairs did not land on the dies as planned (GPU shared with other work?): plain schedule
        c->die_aware = 0;
        --attempt;
        continue;
      }
      c->pairs_computed = cells_only ? screened_pairs : cnt[1];
      if (use_screen) {
        c->info.screen_candidates = (int64_t)cnt[5];
        if (cnt[6]) {
          // The refinement declined: the candidate list is not short.  Three ways to finish, costed with the exact
          // numbers now at hand (units of the exact kernel over every pair): recompute the candidates one by one after
          // all; run the exact kernel on the cells in which the screen found a candidate (LD confined to a band or to
          // blocks: every other cell is ruled out by the screen's bound); or run it on all of this partition's pairs.
          std::vector<uint8_t> flags((size_t)c->plans[1].n_tiles);
          if (!flags.empty())
            WLD_CUDA(c, cudaMemcpyAsync(flags.data(), c->cell_flags.p, flags.size(), cudaMemcpyDeviceToHost, c->stream));
          WLD_CUDA(c, cudaStreamSynchronize(c->stream));
          int64_t n_flagged = 0;
          for (uint8_t f : flags) n_flagged += f != 0;
          c->info.screen_cells = c->plans[1].n_tiles;
          c->info.screen_cells_flagged = n_flagged;
          {
            const double limbs_d = (double)std::max(c->geom.n_limbs, 1);
            const double cost_pairs = cnt[5] <= c->cand_cap ? 455.0 / limbs_d * (double)cnt[5] / (double)std::max<uint64_t>(c->plan_pairs, 1) : 1e9;
            const double cost_cells = 1.3 * (double)n_flagged / (double)std::max<int64_t>(c->plans[1].n_tiles, 1) + 0.02;
            if (cost_pairs <= std::min(cost_cells, 1.0)) {
              refine_only = true;  // use_screen stays set: the survivors come from the candidate list
              --attempt;
              continue;
            }
          }
          {  // the screen's time stays part of the pair stage
            StageTimer& pt = c->timers[WLD_STAGE_PAIR];
            float ms = 0.f;
            if (pt.valid && cudaEventElapsedTime(&ms, pt.beg, pt.end) == cudaSuccess) {
              pt.carry_ms += ms;
              pt.carry_launches += pt.launches;
            }
          }
          use_screen = false;
          ScopedStageTimer tm(c, WLD_STAGE_PAIR_PREP);
          int rc = run_expand_limbs(c, tm, false);
          if (rc == WLD_OK) rc = ensure_tile_plan(c, 0);
          if (rc == WLD_OK && 2 * n_flagged <= c->plans[1].n_tiles) {
            rc = build_cell_plan(c, flags, &n_flagged);
            cells_only = rc == WLD_OK;
            screened_pairs = cnt[1];
            if (rc == WLD_OK) rc = ensure_pair_cap(std::min<uint64_t>(cap, std::max<uint64_t>(c->plans[2].pairs, 1024)));
          } else if (rc == WLD_OK) {
            rc = ensure_pair_cap(cap);  // the exact kernel's default capacity after all
          }
          if (rc != WLD_OK) return rc;
          --attempt;
          continue;
        }
        if (cnt[5] > c->cand_cap) {
          // the count is exact and nothing was written past the end: grow and repeat the screen
          const uint64_t want = cnt[5] + cnt[5] / 16 + 1024;
          c->cand_cap = 0;
          WLD_CUDA(c, c->cand.ensure(sizeof(uint2) * (size_t)want));
          c->cand_cap = c->cand.bytes / sizeof(uint2);
          if (!c->pair_cap_opt) {
            const int rc = ensure_pair_cap(std::min<uint64_t>(cap, c->cand_cap));
            if (rc != WLD_OK) return rc;
          }
          ++c->info.screen_reruns;
          if (attempt == 3) return c->fail(WLD_ERR_NOMEM, "candidate buffer kept overflowing");
          continue;
        }
      }
      if (cnt[0] <= c->pair_cap) {
        c->n_survivors = cnt[0];
        break;
      }
      // overflow protocol: the count is exact, the buffer was not written past its end -> grow, rerun
      const uint64_t want = cnt[0] + cnt[0] / 16 + 1024;
      c->pair_cap = 0;
      WLD_CUDA(c, c->pairs.ensure(sizeof(wld_pair) * (size_t)want));
      c->pair_cap = c->pairs.bytes / sizeof(wld_pair);
      if (attempt == 3) return c->fail(WLD_ERR_NOMEM, "survivor buffer kept overflowing");
    }
    c->info.screen = use_screen ? 1 : cells_only ? 2 : 0;
    if (std::getenv("WLD_DEBUG"))
      std::fprintf(stderr, "[libwld] pair stage of device %d: %s; sample %lld candidates in %lld pairs, screen %lld candidates in %llu pairs, top_min %d, %d limbs\n",
                   c->device, use_screen ? "one-limb screen + exact refinement" : cells_only ? "one-limb screen + exact kernel on the flagged cells" : (c->info.screen_candidates ? "screen gave up -> exact kernel" : "exact kernel"),
                   (long long)c->info.sample_candidates, (long long)c->info.sample_pairs, (long long)c->info.screen_candidates,
                   (unsigned long long)c->pairs_computed, c->quant_top_min, c->geom.n_limbs);
    c->info.kernel = c->pair_kernel;
    c->info.die_schedule = c->die_used;
    c->info.n_limbs = c->geom.n_limbs;
    c->info.limb_bits = c->geom.limb_bits;
    c->info.weight_bits = c->geom.n_limbs * c->geom.limb_bits;
    c->info.k_padded = c->geom.k_padded;
    c->info.gain_bits = c->geom.gain_bits;
    c->info.weight_span_log2 = c->quant_span_log2;
    c->info.weight_rel_err = c->quant_rel_err;
  }
  if (progress) progress(c->pairs_computed, user);
  if (n_survivors) *n_survivors = c->n_survivors;
  if (pairs_computed) *pairs_computed = c->pairs_computed;
  c->stage = Stage::Paired;
  return WLD_OK;
}

int wld_get_pair_weights(wld_ctx* c, double* out, int64_t cap) {
  WLD_CHECK_CTX(c);
  if (c->stage < Stage::Paired) return c->fail(WLD_ERR_STATE, "pair weights requested before wld_ld_pairs");
  if (cap < c->n_seqs || (!out && c->n_seqs)) return c->fail(WLD_ERR_INVALID, "weight buffer too small");
  if (c->n_kept < 2 || c->n_seqs == 0) return c->fail(WLD_ERR_STATE, "the last pair stage had nothing to compute");
  WLD_CUDA(c, cudaMemcpyAsync(out, c->q.p, sizeof(double) * (size_t)c->n_seqs, cudaMemcpyDeviceToHost, c->stream));
  WLD_CUDA(c, cudaStreamSynchronize(c->stream));
  return WLD_OK;
}

int wld_run(wld_ctx* c, const uint8_t* data, int64_t n_seqs, int64_t n_cols, int64_t row_stride, int flags, float min_acgt,
            float min_minor, float max_minor, const float* weights, float r2_threshold, int64_t* n_kept,
            uint64_t* n_survivors, uint64_t* pairs_computed) {
  int rc = wld_load_alignment(c, data, n_seqs, n_cols, row_stride, flags);
  if (rc == WLD_OK) rc = wld_filter_sites(c, min_acgt, min_minor, max_minor, n_kept);
  if (rc == WLD_OK) rc = weights ? wld_set_weights(c, weights, n_seqs) : wld_henikoff(c);
  if (rc == WLD_OK) rc = wld_ld_pairs(c, r2_threshold, nullptr, nullptr, n_survivors, pairs_computed);
  return rc;
}

uint64_t wld_pair_order_key(int64_t n_kept, uint32_t kept_a, uint32_t kept_b) {
  // lib.rs:615-632: n tiles of 256 per edge; linear tile index i -> row = n-1-root(i), col ascending:
  // tile rows bottom-up, columns left to right.  (n-1-row)*n + col is monotone in that order.
  const uint64_t n = (uint64_t)((n_kept + 255) / 256);
  const uint64_t tr = kept_a / 256, tc = kept_b / 256;
  return (n - 1 - tr) * n + tc;
}

int wld_plan_tiles(int64_t n_kept, int n_limbs, int cta_group, int part, int nparts, int sm_count, uint32_t* tiles_mn,
                   uint64_t cap_tiles, uint64_t* n_tiles, uint64_t* n_pairs) {
  if (cta_group != 1 && cta_group != 2) return WLD_ERR_INVALID;
  if (n_kept < 0 || n_limbs < 1 || n_limbs > 4 || nparts < 1 || part < 0 || part >= nparts) return WLD_ERR_INVALID;
  TilePlan plan = plan_tiles(n_kept, n_limbs, part, nparts, sm_count > 0 ? sm_count : kNumSMsB200, cta_group);
  if (n_tiles) *n_tiles = plan.tiles.size();
  if (n_pairs) *n_pairs = plan.pairs;
  if (tiles_mn) {
    if (cap_tiles < plan.tiles.size()) return WLD_ERR_INVALID;
    for (size_t i = 0; i < plan.tiles.size(); ++i) {
      tiles_mn[4 * i] = plan.tiles[i].x;
      tiles_mn[4 * i + 1] = plan.tiles[i].y;
      tiles_mn[4 * i + 2] = plan.tiles[i].z;
      tiles_mn[4 * i + 3] = plan.tiles[i].w;
    }
  }
  return WLD_OK;
}

int wld_plan_cell_tiles(int64_t n_kept, int n_limbs, int cta_group, int part, int nparts, const uint8_t* flags, uint64_t n_flags,
                        uint32_t* tiles, uint64_t cap_tiles, uint64_t* n_tiles, uint64_t* n_pairs) {
  if (cta_group != 1 && cta_group != 2) return WLD_ERR_INVALID;
  if (n_kept < 0 || n_limbs < 1 || n_limbs > 4 || nparts < 1 || part < 0 || part >= nparts || (!flags && n_flags)) return WLD_ERR_INVALID;
  const TilePlan screen = plan_tiles(n_kept, 1, part, nparts, kNumSMsB200, cta_group);
  const TilePlan plan = cut_cell_plan(screen.tiles, flags, (size_t)n_flags, n_kept, n_limbs, cta_group, nullptr);
  if (n_tiles) *n_tiles = plan.tiles.size();
  if (n_pairs) *n_pairs = plan.pairs;
  if (tiles) {
    if (cap_tiles < plan.tiles.size()) return WLD_ERR_INVALID;
    for (size_t i = 0; i < plan.tiles.size(); ++i) {
      tiles[4 * i] = plan.tiles[i].x;
      tiles[4 * i + 1] = plan.tiles[i].y;
      tiles[4 * i + 2] = plan.tiles[i].z;
      tiles[4 * i + 3] = plan.tiles[i].w;
    }
  }
  return WLD_OK;
}

// ---- host <-> device copies of large pageable buffers ------------------------------------------------
// A cudaMemcpy between device memory and PAGEABLE host memory (a Rust Vec<u8> / Vec<PairData>, a numpy array)
// is staged by the driver through a small pinned buffer at a fraction of the PCIe rate.  Large copies are
// therefore cut into chunks and moved by a few host threads, each with its own pinned staging buffer and stream:
// while one thread's chunk is on the bus, the others memcpy theirs between the staging buffer and the caller's
// memory.  Buffers that are already pinned (cudaHostAlloc / cudaHostRegister) take one plain async copy.
}  // extern "C"
namespace {

bool host_is_pinned(const void* p) {
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

int ensure_stagers(wld_ctx* c) {
  if (c->n_stagers) return WLD_OK;
  const unsigned hw = std::thread::hardware_concurrency();
  // half of the host threads (a memcpy into the pinned chunk runs at ~8 GB/s per thread; PCIe 5 x16 takes ~55 GB/s)
  int want = (int)std::min<unsigned>(wld_ctx::kMaxStagers, std::max(2u, hw / 2));
  // ONE pinned allocation for all chunks (each cudaMallocHost call costs milliseconds whatever its size)
  void* base = nullptr;
  while (want >= 1 && cudaMallocHost(&base, kStageChunk * (size_t)want) != cudaSuccess) {
    cudaGetLastError();
    base = nullptr;
    want /= 2;
  }
  if (!base) return c->fail(WLD_ERR_NOMEM, "cannot allocate pinned staging buffers");
  for (int i = 0; i < want; ++i) {
    if (cudaStreamCreateWithFlags(&c->stage_stream[i], cudaStreamNonBlocking) != cudaSuccess) {
      cudaGetLastError();
      break;
    }
    c->stage_buf[i] = static_cast<uint8_t*>(base) + kStageChunk * (size_t)i;
    ++c->n_stagers;
  }
  if (c->n_stagers < 1) {
    cudaFreeHost(base);
    c->stage_buf[0] = nullptr;
    return c->fail(WLD_ERR_NOMEM, "cannot create staging streams");
  }
  return WLD_OK;
}

// dir: cudaMemcpyDeviceToHost or cudaMemcpyHostToDevice.  Row form: `rows` rows of `row_bytes`, pitches in bytes
// (a linear copy is one row).  The context's stream must be idle with respect to the device buffer.
// src_rows (host -> device only): the source rows by pointer instead of base + pitch.
int staged_copy(wld_ctx* c, void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t row_bytes,
                size_t rows, cudaMemcpyKind dir, const uint8_t* const* src_rows) {
  int rc = ensure_stagers(c);
  if (rc != WLD_OK) return rc;
  const bool linear = !src_rows && (rows == 1 || (dst_pitch == row_bytes && src_pitch == row_bytes));
  // unit of work: a run of whole rows (or a byte range of the single row) of at most kStageChunk bytes
  const size_t total = linear ? row_bytes * rows : rows;
  const size_t unit = linear ? kStageChunk : std::max<size_t>(1, kStageChunk / std::max<size_t>(row_bytes, 1));
  if (!linear && row_bytes > kStageChunk) return c->fail(WLD_ERR_UNSUPPORTED, "row longer than the staging chunk");
  std::atomic<size_t> next{0};
  std::atomic<int> failed{0};
  const int device = c->device;
  auto worker = [&](int t) {
    cudaSetDevice(device);
    uint8_t* stage = static_cast<uint8_t*>(c->stage_buf[t]);
    cudaStream_t st = c->stage_stream[t];
    for (;;) {
      const size_t lo = next.fetch_add(unit);
      if (lo >= total || failed.load()) break;
      const size_t cnt = std::min(unit, total - lo);
      cudaError_t e = cudaSuccess;
      if (linear) {
        uint8_t* d = static_cast<uint8_t*>(dst) + lo;
        const uint8_t* s_ = static_cast<const uint8_t*>(src) + lo;
        if (dir == cudaMemcpyDeviceToHost) {
          e = cudaMemcpyAsync(stage, s_, cnt, dir, st);
          if (e == cudaSuccess) e = cudaStreamSynchronize(st);
          if (e == cudaSuccess) std::memcpy(d, stage, cnt);
        } else {
          std::memcpy(stage, s_, cnt);
          e = cudaMemcpyAsync(d, stage, cnt, dir, st);
          if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        }
      } else {
        uint8_t* d = static_cast<uint8_t*>(dst) + lo * dst_pitch;
        const uint8_t* s_ = static_cast<const uint8_t*>(src) + lo * src_pitch;
        if (dir == cudaMemcpyDeviceToHost) {
          e = cudaMemcpy2DAsync(stage, row_bytes, s_, src_pitch, row_bytes, cnt, dir, st);
          if (e == cudaSuccess) e = cudaStreamSynchronize(st);
          if (e == cudaSuccess)
            for (size_t r = 0; r < cnt; ++r) std::memcpy(d + r * dst_pitch, stage + r * row_bytes, row_bytes);
        } else {
          if (src_rows)
            for (size_t r = 0; r < cnt; ++r) std::memcpy(stage + r * row_bytes, src_rows[lo + r], row_bytes);
          else
            for (size_t r = 0; r < cnt; ++r) std::memcpy(stage + r * row_bytes, s_ + r * src_pitch, row_bytes);
          e = cudaMemcpy2DAsync(d, dst_pitch, stage, row_bytes, row_bytes, cnt, dir, st);
          if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        }
      }
      if (e != cudaSuccess) failed.store((int)e);
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < c->n_stagers; ++t) pool.emplace_back(worker, t);
  worker(0);
  for (auto& th : pool) th.join();
  if (failed.load())
    return c->fail(WLD_ERR_CUDA, "staged copy failed: %s", cudaGetErrorString((cudaError_t)failed.load()));
  return WLD_OK;
}

// device -> caller copy of `bytes` (after the stream's pending work), pageable-aware
int copy_out(wld_ctx* c, void* out, const void* dev, size_t bytes, bool out_is_device) {
  if (out_is_device) {
    WLD_CUDA(c, cudaMemcpyAsync(out, dev, bytes, cudaMemcpyDeviceToDevice, c->stream));
    WLD_CUDA(c, cudaStreamSynchronize(c->stream));
    return WLD_OK;
  }
  if (bytes < kStagedMin || host_is_pinned(out)) {
    WLD_CUDA(c, cudaMemcpyAsync(out, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
    WLD_CUDA(c, cudaStreamSynchronize(c->stream));
    return WLD_OK;
  }
  WLD_CUDA(c, cudaStreamSynchronize(c->stream));
  return staged_copy(c, out, bytes, dev, bytes, bytes, 1, cudaMemcpyDeviceToHost);
}

// Orders / maps the survivors as `flags` asks (cached until the next pair stage or append) and returns the
// device buffer that holds them.  *host_fallback is set when the scratch did not fit (caller merges on the host).
int order_for_fetch(wld_ctx* c, int flags, const wld_pair** dev, bool* host_fallback) {
  const bool ordered = !(flags & WLD_FETCH_UNORDERED), parent = !(flags & WLD_FETCH_KEPT_INDEX);
  *host_fallback = false;
  if (!ordered && !parent) {
    *dev = c->pairs.as<wld_pair>();
    return WLD_OK;
  }
  const int key = (ordered ? 1 : 0) | (parent ? 2 : 0);
  if (const char* e = std::getenv("WLD_FORCE_HOST_ORDER"))  // tests: take the no-device-memory path below
    if (e[0] == '1') {
      *host_fallback = true;
      return WLD_OK;
    }
  if (c->sorted_key != key) {
    c->sorted_key = -1;
    const int rc = run_pair_order(c, ordered, parent);
    if (rc == WLD_ERR_NOMEM) {
      *host_fallback = true;
      return WLD_OK;
    }
    if (rc != WLD_OK) return rc;
    c->sorted_key = key;
  }
  *dev = c->sorted.as<wld_pair>();
  return WLD_OK;
}

}  // namespace
extern "C" {

int wld_fetch_pairs_range(wld_ctx* c, uint64_t first, uint64_t count, wld_pair* out, int flags, uint64_t* n_written) {
  WLD_CHECK_CTX(c);
  if (c->stage < Stage::Paired) return c->fail(WLD_ERR_STATE, "wld_fetch_pairs before wld_ld_pairs");
  const uint64_t n = c->n_survivors;
  if (n_written) *n_written = 0;
  if (first > n) return c->fail(WLD_ERR_INVALID, "range starts at %llu, there are %llu pairs", (unsigned long long)first, (unsigned long long)n);
  count = std::min<uint64_t>(count, n - first);
  if (count == 0) return WLD_OK;
  if (!out) return c->fail(WLD_ERR_INVALID, "null output buffer");
  const bool ordered = !(flags & WLD_FETCH_UNORDERED), parent = !(flags & WLD_FETCH_KEPT_INDEX);
  const bool to_device = (flags & WLD_FETCH_DEVICE) != 0;
  const wld_pair* dev = nullptr;
  bool host_fallback = false;
  int rc = order_for_fetch(c, flags, &dev, &host_fallback);
  if (rc != WLD_OK) return rc;
  if (!host_fallback) {
    rc = copy_out(c, out, dev + first, sizeof(wld_pair) * (size_t)count, to_device);
    if (rc == WLD_OK && n_written) *n_written = count;
    return rc;
  }
  // Not enough device memory for the ordering scratch: merge on the host (same order, slower; whole set only).
  if (to_device || first != 0 || count != n)
    return c->fail(WLD_ERR_NOMEM, "no device memory to order %llu pairs (fetch them whole into host memory)", (unsigned long long)n);
  rc = copy_out(c, out, c->pairs.p, sizeof(wld_pair) * (size_t)n, false);
  if (rc != WLD_OK) return rc;
  std::vector<int32_t> smap;
  if (parent) {
    smap.resize((size_t)c->n_kept);
    WLD_CUDA(c, cudaMemcpyAsync(smap.data(), c->site_map.p, sizeof(int32_t) * smap.size(), cudaMemcpyDeviceToHost, c->stream));
    WLD_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  if (ordered) {
    const int64_t nk = c->n_kept;
    std::sort(out, out + n, [nk](const wld_pair& x, const wld_pair& y) {
      const uint64_t kx = wld_pair_order_key(nk, x.site_a, x.site_b), ky = wld_pair_order_key(nk, y.site_a, y.site_b);
      if (kx != ky) return kx < ky;
      if (x.site_a != y.site_a) return x.site_a < y.site_a;
      return x.site_b < y.site_b;
    });
  }
  if (parent)
    for (uint64_t i = 0; i < n; ++i) {
      out[i].site_a = (uint32_t)smap[out[i].site_a];  // lib.rs:662-663
      out[i].site_b = (uint32_t)smap[out[i].site_b];
    }
  if (n_written) *n_written = n;
  return WLD_OK;
}

int wld_fetch_pairs(wld_ctx* c, wld_pair* out, uint64_t cap, int flags, uint64_t* n_written) {
  WLD_CHECK_CTX(c);
  if (c->stage < Stage::Paired) return c->fail(WLD_ERR_STATE, "wld_fetch_pairs before wld_ld_pairs");
  if (n_written) *n_written = 0;
  if (cap < c->n_survivors)
    return c->fail(WLD_ERR_INVALID, "pair buffer holds %llu, need %llu", (unsigned long long)cap, (unsigned long long)c->n_survivors);
  return wld_fetch_pairs_range(c, 0, c->n_survivors, out, flags, n_written);
}

int wld_append_pairs(wld_ctx* c, const wld_pair* src, uint64_t n, int src_is_device) {
  WLD_CHECK_CTX(c);
  if (c->stage < Stage::Paired) return c->fail(WLD_ERR_STATE, "wld_append_pairs before wld_ld_pairs");
  if (n == 0) return WLD_OK;
  if (!src) return c->fail(WLD_ERR_INVALID, "null shard");
  const uint64_t have = c->n_survivors, want = have + n;
  const uint64_t cap = c->pairs.p ? c->pairs.bytes / sizeof(wld_pair) : 0;
  if (cap < want) {  // grow, keeping this context's own survivors
    DevBuf bigger;
    WLD_CUDA(c, bigger.ensure(sizeof(wld_pair) * (size_t)(want + want / 8)));
    if (have) WLD_CUDA(c, cudaMemcpyAsync(bigger.p, c->pairs.p, sizeof(wld_pair) * (size_t)have, cudaMemcpyDeviceToDevice, c->stream));
    WLD_CUDA(c, cudaStreamSynchronize(c->stream));
    c->pairs.release();
    c->pairs = bigger;
    c->pair_cap = c->pairs.bytes / sizeof(wld_pair);
  }
  wld_pair* dst = c->pairs.as<wld_pair>() + have;
  if (src_is_device) {
    WLD_CUDA(c, cudaMemcpyAsync(dst, src, sizeof(wld_pair) * (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
    WLD_CUDA(c, cudaStreamSynchronize(c->stream));  // the caller may release or reuse `src` on return
  } else if (sizeof(wld_pair) * n < kStagedMin || host_is_pinned(src)) {
    WLD_CUDA(c, cudaMemcpyAsync(dst, src, sizeof(wld_pair) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    WLD_CUDA(c, cudaStreamSynchronize(c->stream));
  } else {
    WLD_CUDA(c, cudaStreamSynchronize(c->stream));
    const int rc = staged_copy(c, dst, sizeof(wld_pair) * (size_t)n, src, sizeof(wld_pair) * (size_t)n,
                               sizeof(wld_pair) * (size_t)n, 1, cudaMemcpyHostToDevice);
    if (rc != WLD_OK) return rc;
  }
  c->n_survivors = want;
  c->sorted_key = -1;
  return WLD_OK;
}

int wld_append_pairs_from(wld_ctx* c, wld_ctx* other) {
  WLD_CHECK_CTX(c);
  if (!other || other == c) return c->fail(WLD_ERR_INVALID, "wld_append_pairs_from needs another context");
  if (c->stage < Stage::Paired || other->stage < Stage::Paired)
    return c->fail(WLD_ERR_STATE, "wld_append_pairs_from before wld_ld_pairs on both contexts");
  if (c->n_kept != other->n_kept || c->n_seqs != other->n_seqs)
    return c->fail(WLD_ERR_INVALID, "the contexts hold different site sets");
  const uint64_t n = other->n_survivors;
  if (n == 0) return WLD_OK;
  cudaSetDevice(other->device);
  cudaError_t e = cudaStreamSynchronize(other->stream);  // its survivors are complete
  cudaSetDevice(c->device);
  if (e != cudaSuccess) return c->fail(WLD_ERR_CUDA, "peer context failed: %s", cudaGetErrorString(e));
  const uint64_t have = c->n_survivors, want = have + n;
  const uint64_t cap = c->pairs.p ? c->pairs.bytes / sizeof(wld_pair) : 0;
  if (cap < want) {
    DevBuf bigger;
    WLD_CUDA(c, bigger.ensure(sizeof(wld_pair) * (size_t)(want + want / 8)));
    if (have) WLD_CUDA(c, cudaMemcpyAsync(bigger.p, c->pairs.p, sizeof(wld_pair) * (size_t)have, cudaMemcpyDeviceToDevice, c->stream));
    WLD_CUDA(c, cudaStreamSynchronize(c->stream));
    c->pairs.release();
    c->pairs = bigger;
    c->pair_cap = c->pairs.bytes / sizeof(wld_pair);
  }
  // NVLink / PCIe peer copy (the runtime stages through the host when peer access is not available)
  WLD_CUDA(c, cudaMemcpyPeerAsync(c->pairs.as<wld_pair>() + have, c->device, other->pairs.p, other->device,
                                  sizeof(wld_pair) * (size_t)n, c->stream));
  WLD_CUDA(c, cudaStreamSynchronize(c->stream));
  c->n_survivors = want;
  c->sorted_key = -1;
  return WLD_OK;
}

// ---- introspection -----------------------------------------------------------------------------
int wld_stage_ms(wld_ctx* c, int stage, float* ms) {
  WLD_CHECK_CTX(c);
  if (stage < 0 || stage >= WLD_STAGE_COUNT || !ms) return c->fail(WLD_ERR_INVALID, "bad stage id");
  *ms = 0.f;
  StageTimer& t = c->timers[stage];
  if (!t.valid) return WLD_OK;
  WLD_CUDA(c, cudaEventSynchronize(t.end));
  WLD_CUDA(c, cudaEventElapsedTime(ms, t.beg, t.end));
  *ms += t.carry_ms;
  return WLD_OK;
}

int wld_stage_launches(wld_ctx* c, int stage, int* launches) {
  WLD_CHECK_CTX(c);
  if (stage < 0 || stage >= WLD_STAGE_COUNT || !launches) return c->fail(WLD_ERR_INVALID, "bad stage id");
  *launches = c->timers[stage].valid ? c->timers[stage].launches + c->timers[stage].carry_launches : 0;
  return WLD_OK;
}

int wld_get_pair_info(wld_ctx* c, wld_pair_info* out) {
  WLD_CHECK_CTX(c);
  if (!out) return c->fail(WLD_ERR_INVALID, "null output");
  *out = c->info;
  return WLD_OK;
}

}  // extern "C"
