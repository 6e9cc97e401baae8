// fm_api.cu -- C ABI of libfrogmatch.so (include/frogmatch.h): context, keypoint upload,
// task batching and kernel orchestration.  sm_100a only; there is no CPU fallback.
#include <cstdio>
#include <cstring>
#include <new>

#include "../../include/frogmatch.h"
#include "../../include/frogmatch_debug.h"
#include "fm_all.cuh"
#include "fm_compact.cuh"
#include "fm_exact.cuh"
#include "fm_fast.cuh"
#include "fm_generic.cuh"
#include "fm_host.h"
#include "fm_links.cuh"

using namespace fm;

struct fm_result {
  fm_ctx* ctx = nullptr;
  size_t n_pairs = 0;
  uint64_t total = 0;
  uint32_t flags = 0;
  float ratio = 1.f;
  std::vector<uint32_t> counts;
  std::vector<uint64_t> offsets;
  DevBuf d_out, d_counts, d_dist;
  uint32_t* h_pairs = nullptr;  // pinned
  size_t h_cap = 0;
  float* h_dist = nullptr;  // pinned, FM_FLAG_DISTANCES only
  bool fetched = false;
  // completion: the call's DeviceCounters and per-pair counts land in a pinned block, `done` fires behind them
  void* h_block = nullptr;
  size_t h_block_cap = 0;
  cudaEvent_t done = nullptr;
  bool waited = false;
  EventPool ev;     // this call's phase brackets (and `done`)
  fm_stats stats{};  // launch-side counters; device counters and event times are filled in by finalize()
};

struct fm_links {
  fm_ctx* ctx = nullptr;
  uint64_t total = 0;        // half-links
  uint32_t n_points = 0;     // points of all images of the context
  std::vector<uint64_t> point_base;  // per image index: global id of its point 0 (n_images + 1 entries)
  DevBuf d_off, d_links;
  unsigned long long* h_off = nullptr;  // pinned, after fm_links_fetch
  uint32_t* h_links = nullptr;
  float ms_build = 0.f;
};

namespace {

thread_local std::string g_create_error;

int fail(fm_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg; else g_create_error = msg;
  return code;
}

int cuda_fail(fm_ctx* c, cudaError_t e, const char* what) {
  return fail(c, e == cudaErrorMemoryAllocation ? FM_ERR_NOMEM : FM_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define FM_CUDA(ctx, expr)                                     \
  do {                                                         \
    cudaError_t e__ = (expr);                                  \
    if (e__ != cudaSuccess) return cuda_fail((ctx), e__, #expr); \
  } while (0)

// Upload the ImageDev table, prepare the images uploaded since the last call (one batch of
// kernels) and read the per-image metas (flags, class tables) back.
int sync_images(fm_ctx* c) {
  if (!c->images_dirty) return FM_OK;
  const size_t n = c->h_images.size();
  FM_CUDA(c, c->d_images.ensure(std::max<size_t>(1, n) * sizeof(ImageDev)));
  c->h_metas.assign(n, ImageMeta{});
  if (n) {
    FM_CUDA(c, cudaMemcpyAsync(c->d_images.p, c->h_images.data(), n * sizeof(ImageDev), cudaMemcpyHostToDevice, c->stream));
    {
      Span sp(&c->ev_prep, c->stream, kPhPrep);
      cudaError_t e = fast_prepare_dirty(c);
      if (e != cudaSuccess) return cuda_fail(c, e, "image preparation");
    }
    FM_CUDA(c, cudaMemcpyAsync(c->h_metas.data(), c->d_metas.p, n * sizeof(ImageMeta), cudaMemcpyDeviceToHost, c->stream));
    FM_CUDA(c, cudaStreamSynchronize(c->stream));
    // the bracket above has completed: fold it into the running sum and recycle its events (a long-lived context that
    // re-uploads images without ever clearing must not accumulate events)
    for (auto& sp : c->ev_prep.spans) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) c->ms_prep_acc += ms;
    }
    c->ev_prep.reset();
  }
  c->images_dirty = false;
  return FM_OK;
}

// Small free lists so that several results can be in flight (FM_FLAG_ASYNC) without a cudaMalloc /
// cudaFree (a device-wide synchronisation) per call.
DevBuf take_buf(std::vector<DevBuf>& list, size_t bytes) {
  int best = -1;
  for (size_t i = 0; i < list.size(); i++)
    if (list[i].cap >= bytes && (best < 0 || list[i].cap < list[best].cap)) best = (int)i;
  if (best < 0 && !list.empty()) {  // none fits: grow the largest
    best = 0;
    for (size_t i = 1; i < list.size(); i++)
      if (list[i].cap > list[best].cap) best = (int)i;
  }
  DevBuf b;
  if (best >= 0) {
    b = list[best];
    list.erase(list.begin() + best);
  }
  return b;
}

void give_buf(std::vector<DevBuf>& list, DevBuf& b) {
  if (!b.p) return;
  if (list.size() < 4) list.push_back(b);
  else cudaFree(b.p);
  b = DevBuf{};
}

cudaError_t take_pinned(fm_ctx* c, size_t bytes, void** out, size_t* cap) {
  for (size_t i = 0; i < c->pin_free.size(); i++)
    if (c->pin_free[i].second >= bytes) {
      *out = c->pin_free[i].first;
      *cap = c->pin_free[i].second;
      c->pin_free.erase(c->pin_free.begin() + i);
      return cudaSuccess;
    }
  const size_t want = std::max<size_t>(bytes + bytes / 4, 4096);
  cudaError_t e = cudaMallocHost(out, want);
  if (e == cudaSuccess) *cap = want;
  return e;
}

void give_pinned(fm_ctx* c, void* p, size_t cap) {
  if (!p) return;
  if (c->pin_free.size() < 8) c->pin_free.emplace_back(p, cap);
  else cudaFreeHost(p);
}

// Complete a result: wait for its `done` event, then read the totals, per-pair counts and phase
// times the device left in the pinned block / the call's events.
int finalize(fm_result* r) {
  if (r->waited) return FM_OK;
  fm_ctx* c = r->ctx;
  FM_CUDA(c, cudaSetDevice(c->device));
  FM_CUDA(c, cudaEventSynchronize(r->done));
  const DeviceCounters* hc = static_cast<const DeviceCounters*>(r->h_block);
  const uint32_t* hcounts = reinterpret_cast<const uint32_t*>(static_cast<const char*>(r->h_block) + sizeof(DeviceCounters));
  r->total = hc->running_total;
  r->counts.assign(hcounts, hcounts + r->n_pairs);
  r->offsets.assign(r->n_pairs + 1, 0);
  for (size_t p = 0; p < r->n_pairs; p++) r->offsets[p + 1] = r->offsets[p] + r->counts[p];
  r->stats.scored_pairs = hc->scored_cols;
  r->stats.candidates = hc->rescore.candidates;
  r->stats.rows_rejected_early = hc->rescore.rejected_early;
  if (r->ratio < 1.f && r->stats.rows > r->stats.rows_exact)
    c->reject_hint = (double)hc->rescore.rejected_early / (double)(r->stats.rows - r->stats.rows_exact);
  r->stats.rows_exact += hc->redo_total;
  float acc[kNumPhases] = {0};
  for (auto& s : r->ev.spans) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) acc[s.phase] += ms;
  }
  r->stats.ms_total = acc[kPhTotal];
  r->stats.ms_score = acc[kPhScore];
  r->stats.ms_rescore = acc[kPhRescore] + acc[kPhBands];
  r->stats.ms_exact = acc[kPhExact];
  r->stats.ms_compact = acc[kPhCompact];
  r->waited = true;
  const float ms_prep = c->stats.ms_prep;
  c->stats = r->stats;
  c->stats.ms_prep = ms_prep;
  if (r->offsets[r->n_pairs] != r->total)
    return fail(c, FM_ERR_CUDA, "fm_match: internal error, per-pair counts do not sum to the compacted total");
  return FM_OK;
}

struct Batch {
  uint32_t t0, t1, rows, blocks, chunks, units;
  size_t cursor;  // first entry of this batch in the concatenated prefix arrays
  bool any_exact, any_fast;
};

}  // namespace

extern "C" {

const char* fm_version(void) { return "frogmatch 0.1 (sm_100a: tcgen05 + TMA bulk copy)"; }

int fm_device_count(int* n) {
  if (!n) return FM_ERR_INVALID;
  *n = 0;
  int k = 0;
  cudaError_t e = cudaGetDeviceCount(&k);
  if (e != cudaSuccess) return cuda_fail(nullptr, e, "fm_device_count");
  *n = k;
  return FM_OK;
}

int fm_create(int device, fm_ctx** out) {
  if (!out) return fail(nullptr, FM_ERR_INVALID, "fm_create: out is NULL");
  *out = nullptr;
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0)
    return fail(nullptr, FM_ERR_CUDA, std::string("fm_create: no CUDA device (") + cudaGetErrorString(e) +
                                          "); libfrogmatch has no CPU fallback");
  if (device < 0 || device >= n_dev) return fail(nullptr, FM_ERR_INVALID, "fm_create: device index out of range");
  cudaDeviceProp prop{};
  if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
    return cuda_fail(nullptr, e, "fm_create");
  if (prop.major != 10)
    return fail(nullptr, FM_ERR_UNSUPPORTED, "fm_create: device is sm_" + std::to_string(prop.major) +
                                                 std::to_string(prop.minor) + ", this library is built for sm_100a only");
  fm_ctx* c = new (std::nothrow) fm_ctx();
  if (!c) return fail(nullptr, FM_ERR_NOMEM, "fm_create: out of host memory");
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  if ((e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaMallocHost(reinterpret_cast<void**>(&c->h_pinned), 256)) != cudaSuccess) {
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
    return cuda_fail(nullptr, e, "fm_create");
  }
  c->stream = c->own_stream;
  *out = c;
  return FM_OK;
}

void fm_destroy(fm_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  c->arena.release();
  DevBuf* bufs[] = {&c->s_keys, &c->s_keys_sorted, &c->s_idx, &c->s_idx_sorted, &c->s_norm2, &c->s_sort, &c->s_segs, &c->d_images, &c->d_metas, &c->d_meta_blob, &c->d_rowres, &c->d_rowdist, &c->d_chunk_count, &c->d_chunk_out, &c->d_chunk_stage, &c->d_chunk_status,
                    &c->d_totals, &c->d_bands, &c->d_cands, &c->d_redo, &c->d_taskinfo, &c->d_rowstat, &c->d_surv, &c->d_all, &c->d_all_tasks};
  for (auto* b : bufs) b->release();
  for (auto& b : c->out_free) b.release();
  for (auto& b : c->counts_free) b.release();
  for (auto& b : c->dist_free) b.release();
  for (auto& pb : c->pin_free) cudaFreeHost(pb.first);
  if (c->cache_pinned) cudaFreeHost(c->cache_pinned);
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  c->ev_match.destroy();
  for (auto& e : c->ev_free) e.destroy();
  c->ev_prep.destroy();
  cudaStreamDestroy(c->own_stream);
  delete c;
}

const char* fm_last_error(const fm_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int fm_set_stream(fm_ctx* c, void* s) {
  if (!c) return FM_ERR_INVALID;
  FM_CUDA(c, cudaSetDevice(c->device));
  FM_CUDA(c, cudaStreamSynchronize(c->stream));
  c->stream = s ? reinterpret_cast<cudaStream_t>(s) : c->own_stream;
  return FM_OK;
}

int fm_synchronize(fm_ctx* c) {
  if (!c) return FM_ERR_INVALID;
  FM_CUDA(c, cudaSetDevice(c->device));
  FM_CUDA(c, cudaStreamSynchronize(c->stream));
  return FM_OK;
}

int fm_clear_images(fm_ctx* c) {
  if (!c) return FM_ERR_INVALID;
  for (auto& im : c->images) { im.valid = false; im.slab = nullptr; im.slab_bytes = 0; }
  c->dirty.clear();
  c->images_dirty = true;
  c->dim = 0;
  // the arena is handed out again right away: everything that still reads the old tensors must have finished
  FM_CUDA(c, cudaSetDevice(c->device));
  FM_CUDA(c, cudaStreamSynchronize(c->stream));
  c->arena.reset();
  c->ev_prep.reset();
  c->ms_prep_acc = 0.f;
  return FM_OK;
}

int fm_image_points(const fm_ctx* c, uint32_t img, uint32_t* n) {
  if (!c || !n || img >= c->images.size() || !c->images[img].valid) return FM_ERR_INVALID;
  *n = c->images[img].n;
  return FM_OK;
}

int fm_upload_image(fm_ctx* c, uint32_t img, const float* desc, const float* scale, const float* lap, uint32_t n,
                    uint32_t d) {
  if (!c) return FM_ERR_INVALID;
  if ((n && (!desc || !scale || !lap)) || d == 0) return fail(c, FM_ERR_INVALID, "fm_upload_image: null input or d == 0");
  if (d > (1u << 20)) return fail(c, FM_ERR_UNSUPPORTED, "fm_upload_image: descriptor length above 2^20 is not supported");
  if (img >= 65536u) return fail(c, FM_ERR_INVALID, "fm_upload_image: image index above 65535 (pairs.bin ids are u16, match.cpp:735)");
  if (c->dim != 0 && c->dim != d) return fail(c, FM_ERR_INVALID, "fm_upload_image: all images of a context must share d");
  FM_CUDA(c, cudaSetDevice(c->device));
  c->dim = d;
  if (img >= c->images.size()) c->images.resize(img + 1);
  if (img >= c->h_images.size()) c->h_images.resize(img + 1, ImageDev{});
  Image& im = c->images[img];
  // One slab per image: desc | scale | lap | perm | scale_sorted | norm2_sorted | rowop | colop (1 KB aligned pieces).
  auto pad = [](size_t x) { return (x + 1023) & ~(size_t)1023; };
  const uint32_t n_pad = (n + 255u) & ~255u;  // rows are consumed 256 at a time, columns 64
  const size_t b_desc = pad((size_t)n * d * sizeof(float)), b_vec = pad((size_t)std::max(n, 1u) * sizeof(float));
  const size_t b_op = d == (uint32_t)kD ? pad((size_t)n_pad * kKPad * sizeof(__half)) : 0;
  const size_t need = b_desc + 5 * b_vec + 2 * b_op + 1024;
  if (!im.slab || im.slab_bytes < need) {  // a re-upload that fits reuses the image's slab
    void* p = nullptr;
    FM_CUDA(c, c->arena.alloc(need, &p));
    im.slab = p;
    im.slab_bytes = need;
  }
  char* base = static_cast<char*>(im.slab);
  float* d_desc = reinterpret_cast<float*>(base);
  float* d_scale = reinterpret_cast<float*>(base + b_desc);
  float* d_lap = reinterpret_cast<float*>(base + b_desc + b_vec);
  if (n) {
    FM_CUDA(c, cudaMemcpyAsync(d_desc, desc, (size_t)n * d * sizeof(float), cudaMemcpyDefault, c->stream));
    FM_CUDA(c, cudaMemcpyAsync(d_scale, scale, (size_t)n * sizeof(float), cudaMemcpyDefault, c->stream));
    FM_CUDA(c, cudaMemcpyAsync(d_lap, lap, (size_t)n * sizeof(float), cudaMemcpyDefault, c->stream));
  }
  im.valid = true;
  im.n = n;
  im.d = d;
  ImageDev& v = c->h_images[img];
  v = ImageDev{};
  v.desc = d_desc;
  v.scale = d_scale;
  v.lap = d_lap;
  v.n = n;
  v.d = d;
  v.n_pad = n_pad;
  v.perm = reinterpret_cast<uint32_t*>(base + b_desc + 2 * b_vec);
  v.scale_sorted = reinterpret_cast<float*>(base + b_desc + 3 * b_vec);
  v.norm2_sorted = reinterpret_cast<float*>(base + b_desc + 4 * b_vec);
  v.rowop = b_op ? reinterpret_cast<__half*>(base + b_desc + 5 * b_vec) : nullptr;
  v.colop = b_op ? reinterpret_cast<__half*>(base + b_desc + 5 * b_vec + b_op) : nullptr;
  {
    cudaError_t e = ensure_metas(c, (uint32_t)c->h_images.size());
    if (e != cudaSuccess) return cuda_fail(c, e, "fm_upload_image");
    v.meta = c->d_metas.as<ImageMeta>() + img;
  }
  c->dirty.push_back(img);  // sorted / FP16 tensors are built in one batch at the next fm_match
  c->images_dirty = true;
  return FM_OK;
}

int fm_match(fm_ctx* c, const uint32_t* pair_first, const uint32_t* pair_second, size_t n_pairs, float dist,
             float dist2second, uint32_t flags, fm_result** out) {
  if (!c || !out) return FM_ERR_INVALID;
  *out = nullptr;
  if (n_pairs && (!pair_first || !pair_second)) return fail(c, FM_ERR_INVALID, "fm_match: null pair arrays");
  // match.cpp:320-321: a row with no surviving column has d1 == FLT_MAX and is emitted (with a
  // stale column id) only if sqrt(FLT_MAX) = 1.8e19 < dist.  No caller passes such a threshold
  // (FROG.py uses 1e10); refuse it rather than imitate an index leak.
  if (!(dist < 1.8e19f)) return fail(c, FM_ERR_UNSUPPORTED, "fm_match: distance threshold must be below 1.8e19");
  FM_CUDA(c, cudaSetDevice(c->device));
  for (size_t p = 0; p < n_pairs; p++) {
    uint32_t a = pair_first[p], b = pair_second[p];
    if (a >= c->images.size() || b >= c->images.size() || !c->images[a].valid || !c->images[b].valid)
      return fail(c, FM_ERR_INVALID, "fm_match: pair " + std::to_string(p) + " names an image that was not uploaded");
  }
  int rc = sync_images(c);
  if (rc != FM_OK) return rc;
  const bool sym = flags & FM_FLAG_SYM;
  const bool match_all = flags & FM_FLAG_MATCH_ALL;
  const bool force_exact = (flags & FM_FLAG_FORCE_EXACT) || c->dim != (uint32_t)kD;

  // ---- tasks --------------------------------------------------------------------------------
  std::vector<Task> tasks;
  std::vector<uint32_t> pair_of_task;
  tasks.reserve(n_pairs * (sym ? 2 : 1));
  uint64_t total_rows = 0, desc_pairs = 0;
  auto add_task = [&](uint32_t col, uint32_t row, uint32_t tflags, size_t p) {
    if (force_exact || c->h_metas[col].flags || c->h_metas[row].flags) tflags |= kTaskExact;
    tasks.push_back(Task{col, row, 0, tflags});
    pair_of_task.push_back((uint32_t)p);
    total_rows += c->images[row].n;
    desc_pairs += (uint64_t)c->images[col].n * c->images[row].n;
  };
  for (size_t p = 0; p < n_pairs; p++) {
    add_task(pair_first[p], pair_second[p], 0, p);
    if (sym) add_task(pair_second[p], pair_first[p], kTaskSwap, p);
  }
  const uint32_t n_tasks = (uint32_t)tasks.size();

  // ---- batches + prefix arrays (blocks of 128 rows, compaction chunks, score units) ------------
  uint64_t base_units = 0;
  for (auto& t : tasks)
    if (!(t.flags & kTaskExact)) base_units += (c->images[t.row_img].n + kUnitRows - 1) / kUnitRows;
  uint32_t segs = 1;
  if (base_units > 0 && base_units < (uint64_t)2 * c->sm_count)
    segs = (uint32_t)std::min<uint64_t>(8, ((uint64_t)2 * c->sm_count + base_units - 1) / base_units);
  if (match_all && total_rows > 0xFFFFFFF0ull) return fail(c, FM_ERR_UNSUPPORTED, "fm_match: -all supports at most 2^32 outer-loop rows per call");
  const uint32_t kMaxBatchRows = match_all ? 0xFFFFFFF0u : (24u << 20);  // -all: one batch (its scan spans all rows)
  std::vector<Batch> batches;
  std::vector<uint32_t> blk_off, unit_off;
  std::vector<ChunkDesc> chunk_desc;  // compaction chunks of every batch, batch after batch (fm_compact.cuh)
  std::vector<size_t> chunk_cursor;   // first chunk of each batch in chunk_desc
  for (uint32_t t = 0; t < n_tasks;) {
    chunk_cursor.push_back(chunk_desc.size());
    Batch b{t, t, 0, 0, 0, 0, blk_off.size(), false, false};
    while (b.t1 < n_tasks) {
      const uint32_t nr = c->images[tasks[b.t1].row_img].n;
      if (b.t1 > b.t0 && (uint64_t)b.rows + nr > kMaxBatchRows) break;
      tasks[b.t1].row_off = b.rows;
      blk_off.push_back(b.blocks);
      unit_off.push_back(b.units);
      for (uint32_t r0 = 0; r0 < nr; r0 += kCompactChunk)
        chunk_desc.push_back(ChunkDesc{b.rows + r0, std::min<uint32_t>(kCompactChunk, nr - r0) | ((tasks[b.t1].flags & kTaskSwap) ? 0x80000000u : 0u),
                                       r0, pair_of_task[b.t1]});
      b.rows += nr;
      b.blocks += (nr + 127) / 128;
      b.chunks += (nr + kCompactChunk - 1) / kCompactChunk;
      if (tasks[b.t1].flags & kTaskExact) b.any_exact = true;
      else { b.any_fast = true; b.units += ((nr + kUnitRows - 1) / kUnitRows) * segs; }
      b.t1++;
    }
    blk_off.push_back(b.blocks);
    unit_off.push_back(b.units);
    batches.push_back(b);
    t = b.t1;
  }

  fm_result* r = new (std::nothrow) fm_result();
  if (!r) return fail(c, FM_ERR_NOMEM, "fm_match: out of host memory");
  r->ctx = c;
  r->n_pairs = n_pairs;
  r->flags = flags;
  r->ratio = dist2second;
  r->d_out = take_buf(c->out_free, std::max<uint64_t>(total_rows, 1) * sizeof(uint2));
  r->d_counts = take_buf(c->counts_free, std::max<size_t>(n_pairs, 1) * sizeof(uint32_t));
  const bool want_dist = (flags & FM_FLAG_DISTANCES) && !match_all;
  if (want_dist) r->d_dist = take_buf(c->dist_free, std::max<uint64_t>(total_rows, 1) * sizeof(float));
#define FM_CUDA_R(expr)                       \
  do {                                        \
    cudaError_t e__ = (expr);                 \
    if (e__ != cudaSuccess) {                 \
      int code__ = cuda_fail(c, e__, #expr);  \
      fm_result_free(r);                      \
      return code__;                          \
    }                                         \
  } while (0)

  if (c->ev_match.pool.empty() && !c->ev_free.empty()) {
    c->ev_match = std::move(c->ev_free.back());
    c->ev_free.pop_back();
  }
  c->ev_match.reset();
  fm_stats prev = c->stats;
  c->stats = fm_stats{};
  c->stats.ms_prep = prev.ms_prep;
  c->stats.descriptor_pairs = desc_pairs;
  c->stats.rows = total_rows;

  // metadata blob: tasks | chunk descriptors | blk_off | unit_off   (16-byte records first: alignment)
  const size_t np = blk_off.size();
  const size_t off_chunk = sizeof(Task) * n_tasks;
  const size_t off_blk = off_chunk + sizeof(ChunkDesc) * chunk_desc.size();
  const size_t off_unit = off_blk + sizeof(uint32_t) * np;
  const size_t blob_bytes = off_unit + sizeof(uint32_t) * np;
  // The blob is staged in the result's pinned block (behind the counters and per-pair counts the device
  // writes back) so that the copy is asynchronous whatever its size: a pageable source would make
  // cudaMemcpyAsync wait for the stream, i.e. for the previous FM_FLAG_ASYNC call.
  const size_t off_blob = (sizeof(DeviceCounters) + std::max<size_t>(n_pairs, 1) * sizeof(uint32_t) + 15) & ~(size_t)15;
  FM_CUDA_R(take_pinned(c, off_blob + std::max<size_t>(blob_bytes, 16), &r->h_block, &r->h_block_cap));
  unsigned char* blob = static_cast<unsigned char*>(r->h_block) + off_blob;
  if (n_tasks) memcpy(blob, tasks.data(), sizeof(Task) * n_tasks);
  if (!chunk_desc.empty()) memcpy(blob + off_chunk, chunk_desc.data(), sizeof(ChunkDesc) * chunk_desc.size());
  if (np) {
    memcpy(blob + off_blk, blk_off.data(), sizeof(uint32_t) * np);
    memcpy(blob + off_unit, unit_off.data(), sizeof(uint32_t) * np);
  }
  FM_CUDA_R(c->d_meta_blob.ensure(std::max<size_t>(blob_bytes, 16)));
  FM_CUDA_R(cudaMemcpyAsync(c->d_meta_blob.p, blob, std::max<size_t>(blob_bytes, 16), cudaMemcpyHostToDevice, c->stream));
  const unsigned char* blob_d = c->d_meta_blob.as<unsigned char>();
  const Task* d_tasks = reinterpret_cast<const Task*>(blob_d);
  const ChunkDesc* d_chunk = reinterpret_cast<const ChunkDesc*>(blob_d + off_chunk);
  const uint32_t* d_blk = reinterpret_cast<const uint32_t*>(blob_d + off_blk);
  const uint32_t* d_unit = reinterpret_cast<const uint32_t*>(blob_d + off_unit);

  uint32_t max_rows = 1, max_chunks = 1;
  for (auto& b : batches) { max_rows = std::max(max_rows, b.rows); max_chunks = std::max(max_chunks, b.chunks); }
  FM_CUDA_R(c->d_rowres.ensure((size_t)max_rows * sizeof(uint32_t)));
  if (want_dist) {
    FM_CUDA_R(c->d_rowdist.ensure((size_t)max_rows * sizeof(float)));
    FM_CUDA_R(r->d_dist.ensure(std::max<uint64_t>(total_rows, 1) * sizeof(float)));
  }
  FM_CUDA_R(c->d_chunk_count.ensure((size_t)max_chunks * sizeof(uint32_t)));
  FM_CUDA_R(c->d_chunk_out.ensure((size_t)max_chunks * sizeof(unsigned long long)));
  FM_CUDA_R(c->d_chunk_stage.ensure((size_t)max_chunks * kStageCap * sizeof(uint2)));
  if (!c->d_chunk_status.p) {
    // fresh memory may hold anything: clear it once, then the launch epoch tells stale words from current ones
    FM_CUDA_R(c->d_chunk_status.ensure((size_t)kOnePassChunks * sizeof(unsigned long long)));
    FM_CUDA_R(cudaMemsetAsync(c->d_chunk_status.p, 0, c->d_chunk_status.cap, c->stream));
  }
  FM_CUDA_R(c->d_totals.ensure(sizeof(DeviceCounters)));
  FM_CUDA_R(r->d_out.ensure(std::max<uint64_t>(total_rows, 1) * sizeof(uint2)));
  FM_CUDA_R(r->d_counts.ensure(std::max<size_t>(n_pairs, 1) * sizeof(uint32_t)));
  FM_CUDA_R(cudaMemsetAsync(r->d_counts.p, 0, std::max<size_t>(n_pairs, 1) * sizeof(uint32_t), c->stream));
  FM_CUDA_R(cudaMemsetAsync(c->d_totals.p, 0, sizeof(DeviceCounters), c->stream));

  const ImageDev* d_images = c->d_images.as<ImageDev>();
  uint32_t* d_rowres = c->d_rowres.as<uint32_t>();
  float* d_rowdist = want_dist ? c->d_rowdist.as<float>() : nullptr;
  DeviceCounters* d_counters = c->d_totals.as<DeviceCounters>();
  if (match_all) {
    // ---- -all: count pass, per-task scan, emit pass (fm_all.cuh); list sizes are data-dependent ----
    Span total_span(&c->ev_match, c->stream, kPhTotal);
    std::vector<unsigned long long> task_total(std::max<uint32_t>(n_tasks, 1), 0), task_base(std::max<uint32_t>(n_tasks, 1), 0);
    std::vector<uint32_t> pair_counts(std::max<size_t>(n_pairs, 1), 0);
    unsigned long long grand = 0;
    if (!batches.empty() && batches[0].rows > 0) {
      const Batch& b = batches[0];
      const size_t rows = b.rows;
      FM_CUDA_R(c->d_all.ensure(rows * (3 * sizeof(uint32_t) + sizeof(unsigned long long)) + 16));
      FM_CUDA_R(c->d_all_tasks.ensure((size_t)n_tasks * 2 * sizeof(unsigned long long)));
      unsigned long long* row_off = c->d_all.as<unsigned long long>();
      uint32_t* row_count = reinterpret_cast<uint32_t*>(row_off + rows);
      uint32_t* row_final = row_count + rows;
      uint32_t* row_carry = row_final + rows;
      unsigned long long* d_task_total = c->d_all_tasks.as<unsigned long long>();
      unsigned long long* d_task_base = d_task_total + n_tasks;
      const bool d48 = c->dim == (uint32_t)kD;
      const size_t smem = exact_smem_bytes(kD);
      {
        Span sp(&c->ev_match, c->stream, kPhExact);
        if (d48)
          match_all_kernel<kD, false><<<b.blocks, kExactRows, smem, c->stream>>>(d_images, d_tasks, d_blk, n_tasks, dist, row_count,
                                                                                 row_final, nullptr, nullptr, nullptr, nullptr);
        else
          exact_generic_kernel<1><<<b.blocks, kExactRows, 0, c->stream>>>(d_images, d_tasks, d_blk, n_tasks, dist, 1.f, 0u, nullptr, row_count,
                                                                          row_final, nullptr, nullptr, nullptr, nullptr, nullptr);
        all_scan_kernel<<<(n_tasks + 63) / 64, 64, 0, c->stream>>>(d_images, d_tasks, n_tasks, row_count, row_final, row_carry, row_off,
                                                                    d_task_total);
      }
      FM_CUDA_R(cudaMemcpyAsync(task_total.data(), d_task_total, (size_t)n_tasks * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
      FM_CUDA_R(cudaStreamSynchronize(c->stream));
      for (uint32_t t = 0; t < n_tasks; t++) {
        task_base[t] = grand;
        grand += task_total[t];
        const unsigned long long pc = (unsigned long long)pair_counts[pair_of_task[t]] + task_total[t];
        if (pc > 0xFFFFFFFFull) { fm_result_free(r); return fail(c, FM_ERR_UNSUPPORTED, "fm_match: an -all list exceeds 2^32 pairs (pairs.bin block sizes are u32, match.cpp:734)"); }
        pair_counts[pair_of_task[t]] = (uint32_t)pc;
      }
      FM_CUDA_R(r->d_out.ensure(std::max<unsigned long long>(grand, 1) * sizeof(uint2)));
      FM_CUDA_R(cudaMemcpyAsync(d_task_base, task_base.data(), (size_t)n_tasks * sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
      {
        Span sp(&c->ev_match, c->stream, kPhExact);
        if (d48)
          match_all_kernel<kD, true><<<b.blocks, kExactRows, smem, c->stream>>>(d_images, d_tasks, d_blk, n_tasks, dist, nullptr, nullptr,
                                                                                row_carry, row_off, d_task_base, r->d_out.as<uint2>());
        else
          exact_generic_kernel<2><<<b.blocks, kExactRows, 0, c->stream>>>(d_images, d_tasks, d_blk, n_tasks, dist, 1.f, 0u, nullptr, nullptr,
                                                                          nullptr, row_carry, row_off, d_task_base, r->d_out.as<uint2>(), nullptr);
      }
      FM_CUDA_R(cudaStreamSynchronize(c->stream));  // task_base (pageable) must outlive the copy
      c->stats.kernel_launches += 3;
      c->stats.rows_exact += total_rows;
    }
    // hand the totals to the common tail: DeviceCounters.running_total and the per-pair counts
    DeviceCounters hc{};
    hc.running_total = grand;
    FM_CUDA_R(cudaMemcpyAsync(c->d_totals.p, &hc, sizeof hc, cudaMemcpyHostToDevice, c->stream));
    if (n_pairs)
      FM_CUDA_R(cudaMemcpyAsync(r->d_counts.p, pair_counts.data(), n_pairs * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    FM_CUDA_R(cudaStreamSynchronize(c->stream));
  } else {
    Span total_span(&c->ev_match, c->stream, kPhTotal);
    // Two-phase scoring (score_kernel kVar 1 / 2) pays only when nearly every row fails the reference's ratio test --
    // images with little in common at -d2 < 1 -- and costs ~60 % extra otherwise.  So it is used when the previous
    // call of this kind on this context rejected >= 99 % of its rows, or, inside a synchronous multi-batch call, once
    // the first batch (single pass) has shown that much.
    constexpr double kTwoPhaseMin = 0.99;
    const bool two_phase_possible = dist2second < 1.f && segs == 1 && !force_exact;
    bool two_phase = two_phase_possible && c->reject_hint >= kTwoPhaseMin;
    bool probed = false;
    if (g_debug.two_phase >= 0) {  // test hook: force the choice
      two_phase = two_phase_possible && g_debug.two_phase != 0;
      probed = true;
    }
    for (size_t bi = 0; bi < batches.size(); bi++) {
      const Batch& b = batches[bi];
      const uint32_t nt = b.t1 - b.t0;
      const uint32_t* blk = d_blk + b.cursor;
      const uint32_t* unt = d_unit + b.cursor;
      if (b.rows == 0) continue;
      if (b.any_fast) {
        FastBatchArgs fa{};
        fa.images = d_images;
        fa.tasks = d_tasks + b.t0;
        fa.n_tasks = nt;
        fa.rows = b.rows;
        fa.blk_off = blk;
        fa.blocks128 = b.blocks;
        fa.unit_off = unt;
        fa.units = b.units;
        fa.segs = segs;
        fa.thr = dist;
        fa.ratio = dist2second;
        fa.rowres = d_rowres;
        fa.rowdist = d_rowdist;
        fa.two_phase = two_phase;
        fa.counters = d_counters;
        cudaError_t e = fast_match_batch(c, fa);
        if (e != cudaSuccess) {
          int code = cuda_fail(c, e, "fm_match: tensor-core path");
          fm_result_free(r);
          return code;
        }
      }
      if (b.any_exact) {
        Span sp(&c->ev_match, c->stream, kPhExact);
        if (c->dim == (uint32_t)kD)
          exact_match_kernel<kD><<<b.blocks, kExactRows, exact_smem_bytes(kD), c->stream>>>(
              d_images, d_tasks + b.t0, blk, nt, dist, dist2second, kTaskExact, d_rowres, d_rowdist);
        else
          exact_generic_kernel<0><<<b.blocks, kExactRows, 0, c->stream>>>(d_images, d_tasks + b.t0, blk, nt, dist, dist2second, kTaskExact,
                                                                          d_rowres, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, d_rowdist);
        c->stats.kernel_launches++;
        for (uint32_t t = b.t0; t < b.t1; t++)
          if (tasks[t].flags & kTaskExact) c->stats.rows_exact += c->images[tasks[t].row_img].n;
      }
      {
        Span sp(&c->ev_match, c->stream, kPhCompact);
        CompactArgs ca{};
        ca.chunks = d_chunk + chunk_cursor[bi];
        ca.n_chunks = b.chunks;
        ca.rowres = d_rowres;
        ca.rowdist = d_rowdist;
        ca.chunk_count = c->d_chunk_count.as<uint32_t>();
        ca.chunk_out = c->d_chunk_out.as<unsigned long long>();
        ca.stage = c->d_chunk_stage.as<uint2>();
        ca.pair_count = r->d_counts.as<uint32_t>();
        ca.running_total = &d_counters->running_total;
        ca.out_pairs = r->d_out.as<uint2>();
        ca.out_dist = want_dist ? r->d_dist.as<float>() : nullptr;
        if (b.chunks <= kOnePassChunks) {
          ca.status = c->d_chunk_status.as<unsigned long long>();
          c->compact_epoch = (c->compact_epoch % 0x3FFFFFFEu) + 1u;  // 1 .. 2^30 - 2: never the cleared value 0
          ca.epoch = c->compact_epoch;
          if (want_dist) compact_onepass_kernel<true><<<b.chunks, kCompactThreads, 0, c->stream>>>(ca);
          else compact_onepass_kernel<false><<<b.chunks, kCompactThreads, 0, c->stream>>>(ca);
          c->stats.kernel_launches += 1;
        } else if (g_debug.compact_v >= 2) {
          // persistent, software-pipelined count pass: three CTAs per SM, each with the next chunk's rows in flight
          const uint32_t grid_count = std::min<uint32_t>(b.chunks, (uint32_t)c->sm_count * 3u);
          const uint32_t grid = std::min<uint32_t>(b.chunks, (uint32_t)c->sm_count * 8u);
          if (want_dist) {
            if (g_debug.compact_v >= 3) compact_count3_kernel<true><<<grid_count, kCompactThreads, 0, c->stream>>>(ca);
            else compact_count2_kernel<true><<<grid_count, kCompactThreads, 0, c->stream>>>(ca);
            compact_scan2_kernel<<<1, 1024, 0, c->stream>>>(ca);
            compact_scatter2_kernel<true><<<grid, kCompactThreads, 0, c->stream>>>(ca);
          } else {
            if (g_debug.compact_v >= 3) compact_count3_kernel<false><<<grid_count, kCompactThreads, 0, c->stream>>>(ca);
            else compact_count2_kernel<false><<<grid_count, kCompactThreads, 0, c->stream>>>(ca);
            compact_scan2_kernel<<<1, 1024, 0, c->stream>>>(ca);
            compact_scatter2_kernel<false><<<grid, kCompactThreads, 0, c->stream>>>(ca);
          }
          c->stats.kernel_launches += 3;
        } else if (want_dist) {
          const uint32_t grid = std::min<uint32_t>(b.chunks, (uint32_t)c->sm_count * 8u);  // grid-stride over the chunks
          compact_count_kernel<true><<<b.chunks, kCompactThreads, 0, c->stream>>>(ca);  // (one CTA per chunk measured faster here)
          compact_scan_kernel<<<1, 1024, 0, c->stream>>>(ca);
          compact_scatter_kernel<true><<<grid, kCompactThreads, 0, c->stream>>>(ca);
          c->stats.kernel_launches += 3;
        } else {
          const uint32_t grid = std::min<uint32_t>(b.chunks, (uint32_t)c->sm_count * 8u);
          compact_count_kernel<false><<<b.chunks, kCompactThreads, 0, c->stream>>>(ca);
          compact_scan_kernel<<<1, 1024, 0, c->stream>>>(ca);
          compact_scatter_kernel<false><<<grid, kCompactThreads, 0, c->stream>>>(ca);
          c->stats.kernel_launches += 3;
        }
      }
      if (two_phase_possible && !two_phase && !probed && b.any_fast && !b.any_exact && bi + 1 < batches.size() &&
          !(flags & FM_FLAG_ASYNC)) {
        // a synchronous call with more batches to come: look at what the first one rejected
        probed = true;
        DeviceCounters* hc = reinterpret_cast<DeviceCounters*>(r->h_block);
        FM_CUDA_R(cudaMemcpyAsync(hc, c->d_totals.p, sizeof(DeviceCounters), cudaMemcpyDeviceToHost, c->stream));
        FM_CUDA_R(cudaStreamSynchronize(c->stream));
        two_phase = (double)hc->rescore.rejected_early >= kTwoPhaseMin * (double)b.rows;
      }
    }
  }
  FM_CUDA_R(cudaGetLastError());

  // ---- totals and per-pair counts to a pinned block; (optionally) the lists to the host ----------
  FM_CUDA_R(cudaMemcpyAsync(r->h_block, c->d_totals.p, sizeof(DeviceCounters), cudaMemcpyDeviceToHost, c->stream));
  if (n_pairs)
    FM_CUDA_R(cudaMemcpyAsync(static_cast<char*>(r->h_block) + sizeof(DeviceCounters), r->d_counts.p, n_pairs * sizeof(uint32_t),
                              cudaMemcpyDeviceToHost, c->stream));
  r->done = c->ev_match.get();
  FM_CUDA_R(cudaEventRecord(r->done, c->stream));
  r->stats = c->stats;
  r->ev = std::move(c->ev_match);
  c->ev_match = EventPool{};
  c->last = r;
  if (!(flags & FM_FLAG_ASYNC)) {
    rc = fm_result_wait(r);
    if (rc != FM_OK) { fm_result_free(r); return rc; }
  }
  *out = r;
  return FM_OK;
#undef FM_CUDA_R
}

int fm_result_wait(fm_result* r) {
  if (!r) return FM_ERR_INVALID;
  int rc = finalize(r);
  if (rc != FM_OK) return rc;
  if (!(r->flags & FM_FLAG_DEVICE_ONLY)) return fm_result_fetch(r);
  return FM_OK;
}

int fm_result_fetch(fm_result* r) {
  if (!r) return FM_ERR_INVALID;
  if (r->fetched) return FM_OK;
  fm_ctx* c = r->ctx;
  int rc = finalize(r);
  if (rc != FM_OK) return rc;
  FM_CUDA(c, cudaSetDevice(c->device));
  const size_t bytes = std::max<uint64_t>(r->total, 1) * sizeof(uint2);
  if (c->cache_pinned && c->cache_pinned_cap >= bytes) {
    r->h_pairs = static_cast<uint32_t*>(c->cache_pinned);
    r->h_cap = c->cache_pinned_cap;
    c->cache_pinned = nullptr;
    c->cache_pinned_cap = 0;
  } else {
    const size_t want = bytes + bytes / 4 + 4096;
    FM_CUDA(c, cudaMallocHost(reinterpret_cast<void**>(&r->h_pairs), want));
    r->h_cap = want;
  }
  if (r->d_dist.p && !r->h_dist) FM_CUDA(c, cudaMallocHost(reinterpret_cast<void**>(&r->h_dist), std::max<uint64_t>(r->total, 1) * sizeof(float)));
  if (r->total) {
    FM_CUDA(c, cudaMemcpyAsync(r->h_pairs, r->d_out.p, r->total * sizeof(uint2), cudaMemcpyDeviceToHost, c->stream));
    if (r->h_dist) FM_CUDA(c, cudaMemcpyAsync(r->h_dist, r->d_dist.p, r->total * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    FM_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  r->fetched = true;
  return FM_OK;
}

size_t fm_result_num_pairs(const fm_result* r) { return r ? r->n_pairs : 0; }
uint64_t fm_result_total(const fm_result* r) { return (r && r->waited) ? r->total : 0; }
uint32_t fm_result_count(const fm_result* r, size_t p) { return (r && r->waited && p < r->n_pairs) ? r->counts[p] : 0; }
const uint32_t* fm_result_pairs(const fm_result* r, size_t p) {
  if (!r || p >= r->n_pairs || !r->fetched) return nullptr;
  return r->h_pairs + 2 * r->offsets[p];
}
const float* fm_result_distances(const fm_result* r, size_t p) {
  if (!r || p >= r->n_pairs || !r->fetched || !r->h_dist) return nullptr;
  return r->h_dist + r->offsets[p];
}
const uint32_t* fm_result_device_counts(const fm_result* r) { return r ? r->d_counts.as<uint32_t>() : nullptr; }
const uint32_t* fm_result_device_pairs(const fm_result* r) { return r ? r->d_out.as<uint32_t>() : nullptr; }

void fm_result_free(fm_result* r) {
  if (!r) return;
  fm_ctx* c = r->ctx;
  cudaSetDevice(c->device);
  if (r->done && !r->waited) cudaEventSynchronize(r->done);  // the pinned block is still a copy target
  if (c->last == r) c->last = nullptr;
  // hand the buffers and events back to the context so the next call does not reallocate
  give_buf(c->out_free, r->d_out);
  give_buf(c->counts_free, r->d_counts);
  give_buf(c->dist_free, r->d_dist);
  if (r->h_dist) cudaFreeHost(r->h_dist);
  give_pinned(c, r->h_block, r->h_block_cap);
  if (!r->ev.pool.empty()) {
    if (c->ev_free.size() < 8) c->ev_free.push_back(std::move(r->ev));
    else r->ev.destroy();
  }
  if (r->h_pairs) {
    if (r->h_cap > c->cache_pinned_cap) {
      if (c->cache_pinned) cudaFreeHost(c->cache_pinned);
      c->cache_pinned = r->h_pairs;
      c->cache_pinned_cap = r->h_cap;
    } else {
      cudaFreeHost(r->h_pairs);
    }
  }
  delete r;
}

int fm_get_stats(fm_ctx* c, fm_stats* out) {
  if (!c || !out) return FM_ERR_INVALID;
  FM_CUDA(c, cudaSetDevice(c->device));
  FM_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->last && !c->last->waited) {
    int rc = finalize(c->last);
    if (rc != FM_OK) return rc;
  }
  c->stats.ms_prep = c->ms_prep_acc;
  *out = c->stats;
  return FM_OK;
}

int fm_result_stats(fm_result* r, fm_stats* out) {
  if (!r || !out) return FM_ERR_INVALID;
  int rc = finalize(r);
  if (rc != FM_OK) return rc;
  *out = r->stats;
  return FM_OK;
}

// ---- consumer hand-off: link build (fm_links.cuh) -------------------------------------------------

int fm_links_build(fm_result* r, const uint32_t* pair_first, const uint32_t* pair_second, const uint32_t* block_order,
                   fm_links** out) {
  if (!r || !out) return FM_ERR_INVALID;
  *out = nullptr;
  fm_ctx* c = r->ctx;
  if (!r->waited) return fail(c, FM_ERR_INVALID, "fm_links_build: the result is not complete (fm_result_wait first)");
  if (r->flags & FM_FLAG_MATCH_ALL) return fail(c, FM_ERR_UNSUPPORTED, "fm_links_build: -all results are not supported");
  if (r->n_pairs && (!pair_first || !pair_second)) return fail(c, FM_ERR_INVALID, "fm_links_build: null pair arrays");
  FM_CUDA(c, cudaSetDevice(c->device));
  fm_links* l = new (std::nothrow) fm_links();
  if (!l) return fail(c, FM_ERR_NOMEM, "fm_links_build: out of host memory");
  l->ctx = c;
  const size_t n_img = c->images.size();
  l->point_base.assign(n_img + 1, 0);
  for (size_t i = 0; i < n_img; i++) l->point_base[i + 1] = l->point_base[i] + (c->images[i].valid ? c->images[i].n : 0);
  if (l->point_base[n_img] > 0xFFFFFFF0ull) { delete l; return fail(c, FM_ERR_UNSUPPORTED, "fm_links_build: more than 2^32 points"); }
  l->n_points = (uint32_t)l->point_base[n_img];
  // blocks in file order
  std::vector<LinkBlock> blocks(r->n_pairs);
  std::vector<unsigned long long> entry_off(r->n_pairs + 1, 0);
  for (size_t k = 0; k < r->n_pairs; k++) {
    const size_t p = block_order ? block_order[k] : k;
    if (p >= r->n_pairs || pair_first[p] >= n_img || pair_second[p] >= n_img) { delete l; return fail(c, FM_ERR_INVALID, "fm_links_build: bad block order or pair"); }
    LinkBlock& b = blocks[k];
    b.src = r->offsets[p];
    b.q0 = entry_off[k];
    b.count = r->counts[p];
    b.image1 = pair_first[p];
    b.image2 = pair_second[p];
    b.base1 = (uint32_t)l->point_base[b.image1];
    b.base2 = (uint32_t)l->point_base[b.image2];
    b.pad_ = 0;
    entry_off[k + 1] = entry_off[k] + b.count;
  }
  const unsigned long long n_entries = entry_off[r->n_pairs];
  l->total = 2ull * n_entries;
  auto bail = [&](cudaError_t e, const char* what) { int code = cuda_fail(c, e, what); fm_links_free(l); return code; };
  DevBuf d_blocks, d_eoff, d_deg, d_cursor, d_tmp;
  struct Scratch { DevBuf* b[5]; ~Scratch() { for (auto* x : b) x->release(); } } scratch{{&d_blocks, &d_eoff, &d_deg, &d_cursor, &d_tmp}};
  cudaError_t e;
  const size_t np1 = (size_t)l->n_points + 1;
  if ((e = d_blocks.ensure(std::max<size_t>(1, blocks.size()) * sizeof(LinkBlock))) != cudaSuccess) return bail(e, "fm_links_build");
  if ((e = d_eoff.ensure(entry_off.size() * sizeof(unsigned long long))) != cudaSuccess) return bail(e, "fm_links_build");
  if ((e = d_deg.ensure(np1 * sizeof(uint32_t))) != cudaSuccess) return bail(e, "fm_links_build");
  if ((e = d_cursor.ensure(np1 * sizeof(uint32_t))) != cudaSuccess) return bail(e, "fm_links_build");
  if ((e = d_tmp.ensure(std::max<uint64_t>(l->total, 1) * sizeof(HalfLink))) != cudaSuccess) return bail(e, "fm_links_build");
  if ((e = l->d_off.ensure(np1 * sizeof(unsigned long long))) != cudaSuccess) return bail(e, "fm_links_build");
  if ((e = l->d_links.ensure(std::max<uint64_t>(l->total, 1) * sizeof(uint2))) != cudaSuccess) return bail(e, "fm_links_build");
  cudaEvent_t ev0, ev1;
  cudaEventCreate(&ev0);
  cudaEventCreate(&ev1);
  if (!blocks.empty() && (e = cudaMemcpyAsync(d_blocks.p, blocks.data(), blocks.size() * sizeof(LinkBlock), cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) return bail(e, "fm_links_build");
  if ((e = cudaMemcpyAsync(d_eoff.p, entry_off.data(), entry_off.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) return bail(e, "fm_links_build");
  cudaMemsetAsync(d_deg.p, 0, np1 * sizeof(uint32_t), c->stream);
  cudaMemsetAsync(d_cursor.p, 0, np1 * sizeof(uint32_t), c->stream);
  cudaEventRecord(ev0, c->stream);
  const unsigned grid = (unsigned)((n_entries + 255) / 256);
  const uint2* pairs = r->d_out.as<uint2>();
  if (n_entries)
    links_scatter_kernel<false><<<grid, 256, 0, c->stream>>>(d_blocks.as<LinkBlock>(), d_eoff.as<unsigned long long>(), (uint32_t)r->n_pairs, n_entries,
                                                            pairs, d_deg.as<uint32_t>(), nullptr, nullptr, nullptr);
  links_scan_kernel<<<1, 1024, 0, c->stream>>>(d_deg.as<uint32_t>(), l->n_points, l->d_off.as<unsigned long long>());
  if (n_entries)
    links_scatter_kernel<true><<<grid, 256, 0, c->stream>>>(d_blocks.as<LinkBlock>(), d_eoff.as<unsigned long long>(), (uint32_t)r->n_pairs, n_entries,
                                                           pairs, nullptr, l->d_off.as<unsigned long long>(), d_cursor.as<uint32_t>(), d_tmp.as<HalfLink>());
  if (l->n_points)
    links_sort_kernel<<<(l->n_points + 127) / 128, 128, 0, c->stream>>>(l->d_off.as<unsigned long long>(), l->n_points, d_tmp.as<HalfLink>(),
                                                                        l->d_links.as<uint2>());
  cudaEventRecord(ev1, c->stream);
  if ((e = cudaGetLastError()) != cudaSuccess) return bail(e, "fm_links_build: launch");
  if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) return bail(e, "fm_links_build");  // the scratch buffers and host tables die here
  cudaEventElapsedTime(&l->ms_build, ev0, ev1);
  cudaEventDestroy(ev0);
  cudaEventDestroy(ev1);
  *out = l;
  return FM_OK;
}

uint64_t fm_links_total(const fm_links* l) { return l ? l->total : 0; }

int fm_links_fetch(fm_links* l) {
  if (!l) return FM_ERR_INVALID;
  if (l->h_off) return FM_OK;
  fm_ctx* c = l->ctx;
  FM_CUDA(c, cudaSetDevice(c->device));
  const size_t np1 = (size_t)l->n_points + 1;
  FM_CUDA(c, cudaMallocHost(reinterpret_cast<void**>(&l->h_off), np1 * sizeof(unsigned long long)));
  FM_CUDA(c, cudaMallocHost(reinterpret_cast<void**>(&l->h_links), std::max<uint64_t>(l->total, 1) * sizeof(uint2)));
  FM_CUDA(c, cudaMemcpyAsync(l->h_off, l->d_off.p, np1 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  if (l->total) FM_CUDA(c, cudaMemcpyAsync(l->h_links, l->d_links.p, l->total * sizeof(uint2), cudaMemcpyDeviceToHost, c->stream));
  FM_CUDA(c, cudaStreamSynchronize(c->stream));
  return FM_OK;
}

const uint64_t* fm_links_offsets(const fm_links* l, uint32_t img, uint32_t* n_points) {
  if (!l || !l->h_off || (size_t)img + 1 >= l->point_base.size()) return nullptr;
  if (n_points) *n_points = (uint32_t)(l->point_base[img + 1] - l->point_base[img]);
  return reinterpret_cast<const uint64_t*>(l->h_off) + l->point_base[img];
}
const uint32_t* fm_links_data(const fm_links* l) { return (l && l->h_off) ? l->h_links : nullptr; }
const uint64_t* fm_links_device_offsets(const fm_links* l) { return l ? reinterpret_cast<const uint64_t*>(l->d_off.p) : nullptr; }
const uint32_t* fm_links_device_data(const fm_links* l) { return l ? l->d_links.as<uint32_t>() : nullptr; }
float fm_links_build_ms(const fm_links* l) { return l ? l->ms_build : 0.f; }

void fm_links_free(fm_links* l) {
  if (!l) return;
  cudaSetDevice(l->ctx->device);
  l->d_off.release();
  l->d_links.release();
  if (l->h_off) cudaFreeHost(l->h_off);
  if (l->h_links) cudaFreeHost(l->h_links);
  delete l;
}

// ---- debug / test hooks (include/frogmatch_debug.h) ---------------------------------------------

int fm_debug_set_option(const char* name, int value) {
  if (!name) return FM_ERR_INVALID;
  if (!strcmp(name, "probe")) g_debug.probe = value;
  else if (!strcmp(name, "variant")) g_debug.variant = value;
  else if (!strcmp(name, "pre_tiles")) g_debug.pre_tiles = value;
  else if (!strcmp(name, "two_phase")) g_debug.two_phase = value;
  else if (!strcmp(name, "surv_cap")) g_debug.surv_cap = value;
  else if (!strcmp(name, "compact_v")) g_debug.compact_v = value;
  else return FM_ERR_INVALID;
  return FM_OK;
}

int fm_debug_image(fm_ctx* c, uint32_t img, uint32_t* flags, uint32_t* n_classes, float* class_lap,
                   uint32_t* class_begin, float* max_norm2, uint32_t* perm, float* scale_sorted,
                   uint16_t* rowop, uint16_t* colop) {
  if (!c || img >= c->images.size() || !c->images[img].valid) return FM_ERR_INVALID;
  FM_CUDA(c, cudaSetDevice(c->device));
  int rc = sync_images(c);
  if (rc != FM_OK) return rc;
  const ImageMeta& m = c->h_metas[img];
  const Image& im = c->images[img];
  const ImageDev& v = c->h_images[img];
  if (flags) *flags = m.flags;
  if (n_classes) *n_classes = m.n_classes;
  if (max_norm2) *max_norm2 = m.max_norm2;
  if (class_lap) memcpy(class_lap, m.class_lap, sizeof m.class_lap);
  if (class_begin) memcpy(class_begin, m.class_begin, sizeof m.class_begin);
  if (perm && im.n) FM_CUDA(c, cudaMemcpy(perm, v.perm, im.n * 4, cudaMemcpyDeviceToHost));
  if (scale_sorted && im.n) FM_CUDA(c, cudaMemcpy(scale_sorted, v.scale_sorted, im.n * 4, cudaMemcpyDeviceToHost));
  for (int role = 0; role < 2; role++) {
    uint16_t* dst = role ? colop : rowop;
    const __half* src = role ? v.colop : v.rowop;
    if (!dst) continue;
    if (!src) return fail(c, FM_ERR_INVALID, "fm_debug_image: image has no FP16 operands (d != 48)");
    std::vector<uint8_t> raw((size_t)v.n_pad * 128);
    FM_CUDA(c, cudaMemcpy(raw.data(), src, raw.size(), cudaMemcpyDeviceToHost));
    for (uint32_t s = 0; s < v.n_pad; s++)  // undo the SWIZZLE_128B tile layout
      for (uint32_t q = 0; q < 8; q++)
        memcpy(dst + (size_t)s * 64 + q * 8, raw.data() + (size_t)(s >> 7) * 16384 + sw128_offset(s & 127u, q), 16);
  }
  return FM_OK;
}

int fm_debug_score_unit(fm_ctx* c, uint32_t first_img, uint32_t second_img, uint32_t row_block, float* t_out,
                        uint32_t ld, uint32_t* bands_out, float* cand_t, uint32_t* cand_col) {
  if (!c || first_img >= c->images.size() || second_img >= c->images.size() || !c->images[first_img].valid ||
      !c->images[second_img].valid || c->dim != (uint32_t)kD)
    return FM_ERR_INVALID;
  FM_CUDA(c, cudaSetDevice(c->device));
  int rc = sync_images(c);
  if (rc != FM_OK) return rc;
  const uint32_t nB = c->images[second_img].n;
  const uint32_t n_pad_a = c->h_images[first_img].n_pad;
  if (row_block * kUnitRows >= nB || ld < n_pad_a) return fail(c, FM_ERR_INVALID, "fm_debug_score_unit: bad row block or ld");
  Task task{first_img, second_img, 0, 0};
  const uint32_t blocks = (nB + 127) / 128, units = (nB + kUnitRows - 1) / kUnitRows;
  uint32_t meta[4] = {0, blocks, 0, units};
  DevBuf d_task, d_meta, d_dump;
  FM_CUDA(c, d_task.ensure(sizeof task));
  FM_CUDA(c, d_meta.ensure(sizeof meta));
  FM_CUDA(c, d_dump.ensure((size_t)kUnitRows * ld * sizeof(float)));
  FM_CUDA(c, c->d_bands.ensure((size_t)nB * sizeof(uint2)));
  FM_CUDA(c, c->d_cands.ensure((size_t)nB * kTopK * sizeof(Cand)));
  FM_CUDA(c, cudaMemcpyAsync(d_task.p, &task, sizeof task, cudaMemcpyHostToDevice, c->stream));
  FM_CUDA(c, cudaMemcpyAsync(d_meta.p, meta, sizeof meta, cudaMemcpyHostToDevice, c->stream));
  FM_CUDA(c, cudaMemsetAsync(d_dump.p, 0xFF, (size_t)kUnitRows * ld * sizeof(float), c->stream));  // NaN = "not written"
  FM_CUDA(c, cudaFuncSetAttribute(score_kernel<true, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kScoreSmemBytes));
  bands_kernel<<<blocks, 128, 0, c->stream>>>(c->d_images.as<ImageDev>(), d_task.as<Task>(), d_meta.as<uint32_t>(), 1,
                                              c->d_bands.as<uint2>());
  score_kernel<true, 0, 0><<<1, kScoreThreads, kScoreSmemBytes, c->stream>>>(
      c->d_images.as<ImageDev>(), d_task.as<Task>(), d_meta.as<uint32_t>() + 2, 1, 1, c->d_bands.as<uint2>(),
      c->d_cands.as<Cand>(), nullptr, d_dump.as<float>(), ld, row_block, 0, 0.f, 0.f, nullptr, nullptr, nullptr);
  FM_CUDA(c, cudaGetLastError());
  FM_CUDA(c, cudaStreamSynchronize(c->stream));
  const uint32_t r0 = row_block * kUnitRows, nr = std::min<uint32_t>(kUnitRows, nB - r0);
  if (t_out) FM_CUDA(c, cudaMemcpy(t_out, d_dump.p, (size_t)kUnitRows * ld * sizeof(float), cudaMemcpyDeviceToHost));
  if (bands_out) FM_CUDA(c, cudaMemcpy(bands_out, c->d_bands.as<uint2>() + r0, (size_t)nr * sizeof(uint2), cudaMemcpyDeviceToHost));
  if (cand_t && cand_col) {
    std::vector<Cand> h((size_t)nr * kTopK);
    FM_CUDA(c, cudaMemcpy(h.data(), c->d_cands.as<Cand>() + (size_t)r0 * kTopK, h.size() * sizeof(Cand), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < h.size(); i++) { cand_t[i] = h[i].t; cand_col[i] = h[i].col; }
  }
  d_task.release(); d_meta.release(); d_dump.release();
  return FM_OK;
}

}  // extern "C"
