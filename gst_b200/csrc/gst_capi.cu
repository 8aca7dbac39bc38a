// gst_b200 -- host side of the C ABI declared in include/gst_cuda.h.
//
// Replaces, for the decode path only, the reference's gpu/gpu.cpp + gpu/kernel_cache.cpp
// (OpenCL context / queues / runtime kernel compilation) and the host orchestration of
// codec/decoder.cpp:98-428.  Kernels are compiled ahead of time for sm_100a; ordering is by
// CUDA stream instead of the reference's explicit cl_event DAG on out-of-order queues.
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/gst_cuda.h"
#include "gst_kernels.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define GST_CUDA_TRY(expr)                                                                 \
  do {                                                                                     \
    cudaError_t e_ = (expr);                                                               \
    if (e_ != cudaSuccess)                                                                 \
      return fail(GST_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_),    \
                  __FILE__, __LINE__);                                                     \
  } while (0)

constexpr int kNumWorkStreams = 4;  // gpu/gpu.h:49 kMaxNumWorkQueues
// host-batch workers (gst_decompress_host_batch / gst_load_host_batch): each packs its pages into pinned staging on
// its own thread and stream.  One host thread copies ~13 GB/s and the host memory system saturates around 8 threads
// (measured on the GPU boxes), so 8 workers are what it takes to keep a PCIe 5 x16 link (~55 GB/s) fed.
constexpr int kHostSlots = 8;
constexpr size_t kQuantum = 512;    // the reference's sub-buffer alignment quantum
constexpr size_t kSyncPoolWords = 1 << 18;  // hand-over counters a stream hands out before it zeroes its pool again (1 MiB)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace

struct gst_ctx {
  int device = 0;
  cudaStream_t streams[1 + kNumWorkStreams] = {};
  cudaStream_t host_streams[kHostSlots] = {};
  std::atomic<uint32_t> next_stream{0};
  // PreloadedMemory (codec/decoder.cpp:49-95): bump arena, never reset until freed
  std::mutex arena_mutex;
  uint8_t *arena = nullptr;
  size_t arena_size = 0;
  size_t arena_off = 0;
  // Scratch of calls made without a preallocated arena: one grow-only buffer per stream the context has seen.
  // Calls on one stream are ordered, so the buffer is reused by the next call without any allocation; it is
  // released with the context (or gst_free_scratch).
  struct StreamScratch {
    std::mutex m;  // held from the lookup to the last launch of a call: a concurrent call on the same stream must not
                   // regrow the buffer in between
    uint8_t *ptr = nullptr;
    size_t cap = 0;
    // hand-over counters of the calls on this stream (BatchParams::img_done), also used by calls that take their
    // scratch from the arena: a zeroed pool every call cuts its n words from; when the pool is used up it is
    // zeroed again in stream order (once per ~hundreds of calls), so no kernel ever has to reset a counter
    uint32_t *sync = nullptr;
    size_t sync_cap = 0, sync_used = 0;
    void release() {
      if (ptr) cudaFree(ptr);
      if (sync) cudaFree(sync);
      ptr = nullptr;
      sync = nullptr;
      cap = sync_cap = sync_used = 0;
    }
  };
  std::mutex scratch_mutex;
  std::unordered_map<cudaStream_t, std::unique_ptr<StreamScratch>> stream_scratch;
  // staging of gst_decompress_host_batch / gst_load_host_batch: a pool of slots (pinned staging x 2, device input,
  // device output, stream), grow-only.  A worker thread of a call takes a free slot for the pages it handles, so
  // concurrent callers (the reference's pool threads, demo/photos_sf.cpp:747-830) overlap instead of queueing
  // behind one another.
  std::mutex slot_mutex;
  std::condition_variable slot_cv;
  bool slot_busy[kHostSlots] = {};
  struct HostSlotT {
    uint8_t *pinned[2] = {nullptr, nullptr};
    cudaEvent_t h2d_done[2] = {nullptr, nullptr};
    uint8_t *d_in = nullptr, *d_out = nullptr;
    size_t cap_in = 0, cap_out = 0;
  } host_slots[kHostSlots];
  std::atomic<bool> direct_upload{false};  // gst_ctx_set_direct_upload
  size_t sync_pool_words = kSyncPoolWords;  // (GST_SYNC_POOL_WORDS in the environment: tests make the pool wrap)
  // [0]: status flags the kernels OR into (gst_status_flags), [1]: a zero word
  uint32_t *d_status = nullptr;
  // workspace of the standalone rANS decode / encode entry points (grow-only, one call at a time)
  std::mutex ans_mutex;
  uint8_t *ans_ws = nullptr;
  size_t ans_ws_cap = 0;
  // optional per-kernel timing (gst_profile_*): events around every kernel of every call
  std::mutex prof_mutex;
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_events;  // (kLaunchesPerBatch + 1) per recorded call
};

struct gst_ans_decoder {
  gst_ctx *ctx = nullptr;
  uint32_t lanes = 0;
  uint32_t *table = nullptr;  // device, 2048 packed entries
  uint8_t *freqs = nullptr;   // device, 512 B
};

// Frame streamer: `depth` frames in flight, each slot owns pinned staging, a device input, a
// device output, a stream and a completion event (demo/demo.cpp:145-243 keeps ONE frame in flight
// and blocks on its event, :221).
struct gst_streamer {
  gst_ctx *ctx = nullptr;
  uint32_t width = 0, height = 0, depth = 0;
  int mode = 0;
  size_t frame_bytes = 0;  // decoded bytes per frame
  uint64_t next_ticket = 0;
  struct Slot {
    uint8_t *pinned = nullptr, *d_in = nullptr, *d_out = nullptr;
    size_t cap_in = 0, cap_out = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    uint64_t ticket = UINT64_MAX;  // frame currently (or last) in the slot
    void *frame_dev = nullptr;     // where that frame was decoded to
  };
  std::vector<Slot> slots;
};

namespace {

using HostSlot = gst_ctx::HostSlotT;

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    int cur = -1;
    cudaGetDevice(&cur);
    if (prev >= 0 && cur != prev) cudaSetDevice(prev);
  }
};

int check_dims(const gst_header &h) {
  if (h.width == 0 || h.height == 0 || (h.width % 128) || (h.height % 128))
    return fail(GST_ERR_INVALID, "image dimensions %ux%u must be non-zero multiples of 128", h.width, h.height);
  const uint64_t n = static_cast<uint64_t>(h.width / 4) * (h.height / 4);
  if (n % gst::kGroupSyms)
    return fail(GST_ERR_INVALID, "width*height/16 = %llu must be a multiple of 8192", (unsigned long long)n);
  if (h.palette_bytes % gst::kGroupSyms)
    return fail(GST_ERR_INVALID, "palette_bytes = %u must be a multiple of 8192", h.palette_bytes);
  if ((h.y_cmp_sz | h.chroma_cmp_sz | h.palette_sz | h.indices_sz) & 3u)
    return fail(GST_ERR_INVALID, "compressed stream sizes must be multiples of 4");
  if (n > 0x7FFFFFFFull / 8) return fail(GST_ERR_INVALID, "image too large");
  return GST_OK;
}

int check_header(const gst_header &h) {
  int rc = check_dims(h);
  if (rc) return rc;
  const uint64_t n = static_cast<uint64_t>(h.width / 4) * (h.height / 4);
  // every stream starts with one u32 end offset per group and every group ends with its 32 states
  const uint64_t groups[4] = {2 * n / gst::kGroupSyms, 4 * n / gst::kGroupSyms, h.palette_bytes / gst::kGroupSyms,
                              n / gst::kGroupSyms};
  const uint32_t sizes[4] = {h.y_cmp_sz, h.chroma_cmp_sz, h.palette_sz, h.indices_sz};
  for (int s = 0; s < 4; ++s)
    if (sizes[s] < groups[s] * (4 + 4 * gst::kLanes))
      return fail(GST_ERR_INVALID, "stream %d of %u bytes cannot hold %llu groups", s, sizes[s], (unsigned long long)groups[s]);
  return GST_OK;
}

struct BatchLayout {
  uint32_t n_blocks = 0, groups_per_plane = 0, max_palette = 0;
  size_t off_region = 0, payload_bytes = 0, palette_total = 0, total_cmp = 0;
  // scratch carve-up
  size_t tables_off = 0, sym_off = 0, palette_off = 0, idx_off = 0, total_off = 0, run_off = 0, scratch_bytes = 0;
  bool idx16 = true;
};

int layout_batch(const gst_header *hdrs, uint32_t n, BatchLayout *L) {
  if (!hdrs || n == 0) return fail(GST_ERR_INVALID, "empty batch");
  uint64_t in_total = 0, out_total = 0, pal_total = 0;
  uint32_t max_pal = 0;
  for (uint32_t i = 0; i < n; ++i) {
    int rc = check_header(hdrs[i]);
    if (rc) return rc;
    // codec/decoder.cpp:117-121: all images of one call share their dimensions
    if (hdrs[i].width != hdrs[0].width || hdrs[i].height != hdrs[0].height)
      return fail(GST_ERR_INVALID, "image %u is %ux%u but image 0 is %ux%u: one call needs equal dimensions", i,
                  hdrs[i].width, hdrs[i].height, hdrs[0].width, hdrs[0].height);
    in_total += static_cast<uint64_t>(hdrs[i].y_cmp_sz) + hdrs[i].chroma_cmp_sz + hdrs[i].palette_sz + hdrs[i].indices_sz;
    pal_total += hdrs[i].palette_bytes;
    max_pal = std::max(max_pal, hdrs[i].palette_bytes);
  }
  const uint64_t N = static_cast<uint64_t>(hdrs[0].width / 4) * (hdrs[0].height / 4);
  out_total = 7 * N * n + pal_total;
  // the device-side offset table is cl_uint (codec/decoder.cpp:133-149)
  if (in_total > 0xFFFFFFFFull || out_total > 0xFFFFFFFFull)
    return fail(GST_ERR_INVALID, "batch of %u images overflows the 32-bit stream offsets; split it into pages", n);
  if (n > 65535u) return fail(GST_ERR_INVALID, "batch of %u images: one call takes at most 65535; split it into pages", n);
  L->n_blocks = static_cast<uint32_t>(N);
  L->groups_per_plane = static_cast<uint32_t>(N / gst::kGroupSyms);
  L->max_palette = max_pal;
  L->off_region = align_up(static_cast<size_t>(n) * 8 * sizeof(uint32_t), kQuantum);
  L->payload_bytes = static_cast<size_t>(in_total);
  L->palette_total = static_cast<size_t>(pal_total);
  L->total_cmp = L->off_region + static_cast<size_t>(n) * 2048 + L->payload_bytes;
  size_t off = 0;
  L->tables_off = off;  off += align_up(static_cast<size_t>(n) * 4 * gst::kTableSize * 4, kQuantum);
  L->sym_off = off;     off += align_up(static_cast<size_t>(n) * 6 * N, kQuantum);
  L->palette_off = off; off += align_up(std::max<size_t>(L->palette_total, 16), kQuantum);
  // palette indices < 2^16 everywhere -> the per-block index suffix sums are kept as u16
  L->idx16 = max_pal / 4 <= 65536u;
  L->idx_off = off;     off += align_up(static_cast<size_t>(n) * N * (L->idx16 ? 2 : 4), kQuantum);
  L->total_off = off;   off += align_up(static_cast<size_t>(n) * gst::idx_total_stride(L->groups_per_plane) * 4, kQuantum);
  L->run_off = off;     off += align_up(static_cast<size_t>(n) * (N / gst::kSymsPerLane) * 4, kQuantum);
  L->scratch_bytes = off;
  return GST_OK;
}

struct Taps {
  void *symbols = nullptr, *planes = nullptr, *indices = nullptr;
};

// inline_offsets: n == 1 and cmp_dev holds NO offset table (its first off_region bytes are not read): the eight
// offsets travel in the kernel parameters (the frame streamer uploads a file as it lies)
int decode_batch(gst_ctx *ctx, const gst_header *hdrs, uint32_t n, cudaStream_t stream, const void *cmp_dev,
                 size_t cmp_bytes, void *out_dev, int rgb, const Taps &taps, void *const *wait_events,
                 uint32_t n_wait, void **done_event, bool inline_offsets = false, bool freq_inline = false) {
  if (!ctx) return fail(GST_ERR_INVALID, "null context");
  if (!cmp_dev || !out_dev) return fail(GST_ERR_INVALID, "null device buffer");
  BatchLayout L;
  int rc = layout_batch(hdrs, n, &L);
  if (rc) return rc;
  if (cmp_bytes < L.total_cmp)
    return fail(GST_ERR_INVALID, "compressed buffer holds %zu bytes but the headers describe %zu", cmp_bytes, L.total_cmp);
  DeviceGuard guard(ctx->device);
  for (uint32_t i = 0; i < n_wait; ++i)
    GST_CUDA_TRY(cudaStreamWaitEvent(stream, static_cast<cudaEvent_t>(wait_events[i]), 0));

  // scratch: preallocated arena (bump, never reset), or the grow-only buffer this stream's calls share
  uint8_t *scratch = nullptr;
  bool from_arena = false;
  {
    std::lock_guard<std::mutex> lock(ctx->arena_mutex);
    if (ctx->arena) {
      if (ctx->arena_off + L.scratch_bytes > ctx->arena_size)
        return fail(GST_ERR_NOMEM, "preallocated scratch exhausted: need %zu more bytes, %zu of %zu used",
                    L.scratch_bytes, ctx->arena_off, ctx->arena_size);
      scratch = ctx->arena + ctx->arena_off;
      ctx->arena_off += L.scratch_bytes;
      from_arena = true;
    }
  }
  // the stream's record is held from here to the last launch of the call: a concurrent call on the same stream must
  // neither regrow the buffers in between nor interleave its kernels with ours (the hand-over counters are per stream)
  gst_ctx::StreamScratch *ssp = nullptr;
  {
    std::lock_guard<std::mutex> lock(ctx->scratch_mutex);
    std::unique_ptr<gst_ctx::StreamScratch> &slot = ctx->stream_scratch[stream];
    if (!slot) slot.reset(new gst_ctx::StreamScratch);
    ssp = slot.get();
  }
  std::unique_lock<std::mutex> scratch_lock(ssp->m);
  gst_ctx::StreamScratch &ss = *ssp;
  if (!from_arena) {
    if (ss.cap < L.scratch_bytes) {
      // stream-ordered: the old buffer is released after the work already queued on this stream
      if (ss.ptr) GST_CUDA_TRY(cudaFreeAsync(ss.ptr, stream));
      ss.ptr = nullptr;
      ss.cap = 0;
      const size_t cap = align_up(L.scratch_bytes + L.scratch_bytes / 8, 1 << 20);
      cudaError_t e = cudaMallocAsync(reinterpret_cast<void **>(&ss.ptr), cap, stream);
      if (e != cudaSuccess) return fail(GST_ERR_NOMEM, "scratch allocation of %zu bytes failed: %s", cap, cudaGetErrorString(e));
      ss.cap = cap;
    }
    scratch = ss.ptr;
  }
  uint32_t *img_done = nullptr;
  if (n > 1) {
    const size_t need = align_up(static_cast<size_t>(n), 32);  // (whole cache lines per call)
    if (ss.sync_cap < need) {
      if (ss.sync) GST_CUDA_TRY(cudaFreeAsync(ss.sync, stream));
      ss.sync = nullptr;
      ss.sync_cap = ss.sync_used = 0;
      const size_t cap = std::max<size_t>(4 * need, ctx->sync_pool_words);
      cudaError_t e = cudaMallocAsync(reinterpret_cast<void **>(&ss.sync), cap * sizeof(uint32_t), stream);
      if (e == cudaSuccess) e = cudaMemsetAsync(ss.sync, 0, cap * sizeof(uint32_t), stream);
      if (e != cudaSuccess) return fail(GST_ERR_NOMEM, "hand-over counters for %u images: %s", n, cudaGetErrorString(e));
      ss.sync_cap = cap;
    } else if (ss.sync_used + need > ss.sync_cap) {
      GST_CUDA_TRY(cudaMemsetAsync(ss.sync, 0, ss.sync_cap * sizeof(uint32_t), stream));  // after every earlier call of the stream
      ss.sync_used = 0;
    }
    img_done = ss.sync + ss.sync_used;
    ss.sync_used += need;
  }

  gst::BatchParams p{};
  p.cmp = static_cast<const uint8_t *>(cmp_dev);
  p.cmp_bytes = cmp_bytes;
  p.n_images = n;
  p.blocks_x = hdrs[0].width / 4;
  p.blocks_y = hdrs[0].height / 4;
  p.n_blocks = L.n_blocks;
  p.off_region = static_cast<uint32_t>(L.off_region);
  p.groups_per_plane = L.groups_per_plane;
  p.tables = reinterpret_cast<uint32_t *>(scratch + L.tables_off);
  p.sym_t = scratch + L.sym_off;
  p.palette = scratch + L.palette_off;
  p.palette_cap = L.palette_total;
  p.idx_s = scratch + L.idx_off;
  p.idx16 = L.idx16 ? 1u : 0u;
  p.idx_total = reinterpret_cast<int32_t *>(scratch + L.total_off);
  p.run_end = reinterpret_cast<int32_t *>(scratch + L.run_off);
  p.out = static_cast<uint8_t *>(out_dev);
  p.status = ctx->d_status;
  p.img_done = img_done;
  gst::fill_kernel_constants(&p);
  p.freq_inline = freq_inline ? 1u : 0u;
  if (inline_offsets) {
    if (n != 1) return fail(GST_ERR_INVALID, "inline offsets need a single image");
    const uint32_t N = L.n_blocks;
    const uint32_t out_sz[4] = {2 * N, 4 * N, hdrs[0].palette_bytes, N};
    const uint32_t in_sz[4] = {hdrs[0].y_cmp_sz, hdrs[0].chroma_cmp_sz, hdrs[0].palette_sz, hdrs[0].indices_sz};
    uint32_t oa = 0, ia = 0;
    for (int k = 0; k < 4; ++k) {
      p.off8[k] = oa;
      p.off8[4 + k] = ia;
      oa += out_sz[k];
      ia += in_sz[k];
    }
    p.inline_off = 1;
  }
  p.tap_symbols = static_cast<uint8_t *>(taps.symbols);
  p.tap_planes = static_cast<int8_t *>(taps.planes);
  p.tap_indices = static_cast<int32_t *>(taps.indices);

  cudaEvent_t marks[gst::kLaunchesPerBatch + 1];
  bool profiled = false;
  {
    std::lock_guard<std::mutex> lock(ctx->prof_mutex);
    profiled = ctx->prof_on;
  }
  if (profiled) {
    std::lock_guard<std::mutex> lock(ctx->prof_mutex);
    profiled = ctx->prof_events.size() < 4096 * (gst::kLaunchesPerBatch + 1);  // unread records are capped
  }
  if (profiled)
    for (auto &m : marks) GST_CUDA_TRY(cudaEventCreateWithFlags(&m, cudaEventDefault));
  cudaError_t e = gst::launch_decode_batch(p, rgb, L.max_palette, stream, profiled ? marks : nullptr);
  if (profiled) {
    std::lock_guard<std::mutex> lock(ctx->prof_mutex);
    ctx->prof_events.insert(ctx->prof_events.end(), marks, marks + gst::kLaunchesPerBatch + 1);
  }
  if (e != cudaSuccess) return fail(GST_ERR_CUDA, "decode launch failed: %s", cudaGetErrorString(e));
  if (done_event) {
    cudaEvent_t ev;
    GST_CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDefault));
    GST_CUDA_TRY(cudaEventRecord(ev, stream));
    *done_event = ev;
  }
  return GST_OK;
}

// the context's workspace for the standalone rANS entry points (caller holds ctx->ans_mutex)
int ans_workspace(gst_ctx *ctx, size_t bytes) {
  if (bytes <= ctx->ans_ws_cap) return GST_OK;
  if (ctx->ans_ws) {
    cudaStreamSynchronize(ctx->streams[0]);
    cudaFree(ctx->ans_ws);
    ctx->ans_ws = nullptr;
    ctx->ans_ws_cap = 0;
  }
  const size_t cap = align_up(bytes + bytes / 4, 1 << 16);
  cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&ctx->ans_ws), cap);
  if (e != cudaSuccess) return fail(GST_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", cap, cudaGetErrorString(e));
  ctx->ans_ws_cap = cap;
  return GST_OK;
}

// ans::GenerateHistogram (ans/histogram.cpp:41-123), restated: scale counts to sum M with
// the float rounding rule of the reference, then fix the residual one unit at a time on the
// symbol whose code-length cost changes least (min-heap on the rank).
struct RankedSymbol {
  int symbol;
  double rank;
  bool operator>(const RankedSymbol &o) const { return rank > o.rank; }
};

double freq_change(int count, int new_count, int sign) {
  return std::log2(static_cast<double>(new_count) / static_cast<double>(new_count + sign)) * static_cast<double>(count);
}

int normalize_frequencies(const uint32_t *counts, uint32_t n, int M, std::vector<uint32_t> *out) {
  out->assign(n, 0);
  int sum = 0;
  for (uint32_t i = 0; i < n; ++i) sum = static_cast<int>(static_cast<uint32_t>(sum) + counts[i]);
  if (sum == 0) return fail(GST_ERR_INVALID, "no symbol has a non-zero count");
  int total = 0;
  for (uint32_t i = 0; i < n; ++i) {
    if (counts[i] == 0) continue;
    const double scaled = static_cast<float>(counts[i] * static_cast<uint32_t>(M)) / static_cast<float>(sum);
    const int down = static_cast<int>(scaled);
    const int pick = (scaled * scaled <= static_cast<double>(down * (down + 1))) ? down : down + 1;
    (*out)[i] = static_cast<uint32_t>(std::max(1, pick));
    total += static_cast<int>((*out)[i]);
  }
  int correction = M - total;
  if (correction == 0) return GST_OK;
  const int sign = correction > 0 ? 1 : -1;
  std::vector<RankedSymbol> heap;
  for (uint32_t i = 0; i < n; ++i) {
    if (counts[i] == 0) continue;
    if ((*out)[i] > 1 || correction > 0)
      heap.push_back({static_cast<int>(i), freq_change(static_cast<int>(counts[i]), static_cast<int>((*out)[i]), sign)});
  }
  std::make_heap(heap.begin(), heap.end(), std::greater<RankedSymbol>());
  while (correction != 0) {
    if (heap.empty()) return fail(GST_ERR_INVALID, "cannot normalise frequencies to %d", M);
    std::pop_heap(heap.begin(), heap.end(), std::greater<RankedSymbol>());
    const RankedSymbol s = heap.back();
    heap.pop_back();
    const int i = s.symbol;
    (*out)[i] = static_cast<uint32_t>(static_cast<int>((*out)[i]) + sign);
    correction -= sign;
    if ((*out)[i] > 1 || sign == 1) {
      heap.push_back({i, freq_change(static_cast<int>(counts[i]), static_cast<int>((*out)[i]), sign)});
      std::push_heap(heap.begin(), heap.end(), std::greater<RankedSymbol>());
    }
  }
  return GST_OK;
}

}  // namespace

// ======================================================================================
extern "C" {

const char *gst_last_error(void) { return g_err; }

int gst_ctx_create(int device, gst_ctx **out) {
  if (!out) return fail(GST_ERR_INVALID, "null out pointer");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(GST_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (device < 0 || device >= count) return fail(GST_ERR_INVALID, "device ordinal %d out of range [0,%d)", device, count);
  cudaDeviceProp prop;
  GST_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(GST_ERR_NO_DEVICE, "device %d (%s) is sm_%d%d; this build contains sm_100a code only", device,
                prop.name, prop.major, prop.minor);
  DeviceGuard guard(device);
  gst_ctx *ctx = new (std::nothrow) gst_ctx;
  if (!ctx) return fail(GST_ERR_NOMEM, "out of host memory");
  ctx->device = device;
  if (const char *v = getenv("GST_SYNC_POOL_WORDS")) {
    const long w = atol(v);
    if (w >= 32) ctx->sync_pool_words = static_cast<size_t>(w);
  }
  for (auto &s : ctx->streams) {
    e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
      gst_ctx_destroy(ctx);
      return fail(GST_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(e));
    }
  }
  for (auto &s : ctx->host_streams) {
    e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
      gst_ctx_destroy(ctx);
      return fail(GST_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(e));
    }
  }
  e = cudaMalloc(reinterpret_cast<void **>(&ctx->d_status), 16);
  if (e == cudaSuccess) e = cudaMemset(ctx->d_status, 0, 16);
  if (e != cudaSuccess) {
    gst_ctx_destroy(ctx);
    return fail(GST_ERR_CUDA, "status word allocation failed: %s", cudaGetErrorString(e));
  }
  // keep freed scratch in the pool so per-call cudaMallocAsync does not hit the OS
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t threshold = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
  }
  *out = ctx;
  return GST_OK;
}

void gst_ctx_destroy(gst_ctx *ctx) {
  if (!ctx) return;
  DeviceGuard guard(ctx->device);
  for (auto &s : ctx->streams)
    if (s) {
      cudaStreamSynchronize(s);
      cudaStreamDestroy(s);
    }
  for (auto &s : ctx->host_streams)
    if (s) {
      cudaStreamSynchronize(s);
      cudaStreamDestroy(s);
    }
  if (ctx->arena) cudaFree(ctx->arena);
  for (auto &kv : ctx->stream_scratch)
    if (kv.second) kv.second->release();
  if (ctx->ans_ws) cudaFree(ctx->ans_ws);
  if (ctx->d_status) cudaFree(ctx->d_status);
  for (auto &hs : ctx->host_slots) {
    for (int k = 0; k < 2; ++k) {
      if (hs.pinned[k]) cudaFreeHost(hs.pinned[k]);
      if (hs.h2d_done[k]) cudaEventDestroy(hs.h2d_done[k]);
    }
    if (hs.d_in) cudaFree(hs.d_in);
    if (hs.d_out) cudaFree(hs.d_out);
  }
  for (auto ev : ctx->prof_events) cudaEventDestroy(ev);
  delete ctx;
}

int gst_ctx_device(const gst_ctx *ctx) { return ctx ? ctx->device : -1; }

void *gst_stream_default(gst_ctx *ctx) { return ctx ? ctx->streams[0] : nullptr; }

void *gst_stream_next(gst_ctx *ctx) {
  if (!ctx) return nullptr;
  return ctx->streams[1 + ctx->next_stream.fetch_add(1) % kNumWorkStreams];
}

int gst_ctx_sync(gst_ctx *ctx) {
  if (!ctx) return fail(GST_ERR_INVALID, "null context");
  DeviceGuard guard(ctx->device);
  for (auto &s : ctx->streams) GST_CUDA_TRY(cudaStreamSynchronize(s));
  return GST_OK;
}

int gst_stream_sync(gst_ctx *ctx, void *stream) {
  if (!ctx) return fail(GST_ERR_INVALID, "null context");
  DeviceGuard guard(ctx->device);
  GST_CUDA_TRY(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  return GST_OK;
}

int gst_malloc(gst_ctx *ctx, size_t bytes, void **dptr) {
  if (!ctx || !dptr) return fail(GST_ERR_INVALID, "null argument");
  DeviceGuard guard(ctx->device);
  cudaError_t e = cudaMalloc(dptr, align_up(std::max<size_t>(bytes, 1), 256));
  if (e != cudaSuccess) return fail(GST_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
  return GST_OK;
}

int gst_free(gst_ctx *ctx, void *dptr) {
  if (!ctx) return fail(GST_ERR_INVALID, "null context");
  DeviceGuard guard(ctx->device);
  GST_CUDA_TRY(cudaFree(dptr));
  return GST_OK;
}

int gst_host_alloc(gst_ctx *ctx, size_t bytes, void **hptr) {
  if (!ctx || !hptr) return fail(GST_ERR_INVALID, "null argument");
  DeviceGuard guard(ctx->device);
  cudaError_t e = cudaHostAlloc(hptr, std::max<size_t>(bytes, 1), cudaHostAllocDefault);
  if (e != cudaSuccess) return fail(GST_ERR_NOMEM, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
  return GST_OK;
}

int gst_host_free(gst_ctx *ctx, void *hptr) {
  if (!ctx) return fail(GST_ERR_INVALID, "null context");
  DeviceGuard guard(ctx->device);
  GST_CUDA_TRY(cudaFreeHost(hptr));
  return GST_OK;
}

int gst_upload_async(gst_ctx *ctx, void *stream, void *dst_dev, const void *src_host, size_t bytes) {
  if (!ctx) return fail(GST_ERR_INVALID, "null context");
  DeviceGuard guard(ctx->device);
  GST_CUDA_TRY(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
  return GST_OK;
}

int gst_download_async(gst_ctx *ctx, void *stream, void *dst_host, const void *src_dev, size_t bytes) {
  if (!ctx) return fail(GST_ERR_INVALID, "null context");
  DeviceGuard guard(ctx->device);
  GST_CUDA_TRY(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
  return GST_OK;
}

int gst_download_2d_async(gst_ctx *ctx, void *stream, void *dst_host, size_t dst_pitch, const void *src_dev,
                          size_t src_pitch, size_t width_bytes, size_t rows) {
  if (!ctx) return fail(GST_ERR_INVALID, "null context");
  DeviceGuard guard(ctx->device);
  GST_CUDA_TRY(cudaMemcpy2DAsync(dst_host, dst_pitch, src_dev, src_pitch, width_bytes, rows, cudaMemcpyDeviceToHost,
                                 static_cast<cudaStream_t>(stream)));
  return GST_OK;
}

int gst_memset_async(gst_ctx *ctx, void *stream, void *dst_dev, int value, size_t bytes) {
  if (!ctx) return fail(GST_ERR_INVALID, "null context");
  DeviceGuard guard(ctx->device);
  GST_CUDA_TRY(cudaMemsetAsync(dst_dev, value, bytes, static_cast<cudaStream_t>(stream)));
  return GST_OK;
}

int gst_event_record(gst_ctx *ctx, void *stream, void **event_out) {
  if (!ctx || !event_out) return fail(GST_ERR_INVALID, "null argument");
  DeviceGuard guard(ctx->device);
  cudaEvent_t ev;
  GST_CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDefault));
  cudaError_t e = cudaEventRecord(ev, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) {
    cudaEventDestroy(ev);
    return fail(GST_ERR_CUDA, "cudaEventRecord failed: %s", cudaGetErrorString(e));
  }
  *event_out = ev;
  return GST_OK;
}

int gst_event_wait(void *event) {
  GST_CUDA_TRY(cudaEventSynchronize(static_cast<cudaEvent_t>(event)));
  return GST_OK;
}

int gst_event_elapsed_ms(void *start, void *stop, float *ms) {
  if (!ms) return fail(GST_ERR_INVALID, "null argument");
  GST_CUDA_TRY(cudaEventElapsedTime(ms, static_cast<cudaEvent_t>(start), static_cast<cudaEvent_t>(stop)));
  return GST_OK;
}

void gst_event_destroy(void *event) {
  if (event) cudaEventDestroy(static_cast<cudaEvent_t>(event));
}

// ---- stream format ---------------------------------------------------------------------
int gst_parse_header(const uint8_t *gst, size_t len, gst_header *hdr) {
  if (!gst || !hdr) return fail(GST_ERR_INVALID, "null argument");
  if (len < GST_HEADER_BYTES + 4 * 512) return fail(GST_ERR_INVALID, "file of %zu bytes is too short for a .gst header", len);
  memcpy(hdr, gst, GST_HEADER_BYTES);
  int rc = check_header(*hdr);
  if (rc) return rc;
  const uint64_t need = static_cast<uint64_t>(GST_HEADER_BYTES) + 2048 + hdr->y_cmp_sz + hdr->chroma_cmp_sz +
                        hdr->palette_sz + hdr->indices_sz;
  if (need > len) return fail(GST_ERR_INVALID, "header describes %llu bytes but the file has %zu", (unsigned long long)need, len);
  return GST_OK;
}

size_t gst_packed_size(const gst_header *hdrs, uint32_t n) {
  BatchLayout L;
  if (layout_batch(hdrs, n, &L)) return 0;
  return L.total_cmp;
}

namespace {
// The device input layout of LoadCompressedDXTs (codec/decoder.cpp:430-476, demo/photos_sf.cpp:753-795):
// writes the offsets region (and, with copy_payload, the frequency tables and the payloads) to dst.
int pack_impl(const uint8_t *const *gst_files, const size_t *lens, uint32_t n, uint8_t *dst, size_t dst_cap,
              gst_header *hdrs_out, bool copy_payload, BatchLayout *L_out) {
  if (!gst_files || !lens || !dst || !hdrs_out || n == 0) return fail(GST_ERR_INVALID, "null or empty argument");
  for (uint32_t i = 0; i < n; ++i) {
    int rc = gst_parse_header(gst_files[i], lens[i], &hdrs_out[i]);
    if (rc) return rc;
  }
  BatchLayout L;
  int rc = layout_batch(hdrs_out, n, &L);
  if (rc) return rc;
  if (L_out) *L_out = L;
  const size_t need = copy_payload ? L.total_cmp : L.off_region;
  if (dst_cap < need) return fail(GST_ERR_SMALL, "packed batch needs %zu bytes, buffer has %zu", need, dst_cap);
  uint32_t *out_off = reinterpret_cast<uint32_t *>(dst);
  uint32_t *in_off = out_off + 4 * n;
  memset(dst, 0, L.off_region);
  uint8_t *freqs = dst + L.off_region;
  uint8_t *payload = freqs + static_cast<size_t>(n) * 2048;
  uint32_t in_acc = 0, out_acc = 0;
  for (uint32_t i = 0; i < n; ++i) {
    const gst_header &h = hdrs_out[i];
    const uint32_t N = L.n_blocks;
    const uint32_t in_sz[4] = {h.y_cmp_sz, h.chroma_cmp_sz, h.palette_sz, h.indices_sz};
    const uint32_t out_sz[4] = {2 * N, 4 * N, h.palette_bytes, N};
    if (copy_payload) {
      const uint8_t *src = gst_files[i] + GST_HEADER_BYTES;
      memcpy(freqs + static_cast<size_t>(i) * 2048, src, 2048);
      const size_t body = static_cast<size_t>(in_sz[0]) + in_sz[1] + in_sz[2] + in_sz[3];
      memcpy(payload + in_acc, src + 2048, body);
    }
    for (int s = 0; s < 4; ++s) {
      in_off[4 * i + s] = in_acc;
      out_off[4 * i + s] = out_acc;
      in_acc += in_sz[s];
      out_acc += out_sz[s];
    }
  }
  return GST_OK;
}

}  // namespace

int gst_pack_batch(const uint8_t *const *gst_files, const size_t *lens, uint32_t n, uint8_t *dst, size_t dst_cap,
                   gst_header *hdrs_out) {
  return pack_impl(gst_files, lens, n, dst, dst_cap, hdrs_out, true, nullptr);
}

// ---- scratch ---------------------------------------------------------------------------
size_t gst_required_scratch(const gst_header *hdr) {
  if (!hdr) return 0;
  // codec/decoder.cpp:41-47
  return 4 * static_cast<size_t>(gst::kTableSize) * 6 + 17 * static_cast<size_t>(hdr->width) * hdr->height / 16 +
         hdr->palette_bytes;
}

int gst_preallocate(gst_ctx *ctx, size_t bytes) {
  if (!ctx) return fail(GST_ERR_INVALID, "null context");
  DeviceGuard guard(ctx->device);
  std::lock_guard<std::mutex> lock(ctx->arena_mutex);
  if (ctx->arena) {
    GST_CUDA_TRY(cudaFree(ctx->arena));
    ctx->arena = nullptr;
  }
  ctx->arena_size = ctx->arena_off = 0;
  // every region is rounded up to the 512-byte quantum: leave room for that
  const size_t padded = align_up(bytes, kQuantum) + 16 * kQuantum;
  cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&ctx->arena), padded);
  if (e != cudaSuccess) return fail(GST_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", padded, cudaGetErrorString(e));
  ctx->arena_size = padded;
  return GST_OK;
}

int gst_free_scratch(gst_ctx *ctx) {
  if (!ctx) return fail(GST_ERR_INVALID, "null context");
  DeviceGuard guard(ctx->device);
  GST_CUDA_TRY(cudaDeviceSynchronize());
  {
    std::lock_guard<std::mutex> lock(ctx->scratch_mutex);
    for (auto &kv : ctx->stream_scratch)
      if (kv.second) kv.second->release();
    ctx->stream_scratch.clear();
  }
  std::lock_guard<std::mutex> lock(ctx->arena_mutex);
  if (ctx->arena) GST_CUDA_TRY(cudaFree(ctx->arena));
  ctx->arena = nullptr;
  ctx->arena_size = ctx->arena_off = 0;
  return GST_OK;
}

// ---- decode ----------------------------------------------------------------------------
int gst_load_dxt_batch(gst_ctx *ctx, const gst_header *hdrs, uint32_t n, void *stream, const void *cmp_dev,
                       size_t cmp_bytes, void *out_dev, void *const *wait_events, uint32_t n_wait, void **done_event) {
  return decode_batch(ctx, hdrs, n, static_cast<cudaStream_t>(stream), cmp_dev, cmp_bytes, out_dev, 0, Taps{},
                      wait_events, n_wait, done_event);
}

int gst_load_rgb_batch(gst_ctx *ctx, const gst_header *hdrs, uint32_t n, void *stream, const void *cmp_dev,
                       size_t cmp_bytes, void *out_dev, void *const *wait_events, uint32_t n_wait, void **done_event) {
  return decode_batch(ctx, hdrs, n, static_cast<cudaStream_t>(stream), cmp_dev, cmp_bytes, out_dev, 1, Taps{},
                      wait_events, n_wait, done_event);
}

int gst_load_dxt_batch_tapped(gst_ctx *ctx, const gst_header *hdrs, uint32_t n, void *stream, const void *cmp_dev,
                              size_t cmp_bytes, void *out_dev, void *symbols_dev, void *planes_dev, void *indices_dev) {
  Taps t;
  t.symbols = symbols_dev;
  t.planes = planes_dev;
  t.indices = indices_dev;
  return decode_batch(ctx, hdrs, n, static_cast<cudaStream_t>(stream), cmp_dev, cmp_bytes, out_dev, 0, t, nullptr, 0,
                      nullptr);
}

// Host-to-host batch decode.  The batch is cut into pages (demo/photos_sf.cpp:688); a page is handled by one host
// thread with a staging slot and stream drawn from the context's pool: the thread writes the page's offset table,
// uploads every file in one copy (straight from the caller's buffer when that is pinned, through the slot's pinned
// staging otherwise -- two buffers per slot, so staging page k+1 overlaps the transfers of page k), then enqueues
// decode -> D2H on its stream.  Staging lives in the context and only grows.
namespace {
int host_batch(gst_ctx *ctx, const uint8_t *const *gst_files, const size_t *lens, uint32_t n, uint32_t page, int mode,
               uint8_t *out, size_t out_cap, bool out_on_device);
}

int gst_decompress_host_batch(gst_ctx *ctx, const uint8_t *const *gst_files, const size_t *lens, uint32_t n,
                              uint32_t page, int mode, uint8_t *out, size_t out_cap) {
  return host_batch(ctx, gst_files, lens, n, page, mode, out, out_cap, false);
}

int gst_load_host_batch(gst_ctx *ctx, const uint8_t *const *gst_files, const size_t *lens, uint32_t n,
                        uint32_t page, int mode, void *out_dev, size_t out_cap) {
  return host_batch(ctx, gst_files, lens, n, page, mode, static_cast<uint8_t *>(out_dev), out_cap, true);
}

namespace {
int host_batch(gst_ctx *ctx, const uint8_t *const *gst_files, const size_t *lens, uint32_t n, uint32_t page, int mode,
               uint8_t *out, size_t out_cap, bool out_on_device) {
  if (!ctx || !gst_files || !lens || !out || n == 0) return fail(GST_ERR_INVALID, "null or empty argument");
  if (page == 0 || page > n) page = n;
  gst_header h0;
  int rc = gst_parse_header(gst_files[0], lens[0], &h0);
  if (rc) return rc;
  const size_t per_image = mode ? static_cast<size_t>(h0.width) * h0.height * 3 : static_cast<size_t>(h0.width) * h0.height / 2;
  if (out_cap < per_image * n) return fail(GST_ERR_SMALL, "output needs %zu bytes, buffer has %zu", per_image * n, out_cap);

  const uint32_t n_pages = (n + page - 1) / page;
  // textures staying on the device: the host-side page packing is the bottleneck, use every worker; textures
  // coming back to the host: the D2H copies own the link and the host memory system, four workers are enough
  // (measured: 8 workers +8..20 % / -2 % on the two paths)
  const uint32_t n_slots = std::min<uint32_t>(out_on_device ? kHostSlots : kNumWorkStreams, n_pages);
  std::vector<int> slot_rc(n_slots, GST_OK);
  std::vector<std::string> slot_err(n_slots);

  auto worker = [&](uint32_t si) {
    cudaSetDevice(ctx->device);
    // take a free staging slot (and its stream) from the context's pool; give it back when this worker is done
    uint32_t slot_id = 0;
    {
      std::unique_lock<std::mutex> lock(ctx->slot_mutex);
      ctx->slot_cv.wait(lock, [&] {
        for (uint32_t k = 0; k < kHostSlots; ++k)
          if (!ctx->slot_busy[k]) return true;
        return false;
      });
      while (ctx->slot_busy[slot_id]) ++slot_id;
      ctx->slot_busy[slot_id] = true;
    }
    struct Release {
      gst_ctx *ctx;
      uint32_t id;
      ~Release() {
        {
          std::lock_guard<std::mutex> lock(ctx->slot_mutex);
          ctx->slot_busy[id] = false;
        }
        ctx->slot_cv.notify_one();
      }
    } release{ctx, slot_id};
    HostSlot &s = ctx->host_slots[slot_id];
    cudaStream_t stream = ctx->host_streams[slot_id];
    std::vector<gst_header> hdrs(page);
    auto bail = [&](int code, const char *what, cudaError_t e) {
      slot_rc[si] = code;
      slot_err[si] = std::string(what) + (e != cudaSuccess ? std::string(": ") + cudaGetErrorString(e) : std::string());
    };
    uint32_t turn = 0;
    for (uint32_t pg = si; pg < n_pages; pg += n_slots, ++turn) {
      const uint32_t first = pg * page, cnt = std::min(page, n - first);
      size_t raw = 0;
      for (uint32_t i = 0; i < cnt; ++i) raw += lens[first + i];
      const size_t need = align_up(static_cast<size_t>(cnt) * 32, kQuantum) + raw;  // >= packed size
      const int j = turn & 1;
      cudaError_t e = cudaSuccess;
      if (need > s.cap_in) {
        // grow: drain the stream first, both staging buffers and the device input follow
        e = cudaStreamSynchronize(stream);
        for (int k = 0; k < 2; ++k) {
          if (s.pinned[k]) cudaFreeHost(s.pinned[k]);
          s.pinned[k] = nullptr;
        }
        if (s.d_in) cudaFree(s.d_in);
        s.d_in = nullptr;
        s.cap_in = align_up(need + need / 4, 4096);
        for (int k = 0; k < 2 && e == cudaSuccess; ++k)
          e = cudaHostAlloc(reinterpret_cast<void **>(&s.pinned[k]), s.cap_in, cudaHostAllocDefault);
        if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&s.d_in), s.cap_in);
      }
      if (e == cudaSuccess && !out_on_device && per_image * page > s.cap_out) {
        e = cudaStreamSynchronize(stream);
        if (s.d_out) cudaFree(s.d_out);
        s.d_out = nullptr;
        s.cap_out = per_image * page;
        if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&s.d_out), s.cap_out);
      }
      for (int k = 0; k < 2 && e == cudaSuccess; ++k)
        if (!s.h2d_done[k]) e = cudaEventCreateWithFlags(&s.h2d_done[k], cudaEventDisableTiming);
      if (e != cudaSuccess) return bail(GST_ERR_CUDA, "staging allocation failed", e);
      // the previous upload out of this staging buffer must have left the host
      e = cudaEventSynchronize(s.h2d_done[j]);
      if (e != cudaSuccess) return bail(GST_ERR_CUDA, "staging wait failed", e);
      // The page on the device (demo/photos_sf.cpp:753-806 packs [offsets][all frequency blocks][all payloads]; here
      // the frequency blocks stay in front of their image's streams -- BatchParams::freq_inline -- so that a file is
      // ONE copy): [out_off[4n]][in_off[4n]] padded to 512 | file 0 minus its header | file 1 minus its header ...
      // By default the files are copied into the slot's pinned staging and the page goes up in one DMA (on one GPU
      // 1024 copies of 640 KB cost more in per-copy latency than the packing pass they save: -4 % host to host,
      // -10 % host to device).  With gst_ctx_set_direct_upload a pinned file is DMA-ed from where it lies and the
      // host writes only the offsets region -- for boxes whose host memory, not the link, is the bottleneck.
      for (uint32_t i = 0; i < cnt; ++i) {
        int prc = gst_parse_header(gst_files[first + i], lens[first + i], &hdrs[i]);
        if (prc) {
          slot_rc[si] = prc;
          slot_err[si] = g_err;
          return;
        }
      }
      BatchLayout L;
      int prc = layout_batch(hdrs.data(), cnt, &L);
      if (prc) {
        slot_rc[si] = prc;
        slot_err[si] = g_err;
        return;
      }
      // the output stride and the staging were sized from the first file of the call: every page must match it
      // (inside a page layout_batch has already checked the images against each other)
      if (hdrs[0].width != h0.width || hdrs[0].height != h0.height) {
        slot_rc[si] = GST_ERR_INVALID;
        char msg[160];
        snprintf(msg, sizeof msg, "image %u is %ux%u but image 0 is %ux%u: one call needs equal dimensions", first,
                 hdrs[0].width, hdrs[0].height, h0.width, h0.height);
        slot_err[si] = msg;
        cudaStreamSynchronize(stream);
        return;
      }
      uint8_t *stage = s.pinned[j];
      uint32_t *out_off = reinterpret_cast<uint32_t *>(stage), *in_off = out_off + 4 * cnt;
      memset(stage, 0, L.off_region);
      uint32_t in_acc = 0, out_acc = 0;
      std::vector<uint32_t> base(cnt);
      for (uint32_t i = 0; i < cnt; ++i) {
        const gst_header &h = hdrs[i];
        const uint32_t in_sz[4] = {h.y_cmp_sz, h.chroma_cmp_sz, h.palette_sz, h.indices_sz};
        const uint32_t out_sz[4] = {2 * L.n_blocks, 4 * L.n_blocks, h.palette_bytes, L.n_blocks};
        base[i] = in_acc;
        in_acc += 2048;  // the image's four frequency blocks
        for (int k = 0; k < 4; ++k) {
          in_off[4 * i + k] = in_acc;
          out_off[4 * i + k] = out_acc;
          in_acc += in_sz[k];
          out_acc += out_sz[k];
        }
      }
      const bool direct = ctx->direct_upload.load(std::memory_order_relaxed);
      size_t staged_to = L.off_region;  // bytes of `stage` that have to be uploaded in one piece
      std::vector<uint32_t> dma;         // files that are DMA-ed from where they lie
      for (uint32_t i = 0; i < cnt; ++i) {
        const uint8_t *src = gst_files[first + i] + GST_HEADER_BYTES;
        const size_t body = static_cast<size_t>(in_off[4 * i + 3]) + hdrs[i].indices_sz - base[i];
        bool pinned = false;
        if (direct) {
          cudaPointerAttributes attr;
          pinned = cudaPointerGetAttributes(&attr, src) == cudaSuccess && attr.type == cudaMemoryTypeHost;
          if (!pinned) cudaGetLastError();  // (older drivers report an unregistered pointer as an error)
        }
        if (pinned) {
          dma.push_back(i);
        } else {
          memcpy(stage + L.off_region + base[i], src, body);
          staged_to = L.off_region + base[i] + body;
        }
      }
      // one copy for the offsets region and everything staged behind it (the whole page by default), one per file
      // that is uploaded directly
      e = cudaMemcpyAsync(s.d_in, stage, dma.empty() ? staged_to : L.off_region, cudaMemcpyHostToDevice, stream);
      if (!dma.empty()) {
        for (uint32_t i = 0; i < cnt && e == cudaSuccess; ++i) {
          const size_t body = static_cast<size_t>(in_off[4 * i + 3]) + hdrs[i].indices_sz - base[i];
          const bool is_dma = std::find(dma.begin(), dma.end(), i) != dma.end();
          const uint8_t *src = is_dma ? gst_files[first + i] + GST_HEADER_BYTES : stage + L.off_region + base[i];
          e = cudaMemcpyAsync(s.d_in + L.off_region + base[i], src, body, cudaMemcpyHostToDevice, stream);
        }
      }
      if (e == cudaSuccess) e = cudaEventRecord(s.h2d_done[j], stream);
      if (e != cudaSuccess) return bail(GST_ERR_CUDA, "upload failed", e);
      // the textures either stay in the caller's device buffer (LoadCompressedDXTs into a PBO,
      // demo/photos_sf.cpp:810-821) or come back to the host (DecompressDXT)
      prc = decode_batch(ctx, hdrs.data(), cnt, stream, s.d_in, s.cap_in, out_on_device ? out + per_image * first : s.d_out,
                         mode, Taps{}, nullptr, 0, nullptr, false, true);
      if (prc) {
        slot_rc[si] = prc;
        slot_err[si] = g_err;
        return;
      }
      if (!out_on_device) {
        e = cudaMemcpyAsync(out + per_image * first, s.d_out, per_image * cnt, cudaMemcpyDeviceToHost, stream);
        if (e != cudaSuccess) return bail(GST_ERR_CUDA, "download failed", e);
      }
    }
    cudaError_t e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) bail(GST_ERR_CUDA, "decode failed", e);
  };

  if (n_slots == 1) {
    DeviceGuard guard(ctx->device);
    worker(0);
  } else {
    std::vector<std::thread> pool;
    for (uint32_t si = 0; si < n_slots; ++si) pool.emplace_back(worker, si);
    for (auto &t : pool) t.join();
  }
  for (uint32_t si = 0; si < n_slots; ++si)
    if (slot_rc[si]) return fail(slot_rc[si], "%s", slot_err[si].c_str());
  return GST_OK;
}
}  // namespace

int gst_decompress_host(gst_ctx *ctx, const uint8_t *gst, size_t len, int mode, uint8_t *out, size_t out_cap) {
  const uint8_t *files[1] = {gst};
  const size_t lens[1] = {len};
  return gst_decompress_host_batch(ctx, files, lens, 1, 1, mode, out, out_cap);
}

// ---- frame streamer ---------------------------------------------------------------------
int gst_streamer_create(gst_ctx *ctx, uint32_t width, uint32_t height, uint32_t depth, int mode, gst_streamer **out) {
  if (!ctx || !out) return fail(GST_ERR_INVALID, "null argument");
  if (depth == 0 || depth > 64) return fail(GST_ERR_INVALID, "depth must be 1..64");
  gst_header h{};
  h.width = width;
  h.height = height;
  int rc = check_dims(h);
  if (rc) return rc;
  DeviceGuard guard(ctx->device);
  gst_streamer *st = new (std::nothrow) gst_streamer;
  if (!st) return fail(GST_ERR_NOMEM, "out of host memory");
  st->ctx = ctx;
  st->width = width;
  st->height = height;
  st->depth = depth;
  st->mode = mode ? 1 : 0;
  st->frame_bytes = mode ? static_cast<size_t>(width) * height * 3 : static_cast<size_t>(width) * height / 2;
  st->slots.resize(depth);
  for (auto &s : st->slots) {
    cudaError_t e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&s.d_out), st->frame_bytes);
    if (e == cudaSuccess) s.cap_out = st->frame_bytes;
    if (e != cudaSuccess) {
      gst_streamer_destroy(st);
      return fail(GST_ERR_CUDA, "streamer allocation failed: %s", cudaGetErrorString(e));
    }
  }
  *out = st;
  return GST_OK;
}

void gst_streamer_destroy(gst_streamer *st) {
  if (!st) return;
  DeviceGuard guard(st->ctx->device);
  for (auto &s : st->slots) {
    if (s.stream) {
      cudaStreamSynchronize(s.stream);
      {  // the scratch the context kept for this stream goes with it
        std::lock_guard<std::mutex> lock(st->ctx->scratch_mutex);
        auto it = st->ctx->stream_scratch.find(s.stream);
        if (it != st->ctx->stream_scratch.end()) {
          if (it->second) it->second->release();
          st->ctx->stream_scratch.erase(it);
        }
      }
      cudaStreamDestroy(s.stream);
    }
    if (s.done) cudaEventDestroy(s.done);
    if (s.pinned) cudaFreeHost(s.pinned);
    if (s.d_in) cudaFree(s.d_in);
    if (s.d_out) cudaFree(s.d_out);
  }
  delete st;
}

namespace {
// One frame: [512 unused bytes][file minus its 28-byte header] on the device -- the frequency tables and the four
// streams lie in the file exactly as the decode kernels want them, and the eight offsets of a single image travel in
// the kernel parameters, so nothing is packed.  direct: the copy reads the caller's buffer (which must stay valid,
// and should be pinned, until the frame is waited for); otherwise the frame goes through the slot's pinned staging.
int streamer_submit(gst_streamer *st, const uint8_t *gst, size_t len, void *out_dev, void *out_host, uint64_t *ticket, bool direct) {
  if (!st || !gst) return fail(GST_ERR_INVALID, "null argument");
  gst_header h;
  int rc = gst_parse_header(gst, len, &h);
  if (rc) return rc;
  if (h.width != st->width || h.height != st->height)
    return fail(GST_ERR_INVALID, "frame is %ux%u, the streamer was created for %ux%u", h.width, h.height, st->width, st->height);
  DeviceGuard guard(st->ctx->device);
  const uint64_t t = st->next_ticket;
  gst_streamer::Slot &s = st->slots[t % st->depth];
  // the slot's previous frame (ticket t - depth) must have been decoded: its staging is reused
  if (s.ticket != UINT64_MAX) GST_CUDA_TRY(cudaEventSynchronize(s.done));
  const size_t body = len - GST_HEADER_BYTES;
  const size_t need = kQuantum + body;
  if (need > s.cap_in) {
    if (s.pinned) cudaFreeHost(s.pinned);
    if (s.d_in) cudaFree(s.d_in);
    s.pinned = s.d_in = nullptr;
    s.cap_in = align_up(need + need / 2, 4096);
    GST_CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&s.pinned), s.cap_in, cudaHostAllocDefault));
    GST_CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&s.d_in), s.cap_in));
  }
  const uint8_t *src = gst + GST_HEADER_BYTES;
  if (!direct) {
    memcpy(s.pinned, src, body);
    src = s.pinned;
  }
  GST_CUDA_TRY(cudaMemcpyAsync(s.d_in + kQuantum, src, body, cudaMemcpyHostToDevice, s.stream));  // demo/demo.cpp:189-192
  void *dst = out_dev ? out_dev : s.d_out;
  rc = decode_batch(st->ctx, &h, 1, s.stream, s.d_in, s.cap_in, dst, st->mode, Taps{}, nullptr, 0, nullptr, true);
  if (rc) return rc;
  // read-back on the slot's own stream: the slot's next frame is ordered behind it, and `done` covers it
  if (out_host) GST_CUDA_TRY(cudaMemcpyAsync(out_host, dst, st->frame_bytes, cudaMemcpyDeviceToHost, s.stream));
  GST_CUDA_TRY(cudaEventRecord(s.done, s.stream));
  s.ticket = t;
  s.frame_dev = dst;
  st->next_ticket = t + 1;
  if (ticket) *ticket = t;
  return GST_OK;
}
}  // namespace

int gst_streamer_submit(gst_streamer *st, const uint8_t *gst, size_t len, void *out_dev, uint64_t *ticket) {
  return streamer_submit(st, gst, len, out_dev, nullptr, ticket, false);
}

int gst_streamer_submit_ex(gst_streamer *st, const uint8_t *gst, size_t len, void *out_dev, void *out_host,
                           uint32_t flags, uint64_t *ticket) {
  if (flags & ~static_cast<uint32_t>(GST_SUBMIT_DIRECT)) return fail(GST_ERR_INVALID, "unknown submit flags 0x%x", flags);
  return streamer_submit(st, gst, len, out_dev, out_host, ticket, (flags & GST_SUBMIT_DIRECT) != 0);
}

namespace {
// k consecutive frames as ONE LoadCompressedDXTs-style call on slot `slot_no` (gst_streamer_play): k uploads (straight from
// the caller's buffers when `direct`), the offsets table, two kernel launches, ONE read-back of the k frames.
int streamer_submit_group(gst_streamer *st, uint32_t slot_no, const uint8_t *const *frames, const size_t *lens, uint32_t k,
                          uint8_t *out_dev, uint8_t *out_host, bool direct) {
  gst_streamer::Slot &s = st->slots[slot_no];
  std::vector<gst_header> hdrs(k);
  size_t body_total = 0;
  for (uint32_t i = 0; i < k; ++i) {
    int rc = gst_parse_header(frames[i], lens[i], &hdrs[i]);
    if (rc) return rc;
    if (hdrs[i].width != st->width || hdrs[i].height != st->height)
      return fail(GST_ERR_INVALID, "frame is %ux%u, the streamer was created for %ux%u", hdrs[i].width, hdrs[i].height,
                  st->width, st->height);
    body_total += lens[i] - GST_HEADER_BYTES;
  }
  BatchLayout L;
  int rc = layout_batch(hdrs.data(), k, &L);
  if (rc) return rc;
  // the slot's previous work must be done: its staging and buffers are reused
  if (s.ticket != UINT64_MAX) GST_CUDA_TRY(cudaEventSynchronize(s.done));
  const size_t need = L.off_region + body_total;
  if (need > s.cap_in) {
    if (s.pinned) cudaFreeHost(s.pinned);
    if (s.d_in) cudaFree(s.d_in);
    s.pinned = s.d_in = nullptr;
    s.cap_in = align_up(need + need / 2, 4096);
    GST_CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&s.pinned), s.cap_in, cudaHostAllocDefault));
    GST_CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&s.d_in), s.cap_in));
  }
  if (!out_dev && st->frame_bytes * k > s.cap_out) {
    if (s.d_out) cudaFree(s.d_out);
    s.d_out = nullptr;
    s.cap_out = st->frame_bytes * k;
    GST_CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&s.d_out), s.cap_out));
  }
  // page layout of the host batch loader: [offsets][frame 0 minus its header][frame 1 ...], frequency blocks inline
  uint32_t *out_off = reinterpret_cast<uint32_t *>(s.pinned), *in_off = out_off + 4 * k;
  memset(s.pinned, 0, L.off_region);
  uint32_t in_acc = 0, out_acc = 0;
  std::vector<uint32_t> base(k);
  for (uint32_t i = 0; i < k; ++i) {
    const gst_header &h = hdrs[i];
    const uint32_t in_sz[4] = {h.y_cmp_sz, h.chroma_cmp_sz, h.palette_sz, h.indices_sz};
    const uint32_t out_sz[4] = {2 * L.n_blocks, 4 * L.n_blocks, h.palette_bytes, L.n_blocks};
    base[i] = in_acc;
    in_acc += 2048;
    for (int q = 0; q < 4; ++q) {
      in_off[4 * i + q] = in_acc;
      out_off[4 * i + q] = out_acc;
      in_acc += in_sz[q];
      out_acc += out_sz[q];
    }
  }
  if (direct) {
    GST_CUDA_TRY(cudaMemcpyAsync(s.d_in, s.pinned, L.off_region, cudaMemcpyHostToDevice, s.stream));
    for (uint32_t i = 0; i < k; ++i)
      GST_CUDA_TRY(cudaMemcpyAsync(s.d_in + L.off_region + base[i], frames[i] + GST_HEADER_BYTES, lens[i] - GST_HEADER_BYTES,
                                   cudaMemcpyHostToDevice, s.stream));
  } else {
    for (uint32_t i = 0; i < k; ++i) memcpy(s.pinned + L.off_region + base[i], frames[i] + GST_HEADER_BYTES, lens[i] - GST_HEADER_BYTES);
    GST_CUDA_TRY(cudaMemcpyAsync(s.d_in, s.pinned, need, cudaMemcpyHostToDevice, s.stream));
  }
  uint8_t *dst = out_dev ? out_dev : s.d_out;
  rc = decode_batch(st->ctx, hdrs.data(), k, s.stream, s.d_in, s.cap_in, dst, st->mode, Taps{}, nullptr, 0, nullptr, false, true);
  if (rc) return rc;
  if (out_host) GST_CUDA_TRY(cudaMemcpyAsync(out_host, dst, st->frame_bytes * k, cudaMemcpyDeviceToHost, s.stream));
  GST_CUDA_TRY(cudaEventRecord(s.done, s.stream));
  s.ticket = 0;  // (in use; frames of a group have no tickets of their own)
  s.frame_dev = dst;
  return GST_OK;
}
}  // namespace

// The demo's main loop (demo/demo.cpp:504-600: for every frame load the file, decode it, hand it on) over a sequence
// that is already in host memory.  Frames are taken `group` at a time (GST_PLAY_GROUP in flags, default 8): one
// LoadCompressedDXTs-style call and one read-back per group instead of per frame, `depth` groups in flight -- per
// frame that leaves one host copy (or one upload) and an eighth of everything else, which is what lets one host
// thread keep the PCIe link busy: 600 frames of 1920x1024 with read-back, frames per second on one B200 --
// 30 k one by one, 47 k in groups of 4, 50 k in groups of 16 with staged uploads (the read-back floor is 58 k);
// with GST_SUBMIT_DIRECT 39 k whatever the group (many small uploads slow the read-back beside them down), so
// direct upload pays only when the frames are not read back (89-94 k against 61 k one by one).
int gst_streamer_play(gst_streamer *st, const uint8_t *const *frames, const size_t *lens, uint32_t n, void *out_dev,
                      void *out_host, uint32_t flags) {
  if (!st || !frames || !lens) return fail(GST_ERR_INVALID, "null argument");
  if (flags & ~(GST_SUBMIT_DIRECT | 0xFF00u)) return fail(GST_ERR_INVALID, "unknown play flags 0x%x", flags);
  uint32_t group = (flags >> 8) & 0xFFu;
  if (group == 0) group = 8;
  DeviceGuard guard(st->ctx->device);
  // frames submitted one by one before this call must have drained: their slots are reused
  for (auto &s : st->slots)
    if (s.ticket != UINT64_MAX) GST_CUDA_TRY(cudaEventSynchronize(s.done));
  int rc = GST_OK;
  uint32_t g = 0;
  for (uint32_t f = 0; f < n && rc == GST_OK; f += group, ++g) {
    const uint32_t k = std::min(group, n - f);
    uint8_t *od = out_dev ? static_cast<uint8_t *>(out_dev) + static_cast<size_t>(f) * st->frame_bytes : nullptr;
    uint8_t *oh = out_host ? static_cast<uint8_t *>(out_host) + static_cast<size_t>(f) * st->frame_bytes : nullptr;
    rc = streamer_submit_group(st, g % st->depth, frames + f, lens + f, k, od, oh, (flags & GST_SUBMIT_DIRECT) != 0);
  }
  for (auto &s : st->slots) {
    if (s.ticket != UINT64_MAX) {
      cudaError_t e = cudaEventSynchronize(s.done);
      if (e != cudaSuccess && rc == GST_OK) rc = fail(GST_ERR_CUDA, "frame decode failed: %s", cudaGetErrorString(e));
    }
    s.ticket = UINT64_MAX;
  }
  st->next_ticket += n;
  return rc;
}

int gst_streamer_wait(gst_streamer *st, uint64_t ticket, void **frame_dev) {
  if (!st) return fail(GST_ERR_INVALID, "null streamer");
  gst_streamer::Slot &s = st->slots[ticket % st->depth];
  if (s.ticket != ticket)
    return fail(GST_ERR_INVALID, "frame %llu is not in flight (its slot holds frame %llu)", (unsigned long long)ticket,
                (unsigned long long)s.ticket);
  DeviceGuard guard(st->ctx->device);
  GST_CUDA_TRY(cudaEventSynchronize(s.done));
  if (frame_dev) *frame_dev = s.frame_dev;
  return GST_OK;
}

// ---- standalone rANS decoder -----------------------------------------------------------
int gst_normalize_frequencies(const uint32_t *counts, uint32_t n, uint32_t target_sum, uint32_t *out) {
  if (!counts || !out || n == 0) return fail(GST_ERR_INVALID, "null or empty argument");
  if (target_sum > 0x7FFFFFFFu) return fail(GST_ERR_INVALID, "target sum too large");
  std::vector<uint32_t> h;
  int rc = normalize_frequencies(counts, n, target_sum ? static_cast<int>(target_sum) : gst::kTableSize, &h);
  if (rc) return rc;
  memcpy(out, h.data(), n * sizeof(uint32_t));
  return GST_OK;
}

int gst_build_tables(gst_ctx *ctx, void *stream, const void *freqs_dev, uint32_t n_tables, void *tables_dev) {
  if (!ctx || !freqs_dev || !tables_dev) return fail(GST_ERR_INVALID, "null argument");
  DeviceGuard guard(ctx->device);
  GST_CUDA_TRY(gst::launch_build_tables(static_cast<const uint8_t *>(freqs_dev), n_tables,
                                        static_cast<uint32_t *>(tables_dev), static_cast<cudaStream_t>(stream)));
  return GST_OK;
}

// ---- rANS stream encoder (fixture tooling) --------------------------------------------------
size_t gst_ans_encode_bound(size_t n_symbols) {
  const size_t groups = n_symbols / gst::kGroupSyms;
  return 4 * groups + groups * (gst::kEncGroupCapBytes + 2) + 4;
}

int gst_ans_encode_stream(gst_ctx *ctx, const uint8_t *symbols, size_t n_symbols, uint16_t *freqs_out,
                          uint8_t *stream_out, size_t stream_cap, size_t *stream_bytes) {
  if (!ctx || !symbols || !freqs_out || !stream_out || !stream_bytes) return fail(GST_ERR_INVALID, "null argument");
  if (n_symbols == 0 || n_symbols % gst::kGroupSyms) return fail(GST_ERR_INVALID, "n_symbols must be a positive multiple of 8192");
  if (n_symbols / gst::kGroupSyms > 0x7FFFFFu) return fail(GST_ERR_INVALID, "stream too long for 32-bit offsets");
  // histogram -> normalised frequencies, codec/entropy.cpp:176-190
  std::vector<uint32_t> counts(256, 0);
  for (size_t i = 0; i < n_symbols; ++i) counts[symbols[i]]++;
  uint32_t nz = 0;
  for (uint32_t i = 0; i < 256; ++i) if (counts[i]) nz = i + 1;
  std::vector<uint32_t> F(nz);
  int rc = gst_normalize_frequencies(counts.data(), nz, 0, F.data());
  if (rc != GST_OK) return rc;
  uint16_t f16[256];
  for (uint32_t i = 0; i < 256; ++i) {
    f16[i] = i < nz ? static_cast<uint16_t>(F[i]) : 0;
    if (counts[i] && !f16[i]) return fail(GST_ERR_INVALID, "symbol %u occurs but has frequency 0", i);
  }
  memcpy(freqs_out, f16, sizeof f16);

  const uint32_t groups = static_cast<uint32_t>(n_symbols / gst::kGroupSyms);
  DeviceGuard guard(ctx->device);
  cudaStream_t s = ctx->streams[0];
  // one grow-only device workspace per context, carved up per call (no allocation in steady state):
  // [symbols][per-group scratch][freqs 512][sizes][offsets][output stream (bound)]
  std::lock_guard<std::mutex> ws_lock(ctx->ans_mutex);
  const size_t o_sym = 0;
  const size_t o_scr = o_sym + align_up(n_symbols, 256);
  const size_t o_f = o_scr + align_up(static_cast<size_t>(groups) * gst::kEncGroupCapBytes, 256);
  const size_t o_sizes = o_f + 512;
  const size_t o_offs = o_sizes + align_up(4 * static_cast<size_t>(groups), 256);
  const size_t o_out = o_offs + align_up(4 * static_cast<size_t>(groups), 256);
  const size_t ws_bytes = o_out + align_up(gst_ans_encode_bound(n_symbols), 256);
  int rc2 = ans_workspace(ctx, ws_bytes);
  if (rc2) return rc2;
  uint8_t *d_sym = ctx->ans_ws + o_sym, *d_scratch = ctx->ans_ws + o_scr, *d_out = ctx->ans_ws + o_out;
  uint16_t *d_f = reinterpret_cast<uint16_t *>(ctx->ans_ws + o_f);
  uint32_t *d_sizes = reinterpret_cast<uint32_t *>(ctx->ans_ws + o_sizes), *d_offsets = reinterpret_cast<uint32_t *>(ctx->ans_ws + o_offs);
  std::vector<uint32_t> sizes(groups), offsets(groups);
  size_t total = 0;
  cudaError_t e = cudaMemcpyAsync(d_sym, symbols, n_symbols, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_f, f16, sizeof f16, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = gst::launch_ans_encode(d_sym, groups, d_f, d_scratch, d_sizes, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(sizes.data(), d_sizes, 4 * static_cast<size_t>(groups), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e == cudaSuccess) {
    // offsets are measured from the start of the stream and include the offset table (codec/entropy.cpp:202-231)
    size_t cum = 4 * static_cast<size_t>(groups);
    for (uint32_t g = 0; g < groups; ++g) {
      cum += (sizes[g] + 3u) & ~3u;
      offsets[g] = static_cast<uint32_t>(cum);
    }
    total = (cum + 3) & ~static_cast<size_t>(3);
    if (total > stream_cap) return fail(GST_ERR_INVALID, "stream_out holds %zu bytes, the stream needs %zu", stream_cap, total);
    e = cudaMemsetAsync(d_out, 0, total, s);
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_offsets, offsets.data(), 4 * static_cast<size_t>(groups), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = gst::launch_ans_encode_gather(d_scratch, d_sizes, d_offsets, groups, d_out, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(stream_out, d_out, total, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return fail(GST_ERR_CUDA, "ans encode failed: %s", cudaGetErrorString(e));
  *stream_bytes = total;
  return GST_OK;
}

int gst_ans_rebuild(gst_ans_decoder *d, const uint32_t *F, uint32_t n) {
  if (!d || !F || n == 0 || n > 256) return fail(GST_ERR_INVALID, "need 1..256 symbol counts");
  std::vector<uint32_t> h;
  int rc = normalize_frequencies(F, n, gst::kTableSize, &h);
  if (rc) return rc;
  uint16_t f16[256] = {0};
  for (uint32_t i = 0; i < n; ++i) f16[i] = static_cast<uint16_t>(h[i]);
  DeviceGuard guard(d->ctx->device);
  cudaStream_t s = d->ctx->streams[0];
  // pageable source: the copy is staged before the call returns
  GST_CUDA_TRY(cudaMemcpyAsync(d->freqs, f16, sizeof(f16), cudaMemcpyHostToDevice, s));
  GST_CUDA_TRY(gst::launch_build_tables(d->freqs, 1, d->table, s));
  GST_CUDA_TRY(cudaStreamSynchronize(s));
  return GST_OK;
}

int gst_ans_create(gst_ctx *ctx, const uint32_t *F, uint32_t n, uint32_t lanes, gst_ans_decoder **out) {
  if (!ctx || !out) return fail(GST_ERR_INVALID, "null argument");
  if (lanes == 0 || lanes > gst::kLanes) return fail(GST_ERR_INVALID, "lanes must be 1..32");
  DeviceGuard guard(ctx->device);
  gst_ans_decoder *d = new (std::nothrow) gst_ans_decoder;
  if (!d) return fail(GST_ERR_NOMEM, "out of host memory");
  d->ctx = ctx;
  d->lanes = lanes;
  cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&d->table), gst::kTableSize * 4);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&d->freqs), 512);
  if (e != cudaSuccess) {
    gst_ans_destroy(d);
    return fail(GST_ERR_NOMEM, "cudaMalloc failed: %s", cudaGetErrorString(e));
  }
  int rc = gst_ans_rebuild(d, F, n);
  if (rc) {
    gst_ans_destroy(d);
    return rc;
  }
  *out = d;
  return GST_OK;
}

int gst_ans_table(gst_ans_decoder *d, uint8_t *symbols, uint16_t *freqs, uint16_t *cum_freqs) {
  if (!d) return fail(GST_ERR_INVALID, "null decoder");
  DeviceGuard guard(d->ctx->device);
  std::vector<uint32_t> t(gst::kTableSize);
  GST_CUDA_TRY(cudaMemcpy(t.data(), d->table, t.size() * 4, cudaMemcpyDeviceToHost));
  for (uint32_t slot = 0; slot < gst::kTableSize; ++slot) {
    const uint32_t e = t[slot];
    uint32_t sym, freq, cum;
    gst::unpack_entry(e, slot, &sym, &freq, &cum);
    if (symbols) symbols[slot] = static_cast<uint8_t>(sym);
    if (freqs) freqs[slot] = static_cast<uint16_t>(freq);
    if (cum_freqs) cum_freqs[slot] = static_cast<uint16_t>(cum);
  }
  return GST_OK;
}

int gst_ans_decode(gst_ans_decoder *d, uint32_t lanes, const uint32_t *states, const uint8_t *const *data,
                   const size_t *data_len, uint32_t groups, uint8_t *out) {
  if (!d || !states || !data || !data_len || !out) return fail(GST_ERR_INVALID, "null argument");
  if (lanes == 0 || lanes > d->lanes) return fail(GST_ERR_INVALID, "lanes must be 1..%u", d->lanes);
  if (groups == 0) return GST_OK;
  // the layout ans/ans_ocl.cpp:283-313 builds: [u32 end offsets][pad | words | states]...
  const size_t head = align_up(4 * static_cast<size_t>(groups), 4);
  std::vector<uint8_t> host;
  std::vector<uint32_t> offsets(groups);
  size_t pos = head;
  for (uint32_t g = 0; g < groups; ++g) {
    if (data_len[g] & 1) return fail(GST_ERR_INVALID, "group %u: renorm data must be a whole number of 16-bit words", g);
    pos += align_up(data_len[g], 4) + 4 * static_cast<size_t>(lanes);
    offsets[g] = static_cast<uint32_t>(pos);
  }
  const size_t total = align_up(pos, 16);
  host.assign(total, 0);
  memcpy(host.data(), offsets.data(), 4 * static_cast<size_t>(groups));
  for (uint32_t g = 0; g < groups; ++g) {
    uint8_t *end = host.data() + offsets[g];
    memcpy(end - 4 * lanes, states + static_cast<size_t>(g) * lanes, 4 * static_cast<size_t>(lanes));
    memcpy(end - 4 * lanes - data_len[g], data[g], data_len[g]);
  }
  DeviceGuard guard(d->ctx->device);
  cudaStream_t s = d->ctx->streams[0];
  const size_t out_bytes = static_cast<size_t>(groups) * lanes * gst::kSymsPerLane;
  std::lock_guard<std::mutex> ws_lock(d->ctx->ans_mutex);  // the context's grow-only workspace: [input][output]
  const size_t o_out = align_up(total, 256);
  int rc = ans_workspace(d->ctx, o_out + out_bytes);
  if (rc) return rc;
  uint8_t *dbuf = d->ctx->ans_ws, *dout = d->ctx->ans_ws + o_out;
  cudaError_t e = cudaMemcpyAsync(dbuf, host.data(), total, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = gst::launch_ans_decode_plain(d->table, dbuf, total, groups, lanes, dout, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, dout, out_bytes, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return fail(GST_ERR_CUDA, "ans decode failed: %s", cudaGetErrorString(e));
  return GST_OK;
}

void gst_ans_destroy(gst_ans_decoder *d) {
  if (!d) return;
  DeviceGuard guard(d->ctx->device);
  if (d->table) cudaFree(d->table);
  if (d->freqs) cudaFree(d->freqs);
  delete d;
}

int gst_ctx_set_direct_upload(gst_ctx *ctx, int on) {
  if (!ctx) return fail(GST_ERR_INVALID, "null context");
  ctx->direct_upload.store(on != 0);
  return GST_OK;
}

int gst_status_flags(gst_ctx *ctx, uint32_t *flags, int clear) {
  if (!ctx || !flags) return fail(GST_ERR_INVALID, "null argument");
  DeviceGuard guard(ctx->device);
  GST_CUDA_TRY(cudaMemcpy(flags, ctx->d_status, sizeof(uint32_t), cudaMemcpyDeviceToHost));
  if (clear && *flags) GST_CUDA_TRY(cudaMemset(ctx->d_status, 0, sizeof(uint32_t)));
  return GST_OK;
}

int gst_launches_per_batch(void) { return gst::kLaunchesPerBatch; }

int gst_launches_for_batch(const gst_header *hdrs, uint32_t n) {
  BatchLayout L;
  if (layout_batch(hdrs, n, &L)) return 0;
  return gst::is_small_call(n, L.groups_per_plane, L.max_palette) ? 2 : 3;
}

int gst_profile_enable(gst_ctx *ctx, int on) {
  if (!ctx) return fail(GST_ERR_INVALID, "null context");
  std::lock_guard<std::mutex> lock(ctx->prof_mutex);
  ctx->prof_on = on != 0;
  return GST_OK;
}

int gst_profile_read(gst_ctx *ctx, double *kernel_ms, uint32_t n_kernels, uint64_t *calls) {
  if (!ctx || !kernel_ms || !calls) return fail(GST_ERR_INVALID, "null argument");
  if (n_kernels < static_cast<uint32_t>(gst::kLaunchesPerBatch))
    return fail(GST_ERR_SMALL, "need room for %d kernel times", gst::kLaunchesPerBatch);
  DeviceGuard guard(ctx->device);
  std::vector<cudaEvent_t> evs;
  {
    std::lock_guard<std::mutex> lock(ctx->prof_mutex);
    evs.swap(ctx->prof_events);
  }
  for (uint32_t k = 0; k < n_kernels; ++k) kernel_ms[k] = 0.0;
  const size_t per = gst::kLaunchesPerBatch + 1;
  *calls = evs.size() / per;
  cudaError_t err = cudaSuccess;
  for (size_t c = 0; c < *calls; ++c) {
    cudaError_t e = cudaEventSynchronize(evs[c * per + per - 1]);
    for (size_t k = 0; k + 1 < per && e == cudaSuccess; ++k) {
      float ms = 0.f;
      e = cudaEventElapsedTime(&ms, evs[c * per + k], evs[c * per + k + 1]);
      kernel_ms[k] += ms;
    }
    if (err == cudaSuccess) err = e;
  }
  for (auto ev : evs) cudaEventDestroy(ev);
  if (err != cudaSuccess) return fail(GST_ERR_CUDA, "profile read failed: %s", cudaGetErrorString(err));
  return GST_OK;
}

}  // extern "C"
