// gst_decoder.hpp -- C++ facade over the C ABI (gst_cuda.h) that keeps the reference's decoder
// API surface, so existing callers of GammaUNC/GST's decode path compile against it with the
// OpenCL handle types swapped for the aliases below:
//
//   reference (OpenCL)                      here
//   -------------------------------------   ---------------------------------------------
//   std::unique_ptr<gpu::GPUContext>        std::unique_ptr<gpu::GPUContext> (wraps gst_ctx*)
//   gpu::GPUContext::InitializeOpenCL(gl)   gpu::GPUContext::InitializeCUDA(device)
//   cl_command_queue                        gst_queue  (cudaStream_t as void*)
//   cl_event                                gst_event  (cudaEvent_t as void*, caller destroys)
//   cl_mem                                  gst_mem    (device pointer + size)
//
// Header only; link with -lgst_cuda.  Citations are relative to the reference tree.
#ifndef GST_DECODER_HPP_
#define GST_DECODER_HPP_

#include <cassert>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "gst_cuda.h"

typedef void *gst_queue;
typedef void *gst_event;
struct gst_mem {
  void *ptr;
  size_t bytes;
};

namespace gpu {

// gpu/gpu.h:50-174, reduced to what the decode path uses.
class GPUContext {
 public:
  static std::unique_ptr<GPUContext> InitializeCUDA(int device = 0) {
    gst_ctx *c = nullptr;
    if (gst_ctx_create(device, &c) != GST_OK) return nullptr;  // InitializeOpenCL also returns null on failure
    return std::unique_ptr<GPUContext>(new GPUContext(c));
  }
  ~GPUContext() { gst_ctx_destroy(_ctx); }
  gst_ctx *Handle() const { return _ctx; }
  gst_queue GetDefaultCommandQueue() const { return gst_stream_default(_ctx); }  // gpu/gpu.h:53
  gst_queue GetNextQueue() const { return gst_stream_next(_ctx); }               // gpu/gpu.h:56-59
  void FlushAllQueues() const { gst_ctx_sync(_ctx); }                            // gpu/gpu.h:61-66

  // clCreateBuffer / clReleaseMemObject / clEnqueue{Write,Read}Buffer equivalents
  gst_mem CreateBuffer(size_t bytes) const {
    gst_mem m = {nullptr, bytes};
    if (gst_malloc(_ctx, bytes, &m.ptr) != GST_OK) throw std::runtime_error(gst_last_error());
    return m;
  }
  void ReleaseBuffer(gst_mem m) const { gst_free(_ctx, m.ptr); }
  void WriteBuffer(gst_queue q, gst_mem dst, size_t offset, const void *src, size_t bytes, bool blocking) const {
    if (gst_upload_async(_ctx, q, static_cast<uint8_t *>(dst.ptr) + offset, src, bytes) != GST_OK)
      throw std::runtime_error(gst_last_error());
    if (blocking) gst_stream_sync(_ctx, q);
  }
  void ReadBuffer(gst_queue q, gst_mem src, size_t offset, void *dst, size_t bytes, bool blocking) const {
    if (gst_download_async(_ctx, q, dst, static_cast<const uint8_t *>(src.ptr) + offset, bytes) != GST_OK)
      throw std::runtime_error(gst_last_error());
    if (blocking) gst_stream_sync(_ctx, q);
  }

 private:
  explicit GPUContext(gst_ctx *c) : _ctx(c) {}
  GPUContext(const GPUContext &);
  gst_ctx *_ctx;
};

}  // namespace gpu

namespace GenTC {

// codec/codec_base.h:9-22
struct GenTCHeader {
  uint32_t width;
  uint32_t height;
  uint32_t palette_bytes;
  uint32_t y_cmp_sz;
  uint32_t chroma_cmp_sz;
  uint32_t palette_sz;
  uint32_t indices_sz;

  void Print() const {
    std::printf("Width: %u\nHeight: %u\nNum Palette Entries: %u\nY compressed size: %u\n"
                "Chroma compressed size: %u\nPalette size compressed: %u\nPalette index deltas compressed: %u\n",
                width, height, palette_bytes / 4, y_cmp_sz, chroma_cmp_sz, palette_sz, indices_sz);
  }
  void LoadFrom(const uint8_t *buf) { std::memcpy(this, buf, sizeof(*this)); }  // codec/codec_base.cpp:18-24
};
static_assert(sizeof(GenTCHeader) == sizeof(gst_header), "GenTCHeader and gst_header must be layout compatible");
static const size_t kWaveletBlockDim = 32;

// codec/dxt_image.h:14-21
union PhysicalDXTBlock {
  struct {
    uint16_t ep1;
    uint16_t ep2;
    uint32_t interpolation;
  };
  uint64_t dxt_block;
};

// The decoded-image holder DecompressDXT returns: the subset of GenTC::DXTImage
// (codec/dxt_image.h:39-84) that does not need the encoder.
class DXTImage {
 public:
  DXTImage(int width, int height, const std::vector<uint8_t> &dxt_data)
      : _width(width), _height(height), _blocks(dxt_data.size() / 8) {
    std::memcpy(_blocks.data(), dxt_data.data(), _blocks.size() * 8);
  }
  int Width() const { return _width; }
  int Height() const { return _height; }
  int BlocksWide() const { return _width / 4; }
  int BlocksHigh() const { return _height / 4; }
  const std::vector<PhysicalDXTBlock> &PhysicalBlocks() const { return _blocks; }

 private:
  int _width, _height;
  std::vector<PhysicalDXTBlock> _blocks;
};

namespace detail {
inline const gst_header *AsC(const GenTCHeader *h) { return reinterpret_cast<const gst_header *>(h); }
inline gst_event Load(const std::unique_ptr<gpu::GPUContext> &ctx, const GenTCHeader *hdrs, size_t n, gst_queue queue,
                      gst_mem cmp_data, gst_mem output, uint32_t num_init, const gst_event *init, bool rgb) {
  gst_event done = nullptr;
  int rc = (rgb ? gst_load_rgb_batch : gst_load_dxt_batch)(ctx->Handle(), AsC(hdrs), static_cast<uint32_t>(n), queue,
                                                           cmp_data.ptr, cmp_data.bytes, output.ptr, init, num_init, &done);
  // The reference asserts in debug builds and ignores errors in release builds
  // (gpu/cl_guards.h:80-96); here a failure is at least visible.
  if (rc != GST_OK) {
    std::fprintf(stderr, "GenTC decode failed: %s\n", gst_last_error());
    assert(!"GenTC decode failed");
  }
  return done;
}
}  // namespace detail

// codec/decoder.h:15-36 -------------------------------------------------------------------
// The reference compiles its seven kernels here and checks the device's work-group limits
// (codec/decoder.cpp:558-591).  The CUDA kernels are compiled ahead of time for sm_100a and gst_ctx_create has
// already refused any other device, so a context that exists is a context that can decode.
inline bool InitializeDecoder(const std::unique_ptr<gpu::GPUContext> &gpu_ctx) { return gpu_ctx != nullptr; }

inline DXTImage DecompressDXT(const std::unique_ptr<gpu::GPUContext> &gpu_ctx, const std::vector<uint8_t> &cmp_data) {
  // validate before anything is sized from the header (a short or corrupt buffer must not be read past its end
  // or drive a huge allocation)
  gst_header h;
  if (gst_parse_header(cmp_data.data(), cmp_data.size(), &h) != GST_OK) throw std::runtime_error(gst_last_error());
  GenTCHeader hdr;
  std::memcpy(&hdr, &h, sizeof(hdr));
  std::vector<uint8_t> out(static_cast<size_t>(hdr.width) * hdr.height / 2, 0xFF);
  if (gst_decompress_host(gpu_ctx->Handle(), cmp_data.data(), cmp_data.size(), 0, out.data(), out.size()) != GST_OK)
    throw std::runtime_error(gst_last_error());
  return DXTImage(static_cast<int>(hdr.width), static_cast<int>(hdr.height), out);
}

inline gst_event LoadCompressedDXT(const std::unique_ptr<gpu::GPUContext> &gpu_ctx, const GenTCHeader &hdr,
                                   gst_queue queue, gst_mem cmp_data, gst_mem output, uint32_t num_init,
                                   const gst_event *init) {
  return detail::Load(gpu_ctx, &hdr, 1, queue, cmp_data, output, num_init, init, false);
}
inline gst_event LoadCompressedDXTs(const std::unique_ptr<gpu::GPUContext> &gpu_ctx, const std::vector<GenTCHeader> &hdr,
                                    gst_queue queue, gst_mem cmp_data, gst_mem output, uint32_t num_init,
                                    const gst_event *init) {
  return detail::Load(gpu_ctx, hdr.data(), hdr.size(), queue, cmp_data, output, num_init, init, false);
}
inline gst_event LoadRGB(const std::unique_ptr<gpu::GPUContext> &gpu_ctx, const GenTCHeader &hdr, gst_queue queue,
                         gst_mem cmp_data, gst_mem output, uint32_t num_init, const gst_event *init) {
  return detail::Load(gpu_ctx, &hdr, 1, queue, cmp_data, output, num_init, init, true);
}
inline gst_event LoadRGBs(const std::unique_ptr<gpu::GPUContext> &gpu_ctx, const std::vector<GenTCHeader> &hdr,
                          gst_queue queue, gst_mem cmp_data, gst_mem output, uint32_t num_init, const gst_event *init) {
  return detail::Load(gpu_ctx, hdr.data(), hdr.size(), queue, cmp_data, output, num_init, init, true);
}

inline size_t RequiredScratchMem(const GenTCHeader &hdr) { return gst_required_scratch(detail::AsC(&hdr)); }

// The reference keeps one process-wide arena (codec/decoder.cpp:96 gPreloader).  Here the arena lives in the
// context; the facade remembers the context PreallocateDecompressor was last called on, so that the reference's
// argument-less FreeDecompressor() (codec/decoder.h:37) works as written.  The context must still be alive.
namespace detail {
inline gst_ctx *&PreloadedContext() {
  static gst_ctx *ctx = nullptr;
  return ctx;
}
}  // namespace detail
inline void PreallocateDecompressor(const std::unique_ptr<gpu::GPUContext> &gpu_ctx, size_t req_sz) {
  if (gst_preallocate(gpu_ctx->Handle(), req_sz) != GST_OK) throw std::runtime_error(gst_last_error());
  detail::PreloadedContext() = gpu_ctx->Handle();
}
inline void FreeDecompressor() {
  if (detail::PreloadedContext()) gst_free_scratch(detail::PreloadedContext());
  detail::PreloadedContext() = nullptr;
}
inline void FreeDecompressor(const std::unique_ptr<gpu::GPUContext> &gpu_ctx) {
  gst_free_scratch(gpu_ctx->Handle());
  if (detail::PreloadedContext() == gpu_ctx->Handle()) detail::PreloadedContext() = nullptr;
}

// UploadData (codec/decoder.cpp:430-476): one .gst file -> the device buffer LoadCompressedDXT
// expects (8 offsets at byte 0, file minus header at byte 512).
inline gst_mem UploadData(const std::unique_ptr<gpu::GPUContext> &gpu_ctx, const std::vector<uint8_t> &cmp_data,
                          GenTCHeader *hdr) {
  const uint8_t *files[1] = {cmp_data.data()};
  const size_t lens[1] = {cmp_data.size()};
  gst_header h;
  if (gst_parse_header(files[0], lens[0], &h) != GST_OK) throw std::runtime_error(gst_last_error());
  std::vector<uint8_t> packed(gst_packed_size(&h, 1));
  if (gst_pack_batch(files, lens, 1, packed.data(), packed.size(), &h) != GST_OK)
    throw std::runtime_error(gst_last_error());
  std::memcpy(hdr, &h, sizeof(h));
  gst_mem m = gpu_ctx->CreateBuffer(packed.size());
  gpu_ctx->WriteBuffer(gpu_ctx->GetDefaultCommandQueue(), m, 0, packed.data(), packed.size(), true);
  return m;
}

// The page loop of demo/photos_sf.cpp:688-885 as one call: host .gst files -> textures in a
// caller-owned device buffer (image i at i * W*H/2), pages of `page` images over the work queues.
inline void LoadHostBatch(const std::unique_ptr<gpu::GPUContext> &gpu_ctx, const std::vector<std::vector<uint8_t> > &files,
                          gst_mem output, uint32_t page = 16) {
  std::vector<const uint8_t *> ptrs;
  std::vector<size_t> lens;
  for (const auto &f : files) {
    ptrs.push_back(f.data());
    lens.push_back(f.size());
  }
  if (gst_load_host_batch(gpu_ctx->Handle(), ptrs.data(), lens.data(), static_cast<uint32_t>(files.size()), page, 0,
                          output.ptr, output.bytes) != GST_OK)
    throw std::runtime_error(gst_last_error());
}

// The frame loop of demo/demo.cpp:145-243 (read frameNNNN.gtc, upload, LoadCompressedDXT or LoadRGB,
// wait) with `depth` frames in flight instead of one.
class FrameStreamer {
 public:
  FrameStreamer(const std::unique_ptr<gpu::GPUContext> &gpu_ctx, uint32_t width, uint32_t height, uint32_t depth = 4,
                bool rgb = false) {
    if (gst_streamer_create(gpu_ctx->Handle(), width, height, depth, rgb ? 1 : 0, &_st) != GST_OK)
      throw std::runtime_error(gst_last_error());
  }
  ~FrameStreamer() { gst_streamer_destroy(_st); }
  FrameStreamer(const FrameStreamer &) = delete;
  FrameStreamer &operator=(const FrameStreamer &) = delete;
  // returns the ticket of the frame; blocks only when `depth` frames are already in flight
  uint64_t Submit(const std::vector<uint8_t> &frame, void *out_dev = nullptr) {
    uint64_t t = 0;
    if (gst_streamer_submit(_st, frame.data(), frame.size(), out_dev, &t) != GST_OK) throw std::runtime_error(gst_last_error());
    return t;
  }
  // blocks until the frame is decoded; device address of its DXT1 blocks / RGB texels
  void *Wait(uint64_t ticket) {
    void *p = nullptr;
    if (gst_streamer_wait(_st, ticket, &p) != GST_OK) throw std::runtime_error(gst_last_error());
    return p;
  }

 private:
  gst_streamer *_st = nullptr;
};

}  // namespace GenTC

namespace ans {
namespace ocl {

// ans/ans.h:72-79
static const size_t kANSTableSize = (1 << 11);
static const size_t kNumEncodedSymbols = 256;
static const size_t kThreadsPerEncodingGroup = 32;

inline std::vector<uint32_t> NormalizeFrequencies(const std::vector<uint32_t> &F) {
  std::vector<uint32_t> out(F.size());
  if (gst_normalize_frequencies(F.data(), static_cast<uint32_t>(F.size()), 0, out.data()) != GST_OK) out.clear();
  return out;
}

// ans::ocl::OpenCLDecoder (ans/ans_ocl.h:26-72) on CUDA.
class CUDADecoder {
 public:
  CUDADecoder(const std::unique_ptr<gpu::GPUContext> &ctx, const std::vector<uint32_t> &F, const int num_interleaved)
      : _num_interleaved(num_interleaved), _d(nullptr) {
    if (gst_ans_create(ctx->Handle(), F.data(), static_cast<uint32_t>(F.size()), static_cast<uint32_t>(num_interleaved),
                       &_d) != GST_OK)
      throw std::runtime_error(gst_last_error());
  }
  ~CUDADecoder() { gst_ans_destroy(_d); }

  std::vector<uint8_t> Decode(uint32_t state, const std::vector<uint8_t> &data) const {
    return Flat(std::vector<uint32_t>(1, state), std::vector<std::vector<uint8_t> >(1, data), 1);
  }
  std::vector<std::vector<uint8_t> > Decode(const std::vector<uint32_t> &states, const std::vector<uint8_t> &data) const {
    return Split(Flat(states, std::vector<std::vector<uint8_t> >(1, data), static_cast<uint32_t>(states.size())));
  }
  std::vector<std::vector<uint8_t> > Decode(const std::vector<uint32_t> &states,
                                            const std::vector<std::vector<uint8_t> > &data) const {
    return Split(Flat(states, data, static_cast<uint32_t>(_num_interleaved)));
  }
  void RebuildTable(const std::vector<uint32_t> &F) {
    if (gst_ans_rebuild(_d, F.data(), static_cast<uint32_t>(F.size())) != GST_OK)
      throw std::runtime_error(gst_last_error());
  }
  std::vector<uint8_t> GetSymbols() const {
    std::vector<uint8_t> s(kANSTableSize);
    gst_ans_table(_d, s.data(), nullptr, nullptr);
    return s;
  }
  std::vector<uint16_t> GetFrequencies() const {
    std::vector<uint16_t> f(kANSTableSize);
    gst_ans_table(_d, nullptr, f.data(), nullptr);
    return f;
  }
  std::vector<uint16_t> GetCumulativeFrequencies() const {
    std::vector<uint16_t> c(kANSTableSize);
    gst_ans_table(_d, nullptr, nullptr, c.data());
    return c;
  }

 private:
  CUDADecoder(const CUDADecoder &);
  std::vector<uint8_t> Flat(const std::vector<uint32_t> &states, const std::vector<std::vector<uint8_t> > &data,
                            uint32_t lanes) const {
    std::vector<const uint8_t *> ptrs;
    std::vector<size_t> lens;
    for (const auto &d : data) {
      ptrs.push_back(d.data());
      lens.push_back(d.size());
    }
    std::vector<uint8_t> out(states.size() * kNumEncodedSymbols);
    if (gst_ans_decode(_d, lanes, states.data(), ptrs.data(), lens.data(), static_cast<uint32_t>(data.size()),
                       out.data()) != GST_OK)
      throw std::runtime_error(gst_last_error());
    return out;
  }
  static std::vector<std::vector<uint8_t> > Split(const std::vector<uint8_t> &flat) {
    std::vector<std::vector<uint8_t> > out;
    for (size_t i = 0; i < flat.size(); i += kNumEncodedSymbols)
      out.push_back(std::vector<uint8_t>(flat.begin() + i, flat.begin() + i + kNumEncodedSymbols));
    return out;
  }
  const int _num_interleaved;
  gst_ans_decoder *_d;
};
typedef CUDADecoder OpenCLDecoder;  // drop-in name for code written against ans/ans_ocl.h

// ByteEncoder::EncodeBytes::Run (codec/entropy.cpp:174-265) on the GPU: the 512-byte frequency block followed
// by [u32 end offsets][rANS groups], byte for byte what the reference's CPU encoder returns for the same
// symbols (a multiple of 32 * 256 of them).  Fixture tooling; the decode path does not use it.
inline std::vector<uint8_t> EncodeBytes(const std::unique_ptr<gpu::GPUContext> &gpu_ctx, const std::vector<uint8_t> &symbols) {
  std::vector<uint8_t> out(512 + gst_ans_encode_bound(symbols.size()));
  size_t n = 0;
  if (gst_ans_encode_stream(gpu_ctx->Handle(), symbols.data(), symbols.size(), reinterpret_cast<uint16_t *>(out.data()),
                            out.data() + 512, out.size() - 512, &n) != GST_OK)
    throw std::runtime_error(gst_last_error());
  out.resize(512 + n);
  return out;
}

}  // namespace ocl
}  // namespace ans

#endif  // GST_DECODER_HPP_
