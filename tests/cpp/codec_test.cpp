// C++ counterpart of the reference's codec/test/codec_test.cpp:36-48 and of the batched call
// site demo/photos_sf.cpp:753-821, written against include/gst_decoder.hpp.  The encoder is
// not part of this repository, so the compressed stream and the encoder's PhysicalBlocks()
// come from the committed golden fixture (tests/golden/test1.{gst,dxt}).
// usage: codec_test <test1.gst> <test1.dxt>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <vector>

#include "gst_decoder.hpp"

static std::vector<uint8_t> ReadFile(const char *path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) {
    std::fprintf(stderr, "cannot open %s\n", path);
    std::exit(2);
  }
  return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

#define EXPECT(cond)                                                     \
  do {                                                                   \
    if (!(cond)) {                                                       \
      std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      return 1;                                                          \
    }                                                                    \
  } while (0)

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  const std::vector<uint8_t> cmp_data = ReadFile(argv[1]);
  const std::vector<uint8_t> golden = ReadFile(argv[2]);

  std::unique_ptr<gpu::GPUContext> ctx = gpu::GPUContext::InitializeCUDA(0);
  if (!ctx) {
    std::fprintf(stderr, "no device: %s\n", gst_last_error());
    return 3;
  }
  EXPECT(GenTC::InitializeDecoder(ctx));

  // TEST(GenTC, CanCompressAndDecompressImage)
  GenTC::DXTImage cmp_img = GenTC::DecompressDXT(ctx, cmp_data);
  const GenTC::PhysicalDXTBlock *blks = reinterpret_cast<const GenTC::PhysicalDXTBlock *>(golden.data());
  EXPECT(cmp_img.PhysicalBlocks().size() == golden.size() / 8);
  size_t bad = 0;
  for (size_t i = 0; i < cmp_img.PhysicalBlocks().size(); ++i) bad += blks[i].dxt_block != cmp_img.PhysicalBlocks()[i].dxt_block;
  EXPECT(bad == 0);

  // LoadCompressedDXT on a caller-owned buffer, ordered by an event, with a preallocated arena
  GenTC::GenTCHeader hdr;
  gst_mem cmp_buf = GenTC::UploadData(ctx, cmp_data, &hdr);
  EXPECT(hdr.width == 512 && hdr.height == 512);
  GenTC::PreallocateDecompressor(ctx, 4 * GenTC::RequiredScratchMem(hdr));
  gst_mem out = ctx->CreateBuffer(golden.size() * 3);
  gst_queue q = ctx->GetNextQueue();
  gst_event init = nullptr;
  EXPECT(gst_event_record(ctx->Handle(), ctx->GetDefaultCommandQueue(), &init) == GST_OK);
  gst_event done = GenTC::LoadCompressedDXT(ctx, hdr, q, cmp_buf, out, 1, &init);
  EXPECT(done != nullptr);
  std::vector<uint8_t> host(golden.size());
  EXPECT(gst_event_wait(done) == GST_OK);
  ctx->ReadBuffer(q, out, 0, host.data(), host.size(), true);
  EXPECT(host == golden);
  gst_event_destroy(done);
  gst_event_destroy(init);

  // LoadCompressedDXTs: three copies packed photos_sf-style
  std::vector<const uint8_t *> files(3, cmp_data.data());
  std::vector<size_t> lens(3, cmp_data.size());
  std::vector<gst_header> chdrs(3);
  for (int i = 0; i < 3; ++i) EXPECT(gst_parse_header(files[i], lens[i], &chdrs[i]) == GST_OK);
  std::vector<uint8_t> packed(gst_packed_size(chdrs.data(), 3));
  EXPECT(gst_pack_batch(files.data(), lens.data(), 3, packed.data(), packed.size(), chdrs.data()) == GST_OK);
  gst_mem batch = ctx->CreateBuffer(packed.size());
  ctx->WriteBuffer(q, batch, 0, packed.data(), packed.size(), true);
  std::vector<GenTC::GenTCHeader> hdrs(3, hdr);
  done = GenTC::LoadCompressedDXTs(ctx, hdrs, q, batch, out, 0, nullptr);
  EXPECT(gst_event_wait(done) == GST_OK);
  gst_event_destroy(done);
  for (int i = 0; i < 3; ++i) {
    ctx->ReadBuffer(q, out, i * golden.size(), host.data(), host.size(), true);
    EXPECT(host == golden);
  }
  GenTC::FreeDecompressor();  // the reference's argument-less form (codec/decoder.h:37)

  // the photos_sf page loop as one call (host files -> device textures), ragged last page
  {
    std::vector<std::vector<uint8_t> > batch_files(5, cmp_data);
    gst_mem textures = ctx->CreateBuffer(5 * golden.size());
    GenTC::LoadHostBatch(ctx, batch_files, textures, 2);
    for (int i = 0; i < 5; ++i) {
      ctx->ReadBuffer(q, textures, i * golden.size(), host.data(), host.size(), true);
      EXPECT(host == golden);
    }
  }

  // the demo frame loop with 2 frames in flight
  {
    GenTC::FrameStreamer player(ctx, 512, 512, 2);
    uint64_t t0 = player.Submit(cmp_data), t1 = player.Submit(cmp_data);
    for (uint64_t t : {t0, t1}) {
      void *frame = player.Wait(t);
      gst_mem view;
      view.ptr = frame;
      view.bytes = golden.size();
      ctx->ReadBuffer(q, view, 0, host.data(), host.size(), true);
      EXPECT(host == golden);
    }
  }

  // ans::ocl table interface (ans/ans_ocl_test.cpp:64-108)
  std::vector<uint32_t> F = {3, 2, 1, 4, 3};
  ans::ocl::OpenCLDecoder decoder(ctx, F, 1);
  std::vector<uint32_t> nf = ans::ocl::NormalizeFrequencies(F);
  std::vector<uint8_t> syms = decoder.GetSymbols();
  std::vector<uint16_t> freqs = decoder.GetFrequencies(), cums = decoder.GetCumulativeFrequencies();
  size_t sum = 0;
  for (size_t i = 0; i < nf.size(); ++i) {
    for (uint32_t j = 0; j < nf[i]; ++j) {
      EXPECT(syms[sum + j] == i && freqs[sum + j] == nf[i] && cums[sum + j] == sum);
    }
    sum += nf[i];
  }
  EXPECT(sum == ans::ocl::kANSTableSize);

  ctx->ReleaseBuffer(cmp_buf);
  ctx->ReleaseBuffer(batch);
  ctx->ReleaseBuffer(out);
  std::printf("codec_test: OK (%zu blocks bit-exact)\n", golden.size() / 8);
  return 0;
}
