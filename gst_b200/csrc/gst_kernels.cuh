// gst_b200 -- device-side declarations shared by the kernels and the C-ABI host layer.
//
// The decode path of GammaUNC/GST (.gst stream -> DXT1 blocks), rewritten for sm_100a.
// Reference stages (citations relative to the reference tree):
//   1. ans/build_table.cl:12-83           -> build_tables_kernel
//   2. ans/ans_decode.cl:25-143           -> rans_decode_group() inside every decode kernel
//   3. codec/decode_indices.cl:6-84       -> rans_streams_kernel (scan fused behind the rANS warp) + the carry in wavelet_assemble_kernel
//   4. codec/inverse_wavelet.cl:69-192    -> wavelet_assemble_kernel
//   5. codec/assemble.cl:64-129           -> wavelet_assemble_kernel
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gst {

constexpr int kTableLog = 11;                 // ans/ans.h:73 kANSTableSize = 1 << 11
constexpr int kTableSize = 1 << kTableLog;
constexpr int kSymsPerLane = 256;             // ans/ans.h:74 kNumEncodedSymbols
constexpr int kLanes = 32;                    // ans/ans.h:75 kThreadsPerEncodingGroup
constexpr int kGroupSyms = kSymsPerLane * kLanes;
constexpr uint32_t kRansL = 16u << kTableLog; // ans/ans_decode.cl:7-8  k*M = 2^15
constexpr int kTile = 32;                     // codec/codec_base.h:22 kWaveletBlockDim
constexpr int kTileSyms = kTile * kTile;

// Packed decode-table entry (internal scratch format, 4 B instead of the reference's 6 B
// AnsTableEntry, codec/decoder.cpp:20-24):
//     bits 0..11   freq                 (1..2048)
//     bits 12..19  symbol
//     bits 20..31  slot - cum_freq      (0..freq-1)
// so that the reference update  state' = (state >> 11) * freq + slot - cum_freq  (ans/ans_decode.cl:38-41)
// is SHF + LOP3 + SHF + IMAD.  (Round 1 used  umulhi(state, freq << 21) + bias'  -- one instruction less, all on
// the FMA pipe -- but on sm_100 IMAD.HI costs far more than its four FMA-pipe cycles: the four-instruction
// form is 2.3 % faster, profiles/README.md.)
__host__ __device__ inline uint32_t pack_entry(uint32_t sym, uint32_t freq, uint32_t slot, uint32_t cum) {
  return (freq & 0xFFFu) | ((sym & 0xFFu) << 12) | (((slot - cum) & 0xFFFu) << 20);
}
// {symbol, freq, cum_freq} of slot `slot` back from a packed entry (table read-back API)
__host__ __device__ inline void unpack_entry(uint32_t e, uint32_t slot, uint32_t *sym, uint32_t *freq, uint32_t *cum) {
  *sym = (e >> 12) & 0xFFu;
  *freq = e & 0xFFFu;
  *cum = slot - (e >> 20);
}

// Geometry + buffer description of one LoadCompressedDXTs-style call.  The compressed
// buffer keeps the reference's device layout (codec/decoder.cpp:153-209, 430-476;
// demo/photos_sf.cpp:753-795):
//   [u32 out_off[4B]][u32 in_off[4B]] pad to 512 | B x 4 x 512 B freqs | payloads
struct BatchParams {
  const uint8_t *cmp;        // device, start of the offsets region
  uint64_t cmp_bytes;        // bytes readable from cmp
  uint32_t n_images;         // B
  uint32_t blocks_x, blocks_y;
  uint32_t n_blocks;         // N = blocks_x * blocks_y
  uint32_t off_region;       // bytes of the (padded) offsets region
  uint32_t groups_per_plane; // N / 8192
  // scratch
  uint32_t *tables;          // [B][4][2048] packed entries
  uint8_t *sym_t;            // [B][6 * N/8192 plane-groups][16][32][16 B]: plane symbols, transposed (gst_kernels.cu)
  uint8_t *palette;          // compact: image b at pal_off(b) = out_off[4b+2] - 7N*b - 6N
  uint64_t palette_cap;      // bytes available in `palette`
  void *idx_s;               // [B][N] u16 or u32, transposed like sym_t: sum of the index deltas after block i in its 256-block run
  uint32_t idx16;            // 1: idx_s holds u16 (every palette <= 65536 entries), 0: u32
  int32_t *run_end;          // [B][N/256] group-local inclusive index prefix at the end of every run
  int32_t *idx_total;        // [B][idx_total_stride(N/8192)] sum of the index deltas of each index group
  // outputs
  uint8_t *out;              // DXT1: B * 8N bytes;  RGB8: B * 48N bytes
  uint32_t *status;          // [0]: flags OR-ed in by the kernels (GST_FLAG_*), [1]: a zero word (the palette of an image
                             // whose palette region is unusable)
  // Image-granular hand-over from rans_streams_kernel to wavelet_assemble_kernel (gst_kernels.cu, "hand-over"):
  uint32_t *img_done;        // [B] rans_streams CTAs of image b that have finished (complete at rans_ctas); zero when the
                             // call starts; NULL: no hand-over (single image), the tile kernel waits for the whole grid
  uint32_t rans_ctas;        // CTAs rans_streams_kernel runs per image (set by launch_decode_batch)
  uint32_t freq_inline;      // 1: no frequency region; an image's four frequency blocks precede its Y stream in the payload
  uint32_t inline_off;       // 1: n_images == 1 and the offset table is off8 below, not the first 32 bytes of cmp
  uint32_t off8[8];          // out_off[0..3], in_off[0..3] of that image
  uint32_t kc[8];            // packed constants the wavelet kernel wants in the constant bank (fill_kernel_constants)
  // optional taps for the stage parity tests (NULL in production)
  uint8_t *tap_symbols;      // reference decmp_buf layout: image b stream s at out_off[4b+s]
  int8_t *tap_planes;        // [B][6][N] raster planes (codec/decoder.cpp:280)
  int32_t *tap_indices;      // [B][N] final palette indices (codec/decoder.cpp:302)
};

// words per image in BatchParams::idx_total: whole 128-byte lines, so that no cache line holds totals of two images
__host__ __device__ inline uint32_t idx_total_stride(uint32_t groups_per_plane) { return (groups_per_plane + 31u) & ~31u; }

inline void fill_kernel_constants(BatchParams *p) {
  auto ph = [](int v) { return (static_cast<uint32_t>(v) & 0xFFFFu) * 65537u; };
  p->kc[0] = ph(3); p->kc[1] = ph(1); p->kc[2] = 1u << 31; p->kc[3] = ph(2); p->kc[4] = ph(5);
  p->kc[5] = ph(-254); p->kc[6] = ph(-251); p->kc[7] = 0x10101010u;
}

// launch helpers (gst_kernels.cu)
cudaError_t launch_build_tables(const uint8_t *freqs, uint32_t n_tables, uint32_t *tables, cudaStream_t s);
// max_palette_bytes = max over the batch of GenTCHeader::palette_bytes (sizes the grid)
// marks: NULL, or kLaunchesPerBatch + 1 events recorded around every kernel (profiling)
cudaError_t launch_decode_batch(const BatchParams &p, int rgb_mode, uint32_t max_palette_bytes,
                                cudaStream_t s, cudaEvent_t *marks = nullptr);
cudaError_t launch_ans_decode_plain(const uint32_t *table, const uint8_t *data,
                                    uint64_t data_bytes, uint32_t n_groups, uint32_t n_lanes,
                                    uint8_t *out, cudaStream_t s);
// rANS encode of a symbol stream (fixture tooling).  scratch: n_groups * kEncGroupCapBytes,
// sizes: n_groups u32; the gather writes [u32 end offsets][groups] given offsets[g] = end of group g.
constexpr size_t kEncGroupCapBytes = 2 * kGroupSyms + 4 * kLanes;
cudaError_t launch_ans_encode(const uint8_t *symbols, uint32_t n_groups, const uint16_t *freqs, uint8_t *scratch,
                              uint32_t *sizes, cudaStream_t s);
cudaError_t launch_ans_encode_gather(const uint8_t *scratch, const uint32_t *sizes, const uint32_t *offsets,
                                     uint32_t n_groups, uint8_t *out, cudaStream_t s);
// kernels launch_decode_batch enqueues: 3 (tables, rANS, wavelet + assembly), or 2 for a "small" call -- one with
// too few rANS groups for the separate table kernel to pay (is_small_call): tables built by the consuming CTAs
constexpr int kLaunchesPerBatch = 3;
// (measured on 2048 x 2048 images, 228 groups each: 32 images 5.5 % faster with the fused build, 64 images 0.8 %, 128
// images 0.3 % slower)
constexpr uint32_t kSmallCallGroups = 16384;
bool is_small_call(uint32_t n_images, uint32_t groups_per_plane, uint32_t max_palette_bytes);

}  // namespace gst
