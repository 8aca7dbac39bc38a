/* TEST INFRASTRUCTURE ONLY -- never linked into, loaded by, or called from the
 * product path (gst_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may use this library.
 *
 * gst_oracle: a plain-C CPU restatement of the GammaUNC/GST decode path
 * (.gst stream -> DXT1 blocks / RGB8), stage by stage, following the
 * reference's OpenCL kernels so that every CUDA stage can be checked in
 * isolation.  Citations are relative to /root/reference.
 *
 * PARITY IS PINNED (not "parity unpinned"): tests/test_oracle.py
 * checks this file against
 *   - the reference's own encoder output PhysicalBlocks() on codec/test/test1.png
 *     (the identity of codec/test/codec_test.cpp:36-48), via tests/golden/,
 *   - the unmodified reference CPU code (ans::DecodeInterleaved,
 *     InverseWavelet2D, GenerateHistogram) linked into oracle/_ref/libgst_ref.so,
 *   - the known-answer vectors of codec/test/wavelet_test.cpp:129-158 and
 *     ans/ans_ocl_test.cpp:64-155.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define GSTO_TABLE_SIZE_LOG 11
#define GSTO_TABLE_SIZE (1u << GSTO_TABLE_SIZE_LOG) /* ans/ans.h:73 kANSTableSize */
#define GSTO_SYMS_PER_LANE 256u                     /* ans/ans.h:74 kNumEncodedSymbols */
#define GSTO_LANES 32u                              /* ans/ans.h:75 kThreadsPerEncodingGroup */
#define GSTO_DECODER_L (16u * GSTO_TABLE_SIZE)      /* ans/ans_decode.cl:7-8: k*M = 2^15 */
#define GSTO_TILE 32                                /* codec/codec_base.h:22 kWaveletBlockDim */

typedef struct {
  uint32_t width, height, palette_bytes, y_cmp_sz, chroma_cmp_sz, palette_sz, indices_sz;
} gsto_header; /* codec/codec_base.h:9-20, raw little-endian memcpy (codec_base.cpp:18-24) */

/* ------------------------------------------------------------------------- */
/* Stage 1: ans/build_table.cl:12-83.  256 frequencies (sum 2048) -> for every
 * slot id in [0,2048) the symbol x with cum[x] <= id < cum[x+1], its frequency
 * and its (exclusive) cumulative frequency.  The kernel does an exclusive scan
 * in local memory then an 11-step branch-free binary search per slot; the
 * search below is that same loop (:61-75), so zero-frequency symbols resolve
 * exactly as on the device. */
void gsto_build_table(const uint16_t *freqs /*256*/, uint16_t *t_freq, uint16_t *t_cum,
                      uint8_t *t_sym /* each 2048 */) {
  uint16_t cum[256];
  uint32_t acc = 0;
  for (int i = 0; i < 256; ++i) { /* exclusive scan, ushort wrap as in :14,:26-54 */
    cum[i] = (uint16_t)acc;
    acc += freqs[i];
  }
  for (uint32_t id = 0; id < GSTO_TABLE_SIZE; ++id) {
    uint32_t low = 0, high = 255, x = (high + low) / 2;
    for (int i = 0; i < GSTO_TABLE_SIZE_LOG; ++i) {
      uint32_t too_high = (uint32_t)(id < cum[x]);
      uint32_t too_low = (uint32_t)(x < 255 && cum[x + 1] <= id);
      uint32_t lo1 = (low + 1 > x) ? low + 1 : x;
      uint32_t hi1 = (high - 1 < x) ? high - 1 : x;
      low = too_low * lo1 + (1 - too_low) * low;
      high = too_high * hi1 + (1 - too_high) * high;
      x = (high + low) / 2;
    }
    t_freq[id] = freqs[x];
    t_cum[id] = cum[x];
    t_sym[id] = (uint8_t)x;
  }
}

/* ------------------------------------------------------------------------- */
/* Stage 2: ans/ans_decode.cl:25-74 (ans_decode_single).  `data` points at the
 * start of one stream: [u32 end_offset[groups]][group 0][group 1]...  `lanes`
 * interleaved rANS states run in lock step and share one 16-bit word stream
 * that is consumed backwards, higher lanes first.  Writes lanes*256 symbols at
 * out[(lane + group*lanes)*256 + 255 - i]. */
void gsto_ans_decode_group(const uint16_t *t_freq, const uint16_t *t_cum, const uint8_t *t_sym,
                           const uint8_t *data, uint32_t group, uint32_t lanes, uint8_t *out) {
  uint32_t offset, state[32];
  memcpy(&offset, data + 4 * (size_t)group, 4);                /* :30 */
  memcpy(state, data + offset - 4 * (size_t)lanes, 4 * lanes); /* :31 */
  uint32_t next_to_read = (offset - lanes * 4) / 2;            /* :32, in u16 units */
  for (uint32_t i = 0; i < GSTO_SYMS_PER_LANE; ++i) {
    uint32_t mask = 0;
    uint8_t sym[32];
    for (uint32_t l = 0; l < lanes; ++l) {
      const uint32_t slot = state[l] & (GSTO_TABLE_SIZE - 1);  /* :38-41 */
      state[l] = (state[l] >> GSTO_TABLE_SIZE_LOG) * t_freq[slot] - t_cum[slot] + slot;
      sym[l] = t_sym[slot];
      if (state[l] < GSTO_DECODER_L) mask |= 1u << l;          /* :44-46 */
    }
    const uint32_t total = (uint32_t)__builtin_popcount(mask); /* :51 */
    for (uint32_t l = 0; l < lanes; ++l) {
      if (mask & (1u << l)) {                                  /* :52-57 */
        const uint32_t below = (uint32_t)__builtin_popcount(mask & ((1u << l) - 1));
        const uint32_t skip = total - below - 1;
        uint16_t w;
        memcpy(&w, data + 2 * (size_t)(next_to_read - skip - 1), 2);
        state[l] = (state[l] << 16) | w;
      }
    }
    next_to_read -= total;                                     /* :65 */
    for (uint32_t l = 0; l < lanes; ++l)                       /* :71-72 */
      out[((size_t)l + (size_t)group * lanes) * GSTO_SYMS_PER_LANE + (GSTO_SYMS_PER_LANE - 1 - i)] = sym[l];
  }
}

/* One whole stream as written by ByteEncoder::EncodeBytes (codec/entropy.cpp:199-262). */
void gsto_ans_decode_stream(const uint8_t *freqs512, const uint8_t *stream, size_t num_symbols,
                            uint8_t *out) {
  uint16_t f[256];
  memcpy(f, freqs512, 512);
  uint16_t *tf = (uint16_t *)malloc(2 * GSTO_TABLE_SIZE);
  uint16_t *tc = (uint16_t *)malloc(2 * GSTO_TABLE_SIZE);
  uint8_t *ts = (uint8_t *)malloc(GSTO_TABLE_SIZE);
  gsto_build_table(f, tf, tc, ts);
  const size_t groups = num_symbols / (GSTO_LANES * GSTO_SYMS_PER_LANE);
  for (size_t g = 0; g < groups; ++g)
    gsto_ans_decode_group(tf, tc, ts, stream, (uint32_t)g, GSTO_LANES, out);
  free(tf);
  free(tc);
  free(ts);
}

/* ------------------------------------------------------------------------- */
/* Stage 3: codec/decode_indices.cl:6-84 + host loop codec/decoder.cpp:311-393.
 * Net effect of the multi-pass 128-wide scans: idx[i] = sum_{j<=i}(byte_j-128)
 * in wrapping 32-bit arithmetic. */
void gsto_decode_indices(const uint8_t *deltas, size_t n, int32_t *out) {
  uint32_t acc = 0;
  for (size_t i = 0; i < n; ++i) {
    acc += (uint32_t)((int32_t)deltas[i] - 128); /* :24 */
    out[i] = (int32_t)acc;
  }
}

/* ------------------------------------------------------------------------- */
/* Stage 4: codec/inverse_wavelet.cl.  NormalizeIndex (:14-16): reflect without
 * repeating the edge sample. */
static int gsto_mirror(int idx, int range) {
  int x = idx - (int)(idx >= range) * (idx - range + 2);
  return x < 0 ? -x : x;
}

/* 1-D inverse 5/3 lifting of src[0..len) = [low half | high half] into
 * dst[0..len), with stride (elements) so the same code does rows and columns.
 * Even samples first (:28-44), then odd samples from the finished evens
 * (:46-64).  `/` is C truncating division, as in OpenCL C. */
static void gsto_inv_lift(const int32_t *src, int sstride, int32_t *dst, int dstride, int len) {
  const int mid = len >> 1;
  for (int i = 0; i < len; i += 2) {
    const int prev = mid + gsto_mirror(i - 1, len) / 2;
    const int next = mid + gsto_mirror(i + 1, len) / 2;
    dst[i * dstride] = src[(i / 2) * sstride] - (src[prev * sstride] + src[next * sstride] + 2) / 4;
  }
  for (int i = 1; i < len; i += 2) {
    const int prev = gsto_mirror(i - 1, len);
    const int next = gsto_mirror(i + 1, len);
    dst[i * dstride] = src[(mid + i / 2) * sstride] + (dst[prev * dstride] + dst[next * dstride]) / 2;
  }
}

/* One 32x32 tile: bytes (row-major inside the tile, :94-100) minus 128, then for
 * len = 2,4,8,16,32 a horizontal pass over rows [0,len) followed by a vertical
 * pass over columns [0,len) of the top-left len x len corner (:113-172; the
 * kernel ping-pongs through a transposed scratch tile, which is the same
 * thing).  Output is (char)-truncated (:189). */
void gsto_inverse_wavelet_tile(const uint8_t *in /*1024*/, int8_t *out /*1024 row-major*/) {
  int32_t a[GSTO_TILE * GSTO_TILE], b[GSTO_TILE * GSTO_TILE];
  for (int i = 0; i < GSTO_TILE * GSTO_TILE; ++i) a[i] = (int32_t)in[i] - 128;
  for (int len = 2; len <= GSTO_TILE; len *= 2) {
    for (int y = 0; y < len; ++y) gsto_inv_lift(a + y * GSTO_TILE, 1, b + y * GSTO_TILE, 1, len);
    for (int x = 0; x < len; ++x) gsto_inv_lift(b + x, GSTO_TILE, a + x, GSTO_TILE, len);
  }
  for (int i = 0; i < GSTO_TILE * GSTO_TILE; ++i) out[i] = (int8_t)(uint8_t)(uint32_t)a[i];
}

/* A whole plane: consecutive tiles in the ANS output, tile index =
 * tile_y * tiles_x + tile_x (:94-95), written to a row-major bx x by plane
 * (:175-191). */
void gsto_inverse_wavelet_plane(const uint8_t *in, uint32_t bx, uint32_t by, int8_t *out) {
  const uint32_t tiles_x = bx / GSTO_TILE, tiles_y = by / GSTO_TILE;
  int8_t tile[GSTO_TILE * GSTO_TILE];
  for (uint32_t ty = 0; ty < tiles_y; ++ty)
    for (uint32_t tx = 0; tx < tiles_x; ++tx) {
      gsto_inverse_wavelet_tile(in + (size_t)(ty * tiles_x + tx) * GSTO_TILE * GSTO_TILE, tile);
      for (int y = 0; y < GSTO_TILE; ++y)
        memcpy(out + (size_t)(ty * GSTO_TILE + y) * bx + tx * GSTO_TILE, tile + y * GSTO_TILE, GSTO_TILE);
    }
}

/* ------------------------------------------------------------------------- */
/* Stage 5: codec/assemble.cl.  YCoCgToRGB (:39-46) with truncating division and
 * GetPixel's unmasked shift/or pack into a ushort (:48-62). */
static void gsto_ycocg_to_rgb(int y, int co, int cg, int *r, int *g, int *b) {
  const int t = y - (cg / 2);
  *g = cg + t;
  *b = (t - co) / 2;
  *r = *b + co;
}

static uint16_t gsto_pixel565(int y, int co, int cg) {
  int r, g, b;
  gsto_ycocg_to_rgb(y, co, cg, &r, &g, &b);
  uint32_t px = 0;
  px |= (uint32_t)r << 11;
  px |= (uint32_t)g << 5;
  px |= (uint32_t)b;
  return (uint16_t)px;
}

/* assemble_dxt (:64-81): planes = 6 row-major int8 planes [Y1,Y2,Co1,Cg1,Co2,Cg2]
 * (:27-37); output block = {u16 ep1,u16 ep2,u32 palette[idx]} (codec/dxt_image.h:14-21).
 * Returns -1 if an index points outside the palette (undefined on the device). */
int gsto_assemble_dxt(const int8_t *planes, const int32_t *idx, const uint8_t *palette,
                      size_t palette_bytes, size_t n, uint8_t *out) {
  for (size_t i = 0; i < n; ++i) {
    const uint16_t ep1 = gsto_pixel565(planes[0 * n + i], planes[2 * n + i], planes[3 * n + i]);
    const uint16_t ep2 = gsto_pixel565(planes[1 * n + i], planes[4 * n + i], planes[5 * n + i]);
    const size_t pi = (uint32_t)idx[i];
    if (4 * pi + 4 > palette_bytes) return -1;
    memcpy(out + 8 * i + 0, &ep1, 2);
    memcpy(out + 8 * i + 2, &ep2, 2);
    memcpy(out + 8 * i + 4, palette + 4 * pi, 4);
  }
  return 0;
}

/* assemble_rgb (:83-129): 565 -> 888 by bit replication (:102-108), always the
 * 4-colour palette (2a+b)/3,(a+2b)/3 (:110-111), 16 RGB texels per block
 * written raster into a width x height x 3 image (:117-128).  Stores are
 * uchar-truncated. */
int gsto_assemble_rgb(const int8_t *planes, const int32_t *idx, const uint8_t *palette,
                      size_t palette_bytes, uint32_t bx, uint32_t by, uint8_t *out) {
  const size_t n = (size_t)bx * by;
  for (uint32_t yb = 0; yb < by; ++yb)
    for (uint32_t xb = 0; xb < bx; ++xb) {
      const size_t i = (size_t)yb * bx + xb;
      int pal[4][3];
      gsto_ycocg_to_rgb(planes[0 * n + i], planes[2 * n + i], planes[3 * n + i], &pal[0][0], &pal[0][1], &pal[0][2]);
      gsto_ycocg_to_rgb(planes[1 * n + i], planes[4 * n + i], planes[5 * n + i], &pal[1][0], &pal[1][1], &pal[1][2]);
      for (int e = 0; e < 2; ++e) {
        pal[e][0] = (int)(((uint32_t)pal[e][0] << 3) | (uint32_t)(pal[e][0] >> 2));
        pal[e][1] = (int)(((uint32_t)pal[e][1] << 2) | (uint32_t)(pal[e][1] >> 4));
        pal[e][2] = (int)(((uint32_t)pal[e][2] << 3) | (uint32_t)(pal[e][2] >> 2));
      }
      for (int c = 0; c < 3; ++c) {
        pal[2][c] = (2 * pal[0][c] + pal[1][c]) / 3;
        pal[3][c] = (pal[0][c] + 2 * pal[1][c]) / 3;
      }
      const size_t pi = (uint32_t)idx[i];
      if (4 * pi + 4 > palette_bytes) return -1;
      uint32_t word;
      memcpy(&word, palette + 4 * pi, 4);
      for (int k = 0; k < 16; ++k) {
        const int *rgb = pal[word & 3];
        const size_t x = 4 * (size_t)xb + (k % 4), y = 4 * (size_t)yb + (k / 4);
        uint8_t *o = out + 3 * (4 * (size_t)bx * y + x);
        o[0] = (uint8_t)rgb[0];
        o[1] = (uint8_t)rgb[1];
        o[2] = (uint8_t)rgb[2];
        word >>= 2;
      }
    }
  return 0;
}

/* ------------------------------------------------------------------------- */
/* Whole path for one .gst file (layout: codec/encoder.cpp:122-144):
 * [hdr 28][freq_Y 512][freq_C 512][freq_P 512][freq_I 512][Y][chroma][palette][indices].
 * mode 0 -> DXT1 (8N bytes), mode 1 -> RGB8 (W*H*3 bytes).  Optional taps:
 * symbols (7N+P bytes), planes (6N int8), indices (N int32).  Returns 0, -1 on a
 * malformed container, -2 on a palette index out of range. */
int gsto_decode(const uint8_t *gst, size_t len, int mode, uint8_t *out, uint8_t *symbols_out,
                int8_t *planes_out, int32_t *indices_out) {
  gsto_header h;
  if (len < sizeof(h) + 2048) return -1;
  memcpy(&h, gst, sizeof(h));
  if (h.width == 0 || h.height == 0 || (h.width % 128) || (h.height % 128)) return -1;
  const uint32_t bx = h.width / 4, by = h.height / 4;
  const size_t n = (size_t)bx * by;
  const size_t group = GSTO_LANES * GSTO_SYMS_PER_LANE;
  if ((n % group) || (h.palette_bytes % group)) return -1;
  const size_t sym_sz[4] = {2 * n, 4 * n, h.palette_bytes, n};
  const size_t cmp_sz[4] = {h.y_cmp_sz, h.chroma_cmp_sz, h.palette_sz, h.indices_sz};
  if (sizeof(h) + 2048 + cmp_sz[0] + cmp_sz[1] + cmp_sz[2] + cmp_sz[3] > len) return -1;
  const uint8_t *freqs = gst + sizeof(h), *payload = freqs + 2048;

  const size_t total = 7 * n + h.palette_bytes;
  uint8_t *symbols = (uint8_t *)malloc(total);
  int8_t *planes = (int8_t *)malloc(6 * n);
  int32_t *idx = (int32_t *)malloc(4 * n);
  size_t in_off = 0, out_off = 0, out_offs[4];
  for (int s = 0; s < 4; ++s) { /* offsets as UploadData builds them, codec/decoder.cpp:438-463 */
    out_offs[s] = out_off;
    gsto_ans_decode_stream(freqs + 512 * s, payload + in_off, sym_sz[s], symbols + out_off);
    in_off += cmp_sz[s];
    out_off += sym_sz[s];
  }
  for (int p = 0; p < 6; ++p) gsto_inverse_wavelet_plane(symbols + p * n, bx, by, planes + p * n);
  gsto_decode_indices(symbols + out_offs[3], n, idx);
  int rc = 0;
  if (out) {
    rc = mode == 0 ? gsto_assemble_dxt(planes, idx, symbols + out_offs[2], h.palette_bytes, n, out)
                   : gsto_assemble_rgb(planes, idx, symbols + out_offs[2], h.palette_bytes, bx, by, out);
    if (rc) rc = -2;
  }
  if (symbols_out) memcpy(symbols_out, symbols, total);
  if (planes_out) memcpy(planes_out, planes, 6 * n);
  if (indices_out) memcpy(indices_out, idx, 4 * n);
  free(symbols);
  free(planes);
  free(idx);
  return rc;
}
