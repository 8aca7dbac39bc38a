// gst_b200 -- sm_100a kernels for the .gst -> DXT1 decode path.
//
// Kernel inventory (reference stage it replaces, citations relative to the reference tree):
//   build_tables_kernel   stage 1  ans/build_table.cl:12-83
//   side_streams_kernel   stage 2 for the palette + index streams, with stage 3
//                         (codec/decode_indices.cl:6-84, host loop codec/decoder.cpp:311-393)
//                         fused behind the rANS warp as a group-local prefix sum
//   index_carry_kernel    the cross-group part of stage 3 (exclusive scan of group totals)
//   fused_planes_kernel   stage 2 for the six endpoint planes (ans/ans_decode.cl:25-143),
//                         stage 4 (codec/inverse_wavelet.cl:69-192) and stage 5
//                         (codec/assemble.cl:64-129) in one CTA; the symbol bytes and the
//                         wavelet planes never leave shared memory
//   ans_decode_plain_kernel  the standalone `ans_decode` entry (ans/ans_decode.cl:76-95),
//                         1..32 interleaved lanes, used by the OpenCLDecoder-style API
//
// All arithmetic is integer and follows the reference bit for bit: wrapping u32 rANS
// state, C truncating division in the 5/3 lifting and in YCoCg->RGB, (char) truncation of
// the wavelet output, unmasked shift/or 565 pack.
#include "gst_kernels.cuh"

namespace gst {
namespace {

// ---------------------------------------------------------------------------------------
// small PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void *g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.wait_all;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t lanemask_gt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_gt;" : "=r"(m));
  return m;
}
__device__ __forceinline__ void st_global_cs_v4(void *p, uint4 v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};\n" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

// ---------------------------------------------------------------------------------------
// Stage 2 core: one warp decodes one group of `n_lanes` interleaved rANS streams.
//
// Stream layout (codec/entropy.cpp:199-262): stream = [u32 end_offset[groups]][group 0]...;
// a group ends with its n_lanes 32-bit states, preceded by the shared 16-bit renorm words,
// which are consumed backwards, higher lanes first (ans/ans_decode.cl:30-32,51-65).
//
// The renorm words are staged through a per-warp shared-memory ring filled with cp.async in
// 256-byte, 256-byte-aligned chunks (group ranges are only 4-byte aligned, so windows are
// aligned down in absolute address space).  A checkpoint every 4 symbols tops the ring up;
// 4 symbols consume at most 4*32*2 = 256 B, and a refill is issued whenever fewer than 512 B
// are staged, so a chunk is always complete one checkpoint before its first byte is needed.
constexpr int kRing = 1024;
constexpr int kChunk = 256;
constexpr int kRunStride = 264;                    // one lane's 256 symbols + 8 B pad (66 words:
                                                   // conflict-free 8-byte stores across lanes)
constexpr int kStagePlane = kLanes * kRunStride;   // 8448 B: one decoded group

// emit(m, lo, hi): called 32 times; the 8 symbols at positions q0 = 248 - 8m .. q0 + 7 of this
// lane's 256-symbol run, little-endian packed (lo = q0..q0+3, hi = q0+4..q0+7).
template <class Emit>
__device__ __forceinline__ void rans_decode_group(const uint32_t *__restrict__ tab,
                                                  const uint8_t *__restrict__ stream,
                                                  uint32_t group, uint32_t n_lanes, uint8_t *ring,
                                                  const uint8_t *buf_lo, const uint8_t *buf_hi,
                                                  Emit emit) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t gt = lanemask_gt();
  const bool active = lane < n_lanes;
  const uint32_t ring_s = smem_u32(ring);

  // ans/ans_decode.cl:30.  Clamp so a malformed offset can never leave [buf_lo, buf_hi).
  const uint32_t end = __ldg(reinterpret_cast<const uint32_t *>(stream) + group) & ~3u;
  uintptr_t top = reinterpret_cast<uintptr_t>(stream) + end;
  const uintptr_t lo_ok = reinterpret_cast<uintptr_t>(buf_lo) + 4 * n_lanes;
  const uintptr_t hi_ok = reinterpret_cast<uintptr_t>(buf_hi) & ~static_cast<uintptr_t>(3);
  top = top < lo_ok ? lo_ok : top;
  top = top > hi_ok ? hi_ok : top;
  const uintptr_t a_pos = top - 4 * n_lanes;  // one past the last renorm word

  // ans/ans_decode.cl:31
  uint32_t state = active ? __ldg(reinterpret_cast<const uint32_t *>(a_pos) + lane) : 0u;

  // preload [c_top - 768, c_top), c_top = a_pos rounded up to 256
  const uintptr_t lo16 = (reinterpret_cast<uintptr_t>(buf_lo) + 15) & ~static_cast<uintptr_t>(15);
  const uintptr_t hi16 = reinterpret_cast<uintptr_t>(buf_hi) & ~static_cast<uintptr_t>(15);
  const uintptr_t c_top = (a_pos + 255) & ~static_cast<uintptr_t>(255);
  uintptr_t lo = c_top - 3 * kChunk;
  {
    uintptr_t a = lo + 16 * lane;
    if (a >= lo16 && a + 16 <= hi16) cp_async16(ring_s + (a & (kRing - 1)), reinterpret_cast<const void *>(a));
    a += 512;
    if (lane < 16 && a >= lo16 && a + 16 <= hi16)
      cp_async16(ring_s + (a & (kRing - 1)), reinterpret_cast<const void *>(a));
  }

  uint32_t cur = static_cast<uint32_t>(a_pos);  // low address bits; next word is at cur - 2

#pragma unroll 1
  for (int m = 0; m < 32; ++m) {
    uint32_t acc[2] = {0u, 0u};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      // checkpoint
      cp_async_wait_all();
      __syncwarp();
      if (cur - static_cast<uint32_t>(lo) < 512u) {
        lo -= kChunk;
        const uintptr_t a = lo + 16 * lane;
        if (lane < 16 && a >= lo16 && a + 16 <= hi16)
          cp_async16(ring_s + (a & (kRing - 1)), reinterpret_cast<const void *>(a));
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        // ans/ans_decode.cl:38-41 with freq / (slot - cum_freq) / symbol packed in one word
        const uint32_t e = tab[state & (kTableSize - 1)];
        state = (state >> kTableLog) * ((e >> 8) & 0xFFFu) + (e >> 20);
        // ans/ans_decode.cl:44-57: lanes below L pull one 16-bit word, higher lanes first
        const bool need = active && state < kRansL;
        const uint32_t mask = __ballot_sync(0xffffffffu, need);
        if (need) {
          const uint32_t a = cur - 2u - 2u * __popc(mask & gt);
          const uint32_t w = *reinterpret_cast<const uint16_t *>(ring + (a & (kRing - 1)));
          state = (state << 16) | w;
        }
        cur -= 2u * __popc(mask);  // ans/ans_decode.cl:65
        acc[1 - h] = __byte_perm(acc[1 - h], e, 0x2104);  // acc = acc << 8 | symbol
      }
    }
    emit(m, acc[0], acc[1]);
  }
  cp_async_wait_all();
}

// Copy one staged group (n_runs lane-runs of 256 B) to global memory, coalesced.
__device__ __forceinline__ void copy_stage_to_global(const uint8_t *stage, uint8_t *dst,
                                                     uint32_t n_runs) {
  const uint32_t lane = threadIdx.x & 31;
  for (uint32_t r = 0; r < n_runs; ++r) {
    const uint2 v = *reinterpret_cast<const uint2 *>(stage + r * kRunStride + lane * 8);
    *reinterpret_cast<uint2 *>(dst + r * 256 + lane * 8) = v;
  }
}

// ---------------------------------------------------------------------------------------
// Stage 1.  ans/build_table.cl:12-83: 256 frequencies -> for every slot in [0, 2048) the
// symbol x with cum[x] <= slot < cum[x+1].  The reference scans then binary-searches per slot;
// here every symbol with a non-zero frequency drops its id at slot cum[x] and a max-scan
// spreads it -- the result is determined by the frequencies alone, so it is identical.
__global__ void __launch_bounds__(256) build_tables_kernel(const uint8_t *__restrict__ freqs,
                                                           uint32_t *__restrict__ tables) {
  __shared__ uint32_t s_freq[256];
  __shared__ uint32_t s_cum[256];
  __shared__ uint32_t s_sym[kTableSize];
  __shared__ uint32_t s_warp[8];
  const uint32_t t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const uint16_t *f16 = reinterpret_cast<const uint16_t *>(freqs + 512ull * blockIdx.x);
  const uint32_t f = f16[t];

  // exclusive scan of the 256 frequencies
  uint32_t inc = f;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t n = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += n;
  }
  if (lane == 31) s_warp[warp] = inc;
  for (int i = t; i < kTableSize; i += 256) s_sym[i] = 0;
  __syncthreads();
  uint32_t base = 0;
  for (uint32_t w = 0; w < warp; ++w) base += s_warp[w];
  const uint32_t cum = (base + inc - f) & 0xFFFFu;  // ushort arithmetic, build_table.cl:14
  s_freq[t] = f;
  s_cum[t] = cum;
  if (f != 0 && cum < kTableSize) s_sym[cum] = t;
  __syncthreads();

  // inclusive max-scan over the 2048 slots, 8 consecutive slots per thread
  uint32_t v[8];
  uint32_t run = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    run = max(run, s_sym[t * 8 + i]);
    v[i] = run;
  }
  uint32_t wmax = run;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t n = __shfl_up_sync(0xffffffffu, wmax, d);
    if (lane >= d) wmax = max(wmax, n);
  }
  __syncthreads();
  if (lane == 31) s_warp[warp] = wmax;
  __syncthreads();
  uint32_t prev = __shfl_up_sync(0xffffffffu, wmax, 1);
  if (lane == 0) prev = 0;
  for (uint32_t w = 0; w < warp; ++w) prev = max(prev, s_warp[w]);

  uint32_t *out = tables + static_cast<size_t>(kTableSize) * blockIdx.x;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t slot = t * 8 + i;
    const uint32_t sym = max(v[i], prev);
    out[slot] = pack_entry(sym, s_freq[sym], (slot - s_cum[sym]) & 0xFFFu);
  }
}

// ---------------------------------------------------------------------------------------
// helpers to read the reference's device-side offset table (codec/decoder.cpp:430-463)
struct ImageStreams {
  const uint8_t *stream[4];
  uint32_t out_off[4];
  uint32_t palette_bytes;
  uint32_t pal_off;  // offset of this image's palette inside the compact palette scratch
};

__device__ __forceinline__ ImageStreams image_streams(const BatchParams &p, uint32_t b) {
  ImageStreams s;
  const uint32_t *out_off = reinterpret_cast<const uint32_t *>(p.cmp);
  const uint32_t *in_off = out_off + 4 * p.n_images;
  const uint8_t *payload = p.cmp + p.off_region + 2048ull * p.n_images;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    s.out_off[i] = __ldg(out_off + 4 * b + i);
    s.stream[i] = payload + __ldg(in_off + 4 * b + i);
  }
  s.palette_bytes = s.out_off[3] - s.out_off[2];
  s.pal_off = s.out_off[2] - 7u * p.n_blocks * b - 6u * p.n_blocks;
  return s;
}

__device__ __forceinline__ void load_table(uint32_t *dst, const uint32_t *__restrict__ src,
                                           uint32_t tid, uint32_t n_threads) {
  const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
  uint4 *d4 = reinterpret_cast<uint4 *>(dst);
  for (uint32_t i = tid; i < kTableSize / 4; i += n_threads) d4[i] = __ldg(s4 + i);
}

// ---------------------------------------------------------------------------------------
// Palette + index streams.  One warp per rANS group, 4 warps per CTA, all warps of a CTA on
// the same stream (one table in shared memory).
//   palette groups: symbols -> compact palette scratch (these are the u32 DXT index words,
//                   codec/encoder.cpp:100-108)
//   index groups:   symbols are (delta + 128) per DXT block in raster order
//                   (codec/dxt_image.cpp:610-618); the warp writes the group-local inclusive
//                   prefix sum of (byte - 128) (codec/decode_indices.cl:24) and the group total.
constexpr int kSideWarps = 4;
constexpr int kSideSmem = kTableSize * 4 + kSideWarps * (kRing + kStagePlane);

__global__ void __launch_bounds__(kSideWarps * 32)
    side_streams_kernel(const BatchParams p, uint32_t pal_ctas, uint32_t idx_ctas) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint32_t *tab = reinterpret_cast<uint32_t *>(smem);
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t *ring = smem + kTableSize * 4 + warp * kRing;
  uint8_t *stage = smem + kTableSize * 4 + kSideWarps * kRing + warp * kStagePlane;

  const uint32_t per_image = pal_ctas + idx_ctas;
  const uint32_t b = blockIdx.x / per_image;
  const uint32_t r = blockIdx.x % per_image;
  const bool is_index = r >= pal_ctas;
  const uint32_t type = is_index ? 3u : 2u;
  const ImageStreams is = image_streams(p, b);
  const uint32_t n_groups = is_index ? p.groups_per_plane : is.palette_bytes / kGroupSyms;
  const uint32_t first = (is_index ? r - pal_ctas : r) * kSideWarps;
  if (first >= n_groups) return;

  load_table(tab, p.tables + (4ull * b + type) * kTableSize, threadIdx.x, kSideWarps * 32);
  __syncthreads();
  const uint32_t group = first + warp;
  if (group >= n_groups) return;

  uint32_t bytesum = 0;
  rans_decode_group(tab, is.stream[type], group, kLanes, ring, p.cmp, p.cmp + p.cmp_bytes,
                    [&](int m, uint32_t lo, uint32_t hi) {
                      *reinterpret_cast<uint2 *>(stage + lane * kRunStride + 248 - 8 * m) =
                          make_uint2(lo, hi);
                      bytesum = __dp4a(lo, 0x01010101u, bytesum);
                      bytesum = __dp4a(hi, 0x01010101u, bytesum);
                    });
  __syncwarp();

  if (p.tap_symbols)
    copy_stage_to_global(stage, p.tap_symbols + is.out_off[type] + static_cast<size_t>(group) * kGroupSyms,
                         kLanes);

  if (!is_index) {
    const uint64_t off = static_cast<uint64_t>(is.pal_off) + static_cast<uint64_t>(group) * kGroupSyms;
    if (off + kGroupSyms <= p.palette_cap) copy_stage_to_global(stage, p.palette + off, kLanes);
    return;
  }

  // lane l holds symbols [256 l, 256 l + 256) of the group = consecutive raster blocks
  const uint32_t lane_total = bytesum - 128u * kSymsPerLane;
  uint32_t inc = lane_total;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t n = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += n;
  }
  if (lane == 31) p.idx_total[static_cast<size_t>(b) * p.groups_per_plane + group] = static_cast<int32_t>(inc);
  uint32_t run = inc - lane_total;
  int32_t *dst = p.idx_local + static_cast<size_t>(b) * p.n_blocks +
                 static_cast<size_t>(group) * kGroupSyms + lane * kSymsPerLane;
  const uint8_t *src = stage + lane * kRunStride;
#pragma unroll 2
  for (int k = 0; k < 32; ++k) {
    const uint2 w = *reinterpret_cast<const uint2 *>(src + 8 * k);
    uint32_t o[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      run += ((w.x >> (8 * j)) & 0xFFu) - 128u;
      o[j] = run;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      run += ((w.y >> (8 * j)) & 0xFFu) - 128u;
      o[4 + j] = run;
    }
    *reinterpret_cast<uint4 *>(dst + 8 * k) = make_uint4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<uint4 *>(dst + 8 * k + 4) = make_uint4(o[4], o[5], o[6], o[7]);
  }
}

// Exclusive scan of the per-group totals of one image (the collect_indices passes of
// codec/decode_indices.cl:66-84 collapse to this).
__global__ void __launch_bounds__(32) index_carry_kernel(const BatchParams p) {
  const uint32_t lane = threadIdx.x;
  const int32_t *tot = p.idx_total + static_cast<size_t>(blockIdx.x) * p.groups_per_plane;
  int32_t *car = p.idx_carry + static_cast<size_t>(blockIdx.x) * p.groups_per_plane;
  uint32_t base = 0;
  for (uint32_t i = 0; i < p.groups_per_plane; i += 32) {
    const uint32_t v = (i + lane < p.groups_per_plane) ? static_cast<uint32_t>(tot[i + lane]) : 0u;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t n = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += n;
    }
    if (i + lane < p.groups_per_plane) car[i + lane] = static_cast<int32_t>(base + inc - v);
    base += __shfl_sync(0xffffffffu, inc, 31);
  }
}

// ---------------------------------------------------------------------------------------
// Stage 4 helpers.  1-D inverse 5/3 lifting of v = [low half | high half] in registers,
// codec/inverse_wavelet.cl:28-64 (NormalizeIndex mirror resolved at compile time):
//   even: d[2x]   = s[x]       - (s[mid + max(x-1,0)] + s[mid + x] + 2) / 4
//   odd : d[2x+1] = s[mid + x] + (d[2x] + d[min(2x+2, len-2)]) / 2          ('/' truncates)
template <int LEN>
__device__ __forceinline__ void inverse_lift(int (&v)[LEN]) {
  constexpr int MID = LEN / 2;
  int o[LEN];
#pragma unroll
  for (int x = 0; x < MID; ++x) {
    const int hp = v[MID + (x == 0 ? 0 : x - 1)];
    const int hn = v[MID + x];
    o[2 * x] = v[x] - (hp + hn + 2) / 4;
  }
#pragma unroll
  for (int x = 0; x < MID; ++x) {
    const int ep = o[2 * x];
    const int en = o[(2 * x + 2 == LEN) ? 2 * x : 2 * x + 2];
    o[2 * x + 1] = v[MID + x] + (ep + en) / 2;
  }
#pragma unroll
  for (int i = 0; i < LEN; ++i) v[i] = o[i];
}

constexpr int kWRow = 34;                       // int16 per work-tile row (17 words: odd stride)
constexpr int kWBytes = kTile * kWRow * 2;      // 2176 B per warp

// One level on the top-left LEN x LEN corner of the warp's int16 work tile: rows then
// columns (codec/inverse_wavelet.cl:113-172).  Intermediates are bounded by 128 + 672 per
// level (<= 3488 after five levels) for ANY input bytes, so int16 storage is exact.
template <int LEN>
__device__ __forceinline__ void wavelet_rows(int16_t *W, uint32_t lane) {
  if (lane < LEN) {
    uint32_t *row = reinterpret_cast<uint32_t *>(W + lane * kWRow);
    int v[LEN];
#pragma unroll
    for (int i = 0; i < LEN / 2; ++i) {
      const uint32_t w = row[i];
      v[2 * i] = static_cast<int16_t>(w & 0xFFFFu);
      v[2 * i + 1] = static_cast<int32_t>(w) >> 16;
    }
    inverse_lift<LEN>(v);
#pragma unroll
    for (int i = 0; i < LEN / 2; ++i)
      row[i] = (static_cast<uint32_t>(v[2 * i]) & 0xFFFFu) | (static_cast<uint32_t>(v[2 * i + 1]) << 16);
  }
  __syncwarp();
}

template <int LEN>
__device__ __forceinline__ void wavelet_cols(int16_t *W, uint32_t lane) {
  if (lane < LEN) {
    int v[LEN];
#pragma unroll
    for (int i = 0; i < LEN; ++i) v[i] = W[i * kWRow + lane];
    inverse_lift<LEN>(v);
#pragma unroll
    for (int i = 0; i < LEN; ++i) W[i * kWRow + lane] = static_cast<int16_t>(v[i]);
  }
  __syncwarp();
}

// codec/assemble.cl:39-62: YCoCg667 -> RGB565 with truncating division and an unmasked pack.
__device__ __forceinline__ void ycocg_to_rgb(int y, int co, int cg, int &r, int &g, int &b) {
  const int t = y - (cg / 2);
  g = cg + t;
  b = (t - co) / 2;
  r = b + co;
}
__device__ __forceinline__ uint32_t pack565(int y, int co, int cg) {
  int r, g, b;
  ycocg_to_rgb(y, co, cg, r, g, b);
  return ((static_cast<uint32_t>(r) << 11) | (static_cast<uint32_t>(g) << 5) | static_cast<uint32_t>(b)) & 0xFFFFu;
}
__device__ __forceinline__ int sbyte(uint32_t w, int j) {
  return static_cast<int8_t>((w >> (8 * j)) & 0xFFu);
}

// ---------------------------------------------------------------------------------------
// The fused endpoint-plane kernel.  CTA (b, g) owns tiles [8g, 8g+8) of image b in all six
// planes [Y1,Y2,Co1,Cg1,Co2,Cg2] (codec/assemble.cl:27-37) = one rANS group per plane.
//   phase 1: warps 0..5 rANS-decode one group each into shared memory (48 KB of symbols)
//   phase 2: warp t runs the 5-level inverse wavelet on tile t of every plane in a private
//            int16 work tile (rows in registers, lane = row, then lane = column)
//   phase 3: warp t assembles the 1024 DXT1 blocks (or RGB8 texels) of tile t with coalesced
//            16-byte stores; the palette word comes from the index prefix written by
//            side_streams_kernel plus the per-group carry.
constexpr int kFusedWarps = 8;
constexpr int kFusedAux = 2 * kTableSize * 4 + 6 * kRing;  // tables + rings, reused by work tiles
static_assert(kFusedAux >= kFusedWarps * kWBytes, "work tiles must fit in the aliased region");
constexpr int kFusedSmem = 6 * kStagePlane + kFusedAux;    // 73216 B -> 3 CTAs / SM

template <int RGB>
__global__ void __launch_bounds__(kFusedWarps * 32, 3) fused_planes_kernel(const BatchParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint8_t *stage = smem;
  uint8_t *aux = smem + 6 * kStagePlane;
  uint32_t *tabs = reinterpret_cast<uint32_t *>(aux);
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t b = blockIdx.x / p.groups_per_plane;
  const uint32_t g = blockIdx.x % p.groups_per_plane;
  const ImageStreams is = image_streams(p, b);

  // Y table and chroma table of this image
  load_table(tabs, p.tables + (4ull * b + 0) * kTableSize, threadIdx.x, kFusedWarps * 32);
  load_table(tabs + kTableSize, p.tables + (4ull * b + 1) * kTableSize, threadIdx.x, kFusedWarps * 32);
  __syncthreads();

  if (warp < 6) {
    // Y stream = Y1 || Y2, chroma stream = Co1 || Cg1 || Co2 || Cg2 (codec/encoder.cpp:87,93-95)
    const uint32_t type = warp < 2 ? 0u : 1u;
    const uint32_t group = (warp < 2 ? warp : warp - 2) * p.groups_per_plane + g;
    uint8_t *my_stage = stage + warp * kStagePlane + lane * kRunStride + 248;
    rans_decode_group(tabs + type * kTableSize, is.stream[type], group, kLanes,
                      aux + 2 * kTableSize * 4 + warp * kRing, p.cmp, p.cmp + p.cmp_bytes,
                      [&](int m, uint32_t lo, uint32_t hi) {
                        *reinterpret_cast<uint2 *>(my_stage - 8 * m) = make_uint2(lo, hi);
                      });
  }
  __syncthreads();

  if (p.tap_symbols) {
    for (uint32_t pl = 0; pl < 6; ++pl) {
      const uint32_t type = pl < 2 ? 0u : 1u;
      const uint32_t group = (pl < 2 ? pl : pl - 2) * p.groups_per_plane + g;
      uint8_t *dst = p.tap_symbols + is.out_off[type] + static_cast<size_t>(group) * kGroupSyms;
      for (uint32_t r = warp; r < kLanes; r += kFusedWarps) {
        const uint2 v = *reinterpret_cast<const uint2 *>(stage + pl * kStagePlane + r * kRunStride + lane * 8);
        *reinterpret_cast<uint2 *>(dst + r * 256 + lane * 8) = v;
      }
    }
    __syncthreads();
  }

  // ---- phase 2: inverse wavelet, warp = tile position --------------------------------
  const uint32_t tiles_x = p.blocks_x / kTile;
  const uint32_t tile = g * 8 + warp;
  const uint32_t ty = tile / tiles_x, tx = tile % tiles_x;
  int16_t *W = reinterpret_cast<int16_t *>(aux + warp * kWBytes);
  uint8_t *tile_stage = stage + 4 * warp * kRunStride;  // + plane * kStagePlane

  for (uint32_t pl = 0; pl < 6; ++pl) {
    uint8_t *ts = tile_stage + pl * kStagePlane;
    // bytes -> (byte - 128) as int16, codec/inverse_wavelet.cl:97-100
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const uint32_t row = 4 * s + (lane >> 3), col = 4 * (lane & 7);
      const uint32_t w = *reinterpret_cast<const uint32_t *>(ts + (row >> 3) * kRunStride + (row & 7) * 32 + col);
      const int v0 = static_cast<int>(w & 0xFFu) - 128, v1 = static_cast<int>((w >> 8) & 0xFFu) - 128;
      const int v2 = static_cast<int>((w >> 16) & 0xFFu) - 128, v3 = static_cast<int>(w >> 24) - 128;
      uint32_t *d = reinterpret_cast<uint32_t *>(W + row * kWRow + col);
      d[0] = (static_cast<uint32_t>(v0) & 0xFFFFu) | (static_cast<uint32_t>(v1) << 16);
      d[1] = (static_cast<uint32_t>(v2) & 0xFFFFu) | (static_cast<uint32_t>(v3) << 16);
    }
    __syncwarp();
    wavelet_rows<2>(W, lane);
    wavelet_cols<2>(W, lane);
    wavelet_rows<4>(W, lane);
    wavelet_cols<4>(W, lane);
    wavelet_rows<8>(W, lane);
    wavelet_cols<8>(W, lane);
    wavelet_rows<16>(W, lane);
    wavelet_cols<16>(W, lane);
    wavelet_rows<32>(W, lane);
    // last column pass: (char) truncation (codec/inverse_wavelet.cl:188-190) back into the
    // staging slot of this tile, row-major bytes
    {
      int v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = W[i * kWRow + lane];
      inverse_lift<32>(v);
#pragma unroll
      for (int i = 0; i < 32; ++i) ts[(i >> 3) * kRunStride + (i & 7) * 32 + lane] = static_cast<uint8_t>(v[i]);
      if (p.tap_planes) {
        int8_t *tp = p.tap_planes + (static_cast<size_t>(b) * 6 + pl) * p.n_blocks +
                     static_cast<size_t>(ty * kTile) * p.blocks_x + tx * kTile + lane;
#pragma unroll
        for (int i = 0; i < 32; ++i) tp[static_cast<size_t>(i) * p.blocks_x] = static_cast<int8_t>(v[i]);
      }
    }
    __syncwarp();
  }

  // ---- phase 3: assembly, codec/assemble.cl:64-129 -------------------------------------
  const uint32_t n_entries = is.palette_bytes / 4;
  const bool pal_ok = static_cast<uint64_t>(is.pal_off) + is.palette_bytes <= p.palette_cap && n_entries > 0;
  const uint32_t *pal = reinterpret_cast<const uint32_t *>(p.palette + (pal_ok ? is.pal_off : 0));
  const int32_t *loc_base = p.idx_local + static_cast<size_t>(b) * p.n_blocks;
  const int32_t *carry_base = p.idx_carry + static_cast<size_t>(b) * p.groups_per_plane;

#pragma unroll 1
  for (int k = 0; k < 8; ++k) {
    const uint32_t row = 4 * k + (lane >> 3), col = 4 * (lane & 7);
    const uint32_t gidx = (ty * kTile + row) * p.blocks_x + tx * kTile + col;
    uint32_t pw[6];
#pragma unroll
    for (int pl = 0; pl < 6; ++pl)
      pw[pl] = *reinterpret_cast<const uint32_t *>(tile_stage + pl * kStagePlane + (row >> 3) * kRunStride +
                                                   (row & 7) * 32 + col);
    const int4 loc = __ldg(reinterpret_cast<const int4 *>(loc_base + gidx));
    const uint32_t carry = static_cast<uint32_t>(__ldg(carry_base + gidx / kGroupSyms));
    const uint32_t idx[4] = {carry + static_cast<uint32_t>(loc.x), carry + static_cast<uint32_t>(loc.y),
                             carry + static_cast<uint32_t>(loc.z), carry + static_cast<uint32_t>(loc.w)};
    if (p.tap_indices)
      *reinterpret_cast<uint4 *>(p.tap_indices + static_cast<size_t>(b) * p.n_blocks + gidx) =
          make_uint4(idx[0], idx[1], idx[2], idx[3]);
    uint32_t word[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) word[j] = pal_ok ? __ldg(pal + min(idx[j], n_entries - 1)) : 0u;

    if (!RGB) {
      uint32_t o[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t ep1 = pack565(sbyte(pw[0], j), sbyte(pw[2], j), sbyte(pw[3], j));
        const uint32_t ep2 = pack565(sbyte(pw[1], j), sbyte(pw[4], j), sbyte(pw[5], j));
        o[2 * j] = ep1 | (ep2 << 16);  // PhysicalDXTBlock, codec/dxt_image.h:14-21
        o[2 * j + 1] = word[j];
      }
      uint8_t *dst = p.out + (static_cast<size_t>(b) * p.n_blocks + gidx) * 8;
      st_global_cs_v4(dst, make_uint4(o[0], o[1], o[2], o[3]));
      st_global_cs_v4(dst + 16, make_uint4(o[4], o[5], o[6], o[7]));
    } else {
      // assemble_rgb: 565 -> 888 by bit replication, always the 4-colour palette
      // (codec/assemble.cl:102-111); 16 texels per block, raster RGB8 (:117-128)
      uint8_t texel[4][4][3];  // [block j][palette entry][channel]
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int c0[3], c1[3];
        ycocg_to_rgb(sbyte(pw[0], j), sbyte(pw[2], j), sbyte(pw[3], j), c0[0], c0[1], c0[2]);
        ycocg_to_rgb(sbyte(pw[1], j), sbyte(pw[4], j), sbyte(pw[5], j), c1[0], c1[1], c1[2]);
        c0[0] = static_cast<int>((static_cast<uint32_t>(c0[0]) << 3) | static_cast<uint32_t>(c0[0] >> 2));
        c0[1] = static_cast<int>((static_cast<uint32_t>(c0[1]) << 2) | static_cast<uint32_t>(c0[1] >> 4));
        c0[2] = static_cast<int>((static_cast<uint32_t>(c0[2]) << 3) | static_cast<uint32_t>(c0[2] >> 2));
        c1[0] = static_cast<int>((static_cast<uint32_t>(c1[0]) << 3) | static_cast<uint32_t>(c1[0] >> 2));
        c1[1] = static_cast<int>((static_cast<uint32_t>(c1[1]) << 2) | static_cast<uint32_t>(c1[1] >> 4));
        c1[2] = static_cast<int>((static_cast<uint32_t>(c1[2]) << 3) | static_cast<uint32_t>(c1[2] >> 2));
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          texel[j][0][c] = static_cast<uint8_t>(c0[c]);
          texel[j][1][c] = static_cast<uint8_t>(c1[c]);
          texel[j][2][c] = static_cast<uint8_t>((2 * c0[c] + c1[c]) / 3);
          texel[j][3][c] = static_cast<uint8_t>((c0[c] + 2 * c1[c]) / 3);
        }
      }
      const size_t img_w = 4ull * p.blocks_x;
      uint8_t *img = p.out + static_cast<size_t>(b) * p.n_blocks * 48;
      const size_t x0 = 4ull * (tx * kTile + col), y0 = 4ull * (ty * kTile + row);
#pragma unroll
      for (int yy = 0; yy < 4; ++yy) {
        uint32_t bytes[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) bytes[i] = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
          for (int xx = 0; xx < 4; ++xx) {
            const uint32_t sel = (word[j] >> (2 * (4 * yy + xx))) & 3u;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const uint32_t v = sel == 0 ? texel[j][0][c] : sel == 1 ? texel[j][1][c] : sel == 2 ? texel[j][2][c] : texel[j][3][c];
              const int pos = 12 * j + 3 * xx + c;
              bytes[pos >> 2] |= v << (8 * (pos & 3));
            }
          }
        }
        uint8_t *dst = img + 3 * (img_w * (y0 + yy) + x0);
        st_global_cs_v4(dst, make_uint4(bytes[0], bytes[1], bytes[2], bytes[3]));
        st_global_cs_v4(dst + 16, make_uint4(bytes[4], bytes[5], bytes[6], bytes[7]));
        st_global_cs_v4(dst + 32, make_uint4(bytes[8], bytes[9], bytes[10], bytes[11]));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Standalone decode of [u32 end_offset[n_groups]][groups] with 1..32 interleaved lanes and a
// single table: the `ans_decode` kernel of ans/ans_decode.cl:76-95 as driven by
// ans/ans_ocl.cpp:159-345.  Output: group * n_lanes * 256 + lane * 256 + position.
constexpr int kPlainWarps = 4;
constexpr int kPlainSmem = kTableSize * 4 + kPlainWarps * (kRing + kStagePlane);

__global__ void __launch_bounds__(kPlainWarps * 32)
    ans_decode_plain_kernel(const uint32_t *__restrict__ table, const uint8_t *__restrict__ data,
                            uint64_t data_bytes, uint32_t n_groups, uint32_t n_lanes,
                            uint8_t *__restrict__ out) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint32_t *tab = reinterpret_cast<uint32_t *>(smem);
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t *ring = smem + kTableSize * 4 + warp * kRing;
  uint8_t *stage = smem + kTableSize * 4 + kPlainWarps * kRing + warp * kStagePlane;
  load_table(tab, table, threadIdx.x, kPlainWarps * 32);
  __syncthreads();
  const uint32_t group = blockIdx.x * kPlainWarps + warp;
  if (group >= n_groups) return;
  rans_decode_group(tab, data, group, n_lanes, ring, data, data + data_bytes,
                    [&](int m, uint32_t lo, uint32_t hi) {
                      *reinterpret_cast<uint2 *>(stage + lane * kRunStride + 248 - 8 * m) = make_uint2(lo, hi);
                    });
  __syncwarp();
  copy_stage_to_global(stage, out + static_cast<size_t>(group) * n_lanes * kSymsPerLane, n_lanes);
}

}  // namespace

// ---------------------------------------------------------------------------------------
// launchers
cudaError_t launch_build_tables(const uint8_t *freqs, uint32_t n_tables, uint32_t *tables,
                                cudaStream_t s) {
  if (n_tables == 0) return cudaSuccess;
  build_tables_kernel<<<n_tables, 256, 0, s>>>(freqs, tables);
  return cudaGetLastError();
}

static cudaError_t ensure_attrs() {
  static cudaError_t once = []() {
    cudaError_t e = cudaFuncSetAttribute(fused_planes_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(fused_planes_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(side_streams_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSideSmem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(ans_decode_plain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPlainSmem);
  }();
  return once;
}

cudaError_t launch_decode_batch(const BatchParams &p, int rgb_mode, uint32_t max_palette_bytes,
                                cudaStream_t s, cudaEvent_t *marks) {
  if (p.n_images == 0) return cudaSuccess;
  cudaError_t e = ensure_attrs();
  if (e != cudaSuccess) return e;
  int mark = 0;
  auto stamp = [&]() { return marks ? cudaEventRecord(marks[mark++], s) : cudaSuccess; };
  if ((e = stamp()) != cudaSuccess) return e;
  // stage 1: 4 tables per image, straight from the freq region of the compressed buffer
  e = launch_build_tables(p.cmp + p.off_region, 4 * p.n_images, p.tables, s);
  if (e != cudaSuccess) return e;
  if ((e = stamp()) != cudaSuccess) return e;
  // palette + index streams
  const uint32_t max_pal_groups = max_palette_bytes / kGroupSyms;
  const uint32_t pal_ctas = (max_pal_groups + kSideWarps - 1) / kSideWarps;
  const uint32_t idx_ctas = (p.groups_per_plane + kSideWarps - 1) / kSideWarps;
  side_streams_kernel<<<p.n_images * (pal_ctas + idx_ctas), kSideWarps * 32, kSideSmem, s>>>(p, pal_ctas, idx_ctas);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if ((e = stamp()) != cudaSuccess) return e;
  index_carry_kernel<<<p.n_images, 32, 0, s>>>(p);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if ((e = stamp()) != cudaSuccess) return e;
  const uint32_t grid = p.n_images * p.groups_per_plane;
  if (rgb_mode)
    fused_planes_kernel<1><<<grid, kFusedWarps * 32, kFusedSmem, s>>>(p);
  else
    fused_planes_kernel<0><<<grid, kFusedWarps * 32, kFusedSmem, s>>>(p);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return stamp();
}

cudaError_t launch_ans_decode_plain(const uint32_t *table, const uint8_t *data, uint64_t data_bytes,
                                    uint32_t n_groups, uint32_t n_lanes, uint8_t *out, cudaStream_t s) {
  if (n_groups == 0) return cudaSuccess;
  cudaError_t e = ensure_attrs();
  if (e != cudaSuccess) return e;
  ans_decode_plain_kernel<<<(n_groups + kPlainWarps - 1) / kPlainWarps, kPlainWarps * 32, kPlainSmem, s>>>(
      table, data, data_bytes, n_groups, n_lanes, out);
  return cudaGetLastError();
}

}  // namespace gst
