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
// small PTX helpers.  Shared memory is addressed by its 32-bit shared-window address so the
// hot loops spend no instructions on generic-address arithmetic.
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
  uint32_t v;
  asm volatile("{ .reg .u16 t; ld.shared.u16 t, [%1]; cvt.u32.u16 %0, t; }" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ int lds_s16(uint32_t a) {
  int v;
  asm volatile("{ .reg .s16 t; ld.shared.s16 t, [%1]; cvt.s32.s16 %0, t; }" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts8(uint32_t a, uint32_t v) {
  asm volatile("{ .reg .u16 t; cvt.u16.u32 t, %1; st.shared.u8 [%0], t; }" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts16(uint32_t a, uint32_t v) {
  asm volatile("{ .reg .u16 t; cvt.u16.u32 t, %1; st.shared.u16 [%0], t; }" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t a, uint32_t x, uint32_t y) {
  asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
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
__device__ __forceinline__ void st_global_cs_v4(void *p, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};\n" ::"l"(p), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
// prmt.b32 in its default mode: a selector nibble with bit 3 set replicates the SIGN of the
// selected byte (the __byte_perm intrinsic only documents the low three bits, so use PTX).
template <uint32_t SEL>
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "n"(SEL));
  return d;
}
// signed byte j of w as int; byte pair (j, j+1) as a packed int16x2 word
template <int J>
__device__ __forceinline__ int sext_byte(uint32_t w) {
  return static_cast<int>(prmt<((8 | J) << 12) | ((8 | J) << 8) | ((8 | J) << 4) | J>(w, 0u));
}
template <int J>
__device__ __forceinline__ uint32_t sext_byte_pair(uint32_t w) {
  return prmt<((8 | (J + 1)) << 12) | ((J + 1) << 8) | ((8 | J) << 4) | J>(w, 0u);
}
__device__ __forceinline__ int lo16(uint32_t w) { return static_cast<int>(prmt<0x9910>(w, 0u)); }
__device__ __forceinline__ int hi16(uint32_t w) { return static_cast<int>(w) >> 16; }
__device__ __forceinline__ uint32_t pack16(int a, int b) {
  return __byte_perm(static_cast<uint32_t>(a), static_cast<uint32_t>(b), 0x5410);
}

// ---------------------------------------------------------------------------------------
// Stage 2 core: one warp decodes one group of `n_lanes` interleaved rANS streams.
//
// Stream layout (codec/entropy.cpp:199-262): stream = [u32 end_offset[groups]][group 0]...;
// a group ends with its n_lanes 32-bit states, preceded by the shared 16-bit renorm words,
// which are consumed backwards, higher lanes first (ans/ans_decode.cl:30-32,51-65).
//
// The renorm words are staged through a per-warp, 1 KiB-aligned shared-memory ring filled with
// cp.async in 256-byte, 256-byte-aligned chunks (group ranges are only 4-byte aligned, so the
// windows are aligned down in absolute address space and the ring is indexed by the low
// address bits).  A checkpoint every 4 symbols (which consume at most 4*32*2 = 256 B) tops the
// ring up whenever fewer than 768 B are staged and then waits until at most the two newest
// cp.async groups are pending: a chunk has two checkpoint intervals to land, and because at
// least 512 B are staged at every checkpoint, the 256 B the next four symbols can touch are
// always complete.
//
// Per symbol and lane (ans/ans_decode.cl:38-65):
//   e = table[state & 2047];  state = (state >> 11) * e.freq + e.bias          (bias = slot - cum)
//   lanes whose state fell below L = 2^15 take the next 16-bit word, higher lanes first:
//   word index = next - 1 - popc(ballot & lanes_above_me);  next -= popc(ballot)
// The word load is unconditional (every lane's address lies inside the staged 64 bytes), only
// the state update is predicated.
constexpr int kRing = 1024;
constexpr int kChunk = 256;

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// emit(m, lo, hi): called 32 times; the 8 symbols at positions q0 = 248 - 8m .. q0 + 7 of this
// lane's 256-symbol run, little-endian packed (lo = q0..q0+3, hi = q0+4..q0+7).
template <bool FULL, class Emit>
__device__ __forceinline__ void rans_decode_group(uint32_t tab_s, const uint8_t *__restrict__ stream,
                                                  uint32_t group, uint32_t n_lanes, uint32_t ring_s,
                                                  const uint8_t *buf_lo, const uint8_t *buf_hi, Emit emit) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t gt = lanemask_gt();
  const bool active = FULL || lane < n_lanes;

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

  // preload [c_top - 1024, c_top), c_top = a_pos rounded up to 256
  const uintptr_t lo16 = (reinterpret_cast<uintptr_t>(buf_lo) + 15) & ~static_cast<uintptr_t>(15);
  const uintptr_t hi16 = reinterpret_cast<uintptr_t>(buf_hi) & ~static_cast<uintptr_t>(15);
  const uintptr_t c_top = (a_pos + 255) & ~static_cast<uintptr_t>(255);
  uintptr_t lo = c_top - 4 * kChunk;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const uintptr_t a = lo + 16 * lane + 512 * i;
    if (a >= lo16 && a + 16 <= hi16) cp_async16(ring_s + (a & (kRing - 1)), reinterpret_cast<const void *>(a));
  }
  cp_async_commit();
  cp_async_wait_group<0>();
  __syncwarp();

  uint32_t cur2 = static_cast<uint32_t>(a_pos) - 2u;  // low address bits of the next word

#pragma unroll 1
  for (int m = 0; m < 32; ++m) {
    uint32_t acc[2] = {0u, 0u};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      // checkpoint: top up when < 768 B are staged, then let the two newest groups fly
      if (cur2 + 2u - static_cast<uint32_t>(lo) < 768u) {
        lo -= kChunk;
        const uintptr_t a = lo + 16 * lane;
        if (lane < 16 && a >= lo16 && a + 16 <= hi16)
          cp_async16(ring_s + (a & (kRing - 1)), reinterpret_cast<const void *>(a));
      }
      cp_async_commit();
      cp_async_wait_group<2>();
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint32_t slot_a;  // tab_s + 4 * (state & 2047): one LOP3 + one IMAD
        asm("mad.lo.u32 %0, %1, 4, %2;" : "=r"(slot_a) : "r"(state & (kTableSize - 1)), "r"(tab_s));
        const uint32_t e = lds32(slot_a);
        state = (state >> kTableLog) * ((e >> 8) & 0xFFFu) + (e >> 20);
        const bool need = FULL ? (state < kRansL) : (active && state < kRansL);
        const uint32_t mask = __ballot_sync(0xffffffffu, need);
        const uint32_t a = cur2 - 2u * __popc(mask & gt);
        const uint32_t w = lds_u16(ring_s | (a & (kRing - 1)));
        if (need) state = __byte_perm(w, state, 0x5410);  // state << 16 | w
        cur2 -= 2u * __popc(mask);                        // ans/ans_decode.cl:65
        acc[1 - h] = __byte_perm(acc[1 - h], e, 0x2104);  // acc << 8 | symbol
      }
    }
    emit(m, acc[0], acc[1]);
  }
  cp_async_wait_group<0>();
}

__device__ __forceinline__ void trap_unless_aligned(uint32_t s, uint32_t align) {
  if (s & (align - 1)) __trap();
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
  const uint8_t *payload;
  uint32_t in_off[4];
  uint32_t out_off[4];
  uint32_t palette_bytes;
  uint32_t pal_off;  // offset of this image's palette inside the compact palette scratch
};

__device__ __forceinline__ ImageStreams image_streams(const BatchParams &p, uint32_t b) {
  ImageStreams s;
  const uint4 *tbl = reinterpret_cast<const uint4 *>(p.cmp);
  const uint4 oo = __ldg(tbl + b), io = __ldg(tbl + p.n_images + b);
  s.payload = p.cmp + p.off_region + 2048ull * p.n_images;
  s.out_off[0] = oo.x; s.out_off[1] = oo.y; s.out_off[2] = oo.z; s.out_off[3] = oo.w;
  s.in_off[0] = io.x;  s.in_off[1] = io.y;  s.in_off[2] = io.z;  s.in_off[3] = io.w;
  s.palette_bytes = oo.w - oo.z;
  s.pal_off = oo.z - 7u * p.n_blocks * b - 6u * p.n_blocks;
  return s;
}

__device__ __forceinline__ void load_table(uint32_t dst_s, const uint32_t *__restrict__ src,
                                           uint32_t tid, uint32_t n_threads) {
  const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
  for (uint32_t i = tid; i < kTableSize / 4; i += n_threads) {
    const uint4 v = __ldg(s4 + i);
    sts128(dst_s + 16 * i, v.x, v.y, v.z, v.w);
  }
}

// ---------------------------------------------------------------------------------------
// Palette + index streams.  One warp per rANS group, 8 warps per CTA, all warps of a CTA on
// the same stream (one table in shared memory); no symbol staging, so 6 CTAs fit an SM.
//   palette groups: symbols go straight to the compact palette scratch (they are the u32 DXT
//                   index words, codec/encoder.cpp:100-108)
//   index groups:   symbols are (delta + 128) per DXT block in raster order
//                   (codec/dxt_image.cpp:610-618) and every lane owns 256 consecutive blocks
//                   (a "run").  Stage 3 (codec/decode_indices.cl:24, idx[i] = sum_{j<=i} d[j])
//                   is fused behind the decoder: symbols arrive last-to-first, so the lane
//                   writes S[i] = sum of the deltas AFTER i inside its run, and
//                   idx[i] = run_end[run] - S[i], where run_end is the inclusive prefix at the
//                   end of the run (finished by index_carry_kernel).  S is stored mod 2^16 when
//                   every palette of the batch has <= 65536 entries (idx < 2^16 then makes the
//                   16-bit difference exact), else as 32 bits.
constexpr int kSideWarps = 8;
constexpr int kSideSmem = kSideWarps * kRing + kTableSize * 4;

__global__ void __launch_bounds__(kSideWarps * 32, 6)
    side_streams_kernel(const BatchParams p, uint32_t pal_ctas, uint32_t idx_ctas) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_s = smem_u32(smem);
  trap_unless_aligned(smem_s, kRing);
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ring_s = smem_s + warp * kRing;
  const uint32_t tab_s = smem_s + kSideWarps * kRing;

  const uint32_t per_image = pal_ctas + idx_ctas;
  const uint32_t b = blockIdx.x / per_image;
  const uint32_t r = blockIdx.x % per_image;
  const bool is_index = r >= pal_ctas;
  const uint32_t type = is_index ? 3u : 2u;
  const ImageStreams is = image_streams(p, b);
  const uint32_t n_groups = is_index ? p.groups_per_plane : is.palette_bytes / kGroupSyms;
  const uint32_t first = (is_index ? r - pal_ctas : r) * kSideWarps;
  if (first >= n_groups) return;

  load_table(tab_s, p.tables + (4ull * b + type) * kTableSize, threadIdx.x, kSideWarps * 32);
  __syncthreads();
  const uint32_t group = first + warp;
  if (group >= n_groups) return;
  const uint8_t *stream = is.payload + (is_index ? is.in_off[3] : is.in_off[2]);
  const uint32_t out_off = is_index ? is.out_off[3] : is.out_off[2];
  uint8_t *tap = p.tap_symbols ? p.tap_symbols + out_off + static_cast<size_t>(group) * kGroupSyms + lane * kSymsPerLane + 248
                               : nullptr;

  if (!is_index) {
    const uint64_t off = static_cast<uint64_t>(is.pal_off) + static_cast<uint64_t>(group) * kGroupSyms;
    const bool ok = off + kGroupSyms <= p.palette_cap;
    uint8_t *dst = p.palette + off + lane * kSymsPerLane + 248;
    rans_decode_group<true>(tab_s, stream, group, kLanes, ring_s, p.cmp, p.cmp + p.cmp_bytes,
                            [&](int m, uint32_t lo, uint32_t hi) {
                              if (ok) *reinterpret_cast<uint2 *>(dst - 8 * m) = make_uint2(lo, hi);
                              if (tap) *reinterpret_cast<uint2 *>(tap - 8 * m) = make_uint2(lo, hi);
                            });
    return;
  }

  uint32_t sum = 0;  // sum of (byte - 128) over the symbols decoded so far = positions after the current one
  const size_t run0 = static_cast<size_t>(b) * p.n_blocks + static_cast<size_t>(group) * kGroupSyms + lane * kSymsPerLane + 248;
  uint16_t *dst16 = reinterpret_cast<uint16_t *>(p.idx_s) + run0;
  uint32_t *dst32 = reinterpret_cast<uint32_t *>(p.idx_s) + run0;
  const bool idx16 = p.idx16 != 0;
  rans_decode_group<true>(tab_s, stream, group, kLanes, ring_s, p.cmp, p.cmp + p.cmp_bytes,
                          [&](int m, uint32_t lo, uint32_t hi) {
                            if (tap) *reinterpret_cast<uint2 *>(tap - 8 * m) = make_uint2(lo, hi);
                            uint32_t s[8];
                            s[7] = sum; sum += ((hi >> 24) & 0xFFu) - 128u;
                            s[6] = sum; sum += ((hi >> 16) & 0xFFu) - 128u;
                            s[5] = sum; sum += ((hi >> 8) & 0xFFu) - 128u;
                            s[4] = sum; sum += (hi & 0xFFu) - 128u;
                            s[3] = sum; sum += ((lo >> 24) & 0xFFu) - 128u;
                            s[2] = sum; sum += ((lo >> 16) & 0xFFu) - 128u;
                            s[1] = sum; sum += ((lo >> 8) & 0xFFu) - 128u;
                            s[0] = sum; sum += (lo & 0xFFu) - 128u;
                            if (idx16) {
                              *reinterpret_cast<uint4 *>(dst16 - 8 * m) =
                                  make_uint4(__byte_perm(s[0], s[1], 0x5410), __byte_perm(s[2], s[3], 0x5410),
                                             __byte_perm(s[4], s[5], 0x5410), __byte_perm(s[6], s[7], 0x5410));
                            } else {
                              *reinterpret_cast<uint4 *>(dst32 - 8 * m) = make_uint4(s[0], s[1], s[2], s[3]);
                              *reinterpret_cast<uint4 *>(dst32 - 8 * m + 4) = make_uint4(s[4], s[5], s[6], s[7]);
                            }
                          });
  // group-local inclusive prefix at the end of every run, and the group total
  uint32_t inc = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t n = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += n;
  }
  p.run_end[static_cast<size_t>(b) * (p.n_blocks / kSymsPerLane) + group * kLanes + lane] = static_cast<int32_t>(inc);
  if (lane == 31) p.idx_total[static_cast<size_t>(b) * p.groups_per_plane + group] = static_cast<int32_t>(inc);
}

// The cross-group part of stage 3 (what the collect_indices passes of
// codec/decode_indices.cl:66-84 do): exclusive scan of the group totals of one image, added to
// the group-local run ends.  One CTA per image.
__global__ void __launch_bounds__(256) index_carry_kernel(const BatchParams p) {
  __shared__ uint32_t s_warp[8];
  __shared__ uint32_t s_base;
  const uint32_t t = threadIdx.x, lane = t & 31, warp = t >> 5;
  int32_t *tot = p.idx_total + static_cast<size_t>(blockIdx.x) * p.groups_per_plane;
  int32_t *run_end = p.run_end + static_cast<size_t>(blockIdx.x) * (p.n_blocks / kSymsPerLane);
  if (t == 0) s_base = 0;
  __syncthreads();
  for (uint32_t i0 = 0; i0 < p.groups_per_plane; i0 += 256) {
    const uint32_t i = i0 + t;
    const uint32_t v = i < p.groups_per_plane ? static_cast<uint32_t>(tot[i]) : 0u;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t n = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += n;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t base = s_base;
    for (uint32_t w = 0; w < warp; ++w) base += s_warp[w];
    const uint32_t carry = base + inc - v;  // exclusive
    if (i < p.groups_per_plane) {
      // the 32 runs of group i
      for (uint32_t l = 0; l < kLanes; ++l) run_end[i * kLanes + l] += static_cast<int32_t>(carry);
    }
    __syncthreads();
    if (t == 255) s_base = base + inc;
    __syncthreads();
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

// Work areas (int16, per warp).  Intermediates are bounded by 128 + 672 per level (<= 3488
// after five levels) for ANY input bytes, so int16 storage is exact.
//   Wlow: the 16x16 corners of two tiles, 16 rows of 32 B each (tile A in the warp's X area,
//         tile B in its Y area); the two 16-byte chunks of row r are swapped when (r >> 2) & 1,
//         which makes lane = row 16-byte accesses and lane = column 2-byte accesses conflict
//         free without padding.
//   W   : one 32x32 tile, 32 rows of 64 B, chunk j of row r stored at j ^ ((r >> 1) & 3);
//         rows 0..15 live in the X area, rows 16..31 overlay the tile's own 1 KiB of the
//         symbol stage (its symbols are all in registers by then).
constexpr int kXBytes = 1024;  // per warp: cp.async ring (phase 1) / Wlow tile A / W rows 0..15
constexpr int kYBytes = 512;   // per warp: Wlow tile B

// levels 2..16 on two tiles at once: lanes 0-15 own tile A, lanes 16-31 tile B.
// wl = this lane's Wlow area (X for lanes 0..15, Y for lanes 16..31).
template <int LEN>
__device__ __forceinline__ void low_level(uint32_t wl, uint32_t lane) {
  const uint32_t rc = lane & 15;  // row in the row pass, column in the column pass
  // rows
  if (rc < LEN) {
    const uint32_t row = wl + rc * 32 + (((rc >> 2) & 1) << 4);  // logical chunk 0
    int v[LEN];
    if (LEN == 16) {
      const uint4 a = lds128(row), b = lds128(row ^ 16);
      const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) { v[2 * i] = lo16(w[i]); v[2 * i + 1] = hi16(w[i]); }
    } else if (LEN == 8) {
      const uint4 a = lds128(row);
      const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) { v[2 * i] = lo16(w[i]); v[2 * i + 1] = hi16(w[i]); }
    } else if (LEN == 4) {
      const uint2 a = lds64(row);
      v[0] = lo16(a.x); v[1] = hi16(a.x); v[2] = lo16(a.y); v[3] = hi16(a.y);
    } else {
      const uint32_t a = lds32(row);
      v[0] = lo16(a); v[1] = hi16(a);
    }
    inverse_lift<LEN>(v);
    if (LEN == 16) {
      sts128(row, pack16(v[0], v[1]), pack16(v[2], v[3]), pack16(v[4], v[5]), pack16(v[6], v[7]));
      sts128(row ^ 16, pack16(v[8], v[9]), pack16(v[10], v[11]), pack16(v[12], v[13]), pack16(v[14], v[15]));
    } else if (LEN == 8) {
      sts128(row, pack16(v[0], v[1]), pack16(v[2], v[3]), pack16(v[4], v[5]), pack16(v[6], v[7]));
    } else if (LEN == 4) {
      sts64(row, pack16(v[0], v[1]), pack16(v[2], v[3]));
    } else {
      sts32(row, pack16(v[0], v[1]));
    }
  }
  __syncwarp();
  // columns
  if (rc < LEN) {
    const uint32_t col[2] = {wl + (((rc >> 3) ^ 0) << 4) + (rc & 7) * 2, wl + (((rc >> 3) ^ 1) << 4) + (rc & 7) * 2};
    int v[LEN];
#pragma unroll
    for (int i = 0; i < LEN; ++i) v[i] = lds_s16(col[(i >> 2) & 1] + i * 32);
    inverse_lift<LEN>(v);
#pragma unroll
    for (int i = 0; i < LEN; ++i) sts16(col[(i >> 2) & 1] + i * 32, static_cast<uint32_t>(v[i]));
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
__device__ __forceinline__ uint32_t pack565(int y, int co, int cg) {  // low 16 bits valid
  int r, g, b;
  ycocg_to_rgb(y, co, cg, r, g, b);
  return (static_cast<uint32_t>(r) << 11) | (static_cast<uint32_t>(g) << 5) | static_cast<uint32_t>(b);
}

// 4 sign-extended bytes of a word
__device__ __forceinline__ void sext4(uint32_t w, int (&o)[4]) {
  o[0] = sext_byte<0>(w); o[1] = sext_byte<1>(w); o[2] = sext_byte<2>(w); o[3] = sext_byte<3>(w);
}

// ---------------------------------------------------------------------------------------
// The fused endpoint-plane kernel.  CTA (b, g) owns tiles [8g, 8g+8) of image b in all six
// planes [Y1,Y2,Co1,Cg1,Co2,Cg2] (codec/assemble.cl:27-37) = one rANS group per plane, and
// WARP w OWNS PLANE w through phases 1 and 2, so the only CTA-wide barrier of the data path is
// the one in front of the assembly:
//   phase 1: the warp rANS-decodes the group of its plane into its 8 KiB symbol stage
//   phase 2: the warp runs the 5-level inverse wavelet on the 8 tiles of that group; levels
//            2..16 on two tiles at a time (lane = tile x row / tile x column), level 32 per tile
//            (lane = row in registers, then lane = column); the int8 result replaces the
//            tile's symbols in the stage, row-major
//   phase 3: the six warps share the 64 four-row slabs of the 8 tiles: 4 DXT1 blocks (or 64
//            RGB8 texels) per lane and slab, coalesced 16-byte stores; the palette index is
//            run_end - S (see side_streams_kernel) and its loads run one slab ahead.
//
// Symbol stage: lane l of the rANS warp owns run l = bytes [256 l, 256 l + 256) of the group
// (ans/ans_decode.cl:71); 8-byte chunk c of run r is stored at chunk position c ^ (r & 15) of
// that run, which keeps the lane-strided 8-byte stores of phase 1 bank-conflict free without
// padding.  Tile t = runs 4t..4t+3 = stage bytes [1024 t, 1024 t + 1024), row-major 32x32.
//
// Shared memory (1 KiB aligned): [6 X areas][2 tables][6 Y areas][6 stages] = 74752 B, three
// CTAs (18 warps) per SM.
constexpr int kFusedWarps = 6;
constexpr int kStageBytes = kGroupSyms;                                   // 8192
constexpr int kFusedTabOff = kFusedWarps * kXBytes;                      // 6144
constexpr int kFusedYOff = kFusedTabOff + 2 * kTableSize * 4;            // 22528
constexpr int kFusedStageOff = kFusedYOff + kFusedWarps * kYBytes;       // 25600
constexpr int kFusedSmem = kFusedStageOff + kFusedWarps * kStageBytes;   // 74752
static_assert(kXBytes == kRing, "the X area doubles as the cp.async ring");
static_assert(kFusedStageOff % 1024 == 0, "tile slots must be 16-byte aligned");

template <int RGB>
__global__ void __launch_bounds__(kFusedWarps * 32, 3) fused_planes_kernel(const BatchParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_s = smem_u32(smem);
  trap_unless_aligned(smem_s, kRing);
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tabs_s = smem_s + kFusedTabOff;
  const uint32_t x_s = smem_s + warp * kXBytes;
  const uint32_t y_s = smem_s + kFusedYOff + warp * kYBytes;
  const uint32_t stage0_s = smem_s + kFusedStageOff;
  const uint32_t stage_s = stage0_s + warp * kStageBytes;  // this warp's plane
  const uint32_t b = blockIdx.x / p.groups_per_plane;
  const uint32_t g = blockIdx.x % p.groups_per_plane;
  const ImageStreams is = image_streams(p, b);

  // Y table and chroma table of this image
  load_table(tabs_s, p.tables + (4ull * b + 0) * kTableSize, threadIdx.x, kFusedWarps * 32);
  load_table(tabs_s + kTableSize * 4, p.tables + (4ull * b + 1) * kTableSize, threadIdx.x, kFusedWarps * 32);
  __syncthreads();

  // ---- phase 1: rANS, warp = plane ------------------------------------------------------
  // Y stream = Y1 || Y2, chroma stream = Co1 || Cg1 || Co2 || Cg2 (codec/encoder.cpp:87,93-95)
  const bool chroma = warp >= 2;
  const uint32_t group = (chroma ? warp - 2 : warp) * p.groups_per_plane + g;
  {
    const uint32_t run_s = stage_s + lane * 256;
    const uint32_t sw = lane & 15;
    rans_decode_group<true>(tabs_s + (chroma ? kTableSize * 4 : 0), is.payload + (chroma ? is.in_off[1] : is.in_off[0]),
                            group, kLanes, x_s, p.cmp, p.cmp + p.cmp_bytes,
                            [&](int m, uint32_t lo, uint32_t hi) { sts64(run_s + (((31 - m) ^ sw) << 3), lo, hi); });
  }
  __syncwarp();

  if (p.tap_symbols) {
    uint8_t *dst = p.tap_symbols + (chroma ? is.out_off[1] : is.out_off[0]) + static_cast<size_t>(group) * kGroupSyms;
    for (uint32_t r = 0; r < kLanes; ++r) {
      const uint2 v = lds64(stage_s + r * 256 + ((lane ^ (r & 15)) << 3));
      *reinterpret_cast<uint2 *>(dst + r * 256 + lane * 8) = v;
    }
    __syncwarp();
  }

  // ---- phase 2: inverse wavelet of the 8 tiles of this plane -------------------------------
  const uint32_t tiles_x = p.blocks_x / kTile;
  const uint32_t tile0 = g * 8;
  const uint32_t ty0 = tile0 / tiles_x, tx0 = tile0 % tiles_x;
  const uint32_t wl = (lane & 16) ? y_s : x_s;  // this lane's Wlow area in the low levels

#pragma unroll 1
  for (uint32_t pair = 0; pair < 4; ++pair) {
    // corners: lane -> (tile 2 pair + lane/16, row lane%16), bytes -> (byte - 128) as int16
    {
      const uint32_t t = 2 * pair + (lane >> 4), r = lane & 15;
      // row r of tile t: run 4t + r/8, chunks 4 (r%8) + j; chunk position = chunk ^ (run & 15)
      const uint32_t a0 = stage_s + (4 * t + (r >> 3)) * 256 + ((4 * ((r & 7) ^ (t & 3)) + (r >> 3)) << 3);
      const uint2 a = lds64(a0), c = lds64(a0 ^ 8);
      const uint32_t x0 = a.x ^ 0x80808080u, x1 = a.y ^ 0x80808080u, x2 = c.x ^ 0x80808080u, x3 = c.y ^ 0x80808080u;
      const uint32_t row = wl + r * 32 + (((r >> 2) & 1) << 4);
      sts128(row, sext_byte_pair<0>(x0), sext_byte_pair<2>(x0), sext_byte_pair<0>(x1), sext_byte_pair<2>(x1));
      sts128(row ^ 16, sext_byte_pair<0>(x2), sext_byte_pair<2>(x2), sext_byte_pair<0>(x3), sext_byte_pair<2>(x3));
    }
    __syncwarp();
    low_level<2>(wl, lane);
    low_level<4>(wl, lane);
    low_level<8>(wl, lane);
    low_level<16>(wl, lane);

#pragma unroll 1
    for (uint32_t q = 0; q < 2; ++q) {
      const uint32_t t = 2 * pair + q;
      const uint32_t slot = stage_s + t * kTileSyms;  // this tile's 1 KiB of the stage
      // ---- level 32, rows: lane = row
      {
        int v[32];
        // symbol chunks j = 0..3 of row `lane` sit at sym0 ^ (8 j)
        const uint32_t sym0 = slot + (lane >> 3) * 256 + ((4 * ((lane & 7) ^ (t & 3)) + (lane >> 3)) << 3);
        if (lane < 16) {  // low half of rows 0..15 = the 16x16 result of the lower levels
          const uint32_t row = (q ? y_s : x_s) + lane * 32 + (((lane >> 2) & 1) << 4);
          const uint4 a = lds128(row), c = lds128(row ^ 16);
          const uint32_t w[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) { v[2 * i] = lo16(w[i]); v[2 * i + 1] = hi16(w[i]); }
        } else {
          const uint2 a = lds64(sym0), c = lds64(sym0 ^ 8);
          const uint32_t x[4] = {a.x ^ 0x80808080u, a.y ^ 0x80808080u, c.x ^ 0x80808080u, c.y ^ 0x80808080u};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            v[4 * i] = sext_byte<0>(x[i]); v[4 * i + 1] = sext_byte<1>(x[i]);
            v[4 * i + 2] = sext_byte<2>(x[i]); v[4 * i + 3] = sext_byte<3>(x[i]);
          }
        }
        {
          const uint2 a = lds64(sym0 ^ 16), c = lds64(sym0 ^ 24);
          const uint32_t x[4] = {a.x ^ 0x80808080u, a.y ^ 0x80808080u, c.x ^ 0x80808080u, c.y ^ 0x80808080u};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            v[16 + 4 * i] = sext_byte<0>(x[i]); v[16 + 4 * i + 1] = sext_byte<1>(x[i]);
            v[16 + 4 * i + 2] = sext_byte<2>(x[i]); v[16 + 4 * i + 3] = sext_byte<3>(x[i]);
          }
        }
        __syncwarp();  // every lane holds its row: the tile's symbols and Wlow may be overwritten
        inverse_lift<32>(v);
        const uint32_t wrow = (lane < 16 ? x_s + lane * 64 : slot + (lane - 16) * 64) + (((lane >> 1) & 3) << 4);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          sts128(wrow ^ (j << 4), pack16(v[8 * j], v[8 * j + 1]), pack16(v[8 * j + 2], v[8 * j + 3]),
                 pack16(v[8 * j + 4], v[8 * j + 5]), pack16(v[8 * j + 6], v[8 * j + 7]));
      }
      __syncwarp();
      // ---- level 32, columns: lane = column; (char) truncation (codec/inverse_wavelet.cl:188-190)
      // back into the tile's slot as row-major bytes
      {
        uint32_t colx[4], cols[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const uint32_t o = ((((lane >> 3) ^ s) & 3) << 4) + (lane & 7) * 2;
          colx[s] = x_s + o;
          cols[s] = slot + o - 16 * 64;
        }
        int v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = lds_s16((i < 16 ? colx[(i >> 1) & 3] : cols[(i >> 1) & 3]) + i * 64);
        __syncwarp();  // W is in registers: the slot may take the result
        inverse_lift<32>(v);
#pragma unroll
        for (int i = 0; i < 32; ++i) sts8(slot + i * 32 + lane, static_cast<uint32_t>(v[i]));
        if (p.tap_planes) {
          uint32_t tx = tx0 + t, ty = ty0;
          while (tx >= tiles_x) { tx -= tiles_x; ++ty; }
          int8_t *tp = p.tap_planes + (static_cast<size_t>(b) * 6 + warp) * p.n_blocks +
                       static_cast<size_t>(ty * kTile) * p.blocks_x + tx * kTile + lane;
#pragma unroll
          for (int i = 0; i < 32; ++i) tp[static_cast<size_t>(i) * p.blocks_x] = static_cast<int8_t>(v[i]);
        }
      }
      __syncwarp();
    }
  }
  __syncthreads();

  // ---- phase 3: assembly, codec/assemble.cl:64-129 -------------------------------------
  const uint32_t n_entries = is.palette_bytes / 4;
  const bool pal_ok = static_cast<uint64_t>(is.pal_off) + is.palette_bytes <= p.palette_cap && n_entries > 0;
  const uint32_t *pal = reinterpret_cast<const uint32_t *>(p.palette + (pal_ok ? is.pal_off : 0));
  const size_t img_block0 = static_cast<size_t>(b) * p.n_blocks;
  const int32_t *run_end = p.run_end + static_cast<size_t>(b) * (p.n_blocks / kSymsPerLane);
  const bool idx16 = p.idx16 != 0;
  const uint32_t lane_row = lane >> 3, lane_col = 4 * (lane & 7);

  // slab u = 8 * tile + k: rows 4k..4k+3 of tile (g * 8 + u / 8); first block of this lane
  auto slab_gidx = [&](uint32_t u) -> uint32_t {
    uint32_t tx = tx0 + (u >> 3), ty = ty0;
    while (tx >= tiles_x) { tx -= tiles_x; ++ty; }
    return (ty * kTile + 4 * (u & 7) + lane_row) * p.blocks_x + tx * kTile + lane_col;
  };
  // stage A: suffix sums + run end of the 4 blocks;  stage B: indices -> palette words
  uint32_t sfx[4] = {0u, 0u, 0u, 0u}, re = 0u, word_nx[4] = {0u, 0u, 0u, 0u};
  auto load_sfx = [&](uint32_t gidx) {
    if (idx16) {
      const uint2 sv = __ldg(reinterpret_cast<const uint2 *>(reinterpret_cast<const uint16_t *>(p.idx_s) + img_block0 + gidx));
      sfx[0] = sv.x & 0xFFFFu; sfx[1] = sv.x >> 16; sfx[2] = sv.y & 0xFFFFu; sfx[3] = sv.y >> 16;
    } else {
      const uint4 sv = __ldg(reinterpret_cast<const uint4 *>(reinterpret_cast<const uint32_t *>(p.idx_s) + img_block0 + gidx));
      sfx[0] = sv.x; sfx[1] = sv.y; sfx[2] = sv.z; sfx[3] = sv.w;
    }
    re = static_cast<uint32_t>(__ldg(run_end + gidx / kSymsPerLane));
  };
  auto load_words = [&](uint32_t gidx) {
    uint32_t idx[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      idx[j] = re - sfx[j];
      if (idx16) idx[j] &= 0xFFFFu;
      word_nx[j] = pal_ok ? __ldg(pal + min(idx[j], n_entries - 1)) : 0u;
    }
    if (p.tap_indices)
      *reinterpret_cast<uint4 *>(p.tap_indices + img_block0 + gidx) = make_uint4(idx[0], idx[1], idx[2], idx[3]);
  };

  uint32_t gidx_b = slab_gidx(warp);  // gidx of the slab whose words are being loaded
  load_sfx(gidx_b);
  load_words(gidx_b);
  if (warp + kFusedWarps < 64) {
    gidx_b = slab_gidx(warp + kFusedWarps);
    load_sfx(gidx_b);
  }

#pragma unroll 1
  for (uint32_t u = warp; u < 64; u += kFusedWarps) {
    const uint32_t gidx = slab_gidx(u);
    uint32_t word[4] = {word_nx[0], word_nx[1], word_nx[2], word_nx[3]};
    if (u + kFusedWarps < 64) load_words(gidx_b);  // sfx / re of slab u + 6 arrived during slab u - 6
    if (u + 2 * kFusedWarps < 64) {
      gidx_b = slab_gidx(u + 2 * kFusedWarps);
      load_sfx(gidx_b);
    }
    const uint32_t src = stage0_s + u * 128 + lane * 4;  // rows 4k..4k+3 of the tile, 4 bytes per lane
    uint32_t pw[6];
#pragma unroll
    for (int pl = 0; pl < 6; ++pl) pw[pl] = lds32(src + pl * kStageBytes);

    int y1[4], y2[4], co1[4], cg1[4], co2[4], cg2[4];
    sext4(pw[0], y1); sext4(pw[1], y2); sext4(pw[2], co1); sext4(pw[3], cg1); sext4(pw[4], co2); sext4(pw[5], cg2);

    if (!RGB) {
      uint32_t o[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // PhysicalDXTBlock (codec/dxt_image.h:14-21): u16 ep1, u16 ep2, u32 interpolation
        o[2 * j] = __byte_perm(pack565(y1[j], co1[j], cg1[j]), pack565(y2[j], co2[j], cg2[j]), 0x5410);
        o[2 * j + 1] = word[j];
      }
      uint8_t *dst = p.out + (img_block0 + gidx) * 8;
      st_global_cs_v4(dst, o[0], o[1], o[2], o[3]);
      st_global_cs_v4(dst + 16, o[4], o[5], o[6], o[7]);
    } else {
      // assemble_rgb: 565 -> 888 by bit replication, always the 4-colour palette
      // (codec/assemble.cl:102-111); 16 texels per block, raster RGB8 (:117-128)
      uint32_t pal4[4][4];  // [block j][palette entry] = r | g << 8 | b << 16 (uchar-truncated)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int c0[3], c1[3];
        ycocg_to_rgb(y1[j], co1[j], cg1[j], c0[0], c0[1], c0[2]);
        ycocg_to_rgb(y2[j], co2[j], cg2[j], c1[0], c1[1], c1[2]);
        c0[0] = static_cast<int>((static_cast<uint32_t>(c0[0]) << 3) | static_cast<uint32_t>(c0[0] >> 2));
        c0[1] = static_cast<int>((static_cast<uint32_t>(c0[1]) << 2) | static_cast<uint32_t>(c0[1] >> 4));
        c0[2] = static_cast<int>((static_cast<uint32_t>(c0[2]) << 3) | static_cast<uint32_t>(c0[2] >> 2));
        c1[0] = static_cast<int>((static_cast<uint32_t>(c1[0]) << 3) | static_cast<uint32_t>(c1[0] >> 2));
        c1[1] = static_cast<int>((static_cast<uint32_t>(c1[1]) << 2) | static_cast<uint32_t>(c1[1] >> 4));
        c1[2] = static_cast<int>((static_cast<uint32_t>(c1[2]) << 3) | static_cast<uint32_t>(c1[2] >> 2));
        uint32_t e[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          e[0] |= (static_cast<uint32_t>(c0[c]) & 0xFFu) << (8 * c);
          e[1] |= (static_cast<uint32_t>(c1[c]) & 0xFFu) << (8 * c);
          e[2] |= (static_cast<uint32_t>((2 * c0[c] + c1[c]) / 3) & 0xFFu) << (8 * c);
          e[3] |= (static_cast<uint32_t>((c0[c] + 2 * c1[c]) / 3) & 0xFFu) << (8 * c);
        }
#pragma unroll
        for (int s = 0; s < 4; ++s) pal4[j][s] = e[s];
      }
      // texel coordinates of this lane's first block
      const size_t img_w = 4ull * p.blocks_x;
      uint8_t *img = p.out + static_cast<size_t>(b) * p.n_blocks * 48;
      const uint32_t by = gidx / p.blocks_x, bx = gidx - by * p.blocks_x;
      const size_t x0 = 4ull * bx, y0 = 4ull * by;
#pragma unroll
      for (int yy = 0; yy < 4; ++yy) {
        uint32_t t[16];  // 16 texels of this texel row, r | g << 8 | b << 16
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
          for (int xx = 0; xx < 4; ++xx) {
            const uint32_t sel = (word[j] >> (2 * (4 * yy + xx))) & 3u;
            const uint32_t lo = (sel & 1u) ? pal4[j][1] : pal4[j][0];
            const uint32_t hi = (sel & 1u) ? pal4[j][3] : pal4[j][2];
            t[4 * j + xx] = (sel & 2u) ? hi : lo;
          }
        }
        uint32_t wds[12];  // 48 bytes: 16 x RGB
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          wds[3 * i + 0] = t[4 * i] | (t[4 * i + 1] << 24);
          wds[3 * i + 1] = (t[4 * i + 1] >> 8) | (t[4 * i + 2] << 16);
          wds[3 * i + 2] = (t[4 * i + 2] >> 16) | (t[4 * i + 3] << 8);
        }
        uint8_t *dst = img + 3 * (img_w * (y0 + yy) + x0);
        st_global_cs_v4(dst, wds[0], wds[1], wds[2], wds[3]);
        st_global_cs_v4(dst + 16, wds[4], wds[5], wds[6], wds[7]);
        st_global_cs_v4(dst + 32, wds[8], wds[9], wds[10], wds[11]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Standalone decode of [u32 end_offset[n_groups]][groups] with 1..32 interleaved lanes and a
// single table: the `ans_decode` kernel of ans/ans_decode.cl:76-95 as driven by
// ans/ans_ocl.cpp:159-345.  Output: group * n_lanes * 256 + lane * 256 + position.
constexpr int kPlainWarps = 8;
constexpr int kPlainSmem = kPlainWarps * kRing + kTableSize * 4;

__global__ void __launch_bounds__(kPlainWarps * 32)
    ans_decode_plain_kernel(const uint32_t *__restrict__ table, const uint8_t *__restrict__ data,
                            uint64_t data_bytes, uint32_t n_groups, uint32_t n_lanes,
                            uint8_t *__restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_s = smem_u32(smem);
  trap_unless_aligned(smem_s, kRing);
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tab_s = smem_s + kPlainWarps * kRing;
  load_table(tab_s, table, threadIdx.x, kPlainWarps * 32);
  __syncthreads();
  const uint32_t group = blockIdx.x * kPlainWarps + warp;
  if (group >= n_groups) return;
  uint8_t *dst = out + (static_cast<size_t>(group) * n_lanes + lane) * kSymsPerLane + 248;
  const bool active = lane < n_lanes;
  rans_decode_group<false>(tab_s, data, group, n_lanes, smem_s + warp * kRing, data, data + data_bytes,
                           [&](int m, uint32_t lo, uint32_t hi) {
                             if (active) *reinterpret_cast<uint2 *>(dst - 8 * m) = make_uint2(lo, hi);
                           });
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
  index_carry_kernel<<<p.n_images, 256, 0, s>>>(p);
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
