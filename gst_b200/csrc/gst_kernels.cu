// gst_b200 -- sm_100a kernels for the .gst -> DXT1 decode path.
//
// Kernel inventory (reference stage it replaces, citations relative to the reference tree):
//   build_tables_kernel,     stage 1  ans/build_table.cl:12-83 (flat array of frequency blocks / the blocks of a batch)
//   build_tables_batch_kernel
//   rans_streams_kernel      stage 2  ans/ans_decode.cl:25-143 for all four streams of every image,
//                            one warp per two 32-stream groups.  Plane symbols go to a transposed,
//                            plane-pair-interleaved scratch, palette symbols to the compact palette,
//                            and for the index stream stage 3 (codec/decode_indices.cl:6-84, host loop
//                            codec/decoder.cpp:311-393) is fused behind the rANS warp as a group-local
//                            suffix sum plus a per-group total.  <FT>: stage 1 as well -- the CTA builds its
//                            table in shared memory (calls too small to fill the machine)
//   wavelet_assemble_kernel  stage 4 (codec/inverse_wavelet.cl:69-192) and stage 5
//                            (codec/assemble.cl:64-129): one warp per 32x32 tile, all six planes as
//                            three packed plane pairs; the wavelet planes never leave shared memory; the
//                            cross-group carry of stage 3 is summed here from the group totals
//   ans_encode_kernel,       the entropy stage of the ENCODER (codec/entropy.cpp:174-265), fixture tooling
//   ans_encode_gather_kernel
//   ans_decode_plain_kernel  the standalone `ans_decode` entry (ans/ans_decode.cl:76-95),
//                            1..32 interleaved lanes, used by the OpenCLDecoder-style API
//
// All arithmetic is integer and follows the reference bit for bit: wrapping u32 rANS
// state, C truncating division in the 5/3 lifting and in YCoCg->RGB, (char) truncation of
// the wavelet output, unmasked shift/or 565 pack.
#include <type_traits>

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
// Programmatic dependent launch: the next kernel of the stream may be scheduled once every CTA of this one has
// passed pdl_launch_dependents().  pdl_wait() blocks until this whole kernel has completed and its writes are visible;
// the tile kernels of multi-image calls use the finer, per-image wait below instead (wait_for_image).
// Both are no-ops for kernels launched without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Hand-over of an image from rans_streams_kernel to wavelet_assemble_kernel.  The second kernel is launched as a
// programmatic dependent of the first, so its CTAs become resident as soon as EVERY decode CTA has started (they all
// pass pdl_launch_dependents() first thing) and SMs have room -- that is, during the last, partly filled wave of the
// decode kernel.  Instead of sleeping in griddepcontrol.wait until the whole decode grid has drained, a tile warp
// waits for its own image only: every decode CTA of image b releases its stores with a fence and counts itself into
// img_done[b]; the tile warp acquires the counter.  A warp can only ever wait for CTAs that are already running, so the
// wait cannot deadlock; it is bounded all the same (GST_FLAG_SYNC_TIMEOUT) so that a broken launch order shows up
// as a flag and a failed parity check, not as a hung GPU.  The acquire makes everything the image's decode CTAs
// wrote visible; beyond that, what the tile warps read is either fetched past L1 (cp.async.cg) or lies in cache
// lines that hold data of their own image only (symbols, suffix sums, run ends and palettes are multiples of 128
// bytes per image, the group totals are padded to that: idx_total_stride).
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void wait_for_image(const uint32_t *done, uint32_t target, uint32_t *status) {
  // The whole warp probes (one request: all lanes read the same word) and the vote keeps the loop warp-uniform,
  // so that ptxas does not duplicate the code that follows for a divergent lane 0.  An acquire LOAD is a strong load
  // plus an invalidation of the SM's L1 (LDG.E.STRONG.GPU + CCTL.IVALL); a relaxed probe followed by
  // fence.acq_rel.gpu would add a MEMBAR.ALL.GPU, on which the tile warps were found waiting 10 % of their time.
  uint32_t spins = 0;
#pragma unroll 1
  while (!__all_sync(0xffffffffu, ld_acquire_gpu(done) >= target)) {
    __nanosleep(100);
    if (++spins > (1u << 23)) {  // > 1 s
      atomicOr(status, 4u);
      break;
    }
  }
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
// signed byte j of w as int; byte pair (j, j+1) as a packed int16x2 word.  The scalar
// extractions are dot products with a one-hot selector (IDP.4A / IDP.2A): they issue on the
// FMA pipe, which idles in the wavelet, instead of PRMT / SHF on the saturated ALU pipe.
template <int J>
__device__ __forceinline__ int sext_byte(uint32_t w) {
  int d;
  asm("dp4a.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(1u << (8 * J)), "r"(0));
  return d;
}
template <int J>
__device__ __forceinline__ uint32_t sext_byte_pair(uint32_t w) {
  return prmt<((8 | (J + 1)) << 12) | ((J + 1) << 8) | ((8 | J) << 4) | J>(w, 0u);
}

// ---------------------------------------------------------------------------------------
// Stage 2 core: one warp decodes one group of `n_lanes` interleaved rANS streams.
//
// Stream layout (codec/entropy.cpp:199-262): stream = [u32 end_offset[groups]][group 0]...;
// a group ends with its n_lanes 32-bit states, preceded by the shared 16-bit renorm words,
// which are consumed backwards, higher lanes first (ans/ans_decode.cl:30-32,51-65).
//
// The renorm words are staged through a per-chain shared-memory ring of three 512-byte slots filled with
// cp.async, one 512-byte-aligned chunk of the stream per slot (one 16-byte copy per lane; group ranges are
// only 4-byte aligned, so the windows are aligned down in absolute address space).  Going down the stream
// by 512 bytes goes down one slot, from slot 0 back to slot 2.  A checkpoint every 8 symbols (which
// consume at most 8*32*2 = 512 B) stages the next chunk whenever fewer than 1024 B are staged and then
// waits until only that newest copy is pending: at least 512 complete bytes lie below the next word, which
// is all the next eight symbols can touch, and a chunk has a whole interval (thousands of cycles) to land.
//
// Inside an interval the word address is NOT wrapped: below slot 0 sit 512 bytes that mirror slot 2
// (whatever is staged into slot 2 is staged there too), so an address that runs off the bottom of the
// ring during the eight symbols still finds its word, and the decode step needs no instruction to fold
// the address back -- the checkpoint does it (+1536 when it has left the ring).
//
// Per symbol and lane (ans/ans_decode.cl:38-65):
//   e = table[state & 2047];  state = (state >> 11) * e.freq + slot - e.cum   (entry layout: gst_kernels.cuh)
//   lanes whose state fell below L = 2^15 take the next 16-bit word, higher lanes first:
//   word index = next - 1 - popc(ballot & lanes_above_me);  next -= popc(ballot)
// The word load is unconditional (every lane's address lies inside staged bytes), only
// the state update is predicated.
constexpr int kChunk = 512;
constexpr int kRing = 3 * kChunk;          // the ring proper
constexpr int kRingSlot = kRing + kChunk;  // ring + the mirror below it

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// NC = 1 or 2 groups per warp.  With NC = 2 the warp decodes groups grp[0] and grp[1] of the
// same stream in one interleaved instruction stream: the decode step is a chain of ~12 dependent
// instructions (two of them shared-memory loads), and two independent chains per warp hide
// that latency better than twice the warps would (registers, not warps, are what is left).
// Chain c uses the ring at ring[c].
// emit(m, acc): called 16 times; acc holds the 16 symbols at positions q0 = 240 - 16m .. q0 + 15 of
// this lane's 256-symbol run of every chain:
//   ILV = false: acc[4c + j] = symbols q0 + 4j .. q0 + 4j + 3 of chain c, little-endian packed
//   ILV = true (NC = 2): the two chains byte-interleaved, acc[j] = {chain 0 @ q0 + 2j, chain 1 @ q0 + 2j,
//                chain 0 @ q0 + 2j + 1, chain 1 @ q0 + 2j + 1} -- the layout wavelet_assemble_kernel
//                unpacks with one PRMT per coefficient pair; it costs nothing here because the
//                PRMT that files a symbol away takes any destination byte.
template <bool FULL, int NC, bool ILV, class Emit>
__device__ __forceinline__ void rans_decode_groups(uint32_t tab_s, const uint8_t *__restrict__ stream,
                                                   const uint32_t (&grp)[NC], uint32_t n_lanes, const uint32_t (&ring)[NC],
                                                   const uint8_t *buf_lo, const uint8_t *buf_hi, Emit emit) {
  static_assert(!ILV || NC == 2, "interleaved output needs two chains");
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t gt = lanemask_gt();
  const bool active = FULL || lane < n_lanes;
  const uintptr_t lo_ok = reinterpret_cast<uintptr_t>(buf_lo) + 4 * n_lanes;
  const uintptr_t hi_ok = reinterpret_cast<uintptr_t>(buf_hi) & ~static_cast<uintptr_t>(3);
  const uintptr_t lo16 = (reinterpret_cast<uintptr_t>(buf_lo) + 15) & ~static_cast<uintptr_t>(15);
  const uintptr_t hi16 = reinterpret_cast<uintptr_t>(buf_hi) & ~static_cast<uintptr_t>(15);

  // one 512-byte chunk of the stream (global address `chunk`, 512-aligned) into the ring slot at shared address
  // `dst` of chain c, 16 bytes per lane; slot 2 goes to the mirror below the ring as well
  auto stage = [&](int c, uint32_t dst, uintptr_t chunk) {
    const uintptr_t a = chunk + 16 * lane;
    if (a >= lo16 && a + 16 <= hi16) {
      cp_async16(dst + 16 * lane, reinterpret_cast<const void *>(a));
      if (dst == ring[c] + 2 * kChunk) cp_async16(ring[c] - kChunk + 16 * lane, reinterpret_cast<const void *>(a));
    }
  };
  uint32_t state[NC], pos[NC], mprev[NC], lo_s[NC];  // lo_s: shared address of the slot holding the lowest staged chunk
  uintptr_t lo[NC];                                  // lo: stream address of that chunk
  uint32_t end[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    // ans/ans_decode.cl:30.  The offset table of the stream is caller data too: with inconsistent stream offsets
    // its address can lie outside the buffer, so it is clamped like everything derived from it
    uintptr_t a = reinterpret_cast<uintptr_t>(stream) + 4ull * grp[c];
    a = a < reinterpret_cast<uintptr_t>(buf_lo) ? reinterpret_cast<uintptr_t>(buf_lo) : a;
    a = a + 4 > hi_ok ? hi_ok - 4 : a;
    end[c] = __ldg(reinterpret_cast<const uint32_t *>(a & ~static_cast<uintptr_t>(3))) & ~3u;
  }
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    // Clamp so a malformed offset can never leave [buf_lo, buf_hi).
    uintptr_t top = reinterpret_cast<uintptr_t>(stream) + end[c];
    top = top < lo_ok ? lo_ok : top;
    top = top > hi_ok ? hi_ok : top;
    const uintptr_t a_pos = top - 4 * n_lanes;  // one past the last renorm word
    // ans/ans_decode.cl:31
    state[c] = active ? __ldg(reinterpret_cast<const uint32_t *>(a_pos) + lane) : 0u;
    // preload [c_top - kRing, c_top), c_top = a_pos rounded up to kChunk: the chunk with the first word in slot 2
    const uintptr_t c_top = (a_pos + (kChunk - 1)) & ~static_cast<uintptr_t>(kChunk - 1);
    lo[c] = c_top - kRing;
    lo_s[c] = ring[c];
#pragma unroll
    for (int i = 0; i < kRing / kChunk; ++i) stage(c, ring[c] + kChunk * i, lo[c] + kChunk * i);
    pos[c] = ring[c] + kRing - 2u - static_cast<uint32_t>(c_top - a_pos);  // shared-memory address of the next word
    mprev[c] = 0u;
  }
  cp_async_commit();
  cp_async_wait_group<0>();
  __syncwarp();

  // Word addressing with ONE popcount per symbol.  The reference gives lane l the word at
  //   next - 1 - popc(mask_t & lanes_above_l),  next -= popc(mask_t)     (ans/ans_decode.cl:51-65).
  // Here every lane keeps its own address pos_l(t) = next(t) - popc(mask_t & gt_l) (in bytes): then
  //   pos_l(t+1) = pos_l(t) - popc(mask_t & ~gt_l) - popc(mask_{t+1} & gt_l)
  //              = pos_l(t) - popc((mask_t & ~gt_l) | (mask_{t+1} & gt_l))     (disjoint bit ranges)
  // which is one LOP3 + one POPC instead of two of each.  At a checkpoint the uniform `next` is
  // recovered as pos_l - popc(mask & ~gt_l) and the recurrence restarts with mask = 0.
#pragma unroll 1
  for (int m = 0; m < 16; ++m) {
    uint32_t acc[NC * 4];
#pragma unroll
    for (int j = 0; j < NC * 4; ++j) acc[j] = 0u;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      // checkpoint: stage the next chunk when fewer than two are left, then let only that newest copy fly.
      // The chunk that is topped up replaces words other lanes read during the last 8 symbols.
      __syncwarp();
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        pos[c] -= 2u * __popc(mprev[c] & ~gt);  // uniform again: address of the next unread word
        mprev[c] = 0u;
        if (pos[c] < ring[c]) pos[c] += kRing;  // it ran into the mirror during the last interval: back to slot 2
        // bytes staged at and below the next word, in 1..kRing (pos and lo_s both live in [ring, ring + kRing))
        int32_t staged = static_cast<int32_t>(pos[c] + 2u - lo_s[c]);
        if (staged <= 0) staged += kRing;
        if (staged < 2 * kChunk) {
          lo[c] -= kChunk;
          lo_s[c] = lo_s[c] == ring[c] ? ring[c] + 2 * kChunk : lo_s[c] - kChunk;
          stage(c, lo_s[c], lo[c]);
        }
      }
      cp_async_commit();
      cp_async_wait_group<1>();
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 8; ++k) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          // tab_s | ((state << 2) & 0x1FFC): the shift as a multiply on the FMA pipe, one LOP3 (the table is
          // 8 KiB-aligned in the shared window) -- the mask-then-scale form costs two ALU-pipe instructions
          uint32_t s4;
          asm("mul.lo.u32 %0, %1, 4;" : "=r"(s4) : "r"(state[c]));
          uint32_t slot_a;
          asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(slot_a) : "r"(s4), "n"(4 * kTableSize - 4), "r"(tab_s));
          const uint32_t e = lds32(slot_a);
          // state' = (state >> 11) * freq + slot - cum (gst_kernels.cuh); the symbol shift is a multiply so that it
          // issues on the FMA pipe
          uint32_t sym24;
          asm("mul.lo.u32 %0, %1, 4096;" : "=r"(sym24) : "r"(e));    // e << 12: symbol in the top byte
          state[c] = (state[c] >> 11) * (e & 0xFFFu) + (e >> 20);
          const bool need = FULL ? (state[c] < kRansL) : (active && state[c] < kRansL);
          const uint32_t mask = __ballot_sync(0xffffffffu, need);
          uint32_t sel;  // (mask & gt) | (mprev & ~gt)
          asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(sel) : "r"(mask), "r"(mprev[c]), "r"(gt));
          asm("mad.lo.u32 %0, %1, 0xfffffffe, %0;" : "+r"(pos[c]) : "r"(__popc(sel)));  // pos -= 2 popc, on the FMA pipe
          mprev[c] = mask;
          const uint32_t w = lds_u16(pos[c]);
          uint32_t renorm;  // state << 16 | w
          asm("mad.lo.u32 %0, %1, 65536, %2;" : "=r"(renorm) : "r"(state[c]), "r"(w));
          if (need) state[c] = renorm;
          const int q = 15 - 8 * h - k;  // position inside this 16-symbol piece (symbols arrive last first)
          if (ILV) {
            // byte 2 (q & 1) + c of word q / 2 <- symbol
            constexpr uint32_t kIns[4] = {0x3217u, 0x3270u, 0x3710u, 0x7210u};
            const int byte = 2 * (q & 1) + c;
            uint32_t &a = acc[q >> 1];
            a = byte == 0 ? __byte_perm(a, sym24, kIns[0]) : byte == 1 ? __byte_perm(a, sym24, kIns[1])
              : byte == 2 ? __byte_perm(a, sym24, kIns[2]) : __byte_perm(a, sym24, kIns[3]);
          } else {
            acc[4 * c + (q >> 2)] = __byte_perm(acc[4 * c + (q >> 2)], sym24, 0x2107);  // acc << 8 | symbol
          }
        }
      }
    }
    emit(m, acc);
  }
  cp_async_wait_group<0>();
}

// ---------------------------------------------------------------------------------------
// Stage 1.  ans/build_table.cl:12-83: 256 frequencies -> for every slot in [0, 2048) the
// symbol x with cum[x] <= slot < cum[x+1].  The reference scans then binary-searches per slot;
// here every symbol with a non-zero frequency drops its id at slot cum[x] and a max-scan
// spreads it -- the result is determined by the frequencies alone, so it is identical.
struct TableScratch {
  uint32_t freq[256], cum[256], warp[8];
};
// 256 threads: frequencies f16[256] -> 2048 packed entries at `out` (global, or shared = `work`).  `work`: 2048 words
// of shared memory.
__device__ __forceinline__ void build_table_cta(const uint16_t *__restrict__ f16, uint32_t *work, uint32_t *out, TableScratch &ts) {
  const uint32_t t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const uint32_t f = f16[t];
  // exclusive scan of the 256 frequencies
  uint32_t inc = f;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t n = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += n;
  }
  if (lane == 31) ts.warp[warp] = inc;
  for (int i = t; i < kTableSize; i += 256) work[i] = 0;
  __syncthreads();
  uint32_t base = 0;
  for (uint32_t w = 0; w < warp; ++w) base += ts.warp[w];
  const uint32_t cum = (base + inc - f) & 0xFFFFu;  // ushort arithmetic, build_table.cl:14
  ts.freq[t] = f;
  ts.cum[t] = cum;
  if (f != 0 && cum < kTableSize) work[cum] = t;
  __syncthreads();

  // inclusive max-scan over the 2048 slots, 8 consecutive slots per thread
  uint32_t v[8];
  uint32_t run = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    run = max(run, work[t * 8 + i]);
    v[i] = run;
  }
  uint32_t wmax = run;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t n = __shfl_up_sync(0xffffffffu, wmax, d);
    if (lane >= d) wmax = max(wmax, n);
  }
  __syncthreads();
  if (lane == 31) ts.warp[warp] = wmax;
  __syncthreads();
  uint32_t prev = __shfl_up_sync(0xffffffffu, wmax, 1);
  if (lane == 0) prev = 0;
  for (uint32_t w = 0; w < warp; ++w) prev = max(prev, ts.warp[w]);
  // (every thread has read its eight slots: `out` may be `work`)
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t slot = t * 8 + i;
    const uint32_t sym = max(v[i], prev);
    out[slot] = pack_entry(sym, ts.freq[sym], slot, ts.cum[sym]);
  }
}

__global__ void __launch_bounds__(256) build_tables_kernel(const uint8_t *__restrict__ freqs, uint32_t *__restrict__ tables) {
  pdl_launch_dependents();
  __shared__ uint32_t s_sym[kTableSize];
  __shared__ TableScratch ts;
  build_table_cta(reinterpret_cast<const uint16_t *>(freqs + 512ull * blockIdx.x), s_sym,
                  tables + static_cast<size_t>(kTableSize) * blockIdx.x, ts);
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
  // (a single-image call may carry its eight offsets in the kernel parameters instead of the buffer: the frame
  // streamer uploads a file as it lies and writes no offset table)
  const uint4 oo = p.inline_off ? make_uint4(p.off8[0], p.off8[1], p.off8[2], p.off8[3]) : __ldg(tbl + b);
  const uint4 io = p.inline_off ? make_uint4(p.off8[4], p.off8[5], p.off8[6], p.off8[7]) : __ldg(tbl + p.n_images + b);
  // standard layout: all frequency blocks, then all payloads.  freq_inline: every image's 2048 bytes of frequency
  // blocks sit directly in front of its Y stream, as in the .gst file (in_off counts from the end of the offsets
  // region), so a file is ONE copy
  s.payload = p.cmp + p.off_region + (p.freq_inline ? 0ull : 2048ull * p.n_images);
  s.out_off[0] = oo.x; s.out_off[1] = oo.y; s.out_off[2] = oo.z; s.out_off[3] = oo.w;
  s.in_off[0] = io.x;  s.in_off[1] = io.y;  s.in_off[2] = io.z;  s.in_off[3] = io.w;
  s.palette_bytes = oo.w - oo.z;
  s.pal_off = oo.z - 7u * p.n_blocks * b - 6u * p.n_blocks;
  return s;
}

// the 512-byte frequency block of stream `type` of image b (in_off_y = in_off[4b])
__device__ __forceinline__ const uint8_t *freq_block(const BatchParams &p, uint32_t b, uint32_t type, uint32_t in_off_y) {
  return p.freq_inline ? p.cmp + p.off_region + in_off_y - 2048u + 512u * type
                       : p.cmp + p.off_region + 2048ull * b + 512u * type;
}

// stage 1 for a batch: table 4b + type from the frequency block of that stream, wherever the layout puts it
__global__ void __launch_bounds__(256) build_tables_batch_kernel(const BatchParams p) {
  pdl_launch_dependents();
  __shared__ uint32_t s_sym[kTableSize];
  __shared__ TableScratch ts;
  const uint32_t b = blockIdx.x >> 2, type = blockIdx.x & 3;
  const uint32_t in_off_y = p.inline_off ? p.off8[4] : __ldg(reinterpret_cast<const uint32_t *>(p.cmp) + 4 * (p.n_images + b));
  build_table_cta(reinterpret_cast<const uint16_t *>(freq_block(p, b, type, in_off_y)), s_sym,
                  p.tables + static_cast<size_t>(kTableSize) * blockIdx.x, ts);
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
// Stage 2 for every stream of every image.  One warp per TWO rANS groups (two interleaved chains), 8
// warps per CTA, all warps of a CTA on the same stream of the same image (one 8 KiB table in shared
// memory); 42 KiB of shared memory and 48 registers per thread, so an SM holds 5 CTAs = 40 warps = 80
// independent rANS chains.  The decode loop is a chain of dependent shared-memory lookups (~150 cycles
// per symbol); with this many chains the kernel is bound by the mix of ALU pipe, issue slots and
// shared-memory wavefronts (each 65-72 % busy, profiles/), not by that latency.
//
//   Y / chroma groups: the symbols are the wavelet coefficients of the six endpoint planes
//       (Y = Y1 || Y2, chroma = Co1 || Cg1 || Co2 || Cg2, codec/encoder.cpp:87,93-95), one group =
//       8 tiles of one plane.  A warp takes the same group of the two planes of a pair (Y1|Y2, Co1|Co2,
//       Cg1|Cg2) as its two chains and writes them byte-interleaved to the transposed scratch sym_t:
//       image b, pair-group pg = pair * groups_per_plane + g holds [k = 0..15][lane = 0..31][32 B], the
//       32 bytes being positions 16k..16k+15 of the lane's run as (plane A, plane B) pairs.  One warp
//       store is 1 KiB contiguous (the reference layout would be 32 pieces 256 B apart);
//       wavelet_assemble_kernel reads 16-byte pieces, which is all it needs.
//   palette groups: symbols go straight to the compact palette scratch (they are the u32 DXT
//       index words, codec/encoder.cpp:100-108)
//   index groups: symbols are (delta + 128) per DXT block in raster order
//       (codec/dxt_image.cpp:610-618) and every lane owns 256 consecutive blocks (a "run").
//       Stage 3 (codec/decode_indices.cl:24, idx[i] = sum_{j<=i} d[j]) is fused behind the
//       decoder: symbols arrive last-to-first, so the lane writes S[i] = MINUS the sum of the deltas AFTER
//       i inside its run, and idx[i] = run_end[run] + S[i], where run_end is the inclusive
//       prefix at the end of the run (group-local here, the carry of the earlier groups is added by
//       wavelet_assemble_kernel).  S is stored mod 2^16
//       when every palette of the batch has <= 65536 entries (idx < 2^16 then makes the 16-bit
//       difference exact), else as 32 bits, in the same transposed order as sym_t:
//       [group][k][lane][16 values].
// Shared-memory layout of a decode CTA: n rings (1.5 KiB + 512 B of mirror below each) and the 8 KiB table.  The
// table is indexed by OR-ing low address bits into its base, so it sits on an 8 KiB boundary; the dynamic shared
// window is only guaranteed 1 KiB alignment, and the rings (which need 16-byte alignment only) fill the space
// before and after the table: n * 2 KiB + 8 KiB + less than 2 KiB of slack.
struct RansSmem {
  uint32_t s0, tab, n_before;
  __device__ __forceinline__ explicit RansSmem(uint32_t base) {
    s0 = base;
    tab = (s0 + 4 * kTableSize - 1) & ~static_cast<uint32_t>(4 * kTableSize - 1);
    n_before = (tab - s0) / kRingSlot;
  }
  __device__ __forceinline__ uint32_t ring(uint32_t i) const {  // address of the ring proper; its mirror lies below
    return (i < n_before ? s0 + i * kRingSlot : tab + 4 * kTableSize + (i - n_before) * kRingSlot) + kChunk;
  }
};

constexpr int kRansWarps = 8;
// Two rANS groups (chains) per warp in one instruction stream, 5 CTAs = 40 warps = 80 chains per SM.  Measured
// alternatives: four chains per warp at 3 CTAs (96 chains) 10 % slower, 4 CTAs of two chains 2.4 % slower -- the
// kernel wants warps; one chain per warp for calls that leave SMs idle: no faster (a decode step of a lone warp
// takes ~300 cycles with one chain or two), and the byte-granular stores of a lone plane cost more than they save.
struct RansCfg {
  static constexpr int kChains = 2;
  static constexpr int kGroupsPerCta = kRansWarps * kChains;
  static constexpr int kSmem = kGroupsPerCta * kRingSlot + kTableSize * 4 + kRingSlot;  // + alignment slack
  static constexpr int kCtasPerSm = 5;
};

// CTAs of one image: [Y][chroma][palette][index]
struct StreamGrid {
  uint32_t y_ctas, c_ctas, pal_ctas, idx_ctas;
  __host__ __device__ uint32_t per_image() const { return y_ctas + c_ctas + pal_ctas + idx_ctas; }
};

// Group index inside the Y or chroma stream of plane `half` (0 = A, 1 = B) of plane pair `pair`, group g.
// pair 0 = (Y1, Y2), pair 1 = (Co1, Co2), pair 2 = (Cg1, Cg2), i.e. the two endpoints' planes of one colour
// channel, which the wavelet and the assembly process as the two 16-bit halves of one register.  Stream order
// (codec/encoder.cpp:87,93-95): Y = Y1 || Y2, chroma = Co1 || Cg1 || Co2 || Cg2, every plane groups_per_plane long.
__device__ __forceinline__ uint32_t plane_group(uint32_t pair, uint32_t half, uint32_t g, uint32_t gpp) {
  if (half == 0) return pair == 2 ? gpp + g : g;
  return pair == 0 ? gpp + g : pair == 1 ? 2 * gpp + g : 3 * gpp + g;
}
__device__ __forceinline__ uint8_t *pair_group_base(const BatchParams &p, uint32_t b, uint32_t pair, uint32_t g) {
  return p.sym_t + static_cast<size_t>(b) * 6 * p.n_blocks + (static_cast<size_t>(pair) * p.groups_per_plane + g) * (2 * kGroupSyms);
}

// One warp's share of the Y or chroma stream, two chains: the same group of the two planes of a pair.
template <bool TAP>
__device__ __forceinline__ void rans_plane_pair(const BatchParams &p, uint32_t b, uint32_t type, uint32_t item,
                                                const uint8_t *stream, uint32_t out_off, uint32_t tab_s, const uint32_t (&ring)[2]) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t gpp = p.groups_per_plane;
  const uint32_t pair = type == 0 ? 0u : 1u + item / gpp, g = item % gpp;
  const uint32_t grp[2] = {plane_group(pair, 0, g, gpp), plane_group(pair, 1, g, gpp)};
  uint8_t *dst = pair_group_base(p, b, pair, g) + 15 * 1024 + lane * 32;
  uint8_t *tap = TAP && p.tap_symbols ? p.tap_symbols + out_off + lane * kSymsPerLane + 240 : nullptr;
  rans_decode_groups<true, 2, true>(tab_s, stream, grp, kLanes, ring, p.cmp, p.cmp + p.cmp_bytes,
                                    [&](int m, const uint32_t (&w)[8]) {
                                      *reinterpret_cast<uint4 *>(dst - 1024 * m) = make_uint4(w[0], w[1], w[2], w[3]);
                                      *reinterpret_cast<uint4 *>(dst - 1024 * m + 16) = make_uint4(w[4], w[5], w[6], w[7]);
                                      if (TAP && tap) {
#pragma unroll
                                        for (int c = 0; c < 2; ++c) {
                                          const uint32_t sel = c ? 0x7531u : 0x6420u;
                                          *reinterpret_cast<uint4 *>(tap + static_cast<size_t>(grp[c]) * kGroupSyms - 16 * m) =
                                              make_uint4(__byte_perm(w[0], w[1], sel), __byte_perm(w[2], w[3], sel),
                                                         __byte_perm(w[4], w[5], sel), __byte_perm(w[6], w[7], sel));
                                        }
                                      }
                                    });
}

// One warp's share of the palette or index stream: NC consecutive groups starting at `group`.
template <int NC, bool TAP>
__device__ __forceinline__ void rans_stream_groups(const BatchParams &p, uint32_t b, uint32_t type, uint32_t group,
                                                   const uint8_t *stream, uint32_t out_off, uint32_t pal_off,
                                                   uint32_t tab_s, const uint32_t (&ring)[NC]) {
  const uint32_t lane = threadIdx.x & 31;
  uint32_t grp[NC];

#pragma unroll
  for (int c = 0; c < NC; ++c) grp[c] = group + c;
  uint8_t *tap = TAP && p.tap_symbols ? p.tap_symbols + out_off + static_cast<size_t>(group) * kGroupSyms + lane * kSymsPerLane + 240
                                      : nullptr;
  if (type == 2) {
    const uint64_t off = static_cast<uint64_t>(pal_off) + static_cast<uint64_t>(group) * kGroupSyms;
    const bool ok = off + NC * kGroupSyms <= p.palette_cap;
    uint8_t *dst = p.palette + off + lane * kSymsPerLane + 240;
    rans_decode_groups<true, NC, false>(tab_s, stream, grp, kLanes, ring, p.cmp, p.cmp + p.cmp_bytes,
                                        [&](int m, const uint32_t (&w)[NC * 4]) {
#pragma unroll
                                          for (int c = 0; c < NC; ++c) {
                                            const uint4 v = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
                                            if (ok) *reinterpret_cast<uint4 *>(dst + c * kGroupSyms - 16 * m) = v;
                                            if (TAP && tap) *reinterpret_cast<uint4 *>(tap + c * kGroupSyms - 16 * m) = v;
                                          }
                                        });
    return;
  }

  uint32_t sum[NC];  // MINUS the sum of (byte - 128) over the symbols decoded so far = the positions after the current one
#pragma unroll
  for (int c = 0; c < NC; ++c) sum[c] = 0;
  const size_t t0 = static_cast<size_t>(b) * p.n_blocks + static_cast<size_t>(group) * kGroupSyms + 15 * 512 + lane * 16;
  uint16_t *dst16 = reinterpret_cast<uint16_t *>(p.idx_s) + t0;
  uint32_t *dst32 = reinterpret_cast<uint32_t *>(p.idx_s) + t0;
  const bool idx16 = p.idx16 != 0;
  rans_decode_groups<true, NC, false>(tab_s, stream, grp, kLanes, ring, p.cmp, p.cmp + p.cmp_bytes,
                               [&](int m, const uint32_t (&wa)[NC * 4]) {
#pragma unroll
                                 for (int c = 0; c < NC; ++c) {
                                   const uint32_t w[4] = {wa[4 * c], wa[4 * c + 1], wa[4 * c + 2], wa[4 * c + 3]};
                                   if (TAP && tap) *reinterpret_cast<uint4 *>(tap + c * kGroupSyms - 16 * m) = make_uint4(w[0], w[1], w[2], w[3]);
                                   uint32_t s[16];
#pragma unroll
                                   for (int i = 15; i >= 0; --i) {
                                     s[i] = sum[c];
                                     sum[c] += 128u - ((w[i >> 2] >> (8 * (i & 3))) & 0xFFu);
                                   }
                                   if (idx16) {
                                     uint32_t q[8];
#pragma unroll
                                     for (int i = 0; i < 8; ++i) q[i] = __byte_perm(s[2 * i], s[2 * i + 1], 0x5410);
                                     uint16_t *d = dst16 + c * kGroupSyms - 512 * m;
                                     *reinterpret_cast<uint4 *>(d) = make_uint4(q[0], q[1], q[2], q[3]);
                                     *reinterpret_cast<uint4 *>(d + 8) = make_uint4(q[4], q[5], q[6], q[7]);
                                   } else {
                                     uint32_t *d = dst32 + c * kGroupSyms - 512 * m;
#pragma unroll
                                     for (int i = 0; i < 4; ++i)
                                       *reinterpret_cast<uint4 *>(d + 4 * i) = make_uint4(s[4 * i], s[4 * i + 1], s[4 * i + 2], s[4 * i + 3]);
                                   }
                                 }
                               });
  // group-local inclusive prefix at the end of every run, and the group's total: wavelet_assemble_kernel adds the
  // totals of the earlier groups of the image (the cross-group part of stage 3, codec/decode_indices.cl:66-84)
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    uint32_t inc = 0u - sum[c];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t n = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += n;
    }
    p.run_end[static_cast<size_t>(b) * (p.n_blocks / kSymsPerLane) + (group + c) * kLanes + lane] = static_cast<int32_t>(inc);
    if (lane == 31) p.idx_total[static_cast<size_t>(b) * idx_total_stride(p.groups_per_plane) + group + c] = static_cast<int32_t>(inc);
  }
}

// FT (fused tables): the CTA builds its decode table in shared memory from the 512-byte frequency block of its
// stream (stage 1, ans/build_table.cl) instead of loading the 8 KiB build_tables_kernel left in global memory.  For
// calls too small to fill the machine this removes a launch and its dependency from the critical path; for large
// batches every table would be built by up to eight CTAs instead of once, so they keep the separate kernel.
// LONE: the grid fits the machine at three CTAs per SM, so there is no occupancy to protect and the kernel is not held
// to the 48 registers that 5 CTAs per SM allow -- free of that cap ptxas keeps the table base and the ring addresses
// in registers instead of recomputing them at every checkpoint (72 registers), and a lone group finishes 20 % sooner.
template <bool TAP, bool FT>
__device__ __forceinline__ void rans_streams_cta(const BatchParams &p, const StreamGrid &sg, uint32_t b, uint32_t r, uint8_t *smem) {
  constexpr int NC = RansCfg::kChains;
  const RansSmem lay(smem_u32(smem));
  const uint32_t warp = threadIdx.x >> 5;
  uint32_t ring[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) ring[c] = lay.ring(NC * warp + c);
  // (volatile: the table base stays in its register instead of being rematerialised from S2R + LOP3 at every
  // checkpoint of the decode loop: rans_streams 1.2 % faster)
  uint32_t tab_s;
  asm volatile("mov.u32 %0, %1;" : "=r"(tab_s) : "r"(lay.tab));

  uint32_t type;  // 0 Y, 1 chroma, 2 palette, 3 index
  if (r < sg.y_ctas) type = 0;
  else if ((r -= sg.y_ctas) < sg.c_ctas) type = 1;
  else if ((r -= sg.c_ctas) < sg.pal_ctas) type = 2;
  else { r -= sg.pal_ctas; type = 3; }
  const ImageStreams is = image_streams(p, b);
  // in_off[type] / out_off[type] by selection (a dynamically indexed array would live in local memory)
  const uint32_t in_off = type == 0 ? is.in_off[0] : type == 1 ? is.in_off[1] : type == 2 ? is.in_off[2] : is.in_off[3];
  const uint32_t out_off = type == 0 ? is.out_off[0] : type == 1 ? is.out_off[1] : type == 2 ? is.out_off[2] : is.out_off[3];
  const uint8_t *stream = is.payload + in_off;
  // chains of this CTA's stream: plane streams count one per (plane, group)
  const uint32_t n_chains = type == 0 ? 2 * p.groups_per_plane : type == 1 ? 4 * p.groups_per_plane
                          : type == 2 ? is.palette_bytes / kGroupSyms : p.groups_per_plane;
  const uint32_t first = r * RansCfg::kGroupsPerCta;
  if (first >= n_chains) return;  // (the palette CTAs of the grid are sized for the largest palette of the batch)
  if (FT) {
    // the table is built in place; the scratch of the build (2 KiB) lies in the ring space behind the table, which
    // nobody uses yet.  (As static shared memory it cost the FT kernels their fifth CTA per SM: 5 x (42 KiB + 1 KiB
    // reserved) is all an SM holds.  64 x 2048^2 in one call: 0.190 -> 0.181 ms, configs[1]: 69.6 -> 68.1 us.)
    uint32_t *tab = reinterpret_cast<uint32_t *>(smem + (lay.tab - lay.s0));
    TableScratch &ts = *reinterpret_cast<TableScratch *>(smem + (lay.tab - lay.s0) + 4 * kTableSize);
    build_table_cta(reinterpret_cast<const uint16_t *>(freq_block(p, b, type, is.in_off[0])), tab, tab, ts);
  } else {
    pdl_wait();  // (the tables come from build_tables_batch_kernel)
    load_table(tab_s, p.tables + (4ull * b + type) * kTableSize, threadIdx.x, kRansWarps * 32);
  }
  __syncthreads();
  const uint32_t chain = first + warp * NC;
  if (chain >= n_chains) return;
  if (type < 2) {
    // (plane streams always hold an even number of chains: the two planes of every pair)
    rans_plane_pair<TAP>(p, b, type, chain >> 1, stream, out_off, tab_s, ring);
    return;
  }
  if (chain + 1 < n_chains) {
    rans_stream_groups<2, TAP>(p, b, type, chain, stream, out_off, is.pal_off, tab_s, ring);
  } else {
    const uint32_t ring1[1] = {ring[0]};
    rans_stream_groups<1, TAP>(p, b, type, chain, stream, out_off, is.pal_off, tab_s, ring1);
  }
}

template <bool TAP, bool FT, bool LONE>
__global__ void __launch_bounds__(kRansWarps * 32, LONE ? 3 : RansCfg::kCtasPerSm) rans_streams_kernel(const BatchParams p, const StreamGrid sg) {
  extern __shared__ __align__(1024) uint8_t smem[];
  pdl_launch_dependents();
  const uint32_t per_image = sg.per_image();
  const uint32_t b = blockIdx.x / per_image;
  rans_streams_cta<TAP, FT>(p, sg, b, blockIdx.x % per_image, smem);
  // hand-over: everything this CTA wrote is released to the tile warps of image b (see wait_for_image).  The image
  // index is taken afresh from %ctaid (volatile), so that nothing stays live across the 48-register decode loops for it
  if (p.img_done == nullptr) return;  // (single image: the tile kernel waits for the grid, see wavelet_assemble_kernel)
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t cta;
    asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(cta));
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.img_done + cta / per_image) : "memory");
  }
}

// ---------------------------------------------------------------------------------------
// Stages 4 + 5 on PACKED PLANE PAIRS.
//
// The two endpoints of a DXT1 block go through identical arithmetic: Y1/Y2, Co1/Co2 and Cg1/Cg2 are
// inverse-transformed by the same lifting steps and then converted by the same YCoCg -> 565
// formula, and the block's first word is ep1 | ep2 << 16.  So one 32-bit register carries the
// value of plane A in its low half and of plane B in its high half through the whole kernel,
// and every add works on both (rans_streams_kernel already delivers the coefficients of a pair
// byte-interleaved, so unpacking is one PRMT per pair).
//
// Representation: each half holds x + bias, with bias = kBias = 4096 for every computed value and
// kRaw = 0x1080 = 4224 for a freshly unpacked coefficient (byte | 0x10 << 8 = (byte - 128) + 4224).
// Every wavelet intermediate is bounded by |x| <= 3488 for ANY input bytes (128 + 672 per level), so
// halves stay in [608, 7712]: non-negative, and sums of three never reach 2^16 -- plain 32-bit adds
// never carry across the halves.  The reference's truncating divisions
// (codec/inverse_wavelet.cl:28-64, '/' on ints) are done on T = t + B (t = the dividend, |t| < B, B a multiple of
// the divisor): trunc_fix() below turns T into a value whose FLOOR quotient is trunc(t / k) + B / k with two packed
// 16-bit min / max instructions, the bits a 32-bit shift would carry from the high half into the low half are masked
// off, and the shift itself rides in a LEA.HI together with the add that follows it.
//
// Pipes (profiles/pipe_rates_b200.txt): LOP3 / SHF / PRMT / IADD3 / LEA / VIADDMNMX occupy the ALU pipe for two
// cycles per warp (two-input IADD, VIMNMX: one), IMAD the FMA-heavy pipe for two, and the ALU pipe -- not the issue
// slots -- bounds this kernel.  So two-input adds are written as mad.lo(a, 1, b) (IMAD.IADD) and constants are
// folded into instructions that are there anyway (the add of VIADDMNMX, the addend of LEA.HI, the free multiple of
// 256 in the fields handed to the colour conversion).  IMAD.HI is avoided altogether: it costs far more than the
// four FMA-pipe cycles its issue rate suggests (shift as multiply-high: 4 % slower than LEA.HI + IMAD.IADD).
constexpr int kBias = 4096;
constexpr int kRaw = 0x1080;
__host__ __device__ constexpr uint32_t pk(int v) { return static_cast<uint32_t>(v) * 65537u; }  // v in both halves

__device__ __forceinline__ uint32_t fma_add(uint32_t a, uint32_t b) {  // a + b on the FMA pipe
  uint32_t d;
  asm("mad.lo.u32 %0, %1, 1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t fma_sub_from(uint32_t a, uint32_t b) {  // b - a on the FMA pipe
  uint32_t d;
  asm("mad.lo.u32 %0, %1, 0xffffffff, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
// Constants the kernel wants in registers / the constant bank (BatchParams::kc): as literals ptxas rematerialises
// them with a MOV before every use, and VIADDMNMX takes one immediate only.
//   r3 = ph(3), r1 = ph(1): round-toward-zero addends;
//   cb / cb3 = ph(2) / ph(5), cr / cr3 = ph(-254) / ph(-251): dividend offsets for high-band operands of bias kBias / kRaw
struct ShiftK { uint32_t r3, r1, cb, cb3, cr, cr3; };
__host__ __device__ constexpr uint32_t ph(int v) { return (static_cast<uint32_t>(v) & 0xFFFFu) * 65537u; }  // per-half constant
// Round-toward-zero correction of a packed dividend without a sign test: with T = t + B in each half (B a multiple
// of the divisor k = R + 1, |t| < B) the value  max(T, min(T + R, B + R))  is T + R for t < 0 and has the same
// floor(. / k) as T for t >= 0, so that floor(result / k) = trunc(t / k) + B / k.  Two packed 16-bit min/max
// instructions (VIADDMNMX.S16x2 + VIMNMX.S16x2) replace the shift + mask + multiply-add of a sign-bit test.
__device__ __forceinline__ uint32_t max16x2(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("max.s16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
template <int R, int B>
__device__ __forceinline__ uint32_t trunc_fix(uint32_t T, const ShiftK &sk) {
  return max16x2(T, __viaddmin_s16x2(T, R == 3 ? sk.r3 : sk.r1, ph(B + R)));
}

// (a >> SH) + c as ONE instruction: written as a multiply-high by 2^(32 - SH) with an addend, which ptxas turns
// into LEA.HI (shift + add).  Left to the compiler as C the expression becomes SHF + LOP3 + IADD3.
template <int SH>
__device__ __forceinline__ uint32_t shr_add(uint32_t a, uint32_t c) {
  uint32_t d;
  asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "n"(1u << (32 - SH)), "r"(c));
  return d;
}

// d[2x] = s[x] - (h[x-1] + h[x] + 2) / 4.  HB = bias of the h operands; cD = pk(2048 + kBias - bias of S).
// The dividend T = t + 2^13 = (HP + HN) + c with c = 2 + 2^13 - 2 HB; c rides in the add of the two min / max
// instructions, so the sum itself is one FMA-pipe add:  T3 = max(U + c, min(U + c + 3, 2^13 + 3)).
template <int HB>
__device__ __forceinline__ uint32_t lift_even_p(uint32_t S, uint32_t HP, uint32_t HN, uint32_t cD, const ShiftK &sk) {
  const uint32_t U = fma_add(HP, HN);
  const uint32_t m = __viaddmin_s16x2(U, HB == kBias ? sk.cb3 : sk.cr3, ph(8192 + 3));
  const uint32_t T3 = __viaddmax_s16x2(U, HB == kBias ? sk.cb : sk.cr, m);
  // Q - cD = (T3 >> 2) - cD is one LEA.HI (shift + add), and S - (Q - cD) one FMA-pipe add.  (A multiply-high for
  // the shift was measured 4 % slower: IMAD.HI costs far more than its four issue cycles on the FMA-heavy pipe.)
  const uint32_t Qc = shr_add<2>(T3 & 0xFFFCFFFCu, 0u - cD);   // trunc(t / 4) + 2048 - cD
  return fma_sub_from(Qc, S);
}
// d[2x+1] = h[x] + (d[2x] + d[2x+2]) / 2.  The d operands are computed values (bias kBias); cH = pk(-bias of H).
// LAST: the result goes straight to the (char) truncation, which keeps the low byte of each half only.  With a bias of
// H that is a multiple of 256 the correction cH cannot change that byte and is not added at all; and the bit the shift
// carries from the high half into bit 15 of the low half is left in: it is not in the low byte, and the low half stays
// below 2^15 + 2^13 + 2^13, so it carries nothing back into the high half either.
template <bool LAST = false>
__device__ __forceinline__ uint32_t lift_odd_p(uint32_t H, uint32_t EP, uint32_t EN, uint32_t cH, const ShiftK &sk) {
  const uint32_t T = fma_add(EP, EN);                     // t + 2^13
  const uint32_t T2 = trunc_fix<1, 8192>(T, sk);
  if (LAST) return shr_add<1>(T2, H);
  return shr_add<1>(T2 & 0xFFFEFFFEu, H) + cH;            // h + trunc(t / 2) + 4096 - bias of H
}
// 1-D inverse 5/3 lifting of v = [low half | high half] in registers, codec/inverse_wavelet.cl:28-64
// (NormalizeIndex mirror resolved at compile time).  HB = bias of the high half; the result has bias kBias.
template <int LEN, int HB, bool LAST = false>
__device__ __forceinline__ void inverse_lift_p(uint32_t (&v)[LEN], uint32_t cD, const ShiftK &sk) {
  static_assert(!LAST || HB % 256 == 0, "the dropped correction must be a multiple of 256 per half");
  constexpr int MID = LEN / 2;
  uint32_t o[LEN];
#pragma unroll
  for (int x = 0; x < MID; ++x) o[2 * x] = lift_even_p<HB>(v[x], v[MID + (x == 0 ? 0 : x - 1)], v[MID + x], cD, sk);
#pragma unroll
  for (int x = 0; x < MID; ++x)
    o[2 * x + 1] = lift_odd_p<LAST>(v[MID + x], o[2 * x], o[(2 * x + 2 == LEN) ? 2 * x : 2 * x + 2], pk(-HB), sk);
#pragma unroll
  for (int i = 0; i < LEN; ++i) v[i] = o[i];
}

// Work area of one warp: three planes W[p] (p = plane pair) of 32 rows x 32 packed words, 128 B per
// row, the 16-byte chunk j of row r stored at chunk j ^ (r & 7) (lane = row 16-byte accesses, lane =
// column 4-byte accesses and the assembly's lane = 4 columns accesses are all conflict free).
// Before a row is transformed at level 32 its logical chunks 4..7 hold the row's 64 raw coefficient
// bytes (32 columns x {plane A, plane B}), and the top-left 16x16 words of W[p] are the work area of
// the lower levels: everything is done in place, the tile never leaves these 12 KiB.
constexpr int kWPlane = 4096;
constexpr int kWarpWork = 3 * kWPlane;
__device__ __forceinline__ uint32_t wchunk(uint32_t wp, uint32_t r, uint32_t j) { return wp + r * 128 + (((j ^ r) & 7) << 4); }

// coefficient pair (col 2m + J of both planes) of an interleaved word {A[2m], B[2m], A[2m+1], B[2m+1]}:
// byte | 0x10 << 8 in each half = (byte - 128) + kRaw
template <int J>
__device__ __forceinline__ uint32_t unp(uint32_t w, uint32_t k10) {
  return J == 0 ? __byte_perm(w, k10, 0x4140) : __byte_perm(w, k10, 0x4342);
}
template <int N>  // N words -> 2N coefficient pairs
__device__ __forceinline__ void unp_words(const uint32_t (&w)[N], uint32_t *v, uint32_t k10) {
#pragma unroll
  for (int i = 0; i < N; ++i) { v[2 * i] = unp<0>(w[i], k10); v[2 * i + 1] = unp<1>(w[i], k10); }
}
__device__ __forceinline__ void ld4(uint32_t a, uint32_t *v) { const uint4 t = lds128(a); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
__device__ __forceinline__ void ld2(uint32_t a, uint32_t *v) { const uint2 t = lds64(a); v[0] = t.x; v[1] = t.y; }

// Row r of the level-LEN region: low band = the previous level's result (rows < LEN/2) or raw
// coefficients, high band = raw coefficients.
template <int LEN>
__device__ __forceinline__ void load_row(uint32_t wp, uint32_t r, uint32_t (&v)[LEN], uint32_t k10) {
  if constexpr (LEN == 32) {
    uint32_t w[8];
    if (r < 16) {
#pragma unroll
      for (int j = 0; j < 4; ++j) ld4(wchunk(wp, r, j), v + 4 * j);
    } else {
      ld4(wchunk(wp, r, 4), w); ld4(wchunk(wp, r, 5), w + 4);
      unp_words<8>(w, v, k10);
    }
    ld4(wchunk(wp, r, 6), w); ld4(wchunk(wp, r, 7), w + 4);
    unp_words<8>(w, v + 16, k10);
  } else if constexpr (LEN == 16) {
    uint32_t w[8];
    if (r < 8) {
      ld4(wchunk(wp, r, 0), v); ld4(wchunk(wp, r, 1), v + 4);
      ld4(wchunk(wp, r, 5), w);
      uint32_t w4[4] = {w[0], w[1], w[2], w[3]};
      unp_words<4>(w4, v + 8, k10);
    } else {
      ld4(wchunk(wp, r, 4), w); ld4(wchunk(wp, r, 5), w + 4);
      unp_words<8>(w, v, k10);
    }
  } else if constexpr (LEN == 8) {
    uint32_t w[4];
    if (r < 4) {
      ld4(wchunk(wp, r, 0), v);
      ld2(wchunk(wp, r, 4) + 8, w);
      uint32_t w2[2] = {w[0], w[1]};
      unp_words<2>(w2, v + 4, k10);
    } else {
      ld4(wchunk(wp, r, 4), w);
      unp_words<4>(w, v, k10);
    }
  } else if constexpr (LEN == 4) {
    uint32_t w[2];
    if (r < 2) {
      ld2(wchunk(wp, r, 0), v);
      uint32_t w1[1] = {lds32(wchunk(wp, r, 4) + 4)};
      unp_words<1>(w1, v + 2, k10);
    } else {
      ld2(wchunk(wp, r, 4), w);
      unp_words<2>(w, v, k10);
    }
  } else {  // LEN == 2: the 2x2 corner is all raw (its [0][0] is the tile's DC coefficient)
    uint32_t w1[1] = {lds32(wchunk(wp, r, 4))};
    unp_words<1>(w1, v, k10);
  }
}
template <int LEN>
__device__ __forceinline__ void store_row(uint32_t wp, uint32_t r, const uint32_t (&v)[LEN]) {
  if constexpr (LEN >= 4) {
#pragma unroll
    for (int j = 0; j < LEN / 4; ++j) sts128(wchunk(wp, r, j), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  } else {
    sts64(wchunk(wp, r, 0), v[0], v[1]);
  }
}
// column c of a plane: byte offsets of word (i, c) for i & 7 = 0..7 (add i * 128)
__device__ __forceinline__ void col_offsets(uint32_t c, uint32_t (&off)[8]) {
#pragma unroll
  for (int s = 0; s < 8; ++s) off[s] = ((((c >> 2) ^ s) & 7) << 4) + (c & 3) * 4;
}

// One level (LEN = 2..16) of NP plane pairs (all three, or pair pl0 alone): row pass with lane = (plane pair, row),
// then column pass with lane = (plane pair, column).
template <int LEN, int NP = 3>
__device__ __forceinline__ void low_level_p(uint32_t w_s, uint32_t lane, uint32_t k10, const ShiftK &sk, uint32_t pl0 = 0) {
  constexpr int ITEMS = NP * LEN;
#pragma unroll 1  // (code size: the kernel has to stay inside the 32 KiB L1.5 instruction cache)
  for (int base = 0; base < ITEMS; base += 32) {
    const uint32_t item = base + lane;
    if (item < ITEMS) {
      const uint32_t wp = w_s + (NP == 1 ? pl0 : item / LEN) * kWPlane, r = item % LEN;
      uint32_t v[LEN];
      load_row<LEN>(wp, r, v, k10);
      // low band: the previous level's result (bias kBias) in rows < LEN / 2, raw coefficients below
      inverse_lift_p<LEN, kRaw>(v, (LEN > 2 && r < LEN / 2) ? pk(2048) : pk(2048 + kBias - kRaw), sk);
      store_row<LEN>(wp, r, v);
    }
  }
  __syncwarp();
#pragma unroll 1  // (code size: the kernel has to stay inside the 32 KiB L1.5 instruction cache)
  for (int base = 0; base < ITEMS; base += 32) {
    const uint32_t item = base + lane;
    if (item < ITEMS) {
      const uint32_t wp = w_s + (NP == 1 ? pl0 : item / LEN) * kWPlane, c = item % LEN;
      uint32_t off[8];
      col_offsets(c, off);
      uint32_t v[LEN];
#pragma unroll
      for (int i = 0; i < LEN; ++i) v[i] = lds32(wp + off[i & 7] + i * 128);
      inverse_lift_p<LEN, kBias>(v, pk(2048), sk);
#pragma unroll
      for (int i = 0; i < LEN; ++i) sts32(wp + off[i & 7] + i * 128, v[i]);
    }
  }
  __syncwarp();
}

// The fields the wavelet stores for the assembly are (int8 value + 128) + a multiple of 256 that costs nothing (it is
// part of the XOR constant of the (char) truncation) and is chosen so that no step of the colour conversion needs
// an add of its own for a constant:
constexpr int kBY = 1024, kBO = 0, kBG = 512;
constexpr int kB1 = 512;                             // bias of the dividend -cg
constexpr int kB2 = 256 + kBY - kBO;                 // bias of the dividend t - co
constexpr int kBiasG = 256 + kBG + kBY + kB1 / 2;    // g + kBiasG: a multiple of 2048, so that (g + bias) << 5 only spills into masked bits
constexpr int kBiasB = kB2 / 2;                      // b + kBiasB
constexpr int kBiasR = kBiasB + 128 + kBO;           // r + kBiasR: a multiple of 32
static_assert(kBiasG % 2048 == 0 && kBiasR % 32 == 0 && kB2 % 2 == 0 && kB2 >= 512, "colour conversion biases");
// codec/assemble.cl:39-62 for both endpoints at once: YCoCg667 -> RGB565 with truncating division and
// an unmasked shift/or pack.  Inputs: (int8 value + 128) of plane A | plane B << 16.
//   t = y - cg / 2;  g = cg + t;  b = (t - co) / 2;  r = b + co;  out = r << 11 | g << 5 | b  (low 16 bits)
// Outputs (per half): R = r + kBiasR, G = g + kBiasG, Bq = b + kBiasB.
__device__ __forceinline__ void ycocg_to_rgb_p(uint32_t Y, uint32_t CO, uint32_t CG, const ShiftK &sk, uint32_t &R, uint32_t &G,
                                               uint32_t &Bq) {
  const uint32_t Tn = fma_sub_from(CG, pk(kB1 + 128 + kBG));          // -cg + kB1; trunc(-cg / 2) = -trunc(cg / 2)
  const uint32_t Tb = shr_add<1>(trunc_fix<1, kB1>(Tn, sk) & 0xFFFEFFFEu, Y);   // t + kB1 / 2 + 128 + kBY
  G = fma_add(CG, Tb);                                                // g + kBiasG
  const uint32_t V = fma_sub_from(CO, Tb);                            // (t - co) + kB2
  Bq = (trunc_fix<1, kB2>(V, sk) & 0xFFFEFFFEu) >> 1;                  // b + kB2 / 2
  R = fma_add(Bq, CO);                                                // r + kBiasR
}
__device__ __forceinline__ uint32_t pack565_p(uint32_t Y, uint32_t CO, uint32_t CG, const ShiftK &sk) {  // ep1 | ep2 << 16
  uint32_t R, G, Bq;
  ycocg_to_rgb_p(Y, CO, CG, sk, R, G, Bq);
  uint32_t rs, gs;
  asm("mul.lo.u32 %0, %1, 2048;" : "=r"(rs) : "r"(R));  // << 11 and << 5 on the FMA pipe
  asm("mul.lo.u32 %0, %1, 32;" : "=r"(gs) : "r"(G));
  // b mod 2^16 in each half: a per-half add (VIADD.16x2, FMA pipe) takes the bias off without a borrow between
  // the halves, so the three fields combine in two LOP3
  uint32_t bx;
  asm("add.s16x2 %0, %1, %2;" : "=r"(bx) : "r"(Bq), "r"(ph(-kBiasB)));
  return (rs & 0xF800F800u) | ((gs & 0xFFE0FFE0u) | bx);
}

// ---------------------------------------------------------------------------------------
// Stages 4 + 5.  One warp per 32x32 tile (1024 DXT blocks) of one image, all six planes as three
// plane pairs [Y1|Y2, Co1|Co2, Cg1|Cg2] (codec/assemble.cl:27-37); warps share nothing, so there is
// no CTA barrier at all.
//   fetch   : the tile's coefficients, 3 x 2 KiB, cp.async from sym_t into the raw halves of the W
//             rows: row r of tile tq of a pair-group is the 32-byte pieces k = 2 (r % 8), 2 (r % 8) + 1
//             of run 4 tq + r / 8 (rans_streams_kernel)
//   wavelet : levels 2..16 of the three pairs together (lane = pair x row, then pair x column), level 32
//             per pair (lane = row with the 32 packed coefficients in registers, then lane = column);
//             the (char)-truncated result replaces W
//   assembly: 4 DXT1 blocks (or 64 RGB8 texels) per lane and 4-row slab, coalesced 16-byte
//             stores; palette index = run_end - S.  The index / palette loads of the next slab
//             are in flight while a slab is assembled, those of slab 0 during the wavelet.
// One tile warp per CTA: the warps share nothing, and a CTA of two held its slot until the slower one was done
// (2.529 -> 2.513 ms per step; three per CTA: 2.545).
constexpr int kWaWarps = 1;
constexpr int kWaSmem = kWaWarps * kWarpWork;  // 12288

// IDX16: the index suffix sums are u16 (every palette of the batch has <= 65536 entries), else u32.
template <int RGB, bool TAP, bool IDX16>
__global__ void __launch_bounds__(kWaWarps * 32, 18 / kWaWarps) wavelet_assemble_kernel(const BatchParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t w_s = smem_u32(smem) + warp * kWarpWork;
  // grid = (tiles of an image, images): no division by the tiles per image
  const uint32_t b = blockIdx.y;
  const uint32_t tile = blockIdx.x * kWaWarps + warp;
  if (tile >= p.n_blocks / kTileSyms) return;
  const uint32_t tiles_x = p.blocks_x / kTile;
  const uint32_t ty = tile / tiles_x, tx = tile % tiles_x;
  // everything below reads what rans_streams_kernel wrote for image b.  A single image is complete when the decode
  // grid is, and griddepcontrol.wait is the cheaper wait then (the hand-over costs a lone image 2 us: a fence at the
  // end of its decode CTAs, a probe and a fence here)
  if (p.img_done != nullptr) wait_for_image(p.img_done + b, p.rans_ctas, p.status);
  else pdl_wait();

  // ---- the tile's coefficients: sym_t -> raw halves of the W rows -------------------------------
  {
    const uint32_t tq = tile & 7;
    const uint8_t *src0 = p.sym_t + static_cast<size_t>(b) * 6 * p.n_blocks + static_cast<size_t>(tile >> 3) * (2 * kGroupSyms);
    const size_t pair_stride = static_cast<size_t>(p.groups_per_plane) * (2 * kGroupSyms);  // = 2N
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t pc = lane + 32 * i, r = pc >> 2, j = pc & 3;
        cp_async16(wchunk(w_s + pl * kWPlane, r, 4 + j),
                   src0 + pl * pair_stride + (2 * (r & 7) + (j >> 1)) * 1024 + (4 * tq + (r >> 3)) * 32 + (j & 1) * 16);
      }
    }
    cp_async_commit();
  }

  // out_off[4b..4b+3] of this image (codec/decoder.cpp:430-463): requested now, first used after the
  // lower wavelet levels, so the warp never waits for it
  const uint4 out_off = p.inline_off ? make_uint4(p.off8[0], p.off8[1], p.off8[2], p.off8[3]) : __ldg(reinterpret_cast<const uint4 *>(p.cmp) + b);
  const uint32_t *pal = nullptr;

  // ---- assembly inputs of slab 0 start their trip now ---------------------------------
  const size_t img_block0 = static_cast<size_t>(b) * p.n_blocks;
  // index prefix at the end of a run = its group-local prefix (rans_streams_kernel) + the totals of the earlier
  // index groups of the image.  The 32 blocks of a tile row lie in one 256-block run, so lane l fetches the run end
  // of tile row l once and the slabs pick theirs up with a shuffle.  The carry: the warp sums the totals of
  // the groups before the tile's first one (lanes stride over them, one REDUX), then walks the few groups the 32
  // rows span -- no atomics, no accumulator to zero between calls.
  // The loads start now; the sum is taken once the tile's coefficients have arrived (resolve_run_ends below), so
  // that no warp waits for them alone.
  const uint32_t g_row = (ty * kTile + lane) * p.blocks_x + tx * kTile;
  const uint32_t grp = g_row / kGroupSyms;
  const uint32_t g_first = (ty * kTile * p.blocks_x + tx * kTile) / kGroupSyms;                 // group of tile row 0
  const uint32_t g_last = ((ty * kTile + kTile - 1) * p.blocks_x + tx * kTile) / kGroupSyms;    // ... of tile row 31
  const int32_t *tot = p.idx_total + static_cast<size_t>(b) * idx_total_stride(p.groups_per_plane);
  const int32_t tot_lane = lane < g_first ? __ldg(tot + lane) : 0;   // this lane's share of the groups before the tile
  const int32_t tot_first = __ldg(tot + g_first);
  uint32_t re_row = static_cast<uint32_t>(__ldg(p.run_end + static_cast<size_t>(b) * (p.n_blocks / kSymsPerLane) + g_row / kSymsPerLane));
  auto resolve_run_ends = [&]() {
    int32_t part = tot_lane;
#pragma unroll 1  // (code size: the kernel has to stay inside the 32 KiB L1.5 instruction cache)
    for (uint32_t g = lane + 32; g < g_first; g += 32) part += __ldg(tot + g);   // (images beyond 2048 x 2048 only)
    int32_t carry = __reduce_add_sync(0xffffffffu, part);
    if (grp > g_first) carry += tot_first;
#pragma unroll 1
    for (uint32_t g = g_first + 1; g < g_last; ++g) {                            // (tile rows spanning > 2 groups only)
      const int32_t t = __ldg(tot + g);
      if (grp > g) carry += t;
    }
    re_row += static_cast<uint32_t>(carry);
  };
  // first block of this lane in slab k (rows 4k..4k+3 of the tile)
  const uint32_t gidx0 = (ty * kTile + (lane >> 3)) * p.blocks_x + tx * kTile + 4 * (lane & 7);
  const uint32_t slab_stride = 4 * p.blocks_x;
  // S of 4 blocks: 4 x u16 in raw.x/.y, or 4 x u32
  struct Sfx { uint4 raw; };
  // transposed S: [group][k = pos / 16][run][pos % 16], pos = position inside the 256-block run
  auto sfx_ptr = [&](uint32_t gidx) -> const uint8_t * {
    // = (gidx & ~8191) + ((gidx & 255) >> 4) * 512 + ((gidx & 8191) >> 8) * 16 + (gidx & 15), with h = gidx >> 4
    const uint32_t h = gidx >> 4;
    const uint32_t e32 = (gidx & ~8191u) + ((h & 15u) << 9) + (h & 0x1F0u) + (gidx & 15u);
    return reinterpret_cast<const uint8_t *>(p.idx_s) + (IDX16 ? 2 : 4) * (img_block0 + e32);
  };
  auto load_sfx_at = [&](const uint8_t *sp) -> Sfx {
    Sfx r;
    if (IDX16) {
      const uint2 sv = __ldg(reinterpret_cast<const uint2 *>(sp));
      r.raw = make_uint4(sv.x, sv.y, 0u, 0u);
    } else {
      r.raw = __ldg(reinterpret_cast<const uint4 *>(sp));
    }
    return r;
  };
  auto load_sfx = [&](uint32_t gidx) -> Sfx { return load_sfx_at(sfx_ptr(gidx)); };
  // the suffix sums come from DRAM (the rANS kernel wrote them): pull them into L1 two slabs before
  // they are loaded, so that the load -> index -> palette gather chain of a slab starts on time
  auto prefetch_sfx = [&](const uint8_t *sp) { asm volatile("prefetch.global.L1 [%0];" ::"l"(sp)); };
  // k = slab of these suffix sums.  idx = min(run_end + S, n_entries - 1): with 16-bit sums two indices are one
  // VIADDMNMX.U16x2 (add, wrap and clamp per half); `seen` keeps the largest unclamped index for the status flag.
  uint32_t nmax = 0, seen = 0;  // n_entries - 1 (in both halves when IDX16)
  auto load_words = [&](const Sfx &sf, uint32_t k, uint32_t gidx, uint32_t (&word)[4]) {
    const uint32_t re = __shfl_sync(0xffffffffu, re_row, 4 * k + (lane >> 3));
    uint32_t idx[4];
    if (IDX16) {
      const uint32_t re2 = __byte_perm(re, 0u, 0x1010);  // run end mod 2^16 in both halves
      const uint32_t p0 = __viaddmin_u16x2(re2, sf.raw.x, nmax), p1 = __viaddmin_u16x2(re2, sf.raw.y, nmax);
      seen = __viaddmax_u16x2(re2, sf.raw.x, seen);
      seen = __viaddmax_u16x2(re2, sf.raw.y, seen);
      idx[0] = p0 & 0xFFFFu; idx[1] = p0 >> 16; idx[2] = p1 & 0xFFFFu; idx[3] = p1 >> 16;
    } else {
      const uint32_t u[4] = {re + sf.raw.x, re + sf.raw.y, re + sf.raw.z, re + sf.raw.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        seen = max(seen, u[j]);
        idx[j] = min(u[j], nmax);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) word[j] = __ldg(pal + idx[j]);
    if (TAP && p.tap_indices) {  // the unclamped indices, as the reference computes them
      uint32_t u[4];
      if (IDX16) {
        u[0] = (re + sf.raw.x) & 0xFFFFu; u[1] = (re + (sf.raw.x >> 16)) & 0xFFFFu;
        u[2] = (re + sf.raw.y) & 0xFFFFu; u[3] = (re + (sf.raw.y >> 16)) & 0xFFFFu;
      } else {
        u[0] = re + sf.raw.x; u[1] = re + sf.raw.y; u[2] = re + sf.raw.z; u[3] = re + sf.raw.w;
      }
      *reinterpret_cast<uint4 *>(p.tap_indices + img_block0 + gidx) = make_uint4(u[0], u[1], u[2], u[3]);
    }
  };
  // software pipeline of the assembly inputs: the palette words of slab k + 1 are gathered while slab k
  // is assembled, from suffix sums loaded two slabs earlier and prefetched into L1 two slabs before that.
  // Slot (k + 1) & 1 of sfx / pf / words belongs to slab k + 1 (then k + 3, k + 5): with the slab loop unrolled
  // by two nothing has to be moved from register to register.
  Sfx sfx[2] = {load_sfx(gidx0), load_sfx(gidx0 + slab_stride)};
  const uint8_t *pf[2] = {sfx_ptr(gidx0 + 2 * slab_stride), sfx_ptr(gidx0 + 3 * slab_stride)};  // prefetched, not yet loaded
  prefetch_sfx(pf[0]);
  prefetch_sfx(pf[1]);

  // ---- stage 4: inverse wavelet ------------------------------------------------------------
  uint32_t k10 = 0x10101010u;
  asm volatile("" : "+r"(k10));  // keep it in a register (PRMT takes no immediate source)
  cp_async_wait_group<0>();
  __syncwarp();
  resolve_run_ends();  // (its loads were issued with the tile's: they have landed too)
  // r3 / r1 come from the kernel parameters: as literals ptxas rematerialises them with a MOV before every use
  const ShiftK sk{p.kc[0], p.kc[1], p.kc[3], p.kc[4], p.kc[5], p.kc[6]};
  low_level_p<2>(w_s, lane, k10, sk);
  low_level_p<4>(w_s, lane, k10, sk);
  low_level_p<8>(w_s, lane, k10, sk);
  low_level_p<16>(w_s, lane, k10, sk);

  uint32_t words[2][4];  // palette words of slab k in words[k & 1]
  {  // slab 0: indices -> palette words;  slabs 1, 2: suffix sums in registers
    const uint32_t palette_bytes = out_off.w - out_off.z;
    const uint32_t pal_off = out_off.z - 7u * p.n_blocks * b - 6u * p.n_blocks;  // compact palette scratch
    const uint32_t n_entries = palette_bytes / 4;
    // a palette outside the scratch (malformed offsets) or an empty one: every index reads the zero word, flag 2
    const bool pal_ok = static_cast<uint64_t>(pal_off) + palette_bytes <= p.palette_cap && n_entries > 0;
    pal = pal_ok ? reinterpret_cast<const uint32_t *>(p.palette + pal_off) : p.status + 1;
    nmax = pal_ok ? n_entries - 1 : 0u;
    if (!pal_ok && lane == 0) atomicOr(p.status, 2u);
    if (IDX16) nmax = __byte_perm(nmax, 0u, 0x1010);
    load_words(sfx[0], 0, gidx0, words[0]);
    sfx[0] = load_sfx_at(pf[0]);
    pf[0] = sfx_ptr(gidx0 + 4 * slab_stride);
    prefetch_sfx(pf[0]);
  }

  // level 32, rows: lane = row
#pragma unroll 1
  for (uint32_t pl = 0; pl < 3; ++pl) {
    const uint32_t wp = w_s + pl * kWPlane;
    uint32_t v[32];
    load_row<32>(wp, lane, v, k10);
    inverse_lift_p<32, kRaw>(v, lane < 16 ? pk(2048) : pk(2048 + kBias - kRaw), sk);
    store_row<32>(wp, lane, v);
  }
  __syncwarp();
  // level 32, columns: lane = column; (char) truncation (codec/inverse_wavelet.cl:188-190)
  {
    uint32_t off[8];
    col_offsets(lane, off);
#pragma unroll 1
    for (uint32_t pl = 0; pl < 3; ++pl) {
      const uint32_t wp = w_s + pl * kWPlane;
      uint32_t v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = lds32(wp + off[i & 7] + i * 128);
      inverse_lift_p<32, kBias, true>(v, pk(2048), sk);
      // (char) truncation: low byte of (x + 4096) = x mod 256; ^ 0x80 makes it (int8) x + 128
      const uint32_t kx = pl == 0 ? ph(128 + kBY) : pl == 1 ? ph(128 + kBO) : ph(128 + kBG);
#pragma unroll
      for (int i = 0; i < 32; ++i) sts32(wp + off[i & 7] + i * 128, (v[i] & 0x00FF00FFu) ^ kx);
      if (TAP && p.tap_planes) {
        // reference plane order [Y1, Y2, Co1, Cg1, Co2, Cg2]: pair 0 = planes (0, 1), pair 1 = (2, 4), pair 2 = (3, 5)
        const uint32_t pa = pl == 0 ? 0 : pl == 1 ? 2 : 3, pb = pl == 0 ? 1 : pl == 1 ? 4 : 5;
        int8_t *t0 = p.tap_planes + static_cast<size_t>(b) * 6 * p.n_blocks + static_cast<size_t>(ty * kTile) * p.blocks_x + tx * kTile + lane;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          t0[static_cast<size_t>(pa) * p.n_blocks + static_cast<size_t>(i) * p.blocks_x] = static_cast<int8_t>(v[i] & 0xFFu);
          t0[static_cast<size_t>(pb) * p.n_blocks + static_cast<size_t>(i) * p.blocks_x] = static_cast<int8_t>((v[i] >> 16) & 0xFFu);
        }
      }
    }
  }
  __syncwarp();

  // ---- stage 5: assembly, codec/assemble.cl:64-129 ---------------------------------------
  auto slab = [&](auto q_c, uint32_t k) {
    constexpr int Q = decltype(q_c)::value;  // (k + 1) & 1
    const uint32_t gidx = gidx0 + k * slab_stride;
    const uint32_t (&word)[4] = words[Q ^ 1];
    if (k + 1 < 8) load_words(sfx[Q], k + 1, gidx + slab_stride, words[Q]);  // its S was loaded two slabs ago
    if (k + 3 < 8) sfx[Q] = load_sfx_at(pf[Q]);
    if (k + 5 < 8) {
      pf[Q] = sfx_ptr(gidx + 5 * slab_stride);
      prefetch_sfx(pf[Q]);
    }
    // rows 4k..4k+3 of the tile, 4 blocks per lane: (int8 + 128) of plane A | plane B << 16
    const uint32_t src = wchunk(w_s, 4 * k + (lane >> 3), lane & 7);
    uint32_t Y[4], CO[4], CG[4];
    ld4(src, Y); ld4(src + kWPlane, CO); ld4(src + 2 * kWPlane, CG);

    if (!RGB) {
      uint32_t o[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // PhysicalDXTBlock (codec/dxt_image.h:14-21): u16 ep1, u16 ep2, u32 interpolation
        o[2 * j] = pack565_p(Y[j], CO[j], CG[j], sk);
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
        uint32_t R, G, Bq;
        ycocg_to_rgb_p(Y[j], CO[j], CG[j], sk, R, G, Bq);
        int c0[3] = {static_cast<int>(R & 0xFFFFu) - kBiasR, static_cast<int>(G & 0xFFFFu) - kBiasG, static_cast<int>(Bq & 0xFFFFu) - kBiasB};
        int c1[3] = {static_cast<int>(R >> 16) - kBiasR, static_cast<int>(G >> 16) - kBiasG, static_cast<int>(Bq >> 16) - kBiasB};
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
      const size_t img_w = 4ull * p.blocks_x;
      uint8_t *img = p.out + static_cast<size_t>(b) * p.n_blocks * 48;
      const size_t x0 = 4ull * (tx * kTile + 4 * (lane & 7)), y0 = 4ull * (ty * kTile + 4 * k + (lane >> 3));
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
  };
#pragma unroll 1
  for (uint32_t k = 0; k < 8; k += 2) {
    slab(std::integral_constant<int, 1>{}, k);
    slab(std::integral_constant<int, 0>{}, k + 1);
  }
  // status: a palette index beyond the palette was clamped (a stream the reference encoder would not have written)
  {
    const bool over = IDX16 ? ((seen & 0xFFFFu) > (nmax & 0xFFFFu) || (seen >> 16) > (nmax >> 16)) : seen > nmax;
    if (__any_sync(0xffffffffu, over) && lane == 0) atomicOr(p.status, 1u);
  }
  // (no griddepcontrol.wait on the hand-over path: a warp blocked in it would hold its CTA slot until the whole
  // decode grid has drained, and the early tiles are there to free theirs for the next ones.  The last tiles of the
  // last image cannot finish before every decode CTA has counted itself in, which is the last thing those do.  The
  // counters are not reset here either -- an atomic whose result a warp has to wait for keeps its slot busy for a
  // microsecond per tile: every call gets fresh, zeroed counters from the host side, gst_capi.cu.)
}

// ---------------------------------------------------------------------------------------
// Stages 4 + 5 for calls of so few tiles that the machine is empty (a single image up to 4096 x 4096): one CTA of
// THREE warps per tile instead of one warp.  wavelet_assemble_kernel is built for throughput -- a warp takes a tile
// through ~4000 instructions on its own, 10-14 us when it has an SM sub-partition to itself; here warp w
// transforms plane pair w (the pairs are independent until the assembly), the CTA meets at one barrier, and the
// eight 4-row slabs of the assembly are dealt out to the warps (slabs w, w + 3, w + 6).  The suffix sums and the
// palette words of a warp's slabs are requested before the wavelet, so the assembly finds them in registers.  Same
// arithmetic, same helpers, same scratch layout as wavelet_assemble_kernel; DXT1 output only (RGB8 and the parity
// taps keep the one-warp kernel).
constexpr int kSplitWarps = 3;
constexpr uint32_t kSplitMaxTiles = 7 * 148;  // the grid has to fit the machine in one wave (7 CTAs per SM)

template <bool IDX16>
__global__ void __launch_bounds__(kSplitWarps * 32, 7) wavelet_assemble_split_kernel(const BatchParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t w_s = smem_u32(smem);
  const uint32_t b = blockIdx.y, tile = blockIdx.x;
  const uint32_t tiles_x = p.blocks_x / kTile;
  const uint32_t ty = tile / tiles_x, tx = tile % tiles_x;
  if (p.img_done != nullptr) wait_for_image(p.img_done + b, p.rans_ctas, p.status);
  else pdl_wait();

  // this warp's plane pair: sym_t -> raw halves of the W rows (as in wavelet_assemble_kernel)
  const uint32_t wp = w_s + warp * kWPlane;
  {
    const uint32_t tq = tile & 7;
    const uint8_t *src0 = p.sym_t + static_cast<size_t>(b) * 6 * p.n_blocks + static_cast<size_t>(tile >> 3) * (2 * kGroupSyms) +
                          static_cast<size_t>(warp) * p.groups_per_plane * (2 * kGroupSyms);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t pc = lane + 32 * i, r = pc >> 2, j = pc & 3;
      cp_async16(wchunk(wp, r, 4 + j), src0 + (2 * (r & 7) + (j >> 1)) * 1024 + (4 * tq + (r >> 3)) * 32 + (j & 1) * 16);
    }
    cp_async_commit();
  }
  const uint4 out_off = p.inline_off ? make_uint4(p.off8[0], p.off8[1], p.off8[2], p.off8[3]) : __ldg(reinterpret_cast<const uint4 *>(p.cmp) + b);

  // run ends of the 32 tile rows (lane = tile row) and the carry of the earlier index groups, as in the one-warp kernel
  const size_t img_block0 = static_cast<size_t>(b) * p.n_blocks;
  const uint32_t g_row = (ty * kTile + lane) * p.blocks_x + tx * kTile;
  const uint32_t grp = g_row / kGroupSyms;
  const uint32_t g_first = (ty * kTile * p.blocks_x + tx * kTile) / kGroupSyms;
  const uint32_t g_last = ((ty * kTile + kTile - 1) * p.blocks_x + tx * kTile) / kGroupSyms;
  const int32_t *tot = p.idx_total + static_cast<size_t>(b) * idx_total_stride(p.groups_per_plane);
  int32_t part = 0;
  for (uint32_t g = lane; g < g_first; g += 32) part += __ldg(tot + g);
  const int32_t tot_first = __ldg(tot + g_first);
  uint32_t re_row = static_cast<uint32_t>(__ldg(p.run_end + static_cast<size_t>(b) * (p.n_blocks / kSymsPerLane) + g_row / kSymsPerLane));

  // suffix sums of this warp's slabs k = warp, warp + 3, warp + 6 (transposed layout: see sfx_ptr in the one-warp kernel)
  const uint32_t gidx0 = (ty * kTile + (lane >> 3)) * p.blocks_x + tx * kTile + 4 * (lane & 7);
  const uint32_t slab_stride = 4 * p.blocks_x;
  uint4 sfx[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const uint32_t k = warp + 3 * i;
    sfx[i] = make_uint4(0u, 0u, 0u, 0u);
    if (k < 8) {
      const uint32_t gidx = gidx0 + k * slab_stride, h = gidx >> 4;
      const uint32_t e32 = (gidx & ~8191u) + ((h & 15u) << 9) + (h & 0x1F0u) + (gidx & 15u);
      const uint8_t *sp = reinterpret_cast<const uint8_t *>(p.idx_s) + (IDX16 ? 2 : 4) * (img_block0 + e32);
      if (IDX16) {
        const uint2 sv = __ldg(reinterpret_cast<const uint2 *>(sp));
        sfx[i] = make_uint4(sv.x, sv.y, 0u, 0u);
      } else {
        sfx[i] = __ldg(reinterpret_cast<const uint4 *>(sp));
      }
    }
  }

  uint32_t k10 = 0x10101010u;
  asm volatile("" : "+r"(k10));
  cp_async_wait_group<0>();
  __syncwarp();
  {
    int32_t carry = __reduce_add_sync(0xffffffffu, part);
    if (grp > g_first) carry += tot_first;
    for (uint32_t g = g_first + 1; g < g_last; ++g) {
      const int32_t t = __ldg(tot + g);
      if (grp > g) carry += t;
    }
    re_row += static_cast<uint32_t>(carry);
  }

  // palette words of this warp's slabs: requested now, used after the barrier
  uint32_t words[3][4];
  uint32_t nmax, seen = 0;
  {
    const uint32_t palette_bytes = out_off.w - out_off.z;
    const uint32_t pal_off = out_off.z - 7u * p.n_blocks * b - 6u * p.n_blocks;
    const uint32_t n_entries = palette_bytes / 4;
    const bool pal_ok = static_cast<uint64_t>(pal_off) + palette_bytes <= p.palette_cap && n_entries > 0;
    const uint32_t *pal = pal_ok ? reinterpret_cast<const uint32_t *>(p.palette + pal_off) : p.status + 1;
    nmax = pal_ok ? n_entries - 1 : 0u;
    if (!pal_ok && threadIdx.x == 0) atomicOr(p.status, 2u);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const uint32_t k = warp + 3 * i;
      const uint32_t re = __shfl_sync(0xffffffffu, re_row, (4 * k + (lane >> 3)) & 31);
      uint32_t u[4];
      if (IDX16) {
        u[0] = (re + sfx[i].x) & 0xFFFFu; u[1] = (re + (sfx[i].x >> 16)) & 0xFFFFu;
        u[2] = (re + sfx[i].y) & 0xFFFFu; u[3] = (re + (sfx[i].y >> 16)) & 0xFFFFu;
      } else {
        u[0] = re + sfx[i].x; u[1] = re + sfx[i].y; u[2] = re + sfx[i].z; u[3] = re + sfx[i].w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (k < 8) seen = max(seen, u[j]);
        words[i][j] = k < 8 ? __ldg(pal + min(u[j], nmax)) : 0u;
      }
    }
  }

  // stage 4 for this warp's pair
  const ShiftK sk{p.kc[0], p.kc[1], p.kc[3], p.kc[4], p.kc[5], p.kc[6]};
  low_level_p<2, 1>(w_s, lane, k10, sk, warp);
  low_level_p<4, 1>(w_s, lane, k10, sk, warp);
  low_level_p<8, 1>(w_s, lane, k10, sk, warp);
  low_level_p<16, 1>(w_s, lane, k10, sk, warp);
  {
    uint32_t v[32];
    load_row<32>(wp, lane, v, k10);
    inverse_lift_p<32, kRaw>(v, lane < 16 ? pk(2048) : pk(2048 + kBias - kRaw), sk);
    store_row<32>(wp, lane, v);
  }
  __syncwarp();
  {
    uint32_t off[8];
    col_offsets(lane, off);
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = lds32(wp + off[i & 7] + i * 128);
    inverse_lift_p<32, kBias, true>(v, pk(2048), sk);
    const uint32_t kx = warp == 0 ? ph(128 + kBY) : warp == 1 ? ph(128 + kBO) : ph(128 + kBG);
#pragma unroll
    for (int i = 0; i < 32; ++i) sts32(wp + off[i & 7] + i * 128, (v[i] & 0x00FF00FFu) ^ kx);
  }
  __syncthreads();

  // stage 5 for this warp's slabs
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const uint32_t k = warp + 3 * i;
    if (k < 8) {
      const uint32_t src = wchunk(w_s, 4 * k + (lane >> 3), lane & 7);
      uint32_t Y[4], CO[4], CG[4];
      ld4(src, Y); ld4(src + kWPlane, CO); ld4(src + 2 * kWPlane, CG);
      uint32_t o[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        o[2 * j] = pack565_p(Y[j], CO[j], CG[j], sk);
        o[2 * j + 1] = words[i][j];
      }
      uint8_t *dst = p.out + (img_block0 + gidx0 + k * slab_stride) * 8;
      st_global_cs_v4(dst, o[0], o[1], o[2], o[3]);
      st_global_cs_v4(dst + 16, o[4], o[5], o[6], o[7]);
    }
  }
  if (__any_sync(0xffffffffu, seen > nmax) && lane == 0) atomicOr(p.status, 1u);
}

// ---------------------------------------------------------------------------------------
// Standalone decode of [u32 end_offset[n_groups]][groups] with 1..32 interleaved lanes and a
// single table: the `ans_decode` kernel of ans/ans_decode.cl:76-95 as driven by
// ans/ans_ocl.cpp:159-345.  Output: group * n_lanes * 256 + lane * 256 + position.
constexpr int kPlainWarps = 8;
constexpr int kPlainSmem = kPlainWarps * kRingSlot + kTableSize * 4 + kRingSlot;  // + alignment slack

__global__ void __launch_bounds__(kPlainWarps * 32)
    ans_decode_plain_kernel(const uint32_t *__restrict__ table, const uint8_t *__restrict__ data,
                            uint64_t data_bytes, uint32_t n_groups, uint32_t n_lanes,
                            uint8_t *__restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const RansSmem lay(smem_u32(smem));
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tab_s = lay.tab;
  load_table(tab_s, table, threadIdx.x, kPlainWarps * 32);
  __syncthreads();
  const uint32_t group = blockIdx.x * kPlainWarps + warp;
  if (group >= n_groups) return;
  uint8_t *dst = out + (static_cast<size_t>(group) * n_lanes + lane) * kSymsPerLane + 240;
  const bool active = lane < n_lanes;
  const uint32_t grp[1] = {group};
  const uint32_t ring[1] = {lay.ring(warp)};
  rans_decode_groups<false, 1, false>(tab_s, data, grp, n_lanes, ring, data, data + data_bytes,
                                      [&](int m, const uint32_t (&w)[4]) {
                                        if (active) *reinterpret_cast<uint4 *>(dst - 16 * m) = make_uint4(w[0], w[1], w[2], w[3]);
                                      });
}

// ---------------------------------------------------------------------------------------
// rANS ENCODE of one symbol stream, stream-compatible with ByteEncoder::EncodeBytes
// (codec/entropy.cpp:174-265) = ans::EncodeInterleaved (ans/encode.cpp:224-259) per group of 32 x 256
// symbols with rANS_Encoder::Encode (ans/encode.cpp:56-69): b = 2^16, k = 2^4, M = 2^11.
//   lane j of a warp owns stream j = symbols [256 j, 256 j + 256) of the group, in order; per symbol, streams
//   0..31 in order: while (state >= b k F[s]) emit (state & 0xFFFF), state >>= 16   (at most once: state < 2^31)
//   state = (state / F[s]) * M + B[s] + state % F[s];  the 32 final states follow the words.
// The emission order inside a step is the lane order, so a lane's word lands at
// base + popc(ballot & lanes_below); the decoder reads the same words backwards, higher lanes first.
// Output: the group's bytes (words then states) at scratch + group * kEncGroupCap, the byte count in
// sizes[group].  This is fixture tooling: it removes the CPU entropy coder
// from the path that builds large test sets, and is not part of the decode path.
constexpr int kEncGroupCap = 2 * kGroupSyms + 4 * kLanes;  // every symbol emits a word + the states

__global__ void __launch_bounds__(256) ans_encode_kernel(const uint8_t *__restrict__ symbols, uint32_t n_groups,
                                                         const uint16_t *__restrict__ freqs, uint8_t *__restrict__ scratch,
                                                         uint32_t *__restrict__ sizes) {
  __shared__ uint32_t s_f[256], s_b[256];
  __shared__ uint32_t s_warp[8];
  const uint32_t t = threadIdx.x, lane = t & 31, warp = t >> 5;
  {  // cumulative frequencies: exclusive scan of the 256 normalised frequencies
    const uint32_t f = freqs[t];
    uint32_t inc = f;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t n = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += n;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t base = 0;
    for (uint32_t w = 0; w < warp; ++w) base += s_warp[w];
    s_f[t] = f;
    s_b[t] = base + inc - f;
    __syncthreads();
  }
  const uint32_t group = blockIdx.x * 8 + warp;
  if (group >= n_groups) return;
  const uint8_t *src = symbols + static_cast<size_t>(group) * kGroupSyms + lane * kSymsPerLane;
  uint16_t *words = reinterpret_cast<uint16_t *>(scratch + static_cast<size_t>(group) * kEncGroupCap);
  const uint32_t below = (1u << lane) - 1u;
  uint32_t state = kRansL;  // k * M, ans/encode.cpp:47
  uint32_t n_words = 0;
#pragma unroll 1
  for (int i0 = 0; i0 < kSymsPerLane; i0 += 16) {
    const uint4 v = *reinterpret_cast<const uint4 *>(src + i0);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const uint32_t sym = (w[i >> 2] >> (8 * (i & 3))) & 0xFFu;
      const uint32_t F = s_f[sym], B = s_b[sym];
      // a symbol with F = 0 cannot be coded (the caller's histogram is wrong): treat it as F = 1 so the kernel
      // terminates; the host checks the histogram before launching
      const uint32_t Fs = F ? F : 1u;
      const bool emit = state >= (Fs << 20);  // b k F[s]; F <= 2048 keeps this inside 32 bits
      const uint32_t mask = __ballot_sync(0xffffffffu, emit);
      if (emit) {
        words[n_words + __popc(mask & below)] = static_cast<uint16_t>(state);
        state >>= 16;
      }
      n_words += __popc(mask);
      const uint32_t q = state / Fs;
      state = q * kTableSize + B + (state - q * Fs);
    }
  }
  words[n_words + 2 * lane] = static_cast<uint16_t>(state);  // (the states are only 2-byte aligned here)
  words[n_words + 2 * lane + 1] = static_cast<uint16_t>(state >> 16);
  if (lane == 0) sizes[group] = 2 * n_words + 4 * kLanes;
}

// Gathers the groups into the stream ByteEncoder::EncodeBytes emits after its 512-byte frequency block:
// [u32 end_offset[groups]][group 0][group 1]..., every group padded AT THE FRONT with two zero bytes when
// its size is not a multiple of four (codec/entropy.cpp:213-228).  offsets[g] = end of group g.
__global__ void __launch_bounds__(256) ans_encode_gather_kernel(const uint8_t *__restrict__ scratch,
                                                                const uint32_t *__restrict__ sizes,
                                                                const uint32_t *__restrict__ offsets, uint32_t n_groups,
                                                                uint8_t *__restrict__ out) {
  const uint32_t g = blockIdx.x;
  const uint32_t size = sizes[g], end = offsets[g];
  const uint32_t padded = (size + 3u) & ~3u;
  uint16_t *dst = reinterpret_cast<uint16_t *>(out + end - padded);
  const uint16_t *src = reinterpret_cast<const uint16_t *>(scratch + static_cast<size_t>(g) * kEncGroupCap);
  const uint32_t pad16 = (padded - size) / 2;  // 0 or 1 leading zero word
  if (threadIdx.x == 0 && pad16) dst[0] = 0;
  for (uint32_t i = threadIdx.x; i < size / 2; i += blockDim.x) dst[pad16 + i] = src[i];
  if (threadIdx.x == 0) reinterpret_cast<uint32_t *>(out)[g] = end;
}

}  // namespace

// ---------------------------------------------------------------------------------------
// launchers
cudaError_t launch_build_tables(const uint8_t *freqs, uint32_t n_tables, uint32_t *tables, cudaStream_t s) {
  if (n_tables == 0) return cudaSuccess;
  build_tables_kernel<<<n_tables, 256, 0, s>>>(freqs, tables);
  return cudaGetLastError();
}

static cudaError_t ensure_attrs() {
  static cudaError_t once = []() {
    cudaError_t e = cudaSuccess;
    for (auto *k : {wavelet_assemble_kernel<0, false, true>, wavelet_assemble_kernel<1, false, true>, wavelet_assemble_kernel<0, true, true>,
                    wavelet_assemble_kernel<1, true, true>, wavelet_assemble_kernel<0, false, false>, wavelet_assemble_kernel<1, false, false>,
                    wavelet_assemble_kernel<0, true, false>, wavelet_assemble_kernel<1, true, false>}) {
      if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kWaSmem)) != cudaSuccess) return e;
    }
    for (auto *k : {rans_streams_kernel<false, false, false>, rans_streams_kernel<true, false, false>,
                    rans_streams_kernel<false, true, false>, rans_streams_kernel<true, true, false>,
                    rans_streams_kernel<false, true, true>, rans_streams_kernel<true, true, true>}) {
      if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, RansCfg::kSmem)) != cudaSuccess) return e;
    }
    return cudaFuncSetAttribute(ans_decode_plain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPlainSmem);
  }();
  return once;
}

// A launch that may start while the previous kernel of the stream is still draining (programmatic dependent
// launch): the kernel runs its prologue and blocks in griddepcontrol.wait until that kernel has completed.
template <class... KArgs, class... Args>
static cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

bool is_small_call(uint32_t n_images, uint32_t groups_per_plane, uint32_t max_palette_bytes) {
  return static_cast<uint64_t>(n_images) * (7ull * groups_per_plane + max_palette_bytes / kGroupSyms) <= kSmallCallGroups;
}

cudaError_t launch_decode_batch(const BatchParams &p_in, int rgb_mode, uint32_t max_palette_bytes,
                                cudaStream_t s, cudaEvent_t *marks) {
  if (p_in.n_images == 0) return cudaSuccess;
  BatchParams p = p_in;
  cudaError_t e = ensure_attrs();
  if (e != cudaSuccess) return e;
  int mark = 0;
  auto stamp = [&]() { return marks ? cudaEventRecord(marks[mark++], s) : cudaSuccess; };
  const bool pdl = marks == nullptr;  // (an event between two kernels serialises them anyway)
  if ((e = stamp()) != cudaSuccess) return e;
  // A call whose rANS groups do not fill the machine is bound by the latency of one group (256 dependent decode
  // steps) plus whatever precedes it: the consuming CTAs build their tables themselves, two launches.  Everything
  // else: tables built once per stream by their own kernel, three launches.
  const bool small = is_small_call(p.n_images, p.groups_per_plane, max_palette_bytes);
  if (!small) {
    // stage 1: 4 tables per image, straight from the freq region of the compressed buffer
    build_tables_batch_kernel<<<4 * p.n_images, 256, 0, s>>>(p);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  if ((e = stamp()) != cudaSuccess) return e;
  // stage 2 (+ the group-local part of stage 3): every rANS group of the batch
  const uint32_t per_cta = RansCfg::kGroupsPerCta;
  StreamGrid sg;
  sg.y_ctas = (2 * p.groups_per_plane + per_cta - 1) / per_cta;
  sg.c_ctas = (4 * p.groups_per_plane + per_cta - 1) / per_cta;
  sg.pal_ctas = (max_palette_bytes / kGroupSyms + per_cta - 1) / per_cta;
  sg.idx_ctas = (p.groups_per_plane + per_cta - 1) / per_cta;
  p.rans_ctas = sg.per_image();
  if (p.n_images == 1) p.img_done = nullptr;  // no hand-over for a lone image
  // the stage taps (parity tests only) are a separate instantiation: the production kernels carry none of that code
  const bool taps = p.tap_symbols || p.tap_planes || p.tap_indices;
  const dim3 rans_grid(p.n_images * sg.per_image()), rans_block(kRansWarps * 32);
  if (small && rans_grid.x <= 3u * 148u) {
    e = taps ? launch_kernel(rans_streams_kernel<true, true, true>, rans_grid, rans_block, RansCfg::kSmem, s, false, p, sg)
             : launch_kernel(rans_streams_kernel<false, true, true>, rans_grid, rans_block, RansCfg::kSmem, s, false, p, sg);
  } else if (small) {
    e = taps ? launch_kernel(rans_streams_kernel<true, true, false>, rans_grid, rans_block, RansCfg::kSmem, s, false, p, sg)
             : launch_kernel(rans_streams_kernel<false, true, false>, rans_grid, rans_block, RansCfg::kSmem, s, false, p, sg);
  } else {
    e = taps ? launch_kernel(rans_streams_kernel<true, false, false>, rans_grid, rans_block, RansCfg::kSmem, s, pdl, p, sg)
             : launch_kernel(rans_streams_kernel<false, false, false>, rans_grid, rans_block, RansCfg::kSmem, s, pdl, p, sg);
  }
  if (e != cudaSuccess) return e;
  if ((e = stamp()) != cudaSuccess) return e;
  // stages 4 + 5 (and the cross-group carry of stage 3): one warp per tile
  if (p.n_images > 65535u) return cudaErrorInvalidValue;  // grid.y; gst_capi.cu pages larger batches
  const dim3 grid((p.n_blocks / kTileSyms + kWaWarps - 1) / kWaWarps, p.n_images), block(kWaWarps * 32);
  auto launch = [&](auto kern) { return launch_kernel(kern, grid, block, kWaSmem, s, pdl, p); };
  if (!rgb_mode && !taps && static_cast<uint64_t>(p.n_images) * (p.n_blocks / kTileSyms) <= kSplitMaxTiles) {
    // so few tiles that every one can have a CTA of three warps to itself, all resident at once
    const dim3 sgrid(p.n_blocks / kTileSyms, p.n_images), sblock(kSplitWarps * 32);
    e = p.idx16 ? launch_kernel(wavelet_assemble_split_kernel<true>, sgrid, sblock, kWarpWork, s, pdl, p)
                : launch_kernel(wavelet_assemble_split_kernel<false>, sgrid, sblock, kWarpWork, s, pdl, p);
    if (e != cudaSuccess) return e;
    return stamp();
  }
  const int variant = (rgb_mode ? 1 : 0) | (taps ? 2 : 0) | (p.idx16 ? 4 : 0);
  switch (variant) {
    case 0: e = launch(wavelet_assemble_kernel<0, false, false>); break;
    case 1: e = launch(wavelet_assemble_kernel<1, false, false>); break;
    case 2: e = launch(wavelet_assemble_kernel<0, true, false>); break;
    case 3: e = launch(wavelet_assemble_kernel<1, true, false>); break;
    case 4: e = launch(wavelet_assemble_kernel<0, false, true>); break;
    case 5: e = launch(wavelet_assemble_kernel<1, false, true>); break;
    case 6: e = launch(wavelet_assemble_kernel<0, true, true>); break;
    default: e = launch(wavelet_assemble_kernel<1, true, true>); break;
  }
  if (e != cudaSuccess) return e;
  return stamp();
}

cudaError_t launch_ans_encode(const uint8_t *symbols, uint32_t n_groups, const uint16_t *freqs, uint8_t *scratch,
                              uint32_t *sizes, cudaStream_t s) {
  if (n_groups == 0) return cudaSuccess;
  ans_encode_kernel<<<(n_groups + 7) / 8, 256, 0, s>>>(symbols, n_groups, freqs, scratch, sizes);
  return cudaGetLastError();
}

cudaError_t launch_ans_encode_gather(const uint8_t *scratch, const uint32_t *sizes, const uint32_t *offsets,
                                     uint32_t n_groups, uint8_t *out, cudaStream_t s) {
  if (n_groups == 0) return cudaSuccess;
  ans_encode_gather_kernel<<<n_groups, 256, 0, s>>>(scratch, sizes, offsets, n_groups, out);
  return cudaGetLastError();
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
