/* gst_cuda.h -- C ABI of libgst_cuda.so, the B200 (sm_100a) replacement for the reference's
 * OpenCL runtime layer on the .gst -> DXT1 decode path.
 *
 * Every entry point names the reference interface it replaces (paths relative to the
 * GammaUNC/GST tree).  Plain pointers and sizes only: device buffers are raw CUDA device
 * pointers (void*), streams and events are the CUDA handles cast to void* (cudaStream_t /
 * cudaEvent_t), so a host written in any language can bind this with its FFI.
 *
 * All functions returning int return GST_OK (0) or a negative gst_status; the message of
 * the last failure on the calling thread is available from gst_last_error().  There is no
 * CPU fallback: without a CUDA device gst_ctx_create fails with GST_ERR_NO_DEVICE.
 */
#ifndef GST_CUDA_H_
#define GST_CUDA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum gst_status {
  GST_OK = 0,
  GST_ERR_INVALID = -1,    /* bad argument / malformed header or container        */
  GST_ERR_NO_DEVICE = -2,  /* no usable CUDA device                                */
  GST_ERR_CUDA = -3,       /* a CUDA runtime call failed (see gst_last_error)      */
  GST_ERR_NOMEM = -4,      /* device allocation / preallocated arena exhausted     */
  GST_ERR_SMALL = -5       /* caller-provided output buffer too small              */
} gst_status;

/* codec/codec_base.h:9-20 GenTCHeader -- 7 x u32, raw little-endian memcpy
 * (codec/codec_base.cpp:18-24). */
typedef struct gst_header {
  uint32_t width;
  uint32_t height;
  uint32_t palette_bytes;
  uint32_t y_cmp_sz;
  uint32_t chroma_cmp_sz;
  uint32_t palette_sz;
  uint32_t indices_sz;
} gst_header;

/* ans/ans.h:72-75 and codec/codec_base.h:22 */
#define GST_ANS_TABLE_SIZE 2048u
#define GST_ANS_SYMBOLS_PER_LANE 256u
#define GST_ANS_LANES_PER_GROUP 32u
#define GST_WAVELET_BLOCK_DIM 32u
#define GST_HEADER_BYTES 28u

typedef struct gst_ctx gst_ctx;

const char *gst_last_error(void);

/* ---- context: replaces gpu::GPUContext::InitializeOpenCL (gpu/gpu.cpp:394-529) and
 * GenTC::InitializeDecoder (codec/decoder.cpp:558-591).  One context per GPU / process:
 * device ordinal, 1 default + 4 work streams (gpu/gpu.h:49 kMaxNumWorkQueues), scratch
 * arena.  Fails unless the device is compute capability 10.x. */
int gst_ctx_create(int device, gst_ctx **out);
void gst_ctx_destroy(gst_ctx *ctx);
int gst_ctx_device(const gst_ctx *ctx);
/* gpu/gpu.h:53 GetDefaultCommandQueue, :56-59 GetNextQueue (atomic round-robin) */
void *gst_stream_default(gst_ctx *ctx);
void *gst_stream_next(gst_ctx *ctx);
/* gpu/gpu.h:61-66 FlushAllQueues -- here: synchronise all five streams */
int gst_ctx_sync(gst_ctx *ctx);
int gst_stream_sync(gst_ctx *ctx, void *stream);

/* ---- buffers: replace clCreateBuffer / clEnqueueWriteBuffer / clEnqueueReadBuffer as used
 * by codec/decoder.cpp:430-523 and demo/photos_sf.cpp:798-806.  Sizes are rounded up to
 * 256 B.  Host pointers passed to the async copies should be pinned (gst_host_alloc). */
int gst_malloc(gst_ctx *ctx, size_t bytes, void **dptr);
int gst_free(gst_ctx *ctx, void *dptr);
int gst_host_alloc(gst_ctx *ctx, size_t bytes, void **hptr);
int gst_host_free(gst_ctx *ctx, void *hptr);
int gst_upload_async(gst_ctx *ctx, void *stream, void *dst_dev, const void *src_host, size_t bytes);
int gst_download_async(gst_ctx *ctx, void *stream, void *dst_host, const void *src_dev, size_t bytes);
/* strided read-back (rows of width_bytes): e.g. one block of every texture of a batch */
int gst_download_2d_async(gst_ctx *ctx, void *stream, void *dst_host, size_t dst_pitch,
                          const void *src_dev, size_t src_pitch, size_t width_bytes, size_t rows);
int gst_memset_async(gst_ctx *ctx, void *stream, void *dst_dev, int value, size_t bytes);

/* ---- events: replace cl_event hand-off (codec/decoder.h:19-33).  The caller owns every
 * event returned to it and must destroy it (demo/demo.cpp:224-228). */
int gst_event_record(gst_ctx *ctx, void *stream, void **event_out);
int gst_event_wait(void *event);
int gst_event_elapsed_ms(void *start, void *stop, float *ms);
void gst_event_destroy(void *event);

/* ---- stream format helpers (host only).
 * gst_parse_header: GenTCHeader::LoadFrom (codec/codec_base.cpp:18-24) + the size checks the
 * reference only asserts (codec/encoder.cpp:41-42, codec/decoder.cpp:150,243-244).
 * gst_packed_size / gst_pack_batch: build the device input buffer LoadCompressedDXTs
 * expects, exactly as demo/photos_sf.cpp:753-795 and UploadData (codec/decoder.cpp:430-476):
 *   [u32 out_off[4n]][u32 in_off[4n]] padded to 512 B | n x 2048 B freqs | n payloads.
 * All images of a batch must share width/height (codec/decoder.cpp:117-121). */
int gst_parse_header(const uint8_t *gst, size_t len, gst_header *hdr);
size_t gst_packed_size(const gst_header *hdrs, uint32_t n);
int gst_pack_batch(const uint8_t *const *gst_files, const size_t *lens, uint32_t n,
                   uint8_t *dst, size_t dst_cap, gst_header *hdrs_out);

/* ---- scratch: GenTC::RequiredScratchMem / PreallocateDecompressor / FreeDecompressor
 * (codec/decoder.cpp:41-47,478-485).  gst_required_scratch returns the reference's figure
 * (4*2048*6 + 17N + P), which upper-bounds what this implementation uses (32 KB + 4N + P +
 * 8N/8192), so callers that size by it keep working.  With a preallocated arena regions are
 * bump-allocated and never reset until gst_free_scratch (codec/decoder.cpp:74-82); without
 * one the calls made on a stream share one grow-only buffer the context keeps for that stream
 * (released by gst_free_scratch / gst_ctx_destroy): a call in steady state allocates nothing. */
size_t gst_required_scratch(const gst_header *hdr);
int gst_preallocate(gst_ctx *ctx, size_t bytes);
int gst_free_scratch(gst_ctx *ctx);

/* ---- the decode path.
 * gst_load_dxt_batch  = GenTC::LoadCompressedDXTs (codec/decoder.h:23-25, decoder.cpp:540-544);
 *                       n == 1 is GenTC::LoadCompressedDXT (decoder.h:19-21).
 * gst_load_rgb_batch  = GenTC::LoadRGBs / LoadRGB (codec/decoder.h:27-33).
 * cmp_dev: device buffer in the packed layout above; out_dev: n*W*H/2 bytes of DXT1 blocks
 * (codec/dxt_image.h:14-21) or n*W*H*3 bytes of RGB8.  Asynchronous: the work is ordered
 * after wait_events[0..n_wait) on `stream`; if done_event is non-NULL a new event recorded
 * after the last kernel is returned in it. */
int gst_load_dxt_batch(gst_ctx *ctx, const gst_header *hdrs, uint32_t n, void *stream,
                       const void *cmp_dev, size_t cmp_bytes, void *out_dev,
                       void *const *wait_events, uint32_t n_wait, void **done_event);
int gst_load_rgb_batch(gst_ctx *ctx, const gst_header *hdrs, uint32_t n, void *stream,
                       const void *cmp_dev, size_t cmp_bytes, void *out_dev,
                       void *const *wait_events, uint32_t n_wait, void **done_event);

/* GenTC::DecompressDXT (codec/decoder.h:16-17, decoder.cpp:487-532): host .gst bytes in,
 * host DXT1 blocks out (W*H/2 bytes), blocking.  mode 0 = DXT1, 1 = RGB8 (W*H*3 bytes). */
int gst_decompress_host(gst_ctx *ctx, const uint8_t *gst, size_t len, int mode, uint8_t *out,
                        size_t out_cap);

/* Batched host-to-host convenience used by the end-to-end benchmark: n .gst files with the
 * same dimensions, pinned host staging, H2D, decode, D2H, pipelined over the context's work
 * streams in pages of `page` images (demo/photos_sf.cpp:688 kPageSize = 16; 0 = one page). */
int gst_decompress_host_batch(gst_ctx *ctx, const uint8_t *const *gst_files, const size_t *lens,
                              uint32_t n, uint32_t page, int mode, uint8_t *out, size_t out_cap);

/* The headless form of the photos_sf loader (demo/photos_sf.cpp:688-885): n .gst files in host
 * memory are packed page by page into pinned staging on the context's worker threads, copied
 * to the device and decoded with LoadCompressedDXTs semantics straight into the caller's DEVICE
 * buffer out_dev (image i at i * W*H/2 bytes, or i * W*H*3 in RGB mode) -- where the reference
 * hands the pages to a GL pixel buffer.  Blocking; pages overlap across the work streams. */
int gst_load_host_batch(gst_ctx *ctx, const uint8_t *const *gst_files, const size_t *lens,
                        uint32_t n, uint32_t page, int mode, void *out_dev, size_t out_cap);

/* Upload policy of the two host batch entry points above.  Off (default): every file is copied into pinned staging
 * and a page goes up in one DMA.  On: a file that lies in pinned memory is DMA-ed from where it lies (one copy per
 * file, no host-side packing pass); files in pageable memory are still staged. */
int gst_ctx_set_direct_upload(gst_ctx *ctx, int on);

/* ---- frame streamer: the headless form of the demo player (demo/demo.cpp:145-243,504-600), which
 * reads frameNNNN.gtc, uploads it, calls LoadCompressedDXT / LoadRGB and blocks on the event
 * before the next frame.  Here `depth` frames are in flight: submit() copies the frame into the
 * slot's pinned staging, uploads it and decodes it on the slot's own stream; it only
 * blocks when the slot's previous frame (ticket - depth) is still running.  mode 0 = DXT1, 1 = RGB8
 * (LoadRGB, demo/demo.cpp:208-212).  out_dev NULL decodes into the slot's own device frame.
 * wait() blocks until that frame is complete and returns where it was decoded to; a frame stays
 * valid until `depth` further frames have been submitted. */
typedef struct gst_streamer gst_streamer;
int gst_streamer_create(gst_ctx *ctx, uint32_t width, uint32_t height, uint32_t depth, int mode,
                        gst_streamer **out);
int gst_streamer_submit(gst_streamer *st, const uint8_t *gst, size_t len, void *out_dev,
                        uint64_t *ticket);
/* submit with options.  out_host (may be NULL): the decoded frame is also copied to this host buffer (pinned for full
 * speed) on the slot's own stream, and wait() returns when the copy has landed -- the demo's
 * "decode, then hand the frame on" (demo/demo.cpp:221-237) without a separate download the caller has to order
 * against the slot's next frame.  GST_SUBMIT_DIRECT: no copy into the slot's staging -- the upload reads the
 * caller's .gst buffer, which must stay valid (and should be pinned) until the frame has been waited for. */
#define GST_SUBMIT_DIRECT 1u
int gst_streamer_submit_ex(gst_streamer *st, const uint8_t *gst, size_t len, void *out_dev, void *out_host,
                           uint32_t flags, uint64_t *ticket);
int gst_streamer_wait(gst_streamer *st, uint64_t ticket, void **frame_dev);
/* The player's main loop (demo/demo.cpp:504-600) over n frames already in host memory.  out_dev / out_host, when not
 * NULL, receive frame f at f decoded frames from their start.  Frames are decoded GST_PLAY_GROUP(k) at a time (default
 * 8): one LoadCompressedDXTs-style call and one read-back per group, `depth` groups in flight; GST_SUBMIT_DIRECT as
 * for gst_streamer_submit_ex (it pays when the frames stay on the device; with a read-back the staged upload is
 * faster).  Returns when every frame is done; frames played this way have no tickets. */
#define GST_PLAY_GROUP(k) (((uint32_t)(k) & 0xFFu) << 8)
int gst_streamer_play(gst_streamer *st, const uint8_t *const *frames, const size_t *lens, uint32_t n,
                      void *out_dev, void *out_host, uint32_t flags);
void gst_streamer_destroy(gst_streamer *st);

/* ---- stream sanity flags.  The reference does not validate stream contents (malformed input is undefined behaviour,
 * SURVEY.md section 5); here every access is clamped, and a decode that had to clamp says so: the kernels OR
 *   GST_FLAG_INDEX_CLAMPED   a palette index lay beyond the image's palette (it was clamped to the last entry)
 *   GST_FLAG_PALETTE_RANGE   an image's palette region did not fit its batch (its blocks got index word 0)
 *   GST_FLAG_SYNC_TIMEOUT    internal: a tile waited more than a second for its image's entropy decode (never
 *                            expected; the output of that call is not to be trusted)
 * into one word per context.  gst_status_flags reads it (and clears it when `clear` is non-zero); call it after the
 * work of interest has completed (gst_stream_sync / an event). */
#define GST_FLAG_INDEX_CLAMPED 1u
#define GST_FLAG_PALETTE_RANGE 2u
#define GST_FLAG_SYNC_TIMEOUT 4u
int gst_status_flags(gst_ctx *ctx, uint32_t *flags, int clear);

/* ---- stage taps for the parity tests (not used in production).  Any pointer may be NULL.
 *   symbols_dev : sum(7N + P) bytes, reference decmp_buf layout (codec/decoder.cpp:212)
 *   planes_dev  : n*6N int8, raster planes (codec/decoder.cpp:280)
 *   indices_dev : n*N int32 palette indices (codec/decoder.cpp:302) */
int gst_load_dxt_batch_tapped(gst_ctx *ctx, const gst_header *hdrs, uint32_t n, void *stream,
                              const void *cmp_dev, size_t cmp_bytes, void *out_dev,
                              void *symbols_dev, void *planes_dev, void *indices_dev);

/* ---- standalone rANS decoder: ans::ocl::OpenCLDecoder (ans/ans_ocl.h:26-72).
 * gst_ans_create(F, n, lanes)  = OpenCLDecoder(ctx, F, num_interleaved) (ans_ocl.cpp:60-98);
 *   F are symbol counts, normalised internally to sum 2048 exactly like
 *   ans::ocl::NormalizeFrequencies (ans/ans_ocl_encode.cpp:6-8, ans/histogram.cpp:41-123).
 * gst_ans_rebuild              = RebuildTable (ans_ocl.cpp:100-148)
 * gst_ans_table                = GetSymbols / GetFrequencies / GetCumulativeFrequencies
 *                                (ans_ocl.cpp:349-427); each output holds 2048 entries.
 * gst_ans_decode               = the three Decode overloads (ans_ocl.cpp:155-345): `groups`
 *   independent groups, each `lanes` (<= the decoder's lanes) interleaved streams of 256
 *   symbols; states[g*lanes+l]
 *   is the final encoder state of lane l, data[g]/data_len[g] the group's renorm bytes as
 *   written by the encoder.  out receives groups*lanes*256 symbols, out[(g*lanes+l)*256+i]. */
typedef struct gst_ans_decoder gst_ans_decoder;
/* ans::GenerateHistogram(counts, target_sum) (ans/histogram.cpp:41-123); target_sum 0 means
 * 2048 = ans::ocl::NormalizeFrequencies.  Host only. */
int gst_normalize_frequencies(const uint32_t *counts, uint32_t n, uint32_t target_sum,
                              uint32_t *out);
int gst_ans_create(gst_ctx *ctx, const uint32_t *F, uint32_t n, uint32_t lanes,
                   gst_ans_decoder **out);
int gst_ans_rebuild(gst_ans_decoder *d, const uint32_t *F, uint32_t n);
int gst_ans_table(gst_ans_decoder *d, uint8_t *symbols, uint16_t *freqs, uint16_t *cum_freqs);
int gst_ans_decode(gst_ans_decoder *d, uint32_t lanes, const uint32_t *states,
                   const uint8_t *const *data, const size_t *data_len, uint32_t groups,
                   uint8_t *out);
void gst_ans_destroy(gst_ans_decoder *d);

/* ---- rANS stream ENCODER on the GPU: ByteEncoder::EncodeBytes (codec/entropy.cpp:174-265), i.e.
 * ans::ocl::NormalizeFrequencies of the byte histogram + ans::EncodeInterleaved (ans/encode.cpp:224-259,
 * rANS_Encoder::Encode ans/encode.cpp:56-69) per group of 32 x 256 symbols.  Fixture tooling (it lets
 * large test sets be entropy-coded without the CPU encoder); not on the decode path.
 *   symbols, n_symbols : host bytes, n_symbols a multiple of 8192
 *   freqs_out          : 256 x u16, the normalised frequencies = the 512-byte block of the .gst file
 *   stream_out         : [u32 end_offset[groups]][groups...] padded to 4 bytes, exactly the bytes
 *                        EncodeBytes emits after the frequency block; *stream_bytes receives the size.
 *                        stream_cap >= gst_ans_encode_bound(n_symbols) always suffices. */
size_t gst_ans_encode_bound(size_t n_symbols);
int gst_ans_encode_stream(gst_ctx *ctx, const uint8_t *symbols, size_t n_symbols, uint16_t *freqs_out,
                          uint8_t *stream_out, size_t stream_cap, size_t *stream_bytes);

/* stage 1 on caller buffers: n tables of 256 x u16 frequencies (512 B each, as stored in
 * the .gst file) -> n x 2048 packed u32 entries
 * freq | sym << 12 | (slot - cum_freq) << 20 (gst_kernels.cuh). */
int gst_build_tables(gst_ctx *ctx, void *stream, const void *freqs_dev, uint32_t n_tables,
                     void *tables_dev);

/* number of kernel launches one gst_load_*_batch call enqueues, in launch order:
 * build_tables, rans_streams (all four streams + the group-local index scan), wavelet_assemble.
 * gst_launches_for_batch: the number for these headers -- 2 when the call is too small to fill the GPU (at most
 * 16384 rANS groups): the consuming CTAs build their tables themselves. */
int gst_launches_per_batch(void);
int gst_launches_for_batch(const gst_header *hdrs, uint32_t n);

/* Per-kernel device timing for the benchmark's roofline line (no reference equivalent; the
 * reference only has wall-clock prints, demo/photos_sf.cpp:851,892-893).  While enabled,
 * every gst_load_*_batch call records CUDA events on its stream around each of its kernels.
 * gst_profile_read waits for the recorded calls, adds the elapsed milliseconds of kernel k
 * over all calls since the last read into kernel_ms[k] and returns the number of calls. */
int gst_profile_enable(gst_ctx *ctx, int on);
int gst_profile_read(gst_ctx *ctx, double *kernel_ms, uint32_t n_kernels, uint64_t *calls);

#ifdef __cplusplus
}
#endif
#endif /* GST_CUDA_H_ */
