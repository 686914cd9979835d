/*
 * tinyimgcodec_cuda.h — C ABI of libtinyimgcodec_cuda.so, the B200 (sm_100a) encode path and the decode
 * side that follows it.
 *
 * The reference (clysto/tinyimgcodec) has no FFI/plugin layer: its boundary for the
 * encode hot path is two Python functions,
 *     tinyimgcodec.codec.compress(image, quality=50, auto_generate_huffman_table=False) -> bytes
 *         (tinyimgcodec/codec.py:133-164)
 *     tinyimgcodec.codec.encode(image, quality=50) -> dict
 *         (tinyimgcodec/codec.py:26-43)
 * and the CLI ./encode.py (encode.py:10-19); for the decode side (second half of this header)
 *     tinyimgcodec.codec.decompress(data) -> uint8 H x W      (tinyimgcodec/codec.py:167-189)
 *     tinyimgcodec.codec.decode(data: dict) -> uint8 H x W    (tinyimgcodec/codec.py:46-70).
 * This header is what a ctypes binding placed
 * UNDER those functions calls (INTEGRATION.md shows the stub).  Every entry point
 * returns 0 or a negative TIC_E_* code; no C++ exception, torch type or CUDA type
 * crosses the boundary (streams are passed as void*, i.e. a cudaStream_t value).
 * The caller owns every buffer; the library owns only the opaque handle and the
 * scratch memory behind it.  One handle per GPU; calls on one handle are serialised
 * by the caller (a handle is not thread-safe), different handles run concurrently.
 *
 * There is no CPU fallback: every compute entry point fails with TIC_E_CUDA when no
 * sm_100-class device is usable.
 */
#ifndef TINYIMGCODEC_CUDA_H
#define TINYIMGCODEC_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TIC_OK 0
#define TIC_E_INVALID (-1)    /* bad argument (NULL pointer, negative size, n_images < 0 ...) */
#define TIC_E_CUDA (-2)       /* a CUDA runtime call failed; see tic_last_error() */
#define TIC_E_QUALITY (-3)    /* quality outside 1..99: the reference raises ZeroDivisionError
                                 for 0 (tinyimgcodec/utils.py:50) and KeyError for 100 */
#define TIC_E_CAPACITY (-4)   /* output buffer too small; nothing past out_capacity was written */
#define TIC_E_CATEGORY (-5)   /* a DC size >= 12 or AC size >= 11 met the fixed tables: the
                                 reference raises KeyError (tinyimgcodec/huffman.py:62);
                                 per-image detail is in the status array */
#define TIC_E_UNSUPPORTED (-6) /* auto-table code longer than 32 bits (device limit; the reference
                                 allows up to 255) */
#define TIC_E_TABLE (-7)      /* auto table not serialisable: the reference raises OverflowError from
                                 int2ba (tinyimgcodec/codec.py:76-77,81-83) */

#define TIC_E_STREAM (-8)     /* decode side: at least one stream of the batch is damaged, truncated or has a
                                 header that does not match the caller's dimensions; per-stream detail
                                 (TIC_DSTATUS_*) is in the status array.  The reference swallows these errors
                                 block by block (try/except, tinyimgcodec/codec.py:177-185). */

/* flags for tic_encode_batch */
#define TIC_FLAG_AUTO_HUFFMAN 1u /* per-image tables, tinyimgcodec/codec.py:146-148 */
#define TIC_FLAG_C_VARIANT 2u    /* the stream of the reference's embedded C encoder (c/encode.c, c/img.c:
                                    header flag bit 30, integer AAN FDCT, scaled integer quantiser, one flush
                                    byte).  `quality` is then IMG_Q_BEST..IMG_Q_LOW = 0..3 (c/img.h:22); width
                                    and height must be multiples of 8 (c/encode.c:38-41).  Byte-identical to
                                    the reference binary for every block of the image; the extra block row that
                                    binary appends (c/encode.c:47: one more loop pass at EOF over a stack buffer
                                    the C library has overwritten — its bits differ from run to run of the
                                    reference itself) is not produced.  Not combinable with
                                    TIC_FLAG_AUTO_HUFFMAN. */

#define TIC_FLAG_AUTO_LE_FLAG 4u  /* opt-in, with TIC_FLAG_AUTO_HUFFMAN only: write the header's flag word as the
                                    little-endian bytes 00 00 00 80.  The reference writes it MSB-first
                                    (80 00 00 00, tinyimgcodec/codec.py:111) but reads it back with struct "I"
                                    (codec.py:119), so its own decoder cannot open its auto-table streams; with
                                    this flag it can.  Off by default: byte parity with compress() comes first. */

#define TIC_FLAG_DEBUG_ALL_EXACT 8u /* test hook (tests/test_gpu_parity.py): EVERY coefficient is recomputed by the
                                    float64 exact path, whether the tie guard flagged it or not, and
                                    tic_last_guard_misses() counts the coefficients the exact path changed although
                                    the guard had not flagged them — the guard band of the fast transform
                                    (tensor-core f16 split GEMM, or FP32 butterflies) is sound iff that count is
                                    0.  Same streams, many times slower.  Match: tinyimgcodec/utils.py:32-37,53. */

/* per-image status bits written by the device */
#define TIC_STATUS_CATEGORY 1 /* KeyError case above */
#define TIC_STATUS_TABLE 2    /* auto table not serialisable (OverflowError in the reference,
                                 tinyimgcodec/codec.py:76-77,81-83) */
#define TIC_STATUS_LONGCODE 4 /* auto-table code longer than 32 bits: not supported on the device */

typedef struct tic_handle_s *tic_handle;

/* Library version string, e.g. "tinyimgcodec_cuda 0.1 sm_100a". */
const char *tic_version(void);

/* Create / destroy the per-GPU context (constant tables and a workspace that grows on demand and is
 * reused across calls: per-tile records and positions, and the arena the tiles' bits pass through —
 * never more than min(worst case, out_capacity) + 16 bytes per 128 blocks).  Replaces nothing in the
 * reference (which has no state); it is the price of a device. */
int tic_create(int device, tic_handle *out);
int tic_destroy(tic_handle h);

/* Text of the last error on this handle (never NULL; valid until the next call). */
const char *tic_last_error(tic_handle h);

/* Upper bound on the bytes tic_encode_batch writes for one H x W image with the fixed
 * tables: a 16-byte header (tinyimgcodec/codec.py:102-114), at most 1662 bits per 8x8
 * block (DC 9+11, 63 x (16+10), EOB 4), plus 16 bytes of alignment slack.  With
 * TIC_FLAG_AUTO_HUFFMAN add 1664 bytes for the serialised tables (a stream larger than that is
 * reported as TIC_E_CAPACITY, never written out of bounds).  Pure host arithmetic; usable
 * without a GPU. */
int64_t tic_max_out_bytes(int32_t height, int32_t width);

/* Number of 8x8 blocks of the padded image (ceil(H/8)*ceil(W/8), tinyimgcodec/utils.py:56-61). */
int64_t tic_num_blocks(int32_t height, int32_t width);

/*
 * compress() for a batch of images resident in device memory — replaces the whole body
 * of tinyimgcodec/codec.py:133-164 (encode, run-length, header, Huffman, to_bytes) for
 * n_images independent images in one launch sequence.
 *
 *   d_pixels      n_images device pointers... passed as a HOST array of device addresses;
 *                 image i is heights[i] x widths[i] uint8, row-major, contiguous.
 *   heights/widths HOST arrays, original (unpadded) dimensions; 0 is allowed.
 *   quality       1..99, the same for the whole batch (0..3 = IMG_Q_BEST..IMG_Q_LOW with
 *                 TIC_FLAG_C_VARIANT).
 *   flags         TIC_FLAG_* bits.
 *   d_out         device buffer of out_capacity bytes.  The n streams are written densely,
 *                 each starting on a 16-byte boundary, in image order.
 *   d_out_offsets device int64[n_images]: byte offset of stream i in d_out.
 *   d_out_sizes   device int64[n_images]: byte length of stream i (== len(compress(...))).
 *   d_status      device int32[n_images]: TIC_STATUS_* bits (0 = identical to reference).
 *   stream        cudaStream_t the work is enqueued on (NULL = default stream).
 *
 * Asynchronous with respect to the host: results are valid after the stream is
 * synchronised.  Errors that only the device can see (capacity, category) are reported
 * by tic_encode_finish().
 */
int tic_encode_batch(tic_handle h, const void *const *d_pixels, const int32_t *heights,
                     const int32_t *widths, int32_t n_images, int32_t quality, uint32_t flags,
                     void *d_out, int64_t out_capacity, int64_t *d_out_offsets,
                     int64_t *d_out_sizes, int32_t *d_status, void *stream);

/* Synchronise `stream`, and report what the device saw during the last tic_encode_batch
 * on this handle: TIC_OK, TIC_E_CAPACITY or TIC_E_CATEGORY.  *total_bytes (optional)
 * receives the number of bytes of d_out that are in use (end of the last stream). */
int tic_encode_finish(tic_handle h, void *stream, int64_t *total_bytes);

/*
 * encode() for one image resident in device memory — replaces tinyimgcodec/codec.py:28-36
 * (pad, level shift, block split, FDCT, quantise, zigzag, DC difference).
 *   d_dc  device int32[nblk]      dc[0] absolute, dc[i>0] = difference to block i-1
 *   d_ac  device int32[nblk*63]   zigzag positions 1..63
 * with nblk = tic_num_blocks(height, width), blocks in raster order.
 */
int tic_encode_coeffs(tic_handle h, const void *d_pixels, int32_t height, int32_t width,
                      int32_t quality, int32_t *d_dc, int32_t *d_ac, void *stream);

/*
 * Host-buffer convenience around the two calls above (pinned staging, H2D, launch, D2H):
 * what a ctypes binding of compress() for one numpy image calls.
 *   out / out_capacity  host buffer; *out_size receives len(compress(image, quality)).
 * Synchronous.
 */
int tic_compress_host(tic_handle h, const uint8_t *pixels, int32_t height, int32_t width,
                      int32_t quality, uint32_t flags, uint8_t *out, int64_t out_capacity,
                      int64_t *out_size, int32_t *status);

/* Counters of the last tic_encode_batch on this handle (valid after tic_encode_finish):
 *   [0] kernels launched   [1] tiles   [2] coefficients sent to the exact FP64 path
 *   [3] coefficients the exact path changed   [4] blocks
 *   [5] device time of encode_tiles_kernel, ns   [6] device time of compact_kernel, ns, both summed
 *       over the [7] batches enqueued since the previous tic_encode_finish (CUDA events recorded
 *       on the batches' stream around those two launches; the newest 64 batches at most) */
int tic_last_stats(tic_handle h, int64_t stats[8]);

/* After a batch encoded with TIC_FLAG_DEBUG_ALL_EXACT (and tic_encode_finish): the number of coefficients whose
 * float64 exact value (the reference's, tinyimgcodec/utils.py:32-37,53) differs from the fast path's although the
 * tie guard had not flagged them.  0 proves the guard band on that input; -1 for a null handle. */
int64_t tic_last_guard_misses(tic_handle h);

/* ------------------------------------------------------------------------------------------------
 * Decode side: tinyimgcodec.codec.decompress(data) -> uint8 H x W
 * (tinyimgcodec/codec.py:167-189) and everything under it — parse_header (codec.py:117-130),
 * read_huffman_table (codec.py:87-99), decode_huffman / decode_run_length (tinyimgcodec/huffman.py:
 * 77-98, 36-38), decode (codec.py:46-70: DC cumsum, de-zigzag, dequantise, scipy idct, +128, clip,
 * crop, truncate to uint8).  Pixels are bit-identical to the reference decoder's for every stream the
 * reference decodes without an internal exception, in all three stream forms it understands: fixed
 * tables (flag 0), per-image tables (flag bit 31 as the reference READS it, i.e. the bytes 00 00 00 80),
 * and the embedded C encoder's scaled integer DCT (flag bit 30).
 * ---------------------------------------------------------------------------------------------- */

/* flags for tic_decode_batch */
#define TIC_DFLAG_ACCEPT_BE_FLAG 1u /* opt-in: also treat the header bytes 80 00 00 00 — what compress(...,
                                       auto_generate_huffman_table=True) writes (codec.py:111) and the
                                       reference's own parse_header misreads (codec.py:119,124) — as
                                       "per-image tables follow".  Off by default: the reference decodes such
                                       a stream with the fixed tables (garbage), and so does this library. */

#define TIC_DFLAG_EXACT_ONLY 2u     /* run the float64 IDCT on every block instead of the FP32 pass + float64 pass
                                       over the blocks whose pixels the FP32 pass cannot guarantee.  Same pixels
                                       either way (tests compare the two); for verification and profiling. */

#define TIC_DFLAG_FUSED 4u          /* opt-in: the coefficient pass transforms a block as soon as it is complete instead of
                                       writing its coefficients out for the inverse-transform kernels.  Same pixels either
                                       way (tests compare the two); removes the coefficient buffer's traffic but measures
                                       slower (csrc/tic_decode.cu, dec_write_kernel).  For verification and profiling. */

#define TIC_DFLAG_NO_EARLY_STOP 8u  /* synchronisation rounds: decode a subsequence to its end every time, instead of
                                       stopping where a repeat decode meets the previous one.  Same result (tests compare);
                                       for verification and profiling. */

#define TIC_DFLAG_SYNC_ROUNDS 16u    /* read the "anything repaired?" flag on the host after every synchronisation round
                                       (tic_decode_batch then blocks until the rounds have settled: the first version's
                                       behaviour).  By default two rounds (more once a handle has needed more) are
                                       enqueued blindly, the call returns at once, and tic_decode_finish repeats the
                                       batch with this flag in the rare case that the last of them still repaired. */

/* per-stream status bits of the decode side */
#define TIC_DSTATUS_HEADER 1     /* shorter than 16 bytes, or height / width differ from the caller's */
#define TIC_DSTATUS_CODE 2       /* no codeword matches (ValueError, huffman.py:72-73) or a run passes
                                    coefficient 63 */
#define TIC_DSTATUS_TRUNCATED 4  /* the stream ends before the last block */
#define TIC_DSTATUS_TABLE 8      /* per-image table too large for the device trie (1024 nodes) */
#define TIC_DSTATUS_QUALITY 16   /* quality field 0: ZeroDivisionError in the reference (utils.py:50) */
#define TIC_DSTATUS_RANGE 32     /* a DC value outside int16 (the reference keeps int32) */

/* parse_header's fixed part (codec.py:117-122): the four little-endian words of the 16-byte header.
 * Pure host arithmetic; TIC_E_INVALID when nbytes < 16 (struct.error in the reference). */
int tic_parse_header(const uint8_t *data, int64_t nbytes, int32_t *height, int32_t *width,
                     int32_t *quality, uint32_t *flag);

/*
 * decompress() for a batch of streams resident in device memory.
 *
 *   d_streams     HOST array of n device addresses, each 4-byte aligned; stream i is sizes[i] bytes and
 *                 the buffer must be readable up to the next multiple of 4.
 *   sizes         HOST array, bytes per stream (len(data)).
 *   heights/widths HOST arrays: the dimensions the caller sized d_pixels[i] for (from tic_parse_header);
 *                 a stream whose header disagrees gets TIC_DSTATUS_HEADER and is not written.
 *   d_pixels      HOST array of n device addresses; image i is written as heights[i] x widths[i] uint8,
 *                 row-major, contiguous.
 *   d_status      device int32[n] of TIC_DSTATUS_* bits, or NULL.
 *
 * Enqueued on `stream`; asynchronous (no host synchronisation) unless TIC_DFLAG_SYNC_ROUNDS is given: the
 * self-synchronising Huffman decode is launched a fixed number of times (two on a fresh handle) and whether the
 * last launch still changed an entry state is reported with the batch.  Streams, pixel buffers and d_status must
 * stay valid until tic_decode_finish(), which may decode the batch once more (then checking every round) before
 * it returns; pixels and status are valid after it.
 */
int tic_decode_batch(tic_handle h, const void *const *d_streams, const int64_t *sizes,
                     const int32_t *heights, const int32_t *widths, int32_t n_images, uint32_t flags,
                     void *const *d_pixels, int32_t *d_status, void *stream);

/* Synchronise `stream`, repeat the last tic_decode_batch if its blind synchronisation rounds had not settled, and
 * report it: TIC_OK or TIC_E_STREAM. */
int tic_decode_finish(tic_handle h, void *stream);

/* Host-buffer convenience: what a ctypes binding of decompress() for one `bytes` object calls.
 * out needs height*width bytes (tic_parse_header).  Synchronous. */
int tic_decompress_host(tic_handle h, const uint8_t *data, int64_t nbytes, uint32_t flags,
                        uint8_t *out, int64_t out_capacity, int32_t *status);

/*
 * decode() for one image whose coefficients are resident in device memory — replaces
 * tinyimgcodec/codec.py:46-70: the inverse of tic_encode_coeffs.
 *   d_dc   device int32[nblk]     dc[0] absolute, dc[i>0] differences (np.cumsum, codec.py:53)
 *   d_ac   device int32[nblk*63]  zigzag positions 1..63
 *   scaled_dct  the "scaled_dct" key (codec.py:50,58-62): coefficients of the embedded C encoder,
 *               quality = IMG_Q_BEST..LOW
 *   d_pixels    device uint8[height*width]
 * Errors the device sees (quality 0, a coefficient outside int16) are reported by tic_decode_finish().
 */
int tic_decode_coeffs(tic_handle h, const int32_t *d_dc, const int32_t *d_ac, int32_t height,
                      int32_t width, int32_t quality, int32_t scaled_dct, void *d_pixels, void *stream);

/* Counters of the last tic_decode_batch (valid after tic_decode_finish):
 *   [0] kernels launched  [1] subsequences (1024 stream bits each)  [2] synchronisation rounds
 *   [3] blocks  [4]-[7] device time in ns of: synchronisation rounds, scan, coefficient pass, IDCT (both passes)
 *   [8] blocks the FP32 IDCT pass handed to the exact float64 pass  [9]-[11] reserved */
int tic_decode_stats(tic_handle h, int64_t stats[12]);

#ifdef __cplusplus
}
#endif
#endif /* TINYIMGCODEC_CUDA_H */
