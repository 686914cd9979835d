"""Host-side mirror of the reference's codec entry points, on top of the C ABI.

    compress(image, quality=50, auto_generate_huffman_table=False) -> bytes
        mirrors tinyimgcodec/codec.py:133-164
    encode(image, quality=50) -> dict
        mirrors tinyimgcodec/codec.py:26-43
    decompress(data) -> uint8 H x W
        mirrors tinyimgcodec/codec.py:167-189
    decode(data: dict) -> uint8 H x W
        mirrors tinyimgcodec/codec.py:46-70

Same names, argument meaning and error behaviour as the reference; the work runs in the
sm_100a kernels behind libtinyimgcodec_cuda.so.  PyTorch is used only for device buffers
and streams in the batch API.  There is no CPU path: without the library and a B200 the
calls raise.
"""
import ctypes
import os
import struct
import threading

import numpy as np

from . import _lib

_encoders = {}
_encoders_lock = threading.Lock()


class TicError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libtinyimgcodec_cuda error {code}: {message}")
        self.code = code


class TicStreamError(ValueError):
    """A stream the decoder cannot parse cleanly.  The reference swallows the exception of every damaged
    block (try/except, codec.py:177-185) and returns an image with those blocks zeroed; the B200 path reports
    instead (pass strict=False to get whatever was decoded).  `.status`: TIC_DSTATUS_* bits per stream."""

    def __init__(self, message, status):
        super().__init__(message)
        self.status = status


def parse_header(data):
    """(height, width, quality, flag) — the fixed part of parse_header (codec.py:117-122); struct.error for
    fewer than 16 bytes, like struct.unpack in the reference."""
    if len(data) < 16:
        raise struct.error("unpack requires a buffer of 16 bytes")
    return struct.unpack("IIII", bytes(data[:16]))


def _default_device():
    return int(os.environ.get("TIC_DEVICE", "0"))


def _check_quality(quality):
    """The reference packs quality with struct 'I' (codec.py:103-108) and divides by it
    (utils.py:50): reproduce the exception types it raises."""
    struct.pack("I", quality)  # struct.error for float / negative, like make_header
    if quality == 0:
        raise ZeroDivisionError("division by zero")  # utils.py:50: 5000 / quality
    if quality == 100:
        # utils.py:53 divides by a zero table -> int32 min -> bits_required = 32 -> KeyError
        raise KeyError(32)
    if quality > 100:
        raise ValueError("quality above 100 is outside the codec's domain (negative quantisation table)")
    return int(quality)


C_QFACTORS = {"best": 0, "high": 1, "med": 2, "low": 3}   # c/img.h:22, c/encode.c:19-30


def _check_qfactor(qfactor):
    if isinstance(qfactor, str):
        if qfactor not in C_QFACTORS:
            raise ValueError("Invalid quality factor")   # c/encode.c:28
        return C_QFACTORS[qfactor]
    q = int(qfactor)
    if q not in (0, 1, 2, 3):
        raise ValueError("Invalid quality factor")
    return q


def _as_u8_image(image):
    """`height, width = image.shape` (codec.py:27) then astype(int32) (codec.py:29)."""
    image = np.asarray(image)
    height, width = image.shape  # ValueError for non-2-D input, like the reference
    if image.dtype != np.uint8:
        as_int = image.astype(np.int32)
        if as_int.size and (as_int.min() < 0 or as_int.max() > 255):
            raise ValueError("the B200 path encodes 8-bit grayscale: pixel values must be in 0..255")
        image = as_int.astype(np.uint8)
    return np.ascontiguousarray(image), int(height), int(width)


class Encoder:
    """One per GPU: owns the library handle (include/tinyimgcodec_cuda.h: tic_create)."""

    def __init__(self, device=None):
        self.lib = _lib.load()
        self.device = _default_device() if device is None else int(device)
        h = ctypes.c_void_p()
        rc = self.lib.tic_create(self.device, ctypes.byref(h))
        if rc != _lib.TIC_OK:
            raise TicError(rc, f"tic_create(device={self.device}) failed: no usable sm_100 GPU "
                               "(tinyimgcodec_b200 has no CPU fallback)")
        self.handle = h
        self._lock = threading.Lock()

    def close(self):
        if self.handle:
            self.lib.tic_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _raise(self, rc):
        msg = (self.lib.tic_last_error(self.handle) or b"").decode()
        if rc == _lib.TIC_E_CATEGORY:
            raise KeyError(msg)  # huffman.py:62: category not in the fixed table
        if rc == _lib.TIC_E_UNSUPPORTED:
            raise NotImplementedError(msg)
        if rc == _lib.TIC_E_TABLE:
            raise OverflowError(msg)  # int2ba in write_huffman_table, codec.py:76-77,81-83
        raise TicError(rc, msg)

    # -- single image, host buffers ---------------------------------------------------------
    def compress(self, image, quality=50, auto_generate_huffman_table=False, le_flag_word=False):
        """le_flag_word (auto-table streams only, off by default): write the header's flag word little-endian
        so that the reference's own decoder can open the stream (include/tinyimgcodec_cuda.h,
        TIC_FLAG_AUTO_LE_FLAG); the default is byte identity with the reference's compress()."""
        img, height, width = _as_u8_image(image)
        quality = _check_quality(quality)
        flags = _lib.TIC_FLAG_AUTO_HUFFMAN if auto_generate_huffman_table else 0
        if le_flag_word:
            if not auto_generate_huffman_table:
                raise ValueError("le_flag_word only applies to auto_generate_huffman_table=True")
            flags |= _lib.TIC_FLAG_AUTO_LE_FLAG
        if flags and img.size == 0:
            # calc_huffman_table indexes an empty symbol array (huffman.py:102-103)
            raise IndexError("index 1 is out of bounds for axis 1 with size 0")
        cap = int(self.lib.tic_max_out_bytes(height, width)) + (_lib.AUTO_HEADER_SLACK if flags else 0)
        out = np.empty(cap, dtype=np.uint8)
        size = ctypes.c_int64(0)
        status = ctypes.c_int32(0)
        with self._lock:
            rc = self.lib.tic_compress_host(self.handle, img.ctypes.data, height, width, quality, flags,
                                            out.ctypes.data, cap, ctypes.byref(size), ctypes.byref(status))
            if rc != _lib.TIC_OK:
                self._raise(rc)
        return out[: size.value].tobytes()

    def compress_c(self, image, qfactor="med"):
        """The stream of the reference's embedded C encoder (`c/encode <width> <height> [best|high|med|low]`
        fed the raw pixel rows, c/encode.c:13-66): header flag bit 30, integer FDCT, byte-identical to that
        binary's stdout for every block of the image.  (The binary then codes one more block row from a
        clobbered stack buffer, c/encode.c:47 — different on every run — which is not produced.)"""
        img, height, width = _as_u8_image(image)
        q = _check_qfactor(qfactor)
        if height % 8 or width % 8:
            raise ValueError("Width and height must be multiples of 8")   # c/encode.c:38-41
        cap = int(self.lib.tic_max_out_bytes(height, width))
        out = np.empty(cap, dtype=np.uint8)
        size = ctypes.c_int64(0)
        status = ctypes.c_int32(0)
        with self._lock:
            rc = self.lib.tic_compress_host(self.handle, img.ctypes.data, height, width, q, _lib.TIC_FLAG_C_VARIANT,
                                            out.ctypes.data, cap, ctypes.byref(size), ctypes.byref(status))
            if rc != _lib.TIC_OK:
                self._raise(rc)
        return out[: size.value].tobytes()

    def encode(self, image, quality=50):
        import torch
        img, height, width = _as_u8_image(image)
        q = _check_quality(quality)
        nblk = int(self.lib.tic_num_blocks(height, width))
        dev = torch.device("cuda", self.device)
        with self._lock, torch.cuda.device(dev):
            d_px = torch.from_numpy(img).to(dev) if img.size else torch.empty(0, dtype=torch.uint8, device=dev)
            d_dc = torch.empty(max(nblk, 1), dtype=torch.int32, device=dev)
            d_ac = torch.empty(max(nblk, 1) * 63, dtype=torch.int32, device=dev)
            stream = torch.cuda.current_stream(dev)
            rc = self.lib.tic_encode_coeffs(self.handle, d_px.data_ptr(), height, width, q, d_dc.data_ptr(),
                                            d_ac.data_ptr(), stream.cuda_stream)
            if rc != _lib.TIC_OK:
                self._raise(rc)
            stream.synchronize()
            dc = d_dc[:nblk].cpu().numpy()
            ac = d_ac[: nblk * 63].cpu().numpy().reshape(nblk, 63)
        return {"height": height, "width": width, "quality": quality, "dc": dc, "ac": ac}

    # -- batch, device buffers ----------------------------------------------------------------
    def encode_batch_device(self, d_images, quality=50, out=None, stream=None, auto_generate_huffman_table=False,
                            c_variant=False, debug_all_exact=False):
        """Encode a batch resident in HBM.  `d_images`: a CUDA uint8 tensor (N,H,W) or a list of
        2-D CUDA uint8 tensors.  Returns a DeviceBatchResult; nothing is copied to the host."""
        import torch
        q = _check_qfactor(quality) if c_variant else _check_quality(quality)
        dev = torch.device("cuda", self.device)
        if isinstance(d_images, torch.Tensor):
            # (N, H, W) tensor: the pointer / size arrays are built with numpy, no per-image Python work
            if d_images.dim() != 3:
                raise ValueError("expected an (N, H, W) uint8 tensor")
            if d_images.dtype != torch.uint8 or not d_images.is_contiguous() or d_images.device != dev:
                raise ValueError("images must be a contiguous uint8 tensor on the encoder's GPU")
            n, hgt, wid = (int(v) for v in d_images.shape)
            keep = d_images
            ptrs_np = np.uint64(d_images.data_ptr()) + np.arange(n, dtype=np.uint64) * np.uint64(hgt * wid)
            hs_np = np.full(max(n, 1), hgt, dtype=np.int32)
            ws_np = np.full(max(n, 1), wid, dtype=np.int32)
            worst = n * int(self.lib.tic_max_out_bytes(hgt, wid))
        else:
            tensors = list(d_images)
            n = len(tensors)
            for t in tensors:
                if t.dtype != torch.uint8 or t.dim() != 2 or not t.is_contiguous() or t.device != dev:
                    raise ValueError("images must be contiguous 2-D uint8 tensors on the encoder's GPU")
            keep = tensors
            ptrs_np = np.array([t.data_ptr() for t in tensors], dtype=np.uint64).reshape(-1)
            hs_np = np.array([t.shape[0] for t in tensors] or [0], dtype=np.int32)
            ws_np = np.array([t.shape[1] for t in tensors] or [0], dtype=np.int32)
            worst = sum(int(self.lib.tic_max_out_bytes(t.shape[0], t.shape[1])) for t in tensors)
        if ptrs_np.size == 0:
            ptrs_np = np.zeros(1, dtype=np.uint64)
        ptrs, hs, ws = ptrs_np.ctypes.data, hs_np.ctypes.data, ws_np.ctypes.data
        with torch.cuda.device(dev):
            if out is None:
                slack = _lib.AUTO_HEADER_SLACK if auto_generate_huffman_table else 0
                out = torch.empty(worst + n * slack + 16, dtype=torch.uint8, device=dev)
            meta = torch.empty(max(n, 1) * 3, dtype=torch.int64, device=dev)
            offs, sizes = meta[:n], meta[max(n, 1): max(n, 1) + n]
            stat = meta[2 * max(n, 1):].view(torch.int32)[:n]
            stream = stream or torch.cuda.current_stream(dev)
            with self._lock:
                flags = (_lib.TIC_FLAG_AUTO_HUFFMAN if auto_generate_huffman_table else 0) | \
                        (_lib.TIC_FLAG_C_VARIANT if c_variant else 0) | \
                        (_lib.TIC_FLAG_DEBUG_ALL_EXACT if debug_all_exact else 0)
                rc = self.lib.tic_encode_batch(self.handle, ptrs, hs, ws, n, q, flags, out.data_ptr(), out.numel(),
                                               offs.data_ptr(), sizes.data_ptr(), stat.data_ptr(),
                                               stream.cuda_stream)
                if rc != _lib.TIC_OK:
                    self._raise(rc)
        return DeviceBatchResult(self, out, offs, sizes, stat, stream, n, keep)

    # -- batch, host buffers, pipelined -----------------------------------------------------------
    def compress_batch_pinned(self, h_images, quality=50, chunk=64, out_bytes_per_pixel=0.75, nbuf=4, c_variant=False):
        """End-to-end encode of an (N,H,W) uint8 tensor in PINNED host memory — the batch form of the reference's
        caller (encode.py:10-19: read pixels, compress, write bytes).  Three streams: H2D of chunk c+1 and c+2,
        encode of chunk c and D2H of chunk c-1 run concurrently over `nbuf` device buffer pairs; the host never
        blocks the GPU (it waits for chunk c-1's sizes while chunk c encodes and the next copies are queued).
        Returns (host_buffer, [(offset, size)] per image) with offsets into host_buffer (16-byte aligned).
        out_bytes_per_pixel bounds the compressed size of a CHUNK (device buffer) and of the whole batch (host
        buffer); a batch that compresses worse raises TicError(TIC_E_CAPACITY) naming it.  c_variant: the stream of
        the reference's C encoder (compress_c), `quality` is then "best" | "high" | "med" | "low"."""
        import torch
        n, hgt, wid = (int(v) for v in h_images.shape)
        dev = torch.device("cuda", self.device)
        q = _check_qfactor(quality) if c_variant else _check_quality(quality)
        chunk = max(1, min(int(chunk), max(n, 1)))
        nbuf = max(3, int(nbuf))
        cap = (int(chunk * hgt * wid * out_bytes_per_pixel) + 4096 + 15) & ~15
        nchunks = (n + chunk - 1) // chunk
        with torch.cuda.device(dev):
            key = (chunk, hgt, wid, cap, nbuf)
            if getattr(self, "_pipe_key", None) != key:
                self._pipe = {
                    "d_in": [torch.empty((chunk, hgt, wid), dtype=torch.uint8, device=dev) for _ in range(nbuf)],
                    "d_out": [torch.empty(cap, dtype=torch.uint8, device=dev) for _ in range(nbuf)],
                    "s_in": torch.cuda.Stream(dev), "s_comp": torch.cuda.Stream(dev), "s_out": torch.cuda.Stream(dev),
                }
                self._pipe_key = key
            P = self._pipe
            need = int(n * hgt * wid * out_bytes_per_pixel) + 4096 * max(nchunks, 1)
            h_out = getattr(self, "_h_out", None)
            if h_out is None or h_out.numel() < need:
                h_out = self._h_out = torch.empty(need, dtype=torch.uint8).pin_memory()
            h_meta = getattr(self, "_h_meta", None)   # per chunk: offsets, sizes (int64) and status (int32 in an int64 slot)
            if h_meta is None or h_meta.shape[0] < nchunks or h_meta.shape[2] < chunk:
                h_meta = self._h_meta = torch.empty((max(nchunks, 1), 3, chunk), dtype=torch.int64).pin_memory()
            ev_in = [torch.cuda.Event() for _ in range(nchunks)]
            ev_comp = [torch.cuda.Event() for _ in range(nchunks)]
            ev_out = [torch.cuda.Event() for _ in range(nchunks)]
            keep, index, h_pos = {}, [], 0

            def span(c):
                return c * chunk, min(n, (c + 1) * chunk)

            def issue_h2d(c):
                lo, hi = span(c)
                with torch.cuda.stream(P["s_in"]):
                    if c >= nbuf:
                        P["s_in"].wait_event(ev_comp[c - nbuf])   # input buffer free again
                    P["d_in"][c % nbuf][: hi - lo].copy_(h_images[lo:hi], non_blocking=True)
                    ev_in[c].record(P["s_in"])

            def issue_encode(c):
                lo, hi = span(c)
                P["s_comp"].wait_event(ev_in[c])
                if c >= nbuf:
                    P["s_comp"].wait_event(ev_out[c - nbuf])      # output buffer drained
                res = self.encode_batch_device(P["d_in"][c % nbuf][: hi - lo], q, out=P["d_out"][c % nbuf], stream=P["s_comp"],
                                               c_variant=c_variant)
                with torch.cuda.stream(P["s_comp"]):
                    h_meta[c, 0, : hi - lo].copy_(res.offsets, non_blocking=True)
                    h_meta[c, 1, : hi - lo].copy_(res.sizes, non_blocking=True)
                    h_meta[c, 2, : hi - lo].view(torch.int32)[: hi - lo].copy_(res.status, non_blocking=True)
                    ev_comp[c].record(P["s_comp"])
                keep[c] = res

            def drain(c):
                nonlocal h_pos
                lo, hi = span(c)
                ev_comp[c].synchronize()       # chunk c only: chunk c+1 is encoding, the next H2D copies are queued
                offs, sizes = h_meta[c, 0, : hi - lo].numpy(), h_meta[c, 1, : hi - lo].numpy()
                status = h_meta[c, 2, : hi - lo].view(torch.int32)[: hi - lo].numpy()
                total = int((offs + sizes).max()) if hi > lo else 0
                if total > cap or h_pos + total > h_out.numel():
                    raise TicError(_lib.TIC_E_CAPACITY, f"output buffer too small: images {lo}..{hi - 1} need {total} bytes; "
                                                        f"raise out_bytes_per_pixel (now {out_bytes_per_pixel})")
                if hi > lo and (int(np.bitwise_or.reduce(status)) & _lib.TIC_STATUS_CATEGORY):
                    raise KeyError("coefficient category outside the fixed Huffman tables (reference: KeyError)")
                with torch.cuda.stream(P["s_out"]):
                    P["s_out"].wait_event(ev_comp[c])
                    h_out[h_pos: h_pos + total].copy_(P["d_out"][c % nbuf][:total], non_blocking=True)
                    ev_out[c].record(P["s_out"])
                index.extend((int(h_pos + o), int(s)) for o, s in zip(offs, sizes))
                h_pos += (total + 15) & ~15
                del keep[c]

            try:
                for c in range(min(2, nchunks)):
                    issue_h2d(c)
                for c in range(nchunks):
                    issue_encode(c)
                    if c + 2 < nchunks:
                        issue_h2d(c + 2)
                    if c >= 1:
                        drain(c - 1)
                if nchunks:
                    drain(nchunks - 1)
                P["s_out"].synchronize()
            finally:
                # one finish for the whole batch: collects whatever the device flagged (sticky) and leaves the handle clean
                total = ctypes.c_int64(0)
                with self._lock:
                    rc = self.lib.tic_encode_finish(self.handle, P["s_comp"].cuda_stream, ctypes.byref(total))
                keep.clear()
            if rc not in (_lib.TIC_OK,):
                self._raise(rc)
        return h_out, index

    # -- decode side ----------------------------------------------------------------------------------
    def _raise_decode(self, rc, status):
        msg = (self.lib.tic_last_error(self.handle) or b"").decode()
        if rc == _lib.TIC_E_STREAM:
            bits = int(np.bitwise_or.reduce(np.asarray(status, dtype=np.int64).reshape(-1))) if np.size(status) else 0
            if bits & _lib.TIC_DSTATUS_QUALITY:
                raise ZeroDivisionError("division by zero")   # utils.py:50: 5000 / quality
            raise TicStreamError(msg, status)
        raise TicError(rc, msg)

    def decompress(self, data, strict=True, accept_be_flag=False, exact_only=False, fused=False):
        """tinyimgcodec.codec.decompress (codec.py:167-189) for one stream in host memory."""
        height, width, _, _ = parse_header(data)
        buf = np.frombuffer(bytes(data), dtype=np.uint8)
        out = np.zeros((height, width), dtype=np.uint8)   # a stream the device does not decode at all (strict=False) reads as 0
        status = ctypes.c_int32(0)
        flags = (_lib.TIC_DFLAG_ACCEPT_BE_FLAG if accept_be_flag else 0) | (_lib.TIC_DFLAG_EXACT_ONLY if exact_only else 0) | \
            (_lib.TIC_DFLAG_FUSED if fused else 0)
        with self._lock:
            rc = self.lib.tic_decompress_host(self.handle, buf.ctypes.data, buf.size, flags, out.ctypes.data,
                                              out.size, ctypes.byref(status))
            if rc != _lib.TIC_OK and (strict or rc != _lib.TIC_E_STREAM):
                self._raise_decode(rc, [status.value])
        return out

    def decode_batch_device(self, d_streams, sizes, heights, widths, pixels=None, stream=None,
                            accept_be_flag=False, strict=True, exact_only=False, fused=False, early_stop=True, sync_rounds=False):
        """Decode streams resident in HBM.  `d_streams`: list of CUDA uint8 tensors (each 4-byte aligned), or
        one CUDA uint8 tensor plus `sizes` and byte offsets given as d_streams=(tensor, offsets).  Returns the
        CUDA uint8 (H, W) images (a list of views of one pixel buffer, or one (N, H, W) view when all shapes are
        equal) and the status array (numpy int32)."""
        import torch
        dev = torch.device("cuda", self.device)
        if isinstance(d_streams, tuple):
            base, offsets = d_streams
            ptrs_np = np.uint64(base.data_ptr()) + np.asarray(offsets, dtype=np.uint64)
            keep = base
        else:
            keep = list(d_streams)
            ptrs_np = np.array([t.data_ptr() for t in keep], dtype=np.uint64).reshape(-1)
        n = int(ptrs_np.size)
        sizes_np = np.ascontiguousarray(sizes, dtype=np.int64).reshape(-1)
        hs_np = np.ascontiguousarray(heights, dtype=np.int32).reshape(-1)
        ws_np = np.ascontiguousarray(widths, dtype=np.int32).reshape(-1)
        if not (sizes_np.size == hs_np.size == ws_np.size == n):
            raise ValueError("d_streams, sizes, heights and widths must have one entry per stream")
        npx = hs_np.astype(np.int64) * ws_np.astype(np.int64)
        px_off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum((npx + 15) & ~np.int64(15), out=px_off[1:])
        with torch.cuda.device(dev):
            if pixels is None:
                pixels = torch.empty(int(px_off[-1]) + 16, dtype=torch.uint8, device=dev)
            elif pixels.numel() < int(px_off[-1]):
                raise ValueError("pixel buffer too small")
            out_ptrs = np.uint64(pixels.data_ptr()) + px_off[:-1].astype(np.uint64)
            d_status = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
            stream = stream or torch.cuda.current_stream(dev)
            flags = (_lib.TIC_DFLAG_ACCEPT_BE_FLAG if accept_be_flag else 0) | \
                    (_lib.TIC_DFLAG_EXACT_ONLY if exact_only else 0) | (_lib.TIC_DFLAG_FUSED if fused else 0) | \
                    (0 if early_stop else _lib.TIC_DFLAG_NO_EARLY_STOP) | (_lib.TIC_DFLAG_SYNC_ROUNDS if sync_rounds else 0)
            if n == 0:
                return [], np.zeros(0, dtype=np.int32)
            with self._lock:
                rc = self.lib.tic_decode_batch(self.handle, ptrs_np.ctypes.data, sizes_np.ctypes.data,
                                               hs_np.ctypes.data, ws_np.ctypes.data, n, flags,
                                               out_ptrs.ctypes.data, d_status.data_ptr(), stream.cuda_stream)
                if rc == _lib.TIC_OK:
                    rc = self.lib.tic_decode_finish(self.handle, stream.cuda_stream)
                status = d_status[:n].cpu().numpy() if rc in (_lib.TIC_OK, _lib.TIC_E_STREAM) else None
                if rc != _lib.TIC_OK and (strict or rc != _lib.TIC_E_STREAM):
                    self._raise_decode(rc, status)
                if rc == _lib.TIC_E_STREAM:   # strict=False: streams the device did not decode at all (header mismatch,
                    # quality 0) must not come back as whatever the pixel buffer held before
                    for i in np.nonzero(status & (_lib.TIC_DSTATUS_HEADER | _lib.TIC_DSTATUS_QUALITY))[0]:
                        pixels[int(px_off[i]): int(px_off[i]) + int(npx[i])].zero_()
        del keep
        if n and (hs_np == hs_np[0]).all() and (ws_np == ws_np[0]).all() and int(npx[0]) % 16 == 0:
            # equal shapes, densely packed: one (N, H, W) view instead of N slices (indexable like the list)
            images = pixels[: n * int(npx[0])].view(n, int(hs_np[0]), int(ws_np[0]))
        else:
            images = [pixels[int(px_off[i]): int(px_off[i]) + int(npx[i])].view(int(hs_np[i]), int(ws_np[i]))
                      for i in range(n)]
        return images, status

    def decompress_batch(self, streams, strict=True, accept_be_flag=False, exact_only=False, fused=False, early_stop=True,
                         sync_rounds=False):
        """decompress() for a list of `bytes`: one pinned H2D copy of all streams, one decode, one D2H copy of all
        pixels.  Returns a list of uint8 (H, W) arrays."""
        import torch
        dev = torch.device("cuda", self.device)
        streams = [bytes(s) for s in streams]
        hdrs = [parse_header(s) for s in streams]
        sizes = np.array([len(s) for s in streams], dtype=np.int64)
        offs = np.zeros(len(streams) + 1, dtype=np.int64)
        np.cumsum((sizes + 15) & ~np.int64(15), out=offs[1:])
        with torch.cuda.device(dev):
            h_buf = torch.zeros(int(offs[-1]) + 16, dtype=torch.uint8).pin_memory()
            h_np = h_buf.numpy()
            for s, o in zip(streams, offs[:-1]):
                h_np[o: o + len(s)] = np.frombuffer(s, dtype=np.uint8)
            d_buf = h_buf.to(dev, non_blocking=True)
            imgs, _ = self.decode_batch_device((d_buf, offs[:-1]), sizes, [h[0] for h in hdrs], [h[1] for h in hdrs],
                                               strict=strict, accept_be_flag=accept_be_flag, exact_only=exact_only, fused=fused,
                                               early_stop=early_stop, sync_rounds=sync_rounds)
            return [im.cpu().numpy() for im in imgs]

    def decompress_batch_pinned(self, h_streams, index, heights, widths, chunk=64, nbuf=3, strict=True):
        """The mirror of compress_batch_pinned for the decode side (decompress, codec.py:167-189, for a batch): `h_streams`
        is a PINNED uint8 tensor holding the streams, `index` the [(offset, size)] list compress_batch_pinned returns
        (offsets 16-byte aligned, ascending).  Chunks of `chunk` streams go H2D on a copy stream while the previous chunk
        decodes and the one before returns its pixels.  Returns (pinned pixel buffer, [(offset, height, width)]); for
        equal shapes the buffer is an (N, H, W) tensor."""
        import torch
        n = len(index)
        dev = torch.device("cuda", self.device)
        hs = np.ascontiguousarray(heights, dtype=np.int64).reshape(-1)
        ws = np.ascontiguousarray(widths, dtype=np.int64).reshape(-1)
        if hs.size != n or ws.size != n:
            raise ValueError("index, heights and widths must have one entry per stream")
        offs = np.array([o for o, _ in index], dtype=np.int64)
        sizes = np.array([s for _, s in index], dtype=np.int64)
        if n and (np.any(offs % 4) or np.any(np.diff(offs) < 0)):
            raise ValueError("stream offsets must be 4-byte aligned and ascending")
        npx = hs * ws
        px_off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum((npx + 15) & ~np.int64(15), out=px_off[1:])
        chunk = max(1, min(int(chunk), max(n, 1)))
        nbuf = max(2, int(nbuf))
        nchunks = (n + chunk - 1) // chunk
        spans = [(c * chunk, min(n, (c + 1) * chunk)) for c in range(nchunks)]
        brange = [(int(offs[lo]), int((offs[hi - 1] + sizes[hi - 1] + 15) & ~15)) for lo, hi in spans]
        in_cap = max([b1 - b0 for b0, b1 in brange] or [16]) + 16
        px_cap = max([int(px_off[hi] - px_off[lo]) for lo, hi in spans] or [16]) + 16
        with torch.cuda.device(dev):
            key = ("dec", in_cap, px_cap, nbuf)
            if getattr(self, "_dpipe_key", None) != key:
                self._dpipe = {
                    "d_in": [torch.zeros(in_cap, dtype=torch.uint8, device=dev) for _ in range(nbuf)],
                    "d_px": [torch.empty(px_cap, dtype=torch.uint8, device=dev) for _ in range(nbuf)],
                    "s_in": torch.cuda.Stream(dev), "s_comp": torch.cuda.Stream(dev), "s_out": torch.cuda.Stream(dev),
                }
                self._dpipe_key = key
            P = self._dpipe
            h_px = getattr(self, "_h_px", None)
            if h_px is None or h_px.numel() < int(px_off[-1]) + 16:
                h_px = self._h_px = torch.empty(int(px_off[-1]) + 16, dtype=torch.uint8).pin_memory()
            ev_in = [torch.cuda.Event() for _ in range(nchunks)]
            ev_comp = [torch.cuda.Event() for _ in range(nchunks)]
            ev_out = [torch.cuda.Event() for _ in range(nchunks)]

            def issue_h2d(c):
                b0, b1 = brange[c]
                b1 = min(b1, h_streams.numel())
                with torch.cuda.stream(P["s_in"]):
                    if c >= nbuf:
                        P["s_in"].wait_event(ev_comp[c - nbuf])
                    P["d_in"][c % nbuf][: b1 - b0].copy_(h_streams[b0:b1], non_blocking=True)
                    ev_in[c].record(P["s_in"])

            if nchunks:
                issue_h2d(0)
            for c, (lo, hi) in enumerate(spans):
                if c + 1 < nchunks:
                    issue_h2d(c + 1)
                P["s_comp"].wait_event(ev_in[c])
                if c >= nbuf:
                    P["s_comp"].wait_event(ev_out[c - nbuf])
                # decode_batch_device ends with tic_decode_finish (a stream synchronisation): the call returns when chunk c is
                # decoded; the copies of its neighbours run meanwhile on their own streams
                self.decode_batch_device((P["d_in"][c % nbuf], offs[lo:hi] - brange[c][0]), sizes[lo:hi], hs[lo:hi], ws[lo:hi],
                                         pixels=P["d_px"][c % nbuf], stream=P["s_comp"], strict=strict)
                ev_comp[c].record(P["s_comp"])
                with torch.cuda.stream(P["s_out"]):
                    P["s_out"].wait_event(ev_comp[c])
                    nb = int(px_off[hi] - px_off[lo])
                    h_px[int(px_off[lo]): int(px_off[lo]) + nb].copy_(P["d_px"][c % nbuf][:nb], non_blocking=True)
                    ev_out[c].record(P["s_out"])
            P["s_out"].synchronize()
        if n and (hs == hs[0]).all() and (ws == ws[0]).all() and int(npx[0]) % 16 == 0:
            return h_px[: n * int(npx[0])].view(n, int(hs[0]), int(ws[0])), [(int(px_off[i]), int(hs[i]), int(ws[i])) for i in range(n)]
        return h_px, [(int(px_off[i]), int(hs[i]), int(ws[i])) for i in range(n)]

    def decode(self, data):
        """tinyimgcodec.codec.decode (codec.py:46-70): the dict encode() returns (plus the optional
        "scaled_dct" key parse_header sets, codec.py:122) -> uint8 H x W."""
        import torch
        height, width, quality = int(data["height"]), int(data["width"]), data["quality"]
        scaled = bool(data.get("scaled_dct", False))
        struct.pack("I", quality)
        if quality == 0 and not scaled:
            raise ZeroDivisionError("division by zero")   # utils.py:50
        nblk = int(self.lib.tic_num_blocks(height, width))
        dc = np.ascontiguousarray(data["dc"], dtype=np.int32).reshape(-1)
        ac = np.ascontiguousarray(data["ac"], dtype=np.int32).reshape(-1)
        if dc.size != nblk or ac.size != nblk * 63:
            raise ValueError(f"expected {nblk} dc and {nblk}x63 ac coefficients")
        dev = torch.device("cuda", self.device)
        with self._lock, torch.cuda.device(dev):
            d_dc = torch.from_numpy(dc).to(dev) if nblk else torch.empty(1, dtype=torch.int32, device=dev)
            d_ac = torch.from_numpy(ac).to(dev) if nblk else torch.empty(1, dtype=torch.int32, device=dev)
            d_px = torch.empty(max(height * width, 1), dtype=torch.uint8, device=dev)
            stream = torch.cuda.current_stream(dev)
            rc = self.lib.tic_decode_coeffs(self.handle, d_dc.data_ptr(), d_ac.data_ptr(), height, width,
                                            int(quality), int(scaled), d_px.data_ptr(), stream.cuda_stream)
            if rc == _lib.TIC_OK:
                rc = self.lib.tic_decode_finish(self.handle, stream.cuda_stream)
            if rc != _lib.TIC_OK:
                self._raise_decode(rc, [])
            return d_px[: height * width].cpu().numpy().reshape(height, width)

    def decode_stats(self):
        arr = (ctypes.c_int64 * 12)()
        self.lib.tic_decode_stats(self.handle, arr)
        return {"launches": arr[0], "subsequences": arr[1], "sync_rounds": arr[2], "blocks": arr[3],
                "sync_ms": arr[4] * 1e-6, "scan_ms": arr[5] * 1e-6, "scatter_ms": arr[6] * 1e-6,
                "idct_ms": arr[7] * 1e-6, "exact_blocks": arr[8]}

    def guard_misses(self):
        """After a batch encoded with debug_all_exact=True and finish(): coefficients whose float64 exact value
        differs from the fast path's although the tie guard had not flagged them (must be 0)."""
        return int(self.lib.tic_last_guard_misses(self.handle))

    def stats(self):
        arr = (ctypes.c_int64 * 8)()
        self.lib.tic_last_stats(self.handle, arr)
        return {"launches": arr[0], "tiles": arr[1], "exact_items": arr[2], "exact_changed": arr[3],
                "blocks": arr[4], "encode_kernel_ms_sum": arr[5] * 1e-6, "compact_kernel_ms_sum": arr[6] * 1e-6,
                "timed_batches": arr[7]}


class DeviceBatchResult:
    """Streams of one encode_batch_device call, still in HBM."""

    def __init__(self, enc, out, offsets, sizes, status, stream, n, inputs=None):
        self.enc, self.out, self.offsets, self.sizes, self.status = enc, out, offsets, sizes, status
        self.stream, self.n = stream, n
        self.inputs = inputs   # keeps the pixel tensors alive until the result is dropped
        self.total_bytes = None

    def finish(self):
        """Synchronise and raise what the reference would have raised."""
        total = ctypes.c_int64(0)
        with self.enc._lock:
            rc = self.enc.lib.tic_encode_finish(self.enc.handle, self.stream.cuda_stream, ctypes.byref(total))
            self.total_bytes = total.value
            if rc != _lib.TIC_OK:
                self.enc._raise(rc)
        return self

    def to_bytes(self):
        """One D2H copy of the dense stream buffer, split into per-image bytes objects."""
        if self.total_bytes is None:
            self.finish()
        host = self.out[: self.total_bytes].cpu().numpy()
        offs = self.offsets.cpu().numpy()
        sizes = self.sizes.cpu().numpy()
        return [host[o: o + s].tobytes() for o, s in zip(offs, sizes)]


def get_encoder(device=None):
    device = _default_device() if device is None else int(device)
    with _encoders_lock:
        enc = _encoders.get(device)
        if enc is None:
            enc = _encoders[device] = Encoder(device)
        return enc


def compress(image, quality=50, auto_generate_huffman_table=False, device=None, le_flag_word=False):
    """Drop-in for tinyimgcodec.codec.compress (codec.py:133-164)."""
    return get_encoder(device).compress(image, quality, auto_generate_huffman_table, le_flag_word)


def encode(image, quality=50, device=None):
    """Drop-in for tinyimgcodec.codec.encode (codec.py:26-43)."""
    return get_encoder(device).encode(image, quality)


def decompress(data, device=None, strict=True, accept_be_flag=False, exact_only=False, fused=False):
    """Drop-in for tinyimgcodec.codec.decompress (codec.py:167-189)."""
    return get_encoder(device).decompress(data, strict=strict, accept_be_flag=accept_be_flag, exact_only=exact_only,
                                          fused=fused)


def decode(data, device=None):
    """Drop-in for tinyimgcodec.codec.decode (codec.py:46-70)."""
    return get_encoder(device).decode(data)


def decompress_batch(streams, device=None, strict=True, accept_be_flag=False, exact_only=False, fused=False,
                     early_stop=True, sync_rounds=False):
    """decompress() for a list of streams in one launch sequence."""
    return get_encoder(device).decompress_batch(streams, strict=strict, accept_be_flag=accept_be_flag,
                                                exact_only=exact_only, fused=fused, early_stop=early_stop,
                                                sync_rounds=sync_rounds)


def compress_c(image, qfactor="med", device=None):
    """The reference's embedded C encoder (c/encode.c) for one image: see Encoder.compress_c."""
    return get_encoder(device).compress_c(image, qfactor)


def compress_batch(images, quality=50, device=None, auto_generate_huffman_table=False):
    """compress() for a list of 2-D uint8 arrays (or an (N,H,W) array) in one launch sequence:
    pinned H2D, one encode, one D2H.  Returns a list of bytes objects."""
    import torch
    enc = get_encoder(device)
    dev = torch.device("cuda", enc.device)
    imgs = [_as_u8_image(im)[0] for im in images]
    with torch.cuda.device(dev):
        d_imgs = [torch.from_numpy(im).pin_memory().to(dev, non_blocking=True) if im.size
                  else torch.empty(im.shape, dtype=torch.uint8, device=dev) for im in imgs]
        if auto_generate_huffman_table and any(im.size == 0 for im in imgs):
            raise IndexError("index 1 is out of bounds for axis 1 with size 0")   # huffman.py:102-103
        return enc.encode_batch_device(d_imgs, quality,
                                       auto_generate_huffman_table=auto_generate_huffman_table).to_bytes()
