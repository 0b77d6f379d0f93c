"""
Device-resident batch decoder / encoder: the throughput API over libvc2b200.so.

PyTorch is used only for device memory, streams and (in bench.py) torch.distributed.
Every compute step is a hand-written CUDA kernel behind the C ABI (include/vc2_b200.h).
"""

import ctypes as C

import numpy as np

from . import capi
from .params import Geometry

PAD = 64  # readable slack after the stream bytes (the kernels fetch aligned 16-byte pieces)

# Default for BatchCodec(coeff16=None): whether decoders stage the high-pass coefficients as int16 between the
# unpack and the IDWT where the library supports it (vc2b_coeff16_supported).  Results are identical: a picture
# whose coefficients do not fit is decoded again through the int32 plane.  The compat layer (pseudocode.py)
# hands coefficients to the reference as int32 and always uses the int32 plane.
DEFAULT_COEFF16 = False


class StreamError(Exception):
    """A per-picture decode status other than OK; mirrors the reference's exceptions
    (decoder/exceptions.py: UnexpectedEndOfStream, InvalidSliceYLength)."""

    def __init__(self, picture, status, bad_slice):
        self.picture = picture
        self.status = status
        self.bad_slice = bad_slice
        names = {
            capi.ST_UNEXPECTED_END_OF_STREAM: "UnexpectedEndOfStream",
            capi.ST_INVALID_SLICE_Y_LENGTH: "InvalidSliceYLength",
            capi.ST_INT32_ENVELOPE: "value outside the int32 envelope",
            capi.ST_COEFF16_RANGE: "coefficient outside the int16 plane (decode again with coeff16=False)",
        }
        Exception.__init__(
            self, "picture %d, slice %d: %s" % (picture, bad_slice, names.get(status, status))
        )


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("vc2_conformance_b200 needs a CUDA device (there is no CPU fallback)")
    return torch


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


_geometry_cache = {}


def geometry_for(params):
    """Geometry of a parameter set, cached on the parameter bytes (about a hundred C-ABI calls otherwise)."""
    key = bytes(bytearray(params))
    g = _geometry_cache.get(key)
    if g is None:
        if len(_geometry_cache) > 64:
            _geometry_cache.clear()
        g = _geometry_cache[key] = Geometry(params)
    return g


class BatchCodec(object):
    """Buffers and launch sequence for batches of up to ``max_pictures`` pictures that share one
    parameter set.  All device work is enqueued on ``stream`` (default: torch's current)."""

    def __init__(self, params, max_pictures, max_stream_bytes=None, device="cuda:0", group=128, coeff16=None):
        torch = _torch()
        self.torch = torch
        self.lib = capi.lib()
        self.p = params
        self.g = geometry_for(params)
        self.device = torch.device(device)
        self.max_pictures = int(max_pictures)
        g = self.g
        dev = self.device
        if coeff16 is None:
            coeff16 = DEFAULT_COEFF16
        self.coeff16 = bool(coeff16) and bool(self.lib.vc2b_coeff16_supported(C.byref(params)))
        self._plane32 = set()  # pictures of the last decode whose coefficients are all in the int32 plane
        with torch.cuda.device(dev):
            self.coeffs = torch.empty(self.max_pictures * g.picture_coeffs, dtype=torch.int32, device=dev)
            self.coeffs16 = (torch.empty(self.max_pictures * g.picture_coeffs, dtype=torch.int16, device=dev)
                             if self.coeff16 else None)
            self.pictures = torch.empty(self.max_pictures * g.picture_samples, dtype=torch.uint16, device=dev)
            self.group = max(1, min(int(group), self.max_pictures))
            self.scratch = torch.empty(max(g.scratch_bytes * self.group, 16), dtype=torch.uint8, device=dev)
            self.status = torch.zeros(self.max_pictures * 4, dtype=torch.int32, device=dev)
            self.qindex = torch.zeros(self.max_pictures * g.num_slices, dtype=torch.int32, device=dev)
            self.slice_bits = torch.zeros(self.max_pictures * g.num_slices * 3, dtype=torch.int32, device=dev)
            self.slice_offset = torch.zeros(self.max_pictures * g.num_slices, dtype=torch.int32, device=dev)
            self.pic_offset = torch.zeros(self.max_pictures + 1, dtype=torch.int64, device=dev)
            self.data = None
            self._reserve_data(max_stream_bytes or 1 << 20)

    # -- buffers ---------------------------------------------------------------------
    def _reserve_data(self, nbytes):
        torch = self.torch
        need = int(nbytes) + PAD
        if self.data is None or self.data.numel() < need:
            if self.data is not None:
                # kernels enqueued on a side stream may still read the old buffer: the caching allocator
                # only knows about the stream it was allocated on
                torch.cuda.synchronize(self.device)
            with torch.cuda.device(self.device):
                self.data = torch.zeros((need + 4095) // 4096 * 4096, dtype=torch.uint8, device=self.device)

    @property
    def launches(self):
        """Kernels launched by libvc2b200 in this process so far (vc2b_launch_count)."""
        return int(self.lib.vc2b_launch_count())

    def _stream_ptr(self, stream):
        torch = self.torch
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        return C.c_void_p(s.cuda_stream)

    def _on(self, stream):
        """Context: this codec's device current, and ``stream`` (if given) torch's current stream, so that
        allocations, copies and the library's launches are ordered on one stream of the right device."""
        return _DeviceStream(self.torch, self.device, stream)

    # -- decode --------------------------------------------------------------------------
    def pack_host_streams(self, streams):
        """Concatenate the pictures' transform-data bytes back to back in one pinned host
        buffer.  Returns (pinned uint8 tensor incl. PAD slack, pinned int64 offsets [n+1])."""
        torch = self.torch
        n = len(streams)
        assert n <= self.max_pictures
        offs = np.zeros(n + 1, np.int64)
        for i, s in enumerate(streams):
            offs[i + 1] = offs[i] + len(s)
        host = torch.zeros(int(offs[n]) + PAD, dtype=torch.uint8).pin_memory()
        hv = host.numpy()
        for i, s in enumerate(streams):
            hv[offs[i]:offs[i + 1]] = np.frombuffer(s, np.uint8)
        return host, torch.from_numpy(offs).pin_memory()

    def upload_streams(self, host, offs, n, stream=None):
        """H2D copy of packed streams (pinned uint8 tensor incl. PAD slack, pinned int64 offsets [n+1])."""
        self._reserve_data(host.numel())
        with self._on(stream):
            self.data[: host.numel()].copy_(host, non_blocking=True)
            self.pic_offset[: n + 1].copy_(offs, non_blocking=True)

    def decode_host(self, host, offs, n, stream=None):
        """H2D copy of packed streams (pinned) and decode; everything enqueued on ``stream``."""
        self.upload_streams(host, offs, n, stream)
        self.decode_device(self.data, self.pic_offset, n, stream)

    def decode_streams(self, streams, stream=None, check=True):
        """End-to-end decode of host streams: H2D, slice index, unpack+dequant, IDWT+clip+offset.
        Returns the device picture tensor view [n, picture_samples] (uint16).  With the int16 coefficient
        plane, pictures that do not fit it are decoded again through the int32 plane before returning."""
        n = len(streams)
        host, offs = self.pack_host_streams(streams)
        self.decode_host(host, offs, n, stream)
        if self.coeff16:
            self.redo_coeff16(self.status_host(n, stream), stream=stream)
        if check:
            self.check_status(n, stream)
        return self.pictures[: n * self.g.picture_samples].view(n, self.g.picture_samples)

    def unpack_streams(self, streams, stream=None):
        """transform_data only (decoder/transform_data_syntax.py:116): H2D, slice index, unpack +
        inverse quantisation (+ DC prediction for LD) into ``self.coeffs``; no IDWT."""
        n = len(streams)
        host, offs = self.pack_host_streams(streams)
        self.upload_streams(host, offs, n, stream)
        self.unpack_device(self.data, self.pic_offset, n, stream)

    def unpack_device(self, d_data, d_pic_offset, n, stream=None, first=0, coeff16=None):
        """Slice index + unpack of ``n`` pictures whose transform data is resident in HBM (``d_pic_offset``:
        the offsets of those pictures; results go to picture slots ``first``..)."""
        L, p, g = self.lib, self.p, self.g
        c16 = self.coeff16 if coeff16 is None else (coeff16 and self.coeff16)
        for i in range(first, first + n):  # which plane holds picture i's high-pass coefficients
            self._plane32.discard(i) if c16 else self._plane32.add(i)
        coeffs = self.coeffs[first * g.picture_coeffs:]
        status = self.status[first * 4:]
        qindex = self.qindex[first * g.num_slices:]
        so = self.slice_offset[first * g.num_slices:]
        with self._on(stream):
            sp = self._stream_ptr(stream)
            if not p.is_ld:
                capi.check("vc2b_hq_slice_offsets",
                           L.vc2b_hq_slice_offsets(C.byref(p), _ptr(d_data), _ptr(d_pic_offset), n, _ptr(so),
                                                   _ptr(status), sp))
            if c16:
                capi.check("vc2b_unpack16",
                           L.vc2b_unpack16(C.byref(p), _ptr(d_data), _ptr(d_pic_offset), n, _ptr(so), _ptr(coeffs),
                                           _ptr(self.coeffs16[first * g.picture_coeffs:]), _ptr(qindex),
                                           _ptr(status), sp))
            else:
                capi.check("vc2b_unpack",
                           L.vc2b_unpack(C.byref(p), _ptr(d_data), _ptr(d_pic_offset), n, _ptr(so), _ptr(coeffs),
                                         _ptr(qindex), _ptr(status), sp))

    def decode_device(self, d_data, d_pic_offset, n, stream=None, first=0, coeff16=None):
        """Decode ``n`` pictures whose transform data is already resident in HBM.  With the int16 plane a
        picture may come back with status ST_COEFF16_RANGE: ``redo_coeff16`` decodes those again."""
        self.unpack_device(d_data, d_pic_offset, n, stream, first, coeff16)
        self.idwt_device(n, stream, first, coeff16)

    def idwt_device(self, n, stream=None, first=0, coeff16=None):
        L, p, g = self.lib, self.p, self.g
        c16 = self.coeff16 if coeff16 is None else (coeff16 and self.coeff16)
        coeffs = self.coeffs[first * g.picture_coeffs:]
        pics = self.pictures[first * g.picture_samples:]
        with self._on(stream):
            if c16:
                capi.check("vc2b_idwt16",
                           L.vc2b_idwt16(C.byref(p), _ptr(coeffs), _ptr(self.coeffs16[first * g.picture_coeffs:]), n,
                                         _ptr(pics), _ptr(self.scratch), self.scratch.numel(),
                                         _ptr(self.status[first * 4:]), self._stream_ptr(stream)))
            else:
                capi.check("vc2b_idwt",
                           L.vc2b_idwt(C.byref(p), _ptr(coeffs), n, _ptr(pics), _ptr(self.scratch),
                                       self.scratch.numel(), self._stream_ptr(stream)))

    def redo_coeff16(self, st, d_data=None, d_pic_offset=None, stream=None):
        """Pictures whose status (host array [n, 4] from ``status_host``) is ST_COEFF16_RANGE are decoded
        again through the int32 plane, in place; ``st`` is updated.  Returns the pictures redone."""
        redone = [i for i in range(st.shape[0]) if st[i, 0] == capi.ST_COEFF16_RANGE]
        if not redone:
            return redone
        d_data = self.data if d_data is None else d_data
        d_pic_offset = self.pic_offset if d_pic_offset is None else d_pic_offset
        for i in redone:
            self.decode_device(d_data, d_pic_offset[i:], 1, stream, first=i, coeff16=False)
        with self._on(stream):
            fresh = self.status[: 4 * st.shape[0]].cpu().numpy().reshape(-1, 4)
        for i in redone:
            st[i] = fresh[i]
        return redone

    def check_status(self, n, stream=None):
        st = self.status_host(n, stream)
        for i in range(n):
            if st[i, 0] != capi.ST_OK:
                raise StreamError(i, int(st[i, 0]), int(st[i, 1]))
        return st

    def status_host(self, n, stream=None):
        with self._on(stream):
            return self.status[: 4 * n].cpu().numpy().reshape(n, 4)

    # -- encode --------------------------------------------------------------------------
    def upload_pictures(self, pictures, stream=None):
        """H2D copy of host pictures ([n, picture_samples] uint16 numpy) into ``self.pictures``."""
        torch = self.torch
        pictures = np.ascontiguousarray(pictures, np.uint16).reshape(-1, self.g.picture_samples)
        n = pictures.shape[0]
        assert n <= self.max_pictures
        host = torch.from_numpy(pictures.view(np.int16)).view(torch.uint16).pin_memory()
        with self._on(stream):
            self.pictures[: n * self.g.picture_samples].copy_(host.reshape(-1), non_blocking=True)
        return n

    def dwt_device(self, n, stream=None):
        """picture_encode for the n pictures resident in ``self.pictures`` -> ``self.coeffs``."""
        L, p = self.lib, self.p
        with self._on(stream):
            capi.check("vc2b_dwt",
                       L.vc2b_dwt(C.byref(p), _ptr(self.pictures), n, _ptr(self.coeffs), _ptr(self.scratch),
                                  self.scratch.numel(), self._stream_ptr(stream)))

    def dc_predict_device(self, n, inverse, stream=None):
        L, p = self.lib, self.p
        with self._on(stream):
            capi.check("vc2b_dc_prediction",
                       L.vc2b_dc_prediction(C.byref(p), _ptr(self.coeffs), n, 1 if inverse else 0,
                                            self._stream_ptr(stream)))

    def quantize_device(self, n, picture_bytes, minimum_qindex=0, stream=None, predict=True):
        """quantize_to_fit for every slice (LD: dc prediction is applied first, in place, unless the
        coefficients already carry it: ``predict=False``)."""
        L, p = self.lib, self.p
        with self._on(stream):
            sp = self._stream_ptr(stream)
            if p.is_ld and predict:
                capi.check("vc2b_dc_prediction", L.vc2b_dc_prediction(C.byref(p), _ptr(self.coeffs), n, 1, sp))
            capi.check("vc2b_quantize_to_fit",
                       L.vc2b_quantize_to_fit(C.byref(p), _ptr(self.coeffs), n, int(picture_bytes),
                                              int(minimum_qindex), _ptr(self.qindex), _ptr(self.slice_bits), sp))

    def slice_bits_device(self, n, qindex=0, stream=None):
        """calculate_coeffs_bits of every slice block at a fixed ``qindex`` (no search); fills
        ``self.qindex`` / ``self.slice_bits``."""
        L, p = self.lib, self.p
        with self._on(stream):
            capi.check("vc2b_slice_bits",
                       L.vc2b_slice_bits(C.byref(p), _ptr(self.coeffs), n, int(qindex), None, _ptr(self.qindex),
                                         _ptr(self.slice_bits), self._stream_ptr(stream)))

    def quantized_host(self, n, stream=None):
        """quantize_coeffs of whole pictures at the per-slice indices in ``self.qindex``: host int32
        array [n, picture_coeffs] in the coefficient layout."""
        torch = self.torch
        L, p, g = self.lib, self.p, self.g
        with self._on(stream):
            out = torch.empty(n * g.picture_coeffs, dtype=torch.int32, device=self.device)
            capi.check("vc2b_quantize_slices",
                       L.vc2b_quantize_slices(C.byref(p), _ptr(self.coeffs), n, _ptr(self.qindex), 0, _ptr(out),
                                              self._stream_ptr(stream)))
            return out.cpu().numpy().reshape(n, g.picture_coeffs)

    def pack_device(self, n, picture_bytes, stream=None):
        """Serialise the quantised slices; returns (device uint8 tensor [n, stride], bytes per picture)."""
        torch = self.torch
        L, p = self.lib, self.p
        nbytes = int(L.vc2b_packed_bytes(C.byref(p), int(picture_bytes)))
        if nbytes < 0:
            raise capi.Vc2bError("vc2b_packed_bytes", nbytes)
        stride = (nbytes + 3 + 8) & ~3
        with self._on(stream):
            out = torch.empty(n * stride, dtype=torch.uint8, device=self.device)
            capi.check("vc2b_pack_slices",
                       L.vc2b_pack_slices(C.byref(p), _ptr(self.coeffs), n, int(picture_bytes), _ptr(self.qindex),
                                          _ptr(self.slice_bits), _ptr(out), stride, self._stream_ptr(stream)))
        return out.view(n, stride), nbytes

    def encode_pictures(self, pictures, picture_bytes, minimum_qindex=0, stream=None):
        """End-to-end encode of host pictures ([n, picture_samples] uint16 numpy): H2D, DWT,
        qindex search, slice packing.  Returns list of bytes (transform data per picture)."""
        n = self.upload_pictures(pictures, stream)
        self.dwt_device(n, stream)
        self.quantize_device(n, picture_bytes, minimum_qindex, stream)
        out, nbytes = self.pack_device(n, picture_bytes, stream)
        with self._on(stream):
            host_out = out.cpu().numpy()  # ordered after the pack kernel: same stream
        return [host_out[i, :nbytes].tobytes() for i in range(n)]

    def lossless_layout(self, n, minimum_slice_size_scaler=1, stream=None):
        """make_transform_data_hq_lossless (encoder/pictures.py:528) for the n pictures whose coefficients
        are in ``self.coeffs``: block sizes at qindex 0, the slice_size_scaler the reference would pick
        (ONE scaler for the batch: the largest any picture needs) and every slice's byte offset.
        Returns (slice_size_scaler, codec for that scaler, device offsets [n, slices+1], host totals [n])."""
        torch = self.torch
        L, g = self.lib, self.g
        if self.p.is_ld:
            raise ValueError("lossless coding is an HQ-profile mode")
        from .params import copy_params

        with self._on(stream):
            sp = self._stream_ptr(stream)
            p1 = copy_params(self.p, slice_size_scaler=1)
            capi.check("vc2b_slice_bits",
                       L.vc2b_slice_bits(C.byref(p1), _ptr(self.coeffs), n, 0, None, _ptr(self.qindex),
                                         _ptr(self.slice_bits), sp))
            offs = torch.empty(n * (g.num_slices + 1), dtype=torch.int32, device=self.device)
            maxlen = torch.empty(n, dtype=torch.int32, device=self.device)
            capi.check("vc2b_hq_lossless_layout",
                       L.vc2b_hq_lossless_layout(C.byref(p1), _ptr(self.slice_bits), n, _ptr(offs), _ptr(maxlen), sp))
            max_length = int(maxlen.max().item())
            scaler = max(1, int(minimum_slice_size_scaler), (max_length + 254) // 255)
            ps = copy_params(self.p, slice_size_scaler=scaler)
            capi.check("vc2b_hq_lossless_layout",
                       L.vc2b_hq_lossless_layout(C.byref(ps), _ptr(self.slice_bits), n, _ptr(offs), None, sp))
            offs2 = offs.view(n, g.num_slices + 1)
            totals = offs2[:, g.num_slices].cpu().numpy().astype(np.int64) & 0xffffffff
        return scaler, ps, offs2, totals

    def encode_pictures_lossless(self, pictures, minimum_slice_size_scaler=1, stream=None):
        """Lossless HQ encode of host pictures: H2D, DWT, minimal length fields, slice packing.
        Returns (slice_size_scaler, list of bytes)."""
        torch = self.torch
        L = self.lib
        n = self.upload_pictures(pictures, stream)
        self.dwt_device(n, stream)
        scaler, ps, offs, totals = self.lossless_layout(n, minimum_slice_size_scaler, stream)
        stride = (int(totals.max()) + 3 + 8) & ~3
        with self._on(stream):
            out = torch.empty(n * stride, dtype=torch.uint8, device=self.device)
            capi.check("vc2b_pack_slices_lossless",
                       L.vc2b_pack_slices_lossless(C.byref(ps), _ptr(self.coeffs), n, _ptr(self.qindex),
                                                   _ptr(self.slice_bits), _ptr(offs), _ptr(out), stride,
                                                   self._stream_ptr(stream)))
            host_out = out.cpu().numpy().reshape(n, stride)
        return scaler, [host_out[i, : int(totals[i])].tobytes() for i in range(n)]

    # -- views -----------------------------------------------------------------------------
    def picture_planes(self, i):
        """Host copies of picture i as [Y, C1, C2] uint16 arrays."""
        g = self.g
        with self._on(None):
            flat = self.pictures[i * g.picture_samples:(i + 1) * g.picture_samples].cpu().numpy()
        return [flat[g.plane_off[c]: g.plane_off[c] + g.dims[c][0] * g.dims[c][1]].reshape(g.dims[c][1], g.dims[c][0])
                for c in range(3)]

    def coeff_components(self, i, coeff16=None):
        """Host copies of picture i's coefficients as [Y, C1, C2] int32 arrays (with the int16 plane: the
        level-0 band from the int32 buffer, everything else widened from the int16 buffer)."""
        g = self.g
        c16 = (self.coeff16 and i not in self._plane32) if coeff16 is None else (coeff16 and self.coeff16)
        with self._on(None):
            flat = self.coeffs[i * g.picture_coeffs:(i + 1) * g.picture_coeffs].cpu().numpy()
            if c16:
                wide = self.coeffs16[i * g.picture_coeffs:(i + 1) * g.picture_coeffs].cpu().numpy().astype(np.int32)
                for c in range(3):
                    b0 = g.bands[c][0]
                    lo, cnt = g.comp_off[c] + b0["off"], b0["w"] * b0["h"]
                    wide[lo:lo + cnt] = flat[lo:lo + cnt]
                flat = wide
        return [flat[g.comp_off[c]: g.comp_off[c] + g.comp_coeffs[c]] for c in range(3)]


class _null(object):
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _DeviceStream(object):
    def __init__(self, torch, device, stream):
        self.ctx = [torch.cuda.device(device)]
        if stream is not None:
            self.ctx.append(torch.cuda.stream(stream))

    def __enter__(self):
        for c in self.ctx:
            c.__enter__()
        return self

    def __exit__(self, *a):
        for c in reversed(self.ctx):
            c.__exit__(*a)
        return False


class SequenceDecoder(object):
    """End-to-end decode of a whole sequence held in HOST memory: the pictures' transform data
    goes host -> device, is decoded, and the pictures come back device -> host, in chunks
    pipelined over several CUDA streams so copies overlap kernels.

    This is the throughput form of the per-picture entry points the reference calls in
    decoder/stream.py:128-133 (picture_parse's transform_data, then picture_decode)."""

    def __init__(self, params, chunk=8, nstreams=3, max_chunk_bytes=1 << 20, device="cuda:0", coeff16=True):
        torch = _torch()
        self.torch = torch
        self.p = params
        self.chunk = int(chunk)
        self.device = torch.device(device)
        with torch.cuda.device(self.device):
            self.streams = [torch.cuda.Stream(device=self.device) for _ in range(nstreams)]
            self.codecs = [BatchCodec(params, self.chunk, max_chunk_bytes, device, group=self.chunk, coeff16=coeff16)
                           for _ in range(nstreams)]
        self.g = self.codecs[0].g
        self.status_host = None  # pinned [pictures, 4] int32: every picture's status words (decode)

    @property
    def launches(self):
        """Kernels launched by libvc2b200 in this process so far (vc2b_launch_count)."""
        return int(self.codecs[0].lib.vc2b_launch_count())

    def plan(self, host_bytes, offsets):
        """Pre-computes per-chunk views for a packed host buffer (pinned uint8 tensor with PAD
        bytes of slack) and its [n+1] byte offsets."""
        torch = self.torch
        offsets = np.asarray(offsets, np.int64)
        n = len(offsets) - 1
        chunks = []
        for i0 in range(0, n, self.chunk):
            i1 = min(n, i0 + self.chunk)
            b0, b1 = int(offsets[i0]), int(offsets[i1])
            b0a = b0 & ~15  # keep 16-byte alignment of the device copy
            rel = torch.from_numpy(offsets[i0:i1 + 1] - b0a).pin_memory()
            chunks.append((i0, i1 - i0, host_bytes[b0a:b1 + PAD], rel))
        return chunks

    def decode(self, chunks, out_host):
        """Enqueue the whole sequence; ``out_host`` is a pinned uint16 tensor
        [n, picture_samples] receiving the decoded pictures.  Returns after enqueueing; call
        ``synchronize()`` before reading ``out_host``."""
        torch = self.torch
        S = self.g.picture_samples
        ns = len(self.streams)
        total = sum(c[1] for c in chunks)
        if self.status_host is None or self.status_host.shape[0] < total:
            self.status_host = torch.zeros((total, 4), dtype=torch.int32).pin_memory()
        for k, (i0, n, hbytes, rel) in enumerate(chunks):
            s = self.streams[k % ns]
            codec = self.codecs[k % ns]
            codec.decode_host(hbytes, rel, n, s)
            with torch.cuda.stream(s):
                out_host[i0:i0 + n].copy_(codec.pictures[: n * S].view(n, S), non_blocking=True)
                self.status_host[i0:i0 + n].copy_(codec.status[: 4 * n].view(n, 4), non_blocking=True)

    def finish(self, chunks, out_host):
        """After ``synchronize()``: pictures that did not fit the int16 coefficient plane are decoded again
        through the int32 plane (synchronously) and written to ``out_host``.  Returns the host status array
        [pictures, 4] of the sequence."""
        torch = self.torch
        S = self.g.picture_samples
        total = sum(c[1] for c in chunks)
        st = self.status_host[:total].numpy()
        for k, (i0, n, hbytes, rel) in enumerate(chunks):
            bad = [i for i in range(n) if st[i0 + i, 0] == capi.ST_COEFF16_RANGE]
            if not bad:
                continue
            codec = self.codecs[0]
            s = self.streams[0]
            codec.upload_streams(hbytes, rel, n, s)
            for i in bad:
                codec.decode_device(codec.data, codec.pic_offset[i:], 1, s, first=i, coeff16=False)
                with torch.cuda.stream(s):
                    out_host[i0 + i].copy_(codec.pictures[i * S:(i + 1) * S], non_blocking=True)
                    self.status_host[i0 + i].copy_(codec.status[4 * i:4 * i + 4], non_blocking=True)
            s.synchronize()
        return st

    def fork_from(self, stream):
        """Make every worker stream wait for ``stream`` (start of a timed region)."""
        torch = self.torch
        ev = torch.cuda.Event()
        ev.record(stream)
        for s in self.streams:
            s.wait_event(ev)

    def join_to(self, stream):
        """Make ``stream`` wait for every worker stream (end of a timed region)."""
        torch = self.torch
        for s in self.streams:
            ev = torch.cuda.Event()
            ev.record(s)
            stream.wait_event(ev)

    def synchronize(self):
        for s in self.streams:
            s.synchronize()

    def statuses(self):
        return [c.status_host(self.chunk) for c in self.codecs]


class SequenceEncoder(object):
    """End-to-end encode of a sequence held in HOST memory, the mirror of :class:`SequenceDecoder`: pictures
    go host -> device in chunks, are transformed (picture_encode), quantised to fit (quantize_to_fit as driven
    by make_transform_data_hq_lossy / _ld_lossy, encoder/pictures.py:617,724) and packed, and the transform
    data comes back device -> host, pipelined over several CUDA streams."""

    def __init__(self, params, picture_bytes, chunk=1, nstreams=3, device="cuda:0", minimum_qindex=0):
        torch = _torch()
        self.torch = torch
        self.p = params
        self.picture_bytes = int(picture_bytes)
        self.minimum_qindex = int(minimum_qindex)
        self.chunk = int(chunk)
        self.device = torch.device(device)
        with torch.cuda.device(self.device):
            self.streams = [torch.cuda.Stream(device=self.device) for _ in range(nstreams)]
            self.codecs = [BatchCodec(params, self.chunk, 1 << 16, device, group=self.chunk) for _ in range(nstreams)]
        self.g = self.codecs[0].g
        nbytes = int(self.codecs[0].lib.vc2b_packed_bytes(C.byref(params), self.picture_bytes))
        if nbytes < 0:
            raise capi.Vc2bError("vc2b_packed_bytes", nbytes)
        self.stream_bytes = nbytes
        self.stride = (nbytes + 3 + 8) & ~3

    @property
    def launches(self):
        return int(self.codecs[0].lib.vc2b_launch_count())

    def encode(self, pictures_host, out_host):
        """``pictures_host``: pinned uint16 tensor [n, picture_samples]; ``out_host``: pinned uint8 tensor
        [n, stride] receiving ``stream_bytes`` bytes of transform data per picture.  Returns after
        enqueueing; call ``synchronize()`` before reading ``out_host``."""
        torch = self.torch
        n = pictures_host.shape[0]
        S = self.g.picture_samples
        ns = len(self.streams)
        for k, i0 in enumerate(range(0, n, self.chunk)):
            cnt = min(self.chunk, n - i0)
            s = self.streams[k % ns]
            codec = self.codecs[k % ns]
            with codec._on(s):
                codec.pictures[: cnt * S].copy_(pictures_host[i0:i0 + cnt].reshape(-1), non_blocking=True)
            codec.dwt_device(cnt, s)
            codec.quantize_device(cnt, self.picture_bytes, self.minimum_qindex, s)
            out, _nbytes = codec.pack_device(cnt, self.picture_bytes, s)
            with codec._on(s):
                out_host[i0:i0 + cnt].copy_(out, non_blocking=True)

    def fork_from(self, stream):
        ev = self.torch.cuda.Event()
        ev.record(stream)
        for s in self.streams:
            s.wait_event(ev)

    def join_to(self, stream):
        for s in self.streams:
            ev = self.torch.cuda.Event()
            ev.record(s)
            stream.wait_event(ev)

    def synchronize(self):
        for s in self.streams:
            s.synchronize()
