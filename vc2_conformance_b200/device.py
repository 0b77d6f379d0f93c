"""
Device-resident batch decoder / encoder: the throughput API over libvc2b200.so.

PyTorch is used only for device memory, streams and (in bench.py) torch.distributed.
Every compute step is a hand-written CUDA kernel behind the C ABI (include/vc2_b200.h).
"""

import ctypes as C

import numpy as np

from . import capi
from .params import Geometry

PAD = 64  # readable slack after the stream bytes (the kernels fetch aligned words)


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


class BatchCodec(object):
    """Buffers and launch sequence for batches of up to ``max_pictures`` pictures that share one
    parameter set.  All device work is enqueued on ``stream`` (default: torch's current)."""

    def __init__(self, params, max_pictures, max_stream_bytes=None, device="cuda:0", group=128):
        torch = _torch()
        self.torch = torch
        self.lib = capi.lib()
        self.p = params
        self.g = Geometry(params)
        self.device = torch.device(device)
        self.max_pictures = int(max_pictures)
        g = self.g
        dev = self.device
        with torch.cuda.device(dev):
            self.coeffs = torch.empty(self.max_pictures * g.picture_coeffs, dtype=torch.int32, device=dev)
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
            self.data = torch.zeros((need + 4095) // 4096 * 4096, dtype=torch.uint8, device=self.device)

    @property
    def launches(self):
        """Kernels launched by libvc2b200 in this process so far (vc2b_launch_count)."""
        return int(self.lib.vc2b_launch_count())

    def _stream_ptr(self, stream):
        torch = self.torch
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        return C.c_void_p(s.cuda_stream)

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

    def decode_host(self, host, offs, n, stream=None):
        """H2D copy of packed streams (pinned) and decode; everything enqueued on ``stream``."""
        torch = self.torch
        self._reserve_data(host.numel())
        with (torch.cuda.stream(stream) if stream is not None else _null()):
            self.data[: host.numel()].copy_(host, non_blocking=True)
            self.pic_offset[: n + 1].copy_(offs, non_blocking=True)
        self.decode_device(self.data, self.pic_offset, n, stream)

    def decode_streams(self, streams, stream=None, check=True):
        """End-to-end decode of host streams: H2D, slice index, unpack+dequant, IDWT+clip+offset.
        Returns the device picture tensor view [n, picture_samples] (uint16)."""
        n = len(streams)
        host, offs = self.pack_host_streams(streams)
        self.decode_host(host, offs, n, stream)
        if check:
            self.check_status(n, stream)
        return self.pictures[: n * self.g.picture_samples].view(n, self.g.picture_samples)

    def decode_device(self, d_data, d_pic_offset, n, stream=None, first=0):
        """Decode ``n`` pictures whose transform data is already resident in HBM."""
        L, p, g = self.lib, self.p, self.g
        sp = self._stream_ptr(stream)
        coeffs = self.coeffs[first * g.picture_coeffs:]
        status = self.status[first * 4:]
        qindex = self.qindex[first * g.num_slices:]
        so = self.slice_offset[first * g.num_slices:]
        if not p.is_ld:
            capi.check("vc2b_hq_slice_offsets",
                       L.vc2b_hq_slice_offsets(C.byref(p), _ptr(d_data), _ptr(d_pic_offset), n, _ptr(so),
                                               _ptr(status), sp))
        capi.check("vc2b_unpack",
                   L.vc2b_unpack(C.byref(p), _ptr(d_data), _ptr(d_pic_offset), n, _ptr(so), _ptr(coeffs),
                                 _ptr(qindex), _ptr(status), sp))
        self.idwt_device(n, stream, first)

    def idwt_device(self, n, stream=None, first=0):
        L, p, g = self.lib, self.p, self.g
        sp = self._stream_ptr(stream)
        coeffs = self.coeffs[first * g.picture_coeffs:]
        pics = self.pictures[first * g.picture_samples:]
        capi.check("vc2b_idwt",
                   L.vc2b_idwt(C.byref(p), _ptr(coeffs), n, _ptr(pics), _ptr(self.scratch),
                               self.scratch.numel(), sp))
        levels = max(1, p.dwt_depth + p.dwt_depth_ho)

    def check_status(self, n, stream=None):
        torch = self.torch
        if stream is not None:
            stream.synchronize()
        st = self.status[: 4 * n].cpu().numpy().reshape(n, 4)
        for i in range(n):
            if st[i, 0] != capi.ST_OK:
                raise StreamError(i, int(st[i, 0]), int(st[i, 1]))
        return st

    def status_host(self, n):
        return self.status[: 4 * n].cpu().numpy().reshape(n, 4)

    # -- encode --------------------------------------------------------------------------
    def dwt_device(self, n, stream=None):
        """picture_encode for the n pictures resident in ``self.pictures`` -> ``self.coeffs``."""
        L, p = self.lib, self.p
        sp = self._stream_ptr(stream)
        capi.check("vc2b_dwt",
                   L.vc2b_dwt(C.byref(p), _ptr(self.pictures), n, _ptr(self.coeffs), _ptr(self.scratch),
                              self.scratch.numel(), sp))
        levels = max(1, p.dwt_depth + p.dwt_depth_ho)

    def quantize_device(self, n, picture_bytes, minimum_qindex=0, stream=None):
        """quantize_to_fit for every slice (LD: dc prediction is applied first, in place)."""
        L, p = self.lib, self.p
        sp = self._stream_ptr(stream)
        if p.is_ld:
            capi.check("vc2b_dc_prediction", L.vc2b_dc_prediction(C.byref(p), _ptr(self.coeffs), n, 1, sp))
        capi.check("vc2b_quantize_to_fit",
                   L.vc2b_quantize_to_fit(C.byref(p), _ptr(self.coeffs), n, int(picture_bytes),
                                          int(minimum_qindex), _ptr(self.qindex), _ptr(self.slice_bits), sp))

    def pack_device(self, n, picture_bytes, stream=None):
        """Serialise the quantised slices; returns (device uint8 tensor [n, stride], bytes per picture)."""
        torch = self.torch
        L, p = self.lib, self.p
        sp = self._stream_ptr(stream)
        nbytes = int(L.vc2b_packed_bytes(C.byref(p), int(picture_bytes)))
        if nbytes < 0:
            raise capi.Vc2bError("vc2b_packed_bytes", nbytes)
        stride = (nbytes + 3 + 8) & ~3
        out = torch.empty(n * stride, dtype=torch.uint8, device=self.device)
        capi.check("vc2b_pack_slices",
                   L.vc2b_pack_slices(C.byref(p), _ptr(self.coeffs), n, int(picture_bytes), _ptr(self.qindex),
                                      _ptr(self.slice_bits), _ptr(out), stride, sp))
        return out.view(n, stride), nbytes

    def encode_pictures(self, pictures, picture_bytes, minimum_qindex=0, stream=None):
        """End-to-end encode of host pictures ([n, picture_samples] uint16 numpy): H2D, DWT,
        qindex search, slice packing.  Returns list of bytes (transform data per picture)."""
        torch = self.torch
        pictures = np.ascontiguousarray(pictures, np.uint16).reshape(-1, self.g.picture_samples)
        n = pictures.shape[0]
        host = torch.from_numpy(pictures.view(np.int16)).view(torch.uint16).pin_memory()
        ctx = torch.cuda.stream(stream) if stream is not None else _null()
        with ctx:
            self.pictures[: n * self.g.picture_samples].copy_(host.reshape(-1), non_blocking=True)
        self.dwt_device(n, stream)
        self.quantize_device(n, picture_bytes, minimum_qindex, stream)
        out, nbytes = self.pack_device(n, picture_bytes, stream)
        host_out = out.cpu().numpy()
        return [host_out[i, :nbytes].tobytes() for i in range(n)]

    # -- views -----------------------------------------------------------------------------
    def picture_planes(self, i):
        """Host copies of picture i as [Y, C1, C2] uint16 arrays."""
        g = self.g
        flat = self.pictures[i * g.picture_samples:(i + 1) * g.picture_samples].cpu().numpy()
        return [flat[g.plane_off[c]: g.plane_off[c] + g.dims[c][0] * g.dims[c][1]].reshape(g.dims[c][1], g.dims[c][0])
                for c in range(3)]

    def coeff_components(self, i):
        g = self.g
        flat = self.coeffs[i * g.picture_coeffs:(i + 1) * g.picture_coeffs].cpu().numpy()
        return [flat[g.comp_off[c]: g.comp_off[c] + g.comp_coeffs[c]] for c in range(3)]


class _null(object):
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class SequenceDecoder(object):
    """End-to-end decode of a whole sequence held in HOST memory: the pictures' transform data
    goes host -> device, is decoded, and the pictures come back device -> host, in chunks
    pipelined over several CUDA streams so copies overlap kernels.

    This is the throughput form of the per-picture entry points the reference calls in
    decoder/stream.py:128-133 (picture_parse's transform_data, then picture_decode)."""

    def __init__(self, params, chunk=8, nstreams=3, max_chunk_bytes=1 << 20, device="cuda:0"):
        torch = _torch()
        self.torch = torch
        self.p = params
        self.chunk = int(chunk)
        self.device = torch.device(device)
        with torch.cuda.device(self.device):
            self.streams = [torch.cuda.Stream(device=self.device) for _ in range(nstreams)]
            self.codecs = [BatchCodec(params, self.chunk, max_chunk_bytes, device, group=self.chunk)
                           for _ in range(nstreams)]
        self.g = self.codecs[0].g

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
        for k, (i0, n, hbytes, rel) in enumerate(chunks):
            s = self.streams[k % ns]
            codec = self.codecs[k % ns]
            codec.decode_host(hbytes, rel, n, s)
            with torch.cuda.stream(s):
                out_host[i0:i0 + n].copy_(codec.pictures[: n * S].view(n, S), non_blocking=True)

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
