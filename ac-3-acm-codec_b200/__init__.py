"""ac-3-acm-codec_b200 - Python face of the B200 AC-3 engine.

Thin ctypes layer over the C ABI of liba52_b200.so (include/a52_batch.h,
include/a52.h).  The decode itself is hand-written CUDA for sm_100a
(csrc/a52_decode.cu); torch is used only as the owner of device buffers and
streams.  There is no CPU path: importing works without a GPU (so that the
symbol checks can run), every compute call needs one.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# (A52_B200_LIB: developer knob for A/B builds of the library; the default is the in-tree build)
LIB_PATH = os.environ.get("A52_B200_LIB") or os.path.join(_HERE, "liba52_b200.so")

A52_CHANNEL, A52_MONO, A52_STEREO, A52_3F, A52_2F1R, A52_3F1R, A52_2F2R, A52_3F2R = range(8)
A52_CHANNEL1, A52_CHANNEL2, A52_DOLBY = 8, 9, 10
A52_CHANNEL_MASK, A52_LFE, A52_ADJUST_LEVEL = 15, 16, 32
PCM_F32_PLANAR, PCM_F32_INTERLEAVED, PCM_S16_INTERLEAVED, PCM_S16_WAV = 0, 1, 2, 3
REQ_AS_CODED = 0x100          # include/a52_batch.h: every frame in its own coded mode (libao wav6)
DEVICE_PTRS = 1
DRC_STREAM, DRC_OFF, DRC_TABLE = 0, 1, 2
SLICES_AUTO, SLICES_CHAINED, SLICES_INDEPENDENT = 0, 1, 2
ST_OK, ST_BAD_SYNC, ST_BAD_FRAME, ST_BAD_BLOCK = 0, 1, 2, 16
_NFCH = [2, 1, 2, 3, 3, 4, 4, 5, 1, 1, 2]

EXPORTS = [
    "a52_init", "a52_samples", "a52_syncinfo", "a52_frame", "a52_dynrng", "a52_block", "a52_free",
    "a52_batch_create", "a52_batch_destroy", "a52_batch_last_error", "a52_batch_index", "a52_batch_index_device",
    "a52_batch_frame_stride", "a52_batch_decode", "a52_batch_set_max_frame_bytes", "a52_batch_set_max_stream_frames",
    "a52_batch_launch_count", "a52_batch_kernel_ms", "a52_batch_scan", "a52_batch_set_drc_table", "a52_batch_set_slice_mode", "a52_batch_violations", "a52_ab_imdct", "a52_ab_fp32_peak",
    "AC3_encode_init", "AC3_encode_frame",
    "ac3_batch_create", "ac3_batch_destroy", "ac3_batch_last_error", "ac3_batch_frame_bytes",
    "ac3_batch_encode", "ac3_batch_launch_count", "ac3_batch_kernel_ms",
    "ac3_wav_channel_map", "ac3_acm_bitrate", "ac3_acm_block_align",
    "ac3_stream_open", "ac3_stream_convert", "ac3_stream_frame_bytes", "ac3_stream_close",
]


class CarryStruct(C.Structure):
    _fields_ = [("dither_index", C.c_uint32), ("per_channel", C.c_uint32), ("reserved", C.c_uint32 * 2),
                ("delay", (C.c_float * 128) * 6)]


class FrameScanStruct(C.Structure):
    _fields_ = [("dither_draws", C.c_uint32), ("dynrng", (C.c_int16 * 2) * 6), ("status", C.c_int32)]


SCAN_DTYPE = np.dtype([("dither_draws", np.uint32), ("dynrng", np.int16, (6, 2)), ("status", np.int32)])


class DebugStruct(C.Structure):
    _fields_ = [("exp", C.c_void_p), ("bap", C.c_void_p), ("coef", C.c_void_p), ("info", C.c_void_p)]


def nout_of(flags):
    return _NFCH[flags & A52_CHANNEL_MASK] + (1 if flags & A52_LFE else 0)


_lib = None


def load_library():
    """Load liba52_b200.so; fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("liba52_b200.so is missing - build it with ac-3-acm-codec_b200/build.sh "
                           "(or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    L.a52_batch_create.restype = C.c_void_p
    L.a52_batch_create.argtypes = [C.c_int]
    L.a52_batch_destroy.argtypes = [C.c_void_p]
    L.a52_batch_last_error.restype = C.c_char_p
    L.a52_batch_last_error.argtypes = [C.c_void_p]
    L.a52_batch_index.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
    L.a52_batch_index_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                         C.c_void_p]
    L.a52_batch_frame_stride.restype = C.c_size_t
    L.a52_batch_frame_stride.argtypes = [C.c_int, C.c_int]
    L.a52_batch_set_max_frame_bytes.argtypes = [C.c_void_p, C.c_int]
    L.a52_batch_set_max_stream_frames.argtypes = [C.c_void_p, C.c_int]
    L.a52_batch_launch_count.restype = C.c_long
    L.a52_batch_launch_count.argtypes = [C.c_void_p]
    L.a52_batch_kernel_ms.restype = C.c_double
    L.a52_batch_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    L.a52_batch_decode.argtypes = [
        C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
        C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
        C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.a52_batch_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                 C.c_void_p, C.c_int, C.c_void_p]
    L.a52_batch_set_drc_table.argtypes = [C.c_void_p, C.c_void_p]
    L.a52_batch_set_slice_mode.argtypes = [C.c_void_p, C.c_int]
    L.a52_ab_fp32_peak.restype = C.c_double
    L.a52_ab_fp32_peak.argtypes = [C.c_void_p]
    L.a52_init.restype = C.c_void_p
    L.a52_init.argtypes = [C.c_uint32]
    L.a52_samples.restype = C.POINTER(C.c_float)
    L.a52_samples.argtypes = [C.c_void_p]
    L.a52_syncinfo.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.a52_frame.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float), C.c_float]
    L.a52_dynrng.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.a52_block.argtypes = [C.c_void_p]
    L.a52_free.argtypes = [C.c_void_p]
    # encoder
    L.AC3_encode_init.argtypes = [C.c_int, C.c_int, C.c_int]
    L.AC3_encode_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.ac3_batch_create.restype = C.c_void_p
    L.ac3_batch_create.argtypes = [C.c_int]
    L.ac3_batch_destroy.argtypes = [C.c_void_p]
    L.ac3_batch_last_error.restype = C.c_char_p
    L.ac3_batch_last_error.argtypes = [C.c_void_p]
    L.ac3_batch_frame_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    L.ac3_batch_launch_count.restype = C.c_long
    L.ac3_batch_launch_count.argtypes = [C.c_void_p]
    L.ac3_batch_kernel_ms.restype = C.c_double
    L.ac3_batch_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    L.ac3_batch_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    # encoder front end (include/ac3enc_batch.h)
    L.ac3_wav_channel_map.argtypes = [C.c_int, C.c_void_p]
    L.ac3_acm_bitrate.argtypes = [C.c_int, C.c_uint32]
    L.ac3_acm_block_align.argtypes = [C.c_int, C.c_int]
    L.ac3_stream_open.restype = C.c_void_p
    L.ac3_stream_open.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_int]
    L.ac3_stream_convert.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.c_void_p, C.c_uint32,
                                     C.POINTER(C.c_uint32), C.c_int]
    L.ac3_stream_frame_bytes.argtypes = [C.c_void_p]
    L.ac3_stream_close.argtypes = [C.c_void_p]
    _lib = L
    return L


def index_frames(es):
    """Frame offsets of an elementary stream (host side; a52dec.c:240-309 resync discipline)."""
    L = load_library()
    es = np.ascontiguousarray(es, dtype=np.uint8)
    cap = len(es) // 64 + 2
    off = np.zeros(cap, np.uint64)
    n = L.a52_batch_index(es.ctypes.data, len(es), off.ctypes.data, cap)
    return off[:n].copy()


class BatchDecoder:
    """Batched decoder context bound to one GPU (a52_batch_t)."""

    def __init__(self, device=0):
        self.L = load_library()
        self.ctx = self.L.a52_batch_create(device)
        if not self.ctx:
            raise RuntimeError("a52_batch_create failed: no usable CUDA device %d (there is no CPU fallback)" % device)
        self.device = device

    def close(self):
        if self.ctx:
            self.L.a52_batch_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("a52_batch_decode failed (%d): %s" % (rc, self.L.a52_batch_last_error(self.ctx).decode()))

    def decode_host(self, es, frame_off, stream_first, req_flags, level=1.0, bias=0.0, drc=DRC_STREAM,
                    out_fmt=PCM_F32_PLANAR, carry=None, want_debug=False):
        """Host buffers in, host buffers out (copies inside the call).

        Returns dict(pcm=raw array [nframes, stride/itemsize], status, flags, carry, debug...)."""
        es = np.ascontiguousarray(es, dtype=np.uint8)
        frame_off = np.ascontiguousarray(frame_off, dtype=np.uint64)
        stream_first = np.ascontiguousarray(stream_first, dtype=np.uint32)
        nframes, nstreams = len(frame_off), len(stream_first) - 1
        stride = self.L.a52_batch_frame_stride(req_flags, out_fmt)
        dt = np.int16 if out_fmt >= PCM_S16_INTERLEAVED else np.float32
        pcm = np.zeros((nframes, stride // np.dtype(dt).itemsize), dt)
        status = np.zeros(nframes, np.int32)
        flags = np.zeros(nframes, np.int32)
        cbuf = None
        if carry is not None:
            cbuf = (CarryStruct * nstreams)()
            for i, c in enumerate(carry):
                if c is not None:
                    C.memmove(C.byref(cbuf[i]), C.byref(c), C.sizeof(CarryStruct))
        dbg = None
        out = {}
        if want_debug:
            out["exp"] = np.zeros((nframes, 6, 7, 256), np.uint8)
            out["bap"] = np.zeros((nframes, 6, 7, 256), np.uint8)
            out["coef"] = np.zeros((nframes, 6, 6, 256), np.float32)
            out["info"] = np.zeros((nframes, 6, 16), np.int32)
            dbg = DebugStruct(out["exp"].ctypes.data, out["bap"].ctypes.data, out["coef"].ctypes.data,
                              out["info"].ctypes.data)
        rc = self.L.a52_batch_decode(
            self.ctx, es.ctypes.data, len(es), frame_off.ctypes.data, nframes, stream_first.ctypes.data,
            nstreams, req_flags, level, bias, drc, out_fmt, pcm.ctypes.data, status.ctypes.data,
            flags.ctypes.data, C.byref(cbuf) if cbuf is not None else None,
            C.byref(dbg) if dbg is not None else None, 0, None)
        self._check(rc)
        out.update(pcm=pcm, status=status, flags=flags, carry=cbuf)
        return out

    def scan_host(self, es, frame_off, stream_first, req_flags):
        """a52_batch_scan with host buffers: numpy record array (dither_draws, dynrng[6][2], status) per frame."""
        es = np.ascontiguousarray(es, dtype=np.uint8)
        frame_off = np.ascontiguousarray(frame_off, dtype=np.uint64)
        stream_first = np.ascontiguousarray(stream_first, dtype=np.uint32)
        out = np.zeros(len(frame_off), SCAN_DTYPE)
        assert SCAN_DTYPE.itemsize == C.sizeof(FrameScanStruct) == 32
        rc = self.L.a52_batch_scan(self.ctx, es.ctypes.data, len(es), frame_off.ctypes.data, len(frame_off),
                                   stream_first.ctypes.data, len(stream_first) - 1, req_flags, out.ctypes.data, 0, None)
        self._check(rc)
        return out

    def set_drc_table(self, ranges):
        """float32 [nframes][6][2] for the next DRC_TABLE call (kept alive here), or None."""
        self._drc = None if ranges is None else np.ascontiguousarray(ranges, dtype=np.float32)
        self.L.a52_batch_set_drc_table(self.ctx, self._drc.ctypes.data if self._drc is not None else None)

    def set_slice_mode(self, mode):
        self.L.a52_batch_set_slice_mode(self.ctx, mode)

    def decode_host_into(self, es_ptr, es_bytes, frame_off, stream_first, req_flags, pcm_ptr, status_ptr=0,
                         flags_ptr=0, level=1.0, bias=0.0, drc=DRC_STREAM, out_fmt=PCM_F32_INTERLEAVED):
        """Host pointers (ideally pinned) in and out, no allocation here: the host form of
        a52_batch_decode with caller-owned buffers.  Synchronous."""
        frame_off = np.ascontiguousarray(frame_off, dtype=np.uint64)
        stream_first = np.ascontiguousarray(stream_first, dtype=np.uint32)
        rc = self.L.a52_batch_decode(
            self.ctx, es_ptr, es_bytes, frame_off.ctypes.data, len(frame_off), stream_first.ctypes.data,
            len(stream_first) - 1, req_flags, level, bias, drc, out_fmt, pcm_ptr, status_ptr or None,
            flags_ptr or None, None, None, 0, None)
        self._check(rc)

    def decode_device(self, es_ptr, es_bytes, off_ptr, nframes, first_ptr, nstreams, req_flags, pcm_ptr,
                      status_ptr=0, flags_ptr=0, carry_ptr=0, level=1.0, bias=0.0, drc=DRC_STREAM,
                      out_fmt=PCM_F32_INTERLEAVED, stream=0):
        """Device pointers (e.g. torch tensor .data_ptr()); asynchronous on `stream`.

        off_ptr must hold nframes + 1 offsets (last = es_bytes)."""
        rc = self.L.a52_batch_decode(
            self.ctx, es_ptr, es_bytes, off_ptr, nframes, first_ptr, nstreams, req_flags, level, bias, drc,
            out_fmt, pcm_ptr, status_ptr or None, flags_ptr or None, carry_ptr or None, None, DEVICE_PTRS,
            stream or None)
        self._check(rc)

    def index_device(self, es_ptr, stream_off, frame_off_ptr, max_frames, stream_first_ptr, stream=0):
        """GPU frame indexer over device-resident elementary streams; returns the number of frames."""
        stream_off = np.ascontiguousarray(stream_off, dtype=np.uint64)
        n = self.L.a52_batch_index_device(self.ctx, es_ptr, stream_off.ctypes.data, len(stream_off) - 1, frame_off_ptr,
                                          max_frames, stream_first_ptr, stream or None)
        if n < 0:
            raise RuntimeError("a52_batch_index_device failed (%d): %s" % (n, self.L.a52_batch_last_error(self.ctx).decode()))
        return n

    def set_max_frame_bytes(self, n):
        self.L.a52_batch_set_max_frame_bytes(self.ctx, n)

    def set_max_stream_frames(self, n):
        self.L.a52_batch_set_max_stream_frames(self.ctx, n)

    def launch_count(self):
        return self.L.a52_batch_launch_count(self.ctx)

    def kernel_ms(self):
        n = C.c_int(0)
        ms = self.L.a52_batch_kernel_ms(self.ctx, C.byref(n))
        return ms, n.value

    def frame_stride(self, req_flags, out_fmt):
        return self.L.a52_batch_frame_stride(req_flags, out_fmt)


class EncCarryStruct(C.Structure):
    _fields_ = [("last_samples", (C.c_int16 * 256) * 6), ("csnroffst", C.c_int32), ("started", C.c_int32),
                ("reserved", C.c_int32 * 2)]


class EncDebugStruct(C.Structure):
    _fields_ = [("coef", C.c_void_p), ("exp_shift", C.c_void_p), ("strategy", C.c_void_p),
                ("encoded_exp", C.c_void_p), ("bap", C.c_void_p), ("snr", C.c_void_p)]


class BatchEncoder:
    """Batched encoder context bound to one GPU (ac3_batch_t, include/ac3enc_batch.h)."""

    def __init__(self, device=0):
        self.L = load_library()
        self.ctx = self.L.ac3_batch_create(device)
        if not self.ctx:
            raise RuntimeError("ac3_batch_create failed: no usable CUDA device %d (there is no CPU fallback)" % device)

    def close(self):
        if self.ctx:
            self.L.ac3_batch_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def frame_bytes(self, freq, bitrate, channels):
        return self.L.ac3_batch_frame_bytes(freq, bitrate, channels)

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("ac3_batch_encode failed (%d): %s" % (rc, self.L.ac3_batch_last_error(self.ctx).decode()))

    def encode_host(self, pcm, freq, bitrate, chmap=None, carry=None, want_debug=False):
        """pcm: int16 [nstreams, nframes * 1536, channels].  Returns dict(frames [nstreams, nframes, fb], status...)."""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        ns, nsamp, nch = pcm.shape
        nfr = nsamp // 1536
        pcm = np.ascontiguousarray(pcm[:, : nfr * 1536])
        fb = self.frame_bytes(freq, bitrate, nch)
        if fb <= 0:
            raise ValueError("encoder rejected config")
        out = np.zeros((ns, nfr, fb), np.uint8)
        status = np.zeros((ns, nfr), np.int32)
        cm = np.ascontiguousarray(chmap, dtype=np.uint8) if chmap is not None else None
        cbuf = None
        if carry is not None:
            cbuf = (EncCarryStruct * ns)()
            for i, c in enumerate(carry):
                if c is not None:
                    C.memmove(C.byref(cbuf[i]), C.byref(c), C.sizeof(EncCarryStruct))
        res = {}
        dbg = None
        if want_debug:
            res["coef"] = np.zeros((ns, nfr, 6, 6, 256), np.int32)
            res["exp_shift"] = np.zeros((ns, nfr, 6, 6), np.int8)
            res["strategy"] = np.zeros((ns, nfr, 6, 6), np.uint8)
            res["encoded_exp"] = np.zeros((ns, nfr, 6, 6, 256), np.uint8)
            res["bap"] = np.zeros((ns, nfr, 6, 6, 256), np.uint8)
            res["snr"] = np.zeros((ns, nfr, 2), np.int32)
            dbg = EncDebugStruct(*[res[k].ctypes.data for k in ("coef", "exp_shift", "strategy", "encoded_exp", "bap", "snr")])
        rc = self.L.ac3_batch_encode(self.ctx, pcm.ctypes.data, ns, nfr, freq, bitrate, nch,
                                     cm.ctypes.data if cm is not None else None, out.ctypes.data, status.ctypes.data,
                                     C.byref(cbuf) if cbuf is not None else None,
                                     C.byref(dbg) if dbg is not None else None, 0, None)
        self._check(rc)
        res.update(frames=out, status=status, carry=cbuf, frame_bytes=fb)
        return res

    def encode_device(self, pcm_ptr, nstreams, nframes, freq, bitrate, channels, out_ptr, status_ptr=0, carry_ptr=0,
                      chmap=None, stream=0):
        cm = np.ascontiguousarray(chmap, dtype=np.uint8) if chmap is not None else None
        rc = self.L.ac3_batch_encode(self.ctx, pcm_ptr, nstreams, nframes, freq, bitrate, channels,
                                     cm.ctypes.data if cm is not None else None, out_ptr, status_ptr or None,
                                     carry_ptr or None, None, DEVICE_PTRS, stream or None)
        self._check(rc)

    def launch_count(self):
        return self.L.ac3_batch_launch_count(self.ctx)

    def kernel_ms(self):
        n = C.c_int(0)
        ms = self.L.ac3_batch_kernel_ms(self.ctx, C.byref(n))
        return ms, n.value


def wav_channel_map(channels):
    """chmap[coded channel] = WAVE-order source channel (create_channel_map, AC3ACM.cpp:1631-1662)."""
    m = np.zeros(6, np.uint8)
    if load_library().ac3_wav_channel_map(channels, m.ctypes.data):
        raise ValueError("unsupported channel count %d" % channels)
    return m[:channels]


class PcmToAc3Stream:
    """The ACM wrapper's PCM -> AC-3 stream conversion (stream_convert_pcm, AC3ACM.cpp:1665-1798) on the
    batched encoder: convert(src bytes, dst capacity, start) -> (bytes of src consumed, bytes produced)."""

    def __init__(self, encoder, freq, avg_bytes_per_sec, channels):
        self.L = load_library()
        self.enc = encoder
        self.h = self.L.ac3_stream_open(encoder.ctx, freq, avg_bytes_per_sec, channels)
        if not self.h:
            raise ValueError("format refused (as the ACM stream open would)")
        self.frame_bytes = self.L.ac3_stream_frame_bytes(self.h)

    def convert(self, src, dst_len, start=False):
        src = np.ascontiguousarray(np.frombuffer(bytes(src), np.uint8))
        dst = np.zeros(max(dst_len, 1), np.uint8)
        su, du = C.c_uint32(0), C.c_uint32(0)
        rc = self.L.ac3_stream_convert(self.h, src.ctypes.data if len(src) else None, len(src), C.byref(su),
                                       dst.ctypes.data, dst_len, C.byref(du), 1 if start else 0)
        if rc:
            raise RuntimeError("ac3_stream_convert failed (%d): %s" % (rc, self.L.ac3_batch_last_error(self.enc.ctx).decode()))
        return su.value, bytes(dst[:du.value])

    def close(self):
        if self.h:
            self.L.ac3_stream_close(self.h)
            self.h = None

    def __del__(self):
        self.close()
