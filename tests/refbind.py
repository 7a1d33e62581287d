"""ctypes bindings to the checkers under oracle/ (TEST INFRASTRUCTURE ONLY).

oracle/_ref/liba52_ref.so and oracle/_ref/ac3enc_ref.so are the UNMODIFIED
reference compiled by oracle/Makefile; oracle/liboracle.so is our CPU
restatement.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE = os.path.join(ROOT, "oracle")

A52_CHANNEL, A52_MONO, A52_STEREO, A52_3F, A52_2F1R, A52_3F1R, A52_2F2R, A52_3F2R = range(8)
A52_CHANNEL1, A52_CHANNEL2, A52_DOLBY = 8, 9, 10
A52_CHANNEL_MASK = 15
A52_LFE = 16
A52_ADJUST_LEVEL = 32

NOUT_TBL = [2, 1, 2, 3, 3, 4, 4, 5, 1, 1, 2]


def nout_of(flags):
    return NOUT_TBL[flags & A52_CHANNEL_MASK] + (1 if flags & A52_LFE else 0)


_u8p = C.POINTER(C.c_uint8)
_i8p = C.POINTER(C.c_int8)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)


def _ptr(a, t):
    return a.ctypes.data_as(t)


class RefA52:
    """The reference decoder (liba52) behind ref_* forwarding symbols."""

    def __init__(self, path=None):
        path = path or os.path.join(ORACLE, "_ref", "liba52_ref.so")
        self.lib = L = C.CDLL(path)
        L.ref_a52_init.restype = C.c_void_p
        L.ref_a52_init.argtypes = [C.c_uint32]
        L.ref_a52_samples.restype = _fp
        L.ref_a52_samples.argtypes = [C.c_void_p]
        L.ref_a52_syncinfo.argtypes = [_u8p, _ip, _ip, _ip]
        L.ref_a52_frame.argtypes = [C.c_void_p, _u8p, _ip, _fp, C.c_float]
        L.ref_a52_dynrng.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_a52_block.argtypes = [C.c_void_p]
        L.ref_a52_free.argtypes = [C.c_void_p]
        L.ref_get_expbap.argtypes = [C.c_void_p, C.c_int, _u8p, _i8p]
        L.ref_get_info.argtypes = [C.c_void_p, _ip]
        L.ref_set_lfsr.argtypes = [C.c_void_p, C.c_int]
        L.ref_capture_begin.argtypes = [C.c_void_p]
        L.ref_capture_get.argtypes = [C.c_int, _fp, _ip, _ip]
        L.ref_bit_allocate.argtypes = [C.c_int] * 6 + [_i8p] + [C.c_int] * 5 + [_u8p, _i8p]
        L.ref_imdct.argtypes = [C.c_int, _fp, _fp, C.c_float]
        L.ref_decode_stream.restype = C.c_long
        L.ref_decode_stream.argtypes = [C.c_void_p, _u8p, C.c_long, C.c_int, C.c_float,
                                        C.c_float, _fp, C.c_int, C.c_int]

    def syncinfo(self, buf):
        b = np.ascontiguousarray(np.frombuffer(bytes(buf[:7]), dtype=np.uint8))
        fl, sr, br = C.c_int(0), C.c_int(0), C.c_int(0)
        n = self.lib.ref_a52_syncinfo(_ptr(b, _u8p), C.byref(fl), C.byref(sr), C.byref(br))
        return n, fl.value, sr.value, br.value

    def decode_stream(self, es, req_flags, level=1.0, bias=0.0, dynrng_off=False, want_pcm=True):
        """Decode a whole elementary stream.  Returns (nframes, pcm[nblocks, nout, 256])."""
        es = np.ascontiguousarray(es, dtype=np.uint8)
        n, fl, _, _ = self.syncinfo(es)
        nfr_max = len(es) // max(n, 1) + 1
        # granted output flags are only known after a52_frame; probe on a scratch state
        st = self.lib.ref_a52_init(0)
        f = C.c_int(req_flags)
        lv = C.c_float(level)
        pad = np.concatenate([es[:n], np.zeros(8, np.uint8)])
        rc = self.lib.ref_a52_frame(st, _ptr(pad, _u8p), C.byref(f), C.byref(lv), bias)
        self.lib.ref_a52_free(st)
        if rc:
            return -1, None
        nout = nout_of(f.value)
        out = np.zeros((nfr_max * 6, nout, 256), np.float32) if want_pcm else None
        buf = np.concatenate([es, np.zeros(8, np.uint8)])
        nf = self.lib.ref_decode_stream(None, _ptr(buf, _u8p), len(es), req_flags, level, bias,
                                        _ptr(out, _fp) if want_pcm else None, nout,
                                        1 if dynrng_off else 0)
        if nf < 0:
            return nf, out
        return nf, (out[: nf * 6] if want_pcm else None)

    def decode_dump(self, es, req_flags=A52_3F2R | A52_LFE, level=1.0, bias=0.0, dynrng_off=False,
                    lfsr=None):
        """Frame-by-frame decode recording exp/bap/coefficients per block.

        Returns a list (per frame) of dicts: status, blocks=[{exp[7,256], bap[7,256], info[16],
        coeffs=[(plane, kind, coef[256])...], pcm[nout,256]}].  which index: 0..4 fbw, 5 lfe, 6 cpl.
        """
        L = self.lib
        es = np.ascontiguousarray(es, dtype=np.uint8)
        buf = np.concatenate([es, np.zeros(8, np.uint8)])
        st = L.ref_a52_init(0)
        if lfsr is not None:
            L.ref_set_lfsr(st, lfsr)
        frames = []
        pos = 0
        while pos + 7 <= len(es):
            n, fl, sr, br = self.syncinfo(buf[pos:pos + 7])
            if n == 0 or pos + n > len(es):
                break
            f = C.c_int(req_flags)
            lv = C.c_float(level)
            fr = {"status": 0, "blocks": [], "flags_in": fl, "len": n}
            p = buf[pos:]
            rc = L.ref_a52_frame(st, _ptr(p, _u8p), C.byref(f), C.byref(lv), bias)
            fr["out_flags"] = f.value
            fr["level"] = lv.value
            if rc:
                fr["status"] = 1
                frames.append(fr)
                pos += n
                continue
            if dynrng_off:
                L.ref_a52_dynrng(st, None, None)
            nout = nout_of(f.value)
            for b in range(6):
                L.ref_capture_begin(st)
                rc = L.ref_a52_block(st)
                L.ref_capture_end()
                if rc:
                    fr["status"] = 2 + b
                    break
                blk = {}
                exp = np.zeros((7, 256), np.uint8)
                bap = np.zeros((7, 256), np.int8)
                for w in range(7):
                    L.ref_get_expbap(st, w, _ptr(exp[w], _u8p), _ptr(bap[w], _i8p))
                info = np.zeros(16, np.int32)
                L.ref_get_info(st, _ptr(info, _ip))
                blk["exp"], blk["bap"], blk["info"] = exp, bap, info
                cs = []
                for i in range(L.ref_capture_count()):
                    c = np.zeros(256, np.float32)
                    k, pl = C.c_int(0), C.c_int(0)
                    L.ref_capture_get(i, _ptr(c, _fp), C.byref(k), C.byref(pl))
                    cs.append((pl.value, k.value, c))
                blk["coeffs"] = cs
                sp = L.ref_a52_samples(st)
                blk["pcm"] = np.ctypeslib.as_array(sp, shape=(nout * 256,)).copy().reshape(nout, 256)
                fr["blocks"].append(blk)
            frames.append(fr)
            pos += n
        L.ref_a52_free(st)
        return frames

    def bit_allocate(self, fscod, halfrate, bai11, csnroffst, chbai, exp, end, deltba=None,
                     bndstart=0, start=0, fastleak=0, slowleak=0):
        e = np.zeros(256, np.uint8)
        e[: len(exp)] = exp
        bap = np.zeros(256, np.int8)
        if deltba is not None:
            d = np.ascontiguousarray(deltba, dtype=np.int8)
            self.lib.ref_bit_allocate(fscod, halfrate, bai11, csnroffst, chbai, 1, _ptr(d, _i8p),
                                      bndstart, start, end, fastleak, slowleak, _ptr(e, _u8p), _ptr(bap, _i8p))
        else:
            self.lib.ref_bit_allocate(fscod, halfrate, bai11, csnroffst, chbai, 2, None,
                                      bndstart, start, end, fastleak, slowleak, _ptr(e, _u8p), _ptr(bap, _i8p))
        return bap

    def imdct(self, kind, data, delay, bias=0.0):
        d = np.ascontiguousarray(data, dtype=np.float32).copy()
        dl = np.ascontiguousarray(delay, dtype=np.float32).copy()
        st = self.lib.ref_a52_init(0)  # makes sure the tables are initialised
        self.lib.ref_a52_free(st)
        self.lib.ref_imdct(kind, _ptr(d, _fp), _ptr(dl, _fp), bias)
        return d, dl


class RefAc3Enc:
    """The reference encoder (src/ac3enc/ac3enc.cpp).  Global singleton state: not thread-safe."""

    def __init__(self, path=None):
        path = path or os.path.join(ORACLE, "_ref", "ac3enc_ref.so")
        self.lib = L = C.CDLL(path)
        L.ref_ac3enc_stream.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_short), C.c_int,
                                        _u8p, _u8p]
        L.ref_ac3enc_init.argtypes = [C.c_int, C.c_int, C.c_int]
        L.ref_ac3enc_frame.argtypes = [_u8p, C.POINTER(C.c_short), _u8p]
        L.ref_ac3enc_get.argtypes = [C.c_int, C.c_void_p]

    def encode_stream(self, pcm, freq, bitrate, chmap=None):
        """pcm: int16 [nsamples, nch] interleaved; returns (frame_bytes, uint8[nframes*frame_bytes])."""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        nch = pcm.shape[1]
        nframes = pcm.shape[0] // 1536
        out = np.zeros(nframes * 3840 + 64, np.uint8)
        cm = None
        if chmap is not None:
            cm = np.ascontiguousarray(chmap, dtype=np.uint8)
        fb = self.lib.ref_ac3enc_stream(freq, bitrate, nch, pcm.ctypes.data_as(C.POINTER(C.c_short)),
                                        nframes, _ptr(cm, _u8p) if cm is not None else None,
                                        _ptr(out, _u8p))
        if fb <= 0:
            raise ValueError("reference encoder rejected config")
        return fb, out[: nframes * fb].copy()

    def get(self, what):
        shapes = {0: ((6, 6, 256), np.int32), 1: ((6, 6, 256), np.uint8), 2: ((6, 6), np.uint8),
                  3: ((6, 6, 256), np.uint8), 4: ((6, 6, 256), np.uint8), 5: ((6, 6), np.int8),
                  6: ((3,), np.int32)}
        sh, dt = shapes[what]
        a = np.zeros(sh, dt)
        self.lib.ref_ac3enc_get(what, a.ctypes.data_as(C.c_void_p))
        return a


def have_ref():
    return os.path.exists(os.path.join(ORACLE, "_ref", "liba52_ref.so")) and \
        os.path.exists(os.path.join(ORACLE, "_ref", "ac3enc_ref.so"))


class Oracle:
    """Our CPU restatement (oracle/liboracle.so), same call shapes as RefA52."""

    def __init__(self, path=None):
        path = path or os.path.join(ORACLE, "liboracle.so")
        self.lib = L = C.CDLL(path)
        L.ora_init.restype = C.c_void_p
        L.ora_samples.restype = _fp
        L.ora_samples.argtypes = [C.c_void_p]
        L.ora_syncinfo.argtypes = [_u8p, _ip, _ip, _ip]
        L.ora_frame.argtypes = [C.c_void_p, _u8p, _ip, _fp, C.c_float]
        L.ora_dynrng_off.argtypes = [C.c_void_p]
        L.ora_block.argtypes = [C.c_void_p]
        L.ora_free.argtypes = [C.c_void_p]
        L.ora_get_expbap.argtypes = [C.c_void_p, C.c_int, _u8p, _i8p]
        L.ora_get_info.argtypes = [C.c_void_p, _ip]
        L.ora_set_lfsr.argtypes = [C.c_void_p, C.c_int]
        L.ora_get_coeffs.argtypes = [C.c_void_p, _fp]
        L.ora_bit_allocate.argtypes = [C.c_int] * 6 + [_i8p] + [C.c_int] * 5 + [_u8p, _i8p]
        L.ora_imdct.argtypes = [C.c_int, _fp, _fp, C.c_float]
        L.ora_decode_stream.restype = C.c_long
        L.ora_decode_stream.argtypes = [C.c_void_p, _u8p, C.c_long, C.c_int, C.c_float,
                                        C.c_float, _fp, C.c_int, C.c_int]

    def syncinfo(self, buf):
        b = np.ascontiguousarray(np.frombuffer(bytes(buf[:7]), dtype=np.uint8))
        fl, sr, br = C.c_int(0), C.c_int(0), C.c_int(0)
        n = self.lib.ora_syncinfo(_ptr(b, _u8p), C.byref(fl), C.byref(sr), C.byref(br))
        return n, fl.value, sr.value, br.value

    def decode_stream(self, es, req_flags, level=1.0, bias=0.0, dynrng_off=False, want_pcm=True):
        es = np.ascontiguousarray(es, dtype=np.uint8)
        n, fl, _, _ = self.syncinfo(es)
        nfr_max = len(es) // max(n, 1) + 1
        st = self.lib.ora_init()
        f = C.c_int(req_flags)
        lv = C.c_float(level)
        pad = np.concatenate([es[:n], np.zeros(8, np.uint8)])
        rc = self.lib.ora_frame(st, _ptr(pad, _u8p), C.byref(f), C.byref(lv), bias)
        self.lib.ora_free(st)
        if rc:
            return -1, None
        nout = nout_of(f.value)
        out = np.zeros((nfr_max * 6, nout, 256), np.float32) if want_pcm else None
        buf = np.concatenate([es, np.zeros(8, np.uint8)])
        nf = self.lib.ora_decode_stream(None, _ptr(buf, _u8p), len(es), req_flags, level, bias,
                                        _ptr(out, _fp) if want_pcm else None, nout,
                                        1 if dynrng_off else 0)
        if nf < 0:
            return nf, out
        return nf, (out[: nf * 6] if want_pcm else None)

    def decode_dump(self, es, req_flags=A52_3F2R | A52_LFE, level=1.0, bias=0.0, dynrng_off=False,
                    lfsr=None):
        L = self.lib
        es = np.ascontiguousarray(es, dtype=np.uint8)
        buf = np.concatenate([es, np.zeros(8, np.uint8)])
        st = L.ora_init()
        if lfsr is not None:
            L.ora_set_lfsr(st, lfsr)
        frames = []
        pos = 0
        while pos + 7 <= len(es):
            n, fl, sr, br = self.syncinfo(buf[pos:pos + 7])
            if n == 0 or pos + n > len(es):
                break
            f = C.c_int(req_flags)
            lv = C.c_float(level)
            fr = {"status": 0, "blocks": [], "flags_in": fl, "len": n}
            p = buf[pos:]
            rc = L.ora_frame(st, _ptr(p, _u8p), C.byref(f), C.byref(lv), bias)
            fr["out_flags"] = f.value
            fr["level"] = lv.value
            if rc:
                fr["status"] = 1
                frames.append(fr)
                pos += n
                continue
            if dynrng_off:
                L.ora_dynrng_off(st)
            nout = nout_of(f.value)
            for b in range(6):
                rc = L.ora_block(st)
                if rc:
                    fr["status"] = 2 + b
                    break
                blk = {}
                exp = np.zeros((7, 256), np.uint8)
                bap = np.zeros((7, 256), np.int8)
                for w in range(7):
                    L.ora_get_expbap(st, w, _ptr(exp[w], _u8p), _ptr(bap[w], _i8p))
                info = np.zeros(16, np.int32)
                L.ora_get_info(st, _ptr(info, _ip))
                blk["exp"], blk["bap"], blk["info"] = exp, bap, info
                co = np.zeros((6, 256), np.float32)
                L.ora_get_coeffs(st, _ptr(co, _fp))
                blk["coef"] = co
                sp = L.ora_samples(st)
                blk["pcm"] = np.ctypeslib.as_array(sp, shape=(nout * 256,)).copy().reshape(nout, 256)
                fr["blocks"].append(blk)
            frames.append(fr)
            pos += n
        L.ora_free(st)
        return frames

    def bit_allocate(self, fscod, halfrate, bai11, csnroffst, chbai, exp, end, deltba=None,
                     bndstart=0, start=0, fastleak=0, slowleak=0):
        e = np.zeros(256, np.uint8)
        e[: len(exp)] = exp
        bap = np.zeros(256, np.int8)
        d = None
        if deltba is not None:
            d = np.ascontiguousarray(deltba, dtype=np.int8)
        self.lib.ora_bit_allocate(fscod, halfrate, bai11, csnroffst, chbai, 1 if d is not None else 2,
                                  _ptr(d, _i8p) if d is not None else None,
                                  bndstart, start, end, fastleak, slowleak, _ptr(e, _u8p), _ptr(bap, _i8p))
        return bap

    def imdct(self, kind, data, delay, bias=0.0):
        d = np.ascontiguousarray(data, dtype=np.float32).copy()
        dl = np.ascontiguousarray(delay, dtype=np.float32).copy()
        self.lib.ora_imdct(kind, _ptr(d, _fp), _ptr(dl, _fp), bias)
        return d, dl


class OracleEnc:
    """Our CPU restatement of the encoder (oracle/ac3enc_oracle.c), same call shapes as RefAc3Enc."""

    SHAPES = {0: ((6, 6, 256), np.int32), 1: ((6, 6, 256), np.uint8), 2: ((6, 6), np.uint8),
              3: ((6, 6, 256), np.uint8), 4: ((6, 6, 256), np.uint8), 5: ((6, 6), np.int8), 6: ((4,), np.int32)}

    def __init__(self, path=None):
        path = path or os.path.join(ORACLE, "liboracle.so")
        self.lib = L = C.CDLL(path)
        L.ora_enc_init.restype = C.c_void_p
        L.ora_enc_init.argtypes = [C.c_int, C.c_int, C.c_int]
        L.ora_enc_free.argtypes = [C.c_void_p]
        L.ora_enc_frame_bytes.argtypes = [C.c_void_p]
        L.ora_enc_frame.argtypes = [C.c_void_p, _u8p, C.POINTER(C.c_short), _u8p]
        L.ora_enc_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ora_enc_window.restype = C.POINTER(C.c_int16)
        L.ora_enc_window.argtypes = [C.c_void_p]
        L.ora_enc_stream.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_short), C.c_int, _u8p, _u8p]
        self.st = None

    def encode_stream(self, pcm, freq, bitrate, chmap=None):
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        nch = pcm.shape[1]
        nframes = pcm.shape[0] // 1536
        out = np.zeros(nframes * 3840 + 64, np.uint8)
        cm = np.ascontiguousarray(chmap, dtype=np.uint8) if chmap is not None else None
        fb = self.lib.ora_enc_stream(freq, bitrate, nch, pcm.ctypes.data_as(C.POINTER(C.c_short)), nframes,
                                     _ptr(cm, _u8p) if cm is not None else None, _ptr(out, _u8p))
        if fb <= 0:
            raise ValueError("encoder rejected config")
        return fb, out[: nframes * fb].copy()

    def init(self, freq, bitrate, channels):
        if self.st:
            self.lib.ora_enc_free(self.st)
        self.st = self.lib.ora_enc_init(freq, bitrate, channels)
        return self.lib.ora_enc_frame_bytes(self.st) if self.st else 0

    def frame(self, samples, chmap=None):
        samples = np.ascontiguousarray(samples, dtype=np.int16)
        out = np.zeros(3840 + 64, np.uint8)
        cm = np.ascontiguousarray(chmap, dtype=np.uint8) if chmap is not None else None
        n = self.lib.ora_enc_frame(self.st, _ptr(out, _u8p), samples.ctypes.data_as(C.POINTER(C.c_short)),
                                   _ptr(cm, _u8p) if cm is not None else None)
        return out[:n].copy()

    def get(self, what):
        sh, dt = self.SHAPES[what]
        a = np.zeros(sh, dt)
        self.lib.ora_enc_get(self.st, what, a.ctypes.data_as(C.c_void_p))
        return a

    def window(self):
        st = self.st or self.lib.ora_enc_init(48000, 192000, 2)
        return np.ctypeslib.as_array(self.lib.ora_enc_window(st), shape=(256,)).copy()
