"""Shared helpers for the parity tests."""
import numpy as np

# A/52 standard bap number -> liba52's private code (bit_allocate.c:57-60)
LIBA52_BAP = np.array([0, -1, -2, 3, -3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 14, 16])
NFCHANS = [2, 1, 2, 3, 3, 4, 4, 5]


def relrms(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    d = a - b
    return float(np.sqrt((d * d).mean()) / max(np.sqrt((b * b).mean()), 1e-30))


def s16_of(pcm_bias384):
    """libao's float(bias 384) -> int16 (convert2s16.c:33-41): bit pattern minus 0x43c00000 in wrapping int32
    arithmetic (what the compiled reference does), clamped."""
    i = np.asarray(pcm_bias384, np.float32).view(np.int32).astype(np.int64) - 0x43C00000
    i = (i + 2 ** 31) % 2 ** 32 - 2 ** 31
    return np.clip(i, -32768, 32767).astype(np.int16)


def frame_offsets(es, oracle):
    """Frame offsets by the a52dec.c:240-309 discipline, computed with the oracle's syncinfo."""
    off, pos = [], 0
    while pos + 7 <= len(es):
        n = oracle.syncinfo(es[pos:pos + 7])[0]
        if n == 0:
            pos += 1
            continue
        if pos + n > len(es):
            break
        off.append(pos)
        pos += n
    return np.array(off, np.uint64)
