"""Deterministic synthetic PCM for the corpora of SURVEY.md section 8(d).

LCG x <- 1664525*x + 1013904223 (mod 2^32), seed = 0xA52C0000 + 65536*config + stream.
Per channel: 3 sines (f in [80, 16000] Hz, amplitude 0.1..0.3) + white noise of amplitude 0.02;
the LFE channel (index 5 of a 6-channel layout) uses f in [20, 120] Hz.  Returns int16
[nsamples, nch] interleaved in coded order (L, C, R, LS, RS, LFE for 6 channels).
"""
import numpy as np


class LCG:
    def __init__(self, seed):
        self.x = seed & 0xFFFFFFFF

    def next(self):
        self.x = (1664525 * self.x + 1013904223) & 0xFFFFFFFF
        return self.x

    def uniform(self):
        return self.next() / 4294967296.0


def synth_pcm(config, stream, nch, nsamples, rate=48000, noise=0.02, bursts=False):
    g = LCG(0xA52C0000 + 65536 * config + stream)
    t = np.arange(nsamples, dtype=np.float64) / rate
    out = np.zeros((nsamples, nch), np.float64)
    rng = np.random.RandomState(g.next() & 0x7FFFFFFF)
    for ch in range(nch):
        lfe = (nch == 6 and ch == 5)
        x = np.zeros(nsamples)
        for _ in range(3):
            lo, hi = (20.0, 120.0) if lfe else (80.0, 16000.0)
            f = lo * (hi / lo) ** g.uniform()
            a = 0.1 + 0.2 * g.uniform()
            ph = 2 * np.pi * g.uniform()
            x += a * np.sin(2 * np.pi * f * t + ph)
        x += noise * (2.0 * rng.random_sample(nsamples) - 1.0)
        if bursts:
            # castanet-like bursts every 3..7 blocks
            pos = 0
            while pos < nsamples:
                pos += 256 * (3 + int(g.uniform() * 5))
                n = min(200, nsamples - pos)
                if n > 0:
                    x[pos:pos + n] += 0.5 * (2.0 * rng.random_sample(n) - 1.0) * np.exp(-np.arange(n) / 40.0)
        out[:, ch] = x
    out = np.clip(out, -1.0, 1.0)
    return np.round(out * 32767.0 * 0.9).astype(np.int16)
