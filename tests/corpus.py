"""The config-2 bench corpus: 256 unique 10 s 5.1 streams, bit-reproducible anywhere.

Per channel: 3 sines (f in [80, 16000] Hz, amplitude 0.1 .. 0.3; the LFE channel f in [20, 120] Hz) + white
noise of amplitude 0.02, scaled by 0.9 to int16 - the recipe of tests/synth.py, restated in INTEGER arithmetic
(a 4096-entry sine table addressed by 32-bit phase accumulators, a counter-based hash for the noise) so that
numpy on any host and torch on the GPU produce the same samples bit for bit: the reference encoder
(oracle/_ref, CPU arm) and this repo's encoder (GPU arm) then emit the same bytes, which both arms of
bench.py check against the committed digest tests/golden/c2_corpus.json.

Parameters come from the LCG x <- 1664525 x + 1013904223, seed 0xA52C0000 + 65536 * 2 + stream.
"""
import hashlib
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
UNIQUE = 256
FRAMES = 313
NCH = 6
RATE = 48000
FRAME_BYTES = 1792
DIGEST_PATH = os.path.join(HERE, "golden", "c2_corpus.json")

_TABLE = None


def sine_table():
    """int32 [4096]: round(32767 * sin(2 pi k / 4096))."""
    global _TABLE
    if _TABLE is None:
        k = np.arange(4096, dtype=np.float64)
        _TABLE = np.round(32767.0 * np.sin(2.0 * np.pi * k / 4096.0)).astype(np.int64)
    return _TABLE


def stream_params(stream, config=2):
    """(step[6][3], phase[6][3], amp[6][3], noise_seed[6]) as python ints."""
    x = (0xA52C0000 + 65536 * config + stream) & 0xFFFFFFFF

    def nxt():
        nonlocal x
        x = (1664525 * x + 1013904223) & 0xFFFFFFFF
        return x

    step, phase, amp, seeds = [], [], [], []
    for ch in range(NCH):
        lo, hi = (20.0, 120.0) if ch == 5 else (80.0, 16000.0)
        s3, p3, a3 = [], [], []
        for _ in range(3):
            f = lo * (hi / lo) ** (nxt() / 4294967296.0)
            s3.append(int(f / RATE * 4294967296.0) & 0xFFFFFFFF)
            a3.append(int((0.1 + 0.2 * (nxt() / 4294967296.0)) * 0.9 * 32767.0))
            p3.append(nxt())
        step.append(s3)
        phase.append(p3)
        amp.append(a3)
        seeds.append(nxt())
    return step, phase, amp, seeds


def _params_arrays(streams):
    ps = [stream_params(s) for s in streams]
    step = np.array([p[0] for p in ps], np.int64)      # [S][6][3]
    phase = np.array([p[1] for p in ps], np.int64)
    amp = np.array([p[2] for p in ps], np.int64)
    seed = np.array([p[3] for p in ps], np.int64)      # [S][6]
    return step, phase, amp, seed


NOISE_AMP = 590          # 0.02 * 0.9 * 32767


def synth_numpy(streams, nsamples=FRAMES * 1536):
    """int16 [S][nsamples][6] (coded channel order L C R LS RS LFE), numpy."""
    step, phase, amp, seed = _params_arrays(streams)
    T = sine_table()
    n = np.arange(nsamples, dtype=np.int64)
    out = np.empty((len(streams), nsamples, NCH), np.int16)
    M = 0xFFFFFFFF
    for si in range(len(streams)):
        for ch in range(NCH):
            acc = np.zeros(nsamples, np.int64)
            for i in range(3):
                ph = (phase[si, ch, i] + step[si, ch, i] * n) & M
                acc += amp[si, ch, i] * T[ph >> 20]
            h = (seed[si, ch] + n) & M
            h ^= h >> 16
            h = (h * 0x85EBCA6B) & M
            h ^= h >> 13
            h = (h * 0xC2B2AE35) & M
            h ^= h >> 16
            noise = (((h & 0xFFFF) * (2 * NOISE_AMP)) >> 16) - NOISE_AMP
            v = (acc >> 15) + noise
            out[si, :, ch] = np.clip(v, -32767, 32767).astype(np.int16)
    return out


def synth_torch(streams, device, nsamples=FRAMES * 1536):
    """The same samples computed with torch integer ops on `device`: int16 [S][nsamples][6]."""
    import torch
    step, phase, amp, seed = _params_arrays(streams)
    T = torch.from_numpy(sine_table()).to(device)
    n = torch.arange(nsamples, dtype=torch.int64, device=device)[None, :]
    S = len(streams)
    out = torch.empty((S, nsamples, NCH), dtype=torch.int16, device=device)
    M = 0xFFFFFFFF
    B = 32                                              # streams per pass: bounds the int64 temporaries
    for s0 in range(0, S, B):
        s1 = min(S, s0 + B)
        for ch in range(NCH):
            acc = torch.zeros((s1 - s0, nsamples), dtype=torch.int64, device=device)
            for i in range(3):
                st = torch.from_numpy(step[s0:s1, ch, i]).to(device)[:, None]
                p0 = torch.from_numpy(phase[s0:s1, ch, i]).to(device)[:, None]
                a = torch.from_numpy(amp[s0:s1, ch, i]).to(device)[:, None]
                ph = (p0 + st * n) & M
                acc += a * T[ph >> 20]
            h = (torch.from_numpy(seed[s0:s1, ch]).to(device)[:, None] + n) & M
            h = h ^ (h >> 16)
            h = (h * 0x85EBCA6B) & M
            h = h ^ (h >> 13)
            h = (h * 0xC2B2AE35) & M
            h = h ^ (h >> 16)
            noise = (((h & 0xFFFF) * (2 * NOISE_AMP)) >> 16) - NOISE_AMP
            v = (acc >> 15) + noise
            out[s0:s1, :, ch] = torch.clamp(v, -32767, 32767).to(torch.int16)
    return out


def digest_of(frames):
    """sha256 over uint8 [S][F][1792] (numpy, C order)."""
    return hashlib.sha256(np.ascontiguousarray(frames, dtype=np.uint8).tobytes()).hexdigest()


def load_digest():
    with open(DIGEST_PATH) as f:
        return json.load(f)


def tile_index(nstreams, nframes, unique=UNIQUE, frames=FRAMES):
    """(base[nstreams], idx[nstreams][nframes]): stream s = unique stream s % unique, rotated by 7 * (s // unique) frames."""
    s = np.arange(nstreams)
    base = s % unique
    idx = (np.arange(nframes)[None, :] + 7 * (s // unique)[:, None]) % frames
    return base, idx


# ---- CPU encode of (a part of) the corpus with the reference encoder: checker code, CPU arms only ----
def _enc_worker(streams):
    import sys
    sys.path.insert(0, HERE)
    import refbind
    enc = refbind.RefAc3Enc() if refbind.have_ref() else refbind.OracleEnc()
    out = []
    for s in streams:
        pcm = synth_numpy([s])[0]
        fb, es = enc.encode_stream(np.ascontiguousarray(pcm), RATE, 448000)
        assert fb == FRAME_BYTES
        out.append(np.asarray(es, np.uint8).reshape(-1, FRAME_BYTES)[:FRAMES])
    return out


def encode_cpu(streams, procs=None):
    """uint8 [len(streams)][313][1792] by the reference encoder (or the oracle port), one process per core."""
    import multiprocessing as mp
    streams = list(streams)
    procs = procs or min(len(streams), len(os.sched_getaffinity(0)))
    chunks = [streams[i::procs] for i in range(procs)]
    with mp.get_context("fork").Pool(procs) as pool:
        res = pool.map(_enc_worker, chunks)
    out = np.empty((len(streams), FRAMES, FRAME_BYTES), np.uint8)
    for i, part in enumerate(res):
        for j, fr in enumerate(part):
            out[i + j * procs] = fr
    return out


# ---- config 5: the mixed corpus (any sample rate / channel count) ----
def mixed_cells():
    """(rate, channels, bitrate) cells of BASELINE.json configs[4] inside the reference encoder's feasible region
    (SURVEY.md Appendix C thresholds plus one bitrate step of margin)."""
    rates = [32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 384, 448, 512, 576, 640]
    floor = {32000: [32, 32, 48, 64, 80, 64], 44100: [32, 48, 64, 80, 112, 112], 48000: [32, 48, 80, 96, 112, 112]}   # (3/2+LFE at 48 kHz: 112 kb/s still fails on this corpus)
    cells = []
    for fs in (32000, 44100, 48000):
        for nch in range(1, 7):
            lo = rates.index(floor[fs][nch - 1]) + 1
            for br in rates[lo:]:
                cells.append((fs, nch, br * 1000))
    return cells


def synth_torch_cfg(seeds, nch, rate, nsamples, device):
    """int16 [len(seeds)][nsamples][nch]: the recipe of synth_torch for any channel count / sample rate (the last
    channel of a 6-channel layout is the LFE)."""
    import torch
    T = torch.from_numpy(sine_table()).to(device)
    n = torch.arange(nsamples, dtype=torch.int64, device=device)[None, :]
    out = torch.empty((len(seeds), nsamples, nch), dtype=torch.int16, device=device)
    M = 0xFFFFFFFF
    par = []
    for sd in seeds:
        x = (0xA52C0000 + 65536 * 5 + sd) & M
        row = []
        for ch in range(nch):
            lo, hi = (20.0, 120.0) if (nch == 6 and ch == 5) else (80.0, min(16000.0, 0.45 * rate))
            cur = []
            for _ in range(3):
                x = (1664525 * x + 1013904223) & M
                f = lo * (hi / lo) ** (x / 4294967296.0)
                x = (1664525 * x + 1013904223) & M
                a = int((0.1 + 0.2 * (x / 4294967296.0)) * 0.9 * 32767.0)
                x = (1664525 * x + 1013904223) & M
                cur.append((int(f / rate * 4294967296.0) & M, a, x))
            x = (1664525 * x + 1013904223) & M
            row.append((cur, x))
        par.append(row)
    for ch in range(nch):
        acc = torch.zeros((len(seeds), nsamples), dtype=torch.int64, device=device)
        for i in range(3):
            st = torch.tensor([p[ch][0][i][0] for p in par], dtype=torch.int64, device=device)[:, None]
            a = torch.tensor([p[ch][0][i][1] for p in par], dtype=torch.int64, device=device)[:, None]
            p0 = torch.tensor([p[ch][0][i][2] for p in par], dtype=torch.int64, device=device)[:, None]
            acc += a * T[((p0 + st * n) & M) >> 20]
        h = (torch.tensor([p[ch][1] for p in par], dtype=torch.int64, device=device)[:, None] + n) & M
        h = h ^ (h >> 16); h = (h * 0x85EBCA6B) & M; h = h ^ (h >> 13); h = (h * 0xC2B2AE35) & M; h = h ^ (h >> 16)
        noise = (((h & 0xFFFF) * (2 * NOISE_AMP)) >> 16) - NOISE_AMP
        out[:, :, ch] = torch.clamp((acc >> 15) + noise, -32767, 32767).to(torch.int16)
    return out
