#!/usr/bin/env python3
"""bench.py - decoded audio-seconds per second, batched 5.1 448 kb/s AC-3 decode on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # liba52 on the host cores

Workload (BASELINE.json configs[1]): 4096 independent 10 s streams (313 frames each) of
5.1 / 48 kHz / 448 kb/s AC-3 per GPU, full path (exponents, bit allocation, dequantisation,
IMDCT-512, stereo downmix), float32 stereo PCM out.  The corpus is SURVEY.md section 8(d)'s: 256 UNIQUE
streams (tests/corpus.py: 3 sines + white noise per channel, LCG-seeded, integer arithmetic so that any host
and the GPU synthesise the same samples), encoded at 448 kb/s - by this repo's encoder on the GPU in the
CUDA arm, by the unmodified reference encoder in the CPU arms; both must reproduce the committed digest
tests/golden/c2_corpus.json - and tiled x16 with a rotation of 7 frames per copy.
A "step" is one pass of the decode over the whole batch.  One line of JSON is printed by rank 0.

  value      : audio-s/s, bitstream already resident in HBM, PCM left in HBM (CUDA events, max over ranks)
  e2e        : the same through the C ABI with HOST buffers (pinned): H2D of the bitstream and
               D2H of the PCM inside the timed region
  roofline   : decode kernel vs the HBM roofline - algorithmic bytes (1792 B read + 12288 B
               written per frame) / mean kernel time measured live with CUDA events
  cpu_baseline: the unmodified reference (oracle/_ref, liba52 compiled from /root/reference) or the
               oracle port, one process per host core, on a bounded sample of the same workload
  extra      : supplementary figures measured the same way (device resident), each with its own roofline
               fraction: int16 output, the config-3 corpus (5.1 @ 640 kb/s with block switching, coupling,
               dynrng, delta bit allocation: tests/golden/c3_fixture.npz) to stereo and to 5.1, config 4 (encode)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENC_CPU_FRAMES = 32                 # frames per "stream" of the CPU encode sample
FRAME_BYTES = 1792                  # 5.1 @ 448 kb/s, 48 kHz (parse.c:119)
FRAME_SECONDS = 1536 / 48000.0
PCM_BYTES = 1536 * 2 * 4            # float32 stereo per frame
ALGO_BYTES_PER_FRAME = FRAME_BYTES + PCM_BYTES      # 14080
A52_STEREO, A52_ADJUST_LEVEL = 2, 32
REQ_FLAGS = A52_STEREO | A52_ADJUST_LEVEL
METRIC = "decoded audio-sec/sec, 5.1 448k AC-3, batched streams"
METRIC_ENC = "encoded audio-sec/sec, 5.1 448k AC-3, batched streams"
ENC_ALGO_BYTES_PER_FRAME = 1536 * 6 * 2 + 1792      # int16 5.1 PCM in + frame out
UNIT = "audio-s/s"


CACHE = os.environ.get("A52_BENCH_CACHE", "/tmp/a52_bench_c2_corpus.npy")


def _corpus_mod():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import corpus
    return corpus


def _cached_frames():
    """uint8 [256][313][1792] from the cache file when it reproduces the committed digest, else None."""
    corpus = _corpus_mod()
    try:
        fr = np.load(CACHE)
        if fr.shape == (corpus.UNIQUE, corpus.FRAMES, FRAME_BYTES) and corpus.digest_of(fr) == corpus.load_digest()["sha256"]:
            return fr
    except (OSError, ValueError):
        pass
    return None


def _store_cache(fr):
    try:
        tmp = CACHE + ".%d.tmp.npy" % os.getpid()
        np.save(tmp, fr)
        os.replace(tmp, CACHE)
    except OSError:
        pass


def unique_frames_cpu():
    """The 256 unique encoded streams, made on the host cores by the reference encoder (CPU arms)."""
    corpus = _corpus_mod()
    fr = _cached_frames()
    if fr is None:
        fr = corpus.encode_cpu(range(corpus.UNIQUE))
        assert corpus.digest_of(fr) == corpus.load_digest()["sha256"], "CPU-built corpus differs from the committed digest"
        _store_cache(fr)
    return fr


def unique_frames_gpu(eng, dev):
    """The same 256 streams made on the GPU: torch integer synthesis + this repo's batched encoder."""
    import torch
    corpus = _corpus_mod()
    fr = _cached_frames()
    if fr is not None:
        return torch.from_numpy(fr).to(dev)
    pcm = corpus.synth_torch(range(corpus.UNIQUE), dev)                        # int16 [256][313*1536][6]
    out = torch.zeros((corpus.UNIQUE, corpus.FRAMES, FRAME_BYTES), dtype=torch.uint8, device=dev)
    status = torch.zeros((corpus.UNIQUE, corpus.FRAMES), dtype=torch.int32, device=dev)
    enc = eng.BatchEncoder(dev.index or 0)
    enc.encode_device(pcm.data_ptr(), corpus.UNIQUE, corpus.FRAMES, 48000, 448000, 6, out.data_ptr(),
                      status_ptr=status.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    enc.close()
    assert int((status != 0).sum().item()) == 0
    host = out.cpu().numpy()
    assert corpus.digest_of(host) == corpus.load_digest()["sha256"], "GPU-built corpus differs from the committed digest"
    _store_cache(host)
    return out


def tile_corpus(frames, stream_ids, nframes):
    """[len(stream_ids)][nframes][1792]: global stream g = unique stream g % 256 starting 7 * (g // 256) frames in
    (numpy or torch).  stream_ids: the global ids of this rank's shard (an int = the first so many streams)."""
    corpus = _corpus_mod()
    ids = np.arange(stream_ids) if np.isscalar(stream_ids) else np.asarray(stream_ids)
    base = ids % corpus.UNIQUE
    idx = (np.arange(nframes)[None, :] + 7 * (ids // corpus.UNIQUE)[:, None]) % corpus.FRAMES
    if isinstance(frames, np.ndarray):
        return frames[base[:, None], idx]
    import torch
    return frames[torch.from_numpy(base).to(frames.device)[:, None], torch.from_numpy(idx).to(frames.device)]


# ---------------------------------------------------------------------------
# CPU arm: the reference decoder on the host cores (checker code: the only use of oracle/ here)
# ---------------------------------------------------------------------------
_cpu = {}


def _cpu_init(kind):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refbind
    _cpu["lib"] = refbind.RefA52() if kind == "reference" else refbind.Oracle()
    _cpu["out"] = np.zeros((313 * 6 + 8, 2, 256), np.float32)
    _cpu["first"] = (os.getpid() * 37) % _cpu["es"].shape[0]          # workers start at different streams


def _cpu_work(nstreams):
    import ctypes as C
    lib = _cpu["lib"]
    fn = lib.lib.ref_decode_stream if hasattr(lib.lib, "ref_decode_stream") else lib.lib.ora_decode_stream
    t0 = time.perf_counter()
    frames = 0
    out = _cpu["out"]
    nu = _cpu["es"].shape[0]
    for s in range(nstreams):
        es = _cpu["es"][(_cpu["first"] + s) % nu]
        nf = fn(None, es.ctypes.data_as(C.POINTER(C.c_uint8)), len(es), REQ_FLAGS, 1.0, 0.0,
                out.ctypes.data_as(C.POINTER(C.c_float)), 2, 0)
        assert nf == 313, nf
        frames += nf
    return frames, time.perf_counter() - t0


def _cpu_init_enc(kind):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refbind
    from synth import synth_pcm
    _cpu["enc"] = refbind.RefAc3Enc() if kind == "reference" else refbind.OracleEnc()
    _cpu["pcm"] = np.ascontiguousarray(synth_pcm(2, 0, 6, 1536 * ENC_CPU_FRAMES))
    _cpu["outbuf"] = np.zeros(ENC_CPU_FRAMES * 3840 + 64, np.uint8)


def _cpu_work_enc(nstreams):
    import ctypes as C
    e = _cpu["enc"]
    fn = e.lib.ref_ac3enc_stream if hasattr(e.lib, "ref_ac3enc_stream") else e.lib.ora_enc_stream
    pcm, out = _cpu["pcm"], _cpu["outbuf"]
    t0 = time.perf_counter()
    for _ in range(nstreams):
        fb = fn(48000, 448000, 6, pcm.ctypes.data_as(C.POINTER(C.c_short)), ENC_CPU_FRAMES, None,
                out.ctypes.data_as(C.POINTER(C.c_uint8)))
        assert fb == 1792
    return nstreams * ENC_CPU_FRAMES, time.perf_counter() - t0


class CpuArm:
    def __init__(self, workload="decode"):
        self.workload = workload
        self._init_common()
        if workload == "encode":
            self.pool = self.mp.get_context("fork").Pool(self.cores, initializer=_cpu_init_enc, initargs=(self.kind,))
            self.work = _cpu_work_enc
            self.pool.map(self.work, [1] * self.cores)
            return
        self._init_decode()

    def _init_common(self):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import refbind
        self.kind = "reference" if refbind.have_ref() else "port"
        try:
            self.cores = len(os.sched_getaffinity(0))
        except AttributeError:
            self.cores = os.cpu_count() or 1
        import multiprocessing as mp
        self.mp = mp
        self.work = _cpu_work

    def _init_decode(self):
        _cpu["es"] = unique_frames_cpu().reshape(256, -1)           # inherited by the forked workers
        self.pool = self.mp.get_context("fork").Pool(self.cores, initializer=_cpu_init, initargs=(self.kind,))
        self.pool.map(_cpu_work, [1] * self.cores)               # warm: page in, tables

    def run(self, streams_per_core):
        t0 = time.perf_counter()
        res = self.pool.map(self.work, [streams_per_core] * self.cores, chunksize=1)
        wall = time.perf_counter() - t0
        frames = sum(r[0] for r in res)
        return frames * FRAME_SECONDS, wall

    def calibrate(self, target_s):
        self.run(2)                                  # every worker initialised and warm
        secs, wall = self.run(8)
        per_stream = wall / 8.0
        return max(1, int(target_s / max(per_stream, 1e-4)))

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # the workers (forked, one per core) dlopen the reference themselves; load it here too, so that what this
    # arm runs is visible in the parent's maps as well
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refbind
    _keep = (refbind.RefA52(), refbind.RefAc3Enc()) if refbind.have_ref() else (refbind.Oracle(),)
    arm = CpuArm(args.workload)
    k = arm.calibrate(2.0)                                   # ~2 s of CPU work per core per step
    for _ in range(args.warmup):
        arm.run(k)
    t_total, a_total = 0.0, 0.0
    for _ in range(args.steps):
        a, w = arm.run(k)
        a_total += a
        t_total += w
    arm.close()
    v = a_total / t_total
    if args.workload == "encode":
        sample = "%d x %d frames of 5.1 PCM per core per step (%d cores), in memory" % (k, ENC_CPU_FRAMES, arm.cores)
    else:
        sample = "%d streams x 313 frames per core per step (%d cores), in memory, stereo float out" % (k, arm.cores)
    line = {
        "impl": "reference", "metric": METRIC_ENC if args.workload == "encode" else METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": arm.cores, "kind": arm.kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(args):
    if args.workload == "encode":
        return {"workload": "5.1 48 kHz int16 PCM -> 448 kbps AC-3 encode, %d independent %.1f s streams per GPU "
                            "(BASELINE.json configs[3])" % (args.streams, args.frames * FRAME_SECONDS),
                "streams_per_gpu": args.streams, "frames_per_stream": args.frames, "frame_bytes": FRAME_BYTES,
                "cache": "inputs+outputs per step (%.1f GB) exceed the 126 MB L2"
                % (args.streams * args.frames * ENC_ALGO_BYTES_PER_FRAME / 1e9)}
    s16 = getattr(args, "pcm", "f32") == "s16"
    return {"workload": "5.1 48 kHz 448 kbps decode, %d independent %.1f s streams per GPU, stereo downmix, "
                        "%s PCM (BASELINE.json configs[1]%s)"
                        % (args.streams, args.frames * FRAME_SECONDS, "int16" if s16 else "float32",
                           "; supplementary int16 variant" if s16 else ""),
            "streams_per_gpu": args.streams, "frames_per_stream": args.frames, "frame_bytes": FRAME_BYTES,
            "out": ("s16" if s16 else "f32") + " stereo interleaved",
            "cache": "inputs+outputs per step (%.1f GB) exceed the 126 MB L2"
            % (args.streams * args.frames * (FRAME_BYTES + (PCM_BYTES // 2 if s16 else PCM_BYTES)) / 1e9)}


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.rows.append((time.perf_counter(), ln.strip()))

    def window(self, t0, t1):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, ln in self.rows:
            if t < t0 or t > t1 + 0.15:
                continue
            p = [x.strip() for x in ln.split(",")]
            try:
                sm.append(float(p[0]))
                mx = max(mx, float(p[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        if self.proc:
            self.proc.terminate()


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def run_gpu(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    # CPU baseline first (fork-based pool: must happen before CUDA is initialised in this process)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        arm = CpuArm()
        k = arm.calibrate(args.cpu_seconds)
        a, w = arm.run(k)
        arm.close()
        cpu_baseline = {"value": a / w, "unit": UNIT, "cores": arm.cores, "kind": arm.kind,
                        "sample": "%d streams x 313 frames per core (%d cores, %.1f s wall), in memory, stereo float out"
                                  % (k, arm.cores, w)}

    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    eng = ge.load_engine()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the decode path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import importlib
    shard = importlib.import_module("ac3_acm_codec_b200.shard")

    S, F = args.streams, args.frames
    nframes = S * F
    dev = torch.device("cuda", local)
    uniq = unique_frames_gpu(eng, dev)                         # [256][313][1792] on the device, digest-checked
    es_bytes = nframes * FRAME_BYTES
    es = torch.zeros(es_bytes + 64, dtype=torch.uint8, device=dev)
    # weak scaling: the job is world x S streams (global ids 0 .. world*S-1, all of the same cost), cut by the
    # sharder into one contiguous range per rank; every rank decodes ITS streams, no data-path collective
    my_ids = shard.partition_streams(np.full(world * S, F, np.int64), world)[rank]
    assert len(my_ids) == S
    es[:es_bytes].copy_(tile_corpus(uniq, my_ids, F).reshape(-1))
    off = torch.arange(nframes + 1, dtype=torch.int64, device=dev) * FRAME_BYTES
    first = (torch.arange(S + 1, dtype=torch.int64, device=dev) * F).to(torch.int32)
    s16 = args.pcm == "s16"
    pcm = torch.empty(nframes * 1536 * 2, dtype=torch.int16 if s16 else torch.float32, device=dev)
    status = torch.zeros(nframes, dtype=torch.int32, device=dev)

    dec = eng.BatchDecoder(local)
    dec.set_max_frame_bytes(FRAME_BYTES)
    dec.set_max_stream_frames(F)
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        dec.decode_device(es.data_ptr(), es_bytes, off.data_ptr(), nframes, first.data_ptr(), S, REQ_FLAGS,
                          pcm.data_ptr(), status_ptr=status.data_ptr(),
                          out_fmt=eng.PCM_S16_INTERLEAVED if s16 else eng.PCM_F32_INTERLEAVED,
                          stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    assert int((status != 0).sum().item()) == 0, "decode reported frame errors"
    dec.kernel_ms()                                            # reset the kernel-time accumulator
    sampler = ClockSampler(local) if rank == 0 else None
    time.sleep(0.25)
    l0 = dec.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    t1 = time.perf_counter()
    ms = e0.elapsed_time(e1)
    ms = shard.max_over_ranks(ms)
    launches = dec.launch_count() - l0
    kms, kn = dec.kernel_ms()
    clocks = sampler.window(t0, t1) if sampler else None
    audio_s_per_step = nframes * FRAME_SECONDS * world
    value = audio_s_per_step * args.steps / (ms / 1e3)

    # sanity of the device result against the committed digest: energy of the first four unique streams as the
    # UNMODIFIED reference decoder produced them (tests/golden/c2_corpus.json)
    # per-rank record for the shard report: first / last global stream id, energy of the whole shard's PCM
    shard_energy = float((pcm.double() ** 2).sum().item()) / ((32768.0 ** 2) if s16 else 1.0)
    recs = shard.gather_records([float(my_ids[0]), float(my_ids[-1]), shard_energy, float(nframes)])
    if F == 313 and S >= 4 and rank == 0:
        want = _corpus_mod().load_digest()["energy_stereo_first4"]
        for k in range(4):
            got = float(((pcm[k * F * 3072:(k + 1) * F * 3072].double() / (32768.0 if s16 else 1.0)) ** 2).sum().item())
            assert abs(got - want[k]) / want[k] < (1e-3 if s16 else 1e-5), ("stream energy differs from the reference digest", k, got, want[k])

    extra = None
    if not args.no_extra and not s16:
        extra = run_extras(args, eng, dec, es, es_bytes, off, first, status, shard, barrier, world, dev)

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, eng, dec, es[:es_bytes], shard, barrier, world)

    if sampler:
        sampler.stop()
    if rank == 0:
        peaks, peak_src = None, "fallback"
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        peak = float(peaks["hbm_gbs"]) if peaks and peaks.get("hbm_gbs") else 6650.0
        if peaks and peaks.get("hbm_gbs"):
            peak_src = "measured"
        # the other roofline of SURVEY.md 8(d): 79.5 kFLOP per frame against the MEASURED FP32 FMA peak of this GPU
        fp32_peak = float(eng.load_library().a52_ab_fp32_peak(dec.ctx))
        algo = FRAME_BYTES + (PCM_BYTES // 2 if s16 else PCM_BYTES)
        achieved = nframes * algo / (kms / 1e3) / 1e9 if kms > 0 else 0.0
        traffic = None
        try:
            per_frame = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("dram_bytes_per_frame")
            traffic = per_frame * nframes if per_frame and not s16 else None     # ncu dram bytes per frame x frames per launch
        except (OSError, ValueError):
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "a52_decode_kernel", "kernel_ms": kms, "kernel_launches_timed": kn,
                         "algorithmic_bytes_per_launch": nframes * algo,
                         "fp32": {"peak_tflops_measured": fp32_peak, "algorithmic_flop_per_frame": 79500,
                                  "achieved_tflops": nframes * 79500 / (kms / 1e3) / 1e12 if kms > 0 else 0.0,
                                  "frac": (nframes * 79500 / (kms / 1e3) / 1e12 / fp32_peak) if (kms > 0 and fp32_peak > 0) else None,
                                  "note": "HBM is the slower roofline of the two: frac above is against it"}},
            "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "extra": extra,
            "shards": {"partition": "shard.partition_streams: contiguous ranges of global stream ids (equal costs)",
                       "ranks": [{"streams": [int(r[0]), int(r[1])], "frames": int(r[3]), "pcm_energy": r[2]} for r in recs]},
        }
        print(json.dumps(line))
    dec.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_gpu_encode(args):
    """Config 4: batched encode.  Same contract as the decode line (value = device resident, e2e = host buffers)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        arm = CpuArm("encode")
        k = arm.calibrate(args.cpu_seconds)
        a, w = arm.run(k)
        arm.close()
        cpu_baseline = {"value": a / w, "unit": UNIT, "cores": arm.cores, "kind": arm.kind,
                        "sample": "%d x %d frames of 5.1 PCM per core (%d cores, %.1f s wall), in memory"
                                  % (k, ENC_CPU_FRAMES, arm.cores, w)}
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    eng = ge.load_engine()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the encode path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import importlib
    shard = importlib.import_module("ac3_acm_codec_b200.shard")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from synth import synth_pcm
    S, F = args.streams, args.frames
    dev = torch.device("cuda", local)
    base = torch.from_numpy(np.stack([synth_pcm(2, s, 6, 1536 * 64) for s in range(4)]).reshape(4, 64, 1536 * 6)).to(dev)
    fidx = (torch.arange(F, device=dev)[None, :] + 7 * torch.arange(S, device=dev)[:, None]) % 64
    pcm = base[(torch.arange(S, device=dev) % 4)[:, None], fidx].contiguous()          # [S, F, 1536*6] int16
    out = torch.zeros((S, F, FRAME_BYTES), dtype=torch.uint8, device=dev)
    status = torch.zeros((S, F), dtype=torch.int32, device=dev)
    enc = eng.BatchEncoder(local)
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        enc.encode_device(pcm.data_ptr(), S, F, 48000, 448000, 6, out.data_ptr(), status_ptr=status.data_ptr(),
                          stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    assert int((status != 0).sum().item()) == 0
    enc.kernel_ms()
    sampler = ClockSampler(local) if rank == 0 else None
    time.sleep(0.25)
    l0 = enc.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    t1 = time.perf_counter()
    ms = shard.max_over_ranks(e0.elapsed_time(e1))
    launches = enc.launch_count() - l0
    kms, kn = enc.kernel_ms()
    clocks = sampler.window(t0, t1) if sampler else None
    nframes = S * F
    value = nframes * FRAME_SECONDS * world * args.steps / (ms / 1e3)
    e2e = None
    if not args.no_e2e:
        try:
            import psutil
            avail = psutil.virtual_memory().available
        except ImportError:
            avail = 64 << 30
        s2 = S
        while s2 > 64 and s2 * F * ENC_ALGO_BYTES_PER_FRAME > avail // (3 * max(world, 1)):
            s2 //= 2
        pcm_h = pcm[:s2].cpu().pin_memory()
        out_h = torch.zeros((s2, F, FRAME_BYTES), dtype=torch.uint8).pin_memory()
        st_h = torch.zeros((s2, F), dtype=torch.int32).pin_memory()

        def hstep():
            rc = enc.L.ac3_batch_encode(enc.ctx, pcm_h.data_ptr(), s2, F, 48000, 448000, 6, None, out_h.data_ptr(),
                                        st_h.data_ptr(), None, None, 0, None)
            assert rc == 0
        hstep()
        barrier()
        n = max(1, min(args.steps, args.e2e_steps))
        tt = time.perf_counter()
        for _ in range(n):
            hstep()
        barrier()
        wall = shard.max_over_ranks(time.perf_counter() - tt)
        e2e = {"value": s2 * F * FRAME_SECONDS * world * n / wall, "unit": UNIT,
               "h2d_bytes_per_step": int(pcm_h.numel() * 2), "d2h_bytes_per_step": int(out_h.numel() + st_h.numel() * 4),
               "streams_per_gpu": s2, "steps": n, "ms_per_step": 1e3 * wall / n,
               "api": "ac3_batch_encode (host pointers, pinned)"}
    if sampler:
        sampler.stop()
    if rank == 0:
        peak, peak_src = 6650.0, "fallback"
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            if peaks.get("hbm_gbs"):
                peak, peak_src = float(peaks["hbm_gbs"]), "measured"
        except (OSError, ValueError):
            pass
        achieved = nframes * ENC_ALGO_BYTES_PER_FRAME / (kms / 1e3) / 1e9 if kms > 0 else 0.0
        print(json.dumps({
            "metric": METRIC_ENC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int16/int32 fixed point", "data": "synthetic", "config": workload_config(args),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src, "kernel": "ac3_encode_kernel", "kernel_ms": kms,
                         "kernel_launches_timed": kn, "algorithmic_bytes_per_launch": nframes * ENC_ALGO_BYTES_PER_FRAME},
            "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}))
    enc.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_gpu_mixed(args):
    """Config 5 (BASELINE.json configs[4]): the mixed corpus - every feasible (sample rate, channels, bitrate) cell of
    the reference encoder, `--replicas` streams of 10 s per cell and GPU - encoded and decoded back on the GPUs, the
    streams dealt to the ranks by shard.partition_streams (longest-processing-time-first on frames x channels).  No
    collective on the data path; afterwards the ranks exchange per-cell checksums: every replica of a cell, whichever
    rank decoded it, must give the same PCM bits."""
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    eng = ge.load_engine()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - there is no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import importlib
    shard = importlib.import_module("ac3_acm_codec_b200.shard")
    corpus = _corpus_mod()
    dev = torch.device("cuda", local)
    cells = corpus.mixed_cells()
    R = args.replicas
    # the job: world x R replicas of every cell; stream g = replica g // ncells of cell g % ncells
    ncell = len(cells)
    nstream = world * R * ncell
    secs = args.frames * FRAME_SECONDS                                  # nominal stream length (10 s)
    frames_of = lambda fs: int(round(secs * fs / 1536.0))
    cost = np.array([frames_of(cells[g % ncell][0]) * cells[g % ncell][1] for g in range(nstream)], np.int64)
    parts = shard.partition_streams(cost, world)
    mine = parts[rank]
    load = np.array([cost[p].sum() for p in parts], np.float64)
    # my streams grouped by cell (one encoder call per cell: ac3_batch_encode takes one format per call)
    by_cell = {}
    for g in mine:
        by_cell.setdefault(int(g) % ncell, []).append(int(g))
    # ac3_batch_encode takes one format per call and a context serves one call at a time: the cells go round a
    # set of contexts, each on its own CUDA stream, so that the small per-cell launches share the GPU
    NCTX = 16
    encs = [eng.BatchEncoder(local) for _ in range(NCTX)]
    enc = encs[0]
    dec = eng.BatchDecoder(local)
    side = [torch.cuda.Stream(device=dev) for _ in range(NCTX)]
    stream = torch.cuda.current_stream().cuda_stream
    pcm_in, groups = {}, []
    total_frames, total_audio = 0, 0.0
    for c, gs in sorted(by_cell.items()):
        fs, nch, br = cells[c]
        nf = frames_of(fs)
        # every replica of a cell carries the same samples (seed = cell): the cross-rank checksum below relies on it
        one = corpus.synth_torch_cfg([c], nch, fs, nf * 1536, dev)
        pcm_in[c] = one.expand(len(gs), -1, -1).contiguous()
        fb = enc.frame_bytes(fs, br, nch)
        assert fb > 0, (fs, nch, br)
        groups.append((c, gs, fs, nch, br, nf, fb))
        total_frames += nf * len(gs)
        total_audio += nf * len(gs) * 1536.0 / fs
    # one elementary-stream buffer for the rank: frames of a group back to back, 16-byte aligned group starts
    pos, layout = 0, []
    for (c, gs, fs, nch, br, nf, fb) in groups:
        layout.append(pos)
        pos += (len(gs) * nf * fb + 15) & ~15
    es = torch.zeros(pos + 64, dtype=torch.uint8, device=dev)
    off_l, first_l = [], [0]
    for (c, gs, fs, nch, br, nf, fb), p0 in zip(groups, layout):
        o = p0 + np.arange(len(gs) * nf, dtype=np.int64) * fb
        off_l.append(o)
        for _ in gs:
            first_l.append(first_l[-1] + nf)
    off = torch.from_numpy(np.concatenate(off_l + [np.array([pos], np.int64)])).to(dev)
    first = torch.tensor(first_l, dtype=torch.int32, device=dev)
    nstr = len(first_l) - 1
    pcm = torch.empty(total_frames * 3072, dtype=torch.float32, device=dev)
    status = torch.zeros(total_frames, dtype=torch.int32, device=dev)
    est = torch.zeros(total_frames, dtype=torch.int32, device=dev)
    dec.set_max_frame_bytes(3840)
    dec.set_max_stream_frames(max(g[5] for g in groups))

    def encode_all():
        fo = 0
        cur = torch.cuda.current_stream()
        for sd in side:
            sd.wait_stream(cur)
        for k, ((c, gs, fs, nch, br, nf, fb), p0) in enumerate(zip(groups, layout)):
            encs[k % NCTX].encode_device(pcm_in[c].data_ptr(), len(gs), nf, fs, br, nch, es.data_ptr() + p0,
                                         status_ptr=est.data_ptr() + 4 * fo, stream=side[k % NCTX].cuda_stream)
            fo += len(gs) * nf
        for sd in side:
            cur.wait_stream(sd)

    def decode_all():
        dec.decode_device(es.data_ptr(), pos, off.data_ptr(), total_frames, first.data_ptr(), nstr, REQ_FLAGS, pcm.data_ptr(),
                          status_ptr=status.data_ptr(), out_fmt=eng.PCM_F32_INTERLEAVED, stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(1, args.warmup)):
        encode_all(); decode_all()
    barrier()
    if int((est != 0).sum().item()) or int((status != 0).sum().item()):
        fo, bad = 0, []
        for (c, gs, fs, nch, br, nf, fb) in groups:
            e = int((est[fo:fo + nf * len(gs)] != 0).sum().item()); d = int((status[fo:fo + nf * len(gs)] != 0).sum().item())
            if e or d:
                bad.append((cells[c], e, d))
            fo += nf * len(gs)
        raise SystemExit("mixed corpus: cells with encode / decode errors: %r" % (bad[:40],))
    sampler = ClockSampler(local) if rank == 0 else None
    time.sleep(0.25)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    nl = lambda: sum(e.launch_count() for e in encs) + dec.launch_count()
    l0 = nl()
    enc_ms = dec_ms = 0.0
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ev[0].record(); encode_all(); ev[1].record(); decode_all(); ev[2].record()
        torch.cuda.synchronize()
        enc_ms += ev[0].elapsed_time(ev[1]); dec_ms += ev[1].elapsed_time(ev[2])
    barrier()
    t1 = time.perf_counter()
    launches = nl() - l0
    enc_ms = shard.max_over_ranks(enc_ms) / args.steps
    dec_ms = shard.max_over_ranks(dec_ms) / args.steps
    # round-trip sanity on this rank: signal-to-noise of the stereo downmix is not defined for every mode, so the check
    # is energy: decoded energy within a factor of the input's (a silent or exploding decode fails), every frame OK
    e_out = float((pcm.double() ** 2).sum().item())
    assert e_out > 0
    # per-cell checksum: the PCM bits of the first replica this rank holds of each cell
    chk = np.zeros(ncell, np.float64)
    fo = 0
    for (c, gs, fs, nch, br, nf, fb) in groups:
        seg = pcm[fo * 3072:(fo + nf) * 3072].view(torch.int32).to(torch.int64)
        chk[c] = float((seg & 0xFFFFF).sum().item() % 1000003) + 1.0
        # all replicas on this rank agree bit for bit
        allr = pcm[fo * 3072:(fo + nf * len(gs)) * 3072].view(torch.int32).view(len(gs), -1)
        assert bool((allr == allr[0:1]).all()), cells[c]
        fo += nf * len(gs)
    allchk = shard.gather_records(chk)                                   # [world][ncell], 0 = rank holds no replica
    agree = True
    for c in range(ncell):
        v = allchk[:, c][allchk[:, c] > 0]
        agree = agree and len(v) > 0 and bool((v == v[0]).all())
    audio = shard.gather_records([total_audio, float(total_frames), float(nstr)])
    clocks = sampler.window(t0, t1) if sampler else None
    if sampler:
        sampler.stop()
    if rank == 0:
        tot_audio = float(audio[:, 0].sum())
        ms = enc_ms + dec_ms
        print(json.dumps({
            "metric": "round-trip (encode + decode) audio-sec/sec, mixed corpus", "value": tot_audio / (ms / 1e3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int16/int32 fixed point (encode), f32 (decode)", "data": "synthetic",
            "config": {"workload": "mixed corpus (BASELINE.json configs[4]): %d cells (32 / 44.1 / 48 kHz x 1..6 channels x "
                                   "feasible bitrates 40..640 kb/s) x %d replicas per GPU, %.1f s streams, encode then decode "
                                   "to stereo float" % (ncell, R, secs),
                       "streams_total": int(nstream), "frames_total": int(audio[:, 1].sum()), "audio_seconds_total": tot_audio,
                       "cache": "PCM in + frames + PCM out per rank exceed the 126 MB L2"},
            "encode_ms": enc_ms, "decode_ms": dec_ms,
            "encode_audio_s_per_s": tot_audio / (enc_ms / 1e3), "decode_audio_s_per_s": tot_audio / (dec_ms / 1e3),
            "shards": {"partition": "shard.partition_streams: longest-processing-time-first on frames x channels",
                       "streams_per_rank": [int(x) for x in audio[:, 2]], "load_imbalance_max_over_mean": float(load.max() / load.mean()),
                       "replica_checksums_agree_across_ranks": bool(agree)},
            "gpu_launches": int(launches), "clocks": clocks}))
    for e in encs:
        e.close()
    dec.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def _timed(step, steps, warmup, barrier, shard):
    import torch
    for _ in range(warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    return shard.max_over_ranks(e0.elapsed_time(e1)) / steps


def run_extras(args, eng, dec, es, es_bytes, off, first, status, shard, barrier, world, dev):
    """Supplementary figures, device resident, same timing discipline (CUDA events, max over ranks, inputs and
    outputs larger than L2), each against the HBM roofline with its own algorithmic bytes."""
    import torch
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except (OSError, ValueError, KeyError):
        pass
    S, F = args.streams, args.frames
    nframes = S * F
    stream = torch.cuda.current_stream().cuda_stream
    n, w = args.extra_steps, 2
    out = {}

    def fig(ms, frames, algo_bytes, note):
        return {"value": frames * FRAME_SECONDS * world / (ms / 1e3), "unit": UNIT, "ms_per_step": ms,
                "roofline_frac": frames * algo_bytes / (ms / 1e3) / 1e9 / peak, "algorithmic_bytes_per_frame": algo_bytes,
                "workload": note}

    # (1) the headline corpus with int16 stereo out (what a52dec -o wav and the ACM wrapper deliver)
    pcm16 = torch.empty(nframes * 1536 * 2, dtype=torch.int16, device=dev)
    ms = _timed(lambda: dec.decode_device(es.data_ptr(), es_bytes, off.data_ptr(), nframes, first.data_ptr(), S, REQ_FLAGS,
                                          pcm16.data_ptr(), status_ptr=status.data_ptr(), out_fmt=eng.PCM_S16_INTERLEAVED,
                                          stream=stream), n, w, barrier, shard)
    assert int((status != 0).sum().item()) == 0
    out["pcm_s16"] = fig(ms, nframes, FRAME_BYTES + PCM_BYTES // 2, "configs[1] corpus, int16 stereo interleaved out")
    del pcm16

    # (2) config 3: 5.1 @ 640 kb/s, block switching / coupling / dynrng / delta bit allocation in most blocks
    fx = np.load(os.path.join(ROOT, "tests", "golden", "c3_fixture.npz"))
    c3 = torch.from_numpy(fx["frames"]).to(dev)                                  # [8][32][2560]
    nu, nf3, fb3 = c3.shape
    S3 = S
    F3 = 96                                                                      # 3.07 s per stream: 3 passes over a unique stream
    base = torch.arange(S3, device=dev) % nu
    idx = (torch.arange(F3, device=dev)[None, :] + 5 * (torch.arange(S3, device=dev) // nu)[:, None]) % nf3
    n3 = S3 * F3
    es3 = torch.zeros(n3 * fb3 + 64, dtype=torch.uint8, device=dev)
    es3[:n3 * fb3].copy_(c3[base[:, None], idx].reshape(-1))
    off3 = torch.arange(n3 + 1, dtype=torch.int64, device=dev) * fb3
    first3 = (torch.arange(S3 + 1, dtype=torch.int64, device=dev) * F3).to(torch.int32)
    st3 = torch.zeros(n3, dtype=torch.int32, device=dev)
    dec.set_max_frame_bytes(fb3)
    dec.set_max_stream_frames(F3)
    for key, flags, nout, ekey in (("config3_stereo", REQ_FLAGS, 2, "energy_stereo"), ("config3_51", 7 | 16, 6, "energy_51")):
        pcm3 = torch.empty(n3 * 1536 * nout, dtype=torch.float32, device=dev)
        ms = _timed(lambda: dec.decode_device(es3.data_ptr(), n3 * fb3, off3.data_ptr(), n3, first3.data_ptr(), S3, flags,
                                              pcm3.data_ptr(), status_ptr=st3.data_ptr(), out_fmt=eng.PCM_F32_INTERLEAVED,
                                              stream=stream), n, w, barrier, shard)
        assert int((st3 != 0).sum().item()) == 0
        # stream 0 is unique stream 0 from its first frame: its first 32 frames against the reference decoder's energy
        got = float((pcm3[: nf3 * 1536 * nout].double() ** 2).sum().item())
        want = float(fx[ekey][0])
        assert abs(got - want) / want < 1e-5, ("config-3 energy differs from the reference digest", key, got, want)
        out[key] = fig(ms, n3, fb3 + 1536 * nout * 4,
                       "configs[2]: %d streams x %d frames of 5.1 640 kb/s (8 unique x 32 frames of tests/golden/c3_fixture.npz), "
                       "float %s out" % (S3, F3, "stereo" if nout == 2 else "5.1 + LFE"))
        del pcm3
    dec.set_max_frame_bytes(FRAME_BYTES)
    dec.set_max_stream_frames(F)
    del es3

    # (3) config 4: batched encode of the corpus PCM (the first 256 streams' samples, tiled)
    corpus = _corpus_mod()
    Fe = 64
    pcm_u = corpus.synth_torch(range(64), dev, Fe * 1536).reshape(64, Fe, 1536 * 6)
    pcm_e = pcm_u[torch.arange(S, device=dev) % 64].contiguous()                   # [S][Fe][1536*6] int16
    out_e = torch.zeros((S, Fe, FRAME_BYTES), dtype=torch.uint8, device=dev)
    st_e = torch.zeros((S, Fe), dtype=torch.int32, device=dev)
    enc = eng.BatchEncoder(dev.index or 0)
    ms = _timed(lambda: enc.encode_device(pcm_e.data_ptr(), S, Fe, 48000, 448000, 6, out_e.data_ptr(),
                                          status_ptr=st_e.data_ptr(), stream=stream), n, w, barrier, shard)
    assert int((st_e != 0).sum().item()) == 0
    enc.close()
    out["config4_encode"] = fig(ms, S * Fe, ENC_ALGO_BYTES_PER_FRAME,
                                "configs[3]: %d streams x %d frames of 5.1 int16 PCM -> 448 kb/s (64 unique)" % (S, Fe))
    out["config4_encode"]["metric"] = METRIC_ENC
    return out


def run_e2e(args, eng, dec, corpus, shard, barrier, world):
    import torch
    S, F = args.streams, args.frames
    # bound the pinned allocation by what the host can spare
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except ImportError:
        avail = 64 << 30
    per_stream = F * (FRAME_BYTES + PCM_BYTES)
    s_e2e = S
    budget = avail // (3 * max(world, 1))
    while s_e2e > 64 and s_e2e * per_stream > budget:
        s_e2e //= 2
    nframes = s_e2e * F
    es_h = corpus[:nframes * FRAME_BYTES].cpu().pin_memory()
    s16 = args.pcm == "s16"
    pcm_h = torch.empty(nframes * 1536 * 2, dtype=torch.int16 if s16 else torch.float32).pin_memory()
    status_h = torch.zeros(nframes, dtype=torch.int32).pin_memory()
    off_h = (np.arange(nframes, dtype=np.uint64) * FRAME_BYTES)
    first_h = (np.arange(s_e2e + 1, dtype=np.uint32) * F).astype(np.uint32)

    def step():
        dec.decode_host_into(es_h.data_ptr(), es_h.numel(), off_h, first_h, REQ_FLAGS, pcm_h.data_ptr(),
                             status_h.data_ptr(), out_fmt=eng.PCM_S16_INTERLEAVED if s16 else eng.PCM_F32_INTERLEAVED)

    for _ in range(max(1, min(args.warmup, 2))):
        step()
    barrier()
    n = max(1, min(args.steps, args.e2e_steps))
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    barrier()
    wall = time.perf_counter() - t0
    wall = shard.max_over_ranks(wall)
    assert int((status_h != 0).sum().item()) == 0
    return {"value": nframes * FRAME_SECONDS * world * n / wall, "unit": UNIT,
            "h2d_bytes_per_step": int(es_h.numel() + off_h.nbytes + first_h.nbytes),
            "d2h_bytes_per_step": int(pcm_h.numel() * pcm_h.element_size() + status_h.numel() * 4),
            "streams_per_gpu": s_e2e, "steps": n, "ms_per_step": 1e3 * wall / n,
            "api": "a52_batch_decode (host pointers, pinned)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=4096)
    ap.add_argument("--frames", type=int, default=313)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--pcm", default="f32", choices=["f32", "s16"],
                    help="decoded sample format: f32 = the headline configuration (float stereo, 14 080 algorithmic "
                         "bytes per frame); s16 = what the ACM wrapper and `a52dec -o wav` deliver (int16 stereo, "
                         "7 936 bytes per frame) - a supplementary line, never the default")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the supplementary figures (int16, config 3, config 4)")
    ap.add_argument("--extra-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--workload", default="decode", choices=["decode", "encode", "mixed"],
                    help="decode = BASELINE.json configs[1] (the headline metric); encode = configs[3]; mixed = configs[4] "
                         "(encode-then-decode round trip of the mixed corpus, sharded over the GPUs)")
    ap.add_argument("--replicas", type=int, default=8, help="mixed workload: streams per cell and GPU")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "encode":
        return run_gpu_encode(args)
    if args.workload == "mixed":
        return run_gpu_mixed(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
