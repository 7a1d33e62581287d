"""The tensor-core IMDCT experiment (a52_imdct_ab.inl): FFT transform vs 3xTF32 mma.sync GEMM, same planes, same result.
Builds the 256 x 256 transform matrix from a float64 restatement of imdct.c:258-345 (U / V layout of the kernel),
checks both variants against it, times them with CUDA events and prints / stores one JSON record.
usage: python tools/dev_imdct_ab.py [nplanes] [--once variant]   (--once: a single launch, for ncu)"""
import ctypes as C, json, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as g


def imdct_uv(x):
    """float64: 256 coefficients -> (U[128], V[128]) as imdct512_warp leaves them (natural-order formulation)."""
    m = np.arange(128)
    th = (np.pi / 256) * (m + 64 - 0.25); sg = np.where(m & 1, -1.0, 1.0)
    pre = sg * np.cos(th) + 1j * sg * np.sin(th)
    a, b = x[2 * m], x[255 - 2 * m]
    z = (pre.real * a + pre.imag * b) + 1j * (pre.real * b - pre.imag * a)
    B = np.fft.fft(z)                                   # forward DFT e^{-2 pi j mk / 128}
    i = np.arange(64); p = np.cos((np.pi / 256) * (i + 0.5)) + 1j * np.sin((np.pi / 256) * (i + 0.5))
    B1, B2 = B[i], B[127 - i]
    U = np.zeros(128); V = np.zeros(128)
    U[2 * i] = p.real * B1.real + p.imag * B1.imag
    V[2 * i] = p.imag * B1.real - p.real * B1.imag
    U[2 * i + 1] = -(p.imag * B2.real + p.real * B2.imag)
    V[2 * i + 1] = p.real * B2.real - p.imag * B2.imag
    return np.concatenate([U, V])


def tf32_rna(a):
    b = a.astype(np.float32).view(np.uint32).astype(np.uint64)
    return ((b + 0x1000) & 0xFFFFE000).astype(np.uint32)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else 1 << 20
    once = int(sys.argv[sys.argv.index("--once") + 1]) if "--once" in sys.argv else None
    eng = g.load_engine(); L = eng.load_library()
    L.a52_ab_imdct.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    dec = eng.BatchDecoder(0)
    M = np.stack([imdct_uv(np.eye(256)[k]) for k in range(256)], axis=1)          # [256 out][256 in]
    hi = tf32_rna(M); lo = tf32_rna(M - hi.view(np.float32).astype(np.float64))
    fr = np.zeros((16, 32, 32, 8), np.uint32)
    lane = np.arange(32); gq, t = lane >> 2, lane & 3
    for mt in range(16):
        for ks in range(32):
            r, k = 16 * mt + gq, 8 * ks + t
            for j, (rr, kk) in enumerate([(r, k), (r + 8, k), (r, k + 4), (r + 8, k + 4)]):
                fr[mt, ks, :, j] = hi[rr, kk]; fr[mt, ks, :, 4 + j] = lo[rr, kk]
    afrag = torch.from_numpy(fr.view(np.int32)).cuda()
    gen = torch.Generator(device="cuda").manual_seed(1)
    x = (torch.randn((n, 256), device="cuda", generator=gen) * torch.rand((n, 1), device="cuda", generator=gen)).contiguous()
    y = [torch.zeros((n, 256), device="cuda") for _ in range(2)]
    st = torch.cuda.current_stream().cuda_stream

    def run(v):
        rc = L.a52_ab_imdct(dec.ctx, v, x.data_ptr(), y[v].data_ptr(), n, afrag.data_ptr(), st)
        assert rc == 0
    if once is not None:
        run(once); torch.cuda.synchronize(); return
    rec = {"planes": n, "what": "IMDCT-512 of float planes, global to global (imdct.c:258-345), one B200"}
    ref = (M @ x[:2048].double().cpu().numpy().T).T
    for v, name in ((0, "fft_f32x2"), (1, "mma_3xtf32")):
        for _ in range(3): run(v)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): run(v)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        d = y[v][:2048].double().cpu().numpy() - ref
        rec[name] = {"ms": ms, "planes_per_s": n / (ms / 1e3), "GB_per_s": n * 2048 / (ms / 1e3) / 1e9,
                     "rel_rms_error_vs_float64": float(np.sqrt((d * d).mean()) / np.sqrt((ref * ref).mean()))}
    rec["mma_3xtf32"]["useful_TFLOP_per_s"] = n * 2 * 256 * 256 * 3 / (rec["mma_3xtf32"]["ms"] / 1e3) / 1e12
    rec["fft_over_mma"] = rec["mma_3xtf32"]["ms"] / rec["fft_f32x2"]["ms"]
    print(json.dumps(rec))
    json.dump(rec, open(os.path.join(ROOT, "gpurun_out", "imdct_tc_ab.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
