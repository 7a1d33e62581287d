"""Developer check / profiling target: the config-3 corpus alone (tests/golden/c3_fixture.npz tiled), device resident.
usage: python tools/dev_c3.py [streams] [frames] [stereo|51] [steps]"""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as g
eng = g.load_engine(); dec = eng.BatchDecoder(0)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1776
F = int(sys.argv[2]) if len(sys.argv) > 2 else 32
mode = sys.argv[3] if len(sys.argv) > 3 else "stereo"
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
fx = np.load(os.path.join(ROOT, "tests", "golden", "c3_fixture.npz"))
c3 = torch.from_numpy(fx["frames"]).cuda(); nu, nf3, fb = c3.shape
base = torch.arange(S, device="cuda") % nu
idx = (torch.arange(F, device="cuda")[None, :] + 5 * (torch.arange(S, device="cuda") // nu)[:, None]) % nf3
n = S * F
es = torch.zeros(n * fb + 64, dtype=torch.uint8, device="cuda"); es[:n * fb].copy_(c3[base[:, None], idx].reshape(-1))
off = torch.arange(n + 1, dtype=torch.int64, device="cuda") * fb
first = (torch.arange(S + 1, dtype=torch.int64, device="cuda") * F).to(torch.int32)
flags, nout = ((2 | 32), 2) if mode == "stereo" else ((7 | 16), 6)
pcm = torch.empty(n * 1536 * nout, dtype=torch.float32, device="cuda"); st = torch.zeros(n, dtype=torch.int32, device="cuda")
dec.set_max_frame_bytes(fb); dec.set_max_stream_frames(F)
def step():
    dec.decode_device(es.data_ptr(), n * fb, off.data_ptr(), n, first.data_ptr(), S, flags, pcm.data_ptr(),
                      status_ptr=st.data_ptr(), out_fmt=eng.PCM_F32_INTERLEAVED)
for _ in range(3): step()
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(steps): step()
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / steps
assert int((st != 0).sum()) == 0
print("config 3 (%d x %d) -> %s: %.0f audio-s/s, %.3f ms/step" % (S, F, mode, n * 0.032 / dt, dt * 1e3), flush=True)
