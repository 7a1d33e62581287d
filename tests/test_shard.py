"""Host-side multi-GPU logic on CPU: partitioning, batch cutting, and a world_size-2 gloo run of
the record exchange bench.py uses (the data path itself has no collective)."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _shard_mod(engine):
    import importlib
    return importlib.import_module("ac3_acm_codec_b200.shard")


def test_partition_equal_costs_contiguous(engine):
    sh = _shard_mod(engine)
    for n in (0, 1, 7, 4096):
        for g in (1, 2, 4, 8):
            parts = sh.partition_streams(np.full(n, 313), g)
            assert len(parts) == g
            allidx = np.concatenate(parts) if n else np.zeros(0, np.int64)
            assert (allidx == np.arange(n)).all()
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= 1


def test_partition_lpt_balanced(engine):
    sh = _shard_mod(engine)
    rng = np.random.RandomState(0)
    costs = rng.randint(1, 1000, 300)
    for g in (2, 4, 8):
        parts = sh.partition_streams(costs, g)
        assert sorted(np.concatenate(parts).tolist()) == list(range(300))
        loads = np.array([costs[p].sum() for p in parts])
        assert loads.max() - loads.min() <= costs.max()


def test_shard_batch_repack(engine, c2):
    sh = _shard_mod(engine)
    frames = c2["frames"][:, :3]
    es = frames.reshape(-1)
    # ragged: stream lengths 3, 0, 2, 1 frames
    counts = [3, 0, 2, 1]
    off, first = [], [0]
    for s, n in enumerate(counts):
        for f in range(n):
            off.append((s * 3 + f) * 1792)
        first.append(first[-1] + n)
    off = np.array(off, np.uint64)
    ln = np.full(len(off), 1792)
    sub, soff, sfirst = sh.shard_batch(es, off, ln, first, [2, 3])
    assert sfirst.tolist() == [0, 2, 3] and (soff % 16 == 0).all()
    for k, (s, f) in enumerate([(2, 0), (2, 1), (3, 0)]):
        assert (sub[int(soff[k]):int(soff[k]) + 1792] == frames[s, f]).all()


WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import __graft_entry__ as ge
eng = ge.load_engine()
import importlib
sh = importlib.import_module("ac3_acm_codec_b200.shard")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.RandomState(1)
nstreams = 37
data = [rng.randint(0, 256, rng.randint(1, 9) * 100).astype(np.uint8) for _ in range(nstreams)]
costs = [len(d) for d in data]
mine = sh.partition_streams(costs, world)[rank]
chk = float(sum(int(data[i].astype(np.int64).sum()) * (i + 1) for i in mine))
rec = sh.gather_records([len(mine), sum(costs[i] for i in mine), chk, 10.0 + rank])
tmax = sh.max_over_ranks(10.0 + rank)
if rank == 0:
    total = float(sum(int(d.astype(np.int64).sum()) * (i + 1) for i, d in enumerate(data)))
    assert rec.shape == (world, 4)
    assert rec[:, 0].sum() == nstreams and rec[:, 2].sum() == total, (rec, total)
    assert tmax == 10.0 + world - 1
    print("SHARD_OK", world)
dist.destroy_process_group()
'''


def test_gloo_world2_record_exchange(engine, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29631", str(script), ROOT],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "SHARD_OK 2" in out.stdout
