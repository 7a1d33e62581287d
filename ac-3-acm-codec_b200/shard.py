"""Multi-GPU sharding of a batch of streams: one process per GPU, no collective on the data path.

Streams are independent units, so a rank decodes its share with the same
single-GPU call and only tiny per-rank records (frames, audio seconds, checksum, elapsed ms) are
exchanged afterwards through torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np


def partition_streams(costs, world_size):
    """Assign streams to ranks.  `costs[i]` = cost of stream i (e.g. frames x coded bins).

    Equal costs: contiguous ranges [g*S/G, (g+1)*S/G) so a rank's bitstream is one slice.
    Unequal costs: longest-processing-time-first greedy.  Returns a list of index arrays
    (ascending inside a rank); every stream appears exactly once."""
    costs = np.asarray(costs, dtype=np.int64)
    n = len(costs)
    if world_size <= 1:
        return [np.arange(n, dtype=np.int64)]
    if n == 0 or (costs == costs[0]).all():
        edges = [(g * n) // world_size for g in range(world_size + 1)]
        return [np.arange(edges[g], edges[g + 1], dtype=np.int64) for g in range(world_size)]
    order = np.argsort(-costs, kind="stable")
    load = np.zeros(world_size, np.int64)
    bins = [[] for _ in range(world_size)]
    for i in order:
        g = int(np.argmin(load))
        bins[g].append(int(i))
        load[g] += costs[i]
    return [np.array(sorted(b), dtype=np.int64) for b in bins]


def shard_batch(es, frame_off, frame_len, stream_first, streams):
    """Cut the (es, frame_off, stream_first) triple down to the given streams, repacking every
    frame at a 16-byte aligned offset (what the TMA staging of the decode kernel wants)."""
    es = np.asarray(es, np.uint8)
    frame_off = np.asarray(frame_off, np.uint64)
    frame_len = np.asarray(frame_len, np.int64)
    stream_first = np.asarray(stream_first, np.int64)
    idx = [np.arange(stream_first[s], stream_first[s + 1]) for s in streams]
    nfr = np.array([len(i) for i in idx], np.int64)
    fi = np.concatenate(idx) if idx else np.zeros(0, np.int64)
    ln = frame_len[fi]
    padded = (ln + 15) & ~15
    new_off = np.concatenate([[0], np.cumsum(padded)]).astype(np.uint64)
    out = np.zeros(int(new_off[-1]) + 32, np.uint8)
    for k, f in enumerate(fi):
        out[int(new_off[k]):int(new_off[k]) + int(ln[k])] = es[int(frame_off[f]):int(frame_off[f]) + int(ln[k])]
    first = np.concatenate([[0], np.cumsum(nfr)]).astype(np.uint32)
    return out, new_off[:-1].copy(), first


def gather_records(record, dist=None):
    """All-gather one small float64 record per rank; returns array [world, len(record)]."""
    import torch
    if dist is None:
        import torch.distributed as dist
    rec = torch.tensor(np.asarray(record, np.float64))
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return rec.numpy()[None, :]
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    rec = rec.to(dev)
    out = [torch.zeros_like(rec) for _ in range(dist.get_world_size())]
    dist.all_gather(out, rec)
    return torch.stack(out).cpu().numpy()


def max_over_ranks(value, dist=None):
    import torch
    if dist is None:
        import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
