"""Multi-GPU sharding of independent streams: global stream g lives on rank g mod N for its whole life,
one process per GPU, no collective on the data path.  torch.distributed is used only as the launcher's plumbing:
a barrier around the timed region and the max-over-ranks of the device time."""
import numpy as np


def owner(g, world):
    """Rank that owns global stream g."""
    return g % world


def local_index(g, world):
    """Index of global stream g inside its owner's decoder context."""
    return g // world


def shard_streams(n_global, rank, world):
    """Global ids of the streams rank `rank` owns, in local-index order (round-robin)."""
    return np.arange(rank, n_global, world, dtype=np.int64)


def global_id(local, rank, world):
    return rank + world * local


def rank_corpus(api, rank, world, n_local, n_frames, payload_len, toc, bytes_out=None, frames_out=None):
    """Synthetic packets (SURVEY 8c/8d generator, seeded by GLOBAL stream id) of the streams one rank owns.
    Returns (bytes uint8[n_local*n_frames*payload_len], frames FRAME_DTYPE[n_local][n_frames])."""
    n = n_local * n_frames * payload_len
    b = bytes_out if bytes_out is not None else np.empty(n, np.uint8)
    f = frames_out if frames_out is not None else api.make_frames(n_local, n_frames)
    per = n_frames * payload_len
    for s in range(n_local):  # global ids are interleaved across ranks: one stream at a time
        api.synth_corpus(1, n_frames, payload_len, toc, first_stream=global_id(s, rank, world), bytes_out=b[s * per:(s + 1) * per],
                         frames_out=f[s:s + 1])
    f["offset"] = (np.arange(n_local * n_frames, dtype=np.uint32) * payload_len).reshape(n_local, n_frames)
    return b, f


def max_over_ranks(value, dist=None, device="cpu"):
    """Max of a per-rank scalar (the device time of the timed region) over all ranks; identity without a process group."""
    if dist is None or not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_rate(units_per_rank, world, seconds_max):
    """Whole-job throughput: the units ALL ranks processed divided by the slowest rank's time."""
    return world * units_per_rank / seconds_max


def gather_by_global_id(local, rank, world, dist=None, device="cpu"):
    """Per-stream values of every rank (uint64[n_local] each, stream s of rank r is global stream r + world * s) as one
    array indexed by GLOBAL stream id, on every rank.  Verification plumbing (the result hashes of a sharded run), never
    on the data path."""
    local = np.ascontiguousarray(local, np.uint64)
    if dist is None or not dist.is_available() or not dist.is_initialized() or world == 1:
        return local.copy()
    import torch

    t = torch.from_numpy(local.view(np.int64).copy()).to(device)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    per_rank = torch.stack(parts).cpu().numpy().view(np.uint64).reshape(world, -1)
    return np.ascontiguousarray(per_rank.T).reshape(-1)  # [s][r] -> g = r + world * s
