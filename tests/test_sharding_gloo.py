"""N > 1 path on CPU: two gloo ranks shard a set of streams round-robin (stream g -> rank g mod N), each decodes its
share with no data-path collective, and the union equals the unsharded result.  The decoder here is the oracle (the
GPU is not available on this box): what is under test is the sharding, the corpus addressing by GLOBAL stream id and
the barrier / max-over-ranks plumbing that bench.py uses."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_global, n_frames, out):
    import torch.distributed as dist

    import __graft_entry__ as ge
    from oracle import ref

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = ge.load_package()
    ids = pkg.shard.shard_streams(n_global, rank, world)
    assert all(pkg.shard.owner(int(g), world) == rank for g in ids)
    assert [pkg.shard.local_index(int(g), world) for g in ids] == list(range(len(ids)))
    data, frames = pkg.shard.rank_corpus(pkg.api, rank, world, len(ids), n_frames, 159, 0xFC)
    # "decode" this rank's share: packets are rebuilt from the rank-local corpus, i.e. what the GPU path would be fed
    hashes = np.zeros(n_global, np.uint64)
    dec = ref.PacketDecoder(2)
    for s, g in enumerate(ids):
        ref.lib().ref_opus_init(dec.st, 2)
        pcm = []
        for f in range(n_frames):
            fr = frames[s, f]
            pkt = np.concatenate([np.array([fr["toc"]], np.uint8), data[fr["offset"]: fr["offset"] + fr["len"]]])
            n, p = dec.decode(pkt)
            assert n == 960
            pcm.append(p)
        hashes[g] = ref.fnv_hash(np.concatenate(pcm), 2)
    dec.close()
    dist.barrier()
    import torch

    t = torch.from_numpy(hashes.view(np.int64).copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)  # verification only: disjoint shards, so the sum is the union
    # the gather bench.py uses for its cross-rank result check: per-rank hashes -> one array by global stream id
    by_g = pkg.shard.gather_by_global_id(hashes[ids], rank, world, dist)
    assert np.array_equal(by_g, t.numpy().view(np.uint64))
    slow = pkg.shard.max_over_ranks(1.0 + rank, dist)
    rate = pkg.shard.aggregate_rate(len(ids) * n_frames * 0.02, world, slow)
    if rank == 0:
        np.save(out, np.concatenate([t.numpy().view(np.uint64), np.array([slow, rate]).view(np.uint64)]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_cover_all_streams_exactly_once(tmp_path, ref):
    import torch.multiprocessing as mp

    world, n_global, n_frames = 2, 10, 6
    out = str(tmp_path / "h.npy")
    mp.spawn(_worker, args=(world, _free_port(), n_global, n_frames, out), nprocs=world, join=True)
    got = np.load(out)
    hashes, tail = got[:n_global], got[n_global:].view(np.float64)
    for g in range(n_global):
        want, _, err = ref.decode_synth_stream(0xFC, 160, 2, n_frames, seed=ref.SEED ^ g, want_pcm=False)
        assert err == 0
        assert int(hashes[g]) == want, g
    assert tail[0] == 2.0  # max over ranks of (1 + rank)
    assert tail[1] == pytest.approx(world * (n_global // world) * n_frames * 0.02 / 2.0)


def test_shard_maps_are_a_partition(pkg):
    for world in (1, 2, 4, 8):
        seen = np.concatenate([pkg.shard.shard_streams(65536, r, world) for r in range(world)])
        assert np.array_equal(np.sort(seen), np.arange(65536))
        for r in range(world):
            ids = pkg.shard.shard_streams(1000, r, world)
            assert np.array_equal(ids, pkg.shard.global_id(np.arange(len(ids)), r, world))
