"""N>1 host logic on CPU: world_size-2 (and 3) gloo runs of the time-sharding plan + halo ring shift, checked
against the oracle (sharded FIR / resampler with the exchanged halo as history == the unsharded run), plus the
channel partition. No product compute runs here (that needs a GPU); the oracle stands in as the per-shard filter."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from qdsp_b200 import shard, synth


def test_channel_slice_is_a_partition():
    for nch, world in [(256, 8), (256, 3), (5, 8), (32, 1)]:
        seen = []
        for r in range(world):
            s = shard.channel_slice(nch, world, r)
            seen += list(range(s.start, s.stop))
        assert seen == list(range(nch))
        sizes = [shard.channel_slice(nch, world, r).stop - shard.channel_slice(nch, world, r).start for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


def test_time_shards_cover_stream_on_block_grid():
    for total, world, block, hist in [(1 << 20, 8, 65536, 4094), (1000003, 4, 81920, 401), (5000, 8, 4096, 126), (0, 2, 100, 10)]:
        sh = shard.time_shards(total, world, block, hist)
        assert sum(s.count for s in sh) == total
        pos = 0
        for s in sh:
            assert s.start == pos or s.count == 0
            assert s.start % block == 0 or s.count == 0
            pos += s.count
            need = shard.halo_sources(sh, s.rank)
            assert sum(n for _, _, n in need) == s.halo
        assert sh[0].halo == 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, block, T, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import loader

        P = loader.port()
        taps = (synth.uniform_f32(99, 0, T) / T).astype(np.float32)
        sh = shard.time_shards(total, world, block, T - 1)
        me = sh[rank]
        local = torch.from_numpy(synth.uniform_cf32(3, me.start, me.count))
        halo = shard.exchange_halo(sh, rank, local, T - 1, dist)
        # per-shard FIR with the exchanged halo as history (what qdsp_fir_import_tail + process do on the GPU)
        x = np.concatenate([halo.numpy(), local.numpy()])
        y = P.fir_cf32(taps, x)[T - 1:]
        q.put((rank, me.start, y))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,total,block,T", [(2, 40000, 4096, 127), (3, 30000, 1000, 4095)])
def test_time_sharded_fir_equals_unsharded(world, total, block, T):
    from oracle import loader

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, block, T, q)) for r in range(world)]
    for p in procs:
        p.start()
    parts = sorted([q.get(timeout=120) for _ in range(world)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    y = np.concatenate([p[2] for p in parts])
    taps = (synth.uniform_f32(99, 0, T) / T).astype(np.float32)
    ref = loader.port().fir_cf32(taps, synth.uniform_cf32(3, 0, total))
    assert y.shape == ref.shape
    assert np.array_equal(y.view(np.uint32), ref.view(np.uint32)), "time-sharded FIR must be bit-identical to the unsharded run"
