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


def test_lead_in_shards_cover_stream_on_decimation_grid():
    for per, world, D, T in [(1 << 26, 8, 1280, 10241), (100_030, 3, 50, 401), (5000, 2, 50, 401), (7, 4, 50, 401)]:
        sh = shard.lead_in_shards(per, world, D, T)
        assert sum(s.count for s in sh) == per * world and sh[0].start == 0 and sh[0].lead == 0
        for a, b in zip(sh[:-1], sh[1:]):
            assert a.start + a.count == b.start and b.start % D == 0 and b.first_out == b.start // D
            assert b.lead % D == 0 and (b.lead >= T + D or b.lead == b.start)


def _lead_worker(rank, world, port, per, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import loader

        P = loader.port()
        D, T, blk = 50, 401, 4000
        me = shard.lead_in_shards(per, world, D, T)[rank]
        g0 = me.start - me.lead
        x = synth.fm_cf32(g0, me.lead + me.count, 2_400_000, 250_000, 1000, 5e3, 0.5) + np.float32(0.01) * synth.uniform_cf32(1, g0, me.lead + me.count)
        # what the GPU rank does: seek to g0, process lead + count samples in run() blocks, drop the lead-in outputs
        a, _, _ = P.vfo_fm_window(250e3, 2.4e6, 48e3, 48e3, 5e3, x.astype(np.complex64), blk, g0)
        keep = torch.from_numpy(a[me.lead // D:].copy())
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([keep.numel()], dtype=torch.int64))
        bufs = [torch.zeros(int(s), dtype=torch.float32) for s in sizes]
        if rank == 0:
            bufs[0] = keep
            for r in range(1, world):
                dist.recv(bufs[r], src=r)
            q.put(torch.cat(bufs).numpy())
        else:
            dist.send(keep, dst=0)
    finally:
        dist.destroy_process_group()


def test_lead_in_time_shards_equal_unsharded_chain():
    """World-size-2 gloo run of the no-exchange time sharding the FFT-form channelizer uses at N > 1 (bench.py cfg4): every
    rank runs the oracle chain over lead-in + shard at its stream position and drops the lead-in outputs; the gathered
    audio equals the unsharded run."""
    from oracle import loader

    world, per = 2, 60_030
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_lead_worker, args=(r, world, port, per, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    total = world * per
    x = synth.fm_cf32(0, total, 2_400_000, 250_000, 1000, 5e3, 0.5) + np.float32(0.01) * synth.uniform_cf32(1, 0, total)
    ref, _, _ = loader.port().vfo_fm_window(250e3, 2.4e6, 48e3, 48e3, 5e3, x.astype(np.complex64), 4000, 0)
    assert got.shape == ref.shape
    assert np.abs(got[16:] - ref[16:]).max() <= 1e-6, np.abs(got[16:] - ref[16:]).max()
