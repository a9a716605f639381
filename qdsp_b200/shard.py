"""Multi-GPU partitioning of the hot path (host-side logic; one process per GPU).

Two modes (BASELINE.json north_star):

* channel partition  — the independent channels of a wideband stream are dealt out to ranks; every rank
  reads the same input, no data-path collective (`channel_slice`).
* time sharding      — one long stream is cut into contiguous shards, one per rank; the only exchange is the
  filter history: rank r needs the last H samples that precede its shard (H = tapCount-1 for FIR,
  tapsPerPhase for the resampler / fused chain), which rank r-1 owns (`time_shards`, `exchange_halo`). This is
  the reference's own history carry (`memmove(buffer, &buffer[count], H)`, filter.h:71 / resampling.h:129)
  stretched across devices. Shard boundaries are kept on the reference's run()-block grid so the resampler's
  per-block schedule restart (resampling.h:121) lands exactly where the unsharded run puts it.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


def channel_slice(nch: int, world: int, rank: int) -> slice:
    """Contiguous, balanced deal of `nch` channels over `world` ranks (first `nch % world` ranks get one more)."""
    base, extra = divmod(nch, world)
    start = rank * base + min(rank, extra)
    return slice(start, start + base + (1 if rank < extra else 0))


@dataclass(frozen=True)
class TimeShard:
    rank: int
    start: int        # first sample of the shard (absolute stream index)
    count: int        # samples in the shard
    halo: int         # history samples needed from before `start` (0 for rank 0: zero history)
    first_block: int  # index of the shard's first run() block in the global block grid
    nblocks: int


def time_shards(total: int, world: int, block: int, history: int) -> list[TimeShard]:
    """Cut `total` samples into `world` contiguous shards of whole run() blocks (`block` samples; the global
    last block may be short). Blocks are dealt as evenly as possible; a shard's halo is `history` samples
    (clipped at the stream start)."""
    if total < 0 or world < 1 or block < 1 or history < 0:
        raise ValueError("bad arguments")
    nblocks = (total + block - 1) // block
    base, extra = divmod(nblocks, world)
    shards, b0 = [], 0
    for r in range(world):
        nb = base + (1 if r < extra else 0)
        start = min(b0 * block, total)
        end = min((b0 + nb) * block, total)
        shards.append(TimeShard(r, start, end - start, min(history, start) if r > 0 else 0, b0, nb))
        b0 += nb
    return shards


@dataclass(frozen=True)
class LeadShard:
    rank: int
    start: int        # first sample whose outputs this rank keeps (absolute stream index, on the decimation grid)
    count: int        # samples in the shard
    lead: int         # samples ahead of `start` the rank also consumes; the lead // decim outputs they produce are dropped
    first_out: int    # global index of the first output the rank keeps


def lead_in_shards(per_rank: int, world: int, decim: int, taps: int, extra_rows: int = 0) -> list[LeadShard]:
    """Time shards of a decimating chain that need NO exchange: a stream of world * per_rank samples is cut on the
    decimation grid (run() blocks that are multiples of `decim` keep every output on that grid, resampling.h:121), and
    every rank but the first is handed `lead` more samples ahead of its shard: enough rows for the filter history
    (taps samples) plus one for the FM demodulator's previous phase (demodulator.h:88-92). The rank seeks its handle to
    start - lead (the NCO is closed-form in the stream position), processes lead + count samples and drops the first
    lead // decim outputs. Used by the FFT-form channelizer at N > 1 (bench.py cfg4)."""
    if per_rank < 0 or world < 1 or decim < 1 or taps < 1:
        raise ValueError("bad arguments")
    total = world * per_rank
    rows = -(-taps // decim) + 1 + extra_rows    # the resampler's history is tapsPerPhase = taps samples (resampling.h:129)
    cut = lambda r: total if r >= world else (r * per_rank // decim) * decim
    out = []
    for r in range(world):
        lo, hi = cut(r), cut(r + 1)
        lead = min(rows * decim, lo) if r > 0 else 0
        out.append(LeadShard(r, lo, hi - lo, lead, lo // decim))
    return out


def halo_sources(shards: list[TimeShard], rank: int) -> list[tuple[int, int, int]]:
    """Which ranks own rank's halo: list of (src_rank, src_offset_in_its_shard, n) in stream order. Normally one
    entry (the previous rank's tail); more only when shards are shorter than the history."""
    me = shards[rank]
    need_lo, need_hi = me.start - me.halo, me.start
    out = []
    for s in shards[:rank]:
        lo, hi = max(need_lo, s.start), min(need_hi, s.start + s.count)
        if hi > lo:
            out.append((s.rank, lo - s.start, hi - lo))
    return out


def exchange_halo(shards: list[TimeShard], rank: int, local, history: int, dist):
    """Ring-shift the history tails with point-to-point sends (gloo on CPU tensors in tests; on a GPU box this is the
    FALLBACK path -- bench.py's config 3 lets the FIR kernel read the neighbour's tail directly through a CUDA-IPC peer
    mapping, qdsp_ipc_open + qdsp_fir_process_halo, with no message at all). `local` is this rank's shard as a torch
    tensor of complex64 (any device); returns a tensor of `history` samples: zeros, then whatever precedes the shard.
    The tails travel as float32 VIEWS of the shard (no staging copies), one message per rank pair."""
    import torch

    halo = torch.zeros(history, dtype=local.dtype, device=local.device)
    halo_f = torch.view_as_real(halo)            # [history, 2] float32 view: received straight into place
    local_f = torch.view_as_real(local)
    world = len(shards)
    reqs = []
    pos = history - shards[rank].halo            # right-aligned: the newest sample sits at halo[-1]
    for src, _, n in halo_sources(shards, rank):
        reqs.append(dist.irecv(halo_f[pos:pos + n], src=src))
        pos += n
    for dst in range(rank + 1, world):
        for src, off, n in halo_sources(shards, dst):
            if src == rank:
                reqs.append(dist.isend(local_f[off:off + n], dst=dst))
    for r in reqs:
        r.wait()
    return halo
