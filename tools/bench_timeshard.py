"""tools/bench_timeshard.py — BASELINE config 3: one long cf32 stream through the 4095-tap FIR, time-sharded
across the GPUs of one box. Each rank owns a contiguous shard; the only exchange is the (taps-1)-sample history
tail, ring-shifted with NCCL point-to-point over NVLink and adopted with qdsp_fir_import_tail.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/bench_timeshard.py [--total 1073741824] [--steps 3]
Prints one JSON line (rank 0): whole-job Msamples/s (max over ranks), TFLOP/s, halo bytes, and a bit-exact check of
every shard boundary against an unsharded recomputation of the same window."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from qdsp_b200 import blocks as B, lib, shard  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--total", type=int, default=1 << 30)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--block", type=int, default=1 << 20)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    L = lib.load()
    lib.check(L.qdsp_set_device(local))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    win = B.BlackmanWindow(100e3, 4 * 2.4e6 / 4095, 2.4e6)
    T = win.getTapCount()
    shards = shard.time_shards(args.total, world, args.block, T - 1)
    me = shards[rank]
    x = torch.empty(me.count, dtype=torch.complex64, device="cuda")
    lib.check(L.qdsp_synth_uniform_cf32(x.data_ptr(), 3, me.start, me.count, sp))
    y = torch.empty(me.count, dtype=torch.complex64, device="cuda")
    fir = B.FIR(win)

    def step():
        if world > 1:
            halo = shard.exchange_halo(shards, rank, x, T - 1, dist)   # NCCL send/recv, 32 752 bytes per boundary
        else:
            halo = torch.zeros(T - 1, dtype=torch.complex64, device="cuda")
        fir.import_tail(halo.data_ptr(), -1, sp)
        m = fir.process_device(x.data_ptr(), y.data_ptr(), me.count, stream=sp)
        assert m == me.count
        return halo

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / args.steps

    # boundary check: recompute [start - 2(T-1), start + 4096) unsharded on this GPU, compare the overlap
    ok = 1
    if rank > 0 and me.count > 0:
        w0 = me.start - 2 * (T - 1)
        wn = 2 * (T - 1) + min(4096, me.count)
        xw = torch.empty(wn, dtype=torch.complex64, device="cuda")
        lib.check(L.qdsp_synth_uniform_cf32(xw.data_ptr(), 3, w0, wn, sp))
        yw = torch.empty(wn, dtype=torch.complex64, device="cuda")
        B.FIR(win).process_device(xw.data_ptr(), yw.data_ptr(), wn, stream=sp)
        torch.cuda.synchronize()
        a = torch.view_as_real(y[: wn - 2 * (T - 1)])
        b = torch.view_as_real(yw[2 * (T - 1):])
        ok = int(torch.equal(a, b))
    okt = torch.tensor([ok], dtype=torch.int32, device="cuda")
    if world > 1:
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"config": "3 FIR 4095 taps, time-sharded", "n_gpus": world, "total_samples": args.total,
                          "ms_per_step": ms, "Msamples_s": args.total / ms / 1e3, "tflops_total": 16380.0 * args.total / ms / 1e9,
                          "halo_bytes_per_boundary": (T - 1) * 8, "boundaries_bit_exact": bool(okt.item()), "scaling": "strong"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
