"""tools/size_sweep_cfg1.py — dev probe: configs 1a / 1b (127-tap FIR, FIR + decimate-by-4) at several launch sizes.
Shows how much of a launch is ramp and drain: BASELINE's 2^24 samples are one 52 us launch for 1b.
    python tools/size_sweep_cfg1.py [cfg1a|cfg1b]"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from qdsp_b200 import blocks as B, lib  # noqa: E402

L = lib.load()
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
win = B.BlackmanWindow(300e3, 4 * 2.4e6 / 127, 2.4e6)
only = sys.argv[1] if len(sys.argv) > 1 else None
for logn in (22, 24, 26, 27):
    n = 1 << logn
    nbuf = max(2, (1 << 29) // (n * 8) + 1)          # > L2 in rotation
    nbuf = min(nbuf, 8)
    xs = [B.DevBuf(n * 8) for _ in range(nbuf)]
    ys = [B.DevBuf(n * 8) for _ in range(nbuf)]
    for i, b in enumerate(xs):
        lib.check(L.qdsp_synth_uniform_cf32(b.ptr, 1, i * n, n, sp))
    for name, blk, args in (("cfg1a", B.FIR(win), ()), ("cfg1b", B.PolyphaseResampler(win, 2.4e6, 0.6e6), (524288,))):
        if only and name != only:
            continue
        it = [0]

        def step():
            i = it[0] % nbuf
            it[0] += 1
            blk.process_device(xs[i].ptr, ys[i].ptr, n, *args, stream=sp)

        for _ in range(5):
            step()
        torch.cuda.synchronize()
        k = max(10, (1 << 31) // n // (8 if name == "cfg1a" else 2))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / k
        print(json.dumps({"config": name, "samples": n, "launches": k, "us_per_launch": round(ms * 1e3, 2), "GS_per_s": round(n / ms / 1e6, 1)}), flush=True)
    for b in xs + ys:
        b.free()
