"""dev tool: CostasLoop<4> chunk / warm-up sweep on config-5 input (2^26 samples): time, boundary residual, error vs the
sequential oracle on a prefix."""
import ctypes as C, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qdsp_b200 import blocks as B, lib, synth
from oracle import loader
L = lib.load()
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
n = 1 << 26
npre = 1 << 19
xh = synth.qpsk_cf32(23, 0, npre)
ref = loader.port().costas(4, 0.004, xh)[0]
x = torch.empty(n, dtype=torch.complex64, device="cuda")
reps = n // npre
xt = torch.from_numpy(xh).cuda()
for r in range(reps):
    x[r * npre:(r + 1) * npre] = xt   # periodic extension (phase jumps at the seams are part of the stress)
y = torch.empty(n, dtype=torch.complex64, device="cuda")
for chunk, warm in [(2048, 2048), (1024, 2048), (1024, 1536), (512, 2048)]:
    pl = B.CostasLoop(4, 0.004)
    pl.set_chunking(chunk, warm)
    pl.process_device(x.data_ptr(), y.data_ptr(), n, stream=sp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pl2 = pl
    pl2.set_state(np.zeros(4, np.float32) + np.asarray([0, 0, 1, 0], np.float32))
    e0.record()
    pl2.process_device(x.data_ptr(), y.data_ptr(), n, stream=sp)
    e1.record()
    torch.cuda.synchronize()
    err = float(np.abs(y[:npre].cpu().numpy() - ref).max())
    print(json.dumps({"chunk": chunk, "warmup": warm, "ms": e0.elapsed_time(e1), "GS_s": n / e0.elapsed_time(e1) / 1e6,
                      "residual": pl2.last_residual(), "max_err_vs_sequential_prefix": err}), flush=True)
