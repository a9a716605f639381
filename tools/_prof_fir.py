import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qdsp_b200 import blocks as B, lib
L = lib.load()
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
n = 1 << 24
x = torch.empty(n, dtype=torch.complex64, device="cuda"); y = torch.empty_like(x)
lib.check(L.qdsp_synth_uniform_cf32(x.data_ptr(), 3, 0, n, sp))
f = B.FIR(B.BlackmanWindow(100e3, 4 * 2.4e6 / 4095, 2.4e6))
for _ in range(3):
    f.process_device(x.data_ptr(), y.data_ptr(), n, stream=sp)
torch.cuda.synchronize()
print("ok")
