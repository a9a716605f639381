import numpy as np, sys, os
sys.path.insert(0, os.getcwd())
from oracle import loader
from qdsp_b200 import blocks as B, synth
n = 700_001
for order, gen in [(2, "bpsk"), (4, "qpsk"), (8, "qpsk")]:
    x = (synth.qpsk_cf32 if gen == "qpsk" else synth.bpsk_cf32)(35, 0, n)
    yo, st = loader.port().costas(order, 0.004, x)
    for cuts in ([0, n], [0, 300_003, 300_003 + 2048 * 37 + 5, n]):
        pl = B.CostasLoop(order, 0.004)
        y = np.concatenate([pl.process(x[a:b]) for a, b in zip(cuts[:-1], cuts[1:])])
        d = np.abs(y - yo)
        print(order, cuts, "max", d.max(), "argmax", int(d.argmax()), "resid", pl.last_residual(), flush=True)
