"""oracle/windows.py — TEST INFRASTRUCTURE: window checks of full-size runs against the CPU oracle (oracle/port.c).

A BASELINE-size stream (2^28 .. 2^30 samples) is too long to push through the CPU oracle as a whole, so the
full-size gates compare WINDOWS: the caller hands over a function that returns any slice of the exact input the
GPU consumed and the GPU's output array; the oracle recomputes each window from scratch (with enough pre-roll
for the filter history) and the overlap is compared. Used by tests/test_gpu_fullsize.py and by the `parity`
blocks of bench.py (never inside a timed region)."""
from __future__ import annotations

import numpy as np

from . import loader


def pick_windows(total: int, block: int, k: int, span: int, seams=(), seed: int = 1234):
    """Window centres: every seam (sample index) + k pseudo-random positions; windows are `span` samples long."""
    rng = np.random.default_rng(seed)
    cs = [int(s) for s in seams if 0 < s < total]
    cs += [int(c) for c in rng.integers(span, max(total - span, span + 1), size=k)]
    cs += [span // 2 + 1, total - span // 2 - 1]          # the two ends of the stream
    return sorted(set(min(max(c, span // 2), total - span // 2) for c in cs))


def check_vfofm_windows(get_input, audio: np.ndarray, total: int, block: int, centres, span: int, offset: float,
                        in_sr: float, out_sr: float, bw: float, dev: float, decim: int, taps: int):
    """Fused chain (config 2 / one channel of config 4) with uniform run() blocks of `block` samples
    (block % decim == 0, so output k of the stream is sample k*decim). Returns (max abs error, outputs compared)."""
    assert block % decim == 0
    P = loader.port()
    pre = (taps // decim + 3) * decim            # pre-roll: filter history + the demodulator's previous sample
    skip = taps // decim + 3
    worst, n_cmp = 0.0, 0
    for c in centres:
        lo = max(0, (c - span // 2) // decim * decim - pre)
        hi = min(total, lo + pre + span)
        x = get_input(lo, hi)
        # the run() boundaries that fall inside the window (the resampler restarts its schedule there)
        cuts = [lo] + [b for b in range((lo // block + 1) * block, hi, block)] + [hi]
        sizes = [b - a for a, b in zip(cuts[:-1], cuts[1:])]
        a, _, _ = P.vfo_fm_window(offset, in_sr, out_sr, bw, dev, x, sizes, lo)
        k0 = lo // decim
        first = 0 if lo == 0 else skip
        if lo == 0:
            first = 16                           # start-up: |y| ~ 0, the angle of rounding noise
        g = audio[k0 + first:k0 + len(a)]
        if len(g) == 0:
            continue
        worst = max(worst, float(np.abs(g - a[first:first + len(g)]).max()))
        n_cmp += len(g)
    return worst, n_cmp


def check_fir_windows(get_input, y: np.ndarray | None, get_output, total: int, taps: np.ndarray, centres, span: int):
    """FIR<complex_t> (config 1a / 3): y[i] = sum_j taps[j] x[i-(T-1)+j]. `get_output(lo, hi)` returns GPU outputs.
    Returns (worst window rel-L2, max abs error, outputs compared)."""
    P = loader.port()
    T = len(taps)
    worst_rel, worst_abs, n_cmp = 0.0, 0.0, 0
    for c in centres:
        lo = max(0, c - span // 2 - (T - 1))
        hi = min(total, lo + span + (T - 1))
        x = get_input(lo, hi)
        yo = P.fir_cf32(taps, x)
        first = 0 if lo == 0 else T - 1
        g = get_output(lo + first, hi) if get_output else y[lo + first:hi]
        o = yo[first:]
        d = g - o
        worst_rel = max(worst_rel, float(np.linalg.norm(d) / max(np.linalg.norm(o), 1e-30)))
        worst_abs = max(worst_abs, float(np.abs(d).max()))
        n_cmp += len(o)
    return worst_rel, worst_abs, n_cmp
