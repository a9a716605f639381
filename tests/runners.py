"""Run one tests/cases.py case through (a) the plain-C oracle restatement, (b) the CUDA product path."""
import numpy as np

from oracle import loader
from tests.cases import make_input


def interp_taps():
    """INTERP_TAPS[129][8] of the reference (src/dsp/interpolation_taps.h), from the committed reference vectors."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.npz"))
    return g["interp_taps"]


def _stereo(x):
    return np.ascontiguousarray(x).view(np.float32).reshape(-1, 2)


def run_port(c, x):
    """Returns (output, out_counts or None) from oracle/port.c."""
    P = loader.port()
    k = c["kind"]
    if k == "fir":
        return P.fir_cf32(P.blackman_taps(*c["win"]), x), None
    if k == "fir_f32":
        return P.fir_f32(P.blackman_taps(*c["win"]), x), None
    if k in ("resamp", "resamp_f32"):
        i, d = P.rates_to_ratio(c["in_sr"], c["out_sr"])
        cutoff, tw, wfs = c["win"]
        if c.get("vfo_style"):
            wfs = float(np.float32(c["in_sr"]) * np.float32(i))
        taps = P.blackman_taps(cutoff, tw, wfs, factor=float(i))
        fn = P.resamp_cf32 if k == "resamp" else P.resamp_f32
        y, oc = fn(taps, i, d, x, c["block"])
        return y, oc
    if k == "power_decim":
        return P.power_decim(c["power"], x, c["block"])
    if k == "xlator":
        inc = P.xlator_phase_delta(c["fs"], c["freq"])
        y, _ = P.rotator(x, inc, 1 + 0j, c["block"])
        return y, None
    if k == "fm":
        return P.fm_demod(c["fs"], c["dev"], x), None
    if k == "fm_stereo":
        y = P.fm_demod(c["fs"], c["dev"], x)
        return np.stack([y, y], axis=1), None
    if k == "stereo_fm":
        return P.stereo_fm(c["fs"], c["dev"], x, c["block"]), None
    if k == "vfo":
        a, oc, iq = P.vfo_fm(c["offset"], c["in_sr"], c["out_sr"], c["bw"], 5e3, x, c["block"], want_iq=True)
        return iq, oc
    if k == "vfo_fm":
        return P.vfo_fm(c["offset"], c["in_sr"], c["out_sr"], c["bw"], c["dev"], x, c["block"])
    if k == "channelizer":
        rows = [P.vfo_fm(o, c["in_sr"], c["out_sr"], c["bw"], c["dev"], x, c["block"])[0] for o in c["offsets"]]
        return np.stack(rows), None
    if k == "deemp":
        return P.deemp(c["fs"], c["tau"], _stereo(x)).reshape(-1).view(np.complex64), None
    if k == "agc":
        return P.agc(c["fall"], c["fs"], x, c["block"]), None
    if k == "cagc":
        return P.complex_agc(c["set_point"], c["max_gain"], c["rate"], x), None
    if k == "ffagc":
        return P.ff_agc(x), None
    if k == "costas":
        return P.costas(c["order"], c["bw"], x)[0], None
    if k == "math":
        x2 = make_input(c, "src2")
        return np.stack([P.math(op, x, x2) for op in range(3)]), None
    if k == "layout":
        x2 = make_input(c, "src2") if "src2" in c else None
        y = P.layout(c["op"], x, x2)
        return (np.stack(y) if c["op"] == 3 else y), None
    if k == "volume":
        return P.volume(x, c["volume"], c["call_set"], c["muted"]), None
    if k == "threshold":
        return P.threshold(x), None
    if k == "delay_imag":
        return P.delay_imag(x), None
    if k == "amdemod":
        return P.amdemod(x, c["block"]), None
    if k == "squelch":
        return P.squelch(c["level"], x, c["block"]), None
    if k == "ssb":
        return P.ssbdemod(c["fs"], c["bw"], c["mode"], x, c["block"]), None
    if k == "mm":
        return P.mm(x, c["omega"], c["gain_omega"], c["mu_gain"], c["rel"], interp_taps(), c["block"])
    if k == "msk":
        # MSKDemod = FloatFMDemod -> MMClockRecovery<float> with the constructor's default gains (demodulator.h:503-520)
        a = P.fm_demod(c["fs"], c["dev"], x)
        return P.mm(a, float(np.float32(c["fs"]) / np.float32(c["baud"])), (0.01 * 0.01) / 4, 0.01, 0.005, interp_taps(), c["block"])
    if k == "psk":
        # PSKDemod<ORDER, false> = ComplexAGC(1, 65535, 1e-3) -> FIR(RRCTaps(32, fs, baud, 0.32)) -> CostasLoop -> MM
        y = P.complex_agc(1.0, 65535.0, 10e-4, x)
        # FIR::init asks for 32 taps; RRCTaps::createTaps writes 33 (count | 1, window.h:186) and normalises over all
        # of them; the filter runs over the first 32 (filter.h:24-27)
        y = P.fir_cf32(np.ascontiguousarray(P.rrc_taps(32, c["fs"], c["baud"], 0.32)[:32]), y)
        y = P.costas(c["order"], 0.004, y)[0]
        return P.mm(y, float(np.float32(c["fs"]) / np.float32(c["baud"])), (0.01 * 0.01) / 4, 0.01, 0.005, interp_taps(), c["block"])
    raise ValueError(k)


def run_gpu(c, x, variant=0):
    """Returns (output, out_counts or None) from libqdsp_b200.so through the block mirror."""
    from qdsp_b200 import blocks as B

    k = c["kind"]
    if k in ("fir", "fir_f32"):
        f = B.FIR(B.BlackmanWindow(*c["win"]), np.complex64 if k == "fir" else np.float32)
        f.set_variant(variant)
        # feed block by block like the reference's run() calls (state carried in the handle)
        outs, off = [], 0
        sizes = loader.as_blocks(len(x), c["block"])
        for s in sizes:
            outs.append(f.process(x[off:off + s]))
            off += s
        return np.concatenate(outs), None
    if k in ("resamp", "resamp_f32"):
        cutoff, tw, wfs = c["win"]
        win = B.BlackmanWindow(cutoff, tw, wfs)
        if c.get("vfo_style"):
            i, _ = B.rates_to_ratio(c["in_sr"], c["out_sr"])
            win.setSampleRate(np.float32(c["in_sr"]) * np.float32(i))
        r = B.PolyphaseResampler(win, c["in_sr"], c["out_sr"], np.complex64 if k == "resamp" else np.float32)
        r.set_variant(variant)
        y = r.process(x, c["block"])
        return y, r.last_out_counts
    if k == "power_decim":
        d = B.PowerDecimator(c["power"])
        sizes = loader.as_blocks(len(x), c["block"])
        outs, oc, off = [], [], 0
        for s in sizes:
            y = d.process(x[off:off + s])
            outs.append(y)
            oc.append(len(y))
            off += s
        return np.concatenate(outs), np.asarray(oc, np.int32)
    if k == "xlator":
        xl = B.FrequencyXlator(c["fs"], c["freq"])
        sizes = loader.as_blocks(len(x), c["block"])
        outs, off = [], 0
        for s in sizes:
            outs.append(xl.process(x[off:off + s]))
            off += s
        return np.concatenate(outs), None
    if k in ("fm", "fm_stereo"):
        dm = (B.FloatFMDemod if k == "fm" else B.FMDemod)(c["fs"], c["dev"])
        sizes = loader.as_blocks(len(x), c["block"])
        outs, off = [], 0
        for s in sizes:
            outs.append(dm.process(x[off:off + s]))
            off += s
        y = np.concatenate(outs)
        if k == "fm_stereo":
            y = y.view(np.float32).reshape(-1, 2)
        return y, None
    if k == "stereo_fm":
        y = B.StereoFMDemod(c["fs"], c["dev"]).process(x, c["block"])
        return y.view(np.float32).reshape(-1, 2), None
    if k == "vfo":
        v = B.VFOFM(c["offset"], c["in_sr"], c["out_sr"], c["bw"], 5e3)
        v.set_variant(variant)
        v.want_iq = True
        v.process(x, c["block"])
        return v.last_iq, v.last_out_counts
    if k == "vfo_fm":
        v = B.VFOFM(c["offset"], c["in_sr"], c["out_sr"], c["bw"], c["dev"])
        v.set_variant(variant)
        y = v.process(x, c["block"])
        return y, v.last_out_counts
    if k == "channelizer":
        ch = B.Channelizer(c["offsets"], c["in_sr"], c["out_sr"], c["bw"], c["dev"])
        ch.set_variant(variant)
        return ch.process(x, c["block"]), None
    if k == "deemp":
        d = B.BFMDeemp(c["fs"], c["tau"])
        sizes = loader.as_blocks(len(x), c["block"])
        outs, off = [], 0
        for s in sizes:
            outs.append(d.process(x[off:off + s]))
            off += s
        return np.concatenate(outs), None
    if k == "agc":
        return B.AGC(c["fall"], c["fs"]).process(x, c["block"]), None
    if k == "cagc":
        g = B.ComplexAGC(c["set_point"], c["max_gain"], c["rate"])
        sizes = loader.as_blocks(len(x), c["block"])
        outs, off = [], 0
        for s in sizes:
            outs.append(g.process(x[off:off + s]))
            off += s
        return np.concatenate(outs), None
    if k == "ffagc":
        g = B.FeedForwardAGC()
        sizes = loader.as_blocks(len(x), c["block"])
        outs, off = [], 0
        for s in sizes:
            outs.append(g.process(x[off:off + s]))
            off += s
        return np.concatenate(outs), None
    if k == "math":
        x2 = make_input(c, "src2")
        outs = []
        for cls in (B.Add, B.Substract, B.Multiply):
            blk, off, parts = cls(x.dtype), 0, []
            for s in loader.as_blocks(len(x), c["block"]):
                parts.append(blk.process(x[off:off + s], x2[off:off + s]))
                off += s
            outs.append(np.concatenate(parts))
        return np.stack(outs), None
    if k == "layout":
        op = c["op"]
        sizes = loader.as_blocks(len(x), c["block"])
        if op == 1:
            x2 = make_input(c, "src2")
            blk, off, parts = B.ChannelsToStereo(), 0, []
            for s in sizes:
                parts.append(blk.process(x[off:off + s], x2[off:off + s]))
                off += s
            return np.concatenate(parts), None
        if op == 3:
            blk, off, ls, rs = B.StereoToChannels(), 0, [], []
            for s in sizes:
                l, r = blk.process(x[off:off + s])
                ls.append(l)
                rs.append(r)
                off += s
            return np.stack([np.concatenate(ls), np.concatenate(rs)]), None
        blk = {0: B.MonoToStereo, 2: B.StereoToMono, 4: B.ComplexToStereo, 5: B.ComplexToReal, 6: B.ComplexToImag,
               7: B.RealToComplex}[op]()
        off, parts = 0, []
        for s in sizes:
            parts.append(blk.process(x[off:off + s]))
            off += s
        return np.concatenate(parts), None
    if k in ("volume", "threshold", "delay_imag"):
        if k == "volume":
            blk = B.Volume(c["volume"], x.dtype)
            if c["call_set"]:
                blk.setVolume(c["volume"])
            blk.setMuted(bool(c["muted"]))
        else:
            blk = B.Threshold() if k == "threshold" else B.DelayImag()
        off, parts = 0, []
        for s in loader.as_blocks(len(x), c["block"]):
            parts.append(blk.process(x[off:off + s]))
            off += s
        return np.concatenate(parts), None
    if k == "amdemod":
        return B.AMDemod().process(x, c["block"]), None
    if k == "squelch":
        return B.Squelch(c["level"]).process(x, c["block"]), None
    if k == "ssb":
        blk, off, parts = B.SSBDemod(c["fs"], c["bw"], c["mode"]), 0, []
        for s in loader.as_blocks(len(x), c["block"]):
            parts.append(blk.process(x[off:off + s]))
            off += s
        return np.concatenate(parts), None
    if k == "mm":
        mm = B.MMClockRecovery(c["omega"], c["gain_omega"], c["mu_gain"], c["rel"], interp_taps(), x.dtype)
        y = mm.process(x, c["block"])
        return y, mm.last_out_counts
    if k == "msk":
        d = B.MSKDemod(c["fs"], c["dev"], c["baud"], interp_taps())
        y = d.process(x, c["block"])
        return y, d.last_out_counts
    if k == "psk":
        d = B.PSKDemod(c["order"], c["offset"], c["fs"], c["baud"], interp_taps())
        d.demod.set_chunking(c.get("chunk", 0), c.get("warmup", 0))
        y = d.process(x, c["block"])
        return y, d.last_out_counts
    if k == "costas":
        pl = B.CostasLoop(c["order"], c["bw"])
        pl.set_chunking(c.get("chunk", 0), c.get("warmup", 0))
        sizes = loader.as_blocks(len(x), c["block"])
        outs, off = [], 0
        for s in sizes:
            outs.append(pl.process(x[off:off + s]))
            off += s
        return np.concatenate(outs), None
    raise ValueError(k)


def rel_l2(a, b):
    a = np.asarray(a).astype(np.complex128 if np.iscomplexobj(a) else np.float64)
    b = np.asarray(b).astype(a.dtype)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))
