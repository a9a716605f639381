"""Seeded parity cases shared by tools/make_golden.py (which runs the UNMODIFIED reference on them and
commits the outputs to tests/golden/), tests/test_oracle.py (C restatement vs those outputs) and
tests/test_gpu_parity.py (CUDA path vs both). Sizes are small enough that the fixture file stays small."""
import numpy as np

from qdsp_b200 import synth

FS = 2.4e6
CASES = {
    # config 1a: complex FIR, 127 taps (BlackmanWindow(300e3, 4*fs/127, fs)), ragged blocks
    "fir127": dict(kind="fir", win=(300e3, 4 * FS / 127, FS), src=("uniform", 1, 0, 6000), block=[2048, 1000, 2952]),
    "fir127_f32": dict(kind="fir_f32", win=(300e3, 4 * FS / 127, FS), src=("uniform_f32", 11, 0, 5000), block=1024),
    # config 1b: the same window through PolyphaseResampler fs -> fs/4 (I=1, D=4)
    "decim4": dict(kind="resamp", win=(300e3, 4 * FS / 127, FS), in_sr=FS, out_sr=FS / 4, src=("uniform", 1, 0, 16384), block=4096),
    "decim4_ragged": dict(kind="resamp", win=(300e3, 4 * FS / 127, FS), in_sr=FS, out_sr=FS / 4, src=("uniform", 1, 0, 9001), block=[1001, 4, 3, 0, 4095, 3898]),
    # genuinely rational: 250 kS/s -> 48 kS/s => I=24, D=125, 999 taps, TPP=42 (VFO-style window)
    "rational": dict(kind="resamp", win=(24e3, 24e3, 250e3), in_sr=250e3, out_sr=48e3, vfo_style=True, src=("uniform", 7, 0, 10000), block=[1000, 777, 4096, 1001, 3126]),
    "rational_f32": dict(kind="resamp_f32", win=(3e3, 3e3, 48e3), in_sr=48e3, out_sr=44.1e3, src=("uniform_f32", 8, 0, 6000), block=[1470, 1471, 3059]),
    "interp": dict(kind="resamp", win=(10e3, 5e3, 48e3), in_sr=48e3, out_sr=192e3, src=("uniform", 9, 0, 3000), block=1000),
    "power_decim1": dict(kind="power_decim", power=1, src=("uniform", 12, 0, 4096), block=1024),
    # power > 1 re-reads the input buffer AFTER flush() (resampling.h:235-245): racy against the next swap in the
    # reference, so the golden case is a single run() call
    "power_decim3": dict(kind="power_decim", power=3, src=("uniform", 12, 0, 4096), block=4096),
    # NCO translator alone (float recursive phasor in the reference: short enough that drift is < 1e-5)
    "xlator": dict(kind="xlator", fs=FS, freq=-250e3, src=("uniform", 2, 0, 3000), block=[1000, 777, 1223]),
    "fm": dict(kind="fm", fs=48e3, dev=5e3, src=("fm48k", 0, 5000), block=[1024, 1, 3975]),
    "fm_stereo": dict(kind="fm_stereo", fs=48e3, dev=5e3, src=("fm48k", 0, 2000), block=1000),
    # "next" row: the WFM stereo tail (FloatFMDemod -> pilot FIR<float> 961 taps -> AGC -> matrix) at 240 kS/s
    "stereo_fm": dict(kind="stereo_fm", fs=240e3, dev=75e3, src=("stereo_mpx", 0, 12000), block=[5000, 3000, 4000]),
    # VFO alone and config 2's fused chain: 2.4 MS/s -> 48 kS/s (401 taps, I=1, D=50) + FloatFMDemod
    "vfo": dict(kind="vfo", offset=250e3, in_sr=FS, out_sr=48e3, bw=48e3, src=("cfg2", 0, 40000), block=10000),
    "vfo_fm": dict(kind="vfo_fm", offset=250e3, in_sr=FS, out_sr=48e3, bw=48e3, dev=5e3, src=("cfg2", 0, 81920), block=8192 * 5),
    "vfo_fm_ragged": dict(kind="vfo_fm", offset=250e3, in_sr=FS, out_sr=48e3, bw=48e3, dev=5e3, src=("cfg2", 0, 50021), block=[10000, 7777, 49, 0, 51, 20001, 12143]),
    # config 4 in miniature: 4 channels off one wideband stream (6.144 MS/s -> 48 kS/s: D=128, 1025 taps)
    "channelizer": dict(kind="channelizer", offsets=[-360e3, -120e3, 120e3, 360e3], in_sr=6.144e6, out_sr=48e3, bw=48e3, dev=5e3, src=("cfg4mini", 0, 65536), block=16384),
    # config 5: recurrent blocks
    "deemp": dict(kind="deemp", fs=48e3, tau=50e-6, src=("uniform", 5, 0, 6000), block=[1000, 5000]),
    "agc": dict(kind="agc", fall=20.0, fs=48e3, src=("uniform_f32", 6, 0, 8000), block=[1000, 3000, 4000]),
    "cagc": dict(kind="cagc", set_point=1.0, max_gain=65535.0, rate=1e-3, src=("qpsk_am", 21, 0, 8000), block=4000),
    "ffagc": dict(kind="ffagc", src=("qpsk_am", 22, 0, 6000), block=[500, 600, 2000, 2900]),
    "costas4": dict(kind="costas", order=4, bw=0.004, src=("qpsk", 23, 0, 8000), block=4000),
    "costas2": dict(kind="costas", order=2, bw=0.004, src=("bpsk", 24, 0, 8000), block=4000),
    "costas8": dict(kind="costas", order=8, bw=0.004, src=("qpsk", 25, 0, 4000), block=4000),
    # ---- "next" rows: element-wise / layout / per-block-statistic blocks (math.h, audio.h, convertion.h, processing.h,
    # demodulator.h). Two-input blocks take their second stream from `src2`.
    "math_cf32": dict(kind="math", src=("uniform", 31, 0, 2000), src2=("uniform", 32, 0, 2000), block=[1000, 3, 997]),
    "math_f32": dict(kind="math", src=("uniform_f32", 33, 0, 2001), src2=("uniform_f32", 34, 0, 2001), block=667),
    "mono_to_stereo": dict(kind="layout", op=0, src=("uniform_f32", 35, 0, 1500), block=[1000, 500]),
    "channels_to_stereo": dict(kind="layout", op=1, src=("uniform_f32", 36, 0, 1500), src2=("uniform_f32", 37, 0, 1500), block=[1, 1499]),
    "stereo_to_mono": dict(kind="layout", op=2, src=("uniform", 38, 0, 1500), block=500),
    "stereo_to_channels": dict(kind="layout", op=3, src=("uniform", 39, 0, 1500), block=[700, 800]),
    "complex_to_stereo": dict(kind="layout", op=4, src=("uniform", 40, 0, 1000), block=1000),
    "complex_to_real": dict(kind="layout", op=5, src=("uniform", 41, 0, 1501), block=[1000, 501]),
    "complex_to_imag": dict(kind="layout", op=6, src=("uniform", 42, 0, 1501), block=[1000, 501]),
    "real_to_complex": dict(kind="layout", op=7, src=("uniform_f32", 43, 0, 1500), block=[1000, 500]),
    "volume_f32": dict(kind="volume", volume=0.7, call_set=1, muted=0, src=("uniform_f32", 44, 0, 2002), block=[1001, 1001]),
    # the reference's init() stores the volume but leaves the applied level at 1.0 (processing.h:355-359)
    "volume_stereo_noset": dict(kind="volume", volume=0.7, call_set=0, muted=0, src=("uniform", 45, 0, 1000), block=1000),
    "volume_muted": dict(kind="volume", volume=0.7, call_set=1, muted=1, src=("uniform", 46, 0, 1000), block=500),
    "threshold": dict(kind="threshold", src=("uniform_f32", 47, 0, 2000), block=[1999, 1]),
    "delay_imag": dict(kind="delay_imag", src=("uniform", 48, 0, 2000), block=[1000, 1, 999]),
    "amdemod": dict(kind="amdemod", src=("qpsk_am", 49, 0, 9000), block=[4000, 1000, 4000]),
    # three run() blocks at about -1 dB, -61 dB and -27 dB mean magnitude against a -30 dB gate: pass, mute, pass
    "squelch": dict(kind="squelch", level=-30.0, src=("uniform_steps", 50, [1000, 500, 1500], [1.0, 0.001, 0.05]), block=[1000, 500, 1500]),
    "ssb_usb": dict(kind="ssb", fs=48e3, bw=3e3, mode=0, src=("uniform", 51, 0, 3000), block=[1000, 777, 1223]),
    "ssb_lsb": dict(kind="ssb", fs=48e3, bw=3e3, mode=1, src=("uniform", 52, 0, 3000), block=3000),
    # symbol-timing recovery and the hier demodulators built on it (clock_recovery.h:68-243, demodulator.h:499-682);
    # 4 samples per symbol, the reference's default loop gains
    "mm_cf32": dict(kind="mm", omega=4.0, gain_omega=(0.01 * 0.01) / 4, mu_gain=0.01, rel=0.005, src=("qpsk_clean", 61, 0, 20000), block=[8000, 7000, 5000]),
    "mm_f32": dict(kind="mm", omega=4.0, gain_omega=(0.01 * 0.01) / 4, mu_gain=0.01, rel=0.005, src=("bpsk_real", 62, 0, 16000), block=4000),
    "msk_demod": dict(kind="msk", fs=48e3, dev=3e3, baud=12e3, src=("msk", 63, 0, 16000), block=[6000, 10000]),
    "psk_demod4": dict(kind="psk", order=4, offset=0, fs=48e3, baud=12e3, src=("qpsk_rrc", 64, 0, 24000), block=[8000, 16000]),
    "psk_demod2": dict(kind="psk", order=2, offset=0, fs=48e3, baud=12e3, src=("bpsk_rrc", 65, 0, 16000), block=8000),
    "ssb_dsb": dict(kind="ssb", fs=48e3, bw=3e3, mode=2, src=("uniform", 53, 0, 1000), block=[600, 400]),
}


def make_input(c, key="src") -> np.ndarray:
    src = c[key]
    k = src[0]
    if k == "uniform":
        return synth.uniform_cf32(src[1], src[2], src[3])
    if k == "uniform_f32":
        return synth.uniform_f32(src[1], src[2], src[3])
    if k == "cfg2":
        return synth.cfg2_input(src[1], src[2])
    if k == "fm48k":  # an FM signal already at 48 kS/s (demod-only cases)
        return synth.fm_cf32(src[1], src[2], 48_000, 3_000, 400, 5e3, 0.7) + np.float32(0.01) * synth.uniform_cf32(3, src[1], src[2])
    if k == "stereo_mpx":
        return synth.stereo_mpx_fm_cf32(src[1], src[2])
    if k == "cfg4mini":
        return synth.cfg4_input(src[1], src[2], nch=4, fs=6_144_000, spacing=240_000)
    if k == "qpsk":
        return synth.qpsk_cf32(src[1], src[2], src[3])
    if k == "bpsk":
        return synth.bpsk_cf32(src[1], src[2], src[3])
    if k == "qpsk_clean":
        return synth.qpsk_cf32(src[1], src[2], src[3], sps=4, freq_off=0.0, sigma=0.05)
    if k == "bpsk_real":
        return np.ascontiguousarray(synth.bpsk_cf32(src[1], src[2], src[3], sps=4, freq_off=0.0, sigma=0.05).real)
    if k == "msk":   # continuous-phase binary FSK, +-dev, 4 samples per symbol
        n = np.arange(src[2], src[2] + src[3], dtype=np.int64)
        bits = 1.0 - 2.0 * (synth._splitmix64((np.uint64(src[1] + 77) << np.uint64(40)) ^ (n // 4).astype(np.uint64)) & np.uint64(1)).astype(np.float64)
        ph = 2.0 * np.pi * 3e3 / 48e3 * np.cumsum(bits)
        return (0.8 * np.exp(1j * ph)).astype(np.complex64) + np.float32(0.01) * synth.uniform_cf32(src[1], src[2], src[3])
    if k in ("qpsk_rrc", "bpsk_rrc"):   # rectangular symbols, small carrier offset: the chain's own RRC does the shaping
        f = synth.qpsk_cf32 if k == "qpsk_rrc" else synth.bpsk_cf32
        return f(src[1], src[2], src[3], sps=4, freq_off=0.002, sigma=0.03)
    if k == "uniform_steps":  # uniform noise scaled per run() block
        sizes, gains = src[2], src[3]
        x = synth.uniform_cf32(src[1], 0, int(sum(sizes)))
        g = np.repeat(np.asarray(gains, np.float32), sizes)
        return (x * g).astype(np.complex64)
    if k == "qpsk_am":
        return synth.qpsk_cf32(src[1], src[2], src[3], am_depth=0.5, am_period=3000)
    raise ValueError(k)
