"""oracle/loader.py — TEST INFRASTRUCTURE ONLY.

ctypes bindings for the CPU oracle:
  * ``ref(variant)``  -> oracle/_ref/libqdsp_ref*.so, the UNMODIFIED reference headers compiled
    against the VOLK shim (built in the authoring container, travels to the GPU box prebuilt);
  * ``port()``        -> oracle/liboracle_port.so, the plain-C restatement (always buildable).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. The product package (qdsp_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from functools import lru_cache

import math

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_VARIANTS = {
    "generic": "libqdsp_ref.so",        # parity oracle (VOLK generic semantics)
    "f64nco": "libqdsp_ref_f64nco.so",  # drift-free rotator, for NCO attribution
    "fast": "libqdsp_ref_fast.so",      # SIMD-order dots + -march=x86-64-v3, timing baseline
}

_f = C.c_float
_i = C.c_int
_ll = C.c_longlong
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_dp = C.POINTER(C.c_double)


def build(force: bool = False) -> None:
    """Compile the C port and (when /root/reference exists) the reference drivers."""
    args = ["make", "-s", "-C", HERE, "all"]
    if force:
        subprocess.check_call(["make", "-s", "-C", HERE, "clean"])
    subprocess.check_call(args)


def have_ref(variant: str = "generic") -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", REF_VARIANTS[variant]))


def _fptr(a: np.ndarray):
    return a.ctypes.data_as(_fp)


def _iptr(a: np.ndarray):
    return a.ctypes.data_as(_ip)


def as_blocks(n: int, block: int | list[int] | np.ndarray) -> np.ndarray:
    """Block partition of an n-sample stream: uniform `block` (last one short) or explicit list."""
    if np.isscalar(block):
        b = int(block)
        sizes = [b] * (n // b)
        if n % b:
            sizes.append(n % b)
    else:
        sizes = [int(x) for x in block]
        assert sum(sizes) == n, (sum(sizes), n)
    return np.asarray(sizes, dtype=np.int32)


class Ref:
    """Thin numpy-facing wrapper over oracle/_ref/libqdsp_ref*.so (oracle/ref_driver.cpp)."""

    def __init__(self, variant: str = "generic"):
        path = os.path.join(HERE, "_ref", REF_VARIANTS[variant])
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
        self.variant = variant
        self.lib = L = C.CDLL(path)
        L.ref_blackman_tap_count.argtypes = [_f, _f, _f]
        L.ref_blackman_tap_count.restype = _i
        L.ref_blackman_taps.argtypes = [_f, _f, _f, _fp, _i, _f]
        L.ref_blackman_bandpass_tap_count.argtypes = [_f, _f, _f, _f]
        L.ref_blackman_bandpass_tap_count.restype = _i
        L.ref_blackman_bandpass_taps.argtypes = [_f, _f, _f, _f, _fp, _i, _f]
        L.ref_rrc_taps.argtypes = [_i, _f, _f, _f, _fp]
        for name in ("ref_fir_cf32", "ref_fir_f32"):
            fn = getattr(L, name)
            fn.argtypes = [_f, _f, _f, _fp, _ip, _i, _fp, _dp]
            fn.restype = _ll
        L.ref_resamp_cf32.argtypes = [_f, _f, _f, _f, _f, _i, _fp, _ip, _i, _fp, _ip, _ip, _ip, _dp]
        L.ref_resamp_cf32.restype = _ll
        L.ref_resamp_f32.argtypes = [_f, _f, _f, _f, _f, _fp, _ip, _i, _fp, _ip, _ip, _ip, _dp]
        L.ref_resamp_f32.restype = _ll
        L.ref_power_decim.argtypes = [C.c_uint, _fp, _ip, _i, _fp, _ip]
        L.ref_power_decim.restype = _ll
        L.ref_xlator.argtypes = [_f, _f, _fp, _ip, _i, _fp, _dp]
        L.ref_xlator.restype = _ll
        L.ref_xlator_phase_delta.argtypes = [_f, _f, _fp, _fp]
        L.ref_rotator.argtypes = [_fp, _fp, _f, _f, _fp, _fp, _ip, _i]
        L.ref_vfo.argtypes = [_f, _f, _f, _f, _fp, _ip, _i, _fp, _ip, _dp]
        L.ref_vfo.restype = _ll
        L.ref_vfo_design.argtypes = [_f, _f, _f, _fp, _i, _ip, _ip]
        L.ref_vfo_design.restype = _i
        L.ref_fm_demod.argtypes = [_f, _f, _fp, _ip, _i, _fp, _dp]
        L.ref_fm_demod.restype = _ll
        L.ref_fm_demod_stereo.argtypes = [_f, _f, _fp, _ip, _i, _fp]
        L.ref_fm_demod_stereo.restype = _ll
        L.ref_fast_arctan2.argtypes = [_f, _f]
        L.ref_fast_arctan2.restype = _f
        L.ref_vfo_fm.argtypes = [_f, _f, _f, _f, _f, _fp, _ip, _i, _fp, _ip, _dp]
        L.ref_vfo_fm.restype = _ll
        L.ref_channelizer_fm.argtypes = [_i, _fp, _f, _f, _f, _f, _fp, _ip, _i, _fp, _ll, _dp]
        L.ref_channelizer_fm.restype = _ll
        L.ref_deemp.argtypes = [_f, _f, _fp, _ip, _i, _fp, _dp]
        L.ref_deemp.restype = _ll
        L.ref_agc.argtypes = [_f, _f, _fp, _ip, _i, _fp, _dp]
        L.ref_agc.restype = _ll
        L.ref_complex_agc.argtypes = [_f, _f, _f, _fp, _ip, _i, _fp, _dp]
        L.ref_complex_agc.restype = _ll
        L.ref_ff_agc_cf32.argtypes = [_fp, _ip, _i, _fp, _ip, _ip]
        L.ref_ff_agc_cf32.restype = _ll
        L.ref_stereo_fm.argtypes = [_f, _f, _fp, _ip, _i, _fp]
        L.ref_stereo_fm.restype = _ll
        L.ref_costas.argtypes = [_i, _f, _fp, _ip, _i, _fp, _dp]
        L.ref_costas.restype = _ll
        L.ref_stream_buffer_size.restype = _i

    # -- tap design ---------------------------------------------------------------------------
    def blackman_tap_count(self, cutoff, tw, fs) -> int:
        return int(self.lib.ref_blackman_tap_count(cutoff, tw, fs))

    def blackman_taps(self, cutoff, tw, fs, factor=1.0, count=None) -> np.ndarray:
        n = self.blackman_tap_count(cutoff, tw, fs) if count is None else count
        t = np.empty(n, np.float32)
        self.lib.ref_blackman_taps(cutoff, tw, fs, _fptr(t), n, factor)
        return t

    def blackman_bandpass_taps(self, cutoff, tw, offset, fs, factor=1.0) -> np.ndarray:
        n = int(self.lib.ref_blackman_bandpass_tap_count(cutoff, tw, offset, fs))
        t = np.empty(n, np.float32)
        self.lib.ref_blackman_bandpass_taps(cutoff, tw, offset, fs, _fptr(t), n, factor)
        return t

    def rrc_taps(self, count, fs, baud, alpha) -> np.ndarray:
        t = np.zeros(count | 1, np.float32)
        self.lib.ref_rrc_taps(count, fs, baud, alpha, _fptr(t))
        return t

    def vfo_design(self, in_sr, out_sr, bw):
        i, d = _i(), _i()
        n = int(self.lib.ref_vfo_design(in_sr, out_sr, bw, None, 0, C.byref(i), C.byref(d)))
        t = np.empty(n, np.float32)
        self.lib.ref_vfo_design(in_sr, out_sr, bw, _fptr(t), n, C.byref(i), C.byref(d))
        return t, i.value, d.value

    # -- streaming blocks (x is complex64 / float32 numpy; block = int or list) ----------------
    @staticmethod
    def _cin(x):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        return x, _fptr(x.view(np.float32))

    def fir_cf32(self, cutoff, tw, fs, x, block, timing=False):
        x, px = self._cin(x)
        b = as_blocks(len(x), block)
        y = np.empty(len(x), np.complex64)
        sec = C.c_double()
        n = self.lib.ref_fir_cf32(cutoff, tw, fs, px, _iptr(b), len(b), _fptr(y.view(np.float32)), C.byref(sec))
        assert n == len(x)
        return (y, sec.value) if timing else y

    def fir_f32(self, cutoff, tw, fs, x, block):
        x = np.ascontiguousarray(x, dtype=np.float32)
        b = as_blocks(len(x), block)
        y = np.empty(len(x), np.float32)
        n = self.lib.ref_fir_f32(cutoff, tw, fs, _fptr(x), _iptr(b), len(b), _fptr(y), None)
        assert n == len(x)
        return y

    def resamp_cf32(self, cutoff, tw, win_fs, in_sr, out_sr, x, block, vfo_style=False, timing=False):
        x, px = self._cin(x)
        b = as_blocks(len(x), block)
        cap = int(len(x) * max(1.0, out_sr / in_sr) * 1.01) + 64 * len(b) + 1024
        y = np.empty(cap, np.complex64)
        oc = np.zeros(len(b), np.int32)
        i, d, sec = _i(), _i(), C.c_double()
        n = self.lib.ref_resamp_cf32(cutoff, tw, win_fs, in_sr, out_sr, int(vfo_style), px, _iptr(b), len(b),
                                     _fptr(y.view(np.float32)), _iptr(oc), C.byref(i), C.byref(d), C.byref(sec))
        res = (y[:n].copy(), oc, i.value, d.value)
        return res + (sec.value,) if timing else res

    def resamp_f32(self, cutoff, tw, win_fs, in_sr, out_sr, x, block):
        x = np.ascontiguousarray(x, dtype=np.float32)
        b = as_blocks(len(x), block)
        cap = int(len(x) * max(1.0, out_sr / in_sr) * 1.01) + 64 * len(b) + 1024
        y = np.empty(cap, np.float32)
        oc = np.zeros(len(b), np.int32)
        i, d = _i(), _i()
        n = self.lib.ref_resamp_f32(cutoff, tw, win_fs, in_sr, out_sr, _fptr(x), _iptr(b), len(b), _fptr(y),
                                    _iptr(oc), C.byref(i), C.byref(d), None)
        return y[:n].copy(), oc, i.value, d.value

    def power_decim(self, power, x, block):
        x, px = self._cin(x)
        b = as_blocks(len(x), block)
        y = np.empty(len(x), np.complex64)
        oc = np.zeros(len(b), np.int32)
        n = self.lib.ref_power_decim(power, px, _iptr(b), len(b), _fptr(y.view(np.float32)), _iptr(oc))
        return y[:n].copy(), oc

    def xlator(self, fs, freq, x, block):
        x, px = self._cin(x)
        b = as_blocks(len(x), block)
        y = np.empty(len(x), np.complex64)
        n = self.lib.ref_xlator(fs, freq, px, _iptr(b), len(b), _fptr(y.view(np.float32)), None)
        assert n == len(x)
        return y

    def xlator_phase_delta(self, fs, freq) -> complex:
        re, im = _f(), _f()
        self.lib.ref_xlator_phase_delta(fs, freq, C.byref(re), C.byref(im))
        return complex(re.value, im.value)

    def rotator(self, x, inc: complex, phase: complex, block):
        x, px = self._cin(x)
        b = as_blocks(len(x), block)
        y = np.empty(len(x), np.complex64)
        pr, pi = _f(np.float32(phase.real)), _f(np.float32(phase.imag))
        self.lib.ref_rotator(px, _fptr(y.view(np.float32)), np.float32(inc.real), np.float32(inc.imag),
                             C.byref(pr), C.byref(pi), _iptr(b), len(b))
        return y, complex(pr.value, pi.value)

    def vfo(self, offset, in_sr, out_sr, bw, x, block):
        x, px = self._cin(x)
        b = as_blocks(len(x), block)
        cap = int(len(x) * max(1.0, out_sr / in_sr) * 1.01) + 64 * len(b) + 1024
        y = np.empty(cap, np.complex64)
        oc = np.zeros(len(b), np.int32)
        n = self.lib.ref_vfo(offset, in_sr, out_sr, bw, px, _iptr(b), len(b), _fptr(y.view(np.float32)), _iptr(oc), None)
        return y[:n].copy(), oc

    def fm_demod(self, fs, dev, x, block):
        x, px = self._cin(x)
        b = as_blocks(len(x), block)
        y = np.empty(len(x), np.float32)
        n = self.lib.ref_fm_demod(fs, dev, px, _iptr(b), len(b), _fptr(y), None)
        assert n == len(x)
        return y

    def fm_demod_stereo(self, fs, dev, x, block):
        x, px = self._cin(x)
        b = as_blocks(len(x), block)
        y = np.empty((len(x), 2), np.float32)
        n = self.lib.ref_fm_demod_stereo(fs, dev, px, _iptr(b), len(b), _fptr(y))
        assert n == len(x)
        return y

    def stereo_fm(self, fs, dev, x, block):
        x, px = self._cin(x)
        b = as_blocks(len(x), block)
        y = np.empty((len(x), 2), np.float32)
        n = self.lib.ref_stereo_fm(fs, dev, px, _iptr(b), len(b), _fptr(y))
        assert n == len(x)
        return y

    def fast_arctan2(self, y, x) -> float:
        return float(self.lib.ref_fast_arctan2(y, x))

    def vfo_fm(self, offset, in_sr, out_sr, bw, dev, x, block, timing=False):
        x, px = self._cin(x)
        b = as_blocks(len(x), block)
        cap = int(len(x) * max(1.0, out_sr / in_sr) * 1.01) + 64 * len(b) + 1024
        y = np.empty(cap, np.float32)
        oc = np.zeros(len(b), np.int32)
        sec = C.c_double()
        n = self.lib.ref_vfo_fm(offset, in_sr, out_sr, bw, dev, px, _iptr(b), len(b), _fptr(y), _iptr(oc), C.byref(sec))
        res = (y[:n].copy(), oc)
        return res + (sec.value,) if timing else res

    def channelizer_fm(self, offsets, in_sr, out_sr, bw, dev, x, block, timing=False):
        x, px = self._cin(x)
        b = as_blocks(len(x), block)
        offs = np.ascontiguousarray(offsets, dtype=np.float32)
        cap = int(len(x) * max(1.0, out_sr / in_sr) * 1.01) + 64 * len(b) + 1024
        y = np.zeros((len(offs), cap), np.float32)
        sec = C.c_double()
        per = self.lib.ref_channelizer_fm(len(offs), _fptr(offs), in_sr, out_sr, bw, dev, px, _iptr(b), len(b),
                                          _fptr(y), cap, C.byref(sec))
        res = y[:, :per].copy()
        return (res, sec.value) if timing else res

    def deemp(self, fs, tau, x, block):
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, 2)
        b = as_blocks(len(x), block)
        y = np.empty_like(x)
        n = self.lib.ref_deemp(fs, tau, _fptr(x), _iptr(b), len(b), _fptr(y), None)
        assert n == len(x)
        return y

    def agc(self, fall_rate, fs, x, block):
        x = np.ascontiguousarray(x, dtype=np.float32)
        b = as_blocks(len(x), block)
        y = np.empty_like(x)
        n = self.lib.ref_agc(fall_rate, fs, _fptr(x), _iptr(b), len(b), _fptr(y), None)
        assert n == len(x)
        return y

    def complex_agc(self, set_point, max_gain, rate, x, block):
        x, px = self._cin(x)
        b = as_blocks(len(x), block)
        y = np.empty(len(x), np.complex64)
        n = self.lib.ref_complex_agc(set_point, max_gain, rate, px, _iptr(b), len(b), _fptr(y.view(np.float32)), None)
        assert n == len(x)
        return y

    def ff_agc(self, x, block):
        x, px = self._cin(x)
        b = as_blocks(len(x), block)
        y = np.empty(len(x), np.complex64)
        vc = np.zeros(len(b), np.int32)
        ns = _i()
        n = self.lib.ref_ff_agc_cf32(px, _iptr(b), len(b), _fptr(y.view(np.float32)), _iptr(vc), C.byref(ns))
        return y[:n].copy(), vc[: ns.value].copy()

    def costas(self, order, bw, x, block, timing=False):
        x, px = self._cin(x)
        b = as_blocks(len(x), block)
        y = np.empty(len(x), np.complex64)
        sec = C.c_double()
        n = self.lib.ref_costas(order, bw, px, _iptr(b), len(b), _fptr(y.view(np.float32)), C.byref(sec))
        assert n == len(x)
        return (y, sec.value) if timing else y


    # -- element-wise / layout / per-block-statistic blocks ("next" rows) ----------------------------------
    def _decl_pointwise(self):
        L = self.lib
        if getattr(self, "_pw", False):
            return
        ucp = C.POINTER(C.c_ubyte)
        L.ref_math.argtypes = [_i, _i, _fp, _fp, _ip, _i, _fp]
        L.ref_math.restype = _ll
        L.ref_layout.argtypes = [_i, _fp, _fp, _ip, _i, _fp, _fp]
        L.ref_layout.restype = _ll
        L.ref_volume.argtypes = [_i, _f, _i, _i, _fp, _ip, _i, _fp]
        L.ref_volume.restype = _ll
        L.ref_threshold.argtypes = [_fp, _ip, _i, ucp]
        L.ref_threshold.restype = _ll
        for name in ("ref_delay_imag", "ref_amdemod"):
            getattr(L, name).argtypes = [_fp, _ip, _i, _fp]
            getattr(L, name).restype = _ll
        L.ref_squelch.argtypes = [_f, _fp, _ip, _i, _fp]
        L.ref_squelch.restype = _ll
        L.ref_ssbdemod.argtypes = [_f, _f, _i, _fp, _ip, _i, _fp]
        L.ref_ssbdemod.restype = _ll
        self._pw = True

    @staticmethod
    def _f32view(x):
        x = np.ascontiguousarray(x)
        return x, _fptr(x.view(np.float32))

    def math(self, op, a, b, block):
        self._decl_pointwise()
        a, pa = self._f32view(a)
        b, pb = self._f32view(b)
        bl = as_blocks(len(a), block)
        y = np.empty(len(a), a.dtype)
        n = self.lib.ref_math(op, int(np.iscomplexobj(a)), pa, pb, _iptr(bl), len(bl), _fptr(y.view(np.float32)))
        assert n == len(a)
        return y

    def layout(self, op, x0, x1, block):
        """Returns out0 (and out1 for op 3). Element types follow the QDSP_LAYOUT_* op."""
        self._decl_pointwise()
        x0, p0 = self._f32view(x0)
        p1 = None
        if x1 is not None:
            x1, p1 = self._f32view(x1)
        n = len(x0)
        bl = as_blocks(n, block)
        out_complex = op in (0, 1, 4, 7)
        y0 = np.empty(n, np.complex64 if out_complex else np.float32)
        y1 = np.empty(n, np.float32)
        m = self.lib.ref_layout(op, p0, p1, _iptr(bl), len(bl), _fptr(y0.view(np.float32)), _fptr(y1))
        assert m == n
        return (y0, y1) if op == 3 else y0

    def volume(self, x, volume, call_set, muted, block):
        self._decl_pointwise()
        x, px = self._f32view(x)
        bl = as_blocks(len(x), block)
        y = np.empty(len(x), x.dtype)
        n = self.lib.ref_volume(int(np.iscomplexobj(x)), volume, int(call_set), int(muted), px, _iptr(bl), len(bl),
                                _fptr(y.view(np.float32)))
        assert n == len(x)
        return y

    def threshold(self, x, block):
        self._decl_pointwise()
        x, px = self._f32view(np.asarray(x, np.float32))
        bl = as_blocks(len(x), block)
        y = np.empty(len(x), np.uint8)
        n = self.lib.ref_threshold(px, _iptr(bl), len(bl), y.ctypes.data_as(C.POINTER(C.c_ubyte)))
        assert n == len(x)
        return y

    def _c2x(self, fn, x, block, out_dtype, *pre):
        self._decl_pointwise()
        x, px = self._cin(x)
        bl = as_blocks(len(x), block)
        y = np.empty(len(x), out_dtype)
        n = fn(*pre, px, _iptr(bl), len(bl), _fptr(y.view(np.float32)))
        assert n == len(x)
        return y

    def delay_imag(self, x, block):
        return self._c2x(self.lib.ref_delay_imag, x, block, np.complex64)

    def amdemod(self, x, block):
        return self._c2x(self.lib.ref_amdemod, x, block, np.float32)

    def squelch(self, level, x, block):
        return self._c2x(self.lib.ref_squelch, x, block, np.complex64, _f(level))

    def ssbdemod(self, fs, bw, mode, x, block):
        return self._c2x(self.lib.ref_ssbdemod, x, block, np.float32, _f(fs), _f(bw), int(mode))



    # -- MMClockRecovery / MSKDemod / PSKDemod ("next" row) ------------------------------------------------
    def interp_taps(self) -> np.ndarray:
        t = np.empty((129, 8), np.float32)
        self.lib.ref_interp_taps.argtypes = [_fp]
        self.lib.ref_interp_taps(_fptr(t))
        return t

    def mm(self, x, omega, gain_omega, mu_gain, omega_rel_limit, block):
        x = np.ascontiguousarray(x)
        cplx = int(np.iscomplexobj(x))
        bl = as_blocks(len(x), block)
        y = np.empty(len(x) + 16, x.dtype)
        oc = np.zeros(len(bl), np.int32)
        self.lib.ref_mm.argtypes = [_i, _f, _f, _f, _f, _fp, _ip, _i, _fp, _ip]
        self.lib.ref_mm.restype = _ll
        n = self.lib.ref_mm(cplx, omega, gain_omega, mu_gain, omega_rel_limit, _fptr(x.view(np.float32)), _iptr(bl), len(bl),
                            _fptr(y.view(np.float32)), _iptr(oc))
        return y[:n].copy(), oc

    def msk_demod(self, fs, dev, baud, x, block):
        x, px = self._cin(x)
        bl = as_blocks(len(x), block)
        y = np.empty(len(x) + 16, np.float32)
        oc = np.zeros(len(bl), np.int32)
        self.lib.ref_msk_demod.argtypes = [_f, _f, _f, _fp, _ip, _i, _fp, _ip]
        self.lib.ref_msk_demod.restype = _ll
        n = self.lib.ref_msk_demod(fs, dev, baud, px, _iptr(bl), len(bl), _fptr(y), _iptr(oc))
        return y[:n].copy(), oc

    def psk_demod(self, order, offset, fs, baud, x, block):
        x, px = self._cin(x)
        bl = as_blocks(len(x), block)
        y = np.empty(len(x) + 16, np.complex64)
        oc = np.zeros(len(bl), np.int32)
        self.lib.ref_psk_demod.argtypes = [_i, _i, _f, _f, _fp, _ip, _i, _fp, _ip]
        self.lib.ref_psk_demod.restype = _ll
        n = self.lib.ref_psk_demod(order, int(offset), fs, baud, px, _iptr(bl), len(bl), _fptr(y.view(np.float32)), _iptr(oc))
        return y[:n].copy(), oc


@lru_cache(maxsize=None)
def ref(variant: str = "generic") -> Ref:
    return Ref(variant)


class Port:
    """numpy-facing wrapper over oracle/liboracle_port.so (oracle/port.c, the C restatement)."""

    def __init__(self):
        path = os.path.join(HERE, "liboracle_port.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-s", "-C", HERE, "port"])
        self.lib = L = C.CDLL(path)
        vp = C.c_void_p
        L.port_blackman_tap_count.argtypes = [_f, _f, _f]
        L.port_blackman_tap_count.restype = _i
        L.port_blackman_taps.argtypes = [_f, _f, _f, _fp, _i, _f]
        L.port_blackman_bandpass_taps.argtypes = [_f, _f, _f, _f, _fp, _i, _f]
        L.port_rrc_taps.argtypes = [_i, _f, _f, _f, _fp]
        L.port_rotator.argtypes = [_fp, _fp, _f, _f, _fp, _fp, _i]
        L.port_rotator_f64.argtypes = [_fp, _fp, _f, _f, _dp, _ll]
        L.port_xlator_phase_delta.argtypes = [_f, _f, _fp, _fp]
        L.port_fir_cf32.argtypes = [_fp, _i, _fp, _ll, _fp]
        L.port_fir_f32.argtypes = [_fp, _i, _fp, _ll, _fp]
        L.port_build_tap_phases.argtypes = [_fp, _i, _i, _fp]
        L.port_build_tap_phases.restype = _i
        L.port_rates_to_ratio.argtypes = [_f, _f, _ip, _ip]
        L.port_resamp_cf32.argtypes = [_fp, _i, _i, _i, _fp, _ip, _i, _fp, _ip]
        L.port_resamp_cf32.restype = _ll
        L.port_resamp_f32.argtypes = [_fp, _i, _i, _i, _fp, _ip, _i, _fp, _ip]
        L.port_resamp_f32.restype = _ll
        L.port_resamp_schedule.argtypes = [_i, _i, _i, _ip, _ip]
        L.port_resamp_schedule.restype = _i
        L.port_power_decim.argtypes = [C.c_uint, _fp, _ip, _i, _fp, _ip]
        L.port_power_decim.restype = _ll
        L.port_fast_arctan2.argtypes = [_f, _f]
        L.port_fast_arctan2.restype = _f
        L.port_fm_phasor_speed.argtypes = [_f, _f]
        L.port_fm_phasor_speed.restype = _f
        L.port_fm_demod.argtypes = [_fp, _ll, _f, _fp, _fp]
        L.port_vfo_design.argtypes = [_f, _f, _f, _fp, _i, _ip, _ip]
        L.port_vfo_design.restype = _i
        L.port_vfo_fm.argtypes = [_f, _f, _f, _f, _f, _i, _fp, _ip, _i, _fp, _ip, _fp]
        L.port_vfo_fm.restype = _ll
        L.port_vfo_fm_window.argtypes = [_f, _f, _f, _f, _f, C.c_double, _fp, _ip, _i, _fp, _ip, _fp]
        L.port_vfo_fm_window.restype = _ll
        L.port_xlator_theta.argtypes = [_f, _f]
        L.port_xlator_theta.restype = C.c_double
        L.port_rotator_checkpoints.argtypes = [_f, _f, _fp, _fp, _ip, _i, _fp]
        L.port_rotator_checkpoints.restype = _ll
        L.port_deemp.argtypes = [_f, _f, _fp, _ll, _fp, _fp, _fp]
        L.port_agc.argtypes = [_f, _f, _fp, _ip, _i, _fp, _fp]
        L.port_complex_agc.argtypes = [_f, _f, _f, _fp, _ll, _fp, _fp]
        L.port_ff_agc_cf32.argtypes = [_fp, _ll, _fp]
        L.port_ff_agc_cf32.restype = _ll
        L.port_costas_coeffs.argtypes = [_f, _fp, _fp]
        L.port_costas.argtypes = [_i, _f, _fp, _ll, _fp, _fp]

    @staticmethod
    def _cin(x):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        return x, _fptr(x.view(np.float32))

    def blackman_tap_count(self, cutoff, tw, fs) -> int:
        return int(self.lib.port_blackman_tap_count(cutoff, tw, fs))

    def blackman_taps(self, cutoff, tw, fs, factor=1.0, count=None) -> np.ndarray:
        n = self.blackman_tap_count(cutoff, tw, fs) if count is None else count
        t = np.empty(n, np.float32)
        self.lib.port_blackman_taps(cutoff, tw, fs, _fptr(t), n, factor)
        return t

    def blackman_bandpass_taps(self, cutoff, tw, offset, fs, factor=1.0) -> np.ndarray:
        n = self.blackman_tap_count(cutoff, tw, fs)
        t = np.empty(n, np.float32)
        self.lib.port_blackman_bandpass_taps(cutoff, tw, offset, fs, _fptr(t), n, factor)
        return t

    def rrc_taps(self, count, fs, baud, alpha) -> np.ndarray:
        t = np.zeros(count | 1, np.float32)
        self.lib.port_rrc_taps(count, fs, baud, alpha, _fptr(t))
        return t

    def vfo_design(self, in_sr, out_sr, bw):
        i, d = _i(), _i()
        n = int(self.lib.port_vfo_design(in_sr, out_sr, bw, None, 0, C.byref(i), C.byref(d)))
        t = np.empty(n, np.float32)
        self.lib.port_vfo_design(in_sr, out_sr, bw, _fptr(t), n, C.byref(i), C.byref(d))
        return t, i.value, d.value

    def rates_to_ratio(self, in_sr, out_sr):
        i, d = _i(), _i()
        self.lib.port_rates_to_ratio(in_sr, out_sr, C.byref(i), C.byref(d))
        return i.value, d.value

    def xlator_phase_delta(self, fs, freq) -> complex:
        re, im = _f(), _f()
        self.lib.port_xlator_phase_delta(fs, freq, C.byref(re), C.byref(im))
        return complex(re.value, im.value)

    def rotator(self, x, inc: complex, phase: complex, block):
        x, px = self._cin(x)
        b = as_blocks(len(x), block)
        y = np.empty(len(x), np.complex64)
        yf = y.view(np.float32)
        pr, pi = _f(np.float32(phase.real)), _f(np.float32(phase.imag))
        off = 0
        for cnt in b:
            self.lib.port_rotator(_fptr(x.view(np.float32)[2 * off:]), _fptr(yf[2 * off:]), np.float32(inc.real),
                                  np.float32(inc.imag), C.byref(pr), C.byref(pi), int(cnt))
            off += int(cnt)
        return y, complex(pr.value, pi.value)

    def rotator_f64(self, x, inc: complex, ang: float = 0.0):
        x, px = self._cin(x)
        y = np.empty(len(x), np.complex64)
        a = C.c_double(ang)
        self.lib.port_rotator_f64(px, _fptr(y.view(np.float32)), np.float32(inc.real), np.float32(inc.imag),
                                  C.byref(a), len(x))
        return y, a.value

    def fir_cf32(self, taps, x):
        x, px = self._cin(x)
        taps = np.ascontiguousarray(taps, np.float32)
        y = np.empty(len(x), np.complex64)
        self.lib.port_fir_cf32(_fptr(taps), len(taps), px, len(x), _fptr(y.view(np.float32)))
        return y

    def fir_f32(self, taps, x):
        x = np.ascontiguousarray(x, np.float32)
        taps = np.ascontiguousarray(taps, np.float32)
        y = np.empty(len(x), np.float32)
        self.lib.port_fir_f32(_fptr(taps), len(taps), _fptr(x), len(x), _fptr(y))
        return y

    def tap_phases(self, taps, interp):
        taps = np.ascontiguousarray(taps, np.float32)
        tpp = (len(taps) + interp - 1) // interp
        ph = np.empty((interp, tpp), np.float32)
        self.lib.port_build_tap_phases(_fptr(taps), len(taps), interp, _fptr(ph))
        return ph

    def resamp_cf32(self, taps, interp, decim, x, block):
        x, px = self._cin(x)
        taps = np.ascontiguousarray(taps, np.float32)
        b = as_blocks(len(x), block)
        cap = len(x) * interp // decim + len(b) + 16
        y = np.empty(cap, np.complex64)
        oc = np.zeros(len(b), np.int32)
        n = self.lib.port_resamp_cf32(_fptr(taps), len(taps), interp, decim, px, _iptr(b), len(b),
                                      _fptr(y.view(np.float32)), _iptr(oc))
        return y[:n].copy(), oc

    def resamp_f32(self, taps, interp, decim, x, block):
        x = np.ascontiguousarray(x, np.float32)
        taps = np.ascontiguousarray(taps, np.float32)
        b = as_blocks(len(x), block)
        cap = len(x) * interp // decim + len(b) + 16
        y = np.empty(cap, np.float32)
        oc = np.zeros(len(b), np.int32)
        n = self.lib.port_resamp_f32(_fptr(taps), len(taps), interp, decim, _fptr(x), _iptr(b), len(b), _fptr(y), _iptr(oc))
        return y[:n].copy(), oc

    def resamp_schedule(self, interp, decim, count):
        n = (count * interp) // decim
        ph = np.empty(n, np.int32)
        ix = np.empty(n, np.int32)
        m = self.lib.port_resamp_schedule(interp, decim, count, _iptr(ph), _iptr(ix))
        assert m == n
        return ph, ix

    def power_decim(self, power, x, block):
        x, px = self._cin(x)
        b = as_blocks(len(x), block)
        y = np.empty(len(x), np.complex64)
        oc = np.zeros(len(b), np.int32)
        n = self.lib.port_power_decim(power, px, _iptr(b), len(b), _fptr(y.view(np.float32)), _iptr(oc))
        return y[:n].copy(), oc

    def stereo_fm(self, fs, dev, x, block):
        x, px = self._cin(x)
        b = as_blocks(len(x), block)
        y = np.empty((len(x), 2), np.float32)
        self.lib.port_stereo_fm.argtypes = [_f, _f, _fp, _ip, _i, _fp]
        self.lib.port_stereo_fm.restype = _ll
        n = self.lib.port_stereo_fm(fs, dev, px, _iptr(b), len(b), _fptr(y))
        assert n == len(x)
        return y

    def fast_arctan2(self, y, x) -> float:
        return float(self.lib.port_fast_arctan2(y, x))

    def fm_phasor_speed(self, fs, dev) -> float:
        return float(self.lib.port_fm_phasor_speed(fs, dev))

    def fm_demod(self, fs, dev, x, phase=0.0):
        x, px = self._cin(x)
        y = np.empty(len(x), np.float32)
        st = _f(phase)
        self.lib.port_fm_demod(px, len(x), self.fm_phasor_speed(fs, dev), C.byref(st), _fptr(y))
        return y

    def vfo_fm(self, offset, in_sr, out_sr, bw, dev, x, block, nco_f64=False, want_iq=False):
        x, px = self._cin(x)
        b = as_blocks(len(x), block)
        i, d = self.rates_to_ratio(in_sr, out_sr)
        cap = len(x) * i // d + len(b) + 16
        a = np.empty(cap, np.float32)
        iq = np.empty(cap, np.complex64)
        oc = np.zeros(len(b), np.int32)
        n = self.lib.port_vfo_fm(offset, in_sr, out_sr, bw, dev, int(nco_f64), px, _iptr(b), len(b), _fptr(a),
                                 _iptr(oc), _fptr(iq.view(np.float32)))
        if want_iq:
            return a[:n].copy(), oc, iq[:n].copy()
        return a[:n].copy(), oc

    def vfo_fm_window(self, offset, in_sr, out_sr, bw, dev, x, block, abs_start):
        """The f64-NCO chain over a window x = stream[abs_start : abs_start + len(x)] (resampler from zero history:
        discard the first ceil(T/D)+1 outputs). Returns (audio, out_counts, iq)."""
        x, px = self._cin(x)
        b = as_blocks(len(x), block)
        i, d = self.rates_to_ratio(in_sr, out_sr)
        cap = len(x) * i // d + len(b) + 16
        a = np.empty(cap, np.float32)
        iq = np.empty(cap, np.complex64)
        oc = np.zeros(len(b), np.int32)
        theta = self.lib.port_xlator_theta(in_sr, -offset)
        start = math.fmod(theta * float(abs_start), 2.0 * math.pi) if abs_start < (1 << 50) else 0.0
        # theta * abs_start in float64 loses ~1e-16 * abs_start rad: split the product to keep it exact to ~1e-12
        hi, lo = divmod(int(abs_start), 1 << 20)
        start = math.fmod(math.fmod(theta * float(1 << 20), 2.0 * math.pi) * hi, 2.0 * math.pi) + theta * lo
        n = self.lib.port_vfo_fm_window(offset, in_sr, out_sr, bw, dev, start, px, _iptr(b), len(b), _fptr(a), _iptr(oc),
                                        _fptr(iq.view(np.float32)))
        return a[:n].copy(), oc, iq[:n].copy()

    def rotator_checkpoints(self, inc: complex, block_sizes, phase: complex = 1 + 0j):
        """Phase state of the float32 recursive rotator at the start of every 512-sample run of every call."""
        b = np.ascontiguousarray(block_sizes, dtype=np.int32)
        total = int(sum((int(c) + 511) // 512 for c in b))
        ck = np.empty(max(total, 1) * 2, np.float32)
        pr, pi = _f(np.float32(phase.real)), _f(np.float32(phase.imag))
        k = self.lib.port_rotator_checkpoints(np.float32(inc.real), np.float32(inc.imag), C.byref(pr), C.byref(pi),
                                              _iptr(b), len(b), _fptr(ck))
        assert k == total
        return ck[: 2 * total].view(np.complex64).copy(), complex(pr.value, pi.value)

    def deemp(self, fs, tau, x):
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, 2)
        y = np.empty_like(x)
        l, r = _f(0.0), _f(0.0)
        self.lib.port_deemp(fs, tau, _fptr(x), len(x), _fptr(y), C.byref(l), C.byref(r))
        return y

    def agc(self, fall_rate, fs, x, block):
        x = np.ascontiguousarray(x, dtype=np.float32)
        b = as_blocks(len(x), block)
        y = np.empty_like(x)
        lv = _f(0.0)
        self.lib.port_agc(fall_rate, fs, _fptr(x), _iptr(b), len(b), _fptr(y), C.byref(lv))
        return y

    def complex_agc(self, set_point, max_gain, rate, x, gain=1.0):
        x, px = self._cin(x)
        y = np.empty(len(x), np.complex64)
        g = _f(gain)
        self.lib.port_complex_agc(set_point, max_gain, rate, px, len(x), _fptr(y.view(np.float32)), C.byref(g))
        return y

    def ff_agc(self, x):
        x, px = self._cin(x)
        y = np.empty(len(x), np.complex64)
        n = self.lib.port_ff_agc_cf32(px, len(x), _fptr(y.view(np.float32)))
        return y[:n].copy()

    def costas_coeffs(self, bw):
        a, b = _f(), _f()
        self.lib.port_costas_coeffs(bw, C.byref(a), C.byref(b))
        return a.value, b.value

    def costas(self, order, bw, x, state=None):
        x, px = self._cin(x)
        y = np.empty(len(x), np.complex64)
        st = np.asarray([0, 0, 1, 0] if state is None else state, dtype=np.float32)
        self.lib.port_costas(order, bw, px, len(x), _fptr(y.view(np.float32)), _fptr(st))
        return y, st


    # -- element-wise / layout / per-block-statistic blocks ("next" rows) ----------------------------------
    def _decl_pointwise(self):
        L = self.lib
        if getattr(self, "_pw", False):
            return
        ucp = C.POINTER(C.c_ubyte)
        L.port_math.argtypes = [_i, _i, _fp, _fp, _fp, _ll]
        L.port_layout.argtypes = [_i, _fp, _fp, _fp, _fp, _ll]
        L.port_volume_level.argtypes = [_f]
        L.port_volume_level.restype = _f
        L.port_volume.argtypes = [_f, _i, _fp, _fp, _ll]
        L.port_threshold.argtypes = [_fp, ucp, _ll]
        L.port_delay_imag.argtypes = [_fp, _fp, _ll, _fp]
        L.port_amdemod.argtypes = [_fp, _ip, _i, _fp]
        L.port_squelch.argtypes = [_f, _fp, _ip, _i, _fp]
        L.port_ssb_phase_delta.argtypes = [_f, _f, _i, _fp, _fp]
        L.port_ssbdemod.argtypes = [_f, _f, _i, _fp, _ip, _i, _fp]
        self._pw = True

    @staticmethod
    def _f32view(x):
        x = np.ascontiguousarray(x)
        return x, _fptr(x.view(np.float32))

    def math(self, op, a, b):
        self._decl_pointwise()
        a, pa = self._f32view(a)
        b, pb = self._f32view(b)
        y = np.empty(len(a), a.dtype)
        self.lib.port_math(op, int(np.iscomplexobj(a)), pa, pb, _fptr(y.view(np.float32)), len(a))
        return y

    def layout(self, op, x0, x1=None):
        self._decl_pointwise()
        x0, p0 = self._f32view(x0)
        p1 = None
        if x1 is not None:
            x1, p1 = self._f32view(x1)
        n = len(x0)
        y0 = np.empty(n, np.complex64 if op in (0, 1, 4, 7) else np.float32)
        y1 = np.empty(n, np.float32)
        self.lib.port_layout(op, p0, p1, _fptr(y0.view(np.float32)), _fptr(y1), n)
        return (y0, y1) if op == 3 else y0

    def volume(self, x, volume, call_set, muted):
        self._decl_pointwise()
        x, px = self._f32view(x)
        level = float(self.lib.port_volume_level(volume)) if call_set else 1.0
        y = np.empty(len(x), x.dtype)
        self.lib.port_volume(level, int(muted), px, _fptr(y.view(np.float32)), x.view(np.float32).size)
        return y

    def threshold(self, x):
        self._decl_pointwise()
        x, px = self._f32view(np.asarray(x, np.float32))
        y = np.empty(len(x), np.uint8)
        self.lib.port_threshold(px, y.ctypes.data_as(C.POINTER(C.c_ubyte)), len(x))
        return y

    def delay_imag(self, x, last_im=0.0):
        self._decl_pointwise()
        x, px = self._cin(x)
        y = np.empty(len(x), np.complex64)
        st = _f(last_im)
        self.lib.port_delay_imag(px, _fptr(y.view(np.float32)), len(x), C.byref(st))
        return y

    def amdemod(self, x, block):
        self._decl_pointwise()
        x, px = self._cin(x)
        bl = as_blocks(len(x), block)
        y = np.empty(len(x), np.float32)
        self.lib.port_amdemod(px, _iptr(bl), len(bl), _fptr(y))
        return y

    def squelch(self, level, x, block):
        self._decl_pointwise()
        x, px = self._cin(x)
        bl = as_blocks(len(x), block)
        y = np.empty(len(x), np.complex64)
        self.lib.port_squelch(level, px, _iptr(bl), len(bl), _fptr(y.view(np.float32)))
        return y

    def ssb_phase_delta(self, fs, bw, mode) -> complex:
        self._decl_pointwise()
        re, im = _f(), _f()
        self.lib.port_ssb_phase_delta(fs, bw, mode, C.byref(re), C.byref(im))
        return complex(re.value, im.value)

    def ssbdemod(self, fs, bw, mode, x, block):
        self._decl_pointwise()
        x, px = self._cin(x)
        bl = as_blocks(len(x), block)
        y = np.empty(len(x), np.float32)
        self.lib.port_ssbdemod(fs, bw, mode, px, _iptr(bl), len(bl), _fptr(y))
        return y

    # -- MMClockRecovery ("next" row); `taps` = the caller's INTERP_TAPS[129][8] -------------------------------
    @staticmethod
    def mm_initial_state(omega) -> np.ndarray:
        st = np.zeros(44, np.float32)
        st[0], st[1] = 0.5, omega   # _mu = 0.5, _dynOmega = _omega (clock_recovery.h:86, 233)
        return st

    def mm(self, x, omega, gain_omega, mu_gain, omega_rel_limit, taps, block, state=None):
        x = np.ascontiguousarray(x)
        cplx = int(np.iscomplexobj(x))
        taps = np.ascontiguousarray(taps, np.float32)
        bl = as_blocks(len(x), block)
        y = np.empty(len(x) + 16, x.dtype)
        oc = np.zeros(len(bl), np.int32)
        st = self.mm_initial_state(np.float32(omega)) if state is None else state
        self.lib.port_mm.argtypes = [_i, _f, _f, _f, _f, _fp, _fp, _ip, _i, _fp, _ip, _fp]
        self.lib.port_mm.restype = _ll
        n = self.lib.port_mm(cplx, omega, gain_omega, mu_gain, omega_rel_limit, _fptr(taps), _fptr(x.view(np.float32)),
                             _iptr(bl), len(bl), _fptr(y.view(np.float32)), _iptr(oc), _fptr(st))
        return y[:n].copy(), oc

@lru_cache(maxsize=None)
def port() -> Port:
    return Port()
