"""CPU gate 2: host-side logic of the product and the C-ABI surface (no compute without a GPU)."""
import ctypes as C

import numpy as np
import pytest


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def test_library_exports_every_declared_symbol(qlib):
    from qdsp_b200 import lib

    declared = lib.header_symbols()
    assert len(declared) > 90
    missing = [s for s in declared if not hasattr(qlib, s)]
    assert not missing, missing
    # and the ctypes table covers the header one-to-one
    assert sorted(lib.SIGNATURES) == declared
    assert qlib.qdsp_abi_version() == 1


def test_no_cpu_fallback_without_device(qlib):
    import torch

    from qdsp_b200 import blocks as B
    from qdsp_b200.lib import QdspError

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(QdspError):
        B.FIR(B.BlackmanWindow(300e3, 4 * 2.4e6 / 127, 2.4e6))
    with pytest.raises(QdspError):
        B.VFOFM(250e3, 2.4e6, 48e3, 48e3, 5e3)


def test_product_tap_designer_bit_exact(qlib, golden):
    from qdsp_b200 import blocks as B

    assert np.array_equal(_bits(B.BlackmanWindow(300e3, 4 * 2.4e6 / 127, 2.4e6).createTaps()), _bits(golden["taps_cfg1"]))
    assert np.array_equal(_bits(B.BlackmanWindow(100e3, 4 * 2.4e6 / 4095, 2.4e6).createTaps()), _bits(golden["taps_cfg3"]))
    assert np.array_equal(_bits(B.BlackmanBandpassWindow(15e3, 4e3, 19e3, 240e3).createTaps()), _bits(golden["taps_bandpass"]))
    assert np.array_equal(_bits(B.RRCTaps(31, 4.0, 1.0, 0.35).createTaps()), _bits(golden["taps_rrc"]))
    for name, (i_sr, o_sr) in {"cfg2": (2.4e6, 48e3), "cfg4": (61.44e6, 48e3), "rational": (250e3, 48e3)}.items():
        i, d = C.c_int(), C.c_int()
        n = qlib.qdsp_vfo_design(i_sr, o_sr, 48e3, None, 0, C.byref(i), C.byref(d))
        t = np.empty(n, np.float32)
        qlib.qdsp_vfo_design(i_sr, o_sr, 48e3, t.ctypes.data_as(C.POINTER(C.c_float)), n, C.byref(i), C.byref(d))
        assert (i.value, d.value) == tuple(golden["id_" + name])
        assert np.array_equal(_bits(t), _bits(golden["taps_" + name]))


def test_host_schedule_bit_exact(qlib, port):
    from qdsp_b200 import blocks as B

    for (i, d, n) in [(24, 125, 1000), (1, 50, 819200), (147, 160, 4410), (4, 1, 100), (1, 4, 9001), (1, 1280, 819200)]:
        ph, ix = B.resamp_schedule(i, d, n)
        rph, rix = port.resamp_schedule(i, d, n)
        assert np.array_equal(ph, rph) and np.array_equal(ix, rix)
    assert B.rates_to_ratio(2.4e6, 48e3) == (1, 50)
    assert B.rates_to_ratio(61.44e6, 48e3) == (1, 1280)
    assert B.rates_to_ratio(250e3, 48e3) == (24, 125)
    assert B.rates_to_ratio(48e3, 44.1e3) == port.rates_to_ratio(48e3, 44.1e3)


def test_synth_generators_are_random_access():
    from qdsp_b200 import synth

    a = synth.uniform_cf32(3, 0, 1000)
    b = synth.uniform_cf32(3, 400, 100)
    assert np.array_equal(a[400:500], b)
    assert a.real.min() >= -1 and a.real.max() < 1
    f = synth.cfg2_input(0, 2000)
    g = synth.cfg2_input(1500, 500)
    assert np.array_equal(f[1500:], g)
