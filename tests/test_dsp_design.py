"""Host-side filter design (signalops.dspjl, the DSP.jl stand-in the GPU plan gets its
coefficients from) against scipy.signal and against the oracle's independent
restatement (oracle/dspjl_ref.py)."""
from fractions import Fraction

import numpy as np
import pytest
from scipy import signal as sps

from oracle import dspjl_ref as D
from signalops import dspjl as H

DESIGNS = [("Lowpass", (4000.0,), 48000.0), ("Highpass", (8.0,), 100.0), ("Bandpass", (500.0, 4000.0), 96000.0),
           ("Bandstop", (500.0, 2000.0), 44100.0), ("Lowpass", (3.0,), 10.0), ("Highpass", (2.0,), 48000.0)]
PROTOS = [("butterworth", 5), ("butterworth", 8), ("butterworth", 1), ("chebyshev1", 5, 1.0), ("chebyshev1", 4, 0.5)]


def host_design(kind, bounds, fs, spec):
    proto = H.Butterworth(spec[1]) if spec[0] == "butterworth" else H.Chebyshev1(spec[1], spec[2])
    return H.digitalfilter(getattr(H, kind)(*bounds, fs=fs), proto)


@pytest.mark.parametrize("kind,bounds,fs", DESIGNS)
@pytest.mark.parametrize("spec", PROTOS)
def test_zpk_matches_scipy(kind, bounds, fs, spec):
    z = host_design(kind, bounds, fs, spec)
    zs, ps, ks = D.design_zpk(kind, list(bounds), fs, spec)
    assert len(z.p) == len(ps) and len(z.z) == len(zs)
    assert abs(z.k - ks) <= 1e-12 * abs(ks)
    assert np.max(np.abs(np.sort_complex(np.array(z.p)) - np.sort_complex(ps))) < 1e-11
    assert np.max(np.abs(np.sort_complex(np.array(z.z)) - np.sort_complex(zs))) < 1e-6   # repeated roots at +-1


@pytest.mark.parametrize("kind,bounds,fs", DESIGNS)
@pytest.mark.parametrize("spec", PROTOS)
def test_sos_filtering_agrees_three_ways(kind, bounds, fs, spec):
    x = np.random.default_rng(0).standard_normal(3000)
    sos_h = H.zpk_to_sos(host_design(kind, bounds, fs, spec))
    zs, ps, ks = D.design_zpk(kind, list(bounds), fs, spec)
    coef_o, g_o = D.zpk2sos_dspjl(zs, ps, ks)
    y_o = D.sos_filt(x, coef_o, g_o, np.zeros((coef_o.shape[0], 2)))
    ch = sos_h.coef_table()
    y_h = D.sos_filt(x, ch, sos_h.g, np.zeros((ch.shape[0], 2)))
    y_s = sps.sosfilt(sps.zpk2sos(zs, ps, ks), x)
    scale = np.sqrt(np.mean(y_s ** 2)) + 1e-300
    assert ch.shape == coef_o.shape
    assert np.all(ch[:, 0] == 1.0)                    # monic numerators; the gain is kept aside
    assert np.max(np.abs(y_h - y_o)) < 1e-9 * scale
    assert np.max(np.abs(y_o - y_s)) < 1e-9 * scale


def test_sos_filt_state_streams():
    x = np.random.default_rng(1).standard_normal(1000)
    coef, g = D.zpk2sos_dspjl(*D.design_zpk("Lowpass", [3.0], 10.0, ("butterworth", 5)))
    whole = D.sos_filt(x, coef, g, np.zeros((coef.shape[0], 2)))
    st = np.zeros((coef.shape[0], 2))
    parts = np.concatenate([D.sos_filt(x[i:i + 37], coef, g, st) for i in range(0, 1000, 37)])
    assert np.array_equal(whole, parts)


@pytest.mark.parametrize("ratio", [48000 / 44100, 44100 / 48000, 0.75, np.pi, Fraction(2, 1), Fraction(1, 2),
                                   Fraction(3, 2), Fraction(2, 3), Fraction(3, 1), Fraction(1, 3)])
def test_resample_filter_matches_independent_design(ratio):
    h = H.resample_filter(ratio)
    ho, nphi = D.resample_filter(ratio)
    assert len(h) == len(ho) and len(h) % 2 == 1
    assert np.max(np.abs(h - ho)) < 1e-12
    f = H.FIRFilter(h, ratio)
    f.setphase(f.timedelay())
    r = D.Resampler(ratio)
    assert f.input_deficit == r.st.input_deficit and f.tapsper == r.st.taps_per_phi
    if f.kind == "arbitrary":
        assert f.phi_acc == r.st.phi_acc
        assert np.allclose(f.pfb, np.asarray(r._pfb), rtol=0, atol=1e-12) and np.allclose(f.dpfb, np.asarray(r._dpfb), rtol=0, atol=1e-12)
    elif f.kind in ("rational", "interpolator"):
        assert f.phi_idx == r.st.phi_idx


def test_known_filter_sizes():
    """44.1->48 k has 1185 taps / 38 per phase; 48->44.1 k 1281 / 41; 1/2 has 75."""
    assert len(H.resample_filter(48000 / 44100)) == 1185
    assert H.FIRFilter(H.resample_filter(48000 / 44100), 48000 / 44100).tapsper == 38
    assert len(H.resample_filter(44100 / 48000)) == 1281
    assert H.FIRFilter(H.resample_filter(44100 / 48000), 44100 / 48000).tapsper == 41
    assert len(H.resample_filter(Fraction(1, 2))) == 75


def test_maybe_rationalize():
    from signalops.graph import maybe_rationalize
    assert maybe_rationalize(2.0) == Fraction(2) and maybe_rationalize(2 / 3) == Fraction(2, 3)
    assert maybe_rationalize(0.75) == 0.75 and maybe_rationalize(48000 / 44100) == 48000 / 44100
