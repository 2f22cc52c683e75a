"""Tier-U ops against INDEPENDENT authorities.

FunDSP's source is not available here, so the product's coefficient derivations (csrc/coefs.h) and the oracle's restatement
(oracle/qo_units.h) were written by the same hand: a common-mode error would be invisible to product-vs-oracle parity.  These
tests check BOTH against third-party closed forms — scipy's bilinear transform of the published analog prototypes (Simper's
SVF, the RBJ cookbook shelves / peaking EQ, Butterworth), textbook one-pole sections, the equal-power pan law, numpy's
sine — and put `white()` through distribution / correlation tests.  What stays unpinned after this file: FunDSP's choice of
hash constants (noise seeds, initial phases), its envelope jitter and its wavetable design — data, not algorithms.

CPU only: the product side is read through the lowering's own device-parameter table (qg_net_device_params), the oracle
side through impulse responses."""
import numpy as np
import pytest
from scipy import signal, stats

import quartz_b200 as qb
from quartz_b200 import Net
from tests.graphs import build, pipe
from tests.oracle_ffi import ONet

SR = 48000.0


def product_params(op):
    return Net.str_to_net(op).set_sample_rate(SR).device_params().astype(np.float64)


def oracle_impulse_response(op, n=1 << 15):
    return build({"op": "sr()", "net": pipe("impulse()", op), "n": SR}, ONet).render(n)[:, 0].astype(np.float64)


def svf_response_from_coefs(c, w):
    """frequency response of Simper's trapezoidal SVF (ic1, ic2 state; v1, v2 taps) from its six coefficients, evaluated by
    running the recurrence in f64 on an impulse — no formula shared with coefs.h beyond the tick itself"""
    a1, a2, a3, m0, m1, m2 = c
    n = 1 << 15
    ic1 = ic2 = 0.0
    h = np.empty(n)
    for i in range(n):
        v0 = 1.0 if i == 0 else 0.0
        v3 = v0 - ic2
        v1 = a1 * ic1 + a2 * v3
        v2 = ic2 + a2 * ic1 + a3 * v3
        ic1 = 2.0 * v1 - ic1
        ic2 = 2.0 * v2 - ic2
        h[i] = m0 * v0 + m1 * v1 + m2 * v2
    return np.fft.rfft(h)[np.round(w / (2 * np.pi) * n).astype(int)]


def analog_prototype(mode, q, gain):
    """(b, a) in s, cutoff normalised to 1 rad/s — Cytomic "SvfLinearTrapOptimised2" / RBJ Audio-EQ-Cookbook prototypes"""
    k = 1.0 / q
    A = np.sqrt(gain)
    den = [1.0, k, 1.0]
    return {
        "lowpass": ([1.0], den), "highpass": ([1.0, 0.0, 0.0], den), "bandpass": ([1.0, 0.0], den),
        "notch": ([1.0, 0.0, 1.0], den), "peak": ([1.0, 0.0, -1.0], den), "allpass": ([1.0, -k, 1.0], den),
        "bell": ([1.0, A / q, 1.0], [1.0, 1.0 / (A * q), 1.0]),
        "lowshelf": ([A, A * np.sqrt(A) / q, A * A], [A, np.sqrt(A) / q, 1.0]),
        "highshelf": ([A * A, A * np.sqrt(A) / q, A], [1.0, np.sqrt(A) / q, A]),
    }[mode]


def digital_reference(mode, fc, q, gain, w):
    """scipy's bilinear transform of the analog prototype, prewarped so that fc maps onto itself"""
    b, a = analog_prototype(mode, q, gain)
    w0 = 2.0 * SR * np.tan(np.pi * fc / SR)
    nb, na = len(b) - 1, len(a) - 1
    b = [x / w0 ** (nb - i) for i, x in enumerate(b)]
    a = [x / w0 ** (na - i) for i, x in enumerate(a)]
    bz, az = signal.bilinear(b, a, fs=SR)
    return signal.freqz(bz, az, worN=w)[1]


SVF_CASES = [(m, fc, q, g) for m in ("lowpass", "highpass", "bandpass", "notch", "peak", "allpass") for fc, q in ((300.0, 0.7), (1000.0, 2.0), (9000.0, 8.0))
             for g in (1.0,)] + [(m, fc, q, g) for m in ("bell", "lowshelf", "highshelf") for fc, q, g in ((500.0, 1.0, 4.0), (3000.0, 2.0, 0.25))]


@pytest.mark.parametrize("mode,fc,q,gain", SVF_CASES)
def test_svf_modes_match_the_bilinear_transform_of_their_published_prototypes(mode, fc, q, gain):
    op = f"{mode}({fc},{q},{gain})" if mode in ("bell", "lowshelf", "highshelf") else f"{mode}({fc},{q})"
    n = 1 << 15
    w = 2 * np.pi * np.arange(8, n // 2 - 8, 37) / n          # bin-centred frequencies
    ref = digital_reference(mode, fc, q, gain, w)
    scale = max(1.0, np.abs(ref).max())
    got_p = svf_response_from_coefs(product_params(op), w)
    assert np.abs(got_p - ref).max() <= 2e-5 * scale, ("product coefficients", np.abs(got_p - ref).max())
    got_o = np.fft.rfft(oracle_impulse_response(op, n))[np.round(w / (2 * np.pi) * n).astype(int)]
    assert np.abs(got_o - ref).max() <= 2e-4 * scale, ("oracle impulse response", np.abs(got_o - ref).max())


@pytest.mark.parametrize("fc", [200.0, 3000.0, 15000.0])
def test_butterpass_is_scipys_second_order_butterworth(fc):
    b, a = signal.butter(2, fc, fs=SR)
    p = product_params(f"butterpass({fc})")                   # a1 a2 b0 b1 b2
    assert np.allclose(p, [a[1], a[2], b[0], b[1], b[2]], rtol=2e-5, atol=2e-7), (p, a, b)
    h = oracle_impulse_response(f"butterpass({fc})", 4096)
    ref = signal.lfilter(b, a, np.r_[1.0, np.zeros(4095)])
    assert np.abs(h - ref).max() <= 1e-5


@pytest.mark.parametrize("fc,bw", [(440.0, 20.0), (2500.0, 300.0)])
def test_resonator_poles_zeros_and_bandwidth_independent_gain(fc, bw):
    """two-pole resonator with zeros at z = +-1 (Smith & Angell's constant-gain form): pole radius <-> bandwidth, pole angle
    <-> centre frequency, and an overall (white-noise power) gain that does not depend on the bandwidth"""
    a1, a2, b0, b1, b2 = product_params(f"resonator({fc},{bw})")
    poles = np.roots([1.0, a1, a2])
    assert np.allclose(np.abs(poles), np.exp(-np.pi * bw / SR), rtol=1e-5)            # radius <-> bandwidth
    assert np.allclose(np.abs(np.angle(poles)), 2 * np.pi * fc / SR, rtol=2e-4)       # angle <-> centre frequency
    assert b1 == 0.0 and b2 == -b0                                                     # zeros at z = +1 and z = -1
    w, h = signal.freqz([b0, b1, b2], [1.0, a1, a2], worN=1 << 16)
    assert abs(w[np.abs(h).argmax()] - 2 * np.pi * fc / SR) < 2 * np.pi * bw / SR / 4
    power = []
    for k in (0.5, 1.0, 2.0, 4.0):
        p = product_params(f"resonator({fc},{bw * k})")
        hk = signal.lfilter(p[2:], [1.0, p[0], p[1]], np.r_[1.0, np.zeros((1 << 17) - 1)])
        power.append(float(np.sum(hk * hk)))
    assert max(power) / min(power) < 1.02, power
    ho = oracle_impulse_response(f"resonator({fc},{bw})", 8192)
    assert np.abs(ho - signal.lfilter([b0, b1, b2], [1.0, a1, a2], np.r_[1.0, np.zeros(8191)])).max() <= 1e-5


@pytest.mark.parametrize("fc", [50.0, 1000.0, 8000.0])
def test_one_pole_sections_are_the_textbook_ones(fc):
    c = np.exp(-2 * np.pi * fc / SR)                          # impulse-invariant one-pole
    x = np.r_[1.0, np.zeros(4095)]
    assert np.allclose(product_params(f"lowpole({fc})"), [c], rtol=1e-6)
    assert np.abs(oracle_impulse_response(f"lowpole({fc})", 4096) - signal.lfilter([1 - c], [1, -c], x)).max() <= 1e-6
    assert np.allclose(product_params(f"highpole({fc})"), [c], rtol=1e-6)
    assert np.abs(oracle_impulse_response(f"highpole({fc})", 4096) - signal.lfilter([c, -c], [1, -c], x)).max() <= 1e-6
    d = 1 - 2 * np.pi * fc / SR                               # DC blocker: zero at z = 1, pole just inside
    assert np.allclose(product_params(f"dcblock({fc})"), [d], rtol=1e-6)
    assert np.abs(oracle_impulse_response(f"dcblock({fc})", 4096) - signal.lfilter([1, -1], [1, -d], x)).max() <= 1e-6


@pytest.mark.parametrize("delay", [0.1, 0.5, 0.9])
def test_allpole_is_a_first_order_allpass_with_the_requested_dc_delay(delay):
    eta = (1 - delay) / (1 + delay)                           # Thiran / first-order allpass interpolator
    assert np.allclose(product_params(f"allpole({delay})"), [eta], rtol=1e-6)
    h = oracle_impulse_response(f"allpole({delay})", 4096)
    H = np.fft.rfft(h)
    assert np.abs(np.abs(H) - 1.0).max() < 1e-5               # allpass
    gd = -np.diff(np.unwrap(np.angle(H)))[:4] / (2 * np.pi / 4096)
    assert np.allclose(gd, delay, atol=2e-3)                  # group delay at DC = `delay` samples


def test_pan_is_the_equal_power_law():
    for pan in (-1.0, -0.5, 0.0, 0.3, 1.0):
        l, r = product_params(f"pan({pan})")
        assert abs(l * l + r * r - 1.0) < 1e-6
        assert np.allclose([l, r], [np.cos((pan + 1) * np.pi / 4), np.sin((pan + 1) * np.pi / 4)], atol=1e-6)
        y = build({"op": "sr()", "net": pipe("dc(1)", f"pan({pan})"), "n": SR}, ONet).render(4)
        assert np.allclose(y[0], [l, r], atol=1e-6)


def test_white_noise_is_uniform_and_uncorrelated():
    n = 1 << 18
    x = build(pipe("white()"), ONet).set_salt(1).render(n)[:, 0].astype(np.float64)
    y = build(pipe("white()"), ONet).set_salt(2).render(n)[:, 0].astype(np.float64)
    assert -1.0 <= x.min() and x.max() <= 1.0
    assert abs(x.mean()) < 4 / np.sqrt(3 * n)                 # 4 sigma of the mean of U[-1, 1]
    assert abs(x.var() - 1 / 3) < 0.005
    assert stats.kstest(x, stats.uniform(loc=-1, scale=2).cdf).pvalue > 1e-3
    for lag in (1, 2, 3, 7, 64, 1024):
        assert abs(np.corrcoef(x[:-lag], x[lag:])[0, 1]) < 5 / np.sqrt(n), lag
    assert abs(np.corrcoef(x, y)[0, 1]) < 5 / np.sqrt(n)      # different seeds: independent streams
    spec = np.abs(np.fft.rfft(x.reshape(256, -1), axis=1)) ** 2
    band = spec.mean(axis=0)[1:].reshape(8, -1).mean(axis=1)
    assert band.max() / band.min() < 1.1                      # flat spectrum


@pytest.mark.parametrize("hz", [55.0, 440.0, 9001.5])
def test_sine_is_a_sine_of_the_requested_frequency(hz):
    n = 48000
    x = build({"op": "sr()", "net": pipe(f"sine({hz})"), "n": SR}, ONet).render(n)[:, 0].astype(np.float64)
    ph0 = np.arctan2(x[0], (x[1] - x[0] * np.cos(2 * np.pi * hz / SR)) / np.sin(2 * np.pi * hz / SR))
    ref = np.sin(2 * np.pi * hz * np.arange(n) / SR + ph0)
    # the f32 phase accumulator rounds its increment (< 2^-25 cycles per sample) and its sum: milliradians after one second
    assert np.abs(x - ref).max() < 2 * np.pi * n * 2.0 ** -25 * 0.7


def test_semitone_ratio_and_db_conversions_are_the_usual_definitions():
    xs = np.float32([-12, -1, 0, 0.5, 7, 12, 24])
    for x in xs:
        y = build(pipe(f"dc({float(x)!r})", "semitone_ratio()"), ONet).render(1)[0, 0]
        assert np.isclose(y, 2.0 ** (float(x) / 12), rtol=1e-6)
        y = build(pipe(f"dc({float(x)!r})", "db_amp()"), ONet).render(1)[0, 0]
        assert np.isclose(y, 10.0 ** (float(x) / 20), rtol=1e-6)


def _spectrum(op, n=48000):
    x = build({"op": "sr()", "net": pipe(op), "n": SR}, ONet).render(n)[:, 0].astype(np.float64)
    return x, np.abs(np.fft.rfft(x * np.hanning(n))) / (n / 4), np.fft.rfftfreq(n, 1 / SR)


@pytest.mark.parametrize("op,law", [("saw", lambda k: 1.0 / k), ("square", lambda k: (k % 2) / k), ("triangle", lambda k: (k % 2) / k ** 2)])
def test_wavetable_oscillators_have_the_fourier_series_of_their_waveform(op, law):
    """band-limited saw / square / triangle: harmonic k of the ideal waveform has relative amplitude 1/k, 1/k (odd), 1/k^2
    (odd); nothing between the harmonics (no aliasing), nothing above the table's band limit"""
    f0 = 441.0
    x, X, f = _spectrum(f"{op}({f0})")
    h = np.array([X[int(round(k * f0))] for k in range(1, 21)])
    want = np.array([law(k) for k in range(1, 21)], dtype=np.float64)
    assert np.allclose(h / h[0], want, atol=0.01), (h / h[0], want)
    between = np.array([X[int(round((k + 0.5) * f0))] for k in range(1, 40)])
    assert between.max() < 1e-3 * h[0]                         # -60 dB: no aliased partials
    assert 0.9 < np.abs(x).max() <= 1.0 + 1e-6                 # normalised to full scale


@pytest.mark.parametrize("op,slope", [("pink()", -3.0), ("brown()", -6.0)])
def test_coloured_noises_have_their_spectral_slope(op, slope):
    """pink: -3 dB per octave, brown: -6 dB per octave (power, octave bands between 200 Hz and 12.8 kHz)"""
    x, X, f = _spectrum(op, 1 << 17)
    P = X ** 2
    edges = [200, 400, 800, 1600, 3200, 6400, 12800]
    db = np.array([10 * np.log10(P[(f >= a) & (f < b)].mean()) for a, b in zip(edges[:-1], edges[1:])])
    fit = np.polyfit(np.arange(len(db)), db, 1)[0]
    assert abs(fit - slope) < 0.6, (fit, db)
    assert abs(x.mean()) < 0.05 and 0.05 < x.std() < 1.0
