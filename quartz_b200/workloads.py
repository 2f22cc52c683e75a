"""Synthetic workloads = BASELINE.json `configs`, with the seeded per-voice tables of SURVEY.md section 8(d).

Every workload is described backend-neutrally (graph expression + per-voice op-string parameters + salts) so the
same description drives the GPU bank (bench.py, tests) and the CPU oracle (parity tests, cpu_baseline)."""
import numpy as np

FS = 48000.0
SEED = 0x51574152545A   # "QWARTZ"-ish constant from SURVEY.md section 8(d)


def splitmix64(x):
    x = (np.asarray(x, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15))
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def uniform01(voice, stream):
    """counter-based U[0,1) per (voice, stream)"""
    with np.errstate(over="ignore"):
        h = splitmix64(np.uint64(SEED) ^ (np.asarray(voice, dtype=np.uint64) * np.uint64(0x100000001B3) + np.uint64(stream)))
    return (h >> np.uint64(11)).astype(np.float64) / float(1 << 53)


def _loguniform(u, lo, hi):
    return np.exp(np.log(lo) + u * (np.log(hi) - np.log(lo)))


def _L(op):
    return {"op": op}


def _pipe(*xs):
    return {"op": ">>", "n": 0.0, "inputs": [_L(x) if isinstance(x, str) else x for x in xs]}


def _sr(g):
    return {"op": "sr()", "net": g, "n": FS}


class Workload:
    """name, template graph expression, per-voice raw matrix [V, R], salts [V], samples T, group size G"""

    def __init__(self, name, expr, raw, salts, T, group, voice_expr, bytes_per_unit, bound, note, flops_per_unit=0.0):
        self.name, self.expr, self.raw, self.salts, self.T, self.group = name, expr, raw, salts, T, group
        self.voice_expr = voice_expr          # f(v) -> graph expression of voice v with its own constants
        self.bytes_per_unit = bytes_per_unit  # ALGORITHMIC HBM bytes per voice-sample
        self.flops_per_unit = flops_per_unit  # ALGORITHMIC f32 flops per voice-sample (SURVEY.md 8d / DESIGN.md)
        self.bound = bound
        self.note = note
        self.V = len(salts)


def salts_for(voices):
    with np.errstate(over="ignore"):
        s = splitmix64(np.asarray(voices, dtype=np.uint64) ^ np.uint64(SEED))
    return np.where(s == 0, np.uint64(1), s)


def c1_hello(T=480000):
    """hello_440: sine(440) mono, 10 s at 48 kHz (configs[0])"""
    expr = _sr(_L("sine(440)"))
    return Workload("c1_hello_440", expr, np.zeros((1, 1), np.float32) + 440.0, np.zeros(1, np.uint64), T, 1,
                    lambda v: expr, 4.0, "latency", "1 voice; parity + wall time only")


def c2_lowpass_bank(V=4096, T=2880000, v0=0):
    """noise -> lowpass(hz_v, q_v), per-voice outputs kept (configs[1])"""
    voices = np.arange(v0, v0 + V)
    hz = _loguniform(uniform01(voices, 1), 50.0, 12000.0).astype(np.float32)
    q = (0.5 + uniform01(voices, 2) * 7.5).astype(np.float32)
    raw = np.stack([hz, q], axis=1)
    expr = _sr(_pipe("white()", "lowpass(1000,1)"))
    return Workload("c2_noise_lowpass_bank", expr, raw, salts_for(voices), T, 1,
                    lambda v: _sr(_pipe("white()", f"lowpass({float(hz[v])!r},{float(q[v])!r})")),
                    4.0, "hbm", "4 B written per voice-sample", flops_per_unit=14.0)


def c2_lti_bank(kind="butterpass", V=4096, T=2880000, v0=0):
    """noise -> butterpass(hz_v) / lowpole(hz_v): configs[1]'s shape with the other linear recurrences K2 serves
    (not a BASELINE config: used to time the biquad / one-pole instantiations of the scan kernel)"""
    voices = np.arange(v0, v0 + V)
    hz = _loguniform(uniform01(voices, 1), 500.0 if kind == "butterpass" else 20.0, 12000.0).astype(np.float32)
    expr = _sr(_pipe("white()", f"{kind}(1000)"))
    return Workload(f"c2_noise_{kind}_bank", expr, hz[:, None], salts_for(voices), T, 1,
                    lambda v: _sr(_pipe("white()", f"{kind}({float(hz[v])!r})")), 4.0, "hbm", "4 B written per voice-sample",
                    flops_per_unit=10.0)


def c3_polysynth(V=65536, T=480000, G=32, v0=0, osc="sine"):
    """sine(f_v) -> lowpass(hz_v, q_v) -> * ar(a_v, 1, r_v, 4) -> groups of G voices, scaled 1/G (configs[2]);
    osc = saw / square / triangle / soft_saw gives the same voice with a band-limited wavetable oscillator"""
    voices = np.arange(v0, v0 + V)
    f = _loguniform(uniform01(voices, 3), 55.0, 3520.0).astype(np.float32)
    hz = _loguniform(uniform01(voices, 4), 200.0, 8000.0).astype(np.float32)
    q = (0.5 + uniform01(voices, 5) * 3.5).astype(np.float32)
    a = (0.005 + uniform01(voices, 6) * 0.495).astype(np.float32)
    r = (0.1 + uniform01(voices, 7) * 3.9).astype(np.float32)
    one, four = np.ones(V, np.float32), np.full(V, 4.0, np.float32)
    raw = np.stack([f, hz, q, a, one, r, four], axis=1)

    def voice(v):
        o = _pipe(f"{osc}({float(f[v])!r})", f"lowpass({float(hz[v])!r},{float(q[v])!r})")
        return _sr({"op": "*", "n": 0.0, "inputs": [o, _L(f"ar({float(a[v])!r},1,{float(r[v])!r},4)")]})

    expr = _sr({"op": "*", "n": 0.0, "inputs": [_pipe(f"{osc}(440)", "lowpass(1000,1)"), _L("ar(0.01,1,0.5,4)")]})
    return Workload("c3_polysynth_65536" if osc == "sine" else f"c3_polysynth_{osc}_65536", expr, raw, salts_for(voices), T, G, voice, 4.0 / G, "fp32",
                    "osc->lowpass->envelope, mixed in groups of 32; 32 flop per voice-sample (phase 4, sin 12, SVF 12, "
                    "envelope 2, gain 1, mix 1)", flops_per_unit=32.0)


def hann(n):
    i = np.arange(n, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * i / n)).astype(np.float32)


def spectral_graph(N, J, thr, window, source="white()"):
    """One channel of the spectral-gate patch (structure of assets/spectral-gate) with N-point
    transforms and J overlapping instances (hop N/J): noise input x window -> J x [rfft -> gate -> ifft] -> real part x
    window -> join(J)."""
    hop = N // J
    delays = [j * hop for j in range(J)]
    starts = [(N - d) % N for d in delays]

    def stack(xs):
        return {"op": "|", "n": 0.0, "inputs": xs}

    def win():
        taps = [_L("pass()")] + [_L(f"delay({float(np.float32(d / FS))!r})") for d in delays[1:]]
        return _pipe({"op": "wave()", "arr": [float(x) for x in window]}, f"split({J})", stack(taps))

    gate = _pipe("pol()", stack([_pipe({"op": "^", "n": 0.0, "inputs": [_L(f">({thr!r})"), _L("pass()")]},
                                       {"op": "*", "n": 0.0, "inputs": [_L("pass()"), _L("pass()")]}), _L("pass()")]), "car()")
    chains = [_pipe(f"rfft({N},{s})", gate, f"ifft({N},{s})") for s in starts]
    xw = {"op": "*", "n": 0.0, "inputs": [_pipe(source, f"split({J})"), win()]}
    keep_real = "chan(" + ",".join(["1", "0"] * J) + ")"
    syn = _pipe(xw, stack(chains), keep_real)
    return _sr(_pipe({"op": "*", "n": 0.0, "inputs": [syn, win()]}, f"join({J})"))


def c4_spectral(V=1024, T=1440000, N=2048, J=4, thr=16.0, v0=0):
    """spectral gate: N-pt STFT, hop N/J, V channels (configs[3])"""
    voices = np.arange(v0, v0 + V)
    expr = spectral_graph(N, J, float(thr), hann(N))
    # bytes: 4 B written per channel-sample (input is generated on chip); flops ~ 346 per channel-sample
    lg = int(np.log2(N))
    flops = J * (2.5 * N * lg + 5.0 * N * lg) / N + 16.0   # real FFT + complex inverse per instance and sample, windowing
    return Workload(f"c4_spectral_gate_{N}", expr, None, salts_for(voices), T, 1, lambda v: expr, 4.0, "fp32",
                    f"{J} x [rfft({N}) -> gate -> ifft({N})], hop {N // J}; {flops:.0f} flop per channel-sample for the transforms",
                    flops_per_unit=flops)


MINOR = [0.0, 2.0, 3.0, 5.0, 7.0, 8.0, 10.0, 12.0]


def c5_mixed(V=1048576, T=96000, G=32, v0=0):
    """1M-voice mixed graph (configs[4]): four voice archetypes, V/4 voices each, group-mixed G=32, T = 2 s.
    Returns a LIST of workloads (one bank per archetype — a bank shares one tape)."""
    n = V // 4
    out = []
    # A: pitch-quantised ramp oscillator (in-tree nodes only: ramp, quantize, semitone_ratio)
    va = np.arange(v0, v0 + n)
    rate = (0.5 + uniform01(va, 11) * 7.5).astype(np.float32)
    base = _loguniform(uniform01(va, 12), 55.0, 440.0).astype(np.float32)

    def a_expr(r, b):
        pitch = _pipe(f"dc({r!r})", "ramp()", "mul(24)", {"op": "quantize()", "arr": MINOR}, "semitone_ratio()", f"mul({b!r})")
        return _sr(_pipe(pitch, "ramp()", "mul(TAU)", "sin()"))
    out.append(Workload("c5a_quantised_osc", a_expr(2.0, 110.0), np.stack([rate, np.full(n, 24, np.float32), np.full(n, 12, np.float32),
                        base, np.full(n, 6.2831855, np.float32)], axis=1), salts_for(va), T, G,
                        lambda v: a_expr(float(rate[v]), float(base[v])), 4.0 / G, "fp32",
                        "ramp osc + quantize + semitone_ratio; ~56 flop per voice-sample (2 ramps 6, quantize: floor/div/8-entry scan 22, "
                        "exp2 12, sin 12, 4 muls)", flops_per_unit=56.0))
    # B: noise into a shift register clocked by ramp() >> <(0.5) >> rise(), 8 taps averaged
    vb = np.arange(v0 + n, v0 + 2 * n)
    clk = _loguniform(uniform01(vb, 13), 20.0, 2000.0).astype(np.float32)

    def b_expr(c):
        return _sr(_pipe({"op": "|", "n": 0.0, "inputs": [_L("white()"), _pipe(f"dc({c!r})", "ramp()", "<(0.5)", "rise()")]},
                         "shift_reg()", "join(8)"))
    out.append(Workload("c5b_shift_reg", b_expr(100.0), np.stack([clk, np.full(n, 0.5, np.float32)], axis=1), salts_for(vb), T, G,
                        lambda v: b_expr(float(clk[v])), 4.0 / G, "fp32",
                        "white -> shift_reg clocked by a ramp edge detector; ~30 op per voice-sample (hash 9, ramp 3, compare + rise 3, "
                        "8-tap shift 8, mean 8)", flops_per_unit=30.0))
    # C: one-pole feedback  y = g (x + y[n-1])
    vc = np.arange(v0 + 2 * n, v0 + 3 * n)
    g = (0.5 + uniform01(vc, 14) * 0.49).astype(np.float32)

    def c_expr(gg):
        return _sr(_pipe("white()", {"op": "feedback()", "net": _L(f"mul({gg!r})"), "delay": None}))
    out.append(Workload("c5c_feedback", c_expr(0.9), np.stack([np.zeros(n, np.float32), g], axis=1), salts_for(vc), T, G,
                        lambda v: c_expr(float(g[v])), 4.0 / G, "fp32",
                        "1-sample feedback: the held sample lives in the state region; ~11 op per voice-sample (hash 9, add, mul)",
                        flops_per_unit=11.0))
    # D: delay(1024 samples) + lowpole
    vd = np.arange(v0 + 3 * n, v0 + 4 * n)
    hz = _loguniform(uniform01(vd, 15), 100.0, 8000.0).astype(np.float32)
    dt = float(np.float32(1024.0 / FS))

    def d_expr(h):
        return _sr(_pipe("white()", f"delay({dt!r})", f"lowpole({h!r})"))
    out.append(Workload("c5d_delay_lowpole", d_expr(1000.0), np.stack([np.full(n, dt, np.float32), hz], axis=1), salts_for(vd), T, G,
                        lambda v: d_expr(float(hz[v])), 8.0 + 4.0 / G, "hbm", "1024-sample delay ring in HBM: 8 B state traffic per voice-sample; "
                        "~12 op (hash 9, lowpole 3)", flops_per_unit=12.0))
    return out


WORKLOADS = {"c3_saw": lambda **kw: c3_polysynth(osc="saw", **kw), "c2_butterpass": lambda **kw: c2_lti_bank("butterpass", **kw), "c2_lowpole": lambda **kw: c2_lti_bank("lowpole", **kw),
             "c1": c1_hello, "c2": c2_lowpass_bank, "c3": c3_polysynth, "c4": c4_spectral, "c5": c5_mixed}
