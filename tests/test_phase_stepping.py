"""The time-vector kernel steps a constant-increment f32 phase accumulator (`sine(440)`, `dc(f) >> ramp()`) a BINADE at a time
(csrc/interp.cu: tv_phase_const).  This file restates that algorithm in scalar Python with numpy float32 arithmetic and holds
it BIT FOR BIT to the reference's sample-by-sample recurrence (FunDSP sine: `ph += inc; ph -= floor(ph)`; nodes.rs:476-483 ramp:
`val += inc; if val >= 1 { val -= 1 }`) — the exactness argument of the device code, executable on the CPU: random increments
and start phases, increments that round away, and increments whose low bit sits exactly half an ulp below a binade (ties)."""
import math
import struct

import numpy as np

f32=np.float32
def bits(x): return struct.unpack('<I', struct.pack('<f', float(x)))[0]
def frombits(b): return f32(struct.unpack('<f', struct.pack('<I', b & 0xffffffff))[0])
def naive(ph, inc, n, floor_wrap=True):
    out=np.empty(n, f32); ph=f32(ph); inc=f32(inc)
    for j in range(n):
        out[j]=ph; ph=f32(ph+inc)
        if floor_wrap: ph=f32(ph-f32(math.floor(ph)))
        elif ph>=f32(1): ph=f32(ph-f32(1))
    return out, ph
def fast(ph, inc, n, floor_wrap=True):
    out=np.empty(n, f32); ph=f32(ph); inc=f32(inc)
    fast_ok = inc>0 and inc<0.5
    j=0; segs=0
    while j<n:
        cnt=1; d=f32(0); done=False
        p1=f32(ph+inc)
        eb=bits(ph)&0x7f800000
        if fast_ok and ph>0 and p1<1 and eb>=(40<<23) and (bits(p1)&0x7f800000)==eb:
            ulp=frombits(eb-(23<<23)); top=frombits(eb+(1<<23))
            dd=f32(p1-ph)
            r=f32(inc-f32(f32(math.floor(f32(inc/ulp)))*ulp))
            if dd==0: cnt=n-j; pn=ph; done=True
            elif r!=f32(0.5)*ulp:
                M=int(min(math.ceil((float(top)-float(ph))/float(dd)),1e9))
                d=dd
                if M<=n-j:
                    cnt=M; pn=f32(f32(float(ph)+float(M-1)*float(dd))+inc)
                    if floor_wrap: pn=f32(pn-f32(math.floor(pn)))
                    elif pn>=1: pn=f32(pn-f32(1))
                else:
                    cnt=n-j; pn=f32(float(ph)+float(cnt)*float(dd))
                done=True
        if not done:
            if floor_wrap: p1=f32(p1-f32(math.floor(p1)))
            elif p1>=1: p1=f32(p1-f32(1))
            pn=p1
        for m in range(cnt): out[j+m]=f32(float(ph)+float(m)*float(d))
        j+=cnt; ph=pn; segs+=1
    return out, ph, segs


def _cases():
    rng = np.random.default_rng(1)
    cases = [(440 / 48000, 0.123), (440 / 44100, 0.0), (0.25, 0.0), (0.3333333, 0.9), (1e-7, 0.7), (3e-9, 0.6), (0.49999997, 0.2),
             (2 ** -10, 0.0), (2 ** -10 + 2 ** -30, 0.0), (-0.01, 0.5), (0.75, 0.1)]
    for _ in range(120):
        cases.append((float(f32(np.exp(rng.uniform(np.log(1e-6), np.log(0.49))))), float(f32(rng.uniform(0, 1)))))
    for e in (24, 25, 26, 27, 28):                 # lowest set bit at half an ulp of a high binade: round-to-even alternates
        cases.append((float(f32(2 ** -7 + 2 ** -(e + 1))), 0.1))
    return cases


def test_binade_stepping_is_the_sample_by_sample_recurrence_bit_for_bit():
    total_segments = total_samples = 0
    for inc, ph in _cases():
        for floor_wrap in (True, False):
            n = 2048
            a, pa = naive(ph, inc, n, floor_wrap)
            b, pb, segs = fast(ph, inc, n, floor_wrap)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and bits(pa) == bits(pb), (inc, ph, floor_wrap)
            total_segments += segs
            total_samples += n
    assert total_samples / total_segments > 3.0       # and it does take several samples per sequential step
