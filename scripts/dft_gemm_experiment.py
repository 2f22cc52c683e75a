"""ONE experiment: the 2048-point transforms of configs[3] as dense contractions on the tensor cores.

The north star allows tensor cores "only if a dense DFT/convolution contraction measurably wins".  cuBLAS (tcgen05 on sm_100a)
stands in for the best a hand-written tcgen05 kernel could do: frames [F, N] x twiddle matrix [N, 2N] (real input -> re | im)
for the forward transform, [F, 2N] x [2N, 2N] for the complex inverse.  Reported: time per configs[3] render (2.88 M frames)
in bf16, tf32 and a 3-way bf16 split, next to the transforms' share of the shipped kernel; and what the precision does to the
spectral gate: bins whose `magnitude > threshold` decision differs from the f32 FFT's.
    python scripts/dft_gemm_experiment.py > profiles/r02_dft_gemm_experiment.txt"""
import math
import sys

import torch

N, F, THR = 2048, 32768, 16.0
FRAMES_PER_RENDER = 1024 * 4 * 703
dev = "cuda"
torch.manual_seed(0)
hann = 0.5 - 0.5 * torch.cos(2 * math.pi * torch.arange(N, device=dev, dtype=torch.float64) / N)
x = ((torch.rand(F, N, device=dev, dtype=torch.float64) * 2 - 1) * hann).float()
k = torch.arange(N, device=dev, dtype=torch.float64)
ang = -2 * math.pi * torch.outer(k, k) / N
Wf = torch.cat([torch.cos(ang), torch.sin(ang)], dim=1)                       # [N, 2N]: x @ Wf = (re | im)
Wi = torch.cat([torch.cat([torch.cos(ang), -torch.sin(ang)], 0), torch.cat([torch.sin(ang), torch.cos(ang)], 0)], 1) / N   # [2N, 2N]
ref = torch.fft.fft(x.double(), dim=1)
ref_mag = ref.abs()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def split3(a):
    hi = a.bfloat16()
    r1 = a - hi.float()
    mid = r1.bfloat16()
    lo = (r1 - mid.float()).bfloat16()
    return hi, mid, lo


def report(name, fwd_ms, inv_ms, spec):
    mag = torch.sqrt(spec[:, :N].double() ** 2 + spec[:, N:].double() ** 2)
    err = (mag - ref_mag).abs().max().item()
    flips = ((mag > THR) != (ref_mag > THR)).double().mean().item()
    per_render = (fwd_ms + inv_ms) * FRAMES_PER_RENDER / F
    print(f"{name:34s} forward {fwd_ms:7.3f} ms + inverse {inv_ms:7.3f} ms per {F} frames -> {per_render:7.1f} ms per configs[3] render; "
          f"max |magnitude error| {err:.2e}; gate decisions that differ from the f32 FFT: {flips * 100:.4f} % of bins")


print(f"# {torch.cuda.get_device_name(0)}, torch {torch.__version__}; N = {N}, {F} frames per batch, threshold {THR}")
print("# shipped kernel (K5s): 100.4 ms per render in total, of which the two radix-2 transforms are ~45 ms (profiles/r02_c4_ncu_summary.txt)")
# f32 cuFFT for scale
ms_f, _ = timed(lambda: torch.fft.fft(x, dim=1))
ms_i, _ = timed(lambda: torch.fft.ifft(torch.fft.fft(x, dim=1), dim=1))
print(f"{'cuFFT f32 (library, not bit-exact)':34s} forward {ms_f:7.3f} ms, forward+inverse {ms_i:7.3f} ms -> {ms_i * FRAMES_PER_RENDER / F:7.1f} ms per render")
# bf16 single pass
xb, Wfb, Wib = x.bfloat16(), Wf.bfloat16(), Wi.bfloat16()
ms1, spec = timed(lambda: torch.mm(xb, Wfb, out_dtype=torch.float32))
specb = spec.bfloat16()
ms2, _ = timed(lambda: torch.mm(specb, Wib, out_dtype=torch.float32))
report("bf16 x bf16 (one pass, f32 out)", ms1, ms2, spec)
# tf32
torch.backends.cuda.matmul.allow_tf32 = True
Wf32, Wi32 = Wf.float(), Wi.float()
ms1, spec = timed(lambda: (x @ Wf32))
ms2, _ = timed(lambda: (spec @ Wi32))
report("tf32 (f32 operands, tf32 multiply)", ms1, ms2, spec)
torch.backends.cuda.matmul.allow_tf32 = False
# 3-way bf16 split of both operands: 6 of the 9 partial products (hi*hi, hi*mid, mid*hi, hi*lo, lo*hi, mid*mid), f32 accumulate
xs, ws, wis = split3(x), split3(Wf.float()), split3(Wi.float())
pairs = [(0, 0), (0, 1), (1, 0), (0, 2), (2, 0), (1, 1)]


def split_mm(a3, b3):
    acc = None
    for i, j in pairs:
        p = torch.mm(a3[i], b3[j], out_dtype=torch.float32)        # bf16 operands, f32 accumulator and result
        acc = p if acc is None else acc.add_(p)
    return acc


ms1, spec = timed(lambda: split_mm(xs, ws), reps=3)
ss = split3(spec)
ms2, _ = timed(lambda: split_mm(ss, wis), reps=3)
report("bf16 3-way split (6 products)", ms1, ms2, spec)
# plain f32 (no tensor cores) for reference accuracy
ms1, spec = timed(lambda: (x @ Wf32), reps=2)
report("f32 SIMT GEMM (no tensor cores)", ms1, float("nan"), spec)
