// K5 — frame-parallel spectral path (see spectral.h).  Replaces, for patches of the spectral-gate shape, the per-sample
// streaming of /root/reference/src/nodes.rs:601-700 (`Rfft::tick` / `Ifft::tick`: one bin per sample, a transform every N
// samples) by whole-frame evaluation: a CTA owns ONE frame of one instance of one voice and keeps it in shared memory from
// the N input samples to the N resynthesised samples.
//
// Timing of the reference, restated (n = N, s = start, t = samples since reset, both nodes share n and s):
//   boundaries t_b = -s (mod n).  At t_b the rfft transforms x[t_b - n .. t_b) (zeros before t = 0) and emits bin i of that
//   spectrum at t_b + i (conjugate-mirrored above n/2); the ifft stores chain(bin_i) at position i, and at the NEXT boundary
//   transforms what it collected — positions whose sample time was negative still hold their initial zero — and emits
//   output i at t_b + n + i.  So the ifft output over [t_b, t_b + n) is  IFFT(B)  with
//       B[i] = (tau + i >= 0) ? chain(mirror(FFT(x[tau - n .. tau)))[i]) : 0,   tau = t_b - n.
// The arithmetic of the transforms is the radix-2 decimation-in-time network of every other path (same twiddle table, same
// operation order per butterfly), so threshold decisions in a gate fall exactly where the oracle's do.
#define QG_SPEC_ONLY 1
#include "interp.cu"
#undef QG_SPEC_ONLY

#include <algorithm>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

#include "spectral.h"

namespace qg {

// ------------------------------------------------------------------------------------------------ planner (host)
namespace {

int n_out_of(const Instr& i) {
  switch (i.op) {
    case OP_POL: case OP_CAR: case OP_PAN: case OP_ROTATE: case OP_RFFT: case OP_IFFT: return 2;
    case OP_NOP: return 0;
    default: return 1;
  }
}
int n_in_of(const Instr& i) {
  const uint16_t op = i.op;
  if (op == OP_MOV) return 1;
  if (op == OP_ZERO || op == OP_NOP) return 0;
  if (op >= OP_ADD && op <= OP_PDHALF_UNI) return 2;
  if (op >= OP_LERP && op <= OP_DEXERP11) return 3;
  if (op == OP_SPLINE) return 5;
  if (op >= OP_ABS && op <= OP_MIRROR) return 1;
  if (op == OP_POL || op == OP_CAR || op == OP_ROTATE) return 2;
  if (op == OP_JOIN) return std::min<int>(i.n, 5);
  if (op == OP_DIVN || op == OP_PAN || op == OP_QUANTIZE || op == OP_ARR_GET) return 1;
  if (op == OP_DELAY || op == OP_TICK || op == OP_RFFT) return 1;
  if (op == OP_IFFT) return 2;
  return 0;
}

enum { CLS_RA = 0, CLS_POST = 1, CLS_BIN = 2 };   // BIN of segment r = CLS_BIN + r

struct Planner {
  const Tape& t;
  SpPlan& pl;
  int PS;
  std::vector<int> cls;                       // class of every instruction's outputs
  std::vector<int> seg_of;                    // rfft / ifft instruction -> segment
  std::vector<std::array<int, 5>> src;        // producer instruction of every temporary operand (-1: scalar)
  std::vector<int> rfft_instr, ifft_instr;
  // mini-tape under construction
  std::vector<Instr> mt;
  int slots = 0;
  std::map<std::pair<int, long>, int> memo;   // (instruction, time offset) -> X index of its first output
  bool fail = false;

  Planner(const Tape& t_, SpPlan& p) : t(t_), pl(p), PS((int)(t_.h.n_params + t_.h.n_state)) {}
  int new_slots(int n) { int s = slots; slots += n; return PS + s; }

  bool classify() {
    const int n = (int)t.code.size(), NT = (int)t.h.n_temps;
    std::vector<int> def(NT, -1);
    cls.assign(n, -1); seg_of.assign(n, -1); src.assign(n, {-1, -1, -1, -1, -1});
    if (t.h.n_inputs != 0 || (t.h.flags & TAPE_DIVERGENT)) return false;
    for (int k = 0; k < n; k++) {
      const Instr& I = t.code[k];
      const int ni = n_in_of(I);
      bool any_post = false, any_ra = false;
      int bin = -1;
      for (int q = 0; q < ni; q++) {
        const int x = I.in[q];
        if (x < PS) continue;
        const int p = def[x - PS];
        if (p < 0) return false;
        src[k][q] = p;
        const int c = cls[p];
        if (c == CLS_RA) any_ra = true;
        else if (c == CLS_POST) any_post = true;
        else { if (bin >= 0 && bin != c - CLS_BIN) return false; bin = c - CLS_BIN; }
      }
      if (op_is_stateless(I.op)) {
        if (I.op == OP_NOP) continue;
        if (bin >= 0) { if (any_ra || any_post) return false; cls[k] = CLS_BIN + bin; }
        else cls[k] = any_post ? CLS_POST : CLS_RA;
      } else if (I.op == OP_NOISE || I.op == OP_WAVE || I.op == OP_IMPULSE) {
        cls[k] = CLS_RA;
      } else if (I.op == OP_DELAY || I.op == OP_TICK) {
        if (bin >= 0 || any_post) return false;
        cls[k] = CLS_RA;
      } else if (I.op == OP_RFFT) {
        if (bin >= 0 || any_post) return false;
        if (I.n < 3 || I.n > 12) return false;
        seg_of[k] = (int)rfft_instr.size();
        cls[k] = CLS_BIN + seg_of[k];
        rfft_instr.push_back(k); ifft_instr.push_back(-1);
      } else if (I.op == OP_IFFT) {
        if (bin < 0 || any_ra || any_post) return false;
        const Instr& Rf = t.code[rfft_instr[bin]];
        if (ifft_instr[bin] >= 0 || Rf.n != I.n || t.state_init[Rf.s - t.h.n_params] != t.state_init[I.s - t.h.n_params]) return false;
        ifft_instr[bin] = k; seg_of[k] = bin;
        cls[k] = CLS_POST;
      } else {
        return false;
      }
      const int no = n_out_of(I);
      for (int q = 0; q < no; q++) {
        const int x = (int)I.out + q;
        if (x < PS || x - PS >= NT) return false;
        def[x - PS] = k;
      }
    }
    if (rfft_instr.empty()) return false;
    for (int r : ifft_instr) if (r < 0) return false;
    out_src.clear();
    for (uint16_t o : t.out_x) {
      int p = -1;
      if (o >= PS) { p = def[o - PS]; if (p < 0 || cls[p] >= CLS_BIN) return false; }
      out_src.push_back(p);
    }
    return true;
  }
  std::vector<int> out_src;

  int operand(int k, int q, long d) {
    const int x = t.code[k].in[q];
    if (x < PS) return x;
    const int p = src[k][q];
    return emit(p, d) + (x - (int)t.code[p].out);
  }
  // X index (mini-tape numbering) of the first output of instruction p evaluated `d` samples away from the mini-tape's time
  int emit(int p, long d) {
    const Instr& I = t.code[p];
    if (I.op == OP_IFFT) {   // the resynthesised signal comes from the Y ring, at the mini-tape's own time only
      if (d != 0) { fail = true; return PS; }
      auto it = memo.find({p, 0});
      if (it != memo.end()) return it->second;
      const int base = new_slots(2);
      memo[{p, 0}] = base;
      return base;           // the loads are added in build(), for the components that are actually read
    }
    auto it = memo.find({p, d});
    if (it != memo.end()) return it->second;
    if (d < -(1L << 30)) { fail = true; return PS; }
    Instr J = I;
    J.pad = (uint32_t)(int32_t)d;
    if (I.op == OP_DELAY || I.op == OP_TICK) {
      const long L = I.op == OP_TICK ? 1 : (long)t.rings[I.aux].length;
      J.in[0] = (uint16_t)operand(p, 0, d - L);
      J.pad = (uint32_t)(int32_t)(d - L);          // out = (time + d - L >= 0) ? in : 0
    } else {
      const int ni = n_in_of(I);
      for (int q = 0; q < ni; q++) J.in[q] = (uint16_t)operand(p, q, d);
    }
    const int base = new_slots(n_out_of(I));
    J.out = (uint16_t)base;
    mt.push_back(J);
    memo[{p, d}] = base;
    return base;
  }

  // linear-scan reuse of the mini-tape's temporaries; `live` = X indices read after the mini-tape (kept to the end)
  int compact(std::vector<Instr>& code, std::vector<int*> live, int reserved) {
    const int n = (int)code.size();
    std::vector<int> last(slots, -1);
    for (int k = 0; k < n; k++) {
      const int ni = code[k].op == OP_STREAM_IN ? 0 : n_in_of(code[k]);
      for (int q = 0; q < ni; q++) if (code[k].in[q] >= PS) last[code[k].in[q] - PS] = k;
    }
    for (int* x : live) if (*x >= PS) last[*x - PS] = n;
    std::vector<int> phys(slots, -1), free_at;
    for (int s = 0; s < reserved; s++) { phys[s] = s; free_at.push_back(std::max(last[s], 0)); }
    for (int k = 0; k < n; k++) {
      const int no = n_out_of(code[k]), v0 = code[k].out - PS;
      int die = k;
      for (int q = 0; q < no; q++) die = std::max(die, last[v0 + q]);
      int base = -1;
      for (int s0 = reserved; s0 + no <= (int)free_at.size() && base < 0; s0++) {
        bool ok = true;
        for (int q = 0; q < no; q++) ok = ok && free_at[s0 + q] < k;
        if (ok) base = s0;
      }
      if (base < 0) { base = (int)free_at.size(); free_at.resize(base + no, -1); }
      for (int q = 0; q < no; q++) { phys[v0 + q] = base + q; free_at[base + q] = die; }
    }
    auto rn = [&](int x) { return x < PS ? x : PS + std::max(phys[x - PS], 0); };
    for (Instr& I : code) {
      const int ni = I.op == OP_STREAM_IN ? 0 : n_in_of(I);
      for (int q = 0; q < ni; q++) I.in[q] = (uint16_t)rn(I.in[q]);
      I.out = (uint16_t)rn(I.out);
    }
    for (int* x : live) *x = rn(*x);
    return (int)free_at.size();
  }
  void begin() { mt.clear(); slots = 0; memo.clear(); }

  // How the chain's two results behave when the input bin is conjugated: +1 unchanged, -1 negated, 0 unknown.  The streamed
  // bins above N/2 are conjugates of those below (nodes.rs:640-641); when both results have a known type the frame kernel
  // evaluates the chain on bins 0 .. N/2 only and mirrors the rest.  Only identities that hold BIT FOR BIT are used: any
  // function of unchanged values is unchanged; x+y, x-y keep a common type; products multiply types; |x|, x^2, hypotf, cosf
  // are even and sinf, atan2f(y, .) odd in the device library's implementations (sign handled by copysign / odd polynomials).
  void conj_types(const std::vector<Instr>& code, int xr, int xi, int* t_re, int* t_im) {
    std::map<int, int> ty;
    ty[PS] = +1; ty[PS + 1] = -1;
    auto T = [&](int x) { if (x < PS) return +1; auto it = ty.find(x); return it == ty.end() ? 0 : it->second; };
    for (const Instr& I : code) {
      const int ni = n_in_of(I);
      bool all_even = true;
      for (int q = 0; q < ni; q++) all_even = all_even && T(I.in[q]) == +1;
      const int a = ni > 0 ? T(I.in[0]) : +1, b = ni > 1 ? T(I.in[1]) : +1;
      int o0 = 0, o1 = 0;
      if (all_even) o0 = o1 = +1;
      else switch (I.op) {
        case OP_MOV: case OP_CUBED: case OP_SIN: o0 = a; break;
        case OP_ADD: case OP_SUB: o0 = a == b ? a : 0; break;
        case OP_MUL: o0 = a * b; break;
        case OP_ABS: case OP_SQUARED: case OP_COS: o0 = a != 0 ? +1 : 0; break;
        case OP_HYPOT: o0 = (a != 0 && b != 0) ? +1 : 0; break;
        case OP_ATAN2: o0 = b == +1 ? a : 0; break;                  // atan2f(in0 = y, in1 = x)
        case OP_POL: o0 = (a != 0 && b != 0) ? +1 : 0; o1 = a == +1 ? b : 0; break;   // hypotf(a, b), atan2f(b, a)
        case OP_CAR: o0 = a * (b != 0 ? +1 : 0); o1 = a * b; break;                   // a cosf(b), a sinf(b)
        default: break;
      }
      ty[I.out] = o0;
      if (n_out_of(I) > 1) ty[I.out + 1] = o1;
    }
    *t_re = T(xr); *t_im = T(xi);
    if (*t_re == 0 || *t_im == 0) *t_re = *t_im = 0;
  }

  bool build() {
    if (!classify()) return false;
    const int S = (int)rfft_instr.size();
    int C = 0;
    for (int r = 0; r < S; r++) C = std::max(C, 1 << t.code[rfft_instr[r]].n);
    pl.C = C;
    pl.segs.resize(S);
    for (int r = 0; r < S; r++) {
      const Instr& Rf = t.code[rfft_instr[r]];
      const Instr& If = t.code[ifft_instr[r]];
      SpSegment& sg = pl.segs[r];
      sg.lg = Rf.n; sg.start = (int)t.state_init[Rf.s - t.h.n_params]; sg.tw = (int)Rf.aux2;
      sg.y_re = sg.y_im = -1;
      if (sg.start < 0 || sg.start >= (1 << sg.lg) || If.aux2 != Rf.aux2 || (Rf.aux2 & 1)) return false;
      // ---- pre: the rfft's input as a function of time
      begin();
      int px = operand(rfft_instr[r], 0, 0);
      if (fail) return false;
      pl.n_slots_frame = std::max(pl.n_slots_frame, compact(mt, {&px}, 0));
      sg.pre_lo = (int)pl.code.size();
      pl.code.insert(pl.code.end(), mt.begin(), mt.end());
      sg.pre_hi = (int)pl.code.size();
      sg.pre_x = px;
      // ---- chain: the BIN-class instructions of this segment in program order; slots 0, 1 = the rfft's outputs
      begin();
      slots = 2;
      std::map<int, int> cslot;   // instruction -> X index of its first output
      cslot[rfft_instr[r]] = PS;
      auto cop = [&](int k, int q) -> int {
        const int x = t.code[k].in[q];
        if (x < PS) return x;
        const int p = src[k][q];
        return cslot.at(p) + (x - (int)t.code[p].out);
      };
      for (int k = rfft_instr[r] + 1; k < ifft_instr[r]; k++) {
        if (cls[k] != CLS_BIN + r || seg_of[k] >= 0) continue;
        Instr J = t.code[k];
        const int ni = n_in_of(J);
        for (int q = 0; q < ni; q++) J.in[q] = (uint16_t)cop(k, q);
        const int base = new_slots(n_out_of(J));
        J.out = (uint16_t)base; J.pad = 0;
        cslot[k] = base;
        mt.push_back(J);
      }
      int xr = cop(ifft_instr[r], 0), xi = cop(ifft_instr[r], 1);
      pl.n_slots_frame = std::max(pl.n_slots_frame, compact(mt, {&xr, &xi}, 2));
      sg.ch_lo = (int)pl.code.size();
      pl.code.insert(pl.code.end(), mt.begin(), mt.end());
      sg.ch_hi = (int)pl.code.size();
      sg.rf_x = PS; sg.in_re_x = xr; sg.in_im_x = xi;
      conj_types(mt, xr, xi, &sg.sym_re, &sg.sym_im);
    }
    // chain instructions after their ifft would have been classified BIN only if they read BIN values, which the closed
    // segment check rejects; instructions of class BIN that no ifft consumes are dead code
    // ---- post: outputs as a function of time, reading the Y ring
    begin();
    std::vector<int> ox(t.out_x.size());
    for (size_t c = 0; c < t.out_x.size(); c++) {
      const int o = t.out_x[c], p = out_src[c];
      ox[c] = p < 0 ? o : emit(p, 0) + (o - (int)t.code[p].out);
    }
    if (fail) return false;
    // Y loads for the components the post-graph reads, placed first
    std::vector<Instr> loads;
    auto used = [&](int x) {
      for (const Instr& I : mt) { const int ni = n_in_of(I); for (int q = 0; q < ni; q++) if (I.in[q] == x) return true; }
      for (int o : ox) if (o == x) return true;
      return false;
    };
    for (int r = 0; r < S; r++) {
      auto it = memo.find({ifft_instr[r], 0});
      if (it == memo.end()) continue;
      for (int comp = 0; comp < 2; comp++) {
        if (!used(it->second + comp)) continue;
        Instr J;
        memset(&J, 0, sizeof J);
        J.op = OP_STREAM_IN; J.out = (uint16_t)(it->second + comp); J.aux = (uint32_t)pl.n_streams;
        (comp ? pl.segs[r].y_im : pl.segs[r].y_re) = pl.n_streams++;
        loads.push_back(J);
      }
    }
    mt.insert(mt.begin(), loads.begin(), loads.end());
    std::vector<int*> live;
    for (int& o : ox) live.push_back(&o);
    // (a STREAM_IN writes one slot although its ifft reserved two: compact() only looks at what instructions write)
    pl.n_slots_post = compact(mt, live, 0);
    pl.post_lo = (int)pl.code.size();
    pl.code.insert(pl.code.end(), mt.begin(), mt.end());
    pl.post_hi = (int)pl.code.size();
    for (int o : ox) pl.out_x.push_back((uint16_t)o);
    if (pl.n_streams == 0) return false;   // nothing of the resynthesis reaches an output: leave the tape to the general path
    for (int r = 0; r < S; r++)
      for (int f = 0; f < (C >> pl.segs[r].lg); f++) pl.items.push_back({r, f});
    if (pl.n_slots_frame > 48 || pl.n_slots_post > 96 || pl.code.size() > 512) return false;
    return true;
  }
};

}  // namespace

SpPlan plan_spectral(const Tape& t) {
  SpPlan pl;
  Planner P(t, pl);
  pl.ok = P.build();
  if (!pl.ok) { SpPlan none; return none; }
  if (getenv("QG_DEBUG_PLAN"))
    fprintf(stderr, "spectral plan: %zu segments, %zu instr, %d frame slots, %d post slots, %d streams, sym %d %d\n", pl.segs.size(),
            pl.code.size(), pl.n_slots_frame, pl.n_slots_post, pl.n_streams, pl.segs[0].sym_re, pl.segs[0].sym_im);
  return pl;
}

// ------------------------------------------------------------------------------------------------ device
#include "spectral_fft.cuh"

struct SpSmem { int ps_off, tmp_off, f_off; };   // float offsets: [code] [P + NS scalars] [slots x HB] [transform buffers]

// Evaluates mini-tape [lo, hi) for the n consecutive sample times tbase .. tbase + n - 1 (thread tid owns columns tid,
// tid + nth, ...: a thread only ever reads columns it wrote, so no barrier is needed between instructions).  Times are
// 64-bit only here, at the block level: per sample everything is a 32-bit offset from the block's base.
__device__ __forceinline__ void sp_eval(const Instr* code, int lo, int hi, long tbase, int n, int HB, int PS, const SpSmem& sm,
                                        const SpArgs& a, int v, int tid, int nth) {
  float* ps = QG_SMEM_F + sm.ps_off;
  float* tmp = QG_SMEM_F + sm.tmp_off;
  for (int pc = lo; pc < hi; pc++) {
    const Instr I = code[pc];
    float* o = tmp + ((int)I.out - PS) * HB;
    const long tb = tbase + (long)(int32_t)I.pad;           // time of column 0 for this instruction
    switch (I.op) {
      case OP_NOISE: {
        const uint32_t c0 = __float_as_uint(ps[I.s]) + (uint32_t)tb + 1u;
        for (int j = tid; j < n; j += nth) o[j] = d_noise(c0 + (uint32_t)j);
        break;
      }
      case OP_WAVE: {
        const long len = (long)I.aux2;
        long m0 = ((long)__float_as_uint(ps[I.s]) + tb + tid) % len;
        if (m0 < 0) m0 += len;
        const uint32_t ulen = (uint32_t)len, step = (uint32_t)(nth % len);
        uint32_t m = (uint32_t)m0;
        const float* tab = a.tables + I.aux;
        for (int j = tid; j < n; j += nth) { o[j] = __ldg(tab + m); m += step; if (m >= ulen) m -= ulen; }
        break;
      }
      case OP_IMPULSE: {
        const bool armed = __float_as_uint(ps[I.s]) == 0u;
        const int jz = (tb <= 0 && -tb < (long)n) ? (int)-tb : -1;
        for (int j = tid; j < n; j += nth) o[j] = (armed && j == jz) ? 1.0f : 0.0f;
        break;
      }
      case OP_DELAY: case OP_TICK: {   // the shifted input, silent while the line is still filling
        const int x = I.in[0];
        const int jz = tb >= 0 ? 0 : (-tb > (long)n ? n : (int)-tb);     // columns below jz lie before the stream's start
        if (x < PS) { const float c = ps[x]; for (int j = tid; j < n; j += nth) o[j] = j >= jz ? c : 0.0f; }
        else { const float* src = tmp + (x - PS) * HB; for (int j = tid; j < n; j += nth) o[j] = j >= jz ? src[j] : 0.0f; }
        break;
      }
      case OP_STREAM_IN: {
        const float* y = a.y + ((size_t)I.aux * a.V + v) * (size_t)a.ring;
        const uint32_t r0 = (uint32_t)tbase, mask = (uint32_t)a.ring - 1u;
        for (int j = tid; j < n; j += nth) o[j] = y[(r0 + (uint32_t)j) & mask];
        break;
      }
      default: {
        TvSample L{sm.ps_off, sm.tmp_off, PS, HB, a.tables, tid, nth, n, tid};
        int dummy = 0;
        exec(I, L, dummy);
        break;
      }
    }
  }
}

__device__ __forceinline__ void sp_stage(const SpArgs& a, const SpSmem& sm, int v, int tid, int nth) {
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.code);
    uint4* dst = reinterpret_cast<uint4*>(qg_smem);
    for (int i = tid; i < a.n_code * 2; i += nth) dst[i] = src[i];
  }
  float* ps = QG_SMEM_F + sm.ps_off;
  for (int p = tid; p < a.P; p += nth) ps[p] = a.params[(size_t)p * a.Vp + v];
  for (int s = tid; s < a.NS; s += nth) ps[a.P + s] = a.state_init[(size_t)s * a.Vp + v];
}

// One CTA = one frame: (voice, work item of a round, round).  c0 = first round of this launch.  HB = columns per
// mini-tape block = N_max / 2 + 1 (the non-redundant bins of the largest transform in one block).
__global__ void __launch_bounds__(256, 3) k_spectral_frames(SpArgs a, long c0, int n_rounds, int HB, SpSmem sm) {
  const int tid = threadIdx.x, nth = blockDim.x;
  const int per_voice = a.n_items * n_rounds;
  const int v = (int)(blockIdx.x / (unsigned)per_voice), w = (int)(blockIdx.x % (unsigned)per_voice);
  const SpItem it = a.items[w % a.n_items];
  const long c = c0 + w / a.n_items;
  const SpSegment sg = a.segs[it.seg];
  const int lg = sg.lg, N = 1 << lg, PS = a.P + a.NS;
  const long tb = c * (long)a.C + ((N - sg.start) & (N - 1)) + (long)it.frame * N;   // this frame's output starts here
  const long tau = tb - N;                                                            // its bins were streamed from tau on
  float* yre = sg.y_re >= 0 ? a.y + ((size_t)sg.y_re * a.V + v) * (size_t)a.ring : nullptr;
  float* yim = sg.y_im >= 0 ? a.y + ((size_t)sg.y_im * a.V + v) * (size_t)a.ring : nullptr;
  // only sample times of this call are read by its post pass: columns [i_lo, i_hi) of the frame
  const long lo = a.t0 - tb, hi = a.t0 + a.T - tb;
  if (hi <= 0 || lo >= N) return;
  const int i_lo = lo > 0 ? (int)lo : 0, i_hi = hi < N ? (int)hi : N;
  const uint32_t ymask = (uint32_t)a.ring - 1u, y0 = (uint32_t)tb;
  if (tau + N <= 0) {                            // every position of B predates the stream: the inverse transform of zeros
    for (int i = i_lo + tid; i < i_hi; i += nth) {
      if (yre) yre[(y0 + (uint32_t)i) & ymask] = 0.0f;
      if (yim) yim[(y0 + (uint32_t)i) & ymask] = 0.0f;
    }
    return;
  }
  const Instr* code = reinterpret_cast<const Instr*>(qg_smem);
  sp_stage(a, sm, v, tid, nth);
  float* ps = QG_SMEM_F + sm.ps_off;
  float* tmp = QG_SMEM_F + sm.tmp_off;
  const int f_off = sm.f_off / 2, g_off = f_off + CPAD(N) + 1;     // float2 offsets into the CTA's shared memory
  float2* f = QG_SMEM_C + f_off;
  float2* g = QG_SMEM_C + g_off;
  __syncthreads();
  const float2* tw = reinterpret_cast<const float2*>(a.tables + sg.tw);
  const int sh = 32 - lg;
  // ---- the frame's N input samples x[tau - N + m], bit-reversed into the transform buffer
  {
    const int hb = N < HB - 1 ? N : HB - 1;                  // HB - 1 = N_max / 2: a power of two
    const bool scalar = sg.pre_x < PS;
    const float cs = scalar ? ps[sg.pre_x] : 0.0f;
    const float* src = tmp + (sg.pre_x - PS) * HB;
    for (int m0 = 0; m0 < N; m0 += hb) {
      const long t0 = tau - N + m0;
      if (t0 + hb > 0) sp_eval(code, sg.pre_lo, sg.pre_hi, t0, hb, HB, PS, sm, a, v, tid, nth);
      const int jz = t0 >= 0 ? 0 : (-t0 > (long)hb ? hb : (int)-t0);
      for (int j = tid; j < hb; j += nth) {
        const uint32_t rv = __brev((uint32_t)(m0 + j)) >> sh;
        f[CPAD(rv)] = make_float2(j >= jz ? (scalar ? cs : src[j]) : 0.0f, 0.0f);
      }
    }
  }
  __syncthreads();
  sp_fft(f_off, lg, tw, 1.0f, tid, nth);
  // ---- the bin chain on the streamed bins (conjugate mirror above N/2, nodes.rs:637-642), into the inverse transform's
  // input; positions whose sample time was negative keep the buffer's initial zero.  A chain whose results have a known
  // behaviour under conjugation (plan: sym_re / sym_im) runs on bins 0 .. N/2 only.
  {
    const bool sym = sg.sym_re != 0;
    const int half = N >> 1, nb = half + 1 <= HB ? half + 1 : HB;      // bins per block
    const int i_end = sym ? half + 1 : N;
    const int iz = tau >= 0 ? 0 : (-tau > (long)N ? N : (int)-tau);     // bins below iz predate the stream
    float* c_re = tmp;            // slots 0, 1 = the rfft's two outputs
    float* c_im = tmp + HB;
    const float* o_re = sg.in_re_x < PS ? nullptr : tmp + (sg.in_re_x - PS) * HB;
    const float* o_im = sg.in_im_x < PS ? nullptr : tmp + (sg.in_im_x - PS) * HB;
    const float k_re = o_re ? 0.0f : ps[sg.in_re_x], k_im = o_im ? 0.0f : ps[sg.in_im_x];
    const float s_re = (float)sg.sym_re, s_im = (float)sg.sym_im;
    for (int i0 = 0; i0 < i_end; i0 += nb) {
      const int n = i_end - i0 < nb ? i_end - i0 : nb;
      for (int j = tid; j < n; j += nth) {
        const int i = i0 + j;
        float2 z;
        if (i <= half) z = f[CPAD(i)];
        else { z = f[CPAD(N - i)]; z.y = -z.y; }
        c_re[j] = z.x; c_im[j] = z.y;
      }
      sp_eval(code, sg.ch_lo, sg.ch_hi, tau + i0, n, HB, PS, sm, a, v, tid, nth);
      for (int j = tid; j < n; j += nth) {
        const int i = i0 + j;
        const float2 z = make_float2(o_re ? o_re[j] : k_re, o_im ? o_im[j] : k_im);
        g[CPAD(__brev((uint32_t)i) >> sh)] = i >= iz ? z : make_float2(0.0f, 0.0f);
        if (sym && i > 0 && i < half) {
          const int im = N - i;
          g[CPAD(__brev((uint32_t)im) >> sh)] = im >= iz ? make_float2(z.x * s_re, z.y * s_im) : make_float2(0.0f, 0.0f);
        }
      }
    }
  }
  __syncthreads();
  sp_fft(g_off, lg, tw, -1.0f, tid, nth);
  const float sc = 1.0f / (float)N;
  for (int i = i_lo + tid; i < i_hi; i += nth) {
    const float2 z = g[CPAD(i)];
    if (yre) yre[(y0 + (uint32_t)i) & ymask] = z.x * sc;
    if (yim) yim[(y0 + (uint32_t)i) & ymask] = z.y * sc;
  }
}

// The post-graph: one CTA = HB consecutive output samples of one voice.
__global__ void __launch_bounds__(256) k_spectral_post(SpArgs a, long t_lo, long t_hi, int HB, SpSmem sm) {
  const int tid = threadIdx.x, nth = blockDim.x;
  const int nblk = (int)((t_hi - t_lo + HB - 1) / HB);
  const int v = (int)(blockIdx.x / (unsigned)nblk), b = (int)(blockIdx.x % (unsigned)nblk);
  const long tbase = t_lo + (long)b * HB;
  const int n = (int)(t_hi - tbase < HB ? t_hi - tbase : HB), PS = a.P + a.NS;
  const Instr* code = reinterpret_cast<const Instr*>(qg_smem);
  sp_stage(a, sm, v, tid, nth);
  __syncthreads();
  sp_eval(code, a.post_lo, a.post_hi, tbase, n, HB, PS, sm, a, v, tid, nth);
  const float* ps = QG_SMEM_F + sm.ps_off;
  const float* tmp = QG_SMEM_F + sm.tmp_off;
  const size_t t = (size_t)(tbase - a.t0);
  for (int c = 0; c < a.n_out; c++) {
    const int ox = a.out_x[c];
    float* o = a.frame_major ? a.out + (t * a.V + v) * a.n_out + c : a.out + ((size_t)v * a.n_out + c) * a.T + t;
    const size_t stride = a.frame_major ? (size_t)a.V * a.n_out : 1;
    if (ox < PS) { const float k = ps[ox]; for (int j = tid; j < n; j += nth) o[j * stride] = k; }
    else { const float* src = tmp + (ox - PS) * HB; for (int j = tid; j < n; j += nth) o[j * stride] = src[j]; }
  }
}

// ------------------------------------------------------------------------------------------------ launcher
// Y ring: a power of two of at least two rounds.  Small banks take more rounds per launch to fill the machine.
size_t spectral_y_bytes(const SpPlan& p, long V, int* ring) {
  long mult = 2;
  while (mult < 64 && V * (long)p.items.size() * (mult - 1) < 2048 &&
         (size_t)p.n_streams * (size_t)V * (size_t)(2 * mult) * p.C * 4 <= ((size_t)256 << 20)) mult *= 2;
  // full-size banks: four rounds of ring = three rounds per launch (measured on configs[3]: 103.8 -> 98.7 ms; the launch
  // boundaries drain the machine), as long as the ring stays around the size of the L2
  if (mult < 4 && (size_t)p.n_streams * (size_t)V * (size_t)4 * p.C * 4 <= ((size_t)160 << 20)) mult = 4;
  if (const char* e = getenv("QG_SPECTRAL_RING_MULT")) { const long m = atol(e); if (m >= 2 && m <= 64 && (m & (m - 1)) == 0) mult = m; }
  if (ring) *ring = (int)(mult * p.C);
  return (size_t)p.n_streams * (size_t)V * (size_t)(mult * p.C) * sizeof(float);
}

static long floor_div(long a, long b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

cudaError_t launch_spectral(const SpArgs& a, const SpPlan& p, cudaStream_t stream, int* launches, cudaKernel_t spec_frames,
                            cudaKernel_t spec_post) {
  const int C = p.C, PS = a.P + a.NS;
  int max_n = 0;
  for (const SpSegment& s : p.segs) max_n = std::max(max_n, 1 << s.lg);
  const int HBf = max_n / 2 + 1, HBp = 512;
  SpSmem sf, sp;
  sf.ps_off = sp.ps_off = a.n_code * (int)(sizeof(Instr) / 4);
  sf.tmp_off = sp.tmp_off = sf.ps_off + ((PS + 3) & ~3);
  sf.f_off = (sf.tmp_off + std::max(2, p.n_slots_frame) * HBf + 3) & ~3;   // even: the transform buffers are float2
  sp.f_off = 0;
  const size_t smem_f = (size_t)sf.f_off * 4 + 2 * (size_t)(CPAD(max_n) + 1) * 8, smem_p = (size_t)(sp.tmp_off + std::max(1, p.n_slots_post) * HBp) * 4;
  const bool spec = spec_frames && spec_post;                 // K5s: kernels compiled for this plan (spectral_kernel.cuh)
  const size_t smem_s = 2 * (size_t)(CPAD(max_n) + 1) * 8;    // their only shared memory: the two transform buffers
  const int HBs = 1024;                                       // SP_POST_BLOCK
  cudaError_t e;
  if (spec) {
    if (smem_s > 48 * 1024) {
      e = cudaFuncSetAttribute((const void*)spec_frames, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s);
      if (e != cudaSuccess) return e;
    }
  } else {
    if (smem_f > 200 * 1024 || smem_p > 200 * 1024) return cudaErrorNotSupported;
    e = cudaFuncSetAttribute(k_spectral_frames, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_spectral_post, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p);
    if (e != cudaSuccess) return e;
  }
  const long A = a.t0, E = a.t0 + a.T;
  const long c_first = floor_div(A, C) - 1, c_last = floor_div(E - 1, C);
  const int R = a.ring / C - 1;     // rounds per launch
  cudaError_t le = cudaSuccess;
  auto frames = [&](long c, int n) {
    const long blocks = (long)a.V * a.n_items * n;
    if (spec) {
      SpArgs aa = a;
      void* args[] = {&aa, &c, &n};
      cudaError_t r = cudaLaunchKernel((const void*)spec_frames, dim3((unsigned)blocks), dim3(256), args, smem_s, stream);
      if (r != cudaSuccess) le = r;
    } else {
      k_spectral_frames<<<(unsigned)blocks, 256, smem_f, stream>>>(a, c, n, HBf, sf);
    }
    if (launches) *launches += 1;
  };
  frames(c_first, 1);
  for (long c = c_first + 1; c <= c_last; c += R) {
    const int n = (int)std::min<long>(R, c_last - c + 1);
    frames(c, n);
    const long t_lo = std::max(A, c * (long)C), t_hi = std::min(E, (c + n) * (long)C);
    if (t_hi > t_lo) {
      if (spec) {
        const long nblk = (t_hi - t_lo + HBs - 1) / HBs;
        SpArgs aa = a;
        long lo_ = t_lo, hi_ = t_hi;
        void* args[] = {&aa, &lo_, &hi_};
        cudaError_t r = cudaLaunchKernel((const void*)spec_post, dim3((unsigned)(a.V * nblk)), dim3(256), args, 0, stream);
        if (r != cudaSuccess) le = r;
      } else {
        const long nblk = (t_hi - t_lo + HBp - 1) / HBp;
        k_spectral_post<<<(unsigned)(a.V * nblk), 256, smem_p, stream>>>(a, t_lo, t_hi, HBp, sp);
      }
      if (launches) *launches += 1;
    }
  }
  if (le != cudaSuccess) return le;
  return cudaGetLastError();
}

}  // namespace qg
