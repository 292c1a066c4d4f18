// Fused gate-sweep kernels (K1/K2/K6 of SURVEY.md §2.1) for sm_100a.
//
// One launch = one "sweep": every CTA stages a tile of 2^m amplitudes of one batch element in
// shared memory (coalesced, vectorised 16-byte global accesses), then runs several "rounds".  In a
// round each thread pulls 2^r amplitudes into registers (the r "register bits" of that round),
// applies every gate of the round entirely in registers, and writes them back to the XOR-swizzled
// tile.  After the last round the tile is written back: 2·S bytes of HBM traffic for all gates of
// the sweep instead of 2·S per gate (the reference's einsum-per-gate loop,
// /root/reference/qadence/backends/pyqtorch/backend.py:176).
//
// Per-CTA prologue ("resolve"): parametric gates evaluate sincos in-kernel from the parameter
// table of THIS batch element, and chains of 1-qubit gates on the same target are multiplied into a
// single 2x2 (what `pyq.Merge` does for single-qubit chains, convert_ops.py:277-278, but per CTA and
// for any chain the planner finds).
//
// The adjoint variant (ADJ) walks the same records backwards on ψ and λ simultaneously.  For every
// chain it reduces the four moments  m_P = Im<λ|P|ψ>, P ∈ {I, X, Y, Z}  of the POST-chain states on
// the control-active subspace, un-applies the merged 2x2 from both states, and the epilogue turns
// the moments into the gradient of every parametric gate of the chain:
//   g_i = 2 Re<λ_i| ∂G_i G_i† |ψ_i> = c_I m_I + (W_i c W_i†)·m,  W_i = G_k ··· G_{i+1},
// with ∂G G† = −(i/2)(c_I I + c·σ): RX c=x̂, RY c=ŷ, RZ c=ẑ, PHASE c_I=−1, c=ẑ.   (4·S bytes.)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "b200sv.h"

namespace b200sv {

template <typename real> struct C2;
template <> struct C2<double> { using type = double2; };
template <> struct C2<float> { using type = float2; };

template <typename T> __device__ __forceinline__ T mk(decltype(T::x) x, decltype(T::x) y) {
  T r; r.x = x; r.y = y; return r;
}
template <typename T> __device__ __forceinline__ T cmul(T a, T b) {
  return mk<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

constexpr int RT_STRIDE = 40;  // words per per-round index table

// flat opcodes of resolved gates: base + register index
enum {
  OP_NOP = 0, OP_SCALE = 1, OP_DIAGEXT = 2,
  OP_SWAP = 4,   // + pair index (a<b): 01,02,12,03,13,23 -> 0..5
  OP_MAT = 16, OP_REAL = 24, OP_RX = 32, OP_DIAG = 40, OP_X = 48
};
enum { GR_NONE = 0, GR_CHAIN = 1, GR_SCALE = 2 };

template <typename real> struct __align__(16) ExecGate {
  real m[8];
  unsigned long long ctrl_ext;
  int op, ext_bit, cr, gr;
};

template <typename real> __device__ __forceinline__ int swz(int e);
template <> __device__ __forceinline__ int swz<double>(int e) {
  return e ^ (((e >> 3) ^ (e >> 6) ^ (e >> 9)) & 7);
}
template <> __device__ __forceinline__ int swz<float>(int e) {
  return e ^ (((e >> 4) ^ (e >> 8)) & 15);
}

// ------------------------------------------------------------------------------------------------
// 2x2 complex helpers in double (prologue / epilogue only)
// ------------------------------------------------------------------------------------------------
struct M22 { double v[8]; };  // m00 m01 m10 m11 as (re, im)

__device__ __forceinline__ M22 m22_mul(const M22& a, const M22& b) {  // a * b
  M22 r;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      double re = 0, im = 0;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const double ar = a.v[2 * (2 * i + k)], ai = a.v[2 * (2 * i + k) + 1];
        const double br = b.v[2 * (2 * k + j)], bi = b.v[2 * (2 * k + j) + 1];
        re += ar * br - ai * bi; im += ar * bi + ai * br;
      }
      r.v[2 * (2 * i + j)] = re; r.v[2 * (2 * i + j) + 1] = im;
    }
  return r;
}
__device__ __forceinline__ M22 m22_dag(const M22& a) {
  M22 r;
  r.v[0] = a.v[0]; r.v[1] = -a.v[1]; r.v[2] = a.v[4]; r.v[3] = -a.v[5];
  r.v[4] = a.v[2]; r.v[5] = -a.v[3]; r.v[6] = a.v[6]; r.v[7] = -a.v[7];
  return r;
}
__device__ __forceinline__ M22 m22_eye() { M22 r = {{1, 0, 0, 0, 0, 0, 1, 0}}; return r; }

__device__ __forceinline__ double gate_theta(const b200sv_gate& g, const double* __restrict__ params,
                                             const double* __restrict__ consts, long long batch, long long b,
                                             double* im) {
  if (g.slot >= 0) { *im = 0.0; return params[(long long)g.slot * batch + b]; }
  *im = consts[2 * g.cidx + 1];
  return consts[2 * g.cidx];
}

// forward 2x2 of an elementary 1-qubit record
__device__ M22 elem_matrix(const b200sv_gate& g, const double* __restrict__ params,
                           const double* __restrict__ consts, long long batch, long long b) {
  M22 r = m22_eye();
  double th_im;
  switch (g.kind) {
    case B200SV_GK_MAT: { const double* c = consts + 2 * g.cidx; for (int i = 0; i < 8; ++i) r.v[i] = c[i]; } break;
    case B200SV_GK_REAL: { const double* c = consts + 2 * g.cidx; r.v[0] = c[0]; r.v[2] = c[1]; r.v[4] = c[2]; r.v[6] = c[3]; } break;
    case B200SV_GK_RX: { double s, c; sincos(0.5 * gate_theta(g, params, consts, batch, b, &th_im), &s, &c);
      r.v[0] = c; r.v[6] = c; r.v[2] = 0; r.v[3] = -s; r.v[4] = 0; r.v[5] = -s; } break;
    case B200SV_GK_RY: { double s, c; sincos(0.5 * gate_theta(g, params, consts, batch, b, &th_im), &s, &c);
      r.v[0] = c; r.v[2] = -s; r.v[4] = s; r.v[6] = c; } break;
    case B200SV_GK_RZ: { double s, c; sincos(0.5 * gate_theta(g, params, consts, batch, b, &th_im), &s, &c);
      r.v[0] = c; r.v[1] = -s; r.v[6] = c; r.v[7] = s; } break;
    case B200SV_GK_PHASE: { double s, c; sincos(gate_theta(g, params, consts, batch, b, &th_im), &s, &c);
      r.v[6] = c; r.v[7] = s; } break;
    case B200SV_GK_DIAG: { const double* c = consts + 2 * g.cidx; r.v[0] = c[0]; r.v[1] = c[1]; r.v[6] = c[2]; r.v[7] = c[3]; } break;
    case B200SV_GK_X: r.v[0] = 0; r.v[6] = 0; r.v[2] = 1; r.v[4] = 1; break;
    default: break;
  }
  return r;
}

// ------------------------------------------------------------------------------------------------
// per-CTA resolution of one exec entry (chain leader record + followers) into an ExecGate
// ------------------------------------------------------------------------------------------------
template <typename real>
__device__ void resolve_gate(const b200sv_gate* __restrict__ recs, ExecGate<real>& e,
                             const double* __restrict__ params, const double* __restrict__ consts,
                             long long batch, long long b, int dir, bool adj) {
  const b200sv_gate& g = recs[0];
  e.ctrl_ext = g.ctrl_ext;
  e.ext_bit = g.ext_bit; e.cr = g.ctrl_reg;
  e.gr = GR_NONE;
  const int xk = g.xkind;
  if (xk == B200SV_XK_X) { e.op = OP_X + g.a; return; }
  if (xk == B200SV_XK_SWAP) {
    int a = g.a, c = g.a2; if (a > c) { const int t = a; a = c; c = t; }
    e.op = OP_SWAP + (c * (c - 1)) / 2 + a; return;
  }
  if (xk == B200SV_XK_SCALE) {
    double ti; const double tr = gate_theta(g, params, consts, batch, b, &ti);
    if (dir > 0) { e.m[0] = (real)tr; e.m[1] = (real)ti; }
    else {  // psi <- psi / s (inverse), lambda <- conj(s) lambda (adjoint)
      const double d = tr * tr + ti * ti;
      e.m[0] = (real)(tr / d); e.m[1] = (real)(-ti / d); e.m[2] = (real)tr; e.m[3] = (real)-ti;
    }
    e.op = OP_SCALE;
    if (adj && (g.flags & B200SV_GF_GRAD)) e.gr = GR_SCALE;
    return;
  }
  // 1-qubit chain: product of `chain` elementary matrices, first record applied first
  M22 prod = elem_matrix(g, params, consts, batch, b);
  bool any_grad = (g.flags & B200SV_GF_GRAD) != 0;
  for (int i = 1; i < g.chain; ++i) {
    prod = m22_mul(elem_matrix(recs[i], params, consts, batch, b), prod);
    any_grad |= (recs[i].flags & B200SV_GF_GRAD) != 0;
  }
  if (dir < 0) prod = m22_dag(prod);
  if (adj && any_grad) e.gr = GR_CHAIN;
  if (g.flags & B200SV_GF_MSLOT) {  // unrolled matrix phase: always a general 2x2
#pragma unroll
    for (int i = 0; i < 8; ++i) e.m[i] = (real)prod.v[i];
    e.op = OP_MAT + g.a;
    return;
  }
  switch (xk) {
    case B200SV_XK_RX: e.m[0] = (real)prod.v[0]; e.m[1] = (real)-prod.v[3]; e.op = OP_RX + g.a; break;  // [[c,-is],[-is,c]]
    case B200SV_XK_REAL: e.m[0] = (real)prod.v[0]; e.m[1] = (real)prod.v[2]; e.m[2] = (real)prod.v[4]; e.m[3] = (real)prod.v[6];
      e.op = OP_REAL + g.a; break;
    case B200SV_XK_DIAG: e.m[0] = (real)prod.v[0]; e.m[1] = (real)prod.v[1]; e.m[2] = (real)prod.v[6]; e.m[3] = (real)prod.v[7];
      e.op = g.a >= 0 ? OP_DIAG + g.a : OP_DIAGEXT; break;
    default:
#pragma unroll
      for (int i = 0; i < 8; ++i) e.m[i] = (real)prod.v[i];
      e.op = OP_MAT + g.a; break;
  }
}

// epilogue: turn the four moments of a chain into per-gate gradients
__device__ void chain_gradients(const b200sv_gate* __restrict__ recs, const double* mom,
                                const double* __restrict__ params, const double* __restrict__ consts,
                                double* __restrict__ grads, long long batch, long long b) {
  const int k = recs[0].chain;
  M22 W = m22_eye();
  for (int i = k - 1; i >= 0; --i) {
    const b200sv_gate& g = recs[i];
    if ((g.flags & B200SV_GF_GRAD) && g.grad_row >= 0) {
      // P = c·σ in {X, Y, Z};  A = W P W†  (traceless Hermitian) = [[az, ax - i ay],[ax + i ay, -az]]
      M22 P = {{0, 0, 0, 0, 0, 0, 0, 0}};
      double cI = 0.0;
      if (g.kind == B200SV_GK_RX) { P.v[2] = 1; P.v[4] = 1; }
      else if (g.kind == B200SV_GK_RY) { P.v[3] = -1; P.v[5] = 1; }
      else { P.v[0] = 1; P.v[6] = -1; if (g.kind == B200SV_GK_PHASE) cI = -1.0; }
      const M22 A = m22_mul(m22_mul(W, P), m22_dag(W));
      const double az = A.v[0], ax = A.v[4], ay = A.v[5];
      const double gsum = cI * mom[0] + ax * mom[1] + ay * mom[2] + az * mom[3];
      atomicAdd(&grads[(long long)g.grad_row * batch + b], gsum);
    }
    if (i > 0) W = m22_mul(W, elem_matrix(g, params, consts, batch, b));
  }
}

// ------------------------------------------------------------------------------------------------
// in-register gate application: v[j], j in [0, 2^R); gate target = register bit A
// ------------------------------------------------------------------------------------------------
template <typename real, int R> struct Regs {
  using c2 = typename C2<real>::type;
  static constexpr int NV = 1 << R;

  template <int A, bool CK> static __device__ __forceinline__ void mat_(c2 (&v)[NV], const real* m, int cr) {
    const real m0 = m[0], m1 = m[1], m2 = m[2], m3 = m[3], m4 = m[4], m5 = m[5], m6 = m[6], m7 = m[7];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (j & (1 << A)) continue;
      if (CK && (j & cr) != cr) continue;
      const c2 x = v[j], y = v[j | (1 << A)];
      v[j] = mk<c2>(m0 * x.x - m1 * x.y + m2 * y.x - m3 * y.y, m0 * x.y + m1 * x.x + m2 * y.y + m3 * y.x);
      v[j | (1 << A)] = mk<c2>(m4 * x.x - m5 * x.y + m6 * y.x - m7 * y.y, m4 * x.y + m5 * x.x + m6 * y.y + m7 * y.x);
    }
  }
  template <int A, bool CK> static __device__ __forceinline__ void real_(c2 (&v)[NV], const real* m, int cr) {
    const real m0 = m[0], m1 = m[1], m2 = m[2], m3 = m[3];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (j & (1 << A)) continue;
      if (CK && (j & cr) != cr) continue;
      const c2 x = v[j], y = v[j | (1 << A)];
      v[j] = mk<c2>(m0 * x.x + m1 * y.x, m0 * x.y + m1 * y.y);
      v[j | (1 << A)] = mk<c2>(m2 * x.x + m3 * y.x, m2 * x.y + m3 * y.y);
    }
  }
  template <int A, bool CK> static __device__ __forceinline__ void rx_(c2 (&v)[NV], const real* m, int cr) {
    const real c = m[0], s = m[1];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (j & (1 << A)) continue;
      if (CK && (j & cr) != cr) continue;
      const c2 x = v[j], y = v[j | (1 << A)];
      v[j] = mk<c2>(c * x.x + s * y.y, c * x.y - s * y.x);  // [[c, -i s], [-i s, c]]
      v[j | (1 << A)] = mk<c2>(c * y.x + s * x.y, c * y.y - s * x.x);
    }
  }
  template <int A, bool CK> static __device__ __forceinline__ void diag_(c2 (&v)[NV], const real* m, int cr) {
    const c2 d0 = mk<c2>(m[0], m[1]), d1 = mk<c2>(m[2], m[3]);
    const bool one0 = (d0.x == (real)1 && d0.y == (real)0);  // PHASE / CZ / CPHASE: |0> branch untouched
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (CK && (j & cr) != cr) continue;
      if (j & (1 << A)) v[j] = cmul(v[j], d1);
      else if (!one0) v[j] = cmul(v[j], d0);
    }
  }
  template <int A, bool CK> static __device__ __forceinline__ void x_(c2 (&v)[NV], const real*, int cr) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (j & (1 << A)) continue;
      if (CK && (j & cr) != cr) continue;
      const c2 x = v[j]; v[j] = v[j | (1 << A)]; v[j | (1 << A)] = x;
    }
  }
  template <int A, int B> static __device__ __forceinline__ void swap_(c2 (&v)[NV], int cr) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if ((j & (1 << A)) || !(j & (1 << B))) continue;  // j: bit A = 0, bit B = 1
      if ((j & cr) != cr) continue;
      const int k = (j | (1 << A)) & ~(1 << B);
      const c2 x = v[j]; v[j] = v[k]; v[k] = x;
    }
  }
  static __device__ __forceinline__ void scale_all(c2 (&v)[NV], c2 d, int cr) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if ((j & cr) != cr) continue;
      v[j] = cmul(v[j], d);
    }
  }

#define B200SV_CASES(BASE, FN)                                                                      \
  case BASE + 0: if (cr) FN<0, true>(v, e.m, cr); else FN<0, false>(v, e.m, 0); break;              \
  case BASE + 1: if constexpr (R > 1) { if (cr) FN<1, true>(v, e.m, cr); else FN<1, false>(v, e.m, 0); } break; \
  case BASE + 2: if constexpr (R > 2) { if (cr) FN<2, true>(v, e.m, cr); else FN<2, false>(v, e.m, 0); } break; \
  case BASE + 3: if constexpr (R > 3) { if (cr) FN<3, true>(v, e.m, cr); else FN<3, false>(v, e.m, 0); } break; \
  case BASE + 4: if constexpr (R > 4) { if (cr) FN<4, true>(v, e.m, cr); else FN<4, false>(v, e.m, 0); } break;

  // apply one resolved gate to a register file (thread already known to satisfy ctrl_ext)
  static __device__ __forceinline__ void apply(c2 (&v)[NV], const ExecGate<real>& e,
                                               unsigned long long gthr, bool is_lambda) {
    const int cr = e.cr;
    switch (e.op) {
      B200SV_CASES(OP_MAT, mat_)
      B200SV_CASES(OP_REAL, real_)
      B200SV_CASES(OP_RX, rx_)
      B200SV_CASES(OP_DIAG, diag_)
      B200SV_CASES(OP_X, x_)
      case OP_SWAP + 0: if constexpr (R > 1) swap_<0, 1>(v, cr); break;
      case OP_SWAP + 1: if constexpr (R > 2) swap_<0, 2>(v, cr); break;
      case OP_SWAP + 2: if constexpr (R > 2) swap_<1, 2>(v, cr); break;
      case OP_SWAP + 3: if constexpr (R > 3) swap_<0, 3>(v, cr); break;
      case OP_SWAP + 4: if constexpr (R > 3) swap_<1, 3>(v, cr); break;
      case OP_SWAP + 5: if constexpr (R > 3) swap_<2, 3>(v, cr); break;
      case OP_DIAGEXT: {
        const int bit = (int)((gthr >> e.ext_bit) & 1ull);
        const c2 d = bit ? mk<c2>(e.m[2], e.m[3]) : mk<c2>(e.m[0], e.m[1]);
        if (!(d.x == (real)1 && d.y == (real)0)) scale_all(v, d, cr);
      } break;
      case OP_SCALE: scale_all(v, is_lambda ? mk<c2>(e.m[2], e.m[3]) : mk<c2>(e.m[0], e.m[1]), 0); break;
      default: break;
    }
  }

  // ---- adjoint: four moments Im<λ|P|ψ>, P = I, X, Y, Z, on the control-active subspace ----------
  template <int A>
  static __device__ __forceinline__ void moments_(const c2 (&p)[NV], const c2 (&l)[NV], int cr, double (&mo)[4]) {
    // the kernel is issue-bound (an FP64 instruction occupies the issue port for two cycles): pure FMA
    // chains, 12 per amplitude pair, no separate adds
    double s00 = 0, s11 = 0, x_im = 0, y_re = 0;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (j & (1 << A)) continue;
      if ((j & cr) != cr) continue;
      const int k = j | (1 << A);
      const double p0x = p[j].x, p0y = p[j].y, p1x = p[k].x, p1y = p[k].y;
      const double l0x = l[j].x, l0y = l[j].y, l1x = l[k].x, l1y = l[k].y;
      s00 = fma(l0x, p0y, s00); s00 = fma(-l0y, p0x, s00);      // Im conj(l0) p0
      s11 = fma(l1x, p1y, s11); s11 = fma(-l1y, p1x, s11);      // Im conj(l1) p1
      x_im = fma(l0x, p1y, x_im); x_im = fma(-l0y, p1x, x_im);  // Im<λ|X|ψ> = Im(conj(l0) p1 + conj(l1) p0)
      x_im = fma(l1x, p0y, x_im); x_im = fma(-l1y, p0x, x_im);
      y_re = fma(l1x, p0x, y_re); y_re = fma(l1y, p0y, y_re);   // Im<λ|Y|ψ> = Re(conj(l1) p0 - conj(l0) p1)
      y_re = fma(-l0x, p1x, y_re); y_re = fma(-l0y, p1y, y_re);
    }
    mo[0] = s00 + s11; mo[1] = x_im; mo[2] = y_re; mo[3] = s00 - s11;
  }
  static __device__ __forceinline__ void moments(const c2 (&p)[NV], const c2 (&l)[NV], const ExecGate<real>& e,
                                                 unsigned long long gthr, double (&mo)[4]) {
    const int cr = e.cr;
    mo[0] = mo[1] = mo[2] = mo[3] = 0.0;
    if (e.op == OP_DIAGEXT) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        if ((j & cr) != cr) continue;
        s += (double)l[j].x * (double)p[j].y - (double)l[j].y * (double)p[j].x;
      }
      mo[0] = s; mo[3] = ((gthr >> e.ext_bit) & 1ull) ? -s : s;
      return;
    }
    switch (e.op & 7) {
      case 0: moments_<0>(p, l, cr, mo); break;
      case 1: if constexpr (R > 1) moments_<1>(p, l, cr, mo); break;
      case 2: if constexpr (R > 2) moments_<2>(p, l, cr, mo); break;
      case 3: if constexpr (R > 3) moments_<3>(p, l, cr, mo); break;
      default: break;
    }
  }
  // ---- straight-line matrix phases, specialised on the set MASK of occupied register slots ----------
  template <int MASK>
  static __device__ __forceinline__ void mphase_fwd(c2 (&p)[NV], const ExecGate<real>* eg, const int* ms) {
    if constexpr (MASK & 1) mat_<0, false>(p, eg[ms[0]].m, 0);
    if constexpr (R > 1 && (MASK & 2)) mat_<1, false>(p, eg[ms[1]].m, 0);
    if constexpr (R > 2 && (MASK & 4)) mat_<2, false>(p, eg[ms[2]].m, 0);
    if constexpr (R > 3 && (MASK & 8)) mat_<3, false>(p, eg[ms[3]].m, 0);
  }
  template <int MASK, int A, bool WITH_L, typename RED>
  static __device__ __forceinline__ void mslot_bwd(c2 (&p)[NV], c2 (&l)[WITH_L ? NV : 1], const ExecGate<real>* eg, const int* ms, RED& reduce) {
    if constexpr (A < R && ((MASK >> A) & 1)) {
      const ExecGate<real>& e = eg[ms[A]];
      if constexpr (WITH_L) {
        if (e.gr == GR_CHAIN) {  // uniform across the CTA
          double mo[4];
          moments_<A>(p, l, 0, mo);
          reduce(mo, 4, ms[A]);
        }
      }
      mat_<A, false>(p, e.m, 0);
      if constexpr (WITH_L) mat_<A, false>(l, e.m, 0);
    }
  }
  template <int MASK, bool WITH_L, typename RED>
  static __device__ __forceinline__ void mphase_bwd(c2 (&p)[NV], c2 (&l)[WITH_L ? NV : 1], const ExecGate<real>* eg, const int* ms, RED& reduce) {
    mslot_bwd<MASK, 3, WITH_L>(p, l, eg, ms, reduce);
    mslot_bwd<MASK, 2, WITH_L>(p, l, eg, ms, reduce);
    mslot_bwd<MASK, 1, WITH_L>(p, l, eg, ms, reduce);
    mslot_bwd<MASK, 0, WITH_L>(p, l, eg, ms, reduce);
  }

  static __device__ __forceinline__ double grad_scale(const c2 (&p)[NV], const c2 (&l)[NV]) {
    double acc = 0.0;  // 2 Re <λ|ψ_before>
#pragma unroll
    for (int j = 0; j < NV; ++j) acc += (double)l[j].x * (double)p[j].x + (double)l[j].y * (double)p[j].y;
    return 2.0 * acc;
  }
};

__device__ __forceinline__ int byte_of(unsigned long long w, int i) { return (int)((w >> (8 * i)) & 0xffull); }

// ------------------------------------------------------------------------------------------------
// the sweep kernel
// ------------------------------------------------------------------------------------------------
#ifndef B200SV_MINB_FWD_D
#define B200SV_MINB_FWD_D 4  // complex128 forward, r >= 3: CTAs per SM the register allocation targets
#endif
#ifndef B200SV_ADJ_D_MAXT
#define B200SV_ADJ_D_MAXT 128  // threads per CTA of the complex128 adjoint kernel (r = 3)
#endif
#ifndef B200SV_MINB_FWD_D_R4
#define B200SV_MINB_FWD_D_R4 2
#endif
#ifndef B200SV_MINB_FWD_D_R2
#define B200SV_MINB_FWD_D_R2 3
#endif
#ifndef B200SV_MINB_ADJ_D_R2
#define B200SV_MINB_ADJ_D_R2 4
#endif
#ifndef B200SV_MINB_ADJ_D
#define B200SV_MINB_ADJ_D 4  // complex128 adjoint (128-thread CTAs)
#endif
// launch geometry per instantiation: (max threads per CTA, min CTAs per SM)
template <typename real, int R, bool ADJ> struct SweepCfg {
  static constexpr int MAXT = (ADJ && sizeof(real) == 8 && R >= 3) ? B200SV_ADJ_D_MAXT : 256;
  static constexpr int MINB = (ADJ && sizeof(real) == 8) ? (R <= 2 ? B200SV_MINB_ADJ_D_R2 : B200SV_MINB_ADJ_D)
                              : ((sizeof(real) == 8 && R >= 4) ? B200SV_MINB_FWD_D_R4 : (sizeof(real) == 8 && R == 3) ? B200SV_MINB_FWD_D : ((sizeof(real) == 8 && R == 2) ? B200SV_MINB_FWD_D_R2 : (ADJ ? 2 : 3)));
};

// 16-byte (complex128) / 8-byte (complex64) asynchronous global -> shared copies (LDGSTS)
template <int BYTES> __device__ __forceinline__ void cp_async(void* smem, const void* gmem);
template <> __device__ __forceinline__ void cp_async<16>(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
template <> __device__ __forceinline__ void cp_async<8>(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// Persistent kernel: each CTA owns a contiguous range of tiles and double-buffers them — the
// cp.async loads of tile k+1 are in flight while tile k is processed.
// GEN = false: the sweep has no generic-phase entries (pure merged-2x2 + permutation rounds, e.g.
// every sweep of an HEA): the interpreter is compiled out, which shrinks registers and code.
// PEER = true: the register is sharded over several GPUs by its highest-order qubits and this kernel addresses ALL
// shards directly: amplitude index bits >= peers.n_local select the owning GPU's buffer (NVLink peer memory, mapped
// with CUDA IPC), the rest is the offset inside it.  A gate on a "global" qubit is then just a sweep whose tiles
// straddle shards — the exchange is fused into the sweep's own tile loads and stores (no separate all-to-all pass,
// no staging buffer).  Every tile is read and written by exactly one CTA of one GPU, so there is no intra-sweep
// hazard; sweeps are separated by a cross-GPU barrier on the stream.
struct PeerTable {
  void* p[16];            // state buffers of ranks 0..n_peers-1 (p[rank] is the local one)
  void* pl[16];           // adjoint: the lambda shards of the same ranks
  int n_local;            // index bits below this are the offset inside a shard
  long long tile_offset;  // this GPU processes tiles [tile_offset, tile_offset + n_tiles) of the sweep
  long long n_tiles;
};

template <typename real, int R, bool ADJ, bool GEN, bool PEER = false>
__global__ void __launch_bounds__(SweepCfg<real, R, ADJ>::MAXT, SweepCfg<real, R, ADJ>::MINB)
sweep_kernel(typename C2<real>::type* __restrict__ psi_g, typename C2<real>::type* __restrict__ lam_g,
             const double* __restrict__ params, double* __restrict__ grads,
             const b200sv_sweep* __restrict__ sweeps, const b200sv_round* __restrict__ rounds,
             const b200sv_gate* __restrict__ gates, const double* __restrict__ consts,
             const ExecGate<real>* __restrict__ xtab, double* __restrict__ mom_g, int sweep_idx,
             int n, long long batch, unsigned long long index_hi, int dir_dbg, long long tiles_per_cta, int nbuf,
             const PeerTable peers) {
  constexpr int dir = ADJ ? -1 : +1;  // the adjoint kernel always walks backwards, the plain one forwards
  const int dbg = dir_dbg & 0xff;  // timing experiments only (B200SV_DBG): 1 = skip math, 2 = skip rounds
  using c2 = typename C2<real>::type;
  using RG = Regs<real, R>;
  constexpr int NV = 1 << R;
  extern __shared__ __align__(16) unsigned char smem_raw[];

  const b200sv_sweep& sw = sweeps[sweep_idx];
  const int m = sw.m;
  const int tile = 1 << m;
  const int nt = blockDim.x;  // == 2^(m-R)
  const int tid = threadIdx.x;
  const int n_exec = sw.n_exec;
  const int n_rounds = sw.n_rounds;
  constexpr int NSTATE = ADJ ? 2 : 1;

  // shared memory: 2 buffers x (psi [, lambda]) tiles | offset tables | resolved gates | moments | round tables
  c2* bufs = reinterpret_cast<c2*>(smem_raw);
  unsigned long long* t_lo = reinterpret_cast<unsigned long long*>(bufs + nbuf * NSTATE * tile);
  unsigned long long* t_hi = t_lo + 64;
  ExecGate<real>* eg = reinterpret_cast<ExecGate<real>*>(t_hi + 64);
  double* gsum = reinterpret_cast<double*>(eg + n_exec);  // ADJ: [n_exec][nwarps][4] moment partials
  const int nwarps = (nt + 31) >> 5;
  // per-round index tables (RT_STRIDE words each): [0,16) low thread nibble, [16,32) high nibble,
  // [32,32+R) register bits, [36] per-tile external-control constant; each word =
  // swz(plain tile index part) | swz(permuted part) << 16
  unsigned* rtab = reinterpret_cast<unsigned*>(gsum + (ADJ ? 4 * n_exec * nwarps : 0));

  const long long tiles_per_state = 1ll << sw.n_outer;
  const long long total_tiles = PEER ? peers.n_tiles : (batch << sw.n_outer);
  const long long tile_off = PEER ? peers.tile_offset : 0;
  const unsigned long long local_mask = PEER ? ((1ull << peers.n_local) - 1ull) : ~0ull;
  auto gaddr = [&](c2* base_ptr, unsigned long long idx) -> c2* {  // basis index -> address (through the peer table)
    if constexpr (PEER) return reinterpret_cast<c2*>(peers.p[idx >> peers.n_local]) + (idx & local_mask);
    else return base_ptr + idx;
  };
  auto laddr = [&](c2* base_ptr, unsigned long long idx) -> c2* {  // the same for lambda
    if constexpr (PEER) return reinterpret_cast<c2*>(peers.pl[idx >> peers.n_local]) + (idx & local_mask);
    else return base_ptr + idx;
  };
  const long long tile_first = (long long)blockIdx.x * tiles_per_cta;
  long long tile_end = tile_first + tiles_per_cta;
  if (tile_end > total_tiles) tile_end = total_tiles;

  // ---- one-time prologue: offset tables (global offsets of tile indices, per-round index tables) ----
  const int nthr_bits = m - R;
  for (int v = tid; v < 64; v += nt) {
    unsigned long long lo = 0, hi = 0;
    for (int i = 0; i < 6; ++i) {
      if (i < m && ((v >> i) & 1)) lo |= 1ull << sw.tile_bits[i];
      if (i + 6 < m && ((v >> i) & 1)) hi |= 1ull << sw.tile_bits[i + 6];
    }
    t_lo[v] = lo; t_hi[v] = hi;
  }
  for (int w = tid; w < n_rounds * RT_STRIDE; w += nt) {
    const int rloc = w / RT_STRIDE, i = w % RT_STRIDE;
    const b200sv_round& rd = rounds[sw.first_round + rloc];
    const bool perm = (rd.flags & B200SV_RF_PERM) != 0;
    unsigned plain = 0, permd = 0;
    if (i < 32) {
      const int half = i >> 4;
      for (int k = 0; k < 4; ++k) {
        const int tb = k + 4 * half;
        if (tb < nthr_bits && ((i >> k) & 1)) {
          const int pos = rd.thrpos[tb];
          plain |= 1u << pos;
          permd ^= perm ? (unsigned)rd.pcol[pos] : (1u << pos);
        }
      }
      if (half == 0 && perm) permd ^= rd.pconst;
    } else if (i - 32 < R) {
      const int pos = rd.regpos[i - 32];
      plain = 1u << pos;
      permd = perm ? (unsigned)rd.pcol[pos] : (1u << pos);
    }
    rtab[w] = (unsigned)swz<real>((int)plain) | ((unsigned)swz<real>((int)permd) << 16);
  }
  if (ADJ) for (int g = tid; g < 4 * n_exec * nwarps; g += nt) gsum[g] = 0.0;
  __syncthreads();

  auto tile_base = [&](long long t) -> unsigned long long {  // CTA-index bits of tile t deposited into the basis index
    const unsigned long long c = (unsigned long long)((t + tile_off) % tiles_per_state);
    unsigned long long base = 0;
    if (sw.n_seg >= 0) {  // runs of consecutive outer bits: a few shifts instead of a bit loop
      for (int k = 0; k < sw.n_seg; ++k)
        base |= ((c >> sw.seg[k].src) & ((1ull << sw.seg[k].len) - 1ull)) << sw.seg[k].dst;
    } else {
      for (int k = 0; k < sw.n_outer; ++k) base |= ((c >> k) & 1ull) << sw.outer_bits[k];
    }
    return base;
  };
  auto issue_loads = [&](long long t, unsigned long long base, int which) {  // asynchronous coalesced global -> swizzled shared
    const long long off = ((t / tiles_per_state) << n) + (long long)base;
    c2* dst = bufs + which * NSTATE * tile;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int e = i * nt + tid;
      const long long g = off + (long long)(t_lo[e & 63] + t_hi[e >> 6]);
      cp_async<sizeof(c2)>(dst + swz<real>(e), gaddr(psi_g, (unsigned long long)g));
      if constexpr (ADJ) cp_async<sizeof(c2)>(dst + tile + swz<real>(e), laddr(lam_g, (unsigned long long)g));
    }
    cp_async_commit();
  };
  auto flush_moments = [&](long long bb) {  // ADJ: this CTA's partial moments of batch element bb -> global
    for (int x = tid; x < n_exec; x += nt) {
      if (eg[x].gr == GR_NONE) continue;
      double* dst = mom_g + ((long long)(sw.first_exec + x) * batch + bb) * 4;
      const int nq = eg[x].gr == GR_SCALE ? 1 : 4;
      for (int q = 0; q < 4; ++q) {
        double acc = 0.0;
        for (int w = 0; w < nwarps; ++w) { acc += gsum[(x * nwarps + w) * 4 + q]; gsum[(x * nwarps + w) * 4 + q] = 0.0; }
        if (q < nq) atomicAdd(dst + q, acc);
      }
    }
  };

  long long cur_b = -1;
  unsigned long long next_base = 0;
  if (tile_first < tile_end) { next_base = tile_base(tile_first); issue_loads(tile_first, next_base, 0); }
  for (long long t = tile_first; t < tile_end; ++t) {
  const int which = (int)((t - tile_first) & (nbuf - 1));
  const long long b = t / tiles_per_state;
  const unsigned long long cta_base = next_base;
  const unsigned long long cta_index = cta_base | index_hi;
  c2* tile_psi = bufs + which * NSTATE * tile;
  c2* tile_lam = tile_psi + (ADJ ? tile : 0);
  if (t + 1 < tile_end && nbuf == 2) { next_base = tile_base(t + 1); issue_loads(t + 1, next_base, which ^ 1); cp_async_wait<1>(); }
  else if (t > tile_first && nbuf == 1) { issue_loads(t, cta_base, 0); cp_async_wait<0>(); }
  else cp_async_wait<0>();
  if (nbuf == 1 && t + 1 < tile_end) next_base = tile_base(t + 1);
  if (b != cur_b) {  // new batch element: its resolved gates (one coalesced copy from the plan-wide table)
    if (ADJ && cur_b >= 0) { flush_moments(cur_b); __syncthreads(); }  // eg[].gr is read by the flush
    constexpr int W = sizeof(ExecGate<real>) / 16;
    const uint4* src = reinterpret_cast<const uint4*>(xtab);
    uint4* dst = reinterpret_cast<uint4*>(eg);
    for (int w = tid; w < n_exec * W; w += nt) {
      const int x = w / W, k = w % W;
      dst[w] = src[((long long)(sw.first_exec + x) * batch + b) * W + k];
    }
    cur_b = b;
  }
  for (int rr = tid; rr < n_rounds; rr += nt) {  // per-tile constant of the permutation phase: external (outer / rank bit) controls
    const b200sv_round& rd = rounds[sw.first_round + rr];
    unsigned ext = 0;
    for (int k = 0; k < (int)rd.n_pext; ++k)
      if ((cta_index >> rd.pext[k].bit) & 1ull) ext ^= rd.pext[k].vec;
    rtab[rr * RT_STRIDE + 36] = (unsigned)swz<real>((int)ext) << 16;
  }
  __syncthreads();  // tile `which` has landed; eg / rtab[36] visible
  const long long state_off = b << n;

  // ---- rounds ----
  for (int rr = 0; rr < ((dbg & 2) ? 0 : sw.n_rounds); ++rr) {
    const int rloc = dir > 0 ? rr : sw.n_rounds - 1 - rr;
    const int ridx = sw.first_round + rloc;
    const uint4* rq = reinterpret_cast<const uint4*>(rounds + ridx);
    const uint4 rw1 = rq[1];  // thrpos[4:12] first_exec n_exec
    const uint4 rw2 = rq[2];  // mslot[0:4]
    const int first_exec = (int)rw1.z, n_ex = (int)rw1.w;
    const int ms[4] = {(int)rw2.x, (int)rw2.y, (int)rw2.z, (int)rw2.w};
    const bool has_perm = (rounds[ridx].flags & B200SV_RF_PERM) != 0;
    const unsigned* T = rtab + rloc * RT_STRIDE;
    const unsigned u = T[tid & 15] ^ T[16 + (tid >> 4)] ^ T[36];
    const int st = (int)(u & 0xffffu), sp = (int)(u >> 16);  // swizzled plain / permuted base index
    int ro[R], rp[R];
#pragma unroll
    for (int a = 0; a < R; ++a) { const unsigned w = T[32 + a]; ro[a] = (int)(w & 0xffffu); rp[a] = (int)(w >> 16); }
    auto sidx = [&](int j) {  // swizzled tile index of register j before the permutation phase
      int o = st;
#pragma unroll
      for (int a = 0; a < R; ++a)
        if ((j >> a) & 1) o ^= ro[a];
      return o;
    };
    auto pidx = [&](int j) {  // ... after it
      int o = sp;
#pragma unroll
      for (int a = 0; a < R; ++a)
        if ((j >> a) & 1) o ^= rp[a];
      return o;
    };
    unsigned long long gthr = 0;
    if constexpr (GEN) {
      if (n_ex > 0) {  // index bits of this thread (controls / external diagonals of the generic phase)
        const int tbase = swz<real>(st);  // swz is an involution
        gthr = cta_base | t_lo[tbase & 63] | t_hi[tbase >> 6] | index_hi;
      }
    }

    c2 p[NV];
    c2 l[ADJ ? NV : 1];
    if constexpr (dir > 0) {
#pragma unroll
      for (int j = 0; j < NV; ++j) p[j] = tile_psi[sidx(j)];
    } else {
#pragma unroll
      for (int j = 0; j < NV; ++j) p[j] = tile_psi[pidx(j)];
      if constexpr (ADJ) {
#pragma unroll
        for (int j = 0; j < NV; ++j) l[j] = tile_lam[pidx(j)];
      }
    }
    if (has_perm) __syncthreads();  // every register file is loaded before anyone stores at permuted addresses

    if constexpr (dir > 0) {
      // ---- M phase: one merged 2x2 per occupied register bit; straight-line code per slot mask ----
      {
        const int mask = (dbg & 1) ? 0 : (ms[0] >= 0 ? 1 : 0) | (ms[1] >= 0 ? 2 : 0) | (ms[2] >= 0 ? 4 : 0) | (ms[3] >= 0 ? 8 : 0);
        // full rounds take the straight-line path; partial rounds a compact per-slot path (specialising
        // every mask bloats the kernel past the instruction cache)
        constexpr int FULL = (1 << R) - 1;
        if (mask == FULL) RG::template mphase_fwd<FULL>(p, eg, ms);
        else {
#pragma unroll
          for (int a = 0; a < R; ++a) {
            if (mask & (1 << a)) {
              const real* mm = eg[ms[a]].m;
              if (a == 0) RG::template mat_<0, false>(p, mm, 0);
              if constexpr (R > 1) { if (a == 1) RG::template mat_<1, false>(p, mm, 0); }
              if constexpr (R > 2) { if (a == 2) RG::template mat_<2, false>(p, mm, 0); }
              if constexpr (R > 3) { if (a == 3) RG::template mat_<3, false>(p, mm, 0); }
            }
          }
        }
      }
      // ---- G phase ----
      if constexpr (GEN) {
        for (int gi = 0; gi < n_ex; ++gi) {
          const ExecGate<real>& e = eg[first_exec + gi];
          if ((gthr & e.ctrl_ext) == e.ctrl_ext) RG::apply(p, e, gthr, false);
        }
      }
#pragma unroll
      for (int j = 0; j < NV; ++j) tile_psi[pidx(j)] = p[j];
    } else {
      // every resolved entry lives in exactly one round, so each warp owns one gsum slot per entry:
      // plain stores, no atomics (nt < 32: a single partial warp, lanes reduce through full-mask shuffles
      // of the active lanes only)
      auto reduce_moments = [&](double (&mo)[4], int nmo, int x) {
        if (nt >= 32 && nmo == 4) {
          // transposed butterfly: 6 double shuffles for 4 values instead of 20; totals land in lanes 0, 8, 16, 24
          const int lane = tid & 31;
          const bool h16 = (lane & 16) != 0, h8 = (lane & 8) != 0;
          const double r0 = __shfl_xor_sync(0xffffffffu, h16 ? mo[0] : mo[2], 16);
          const double r1 = __shfl_xor_sync(0xffffffffu, h16 ? mo[1] : mo[3], 16);
          const double a0 = (h16 ? mo[2] : mo[0]) + r0, a1 = (h16 ? mo[3] : mo[1]) + r1;
          double v = (h8 ? a1 : a0) + __shfl_xor_sync(0xffffffffu, h8 ? a0 : a1, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          if ((lane & 7) == 0) gsum[(x * nwarps + (tid >> 5)) * 4 + (lane >> 3)] += v;
          return;
        }
        const unsigned mask = nt >= 32 ? 0xffffffffu : ((1u << nt) - 1u);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (q < nmo) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              if (o < nt) mo[q] += __shfl_xor_sync(mask, mo[q], o);
            }
          }
        }
        if ((tid & 31) == 0)
          for (int q = 0; q < nmo; ++q) gsum[(x * nwarps + (tid >> 5)) * 4 + q] += mo[q];
      };
      // ---- G phase, reversed ----
      if constexpr (GEN) {
      for (int gi = n_ex - 1; gi >= 0; --gi) {
        const int x = first_exec + gi;
        const ExecGate<real>& e = eg[x];
        const bool active = (gthr & e.ctrl_ext) == e.ctrl_ext;
        if constexpr (ADJ) {
          if (e.gr != GR_NONE) {  // uniform across the CTA
            double mo[4] = {0.0, 0.0, 0.0, 0.0};
            int nmo = 4;
            if (e.gr == GR_SCALE) {
              RG::apply(p, e, gthr, false);  // ψ_before = ψ / s
              mo[0] = RG::grad_scale(p, l);
              RG::apply(l, e, gthr, true);
              nmo = 1;
            } else if (active) {
              RG::moments(p, l, e, gthr, mo);
            }
            reduce_moments(mo, nmo, x);
            if (e.gr == GR_SCALE) continue;
          }
          if (active) { RG::apply(p, e, gthr, false); RG::apply(l, e, gthr, true); }
        } else {
          if (active) RG::apply(p, e, gthr, false);
        }
      }
      }
      // ---- M phase, reversed: moments of the post-chain states, then un-apply the merged 2x2 ----
      {
        const int mask = (dbg & 1) ? 0 : (ms[0] >= 0 ? 1 : 0) | (ms[1] >= 0 ? 2 : 0) | (ms[2] >= 0 ? 4 : 0) | (ms[3] >= 0 ? 8 : 0);
        constexpr int FULL = (1 << R) - 1;
        if (mask == FULL) RG::template mphase_bwd<FULL, ADJ>(p, l, eg, ms, reduce_moments);
        else {
          if constexpr (R > 3) { if (mask & 8) RG::template mphase_bwd<8, ADJ>(p, l, eg, ms, reduce_moments); }
          if constexpr (R > 2) { if (mask & 4) RG::template mphase_bwd<4, ADJ>(p, l, eg, ms, reduce_moments); }
          if constexpr (R > 1) { if (mask & 2) RG::template mphase_bwd<2, ADJ>(p, l, eg, ms, reduce_moments); }
          if (mask & 1) RG::template mphase_bwd<1, ADJ>(p, l, eg, ms, reduce_moments);
        }
      }
#pragma unroll
      for (int j = 0; j < NV; ++j) tile_psi[sidx(j)] = p[j];
      if constexpr (ADJ) {
#pragma unroll
        for (int j = 0; j < NV; ++j) tile_lam[sidx(j)] = l[j];
      }
    }
    __syncthreads();
  }

  // ---- write the tile back ----
  {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int e = i * nt + tid;
      *gaddr(psi_g, (unsigned long long)(state_off + (long long)(cta_base + t_lo[e & 63] + t_hi[e >> 6]))) = tile_psi[swz<real>(e)];
    }
    if (ADJ) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int e = i * nt + tid;
        *laddr(lam_g, (unsigned long long)(state_off + (long long)(cta_base + t_lo[e & 63] + t_hi[e >> 6]))) = tile_lam[swz<real>(e)];
      }
    }
  }
  __syncthreads();  // buffer `which` is free again before the next iteration's prefetch overwrites it
  }  // tile loop
  if (ADJ && cur_b >= 0) flush_moments(cur_b);
}

// ------------------------------------------------------------------------------------------------
// resolve kernel: one thread per (resolved entry, batch element) evaluates sincos of the parametric
// angles and multiplies the chain into one 2x2 — once per launch instead of once per CTA
// ------------------------------------------------------------------------------------------------
template <typename real>
__global__ void resolve_kernel(const b200sv_gate* __restrict__ gates, const int* __restrict__ exec_leader,
                               const double* __restrict__ params, const double* __restrict__ consts,
                               ExecGate<real>* __restrict__ xtab, int x_first, int x_count, long long batch,
                               int dir, int adj) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)x_count * batch) return;
  const int x = x_first + (int)(t / batch);
  const long long b = t % batch;
  ExecGate<real> e;
  resolve_gate<real>(gates + exec_leader[x], e, params, consts, batch, b, dir, adj != 0);
  xtab[(long long)x * batch + b] = e;
}

// gradient kernel: moments of every chain -> gradient of every parametric gate of the chain
template <typename real>
__global__ void gradient_kernel(const b200sv_gate* __restrict__ gates, const int* __restrict__ exec_leader,
                                const double* __restrict__ params, const double* __restrict__ consts,
                                const ExecGate<real>* __restrict__ xtab, const double* __restrict__ mom_g,
                                double* __restrict__ grads, int x_first, int x_count, long long batch) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)x_count * batch) return;
  const int x = x_first + (int)(t / batch);
  const long long b = t % batch;
  const int gr = xtab[(long long)x * batch + b].gr;
  if (gr == GR_NONE) return;
  const b200sv_gate* rec = gates + exec_leader[x];
  const double* mom = mom_g + ((long long)x * batch + b) * 4;
  if (gr == GR_CHAIN) chain_gradients(rec, mom, params, consts, grads, batch, b);
  else if (rec->grad_row >= 0) atomicAdd(&grads[(long long)rec->grad_row * batch + b], mom[0]);
}

template <typename real, bool ADJ>
inline size_t sweep_smem_bytes(int m, int n_exec, int n_rounds, int nbuf) {
  size_t tile = (size_t)1 << m;
  size_t s = (size_t)nbuf * tile * 2 * sizeof(real) * (ADJ ? 2 : 1);  // nbuf = 2: double-buffered persistent CTAs
  s += 128 * sizeof(unsigned long long);
  s += (size_t)n_exec * sizeof(ExecGate<real>);
  s += (size_t)n_exec * 4 * sizeof(double) * 8;  // up to 8 warps per CTA
  s += (size_t)n_rounds * RT_STRIDE * sizeof(unsigned);
  return s;
}

}  // namespace b200sv
