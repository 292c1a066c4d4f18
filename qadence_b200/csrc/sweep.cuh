// Fused gate-sweep kernels (K1/K2/K6 of SURVEY.md §2.1) for sm_100a.
//
// One launch = one "sweep": every CTA stages a tile of 2^m amplitudes of one batch element in
// shared memory (coalesced, vectorised 16-byte global accesses), then runs several "rounds".  In a
// round each thread pulls 2^r amplitudes into registers (the r "register bits" of that round),
// applies every gate of the round entirely in registers (sincos for parametric rotations is
// evaluated in-kernel, once per CTA), and writes them back to the XOR-swizzled tile.  After the
// last round the tile is written back: 2·S bytes of HBM traffic for all gates of the sweep
// instead of 2·S per gate (the reference's einsum-per-gate loop,
// /root/reference/qadence/backends/pyqtorch/backend.py:176).
//
// The adjoint variant (ADJ) walks the same records backwards on ψ and λ simultaneously and reduces
// the per-parameter gradient inner products 2·Re<λ|∂U|ψ> in the same pass (4·S bytes).
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

// execution kinds after per-CTA resolution of parameters
enum { EX_MAT = 0, EX_REAL, EX_RX, EX_DIAG, EX_X, EX_SWAP, EX_SCALE, EX_NOP };
enum { GR_NONE = 0, GR_RX, GR_RY, GR_RZ, GR_PHASE, GR_SCALE };

template <typename real> struct ExecGate {
  real m[8];
  unsigned long long ctrl_ext;
  int ex, a, a2, ext_bit, ctrl_reg, gr;
};

template <typename real> __device__ __forceinline__ int swz(int e);
template <> __device__ __forceinline__ int swz<double>(int e) {
  return e ^ (((e >> 3) ^ (e >> 6) ^ (e >> 9)) & 7);
}
template <> __device__ __forceinline__ int swz<float>(int e) {
  return e ^ (((e >> 4) ^ (e >> 8)) & 15);
}

// ------------------------------------------------------------------------------------------------
// per-CTA resolution of one gate record into an ExecGate (matrix entries for THIS batch element)
// ------------------------------------------------------------------------------------------------
template <typename real>
__device__ void resolve_gate(const b200sv_gate& g, ExecGate<real>& e, const double* __restrict__ params,
                             const double* __restrict__ consts, long long batch, long long b, int dir,
                             bool want_grad) {
  e.ctrl_ext = g.ctrl_ext;
  e.a = g.a; e.a2 = g.a2; e.ext_bit = g.ext_bit; e.ctrl_reg = g.ctrl_reg;
  e.gr = GR_NONE;
  double theta = 0.0, theta_im = 0.0;
  if (g.kind >= B200SV_GK_RX && g.kind <= B200SV_GK_PHASE || g.kind == B200SV_GK_SCALE) {
    if (g.slot >= 0) theta = params[(long long)g.slot * batch + b];
    else { theta = consts[2 * g.cidx]; theta_im = consts[2 * g.cidx + 1]; }
  }
  const bool grad = want_grad && (g.flags & B200SV_GF_GRAD) && g.grad_row >= 0;
  switch (g.kind) {
    case B200SV_GK_MAT: {
      const double* c = consts + 2 * g.cidx;
      if (dir > 0) { for (int i = 0; i < 8; ++i) e.m[i] = (real)c[i]; }
      else {  // conjugate transpose
        e.m[0] = (real)c[0]; e.m[1] = (real)-c[1]; e.m[2] = (real)c[4]; e.m[3] = (real)-c[5];
        e.m[4] = (real)c[2]; e.m[5] = (real)-c[3]; e.m[6] = (real)c[6]; e.m[7] = (real)-c[7];
      }
      e.ex = EX_MAT; break;
    }
    case B200SV_GK_REAL: {
      const double* c = consts + 2 * g.cidx;  // m00 m01 m10 m11
      e.m[0] = (real)c[0]; e.m[3] = (real)c[3];
      e.m[1] = (real)(dir > 0 ? c[1] : c[2]); e.m[2] = (real)(dir > 0 ? c[2] : c[1]);
      e.ex = EX_REAL; break;
    }
    case B200SV_GK_RX: {
      double s, c; sincos(0.5 * theta, &s, &c); if (dir < 0) s = -s;
      e.m[0] = (real)c; e.m[1] = (real)s; e.ex = EX_RX; if (grad) e.gr = GR_RX; break;
    }
    case B200SV_GK_RY: {
      double s, c; sincos(0.5 * theta, &s, &c); if (dir < 0) s = -s;
      e.m[0] = (real)c; e.m[1] = (real)-s; e.m[2] = (real)s; e.m[3] = (real)c;
      e.ex = EX_REAL; if (grad) e.gr = GR_RY; break;
    }
    case B200SV_GK_RZ: {
      double s, c; sincos(0.5 * theta, &s, &c); if (dir < 0) s = -s;
      e.m[0] = (real)c; e.m[1] = (real)-s; e.m[2] = (real)c; e.m[3] = (real)s;
      e.ex = EX_DIAG; if (grad) e.gr = GR_RZ; break;
    }
    case B200SV_GK_PHASE: {
      double s, c; sincos(theta, &s, &c); if (dir < 0) s = -s;
      e.m[0] = (real)1; e.m[1] = (real)0; e.m[2] = (real)c; e.m[3] = (real)s;
      e.ex = EX_DIAG; if (grad) e.gr = GR_PHASE; break;
    }
    case B200SV_GK_DIAG: {
      const double* c = consts + 2 * g.cidx;
      const double sg = dir > 0 ? 1.0 : -1.0;
      e.m[0] = (real)c[0]; e.m[1] = (real)(sg * c[1]); e.m[2] = (real)c[2]; e.m[3] = (real)(sg * c[3]);
      e.ex = EX_DIAG; break;
    }
    case B200SV_GK_X: e.ex = EX_X; break;
    case B200SV_GK_SWAP: e.ex = EX_SWAP; break;
    case B200SV_GK_SCALE: {
      if (dir > 0) { e.m[0] = (real)theta; e.m[1] = (real)theta_im; }
      else {  // psi <- psi / s (inverse), lambda <- conj(s) lambda (adjoint)
        const double d = theta * theta + theta_im * theta_im;
        e.m[0] = (real)(theta / d); e.m[1] = (real)(-theta_im / d);
        e.m[2] = (real)theta; e.m[3] = (real)-theta_im;
      }
      e.ex = EX_SCALE; if (grad) e.gr = GR_SCALE; break;
    }
    default: e.ex = EX_NOP; break;
  }
}

// ------------------------------------------------------------------------------------------------
// in-register gate application: v[j], j in [0, 2^R); gate target = register bit A
// ------------------------------------------------------------------------------------------------
template <typename real, int R> struct Regs {
  using c2 = typename C2<real>::type;
  static constexpr int NV = 1 << R;

  template <int A> static __device__ __forceinline__ void mat(c2 (&v)[NV], const real* m, int cr) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (j & (1 << A)) continue;
      if ((j & cr) != cr) continue;
      const c2 x = v[j], y = v[j | (1 << A)];
      v[j] = mk<c2>(m[0] * x.x - m[1] * x.y + m[2] * y.x - m[3] * y.y,
                    m[0] * x.y + m[1] * x.x + m[2] * y.y + m[3] * y.x);
      v[j | (1 << A)] = mk<c2>(m[4] * x.x - m[5] * x.y + m[6] * y.x - m[7] * y.y,
                               m[4] * x.y + m[5] * x.x + m[6] * y.y + m[7] * y.x);
    }
  }
  template <int A> static __device__ __forceinline__ void realm(c2 (&v)[NV], const real* m, int cr) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (j & (1 << A)) continue;
      if ((j & cr) != cr) continue;
      const c2 x = v[j], y = v[j | (1 << A)];
      v[j] = mk<c2>(m[0] * x.x + m[1] * y.x, m[0] * x.y + m[1] * y.y);
      v[j | (1 << A)] = mk<c2>(m[2] * x.x + m[3] * y.x, m[2] * x.y + m[3] * y.y);
    }
  }
  template <int A> static __device__ __forceinline__ void rx(c2 (&v)[NV], real c, real s, int cr) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (j & (1 << A)) continue;
      if ((j & cr) != cr) continue;
      const c2 x = v[j], y = v[j | (1 << A)];
      // [[c, -i s], [-i s, c]]
      v[j] = mk<c2>(c * x.x + s * y.y, c * x.y - s * y.x);
      v[j | (1 << A)] = mk<c2>(c * y.x + s * x.y, c * y.y - s * x.x);
    }
  }
  template <int A> static __device__ __forceinline__ void xperm(c2 (&v)[NV], int cr) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (j & (1 << A)) continue;
      if ((j & cr) != cr) continue;
      const c2 x = v[j]; v[j] = v[j | (1 << A)]; v[j | (1 << A)] = x;
    }
  }
  template <int A, int B> static __device__ __forceinline__ void swapab(c2 (&v)[NV], int cr) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if ((j & (1 << A)) || !(j & (1 << B))) continue;  // j: bit A = 0, bit B = 1
      if ((j & cr) != cr) continue;
      const int k = (j | (1 << A)) & ~(1 << B);
      const c2 x = v[j]; v[j] = v[k]; v[k] = x;
    }
  }
  static __device__ __forceinline__ void diag_in(c2 (&v)[NV], const real* m, int a, int cr) {
    const c2 d0 = mk<c2>(m[0], m[1]), d1 = mk<c2>(m[2], m[3]);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if ((j & cr) != cr) continue;
      v[j] = cmul(v[j], ((j >> a) & 1) ? d1 : d0);
    }
  }
  static __device__ __forceinline__ void scale_all(c2 (&v)[NV], c2 d, int cr) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if ((j & cr) != cr) continue;
      v[j] = cmul(v[j], d);
    }
  }

#define B200SV_DISPATCH_A(a_, STMT)                                   \
  switch (a_) {                                                       \
    case 0: { constexpr int A = 0; STMT; } break;                     \
    case 1: if constexpr (R > 1) { constexpr int A = 1; STMT; } break; \
    case 2: if constexpr (R > 2) { constexpr int A = 2; STMT; } break; \
    case 3: if constexpr (R > 3) { constexpr int A = 3; STMT; } break; \
    case 4: if constexpr (R > 4) { constexpr int A = 4; STMT; } break; \
    default: break;                                                   \
  }

  static __device__ __forceinline__ void swap_dispatch(c2 (&v)[NV], int a, int b, int cr) {
    if (a > b) { const int t = a; a = b; b = t; }
    // a < b
    switch (a * 8 + b) {
      case 0 * 8 + 1: if constexpr (R > 1) swapab<0, 1>(v, cr); break;
      case 0 * 8 + 2: if constexpr (R > 2) swapab<0, 2>(v, cr); break;
      case 1 * 8 + 2: if constexpr (R > 2) swapab<1, 2>(v, cr); break;
      case 0 * 8 + 3: if constexpr (R > 3) swapab<0, 3>(v, cr); break;
      case 1 * 8 + 3: if constexpr (R > 3) swapab<1, 3>(v, cr); break;
      case 2 * 8 + 3: if constexpr (R > 3) swapab<2, 3>(v, cr); break;
      case 0 * 8 + 4: if constexpr (R > 4) swapab<0, 4>(v, cr); break;
      case 1 * 8 + 4: if constexpr (R > 4) swapab<1, 4>(v, cr); break;
      case 2 * 8 + 4: if constexpr (R > 4) swapab<2, 4>(v, cr); break;
      case 3 * 8 + 4: if constexpr (R > 4) swapab<3, 4>(v, cr); break;
      default: break;
    }
  }

  // apply one resolved gate to a register file (thread already known to satisfy ctrl_ext)
  static __device__ __forceinline__ void apply(c2 (&v)[NV], const ExecGate<real>& e,
                                               unsigned long long gthr, bool is_lambda) {
    const int cr = e.ctrl_reg;
    switch (e.ex) {
      case EX_MAT: B200SV_DISPATCH_A(e.a, mat<A>(v, e.m, cr)); break;
      case EX_REAL: B200SV_DISPATCH_A(e.a, realm<A>(v, e.m, cr)); break;
      case EX_RX: B200SV_DISPATCH_A(e.a, rx<A>(v, e.m[0], e.m[1], cr)); break;
      case EX_X: B200SV_DISPATCH_A(e.a, xperm<A>(v, cr)); break;
      case EX_SWAP: swap_dispatch(v, e.a, e.a2, cr); break;
      case EX_DIAG:
        if (e.a >= 0) diag_in(v, e.m, e.a, cr);
        else {
          const int bit = (int)((gthr >> e.ext_bit) & 1ull);
          const c2 d = bit ? mk<c2>(e.m[2], e.m[3]) : mk<c2>(e.m[0], e.m[1]);
          if (!(d.x == (real)1 && d.y == (real)0)) scale_all(v, d, cr);
        }
        break;
      case EX_SCALE:
        scale_all(v, is_lambda ? mk<c2>(e.m[2], e.m[3]) : mk<c2>(e.m[0], e.m[1]), 0);
        break;
      default: break;
    }
  }

  // ---- adjoint gradient partials, evaluated on the POST-gate ψ, λ ----
  //  RX/RY/RZ: g = Im<λ|P|ψ> on the control-active subspace; PHASE: g = -2 Im(conj(λ1) ψ1).
  template <int A>
  static __device__ __forceinline__ double grad_pair(const c2 (&p)[NV], const c2 (&l)[NV], int gr, int cr) {
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (j & (1 << A)) continue;
      if ((j & cr) != cr) continue;
      const int k = j | (1 << A);
      const double p0x = p[j].x, p0y = p[j].y, p1x = p[k].x, p1y = p[k].y;
      const double l0x = l[j].x, l0y = l[j].y, l1x = l[k].x, l1y = l[k].y;
      if (gr == GR_RX) acc += (l0x * p1y - l0y * p1x) + (l1x * p0y - l1y * p0x);
      else if (gr == GR_RY) acc += -(l0x * p1x + l0y * p1y) + (l1x * p0x + l1y * p0y);
      else if (gr == GR_RZ) acc += (l0x * p0y - l0y * p0x) - (l1x * p1y - l1y * p1x);
      else if (gr == GR_PHASE) acc += -2.0 * (l1x * p1y - l1y * p1x);
    }
    return acc;
  }
  static __device__ __forceinline__ double grad_ext(const c2 (&p)[NV], const c2 (&l)[NV], int gr, int cr, int bit) {
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if ((j & cr) != cr) continue;
      acc += (double)l[j].x * (double)p[j].y - (double)l[j].y * (double)p[j].x;  // Im(conj(l) p)
    }
    if (gr == GR_RZ) return bit ? -acc : acc;
    if (gr == GR_PHASE) return bit ? -2.0 * acc : 0.0;
    return 0.0;
  }
  static __device__ __forceinline__ double grad_scale(const c2 (&p)[NV], const c2 (&l)[NV]) {
    double acc = 0.0;  // 2 Re <λ|ψ_before>
#pragma unroll
    for (int j = 0; j < NV; ++j) acc += (double)l[j].x * (double)p[j].x + (double)l[j].y * (double)p[j].y;
    return 2.0 * acc;
  }
  static __device__ __forceinline__ double gradient(const c2 (&p)[NV], const c2 (&l)[NV],
                                                    const ExecGate<real>& e, unsigned long long gthr) {
    double g = 0.0;
    if (e.a >= 0) { B200SV_DISPATCH_A(e.a, g = grad_pair<A>(p, l, e.gr, e.ctrl_reg)); }
    else g = grad_ext(p, l, e.gr, e.ctrl_reg, (int)((gthr >> e.ext_bit) & 1ull));
    return g;
  }
};

// ------------------------------------------------------------------------------------------------
// the sweep kernel
// ------------------------------------------------------------------------------------------------
template <typename real, int R, bool ADJ>
__global__ void __launch_bounds__(256)
sweep_kernel(typename C2<real>::type* __restrict__ psi_g, typename C2<real>::type* __restrict__ lam_g,
             const double* __restrict__ params, double* __restrict__ grads,
             const b200sv_sweep* __restrict__ sweeps, const b200sv_round* __restrict__ rounds,
             const b200sv_gate* __restrict__ gates, const double* __restrict__ consts, int sweep_idx,
             int n, long long batch, unsigned long long index_hi, int dir) {
  using c2 = typename C2<real>::type;
  using RG = Regs<real, R>;
  constexpr int NV = 1 << R;
  extern __shared__ __align__(16) unsigned char smem_raw[];

  const b200sv_sweep& sw = sweeps[sweep_idx];
  const int m = sw.m;
  const int tile = 1 << m;
  const int nt = blockDim.x;  // == 2^(m-R)
  const int tid = threadIdx.x;
  const int n_gates = sw.n_gates;

  c2* tile_psi = reinterpret_cast<c2*>(smem_raw);
  c2* tile_lam = tile_psi + (ADJ ? tile : 0);
  unsigned long long* t_lo = reinterpret_cast<unsigned long long*>(tile_lam + tile);
  unsigned long long* t_hi = t_lo + 64;
  ExecGate<real>* eg = reinterpret_cast<ExecGate<real>*>(t_hi + 64);
  double* gsum = reinterpret_cast<double*>(eg + n_gates);

  const long long tiles_per_state = 1ll << sw.n_outer;
  const long long b = blockIdx.x / tiles_per_state;
  const unsigned long long c = blockIdx.x % tiles_per_state;
  unsigned long long cta_base = 0;
  for (int k = 0; k < sw.n_outer; ++k) cta_base |= ((c >> k) & 1ull) << sw.outer_bits[k];

  // ---- prologue: offset tables, resolved gates ----
  for (int v = tid; v < 64; v += nt) {
    unsigned long long lo = 0, hi = 0;
    for (int i = 0; i < 6; ++i) {
      if (i < m && ((v >> i) & 1)) lo |= 1ull << sw.tile_bits[i];
      if (i + 6 < m && ((v >> i) & 1)) hi |= 1ull << sw.tile_bits[i + 6];
    }
    t_lo[v] = lo; t_hi[v] = hi;
  }
  for (int g = tid; g < n_gates; g += nt) {
    resolve_gate<real>(gates[sw.first_gate + g], eg[g], params, consts, batch, b, dir, ADJ);
    if (ADJ) gsum[g] = 0.0;
  }
  __syncthreads();

  // ---- stage the tile: coalesced global -> swizzled shared ----
  const long long state_off = b << n;
  {
    c2 buf[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int e = i * nt + tid;
      buf[i] = psi_g[state_off + (long long)(cta_base + t_lo[e & 63] + t_hi[e >> 6])];
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) tile_psi[swz<real>(i * nt + tid)] = buf[i];
    if (ADJ) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int e = i * nt + tid;
        buf[i] = lam_g[state_off + (long long)(cta_base + t_lo[e & 63] + t_hi[e >> 6])];
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) tile_lam[swz<real>(i * nt + tid)] = buf[i];
    }
  }
  __syncthreads();

  // ---- rounds ----
  const int nthr_bits = m - R;
  for (int rr = 0; rr < sw.n_rounds; ++rr) {
    const int ridx = dir > 0 ? sw.first_round + rr : sw.first_round + sw.n_rounds - 1 - rr;
    const b200sv_round rd = rounds[ridx];
    int tbase = 0;
    for (int i = 0; i < nthr_bits; ++i) tbase |= ((tid >> i) & 1) << rd.thrpos[i];
    int idx[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      int o = tbase;
#pragma unroll
      for (int a = 0; a < R; ++a)
        if ((j >> a) & 1) o |= 1 << rd.regpos[a];
      idx[j] = swz<real>(o);
    }
    const unsigned long long gthr = cta_base | t_lo[tbase & 63] | t_hi[tbase >> 6] | index_hi;

    c2 p[NV];
    c2 l[ADJ ? NV : 1];
#pragma unroll
    for (int j = 0; j < NV; ++j) p[j] = tile_psi[idx[j]];
    if constexpr (ADJ) {
#pragma unroll
      for (int j = 0; j < NV; ++j) l[j] = tile_lam[idx[j]];
    }

    for (int gi = 0; gi < rd.n_gates; ++gi) {
      const int g = (dir > 0 ? rd.first_gate + gi : rd.first_gate + rd.n_gates - 1 - gi) - sw.first_gate;
      const ExecGate<real>& e = eg[g];
      const bool active = (gthr & e.ctrl_ext) == e.ctrl_ext;
      if constexpr (ADJ) {
        if (e.gr != GR_NONE) {  // uniform across the CTA
          double part = 0.0;
          if (e.gr == GR_SCALE) {
            if (active) RG::apply(p, e, gthr, false);  // ψ_before = ψ / s
            part = RG::grad_scale(p, l);
            RG::apply(l, e, gthr, true);
          } else if (active) {
            part = RG::gradient(p, l, e, gthr);
          }
          if (nt >= 32) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if ((tid & 31) == 0) atomicAdd(&gsum[g], part);
          } else {
            atomicAdd(&gsum[g], part);
          }
          if (e.gr == GR_SCALE) continue;
        }
        if (active) { RG::apply(p, e, gthr, false); RG::apply(l, e, gthr, true); }
      } else {
        if (active) RG::apply(p, e, gthr, false);
      }
    }

#pragma unroll
    for (int j = 0; j < NV; ++j) tile_psi[idx[j]] = p[j];
    if constexpr (ADJ) {
#pragma unroll
      for (int j = 0; j < NV; ++j) tile_lam[idx[j]] = l[j];
    }
    __syncthreads();
  }

  // ---- write the tile back ----
  {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int e = i * nt + tid;
      psi_g[state_off + (long long)(cta_base + t_lo[e & 63] + t_hi[e >> 6])] = tile_psi[swz<real>(e)];
    }
    if (ADJ) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int e = i * nt + tid;
        lam_g[state_off + (long long)(cta_base + t_lo[e & 63] + t_hi[e >> 6])] = tile_lam[swz<real>(e)];
      }
      for (int g = tid; g < n_gates; g += nt) {
        const b200sv_gate& gr = gates[sw.first_gate + g];
        if ((gr.flags & B200SV_GF_GRAD) && gr.grad_row >= 0)
          atomicAdd(&grads[(long long)gr.grad_row * batch + b], gsum[g]);
      }
    }
  }
}

template <typename real, bool ADJ>
inline size_t sweep_smem_bytes(int m, int n_gates) {
  size_t tile = (size_t)1 << m;
  size_t s = tile * 2 * sizeof(real) * (ADJ ? 2 : 1);
  s += 128 * sizeof(unsigned long long);
  s += (size_t)n_gates * sizeof(ExecGate<real>);
  s += (size_t)n_gates * sizeof(double);
  return s;
}

}  // namespace b200sv
