// Lean fused-sweep kernel (round 2) for sm_100a: sweeps that consist only of merged 1-qubit chains ("M slots") and
// index permutations (X / CNOT / SWAP folded into addressing) — every sweep of an HEA, a feature map, a QNN.
//
// What differs from `sweep_kernel` (sweep.cuh), which stays the executor of sweeps with generic-phase entries:
//  * tiles move with TMA: one `cp.async.bulk.tensor` per state per tile (SASS UTMALDG / UTMASTG) through a tensor map
//    that describes the tile as a <= 5-D box of the [batch][2^n] state (low run x window x ...), completion on an
//    mbarrier; no per-element address arithmetic, no LDGSTS issue slots.  Tiles that are not such a box fall back to
//    cp.async gathers inside the same kernel.
//  * persistent CTAs that claim chunks of tiles from a device counter (dynamic scheduling: the warp schedulers favour the
//    older CTAs of an SM, a static split leaves the young ones alone at the end); index tables are built ONCE on the host
//    (plan.py: `lean tables`) and copied once per CTA.
//  * normalised 2x2:  U = nu * [[1, b], [g, d]] after moving the largest entry to position (0,0) by flipping the pair on
//    the load side (column flip) and / or the store side (row flip) — both are address XORs.  Applying the normalised
//    matrix costs 12 FP64 instructions per amplitude pair instead of 16; the scalars nu of a sweep are multiplied
//    together per batch element by the resolve kernel and applied once, in the last round of the sweep.  In the adjoint
//    walk psi and lambda carry the SAME pending scalar s, so a measured moment Im<lambda'|P|psi'> is the true one divided
//    by |s|^2 (real): the gradient kernel rescales.  A column flip conjugates P by X: the Y and Z moments change sign.
//  * tile buffers are aligned to their own size in shared memory, so every LDS / STS address is `base ^ mask`: one LOP3
//    per access, lambda at an immediate offset from psi.
//  * one copy of the round body (slots are skipped by a uniform branch), to stay inside the instruction cache.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "b200sv.h"
#include "sweep.cuh"

namespace b200sv {

constexpr int LT_STRIDE = B200SV_LT_STRIDE;
constexpr unsigned LT_EMPTY = 0xffffu;

template <typename real> struct __align__(16) LeanGate {
  real m[6];          // normalised 2x2 [[1, b], [g, d]]: b.re b.im g.re g.im d.re d.im
  real mscale;        // adjoint: measured moments of this entry are multiplied by mscale = |pending scalar|^2
  unsigned flipmask;  // XOR-ed (AND register-bit table word) into the packed load|store address: 0xffff column flip, 0xffff0000 row flip
  int gr;             // GR_* | column flip << 8
};
static_assert(sizeof(LeanGate<double>) == 64 && sizeof(LeanGate<float>) == 48, "LeanGate layout");

struct TmaGeom {
  int rank;       // 0: this sweep's tile is not a TMA box (cp.async gathers)
  int top_shift;  // top coordinate = (tile counter >> src[rank-1]) + (batch element << top_shift)
  unsigned char src[5], nbits[5], is_tile[5];
};

template <typename real> struct LeanArgs {
  typename C2<real>::type* psi;
  typename C2<real>::type* lam;
  const unsigned* tab;           // [n_rounds][LT_STRIDE] of this sweep, this direction
  const LeanGate<real>* gtab;    // [batch][n_exec_total]
  const typename C2<real>::type* sscal;  // [batch]: product of the nu of this sweep
  double* mom;                   // [n_exec_total][batch][4]
  int n;
  int n_exec_total;
  long long batch;
  unsigned long long index_hi;
  long long total_tiles;
  unsigned* tile_counter;  // device counter of this launch (zeroed beforehand): next unclaimed tile
  int chunk;               // consecutive tiles claimed at once (keeps a CTA on one batch element)
  int dbg;  // timing experiments only: 1 = skip the 2x2 math, 2 = skip the rounds (results are wrong in both)
  unsigned long long* trace;  // diagnostics (B200SV_LEAN_TRACE): [cta][trace_tiles][4] globaltimer stamps, else nullptr
  int trace_tiles;
  TmaGeom tg;
};

// ---- PTX helpers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <typename c2> __device__ __forceinline__ c2 lds_c2(unsigned addr);
template <> __device__ __forceinline__ double2 lds_c2<double2>(unsigned addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
  return v;
}
template <> __device__ __forceinline__ float2 lds_c2<float2>(unsigned addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_c2(unsigned addr, double2 v) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void sts_c2(unsigned addr, float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
template <int BYTES> __device__ __forceinline__ void cp_async_s(unsigned smem, const void* gmem);
template <> __device__ __forceinline__ void cp_async_s<16>(unsigned smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem), "l"(gmem) : "memory");
}
template <> __device__ __forceinline__ void cp_async_s<8>(unsigned smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem), "l"(gmem) : "memory");
}

__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mbar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned mbar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(mbar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
  // try_wait suspends the thread for a hardware time slice; a tile arrives within microseconds.  A wait that outlives ~2^26
  // slices means the TMA never completed (bad descriptor): trap instead of hanging the device.
  for (unsigned spins = 0; !mbar_try_wait(mbar, parity); ++spins)
    if (spins > (1u << 26)) __trap();
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

__device__ __forceinline__ void tma_load(const CUtensorMap* tm, unsigned dst, unsigned mbar, int rank, const int (&c)[5]) {
  const unsigned long long d = reinterpret_cast<unsigned long long>(tm);
  switch (rank) {
    case 2:
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                   "l"(d), "r"(mbar), "r"(c[0]), "r"(c[1])
                   : "memory");
      break;
    case 3:
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
                   "l"(d), "r"(mbar), "r"(c[0]), "r"(c[1]), "r"(c[2])
                   : "memory");
      break;
    case 4:
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
                   "l"(d), "r"(mbar), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3])
                   : "memory");
      break;
    default:
      asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
                   "l"(d), "r"(mbar), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4])
                   : "memory");
      break;
  }
}
__device__ __forceinline__ void tma_store(const CUtensorMap* tm, unsigned src, int rank, const int (&c)[5]) {
  const unsigned long long d = reinterpret_cast<unsigned long long>(tm);
  switch (rank) {
    case 2:
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(d), "r"(src), "r"(c[0]), "r"(c[1]) : "memory");
      break;
    case 3:
      asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(d), "r"(src), "r"(c[0]), "r"(c[1]),
                   "r"(c[2])
                   : "memory");
      break;
    case 4:
      asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(d), "r"(src), "r"(c[0]), "r"(c[1]),
                   "r"(c[2]), "r"(c[3])
                   : "memory");
      break;
    default:
      asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(d), "r"(src), "r"(c[0]),
                   "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4])
                   : "memory");
      break;
  }
}

// ---- in-register arithmetic ----------------------------------------------------------------------------------------
template <typename real, int R> struct LeanRegs {
  using c2 = typename C2<real>::type;
  static constexpr int NV = 1 << R;

  // v <- [[1, b], [g, d]] v on register bit A: 12 FP instructions per amplitude pair
  template <int A> static __device__ __forceinline__ void nmat(c2 (&v)[NV], const real* __restrict__ m) {
    const real br = m[0], bi = m[1], gr = m[2], gi = m[3], dr = m[4], di = m[5];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (j & (1 << A)) continue;
      const int k = j | (1 << A);
      // ordered so that every result can be written over its own input (4 temporaries per pair): no register moves
      // where the skipped-slot path and this one meet
      real t1 = gr * v[j].x, t2 = gr * v[j].y;
      t1 = fma(-gi, v[j].y, t1);
      t2 = fma(gi, v[j].x, t2);
      v[j].x = fma(br, v[k].x, v[j].x);
      v[j].y = fma(br, v[k].y, v[j].y);
      v[j].x = fma(-bi, v[k].y, v[j].x);
      v[j].y = fma(bi, v[k].x, v[j].y);
      t1 = fma(dr, v[k].x, t1);
      t2 = fma(di, v[k].x, t2);
      v[k].x = fma(-di, v[k].y, t1);
      v[k].y = fma(dr, v[k].y, t2);
    }
  }
  // four moments Im<l|P|p>, P = I, X, Y, Z on register bit A (pure FMA chains, 12 per amplitude pair)
  template <int A> static __device__ __forceinline__ void moments(const c2 (&p)[NV], const c2 (&l)[NV], double (&mo)[4]) {
    // two independent accumulator sets (even / odd pairs): 8 FMA chains in flight instead of 4 — at 4 warps per scheduler
    // the FP64 result latency is otherwise exposed (ncu: "wait" stalls on these lines)
    real s00 = 0, s11 = 0, x_im = 0, y_re = 0, s00b = 0, s11b = 0, x_imb = 0, y_reb = 0;
    int pair = 0;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (j & (1 << A)) continue;
      const int k = j | (1 << A);
      if ((pair++ & 1) == 0) {
        s00 = fma(l[j].x, p[j].y, s00); s00 = fma(-l[j].y, p[j].x, s00);
        s11 = fma(l[k].x, p[k].y, s11); s11 = fma(-l[k].y, p[k].x, s11);
        x_im = fma(l[j].x, p[k].y, x_im); x_im = fma(-l[j].y, p[k].x, x_im);
        x_im = fma(l[k].x, p[j].y, x_im); x_im = fma(-l[k].y, p[j].x, x_im);
        y_re = fma(l[k].x, p[j].x, y_re); y_re = fma(l[k].y, p[j].y, y_re);
        y_re = fma(-l[j].x, p[k].x, y_re); y_re = fma(-l[j].y, p[k].y, y_re);
      } else {
        s00b = fma(l[j].x, p[j].y, s00b); s00b = fma(-l[j].y, p[j].x, s00b);
        s11b = fma(l[k].x, p[k].y, s11b); s11b = fma(-l[k].y, p[k].x, s11b);
        x_imb = fma(l[j].x, p[k].y, x_imb); x_imb = fma(-l[j].y, p[k].x, x_imb);
        x_imb = fma(l[k].x, p[j].y, x_imb); x_imb = fma(-l[k].y, p[j].x, x_imb);
        y_reb = fma(l[k].x, p[j].x, y_reb); y_reb = fma(l[k].y, p[j].y, y_reb);
        y_reb = fma(-l[j].x, p[k].x, y_reb); y_reb = fma(-l[j].y, p[k].y, y_reb);
      }
    }
    s00 += s00b; s11 += s11b; x_im += x_imb; y_re += y_reb;
    mo[0] = (double)s00 + (double)s11; mo[1] = (double)x_im; mo[2] = (double)y_re; mo[3] = (double)s00 - (double)s11;
  }
};

template <int I, int N, typename F> __device__ __forceinline__ void static_for_up(F&& f) {
  if constexpr (I < N) { f(std::integral_constant<int, I>{}); static_for_up<I + 1, N>(f); }
}
template <int I, typename F> __device__ __forceinline__ void static_for_down(F&& f) {
  if constexpr (I >= 0) { f(std::integral_constant<int, I>{}); static_for_down<I - 1>(f); }
}

// launch geometry per instantiation: CTAs per SM the register allocation targets (threads per CTA = 2^(M-R))
#ifndef B200SV_LEAN_REGS_FWD_D_R4
#define B200SV_LEAN_REGS_FWD_D_R4 128  // complex128 forward, r = 4: 8 CTAs of 64 threads (2^10-amplitude tiles, packed layout); measured 26.2 vs 28.1 ms per pass at 168
#endif
#ifndef B200SV_LEAN_REGS_64
#define B200SV_LEAN_REGS_64 128  // 64 data registers (complex128 adjoint r = 3, complex64 r = 4 / adjoint r = 3 ...): 4 CTAs of 128 threads
#endif
#ifndef B200SV_LEAN_REGS_ADJ_D_R2
#define B200SV_LEAN_REGS_ADJ_D_R2 80  // experimental r = 2 adjoint tiles: 6 warps per scheduler
#endif
template <typename real, int R, int M, bool ADJ> struct LeanCfg {
  static constexpr int NT = 1 << (M - R);
  static constexpr int DATA_REGS = (1 << R) * (ADJ ? 2 : 1) * (int)(2 * sizeof(real) / 4);
  // registers per thread the kernel may use = 65536 / (NT * MINB), capped at 255
  // registers per thread the kernel may use = 65536 / (NT * MINB), capped at 255
  static constexpr int REGS_WANTED = DATA_REGS >= 128 ? 255 : (DATA_REGS >= 64 ? ((sizeof(real) == 8 && !ADJ && R == 4) ? B200SV_LEAN_REGS_FWD_D_R4 : B200SV_LEAN_REGS_64)
                                                                : ((sizeof(real) == 8 && ADJ && R == 2) ? B200SV_LEAN_REGS_ADJ_D_R2 : 128));
  static constexpr int MINB_RAW = 65536 / (NT * REGS_WANTED);
  static constexpr int MINB = MINB_RAW < 1 ? 1 : MINB_RAW;
  // Shared-memory layout.  ALIGNED: the tile buffers are aligned to their own size inside the window, so a thread's address is
  // base | offset and the per-amplitude masks are XOR-ed straight into it (cheapest addressing) — at the price of up to one
  // tile of slack per CTA.  Packed (tiles first, tables behind, address = base + (offset ^ mask), one more add per access):
  // taken where the slack would cost resident CTAs (B200: 227 KB per SM, 1 KB reserved per CTA).
  static constexpr int TILE_BYTES = (1 << M) * 2 * (int)sizeof(real);
  static constexpr int ALIGNED_SMEM = ((ADJ ? 2 : 1) + 1) * TILE_BYTES + 1024;
  static constexpr bool ALIGNED = MINB * ALIGNED_SMEM <= 227 * 1024;
};

template <typename real, int R, bool ADJ> __host__ __device__ inline size_t lean_table_bytes(int n_rounds, int n_exec, int nt) {
  size_t s = 64;                                               // 2 mbarriers, sweep scalar
  s += (size_t)(n_exec + 1) * sizeof(LeanGate<real>);
  if (ADJ) s += (size_t)n_exec * (nt / 32) * 4 * sizeof(double);
  s += (size_t)n_rounds * LT_STRIDE * sizeof(unsigned);
  s += (size_t)((n_rounds + 3) & ~3) * sizeof(unsigned);
  s += 128 * sizeof(unsigned long long);
  return (s + 15) & ~(size_t)15;
}
template <typename real, int R, int M, bool ADJ> inline size_t lean_smem_bytes(int n_rounds, int n_exec) {
  const size_t tile_bytes = ((size_t)1 << M) * 2 * sizeof(real);
  const size_t tables = lean_table_bytes<real, R, ADJ>(n_rounds, n_exec, 1 << (M - R));
  if (!LeanCfg<real, R, M, ADJ>::ALIGNED) return (size_t)(ADJ ? 2 : 1) * tile_bytes + tables;
  // aligned buffers: one tile of slack, split before / after the buffers; the tables go into whichever part holds them (the
  // kernel decides from the actual window address)
  size_t slack = tile_bytes;
  while (slack / 2 < tables && slack < 4 * tile_bytes) slack += tile_bytes;
  return (size_t)(ADJ ? 2 : 1) * tile_bytes + slack;
}

template <int R> __device__ __forceinline__ int lean_reduce_index(int lane) {
  // value kept by a lane after the levels 16, 8, 4, 2 (level 1 leaves both lanes of a pair with the total); -1: padding
  static_assert(4 * R <= 16, "one value per lane pair");
  int n = 4 * R, base = 0;
#pragma unroll
  for (int bit = 4; bit >= 1; --bit) {
    const int half = (n + 1) >> 1;
    if ((lane >> bit) & 1) { base += half; n = n - half; }
    else n = half;
  }
  return n >= 1 ? base : -1;
}

template <int R> __device__ __forceinline__ double lean_reduce_moments(double (&mo)[4 * R], int lane) {
  constexpr int N0 = 4 * R;
  double v[N0];
#pragma unroll
  for (int i = 0; i < N0; ++i) v[i] = mo[i];
  int n = N0;  // compile-time after unrolling
#pragma unroll
  for (int bit = 4; bit >= 1; --bit) {
    if (n > 1) {
      const int half = (n + 1) >> 1;  // low lanes keep [0, half), high lanes keep [half, n) (padded with zero)
      const bool hi = ((lane >> bit) & 1) != 0;
#pragma unroll
      for (int i = 0; i < N0; ++i) {
        if (i < half) {
          const double other_hi = (i + half < n) ? v[i + half] : 0.0;
          const double keep = hi ? other_hi : v[i];
          const double send = hi ? v[i] : other_hi;
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1 << bit);
        }
      }
      n = half;
    } else {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1 << bit);
    }
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
  return v[0];
}

// ------------------------------------------------------------------------------------------------------------------
template <typename real, int R, int M, bool ADJ>
__global__ void __launch_bounds__(LeanCfg<real, R, M, ADJ>::NT, LeanCfg<real, R, M, ADJ>::MINB)
lean_sweep_kernel(const __grid_constant__ LeanArgs<real> A, const __grid_constant__ b200sv_sweep sw,
                  const __grid_constant__ CUtensorMap tm_psi, const __grid_constant__ CUtensorMap tm_lam) {
  using c2 = typename C2<real>::type;
  using RG = LeanRegs<real, R>;
  constexpr int NV = 1 << R;
  constexpr int NT = 1 << (M - R);
  constexpr int NWARP = NT / 32;
  constexpr int NSTATE = ADJ ? 2 : 1;
  constexpr unsigned TILE_BYTES = (1u << M) * (unsigned)sizeof(c2);
  static_assert(NT >= 32 && NT <= 256, "lean sweeps run with 32..256 threads");
  static_assert(TILE_BYTES <= 65536u, "packed 16-bit byte offsets");
  extern __shared__ __align__(1024) unsigned char lean_smem[];
  unsigned char* const smem_raw = lean_smem;

  const int tid = threadIdx.x;
  const int n_rounds = sw.n_rounds, n_exec = sw.n_exec;
  const bool use_tma = A.tg.rank > 0;

  // ---- shared-memory carve-up (LeanCfg::ALIGNED): buffers aligned to TILE_BYTES with the tables in the slack, or packed
  constexpr bool ALIGNED = LeanCfg<real, R, M, ADJ>::ALIGNED;
  const unsigned sbase = smem_u32(smem_raw);
  const unsigned buf_bytes = (unsigned)NSTATE * TILE_BYTES;
  const unsigned tbl_bytes = (unsigned)lean_table_bytes<real, R, ADJ>(n_rounds, n_exec, NT);
  const unsigned pad = ALIGNED ? ((0u - sbase) & (TILE_BYTES - 1u)) : 0u;
  unsigned char* tbl = (ALIGNED && pad >= tbl_bytes) ? smem_raw : smem_raw + pad + buf_bytes;
  unsigned char* const tile_ptr = smem_raw + pad;
  const unsigned tile_psi = sbase + pad;
  const unsigned mbar0 = smem_u32(tbl);
  volatile long long* s_tile = reinterpret_cast<volatile long long*>(tbl + 16);  // the tile this CTA works on next (-1: none)
  c2* sscal_s = reinterpret_cast<c2*>(tbl + 32);
  LeanGate<real>* gt = reinterpret_cast<LeanGate<real>*>(tbl + 64);
  double* gsum = reinterpret_cast<double*>(gt + n_exec + 1);
  unsigned* tab_s = reinterpret_cast<unsigned*>(gsum + (ADJ ? n_exec * NWARP * 4 : 0));
  unsigned* ext_s = tab_s + n_rounds * LT_STRIDE;
  unsigned long long* t_lo = reinterpret_cast<unsigned long long*>(ext_s + ((n_rounds + 3) & ~3));
  unsigned long long* t_hi = t_lo + 64;

  const long long tiles_per_state = 1ll << sw.n_outer;

  // ---- one-time prologue ----
  for (int w = tid; w < n_rounds * LT_STRIDE; w += NT) tab_s[w] = A.tab[w];
  if (!use_tma) {
    for (int v = tid; v < 64; v += NT) {
      unsigned long long lo = 0, hi = 0;
      for (int i = 0; i < 6; ++i) {
        if (i < M && ((v >> i) & 1)) lo |= 1ull << sw.tile_bits[i];
        if (i + 6 < M && ((v >> i) & 1)) hi |= 1ull << sw.tile_bits[i + 6];
      }
      t_lo[v] = lo; t_hi[v] = hi;
    }
  }
  if (ADJ) for (int g = tid; g < 4 * n_exec * NWARP; g += NT) gsum[g] = 0.0;
  if (tid == 0) {
    mbar_init(mbar0, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  auto tile_base = [&](long long t) -> unsigned long long {  // CTA-index bits of tile t deposited into the basis index
    const unsigned long long c = (unsigned long long)(t % tiles_per_state);
    unsigned long long base = 0;
    if (sw.n_seg >= 0) {
      for (int k = 0; k < sw.n_seg; ++k) base |= ((c >> sw.seg[k].src) & ((1ull << sw.seg[k].len) - 1ull)) << sw.seg[k].dst;
    } else {
      for (int k = 0; k < sw.n_outer; ++k) base |= ((c >> k) & 1ull) << sw.outer_bits[k];
    }
    return base;
  };
  auto tma_coords = [&](long long t, int (&crd)[5]) {
    const unsigned long long c = (unsigned long long)(t % tiles_per_state);
    const long long b = t / tiles_per_state;
#pragma unroll
    for (int d = 0; d < 5; ++d) {
      crd[d] = 0;
      if (d < A.tg.rank && !A.tg.is_tile[d]) crd[d] = (int)((c >> A.tg.src[d]) & ((1ull << A.tg.nbits[d]) - 1ull));
    }
    crd[A.tg.rank - 1] += (int)(b << A.tg.top_shift);
  };
  auto issue_load_tma = [&](long long t) {  // thread 0 only
    int crd[5];
    tma_coords(t, crd);
    mbar_expect_tx(mbar0, NSTATE * TILE_BYTES);
    tma_load(&tm_psi, tile_psi, mbar0, A.tg.rank, crd);
    if constexpr (ADJ) tma_load(&tm_lam, tile_psi + TILE_BYTES, mbar0, A.tg.rank, crd);
  };
  auto issue_load_manual = [&](long long t, unsigned long long base) {  // all threads: coalesced gathers into the swizzled layout
    const long long off = ((t / tiles_per_state) << A.n) + (long long)base;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int e = i * NT + tid;
      const long long g = off + (long long)(t_lo[e & 63] + t_hi[e >> 6]);
      const unsigned dst = tile_psi + (unsigned)swz<real>(e) * (unsigned)sizeof(c2);
      cp_async_s<sizeof(c2)>(dst, A.psi + g);
      if constexpr (ADJ) cp_async_s<sizeof(c2)>(dst + TILE_BYTES, A.lam + g);
    }
    cp_async_commit();
  };
  auto flush_moments = [&](long long bb) {
    for (int x = tid; x < n_exec; x += NT) {
      if ((gt[x].gr & 0xff) == GR_NONE) continue;
      double* dst = A.mom + ((long long)(sw.first_exec + x) * A.batch + bb) * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        double acc = 0.0;
        for (int w = 0; w < NWARP; ++w) { acc += gsum[(x * NWARP + w) * 4 + q]; gsum[(x * NWARP + w) * 4 + q] = 0.0; }
        atomicAdd(dst + q, acc);
      }
    }
  };
  // Dynamic tile scheduling: CTAs claim chunks of consecutive tiles from a device counter.  The warp schedulers favour
  // the older of the CTAs that share an SM (measured: 17 / 22 / 33 / 49 us per tile for the 1st .. 4th CTA of an SM), so a
  // static split leaves the late CTAs running alone at the end of the sweep.
  long long chunk_end = 0;  // thread 0 only
  auto claim = [&](long long after) -> long long {  // thread 0 only
    if (after >= 0 && after + 1 < chunk_end) return after + 1;
    const long long c = (long long)atomicAdd(A.tile_counter, (unsigned)A.chunk);
    if (c >= A.total_tiles) return -1;
    chunk_end = c + A.chunk < A.total_tiles ? c + A.chunk : A.total_tiles;
    return c;
  };
  if (tid == 0) {
    const long long t0 = claim(-1);
    s_tile[0] = t0;
    if (use_tma && t0 >= 0) issue_load_tma(t0);
  }
  __syncthreads();
  if (!use_tma && s_tile[0] >= 0) issue_load_manual(s_tile[0], tile_base(s_tile[0]));

  long long cur_b = -1;
  unsigned phase0 = 0;
  long long traced = 0;
  while (true) {
    const long long t = s_tile[0];
    if (t < 0) break;
    const long long b = t / tiles_per_state;
    const unsigned long long cta_base = tile_base(t);
    const unsigned long long cta_index = cta_base | A.index_hi;

    if (b != cur_b) {  // new batch element: its normalised gates and the sweep scalar
      if (ADJ && cur_b >= 0) { flush_moments(cur_b); __syncthreads(); }
      constexpr int W = sizeof(LeanGate<real>) / 16;
      const uint4* src = reinterpret_cast<const uint4*>(A.gtab + ((long long)b * A.n_exec_total + sw.first_exec));
      uint4* dst = reinterpret_cast<uint4*>(gt);
      for (int w = tid; w < n_exec * W; w += NT) dst[w] = src[w];
      if (tid == 0) {
        sscal_s[0] = A.sscal[b];
        LeanGate<real> dummy = {};
        gt[n_exec] = dummy;
      }
      cur_b = b;
    }
    for (int rr = tid; rr < n_rounds; rr += NT) {  // per-tile constant of the permutation side: external control bits
      const unsigned* T = tab_s + rr * LT_STRIDE;
      const unsigned w = T[43];
      const int npx = (int)(w & 7u);
      unsigned e = 0;
      for (int k = 0; k < npx; ++k)
        if ((cta_index >> ((w >> (8 + 6 * k)) & 63u)) & 1ull) e ^= T[44 + k];
      ext_s[rr] = e;
    }
    if (use_tma) { mbar_wait(mbar0, phase0); phase0 ^= 1u; }
    else cp_async_wait<0>();
    __syncthreads();  // tile landed (all threads' gathers), gates / ext visible
    const bool tracing = A.trace != nullptr && tid == 0 && traced < A.trace_tiles;
    unsigned long long* trow = tracing ? A.trace + ((long long)blockIdx.x * A.trace_tiles + traced) * 4 : nullptr;
    if (tracing) {
      unsigned long long now; unsigned smid;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      trow[0] = now; trow[3] = smid;
    }
    ++traced;

    const c2 sc = sscal_s[0];
    const bool scale = !(sc.x == (real)1 && sc.y == (real)0);

    // ---- rounds ----
#pragma unroll 1
    for (int rr = 0; rr < ((A.dbg & 2) ? 0 : n_rounds); ++rr) {
      const int rloc = ADJ ? n_rounds - 1 - rr : rr;
      const unsigned* T = tab_s + rloc * LT_STRIDE;
      unsigned base = T[tid & 15] ^ T[16 + (tid >> 4)] ^ ext_s[rloc];
      const bool need_sync = (T[37] & 1u) != 0;
      unsigned ro[R];
      int ms[R];
#pragma unroll
      for (int a = 0; a < R; ++a) {
        ro[a] = T[32 + a];
        ms[a] = (int)T[38 + a];
        if (ms[a] != (int)LT_EMPTY) base ^= ro[a] & gt[ms[a]].flipmask;
      }
      // ALIGNED: the buffer is TILE_BYTES-aligned, OR == ADD and the later XORs stay inside; packed: offsets, base added per access
      const unsigned la0 = ALIGNED ? (tile_psi | (base & 0xffffu)) : (base & 0xffffu);
      const unsigned sa0 = ALIGNED ? (tile_psi | (base >> 16)) : (base >> 16);
      auto xmask = [&](int j, int sh) {
        unsigned o = 0;
#pragma unroll
        for (int a = 0; a < R; ++a)
          if ((j >> a) & 1) o ^= (ro[a] >> sh) & 0xffffu;
        return o;
      };

      c2 p[NV];
      c2 l[ADJ ? NV : 1];
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const unsigned a = la0 ^ xmask(j, 0);
#ifdef B200SV_LEAN_DBGX  // timing experiments only (results invalid): bit 4 = no tile traffic in the rounds, bit 8 = no barriers
        if (A.dbg & 4) {
          p[j].x = (real)(a & 1023u) * (real)0.001; p[j].y = (real)(j + 1) * (real)0.01;
          if constexpr (ADJ) { l[j].x = p[j].y; l[j].y = p[j].x; }
          continue;
        }
#endif
        // plain shared-window accesses (the buffers start the dynamic window): the base folds into the instruction
        if constexpr (ALIGNED) {
          p[j] = lds_c2<c2>(a);
          if constexpr (ADJ) l[j] = lds_c2<c2>(a + TILE_BYTES);
        } else {
          p[j] = *reinterpret_cast<const c2*>(tile_ptr + a);
          if constexpr (ADJ) l[j] = *reinterpret_cast<const c2*>(tile_ptr + a + TILE_BYTES);
        }
      }
#ifdef B200SV_LEAN_DBGX
      if (need_sync && !(A.dbg & 8)) __syncthreads();
#else
      if (need_sync) __syncthreads();  // every register file is loaded before anyone stores at permuted / re-laid-out addresses
#endif

      if (!(A.dbg & 1)) {
        // Full rounds (every slot holds a chain — all rounds of an HEA sweep but its first / last) take a straight-line copy of
        // the body: no per-slot tests, branches and reconvergence points (~70 of ~850 instructions of an adjoint round)
        bool full = true;
#pragma unroll
        for (int a = 0; a < R; ++a) {
          full &= ms[a] != (int)LT_EMPTY;
          if constexpr (ADJ) full &= (gt[ms[a] != (int)LT_EMPTY ? ms[a] : n_exec].gr & 0xff) == GR_CHAIN;
        }
#ifdef B200SV_LEAN_DBGX
        if (A.dbg & (16 | 32 | 64)) full = false;  // the timing experiments live in the general body; bit 64: general body only
#endif
        // measured (cfg2, ms per pass): adjoint r = 3 93.2 -> 89.5 with the straight-line copy; forward r = 4 26.2 -> 27.8 and
        // adjoint r = 4 85.6 -> 88.4 (the second copy costs registers there): taken for the r <= 3 adjoint instantiations only
        constexpr bool HAS_FAST = ADJ && R <= 3;
        if (!HAS_FAST) full = false;
        auto math = [&](auto fast_tag) {
          constexpr bool FAST = decltype(fast_tag)::value;
          if constexpr (!ADJ) {
            static_for_up<0, R>([&](auto ic) {
              constexpr int a = decltype(ic)::value;
              if (FAST || ms[a] != (int)LT_EMPTY) RG::template nmat<a>(p, gt[ms[a]].m);
            });
          } else {
            // The chains of one round act on different qubits, so <lambda|P_a|psi> does not care whether the OTHER slots
            // are already un-applied (their 2x2 commute with P_a and cancel between bra and ket up to the real factor the
            // resolve kernel tracks): all moments of the round are taken first, from the same registers, and reduced together.
            bool any = FAST;
            if constexpr (!FAST) {
#pragma unroll
              for (int a = 0; a < R; ++a) any |= ms[a] != (int)LT_EMPTY && (gt[ms[a]].gr & 0xff) == GR_CHAIN;
#ifdef B200SV_LEAN_DBGX
              if (A.dbg & 16) any = false;  // bit 16: no moments at all
#endif
            }
            if (any) {  // uniform across the CTA
              double mo[4 * R];
              static_for_up<0, R>([&](auto ic) {
                constexpr int a = decltype(ic)::value;
                double m4[4] = {0.0, 0.0, 0.0, 0.0};
                if (FAST || (ms[a] != (int)LT_EMPTY && (gt[ms[a]].gr & 0xff) == GR_CHAIN)) RG::template moments<a>(p, l, m4);
                mo[4 * a] = m4[0]; mo[4 * a + 1] = m4[1]; mo[4 * a + 2] = m4[2]; mo[4 * a + 3] = m4[3];
              });
              const int lane = tid & 31;
              bool reduce = true;
#ifdef B200SV_LEAN_DBGX
              if (A.dbg & 32) {  // bit 32: moments without the cross-lane reduction
                double chk = 0;
#pragma unroll
                for (int i = 0; i < 4 * R; ++i) chk += mo[i];
                if (chk == 123456.789) gsum[0] += chk;
                reduce = false;
              }
#endif
              if (reduce) {
                const double tot = lean_reduce_moments<R>(mo, lane);
                const int vi = lean_reduce_index<R>(lane);
                if ((lane & 1) == 0 && vi >= 0) {
                  const int x = ms[vi >> 2];
                  if (FAST || x != (int)LT_EMPTY) gsum[(x * NWARP + (tid >> 5)) * 4 + (vi & 3)] += tot;
                }
              }
            }
            static_for_down<R - 1>([&](auto ic) {
              constexpr int a = decltype(ic)::value;
              if (FAST || ms[a] != (int)LT_EMPTY) {
                const LeanGate<real>& g = gt[ms[a]];
                RG::template nmat<a>(p, g.m);
                RG::template nmat<a>(l, g.m);
              }
            });
          }
        };
        if constexpr (HAS_FAST) {
          if (full) math(std::true_type{});
          else math(std::false_type{});
        } else {
          math(std::false_type{});
        }
      }
      if (rr == n_rounds - 1 && scale) {  // the sweep's pending scalar, once per amplitude
#pragma unroll
        for (int j = 0; j < NV; ++j) {  // in place (one temporary), see nmat
          real t1 = p[j].x * sc.y;
          p[j].x = p[j].x * sc.x;
          p[j].x = fma(-p[j].y, sc.y, p[j].x);
          p[j].y = fma(p[j].y, sc.x, t1);
          if constexpr (ADJ) {
            t1 = l[j].x * sc.y;
            l[j].x = l[j].x * sc.x;
            l[j].x = fma(-l[j].y, sc.y, l[j].x);
            l[j].y = fma(l[j].y, sc.x, t1);
          }
        }
      }
#ifdef B200SV_LEAN_DBGX
      if (A.dbg & 4) {  // keep the arithmetic alive without the stores
        real chk = 0;
#pragma unroll
        for (int j = 0; j < NV; ++j) { chk += p[j].x + p[j].y; if constexpr (ADJ) chk += l[j].x + l[j].y; }
        if (chk == (real)123456.789) sts_c2(ALIGNED ? sa0 : tile_psi + sa0, p[0]);
      } else
#endif
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const unsigned a = sa0 ^ xmask(j, 16);
        if constexpr (ALIGNED) {
          sts_c2(a, p[j]);
          if constexpr (ADJ) sts_c2(a + TILE_BYTES, l[j]);
        } else {
          *reinterpret_cast<c2*>(tile_ptr + a) = p[j];
          if constexpr (ADJ) *reinterpret_cast<c2*>(tile_ptr + a + TILE_BYTES) = l[j];
        }
      }
      if (use_tma && rr == n_rounds - 1) fence_proxy_async();  // generic-proxy writes -> visible to the TMA store
#ifdef B200SV_LEAN_DBGX
      if (!(A.dbg & 8) || rr == n_rounds - 1) __syncthreads();
#else
      __syncthreads();
#endif
    }
    if (tracing) { unsigned long long now; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now)); trow[1] = now; }

    // ---- write the tile back, claim the next one ----
    if (use_tma) {
      if (tid == 0) {
        int crd[5];
        tma_coords(t, crd);
        tma_store(&tm_psi, tile_psi, A.tg.rank, crd);
        if constexpr (ADJ) tma_store(&tm_lam, tile_psi + TILE_BYTES, A.tg.rank, crd);
        bulk_commit();
        const long long nxt = claim(t);
        bulk_wait_read0();  // the store has drained the buffer
        if (nxt >= 0) issue_load_tma(nxt);
        s_tile[0] = nxt;
        if (tracing) { unsigned long long now; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now)); trow[2] = now; }
      }
      __syncthreads();  // publishes s_tile
    } else {
      const long long state_off = b << A.n;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int e = i * NT + tid;
        const long long g = state_off + (long long)(cta_base + t_lo[e & 63] + t_hi[e >> 6]);
        const unsigned src = tile_psi + (unsigned)swz<real>(e) * (unsigned)sizeof(c2);
        A.psi[g] = lds_c2<c2>(src);
        if constexpr (ADJ) A.lam[g] = lds_c2<c2>(src + TILE_BYTES);
      }
      if (tid == 0) s_tile[0] = claim(t);
      __syncthreads();
      const long long nxt = s_tile[0];
      if (nxt >= 0) issue_load_manual(nxt, tile_base(nxt));
    }
  }
  if (ADJ && cur_b >= 0) { __syncthreads(); flush_moments(cur_b); }
  if (use_tma && tid == 0) bulk_wait_read0();  // shared memory must outlive the last store's reads
}

// ------------------------------------------------------------------------------------------------------------------
// resolve kernel of the lean sweeps: one thread per (sweep, batch element) walks the sweep's M slots in EXECUTION order,
// multiplies each chain into one 2x2 (sincos evaluated here, once per launch), normalises it and keeps the running
// product of the scalars nu
// ------------------------------------------------------------------------------------------------------------------
template <typename real>
__global__ void resolve_lean_kernel(const b200sv_sweep* __restrict__ sweeps, const b200sv_round* __restrict__ rounds,
                                    const b200sv_gate* __restrict__ gates, const int* __restrict__ exec_leader,
                                    const double* __restrict__ params, const double* __restrict__ consts,
                                    LeanGate<real>* __restrict__ gtab, typename C2<real>::type* __restrict__ sscal,
                                    unsigned* __restrict__ tile_counters, int s_first, int s_count, int n_exec_total,
                                    long long batch, int dir) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)s_count * batch) return;
  const int s = s_first + (int)(t / batch);
  const long long b = t % batch;
  const b200sv_sweep& sw = sweeps[s];
  if (sw.io_mode == B200SV_IO_GENERIC) return;
  double acc_r = 1.0, acc_i = 0.0;
  for (int rr = 0; rr < sw.n_rounds; ++rr) {
    const b200sv_round& rd = rounds[sw.first_round + (dir > 0 ? rr : sw.n_rounds - 1 - rr)];
    const double round_scale = acc_r * acc_r + acc_i * acc_i;  // all moments of a round are taken before its un-applies
    for (int ai = 0; ai < sw.r; ++ai) {
      const int a = dir > 0 ? ai : sw.r - 1 - ai;
      const int x = a < 4 ? rd.mslot[a] : rd.mslot4;
      if (x < 0) continue;
      const b200sv_gate* recs = gates + exec_leader[sw.first_exec + x];
      M22 prod = elem_matrix(recs[0], params, consts, batch, b);
      bool any_grad = (recs[0].flags & B200SV_GF_GRAD) != 0;
      for (int i = 1; i < recs[0].chain; ++i) {
        prod = m22_mul(elem_matrix(recs[i], params, consts, batch, b), prod);
        any_grad |= (recs[i].flags & B200SV_GF_GRAD) != 0;
      }
      if (dir < 0) prod = m22_dag(prod);
      // pivot: largest entry -> position (0,0) through a row flip rf and a column flip cf: E[i][j] = U[i^rf][j^cf]
      int best = 0;
      double bm = -1.0;
      for (int e = 0; e < 4; ++e) {
        const double mag = prod.v[2 * e] * prod.v[2 * e] + prod.v[2 * e + 1] * prod.v[2 * e + 1];
        if (mag > bm) { bm = mag; best = e; }
      }
      const int rf = best >> 1, cf = best & 1;
      const double nr = prod.v[2 * best], ni = prod.v[2 * best + 1];
      const double inv_r = bm > 0.0 ? nr / bm : 0.0, inv_i = bm > 0.0 ? -ni / bm : 0.0;  // 1 / nu
      LeanGate<real> g;
      const int order[3] = {(0 ^ rf) * 2 + (1 ^ cf), (1 ^ rf) * 2 + (0 ^ cf), (1 ^ rf) * 2 + (1 ^ cf)};  // E01, E10, E11
      for (int k = 0; k < 3; ++k) {
        const double er = prod.v[2 * order[k]], ei = prod.v[2 * order[k] + 1];
        g.m[2 * k] = (real)(er * inv_r - ei * inv_i);
        g.m[2 * k + 1] = (real)(er * inv_i + ei * inv_r);
      }
      g.mscale = (real)round_scale;
      g.flipmask = (cf ? 0xffffu : 0u) | (rf ? 0xffff0000u : 0u);
      g.gr = ((dir < 0 && any_grad) ? GR_CHAIN : GR_NONE) | (cf << 8);
      const double tr = acc_r * nr - acc_i * ni, ti = acc_r * ni + acc_i * nr;
      acc_r = tr; acc_i = ti;
      gtab[b * n_exec_total + sw.first_exec + x] = g;
    }
  }
  typename C2<real>::type sc;
  sc.x = (real)acc_r; sc.y = (real)acc_i;
  sscal[(long long)s * batch + b] = sc;
  if (b == 0) tile_counters[s] = 0u;  // dynamic tile scheduling of this sweep's launch starts at tile 0
}

}  // namespace b200sv
