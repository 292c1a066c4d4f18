// b200sv C-ABI implementation (see include/b200sv.h).  sm_100a only; no CPU fallback.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef B200SV_PERSISTENT_DEFAULT
#define B200SV_PERSISTENT_DEFAULT 0
#endif

#include <mutex>
#include <type_traits>

#include <nvtx3/nvToolsExt.h>  // header-only (NVTX v3): ranges cost nothing unless a tool injects itself

#include "b200sv.h"
#include "sweep.cuh"
#include "lean.cuh"

using namespace b200sv;

#define CK(expr)                               \
  do {                                         \
    cudaError_t _e = (expr);                   \
    if (_e != cudaSuccess) return (int)_e;     \
  } while (0)

#define LAUNCH_CK()                            \
  do {                                         \
    cudaError_t _e = cudaGetLastError();       \
    if (_e != cudaSuccess) return (int)_e;     \
  } while (0)

namespace {

// One NVTX range per sweep launch (B200SV_NVTX=1): "b200sv fwd sweep 3 lean-tma m10 r4 rounds 6" — SURVEY.md §5's tracing hook.
struct SweepRange {
  bool on;
  SweepRange(const char* dir, int s, const b200sv_sweep& sw, const char* kind) {
    static const bool enabled = [] { const char* e = getenv("B200SV_NVTX"); return e && atoi(e) != 0; }();
    on = enabled;
    if (on) {
      char name[160];
      snprintf(name, sizeof(name), "b200sv %s sweep %d %s m%d r%d rounds %d entries %d", dir, s, kind, sw.m, sw.r, sw.n_rounds, sw.n_exec);
      nvtxRangePushA(name);
    }
  }
  ~SweepRange() { if (on) nvtxRangePop(); }
};

// ------------------------------------------------------------------------------------------------
// sweep launchers
// ------------------------------------------------------------------------------------------------
template <typename real> struct Workspace {
  ExecGate<real>* xtab;       // generic sweeps: [n_exec_total][batch]
  double* mom;                // adjoint: [n_exec_total][batch][4]
  LeanGate<real>* ltab;       // lean sweeps: [batch][n_exec_total]
  typename C2<real>::type* sscal;  // lean sweeps: [n_sweeps][batch]
  unsigned* tile_counters;    // lean sweeps: [n_sweeps] next unclaimed tile (dynamic scheduling)
};

inline size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

template <typename real>
Workspace<real> carve_workspace(void* ws, const b200sv_plan* plan, int64_t batch) {
  Workspace<real> w;
  unsigned char* p = reinterpret_cast<unsigned char*>(ws);
  const size_t ne = (size_t)plan->n_exec_total * (size_t)batch;
  w.xtab = reinterpret_cast<ExecGate<real>*>(p);
  p += align256(ne * sizeof(ExecGate<real>));
  w.mom = reinterpret_cast<double*>(p);
  p += align256(ne * 4 * sizeof(double));
  w.ltab = reinterpret_cast<LeanGate<real>*>(p);
  p += align256(ne * sizeof(LeanGate<real>));
  w.sscal = reinterpret_cast<typename C2<real>::type*>(p);
  p += align256((size_t)plan->n_sweeps * (size_t)batch * 2 * sizeof(real));
  w.tile_counters = reinterpret_cast<unsigned*>(p);
  return w;
}

template <typename real>
size_t workspace_bytes(const b200sv_plan* plan, int64_t batch, bool /*adjoint*/) {
  const size_t ne = (size_t)plan->n_exec_total * (size_t)batch;
  size_t bytes = align256(ne * sizeof(ExecGate<real>)) + align256(ne * 4 * sizeof(double)) + align256(ne * sizeof(LeanGate<real>));
  bytes += align256((size_t)plan->n_sweeps * (size_t)batch * 2 * sizeof(real));
  bytes += align256((size_t)plan->n_sweeps * sizeof(unsigned));
  return bytes + 256;
}

// per-device launch state of one kernel instantiation (the opt-in shared-memory size and the occupancy are properties of
// the (device, function) pair: a second device in the same process needs its own cudaFuncSetAttribute)
struct DeviceLaunchState {
  std::mutex mu;
  size_t configured[64] = {};
  size_t occ_smem[64] = {};
  int occ_blocks[64] = {};
};

template <typename K>
int prepare_kernel(K kern, DeviceLaunchState& stt, size_t smem, int threads, int* resident_ctas) {
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return B200SV_E_UNSUPPORTED;
  std::lock_guard<std::mutex> lock(stt.mu);
  if (smem > stt.configured[dev]) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stt.configured[dev] = smem;
  }
  if (resident_ctas) {
    if (stt.occ_blocks[dev] == 0 || stt.occ_smem[dev] != smem) {
      int sms = 0, per_sm = 0;
      CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
      stt.occ_blocks[dev] = sms * (per_sm > 0 ? per_sm : 1);
      stt.occ_smem[dev] = smem;
    }
    *resident_ctas = stt.occ_blocks[dev];
  }
  return 0;
}

template <typename real, int R, bool ADJ, bool GEN, bool PEER = false>
int launch_sweep(void* psi, void* lam, const double* params, double* grads, const b200sv_plan* plan,
                 const Workspace<real>& ws, int s, int n, int64_t batch, uint64_t index_hi, int dir, cudaStream_t st,
                 const PeerTable* peers = nullptr) {
  using c2 = typename C2<real>::type;
  const b200sv_sweep& sw = plan->sweeps_host[s];
  SweepRange nvtx_range(ADJ ? "adj" : "fwd", s, sw, PEER ? "generic-peer" : "generic");
  static const int persistent = [] { const char* e = getenv("B200SV_PERSISTENT"); return e ? atoi(e) : B200SV_PERSISTENT_DEFAULT; }();
  const int nbuf = persistent == 1 ? 2 : 1;  // 1: persistent + double buffer, 2: persistent, single buffer (table reuse only)
  static const int dbg_flags = [] { const char* e = getenv("B200SV_DBG"); return e ? (atoi(e) & 0xff) : 0; }();
  const size_t smem = sweep_smem_bytes<real, ADJ>(sw.m, sw.n_exec, sw.n_rounds, nbuf);
  auto kern = sweep_kernel<real, R, ADJ, GEN, PEER>;
  static DeviceLaunchState lstate;  // per instantiation, per device inside
  const int nt = 1 << (sw.m - R);
  if (nt > SweepCfg<real, R, ADJ>::MAXT) return B200SV_E_ARG;
  const long long total = PEER ? peers->n_tiles : (((long long)batch) << sw.n_outer);
  if (total <= 0) return PEER ? 0 : B200SV_E_ARG;  // a rank may have no tile of a very small sweep
  int resident = 0;  // persistent grid: CTAs that fit on the device at once
  if (int rc = prepare_kernel(kern, lstate, smem, nt, persistent ? &resident : nullptr)) return rc;
  long long grid = (!persistent || total < resident) ? total : resident;
  const long long per_cta = (total + grid - 1) / grid;
  grid = (total + per_cta - 1) / per_cta;
  if (grid > 2147483647ll) return B200SV_E_ARG;
  kern<<<(unsigned)grid, nt, smem, st>>>((c2*)psi, (c2*)lam, params, grads, plan->sweeps, plan->rounds,
                                         plan->gates, plan->consts, ws.xtab, ws.mom, s, n, (long long)batch,
                                         (unsigned long long)index_hi, (dir < 0 ? 0x100 : 0) | dbg_flags, per_cta, nbuf,
                                         peers ? *peers : PeerTable{});
  LAUNCH_CK();
  return 0;
}

// sweeps [first, last) of a plan over the FULL index space of a register sharded over `n_peers` GPUs (peer-memory
// addressing, see PeerTable in sweep.cuh), forward or adjoint; the caller separates sweeps by a cross-GPU barrier.  In the
// adjoint direction every rank accumulates the gradient contributions of ITS tiles; the caller all-reduces `grads`.
template <typename real, bool ADJ>
int run_plan_peers(void* const* peer_states, void* const* peer_lams, int n_peers, int rank, int n_total, const double* params,
                   double* grads, const b200sv_plan* plan, int first, int last, void* workspace, cudaStream_t st) {
  if (first >= last) return 0;
  int g = 0;
  while ((1 << g) < n_peers) ++g;
  if ((1 << g) != n_peers || n_peers > 16 || rank < 0 || rank >= n_peers) return B200SV_E_ARG;
  for (int i = 0; i < n_peers; ++i)
    if (!peer_states[i] || (ADJ && !peer_lams[i])) return B200SV_E_ARG;
  const Workspace<real> ws = carve_workspace<real>(workspace, plan, 1);
  const int x_first = plan->sweeps_host[first].first_exec;
  const int x_count = plan->sweeps_host[last - 1].first_exec + plan->sweeps_host[last - 1].n_exec - x_first;
  if (x_count > 0) {
    resolve_kernel<real><<<(unsigned)((x_count + 127) / 128), 128, 0, st>>>(plan->gates, plan->exec_leader, params, plan->consts, ws.xtab,
                                                                          x_first, x_count, 1ll, ADJ ? -1 : +1, ADJ ? 1 : 0);
    LAUNCH_CK();
    if (ADJ) CK(cudaMemsetAsync(ws.mom + (size_t)x_first * 4, 0, (size_t)x_count * 4 * sizeof(double), st));
  }
  for (int k = 0; k < last - first; ++k) {
    const int s = ADJ ? last - 1 - k : first + k;
    const b200sv_sweep& sw = plan->sweeps_host[s];
    if (sw.m < sw.r || sw.m - sw.r > 8 || sw.m > 12 || sw.n_outer != n_total - sw.m) return B200SV_E_ARG;
    PeerTable pt{};
    for (int i = 0; i < n_peers; ++i) { pt.p[i] = peer_states[i]; pt.pl[i] = ADJ ? peer_lams[i] : nullptr; }
    pt.n_local = n_total - g;
    const long long tiles = 1ll << sw.n_outer;
    const long long per_rank = (tiles + n_peers - 1) / n_peers;
    pt.tile_offset = per_rank * rank;
    pt.n_tiles = pt.tile_offset >= tiles ? 0 : (pt.tile_offset + per_rank > tiles ? tiles - pt.tile_offset : per_rank);
    const bool gen = sw.n_gen > 0;
    void* lam0 = ADJ ? peer_lams[rank] : nullptr;
    int rc = B200SV_E_ARG;
#define B200SV_LAUNCH_PEER(RR)                                                                                                            \
  rc = gen ? launch_sweep<real, RR, ADJ, true, true>(peer_states[rank], lam0, params, grads, plan, ws, s, n_total, 1, 0, ADJ ? -1 : +1, st, &pt) \
           : launch_sweep<real, RR, ADJ, false, true>(peer_states[rank], lam0, params, grads, plan, ws, s, n_total, 1, 0, ADJ ? -1 : +1, st, &pt)
    switch (sw.r) {
      case 3: B200SV_LAUNCH_PEER(3); break;
      case 4: B200SV_LAUNCH_PEER(4); break;
      default: return B200SV_E_ARG;  // peer sweeps are built with the large-state tile geometry only
    }
#undef B200SV_LAUNCH_PEER
    if (rc) return rc;
  }
  if (ADJ && x_count > 0) {
    gradient_kernel<real><<<(unsigned)((x_count + 127) / 128), 128, 0, st>>>(plan->gates, plan->exec_leader, params, plan->consts, ws.xtab,
                                                                           ws.mom, grads, x_first, x_count, 1ll);
    LAUNCH_CK();
  }
  return 0;
}

template <typename real, bool ADJ>
int launch_sweep_r(void* psi, void* lam, const double* params, double* grads, const b200sv_plan* plan,
                   const Workspace<real>& ws, int s, int n, int64_t batch, uint64_t index_hi, int dir, cudaStream_t st) {
  const b200sv_sweep& sw = plan->sweeps_host[s];
  if (sw.m < sw.r || sw.m - sw.r > 8 || sw.m > 12 || sw.n_outer != n - sw.m) return B200SV_E_ARG;
  bool gen = false;  // any generic-phase entry in this sweep?
  int n_m = 0;
  {
    // resolved entries that are not matrix-phase slots belong to the generic phase; the planner stores the
    // count of generic entries per round, their sum is n_exec minus the number of M slots — recomputed from
    // the host copy of the sweep: n_gen is kept in `sw.n_gen`.
    static const int force_gen = [] { const char* e = getenv("B200SV_FORCE_GEN"); return e ? atoi(e) : 0; }();
    gen = sw.n_gen > 0 || force_gen;
    (void)n_m;
  }
#define B200SV_LAUNCH(RR)                                                                                                   \
  return gen ? launch_sweep<real, RR, ADJ, true>(psi, lam, params, grads, plan, ws, s, n, batch, index_hi, dir, st)           \
             : launch_sweep<real, RR, ADJ, false>(psi, lam, params, grads, plan, ws, s, n, batch, index_hi, dir, st)
  switch (sw.r) {
    case 1: B200SV_LAUNCH(1);
    case 2: B200SV_LAUNCH(2);
    case 3: B200SV_LAUNCH(3);
    case 4: B200SV_LAUNCH(4);
    default: return B200SV_E_ARG;
  }
#undef B200SV_LAUNCH
}

// ------------------------------------------------------------------------------------------------
// lean sweeps: TMA tensor maps, persistent launch
// ------------------------------------------------------------------------------------------------
typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

TmapEncodeFn tmap_encode_fn() {
  static TmapEncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      p = nullptr;
    }
    return reinterpret_cast<TmapEncodeFn>(p);
  }();
  return fn;
}

// The state [batch][2^n] (complex, `amp_bytes` per amplitude) as the tensor `tm` describes it (see b200sv_tma): elements are
// 8 bytes wide, so a complex128 amplitude is two elements and dimension 0 carries one extra bit.
int make_tensor_map(CUtensorMap* out, void* state, int amp_bytes, int64_t batch, const b200sv_tma& tm) {
  TmapEncodeFn enc = tmap_encode_fn();
  if (!enc) return B200SV_E_UNSUPPORTED;
  const int rank = tm.rank & 0xff;
  const bool swizzle = (tm.rank & 0x100) != 0;
  if (rank < 2 || rank > 5 || tm.start[0] != 0 || !tm.is_tile[0]) return B200SV_E_ARG;
  const int ebit = amp_bytes == 16 ? 1 : 0;
  cuuint64_t size[5], stride[5];
  cuuint32_t box[5], estr[5];
  for (int d = 0; d < rank; ++d) {
    const int bits = tm.nbits[d] + (d == 0 ? ebit : 0);
    size[d] = 1ull << bits;
    if (d == rank - 1) size[d] = (cuuint64_t)batch << tm.nbits[d];
    box[d] = tm.is_tile[d] ? (cuuint32_t)size[d] : 1u;
    estr[d] = 1;
    if (box[d] > 256u || size[d] > (1ull << 32)) return B200SV_E_UNSUPPORTED;
    if (d >= 1) {
      stride[d - 1] = (cuuint64_t)amp_bytes << tm.start[d];
      if (stride[d - 1] % 16 != 0 || stride[d - 1] >= (1ull << 40)) return B200SV_E_UNSUPPORTED;
    }
  }
  const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT64, (cuuint32_t)rank, state, size, stride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : B200SV_E_ARG;
}

// (real, R, M, ADJ) combinations lean_sweep_kernel is compiled for
#ifdef B200SV_LEAN_EXPERIMENTAL  // tuning builds only (tools/gpu_call12.sh)
#define B200SV_LEAN_EXTRA(X) X(double, 5, 11, false) X(double, 5, 10, false) X(double, 4, 9, true)
#else
#define B200SV_LEAN_EXTRA(X)
#endif
#define B200SV_LEAN_INSTANCES(X)                                                                          \
  B200SV_LEAN_EXTRA(X)                                                                                    \
  X(double, 4, 11, false) X(double, 5, 12, false) X(double, 4, 10, false)                                 \
  X(double, 4, 10, true) X(double, 3, 10, true) X(double, 4, 11, true) X(double, 3, 11, true)             \
  X(float, 4, 12, false) X(float, 5, 12, false) X(float, 4, 11, false)                                    \
  X(float, 3, 11, true) X(float, 4, 12, true) X(float, 4, 11, true)

template <typename real, int R, int M, bool ADJ>
int launch_lean(void* psi, void* lam, const b200sv_plan* plan, const Workspace<real>& ws, int s, int n, int64_t batch,
                uint64_t index_hi, cudaStream_t st) {
  using c2 = typename C2<real>::type;
  const b200sv_sweep& sw = plan->sweeps_host[s];
  SweepRange nvtx_range(ADJ ? "adj" : "fwd", s, sw, sw.io_mode == B200SV_IO_LEAN_TMA ? "lean-tma" : "lean-gather");
  static const int dbg_flags = [] { const char* e = getenv("B200SV_DBG"); return e ? (atoi(e) & 0xff) : 0; }();
  static const int occ_env = [] { const char* e = getenv("B200SV_LEAN_CTAS_PER_SM"); return e ? atoi(e) : 0; }();
  const bool tma = sw.io_mode == B200SV_IO_LEAN_TMA;
  const size_t smem = lean_smem_bytes<real, R, M, ADJ>(sw.n_rounds, sw.n_exec);
  if (smem > 227 * 1024) return B200SV_E_UNSUPPORTED;
  auto kern = lean_sweep_kernel<real, R, M, ADJ>;
  static DeviceLaunchState lstate;
  constexpr int NT = 1 << (M - R);
  int resident = 0;
  if (int rc = prepare_kernel(kern, lstate, smem, NT, &resident)) return rc;
  if (occ_env > 0) {  // tuning experiments: cap the CTAs per SM
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (sms * occ_env < resident) resident = sms * occ_env;
  }
  const long long total = ((long long)batch) << sw.n_outer;
  if (total <= 0 || total > 0x7fffffffll) return B200SV_E_ARG;
  const long long grid = total < resident ? total : resident;
  // chunks of consecutive tiles: small, so that the tail (the last chunk of the slowest CTA) stays short; the price is a
  // reload of the ~2 KB gate table and a flush of the moment partials whenever a chunk starts in another batch element
  static const int chunk_env = [] { const char* e = getenv("B200SV_LEAN_CHUNK"); return e ? atoi(e) : 0; }();
  long long chunk = chunk_env > 0 ? chunk_env : 4;
  if (chunk > total / grid) chunk = total / grid;
  if (chunk < 1) chunk = 1;
  const long long per_cta = (total + grid - 1) / grid;  // (diagnostics header only)

  LeanArgs<real> a{};
  a.psi = (c2*)psi;
  a.lam = (c2*)lam;
  a.tab = (ADJ ? plan->lean_tab_bwd : plan->lean_tab_fwd) + (size_t)sw.first_round * LT_STRIDE;
  a.gtab = ws.ltab;
  a.sscal = ws.sscal + (size_t)s * batch;
  a.mom = ws.mom;
  a.n = n;
  a.n_exec_total = plan->n_exec_total;
  a.batch = batch;
  a.index_hi = index_hi;
  a.total_tiles = total;
  a.tile_counter = ws.tile_counters + s;
  a.chunk = (int)chunk;
  a.dbg = dbg_flags;
  CUtensorMap tm_psi{}, tm_lam{};
  if (tma) {
    if (!plan->tma_host) return B200SV_E_ARG;
    const b200sv_tma& tm = plan->tma_host[s];
    if (int rc = make_tensor_map(&tm_psi, psi, (int)sizeof(c2), batch, tm)) return rc;
    if (ADJ) { if (int rc = make_tensor_map(&tm_lam, lam, (int)sizeof(c2), batch, tm)) return rc; }
    a.tg.rank = tm.rank & 0xff;
    a.tg.top_shift = tm.nbits[a.tg.rank - 1];
    for (int d = 0; d < 5; ++d) { a.tg.src[d] = tm.src[d]; a.tg.nbits[d] = tm.nbits[d]; a.tg.is_tile[d] = tm.is_tile[d]; }
  }
  // diagnostics: B200SV_LEAN_TRACE=<file prefix> records per-tile globaltimer stamps of the first 64 tiles of every CTA
  // (tile landed / rounds done / store+next load issued, SM id) and writes them to <prefix>.<sweep>.<fwd|adj>.bin.  This
  // path synchronises and allocates: never set it outside a debugging session.
  const char* trace_prefix = getenv("B200SV_LEAN_TRACE");
  unsigned long long* d_trace = nullptr;
  const int trace_tiles = 64;
  if (trace_prefix && *trace_prefix) {
    CK(cudaMalloc(&d_trace, (size_t)grid * trace_tiles * 4 * sizeof(unsigned long long)));
    CK(cudaMemsetAsync(d_trace, 0, (size_t)grid * trace_tiles * 4 * sizeof(unsigned long long), st));
    a.trace = d_trace;
    a.trace_tiles = trace_tiles;
  }
  kern<<<(unsigned)grid, NT, smem, st>>>(a, sw, tm_psi, tm_lam);
  LAUNCH_CK();
  if (d_trace) {
    CK(cudaStreamSynchronize(st));
    const size_t nwords = (size_t)grid * trace_tiles * 4;
    unsigned long long* h = (unsigned long long*)malloc(nwords * sizeof(unsigned long long));
    if (h) {
      CK(cudaMemcpy(h, d_trace, nwords * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
      char path[512];
      snprintf(path, sizeof(path), "%s.%d.%s.bin", trace_prefix, s, ADJ ? "adj" : "fwd");
      if (FILE* f = fopen(path, "wb")) {
        const long long hdr[4] = {grid, trace_tiles, per_cta, (long long)NT};
        fwrite(hdr, sizeof(hdr), 1, f);
        fwrite(h, sizeof(unsigned long long), nwords, f);
        fclose(f);
      }
      free(h);
    }
    CK(cudaFree(d_trace));
  }
  return 0;
}

template <typename real, bool ADJ>
int launch_lean_r(void* psi, void* lam, const b200sv_plan* plan, const Workspace<real>& ws, int s, int n, int64_t batch,
                  uint64_t index_hi, cudaStream_t st) {
  const b200sv_sweep& sw = plan->sweeps_host[s];
  if (sw.n_outer != n - sw.m || !plan->lean_tab_fwd || !plan->lean_tab_bwd) return B200SV_E_ARG;
#define B200SV_X(REAL, RR, MM, AA)                                                                                    \
  if constexpr (std::is_same<real, REAL>::value && ADJ == AA) {                                                       \
    if (sw.r == RR && sw.m == MM) return launch_lean<real, RR, MM, ADJ>(psi, lam, plan, ws, s, n, batch, index_hi, st); \
  }
  B200SV_LEAN_INSTANCES(B200SV_X)
#undef B200SV_X
  return B200SV_E_UNSUPPORTED;
}

bool lean_supported(int dtype, int m, int r, bool adjoint) {
#define B200SV_X(REAL, RR, MM, AA) \
  if ((dtype == B200SV_C128) == std::is_same<REAL, double>::value && adjoint == AA && r == RR && m == MM) return true;
  B200SV_LEAN_INSTANCES(B200SV_X)
#undef B200SV_X
  return false;
}

// gradient kernel: moments of every chain -> gradient of every parametric gate of the chain (generic and lean sweeps)
template <typename real>
__global__ void gradient2_kernel(const b200sv_gate* __restrict__ gates, const int* __restrict__ exec_leader,
                                 const uint8_t* __restrict__ exec_mode, const double* __restrict__ params,
                                 const double* __restrict__ consts, const ExecGate<real>* __restrict__ xtab,
                                 const LeanGate<real>* __restrict__ ltab, const double* __restrict__ mom_g,
                                 double* __restrict__ grads, int x_first, int x_count, int n_exec_total, long long batch) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)x_count * batch) return;
  const int x = x_first + (int)(t / batch);
  const long long b = t % batch;
  const b200sv_gate* rec = gates + exec_leader[x];
  const double* mom = mom_g + ((long long)x * batch + b) * 4;
  if (exec_mode[x] == B200SV_IO_GENERIC) {
    const int gr = xtab[(long long)x * batch + b].gr;
    if (gr == GR_CHAIN) chain_gradients(rec, mom, params, consts, grads, batch, b);
    else if (gr == GR_SCALE && rec->grad_row >= 0) atomicAdd(&grads[(long long)rec->grad_row * batch + b], mom[0]);
    return;
  }
  const LeanGate<real> g = ltab[b * n_exec_total + x];
  if ((g.gr & 0xff) != GR_CHAIN) return;
  // measured with both states divided by the pending scalar s (real factor |s|^2) and, after a column flip, with the
  // pair relabelled by X: X P X = P for I, X and -P for Y, Z
  const double sc = (double)g.mscale, sg = (g.gr >> 8) & 1 ? -1.0 : 1.0;
  const double m4[4] = {sc * mom[0], sc * mom[1], sg * sc * mom[2], sg * sc * mom[3]};
  chain_gradients(rec, m4, params, consts, grads, batch, b);
}

// forward (ADJ = false) or adjoint (ADJ = true) execution of sweeps [first, last)
template <typename real, bool ADJ>
int run_plan(void* psi, void* lam, int n, int64_t batch, const double* params, double* grads,
             const b200sv_plan* plan, int first, int last, uint64_t index_hi, void* workspace, cudaStream_t st) {
  if (first >= last) return 0;
  const Workspace<real> ws = carve_workspace<real>(workspace, plan, batch);
  const int x_first = plan->sweeps_host[first].first_exec;
  const int x_last = plan->sweeps_host[last - 1].first_exec + plan->sweeps_host[last - 1].n_exec;
  const int x_count = x_last - x_first;
  const long long nthreads = (long long)x_count * batch;
  const unsigned rgrid = (unsigned)((nthreads + 127) / 128);
  bool any_generic = false, any_lean = false, lean_ok = true;
  for (int s = first; s < last; ++s) {
    const b200sv_sweep& sw = plan->sweeps_host[s];
    (sw.io_mode == B200SV_IO_GENERIC ? any_generic : any_lean) = true;
    // a plan marked for the lean kernel but executed in the direction / dtype it was not built for: generic kernels
    if (sw.io_mode != B200SV_IO_GENERIC && !lean_supported(sizeof(real) == 8 ? B200SV_C128 : B200SV_C64, sw.m, sw.r, ADJ)) lean_ok = false;
  }
  if (any_lean && (!plan->exec_mode || !plan->lean_tab_fwd || !plan->lean_tab_bwd)) return B200SV_E_ARG;
  if (!lean_ok) { any_generic = true; any_lean = false; }
  if (x_count > 0 && any_generic) {
    resolve_kernel<real><<<rgrid, 128, 0, st>>>(plan->gates, plan->exec_leader, params, plan->consts, ws.xtab, x_first,
                                                x_count, (long long)batch, ADJ ? -1 : +1, ADJ ? 1 : 0);
    LAUNCH_CK();
  }
  if (any_lean) {
    const long long nt = (long long)(last - first) * batch;
    resolve_lean_kernel<real><<<(unsigned)((nt + 63) / 64), 64, 0, st>>>(plan->sweeps, plan->rounds, plan->gates, plan->exec_leader,
                                                                         params, plan->consts, ws.ltab, ws.sscal, ws.tile_counters, first,
                                                                         last - first, plan->n_exec_total, (long long)batch, ADJ ? -1 : +1);
    LAUNCH_CK();
  }
  if (ADJ && x_count > 0) CK(cudaMemsetAsync(ws.mom + (size_t)x_first * batch * 4, 0, (size_t)x_count * batch * 4 * sizeof(double), st));
  auto one = [&](int s) -> int {
    if (!any_lean || plan->sweeps_host[s].io_mode == B200SV_IO_GENERIC)
      return launch_sweep_r<real, ADJ>(psi, lam, params, grads, plan, ws, s, n, batch, index_hi, ADJ ? -1 : +1, st);
    return launch_lean_r<real, ADJ>(psi, lam, plan, ws, s, n, batch, index_hi, st);
  };
  if (!ADJ) {
    for (int s = first; s < last; ++s)
      if (int rc = one(s)) return rc;
  } else {
    for (int s = last - 1; s >= first; --s)
      if (int rc = one(s)) return rc;
    if (x_count > 0) {
      if (any_lean) {
        gradient2_kernel<real><<<rgrid, 128, 0, st>>>(plan->gates, plan->exec_leader, plan->exec_mode, params, plan->consts, ws.xtab,
                                                      ws.ltab, ws.mom, grads, x_first, x_count, plan->n_exec_total, (long long)batch);
      } else {
        gradient_kernel<real><<<rgrid, 128, 0, st>>>(plan->gates, plan->exec_leader, params, plan->consts, ws.xtab, ws.mom,
                                                     grads, x_first, x_count, (long long)batch);
      }
      LAUNCH_CK();
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
template <typename c2> __global__ void zero_state_kernel(c2* psi, int n, long long batch) {
  const long long total = batch << n;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long mask = (1ll << n) - 1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    c2 v; v.x = (i & mask) == 0 ? 1 : 0; v.y = 0;
    psi[i] = v;
  }
}

__device__ __forceinline__ double block_reduce_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
  if (w == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  __syncthreads();
  return v;  // valid in thread 0
}

struct TermS {  // per-CTA resolved term
  unsigned long long x, z;
  double cr, ci;
};

// Single-Z diagonal terms (the first n_single terms; total_magnetization, detunings, the linear part of
// N_i N_j expansions) are additive over index bits:  sum_t c_t (-1)^{bit_t(i)} = c0 - 2 sum_{bits set in i} cbit[bit],
// so a thread evaluates them once for its fixed index bits and adds a per-k correction.
struct SingleZ {
  double cbr[64], cbi[64];
  double c0r, c0i;
};

constexpr int MAX_TERMS_SMEM = 1024;
constexpr int EXP_THREADS = 256;

__device__ __forceinline__ void resolve_terms(TermS* sh, const b200sv_pterm* terms, int T,
                                              const double* coef, long long coef_ld, long long b,
                                              SingleZ* sz = nullptr, int n_single = 0) {
  if (sz != nullptr) {
    for (int q = threadIdx.x; q < 64; q += blockDim.x) { sz->cbr[q] = 0.0; sz->cbi[q] = 0.0; }
    if (threadIdx.x == 0) { sz->c0r = 0.0; sz->c0i = 0.0; }
    __syncthreads();
  }
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const b200sv_pterm tm = terms[t];
    const long long bb = coef_ld == 1 ? 0 : b;
    double cr = coef[((long long)tm.coef_row * coef_ld + bb) * 2];
    double ci = coef[((long long)tm.coef_row * coef_ld + bb) * 2 + 1];
    // multiply by (-i)^{n_y}
    switch (tm.n_y & 3) {
      case 1: { const double t0 = cr; cr = ci; ci = -t0; } break;
      case 2: cr = -cr; ci = -ci; break;
      case 3: { const double t0 = cr; cr = -ci; ci = t0; } break;
      default: break;
    }
    if (t < n_single) {  // xmask == 0, exactly one z bit (guaranteed by the caller)
      const int bit = __ffsll((long long)tm.zmask) - 1;
      atomicAdd(&sz->cbr[bit], cr); atomicAdd(&sz->cbi[bit], ci);
      atomicAdd(&sz->c0r, cr); atomicAdd(&sz->c0i, ci);
    } else {
      sh[t - n_single].x = tm.xmask; sh[t - n_single].z = tm.zmask; sh[t - n_single].cr = cr; sh[t - n_single].ci = ci;
    }
  }
}

// weight of the single-Z terms at index gi: c0 - 2 * sum over set bits
__device__ __forceinline__ void single_z_weight(const SingleZ& sz, unsigned long long gi, double& wr, double& wi) {
  double ar = 0.0, ai = 0.0;
  while (gi) {
    const int bit = __ffsll((long long)gi) - 1;
    gi &= gi - 1;
    ar += sz.cbr[bit]; ai += sz.cbi[bit];
  }
  wr = sz.c0r - 2.0 * ar; wi = sz.c0i - 2.0 * ai;
}

// Index decomposition of the streaming Pauli kernels: i = blk * 2^14 + g * 2^11 + k * 2^8 + tid  (g, k < 8).  The single-Z
// part of a Pauli sum is additive over index bits, w(i) = c0 + d(tid) + d(k << 8) + d(g << 11) + d(blk << 14 | index_hi) with
// d(x) = -2 * sum_{bits set in x} c_bit: the four pieces are computed once per CTA (tables in shared memory), so a
// total_magnetization-like observable costs three additions per amplitude instead of a loop over the set bits.
constexpr int PK_K = 8, PK_G = 8, PK_TILE = EXP_THREADS * PK_K * PK_G;

struct DiagTables {
  double kr[PK_K], ki[PK_K], gr[PK_G], gi[PK_G], br, bi;
};

__device__ __forceinline__ void diag_tables(const SingleZ& sz, DiagTables* dt, long long blk, unsigned long long index_hi, int n_single,
                                            double& tr, double& ti) {
  tr = 0.0; ti = 0.0;
  if (n_single > 0) {
    double wr, wi;
    single_z_weight(sz, (unsigned long long)threadIdx.x, wr, wi);
    tr = wr - sz.c0r; ti = wi - sz.c0i;
  }
  if (threadIdx.x < PK_K + PK_G + 1) {
    const int t = threadIdx.x;
    unsigned long long x = t < PK_K ? ((unsigned long long)t << 8) : t < PK_K + PK_G ? ((unsigned long long)(t - PK_K) << 11)
                                                                                     : (((unsigned long long)blk << 14) | index_hi);
    double wr = 0.0, wi = 0.0;
    if (n_single > 0) single_z_weight(sz, x, wr, wi);
    else { wr = sz.c0r; wi = sz.c0i; }
    if (t < PK_K) { dt->kr[t] = wr - sz.c0r; dt->ki[t] = wi - sz.c0i; }
    else if (t < PK_K + PK_G) { dt->gr[t - PK_K] = wr - sz.c0r; dt->gi[t - PK_K] = wi - sz.c0i; }
    else { dt->br = wr; dt->bi = wi; }  // includes c0 once
  }
}

// E[b] += Re sum_i conj(psi_i) sum_t c_t (-1)^{popc(i&z_t)} psi[i^x_t]
// PEER: psi is this rank's shard of a register sharded by its top qubits; the partner amplitude psi[i ^ x] of a term with
// X / Y factors on global qubits lives in another rank's shard and is read through NVLink peer memory (pt.p[owner]).
template <typename c2, bool PEER = false, bool DIAG = false>
__global__ void __launch_bounds__(EXP_THREADS)
pauli_expect_kernel(const c2* __restrict__ psi, int n, long long batch, const b200sv_pterm* __restrict__ terms,
                    int T, int n_single, const double* __restrict__ coef, long long coef_ld, double* __restrict__ out,
                    unsigned long long index_hi, long long blocks_per_state, const PeerTable pt,
                    const double* __restrict__ diag = nullptr, long long diag_ld = 1) {
  // DIAG: the real parts of `diag` [diag_ld][2^n] (b200sv_pauli_diagonal) are the weights of ALL diagonal terms; the term list
  // holds the X / Y terms only.  A constant observable with ZZ.. strings reads 16 bytes of an L2-resident table per amplitude
  // instead of evaluating every string's parity per amplitude (20 strings on 20 qubits: 5.9 ms -> the stream's 0.7 ms)
  __shared__ TermS sh[MAX_TERMS_SMEM];
  __shared__ SingleZ sz;
  __shared__ DiagTables dt;
  __shared__ double red[32];
  // DIAG with one table for the whole batch: consecutive CTAs take the SAME index range of consecutive batch elements, so a
  // table slice is read from DRAM once and served from L2 for the rest of the batch
  const bool b_inner = DIAG && diag_ld == 1;
  const long long b = b_inner ? blockIdx.x % batch : blockIdx.x / blocks_per_state;
  const long long blk = b_inner ? blockIdx.x / batch : blockIdx.x % blocks_per_state;
  resolve_terms(sh, terms, T, coef, coef_ld, b, &sz, n_single);
  __syncthreads();
  T -= n_single;
  double tr, ti;
  diag_tables(sz, &dt, blk, index_hi, n_single, tr, ti);
  __syncthreads();
  const long long N = 1ll << n;
  const c2* p = psi + (b << n);
  double acc = 0.0;
  const long long base = blk * (long long)PK_TILE + threadIdx.x;
  const double wt = tr + dt.br;
#pragma unroll 1
  for (int g = 0; g < PK_G; ++g) {
    const long long gbase = base + (long long)g * (EXP_THREADS * PK_K);
    if (gbase >= N) break;
    // all of the group's amplitudes are requested before the first one is used: the kernel is a pure stream (1·S bytes)
    // and needs several independent 16-byte loads in flight per thread to fill the HBM pipe
    c2 av[PK_K];
    double dw[DIAG ? PK_K : 1];
    const double2* dg = DIAG ? reinterpret_cast<const double2*>(diag) + ((diag_ld == 1 ? 0 : b) << n) : nullptr;
#pragma unroll
    for (int k = 0; k < PK_K; ++k) {
      const long long i = gbase + (long long)k * EXP_THREADS;
      av[k].x = 0; av[k].y = 0;
      if constexpr (DIAG) dw[k] = 0.0;
      if (i < N) {
        av[k] = __ldcs(p + i);
        if constexpr (DIAG) dw[k] = dg[i].x;
      }
    }
    const double wg = wt + dt.gr[g];
#pragma unroll
    for (int k = 0; k < PK_K; ++k) {
      const long long i = gbase + (long long)k * EXP_THREADS;
      if (i >= N) break;
      const c2 a = av[k];
      double wr = wg + dt.kr[k];  // diagonal weight of the single-Z terms
      if constexpr (DIAG) wr += dw[k];
      double sr = 0.0, si = 0.0;  // off-diagonal sum
      if (T > 0) {
        const unsigned long long gi = (unsigned long long)i | index_hi;
        for (int t = 0; t < T; ++t) {
          const TermS tm = sh[t];
          const double sg = (__popcll(gi & tm.z) & 1) ? -1.0 : 1.0;
          if (tm.x == 0ull) wr += sg * tm.cr;
          else {
            c2 q;
            if constexpr (PEER) {
              const unsigned long long j = gi ^ tm.x;
              q = reinterpret_cast<const c2*>(pt.p[j >> pt.n_local])[j & ((1ull << pt.n_local) - 1ull)];
            } else {
              q = p[i ^ (long long)tm.x];
            }
            const double cr = sg * tm.cr, ci = sg * tm.ci;
            sr += cr * (double)q.x - ci * (double)q.y;
            si += cr * (double)q.y + ci * (double)q.x;
          }
        }
      }
      const double ax = a.x, ay = a.y;
      acc += wr * (ax * ax + ay * ay) + (ax * sr + ay * si);
    }
  }
  acc = block_reduce_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(&out[b], acc);
}

// out = alpha * (sum_t c_t P_t psi) + beta * acc
template <typename c2, bool PEER = false, bool DIAG = false>
__global__ void __launch_bounds__(EXP_THREADS)
pauli_apply_kernel(const c2* __restrict__ psi, c2* out, const c2* acc, int n, long long batch,
                   const b200sv_pterm* __restrict__ terms, int T, int n_single, const double* __restrict__ coef,
                   long long coef_ld, double ar, double ai, double br, double bi,
                   unsigned long long index_hi, long long blocks_per_state, const PeerTable pt,
                   const double* __restrict__ diag = nullptr, long long diag_ld = 1, const double* __restrict__ diag_scale = nullptr) {
  // diag_scale (optional): complex factor [batch][2] of the table per batch element (the adjoint pass weights observable o by
  // the cotangent of E[b, o]; the term coefficients carry the same factor in `coef`)
  // diag (optional): precomputed complex weights [diag_ld][2^n] of ALL diagonal terms of the sum (b200sv_pauli_diagonal); the
  // term list then holds only the X / Y terms — the Krylov matvec of a Rydberg Hamiltonian (66 ZZ + 12 Z + 12 X terms)
  // evaluates 12 terms per amplitude instead of 90
  __shared__ TermS sh[MAX_TERMS_SMEM];
  __shared__ SingleZ sz;
  __shared__ DiagTables dt;
  const bool b_inner = DIAG && diag_ld == 1;  // see pauli_expect_kernel
  const long long b = b_inner ? blockIdx.x % batch : blockIdx.x / blocks_per_state;
  const long long blk = b_inner ? blockIdx.x / batch : blockIdx.x % blocks_per_state;
  resolve_terms(sh, terms, T, coef, coef_ld, b, &sz, n_single);
  __syncthreads();
  T -= n_single;
  double tr, ti;
  diag_tables(sz, &dt, blk, index_hi, n_single, tr, ti);
  __syncthreads();
  const long long N = 1ll << n;
  const c2* p = psi + (b << n);
  const long long base = blk * (long long)PK_TILE + threadIdx.x;
  const double wtr = tr + dt.br, wti = ti + dt.bi;
#pragma unroll 1
  for (int g = 0; g < PK_G; ++g) {
    const long long gbase = base + (long long)g * (EXP_THREADS * PK_K);
    if (gbase >= N) break;
    c2 av[PK_K], zv[PK_K];  // loads first (memory-level parallelism), see pauli_expect_kernel
    double2 dv[DIAG ? PK_K : 1];  // (a template parameter: the 16 extra registers cost the plain kernel a third of its speed)
    const double2* dg = DIAG ? reinterpret_cast<const double2*>(diag) + ((diag_ld == 1 ? 0 : b) << n) : nullptr;
#pragma unroll
    for (int k = 0; k < PK_K; ++k) {
      const long long i = gbase + (long long)k * EXP_THREADS;
      av[k].x = 0; av[k].y = 0; zv[k].x = 0; zv[k].y = 0;
      if constexpr (DIAG) { dv[k].x = 0.0; dv[k].y = 0.0; }
      if (i < N) {
        av[k] = __ldcs(p + i);
        if (acc != nullptr) zv[k] = acc[(b << n) + i];
        if constexpr (DIAG) dv[k] = dg[i];
      }
    }
    const double wgr = wtr + dt.gr[g], wgi = wti + dt.gi[g];
#pragma unroll
    for (int k = 0; k < PK_K; ++k) {
      const long long i = gbase + (long long)k * EXP_THREADS;
      if (i >= N) break;
      const c2 a = av[k];
      double wr = wgr + dt.kr[k], wi = wgi + dt.ki[k], sr = 0.0, si = 0.0;
      if constexpr (DIAG) {
        if (diag_scale != nullptr) {
          const double sx = diag_scale[2 * b], sy = diag_scale[2 * b + 1];
          wr += sx * dv[k].x - sy * dv[k].y;
          wi += sx * dv[k].y + sy * dv[k].x;
        } else {
          wr += dv[k].x; wi += dv[k].y;
        }
      }
      if (T > 0) {
        const unsigned long long gi = (unsigned long long)i | index_hi;
        for (int t = 0; t < T; ++t) {
          const TermS tm = sh[t];
          const double sg = (__popcll(gi & tm.z) & 1) ? -1.0 : 1.0;
          if (tm.x == 0ull) { wr += sg * tm.cr; wi += sg * tm.ci; }
          else {
            c2 q;
            if constexpr (PEER) {
              const unsigned long long j = gi ^ tm.x;
              q = reinterpret_cast<const c2*>(pt.p[j >> pt.n_local])[j & ((1ull << pt.n_local) - 1ull)];
            } else {
              q = p[i ^ (long long)tm.x];
            }
            const double cr = sg * tm.cr, ci = sg * tm.ci;
            sr += cr * (double)q.x - ci * (double)q.y;
            si += cr * (double)q.y + ci * (double)q.x;
          }
        }
      }
      sr += wr * (double)a.x - wi * (double)a.y;
      si += wr * (double)a.y + wi * (double)a.x;
      double rr = ar * sr - ai * si, ri = ar * si + ai * sr;
      if (acc != nullptr) {
        const c2 z = zv[k];
        rr += br * (double)z.x - bi * (double)z.y;
        ri += br * (double)z.y + bi * (double)z.x;
      }
      c2 o; o.x = rr; o.y = ri;
      __stcs(out + (b << n) + i, o);
    }
  }
}

// out[r][i] = sum_t coef_t(r) (-1)^{popc(i & z_t)}  (all terms diagonal): the weight table b200sv_pauli_apply_diag consumes
__global__ void __launch_bounds__(256)
pauli_diagonal_kernel(const b200sv_pterm* __restrict__ terms, int T, int n_single, const double* __restrict__ coef, long long coef_ld,
                      int n, double2* __restrict__ out, unsigned long long index_hi, long long blocks_per_state) {
  __shared__ TermS sh[MAX_TERMS_SMEM];
  __shared__ SingleZ sz;
  const long long r = blockIdx.x / blocks_per_state, blk = blockIdx.x % blocks_per_state;
  resolve_terms(sh, terms, T, coef, coef_ld, r, &sz, n_single);
  __syncthreads();
  T -= n_single;
  const long long N = 1ll << n;
  for (int k = 0; k < 4; ++k) {
    const long long i = blk * 1024ll + k * 256 + threadIdx.x;
    if (i >= N) break;
    const unsigned long long gi = (unsigned long long)i | index_hi;
    double wr = sz.c0r, wi = sz.c0i;
    if (n_single > 0) single_z_weight(sz, gi, wr, wi);
    for (int t = 0; t < T; ++t) {
      const TermS tm = sh[t];
      const double sg = (__popcll(gi & tm.z) & 1) ? -1.0 : 1.0;
      wr += sg * tm.cr; wi += sg * tm.ci;
    }
    out[(r << n) + i] = make_double2(wr, wi);
  }
}

// psi[i] *= exp(-i * sign * t_b * d(i))
template <typename c2>
__global__ void __launch_bounds__(256)
diag_evolve_kernel(c2* psi, int n, long long batch, const b200sv_pterm* __restrict__ terms, int T, int n_single,
                   const double* __restrict__ coef, long long coef_ld, const double* __restrict__ tt,
                   long long t_ld, double sign, unsigned long long index_hi, long long blocks_per_state) {
  __shared__ TermS sh[MAX_TERMS_SMEM];
  __shared__ SingleZ sz;
  const long long b = blockIdx.x / blocks_per_state;
  const long long blk = blockIdx.x % blocks_per_state;
  resolve_terms(sh, terms, T, coef, coef_ld, b, &sz, n_single);
  __syncthreads();
  T -= n_single;
  const double t = tt[t_ld == 1 ? 0 : b] * sign;
  const long long N = 1ll << n;
  c2* p = psi + (b << n);
  const long long base = blk * (long long)(256 * 4);
#pragma unroll 1
  for (int k = 0; k < 4; ++k) {
    const long long i = base + (long long)k * 256 + threadIdx.x;
    if (i >= N) break;
    const unsigned long long gi = (unsigned long long)i | index_hi;
    double d = 0.0;
    if (n_single > 0) { double di; single_z_weight(sz, gi, d, di); }
    for (int q = 0; q < T; ++q) {
      const TermS tm = sh[q];
      d += (__popcll(gi & tm.z) & 1) ? -tm.cr : tm.cr;
    }
    double s, c;
    sincos(-t * d, &s, &c);
    const c2 a = p[i];
    c2 o; o.x = (double)a.x * c - (double)a.y * s; o.y = (double)a.x * s + (double)a.y * c;
    p[i] = o;
  }
}

// ---- batched BLAS-1 --------------------------------------------------------------------------
template <typename c2>
__global__ void __launch_bounds__(256)
dot_kernel(const c2* __restrict__ x, const c2* __restrict__ y, int n, double* out, long long blocks_per_state) {
  __shared__ double red[32];
  const long long b = blockIdx.x / blocks_per_state;
  const long long blk = blockIdx.x % blocks_per_state;
  const long long N = 1ll << n;
  double re = 0.0, im = 0.0;
  const long long base = blk * 2048ll;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const long long i = base + k * 256 + threadIdx.x;
    if (i < N) {
      const c2 a = x[(b << n) + i], c = y[(b << n) + i];
      re += (double)a.x * (double)c.x + (double)a.y * (double)c.y;   // conj(a) * c
      im += (double)a.x * (double)c.y - (double)a.y * (double)c.x;
    }
  }
  re = block_reduce_sum(re, red);
  im = block_reduce_sum(im, red);
  if (threadIdx.x == 0) { atomicAdd(&out[2 * b], re); atomicAdd(&out[2 * b + 1], im); }
}

template <typename c2>
__global__ void axpy_kernel(const double* __restrict__ a, const c2* __restrict__ x, c2* y, int n, long long batch) {
  const long long total = batch << n;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long b = i >> n;
    const double ar = a[2 * b], ai = a[2 * b + 1];
    const c2 xv = x[i]; c2 yv = y[i];
    yv.x = (double)yv.x + ar * (double)xv.x - ai * (double)xv.y;
    yv.y = (double)yv.y + ar * (double)xv.y + ai * (double)xv.x;
    y[i] = yv;
  }
}

template <typename c2>
__global__ void scal_kernel(const double* __restrict__ a, c2* x, int n, long long batch) {
  const long long total = batch << n;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long b = i >> n;
    const double ar = a[2 * b], ai = a[2 * b + 1];
    const c2 xv = x[i]; c2 o;
    o.x = ar * (double)xv.x - ai * (double)xv.y;
    o.y = ar * (double)xv.y + ai * (double)xv.x;
    x[i] = o;
  }
}

template <typename c2>
__global__ void lincomb_kernel(const double* __restrict__ c, const c2* __restrict__ V, int K, c2* out, int n,
                               long long batch) {
  const long long total = batch << n;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long b = i >> n;
    double re = 0.0, im = 0.0;
    for (int k = 0; k < K; ++k) {
      const double cr = c[((long long)k * batch + b) * 2], ci = c[((long long)k * batch + b) * 2 + 1];
      const c2 v = V[(long long)k * total + i];
      re += cr * (double)v.x - ci * (double)v.y;
      im += cr * (double)v.y + ci * (double)v.x;
    }
    c2 o; o.x = re; o.y = im;
    out[i] = o;
  }
}

// ---- fused Lanczos step (Krylov HamEvo): three passes over the state instead of nine BLAS-1 calls ---------------------
// acc[b] = {a, c_re, c_im, n2} accumulators of this step (zeroed by the caller once for all steps)
//   lanczos_a:    w <- w - beta_prev[b] * v_prev (if any);  a  += Re <v, w>
//   lanczos_orth: w <- w - a * v;                            c  += <v, w>,  n2 += ||w||^2
//   lanczos_next: v_next = (w - c v) / b,  b = sqrt(n2 - |c|^2)  (v is normalised);  alpha = a + Re c,  beta = b
__device__ __forceinline__ void block_reduce_add(double v, double* sh, double* dst) {
  const double t = block_reduce_sum(v, sh);
  if (threadIdx.x == 0) atomicAdd(dst, t);
}

template <typename c2>
__global__ void __launch_bounds__(256)
lanczos_a_kernel(c2* __restrict__ w, const c2* __restrict__ v, const c2* __restrict__ vprev, const double* __restrict__ beta_prev,
                 double* __restrict__ acc, int n, long long blocks_per_state) {
  __shared__ double red[32];
  const long long b = blockIdx.x / blocks_per_state, blk = blockIdx.x % blocks_per_state;
  const long long N = 1ll << n, base = (b << n) + blk * 2048ll;
  const double bp = (vprev != nullptr) ? beta_prev[b] : 0.0;
  double a = 0.0;
  c2 wv[8], vv[8], pv[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const long long i = blk * 2048ll + k * 256 + threadIdx.x;
    wv[k].x = 0; wv[k].y = 0; vv[k] = wv[k]; pv[k] = wv[k];
    if (i < N) { wv[k] = w[base + k * 256 + threadIdx.x]; vv[k] = v[base + k * 256 + threadIdx.x]; if (vprev != nullptr) pv[k] = vprev[base + k * 256 + threadIdx.x]; }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const long long i = blk * 2048ll + k * 256 + threadIdx.x;
    if (i >= N) break;
    const double wr = (double)wv[k].x - bp * (double)pv[k].x, wi = (double)wv[k].y - bp * (double)pv[k].y;
    a += (double)vv[k].x * wr + (double)vv[k].y * wi;
    if (vprev != nullptr) { c2 o; o.x = wr; o.y = wi; w[base + k * 256 + threadIdx.x] = o; }
  }
  block_reduce_add(a, red, acc + 4 * b);
}

template <typename c2>
__global__ void __launch_bounds__(256)
lanczos_orth_kernel(c2* __restrict__ w, const c2* __restrict__ v, double* __restrict__ acc, int n, long long blocks_per_state) {
  __shared__ double red[32];
  const long long b = blockIdx.x / blocks_per_state, blk = blockIdx.x % blocks_per_state;
  const long long N = 1ll << n, base = (b << n) + blk * 2048ll;
  const double a = acc[4 * b];
  double cr = 0.0, ci = 0.0, n2 = 0.0;
  c2 wv[8], vv[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const long long i = blk * 2048ll + k * 256 + threadIdx.x;
    wv[k].x = 0; wv[k].y = 0; vv[k] = wv[k];
    if (i < N) { wv[k] = w[base + k * 256 + threadIdx.x]; vv[k] = v[base + k * 256 + threadIdx.x]; }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const long long i = blk * 2048ll + k * 256 + threadIdx.x;
    if (i >= N) break;
    const double vr = vv[k].x, vi = vv[k].y;
    const double wr = (double)wv[k].x - a * vr, wi = (double)wv[k].y - a * vi;
    cr += vr * wr + vi * wi;  // conj(v) * w
    ci += vr * wi - vi * wr;
    n2 += wr * wr + wi * wi;
    c2 o; o.x = wr; o.y = wi;
    w[base + k * 256 + threadIdx.x] = o;
  }
  block_reduce_add(cr, red, acc + 4 * b + 1);
  block_reduce_add(ci, red, acc + 4 * b + 2);
  block_reduce_add(n2, red, acc + 4 * b + 3);
}

template <typename c2>
__global__ void __launch_bounds__(256)
lanczos_next_kernel(c2* __restrict__ w, const c2* __restrict__ v, const double* __restrict__ acc, double* __restrict__ alpha,
                    double* __restrict__ beta, int n, long long blocks_per_state) {
  const long long b = blockIdx.x / blocks_per_state, blk = blockIdx.x % blocks_per_state;
  const long long N = 1ll << n, base = (b << n) + blk * 2048ll;
  const double a = acc[4 * b], cr = acc[4 * b + 1], ci = acc[4 * b + 2], n2 = acc[4 * b + 3];
  const double al = a + cr;
  double b2 = n2 - (cr * cr + ci * ci);
  if (b2 < 0.0) b2 = 0.0;
  double bb = sqrt(b2);
  const double scale = fabs(al) > 1.0 ? fabs(al) : 1.0;
  const bool tiny = bb <= 1e-14 * scale;  // invariant subspace reached: the basis ends here
  const double inv = tiny ? 0.0 : 1.0 / bb;
  if (tiny) bb = 0.0;
  if (blk == 0 && threadIdx.x == 0) { alpha[b] = al; beta[b] = bb; }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const long long i = blk * 2048ll + k * 256 + threadIdx.x;
    if (i >= N) break;
    const c2 wv = w[base + k * 256 + threadIdx.x], vv = v[base + k * 256 + threadIdx.x];
    const double vr = vv.x, vi = vv.y;
    c2 o;
    o.x = ((double)wv.x - (cr * vr - ci * vi)) * inv;
    o.y = ((double)wv.y - (cr * vi + ci * vr)) * inv;
    w[base + k * 256 + threadIdx.x] = o;
  }
}

// ---- batched tridiagonal exponential (Lanczos small problem) -----------------------------------
constexpr int TRI_MAX = 64;
__global__ void tridiag_expm_kernel(const double* __restrict__ alpha, const double* __restrict__ beta,
                                    const double* __restrict__ tt, int m, long long batch, double* __restrict__ y,
                                    double* __restrict__ err) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  double a[TRI_MAX], c[TRI_MAX];
  double yr[TRI_MAX], yi[TRI_MAX], tr[TRI_MAX], ti[TRI_MAX], ar[TRI_MAX], ai[TRI_MAX];
  double nb = 0.0;
  for (int j = 0; j < m; ++j) { a[j] = alpha[(long long)j * batch + b]; c[j] = beta[(long long)j * batch + b]; }
  for (int j = 0; j < m; ++j) {
    const double r = fabs(a[j]) + (j > 0 ? fabs(c[j - 1]) : 0.0) + (j + 1 < m ? fabs(c[j]) : 0.0);
    nb = fmax(nb, r);
  }
  const double tau = tt[b];
  int steps = (int)ceil(nb * fabs(tau));
  if (steps < 1) steps = 1;
  const double h = tau / steps;
  for (int j = 0; j < m; ++j) { yr[j] = j == 0 ? 1.0 : 0.0; yi[j] = 0.0; }
  for (int s = 0; s < steps; ++s) {
    for (int j = 0; j < m; ++j) { tr[j] = yr[j]; ti[j] = yi[j]; ar[j] = yr[j]; ai[j] = yi[j]; }
    for (int k = 1; k <= 22; ++k) {  // term <- (-i h / k) T term ; ||h T|| <= 1 so 22 terms reach 1e-21
      const double f = h / k;
      double pr = 0.0, pi = 0.0;  // previous row's OLD value
      for (int j = 0; j < m; ++j) {
        const double cr = tr[j], ci = ti[j];
        double wr = a[j] * cr, wi = a[j] * ci;
        if (j > 0) { wr += c[j - 1] * pr; wi += c[j - 1] * pi; }
        if (j + 1 < m) { wr += c[j] * tr[j + 1]; wi += c[j] * ti[j + 1]; }
        pr = cr; pi = ci;
        // (-i f) (wr + i wi) = f wi - i f wr
        tr[j] = f * wi; ti[j] = -f * wr;
        ar[j] += tr[j]; ai[j] += ti[j];
      }
    }
    for (int j = 0; j < m; ++j) { yr[j] = ar[j]; yi[j] = ai[j]; }
  }
  for (int j = 0; j < m; ++j) { y[((long long)j * batch + b) * 2] = yr[j]; y[((long long)j * batch + b) * 2 + 1] = yi[j]; }
  err[b] = fabs(c[m - 1]) * sqrt(yr[m - 1] * yr[m - 1] + yi[m - 1] * yi[m - 1]);
}

// ---- dense k-qubit constant matrix, out of place ---------------------------------------------
struct DenseArgs { int bits[10]; };

template <typename c2>
__global__ void dense_kernel(const c2* __restrict__ psi, c2* __restrict__ out, int n, long long batch,
                             const double* __restrict__ mat, DenseArgs da, int k) {
  const long long total = batch << n;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int dim = 1 << k;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long b = i >> n;
    const long long idx = i & ((1ll << n) - 1);
    int row = 0;
    long long cleared = idx;
    for (int q = 0; q < k; ++q) {  // operand bit (k-1-q) <-> targets[q]
      const int bit = da.bits[q];
      row |= (int)((idx >> bit) & 1) << (k - 1 - q);
      cleared &= ~(1ll << bit);
    }
    double re = 0.0, im = 0.0;
    for (int col = 0; col < dim; ++col) {
      long long src = cleared;
      for (int q = 0; q < k; ++q) src |= (long long)((col >> (k - 1 - q)) & 1) << da.bits[q];
      const c2 v = psi[(b << n) + src];
      const double mr = mat[2 * ((long long)row * dim + col)], mi = mat[2 * ((long long)row * dim + col) + 1];
      re += mr * (double)v.x - mi * (double)v.y;
      im += mr * (double)v.y + mi * (double)v.x;
    }
    c2 o; o.x = re; o.y = im;
    out[i] = o;
  }
}

// ---- sampling ----------------------------------------------------------------------------------
template <typename c2>
__global__ void probs_kernel(const c2* __restrict__ psi, double* __restrict__ probs, long long total) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const double re = psi[i].x, im = psi[i].y;
    probs[i] = __dadd_rn(__dmul_rn(re, re), __dmul_rn(im, im));  // no FMA contraction: bit-reproducible
  }
}

__global__ void leaf_kernel(const double* __restrict__ probs, double* __restrict__ tree, int n, long long batch,
                            int chunk, long long L) {
  const long long total = batch * L;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += stride) {
    const long long b = j / L, leaf = j % L;
    const double* p = probs + (b << n) + leaf * chunk;
    double run = 0.0;
    for (int k = 0; k < chunk; ++k) run = __dadd_rn(run, p[k]);
    tree[b * 2 * L + leaf] = run;
  }
}

// one CTA per batch element builds the inner levels: level l at offset off_l, off_0 = 0, size L>>l
__global__ void tree_kernel(double* __restrict__ tree, long long L) {
  double* t = tree + (long long)blockIdx.x * 2 * L;
  long long off = 0, sz = L;
  while (sz > 1) {
    const long long nxt = off + sz;
    for (long long i = threadIdx.x; i < sz / 2; i += blockDim.x)
      t[nxt + i] = __dadd_rn(t[off + 2 * i], t[off + 2 * i + 1]);
    __syncthreads();
    off = nxt; sz >>= 1;
  }
}

__global__ void descend_kernel(const double* __restrict__ probs, const double* __restrict__ tree,
                               const double* __restrict__ uniforms, long long* __restrict__ out, int n,
                               long long batch, long long shots, int chunk, long long L, int levels) {
  const long long total = batch * shots;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += stride) {
    const long long b = j / shots;
    const double* t = tree + b * 2 * L;
    // offsets of levels: off_l = sum_{q<l} L>>q = 2L - (L >> (l-1)) for l>=1
    long long off_root = 2 * L - 2;  // level `levels-1` has size 1 (for L == 1: off 0)
    if (L == 1) off_root = 0;
    double v = __dmul_rn(uniforms[j], t[off_root]);
    long long node = 0;
    for (int lvl = levels - 2; lvl >= 0; --lvl) {
      const long long off = lvl == 0 ? 0 : 2 * L - (L >> (lvl - 1));
      const double left = t[off + 2 * node];
      if (v < left) node = 2 * node;
      else { v = __dadd_rn(v, -left); node = 2 * node + 1; }
    }
    const double* p = probs + (b << n) + node * chunk;
    double acc = 0.0;
    int sel = chunk - 1;
    for (int k = 0; k < chunk; ++k) {
      acc = __dadd_rn(acc, p[k]);
      if (v < acc) { sel = k; break; }
    }
    out[j] = node * chunk + sel;
  }
}

// ---- index permutations ------------------------------------------------------------------------
struct PermArgs { int perm[64]; };

template <typename c2>
__global__ void permute_kernel(const c2* __restrict__ in, c2* __restrict__ out, int n, long long batch, PermArgs pa) {
  const long long total = batch << n;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += stride) {
    // gather form: out[o] = in[P^{-1}(o)], with source bit j found at output position perm[j]
    const long long b = o >> n;
    const long long oi = o & ((1ll << n) - 1);
    long long src = 0;
    for (int j = 0; j < n; ++j) src |= ((oi >> pa.perm[j]) & 1ll) << j;
    out[o] = in[(b << n) + src];
  }
}

inline int grid_for(long long total, int threads) {
  long long g = (total + threads - 1) / threads;
  const long long cap = 148ll * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int b200sv_abi_version(void) { return B200SV_ABI_VERSION; }

int b200sv_lean_supported(int dtype, int m, int r, int adjoint) {
  if (dtype != B200SV_C64 && dtype != B200SV_C128) return 0;
  return lean_supported(dtype, m, r, adjoint != 0) ? 1 : 0;
}

int b200sv_device_sm_count(void) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return B200SV_E_NODEVICE; }
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    cudaGetLastError();
    return B200SV_E_NODEVICE;
  }
  return sms;
}

int b200sv_init_zero_state(void* state, int dtype, int n, int64_t batch, void* stream) {
  if (!state || n < 0 || n > 40 || batch <= 0) return B200SV_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)batch << n;
  if (dtype == B200SV_C128) zero_state_kernel<double2><<<grid_for(total, 256), 256, 0, st>>>((double2*)state, n, batch);
  else if (dtype == B200SV_C64) zero_state_kernel<float2><<<grid_for(total, 256), 256, 0, st>>>((float2*)state, n, batch);
  else return B200SV_E_ARG;
  LAUNCH_CK();
  return 0;
}

int64_t b200sv_plan_workspace_bytes(const b200sv_plan* plan, int dtype, int64_t batch, int adjoint) {
  if (!plan || batch <= 0) return B200SV_E_ARG;
  if (dtype == B200SV_C128) return (int64_t)workspace_bytes<double>(plan, batch, adjoint != 0);
  if (dtype == B200SV_C64) return (int64_t)workspace_bytes<float>(plan, batch, adjoint != 0);
  return B200SV_E_ARG;
}

int b200sv_apply_plan(void* state, int dtype, int n, int64_t batch, const double* params,
                      const b200sv_plan* plan, int first, int last, uint64_t index_hi, void* workspace,
                      void* stream) {
  if (!state || !plan || !workspace || first < 0 || last > plan->n_sweeps || batch <= 0) return B200SV_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B200SV_C128) return run_plan<double, false>(state, nullptr, n, batch, params, nullptr, plan, first, last, index_hi, workspace, st);
  if (dtype == B200SV_C64) return run_plan<float, false>(state, nullptr, n, batch, params, nullptr, plan, first, last, index_hi, workspace, st);
  return B200SV_E_ARG;
}

int b200sv_adjoint_plan(void* psi, void* lam, int dtype, int n, int64_t batch, const double* params,
                        double* grads, const b200sv_plan* plan, int first, int last, uint64_t index_hi,
                        void* workspace, void* stream) {
  if (!psi || !lam || !plan || !grads || !workspace || first < 0 || last > plan->n_sweeps || batch <= 0) return B200SV_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B200SV_C128) return run_plan<double, true>(psi, lam, n, batch, params, grads, plan, first, last, index_hi, workspace, st);
  if (dtype == B200SV_C64) return run_plan<float, true>(psi, lam, n, batch, params, grads, plan, first, last, index_hi, workspace, st);
  return B200SV_E_ARG;
}

static int check_terms(int n_terms, int n_single = 0) {
  if (n_single < 0 || n_single > n_terms) return B200SV_E_ARG;
  return (n_terms < 0 || n_terms - n_single > MAX_TERMS_SMEM) ? B200SV_E_UNSUPPORTED : 0;
}

int b200sv_pauli_expectation(const void* psi, int dtype, int n, int64_t batch, const b200sv_pterm* terms,
                             int n_terms, int n_single, const double* coef, int64_t coef_ld, double* out,
                             uint64_t index_hi, void* stream) {
  if (!psi || !out || batch <= 0 || (coef_ld != 1 && coef_ld != batch)) return B200SV_E_ARG;
  if (int rc = check_terms(n_terms, n_single)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const long long N = 1ll << n;
  const long long per = PK_TILE;
  const long long bps = (N + per - 1) / per;
  const long long grid = bps * batch;
  if (grid > 2147483647ll) return B200SV_E_ARG;
  if (dtype == B200SV_C128)
    pauli_expect_kernel<double2><<<(unsigned)grid, EXP_THREADS, 0, st>>>((const double2*)psi, n, batch, terms, n_terms, n_single, coef, coef_ld, out, index_hi, bps, PeerTable{});
  else if (dtype == B200SV_C64)
    pauli_expect_kernel<float2><<<(unsigned)grid, EXP_THREADS, 0, st>>>((const float2*)psi, n, batch, terms, n_terms, n_single, coef, coef_ld, out, index_hi, bps, PeerTable{});
  else return B200SV_E_ARG;
  LAUNCH_CK();
  return 0;
}

int b200sv_pauli_apply(const void* psi, void* out, const void* acc, int dtype, int n, int64_t batch,
                       const b200sv_pterm* terms, int n_terms, int n_single, const double* coef, int64_t coef_ld,
                       double ar, double ai, double br, double bi, uint64_t index_hi, void* stream) {
  if (!psi || !out || psi == out || batch <= 0 || (coef_ld != 1 && coef_ld != batch)) return B200SV_E_ARG;
  if (int rc = check_terms(n_terms, n_single)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const long long N = 1ll << n;
  const long long bps = (N + PK_TILE - 1) / PK_TILE;
  const long long grid = bps * batch;
  if (grid > 2147483647ll) return B200SV_E_ARG;
  if (br == 0.0 && bi == 0.0) acc = nullptr;
  if (dtype == B200SV_C128)
    pauli_apply_kernel<double2><<<(unsigned)grid, 256, 0, st>>>((const double2*)psi, (double2*)out, (const double2*)acc, n, batch, terms, n_terms, n_single, coef, coef_ld, ar, ai, br, bi, index_hi, bps, PeerTable{});
  else if (dtype == B200SV_C64)
    pauli_apply_kernel<float2><<<(unsigned)grid, 256, 0, st>>>((const float2*)psi, (float2*)out, (const float2*)acc, n, batch, terms, n_terms, n_single, coef, coef_ld, ar, ai, br, bi, index_hi, bps, PeerTable{});
  else return B200SV_E_ARG;
  LAUNCH_CK();
  return 0;
}

int b200sv_pauli_diagonal(const b200sv_pterm* terms, int n_terms, int n_single, const double* coef, int64_t coef_ld, int n,
                          int64_t rows, double* out, uint64_t index_hi, void* stream) {
  if (!out || !coef || n < 0 || n > 40 || rows <= 0 || (coef_ld != 1 && coef_ld != rows)) return B200SV_E_ARG;
  if (int rc = check_terms(n_terms, n_single)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const long long N = 1ll << n;
  const long long bps = (N + 1023) / 1024;
  const long long grid = bps * rows;
  if (grid > 2147483647ll) return B200SV_E_ARG;
  pauli_diagonal_kernel<<<(unsigned)grid, 256, 0, st>>>(terms, n_terms, n_single, coef, coef_ld, n, (double2*)out, index_hi, bps);
  LAUNCH_CK();
  return 0;
}

int b200sv_pauli_expectation_diag(const void* psi, int dtype, int n, int64_t batch, const b200sv_pterm* terms, int n_terms,
                                  const double* coef, int64_t coef_ld, const double* diag, int64_t diag_ld, double* out, void* stream) {
  if (!psi || !out || !diag || batch <= 0 || (coef_ld != 1 && coef_ld != batch) || (diag_ld != 1 && diag_ld != batch)) return B200SV_E_ARG;
  if (n_terms > 0 && (!terms || !coef)) return B200SV_E_ARG;
  if (int rc = check_terms(n_terms, 0)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const long long N = 1ll << n;
  const long long bps = (N + PK_TILE - 1) / PK_TILE;
  const long long grid = bps * batch;
  if (grid > 2147483647ll) return B200SV_E_ARG;
  if (dtype == B200SV_C128)
    pauli_expect_kernel<double2, false, true><<<(unsigned)grid, EXP_THREADS, 0, st>>>((const double2*)psi, n, batch, terms, n_terms, 0, coef, coef_ld, out, 0ull, bps, PeerTable{}, diag, (long long)diag_ld);
  else if (dtype == B200SV_C64)
    pauli_expect_kernel<float2, false, true><<<(unsigned)grid, EXP_THREADS, 0, st>>>((const float2*)psi, n, batch, terms, n_terms, 0, coef, coef_ld, out, 0ull, bps, PeerTable{}, diag, (long long)diag_ld);
  else return B200SV_E_ARG;
  LAUNCH_CK();
  return 0;
}

int b200sv_pauli_apply_diag(const void* psi, void* out, const void* acc, int dtype, int n, int64_t batch,
                            const b200sv_pterm* terms, int n_terms, const double* coef, int64_t coef_ld, const double* diag,
                            int64_t diag_ld, const double* diag_scale, double ar, double ai, double br, double bi, void* stream) {
  if (!psi || !out || psi == out || !diag || batch <= 0 || (coef_ld != 1 && coef_ld != batch) || (diag_ld != 1 && diag_ld != batch)) return B200SV_E_ARG;
  if (n_terms > 0 && (!terms || !coef)) return B200SV_E_ARG;
  if (int rc = check_terms(n_terms, 0)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const long long N = 1ll << n;
  const long long bps = (N + PK_TILE - 1) / PK_TILE;
  const long long grid = bps * batch;
  if (grid > 2147483647ll) return B200SV_E_ARG;
  if (br == 0.0 && bi == 0.0) acc = nullptr;
  if (dtype == B200SV_C128)
    pauli_apply_kernel<double2, false, true><<<(unsigned)grid, 256, 0, st>>>((const double2*)psi, (double2*)out, (const double2*)acc, n, batch, terms, n_terms, 0, coef, coef_ld, ar, ai, br, bi, 0ull, bps, PeerTable{}, diag, (long long)diag_ld, diag_scale);
  else if (dtype == B200SV_C64)
    pauli_apply_kernel<float2, false, true><<<(unsigned)grid, 256, 0, st>>>((const float2*)psi, (float2*)out, (const float2*)acc, n, batch, terms, n_terms, 0, coef, coef_ld, ar, ai, br, bi, 0ull, bps, PeerTable{}, diag, (long long)diag_ld, diag_scale);
  else return B200SV_E_ARG;
  LAUNCH_CK();
  return 0;
}

int b200sv_diag_evolve(void* psi, int dtype, int n, int64_t batch, const b200sv_pterm* terms, int n_terms,
                       int n_single, const double* coef, int64_t coef_ld, const double* t, int64_t t_ld, double sign,
                       uint64_t index_hi, void* stream) {
  if (!psi || !t || batch <= 0 || (coef_ld != 1 && coef_ld != batch) || (t_ld != 1 && t_ld != batch)) return B200SV_E_ARG;
  if (int rc = check_terms(n_terms, n_single)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const long long N = 1ll << n;
  const long long bps = (N + 1023) / 1024;
  const long long grid = bps * batch;
  if (grid > 2147483647ll) return B200SV_E_ARG;
  if (dtype == B200SV_C128)
    diag_evolve_kernel<double2><<<(unsigned)grid, 256, 0, st>>>((double2*)psi, n, batch, terms, n_terms, n_single, coef, coef_ld, t, t_ld, sign, index_hi, bps);
  else if (dtype == B200SV_C64)
    diag_evolve_kernel<float2><<<(unsigned)grid, 256, 0, st>>>((float2*)psi, n, batch, terms, n_terms, n_single, coef, coef_ld, t, t_ld, sign, index_hi, bps);
  else return B200SV_E_ARG;
  LAUNCH_CK();
  return 0;
}

int b200sv_dot(const void* x, const void* y, int dtype, int n, int64_t batch, double* out, void* stream) {
  if (!x || !y || !out || batch <= 0) return B200SV_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaMemsetAsync(out, 0, sizeof(double) * 2 * batch, st));
  const long long N = 1ll << n;
  const long long bps = (N + 2047) / 2048;
  const long long grid = bps * batch;
  if (grid > 2147483647ll) return B200SV_E_ARG;
  if (dtype == B200SV_C128) dot_kernel<double2><<<(unsigned)grid, 256, 0, st>>>((const double2*)x, (const double2*)y, n, out, bps);
  else if (dtype == B200SV_C64) dot_kernel<float2><<<(unsigned)grid, 256, 0, st>>>((const float2*)x, (const float2*)y, n, out, bps);
  else return B200SV_E_ARG;
  LAUNCH_CK();
  return 0;
}

int b200sv_axpy(const double* a, const void* x, void* y, int dtype, int n, int64_t batch, void* stream) {
  if (!a || !x || !y || batch <= 0) return B200SV_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)batch << n;
  if (dtype == B200SV_C128) axpy_kernel<double2><<<grid_for(total, 256), 256, 0, st>>>(a, (const double2*)x, (double2*)y, n, batch);
  else if (dtype == B200SV_C64) axpy_kernel<float2><<<grid_for(total, 256), 256, 0, st>>>(a, (const float2*)x, (float2*)y, n, batch);
  else return B200SV_E_ARG;
  LAUNCH_CK();
  return 0;
}

int b200sv_scal(const double* a, void* x, int dtype, int n, int64_t batch, void* stream) {
  if (!a || !x || batch <= 0) return B200SV_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)batch << n;
  if (dtype == B200SV_C128) scal_kernel<double2><<<grid_for(total, 256), 256, 0, st>>>(a, (double2*)x, n, batch);
  else if (dtype == B200SV_C64) scal_kernel<float2><<<grid_for(total, 256), 256, 0, st>>>(a, (float2*)x, n, batch);
  else return B200SV_E_ARG;
  LAUNCH_CK();
  return 0;
}

int b200sv_lincomb(const double* c, const void* V, int K, void* out, int dtype, int n, int64_t batch, void* stream) {
  if (!c || !V || !out || K <= 0 || batch <= 0) return B200SV_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)batch << n;
  if (dtype == B200SV_C128) lincomb_kernel<double2><<<grid_for(total, 256), 256, 0, st>>>(c, (const double2*)V, K, (double2*)out, n, batch);
  else if (dtype == B200SV_C64) lincomb_kernel<float2><<<grid_for(total, 256), 256, 0, st>>>(c, (const float2*)V, K, (float2*)out, n, batch);
  else return B200SV_E_ARG;
  LAUNCH_CK();
  return 0;
}

int b200sv_lanczos_step(void* w, const void* v, const void* v_prev, const double* beta_prev, double* acc, double* alpha,
                        double* beta, int dtype, int n, int64_t batch, void* stream) {
  if (!w || !v || !acc || !alpha || !beta || (v_prev && !beta_prev) || n < 0 || n > 40 || batch <= 0 || w == v) return B200SV_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const long long N = 1ll << n;
  const long long bps = (N + 2047) / 2048;
  const long long grid = bps * batch;
  if (grid > 2147483647ll) return B200SV_E_ARG;
  if (dtype == B200SV_C128) {
    lanczos_a_kernel<double2><<<(unsigned)grid, 256, 0, st>>>((double2*)w, (const double2*)v, (const double2*)v_prev, beta_prev, acc, n, bps);
    lanczos_orth_kernel<double2><<<(unsigned)grid, 256, 0, st>>>((double2*)w, (const double2*)v, acc, n, bps);
    lanczos_next_kernel<double2><<<(unsigned)grid, 256, 0, st>>>((double2*)w, (const double2*)v, acc, alpha, beta, n, bps);
  } else if (dtype == B200SV_C64) {
    lanczos_a_kernel<float2><<<(unsigned)grid, 256, 0, st>>>((float2*)w, (const float2*)v, (const float2*)v_prev, beta_prev, acc, n, bps);
    lanczos_orth_kernel<float2><<<(unsigned)grid, 256, 0, st>>>((float2*)w, (const float2*)v, acc, n, bps);
    lanczos_next_kernel<float2><<<(unsigned)grid, 256, 0, st>>>((float2*)w, (const float2*)v, acc, alpha, beta, n, bps);
  } else return B200SV_E_ARG;
  LAUNCH_CK();
  return 0;
}

int b200sv_tridiag_expm(const double* alpha, const double* beta, const double* t, int m, int64_t batch,
                        double* y, double* err, void* stream) {
  if (!alpha || !beta || !t || !y || !err || m < 1 || m > TRI_MAX || batch <= 0) return B200SV_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  tridiag_expm_kernel<<<(unsigned)((batch + 63) / 64), 64, 0, st>>>(alpha, beta, t, m, (long long)batch, y, err);
  LAUNCH_CK();
  return 0;
}

int b200sv_apply_dense(const void* psi, void* out, int dtype, int n, int64_t batch, const double* matrix,
                       const int32_t* target_bits, int k, void* stream) {
  if (!psi || !out || psi == out || !matrix || !target_bits || k < 1 || k > 10 || k > n || batch <= 0) return B200SV_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  DenseArgs da;
  for (int q = 0; q < k; ++q) {
    if (target_bits[q] < 0 || target_bits[q] >= n) return B200SV_E_ARG;
    da.bits[q] = target_bits[q];
  }
  const long long total = (long long)batch << n;
  if (dtype == B200SV_C128) dense_kernel<double2><<<grid_for(total, 256), 256, 0, st>>>((const double2*)psi, (double2*)out, n, batch, matrix, da, k);
  else if (dtype == B200SV_C64) dense_kernel<float2><<<grid_for(total, 256), 256, 0, st>>>((const float2*)psi, (float2*)out, n, batch, matrix, da, k);
  else return B200SV_E_ARG;
  LAUNCH_CK();
  return 0;
}

int b200sv_sample(const void* psi, int dtype, int n, int64_t batch, double* probs, double* tree,
                  const double* uniforms, int64_t n_shots, int64_t* out_idx, void* stream) {
  if (!psi || !probs || !tree || !uniforms || !out_idx || batch <= 0 || n_shots <= 0 || n < 0 || n > 34) return B200SV_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const long long N = 1ll << n;
  const int chunk = (int)(N < 256 ? N : 256);
  const long long L = N / chunk;
  int levels = 1;
  for (long long s = L; s > 1; s >>= 1) ++levels;
  const long long total = (long long)batch << n;
  if (dtype == B200SV_C128) probs_kernel<double2><<<grid_for(total, 256), 256, 0, st>>>((const double2*)psi, probs, total);
  else if (dtype == B200SV_C64) probs_kernel<float2><<<grid_for(total, 256), 256, 0, st>>>((const float2*)psi, probs, total);
  else return B200SV_E_ARG;
  LAUNCH_CK();
  leaf_kernel<<<grid_for(batch * L, 128), 128, 0, st>>>(probs, tree, n, batch, chunk, L);
  LAUNCH_CK();
  if (batch > 2147483647ll) return B200SV_E_ARG;
  tree_kernel<<<(unsigned)batch, 256, 0, st>>>(tree, L);
  LAUNCH_CK();
  descend_kernel<<<grid_for(batch * n_shots, 128), 128, 0, st>>>(probs, tree, uniforms, (long long*)out_idx, n, batch, n_shots, chunk, L, levels);
  LAUNCH_CK();
  return 0;
}

int b200sv_permute_bits(const void* in, void* out, int dtype, int n, int64_t batch, const int32_t* perm, void* stream) {
  if (!in || !out || in == out || !perm || n < 0 || n > 40 || batch <= 0) return B200SV_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  PermArgs pa;
  uint64_t seen = 0;
  for (int j = 0; j < n; ++j) {
    if (perm[j] < 0 || perm[j] >= n || ((seen >> perm[j]) & 1)) return B200SV_E_ARG;
    seen |= 1ull << perm[j];
    pa.perm[j] = perm[j];
  }
  const long long total = (long long)batch << n;
  if (dtype == B200SV_C128) permute_kernel<double2><<<grid_for(total, 256), 256, 0, st>>>((const double2*)in, (double2*)out, n, batch, pa);
  else if (dtype == B200SV_C64) permute_kernel<float2><<<grid_for(total, 256), 256, 0, st>>>((const float2*)in, (float2*)out, n, batch, pa);
  else return B200SV_E_ARG;
  LAUNCH_CK();
  return 0;
}

int b200sv_bit_reverse(const void* in, void* out, int dtype, int n, int64_t batch, void* stream) {
  if (n < 0 || n > 40) return B200SV_E_ARG;
  int32_t perm[64];
  for (int j = 0; j < n; ++j) perm[j] = n - 1 - j;
  return b200sv_permute_bits(in, out, dtype, n, batch, perm, stream);
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------------
// Device-side parameter embedding: one thread per (slot, batch element), slot uniform per CTA row
// ------------------------------------------------------------------------------------------------------------------
template <bool BWD>
__global__ void __launch_bounds__(128) embed_kernel(const b200sv_enode* __restrict__ nodes, const int32_t* __restrict__ slot_off,
                                                    const double* __restrict__ consts, const double* __restrict__ var,
                                                    const double* __restrict__ feat, long long batch, const double* __restrict__ gout,
                                                    double* __restrict__ out) {
  const int slot = blockIdx.y;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= batch) return;
  const int first = slot_off[slot], n = slot_off[slot + 1] - first;
  const b200sv_enode* nd = nodes + first;
  double val[B200SV_EMBED_MAX_NODES];
  for (int i = 0; i < n; ++i) {
    const b200sv_enode q = nd[i];
    const double x = (q.op >= B200SV_EO_ADD) ? val[q.a] : 0.0;
    double r;
    switch (q.op) {
      case B200SV_EO_CONST: r = consts[q.arg]; break;
      case B200SV_EO_VAR: r = var[q.arg]; break;
      case B200SV_EO_FEAT: r = feat[(long long)q.arg * batch + e]; break;
      case B200SV_EO_ADD: r = x + val[q.b]; break;
      case B200SV_EO_MUL: r = x * val[q.b]; break;
      case B200SV_EO_POW: r = pow(x, val[q.b]); break;
      case B200SV_EO_SIN: r = sin(x); break;
      case B200SV_EO_COS: r = cos(x); break;
      case B200SV_EO_TAN: r = tan(x); break;
      case B200SV_EO_TANH: r = tanh(x); break;
      case B200SV_EO_LOG: r = log(x); break;
      case B200SV_EO_EXP: r = exp(x); break;
      case B200SV_EO_ACOS: r = acos(x); break;
      case B200SV_EO_ASIN: r = asin(x); break;
      case B200SV_EO_ATAN: r = atan(x); break;
      case B200SV_EO_ABS: r = fabs(x); break;
      default: r = x > 0.0 ? 1.0 : (x == 0.0 ? 0.5 : 0.0); break;  // HEAVISIDE (torch.heaviside: NaN -> 0)
    }
    val[i] = r;
  }
  if constexpr (!BWD) {
    out[(long long)slot * batch + e] = n > 0 ? val[n - 1] : 0.0;
  } else {
  double adj[B200SV_EMBED_MAX_NODES];
  for (int i = 0; i < n; ++i) adj[i] = 0.0;
  if (n > 0) adj[n - 1] = gout[(long long)slot * batch + e];
  for (int i = n - 1; i >= 0; --i) {
    const b200sv_enode q = nd[i];
    const double g = adj[i];
    if (q.op == B200SV_EO_VAR || q.op == B200SV_EO_FEAT) {
      out[(long long)q.load * batch + e] = g;
      continue;
    }
    if (q.op < B200SV_EO_ADD) continue;
    const double x = val[q.a];
    switch (q.op) {
      case B200SV_EO_ADD: adj[q.a] += g; adj[q.b] += g; break;
      case B200SV_EO_MUL: adj[q.a] += g * val[q.b]; adj[q.b] += g * x; break;
      case B200SV_EO_POW: {
        const double y = val[q.b];
        adj[q.a] += g * y * pow(x, y - 1.0);
        if (nd[q.b].op != B200SV_EO_CONST) adj[q.b] += g * val[i] * log(x);
        break;
      }
      case B200SV_EO_SIN: adj[q.a] += g * cos(x); break;
      case B200SV_EO_COS: adj[q.a] -= g * sin(x); break;
      case B200SV_EO_TAN: adj[q.a] += g * (1.0 + val[i] * val[i]); break;
      case B200SV_EO_TANH: adj[q.a] += g * (1.0 - val[i] * val[i]); break;
      case B200SV_EO_LOG: adj[q.a] += g / x; break;
      case B200SV_EO_EXP: adj[q.a] += g * val[i]; break;
      case B200SV_EO_ACOS: adj[q.a] -= g / sqrt(1.0 - x * x); break;
      case B200SV_EO_ASIN: adj[q.a] += g / sqrt(1.0 - x * x); break;
      case B200SV_EO_ATAN: adj[q.a] += g / (1.0 + x * x); break;
      case B200SV_EO_ABS: adj[q.a] += x > 0.0 ? g : (x < 0.0 ? -g : 0.0); break;
      default: break;  // HEAVISIDE: zero gradient
    }
  }
  }
}

static int embed_launch(bool bwd, const b200sv_enode* nodes, const int32_t* slot_off, int n_slots, const double* consts, const double* var,
                        const double* feat, int64_t batch, const double* gout, double* out, void* stream) {
  if (!nodes || !slot_off || n_slots < 0 || n_slots > 65535 || batch <= 0 || !out || (bwd && !gout)) return B200SV_E_ARG;
  if (n_slots == 0) return 0;
  dim3 grid((unsigned)((batch + 127) / 128), (unsigned)n_slots);
  cudaStream_t st = (cudaStream_t)stream;
  if (bwd) embed_kernel<true><<<grid, 128, 0, st>>>(nodes, slot_off, consts, var, feat, batch, gout, out);
  else embed_kernel<false><<<grid, 128, 0, st>>>(nodes, slot_off, consts, var, feat, batch, nullptr, out);
  LAUNCH_CK();
  return 0;
}

extern "C" {
int b200sv_apply_plan_peers(void* const* peer_states, int n_peers, int rank, int dtype, int n_total, const double* params,
                            const b200sv_plan* plan, int first_sweep, int last_sweep, void* workspace, void* stream) {
  if (!peer_states || !plan || n_total < 1 || n_total > 48 || first_sweep < 0 || last_sweep > plan->n_sweeps) return B200SV_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (!workspace) return B200SV_E_ARG;
  if (dtype == B200SV_C128) return run_plan_peers<double, false>(peer_states, nullptr, n_peers, rank, n_total, params, nullptr, plan, first_sweep, last_sweep, workspace, st);
  if (dtype == B200SV_C64) return run_plan_peers<float, false>(peer_states, nullptr, n_peers, rank, n_total, params, nullptr, plan, first_sweep, last_sweep, workspace, st);
  return B200SV_E_ARG;
}

int b200sv_adjoint_plan_peers(void* const* peer_states, void* const* peer_lams, int n_peers, int rank, int dtype, int n_total,
                              const double* params, double* grads, const b200sv_plan* plan, int first_sweep, int last_sweep,
                              void* workspace, void* stream) {
  if (!peer_states || !peer_lams || !plan || !grads || !workspace || n_total < 1 || n_total > 48 || first_sweep < 0 ||
      last_sweep > plan->n_sweeps)
    return B200SV_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B200SV_C128) return run_plan_peers<double, true>(peer_states, peer_lams, n_peers, rank, n_total, params, grads, plan, first_sweep, last_sweep, workspace, st);
  if (dtype == B200SV_C64) return run_plan_peers<float, true>(peer_states, peer_lams, n_peers, rank, n_total, params, grads, plan, first_sweep, last_sweep, workspace, st);
  return B200SV_E_ARG;
}

static int peer_table(void* const* peer_states, int n_peers, int rank, int n_total, PeerTable* pt) {
  int g = 0;
  while ((1 << g) < n_peers) ++g;
  if (!peer_states || (1 << g) != n_peers || n_peers > 16 || rank < 0 || rank >= n_peers || n_total <= g || n_total > 48) return B200SV_E_ARG;
  for (int i = 0; i < n_peers; ++i) {
    if (!peer_states[i]) return B200SV_E_ARG;
    pt->p[i] = peer_states[i];
  }
  pt->n_local = n_total - g;
  return 0;
}

int b200sv_pauli_expectation_peers(void* const* peer_states, int n_peers, int rank, int dtype, int n_total,
                                   const b200sv_pterm* terms, int n_terms, int n_single, const double* coef, double* out,
                                   void* stream) {
  PeerTable pt{};
  if (int rc = peer_table(peer_states, n_peers, rank, n_total, &pt)) return rc;
  if (!out || !coef) return B200SV_E_ARG;
  if (int rc = check_terms(n_terms, n_single)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int n = pt.n_local;
  const long long N = 1ll << n;
  const long long per = PK_TILE;
  const long long bps = (N + per - 1) / per;
  if (bps > 2147483647ll) return B200SV_E_ARG;
  const unsigned long long index_hi = (unsigned long long)rank << n;
  if (dtype == B200SV_C128)
    pauli_expect_kernel<double2, true><<<(unsigned)bps, EXP_THREADS, 0, st>>>((const double2*)peer_states[rank], n, 1, terms, n_terms, n_single, coef, 1, out, index_hi, bps, pt);
  else if (dtype == B200SV_C64)
    pauli_expect_kernel<float2, true><<<(unsigned)bps, EXP_THREADS, 0, st>>>((const float2*)peer_states[rank], n, 1, terms, n_terms, n_single, coef, 1, out, index_hi, bps, pt);
  else return B200SV_E_ARG;
  LAUNCH_CK();
  return 0;
}

int b200sv_pauli_apply_peers(void* const* peer_states, int n_peers, int rank, void* out_local, int dtype, int n_total,
                             const b200sv_pterm* terms, int n_terms, int n_single, const double* coef, double alpha_re,
                             double alpha_im, void* stream) {
  PeerTable pt{};
  if (int rc = peer_table(peer_states, n_peers, rank, n_total, &pt)) return rc;
  if (!out_local || !coef) return B200SV_E_ARG;
  for (int i = 0; i < n_peers; ++i)
    if (peer_states[i] == out_local) return B200SV_E_ARG;  // out of place
  if (int rc = check_terms(n_terms, n_single)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int n = pt.n_local;
  const long long N = 1ll << n;
  const long long bps = (N + PK_TILE - 1) / PK_TILE;
  if (bps > 2147483647ll) return B200SV_E_ARG;
  const unsigned long long index_hi = (unsigned long long)rank << n;
  if (dtype == B200SV_C128)
    pauli_apply_kernel<double2, true><<<(unsigned)bps, 256, 0, st>>>((const double2*)peer_states[rank], (double2*)out_local, nullptr, n, 1, terms, n_terms, n_single, coef, 1, alpha_re, alpha_im, 0.0, 0.0, index_hi, bps, pt);
  else if (dtype == B200SV_C64)
    pauli_apply_kernel<float2, true><<<(unsigned)bps, 256, 0, st>>>((const float2*)peer_states[rank], (float2*)out_local, nullptr, n, 1, terms, n_terms, n_single, coef, 1, alpha_re, alpha_im, 0.0, 0.0, index_hi, bps, pt);
  else return B200SV_E_ARG;
  LAUNCH_CK();
  return 0;
}

int b200sv_ipc_alloc(int64_t bytes, void** ptr, void* handle64) {
  if (bytes <= 0 || !ptr || !handle64) return B200SV_E_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  CK(cudaMalloc(ptr, (size_t)bytes));
  const cudaError_t e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), *ptr);
  if (e != cudaSuccess) {  // do not leak the allocation when the handle cannot be made
    cudaFree(*ptr);
    *ptr = nullptr;
    return (int)e;
  }
  return 0;
}
int b200sv_ipc_open(const void* handle64, void** ptr) {
  if (!handle64 || !ptr) return B200SV_E_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  CK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
int b200sv_ipc_close(void* ptr) {
  if (!ptr) return B200SV_E_ARG;
  CK(cudaIpcCloseMemHandle(ptr));
  return 0;
}
int b200sv_ipc_free(void* ptr) {
  if (!ptr) return B200SV_E_ARG;
  CK(cudaFree(ptr));
  return 0;
}

int b200sv_embed_forward(const b200sv_enode* nodes, const int32_t* slot_off, int n_slots, const double* consts, const double* var,
                         const double* feat, int64_t batch, double* out, void* stream) {
  return embed_launch(false, nodes, slot_off, n_slots, consts, var, feat, batch, nullptr, out, stream);
}
int b200sv_embed_backward(const b200sv_enode* nodes, const int32_t* slot_off, int n_slots, const double* consts, const double* var,
                          const double* feat, int64_t batch, const double* gout, double* gloads, void* stream) {
  return embed_launch(true, nodes, slot_off, n_slots, consts, var, feat, batch, gout, gloads, stream);
}
}  // extern "C"
