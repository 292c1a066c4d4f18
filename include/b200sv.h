/*
 * b200sv — C-ABI of the B200-native state-vector backend for qadence.
 *
 * The reference (pasqal-io/qadence v1.11.5) is 100 % Python and has no FFI: its hot path calls the
 * third-party `pyqtorch` package, which executes one `torch.einsum` per gate.  Every entry point
 * below therefore cites the *reference call site it replaces* (paths relative to /root/reference).
 * A maintainer binds these with `ctypes`; no torch types cross this boundary:
 * plain device pointers, sizes, a dtype enum and a `cudaStream_t` passed as `void*`.
 *
 * Conventions
 *   - states are `[batch][2^n]` interleaved complex (re, im), float32 or float64 per component,
 *     batch-major, basis index bit (n-1-q) <-> qubit q (qubit 0 = most significant bit;
 *     tests/backends/test_endianness.py:113-138);
 *   - "bit position" always means the position in the basis index (0 = least significant);
 *   - parameter tables are `double[n_slots][batch]` on the device;
 *   - every function returns 0 on success, a positive `cudaError_t`, or a negative B200SV_E_* code;
 *     nothing throws, nothing allocates device memory (callers pass workspaces), all work is
 *     enqueued on the given stream and the call returns without synchronising.
 */
#ifndef B200SV_H
#define B200SV_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200SV_ABI_VERSION 2

enum { B200SV_C64 = 0, B200SV_C128 = 1 };

enum {
  B200SV_E_ARG = -1,      /* invalid argument */
  B200SV_E_UNSUPPORTED = -2,
  B200SV_E_NODEVICE = -3,
};

/* ---- fused sweep plan records (built by qadence_b200/plan.py, packed little-endian) ---------- */

enum {
  B200SV_GK_MAT = 0,   /* general complex 2x2 from the constant pool (4 complex128 at cidx)        */
  B200SV_GK_REAL = 1,  /* real 2x2 from the constant pool (H, ...): 4 doubles packed in 2 complex  */
  B200SV_GK_RX = 2,    /* exp(-i θ X/2), θ = params[slot][b] or constant at cidx                   */
  B200SV_GK_RY = 3,
  B200SV_GK_RZ = 4,
  B200SV_GK_PHASE = 5, /* diag(1, e^{iθ})                                                           */
  B200SV_GK_DIAG = 6,  /* diag(d0, d1) from the constant pool (Z, S, T, N, ...)                    */
  B200SV_GK_X = 7,     /* Pauli X as a register permutation (CNOT / Toffoli with controls)         */
  B200SV_GK_SWAP = 8,  /* swap of two register bits (SWAP / CSWAP)                                  */
  B200SV_GK_SCALE = 9  /* ψ <- s ψ, s = params[slot][b] or complex constant at cidx                */
};

#define B200SV_GF_GRAD 1 /* gate.flags: parametric gate whose gradient is requested (adjoint)      */

/* execution class of a resolved entry (chosen by the planner; a chain of 1-qubit gates on one
 * target is multiplied into a single 2x2 per CTA and executed as the class of the product) */
enum {
  B200SV_XK_MAT = 0,  /* general complex 2x2                      */
  B200SV_XK_REAL = 1, /* real 2x2 (RY, H, chains of those)        */
  B200SV_XK_RX = 2,   /* [[c, -i s], [-i s, c]] (RX chains)       */
  B200SV_XK_DIAG = 3, /* diagonal (RZ, PHASE, Z, S, T, N, chains) */
  B200SV_XK_X = 4,
  B200SV_XK_SWAP = 5,
  B200SV_XK_SCALE = 6
};

typedef struct b200sv_gate {   /* 56 bytes */
  int32_t kind;       /* B200SV_GK_*                                                               */
  int32_t a;          /* register index (0..r-1) of the target inside its round; -1: external      */
  int32_t a2;         /* second register index (SWAP), else -1                                     */
  int32_t ext_bit;    /* a == -1: global bit position of a diagonal gate's target                  */
  uint64_t ctrl_ext;  /* all-ones controls outside the round's register bits (global bit mask)     */
  int32_t ctrl_reg;   /* all-ones controls inside the round's register bits (register-index mask)  */
  int32_t slot;       /* parameter-table row, or -1 when the parameter is a constant               */
  int32_t cidx;       /* constant-pool offset in complex128 units, or -1                           */
  int32_t grad_row;   /* adjoint: row of the gradient table this gate accumulates into, or -1      */
  int32_t flags;
  int32_t chain;      /* >= 1: leader of `chain` consecutive records merged into one 2x2; 0: follower */
  int32_t xidx;       /* index of the resolved entry inside its sweep (followers repeat the leader's) */
  int32_t xkind;      /* B200SV_XK_* of the resolved entry (leader)                                */
} b200sv_gate;

#define B200SV_GF_MSLOT 2 /* gate.flags: chain executed in the round's unrolled matrix phase      */
#define B200SV_XK_PERM 7  /* X / CNOT / SWAP folded into the round's index permutation (not executed) */
#define B200SV_RF_PERM 1  /* round.flags: the permutation phase is not the identity                  */

/* A round = three phases over the 2^r amplitudes each thread holds in registers:
 *   M: one merged 2x2 per register bit (mslot[a], unrolled, no dispatch);
 *   G: generic resolved entries [first_exec, first_exec + n_exec) (controlled gates, external
 *      diagonals, scales ...), interpreted in order;
 *   P: uncontrolled X, singly-controlled X (CNOT) and SWAP gates on tile bits, composed into one
 *      GF(2)-affine map of the tile index  pi(i) = XOR_{bits of i} pcol[bit] ^ pconst ^ XOR_k pext[k].vec
 *      (k over external control bits that are 1 for this CTA) and applied for free by storing the
 *      registers at pi(index) instead of index.  Backward execution loads from pi(index). */
typedef struct b200sv_round {  /* 112 bytes */
  int32_t first_gate; /* absolute index into the gate array                                        */
  int32_t n_gates;
  uint8_t regpos[4];  /* tile-local bit positions held in registers (ascending)                    */
  uint8_t thrpos[12]; /* tile-local bit position of thread-index bit i                             */
  int32_t first_exec; /* G-phase resolved entries of this round, relative to the sweep             */
  int32_t n_exec;
  int32_t mslot[4];   /* M-phase: resolved-entry index (relative to the sweep) for register bit a, or -1 */
  uint16_t pcol[12];  /* P-phase: image of tile-local bit i                                        */
  uint16_t pconst;
  uint16_t n_pext;
  struct { uint8_t bit; uint8_t pad; uint16_t vec; } pext[4]; /* external control bit (global position) -> xor vector */
  int32_t flags;      /* B200SV_RF_*                                                               */
  int32_t mslot4;     /* M-phase entry of register bit 4 (lean sweeps with r = 5), or -1                   */
  uint8_t regpos4;    /* tile-local bit position of register bit 4                                         */
  uint8_t pad4[11];
} b200sv_round;

typedef struct b200sv_sweep {  /* 128 bytes: one kernel launch = one read + one write of the state */
  int32_t m;            /* tile bits: 2^m amplitudes staged in shared memory per CTA               */
  int32_t r;            /* register bits: 2^r amplitudes per thread per round                      */
  int32_t first_round;
  int32_t n_rounds;
  int32_t first_gate;
  int32_t n_gates;
  int32_t n_outer;      /* = n_local - m                                                           */
  int32_t n_exec;       /* resolved entries (chains) in this sweep                                  */
  uint8_t tile_bits[16];  /* global bit position of tile-local bit i (ascending)                   */
  uint8_t outer_bits[44]; /* global bit positions enumerated by the CTA index (ascending)          */
  uint8_t io_mode;      /* B200SV_IO_*: which kernel runs the sweep and how its tiles move                 */
  uint8_t pad_io[3];
  int32_t first_exec;   /* offset of this sweep's resolved entries in the plan-wide tables          */
  int16_t n_seg;        /* CTA index -> base index: sum_s ((c >> src) & (2^len - 1)) << dst; -1: bit loop */
  int16_t n_gen;        /* generic-phase entries in this sweep (0: the lean kernel variant is launched) */
  struct { uint8_t src, len, dst, pad; } seg[6];
} b200sv_sweep;

/* Which kernel executes a sweep (chosen by the planner, b200sv_lean_supported):
 *   GENERIC   sweep_kernel: any gate mix (controlled gates, external diagonals, scales), cp.async tile staging;
 *   LEAN      lean_sweep_kernel: merged 1-qubit chains + folded permutations only, persistent CTAs, normalised 2x2,
 *             tiles gathered with cp.async (the tile is not a box of the state tensor);
 *   LEAN_TMA  the same kernel, tiles staged by TMA (cp.async.bulk.tensor through the box described by b200sv_tma). */
enum { B200SV_IO_GENERIC = 0, B200SV_IO_LEAN = 1, B200SV_IO_LEAN_TMA = 2 };

/* The tile of a LEAN_TMA sweep as a box of the state seen as a rank-`rank` tensor (innermost dimension first).
 * Dimension d covers index bits [start[d], start[d] + nbits[d]) of the amplitude index; a tile dimension is taken whole
 * (box = size), a gap dimension one element at a time: its coordinate is bits [src[d], src[d] + nbits[d]) of the tile
 * counter.  The last dimension is the remaining high index bits times the batch. */
typedef struct b200sv_tma {  /* 24 bytes */
  int32_t rank;       /* bits 0..7: rank; bit 8: CU_TENSOR_MAP_SWIZZLE_128B (the lean tables then address the staged tile
                         with 16-byte chunk bits 4..6 XOR-ed with address bits 7..9) */
  uint8_t start[5], nbits[5], is_tile[5], src[5];
} b200sv_tma;

/* Lean index tables: B200SV_LT_STRIDE uint32 words per round and direction (built by qadence_b200/plan.py).
 *   [0,16)  XOR contribution of thread-index bits 0..3 (entry v = XOR over the set bits of v), [16,32) of bits 4..7:
 *           low 16 bits = byte offset on the LOAD side of the round, high 16 bits = on the STORE side.  The kernel only
 *           follows these offsets; the planner decides what they mean: either every round applies its folded X / CNOT / SWAP
 *           permutation physically (forward: load = plain index, store = permuted index; backward: the other way round),
 *           or the permutations are tracked — a round loads the amplitude of position x from S_k(x) and stores it in place,
 *           S_{k+1} = S_k o pi_k^-1, and only the last round of the sweep stores to the true positions (default);
 *   [32,37) contribution of register bit a;   [37] bit 0: some thread stores where another one loads, i.e. a barrier is
 *           needed between the round's loads and stores;
 *   [38,43) resolved-entry index (relative to the sweep) of register bit a's merged 2x2, or 0xffff;
 *   [43]    n_ext | bit_k << (8 + 6k): index bits outside the tile (controls of folded CNOTs) whose value shifts the
 *           offsets of this round;  [44,48) their XOR vectors (load half and store half like the entries above). */
#define B200SV_LT_STRIDE 48

/* Device-resident plan: the three arrays above plus the constant pool, uploaded once at convert
 * time (replaces the nn.Module tree built by backends/pyqtorch/convert_ops.py:177-351). */
typedef struct b200sv_plan {
  const b200sv_sweep* sweeps;   /* device */
  const b200sv_round* rounds;   /* device */
  const b200sv_gate* gates;     /* device */
  const double* consts;         /* device, complex128 pool as (re, im) pairs */
  const b200sv_sweep* sweeps_host; /* host copy of `sweeps` (launch geometry) */
  int32_t n_sweeps;
  int32_t n_exec_total;         /* resolved entries over all sweeps */
  const int32_t* exec_leader;   /* device: gate-record index of the leader of resolved entry x */
  const uint32_t* lean_tab_fwd; /* device: [n_rounds_total][B200SV_LT_STRIDE], forward execution */
  const uint32_t* lean_tab_bwd; /* device: the same for backward (adjoint) execution */
  const b200sv_tma* tma_host;   /* host: one record per sweep (rank 0 unless io_mode == LEAN_TMA) */
  const uint8_t* exec_mode;     /* device: io_mode of the sweep that owns resolved entry x */
} b200sv_plan;

/* ---- Pauli-sum records (observables, Rydberg / Pauli-like HamEvo generators) ----------------- */
typedef struct b200sv_pterm {  /* 32 bytes: coef * (-i)^{nY} * (-1)^{popc(i & zmask)} * psi[i ^ xmask] */
  uint64_t xmask;     /* bit positions carrying X or Y */
  uint64_t zmask;     /* bit positions carrying Z or Y */
  int32_t coef_row;   /* row of the coefficient table */
  int32_t n_y;        /* number of Y factors (phase (-i)^{n_y}) */
  int32_t pad[2];
} b200sv_pterm;

/* ---- entry points ---------------------------------------------------------------------------- */

int b200sv_abi_version(void);
/* 1 if lean_sweep_kernel is instantiated for tiles of 2^m amplitudes with r register bits (host-side query used by the
 * planner to choose io_mode; needs no device). */
int b200sv_lean_supported(int dtype, int m, int r, int adjoint);
/* number of SMs of the current device, or a negative error code (no device => B200SV_E_NODEVICE:
 * there is no CPU fallback anywhere in this library). */
int b200sv_device_sm_count(void);

/* |0...0> for every batch element.  Replaces pyq.QuantumCircuit.init_state
 * (backends/pyqtorch/backend.py:171). */
int b200sv_init_zero_state(void* state, int dtype, int n_qubits, int64_t batch, void* stream);

/* Workspace (device bytes) that apply / adjoint need for `batch` states: the table of resolved gates
 * (one 2x2 per chain and batch element, sincos evaluated ONCE per launch by a small resolve kernel
 * instead of once per CTA) and, for the adjoint, the moment accumulators.  Never allocated here. */
int64_t b200sv_plan_workspace_bytes(const b200sv_plan* plan, int dtype, int64_t batch, int adjoint);

/* Forward execution of sweeps [first, last) of a plan, in place.  Replaces the per-gate einsum loop
 * `circuit.native.run(state, values)` (backends/pyqtorch/backend.py:176).
 * `index_hi` is OR-ed into the basis index seen by control / diagonal predicates (rank bits of a
 * state sharded by its highest-order qubits); 0 on a single GPU. */
int b200sv_apply_plan(void* state, int dtype, int n_qubits, int64_t batch, const double* params,
                      const b200sv_plan* plan, int first, int last, uint64_t index_hi, void* workspace,
                      void* stream);

/* Adjoint (reverse) execution of sweeps [first, last) of a plan on psi AND lambda, accumulating
 * grads[grad_row][b] += 2 Re<lambda| dU/dθ |psi> for every gate flagged B200SV_GF_GRAD.
 * Replaces pyqtorch.differentiation.AdjointExpectation.backward
 * (call site engines/torch/differentiable_expectation.py:153-165). */
int b200sv_adjoint_plan(void* psi, void* lam, int dtype, int n_qubits, int64_t batch,
                        const double* params, double* grads, const b200sv_plan* plan, int first,
                        int last, uint64_t index_hi, void* workspace, void* stream);

/* In the three Pauli-sum entry points the first `n_single` terms must be single-Z strings (xmask == 0,
 * exactly one zmask bit): they are additive over index bits and are evaluated once per thread instead of
 * once per amplitude (total_magnetization, detunings, the linear part of N_i N_j).  Pass 0 if unsorted. */

/* out[b] (+)= Re <psi_b| sum_t coef[row_t][b] P_t |psi_b>.  coef is a device table
 * `double[n_rows][coef_ld][2]` (complex), coef_ld = batch or 1 (broadcast).
 * Replaces pyq.Observable.expectation (backends/pyqtorch/backend.py:208,249). */
int b200sv_pauli_expectation(const void* psi, int dtype, int n_qubits, int64_t batch,
                             const b200sv_pterm* terms, int n_terms, int n_single, const double* coef,
                             int64_t coef_ld, double* out, uint64_t index_hi, void* stream);

/* out = alpha * (sum_t coef_t P_t psi) + beta * acc   (acc may be NULL when beta == 0, may alias out).
 * With alpha=1, beta=0 this is lambda = O psi (the adjoint's projected state) and the matrix-free
 * matvec of Krylov HamEvo.  Replaces pyq.Add/Scale/Sequence applied term by term
 * (backends/pyqtorch/convert_ops.py:223,276-280).  Terms whose xmask has bits >= n_qubits are
 * rejected (global qubits must be made local first). */
int b200sv_pauli_apply(const void* psi, void* out, const void* acc, int dtype, int n_qubits,
                       int64_t batch, const b200sv_pterm* terms, int n_terms, int n_single, const double* coef,
                       int64_t coef_ld, double alpha_re, double alpha_im, double beta_re,
                       double beta_im, uint64_t index_hi, void* stream);

/* The diagonal part of a Pauli sum as a table: out[r][i] = sum_t coef[row_t][r] (-1)^{popc(i & zmask_t)} (complex, `rows` = 1
 * or the batch; all terms must have xmask == 0), and the matvec that consumes it:
 *     out = alpha * (diag[b][i] psi_i + sum_{X/Y terms t} coef_t P_t psi) + beta * acc.
 * The Krylov propagator computes the table once per evolution: a Rydberg generator (66 N_iN_j + detunings + 12 drive terms,
 * analog/hamiltonian_terms.py:16-72) then costs 12 instead of 90 terms per amplitude and matvec.  Same replaced call site
 * as b200sv_pauli_apply. */
int b200sv_pauli_diagonal(const b200sv_pterm* terms, int n_terms, int n_single, const double* coef, int64_t coef_ld,
                          int n_qubits, int64_t rows, double* out, uint64_t index_hi, void* stream);
int b200sv_pauli_apply_diag(const void* psi, void* out, const void* acc, int dtype, int n_qubits, int64_t batch,
                            const b200sv_pterm* terms, int n_terms, const double* coef, int64_t coef_ld, const double* diag,
                            int64_t diag_ld, const double* diag_scale, double alpha_re, double alpha_im, double beta_re,
                            double beta_im, void* stream);
/* ... and the expectation value with the same table: out[b] += Re sum_i conj(psi_i) (Re diag[b][i] psi_i + sum_{X/Y terms} ...).
 * `diag_scale` above (nullable) is a complex factor [batch][2] of the table per batch element.  Observables with constant
 * multi-qubit Z strings (Ising / zz_hamiltonian observables, constructors/hamiltonians.py) take this path: the table
 * (2^n complex numbers, L2-resident for n <= 22) is built once per observable.  Replaces obs.native.expectation
 * (backends/pyqtorch/backend.py:208,249). */
int b200sv_pauli_expectation_diag(const void* psi, int dtype, int n_qubits, int64_t batch, const b200sv_pterm* terms,
                                  int n_terms, const double* coef, int64_t coef_ld, const double* diag, int64_t diag_ld,
                                  double* out, void* stream);

/* psi[b][i] *= exp(-i t[b] d_b(i)), d_b(i) = sum_t coef[row_t][b] (-1)^{popc(i & zmask_t)} evaluated
 * on the fly from the index bits (all terms must have xmask == 0).  The diagonal Rydberg interaction
 * sum_{i<j} C6/r^6 N_i N_j of analog blocks (analog/hamiltonian_terms.py:16-45) takes this path.
 * Replaces pyq.HamiltonianEvolution's dense matrix_exp (backends/pyqtorch/convert_ops.py:258-269). */
int b200sv_diag_evolve(void* psi, int dtype, int n_qubits, int64_t batch, const b200sv_pterm* terms,
                       int n_terms, int n_single, const double* coef, int64_t coef_ld, const double* t,
                       int64_t t_ld, double sign, uint64_t index_hi, void* stream);

/* Batched BLAS-1 on states (Krylov / Lanczos plumbing), all per batch element b:
 *   dot:   out[b] (complex, 2 doubles) = <x_b|y_b>
 *   axpy:  y_b += a[b] * x_b         (a complex table [batch][2])
 *   scal:  x_b *= a[b]
 *   lincomb: out_b = sum_k c[k][b] * V_k,b   (V: K states stacked [K][batch][2^n]) */
int b200sv_dot(const void* x, const void* y, int dtype, int n_qubits, int64_t batch, double* out,
               void* stream);
int b200sv_axpy(const double* a, const void* x, void* y, int dtype, int n_qubits, int64_t batch,
                void* stream);
int b200sv_scal(const double* a, void* x, int dtype, int n_qubits, int64_t batch, void* stream);
int b200sv_lincomb(const double* c, const void* V, int K, void* out, int dtype, int n_qubits,
                   int64_t batch, void* stream);

/* One Lanczos step after the matvec (Krylov HamEvo), fused into three passes over the state, per batch element b:
 *     w <- w - beta_prev[b] v_prev   (v_prev may be NULL: first step);   a = Re<v, w>;   w <- w - a v;
 *     c = <v, w> (one re-orthogonalisation pass);   v_next = (w - c v) / b,  b = sqrt(||w||^2 - |c|^2)  written over w;
 *     alpha[b] = a + Re c,  beta[b] = b   (b = 0 and v_next = 0 when the Krylov space is exhausted).
 * `acc` is a scratch of 4 doubles per batch element that the caller zeroes before the call.  Replaces nine batched BLAS-1
 * launches (dot / axpy / scal / copy) and the torch glue between them per step of hamevo.krylov_expm. */
int b200sv_lanczos_step(void* w, const void* v, const void* v_prev, const double* beta_prev, double* acc, double* alpha,
                        double* beta, int dtype, int n_qubits, int64_t batch, void* stream);

/* Krylov plumbing: y_b = exp(-i t_b T_b) e_1 for the batched real symmetric tridiagonal T_b built by
 * Lanczos (diagonal alpha[j][b], off-diagonal beta[j][b], j < m <= 64), by scaled Taylor series on
 * the device (no dense eigendecomposition, no host round trip), plus the a-posteriori error estimate
 * err[b] = |beta[m-1][b]| * |y_b[m-1]|.  y is `double[m][batch][2]`. */
int b200sv_tridiag_expm(const double* alpha, const double* beta, const double* t, int m, int64_t batch,
                        double* y, double* err, void* stream);

/* Constant dense 2^k x 2^k matrix (complex128, row-major, targets[0] = most significant operand
 * bit) applied out of place: out = (M ⊗ I) psi.  MatrixBlock / Projector(ket,bra)
 * (backends/pyqtorch/convert_ops.py:271-272, 287-294).  k <= 10. */
int b200sv_apply_dense(const void* psi, void* out, int dtype, int n_qubits, int64_t batch,
                       const double* matrix, const int32_t* target_bits, int k, void* stream);

/* Sampling (replaces pyq.QuantumCircuit.sample, backends/pyqtorch/backend.py:300-302): inverse-CDF
 * on SUPPLIED uniforms through a fixed-shape sum tree, bit-reproducible on any machine:
 *   probs[b][i] = re*re + im*im (separately rounded), leaves = sequential sums of `chunk`
 *   consecutive probabilities, inner nodes = left + right; descend with v = u*root.
 * `tree` is a workspace of 2*(2^n/chunk) doubles per batch element (chunk = min(2^n, 256));
 * `probs` 2^n doubles per batch element.  out_idx[b][s] = sampled basis index (int64). */
int b200sv_sample(const void* psi, int dtype, int n_qubits, int64_t batch, double* probs,
                  double* tree, const double* uniforms, int64_t n_shots, int64_t* out_idx,
                  void* stream);

/* out[b][rev_n(i)] = in[b][i]: Endianness.LITTLE views (transpile/invert.py:95-109). */
int b200sv_bit_reverse(const void* in, void* out, int dtype, int n_qubits, int64_t batch,
                       void* stream);

/* out[b][P(i)] = in[b][i] where P moves index bit j to position perm[j] (a bijection on n bits);
 * used to re-index local amplitudes around the all-to-all of a global<->local qubit swap. */
int b200sv_permute_bits(const void* in, void* out, int dtype, int n_qubits, int64_t batch,
                        const int32_t* perm, void* stream);

/* One register sharded over several GPUs of a box by its highest-order qubits, WITHOUT exchange passes: forward sweeps
 * [first_sweep, last_sweep) of a plan built for all `n_total` qubits, where `peer_states[r]` is the shard of rank r
 * (2^(n_total - log2 n_peers) amplitudes; peer_states[rank] is this GPU's own buffer, the others are NVLink peer
 * mappings opened with b200sv_ipc_open).  Index bits >= n_local select the buffer, so a gate on a "global" qubit is a
 * sweep whose tiles straddle shards: the exchange that the reference design would do as a separate all-to-all
 *  is fused into the sweep's own tile loads / stores.  This GPU processes the rank-th slice of the
 * sweep's tiles; every tile is read and written by exactly one CTA on one GPU.  The CALLER puts a cross-GPU barrier on
 * the stream between two sweeps.  Tile geometry r must be 3 or 4. */
int b200sv_apply_plan_peers(void* const* peer_states, int n_peers, int rank, int dtype, int n_total, const double* params,
                            const b200sv_plan* plan, int first_sweep, int last_sweep, void* workspace, void* stream);

/* The adjoint counterpart: sweeps [first_sweep, last_sweep) walked backwards on psi AND lambda, both sharded the same way
 * (`peer_lams[r]` = the lambda shard of rank r).  Every rank accumulates into ITS `grads` (double[n_rows], batch 1) the
 * contributions 2 Re<lambda| dU |psi> of the tiles it processes; the caller all-reduces (SUM) the gradient vector over the
 * ranks and puts a cross-GPU barrier on the stream between two sweeps.  Replaces
 * pyqtorch.differentiation.AdjointExpectation.backward for a register that does not fit one GPU. */
int b200sv_adjoint_plan_peers(void* const* peer_states, void* const* peer_lams, int n_peers, int rank, int dtype, int n_total,
                              const double* params, double* grads, const b200sv_plan* plan, int first_sweep, int last_sweep,
                              void* workspace, void* stream);

/* Pauli-sum observable on a sharded register: out[0] += this rank's part of Re<psi| sum_t coef_t P_t |psi> (rows of its own
 * shard; the partner amplitudes psi[i ^ xmask] of terms with X / Y on global qubits are read from the peer shards).  The
 * caller all-reduces (SUM) `out`.  `coef` is `double[n_rows][1][2]`.  Replaces pyq.Observable.expectation
 * (backends/pyqtorch/backend.py:208,249) on a sharded state. */
int b200sv_pauli_expectation_peers(void* const* peer_states, int n_peers, int rank, int dtype, int n_total,
                                   const b200sv_pterm* terms, int n_terms, int n_single, const double* coef, double* out,
                                   void* stream);

/* out_local = alpha * (sum_t coef_t P_t psi) restricted to this rank's rows: lambda = O psi of the adjoint pass on a sharded
 * register (out_local must not be one of the psi shards; barrier before — psi final everywhere — and after). */
int b200sv_pauli_apply_peers(void* const* peer_states, int n_peers, int rank, void* out_local, int dtype, int n_total,
                             const b200sv_pterm* terms, int n_terms, int n_single, const double* coef, double alpha_re,
                             double alpha_im, void* stream);

/* Device memory that other processes of the box can map (cudaMalloc + CUDA IPC); `handle64` is 64 opaque bytes to ship to
 * the peers, who call b200sv_ipc_open (enables peer access lazily) and b200sv_ipc_close. */
int b200sv_ipc_alloc(int64_t bytes, void** ptr, void* handle64);
int b200sv_ipc_open(const void* handle64, void** ptr);
int b200sv_ipc_close(void* ptr);
int b200sv_ipc_free(void* ptr);

/* Device-side parameter embedding (replaces the per-call host loop of qadence/blocks/embedding.py:118-167, which
 * evaluates one sympytorch module per unique expression and builds a dict, and its autograd graph).
 * Every slot of the gate-parameter table is a small expression DAG over root parameters, stored as `b200sv_enode`
 * records in topological order (operands `a`, `b` are node indices relative to the slot's first node; the slot's value is
 * its last node):
 *     out[slot][e] = f_slot(var[0..], feat[0..][e])            forward
 *     gloads[L][e] = gout[slot][e] * d f_slot / d(load L)      backward: one row per VAR / FEAT node (`load` = L)
 * `var` holds the batch-independent (variational) roots, `feat` is [n_feat][batch] (feature inputs).  The function set is
 * the reference's SYMPY_TO_PYQ_MAPPING (backends/pyqtorch/convert_ops.py:46-61).  Slots are limited to 32 nodes. */
enum {
  B200SV_EO_CONST = 0, /* arg = index into consts */
  B200SV_EO_VAR = 1,   /* arg = index into var,  load = row of gloads */
  B200SV_EO_FEAT = 2,  /* arg = row of feat,     load = row of gloads */
  B200SV_EO_ADD = 3, B200SV_EO_MUL = 4, B200SV_EO_POW = 5,
  B200SV_EO_SIN = 6, B200SV_EO_COS = 7, B200SV_EO_TAN = 8, B200SV_EO_TANH = 9, B200SV_EO_LOG = 10, B200SV_EO_EXP = 11,
  B200SV_EO_ACOS = 12, B200SV_EO_ASIN = 13, B200SV_EO_ATAN = 14, B200SV_EO_ABS = 15,
  B200SV_EO_HEAVISIDE = 16 /* 0 / 0.5 / 1, zero gradient (parameters.py:196-205) */
};
#define B200SV_EMBED_MAX_NODES 32
typedef struct b200sv_enode {
  int32_t op, a, b, arg, load, pad0, pad1, pad2;
} b200sv_enode;

int b200sv_embed_forward(const b200sv_enode* nodes, const int32_t* slot_off, int n_slots, const double* consts,
                         const double* var, const double* feat, int64_t batch, double* out, void* stream);
int b200sv_embed_backward(const b200sv_enode* nodes, const int32_t* slot_off, int n_slots, const double* consts,
                          const double* var, const double* feat, int64_t batch, const double* gout,
                          double* gloads, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200SV_H */
