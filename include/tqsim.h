/*
 * tqsim.h -- C ABI of libtqsim: batched fp64 circuit simulation + energy evaluation on B200 (sm_100a).
 *
 * This is the drop-in boundary for the ONE hot path of Aqasch/TensorRL-QAS:
 *   environments/VQAs/VQE_qulacs*.py : Parametric_Circuit.construct_ansatz / get_energy_qulacs / get_exp_val
 * The reference has no FFI of its own (pure Python on top of the qulacs wheel); every entry point below
 * names the reference call site(s) whose work it replaces.  All paths are relative to the reference tree.
 *
 * Conventions (qulacs semantics):
 *   - little-endian basis: amplitude index i, qubit k  <->  bit (i >> k) & 1
 *   - RX/RY/RZ(theta) = exp(+i * theta/2 * P)            (opposite sign to qiskit)
 *   - CNOT: q0 = control, q1 = target
 *   - complex128 is passed as interleaved (re, im) doubles
 *
 * Error model: every function returns 0 on success and a negative TQ_E* code on failure; nothing throws or
 * aborts across the ABI.  tq_last_error(h) returns a human-readable message for the last failure on that
 * handle (or, with h == NULL, for the last failed tq_create on the calling thread).
 *
 * Ownership: the caller owns every buffer it passes in.  The library owns only the scratch it allocates for a
 * handle (state vectors, per-gate coefficient tables, partial sums) and frees it in tq_destroy.
 * Threading: a handle is single-threaded; distinct handles may be driven from distinct host threads.
 * Streams: *_dev entry points enqueue on the caller's stream (a cudaStream_t passed as void*, NULL = default
 * stream) and return without synchronising; *_host entry points copy in, run and copy out, and return when
 * the result is in the caller's host buffer.  A handle owns ONE set of scratch buffers (states, block matrices, partial
 * sums): keep at most one stream in flight per handle -- enqueueing on a second stream (or mixing a caller stream with
 * a *_host call) before the first has finished is a data race on that scratch.  The tq_set_* calls wait for the stream
 * used last.
 * There is NO CPU fallback: without a CUDA device tq_create fails with TQ_ENODEV.
 */
#ifndef TQSIM_H
#define TQSIM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TQ_VERSION 100

/* gate kinds for tq_set_circuit */
enum {
    TQ_RX = 0,     /* environments/VQAs/VQE_qulacs.py:36  add_parametric_RX_gate(q0, theta) */
    TQ_RY = 1,     /* environments/VQAs/VQE_qulacs.py:38  add_parametric_RY_gate(q0, theta) */
    TQ_RZ = 2,     /* environments/VQAs/VQE_qulacs.py:40  add_parametric_RZ_gate(q0, theta) */
    TQ_CNOT = 3,   /* environments/VQAs/VQE_qulacs.py:25  add_gate(CNOT(ctrl=q0, targ=q1)) */
    TQ_X = 4,      /* fixed Pauli gates (explicit noise insertions, state preparation in tests) */
    TQ_Y = 5,
    TQ_Z = 6,
    TQ_DEPOL1 = 7, /* environments/VQAs/VQE_qulacs_noise.py:45-54  DepolarizingNoise(q0, p)  (p in `fixed`) */
    TQ_DEPOL2 = 8  /* environments/VQAs/VQE_qulacs_noise.py:32-33  TwoQubitDepolarizingNoise(q0, q1, p) */
};

/* error codes */
enum {
    TQ_OK = 0,
    TQ_EINVAL = -1,  /* bad argument */
    TQ_ENODEV = -2,  /* no usable CUDA device / wrong architecture */
    TQ_ECUDA = -3,   /* a CUDA runtime call failed (message has the CUDA error string) */
    TQ_ENOMEM = -4,  /* device or host allocation failed */
    TQ_ESTATE = -5   /* call sequence error (e.g. energy before a circuit/Hamiltonian was set) */
};

typedef struct tq_context* tq_handle;

/* library/ABI version (TQ_VERSION of the build) */
int tq_version(void);

/* One handle per (n_qubits, device).  Replaces qulacs.QuantumState(n) + ParametricQuantumCircuit(n) construction:
 * environments/VQAs/VQE_qulacs.py:8-10, :81. */
int tq_create(int n_qubits, int device_id, tq_handle* out);
int tq_destroy(tq_handle h);
const char* tq_last_error(tq_handle h);

/* Hamiltonian as a Pauli sum.  Term t = coeff[t] * prod_q P_q with P_q = X if only xmask bit q, Z if only zmask
 * bit q, Y if both.  Bits are qubit indices of the little-endian state.  coeff_im may be NULL (all real).
 * Data contract: dmrg-to-qc/mol_data/*.npz keys `paulis`/`weights` (dmrg-to-qc/heisenberg_model.py:21-72). */
int tq_set_pauli_hamiltonian(tq_handle h, int n_terms, const uint64_t* xmask, const uint64_t* zmask,
                             const double* coeff_re, const double* coeff_im);

/* Hamiltonian as the dense 2^n x 2^n complex128 matrix the reference passes as `op`
 * (environments/VQAs/VQE_qulacs.py:84-85, E = Re(conj(psi).T @ op @ psi)); row-major; n_qubits <= 12 (the state must be
 * one tile; TQ_EINVAL otherwise -- larger registers take tq_set_pauli_hamiltonian).
 * The library keeps the non-zero entries and evaluates exactly that bilinear form over them. */
int tq_set_dense_hamiltonian(tq_handle h, const double* h_matrix_host);

/* Initial state: NULL = |0...0> (VQE_qulacs.py:81); otherwise 2^n complex128 copied from host
 * (state.load(TN_state), environments/VQAs/VQE_qulacs_TN_notin_RL.py:82-83). */
int tq_set_init_state(tq_handle h, const double* psi_host_or_null);

/* The circuit shared by all elements of a batch: G gates in application order.
 *   kind[g]       one of TQ_*
 *   q0[g], q1[g]  qubits (q1 ignored for 1-qubit kinds; CNOT: q0 control, q1 target)
 *   param_idx[g]  rotations: column of the parameter matrix that holds theta, or -1 to use fixed[g];
 *                 TQ_DEPOL1/2: noise-slot index (column of the trajectory code matrix), or -1
 *   fixed[g]      rotations with param_idx -1: theta; TQ_DEPOL1/2: probability p
 * Replaces construct_ansatz's add_gate/add_parametric_* calls (VQE_qulacs.py:12-44) and the
 * set_parameter loop (VQE_qulacs.py:73-74): parameters are bound per batch element at evaluation time. */
int tq_set_circuit(tq_handle h, int n_gates, const int32_t* kind, const int32_t* q0, const int32_t* q1,
                   const int32_t* param_idx, const double* fixed, int n_params);

/* Energies of B parameter sets on the pure-state path: out[b] = <psi(params[b])| H |psi(params[b])>.
 * params is [B][ld_params] row-major (ld_params >= n_params).  TQ_DEPOL* gates are skipped (noise-free).
 * Replaces get_energy_qulacs + get_exp_val (VQE_qulacs.py:47-86), B calls at a time. */
int tq_energy_batch(tq_handle h, int batch, const double* params_dev, int ld_params, double* out_dev,
                    void* stream);
int tq_energy_batch_host(tq_handle h, int batch, const double* params_host, int ld_params, double* out_host);

/* Energies of n_problems DIFFERENT problems in one launch: problem i is handles[i] -- its own circuit, Hamiltonian and
 * initial state -- evaluated at params_host[i] (n_params doubles of that handle; the pointer may be NULL when the circuit
 * has no parameters) and, when codes_host and codes_host[i] are non-NULL, with that trajectory-noise code row.  All
 * handles share the device and n_qubits, and every problem must fit one tile (n_qubits <= 12).  out_host[i] receives
 * energy i.  Replaces B independent environments each calling get_energy_qulacs once (one COBYLA round of a lock-step
 * multi-environment driver; environments/environment_qulacs.py:429-433 called from B environments at a time). */
int tq_energy_multi_host(int n_problems, tq_handle* handles, const double* const* params_host,
                         const uint8_t* const* codes_host, double* out_host);

/* Pauli-trajectory noise on the pure-state path (what the reference's noise environments actually run:
 * one sampled Pauli per noise gate and per evaluation, VQE_qulacs_noise.py:31-33,44-54).
 * codes is [B][ld_codes] uint8: for a TQ_DEPOL1 gate with slot s, codes[b][s] in {0:I,1:X,2:Y,3:Z} acts on q0;
 * for TQ_DEPOL2, codes[b][s] = pa + 4*pb with pa on q0 and pb on q1.  The caller samples the codes. */
int tq_energy_traj_batch(tq_handle h, int batch, const double* params_dev, int ld_params,
                         const uint8_t* codes_dev, int ld_codes, double* out_dev, void* stream);
int tq_energy_traj_batch_host(tq_handle h, int batch, const double* params_host, int ld_params,
                              const uint8_t* codes_host, int ld_codes, double* out_host);

/* Exact density-matrix path: rho is evolved as a 2n-qubit vector, TQ_DEPOL* gates are the exact channels
 * (mean of the reference's sampled trajectories); out[b] = Tr(rho H).  n_qubits <= 13. */
int tq_energy_dm_batch(tq_handle h, int batch, const double* params_dev, int ld_params, double* out_dev,
                       void* stream);
int tq_energy_dm_batch_host(tq_handle h, int batch, const double* params_host, int ld_params, double* out_host);

/* Final states psi(params[b]) for b in [0, B): states is [B][2^n] complex128 (debugging / parity;
 * replaces state.get_vector(), VQE_qulacs.py:84). */
int tq_state_batch(tq_handle h, int batch, const double* params_dev, int ld_params, double* states_dev,
                   void* stream);
int tq_state_batch_host(tq_handle h, int batch, const double* params_host, int ld_params, double* states_host);

/* Device-resident states evolved IN PLACE: states_dev is [B][2^n] complex128 and every element starts from what the
 * buffer holds (not from the handle's initial state; the states need not be normalised -- every kernel is linear).  The
 * handle's circuit is applied (a gate-free circuit leaves the buffer untouched) and, when energies_dev is non-NULL,
 * energies_dev[b] = <psi_b| H |psi_b> of the evolved state.  This is the building block of single-state sharding
 * (tensorrl_qas_b200/sharded.py): each rank evolves its 2^(n-g)-amplitude shard through the
 * gates that are local in the current qubit layout and evaluates its part of the Pauli sum; the qubit exchange between
 * layouts is one all-to-all.  Replaces circuit.update_quantum_state(state) on a state that already holds amplitudes
 * (environments/VQAs/VQE_qulacs_TN_notin_RL.py:82-84) for registers too large for one GPU's share of the batch. */
int tq_evolve_states(tq_handle h, int batch, const double* params_dev, int ld_params, double* states_dev,
                     double* energies_dev, void* stream);
/* Evolution of one shard with the qubit exchange fused into the write-back (single-state sharding over n_ranks = 2, 4 or
 * 8 GPUs of one NVSwitch box).  shard_dev holds this rank's 2^n amplitudes (n = the handle's qubit count = local qubits);
 * the circuit is applied and the LAST tile pass stores every amplitude where it lives after the rank bits have been
 * swapped with the top log2(n_ranks) local qubits: amplitude idx goes to rank c = idx >> (n - g), at index
 * (rank << (n - g)) | (idx & (2^(n-g) - 1)) of recv_ptrs[c] -- rank c's receive buffer (device pointers valid on this
 * device: peer memory opened with tq_ipc_open, or local buffers; recv_ptrs is a HOST array of n_ranks addresses and must
 * not alias shard_dev).  That replaces the all-to-all between two segments (tensorrl_qas_b200/sharded.py): the NVLink
 * transfer runs tile by tile under the tensor-core work of the same kernel.  The caller synchronises the ranks (any
 * stream-ordered collective) before the receive buffers are read.  Needs tensor-core passes (n >= 9). */
int tq_evolve_states_exchange(tq_handle h, const double* params_dev, int ld_params, double* shard_dev, int n_ranks,
                              int rank, const uint64_t* recv_ptrs, void* stream);
/* Plumbing for the above: shard buffers as plain cudaMalloc allocations (base pointers, so that they can be exported), and
 * CUDA IPC handles (64 bytes) to map a peer rank's buffer into this process. */
int tq_device_alloc(int device, uint64_t bytes, void** out);
int tq_device_free(int device, void* p);
int tq_ipc_export(int device, void* base, unsigned char* handle64);
int tq_ipc_open(int device, const unsigned char* handle64, void** out);
int tq_ipc_close(int device, void* p);

/* Final density matrices, [B][4^n] complex128, entry rho[r][c] at index r + (c << n). */
int tq_dm_batch_host(tq_handle h, int batch, const double* params_host, int ld_params, double* rho_host);

/* Ask/tell COBYLA for the path's unconstrained problems (tensorrl_qas_b200/csrc/tq_cobyla.cpp; host only, no GPU).
 * Optional replacement for scipy.optimize.minimize(cost, x0, method="COBYLA", options={"maxiter": maxfun}) as the
 * reference calls it (environments/environment_qulacs.py:436-441: no constraints; scipy's defaults rhobeg = 1.0,
 * rhoend = tol = 1e-4).  Powell's COBYLA restated for m = 0; NOT trajectory-identical to scipy >= 1.16 (PRIMA), so the
 * drop-in environments keep scipy unless asked (TQ_OPTIMIZER=native).
 *   create: the first point to evaluate is x0.   ask: 0 = x_out holds the next point, 1 = finished.
 *   tell: value at the point last asked; returns 0 = ask again, 1 = finished.   result: best point, its value, number of
 *   evaluations, status (1 = rho reached rhoend, 2 = maxfun evaluations used, 3 = rounding errors in the simplex). */
typedef struct tq_cobyla* tq_cobyla_handle;
int tq_cobyla_create(int n, const double* x0, double rhobeg, double rhoend, int maxfun, tq_cobyla_handle* out);
int tq_cobyla_destroy(tq_cobyla_handle h);
int tq_cobyla_ask(tq_cobyla_handle h, double* x_out);
int tq_cobyla_tell(tq_cobyla_handle h, double f);
int tq_cobyla_result(tq_cobyla_handle h, double* x_out, double* f_out, int* nfev_out, int* status_out);

/* Introspection of the compiled plan of the current circuit (for DESIGN/bench accounting):
 *   info[0] gate passes over the state, info[1] expectation-only passes, info[2] tile qubits k,
 *   info[3] kernel launches per energy call, info[4] Hamiltonian flip-mask groups M, info[5] non-zeros kept
 *   of a dense Hamiltonian (0 if Pauli), info[6] unitary gates G, info[7] rotation gates.
 * which: 0 = pure-state plan, 1 = density-matrix plan, 2 = pure-state plan with trajectory-noise slots. */
int tq_plan_info(tq_handle h, int which, int64_t* info8);

/* Work counts of the compiled plan (bench.py's FP64 tensor-core accounting):
 *   counts[0] dense blocks executed on the FP64 tensor cores (each = 16 FMA per amplitude: a 4x4 complex block as an
 *             8x8 real DMMA product), counts[1] lane<->register qubit exchanges by shuffle, counts[2] gate windows,
 *             counts[3] expectation windows, counts[4] FP64-pipe register windows (tiles below 2^9 amplitudes and
 *             density matrices), counts[5] expectation-only passes that stream the state straight from HBM,
 *             counts[6] (dense block, tile) pairs executed per batch element with the current initial state (a run
 *             from |0...0> skips the tiles and amplitudes nothing has populated yet), counts[7] launches of the streaming
 *             pass kernel (persistent CTAs, TMA tile I/O; tq_stream.cu) issued by this handle so far. */
int tq_plan_counts(tq_handle h, int which, int64_t* counts8);

/* Number of kernel launches issued by this handle since creation (bench.py's gpu_launches). */
int64_t tq_launch_count(tq_handle h);

/* Compiled plans are cached per handle, keyed on the gate list (kinds, qubits, parameter slots, fixed angles; angles bound
 * through param_idx are run-time parameters and do not enter the key): tq_set_circuit with the circuit that is already
 * bound is a no-op, a circuit seen before gets its plan back without recompilation (least recently used of
 * TQ_PLAN_CACHE = 64 circuits evicted; setting a Hamiltonian drops the cache).  The reference rebuilds its circuit twice
 * per environment step (environments/environment_qulacs.py:409-411, 423-425) and revisits the same structures episode
 * after episode.  stats4: plans restored from the cache, circuits compiled afresh, no-op calls, circuits cached now. */
int tq_plan_cache_stats(tq_handle h, int64_t* stats4);

/* Per-launch device timing (bench.py's measured roofline; replaces profiler-derived constants).  While enabled, every
 * kernel launch of the handle is bracketed by a CUDA-event pair on the launch stream.  tq_profile_read waits for the
 * recorded launches, returns them in issue order and clears the list:
 *   kind[i]         0 prep_matrices  1 tile_pass (FP64-pipe windows)  2 tile_pass_mma  3 tile_stream (gate pass)
 *                   4 tile_stream (gate pass + expectation windows)  5 tile_stream (expectation-only sub-passes)
 *                   6 expect_direct  7 reduce_partials  8 dm_expect  9 table launch (tq_energy_multi_host)
 *   ms[i]           device time of the launch
 *   model_bytes[i]  HBM bytes the compiled plan moves in it (populated part of the state in, whole tiles out)
 *   alg_bytes[i]    bytes the reference's one-state-pass-per-gate model charges for the gates / Hamiltonian groups the
 *                   launch covers (16 * 2^n * (2 G + M) per evaluation)
 * Any output array may be NULL. */
int tq_profile_enable(tq_handle h, int on);
int tq_profile_read(tq_handle h, int max_records, int32_t* kind, float* ms, double* model_bytes, double* alg_bytes,
                    int* n_out);
/* FP64 tensor-core work (flops: 2 x 16 FMA per amplitude and fused dense block, known zeros left out) of the records the
 * last tq_profile_read handed out, same order -- what bench.py divides by the launch time for the tensor-pipe roofline. */
int tq_profile_read_flops(tq_handle h, int max_records, double* dmma_flops, int* n_out);

/* Measured FP64 peak of the device in TFLOP/s: which = 0 mma.sync.m8n8k4.f64 (the tensor-core instruction of the fused
 * dense blocks), 1 DFMA on the FP64 pipe.  Best of three timed runs of a register-resident kernel. */
int tq_fp64_peak(int device, int which, double* tflops_out);

/* Diagnostics -- planner dry run, needs no GPU: text dump ("PASS lead=.. local=a,b,.." / "OP op a b t flags fixed"
 * lines) of the tile passes the circuit compiler produces for a gate list.  which as in tq_plan_info (2 =
 * trajectory-noise plan); which | 16 additionally assigns a synthetic Hamiltonian built from cover_masks to the passes and
 * reports the streaming-kernel layouts ("STREAMABLE", "STREAM", "STREAMCHECK" lines).  The caller releases the string
 * with tq_free. */
char* tq_plan_dump(int n_qubits, int n_gates, const int32_t* kind, const int32_t* q0, const int32_t* q1,
                   const int32_t* param_idx, const double* fixed, int which, int tile_bits, int low_bits,
                   int n_cover, const uint64_t* cover_masks);
void tq_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* TQSIM_H */
