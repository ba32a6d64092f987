/*
 * tq_oracle.c -- CPU restatement (plain C + pthreads) of the TensorRL-QAS hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED against qulacs itself: qulacs (unpinned third-party dependency, reference requirements.txt:2) is
 * not importable in the build container and the reference ships no tests or golden vectors for this path
 * (SURVEY.md section 8c).  What IS pinned: (1) the gate order / parameter mapping / dense expectation, by running
 * the reference's own environments/VQAs/VQE_qulacs*.py against oracle/np_oracle.py's qulacs-shaped classes
 * (tests/golden/make_golden.py); (2) the conventions, through the shipped artefacts: the MPS init circuits
 * (dmrg-to-qc/init_state_circ, .qpy files) evaluated on the shipped Hamiltonians (dmrg-to-qc/mol_data, .npz files) land
 * within a few mHa above the shipped ground-state `eigvals` only with exactly these sign / bit-order choices.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this file's
 * shared object; the product path (libtqsim) never does.
 *
 * Algorithm = what qulacs + numpy do for the reference, one full-state pass per gate:
 *   state init         environments/VQAs/VQE_qulacs.py:81            QuantumState(n) -> |0...0>
 *   state.load         environments/VQAs/VQE_qulacs_TN_notin_RL.py:83
 *   gates              environments/VQAs/VQE_qulacs.py:25,36-40      CNOT(ctrl,targ), RX/RY/RZ = exp(+i theta/2 P)
 *   set_parameter      environments/VQAs/VQE_qulacs.py:73-74         angle j <- params[j]
 *   dense expectation  environments/VQAs/VQE_qulacs.py:84-85         Re(conj(psi).T @ op @ psi)
 *   noise gates        environments/VQAs/VQE_qulacs_noise.py:31-33,44-54   one sampled Pauli per noise gate
 *   Pauli-sum form     dmrg-to-qc/heisenberg_model.py:21-72 (term list), SURVEY.md appendix A (index formula)
 * Little-endian: qubit k <-> bit k of the amplitude index.  complex128 = interleaved (re, im) doubles.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

enum { K_RX = 0, K_RY = 1, K_RZ = 2, K_CNOT = 3, K_X = 4, K_Y = 5, K_Z = 6, K_DEPOL1 = 7, K_DEPOL2 = 8 };

typedef struct { double re, im; } c128;

static inline c128 cmul(c128 a, c128 b) { c128 r = { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re }; return r; }
static inline c128 cadd(c128 a, c128 b) { c128 r = { a.re + b.re, a.im + b.im }; return r; }

/* v <- (m on bit q) v for a vector of 2^nbits entries; m = {m00, m01, m10, m11}.
 * Pair loop in (block, offset) form so the compiler can vectorise it; diagonal matrices (RZ, Z) skip the mixing. */
static void apply_1q(int nbits, c128* v, int q, const c128 m[4]) {
    const uint64_t dim = 1ull << nbits, bit = 1ull << q;
    const int diagonal = m[1].re == 0.0 && m[1].im == 0.0 && m[2].re == 0.0 && m[2].im == 0.0;
    for (uint64_t base = 0; base < dim; base += 2 * bit) {
        c128* lo = v + base;
        c128* hi = v + base + bit;
        if (diagonal) {
            for (uint64_t j = 0; j < bit; ++j) { lo[j] = cmul(m[0], lo[j]); hi[j] = cmul(m[3], hi[j]); }
        } else {
            for (uint64_t j = 0; j < bit; ++j) {
                c128 a0 = lo[j], a1 = hi[j];
                lo[j] = cadd(cmul(m[0], a0), cmul(m[1], a1));
                hi[j] = cadd(cmul(m[2], a0), cmul(m[3], a1));
            }
        }
    }
}

static void apply_cnot(int nbits, c128* v, int ctrl, int targ) {
    const uint64_t dim = 1ull << nbits, cb = 1ull << ctrl, tb = 1ull << targ;
    for (uint64_t base = 0; base < dim; base += 2 * tb) {
        c128* lo = v + base;
        c128* hi = v + base + tb;
        for (uint64_t j = 0; j < tb; ++j)
            if ((base + j) & cb) { c128 t = lo[j]; lo[j] = hi[j]; hi[j] = t; }
    }
}

/* 2x2 matrix of a 1-qubit kind; conj != 0 gives the complex conjugate (column side of a density matrix) */
static void gate_matrix(int kind, double theta, int conj, c128 m[4]) {
    const double c = cos(0.5 * theta), s = sin(0.5 * theta);
    const double sg = conj ? -1.0 : 1.0;
    memset(m, 0, 4 * sizeof(c128));
    switch (kind) {
    case K_RX: m[0].re = c; m[1].im = sg * s; m[2].im = sg * s; m[3].re = c; break;       /* cos I + i sin X */
    case K_RY: m[0].re = c; m[1].re = s; m[2].re = -s; m[3].re = c; break;                /* cos I + i sin Y */
    case K_RZ: m[0].re = c; m[0].im = sg * s; m[3].re = c; m[3].im = -sg * s; break;      /* diag(e^{+it/2}, e^{-it/2}) */
    case K_X: m[1].re = 1; m[2].re = 1; break;
    case K_Y: m[1].im = -sg; m[2].im = sg; break;                                         /* [[0,-i],[i,0]] */
    case K_Z: m[0].re = 1; m[3].re = -1; break;
    default: m[0].re = 1; m[3].re = 1; break;
    }
}

/* one gate on a pure state */
void orc_apply_gate(int n, double* psi, int kind, int q0, int q1, double theta) {
    c128 m[4];
    if (kind == K_CNOT) { apply_cnot(n, (c128*)psi, q0, q1); return; }
    if (kind == K_DEPOL1 || kind == K_DEPOL2) return;
    gate_matrix(kind, theta, 0, m);
    apply_1q(n, (c128*)psi, q0, m);
}

static void apply_pauli_code(int n, c128* v, int q, int code) {  /* 0 I, 1 X, 2 Y, 3 Z */
    c128 m[4];
    if (code == 0) return;
    gate_matrix(K_X + code - 1, 0.0, 0, m);
    apply_1q(n, v, q, m);
}

/*
 * Run a gate list on a pure state.  params: this element's angle row; codes: this element's noise-code row or
 * NULL (noise gates skipped).  Same argument meaning as tq_set_circuit (include/tqsim.h).
 */
void orc_run_circuit(int n, double* psi, int G, const int32_t* kind, const int32_t* q0, const int32_t* q1,
                     const int32_t* pidx, const double* fixed, const double* params, const uint8_t* codes) {
    c128* v = (c128*)psi;
    for (int g = 0; g < G; ++g) {
        int k = kind[g];
        if (k == K_DEPOL1) {
            if (codes && pidx[g] >= 0) apply_pauli_code(n, v, q0[g], codes[pidx[g]] & 3);
        } else if (k == K_DEPOL2) {
            if (codes && pidx[g] >= 0) {
                apply_pauli_code(n, v, q0[g], codes[pidx[g]] & 3);
                apply_pauli_code(n, v, q1[g], (codes[pidx[g]] >> 2) & 3);
            }
        } else {
            double th = (k <= K_RZ) ? (pidx[g] >= 0 ? params[pidx[g]] : fixed[g]) : 0.0;
            orc_apply_gate(n, psi, k, q0[g], q1[g], th);
        }
    }
}

/* Re(psi^dagger H psi), H dense row-major 2^n x 2^n complex128 (VQE_qulacs.py:85) */
double orc_expect_dense(int n, const double* psi, const double* H) {
    const uint64_t dim = 1ull << n;
    const c128* v = (const c128*)psi;
    const c128* h = (const c128*)H;
    double acc = 0.0;
    for (uint64_t r = 0; r < dim; ++r) {
        c128 row = { 0, 0 };
        for (uint64_t c = 0; c < dim; ++c) row = cadd(row, cmul(h[r * dim + c], v[c]));
        acc += v[r].re * row.re + v[r].im * row.im; /* Re(conj(v_r) * row) */
    }
    return acc;
}

static inline int parity64(uint64_t x) { return __builtin_parityll(x); }

/* sum_t coeff_t <psi| P_t |psi>, P_t |i> = i^{ny} (-1)^{popcount(i & z)} |i ^ x>   */
double orc_expect_pauli(int n, const double* psi, int T, const uint64_t* xm, const uint64_t* zm,
                        const double* cre, const double* cim) {
    const uint64_t dim = 1ull << n;
    const c128* v = (const c128*)psi;
    double total = 0.0;
    for (int t = 0; t < T; ++t) {
        const uint64_t x = xm[t], z = zm[t];
        const int ny = __builtin_popcountll(x & z) & 3;
        c128 acc = { 0, 0 };
        for (uint64_t i = 0; i < dim; ++i) {
            c128 a = v[i], b = v[i ^ x];
            c128 p = { b.re * a.re + b.im * a.im, b.re * a.im - b.im * a.re }; /* conj(b) * a */
            if (parity64(i & z)) { acc.re -= p.re; acc.im -= p.im; } else { acc.re += p.re; acc.im += p.im; }
        }
        /* multiply by i^ny */
        c128 ph = acc;
        if (ny == 1) { ph.re = -acc.im; ph.im = acc.re; }
        else if (ny == 2) { ph.re = -acc.re; ph.im = -acc.im; }
        else if (ny == 3) { ph.re = acc.im; ph.im = -acc.re; }
        c128 w = { cre[t], cim ? cim[t] : 0.0 };
        total += w.re * ph.re - w.im * ph.im;
    }
    return total;
}

/* ---- batch driver: a pool of pthreads pulling element indices from a shared counter ---- */
typedef struct {
    int n, G, B, ld, ldc, ham_kind, T, dm;
    const double* init; const int32_t *kind, *q0, *q1, *pidx; const double *fixed, *params; const uint8_t* codes;
    const double* H; const uint64_t *xm, *zm; const double *cre, *cim; double* out;
    volatile int next;
} batch_job;

void orc_dm_run(int n, const double* init, int G, const int32_t* kind, const int32_t* q0, const int32_t* q1,
                const int32_t* pidx, const double* fixed, const double* params, double* rho_out);
double orc_dm_expect_dense(int n, const double* rho_, const double* H);
double orc_dm_expect_pauli(int n, const double* rho_, int T, const uint64_t* xm, const uint64_t* zm,
                           const double* cre, const double* cim);

static void* batch_worker(void* arg) {
    batch_job* j = (batch_job*)arg;
    const uint64_t len = 1ull << (j->dm ? 2 * j->n : j->n);
    double* buf = (double*)malloc(len * 2 * sizeof(double));
    for (;;) {
        int b = __sync_fetch_and_add(&j->next, 1);
        if (b >= j->B) break;
        const double* prm = j->params + (size_t)b * j->ld;
        if (j->dm) {
            orc_dm_run(j->n, j->init, j->G, j->kind, j->q0, j->q1, j->pidx, j->fixed, prm, buf);
            j->out[b] = j->ham_kind == 0 ? orc_dm_expect_dense(j->n, buf, j->H)
                                         : orc_dm_expect_pauli(j->n, buf, j->T, j->xm, j->zm, j->cre, j->cim);
        } else {
            if (j->init) memcpy(buf, j->init, len * 2 * sizeof(double));
            else { memset(buf, 0, len * 2 * sizeof(double)); buf[0] = 1.0; }
            orc_run_circuit(j->n, buf, j->G, j->kind, j->q0, j->q1, j->pidx, j->fixed, prm,
                            j->codes ? j->codes + (size_t)b * j->ldc : NULL);
            j->out[b] = j->ham_kind == 0 ? orc_expect_dense(j->n, buf, j->H)
                                         : orc_expect_pauli(j->n, buf, j->T, j->xm, j->zm, j->cre, j->cim);
        }
    }
    free(buf);
    return NULL;
}

int orc_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

static int run_batch(batch_job* j, int nthreads) {
    if (nthreads <= 0) nthreads = orc_max_threads();
    if (nthreads > j->B) nthreads = j->B > 0 ? j->B : 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t tid[256];
    j->next = 0;
    for (int t = 1; t < nthreads; ++t) pthread_create(&tid[t], NULL, batch_worker, j);
    batch_worker(j);
    for (int t = 1; t < nthreads; ++t) pthread_join(tid[t], NULL);
    return nthreads;
}

/*
 * Batched energies (the CPU baseline): element b starts from init (or |0..0> when NULL), runs the circuit with
 * params[b*ld .. ] (and codes[b*ldc ..] when given) and takes the expectation.  ham_kind 0 = dense (H), 1 = Pauli.
 * Threads across elements (nthreads <= 0: all online cores); returns the number of threads used.
 */
int orc_energy_batch(int n, const double* init, int G, const int32_t* kind, const int32_t* q0, const int32_t* q1,
                     const int32_t* pidx, const double* fixed, int B, const double* params, int ld,
                     const uint8_t* codes, int ldc, int ham_kind, const double* H, int T, const uint64_t* xm,
                     const uint64_t* zm, const double* cre, const double* cim, double* out, int nthreads) {
    batch_job j = { n, G, B, ld, ldc, ham_kind, T, 0, init, kind, q0, q1, pidx, fixed, params, codes,
                    H, xm, zm, cre, cim, out, 0 };
    return run_batch(&j, nthreads);
}

/* final state of one element (parity of the whole state, not just the energy) */
void orc_state(int n, const double* init, int G, const int32_t* kind, const int32_t* q0, const int32_t* q1,
               const int32_t* pidx, const double* fixed, const double* params, const uint8_t* codes, double* psi) {
    const uint64_t dim = 1ull << n;
    if (init) memcpy(psi, init, dim * 2 * sizeof(double));
    else { memset(psi, 0, dim * 2 * sizeof(double)); psi[0] = 1.0; }
    orc_run_circuit(n, psi, G, kind, q0, q1, pidx, fixed, params, codes);
}

/* ------------------------------------------------------------------ density matrix ------------------------- */
/* rho[r][c] stored at index r + (c << n): a 2n-"qubit" vector; U rho U^dagger = U on bit q, conj(U) on bit q+n.  */

static void dm_apply_pauli_both(int n, c128* rho, int q, int code) { /* rho <- P rho P */
    c128 m[4];
    if (code == 0) return;
    gate_matrix(K_X + code - 1, 0.0, 0, m);
    apply_1q(2 * n, rho, q, m);
    gate_matrix(K_X + code - 1, 0.0, 1, m);
    apply_1q(2 * n, rho, q + n, m);
}

/* exact channels by their definition: rho -> (1-p) rho + p/3 sum_{P in X,Y,Z} P rho P   (1 qubit)
 *                                     rho -> (1-p) rho + p/15 sum_{P != II} P rho P     (2 qubits) */
static void dm_depol(int n, c128* rho, int qa, int qb, double p, int two) {
    const uint64_t len = 1ull << (2 * n);
    c128* acc = (c128*)calloc(len, sizeof(c128));
    c128* tmp = (c128*)malloc(len * sizeof(c128));
    const int ncodes = two ? 16 : 4;
    const double w = two ? p / 15.0 : p / 3.0;
    for (int code = 1; code < ncodes; ++code) {
        memcpy(tmp, rho, len * sizeof(c128));
        dm_apply_pauli_both(n, tmp, qa, code & 3);
        if (two) dm_apply_pauli_both(n, tmp, qb, (code >> 2) & 3);
        for (uint64_t i = 0; i < len; ++i) { acc[i].re += tmp[i].re; acc[i].im += tmp[i].im; }
    }
    for (uint64_t i = 0; i < len; ++i) {
        rho[i].re = (1.0 - p) * rho[i].re + w * acc[i].re;
        rho[i].im = (1.0 - p) * rho[i].im + w * acc[i].im;
    }
    free(acc);
    free(tmp);
}

/* rho <- circuit(rho); init = pure state to start from (NULL = |0..0>) */
void orc_dm_run(int n, const double* init, int G, const int32_t* kind, const int32_t* q0, const int32_t* q1,
                const int32_t* pidx, const double* fixed, const double* params, double* rho_out) {
    const uint64_t dim = 1ull << n;
    c128* rho = (c128*)rho_out;
    c128 m[4];
    memset(rho, 0, dim * dim * sizeof(c128));
    if (init) {
        const c128* v = (const c128*)init;
        for (uint64_t c = 0; c < dim; ++c)
            for (uint64_t r = 0; r < dim; ++r) { /* rho = |v><v| : rho[r][c] = v_r conj(v_c) */
                rho[r + (c << n)].re = v[r].re * v[c].re + v[r].im * v[c].im;
                rho[r + (c << n)].im = v[r].im * v[c].re - v[r].re * v[c].im;
            }
    } else rho[0].re = 1.0;
    for (int g = 0; g < G; ++g) {
        int k = kind[g];
        if (k == K_CNOT) {
            apply_cnot(2 * n, rho, q0[g], q1[g]);
            apply_cnot(2 * n, rho, q0[g] + n, q1[g] + n);
        } else if (k == K_DEPOL1) dm_depol(n, rho, q0[g], 0, fixed[g], 0);
        else if (k == K_DEPOL2) dm_depol(n, rho, q0[g], q1[g], fixed[g], 1);
        else {
            double th = (k <= K_RZ) ? (pidx[g] >= 0 ? params[pidx[g]] : fixed[g]) : 0.0;
            gate_matrix(k, th, 0, m);
            apply_1q(2 * n, rho, q0[g], m);
            gate_matrix(k, th, 1, m);
            apply_1q(2 * n, rho, q0[g] + n, m);
        }
    }
}

/* Re Tr(rho H), H dense */
double orc_dm_expect_dense(int n, const double* rho_, const double* H) {
    const uint64_t dim = 1ull << n;
    const c128* rho = (const c128*)rho_;
    const c128* h = (const c128*)H;
    double acc = 0.0;
    for (uint64_t r = 0; r < dim; ++r)
        for (uint64_t c = 0; c < dim; ++c) { /* Tr(rho H) = sum_{r,c} rho[r][c] H[c][r] */
            c128 a = rho[r + (c << n)], b = h[c * dim + r];
            acc += a.re * b.re - a.im * b.im;
        }
    return acc;
}

/* Re Tr(rho sum_t coeff_t P_t):  Tr(rho P) = sum_i rho[i^x][i] * phase_P(i)  with P|i> = phase_P(i) |i^x> */
double orc_dm_expect_pauli(int n, const double* rho_, int T, const uint64_t* xm, const uint64_t* zm,
                           const double* cre, const double* cim) {
    const uint64_t dim = 1ull << n;
    const c128* rho = (const c128*)rho_;
    double total = 0.0;
    for (int t = 0; t < T; ++t) {
        const uint64_t x = xm[t], z = zm[t];
        const int ny = __builtin_popcountll(x & z) & 3;
        c128 acc = { 0, 0 };
        for (uint64_t i = 0; i < dim; ++i) {
            /* <i^x| P |i> = phase(i); Tr(rho P) = sum_i <i| rho P |i> = sum_i rho[i][i^x] * <i^x|P|i> */
            c128 a = rho[i + ((i ^ x) << n)];
            if (parity64(i & z)) { acc.re -= a.re; acc.im -= a.im; } else { acc.re += a.re; acc.im += a.im; }
        }
        c128 ph = acc;
        if (ny == 1) { ph.re = -acc.im; ph.im = acc.re; }
        else if (ny == 2) { ph.re = -acc.re; ph.im = -acc.im; }
        else if (ny == 3) { ph.re = acc.im; ph.im = -acc.re; }
        c128 w = { cre[t], cim ? cim[t] : 0.0 };
        total += w.re * ph.re - w.im * ph.im;
    }
    return total;
}

/* batched density-matrix energies; ham_kind as in orc_energy_batch */
int orc_dm_energy_batch(int n, const double* init, int G, const int32_t* kind, const int32_t* q0,
                        const int32_t* q1, const int32_t* pidx, const double* fixed, int B, const double* params,
                        int ld, int ham_kind, const double* H, int T, const uint64_t* xm, const uint64_t* zm,
                        const double* cre, const double* cim, double* out, int nthreads) {
    batch_job j = { n, G, B, ld, 0, ham_kind, T, 1, init, kind, q0, q1, pidx, fixed, params, NULL,
                    H, xm, zm, cre, cim, out, 0 };
    return run_batch(&j, nthreads);
}
