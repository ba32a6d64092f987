"""The BASELINE.json configurations C1-C5 as concrete workloads, shared by bench.py
(`--workload`) and tests/test_configs.py.  Inputs come from the committed fixtures under tests/golden/ (made from the
reference's shipped artefacts by the scripts next to them), never from /root/reference at run time.

  C1   4 qubits, LiH-4q dense Hamiltonian (the shipped parity file), loaded state from a 27-gate brickwork, 20 agent gates;
       the reference's own regime: B = 1 per call (latency), and B = 4096
  C2   BeH2-6q trainable-environment circuit (shipped QPY, mirrored / negated / float32 angles) + 20 agent gates, B = 256
  C3   H2O-8q fixed environment: TN state from the shipped QPY loaded, bit-reversed Hamiltonian, 20 agent gates, B = 4096
  C4   H2O-8q exact depolarising density matrix (2^16 entries): shipped QPY circuit + 40 agent gates, p1 = 0.01, p2 = 0.05, B = 64
  C5   20-qubit Heisenberg chain, 440-gate brickwork circuit from |0...0>, B = 64 per GPU (the headline)
  C5L  C5 with a loaded (dense, random) initial state: the known-zero skipping of a run from |0...0> is off
  C5G  20 qubits, 440 gates straight from the generic generator (no brick structure), same Hamiltonian
"""
import os
from dataclasses import dataclass, field

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(ROOT, "tests", "golden")
NAMES = ("C1", "C2", "C3", "C4", "C5", "C5L", "C5G")


@dataclass
class Workload:
    name: str
    title: str
    n: int
    gl: object                      # tensorrl_qas_b200.circuit.GateList
    mode: str = "pure"              # "pure" | "dm"
    dense: object = None            # dense Hamiltonian (n <= 12) ...
    pauli: object = None            # ... or (xmask, zmask, coeff)
    init: object = None             # loaded initial state or None (|0...0>)
    batch: int = 64                 # parameter sets per GPU and step
    groups: int = 0                 # Hamiltonian flip-mask groups M (SURVEY.md section 8d traffic model)
    bound: str = "hbm"              # what bounds the path at this size (stated in the bench line)
    bound_note: str = ""
    eig_min: float = None
    extra: dict = field(default_factory=dict)

    def params(self, batch=None, seed0=1000):
        from tensorrl_qas_b200.circuit import parameter_batch
        return parameter_batch(self.gl, self.batch if batch is None else batch, seed0=seed0)

    def bind(self, sim):
        """Upload circuit / Hamiltonian / initial state to a tensorrl_qas_b200.Simulator."""
        if self.pauli is not None:
            sim.set_pauli_hamiltonian(*self.pauli)
        else:
            sim.set_dense_hamiltonian(self.dense)
        if self.init is not None:
            sim.set_init_state(self.init)
        sim.set_circuit(self.gl)
        return sim

    def oracle_energies(self, params, nthreads=0, return_threads=False):
        """The CPU restatement of the reference path on these inputs (test / baseline infrastructure)."""
        from oracle import c_oracle
        kw = dict(dense=self.dense) if self.pauli is None else dict(pauli=self.pauli)
        if self.mode == "dm":
            e = c_oracle.dm_energies(self.gl, params, nthreads=nthreads, **kw)
            return (e, c_oracle.max_threads() if nthreads == 0 else nthreads) if return_threads else e
        return c_oracle.energies(self.gl, params, init=self.init, nthreads=nthreads, return_threads=return_threads, **kw)

    def algorithmic_bytes_per_eval(self):
        """qulacs' own traffic model -- one read + one write of the state per gate, one read per
        flip-mask group, one read of a loaded initial state (density matrix: 4^n entries)."""
        state = 16 << (2 * self.n if self.mode == "dm" else self.n)
        g = self.gl.n_unitary   # (the survey's figure counts the unitary gates, G = 190 for C4)
        return state * (2 * g + max(self.groups, 1) + (1 if self.init is not None else 0))


def _case(key):
    import sys
    tests = os.path.join(ROOT, "tests")
    if tests not in sys.path:
        sys.path.insert(0, tests)
    from golden_util import Case
    return Case(key)


def _noisy(gl, p1=0.01, p2=0.05):
    from tensorrl_qas_b200.circuit import GateList
    out = GateList(gl.n_qubits)
    for kind, q0, q1, pidx, fixed in gl.tuples():
        if kind == 3:
            out.add_cnot(q0, q1)
            out.add_depol2(q0, q1, p2)
        else:
            out.add_rotation(kind, q0, fixed)
            out.add_depol1(q0, p1)
    return out


def _small_state(gl, params):
    """Statevector of a small rotation / CNOT circuit from |0...0> in plain numpy (conventions of include/tqsim.h:
    little-endian, R_P(theta) = exp(+i theta/2 P), CNOT(q0 = control, q1 = target)).  Only used to MAKE the loaded initial
    state of C1 (a workload input); every measured or checked quantity comes from libtqsim / the oracle."""
    n = gl.n_qubits
    psi = np.zeros(1 << n, dtype=np.complex128)
    psi[0] = 1.0
    idx = np.arange(1 << n)
    P = {0: np.array([[0, 1], [1, 0]], dtype=complex), 1: np.array([[0, -1j], [1j, 0]]), 2: np.diag([1.0 + 0j, -1.0])}
    for kind, q0, q1, pidx, fixed in gl.tuples():
        if kind == 3:
            sel = ((idx >> q0) & 1) == 1
            psi = np.where(sel, psi[idx ^ (1 << q1)], psi)
            continue
        theta = params[pidx] if pidx >= 0 else fixed
        U = np.cos(theta / 2) * np.eye(2) + 1j * np.sin(theta / 2) * P[kind]
        bit = (idx >> q0) & 1
        partner = psi[idx ^ (1 << q0)]
        psi = np.where(bit == 0, U[0, 0] * psi + U[0, 1] * partner, U[1, 1] * psi + U[1, 0] * partner)
    return psi


def _groups_of(xmask):
    return len(set(int(v) for v in xmask))


def build(name):
    from tensorrl_qas_b200 import loaders
    from tensorrl_qas_b200.circuit import append_random_gates, brickwork_circuit, synthetic_circuit
    name = name.upper()
    if name == "C1":
        n = 4
        g = np.load(os.path.join(GOLDEN, "lih_4q_parity.npz"))
        H = g["hamiltonian"].astype(np.complex128)
        init_gl = synthetic_circuit(n, 27, 0)
        from tensorrl_qas_b200.circuit import parameter_batch
        init = _small_state(init_gl, parameter_batch(init_gl, 1)[0])   # an INPUT of the workload: the "TN state" to load
        x, _z, _c = loaders.dense_to_pauli(H)
        return Workload("C1", "C1: 4-qubit LiH (shipped dense parity Hamiltonian), loaded state, 20 agent gates",
                        n, synthetic_circuit(n, 20, 1), dense=H, init=init, batch=4096, groups=_groups_of(x),
                        bound="latency", bound_note="state = 256 B: launch latency per call; batched: FP64 pipe / shared memory",
                        eig_min=float(g["eigvals"].min()))
    if name == "C2":
        c = _case("beh2_6q")
        gl = append_random_gates(c.gatelist("in"), 20, np.random.default_rng(2))
        x, _z = c.masks(False)
        return Workload("C2", "C2: BeH2-6q trainable environment (shipped QPY circuit + 20 agent gates), 256 parameter sets",
                        c.n, gl, dense=c.dense(False), batch=256, groups=_groups_of(x), bound="smem",
                        bound_note="state = 1 KiB, resident in shared memory: FP64 pipe / shared-memory exchanges",
                        eig_min=c.eig_min)
    if name == "C3":
        c = _case("h2o_8q")
        x, _z = c.masks(True)
        return Workload("C3", "C3: H2O-8q fixed environment (TN state loaded, bit-reversed Hamiltonian, 20 agent gates)",
                        c.n, synthetic_circuit(c.n, 20, 3), dense=c.dense(True), init=c.g["notin_tn_state"], batch=4096,
                        groups=_groups_of(x), bound="smem",
                        bound_note="state = 4 KiB, resident in shared memory: the sparse bilinear form (2176 non-zeros) on the FP64 pipe",
                        eig_min=c.eig_min)
    if name == "C4":
        c = _case("h2o_8q")
        gl = _noisy(append_random_gates(c.gatelist("in"), 40, np.random.default_rng(4)))
        x, _z = c.masks(False)
        return Workload("C4", "C4: H2O-8q exact depolarising density matrix (shipped QPY circuit + 40 agent gates, p1 = 0.01, p2 = 0.05)",
                        c.n, gl, mode="dm", dense=c.dense(False), batch=64, groups=_groups_of(x), bound="l2",
                        bound_note="rho = 1 MiB per element: L2-resident between its passes, FP64 pipe in the register windows",
                        eig_min=c.eig_min)
    if name in ("C5", "C5L", "C5G"):
        n = 20
        paulis, w = loaders.heisenberg_terms(n)
        x, z = loaders.pauli_masks(paulis, n)
        gl = synthetic_circuit(n, 440, 5) if name == "C5G" else brickwork_circuit(n, 21, 41, 5)
        init = None
        title = {"C5": "C5: 20-qubit Heisenberg chain energy sweep, brickwork synthetic circuit",
                 "C5L": "C5L: C5 from a loaded dense initial state (known-zero skipping off)",
                 "C5G": "C5G: 20-qubit Heisenberg chain, 440 gates from the generic generator (no brick structure)"}[name]
        if name == "C5L":
            rng = np.random.default_rng(55)
            init = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
            init /= np.linalg.norm(init)
        return Workload(name, title, n, gl, pauli=(x, z, w), init=init, batch=64, groups=_groups_of(x), bound="hbm",
                        bound_note="64 x 16 MiB of states per GPU and pass >> 126 MB L2: HBM streaming + FP64 tensor cores")
    raise ValueError(f"unknown workload {name!r}; one of {NAMES}")
