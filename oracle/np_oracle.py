"""numpy restatement of the qulacs entry points the TensorRL-QAS hot path uses.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED against qulacs itself (qulacs is an unpinned third-party wheel, reference requirements.txt:2, that
cannot be imported or installed in the build container; the reference has no tests).  The semantics restated here
are qulacs' public ones, cross-checked against the reference's own sign/mirror handling (SURVEY.md section 8c /
appendix A):

  * little-endian basis: qubit k <-> bit k of the amplitude index
  * RX/RY/RZ(theta) = exp(+i theta/2 P)     -- reference compensates with `-gate_param`,
                                               environments/environment_qulacs.py:305-311
  * CNOT(control, target)                    -- environments/VQAs/VQE_qulacs.py:25
  * DepolarizingNoise(q, p): X, Y, Z each with probability p/3 on a pure state, one draw per update
    TwoQubitDepolarizingNoise(a, b, p): each of the 15 non-identity two-qubit Paulis with probability p/15
    (environments/VQAs/VQE_qulacs_noise.py:31-33,44-54)

The 11 entry points (inner boundary) are exposed with qulacs' names so that the reference's
own environments/VQAs/VQE_qulacs*.py can be imported and run verbatim on top of this module
(`install_as_qulacs()`); that is how tests/golden/make_golden.py produces the committed golden vectors.

Nothing under tensorrl_qas_b200/ imports this file.
"""
import sys
import types

import numpy as np

_rng = np.random.default_rng(0)
noise_log = []  # Pauli code drawn by every probabilistic gate application, in order (0 = identity)


def seed(s):
    """Seed the explicit RNG used by the probabilistic noise gates (qulacs' own RNG is unseeded in the reference)."""
    global _rng
    _rng = np.random.default_rng(s)


def _mat(kind, theta=0.0):
    c, s = np.cos(0.5 * theta), np.sin(0.5 * theta)
    if kind == "RX":
        return np.array([[c, 1j * s], [1j * s, c]], dtype=np.complex128)
    if kind == "RY":
        return np.array([[c, s], [-s, c]], dtype=np.complex128)
    if kind == "RZ":
        return np.array([[c + 1j * s, 0], [0, c - 1j * s]], dtype=np.complex128)
    if kind == "X":
        return np.array([[0, 1], [1, 0]], dtype=np.complex128)
    if kind == "Y":
        return np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
    if kind == "Z":
        return np.array([[1, 0], [0, -1]], dtype=np.complex128)
    if kind == "I":
        return np.eye(2, dtype=np.complex128)
    raise ValueError(kind)


def apply_1q(vec, nbits, q, m):
    """vec <- (m on bit q) vec, little-endian.  Returns a new array."""
    v = vec.reshape(1 << (nbits - 1 - q), 2, 1 << q)
    out = np.empty_like(v)
    out[:, 0, :] = m[0, 0] * v[:, 0, :] + m[0, 1] * v[:, 1, :]
    out[:, 1, :] = m[1, 0] * v[:, 0, :] + m[1, 1] * v[:, 1, :]
    return out.reshape(-1)


def apply_cnot(vec, nbits, ctrl, targ):
    idx = np.arange(1 << nbits)
    src = np.where((idx >> ctrl) & 1, idx ^ (1 << targ), idx)
    return vec[src]


class QuantumState:
    """qulacs.QuantumState: |0...0> on construction (environments/VQAs/VQE_qulacs.py:81)."""

    def __init__(self, n):
        self.n = int(n)
        self.vec = np.zeros(1 << self.n, dtype=np.complex128)
        self.vec[0] = 1.0

    def set_zero_state(self):
        self.vec[:] = 0
        self.vec[0] = 1.0

    def load(self, v):
        """environments/VQAs/VQE_qulacs_TN_notin_RL.py:83"""
        v = np.asarray(v, dtype=np.complex128).reshape(-1)
        if v.shape[0] != self.vec.shape[0]:
            raise ValueError("state dimension mismatch")
        self.vec = v.copy()

    def get_vector(self):
        return self.vec.copy()

    def get_qubit_count(self):
        return self.n


class _Gate:
    def __init__(self, kind, q0, q1=-1, value=0.0):
        self.kind, self.q0, self.q1, self.value = kind, int(q0), int(q1), float(value)

    def update_quantum_state(self, state):
        n, k = state.n, self.kind
        if k == "CNOT":
            state.vec = apply_cnot(state.vec, n, self.q0, self.q1)
        elif k == "DEPOL1":
            u = _rng.random()
            p = self.value
            code = 0
            if u < p:
                code = 1 + min(int(u / (p / 3.0)), 2)
                state.vec = apply_1q(state.vec, n, self.q0, _mat("XYZ"[code - 1]))
            noise_log.append(code)
        elif k == "DEPOL2":
            u = _rng.random()
            p = self.value
            code = 0
            if u < p:
                code = 1 + min(int(u / (p / 15.0)), 14)  # 1..15 = pa + 4 pb
                pa, pb = code & 3, code >> 2
                if pa:
                    state.vec = apply_1q(state.vec, n, self.q0, _mat("XYZ"[pa - 1]))
                if pb:
                    state.vec = apply_1q(state.vec, n, self.q1, _mat("XYZ"[pb - 1]))
            noise_log.append(code)
        else:
            state.vec = apply_1q(state.vec, n, self.q0, _mat(k, self.value))


def CNOT(control, target):
    return _Gate("CNOT", control, target)


def RX(q, theta):
    return _Gate("RX", q, value=theta)


def RY(q, theta):
    return _Gate("RY", q, value=theta)


def RZ(q, theta):
    return _Gate("RZ", q, value=theta)


def X(q):
    return _Gate("X", q)


def Y(q):
    return _Gate("Y", q)


def Z(q):
    return _Gate("Z", q)


def DepolarizingNoise(q, p):
    return _Gate("DEPOL1", q, value=p)


def TwoQubitDepolarizingNoise(a, b, p):
    return _Gate("DEPOL2", a, b, value=p)


class QuantumCircuit:
    def __init__(self, n):
        self.n = int(n)
        self.gates = []

    def add_gate(self, g):
        self.gates.append(g)

    def get_gate_count(self):
        return len(self.gates)

    def update_quantum_state(self, state):
        for g in self.gates:
            g.update_quantum_state(state)


class ParametricQuantumCircuit(QuantumCircuit):
    """Parameter index = order of add_parametric_* calls (environments/VQAs/VQE_qulacs.py:36-40,73-74)."""

    def __init__(self, n):
        super().__init__(n)
        self.param_gates = []

    def _add_param(self, kind, q, theta):
        g = _Gate(kind, q, value=theta)
        self.gates.append(g)
        self.param_gates.append(g)

    def add_parametric_RX_gate(self, q, theta):
        self._add_param("RX", q, theta)

    def add_parametric_RY_gate(self, q, theta):
        self._add_param("RY", q, theta)

    def add_parametric_RZ_gate(self, q, theta):
        self._add_param("RZ", q, theta)

    def get_parameter_count(self):
        return len(self.param_gates)

    def get_parameter(self, j):
        return self.param_gates[int(j)].value

    def set_parameter(self, j, v):
        self.param_gates[int(j)].value = float(v)


def install_as_qulacs():
    """Register this module as `qulacs` / `qulacs.gate` so the reference's VQA modules import unchanged."""
    me = sys.modules[__name__]
    q = types.ModuleType("qulacs")
    for name in ("QuantumState", "QuantumCircuit", "ParametricQuantumCircuit"):
        setattr(q, name, getattr(me, name))
    g = types.ModuleType("qulacs.gate")
    names = ("CNOT", "RX", "RY", "RZ", "X", "Y", "Z", "DepolarizingNoise", "TwoQubitDepolarizingNoise")
    for name in names:
        setattr(g, name, getattr(me, name))
    g.__all__ = list(names)
    q.gate = g
    sys.modules["qulacs"] = q
    sys.modules["qulacs.gate"] = g
    return q


# ---------------------------------------------------------------------------------------------------------------
# functional helpers on plain gate lists (same argument meaning as tq_set_circuit, include/tqsim.h)
# ---------------------------------------------------------------------------------------------------------------
KIND_NAMES = ("RX", "RY", "RZ", "CNOT", "X", "Y", "Z", "DEPOL1", "DEPOL2")


def run_circuit(n, gates, params, init=None, codes=None):
    """gates: iterable of (kind:int, q0, q1, param_idx, fixed).  Returns the final 2^n complex128 vector."""
    vec = np.zeros(1 << n, dtype=np.complex128)
    if init is None:
        vec[0] = 1.0
    else:
        vec[:] = np.asarray(init, dtype=np.complex128)
    for kind, q0, q1, pidx, fixed in gates:
        name = KIND_NAMES[kind]
        if name == "CNOT":
            vec = apply_cnot(vec, n, q0, q1)
        elif name == "DEPOL1":
            if codes is not None and pidx >= 0 and codes[pidx] & 3:
                vec = apply_1q(vec, n, q0, _mat("XYZ"[(codes[pidx] & 3) - 1]))
        elif name == "DEPOL2":
            if codes is not None and pidx >= 0:
                pa, pb = codes[pidx] & 3, (codes[pidx] >> 2) & 3
                if pa:
                    vec = apply_1q(vec, n, q0, _mat("XYZ"[pa - 1]))
                if pb:
                    vec = apply_1q(vec, n, q1, _mat("XYZ"[pb - 1]))
        elif name in ("RX", "RY", "RZ"):
            theta = float(params[pidx]) if pidx >= 0 else float(fixed)
            vec = apply_1q(vec, n, q0, _mat(name, theta))
        else:
            vec = apply_1q(vec, n, q0, _mat(name))
    return vec


def expect_dense(psi, H):
    """Verbatim form of environments/VQAs/VQE_qulacs.py:85."""
    return (np.conj(psi).T @ H @ psi).real


def pauli_matrix_le(n, xmask, zmask):
    """Dense matrix of one Pauli string given as little-endian bit masks (tests only, small n)."""
    mats = []
    for q in range(n - 1, -1, -1):  # kron's first factor is the most significant bit
        xb, zb = (xmask >> q) & 1, (zmask >> q) & 1
        mats.append(_mat("IXZY"[xb + 2 * zb]))
    out = np.array([[1.0 + 0j]])
    for m in mats:
        out = np.kron(out, m)
    return out


def expect_pauli(psi, xmask, zmask, coeff):
    """sum_t coeff_t <psi|P_t|psi> with P|i> = i^{ny} (-1)^{popcount(i & z)} |i ^ x>."""
    dim = psi.shape[0]
    idx = np.arange(dim, dtype=np.uint64)
    total = 0.0
    for x, z, w in zip(xmask, zmask, coeff):
        x, z = int(x), int(z)
        ny = bin(x & z).count("1") & 3
        par = np.zeros(dim, dtype=np.int64)
        t = idx & np.uint64(z)
        while t.any():
            par ^= (t & np.uint64(1)).astype(np.int64)
            t >>= np.uint64(1)
        sign = 1.0 - 2.0 * par
        val = np.sum(np.conj(psi[(idx ^ np.uint64(x)).astype(np.int64)]) * sign * psi)
        total += (w * (1j ** ny) * val).real
    return float(total)
