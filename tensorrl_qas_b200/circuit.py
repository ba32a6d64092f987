"""Gate lists in the layout of tq_set_circuit (include/tqsim.h) and the reference's tensor encoding of circuits.

Reference behaviour restated here (no qulacs objects are built -- the kernels consume the arrays directly):
  * `decode_state_tensor`  == Parametric_Circuit.construct_ansatz, environments/VQAs/VQE_qulacs.py:12-44
    (noise variants: environments/VQAs/VQE_qulacs_noise.py:31-33,44-54)
"""
import numpy as np

KIND = {"RX": 0, "RY": 1, "RZ": 2, "CNOT": 3, "X": 4, "Y": 5, "Z": 6, "DEPOL1": 7, "DEPOL2": 8}
KIND_NAME = {v: k for k, v in KIND.items()}


class GateList:
    """Gates in application order; rotations own parameter columns in order of appearance (like qulacs'
    ParametricQuantumCircuit, VQE_qulacs.py:36-40,73-74), noise gates own code-slot columns."""

    def __init__(self, n_qubits):
        self.n_qubits = int(n_qubits)
        self.kind, self.q0, self.q1, self.pidx, self.fixed = [], [], [], [], []
        self.n_params = 0
        self.n_slots = 0
        self.initial_angles = []

    def _add(self, kind, q0, q1, pidx, fixed):
        self.kind.append(kind)
        self.q0.append(int(q0))
        self.q1.append(int(q1))
        self.pidx.append(int(pidx))
        self.fixed.append(float(fixed))

    def add_cnot(self, control, target):
        self._add(KIND["CNOT"], control, target, -1, 0.0)

    def add_rotation(self, axis, qubit, theta, parametric=True):
        """axis 0/1/2 = X/Y/Z.  Parametric rotations take the next parameter column; theta is its initial value."""
        if parametric:
            self._add(int(axis), qubit, 0, self.n_params, float(theta))
            self.n_params += 1
            self.initial_angles.append(float(theta))
        else:
            self._add(int(axis), qubit, 0, -1, float(theta))

    def add_pauli(self, name, qubit):
        self._add(KIND[name], qubit, 0, -1, 0.0)

    def add_depol1(self, qubit, p):
        self._add(KIND["DEPOL1"], qubit, 0, self.n_slots, p)
        self.n_slots += 1

    def add_depol2(self, a, b, p):
        self._add(KIND["DEPOL2"], a, b, self.n_slots, p)
        self.n_slots += 1

    def __len__(self):
        return len(self.kind)

    def arrays(self):
        return (np.asarray(self.kind, dtype=np.int32), np.asarray(self.q0, dtype=np.int32),
                np.asarray(self.q1, dtype=np.int32), np.asarray(self.pidx, dtype=np.int32),
                np.asarray(self.fixed, dtype=np.float64))

    def tuples(self):
        return list(zip(self.kind, self.q0, self.q1, self.pidx, self.fixed))

    @property
    def n_unitary(self):
        return sum(1 for k in self.kind if k <= KIND["Z"])

    def count(self, name):
        return sum(1 for k in self.kind if k == KIND[name])


def decode_state_tensor(state, n_qubits, noise=None):
    """(L, n+6, n) float32 circuit encoding -> GateList, in construct_ansatz's order (VQE_qulacs.py:12-44):
    per layer, CNOTs in row-major order of the one-hot block [targ][ctrl], then rotations in row-major order of
    the [axis][qubit] block (all RX by qubit, then RY, then RZ) with angle state[l][n+3+axis][qubit] promoted
    float32 -> float64 exactly.  noise = (p1, p2) appends DepolarizingNoise(q, p1) after each rotation and
    TwoQubitDepolarizingNoise(ctrl, targ, p2) after each CNOT (VQE_qulacs_noise.py:31-33,44-54)."""
    arr = np.asarray(state.detach().cpu().numpy() if hasattr(state, "detach") else state)
    n = int(n_qubits)
    gl = GateList(n)
    # one nonzero() over the whole tensor instead of two per layer: row-major order of (layer, row, column) is exactly
    # the per-layer row-major order the reference walks
    cl, ct, cc = np.nonzero(arr[:, :n, :] == 1)
    rl, ra, rq = np.nonzero(arr[:, n:n + 3, :] == 1)
    angles = arr[rl, n + 3 + ra, rq].astype(np.float64)   # float32 -> float64, exact
    n_c, n_r = len(cl), len(rl)
    # merge the two streams: within a layer all CNOTs come before the rotations (stable sort on 2 * layer + is_rotation)
    order = np.argsort(np.concatenate([2 * cl, 2 * rl + 1]), kind="stable")
    kind = np.concatenate([np.full(n_c, KIND["CNOT"]), ra])[order]
    q0 = np.concatenate([cc, rq])[order]
    q1 = np.concatenate([ct, np.zeros(n_r, dtype=np.int64)])[order]
    pidx = np.concatenate([np.full(n_c, -1), np.arange(n_r)])[order]
    fixed = np.concatenate([np.zeros(n_c), angles])[order]
    if noise is not None:   # every gate is followed by its depolarising channel; noise slots count gates in order
        is_c = kind == KIND["CNOT"]
        total = n_c + n_r
        slots = np.arange(total)
        kind = np.stack([kind, np.where(is_c, KIND["DEPOL2"], KIND["DEPOL1"])], axis=1).reshape(-1)
        q0 = np.repeat(q0, 2)
        q1 = np.repeat(q1, 2)
        pidx = np.stack([pidx, slots], axis=1).reshape(-1)
        fixed = np.stack([fixed, np.where(is_c, noise[1], noise[0])], axis=1).reshape(-1)
        gl.n_slots = total
    gl.kind, gl.q0, gl.q1, gl.pidx, gl.fixed = kind.tolist(), q0.tolist(), q1.tolist(), pidx.tolist(), fixed.tolist()
    gl.n_params = n_r
    gl.initial_angles = angles.tolist()
    return gl


def synthetic_circuit(n_qubits, n_gates, seed, qubits=None):
    """SURVEY.md section 8d generator: gate i is a rotation with probability 0.6 (axis uniform in X/Y/Z, qubit
    uniform, theta ~ U(-pi, pi)) else a CNOT (control uniform, target uniform != control)."""
    rng = np.random.default_rng(seed)
    gl = GateList(n_qubits)
    append_random_gates(gl, n_gates, rng, qubits)
    return gl


def append_random_gates(gl, n_gates, rng, qubits=None):
    qs = list(range(gl.n_qubits)) if qubits is None else list(qubits)
    for _ in range(n_gates):
        if rng.random() < 0.6 or len(qs) < 2:
            gl.add_rotation(int(rng.integers(3)), qs[int(rng.integers(len(qs)))], float(rng.uniform(-np.pi, np.pi)))
        else:
            c = int(rng.integers(len(qs)))
            t = int(rng.integers(len(qs) - 1))
            t = t + 1 if t >= c else t
            gl.add_cnot(qs[c], qs[t])
    return gl


def brickwork_circuit(n_qubits, gates_per_brick, n_agent_gates, seed):
    """C5-shaped synthetic circuit: one staircase layer of n-1 two-qubit bricks on
    (q, q+1), each `gates_per_brick` generator gates restricted to its two qubits (the shape of a transpiled SU(4)
    MPS brick), followed by `n_agent_gates` generator gates on all qubits."""
    rng = np.random.default_rng(seed)
    gl = GateList(n_qubits)
    for q in range(n_qubits - 1):
        append_random_gates(gl, gates_per_brick, rng, (q, q + 1))
    append_random_gates(gl, n_agent_gates, rng)
    return gl


def parameter_batch(gl, batch, seed0=1000):
    """theta_b = theta_0 + U(-0.1, 0.1), one generator per element seeded 1000 + b."""
    base = np.asarray(gl.initial_angles, dtype=np.float64)
    out = np.empty((batch, max(gl.n_params, 1)), dtype=np.float64)
    for b in range(batch):
        rng = np.random.default_rng(seed0 + b)
        out[b, :gl.n_params] = base + rng.uniform(-0.1, 0.1, size=gl.n_params)
    return out
