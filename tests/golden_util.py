"""Access to tests/golden/vqa_golden.npz (made by tests/golden/make_golden.py from the reference's own code + data)."""
import os

import numpy as np

from tensorrl_qas_b200 import loaders
from tensorrl_qas_b200.circuit import GateList, decode_state_tensor

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vqa_golden.npz")
CASES = ("beh2_6q", "h2o_8q", "ch2_8q", "heis_5q")
_cache = {}


def golden():
    if "g" not in _cache:
        _cache["g"] = dict(np.load(GOLDEN, allow_pickle=False))
    return _cache["g"]


class Case:
    def __init__(self, key):
        g = golden()
        self.key = key
        self.n = int(g[f"{key}/n"])
        self.paulis = [str(s) for s in g[f"{key}/paulis"]]
        self.weights = g[f"{key}/weights"]
        self.eig_min = float(g[f"{key}/eig_min"])
        self.h00 = float(g[f"{key}/h00"])
        self.g = {k.split("/", 1)[1]: v for k, v in g.items() if k.startswith(key + "/")}

    def masks(self, reversed_h):
        """Pauli masks for the un-reversed matrix (trainable envs) or the bit-reversed one (fixed envs)."""
        return loaders.pauli_masks(self.paulis, self.n, char0_is_msb=not reversed_h)

    def dense(self, reversed_h):
        """Dense Hamiltonian rebuilt from the Pauli list (the npz matrix equals this sum exactly)."""
        from oracle.np_oracle import pauli_matrix_le
        x, z = self.masks(reversed_h)
        dim = 1 << self.n
        H = np.zeros((dim, dim), dtype=np.complex128)
        for xm, zm, w in zip(x, z, self.weights):
            H += w * pauli_matrix_le(self.n, int(xm), int(zm))
        return H

    def init_circuit(self):
        c = loaders.InitCircuit(self.n)
        for name, q0, q1, th in zip(self.g["init_name"], self.g["init_q0"], self.g["init_q1"], self.g["init_theta"]):
            name = str(name)
            c.ops.append((name, (int(q0), int(q1)) if name == "cx" else (int(q0),), None if name == "cx" else float(th)))
        return c

    def gatelist(self, which, noise=None):
        t = self.g["in_tensor" if which == "in" else "notin_tensor"]
        return decode_state_tensor(t, self.n, noise=noise)
