"""qiskit-free readers for the artefacts the TensorRL-QAS environments consume.

  * `load_qpy_circuit`   -- QPY (format version 14, written by qiskit 2.0.0) circuits over {rx, ry, rz, cx}:
                            replaces qiskit.qpy.load at environments/environment_qulacs.py:77-82
  * `InitCircuit.layers` -- ASAP layering == circuit_to_dag(...).layers(), environments/environment_qulacs.py:85-91
  * `load_hamiltonian`   -- dmrg-to-qc/mol_data/*.npz (environments/environment_qulacs.py:100-106)
  * `reverse_qargs`      -- Operator(H).reverse_qargs().to_matrix(), environments/environment_qulacs_TN_notin_agent.py:162
  * `pauli_masks`        -- Pauli strings -> (xmask, zmask) over the little-endian state index
  * `heisenberg_terms`   -- term list of dmrg-to-qc/heisenberg_model.py:21-72 for any n (no dense matrix)
"""
import struct
from dataclasses import dataclass, field

import numpy as np

# ------------------------------------------------------------------------------------------------ QPY ----------
_FILE_HEADER = "!6sBBBBQ"          # magic, qpy version, qiskit major/minor/patch, number of programs
_CIRCUIT_HEADER_V12 = "!H1cHIIQIQI"  # name_size, phase type, phase size, qubits, clbits, metadata, registers, instrs, vars
_REGISTER_V4 = "!1c?IH?"            # type, standalone, size, name_size, in_circuit
_INSTRUCTION_V2 = "!HHHII?HqII"     # name, label, n_params, n_qargs, n_cargs, has_cond, cond_reg, cond_val, ctrl, ctrl_state
_GATES = {"RXGate": "rx", "RYGate": "ry", "RZGate": "rz", "CXGate": "cx"}


@dataclass
class InitCircuit:
    """A circuit over {rx, ry, rz, cx} in qiskit conventions (qubit p = bit p, R_P(t) = exp(-i t/2 P))."""
    n_qubits: int
    ops: list = field(default_factory=list)  # (name, (qubits...), angle or None) in program order
    name: str = ""

    def layers(self):
        """ASAP layers: an op goes to layer max(level of its qubits); depth() = number of layers."""
        level = [0] * self.n_qubits
        out = []
        for name, qs, angle in self.ops:
            l = max(level[q] for q in qs)
            while len(out) <= l:
                out.append([])
            out[l].append((name, qs, angle))
            for q in qs:
                level[q] = l + 1
        return out

    def depth(self):
        return len(self.layers())

    def count(self, name):
        return sum(1 for o in self.ops if o[0] == name)


class _Reader:
    def __init__(self, data):
        self.d, self.o = data, 0

    def unpack(self, fmt):
        size = struct.calcsize(fmt)
        if self.o + size > len(self.d):
            raise ValueError("truncated QPY file")
        vals = struct.unpack_from(fmt, self.d, self.o)
        self.o += size
        return vals

    def take(self, n):
        if self.o + n > len(self.d):
            raise ValueError("truncated QPY file")
        b = self.d[self.o:self.o + n]
        self.o += n
        return b


def load_qpy_circuit(path):
    """Parse the first circuit of a QPY v10-v14 file whose instructions are RX/RY/RZ/CX gates with float parameters."""
    with open(path, "rb") as f:
        return load_qpy_bytes(f.read(), path)


def load_qpy_bytes(data, path="<bytes>"):
    """load_qpy_circuit on the file's contents (`path` only labels error messages)."""
    r = _Reader(bytes(data))
    magic, version, _maj, _min, _pat, n_prog = r.unpack(_FILE_HEADER)
    if magic != b"QISKIT":
        raise ValueError(f"{path}: not a QPY file")
    if version < 10 or version > 14:
        raise ValueError(f"{path}: unsupported QPY version {version} (the shipped circuits are versions 10, 12 and 14)")
    r.take(1)  # symbolic-expression encoding tag
    if n_prog < 1:
        raise ValueError(f"{path}: no circuits")
    if r.take(1) != b"q":
        raise ValueError(f"{path}: first program is not a circuit")
    if version >= 12:
        (name_size, _ptype, phase_size, n_qubits, _n_clbits, meta_size, n_regs, n_instr, n_vars) = r.unpack(_CIRCUIT_HEADER_V12)
    else:  # versions 10/11: same header without the classical-variable count
        (name_size, _ptype, phase_size, n_qubits, _n_clbits, meta_size, n_regs, n_instr) = r.unpack(_CIRCUIT_HEADER_V12[:-1])
        n_vars = 0
    name = r.take(name_size).decode()
    r.take(phase_size)
    r.take(meta_size)
    if n_vars:
        raise ValueError(f"{path}: circuits with classical variables are not supported")
    for _ in range(n_regs):
        _t, _standalone, size, rname_size, _in_circ = r.unpack(_REGISTER_V4)
        r.take(rname_size)
        r.take(8 * size)
    (n_custom,) = r.unpack("!Q")  # custom-operation table precedes the instruction stream
    if n_custom:
        raise ValueError(f"{path}: circuits with custom operations are not supported")
    circ = InitCircuit(n_qubits=n_qubits, name=name)
    for i in range(n_instr):
        nsz, lsz, n_par, n_qargs, n_cargs, has_cond, _creg, _cval, _nctrl, _cstate = r.unpack(_INSTRUCTION_V2)
        gname = r.take(nsz).decode()
        r.take(lsz)
        if has_cond or n_cargs:
            raise ValueError(f"{path}: instruction {i} is conditional / classical")
        if gname not in _GATES:
            raise ValueError(f"{path}: instruction {i} is {gname}; only RX/RY/RZ/CX are supported")
        qs = []
        for _ in range(n_qargs):
            kind, idx = r.unpack("!1cI")
            if kind != b"q":
                raise ValueError(f"{path}: unexpected bit type in instruction {i}")
            qs.append(idx)
        angle = None
        for _ in range(n_par):
            ptype, psize = r.unpack("!1cQ")
            raw = r.take(psize)
            if ptype != b"f" or psize != 8:
                raise ValueError(f"{path}: instruction {i} has a non-float parameter")
            angle = struct.unpack("<d", raw)[0]
        short = _GATES[gname]
        if (short == "cx") != (len(qs) == 2) or (short != "cx" and angle is None):
            raise ValueError(f"{path}: malformed {gname} record")
        circ.ops.append((short, tuple(qs), angle))
    return circ


def init_circuit_gatelist(circ, parametric=False):
    """The init circuit as a qulacs-convention GateList acting on the same qubits: qiskit r_P(t) == qulacs R_P(-t)
    (the sign flip the reference applies at environments/environment_qulacs.py:305-311)."""
    from .circuit import GateList
    gl = GateList(circ.n_qubits)
    axis = {"rx": 0, "ry": 1, "rz": 2}
    for name, qs, angle in circ.ops:
        if name == "cx":
            gl.add_cnot(qs[0], qs[1])
        else:
            gl.add_rotation(axis[name], qs[0], -angle, parametric=parametric)
    return gl


# ------------------------------------------------------------------------------------------ Hamiltonians -------
def pauli_masks(paulis, n_qubits, char0_is_msb=True):
    """Pauli strings -> (xmask, zmask) uint64 arrays over the little-endian state index.

    In the shipped npz files string character p is the p-th kron factor, i.e. bit n-1-p of the matrix index
    (`char0_is_msb=True`, what the un-reversed `hamiltonian` of environments/environment_qulacs.py:106 means).
    With the bit-reversed matrix of the fixed environments (environment_qulacs_TN_notin_agent.py:162) character
    p sits on bit p (`char0_is_msb=False`)."""
    x = np.zeros(len(paulis), dtype=np.uint64)
    z = np.zeros(len(paulis), dtype=np.uint64)
    for t, s in enumerate(paulis):
        s = str(s)
        if len(s) != n_qubits:
            raise ValueError(f"Pauli string {s!r} does not have {n_qubits} characters")
        for p, ch in enumerate(s):
            b = np.uint64(1 << ((n_qubits - 1 - p) if char0_is_msb else p))
            if ch in "XY":
                x[t] |= b
            if ch in "ZY":
                z[t] |= b
            if ch not in "IXYZ":
                raise ValueError(f"bad Pauli character {ch!r}")
    return x, z


def load_hamiltonian(path):
    """npz data contract: returns dict with hamiltonian (complex128), eigvals, weights, paulis (or None)."""
    d = np.load(path, allow_pickle=True)
    out = {"hamiltonian": np.asarray(d["hamiltonian"], dtype=np.complex128), "eigvals": np.asarray(d["eigvals"]),
           "weights": np.asarray(d["weights"]), "paulis": None,
           "energy_shift": d["energy_shift"] if "energy_shift" in d.files else 0}
    if "paulis" in d.files:
        out["paulis"] = [str(s) for s in d["paulis"]]
    return out


def reverse_qargs(H):
    """Bit-reversal permutation of rows and columns: qiskit Operator(H).reverse_qargs().to_matrix()."""
    H = np.asarray(H)
    dim = H.shape[0]
    n = dim.bit_length() - 1
    idx = np.arange(dim)
    rev = np.zeros(dim, dtype=np.int64)
    for b in range(n):
        rev |= ((idx >> b) & 1) << (n - 1 - b)
    return H[np.ix_(rev, rev)]


def heisenberg_terms(n):
    """(paulis, weights) of dmrg-to-qc/heisenberg_model.py:21-72: open chain sum_i XX+YY+ZZ on (i, i+1), then sum_i Z_i."""
    paulis, weights = [], []
    for i in range(n - 1):
        for ch in "XYZ":
            s = ["I"] * n
            s[i] = s[i + 1] = ch
            paulis.append("".join(s))
            weights.append(1.0)
    for i in range(n):
        s = ["I"] * n
        s[i] = "Z"
        paulis.append("".join(s))
        weights.append(1.0)
    return paulis, np.asarray(weights)


def dense_to_pauli(H, tol=0.0):
    """Pauli decomposition of a dense matrix by trace projection (fast Walsh-Hadamard over the flip-mask diagonals).
    Returns (xmask, zmask, coeff) over the matrix's own little-endian index bits; coeff complex128."""
    H = np.asarray(H, dtype=np.complex128)
    dim = H.shape[0]
    n = dim.bit_length() - 1
    idx = np.arange(dim)
    xs, zs, cs = [], [], []
    for x in range(dim):
        f = H[idx ^ x, idx]  # <i^x|H|i> = sum_z c_{x,z} i^{ny} (-1)^{popcount(i&z)}
        if not f.any():
            continue
        g = f.copy()
        h = 1
        while h < dim:  # Walsh-Hadamard transform over i
            g = g.reshape(-1, 2, h)
            g = np.stack([g[:, 0, :] + g[:, 1, :], g[:, 0, :] - g[:, 1, :]], axis=1).reshape(-1)
            h *= 2
        g /= dim
        for z in np.nonzero(np.abs(g) > tol)[0]:
            ny = bin(x & int(z)).count("1") & 3
            xs.append(x)
            zs.append(int(z))
            cs.append(g[z] / (1j ** ny))
    return np.asarray(xs, dtype=np.uint64), np.asarray(zs, dtype=np.uint64), np.asarray(cs, dtype=np.complex128)
