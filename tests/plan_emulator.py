"""numpy execution of the tile passes the circuit compiler emits (tq_plan_dump), used on CPU to check that a plan
is equivalent to the gate list it was compiled from.  Test-only: mirrors the DevOp semantics documented in
tensorrl_qas_b200/csrc/tq_plan.h on the full vector (tile position p of a pass = physical bit local[p])."""
import numpy as np

from oracle.np_oracle import _mat, apply_1q, apply_cnot

(OP_RX, OP_RY, OP_RZ, OP_RZ_NL, OP_CNOT, OP_CNOT_NL, OP_X, OP_Y, OP_Z, OP_Z_NL, OP_PAULI1, OP_PAULI2, OP_DEPOL1_DM,
 OP_DEPOL2_DM) = range(14)
FLAG_CONJ = 1


def _pauli_both(vec, nbits, n, q, code):
    if code == 0:
        return vec
    m = _mat("XYZ"[code - 1])
    vec = apply_1q(vec, nbits, q, m)
    return apply_1q(vec, nbits, q + n, np.conj(m))


def _depol_dm(vec, nbits, qubits, p):
    """exact channel by definition on the vectorised density matrix (row bits q, column bits q + n)."""
    n = nbits // 2
    two = len(qubits) == 2
    acc = np.zeros_like(vec)
    for code in range(1, 16 if two else 4):
        t = _pauli_both(vec, nbits, n, qubits[0], code & 3)
        if two:
            t = _pauli_both(t, nbits, n, qubits[1], (code >> 2) & 3)
        acc += t
    return (1 - p) * vec + (p / (15.0 if two else 3.0)) * acc


def run_plan(passes, nbits, params, init=None, codes=None):
    vec = np.zeros(1 << nbits, dtype=np.complex128)
    if init is None:
        vec[0] = 1
    else:
        vec[:] = init
    for p in passes:
        loc = p["local"]
        for op, a, b, t, flags, fixed in p["ops"]:
            conj = bool(flags & FLAG_CONJ)
            if op in (OP_RX, OP_RY, OP_RZ, OP_RZ_NL):
                q = a if op == OP_RZ_NL else loc[a]
                theta = params[t] if t >= 0 else fixed
                m = _mat(("RX", "RY", "RZ", "RZ")[op], theta)
                vec = apply_1q(vec, nbits, q, np.conj(m) if conj else m)
            elif op == OP_CNOT:
                vec = apply_cnot(vec, nbits, loc[a], loc[b])
            elif op == OP_CNOT_NL:
                vec = apply_cnot(vec, nbits, a, loc[b])
            elif op in (OP_X, OP_Y, OP_Z, OP_Z_NL):
                q = a if op == OP_Z_NL else loc[a]
                m = _mat({OP_X: "X", OP_Y: "Y", OP_Z: "Z", OP_Z_NL: "Z"}[op])
                vec = apply_1q(vec, nbits, q, np.conj(m) if conj else m)
            elif op == OP_PAULI1:
                c = codes[t] & 3
                if c:
                    vec = apply_1q(vec, nbits, loc[a], _mat("XYZ"[c - 1]))
            elif op == OP_PAULI2:
                ca, cb = codes[t] & 3, (codes[t] >> 2) & 3
                if ca:
                    vec = apply_1q(vec, nbits, loc[a], _mat("XYZ"[ca - 1]))
                if cb:
                    vec = apply_1q(vec, nbits, loc[b], _mat("XYZ"[cb - 1]))
            elif op == OP_DEPOL1_DM:
                assert loc[b] == loc[a] + nbits // 2
                vec = _depol_dm(vec, nbits, [loc[a]], fixed)
            elif op == OP_DEPOL2_DM:
                qa, qb = loc[a & 0xff], loc[(a >> 8) & 0xff]
                assert loc[b & 0xff] == qa + nbits // 2 and loc[(b >> 8) & 0xff] == qb + nbits // 2
                vec = _depol_dm(vec, nbits, [qa, qb], fixed)
            else:
                raise ValueError(op)
    return vec


(W_ROT_X, W_ROT_Y, W_ROT_Z, W_PHASE, W_CX_WW, W_CX_OW, W_X, W_Y, W_Z, W_Z_OUT, W_PAULI, W_DEPOL1, W_DEPOL2) = range(13)


def run_plan_windows(passes, nbits, params, init=None, codes=None):
    """Same as run_plan but from the register-window schedule (what the kernel executes): register bit r of a window
    is tile position wpos[r], i.e. physical bit local[wpos[r]]; qsel is already a physical bit."""
    vec = np.zeros(1 << nbits, dtype=np.complex128)
    if init is None:
        vec[0] = 1
    else:
        vec[:] = init
    for p in passes:
        loc = p["local"]
        k_eff = max(len(loc), 4)
        for w in p["windows"]:
            assert len(w["wpos"]) == 4 and len(set(w["wpos"])) == 4
            assert sorted(w["wpos"] + w["tpos"]) == list(range(k_eff))
            assert len(w["ops"]) <= 256
            phys = [loc[q] if q < len(loc) else None for q in w["wpos"]]
            for code, rb, rb2, qsel, flags, t, fixed in w["ops"]:
                conj = bool(flags & FLAG_CONJ)
                if code in (W_ROT_X, W_ROT_Y, W_ROT_Z, W_PHASE):
                    q = qsel if code == W_PHASE else phys[rb]
                    theta = params[t] if t >= 0 else fixed
                    m = _mat(("RX", "RY", "RZ", "RZ")[code], theta)
                    vec = apply_1q(vec, nbits, q, np.conj(m) if conj else m)
                elif code == W_CX_WW:
                    vec = apply_cnot(vec, nbits, phys[rb], phys[rb2])
                elif code == W_CX_OW:
                    assert qsel not in [x for x in phys if x is not None]
                    vec = apply_cnot(vec, nbits, qsel, phys[rb])
                elif code in (W_X, W_Y, W_Z, W_Z_OUT):
                    q = qsel if code == W_Z_OUT else phys[rb]
                    m = _mat({W_X: "X", W_Y: "Y", W_Z: "Z", W_Z_OUT: "Z"}[code])
                    vec = apply_1q(vec, nbits, q, np.conj(m) if conj else m)
                elif code == W_PAULI:
                    c = (codes[t] >> rb2) & 3
                    if c:
                        vec = apply_1q(vec, nbits, phys[rb], _mat("XYZ"[c - 1]))
                elif code == W_DEPOL1:
                    assert phys[rb2] == phys[rb] + nbits // 2
                    vec = _depol_dm(vec, nbits, [phys[rb]], fixed)
                elif code == W_DEPOL2:
                    qa, qb = phys[rb & 3], phys[(rb >> 2) & 3]
                    assert phys[rb2 & 3] == qa + nbits // 2 and phys[(rb2 >> 2) & 3] == qb + nbits // 2
                    vec = _depol_dm(vec, nbits, [qa, qb], fixed)
                else:
                    raise ValueError(code)
    return vec


def check_invariants(passes, nbits, tile_bits, low_bits):
    """structural checks: tile size, forced low bits, every position operand is a valid tile position"""
    k = min(tile_bits, nbits)
    for p in passes:
        loc = p["local"]
        assert len(loc) == k and loc == sorted(loc) and len(set(loc)) == k
        if nbits > k:
            assert loc[:low_bits] == list(range(low_bits))
        lead = 0
        while lead < k and loc[lead] == lead:
            lead += 1
        assert p["lead"] == lead
        for op, a, b, t, flags, fixed in p["ops"]:
            if op in (OP_RX, OP_RY, OP_RZ, OP_X, OP_Y, OP_Z, OP_PAULI1):
                assert 0 <= a < k
            if op in (OP_CNOT, OP_PAULI2, OP_DEPOL1_DM):
                assert 0 <= a < k and 0 <= b < k and a != b
            if op == OP_CNOT_NL:
                assert 0 <= b < k and a not in loc
            if op in (OP_RZ_NL, OP_Z_NL):
                assert a not in loc and 0 <= a < nbits
