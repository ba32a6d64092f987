"""numpy execution of what the circuit compiler emits (tq_plan_dump): fused-block matrix programs, tile passes and
register-window schedules.  Used on CPU to check that a compiled plan is equivalent to the gate list it came from.
Test-only: mirrors the op semantics documented in tensorrl_qas_b200/csrc/tq_plan.h on the full vector (tile
position p of a pass = physical bit local[p]; register bit r of a window = tile position wpos[r])."""
import numpy as np

from oracle.np_oracle import _mat, apply_1q, apply_cnot

OP_U2, OP_U1, OP_D1, OP_D1_NL, OP_CNOT, OP_CNOT_NL, OP_DEPOL1_DM, OP_DEPOL2_DM = range(8)
W_U2, W_U1, W_D1, W_D1_OUT, W_CX_WW, W_CX_OW, W_DEPOL1, W_DEPOL2 = range(8)
MG_RX, MG_RY, MG_RZ, MG_CX, MG_X, MG_Y, MG_Z, MG_PAULI_SLOT = range(8)
FLAG_CONJ, FLAG_SWAP = 1, 2
M_U2, M_SWAPQL, M_CX_OUT, M_CX_RR, M_EXPC, M_EXPD = range(16, 22)


def block_matrix(mat, params, codes):
    """Matrix of a fused block (index bit 0 = block qubit 0), by running its program on the identity."""
    nq = mat["nq"]
    dim = 1 << nq
    M = np.eye(dim, dtype=np.complex128)
    for kind, lq, pidx, fixed in mat["gates"]:
        if kind == MG_CX:
            G = np.zeros((4, 4), dtype=np.complex128)
            for i in range(4):
                G[i ^ ((2 >> lq) if (i >> lq) & 1 else 0), i] = 1  # control bit lq set -> flip the other bit
            M = G @ M
            continue
        if kind == MG_PAULI_SLOT:
            code = (codes[pidx] >> int(fixed)) & 3
            if code == 0:
                continue
            g = _mat("XYZ"[code - 1])
        elif kind <= MG_RZ:
            g = _mat(("RX", "RY", "RZ")[kind], params[pidx] if pidx >= 0 else fixed)
        else:
            g = _mat({MG_X: "X", MG_Y: "Y", MG_Z: "Z"}[kind])
        if nq == 1:
            M = g @ M
        else:
            G = np.kron(g, np.eye(2)) if lq == 1 else np.kron(np.eye(2), g)  # kron's first factor = high bit
            M = G @ M
    if mat["diag"]:
        assert np.abs(M - np.diag(np.diag(M))).max() == 0
    return M


def apply_2q(vec, nbits, q0, q1, M):
    """M's index bit 0 = q0, bit 1 = q1."""
    idx = np.arange(1 << nbits)
    base = idx[((idx >> q0) & 1 == 0) & ((idx >> q1) & 1 == 0)]
    sub = np.stack([vec[base | (((k & 1) << q0) | ((k >> 1) << q1))] for k in range(4)])
    out = M @ sub
    res = vec.copy()
    for k in range(4):
        res[base | (((k & 1) << q0) | ((k >> 1) << q1))] = out[k]
    return res


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


def _init(nbits, init):
    vec = np.zeros(1 << nbits, dtype=np.complex128)
    if init is None:
        vec[0] = 1
    else:
        vec[:] = init
    return vec


def run_plan(plan, nbits, params, init=None, codes=None):
    """Execute the tile-level ops of every pass."""
    vec = _init(nbits, init)
    mats = [block_matrix(m, params, codes) for m in plan["mats"]]
    for p in plan["passes"]:
        loc = p["local"]
        for op, a, b, t, flags, fixed in p["ops"]:
            M = None if t < 0 or op >= OP_CNOT else (np.conj(mats[t]) if flags & FLAG_CONJ else mats[t])
            if op == OP_U2:
                assert loc[a] < loc[b]
                vec = apply_2q(vec, nbits, loc[a], loc[b], M)
            elif op == OP_U1:
                vec = apply_1q(vec, nbits, loc[a], M)
            elif op in (OP_D1, OP_D1_NL):
                vec = apply_1q(vec, nbits, a if op == OP_D1_NL else loc[a], M)
            elif op == OP_CNOT:
                vec = apply_cnot(vec, nbits, loc[a], loc[b])
            elif op == OP_CNOT_NL:
                vec = apply_cnot(vec, nbits, a, loc[b])
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


def run_plan_windows(plan, nbits, params, init=None, codes=None):
    """Execute the register-window schedule (what the kernel runs); qsel is already a physical bit."""
    vec = _init(nbits, init)
    mats = [block_matrix(m, params, codes) for m in plan["mats"]]
    for p in plan["passes"]:
        loc = p["local"]
        k_eff = max(len(loc), 4)
        for w in p["windows"]:
            if w.get("mma"):
                if init is None:
                    # warps the kernel lets idle (dead warp-index bits) must really hold zeros: every amplitude whose
                    # index has such a qubit set is zero on entry (the run started from |0...0>)
                    idx = np.arange(1 << nbits)
                    for i, pos in enumerate(w["w"]):
                        if (w["dead"] >> i) & 1:
                            assert not np.any(vec[((idx >> loc[pos]) & 1) == 1]), "idle warp holds non-zero amplitudes"
                else:
                    pass  # with a loaded state the kernel ignores the dead bits
                vec = _run_mma_window(vec, nbits, loc, w, mats, from_zero=init is None)
                continue
            assert len(w["wpos"]) == 4 and len(set(w["wpos"])) == 4
            assert sorted(w["wpos"] + w["tpos"]) == list(range(k_eff))
            assert len(w["ops"]) <= 32
            phys = [loc[q] if q < len(loc) else None for q in w["wpos"]]
            for code, rb, rb2, qsel, flags, t, fixed in w["ops"]:
                M = None
                if code <= W_D1_OUT:
                    M = np.conj(mats[t]) if flags & FLAG_CONJ else mats[t]
                if code == W_U2:
                    assert rb < rb2
                    if M.shape != (4, 4):
                        raise AssertionError("W_U2 needs a two-qubit matrix")
                    if flags & FLAG_SWAP:  # the block's qubit 0 sits on rb2
                        vec = apply_2q(vec, nbits, phys[rb2], phys[rb], M)
                    else:
                        vec = apply_2q(vec, nbits, phys[rb], phys[rb2], M)
                elif code in (W_U1, W_D1):
                    vec = apply_1q(vec, nbits, phys[rb], M)
                elif code == W_D1_OUT:
                    assert qsel not in [x for x in phys if x is not None]
                    vec = apply_1q(vec, nbits, qsel, M)
                elif code == W_CX_WW:
                    vec = apply_cnot(vec, nbits, phys[rb], phys[rb2])
                elif code == W_CX_OW:
                    assert qsel not in [x for x in phys if x is not None]
                    vec = apply_cnot(vec, nbits, qsel, phys[rb])
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


def _run_mma_window(vec, nbits, loc, w, mats, from_zero=False):
    """Tensor-core window (tq_plan.h "DMMA windows"): ops act on (QL, register bit) pairs; M_SWAPQL only relabels."""
    k = len(loc)
    assert k >= 9 and len(w["r"]) == 5 and len(w["g"]) == 3 and len(w["w"]) == k - 9
    assert sorted(w["r"] + [w["ql"]] + w["g"] + w["w"]) == list(range(k)), "entry layout is not a permutation"
    assert len(w["ops"]) <= 32
    r, ql = list(w["r"]), w["ql"]
    fixed_pos = set(w["g"] + w["w"])
    for code, rb, rb2, qsel, flags, t, fixed in w["ops"]:
        inside = [loc[q] for q in r] + [loc[ql]]
        if code == M_U2:
            M = mats[t]
            pq, px = loc[ql], loc[r[rb]]
            if from_zero and rb2 <= 3:
                # flags bits 1..5: register bits the kernel treats as all-zero before this block
                idx = np.arange(1 << nbits)
                for bit_r in range(5):
                    if (flags >> (1 + bit_r)) & 1:
                        assert not np.any(vec[((idx >> loc[r[bit_r]]) & 1) == 1]), "dead register bit holds amplitudes"
            if rb2 == 0:      # 4x4, index bit 0 = QL
                assert M.shape == (4, 4)
                vec = apply_2q(vec, nbits, pq, px, M)
            elif rb2 == 1:    # 4x4, index bit 0 = RX
                assert M.shape == (4, 4)
                vec = apply_2q(vec, nbits, px, pq, M)
            elif rb2 == 2:
                assert M.shape == (2, 2)
                vec = apply_1q(vec, nbits, px, M)
            elif rb2 == 3:
                assert M.shape == (2, 2)
                vec = apply_1q(vec, nbits, pq, M)
            else:             # diagonal block on a bit outside the window
                assert rb2 == 4 and M.shape == (2, 2) and M[0, 1] == 0 and M[1, 0] == 0 and qsel not in inside
                # the phase is part of the B fragment, which all rows (lane bits 2..4) of a DMMA share
                assert qsel not in [loc[q] for q in w["g"]], "diagonal block on a lane-group qubit"

                vec = apply_1q(vec, nbits, qsel, M)
            if rb != 0:       # results land in adjacent registers: register bits 0 and rb trade qubits
                r[0], r[rb] = r[rb], r[0]
            if flags & 1:     # output roles exchanged (row-permuted matrix): QL <-> register bit 0
                assert rb2 <= 3
                r[0], ql = ql, r[0]
        elif code == M_SWAPQL:
            r[rb], ql = ql, r[rb]
        elif code == M_CX_OUT:
            assert qsel not in inside
            vec = apply_cnot(vec, nbits, qsel, loc[r[rb]])
        else:
            raise ValueError(code)
    assert r == w["rout"] and ql == w["qlout"], "exit layout does not match the swaps"
    assert not (set(r) | {ql}) & fixed_pos
    if not w["ops"]:
        assert w["flags"] & 1, "layout-only windows are read-only"
    return vec


def check_invariants(plan, nbits, tile_bits, low_bits):
    """structural checks: tile size, forced low bits, every position operand is a valid tile position"""
    k = min(tile_bits, nbits)
    for p in plan["passes"]:
        loc = p["local"]
        assert len(loc) == k and loc == sorted(loc) and len(set(loc)) == k
        if nbits > k:
            assert loc[:low_bits] == list(range(low_bits))
        lead = 0
        while lead < k and loc[lead] == lead:
            lead += 1
        assert p["lead"] == lead
        n_real = sum(1 for w in p["windows"] for o in w["ops"] if o[0] != M_SWAPQL)
        assert n_real == len(p["ops"])
        for op, a, b, t, flags, fixed in p["ops"]:
            if op in (OP_U1, OP_D1):
                assert 0 <= a < k
            if op in (OP_U2, OP_CNOT, OP_DEPOL1_DM):
                assert 0 <= a < k and 0 <= b < k and a != b
            if op == OP_CNOT_NL:
                assert 0 <= b < k and a not in loc
            if op == OP_D1_NL:
                assert a not in loc and 0 <= a < nbits
            if op <= OP_D1_NL:
                assert 0 <= t < len(plan["mats"])
