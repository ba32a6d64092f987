"""ctypes front end of oracle/libtq_oracle.so (the C restatement in oracle/tq_oracle.c).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the
product package."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtq_oracle.so")
_lib = None

_i32p = ctypes.POINTER(ctypes.c_int32)
_f64p = ctypes.POINTER(ctypes.c_double)
_u64p = ctypes.POINTER(ctypes.c_uint64)
_u8p = ctypes.POINTER(ctypes.c_uint8)


def build():
    src = os.path.join(_HERE, "tq_oracle.c")
    if os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= os.path.getmtime(src):
        return LIB_PATH
    subprocess.run(["make", "-C", _HERE, "libtq_oracle.so"], check=True, capture_output=True)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = ctypes.CDLL(LIB_PATH)
        circ = [ctypes.c_int, _i32p, _i32p, _i32p, _i32p, _f64p]
        ham = [ctypes.c_int, _f64p, ctypes.c_int, _u64p, _u64p, _f64p, _f64p]
        L.orc_state.argtypes = [ctypes.c_int, _f64p] + circ + [_f64p, _u8p, _f64p]
        L.orc_state.restype = None
        L.orc_energy_batch.argtypes = ([ctypes.c_int, _f64p] + circ + [ctypes.c_int, _f64p, ctypes.c_int, _u8p,
                                                                        ctypes.c_int] + ham + [_f64p, ctypes.c_int])
        L.orc_energy_batch.restype = ctypes.c_int
        L.orc_dm_run.argtypes = [ctypes.c_int, _f64p] + circ + [_f64p, _f64p]
        L.orc_dm_run.restype = None
        L.orc_dm_energy_batch.argtypes = ([ctypes.c_int, _f64p] + circ + [ctypes.c_int, _f64p, ctypes.c_int] + ham +
                                          [_f64p, ctypes.c_int])
        L.orc_dm_energy_batch.restype = ctypes.c_int
        L.orc_expect_dense.argtypes = [ctypes.c_int, _f64p, _f64p]
        L.orc_expect_dense.restype = ctypes.c_double
        L.orc_expect_pauli.argtypes = [ctypes.c_int, _f64p, ctypes.c_int, _u64p, _u64p, _f64p, _f64p]
        L.orc_expect_pauli.restype = ctypes.c_double
        L.orc_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


class _Circ:
    def __init__(self, gl):
        self.kind, self.q0, self.q1, self.pidx, self.fixed = gl.arrays()
        self.G = len(gl)
        self.n = gl.n_qubits

    def args(self):
        return [self.G, _p(self.kind, _i32p), _p(self.q0, _i32p), _p(self.q1, _i32p), _p(self.pidx, _i32p),
                _p(self.fixed, _f64p)]


class _Ham:
    """dense=H (2^n x 2^n complex128) or pauli=(xmask, zmask, coeff)."""

    def __init__(self, dense=None, pauli=None):
        self.kind = 0 if dense is not None else 1
        self.H = None if dense is None else np.ascontiguousarray(dense, dtype=np.complex128)
        if pauli is not None:
            x, z, c = pauli
            self.x = np.ascontiguousarray(x, dtype=np.uint64)
            self.z = np.ascontiguousarray(z, dtype=np.uint64)
            c = np.asarray(c)
            self.cre = np.ascontiguousarray(c.real, dtype=np.float64)
            self.cim = np.ascontiguousarray(c.imag, dtype=np.float64) if np.iscomplexobj(c) else None
            self.T = len(self.x)
        else:
            self.x = self.z = self.cre = self.cim = None
            self.T = 0

    def args(self):
        return [self.kind, _p(self.H, _f64p), self.T, _p(self.x, _u64p), _p(self.z, _u64p), _p(self.cre, _f64p),
                _p(self.cim, _f64p)]


def _prep_params(params):
    p = np.ascontiguousarray(params, dtype=np.float64)
    if p.ndim == 1:
        p = p.reshape(1, -1)
    if p.shape[1] == 0:
        p = np.zeros((p.shape[0], 1))
    return p


def _prep_init(init):
    return None if init is None else np.ascontiguousarray(init, dtype=np.complex128).reshape(-1)


def state(gl, params, init=None, codes=None):
    c = _Circ(gl)
    p = _prep_params(params)[0]
    ini = _prep_init(init)
    cd = None if codes is None else np.ascontiguousarray(codes, dtype=np.uint8).reshape(-1)
    out = np.empty(1 << c.n, dtype=np.complex128)
    lib().orc_state(c.n, _p(ini, _f64p), *c.args(), _p(p, _f64p), _p(cd, _u8p), _p(out, _f64p))
    return out


def energies(gl, params, dense=None, pauli=None, init=None, codes=None, nthreads=0, return_threads=False):
    c = _Circ(gl)
    h = _Ham(dense, pauli)
    p = _prep_params(params)
    ini = _prep_init(init)
    cd = None
    ldc = 0
    if codes is not None:
        cd = np.ascontiguousarray(codes, dtype=np.uint8)
        if cd.ndim == 1:
            cd = cd.reshape(1, -1)
        ldc = cd.shape[1]
    out = np.empty(p.shape[0], dtype=np.float64)
    used = lib().orc_energy_batch(c.n, _p(ini, _f64p), *c.args(), p.shape[0], _p(p, _f64p), p.shape[1],
                                  _p(cd, _u8p), ldc, *h.args(), _p(out, _f64p), nthreads)
    return (out, used) if return_threads else out


def density_matrix(gl, params, init=None):
    """rho as [col][row] (entry rho[r][c] at out[c, r]) to match the r + (c << n) layout."""
    c = _Circ(gl)
    p = _prep_params(params)[0]
    ini = _prep_init(init)
    dim = 1 << c.n
    out = np.empty((dim, dim), dtype=np.complex128)
    lib().orc_dm_run(c.n, _p(ini, _f64p), *c.args(), _p(p, _f64p), _p(out, _f64p))
    return out


def dm_energies(gl, params, dense=None, pauli=None, init=None, nthreads=0):
    c = _Circ(gl)
    h = _Ham(dense, pauli)
    p = _prep_params(params)
    ini = _prep_init(init)
    out = np.empty(p.shape[0], dtype=np.float64)
    lib().orc_dm_energy_batch(c.n, _p(ini, _f64p), *c.args(), p.shape[0], _p(p, _f64p), p.shape[1], *h.args(),
                              _p(out, _f64p), nthreads)
    return out


def expect_dense(psi, H):
    psi = np.ascontiguousarray(psi, dtype=np.complex128)
    H = np.ascontiguousarray(H, dtype=np.complex128)
    n = int(np.log2(psi.shape[0]))
    return lib().orc_expect_dense(n, _p(psi, _f64p), _p(H, _f64p))


def expect_pauli(psi, xmask, zmask, coeff):
    psi = np.ascontiguousarray(psi, dtype=np.complex128)
    h = _Ham(pauli=(xmask, zmask, coeff))
    n = int(np.log2(psi.shape[0]))
    return lib().orc_expect_pauli(n, _p(psi, _f64p), h.T, _p(h.x, _u64p), _p(h.z, _u64p), _p(h.cre, _f64p),
                                  _p(h.cim, _f64p))


def max_threads():
    return lib().orc_max_threads()
