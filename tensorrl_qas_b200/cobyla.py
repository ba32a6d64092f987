"""Native ask/tell COBYLA (tq_cobyla_* in include/tqsim.h) -- optional stand-in for the reference's
`scipy.optimize.minimize(cost, x0, method="COBYLA", options={"maxiter": 1000})` (environments/environment_qulacs.py:436-441).

Under scipy >= 1.16 COBYLA is pure Python (2-12 ms of interpreter time per iteration), far more than
an energy evaluation on the GPU; this optimiser costs microseconds per iteration and, being ask/tell, lets ONE host loop
drive B optimisers against ONE batched launch per round (`minimize_many`).  Its trajectory is not scipy's, so the drop-in
environments use it only when asked (`TQ_OPTIMIZER=native` or `env.optimizer = "native"`)."""
import ctypes

import numpy as np

from . import _lib

STATUS = {0: "running", 1: "rho reached rhoend", 2: "maximum number of function evaluations reached",
          3: "rounding errors are damaging the simplex"}


class NativeCobyla:
    """One optimiser: `ask()` -> point to evaluate (None when finished), `tell(f)`, `result()`."""

    def __init__(self, x0, rhobeg=1.0, rhoend=1e-4, maxfun=1000):
        self._L = _lib.lib()
        x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(-1)
        self.n = x0.shape[0]
        self._h = ctypes.c_void_p()
        rc = self._L.tq_cobyla_create(self.n, x0.ctypes.data_as(_lib.c_dbl_p), float(rhobeg), float(rhoend), int(maxfun),
                                      ctypes.byref(self._h))
        if rc != 0:
            self._h = None
            raise ValueError("tq_cobyla_create: bad arguments (n >= 1, 0 < rhoend <= rhobeg, maxfun >= 1)")
        self._buf = np.empty(self.n, dtype=np.float64)

    def close(self):
        if getattr(self, "_h", None):
            self._L.tq_cobyla_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def ask(self):
        rc = self._L.tq_cobyla_ask(self._h, self._buf.ctypes.data_as(_lib.c_dbl_p))
        if rc < 0:
            raise RuntimeError("tq_cobyla_ask failed")
        return None if rc == 1 else self._buf.copy()

    def tell(self, f):
        """True while the optimiser wants another evaluation."""
        rc = self._L.tq_cobyla_tell(self._h, float(f))
        if rc < 0:
            raise RuntimeError("tq_cobyla_tell without a pending ask")
        return rc == 0

    def result(self):
        x = np.empty(self.n, dtype=np.float64)
        f = ctypes.c_double()
        nfev, status = ctypes.c_int32(), ctypes.c_int32()
        rc = self._L.tq_cobyla_result(self._h, x.ctypes.data_as(_lib.c_dbl_p), ctypes.byref(f), ctypes.byref(nfev),
                                      ctypes.byref(status))
        if rc != 0:
            raise RuntimeError("tq_cobyla_result while an evaluation is pending")
        return {"x": x, "fun": f.value, "nfev": nfev.value, "status": status.value, "success": status.value == 1,
                "message": STATUS.get(status.value, "?")}


def minimize(fun, x0, maxiter=1000, rhobeg=1.0, tol=1e-4):
    """Same call shape and result keys as the reference's use of scipy (x, fun, nfev, success, status, message)."""
    opt = NativeCobyla(x0, rhobeg, tol, maxiter)
    x = opt.ask()
    while x is not None:
        if not opt.tell(fun(x)):
            break
        x = opt.ask()
    res = opt.result()
    opt.close()
    return res


def minimize_many(batch_fun, x0s, maxiter=1000, rhobeg=1.0, tol=1e-4):
    """Lock-step optimisation of B independent problems: every round asks all running optimisers for their next point and
    calls `batch_fun(indices, points)` ONCE (-> one value per point; e.g. `energies_multi`, one launch for B different
    circuits), then tells them.  Returns the list of result dicts and the number of rounds."""
    budgets = list(maxiter) if hasattr(maxiter, "__len__") else [maxiter] * len(x0s)   # one budget, or one per problem
    opts = [NativeCobyla(x0, rhobeg, tol, m) for x0, m in zip(x0s, budgets)]
    running = list(range(len(opts)))
    rounds = 0
    while running:
        pts = [opts[i].ask() for i in running]
        vals = batch_fun(running, pts)
        rounds += 1
        running = [i for i, v in zip(running, vals) if opts[i].tell(v)]
    out = [o.result() for o in opts]
    for o in opts:
        o.close()
    return out, rounds
