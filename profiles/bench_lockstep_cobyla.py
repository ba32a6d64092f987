"""End-to-end COBYLA rounds of B different environments (BASELINE.json config "256 batched environments").

B BeH2-6q problems -- the shipped MPS circuit with FIXED angles (the reference's fixed environments) plus 12 agent gates
whose angles COBYLA optimises, a different agent circuit per environment.  Compared:
  native lock-step : tensorrl_qas_b200.cobyla.minimize_many (ask all -> ONE tq_energy_multi_host launch -> tell all)
  scipy, serial    : scipy.optimize.minimize(method="COBYLA") per environment, one tq_energy_batch_host call per evaluation
                     (what the reference's loop does with the drop-in shims), on a sample of the environments
Prints one JSON line.  python profiles/bench_lockstep_cobyla.py [--envs 256] [--scipy-sample 16]"""
import argparse
import json
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=256)
    ap.add_argument("--scipy-sample", type=int, default=16)
    args = ap.parse_args()
    import scipy.optimize
    from golden_util import Case
    from tensorrl_qas_b200 import Simulator, cobyla
    from tensorrl_qas_b200.circuit import GateList, append_random_gates
    from tensorrl_qas_b200.simulator import energies_multi

    c = Case("beh2_6q")
    H = c.dense(False)
    base = c.gatelist("in")
    sims, x0s = [], []
    for b in range(args.envs):
        gl = GateList(c.n)
        for kind, q0, q1, pidx, fixed in base.tuples():     # the MPS part: angles fixed
            if kind <= 2:
                gl.add_rotation(kind, q0, base.initial_angles[pidx], parametric=False)
            else:
                gl.add_cnot(q0, q1)
        rng = np.random.default_rng(b)
        while gl.n_params < 12:                              # the agent's part: 12 trainable rotations + some CNOTs
            append_random_gates(gl, 1, rng)
        s = Simulator(c.n)
        s.set_circuit(gl)
        s.set_dense_hamiltonian(H)
        sims.append(s)
        x0s.append(np.asarray(gl.initial_angles, dtype=np.float64))

    def batch(indices, points):
        return energies_multi([sims[i] for i in indices], points)

    cobyla.minimize_many(batch, x0s[:8], maxiter=50)          # warm-up (plans, pinned buffers)
    t0 = time.perf_counter()
    res, rounds = cobyla.minimize_many(batch, x0s, maxiter=1000)
    t_native = time.perf_counter() - t0
    evals_native = sum(r["nfev"] for r in res)

    m = min(args.scipy_sample, args.envs)
    t0 = time.perf_counter()
    sres = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(m):
            sres.append(scipy.optimize.minimize(lambda x, i=i: float(sims[i].energies(x.reshape(1, -1))[0]), x0s[i],
                                                method="COBYLA", options={"maxiter": 1000}))
    t_scipy = time.perf_counter() - t0
    evals_scipy = sum(int(r.nfev) for r in sres)
    d = [res[i]["fun"] - float(sres[i].fun) for i in range(m)]
    print(json.dumps({
        "workload": f"{args.envs} different BeH2-6q environments, 12 trainable angles each, COBYLA to rhoend = 1e-4",
        "native_lockstep": {"seconds": t_native, "rounds": rounds, "evals": evals_native,
                            "evals_per_s": evals_native / t_native, "ms_per_round": 1e3 * t_native / rounds,
                            "envs_per_s": args.envs / t_native},
        "scipy_serial": {"envs": m, "seconds": t_scipy, "evals": evals_scipy, "evals_per_s": evals_scipy / t_scipy,
                         "ms_per_eval": 1e3 * t_scipy / evals_scipy, "envs_per_s": m / t_scipy},
        "speedup_envs_per_s": (args.envs / t_native) / (m / t_scipy),
        "final_energy_native_minus_scipy": {"max": max(d), "min": min(d), "mean": float(np.mean(d))},
    }))


if __name__ == "__main__":
    main()
