"""Native ask/tell COBYLA (tq_cobyla_*, tensorrl_qas_b200/cobyla.py): host-only, runs without a GPU.  It is NOT pinned to
scipy's trajectory (scipy >= 1.16 ships PRIMA's COBYLA); what is checked: convergence to known minima, the evaluation
budget, the ask/tell protocol, determinism, lock-step batching, a VQE cost against scipy's result, and the environment
switch."""
import warnings

import numpy as np
import pytest
import scipy.optimize

from oracle import c_oracle
from golden_util import Case
from tensorrl_qas_b200 import cobyla

import env_fixture as fx
import test_env as te


def test_convex_quadratics_reach_the_minimum(built_lib):
    rng = np.random.default_rng(0)
    for n in (1, 2, 5, 12, 40):
        A = rng.normal(size=(n, n))
        Q = A @ A.T / n + np.eye(n)
        b = rng.normal(size=n)
        xs = np.linalg.solve(Q, b)
        calls = []

        def f(x):
            calls.append(1)
            return 0.5 * x @ Q @ x - b @ x

        res = cobyla.minimize(f, np.zeros(n), maxiter=4000)
        assert res["nfev"] == len(calls) <= 4000
        assert res["success"] and res["status"] == 1
        assert res["fun"] - f(xs) < 1e-6 * max(1.0, abs(f(xs)))
        assert np.abs(res["x"] - xs).max() < 5e-3
        assert abs(f(res["x"]) - res["fun"]) < 1e-15          # the reported value belongs to the reported point


def test_budget_protocol_and_determinism(built_lib):
    def f(x):
        return float(np.sum(np.cos(x)) + 0.1 * np.sum(x * x))

    x0 = np.linspace(-1, 1, 7)
    r = cobyla.minimize(f, x0, maxiter=9)
    assert r["nfev"] == 9 and r["status"] == 2 and not r["success"]
    assert r["fun"] <= f(x0)                                   # never worse than the best point evaluated
    r1, r2 = cobyla.minimize(f, x0), cobyla.minimize(f, x0)
    assert r1["nfev"] == r2["nfev"] and np.array_equal(r1["x"], r2["x"]) and r1["fun"] == r2["fun"]
    opt = cobyla.NativeCobyla(x0)
    assert np.array_equal(opt.ask(), x0)                       # the first point is x0
    with pytest.raises(RuntimeError):
        opt.result()                                           # an evaluation is pending
    assert opt.tell(f(x0))
    x1 = opt.ask()
    assert np.count_nonzero(x1 != x0) == 1 and abs((x1 - x0).sum() - 1.0) < 1e-15   # first simplex vertex: x0 + rhobeg e_0
    with pytest.raises(ValueError):
        cobyla.NativeCobyla(np.zeros(0))
    with pytest.raises(ValueError):
        cobyla.NativeCobyla(np.zeros(3), rhobeg=1e-6, rhoend=1e-4)


def test_lockstep_batches_equal_single_runs(built_lib):
    rng = np.random.default_rng(3)
    shifts = [rng.normal(size=n) for n in (3, 6, 6, 10)]

    def f(i, x):
        return float(np.sum((x - shifts[i]) ** 2) + np.sum(np.sin(x)))

    singles = [cobyla.minimize(lambda x, i=i: f(i, x), np.zeros(len(s))) for i, s in enumerate(shifts)]
    sizes = []

    def batch(indices, points):
        sizes.append(len(indices))
        return [f(i, x) for i, x in zip(indices, points)]

    many, rounds = cobyla.minimize_many(batch, [np.zeros(len(s)) for s in shifts])
    assert rounds == max(r["nfev"] for r in singles) and sizes[0] == 4 and sizes[-1] >= 1
    for a, b in zip(singles, many):
        assert a["nfev"] == b["nfev"] and np.array_equal(a["x"], b["x"]) and a["fun"] == b["fun"]


def test_vqe_cost_against_scipy(built_lib):
    """H2O-8q golden circuit (reference code + data), perturbed angles as the start.  Measured in round 1 (1000
    evaluations): all 138 angles -- scipy (PRIMA) -73.2902, native -73.2768 from a start of -70.18 (PRIMA is the better
    optimiser per evaluation in high dimension; the native one wins on host time per iteration); the last 12 angles --
    scipy -72.5508, native -72.5514."""
    c = Case("h2o_8q")
    gl = c.gatelist("in")
    x, z = c.masks(False)
    ham = (x, z, c.weights)
    x0 = np.asarray(c.g["in_X"][0], dtype=np.float64)

    def cost(p):
        return float(c_oracle.energies(gl, np.asarray(p)[None, :], pauli=ham, nthreads=1)[0])

    idx = np.arange(len(x0) - 12, len(x0))

    def cost12(p):
        q = x0.copy()
        q[idx] = p
        return cost(q)

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        s = scipy.optimize.minimize(cost12, x0[idx], method="COBYLA", options={"maxiter": 1000})
    r = cobyla.minimize(cost12, x0[idx], maxiter=1000)
    assert r["nfev"] <= 1000 and r["fun"] <= cost(x0) + 1e-12
    assert r["fun"] <= s.fun + 2e-3                            # the fixed environments' regime (a dozen angles)
    full = cobyla.minimize(cost, x0, maxiter=1000)
    assert full["nfev"] <= 1000
    assert cost(x0) - full["fun"] >= 0.99 * (cost(x0) - (-73.29015961931559))   # 99 % of scipy's gain on all 138 angles
    assert full["fun"] >= c.eig_min - 1e-9                     # variational bound


@pytest.mark.parametrize("key", ["fixed_beh2", "fixed_h2o8"])
def test_environment_switch(key, tmp_path, monkeypatch):
    """env.optimizer = "native" (or TQ_OPTIMIZER=native) routes CircuitEnv.scipy_optim through the native optimiser; scipy
    stays the default.  On the reference's fixed-environment episodes (a few angles per step) it lands on the energies of
    the scipy-driven golden episodes to 1e-5 Ha with the same masks, gate placement and done flags (measured: |dE| <= 1e-6,
    19-150 evaluations per step against scipy's 20-147)."""
    ep, env, table = te._make_env(key, tmp_path, monkeypatch, "oracle")
    assert getattr(env, "optimizer", None) in (None, "scipy")
    env.optimizer = "native"
    env.reset()
    d = ep.d
    optimised = 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(ep.n_steps):
            assert [int(a) for a in env.illegal_action_new()] == ep.illegal(i)
            obs, reward, done = env.step(list(table[int(d["action"][i])]))
            n = env.num_qubits
            assert np.array_equal(env.state.numpy()[:, :n + 3], d["state"][i][:, :n + 3])
            assert 1 <= int(env.nfev) <= env.global_iters
            optimised += int(env.nfev) > 1
            assert abs(float(env.energy) - float(d["energy"][i])) < 1e-5, i
            assert float(env.energy) >= float(env.min_eig) - 1e-9
            assert done == int(d["done"][i])
            if done:
                break
    assert optimised >= 5


def test_pinned_trajectories(built_lib):
    """Regression pins of the optimiser's own path (round 1 values; plain double arithmetic, no fused multiply-adds on the
    host side): evaluation counts are exact, values to 1e-12."""
    def quad(x):
        return float(np.sum((x - np.arange(len(x))) ** 2) + 0.5 * x[0] * x[1])

    r = cobyla.minimize(quad, np.zeros(12))
    assert (r["nfev"], r["status"]) == (293, 1)
    assert abs(r["fun"] - (-0.06666665549933619)) < 1e-12
    assert abs(r["x"][0] - (-0.2666608162230507)) < 1e-9 and abs(r["x"][-1] - 10.999971975379118) < 1e-9

    def wells(x):
        return float(np.sum(np.cos(x)) + 0.1 * np.sum(x * x))

    r = cobyla.minimize(wells, np.linspace(-1, 1, 7))
    assert (r["nfev"], r["status"]) == (150, 1)
    assert abs(r["fun"] - (-1.2662883372511278)) < 1e-12
