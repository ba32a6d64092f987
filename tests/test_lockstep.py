"""Lock-step multi-environment evaluation (tensorrl_qas_b200/lockstep.py, tq_energy_multi_host)."""
import importlib
import warnings

import numpy as np
import pytest
import torch

import env_fixture as fx


def test_lockstep_group_batches_rounds_and_handles_uneven_workers():
    """Workers with different numbers of evaluations: every round holds one request per still-running worker, results
    go back to the right worker, a finished worker no longer blocks the others."""
    from tensorrl_qas_b200.VQAs import _backend
    from tensorrl_qas_b200.lockstep import run_lockstep
    sizes = []

    def evaluate_round(items):
        sizes.append(len(items))
        return [float(np.sum(p)) + 1000.0 * sim for sim, p, _ in items]

    def make_task(i, n_evals):
        def task():
            out = []
            for k in range(n_evals):
                out.append(_backend.evaluate(i, np.array([k, 0.5 * i])))   # `sim` is just a tag here
            return out
        return task

    counts = [5, 2, 7, 1]
    results, group = run_lockstep([make_task(i, c) for i, c in enumerate(counts)], evaluate_round=evaluate_round)
    for i, c in enumerate(counts):
        assert results[i] == [k + 0.5 * i + 1000.0 * i for k in range(c)]
    assert sizes == [4, 3, 2, 2, 2, 1, 1] and group.evaluations == sum(counts) and group.rounds == 7
    # outside a lock-step thread the hook is inert
    assert getattr(_backend._ctx, "group", None) is None


def test_lockstep_propagates_errors_without_deadlock():
    from tensorrl_qas_b200.VQAs import _backend
    from tensorrl_qas_b200.lockstep import run_lockstep

    def bad_round(items):
        raise RuntimeError("device lost")

    with pytest.raises(RuntimeError, match="device lost"):
        run_lockstep([lambda: _backend.evaluate(0, np.zeros(1)), lambda: _backend.evaluate(1, np.zeros(1))],
                     evaluate_round=bad_round)

    def task_raises():
        raise ValueError("boom")

    with pytest.raises(ValueError, match="boom"):
        run_lockstep([task_raises, lambda: _backend.evaluate(1, np.ones(1))], evaluate_round=lambda it: [0.0] * len(it))


@pytest.mark.gpu
def test_multi_problem_launch_matches_individual_calls(built_lib, oracle):
    """tq_energy_multi_host: B different circuits / Hamiltonians / initial states in one launch, bit-identical to B
    separate calls; both kernel families (n = 6: FP64-pipe windows, n = 10: tensor-core windows)."""
    from tensorrl_qas_b200 import Simulator, loaders
    from tensorrl_qas_b200.circuit import parameter_batch, synthetic_circuit
    from tensorrl_qas_b200.simulator import energies_multi
    for n in (6, 10):
        sims, params, want = [], [], []
        for i in range(9):
            gl = synthetic_circuit(n, 10 + 7 * i, 100 + i)
            rng = np.random.default_rng(i)
            A = rng.normal(size=(1 << n, 1 << n))
            H = A + A.T
            s = Simulator(n)
            s.set_circuit(gl)
            s.set_dense_hamiltonian(H)
            init = None
            if i % 3 == 1:
                init = oracle.state(synthetic_circuit(n, 12, 7 + i), parameter_batch(synthetic_circuit(n, 12, 7 + i), 1)[0])
                s.set_init_state(init)
            p = parameter_batch(gl, 1)[0]
            sims.append(s)
            params.append(p.astype(np.float32) if i % 2 else p)      # float32 angle vectors are promoted exactly
            want.append(oracle.energies(gl, np.asarray(params[-1], dtype=np.float64).reshape(1, -1), dense=H, init=init)[0])
        got = energies_multi(sims, params)
        single = np.array([s.energies(np.asarray(p, dtype=np.float64).reshape(1, -1))[0] for s, p in zip(sims, params)])
        assert np.array_equal(got, single)
        assert np.abs(got - np.array(want)).max() < 1e-10
        for s in sims:
            s.close()


def test_group_optimise_runs_whole_problems_in_one_host_loop(built_lib):
    """Native-optimiser lock-step: each worker hands its WHOLE problem to the group (one thread hand-over per
    optimisation, not per evaluation); the coordinator probes the cost closures for their requests and evaluates every
    round as one batch.  Results equal stand-alone runs bit for bit; workers that only need single evaluations, or finish
    early, mix in."""
    from tensorrl_qas_b200 import cobyla
    from tensorrl_qas_b200.VQAs import _backend
    from tensorrl_qas_b200.lockstep import run_lockstep
    shifts = [np.linspace(-1, 1, n) * (i + 1) for i, n in enumerate((3, 5, 8))]
    sizes = []

    def f(i, p):
        return float(np.sum((p - shifts[i]) ** 2) + np.sum(np.cos(p)))

    def evaluate_round(items):
        sizes.append(len(items))
        return [f(tag, p) for tag, p, _ in items]

    def make_task(i):
        def task():
            first = _backend.evaluate(i, np.zeros(len(shifts[i])))               # a plain evaluation first (get_energy)
            res = _backend._ctx.group.optimise(i, lambda x: _backend.evaluate(i, x), np.zeros(len(shifts[i])), 1000)
            last = _backend.evaluate(i, res["x"])                               # ... and one after (the step's energy)
            return first, res, last
        return task

    def lone_task():
        return [_backend.evaluate(0, np.full(3, 0.1 * k)) for k in range(4)]

    results, group = run_lockstep([make_task(i) for i in range(3)] + [lone_task], evaluate_round=evaluate_round)
    for i in range(3):
        first, res, last = results[i]
        single = cobyla.minimize(lambda x: f(i, x), np.zeros(len(shifts[i])))
        assert first == f(i, np.zeros(len(shifts[i])))
        assert res["nfev"] == single["nfev"] and np.array_equal(res["x"], single["x"]) and res["fun"] == single["fun"]
        assert last == res["fun"]
    assert results[3] == [f(0, np.full(3, 0.1 * k)) for k in range(4)]
    assert max(sizes) >= 3 and group.evaluations == sum(sizes)

    # a cost that post-processes the energy cannot be probed: loud failure, no dead-lock
    def bad_task():
        return _backend._ctx.group.optimise(0, lambda x: _backend.evaluate(0, x) + 1.0, np.zeros(2), 50)

    with pytest.raises(RuntimeError, match="exactly one energy evaluation"):
        run_lockstep([bad_task], evaluate_round=evaluate_round)


@pytest.mark.gpu
@pytest.mark.parametrize("optimizer", ["scipy", "native"])
@pytest.mark.parametrize("module_name,key", [("environment_qulacs_TN_notin_agent", "fixed_beh2"),
                                              ("environment_qulacs", "trainable_beh2"),
                                              ("environment_qulacs_noise", "noise_trainable_h2o8")])
def test_lockstep_envs_reproduce_serial_trajectories(module_name, key, optimizer, tmp_path, monkeypatch):
    """B environments stepped in lock-step (one launch per COBYLA round) follow exactly the trajectories of the same
    environments stepped one after the other -- with scipy's COBYLA and with the library's own (TQ_OPTIMIZER=native)."""
    monkeypatch.setenv("TQ_OPTIMIZER", optimizer)
    from tensorrl_qas_b200.VQAs import _backend
    from tensorrl_qas_b200.environments.utils import utils
    from tensorrl_qas_b200.lockstep import LockstepEnvs
    ep = fx.Episode(key)
    fx.materialize(str(tmp_path), ep)
    monkeypatch.setenv("TQ_DATA_ROOT", str(tmp_path))
    mod = importlib.import_module(f"tensorrl_qas_b200.environments.{module_name}")
    conf = ep.conf
    conf["non_local_opt"]["global_iters"] = 60
    B, steps = 4, 4
    table = utils.dictionary_of_actions(conf["env"]["num_qubits"])
    rng = np.random.default_rng(5)
    plans = [[int(rng.integers(len(table))) for _ in range(steps)] for _ in range(B)]
    noisy = "noise" in module_name

    def fresh():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return [mod.CircuitEnv(conf, device=torch.device("cpu")) for _ in range(B)]

    # serial reference: each environment with its own noise stream
    _backend.reset_backends()
    serial = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for b, env in enumerate(fresh()):
            if noisy:
                mod.CircuitEnv.vc.seed(100 + b)
            env.reset()
            rec = []
            for a in plans[b]:
                obs, rwd, done = env.step(list(table[a]))
                rec.append((obs.numpy().copy(), float(rwd), done, float(env.energy), int(env.nfev)))
                if done:
                    break
            serial.append(rec)
    _backend.reset_backends()
    envs = fresh()
    ls = LockstepEnvs(envs, seeds=[100 + b for b in range(B)] if noisy else None)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ls.reset_all()
        got = [[] for _ in range(B)]
        alive = list(range(B))
        for t in range(steps):
            out = LockstepEnvs([envs[b] for b in alive], seeds=None)   # (sub-group of the still-running environments)
            out._rngs = None if ls._rngs is None else [ls._rngs[b] for b in alive]
            out._tag = ls._tag
            res = out.step_all([list(table[plans[b][t]]) for b in alive])
            assert out.last_group.rounds < out.last_group.evaluations or len(alive) == 1
            nxt = []
            for b, (obs, rwd, done) in zip(alive, res):
                got[b].append((obs.numpy().copy(), float(rwd), done, float(envs[b].energy), int(envs[b].nfev)))
                if not done:
                    nxt.append(b)
            alive = nxt
            if not alive:
                break
    for b in range(B):
        assert len(got[b]) == len(serial[b])
        for t, ((o1, r1, d1, e1, n1), (o0, r0, d0, e0, n0)) in enumerate(zip(got[b], serial[b])):
            assert (e1, n1, r1, d1) == (e0, n0, r0, d0), f"environment {b}, step {t}: (energy, nfev, reward, done)"
            assert np.array_equal(o1, o0), f"environment {b}, step {t}: observation"

    _backend.reset_backends()
