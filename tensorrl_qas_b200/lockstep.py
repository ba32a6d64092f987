"""Lock-step evaluation of several environments (BASELINE.json config "256 batched environments").

The reference runs ONE environment: `CircuitEnv.step()` calls scipy's COBYLA, whose cost closure evaluates one energy at
a time (environments/environment_qulacs.py:429-445).  Environments are independent, so B of them can advance together:
every environment runs its unmodified `step()` in a worker thread, and each time all still-running workers have asked for
an energy, the B requests -- B DIFFERENT circuits / Hamiltonians / angle vectors -- go to the GPU as ONE launch
(`tq_energy_multi_host`).  Per environment nothing changes: the same sequence of evaluations with the same arguments and,
because every problem is evaluated by its own CTA with a fixed reduction order, bit-identical energies -- so trajectories
are those of the serial run (tested).

Under scipy >= 1.16 COBYLA itself is pure Python (pyprima, 2-12 ms per iteration) and holds the GIL, so
with scipy the wall-clock gain is bounded by the optimiser, not by the simulator (the GPU side of a round is one ~40 us
launch whatever B is; measured: 64 environments in lock-step are SLOWER than one after the other).  With the library's own
ask/tell COBYLA (`TQ_OPTIMIZER=native`, tensorrl_qas_b200/cobyla.py) a worker hands its whole optimisation problem to the
group (`LockstepGroup.optimise`): one thread hand-over per `step()`, all optimisers driven by one host loop, one launch per
round -- 617 environment steps per second for 64 environments against 54-66 serial with scipy.
"""
import itertools
import threading

import numpy as np

from .VQAs import _backend
from .simulator import energies_multi


class LockstepGroup:
    """Rendezvous of `n_workers` threads: `submit` blocks until every active worker has submitted (or retired), then the
    last arrival evaluates the whole round with `evaluate_round(items) -> energies` and wakes the others."""

    def __init__(self, n_workers, evaluate_round=None):
        self.n_active = int(n_workers)
        self.cv = threading.Condition()
        self.pending = {}
        self.opt_pending = {}   # whole optimisation problems (native COBYLA): worker -> (cost, x0, maxiter, slot_key, rng)
        self.results = {}
        self.rounds = 0
        self.evaluations = 0
        self.failure = None
        self.evaluate_round = evaluate_round or self._gpu_round

    @staticmethod
    def _gpu_round(items):
        sims = [it[0] for it in items]
        params = [it[1] for it in items]
        codes = [it[2] for it in items]
        return energies_multi(sims, params, codes if any(c is not None for c in codes) else None)

    def _waiting(self):
        return len(self.pending) + len(self.opt_pending)

    def _flush(self):
        if self.opt_pending:
            self._optimise_round()
        if not self.pending:
            self.cv.notify_all()
            return
        order = sorted(self.pending)
        try:
            energies = self.evaluate_round([self.pending[w] for w in order])
            for w, e in zip(order, energies):
                self.results[w] = np.float64(e)
        except BaseException as exc:  # wake everybody up with the error instead of dead-locking the group
            self.failure = exc
            for w in order:
                self.results[w] = exc
        self.rounds += 1
        self.evaluations += len(order)
        self.pending.clear()
        self.cv.notify_all()

    def _optimise_round(self):
        """Every waiting optimisation problem in one host loop (`cobyla.minimize_many`): per round the coordinator probes
        each running cost closure for its (handle, angles, noise codes) -- in that environment's evaluation context, so
        handles and noise streams are the ones the environment would use itself -- and evaluates the round as one launch.
        Each optimiser sees exactly the evaluations of its serial run."""
        from . import cobyla
        order = sorted(self.opt_pending)
        problems = [self.opt_pending[w] for w in order]
        ctx = _backend._ctx
        saved = (getattr(ctx, "slot_key", None), getattr(ctx, "rng", None))

        def batch(indices, points):
            items = []
            for i, x in zip(indices, points):
                cost, _, _, slot_key, rng = problems[i]
                ctx.slot_key, ctx.rng, ctx.capture = slot_key, rng, []
                try:
                    probe = cost(x)
                    got = ctx.capture
                finally:
                    ctx.capture = None
                if len(got) != 1 or float(probe) != 0.0:   # the probe must hand back the placeholder energy untouched
                    raise RuntimeError("lock-step optimisation needs a cost that is exactly one energy evaluation")
                items.append(got[0])
            self.rounds += 1
            self.evaluations += len(items)
            return self.evaluate_round(items)

        try:
            results, _ = cobyla.minimize_many(batch, [p[1] for p in problems], maxiter=[p[2] for p in problems])
            for w, r in zip(order, results):
                self.results[w] = r
        except BaseException as exc:
            self.failure = exc
            for w in order:
                self.results[w] = exc
        finally:
            ctx.slot_key, ctx.rng = saved
        self.opt_pending.clear()

    def optimise(self, worker, cost, x0, maxiter):
        """COBYLA (the library's own) of `cost` from `x0`, run together with the other workers' problems; blocks until
        this worker's result dict (x, fun, nfev, ...) is ready."""
        ctx = _backend._ctx
        with self.cv:
            self.opt_pending[worker] = (cost, np.array(x0, dtype=np.float64).reshape(-1), int(maxiter),
                                        getattr(ctx, "slot_key", None), getattr(ctx, "rng", None))
            if self._waiting() >= self.n_active:
                self._flush()
            else:
                while worker not in self.results:
                    self.cv.wait()
            out = self.results.pop(worker)
        if isinstance(out, BaseException):
            raise out
        return out

    def submit(self, worker, sim, params, codes=None):
        with self.cv:
            self.pending[worker] = (sim, np.array(params, dtype=np.float64).reshape(-1), codes)
            if self._waiting() >= self.n_active:
                self._flush()
            else:
                while worker not in self.results:
                    self.cv.wait()
            out = self.results.pop(worker)
        if isinstance(out, BaseException):
            raise out
        return out

    def retire(self, worker):
        """The worker will not submit again (its step finished or failed)."""
        with self.cv:
            self.n_active -= 1
            if self._waiting() and self._waiting() >= self.n_active:
                self._flush()


def run_lockstep(tasks, seeds=None, evaluate_round=None, slot_prefix="lockstep"):
    """Run the callables `tasks` (one per environment) in worker threads whose energy evaluations are batched round by
    round.  seeds[i] (optional) seeds worker i's private noise generator.  Returns (results, group); an exception in any
    worker is re-raised here after all workers have stopped.  The workers' handles stay cached under `slot_prefix` for the
    next call with the same prefix; `_backend.release_slots(slot_prefix)` drops them."""
    n = len(tasks)
    group = LockstepGroup(n, evaluate_round)
    results, errors = [None] * n, [None] * n

    def work(i):
        _backend._ctx.group = group
        _backend._ctx.worker = i
        _backend._ctx.slot_key = (slot_prefix, i)
        _backend._ctx.rng = np.random.default_rng(seeds[i]) if seeds is not None else None
        try:
            results[i] = tasks[i]()
        except BaseException as exc:
            errors[i] = exc
        finally:
            group.retire(i)
            _backend._ctx.group = None
            _backend._ctx.slot_key = None
            _backend._ctx.rng = None

    threads = [threading.Thread(target=work, args=(i,), daemon=True) for i in range(n)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in errors:
        if e is not None:
            raise e
    return results, group


_tags = itertools.count()


class LockstepEnvs:
    """B environments advanced together: `reset_all()` and `step_all(actions)` mirror `CircuitEnv.reset()` / `.step()`
    element-wise.  Each environment keeps its own libtqsim handle (its circuit differs from the others')."""

    def __init__(self, envs, seeds=None):
        self.envs = list(envs)
        self.seeds = seeds
        self._rngs = None if seeds is None else [np.random.default_rng(s) for s in seeds]
        self.last_group = None
        self._tag = f"lockstep-{next(_tags)}"   # (not id(self): a recycled id would inherit a dead driver's handles)

    def close(self):
        """Release the per-environment libtqsim handles of this driver."""
        _backend.release_slots(self._tag)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def _run(self, tasks):
        n = len(tasks)
        group = LockstepGroup(n)
        results, errors = [None] * n, [None] * n

        def work(i):
            _backend._ctx.group = group
            _backend._ctx.worker = i
            _backend._ctx.slot_key = (self._tag, i)
            _backend._ctx.rng = self._rngs[i] if self._rngs is not None else None   # persists across steps
            try:
                results[i] = tasks[i]()
            except BaseException as exc:
                errors[i] = exc
            finally:
                group.retire(i)
                _backend._ctx.group = None
                _backend._ctx.slot_key = None
                _backend._ctx.rng = None

        threads = [threading.Thread(target=work, args=(i,), daemon=True) for i in range(n)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for e in errors:
            if e is not None:
                raise e
        self.last_group = group
        return results

    def reset_all(self):
        return self._run([env.reset for env in self.envs])

    def step_all(self, actions, train_flag=True):
        if len(actions) != len(self.envs):
            raise ValueError("one action per environment")
        return self._run([(lambda e=env, a=act: e.step(a, train_flag)) for env, act in zip(self.envs, actions)])

    def illegal_actions_all(self):
        return [env.illegal_action_new() for env in self.envs]
