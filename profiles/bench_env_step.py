"""Host cost of re-binding circuits: what one `CircuitEnv.step()` pays outside COBYLA for its two
`construct_ansatz` + first-evaluation rounds (environments/environment_qulacs.py:409-411, 423-425), with the plan cache of
tq_set_circuit on (default) and off (TQ_PLAN_CACHE=0).  Prints one JSON line.

    python profiles/bench_env_step.py > profiles/env_step_r02.json"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def measure(cache):
    os.environ["TQ_PLAN_CACHE"] = "64" if cache else "0"
    import bench_workloads
    from tensorrl_qas_b200 import Simulator
    from tensorrl_qas_b200.circuit import append_random_gates
    w = bench_workloads.build("C2")            # BeH2-6q trainable environment circuit (128+ gates)
    sim = Simulator(w.n, 0)
    sim.set_dense_hamiltonian(w.dense)
    # an "episode": the circuit grows by one agent gate per step; every step binds its circuit twice and evaluates once
    import copy
    from tensorrl_qas_b200.circuit import parameter_batch
    rng = np.random.default_rng(0)
    episode = []
    gl = w.gl
    for _ in range(20):
        gl = append_random_gates(copy.deepcopy(gl), 1, rng)
        episode.append((gl, parameter_batch(gl, 1)))
    out = {}
    for label, n_rep in (("first_episode", 1), ("repeated_episode", 5)):
        t_bind = t_eval = 0.0
        for _ in range(n_rep):
            for g, p in episode:
                t0 = time.perf_counter()
                sim.set_circuit(g)
                sim.set_circuit(g)
                t1 = time.perf_counter()
                sim.energies(p)
                t2 = time.perf_counter()
                t_bind += t1 - t0
                t_eval += t2 - t1
        steps = n_rep * len(episode)
        out[label] = {"us_bind_two_set_circuit_per_step": 1e6 * t_bind / steps,
                      "us_first_energy_per_step": 1e6 * t_eval / steps}
    out["cache_stats"] = sim.plan_cache_stats()
    sim.close()
    return out


if __name__ == "__main__":
    res = {"workload": "C2 circuit growing by one gate per step, 20 steps; per step: tq_set_circuit x 2 + one tq_energy_batch_host",
           "plan_cache_on": measure(True), "plan_cache_off": measure(False)}
    print(json.dumps(res))
