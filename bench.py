#!/usr/bin/env python
"""bench.py -- energy evals/sec of the batched fp64 circuit-simulation hot path (BASELINE.json metric).

Default workload: 20-qubit open Heisenberg chain (77 Pauli terms, 20 flip-mask groups),
brickwork-shaped synthetic circuit of 440 gates (seed 5), 64 parameter sets per GPU (theta_0 + U(-0.1, 0.1), seeds
1000 + b).  One "step" = the energies of the whole batch.  States (64 x 16 MiB = 1 GiB per GPU) never fit the 126 MB L2,
so every pass streams from HBM; no explicit L2 flush is needed.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload C1|C2|C3|C4|C5|C5L|C5G]
                    [--scaling weak|strong] [--batch B]

N > 1 is launched by torchrun, one rank per GPU; the batch is sharded (weak: B per rank; strong: B split over the ranks)
and the only collective is the all-gather of the per-circuit energies.  `--impl reference` times the CPU restatement of
the reference path (oracle/tq_oracle.c; qulacs itself is probed for and used when importable) on the host cores.

What the JSON line carries besides the contract keys:
  roofline            the dominant kernel of the step, timed live with CUDA events around every launch (tq_profile_*):
                      `achieved` / `frac` on the ALGORITHMIC bytes of the reference's one-pass-per-gate model (SURVEY.md 8d;
                      fusion makes this exceed 1), `measured` on the bytes the compiled plan really moves, every kernel's
                      share of the step in `kernels`, the FP64 tensor-core work against a DMMA peak measured in the same run
  cpu_baseline        the oracle port (and qulacs when present) on a bounded sample of the same workload
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "energy_evals_per_sec"
UNIT = "evals/s"


def config_dict(w, n_gpus, batch_per_gpu, scaling):
    gl = w.gl
    return {
        "workload": w.title, "name": w.name, "n_qubits": w.n, "gates": int(gl.n_unitary), "rotations": int(gl.n_params),
        "cnots": int(gl.count("CNOT")), "noise_channels": int(len(gl) - gl.n_unitary), "mode": w.mode,
        "hamiltonian": "Pauli sum" if w.pauli is not None else "dense", "flip_mask_groups": int(w.groups),
        "initial_state": "loaded" if w.init is not None else "|0...0>",
        "batch_per_gpu": int(batch_per_gpu), "global_batch": int(batch_per_gpu * n_gpus),
        "sharding": f"batch x{n_gpus} ({scaling})" if n_gpus > 1 else "none",
        "l2": ("working set %.0f MiB of states per GPU and pass %s 126 MB L2" %
               (batch_per_gpu * (16 << (2 * w.n if w.mode == "dm" else w.n)) / 2 ** 20,
                ">" if batch_per_gpu * (16 << (2 * w.n if w.mode == "dm" else w.n)) > 126e6 else "<=")) +
              ("; inputs larger than L2, no flush needed" if w.bound == "hbm" else
               "; the path is not HBM-bound at this size (see roofline.bound_note)"),
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for t, line in self.rows:
            if not (t0 - 0.05 <= t <= t1 + 0.15):
                continue
            f = [v.strip() for v in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------- CPU arms ----------------
def probe_qulacs():
    try:
        import qulacs  # noqa: F401
        return True, getattr(qulacs, "__version__", "?")
    except Exception as exc:  # ModuleNotFoundError here and on the GPU boxes of this pool
        return False, f"{type(exc).__name__}: {exc}"


def qulacs_energies(w, params):
    """The reference's own loop (environments/VQAs/VQE_qulacs.py:66-86) on this workload: ParametricQuantumCircuit,
    set_parameter per angle, update_quantum_state, expectation.  Only reachable where `import qulacs` works."""
    import qulacs
    from qulacs import ParametricQuantumCircuit, QuantumState
    from qulacs.gate import CNOT
    n = w.n
    circ = ParametricQuantumCircuit(n)
    slots = []
    for kind, q0, q1, pidx, fixed in w.gl.tuples():
        if kind == 3:
            circ.add_gate(CNOT(q0, q1))
        elif kind in (0, 1, 2):
            (circ.add_parametric_RX_gate, circ.add_parametric_RY_gate, circ.add_parametric_RZ_gate)[kind](q0, fixed)
            slots.append(pidx)
    obs = None
    if w.pauli is not None:
        obs = qulacs.Observable(n)
        for xm, zm, c in zip(*w.pauli):
            s = " ".join(f"{'IXZY'[((int(xm) >> q) & 1) | (((int(zm) >> q) & 1) << 1)]} {q}" for q in range(n)
                         if ((int(xm) | int(zm)) >> q) & 1)
            obs.add_operator(float(c), s)
    out = np.empty(len(params))
    for b, row in enumerate(params):
        for i, pidx in enumerate(slots):
            if pidx >= 0:
                circ.set_parameter(i, float(row[pidx]))
        st = QuantumState(n)
        if w.init is not None:
            st.load(np.asarray(w.init, dtype=np.complex128))
        circ.update_quantum_state(st)
        if obs is not None:
            out[b] = obs.get_expectation_value(st)
        else:
            psi = st.get_vector()
            out[b] = float((np.conj(psi).T @ w.dense @ psi).real)
    return out


def cpu_sample(w, params, n_evals, threads=0):
    p = params[:n_evals]
    t0 = time.perf_counter()
    e, used = w.oracle_energies(p, nthreads=threads, return_threads=True)
    dt = time.perf_counter() - t0
    return e, n_evals / dt, used, dt


def bounded_sample(w, batch, cores, seconds=3.0):
    """Evaluations for about `seconds` of oracle work per repetition (the oracle runs one evaluation per thread; at least
    one evaluation per core, at most the batch)."""
    state_bytes = 16 << (2 * w.n if w.mode == "dm" else w.n)
    passes = 2 * len(w.gl) + max(w.groups, 1)
    sec_per_eval = state_bytes * passes / 4e9 + 2e-5      # ~4 GB/s per core of state traffic, 20 us floor
    target = max(cores, int(seconds * cores / max(sec_per_eval, 1e-9)))
    return int(max(1, min(batch, target, 65536)))


def run_reference(args, rank, world):
    """CPU arm: the reference path on the host cores -- qulacs when it is importable, else its restatement
    (oracle/tq_oracle.c, one full-state pass per gate) on all host threads."""
    if rank != 0:
        return
    import bench_workloads
    from oracle import c_oracle
    w = bench_workloads.build(args.workload)
    batch = args.batch or w.batch
    cores = c_oracle.max_threads()
    sample = bounded_sample(w, batch * world, cores)
    params = w.params(sample)
    have_q, q_info = probe_qulacs()
    kind = "port"
    run = lambda: cpu_sample(w, params, sample)[2]   # noqa: E731
    if have_q:
        kind = "qulacs"
        run = lambda: (qulacs_energies(w, params), int(os.environ.get("OMP_NUM_THREADS", cores)))[1]   # noqa: E731
    for _ in range(args.warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        used = run()
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(w, args.gpus, batch, args.scaling),
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": kind,
                         "sample": f"{sample} evaluations of the workload per step ("
                                   + ("qulacs " + q_info if have_q else "oracle/tq_oracle.c: one full-state pass per gate, one "
                                      "evaluation per thread; qulacs probe: " + q_info) + ")"},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- GPU arm -----------------
def main():
    import bench_workloads
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C5", choices=list(bench_workloads.NAMES) + [n.lower() for n in bench_workloads.NAMES])
    ap.add_argument("--batch", type=int, default=0, help="parameter sets per GPU (weak) / in total (strong); 0 = the workload's")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    args.workload = args.workload.upper()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from tensorrl_qas_b200 import Simulator
    from tensorrl_qas_b200.simulator import fp64_peak

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (tensorrl_qas_b200 has no CPU fallback)")
    if world != args.gpus:
        raise SystemExit(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; launch N>1 with torchrun")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # NCCL announces itself on STDOUT ("NCCL version ...") when the communicator is created; the contract is ONE JSON
    # line on stdout, so file descriptor 1 points at stderr until the warm-up (first collective) is over
    saved_stdout = None
    if world > 1:
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)

    w = bench_workloads.build(args.workload)
    total = args.batch or w.batch
    if args.scaling == "strong":      # the global batch is fixed; every rank takes its contiguous share
        if total % world:
            raise SystemExit("bench.py: --scaling strong needs a batch divisible by the number of GPUs")
        B = total // world
    else:
        B = total
    # each rank owns a contiguous slice of the global batch: elements [rank*B, (rank+1)*B)
    params = w.params(B, seed0=1000 + rank * B)
    sim = Simulator(w.n, local_rank)
    w.bind(sim)
    which = 1 if w.mode == "dm" else 0
    info = sim.plan_info(which)

    from tensorrl_qas_b200.sharding import HostBatchGather, OverlappedGather
    p_dev = torch.from_numpy(params).to(dev)
    out = torch.empty(B, dtype=torch.float64, device=dev)
    # the path's only collective: every rank receives all B * world energies.  It runs on a side stream under the next
    # step's kernels (two result buffers alternate); sync_all() at the end of the timed region waits for all of them
    gather = OverlappedGather(B, world, dev) if world > 1 else None

    def step():
        if gather is None:
            sim.energies_dev(p_dev, out=out, mode=w.mode)
            return out
        buf = gather.local_buffer()
        sim.energies_dev(p_dev, out=buf, mode=w.mode)
        gather.submit()
        return buf

    # N > 1: the host caller needs every rank's energies -- host buffers in, the gathered vector out, one synchronisation
    host_gather = HostBatchGather(sim, B, params.shape[1], world, dev, mode=w.mode) if world > 1 else None

    def host_step():
        if host_gather is not None:
            e_all = host_gather(params)
            return e_all[rank * B:(rank + 1) * B].copy(), e_all
        e = sim.energies_dm(params) if w.mode == "dm" else sim.energies(params)
        return e, e

    def sync_all():
        if gather is not None:
            gather.wait()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    sync_all()
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)

    # ---------------- timed region: device time with CUDA events on the launch stream, max over ranks ----------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    time.sleep(0.25 if rank == 0 else 0)
    sync_all()
    launches0 = sim.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    sync_all()
    t_wall1 = time.perf_counter()
    launches = sim.launch_count - launches0
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    evals = B * world * args.steps
    value = evals / (ms * 1e-3)

    # ---------------- end-to-end through the host-buffer API: H2D of the angles + D2H of the energies inside -----
    for _ in range(2):
        host_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e_host, _ = host_step()
    t_e2e = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    e2e_value = evals / t_e2e
    e_dev = (gather.local[(gather.i - 1) % 2] if gather is not None else out).cpu().numpy()
    assert np.array_equal(e_host, e_dev), "host and device entry points disagree"
    if gather is not None:   # the gathered vector holds every rank's slice in rank order
        full = gather.wait().cpu().numpy()
        assert np.array_equal(full[rank * B:(rank + 1) * B], e_dev), "all-gather misplaced this rank's energies"

    # ---------------- per-launch device times of a few more steps (CUDA events around every kernel launch) --------
    prof_steps = 3
    sim.profile(True)
    for _ in range(prof_steps):
        sim.energies_dev(p_dev, out=out, mode=w.mode)
    recs = sim.profile_read()
    sim.profile(False)

    # latency regime of the reference (one cost evaluation per call): C1-C3 report it beside the batched number
    latency = None
    if world == 1 and w.n <= 12 and w.mode == "pure":
        one = params[:1].copy()
        for _ in range(20):
            sim.energies(one)
        n_lat = 300
        t0 = time.perf_counter()
        for _ in range(n_lat):
            sim.energies(one)
        latency = {"us_per_eval_B1_host_call": 1e6 * (time.perf_counter() - t0) / n_lat,
                   "note": "one tq_energy_batch_host call per evaluation (the reference's COBYLA loop shape): angles and "
                           "energy through pinned host memory, one launch, polled completion"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline from the live per-launch times -------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    per_step = len(recs) // prof_steps
    kernels = []
    for i in range(per_step):   # launch i of a step, averaged over the profiled steps
        rs = [recs[s * per_step + i] for s in range(prof_steps)]
        kernels.append({"kernel": rs[0]["kernel"], "ms": float(np.mean([r["ms"] for r in rs])),
                        "model_bytes": rs[0]["model_bytes"], "alg_bytes": rs[0]["alg_bytes"],
                        "dmma_flops": rs[0].get("dmma_flops", 0.0)})
    step_ms_prof = sum(k["ms"] for k in kernels) or 1e-9
    dmma_peak = fp64_peak(local_rank, 0)
    dfma_peak = fp64_peak(local_rank, 1)
    for k in kernels:
        k["share_of_step"] = k["ms"] / step_ms_prof
        k["model_gbs"] = k["model_bytes"] / (k["ms"] * 1e-3) / 1e9
        k["frac_of_hbm_peak_model"] = k["model_gbs"] / peak
        # FP64 tensor-core work of the launch against the DMMA peak measured in this run: with the HBM fraction beside
        # it, this says which roof the launch is under
        k["dmma_tflops"] = k["dmma_flops"] / (k["ms"] * 1e-3) / 1e12
        k["frac_of_fp64_tensor_peak"] = k["dmma_tflops"] / dmma_peak if dmma_peak else None
    dom = max(kernels, key=lambda k: k["ms"]) if kernels else None
    bytes_eval = w.algorithmic_bytes_per_eval()
    per_gpu_evals_s = B * args.steps / (ms * 1e-3)
    model_step = sum(k["model_bytes"] for k in kernels)
    # measured DRAM traffic of one step: only from an ncu capture of THIS plan (profiles/traffic.json names the launch
    # sequence it was taken on); anything else would be a stale constant
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    sig = [k["kernel"] for k in kernels]
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("batch") == B and tj.get("workload") == w.name and tj.get("launch_sequence") == sig:
            traffic = tj.get("dram_bytes_per_step")
    counts = sim.plan_counts(which)
    tile_amps = 1 << info["tile_bits"]
    dmma_flop_eval = 2.0 * 16.0 * counts["tensor_core_block_tiles"] * tile_amps
    dmma_tflops = dmma_flop_eval * per_gpu_evals_s / 1e12
    roofline = None
    if dom is not None:
        alg_gbs = dom["alg_bytes"] / (dom["ms"] * 1e-3) / 1e9
        # The headline is the roof the dominant launch is actually under: the larger of its two measured fractions (HBM
        # on the bytes the compiled plan moves, FP64 tensor pipe on the block flops it executes).  SURVEY.md 8d's
        # algorithmic-bytes figure (one state pass per gate, the reference's model) stays beside it as `algorithmic`:
        # a launch fuses many gates per pass, so that fraction is > 1 and is a statement about fusion, not bandwidth.
        tensor_bound = (dom["frac_of_fp64_tensor_peak"] or 0.0) > dom["frac_of_hbm_peak_model"]
        roofline = {
            "bound": "tensor" if tensor_bound else "hbm", "kernel": dom["kernel"],
            "achieved": dom["dmma_tflops"] if tensor_bound else dom["model_gbs"],
            "peak": dmma_peak if tensor_bound else peak, "unit": "TFLOP/s" if tensor_bound else "GB/s",
            "frac": dom["frac_of_fp64_tensor_peak"] if tensor_bound else dom["frac_of_hbm_peak_model"],
            "traffic": traffic,
            "peak_source": ("FP64 mma.sync.m8n8k4 peak measured in this run (tq_fp64_peak; MEASURED_PEAKS.json has no FP64 "
                            "figure; DMMA and DFMA share one datapath: profiles/microbench/fp64_mixed.cu)"
                            if tensor_bound else peak_src),
            "stated_bound": w.bound, "bound_note": w.bound_note,
            "ms_per_launch": dom["ms"],
            "note": "achieved / peak / frac: the dominant launch against the roof it is under, from its CUDA-event time in "
                    "this run -- FP64 tensor-core flops of the fused blocks it executes (2 x 16 FMA per amplitude and block, "
                    "known zeros left out) or the HBM bytes the compiled plan moves in it; `algorithmic` is SURVEY.md 8d's "
                    "figure for the same launch (bytes of the reference's one-state-pass-per-gate model over the same time: "
                    "> 1 x the HBM peak because the launch fuses many gates per pass)",
            "algorithmic": {"bytes_per_launch": dom["alg_bytes"], "gbs": alg_gbs, "hbm_peak_gbs": peak,
                            "frac_of_hbm_peak": alg_gbs / peak, "hbm_peak_source": peak_src},
            "measured": {"share_of_step": dom["share_of_step"], "model_bytes_per_launch": dom["model_bytes"],
                         "model_gbs": dom["model_gbs"], "frac_of_peak_model": dom["frac_of_hbm_peak_model"],
                         "dmma_tflops": dom["dmma_tflops"], "frac_of_fp64_tensor_peak": dom["frac_of_fp64_tensor_peak"],
                         "binding_roof": ("fp64 tensor pipe" if (dom["frac_of_fp64_tensor_peak"] or 0.0) >
                                          dom["frac_of_hbm_peak_model"] else "hbm"),
                         "frac_of_binding_roof": max(dom["frac_of_fp64_tensor_peak"] or 0.0,
                                                     dom["frac_of_hbm_peak_model"])},
            "step": {"ms_profiled": step_ms_prof, "algorithmic_bytes_per_eval": bytes_eval,
                     "algorithmic_gbs": bytes_eval * per_gpu_evals_s / 1e9,
                     "frac_algorithmic": bytes_eval * per_gpu_evals_s / 1e9 / peak,
                     "model_bytes_per_step": model_step, "model_gbs": model_step / (ms / args.steps * 1e-3) / 1e9,
                     "frac_of_peak_model": model_step / (ms / args.steps * 1e-3) / 1e9 / peak,
                     "traffic_ncu_bytes_per_step": traffic},
            "kernels": kernels,
            "fp64_tensor": {"achieved": dmma_tflops, "peak": dmma_peak, "unit": "TFLOP/s",
                            "frac": dmma_tflops / dmma_peak if dmma_peak else None,
                            "fp64_pipe_peak": dfma_peak, "blocks_per_eval": counts["tensor_core_blocks"],
                            "block_tiles_per_eval": counts["tensor_core_block_tiles"], "flop_per_eval": dmma_flop_eval,
                            "peak_source": "measured in this run (tq_fp64_peak: mma.sync.m8n8k4.f64 / DFMA chains, best of 3)"},
        }

    # ---------------- CPU baseline beside it: the reference path on the host cores, bounded sample ----------------
    cpu = None
    have_q, q_info = probe_qulacs()
    if not args.no_cpu_baseline and world == 1:   # rank 0 at N = 1 only (task contract)
        from oracle import c_oracle
        cores = c_oracle.max_threads()
        sample = bounded_sample(w, B, cores, seconds=5.0)
        e_cpu, rate1, used, dt1 = cpu_sample(w, params, sample)
        _, rate2, _, dt2 = cpu_sample(w, params, sample)
        rate, dt = 2 * sample / (dt1 + dt2), dt1 + dt2
        err = float(np.abs(e_cpu - e_host[:len(e_cpu)]).max())
        cpu = {"value": rate, "unit": UNIT, "cores": used, "kind": "port",
               "sample": f"{len(e_cpu)} of the {B} evaluations of one step, twice ({dt:.1f} s), oracle/tq_oracle.c",
               "max_abs_dE_vs_gpu": err, "qulacs_probe": q_info if not have_q else f"qulacs {q_info}"}
        if have_q:   # the reference's own backend is here: time its loop, pin the oracle and the GPU path against it
            nq = min(sample, 128)
            t0 = time.perf_counter()
            e_q = qulacs_energies(w, params[:nq])
            dtq = time.perf_counter() - t0
            cpu.update({"kind": "qulacs", "value": nq / dtq, "cores": int(os.environ.get("OMP_NUM_THREADS", cores)),
                        "sample": f"{nq} evaluations through qulacs' ParametricQuantumCircuit ({dtq:.1f} s)",
                        "oracle_port_value": rate, "max_abs_dE_qulacs_vs_gpu": float(np.abs(e_q - e_host[:nq]).max()),
                        "max_abs_dE_qulacs_vs_oracle": float(np.abs(e_q - e_cpu[:nq]).max())})
            np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"qulacs_{w.name}.npz"), params=params[:nq], energies=e_q)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config_dict(w, world, B, args.scaling),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(params.nbytes),
                "d2h_bytes_per_step": int(8 * B * world) if world > 1 else int(8 * B),
                "includes_gather": world > 1,
                "api": ("tensorrl_qas_b200.sharding.HostBatchGather (pinned H2D, tq_energy_batch, NCCL all-gather, pinned D2H "
                        "of all ranks' energies, one synchronisation)" if world > 1 else
                        "tq_energy_batch_host / tq_energy_dm_batch_host (pinned staging inside the call)")},
        "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        "latency": latency, "plan": dict(info, **counts),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
