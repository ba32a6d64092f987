#!/usr/bin/env python
"""bench.py -- energy evals/sec of the batched fp64 statevector hot path (BASELINE.json metric).

Workload: 20-qubit open Heisenberg chain (77 Pauli terms, 20 flip-mask groups),
brickwork-shaped synthetic circuit of 440 gates (seed 5), 64 parameter sets per GPU (theta_0 + U(-0.1, 0.1),
seeds 1000 + b).  One "step" = the energies of the whole batch.  States (64 x 16 MiB = 1 GiB per GPU) never fit
the 126 MB L2, so every pass streams from HBM; no explicit L2 flush is needed.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 is launched by torchrun, one rank per GPU; the batch is sharded (64 per rank, weak scaling) and the only
collective is the all-gather of the per-circuit energies.  `--impl reference` times the CPU restatement of the
reference path (oracle/tq_oracle.c; qulacs itself is not installable here) on the host cores.
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

N_QUBITS = 20
GATES_PER_BRICK = 21
AGENT_GATES = 41
CIRCUIT_SEED = 5
BATCH_PER_GPU = 64
FP64_DMMA_PEAK_TFLOPS = 37.2   # profiles/microbench/fp64_peak.cu on this pool's B200
METRIC = "energy_evals_per_sec"
UNIT = "evals/s"


def workload():
    from tensorrl_qas_b200 import loaders
    from tensorrl_qas_b200.circuit import brickwork_circuit, parameter_batch
    gl = brickwork_circuit(N_QUBITS, GATES_PER_BRICK, AGENT_GATES, CIRCUIT_SEED)
    paulis, w = loaders.heisenberg_terms(N_QUBITS)
    x, z = loaders.pauli_masks(paulis, N_QUBITS)
    return gl, (x, z, w), parameter_batch


def config_dict(n_gpus, gl, groups):
    return {
        "workload": "C5: 20-qubit Heisenberg chain energy sweep, brickwork synthetic circuit",
        "n_qubits": N_QUBITS, "gates": len(gl), "rotations": gl.n_params, "cnots": gl.count("CNOT"),
        "pauli_terms": 3 * (N_QUBITS - 1) + N_QUBITS, "flip_mask_groups": groups,
        "batch_per_gpu": BATCH_PER_GPU, "global_batch": BATCH_PER_GPU * n_gpus,
        "sharding": f"batch x{n_gpus}" if n_gpus > 1 else "none",
        "l2": "working set 1 GiB of states per GPU per pass > 126 MB L2 (no flush needed)",
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


def algorithmic_bytes_per_eval(n, gates, groups, init_loaded):
    """qulacs' own traffic model, one read + one write of the state per gate, one read per
    flip-mask group, one read of a loaded initial state."""
    return 16 * (1 << n) * (2 * gates + groups + (1 if init_loaded else 0))


def cpu_baseline_sample(gl, ham, params, n_evals, threads=0):
    from oracle import c_oracle
    p = params[:n_evals]
    t0 = time.perf_counter()
    e, used = c_oracle.energies(gl, p, pauli=ham, nthreads=threads, return_threads=True)
    dt = time.perf_counter() - t0
    return e, n_evals / dt, used, dt


def run_reference(args, rank, world):
    """CPU arm: the restated reference algorithm (one full-state pass per gate) on all host threads."""
    if rank != 0:
        return
    from oracle import c_oracle
    gl, ham, parameter_batch = workload()
    cores = c_oracle.max_threads()
    sample = max(cores, 1)
    params = parameter_batch(gl, sample)
    for _ in range(args.warmup):
        cpu_baseline_sample(gl, ham, params, min(sample, cores))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, _, used, _ = cpu_baseline_sample(gl, ham, params, sample)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.gpus, gl, 20),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port",
                         "sample": f"{sample} evals per step of the same 20q/440-gate workload "
                                   "(oracle/tq_oracle.c, qulacs-shaped: one state pass per gate; qulacs not installable)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="parameter sets per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from tensorrl_qas_b200 import Simulator

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

    gl, ham, parameter_batch = workload()
    B = args.batch
    # each rank owns a contiguous slice of the global batch: elements [rank*B, (rank+1)*B)
    params = np.stack([parameter_batch(gl, 1, seed0=1000 + rank * B + b)[0] for b in range(B)])
    sim = Simulator(N_QUBITS, local_rank)
    sim.set_pauli_hamiltonian(*ham)
    sim.set_circuit(gl)
    info = sim.plan_info(0)

    from tensorrl_qas_b200.sharding import gather_energies
    p_dev = torch.from_numpy(params).to(dev)
    out = torch.empty(B, dtype=torch.float64, device=dev)

    def step():
        sim.energies_dev(p_dev, out=out)
        if world > 1:  # the path's only collective: every rank receives all B * world energies
            gather_energies(out, B * world)

    def sync_all():
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
        sim.energies(params)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e_host = sim.energies(params)
    t_e2e = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    e2e_value = evals / t_e2e
    assert np.array_equal(e_host, out.cpu().numpy()), "host and device entry points disagree"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (tile_pass_kernel) --------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    bytes_eval = algorithmic_bytes_per_eval(N_QUBITS, len(gl), info["groups"], False)
    per_gpu_evals_s = B * args.steps / (ms * 1e-3)
    achieved = bytes_eval * per_gpu_evals_s / 1e9
    state_bytes = 16 << N_QUBITS
    passes_rw = info["gate_passes"]
    passes_ro = info["expectation_passes"]
    moved = state_bytes * (2 * passes_rw - 1 + passes_ro)  # first pass does not read, last gate pass writes iff followed
    traffic = None  # ncu dram__bytes_read.sum + dram__bytes_write.sum over the launches of one step (B = 64)
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("batch") == B:
            traffic = tj.get("dram_bytes_per_step")
    counts = sim.plan_counts(0)
    # FP64 tensor-core work: every dense block is an 8x8 real DMMA product = 16 FMA per amplitude
    # (block-tile pairs actually executed: tiles that are still all-zero early in a run from |0...0> are skipped)
    tile_amps = 1 << info["tile_bits"]
    dmma_flop_eval = 2.0 * 16.0 * counts["tensor_core_block_tiles"] * tile_amps
    dmma_tflops = dmma_flop_eval * per_gpu_evals_s / 1e12
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
        "peak_source": peak_src,
        "kernel": f"tile_pass_mma_kernel ({passes_rw} gate-pass launches per step, 78 % of the step) + expect_direct_kernel "
                  f"(the {passes_ro} expectation-only passes in one persistent launch, 21 %); per-step totals",
        "algorithmic_bytes_per_eval": bytes_eval, "algorithmic_bytes_per_step": bytes_eval * B,
        "note": "algorithmic bytes = qulacs' unfused model 16*2^n*(2G+M) (SURVEY.md 8d); the fused passes move far "
                "fewer real bytes, so frac > 1 is expected -- hbm_actual has the real traffic, fp64_tensor the bound "
                "the fused kernel actually runs against",
        "hbm_actual": {"bytes_per_eval_model": moved, "gbs_model": moved * per_gpu_evals_s / 1e9,
                       "frac_of_peak_model": moved * per_gpu_evals_s / 1e9 / peak,
                       "bytes_per_eval_ncu": (traffic / B) if traffic else None,
                       "gbs_ncu": (traffic / B * per_gpu_evals_s / 1e9) if traffic else None,
                       "frac_of_peak_ncu": (traffic / B * per_gpu_evals_s / 1e9 / peak) if traffic else None,
                       "gate_passes": passes_rw, "expectation_passes": passes_ro,
                       "note": "model = one read + one write of the state per gate pass, one read per expectation pass; "
                               "ncu = measured DRAM bytes of one step (profiles/traffic.json): the first gate pass of a "
                               "run from |0...0> only touches the tiles that are not all-zero yet"},
        "fp64_tensor": {"achieved": dmma_tflops, "peak": FP64_DMMA_PEAK_TFLOPS, "unit": "TFLOP/s",
                        "frac": dmma_tflops / FP64_DMMA_PEAK_TFLOPS, "blocks_per_eval": counts["tensor_core_blocks"],
                        "block_tiles_per_eval": counts["tensor_core_block_tiles"],
                        "dense_block_tiles_per_eval": counts["tensor_core_blocks"] * ((1 << N_QUBITS) // tile_amps),
                        "flop_per_eval": dmma_flop_eval,
                        "peak_source": "measured: profiles/microbench/fp64_peak.cu, mma.sync.m8n8k4.f64 on this pool's "
                                       "B200 (63.9 FMA/clk/SM x 148 SMs x 1.965 GHz); MEASURED_PEAKS.json has no FP64 figure"},
    }

    # ---------------- CPU baseline beside it: the oracle port on the host cores, bounded sample -------------------
    cpu = None
    if not args.no_cpu_baseline and world == 1:   # rank 0 at N = 1 only (task contract)
        from oracle import c_oracle
        # bounded sample, about 10 s of CPU work: the evaluations of one step, as many of them as 16 threads finish in
        # ~5 s each way (two timed repetitions; the whole step where the host has >= 16 cores)
        cores = c_oracle.max_threads()
        sample = min(B, max(cores, 1) * 4)
        e_cpu, rate1, used, dt1 = cpu_baseline_sample(gl, ham, params, sample)
        _, rate2, _, dt2 = cpu_baseline_sample(gl, ham, params, sample)
        rate, dt = 2 * sample / (dt1 + dt2), dt1 + dt2
        err = float(np.abs(e_cpu - e_host[:len(e_cpu)]).max())
        cpu = {"value": rate, "unit": UNIT, "cores": used, "kind": "port",
               "sample": f"{len(e_cpu)} of the {B} evals of one step, twice ({dt:.1f} s), oracle/tq_oracle.c",
               "max_abs_dE_vs_gpu": err}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config_dict(world, gl, info["groups"]),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(params.nbytes),
                "d2h_bytes_per_step": int(8 * B)},
        "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        "plan": dict(info, **counts),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
