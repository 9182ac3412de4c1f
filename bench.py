#!/usr/bin/env python
"""bench.py -- env-steps/sec (agents x envs x steps / s) of the GCBF+ rollout hot path.

Contract: `python bench.py --gpus N --steps K --warmup W [--impl reference]`
(N > 1: launched by torch.distributed.run, one rank per GPU).  One JSON line on rank 0.

A "step" is one steady-state T = 256-step closed-loop rollout of E envs per
GPU of the BASELINE.json config `DoubleIntegrator n=512, 16 envs, obs 8, n-rays 32`:
per env-step {actor GNN forward, a = 2 pi + u_ref, clip, Euler, reward/cost, LiDAR ray cast,
radius neighbour lists} with the reference's pretrained DoubleIntegrator weights.
`value` : inputs resident in HBM (CUDA events around K replays of the rollout CUDA graph).
`e2e`   : through RolloutEngine with HOST (pinned) initial conditions, H2D + D2H inside.
`--impl reference`: the restated reference (dense padded N x N formulation of
gcbfplus/utils/graph.py + nn/gnn.py, torch-CPU fp32, all host threads) on a bounded
sample -- JAX is not installable in this image, so the CPU oracle is the arm.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "env-steps/sec (agents x envs x steps/s) DoubleIntegrator n=512"
ENV_ID, N_AGENTS, ENVS_PER_GPU, N_OBS, N_RAYS, AREA, T_STEPS = "DoubleIntegrator", 512, 16, 8, 32, 32.0, 256
F_EDGE, F_NODE = 267520, 461312          # FLOP per real edge / per agent (actor, DI)
B_ALG = 312                              # algorithmic bytes per agent-env-step (DI, recording on)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ======================================================================================== ours
def run_ours(args):
    import torch
    import torch.distributed as dist
    from helpers import GOLDEN, product_algo
    from gcbfplus_b200 import _lib
    from gcbfplus_b200.env import make_env
    from gcbfplus_b200.trainer.rollout import RolloutEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load(build_if_missing=False)

    E, N, T = args.envs_per_gpu, N_AGENTS, args.T
    env = make_env(ENV_ID, N, area_size=AREA, num_obs=N_OBS, n_rays=N_RAYS, device=dev)
    algo = product_algo(env, ENV_ID)
    g0 = env.reset(1000 + rank, n_envs=E)
    eng = RolloutEngine(env, E, T=T, n_obs=N_OBS)
    eng.set_params(algo.actor_params)
    # host-side (pinned) copies for the e2e leg
    h_agent = g0.agent.cpu().pin_memory()
    h_goal = g0.goal.cpu().pin_memory()
    h_obs = g0.obstacle.packed.cpu().pin_memory()
    h_rew = torch.empty(T, E).pin_memory()
    h_cost = torch.empty(T, E).pin_memory()
    h_final = torch.empty(E, N, env.state_dim).pin_memory()
    eng.set_initial(g0.agent, g0.goal, g0.obstacle)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up (first run captures the CUDA graph)
    for _ in range(max(args.warmup, 3)):
        eng.run(check=False)
    torch.cuda.synchronize()
    eng.check_overflow()
    n_edges = eng.counters[:, 0].float().mean().item()
    deg_real = n_edges / (E * N)

    # ---- value: device-resident inputs
    sampler = ClockSampler(local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if rank == 0:
        sampler.start()
    ev0.record()
    for _ in range(args.steps):
        eng.run(check=False)
    ev1.record()
    barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.stop() if rank == 0 else {}
    eng.check_overflow()
    ms_per_step = ms_total / args.steps
    value = N * E * world * T / (ms_per_step * 1e-3)

    # ---- e2e: host buffers in, host results out, every step
    def e2e_step():
        eng.agent[0].copy_(h_agent, non_blocking=True)
        eng.goal.copy_(h_goal, non_blocking=True)
        eng.obstacles.copy_(h_obs, non_blocking=True)
        eng.run(check=False)
        h_rew.copy_(eng.rewards, non_blocking=True)
        h_cost.copy_(eng.costs, non_blocking=True)
        h_final.copy_(eng.agent[T], non_blocking=True)

    e2e_step()
    barrier()
    ev0.record()
    for _ in range(args.steps):
        e2e_step()
    ev1.record()
    barrier()
    ms_e2e = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    e2e_value = N * E * world * T / (ms_e2e * 1e-3)
    h2d = (h_agent.numel() + h_goal.numel() + h_obs.numel()) * 4
    d2h = (h_rew.numel() + h_cost.numel() + h_final.numel()) * 4

    # ---- train step (update_inner, reported separately per SURVEY 8d): minibatch of 256 graphs of the
    # recorded rollout, sharded over ranks, incl. the denominator + packed-gradient all-reduces
    train = None if args.no_train else train_step_bench(torch, dist if world > 1 else None, env, algo, eng, rank, world,
                                                        args, max_over_ranks, barrier)

    # ---- roofline of the dominant kernel (fp32 GEMM 256x256 over the edge rows), timed alone
    roof = gemm_roofline(torch, _lib, dev, int(n_edges), E * N) if rank == 0 else None
    cpu = cpu_baseline(args, steps=1) if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None

    if rank == 0:
        hbm, tf_burst, tf_sus, src = load_peaks()
        flop_step = (deg_real * F_EDGE + F_NODE) * N * E * world          # per env-step, whole job
        out = {
            "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{ENV_ID} n={N} envs/gpu={E} obs={N_OBS} n_rays={N_RAYS} area={AREA} "
                                   f"T={T} rollout (configs[2])", "step": f"one {T}-step rollout of {E} envs per GPU",
                       "weights": "reference pretrained DoubleIntegrator gcbf+ (tests/golden fixture)",
                       "l2": "no flush: each rollout streams ~0.6 GB of trajectory records (> 126 MB L2)",
                       "deg_real": deg_real, "edges_per_step": n_edges},
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": eng.launches_per_run * args.steps,
            "clocks": clocks,
            "roofline": roof,
            "rollout_rooflines": {
                "hbm_frac_of_" + src: value * B_ALG / (hbm * 1e9 * world),
                "algorithmic_tflops_per_gpu": value / (N * E * world) * flop_step / 1e12 / world,
                "note": "whole rollout vs HBM (312 B/agent-step) and achieved TFLOP/s per GPU counting the reference's "
                        "unfolded F_edge/F_node (SURVEY 8d); the path is latency-bound (8 dependent launches per "
                        "env-step over 8192 agents), not HBM-bound"},
            "train_step": train,
            "cpu_baseline": cpu,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def train_step_bench(torch, dist, env, algo, eng, rank, world, args, max_over_ranks, barrier):
    """GCBF+ update_inner throughput: global minibatch of 256 graphs (N=512) drawn from the recorded
    rollout, B/world graphs per rank, u_qp := u_ref + 0.1 (QP labels are an input)."""
    from gcbfplus_b200.algo import train as T
    B_glob = 256
    B = max(B_glob // world, 1)
    ro = eng.result()
    tsel = torch.arange(B, device=env.device) % eng.T
    esel = torch.arange(B, device=env.device) % eng.E
    agent = eng.agent[tsel, esel].contiguous()
    hits = eng.hits[tsel, esel].contiguous()
    goal = eng.goal[esel].contiguous()
    old = env.edge_cap_per_agent
    env.edge_cap_per_agent = 4
    g = env.get_graph(agent, goal, None, hits=hits)
    env.edge_cap_per_agent = old
    obs_rep = ro.obstacle.select(esel.cpu().numpy()) if hasattr(ro.obstacle, "select") else None
    gm = g._replace(obstacle=obs_rep)
    unsafe = env.unsafe_mask(gm)
    safe = ~unsafe
    u_qp = env.u_ref(g) + 0.1
    algo._trainer_state = None
    for _ in range(2):
        T.train_minibatch(algo, g, safe, unsafe, u_qp, apply=True)
    torch.cuda.synchronize()
    g.check_overflow()
    n_it = 5
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = env.lib.gcbf_launch_count()
    barrier()
    ev0.record()
    for _ in range(n_it):
        T.train_minibatch(algo, g, safe, unsafe, u_qp, apply=True)
    ev1.record()
    barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1)) / n_it
    launches = (env.lib.gcbf_launch_count() - n0) // n_it
    n_edges = g.n_edge
    N = env.num_agents
    flops = 9.0 * (n_edges * F_EDGE + B * N * F_NODE) * world       # 3 passes x (fwd + ~2x bwd), SURVEY 8d
    return {"ms_per_minibatch": ms, "graphs_per_s": B * world / (ms * 1e-3),
            "agent_samples_per_s": B * world * N / (ms * 1e-3), "global_batch_graphs": B * world,
            "graphs_per_rank": B, "edges_per_rank": n_edges, "launches_per_step": int(launches),
            "approx_tflops_per_gpu": flops / world / (ms * 1e-3) / 1e12,
            "allreduce_bytes_per_step": 4 * (algo._trainer_state.packed.numel() + 4) if world > 1 else 0}


def gemm_roofline(torch, _lib, dev, n_edges: int, n_agents: int):
    """Dominant kernel = tc::gemm_tc_kernel (tcgen05 kind::tf32, 3xTF32 operand split, TMA + mbarrier pipeline,
    fp32 accumulators in TMEM): ~49% of a rollout step and ~45% of a train step.  Timed alone with CUDA events
    on the launching stream, L2 flushed between launches.  `achieved` counts ALGORITHMIC flops (2 M K N, one
    fp32-equivalent GEMM); the tensor pipe executes 3x that in TF32 MMAs."""
    hbm, tf_burst, tf_sus, src = load_peaks()
    lib = _lib.load()
    st = torch.cuda.current_stream(dev).cuda_stream
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)

    def time_tc(M, K, Nn, simt=False):
        A = torch.randn(M, K, device=dev)
        W = torch.randn(K, Nn, device=dev) * 0.05
        Bt = W.t().contiguous()
        Bh, Bl = torch.empty_like(Bt), torch.empty_like(Bt)
        _lib.check(lib.gcbf_split_tf32(Bt.data_ptr(), Bh.data_ptr(), Bl.data_ptr(), Bt.numel(), st), "split")
        b = torch.zeros(Nn, device=dev)
        Cout = torch.empty(M, Nn, device=dev)
        times = []
        for it in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if simt:
                _lib.check(lib.gcbf_gemm_nn(0, 0, A.data_ptr(), W.data_ptr(), b.data_ptr(), None, Cout.data_ptr(), None,
                                            None, M, M, K, Nn, st), "gcbf_gemm_nn")
            else:
                _lib.check(lib.gcbf_gemm_tc(0, 0, A.data_ptr(), Bh.data_ptr(), Bl.data_ptr(), b.data_ptr(), None,
                                            Cout.data_ptr(), None, None, M, M, K, Nn, st), "gcbf_gemm_tc")
            e1.record()
            torch.cuda.synchronize()
            if it >= 3:
                times.append(e0.elapsed_time(e1))
        return sum(times) / len(times)

    M, K, Nn = max(n_edges, 128), 256, 128                      # message layer of the rollout (edge rows)
    ms = time_tc(M, K, Nn)
    flops = 2.0 * M * K * Nn
    achieved = flops / (ms * 1e-3) / 1e12
    Mt = 200000                                                 # train-step sized launch (256 graphs x 512 agents)
    ms_t = time_tc(Mt, 256, 256)
    ach_t = 2.0 * Mt * 256 * 256 / (ms_t * 1e-3) / 1e12
    ms_s = time_tc(Mt, 256, 256, simt=True)
    alg_bytes = 4.0 * (M * K + K * Nn + M * Nn)
    return {"kernel": "tc::gemm_tc_kernel<128,EPI_BIAS> (tcgen05 3xTF32) M=%d K=256 N=128: folded message layer "
                      "over the rollout's edge rows" % M,
            "bound": "tensor", "achieved": achieved, "peak": tf_burst, "unit": "TFLOP/s", "frac": achieved / tf_burst,
            "peak_source": src + " bf16 cuBLAS burst (kernel timed alone)",
            "us_per_launch": ms * 1e3, "algorithmic_flops": flops, "algorithmic_bytes": alg_bytes, "traffic": None,
            "tf32_mma_tflops": 3 * achieved, "frac_of_tf32_peak": 3 * achieved / (tf_burst / 2),
            "train_shape": {"M": Mt, "K": 256, "N": 256, "us_per_launch": ms_t * 1e3, "achieved": ach_t,
                            "frac": ach_t / tf_burst, "tf32_mma_tflops": 3 * ach_t,
                            "frac_of_tf32_peak": 3 * ach_t / (tf_burst / 2),
                            "simt_fp32_kernel_tflops": 2.0 * Mt * 256 * 256 / (ms_s * 1e-3) / 1e12,
                            # dram__bytes_read.sum + dram__bytes_write.sum of this launch from the committed ncu
                            # --set full capture (profiles/r01_final_summary.md), not re-measured here
                            "traffic": 205.6e6 + 151.5e6, "algorithmic_bytes": 4.0 * (Mt * 256 + 256 * 256 + Mt * 256)},
            "note": "rollout launches are single-wave (110 row tiles on 148 SMs): latency-bound; the train-shape "
                    "line shows the kernel at scale.  TF32 dense peak taken as half the measured bf16 peak."}


# ======================================================================================== CPU arm
def cpu_baseline(args, steps: int = 1, verbose: bool = False):
    """Restated reference (dense padded formulation), torch-CPU fp32, all host threads, on a
    bounded sample: 1 env of the same workload (n=512, obs 8, 32 rays), `steps` env-steps."""
    import numpy as np
    import torch
    from helpers import oracle_env, oracle_params
    from oracle.algo import act
    from oracle.geometry import Rectangle
    ncpu = os.cpu_count() or 1
    rng = np.random.Generator(np.random.PCG64(0))
    ap, _ = oracle_params(ENV_ID)

    def probe(threads: int) -> float:
        """Seconds for one dense policy forward at n=128 with `threads` intra-op threads."""
        torch.set_num_threads(threads)
        n = 128
        e = oracle_env(ENV_ID, n, 16.0, 0, N_RAYS)
        ag = torch.zeros(n, 4)
        ag[:, :2] = torch.rand(n, 2) * 16.0
        with torch.no_grad():
            g = e.get_graph(ag, ag.flip(0).clone(), None)
            act(e, ap, g)
            t0 = time.perf_counter()
            act(e, ap, g)
            return time.perf_counter() - t0

    cands = sorted({c for c in (ncpu, ncpu // 2, 64, 32, 16, 8) if 1 <= c <= ncpu}, reverse=True)
    timings = {c: probe(c) for c in cands}
    cores = min(timings, key=timings.get)            # the thread count the restated reference runs fastest with
    torch.set_num_threads(cores)
    N = N_AGENTS
    oenv = oracle_env(ENV_ID, N, AREA, N_OBS, N_RAYS)
    obs = Rectangle.create(rng.uniform(0, AREA, (N_OBS, 2)), rng.uniform(0.1, 0.5, N_OBS),
                           rng.uniform(0.1, 0.5, N_OBS), rng.uniform(0, 2 * np.pi, N_OBS))
    agent = torch.zeros(N, 4)
    agent[:, :2] = torch.from_numpy(rng.uniform(0, AREA, (N, 2)).astype(np.float32))
    goal = torch.zeros(N, 4)
    goal[:, :2] = torch.from_numpy(rng.uniform(0, AREA, (N, 2)).astype(np.float32))
    times = []
    with torch.no_grad():
        g = oenv.get_graph(agent, goal, obs)
        for s in range(steps + 1):
            t0 = time.perf_counter()
            a = act(oenv, ap, g)                     # dense: all 2N^2 + NR padded edges, like the reference
            g, r, c = oenv.step(g, a)
            dt = time.perf_counter() - t0
            if s > 0 or steps == 0:
                times.append(dt)
    sec = sum(times) / max(len(times), 1)
    # same step with the masked edges dropped first (the CUDA path's formulation) -> separates the algorithmic
    # (dense -> sparse) factor from the hardware (CPU -> B200) one, BASELINE.md section 3
    sparse_times = []
    with torch.no_grad():
        for s in range(2):
            t0 = time.perf_counter()
            gs = oenv.sparsify(g)
            a = act(oenv, ap, gs)
            g, r, c = oenv.step(gs, a)
            if s > 0:
                sparse_times.append(time.perf_counter() - t0)
    sec_sparse = sum(sparse_times) / max(len(sparse_times), 1)
    return {"value": N / sec, "unit": "env-steps/s", "cores": cores, "kind": "port", "sparse_value": N / sec_sparse,
            "sample": f"1 env x n={N} x {len(times)} env-step(s) after 1 warm-up step, dense reference formulation "
                      f"({2 * N * N + N * N_RAYS} padded edges/graph), torch-CPU fp32; {sec:.2f} s per env-step; "
                      f"{cores} of {ncpu} host threads (fastest of {cands} on a n=128 probe)",
            "sec_per_env_step": sec}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    t0 = time.perf_counter()
    cpu = cpu_baseline(args, steps=max(1, min(args.steps, 3)))
    value = cpu["value"]
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": cpu["sec_per_env_step"] * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"{ENV_ID} n={N_AGENTS} obs={N_OBS} n_rays={N_RAYS} area={AREA} (configs[2]); "
                                  "bounded sample: 1 env, per step 1 env-step",
                      "note": "JAX/Flax/jraph are not installable in this image: the arm is the restated "
                              "reference (oracle, dense formulation), not the JAX code"},
           "cpu_baseline": {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "sparse_value")},
           "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--T", type=int, default=T_STEPS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
