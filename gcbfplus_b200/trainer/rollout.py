"""Batched closed-loop rollout -- replaces jit(vmap(rollout)) of gcbfplus/trainer/utils.py:25-55
and trainer/trainer.py:81-87.

Two device paths with the same arithmetic (bit-identical results, tests/test_gpu_rollout.py):
* persistent (default where supported: 2-D environments, n <= 512): ONE kernel launch for the whole T-step rollout, one
  thread-block cluster per environment looping over the steps (csrc/rollout_persist.cu);
* 5-launch env-step (gcbf_rollout_step), the whole T-step loop captured in one CUDA graph: LinearDrone, n > 512,
  GCBF_PERSISTENT=0, the u_ref policy.
No host sync inside the loop on either path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from .. import _lib
from ..algo.params import NetParams
from ..utils.graph import SwarmGraph
from .data import Rollout


class _Chain:
    """Per-chain scratch: a contiguous slice [e0, e1) of the environments with its own topology arrays,
    policy output and activation workspace, so that chains are independent branches of the CUDA graph."""

    def __init__(self, eng: "RolloutEngine", e0: int, e1: int):
        env, dev = eng.env, eng.env.device
        self.e0, self.e1 = e0, e1
        E, N, nu = e1 - e0, env.num_agents, env.action_dim
        f32, i32 = torch.float32, torch.int32
        self.desc = env.desc(E, eng.O)
        cap = self.desc.edge_cap
        self.pi = torch.zeros(E, N, nu, dtype=f32, device=dev)
        # edge lists are double-buffered: step t reads half t % 2 while the graph of state t+1 is written into the
        # other half (gcbf_rollout_step's fused tail + graph build reads the old lists for the step cost)
        self.row_start = torch.zeros(2, E * N, dtype=i32, device=dev)
        self.row_deg = torch.zeros(2, E * N, dtype=i32, device=dev)
        self.edge_recv = torch.zeros(2, cap, dtype=i32, device=dev)
        self.edge_src = torch.zeros(2, cap, dtype=i32, device=dev)
        self.counters = torch.zeros(eng.T + 1, 4, dtype=i32, device=dev)
        n_ws = env.lib.gcbf_rollout_workspace_floats(C.byref(self.desc))
        self.ws = torch.empty(int(n_ws), dtype=f32, device=dev)
        self.stream = None


class RolloutEngine:
    def __init__(self, env, n_envs: int, T: Optional[int] = None, n_obs: Optional[int] = None,
                 use_cuda_graph: bool = True, policy: str = "actor", n_chains: Optional[int] = None,
                 persistent: Optional[bool] = None):
        """policy: 'actor' (a = 2 pi + u_ref, algo.step) or 'u_ref' (test.py --u-ref).
        n_chains: the environments are split into independent chains that run as parallel branches of
        the CUDA graph (each per-step kernel is latency-bound and fills a fraction of the 148 SMs, so
        concurrent chains overlap their launch / tail latencies).  Results do not depend on it."""
        self.env = env
        self.E = n_envs
        self.T = T or env.max_episode_steps
        self.O = env.params["n_obs"] if n_obs is None else n_obs
        self.policy = policy
        self.use_cuda_graph = use_cuda_graph
        dev = env.device
        E, T, N = self.E, self.T, env.num_agents
        sd, nu, R, pd = env.state_dim, env.action_dim, env.n_hits, env.pos_dim
        f32, i32 = torch.float32, torch.int32
        self.agent = torch.zeros(T + 1, E, N, sd, dtype=f32, device=dev)
        self.hits = torch.zeros(T + 1, E, N, R, pd, dtype=f32, device=dev)
        self.goal = torch.zeros(E, N, sd, dtype=f32, device=dev)
        self.obs_w = 16 if pd == 2 else 4
        self.obstacles = torch.zeros(E, max(self.O, 1), self.obs_w, dtype=f32, device=dev)
        self.actions = torch.zeros(T, E, N, nu, dtype=f32, device=dev)
        self.rewards = torch.zeros(T, E, dtype=f32, device=dev)
        self.costs = torch.zeros(T, E, dtype=f32, device=dev)
        if n_chains is None:
            n_chains = 1   # measured: no gain at fixed E (each chain's step latency does not shrink with its batch)
        if not use_cuda_graph:
            n_chains = 1
        assert E % n_chains == 0
        per = E // n_chains
        self.chains = [_Chain(self, c * per, (c + 1) * per) for c in range(n_chains)]
        self.desc = self.chains[0].desc
        self.params_buf = torch.zeros(_lib.param_count(env.edge_dim, nu), dtype=f32, device=dev)
        # folded inference weights (gcbf_prepare_infer), rebuilt by set_params()
        self.infer_blob = torch.zeros(int(env.lib.gcbf_infer_count(env.edge_dim, nu)), dtype=f32, device=dev)
        self.use_tc = 1 if _lib.USE_TC else 0
        # persistent single-launch rollout (csrc/rollout_persist.cu) where the library supports the configuration
        # the persistent kernel splits the edge capacity evenly over the environments (and, in pair mode, over the pairs
        # of an environment), so it gets a roomier descriptor than the pooled lists of the 5-launch path: 48 rows per
        # agent (or the worst case 1 + (N - 1) + R when that is smaller)
        per_agent = min(1 + (N - 1) + R, max(2 * env.edge_cap_per_agent, 48))
        self._pdesc = env.desc(E, self.O, edge_cap=E * N * per_agent)
        level = int(env.lib.gcbf_rollout_persistent_supported(C.byref(self._pdesc))) \
            if (policy == "actor" and self.use_tc and len(self.chains) == 1) else 0
        ok = level > 0
        if persistent is None:
            # default only where every environment's cluster is resident at once (level 2): on a B200 at most 15
            # clusters of 8 CTAs fit, so 16 environments of n = 512 would run in two rounds (measured: 134 vs 76 us / step)
            # DubinsCar is opt-in: its persistent rollout agrees with the 5-launch path only to closed-loop rounding.
            # profiles/r02_dubins_persistent_vs_5launch.log: states and actions are identical while every speed is 0
            # (step 0); from the first step with v != 0 a few policy outputs differ by 1-2 ulp (1.8e-7) -- the
            # (v cos th, v sin th) edge features are evaluated in two translation units -- which the closed loop
            # amplifies to 1.6e-5 after 24 steps.  Env-sharded runs must not mix two numeric paths.
            persistent = (level == 2 and env.ENV_ID != "DubinsCar" and os.environ.get("GCBF_PERSISTENT", "1") != "0")
        if persistent and not ok:
            raise ValueError("persistent rollout unsupported for this configuration (2-D env, n <= 512, tensor-core path, "
                             "actor policy, one chain)")
        self.persistent = bool(persistent)
        #: optional [T + 1, 8] int64 device tensor: in-kernel %globaltimer stamps of the persistent rollout (set before
        #: the first run(); see gcbf_rollout_persistent in include/gcbf_b200.h)
        self.phase_stamps: Optional[torch.Tensor] = None
        self._pws = None
        if self.persistent:
            n = env.lib.gcbf_rollout_persistent_workspace_floats(C.byref(self._pdesc))
            self._pws = torch.empty(int(n), dtype=f32, device=dev)
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self.launches_per_run = 0
        self._obstacle_obj = None

    @property
    def counters(self) -> torch.Tensor:
        """[T+1, 4]: per step total edge count (col 0) and overflow flag (col 1) over all chains."""
        c = torch.stack([ch.counters for ch in self.chains])
        return torch.stack([c[:, :, 0].sum(0), c[:, :, 1].amax(0), c[:, :, 2].sum(0), c[:, :, 3].sum(0)], dim=1)

    # ------------------------------------------------------------------ one env step (enqueue only)
    def _build(self, ch: _Chain, t: int, stream: int) -> None:
        env, d = self.env, ch.desc
        rc = env.lib.gcbf_graph_build(C.byref(d), self.agent[t, ch.e0].data_ptr(),
                                      self.obstacles[ch.e0].data_ptr() if self.O > 0 else None,
                                      env.ray_table.data_ptr(), self.hits[t, ch.e0].data_ptr(),
                                      ch.row_start[t % 2].data_ptr(), ch.row_deg[t % 2].data_ptr(),
                                      ch.edge_recv[t % 2].data_ptr(), ch.edge_src[t % 2].data_ptr(),
                                      ch.counters[t].data_ptr(), 1, stream)
        _lib.check(rc, "gcbf_graph_build")

    def _step(self, ch: _Chain, t: int, stream: int) -> None:
        env, d = self.env, ch.desc
        obs = self.obstacles[ch.e0].data_ptr() if self.O > 0 else None
        b = t % 2
        if self.policy == "actor":      # algo.step + env.step + get_graph(next) in one call (6 launches)
            rc = env.lib.gcbf_rollout_step(
                C.byref(d), self.params_buf.data_ptr(), self.infer_blob.data_ptr(), self.use_tc,
                self.agent[t, ch.e0].data_ptr(), self.goal[ch.e0].data_ptr(), obs, env.ray_table.data_ptr(),
                self.hits[t, ch.e0].data_ptr(), ch.row_start[b].data_ptr(), ch.row_deg[b].data_ptr(),
                ch.edge_recv[b].data_ptr(), ch.edge_src[b].data_ptr(), ch.counters[t].data_ptr(),
                self.actions[t, ch.e0].data_ptr(), self.agent[t + 1, ch.e0].data_ptr(),
                self.hits[t + 1, ch.e0].data_ptr(), ch.row_start[1 - b].data_ptr(), ch.row_deg[1 - b].data_ptr(),
                ch.edge_recv[1 - b].data_ptr(), ch.edge_src[1 - b].data_ptr(), ch.counters[t + 1].data_ptr(),
                self.rewards[t, ch.e0:].data_ptr(), self.costs[t, ch.e0:].data_ptr(), ch.ws.data_ptr(), ch.ws.numel(),
                stream)
            _lib.check(rc, "gcbf_rollout_step")
            return
        rc = env.lib.gcbf_env_step(C.byref(d), self.agent[t, ch.e0].data_ptr(), self.goal[ch.e0].data_ptr(), obs,
                                   None, ch.row_start[b].data_ptr(), ch.row_deg[b].data_ptr(), ch.edge_src[b].data_ptr(),
                                   self.actions[t, ch.e0].data_ptr(), self.agent[t + 1, ch.e0].data_ptr(),
                                   self.rewards[t, ch.e0:].data_ptr(), self.costs[t, ch.e0:].data_ptr(), 2, stream)
        _lib.check(rc, "gcbf_env_step")
        self._build(ch, t + 1, stream)

    def _enqueue_chain(self, ch: _Chain, stream: int) -> None:
        self._build(ch, 0, stream)
        for t in range(self.T):
            self._step(ch, t, stream)

    def _enqueue_persistent(self, n_steps: int, stream: int) -> None:
        env, ch = self.env, self.chains[0]
        rc = env.lib.gcbf_rollout_persistent(
            C.byref(self._pdesc), int(n_steps), self.params_buf.data_ptr(), self.infer_blob.data_ptr(), self.goal.data_ptr(),
            self.obstacles.data_ptr() if self.O > 0 else None, env.ray_table.data_ptr(), self.agent.data_ptr(),
            self.hits.data_ptr(), self.actions.data_ptr(), self.rewards.data_ptr(), self.costs.data_ptr(),
            ch.counters.data_ptr(), self._pws.data_ptr(), self._pws.numel(),
            self.phase_stamps.data_ptr() if self.phase_stamps is not None else None, stream)
        _lib.check(rc, "gcbf_rollout_persistent")

    def _enqueue_all(self) -> None:
        dev = self.env.device
        main = torch.cuda.current_stream(dev)
        if self.persistent:
            self._enqueue_persistent(self.T, main.cuda_stream)
            return
        if len(self.chains) == 1:
            self._enqueue_chain(self.chains[0], main.cuda_stream)
            return
        for ch in self.chains:                              # fork: parallel branches
            if ch.stream is None:
                ch.stream = torch.cuda.Stream(dev)
            ch.stream.wait_stream(main)
            with torch.cuda.stream(ch.stream):
                self._enqueue_chain(ch, ch.stream.cuda_stream)
        for ch in self.chains:                              # join
            main.wait_stream(ch.stream)

    # ------------------------------------------------------------------ public
    def set_initial(self, agent0: torch.Tensor, goal: torch.Tensor, obstacle) -> None:
        """Initial conditions: [E,N,sd] x2 (device or pinned host) + obstacle container."""
        self.agent[0].copy_(agent0.reshape(self.agent[0].shape), non_blocking=True)
        self.goal.copy_(goal.reshape(self.goal.shape), non_blocking=True)
        if self.O > 0:
            packed = obstacle.packed if hasattr(obstacle, "packed") else obstacle
            self.obstacles.copy_(packed.reshape(self.obstacles.shape), non_blocking=True)
        self._obstacle_obj = obstacle

    def set_params(self, params: NetParams) -> None:
        self.params_buf.copy_(params.flat, non_blocking=True)
        env = self.env
        _lib.check(env.lib.gcbf_prepare_infer(env.edge_dim, env.action_dim, _lib.ptr(self.params_buf),
                                              _lib.ptr(self.infer_blob), env._stream()), "gcbf_prepare_infer")

    def run(self, check: bool = True) -> None:
        """Run the T-step rollout from the current initial conditions (async)."""
        dev = self.env.device
        for ch in self.chains:
            ch.counters.zero_()
        lib = self.env.lib
        if self.use_cuda_graph:
            if self._graph is None:
                # warm-up outside capture (lazy module load, function attributes)
                st = torch.cuda.current_stream(dev).cuda_stream
                if self.persistent:
                    self._enqueue_persistent(min(self.T, 1), st)
                    self.chains[0].counters.zero_()
                else:
                    for ch in self.chains:
                        self._build(ch, 0, st)
                        self._step(ch, 0, st)
                torch.cuda.synchronize(dev)
                g = torch.cuda.CUDAGraph()
                n0 = lib.gcbf_launch_count()
                with torch.cuda.graph(g):
                    self._enqueue_all()
                self.launches_per_run = int(lib.gcbf_launch_count() - n0)
                self._graph = g
            self._graph.replay()
        else:
            n0 = lib.gcbf_launch_count()
            self._enqueue_all()
            self.launches_per_run = int(lib.gcbf_launch_count() - n0)
        if check:
            self.check_overflow()

    def check_overflow(self) -> None:
        c = self.counters.cpu()
        if int(c[:, 1].max()) != 0:
            cap = self._pdesc.edge_cap if self.persistent else self.desc.edge_cap
            raise RuntimeError(f"edge capacity overflow during rollout: up to {int(c[:, 0].max())} edges; edge_cap={cap}"
                               + (" split evenly over the environments / CTA pairs (persistent kernel)" if self.persistent
                                  else " per chain") + "; raise env.edge_cap_per_agent")

    def result(self) -> Rollout:
        """trainer/data.py Rollout in the reference's (b, T) order (views/transposes of the record)."""
        dones = torch.zeros(self.E, self.T, dtype=torch.bool, device=self.env.device)
        return Rollout(agent=self.agent.transpose(0, 1), goal=self.goal, hits=self.hits.transpose(0, 1),
                       obstacle=self._obstacle_obj, actions=self.actions.transpose(0, 1),
                       rewards=self.rewards.transpose(0, 1), costs=self.costs.transpose(0, 1), dones=dones,
                       log_pis=None, n_edges=self.counters[:, 0])
