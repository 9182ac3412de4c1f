"""Trainer -- gcbfplus/trainer/trainer.py:18-143 on the B200 rollout engine + CUDA train step.
wandb logging is optional (offline / absent -> metrics are printed and kept in `self.history`)."""
from __future__ import annotations

import os
from time import time

import numpy as np

from .rollout import RolloutEngine
from .utils import eval_metrics, rollout


class Trainer:

    def __init__(self, env, env_test, algo, n_env_train: int, n_env_test: int, log_dir: str, seed: int, params: dict,
                 save_log: bool = True):
        self.env = env
        self.env_test = env_test
        self.algo = algo
        self.n_env_train = n_env_train
        self.n_env_test = n_env_test
        self.log_dir = log_dir
        self.seed = seed
        if Trainer._check_params(params):
            self.params = params
        if save_log:
            os.makedirs(log_dir, exist_ok=True)
            self.model_dir = os.path.join(log_dir, "models")
            os.makedirs(self.model_dir, exist_ok=True)
        self.wandb = None
        if os.environ.get("WANDB_MODE", "disabled") != "disabled":
            try:
                import wandb
                wandb.init(name=params["run_name"], project="gcbf-b200", dir=self.log_dir)
                self.wandb = wandb
            except Exception as e:  # pragma: no cover
                print(f"wandb unavailable ({e}); logging to stdout only")
        self.save_log = save_log
        self.steps = params["training_steps"]
        self.eval_interval = params["eval_interval"]
        self.eval_epi = params["eval_epi"]
        self.save_interval = params["save_interval"]
        self.update_steps = 0
        self.rng = np.random.Generator(np.random.PCG64(seed))         # trainer.py:62 key stream
        self.history = []

    @staticmethod
    def _check_params(params: dict) -> bool:
        """trainer/trainer.py:64-74."""
        assert "run_name" in params, "run_name not found in params"
        assert "training_steps" in params, "training_steps not found in params"
        assert "eval_interval" in params, "eval_interval not found in params"
        assert params["eval_interval"] > 0, "eval_interval must be positive"
        assert "eval_epi" in params, "eval_epi not found in params"
        assert params["eval_epi"] >= 1, "eval_epi must be greater than or equal to 1"
        assert "save_interval" in params, "save_interval not found in params"
        assert params["save_interval"] > 0, "save_interval must be positive"
        return True

    def _log(self, info: dict) -> None:
        self.history.append(dict(info, update_steps=self.update_steps))
        if self.wandb is not None:
            self.wandb.log(info, step=self.update_steps)

    def train(self):
        """trainer/trainer.py:76-143."""
        start_time = time()
        train_engine = RolloutEngine(self.env, self.n_env_train)
        test_engine = RolloutEngine(self.env_test, self.n_env_test)
        test_seed = int(np.random.Generator(np.random.PCG64(self.seed)).integers(0, 2 ** 31 - 1))   # fixed test keys
        for step in range(0, self.steps + 1):
            if step % self.eval_interval == 0:
                ro = rollout(self.env_test, test_engine, self.algo.actor_params, test_seed)
                info = eval_metrics(self.env_test, ro)
                eval_info = {k: v for k, v in info.items() if k.startswith("eval/")}
                eval_info["step"] = step
                self._log(eval_info)
                print(f"step: {step:3}, time: {time() - start_time:5.0f}s, reward: {info['eval/reward']:9.4f}, "
                      f"min/max reward: {info['reward_min']:7.2f}/{info['reward_max']:7.2f}, "
                      f"cost: {info['eval/cost']:8.4f}, unsafe_frac: {info['eval/unsafe_frac']:6.2f}, "
                      f"finish: {info['eval/finish']:6.2f}")
                if self.save_log and step % self.save_interval == 0:
                    self.algo.save(os.path.join(self.model_dir), step)
            key = int(self.rng.integers(0, 2 ** 31 - 1))
            ro = rollout(self.env, train_engine, self.algo.actor_params, key)
            update_info = self.algo.update(ro, step)
            self._log(update_info)
            self.update_steps += 1
