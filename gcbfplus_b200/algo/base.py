"""Controller interface of the B200 path.

Same attribute and method names as the reference's abstract controller (gcbfplus/algo/base.py:10-68), because
train.py / test.py / Trainer are written against them: `node_dim`, `edge_dim`, `action_dim`, `n_agents`, `config`,
`actor_params`, `act`, `step`, `update`, `save`, `load`.  Tensors are torch CUDA fp32 with a leading graph-batch
dimension where the reference vmaps over single graphs."""
from ..env.base import MultiAgentEnv

_DIMS = ("node_dim", "edge_dim", "action_dim", "n_agents")


def _abstract(name: str):
    def method(self, *args, **kwargs):
        raise NotImplementedError(f"{type(self).__name__} must implement {name}()")
    method.__name__ = name
    return method


class MultiAgentController:
    """Holds the environment and the four sizes; everything else is the subclass's job."""

    def __init__(self, env: MultiAgentEnv, node_dim: int, edge_dim: int, action_dim: int, n_agents: int):
        self._env = env
        self._sizes = dict(zip(_DIMS, (int(node_dim), int(edge_dim), int(action_dim), int(n_agents))))

    # node_dim / edge_dim / action_dim / n_agents: read-only views of the constructor arguments
    def __getattr__(self, name):
        sizes = self.__dict__.get("_sizes", {})
        if name in sizes:
            return sizes[name]
        raise AttributeError(f"{type(self).__name__!s} has no attribute {name!r}")

    @property
    def config(self) -> dict:                     # hyper-parameters written next to the checkpoints (config.yaml)
        raise NotImplementedError

    @property
    def actor_params(self):                       # what rollout / test code hands back to act() / step()
        raise NotImplementedError

    # act(graph, params=None) -> action [G, N, nu];  step(graph, key, params=None) -> (action, log_pi)
    act = _abstract("act")
    step = _abstract("step")
    # update(rollout, step) -> info dict;  save / load(dir, step): <dir>/<step>/{actor,cbf}.pkl
    update = _abstract("update")
    save = _abstract("save")
    load = _abstract("load")

