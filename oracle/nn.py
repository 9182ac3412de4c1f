"""Oracle networks (TEST INFRASTRUCTURE ONLY): the GNN layer, CBF net and
deterministic policy restated from gcbfplus/nn/gnn.py:44-104, nn/mlp.py:6-30,
algo/module/cbf.py:12-53, algo/module/policy.py:63-128.

Parameters are nested dicts with the reference's flax names:
``params/GNN_0/GNNLayer_0/{msg/Dense_0,msg/Dense_1,Dense_0,attn/Dense_0,attn/Dense_1,
Dense_1,update/Dense_0,update/Dense_1,Dense_2}``, ``CBFHead|PolicyHead/{Dense_0,Dense_1}``,
``Dense_0`` (cbf out) / ``OutputDense`` (actor out); each ``{kernel [in,out], bias [out]}``.
Third-party semantics restated from their public definitions:
flax ``nn.Dense``: y = x @ kernel + bias; jraph ``segment_softmax``:
exp(x - segment_max) / segment_sum; ``segment_sum``: scatter-add.
"""
from __future__ import annotations

import math
import pickle
from typing import Dict

import numpy as np
import torch

from .envs import Graph


# --------------------------------------------------------------------------- parameter IO
class _RefUnpickler(pickle.Unpickler):
    """Loads the reference's pickled jax.Array leaves without JAX."""

    def find_class(self, module, name):
        if module.startswith("jax") and name == "_reconstruct_array":
            def rec(fun, args, arr_state, aval_state):
                arr = fun(*args)
                arr.__setstate__(arr_state)
                return arr
            return rec
        if module.startswith("numpy.core"):
            module = module.replace("numpy.core", "numpy._core")
        return super().find_class(module, name)


def load_ref_pickle(path: str) -> dict:
    with open(path, "rb") as f:
        return _RefUnpickler(f).load()


def flatten_params(tree: dict, prefix: str = "") -> Dict[str, np.ndarray]:
    out = {}
    for k, v in tree.items():
        if isinstance(v, dict):
            out.update(flatten_params(v, prefix + k + "/"))
        else:
            out[prefix + k] = np.asarray(v)
    return out


def unflatten_params(flat: Dict[str, np.ndarray]) -> dict:
    tree: dict = {}
    for k, v in flat.items():
        parts = k.split("/")
        d = tree
        for p in parts[:-1]:
            d = d.setdefault(p, {})
        d[parts[-1]] = v
    return tree


def layer_specs(edge_dim: int, out_dim: int, kind: str):
    """(flax path, in, out) in forward order for one network (kind = 'cbf' | 'actor')."""
    g = "params/GNN_0/GNNLayer_0/"
    head = "CBFHead" if kind == "cbf" else "PolicyHead"
    last = "Dense_0" if kind == "cbf" else "OutputDense"
    return [
        (g + "msg/Dense_0", edge_dim + 6, 256), (g + "msg/Dense_1", 256, 256), (g + "Dense_0", 256, 128),
        (g + "attn/Dense_0", 128, 128), (g + "attn/Dense_1", 128, 128), (g + "Dense_1", 128, 1),
        (g + "update/Dense_0", 131, 256), (g + "update/Dense_1", 256, 256), (g + "Dense_2", 256, 128),
        (f"params/{head}/Dense_0", 128, 256), (f"params/{head}/Dense_1", 256, 256),
        (f"params/{last}", 256, out_dim),
    ]


def init_params(edge_dim: int, out_dim: int, kind: str, seed: int) -> dict:
    """xavier_uniform kernels, zero biases (nn/utils.py:21; flax Dense default bias init).
    NumPy PCG64 stream (the reference's jax.random stream is not reproducible here)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    flat = {}
    for path, fi, fo in layer_specs(edge_dim, out_dim, kind):
        lim = math.sqrt(6.0 / (fi + fo))
        flat[path + "/kernel"] = rng.uniform(-lim, lim, size=(fi, fo)).astype(np.float32)
        flat[path + "/bias"] = np.zeros((fo,), dtype=np.float32)
    return unflatten_params(flat)


def to_torch(params: dict, dtype=torch.float32, requires_grad: bool = False) -> Dict[str, torch.Tensor]:
    flat = flatten_params(params)
    return {k: torch.tensor(v, dtype=dtype, requires_grad=requires_grad) for k, v in flat.items()}


# --------------------------------------------------------------------------- forward
def _dense(p, path, x):
    return x @ p[path + "/kernel"] + p[path + "/bias"]


def gnn_layer(p: Dict[str, torch.Tensor], nodes, edges, senders, receivers):
    """nn/gnn.py:22-75 for one layer.  Returns new node features [n_nodes, 128]."""
    g = "params/GNN_0/GNNLayer_0/"
    n_nodes = nodes.shape[0]
    feats = torch.cat([edges, nodes[senders], nodes[receivers]], dim=-1)           # gnn.py:54
    x = torch.relu(_dense(p, g + "msg/Dense_0", feats))                            # MLP(256,256), act_final=False
    x = _dense(p, g + "msg/Dense_1", x)
    msg = _dense(p, g + "Dense_0", x)                                              # gnn.py:56 -> 128
    gf = torch.relu(_dense(p, g + "attn/Dense_0", msg))                            # gnn.py:66
    gf = _dense(p, g + "attn/Dense_1", gf)
    gate = _dense(p, g + "Dense_1", gf).squeeze(-1)                                # gnn.py:67
    seg_max = torch.full((n_nodes,), -float("inf"), dtype=gate.dtype)
    seg_max = seg_max.scatter_reduce(0, receivers, gate.detach(), reduce="amax", include_self=True)
    ex = torch.exp(gate - seg_max[receivers])
    denom = torch.zeros(n_nodes, dtype=gate.dtype).index_add(0, receivers, ex)
    attn = ex / denom[receivers]                                                   # segment_softmax
    aggr = torch.zeros(n_nodes, msg.shape[1], dtype=msg.dtype).index_add(0, receivers, attn[:, None] * msg)
    u = torch.cat([nodes, aggr], dim=-1)                                           # gnn.py:60
    u = torch.relu(_dense(p, g + "update/Dense_0", u))
    u = _dense(p, g + "update/Dense_1", u)
    return _dense(p, g + "Dense_2", u)                                             # gnn.py:62


def net_forward(p: Dict[str, torch.Tensor], graph: Graph, kind: str) -> torch.Tensor:
    """CBFNet (cbf.py:12-21) or Deterministic (policy.py:63-73): GNN -> agent rows
    (type_nodes(0, n_agents), graph.py:112-124) -> head MLP -> tanh(Dense)."""
    x = gnn_layer(p, graph.nodes.to(p["params/GNN_0/GNNLayer_0/Dense_0/bias"].dtype),
                  graph.edges, graph.senders, graph.receivers)
    x = x[: graph.n_agents]
    head = "CBFHead" if kind == "cbf" else "PolicyHead"
    last = "Dense_0" if kind == "cbf" else "OutputDense"
    x = torch.relu(_dense(p, f"params/{head}/Dense_0", x))
    x = _dense(p, f"params/{head}/Dense_1", x)
    return torch.tanh(_dense(p, f"params/{last}", x))
