"""Network parameters: flat fp32 device buffers <-> the reference's nested flax dict.

Layout of one network (CBF or actor) = the 12 Dense layers in forward order
(gcbfplus/nn/gnn.py:44-104, algo/module/cbf.py:12-53, algo/module/policy.py:63-128;
names per SURVEY A.3), kernel [in, out] row-major then bias, each 16-byte aligned;
offsets come from libgcbf_b200 (gcbf_param_offsets) so C and Python cannot drift.
Checkpoints keep the reference format: pickle of {'params': nested dict} with NumPy leaves
(gcbfplus/algo/gcbf.py:344-357); the reference's own pickles (jax.Array leaves) load
through a stub unpickler, no JAX needed.
"""
from __future__ import annotations

import math
import pickle
from typing import Dict, List, Tuple

import numpy as np
import torch

from .. import _lib


def layer_specs(edge_dim: int, out_dim: int, kind: str) -> List[Tuple[str, int, int]]:
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


def flatten_tree(tree: dict, prefix: str = "") -> Dict[str, np.ndarray]:
    out = {}
    for k, v in tree.items():
        if isinstance(v, dict):
            out.update(flatten_tree(v, prefix + k + "/"))
        else:
            out[prefix + k] = np.asarray(v)
    return out


def unflatten_tree(flat: Dict[str, np.ndarray]) -> dict:
    tree: dict = {}
    for k, v in flat.items():
        parts = k.split("/")
        d = tree
        for p in parts[:-1]:
            d = d.setdefault(p, {})
        d[parts[-1]] = v
    return tree


class _RefUnpickler(pickle.Unpickler):
    """Reads reference checkpoints whose leaves are pickled jax.Array objects."""

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


def load_pickle(path: str) -> dict:
    with open(path, "rb") as f:
        return _RefUnpickler(f).load()


class NetParams:
    """One network's parameters as a flat fp32 device buffer."""

    def __init__(self, edge_dim: int, out_dim: int, kind: str, device="cuda"):
        assert kind in ("cbf", "actor")
        self.edge_dim, self.out_dim, self.kind = edge_dim, out_dim, kind
        self.specs = layer_specs(edge_dim, out_dim, kind)
        self.offsets = _lib.param_offsets(edge_dim, out_dim)
        self.count = _lib.param_count(edge_dim, out_dim)
        self.flat = torch.zeros(self.count, dtype=torch.float32, device=device)
        self._flat_t = None   # transposed GEMM weights for the tensor-core path (gcbf_prepare_params)

    def prepared(self, stream: int = None):
        """Transposed GEMM weights (K-major B operands of the tcgen05 path), recomputed from `flat`.
        Returns None when the tensor-core path is disabled (GCBF_TENSOR_CORES=0)."""
        if not _lib.USE_TC:
            return None
        lib = _lib.load()
        if self._flat_t is None:
            n = lib.gcbf_params_t_count(self.edge_dim, self.out_dim)
            self._flat_t = torch.empty(int(n), dtype=torch.float32, device=self.flat.device)
        if stream is None:
            stream = torch.cuda.current_stream(self.flat.device).cuda_stream
        _lib.check(lib.gcbf_prepare_params(self.edge_dim, self.out_dim, _lib.ptr(self.flat), _lib.ptr(self._flat_t),
                                           stream), "gcbf_prepare_params")
        return self._flat_t

    # ---- init (nn/utils.py:21 xavier_uniform kernels, zero biases) ----
    def init_xavier(self, seed: int) -> "NetParams":
        rng = np.random.Generator(np.random.PCG64(seed))
        host = np.zeros(self.count, dtype=np.float32)
        for i, (_, fi, fo) in enumerate(self.specs):
            lim = math.sqrt(6.0 / (fi + fo))
            w = rng.uniform(-lim, lim, size=(fi, fo)).astype(np.float32)
            host[self.offsets[2 * i]: self.offsets[2 * i] + fi * fo] = w.reshape(-1)
        self.flat.copy_(torch.from_numpy(host))
        return self

    # ---- nested dict <-> flat ----
    def from_tree(self, tree: dict) -> "NetParams":
        flat = flatten_tree(tree)
        host = np.zeros(self.count, dtype=np.float32)
        for i, (path, fi, fo) in enumerate(self.specs):
            w = np.asarray(flat[path + "/kernel"], dtype=np.float32)
            b = np.asarray(flat[path + "/bias"], dtype=np.float32)
            if w.shape != (fi, fo) or b.shape != (fo,):
                raise ValueError(f"{path}: expected kernel {(fi, fo)}, got {w.shape}")
            host[self.offsets[2 * i]: self.offsets[2 * i] + fi * fo] = w.reshape(-1)
            host[self.offsets[2 * i + 1]: self.offsets[2 * i + 1] + fo] = b
        self.flat.copy_(torch.from_numpy(host))
        return self

    def to_tree(self) -> dict:
        host = self.flat.detach().cpu().numpy()
        flat = {}
        for i, (path, fi, fo) in enumerate(self.specs):
            flat[path + "/kernel"] = host[self.offsets[2 * i]: self.offsets[2 * i] + fi * fo].reshape(fi, fo).copy()
            flat[path + "/bias"] = host[self.offsets[2 * i + 1]: self.offsets[2 * i + 1] + fo].copy()
        return unflatten_tree(flat)

    def n_real(self) -> int:
        return sum(fi * fo + fo for _, fi, fo in self.specs)

    def clone(self) -> "NetParams":
        out = NetParams(self.edge_dim, self.out_dim, self.kind, device=self.flat.device)
        out.flat.copy_(self.flat)
        return out

    def save(self, path: str) -> None:
        with open(path, "wb") as f:
            pickle.dump(self.to_tree(), f)

    def load(self, path: str) -> "NetParams":
        return self.from_tree(load_pickle(path))
