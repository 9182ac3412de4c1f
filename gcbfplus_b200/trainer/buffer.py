"""MaskedReplayBuffer -- gcbfplus/trainer/buffer.py:57-93 kept on the device.

Same semantics as the reference (FIFO of the last `size` rollouts / unsafe graphs, uniform sampling
with replacement from NumPy's global RNG, gcbf_plus.py:232-280) on compact per-graph records
{agent, hits, goal, safe, unsafe} instead of dense GraphsTuples."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch


class MaskedReplayBuffer:

    def __init__(self, size: int):
        self._size = size
        self._data: Optional[Dict[str, torch.Tensor]] = None   # per-graph arrays, rollout-major
        self._T = 1                                            # graphs per stored item (T for rollouts, 1 for graphs)

    # ---- rollouts (buffer.append(rollout, safe, unsafe), buffer.py:66-80)
    def append_rollouts(self, flat: Dict[str, torch.Tensor], b: int, T: int) -> None:
        self._T = T
        self._append(flat, self._size * T)

    # ---- single graphs (unsafe_buffer.append(rollout[unsafe_multi_mask], ...), gcbf_plus.py:246-251)
    def append_graphs(self, flat: Dict[str, torch.Tensor], mask: torch.Tensor) -> None:
        self._T = 1
        sel = torch.nonzero(mask.reshape(-1), as_tuple=False).squeeze(-1)
        self._append({k: v[sel] for k, v in flat.items()}, self._size)

    def _append(self, flat: Dict[str, torch.Tensor], cap_graphs: int) -> None:
        if self._data is None:
            self._data = {k: v.clone() for k, v in flat.items()}
        else:
            self._data = {k: torch.cat([self._data[k], flat[k]], dim=0) for k in flat}
        n = next(iter(self._data.values())).shape[0]
        if n > cap_graphs:
            self._data = {k: v[-cap_graphs:].contiguous() for k, v in self._data.items()}

    @property
    def n_items(self) -> int:
        if self._data is None:
            return 0
        return next(iter(self._data.values())).shape[0] // self._T

    @property
    def length(self) -> int:
        """Number of stored graphs (`n_data` of buffer.py:89-93)."""
        return 0 if self._data is None else int(next(iter(self._data.values())).shape[0])

    def sample_rollouts(self, batch_size: int) -> Dict[str, torch.Tensor]:
        """buffer.py:82-85: `batch_size` whole rollouts, with replacement -> batch_size*T graphs."""
        idx = np.random.randint(0, self.n_items, batch_size)
        gi = (idx[:, None] * self._T + np.arange(self._T)[None, :]).reshape(-1)
        return self.get_data(gi)

    def sample_graphs(self, n: int) -> Dict[str, torch.Tensor]:
        idx = np.random.randint(0, self.length, n)
        return self.get_data(idx)

    def get_data(self, idx: np.ndarray) -> Dict[str, torch.Tensor]:
        dev = next(iter(self._data.values())).device
        t = torch.from_numpy(np.asarray(idx, dtype=np.int64)).to(dev)
        return {k: v[t] for k, v in self._data.items()}
