"""Word prototypes and minimum-jerk trajectories on the GPU.

Mirror of the geometry in src/shared/keyboard.py: ``QWERTYKeyboard`` (key centres :654-673, ``get_word_prototype``
:710-765, ``get_minimum_jerk_trajectory`` :821-864) and ``generate_minimum_jerk_trajectory`` (:389-514), batched over
words: one thread block per word in csrc/keyboard.cu through the C ABI (wgg_word_prototypes, wgg_minimum_jerk).  The key
layout table is host arithmetic (26 entries); the random key / midpoint offsets are drawn on the host with
``np.random.normal`` in the reference's order, so ``np.random.seed(s)`` reproduces the reference's trajectories.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib


@dataclass
class KeyboardConfig:
    """Field names and defaults of src/shared/config.py:101-113."""
    width: float = 1.0
    height: float = 1.0
    rows: Tuple[str, ...] = ("qwertyuiop", "asdfghjkl", "zxcvbnm")
    row_offsets: Tuple[float, ...] = (0.0, 0.05, 0.15)
    key_width: float = 0.1
    key_height: float = 0.333


class QWERTYKeyboard:
    def __init__(self, config: Optional[KeyboardConfig] = None, device="cuda"):
        self.config = config or KeyboardConfig()
        self.device = torch.device(device)
        self.key_centers = self._compute_key_centers()

    def _compute_key_centers(self) -> Dict[str, Tuple[float, float]]:
        """keyboard.py:654-673: keys span x in [-0.9, 0.9] (minus the row offset), rows at y = -2/3, 0, 2/3."""
        centers = {}
        rows = self.config.rows
        for r, (row, off) in enumerate(zip(rows, self.config.row_offsets)):
            y = -1 + (r + 0.5) * (2.0 / len(rows))
            span, start = 1.8 - off, -0.9 + off / 2
            for i, key in enumerate(row):
                centers[key.lower()] = (start + (i + 0.5) * (span / len(row)), y)
        return centers

    def _key_positions(self, word: str) -> List[Tuple[float, float]]:
        return [self.key_centers[c] for c in word.lower() if c in self.key_centers]   # keyboard.py:679-686

    def _pack(self, words: Sequence[str], limit: int):
        pos = [self._key_positions(w) for w in words]
        maxk = max([len(p) for p in pos] + [1])
        if maxk > limit:
            raise ValueError(f"a word has {maxk} keys; the kernel handles at most {limit}")
        keys = np.zeros((len(words), maxk, 2), np.float64)
        nk = np.zeros(len(words), np.int32)
        for i, p in enumerate(pos):
            nk[i] = len(p)
            if p:
                keys[i, :len(p)] = np.asarray(p, np.float64)
        return keys, nk, maxk

    def get_word_prototypes(self, words: Sequence[str], num_points: int = 128) -> torch.Tensor:
        """(n, num_points, 3) fp32 device tensor; row i = the reference's get_word_prototype(words[i], num_points)."""
        keys, nk, maxk = self._pack(words, 64)
        n = len(words)
        out = torch.empty(n, num_points, 3, dtype=torch.float32, device=self.device)
        if n == 0:
            return out
        kd = torch.from_numpy(keys).to(self.device)
        nd = torch.from_numpy(nk).to(self.device)
        c = _lib.ctx(self.device)
        _lib.check(_lib.lib().wgg_word_prototypes(c, kd.data_ptr(), nd.data_ptr(), n, maxk, num_points, out.data_ptr(),
                                                  _lib.stream(self.device)), c)
        return out

    def get_word_prototype(self, word: str, num_points: int = 128) -> np.ndarray:
        return self.get_word_prototypes([word], num_points)[0].cpu().numpy()

    def get_minimum_jerk_trajectories(self, words: Sequence[str], num_points: int = 128, include_midpoints: bool = True,
                                      offset_std: float = 0.0) -> torch.Tensor:
        """(n, num_points, 3) fp32 device tensor; row i = the reference's get_minimum_jerk_trajectory(words[i], ...).
        With offset_std > 0 the offsets are drawn from numpy's global generator word by word exactly as the reference
        draws them: (k - 2, 2) key offsets, then k - 1 midpoint offsets (only when midpoints are included and k > 2)."""
        keys, nk, maxk = self._pack(words, 32)
        n = len(words)
        out = torch.empty(n, num_points, 3, dtype=torch.float32, device=self.device)
        if n == 0:
            return out
        kn = mn = None
        if offset_std > 0:
            kn = np.zeros((n, maxk, 2), np.float64)
            mn = np.zeros((n, maxk), np.float64)
            for i, k in enumerate(nk):
                if k > 2:
                    kn[i, :k - 2] = np.random.normal(0, offset_std, (k - 2, 2))
                    if include_midpoints:
                        for j in range(k - 1):
                            mn[i, j] = np.random.normal(0, offset_std * 0.5)
        kd = torch.from_numpy(keys).to(self.device)
        nd = torch.from_numpy(nk).to(self.device)
        knd = torch.from_numpy(kn).to(self.device) if kn is not None else None
        mnd = torch.from_numpy(mn).to(self.device) if mn is not None else None
        c = _lib.ctx(self.device)
        _lib.check(_lib.lib().wgg_minimum_jerk(c, kd.data_ptr(), nd.data_ptr(), n, maxk,
                                               knd.data_ptr() if knd is not None else None,
                                               mnd.data_ptr() if mnd is not None else None, int(include_midpoints),
                                               num_points, out.data_ptr(), _lib.stream(self.device)), c)
        return out

    def get_minimum_jerk_trajectory(self, word: str, num_points: int = 128, include_midpoints: bool = True,
                                    offset_std: float = 0.0) -> np.ndarray:
        return self.get_minimum_jerk_trajectories([word], num_points, include_midpoints, offset_std)[0].cpu().numpy()
