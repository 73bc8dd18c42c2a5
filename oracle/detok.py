"""
numpy restatement of the action de-tokeniser and un-normaliser (TEST INFRASTRUCTURE — see oracle/__init__.py).

  decode_token_ids_to_actions  /root/reference/prismatic/vla/action_tokenizer.py:49-68
                               == /root/reference/prismatic/extern/hf/modeling_prismatic.py:522-525
  unnormalize_actions          modeling_prismatic.py:528-535 == prismatic/models/vlms/prismatic.py:674-685
Pinned by tests/golden/detok_golden.json, which `oracle/gen_golden.py` produced by running the reference's own
`ActionTokenizer` / `Solver` code from /root/reference.
"""

from __future__ import annotations

import numpy as np


def bin_centers(n_bins: int = 256) -> np.ndarray:
    edges = np.linspace(-1, 1, n_bins)
    return (edges[:-1] + edges[1:]) / 2.0


def decode_token_ids_to_actions(token_ids: np.ndarray, vocab_size: int = 32000, n_bins: int = 256) -> np.ndarray:
    centers = bin_centers(n_bins)
    d = vocab_size - np.asarray(token_ids)
    d = np.clip(d - 1, a_min=0, a_max=centers.shape[0] - 1)
    return centers[d]


def unnormalize_actions(normalized: np.ndarray, stats: dict) -> np.ndarray:
    mask = stats.get("mask", np.ones_like(stats["q01"], dtype=bool))
    high, low = np.array(stats["q99"]), np.array(stats["q01"])
    return np.where(mask, 0.5 * (normalized + 1) * (high - low) + low, normalized)
