"""
Host-side 256-bin action (de)tokeniser.

Same arithmetic as /root/reference/prismatic/vla/action_tokenizer.py:28-68 (and its HF twin,
extern/hf/modeling_prismatic.py:522-525): bin edges `linspace(-1, 1, 256)`, 255 mid-points, and a token id `t` decodes
to centre index `clip(vocab_size - t - 1, 0, 254)`. The device twin is `emx_detokenize_actions`
(include/emmax.h) which the model uses on the generated ids without leaving the GPU; this class is what the text
round trip in `Solver` needs (decode -> re-tokenise), and it is what the tests check the kernel against.
"""

from __future__ import annotations

from typing import Any, List, Union

import numpy as np


class ActionTokenizer:
    def __init__(self, tokenizer: Any, bins: int = 256, min_action: int = -1, max_action: int = 1) -> None:
        self.tokenizer, self.n_bins, self.min_action, self.max_action = tokenizer, bins, min_action, max_action
        self.bins = np.linspace(min_action, max_action, bins)
        self.bin_centers = 0.5 * (self.bins[:-1] + self.bins[1:])
        # the reference reserves the last `bins` ids; "+ 1" because digitize() is 1-based (action_tokenizer.py:36)
        self.action_token_begin_idx = int(tokenizer.vocab_size - (bins + 1))

    def __call__(self, action: np.ndarray) -> Union[str, List[str]]:
        clipped = np.clip(action, a_min=float(self.min_action), a_max=float(self.max_action))
        which_bin = np.digitize(clipped, self.bins)  # in [1, bins]
        ids = self.tokenizer.vocab_size - which_bin
        if which_bin.ndim == 1:
            return self.tokenizer.decode(list(ids))
        return self.tokenizer.batch_decode(ids.tolist())

    def decode_token_ids_to_actions(self, action_token_ids: np.ndarray) -> np.ndarray:
        idx = self.tokenizer.vocab_size - np.asarray(action_token_ids)
        idx = np.clip(idx - 1, a_min=0, a_max=self.bin_centers.shape[0] - 1)
        return self.bin_centers[idx]

    @property
    def vocab_size(self) -> int:
        return self.n_bins
