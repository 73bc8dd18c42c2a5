"""
Tokenizer plumbing.

The reference uses the gated `meta-llama/Llama-2-7b-hf` SentencePiece tokenizer
(/root/reference/prismatic/models/backbones/llm/llama2.py:55-102; module-level load at vla/solver.py:188). It is not
available offline, so `SyntheticLlamaTokenizer` is a deterministic stand-in with the properties the hot path relies on:

  * `vocab_size == 32000`, BOS=1, EOS=2, UNK=0, PAD=32000 (added token, llama2.py:74-76)
  * id 29871 is the SentencePiece "▁" dummy prefix that is prepended to every encoded text
    (the reason vla/solver.py:125-126 drops the first decoded value, and modeling_prismatic.py:513-516 appends it)
  * the last 256 ids (31744..31999) are single code points, so action tokens survive decode -> re-encode
    (vla/action_tokenizer.py:38-47, solver.py:121-124)

When a real tokenizer directory is supplied, `load_tokenizer` defers to `transformers.AutoTokenizer`.
"""

from __future__ import annotations

import os
from typing import Any, Dict, List, Optional, Sequence, Union

import torch

_KEYWORDS = ["POLICIES:", "MOVEMENT:", "NEXT GRIPPER:", "REASONING:", "SUBTASK:", "INSTRUCTION:", "CURRENT GRIPPER:"]


class _Encoding(dict):
    """Minimal `BatchEncoding`: attribute + key access, `.to(device)`."""

    def __getattr__(self, k: str) -> Any:
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def to(self, device: Any) -> "_Encoding":
        return _Encoding({k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in self.items()})


class SyntheticLlamaTokenizer:
    vocab_size = 32000
    unk_token_id, bos_token_id, eos_token_id, pad_token_id = 0, 1, 2, 32000
    prefix_id = 29871
    byte_lo = 3  # ids 3..258 = bytes 0x00..0xFF (as in Llama's byte fallback)
    keyword_lo = 300
    alias_lo = 400  # ids that decode to "\n" / ";" but are never produced by encode (lets scripts stay unique-id)
    filler_lo, filler_hi = 1000, 29000
    action_id_lo, action_id_hi = 31744, 31999
    model_input_names = ["input_ids", "attention_mask"]
    padding_side = "right"
    model_max_length = 2048
    name_or_path = "synthetic-llama-2"

    def __init__(self) -> None:
        self._id2piece: Dict[int, str] = {}
        self._piece2id: Dict[str, int] = {}
        for i, kw in enumerate(_KEYWORDS):
            self._id2piece[self.keyword_lo + i] = kw
            self._piece2id[kw] = self.keyword_lo + i
        for i in range(16):
            self._id2piece[self.alias_lo + i] = "\n"
            self._id2piece[self.alias_lo + 16 + i] = ";"
        for i in range(self.filler_lo, self.filler_hi):
            ch = chr(0x20000 + i)
            self._id2piece[i] = ch
            self._piece2id[ch] = i
        for i in range(self.action_id_lo, self.action_id_hi + 1):
            ch = chr(0x4E00 + i - self.action_id_lo)
            self._id2piece[i] = ch
            self._piece2id[ch] = i

    # --- helpers used by the scripted synthetic checkpoints -------------------------------------------------------
    def key_id(self, kw: str) -> int:
        return self._piece2id[kw]

    def newline_id(self, k: int) -> int:
        return self.alias_lo + k

    def semicolon_id(self, k: int) -> int:
        return self.alias_lo + 16 + k

    def __len__(self) -> int:
        return self.vocab_size + 1  # + <PAD>

    # --- encode ---------------------------------------------------------------------------------------------------
    def _encode_one(self, text: str, add_special_tokens: bool) -> List[int]:
        ids: List[int] = [self.bos_token_id] if add_special_tokens else []
        if len(text) == 0:
            return ids
        ids.append(self.prefix_id)
        i, n = 0, len(text)
        while i < n:
            for kw in _KEYWORDS:
                if text.startswith(kw, i):
                    ids.append(self._piece2id[kw])
                    i += len(kw)
                    break
            else:
                ch = text[i]
                if ch in self._piece2id:
                    ids.append(self._piece2id[ch])
                else:
                    ids.extend(self.byte_lo + b for b in ch.encode("utf-8"))
                i += 1
        return ids

    def __call__(
        self,
        text: Union[str, Sequence[str]],
        add_special_tokens: bool = True,
        return_tensors: Optional[str] = None,
        padding: Any = False,
        truncation: Any = None,
        max_length: Optional[int] = None,
        **_: Any,
    ) -> _Encoding:
        batched = not isinstance(text, str)
        texts = list(text) if batched else [text]
        seqs = [self._encode_one(t, add_special_tokens) for t in texts]
        if truncation:
            lim = max_length or self.model_max_length
            seqs = [s[:lim] for s in seqs]
        masks = [[1] * len(s) for s in seqs]
        if return_tensors in ("pt", "PYTORCH") or str(return_tensors).endswith("PYTORCH"):
            L = max(len(s) for s in seqs)
            if any(len(s) != L for s in seqs):
                if not padding:
                    raise ValueError("Unable to create tensor: sequences differ in length and `padding` is off")
                masks = [m + [0] * (L - len(m)) for m in masks]
                seqs = [s + [self.pad_token_id] * (L - len(s)) for s in seqs]
            return _Encoding(input_ids=torch.tensor(seqs, dtype=torch.long), attention_mask=torch.tensor(masks, dtype=torch.long))
        if not batched:
            return _Encoding(input_ids=seqs[0], attention_mask=masks[0])
        return _Encoding(input_ids=seqs, attention_mask=masks)

    def encode(self, text: str, add_special_tokens: bool = True) -> List[int]:
        return self._encode_one(text, add_special_tokens)

    # --- decode ---------------------------------------------------------------------------------------------------
    def decode(self, token_ids: Any = None, skip_special_tokens: bool = False, **kw: Any) -> str:
        if token_ids is None:
            token_ids = kw.get("sequences")
        if isinstance(token_ids, torch.Tensor):
            token_ids = token_ids.tolist()
        elif hasattr(token_ids, "tolist"):
            token_ids = token_ids.tolist()
        if isinstance(token_ids, int):
            token_ids = [token_ids]
        out: List[str] = []
        pending = bytearray()

        def flush() -> None:
            if pending:
                out.append(pending.decode("utf-8", errors="replace"))
                pending.clear()

        for t in token_ids:
            t = int(t)
            if self.byte_lo <= t < self.byte_lo + 256:
                pending.append(t - self.byte_lo)
                continue
            flush()
            if t in (self.unk_token_id, self.bos_token_id, self.eos_token_id, self.pad_token_id):
                if not skip_special_tokens:
                    out.append({0: "<unk>", 1: "<s>", 2: "</s>", 32000: "<PAD>"}[t])
            elif t == self.prefix_id:
                out.append(" ")
            elif t in self._id2piece:
                out.append(self._id2piece[t])
            else:
                out.append("�")
        flush()
        s = "".join(out)
        return s[1:] if s.startswith(" ") else s  # SentencePiece strips the dummy-prefix space

    def batch_decode(self, sequences: Any = None, skip_special_tokens: bool = False, **kw: Any) -> List[str]:
        if isinstance(sequences, torch.Tensor):
            sequences = sequences.tolist()
        return [self.decode(s, skip_special_tokens=skip_special_tokens) for s in sequences]


def load_tokenizer(path: Optional[str]) -> Any:
    """Real tokenizer if the directory has one. `path=None` (synthetic models) gives the synthetic stand-in; a checkpoint directory
    WITHOUT tokenizer files also does, but loudly: its vocabulary is not the checkpoint's, so prompts and decoded text are only
    meaningful for the seeded synthetic weights."""
    if path is not None and os.path.isdir(path):
        if any(os.path.exists(os.path.join(path, f)) for f in ("tokenizer.model", "tokenizer.json")):
            from transformers import AutoTokenizer

            return AutoTokenizer.from_pretrained(path, model_max_length=2048, padding_side="right")
        import warnings

        warnings.warn(
            f"{path!r} holds no tokenizer.model / tokenizer.json: falling back to the SYNTHETIC Llama-shaped tokenizer. Prompts, decoded "
            "reasoning text and the Solver round trip will NOT match a real checkpoint's vocabulary - copy the checkpoint's tokenizer files "
            "into the directory (or pass tokenizer=...) for real weights.",
            UserWarning,
            stacklevel=2,
        )
    return SyntheticLlamaTokenizer()
