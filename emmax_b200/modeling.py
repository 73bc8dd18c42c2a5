"""
Drop-in model surface: `AutoModelForVision2Seq` / `OpenVLAForActionPrediction`.

Mirrors the public behaviour of
  * /root/reference/prismatic/extern/hf/modeling_prismatic.py:492-566  (`OpenVLAForActionPrediction`: `predict_action`,
    `_check_unnorm_key`, `get_action_dim`, `get_action_stats`, `bins`, `bin_centers`, `vocab_size`, `norm_stats`)
  * /root/reference/prismatic/extern/hf/modeling_prismatic.py:291-485  (`forward` inference subset,
    `prepare_inputs_for_generation` batch-size rule)
  * /root/reference/prismatic/models/vlms/prismatic.py:627-696          (`generate_actions(image, prompt_text, type, **kw)`)
  * README.md:35-47 of the reference (`from_pretrained(...)`, `generate_actions(inputs, tokenizer, ...)`)
so that experiments/robot/openvla_utils.py:43-72,169,215-217 and vla-scripts/extern/verify_openvla.py:30-85 run
unchanged against this class. All compute goes through emmax_b200.engine.Engine (libemmax.so); there is no eager path.
"""

from __future__ import annotations

import json
import os
from collections.abc import Mapping
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib
from .action_tokenizer import ActionTokenizer
from .configuration import OpenVLAConfig, emma_x_config
from .engine import Engine
from .processing import PrismaticImageProcessor
from .prompting import PurePromptBuilder
from .solver import Solver, unnormalize
from .tokenization import load_tokenizer


@dataclass
class PrismaticCausalLMOutputWithPast:
    loss: Optional[torch.Tensor] = None
    logits: Optional[torch.Tensor] = None
    past_key_values: Any = None
    hidden_states: Any = None
    attentions: Any = None
    projector_features: Optional[torch.Tensor] = None


class _Namespace:
    def __init__(self, **kw: Any) -> None:
        self.__dict__.update(kw)


def _load_checkpoint_tensors(path: str) -> Dict[str, torch.Tensor]:
    """safetensors shards (HF export, convert_openvla_weights_to_hf.py:244-250) or a torch `.pt` state dict."""
    from safetensors.torch import load_file

    index = os.path.join(path, "model.safetensors.index.json")
    sd: Dict[str, torch.Tensor] = {}
    if os.path.exists(index):
        with open(index) as f:
            shards = sorted(set(json.load(f)["weight_map"].values()))
        for s in shards:
            sd.update(load_file(os.path.join(path, s)))
    elif os.path.exists(os.path.join(path, "model.safetensors")):
        sd = load_file(os.path.join(path, "model.safetensors"))
    else:
        pts = [f for f in os.listdir(path) if f.endswith((".pt", ".bin"))]
        if not pts:
            raise FileNotFoundError(f"no weights under {path}")
        sd = torch.load(os.path.join(path, pts[0]), map_location="cpu")
        sd = sd.get("model", sd)
        if isinstance(sd.get("llm_backbone"), dict):  # a native Prismatic checkpoint (component dicts): apply the converter's name map
            from .load import remap_native_state_dict

            sd = remap_native_state_dict(sd)
    return sd


class OpenVLAForActionPrediction:
    config_class = OpenVLAConfig

    def __init__(self, config: OpenVLAConfig, state_dict: Dict[str, torch.Tensor], tokenizer: Any = None,
                 max_context: int = 2048, max_batch: int = 1) -> None:  # fmt: skip
        if config.use_fused_vision_backbone is None:
            raise ValueError("Missing config field `use_fused_vision_backbone`")
        self.config = config
        self.norm_stats = config.norm_stats
        self.proprio_norm_stats: Optional[dict] = None
        self.bins = np.linspace(-1, 1, config.n_action_bins)
        self.bin_centers = (self.bins[:-1] + self.bins[1:]) / 2.0
        # de-tokenisation vocabulary: the tokenizer's 32000, i.e. padded embedding rows minus the pad (modeling_prismatic.py:504)
        self.vocab_size = config.text_config.vocab_size - config.pad_to_multiple_of
        self.pad_token_id = config.pad_token_id
        self.training = False
        self._sd: Optional[Dict[str, torch.Tensor]] = state_dict
        self._engine: Optional[Engine] = None
        self._max_context, self._max_batch = max_context, max_batch
        self._tokenizer = tokenizer if tokenizer is not None else load_tokenizer(None)
        self.action_tokenizer = ActionTokenizer(self._tokenizer)
        self.solver = Solver(self.action_tokenizer, verbose=False)
        # attribute paths the native callers touch (prismatic.py:630: vision_backbone.image_transform, llm_backbone.tokenizer)
        self.vision_backbone = _Namespace(image_transform=PrismaticImageProcessor(image_resize_strategy=config.image_resize_strategy).apply_transform)
        self.llm_backbone = _Namespace(tokenizer=self._tokenizer, half_precision_dtype=torch.bfloat16)
        self._d_stats: Dict[str, Tuple[torch.Tensor, ...]] = {}

    # ---- construction ------------------------------------------------------------------------------------------
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, *_, torch_dtype: Any = torch.bfloat16,
                        attn_implementation: Optional[str] = None, low_cpu_mem_usage: bool = True,
                        trust_remote_code: bool = True, load_in_8bit: bool = False, load_in_4bit: bool = False,
                        **kwargs: Any) -> "OpenVLAForActionPrediction":  # fmt: skip
        if load_in_8bit or load_in_4bit:
            raise NotImplementedError("emmax_b200 runs bf16 weights; bitsandbytes quantisation is out of scope")
        if torch_dtype not in (torch.bfloat16, None, "bfloat16", "auto"):
            raise NotImplementedError(f"emmax_b200 computes in bf16 (reference HF path); got torch_dtype={torch_dtype}")
        path = pretrained_model_name_or_path
        if not os.path.isdir(path):
            raise FileNotFoundError(f"{path!r} is not a local directory (offline: no hub access); use from_synthetic() for a random-init model")
        config = OpenVLAConfig.from_pretrained(path)
        stats = os.path.join(path, "dataset_statistics.json")
        if os.path.isfile(stats):  # openvla_utils.py:59-70
            with open(stats) as f:
                config.norm_stats = json.load(f)
        return cls(config, _load_checkpoint_tensors(path), tokenizer=load_tokenizer(path), **kwargs)

    @classmethod
    def from_synthetic(cls, config: Optional[OpenVLAConfig] = None, seed: int = 0, device: Union[str, torch.device] = "cuda",
                       script: Optional[List[int]] = None, script_prev: Optional[int] = None, **kwargs: Any) -> "OpenVLAForActionPrediction":  # fmt: skip
        from .synthetic import make_state_dict

        config = config or emma_x_config()
        gen_device = device if config.text_config.hidden_size >= 1024 else "cpu"  # toy configs: CPU RNG, reproducible anywhere
        sd = make_state_dict(config, seed=seed, device=gen_device, script=script, script_prev=script_prev)
        return cls(config, sd, **kwargs).to(device)

    def to(self, *args: Any, **kwargs: Any) -> "OpenVLAForActionPrediction":
        device = kwargs.get("device")
        for a in args:
            if isinstance(a, (str, torch.device)):
                device = a
            elif isinstance(a, torch.dtype):
                self._check_dtype(a)
        if isinstance(kwargs.get("dtype"), torch.dtype):
            self._check_dtype(kwargs["dtype"])
        if device is None:
            return self
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.EmxError("emmax_b200 has no CPU path (libemmax.so is sm_100a only)")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if self._engine is None:
            self._engine = Engine(self.config, self._sd, device, max_batch=self._max_batch, max_context=self._max_context)
            self._sd = None
        elif self._engine.device != device:
            raise NotImplementedError("moving a materialised engine between devices is not supported; build a replica instead")
        return self

    @staticmethod
    def _check_dtype(dtype: torch.dtype) -> None:
        if dtype == torch.float16:
            # experiments/robot/robot_utils.py:42 moves the natively loaded model with `.to(device, dtype=torch.float16)`
            import warnings

            warnings.warn("emmax_b200 computes in bf16 (the reference HF path's dtype, openvla_utils.py:46); the float16 request is ignored", stacklevel=3)
        elif dtype != torch.bfloat16:
            raise NotImplementedError("emmax_b200 computes in bf16")

    def cuda(self, device: Optional[int] = None) -> "OpenVLAForActionPrediction":
        return self.to(torch.device("cuda", device if device is not None else torch.cuda.current_device()))

    def eval(self) -> "OpenVLAForActionPrediction":
        return self

    @property
    def device(self) -> torch.device:
        return self._engine.device if self._engine is not None else torch.device("cpu")

    @property
    def engine(self) -> Engine:
        if self._engine is None:
            raise _lib.EmxError("model is not on a CUDA device yet: call .to('cuda:0') (there is no CPU path)")
        return self._engine

    # ---- generation ----------------------------------------------------------------------------------------------
    def _check_generate_inputs(self, input_ids: Optional[torch.Tensor], inputs_embeds: Optional[torch.Tensor] = None) -> None:
        if ((input_ids is not None) and (input_ids.shape[0] > 1)) or ((inputs_embeds is not None) and (inputs_embeds.shape[0] > 1)):
            raise ValueError("Generation with batch size > 1 is not currently supported!")

    @torch.no_grad()
    def generate(self, input_ids: Optional[torch.Tensor] = None, pixel_values: Optional[torch.Tensor] = None,
                 attention_mask: Optional[torch.Tensor] = None, max_new_tokens: Optional[int] = None,
                 max_length: Optional[int] = None, min_length: int = 0, do_sample: bool = False, temperature: float = 0.0,
                 eos_token_id: Union[int, None, str] = "default", **kwargs: Any) -> torch.Tensor:  # fmt: skip
        """Greedy `GenerationMixin.generate`: returns [1, P + T] ids (prompt ids without the patch positions + new ids)."""
        self._check_generate_inputs(input_ids, kwargs.get("inputs_embeds"))
        if do_sample:
            raise NotImplementedError("only greedy decoding (do_sample=False) is implemented, as used by the reference callers")
        if pixel_values is None:
            raise ValueError("language-only generation is not part of the accelerated path: pass `pixel_values`")
        if isinstance(pixel_values, dict):  # native path passes {"dino": [1,3,h,w], "siglip": [1,3,h,w]} (prismatic.py:646-652)
            pixel_values = torch.cat([pixel_values["dino"], pixel_values["siglip"]], dim=1)
        P = input_ids.shape[1]
        if max_new_tokens is None:
            max_new_tokens = (max_length - P) if max_length is not None else 20
        eos = self.config.text_config.eos_token_id if eos_token_id == "default" else eos_token_id
        if max_new_tokens <= 0:
            return input_ids
        new, _ = self.engine.generate(input_ids.to(self.device), pixel_values.to(self.device, torch.bfloat16), max_new_tokens, eos_token_id=eos)
        return torch.cat([input_ids.to(self.device), new.to(torch.long)[None]], dim=1)

    def _device_stats(self, key: str) -> Tuple[torch.Tensor, ...]:
        if key not in self._d_stats:
            st = self.norm_stats[key]["action"]
            mask = st.get("mask", np.ones_like(st["q01"], dtype=bool))
            self._d_stats[key] = (
                torch.tensor(np.array(st["q01"], dtype=np.float64), device=self.device),
                torch.tensor(np.array(st["q99"], dtype=np.float64), device=self.device),
                torch.tensor(np.array(mask, dtype=np.uint8), device=self.device),
            )
        return self._d_stats[key]

    def detokenize_on_device(self, token_ids: torch.Tensor, unnorm_key: Optional[str] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """ids [n] (int32, device) -> (normalized [n], actions [n]) fp64 on device, via emx_detokenize_actions."""
        key = self._check_unnorm_key(self.norm_stats, unnorm_key)
        q01, q99, mask = self._device_stats(key)
        ids = token_ids.to(device=self.device, dtype=torch.int32).contiguous()
        n = ids.numel()
        norm = torch.empty(n, dtype=torch.float64, device=self.device)
        act = torch.empty(n, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):  # launch on a stream of the engine's GPU, whatever the caller's current device is
            _lib.call("emx_detokenize_actions", ids.data_ptr(), n, self.vocab_size, self.config.n_action_bins, q01.data_ptr(), q99.data_ptr(),
                      mask.data_ptr(), q01.numel(), norm.data_ptr(), act.data_ptr(), _lib.stream())  # fmt: skip
        return norm, act

    @torch.no_grad()
    def predict_action(self, input_ids: Optional[torch.Tensor] = None, unnorm_key: Optional[str] = None, **kwargs: Any) -> np.ndarray:
        """modeling_prismatic.py:506-537: append 29871 if absent, generate action_dim tokens, de-tokenise, un-normalise."""
        if not torch.all(input_ids[:, -1] == 29871):
            input_ids = torch.cat((input_ids, torch.tensor([[29871]], dtype=input_ids.dtype, device=input_ids.device)), dim=1)
        n = self.get_action_dim(unnorm_key)
        kwargs.pop("max_new_tokens", None)
        # EOS stays active exactly as in the reference's `self.generate(...)` call (modeling_prismatic.py:519): a model that emits </s> early
        # stops there, and `[-n:]` below then reaches back into the prompt ids just as the reference's slice does
        generated = self.generate(input_ids, max_new_tokens=n, **kwargs)
        _, actions = self.detokenize_on_device(generated[0, -n:], unnorm_key)
        return actions.cpu().numpy()

    @torch.no_grad()
    def generate_actions(self, *args: Any, **kwargs: Any) -> Tuple[Any, str]:
        """Two call forms:
        README.md:44-47  `generate_actions(inputs, tokenizer, do_sample=False, max_new_tokens=512)` -> (action[7], reasoning)
        prismatic.py:627 `generate_actions(image, prompt_text, type, **gen_kwargs)` -> (list of action[7] | proprio[7], text)
        """
        first = args[0] if args else kwargs.get("inputs", kwargs.get("image"))
        if isinstance(first, Mapping):
            inputs = first
            tokenizer = args[1] if len(args) > 1 else kwargs.pop("tokenizer", self._tokenizer)
            kwargs.pop("inputs", None)
            actions, text = self._generate_actions_ids(inputs["input_ids"], inputs["pixel_values"], tokenizer, "act", **kwargs)
            return actions[0], text
        image = first
        prompt_text = args[1] if len(args) > 1 else kwargs.pop("prompt_text")
        kind = args[2] if len(args) > 2 else kwargs.pop("type")
        kwargs.pop("image", None)
        tok = self.llm_backbone.tokenizer
        input_ids = tok(prompt_text, truncation=True, return_tensors="pt").input_ids
        pixel_values = self.vision_backbone.image_transform(image)[None, ...]
        return self._generate_actions_ids(input_ids, pixel_values, tok, kind, **kwargs)

    # ---- batched requests: an extension over the reference, whose cached generation asserts batch size 1 --------------
    @torch.no_grad()
    def generate_batch(self, input_ids: Sequence[torch.Tensor], pixel_values: Sequence[torch.Tensor], max_new_tokens: Union[int, Sequence[int]] = 512,
                       eos_token_id: Union[int, None, str] = "default", continuous: bool = False, order: str = "fifo") -> List[torch.Tensor]:  # fmt: skip
        """N independent requests (prompt ids [1, n_i], pixel_values [1, 6, h, w] each; N robots or N simulator environments) decoded
        8 at a time with ONE pass over the weights per token for the whole group (Engine.generate_batch / emx_decode_batch_step) - where
        the reference raises "Generation with batch size > 1 is not currently supported!" (modeling_prismatic.py:460-463) and a caller
        has to loop. Every request's result is what its own bs=1 `generate` returns: [1, n_i + T_i] ids.
        continuous=True: the N requests are a stream through the 8 sequence slots (Engine.serve): a slot is refilled with the next request as
        soon as its sequence ends instead of waiting for the longest sequence of its group of 8; `order` ("fifo" | "longest_first") is the
        admission order of that stream (results always come back in request order)."""
        N = len(input_ids)
        limits = [int(max_new_tokens)] * N if isinstance(max_new_tokens, int) else [int(x) for x in max_new_tokens]
        eos = self.config.text_config.eos_token_id if eos_token_id == "default" else eos_token_id
        eng, dev = self.engine, self.device
        if continuous:
            new = eng.serve([(input_ids[i], pixel_values[i], limits[i]) for i in range(N)], eos_token_id=eos, order=order)
            return [torch.cat([input_ids[i].to(dev), new[i].to(torch.long)[None]], dim=1) for i in range(N)]
        group = min(_lib.MAX_DECODE_BATCH, eng.max_batch)
        out: List[torch.Tensor] = []
        for g0 in range(0, N, group):
            ids = [x.to(dev) for x in input_ids[g0 : g0 + group]]
            pv = torch.cat([x.to(dev, torch.bfloat16) for x in pixel_values[g0 : g0 + group]], dim=0)
            same = len({x.shape[1] for x in ids}) == 1
            new, _ = eng.generate_batch(torch.cat(ids, dim=0) if same else ids, pv, limits[g0 : g0 + group], eos_token_id=eos)
            out += [torch.cat([ids[b], new[b].to(torch.long)[None]], dim=1) for b in range(len(ids))]
        return out

    @torch.no_grad()
    def predict_action_batch(self, inputs: Sequence[Mapping], unnorm_key: Optional[str] = None) -> np.ndarray:
        """`predict_action` (modeling_prismatic.py:506-537) for N processor outputs at once: [N, action_dim] float64, row i identical to
        `predict_action(**inputs[i], unnorm_key=unnorm_key)` — N simulator environments / robots share each pass over the weights."""
        n = self.get_action_dim(unnorm_key)
        ids = []
        for i in inputs:
            x = i["input_ids"]
            if not torch.all(x[:, -1] == 29871):
                x = torch.cat((x, torch.tensor([[29871]], dtype=x.dtype, device=x.device)), dim=1)
            ids.append(x)
        gen = self.generate_batch(ids, [i["pixel_values"] for i in inputs], max_new_tokens=n)
        tail = torch.stack([g[0, -n:] for g in gen]).reshape(-1)
        _, actions = self.detokenize_on_device(tail, unnorm_key)
        return actions.cpu().numpy().reshape(len(gen), n)

    @torch.no_grad()
    def generate_actions_batch(self, inputs: Sequence[Mapping], tokenizer: Any = None, type: str = "act",  # noqa: A002
                               max_new_tokens: Union[int, Sequence[int]] = 512, do_sample: bool = False,
                               continuous: bool = False, order: str = "fifo") -> List[Tuple[Any, str]]:
        """`generate_actions(inputs, tokenizer, ...)` (README.md:44-47) for a list of processor outputs at once: one (action[7], reasoning)
        per request, each identical to its own bs=1 call. continuous=True / order: see generate_batch."""
        if do_sample:
            raise NotImplementedError("only greedy decoding (do_sample=False) is implemented, as used by the reference callers")
        tokenizer = tokenizer if tokenizer is not None else self._tokenizer
        gen = self.generate_batch([i["input_ids"] for i in inputs], [i["pixel_values"] for i in inputs], max_new_tokens, continuous=continuous, order=order)
        res = []
        for i, g in zip(inputs, gen):
            acts, text = self._parse_generated(g, i["input_ids"].shape[1], tokenizer, type)
            res.append((acts[0] if type == "act" else acts, text))
        return res

    def _generate_actions_ids(self, input_ids: torch.Tensor, pixel_values: torch.Tensor, tokenizer: Any, kind: str, **gen_kwargs: Any):
        generated = self.generate(input_ids=input_ids, pixel_values=pixel_values, **gen_kwargs)
        return self._parse_generated(generated, input_ids.shape[1], tokenizer, kind)

    def _parse_generated(self, generated: torch.Tensor, n_prompt: int, tokenizer: Any, kind: str):
        input_ids = generated[:, :n_prompt]
        text = tokenizer.decode(generated[0, input_ids.shape[1] :], skip_special_tokens=True).strip()
        solver = self.solver if tokenizer is self._tokenizer else Solver(ActionTokenizer(tokenizer), verbose=False)
        if kind == "act":
            policies, _reasoning = solver.extract_action_policies(text)
            stats = self.get_action_stats(None)
            return [unnormalize(a, stats) for a in policies], text
        if kind == "pos":
            require_unorm, delta = solver.extract_movement_plan(text)
            proprio = delta
            if require_unorm:
                proprio = unnormalize(delta, self.get_proprio_stats(), low_key="Q1", high_key="Q99")
            return proprio, text
        raise ValueError(f"unknown generate_actions type {kind!r} (expected 'act' or 'pos')")

    # ---- forward (inference subset of modeling_prismatic.py:291-447) -----------------------------------------------
    @torch.no_grad()
    def forward(self, input_ids: Optional[torch.Tensor] = None, attention_mask: Optional[torch.Tensor] = None,
                pixel_values: Optional[torch.Tensor] = None, labels: Optional[torch.Tensor] = None,
                inputs_embeds: Optional[torch.Tensor] = None, past_key_values: Any = None, use_cache: Optional[bool] = None,
                output_attentions: Optional[bool] = None, output_hidden_states: Optional[bool] = None,
                output_projector_features: Optional[bool] = None, return_dict: Optional[bool] = None) -> PrismaticCausalLMOutputWithPast:  # fmt: skip
        if labels is not None or inputs_embeds is not None or output_attentions or output_hidden_states:
            raise NotImplementedError("training / introspection outputs are outside the accelerated inference path")
        eng = self.engine
        with torch.cuda.device(eng.device):
            return self._forward_on_device(eng, input_ids, pixel_values, past_key_values, output_projector_features)

    def _forward_on_device(self, eng: Engine, input_ids, pixel_values, past_key_values, output_projector_features):
        if input_ids.shape[1] == 1 and past_key_values is not None:
            assert input_ids.shape[0] == 1, "Generation is only currently supported for batch size of 1!"
            return self._forward_cached(input_ids)
        if pixel_values is None:
            raise NotImplementedError("language-only forward is outside the accelerated path")
        if input_ids.shape[0] != pixel_values.shape[0]:
            raise ValueError("Non-homogenous batch of (text, image) input -- forward() does not support mixed batches!")
        ws = eng.prefill(input_ids.to(self.device), pixel_values.to(self.device, torch.bfloat16), use_graph=False)
        B, S, H, V = input_ids.shape[0], ws["S"], eng.t.hidden_size, eng.t.vocab_size
        # full-sequence logits like the reference (the generate path only ever computes the last row)
        normed = torch.empty_like(ws["x"])
        _lib.call("emx_rmsnorm", ws["x"].data_ptr(), eng.final_norm.data_ptr(), normed.data_ptr(), B * S, H, eng.t.rms_norm_eps, _lib.stream())
        logits = torch.empty((B * S, V), dtype=torch.bfloat16, device=self.device)
        eng.gemm(normed, eng.lm_head, logits)
        st = eng.d_state
        st[0:1].copy_(ws["first"][0:1]), st[1].fill_(S), st[2].fill_(0), st[3].zero_()
        self._cached_pos = S  # host mirror of the KV length, for the capacity check of the caller-driven cached loop
        return PrismaticCausalLMOutputWithPast(
            logits=logits.view(B, S, V), past_key_values=eng,
            projector_features=ws["patches"].view(B, -1, H).clone() if output_projector_features else None,
        )  # fmt: skip

    def _forward_cached(self, input_ids: torch.Tensor) -> PrismaticCausalLMOutputWithPast:
        import ctypes as C

        eng = self.engine
        V = eng.t.vocab_size
        pos = getattr(self, "_cached_pos", None)
        if pos is None:
            raise ValueError("cached forward() without a preceding multimodal forward(): there is no KV cache to extend")
        if pos >= eng.max_context:
            raise ValueError(f"KV cache is full: {pos} positions cached, engine capacity max_context={eng.max_context}")
        self._cached_pos = pos + 1
        logits = torch.empty((1, 1, V), dtype=torch.float32, device=self.device)
        eng.d_state[0:1].copy_(input_ids.reshape(-1)[:1].to(device=self.device, dtype=torch.int32))
        p = eng._decode_params(0)
        p.logits_out = logits.data_ptr()
        _lib.check(_lib.load().emx_decode_step(C.byref(p), _lib.stream()))
        return PrismaticCausalLMOutputWithPast(logits=logits.to(torch.bfloat16), past_key_values=eng)

    __call__ = forward

    # ---- statistics helpers (modeling_prismatic.py:539-566; prismatic/models/vlas/openvla.py:106-137) ---------------
    @staticmethod
    def _check_unnorm_key(norm_stats: Dict[str, Dict[str, Any]], unnorm_key: Optional[str]) -> str:
        if unnorm_key is None and len(norm_stats) != 1:
            raise ValueError(
                f"Your model was trained on more than one dataset. "
                f"Please pass a `unnorm_key` from the following options to choose the statistics used for "
                f"de-normalizing actions: {norm_stats.keys()}"
            )
        unnorm_key = unnorm_key if unnorm_key is not None else next(iter(norm_stats.keys()))
        if unnorm_key not in norm_stats:
            raise ValueError(
                f"The `unnorm_key` you chose ({unnorm_key = }) is not in the available statistics. "
                f"Please choose from: {norm_stats.keys()}"
            )
        return unnorm_key

    def get_action_dim(self, unnorm_key: Optional[str] = None) -> int:
        return len(self.norm_stats[self._check_unnorm_key(self.norm_stats, unnorm_key)]["action"]["q01"])

    def get_action_stats(self, unnorm_key: Optional[str] = None) -> Dict[str, Any]:
        return self.norm_stats[self._check_unnorm_key(self.norm_stats, unnorm_key)]["action"]

    def get_proprio_stats(self) -> Optional[dict]:
        return self.proprio_norm_stats

    def get_prompt_builder(self, system_prompt: Optional[str] = None) -> PurePromptBuilder:
        return PurePromptBuilder("prismatic", system_prompt=system_prompt)


class AutoModelForVision2Seq:
    """`AutoModelForVision2Seq.from_pretrained(path, attn_implementation=..., torch_dtype=torch.bfloat16, ...)`
    (README.md:35-41; experiments/robot/openvla_utils.py:43-51). transformers 5.x no longer ships this Auto class."""

    @staticmethod
    def from_pretrained(path: str, *args: Any, **kwargs: Any) -> OpenVLAForActionPrediction:
        return OpenVLAForActionPrediction.from_pretrained(path, *args, **kwargs)

    @staticmethod
    def register(*_: Any, **__: Any) -> None:
        return None


class AutoConfig:
    @staticmethod
    def from_pretrained(path: str, **kwargs: Any) -> OpenVLAConfig:
        return OpenVLAConfig.from_pretrained(path, **kwargs)

    @staticmethod
    def register(*_: Any, **__: Any) -> None:
        return None
