"""
emmax_b200 — B200-native (sm_100a) implementation of Emma-X's `generate_actions` hot path behind the reference's
Python surface (`AutoModelForVision2Seq` / `AutoProcessor` / `generate_actions` / `predict_action`).

    from emmax_b200 import AutoModelForVision2Seq, AutoProcessor
    vla = AutoModelForVision2Seq.from_pretrained(path, torch_dtype=torch.bfloat16).to("cuda:0")
    processor = AutoProcessor.from_pretrained(path)
    prompt, image = processor.get_prompt(task_label, image)
    inputs = processor(prompt, image).to("cuda:0", dtype=torch.bfloat16)
    action, reasoning = vla.generate_actions(inputs, processor.tokenizer, do_sample=False, max_new_tokens=512)

All compute runs in hand-written CUDA kernels (libemmax.so, C ABI in include/emmax.h). No CPU / eager fallback exists.
"""

from .action_tokenizer import ActionTokenizer
from .configuration import OpenVLAConfig, PrismaticConfig, emma_x_config, tiny_config
from .load import load_vla
from .modeling import AutoConfig, AutoModelForVision2Seq, OpenVLAForActionPrediction, PrismaticCausalLMOutputWithPast
from .processing import AutoImageProcessor, AutoProcessor, BatchFeature, PrismaticImageProcessor, PrismaticProcessor
from .prompting import PurePromptBuilder, emma_x_prompt, openvla_prompt
from .simpler_policy import BatchedOpenVLAInference, OpenVLAInference
from .solver import Solver
from .tokenization import SyntheticLlamaTokenizer

__all__ = [
    "ActionTokenizer", "AutoConfig", "BatchedOpenVLAInference", "AutoImageProcessor", "AutoModelForVision2Seq", "AutoProcessor", "BatchFeature",
    "OpenVLAConfig", "OpenVLAForActionPrediction", "OpenVLAInference", "PrismaticCausalLMOutputWithPast", "PrismaticConfig",
    "PrismaticImageProcessor", "PrismaticProcessor", "PurePromptBuilder", "Solver", "SyntheticLlamaTokenizer",
    "emma_x_config", "emma_x_prompt", "load_vla", "openvla_prompt", "tiny_config",
]  # fmt: skip
