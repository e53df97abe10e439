"""mapf_gpt_b200 -- B200-native rollout engine behind the MAPF-GPT inference surface.

Only the hot path: POGEMA grid step -> FOV tokenizer -> GPT forward -> action
(reference: mapf_gpt/inference.py, observation_generator.{h,cpp}, model.py).
"""
from .weights import GPTConfig, model_config, random_init, load_checkpoint, save_checkpoint  # noqa: F401

__all__ = ["GPTConfig", "model_config", "random_init", "load_checkpoint", "save_checkpoint"]
