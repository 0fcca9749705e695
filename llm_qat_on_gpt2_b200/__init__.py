"""B200-native switchable-precision fake-quant linear path for GPT-2 QAT.

Drop-in replacements for the hot-path classes of Laurence-Wu/LLM-QAT-on-gpt2
(part1_switchable_precision.{quantization, quantization_methods, lora, switchable_batchnorm,
models_sp}), backed by hand-written sm_100a CUDA kernels behind a C ABI (include/spq_b200.h,
libspq_b200.so).  There is no CPU path: the modules raise on non-CUDA tensors or when the library
has not been built (`python -m llm_qat_on_gpt2_b200.build`).
"""
from .quantization import LearnableFakeQuantize, calibrate_many
from .quantization_methods import (apply_log_quantization, apply_minmax_quantization, quantize_codes,
                                   LogQuantizationFunction, MinMaxQuantizationFunction)
from .lora import LoRALayer, SPLinearWithLoRA, linear_fp
from .switchable_batchnorm import SwitchableLayerNorm
from .models_sp import SPAttention, SPMLP, SPBlock, SPModel, SPLMHeadModel

__all__ = [
    "LearnableFakeQuantize", "calibrate_many", "apply_minmax_quantization", "apply_log_quantization", "quantize_codes",
    "MinMaxQuantizationFunction", "LogQuantizationFunction", "LoRALayer", "SPLinearWithLoRA", "linear_fp",
    "SwitchableLayerNorm", "SPAttention", "SPMLP", "SPBlock", "SPModel", "SPLMHeadModel",
]
