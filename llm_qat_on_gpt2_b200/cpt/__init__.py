"""Cyclic-precision-training (part2) variant of the path on the same B200 kernels.

Drop-ins for part2_cyclic_precision_training/{quantization.py, cpt_model.py}: the multi-bit
LearnableFakeQuantize (one object holds scales / zero_points per bit-width), GradientQuantizer,
LoRAAdapter, CPTLinear and the GPT-2 wrapper classes around them.
"""
from .quantization import GradientQuantizer, LearnableFakeQuantize
from .cpt_model import LoRAAdapter, CPTLinear, CPTSelfAttention, CPTMLP, CPTBlock, CPTModel

from .cyclic_scheduler import CyclicPrecisionScheduler, PrecisionRangeTest
from .calibration import CalibrationManager
from .training import CPTTrainer

__all__ = ["CyclicPrecisionScheduler", "PrecisionRangeTest", "CalibrationManager", "CPTTrainer", "GradientQuantizer", "LearnableFakeQuantize", "LoRAAdapter", "CPTLinear", "CPTSelfAttention",
           "CPTMLP", "CPTBlock", "CPTModel"]
