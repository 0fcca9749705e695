"""Training-step helper for the switchable-precision path (host side).

The reference recalibrates the LoRA A/B quantisers of every SPLinearWithLoRA on their own weights before each
student step (p1/train_sp.py:125-163, 362-364).  Here that costs, per linear, one calibration, two dequantise
launches, one scale-preparation launch and four operand builds: ~350 small launches per step for GPT-2 small,
all with static shapes and addresses.  `LoRARefresher` captures them once into a CUDA graph and replays it: one
graph launch and one device->host flag read per step.  Tensors produced inside the graph live in its private
pool and are rewritten in place by every replay, so call `refresh()` only between optimizer steps (no autograd
graph of an earlier forward may still be waiting for its backward).
"""
from typing import List

import torch

from .quantization import calibrate_many


class LoRARefresher:
    def __init__(self, linears: List[torch.nn.Module], bits: int, with_backward_operands: bool = True):
        self.linears = [m for m in linears if m.__class__.__name__ == 'SPLinearWithLoRA']
        self.bits = bits
        self.key = f'{bits}bit'
        self.with_bwd = with_backward_operands
        self.graph = None
        self.finish = None
        self._sig = None
        adapters = [m.lora_adapters[self.key] for m in self.linears]
        self.active = [(m, lo) for m, lo in zip(self.linears, adapters) if lo.enabled and lo.scaling != 0]
        self.quantizers = [q for _, lo in self.active for q in (lo.quantize_A, lo.quantize_B)]
        self.weights = [w for _, lo in self.active for w in (lo.lora_A, lo.lora_B)]

    # ------------------------------------------------------------------------------------------
    def _signature(self):
        # everything the captured launches hard-wire: parameter / buffer addresses and the static cache levels
        sig = []
        for m, lo in self.active:
            qi = m.quantizers_input[self.key]
            sig.append((lo.lora_A.data_ptr(), lo.lora_B.data_ptr(), lo.quantize_A.scale.data_ptr(),
                        lo.quantize_B.scale.data_ptr(), qi.generation, qi.scale.data_ptr(),
                        m.quantizers_weight[self.key].generation, m.linear.weight.data_ptr(), m.linear.weight._version))
        return tuple(sig)

    def _build_all(self):
        for m, _ in self.active:
            m._operands_for(self.bits, True)
            if self.with_bwd:
                m._backward_operands_for(self.bits, True)

    def _eager(self):
        calibrate_many(self.quantizers, [w.data for w in self.weights])
        self._build_all()

    def refresh(self) -> None:
        """Recalibrate the active LoRA quantisers on the current LoRA weights and rebuild every operand that
        depends on them, for all linears."""
        if not self.active:
            return
        with torch.no_grad():
            sig = self._signature()
            if self.graph is None or sig != self._sig:
                # (re)capture: a first eager pass settles buffer shapes and the input / weight cache levels
                self._eager()
                self._eager()
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                for m, _ in self.active:                      # force a rebuild of the LoRA levels inside the capture
                    m._op_cache.get(self.bits, {}).pop('lora', None)
                    ent = m._op_cache.get(self.bits, {}).get('input')
                    if ent is not None:
                        ent['lora'] = None
                with torch.cuda.graph(g):
                    fin = calibrate_many(self.quantizers, [w.data for w in self.weights], defer=True)
                    for q in self.quantizers:                 # the captured rebuild must see "calibrated" quantisers
                        q.calibrated, q.collecting_stats = True, False
                    self._build_all()
                self.graph, self.finish, self._sig = g, fin, self._signature()
            self.graph.replay()
            redone = self.finish()
            if redone:
                # a log quantiser without data (fresh all-zero lora_B) went through the reference's default-shape
                # path on the host side: its scales are not the ones the replay used -- rebuild eagerly this once
                self._build_all()
                self._sig = None
            else:
                for m, _ in self.active:
                    m._restamp_lora_keys(self.bits, self.with_bwd)
