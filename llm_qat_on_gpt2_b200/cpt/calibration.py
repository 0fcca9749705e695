"""Calibration driver of the CPT variant -- same entry points as part2_cyclic_precision_training/calibration.py
(`CalibrationManager` :8: `_calibrate_precision` :17, `ensure_calibrated` :85, `calibrate_gradient_quantizers` :98,
`calibrate_lora_weight_quantizers` :153), host logic only: it toggles the quantisers' state machines in upstream's
order and lets the kernels do the statistics.

Per width: weight quantisers on their own weights -> input quantisers over `num_batches` no-grad forwards with the LoRA
branch off -> the width's LoRA weight quantiser on A then B (one quantiser object sees both, :198-200).  The gradient
quantisers collect their statistics during ONE backward at 32 bits, exactly as upstream does (:117-131) -- at that
width CPTLinear runs without its adapter, so no LoRA gradient exists and they stay uncalibrated (identity), which is
upstream's observable behaviour.  `data_parallel_group`: MIN/MAX all-reduce of the input statistics (dp.py).
"""
from __future__ import annotations

from typing import Iterable, List

import torch

from .. import dp


class CalibrationManager:
    def __init__(self, model, train_loader, device, data_parallel_group=None):
        self.model = model
        self.train_loader = train_loader
        self.device = device
        self.group = data_parallel_group
        self.calibrated_bits = set()
        self.gradient_calibrated = False
        self.lora_calibrated_bits = set()

    def _linears(self):
        return [m for m in self.model.modules() if m.__class__.__name__ == 'CPTLinear']

    def _batches(self, n):
        it = iter(self.train_loader)
        for _ in range(n):
            try:
                yield next(it)
            except StopIteration:
                return

    def _calibrate_precision(self, bits: int, num_batches: int):
        if bits >= 32:
            return
        linears = self._linears()
        with torch.no_grad():
            for m in linears:
                qw = m.quantizer_weight
                qw.set_num_bits(bits)
                qw.start_calibration()
                qw(m.linear.weight.data)
                qw.finish_calibration(debug=False)
            for m in linears:
                m.quantizer_input.set_num_bits(bits)
                m.quantizer_input.start_calibration()
            self.model.disable_lora_for_calibration()
            for batch in self._batches(num_batches):
                self.model(batch['input_ids'].to(self.device, non_blocking=True))
            self.model.enable_lora_after_calibration()
            dp.finish_calibration_many([m.quantizer_input for m in linears], self.group)

    def ensure_calibrated(self, bits: int, num_batches: int = 10):
        if bits >= 32:
            return
        if bits not in self.calibrated_bits:
            self.model.set_precision(bits)
            self._calibrate_precision(bits, num_batches=num_batches)
            self.calibrated_bits.add(bits)
        if bits not in self.lora_calibrated_bits:
            self.calibrate_lora_weight_quantizers([bits])
            self.lora_calibrated_bits.add(bits)

    def calibrate_gradient_quantizers(self, precision: int = 32):
        """Upstream runs the collecting backward at 32 bits (:117).  `precision` other than 32 is an extension used by
        the tests to calibrate the gradient quantisers on real LoRA gradients."""
        quantizers = []
        for m in self.model.modules():
            if m.__class__.__name__ == 'LoRAAdapter':
                for q in (m.grad_quantizer_A, m.grad_quantizer_B):
                    if q is not None:
                        q.start_calibration()
                        quantizers.append(q)
        if not quantizers:
            return
        was_training, original = self.model.training, self.model.current_precision
        self.model.set_precision(precision)
        self.model.train()
        for batch in self._batches(1):
            ids = batch['input_ids'].to(self.device)
            labels = batch.get('labels', batch['input_ids']).to(self.device)
            self.model(ids, labels=labels).loss.backward()
            self.model.zero_grad()
        self.model.set_precision(original)
        if not was_training:
            self.model.eval()
        for q in quantizers:
            q.finish_calibration(debug=False)
        self.gradient_calibrated = True

    def calibrate_lora_weight_quantizers(self, bit_widths: Iterable[int]):
        for bits in bit_widths:
            if bits >= 32:
                continue
            key = f'{bits}bit'
            with torch.no_grad():
                for m in self._linears():
                    sl = getattr(m, 'shared_lora', None)
                    if sl is None or sl.lora_A is None or key not in m.lora_weight_quantizers:
                        continue
                    q = m.lora_weight_quantizers[key]
                    q.set_num_bits(bits)
                    q.start_calibration()
                    q(sl.lora_A)
                    q(sl.lora_B)
                    q.finish_calibration(debug=False)


__all__: List[str] = ["CalibrationManager"]
