"""Cyclic-precision schedule and precision range test (reference:
part2_cyclic_precision_training/cyclic_scheduler.py:5-123).  Host-side control of `model.set_precision(bits)`:
no device work of its own -- the forward passes it triggers run on the CUDA path of cpt_model / models_sp."""
import math
from typing import List, Tuple

import torch


class CyclicPrecisionScheduler:
    """Precision as a periodic function of the epoch, snapped to the configured bit widths (reference :5-43).
    `total_epochs / total_cycles` epochs form one cycle: cosine rises min -> max -> min within it, triangular is
    the piecewise-linear version."""

    def __init__(self, bit_widths: List[int] = [4, 6, 8], schedule_type: str = 'cosine', total_epochs: int = 160,
                 total_cycles: int = 32):
        self.bit_widths = sorted(bit_widths)
        self.min_bits = min(bit_widths)
        self.max_bits = max(bit_widths)
        self.schedule_type = schedule_type
        self.total_epochs = total_epochs
        self.total_cycles = total_cycles
        self.epochs_per_cycle = total_epochs / total_cycles
        self.global_cycle = 0
        self.cycle_count = 0
        self.current_epoch = 0

    def get_precision_for_epoch(self, epoch: int) -> int:
        phase = float(epoch % self.epochs_per_cycle) / self.epochs_per_cycle          # in [0, 1)
        span = self.max_bits - self.min_bits
        if self.schedule_type == 'cosine':
            value = self.min_bits + 0.5 * span * (1 - math.cos(phase * 2 * math.pi))
        elif self.schedule_type == 'triangular':
            value = self.min_bits + span * (2 * phase) if phase < 0.5 else self.max_bits - span * (2 * (phase - 0.5))
        else:
            raise ValueError(f"Unknown schedule type: {self.schedule_type}")
        return self._round_to_nearest_bitwidth(value)

    def _round_to_nearest_bitwidth(self, precision: float) -> int:
        gaps = [abs(precision - bw) for bw in self.bit_widths]
        return self.bit_widths[gaps.index(min(gaps))]                                  # first of equal gaps, as upstream


class PrecisionRangeTest:
    """Finds the lowest useful precision by sweeping bits upward and watching next-token accuracy (reference
    :45-123): returns the first width whose relative accuracy gain over the previous one exceeds `threshold`, or
    (after three widths) the first whose gain drops below 0.5 %."""

    EARLY_STOP = 0.005

    def __init__(self, model, start_bits: int, max_bits: int, threshold: float, test_iterations: int, target_bits: int):
        self.model = model
        self.start_bits = start_bits
        self.max_bits = max_bits
        self.threshold = threshold
        self.test_iterations = test_iterations
        self.target_bits = target_bits

    def _accuracy_at(self, bits, dataloader, device):
        self.model.set_precision(bits)
        hits = seen = batches = 0
        loss_sum = 0.0
        with torch.no_grad():
            for i, batch in enumerate(dataloader):
                if i >= self.test_iterations:
                    break
                ids = batch['input_ids'].to(device)
                labels = batch['labels'].to(device)
                mask_in = batch.get('attention_mask')
                if mask_in is not None:
                    mask_in = mask_in.to(device)
                out = self.model(ids, labels=labels, attention_mask=mask_in)
                loss_sum += out['loss'].item()
                batches += 1
                pred = out['logits'].argmax(dim=-1)[..., :-1]
                tgt = labels[..., 1:]
                valid = tgt != -100
                hits += ((pred == tgt) & valid).sum().item()
                seen += valid.sum().item()
        return {'accuracy': hits / seen if seen > 0 else 0, 'loss': loss_sum / batches if batches > 0 else float('inf')}

    def find_lower_bound(self, dataloader, criterion=None) -> int:
        self.model.train()
        device = next(self.model.parameters()).device
        metrics = {}
        for bits in range(self.start_bits, self.max_bits + 1):
            metrics[bits] = self._accuracy_at(bits, dataloader, device)
            if bits > self.start_bits:
                prev = metrics[bits - 1]['accuracy']
                gain = (metrics[bits]['accuracy'] - prev) / max(prev, 1e-6)
                if gain > self.threshold:
                    return bits
                if gain < self.EARLY_STOP and bits >= self.start_bits + 3:
                    return bits
        best, best_gain = self.start_bits, 0
        for bits in range(self.start_bits + 1, min(self.start_bits + 4, self.max_bits + 1)):
            if bits in metrics and bits - 1 in metrics:
                gain = metrics[bits]['accuracy'] - metrics[bits - 1]['accuracy']
                if gain > best_gain:
                    best, best_gain = bits, gain
        return best

    def find_bounds(self, dataloader, criterion=None) -> Tuple[int, int]:
        lower = min(self.find_lower_bound(dataloader, criterion), self.target_bits)
        return lower, min(self.target_bits + 4, self.max_bits)
