"""Multi-bit LearnableFakeQuantize and GradientQuantizer of the CPT variant on the B200 kernels --
drop-ins for part2_cyclic_precision_training/quantization.py (:14-26 GradientQuantizer, :28-300
LearnableFakeQuantize).

Differences from the part1 class that this mirrors: `scales[bits]` / `zero_points[bits]` dicts and
a `calibrated_bits` set instead of single `scale` / `zero_point` buffers (so precision can cycle
without recalibrating), a state_dict that flattens those dicts into `_scales_{b}` /
`_zero_points_{b}` / `_calibrated_bits` entries, `set_num_bits` that does not reset anything, and
a forward that silently passes x through at an uncalibrated width outside training.

The level index of the log quantiser is the same as part1's; part2 forms the dequantised value as
`L/(2n) + 0.5` without part1's `* (2^b - 1) / (2^b - 1)` round trip (p2/quantization_methods.py:40),
a <= 1 ulp difference in the exponent that is far inside the 1e-3 output tolerance, so the same
kernel serves both.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .. import _lib
from ..quantization import LearnableFakeQuantize as _SPFakeQuantize
from ..quantization_methods import apply_log_quantization, apply_minmax_quantization


class GradientQuantizer(torch.autograd.Function):
    """Identity forward; backward fake-quantises the gradient with `quantizer` when that quantiser is
    collecting statistics or is calibrated for its width (p2/quantization.py:14-26)."""

    @staticmethod
    def forward(ctx, input, quantizer):
        ctx.quantizer = quantizer
        return input

    @staticmethod
    def backward(ctx, grad_output):
        q = ctx.quantizer
        if q is not None and (q.collecting_stats or q.num_bits in q.calibrated_bits):
            return q(grad_output), None
        return grad_output, None


class LearnableFakeQuantize(_SPFakeQuantize):
    def __init__(self, num_bits, channel_dim=0, quantizer_type='minmax', eps=1e-5, symmetric=True,
                 per_channel=True, is_input=False):
        super().__init__(num_bits, channel_dim=channel_dim, quantizer_type=quantizer_type, eps=eps,
                         symmetric=symmetric, per_channel=per_channel, is_input=is_input)
        # the CPT class has no scale / zero_point buffers (p2/quantization.py:42-47)
        del self._buffers['scale']
        del self._buffers['zero_point']
        self.scales = {}
        self.zero_points = {}
        self.calibrated_bits = set()

    # `calibrated` of the part1 class maps onto the per-bits set
    @property
    def calibrated(self):
        return self.num_bits in self.calibrated_bits

    @calibrated.setter
    def calibrated(self, value):
        pass

    @property
    def scale(self):
        return self.scales[self.num_bits]

    @property
    def zero_point(self):
        return self.zero_points[self.num_bits]

    # ---------------------------------------------------------------- state_dict (p2 :54-141)
    def state_dict(self, *args, destination=None, prefix='', keep_vars=False, **kwargs):
        state = super().state_dict(*args, destination=destination, prefix=prefix, keep_vars=keep_vars, **kwargs)
        for bits, t in self.scales.items():
            state[f'{prefix}_scales_{bits}'] = t if keep_vars else t.clone()
        for bits, t in self.zero_points.items():
            state[f'{prefix}_zero_points_{bits}'] = t if keep_vars else t.clone()
        if self.calibrated_bits:
            state[f'{prefix}_calibrated_bits'] = list(self.calibrated_bits)
        return state

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        self.scales, self.zero_points, self.calibrated_bits = {}, {}, set()
        dev = self.running_min.device
        for key in [k for k in state_dict if k.startswith(prefix)]:
            suffix = key[len(prefix):]
            if suffix.startswith('_scales_') or suffix.startswith('_zero_points_'):
                is_scale = suffix.startswith('_scales_')
                try:
                    bits = int(suffix[len('_scales_'):] if is_scale else suffix[len('_zero_points_'):])
                except ValueError:
                    continue
                (self.scales if is_scale else self.zero_points)[bits] = state_dict.pop(key).clone().to(dev)
                if is_scale:
                    self.calibrated_bits.add(bits)
            elif suffix == '_calibrated_bits':
                v = state_dict.pop(key)
                if isinstance(v, list):
                    self.calibrated_bits = set(v)
        sk, zk = prefix + 'scale', prefix + 'zero_point'
        if sk in state_dict and zk in state_dict:           # a part1-style checkpoint
            self.scales[self.num_bits] = state_dict.pop(sk).clone().to(dev)
            self.zero_points[self.num_bits] = state_dict.pop(zk).clone().to(dev)
            self.calibrated_bits.add(self.num_bits)
        for name in ('running_min', 'running_max'):
            key = prefix + name
            if key in state_dict:
                t = state_dict[key]
                if self.is_input and t.dim() == 3 and t.shape[1] > 1:
                    take_min = self.quantizer_type != 'log' and 'min' in name
                    t = t.min(dim=1, keepdim=True)[0] if take_min else t.max(dim=1, keepdim=True)[0]
                    state_dict[key] = t
                getattr(self, name).resize_as_(t)
        nn.Module._load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys,
                                        unexpected_keys, error_msgs)
        self.generation += 1

    def _apply(self, fn, *a, **k):
        # .to(device) / .cuda() must move the per-bits tensors too (they are not registered buffers)
        out = super()._apply(fn, *a, **k)
        self.scales = {b: fn(t) for b, t in self.scales.items()}
        self.zero_points = {b: fn(t) for b, t in self.zero_points.items()}
        return out

    # ---------------------------------------------------------------- host state
    def set_num_bits(self, value):
        # p2 :143-146 -- switching width keeps every width's calibration
        self.num_bits = max(1, min(value, 32))
        self._update_quant_range()

    def start_calibration(self):
        self.collecting_stats = True
        self.num_batches_collected = 0
        self.temp_min = None
        self.temp_max = None
        self._first_shape = None
        if self._stat_state is not None:
            self._stat_state.zero_()

    def finish_calibration(self, debug=False):
        if self.num_batches_collected > 0 and self.temp_min is not None:
            tmin, tmax = self.temp_min, self.temp_max
            if self.stats_sync_hook is not None:
                self.stats_sync_hook(self, tmin, tmax)
            if self.quantizer_type == 'log':
                had_data = self._stat_flag_host if self._stat_flag_host is not None else int(self._stat_state.item())
            else:
                had_data = 1
            if not had_data:
                shape = self._log_default_shape()
                log_eps = float(np.log2(np.float64(np.float32(self.eps))).astype(np.float32))
                tmin = torch.full(shape, log_eps, dtype=torch.float32, device=tmin.device)
                tmax = tmin.clone()
            with torch.no_grad():
                self.running_min.resize_as_(tmin).copy_(tmin)
                self.running_max.resize_as_(tmax).copy_(tmax)
                scale, zp = torch.empty_like(tmin), torch.empty_like(tmin)
                qt = _lib.QTYPE.get(self.quantizer_type, _lib.MINMAX)
                _lib.finish_calibration(self.running_min, self.running_max, qt, self.symmetric, self.num_bits,
                                        self.eps, scale, zp)
                self.scales[self.num_bits] = scale
                self.zero_points[self.num_bits] = zp
            self.calibrated_bits.add(self.num_bits)
            self.collecting_stats = False
            self.temp_min = None
            self.temp_max = None
            self.generation += 1
        else:
            self.collecting_stats = False

    # ---------------------------------------------------------------- forward (p2 :257-285)
    def forward(self, x):
        if self.num_bits >= 32:
            return x
        if self.collecting_stats:
            self._collect_statistics_batch(x)
            return x
        if self.num_bits not in self.calibrated_bits:
            if self.training and torch.is_grad_enabled():
                raise RuntimeError(
                    f"FATAL: Quantizer not calibrated for {self.num_bits}-bit precision during training!\n"
                    f"  Calibrated bits: {self.calibrated_bits}\n"
                    f"  Available scales: {list(self.scales.keys())}\n"
                    f"  Available zero_points: {list(self.zero_points.keys())}\n"
                    f"  This indicates a bug in the calibration logic.\n"
                    f"  Training cannot proceed with uncalibrated quantizers.")
            return x
        scale, zero_point = self.scales[self.num_bits], self.zero_points[self.num_bits]
        if self.quantizer_type == 'minmax':
            return apply_minmax_quantization(x, scale, zero_point, self.num_bits, self.symmetric)
        elif self.quantizer_type == 'log':
            return apply_log_quantization(x, zero_point, scale, self.num_bits, self.symmetric)
        raise ValueError(f"Unknown quantizer type: {self.quantizer_type}. Supported types: 'minmax', 'log'")

    def ready(self) -> bool:
        return self.num_bits < 32 and not self.collecting_stats and self.num_bits in self.calibrated_bits and \
            self.quantizer_type in ('minmax', 'log')
