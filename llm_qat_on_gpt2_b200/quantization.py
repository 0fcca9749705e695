"""LearnableFakeQuantize on the B200 kernels -- drop-in for the reference class of the same name
(part1_switchable_precision/quantization.py:15-239; byte-identical copy in part5_squad).

Same constructor, attributes, buffers (`scale`, `zero_point`, `running_min`, `running_max`),
state machine (`start_calibration` / forward-in-collecting-mode / `finish_calibration`),
`set_num_bits`, checkpoint resize semantics and exceptions.  Device work:
  * statistics      -> spq_minmax_stats      (one streaming pass, no host sync per batch)
  * scale / zp      -> spq_finish_calibration (true IEEE division, as torch-CPU)
  * quantise        -> spq_fake_quantize      (bit-exact codes)
The log quantiser keeps its `any(|x| > eps)` test on the device; the host reads that flag once,
in `finish_calibration`, because the reference gives the statistics a different *shape* when no
batch ever exceeded eps (p1/quantization.py:164-172, 194-197).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .quantization_methods import apply_log_quantization, apply_minmax_quantization


class LearnableFakeQuantize(nn.Module):
    def __init__(self, num_bits, channel_dim=0, quantizer_type='minmax', eps=1e-5, symmetric=True,
                 per_channel=True, is_input=False):
        super().__init__()
        self.num_bits = max(1, min(num_bits, 32))
        self.symmetric = symmetric
        self.per_channel = per_channel
        self.channel_dim = channel_dim if per_channel else None
        self.quantizer_type = quantizer_type
        self.eps = eps
        self.is_input = is_input
        self._update_quant_range()

        self.register_buffer('scale', torch.ones(1))
        self.register_buffer('zero_point', torch.zeros(1))
        self.register_buffer('running_min', torch.zeros(1))
        self.register_buffer('running_max', torch.zeros(1))

        self.calibrated = False
        self.collecting_stats = False
        self.num_batches_collected = 0
        self.temp_min = None
        self.temp_max = None
        # not in the reference: bookkeeping for the device-side statistics and operand caches
        self._stat_state = None        # int32[1] on device: bit0 = some batch exceeded eps
        self._first_shape = None       # shape of the first collected tensor (log default-shape quirk)
        self.generation = 0            # bumped whenever scale / zero_point may have changed
        # optional hook: called with (temp_min, temp_max) before the scale is computed, so a
        # data-parallel driver can MIN/MAX all-reduce the statistics (see dp.py)
        self.stats_sync_hook = None
        self._stat_flag_host = None    # set by dp.finish_calibration_many to avoid a per-quantiser host read
        self.stats_stream = None       # optional torch.cuda.Stream for the statistics pass (training.GraphedCalibratedForward)
        self._stats_x = None

    # ---------------------------------------------------------------- checkpoint loading
    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys,
                              unexpected_keys, error_msgs):
        # reference :40-75 -- buffers take the checkpoint's shapes; legacy per-token input
        # statistics ([B, T>1, C]) are folded over dim 1
        for name in ('scale', 'zero_point', 'running_min', 'running_max'):
            key = prefix + name
            if key not in state_dict:
                continue
            buf = getattr(self, name, None)
            if buf is None:
                continue
            t = state_dict[key]
            if self.is_input and t.dim() == 3 and t.shape[1] > 1:
                take_min = self.quantizer_type != 'log' and 'min' in name
                t = t.min(dim=1, keepdim=True)[0] if take_min else t.max(dim=1, keepdim=True)[0]
                state_dict[key] = t
            buf.resize_as_(t)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys,
                                      unexpected_keys, error_msgs)
        if prefix + 'scale' in state_dict and prefix + 'zero_point' in state_dict:
            self.calibrated = True
        self.generation += 1

    # ---------------------------------------------------------------- host state
    def set_num_bits(self, value):
        old_bits = self.num_bits
        self.num_bits = max(1, min(value, 32))
        self._update_quant_range()
        if old_bits != self.num_bits:
            print(f"    Reset calibration for {self.quantizer_type} quantizer: {old_bits} -> {self.num_bits} bits")
            self.calibrated = False
            self.generation += 1

    def _update_quant_range(self):
        if self.symmetric:
            self.quant_min = -(2 ** (self.num_bits - 1))
            self.quant_max = 2 ** (self.num_bits - 1) - 1
        else:
            self.quant_min = 0
            self.quant_max = 2 ** self.num_bits - 1

    def start_calibration(self):
        self.collecting_stats = True
        self.calibrated = False
        self.num_batches_collected = 0
        self.temp_min = None
        self.temp_max = None
        self._first_shape = None
        if self._stat_state is not None:
            self._stat_state.zero_()

    # ---------------------------------------------------------------- calibration
    def _stat_layout(self, x):
        """(2-D view, broadcast mode, keepdim statistics shape) for this quantiser on x."""
        nd = x.dim()
        if self.per_channel and self.channel_dim is not None and nd > 0:
            cd = self.channel_dim if self.channel_dim >= 0 else nd + self.channel_dim
            shape = [1] * nd
            shape[cd] = x.shape[cd]
            if cd == nd - 1:
                return x.reshape(-1, x.shape[-1]), _lib.PER_COL, shape
            if cd == 0:
                return x.reshape(x.shape[0], -1), _lib.PER_ROW, shape
            # interior channel dim: bring it last (a copy; no reference call site does this)
            xp = x.movedim(cd, -1).contiguous()
            return xp.reshape(-1, xp.shape[-1]), _lib.PER_COL, shape
        return x.reshape(-1, x.shape[-1] if nd > 0 else 1), _lib.PER_TENSOR, [1] * nd

    def _collect_statistics_batch(self, x):
        if not x.is_cuda:
            raise RuntimeError("calibration statistics are computed on the CUDA kernels only (no CPU "
                               f"fallback); got a tensor on {x.device}")
        with torch.no_grad():
            xc = x.detach()
            keep_half = xc.dtype == torch.float16 and not (
                self.per_channel and self.channel_dim is not None and xc.dim() > 1
                and self.channel_dim % xc.dim() == 0)          # per-row (weight) statistics take float32
            if xc.dtype != torch.float32 and not keep_half:
                xc = xc.float()
            xc = xc.contiguous()
            if xc.numel() == 0:
                raise RuntimeError("cannot collect statistics of an empty tensor")
            x2d, bcast, stat_shape = self._stat_layout(xc)
            if self._stat_state is None or self._stat_state.device != xc.device:
                self._stat_state = torch.zeros(1, dtype=torch.int32, device=xc.device)
            first = self.temp_min is None
            if first:
                self.temp_min = torch.empty(stat_shape, dtype=torch.float32, device=xc.device)
                self.temp_max = torch.empty(stat_shape, dtype=torch.float32, device=xc.device)
                self._first_shape = tuple(xc.shape)
            side = self.stats_stream
            if side is None:
                _lib.minmax_stats(x2d, bcast, self.quantizer_type == 'log', self.eps, self.temp_min, self.temp_max,
                                  accumulate=not first, state=self._stat_state)
            else:
                # the statistics pass runs beside the caller's own pass over x (the row-scaling of the calibration
                # GEMM's operand): two streaming readers of the same tensor at the same time, the second one is served
                # by L2.  The caller joins with `join_stats()` before x can be released (SPLinearWithLoRA.forward)
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    _lib.minmax_stats(x2d, bcast, self.quantizer_type == 'log', self.eps, self.temp_min, self.temp_max,
                                      accumulate=not first, state=self._stat_state)
                self._stats_x = x2d                      # keeps the (possibly converted) input alive until the join
            self.num_batches_collected += 1

    def _stats_targets_lastdim(self, x_shape, device):
        """Bookkeeping of `_collect_statistics_batch` for a producer that computes the statistics itself (the
        LayerNorm-fused activation pass, lora.py): returns (temp_min, temp_max, accumulate, state) for a float32 input of
        shape `x_shape` whose channel dim is the LAST one, or None when this quantiser's layout is a different one."""
        nd = len(x_shape)
        if not (self.per_channel and self.channel_dim is not None and nd > 1):
            return None
        cd = self.channel_dim if self.channel_dim >= 0 else nd + self.channel_dim
        if cd != nd - 1:
            return None
        stat_shape = [1] * nd
        stat_shape[cd] = x_shape[cd]
        if self._stat_state is None or self._stat_state.device != device:
            self._stat_state = torch.zeros(1, dtype=torch.int32, device=device)
        first = self.temp_min is None
        if first:
            self.temp_min = torch.empty(stat_shape, dtype=torch.float32, device=device)
            self.temp_max = torch.empty(stat_shape, dtype=torch.float32, device=device)
            self._first_shape = tuple(x_shape)
        self.num_batches_collected += 1
        return self.temp_min, self.temp_max, (not first), self._stat_state

    def join_stats(self):
        """Make the current stream wait for a statistics pass issued on `stats_stream` (no-op otherwise)."""
        if self._stats_x is not None:
            torch.cuda.current_stream().wait_stream(self.stats_stream)
            self._stats_x = None

    def _log_default_shape(self):
        # reference :164-172: the CHANNEL dim is the one set to 1
        shape = list(self._first_shape)
        if self.per_channel and self.channel_dim is not None:
            cd = self.channel_dim if self.channel_dim >= 0 else len(shape) + self.channel_dim
            shape[cd] = 1
            return shape
        return []

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
                # no batch exceeded eps: the reference's statistics are log2(eps) in the
                # "default shape" (e.g. [r, 1] for a fresh all-zero lora_B of shape [r, N])
                shape = self._log_default_shape()
                log_eps = float(np.log2(np.float64(np.float32(self.eps))).astype(np.float32))
                tmin = torch.full(shape, log_eps, dtype=torch.float32, device=tmin.device)
                tmax = tmin.clone()
            with torch.no_grad():
                self.running_min.resize_as_(tmin).copy_(tmin)
                self.running_max.resize_as_(tmax).copy_(tmax)
                self.scale.resize_as_(tmin)
                self.zero_point.resize_as_(tmin)
                qt = _lib.QTYPE.get(self.quantizer_type, _lib.MINMAX)
                _lib.finish_calibration(self.running_min, self.running_max, qt, self.symmetric, self.num_bits,
                                        self.eps, self.scale, self.zero_point)
                if debug:
                    print(f"         Computed scale: mean={self.scale.mean().item():.6f}")
            self.calibrated = True
            self.collecting_stats = False
            self.temp_min = None
            self.temp_max = None
            self.generation += 1
        else:
            self.collecting_stats = False
            if debug:
                print(f"      ⚠️ No statistics collected for {self.num_bits}-bit {self.quantizer_type} quantizer")

    # ---------------------------------------------------------------- forward
    def forward(self, x):
        if self.num_bits >= 32:
            return x
        if self.collecting_stats:
            self._collect_statistics_batch(x)
            return x
        if not self.calibrated:
            raise RuntimeError(
                f"Quantizer not calibrated. Please run calibration first for {self.quantizer_type} quantizer.")
        if self.quantizer_type == 'minmax':
            return self._quantize_minmax(x)
        elif self.quantizer_type == 'log':
            return self._quantize_log(x)
        else:
            raise ValueError(
                f"Unknown quantizer type: {self.quantizer_type}. Supported types: 'minmax', 'log'")

    def _quantize_minmax(self, x):
        return apply_minmax_quantization(x, self.scale, self.zero_point, self.num_bits, self.symmetric)

    def _quantize_log(self, x):
        return apply_log_quantization(x, self.zero_point, self.scale, self.num_bits, self.symmetric)

    # ---------------------------------------------------------------- helpers for the fused linear
    def ready(self) -> bool:
        """True when forward would take the plain quantise branch (fused kernels may be used)."""
        return self.num_bits < 32 and not self.collecting_stats and self.calibrated and \
            self.quantizer_type in ('minmax', 'log')

    def abs_bound(self) -> torch.Tensor:
        """Per-channel upper bound of |q(x)| implied by the calibrated parameters (shape of scale)."""
        if self.quantizer_type == 'log':
            return torch.exp2(self.zero_point + self.scale.clamp(min=0))
        if self.symmetric:
            return self.scale * float(2 ** (self.num_bits - 1) - 1)
        hi = float(2 ** self.num_bits - 1)
        return torch.maximum((self.zero_point).abs(), (hi - self.zero_point).abs()) * self.scale


_calib_tables = {}


def calibrate_many(quantizers, tensors, defer: bool = False):
    """`q.start_calibration(); q(w); q.finish_calibration()` for every pair, in ONE kernel launch and one
    device->host read (spq_calibrate_many).  The reference recalibrates the LoRA A/B quantisers of all linears
    on their own weights every training step (p1/train_sp.py:125-163); one launch per quantiser phase made that
    part of the step host-bound.  Results are bit-identical to the per-quantiser calls (tests/test_gpu_modules.py).
    32-bit quantisers and log quantisers whose tensor has no |x| > eps take the single-tensor path.

    defer=True (CUDA-graph capture, see training.LoRARefresher): only the kernel is launched; the returned
    callable does the host part later (flag read, quantiser state) and returns the number of quantisers that
    had to take the single-tensor path."""
    import struct
    qs, ws = list(quantizers), list(tensors)
    if len(qs) != len(ws):
        raise ValueError("calibrate_many: one tensor per quantiser")
    fast, slow = [], []
    for q, w in zip(qs, ws):
        ok = (q.num_bits < 32 and w.is_cuda and w.dtype == torch.float32 and w.is_contiguous() and w.numel() > 0
              and q.stats_sync_hook is None and q.quantizer_type in ('minmax', 'log'))
        (fast if ok else slow).append((q, w))
    if defer and slow:
        raise RuntimeError("calibrate_many(defer=True): every quantiser must be eligible for the one-launch path")
    for q, w in slow:
        q.start_calibration(); q(w); q.finish_calibration()
    if not fast:
        return (lambda: 0) if defer else None
    dev = fast[0][1].device
    layouts = []
    with torch.no_grad():
        for q, w in fast:
            wd = w.detach()
            x2d, bcast, stat_shape = q._stat_layout(wd)
            if x2d.data_ptr() != wd.data_ptr():          # interior channel dim (copy): not a weight layout
                layouts.append(None)
                continue
            for name in ('running_min', 'running_max', 'scale', 'zero_point'):
                buf = getattr(q, name)
                if list(buf.shape) != list(stat_shape):
                    buf.resize_(stat_shape)
            layouts.append((x2d, bcast))
    key = tuple((w.data_ptr(), tuple(w.shape), q.running_min.data_ptr(), q.running_max.data_ptr(), q.scale.data_ptr(),
                 q.zero_point.data_ptr(), q.num_bits, q.symmetric, q.quantizer_type, q.eps)
                for (q, w), lay in zip(fast, layouts) if lay is not None)
    ent = _calib_tables.get(key)
    if ent is None:
        rows_b, max_blocks = [], 1
        for (q, w), lay in zip(fast, layouts):
            if lay is None:
                continue
            x2d, bcast = lay
            rows, cols = x2d.shape
            levels = float(2 ** (q.num_bits - 1) - 1) if q.symmetric else float(2 ** q.num_bits - 1)
            blocks = (cols + 31) // 32 if bcast == _lib.PER_COL else (rows + 7) // 8 if bcast == _lib.PER_ROW else 1
            max_blocks = max(max_blocks, blocks)
            rows_b.append(struct.pack(_lib.CALIB_JOB_FORMAT, x2d.data_ptr(), rows, cols, q.running_min.data_ptr(),
                                      q.running_max.data_ptr(), q.scale.data_ptr(), q.zero_point.data_ptr(), bcast,
                                      _lib.QTYPE[q.quantizer_type], int(q.symmetric), levels, float(q.eps), 0))
        if defer and torch.cuda.is_current_stream_capturing():
            # building the job table is a pageable host->device copy: not capturable.  LoRARefresher runs the same
            # call eagerly before it captures, so a miss here means the addresses changed in between
            raise RuntimeError("calibrate_many(defer=True): job table not built yet; run the call once outside "
                               "the CUDA-graph capture")
        table = torch.frombuffer(bytearray(b"".join(rows_b)), dtype=torch.uint8).to(dev)
        ent = (table, len(rows_b), max_blocks, torch.zeros(len(rows_b), dtype=torch.int32, device=dev))
        while len(_calib_tables) >= 64:
            # evict the oldest entry only; a captured graph that launches with an evicted table keeps it alive
            # through the `finish` closure it holds (tables carry raw device pointers)
            _calib_tables.pop(next(iter(_calib_tables)))
        _calib_tables[key] = ent
    table, n_jobs, max_blocks, flags = ent
    if n_jobs:
        _lib.calibrate_many(table, n_jobs, max_blocks, flags)

    def finish(redo: bool = True, _keep=ent) -> int:      # _keep: the job table outlives any cache eviction
        """redo=False (whole-step CUDA graphs, training.SPTrainer): a log quantiser whose tensor had nothing above
        eps keeps what the kernel wrote -- log2(eps) statistics and a zero range in the per-channel layout, the same
        VALUES as the reference's default-shape quirk (p1/quantization.py:164-172, 194-197), so every dequantised
        number is identical; only the buffer shape differs until the next eager calibration."""
        got = flags.tolist() if n_jobs else []            # the one device->host read
        it = iter(got)
        redone = 0
        for (q, w), lay in zip(fast, layouts):
            had = next(it) if lay is not None else 0
            if lay is None or (not had and redo):
                q.start_calibration(); q(w); q.finish_calibration()       # no data above eps / odd layout
                redone += 1
                continue
            if not had:
                redone += 1
            q.calibrated = True
            q.collecting_stats = False
            q.num_batches_collected = 1
            q.temp_min = q.temp_max = None
            q._first_shape = tuple(w.shape)
            q.generation += 1
        return redone

    if defer:
        return finish
    finish()
    return None


def pow2_ceil(t: torch.Tensor) -> torch.Tensor:
    """Smallest power of two >= t (elementwise, t > 0; zeros and non-finite map to 1)."""
    safe = torch.where(torch.isfinite(t) & (t > 0), t, torch.ones_like(t))
    return torch.exp2(torch.ceil(torch.log2(safe)))


__all__ = ["LearnableFakeQuantize", "pow2_ceil"]
_ = math
