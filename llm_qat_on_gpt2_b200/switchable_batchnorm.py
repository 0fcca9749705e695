"""SwitchableLayerNorm on the B200 kernels -- drop-in for the reference class of the same name
(part1_switchable_precision/switchable_batchnorm.py:7-109): one (weight, bias) pair per
precision in `weights[str(p)]` / `biases[str(p)]`, `set_precision`, and the `ln_layers[str(p)]`
`.weight/.bias` `.data/.requires_grad` views the reference's weight loaders assign through
(p1/models_sp.py:347-357).  Forward and backward are one row-resident kernel each
(csrc/spq_layernorm.cu) instead of eight eager kernels plus autograd.
"""
from __future__ import annotations

from typing import List, Union

import torch
import torch.nn as nn

from . import _lib


class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps, ncols):
        if not x.is_cuda:
            raise RuntimeError("SwitchableLayerNorm runs on the CUDA kernels only (no CPU fallback); "
                               f"got a tensor on {x.device}")
        x2d = x.reshape(-1, ncols)
        if x2d.dtype != torch.float32:
            x2d = x2d.float()
        x2d = x2d.contiguous()
        rows = x2d.shape[0]
        w = weight.detach().reshape(-1).float().contiguous()
        b = bias.detach().reshape(-1).float().contiguous()
        y = torch.empty_like(x2d)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        if rows:
            _lib.layernorm_fwd(x2d, w, b, eps, y, mean, rstd)
        ctx.save_for_backward(x2d, w, mean, rstd)
        ctx.x_shape, ctx.w_shape = x.shape, weight.shape
        ctx.params = (weight, bias)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, gy):
        x2d, w, mean, rstd = ctx.saved_tensors
        rows, cols = x2d.shape
        g2d = gy.reshape(-1, cols)
        if g2d.dtype != torch.float32:
            g2d = g2d.float()
        g2d = g2d.contiguous()
        dx = torch.empty_like(x2d)
        need_p = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        from .lora import _grad_sink
        wp, bp = ctx.params
        sw = _grad_sink(wp, ctx.w_shape) if (ctx.needs_input_grad[1] and ctx.needs_input_grad[2]) else None
        sb = _grad_sink(bp, ctx.w_shape) if sw is not None else None
        if sw is not None and sb is not None:
            # the training driver owns the .grad buffers: the column sums are added straight into them -- on the driver's
            # side stream when it offers one (lora._GradSide: nothing downstream waits for the fold)
            from .lora import _GradSide
            if _GradSide.stream is not None:
                ws = _lib.layernorm_bwd_split(g2d, x2d, w, mean, rstd, dx)
                lane = _GradSide.fork(ws)
                with torch.cuda.stream(lane):
                    _lib.layernorm_bwd_finalize(ws, rows, cols, sw.view(-1), sb.view(-1), accumulate_params=True)
            else:
                _lib.layernorm_bwd(g2d, x2d, w, mean, rstd, dx, sw.view(-1), sb.view(-1), accumulate_params=True)
            return (dx.view(ctx.x_shape) if ctx.needs_input_grad[0] else None, None, None, None, None)
        dw = torch.empty(cols, dtype=torch.float32, device=gy.device) if need_p else None
        db = torch.empty(cols, dtype=torch.float32, device=gy.device) if need_p else None
        _lib.layernorm_bwd(g2d, x2d, w, mean, rstd, dx, dw, db)
        return (dx.view(ctx.x_shape) if ctx.needs_input_grad[0] else None,
                dw.view(ctx.w_shape) if ctx.needs_input_grad[1] else None,
                db.view(ctx.w_shape) if ctx.needs_input_grad[2] else None, None, None)


class _ParamView:
    """`.data` / `.requires_grad` window onto one entry of a ParameterDict."""

    def __init__(self, params: nn.ParameterDict, key: str):
        self._params, self._key = params, key

    @property
    def data(self):
        return self._params[self._key].data

    @data.setter
    def data(self, value):
        self._params[self._key].data = value

    @property
    def requires_grad(self):
        return self._params[self._key].requires_grad

    @requires_grad.setter
    def requires_grad(self, value):
        self._params[self._key].requires_grad = value


class _LayerNormView:
    """What `ln_layers[str(p)]` hands out: an object with `.weight` and `.bias` views."""

    def __init__(self, owner: "SwitchableLayerNorm", key: str):
        self._owner, self._key = owner, key

    @property
    def weight(self):
        return _ParamView(self._owner.weights, self._key)

    @property
    def bias(self):
        return _ParamView(self._owner.biases, self._key)


class SwitchableLayerNorm(nn.Module):
    def __init__(self, normalized_shape: Union[int, List[int], torch.Size],
                 precision_levels: List[int] = [6, 8, 16, 32], eps: float = 1e-5):
        super().__init__()
        if isinstance(normalized_shape, int):
            normalized_shape = (normalized_shape,)
        self.normalized_shape = tuple(normalized_shape)
        self.precision_levels = sorted(precision_levels)
        self.eps = eps
        self.weights = nn.ParameterDict()
        self.biases = nn.ParameterDict()
        for precision in self.precision_levels:
            self.weights[str(precision)] = nn.Parameter(torch.ones(normalized_shape))
            self.biases[str(precision)] = nn.Parameter(torch.zeros(normalized_shape))
        self.current_precision = max(self.precision_levels)
        self.ln_layers = {str(p): _LayerNormView(self, str(p)) for p in self.precision_levels}
        self._ncols = 1
        for d in self.normalized_shape:
            self._ncols *= int(d)

    def set_precision(self, precision: int) -> int:
        if precision not in self.precision_levels:
            raise ValueError(f"Precision {precision} not supported. Available: {self.precision_levels}")
        self.current_precision = precision
        return self.current_precision

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        key = str(self.current_precision)
        return _LayerNormFn.apply(x, self.weights[key], self.biases[key], self.eps, self._ncols)
