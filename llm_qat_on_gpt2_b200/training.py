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


class _DistillKL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, s_logits, t_logits, temperature):
        from . import _lib
        B, T, V = s_logits.shape
        rows = B * (T - 1)

        def as2d(x):
            # [B, T, V] -> [B*T, V] without a copy, also for the row-padded views the LM head returns
            if x.dtype != torch.float32:
                x = x.float()
            if x.stride(-1) == 1 and x.stride(0) == T * x.stride(1):
                return x.as_strided((B * T, V), (x.stride(1), 1))
            return x.reshape(B * T, V).contiguous()
        s2d, t2d = as2d(s_logits.detach()), as2d(t_logits.detach())
        row_loss, grad = _lib.distill_kl(s2d, t2d, temperature, T, temperature * temperature / rows,
                                         want_grad=ctx.needs_input_grad[0])
        ctx.save_for_backward(grad)
        ctx.shape = s_logits.shape
        return row_loss.sum() * (temperature * temperature / rows)

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (grad * g).view(ctx.shape) if grad is not None else None, None, None


def distillation_kl_loss(student_logits, teacher_logits, temperature: float = 3.0):
    """`F.kl_div(log_softmax(s[:, :-1] / T), log_softmax(t[:, :-1] / T), reduction='batchmean', log_target=True)
    * T**2` of the reference's distillation step (p1/distillation_manager.py:64-80), value and gradient from ONE
    kernel (spq_distill_kl) instead of ~10 elementwise passes over the [B, T, V] logits.  Logits may be the
    row-padded views the LM head returns."""
    if not (student_logits.is_cuda and student_logits.dim() == 3 and teacher_logits.shape == student_logits.shape):
        raise RuntimeError("distillation_kl_loss expects CUDA logits of shape [B, T, V] (no CPU fallback)")
    return _DistillKL.apply(student_logits, teacher_logits, float(temperature))


def distillation_loss(student_outputs, teacher_outputs, temperature: float = 3.0, alpha_kl: float = 1.0,
                      alpha_feature: float = 1e-7, feature_layers=None, accumulative: bool = False, rng=None):
    """`DistillationManager.compute_distillation_loss` (p1/distillation_manager.py:64-116) on output dicts with
    'logits' and (optionally) 'hidden_states': alpha_kl * KL(T) [fused kernel] + alpha_feature * MSE of the hidden
    states of one randomly chosen layer (or the mean over `feature_layers` when accumulative).  The models return
    detached hidden-state copies, as upstream (p1/models_sp.py:323), so the feature term carries no gradient."""
    import random
    import torch.nn.functional as F
    kl = distillation_kl_loss(student_outputs['logits'], teacher_outputs['logits'], temperature)
    feature = None
    hs, ht = student_outputs.get('hidden_states'), teacher_outputs.get('hidden_states')
    if hs and ht:
        n = min(len(hs), len(ht))
        layers = [l for l in (feature_layers or range(n)) if l < n]
        if layers:
            if accumulative:
                feature = sum(F.mse_loss(hs[l], ht[l], reduction='mean') for l in layers) / len(layers)
            else:
                l = (rng or random).choice(layers)
                feature = F.mse_loss(hs[l], ht[l], reduction='mean')
    return alpha_kl * kl if feature is None else alpha_kl * kl + alpha_feature * feature


class GraphedNoGradForward:
    """`model(ids, **kwargs)` under no_grad with fixed shapes -- the 32-bit teacher forward of the distillation step
    (p1/distillation_manager.py:34-62) -- replayed as ONE CUDA graph: ~200 launches of the eager forward become one.
    The precision must be set before the first call (the captured launches are those of that precision) and the
    returned tensors are static buffers, overwritten by the next call.

    The capture hard-wires the fp16 weight operands the model's caches held at capture time.  Upstream trains the
    base linears and every LayerNorm pair at the teacher width (models_sp.unfreeze_weights(32)), so the signature
    covers (data_ptr, _version) of every parameter and buffer the forward reads: any optimizer update, `.data`
    assignment or recalibration recaptures, and the graph keeps its own references to the cached operands so that
    an eager forward rebuilding a cache in between cannot free memory a replay would read."""

    def __init__(self, model, **kwargs):
        self.model, self.kwargs = model, kwargs
        self.graph = None
        self.ids = None
        self.out = None
        self._sig = None
        self._keep = None

    def _signature(self, ids):
        prec = self.model.get_current_precision() if hasattr(self.model, 'get_current_precision') else None
        tensors = tuple((t.data_ptr(), t._version) for t in self.model.parameters())
        gens = tuple(m.generation for m in self.model.modules() if hasattr(m, 'generation'))
        return (tuple(ids.shape), ids.dtype, self.model.training, prec, tensors, gens)

    def _cached_operands(self):
        keep = []
        for m in self.model.modules():
            for attr in ('_fp_cache', '_lm_head_cache'):
                c = getattr(m, attr, None)
                if c is not None:
                    keep.append((c.fwd, c.bwd))
            oc = getattr(m, '_op_cache', None)
            if oc:
                keep.append({b: dict(ent) for b, ent in oc.items()})
            cc = getattr(m, '_calib_cache', None)
            if cc:
                keep.append(dict(cc))
        return keep

    def __call__(self, ids):
        sig = self._signature(ids)
        if self.graph is None or sig != self._sig:
            self.ids = ids.clone()
            with torch.no_grad():
                for _ in range(2):                         # warm-up: operand caches, cuDNN plans, allocator
                    self.model(self.ids, **self.kwargs)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self.out = self.model(self.ids, **self.kwargs)
            self.graph, self._sig, self._keep = g, sig, self._cached_operands()
        self.ids.copy_(ids)
        self.graph.replay()
        return self.out
