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

from .lora import side_stream_grads
from .quantization import calibrate_many


class LoRARefresher:
    def __init__(self, linears: List[torch.nn.Module], bits: int, with_backward_operands: bool = True, n_side: int = 8):
        self.linears = [m for m in linears if m.__class__.__name__ == 'SPLinearWithLoRA']
        self.bits = bits
        self.key = f'{bits}bit'
        self.with_bwd = with_backward_operands
        self.graph = None
        self.finish = None
        self._sig = None
        self._side = None
        self.n_side = n_side
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
        # per linear: two dequantise launches, one scale-preparation launch, four operand builds -- a few CTAs each and
        # independent across linears.  Round-robin on side streams: parallel branches when captured into a graph
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = [torch.cuda.Stream() for _ in range(self.n_side)]
        lanes = self._side if self._side else [main]
        for st in self._side:
            st.wait_stream(main)
        for i, (m, _) in enumerate(self.active):
            with torch.cuda.stream(lanes[i % len(lanes)]):
                m._operands_for(self.bits, True)
                if self.with_bwd:
                    m._backward_operands_for(self.bits, True)
        for st in self._side:
            main.wait_stream(st)

    def _eager(self):
        calibrate_many(self.quantizers, [w.data for w in self.weights])
        self._build_all()

    # pieces a larger capture (SPTrainer's whole micro-step graph) composes ------------------------------------
    def invalidate(self):
        """Drop the LoRA-level cache entries so that the next `body()` rebuilds them (inside a capture: into the
        graph's private pool, where every replay rewrites them in place)."""
        for m, _ in self.active:
            m._op_cache.get(self.bits, {}).pop('lora', None)
            ent = m._op_cache.get(self.bits, {}).get('input')
            if ent is not None:
                ent['lora'] = None

    def body(self):
        """The device work of one refresh, host-sync free (capturable): one calibration launch for all LoRA
        quantisers + the operand rebuild.  Returns the deferred host part (flag read, quantiser bookkeeping)."""
        with torch.no_grad():
            fin = calibrate_many(self.quantizers, [w.data for w in self.weights], defer=True)
            for q in self.quantizers:                     # the rebuild must see "calibrated" quantisers
                q.calibrated, q.collecting_stats = True, False
            self._build_all()
        return fin

    def restamp(self):
        for m, _ in self.active:
            m._restamp_lora_keys(self.bits, self.with_bwd)

    def replay(self):
        """The device part of `refresh()` as one graph replay (captured on first use / when an address or a static
        cache level changed); returns the deferred host part (`finish(redo=...)`, one device->host flag read)."""
        if not self.active:
            return lambda redo=True: 0
        with torch.no_grad():
            sig = self._signature()
            if self.graph is None or sig != self._sig:
                # (re)capture: a first eager pass settles buffer shapes and the input / weight cache levels
                self._eager()
                self._eager()
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                self.invalidate()                             # force a rebuild of the LoRA levels inside the capture
                with torch.cuda.graph(g):
                    fin = self.body()
                self.graph, self.finish, self._sig = g, fin, self._signature()
            self.graph.replay()
            self.restamp()
        return self.finish

    def refresh(self) -> None:
        """Recalibrate the active LoRA quantisers on the current LoRA weights and rebuild every operand that
        depends on them, for all linears."""
        if not self.active:
            return
        redone = self.replay()()
        if redone:
            # a log quantiser without data (fresh all-zero lora_B) went through the reference's default-shape
            # path on the host side: its scales are not the ones the replay used -- rebuild eagerly this once
            with torch.no_grad():
                self._build_all()
            self._sig = None


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


class _LMHeadLossFn(torch.autograd.Function):
    """loss(lm_head(hidden)) with the LM-head backward fed by the loss kernel's fp16 gradient operand: the float32
    [B, T, V] dlogits matrix, its row-scaling pass and torch's softmax passes never exist."""

    @staticmethod
    def forward(ctx, hidden, weight, cache, kind, other, temperature):
        from . import _lib
        from .lora import _as_2d_act, _rowscaled_f16
        V, C = weight.shape
        B, T = hidden.shape[0], hidden.shape[1]
        x2d = _as_2d_act(hidden, C, max_cols=8192)
        M = x2d.shape[0]
        x16, rs = _rowscaled_f16(x2d)
        w16, pw = cache.get(weight, transposed=False)
        ld = (V + 31) // 32 * 32
        logits = torch.empty((M, ld), dtype=torch.float32, device=hidden.device)
        s2d = logits[:, :V]
        _lib.qgemm(x16, w16, M, V, C, s2d, row_scale=rs, col_scale=pw)
        if kind == 'kl':
            t = other
            if t.dtype != torch.float32:
                t = t.float()
            if t.stride(-1) == 1 and t.stride(0) == T * t.stride(1):
                t2d = t.as_strided((M, V), (t.stride(1), 1))              # the row-padded view the LM head returns
            else:
                t2d = t.reshape(M, V).contiguous()
            row_loss, _, g16, eg, _ = _lib.softmax_loss_grad16(s2d, 'kl', t2d=t2d, temperature=temperature, seq_len=T)
            rows = B * (T - 1)
            loss = row_loss.sum() * (temperature * temperature / rows)
            factor = torch.full((1,), temperature / rows, dtype=torch.float32, device=hidden.device)
        else:
            labels = other
            targets = torch.full_like(labels, -100)
            targets[..., :-1] = labels[..., 1:]                             # position t is scored against token t+1
            row_loss, row_valid, g16, eg, _ = _lib.softmax_loss_grad16(s2d, 'ce', targets=targets)
            factor = (1.0 / row_valid.sum()).reshape(1)
            loss = row_loss.sum() * factor[0]
        ctx.cache, ctx.x_shape, ctx.dims = cache, hidden.shape, (M, V, C)
        ctx.save_for_backward(g16, eg, factor, weight)
        return loss

    @staticmethod
    def backward(ctx, gout):
        from . import _lib
        g16, eg, factor, weight = ctx.saved_tensors
        M, V, C = ctx.dims
        gx = None
        if ctx.needs_input_grad[0]:
            wt16, pk = ctx.cache.get(weight, transposed=True)
            gx = torch.empty((M, C), dtype=torch.float32, device=g16.device)
            _lib.qgemm(g16, wt16, M, C, V, gx, row_scale=eg * (gout * factor), col_scale=pk)
            gx = gx.view(ctx.x_shape)
        return gx, None, None, None, None, None


def lm_head_loss(model, hidden, *, labels=None, teacher_logits=None, temperature: float = 3.0):
    """Loss of the tied LM head on the final hidden states [B, T, C] of an SPLMHeadModel:
      labels given          -> the next-token cross-entropy of SPLMHeadModel.forward (p1/models_sp.py:441-449);
      teacher_logits given  -> KL(T) * T^2 of DistillationManager.compute_distillation_loss (p1/distillation_manager.py:64-80).
    Same values as `F.cross_entropy` / `distillation_kl_loss` on `model.lm_head(hidden)`; gradient flows to `hidden`
    only (the tied embedding is frozen on this path)."""
    if (labels is None) == (teacher_logits is None):
        raise ValueError("lm_head_loss: give either labels or teacher_logits")
    if not hidden.is_cuda:
        raise RuntimeError("lm_head_loss runs on the CUDA kernels only (no CPU fallback)")
    w = model.lm_head.weight
    if torch.is_grad_enabled() and w.requires_grad:
        raise RuntimeError("lm_head_loss does not produce a gradient for the (tied) LM-head weight; freeze it or use model(...)")
    if labels is not None:
        return _LMHeadLossFn.apply(hidden, w, model._lm_head_cache, 'ce', labels, 1.0)
    return _LMHeadLossFn.apply(hidden, w, model._lm_head_cache, 'kl', teacher_logits, float(temperature))


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
        self._params_for = {}

    def _read_params(self, prec):
        """Parameters a forward at `prec` reads: everything except the LoRA adapters and LayerNorm pairs of OTHER
        precisions (those change every optimizer step of a student width and must not force a recapture)."""
        got = self._params_for.get(prec)
        if got is None:
            got = []
            for n, p in self.model.named_parameters():
                parts = n.split('.')
                if 'lora_adapters' in parts and (prec is None or prec >= 32 or f'{prec}bit' not in parts):
                    continue
                if len(parts) >= 2 and parts[-2] in ('weights', 'biases') and prec is not None and parts[-1] != str(prec):
                    continue
                got.append(p)
            self._params_for[prec] = got
        return got

    def _signature(self, ids):
        prec = self.model.get_current_precision() if hasattr(self.model, 'get_current_precision') else None
        tensors = tuple((t.data_ptr(), t._version) for t in self._read_params(prec))
        gens = tuple(m.generation for m in self.model.modules() if hasattr(m, 'generation')) if (prec or 32) < 32 else ()
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


class GraphedCalibratedForward:
    """BASELINE configs[1] as two CUDA-graph replays: the calibration pass of upstream's CalibrationManager
    (p1/train_sp.py:47-83 -- start_calibration() on every input quantiser, LoRA disabled, one forward of the
    transformer body on the batch, finish_calibration()) followed by `model(ids, labels=ids)` at the same width.

    Eagerly that step is ~630 launches of 5-200 us kernels with one device->host read in the middle (the log
    quantisers' any(|x| > eps) flags), after which the host has to refill an empty launch queue.  Here:

      graph 1   statistics pass: every input quantiser's min / max (log2 domain for log quantisers) into static
                buffers, plus the packed MIN / MAX / flag buckets when data parallel
      (eager)   world > 1 only: three NCCL all-reduces on the packed buckets -- no host synchronisation
      graph 2   unpack, finish_calibration of every quantiser, the operand rebuild that depends on the new input
                scales (prep_linear_scales, W / LoRA operands), the quantised forward, the LM head and the loss

    The flag read is deferred: a log quantiser whose input had nothing above eps keeps what the statistics kernel
    wrote (log2(eps), zero range, per-channel layout -- the same dequantised VALUES as upstream's default-shape
    path, p1/quantization.py:164-172, 194-197); `nodata_count()` reports how many quantiser calibrations took that
    path since construction (one device->host read, outside the step).

    Returned tensors are static buffers, overwritten by the next call.  The capture hard-wires parameter / buffer
    addresses and the weight- and LoRA-level operand caches: the signature covers (data_ptr, _version) of every
    parameter the forward reads and the generation of every weight / LoRA quantiser; a change recaptures."""

    def __init__(self, model, group=None, with_labels: bool = True, n_side: int = 8, overlap_stats: bool = True):
        self.model, self.group, self.with_labels = model, group, with_labels
        self.linears = [m for m in model.modules() if m.__class__.__name__ == 'SPLinearWithLoRA']
        self.world = torch.distributed.get_world_size(group) if (torch.distributed.is_available()
                                                                  and torch.distributed.is_initialized()) else 1
        self.g1 = self.g2 = None
        self.ids = None
        self.out = None
        self._sig = None
        self._keep = None
        self._buckets = None
        self._nodata = None
        self._temps = None
        self._side = None
        self._stats_stream = None
        self.n_side = n_side
        self.overlap_stats = overlap_stats
        self.kernels_per_replay = 0
        self._helper = GraphedNoGradForward(model)        # parameter bookkeeping shared with the teacher graph

    def _input_quantizers(self, bits):
        return [m.quantizers_input[f'{bits}bit'] for m in self.linears]

    def _signature(self, ids):
        bits = self.model.get_current_precision()
        key = f'{bits}bit'
        tensors = tuple((t.data_ptr(), t._version) for t in self._helper._read_params(bits))
        gens = []
        for m in self.linears:
            lo = m.lora_adapters[key]
            qi = m.quantizers_input[key]
            gens.append((m.quantizers_weight[key].generation, lo.quantize_A.generation, lo.quantize_B.generation,
                         lo.enabled, qi.scale.data_ptr(), qi.zero_point.data_ptr(), qi.running_min.data_ptr(),
                         qi.running_max.data_ptr(), tuple(qi.scale.shape)))
        return (tuple(ids.shape), ids.dtype, self.model.training, bits, tensors, tuple(gens))

    # -------------------------------------------------------------------------------- the two bodies
    def _stats_body(self, qs):
        m = self.model
        if self.overlap_stats and self._stats_stream is None:
            self._stats_stream = torch.cuda.Stream()
        for q in qs:
            q.start_calibration()
            q.stats_stream = self._stats_stream if self.overlap_stats else None
        m.disable_lora_for_calibration()
        try:
            m.transformer(self.ids)                          # statistics only need the transformer body
        finally:
            m.enable_lora_after_calibration()
            for q in qs:
                q.join_stats()
                q.stats_stream = None
        live = [q for q in qs if q.temp_min is not None]
        self._temps = [(q.temp_min, q.temp_max) for q in live]       # finish_calibration drops the quantisers' references
        flags = [q._stat_state for q in live if q.quantizer_type == 'log' and q._stat_state is not None]
        fl = torch.cat(flags) if flags else None
        if self.world > 1:
            mins = torch.cat([q.temp_min.reshape(-1) for q in live])
            maxs = torch.cat([q.temp_max.reshape(-1) for q in live])
        else:
            mins = maxs = None
        return live, mins, maxs, fl

    def _forward_body(self, qs, live, mins, maxs, fl):
        if self.world > 1:
            off = 0
            for q in live:
                n = q.temp_min.numel()
                q.temp_min.copy_(mins[off:off + n].view_as(q.temp_min))
                q.temp_max.copy_(maxs[off:off + n].view_as(q.temp_max))
                off += n
        if fl is not None:
            self._nodata += (fl == 0).sum()
        # per linear: finish_calibration -> prep_linear_scales -> W / LoRA operand builds: five launches of one to a
        # few CTAs each, independent across the 48 linears.  Issued round-robin on side streams they become parallel
        # branches of the captured graph (~2 ms of serialised small kernels per step otherwise); the forward below
        # finds every operand cache current
        bits = self.model.get_current_precision()
        key = f'{bits}bit'
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = [torch.cuda.Stream() for _ in range(self.n_side)]
        lanes = self._side if self._side else [main]         # n_side = 0: everything in line (A/B switch)
        for st in self._side:
            st.wait_stream(main)
        for i, (m, q) in enumerate(zip(self.linears, qs)):
            with torch.cuda.stream(lanes[i % len(lanes)]):
                hook, q.stats_sync_hook = q.stats_sync_hook, None
                q._stat_flag_host = 1                        # deferred (class docstring)
                try:
                    q.finish_calibration()
                finally:
                    q.stats_sync_hook = hook
                    q._stat_flag_host = None
                lo = m.lora_adapters[key]
                if q.ready() and m.quantizers_weight[key].ready():
                    m._operands_for(bits, bool(lo.enabled and lo.scaling != 0))
        for st in self._side:
            main.wait_stream(st)
        if self.with_labels:
            return self.model(self.ids, labels=self.ids)
        return self.model(self.ids)

    def _exchange(self, mins, maxs, fl):
        import torch.distributed as dist
        dist.all_reduce(mins, op=dist.ReduceOp.MIN, group=self.group)
        dist.all_reduce(maxs, op=dist.ReduceOp.MAX, group=self.group)
        if fl is not None:
            dist.all_reduce(fl, op=dist.ReduceOp.MAX, group=self.group)

    def _eager(self, qs):
        live, mins, maxs, fl = self._stats_body(qs)
        if self.world > 1:
            self._exchange(mins, maxs, fl)
        return self._forward_body(qs, live, mins, maxs, fl)

    # -------------------------------------------------------------------------------- call
    def __call__(self, ids):
        from . import _lib
        sig = self._signature(ids)
        if self.g1 is None or sig != self._sig:
            bits = self.model.get_current_precision()
            if bits >= 32:
                raise RuntimeError("GraphedCalibratedForward: set a quantised precision first (nothing to calibrate at 32 bits)")
            qs = self._input_quantizers(bits)
            if self.world > 1:
                # every rank must bring the same quantisers (checked once per capture, dp.sync_calibration_stats
                # checks it on every call)
                import torch.distributed as dist
                mine = torch.tensor([len(qs), sum(q.scale.numel() for q in qs)], dtype=torch.int64,
                                    device=next(self.model.parameters()).device)
                lo, hi = mine.clone(), mine.clone()
                dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.group)
                dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=self.group)
                if not torch.equal(lo, hi):
                    raise RuntimeError("GraphedCalibratedForward: ranks disagree on the input quantisers")
            dev = next(self.model.parameters()).device
            self.ids = torch.empty(ids.shape, dtype=ids.dtype, device=dev)      # `ids` may live in pinned host memory
            self.ids.copy_(ids)
            if self._nodata is None:
                self._nodata = torch.zeros((), dtype=torch.int64, device=dev)
            with torch.no_grad():
                for _ in range(2):                           # warm-up: operand caches, cuDNN plans, workspaces, shapes
                    self._eager(qs)
                torch.cuda.synchronize()
                self._nodata.zero_()
                pool = torch.cuda.graph_pool_handle()
                n0 = _lib.launch_count()
                g1 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g1, pool=pool):
                    live, mins, maxs, fl = self._stats_body(qs)
                g2 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g2, pool=pool):
                    self.out = self._forward_body(qs, live, mins, maxs, fl)
                self.kernels_per_replay = _lib.launch_count() - n0
            self.g1, self.g2, self._buckets = g1, g2, (mins, maxs, fl)
            self._keep = (self._helper._cached_operands(), live, self._temps)
            self._sig = self._signature(ids)
        self.ids.copy_(ids, non_blocking=True)
        self.g1.replay()
        if self.world > 1:
            self._exchange(*self._buckets)
        self.g2.replay()
        return self.out

    def nodata_count(self) -> int:
        return 0 if self._nodata is None else int(self._nodata.item())


# ----------------------------------------------------------------------------------------------------------------
# The switchable-precision training step (BASELINE.json configs[2]; p1/train_sp.py:341-397)
# ----------------------------------------------------------------------------------------------------------------

class FlatTrainState:
    """The path's trainable parameters -- LoRA A/B of every student width and every LayerNorm pair of the widths in
    use (north_star (d); README.md:108-112 of the reference) -- laid out in ONE flat float32 buffer, grouped in
    segments: [LayerNorm pairs of the teacher width | per student width b: LayerNorm pairs @b, LoRA A/B @b].

      * `p.data` of each parameter is a view into `flat_param`, `p.grad` a view into `flat_grad`: autograd
        accumulates straight into the buffer the NCCL all-reduce sends (no pack / scatter kernels, no `.grad`
        reallocation, static addresses for CUDA graphs);
      * AdamW state (`exp_avg`, `exp_avg_sq`) is flat too; a segment is updated only in optimizer steps in which it
        received a gradient, with its own step count -- what torch.optim.AdamW does for parameters whose .grad is
        None after `zero_grad(set_to_none=True)` (p1/train_sp.py:344).
    Everything else in the model is frozen (requires_grad=False)."""

    def __init__(self, model, teacher_bits, student_bits):
        self.model = model
        segs = {teacher_bits: []}
        for b in student_bits:
            segs[b] = []
        frozen = []
        for name, p in model.named_parameters():
            parts = name.split('.')
            seg = None
            if 'lora_adapters' in parts:
                b = int(parts[parts.index('lora_adapters') + 1][:-3])
                if b in student_bits and parts[-1] in ('lora_A', 'lora_B'):
                    seg = b
            elif len(parts) >= 2 and parts[-2] in ('weights', 'biases') and parts[-1].isdigit() and int(parts[-1]) in segs:
                seg = int(parts[-1])
            if seg is None:
                frozen.append(p)
            else:
                segs[seg].append((name, p))
        dev = next(model.parameters()).device
        total, self.segments, self.slots = 0, {}, {}
        for key, items in segs.items():
            start = total
            for name, p in items:
                self.slots[name] = (p, total, p.numel())
                total += (p.numel() + 3) // 4 * 4
            self.segments[key] = (start, total)
        self.numel = total
        self.flat_param = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self.steps = {k: 0 for k in self.segments}
        with torch.no_grad():
            for p in frozen:
                p.requires_grad_(False)
            for name, (p, off, n) in self.slots.items():
                view = self.flat_param[off:off + n].view_as(p)
                view.copy_(p.data)
                p.data = view
                p.requires_grad_(True)
                p.grad = self.flat_grad[off:off + n].view_as(p)
                p._spq_accumulate_in_place = True      # backward kernels add into p.grad themselves (lora._grad_sink)
        self.params = [p for p, _, _ in self.slots.values()]

    def seg_grad(self, key):
        a, b = self.segments[key]
        return self.flat_grad[a:b]

    def zero_grad(self):
        self.flat_grad.zero_()
        for p, off, n in self.slots.values():             # something replaced a .grad (set_to_none, a new tensor)
            g = p.grad
            if g is None or g.data_ptr() != self.flat_grad.data_ptr() + 4 * off:
                p.grad = self.flat_grad[off:off + n].view_as(p)

    def bump_versions(self):
        """The fused optimiser writes through raw pointers: tell torch (and the operand caches keyed on
        `_version`) that the parameters changed."""
        inc = torch.autograd.graph.increment_version
        for p in self.params:
            inc(p)


class SPTrainer:
    """One optimizer step of upstream's switchable-precision training (`train_step`, p1/train_sp.py:341-397) on
    this repo's modules, batch-sharded data parallel:

        zero_grad                                                                 :344
        micro-step 0      : teacher width, CE loss with labels = ids, backward;   :354-356, 320-323
                            then a no-grad forward whose logits + 13 hidden states are the distillation targets
                            (DistillationManager.update_teacher)                   :325-326, distillation_manager.py:34-62
        micro-steps 1..G-1: width = random.choice(student widths); LoRA quantisers recalibrated on their weights
                            (calibrate_lora_only); forward; KL(T) * T^2 + 1e-7 * MSE(one random hidden layer);
                            backward                                               :357-380, distillation_manager.py:64-116
        every loss is divided by G (gradient accumulation)                         :339
        clip_grad_norm_(1.0); AdamW step; cosine LR advanced once per micro-step   :380, 390-393

    B200 mapping: each micro-step (recalibration + forward + loss + backward) is ONE CUDA-graph replay per width;
    gradients accumulate into `FlatTrainState.flat_grad`; a segment's all-reduce (NCCL over NVLink, async) is
    launched right after the LAST micro-step that touches it, so it overlaps the remaining micro-steps and only
    the final width's segment is exposed; 1/world and the clip coefficient are folded into the fused AdamW kernel
    (`spq_grad_sumsq`, `spq_adamw_flat`).  Upstream's GradScaler is a no-op here (gradients are float32; its
    power-of-two scale cancels exactly) and is not reproduced; the teacher cache of DistillationManager is the
    static output of the teacher graph.  All ranks draw the same widths / feature layers from `rng`."""

    def __init__(self, model, bit_widths, *, grad_accum=8, lr=1e-4, weight_decay=0.01, betas=(0.9, 0.999), eps=1e-8,
                 max_grad_norm=1.0, temperature=3.0, alpha_kl=1.0, alpha_feature=1e-7, total_lr_steps=None,
                 group=None, rng=None, use_graphs=True, grad_side_stream=True):
        import random
        self.model = model
        self.teacher_bits = max(bit_widths)
        self.student_bits = [b for b in bit_widths if b != self.teacher_bits]
        self.G, self.lr0, self.wd, self.betas, self.eps = grad_accum, lr, weight_decay, betas, eps
        self.max_norm, self.T, self.alpha_kl, self.alpha_feature = max_grad_norm, temperature, alpha_kl, alpha_feature
        self.total_lr_steps = total_lr_steps
        self.group = group
        self.rng = rng if rng is not None else random.Random(0)
        self.use_graphs = use_graphs
        self.world = torch.distributed.get_world_size(group) if (torch.distributed.is_available()
                                                                  and torch.distributed.is_initialized()) else 1
        self.state = FlatTrainState(model, self.teacher_bits, self.student_bits)
        linears = [m for m in model.modules() if m.__class__.__name__ == 'SPLinearWithLoRA']
        self.refreshers = {b: LoRARefresher(linears, b) for b in self.student_bits}
        self.dev = self.state.flat_param.device
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=self.dev)
        self.loss_buf = torch.zeros(grad_accum, 2, dtype=torch.float32, device=self.dev)
        self.layer_sel = torch.zeros(1, dtype=torch.int32, device=self.dev)
        import os
        self.grad_stream = torch.cuda.Stream() if (grad_side_stream and os.environ.get('SPQ_GRAD_SIDE', '1') != '0') else None
        self.ids = None
        self.graphs, self.outs, self.finishers, self.sigs = {}, {}, {}, {}
        self._frozen = None
        self.pool = None
        self.micro_steps_done = 0
        self.n_hidden = None
        self.phase_events = None          # set to a list to record (name, start_event, end_event) per phase

    # ------------------------------------------------------------------------------------------ bodies
    def _teacher_body(self):
        m = self.model
        m.set_precision(self.teacher_bits)
        self._mark_inner('teacher_forward')
        hidden, _ = m.transformer(self.ids, output_hidden_states=True)
        loss = lm_head_loss(m, hidden, labels=self.ids) / self.G           # outputs['loss'] of model(ids, labels=ids)
        self._mark_inner('teacher_backward')
        loss.backward()
        self._mark_inner('teacher_cache_forward')
        with torch.no_grad():                                              # DistillationManager.update_teacher
            t = m(self.ids, output_hidden_states=True, return_dict=True)
        return {'loss': loss.detach().reshape(1), 'logits': t['logits'], 'hidden': t['hidden_states']}

    def _student_body(self, bits):
        from . import _lib
        m = self.model
        m.set_precision(bits)
        self._mark_inner('student_forward')
        hidden, hs = m.transformer(self.ids, output_hidden_states=True)
        tch = self.outs[self.teacher_bits]
        self._mark_inner('distillation_loss')
        kl = lm_head_loss(m, hidden, teacher_logits=tch['logits'], temperature=self.T)
        with torch.no_grad():
            # F.mse_loss(student_hidden[l], teacher_hidden[l]) for the ONE layer drawn for this micro-step: the index is
            # read on the device (self.layer_sel, set by train_step before the replay), so the captured graph serves
            # every draw without evaluating all 13 pairs
            ht = tch['hidden']
            n = min(len(hs), len(ht))
            feat = _lib.mse_select([h.contiguous() for h in hs[:n]], [h.contiguous() for h in ht[:n]], self.layer_sel)
        self._mark_inner('student_backward')
        # the LoRA weight-gradient GEMMs (dA, dB: ~11 % of the step's kernel time, nothing downstream reads them) run on a
        # side stream beside the dX chain: parallel branches of the captured graph, joined when the block exits
        with side_stream_grads(self.grad_stream):
            ((self.alpha_kl / self.G) * kl).backward()      # the feature term carries no gradient (detached copies)
        return {'kl': kl.detach().reshape(1), 'feat': feat}

    def _refresh(self, bits):
        """calibrate_lora_only(bits) (p1/train_sp.py:125-163, 362-364).  Upstream repeats it before every student
        micro-step; the LoRA weights only change at the optimizer step, so it runs before the FIRST micro-step of each
        width in an optimizer step -- the repeats would reproduce the same numbers bit for bit (cached, like q(W))."""
        self._mark_inner('lora_recalibration')
        r = self.refreshers[bits]
        self.finishers[bits] = r.replay() if self.use_graphs else r.body()

    def _body(self, bits):
        return self._teacher_body() if bits == self.teacher_bits else self._student_body(bits)

    # ------------------------------------------------------------------------------------------ graphs
    def _signature(self, bits):
        """What a captured micro-step hard-wires beyond the live-read trainable parameters: the frozen tensors
        (their fp16 operands are cached by address / version) and, for a student width, the calibration state."""
        if self._frozen is None:
            own = {id(p) for p in self.state.params}
            self._frozen = [p for p in self.model.parameters() if id(p) not in own]
        sig = (tuple(self.ids.shape), self.model.training, tuple((p.data_ptr(), p._version) for p in self._frozen))
        return sig if bits == self.teacher_bits else sig + (self.refreshers[bits]._signature(),)

    def _ensure(self, bits):
        if not self.use_graphs:
            return
        sig = self._signature(bits)
        if bits in self.graphs and self.sigs.get(bits) == sig:
            return
        if bits != self.teacher_bits:
            self._refresh(bits)            # (captures the refresher's own graph on first use; its outputs are static)
        # warm-up on the eager path: operand caches, cuDNN plans, workspaces
        for _ in range(2):
            self.outs[bits] = self._body(bits)
        torch.cuda.synchronize()
        if self.pool is None:
            self.pool = torch.cuda.graph_pool_handle()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, pool=self.pool):
            self.outs[bits] = self._body(bits)
        self.graphs[bits], self.sigs[bits] = g, self._signature(bits)
        if bits == self.teacher_bits:
            for b in self.student_bits:                    # the students read the teacher graph's static outputs
                self.graphs.pop(b, None)
        self.state.zero_grad()                                               # the warm-ups accumulated into it

    def _run(self, bits):
        if self.use_graphs:
            self.graphs[bits].replay()
        else:
            self.outs[bits] = self._body(bits)
        return self.outs[bits]

    # ------------------------------------------------------------------------------------------ LR schedule
    def lr_at(self, micro_steps):
        """CosineAnnealingLR(T_max = iterations * G, eta_min = 0) stepped once per micro-step (p1/train_sp.py:380, 456-457)."""
        import math
        if not self.total_lr_steps:
            return self.lr0
        return self.lr0 * (1.0 + math.cos(math.pi * micro_steps / self.total_lr_steps)) / 2.0

    # ------------------------------------------------------------------------------------------ the step
    def _mark(self, name):
        if self.phase_events is None:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.phase_events.append((name, e))
        return e

    def _mark_inner(self, name):
        # phases inside a micro-step can only be separated on the eager path (a replay is one launch)
        if not self.use_graphs and not torch.cuda.is_current_stream_capturing():
            self._mark(name)

    def train_step(self, ids):
        """ids: [B, T] int64 token ids of this rank's shard (CUDA, or pinned host memory -> copied asynchronously).
        Returns {'loss': total loss as upstream reports it (sum of the G micro-step losses, each / G), 'precisions': [...]}
        -- one device->host read per optimizer step."""
        from . import _lib
        import torch.distributed as dist
        st = self.state
        if self.ids is None or self.ids.shape != ids.shape:
            self.ids = torch.empty(ids.shape, dtype=torch.int64, device=self.dev)
            self.graphs.clear(); self.outs.clear()
        self.ids.copy_(ids, non_blocking=True)
        n_layers_p1 = len(self.model.transformer.h) + 1
        # upstream's draw order: per student micro-step, first the width (train_step), then the feature layer
        # (compute_distillation_loss)
        sched, layers = [self.teacher_bits], [None]
        for _ in range(1, self.G):
            sched.append(self.rng.choice(self.student_bits))
            layers.append(self.rng.choice(list(range(n_layers_p1))))
        last_use = {b: i for i, b in enumerate(sched)}
        for b in dict.fromkeys(sched):
            self._ensure(b)
        self._mark('zero_grad')
        st.zero_grad()
        works = []
        refreshed = set()
        for i, bits in enumerate(sched):
            if bits != self.teacher_bits and bits not in refreshed:
                self._mark(f'lora_recalibration_{bits}')
                if bits not in self.finishers:          # (_ensure already refreshed a width it just captured)
                    self._refresh(bits)
                refreshed.add(bits)
            self._mark(f'micro_step_{bits}')
            if bits != self.teacher_bits:
                self.layer_sel.fill_(layers[i])          # the feature layer of this micro-step (read by spq_mse_select)
            out = self._run(bits)
            if bits == self.teacher_bits:
                self.loss_buf[i, 0:1].copy_(out['loss'])
                self.loss_buf[i, 1].zero_()
            else:
                self.loss_buf[i, 0:1].copy_(out['kl'])
                self.loss_buf[i, 1:2].copy_(out['feat'])
            if self.world > 1 and last_use[bits] == i:
                # this segment is final: its all-reduce runs on NCCL's stream under the remaining micro-steps
                works.append(dist.all_reduce(st.seg_grad(bits), op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self._mark('allreduce_wait')
        for w in works:
            w.wait()
        self._mark('optimizer')
        self.micro_steps_done += self.G
        lr = self.lr_at(self.micro_steps_done)
        inv_world = 1.0 / self.world
        _lib.grad_sumsq(st.flat_grad, self.sumsq, scale=inv_world)
        for b in dict.fromkeys(sched):
            a, e = st.segments[b]
            if e == a:
                continue
            st.steps[b] += 1
            _lib.adamw_flat(st.flat_param[a:e], st.flat_grad[a:e], st.exp_avg[a:e], st.exp_avg_sq[a:e], lr, self.betas,
                            self.eps, self.wd, st.steps[b], grad_scale=inv_world, total_sumsq=self.sumsq,
                            max_norm=self.max_norm)
        st.bump_versions()
        self._mark('end')
        vals = self.loss_buf.tolist()                                         # the one device->host read of the step
        for b, fin in list(self.finishers.items()):
            fin(redo=False)
        self.finishers.clear()
        total = 0.0
        for i, bits in enumerate(sched):
            if bits == self.teacher_bits:
                total += vals[i][0]
            else:
                total += (self.alpha_kl * vals[i][0] + self.alpha_feature * vals[i][1]) / self.G
        return {'loss': total, 'precisions': sched, 'lr': lr}

    def phase_times_ms(self):
        """After a step run with `phase_events = []`: [(phase, ms)] between consecutive marks (synchronises)."""
        torch.cuda.synchronize()
        ev = self.phase_events or []
        return [(ev[i][0], ev[i][1].elapsed_time(ev[i + 1][1])) for i in range(len(ev) - 1)]
