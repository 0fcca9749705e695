"""Cyclic-precision training step (BASELINE.json configs[3]; p2/main_cpt.py:30-60 `train_epoch_with_cpt`, one batch):

    model.set_precision(bits); optimizer.zero_grad(); loss = model(ids, labels).loss; loss.backward()
    clip_grad_norm_(1.0); optimizer.step(); lr_scheduler.step()

with `bits` moving along `CyclicPrecisionScheduler` -- per STEP here (configs[3]), per epoch upstream.  The shared LoRA
adapter changes at every optimizer step and the width at every step, so the LoRA-level GEMM operands of every
CPTLinear (dequantised A / B at the step's width, their power-of-two normalisers, four fp16 operand builds: ~15 small
launches per linear, 97 linears in GPT-2 medium) are rebuilt each step.  `CPTTrainer` captures one CUDA graph per
width -- operand rebuild + forward + loss + backward -- so a step is one replay plus three optimiser launches; the
trainable parameters (shared LoRA A / B, LayerNorm weights and biases: p2/main_cpt.py:96-150) live in one flat buffer
whose gradient twin is what the data-parallel all-reduce sends (`spq_grad_sumsq`, `spq_adamw_flat`)."""
from __future__ import annotations

import math

import torch

from .. import _lib
from ..lora import side_stream_grads


class CPTTrainer:
    def __init__(self, model, *, lr=1e-4, weight_decay=0.01, betas=(0.9, 0.999), eps=1e-8, max_grad_norm=1.0,
                 total_lr_steps=None, group=None, use_graphs=True):
        self.model = model
        self.lr0, self.wd, self.betas, self.eps, self.max_norm = lr, weight_decay, betas, eps, max_grad_norm
        self.total_lr_steps = total_lr_steps
        self.group = group
        self.use_graphs = use_graphs
        self.world = torch.distributed.get_world_size(group) if (torch.distributed.is_available()
                                                                  and torch.distributed.is_initialized()) else 1
        named = []
        for n, p in model.named_parameters():
            parts = n.split('.')
            train = ('shared_lora' in parts and parts[-1] in ('lora_A', 'lora_B')) or \
                    (len(parts) >= 2 and parts[-2] in ('ln_1', 'ln_2', 'ln_f'))
            if train:
                named.append((n, p))
            else:
                p.requires_grad_(False)
        dev = next(model.parameters()).device
        self.slots, total = {}, 0
        for n, p in named:
            self.slots[n] = (p, total, p.numel())
            total += (p.numel() + 3) // 4 * 4
        self.numel = total
        self.flat_param = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for n, (p, off, cnt) in self.slots.items():
                view = self.flat_param[off:off + cnt].view_as(p)
                view.copy_(p.data)
                p.data = view
                p.requires_grad_(True)
                p.grad = self.flat_grad[off:off + cnt].view_as(p)
                p._spq_accumulate_in_place = True      # the gradient GEMM's fold pass adds into p.grad (lora._grad_sink)
        self.params = [p for p, _, _ in self.slots.values()]
        self.linears = [m for m in model.modules() if m.__class__.__name__ == 'CPTLinear']
        self.dev = dev
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self.ids = None
        self.graphs, self.outs, self.sigs = {}, {}, {}
        self.pool = None
        self.steps_done = 0
        self._frozen = None
        import os
        self.grad_stream = torch.cuda.Stream() if os.environ.get('SPQ_GRAD_SIDE', '1') != '0' else None

    # ------------------------------------------------------------------------------------------
    def _body(self, bits):
        m = self.model
        m.set_precision(bits)
        out = m(self.ids, labels=self.ids)
        with side_stream_grads(self.grad_stream):       # LoRA weight-gradient GEMMs beside the dX chain (lora._GradSide)
            out.loss.backward()
        return out.loss.detach().reshape(1)

    def _signature(self, bits):
        if self._frozen is None:
            own = {id(p) for p in self.params}
            self._frozen = [p for p in self.model.parameters() if id(p) not in own]
        gens = tuple((l.quantizer_weight.generation, l.quantizer_input.generation,
                      l.lora_weight_quantizers[f'{bits}bit'].generation if f'{bits}bit' in l.lora_weight_quantizers else 0,
                      l.shared_lora.grad_quantizer_A.generation if l.shared_lora.grad_quantizer_A is not None else 0,
                      l.shared_lora.grad_quantizer_B.generation if l.shared_lora.grad_quantizer_B is not None else 0)
                     for l in self.linears)
        return (tuple(self.ids.shape), self.model.training, tuple((p.data_ptr(), p._version) for p in self._frozen), gens)

    def _invalidate_lora_levels(self, bits):
        for l in self.linears:
            ent = l._op_cache.get(bits)
            if ent is not None:
                ent['lora'] = None
                ent['bwd'] = None

    def _ensure(self, bits):
        if not self.use_graphs:
            return
        sig = self._signature(bits)
        if bits in self.graphs and self.sigs.get(bits) == sig:
            return
        for _ in range(2):                                   # warm-up: base-level operand caches, cuDNN plans, workspaces
            self._body(bits)
        torch.cuda.synchronize()
        self._invalidate_lora_levels(bits)                   # rebuilt inside the capture -> rewritten by every replay
        if self.pool is None:
            self.pool = torch.cuda.graph_pool_handle()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, pool=self.pool):
            self.outs[bits] = self._body(bits)
        self.graphs[bits], self.sigs[bits] = g, self._signature(bits)

    def lr_at(self, step):
        if not self.total_lr_steps:
            return self.lr0
        return self.lr0 * (1.0 + math.cos(math.pi * step / self.total_lr_steps)) / 2.0

    def zero_grad(self):
        self.flat_grad.zero_()
        for p, off, cnt in self.slots.values():
            g = p.grad
            if g is None or g.data_ptr() != self.flat_grad.data_ptr() + 4 * off:
                p.grad = self.flat_grad[off:off + cnt].view_as(p)

    def train_step(self, ids, bits, read_loss=True):
        """One optimizer step at width `bits` on this rank's shard `ids` [B, T] (CUDA or pinned host)."""
        import torch.distributed as dist
        if self.ids is None or self.ids.shape != ids.shape:
            self.ids = torch.empty(ids.shape, dtype=torch.int64, device=self.dev)
            self.graphs.clear(); self.outs.clear()
        self.ids.copy_(ids, non_blocking=True)
        self._ensure(bits)
        self.zero_grad()
        if self.use_graphs:
            self.graphs[bits].replay()
            loss = self.outs[bits]
        else:
            loss = self._body(bits)
        if self.world > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group)
        inv_world = 1.0 / self.world
        lr = self.lr_at(self.steps_done)                     # optimizer.step() precedes lr_scheduler.step() upstream
        self.steps_done += 1
        _lib.grad_sumsq(self.flat_grad, self.sumsq, scale=inv_world)
        _lib.adamw_flat(self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq, lr, self.betas, self.eps, self.wd,
                        self.steps_done, grad_scale=inv_world, total_sumsq=self.sumsq, max_norm=self.max_norm)
        inc = torch.autograd.graph.increment_version
        for p in self.params:
            inc(p)
        return {'loss': float(loss.item()) if read_loss else loss, 'bits': bits, 'lr': lr}
