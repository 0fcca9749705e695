"""fp32 restatement of upstream's `train_step` (p1/train_sp.py:341-397) assembled from UNMODIFIED upstream components
(baseline/_ref): the upstream SPLMHeadModel, `CalibrationManager.calibrate_lora_only`, `DistillationManager.update_teacher`
/ `compute_distillation_loss`, torch AdamW, `clip_grad_norm_`, CosineAnnealingLR.

TEST INFRASTRUCTURE ONLY.  Upstream's own `train_step` hard-codes `torch.amp.autocast('cuda')` + GradScaler
(p1/train_sp.py:319, 379, 390-393); a 1e-3 comparison needs the float32 numerics, so the loop is restated here line by
line without the autocast context and without the scaler (whose power-of-two factor cancels).  The draw order of
`random` (width per student micro-step, then the feature layer inside compute_distillation_loss) is upstream's.
"""
import random
import types

import torch

from . import upstream


def make_config(grad_accum=8, temperature=3.0, alpha_kl=1.0, alpha_feature=1e-7, max_grad_norm=1.0):
    return types.SimpleNamespace(gradient_accumulation_steps=grad_accum, max_grad_norm=max_grad_norm,
                                 distill_temperature=temperature, distill_alpha_kl=alpha_kl,
                                 distill_alpha_feature=alpha_feature, cache_size=32, feature_layers=None)


class UpstreamTrainStep:
    def __init__(self, model, bit_widths, config, lr=1e-4, weight_decay=0.01, betas=(0.9, 0.999), eps=1e-8,
                 total_lr_steps=None):
        self.model, self.cfg, self.bit_widths = model, config, list(bit_widths)
        train_sp = upstream.p1_bare("train_sp")
        dm = upstream.p1_bare("distillation_manager")
        self.calib = train_sp.CalibrationManager(model, [], next(model.parameters()).device)
        self.distill = dm.DistillationManager(model, max(bit_widths), config)
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.opt = torch.optim.AdamW(self.params, lr=lr, weight_decay=weight_decay, betas=betas, eps=eps)
        self.sched = (torch.optim.lr_scheduler.CosineAnnealingLR(self.opt, T_max=total_lr_steps)
                      if total_lr_steps else None)

    def step(self, input_ids, clip=True, apply=True):
        """One optimizer step; returns (total_loss, precisions, {name: accumulated grad before clipping})."""
        cfg, model = self.cfg, self.model
        self.opt.zero_grad(set_to_none=True)                                   # :344
        model.train()
        total, used = 0.0, []
        teacher = max(self.bit_widths)
        students = [b for b in self.bit_widths if b != teacher]
        with upstream.quiet():
            for bit_step in range(cfg.gradient_accumulation_steps):            # :353
                precision = teacher if bit_step == 0 else random.choice(students)
                used.append(precision)
                if precision < 32:                                             # :362-364
                    model.set_precision(precision)
                    self.calib.calibrate_lora_only(precision, num_batches=2)
                model.set_precision(precision)                                 # compute_loss_single_precision :317
                if precision == teacher:
                    out = model(input_ids, labels=input_ids, output_hidden_states=True, return_dict=True)
                    loss = out['loss']
                    with torch.no_grad():
                        self.distill.update_teacher(input_ids, None)
                else:
                    out = model(input_ids, output_hidden_states=True, return_dict=True)
                    loss = self.distill.compute_distillation_loss(out, input_ids)
                loss = loss / cfg.gradient_accumulation_steps                  # :339
                total += loss.detach().item()
                loss.backward()
                if self.sched is not None:
                    self.sched.step()                                          # :380
        grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
        if clip:
            torch.nn.utils.clip_grad_norm_(model.parameters(), cfg.max_grad_norm)   # :391
        if apply:
            self.opt.step()
        return total, used, grads
