"""GPT-2 wrapper around the B200 fake-quant linear path -- drop-in for
part1_switchable_precision/models_sp.py (SPAttention :18, SPMLP :78, SPBlock :130, SPModel :173,
SPLMHeadModel :390).  Same class names (the reference's calibration helpers match modules by
`__class__.__name__ == 'SPLinearWithLoRA'`, :239), constructor arguments, `set_precision`
fan-out, `disable/enable_lora_for_calibration`, `verify_precision_consistency`, `forward`
signature / return conventions, `generate` and `load_pretrained_weights`.

Only the hot-path modules run on this repo's kernels (SPLinearWithLoRA, SwitchableLayerNorm and
the tied LM head, which is the same tcgen05 GEMM with unquantised operands).  Attention, GELU,
embeddings, the loss and sampling stay stock PyTorch, as SURVEY.md section 8 scopes them; the
causal attention goes through torch's fused SDPA instead of materialising [B,H,T,T] scores, and
the reference's per-4-blocks `torch.cuda.empty_cache()` (:327-328) is not reproduced.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.utils.checkpoint import checkpoint

from . import _lib
from .lora import SPLinearWithLoRA, _FpWeightCache, linear_fp
from .switchable_batchnorm import SwitchableLayerNorm


def _sp_linear(config, n_in, n_out, bit_widths):
    return SPLinearWithLoRA(n_in, n_out, bit_widths=bit_widths,
                            lora_rank_per_bit=config.lora_rank_per_bit,
                            lora_alpha_per_bit=config.lora_alpha_per_bit,
                            quantizer_per_bit=config.quantizer_per_bit,
                            per_channel=getattr(config, 'per_channel_quantization', True))


class SPAttention(nn.Module):
    def __init__(self, config, bit_widths):
        super().__init__()
        self.n_head = config.n_head
        self.n_embd = config.n_embd
        self.head_dim = self.n_embd // self.n_head
        self.bit_widths = bit_widths
        self.c_attn = _sp_linear(config, config.n_embd, 3 * config.n_embd, bit_widths)
        self.c_proj = _sp_linear(config, config.n_embd, config.n_embd, bit_widths)
        # kept for state_dict compatibility with the reference (:50); SDPA does the masking
        self.register_buffer("bias", torch.tril(torch.ones(config.n_positions, config.n_positions)))
        # 'fp32': exact-softmax fp32 attention; 'fp16': flash attention on fp16 q/k/v (what the
        # reference's AMP training loop effectively runs, p1/train_sp.py:319)
        self.attention_dtype = getattr(config, 'attention_dtype', 'fp32')

    def set_precision(self, bits) -> int:
        self.current_bit_width = bits
        self.c_attn.set_precision(bits)
        self.c_proj.set_precision(bits)
        return self.current_bit_width

    def forward(self, hidden_states, attention_mask=None, residual=None, pre_norm=None):
        # attention_mask is accepted and ignored, as in the reference (:58-76); `residual` (SPBlock) is added
        # to the projection output (inside c_proj's GEMM epilogue when autograd is off); `pre_norm` (SPBlock: ln_1) is
        # applied to hidden_states first, inside c_attn's activation-side kernel when autograd is off
        B, T, C = hidden_states.shape
        half = self.attention_dtype == 'fp16'
        qkv = self.c_attn(hidden_states, out_half=half, pre_norm=pre_norm)
        q, k, v = qkv.split(self.n_embd, dim=2)
        q = q.view(B, T, self.n_head, self.head_dim).transpose(1, 2)
        k = k.view(B, T, self.n_head, self.head_dim).transpose(1, 2)
        v = v.view(B, T, self.n_head, self.head_dim).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v, is_causal=True)     # torch picks cuDNN's sm_100 fused attention
        o = o.transpose(1, 2).contiguous().view(B, T, C)
        return self.c_proj(o, residual=residual)


class SPMLP(nn.Module):
    def __init__(self, config, bit_widths=None):
        super().__init__()
        if bit_widths is None:
            bit_widths = getattr(config, 'bit_widths', [6, 8, 16, 32])
        self.bit_widths = bit_widths
        for attr in ('lora_rank_per_bit', 'lora_alpha_per_bit'):
            if not hasattr(config, attr):
                raise AttributeError(
                    f"Config missing required switchable precision attributes: {attr}\n"
                    "Required: lora_rank_per_bit, lora_alpha_per_bit")
        self.c_fc = _sp_linear(config, config.n_embd, 4 * config.n_embd, bit_widths)
        self.c_proj = _sp_linear(config, 4 * config.n_embd, config.n_embd, bit_widths)
        self.act = nn.GELU()          # exact-erf GELU, as the reference (:114)
        # 'fp32' (default): the GELU output is float32, as upstream's fp32 path.  'fp16': under no_grad the c_fc epilogue
        # stores gelu(.) as float16 -- what upstream's AMP loop holds there (autocast runs F.linear and GELU in fp16,
        # p1/train_sp.py:319) -- and c_proj's statistics / quantise kernels read 2 bytes per element instead of 4
        self.activation_dtype = getattr(config, 'mlp_activation_dtype', 'fp32')

    def set_precision(self, bits) -> int:
        if bits not in self.bit_widths:
            raise ValueError(f"Bit width {bits} not in configured widths {self.bit_widths}")
        self.c_fc.set_precision(bits)
        self.c_proj.set_precision(bits)
        return bits

    def forward(self, hidden_states, residual=None, pre_norm=None):
        # exact-erf GELU: fused into c_fc's GEMM epilogue when autograd is off, a separate pass otherwise;
        # `residual` (SPBlock) is added to the projection output, `pre_norm` (SPBlock: ln_2) applied to the input first
        half = self.activation_dtype == 'fp16' and not torch.is_grad_enabled()
        return self.c_proj(self.c_fc(hidden_states, fuse_gelu=True, out_half=half, pre_norm=pre_norm), residual=residual)


class SPBlock(nn.Module):
    def __init__(self, config, bit_widths):
        super().__init__()
        self.ln_1 = SwitchableLayerNorm(config.n_embd, precision_levels=bit_widths, eps=config.layer_norm_epsilon)
        self.attn = SPAttention(config, bit_widths)
        self.ln_2 = SwitchableLayerNorm(config.n_embd, precision_levels=bit_widths, eps=config.layer_norm_epsilon)
        self.mlp = SPMLP(config, bit_widths)

    def set_precision(self, bits) -> int:
        self.ln_1.set_precision(bits)
        self.attn.set_precision(bits)
        self.ln_2.set_precision(bits)
        self.mlp.set_precision(bits)
        return bits

    def forward(self, hidden_states, attention_mask=None, use_checkpoint=False):
        if use_checkpoint:
            return checkpoint(self._forward, hidden_states, attention_mask)
        return self._forward(hidden_states, attention_mask)

    def _forward(self, hidden_states, attention_mask=None):
        # residual stream (reference :139-147): x + attn(ln_1(x)), then x + mlp(ln_2(x)); the adds ride in the
        # c_proj epilogues when autograd is off
        if torch.is_grad_enabled():
            hidden_states = self.attn(self.ln_1(hidden_states), attention_mask, residual=hidden_states)
            hidden_states = self.mlp(self.ln_2(hidden_states), residual=hidden_states)
            return hidden_states
        # no_grad: each LayerNorm has exactly one consumer (c_attn / c_fc), which normalises the rows inside its own
        # activation-side kernel -- the float32 LayerNorm output is never stored
        hidden_states = self.attn(hidden_states, attention_mask, residual=hidden_states, pre_norm=self.ln_1)
        hidden_states = self.mlp(hidden_states, residual=hidden_states, pre_norm=self.ln_2)
        return hidden_states


class SPModel(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.bit_widths = config.bit_widths
        self.current_bit_width = max(self.bit_widths)
        self.wte = nn.Embedding(config.vocab_size, config.n_embd)
        self.wpe = nn.Embedding(config.n_positions, config.n_embd)
        self.drop = nn.Dropout(config.embd_pdrop)
        self.h = nn.ModuleList([SPBlock(config, bit_widths=self.bit_widths) for _ in range(config.n_layer)])
        self.ln_f = SwitchableLayerNorm(config.n_embd, precision_levels=self.bit_widths,
                                        eps=config.layer_norm_epsilon)

    def _sp_linears(self):
        return [m for m in self.modules() if m.__class__.__name__ == 'SPLinearWithLoRA']

    def _layer_norms(self):
        for block in self.h:
            yield block.ln_1
            yield block.ln_2
        yield self.ln_f

    def unfreeze_weights(self, bits):
        # reference :197-222 -- at the teacher width every LN pair and every base linear trains
        if bits != 32:
            return
        for ln in self._layer_norms():
            if isinstance(ln, SwitchableLayerNorm):
                for p in ln.precision_levels:
                    ln.weights[str(p)].requires_grad = True
                    ln.biases[str(p)].requires_grad = True
        for block in self.h:
            for lin in (block.attn.c_attn, block.attn.c_proj, block.mlp.c_fc, block.mlp.c_proj):
                lin.linear.weight.requires_grad = True
                lin.linear.bias.requires_grad = True

    def set_precision(self, bits) -> int:
        if bits not in self.bit_widths:
            raise ValueError(f"Bit width {bits} not in configured widths {self.bit_widths}")
        self.current_bit_width = bits
        for block in self.h:
            block.set_precision(bits)
        self.ln_f.set_precision(bits)
        return self.current_bit_width

    def disable_lora_for_calibration(self):
        for module in self._sp_linears():
            module.calibration_mode = True

    def enable_lora_after_calibration(self):
        for module in self._sp_linears():
            module.calibration_mode = False

    def verify_precision_consistency(self) -> Tuple[bool, Dict]:
        want = self.current_bit_width
        details = {'expected': want, 'mismatches': [], 'components': {}}

        def note(name, got):
            details['components'][name] = got
            if got != want:
                details['mismatches'].append(f'{name}: {got} (expected {want})')

        for i, block in enumerate(self.h):
            if isinstance(block.ln_1, SwitchableLayerNorm):
                note(f'block_{i}_ln_1', block.ln_1.current_precision)
            if isinstance(block.ln_2, SwitchableLayerNorm):
                note(f'block_{i}_ln_2', block.ln_2.current_precision)
            if hasattr(block.attn, 'current_bit_width'):
                note(f'block_{i}_attn', block.attn.current_bit_width)
            for nm, lin in (('mlp_c_fc', block.mlp.c_fc), ('mlp_c_proj', block.mlp.c_proj)):
                if hasattr(lin, 'current_precision'):
                    note(f'block_{i}_{nm}', lin.current_precision)
        if isinstance(self.ln_f, SwitchableLayerNorm):
            note('ln_f', self.ln_f.current_precision)
        return len(details['mismatches']) == 0, details

    def get_current_precision(self):
        return self.current_bit_width

    def forward(self, input_ids=None, inputs_embeds=None, attention_mask=None, use_checkpoint=False,
                output_hidden_states=False):
        if inputs_embeds is not None:
            hidden_states = inputs_embeds
        else:
            if input_ids is None:
                raise ValueError("Either input_ids or inputs_embeds must be provided")
            T = input_ids.shape[1]
            position_ids = torch.arange(0, T, dtype=torch.long, device=input_ids.device).unsqueeze(0)
            hidden_states = self.drop(self.wte(input_ids) + self.wpe(position_ids))

        # upstream returns `hidden_states.clone().detach()` copies (:323); nothing on this path writes a hidden state
        # in place (every block output is a fresh tensor), so the detached tensors themselves are returned and the
        # 13 copies per forward are skipped
        all_hidden_states = [] if output_hidden_states else None
        for block in self.h:
            if output_hidden_states:
                all_hidden_states.append(hidden_states.detach())
            hidden_states = block(hidden_states, attention_mask, use_checkpoint)
        hidden_states = self.ln_f(hidden_states)
        if output_hidden_states:
            all_hidden_states.append(hidden_states.detach())
            return hidden_states, all_hidden_states
        return hidden_states

    def load_pretrained_weights(self, pretrained_model, device='cuda'):
        # reference :338-388 -- copy a HF GPT2Model in (Conv1D weights are transposed) and freeze
        def put(param, value):
            param.data = value
            param.requires_grad = False

        put(self.wte.weight, pretrained_model.wte.weight.data.clone())
        put(self.wpe.weight, pretrained_model.wpe.weight.data.clone())
        for i in range(min(len(self.h), len(pretrained_model.h))):
            src, dst = pretrained_model.h[i], self.h[i]
            for ln_dst, ln_src in ((dst.ln_1, src.ln_1), (dst.ln_2, src.ln_2)):
                for key in ln_dst.ln_layers:
                    view = ln_dst.ln_layers[key]
                    view.weight.data = ln_src.weight.data.clone()
                    view.bias.data = ln_src.bias.data.clone()
                    view.weight.requires_grad = False
                    view.bias.requires_grad = False
            for lin_dst, lin_src in ((dst.attn.c_attn, src.attn.c_attn), (dst.attn.c_proj, src.attn.c_proj),
                                     (dst.mlp.c_fc, src.mlp.c_fc), (dst.mlp.c_proj, src.mlp.c_proj)):
                put(lin_dst.linear.weight, lin_src.weight.data.t().contiguous())
                put(lin_dst.linear.bias, lin_src.bias.data.clone())
        for key in self.ln_f.ln_layers:
            view = self.ln_f.ln_layers[key]
            view.weight.data = pretrained_model.ln_f.weight.data.clone()
            view.bias.data = pretrained_model.ln_f.bias.data.clone()
            view.weight.requires_grad = False
            view.bias.requires_grad = False
        print("✅ Loaded pretrained weights with S-BN support")
        print("   - All precision-specific LayerNorm layers initialized")
        return self


class SPLMHeadModel(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.transformer = SPModel(config)
        self.lm_head = nn.Linear(config.n_embd, config.vocab_size, bias=False)
        self.lm_head.weight = self.transformer.wte.weight          # tied, unquantised (:396-398)
        self._lm_head_cache = _FpWeightCache()

    def set_precision(self, bits) -> int:
        return self.transformer.set_precision(bits)

    def disable_lora_for_calibration(self):
        self.transformer.disable_lora_for_calibration()

    def enable_lora_after_calibration(self):
        self.transformer.enable_lora_after_calibration()

    def verify_precision_consistency(self) -> Tuple[bool, Dict]:
        return self.transformer.verify_precision_consistency()

    def get_current_precision(self):
        return self.transformer.get_current_precision()

    def forward(self, input_ids=None, inputs_embeds=None, labels=None, attention_mask=None,
                use_checkpoint=False, output_hidden_states=False, return_dict=False):
        out = self.transformer(input_ids, inputs_embeds=inputs_embeds, attention_mask=attention_mask,
                               use_checkpoint=use_checkpoint, output_hidden_states=output_hidden_states)
        hidden_states, all_hidden_states = out if output_hidden_states else (out, None)

        # evaluation with labels: the LM-head GEMM also leaves the per-row log-sum-exp partials, so the loss
        # needs no second pass over the [B, T, V] logits
        fused_ce = labels is not None and not (torch.is_grad_enabled() and
                                               (hidden_states.requires_grad or self.lm_head.weight.requires_grad))
        lse_parts = [] if fused_ce else None
        logits = linear_fp(hidden_states, self.lm_head.weight, None, self._lm_head_cache, lse_out=lse_parts)

        loss = None
        if labels is not None:
            # next-token loss of the reference (:441-449) without its 2 x [B,T,V] shifted copies:
            # position t is scored against labels[t+1]; the last position is ignored, so the mean
            # runs over the same B*(T-1) terms
            targets = torch.full_like(labels, -100)
            targets[..., :-1] = labels[..., 1:]
            if fused_ce and lse_parts:
                loss = _lib.cross_entropy_from_parts(lse_parts[0], logits.view(-1, logits.size(-1)), targets)
            elif torch.is_grad_enabled() and logits.requires_grad:
                loss = F.cross_entropy(logits.reshape(-1, logits.size(-1)), targets.reshape(-1), ignore_index=-100)
            else:
                # evaluation: one fused pass over the (stride-padded) logits
                loss = _lib.cross_entropy_fwd(logits.view(-1, logits.size(-1)), targets)

        if return_dict or output_hidden_states:
            return {'loss': loss, 'logits': logits, 'hidden_states': all_hidden_states}
        return {'loss': loss, 'logits': logits} if loss is not None else logits

    def generate(self, input_ids, max_length=100, temperature=1.0, do_sample=True, top_k=50, top_p=0.95,
                 eos_token_id=None, attention_mask=None):
        # reference :460-507 -- full-prefix re-forward per token, top-k then nucleus filtering
        self.eval()
        with torch.no_grad():
            mask = attention_mask
            for _ in range(max_length - input_ids.shape[1]):
                outputs = self.forward(input_ids, attention_mask=mask)
                logits = outputs if not isinstance(outputs, dict) else outputs['logits']
                nxt = logits[:, -1, :] / temperature
                if do_sample:
                    if top_k > 0:
                        kth = torch.topk(nxt, top_k)[0][..., -1, None]
                        nxt[nxt < kth] = float('-inf')
                    if top_p < 1.0:
                        sorted_logits, sorted_idx = torch.sort(nxt, descending=True)
                        cum = torch.cumsum(F.softmax(sorted_logits, dim=-1), dim=-1)
                        drop = cum > top_p
                        drop[..., 1:] = drop[..., :-1].clone()
                        drop[..., 0] = 0
                        nxt[drop.scatter(1, sorted_idx, drop)] = float('-inf')
                    next_tokens = torch.multinomial(F.softmax(nxt, dim=-1), num_samples=1)
                else:
                    next_tokens = torch.argmax(nxt, dim=-1, keepdim=True)
                input_ids = torch.cat([input_ids, next_tokens], dim=1)
                if mask is not None:
                    mask = torch.cat([mask, torch.ones((mask.shape[0], 1), dtype=mask.dtype, device=mask.device)], dim=1)
                if eos_token_id is not None and (next_tokens == eos_token_id).all():
                    break
                if input_ids.shape[1] >= self.config.n_positions:
                    break
        return input_ids

    def load_pretrained_weights(self, pretrained_model, device='cuda'):
        self.transformer.load_pretrained_weights(pretrained_model.transformer, device)
        self.lm_head.weight = self.transformer.wte.weight
        print("LM head weights tied to token embeddings")
        return self
