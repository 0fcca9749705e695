"""Checkpoint / deploy formats next to the hot path (reference: part1_switchable_precision/
deploy.py).  Same file layouts and dictionary keys as upstream, so its evaluation loaders read what this writes:

  convert_to_int8(model)            per-tensor absmax/127 (or min/max/255) integer view of every SPLinearWithLoRA
                                    weight at the module's current precision       (reference :5-64)
  save_int8_checkpoint(...)         {'int8_state_dict', 'model_info', configs}      (reference :66-123)
  save_sp_checkpoints(...)          one FP32 state_dict .pth per non-32-bit width   (reference :125-183)
  load_model_for_evaluation(...)    rebuild an SPLMHeadModel from such a file       (reference :185-253)
  export_integer_weights(model, b)  NEW: the true integer codes and per-channel scales of the *calibrated* weight
                                    quantisers, straight from the CUDA quantise kernel (spq_fake_quantize)

File IO is not a hot path: the arithmetic of convert_to_int8 is a handful of torch reductions per layer, kept in
torch so that it also runs on a CPU copy of the model, exactly like upstream.
"""
import os
import time

import torch


def _is_sp_linear(module) -> bool:
    return 'LinearWithLoRA' in module.__class__.__name__ and hasattr(module, 'linear')


def _weight_quantizer(module):
    if hasattr(module, 'quantize_weight'):
        return module.quantize_weight
    if hasattr(module, 'quantizers_weight'):
        return module.quantizers_weight[f"{getattr(module, 'current_bits', 8)}bit"]
    return None


def convert_to_int8(model):
    model.eval()
    out = {}
    with torch.no_grad():
        for name, module in model.named_modules():
            if not _is_sp_linear(module):
                continue
            wq = _weight_quantizer(module)
            if wq is None:
                continue
            w = module.linear.weight.data
            if wq.symmetric:
                top = w.abs().max()
                scale = top / 127.0 if top > 0 else 1.0
                zero_point = torch.tensor(0, dtype=torch.int32)
                q = torch.round(w / scale).clamp(-128, 127).to(torch.int8)
            else:
                lo, hi = w.min(), w.max()
                scale = (hi - lo) / 255.0 if hi > lo else 1.0
                zero_point = torch.round(-lo / scale).clamp(0, 255).to(torch.int32)
                q = torch.round((w - lo) / scale).clamp(0, 255).to(torch.uint8)
            prefix = f"{name}." if name else ""
            out[prefix + "weight_int8"] = q.cpu()
            out[prefix + "scale"] = torch.scalar_tensor(scale, dtype=torch.float32).cpu()
            out[prefix + "zero_point"] = zero_point.cpu()
            if module.linear.bias is not None:
                out[prefix + "bias"] = module.linear.bias.data.cpu()
            # upstream looks for `.lora` / `.loras`; the switchable module keeps its adapters in `.lora_adapters`,
            # so (as upstream) no LoRA tensors are exported for it
            lora = getattr(module, 'lora', None)
            if lora is None and hasattr(module, 'loras'):
                lora = module.loras[f"{getattr(module, 'current_bits', 8)}bit"]
            if lora is not None:
                out[prefix + "lora.A"] = lora.lora_A.data.cpu()
                out[prefix + "lora.B"] = lora.lora_B.data.cpu()
                out[prefix + "lora.scaling"] = torch.scalar_tensor(lora.scaling, dtype=torch.float32).cpu()
    return out


def save_int8_checkpoint(model, filepath, model_config=None, training_config=None, target_bits=None):
    if target_bits is not None and hasattr(model, 'set_precision'):
        model.set_precision(target_bits)
    int8_sd = convert_to_int8(model)
    mb = 1024 * 1024
    fp32_params = sum(p.numel() for p in model.parameters())
    int8_params = sum(t.numel() for k, t in int8_sd.items() if 'int8' in k)
    meta_mb = sum(t.numel() * 4 / mb for k, t in int8_sd.items() if 'int8' not in k)
    total_mb = int8_params / mb + meta_mb
    ckpt = {
        'int8_state_dict': int8_sd,
        'model_info': {
            'fp32_params': fp32_params, 'fp32_size_mb': fp32_params * 4 / mb, 'int8_params': int8_params,
            'int8_size_mb': int8_params / mb, 'metadata_size_mb': meta_mb, 'total_size_mb': total_mb,
            'compression_ratio': (fp32_params * 4 / mb) / total_mb if total_mb > 0 else 0, 'target_bits': target_bits,
        },
    }
    if model_config:
        ckpt['model_config'] = model_config.__dict__
        ckpt['bit_widths'] = getattr(model_config, 'bit_widths', None)
    if training_config:
        ckpt['training_config'] = training_config.__dict__
    torch.save(ckpt, filepath)
    return ckpt


def save_sp_checkpoints(model, base_filename, model_config, training_config=None):
    """One `<base>_<bits>bit_FP32_<timestamp>.pth` per configured width below 32; returns {bits: path}."""
    widths = getattr(model_config, 'bit_widths', [6, 8, 16, 32])
    stamp = time.strftime('%Y%m%d_%H%M%S')
    saved = {}
    for bits in widths:
        if bits == 32:
            continue
        model.set_precision(bits)
        path = f"{base_filename}_{bits}bit_FP32_{stamp}.pth"
        ckpt = {'model_state_dict': model.state_dict(), 'model_config': model_config.__dict__,
                'training_config': training_config.__dict__ if training_config else None, 'bit_width': bits,
                'timestamp': stamp}
        for attempt in range(3):
            try:
                torch.save(ckpt, path, pickle_protocol=4)
                back = torch.load(path, map_location='cpu', weights_only=False)
                if back['bit_width'] != bits:
                    raise RuntimeError("checkpoint read back with a different bit width")
                saved[bits] = path
                break
            except Exception:
                if os.path.exists(path):
                    try:
                        os.remove(path)
                    except OSError:
                        pass
                if attempt == 2:
                    break
                time.sleep(1.0)
    return saved


def load_model_for_evaluation(checkpoint_path, config=None, target_bits=None, device='cuda'):
    """Rebuild an SPLMHeadModel (per-tensor quantisation, as upstream's evaluation path) and load the file
    strictly.  `config` needs the ModelConfig attributes (vocab_size ... quantizer_per_bit); when omitted they
    are taken from the checkpoint's 'model_config'."""
    from types import SimpleNamespace
    from transformers import GPT2Config
    from .models_sp import SPLMHeadModel
    device = torch.device(device)
    ckpt = torch.load(checkpoint_path, map_location='cpu', weights_only=False)
    if config is None:
        if 'model_config' not in ckpt:
            raise ValueError("No config provided and checkpoint doesn't contain model_config")
        config = SimpleNamespace(**ckpt['model_config'])
    config.per_channel_quantization = False
    if target_bits is None:
        target_bits = ckpt.get('bit_width', ckpt.get('metadata', {}).get('target_bits', 16))
    g = GPT2Config(vocab_size=config.vocab_size, n_positions=config.n_positions, n_embd=config.n_embd,
                   n_layer=config.n_layer, n_head=config.n_head, layer_norm_epsilon=config.layer_norm_epsilon,
                   use_cache=False, bos_token_id=50256, eos_token_id=50256)
    g.lora_rank_per_bit = config.lora_rank_per_bit
    g.lora_alpha_per_bit = config.lora_alpha_per_bit
    g.quantizer_per_bit = config.quantizer_per_bit
    g.bit_widths = config.bit_widths
    g.per_channel_quantization = False
    model = SPLMHeadModel(g)
    model.load_state_dict(ckpt['model_state_dict'] if 'model_state_dict' in ckpt else ckpt)
    model.set_precision(target_bits)
    return model.to(device).eval()


def pack_int4(codes: torch.Tensor) -> torch.Tensor:
    """[..., K] integer codes in [-8, 7] (or [0, 15]) -> uint8 [..., ceil(K / 2)]: element 2i in the low nibble, 2i+1 in
    the high nibble, two's complement; an odd K is padded with a zero nibble."""
    c = codes.to(torch.int16)
    if c.shape[-1] % 2:
        c = torch.nn.functional.pad(c, (0, 1))
    lo, hi = c[..., 0::2] & 0xF, c[..., 1::2] & 0xF
    return (lo | (hi << 4)).to(torch.uint8)


def unpack_int4(packed: torch.Tensor, K: int, signed: bool = True) -> torch.Tensor:
    """Inverse of pack_int4: uint8 [..., ceil(K / 2)] -> int8 [..., K]."""
    p = packed.to(torch.int16)
    both = torch.stack((p & 0xF, (p >> 4) & 0xF), dim=-1).reshape(*packed.shape[:-1], -1)[..., :K]
    if signed:
        both = torch.where(both >= 8, both - 16, both)
    return both.to(torch.int8)


def export_integer_weights(model, bits, packed: bool = False):
    """True integer weights of every SPLinearWithLoRA at `bits`: the codes the calibrated weight quantiser assigns
    (int8 when they fit, else int32; log quantisers: level index + int8 sign) with its scale / zero-point --
    `codes` from the same kernel the forward uses, so `dequant == (codes - zero_point) * scale` exactly for
    min-max quantisers.  `packed=True`: widths <= 4 store two codes per byte (`codes_packed`, see pack_int4;
    `codes_shape` keeps [N, K]) -- the packed int4 container upstream's `convert_to_int8` (p1/deploy.py:5-30) lacks.
    CUDA only."""
    from .quantization_methods import quantize_codes
    out = {}
    key = f"{bits}bit"
    with torch.no_grad():
        for name, m in model.named_modules():
            if m.__class__.__name__ != 'SPLinearWithLoRA' or key not in m.quantizers_weight:
                continue
            q = m.quantizers_weight[key]
            if not q.calibrated:
                raise RuntimeError(f"{name}: weight quantiser for {bits} bits is not calibrated")
            dq, codes, sign = quantize_codes(m.linear.weight.data, q.scale, q.zero_point, q.num_bits, q.symmetric,
                                             q.quantizer_type)
            small = int(codes.abs().max()) <= 127
            prefix = f"{name}." if name else ""
            lo, hi = int(codes.min()), int(codes.max())
            if packed and q.num_bits <= 4 and ((lo >= -8 and hi <= 7) or (lo >= 0 and hi <= 15)):
                out[prefix + "codes_packed"] = pack_int4(codes).cpu()
                out[prefix + "codes_shape"] = tuple(codes.shape)
                out[prefix + "codes_signed"] = lo < 0 or hi <= 7
            else:
                out[prefix + "codes"] = codes.to(torch.int8 if small else torch.int32).cpu()
            if sign is not None:
                out[prefix + "sign"] = sign.cpu()
            out[prefix + "scale"] = q.scale.detach().cpu()
            out[prefix + "zero_point"] = q.zero_point.detach().cpu()
            out[prefix + "quantizer_type"] = q.quantizer_type
            if m.linear.bias is not None:
                out[prefix + "bias"] = m.linear.bias.data.cpu()
    return out
