"""Importer for the UNMODIFIED upstream packages installed under baseline/_ref/ (tools/install_ref.py).

TEST INFRASTRUCTURE ONLY: used by tests/, bench.py's reference / cpu_baseline legs and nothing else.
The product package (llm_qat_on_gpt2_b200/) never imports this module.

part1 (and part5) import each other both as `part1_switchable_precision.x` and, in the drivers, as bare
`distillation_manager` etc. (p1/train_sp.py:13-16); part2 uses bare names only (`from quantization import
...`, p2/cpt_model.py:9) that clash with part1's, so part2 is loaded under private module names.
"""
from __future__ import annotations

import contextlib
import importlib
import importlib.util
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.environ.get("SPQ_UPSTREAM_DIR", os.path.join(ROOT, "baseline", "_ref"))


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "part1_switchable_precision", "lora.py"))


def manifest() -> dict:
    with open(os.path.join(REF_DIR, "MANIFEST.json")) as fh:
        return json.load(fh)


def _ensure_path():
    if not available():
        raise RuntimeError(f"upstream copy not found under {REF_DIR}: run `python tools/install_ref.py` in the "
                           "build container (it ships to the GPU box with the gpurun snapshot)")
    sys.dont_write_bytecode = True
    for p in (os.path.join(REF_DIR, "part1_switchable_precision"), REF_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)


@contextlib.contextmanager
def quiet():
    """The upstream classes print on calibration / precision switches."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def p1(name: str):
    """part1_switchable_precision.<name> of the upstream copy (e.g. 'lora', 'models_sp', 'train_sp')."""
    _ensure_path()
    return importlib.import_module(f"part1_switchable_precision.{name}")


def p1_bare(name: str):
    """Upstream driver modules that are written for bare imports (train_sp, distillation_manager)."""
    _ensure_path()
    return importlib.import_module(name)


_p2_cache = {}


def p2(name: str):
    """part2_cyclic_precision_training/<name>.py under a private module name.  Its bare imports
    (`quantization`, `quantization_methods`, `cpt_model`) are resolved against part2 while it loads and the
    global names are restored afterwards, so part1's modules of the same names stay untouched."""
    _ensure_path()
    if name in _p2_cache:
        return _p2_cache[name]
    base = os.path.join(REF_DIR, "part2_cyclic_precision_training")
    clash = ("quantization", "quantization_methods", "cpt_model", "cyclic_scheduler", "calibration", "config_cpt")
    saved = {k: sys.modules.pop(k) for k in clash if k in sys.modules}
    sys.path.insert(0, base)
    try:
        for k in clash:                                  # give part2's own modules the bare names while loading
            if f"_upstream_p2_{k}" in sys.modules:
                sys.modules[k] = sys.modules[f"_upstream_p2_{k}"]
        mod = importlib.import_module(name)
        for k in clash:
            m = sys.modules.get(k)
            if m is not None and getattr(m, "__file__", "").startswith(base):
                sys.modules[f"_upstream_p2_{k}"] = m
                _p2_cache[k] = m
    finally:
        sys.path.remove(base)
        for k in clash:
            sys.modules.pop(k, None)
        sys.modules.update(saved)
    _p2_cache[name] = mod
    return mod


def gpt2_config(n_layer=12, n_embd=768, n_head=12, bit_widths=(4, 8, 32), quantizer_per_bit=None, rank=64, alpha=64,
                per_channel=True, embd_pdrop=0.0, vocab_size=50257, n_positions=1024):
    """GPT2Config with the extra attributes upstream monkey-patches on (p1/main_sp.py:26-46)."""
    from transformers import GPT2Config
    cfg = GPT2Config(vocab_size=vocab_size, n_positions=n_positions, n_embd=n_embd, n_layer=n_layer, n_head=n_head,
                     layer_norm_epsilon=1e-5, embd_pdrop=embd_pdrop)
    bw = list(bit_widths)
    cfg.bit_widths = bw
    cfg.lora_rank_per_bit = {b: (rank if b < 32 else 0) for b in bw}
    cfg.lora_alpha_per_bit = {b: (alpha if b < 32 else 0) for b in bw}
    qpb = quantizer_per_bit or {b: ("minmax" if b <= 4 else "log") for b in bw if b < 32}     # p1/config_sp.py:14-30
    qpb = dict(qpb)
    qpb.setdefault(32, None)
    cfg.quantizer_per_bit = qpb
    cfg.per_channel_quantization = per_channel
    return cfg
