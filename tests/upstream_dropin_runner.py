#!/usr/bin/env python
"""Run UNMODIFIED upstream callers on top of this repo's drop-in modules, in a fresh
process so the sys.modules aliasing stays contained.  Prints one JSON object.

    python tests/upstream_dropin_runner.py part5_tests     # part5_squad/tests/{test_model,test_training_step,
                                                           #   test_loss,test_distillation}.py, every test_* function
    python tests/upstream_dropin_runner.py calibration     # p1/train_sp.py CalibrationManager + train_step

The aliases make `part1_switchable_precision.{quantization,quantization_methods,lora,switchable_batchnorm}` and the
byte-identical `part5_squad.{quantization,lora,switchable_batchnorm}` resolve to llm_qat_on_gpt2_b200's modules;
everything else (models_squad, train_squad, train_sp, distillation managers, the tests) is upstream's own file from
baseline/_ref.  Upstream's tests build CPU tensors; `torch.set_default_device('cuda')` puts them on the GPU, the only
place the drop-in runs.
"""
import contextlib
import io
import json
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True


def alias_dropin(models_too=False):
    import importlib
    import llm_qat_on_gpt2_b200  # noqa: F401
    names = ["quantization", "quantization_methods", "lora", "switchable_batchnorm"] + (["models_sp"] if models_too else [])
    for name in names:
        mod = importlib.import_module(f"llm_qat_on_gpt2_b200.{name}")
        sys.modules[f"part1_switchable_precision.{name}"] = mod
        if name != "models_sp":
            sys.modules[f"part5_squad.{name}"] = mod
        sys.modules[name] = mod


def run_part5_tests():
    import torch
    from oracle import upstream as up
    from llm_qat_on_gpt2_b200 import _lib
    # the package __init__ of part1 / part5 import their training drivers (datasets etc.); the tests only need the
    # packages to exist as namespaces for `part5_squad.lora`-style imports, which the aliases provide
    import types
    for pkg in ("part1_switchable_precision", "part5_squad"):
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(up.REF_DIR, pkg)]
        sys.modules[pkg] = m
    alias_dropin()
    p5 = os.path.join(up.REF_DIR, "part5_squad")
    sys.path.insert(0, os.path.join(p5, "tests"))
    sys.path.insert(0, p5)
    torch.set_default_device("cuda")
    torch.manual_seed(0)
    results = {}
    n0 = _lib.launch_count()
    for modname in ("test_model", "test_training_step", "test_loss", "test_distillation"):
        mod = __import__(modname)
        for fn in sorted(n for n in dir(mod) if n.startswith("test_") and callable(getattr(mod, n))):
            buf = io.StringIO()
            try:
                with contextlib.redirect_stdout(buf):
                    getattr(mod, fn)()
                results[f"{modname}.{fn}"] = "passed"
            except Exception:
                results[f"{modname}.{fn}"] = "FAILED: " + traceback.format_exc()[-1500:]
    import models_squad
    lin_cls = models_squad.SPLinearWithLoRA
    return {"results": results, "kernel_launches": _lib.launch_count() - n0,
            "linear_class_module": lin_cls.__module__, "watchdog": _lib.debug_status(),
            "models_squad_file": models_squad.__file__}


def run_calibration():
    """Upstream CalibrationManager.calibrate_all_precisions + ensure_calibrated + two upstream train_step()s (AMP
    autocast + GradScaler, grad-accum 3, teacher at 32 then random student bits) on upstream's models_sp built on the
    drop-in linears / layer norms; then the calibrated parameters are compared with an all-upstream run on
    torch-CUDA from the same seed and batches."""
    import random
    import types
    import torch
    from oracle import upstream as up
    from llm_qat_on_gpt2_b200 import _lib

    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda")
    bit_widths = (4, 8, 32)
    g = torch.Generator().manual_seed(3)
    batches = [{"input_ids": torch.randint(0, 50257, (2, 64), generator=g)} for _ in range(4)]

    def build(models_mod):
        cfg = up.gpt2_config(n_layer=2, bit_widths=bit_widths)
        torch.manual_seed(0)
        model = models_mod.SPLMHeadModel(cfg).to(dev)
        with torch.no_grad():
            model.transformer.wte.weight.normal_(0, 0.02)
            model.transformer.wpe.weight.normal_(0, 0.01)
        return model

    # --- all-upstream run first (its modules are imported under their own names before any alias exists)
    with up.quiet():
        ref = build(up.p1("models_sp"))
        init_state = {k: v.detach().clone() for k, v in ref.state_dict().items()}
        ref_train = up.p1_bare("train_sp")
        mgr = ref_train.CalibrationManager(ref, batches, dev)
        mgr.calibrate_all_precisions([4, 8], num_batches=3)
    ref_state = {k: v.detach().clone() for k, v in ref.state_dict().items()}

    # --- drop-in run: forget upstream's hot modules, alias ours, re-import upstream's models_sp / train_sp on top
    for k in [k for k in sys.modules if k.startswith("part1_switchable_precision") or k in (
            "train_sp", "distillation_manager", "quantization", "quantization_methods", "lora", "switchable_batchnorm")]:
        del sys.modules[k]
    pkg = types.ModuleType("part1_switchable_precision")
    pkg.__path__ = [os.path.join(up.REF_DIR, "part1_switchable_precision")]
    sys.modules["part1_switchable_precision"] = pkg
    alias_dropin()
    import importlib
    models_mod = importlib.import_module("part1_switchable_precision.models_sp")          # upstream file, our linears
    train_mod = importlib.import_module("train_sp")
    assert models_mod.SPLinearWithLoRA.__module__ == "llm_qat_on_gpt2_b200.lora"
    n0 = _lib.launch_count()
    with up.quiet():
        ours = build(models_mod)
        ours.load_state_dict(init_state, strict=True)                 # identical parameters, whatever the init order
        mgr2 = train_mod.CalibrationManager(ours, batches, dev)
        mgr2.calibrate_all_precisions([4, 8], num_batches=3)
        for b in (4, 8):
            mgr2.ensure_calibrated(b)                                  # prints q.scale.mean() etc. (p1/train_sp.py:176-229)
    cmp = {}
    for k, v in ours.state_dict().items():
        if k.rsplit(".", 1)[-1] in ("scale", "zero_point", "running_min", "running_max") and "lora_adapters" not in k:
            r = ref_state[k]
            if tuple(r.shape) != tuple(v.shape):
                cmp[k] = {"shape": [list(v.shape), list(r.shape)]}
                continue
            d = (v.contiguous().view(torch.int32).long() - r.contiguous().view(torch.int32).long()).abs()
            kind = "minmax" if ".4bit." in k else "log"
            ent = cmp.setdefault(kind + ("_input" if "quantizers_input" in k else "_weight"), {"max_ulp": 0, "tensors": 0, "differ": 0})
            ent["max_ulp"] = max(ent["max_ulp"], int(d.max()))
            ent["tensors"] += 1
            ent["differ"] += int((d > 0).any())

    # --- upstream train_step through the drop-in (teacher fwd+bwd @32 + cache forward, students at random bits)
    cfgT = types.SimpleNamespace(gradient_accumulation_steps=3, max_grad_norm=1.0, distill_temperature=3.0,
                                 distill_alpha_kl=1.0, distill_alpha_feature=1e-7, cache_size=32, feature_layers=None,
                                 num_iterations=4)
    random.seed(0)
    ours.train()
    for n, p in ours.named_parameters():
        p.requires_grad_("lora_" in n or "ln_" in n)
    with up.quiet():
        dm = importlib.import_module("distillation_manager").DistillationManager(ours, 32, cfgT)
        opt = torch.optim.AdamW([p for p in ours.parameters() if p.requires_grad], lr=1e-3)
        sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=12)
        scaler = torch.amp.GradScaler("cuda")
        before = {n: p.detach().clone() for n, p in ours.named_parameters() if p.requires_grad}
        losses = []
        it = iter(batches)
        batch = None
        for iteration in range(2):
            total, batch = train_mod.train_step(ours, it, batches, opt, scaler, list(bit_widths), dm, cfgT, iteration,
                                                None, mgr2, sched, batch)
            losses.append(float(total))
    moved = sum(int(not torch.equal(before[n], p.detach())) for n, p in ours.named_parameters() if n in before)
    return {"calibration_vs_upstream_cuda": cmp, "train_step_losses": losses, "params_moved": moved,
            "params_trainable": len(before), "kernel_launches": _lib.launch_count() - n0, "watchdog": _lib.debug_status(),
            "cache": dm.get_cache_stats()}


if __name__ == "__main__":
    what = sys.argv[1]
    out = run_part5_tests() if what == "part5_tests" else run_calibration()
    print("RESULT " + json.dumps(out))
