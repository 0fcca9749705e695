"""ncu launch list of ONE CPT training step (GPT-2 medium, width 6): warm up / capture, then cudaProfilerStart/Stop around a step.
    ncu --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_cpt.csv python tools/cpt_profile.py"""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llm_qat_on_gpt2_b200.cpt import CPTModel, CalibrationManager, CPTTrainer

dev = torch.device("cuda")
widths = [6]
mc = types.SimpleNamespace(vocab_size=50257, n_positions=1024, n_embd=1024, n_layer=24, n_head=16, layer_norm_epsilon=1e-5, embd_pdrop=0.1,
                           bit_widths=widths + [32], shared_lora_rank=16, shared_lora_alpha=32,
                           quantizer_per_bit={**{b: "log" for b in widths}, 32: None}, gradient_bits=8, attention_dtype="fp16")
torch.manual_seed(0)
model = CPTModel({"model": mc, "training": types.SimpleNamespace(target_bits=6)}).to(dev)
with torch.no_grad():
    for m in model.modules():
        if m.__class__.__name__ == "LoRAAdapter" and m.lora_B is not None:
            m.lora_B.normal_(0, 0.02)
g = torch.Generator().manual_seed(1)
loader = [{"input_ids": torch.randint(0, 50257, (32, 256), generator=g)}]
mgr = CalibrationManager(model, loader, dev)
mgr.calibrate_gradient_quantizers()
mgr.ensure_calibrated(6, num_batches=1)
model.train()
tr = CPTTrainer(model, use_graphs=True)
ids = loader[0]["input_ids"].to(dev)
for _ in range(3):
    tr.train_step(ids, 6, read_loss=False)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.train_step(ids, 6, read_loss=False)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
