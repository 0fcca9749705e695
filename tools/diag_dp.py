"""Diagnostic: gradient of one SPTrainer step on a global batch vs the average over its two shards (one process)."""
import sys, os, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import upstream as up
from llm_qat_on_gpt2_b200 import SPLMHeadModel
from llm_qat_on_gpt2_b200.training import SPTrainer
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import test_gpu_training as T

att = sys.argv[1] if len(sys.argv) > 1 else "fp16"
G = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = up.gpt2_config(n_layer=4, bit_widths=(4, 8, 32), embd_pdrop=0.0)
cfg.attention_dtype = att
torch.manual_seed(0)
with up.quiet():
    model = SPLMHeadModel(cfg).cuda()
with torch.no_grad():
    model.transformer.wte.weight.normal_(0, 0.02); model.transformer.wpe.weight.normal_(0, 0.01)
    for n, p in model.named_parameters():
        if n.endswith("lora_B"):
            p.normal_(0, 0.02)
g = torch.Generator().manual_seed(1)
ids = torch.randint(0, 50257, (8, 256), generator=g).cuda()
T._calibrate(model, (4, 8), [ids])
model.train()

def grads(batch):
    tr = SPTrainer(model, [4, 8, 32], grad_accum=G, lr=0.0, weight_decay=0.0, rng=random.Random(3), use_graphs=False)
    out = tr.train_step(batch)
    g = tr.state.flat_grad.clone()
    segs = dict(tr.state.segments)
    return g, segs, out

def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))

g_all, segs, o_all = grads(ids)
g_all2, _, _ = grads(ids)
g0, _, o0 = grads(ids[:4]); g1, _, o1 = grads(ids[4:])
g_avg = (g0 + g1) / 2
print("attention", att, "G", G, "sched", o_all["precisions"], "losses", o_all["loss"], (o0["loss"] + o1["loss"]) / 2)
print("repeat same batch:", rel(g_all2, g_all))
for k, (a, b) in segs.items():
    print(f"segment {k}: shard-average vs global {rel(g_avg[a:b], g_all[a:b]):.3e}   |g| {float(g_all[a:b].norm()):.3e}")
