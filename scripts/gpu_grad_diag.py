"""Diagnostic: where does the gradient difference CUDA-vs-oracle come from? (forward alignment under EMU, growth with depth)"""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_manifest
from dose_prediction_b200 import networks, synth
from dose_prediction_b200.training import DoseTrainer
from oracle import precision_probe, synth_ckpt, torch_ref

S, B = 32, 2
tokens = (S // 16) ** 3
man = [(k, ([1, tokens, s[2]] if k.endswith("position_embeddings") else s)) for k, s, *_ in load_manifest("dose_pyfer")]
sd = synth_ckpt.make_state_dict(man, seed=0)
model = networks.Model(9, 1, [-1, 16, 32, 64, 128, 256], img_size=(S,) * 3)
model.load_state_dict(sd, strict=True)
model = model.to("cuda:0").train()
vol = synth.make_batch(B, S, seed=1234)
gen = torch.Generator().manual_seed(5)
probe = [torch.randn(B, 1, S >> i, S >> i, S >> i, generator=gen) / (S >> i) ** 1.5 for i in range(4)]
torch.set_num_threads(os.cpu_count() or 1)
res = {}
for tag, emu in (("plain", None), ("emu", precision_probe.recipe_dose)):
    torch_ref.EMU = emu
    _, g, _, outs = torch_ref.dose_pyfer_train_step(sd, vol["dose_input"], vol["gt"], probe=probe)
    torch_ref.EMU = None
    res[tag] = (g, outs)
tr = DoseTrainer(model, B, S, probe=probe)
tr.forward_backward(vol["dose_input"].to("cuda:0"), vol["gt"].to("cuda:0"))
torch.cuda.synchronize()
mine = tr.grads(); outs = tr.outputs()
rel = torch_ref.rel_l2
for tag in ("plain", "emu"):
    g, o = res[tag]
    print(tag, "forward rel per head", [rel(a.cpu(), b) for a, b in zip(outs[1], o[1])])
    rows = [(n, rel(mine[n].cpu(), r)) for n, r in g.items() if float(r.norm()) > 1e-4 * max(float(v.norm()) for v in g.values())]
    for n, r in rows:
        if any(t in n for t in ("dose_convertors", "decoder1.conv_block.cov_.conv.0.weight", "decoder1.conv_block.cov_.conv_7.0.conv.3.weight",
                                "decoder1.conv_block.cov_.conv_7.0.conv.0.weight", "decoder1.transp_conv", "decoder2.conv_block.cov_.conv.0.weight",
                                "decoder2.conv_block.cov_.conv_7.0.conv.0.weight", "decoder3.conv_block.cov_.conv_7.0.conv.0.weight",
                                "decoder4.conv_block.cov_.conv_7.0.conv.0.weight", "decoder4.transp_conv", "skip1.layer.conv2", "skip1.layer.conv1",
                                "blocks.7.mlp.linear2.weight", "blocks.0.mlp.linear1.weight", "patch_embeddings.1.weight")):
            print("  ", tag, f"{r:.3e}", n)
print("oracle plain vs emu grads:", sorted(((rel(res['emu'][0][n], res['plain'][0][n]), n) for n in res['plain'][0] if float(res['plain'][0][n].norm()) > 0), reverse=True)[:3])
