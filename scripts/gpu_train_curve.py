"""Sanity aid: 40 DOSE-PYFER training steps at 64^3 over four rotating synthetic samples; prints the loss curve."""
import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch
from conftest import load_manifest
from dose_prediction_b200 import networks, synth
from dose_prediction_b200.training import DoseTrainer
from oracle import synth_ckpt
size=64
man=[(k,([1,(size//16)**3,s[2]] if k.endswith("position_embeddings") else s)) for k,s,*_ in load_manifest("dose_pyfer")]
sd=synth_ckpt.make_state_dict(man, seed=0)
m=networks.Model(9,1,[-1,16,32,64,128,256],img_size=(size,)*3); m.load_state_dict(sd); m.cuda().train()
tr=DoseTrainer(m,2,size,lr=3e-4)
losses=[]
for step in range(40):
    vol=synth.make_batch(2,size,seed=100+2*(step%4))
    losses.append(float(tr.step(vol["dose_input"].cuda(), vol["gt"].cuda())))
print(["%.3f"%l for l in losses])
print("found_inf", int(tr.found_inf))
