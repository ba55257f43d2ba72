"""TEST INFRASTRUCTURE — generate tests/golden/* from the REAL reference modules.

Run in the build container only (needs /root/reference):  python -m oracle.make_golden
Weights: oracle.synth_ckpt (seeded, per key).  Inputs: dose_prediction_b200.synth (seeded).
Outputs are what the reference's own nn.Modules (imported unmodified through ref_loader) return.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dose_prediction_b200 import synth  # noqa: E402
from oracle import ref_loader, synth_ckpt, torch_ref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
DOSE_SEED, SEG_SEED = 0, 1


def _np(t):
    return t.detach().cpu().numpy().astype(np.float32)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    # ---- drop-in contract: state_dict manifests at the headline size (128^3)
    for name, mod in (("dose_pyfer", ref_loader.build_dose(128)), ("oar_transeg", ref_loader.build_seg(128)),
                      ("oar_transeg_2ch", ref_loader.build_seg(128, in_channels=2))):
        with open(os.path.join(OUT, f"manifest_{name}.json"), "w") as f:
            json.dump(synth_ckpt.manifest_of(mod), f, indent=0)
    old = ref_loader.oar_transeg_old()
    old_model = old.TRANSEG(in_channels=1, out_channels=8, img_size=(96, 96, 96), feature_size=16,
                            hidden_size=768, mlp_dim=3072, num_heads=12, pos_embed="perceptron",
                            norm_name="instance", res_block=True, conv_block=True, dropout_rate=0.0)
    with open(os.path.join(OUT, "manifest_transeg_old_96.json"), "w") as f:
        json.dump(synth_ckpt.manifest_of(old_model), f, indent=0)

    # ---- 32^3 end-to-end vectors (batch 2 so per-instance statistics are exercised)
    vol = synth.make_batch(2, 32, seed=1234)
    dose = ref_loader.build_dose(32).eval()
    dose.load_state_dict(synth_ckpt.make_state_dict(synth_ckpt.manifest_of(dose), DOSE_SEED), strict=True)
    seg = ref_loader.build_seg(32).eval()
    seg.load_state_dict(synth_ckpt.make_state_dict(synth_ckpt.manifest_of(seg), SEG_SEED), strict=True)
    with torch.no_grad():
        out = dose(vol["dose_input"])
        logits = seg(vol["ct"])
    np.savez_compressed(os.path.join(OUT, "dose32.npz"), out_A=_np(out[0]),
                        **{f"d{i}": _np(t) for i, t in enumerate(out[1])})
    np.savez_compressed(os.path.join(OUT, "seg32.npz"), logits=_np(logits))

    # ---- the OldModels TRANSEG (what the reference's seg checkpoints / LinkedNet use), 32^3
    old32 = old.TRANSEG(in_channels=1, out_channels=8, img_size=(32, 32, 32), feature_size=16, hidden_size=768, mlp_dim=3072,
                        num_heads=12, pos_embed="perceptron", norm_name="instance", res_block=True, conv_block=True,
                        dropout_rate=0.0).eval()
    old32.load_state_dict(synth_ckpt.make_state_dict(synth_ckpt.manifest_of(old32), 2), strict=True)
    with torch.no_grad():
        np.savez_compressed(os.path.join(OUT, "seg_old32.npz"), logits=_np(old32(vol["ct"][:1])))

    # ---- hand-off exactly as LinkedNet.test_step does it (train_light_linked_model.py:152-169)
    cfg = ref_loader.seg_config()
    from monai.data import decollate_batch
    ct, ptv = vol["ct"][:1], vol["ptv"][:1]
    with torch.no_grad():
        oars = seg(ct)
        oars = [cfg.post_pred(i) for i in decollate_batch(oars)][0]
        oars = torch.permute(oars, (0, 3, 2, 1))
        ct_t = torch.permute(ct, (0, 1, 4, 3, 2))
        oars = torch.unsqueeze(oars, dim=0)[:, 1:, :, :, :]
        structures = torch.cat((ptv, oars, ct_t), dim=1)
        pred = dose(structures)[1][0]
    np.savez_compressed(os.path.join(OUT, "cascade32.npz"), structures=_np(structures), dose=_np(pred))

    # ---- GenLoss (loss.py:69-119) on the reference's own outputs
    loss_mod = ref_loader.loss().GenLoss(im_size=32)
    with torch.no_grad():
        val = loss_mod(out, vol["gt"], casecade=True, freez=True, delta1=10, delta2=8)
    np.savez_compressed(os.path.join(OUT, "genloss32.npz"), loss=np.float64(val.item()))

    # ---- training step (train_light_pyfer.py:85-88,122-143): reference modules in train mode + GenLoss + autograd,
    # then one AdamW step; per-parameter gradient norms, sampled gradient / updated-parameter entries, BN running stats
    tm = ref_loader.build_dose(32).train()
    tm.load_state_dict(synth_ckpt.make_state_dict(synth_ckpt.manifest_of(tm), DOSE_SEED), strict=True)
    for n, p_ in tm.named_parameters():
        if "net_A" in n or "conv_out_A" in n:
            p_.requires_grad = False
    params = [(n, p_) for n, p_ in tm.named_parameters() if p_.requires_grad]
    opt = torch.optim.AdamW([p_ for _, p_ in params], lr=1e-4, weight_decay=1e-4)
    loss = loss_mod(tm(vol["dose_input"]), vol["gt"], casecade=True, freez=True, delta1=10, delta2=8)
    loss.backward()
    rec = {"loss": np.float64(loss.item())}
    names, norms = [], []
    for n, p_ in params:
        g_ = p_.grad if p_.grad is not None else torch.zeros_like(p_)
        names.append(n)
        norms.append(float(g_.double().norm()))
        rec["g/" + n] = _np(g_.flatten()[torch_ref.sample_idx(g_.numel())])
    opt.step()
    for n, p_ in params:
        rec["p/" + n] = _np(p_.detach().flatten()[torch_ref.sample_idx(p_.numel())])
    bn = "net_B.decoder.decoder1.conv_block.cov_.conv_7.0.conv.1."
    sd_after = tm.state_dict()
    rec["running_mean"] = _np(sd_after[bn + "running_mean"])
    rec["running_var"] = _np(sd_after[bn + "running_var"])
    rec["names"] = np.array(names)
    rec["norms"] = np.array(norms, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "train32.npz"), **rec)

    # ---- seg training step (train_light_transeg.py:184-198): the reference module in train mode + autograd; the loss is
    # oracle.torch_ref.dice_ce_loss (monai's DiceCELoss is an un-vendored dependency, restated there)
    sm = ref_loader.build_seg(32).train()
    sm.load_state_dict(synth_ckpt.make_state_dict(synth_ckpt.manifest_of(sm), SEG_SEED), strict=True)
    label = synth.oar_labels(vol["oars"])
    sloss = torch_ref.dice_ce_loss(sm(vol["ct"]), label)
    sloss.backward()
    srec = {"loss": np.float64(sloss.item())}
    snames, snorms = [], []
    for n, p_ in sm.named_parameters():
        g_ = p_.grad if p_.grad is not None else torch.zeros_like(p_)
        snames.append(n)
        snorms.append(float(g_.double().norm()))
        srec["g/" + n] = _np(g_.flatten()[torch_ref.sample_idx(g_.numel())])
    srec["names"] = np.array(snames)
    srec["norms"] = np.array(snorms, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "segtrain32.npz"), **srec)

    # ---- per-volume evaluation (evaluate_openKBP.py): the reference's own numpy functions on a synthetic prediction
    ev = ref_loader.evaluate()
    v64 = synth.make_volume(64, seed=99)
    gt_gy = (v64["gt"][0, 0].numpy() * 70).astype(np.float32)
    pmask = v64["gt"][0, 1].numpy()
    rng = np.random.default_rng(7)
    raw_pred = (gt_gy / 70 + rng.normal(0, 0.03, gt_gy.shape)).astype(np.float32)
    pred_gy = raw_pred.copy()
    pred_gy[np.logical_or(pmask < 1, pred_gy < 0)] = 0            # train_light_pyfer.py:210-213
    pred_gy = 70. * pred_gy
    erec = {"dose_dif": np.float64(ev.get_3D_Dose_dif(pred_gy, gt_gy, pmask)),
            "ivs": np.array([ev.IVS(pred_gy, gt_gy, lv) for lv in np.linspace(0, 70, 101)], dtype=np.float64)}
    sp = (3.906, 3.906, 2.5)
    difs = []
    for name, m in synth.structures(v64).items():
        m = m[0, 0].numpy()
        if not np.any(m):
            continue
        mode = "target" if name.startswith("PTV") else "OAR"
        a, b = ev.get_DVH_metrics(pred_gy, m, mode=mode, spacing=sp), ev.get_DVH_metrics(gt_gy, m, mode=mode, spacing=sp)
        for k_ in b:
            erec["pre" + name + "_" + k_] = np.float64(a[k_])
            erec["gt_" + name + "_" + k_] = np.float64(b[k_])
            difs.append(abs(b[k_] - a[k_]))
    erec["dvh_dif"] = np.float64(np.mean(difs))
    np.savez_compressed(os.path.join(OUT, "eval64.npz"), **erec)

    # ---- input pipeline: the reference's own transform classes (dataloader_OpenKBP_monai.py:84-146) on a small raw case
    dl = ref_loader.dataloader()
    rng = np.random.default_rng(11)
    shp = (12, 10, 16)
    raw = {"CT": rng.integers(-2000, 3000, shp).astype(np.int16), "dose": (rng.random(shp) * 75).astype(np.float32),
           "dose_mask": (rng.random(shp) < 0.6).astype(np.uint8)}
    for n in ("Brainstem", "SpinalCord", "LeftParotid", "Mandible", "PTV70", "PTV56"):
        raw[n] = (rng.random(shp) < 0.2).astype(np.uint8)
    dd = {k_: np.transpose(v_, (2, 1, 0)) for k_, v_ in raw.items()}          # monai Transposed(indices=[2,1,0])
    dd = dl.Empty2FullOAR()(dd)
    dd = dl.NormalizePTVTr()(dd)
    dd = dl.MyIntensityNormalTransform(a_min=-1024, a_max=1500)(dd)
    dd = dl.NormalizeDoseTr()(dd)
    inp_ref = np.stack([dd["PTV"]] + [dd[n] for n in dl.OAR_NAMES] + [dd["CT"]]).astype(np.float32)     # ConcatItemsd 'Input'
    gt_ref = np.stack([dd["dose"], dd["dose_mask"]]).astype(np.float32)                                   # ConcatItemsd 'GT'
    np.savez_compressed(os.path.join(OUT, "pipeline12.npz"), input=inp_ref, gt=gt_ref)

    # ---- sliding window (monai restatement; seg net built for 32^3 scanned over a 48^3 CT)
    from monai.inferers import sliding_window_inference
    ct48 = synth.make_volume(48, seed=77)["ct"]
    with torch.no_grad():
        sw = sliding_window_inference(ct48, (32, 32, 32), 4, seg)
    np.savez_compressed(os.path.join(OUT, "sliding48.npz"), logits=_np(sw[:, :, ::2, ::2, ::2]))
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
