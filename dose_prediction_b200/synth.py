"""Synthetic OpenKBP-shaped inputs (there is no dataset offline).

Channel contract of the dose net input, reference DosePrediction/DataLoader/dataloader_OpenKBP_monai.py:
  :195-197  Input = [PTV, Brainstem, SpinalCord, RightParotid, LeftParotid, Esophagus, Larynx, Mandible, CT]
  :116-125  PTV in {0, 56/70, 63/70, 1.0} (nested targets);  :137-146 CT = clip(HU,-1024,1500)/1000
  :128-134,199-201  GT = [dose/70, possible_dose_mask]
Everything is a closed-form function of (seed, voxel index) plus a seeded torch CPU generator, so
the build container and the GPU box produce bit-identical volumes.
"""
import torch


def _grid(size):
    ax = torch.linspace(-1.0, 1.0, size)
    return torch.meshgrid(ax, ax, ax, indexing="ij")


def _ellipsoid(g, c, r):
    return (((g[0] - c[0]) / r[0]) ** 2 + ((g[1] - c[1]) / r[1]) ** 2 + ((g[2] - c[2]) / r[2]) ** 2) <= 1.0


def make_volume(size: int = 128, seed: int = 1234):
    """Returns dict(ct [1,1,S,S,S], ptv [1,1,...], oars [1,7,...], dose_input [1,9,...], gt [1,2,...]) fp32."""
    gen = torch.Generator(device="cpu").manual_seed(int(seed))
    g = _grid(size)
    u = torch.rand(64, generator=gen)
    # CT: smooth anatomy-like field + noise, clipped to the loader's window
    ph = u[:6] * 6.283
    ct = 0.45 * torch.sin(3.1 * g[0] + ph[0]) * torch.cos(2.3 * g[1] + ph[1]) \
        + 0.35 * torch.sin(4.7 * g[2] + ph[2]) + 0.25 * torch.cos(5.3 * (g[0] + g[1]) + ph[3])
    ct = ct + 0.15 * torch.randn(ct.shape, generator=gen)
    body = _ellipsoid(g, (0.0, 0.0, 0.0), (0.92, 0.85, 0.95))
    ct = torch.where(body, ct, torch.full_like(ct, -1.0)).clamp_(-1.024, 1.5)
    # PTV: nested targets 56/63/70 Gy -> 0.8/0.9/1.0
    c = (u[6:9] - 0.5) * 0.3
    ptv = torch.zeros_like(ct)
    for val, rad in ((0.8, 0.42), (0.9, 0.30), (1.0, 0.18)):
        ptv = torch.where(_ellipsoid(g, c, (rad, rad * 0.9, rad * 1.1)), torch.full_like(ptv, val), ptv)
    # 7 disjoint OAR ellipsoids on a ring around the target
    oars = []
    for k in range(7):
        ang = 6.283 * (k + u[10 + k] * 0.3) / 7.0
        cen = (0.62 * torch.cos(ang), 0.58 * torch.sin(ang), (u[20 + k] - 0.5) * 0.8)
        rad = (0.10 + 0.06 * u[30 + k], 0.09 + 0.05 * u[37 + k], 0.12 + 0.10 * u[44 + k])
        oars.append(_ellipsoid(g, cen, rad).float())
    oars = torch.stack(oars)
    dist = torch.sqrt((g[0] - c[0]) ** 2 + (g[1] - c[1]) ** 2 + (g[2] - c[2]) ** 2)
    dose = (1.08 * torch.exp(-(dist / 0.55) ** 2)).clamp_(0, 1.1) * body
    gt = torch.stack((dose, body.float()))[None]
    ct, ptv = ct[None, None], ptv[None, None]
    dose_input = torch.cat((ptv, oars[None], ct), dim=1)
    return {"ct": ct.contiguous(), "ptv": ptv.contiguous(), "oars": oars[None].contiguous(),
            "dose_input": dose_input.contiguous(), "gt": gt.contiguous()}


def make_batch(batch: int, size: int = 128, seed: int = 1234):
    vols = [make_volume(size, seed + i) for i in range(batch)]
    return {k: torch.cat([v[k] for v in vols]) for k in vols[0]}


def oar_labels(oars):
    """[B,7,...] binary OAR masks -> [B,1,...] fp32 class-index map (0 = background, k+1 = OAR k), the 'OARs' label
    volume the seg training step consumes (OARSegmentation/train_light_transeg.py:185)."""
    idx = torch.arange(1, oars.shape[1] + 1, dtype=oars.dtype).view(1, -1, 1, 1, 1)
    return (oars * idx).amax(dim=1, keepdim=True).contiguous()


def structures(vol):
    """{structure name: [B,1,...] binary mask} in the reference's naming (evaluate_openKBP.py:176-186): the seven
    OARs and the three nested targets PTV70 / PTV63 / PTV56 (PTV map values 1.0 / 0.9 / 0.8)."""
    names = ["Brainstem", "SpinalCord", "RightParotid", "LeftParotid", "Esophagus", "Larynx", "Mandible"]
    out = {n: vol["oars"][:, i:i + 1].contiguous() for i, n in enumerate(names)}
    for n, val in (("PTV70", 1.0), ("PTV63", 0.9), ("PTV56", 0.8)):
        out[n] = (vol["ptv"] - val).abs().lt(1e-6).float()
    return out
