"""TEST INFRASTRUCTURE — deterministic synthetic checkpoints.

The reference ships no checkpoints (PretrainedModels/*/PATH_TO_PRETRANED_MODELS are Drive links), and
168 M parameters cannot be committed as fixtures.  Instead every tensor of a state_dict is generated
from (key, shape, seed) alone, with the magnitudes of the reference's own initialisers:
  * net_A convs: kaiming-uniform(relu) bound sqrt(6/fan_in)           (c3d.py:127-142)
  * patch-embedding Linear / position embeddings: N(0, 0.02)         (monai PatchEmbeddingBlock)
  * every other conv/linear/deconv weight: U(+-1/sqrt(fan_in))        (torch default)
  * biases U(+-0.05); norm scales 1+U(+-0.1); norm shifts U(+-0.1)
  * BatchNorm running_mean U(+-0.2), running_var U(0.5,1.5) so BN folding is exercised
so that the build container (real reference modules) and the GPU box (product modules / torch_ref)
reconstruct bit-identical weights without shipping them.
"""
import math
import zlib

import torch


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) * 2654435761 + seed * 97) % (2 ** 63))
    return g


def _uniform(shape, lo, hi, g):
    return torch.empty(shape, dtype=torch.float32).uniform_(lo, hi, generator=g)


def make_tensor(key: str, shape, dtype=torch.float32, seed: int = 0) -> torch.Tensor:
    shape = tuple(shape)
    g = _gen(key, seed)
    leaf = key.rsplit(".", 1)[-1]
    if leaf == "num_batches_tracked":
        return torch.zeros(shape, dtype=torch.int64)
    if leaf == "running_mean":
        return _uniform(shape, -0.2, 0.2, g)
    if leaf == "running_var":
        return _uniform(shape, 0.5, 1.5, g)
    if leaf == "cls_token":
        return torch.zeros(shape)
    if leaf == "position_embeddings":
        return torch.empty(shape).normal_(0.0, 0.02, generator=g)
    if leaf == "bias":
        return _uniform(shape, -0.1, 0.1, g) if _is_norm(key) else _uniform(shape, -0.05, 0.05, g)
    if leaf == "weight":
        if len(shape) == 1:
            return 1.0 + _uniform(shape, -0.1, 0.1, g)
        if "patch_embeddings" in key:
            return torch.empty(shape).normal_(0.0, 0.02, generator=g)
        fan_in = shape[1] * math.prod(shape[2:])
        bound = math.sqrt(6.0 / fan_in) if "net_A." in key else 1.0 / math.sqrt(fan_in)
        return _uniform(shape, -bound, bound, g)
    raise KeyError(f"synth_ckpt: unrecognised state_dict leaf in {key!r}")


def _is_norm(key: str) -> bool:
    # norm shifts live beside a 1-D weight; callers pass full manifests so this is only a magnitude hint
    return any(t in key for t in (".norm", "single_conv.1.", ".conv.1.", ".conv.4.", "norm1", "norm2"))


def make_state_dict(manifest, seed: int = 0):
    """manifest: iterable of (key, shape) or (key, shape, dtype-string); returns an ordered dict."""
    out = {}
    for ent in manifest:
        key, shape = ent[0], ent[1]
        out[key] = make_tensor(key, shape, seed=seed)
    return out


def manifest_of(module_or_sd):
    sd = module_or_sd.state_dict() if hasattr(module_or_sd, "state_dict") else module_or_sd
    return [(k, list(v.shape), str(v.dtype).replace("torch.", "")) for k, v in sd.items()]
