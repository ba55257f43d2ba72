"""TEST INFRASTRUCTURE — import the reference's model files unmodified (build container only).

Puts `oracle/monai_compat` (for `import monai`) and `/root/reference` on sys.path and imports
  DosePrediction.Models.Networks.dose_pyfer   (dose_pyfer.py:325 Model)
  OARSegmentation.Models.Networks.oar_transeg (oar_transeg.py:14 Model)
  DosePrediction.Train.loss                   (loss.py:50 GenLoss)
On the GPU box (no /root/reference there) the same imports resolve against oracle/_ref, the git-ignored copy of exactly
these files that oracle/make_ref.py makes in the build container; only tests/, bench.py's reference arm / cpu_baseline
and __graft_entry__.smoke() may use this module, never the product.
"""
import contextlib
import importlib
import io
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_PREBUILT = os.path.join(_HERE, "_ref")          # oracle/make_ref.py: unmodified copies of the files the path imports


def _default_root():
    if os.environ.get("DP_REFERENCE_ROOT"):
        return os.environ["DP_REFERENCE_ROOT"]
    if os.path.isdir("/root/reference/DosePrediction"):
        return "/root/reference"
    return _PREBUILT                                # the GPU box: /root/reference does not exist there


REFERENCE_ROOT = _default_root()
_COMPAT = os.path.join(_HERE, "monai_compat")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "DosePrediction"))


def _ensure_path():
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for p in (REFERENCE_ROOT, _COMPAT):
        if p not in sys.path:
            sys.path.insert(0, p)


def _imp(name):
    _ensure_path()
    with contextlib.redirect_stdout(io.StringIO()):
        return importlib.import_module(name)


def dose_pyfer():
    return _imp("DosePrediction.Models.Networks.dose_pyfer")


def oar_transeg():
    return _imp("OARSegmentation.Models.Networks.oar_transeg")


def oar_transeg_old():
    return _imp("OARSegmentation.OldModels.Networks.oar_transeg")


def loss():
    return _imp("DosePrediction.Train.loss")


def evaluate():
    """DosePrediction/Evaluate/evaluate_openKBP.py, unmodified.  Its module-level imports of SimpleITK, matplotlib and
    monai.metrics are unused by the functions the oracle is pinned against (IVS, get_3D_Dose_dif, get_DVH_metrics);
    stand-in modules are registered for whichever of them is not installed."""
    import types
    for name in ("SimpleITK", "matplotlib", "matplotlib.pyplot", "matplotlib.colors", "monai.metrics"):
        _ensure_path()
        try:
            importlib.import_module(name)
        except Exception:
            m = types.ModuleType(name)
            if name == "monai.metrics":
                m.DiceMetric = lambda *a, **k: None
            sys.modules[name] = m
            if "." in name:
                parent, child = name.rsplit(".", 1)
                if parent in sys.modules:
                    setattr(sys.modules[parent], child, m)
    return _imp("DosePrediction.Evaluate.evaluate_openKBP")


def dataloader():
    """DosePrediction/DataLoader/dataloader_OpenKBP_monai.py, unmodified.  The monai names it imports at module level
    (dictionary transforms, datasets) are only used inside prepare_data / get_dataset; placeholders are registered for
    the ones oracle/monai_compat does not restate, so that the reference's OWN transform classes (Empty2FullOAR,
    NormalizePTVTr, MyIntensityNormalTransform, NormalizeDoseTr) can be imported and run."""
    _ensure_path()
    import monai.data
    import monai.transforms
    for mod, names in ((monai.data, ("Dataset", "DataLoader", "CacheDataset", "list_data_collate")),
                       (monai.transforms, ("Compose", "LoadImaged", "ToTensord", "AddChanneld", "Orientationd", "ConcatItemsd",
                                           "DeleteItemsd", "RandCropByPosNegLabeld", "RandFlipd", "RandShiftIntensityd",
                                           "RandRotate90d", "Transposed"))):
        for n in names:
            if not hasattr(mod, n):
                setattr(mod, n, type(n, (), {"__init__": lambda self, *a, **k: None}))
    return _imp("DosePrediction.DataLoader.dataloader_OpenKBP_monai")


def seg_config():
    return _imp("OARSegmentation.config")


def build_dose(img=128, **kw):
    """Reference DOSE-PYFER with the hot-path ctor args (train_light_pyfer.py:73-83)."""
    m = dose_pyfer()
    args = dict(in_ch=9, out_ch=1, list_ch_A=[-1, 16, 32, 64, 128, 256], feature_size=16,
                img_size=(img, img, img), num_layers=8, num_heads=6, act="mish",
                mode_multi_dec=True, multiS_conv=True)
    args.update(kw)
    with contextlib.redirect_stdout(io.StringIO()):
        return m.Model(**args)


def build_seg(img=128, in_channels=1, **kw):
    """Reference OAR-TRANSEG with the hot-path ctor args (train_light_transeg.py:110-124)."""
    m = oar_transeg()
    args = dict(in_channels=in_channels, out_channels=8, img_size=(img, img, img), feature_size=16,
                hidden_size=768, mlp_dim=3072, num_heads=12, pos_embed="perceptron",
                norm_name="instance", res_block=True, conv_block=True, dropout_rate=0.0)
    args.update(kw)
    return m.Model(**args)
