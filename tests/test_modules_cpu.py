"""Drop-in boundary on CPU: constructor surface, state_dict layout (keys/shapes/order) and the
no-fallback rule.  The reference's own modules are used when the tree is present."""
import pytest
import torch

from conftest import load_manifest
from dose_prediction_b200 import networks
from oracle import ref_loader, synth_ckpt


def _manifest_of(m):
    return [(k, list(v.shape), str(v.dtype).replace("torch.", "")) for k, v in m.state_dict().items()]


def _want(name):
    return [(k, list(s), d) for k, s, d in load_manifest(name)]


def test_dose_pyfer_state_dict_layout_matches_reference_manifest():
    m = networks.Model(9, 1, [-1, 16, 32, 64, 128, 256], feature_size=16, img_size=(128, 128, 128), num_layers=8,
                       num_heads=6, act="mish", mode_multi_dec=True, multiS_conv=True)
    assert _manifest_of(m) == _want("dose_pyfer")


@pytest.mark.parametrize("in_ch,name", [(1, "oar_transeg"), (2, "oar_transeg_2ch")])
def test_oar_transeg_state_dict_layout_matches_reference_manifest(in_ch, name):
    m = networks.OARTranseg(in_channels=in_ch, out_channels=8, img_size=(128, 128, 128), feature_size=16, hidden_size=768,
                            mlp_dim=3072, num_heads=12, pos_embed="perceptron", norm_name="instance", res_block=True,
                            conv_block=True, dropout_rate=0.0)
    assert _manifest_of(m) == _want(name)


def test_constructor_validation_matches_reference():
    with pytest.raises(ValueError, match="dropout_rate"):
        networks.OARTranseg(1, 8, 32, dropout_rate=1.5, pos_embed="perceptron")
    with pytest.raises(ValueError, match="divisible by num_heads"):
        networks.OARTranseg(1, 8, 32, hidden_size=770, pos_embed="perceptron")
    with pytest.raises(ValueError, match="divisible by num_heads"):
        networks.ViTEncoder(25, 32, hidden_size=768, num_heads=7, pos_embed="perceptron")


def test_no_cpu_fallback():
    m = networks.OARTranseg(1, 8, 32, pos_embed="perceptron").eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 1, 32, 32, 32))
    # train mode is the autograd shim (training.autograd_forward_seg): CUDA only as well
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        networks.OARTranseg(1, 8, 32, pos_embed="perceptron").train()(torch.zeros(1, 1, 32, 32, 32))
    # sub-networks keep an inference-only forward
    with pytest.raises(RuntimeError, match="eval"):
        networks.BaseUNet(9, [-1, 16, 32, 64, 128, 256]).train()(torch.zeros(1, 9, 32, 32, 32))


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree only exists in the build container")
def test_state_dict_round_trips_with_live_reference_modules():
    ref = ref_loader.build_dose(32)
    ours = networks.Model(9, 1, [-1, 16, 32, 64, 128, 256], img_size=(32, 32, 32))
    ours.load_state_dict(ref.state_dict(), strict=True)
    ref.load_state_dict(ours.state_dict(), strict=True)
    assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())     # positional copies rely on order
    rs = ref_loader.build_seg(32)
    os_ = networks.OARTranseg(1, 8, (32, 32, 32), pos_embed="perceptron")
    os_.load_state_dict(rs.state_dict(), strict=True)
    assert list(os_.state_dict().keys()) == list(rs.state_dict().keys())


def test_synthetic_checkpoint_is_deterministic():
    man = [("net_A.encoder.encoder_1.0.single_conv.0.weight", [16, 9, 3, 3, 3]), ("x.running_var", [4])]
    a, b = synth_ckpt.make_state_dict(man, 3), synth_ckpt.make_state_dict(man, 3)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert not torch.equal(a[man[0][0]], synth_ckpt.make_state_dict(man, 4)[man[0][0]])


def test_old_transeg_state_dict_layout_matches_reference_manifest():
    m = networks.TRANSEG(in_channels=1, out_channels=8, img_size=(96, 96, 96), feature_size=16, hidden_size=768, mlp_dim=3072,
                         num_heads=12, pos_embed="perceptron", norm_name="instance", res_block=True, conv_block=True)
    assert _manifest_of(m) == _want("transeg_old_96")


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree only exists in the build container")
def test_variant_constructors_match_live_reference_layout():
    ref = ref_loader.build_dose(32, multiS_conv=False, act="relu")
    ours = networks.Model(9, 1, [-1, 16, 32, 64, 128, 256], img_size=(32, 32, 32), multiS_conv=False, act="relu")
    assert _manifest_of(ours) == _manifest_of(ref)
    ours.load_state_dict(ref.state_dict(), strict=True)
    ref2 = ref_loader.build_dose(32, mode_multi_dec=False)
    ours2 = networks.Model(9, 1, [-1, 16, 32, 64, 128, 256], img_size=(32, 32, 32), mode_multi_dec=False)
    assert _manifest_of(ours2) == _manifest_of(ref2)
    ref3 = ref_loader.build_seg(32, pos_embed="conv")
    ours3 = networks.OARTranseg(1, 8, (32, 32, 32), pos_embed="conv")
    assert _manifest_of(ours3) == _manifest_of(ref3)


def test_reference_named_modules_export_the_reference_class_names():
    """a user of the reference swaps the import path, not the names."""
    from dose_prediction_b200 import base_blocks, c3d, dose_pyfer, networks, oar_transeg
    assert dose_pyfer.Model is networks.Model and dose_pyfer.create_pretrained_unet is networks.create_pretrained_unet
    assert {"ViTEncoder", "PyMSCDecoder", "MainSubsetModel"} <= set(dose_pyfer.__all__)
    assert oar_transeg.Model is networks.OARTranseg and oar_transeg.TRANSEG is networks.TRANSEG
    assert c3d.BaseUNet is networks.BaseUNet and base_blocks.ModifiedUnetrUpBlock is networks.ModifiedUnetrUpBlock
    import inspect
    for cls, params in ((base_blocks.ModifiedUnetrUpBlock, ["inp", "skip"]), (dose_pyfer.ViTEncoder, ["x_in"]),
                        (dose_pyfer.PyMSCDecoder, ["out_encoder"]), (oar_transeg.Model, ["x_in"]), (dose_pyfer.Model, ["x"])):
        assert list(inspect.signature(cls.forward).parameters)[1:] == params, cls


def test_cpu_tensors_and_train_mode_sub_blocks_fail_loudly():
    import pytest
    import torch
    from dose_prediction_b200 import networks
    blk = networks.ModifiedUnetrUpBlock(3, 32, 16, 2)
    with pytest.raises(RuntimeError):
        blk(torch.zeros(1, 32, 4, 4, 4), torch.zeros(1, 16, 8, 8, 8))          # train mode
    with pytest.raises(RuntimeError):
        blk.eval()(torch.zeros(1, 32, 4, 4, 4), torch.zeros(1, 16, 8, 8, 8))   # CPU tensor: no fallback
