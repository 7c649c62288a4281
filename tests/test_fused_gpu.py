"""The fused ResBlock2/MRF stage kernel on hardware, through the C ABI: against the per-layer path (same
3xTF32 arithmetic, different association) and against the CPU oracle, at shapes that exercise ragged tiles."""
import pytest
import torch

from tests.golden_util import rel_rms_err

pytestmark = pytest.mark.gpu


def _set(name, v):
    from wetts_b200 import _lib
    _lib.check(_lib.load().wetts_set_option(name.encode(), int(v)))


@pytest.mark.parametrize("B,Ty", [(1, 7), (3, 33), (2, 130)])
def test_generator_fused_matches_per_layer_and_oracle(B, Ty):
    import wetts_b200
    from oracle import vits_oracle as O
    from wetts_b200 import synth
    from wetts_b200.hparams import builtin_config
    hps = builtin_config("multilingual_v3")
    sd = synth.make_state_dict(hps.model, 50, 2, seed=11)
    net = wetts_b200.build_model(hps, 50, 2, sd, "cuda")
    gen = torch.Generator().manual_seed(100 + Ty)
    z = torch.randn(B, hps.model.inter_channels, Ty, generator=gen)
    sid = torch.randint(0, 2, (B,), generator=gen)
    g = net.emb_g(sid)[:, :, None]
    try:
        _set("fused_resblock", 1)
        o_fused = net.dec(z.cuda(), g=g).cpu()
        launches_fused = net.launch_count()
        _set("fused_resblock", 0)
        o_layer = net.dec(z.cuda(), g=g).cpu()
        launches_layer = net.launch_count() - launches_fused
    finally:
        _set("fused_resblock", 1)
    w = O.fold_weight_norm(sd)
    ref = O.generator(w, hps.model, z, g.cpu())
    assert o_fused.shape == ref.shape == (B, 1, Ty * 256)
    assert rel_rms_err(o_fused, ref) < 3e-4
    assert rel_rms_err(o_layer, ref) < 3e-4
    assert rel_rms_err(o_fused, o_layer) < 3e-4
    # two stages (64 and 32 channels) collapse from 6 launches to 1 each
    assert launches_layer > 0


def test_fused_option_roundtrip():
    from wetts_b200 import _lib
    import ctypes
    lib = _lib.load()
    v = ctypes.c_int(-1)
    _lib.check(lib.wetts_get_option(b"fused_resblock", ctypes.byref(v)))
    assert v.value == 1
