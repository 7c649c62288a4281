"""GPU parity of the f16-split fused MRF stage kernel (tensor_format = 16) through the C ABI: against the
reference-generated fixtures (ResBlock2 = v3, ResBlock1 = v1 recipes), against the 3xTF32 path, against the CPU oracle on
a ragged batch whose length is not a multiple of the 128-sample tile, and the length-aware mode (valid samples must be
BIT-identical to the full computation: the same work items compute the same values, only the item list differs)."""
import pytest
import torch

from tests.golden_util import load_case, rel_rms_err

pytestmark = pytest.mark.gpu

GEN_TOL = 3e-4      # generator block tolerance (40-conv stack), as in tests/test_parity_gpu.py


def _net(name):
    import wetts_b200
    hps, sd, g, t = load_case(name)
    net = wetts_b200.build_model(hps, int(g["n_vocab"]), int(g["n_speakers"]), sd, "cuda")
    return net, hps, g, t


@pytest.mark.parametrize("name", ["v3_ragged", "v3_tx128", "v1_ragged", "aishell3_long"])
def test_generator_f16_path_matches_reference_fixture(name):
    net, hps, g, t = _net(name)
    dev = net.device
    gvec = net.emb_g(t["sid"])[:, :, None] if int(g["n_speakers"]) > 0 else None
    Ty = t["z"].shape[2]
    ym = (torch.arange(Ty)[None, :] < t["y_lengths"][:, None]).float()[:, None].to(dev)
    z = t["z"].to(dev) * ym
    net.set_option("tensor_format", 16)
    o16 = net.dec(z, g=gvec)
    net.set_option("tensor_format", 32)
    o32 = net.dec(z, g=gvec)
    torch.cuda.synchronize()
    e16, e32 = rel_rms_err(o16.cpu(), t["o"]), rel_rms_err(o32.cpu(), t["o"])
    print(f"{name}: generator err/rms f16-split {e16:.3e}, 3xTF32 {e32:.3e}, f16 vs tf32 {rel_rms_err(o16.cpu(), o32.cpu()):.3e}")
    assert e16 < GEN_TOL and e32 < GEN_TOL


def test_f16_path_end_to_end_v3_tx128():
    """whole infer with the f16 stage kernels: teacher-forced durations, injected noise"""
    net, hps, g, t = _net("v3_tx128")
    net.set_option("tensor_format", 16)
    ns, ls, nsw = [float(v) for v in g["scales"]]
    o, _, _, (z, z_p, _, _) = net.infer(t["x"], t["x_lengths"], t["sid"], ns, ls, nsw, noise_z=t["noise_z"],
                                        durations=t["w_ceil"], return_attn=False)
    assert torch.equal(net.last_y_lengths.cpu(), t["y_lengths"])
    assert rel_rms_err(o.cpu(), t["o"]) < 1e-3


@pytest.mark.parametrize("cfg_name,n_spk", [("multilingual_v3", 2), ("baker_v1", 1)])
def test_length_aware_mode_leaves_valid_samples_bit_identical(cfg_name, n_spk):
    import wetts_b200
    from wetts_b200 import synth
    from wetts_b200.hparams import builtin_config
    hps = builtin_config(cfg_name)
    sd = synth.make_state_dict(hps.model, 64, n_spk, seed=11)
    net = wetts_b200.build_model(hps, 64, n_spk, sd, "cuda")
    net.set_option("tensor_format", 16)
    gen = torch.Generator().manual_seed(5)
    B, Tx = 4, 48
    x = torch.randint(0, 64, (B, Tx), generator=gen)
    lens = torch.tensor([48, 17, 33, 5])
    sid = torch.randint(0, n_spk, (B,), generator=gen)
    dur = torch.randint(1, 6, (B, 1, Tx), generator=gen).float() * (torch.arange(Tx)[None, None, :] < lens[:, None, None])
    nz = torch.randn(B, 192, int(dur.sum(-1).max()), generator=gen)
    nw = torch.randn(B, 2, Tx, generator=gen)

    def run():
        o, _, ym, _ = net.infer(x, lens, sid, 0.667, 1.0, 0.8, noise_w=nw, noise_z=nz, durations=dur, return_attn=False)
        torch.cuda.synchronize()
        return o, net.last_y_lengths.clone()

    o_full, yl = run()
    net.set_option("length_aware", 1)
    o_la, yl2 = run()
    assert torch.equal(yl, yl2)
    for b in range(B):
        n = int(yl[b]) * 256
        assert torch.equal(o_la[b, 0, :n], o_full[b, 0, :n]), f"utterance {b}: valid samples differ in length-aware mode"


def test_f16_path_against_oracle_random_ragged():
    from oracle import vits_oracle as O
    from wetts_b200 import synth
    from wetts_b200.hparams import builtin_config
    import wetts_b200
    for cfg_name, n_spk in (("multilingual_v3", 2), ("baker_v1", 1)):
        hps = builtin_config(cfg_name)
        sd = synth.make_state_dict(hps.model, 100, n_spk, seed=7)
        gen = torch.Generator().manual_seed(123)
        B, T = 3, 37                         # 37 frames: every stage length is ragged against the 128-sample tile
        z = torch.randn(B, 192, T, generator=gen)
        sid = torch.randint(0, n_spk, (B,), generator=gen)
        w = O.fold_weight_norm(sd)
        gv = w["emb_g.weight"][sid][:, :, None]
        ref = O.generator(w, hps.model, z, gv)
        net = wetts_b200.build_model(hps, 100, n_spk, sd, "cuda")
        net.set_option("tensor_format", 16)
        o = net.dec(z, g=net.emb_g(sid)[:, :, None])
        assert rel_rms_err(o.cpu(), ref) < GEN_TOL


def test_work_item_sizes_of_the_c32_stage_are_bit_identical_and_match_the_oracle():
    """The 32-channel ResBlock2 stage kernel picks 384-sample work items for large launches (what bench.py runs) and
    128-sample items for small ones (what the fixture tests above run): every size must give the same bits, and the
    large-launch route is checked against the CPU oracle directly (ragged against 128, 256 and 384)."""
    import ctypes
    import wetts_b200
    from oracle import vits_oracle as O
    from wetts_b200 import _lib, synth
    from wetts_b200.hparams import builtin_config
    lib = _lib.load()
    hps = builtin_config("multilingual_v3")
    sd = synth.make_state_dict(hps.model, 100, 2, seed=7)
    gen = torch.Generator().manual_seed(321)
    B, T = 8, 301                            # 8 x 77056 samples at the last stage: 1608 items of 384 >= 8 per SM
    z = torch.randn(B, 192, T, generator=gen)
    sid = torch.randint(0, 2, (B,), generator=gen)
    net = wetts_b200.build_model(hps, 100, 2, sd, "cuda")
    net.set_option("tensor_format", 16)
    gv = net.emb_g(sid)[:, :, None]
    last = ctypes.c_int(0)
    outs = {}
    try:
        for rows in (0, 128, 256, 384):
            _lib.check(lib.wetts_set_option(b"mrf_item_rows", rows))
            outs[rows] = net.dec(z, g=gv).clone()
            torch.cuda.synchronize()
            _lib.check(lib.wetts_get_option(b"mrf_item_rows_last", ctypes.byref(last)))
            assert last.value == (rows if rows else 384), f"requested {rows}, the launch used {last.value}"
    finally:
        _lib.check(lib.wetts_set_option(b"mrf_item_rows", 0))
    for rows in (128, 256, 384):
        assert torch.equal(outs[rows], outs[0]), f"{rows}-sample items differ from the default"
    w = O.fold_weight_norm(sd)
    ref = O.generator(w, hps.model, z[:2], w["emb_g.weight"][sid[:2]][:, :, None])
    assert rel_rms_err(outs[0][:2].cpu(), ref) < GEN_TOL
