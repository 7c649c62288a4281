"""GPU parity on the wider reference fixtures -- the ones that reach the route bench.py measures: every text-encoder /
duration-predictor / flow / generator convolution takes the tcgen05 kernel at T >= 64.

  aishell3_long  BASELINE.json configs[4] family: AISHELL-3 v1, 218 speakers, SDP, HiFi-GAN V1, 128 / 80 phonemes
  baker_v1_cli   configs[0]: Baker v1, the CLI utterance at the CLI's scales
  v3_tx128       configs[2]'s route: multilingual v3, DurationPredictor, ragged 128 / 97 / 64 phonemes

All gating (no xfail).  Tolerances, relative to rms(reference) as everywhere: blocks 1e-4 (generator / flow 3e-4),
end to end 1e-3, integer outputs exact -- including the frame counts from the GPU's OWN durations on the tensor-core
route.  Every case runs on both operand formats of the tensor-pipe kernels:
  fmt16 (default): f16 split, K = 16 channels per MMA.  Measured on B200 (profiles/r02_tc_numerics.txt): text encoder
         1.0e-5, z_p 3.2e-5 .. 3.4e-5, generator 3e-6 .. 1.4e-5 -- everything inside the 1e-4 block tolerance.
  fmt32: 3xTF32, K = 8 per MMA: twice the sequential fp32 accumulations into TMEM, whose truncation is the dominant
         error of the tensor route (profiles/r02_mma_numerics_and_rates.txt: it grows linearly with the number of
         accumulations and halves when the accumulations are dealt over two accumulators).  Text encoder 5e-5 .. 6e-5;
         z_p = m_p + noise * exp(logs_p) * noise_scale multiplies the error of logs_p by up to |noise| exp(logs_p) 0.667
         (11x on aishell3_long): 1.5e-4 .. 2.1e-4, so z_p / z are held to 3e-4 on THIS format only.
This is what round 1 could not explain (its CPU probe modelled the operand split, not the accumulation): not a staging
or masking bug -- the fp32 SIMT route is at 3e-6 on the same inputs, the f16 route at 1e-5."""
import pytest
import torch

from tests.golden_util import WIDE_CASES, load_case, rel_rms_err

pytestmark = pytest.mark.gpu

BLOCK_TOL = 1e-4
E2E_TOL = 1e-3


@pytest.mark.parametrize("fmt", [32, 16])
@pytest.mark.parametrize("name", WIDE_CASES)
def test_wide_fixture_end_to_end_and_blocks(name, fmt):
    """fmt = operand format of the tensor-pipe kernels: 32 = 3xTF32, 16 = f16 split (both must hold the same bounds)"""
    import wetts_b200
    hps, sd, g, t = load_case(name)
    net = wetts_b200.build_model(hps, int(g["n_vocab"]), int(g["n_speakers"]), sd, "cuda")
    net.set_option("tensor_format", fmt)
    name = f"{name}[fmt{fmt}]"
    dev = net.device
    ns, ls, nsw = [float(v) for v in g["scales"]]
    o, attn, y_mask, (z, z_p, m_p, logs_p) = net.infer(
        t["x"], t["x_lengths"], t["sid"], noise_scale=ns, length_scale=ls, noise_scale_w=nsw,
        noise_w=t["noise_w"], noise_z=t["noise_z"], durations=t["w_ceil"])
    torch.cuda.synchronize()
    assert torch.equal(net.last_y_lengths.cpu(), t["y_lengths"])
    e_zp, e_z, e_o = rel_rms_err(z_p.cpu(), t["z_p"]), rel_rms_err(z.cpu(), t["z"]), rel_rms_err(o.cpu(), t["o"])
    print(f"{name}: e2e z_p {e_zp:.3e}  z {e_z:.3e}  o {e_o:.3e}")
    assert e_zp < (BLOCK_TOL if fmt == 16 else BLOCK_TOL * 3)    # 3xTF32: accumulation error x exp(logs_p) amplification
    assert e_z < BLOCK_TOL * 3
    assert e_o < E2E_TOL
    a = attn[:, 0].cpu()
    valid = t["attn_rowsum"] > 0
    assert torch.equal(a.sum(-1), t["attn_rowsum"])
    assert torch.equal(a.argmax(-1)[valid].int(), t["attn_argmax"][valid])
    # blocks given the reference's intermediates
    gvec = net.emb_g(t["sid"])[:, :, None] if int(g["n_speakers"]) > 0 else None
    h, m, logs, x_mask = net.enc_p(t["x"], t["x_lengths"])
    e_h, e_m, e_l = rel_rms_err(h.cpu(), t["h"]), rel_rms_err(m.cpu(), t["m_p_tx"]), rel_rms_err(logs.cpu(), t["logs_p_tx"])
    print(f"{name}: text encoder h {e_h:.3e}  m {e_m:.3e}  logs {e_l:.3e}")
    assert e_h < BLOCK_TOL and e_m < BLOCK_TOL and e_l < BLOCK_TOL
    if net.use_sdp:
        logw = net.dp(t["h"].to(dev), x_mask, g=gvec, reverse=True, noise_scale=nsw, noise=t["noise_w"])
    else:
        logw = net.dp(t["h"].to(dev), x_mask, g=gvec)
    e_w = rel_rms_err(logw.cpu(), t["logw"])
    print(f"{name}: logw {e_w:.3e}")
    assert e_w < 2e-4
    Ty = t["z"].shape[2]
    ym = (torch.arange(Ty)[None, :] < t["y_lengths"][:, None]).float()[:, None].to(dev)
    e_f = rel_rms_err(net.flow(t["z_p"].to(dev), ym, g=gvec, reverse=True).cpu(), t["z"])
    e_g = rel_rms_err(net.dec(t["z"].to(dev) * ym, g=gvec).cpu(), t["o"])
    print(f"{name}: flow {e_f:.3e}  generator {e_g:.3e}")
    assert e_f < BLOCK_TOL
    assert e_g < BLOCK_TOL * 3


@pytest.mark.parametrize("fmt", [32, 16])
@pytest.mark.parametrize("name", WIDE_CASES)
def test_own_durations_on_the_tensor_core_route_are_exact(name, fmt):
    """No teacher forcing: ceil(exp(logw) * length_scale) from the GPU's own text encoder + duration predictor (tcgen05
    route at Tx >= 64) must land on the reference's integers -- the numerator of the benchmark's metric."""
    import wetts_b200
    hps, sd, g, t = load_case(name)
    net = wetts_b200.build_model(hps, int(g["n_vocab"]), int(g["n_speakers"]), sd, "cuda")
    net.set_option("tensor_format", fmt)
    ns, ls, nsw = [float(v) for v in g["scales"]]
    net.infer(t["x"], t["x_lengths"], t["sid"], noise_scale=ns, length_scale=ls, noise_scale_w=nsw,
              noise_w=t["noise_w"], noise_z=t["noise_z"], return_attn=False)
    assert torch.equal(net.last_y_lengths.cpu(), t["y_lengths"])


def test_text_encoder_tcgen05_vs_simt_at_tx128():
    """The text encoder at Tx = 128 (ragged 128 / 80) through the tcgen05 route against the fp32 SIMT route and the CPU
    oracle.  Both must be inside the 1e-4 block tolerance; the SIMT route (summation-order noise only) inside 2e-5.
    The measured errors are printed (run with -s / -rA)."""
    import wetts_b200
    from oracle import vits_oracle as O
    from wetts_b200 import synth
    from wetts_b200.hparams import builtin_config
    hps = builtin_config("aishell3_v1")
    sd = synth.make_state_dict(hps.model, 256, 218, seed=hps.train.seed)
    net = wetts_b200.build_model(hps, 256, 218, sd, "cuda")
    gen = torch.Generator().manual_seed(5682)
    x = torch.randint(0, 256, (2, 128), generator=gen)
    lens = torch.tensor([128, 80])
    ref = O.text_encoder(O.fold_weight_norm(sd), hps.model, x, lens)
    out = {}
    for tc in (1, 0):
        net.set_option("tensor_cores", tc)
        out[tc] = [v.cpu() for v in net.enc_p(x, lens)[:3]]
    net.set_option("tensor_cores", -1)
    for i, nm in enumerate(("h", "m", "logs")):
        e_tc, e_simt = rel_rms_err(out[1][i], ref[i]), rel_rms_err(out[0][i], ref[i])
        print(f"text encoder Tx=128 {nm}: tcgen05 vs oracle {e_tc:.3e}, SIMT vs oracle {e_simt:.3e}, "
              f"tcgen05 vs SIMT {rel_rms_err(out[1][i], out[0][i]):.3e}")
    for i in range(3):
        assert rel_rms_err(out[0][i], ref[i]) < 2e-5      # fp32 SIMT route: summation-order noise only
        assert rel_rms_err(out[1][i], ref[i]) < 1e-4      # tensor-pipe route: the stated block tolerance


@pytest.mark.parametrize("name", ["aishell3_long", "v3_tx128"])
def test_text_encoder_with_tensor_core_attention(name):
    """attention_tensor_cores = 1: QK^T and PV of the text encoder's attention run as f16-split tcgen05 MMAs
    (attn_tc.cu, 64 <= Tx <= 128); the encoder outputs must stay inside the block tolerance against the fixture."""
    import wetts_b200
    hps, sd, g, t = load_case(name)
    net = wetts_b200.build_model(hps, int(g["n_vocab"]), int(g["n_speakers"]), sd, "cuda")
    net.set_option("attention_tensor_cores", 0)
    h0, m0, l0, _ = net.enc_p(t["x"], t["x_lengths"])
    n0 = net.launch_count()
    net.set_option("attention_tensor_cores", 1)
    h1, m1, l1, _ = net.enc_p(t["x"], t["x_lengths"])
    torch.cuda.synchronize()
    net.check_faults()
    for nm, a, ref in (("h", h1, t["h"]), ("m", m1, t["m_p_tx"]), ("logs", l1, t["logs_p_tx"])):
        e = rel_rms_err(a.cpu(), ref)
        print(f"{name}: text encoder with tensor-pipe attention {nm} {e:.3e}")
        assert e < BLOCK_TOL
    assert not torch.equal(h0, h1), "the option must change the arithmetic route"
    assert net.launch_count() > n0
