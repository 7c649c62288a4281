"""L2 adapters on the GPU: ONNX-session contract, streaming chunk decode, int16 output stage,
tensor-core vs SIMT path agreement, error behaviour of the C ABI."""
import numpy as np
import pytest
import torch

import wetts_b200
from wetts_b200 import _lib, synth
from wetts_b200.hparams import builtin_config
from wetts_b200.session import InferenceSession, StreamingVits, depadding, split_to_chunks, to_int16
from tests.golden_util import load_case, rel_rms_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def v3():
    hps = builtin_config("multilingual_v3")
    sd = synth.make_state_dict(hps.model, 80, 2, seed=11)
    return hps, wetts_b200.build_model(hps, 80, 2, sd, "cuda")


def test_session_contract_matches_infer(v3):
    hps, net = v3
    gen = torch.Generator().manual_seed(5)
    x = torch.randint(0, 80, (2, 24), generator=gen)
    lens = torch.tensor([24, 17])
    sid = torch.tensor([1, 0])
    feeds = {"input": x.numpy(), "input_lengths": lens.numpy(),
             "scales": np.array([[0.667, 2.8, 0.8]] * 2, dtype=np.float32), "sid": sid.numpy()}
    sess = InferenceSession(net, "full")
    assert [i.name for i in sess.get_inputs()] == ["input", "input_lengths", "scales", "sid"]
    torch.manual_seed(77)
    out = sess.run(None, feeds)[0]
    torch.manual_seed(77)
    o, _, y_mask, (z, *_rest) = net.infer(x, lens, sid, 0.667, 2.8, 0.8)
    assert out.shape == tuple(o.shape) and out.dtype == np.float32
    assert np.array_equal(out, o.cpu().numpy())
    # encoder graph -> z [B, L, 192] (masked, time-major); decoder graph vocodes it
    torch.manual_seed(77)
    zt = InferenceSession(net, "encoder").run(None, feeds)[0]
    assert zt.shape == (2, z.shape[2], 192)
    assert np.allclose(zt, (z * y_mask).transpose(1, 2).cpu().numpy())
    dec = InferenceSession(net, "decoder").run(None, {"z": zt, "sid": sid.numpy()})[0]
    g = net.emb_g(sid)[:, :, None]
    ref = net.dec(torch.from_numpy(zt).transpose(1, 2).cuda(), g=g)
    assert np.array_equal(dec, ref.cpu().numpy())
    with pytest.raises(ValueError):
        sess.run(None, {"input": x.numpy()})


def test_streaming_chunks_match_full_decode(v3):
    hps, net = v3
    torch.manual_seed(3)
    phon = list(range(1, 41))
    st = StreamingVits(net, chunk_size=40, pad_size=12, scales=(0.667, 2.8, 0.8))
    torch.manual_seed(9)
    st.set_input(phon, 1)
    L = sum(c.shape[1] for c in split_to_chunks(torch.cat(st.chunks[:1], 1), -1, 0)) if False else None
    pieces, done = [], False
    while not done:
        a, done = st.stream_decode()
        if a is not None:
            pieces.append(a)
    audio = torch.cat(pieces)
    torch.manual_seed(9)
    z = net.export_encoder_forward(torch.tensor([phon]), torch.tensor([40]), torch.tensor([[0.667, 2.8, 0.8]]),
                                   torch.tensor([1]))
    full = net.export_decoder_forward(z, torch.tensor([1]))[0, 0] * 32767.0
    assert audio.shape == full.shape
    # pad (12 frames) covers the generator's receptive field (+-10.1 frames)
    assert rel_rms_err(audio.cpu(), full.cpu()) < 1e-3


def test_chunk_helpers_match_reference_semantics():
    z = torch.arange(2 * 95 * 3, dtype=torch.float32).reshape(2, 95, 3)
    ch = split_to_chunks(z, 40, 10)
    assert [c.shape[1] for c in ch] == [50, 60, 25]
    assert torch.equal(ch[1], z[:, 30:90])
    a = torch.arange(60 * 256, dtype=torch.float32)[None]
    assert depadding(a[:, : 50 * 256], 3, 0, 40, 10).shape[1] == 40 * 256
    assert depadding(a, 3, 1, 40, 10).shape[1] == 40 * 256
    assert depadding(a[:, : 25 * 256], 3, 2, 40, 10).shape[1] == 15 * 256
    assert split_to_chunks(z, -1, 0)[0] is z


def test_int16_output_stage():
    a = torch.tensor([[[0.5, -0.25, 0.001]], [[0.01, 0.02, -0.04]]], device="cuda")
    s = to_int16(a, "scale").cpu()
    assert s.dtype == torch.int16 and s[0, 0, 0] == int(0.5 * 32767)
    p = to_int16(a, "peak").cpu()
    assert abs(int(p[0, 0, 0]) - int(32767 * 0.6)) <= 1 and abs(int(p[1, 0, 2]) + int(32767 * 0.6)) <= 1
    b = to_int16(a, "peak_batch").cpu()
    assert abs(int(b[0, 0, 0]) - int(32767 * 0.6)) <= 1 and abs(int(b[1, 0, 2])) < 3000


def test_tensor_core_and_simt_paths_agree():
    """Same inputs through the tcgen05 3xTF32 path and the fp32 SIMT path."""
    hps, sd, g, t = load_case("v3_ragged")
    lib = _lib.load()
    outs = []
    try:
        for tc in (0, 1):
            _lib.check(lib.wetts_set_option(b"tensor_cores", tc))
            net = wetts_b200.build_model(hps, int(g["n_vocab"]), int(g["n_speakers"]), sd, "cuda")
            ns, ls, nsw = [float(v) for v in g["scales"]]
            o, *_ = net.infer(t["x"], t["x_lengths"], t["sid"], ns, ls, nsw, noise_w=t["noise_w"],
                              noise_z=t["noise_z"], durations=t["w_ceil"])
            outs.append(o.cpu())
    finally:
        _lib.check(lib.wetts_set_option(b"tensor_cores", 1))
    assert rel_rms_err(outs[1], outs[0]) < 3e-4
    assert rel_rms_err(outs[0], t["o"]) < 1e-4 and rel_rms_err(outs[1], t["o"]) < 1e-3


def test_padded_tail_is_computed_like_the_reference():
    """Finding 9: valid samples near an utterance's end depend on the padded tail being run through
    the whole conv stack; compare the FULL padded output of a ragged batch with the fixture."""
    hps, sd, g, t = load_case("v3_ragged")
    net = wetts_b200.build_model(hps, int(g["n_vocab"]), int(g["n_speakers"]), sd, "cuda")
    ns, ls, nsw = [float(v) for v in g["scales"]]
    o, *_ = net.infer(t["x"], t["x_lengths"], t["sid"], ns, ls, nsw, noise_w=t["noise_w"], noise_z=t["noise_z"],
                      durations=t["w_ceil"])
    short = int(t["y_lengths"].argmin())
    n_valid = int(t["y_lengths"][short]) * 256
    tail = o[short, 0, n_valid:].cpu()
    assert tail.abs().max() > 0                       # the tail is not silence
    assert rel_rms_err(tail, t["o"][short, 0, n_valid:]) < 1e-3


def test_c_abi_error_reporting(v3):
    hps, net = v3
    lib = _lib.load()
    e = net._engine
    # undersized workspace -> non-zero status + message, no crash
    z = torch.zeros(1, 192, 8, device="cuda")
    out = torch.empty(1, 1, 8 * 256, device="cuda")
    ws = torch.empty(1024, dtype=torch.uint8, device="cuda")
    import ctypes as C
    rc = lib.wetts_generator_forward(e.handle, C.c_void_p(z.data_ptr()), None, None, 1, 8, C.c_void_p(out.data_ptr()),
                                     C.c_void_p(ws.data_ptr()), ws.numel(), None)
    assert rc != 0 and b"workspace too small" in lib.wetts_last_error()
    assert lib.wetts_set_option(b"no_such_option", 1) != 0
    # missing checkpoint tensor
    bad = wetts_b200.SynthesizerTrn(80, 513, 32, n_speakers=2, **hps.model)
    sd = synth.make_state_dict(hps.model, 80, 2, seed=11)
    del sd["dec.conv_post.weight"]
    bad.load_state_dict(sd)
    with pytest.raises(wetts_b200.WettsError, match="dec.conv_post.weight"):
        bad.to("cuda")


def test_int16_output_stage_matches_the_callers_formulas():
    """wetts_audio_to_int16 against the three caller formulas (inference.py:101-105 per utterance, gpu_triton
    model.py:150-151 batch-global, cli/model.py:60 plain scale).  Plain scale: exact.  Peak modes: the kernel
    evaluates 32767 / max(0.01, peak) * 0.6 in fp32 in the reference's operation order; a torch re-evaluation may
    round the gain one ulp differently (e.g. `scalar / tensor` is reciprocal-times-scalar in torch), which moves a
    truncated int16 sample by at most one step."""
    gen = torch.Generator().manual_seed(3)
    a = (torch.randn(5, 1, 4099, generator=gen) * torch.tensor([0.3, 0.02, 1.7, 0.004, 0.9])[:, None, None]).cuda()
    lens = torch.tensor([4099, 1000, 37, 4099, 2048])

    def ref(mode):
        x = a.float()
        full = x.new_tensor(32767.0)
        if mode == "scale":
            g = full
        elif mode == "peak_batch":
            g = torch.div(full, x.abs().max().clamp_min(0.01)) * 0.6
        else:
            flat = x.reshape(5, -1) * (torch.arange(4099, device="cuda")[None, :] < lens.cuda()[:, None])
            g = (torch.div(full, flat.abs().amax(dim=1).clamp_min(0.01)) * 0.6).reshape(5, 1, 1)
        return (x * g).clamp(-32767.0, 32767.0).to(torch.int16)

    assert torch.equal(to_int16(a, "scale"), ref("scale"))
    for mode, kw in (("peak_batch", {}), ("peak", {"lengths": lens})):
        got, want = to_int16(a, mode, **kw), ref(mode)
        assert got.dtype == torch.int16 and got.shape == want.shape
        assert int((got.int() - want.int()).abs().max()) <= 1
    # per-utterance gain: the loudest valid sample of every row lands on 0.6 full scale (or clips when the peak
    # lies outside the valid prefix)
    p = to_int16(a, "peak", lengths=lens).int().reshape(5, -1)
    for b in range(5):
        assert abs(int(p[b, : int(lens[b])].abs().max()) - int(32767 * 0.6)) <= 2
    assert to_int16(a[:, 0], "scale").shape == (5, 4099)
    with pytest.raises(wetts_b200.WettsError):
        to_int16(a.cpu(), "scale")
