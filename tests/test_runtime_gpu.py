"""C++ runtime shim on the GPU: `vits_main` (runtime/) against the Python adapters over the same engine.
The decoder and the chunked streaming decode must be BIT-identical to `export_decoder_forward` /
`split_to_chunks` + `depadding` (same kernels, same chunking rules, vits_model.cc:96-153); the full path must
produce the number of samples the deterministic duration predictor implies."""
import subprocess

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup(tmp_path_factory):
    import wetts_b200
    from wetts_b200 import build as _build
    from wetts_b200 import synth
    from wetts_b200.flat import write_flat
    from wetts_b200.hparams import builtin_config
    hps = builtin_config("multilingual_v3")
    sd = synth.make_state_dict(hps.model, 80, 2, seed=11)
    d = tmp_path_factory.mktemp("rt")
    path = str(d / "model.wb2")
    write_flat(path, hps, 80, 2, sd)
    net = wetts_b200.build_model(hps, 80, 2, sd, "cuda")
    return _build.build_runtime(), path, net, d


def _run(exe, args):
    r = subprocess.run([exe] + args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_decoder_and_streaming_are_bit_identical_to_the_python_adapters(setup):
    from wetts_b200.session import depadding, split_to_chunks
    exe, path, net, d = setup
    gen = torch.Generator().manual_seed(5)
    x = torch.randint(0, 80, (1, 60), generator=gen)
    lens, sid = torch.tensor([60]), torch.tensor([1])
    scales = torch.tensor([[0.667, 2.0, 0.8]])
    z = net.export_encoder_forward(x, lens, scales, sid)             # [1, L, 192]
    L = z.shape[1]
    assert L > 100                                                   # several 40-frame chunks
    zf = str(d / "z.f32")
    z[0].cpu().numpy().astype("<f4").tofile(zf)
    ref = (net.export_decoder_forward(z, sid)[0, 0] * 32767.0).cpu().numpy()
    out = str(d / "full.f32")
    _run(exe, ["--weights", path, "--decode_z", zf, "--sid", "1", "--f32", out, "--wav", str(d / "full.wav")])
    got = np.fromfile(out, dtype="<f4")
    assert got.shape == ref.shape == (L * 256,)
    assert np.array_equal(got, ref)
    # 16-bit wav: header + clipped samples
    wav = open(d / "full.wav", "rb").read()
    assert wav[:4] == b"RIFF" and wav[8:12] == b"WAVE" and len(wav) == 44 + 2 * L * 256
    # chunked streaming decode (chunk 40, pad 10)
    chunks = split_to_chunks(z, 40, 10)
    pieces = []
    for i, c in enumerate(chunks):
        a = net.export_decoder_forward(c.contiguous(), sid)[:, 0]
        pieces.append(depadding(a, len(chunks), i, 40, 10))
    ref_s = (torch.cat(pieces, dim=1)[0] * 32767.0).cpu().numpy()
    out_s = str(d / "stream.f32")
    _run(exe, ["--weights", path, "--decode_z", zf, "--sid", "1", "--stream", "--chunk", "40", "--pad", "10", "--f32", out_s])
    got_s = np.fromfile(out_s, dtype="<f4")
    assert got_s.shape == ref_s.shape == (L * 256,)
    assert np.array_equal(got_s, ref_s)


def test_full_path_matches_the_deterministic_duration_predictor(setup):
    exe, path, net, d = setup
    ids = [3, 17, 42, 8, 8, 61, 29, 5, 77, 13, 40, 2]
    x = torch.tensor([ids])
    o, *_ = net.infer(x, torch.tensor([len(ids)]), torch.tensor([0]), 0.667, 1.0, 0.8, return_attn=False)
    frames = int(net.last_y_lengths[0])
    out = str(d / "tts.f32")
    txt = _run(exe, ["--weights", path, "--phonemes", " ".join(map(str, ids)), "--sid", "0", "--f32", out])
    got = np.fromfile(out, dtype="<f4")
    assert f"frames {frames} " in txt and got.shape == (frames * 256,)       # v3: DurationPredictor, no noise in durations
    assert np.isfinite(got).all()
    ref_rms = float((o[0, 0] * 32767.0).pow(2).mean().sqrt())
    assert 0.5 * ref_rms < float(np.sqrt((got.astype(np.float64) ** 2).mean())) < 2.0 * ref_rms   # different noise draw
