"""Size-independent properties at BASELINE.json's full size (multilingual v3, 256 utterances x 128 phonemes):
the oracle cannot run this shape in seconds, so the CUDA path is checked through properties the domain offers.

* batch-composition invariance: utterances are independent, so the waveform of an utterance must
  be BIT-identical whether it is synthesised inside the full batch or inside an 8-utterance sub-batch (same
  injected noise, same teacher-forced durations -> same Ty, hence the same padded tail, finding 9);
* determinism: the same call twice gives bit-identical output (no atomics / order-dependent accumulation);
* length bookkeeping: y_lengths = sum of the forced durations, audio length = Ty * 256, y_mask row sums = y_lengths;
* the sub-batch is also checked against the CPU oracle (stated fp32 tolerance), which anchors the full batch."""
import pytest
import torch

from tests.golden_util import rel_rms_err

pytestmark = pytest.mark.gpu


def test_full_size_batch_invariance_determinism_and_oracle_anchor():
    import wetts_b200
    from oracle import vits_oracle as O
    from wetts_b200 import synth
    from wetts_b200.hparams import builtin_config
    hps = builtin_config("multilingual_v3")
    n_vocab, n_spk, B, Tx = 256, 2, 256, 128
    sd = synth.make_state_dict(hps.model, n_vocab, n_spk, seed=hps.train.seed)
    net = wetts_b200.build_model(hps, n_vocab, n_spk, sd, "cuda")
    gen = torch.Generator().manual_seed(2024)
    x = torch.randint(0, n_vocab, (B, Tx), generator=gen)
    lens = torch.full((B,), Tx, dtype=torch.long)
    lens[::7] = torch.randint(64, Tx, (len(lens[::7]),), generator=gen)          # ragged text lengths
    sid = torch.randint(0, n_spk, (B,), generator=gen)
    # teacher-forced integer durations (1..8 frames per phoneme inside the text, 0 outside); every utterance is
    # then padded with phoneme 0's duration so that all utterances have the same Ty (same padded tail everywhere)
    dur = torch.randint(1, 9, (B, 1, Tx), generator=gen).float()
    dur = dur * (torch.arange(Tx)[None, None, :] < lens[:, None, None])
    tot = dur.sum(-1, keepdim=True)
    Ty = int(tot.max())
    dur[:, :, 0] += (Ty - tot)[:, :, 0]
    assert torch.all(dur.sum(-1) == Ty)
    noise_z = torch.randn(B, hps.model.inter_channels, Ty, generator=gen)

    def run(idx):
        o, _, y_mask, (z, *_r) = net.infer(x[idx], lens[idx], sid[idx], 0.667, 1.0, 0.8, noise_z=noise_z[idx],
                                           durations=dur[idx], return_attn=False)
        torch.cuda.synchronize()
        return o, y_mask, z, net.last_y_lengths.clone()

    full = torch.arange(B)
    o1, y_mask, z1, yl = run(full)
    assert o1.shape == (B, 1, Ty * 256)
    assert torch.equal(yl.cpu(), torch.full((B,), Ty, dtype=torch.long))
    assert torch.equal(y_mask[:, 0].sum(-1).long().cpu(), yl.cpu())
    o2, _, z2, _ = run(full)
    assert torch.equal(o1, o2) and torch.equal(z1, z2), "same call twice must be bit-identical"
    sub = torch.tensor([0, 7, 31, 100, 128, 200, 254, 255])
    o3, _, z3, _ = run(sub)
    assert torch.equal(z3, z1[sub.cuda()]), "flow output must not depend on the batch composition"
    assert torch.equal(o3, o1[sub.cuda()]), "waveform must not depend on the batch composition"
    # anchor: two of the sub-batch utterances against the CPU oracle
    two = sub[:2]
    r = O.infer(sd, hps.model, x[two], lens[two], sid[two], 0.667, 1.0, 0.8, noise_z=noise_z[two], durations=dur[two])
    assert torch.equal(r["y_lengths"], torch.full((2,), Ty, dtype=torch.long))
    assert rel_rms_err(o1[two.cuda()].cpu(), r["o"]) < 1e-3
    assert rel_rms_err(z1[two.cuda()].cpu(), r["z"]) < 3e-4
