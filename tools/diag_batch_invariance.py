"""Does an utterance's waveform depend on which other utterances share its batch, or on what ran before in the same
workspace?  (diagnostic for the NCCL sharded-vs-unsharded test; one GPU)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, wetts_b200
from wetts_b200 import synth
from wetts_b200.hparams import builtin_config

dev = torch.device("cuda", 0)
hps = builtin_config("multilingual_v3")
n_vocab, n_spk = 64, 2
sd = synth.make_state_dict(hps.model, n_vocab, n_spk, seed=1234)
net = wetts_b200.build_model(hps, n_vocab, n_spk, sd, dev)
gen = torch.Generator().manual_seed(3)
B, Tx = 7, 24
lens = torch.tensor([24, 9, 17, 13, 24, 5, 20])
x = torch.randint(1, n_vocab, (B, Tx), generator=gen) * (torch.arange(Tx)[None, :] < lens[:, None])
sid = torch.randint(0, n_spk, (B,), generator=gen)
dur = torch.randint(3, 7, (B, 1, Tx), generator=gen).float() * (torch.arange(Tx)[None, None, :] < lens[:, None, None])
Tmax = int(dur.sum(-1).max())
noise_z = torch.randn(B, 192, Tmax, generator=gen)
kw = dict(noise_scale=0.667, length_scale=1.0, noise_scale_w=0.8, return_attn=False)


def run(idx):
    idx = torch.tensor(idx)
    o, _, ym, (z, z_p, m_p, logs_p) = net.infer(x[idx], lens[idx], sid[idx], noise_z=noise_z[idx], durations=dur[idx], **kw)
    n = (ym.sum((1, 2)) * 256).long().tolist()
    return [o[i, 0, :n[i]].clone() for i in range(len(n))], z.clone(), z_p.clone(), ym.clone()


def cmp(tag, a, b):
    d = float((a - b).abs().max()) if a.shape == b.shape else float("nan")
    print(f"  {tag}: shapes {tuple(a.shape)} {tuple(b.shape)} max|diff| {d:.3e}")


full1, zf, zpf, ymf = run(list(range(B)))
print("frames per utterance:", [int(v) for v in ymf.sum((1, 2)).tolist()], "Tmax", Tmax)
for sub in ([0, 4, 2, 5], [0, 1, 2, 3], [0], [4, 0], [1, 0]):
    outs, z, z_p, ym = run(sub)
    print("subset", sub)
    for k, i in enumerate(sub):
        cmp(f"utt {i} waveform vs full", outs[k], full1[i])
        n = int(ym[k].sum())
        cmp(f"utt {i} z_p", z_p[k, :, :n], zpf[i, :, :n])
        cmp(f"utt {i} z", z[k, :, :n], zf[i, :, :n])
full2, _, _, _ = run(list(range(B)))
print("full again (after the subsets, same workspace)")
for i in range(B):
    cmp(f"utt {i}", full2[i], full1[i])
for fmt, name in ((32, "tensor_format 32"), (16, "tensor_cores 0")):
    if name == "tensor_cores 0":
        net.set_option("tensor_cores", 0)
    else:
        net.set_option("tensor_format", fmt)
    f, _, _, _ = run(list(range(B)))
    s, _, _, _ = run([0, 4, 2, 5])
    print(name)
    cmp("utt 0 subset vs full", s[0], f[0])
    net.set_option("tensor_format", 16)
