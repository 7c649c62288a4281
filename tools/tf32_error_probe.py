"""CPU probe: how far does a 3xTF32 evaluation of the text encoder's convolutions move m / logs / z_p on a
fixture?  (Used to judge the z_p tolerance of tests/test_zz_widecases_gpu.py; test infrastructure only.)

Every F.conv1d inside oracle.vits_oracle.text_encoder is replaced by hi*hi + hi*lo + lo*hi with operands rounded
to TF32 (round-to-nearest-away on 13 dropped mantissa bits), accumulated in fp64 and rounded to fp32 once -- a
slightly optimistic model of the tensor-core path (TMEM accumulates in fp32)."""
import sys
import os
import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vits_oracle as O  # noqa: E402
from tests.golden_util import load_case, rel_rms_err  # noqa: E402


def tf32(x):
    u = x.contiguous().view(torch.int32)
    return ((u + 0x1000) & ~0x1FFF).view(torch.float32)


def conv3xtf32(x, w, b=None, **kw):
    xh = tf32(x); xl = tf32(x - xh)
    wh = tf32(w); wl = tf32(w - wh)
    d = torch.float64
    y = (_conv(xh.to(d), wh.to(d), None, **kw) + _conv(xh.to(d), wl.to(d), None, **kw) + _conv(xl.to(d), wh.to(d), None, **kw))
    if b is not None:
        y = y + b.to(d)[None, :, None]
    return y.float()


_conv = F.conv1d


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "aishell3_long"
    hps, sd, g, t = load_case(name)
    w = O.fold_weight_norm(sd)
    ref = O.text_encoder(w, hps.model, t["x"], t["x_lengths"])
    O.F.conv1d = conv3xtf32
    try:
        got = O.text_encoder(w, hps.model, t["x"], t["x_lengths"])
    finally:
        O.F.conv1d = _conv
    names = ["h", "m", "logs"]
    for n, a, b in zip(names, got, ref):
        print(f"{name}: {n:5s} max|d| = {float((a - b).abs().max()):.3e}   rel to rms = {rel_rms_err(a, b):.3e}")
    # effect on z_p through exp(logs) with the fixture's noise
    dm, dl = (got[1] - ref[1]), (got[2] - ref[2])
    print(f"{name}: max |d logs| * 11 (noise*exp(logs)*0.667 amplification) = {float(dl.abs().max()) * 11:.3e} "
          f"vs block tolerance 1e-4 * rms(z_p) ~ 1e-4")


if __name__ == "__main__":
    main()
