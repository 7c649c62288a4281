"""CPU probe: error of a split-operand tensor-core evaluation of the generator / flow against fp32 (fp64-checked).

Two splittings of an fp32 operand x into tensor-core operands are compared on the reference-generated fixtures:
  tf32x3 : hi = tf32(x), lo = tf32(x - hi)                       (kind::tf32, K = 8 per MMA)
  f16x3  : hi = f16(x),  lo' = f16((x - hi) * 2^11)              (kind::f16,  K = 16 per MMA; small terms accumulate
           in their own accumulator columns and are scaled by 2^-11 in the epilogue)
Both keep 22 significand bits (tf32 and fp16 both carry 11); fp16 adds a range limit (|x| < 65504) and a subnormal
floor (absolute error <= 2^-25 for |x| < 2^-14).  Products hi*hi + hi*lo + lo*hi, fp64 accumulation (optimistic for
both in the same way).  Test infrastructure only."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vits_oracle as O  # noqa: E402
from tests.golden_util import load_case, rel_rms_err  # noqa: E402

_conv, _convT = F.conv1d, F.conv_transpose1d
MAXABS = {"v": 0.0}


def tf32(x):
    u = x.contiguous().view(torch.int32)
    return ((u + 0x1000) & ~0x1FFF).view(torch.float32)


def split_tf32(x):
    hi = tf32(x)
    return hi, tf32(x - hi), 1.0


def split_f16(x):
    hi = x.half().float()
    lo = ((x - hi) * 2048.0).half().float()
    return hi, lo, 1.0 / 2048.0


def make(split):
    def conv(x, w, b=None, **kw):
        MAXABS["v"] = max(MAXABS["v"], float(x.abs().max()))
        xh, xl, s = split(x)
        wh, wl, s2 = split(w)
        d = torch.float64
        main = _conv(xh.to(d), wh.to(d), None, **kw)
        small = _conv(xh.to(d), wl.to(d), None, **kw) + _conv(xl.to(d), wh.to(d), None, **kw)
        y = main.float() + small.float() * s          # two fp32 accumulators, combined in the epilogue
        if b is not None:
            y = y + b[None, :, None]
        return y

    def convT(x, w, b=None, **kw):
        xh, xl, s = split(x)
        wh, wl, s2 = split(w)
        d = torch.float64
        main = _convT(xh.to(d), wh.to(d), None, **kw)
        small = _convT(xh.to(d), wl.to(d), None, **kw) + _convT(xl.to(d), wh.to(d), None, **kw)
        y = main.float() + small.float() * s
        if b is not None:
            y = y + b[None, :, None]
        return y
    return conv, convT


def main():
    for name in sys.argv[1:] or ["v3_ragged", "v1_ragged"]:
        hps, sd, g, t = load_case(name)
        w = O.fold_weight_norm(sd)
        gv = w["emb_g.weight"][t["sid"]][:, :, None] if int(g["n_speakers"]) > 0 else None
        Ty = t["z"].shape[2]
        ym = (torch.arange(Ty)[None, :] < t["y_lengths"][:, None]).float()[:, None]
        ref_o = O.generator(w, hps.model, t["z"] * ym, gv)
        ref_z = O.flow_reverse(w, hps.model, t["z_p"], ym, gv)
        print(f"{name}: fp32 oracle vs fixture: o {rel_rms_err(ref_o, t['o']):.2e}  z {rel_rms_err(ref_z, t['z']):.2e}")
        for label, split in (("tf32x3", split_tf32), ("f16x3 ", split_f16)):
            conv, convT = make(split)
            O.F.conv1d, O.F.conv_transpose1d = conv, convT
            MAXABS["v"] = 0.0
            try:
                o = O.generator(w, hps.model, t["z"] * ym, gv)
                z = O.flow_reverse(w, hps.model, t["z_p"], ym, gv)
            finally:
                O.F.conv1d, O.F.conv_transpose1d = _conv, _convT
            print(f"{name}: {label} generator err/rms {rel_rms_err(o, t['o']):.3e}   flow err/rms {rel_rms_err(z, t['z']):.3e}"
                  f"   max |activation| seen {MAXABS['v']:.1f}")


if __name__ == "__main__":
    main()
