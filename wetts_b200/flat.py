"""Flat weights file for the C++ runtime shim (runtime/vits_model.cc): the engine configuration
(`wetts_vits_config`, byte for byte as in include/wetts_b200.h) followed by every checkpoint tensor under its
reference state-dict key.  A C++ host cannot unpickle a `.pth`; this is the hand-over format.

    magic "WETTSB2\\0" | u32 version=1 | u32 sizeof(config) | config | i32 sampling_rate | u32 n_tensors
    per tensor: u16 name_len | name | u8 ndim | i64 dims[ndim] | f32 data[numel]

    python -m wetts_b200.flat --config cfg.json --checkpoint G_x.pth --n_vocab N --n_speakers S --out model.wb2
"""
import struct

import numpy as np
import torch

MAGIC = b"WETTSB2\x00"


def write_flat(path, hps, n_vocab, n_speakers, state_dict):
    """Writes the config and `state_dict` (reference keys; `enc_q.*` skipped, as the engine ignores it)."""
    from .models import SynthesizerTrn
    net = SynthesizerTrn(n_vocab, hps.data.filter_length // 2 + 1, hps.train.segment_size // hps.data.hop_length,
                         n_speakers=n_speakers, **hps.model)
    blob = bytes(memoryview(net._engine.cfg))
    tensors = [(k, v) for k, v in state_dict.items() if torch.is_tensor(v) and not k.startswith("enc_q.")]
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<II", 1, len(blob)))
        f.write(blob)
        f.write(struct.pack("<iI", int(hps.data.sampling_rate), len(tensors)))
        for name, t in tensors:
            a = np.ascontiguousarray(t.detach().cpu().to(torch.float32).numpy())
            nb = name.encode()
            f.write(struct.pack("<H", len(nb)))
            f.write(nb)
            f.write(struct.pack("<B", a.ndim))
            f.write(struct.pack("<%dq" % a.ndim, *a.shape))
            f.write(a.tobytes())
    return len(tensors)


def main():
    import argparse
    from .hparams import get_hparams_from_file
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True)
    ap.add_argument("--checkpoint", required=True)
    ap.add_argument("--n_vocab", type=int, required=True)
    ap.add_argument("--n_speakers", type=int, required=True)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    hps = get_hparams_from_file(a.config)
    ck = torch.load(a.checkpoint, map_location="cpu")
    sd = ck["model"] if "model" in ck else ck
    n = write_flat(a.out, hps, a.n_vocab, a.n_speakers, sd)
    print(f"wrote {a.out}: {n} tensors")


if __name__ == "__main__":
    main()
