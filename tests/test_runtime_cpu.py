"""C++ runtime shim (runtime/vits_model.{h,cc}) without a GPU: it builds, the flat weights
file written by wetts_b200/flat.py has the documented layout, and the CLI fails loudly (no CPU path) when there
is no CUDA device or the file is damaged."""
import os
import struct
import subprocess

import numpy as np
import pytest
import torch

from wetts_b200 import build as _build
from wetts_b200 import synth
from wetts_b200.flat import MAGIC, write_flat
from wetts_b200.hparams import builtin_config


@pytest.fixture(scope="module")
def exe():
    _build.build()
    return _build.build_runtime()


@pytest.fixture(scope="module")
def flat_file(tmp_path_factory):
    hps = builtin_config("multilingual_v3")
    sd = synth.make_state_dict(hps.model, 40, 2, seed=3)
    path = str(tmp_path_factory.mktemp("flat") / "model.wb2")
    n = write_flat(path, hps, 40, 2, sd)
    return path, sd, n, hps


def test_flat_file_layout(flat_file):
    path, sd, n, hps = flat_file
    raw = open(path, "rb").read()
    assert raw[:8] == MAGIC
    version, cfg_bytes = struct.unpack_from("<II", raw, 8)
    assert version == 1
    off = 16 + cfg_bytes
    n_vocab, n_speakers, inter = struct.unpack_from("<iii", raw, 16)          # first fields of wetts_vits_config
    assert (n_vocab, n_speakers, inter) == (40, 2, hps.model["inter_channels"])
    sr, n_tensors = struct.unpack_from("<iI", raw, off)
    off += 8
    assert sr == hps.data.sampling_rate and n_tensors == n
    seen = {}
    for _ in range(n_tensors):
        (ln,) = struct.unpack_from("<H", raw, off); off += 2
        name = raw[off:off + ln].decode(); off += ln
        (nd,) = struct.unpack_from("<B", raw, off); off += 1
        dims = struct.unpack_from("<%dq" % nd, raw, off); off += 8 * nd
        numel = int(np.prod(dims)) if nd else 1
        seen[name] = np.frombuffer(raw, dtype="<f4", count=numel, offset=off).reshape(dims); off += 4 * numel
    assert off == len(raw)
    assert not any(k.startswith("enc_q.") for k in seen)
    for k in ("dec.ups.0.weight_v", "dec.ups.0.weight_g", "enc_p.emb.weight", "flow.flows.0.post.weight"):
        assert np.array_equal(seen[k], sd[k].numpy())


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_cli_fails_loudly_without_gpu(exe, flat_file):
    path = flat_file[0]
    r = subprocess.run([exe, "--weights", path, "--phonemes", "1 2 3"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "vits_main:" in r.stderr      # cudaSetDevice / engine creation error, no fallback


def test_cli_rejects_damaged_files(exe, flat_file, tmp_path):
    raw = open(flat_file[0], "rb").read()
    bad_magic = tmp_path / "bad_magic.wb2"
    bad_magic.write_bytes(b"NOTAFILE" + raw[8:])
    truncated = tmp_path / "truncated.wb2"
    truncated.write_bytes(raw[: len(raw) // 2])
    for p, msg in ((bad_magic, "not a wetts_b200 flat weights file"), (truncated, "truncated")):
        r = subprocess.run([exe, "--weights", str(p), "--phonemes", "1 2 3"], capture_output=True, text=True, timeout=120)
        assert r.returncode == 1 and msg in r.stderr, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 64 and "usage:" in r.stderr
