"""Shared helpers for tests that use the reference-generated fixtures in tests/golden/."""
import os

import numpy as np
import torch

from wetts_b200 import synth
from wetts_b200.hparams import builtin_config

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["v3_ragged", "v3_single", "v1_ragged", "v2_short"]
# BASELINE.json configs[4] family (AISHELL-3 v1, 218 speakers, 128/80 phonemes), configs[0] (Baker v1, the CLI
# utterance, batch 1, CLI scales) and the benchmark's own route (multilingual v3 at Tx = 128, ragged 128/97/64: every
# text-encoder / duration-predictor conv takes the tcgen05 kernel): pinned on the CPU with the other fixtures; on the
# GPU they gate in tests/test_zz_widecases_gpu.py.
WIDE_CASES = ["aishell3_long", "baker_v1_cli", "v3_tx128"]
# the vits2_vocos_v1 recipe (Vocos iSTFT decoder, VITS2 'pre_conv' transformer flows, SDP)
VITS2_CASES = ["vits2_vocos_short"]


def load_case(name):
    d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: d[k] for k in d.files}
    hps = builtin_config(str(g["config"]))
    sd = synth.make_state_dict(hps.model, int(g["n_vocab"]), int(g["n_speakers"]), seed=int(g["ckpt_seed"]))
    fp = synth.fingerprint(sd)
    assert abs(fp - float(g["fingerprint"])) <= 1e-9 * abs(fp), "synthetic checkpoint differs from the fixture's"
    t = {k: torch.from_numpy(v) for k, v in g.items() if v.dtype.kind in "fi" and v.ndim > 0}
    return hps, sd, g, t


def rel_rms_err(a, b):
    """max |a-b| relative to rms(b) -- the tolerance form SURVEY.md §8(c) states."""
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.pow(2).mean().sqrt().clamp_min(1e-30))
