"""Pins oracle/vits_oracle.py against fixtures produced by the REAL reference
(oracle/gen_golden.py).  CPU only."""
import pytest
import torch

from oracle import vits_oracle as O
from tests.golden_util import CASES, VITS2_CASES, WIDE_CASES, load_case, rel_rms_err

TOL = 2e-5  # relative to rms; both sides are fp32 CPU, differences are summation order only


@pytest.mark.parametrize("name", CASES + WIDE_CASES + VITS2_CASES)
def test_oracle_matches_reference_fixture(name):
    hps, sd, g, t = load_case(name)
    ns, ls, nsw = [float(v) for v in g["scales"]]
    torch.set_num_threads(4)
    r = O.infer(sd, hps.model, t["x"], t["x_lengths"], t["sid"], noise_scale=ns, length_scale=ls,
                noise_scale_w=nsw, noise_w=t["noise_w"], noise_z=t["noise_z"], durations=t["w_ceil"])
    assert rel_rms_err(r["h"], t["h"]) < TOL
    assert rel_rms_err(r["m_p_tx"], t["m_p_tx"]) < TOL
    assert rel_rms_err(r["logs_p_tx"], t["logs_p_tx"]) < TOL
    assert rel_rms_err(r["logw"], t["logw"]) < 1e-4
    assert torch.equal(r["y_lengths"], t["y_lengths"])
    assert rel_rms_err(r["z_p"], t["z_p"]) < TOL
    assert rel_rms_err(r["z"], t["z"]) < TOL
    assert rel_rms_err(r["o"], t["o"]) < 1e-4
    # frame->phoneme map equals the argmax of the reference's one-hot attn on valid frames
    idx = r["attn_idx"]
    valid = idx >= 0
    assert torch.equal(idx[valid].int(), t["attn_argmax"][valid])
    assert torch.equal(valid.float(), t["attn_rowsum"])


@pytest.mark.parametrize("name", ["v3_ragged", "v1_ragged", "baker_v1_cli"])
def test_oracle_own_durations_match_reference(name):
    """Without teacher forcing the oracle's ceil(exp(logw)) lands on the same integers."""
    hps, sd, g, t = load_case(name)
    ns, ls, nsw = [float(v) for v in g["scales"]]
    r = O.infer(sd, hps.model, t["x"], t["x_lengths"], t["sid"], noise_scale=ns, length_scale=ls,
                noise_scale_w=nsw, noise_w=t["noise_w"], noise_z=t["noise_z"])
    assert torch.equal(r["w_ceil"], t["w_ceil"])
