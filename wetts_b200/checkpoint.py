"""Checkpoint loading in the reference's format.

Mirrors `load_checkpoint(checkpoint_path, model, optimizer=None)` of the reference
(wetts/vits/utils/task.py:31-56): a `torch.save`d dict {"model", "iteration", "optimizer",
"learning_rate"} whose "model" entry is a SynthesizerTrn state dict carrying
weight_g/weight_v pairs (inference.py does not remove weight-norm).  Folding and re-layout
happen on the device inside the engine (wetts_vits_finalize).
"""
import os

import torch


def load_checkpoint(checkpoint_path, model, optimizer=None, *, trust_pickle=False):
    """`trust_pickle=True` restores the reference's behaviour (full unpickling, task.py:33) for checkpoints
    that carry non-tensor objects; the default only materialises tensors and plain containers, so an
    untrusted .pth cannot execute code."""
    assert os.path.isfile(checkpoint_path), checkpoint_path
    ckpt = torch.load(checkpoint_path, map_location="cpu", weights_only=not trust_pickle)
    state = ckpt["model"] if isinstance(ckpt, dict) and "model" in ckpt else ckpt
    model.load_state_dict(state)
    return model, optimizer, ckpt.get("learning_rate", None), ckpt.get("iteration", 0)
