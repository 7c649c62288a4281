"""Recipe: make the UNMODIFIED reference implementation of the path available next to the oracle.

    python -m oracle.build_ref          # authoring container (needs /root/reference); idempotent

The reference's hot path is pure Python/PyTorch (`wetts/vits/model/*.py`, `wetts/vits/utils/*.py`), so "building"
it means placing those files, byte for byte, where an interpreter can import them: `oracle/_ref/wetts_vits/`.
`oracle/_ref/` is git-ignored (reference sources never enter this repository's history) but not gpurun-ignored,
so it travels to the GPU box with the snapshot, where `/root/reference` does not exist.  It is used for two
things only (test infrastructure, like everything under oracle/):
  * `bench.py --impl reference` / `cpu_baseline`: the CPU arm times the reference's own `SynthesizerTrn.infer`
    (kind "reference") instead of the oracle port when this directory is present;
  * `tests/test_oracle_golden.py` cross-checks the oracle against the live reference when available.
A manifest with the sha256 of every file is written so a reader can verify nothing was edited.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/wetts/vits"
DST = os.path.join(HERE, "_ref", "wetts_vits")
SUBDIRS = ("model", "utils")


def build(verbose=False):
    """Returns the destination directory if the reference is (now) available there, else None."""
    if not os.path.isdir(SRC):
        return DST if os.path.isfile(os.path.join(DST, "MANIFEST.json")) else None
    manifest = {}
    for sub in SUBDIRS:
        os.makedirs(os.path.join(DST, sub), exist_ok=True)
        for name in sorted(os.listdir(os.path.join(SRC, sub))):
            if not name.endswith(".py"):
                continue
            s, d = os.path.join(SRC, sub, name), os.path.join(DST, sub, name)
            data = open(s, "rb").read()
            if not os.path.exists(d) or open(d, "rb").read() != data:
                shutil.copyfile(s, d)
            manifest[f"{sub}/{name}"] = hashlib.sha256(data).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "files": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print(f"oracle/_ref/wetts_vits: {len(manifest)} files from {SRC}")
    return DST


if __name__ == "__main__":
    sys.exit(0 if build(verbose=True) else 1)
