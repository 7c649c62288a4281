"""Config handling that accepts the reference's JSON recipe files unchanged.

Mirrors the interface of the reference's recursive attribute bag
(reference: wetts/vits/utils/task.py:250-255 `get_hparams_from_file`,
task.py:273-303 `HParams`): attribute access, `[]` access, `in`, `len`,
`keys/items/values`, nested dicts become nested bags.  `**hps.model` splats
because the bag implements the mapping protocol (`keys` + `__getitem__`).
"""
import json
import os

_CONFIG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs")


class HParams:
    def __init__(self, **entries):
        for name, value in entries.items():
            self[name] = HParams(**value) if isinstance(value, dict) else value

    # mapping protocol -------------------------------------------------
    def keys(self):
        return vars(self).keys()

    def items(self):
        return vars(self).items()

    def values(self):
        return vars(self).values()

    def __len__(self):
        return len(vars(self))

    def __getitem__(self, name):
        return getattr(self, name)

    def __setitem__(self, name, value):
        setattr(self, name, value)

    def __contains__(self, name):
        return name in vars(self)

    def __repr__(self):
        return repr(vars(self))

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, HParams) else v) for k, v in self.items()}


def get_hparams_from_file(config_path):
    with open(config_path, "r") as f:
        return HParams(**json.load(f))


def builtin_config(name):
    """Load one of the recipe shapes shipped with this package, e.g.
    'multilingual_v3', 'baker_v1', 'aishell3_v1' (values equal the reference's
    examples/*/configs/{v1,v2,v3}.json)."""
    return get_hparams_from_file(os.path.join(_CONFIG_DIR, name + ".json"))
