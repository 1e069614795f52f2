"""JSON / pickle persistence for the dataclass configs (reference: wtracker/utils/config_base.py:14-88),
without the tkinter file dialogs: a missing path is an error here."""

from __future__ import annotations

import json
import pickle
from dataclasses import dataclass, is_dataclass
from typing import Any


def _plain(obj: Any) -> Any:
    if is_dataclass(obj):
        return {k: _plain(v) for k, v in obj.__dict__.items()}
    if isinstance(obj, (list, tuple)):
        return [_plain(v) for v in obj]
    if isinstance(obj, dict):
        return {k: _plain(v) for k, v in obj.items()}
    return obj


@dataclass
class ConfigBase:
    @classmethod
    def load_json(cls, path: str):
        if path is None:
            raise ValueError("a path is required (wtracker_b200 has no file dialogs)")
        with open(path, "r") as f:
            data = json.load(f)
        obj = cls.__new__(cls)
        obj.__dict__.update(data)
        return obj

    def save_json(self, path: str) -> None:
        if path is None:
            raise ValueError("a path is required (wtracker_b200 has no file dialogs)")
        with open(path, "w") as f:
            json.dump(_plain(self), f, indent=4)

    @classmethod
    def load_pickle(cls, path: str):
        with open(path, "rb") as f:
            return pickle.load(f)

    def save_pickle(self, path: str) -> None:
        with open(path, "wb") as f:
            pickle.dump(self, f)
