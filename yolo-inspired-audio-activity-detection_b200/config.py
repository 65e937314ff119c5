"""config/config.yaml surface (reference: config/config.yaml:1-98, train.py:30-34)."""
from __future__ import annotations

import copy
import os
from typing import Any, Dict

import yaml

DEFAULT_CONFIG_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config", "config.yaml")


def load_config(path: str = DEFAULT_CONFIG_PATH) -> Dict[str, Any]:
    with open(path, "r") as f:
        return yaml.safe_load(f)


def default_config() -> Dict[str, Any]:
    return copy.deepcopy(load_config(DEFAULT_CONFIG_PATH))
