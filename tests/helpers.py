"""Shared helpers for the test-suite (oracle side)."""
from __future__ import annotations

import numpy as np


def unhex(s):
    return None if s is None else float.fromhex(s)


def bits_from_hex(h: str, shape) -> np.ndarray:
    return np.frombuffer(bytes.fromhex(h), dtype=np.uint16).reshape(shape).copy()
