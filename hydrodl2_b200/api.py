"""Model discovery and loading — same behaviour as ``hydrodl2/api/methods.py``.

``load_model(model, ver_name)`` (methods.py:78-139) returns the *uninstantiated*
class: the name is snake-cased, the directory is its first token, the class is
looked up by ``ver_name`` and, on a miss, falls back to the alphabetically first
class of the module with a warning — exactly the reference's rule, so callers
written against hydrodl2 keep working (including its quirks).
"""

from __future__ import annotations

import importlib
import logging
import os
import re

from torch.nn import Module

log = logging.getLogger('hydrodl2_b200')

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))


def _scan(kind: str) -> dict[str, list[str]]:
    root = os.path.join(_PKG_DIR, kind)
    found: dict[str, list[str]] = {}
    if not os.path.isdir(root):
        return found
    for d in sorted(os.listdir(root)):
        p = os.path.join(root, d)
        if not os.path.isdir(p) or d.startswith('_'):
            continue
        files = sorted(f[:-3] for f in os.listdir(p) if f.endswith('.py') and not f.startswith('_'))
        found[d] = files
    return found


def available_models() -> dict[str, list[str]]:
    """methods.py:18-35 — {model directory: [model files]}."""
    return _scan('models')


def _list_available_models() -> list[str]:
    """methods.py:38-55."""
    return [f for files in _scan('models').values() for f in files]


def available_modules() -> dict[str, list[str]]:
    """methods.py:58-75 (the reference ships only placeholders here)."""
    return _scan('modules')


def load_model(model: str, ver_name: str = None) -> Module:
    """methods.py:78-139."""
    if ver_name is None:
        ver_name = model
    model = re.sub(r'([a-z])([A-Z])', r'\1_\2', model).lower()
    model_dir = model.split('_')[0].lower()
    if not os.path.exists(os.path.join(_PKG_DIR, 'models', model_dir, f'{model}.py')):
        raise ImportError(f"Model '{model}' not found.")
    try:
        module = importlib.import_module(f'{__package__}.models.{model_dir}.{model}')
    except ImportError as e:
        raise ImportError(f"Model '{model}' not found.") from e
    try:
        cls = getattr(module, ver_name)
    except AttributeError as e:
        classes = [a for a in dir(module) if isinstance(getattr(module, a), type) and a != 'Any']
        if not classes:
            raise ImportError(f"Model version '{model}' not found.") from e
        log.warning(
            f"Model class '{ver_name}' not found in module '{module.__file__}'. "
            f"Falling back to the first available: '{classes[0]}'."
        )
        cls = getattr(module, classes[0])
    return cls


def load_module():
    """methods.py:142-144."""
    raise NotImplementedError("This function is not yet implemented.")
