"""Test infrastructure: the two pieces of the reference's plug-in machinery the on-the-fly path goes
through, restated so that the GPU box (which has no /root/reference) can exercise the plug-ins.

* ``Registry`` / ``build_from_cfg``  -- stand-ins for reference baseline/utils/registry.py:12-84
* ``runner_to_cuda``                 -- stand-in for reference baseline/engine/runner.py:125-152 (``Runner.to_cuda``)

tests/test_plugins.py holds both to the reference where it is mounted: the real ``registry.py`` is loaded by
path and must behave identically on the same calls, and ``runner_to_cuda`` is run side by side with the method cut
out of the reference's runner.py on the same batches.  Nothing in the product imports this module.
"""
import importlib.util
import inspect
import os
import sys

import numpy as np
import torch

REF_ROOT = "/root/reference"


def load_reference_registry():
    """The reference's own registry module, loaded by path (it needs only ``six``); None if absent."""
    path = os.path.join(REF_ROOT, "baseline", "utils", "registry.py")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("_ref_registry", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class Registry:
    """Behavioural stand-in for reference baseline/utils/registry.py:12-51 (name, get, register_module)."""

    def __init__(self, name):
        self.name = name
        self._items = {}

    def get(self, key):
        return self._items.get(key)

    def register_module(self, cls):
        if not inspect.isclass(cls):
            raise TypeError(f"module must be a class, but got {type(cls)}")
        if cls.__name__ in self._items:
            raise KeyError(f"{cls.__name__} is already registered in {self.name}")
        self._items[cls.__name__] = cls
        return cls


def build_from_cfg(cfg, registry, default_args=None):
    """Stand-in for reference baseline/utils/registry.py:54-84: the constructor's kwargs are the cfg dict minus
    'type'; default_args (the reference passes dict(cfg=cfg)) only fill what the cfg dict does not set."""
    if not isinstance(cfg, dict) or "type" not in cfg:
        raise AssertionError("cfg must be a dict with a 'type' key")
    kind = cfg["type"]
    if isinstance(kind, str):
        cls = registry.get(kind)
        if cls is None:
            raise KeyError(f"{kind} is not in the {registry.name} registry")
    elif inspect.isclass(kind):
        cls = kind
    else:
        raise TypeError(f"type must be a str or valid type, but got {type(kind)}")
    kwargs = {k: v for k, v in cfg.items() if k != "type"}
    for k, v in (default_args or {}).items():
        kwargs.setdefault(k, v)
    return cls(**kwargs)


def registries():
    """(module providing Registry/build_from_cfg, is_reference): the real one where mounted."""
    ref = load_reference_registry()
    if ref is not None:
        return ref, True
    return sys.modules[__name__], False


def _to_device(t):
    return t.cuda()


def runner_to_cuda(self, batch):
    """What reference baseline/engine/runner.py:125-152 (``Runner.to_cuda``) does to a collated batch, in our own words:
    'meta' and 'image_name' stay; a LIST of tensors (or ndarrays) is stacked along a new first axis -- which needs
    equal shapes, hence ``PointBatch`` -- and moved; a list of anything else is moved item by item; every other
    entry is moved with ``.cuda(non_blocking=True)``.  tests/test_plugins.py runs it side by side with the method
    extracted from the reference file."""
    for key in batch:
        if key in ("meta", "image_name"):
            continue
        value = batch[key]
        if not isinstance(value, list):
            batch[key] = value.cuda(non_blocking=True)
        elif isinstance(value[0], torch.Tensor):
            batch[key] = _to_device(torch.cat([v.unsqueeze(0) for v in value], dim=0))
        elif isinstance(value[0], np.ndarray):
            batch[key] = _to_device(torch.cat([torch.from_numpy(v).unsqueeze(0) for v in value], dim=0))
        else:
            batch[key] = [v.cuda() for v in value]
    return batch


def load_reference_to_cuda():
    """``Runner.to_cuda`` cut out of the reference's runner.py (the module itself needs mmcv & co.); None if absent."""
    import ast
    import textwrap
    path = os.path.join(REF_ROOT, "baseline", "engine", "runner.py")
    if not os.path.exists(path):
        return None
    src = open(path).read()
    tree = ast.parse(src)
    fn = next(n for c in ast.walk(tree) if isinstance(c, ast.ClassDef) and c.name == "Runner"
              for n in c.body if isinstance(n, ast.FunctionDef) and n.name == "to_cuda")
    ns = {"torch": torch, "np": np}
    exec(textwrap.dedent(ast.get_source_segment(src, fn)), ns)
    return ns["to_cuda"]
