"""Test infrastructure: the two pieces of the reference's plug-in machinery the on-the-fly path goes
through, restated so that the GPU box (which has no /root/reference) can exercise the plug-ins.

* ``Registry`` / ``build_from_cfg``  -- reference baseline/utils/registry.py:12-84
* ``runner_to_cuda``                 -- reference baseline/engine/runner.py:125-152 (``Runner.to_cuda``)

tests/test_plugins.py holds both to the reference where it is mounted: the real ``registry.py`` is loaded by
path and must behave identically on the same calls, and ``runner_to_cuda`` must have the same AST as the
method in the reference file.  Nothing in the product imports this module.
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


class Registry(object):
    # reference baseline/utils/registry.py:12-51
    def __init__(self, name):
        self._name = name
        self._module_dict = dict()

    @property
    def name(self):
        return self._name

    def get(self, key):
        return self._module_dict.get(key, None)

    def register_module(self, cls):
        if not inspect.isclass(cls):
            raise TypeError('module must be a class, but got {}'.format(type(cls)))
        if cls.__name__ in self._module_dict:
            raise KeyError('{} is already registered in {}'.format(cls.__name__, self.name))
        self._module_dict[cls.__name__] = cls
        return cls


def build_from_cfg(cfg, registry, default_args=None):
    # reference baseline/utils/registry.py:54-84: kwargs = cfg minus 'type', default_args fill the gaps
    assert isinstance(cfg, dict) and 'type' in cfg
    args = cfg.copy()
    obj_type = args.pop('type')
    if isinstance(obj_type, str):
        obj_cls = registry.get(obj_type)
        if obj_cls is None:
            raise KeyError('{} is not in the {} registry'.format(obj_type, registry.name))
    elif inspect.isclass(obj_type):
        obj_cls = obj_type
    else:
        raise TypeError('type must be a str or valid type, but got {}'.format(type(obj_type)))
    if default_args is not None:
        for name, value in default_args.items():
            args.setdefault(name, value)
    return obj_cls(**args)


def registries():
    """(module providing Registry/build_from_cfg, is_reference): the real one where mounted."""
    ref = load_reference_registry()
    if ref is not None:
        return ref, True
    return sys.modules[__name__], False


def runner_to_cuda(self, batch):
    for k in batch:
        if k == 'meta':
            continue
        if k == 'image_name':
            continue
        if isinstance(batch[k], list):
            if isinstance(batch[k][0], torch.Tensor):
                batch[k] = [ item.unsqueeze(0) for item in batch[k]]
                batch[k] = torch.cat(batch[k], dim=0).cuda()
            elif isinstance(batch[k][0], np.ndarray):
                batch[k] = [ torch.from_numpy(item).unsqueeze(0) for item in batch[k]]
                batch[k] = torch.cat(batch[k], dim=0).cuda()
            else:
                batch[k] = [item.cuda() for item in batch[k]]
        else:
            batch[k] = batch[k].cuda(non_blocking=True)

    return batch
