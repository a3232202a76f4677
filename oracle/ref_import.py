"""Import the UNMODIFIED reference modules in the build container (TEST INFRASTRUCTURE ONLY).

The reference's hot-path modules import packages that are not installed here
(torch_geometric, torch_scatter, matplotlib, gym, stable_baselines3, ...;
SURVEY.md §8c).  None of them is needed for the arithmetic of the s2v path
except torch_scatter's scatter_mean/scatter_add and PyG's to_dense_adj, which
``oracle/s2v_ref.py`` restates.  This module installs a meta-path finder that
fabricates empty stand-ins for the absent packages so that
``modules.gnn.comb_opt`` and ``modules.ppo.models_basic_lstm`` import from
/root/reference as they are, then patches in the functional restatements.

Only usable where /root/reference exists (the build container); never at GPU
run time.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("SIMMODEL_REFERENCE", "/root/reference")

_ABSENT_ROOTS = {"torch_geometric", "torch_scatter", "torch_sparse", "matplotlib", "gym", "stable_baselines3",
                 "sb3_contrib", "dotmap", "gurobipy", "plotly", "seaborn"}


class _StubLoader(importlib.abc.Loader):
    def create_module(self, spec):
        mod = types.ModuleType(spec.name)
        mod.__path__ = []

        def _getattr(name):
            if name.startswith("__"):
                raise AttributeError(name)
            return type(name, (), {"__init__": lambda self, *a, **k: None})

        mod.__getattr__ = _getattr
        return mod

    def exec_module(self, module):
        pass


class _StubFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in _ABSENT_ROOTS:
            return importlib.machinery.ModuleSpec(name, _StubLoader(), is_package=True)
        return None


_installed = False


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "modules", "gnn"))


def load_reference():
    """Returns (comb_opt, models_basic_lstm) modules of the real reference."""
    global _installed
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True  # /root/reference is read-only
    if not _installed:
        sys.meta_path.append(_StubFinder())
        sys.path.insert(0, REFERENCE_ROOT)
        _installed = True
    from oracle import gat_ref, s2v_ref
    # the gat2 branch subclasses PyG classes at import time (comb_opt.py:48-62, models_basic_lstm.py:10-11,326):
    # put the restatements (oracle/gat_ref.py) into the fabricated torch_geometric modules BEFORE importing
    import torch_geometric.nn.conv as _pyg_conv
    import torch_geometric.nn.models as _pyg_models
    import torch_geometric.nn.models.basic_gnn as _pyg_basic
    _pyg_conv.GATv2Conv, _pyg_conv.MessagePassing = gat_ref.GATv2Conv, gat_ref.MessagePassing
    _pyg_basic.BasicGNN = gat_ref.BasicGNN
    _pyg_models.GAT = lambda **kw: torch.nn.Identity()     # QNet_GAT.__init__ (GATv1, not on our path) must construct
    import modules.gnn.comb_opt as comb_opt
    import modules.ppo.models_basic_lstm as models_basic_lstm

    comb_opt.scatter_mean = lambda src, index, dim=0: s2v_ref.scatter_mean(src, index)
    comb_opt.scatter_add = lambda src, index, dim=0: s2v_ref.scatter_add(src, index)
    models_basic_lstm.to_dense_adj = s2v_ref.to_dense_adj

    class _Batch:
        @staticmethod
        def from_data_list(data_list):
            if all(hasattr(d, "edge_index") and hasattr(d, "x") for d in data_list):      # gat2 branch
                return gat_ref.batch_from_data_list(data_list)
            return s2v_ref.batch_from_sizes([d.num_nodes for d in data_list])

    comb_opt.Batch = _Batch
    models_basic_lstm.Batch = _Batch
    models_basic_lstm.Data = gat_ref.Data
    return comb_opt, models_basic_lstm
