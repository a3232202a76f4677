"""simmodel_b200: B200-native s2v / GNN-embedding hot path of rvdweerd/simmodel.

Public surface (mirrors the reference's own interface for this path):
  QNet, QFunction, Batch, Data      modules/gnn/comb_opt.py:184-366
  PPO_GNN_Single_LSTM               modules/ppo/models_basic_lstm.py:407-566 (qnet='s2v' | 'gat2')
  QNet_GATv2, GATv2, GATv2Conv      modules/gnn/comb_opt.py:48-62,156-182 (--qnet gat2)
  sb3.Struc2VecExtractor, sb3.s2v_ACNetwork, sb3.DeployablePPOPolicy   modules/ppo/models_sb3_s2v.py
  ngo.FeatureExtractor, ngo.Actor, ngo.Critic, ngo.MaskablePPOPolicy   modules/ppo/models_ngo.py
  GraphLayout, ops                  the native (x, edge_index, ptr, n_pad, mask) contract
  torch_ops                         the same entry points as torch.library custom ops (torch.ops.s2v_b200.*)
The arithmetic lives in csrc/*.cu behind the C ABI of include/s2v_b200.h
(libs2v_b200.so, built by ``python -m simmodel_b200.build``); there is no CPU fallback.
"""
from . import graphs, ops                                   # noqa: F401
from ._lib import S2VError, LIB_PATH                         # noqa: F401
from .layout import GraphLayout                              # noqa: F401
from .gat import GATv2, GATv2Conv, QNet_GATv2                # noqa: F401
from .ppo import PPO_GNN_Single_LSTM                         # noqa: F401
from . import ngo, sb3, torch_ops                            # noqa: F401
from .qnet import Batch, Data, QFunction, QNet, ranges_from_ptr  # noqa: F401

__all__ = ["QNet", "QNet_GATv2", "GATv2", "GATv2Conv", "QFunction", "Batch", "Data", "PPO_GNN_Single_LSTM", "GraphLayout", "ops", "graphs",
           "S2VError", "ranges_from_ptr"]
