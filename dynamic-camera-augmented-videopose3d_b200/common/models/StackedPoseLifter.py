"""Drop-in replacement for the reference's common/models/StackedPoseLifter.py, backed by the sm_100a kernels of the
temporal stack (vp3d_b200.lifter). Same constructor, attributes and state_dict layout (`mlp_layers.{0,3,...}.weight /
.bias`, reference :21-34), so run.py's `from common.models.StackedPoseLifter import *`, the isinstance dispatch
(run.py:474,517,714) and checkpoints keep working. The nn.Linear / nn.ReLU / nn.Dropout children are parameter containers
in the reference's order; their forward is never called. CUDA tensors only -- there is no CPU fallback.
"""
import torch
import torch.nn as nn

from vp3d_b200 import lifter as _engine


class StackedPoseLifter(nn.Module):

    def __init__(self, num_joints: int, features: int, num_layers: int, layer_size: int, dropout: float = 0.25):
        super().__init__()
        self.num_joints, self.features = num_joints, features
        self.num_layers, self.layer_size = num_layers, layer_size
        self.operand_dtype = None
        # (Linear, ReLU, Dropout) for the input layer and each of the num_layers hidden layers, then the output Linear:
        # the module indices 0, 3, 6, ... are the state_dict keys (nn.Linear draws its default init in this order too)
        widths = [num_joints * features * 2] + [layer_size] * (num_layers + 1)
        mods = []
        for fan_in, fan_out in zip(widths[:-1], widths[1:]):
            mods += [nn.Linear(fan_in, fan_out), nn.ReLU(inplace=True), nn.Dropout(dropout)]
        mods.append(nn.Linear(layer_size, num_joints * features))
        self.mlp_layers = nn.ModuleList(mods)
        self.dropout = nn.Dropout(dropout)

    def forward(self, input_3d_transformer: torch.Tensor, input_3d_FCN: torch.Tensor):
        """(B, 1, J, features) x 2 -- or any (B, ...) with J * features trailing elements, run.py:521-522 passes squeezed
        (T, J, features) -- to (B, 1, J, features)."""
        return _engine.forward(self, input_3d_transformer, input_3d_FCN)
