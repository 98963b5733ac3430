"""CPU oracle for common/models/StackedPoseLifter.py and the sliding-window evaluator of the camera-aware sibling models.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): only tests/ may import this module.

Functional restatement: the model is a dict of tensors under the reference's state_dict keys
(`mlp_layers.{0,3,6,...}.weight / .bias`, StackedPoseLifter.py:21-34) and the forward spells out what the nn.ModuleList
does in eval mode (Linear -> ReLU -> [Dropout = identity]) x (1 + num_layers) -> Linear (StackedPoseLifter.py:37-56).
Backward comes from torch autograd, as in the reference (run.py:485).
Parity pin: tests/golden/lifter.npz, produced by importing the real reference (tests/golden/make_golden_lifter.py).
"""
import torch
import torch.nn.functional as F


def linear_indices(num_layers):
    """Positions of the nn.Linear modules inside mlp_layers (StackedPoseLifter.py:21-32): 0, 3, ..., 3 * (num_layers + 1)."""
    return [3 * i for i in range(num_layers + 2)]


def init_state(num_joints, features, num_layers, layer_size, seed=0):
    """Random parameters with the reference's shapes and nn.Linear's default init range (uniform +-1/sqrt(fan_in))."""
    g = torch.Generator().manual_seed(seed)
    dims = [num_joints * features * 2] + [layer_size] * (num_layers + 1) + [num_joints * features]
    sd = {}
    for idx, (fan_in, fan_out) in zip(linear_indices(num_layers), zip(dims[:-1], dims[1:])):
        bound = 1.0 / fan_in ** 0.5
        sd['mlp_layers.%d.weight' % idx] = (torch.rand(fan_out, fan_in, generator=g) * 2 - 1) * bound
        sd['mlp_layers.%d.bias' % idx] = (torch.rand(fan_out, generator=g) * 2 - 1) * bound
    return sd


def forward(sd, input_3d_transformer, input_3d_fcn, num_joints, features):
    """StackedPoseLifter.forward in eval mode (:37-56): (B, ..., J, F) x 2 -> (B, 1, J, F)."""
    a = input_3d_transformer.reshape(input_3d_transformer.size(0), -1)      # :47
    b = input_3d_fcn.reshape(input_3d_fcn.size(0), -1)                        # :48
    x = torch.cat((a, b), dim=-1)                                             # :50
    idx = sorted(int(k.split('.')[1]) for k in sd if k.endswith('.weight'))
    for i in idx[:-1]:
        x = F.relu(F.linear(x, sd['mlp_layers.%d.weight' % i], sd['mlp_layers.%d.bias' % i]))   # Linear, ReLU, Dropout(eval)
    x = F.linear(x, sd['mlp_layers.%d.weight' % idx[-1]], sd['mlp_layers.%d.bias' % idx[-1]])
    return x.view(x.size(0), 1, num_joints, features)                         # :55


def sliding_window(model_fn, inputs_2d, inputs_cam, window_size, num_joints_out, out_features):
    """CamLSTMBase.sliding_window (CamLSTM.py:33-44; CamTransformer.py:72-91 is the same): every window of
    `window_size` consecutive frames of ONE sequence becomes a batch element; the model maps a window to one pose."""
    _, T, J, _ = inputs_2d.shape
    n_windows = T - window_size + 1
    if n_windows <= 0:
        raise ValueError("window_size larger than sequence length")
    win_2d = torch.stack([inputs_2d[0, i:i + window_size] for i in range(n_windows)])      # (n_windows, W, J, F)
    win_cam = torch.stack([inputs_cam[0, i:i + window_size] for i in range(n_windows)])    # (n_windows, W, 3, 4)
    out = model_fn(win_2d, win_cam)
    return out.view(1, n_windows, num_joints_out, out_features)
