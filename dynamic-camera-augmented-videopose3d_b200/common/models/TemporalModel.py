"""Drop-in replacement for the reference's common/models/TemporalModel.py, backed by sm_100a CUDA kernels.

Same classes, constructor signatures, attributes and state_dict layout as the reference
(TemporalModelBase :10-76, TemporalModel :79-138, TemporalModelOptimized1f :141-198), so run.py's
`from common.models.TemporalModel import *`, load_state_dict / state_dict, set_bn_momentum and
isinstance(m, TemporalModelBase) dispatch keep working. The nn.Conv1d / nn.BatchNorm1d children are *parameter
containers only* (identical names, shapes, default initialisation and RNG consumption); their forward is never called.
forward() runs the tcgen05 implicit-GEMM kernels of libvp3d_b200.so. CUDA tensors only -- there is no CPU fallback.

Extra (non-reference) knob: `model.operand_dtype` in {'fp16', 'bf16', 'tf32'} (default: env VP3D_DTYPE or 'fp16')
selects the tensor-core operand type; accumulation is fp32 in all cases.
"""
import torch
import torch.nn as nn

from vp3d_b200 import temporal as _engine


class TemporalModelBase(nn.Module):
    """
    Do not instantiate this class.
    """

    _strided = False

    def __init__(self, num_joints_in, in_features, num_joints_out,
                 filter_widths, causal, dropout, channels):
        super().__init__()
        for width in filter_widths:
            assert width % 2 != 0, 'Only odd filter widths are supported'

        self.num_joints_in = num_joints_in
        self.in_features = in_features
        self.num_joints_out = num_joints_out
        self.filter_widths = filter_widths
        self.operand_dtype = None

        # registration order matters for state_dict()/parameters() ordering: drop, relu, expand_bn, shrink
        self.drop = nn.Dropout(dropout)
        self.relu = nn.ReLU(inplace=True)
        self.pad = [filter_widths[0] // 2]
        self.expand_bn = nn.BatchNorm1d(channels, momentum=0.1)
        self.shrink = nn.Conv1d(channels, num_joints_out * 3, 1)

    def _build_stack(self, channels, causal, dense):
        """Creates expand_conv / layers_conv / layers_bn and fills pad / causal_shift for either variant."""
        widths = self.filter_widths
        first = widths[0]
        in_ch = self.num_joints_in * self.in_features
        if self._strided:
            self.expand_conv = nn.Conv1d(in_ch, channels, first, stride=first, bias=False)
        else:
            self.expand_conv = nn.Conv1d(in_ch, channels, first, bias=False)
        self.causal_shift = [first // 2 if causal else 0]

        convs, norms = [], []
        reach = first  # dilation the next block would use in the dilated model
        for width in widths[1:]:
            half_span = (width - 1) * reach // 2
            self.pad.append(half_span)
            if self._strided:
                self.causal_shift.append(width // 2 if causal else 0)
                convs.append(nn.Conv1d(channels, channels, width, stride=width, bias=False))
            else:
                self.causal_shift.append(width // 2 * reach if causal else 0)
                if dense:
                    convs.append(nn.Conv1d(channels, channels, 2 * half_span + 1, dilation=1, bias=False))
                else:
                    convs.append(nn.Conv1d(channels, channels, width, dilation=reach, bias=False))
            norms.append(nn.BatchNorm1d(channels, momentum=0.1))
            convs.append(nn.Conv1d(channels, channels, 1, dilation=1, bias=False))
            norms.append(nn.BatchNorm1d(channels, momentum=0.1))
            reach *= width
        self.layers_conv = nn.ModuleList(convs)
        self.layers_bn = nn.ModuleList(norms)

    def set_bn_momentum(self, momentum):
        self.expand_bn.momentum = momentum
        for bn in self.layers_bn:
            bn.momentum = momentum

    def receptive_field(self):
        """
        Return the total receptive field of this model as # of frames.
        """
        return 1 + 2 * sum(self.pad)

    def total_causal_shift(self):
        """
        Return the asymmetric offset for sequence padding.
        The returned value is typically 0 if causal convolutions are disabled,
        otherwise it is half the receptive field.
        (Kept bug-compatible with the reference: TemporalModel(causal=True) stores shifts that already include the
        dilation, so this returns 7381 rather than 121 for 3,3,3,3,3 -- SURVEY appendix A.)
        """
        total = self.causal_shift[0]
        scale = self.filter_widths[0]
        for i in range(1, len(self.filter_widths)):
            total += self.causal_shift[i] * scale
            scale *= self.filter_widths[i]
        return total

    def forward(self, x):
        assert len(x.shape) == 4
        assert x.shape[-2] == self.num_joints_in
        assert x.shape[-1] == self.in_features

        n, t = x.shape[0], x.shape[1]
        y = self._forward_blocks(x.reshape(n, t, -1))     # channels-last in, channels-last out: (N, T', 3*J_out)
        return y.view(n, -1, self.num_joints_out, 3)

    def _forward_blocks(self, x):
        if self.training:
            from vp3d_b200 import training as _training
            return _training.forward_train(self, x)
        return _engine.forward_eval(self, x)


class TemporalModel(TemporalModelBase):
    """
    Reference 3D pose estimation model with temporal convolutions.
    This implementation can be used for all use-cases.
    """

    def __init__(self, num_joints_in, in_features, num_joints_out,
                 filter_widths, causal=False, dropout=0.25, channels=1024, dense=False):
        """
        Initialize this model.

        Arguments:
        num_joints_in -- number of input joints (e.g. 17 for Human3.6M)
        in_features -- number of input features for each joint (typically 2 for 2D input)
        num_joints_out -- number of output joints (can be different than input)
        filter_widths -- list of convolution widths, which also determines the # of blocks and receptive field
        causal -- use causal convolutions instead of symmetric convolutions (for real-time applications)
        dropout -- dropout probability
        channels -- number of convolution channels
        dense -- use regular dense convolutions instead of dilated convolutions (ablation experiment)
        """
        super().__init__(num_joints_in, in_features, num_joints_out, filter_widths, causal, dropout, channels)
        self._build_stack(channels, causal, dense)


class TemporalModelOptimized1f(TemporalModelBase):
    """
    3D pose estimation model optimized for single-frame batching, i.e.
    where batches have input length = receptive field, and output length = 1.
    This scenario is only used for training when stride == 1.

    Strided convolutions replace the dilated ones; on channels-last data every layer is a plain GEMM on a reshaped
    view. The weights are interchangeable with TemporalModel.
    """

    _strided = True

    def __init__(self, num_joints_in, in_features, num_joints_out,
                 filter_widths, causal=False, dropout=0.25, channels=1024):
        """
        Initialize this model.

        Arguments:
        num_joints_in -- number of input joints (e.g. 17 for Human3.6M)
        in_features -- number of input features for each joint (typically 2 for 2D input)
        num_joints_out -- number of output joints (can be different than input)
        filter_widths -- list of convolution widths, which also determines the # of blocks and receptive field
        causal -- use causal convolutions instead of symmetric convolutions (for real-time applications)
        dropout -- dropout probability
        channels -- number of convolution channels
        """
        super().__init__(num_joints_in, in_features, num_joints_out, filter_widths, causal, dropout, channels)
        self._build_stack(channels, causal, dense=False)
