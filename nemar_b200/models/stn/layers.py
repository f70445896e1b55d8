"""STN building blocks on the B200 engine.  Attribute names follow the reference (models/stn/layers.py:73-106,
158-185,218-242) so state_dict keys match: Conv.conv2d / Conv.resnet_block, DownBlock.conv_0,
ResnetTransformer.model.<i>.conv_block.<1|5>."""
from functools import partial

import torch
import torch.nn as nn

from ...engine import functional as F
from ...engine import lib as L
from .. import networks as N

ACTS = {"relu": L.ACT_RELU, "leaky_relu": L.ACT_LRELU, "tanh": L.ACT_TANH, None: L.ACT_NONE}


def get_init_function(activation, init_function, **kwargs):
    """Reference layers.py:25-55 (including its quirks: 'zeros' is N(0,1e-5); None falls back to kaiming)."""
    a = 0.0
    if activation == "leaky_relu":
        a = kwargs.get("negative_slope", 0.2)
    gain = kwargs.get("gain", 0.02)
    if isinstance(init_function, str):
        if init_function == "kaiming":
            act = "relu" if activation is None else activation
            return partial(nn.init.kaiming_normal_, a=a, nonlinearity=act, mode="fan_in")
        if init_function == "dirac":
            return nn.init.dirac_
        if init_function == "xavier":
            act = "relu" if activation is None else activation
            return partial(nn.init.xavier_normal_, gain=nn.init.calculate_gain(nonlinearity=act, param=a))
        if init_function == "normal":
            return partial(nn.init.normal_, mean=0.0, std=gain)
        if init_function == "orthogonal":
            return partial(nn.init.orthogonal_, gain=gain)
        if init_function == "zeros":
            return partial(nn.init.normal_, mean=0.0, std=1e-5)
        return None
    if init_function is None:
        if activation in ("relu", "leaky_relu"):
            return partial(nn.init.kaiming_normal_, a=a, nonlinearity=activation)
        if activation in ("tanh", "sigmoid"):
            return partial(nn.init.xavier_normal_, gain=nn.init.calculate_gain(nonlinearity=activation, param=a))
        return None
    return init_function


class ResnetTransformer(nn.Module):
    """n reflect-padded InstanceNorm ResnetBlocks (reference layers.py:218-242)."""

    def __init__(self, dim, n_blocks, init_func):
        super().__init__()
        self.model = N.Holder()
        for i in range(n_blocks):
            self.model.add_module(str(i), N.ResnetBlock(dim, use_dropout=False, use_bias=True))
        self.n_blocks = n_blocks
        init_ = get_init_function("relu", init_func)
        for m in self.model.modules():
            if isinstance(m, N.Conv):
                init_(m.weight)
                if m.bias is not None:
                    m.bias.data.zero_()

    def run(self, t_padded, out_pad=0):
        for i in range(self.n_blocks):
            t_padded = getattr(self.model, str(i)).run(t_padded, out_pad=1 if i + 1 < self.n_blocks else out_pad)
        return t_padded


class Conv(nn.Module):
    """conv -> InstanceNorm? -> activation -> resblock?  (reference layers.py:73-106)"""

    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, bias=True, activation="relu",
                 init_func="kaiming", use_norm=False, use_resnet=False, **kwargs):
        super().__init__()
        self.conv2d = N.Conv(in_channels, out_channels, kernel_size, stride, padding, bias=bias)
        self.resnet_block = ResnetTransformer(out_channels, 1, init_func) if use_resnet else None
        self.use_norm = use_norm
        self.act = ACTS[activation]
        init_ = get_init_function(activation, init_func)
        init_(self.conv2d.weight)
        if self.conv2d.bias is not None:
            self.conv2d.bias.data.zero_()

    def run(self, x, x_pad=0, out_pad=0, out_f32=False):
        inner_pad = 1 if self.resnet_block is not None else out_pad
        if self.use_norm:
            t = N.conv_in_act(self.conv2d, x, x_pad, self.act, out_pad=inner_pad)
        else:
            t = N.conv_act(self.conv2d, x, x_pad, self.act, out_pad=inner_pad, out_f32=out_f32)
        if self.resnet_block is not None:
            t = self.resnet_block.run(t, out_pad=out_pad)
        return t


class DownBlock(nn.Module):
    """Conv (+Conv) then MaxPool2d(2); returns (pooled, skip) (reference layers.py:158-185)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, bias=False, activation="relu",
                 init_func="kaiming", use_norm=False, use_resnet=False, skip=True, refine=False, pool=True, **kwargs):
        super().__init__()
        self.conv_0 = Conv(in_channels, out_channels, kernel_size, stride, padding, bias=bias, activation=activation,
                           init_func=init_func, use_norm=use_norm, use_resnet=use_resnet)
        self.conv_1 = None
        if refine:
            self.conv_1 = Conv(out_channels, out_channels, kernel_size, stride, padding, bias=bias,
                               activation=activation, init_func=init_func, use_norm=use_norm, use_resnet=use_resnet)
        self.skip, self.pool = skip, pool

    def run(self, x):
        x = skip = self.conv_0.run(x)
        if self.conv_1 is not None:
            x = skip = self.conv_1.run(x)
        if self.pool:
            x = F.MaxPool2Fn.apply(x)
        return (x, skip) if self.skip else x
