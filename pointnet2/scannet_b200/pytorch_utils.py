"""Shared-MLP building blocks -- mirror of /root/reference/pointnet2/pytorch_utils.py
for the parts the SA/FP path instantiates: ``SharedMLP`` (ref :11-36), ``Conv1d``
(:123-154), ``Conv2d`` (:157-188) and the BatchNorm wrappers (:39-64).

Module and parameter names are the reference's, so state dicts are
interchangeable (``...layer0.conv.weight``, ``...layer0.bn.bn.running_mean``),
and construction consumes the RNG in the reference's order (conv default init,
then ``init(weight)``), so the same seed yields the same parameters.
``FC``, ``Conv3d`` and ``BNMomentumScheduler`` (ref :191-297) are never
instantiated by the semseg models and are out of scope.
"""
import torch.nn as nn


class _BNWrap(nn.Sequential):
    """``<name>bn`` child holding the torch BatchNorm; weight 1, bias 0 (ref :39-48)."""

    def __init__(self, channels, norm_cls, name=""):
        super().__init__()
        self.add_module(name + "bn", norm_cls(channels))
        nn.init.constant_(self[0].weight, 1.0)
        nn.init.constant_(self[0].bias, 0)


class BatchNorm1d(_BNWrap):
    def __init__(self, in_size, *, name=""):
        super().__init__(in_size, nn.BatchNorm1d, name)


class BatchNorm2d(_BNWrap):
    def __init__(self, in_size, name=""):
        super().__init__(in_size, nn.BatchNorm2d, name)


class _ConvUnit(nn.Sequential):
    """conv (bias only without BN) -> BN -> activation; ``preact`` puts BN and
    activation in front of the conv (ref :67-121)."""

    conv_cls = None
    norm_cls = None

    def __init__(self, in_size, out_size, *, kernel_size, stride, padding,
                 activation=nn.ReLU(inplace=True), bn=False, init=nn.init.kaiming_normal_,
                 bias=True, preact=False, name=""):
        super().__init__()
        bias = bias and not bn
        conv = self.conv_cls(in_size, out_size, kernel_size=kernel_size, stride=stride,
                             padding=padding, bias=bias)
        init(conv.weight)
        if bias:
            nn.init.constant_(conv.bias, 0)
        norm = self.norm_cls(in_size if preact else out_size) if bn else None

        def add_norm_act():
            if norm is not None:
                self.add_module(name + "bn", norm)
            if activation is not None:
                self.add_module(name + "activation", activation)

        if preact:
            add_norm_act()
        self.add_module(name + "conv", conv)
        if not preact:
            add_norm_act()


class Conv1d(_ConvUnit):
    conv_cls = nn.Conv1d
    norm_cls = BatchNorm1d

    def __init__(self, in_size, out_size, *, kernel_size=1, stride=1, padding=0, **kw):
        super().__init__(in_size, out_size, kernel_size=kernel_size, stride=stride,
                         padding=padding, **kw)


class Conv2d(_ConvUnit):
    conv_cls = nn.Conv2d
    norm_cls = BatchNorm2d

    def __init__(self, in_size, out_size, *, kernel_size=(1, 1), stride=(1, 1),
                 padding=(0, 0), **kw):
        super().__init__(in_size, out_size, kernel_size=kernel_size, stride=stride,
                         padding=padding, **kw)


class SharedMLP(nn.Sequential):
    """Per-point MLP: a chain of 1x1 Conv2d (+BN) + ReLU named ``layer{i}`` (ref :11-36)."""

    def __init__(self, args, *, bn=False, activation=nn.ReLU(inplace=True), preact=False,
                 first=False, name=""):
        super().__init__()
        for i in range(len(args) - 1):
            plain = first and preact and i == 0
            self.add_module(
                name + "layer{}".format(i),
                Conv2d(args[i], args[i + 1], bn=(not plain) and bn,
                       activation=None if plain else activation, preact=preact))
