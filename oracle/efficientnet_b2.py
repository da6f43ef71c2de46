"""Oracle restatement of ``efficientnet_pytorch==0.7.1`` (EfficientNet-B2 feature path).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  PARITY UNPINNED at this
boundary: the third-party package (requirements.txt:8 of the reference) is not
vendored in /root/reference and cannot be installed here.  What is restated is
the published algorithm of that package, anchored on the reference's own call
sites and in-repo restatements:

* call sites: captioning/models/hf_wrapper.py:12-13,218-241 (``EfficientNet(blocks_args,
  global_params)``, ``extract_features``, ``_change_in_channels(1)``)
* constructor structure: captioning/models/eff_latent_encoder.py:76-118 (MBConv block),
  :123-206 (network), state-dict key list :263-290
* the canonical input shape (1, 64, 1001): captioning/models/flops_counting_model.py:330

Module and parameter names match the package so that a ``state_dict`` is
interchangeable (``_conv_stem``, ``_bn0``, ``_blocks.N._expand_conv`` ...).
"""
import math
from dataclasses import dataclass, replace
from typing import List, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class BlockArgs:
    num_repeat: int
    kernel_size: int
    stride: int
    expand_ratio: int
    input_filters: int
    output_filters: int
    se_ratio: float
    id_skip: bool = True


@dataclass
class GlobalParams:
    width_coefficient: float = 1.1
    depth_coefficient: float = 1.2
    image_size: int = 260
    dropout_rate: float = 0.3
    batch_norm_momentum: float = 0.99
    batch_norm_epsilon: float = 1e-3
    drop_connect_rate: float = 0.2
    depth_divisor: int = 8
    include_top: bool = False


# efficientnet-b0 base stages; B2 scales width x1.1 and depth x1.2
BASE_STAGES = [
    BlockArgs(1, 3, 1, 1, 32, 16, 0.25),
    BlockArgs(2, 3, 2, 6, 16, 24, 0.25),
    BlockArgs(2, 5, 2, 6, 24, 40, 0.25),
    BlockArgs(3, 3, 2, 6, 40, 80, 0.25),
    BlockArgs(3, 5, 1, 6, 80, 112, 0.25),
    BlockArgs(4, 5, 2, 6, 112, 192, 0.25),
    BlockArgs(1, 3, 1, 6, 192, 320, 0.25),
]


def get_model_params(model_name="efficientnet-b2", override=None):
    assert model_name == "efficientnet-b2"
    gp = GlobalParams()
    if override:
        gp = replace(gp, **override)
    return [replace(b) for b in BASE_STAGES], gp


def round_filters(filters: int, gp: GlobalParams) -> int:
    mult, div = gp.width_coefficient, gp.depth_divisor
    filters = filters * mult
    new = max(div, int(filters + div / 2) // div * div)
    if new < 0.9 * filters:
        new += div
    return int(new)


def round_repeats(repeats: int, gp: GlobalParams) -> int:
    return int(math.ceil(gp.depth_coefficient * repeats))


def out_image_size(size: Tuple[int, int], stride: int) -> Tuple[int, int]:
    return (int(math.ceil(size[0] / stride)), int(math.ceil(size[1] / stride)))


def static_same_pad(image_size: Tuple[int, int], k: int, s: int) -> Tuple[int, int, int, int]:
    """(left, right, top, bottom) zero padding of TF 'SAME' computed ONCE for
    ``image_size`` (260-derived), then applied to whatever input arrives."""
    ih, iw = image_size
    oh, ow = math.ceil(ih / s), math.ceil(iw / s)
    pad_h = max((oh - 1) * s + (k - 1) + 1 - ih, 0)
    pad_w = max((ow - 1) * s + (k - 1) + 1 - iw, 0)
    return (pad_w // 2, pad_w - pad_w // 2, pad_h // 2, pad_h - pad_h // 2)


class Conv2dStaticSamePadding(nn.Conv2d):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, image_size=None, **kw):
        super().__init__(in_channels, out_channels, kernel_size, stride, **kw)
        assert image_size is not None
        if isinstance(image_size, int):
            image_size = (image_size, image_size)
        self.pads = static_same_pad(image_size, self.kernel_size[0], self.stride[0])

    def forward(self, x):
        if any(self.pads):
            x = F.pad(x, self.pads)
        return F.conv2d(x, self.weight, self.bias, self.stride, 0, self.dilation, self.groups)


def swish(x):
    return x * torch.sigmoid(x)


class MBConvBlock(nn.Module):
    def __init__(self, args: BlockArgs, gp: GlobalParams, image_size):
        super().__init__()
        self.args = args
        mom, eps = 1 - gp.batch_norm_momentum, gp.batch_norm_epsilon
        inp = args.input_filters
        oup = inp * args.expand_ratio
        if args.expand_ratio != 1:
            self._expand_conv = Conv2dStaticSamePadding(inp, oup, 1, image_size=image_size, bias=False)
            self._bn0 = nn.BatchNorm2d(oup, momentum=mom, eps=eps)
        self._depthwise_conv = Conv2dStaticSamePadding(
            oup, oup, args.kernel_size, stride=args.stride, image_size=image_size, groups=oup, bias=False)
        self._bn1 = nn.BatchNorm2d(oup, momentum=mom, eps=eps)
        image_size = out_image_size(image_size, args.stride)
        nsq = max(1, int(args.input_filters * args.se_ratio))
        self._se_reduce = Conv2dStaticSamePadding(oup, nsq, 1, image_size=(1, 1))
        self._se_expand = Conv2dStaticSamePadding(nsq, oup, 1, image_size=(1, 1))
        self._project_conv = Conv2dStaticSamePadding(oup, args.output_filters, 1, image_size=image_size, bias=False)
        self._bn2 = nn.BatchNorm2d(args.output_filters, momentum=mom, eps=eps)
        self.has_skip = args.id_skip and args.stride == 1 and args.input_filters == args.output_filters

    def forward(self, inputs):
        x = inputs
        if self.args.expand_ratio != 1:
            x = swish(self._bn0(self._expand_conv(x)))
        x = swish(self._bn1(self._depthwise_conv(x)))
        s = F.adaptive_avg_pool2d(x, 1)
        s = self._se_expand(swish(self._se_reduce(s)))
        x = torch.sigmoid(s) * x
        x = self._bn2(self._project_conv(x))
        if self.has_skip:
            # drop-connect only acts in training; inference path is the plain residual
            x = x + inputs
        return x


class EfficientNet(nn.Module):
    """Feature extractor only (``include_top=False``)."""

    def __init__(self, blocks_args: List[BlockArgs] = None, global_params: GlobalParams = None):
        super().__init__()
        if blocks_args is None:
            blocks_args, global_params = get_model_params()
        gp = global_params
        self._global_params = gp
        mom, eps = 1 - gp.batch_norm_momentum, gp.batch_norm_epsilon
        image_size = (gp.image_size, gp.image_size)
        out_c = round_filters(32, gp)
        self._conv_stem = Conv2dStaticSamePadding(3, out_c, 3, stride=2, image_size=image_size, bias=False)
        self._bn0 = nn.BatchNorm2d(out_c, momentum=mom, eps=eps)
        image_size = out_image_size(image_size, 2)
        self._blocks = nn.ModuleList()
        self.block_cfgs = []
        for st in blocks_args:
            st = replace(st, input_filters=round_filters(st.input_filters, gp),
                         output_filters=round_filters(st.output_filters, gp),
                         num_repeat=round_repeats(st.num_repeat, gp))
            self._blocks.append(MBConvBlock(st, gp, image_size))
            self.block_cfgs.append((st, image_size))
            image_size = out_image_size(image_size, st.stride)
            rep = replace(st, input_filters=st.output_filters, stride=1)
            for _ in range(st.num_repeat - 1):
                self._blocks.append(MBConvBlock(rep, gp, image_size))
                self.block_cfgs.append((rep, image_size))
            last = st.output_filters
        self._conv_head = Conv2dStaticSamePadding(last, round_filters(1280, gp), 1, image_size=image_size, bias=False)
        self._bn1 = nn.BatchNorm2d(round_filters(1280, gp), momentum=mom, eps=eps)

    def _change_in_channels(self, in_channels):
        gp = self._global_params
        self._conv_stem = Conv2dStaticSamePadding(
            in_channels, round_filters(32, gp), 3, stride=2,
            image_size=(gp.image_size, gp.image_size), bias=False)

    def extract_features(self, x):
        x = swish(self._bn0(self._conv_stem(x)))
        for blk in self._blocks:
            x = blk(x)
        return swish(self._bn1(self._conv_head(x)))


def layer_plan():
    """Static description of the 23 blocks: used by tests to cross-check the CUDA
    library's own plan (channels, kernel, stride, pads)."""
    net = EfficientNet()
    plan = []
    for (a, img), blk in zip(net.block_cfgs, net._blocks):
        plan.append(dict(cin=a.input_filters, cout=a.output_filters, expand=a.expand_ratio,
                         k=a.kernel_size, s=a.stride, pads=blk._depthwise_conv.pads,
                         nsq=blk._se_reduce.out_channels, skip=blk.has_skip))
    return plan
