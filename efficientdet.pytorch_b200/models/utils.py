"""Architecture tables and parameter-holder layers of the EfficientNet backbone.

Mirrors the public names of the reference's ``models/utils.py`` that other modules and user code
touch (``GlobalParams``, ``BlockArgs``, ``round_filters``, ``round_repeats``, ``get_model_params``,
``efficientnet_params``, ``get_same_padding_conv2d``, ``Conv2dStaticSamePadding``,
``MemoryEfficientSwish``, ``drop_connect``, ``load_pretrained_weights``) -- but the layers here are
*parameter holders*: arithmetic happens in the fused sm_100a kernels driven from
``efficientnet.py``; calling a holder on its own routes to the same kernels (never to ATen conv).
"""
import collections
import math
import os

import torch
from torch import nn

GlobalParams = collections.namedtuple('GlobalParams', [
    'batch_norm_momentum', 'batch_norm_epsilon', 'dropout_rate', 'num_classes', 'width_coefficient',
    'depth_coefficient', 'depth_divisor', 'min_depth', 'drop_connect_rate', 'image_size'])
BlockArgs = collections.namedtuple('BlockArgs', [
    'kernel_size', 'num_repeat', 'input_filters', 'output_filters', 'expand_ratio', 'id_skip', 'stride', 'se_ratio'])
GlobalParams.__new__.__defaults__ = (None,) * len(GlobalParams._fields)
BlockArgs.__new__.__defaults__ = (None,) * len(BlockArgs._fields)

# name -> (width, depth, nominal resolution, dropout)          reference models/utils.py:171-184
_COEFFS = {'efficientnet-b0': (1.0, 1.0, 224, 0.2), 'efficientnet-b1': (1.0, 1.1, 240, 0.2),
           'efficientnet-b2': (1.1, 1.2, 260, 0.3), 'efficientnet-b3': (1.2, 1.4, 300, 0.3),
           'efficientnet-b4': (1.4, 1.8, 380, 0.4), 'efficientnet-b5': (1.6, 2.2, 456, 0.4),
           'efficientnet-b6': (1.8, 2.6, 528, 0.5), 'efficientnet-b7': (2.0, 3.1, 600, 0.5)}

# (kernel, repeats, in, out, expand, stride).  Stages 5 and 7 are stride 2 in this detector's
# backbone (reference models/utils.py:267-268) so the seven stage outputs sit at strides 2..128.
_STAGE_TABLE = ((3, 1, 32, 16, 1, 1), (3, 2, 16, 24, 6, 2), (5, 2, 24, 40, 6, 2), (3, 3, 40, 80, 6, 2),
                (5, 3, 80, 112, 6, 2), (5, 4, 112, 192, 6, 2), (3, 1, 192, 320, 6, 2))


def efficientnet_params(model_name):
    return _COEFFS[model_name]


def round_filters(filters, global_params):
    mult = global_params.width_coefficient
    if not mult:
        return filters
    div = global_params.depth_divisor
    floor_ = global_params.min_depth or div
    scaled = filters * mult
    rounded = max(floor_, int(scaled + div / 2) // div * div)
    if rounded < 0.9 * scaled:
        rounded += div
    return int(rounded)


def round_repeats(repeats, global_params):
    mult = global_params.depth_coefficient
    return int(math.ceil(mult * repeats)) if mult else repeats


def efficientnet(width_coefficient=None, depth_coefficient=None, dropout_rate=0.2, drop_connect_rate=0.2,
                 image_size=None, num_classes=1000):
    blocks = [BlockArgs(kernel_size=k, num_repeat=r, input_filters=i, output_filters=o, expand_ratio=e,
                        id_skip=True, stride=[s], se_ratio=0.25) for (k, r, i, o, e, s) in _STAGE_TABLE]
    gp = GlobalParams(batch_norm_momentum=0.99, batch_norm_epsilon=1e-3, dropout_rate=dropout_rate,
                      drop_connect_rate=drop_connect_rate, num_classes=num_classes,
                      width_coefficient=width_coefficient, depth_coefficient=depth_coefficient, depth_divisor=8,
                      min_depth=None, image_size=image_size)
    return blocks, gp


def get_model_params(model_name, override_params):
    if not model_name.startswith('efficientnet'):
        raise NotImplementedError('model name is not pre-defined: %s' % model_name)
    w, d, s, p = efficientnet_params(model_name)
    blocks, gp = efficientnet(width_coefficient=w, depth_coefficient=d, dropout_rate=p, image_size=s)
    if override_params:
        gp = gp._replace(**override_params)
    return blocks, gp


def static_same_pad(kernel, stride, image_size):
    """(left, right, top, bottom) of the reference's Conv2dStaticSamePadding: TF-'SAME' evaluated once
    for the *nominal* image size (reference models/utils.py:134-149)."""
    ih, iw = image_size if isinstance(image_size, (list, tuple)) else (image_size, image_size)
    ph = max((math.ceil(ih / stride) - 1) * stride + kernel - ih, 0)
    pw = max((math.ceil(iw / stride) - 1) * stride + kernel - iw, 0)
    return (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2)


class Conv2dStaticSamePadding(nn.Conv2d):
    """Parameter holder with the reference's name, ctor and ``weight``/``bias`` layout.  ``same_pad``
    carries the static (left, right, top, bottom) padding consumed by the fused kernels."""

    def __init__(self, in_channels, out_channels, kernel_size, image_size=None, **kwargs):
        super().__init__(in_channels, out_channels, kernel_size, **kwargs)
        assert image_size is not None
        self.stride = self.stride if len(self.stride) == 2 else [self.stride[0]] * 2
        self.same_pad = static_same_pad(self.kernel_size[0], self.stride[0], image_size)

    def forward(self, x):
        raise RuntimeError('Conv2dStaticSamePadding is a parameter holder in the B200 build; its arithmetic is '
                           'fused into MBConvBlock / EfficientNet kernels')


class Identity(nn.Module):
    def forward(self, x):
        return x


def get_same_padding_conv2d(image_size=None):
    if image_size is None:
        raise NotImplementedError('dynamic same padding is not used by EfficientDet (image_size is always set)')
    from functools import partial
    return partial(Conv2dStaticSamePadding, image_size=image_size)


class MemoryEfficientSwish(nn.Module):
    """Kept for API parity (``set_swish``); swish is fused into the conv epilogues."""

    def forward(self, x):
        raise RuntimeError('swish is fused into the B200 conv kernels; it is not a stand-alone layer here')


Swish = MemoryEfficientSwish


def drop_connect_scale(batch, p, device):
    """The per-sample multiplier of the reference's drop_connect (models/utils.py:79-90):
    floor(keep_prob + U[0,1)) / keep_prob, drawn with the same torch.rand([B,1,1,1]) call."""
    keep = 1 - p
    u = torch.rand([batch, 1, 1, 1], dtype=torch.float32, device=device)
    return (torch.floor(keep + u) / keep).reshape(batch)


def load_pretrained_weights(model, model_name, load_fc=True):
    """Offline-safe: loads ``$EFFDET_PRETRAINED_DIR/<model_name>.pth`` when present, else leaves the
    random init in place (the reference downloads from a bucket, models/utils.py:305-328; there is
    no network here and EfficientDet.__init__ overwrites every conv weight anyway, :47-53)."""
    root = os.environ.get('EFFDET_PRETRAINED_DIR')
    if not root:
        return False
    path = os.path.join(root, model_name + '.pth')
    if not os.path.exists(path):
        return False
    sd = torch.load(path, map_location='cpu')
    if not load_fc:
        sd.pop('_fc.weight', None)
        sd.pop('_fc.bias', None)
    model.load_state_dict(sd, strict=False)
    return True
