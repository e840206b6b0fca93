"""EfficientNet backbone of the detector, B200-native.

Public surface and state-dict keys follow the reference's ``models/efficientnet.py``
(``EfficientNet.from_name / from_pretrained / get_list_features / extract_features``; ``MBConvBlock`` children
``_expand_conv, _bn0, _depthwise_conv, _bn1, _se_reduce, _se_expand, _project_conv, _bn2``).  The layers are parameter
holders: one MBConv block is ONE autograd node whose forward is five launches

    expand GEMM + BN + swish | depthwise + BN + swish | SE mean | SE gate | project GEMM x gate + BN (+ skip)

instead of ~15 ATen ops (reference :75-105).  BatchNorm is frozen (eval statistics, reference
``models/efficientdet.py:88-92``) and folded into the conv epilogues; swish keeps only the pre-activation.
"""
from torch import nn

from . import _ops
from .utils import (MemoryEfficientSwish, drop_connect_scale, efficientnet_params, get_model_params,
                    get_same_padding_conv2d, load_pretrained_weights, round_filters, round_repeats)

_VALID_NAMES = ['efficientnet-b%d' % i for i in range(8)]


def _frozen_bn(bn):
    """(gamma, beta, running_mean, running_var) of a holder BatchNorm2d, in kernel-argument order.
    The kernels implement the frozen (running-statistics) BatchNorm the detector is trained with
    (models/efficientdet.py:88-92, train.py:100-102); a BatchNorm left in batch-statistics mode would silently
    compute something else than the reference, so it is refused."""
    if bn.training:
        raise _ops.N.EffdetNativeError(
            'BatchNorm2d is in batch-statistics (training) mode; the B200 hot path implements the frozen BatchNorm '
            'the reference trains with -- call freeze_bn() (or .eval()) after .train(), as train.py:100-102 does')
    return [bn.weight, bn.bias, bn.running_mean, bn.running_var]


def _make_bn(channels, global_params):
    return nn.BatchNorm2d(num_features=channels, momentum=1 - global_params.batch_norm_momentum,
                          eps=global_params.batch_norm_epsilon)


class MBConvBlock(nn.Module):
    """Mobile inverted bottleneck with squeeze-excite.  ``forward(inputs, drop_connect_rate)`` takes and returns
    logical NCHW tensors (channels_last memory); ``forward_nhwc`` is the zero-copy internal entry."""

    def __init__(self, block_args, global_params):
        super().__init__()
        self._block_args = block_args
        self._bn_mom = 1 - global_params.batch_norm_momentum
        self._bn_eps = global_params.batch_norm_epsilon
        self.id_skip = block_args.id_skip
        ratio = block_args.se_ratio
        self.has_se = ratio is not None and 0 < ratio <= 1
        if not self.has_se:
            raise NotImplementedError('the B200 MBConv kernel chain always includes squeeze-excite '
                                      '(every EfficientDet block has se_ratio=0.25)')
        conv = get_same_padding_conv2d(image_size=global_params.image_size)
        c_in, c_out = block_args.input_filters, block_args.output_filters
        c_mid = c_in * block_args.expand_ratio
        if block_args.expand_ratio != 1:                     # 1x1 expansion (absent in the first stage)
            self._expand_conv = conv(in_channels=c_in, out_channels=c_mid, kernel_size=1, bias=False)
            self._bn0 = _make_bn(c_mid, global_params)
        self._depthwise_conv = conv(in_channels=c_mid, out_channels=c_mid, groups=c_mid,
                                    kernel_size=block_args.kernel_size, stride=block_args.stride, bias=False)
        self._bn1 = _make_bn(c_mid, global_params)
        c_sq = max(1, int(c_in * ratio))                     # squeeze width is relative to the block INPUT
        self._se_reduce = conv(in_channels=c_mid, out_channels=c_sq, kernel_size=1)
        self._se_expand = conv(in_channels=c_sq, out_channels=c_mid, kernel_size=1)
        self._project_conv = conv(in_channels=c_mid, out_channels=c_out, kernel_size=1, bias=False)
        self._bn2 = _make_bn(c_out, global_params)
        self._swish = MemoryEfficientSwish()

    # -- static description handed to the kernels ------------------------------------------------------------
    def _has_skip(self):
        # reference :100 -- `stride == 1` is False for the list [1] carried by the first block of a stage, so only
        # the repeat blocks of a stage get the residual (and drop-connect)
        a = self._block_args
        return bool(self.id_skip and a.stride == 1 and a.input_filters == a.output_filters)

    def _kernel_cfg(self):
        a = self._block_args
        stride = a.stride[0] if isinstance(a.stride, (list, tuple)) else a.stride
        left, right, top, bottom = self._depthwise_conv.same_pad
        return dict(k=a.kernel_size, s=stride, eps=self._bn_eps, expand=a.expand_ratio != 1, skip=self._has_skip(),
                    pad_t=top, pad_l=left, pad_h=top + bottom, pad_w=left + right)

    def _params(self):
        chain = []
        if self._block_args.expand_ratio != 1:
            chain += [self._expand_conv.weight] + _frozen_bn(self._bn0)
        chain += [self._depthwise_conv.weight] + _frozen_bn(self._bn1)
        chain += [self._se_reduce.weight, self._se_reduce.bias, self._se_expand.weight, self._se_expand.bias]
        chain += [self._project_conv.weight] + _frozen_bn(self._bn2)
        return chain

    # -- execution ------------------------------------------------------------------------------------------
    def forward_nhwc(self, x, drop_connect_rate=None):
        cfg = self._kernel_cfg()
        keep_scale = None
        if cfg['skip'] and drop_connect_rate and self.training:
            keep_scale = drop_connect_scale(x.shape[0], drop_connect_rate, x.device)
        return _ops.MBConvFn.apply(x, keep_scale, cfg, *self._params())

    def forward(self, inputs, drop_connect_rate=None):
        x = _ops.to_nhwc(inputs, 'MBConvBlock input')
        return _ops.to_nchw_view(self.forward_nhwc(x, drop_connect_rate))

    def set_swish(self, memory_efficient=True):
        """API parity; swish is always fused (and always 'memory efficient': only z is kept)."""


class EfficientNet(nn.Module):
    """Feature extractor: ``forward(images[B,3,H,W])`` -> list of the 7 stage outputs (strides 2..128)."""

    def __init__(self, blocks_args=None, global_params=None):
        super().__init__()
        assert isinstance(blocks_args, list), 'blocks_args should be a list'
        assert len(blocks_args) > 0, 'block args must be greater than 0'
        self._global_params = global_params
        self._blocks_args = blocks_args
        conv = get_same_padding_conv2d(image_size=global_params.image_size)

        stem_width = round_filters(32, global_params)
        self._conv_stem = conv(3, stem_width, kernel_size=3, stride=2, bias=False)
        self._bn0 = _make_bn(stem_width, global_params)

        # one MBConvBlock per repeat; the stored per-stage args end up describing the REPEAT blocks
        # (input == output filters, stride 1) exactly as the reference leaves them (:149-166)
        self._blocks = nn.ModuleList([])
        for stage in range(len(self._blocks_args)):
            args = self._blocks_args[stage]
            args = args._replace(input_filters=round_filters(args.input_filters, global_params),
                                 output_filters=round_filters(args.output_filters, global_params),
                                 num_repeat=round_repeats(args.num_repeat, global_params))
            self._blocks.append(MBConvBlock(args, global_params))
            if args.num_repeat > 1:
                args = args._replace(input_filters=args.output_filters, stride=1)
            self._blocks_args[stage] = args
            self._blocks.extend(MBConvBlock(args, global_params) for _ in range(args.num_repeat - 1))

        # classifier tail: never executed by the detector, but part of the checkpoint schema
        tail_in = self._blocks_args[-1].output_filters
        tail_out = round_filters(1280, global_params)
        self._conv_head = conv(tail_in, tail_out, kernel_size=1, bias=False)
        self._bn1 = _make_bn(tail_out, global_params)
        self._avg_pooling = nn.AdaptiveAvgPool2d(1)
        self._dropout = nn.Dropout(global_params.dropout_rate)
        self._fc = nn.Linear(tail_out, global_params.num_classes)
        self._swish = MemoryEfficientSwish()

    def set_swish(self, memory_efficient=True):
        for blk in self._blocks:
            blk.set_swish(memory_efficient)

    def get_list_features(self):
        return [args.output_filters for args in self._blocks_args]

    # -- execution ------------------------------------------------------------------------------------------
    def extract_features_nhwc(self, inputs):
        _ops.check_cuda_f32(inputs, 'EfficientNet input')
        stem_bn = self._bn0
        _frozen_bn(stem_bn)
        x = _ops.StemFn.apply(inputs, self._conv_stem.weight, stem_bn.weight, stem_bn.bias, stem_bn.running_mean,
                              stem_bn.running_var, stem_bn.eps)
        total = len(self._blocks)
        base_rate = self._global_params.drop_connect_rate
        outputs, stage, done_in_stage = [], 0, 0
        for position, blk in enumerate(self._blocks):
            rate = base_rate * float(position) / total if base_rate else base_rate     # reference :200-203
            x = blk.forward_nhwc(x, drop_connect_rate=rate)
            done_in_stage += 1
            if done_in_stage == self._blocks_args[stage].num_repeat:                    # last block of the stage
                outputs.append(x)
                stage, done_in_stage = stage + 1, 0
        return outputs

    def extract_features(self, inputs):
        return [_ops.to_nchw_view(t) for t in self.extract_features_nhwc(inputs)]

    def forward(self, inputs):
        return self.extract_features(inputs)

    # -- construction helpers -----------------------------------------------------------------------------------
    @classmethod
    def _check_model_name_is_valid(cls, model_name, also_need_pretrained_weights=False):
        allowed = _VALID_NAMES[:4] if also_need_pretrained_weights else _VALID_NAMES
        if model_name not in allowed:
            raise ValueError('model_name should be one of: ' + ', '.join(allowed))

    @classmethod
    def from_name(cls, model_name, override_params=None):
        cls._check_model_name_is_valid(model_name)
        blocks_args, global_params = get_model_params(model_name, override_params)
        return cls(blocks_args, global_params)

    @classmethod
    def from_pretrained(cls, model_name, num_classes=1000):
        net = cls.from_name(model_name, override_params={'num_classes': num_classes})
        load_pretrained_weights(net, model_name, load_fc=(num_classes == 1000))
        return net

    @classmethod
    def get_image_size(cls, model_name):
        cls._check_model_name_is_valid(model_name)
        return efficientnet_params(model_name)[2]
