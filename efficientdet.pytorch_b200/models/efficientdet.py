"""``EfficientDet`` -- same constructor, attributes, call convention and state-dict schema as the
reference's ``models/efficientdet.py``; everything under ``forward`` is sm_100a kernels.

  train :  model([images[B,3,H,W], annotations[B,G,5]]) -> (cls_loss[1], reg_loss[1])   (:57-68)
  eval  :  model(image[1,3,H,W]) -> [scores[K], classes[K] int64, boxes[K,4]]            (:69-86)
"""
import math

import torch
import torch.nn as nn

from . import _ops
from .bifpn import BIFPN
from .efficientnet import EfficientNet
from .losses import FocalLoss
from .module import Anchors, BBoxTransform, ClipBoxes
from .retinahead import RetinaHead

MODEL_MAP = {'efficientdet-d%d' % i: 'efficientnet-b%d' % min(i, 6) for i in range(8)}


class EfficientDet(nn.Module):
    def __init__(self, num_classes, network='efficientdet-d0', D_bifpn=3, W_bifpn=88, D_class=3, is_training=True,
                 threshold=0.01, iou_threshold=0.5):
        super().__init__()
        self.backbone = EfficientNet.from_pretrained(MODEL_MAP[network])
        self.is_training = is_training
        self.neck = BIFPN(in_channels=self.backbone.get_list_features()[-5:], out_channels=W_bifpn,
                          stack=D_bifpn, num_outs=5)
        self.bbox_head = RetinaHead(num_classes=num_classes, in_channels=W_bifpn)
        self.anchors = Anchors()
        self.regressBoxes = BBoxTransform()
        self.clipBoxes = ClipBoxes()
        self.threshold = threshold
        self.iou_threshold = iou_threshold
        # the reference re-initialises every conv in the model, backbone included (:47-53)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2. / n))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
        self.freeze_bn()
        self.criterion = FocalLoss()

    def forward(self, inputs):
        if self.is_training:
            inputs, annotations = inputs
        feats = self.extract_feat_nhwc(inputs)
        classification, regression = self.bbox_head.forward_concat_nhwc(feats)
        anchors = self.anchors(inputs)
        if self.is_training:
            return self.criterion(classification, regression, anchors, annotations)
        det = _ops.detect_image0(classification, regression, anchors, inputs.shape[2], inputs.shape[3],
                                 self.threshold, self.iou_threshold)
        if det is None:
            print('No boxes to NMS')
            return [torch.zeros(0), torch.zeros(0), torch.zeros(0, 4)]
        return det

    def freeze_bn(self):
        """BatchNorm always runs on its running statistics (reference :88-92); here that is structural
        -- BN is folded into the conv epilogues -- the call keeps the modules in eval mode for parity."""
        for layer in self.modules():
            if isinstance(layer, nn.BatchNorm2d):
                layer.eval()

    def extract_feat_nhwc(self, img):
        return self.neck.forward_nhwc(self.backbone.extract_features_nhwc(img)[-5:])

    def extract_feat(self, img):
        return tuple(_ops.to_nchw_view(t) for t in self.extract_feat_nhwc(img))
