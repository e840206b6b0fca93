"""Device-side versions of the two steps that sit either side of the hot path in the reference's loops
(SURVEY.md 8(f) ranks 2 and 3), as thin host wrappers over csrc/pipeline.cu:

  DeviceCollater      Normalizer -> Augmenter flip -> zero pad -> collater -> .cuda().float()
                      (datasets/augmentation.py:69-91,111-150; train.py:105-106): the DataLoader workers only decode
                      (and resize) to uint8; everything else happens in two launches on the GPU, bit-identical to NumPy.
  select_detections   eval.py:105-128 after NMS: boxes /= scale, score threshold, top-100, per-label split, without
                      the three .cpu().numpy() round trips per image.
"""
import ctypes

import numpy as np
import torch

from . import _native as N

MEAN = (0.485, 0.456, 0.406)        # datasets/augmentation.py:144-145
STD = (0.229, 0.224, 0.225)


class DeviceCollater:
    """collate_fn replacement: call with a list of samples {'img': uint8 [h,w,3] ndarray, 'annot': [n,5] float64 ndarray,
    optional 'flip': bool, optional 'scale': float}; returns (images float32 [B,3,S,S], annotations float32 [B,G,5])
    on `device` -- what train.py:105-106 hands to the model."""

    def __init__(self, common_size=512, device='cuda:0'):
        self.S = int(common_size)
        self.device = torch.device(device)
        self._mean = (ctypes.c_double * 3)(*MEAN)
        self._std = (ctypes.c_double * 3)(*STD)

    def __call__(self, samples):
        B, S, dev = len(samples), self.S, self.device
        imgs = [np.ascontiguousarray(s['img']) for s in samples]
        for im in imgs:
            if im.dtype != np.uint8 or im.ndim != 3 or im.shape[2] != 3 or im.shape[0] > S or im.shape[1] > S:
                raise N.EffdetNativeError('DeviceCollater: images must be uint8 [h,w,3] with h,w <= %d, got %s %s'
                                          % (S, im.dtype, im.shape))
        sizes = np.array([[im.shape[0], im.shape[1]] for im in imgs], dtype=np.int32)
        nbytes = [im.size for im in imgs]
        offs = np.concatenate([[0], np.cumsum(nbytes)[:-1]]).astype(np.int64)
        flat = torch.from_numpy(np.concatenate([im.reshape(-1) for im in imgs])).pin_memory()
        flips = np.array([1 if s.get('flip') else 0 for s in samples], dtype=np.uint8)
        anns = [np.asarray(s['annot'], dtype=np.float64).reshape(-1, 5) for s in samples]
        counts = np.array([a.shape[0] for a in anns], dtype=np.int32)
        row_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        G = max(int(counts.max()), 1)                                  # collater: one all -1 row when nobody has boxes
        rows = np.concatenate(anns, axis=0) if counts.sum() else np.zeros((1, 5))
        scales = np.array([float(s.get('scale', 1.0)) for s in samples], dtype=np.float64)
        d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev, non_blocking=True)   # noqa: E731
        pix_d, offs_d, hw_d, flip_d = flat.to(dev, non_blocking=True), d(offs), d(sizes), d(flips)
        rows_d, roff_d, sc_d, w_d = d(rows), d(row_off), d(scales), d(sizes[:, 1].copy())
        out = torch.empty((B, 3, S, S), device=dev, dtype=torch.float32)
        ann = torch.empty((B, G, 5), device=dev, dtype=torch.float32)
        N.call('effdet_normalize_pad', out, N.ptr(pix_d), N.ptr(offs_d), N.ptr(hw_d), N.ptr(flip_d), N.f32(out), B, S,
               self._mean, self._std, nbytes=float(flat.numel() + 4 * out.numel()))
        N.call('effdet_collate_annots', out, N.ptr(rows_d), N.ptr(roff_d), N.ptr(sc_d), N.ptr(flip_d), N.ptr(w_d), N.f32(ann),
               B, G)
        return out, ann


def select_detections(scores, labels, boxes, scale, score_threshold=0.05, max_detections=100, num_classes=80):
    """eval.py:105-128 on the device.  scores [n] f32, labels [n] i64, boxes [n,4] f32 (the model's eval-mode output).
    -> (dets [m,5] f32 grouped by label, labels [m] i32, class_offsets [num_classes+1] i32): detections of label c are
    dets[class_offsets[c]:class_offsets[c+1]], each (x1,y1,x2,y2,score) with boxes already divided by scale."""
    dev = scores.device
    n = int(scores.numel())
    dets = torch.empty((max_detections, 5), device=dev, dtype=torch.float32)
    labs = torch.empty((max_detections,), device=dev, dtype=torch.int32)
    offs = torch.empty((num_classes + 1,), device=dev, dtype=torch.int32)
    cnt = torch.empty((1,), device=dev, dtype=torch.int32)
    sc = scores.contiguous() if n else None
    N.call('effdet_eval_select', dets, N.f32(sc, 'scores'), N.ptr(labels.contiguous() if n else None),
           N.f32(boxes.contiguous() if n else None, 'boxes'), n, float(scale), float(score_threshold), int(max_detections),
           int(num_classes), N.f32(dets), N.ptr(labs), N.ptr(offs), N.ptr(cnt))
    m = int(cnt.item())
    return dets[:m], labs[:m], offs
