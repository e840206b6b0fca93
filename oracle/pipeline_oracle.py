"""TEST INFRASTRUCTURE (oracle): NumPy restatement of the steps either side of the hot path in the reference, used only
by tests/ to check the device kernels of efficientdet.pytorch_b200/csrc/pipeline.cu.  Pinned against the reference's own
code by tests/golden/make_pipeline_golden.py (which executes the class / function bodies cut out of
/root/reference/datasets/augmentation.py and /root/reference/eval.py) -> tests/golden/pipeline_*.npz.

  normalize_pad_collate : Normalizer (datasets/augmentation.py:141-150) -> Augmenter flip (:118-138) -> zero pad to the
                          common size (Resizer without the cv2.resize, :111-113) -> collater (:69-91) -> .float() (train.py:105)
  select_detections     : eval.py:108-128 (score threshold, argsort(-scores)[:max_detections], per-label split)
"""
import numpy as np

MEAN = np.array([[[0.485, 0.456, 0.406]]])          # datasets/augmentation.py:144-145 (float64)
STD = np.array([[[0.229, 0.224, 0.225]]])


def normalize_pad_collate(images_u8, annots, flips, common_size, scales=None):
    """images_u8: list of [h,w,3] uint8 (h,w <= common_size); annots: list of [n,5] float64; flips: list of bool;
    scales: per-image float64 box scale (Resizer's `scale`, None = 1).  -> (float32 [B,3,S,S], float32 [B,G,5])"""
    B = len(images_u8)
    imgs, anns = [], []
    for b in range(B):
        img = (images_u8[b].astype(np.float32) - MEAN) / STD                 # Normalizer: float64 result
        ann = annots[b].astype(np.float64).copy()
        if flips[b]:                                                          # Augmenter
            img = img[:, ::-1, :]
            cols = img.shape[1]
            x1, x2 = ann[:, 0].copy(), ann[:, 2].copy()
            ann[:, 0] = cols - x2
            ann[:, 2] = cols - x1
        new = np.zeros((common_size, common_size, 3))                         # Resizer: pad (no cv2.resize here)
        new[0:img.shape[0], 0:img.shape[1]] = img
        if scales is not None:
            ann[:, :4] *= scales[b]
        imgs.append(new)
        anns.append(ann)
    stacked = np.stack(imgs, axis=0)                                          # collater
    G = max(a.shape[0] for a in anns)
    if G > 0:
        pad = np.ones((B, G, 5), dtype=np.float32) * -1
        for b, a in enumerate(anns):
            if a.shape[0] > 0:
                pad[b, :a.shape[0], :] = a
    else:
        pad = np.ones((B, 1, 5), dtype=np.float32) * -1
    return stacked.transpose(0, 3, 1, 2).astype(np.float32), pad              # permute(0,3,1,2) + .float()


def select_detections(scores, labels, boxes, scale, score_threshold, max_detections, num_classes):
    """eval.py:105-128 -> list over labels of [k,5] arrays (x1,y1,x2,y2,score).  np.argsort(-scores) is evaluated as a
    STABLE sort here (the reference's default quicksort is unspecified on ties)."""
    boxes = boxes.astype(np.float32).copy()
    boxes /= scale
    indices = np.where(scores > score_threshold)[0]
    out = [np.zeros((0, 5)) for _ in range(num_classes)]
    if indices.shape[0] > 0:
        sc = scores[indices]
        order = np.argsort(-sc, kind='stable')[:max_detections]
        image_boxes = boxes[indices[order], :]
        image_scores = sc[order]
        image_labels = labels[indices[order]]
        det = np.concatenate([image_boxes, np.expand_dims(image_scores, axis=1), np.expand_dims(image_labels, axis=1)], axis=1)
        for label in range(num_classes):
            out[label] = det[det[:, -1] == label, :-1]
    return out
