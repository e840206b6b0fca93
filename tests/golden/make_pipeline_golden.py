"""Pin oracle/pipeline_oracle.py against the reference's OWN code.  datasets/augmentation.py cannot be imported here
(albumentations is not installed), so the definitions this path uses (Normalizer, Augmenter, collater and the padding
lines of Resizer) are cut out of the reference file with `ast` and executed unmodified; eval.py's selection block
(lines 105-128, inside a function that needs a dataset + a CUDA model) is executed from its source lines with the
loop variables bound.  Outputs -> tests/golden/pipeline_input.npz, pipeline_eval.npz.

usage: python tests/golden/make_pipeline_golden.py     (authoring container only: needs /root/reference)"""
import ast
import os
import sys
import textwrap

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('EFFDET_REFERENCE', '/root/reference')
sys.path.insert(0, os.path.join(REPO, 'oracle'))
import pipeline_oracle as P  # noqa: E402

src = open(os.path.join(REF, 'datasets', 'augmentation.py')).read()
tree = ast.parse(src)
want = {'collater', 'Augmenter', 'Normalizer'}
ns = {'np': np, 'torch': torch}
for node in tree.body:
    if getattr(node, 'name', None) in want:
        exec(compile(ast.Module([node], []), 'augmentation.py', 'exec'), ns)

rng = np.random.RandomState(7)
S = 96
sizes = [(96, 64), (50, 96), (96, 96), (33, 47)]
images = [rng.randint(0, 256, size=(h, w, 3)).astype(np.uint8) for h, w in sizes]
annots = [np.concatenate([np.sort(rng.rand(n, 2) * w, axis=1)[:, [0]], rng.rand(n, 1) * h, np.sort(rng.rand(n, 2) * w, axis=1)[:, [1]],
                          rng.rand(n, 1) * h, rng.randint(0, 80, size=(n, 1)).astype(np.float64)], axis=1)
          for (h, w), n in zip(sizes, [3, 0, 5, 1])]
flips = [True, False, True, False]

# reference: Normalizer -> (forced) Augmenter -> pad lines of Resizer (scale 1: no cv2.resize) -> collater -> .float()
samples = []
for img, ann, fl in zip(images, annots, flips):
    s = ns['Normalizer']()({'img': img, 'annot': ann.copy()})
    if fl:
        s = ns['Augmenter']()(s, flip_x=2.0)                 # np.random.rand() < 2.0 -> always flips
    image, a = s['img'], s['annot']
    new_image = np.zeros((S, S, 3))                          # datasets/augmentation.py:111-112
    new_image[0:image.shape[0], 0:image.shape[1]] = image
    samples.append({'img': torch.from_numpy(new_image), 'annot': torch.from_numpy(a), 'scale': 1.0})
ref_imgs, ref_ann = ns['collater'](samples)
ref_imgs = ref_imgs.float()                                  # train.py:105
o_imgs, o_ann = P.normalize_pad_collate(images, annots, flips, S)
assert np.array_equal(ref_imgs.numpy(), o_imgs), 'oracle != reference (images)'
assert np.array_equal(ref_ann.numpy(), o_ann), 'oracle != reference (annotations)'
flat = np.concatenate([im.reshape(-1) for im in images])
np.savez_compressed(os.path.join(HERE, 'pipeline_input.npz'), pixels=flat, sizes=np.array(sizes, dtype=np.int32),
                    flips=np.array(flips, dtype=np.uint8), common=np.array([S]), ann_rows=np.concatenate(annots, axis=0),
                    ann_counts=np.array([a.shape[0] for a in annots], dtype=np.int32), out_images=ref_imgs.numpy(),
                    out_annots=ref_ann.numpy())

# eval.py:105-128 executed from the reference source with the loop variables bound
lines = open(os.path.join(REF, 'eval.py')).read().splitlines()
start = next(i for i, l in enumerate(lines) if '# correct boxes for image scale' in l)       # eval.py:107
stop = next(i for i, l in enumerate(lines) if i > start and "print('{}/{}'" in l)              # eval.py:135
block = textwrap.dedent('\n'.join(lines[start:stop]))        # scale correction ... per-label copy (both branches)
assert block.lstrip().startswith('# correct boxes') and 'all_detections[index][label]' in block, block[:80]


class _DS:
    def num_classes(self):
        return 7


n = 300
scores = rng.rand(n).astype(np.float32)
labels = rng.randint(0, 7, size=n).astype(np.int64)
boxes = (rng.rand(n, 4) * 500).astype(np.float32)
env = dict(np=np, scores=scores.copy(), labels=labels.copy(), boxes=boxes.copy(), scale=0.7371, score_threshold=0.35,
           max_detections=100, dataset=_DS(), index=0, all_detections=[[None] * 7])
exec(block, env)
ref = env['all_detections'][0]
got = P.select_detections(scores, labels, boxes, 0.7371, 0.35, 100, 7)
for c in range(7):
    assert np.array_equal(np.asarray(ref[c]), np.asarray(got[c])), 'oracle != reference (eval label %d)' % c
np.savez_compressed(os.path.join(HERE, 'pipeline_eval.npz'), scores=scores, labels=labels, boxes=boxes,
                    scale=np.array([0.7371]), thr=np.array([0.35]), max_det=np.array([100]), num_classes=np.array([7]),
                    **{'label%d' % c: np.asarray(ref[c]) for c in range(7)})
print('pipeline goldens written; oracle == reference code on both')
