"""Steps either side of the hot path: oracle/pipeline_oracle.py against the fixtures
generated from the reference's own code (tests/golden/make_pipeline_golden.py), and -- on the GPU -- the device kernels
of csrc/pipeline.cu against the same fixtures, bit for bit."""
import os

import numpy as np
import pytest
import torch

import pipeline_oracle as P

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _input_case():
    st = np.load(os.path.join(G, 'pipeline_input.npz'))
    sizes = st['sizes']
    images, off = [], 0
    for h, w in sizes:
        images.append(st['pixels'][off:off + h * w * 3].reshape(h, w, 3))
        off += h * w * 3
    annots, r = [], 0
    for n in st['ann_counts']:
        annots.append(st['ann_rows'][r:r + n])
        r += n
    return st, images, annots, [bool(f) for f in st['flips']], int(st['common'][0])


def test_oracle_input_path_equals_reference_fixture():
    st, images, annots, flips, S = _input_case()
    imgs, ann = P.normalize_pad_collate(images, annots, flips, S)
    assert imgs.dtype == np.float32 and np.array_equal(imgs, st['out_images'])
    assert ann.dtype == np.float32 and np.array_equal(ann, st['out_annots'])
    # Resizer's box scaling (float64 multiply before the float32 store)
    _, ann2 = P.normalize_pad_collate(images, annots, flips, S, scales=[0.5, 1.0, 0.731, 2.0])
    assert np.array_equal(ann2[2, :5, 0], (np.float64(st['out_annots'][2, :5, 0]) * 0 + (S - annots[2][:, 2]) * 0.731).astype(np.float32))


def test_oracle_eval_selection_equals_reference_fixture():
    st = np.load(os.path.join(G, 'pipeline_eval.npz'))
    K = int(st['num_classes'][0])
    got = P.select_detections(st['scores'], st['labels'], st['boxes'], float(st['scale'][0]), float(st['thr'][0]),
                              int(st['max_det'][0]), K)
    total = 0
    for c in range(K):
        assert np.array_equal(np.asarray(got[c]), st['label%d' % c]), c
        total += got[c].shape[0]
    assert total == 100                                        # the fixture has more than max_detections candidates


@pytest.mark.gpu
def test_device_collater_bit_exact():
    from models.pipeline import DeviceCollater
    st, images, annots, flips, S = _input_case()
    samples = [dict(img=im, annot=a, flip=f) for im, a, f in zip(images, annots, flips)]
    imgs, ann = DeviceCollater(common_size=S, device='cuda:0')(samples)
    assert imgs.dtype == torch.float32 and tuple(imgs.shape) == (4, 3, S, S)
    assert np.array_equal(imgs.cpu().numpy(), st['out_images'])
    assert np.array_equal(ann.cpu().numpy(), st['out_annots'])
    scales = [0.5, 1.0, 0.731, 2.0]
    _, want = P.normalize_pad_collate(images, annots, flips, S, scales=scales)
    for s_, sc in zip(samples, scales):
        s_['scale'] = sc
    _, ann2 = DeviceCollater(common_size=S, device='cuda:0')(samples)
    assert np.array_equal(ann2.cpu().numpy(), want)
    # a batch without any box: collater emits one all -1 row per image
    _, ann3 = DeviceCollater(common_size=S, device='cuda:0')([dict(img=images[1], annot=np.zeros((0, 5)))])
    assert tuple(ann3.shape) == (1, 1, 5) and bool((ann3 == -1).all())


@pytest.mark.gpu
def test_device_eval_selection_bit_exact():
    from models.pipeline import select_detections
    st = np.load(os.path.join(G, 'pipeline_eval.npz'))
    K = int(st['num_classes'][0])
    dev = torch.device('cuda:0')
    dets, labs, offs = select_detections(torch.from_numpy(st['scores']).to(dev), torch.from_numpy(st['labels']).to(dev),
                                         torch.from_numpy(st['boxes']).to(dev), float(st['scale'][0]), float(st['thr'][0]),
                                         int(st['max_det'][0]), K)
    offs = offs.cpu().numpy()
    dets, labs = dets.cpu().numpy(), labs.cpu().numpy()
    assert dets.shape[0] == 100 and offs[-1] == 100
    for c in range(K):
        want = st['label%d' % c]
        got = dets[offs[c]:offs[c + 1]]
        assert got.shape == want.shape and np.array_equal(got.astype(np.float64), want), c
        assert (labs[offs[c]:offs[c + 1]] == c).all()
    # nothing above the threshold -> empty result (eval.py:130-133)
    d2, l2, o2 = select_detections(torch.from_numpy(st['scores']).to(dev), torch.from_numpy(st['labels']).to(dev),
                                   torch.from_numpy(st['boxes']).to(dev), 1.0, 2.0, 100, K)
    assert d2.shape[0] == 0 and int(o2[-1]) == 0
