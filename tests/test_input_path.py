"""Input path: the oracle's restatement of the reference transform against PIL + torchvision (the third-party code the
reference calls, mini_imagenet.py:50-56), the coefficient tables the kernel uses, the episode sampler against the
reference's own sampler; -m gpu: the kernel against the oracle, bit for bit."""
import importlib.util
import os

import numpy as np
import pytest
import torch

import sun_oracle as O
from sunb200 import input as I

REF_SAMPLER = "/root/reference/test_phase/datasets/samplers.py"


def _images(n, seed=0, size=84):
    return np.random.default_rng(seed).integers(0, 256, (n, size, size, 3), dtype=np.uint8)


def test_oracle_transform_matches_pil_and_torchvision():
    PIL = pytest.importorskip("PIL.Image")
    tv = pytest.importorskip("torchvision.transforms")
    t = tv.Compose([tv.Resize((88, 88)), tv.CenterCrop(80), tv.ToTensor(),
                    tv.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    for img in _images(4, seed=1):
        ref = t(PIL.fromarray(img))
        assert torch.equal(O.default_transform(img), ref)          # bit-exact (integer resampling, IEEE fp32 ops)
    flat = np.full((84, 84, 3), 255, dtype=np.uint8)               # saturation / rounding corner
    assert torch.equal(O.default_transform(flat), t(PIL.fromarray(flat)))


def test_resize_tables_reproduce_the_oracle():
    """The (first index, weights) tables of the kernel applied with plain integer numpy == the oracle's resize + crop."""
    tmin, tk = I.resize_tables(84, 88, 80)
    assert tmin.shape == (80,) and tk.shape == (80, 3) and (tk.sum(1) == (1 << 22)).all()
    img = _images(1, seed=2)[0].astype(np.int64)
    h = np.zeros((84, 80, 3), dtype=np.int64)
    for x in range(80):
        acc = np.full((84, 3), 1 << 21, dtype=np.int64)
        for j in range(3):
            if tk[x, j]:
                acc += img[:, tmin[x] + j] * tk[x, j]
        h[:, x] = np.clip(acc >> 22, 0, 255)
    v = np.zeros((80, 80, 3), dtype=np.int64)
    for y in range(80):
        acc = np.full((80, 3), 1 << 21, dtype=np.int64)
        for j in range(3):
            if tk[y, j]:
                acc += h[tmin[y] + j] * tk[y, j]
        v[y] = np.clip(acc >> 22, 0, 255)
    ks = O._pil_bilinear_coeffs(84, 88)
    ref = O._pil_resample_axis(O._pil_resample_axis(img.astype(np.uint8), ks, 1), ks, 0)[4:84, 4:84]
    assert np.array_equal(v, ref)


@pytest.mark.skipif(not os.path.exists(REF_SAMPLER), reason="reference checkout not present (GPU box)")
def test_sampler_matches_reference_sampler():
    from datasets import CategoriesSampler
    spec = importlib.util.spec_from_file_location("ref_samplers", REF_SAMPLER)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    label = np.repeat(np.arange(20), 30)
    for args in ((7, 5, 16, 1), (3, 10, 6, 8)):
        np.random.seed(12345)
        a = [b.clone() for b in ref.CategoriesSampler(label, *args)]
        np.random.seed(12345)
        b = [b.clone() for b in CategoriesSampler(label, *args)]
        assert len(a) == len(b) == args[0]
        assert all(torch.equal(x, y) for x, y in zip(a, b))          # index work: bit-exact


def test_sampler_episode_structure():
    from datasets import CategoriesSampler
    label = np.repeat(np.arange(12), 25)
    np.random.seed(3)
    for batch in CategoriesSampler(label, 4, 5, 16, ep_per_batch=2):
        grid = batch.view(2, 5, 16)
        lab = torch.as_tensor(label)[grid]
        assert (lab == lab[:, :, :1]).all()                               # one class per row
        for e in range(2):
            assert len(set(lab[e, :, 0].tolist())) == 5                   # classes without replacement
            assert len(set(grid[e].flatten().tolist())) == 80             # images without replacement


@pytest.mark.gpu
def test_preprocess_kernel_bit_exact():
    imgs = _images(37, seed=5)
    data = torch.from_numpy(imgs).cuda()
    out = I.preprocess_u8(data)
    ref = torch.stack([O.default_transform(im) for im in imgs])
    assert out.shape == (37, 3, 80, 80) and torch.equal(out.cpu(), ref)
    idx = torch.tensor([36, 0, 5, 5, 17], dtype=torch.int64)
    assert torch.equal(I.preprocess_u8(data, idx.cuda()).cpu(), ref[idx])
    with pytest.raises(IndexError):
        I.preprocess_u8(data, torch.tensor([37]).cuda())


@pytest.mark.gpu
def test_device_store_episode_matches_host_pipeline():
    """A sampler batch through the device store == the reference pipeline (transform per image, stack, split)."""
    from datasets import CategoriesSampler, DeviceImageStore
    import utils.few_shot as fs
    imgs = _images(120, seed=6)
    label = np.repeat(np.arange(8) + 3, 15)
    store = DeviceImageStore(imgs, label)
    assert store.n_classes == 8 and store.label[0] == 0
    np.random.seed(1)
    batch = next(iter(CategoriesSampler(store.label, 1, 5, 6, ep_per_batch=2)))
    x, y = store.batch(batch)
    ref = torch.stack([O.default_transform(imgs[i]) for i in batch.tolist()])
    assert torch.equal(x.cpu(), ref)
    xs, xq = fs.split_shot_query(x, 5, 1, 5, ep_per_batch=2)
    rs, rq = O.split_shot_query(ref, 5, 1, 5, 2)
    assert torch.equal(xs.cpu(), rs) and torch.equal(xq.cpu(), rq)
    assert torch.equal(y, torch.as_tensor(store.label)[batch])
