"""CPU: the pieces around the CLIP tower that make CLIP_anno.py drop in — the PIL preprocess (openai/CLIP's transform, which the
reference applies per image at data/pix3d.py:286-288) and the annotation CSV (CLIP_anno.py:98-127, read back by data/pix3d.py:95-108)."""
import csv
import os

import numpy as np
import pytest
import torch

import refharness


def _pil(h, w, seed):
    from PIL import Image
    rng = np.random.RandomState(seed)
    return Image.fromarray(rng.randint(0, 256, size=(h, w, 3), dtype=np.uint8), "RGB")


@pytest.mark.parametrize("h,w", [(224, 224), (300, 451), (500, 333), (97, 224), (224, 100)])
def test_pil_preprocess_equals_torchvision_clip_transform(h, w):
    """openai/CLIP `_transform(224)`: Resize(224, BICUBIC) -> CenterCrop(224) -> RGB -> ToTensor -> Normalize(mean, std)."""
    tv = pytest.importorskip("torchvision")
    from torchvision.transforms import CenterCrop, Compose, InterpolationMode, Normalize, Resize, ToTensor
    from shapeclipper_b200 import clip
    ref = Compose([Resize(224, interpolation=InterpolationMode.BICUBIC), CenterCrop(224), lambda im: im.convert("RGB"), ToTensor(),
                   Normalize(clip.CLIP_MEAN, clip.CLIP_STD)])
    im = _pil(h, w, h * 1000 + w)
    got, want = clip.preprocess(im), ref(im)
    assert got.shape == (3, 224, 224) and got.dtype == torch.float32
    assert torch.equal(got, want)


def test_tensor_preprocess_is_close_to_the_pil_path():
    from shapeclipper_b200 import clip
    from PIL import Image
    yy, xx = np.mgrid[0:300, 0:400]
    arr = np.stack([127 + 120 * np.sin(xx / 37.0), 127 + 120 * np.cos(yy / 29.0), (xx + yy) * 255.0 / 700], -1).astype(np.uint8)
    im = Image.fromarray(arr, "RGB")                       # a smooth image: the two bicubic kernels agree on it
    x = torch.from_numpy(np.asarray(im).copy()).permute(2, 0, 1).float().div(255)
    a, b = clip.preprocess(im), clip.preprocess(x.unsqueeze(0))[0]
    assert a.shape == b.shape
    assert float((a - b).abs().mean()) < 0.02            # different bicubic kernels (PIL vs torch antialias): same image


def _fake_matches(n, k, seed):
    g = torch.Generator().manual_seed(seed)
    idx = torch.stack([torch.cat([torch.tensor([i]), torch.randperm(n, generator=g)[:k - 1]]) for i in range(n)])
    val = torch.cat([torch.ones(n, 1), torch.rand(n, k - 1, generator=g)], 1)
    return idx, val


def test_save_anno_round_trip(tmp_path):
    from shapeclipper_b200 import clip_anno
    n, k = 23, 6
    labels = ["img_processed/chair/%04d.png" % ((i * 7) % n) for i in range(n)]           # not sorted: the writer sorts rows
    idx, val = _fake_matches(n, k, 0)
    path = clip_anno.save_anno(str(tmp_path), "chair", "train", labels, idx, val, k_nearest=k)
    assert os.path.basename(path) == "chair_train.csv"
    rows = list(csv.reader(open(path)))
    assert rows[0] == ["Query"] + ["Top_%d" % i for i in range(1, k)] + ["Top_%d_score" % i for i in range(1, k)]
    assert [r[0] for r in rows[1:]] == sorted(labels)
    d = clip_anno.load_anno(path, k_nearest=5)
    for i, lab in enumerate(labels):
        assert d[lab] == [labels[j] for j in idx[i, 1:].tolist()]
    row = next(r for r in rows[1:] if r[0] == labels[3])
    assert row[k:] == ["%.4f" % v for v in val[3, 1:].tolist()]


@pytest.mark.skipif(not refharness.reference_available(), reason="reference modules absent")
def test_save_anno_is_byte_equal_to_the_reference_writer(tmp_path):
    """NN_annotator.save_anno (CLIP_anno.py:98-127) itself, `clip` stubbed (never called), against clip_anno.save_anno."""
    import importlib
    from shapeclipper_b200 import clip_anno
    refharness.import_reference()
    refharness._stub("clip")
    anno = importlib.import_module("CLIP_anno")
    ann = anno.NN_annotator.__new__(anno.NN_annotator)
    ann.split = "val"
    n, k = 31, 6
    labels = ["chair/%03d.png" % ((i * 11) % n) for i in range(n)]
    idx, val = _fake_matches(n, k, 1)
    opt = refharness.load_reference_opt()
    opt.anno_root = str(tmp_path / "ref")
    ann.save_anno(opt, lambda root, label: (os.path.join(root, label), None), labels, [r for r in idx], val, k_nearest=k, category_set="custom")
    ours = clip_anno.save_anno(str(tmp_path / "ours"), opt.data[opt.data.dataset].cat.replace(", ", "_"), "val", labels, idx, val, k_nearest=k)
    ref_path = os.path.join(opt.anno_root, os.path.basename(ours))
    assert open(ours, "rb").read() == open(ref_path, "rb").read()


def test_shim_registers_clip_when_openai_clip_is_absent():
    import importlib.util
    import sys
    saved = {k: sys.modules.get(k) for k in ("clip", "model.renderer", "model.implicit", "chamfer_3D")}
    sys.modules.pop("clip", None)
    if importlib.util.find_spec("clip") is not None:
        if saved["clip"] is not None:
            sys.modules["clip"] = saved["clip"]
        pytest.skip("a real clip package is installed")
    try:
        from shapeclipper_b200 import shim
        shim.install()
        import clip
        assert clip.__name__ == "shapeclipper_b200.clip" and callable(clip.load) and callable(clip.preprocess)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
