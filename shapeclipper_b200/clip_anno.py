"""CLIP nearest-neighbour annotation end to end on the GPU, with the reference's file format (CLIP_anno.py):

    preprocess -> encode_image -> F.normalize -> all-pairs cosine top-k (or opt.thres sampling) -> CSV

`save_anno` writes exactly what NN_annotator.save_anno does (CLIP_anno.py:98-127: header Query, Top_1..Top_{k-1}, Top_1_score..;
the query's own index at rank 0 dropped; scores as '{:.4f}'; rows sorted by the query path) and `load_anno` parses it the way the
training dataset does (data/pix3d.py:95-108), so a bank refreshed here is read by the reference's loader unchanged.
"""
import csv
import os

import torch

from . import clip as scclip


def save_anno(anno_root, category_name, split, labels, index_topk, value_topk, k_nearest=6, label2path=None):
    """labels: N relative image paths (or labels mapped by label2path('', label)[0], as in the reference); index_topk [N][k],
    value_topk [N,k] from calc_matches. Returns the csv path."""
    label2path = label2path or (lambda root, label: (os.path.join(root, label), None))
    csv_path = os.path.join(anno_root, "{}_{}.csv".format(category_name, split))
    os.makedirs(anno_root, exist_ok=True)
    header = ["Query"] + ["Top_{}".format(i) for i in range(1, k_nearest)] + ["Top_{}_score".format(i) for i in range(1, k_nearest)]
    idx = index_topk.cpu().tolist() if isinstance(index_topk, torch.Tensor) else [list(map(int, r)) for r in index_topk]
    val = value_topk.cpu().tolist() if isinstance(value_topk, torch.Tensor) else [list(map(float, r)) for r in value_topk]
    rows = []
    for i, label in enumerate(labels):
        row = [label2path("", label)[0]]
        row += [label2path("", labels[j])[0] for j in idx[i][1:]]
        row += ["{:.4f}".format(v) for v in val[i][1:]]
        rows.append(row)
    rows.sort(key=lambda r: r[0])
    with open(csv_path, "w") as f:
        w = csv.writer(f)
        w.writerow(header)
        w.writerows(rows)
    return csv_path


def load_anno(csv_path, k_nearest=5, name_from_path=None):
    """{key(query): [key(neighbour) x k_nearest]} as data/pix3d.py:95-108 builds it; key = name_from_path(path) (identity by default)."""
    key = name_from_path or (lambda p: p)
    with open(csv_path, "r") as f:
        rows = list(csv.reader(f))[1:]
    return {key(r[0]): [key(p) for p in r[1:1 + k_nearest]] for r in rows}


@torch.no_grad()
def annotate(images, labels, anno_root, category_name, split, model=None, k_nearest=6, thres=None, batch_size=32, device="cuda",
             precision="split"):
    """images: iterable of PIL images or an [N,3,h,w] float tensor in [0,1]. Encodes them in batches (CLIP_anno.py:151-170), finds
    each image's neighbours and writes the annotation file. Returns (csv path, indices, values)."""
    if model is None:
        model, _ = scclip.load("ViT-L/14", device, precision=precision)       # the reference's model (CLIP_anno.py:16)
    feats, batch = [], []

    def flush():
        if batch:
            x = torch.stack(batch).to(device)
            feats.append(torch.nn.functional.normalize(model.encode_image(x).float(), dim=-1))
            batch.clear()
    for im in images:
        t = scclip.preprocess(im)
        batch.append(t.reshape(3, t.shape[-2], t.shape[-1]).cpu())
        if len(batch) == batch_size:
            flush()
    flush()
    f = torch.cat(feats, 0)
    idx, val = scclip.calc_matches(f, k_nearest=k_nearest, thres=thres)
    return save_anno(anno_root, category_name, split, labels, idx, val, k_nearest=k_nearest), idx, val
