"""Weight blobs and synthetic inputs for the ViT-B/16 path.

The reference keeps its model as 152 raw float32 blobs ``Network/Weight_<idx>_*.bin``
in torchvision ``state_dict()`` order, loaded by ``load_weights``
(/root/reference/MulticoreMainProject/Network.c:134-218), which also rounds every
value to 6 decimals (Network.c:208-211).  36 of the blobs (every in_proj / fc1 /
fc2 weight) are missing from the reference checkout, and the GPU box has no
reference at all, so this module can produce the full set three ways:

* ``load_blobs(dir)``        -- read whatever ``Weight_<idx>_*.bin`` files exist;
* ``synthetic_blobs(...)``   -- seeded random-init blobs of the right shapes;
* ``model_blobs(dir, ...)``  -- bundled blobs where present, synthetic fill for
  the rest (the recipe SURVEY.md section 8d fixes: N(0, 0.02^2), numpy
  ``default_rng(0)`` drawn in blob-index order, then the 6-decimal rounding).
"""
from __future__ import annotations

import os
import re

import numpy as np

EMBED, HEADS, DEPTH, HIDDEN, CLASSES, PATCH, NBLOBS = 768, 12, 12, 3072, 1000, 16, 152


def tokens(img: int) -> int:
    return (img // PATCH) ** 2 + 1


# model variants with the reference's blob order: name -> (patch, embed, depth, heads, hidden)
VARIANTS = {
    "b16": (16, 768, 12, 12, 3072),   # the reference's macros (ViT_seq.c:10-17)
    "b32": (32, 768, 12, 12, 3072),   # patch_size 32
    "s16": (16, 384, 12, 6, 1536),    # embed_dim 384, num_heads 6
    "l16": (16, 1024, 24, 16, 4096),  # depth 24: beyond what the reference's unrolled encoder calls express
}


def blob_shapes(img: int = 224, variant: str = "b16"):
    """Shapes of the 8 + 12*depth blobs (152 for depth 12; index use: ViT_seq.c:437-513)."""
    patch, embed, depth, _heads, hidden = VARIANTS[variant]
    t = (img // patch) ** 2 + 1
    shapes = [(embed,), (embed, 3, patch, patch), (embed,), (t, embed)]
    for _ in range(depth):
        shapes += [(embed,), (embed,), (3 * embed, embed), (3 * embed,), (embed, embed), (embed,),
                   (embed,), (embed,), (hidden, embed), (hidden,), (embed, hidden), (embed,)]
    shapes += [(embed,), (embed,), (CLASSES, embed), (CLASSES,)]
    assert len(shapes) == 8 + 12 * depth
    return shapes


def variant_blobs(variant: str, img: int = 224, seed: int = 0, std: float = 0.02):
    """fully synthetic blobs of a model variant: N(0, std^2), LayerNorm gammas 1 + N(0, std^2), 6-decimal rounding"""
    rng = np.random.default_rng(seed)
    shapes = blob_shapes(img, variant)
    n = len(shapes)
    out = []
    for idx, shp in enumerate(shapes):
        a = rng.standard_normal(int(np.prod(shp)), dtype=np.float32) * np.float32(std)
        if idx == n - 4 or (4 <= idx < n - 4 and (idx - 4) % 12 in (0, 6)):
            a = a + np.float32(1.0)
        out.append(round6(a))
    return out


def is_ln_gamma(idx: int) -> bool:
    if idx == 148:
        return True
    return 4 <= idx < 148 and (idx - 4) % 12 in (0, 6)


def round6(x: np.ndarray) -> np.ndarray:
    """roundf(x * 1000000.0f) / 1000000.0f in float32 (Network.c:208-211)."""
    t = x.astype(np.float32) * np.float32(1000000.0)
    tr = np.trunc(t)
    r = np.where(np.abs(t - tr) >= np.float32(0.5), tr + np.sign(t), tr).astype(np.float32)
    return (r / np.float32(1000000.0)).astype(np.float32)


_NAME = re.compile(r"^Weight_(\d+)_.*\.bin$")


def load_blobs(directory: str, rounded: bool = True):
    """{idx: flat float32 array} for every Weight_<idx>_*.bin in ``directory``."""
    out = {}
    if not directory or not os.path.isdir(directory):
        return out
    for name in sorted(os.listdir(directory)):
        m = _NAME.match(name)
        if not m:
            continue
        idx = int(m.group(1))
        if 0 <= idx < NBLOBS:
            a = np.fromfile(os.path.join(directory, name), dtype=np.float32)
            out[idx] = round6(a) if rounded else a
    return out


def synthetic_blobs(img: int = 224, seed: int = 0, only=None, std: float = 0.02):
    """Seeded random-init blobs: N(0, std^2) everywhere, LayerNorm gammas 1 + N(0, std^2)."""
    rng = np.random.default_rng(seed)
    out = {}
    for idx, shp in enumerate(blob_shapes(img)):
        if only is not None and idx not in only:
            continue
        a = rng.standard_normal(int(np.prod(shp)), dtype=np.float32) * np.float32(std)
        if is_ln_gamma(idx):
            a = a + np.float32(1.0)
        out[idx] = round6(a)
    return out


def model_blobs(directory: str | None = None, img: int = 224, seed: int = 0):
    """152 flat float32 blobs: bundled where present, seeded synthetic otherwise.

    For img != 224 the bundled [197,768] position embedding does not fit and is
    replaced by a synthetic [T,768] one (``default_rng(1)``).
    """
    have = load_blobs(directory) if directory else {}
    shapes = blob_shapes(img)
    if img != 224 and 3 in have:
        del have[3]
    have = {i: a for i, a in have.items() if a.size == int(np.prod(shapes[i]))}
    missing = [i for i in range(NBLOBS) if i not in have]
    fill = synthetic_blobs(img, seed, only=set(missing))
    if img != 224:
        rng = np.random.default_rng(1)
        fill[3] = round6(rng.standard_normal(tokens(img) * EMBED, dtype=np.float32) * np.float32(0.02))
    blobs = [have[i] if i in have else fill[i] for i in range(NBLOBS)]
    return blobs


def synthetic_images(n: int, img: int = 224, seed: int = 1234) -> np.ndarray:
    """[n,3,img,img] float32 ~ N(0,1) (ImageNet-normalised pixels have that scale)."""
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n, 3, img, img), dtype=np.float32)


def load_image_file(path: str) -> np.ndarray:
    """The reference's image container: 4 x int32 (n,c,h,w) + float32 data (Network.c:26-109)."""
    with open(path, "rb") as f:
        n, c, h, w = np.fromfile(f, dtype=np.int32, count=4)
        data = np.fromfile(f, dtype=np.float32, count=int(n) * int(c) * int(h) * int(w))
    return data.reshape(int(n), int(c), int(h), int(w))


def write_image_file(path: str, images: np.ndarray) -> None:
    images = np.ascontiguousarray(images, dtype=np.float32)
    with open(path, "wb") as f:
        np.asarray(images.shape, dtype=np.int32).tofile(f)
        images.tofile(f)
