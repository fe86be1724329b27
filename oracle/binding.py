"""ctypes bindings for the CPU checkers -- TEST INFRASTRUCTURE ONLY.

Two checkers live behind this module:

* ``Oracle``    -- oracle/_build/libvit_oracle.so, our C restatement
  (oracle/vit_oracle.c) of the reference's sequential forward.
* ``Reference`` -- oracle/_ref/libvit_ref{224,384}.so, the reference's own
  ViT_seq.c compiled unmodified by oracle/Makefile (entry point ``ViT_seq``,
  /root/reference/MulticoreMainProject/ViT_seq.c:402, and its stage functions).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this file; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libvit_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")

NBLOBS = 152
_f32p = C.POINTER(C.c_float)


def build(quiet: bool = True) -> None:
    """Compile the checkers (oracle always; _ref only where /root/reference exists)."""
    subprocess.run(["make", "-C", HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _fp(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_f32p)


class ImageData(C.Structure):
    """Mirror of the reference struct, Network.h:7-14."""
    _fields_ = [("n", C.c_int), ("c", C.c_int), ("h", C.c_int), ("w", C.c_int), ("data", _f32p)]


class Network(C.Structure):
    """Mirror of the reference struct, Network.h:19-23."""
    _fields_ = [("data", _f32p), ("size", C.c_size_t)]


def make_image_structs(images: np.ndarray):
    """images [n,3,h,w] float32 -> (ImageData array, keepalive)."""
    n, c, h, w = images.shape
    arr = (ImageData * n)()
    keep = []
    for i in range(n):
        buf = np.ascontiguousarray(images[i], dtype=np.float32)
        keep.append(buf)
        arr[i].n, arr[i].c, arr[i].h, arr[i].w = n, c, h, w
        arr[i].data = _fp(buf)
    return arr, keep


def make_network_structs(weights):
    arr = (Network * NBLOBS)()
    keep = []
    for i, wt in enumerate(weights):
        buf = np.ascontiguousarray(wt, dtype=np.float32).reshape(-1)
        keep.append(buf)
        arr[i].data = _fp(buf)
        arr[i].size = buf.size
    return arr, keep


def make_prob_rows(n: int, classes: int = 1000):
    out = np.zeros((n, classes), dtype=np.float32)
    rows = (_f32p * n)()
    for i in range(n):
        rows[i] = out[i].ctypes.data_as(_f32p)
    return rows, out


class Oracle:
    """Our restatement (oracle/vit_oracle.c)."""

    # -D variants built by oracle/Makefile: name -> (patch, embed, heads, hidden)
    VARIANTS = {None: (16, 768, 12, 3072, 12), "b16": (16, 768, 12, 3072, 12), "b32": (32, 768, 12, 3072, 12),
                "s16": (16, 384, 6, 1536, 12), "l16": (16, 1024, 16, 4096, 24)}

    def __init__(self, variant=None):
        self.patch, self.embed, self.heads, self.hidden, self.depth = self.VARIANTS[variant]
        self.nblobs = 8 + 12 * self.depth
        so = ORACLE_SO if variant in (None, "b16") else ORACLE_SO.replace("libvit_oracle.so", f"libvit_oracle_{variant}.so")
        if not os.path.exists(so):
            build()
        self.lib = C.CDLL(so)
        L = self.lib
        L.vit_oracle_tokens.restype = C.c_int
        L.vit_oracle_forward.restype = C.c_int
        L.vit_oracle_forward.argtypes = [_f32p, C.c_int, C.c_int, C.POINTER(_f32p), _f32p, _f32p, _f32p]
        L.vit_oracle_set_threads.argtypes = [C.c_int]
        L.vit_oracle_get_threads.restype = C.c_int

    def set_threads(self, n: int):
        self.lib.vit_oracle_set_threads(n)

    def threads(self) -> int:
        return self.lib.vit_oracle_get_threads()

    @staticmethod
    def _wptrs(weights):
        """array of blob pointers; a None entry becomes NULL (the C side rejects it)"""
        keep = [None if w is None else np.ascontiguousarray(w, dtype=np.float32).reshape(-1) for w in weights]
        arr = (_f32p * len(keep))()
        for i, w in enumerate(keep):
            if w is not None:
                arr[i] = _fp(w)
        return arr, keep

    def forward(self, images: np.ndarray, weights, want_logits=True, want_stages=False):
        """images [n,3,img,img] -> dict(probs, logits, stages)."""
        images = np.ascontiguousarray(images, dtype=np.float32)
        n, c, h, w = images.shape
        assert c == 3 and h == w
        if len(weights) != self.nblobs:
            raise ValueError(f"need {self.nblobs} weight blobs, got {len(weights)}")
        T = self.lib.vit_oracle_tokens(h)
        probs = np.zeros((n, 1000), np.float32)
        logits = np.zeros((n, 1000), np.float32) if want_logits else None
        stages = np.zeros((self.depth + 1, T, self.embed), np.float32) if want_stages else None
        wp, keep = self._wptrs(weights)
        rc = self.lib.vit_oracle_forward(_fp(images), n, h, wp, _fp(probs),
                                         _fp(logits) if want_logits else None,
                                         _fp(stages) if want_stages else None)
        if rc != 0:
            raise ValueError("vit_oracle_forward rejected its arguments")
        return {"probs": probs, "logits": logits, "stages": stages}

    # --- stage functions (for per-kernel parity tests) ---
    def layer_norm(self, x, g, b):
        x = np.ascontiguousarray(x, np.float32)
        y = np.empty_like(x)
        self.lib.vit_oracle_layer_norm(_fp(x), _fp(y), C.c_int(x.shape[0]), _fp(np.ascontiguousarray(g, np.float32)),
                                       _fp(np.ascontiguousarray(b, np.float32)))
        return y

    def linear(self, x, w, b, gelu=False):
        x = np.ascontiguousarray(x, np.float32)
        w = np.ascontiguousarray(w, np.float32)
        b = np.ascontiguousarray(b, np.float32)
        t, k = x.shape
        n = w.shape[0]
        y = np.empty((t, n), np.float32)
        self.lib.vit_oracle_linear(_fp(x), _fp(y), C.c_int(t), C.c_int(k), C.c_int(n), _fp(w), _fp(b), C.c_int(int(gelu)))
        return y

    def attention_core(self, q, k, v):
        q, k, v = (np.ascontiguousarray(a, np.float32) for a in (q, k, v))
        o = np.empty_like(q)
        self.lib.vit_oracle_attention_core(_fp(q), _fp(k), _fp(v), _fp(o), C.c_int(q.shape[0]))
        return o

    def mha(self, x, w_in, b_in, w_out, b_out):
        x = np.ascontiguousarray(x, np.float32)
        y = np.empty_like(x)
        args = [np.ascontiguousarray(a, np.float32) for a in (w_in, b_in, w_out, b_out)]
        self.lib.vit_oracle_mha(_fp(x), _fp(y), C.c_int(x.shape[0]), *[_fp(a) for a in args])
        return y

    def patch_embed(self, image, cls, conv_w, conv_b, pos):
        image = np.ascontiguousarray(image, np.float32)
        img = image.shape[-1]
        T = self.lib.vit_oracle_tokens(img)
        out = np.empty((T, self.embed), np.float32)
        args = [np.ascontiguousarray(a, np.float32) for a in (cls, conv_w, conv_b, pos)]
        self.lib.vit_oracle_patch_embed(_fp(image), C.c_int(img), *[_fp(a) for a in args], _fp(out))
        return out

    def encoder(self, x, layer_weights):
        x = np.ascontiguousarray(x, np.float32)
        y = np.empty_like(x)
        wp, keep = self._wptrs(layer_weights)
        self.lib.vit_oracle_encoder(_fp(x), _fp(y), C.c_int(x.shape[0]), wp)
        return y

    def softmax(self, logits):
        logits = np.ascontiguousarray(logits, np.float32)
        p = np.empty_like(logits)
        self.lib.vit_oracle_softmax(_fp(logits), _fp(p), C.c_int(logits.size))
        return p


class Reference:
    """The reference's own ViT_seq.c, compiled unmodified into oracle/_ref."""

    def __init__(self, img: int = 224, variant=None):
        """variant: None (the unmodified macros), "b32" (patch_size 32) or "s16" (embed_dim 384, num_heads 6)"""
        path = os.path.join(REF_DIR, f"libvit_ref{img}.so" if variant in (None, "b16") else f"libvit_ref{img}_{variant}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.img = img
        self.lib = C.CDLL(path)
        self.lib.ViT_seq.argtypes = [C.POINTER(ImageData), C.POINTER(Network), C.POINTER(_f32p)]
        self.lib.ViT_seq.restype = None

    @staticmethod
    def available(img: int = 224, variant=None) -> bool:
        name = f"libvit_ref{img}.so" if variant in (None, "b16") else f"libvit_ref{img}_{variant}.so"
        return os.path.exists(os.path.join(REF_DIR, name))

    def forward(self, images: np.ndarray, weights) -> np.ndarray:
        """ViT_seq(image, networks, probabilities) -> probs [n,1000]."""
        images = np.ascontiguousarray(images, dtype=np.float32)
        assert images.shape[2] == self.img
        imgs, k1 = make_image_structs(images)
        nets, k2 = make_network_structs(weights)
        rows, out = make_prob_rows(images.shape[0])
        self.lib.ViT_seq(imgs, nets, rows)
        return out

    def linear(self, x, w, b):
        """linear_layer_seq, ViT_seq.c:295 (Network structs passed by value)."""
        x = np.ascontiguousarray(x, np.float32)
        w = np.ascontiguousarray(w, np.float32)
        b = np.ascontiguousarray(b, np.float32)
        t, k = x.shape
        n = w.shape[0]
        y = np.empty((t, n), np.float32)
        f = self.lib.linear_layer_seq
        f.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int, Network, Network]
        f.restype = None
        f(_fp(x), _fp(y), t, k, n, Network(_fp(w.reshape(-1)), w.size), Network(_fp(b), b.size))
        return y

    def layer_norm(self, x, g, b):
        x = np.ascontiguousarray(x, np.float32)
        g = np.ascontiguousarray(g, np.float32)
        b = np.ascontiguousarray(b, np.float32)
        y = np.empty_like(x)
        f = self.lib.layer_norm_seq
        f.argtypes = [_f32p, _f32p, Network, Network]
        f.restype = None
        f(_fp(x), _fp(y), Network(_fp(g), g.size), Network(_fp(b), b.size))
        return y

    def mha(self, x, w_in, b_in, w_out, b_out):
        x = np.ascontiguousarray(x, np.float32)
        args = [np.ascontiguousarray(a, np.float32).reshape(-1) for a in (w_in, b_in, w_out, b_out)]
        y = np.empty_like(x)
        f = self.lib.multihead_attn_seq
        f.argtypes = [_f32p, _f32p, Network, Network, Network, Network]
        f.restype = None
        f(_fp(x), _fp(y), *[Network(_fp(a), a.size) for a in args])
        return y
