"""vit-with-opencl_b200 -- B200-native ViT-B/16 inference engine (host-side mirror).

The product is ``libvit_b200.so``: a C host (host/*.c) over hand-written sm_100a
CUDA kernels (csrc/*.cu) that exports the reference's entry point
``ViT_opencl(ImageData*, Network*, float**)`` (ViT_opencl.h:6) plus the resident
engine API of include/vit_b200.h.  This module is only a ctypes mirror of those C
interfaces for the Python tests and bench.py -- there is no Python compute path
and no CPU fallback: if the shared library is missing or no B200 is present,
calls fail loudly.

The directory name contains a hyphen, so it is imported through
``__graft_entry__.load_package()`` under the module name ``vit_with_opencl_b200``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import compare, synth  # noqa: F401  (re-export)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvit_b200.so")
INCLUDE_DIR = os.path.join(os.path.dirname(HERE), "include")

FP32, BF16, FP8 = 0, 1, 2
NBLOBS, CLASSES = 152, 1000
EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RESIDUAL, EPI_PATCH_EMBED = 0, 1, 2, 3

_f32p = C.POINTER(C.c_float)


class VitError(RuntimeError):
    pass


class ImageData(C.Structure):
    """Reference struct, Network.h:7-14 (vitb200_image in include/vit_b200.h)."""
    _fields_ = [("n", C.c_int), ("c", C.c_int), ("h", C.c_int), ("w", C.c_int), ("data", _f32p)]


class Network(C.Structure):
    """Reference struct, Network.h:19-23 (vitb200_blob in include/vit_b200.h)."""
    _fields_ = [("data", _f32p), ("size", C.c_size_t)]


class GemmDesc(C.Structure):
    """vitcu_gemm_desc, include/vit_cuda_layer.h."""
    _fields_ = [("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("lda", C.c_size_t), ("epilogue", C.c_int),
                ("bias", C.c_void_p), ("residual", C.c_void_p), ("pos", C.c_void_p), ("patches", C.c_int),
                ("tokens", C.c_int), ("out_bf16", C.c_int), ("ldc", C.c_size_t),
                ("ln_stats", C.c_void_p), ("ln_slots", C.c_int), ("ln_colsum", C.c_void_p),
                ("emit_bf16", C.c_void_p), ("emit_stats", C.c_void_p),
                ("acc_scale", C.c_float), ("out_fp8", C.c_int), ("out_scale", C.c_float), ("emit_fp8", C.c_int),
                ("emit_scale", C.c_float), ("accumulate", C.c_int)]


def build(verbose: bool = False) -> str:
    """Compile libvit_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.run(["make", "-C", HERE, "-j8", "all"], check=True,
                   stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """The C-ABI library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VitError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.vitb200_last_error.restype = C.c_char_p
        L.vitcu_last_error.restype = C.c_char_p
        L.vitb200_device_count.restype = C.c_int
        L.vitb200_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int]
        L.vitb200_create_model.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.vitb200_model_from_blobs.argtypes = [C.POINTER(Network), C.c_void_p, C.c_void_p]
        L.vitb200_destroy.argtypes = [C.c_void_p]
        L.vitb200_destroy.restype = None
        L.vitb200_load_weights.argtypes = [C.c_void_p, C.POINTER(Network)]
        L.vitb200_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.vitb200_forward_structs.argtypes = [C.c_void_p, C.POINTER(ImageData), C.c_int, C.POINTER(_f32p)]
        L.vitb200_forward_topk.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.vitb200_stage_images.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.vitb200_forward_resident.argtypes = [C.c_void_p, C.c_int]
        L.vitb200_read_probs.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.vitb200_last_forward_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.vitb200_profile_timeline.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int)]
        L.vitb200_profile_gemms.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int)]
        L.vitb200_time_resident.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float)]
        L.vitb200_set_stop_after_layer.argtypes = [C.c_void_p, C.c_int]
        L.vitb200_read_tokens.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.vitb200_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
        L.vitb200_host_free.argtypes = [C.c_void_p]
        L.vitb200_kernels_per_forward.argtypes = [C.c_void_p]
        L.vitb200_tokens.argtypes = [C.c_void_p]
        L.vitb200_embed.argtypes = [C.c_void_p]
        L.ViT_opencl.argtypes = [C.POINTER(ImageData), C.POINTER(Network), C.POINTER(_f32p)]
        L.ViT_opencl.restype = None
        L.vitb200_release_persistent.restype = None
        L.vitb200_last_call_stats.restype = None
        L.vitb200_last_call_stats.argtypes = [C.c_void_p]
        # device layer
        L.vitcu_malloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
        L.vitcu_free.argtypes = [C.c_void_p]
        L.vitcu_memcpy_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.vitcu_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.vitcu_memset.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p]
        L.vitcu_device_sync.argtypes = []
        L.vitcu_set_device.argtypes = [C.c_int]
        L.vitcu_device_info.argtypes = [C.c_int, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                        C.POINTER(C.c_size_t)]
        L.vitcu_event_create.argtypes = [C.POINTER(C.c_void_p)]
        L.vitcu_event_destroy.argtypes = [C.c_void_p]
        L.vitcu_event_record.argtypes = [C.c_void_p, C.c_void_p]
        L.vitcu_event_sync.argtypes = [C.c_void_p]
        L.vitcu_event_elapsed_ms.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
        L.vitcu_launch_count.restype = C.c_ulonglong
        L.vitcu_launch_count_of.restype = C.c_ulonglong
        L.vitcu_launch_count_of.argtypes = [C.c_char_p]
        L.vitcu_launch_count_reset.restype = None
        L.vitcu_f32_to_bf16.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.vitcu_patch_gather.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.vitcu_patch_embed_tc.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                           C.c_void_p]
        L.vitcu_cls_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.vitcu_layernorm_ex.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_int, C.c_void_p]
        L.vitcu_attention_ex.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.vitcu_patch_gather_ex.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.vitcu_cls_rows_ex.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.vitcu_layernorm.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_void_p]
        L.vitcu_sgemm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(GemmDesc), C.c_void_p]
        L.vitcu_gemm_bf16.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(GemmDesc), C.c_void_p]
        L.vitcu_gemm_bf16x3.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(GemmDesc), C.c_void_p]
        L.vitcu_ln_fold_weights.argtypes = [C.c_void_p] * 7 + [C.c_int, C.c_int, C.c_void_p]
        L.vitcu_rowstats_cast.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.vitcu_gemm_bf16_emit_supported.argtypes = [C.c_int, C.c_int]
        L.vitcu_gemm_e4m3.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(GemmDesc), C.c_void_p]
        L.vitcu_absmax_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]
        L.vitcu_absmax_bf16.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.vitcu_fp8_quant_weights.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.vitcu_split3.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
        L.vitcu_patch_embed_tc_acc.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                               C.c_int, C.c_void_p]
        L.vitcu_token_rows_init.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.vitcu_split3_gelu.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
        L.vitcu_layernorm_zero.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                           C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
        L.vitcu_attention.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.vitcu_softmax_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.vitcu_topk_rows.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.vitcu_host_is_pinned.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        _lib = L
    return _lib


def _check(rc: int, layer: bool = False):
    if rc != 0:
        L = lib()
        msg = (L.vitcu_last_error() if layer else L.vitb200_last_error()).decode(errors="replace")
        raise VitError(f"error {rc}: {msg}")


def device_count() -> int:
    return lib().vitb200_device_count()


def launch_counts() -> dict:
    """launches per kernel family since the last vitcu_launch_count_reset (a captured launch counts once)"""
    L = lib()
    names = ("gemm_bf16_tc2_kernel", "gemm_bf16_tc_kernel", "attention_tc_kernel", "attention_flash_tc_kernel",
             "attention_simt_kernel", "sgemm_kernel", "layernorm_kernel", "patch_embed_tc_kernel",
             "attention_duo_tc_kernel")
    return {n: int(L.vitcu_launch_count_of(n.encode())) for n in names}


def make_network_structs(blobs):
    arr = (Network * max(NBLOBS, len(blobs)))()
    keep = []
    for i, b in enumerate(blobs):
        if b is None:
            arr[i].data = None
            arr[i].size = 0
            continue
        buf = np.ascontiguousarray(b, dtype=np.float32).reshape(-1)
        keep.append(buf)
        arr[i].data = buf.ctypes.data_as(_f32p)
        arr[i].size = buf.size
    return arr, keep


def make_image_structs(images: np.ndarray):
    n, c, h, w = images.shape
    arr = (ImageData * n)()
    keep = []
    for i in range(n):
        buf = np.ascontiguousarray(images[i], dtype=np.float32)
        keep.append(buf)
        arr[i].n, arr[i].c, arr[i].h, arr[i].w = n, c, h, w
        arr[i].data = buf.ctypes.data_as(_f32p)
    return arr, keep


def vit_opencl(images: np.ndarray, blobs) -> np.ndarray:
    """Call the drop-in entry point exactly as Main.c:54 does; returns probs [n,1000]."""
    imgs, k1 = make_image_structs(np.ascontiguousarray(images, np.float32))
    nets, k2 = make_network_structs(blobs)
    n = images.shape[0]
    out = np.zeros((n, CLASSES), np.float32)
    rows = (_f32p * n)(*[out[i].ctypes.data_as(_f32p) for i in range(n)])
    lib().ViT_opencl(imgs, nets, rows)
    return out


class CallStats(C.Structure):
    """vitb200_call_stats (include/vit_b200.h): where the last ViT_opencl call spent its time"""
    _fields_ = [("images", C.c_int), ("gpus", C.c_int), ("wall_s", C.c_double), ("create_s", C.c_double),
                ("weights_s", C.c_double), ("forward_s", C.c_double), ("teardown_s", C.c_double)]


class Model(C.Structure):
    """vitb200_model (include/vit_b200.h): run-time values of the reference's model macros"""
    _fields_ = [("img", C.c_int), ("patch", C.c_int), ("embed", C.c_int), ("depth", C.c_int), ("heads", C.c_int),
                ("hidden", C.c_int)]

    @classmethod
    def variant(cls, name: str, img: int = 224):
        patch, embed, depth, heads, hidden = synth.VARIANTS[name]
        return cls(img, patch, embed, depth, heads, hidden)


class PinnedArray:
    """float32 numpy view over pinned host memory (vitb200_host_alloc)."""

    def __init__(self, shape):
        self.shape = tuple(int(s) for s in shape)
        nbytes = int(np.prod(self.shape)) * 4
        self.ptr = C.c_void_p()
        _check(lib().vitb200_host_alloc(C.byref(self.ptr), nbytes))
        buf = (C.c_float * (nbytes // 4)).from_address(self.ptr.value)
        self.array = np.frombuffer(buf, dtype=np.float32).reshape(self.shape)

    def free(self):
        if self.ptr:
            self.array = None
            lib().vitb200_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Engine:
    """Resident engine on one GPU (include/vit_b200.h)."""

    def __init__(self, device: int = 0, img: int = 224, precision: int = BF16, max_batch: int = 256, model=None):
        self.h = C.c_void_p()
        self.img, self.precision, self.max_batch = img, precision, max_batch
        if model is None:
            _check(lib().vitb200_create(C.byref(self.h), device, img, precision, max_batch))
        else:  # a Model, or a variant name from synth.VARIANTS
            m = Model.variant(model, img) if isinstance(model, str) else model
            self.img = m.img
            _check(lib().vitb200_create_model(C.byref(self.h), device, C.byref(m), precision, max_batch))
        self.tokens = lib().vitb200_tokens(self.h)
        self.embed = lib().vitb200_embed(self.h)

    def load_weights(self, blobs):
        nets, keep = make_network_structs(blobs)
        _check(lib().vitb200_load_weights(self.h, nets))

    def forward(self, images: np.ndarray, want_logits: bool = False):
        """images: float32 [n,3,img,img] (numpy array, pinned or not) -> probs (and logits)."""
        assert images.dtype == np.float32 and images.flags["C_CONTIGUOUS"]
        n = images.shape[0]
        probs = np.empty((n, CLASSES), np.float32)
        logits = np.empty((n, CLASSES), np.float32) if want_logits else None
        _check(lib().vitb200_forward(self.h, images.ctypes.data, n, probs.ctypes.data,
                                     logits.ctypes.data if want_logits else None))
        return (probs, logits) if want_logits else probs

    def forward_into(self, images: np.ndarray, probs: np.ndarray):
        _check(lib().vitb200_forward(self.h, images.ctypes.data, images.shape[0], probs.ctypes.data, None))

    def forward_structs(self, images: np.ndarray) -> np.ndarray:
        imgs, keep = make_image_structs(images)
        n = images.shape[0]
        out = np.zeros((n, CLASSES), np.float32)
        rows = (_f32p * n)(*[out[i].ctypes.data_as(_f32p) for i in range(n)])
        _check(lib().vitb200_forward_structs(self.h, imgs, n, rows))
        return out

    def forward_topk(self, images: np.ndarray, k: int = 1):
        """labels [n,k] int32 and their probabilities [n,k], most probable first (vitb200_forward_topk)"""
        assert images.dtype == np.float32 and images.flags["C_CONTIGUOUS"]
        n = images.shape[0]
        labels = np.zeros((n, k), np.int32)
        probs = np.zeros((n, k), np.float32)
        _check(lib().vitb200_forward_topk(self.h, images.ctypes.data, n, k, labels.ctypes.data, probs.ctypes.data))
        return labels, probs

    def stage(self, images: np.ndarray):
        assert images.dtype == np.float32 and images.flags["C_CONTIGUOUS"]
        _check(lib().vitb200_stage_images(self.h, images.ctypes.data, images.shape[0]))

    def forward_resident(self, n: int) -> float:
        _check(lib().vitb200_forward_resident(self.h, n))
        ms = C.c_float()
        _check(lib().vitb200_last_forward_ms(self.h, C.byref(ms)))
        return ms.value

    def profile_gemms(self, n: int, iters: int = 3):
        """(summed GEMM ms per forward, GEMM launches per forward), CUDA events around every launch in situ"""
        ms, k = C.c_float(), C.c_int()
        _check(lib().vitb200_profile_gemms(self.h, n, iters, C.byref(ms), C.byref(k)))
        return ms.value, k.value

    def profile_timeline(self, n: int, iters: int = 3):
        """in-situ spans (kernel + gap to the next launch) of one eager forward, averaged over `iters`:
        {kind: (ms per forward, launches)} for gemm / attention / layernorm / other"""
        ms, cnt = (C.c_float * 4)(), (C.c_int * 4)()
        acc = [0.0] * 4
        for _ in range(iters):
            _check(lib().vitb200_profile_timeline(self.h, n, ms, cnt))
            acc = [a + m for a, m in zip(acc, ms)]
        return {k: (acc[i] / iters, cnt[i]) for i, k in enumerate(("gemm", "attention", "layernorm", "other"))}

    def time_resident(self, n: int, iters: int) -> float:
        ms = C.c_float()
        _check(lib().vitb200_time_resident(self.h, n, iters, C.byref(ms)))
        return ms.value

    def read_probs(self, n: int, want_logits: bool = True):
        probs = np.empty((n, CLASSES), np.float32)
        logits = np.empty((n, CLASSES), np.float32) if want_logits else None
        _check(lib().vitb200_read_probs(self.h, n, probs.ctypes.data, logits.ctypes.data if want_logits else None))
        return probs, logits

    def stop_after_layer(self, layer: int):
        _check(lib().vitb200_set_stop_after_layer(self.h, layer))

    def read_tokens(self, n: int) -> np.ndarray:
        x = np.empty((n, self.tokens, self.embed), np.float32)
        _check(lib().vitb200_read_tokens(self.h, n, x.ctypes.data))
        return x

    @property
    def kernels_per_forward(self) -> int:
        return lib().vitb200_kernels_per_forward(self.h)

    def close(self):
        if self.h:
            lib().vitb200_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------
# thin helpers over the device layer, for per-kernel parity tests
# ---------------------------------------------------------------------------
class DeviceBuffer:
    def __init__(self, nbytes: int):
        self.ptr = C.c_void_p()
        self.nbytes = int(nbytes)
        _check(lib().vitcu_malloc(C.byref(self.ptr), self.nbytes), layer=True)

    @classmethod
    def from_numpy(cls, a: np.ndarray):
        a = np.ascontiguousarray(a)
        b = cls(a.nbytes)
        _check(lib().vitcu_memcpy_h2d(b.ptr, a.ctypes.data, a.nbytes, None), layer=True)
        _check(lib().vitcu_device_sync(), layer=True)
        return b

    def to_numpy(self, dtype, shape) -> np.ndarray:
        out = np.empty(shape, dtype=dtype)
        assert out.nbytes <= self.nbytes
        _check(lib().vitcu_device_sync(), layer=True)
        _check(lib().vitcu_memcpy_d2h(out.ctypes.data, self.ptr, out.nbytes, None), layer=True)
        _check(lib().vitcu_device_sync(), layer=True)
        return out

    def free(self):
        if self.ptr:
            lib().vitcu_free(self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def bf16_bits_to_f32(bits: np.ndarray) -> np.ndarray:
    return (bits.astype(np.uint32) << 16).view(np.float32)


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """round-to-nearest-even, like __float2bfloat16_rn"""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    r = ((u >> 16) & 1) + 0x7FFF
    return ((u + r) >> 16).astype(np.uint16)


def layer_check(rc: int):
    _check(rc, layer=True)
