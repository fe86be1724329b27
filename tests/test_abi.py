"""CPU tests of the boundary: the C-ABI library loads without a GPU, exports
every symbol the headers declare, mirrors the reference structs, and fails
loudly (no CPU fallback) when asked to compute without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(vitcu_\w+|vitb200_\w+|ViT_opencl)\s*\(", src)
    # typedef'd struct names etc. never match "name(", so these are functions
    return sorted(set(names))


@pytest.mark.parametrize("header", ["vit_b200.h", "vit_cuda_layer.h"])
def test_every_declared_symbol_is_exported(pkg, header):
    L = pkg.lib()
    names = _declared_functions(header)
    assert len(names) > 10
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in include/{header} but not exported: {missing}"


def test_reference_entry_point_is_exported(pkg):
    assert hasattr(pkg.lib(), "ViT_opencl")


def test_struct_layouts_match_reference(pkg):
    # Network.h:7-14 -> 4 ints + pointer; Network.h:19-23 -> pointer + size_t
    assert C.sizeof(pkg.ImageData) == 24 and pkg.ImageData.data.offset == 16
    assert C.sizeof(pkg.Network) == 16 and pkg.Network.size.offset == 8


def test_no_cpu_fallback(pkg):
    if pkg.device_count() > 0:
        pytest.skip("a GPU is visible; this test is about the GPU-less box")
    with pytest.raises(pkg.VitError):
        pkg.Engine(0, 224, pkg.BF16, 2)


def test_argument_validation_without_gpu(pkg):
    L = pkg.lib()
    h = C.c_void_p()
    assert L.vitb200_create(C.byref(h), 0, 225, pkg.FP32, 1) != 0  # not a multiple of 16
    assert b"multiple of 16" in L.vitb200_last_error()
    assert L.vitb200_create(C.byref(h), 0, 224, 7, 1) != 0
    assert L.vitb200_create(None, 0, 224, 0, 1) != 0


def test_round6_matches_reference_loader(pkg, tmp_path):
    """synth.round6 == the rounding load_weights applies (Network.c:208-211)"""
    from oracle import binding
    if not binding.Reference.available(224):
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(0)
    raw = (rng.standard_normal(4096, dtype=np.float32) * np.float32(0.3)).astype(np.float32)
    raw[:4] = [0.4999995, -0.4999995, 1.5e-7, -2.5e-6]
    d = tmp_path / "Network"
    d.mkdir()
    raw.tofile(d / "Weight_5_some_name.bin")
    ref = binding.Reference(224).lib
    nets = (binding.Network * 152)()
    ref.load_weights.argtypes = [C.c_char_p, C.POINTER(binding.Network), C.c_int]
    ref.load_weights(str(d).encode(), nets, 152)
    assert nets[5].size == raw.size and not nets[4].data
    got = np.ctypeslib.as_array(nets[5].data, shape=(raw.size,)).copy()
    assert np.array_equal(got, pkg.synth.round6(raw))
    assert np.array_equal(pkg.synth.load_blobs(str(d))[5], got)


def test_blob_shapes(pkg):
    s224, s384 = pkg.synth.blob_shapes(224), pkg.synth.blob_shapes(384)
    assert len(s224) == 152 and s224[3] == (197, 768) and s384[3] == (577, 768)
    assert s224[6] == (2304, 768) and s224[12] == (3072, 768) and s224[14] == (768, 3072)
    assert s224[150] == (1000, 768)
    total = sum(int(np.prod(s)) for s in s224)
    assert total == 86567656


def test_bf16_helpers_roundtrip(pkg):
    x = np.array([1.0, -2.5, 3.14159, 1e-8, 65504.0, 0.1], np.float32)
    b = pkg.f32_to_bf16_bits(x)
    y = pkg.bf16_bits_to_f32(b)
    assert np.all(np.abs(y - x) <= np.abs(x) * 2.0 ** -8)
    # round-to-nearest-even on a tie
    tie = np.array([np.uint32(0x3F808000)], np.uint32).view(np.float32)
    assert pkg.f32_to_bf16_bits(tie)[0] == 0x3F80


# ---------------------------------------------------------------- host logic that needs no GPU
def test_stager_gathers_blocks_on_all_threads(pkg):
    """host/vit_stage.c: the thread pool that gathers the reference's per-image buffers (R/Network.c:84-105)
    into pinned staging -- here into plain memory: byte-exact for block sizes that are not a multiple of
    the 128 KB piece, for per-image pointers and for one contiguous source, and reusable job after job"""
    L = pkg.lib()
    L.vit_stager_create.restype = C.c_void_p
    L.vit_stager_create.argtypes = [C.c_int]
    L.vit_stager_destroy.argtypes = [C.c_void_p]
    L.vit_stager_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    L.vit_stager_copy.restype = None
    rng = np.random.default_rng(5)
    for threads in (1, 4):
        st = L.vit_stager_create(threads)
        assert st
        try:
            for elems, count in ((150528, 9), (33333, 37), (7, 3), (150528, 1)):
                blocks = [rng.standard_normal(elems, dtype=np.float32) for _ in range(count)]
                structs, keep = pkg.make_image_structs(np.stack(blocks).reshape(count, 1, 1, elems))
                for i, b in enumerate(blocks):          # separate allocations, like load_image_data
                    structs[i].data = b.ctypes.data_as(C.POINTER(C.c_float))
                dst = np.zeros((count, elems), np.float32)
                L.vit_stager_copy(st, dst.ctypes.data, structs, None, elems * 4, count)
                assert np.array_equal(dst, np.stack(blocks))
                src = np.ascontiguousarray(np.stack(blocks))
                dst2 = np.zeros_like(dst)
                L.vit_stager_copy(st, dst2.ctypes.data, None, src.ctypes.data, elems * 4, count)
                assert np.array_equal(dst2, src)
        finally:
            L.vit_stager_destroy(st)


def test_model_from_blobs_reads_dims_off_sizes(pkg):
    """vitb200_model_from_blobs (host logic, no device): B/16, B/32 and S/16 blob sets, and the fall-back"""
    L = pkg.lib()
    for variant, img in (("b16", 224), ("b32", 224), ("s16", 224), ("b32", 384)):
        shapes = pkg.synth.blob_shapes(img, variant)
        blobs = [np.zeros(int(np.prod(s)), np.float32) for s in shapes]
        nets, keep = pkg.make_network_structs(blobs)
        m = pkg.Model()
        assert L.vitb200_model_from_blobs(nets, None, C.byref(m)) == 0
        patch, embed, depth, heads, hidden = pkg.synth.VARIANTS[variant]
        assert (m.img, m.patch, m.embed, m.depth, m.heads, m.hidden) == (img, patch, embed, depth, heads, hidden)
    blobs[1] = None  # conv filters absent (Network.c:144-148 leaves {NULL,0}): ViT-B/16 defaults, load_weights names the blob
    nets, keep = pkg.make_network_structs(blobs)
    m = pkg.Model()
    assert L.vitb200_model_from_blobs(nets, None, C.byref(m)) == 0
    assert (m.patch, m.embed, m.hidden) == (16, 768, 3072)


def test_shard_plan_covers_every_image_once():
    """host logic of the multi-GPU split (vit_opencl.c): contiguous, disjoint, complete"""
    import __graft_entry__ as g
    pkg = g.load_package()
    L = pkg.lib()
    L.vitb200_shard_plan.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    for n in (1, 2, 7, 100, 256, 600, 4096, 4097):
        for gpus in (1, 2, 3, 4, 8, 64):
            first, count = (C.c_int * 64)(), (C.c_int * 64)()
            used = L.vitb200_shard_plan(n, gpus, first, count)
            assert 1 <= used <= min(gpus, n)
            nxt = 0
            for k in range(used):
                assert first[k] == nxt and count[k] > 0
                nxt += count[k]
            assert nxt == n
            assert max(count[:used]) - min(count[:used]) <= max(count[:used])  # last shard may be short
            assert max(count[:used]) == -(-n // min(gpus, n))
    assert L.vitb200_shard_plan(0, 4, (C.c_int * 4)(), (C.c_int * 4)()) == 0
    assert L.vitb200_shard_plan(4096, 8, (C.c_int * 8)(), (C.c_int * 8)()) == 8


def test_chunk_schedule_of_a_call():
    """host logic of the chunk pipeline (vit_engine.c: vitb200_next_chunk): the chunks of a call cover its images once,
    none is larger than the engine's chunk; a call of at least one chunk opens with a quarter chunk (head split);
    with pageable sources (tail split) at most a quarter chunk is left after the last upload of a multi-chunk call"""
    import __graft_entry__ as g
    L = g.load_package().lib()
    L.vitb200_next_chunk.argtypes = [C.c_int] * 5

    def schedule(n, chunk, head, tail):
        done, sizes = 0, []
        while done < n:
            b = L.vitb200_next_chunk(n, done, chunk, head, tail)
            assert 0 < b <= chunk and b <= n - done
            sizes.append(b)
            done += b
        assert L.vitb200_next_chunk(n, done, chunk, head, tail) == 0
        return sizes

    for chunk in (64, 256):
        q = chunk // 4
        for n in (1, q, chunk - 1, chunk, chunk + 1, 2 * chunk, 2 * chunk + q, 4096, 4097, 512):
            plain = schedule(n, chunk, 0, 0)
            assert sum(plain) == n and all(b == chunk for b in plain[:-1])
            for head, tail in ((1, 0), (0, 1), (1, 1)):
                s = schedule(n, chunk, head, tail)
                assert sum(s) == n
                if head and n >= chunk:
                    assert s[0] == q
                if not head:
                    assert s[0] == min(n, chunk)
                if tail and len(s) > 1:
                    assert s[-1] <= max(q, chunk // 2)   # never a whole chunk after the last upload
                if tail and len(plain) > 1 and plain[-1] > chunk // 2:
                    assert s[-1] == q
    assert schedule(256, 256, 1, 1) == [64, 128, 64]
    assert schedule(512, 256, 1, 1) == [64, 256, 128, 64]
    assert schedule(512, 256, 1, 0) == [64, 256, 192]
    assert schedule(100, 256, 1, 1) == [100]
    assert L.vitb200_next_chunk(0, 0, 256, 1, 1) == 0


def test_dispatch_predicates_at_the_shapes_of_the_path():
    """host-side kernel selection (no device needed: the SM count falls back to the B200's 148): which chunks run the
    CTA-pair GEMMs with the folded LayerNorm (vitcu_gemm_bf16_emit_supported -- also the condition under which the chunk
    schedule may cut quarter chunks) and which run the K-sliced single-CTA chain (vitcu_gemm_split_k_pays: batch-1 latency)"""
    import __graft_entry__ as g
    L = g.load_package().lib()
    T, D, HID = 197, 768, 3072
    # 256, 64 and 32 images are CTA-pair chunks, 16 images and fewer are not
    assert [L.vitcu_gemm_bf16_emit_supported(b * T, D) for b in (256, 64, 32, 16, 4, 1)] == [1, 1, 1, 0, 0, 0]
    assert L.vitcu_gemm_bf16_emit_supported(256 * T, 384) == 0          # widths that are not a multiple of 256 never pair
    # K slicing pays while the 128 x 128 tiles of a product occupy at most half of the SMs
    assert [L.vitcu_gemm_split_k_pays(T, n) for n in (3 * D, HID, D)] == [1, 1, 1]            # one 224 x 224 image
    assert [L.vitcu_gemm_split_k_pays(2 * T, n) for n in (3 * D, HID)] == [1, 0]              # two images: fc1 is 96 tiles
    assert L.vitcu_gemm_split_k_pays(577, 3 * D) == 0 and L.vitcu_gemm_split_k_pays(256 * T, D) == 0


def test_ctypes_mirrors_match_the_c_headers(tmp_path):
    """the ctypes structures of the Python mirror against the C headers themselves: a C program compiled from
    include/*.h prints sizeof / offsetof of every mirrored struct, so that a field added on one side only fails here
    instead of silently shifting the arguments of the GPU tests"""
    import subprocess
    import __graft_entry__ as g
    pkg = g.load_package()
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stddef.h>
#include <stdio.h>
#include "vit_b200.h"
#include "vit_cuda_layer.h"
int main(void)
{
    printf("gemm_desc %zu %zu %zu %zu %zu %zu\n", sizeof(vitcu_gemm_desc), offsetof(vitcu_gemm_desc, lda), offsetof(vitcu_gemm_desc, ldc),
           offsetof(vitcu_gemm_desc, ln_colsum), offsetof(vitcu_gemm_desc, emit_scale), offsetof(vitcu_gemm_desc, accumulate));
    printf("call_stats %zu %zu %zu\n", sizeof(vitb200_call_stats), offsetof(vitb200_call_stats, wall_s), offsetof(vitb200_call_stats, teardown_s));
    printf("model %zu %zu\n", sizeof(vitb200_model), offsetof(vitb200_model, hidden));
    printf("image %zu %zu\n", sizeof(vitb200_image), offsetof(vitb200_image, data));
    printf("blob %zu %zu\n", sizeof(vitb200_blob), offsetof(vitb200_blob, size));
    return 0;
}
''')
    exe = tmp_path / "abi"
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    subprocess.run(["gcc", "-std=c11", "-I", inc, "-o", str(exe), str(src)], check=True)
    out = dict((l.split()[0], [int(v) for v in l.split()[1:]]) for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    G, S, M = pkg.GemmDesc, pkg.CallStats, pkg.Model
    assert out["gemm_desc"] == [C.sizeof(G), G.lda.offset, G.ldc.offset, G.ln_colsum.offset, G.emit_scale.offset, G.accumulate.offset]
    assert out["call_stats"] == [C.sizeof(S), S.wall_s.offset, S.teardown_s.offset]
    assert out["model"] == [C.sizeof(M), M.hidden.offset]
    assert out["image"] == [C.sizeof(pkg.ImageData), pkg.ImageData.data.offset]
    assert out["blob"] == [C.sizeof(pkg.Network), pkg.Network.size.offset]
