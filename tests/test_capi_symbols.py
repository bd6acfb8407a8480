"""CPU: the C-ABI library builds, loads without a GPU and exports every symbol include/hh_b200.h declares; the ctypes
binding covers exactly that set; compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "hh_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hh_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from helping_hand_for_egocentric_videos_b200 import _lib
    return _lib


def test_header_symbols_are_exported_and_bound(lib):
    names = _declared()
    assert len(names) >= 30
    dll = lib.load()
    for n in names:
        assert hasattr(dll, n), "libhh_b200.so does not export %s" % n
    assert sorted(lib.SIGNATURES) == names, set(lib.SIGNATURES) ^ set(names)


def test_library_has_no_driver_link_dependency(lib):
    import subprocess
    out = subprocess.run(["ldd", lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out and "libtorch" not in out and "libnccl" not in out


def test_bad_arguments_are_reported_not_crashed(lib):
    dll = lib.load()
    h = C.c_void_p()
    cfg = lib.EncoderCfg(224, 14, 16, 1000, 24, 16, 4000)        # embed_dim not a multiple of 128
    assert dll.hh_encoder_create(C.byref(h), C.byref(cfg)) == -2
    assert "embed_dim" in lib.last_error()
    dcfg = lib.DecoderCfg(512, 8, 6, 2048, 1, 22048, 1024, 4, 256, 1)   # num_queries = 1 unsupported... (0 < Q <= 16 ok)
    dcfg.num_queries = 40
    assert dll.hh_decoder_create(C.byref(h), C.byref(dcfg)) == -2
    assert dll.hh_encoder_set_weight(None, b"x", None, 0, None) == -1


def test_engine_rejects_unknown_or_misshapen_weights(lib):
    if torch.cuda.is_available():
        pytest.skip("CPU-only check of the host logic")
    dll = lib.load()
    h = C.c_void_p()
    cfg = lib.EncoderCfg(56, 14, 2, 128, 1, 2, 512)
    assert dll.hh_encoder_create(C.byref(h), C.byref(cfg)) == 0
    buf = (C.c_float * 128)()
    assert dll.hh_encoder_set_weight(h, b"not.a.key", buf, 128, None) == -2
    assert "unknown parameter" in lib.last_error()
    assert dll.hh_encoder_set_weight(h, b"norm.weight", buf, 64, None) == -2
    assert "expected 128" in lib.last_error()
    assert dll.hh_encoder_flops_per_clip(h) > 0
    dll.hh_encoder_destroy(h)


def test_flops_formula_matches_survey(lib):
    """3415.41 GF encoder + 33.15 GF decoder per clip at L/14, T=16, Q=13, no traj."""
    dll = lib.load()
    h = C.c_void_p()
    cfg = lib.EncoderCfg(224, 14, 16, 1024, 24, 16, 4096)
    assert dll.hh_encoder_create(C.byref(h), C.byref(cfg)) == 0
    assert abs(dll.hh_encoder_flops_per_clip(h) / 1e9 - 3415.41) < 0.5
    dll.hh_encoder_destroy(h)
    d = C.c_void_p()
    dcfg = lib.DecoderCfg(512, 8, 6, 2048, 13, 22048, 1024, 16, 256, 0)
    assert dll.hh_decoder_create(C.byref(d), C.byref(dcfg)) == 0
    assert abs(dll.hh_decoder_flops_per_clip(d, 16) / 1e9 - 33.15) < 0.3
    dll.hh_decoder_destroy(d)
    cfg4 = lib.EncoderCfg(224, 14, 4, 1024, 24, 16, 4096)
    assert dll.hh_encoder_create(C.byref(h), C.byref(cfg4)) == 0
    assert abs(dll.hh_encoder_flops_per_clip(h) / 1e9 - 853.25) < 0.3
    dll.hh_encoder_destroy(h)


def test_no_cpu_fallback():
    from helping_hand_for_egocentric_videos_b200 import ops
    from helping_hand_for_egocentric_videos_b200.model import LaviLa
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.sim_matrix(torch.randn(2, 8), torch.randn(3, 8))
    vis = LaviLa.SpaceTimeTransformer(img_size=56, patch_size=14, embed_dim=128, depth=1, num_heads=2, num_frames=2,
                                      ln_pre=True, act_layer=LaviLa.QuickGELU, time_init='zeros', num_classes=0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vis.forward_features(torch.randn(1, 2, 3, 56, 56))
