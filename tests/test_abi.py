"""CPU: the C-ABI library loads and exports every symbol include/dualvar_b200.h declares; the
product path refuses to run without a GPU / without the extension (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

from dualvar_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header="dualvar_b200.h"):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dv_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.exported_symbols()) == names


def test_diagnostics_live_in_their_own_library():
    """The product library exports no dv_debug_* symbol; the diagnostics build (make diag) exports the product ABI plus
    everything include/dualvar_b200_diag.h declares."""
    lib = _lib.load()
    dbg = [n for n in _declared("dualvar_b200_diag.h") if n.startswith("dv_debug_")]
    assert len(dbg) >= 3
    for n in dbg:
        assert not hasattr(lib, n), f"{n} must not be in the product library"
    path = os.path.join(ROOT, "dualvar_b200", "lib", "libdualvar_b200_diag.so")
    if not os.path.exists(path):
        pytest.skip("diagnostics library not built (make diag)")
    diag = ctypes.CDLL(path)
    for n in dbg + _declared():
        assert hasattr(diag, n), n


def test_comm_status_word_starts_clean_and_timeout_is_settable():
    """dv_comm_status / dv_comm_set_timeout (peer-exchange failure reporting) work without a GPU: no exchange has run, so
    the status is 0 and asking for it allocates nothing."""
    lib = _lib.load()
    peer, seq = ctypes.c_int(7), ctypes.c_int64(7)
    assert lib.dv_comm_set_timeout(ctypes.c_double(12.5)) == 0
    assert lib.dv_comm_status(ctypes.byref(peer), ctypes.byref(seq), 1) == 0
    assert lib.dv_comm_set_timeout(ctypes.c_double(600.0)) == 0


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.dv_version() >= 1
    g = _lib.make_geom(1, 1, 1, 1, 8, 8, (1, 1, 1), (1, 1, 1), (0, 0, 0))
    rc = lib.dv_pack_conv_weight(None, None, None, ctypes.byref(g), None)
    assert rc != 0 and b"NULL" in lib.dv_last_error()
    bad = _lib.make_geom(1, 4, 4, 4, 8, 8, (3, 3, 3), (3, 1, 1), (1, 1, 1))
    rc = lib.dv_conv3d_fprop_bf16(ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), None, None,
                                  ctypes.byref(bad), None)
    assert rc != 0 and b"stride" in lib.dv_last_error()


def test_geometry_helper_matches_conv_formula():
    g = _lib.make_geom(2, 16, 112, 112, 3, 83, (1, 7, 7), (1, 2, 2), (0, 3, 3))
    assert (g.To, g.Ho, g.Wo) == (16, 56, 56) and g.Cin_p == 8 and g.Cout_p == 88 and g.taps == 49


def test_missing_extension_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "_LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.DualVarNativeError):
        _lib.load()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU error path")
def test_cpu_tensors_are_rejected_not_silently_computed():
    from types import SimpleNamespace
    from dualvar_b200 import models as PM
    m = PM.SimCLR_TimeSeriesV4("r3d", 128, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc",
                               SimpleNamespace(shufflerank_theta=0.05))
    with pytest.raises(_lib.DualVarNativeError):
        m(torch.zeros(2, 3, 3, 8, 32, 32))
    with pytest.raises(_lib.DualVarNativeError):
        m.encoder_q[0](torch.zeros(1, 3, 8, 32, 32))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "dualvar_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn


def test_new_entry_points_validate_arguments_without_a_gpu():
    """Status codes + dv_last_error() of the fp32-mode and frame-staging entry points: every bad call is refused on the
    host side, before any CUDA call (so this runs on the CPU-only build box)."""
    lib = _lib.load()
    P = ctypes.c_void_p(64)      # a non-NULL, never dereferenced pointer

    def err():
        return lib.dv_last_error().decode()

    # split planes: 1..3 planes, planes must not overlap and must keep 16-byte alignment
    assert lib.dv_f32_split(P, P, 1024, 4, 1024, None) != 0 and "n_planes" in err()
    assert lib.dv_f32_split(P, P, 100, 2, 1024, None) != 0 and "plane_stride" in err()
    assert lib.dv_f32_split(P, P, 1028, 2, 1024, None) != 0 and "plane_stride" in err()
    assert lib.dv_f32_split_planes(None, P, 8, 2, None) != 0
    # BatchNorm apply into a concat slice: slice must fit the row and stay 8-channel aligned
    assert lib.dv_f32_bn_apply(P, P, None, None, None, P, None, 0, 3, 10, 64, 64, 8, 1, None) != 0 and "f32_bn_apply" in err()
    assert lib.dv_f32_bn_apply(P, P, P, None, None, P, None, 0, 3, 10, 64, 64, 0, 1, None) != 0 and "go together" in err()
    assert lib.dv_f32_bn_bwd_reduce(P, None, None, P, None, P, 10, 64, 64, 0, 1, None) != 0       # relu without a mask source
    assert lib.dv_f32_bn_bwd_apply(P, None, None, P, P, P, None, 640, 3, None, 10, 64, 64, 0, 1, None) != 0   # no dy planes
    # convolutions: geometry is checked exactly like the bf16 entry points
    bad = _lib.make_geom(1, 4, 4, 4, 8, 8, (3, 3, 3), (3, 1, 1), (1, 1, 1))
    assert lib.dv_conv3d_fprop_f32acc(P, P, P, None, ctypes.byref(bad), 0, None) != 0 and "stride" in err()
    assert lib.dv_conv3d_dgrad_f32acc(P, P, P, ctypes.byref(bad), 1, None) != 0 and "stride" in err()
    good = _lib.make_geom(1, 4, 8, 8, 8, 8, (3, 3, 3), (1, 1, 1), (1, 1, 1))
    assert lib.dv_conv3d_fprop_f32acc(None, P, P, None, ctypes.byref(good), 0, None) != 0 and "NULL" in err()
    assert lib.dv_conv3d_stem_fprop_f32acc(P, P, P, None, ctypes.byref(good), 0, None) != 0 and "stem path" in err()
    # frame staging: the crop must fit the scaled frame
    assert lib.dv_frames_scale_crop_u8(P, P, P, P, 1, 3, 16, 240, 320, 128, 171, 144, 112, None) != 0 and "does not fit" in err()
    assert lib.dv_frames_scale_crop_u8(P, P, P, None, 1, 3, 16, 240, 320, 128, 171, 112, 112, None) != 0 and "NULL" in err()
    assert lib.dv_frames_scale_crop_u8(P, P, P, P, 0, 3, 16, 240, 320, 128, 171, 112, 112, None) != 0 and "empty" in err()
    # the host-only table accessor reports a short buffer instead of overrunning it
    buf = (ctypes.c_int32 * 16)()
    ks = ctypes.c_int32(0)
    assert lib.dv_frames_axis_table_host(320, 128, buf, 16, ctypes.byref(ks)) != 0 and "buffer holds" in err()
    assert ks.value == 11
    # ingest into planes: plane stride must cover one plane
    assert lib.dv_ingest_clips_planes(P, 0, P, 8, 2, None, 1, 1, 1, 1, 1, 3, 4, 8, 8, 0, 1, 0, None, None, 0, None) != 0 \
        and "plane_stride" in err()


def test_sim_ce_block_count_matches_the_host_mirror():
    """objectives._sim_blocks sizes the partials buffer of the fused similarity / log-sum-exp launch: it must equal the
    library's own column-block count (host-only entry point: no GPU needed)."""
    from dualvar_b200 import _lib
    from dualvar_b200.objectives import _sim_blocks
    lib = _lib.load()
    for C in (1, 2, 127, 128, 129, 1024, 16384, 16385):
        assert lib.dv_sim_ce_blocks(C) == _sim_blocks(C)
