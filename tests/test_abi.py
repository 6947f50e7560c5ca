"""The C-ABI library loads on a CPU-only box and exports every symbol include/b200mel.h declares.
No compute calls (there is no GPU here); argument validation that happens before any CUDA call is checked."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200mel.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200mel_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_expected_surface():
    names = declared_functions()
    for must in ("b200mel_version", "b200mel_last_error", "b200mel_plan_create", "b200mel_plan_destroy",
                 "b200mel_out_frames", "b200mel_forward", "b200mel_forward_host", "b200mel_mel_filterbank",
                 "b200mel_hann_window", "b200mel_plan_set_filterbank", "b200mel_launch_count"):
        assert must in names


def test_library_exports_every_declared_symbol(built_lib):
    handle = C.CDLL(built_lib.LIB_PATH)
    for name in declared_functions():
        assert hasattr(handle, name), f"{name} declared in include/b200mel.h but not exported"
    assert set(declared_functions()) == set(built_lib.SYMBOLS), "ctypes binding and header disagree"
    assert handle.b200mel_version() == int(re.search(r"#define B200MEL_VERSION (\d+)", open(HEADER).read()).group(1))


def test_struct_layouts_match_header(built_lib):
    # 13 x 4-byte fields and 8 x 4-byte fields, no padding (the library rejects a struct_size mismatch)
    assert C.sizeof(built_lib.Config) == 13 * 4
    assert C.sizeof(built_lib.Epilogue) == 8 * 4
    hdr = open(HEADER).read()
    cfg_fields = re.search(r"typedef struct b200mel_config \{(.*?)\} b200mel_config;", hdr, re.S).group(1)
    cfg_fields = re.sub(r"/\*.*?\*/", "", cfg_fields, flags=re.S)
    names = re.findall(r"\b(?:int32_t|float)\s+(\w+);", cfg_fields)
    assert names == [f[0] for f in built_lib.Config._fields_]
    epi_fields = re.search(r"typedef struct b200mel_epilogue \{(.*?)\} b200mel_epilogue;", hdr, re.S).group(1)
    epi_fields = re.sub(r"/\*.*?\*/", "", epi_fields, flags=re.S)
    assert re.findall(r"\b(?:int32_t|float)\s+(\w+);", epi_fields) == [f[0] for f in built_lib.Epilogue._fields_]


def test_io_struct_matches_header(built_lib):
    hdr = open(HEADER).read()
    body = re.search(r"typedef struct b200mel_io \{(.*?)\} b200mel_io;", hdr, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        decl = re.sub(r"^(const\s+)?(float|int32_t|int64_t)\s*", "", decl)
        names += [n.strip().lstrip("*").strip() for n in decl.split(",")]
    assert names == [f[0] for f in built_lib.IO._fields_]
    # pointers / int64 are 8-byte aligned: 2 x int32, then 9 x 8 bytes, 2 x 4 bytes, 2 pointers, 2 x int32
    assert C.sizeof(built_lib.IO) == 8 + 9 * 8 + 8 + 2 * 8 + 8
    lib = built_lib.lib()
    io = built_lib.IO()
    io.struct_size = 4
    assert lib.b200mel_forward_io(None, C.byref(io), None, None) == built_lib.EINVAL


def test_error_codes_without_gpu(built_lib):
    import torch

    lib = built_lib.lib()
    h = C.c_void_p()
    assert lib.b200mel_plan_create(None, C.byref(h)) == built_lib.EINVAL
    assert b"null" in lib.b200mel_last_error()
    bad = built_lib.make_config(22050, 1000, 1000, 256, 80)
    assert lib.b200mel_plan_create(C.byref(bad), C.byref(h)) == built_lib.EUNSUP  # n_fft not 1024/2048
    bad = built_lib.make_config(22050, 1024, 2048, 256, 80)
    assert lib.b200mel_plan_create(C.byref(bad), C.byref(h)) == built_lib.EINVAL  # win > n_fft
    bad = built_lib.make_config(22050, 1024, 1024, 0, 80)
    assert lib.b200mel_plan_create(C.byref(bad), C.byref(h)) == built_lib.EINVAL
    bad = built_lib.make_config(22050, 1024, 1024, 256, 80)
    bad.struct_size = 4
    assert lib.b200mel_plan_create(C.byref(bad), C.byref(h)) == built_lib.EINVAL
    assert b"struct_size" in lib.b200mel_last_error()
    if not torch.cuda.is_available():
        ok = built_lib.make_config(22050, 1024, 1024, 256, 80, 0.0, 8000.0)
        assert lib.b200mel_plan_create(C.byref(ok), C.byref(h)) == built_lib.ENODEV  # loud: no CPU fallback
        assert b"no CPU fallback" in lib.b200mel_last_error()
        with pytest.raises(built_lib.B200MelError):
            built_lib.Plan(ok, 0)
    assert lib.b200mel_forward(None, None, 1, 1, 1, None, None, None, 0, None, None, None) == built_lib.EINVAL
    assert lib.b200mel_plan_destroy(None) == built_lib.OK
    out = np.zeros(4, dtype=np.float32)
    assert lib.b200mel_hann_window(8, 4, out.ctypes.data) == built_lib.EINVAL
    assert lib.b200mel_mel_filterbank(22050, 1024, 80, 9000.0, 8000.0, 0, 1, out.ctypes.data) == built_lib.EINVAL
    with pytest.raises(ValueError):
        built_lib.check(built_lib.EINVAL)


def test_wave_operator_argument_validation_without_gpu(built_lib):
    """Argument checks of the waveform-side entry points happen before any CUDA call."""
    import torch

    lib = built_lib.lib()
    buf = np.zeros(64, dtype=np.float32)
    p = buf.ctypes.data
    assert lib.b200mel_preemphasis(None, 1, 16, 16, 0.97, p, 16, None) == built_lib.EINVAL
    assert lib.b200mel_preemphasis(p, 1, 1, 1, 0.97, p, 1, None) == built_lib.EINVAL  # reflect pad needs L >= 2
    assert b"L >= 2" in lib.b200mel_last_error()
    assert lib.b200mel_preemphasis(p, 2, 16, 8, 0.97, p, 16, None) == built_lib.EINVAL  # row stride < L
    assert lib.b200mel_preemphasis(p, 0, 16, 16, 0.97, p, 16, None) == built_lib.OK  # empty batch: nothing to do
    assert lib.b200mel_volume_norm(p, 8, -11.5, p, None, None) == built_lib.EINVAL
    assert lib.b200mel_volume_norm(p, 0, -11.5, p, None, None) == built_lib.OK
    assert lib.b200mel_mel_to_mfcc(p, p, 1, 0, 4, 4, p, None) == built_lib.EINVAL
    assert lib.b200mel_mel_to_mfcc(p, p, 1, 200, 4, 4, p, None) == built_lib.EUNSUP  # more than 128 mel rows
    assert lib.b200mel_mel_to_mfcc(None, p, 1, 8, 4, 4, p, None) == built_lib.EINVAL
    if not torch.cuda.is_available():  # valid arguments, no device: loud, no fallback
        assert lib.b200mel_preemphasis(p, 1, 16, 16, 0.97, p, 16, None) == built_lib.ENODEV
        assert b"no CPU fallback" in lib.b200mel_last_error()
        assert lib.b200mel_mel_to_mfcc(p, p, 1, 8, 4, 2, p, None) == built_lib.ENODEV


def test_missing_library_fails_loudly(monkeypatch, built_lib):
    monkeypatch.setattr(built_lib, "_lib", None)
    monkeypatch.setattr(built_lib, "LIB_PATH", "/nonexistent/libb200mel.so")
    with pytest.raises(ImportError, match="no CPU"):
        built_lib.lib()


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under pytorch_sound_b200/ may import it."""
    pkg = os.path.join(ROOT, "pytorch_sound_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), os.path.join(dp, f)
                assert "mel_oracle" not in src, os.path.join(dp, f)
