"""CPU: libtango_b200.so loads without a GPU and exports every symbol include/tango_b200.h declares."""
import ctypes
import os
import re

import torch

from tango_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "tango_b200.h")).read()
    return sorted(set(re.findall(r"\b(tng_[a-z0-9_]+)\s*\(", src)))


def test_exports_match_header():
    lib = L.load()
    names = header_symbols()
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), f"{n} not exported"
    assert sorted(L.SYMBOLS) == names
    assert lib.tng_version() >= 100


def test_struct_sizes_match_c_layout():
    # tng_aview: ptr + 7 x int64; tng_kgroup: 6 x int32
    assert ctypes.sizeof(L.AView) == 64 and ctypes.sizeof(L.KGroup) == 24
    assert ctypes.sizeof(L.GemmDesc) % 8 == 0


def test_compute_call_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        return
    lib = L.load()
    d = L.GemmDesc()
    d.n_aviews, d.n_groups, d.W, d.H, d.NB, d.Ncols, d.Ktot = 1, 1, 128, 1, 1, 64, 64
    d.g[0] = L.KGroup(0, 0, 0, 0, 0, 1)
    d.a[0] = L.AView(0, 64, 128, 1, 1, 64, 8192, 8192)
    buf = ctypes.create_string_buffer(64)
    d.out_f32 = ctypes.addressof(buf)
    d.ld_f32 = 64
    rc = lib.tng_conv_gemm(ctypes.byref(d), None)
    assert rc != 0 and len(lib.tng_last_error()) > 0
