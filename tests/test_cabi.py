"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU
and exports every symbol include/ofdis_b200.h declares; argument validation that
does not need a device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from of_dis_b200 import build

    return ctypes.CDLL(build.build())


def test_every_declared_symbol_is_exported(built_lib):
    hdr = open(os.path.join(ROOT, "include", "ofdis_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(ofdis_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(built_lib, name), name
    from of_dis_b200 import api

    assert sorted(api.EXPORTS) == declared


def test_create_rejects_bad_arguments_without_touching_the_gpu(built_lib):
    from of_dis_b200 import params

    prm = params.operating_point(2, 1024)
    h = ctypes.c_void_p()
    cp = prm.to_c()
    # width not divisible by 2^sc_f (oflow.h:87)
    assert built_lib.ofdis_create(ctypes.byref(h), 0, None, ctypes.byref(cp), 2, 1000, 448, 8, 1) == -1
    # a refinement level taller than 1024 rows (one SOR thread per row): valid in the reference, not built
    assert built_lib.ofdis_create(ctypes.byref(h), 0, None, ctypes.byref(cp), 2, 1024, 8224, 8, 1) == -3
    cp.noc = 2
    assert built_lib.ofdis_create(ctypes.byref(h), 0, None, ctypes.byref(cp), 2, 1024, 448, 8, 1) == -1
    assert built_lib.ofdis_destroy(None) == 0


def test_missing_library_fails_loudly(monkeypatch):
    from of_dis_b200 import api

    monkeypatch.setattr(api, "_lib", None)
    monkeypatch.setattr(api, "LIB_PATH", "/nonexistent/libofdis_b200.so")
    with pytest.raises(api.OfdisError):
        api.lib()
