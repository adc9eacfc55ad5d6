"""CPU-side checks of the drop-in boundary: libbisbm.so loads, exports every symbol that
include/bisbm.h declares, and has NO CPU execution path (compute fails loudly without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib(pkg):
    pkg.build.build_lib()
    return pkg.host.load_library()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "bisbm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bisbm_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported(lib, host):
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), "libbisbm.so does not export " + s
        assert s in host.SIGNATURES, "host.py does not bind " + s
    for s in host.SIGNATURES:
        assert s in syms, "host.py binds undeclared symbol " + s


def test_version_and_error_string(lib):
    assert b"sm_100a" in lib.bisbm_version()
    assert isinstance(lib.bisbm_last_error(), bytes)


def test_argument_errors_do_not_need_a_gpu(lib, host):
    h = C.c_void_p()
    rc = lib.bisbm_create(0, 0, 0, None, None, 0, C.byref(h))
    assert rc == 1 and b"node count" in lib.bisbm_last_error()
    rc = lib.bisbm_set_chains(None, 1, None, None, None, 1.0)
    assert rc == 1
    assert lib.bisbm_destroy(None) == 0


def test_no_cpu_fallback(lib, host):
    """Without a CUDA device every compute entry point must fail with BISBM_ERR_CUDA."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device path cannot be exercised")
    edges = np.array([[0, 2], [1, 3]], dtype=np.uint32)
    with pytest.raises(host.BisbmError) as ei:
        host.Graph(edges, 2, 2)
    assert ei.value.code == 2 and "no CUDA device" in str(ei.value)
    with pytest.raises(host.BisbmError) as ei:      # the edge-list text entry has no host-side parser to fall back on either
        host.Graph("0 2\n1 3\n", 2, 2)
    assert ei.value.code == 2 and "no CUDA device" in str(ei.value)


def test_product_never_imports_the_oracle():
    """Only tests/, bench.py and __graft_entry__.smoke() may touch oracle/."""
    pk = os.path.join(ROOT, "bipartitesbm-mcmc_b200")
    for dp, _, fs in os.walk(pk):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "liboracle" not in txt and "libref" not in txt and "bisbm_oracle" not in txt, f
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
