"""The C-ABI library loads and exports every symbol include/rtk.h declares (no compute calls)."""
import ctypes
import os
import re

from common import PRODUCT_LIB, ROOT, ensure_built


def test_library_exports_every_declared_symbol():
    ensure_built()
    hdr = open(os.path.join(ROOT, "include", "rtk.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(rtk_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 20
    lib = ctypes.CDLL(PRODUCT_LIB)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product must fail loudly, not compute on the CPU."""
    import ratatosk_b200 as rb
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    try:
        rb.Context(0)
    except rb.RtkError as e:
        assert "no CUDA device" in str(e)
    else:
        raise AssertionError("Context() succeeded without a GPU")


def test_product_does_not_link_test_infrastructure():
    out = os.popen("ldd %s" % PRODUCT_LIB).read()
    assert "oracle" not in out and "hostsim" not in out and "ref_seams" not in out
