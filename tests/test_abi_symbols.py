"""The C-ABI libraries load without a GPU and export every symbol their headers declare."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared(header):
    src = open(os.path.join(ROOT, header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:vf|orc)_\w+)\s*\(", src)))


def test_libvlitefast_exports_header(pkg):
    L = pkg.lib()
    names = declared("include/vlitefast.h")
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "libvlitefast.so lacks %s" % n
    # every entry point has its prototype declared to ctypes: an undeclared one would pass pointers as 32-bit ints
    unbound = [n for n in names if getattr(L, n).argtypes is None]
    assert not unbound, "binding.py declares no argtypes for %s" % unbound


def test_hostlib_exports_headers(pkg):
    L = pkg.hostlib()
    hdrs = [f for f in os.listdir(os.path.join(ROOT, "vlite-fast_b200", "host")) if f.endswith(".h")]
    assert hdrs
    for hd in hdrs:
        for n in declared(os.path.join("vlite-fast_b200", "host", hd)):
            assert hasattr(L, n), "libvlitehost.so lacks %s (%s)" % (n, hd)


def test_oracle_exports_header(orc):
    L = orc.liba()
    for n in declared("oracle/vlite_oracle.h"):
        assert hasattr(L, n), n


def test_config_defaults_are_the_reference_defaults(pkg):
    cfg = pkg.VfConfig()
    assert pkg.lib().vf_config_default(ctypes.byref(cfg)) == 0
    # src/process_baseband.h:16-55, src/process_baseband.cu:34,349-351
    assert (cfg.nfft, cfg.nscrunch, cfg.ffts_per_seg, cfg.nkurto) == (12500, 8, 1024, 500)
    assert (cfg.chanmin, cfg.chanmax, cfg.nbit, cfg.npol, cfg.rfi_mode) == (2155, 6250, 2, 1, 2)
    assert (cfg.dag_thresh, cfg.min_weight) == (3.0, 0.2)      # DAG_THRESH, MIN_WEIGHT, src/process_baseband.h:42,45
    assert ctypes.sizeof(cfg) == 96


def test_testing_hooks_are_not_in_the_product_library(pkg):
    """the monolithic channeliser and vf_debug_division exist in libvlitefast_testing.so only"""
    L, LT = pkg.lib(), pkg.lib(testing=True)
    assert not hasattr(L, "vf_debug_division") and hasattr(LT, "vf_debug_division")
    for n in declared("include/vlitefast.h"):
        assert hasattr(LT, n), n
    cfg = pkg.VfConfig()
    L.vf_config_default(ctypes.byref(cfg))
    cfg.k1_threads = 640
    h = ctypes.c_void_p()
    assert L.vf_create(ctypes.byref(cfg), ctypes.byref(h)) == 1 and not h


def test_strerror_and_bad_config(pkg):
    L = pkg.lib()
    assert L.vf_strerror(0) == b"ok"
    assert b"CUDA" in L.vf_strerror(20)
    cfg = pkg.VfConfig()
    L.vf_config_default(ctypes.byref(cfg))
    h = ctypes.c_void_p()
    cfg.nbit = 3
    assert L.vf_create(ctypes.byref(cfg), ctypes.byref(h)) == 1 and not h
    cfg.nbit = 2
    cfg.ffts_per_seg = 12
    assert L.vf_create(ctypes.byref(cfg), ctypes.byref(h)) == 1 and not h


def test_no_cpu_fallback_without_device(pkg):
    """Without a CUDA device vf_create must fail (VF_ERR_NODEV), never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.VfError) as e:
        pkg.Pipeline(ffts_per_seg=8)
    assert e.value.code == 25
