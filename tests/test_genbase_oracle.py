"""CPU: the numpy restatement of the GPU generator (oracle/gen_oracle.py): Philox known answers, the sweep lengths
of the reference's default DM, statistics of the noise, and that libvlitegen exports its header."""
import ctypes
import importlib.util
import os

import numpy as np

from conftest import ROOT
from test_abi_symbols import declared


def load_gen_oracle():
    spec = importlib.util.spec_from_file_location("gen_oracle", os.path.join(ROOT, "oracle", "gen_oracle.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_philox_known_answers():
    """Random123 kat_vectors, philox4x32 10 rounds"""
    g = load_gen_oracle()
    out = g.philox4x32_10(0, 0, 0, 0, 0, 0)
    assert [int(x) for x in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    out = g.philox4x32_10(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff)
    assert [int(x) for x in out] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    out = g.philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)
    assert [int(x) for x in out] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_sweep_of_the_default_dm():
    """DM 30 between 320 and 384 MHz: 0.21 s + 0.16 s, the 'will lose 0.37 to edge effects' of src/genbase.cu:203-206"""
    g = load_gen_oracle()
    n_lo, n_hi = g.sweep_samples(30.0)
    assert n_lo % 2 == 0 and n_hi % 2 == 0 and n_hi > n_lo
    assert abs((n_lo + n_hi) / g.RATE - 0.3715) < 2e-3


def test_noise_is_standard_normal_and_tiles():
    g = load_gen_oracle()
    x = g.noise(0, 400000, 0, 42, 10 ** 12, 1, 1.0)
    assert abs(x.mean()) < 5e-3 and abs(x.std() - 1) < 5e-3
    k = ((x / x.std()) ** 4).mean()
    assert abs(k - 3) < 0.05
    # a block that starts off a multiple of 4 sees the same stream
    y = g.noise(1003, 5000, 0, 42, 10 ** 12, 1, 1.0)
    assert np.array_equal(y, x[1003:6003])
    # the pulse: 3 % of each period scaled by the amplitude
    p = g.noise(0, 40000, 1, 7, 10000, 1, 2.0)
    q = g.noise(0, 40000, 1, 7, 10000, 1, 1.0)
    on = (np.arange(40000) % 10000) < 300
    assert np.array_equal(p[~on], q[~on]) and np.array_equal(p[on], np.float32(2.0) * q[on])
    # every second period only
    r = g.noise(0, 40000, 1, 7, 10000, 2, 2.0)
    odd = on & ((np.arange(40000) // 10000) % 2 == 1)
    assert np.array_equal(r[odd], q[odd]) and np.array_equal(r[on & ~odd], p[on & ~odd])


def test_libvlitegen_exports_header(pkg):
    L = pkg.genlib()
    src = open(os.path.join(ROOT, "include", "vlitegen.h")).read()
    import re
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = sorted(set(re.findall(r"\b(vfg_\w+)\s*\(", src)))
    assert len(names) >= 9
    for n in names:
        assert hasattr(L, n), n
    cfg = pkg.VfgConfig()
    assert L.vfg_config_default(ctypes.byref(cfg)) == 0
    assert cfg.dm == 30 and cfg.pulse_period == 0.5 and cfg.buflen == 32000000 and cfg.seed == 42   # src/genbase.cu:81-88,203
