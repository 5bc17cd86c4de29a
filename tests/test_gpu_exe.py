"""-m gpu: the C process_baseband executable end to end: VDIF file -> ring shim
-> GPU -> SIGPROC .fil, against the same segments pushed through the ctypes
binding, and the synthetic streaming mode."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from test_host_c import parse_sigproc

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "vlite-fast_b200", "bin")


@pytest.mark.skipif(not os.path.exists(os.path.join(BIN, "process_baseband")), reason="executables not built")
def test_file_replay_matches_binding(pkg, tmp_path):
    vdif = tmp_path / "one.vdif"
    subprocess.run([os.path.join(BIN, "genbase"), "-o", str(vdif), "-t", "1", "-r", "11", "-f", "-n", "4"], check=True)
    assert os.path.getsize(vdif) == 257638400
    r = subprocess.run([os.path.join(BIN, "process_baseband"), "-f", str(vdif), "-D", str(tmp_path), "-b", "8", "-r", "2",
                        "-a", "4", "-j", "-n", "2"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    summ = json.loads(r.stdout.strip().splitlines()[-1])
    assert summ["seconds"] == 1 and summ["segments"] == 10 and summ["exit"] == 0
    fils = sorted(f for f in os.listdir(tmp_path) if f.endswith(".fil"))
    assert len(fils) == 2 and fils[1].endswith("_muos_ea04_kur.fil") and fils[0].endswith("_muos_ea04.fil")
    frames = np.fromfile(vdif, np.uint8).reshape(-1, 5032)
    with pkg.Pipeline(ffts_per_seg=1024, nbit=8, npol=1, rfi_mode=2) as p:
        want_main, want_raw = [], []
        for s in range(10):
            m, rw = p.process_vdif(np.ascontiguousarray(frames[s * 5120:(s + 1) * 5120]).reshape(-1), s * 2560)
            want_main.append(m); want_raw.append(rw)
    for name, want in ((fils[1], want_main), (fils[0], want_raw)):
        buf = open(tmp_path / name, "rb").read()
        d, order, used = parse_sigproc(buf)
        assert d["nchans"] == 4096 and d["nbits"] == 8 and d["telescope_id"] == 4 and d["source_name"] == "REPLAY"
        data = np.frombuffer(buf[used:], np.uint8)
        assert data.size == 10 * 128 * 4096
        assert np.array_equal(data, np.concatenate(want)), name


@pytest.mark.skipif(not os.path.exists(os.path.join(BIN, "process_baseband")), reason="executables not built")
def test_synthetic_stream(tmp_path):
    r = subprocess.run([os.path.join(BIN, "process_baseband"), "-S", "2", "-L", "3", "-F", "-w", "0", "-j"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    summ = json.loads(r.stdout.strip().splitlines()[-1])
    assert summ["seconds"] == 6 and summ["segments"] == 60 and summ["x_realtime"] > 1
