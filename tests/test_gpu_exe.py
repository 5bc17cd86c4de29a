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


@pytest.mark.skipif(not os.path.exists(os.path.join(BIN, "vf_dada_db")), reason="executables not built")
def test_shared_memory_ring_matches_file_replay(tmp_path):
    """the reference's arrangement (scripts/start_dada:13): dada_db makes the ring, a producer process fills it,
    process_baseband -k reads it; same bytes as the file replay of the same data"""
    key = "%x" % (0x5000 + os.getpid() % 0x1000)
    subprocess.run([os.path.join(BIN, "vf_dada_db"), "-k", key, "-d"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.run([os.path.join(BIN, "vf_dada_db"), "-k", key, "-n", "3"], check=True, stdout=subprocess.DEVNULL)
    try:
        d_ring, d_file = tmp_path / "ring", tmp_path / "file"
        d_ring.mkdir(); d_file.mkdir()
        gen = ["-t", "2", "-r", "21", "-f", "-n", "6"]
        reader = subprocess.Popen([os.path.join(BIN, "process_baseband"), "-k", key, "-s", "-D", str(d_ring), "-b", "8", "-r", "2", "-j"],
                                  stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        subprocess.run([os.path.join(BIN, "genbase"), "-k", key] + gen, check=True, timeout=300)
        out, err = reader.communicate(timeout=300)
        assert reader.returncode == 0, err[-2000:]
        summ = json.loads(out.strip().splitlines()[-1])
        assert summ["seconds"] == 2 and summ["segments"] == 20 and summ["exit"] == 0
        vdif = tmp_path / "two.vdif"
        subprocess.run([os.path.join(BIN, "genbase"), "-o", str(vdif)] + gen, check=True)
        r = subprocess.run([os.path.join(BIN, "process_baseband"), "-f", str(vdif), "-D", str(d_file), "-b", "8", "-r", "2", "-a", "6", "-j", "-n", "2"],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        a, b = sorted(os.listdir(d_ring)), sorted(os.listdir(d_file))
        assert len(a) == 2 and len(b) == 2
        for fa, fb in zip(a, b):
            da, db = open(d_ring / fa, "rb").read(), open(d_file / fb, "rb").read()
            _, _, ua = parse_sigproc(da); _, _, ub = parse_sigproc(db)
            assert da[ua:] == db[ub:] and len(da) - ua == 20 * 128 * 4096
    finally:
        subprocess.run([os.path.join(BIN, "vf_dada_db"), "-k", key, "-d"], stdout=subprocess.DEVNULL)


_RING_READER = r"""
import ctypes as C, sys
L = C.CDLL(sys.argv[1])
L.vf_ring_connect_shm.restype = C.c_void_p; L.vf_ring_connect_shm.argtypes = [C.c_int]
L.vf_ring_header_read.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
L.vf_ring_read.restype = C.c_ssize_t; L.vf_ring_read.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
L.vf_ring_destroy.argtypes = [C.c_void_p]
r = L.vf_ring_connect_shm(int(sys.argv[2], 16))
hdr = C.create_string_buffer(4096)
assert L.vf_ring_header_read(r, hdr, 120000) == 0
open(sys.argv[3] + ".hdr", "wb").write(hdr.value)
buf = C.create_string_buffer(1 << 20)
with open(sys.argv[3], "wb") as f:
    while True:
        n = L.vf_ring_read(r, buf, len(buf))
        f.write(buf.raw[:n])
        if n < len(buf):
            break
L.vf_ring_destroy(r)
"""


@pytest.mark.skipif(not os.path.exists(os.path.join(BIN, "vf_dada_db")), reason="executables not built")
def test_output_rings(pkg, tmp_path):
    """-C: the main stream segment by segment (co-adder ring, src/process_baseband.cu:1416-1422);
    -K: the first 10 s at once, then second by second (heimdall ring, :1482-1494).  Both carry the
    psrdada header of :136-201 and, over 12 s, exactly the bytes of the _kur filterbank file."""
    import sys
    base = 0x4000 + (os.getpid() % 0x800) * 2
    kout, kco = "%x" % base, "%x" % (base + 1)
    for k in (kout, kco):
        subprocess.run([os.path.join(BIN, "vf_dada_db"), "-k", k, "-d"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        subprocess.run([os.path.join(BIN, "vf_dada_db"), "-k", k, "-b", "1310720", "-n", "16"], check=True, stdout=subprocess.DEVNULL)
    try:
        script = tmp_path / "reader.py"
        script.write_text(_RING_READER)
        lib = pkg.hostlib()._name
        readers = [subprocess.Popen([sys.executable, str(script), lib, k, str(tmp_path / n)]) for k, n in ((kout, "out.bin"), (kco, "co.bin"))]
        r = subprocess.run([os.path.join(BIN, "process_baseband"), "-S", "2", "-L", "6", "-F", "-D", str(tmp_path), "-b", "2", "-r", "2",
                            "-K", kout, "-C", kco, "-j"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        for p in readers:
            assert p.wait(timeout=60) == 0
        fil = [f for f in os.listdir(tmp_path) if f.endswith("_kur.fil")]
        assert len(fil) == 1
        buf = open(tmp_path / fil[0], "rb").read()
        _, _, used = parse_sigproc(buf)
        data = buf[used:]
        assert len(data) == 120 * 128 * 4096 // 4
        assert open(tmp_path / "co.bin", "rb").read() == data
        assert open(tmp_path / "out.bin", "rb").read() == data
        hdr = open(tmp_path / "out.bin.hdr", "rb").read().decode()
        assert "NCHAN 4096" in hdr.replace("  ", " ") or "NCHAN" in hdr
        assert fil[0] in hdr                                          # SIGPROC_FILE key, :197
    finally:
        for k in (kout, kco):
            subprocess.run([os.path.join(BIN, "vf_dada_db"), "-k", k, "-d"], stdout=subprocess.DEVNULL)


@pytest.mark.skipif(not os.path.exists(os.path.join(BIN, "process_baseband")), reason="executables not built")
def test_frames_anywhere_in_the_second_and_statistics_dumps(pkg, tmp_path):
    """(1) ADVICE r01: the executable hands the whole one-second block to the library, so frames may sit anywhere in
    the second (the reference places them by header across the second, src/process_baseband.cu:1015-1035); a frame of
    another second is skipped with a warning, not fatal.  (2) -W / -H: the WRITE_KURTO / DOHISTO dumps
    (:1378-1393, :1444-1450) next to the filterbank equal vf_get_stats of every segment."""
    vdif = tmp_path / "one.vdif"
    subprocess.run([os.path.join(BIN, "genbase"), "-o", str(vdif), "-t", "1", "-r", "13", "-f", "-n", "4"], check=True)
    frames = np.fromfile(vdif, np.uint8).reshape(-1, 5032).copy()
    out = {}
    for tag, extra in (("plain", []), ("dumps", ["-W", "-H", "-T"])):
        d = tmp_path / tag
        d.mkdir()
        r = subprocess.run([os.path.join(BIN, "process_baseband"), "-f", str(vdif), "-D", str(d), "-b", "8", "-r", "2", "-a", "4", "-j", "-n", "2", "-o"] + extra,
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        out[tag] = {f: open(d / f, "rb").read() for f in sorted(os.listdir(d))}
        if extra:
            assert "Channelise.." in r.stderr and "Normalise..." in r.stderr
    fil = [f for f in out["plain"] if f.endswith("_kur.fil")][0]
    assert out["plain"][fil] == out["dumps"][fil]                     # block path == per-segment path
    # shuffled across the whole second, one frame re-stamped with another second
    rng = np.random.default_rng(5)
    sh = frames[rng.permutation(frames.shape[0])].copy()
    victim = int(np.flatnonzero((sh[:, 4:8].copy().view(np.uint32)[:, 0] & 0xFFFFFF) == 1234)[0])
    sec = int(sh[victim, :4].copy().view(np.uint32)[0] & 0x3FFFFFFF)
    # keep the first frame in place: the executable checks the alignment of the observation's first frame (:843-849)
    first = int(np.flatnonzero(((sh[:, 4:8].copy().view(np.uint32)[:, 0] & 0xFFFFFF) == 0) & (((sh[:, 12:16].copy().view(np.uint32)[:, 0] >> 16) & 0x3FF) == 0))[0])
    sh[[0, first]] = sh[[first, 0]]
    if victim == 0 or victim == first:
        victim = 7
    pol = ((sh[victim, 12:16].copy().view(np.uint32)[0] >> 16) & 0x3FF) != 0
    fno = int(sh[victim, 4:8].copy().view(np.uint32)[0] & 0xFFFFFF)
    sh[victim, :4] = np.frombuffer(np.uint32(sec + 1).tobytes(), np.uint8)
    shf = tmp_path / "shuffled.vdif"
    sh.tofile(shf)
    d = tmp_path / "shuffled"
    d.mkdir()
    r = subprocess.run([os.path.join(BIN, "process_baseband"), "-f", str(shf), "-D", str(d), "-b", "8", "-r", "2", "-a", "4", "-j", "-n", "2", "-o"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "of another second skipped" in r.stderr
    # the same stream with that frame's samples dropped, through the binding
    ordered = frames.copy()
    idx = 2 * fno + int(pol)
    ordered[idx, 3] |= 0x80                                           # invalid bit: an absent frame
    with pkg.Pipeline(ffts_per_seg=1024, nbit=8, npol=1, rfi_mode=2) as p:
        main, raw, rep, warned = p.process_vdif_block(np.ascontiguousarray(ordered).reshape(-1), 0, sec, 10)
    got = open(d / [f for f in os.listdir(d) if f.endswith("_kur.fil")][0], "rb").read()
    _, _, used = parse_sigproc(got)
    assert np.array_equal(np.frombuffer(got[used:], np.uint8), main.reshape(-1))
    # dumps against the binding, segment by segment
    base = fil[:-len("_kur.fil")]
    kur = np.frombuffer(out["dumps"][base + ".kurto"], np.float32).reshape(10, 2 * 1024 * 25)
    kfb = np.frombuffer(out["dumps"][base + ".block_kurto"], np.float32).reshape(10, 2 * 1024)
    wts = np.frombuffer(out["dumps"][base + ".weights"], np.float32).reshape(10, 1024)
    his = np.frombuffer(out["dumps"][base + ".histo"], np.uint32).reshape(10, 512)
    with pkg.Pipeline(ffts_per_seg=1024, nbit=8, npol=1, rfi_mode=2, keep_stats=1, do_histo=1) as p:
        for s in range(10):
            p.process_vdif(np.ascontiguousarray(frames[s * 5120:(s + 1) * 5120]).reshape(-1), s * 2560)
            st = p.get_stats()
            assert np.array_equal(st["kur"], kur[s], equal_nan=True) and np.array_equal(st["kur_fb"], kfb[s], equal_nan=True), s
            assert np.array_equal(st["weights"][:1024], wts[s]) and np.array_equal(st["histo"], his[s]), s
