"""Host-side C helpers (libvlitehost.so): VDIF header access, SIGPROC / psrdada
headers, file naming, ring-buffer shim, generator.  No GPU needed."""
import ctypes as C
import struct
import threading

import numpy as np
import pytest


class VdifHeader(C.Structure):
    _fields_ = [("w", C.c_uint32 * 8)]


class ObsInfo(C.Structure):
    _fields_ = [("name", C.c_char * 128), ("station_id", C.c_int), ("ra", C.c_double), ("dec", C.c_double),
                ("scanstart", C.c_double)]


@pytest.fixture(scope="module")
def H(pkg):
    L = pkg.hostlib()
    L.vf_vdif_frame_dmjd.restype = C.c_double
    L.vf_vdif_frame_dmjd.argtypes = [C.c_void_p, C.c_int]
    L.vf_vdif_to_unixepoch.restype = C.c_long
    L.vf_vdif_epoch_sec_offset.restype = C.c_uint32
    L.vf_sigproc_header_to_buffer.restype = C.c_long
    L.vf_sigproc_header_to_buffer.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.vf_ring_create.restype = C.c_void_p
    L.vf_ring_create.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p]
    for f in ("vf_ring_destroy", "vf_ring_shutdown", "vf_ring_end_of_data", "vf_ring_block_read_close"):
        getattr(L, f).argtypes = [C.c_void_p]
    L.vf_ring_get_nfull.restype = C.c_uint64; L.vf_ring_get_nfull.argtypes = [C.c_void_p]
    L.vf_ring_write.restype = C.c_ssize_t; L.vf_ring_write.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.vf_ring_read.restype = C.c_ssize_t; L.vf_ring_read.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.vf_ring_header_write.argtypes = [C.c_void_p, C.c_char_p]
    L.vf_ring_header_read.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    L.vf_ring_block_write_open.restype = C.c_void_p; L.vf_ring_block_write_open.argtypes = [C.c_void_p]
    L.vf_ring_block_write_close.argtypes = [C.c_void_p, C.c_uint64]
    L.vf_ring_block_read_open.restype = C.c_void_p
    L.vf_ring_block_read_open.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    return L


def make_header(H, seconds, frame, epoch=30, station=7, thread=1):
    h = VdifHeader()
    H.vf_vdif_set(C.byref(h), seconds, frame, epoch, station, thread)
    return h


def test_vdif_bit_layout(H):
    """analysis/baseband.py:19-28 of the reference: seconds word0[0:30], frame word1[0:24],
    epoch word1[24:30], length/8 word2[0:24], station word3[0:16], thread word3[16:26]"""
    h = make_header(H, 123456789 % (1 << 30), 25599, epoch=30, station=0x1234, thread=1)
    assert h.w[0] & 0x3FFFFFFF == 123456789 % (1 << 30)
    assert h.w[1] & 0xFFFFFF == 25599 and (h.w[1] >> 24) & 0x3F == 30
    assert (h.w[2] & 0xFFFFFF) * 8 == 5032
    assert h.w[3] & 0xFFFF == 0x1234 and (h.w[3] >> 16) & 0x3FF == 1 and (h.w[3] >> 26) & 0x1F == 7
    assert H.vf_vdif_thread_id(C.byref(h)) == 1 and H.vf_vdif_frame_number(C.byref(h)) == 25599
    assert H.vf_vdif_station_id(C.byref(h)) == 0x1234 and H.vf_vdif_frame_bytes(C.byref(h)) == 5032
    assert H.vf_vdif_frame_second(C.byref(h)) == (123456789 % (1 << 30)) % 86400


def test_vdif_time(H):
    # epoch 30 = 2015-01-01 (MJD 57023, unix 1420070400); epoch 31 = 2015-07-01 (MJD 57204)
    h = make_header(H, 86400 * 3 + 3661, 12800, epoch=30)
    assert H.vf_vdif_frame_mjd(C.byref(h)) == 57026
    assert H.vf_vdif_frame_mjd_sec(C.byref(h)) == 3661
    assert H.vf_vdif_to_unixepoch(C.byref(h)) == 1420070400 + 86400 * 3 + 3661
    assert abs(H.vf_vdif_frame_dmjd(C.byref(h), 25600) - (57026 + (3661 + 0.5) / 86400)) < 1e-12
    h = make_header(H, 0, 0, epoch=31)
    assert H.vf_vdif_frame_mjd(C.byref(h)) == 57204


def parse_sigproc(buf):
    out, i = {}, 0

    def s():
        nonlocal i
        n = struct.unpack_from("<i", buf, i)[0]; i += 4
        v = buf[i:i + n].decode(); i += n
        return v
    assert s() == "HEADER_START"
    ints = {"barycentric", "telescope_id", "data_type", "nchans", "nbits", "nifs"}
    order = []
    while True:
        k = s()
        if k == "HEADER_END":
            break
        order.append(k)
        if k == "source_name":
            out[k] = s()
        elif k in ints:
            out[k] = struct.unpack_from("<i", buf, i)[0]; i += 4
        else:
            out[k] = struct.unpack_from("<d", buf, i)[0]; i += 8
    return out, order, i


def test_sigproc_header(H):
    obs = ObsInfo(b"B0329+54", 12, 0.9338, 0.9528, 0.0)
    h = make_header(H, 86400 * 10 + 100, 0, epoch=30, station=12, thread=0)
    buf = C.create_string_buffer(1024)
    n = H.vf_sigproc_header_to_buffer(buf, 1024, C.byref(obs), C.byref(h), 2, 1)
    assert n > 0
    d, order, used = parse_sigproc(buf.raw[:n])
    assert used == n
    # field order of src/process_baseband.cu:243-268
    assert order == ["source_name", "barycentric", "telescope_id", "src_raj", "src_dej", "data_type", "fch1",
                     "foff", "nchans", "nbits", "tstart", "tsamp", "nifs"]
    chbw = -64.0 / 6251
    assert d["source_name"] == "B0329+54" and d["telescope_id"] == 12 and d["data_type"] == 1
    assert d["fch1"] == 384 + (2155 - 0.5) * chbw and d["foff"] == chbw
    assert d["nchans"] == 4096 and d["nbits"] == 2 and d["nifs"] == 1
    assert d["tsamp"] == 12500 / 128000000 * 8
    assert abs(d["tstart"] - (57033 + 100 / 86400)) < 1e-12
    # ra 0.9338 rad = 3h34m..; dec 0.9528 rad = 54d35m..
    assert 33400 < d["src_raj"] < 33500 and 543400 < d["src_dej"] < 543700


def test_ascii_header_and_names(H):
    hdr = C.create_string_buffer(4096)
    obs = ObsInfo(b"J0534+2200", 3, 1.4597, 0.3842, 57000.5)
    h = make_header(H, 86400 + 3723, 0, epoch=30)
    assert H.vf_write_psrdada_header(hdr, C.byref(obs), C.byref(h), 2, 1, b"/data/x.fil") == 0
    text = hdr.value.decode()
    kv = {l.split()[0]: l.split(None, 1)[1] for l in text.strip().split("\n")}
    # key set of src/process_baseband.cu:172-197
    for k in ("STATIONID", "BEAM", "RA", "DEC", "NAME", "SCANSTART", "NCHAN", "BANDWIDTH", "CFREQ", "NPOL", "NBIT",
              "TSAMP", "UTC_START", "UNIXEPOCH", "VDIF_MJD", "VDIF_SEC", "SIGPROC_FILE"):
        assert k in kv, k
    assert kv["NCHAN"] == "4096" and kv["NBIT"] == "2" and kv["UTC_START"] == "2015-01-02-01:02:03"
    assert abs(float(kv["TSAMP"]) - 781.25) < 1e-9
    assert abs(float(kv["BANDWIDTH"]) - 4096 * -64 / 6251) < 1e-5
    assert abs(float(kv["CFREQ"]) - (384 + 0.5 * (2155 + 6250 - 1) * -64 / 6251)) < 1e-5
    v = C.c_int()
    assert H.vf_ascii_header_get(hdr, b"NCHAN", b"%d", C.byref(v)) == 1 and v.value == 4096
    assert H.vf_ascii_header_set(hdr, C.c_size_t(4096), b"NCHAN", b"%d", 17) == 0
    assert H.vf_ascii_header_get(hdr, b"NCHAN", b"%d", C.byref(v)) == 1 and v.value == 17
    assert H.vf_ascii_header_get(hdr, b"NOPE", b"%d", C.byref(v)) == -1
    name = C.create_string_buffer(256)
    H.vf_fb_filename(name, C.c_size_t(256), b"/mnt/ssd/fildata", C.byref(h), 3, 1)
    assert name.value == b"/mnt/ssd/fildata/20150102_010203_muos_ea03_kur.fil"
    H.vf_fb_filename(name, C.c_size_t(256), b".", C.byref(h), 11, 0)
    assert name.value == b"./20150102_010203_muos_ea11.fil"


def test_ring_byte_stream_and_eod(H):
    r = H.vf_ring_create(4, 1000, None)
    data = np.random.default_rng(0).integers(0, 256, 10500, dtype=np.uint8)
    got = []

    def writer():
        assert H.vf_ring_header_write(r, b"NAME test\nSTATIONID 5\n") == 0
        for i in range(0, data.size, 700):
            chunk = data[i:i + 700]
            assert H.vf_ring_write(r, chunk.ctypes.data, chunk.size) == chunk.size
        assert H.vf_ring_end_of_data(r) == 0

    t = threading.Thread(target=writer); t.start()
    hdr = C.create_string_buffer(4096)
    assert H.vf_ring_header_read(r, hdr, 5000) == 0 and b"STATIONID 5" in hdr.value
    buf = np.empty(333, np.uint8)
    while True:
        n = H.vf_ring_read(r, buf.ctypes.data, buf.size)
        got.append(buf[:n].copy())
        if n < buf.size:
            break
    t.join()
    assert np.array_equal(np.concatenate(got), data)
    assert H.vf_ring_read(r, buf.ctypes.data, buf.size) == 0          # stays at EOD
    assert H.vf_ring_header_read(r, hdr, 50) == 1                      # no new observation: timeout
    H.vf_ring_destroy(r)


def test_ring_blocks_and_backpressure(H):
    r = H.vf_ring_create(2, 64, None)
    state = {"written": 0}

    def writer():
        H.vf_ring_header_write(r, b"NAME blocks\n")
        for i in range(6):
            p = H.vf_ring_block_write_open(r)
            C.memset(p, i + 1, 64)
            H.vf_ring_block_write_close(r, 64 if i < 5 else 10)
            state["written"] = i + 1
        H.vf_ring_end_of_data(r)

    t = threading.Thread(target=writer); t.start()
    hdr = C.create_string_buffer(4096)
    assert H.vf_ring_header_read(r, hdr, 5000) == 0
    import time
    time.sleep(0.2)
    assert state["written"] == 2 and H.vf_ring_get_nfull(r) == 2       # the writer blocks on a full ring
    sizes = []
    for i in range(6):
        nb = C.c_uint64()
        p = H.vf_ring_block_read_open(r, C.byref(nb))
        assert p and C.string_at(p, 1) == bytes([i + 1])
        sizes.append(nb.value)
        H.vf_ring_block_read_close(r)
    nb = C.c_uint64()
    assert not H.vf_ring_block_read_open(r, C.byref(nb))               # EOD
    t.join()
    assert sizes == [64] * 5 + [10]
    H.vf_ring_destroy(r)


def test_generator_is_deterministic_and_framed(pkg):
    g = pkg.GenParams.default(seed=5, rfi_amp=30, rfi_burst_every=2, drop_period=50, drop_len=1)
    a = pkg.gen_samples(g, 2, 1, 1000, 50000)
    b = pkg.gen_samples(g, 2, 1, 1000, 50000)
    assert np.array_equal(a, b)
    assert np.array_equal(pkg.gen_samples(g, 2, 1, 21000, 10000), a[20000:30000])   # pure function of the index
    assert not np.array_equal(pkg.gen_samples(g, 3, 1, 1000, 50000), a)
    assert 100 < a[a != 0].mean() < 156 and 10 < a[a != 0].std() < 30
    fr = pkg.gen_vdif_second(g, 2, 77, 10, 4).reshape(8, 5032)
    for i in range(8):
        w = fr[i, :32].view(np.uint32)
        assert w[0] == 77 and (w[1] & 0xFFFFFF) == 10 + i // 2 and ((w[3] >> 16) & 0x3FF) == i % 2 and (w[3] & 0xFFFF) == 2
        want = pkg.gen_samples(g, 2, i % 2, (77 * 25600 + 10 + i // 2) * 5000, 5000)
        assert np.array_equal(fr[i, 32:], want)


def test_control_commands_over_udp(H):
    """get_cmds / test_for_cmd of src/utils.c:174-220 over a UDP socket (unicast here: the build
    container has no multicast route; a multicast group additionally joins the group)"""
    H.vf_mc_open.argtypes = [C.c_char_p, C.c_int]
    H.vf_mc_send.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int]
    H.vf_get_cmds.argtypes = [C.POINTER(C.c_int * 5), C.c_int]
    import socket as pysock, time
    s = pysock.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    sock = H.vf_mc_open(b"127.0.0.1", port)
    assert sock >= 0
    assert H.vf_test_for_cmd(ord("Q"), sock) == 0                # nothing queued, does not block
    assert H.vf_mc_send(b"127.0.0.1", port, b"NS", 2) == 2
    time.sleep(0.05)
    cmds = (C.c_int * 5)()
    H.vf_get_cmds(C.byref(cmds), sock)
    assert list(cmds) == [1, 0, 0, 0, 1]
    assert H.vf_mc_send(b"127.0.0.1", port, b"E", 1) == 1
    assert H.vf_mc_send(b"127.0.0.1", port, b"Q", 1) == 1
    time.sleep(0.05)
    assert H.vf_test_for_cmd(ord("Q"), sock) == 1                # searches every queued datagram
    assert H.vf_test_for_cmd(ord("Q"), sock) == 0
    H.vf_mc_close(sock)


# ---- shared-memory ring (psrdada-style, between processes) -------------------
_SHM_WRITER = r"""
import ctypes as C, sys
L = C.CDLL(sys.argv[1])
L.vf_ring_connect_shm.restype = C.c_void_p; L.vf_ring_connect_shm.argtypes = [C.c_int]
L.vf_ring_header_write.argtypes = [C.c_void_p, C.c_char_p]
L.vf_ring_block_write_open.restype = C.c_void_p; L.vf_ring_block_write_open.argtypes = [C.c_void_p]
L.vf_ring_block_write_close.argtypes = [C.c_void_p, C.c_uint64]
L.vf_ring_end_of_data.argtypes = [C.c_void_p]; L.vf_ring_destroy.argtypes = [C.c_void_p]
r = L.vf_ring_connect_shm(int(sys.argv[2]))
assert r
assert L.vf_ring_header_write(r, b"NAME shm\nSTATIONID 9\n") == 0
for i in range(7):
    p = L.vf_ring_block_write_open(r)
    assert p
    C.memset(p, i + 1, 4096)
    assert L.vf_ring_block_write_close(r, 4096 if i < 6 else 100) == 0
assert L.vf_ring_end_of_data(r) == 0
L.vf_ring_destroy(r)
"""


def _shm_api(H):
    H.vf_ring_create_shm.restype = C.c_void_p; H.vf_ring_create_shm.argtypes = [C.c_int, C.c_uint64, C.c_uint64]
    H.vf_ring_connect_shm.restype = C.c_void_p; H.vf_ring_connect_shm.argtypes = [C.c_int]
    H.vf_ring_remove_shm.argtypes = [C.c_int]
    H.vf_ring_data_base.restype = C.c_void_p; H.vf_ring_data_base.argtypes = [C.c_void_p]
    H.vf_ring_get_nbufs.restype = C.c_uint64; H.vf_ring_get_nbufs.argtypes = [C.c_void_p]
    H.vf_ring_get_bufsz.restype = C.c_uint64; H.vf_ring_get_bufsz.argtypes = [C.c_void_p]


def test_shm_ring_between_processes(H, pkg, tmp_path):
    """a ring in SysV shared memory (what psrdada's dada_db makes): another process attaches by key,
    writes an observation, this one reads it with back-pressure (3 blocks for 7 written)"""
    import os, subprocess, sys
    _shm_api(H)
    key = 0x6000 + os.getpid() % 0x1000
    H.vf_ring_remove_shm(key)
    r = H.vf_ring_create_shm(key, 3, 4096)
    assert r
    assert not H.vf_ring_create_shm(key, 3, 4096)                     # the key exists: refused, like dada_db
    try:
        assert H.vf_ring_get_nbufs(r) == 3 and H.vf_ring_get_bufsz(r) == 4096
        script = tmp_path / "writer.py"
        script.write_text(_SHM_WRITER)
        w = subprocess.Popen([sys.executable, str(script), pkg.hostlib()._name, str(key)])
        hdr = C.create_string_buffer(4096)
        assert H.vf_ring_header_read(r, hdr, 20000) == 0 and b"STATIONID 9" in hdr.value
        nb = C.c_uint64()
        for i in range(7):
            p = H.vf_ring_block_read_open(r, C.byref(nb))
            assert p and C.string_at(p, 2) == bytes([i + 1]) * 2
            assert nb.value == (4096 if i < 6 else 100)
            H.vf_ring_block_read_close(r)
        assert not H.vf_ring_block_read_open(r, C.byref(nb))          # end of data
        assert w.wait(timeout=20) == 0
        assert H.vf_ring_header_read(r, hdr, 50) == 1                  # no second observation
    finally:
        H.vf_ring_destroy(r)                                           # the creator removes the segment
    assert not H.vf_ring_connect_shm(key)


def test_dada_db_and_genbase_into_ring(H, pkg, tmp_path):
    """vf_dada_db makes the ring, genbase -k fills it from another process; the blocks equal genbase -o"""
    import os, subprocess
    from conftest import ROOT
    BIN = os.path.join(ROOT, "vlite-fast_b200", "bin")
    if not os.path.exists(os.path.join(BIN, "vf_dada_db")):
        pytest.skip("executables not built")
    _shm_api(H)
    key = 0x7000 + os.getpid() % 0x1000
    subprocess.run([os.path.join(BIN, "vf_dada_db"), "-k", "%x" % key, "-d"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    out = subprocess.run([os.path.join(BIN, "vf_dada_db"), "-k", "%x" % key, "-n", "2"], check=True, stdout=subprocess.PIPE, text=True).stdout
    assert "nbufs=2 bufsz=257638400" in out
    try:
        vdif = tmp_path / "one.vdif"
        subprocess.run([os.path.join(BIN, "genbase"), "-o", str(vdif), "-t", "1", "-r", "5", "-n", "3"], check=True)
        w = subprocess.Popen([os.path.join(BIN, "genbase"), "-k", "%x" % key, "-t", "1", "-r", "5", "-n", "3"])
        r = H.vf_ring_connect_shm(key)
        assert r
        hdr = C.create_string_buffer(4096)
        assert H.vf_ring_header_read(r, hdr, 60000) == 0
        assert b"GENBASE" in hdr.value and b"STATIONID" in hdr.value
        nb = C.c_uint64()
        p = H.vf_ring_block_read_open(r, C.byref(nb))
        assert p and nb.value == 257638400
        got = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nb.value,))
        assert np.array_equal(got, np.fromfile(vdif, np.uint8))
        H.vf_ring_block_read_close(r)
        assert not H.vf_ring_block_read_open(r, C.byref(nb))
        assert w.wait(timeout=60) == 0
        H.vf_ring_destroy(r)
    finally:
        subprocess.run([os.path.join(BIN, "vf_dada_db"), "-k", "%x" % key, "-d"], check=True, stdout=subprocess.DEVNULL)
    assert not H.vf_ring_connect_shm(key)


def test_psrdada_backed_ring_type_checks():
    """the ring surface of process_baseband on psrdada's own calls (host/psrdada/vf_ring_psrdada.c)
    compiles against headers carrying psrdada's prototypes (compile-only: psrdada is not in the image)"""
    import os
    import subprocess
    from conftest import ROOT
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "vlite-fast_b200", "host"), "psrdada-check", "HOSTCC=/usr/bin/gcc"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "warning" not in r.stdout


def test_float_thresholds_equal_the_double_compares():
    """the 2-bit digitiser compares a float with double thresholds (src/pb_kernels.cu:660-663); the kernel uses the
    smallest float >= each threshold, which gives the same answer for every float"""
    for d, bits in ((-0.6109, 0xbf1c63f1), (0.3970, 0x3ecb4396), (1.4050, 0x3fb3d70b)):
        f = np.array([bits], np.uint32).view(np.float32)[0]
        assert float(f) >= d and float(np.nextafter(f, np.float32(-np.inf))) < d
        xs = np.nextafter(f, np.float32(-np.inf)), f, np.nextafter(f, np.float32(np.inf))
        for x in xs:
            assert (float(x) < d) == (x < f)
