"""TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

ctypes views of the CPU oracle (oracle/liboracle.so, "oracle A") and of the
reference-kernel replay (oracle/_ref/libvlite_ref.so, "oracle B", needs a GPU).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
legs may import this module."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
NFFT, NCHAN, NCHANOUT, CHANMIN, NSUB, NSCRUNCH = 12500, 6251, 4096, 2155, 25, 8

_liba = None
_libb = None


def liba():
    global _liba
    if _liba is None:
        p = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(p):
            raise ImportError("oracle/liboracle.so not built: make -C oracle")
        L = C.CDLL(p)
        vp, i, sz = C.c_void_p, C.c_int, C.c_size_t
        L.orc_create.argtypes = [i, i, i, i, i]; L.orc_create.restype = vp
        L.orc_destroy.argtypes = [vp]
        L.orc_reset_bandpass.argtypes = [vp]
        L.orc_out_bytes.argtypes = [vp]; L.orc_out_bytes.restype = sz
        L.orc_process_segment.argtypes = [vp, vp, vp, vp, vp]
        for n in ("pow", "kur", "dag", "pow_fb", "kur_fb", "dag_fb", "weights", "ave_main", "ave_raw",
                  "power_main", "power_raw", "bp_main", "bp_raw", "histo"):
            f = getattr(L, "orc_get_" + n)
            f.argtypes = [vp]; f.restype = vp
        L.orc_set_bandpass.argtypes = [vp, i, vp]
        L.orc_fft_plan_create.argtypes = [i]; L.orc_fft_plan_create.restype = vp
        L.orc_fft_plan_destroy.argtypes = [vp]
        L.orc_fft_scratch_floats.argtypes = [vp]; L.orc_fft_scratch_floats.restype = sz
        L.orc_rfft.argtypes = [vp, vp, vp, vp]
        L.orc_stage_convert.argtypes = [vp, vp, sz]
        L.orc_stage_kurtosis.argtypes = [vp, vp, vp, sz]
        L.orc_stage_dagostino.argtypes = [vp, vp, sz, i]
        L.orc_digitise.argtypes = [vp, vp, i, i, i]
        L.orc_dagostino_constants.argtypes = [i, vp]
        _liba = L
    return _liba


def _view(ptr, n, dtype=np.float32):
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float if dtype == np.float32 else C.c_uint32)),
                                 shape=(n,)).copy()


class OracleChain:
    """CPU restatement of the chain (vlite_oracle.c)."""

    def __init__(self, ffts_per_seg=1024, nbit=2, npol=1, rfi_mode=2, nthreads=0):
        self.L = liba()
        self.T, self.nbit, self.npol, self.mode = ffts_per_seg, nbit, npol, rfi_mode
        self.c = self.L.orc_create(ffts_per_seg, nbit, npol, rfi_mode, nthreads)
        if not self.c:
            raise ValueError("orc_create rejected the configuration")
        self.out_bytes = self.L.orc_out_bytes(self.c)

    def close(self):
        if self.c:
            self.L.orc_destroy(self.c)
            self.c = None

    def __del__(self):
        self.close()

    def reset_bandpass(self):
        self.L.orc_reset_bandpass(self.c)

    def process_segment(self, pol0, pol1):
        assert pol0.dtype == np.uint8 and pol0.size == self.T * NFFT and pol1.size == pol0.size
        main = np.empty(self.out_bytes, np.uint8)
        raw = np.empty(self.out_bytes, np.uint8) if self.mode == 2 else None
        rc = self.L.orc_process_segment(self.c, pol0.ctypes.data, pol1.ctypes.data, main.ctypes.data,
                                        raw.ctypes.data if raw is not None else None)
        assert rc == 0
        return main, raw

    def get(self, name):
        T = self.T
        sizes = {"pow": 2 * T * NSUB, "kur": 2 * T * NSUB, "dag": 2 * T * NSUB, "pow_fb": 2 * T,
                 "kur_fb": 2 * T, "dag_fb": 2 * T, "weights": 2 * T,
                 "ave_main": self.npol * (T // 8) * NCHAN, "ave_raw": self.npol * (T // 8) * NCHAN,
                 "power_main": 2 * T * NCHAN, "power_raw": 2 * T * NCHAN, "bp_main": 2 * NCHAN,
                 "bp_raw": 2 * NCHAN, "histo": 512}
        ptr = getattr(self.L, "orc_get_" + name)(self.c)
        return _view(ptr, sizes[name], np.uint32 if name == "histo" else np.float32)

    def set_bandpass(self, which, bp):
        bp = np.ascontiguousarray(bp, np.float32)
        assert bp.size == 2 * NCHAN
        self.L.orc_set_bandpass(self.c, which, bp.ctypes.data)

    # trimmed views in the library's layouts
    def ave_trimmed(self, which="main"):
        a = self.get("ave_" + which).reshape(self.npol, self.T // 8, NCHAN)
        return a[:, :, CHANMIN:CHANMIN + NCHANOUT].copy()

    def power_trimmed(self, which="main"):
        p = self.get("power_" + which).reshape(2, self.T, NCHAN)[:, :, CHANMIN:CHANMIN + NCHANOUT]
        return np.ascontiguousarray(p.transpose(1, 2, 0))     # [T][4096][2]

    def mask(self):
        """uint32[T]: bit j = sub-block j excised (dag > 3)."""
        d = self.get("dag")[: self.T * NSUB].reshape(self.T, NSUB)
        bits = (d > np.float32(3.0)).astype(np.uint32)
        return (bits << np.arange(NSUB, dtype=np.uint32)).sum(axis=1).astype(np.uint32)


def rfft(x):
    """orc_rfft of float32 rows; returns complex64 [.., n/2+1]."""
    L = liba()
    x = np.ascontiguousarray(x, np.float32)
    n = x.shape[-1]
    pl = L.orc_fft_plan_create(n)
    assert pl
    scr = np.empty(L.orc_fft_scratch_floats(pl), np.float32)
    rows = x.reshape(-1, n)
    out = np.empty((rows.shape[0], n // 2 + 1), np.complex64)
    for r in range(rows.shape[0]):
        L.orc_rfft(pl, rows[r].ctypes.data, out[r].ctypes.data, scr.ctypes.data)
    L.orc_fft_plan_destroy(pl)
    return out.reshape(x.shape[:-1] + (n // 2 + 1,))


def libb():
    global _libb
    if _libb is None:
        p = os.path.join(_HERE, "_ref", "libvlite_ref.so")
        if not os.path.exists(p):
            raise ImportError("oracle/_ref/libvlite_ref.so not built (needs /root/reference at build time)")
        L = C.CDLL(p)
        vp, i = C.c_void_p, C.c_int
        L.ref_ffts_per_seg.restype = i
        L.ref_create.argtypes = [i, i, i, i, i, i, C.POINTER(vp)]
        L.ref_reset_bandpass.argtypes = [vp]
        L.ref_out_bytes.argtypes = [vp]; L.ref_out_bytes.restype = C.c_size_t
        L.ref_process_segment.argtypes = [vp, vp, vp, vp, vp, i]
        L.ref_time_device.argtypes = [vp, vp, i, C.POINTER(C.c_float)]
        L.ref_last_ms.argtypes = [vp]; L.ref_last_ms.restype = C.c_float
        L.ref_get.argtypes = [vp, i, vp]; L.ref_get.restype = C.c_long
        L.ref_destroy.argtypes = [vp]
        _libb = L
    return _libb


class RefChain:
    """The reference's own kernels + cuFFT in the reference's launch order (GPU)."""
    WHICH = {"pow": 0, "kur": 1, "dag": 2, "pow_fb": 3, "kur_fb": 4, "dag_fb": 5, "weights": 6,
             "ave_main": 7, "ave_raw": 8, "power_main": 9, "power_raw": 10, "bp_main": 11, "bp_raw": 12,
             "histo": 13, "weights_work": 14}

    def __init__(self, nbit=2, npol=1, rfi_mode=2, keep_det=False, do_histo=False, inject_frb=False):
        self.L = libb()
        self.T = self.L.ref_ffts_per_seg()
        self.nbit, self.npol, self.mode = nbit, npol, rfi_mode
        self.c = C.c_void_p()
        rc = self.L.ref_create(nbit, npol, rfi_mode, int(keep_det), int(do_histo), int(inject_frb), C.byref(self.c))
        if rc:
            raise RuntimeError("ref_create failed: %d" % rc)
        self.out_bytes = self.L.ref_out_bytes(self.c)

    def close(self):
        if self.c:
            self.L.ref_destroy(self.c)
            self.c = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset_bandpass(self):
        assert self.L.ref_reset_bandpass(self.c) == 0

    def process_segment(self, pol0, pol1, inject_frb_now=0):
        assert pol0.size == self.T * NFFT
        main = np.empty(self.out_bytes, np.uint8)
        raw = np.empty(self.out_bytes, np.uint8) if self.mode == 2 else None
        rc = self.L.ref_process_segment(self.c, pol0.ctypes.data, pol1.ctypes.data, main.ctypes.data,
                                        raw.ctypes.data if raw is not None else None, inject_frb_now)
        if rc:
            raise RuntimeError("ref_process_segment failed: %d" % rc)
        return main, raw

    def time_device(self, d_in, nseg):
        ms = C.c_float()
        rc = self.L.ref_time_device(self.c, d_in, nseg, C.byref(ms))
        if rc:
            raise RuntimeError("ref_time_device failed: %d" % rc)
        return ms.value

    def get(self, name):
        T = self.T
        n = max(2 * T * NCHAN, 2 * T * NSUB)
        buf = np.empty(n, np.uint32 if name == "histo" else np.float32)
        got = self.L.ref_get(self.c, self.WHICH[name], buf.ctypes.data)
        if got < 0:
            raise RuntimeError("ref_get(%s) unavailable" % name)
        return buf[:got].copy()

    def ave_trimmed(self, which="main"):
        a = self.get("ave_" + which).reshape(self.npol, self.T // 8, NCHAN)
        return a[:, :, CHANMIN:CHANMIN + NCHANOUT].copy()

    def power_trimmed(self, which="main"):
        p = self.get("power_" + which).reshape(2, self.T, NCHAN)[:, :, CHANMIN:CHANMIN + NCHANOUT]
        return np.ascontiguousarray(p.transpose(1, 2, 0))

    def mask(self):
        d = self.get("dag")[: self.T * NSUB].reshape(self.T, NSUB)
        bits = (d > np.float32(3.0)).astype(np.uint32)
        return (bits << np.arange(NSUB, dtype=np.uint32)).sum(axis=1).astype(np.uint32)
