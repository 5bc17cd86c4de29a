"""ctypes bindings of include/vlitefast.h (libvlitefast.so) and of the host
helpers (libvlitehost.so).  Every call goes through the C ABI; nothing here
computes."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

NFFT = 12500          # src/process_baseband.h:20
NCHANOUT = 4096       # CHANMAX-CHANMIN+1, :53-54
NSCRUNCH = 8          # :24
NSUB = 25             # NFFT/NKURTO
VD_FRM = 5032         # :16
VD_DAT = 5000         # :17
FRAMES_PER_SEC = 25600  # :19

VF_ERRORS = {1: "ARG", 20: "CUDA", 21: "NOMEM", 22: "STATE", 23: "VDIF", 24: "NCCL", 25: "NODEV"}


class VfError(RuntimeError):
    def __init__(self, code, detail=""):
        self.code = code
        super().__init__("libvlitefast error %d (%s): %s" % (code, VF_ERRORS.get(code, "?"), detail))


class VfConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "abi_version", "nfft", "nscrunch", "ffts_per_seg", "nkurto", "chanmin", "chanmax", "nbit",
        "npol", "rfi_mode", "do_histo", "keep_stats", "keep_power", "inject_frb", "gpu_id",
        "n_antennas", "k1_threads", "power_segments", "max_batch_segments", "numa_pin")] + [
        ("dag_thresh", C.c_double), ("min_weight", C.c_double)]


def _load(name):
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        raise ImportError("%s is not built: run `make` (or __graft_entry__.build()) first; "
                          "there is no Python or CPU fallback" % path)
    return C.CDLL(path, mode=C.RTLD_LOCAL)


_lib = {}
_hostlib = None


def lib(testing=False):
    """libvlitefast.so with argument types declared.  testing=True: libvlitefast_testing.so, the same sources
    built with -DVF_TESTING (monolithic channeliser variants + vf_debug_division, csrc/vf_testing.h) -- for the
    A/B tests only."""
    if testing in _lib:
        return _lib[testing]
    L = _load("libvlitefast_testing.so" if testing else "libvlitefast.so")
    vp, u8p, i, sz = C.c_void_p, C.c_void_p, C.c_int, C.c_size_t
    pp = C.POINTER(C.c_void_p)
    L.vf_config_default.argtypes = [C.POINTER(VfConfig)]
    L.vf_create.argtypes = [C.POINTER(VfConfig), C.POINTER(vp)]
    L.vf_destroy.argtypes = [vp]
    L.vf_strerror.argtypes = [i]; L.vf_strerror.restype = C.c_char_p
    L.vf_last_error.argtypes = [vp]; L.vf_last_error.restype = C.c_char_p
    L.vf_segment_out_bytes.argtypes = [vp]; L.vf_segment_out_bytes.restype = sz
    L.vf_segment_in_samples.argtypes = [vp]; L.vf_segment_in_samples.restype = sz
    L.vf_process_segment.argtypes = [vp, i, u8p, u8p, sz, u8p, u8p, C.POINTER(sz)]
    L.vf_process_batch.argtypes = [vp, i, pp, pp, sz, pp, pp]
    L.vf_process_vdif.argtypes = [vp, i, vp, sz, C.c_uint32, u8p, u8p, C.POINTER(sz)]
    L.vf_submit_async.argtypes = [vp, i, i, pp, pp, sz, pp, pp]
    L.vf_wait.argtypes = [vp, i]
    L.vf_submit_vdif_async.argtypes = [vp, i, i, vp, sz, C.c_uint32, u8p, u8p]
    L.vf_submit_vdif_block_async.argtypes = [vp, i, i, vp, sz, C.c_uint32, C.c_long, i, u8p, u8p]
    L.vf_vdif_report.argtypes = [vp, i, C.POINTER(C.c_uint * 5)]
    L.vf_reserve_vdif_blocks.argtypes = [vp, i]
    L.vf_submit_block_async.argtypes = [vp, i, i, i, vp, vp, vp]
    L.vf_process_device.argtypes = [vp, i, i, vp, vp, vp]
    L.vf_sync.argtypes = [vp]
    L.vf_set_serial.argtypes = [vp, i]
    L.vf_timer_begin.argtypes = [vp]
    L.vf_timer_end.argtypes = [vp, C.POINTER(C.c_float)]
    if testing:
        L.vf_debug_division.argtypes = [vp, vp, vp, vp, vp, sz]
    fp = C.POINTER(C.c_float)
    L.vf_last_elapsed_ms.argtypes = [vp, fp, fp, fp]
    L.vf_slot_elapsed_ms.argtypes = [vp, i, fp, fp, fp]
    L.vf_bind_thread_to_gpu.argtypes = [i, C.c_char_p, sz]
    L.vf_host_alloc.argtypes = [C.POINTER(vp), sz]
    L.vf_host_register.argtypes = [vp, sz]
    L.vf_host_unregister.argtypes = [vp]
    L.vf_host_free.argtypes = [vp]
    L.vf_get_stats.argtypes = [vp, i] + [vp] * 8
    L.vf_get_mask.argtypes = [vp, i, vp]
    L.vf_get_power_f32.argtypes = [vp, i, i, vp]
    L.vf_get_detected_power.argtypes = [vp, i, i, vp]
    L.vf_get_bandpass.argtypes = [vp, i, i, vp]
    L.vf_set_bandpass.argtypes = [vp, i, i, vp]
    L.vf_reset_bandpass.argtypes = [vp, i]
    L.vf_set_frb_injection.argtypes = [vp, i, C.c_float, C.c_float, C.c_float]
    L.vf_coadd_init.argtypes = [vp, i, i, vp]
    L.vf_coadd_unique_id.argtypes = [vp]
    L.vf_coadd_segment.argtypes = [vp, i, i, vp, vp]
    L.vf_coadd_batch.argtypes = [vp, i, i, i, vp, vp, i]
    _lib[testing] = L
    return L


class GenParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("pulse_period", C.c_int), ("pulse_width", C.c_int),
                ("pulse_amp_q8", C.c_int * 2), ("rfi_period", C.c_int), ("rfi_width", C.c_int),
                ("rfi_amp", C.c_int), ("rfi_burst_every", C.c_int), ("tone_step", C.c_int),
                ("tone_amp", C.c_int), ("drop_period", C.c_int), ("drop_len", C.c_int),
                ("drop_pol_skew", C.c_int)]

    @classmethod
    def default(cls, **kw):
        g = cls()
        hostlib().vf_gen_defaults(C.byref(g))
        for k, v in kw.items():
            if k == "pulse_amp_q8":
                g.pulse_amp_q8[0], g.pulse_amp_q8[1] = v
            else:
                setattr(g, k, v)
        return g


def hostlib():
    """libvlitehost.so (plain C helpers: generator, VDIF, SIGPROC, ring shim)."""
    global _hostlib
    if _hostlib is not None:
        return _hostlib
    L = _load("libvlitehost.so")
    L.vf_gen_defaults.argtypes = [C.POINTER(GenParams)]
    L.vf_gen_samples.argtypes = [C.POINTER(GenParams), C.c_int, C.c_int, C.c_uint64, C.c_size_t, C.c_void_p]
    L.vf_gen_vdif_second.argtypes = [C.POINTER(GenParams), C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
    L.vf_gen_vdif_second.restype = C.c_size_t
    _hostlib = L
    return L


def gen_samples(g, antenna, pol, sample0, n, out=None):
    """n unsigned 8-bit samples of (antenna, pol) from absolute sample index sample0."""
    if out is None:
        out = np.empty(n, dtype=np.uint8)
    assert out.dtype == np.uint8 and out.size >= n and out.flags.c_contiguous
    hostlib().vf_gen_samples(C.byref(g), antenna, pol, sample0, n, out.ctypes.data)
    return out


def gen_vdif_second(g, antenna, second, first_frame, nframes):
    out = np.empty(nframes * 2 * VD_FRM, dtype=np.uint8)
    n = hostlib().vf_gen_vdif_second(C.byref(g), antenna, second, first_frame, nframes, out.ctypes.data)
    assert n == out.size
    return out


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return a
    assert a.flags.c_contiguous
    return a.ctypes.data


def bind_thread_to_gpu(gpu_id):
    """vf_bind_thread_to_gpu: returns the CPU list applied ('' when the platform exposes none)"""
    buf = C.create_string_buffer(1024)
    rc = lib().vf_bind_thread_to_gpu(gpu_id, buf, 1024)
    if rc:
        raise VfError(rc, "vf_bind_thread_to_gpu")
    return buf.value.decode()


class Pipeline:
    """One vf_handle.  Keyword arguments are vf_config fields."""

    def __init__(self, testing=False, **kw):
        L = lib(testing)
        cfg = VfConfig()
        L.vf_config_default(C.byref(cfg))
        for k, v in kw.items():
            if not hasattr(cfg, k):
                raise TypeError("unknown vf_config field %r" % k)
            setattr(cfg, k, v)
        self.cfg = cfg
        self.L = L
        self.h = C.c_void_p()
        rc = L.vf_create(C.byref(cfg), C.byref(self.h))
        if rc:
            detail = L.vf_last_error(self.h).decode() if self.h else L.vf_strerror(rc).decode()
            if self.h:
                L.vf_destroy(self.h)
                self.h = C.c_void_p()
            raise VfError(rc, detail)
        self.T = cfg.ffts_per_seg
        self.ntime = self.T // NSCRUNCH
        self.n_ant = cfg.n_antennas
        self.nsamp = L.vf_segment_in_samples(self.h)
        self.out_bytes = L.vf_segment_out_bytes(self.h)

    def close(self):
        if getattr(self, "h", None):
            self.L.vf_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc:
            raise VfError(rc, self.L.vf_last_error(self.h).decode())

    # ---- processing -------------------------------------------------------
    def process_segment(self, pol0, pol1, antenna=0):
        """vf_process_segment: returns (fb_main, fb_raw or None)."""
        main = np.empty(self.out_bytes, np.uint8)
        raw = np.empty(self.out_bytes, np.uint8) if self.cfg.rfi_mode == 2 else None
        nb = C.c_size_t()
        self._ck(self.L.vf_process_segment(self.h, antenna, _ptr(pol0), _ptr(pol1), pol0.size,
                                           _ptr(main), _ptr(raw), C.byref(nb)))
        assert nb.value == self.out_bytes
        return main, raw

    def _arrays(self, lst):
        arr = (C.c_void_p * len(lst))()
        for i, a in enumerate(lst):
            arr[i] = _ptr(a)
        return arr

    def process_batch(self, pol0s, pol1s):
        n = len(pol0s)
        mains = [np.empty(self.out_bytes, np.uint8) for _ in range(n)]
        raws = [np.empty(self.out_bytes, np.uint8) for _ in range(n)] if self.cfg.rfi_mode == 2 else None
        self._ck(self.L.vf_process_batch(self.h, n, self._arrays(pol0s), self._arrays(pol1s), pol0s[0].size,
                                         self._arrays(mains), self._arrays(raws) if raws else None))
        return mains, raws

    def submit_async(self, slot, pol0s, pol1s, mains, raws=None):
        n = len(pol0s)
        self._ck(self.L.vf_submit_async(self.h, slot, n, self._arrays(pol0s), self._arrays(pol1s), self.nsamp,
                                        self._arrays(mains), self._arrays(raws) if raws else None))

    def submit_block_async(self, slot, n_ant, n_seg, block, mains, raws=None):
        """vf_submit_block_async: block is one host array [n_seg][n_ant][2][nsamp]; outputs [n_seg][n_ant][out_bytes]"""
        assert block.size == n_seg * n_ant * 2 * self.nsamp and mains.size == n_seg * n_ant * self.out_bytes
        self._ck(self.L.vf_submit_block_async(self.h, slot, n_ant, n_seg, _ptr(block), _ptr(mains), _ptr(raws)))

    def wait(self, slot):
        self._ck(self.L.vf_wait(self.h, slot))

    def process_vdif(self, frames, first_frame, antenna=0):
        main = np.empty(self.out_bytes, np.uint8)
        raw = np.empty(self.out_bytes, np.uint8) if self.cfg.rfi_mode == 2 else None
        nb = C.c_size_t()
        self._ck(self.L.vf_process_vdif(self.h, antenna, _ptr(frames), frames.size // VD_FRM, first_frame,
                                        _ptr(main), _ptr(raw), C.byref(nb)))
        return main, raw

    def process_vdif_block(self, frames, first_frame, expect_second, n_seg, antenna=0, slot=0):
        """vf_submit_vdif_block_async + vf_wait: returns (fb_main [n_seg][out_bytes], fb_raw or None, report, warned);
        report = (outside the block, other second, placed, invalid bit, frames of a complete block)"""
        main = np.empty((n_seg, self.out_bytes), np.uint8)
        raw = np.empty((n_seg, self.out_bytes), np.uint8) if self.cfg.rfi_mode == 2 else None
        self._ck(self.L.vf_submit_vdif_block_async(self.h, slot, antenna, _ptr(frames), frames.size // VD_FRM, first_frame,
                                                   expect_second, n_seg, _ptr(main), _ptr(raw)))
        rc = self.L.vf_wait(self.h, slot)
        if rc not in (0, 23):
            self._ck(rc)
        rep = (C.c_uint * 5)()
        self._ck(self.L.vf_vdif_report(self.h, slot, C.byref(rep)))
        return main, raw, tuple(rep), rc == 23

    def process_device(self, n_ant, n_seg, d_in, d_main, d_raw=None):
        """device pointers (ints); asynchronous, see sync()."""
        self._ck(self.L.vf_process_device(self.h, n_ant, n_seg, d_in, d_main, d_raw))

    def debug_division(self, p, b):
        p = np.ascontiguousarray(p, np.float32); b = np.ascontiguousarray(b, np.float32)
        qp = np.empty_like(p); qr = np.empty_like(p)
        self._ck(self.L.vf_debug_division(self.h, _ptr(p), _ptr(b), _ptr(qp), _ptr(qr), p.size))
        return qp, qr

    def set_serial(self, serial):
        self._ck(self.L.vf_set_serial(self.h, int(serial)))

    def sync(self):
        self._ck(self.L.vf_sync(self.h))

    def timer_begin(self):
        self._ck(self.L.vf_timer_begin(self.h))

    def timer_end(self):
        ms = C.c_float()
        self._ck(self.L.vf_timer_end(self.h, C.byref(ms)))
        return ms.value

    def last_elapsed_ms(self):
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        self._ck(self.L.vf_last_elapsed_ms(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def slot_elapsed_ms(self, slot):
        """(total, K1, K2) device ms of the last asynchronous submission on `slot`, after wait(slot)"""
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        self._ck(self.L.vf_slot_elapsed_ms(self.h, slot, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    # ---- inspection ---------------------------------------------------------
    def get_stats(self, antenna=0):
        T = self.T
        out = {}
        if self.cfg.keep_stats and self.cfg.rfi_mode:
            for k, n in (("pow", 2 * T * NSUB), ("kur", 2 * T * NSUB), ("dag", 2 * T * NSUB),
                         ("pow_fb", 2 * T), ("kur_fb", 2 * T), ("dag_fb", 2 * T)):
                out[k] = np.empty(n, np.float32)
        if self.cfg.rfi_mode:
            out["weights"] = np.empty(2 * T, np.float32)
        if self.cfg.do_histo:
            out["histo"] = np.empty(512, np.uint32)
        order = ("pow", "kur", "dag", "pow_fb", "kur_fb", "dag_fb", "weights", "histo")
        self._ck(self.L.vf_get_stats(self.h, antenna, *[_ptr(out.get(k)) for k in order]))
        return out

    def get_mask(self, antenna=0):
        m = np.empty(self.T, np.uint32)
        self._ck(self.L.vf_get_mask(self.h, antenna, _ptr(m)))
        return m

    def get_power_f32(self, antenna=0, which=0):
        out = np.empty((self.cfg.npol, self.ntime, NCHANOUT), np.float32)
        self._ck(self.L.vf_get_power_f32(self.h, antenna, which, _ptr(out)))
        return out

    def get_detected_power(self, antenna=0, which=0):
        out = np.empty((self.T, NCHANOUT, 2), np.float32)
        self._ck(self.L.vf_get_detected_power(self.h, antenna, which, _ptr(out)))
        return out

    def get_bandpass(self, antenna=0, which=0):
        out = np.empty((2, NCHANOUT), np.float32)
        self._ck(self.L.vf_get_bandpass(self.h, antenna, which, _ptr(out)))
        return out

    def set_bandpass(self, bp, antenna=0, which=0):
        bp = np.ascontiguousarray(bp, np.float32)
        assert bp.shape == (2, NCHANOUT)
        self._ck(self.L.vf_set_bandpass(self.h, antenna, which, _ptr(bp)))

    def reset_bandpass(self, antenna=-1):
        self._ck(self.L.vf_reset_bandpass(self.h, antenna))

    def set_frb_injection(self, nfft_since_frb, dm=80.0, width=20.48, amp=1.05):
        self._ck(self.L.vf_set_frb_injection(self.h, nfft_since_frb, dm, width, amp))

    # ---- co-add -------------------------------------------------------------
    def coadd_unique_id(self):
        buf = (C.c_char * 128)()
        rc = self.L.vf_coadd_unique_id(buf)
        if rc:
            raise VfError(rc, "ncclGetUniqueId")
        return bytes(buf)

    def coadd_init(self, nranks=1, rank=0, unique_id=None):
        buf = C.create_string_buffer(unique_id, 128) if unique_id else None
        self._ck(self.L.vf_coadd_init(self.h, nranks, rank, buf))

    def coadd_batch(self, root, total_antennas, n_seg, want=True, wait=True):
        n = self.cfg.npol * self.ntime * NCHANOUT
        fb = np.empty((n_seg, n * self.cfg.nbit // 8), np.uint8) if want else None
        sm = np.empty((n_seg, self.cfg.npol, self.ntime, NCHANOUT), np.float32) if want else None
        self._ck(self.L.vf_coadd_batch(self.h, root, total_antennas, n_seg, _ptr(fb), _ptr(sm), 1 if (wait or want) else 0))
        return fb, sm

    def coadd_batch_into(self, root, total_antennas, n_seg, fb_host=None, sum_host=None, wait=False):
        """vf_coadd_batch into caller-owned (pinned) host arrays; None = not wanted (every rank but the root)"""
        self._ck(self.L.vf_coadd_batch(self.h, root, total_antennas, n_seg, _ptr(fb_host), _ptr(sum_host), 1 if wait else 0))

    def coadd_segment(self, root, total_antennas, want=True):
        n = self.cfg.npol * self.ntime * NCHANOUT
        fb = np.empty(n * self.cfg.nbit // 8, np.uint8) if want else None
        sm = np.empty((self.cfg.npol, self.ntime, NCHANOUT), np.float32) if want else None
        self._ck(self.L.vf_coadd_segment(self.h, root, total_antennas, _ptr(fb), _ptr(sm)))
        return fb, sm


# ---- libvlitegen.so: GPU baseband generator (include/vlitegen.h) -------------
class VfgConfig(C.Structure):
    _fields_ = [("dm", C.c_double), ("pulse_period", C.c_double), ("ampl", C.c_float * 2), ("skip_period", C.c_int),
                ("add_rfi", C.c_int), ("seed", C.c_ulonglong), ("buflen", C.c_longlong), ("gpu_id", C.c_int),
                ("reserved", C.c_int * 7)]


_genlib = None


def genlib():
    global _genlib
    if _genlib is not None:
        return _genlib
    L = _load("libvlitegen.so")
    vp = C.c_void_p
    L.vfg_config_default.argtypes = [C.POINTER(VfgConfig)]
    L.vfg_create.argtypes = [C.POINTER(VfgConfig), C.POINTER(vp)]
    L.vfg_destroy.argtypes = [vp]
    L.vfg_last_error.argtypes = [vp]; L.vfg_last_error.restype = C.c_char_p
    L.vfg_block_samples.argtypes = [vp]; L.vfg_block_samples.restype = C.c_longlong
    L.vfg_sweep_samples.argtypes = [vp]; L.vfg_sweep_samples.restype = C.c_longlong
    L.vfg_generate.argtypes = [vp, vp, vp, C.c_size_t]
    L.vfg_last_block_f32.argtypes = [vp, C.c_int, vp]
    L.vfg_generate_vdif_second.argtypes = [vp, C.c_int, C.c_uint32, vp]
    _genlib = L
    return L


class GpuGenerator:
    """vfg_* handle: coherent-dispersion baseband generator on the GPU."""

    def __init__(self, **kw):
        self.L = genlib()
        self.cfg = VfgConfig()
        self.L.vfg_config_default(C.byref(self.cfg))
        for k, v in kw.items():
            if k == "ampl":
                self.cfg.ampl[0], self.cfg.ampl[1] = v
            else:
                setattr(self.cfg, k, v)
        self.h = C.c_void_p()
        rc = self.L.vfg_create(C.byref(self.cfg), C.byref(self.h))
        if rc:
            msg = self.L.vfg_last_error(self.h).decode() if self.h else "no CUDA device"
            self.L.vfg_destroy(self.h)
            self.h = None
            raise VfError(rc, msg)
        self.block_samples = self.L.vfg_block_samples(self.h)
        self.sweep_samples = self.L.vfg_sweep_samples(self.h)

    def _ck(self, rc):
        if rc:
            raise VfError(rc, self.L.vfg_last_error(self.h).decode())

    def generate(self, n):
        p0, p1 = np.empty(n, np.uint8), np.empty(n, np.uint8)
        self._ck(self.L.vfg_generate(self.h, p0.ctypes.data, p1.ctypes.data, n))
        return p0, p1

    def last_block_f32(self, pol):
        out = np.empty(self.block_samples, np.float32)
        self._ck(self.L.vfg_last_block_f32(self.h, pol, out.ctypes.data))
        return out

    def vdif_second(self, station, second):
        out = np.empty(FRAMES_PER_SEC * 2 * VD_FRM, np.uint8)
        self._ck(self.L.vfg_generate_vdif_second(self.h, station, second, out.ctypes.data))
        return out

    def close(self):
        if self.h:
            self.L.vfg_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
