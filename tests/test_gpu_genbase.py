"""-m gpu: libvlitegen (GPU baseband generator) against its numpy restatement, its streaming
and VDIF framing, and the dispersed pulse it injects seen through the channeliser."""
import numpy as np
import pytest

from test_genbase_oracle import load_gen_oracle

pytestmark = pytest.mark.gpu
SMALL = dict(dm=0.05, pulse_period=0.002, ampl=(0.5, 0.05), seed=11, buflen=1 << 20)


@pytest.mark.parametrize("rfi", [0, 1])
def test_blocks_match_the_numpy_restatement(pkg, rfi):
    g = load_gen_oracle()
    with pkg.GpuGenerator(add_rfi=rfi, **SMALL) as gen:
        n = gen.block_samples
        n_lo, n_hi = g.sweep_samples(SMALL["dm"])
        assert gen.sweep_samples == n_lo + n_hi and n == SMALL["buflen"] - n_lo - n_hi
        for blk in range(2):
            p0, p1 = gen.generate(n)
            for pol, got in ((0, p0), (1, p1)):
                want_u, want_v = g.block(blk, pol, dm=SMALL["dm"], pulse_period=SMALL["pulse_period"], ampl=SMALL["ampl"][pol],
                                         add_rfi=rfi, seed=SMALL["seed"], buflen=SMALL["buflen"])
                v = gen.last_block_f32(pol)
                # voltages: fp32 FFT of 2^20 points against float64, 1e-5 of the rms
                assert np.abs(v - want_v).max() < 2e-5 * want_v.std() + 2e-5
                d = np.abs(got.astype(int) - want_u.astype(int))
                assert d.max() <= 1 and (d != 0).mean() < 1e-3, (d.max(), (d != 0).mean())
                assert abs(got.astype(float).mean() - 128) < 0.2          # digitiser centred on 128 (:700-706)


def test_stream_is_independent_of_the_request_size(pkg):
    with pkg.GpuGenerator(**SMALL) as a, pkg.GpuGenerator(**SMALL) as b:
        n = a.block_samples
        whole0, whole1 = a.generate(2 * n + 777)
        parts0, parts1 = [], []
        for m in (1000, n - 3, 5, n - 225):
            q0, q1 = b.generate(m)
            parts0.append(q0); parts1.append(q1)
        assert np.array_equal(np.concatenate(parts0), whole0) and np.array_equal(np.concatenate(parts1), whole1)


def test_vdif_second_frames(pkg):
    cfg = dict(dm=1.0, pulse_period=0.01, ampl=(0.3, 0.3), seed=5, buflen=1 << 24)
    with pkg.GpuGenerator(**cfg) as a, pkg.GpuGenerator(**cfg) as b:
        blk = a.vdif_second(7, 18000).reshape(-1, pkg.VD_FRM)
        p0, p1 = b.generate(128000000)
        hdr = blk[:, :32].copy().view(np.uint32)
        assert np.all(hdr[:, 0] == 18000) and np.array_equal(hdr[0::2, 1] & 0xFFFFFF, np.arange(25600))
        assert np.all((hdr[0::2, 3] >> 16) & 0x3FF == 0) and np.all((hdr[1::2, 3] >> 16) & 0x3FF == 1)
        assert np.all(hdr[:, 3] & 0xFFFF == 7) and np.all((hdr[:, 2] & 0xFFFFFF) * 8 == 5032)
        assert np.array_equal(blk[0::2, 32:].reshape(-1), p0) and np.array_equal(blk[1::2, 32:].reshape(-1), p1)


def test_dispersed_pulse_through_the_channeliser(pkg):
    """a DM 10 pulse every 0.1 s: in the detected power the pulse arrives later at lower sky frequency (higher
    channel number, SURVEY.md 8a note 9) by the cold-plasma delay between the two sub-bands"""
    dm, period = 10.0, 0.1
    with pkg.GpuGenerator(dm=dm, pulse_period=period, ampl=(3.0, 3.0), seed=3, buflen=1 << 25) as gen:
        T = 1024
        with pkg.Pipeline(ffts_per_seg=T, nbit=8, npol=1, rfi_mode=0) as p:
            prof = []
            for s in range(4):
                p0, p1 = gen.generate(T * pkg.NFFT)
                p.process_segment(p0, p1)
                det = p.get_detected_power(0, 0).reshape(T, pkg.NCHANOUT, 2).sum(axis=2)
                prof.append(np.stack([det[:, :512].mean(axis=1), det[:, -512:].mean(axis=1)]))
    prof = np.concatenate(prof, axis=1)                    # [2 sub-bands][4096 time steps of 97.66 us]
    step = pkg.NFFT / 128e6
    nper = int(round(period / step))
    fold = prof[:, :(prof.shape[1] // nper) * nper].reshape(2, -1, nper).mean(axis=1)
    fold = fold / np.median(fold, axis=1, keepdims=True)
    assert fold.max(axis=1).min() > 1.5                    # the pulse is there in both sub-bands
    # centre of the on-pulse window in each sub-band
    def centre(x):
        on = x > 0.5 * (x.max() + 1)
        idx = np.arange(nper)[on]
        return idx.mean()
    # channel c <-> 384 - (2155 + c) * 64 / 6250 MHz
    f_hi = 384 - (2155 + 256) * 64 / 6250.
    f_lo = 384 - (2155 + 4096 - 256) * 64 / 6250.
    want = 4.149e3 * dm * (f_lo ** -2 - f_hi ** -2)        # seconds
    got = ((centre(fold[1]) - centre(fold[0])) % nper) * step
    assert abs(got - want) < 0.15 * want + 2 * step, (got, want)
