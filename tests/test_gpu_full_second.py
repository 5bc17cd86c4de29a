"""-m gpu: whole antenna-seconds at the reference's geometry (10 segments of 1024 FFTs, carried bandpass) against
the reference's OWN kernels (oracle B, oracle/_ref): the north_star input -- genbase-style VDIF-rate baseband with
an injected DISPERSED pulse, canonical flags of scripts/baseband_test:21 -- and BASELINE.json configs[3], eight
antennas in one batched launch pair.  Bars: masks / weights / statistics bit-exact, detected power within 1e-5
(norm-wise), digitised samples within 1 LSB in < 1e-4 of the samples with no small-size allowance."""
import os

import numpy as np
import pytest

from conftest import ROOT, byte_diff

pytestmark = pytest.mark.gpu

HAVE_REF = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libvlite_ref.so"))
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built")

T, NSEG, NFFT = 1024, 10, 12500
REL, BYTE_FRAC = 1e-5, 1e-4


def strict_bytes(a, b, nbit, what):
    worst, frac = byte_diff(a, b, nbit)
    assert worst <= 1, "%s: code differs by %d" % (what, worst)
    assert frac < BYTE_FRAC, "%s: %.3g of samples differ" % (what, frac)
    return frac


@needs_ref
def test_canonical_dispersed_pulse_second_vs_reference_kernels(pkg, orc):
    """genbase -d 10 -a 0.05 -s 0.1 -p 0.5 -r 102 -f (scripts/baseband_test:21; chirp src/genbase.cu:525-552) ->
    process_baseband -b 2 -P 1 -r 2 (:26): all ten segments of the second through libvlitefast and through the
    reference kernels, segment by segment, the bandpass carried from one to the next in both."""
    with pkg.GpuGenerator(dm=10.0, pulse_period=0.5, ampl=(0.05, 0.005), seed=102, add_rfi=1) as gen:
        p0, p1 = gen.generate(NSEG * T * NFFT)
    r = orc.RefChain(2, 1, 2, keep_det=True)
    worst_frac, nmasked = 0.0, 0
    with pkg.Pipeline(ffts_per_seg=T, nbit=2, npol=1, rfi_mode=2, keep_stats=1, keep_power=1) as p:
        for s in range(NSEG):
            a, b = p0[s * T * NFFT:(s + 1) * T * NFFT], p1[s * T * NFFT:(s + 1) * T * NFFT]
            main, raw = p.process_segment(a, b)
            rmain, rraw = r.process_segment(a, b)
            mask = p.get_mask()
            assert np.array_equal(mask, r.mask()), s
            nmasked += int(np.count_nonzero(mask))
            st = p.get_stats()
            for k in ("pow", "kur", "dag", "pow_fb", "kur_fb", "dag_fb", "weights"):
                assert np.array_equal(st[k], r.get(k), equal_nan=True), (k, s)
            for which, name in ((0, "main"), (1, "raw")):
                det, rdet = p.get_detected_power(0, which), r.power_trimmed(name)
                assert np.abs(det - rdet).max() / rdet.mean() < REL, (name, s)
                l2 = np.sqrt(((det - rdet).astype(np.float64) ** 2).sum() / (rdet.astype(np.float64) ** 2).sum())
                assert l2 < 1e-6, (name, s)
                assert np.abs(p.get_power_f32(0, which) - r.ave_trimmed(name)).max() < 2e-4, (name, s)
            worst_frac = max(worst_frac, strict_bytes(main, rmain, 2, "main seg %d" % s),
                             strict_bytes(raw, rraw, 2, "raw seg %d" % s))
    r.close()
    # the -f RFI of genbase (src/genbase.cu:671-687) must actually exercise the excision
    assert nmasked > NSEG * T // 4
    print("canonical second: %d of %d time steps excised somewhere, worst byte fraction %.2e" % (nmasked, NSEG * T, worst_frac))


@needs_ref
def test_eight_antennas_one_batched_launch_vs_reference_kernels(pkg, orc):
    """BASELINE.json configs[3]: 8 antennas x 10 segments in ONE vf_process_device call (one launch pair, the
    channeliser sees 80 (segment, antenna) pairs, the normaliser walks the segments in time order) against eight
    independent replays of the reference's kernel sequence."""
    import torch
    n = 8
    gen = dict(seed=102, rfi_amp=60, rfi_burst_every=16)
    g = pkg.GenParams.default(**gen)
    host = np.empty((NSEG, n, 2, T * NFFT), np.uint8)
    for a in range(n):
        for s in range(NSEG):
            for pol in range(2):
                pkg.gen_samples(g, a, pol, s * T * NFFT, T * NFFT, host[s, a, pol])
    d_in = torch.from_numpy(host).cuda()
    with pkg.Pipeline(ffts_per_seg=T, nbit=8, npol=1, rfi_mode=2, n_antennas=n, keep_power=1, max_batch_segments=NSEG) as p:
        d_main = torch.zeros((NSEG, n, p.out_bytes), dtype=torch.uint8, device="cuda")
        d_raw = torch.zeros_like(d_main)
        torch.cuda.synchronize()
        p.process_device(n, NSEG, d_in.data_ptr(), d_main.data_ptr(), d_raw.data_ptr())
        p.sync()
        m, rw = d_main.cpu().numpy(), d_raw.cpu().numpy()
        masks = [p.get_mask(a) for a in range(n)]
        dets = [p.get_detected_power(a, 0) for a in range(n)]
        aves = [p.get_power_f32(a, 0) for a in range(n)]
        bps = [p.get_bandpass(a, 0) for a in range(n)]
    del d_in
    worst = 0.0
    for a in range(n):
        r = orc.RefChain(8, 1, 2, keep_det=True)
        for s in range(NSEG):
            rmain, rraw = r.process_segment(host[s, a, 0], host[s, a, 1])
            worst = max(worst, strict_bytes(m[s, a], rmain, 8, "main ant %d seg %d" % (a, s)),
                        strict_bytes(rw[s, a], rraw, 8, "raw ant %d seg %d" % (a, s)))
        # state after the last segment
        assert np.array_equal(masks[a], r.mask()), a
        rdet = r.power_trimmed("main")
        assert np.abs(dets[a] - rdet).max() / rdet.mean() < REL, a
        assert np.abs(aves[a] - r.ave_trimmed("main")).max() < 2e-4, a
        rbp = r.get("bp_main").reshape(2, 6251)[:, 2155:2155 + 4096]
        np.testing.assert_allclose(bps[a], rbp, rtol=2e-5)
        r.close()
    assert not np.array_equal(m[:, 0], m[:, 1])
    print("8 antennas x 10 segments batched: worst byte fraction %.2e" % worst)
