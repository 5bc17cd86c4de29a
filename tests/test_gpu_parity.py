"""-m gpu: libvlitefast (through the C ABI) against the CPU oracle on the same
seeded inputs.  Bars (BASELINE.json north_star): RFI masks and weights
bit-exact; float spectra / filterbank powers within 1e-5 relative (norm-wise:
max |diff| / mean power); digitised samples bit-exact
except rounding-boundary samples, <= 1 LSB in < 1e-4 of samples."""
import numpy as np
import pytest

from conftest import make_input, byte_diff, RFI

pytestmark = pytest.mark.gpu

REL = 1e-5          # north_star tolerance for float spectra / powers
BYTE_FRAC = 1e-4    # north_star allowance for rounding-boundary samples


def check_bytes(a, b, nbit, what):
    worst, frac = byte_diff(a, b, nbit)
    nsamp = a.size * (8 // nbit)
    assert worst <= 1, "%s: code differs by %d" % (what, worst)
    # at small sizes 1e-4 of the samples is a handful: allow 3 boundary samples
    assert frac < BYTE_FRAC or frac * nsamp <= 3, "%s: %.3g of samples differ" % (what, frac)


def run_both(pkg, orc, T, nbit, npol, mode, seed=1, nseg=1, **gen):
    p = pkg.Pipeline(ffts_per_seg=T, nbit=nbit, npol=npol, rfi_mode=mode, keep_stats=1, keep_power=1, do_histo=1)
    o = orc.OracleChain(T, nbit, npol, mode)
    res = []
    for s in range(nseg):
        p0, p1 = make_input(pkg, T, seed=seed, sample0=s * T * 12500, **gen)
        res.append((p.process_segment(p0, p1), o.process_segment(p0, p1)))
    return p, o, res


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("nbit,npol", [(2, 1), (8, 1), (4, 2), (8, 2)])
def test_segment_matches_oracle(pkg, orc, mode, nbit, npol):
    T = 32
    p, o, res = run_both(pkg, orc, T, nbit, npol, mode, seed=10 + mode, **RFI)
    (main, raw), (omain, oraw) = res[0]
    if mode:
        assert np.array_equal(p.get_mask(), o.mask())
        st = p.get_stats()
        assert np.count_nonzero(p.get_mask()) > 0
        for k in ("pow", "kur", "pow_fb", "kur_fb", "weights"):
            assert np.array_equal(st[k], o.get(k), equal_nan=True), k
        # powf is the one operation that differs between glibc (oracle) and libdevice (GPU)
        np.testing.assert_allclose(st["dag"], o.get("dag"), rtol=1e-5, atol=5e-6)
        np.testing.assert_allclose(st["dag_fb"], o.get("dag_fb"), rtol=1e-5, atol=2e-5)
        assert np.array_equal(st["histo"], o.get("histo"))
    det, odet = p.get_detected_power(0, 0), o.power_trimmed("main")
    assert np.abs(det - odet).max() / odet.mean() < REL
    ave, oave = p.get_power_f32(0, 0), o.ave_trimmed("main")
    assert np.abs(ave - oave).max() < 2e-4          # unit-variance quantity: absolute
    check_bytes(main, omain, nbit, "main")
    if mode == 2:
        check_bytes(raw, oraw, nbit, "raw")
        assert np.abs(p.get_power_f32(0, 1) - o.ave_trimmed("raw")).max() < 2e-4
    bp, obp = p.get_bandpass(0, 0), o.get("bp_main").reshape(2, 6251)[:, 2155:2155 + 4096]
    np.testing.assert_allclose(bp, obp, rtol=2e-5)
    p.close()


def test_bandpass_state_across_segments(pkg, orc):
    T = 16
    p, o, res = run_both(pkg, orc, T, 8, 1, 2, seed=4, nseg=4, **RFI)
    for i, ((main, raw), (omain, oraw)) in enumerate(res):
        check_bytes(main, omain, 8, "main seg %d" % i)
        check_bytes(raw, oraw, 8, "raw seg %d" % i)
    p.close()


def test_clean_input(pkg, orc):
    T = 16
    p, o, res = run_both(pkg, orc, T, 2, 1, 2, seed=77)
    (main, raw), (omain, oraw) = res[0]
    assert np.array_equal(p.get_mask(), o.mask())
    check_bytes(main, omain, 2, "main")
    check_bytes(raw, oraw, 2, "raw")
    p.close()


def test_all_dropped_segment(pkg, orc):
    """every frame dropped (bytes 0): all excised, weight 0, bandpass set to 1"""
    T = 8
    z = np.zeros(T * 12500, np.uint8)
    with pkg.Pipeline(ffts_per_seg=T, nbit=2, npol=1, rfi_mode=1, keep_stats=1) as p:
        main, _ = p.process_segment(z, z)
        assert np.all(p.get_mask() == (1 << 25) - 1)
        assert np.all(p.get_stats()["weights"] == 0)
        assert np.all(p.get_bandpass() == 1)
        assert np.all(main == 0b01010101)


def test_partially_dropped_frames(pkg, orc):
    T = 16
    gen = dict(drop_period=7, drop_len=2, drop_pol_skew=3, **RFI)
    p, o, res = run_both(pkg, orc, T, 8, 2, 2, seed=8, **gen)
    (main, raw), (omain, oraw) = res[0]
    assert np.array_equal(p.get_mask(), o.mask())
    assert np.array_equal(p.get_stats()["weights"], o.get("weights"))
    check_bytes(main, omain, 8, "main")
    check_bytes(raw, oraw, 8, "raw")
    p.close()


def test_batch_equals_single(pkg):
    T, n = 16, 3
    ins = [make_input(pkg, T, seed=30, antenna=a, **RFI) for a in range(n)]
    with pkg.Pipeline(ffts_per_seg=T, nbit=8, rfi_mode=2, n_antennas=n) as pb:
        mains, raws = pb.process_batch([i[0] for i in ins], [i[1] for i in ins])
    for a in range(n):
        with pkg.Pipeline(ffts_per_seg=T, nbit=8, rfi_mode=2) as p1:
            m, r = p1.process_segment(*ins[a])
        assert np.array_equal(m, mains[a]) and np.array_equal(r, raws[a])
    assert not np.array_equal(mains[0], mains[1])


def test_vdif_input_equals_planar(pkg):
    T = 16
    nfr = T * 12500 // 5000
    g = pkg.GenParams.default(seed=12, **RFI)
    first = 400
    frames = pkg.gen_vdif_second(g, 0, 1234, first, nfr)
    s0 = (1234 * 25600 + first) * 5000
    p0 = pkg.gen_samples(g, 0, 0, s0, T * 12500)
    p1 = pkg.gen_samples(g, 0, 1, s0, T * 12500)
    # shuffle frame order: depacketising goes by header, not by position
    fr = frames.reshape(-1, 5032)
    perm = np.random.default_rng(0).permutation(fr.shape[0])
    shuffled = np.ascontiguousarray(fr[perm]).reshape(-1)
    with pkg.Pipeline(ffts_per_seg=T, nbit=8, rfi_mode=2) as p:
        m, r = p.process_segment(p0, p1)
        p.reset_bandpass()
        mv, rv = p.process_vdif(shuffled, first)
        assert np.array_equal(m, mv) and np.array_equal(r, rv)
        p.reset_bandpass()
        with pytest.raises(pkg.VfError) as e:
            p.process_vdif(shuffled, first + 1)          # one frame pair falls outside
        assert e.value.code == 23


def test_async_double_buffer_equals_sync(pkg):
    T, nseg = 16, 6
    segs = [make_input(pkg, T, seed=40, sample0=s * T * 12500, **RFI) for s in range(nseg)]
    with pkg.Pipeline(ffts_per_seg=T, nbit=8, rfi_mode=1) as p:
        want = [p.process_segment(*s)[0] for s in segs]
    with pkg.Pipeline(ffts_per_seg=T, nbit=8, rfi_mode=1) as p:
        outs = [np.empty(p.out_bytes, np.uint8) for _ in range(nseg)]
        for s in range(nseg):
            if s >= 2:
                p.wait(s & 1)
            p.submit_async(s & 1, [segs[s][0]], [segs[s][1]], [outs[s]])
        p.wait(0); p.wait(1)
        with pytest.raises(pkg.VfError):
            p.wait(0)
        # per-stage device times of an asynchronous submission (the executable's RT_PROFILE table)
        for slot in (0, 1):
            total, k1, k2 = p.slot_elapsed_ms(slot)
            assert 0 < k1 < total and 0 < k2 < total and k1 + k2 <= total * 1.001, (total, k1, k2)
    for s in range(nseg):
        assert np.array_equal(outs[s], want[s]), s


def test_host_block_equals_per_segment(pkg):
    """vf_submit_block_async (a block of segments x antennas from one host buffer, one launch pair) gives the bytes of
    per-segment vf_process_batch calls, on both slots, with the bandpass carried from block to block"""
    T, nseg, nant, nblk = 16, 3, 2, 3
    rng_in = np.empty((nblk, nseg, nant, 2, T * 12500), np.uint8)
    for b in range(nblk):
        for s in range(nseg):
            for a in range(nant):
                p0, p1 = make_input(pkg, T, seed=61, antenna=a, sample0=(b * nseg + s) * T * 12500, **RFI)
                rng_in[b, s, a, 0], rng_in[b, s, a, 1] = p0, p1
    with pkg.Pipeline(ffts_per_seg=T, nbit=2, rfi_mode=2, n_antennas=nant) as p:
        want = [p.process_batch([rng_in[b, s, a, 0] for a in range(nant)], [rng_in[b, s, a, 1] for a in range(nant)])
                for b in range(nblk) for s in range(nseg)]
    with pkg.Pipeline(ffts_per_seg=T, nbit=2, rfi_mode=2, n_antennas=nant, max_batch_segments=nseg) as p:
        mains = np.zeros((nblk, nseg, nant, p.out_bytes), np.uint8)
        raws = np.zeros_like(mains)
        for b in range(nblk):
            if b >= 2:
                p.wait(b & 1)
            p.submit_block_async(b & 1, nant, nseg, rng_in[b], mains[b], raws[b])
        p.wait(1); p.wait(0)
        with pytest.raises(pkg.VfError):
            p.submit_block_async(0, nant, nseg + 1, np.zeros((nseg + 1, nant, 2, T * 12500), np.uint8),
                                 np.zeros((nseg + 1, nant, p.out_bytes), np.uint8), None)
    for b in range(nblk):
        for s in range(nseg):
            wm, wr = want[b * nseg + s]
            for a in range(nant):
                assert np.array_equal(mains[b, s, a], wm[a]), (b, s, a)
                assert np.array_equal(raws[b, s, a], wr[a]), (b, s, a)


def test_device_resident_equals_host(pkg):
    import torch
    T, nseg, n = 16, 3, 2
    data = np.empty((nseg, n, 2, T * 12500), np.uint8)
    for s in range(nseg):
        for a in range(n):
            data[s, a, 0], data[s, a, 1] = make_input(pkg, T, seed=50, antenna=a, sample0=s * T * 12500, **RFI)
    with pkg.Pipeline(ffts_per_seg=T, nbit=2, rfi_mode=2, n_antennas=n) as p:
        want = [p.process_batch([data[s, a, 0] for a in range(n)], [data[s, a, 1] for a in range(n)]) for s in range(nseg)]
    with pkg.Pipeline(ffts_per_seg=T, nbit=2, rfi_mode=2, n_antennas=n) as p:
        d_in = torch.from_numpy(data).cuda()
        d_main = torch.zeros((nseg, n, p.out_bytes), dtype=torch.uint8, device="cuda")
        d_raw = torch.zeros_like(d_main)
        torch.cuda.synchronize()
        p.process_device(n, nseg, d_in.data_ptr(), d_main.data_ptr(), d_raw.data_ptr())
        p.sync()
        total, k1, k2 = p.last_elapsed_ms()
        assert total > 0 and k1 > 0 and k2 > 0
        m, r = d_main.cpu().numpy(), d_raw.cpu().numpy()
    for s in range(nseg):
        for a in range(n):
            assert np.array_equal(m[s, a], want[s][0][a]) and np.array_equal(r[s, a], want[s][1][a])


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_pipelined_channeliser_stress(pkg, mode):
    """the two warp groups of the pipelined channeliser hand buffers over through named barriers and mbarriers:
    odd item counts, more or fewer items than CTAs, every RFI mode, dropped frames, repeated launches"""
    gen = dict(drop_period=9, drop_len=1, drop_pol_skew=2, **RFI)
    for T, n in ((8, 1), (40, 3), (136, 2), (152, 1)):
        ins = [make_input(pkg, T, seed=90 + T, antenna=a, **gen) for a in range(n)]
        outs = []
        for nt in (640, 1, 0):
            with pkg.Pipeline(testing=True, ffts_per_seg=T, nbit=8, npol=2, rfi_mode=mode, n_antennas=n, k1_threads=nt) as p:
                runs = []
                for _ in range(3):
                    m, r = p.process_batch([i[0] for i in ins], [i[1] for i in ins])
                    runs.append((m, r, [p.get_mask(a) for a in range(n)] if mode else None))
                outs.append(runs)
        # round 1's pipelined kernel (1) has the arithmetic of the monolithic one (640): identical bits; the product
        # kernel (0) has another transform: identical masks, bytes equal up to samples on a rounding boundary
        for a_run, b_run, c_run in zip(*outs):
            for a in range(n):
                assert np.array_equal(a_run[0][a], b_run[0][a])
                check_bytes(a_run[0][a], c_run[0][a], 8, "main T=%d antenna %d" % (T, a))
                if mode == 2:
                    assert np.array_equal(a_run[1][a], b_run[1][a])
                    check_bytes(a_run[1][a], c_run[1][a], 8, "raw T=%d antenna %d" % (T, a))
                if mode:
                    assert np.array_equal(a_run[2][a], b_run[2][a])
                    assert np.array_equal(a_run[2][a], c_run[2][a])
        # and the product kernel is deterministic from launch to launch
        with pkg.Pipeline(testing=True, ffts_per_seg=T, nbit=8, npol=2, rfi_mode=mode, n_antennas=n, k1_threads=0) as p:
            for r_i, c_run in enumerate(outs[2]):
                m, r = p.process_batch([i[0] for i in ins], [i[1] for i in ins])
                for a in range(n):
                    assert np.array_equal(m[a], c_run[0][a]), (T, r_i, a)


def test_batched_launches_equal_per_segment(pkg):
    """vf_process_device covers consecutive segments with one launch pair (max_batch_segments): same bytes, masks
    and detected power of the last segment as one launch pair per segment, for whole and ragged batches"""
    import torch
    T, nseg, n = 256, 5, 2
    data = np.empty((nseg, n, 2, T * 12500), np.uint8)
    for s in range(nseg):
        for a in range(n):
            data[s, a, 0], data[s, a, 1] = make_input(pkg, T, seed=51, antenna=a, sample0=s * T * 12500, **RFI)
    d_in = torch.from_numpy(data).cuda()
    res = []
    for mb in (1, 0, 2):
        with pkg.Pipeline(ffts_per_seg=T, nbit=4, npol=2, rfi_mode=2, n_antennas=n, max_batch_segments=mb, keep_power=1) as p:
            d_main = torch.zeros((nseg, n, p.out_bytes), dtype=torch.uint8, device="cuda")
            d_raw = torch.zeros_like(d_main)
            for _ in range(2):          # twice: the second pass starts from a settled bandpass
                p.process_device(n, nseg, d_in.data_ptr(), d_main.data_ptr(), d_raw.data_ptr())
            p.sync()
            res.append((d_main.cpu().numpy(), d_raw.cpu().numpy(), [p.get_mask(a) for a in range(n)],
                        [p.get_detected_power(a, 0) for a in range(n)], [p.get_power_f32(a, 0) for a in range(n)],
                        [p.get_bandpass(a, 0) for a in range(n)]))
    for r in res[1:]:
        assert np.array_equal(res[0][0], r[0]) and np.array_equal(res[0][1], r[1])
        for a in range(n):
            assert np.array_equal(res[0][2][a], r[2][a])
            assert np.array_equal(res[0][3][a], r[3][a]) and np.array_equal(res[0][4][a], r[4][a])
            assert np.array_equal(res[0][5][a], r[5][a])


def test_k1_thread_variants_agree(pkg):
    """Testing build, A/B: the product channeliser (0: one polarisation per thread group, 6250-point real-input FFT)
    against round 1's kernels with the two-for-one 12500-point FFT -- pipelined (1) and monolithic at three CTA
    sizes.  Masks, weights, statistics and histogram do not depend on the FFT: identical.  The old variants agree
    with each other bit for bit; the new transform rounds differently, so its bytes agree with theirs up to samples
    on a rounding boundary.  T = 1024 makes every CTA of the pipelined kernels draw many items from the counter."""
    for T, n_seg in ((16, 1), (1024, 3)):
        _k1_variants(pkg, T, n_seg)


def _k1_variants(pkg, T, n_seg):
    p0, p1 = make_input(pkg, T, seed=60, **RFI)
    outs = []
    for nt in (0, 1, 320, 512, 640) if T == 16 else (0, 1, 640):
        with pkg.Pipeline(testing=True, ffts_per_seg=T, nbit=8, rfi_mode=2, k1_threads=nt, keep_stats=1, do_histo=1) as p:
            for _ in range(n_seg):          # several launches: the item counter is never reset
                o = p.process_segment(p0, p1)
            outs.append(o + (p.get_mask(), p.get_stats()))
    for o in outs[2:]:
        assert np.array_equal(outs[1][0], o[0]) and np.array_equal(outs[1][1], o[1])
    for o in outs[1:]:
        assert np.array_equal(outs[0][2], o[2])
        for k in ("pow", "kur", "dag", "pow_fb", "kur_fb", "dag_fb", "weights", "histo"):
            assert np.array_equal(outs[0][3][k], o[3][k], equal_nan=True), k
    check_bytes(outs[0][0], outs[1][0], 8, "main, new against old transform")
    check_bytes(outs[0][1], outs[1][1], 8, "raw, new against old transform")


def test_coadd_single_process(pkg, orc):
    """co-add of 3 antennas on one GPU = digitised (sum of the oracle's tiles / sqrt 3)"""
    T, n = 16, 3
    ins = [make_input(pkg, T, seed=70, antenna=a, **RFI) for a in range(n)]
    with pkg.Pipeline(ffts_per_seg=T, nbit=8, rfi_mode=1, n_antennas=n, keep_power=1) as p:
        p.process_batch([i[0] for i in ins], [i[1] for i in ins])
        p.coadd_init()
        fb, sm = p.coadd_segment(0, n)
    tiles = []
    for a in range(n):
        o = orc.OracleChain(T, 8, 1, 1)
        o.process_segment(*ins[a])
        tiles.append(o.ave_trimmed("main"))
    osum = tiles[0] + tiles[1] + tiles[2]
    assert np.abs(sm - osum).max() < 5e-4
    full = np.zeros((1, T // 8, 6251), np.float32)
    full[:, :, 2155:2155 + 4096] = osum * np.float32(1 / np.sqrt(3.0))
    want = np.empty(fb.size, np.uint8)
    orc.liba().orc_digitise(full.ctypes.data, want.ctypes.data, T // 8, 1, 8)
    check_bytes(fb, want, 8, "coadd")


@pytest.mark.parametrize("T,nbit,npol", [(8, 2, 1), (24, 4, 2), (72, 2, 2), (136, 8, 1)])
def test_ragged_segment_lengths(pkg, orc, T, nbit, npol):
    """segment lengths that are not a multiple of the normaliser's 64-step chunk"""
    p, o, res = run_both(pkg, orc, T, nbit, npol, 2, seed=T, nseg=2, **RFI)
    for (main, raw), (omain, oraw) in res:
        check_bytes(main, omain, nbit, "main")
        check_bytes(raw, oraw, nbit, "raw")
    assert np.array_equal(p.get_mask(), o.mask())
    p.close()


def test_coadd_batch_of_segments(pkg, orc):
    """tiles of 3 consecutive segments kept and co-added in one call"""
    T, n, nseg = 16, 2, 3
    with pkg.Pipeline(ffts_per_seg=T, nbit=8, rfi_mode=1, n_antennas=n, keep_power=1, power_segments=4) as p:
        p.coadd_init()
        oracles = [orc.OracleChain(T, 8, 1, 1) for _ in range(n)]
        want = []
        for s in range(nseg + 2):                  # two extra segments first: the tile ring wraps
            ins = [make_input(pkg, T, seed=80, antenna=a, sample0=s * T * 12500, **RFI) for a in range(n)]
            p.process_batch([i[0] for i in ins], [i[1] for i in ins])
            tiles = []
            for a in range(n):
                oracles[a].process_segment(*ins[a])
                tiles.append(oracles[a].ave_trimmed("main"))
            want.append(tiles[0] + tiles[1])
        fb, sm = p.coadd_batch(0, n, nseg)
        for i in range(nseg):
            assert np.abs(sm[i] - want[2 + i]).max() < 5e-4, i
        assert fb.shape == (nseg, T // 8 * 4096)
        with pytest.raises(pkg.VfError):
            p.coadd_batch(0, n, 5)


def test_coadd_batch_after_batched_launches(pkg):
    """the ring of kept f32 tiles is filled from inside a batched launch (segment s of the launch -> tile
    (counter + s) % ring): the co-add of the last segments equals the one after per-segment launches, also when
    the ring wraps inside a launch"""
    import torch
    T, n, nseg = 64, 2, 6
    data = np.empty((nseg, n, 2, T * 12500), np.uint8)
    for s in range(nseg):
        for a in range(n):
            data[s, a, 0], data[s, a, 1] = make_input(pkg, T, seed=81, antenna=a, sample0=s * T * 12500, **RFI)
    d_in = torch.from_numpy(data).cuda()
    res = []
    for mb in (1, 0, 4):
        with pkg.Pipeline(ffts_per_seg=T, nbit=8, rfi_mode=2, n_antennas=n, keep_power=1, power_segments=4,
                          max_batch_segments=mb) as p:
            p.coadd_init()
            d_main = torch.zeros((nseg, n, p.out_bytes), dtype=torch.uint8, device="cuda")
            d_raw = torch.zeros_like(d_main)
            p.process_device(n, nseg, d_in.data_ptr(), d_main.data_ptr(), d_raw.data_ptr())
            p.sync()
            fb, sm = p.coadd_batch(0, n, 4)
            res.append((fb.copy(), sm.copy(), p.get_power_f32(1, 0)))
    for r in res[1:]:
        assert np.array_equal(res[0][0], r[0]) and np.array_equal(res[0][1], r[1]) and np.array_equal(res[0][2], r[2])


def test_vdif_missing_and_invalid_frames_become_dropped_samples(pkg, orc):
    """frames the writer never delivered (absent, or marked invalid) are zeros = dropped data"""
    T = 16
    nfr = T * 12500 // 5000
    g = pkg.GenParams.default(seed=15, **RFI)
    frames = pkg.gen_vdif_second(g, 0, 99, 0, nfr).reshape(-1, 5032).copy()
    p0 = pkg.gen_samples(g, 0, 0, 99 * 25600 * 5000, T * 12500).copy()
    p1 = pkg.gen_samples(g, 0, 1, 99 * 25600 * 5000, T * 12500).copy()
    # frame pair 3 is absent, frame 7 of thread 1 carries the VDIF invalid bit
    keep = np.ones(frames.shape[0], bool)
    keep[6] = keep[7] = False
    frames[15, 3] |= 0x80                      # word0 bit 31 (little endian: top bit of byte 3)
    p0[3 * 5000:4 * 5000] = 0; p1[3 * 5000:4 * 5000] = 0
    p1[7 * 5000:8 * 5000] = 0
    with pkg.Pipeline(ffts_per_seg=T, nbit=8, rfi_mode=2, keep_stats=1) as p:
        want = p.process_segment(p0, p1)
        wmask = p.get_mask()
        p.reset_bandpass()
        got = p.process_vdif(np.ascontiguousarray(frames[keep]).reshape(-1), 0)
        assert np.array_equal(p.get_mask(), wmask)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    assert np.count_nonzero(wmask) > 0


def test_batch_statistics_are_per_antenna(pkg, orc):
    T, n = 16, 2
    ins = [make_input(pkg, T, seed=31, antenna=a, **RFI) for a in range(n)]
    with pkg.Pipeline(ffts_per_seg=T, nbit=2, rfi_mode=1, n_antennas=n, keep_stats=1, do_histo=1, keep_power=1) as p:
        p.process_batch([i[0] for i in ins], [i[1] for i in ins])
        for a in range(n):
            o = orc.OracleChain(T, 2, 1, 1)
            o.process_segment(*ins[a])
            st = p.get_stats(a)
            assert np.array_equal(p.get_mask(a), o.mask())
            for k in ("pow", "kur", "weights", "histo"):
                assert np.array_equal(st[k], o.get(k), equal_nan=True), (a, k)
            assert np.abs(p.get_power_f32(a, 0) - o.ave_trimmed("main")).max() < 2e-4


def test_argument_errors(pkg):
    with pkg.Pipeline(ffts_per_seg=8, rfi_mode=2, n_antennas=1) as p:
        z = np.zeros(8 * 12500, np.uint8)
        with pytest.raises(pkg.VfError) as e:
            p.process_segment(z[:-1], z[:-1])                 # wrong segment length
        assert e.value.code == 1
        with pytest.raises(pkg.VfError) as e:
            p.process_batch([z, z], [z, z])                   # more antennas than the handle has
        assert e.value.code == 1
        with pytest.raises(pkg.VfError) as e:
            p.get_power_f32()                                 # keep_power not set
        assert e.value.code == 22
        with pytest.raises(pkg.VfError) as e:
            p.set_frb_injection(0)                            # inject_frb not set
        assert e.value.code == 22
    with pytest.raises(pkg.VfError) as e:
        pkg.Pipeline(ffts_per_seg=8, gpu_id=99)
    assert e.value.code == 25


def test_packed_division_is_correctly_rounded(pkg):
    """the normaliser's packed division against CUDA's div.rn.f32, bit for bit"""
    rng = np.random.default_rng(3)
    n = 1 << 22
    p = np.exp(rng.uniform(np.log(1e-3), np.log(1e9), n)).astype(np.float32)
    b = np.exp(rng.uniform(np.log(1e-2), np.log(1e8), n)).astype(np.float32)
    p[:64] = 0.0
    p[64:128] = np.inf
    b[128:192] = 1e-35            # outside the fast range: plain division
    b[192:256] = 1e32
    p[256:512] = b[256:512] * np.float32(11.0)
    p[512:768] = np.nextafter(b[512:768], np.float32(np.inf))
    with pkg.Pipeline(testing=True, ffts_per_seg=8) as pl:
        qp, qr = pl.debug_division(p, b)
    ok = (qp.view(np.uint32) == qr.view(np.uint32)) | (np.isnan(qp) & np.isnan(qr))
    ok[64:128] = True             # +inf dividends: the quotient is never used (weight 0)
    assert ok.all(), (int((~ok).sum()), p[~ok][:4], b[~ok][:4], qp[~ok][:4], qr[~ok][:4])
    with np.errstate(all="ignore"):
        want = (p.astype(np.float64) / b.astype(np.float64)).astype(np.float32)
    sel = np.isfinite(want) & (want > 1e-30)
    assert np.array_equal(qr[sel], want[sel])


def test_single_antenna_entry_points_on_a_multi_antenna_handle(pkg):
    """vf_process_segment / vf_process_vdif take the antenna: its own bandpass, statistics and kept tile"""
    T, n = 16, 3
    nfr = T * 12500 // 5000
    g = pkg.GenParams.default(seed=33, **RFI)
    ins = [(pkg.gen_samples(g, a, 0, 0, T * 12500), pkg.gen_samples(g, a, 1, 0, T * 12500)) for a in range(n)]
    ins2 = [(pkg.gen_samples(g, a, 0, T * 12500, T * 12500), pkg.gen_samples(g, a, 1, T * 12500, T * 12500)) for a in range(n)]
    with pkg.Pipeline(ffts_per_seg=T, nbit=8, rfi_mode=2, n_antennas=n, keep_power=1, keep_stats=1) as pb:
        pb.process_batch([i[0] for i in ins], [i[1] for i in ins])
        want = pb.process_batch([i[0] for i in ins2], [i[1] for i in ins2])       # second segment: carried bandpass
        wmask = [pb.get_mask(a) for a in range(n)]
        wave = [pb.get_power_f32(a, 0) for a in range(n)]
    with pkg.Pipeline(ffts_per_seg=T, nbit=8, rfi_mode=2, n_antennas=n, keep_power=1, keep_stats=1) as p:
        for a in (2, 0, 1):                                   # any order, one antenna at a time
            p.process_segment(*ins[a], antenna=a)
        got = {}
        for a in (1, 2, 0):
            if a == 1:      # the raw-frame form for one of them
                frames = pkg.gen_vdif_second(g, a, 0, nfr, nfr)      # frames nfr..2nfr-1 of second 0 = samples T*12500..
                got[a] = p.process_vdif(frames, nfr, antenna=a)
            else:
                got[a] = p.process_segment(*ins2[a], antenna=a)
        for a in range(n):
            assert np.array_equal(got[a][0], want[0][a]) and np.array_equal(got[a][1], want[1][a]), a
            assert np.array_equal(p.get_mask(a), wmask[a]), a
            assert np.array_equal(p.get_power_f32(a, 0), wave[a]), a
        with pytest.raises(pkg.VfError) as e:
            p.process_segment(*ins[0], antenna=n)
        assert e.value.code == 1


def test_statistics_follow_the_slot_of_the_last_segment(pkg, orc):
    """statistics and histogram buffers are per slot; after segments that alternate between the two
    slots (vf_process_device, vf_submit_async x2) vf_get_stats returns those of the last segment, unmixed"""
    import torch
    T, nseg = 64, 3
    segs = [make_input(pkg, T, seed=21, sample0=s * T * 12500, **RFI) for s in range(nseg)]
    o = orc.OracleChain(T, 8, 1, 2)
    for s in segs:
        o.process_segment(*s)
    def check(p):
        st = p.get_stats()
        for k in ("pow", "kur", "pow_fb", "kur_fb", "weights", "histo"):
            assert np.array_equal(st[k], o.get(k), equal_nan=True), k
        assert np.array_equal(p.get_mask(), o.mask())
    data = np.stack([np.stack(s) for s in segs])[:, None]          # [seg][1][2][n]
    with pkg.Pipeline(ffts_per_seg=T, nbit=8, rfi_mode=2, keep_stats=1, do_histo=1) as p:
        d_in = torch.from_numpy(np.ascontiguousarray(data)).cuda()
        d_main = torch.zeros((nseg, 1, p.out_bytes), dtype=torch.uint8, device="cuda")
        d_raw = torch.zeros_like(d_main)
        p.process_device(1, nseg, d_in.data_ptr(), d_main.data_ptr(), d_raw.data_ptr())
        check(p)
    with pkg.Pipeline(ffts_per_seg=T, nbit=8, rfi_mode=2, keep_stats=1, do_histo=1) as p:
        outs = [(np.empty(p.out_bytes, np.uint8), np.empty(p.out_bytes, np.uint8)) for _ in range(nseg)]
        for s in range(nseg):
            if s >= 2:
                p.wait(s & 1)
            p.submit_async(s & 1, [segs[s][0]], [segs[s][1]], [outs[s][0]], [outs[s][1]])
        p.wait(1); p.wait(0)
        check(p)


def _row_kept(w, min_weight=0.2):
    """tscrunch_weights' row rule (src/pb_kernels.cu:616-623) from the step weights [T]"""
    T = w.size
    kept = np.zeros(T // 8, bool)
    for t8 in range(T // 8):
        ws = np.float32(0)
        for j in range(8):
            x = w[t8 * 8 + j]
            if x != 0 and float(x) >= min_weight:
                ws = np.float32(ws + x)
        kept[t8] = float(np.float32(ws / np.float32(8))) >= min_weight
    return kept


def test_coadd_counts_the_antennas_that_kept_the_row(pkg, orc):
    """values and counts are summed, the sum is divided by sqrt (count): an antenna whose scrunched row
    was zeroed (dropped data) does not dilute the others"""
    T, n = 32, 3
    ins = [list(make_input(pkg, T, seed=71, antenna=a, **RFI)) for a in range(n)]
    ins[1][0] = ins[1][0].copy(); ins[1][1] = ins[1][1].copy()
    ins[1][0][:16 * 12500] = 0; ins[1][1][:16 * 12500] = 0          # antenna 1: time steps 0..15 dropped -> rows 0, 1 zeroed
    ins[2][0] = ins[2][0].copy(); ins[2][1] = ins[2][1].copy()
    ins[2][0][8 * 12500:16 * 12500] = 0; ins[2][1][8 * 12500:16 * 12500] = 0      # antenna 2: row 1
    with pkg.Pipeline(ffts_per_seg=T, nbit=8, rfi_mode=1, n_antennas=n, keep_power=1) as p:
        p.process_batch([i[0] for i in ins], [i[1] for i in ins])
        p.coadd_init()
        fb, sm = p.coadd_segment(0, n)
    tiles, kept = [], []
    for a in range(n):
        o = orc.OracleChain(T, 8, 1, 1)
        o.process_segment(*ins[a])
        tiles.append(o.ave_trimmed("main"))
        kept.append(_row_kept(o.get("weights")[:T]))
    cnt = np.sum(kept, axis=0)
    assert list(cnt) == [2, 1, 3, 3]
    osum = tiles[0] + tiles[1] + tiles[2]
    assert np.all(tiles[1][0, :2] == 0) and np.all(tiles[2][0, 1] == 0)
    assert np.abs(sm - osum).max() < 5e-4
    full = np.zeros((1, T // 8, 6251), np.float32)
    full[:, :, 2155:2155 + 4096] = osum / np.sqrt(cnt.astype(np.float32))[None, :, None]
    want = np.empty(fb.size, np.uint8)
    orc.liba().orc_digitise(full.ctypes.data, want.ctypes.data, T // 8, 1, 8)
    check_bytes(fb, want, 8, "coadd with counts")


def test_coadd_sums_only_the_antennas_of_the_last_launch(pkg, orc):
    """a handle for 3 antennas that last processed 2 co-adds those 2 (no stale tile)"""
    T, n = 16, 3
    ins = [make_input(pkg, T, seed=72, antenna=a, **RFI) for a in range(n)]
    with pkg.Pipeline(ffts_per_seg=T, nbit=8, rfi_mode=0, n_antennas=n, keep_power=1) as p:
        p.coadd_init()
        p.process_batch([i[0] for i in ins], [i[1] for i in ins])
        p.reset_bandpass()
        p.process_batch([i[0] for i in ins[:2]], [i[1] for i in ins[:2]])
        fb, sm = p.coadd_segment(0, 2)
        t0, t1 = p.get_power_f32(0, 0), p.get_power_f32(1, 0)
    assert np.array_equal(sm, t0 + t1)


def test_thresholds_are_configurable(pkg):
    """dag_thresh / min_weight (DAG_THRESH, MIN_WEIGHT of src/process_baseband.h:42,45) are carried by vf_config"""
    T = 128
    p0, p1 = make_input(pkg, T, seed=14, **RFI)
    with pkg.Pipeline(ffts_per_seg=T, nbit=8, rfi_mode=1, keep_stats=1, keep_power=1) as p:
        p.process_segment(p0, p1)
        dag, m_def = p.get_stats()["dag"][:T * 25].reshape(T, 25), p.get_mask()
        zero_def = np.all(p.get_power_f32(0, 0)[0] == 0, axis=1)
    outcomes = set()
    for thr, mw in ((2.0, 0.5), (1.5, 0.8), (1.0, 0.9)):
        with pkg.Pipeline(ffts_per_seg=T, nbit=8, rfi_mode=1, keep_stats=1, dag_thresh=thr, min_weight=mw, keep_power=1) as p:
            p.process_segment(p0, p1)
            m, w = p.get_mask(), p.get_stats()["weights"][:T]
            ave = p.get_power_f32(0, 0)
        want = ((dag > np.float32(thr)).astype(np.uint32) << np.arange(25, dtype=np.uint32)).sum(axis=1).astype(np.uint32)
        assert np.array_equal(m, want) and not np.array_equal(m, m_def)
        kept = _row_kept(w, mw)
        assert np.array_equal(np.all(ave[0] == 0, axis=1), ~kept), (thr, mw)
        outcomes.add((bool(kept.any()), bool((~kept).any())))
    assert not zero_def.all()
    assert any(o[1] for o in outcomes)                       # some setting zeroes rows the defaults keep
    with pytest.raises(pkg.VfError):
        pkg.Pipeline(ffts_per_seg=T, dag_thresh=-1.0)


def test_vdif_block_places_frames_across_the_whole_block(pkg):
    """ADVICE r01 / src/process_baseband.cu:1015-1035: every frame of the block is placed by (thread, frame number)
    wherever it sits in the buffer; a frame of another second is skipped and counted, an absent frame leaves zeros
    (dropped samples); none of that is fatal, and the block goes through ONE launch pair"""
    T, nseg = 16, 5
    per_seg = T * 12500 // 5000                          # frames per pol per segment
    nfr = per_seg * nseg
    g = pkg.GenParams.default(seed=17, **RFI)
    sec, first = 4321, 800
    frames = pkg.gen_vdif_second(g, 0, sec, first, nfr).reshape(-1, 5032).copy()
    s0 = (sec * 25600 + first) * 5000
    p0 = pkg.gen_samples(g, 0, 0, s0, nseg * T * 12500).copy()
    p1 = pkg.gen_samples(g, 0, 1, s0, nseg * T * 12500).copy()
    # frame 50 of thread 0 is absent, frame 120 of thread 1 carries another second, frame 130 (both threads) the invalid bit
    keep = np.ones(frames.shape[0], bool)
    keep[2 * 50] = False
    frames[2 * 120 + 1, :4] = np.frombuffer(np.uint32(sec + 1).tobytes(), np.uint8)
    frames[2 * 130, 3] |= 0x80; frames[2 * 130 + 1, 3] |= 0x80
    p0[50 * 5000:51 * 5000] = 0
    p1[120 * 5000:121 * 5000] = 0
    p0[130 * 5000:131 * 5000] = 0; p1[130 * 5000:131 * 5000] = 0
    fr = frames[keep]
    perm = np.random.default_rng(1).permutation(fr.shape[0])        # any order across the WHOLE block
    shuffled = np.ascontiguousarray(fr[perm]).reshape(-1)
    with pkg.Pipeline(ffts_per_seg=T, nbit=8, rfi_mode=2) as p:
        want = [p.process_segment(p0[s * T * 12500:(s + 1) * T * 12500], p1[s * T * 12500:(s + 1) * T * 12500]) for s in range(nseg)]
        wmask = p.get_mask()
        p.reset_bandpass()
        main, raw, rep, warned = p.process_vdif_block(shuffled, first, sec, nseg)
        assert warned and rep == (0, 1, 2 * nfr - 4, 2, 2 * nfr)
        assert np.array_equal(p.get_mask(), wmask)
        for s in range(nseg):
            assert np.array_equal(main[s], want[s][0]) and np.array_equal(raw[s], want[s][1]), s
        # a clean block: no warning
        p.reset_bandpass()
        clean = pkg.gen_vdif_second(g, 0, sec, first, nfr)
        _, _, rep, warned = p.process_vdif_block(clean, first, sec, nseg, slot=1)
        assert not warned and rep == (0, 0, 2 * nfr, 0, 2 * nfr)
    with pkg.Pipeline(ffts_per_seg=T, nbit=8, rfi_mode=2, keep_stats=1) as p:
        with pytest.raises(pkg.VfError) as e:                      # statistics dumps are per segment
            p.process_vdif_block(clean, first, sec, nseg)
        assert e.value.code == 22
