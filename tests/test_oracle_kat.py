"""CPU suite: hand-written known-answer tests of the oracle's restatement of the
reference-authored stages."""
import numpy as np

T0 = 1.7e9


def _ev(x, y, t, p):
    return (np.array(x, np.uint16), np.array(y, np.uint16), np.array(t, np.float64),
            np.array(p, np.uint8))


def test_refractory_filter_and_polarity_rearm(ora):
    """event_detector.cc:149-166: a same-polarity event within 10 ms is not ACCEPTED unless
    the opposite polarity fired in between; sae_latest always advances."""
    s = ora.Sae(32, 32)
    s.update(*_ev([5, 5], [6, 6], [T0 + 0.100, T0 + 0.105], [1, 1]))
    sae0, sae1, lat0, lat1 = s.planes()
    assert sae1[6, 5] == T0 + 0.100 and lat1[6, 5] == T0 + 0.105 and sae0[6, 5] == 0
    # 20 ms later: accepted again
    s.update(*_ev([5], [6], [T0 + 0.125], [1]))
    assert s.planes()[1][6, 5] == T0 + 0.125
    # opposite polarity re-arms: +, -, + within 2 ms are all accepted
    s2 = ora.Sae(32, 32)
    s2.update(*_ev([1, 1, 1], [1, 1, 1], [T0 + 0.200, T0 + 0.201, T0 + 0.202], [1, 0, 1]))
    p = s2.planes()
    assert p[1][1, 1] == T0 + 0.202 and p[0][1, 1] == T0 + 0.201


def test_time_surface_values(ora):
    """event_detector.cc:230-267: 255*(+-exp(-dt/20ms)+1)/2 -> u8, empty pixel 128."""
    s = ora.Sae(16, 16)
    s.update(*_ev([2, 3, 4], [2, 2, 2], [T0, T0, T0 - 0.020], [1, 0, 1]))
    ts = s.time_surface(T0)
    assert ts[0, 0] == 128
    assert ts[2, 2] == 255 and ts[2, 3] == 0
    assert ts[2, 4] == int(np.rint(127.5 * np.exp(-1.0) + 127.5))
    ts_np = s.time_surface(T0, ignore_polarity=1)
    assert ts_np[0, 0] == 0 and ts_np[2, 2] == 255 and ts_np[2, 3] == 255


def _paint(s, pts, t):
    xs, ys = zip(*pts)
    s.update(*_ev(xs, ys, [t] * len(xs), [1] * len(xs)))


def test_arc_star_corner_vs_edge(ora):
    """event_detector.cc:308-544 on synthetic SAEs: an L-shaped front is a corner, a straight
    edge is not, an isolated event is not."""
    W = H = 64
    cx = cy = 32
    # straight vertical edge sweeping right: newest column at x = 32
    s = ora.Sae(W, H)
    for k, x in enumerate(range(20, 33)):
        _paint(s, [(x, y) for y in range(10, 54)], T0 + 0.02 * k)
    assert not s.is_corner(T0 + 0.02 * 12, cx, cy, 1)
    # quarter-plane (corner of a square) growing diagonally: fresh pixels form an L
    s = ora.Sae(W, H)
    for k, d in enumerate(range(20, 33)):
        _paint(s, [(d, y) for y in range(10, d + 1)] + [(x, d) for x in range(10, d + 1)], T0 + 0.02 * k)
    assert s.is_corner(T0 + 0.02 * 12, cx, cy, 1)
    # isolated event on an empty surface: every ring element is 0 -> not a corner
    s = ora.Sae(W, H)
    _paint(s, [(cx, cy)], T0)
    assert not s.is_corner(T0, cx, cy, 1)
    # too close to the border (MIN_DIST + 1)
    assert not s.is_corner(T0, 5, 5, 1)
    # an event superseded by the opposite polarity at its pixel is rejected first
    s.update(*_ev([cx], [cy], [T0 + 0.001], [0]))
    assert not s.is_corner(T0, cx, cy, 1)


def test_disc_r10_has_317_pixels(ora):
    hw = ora.disc_half_widths(10)
    assert int((2 * hw + 1).sum() * 2 - (2 * hw[0] + 1)) == 317
    m = np.zeros((41, 41), np.uint8)
    ora.fill_disc(m, 20, 20, 10)
    assert int((m == 255).sum()) == 317


def test_set_mask_orders_by_track_count(ora):
    """feature_tracker.cpp:123-151: older tracks win the min-distance conflict."""
    pts = np.array([[50.0, 50.0], [55.0, 50.0], [100.0, 100.0]], np.float32)
    ids = np.array([1, 2, 3], np.int32)
    cnt = np.array([2, 7, 1], np.int32)
    kp, ki, kc, mask = ora.set_mask(346, 260, 10, pts, ids, cnt)
    assert ki.tolist() == [2, 3] and kc.tolist() == [7, 1]
    assert mask[50, 55] == 255 and mask[50, 66] == 0 and mask[100, 100] == 255


def test_lk_recovers_subpixel_translation(ora):
    rng = np.random.default_rng(1)
    yy, xx = np.mgrid[0:120, 0:160].astype(np.float64)
    def img(dx, dy):
        v = 128 + 60 * np.sin((xx - dx) / 7.0) * np.cos((yy - dy) / 9.0) + 40 * np.sin((xx - dx + yy - dy) / 13.0)
        return np.clip(np.rint(v), 0, 255).astype(np.uint8)
    a, b = img(0, 0), img(1.75, -0.6)
    pts = np.stack([rng.uniform(30, 130, 20), rng.uniform(30, 90, 20)], 1).astype(np.float32)
    out, st = ora.calc_optical_flow_pyr_lk(a, b, pts, None, max_level=3)
    assert st.all()
    d = out - pts
    assert np.abs(d[:, 0] - 1.75).max() < 0.08 and np.abs(d[:, 1] + 0.6).max() < 0.08


def test_lift_projective_inverts_distortion(ora):
    from esvio_b200 import synth
    cam = synth.CAM_DAVIS346[0]
    # distort a normalised point with the radial-tangential model, project, lift back
    x, y = 0.21, -0.13
    r2 = x * x + y * y
    rad = cam["k1"] * r2 + cam["k2"] * r2 * r2
    xd = x + x * rad + 2 * cam["p1"] * x * y + cam["p2"] * (r2 + 2 * x * x)
    yd = y + y * rad + 2 * cam["p2"] * x * y + cam["p1"] * (r2 + 2 * y * y)
    u, v = cam["fx"] * xd + cam["cx"], cam["fy"] * yd + cam["cy"]
    lx, ly = ora.lift_projective(cam, u, v)
    assert abs(lx - x) < 1e-6 and abs(ly - y) < 1e-6


def test_oracle_tracker_first_publish_and_ids(ora):
    from esvio_b200 import synth
    cfg = synth.default_config(346, 260)
    t = ora.OracleTracker(cfg, use_cv2=False)
    s = synth.StereoEventStream(346, 260, 1.0e6)
    L, R, tr = s.stereo_window(0)
    o = t.track(tr, L, R, True)
    assert len(o["id"]) > 0 and (o["track_cnt"] == 1).all() and (o["vx"] == 0).all()
    assert o["id"].tolist() == list(range(len(o["id"])))      # n_id++ in selection order
    assert set(o["id_right"]) <= set(o["id"])
    L, R, tr = s.stereo_window(1)
    o2 = t.track(tr, L, R, False)
    assert (o2["track_cnt"] == 2).all() and set(o2["id"]) <= set(o["id"])


# ---- motion-compensated SAE (event_detector.cc:102-147,547-591) ----
MOTION = dict(state_v=(1.0, 0.5, 0.2), v_pre=(0.9, 0.45, 0.25), accel=(4.0, 3.0, 2.0),
              omega=(0.5, -0.3, 0.8), t1=0.0333, K=(226.38, 226.15, 173.0, 130.0))


def test_matrix_exponential_is_a_rotation(ora):
    """Matrix3f::exp() of a skew matrix = Rodrigues' rotation, through all three Pade branches
    (L1 norm < 0.426, < 1.88, beyond with squarings)."""
    rng = np.random.default_rng(5)
    for scale in (0.05, 0.3, 1.0, 3.0):
        for _ in range(20):
            v = rng.normal(size=3) * scale
            K = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
            th = np.linalg.norm(v)
            ref = np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * (K @ K)
            R = ora.mat3_exp_f(K)
            assert np.abs(R - ref).max() < 2e-6 * max(1.0, scale)
            assert abs(np.linalg.det(R.astype(np.float64)) - 1.0) < 1e-5


def test_motion_correction_kat(ora):
    W, H = 346, 260
    m = dict(MOTION)
    assert ora.motion_active(m)
    assert not ora.motion_active(dict(m, accel=(3.0, 3.0, 2.0)))      # |a| = 4.69 <= 5
    # border pixels (<= 6 from the left/top, > W-6 / H-6) are never moved (ed.cc:552-553)
    for x, y in ((6, 100), (100, 6), (341, 100), (100, 255)):
        assert ora.motion_correct(m, W, H, x, y, 0.02) == (x, y)
    # no rotation, no translation: the warp is K * I * K^-1 in float -- it returns the pixel or,
    # where the float product lands a hair below the integer, its floor neighbour
    still = dict(m, omega=(0, 0, 0), state_v=(0, 0, 0), v_pre=(0, 0, 0))
    for x, y in ((50, 60), (173, 130), (300, 200)):
        ox, oy = ora.motion_correct(still, W, H, x, y, 0.02)
        assert x - 1 <= ox <= x and y - 1 <= oy <= y
    # pure roll about the optical axis by w*dt: pixels rotate about (cx, cy) the other way
    roll = dict(still, omega=(0, 0, 2.0))
    ox, oy = ora.motion_correct(roll, W, H, 273, 130, 0.05)   # 100 px right of the centre
    ang = -2.0 * 0.05
    assert abs(ox - (173 + 100 * np.cos(ang))) <= 1.5 and abs(oy - (130 + 100 * np.sin(ang))) <= 1.5
    # a warp that leaves the sensor keeps the raw pixel (ed.cc:574-582)
    far = dict(still, omega=(0, 40.0, 0))
    assert ora.motion_correct(far, W, H, 300, 130, 0.03) == (300, 130)


def test_sae_update_mc_gates(ora):
    """The warp applies only when dt_window > 0, (t - t0)/dt_window < 1 and |accel| > 5
    (feature_tracker.cpp:628, event_detector.cc:125); otherwise it is the plain update."""
    W, H = 346, 260
    rng = np.random.default_rng(9)
    n = 4000
    x = rng.integers(0, W, n).astype(np.uint16)
    y = rng.integers(0, H, n).astype(np.uint16)
    t = np.sort(rng.uniform(0.0, 0.0333, n)) + 1000.0
    p = rng.integers(0, 2, n).astype(np.uint8)
    plain = ora.Sae(W, H)
    plain.update(x, y, t, p)
    for m in (dict(MOTION, t1=1000.0 + 0.0333, accel=(1.0, 1.0, 1.0)),   # below the threshold
              dict(MOTION, t1=999.0)):                                   # header stamp before t0
        s = ora.Sae(W, H)
        s.update_mc(x, y, t, p, m, t[0])
        assert all(np.array_equal(a, b) for a, b in zip(s.planes(), plain.planes()))
    s = ora.Sae(W, H)
    s.update_mc(x, y, t, p, dict(MOTION, t1=1000.0 + 0.0333), t[0])
    assert not all(np.array_equal(a, b) for a, b in zip(s.planes(), plain.planes()))
    # same number of events landed, only somewhere else
    assert (s.planes()[2] > 0).sum() + (s.planes()[3] > 0).sum() > 0


# ---- cv::goodFeaturesToTrack restatement (frame path) ----
def _square_image(size=48, squares=((12, 12, 12, 200),)):
    img = np.zeros((size, size), np.uint8)
    for x0, y0, s, v in squares:
        img[y0:y0 + s, x0:x0 + s] = v
    return img


def test_good_features_find_the_corners_of_a_square(ora):
    """A bright square on black has exactly four minimum-eigenvalue maxima, one per corner;
    edges (one large eigenvalue only) and flat areas give nothing."""
    img = _square_image()
    pts = ora.good_features_to_track(img, 10, 0.01, 5.0)
    assert len(pts) == 4
    want = np.array([[11.5, 11.5], [23.5, 11.5], [11.5, 23.5], [23.5, 23.5]])
    d = np.abs(pts[:, None, :] - want[None, :, :]).max(2)          # Chebyshev distance
    assert (d.min(0) <= 1.0).all() and (d.min(1) <= 1.0).all()
    eig = ora.corner_min_eigen_val(img)
    assert eig[18, 12] < 1e-3 * eig.max() and eig[5, 5] == 0.0      # edge midpoint, flat area
    assert len(ora.good_features_to_track(np.full((40, 40), 9, np.uint8), 10, 0.01, 5.0)) == 0


def test_good_features_mask_quality_cap_and_spacing(ora):
    img = _square_image(64, ((8, 8, 12, 200), (40, 40, 12, 12)))     # strong and weak square
    strong = ora.good_features_to_track(img, 0, 0.01, 1.0)
    # quality 0.01: the weak square's corners ((12/200)^2 of the strong ones) stay below 1 % of max
    assert len(strong) == 4 and strong.max() < 30
    both = ora.good_features_to_track(img, 0, 1e-5, 1.0)
    assert len(both) == 8 and (both[:4].max() < 30) and (both[4:].min() > 30)   # best first
    # mask: only pixels with a non-zero mask may become corners, and the maximum is taken
    # under the mask, so the weak square alone passes the 1 % test again
    mask = np.zeros_like(img)
    mask[32:, 32:] = 255
    weak = ora.good_features_to_track(img, 0, 0.01, 1.0, mask)
    assert len(weak) == 4 and weak.min() > 30
    # maxCorners cuts the sorted list; a large minimum distance keeps one corner per square
    assert np.array_equal(ora.good_features_to_track(img, 2, 1e-5, 1.0), both[:2])
    far = ora.good_features_to_track(img, 0, 1e-5, 20.0)
    assert len(far) == 2 and np.array_equal(far[0], both[0])
    # min distance < 1: no spacing rule at all
    assert np.array_equal(ora.good_features_to_track(img, 0, 1e-5, 0.5), both)


def test_time_window_shard_rule_reproduces_the_sequential_sae(ora):
    """SURVEY.md 8e row 3 (time-window shard), stated and checked on the oracle: the acceptance
    test of createSAE (event_detector.cc:157) reads only `sae_latest`, and `sae_latest` after a
    window is "last event time per pixel and polarity".  So window k can be processed on its
    own rank from a carry-in of `sae_latest` alone (an exclusive "last non-empty" scan of the
    ranks' local last-event planes), and `sae` follows by an inclusive scan of the per-window
    accepted times.  The planes and the time surface after every window must equal the
    sequential ones."""
    from esvio_b200 import synth
    W, H, n_win = 346, 260, 4
    s = synth.StereoEventStream(W, H, 1.0e6)
    wins = [s.stereo_window(k)[0] for k in range(n_win)]
    n = W * H

    def view(sae, which, k):
        arr = getattr(sae._h.contents, which)[k]
        return np.ctypeslib.as_array(arr, shape=(n,))

    # sequential reference
    seq = ora.Sae(W, H)
    seq_states = []
    for x, y, t, p in (w[:4] for w in wins):
        seq.update(x, y, t, p)
        seq_states.append((seq.planes(), seq.time_surface(float(t[-1]))))

    # phase A (every rank on its own): local last-event time per pixel and polarity
    local_latest = []
    for x, y, t, p in (w[:4] for w in wins):
        ll = np.zeros((2, n))
        idx = y.astype(np.int64) * W + x
        for pol in (0, 1):
            m = p == pol
            np.maximum.at(ll[pol], idx[m], t[m])          # times ascend: max = last
        local_latest.append(ll)
    # phase B: exclusive "last non-empty" scan -> carry-in of sae_latest for every rank
    carry, acc = [], np.zeros((2, n))
    for ll in local_latest:
        carry.append(acc.copy())
        acc = np.where(ll > 0, ll, acc)
    # phase C (every rank on its own): the normal update from (latest = carry, sae = sentinel)
    SENT = -1.0
    accepted = []
    for k, (x, y, t, p) in enumerate(w[:4] for w in wins):
        r = ora.Sae(W, H)
        for pol in (0, 1):
            view(r, "latest", pol)[:] = carry[k][pol]
            view(r, "sae", pol)[:] = SENT
        r.update(x, y, t, p)
        pl = r.planes()
        accepted.append(np.stack([pl[0].ravel(), pl[1].ravel()]))
        # sae_latest after the window needs no further exchange
        for pol in (0, 1):
            assert np.array_equal(pl[2 + pol], seq_states[k][0][2 + pol]), (k, pol)
    # phase D: inclusive scan of the accepted times -> sae after every window, then the surface
    sae_acc = np.zeros((2, n))
    for k in range(n_win):
        sae_acc = np.where(accepted[k] != SENT, accepted[k], sae_acc)
        for pol in (0, 1):
            assert np.array_equal(sae_acc[pol].reshape(H, W), seq_states[k][0][pol]), (k, pol)
        r = ora.Sae(W, H)
        for pol in (0, 1):
            view(r, "sae", pol)[:] = sae_acc[pol]
        t_ref = float(wins[k][2][-1])
        assert np.array_equal(r.time_surface(t_ref), seq_states[k][1]), k
