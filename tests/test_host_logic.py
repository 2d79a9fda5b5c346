"""Host-side pieces of the package against the oracle / the reference's golden vectors.  CPU only."""
import os
import sys

import numpy as np
import pytest

import __graft_entry__ as ge
from conftest import GOLDEN
from oracle import scene_ref, synth


@pytest.fixture(scope='module')
def pkg():
    return ge.load_package()


def _mod(pkg, name):
    return sys.modules[pkg.__name__ + '.' + name]


def test_calibration_and_focal(pkg):
    cam = _mod(pkg, 'camera')
    kat = np.load(os.path.join(GOLDEN, 'kat_functions.npz'))
    assert np.array_equal(cam.compute_calibration_matrix(1, 100, kat['proj_K'], (1280, 720)), kat['calib_land'])
    assert np.array_equal(cam.compute_calibration_matrix(1, 100, kat['calib_K2'], (512, 512)), kat['calib_square'])
    assert np.array_equal(cam.compute_calibration_matrix(1, 100, kat['calib_K2'], (480, 640)), kat['calib_port'])
    assert abs(cam.get_focal(720, 60) - float(kat['focal_720_60'])) < 1e-9


def test_model_loader_matches_the_oracle_loader(pkg, model_dir):
    io = _mod(pkg, 'smpl_io')
    a = io.load_smpl_model(model_dir)
    b = synth.load_model_tensors(model_dir)
    for k in ('v_template', 'faces', 'shapedirs', 'posedirs', 'J_regressor', 'lbs_weights', 'parents', 'J_regressor_alphapose'):
        assert a[k].dtype == b[k].dtype and np.array_equal(a[k], b[k]), k
    with pytest.raises(FileNotFoundError):
        io.load_smpl_model('/nonexistent/dir')
    # the 17-joint H36M regressor is stored in H36M order and used in the reference layer's row order (smpl.py:240-242)
    raw = np.load(os.path.join(model_dir, 'J_regressor_h36m.npy'))
    assert a['J_regressor_h36m17'].shape == (17, 6890)
    assert np.array_equal(a['J_regressor_h36m17'][0], raw[6].astype(np.float32)) and np.array_equal(a['J_regressor_h36m17'][14], raw[0].astype(np.float32))


def test_fillin_matches_the_reference_loop(pkg):
    sc = _mod(pkg, 'scene')
    rng = np.random.default_rng(0)
    for shape, ks in (((23, 31), 7), ((17, 12), 3), ((20, 20, 3), 11)):
        x = rng.random(shape).astype(np.float32)
        if len(shape) == 3:
            x = (x * 255).astype(np.uint8)
        mask = (rng.random(shape[:2]) > 0.6).astype(np.float32)
        mask[5:12, 4:9] = 0
        ax, am = sc.fillin_values(x, mask, ks)
        bx, bm = scene_ref.fillin_values(x, mask, ks)
        assert np.array_equal(am, bm)
        assert np.array_equal(ax, bx) if x.dtype == np.uint8 else np.allclose(ax, bx, rtol=0, atol=1e-7)
    # a fully valid mask is a no-op, an empty one stays empty
    x = rng.random((8, 9)).astype(np.float32)
    assert np.array_equal(sc.fillin_values(x, np.ones((8, 9), np.float32), 5)[0], x)
    assert sc.fillin_values(x, np.zeros((8, 9), np.float32), 5)[1].max() == 0


def test_scene_aggregation_and_postprocess(pkg):
    pytest.importorskip('cv2')
    sc = _mod(pkg, 'scene')
    rng = np.random.default_rng(1)
    T, H, W = 7, 24, 32
    depths = (2 + rng.random((T, H, W)) * 3).astype(np.float32)
    images = rng.integers(0, 255, (T, H, W, 3), dtype=np.uint8)
    back = (rng.random((T, H, W)) > 0.3).astype(np.float32)
    back[:, 3:6, 3:6] = 0                                    # never seen as background
    ai, ad, am = sc.aggregate_scene_geometry_median(depths, images, back)
    bi, bd, bm = scene_ref.aggregate_scene_median(depths, images, back)
    assert np.array_equal(am, bm) and np.array_equal(ad[am], bd[bm]) and np.array_equal(ai[am], bi[bm])
    pa = sc.postprocess_depthmap(ad, am, use_bilateral_filter=True)
    pb = scene_ref.postprocess_depthmap(bd, bm, use_bilateral_filter=True)
    assert np.allclose(pa, pb, atol=1e-6)


def test_log_from_loss_block(pkg):
    sh = _mod(pkg, 'sharding')
    L = np.arange(16, dtype=np.float32) + 1
    log = sh.log_from_loss_block(L, 4)
    assert log['loss_pose24j'] == 0.25 and log['reg_foot_sliding'] == 7 / 4 and log['reg_vel'] == 8 and log['reg_filter_verts'] == 9
    assert list(log) == ['loss_pose24j', 'loss_depth', 'loss_silhouette', 'reg_ref_poses', 'reg_scale', 'reg_contact',
                         'reg_foot_sliding', 'reg_vel', 'reg_filter_verts']


def test_frame_ranges(pkg):
    sh = _mod(pkg, 'sharding')
    for T, B, world in ((512, 8, 8), (200, 10, 3), (7, 2, 2), (5, 8, 2), (1000, 8, 8), (33, 4, 4)):
        ranges = [sh.frame_range(T, B, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == T
        for (a0, a1), (b0, b1) in zip(ranges[:-1], ranges[1:]):
            assert a1 == b0 and a0 <= a1
        for (a0, a1) in ranges[:-1]:
            assert a0 % B == 0 and (a1 % B == 0 or a1 == T)       # inner edges are batch edges
        sizes = [(a1 - a0 + B - 1) // B for a0, a1 in ranges]
        assert max(sizes) - min(sizes) <= 1
    ranges = [sh.frame_range(5, 8, r, 2) for r in range(2)]
    assert sh.neighbours(0, 2, ranges) == (None, None)              # rank 1 owns nothing -> no neighbour
    ranges = [sh.frame_range(512, 8, r, 8) for r in range(8)]
    assert sh.neighbours(3, 8, ranges) == (2, 4) and sh.neighbours(0, 8, ranges) == (None, 1) and sh.neighbours(7, 8, ranges) == (6, None)


@pytest.mark.skipif(not os.path.isdir('/root/reference/mhmocap'), reason='differential check against the reference needs /root/reference (build container only)')
def test_one_euro_over_time_matches_the_reference_filter(pkg):
    """`get_filtered_vertices_by_smpl` drives the reference's OneEuroFilter with time stamps i / frame_rate (optimizer.py:643-648);
    the host mirror reproduces the class bit for bit (dx0 = zeros: the reference passes the int 0, which its own assert rejects)."""
    sys.path.insert(0, '/root/reference')
    try:
        from mhmocap.one_euro_filter import OneEuroFilter
    finally:
        sys.path.remove('/root/reference')
    opt = _mod(pkg, 'optimizer')
    rng = np.random.default_rng(4)
    for shape, mc, beta in (((9, 3, 1, 3), 0.004, 0.7), ((6, 2, 72), 0.1, 0.1)):
        x = (rng.normal(0, 0.3, shape).cumsum(0)).astype(np.float32)
        ref = x.copy()
        f = OneEuroFilter(0, ref[0], dx0=np.zeros_like(ref[0]), min_cutoff=mc, beta=beta, d_cutoff=1.0)
        for i in range(1, len(ref)):
            ref[i] = f(i / 25, ref[i])
        got = opt.one_euro_over_time(x, mc, beta, 25)
        assert got.dtype == np.float32 and np.array_equal(got, ref)
        assert np.array_equal(got[0], x[0]) and not np.array_equal(got[1:], x[1:])
    one = opt.one_euro_over_time(x[:1], 0.1, 0.1)
    assert np.array_equal(one, x[:1])                                             # a single frame passes through


@pytest.mark.parametrize('count,N,HW', [(1, 1, 1), (3, 8, 1000), (2, 32, 4099), (5, 3, 2048 * 3 + 17)])
def test_host_mask_packing_matches_numpy(pkg, count, N, HW):
    """The host side of the ingest (``optimizer.py:396-409``): float32 {0., 1.} instance masks -> one 32-bit plane per frame, bit n =
    person n, packed by all cores with the widest vector unit of the host (function multi-versioning); ragged sizes that straddle
    the 2048-pixel work items and frame boundaries, and the non-binary flag."""
    L = _mod(pkg, '_lib')
    rng = np.random.default_rng(count * 131 + N)
    seg = (rng.random((count, N, HW)) > 0.6).astype(np.float32)
    out = np.full((count, HW), 0xdeadbeef, np.uint32)
    rc = L.lib.mh_debug_pack_masks(L.ptr(seg), count, N, HW, L.ptr(out))
    assert rc == 0
    ref = np.zeros((count, HW), np.uint32)
    for n in range(N):
        ref |= (seg[:, n] != 0).astype(np.uint32) << np.uint32(n)
    assert np.array_equal(out, ref)
    seg[count - 1, N - 1, HW - 1] = 0.5                                        # not a mask value: reported, still counted as set
    rc = L.lib.mh_debug_pack_masks(L.ptr(seg), count, N, HW, L.ptr(out))
    assert rc == 1 and (out[count - 1, HW - 1] >> np.uint32(N - 1)) & 1 == 1
    assert L.lib.mh_debug_pack_masks(L.ptr(seg), 0, N, HW, L.ptr(out)) < 0     # bad arguments
    assert L.lib.mh_pool_bytes() == 0                                           # nothing parked without a device


def test_sparse_joints_key_is_validated(pkg, model_dir):
    """``smpl_sparse_joints_key`` must name a 17-joint output of the reference's SMPL layer (``optimizer.py:40``, ``smpl.py:368-381``)."""
    with pytest.raises(ValueError, match='smpl_sparse_joints_key'):
        pkg.SMPLDepthSequenceOptimizer(image_size=(64, 48), num_frames=2, cam_K=np.eye(3, dtype=np.float32), device='cuda:0',
                                       smpl_model_parameters_path=model_dir, smpl_sparse_joints_key='joints_smpl24')
