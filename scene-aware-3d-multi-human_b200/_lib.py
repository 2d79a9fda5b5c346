"""ctypes binding of ``libmhopt.so`` (C ABI declared in ``include/mhopt.h``).

There is NO CPU fallback: if the library is missing or fails to load, importing
this module raises.  Every entry point returns 0 or a negative error code; ``check``
turns the code + ``mh_last_error`` into a Python exception.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libmhopt.so')

c_int32, c_int64, c_float = ctypes.c_int32, ctypes.c_int64, ctypes.c_float
c_void_p = ctypes.c_void_p
FP = ctypes.POINTER(ctypes.c_float)
IP = ctypes.POINTER(ctypes.c_int32)

(P_POSES_T, P_POSES_SMPL, P_BETAS, P_ZMIN_LIN, P_ZMAX_LIN, P_XSCALE, P_BETAS_REF) = range(7)
(L_POSE2D, L_DEPTH, L_SILHOUETTE, L_REF_POSES, L_SCALE, L_CONTACT, L_FOOT, L_VEL, L_FILTER_VERTS, L_INIT_2D) = range(10)
L_COUNT = 16
(BUF_SHARED, BUF_HALO_SEND, BUF_HALO_RECV, BUF_CARRY_OUT, BUF_CARRY_IN, BUF_GRADS, BUF_VERTS, BUF_FILTERED,
 BUF_PARAMS, BUF_MEDIAN_HIST, BUF_MEDIAN_AUX) = range(11)
LD3V = 20672
V, F = 6890, 13776


class MhDims(ctypes.Structure):
    _fields_ = [('T', c_int32), ('N', c_int32), ('H', c_int32), ('W', c_int32), ('V', c_int32), ('F', c_int32),
                ('B', c_int32), ('device', c_int32), ('rank', c_int32), ('world', c_int32), ('t0', c_int32),
                ('T_total', c_int32), ('M_max', c_int64)]


class MhModel(ctypes.Structure):
    _fields_ = [('v_template', FP), ('shapedirs', FP), ('posedirs', FP), ('J_regressor', FP), ('lbs_weights', FP),
                ('parents', IP), ('faces', IP), ('reg17', FP)]


class MhCoefs(ctypes.Structure):
    _fields_ = [('proj2d', c_float), ('depth', c_float), ('silhouette', c_float), ('reg_velocity', c_float),
                ('reg_verts_filter', c_float), ('reg_poses', c_float), ('reg_scales', c_float), ('reg_contact', c_float),
                ('reg_foot_sliding', c_float), ('joint_confidence_thr', c_float), ('eps', c_float)]


class MhError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                          f'(nvcc, sm_100a).  There is no CPU fallback.')
    lib = ctypes.CDLL(LIB_PATH)
    ctx = c_void_p
    sig = {
        'mh_create': (c_int32, [ctypes.POINTER(ctx), ctypes.POINTER(MhDims)]),
        'mh_destroy': (None, [ctx]),
        'mh_last_error': (ctypes.c_char_p, [ctx]),
        'mh_version': (ctypes.c_char_p, []),
        'mh_set_batch': (c_int32, [ctx, c_int32]),
        'mh_set_model': (c_int32, [ctx, ctypes.POINTER(MhModel)]),
        'mh_set_camera': (c_int32, [ctx, FP, FP, FP]),
        'mh_set_coefs': (c_int32, [ctx, ctypes.POINTER(MhCoefs)]),
        'mh_set_joint_weights': (c_int32, [ctx, FP]),
        'mh_set_optimize_scale': (c_int32, [ctx, c_int32]),
        'mh_ingest_frames': (c_int32, [ctx, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
        'mh_ingest_frames_u8': (c_int32, [ctx, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
        'mh_finalize_ingest': (c_int32, [ctx, c_void_p]),
        'mh_comm_unique_id': (c_int32, [ctx, c_void_p]),
        'mh_comm_create': (c_int32, [ctx, c_void_p, ctypes.POINTER(c_void_p)]),
        'mh_comm_destroy': (None, [c_void_p]),
        'mh_set_comm': (c_int32, [ctx, c_void_p, c_int32, c_int32]),
        'mh_has_comm': (c_int32, [ctx]),
        'mh_fit_cycle': (c_int32, [ctx, c_float, c_void_p]),
        'mh_fit_cycle_grads': (c_int32, [ctx, c_void_p]),
        'mh_init_cycle': (c_int32, [ctx, c_float, c_int32, c_void_p]),
        'mh_scene_update_from_median': (c_int32, [ctx, c_int32, c_int32, c_void_p, c_void_p]),
        'mh_postprocess_depthmap': (c_int32, [ctx, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
        'mh_set_scene': (c_int32, [ctx, c_void_p, c_int64, c_void_p]),
        'mh_set_scene_from_depth': (c_int32, [ctx, c_void_p, c_void_p, c_void_p]),
        'mh_set_param': (c_int32, [ctx, c_int32, c_void_p, c_int64, c_void_p]),
        'mh_get_param': (c_int32, [ctx, c_int32, c_void_p, c_int64]),
        'mh_get_grad': (c_int32, [ctx, c_int32, c_void_p, c_int64]),
        'mh_device_view': (c_int32, [ctx, c_int32, ctypes.POINTER(c_void_p), ctypes.POINTER(c_int64)]),
        'mh_reset_optimizer': (c_int32, [ctx, c_void_p]),
        'mh_smpl_forward': (c_int32, [ctx, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
        'mh_smpl_regress': (c_int32, [ctx, c_void_p, c_void_p, c_int64, c_void_p, c_int32, c_void_p]),
        'mh_init_begin': (c_int32, [ctx, c_void_p, c_void_p, c_void_p, c_float, c_void_p]),
        'mh_init_grads': (c_int32, [ctx, c_int32, c_int32, c_void_p]),
        'mh_init_update': (c_int32, [ctx, c_float, c_int32, c_void_p]),
        'mh_halo_pack': (c_int32, [ctx, c_void_p]),
        'mh_fit_grads': (c_int32, [ctx, c_int32, c_int32, c_void_p]),
        'mh_fit_update': (c_int32, [ctx, c_float, c_void_p]),
        'mh_read_losses': (c_int32, [ctx, c_void_p, c_void_p]),
        'mh_refresh_filters': (c_int32, [ctx, c_float, c_float, c_float, c_float, c_float, c_int32, c_void_p]),
        'mh_clear_filters': (c_int32, [ctx]),
        'mh_refresh_filters_flag': (c_int32, [ctx, c_int32]),
        'mh_one_euro_filter': (c_int32, [ctx, c_void_p, c_void_p, c_int32, c_int64, c_float, c_float, c_float]),
        'mh_scene_depths': (c_int32, [ctx, c_int32, c_int32, c_void_p]),
        'mh_scene_set_back': (c_int32, [ctx, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
        'mh_scene_median_pass': (c_int32, [ctx, c_int32, c_int32, c_void_p]),
        'mh_scene_median_finish': (c_int32, [ctx, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
        'mh_forward_only': (c_int32, [ctx, c_void_p]),
        'mh_debug_render': (c_int32, [ctx, c_int32, c_int32, c_void_p, c_void_p]),
        'mh_synth_planes': (c_int32, [ctx, c_float, c_float, c_void_p]),
        'mh_read_planes': (c_int32, [ctx, c_int32, c_int32, c_void_p, c_void_p]),
        'mh_set_timing': (c_int32, [ctx, c_int32]),
        'mh_read_timing': (c_int32, [ctx, c_void_p, ctypes.POINTER(c_int32)]),
        'mh_debug_set_render_caps': (c_int32, [ctx, c_int32, c_int32, c_int32]),
        'mh_debug_knn_stats': (c_int32, [ctx, c_void_p, c_void_p]),
        'mh_debug_pack_masks': (c_int32, [c_void_p, c_int32, c_int32, c_int64, c_void_p]),
        'mh_pool_trim': (None, []),
        'mh_pool_bytes': (c_int64, []),
        'mh_debug_gemm_fwd': (c_int32, [ctx, c_void_p, c_void_p, c_int32, c_int32]),
        'mh_debug_gemm_bwd': (c_int32, [ctx, c_void_p, c_void_p, c_int32, c_int32]),
        'mh_render_profile': (c_int32, [ctx, c_int32, c_void_p]),
        'mh_launch_count': (c_int64, [ctx]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)          # AttributeError here = the library does not match include/mhopt.h
        fn.restype = res
        fn.argtypes = args
    return lib, sorted(sig)


lib, SYMBOLS = _load()


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def ptr(a):
    """Host pointer of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags['C_CONTIGUOUS']
    return a.ctypes.data_as(c_void_p)


class Context(object):
    """Owns one ``mh_ctx`` (one per GPU / rank)."""

    def __init__(self, T, N, H, W, B=None, device=0, rank=0, world=1, t0=0, T_total=None, M_max=None):
        self.dims = MhDims(T=T, N=N, H=H, W=W, V=V, F=F, B=B if B else T, device=device, rank=rank, world=world, t0=t0,
                           T_total=T_total if T_total is not None else T, M_max=M_max if M_max is not None else H * W)
        self.h = c_void_p()
        rc = lib.mh_create(ctypes.byref(self.h), ctypes.byref(self.dims))
        if rc != 0:
            msg = lib.mh_last_error(self.h).decode() if self.h else 'mh_create failed'
            if self.h:
                lib.mh_destroy(self.h)
                self.h = c_void_p()
            raise MhError(f'mh_create: {msg} (code {rc})')
        self._keep = []

    def check(self, rc):
        if rc != 0:
            raise MhError(f'{lib.mh_last_error(self.h).decode()} (code {rc})')

    def call(self, name, *args):
        self.check(getattr(lib, name)(self.h, *args))

    def close(self):
        if self.h:
            lib.mh_destroy(self.h)
            self.h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers -------------------------------------------------------------------------------
    def set_model(self, model, reg17=None):
        """model: dict of numpy arrays (keys of ``smpl_io.load_smpl_model``); ``reg17``: the (17, 6890) regressor of the sparse joints
        the 2-D terms use (``smpl_sparse_joints_key`` of the reference, ``optimizer.py:40, 696, 750``; default: AlphaPose)."""
        arrs = {k: f32(model[k]) for k in ('v_template', 'shapedirs', 'posedirs', 'J_regressor', 'lbs_weights')}
        arrs['reg17'] = f32(model['J_regressor_alphapose'] if reg17 is None else reg17)
        parents = np.ascontiguousarray(model['parents'], dtype=np.int32)
        faces = np.ascontiguousarray(model['faces'], dtype=np.int32)
        assert arrs['v_template'].shape == (V, 3) and arrs['shapedirs'].shape == (V, 3, 10), 'SMPL topology expected'
        assert arrs['posedirs'].shape == (207, 3 * V) and arrs['J_regressor'].shape == (24, V)
        assert arrs['lbs_weights'].shape == (V, 24) and faces.shape == (F, 3) and arrs['reg17'].shape == (17, V)
        m = MhModel(v_template=arrs['v_template'].ctypes.data_as(FP), shapedirs=arrs['shapedirs'].ctypes.data_as(FP),
                    posedirs=arrs['posedirs'].ctypes.data_as(FP), J_regressor=arrs['J_regressor'].ctypes.data_as(FP),
                    lbs_weights=arrs['lbs_weights'].ctypes.data_as(FP), parents=parents.ctypes.data_as(IP),
                    faces=faces.ctypes.data_as(IP), reg17=arrs['reg17'].ctypes.data_as(FP))
        self.call('mh_set_model', ctypes.byref(m))

    def set_camera(self, K, Kndc, Kd=None):
        K = f32(K).reshape(9)
        Kndc = f32(Kndc).reshape(16)
        kd = f32(Kd).reshape(5) if Kd is not None else None
        self.call('mh_set_camera', K.ctypes.data_as(FP), Kndc.ctypes.data_as(FP), kd.ctypes.data_as(FP) if kd is not None else None)

    def set_coefs(self, **kw):
        c = MhCoefs(**{k: float(v) for k, v in kw.items()})
        self.call('mh_set_coefs', ctypes.byref(c))

    def set_param(self, which, arr, stream=None):
        a = f32(arr).reshape(-1)
        self._keep.append(a)
        self.call('mh_set_param', which, ptr(a), a.size, stream)

    def get_param(self, which, shape):
        out = np.empty(shape, np.float32)
        self.call('mh_get_param', which, ptr(out), out.size)
        return out

    def get_grad(self, which, shape):
        out = np.empty(shape, np.float32)
        self.call('mh_get_grad', which, ptr(out), out.size)
        return out

    def device_view(self, which):
        p, n = c_void_p(), c_int64()
        self.call('mh_device_view', which, ctypes.byref(p), ctypes.byref(n))
        return p.value, n.value

    def read_losses(self, stream=None):
        out = np.zeros(L_COUNT, np.float32)
        self.call('mh_read_losses', ptr(out), stream)
        return out

    def read_timing(self, n=64):
        out = np.zeros((n, 6), np.float32)
        k = c_int32(n)
        self.call('mh_read_timing', ptr(out), ctypes.byref(k))
        return out[:k.value]

    def launches(self):
        return int(lib.mh_launch_count(self.h))
