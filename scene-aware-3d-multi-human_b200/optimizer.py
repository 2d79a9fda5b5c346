"""Drop-in replacement of the reference's ``mhmocap.optimizer`` for the B200.

Keeps the public interface of ``SMPLDepthSequenceOptimizer`` (``mhmocap/optimizer.py:146-770``) -- constructor
keywords, ``init_optimized_variables``, ``fit``, ``get_optimized_variables``, ``update_scene_pointcloud``,
``one_euro_filter`` -- so that the reference's ``Predictor`` (``mhmocap/predict.py:290-306, 332-344``) drives it
unchanged, while every per-iteration computation runs in ``libmhopt.so`` (hand-written sm_100a kernels behind the C
ABI of ``include/mhopt.h``).  There is no CPU path: constructing the optimiser without a CUDA device raises.

Differences from the reference that a caller can observe (all documented in DESIGN.md):
  * the dataloader is consumed ONCE (first cycle): the modalities stay resident in HBM, keyed by ``idxs``;
    with ``shuffle=False`` the objective is identical, with ``shuffle=True`` the reference's foot-sliding term
    pairs random frames (``optimizer.py:512-518``) whereas this implementation always pairs frame t with t-1
    inside batches of ``batch_size`` consecutive frames;
  * ``fit(num_iter <= 30)`` returns (scene outputs ``None``) instead of raising ``UnboundLocalError``
    (``optimizer.py:595``);
  * with ``torch.distributed`` initialised, frames are sharded over the ranks (``sharding.py``).  Shard edges must be
    multiples of the dataloader's batch size, which the reference API only reveals in ``fit``: without the
    ``batch_size=`` extension of ``init_optimized_variables`` every rank runs the (cheap) translation init on the WHOLE
    sequence and the frames are sharded when ``fit`` sees the first batch.
"""
import ctypes
import math
import os

import numpy as np
import torch

from . import _lib
from . import camera
from . import scene as scene_ops
from . import sharding
from . import smpl_io

L = _lib
# outputs of the reference's SMPL layer with 17 joints (smpl.py:368-381); 'joints_mupots' only exists there when the layer is given
# that regressor, which the reference optimiser does not do (optimizer.py:69-72) -- accepted here as an extension
_SPARSE_REGRESSORS = {'joints_alphapose': 'J_regressor_alphapose', 'joints_h36m17': 'J_regressor_h36m17', 'joints_mupots': 'J_regressor_mupots'}
_LIB_COMMS = {}              # (process group, rank, world, device) -> handle of mh_comm_create, kept for the life of the process


def _device_ordinal(device):
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError('scene-aware-3d-multi-human_b200 needs a CUDA device (sm_100a); there is no CPU path')
        return torch.cuda.current_device()
    dev = torch.device(device)
    if dev.type != 'cuda':
        raise RuntimeError(f'device {dev} requested: this implementation runs on CUDA devices only (no CPU path)')
    return dev.index if dev.index is not None else torch.cuda.current_device()


class _DevView(object):
    """Exposes a raw device pointer of the library as a ``__cuda_array_interface__`` object."""
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {'shape': (int(n),), 'typestr': '<f4', 'data': (int(ptr), False), 'version': 2}


def one_euro_over_time(x, min_cutoff, beta, frame_rate=25, d_cutoff=1.0):
    """One-Euro filter of ``x (T, ...)`` along axis 0 with time stamps ``i / frame_rate`` (``one_euro_filter.py:16-53`` driven as in
    ``optimizer.py:643-648``: first sample passes through, ``dx0 = 0``).  Host numpy: the arrays are T x N x 75 floats."""
    x = np.array(x, copy=True)
    x_prev, dx_prev, t_prev = x[0].copy(), np.zeros_like(x[0]), 0.0

    def alpha(t_e, cutoff):
        r = 2 * math.pi * cutoff * t_e
        return r / (r + 1)

    for i in range(1, len(x)):
        t = i / frame_rate
        t_e = t - t_prev
        dx = (x[i] - x_prev) / t_e
        a_d = alpha(t_e, d_cutoff)
        dx_hat = a_d * dx + (1 - a_d) * dx_prev
        a = alpha(t_e, min_cutoff + beta * np.abs(dx_hat))
        x_hat = a * x[i] + (1 - a) * x_prev
        # the reference stores the time stamp through its mask, i.e. as an array of x's dtype: from the second step on t_e is float32
        x_prev, dx_prev, t_prev = x_hat, dx_hat, np.full_like(x_hat, t)
        x[i] = x_hat
    return x


class SMPLOptimizerBase(object):
    """Model paths and joint weights of ``SMPLOptimizerBase.__init__`` (``optimizer.py:35-131``)."""

    def __init__(self, device=None, smpl_model_parameters_path='model_data/parameters',
                 smpl_J_reg_extra_path='J_regressor_extra.npy', smpl_J_reg_h37m_path='J_regressor_h36m.npy',
                 smpl_J_reg_alphapose_path='SMPL_AlphaPose_Regressor_RMSprop_6.npy',
                 smpl_sparse_joints_key='joints_alphapose', pose24j_weights=None, pose17j_weights=None,
                 process_group=None, scene_update=True, max_scene_points=None, allow_partial_loader=False):
        self.device_ordinal = _device_ordinal(device)
        self.device = torch.device('cuda', self.device_ordinal)
        # the sparse joints the 2-D terms compare with ``pose2d`` (``optimizer.py:40, 696, 750``): either 17-joint regressor of the
        # reference's SMPL layer (``smpl.py:376-381``)
        if smpl_sparse_joints_key not in _SPARSE_REGRESSORS:
            raise ValueError(f"smpl_sparse_joints_key must be one of {sorted(_SPARSE_REGRESSORS)} (17-joint regressors, "
                             f"smpl.py:376-381), got '{smpl_sparse_joints_key}'")
        self.smpl_model_parameters_path = os.path.abspath(smpl_model_parameters_path)
        self.smpl_sparse_joints_key = smpl_sparse_joints_key
        self.model = smpl_io.load_smpl_model(smpl_model_parameters_path, alphapose_regressor=smpl_J_reg_alphapose_path,
                                             h36m_regressor=smpl_J_reg_h37m_path)
        if _SPARSE_REGRESSORS[smpl_sparse_joints_key] not in self.model:
            raise FileNotFoundError(f"the regressor of '{smpl_sparse_joints_key}' is not in {smpl_model_parameters_path}")
        self.faces_smpl = self.model['faces']
        w24 = np.ones(24, np.float32) if pose24j_weights is None else np.array(pose24j_weights, np.float32)
        self.pose24j_weights = len(w24) * w24 / np.sum(w24)
        w17 = np.ones(17, np.float32) if pose17j_weights is None else np.array(pose17j_weights, np.float32)
        self.pose17j_weights = (len(w17) * w17 / np.sum(w17)).astype(np.float32)          # optimizer.py:127-130
        self.group = process_group
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.rank = torch.distributed.get_rank(process_group)
            self.world = torch.distributed.get_world_size(process_group)
        else:
            self.rank, self.world = 0, 1
        self.scene_update = scene_update
        self.max_scene_points = max_scene_points
        # extension: with several ranks the dataloader of a rank may deliver only that rank's frames (the reference API iterates the
        # whole sequence on every process)
        self.partial_loader_ok = bool(allow_partial_loader)


class SMPLDepthSequenceOptimizer(SMPLOptimizerBase):
    """Space-time SMPL fit of a whole sequence (``optimizer.py:146-770``) on ``libmhopt.so``."""

    def __init__(self, image_size, num_frames, fov=60, focal_length=None, znear=1.0, zfar=100.0, cam_K=None,
                 cam_dist_coef=None, proj2d_loss_coef=1.0, depth_loss_coef=1.0, silhouette_loss_coef=1.0,
                 reg_velocity_coef=1.0, reg_verts_filter_coef=1.0, reg_poses_coef=1.0, reg_scales_coef=1.0,
                 reg_contact_coef=1.0, reg_foot_sliding_coef=1.0, joint_confidence_thr=0.5, eps=1e-3, **kargs):
        """``image_size`` is (W, H) as the callers pass it (``predict.py:292``, ``datautils.py:601``)."""
        super().__init__(**kargs)
        if focal_length is None:
            focal_length = camera.get_focal(min(image_size), fov)
        if cam_K is None:
            # the reference's fallback (optimizer.py:193-197), axes as written there
            self.cam_K = np.array([[focal_length, 0, image_size[1] / 2.0], [0, focal_length, image_size[0] / 2.0],
                                   [0, 0, 1]], dtype=np.float32)
        else:
            self.cam_K = np.asarray(cam_K).astype(np.float32)
        self.cam_dist_coef = cam_dist_coef
        self.znear, self.zfar = znear, zfar
        self.K_ndc = camera.compute_calibration_matrix(znear, zfar, self.cam_K, image_size)       # optimizer.py:206
        self.coefs = dict(proj2d=proj2d_loss_coef, depth=depth_loss_coef, silhouette=silhouette_loss_coef,
                          reg_velocity=reg_velocity_coef, reg_verts_filter=reg_verts_filter_coef, reg_poses=reg_poses_coef,
                          reg_scales=reg_scales_coef, reg_contact=reg_contact_coef, reg_foot_sliding=reg_foot_sliding_coef,
                          joint_confidence_thr=joint_confidence_thr, eps=eps)
        self.proj2d_loss_coef, self.depth_loss_coef, self.silhouette_loss_coef = proj2d_loss_coef, depth_loss_coef, silhouette_loss_coef
        self.reg_velocity_coef, self.reg_verts_filter_coef, self.reg_poses_coef = reg_velocity_coef, reg_verts_filter_coef, reg_poses_coef
        self.reg_scales_coef, self.reg_contact_coef, self.reg_foot_sliding_coef = reg_scales_coef, reg_contact_coef, reg_foot_sliding_coef
        self.joint_confidence_thr, self.eps = joint_confidence_thr, eps
        self.num_frames = num_frames
        self.img_w, self.img_h = image_size
        self.ctx = None
        self.scene_depth = None
        self.scene_pcd = None
        self.scene_img = None
        self.scene_mask = None
        self.poses_T_filtered = None
        self.verts_filtered = None
        self._ingested = False
        self._images = None
        self._backmasks = None

    # ------------------------------------------------------------------------------------------ context
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _make_context(self, T_total, N, B=None):
        """Device context of this rank.  With several ranks and the batch size still unknown (the reference API passes it
        only through the dataloader of ``fit``) the context covers the WHOLE sequence, replicated, and ``_reshard`` replaces
        it when the first batch arrives; ``self._dist`` says whether the per-cycle exchanges run."""
        self.num_people = N
        self.T_total = T_total
        Bq = B if B else T_total
        self._dist = self.world > 1 and bool(B)
        if self._dist:
            self.ranges = [sharding.frame_range(T_total, Bq, r, self.world) for r in range(self.world)]
            self.t0, self.t1 = self.ranges[self.rank]
            empty = [r for r, (a, b) in enumerate(self.ranges) if b <= a]
            if empty:
                # raised identically on EVERY rank (it only depends on T, B, world): nobody is left waiting in a collective
                raise RuntimeError(f'ranks {empty} own no frames: {T_total} frames in batches of {Bq} over {self.world} ranks; '
                                   f'use at most {(T_total + Bq - 1) // Bq} ranks')
            self.prev, self.next = sharding.neighbours(self.rank, self.world, self.ranges)
        else:
            self.ranges = [(0, T_total)]
            self.t0, self.t1 = 0, T_total
            self.prev = self.next = None
        self.T_local = self.t1 - self.t0
        torch.cuda.set_device(self.device)
        self.ctx = L.Context(self.T_local, N, self.img_h, self.img_w, B=Bq, device=self.device_ordinal,
                             rank=self.rank if self._dist else 0, world=self.world if self._dist else 1, t0=self.t0, T_total=T_total,
                             M_max=self.max_scene_points if self.max_scene_points else self.img_h * self.img_w)
        self.ctx.set_model(self.model, self.sparse_regressor())
        self.ctx.set_camera(self.cam_K, self.K_ndc, self.cam_dist_coef)
        self.ctx.set_coefs(**self.coefs)
        w = np.ascontiguousarray(self.pose17j_weights, np.float32)
        self.ctx.call('mh_set_joint_weights', w.ctypes.data_as(L.FP))
        self._views = {}
        self._lib_comm = self._dist and self._setup_lib_comm()

    def sparse_regressor(self):
        """(17, 6890) regressor of ``smpl_sparse_joints_key``.  The layer returns 'joints_h36m17' relative to its pelvis joint
        (``smpl.py:369-373``: ``joints - joints[:, 14]``), which is the regressor with row 14 subtracted from every row; its rows sum
        to 0, so the body translation enters the absolute joints once, as ``scale * joints + poses_T`` does (``optimizer.py:750``)."""
        reg = self.model[_SPARSE_REGRESSORS[self.smpl_sparse_joints_key]]
        if self.smpl_sparse_joints_key == 'joints_h36m17':
            reg = np.ascontiguousarray(reg - reg[14:15])
        return reg

    def _setup_lib_comm(self):
        """Hand the per-cycle exchanges (halo frames, all-reduce of the shared leaves) to a communicator the library owns
        (``mh_set_comm``: NCCL bound at run time), so that one C call enqueues a whole cycle.  Falls back to ``torch.distributed`` on
        the library's device buffers when NCCL cannot be loaded, when the process group is not NCCL (the CPU tests use gloo) or
        when ``MH_LIB_COMM=0``.  Collective: every rank of the group takes the same branch."""
        dist = torch.distributed
        if not (dist.is_available() and dist.is_initialized()):
            return False                                                    # ranks emulated inside one process (tests)
        if os.environ.get('MH_LIB_COMM', '1') == '0' or dist.get_backend(self.group) != 'nccl':
            return False
        # the communicator is created once per process and process group (ncclCommInitRank takes seconds on an 8-GPU node) and
        # attached to every context this process creates afterwards
        key = (id(self.group) if self.group is not None else 0, self.rank, self.world, self.device_ordinal)
        prev, nxt = (-1 if self.prev is None else self.prev), (-1 if self.next is None else self.next)
        if key in _LIB_COMMS:
            self.ctx.call('mh_set_comm', _LIB_COMMS[key], prev, nxt)
            return True
        uid = np.zeros(128, np.uint8)
        ok = 1
        if self.rank == 0:
            try:
                self.ctx.call('mh_comm_unique_id', L.ptr(uid))
            except L.MhError:
                ok = 0
        box = [(ok, uid.tobytes())]
        src = dist.get_global_rank(self.group, 0) if self.group is not None else 0
        dist.broadcast_object_list(box, src=src, group=self.group)
        ok, raw = box[0]
        if not ok:
            return False
        uid = np.frombuffer(raw, np.uint8).copy()
        flag = torch.ones(1, device=self.device)
        handle = ctypes.c_void_p()
        try:
            self.ctx.call('mh_comm_create', L.ptr(uid), ctypes.byref(handle))
        except L.MhError:
            flag.zero_()
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)         # all or nobody
        if flag.item() <= 0:
            return False
        _LIB_COMMS[key] = handle
        self.ctx.call('mh_set_comm', handle, prev, nxt)
        return True

    def _view(self, which):
        """torch tensor aliasing a device buffer of the library (for the NCCL plumbing)."""
        if which not in self._views:
            p, n = self.ctx.device_view(which)
            self._views[which] = torch.as_tensor(_DevView(p, n), device=self.device)
        return self._views[which]

    def _set_batch(self, B):
        if self.world > 1:
            if not self._dist:
                self._reshard(B)
            elif [sharding.frame_range(self.T_total, B, r, self.world) for r in range(self.world)] != self.ranges:
                raise RuntimeError(f'the dataloader batch size {B} changes the frame sharding chosen at init; pass '
                                   f'batch_size={B} to init_optimized_variables (or none at all)')
        self.ctx.call('mh_set_batch', B)
        self.batch_size = B

    def _leaves_to_host(self):
        """All six leaves of the local context as host arrays (whole-sequence shapes when the context is replicated)."""
        ctx, N, T = self.ctx, self.num_people, self.T_local
        return dict(poses_T=ctx.get_param(L.P_POSES_T, (T, N, 3)), poses_smpl=ctx.get_param(L.P_POSES_SMPL, (T, N, 72)),
                    betas=ctx.get_param(L.P_BETAS, (1, N, 10)), betas_ref=ctx.get_param(L.P_BETAS_REF, (1, N, 10)),
                    zmin_lin=ctx.get_param(L.P_ZMIN_LIN, (T,)), zmax_lin=ctx.get_param(L.P_ZMAX_LIN, (T,)),
                    xscale=ctx.get_param(L.P_XSCALE, (N,)))

    def _reshard(self, B):
        """Replace the replicated whole-sequence context of the init stage by this rank's frame shard (batch size ``B`` is
        known now); the leaves move over, frame-indexed ones sliced to the shard."""
        s = self._leaves_to_host()
        N = self.num_people
        self.ctx.close()
        self._make_context(self.T_total, N, B)
        ctx, st = self.ctx, self._stream()
        sl = slice(self.t0, self.t1)
        ctx.call('mh_set_optimize_scale', int(self.optim_scale_factor))
        ctx.set_param(L.P_XSCALE, s['xscale'], st)
        ctx.set_param(L.P_POSES_T, s['poses_T'][sl], st)
        ctx.set_param(L.P_POSES_SMPL, s['poses_smpl'][sl], st)
        ctx.set_param(L.P_BETAS, s['betas'], st)
        ctx.set_param(L.P_BETAS_REF, s['betas_ref'], st)
        ctx.set_param(L.P_ZMIN_LIN, s['zmin_lin'][sl], st)
        ctx.set_param(L.P_ZMAX_LIN, s['zmax_lin'][sl], st)
        ctx.call('mh_clear_filters')
        ctx.call('mh_set_scene', None, 0, st)
        torch.cuda.current_stream(self.device).synchronize()

    def _exchange_halo(self):
        if not self._dist:
            return 0, 0
        self.ctx.call('mh_halo_pack', self._stream())
        n = self.num_people * 75
        send = self._view(L.BUF_HALO_SEND).view(2, n)
        recv = self._view(L.BUF_HALO_RECV).view(2, n)
        hp, hn = sharding.exchange_halo(send, recv, self.prev, self.next, self.group)
        return int(hp), int(hn)

    # ------------------------------------------------------------------------------------------ init
    def init_optimized_variables(self, pose2d, poses_smpl, betas_smpl, valid_smpl, scale_factor=None, num_iter=100,
                                 batch_size=None):
        """``optimizer.py:262-321``; arrays cover the WHOLE sequence on every rank.  ``batch_size`` (extension) fixes
        the frame sharding when it is known before ``fit``."""
        assert (pose2d.shape[:2] == poses_smpl.shape[:2] == betas_smpl.shape[:2] == valid_smpl.shape[:2]), (
            f'Error: invalid inputs {pose2d.shape}, {poses_smpl.shape}, {betas_smpl.shape}, {valid_smpl.shape}')
        T, N = pose2d.shape[0:2]
        assert tuple(pose2d.shape[2:]) == (17, 3), f'pose2d must be (T, N, 17, 3) AlphaPose joints, got {pose2d.shape}'
        assert tuple(poses_smpl.shape[2:]) == (72,) and tuple(betas_smpl.shape[2:]) == (10,), (poses_smpl.shape, betas_smpl.shape)
        if self.ctx is not None:
            self.ctx.close()
        self._make_context(T, N, batch_size)
        ctx, st = self.ctx, self._stream()
        sl = slice(self.t0, self.t1)
        if scale_factor is not None:
            xs = (np.log(scale_factor) / np.log(1.1)).astype(np.float32)
            self.optim_scale_factor = False
        else:
            xs = np.zeros(N, np.float32)
            self.optim_scale_factor = True
        ctx.call('mh_set_optimize_scale', int(self.optim_scale_factor))
        ctx.set_param(L.P_XSCALE, xs.reshape(N), st)
        optim_log = self.__init_global_poses(pose2d[sl], poses_smpl[sl], betas_smpl[sl], num_iter)
        # remaining leaves (optimizer.py:291-303)
        poses_T_local = ctx.get_param(L.P_POSES_T, (self.T_local, N, 3))
        max_z = np.clip(np.max(poses_T_local[..., 2], axis=1), 2, None)
        ctx.set_param(L.P_POSES_SMPL, poses_smpl[sl], st)
        avg_betas = np.mean(betas_smpl, axis=0, keepdims=True).astype(np.float32)              # (1, N, 10), mean over ALL frames
        ctx.set_param(L.P_BETAS, avg_betas, st)
        ctx.set_param(L.P_BETAS_REF, avg_betas, st)
        self.valid_smpl = (valid_smpl > 0.7).astype(np.float32)
        ctx.set_param(L.P_ZMIN_LIN, np.ones_like(max_z), st)
        ctx.set_param(L.P_ZMAX_LIN, 2.0 * max_z, st)
        self.scene_depth = self.scene_pcd = None
        self.poses_T_filtered = self.verts_filtered = None
        ctx.call('mh_clear_filters')
        ctx.call('mh_set_scene', None, 0, st)
        self._ingested = False
        return optim_log

    def __init_global_poses(self, pose2d, poses_smpl, betas_smpl, num_iter, joints_thr=0.15):
        """Hot loop A (``optimizer.py:710-770``): Adam(lr .5, betas (.5,.5), eps 1e-6) + ExponentialLR(.95) on poses_T."""
        ctx, st = self.ctx, self._stream()
        p2d, th, be = L.f32(pose2d), L.f32(poses_smpl), L.f32(betas_smpl)
        ctx.call('mh_init_begin', L.ptr(p2d), L.ptr(th), L.ptr(be), joints_thr, st)
        lr = 0.5
        log = []
        count = float(self.T_total * self.num_people * 17 * 2)
        for it in range(num_iter):
            if self._lib_comm or not self._dist:
                ctx.call('mh_init_cycle', lr, it + 1, st)                  # halo, gradients, all-reduce, Adam step: one call
            else:
                hp, hn = self._exchange_halo()
                ctx.call('mh_init_grads', hp, hn, st)
                sharding.allreduce_shared(self._view(L.BUF_SHARED), self.group)
                ctx.call('mh_init_update', lr, it + 1, st)
            lr *= 0.95
            losses = ctx.read_losses(st)
            log.append({'loss_2d': np.float32(losses[L.L_INIT_2D] / count)})
        return log

    # ------------------------------------------------------------------------------------------ ingest
    def _ingest(self, dataloader):
        """One pass over the dataloader (``optimizer.py:394-400``): every modality goes to the device once."""
        ctx, st = self.ctx, self._stream()
        H, W, N = self.img_h, self.img_w, self.num_people
        seen = np.zeros(self.T_total, bool)
        self.h2d_bytes = 0
        B = None
        stream = torch.cuda.current_stream(self.device)
        inflight = []                                  # (event, host buffers): the copies are asynchronous, the buffers may be temporaries
        keep_scene = self.scene_update
        self._have_images = False
        for data in dataloader:
            idxs = np.asarray(data['idxs']).astype(np.int64).reshape(-1)
            if B is None:
                B = len(idxs)
                self._set_batch(B)
                ctx = self.ctx                         # _set_batch may have replaced the replicated init context by this rank's shard
            arr = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in data.items()
                   if k in ('depths', 'seg_mask', 'pose2d', 'poses_smpl', 'images', 'backmasks')}
            seen[idxs] = True
            # runs of consecutive frames owned by this rank go to the device in one call
            j = 0
            while j < len(idxs):
                t = int(idxs[j])
                if not (self.t0 <= t < self.t1):
                    j += 1
                    continue
                e = j + 1
                while e < len(idxs) and int(idxs[e]) == int(idxs[e - 1]) + 1 and int(idxs[e]) < self.t1:
                    e += 1
                tl, cnt = t - self.t0, e - j
                assert tuple(arr['pose2d'].shape[2:]) == (17, 3), f"pose2d must be (B, N, 17, 3), got {arr['pose2d'].shape}"
                dep = L.f32(arr['depths'][j:e]); p2d = L.f32(arr['pose2d'][j:e])
                th = L.f32(arr['poses_smpl'][j:e]); vl = L.f32(self.valid_smpl[t:t + cnt].reshape(cnt, N))
                seg = arr['seg_mask'][j:e]
                if seg.dtype in (np.uint8, np.bool_):                      # compact masks: a quarter of the bytes (extension)
                    seg = np.ascontiguousarray(seg).view(np.uint8)
                    entry = 'mh_ingest_frames_u8'
                else:                                                      # float32 {0., 1.} as the reference dataset delivers them
                    seg = L.f32(seg)
                    entry = 'mh_ingest_frames'
                assert dep.shape == (cnt, H, W) and seg.shape == (cnt, N, H, W), (dep.shape, seg.shape)
                ctx.call(entry, tl, cnt, L.ptr(dep), L.ptr(seg), L.ptr(p2d), L.ptr(th), L.ptr(vl), st)
                held = [dep, seg, p2d, th, vl]
                self.h2d_bytes += dep.nbytes + seg.nbytes + p2d.nbytes + th.nbytes + vl.nbytes
                if keep_scene:
                    # backmasks / images stay on the device for the scene median (optimizer.py:399-400, 579-582)
                    bk = np.ascontiguousarray(np.asarray(arr['backmasks'][j:e]) != 0).astype(np.uint8)
                    im = np.ascontiguousarray(arr['images'][j:e], dtype=np.uint8) if 'images' in arr else None
                    assert bk.shape == (cnt, H, W) and (im is None or im.shape == (cnt, H, W, 3))
                    ctx.call('mh_scene_set_back', tl, cnt, L.ptr(bk), L.ptr(im), st)
                    held += [bk, im]
                    self._have_images = im is not None
                    self.h2d_bytes += bk.nbytes + (im.nbytes if im is not None else 0)
                # keep the host buffers of the last two calls alive instead of synchronising after every call: the next batch is
                # fetched / converted on the host while this one is still being copied
                ev = torch.cuda.Event()
                ev.record(stream)
                inflight.append((ev, held))
                while len(inflight) > 2:
                    inflight.pop(0)[0].synchronize()
                j = e
        stream.synchronize()
        inflight.clear()
        if not seen[self.t0:self.t1].all() or not (seen.all() or self.partial_loader_ok):
            raise RuntimeError(f'the dataloader did not deliver frames {np.nonzero(~seen)[0][:8]}...')
        ctx.call('mh_finalize_ingest', st)
        self._ingested = True

    # ------------------------------------------------------------------------------------------ fit
    def fit(self, dataloader, num_iter=250, min_cutoff1=0.01, min_cutoff2=0.001, beta1=0.02, beta2=0.5,
            update_filters_every=25, verbose=False, start_cycle=0):
        """Hot loop B (``optimizer.py:324-602``): RMSprop(lr .01, alpha .5, momentum .9) + ExponentialLR(.99), one
        step per pass over the video.  Returns the reference's ``optim_log``.  ``start_cycle`` (extension) resumes the cycle
        schedule -- learning rate, filter refreshes, scene updates -- at that cycle: cycles ``start_cycle .. num_iter - 1`` run."""
        ctx, st = self.ctx, self._stream()
        if not self._ingested:
            import time
            t0 = time.perf_counter()
            self._ingest(dataloader)
            self.ingest_seconds = time.perf_counter() - t0
        if not self.optim_scale_factor:
            print('WARNING!!! Not optimizing scale_factor!')
        ctx.call('mh_reset_optimizer', st)
        n_batches = (self.T_total + self.batch_size - 1) // self.batch_size
        lr = 0.01 * 0.99 ** start_cycle
        optim_log = []
        cycles = range(start_cycle, num_iter)
        if verbose:
            from tqdm import tqdm
            cycles = tqdm(cycles)
        ma = None
        for cycle in cycles:
            if (cycle >= 30) and (cycle % update_filters_every == 0):
                self._refresh_filters(min_cutoff1, beta1, min_cutoff2, beta2)
            if cycle >= 30 and self.scene_update:
                # the scene is rebuilt between this cycle's gradients (which still see the previous cloud) and the step (optimizer.py:578-587)
                self._cycle(None)
                self._update_scene_geometry(read_back=(cycle == num_iter - 1))
                ma = True
                ctx.call('mh_fit_update', lr, st)
            else:
                self._cycle(lr)
            lr *= 0.99
            losses = ctx.read_losses(st)
            optim_log.append(sharding.log_from_loss_block(losses, n_batches))
        if ma is not None:
            # the median image only depends on constant inputs: evaluated once instead of every cycle (optimizer.py:581, 595-600)
            scene_mask = self._median_result(0)[1].copy()
            scene_img = self._device_median(1) if self._have_images else None
            if scene_img is not None:
                while scene_mask.min() == 0:
                    scene_img, scene_mask = scene_ops.fillin_values(scene_img, scene_mask, filter_size=11)
            self.scene_img, self.scene_mask = scene_img, scene_mask
        return optim_log

    def _cycle(self, lr):
        """One optimisation cycle: halo exchange, gradients, all-reduce of the shared leaves and -- unless ``lr`` is None -- the
        RMSprop step."""
        st = self._stream()
        if self._lib_comm or not self._dist:                                # one C call, nothing but kernels and NCCL on the stream
            if lr is None:
                self.ctx.call('mh_fit_cycle_grads', st)
            else:
                self.ctx.call('mh_fit_cycle', lr, st)
            return
        hp, hn = self._exchange_halo()
        self.ctx.call('mh_fit_grads', hp, hn, st)
        sharding.allreduce_shared(self._view(L.BUF_SHARED), self.group)
        if lr is not None:
            self.ctx.call('mh_fit_update', lr, st)

    def step_device_only(self, lr):
        """One cycle without any host readback (bench inner loop)."""
        self._cycle(lr)

    def _refresh_filters(self, mc1, b1, mc2, b2, frame_rate=25):
        """``optimizer.py:383-392``: the One-Euro scan is sequential in time, so the ranks run it one after the other,
        handing the filter state over; then the filtered boundary frames are exchanged for the halo slots."""
        ctx, st = self.ctx, self._stream()
        if self._dist:
            sharding.pass_carry(None, self._view(L.BUF_CARRY_IN), self.prev, self.next, self.group)
        ctx.call('mh_refresh_filters', mc1, b1, mc2, b2, float(frame_rate), int(self.prev is None), st)
        if self._dist:
            sharding.send_carry(self._view(L.BUF_CARRY_OUT), self.next, self.group)
            row = self.num_people * L.LD3V
            F = self._view(L.BUF_FILTERED).view(self.T_local + 2, row)
            send = torch.stack([F[1], F[self.T_local]])
            recv = torch.empty_like(send)
            hp, hn = sharding.exchange_halo(send, recv, self.prev, self.next, self.group)
            if hp:
                F[0].copy_(recv[0])
            if hn:
                F[self.T_local + 1].copy_(recv[1])
        self.poses_T_filtered = True
        self.verts_filtered = True

    def _device_median(self, which):
        """Masked temporal median over ALL frames (``fhsog.py:180-202``) by exact radix selection on the device; the
        per-pixel digit histograms are summed over the ranks between passes (``csrc/mh_scene.cu``).  Returns host arrays."""
        self._median_passes(which)
        return self._median_result(which)

    def _median_passes(self, which):
        ctx, st = self.ctx, self._stream()
        HW = self.img_h * self.img_w
        npass = 10 if which == 0 else 4
        planes = 1 if which == 0 else 3
        for p in range(npass):
            ctx.call('mh_scene_median_pass', which, p, st)
            if self._dist:
                if p < npass - 1:
                    hist = self._view(L.BUF_MEDIAN_HIST)[:(1 if p == 0 else 16 * planes) * HW]
                    torch.distributed.all_reduce(hist, op=torch.distributed.ReduceOp.SUM, group=self.group)
                else:
                    aux = self._view(L.BUF_MEDIAN_AUX)
                    torch.distributed.all_reduce(aux[:planes * HW], op=torch.distributed.ReduceOp.SUM, group=self.group)
                    torch.distributed.all_reduce(aux[3 * HW:(3 + planes) * HW], op=torch.distributed.ReduceOp.MIN, group=self.group)

    def _median_result(self, which):
        ctx, st = self.ctx, self._stream()
        if which == 0:
            depth = np.empty((self.img_h, self.img_w), np.float32)
            mask = np.empty((self.img_h, self.img_w), np.uint8)
            ctx.call('mh_scene_median_finish', 0, L.ptr(depth), L.ptr(mask), None, st)
            return depth, mask.astype(bool)
        img = np.empty((self.img_h, self.img_w, 3), np.uint8)
        ctx.call('mh_scene_median_finish', 1, None, None, L.ptr(img), st)
        return img

    def _update_scene_geometry(self, read_back=False):
        """``optimizer.py:578-584``, all on the device: masked temporal median of the per-frame scene depths (``mh_scene.cu``) ->
        ``postprocess_depthmap`` (bilateral filter, Sobel edge mask, erosions, fill-in sweeps: ``mh_scenepost.cu``) -> scene point
        cloud.  Every rank computes the same map.  ``read_back``: fetch the post-processed depth map (the ``scene_depth`` output)."""
        self._median_passes(0)
        out = np.empty((self.img_h, self.img_w), np.float32) if read_back else None
        self.ctx.call('mh_scene_update_from_median', 1, 7, L.ptr(out), self._stream())
        self.scene_pcd = True
        self.scene_depth = out if read_back else True

    def postprocess_depthmap(self, depth, mask=None, fillin_ksize=7, use_bilateral_filter=False):
        """``mhmocap.utils.postprocess_depthmap`` (``utils.py:174-209``) on the device (``scene.postprocess_depthmap`` is the host
        mirror the parity tests compare it with)."""
        d = L.f32(depth)
        m = None if mask is None else np.ascontiguousarray(np.asarray(mask) > 0).astype(np.uint8)
        out = np.empty_like(d)
        self.ctx.call('mh_postprocess_depthmap', L.ptr(d), L.ptr(m), int(bool(use_bilateral_filter)), int(fillin_ksize), L.ptr(out), self._stream())
        return out

    def update_scene_pointcloud(self, scene_depth, scene_mask):
        """``optimizer.py:605-616``: inverse-project the pixel centres with the scene depth, keep ``mask > 0.5``."""
        d = L.f32(scene_depth)
        m = np.ascontiguousarray(np.asarray(scene_mask).astype(np.float32) > 0.5).astype(np.uint8)
        self.ctx.call('mh_set_scene_from_depth', L.ptr(d), L.ptr(m), self._stream())
        self.scene_depth = scene_depth
        self.scene_pcd = True

    def set_scene_pcd(self, pcd):
        """Extension: hand the scene cloud (M, 3) over directly (frozen-scene runs)."""
        p = L.f32(pcd).reshape(-1, 3)
        self.ctx.call('mh_set_scene', L.ptr(p), p.shape[0], self._stream())
        torch.cuda.current_stream(self.device).synchronize()
        self.scene_pcd = True
        self.scene_depth = True

    # ------------------------------------------------------------------------------------------ outputs
    def _gather_frames(self, local):
        if not self._dist:
            return local
        parts = [None] * self.world
        torch.distributed.all_gather_object(parts, local, group=self.group)
        return np.concatenate([p for p in parts if p is not None and len(p)], axis=0)

    def get_optimized_variables(self):
        """``optimizer.py:619-636`` (same keys and shapes)."""
        ctx, N, T = self.ctx, self.num_people, self.T_local
        xs = ctx.get_param(L.P_XSCALE, (1, N, 1, 1))
        zmin_lin = ctx.get_param(L.P_ZMIN_LIN, (T, 1, 1))
        zmax_lin = ctx.get_param(L.P_ZMAX_LIN, (T, 1, 1))
        min_z = np.log(1.0 + np.exp(zmin_lin)).astype(np.float32)                                  # transforms.py:296
        max_z = (min_z + 1.0 + np.log(1.0 + np.exp(zmax_lin))).astype(np.float32)
        return {
            'scale_factor': np.power(np.float32(1.1), xs).astype(np.float32),
            'poses_T': self._gather_frames(ctx.get_param(L.P_POSES_T, (T, N, 1, 3))),
            'poses_smpl': self._gather_frames(ctx.get_param(L.P_POSES_SMPL, (T, N, 72))),
            'betas_smpl': ctx.get_param(L.P_BETAS, (1, N, 10)),
            'valid_smpl': self.valid_smpl,
            'min_z': self._gather_frames(min_z),
            'max_z': self._gather_frames(max_z),
            'scene_depth': self.scene_depth if isinstance(self.scene_depth, np.ndarray) else None,
            'scene_img': self.scene_img,
            'scene_mask': self.scene_mask,
        }

    def one_euro_filter(self, x, min_cutoff=0.1, beta=0.02, frame_rate=25):
        """``optimizer.py:664-675`` on the device: x (T, ...) -> filtered float32 tensor on ``self.device``."""
        y = L.f32(x.detach().cpu().numpy() if torch.is_tensor(x) else x)
        out = np.empty_like(y)
        T = y.shape[0]
        self.ctx.call('mh_one_euro_filter', L.ptr(y), L.ptr(out), T, y.size // T, float(min_cutoff), float(beta), float(frame_rate))
        return torch.from_numpy(out).to(self.device)

    def smpl_forward(self, betas, poses, want_verts=True):
        """``SMPL.forward`` (``smpl.py:297-399``) through the kernels: (verts (nb,6890,3) | None, joints_alphapose (nb,17,3))."""
        b, p = L.f32(betas).reshape(-1, 10), L.f32(poses).reshape(-1, 72)
        verts = np.empty((b.shape[0], L.V, 3), np.float32) if want_verts else None
        j17 = np.empty((b.shape[0], 17, 3), np.float32)
        self.ctx.call('mh_smpl_forward', L.ptr(b), L.ptr(p), b.shape[0], L.ptr(verts), L.ptr(j17))
        return verts, j17

    def predict(self, poses_T, poses_smpl, betas_smpl, scale_factor):
        """``SMPLOptimizerBase.predict`` (``optimizer.py:133-143``; not called by the reference's drivers): absolute vertices and
        sparse joints of the given bodies, ``scale_factor * SMPL(betas, poses) + poses_T`` with numpy broadcasting as there."""
        if self.ctx is None:
            raise RuntimeError('predict needs the device context: call init_optimized_variables first')
        verts, joints = self.smpl_forward(betas_smpl, poses_smpl)
        return scale_factor * verts + poses_T, scale_factor * joints + poses_T

    def get_filtered_vertices_by_smpl(self, min_cutoff_T=0.004, min_cutoff_angles=0.1, beta_T=0.7, beta_angles=0.1, frame_rate=25):
        """``optimizer.py:639-661`` (not called by the reference's drivers): One-Euro filter of the optimised translations and pose
        angles over time (here with the proper time stamps ``i / frame_rate``, unlike ``one_euro_filter``), then SMPL on the
        filtered poses; absolute vertices (T, N, 6890, 3) as a float32 tensor on ``self.device``.  Single-rank only."""
        if self._dist:
            raise NotImplementedError('get_filtered_vertices_by_smpl: gather the variables with get_optimized_variables() first')
        ctx, N, T = self.ctx, self.num_people, self.T_local
        pT = one_euro_over_time(ctx.get_param(L.P_POSES_T, (T, N, 1, 3)), min_cutoff_T, beta_T, frame_rate)
        th = one_euro_over_time(ctx.get_param(L.P_POSES_SMPL, (T, N, 72)), min_cutoff_angles, beta_angles, frame_rate)
        betas = np.tile(ctx.get_param(L.P_BETAS, (1, N, 10)), (T, 1, 1))
        verts, _ = self.smpl_forward(betas.reshape(-1, 10), th.reshape(-1, 72))
        scale = np.power(np.float32(1.1), ctx.get_param(L.P_XSCALE, (1, N, 1, 1))).astype(np.float32)
        return torch.from_numpy((scale * verts.reshape(T, N, -1, 3) + pT).astype(np.float32)).to(self.device)

    def smpl_regress(self, betas, poses, regressor):
        """SMPL forward + ``regressor (J, 6890) . verts`` on the device (``smpl.py:376-389`` with any of the reference's regressors:
        ``joints_mupots``, ``joints_h36m17``, ``joints_extra9``); local joints (nb, J, 3), used by ``evaluation.SMPLJoints``."""
        b, p = L.f32(betas).reshape(-1, 10), L.f32(poses).reshape(-1, 72)
        R = np.asarray(regressor, np.float32)
        if R.ndim == 2 and R.shape[0] == L.V and R.shape[1] != L.V:            # the reference's .npy files are stored (6890, J)
            R = R.T
        R = L.f32(np.ascontiguousarray(R))
        if R.ndim != 2 or R.shape[1] != L.V:
            raise ValueError(f'regressor must be (J, {L.V}) or ({L.V}, J), got {R.shape}')
        out = np.empty((b.shape[0], R.shape[0], 3), np.float32)
        self.ctx.call('mh_smpl_regress', L.ptr(b), L.ptr(p), b.shape[0], L.ptr(R), R.shape[0], L.ptr(out))
        return out
