"""Restatement of ``SMPLDepthSequenceOptimizer`` (``mhmocap/optimizer.py``) on torch-CPU autograd.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): the teacher-forced parity
checker (same parameters in -> same losses / gradients out) and the "port" timed
as the CPU baseline.  Written from the reference's algorithm; each block cites
the ``optimizer.py`` lines it follows, including the batch-partition quirks
Q1-Q4 of SURVEY.md section 8a.  The frame batches are contiguous index segments
(``shuffle=False`` semantics).
"""
import numpy as np
import torch

from . import refmath as rm
from . import raster
from . import scene_ref

COEF_KEYS = ['proj2d', 'depth', 'silhouette', 'reg_velocity', 'reg_verts_filter', 'reg_poses',
             'reg_scales', 'reg_contact', 'reg_foot_sliding']


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


class FitRef(object):
    def __init__(self, model, image_size, num_frames, cam_K, coefs, cam_dist_coef=None,
                 joint_confidence_thr=0.5, eps=1e-3, znear=1.0, zfar=100.0, pose17j_weights=None):
        """model: dict from ``oracle.synth.load_model_tensors``; image_size (W,H);
        coefs: dict over COEF_KEYS (``optimizer.py:159-167``)."""
        self.m = {k: (_t(v) if isinstance(v, np.ndarray) and v.dtype == np.float32 else v) for k, v in model.items()}
        self.m['parents'] = [int(p) for p in model['parents']]
        self.faces = _t(model['faces'].astype(np.int64))
        self.W, self.H = image_size
        self.T = num_frames
        self.cam_K = np.asarray(cam_K, np.float32)
        self.Kd = cam_dist_coef
        self.coefs = dict(coefs)
        self.thr = joint_confidence_thr
        w17 = np.ones(17, np.float32) if pose17j_weights is None else np.asarray(pose17j_weights, np.float32)
        self.pose_weights = _t((len(w17) * w17 / np.sum(w17)).astype(np.float32)).view(1, 1, 17, 1)      # optimizer.py:127-130, 259
        self.eps = eps
        self.Kndc = _t(rm.compute_calibration_matrix(znear, zfar, self.cam_K, image_size))   # :206
        self.scene_pcd = None
        self.scene_depth = None
        self.verts_filtered = None
        self.poses_T_filtered = None

    # ------------------------------------------------------------------ SMPL
    def _smpl(self, betas, poses):
        out = rm.smpl_forward(self.m, betas, poses)
        j17 = rm.regress_joints(self.m['J_regressor_alphapose'], out['verts'])          # smpl.py:375-377
        return out['verts'], j17

    # ------------------------------------------------------------------ init (hot loop A)
    def init_optimized_variables(self, pose2d, poses_smpl, betas_smpl, valid_smpl, scale_factor=None,
                                 num_iter=100, joints_thr=0.15):
        """``optimizer.py:262-321`` + ``__init_global_poses`` (``:710-770``)."""
        T, N = pose2d.shape[:2]
        self.N = N
        if scale_factor is not None:
            xs = np.log(scale_factor) / np.log(1.1)
            self.xscale = _t(xs[None, :, None, None].astype(np.float32))
            self.optim_scale = False
        else:
            self.xscale = torch.zeros(1, N, 1, 1, requires_grad=True)
            self.optim_scale = True
        poses_T = torch.tensor(np.tile(np.array([[[[0, 0, 1]]]], np.float32), (T, N, 1, 1)), requires_grad=True)
        th = _t(poses_smpl.astype(np.float32)); be = _t(betas_smpl.astype(np.float32))
        vis = _t((pose2d[..., 2:] > joints_thr).astype(np.float32))
        gt = _t(pose2d[..., 0:2].astype(np.float32))
        opt = torch.optim.Adam([poses_T], lr=0.5, betas=(0.5, 0.5), eps=1e-6)
        sch = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.95)
        Kt = _t(self.cam_K)[None].expand(T * N, 3, 3)
        log = []
        with torch.no_grad():
            _, j17 = self._smpl(be.view(T * N, -1), th.view(T * N, -1))          # inputs constant over iterations
        for _ in range(num_iter):
            opt.zero_grad()
            scale = torch.pow(1.1, self.xscale)
            g3d = scale * j17.view(T, N, -1, 3) + poses_T
            p2d = rm.camera_projection(g3d.view(T * N, -1, 3), Kt, self.Kd).view(T, N, -1, 2)
            loss_2d = torch.mean(torch.square(self.pose_weights * vis * p2d - self.pose_weights * vis * gt))   # MSELoss(mean) :740, 754-756
            speed = torch.sum(torch.square(poses_T[1:] - poses_T[:-1]))
            loss = self.coefs['proj2d'] * loss_2d + self.coefs['reg_velocity'] * speed
            log.append({'loss_2d': loss_2d.detach().numpy()})
            loss.backward()
            opt.step(); sch.step()
        pT = poses_T.detach().numpy()
        self.set_variables(pT, poses_smpl, np.mean(betas_smpl, axis=0, keepdims=True), valid_smpl)
        return log

    def set_variables(self, poses_T, poses_smpl, betas, valid_smpl, zmin_lin=None, zmax_lin=None, xscale=None):
        """Leaves of ``fit`` (``optimizer.py:291-303``)."""
        self.N = poses_T.shape[1]
        self.poses_T = torch.tensor(poses_T.astype(np.float32), requires_grad=True)
        self.poses_smpl = torch.tensor(poses_smpl.astype(np.float32), requires_grad=True)
        self.betas = torch.tensor(betas.astype(np.float32), requires_grad=True)
        self.betas_ref = torch.tensor(betas.astype(np.float32))
        self.valid_smpl = _t((valid_smpl > 0.7).astype(np.float32))
        max_z = np.clip(np.max(poses_T[..., 2:], axis=1), 2, None)
        self.zmin_lin = torch.tensor(np.ones_like(max_z) if zmin_lin is None else zmin_lin, requires_grad=True)
        self.zmax_lin = torch.tensor(2.0 * max_z if zmax_lin is None else zmax_lin, requires_grad=True)
        if xscale is not None:
            self.xscale = torch.tensor(xscale.astype(np.float32), requires_grad=True)
            self.optim_scale = True
        elif not hasattr(self, 'xscale'):
            self.xscale = torch.zeros(1, self.N, 1, 1, requires_grad=True)
            self.optim_scale = True

    def leaves(self):
        lv = [self.poses_T, self.poses_smpl, self.betas, self.zmin_lin, self.zmax_lin]
        if self.optim_scale:
            lv.append(self.xscale)
        return lv

    # ------------------------------------------------------------------ one batch
    def batch_loss(self, data, idxs):
        """Loss of one frame batch (``optimizer.py:394-544``).  ``data``: dict of
        full-sequence numpy arrays; ``idxs``: int array of the frames in this batch."""
        B, N, H, W = len(idxs), self.N, self.H, self.W
        ix = torch.from_numpy(np.asarray(idxs, np.int64))
        scale = torch.pow(1.1, self.xscale)                                              # :681
        min_z = rm.softplus(self.zmin_lin[ix])                                           # :683
        max_z = min_z.detach().clone() + 1.0 + rm.softplus(self.zmax_lin[ix])            # :684-688
        poses = self.poses_smpl[ix].view(-1, 72)
        betas = self.betas.tile((B, 1, 1)).view(-1, 10)
        verts, j17 = self._smpl(betas, poses)
        pT = self.poses_T[ix]
        verts_abs = scale * verts.view(B, N, -1, 3) + pT                                 # :702
        joints_abs = scale * j17.view(B, N, -1, 3) + pT                                  # :703
        pose2d = _t(data['pose2d'][idxs]); seg = _t(data['seg_mask'][idxs]); depths = _t(data['depths'][idxs])
        thr_scores = torch.ge(pose2d[..., 2:3], self.thr).float()                        # :404
        pose2d_valid = torch.ge(torch.sum(thr_scores, dim=(2, 3)), 2).float()            # :405
        smpl_valid = self.valid_smpl[ix].float()
        mask_valid = torch.ge(torch.sum(seg, dim=(2, 3)), 0.005 * H * W).float()         # :407-409
        # 2D term :414-420
        Kt = _t(self.cam_K)[None].expand(B * N, 3, 3)
        j2d = rm.camera_projection(joints_abs.view(B * N, -1, 3), Kt, self.Kd).view(B, N, -1, 2)
        norm = torch.tensor([[[[float(W), float(H)]]]])
        pmask = self.pose_weights * thr_scores                                            # :420
        loss_pose = torch.sum(torch.square(pmask * j2d / norm - pmask * pose2d[..., 0:2] / norm))
        # raster terms :425-475
        target_disp = depths * (1.0 / min_z - 1.0 / max_z) + 1.0 / max_z
        zb, al = [], []
        for b in range(B):
            for n in range(N):
                z0, a = raster.render_person(verts_abs[b, n], self.faces, self.Kndc, H, W)
                zb.append(z0); al.append(a)
        zbuf = torch.stack(zb).view(B, N, H, W)
        alpha = torch.stack(al).view(B, N, H, W)
        eroded = rm.erode5_twice3(seg.view(B * N, 1, H, W)).view(B, N, H, W)
        sup = torch.gt(zbuf, 0).float() * eroded * pose2d_valid.unsqueeze(-1).unsqueeze(-1)   # :432-438
        zdisp = 1.0 / torch.clamp(zbuf + 0.2, self.eps)                                  # :440
        loss_depth = rm.avg_depth_loss(zdisp, target_disp.unsqueeze(1), sup)             # :442
        order = torch.argsort(pT[..., 0, 2], dim=1)                                      # :450
        loss_sil = 0
        for j in range(B):
            acc = torch.zeros(H, W)
            for q in range(N):
                p = int(order[j, q])
                if float(mask_valid[j, q] * pose2d_valid[j, q]) > 0:                     # gate indexed by POSITION q (:472)
                    loss_sil = loss_sil + rm.masked_mse_loss(alpha[j, p], seg[j, p], 1 - acc)
                acc = torch.gt(acc + seg[j, p], 0).float()                               # :475
        # contact + foot sliding :483-518
        reg_contact = 0
        reg_foot = 0
        if self.scene_pcd is not None:
            gv = verts_abs
            low_idx = torch.argmax(gv[..., 1:2], dim=2, keepdim=True).tile((1, 1, 1, 3)).long()
            low = torch.gather(gv, 2, low_idx)                                           # (B,N,1,3)
            d2 = torch.sum(torch.pow(self.scene_pcd - low, 2), -1)                       # (B,N,M)
            nn = torch.topk(d2, 32, dim=-1, largest=False).indices                       # set of 32 nearest (:495)
            pts = self.scene_pcd[0, 0][nn]                                               # (B,N,32,3)
            mean_pt = torch.mean(pts, dim=2, keepdim=True)
            cdv = (mean_pt - low)[..., 1:2]
            target = pT.detach().clone()
            target[..., 1:2] += cdv + 0.02
            reg_contact = torch.sum(torch.abs(pT - target.detach().clone()))             # :506
            in_thr = torch.gt(cdv, -0.20)
            low_t = low[1:]; in_t = in_thr[1:]
            low_tm1 = torch.gather(gv[:-1], 2, low_idx[1:])
            reg_foot = torch.sum(torch.abs(in_t * low_t - in_t * low_tm1)) / torch.clamp(torch.sum(in_t), 1)
        # priors :523-532
        ref = _t(data['poses_smpl'][idxs])
        reg_poses = torch.sum(torch.abs(smpl_valid * ref - smpl_valid * self.poses_smpl[ix]))
        reg_poses = reg_poses + B * torch.sum(torch.abs(self.betas - self.betas_ref))
        reg_scale_avg = torch.square(torch.sum(scale - 1.0))
        reg_scale_person = torch.mean(torch.square(scale - 1.0))
        c = self.coefs
        total = (c['proj2d'] * loss_pose + c['depth'] * loss_depth + c['silhouette'] * loss_sil
                 + c['reg_poses'] * reg_poses + c['reg_scales'] * reg_scale_person
                 + float(c['reg_scales'] > 0) * reg_scale_avg
                 + c['reg_contact'] * reg_contact + c['reg_foot_sliding'] * reg_foot)    # :535-542
        f = lambda v: float(v) if not isinstance(v, int) else 0.0
        log = {'loss_pose24j': f(loss_pose), 'loss_depth': f(loss_depth), 'loss_silhouette': f(loss_sil),
               'reg_ref_poses': f(reg_poses), 'reg_scale': f(reg_scale_avg + reg_scale_person),
               'reg_contact': f(reg_contact), 'reg_foot_sliding': f(reg_foot)}
        return total, log, target_disp.detach()

    # ------------------------------------------------------------------ one cycle
    def full_sequence_verts(self):
        verts, _ = self._smpl(self.betas.tile((self.T, 1, 1)).view(-1, 10), self.poses_smpl.view(-1, 72))
        return torch.pow(1.1, self.xscale) * verts.view(self.T, self.N, -1, 3) + self.poses_T

    def cycle_grads(self, data, batches):
        """zero_grad + all batch backwards + temporal backward (``optimizer.py:376-575``
        minus the filter refresh and the scene update).  Returns (log dict, list of
        per-batch target_disp).  Gradients are left in ``.grad`` of ``leaves()``."""
        for p in self.leaves():
            p.grad = None
        logs, tds = [], []
        for idxs in batches:
            total, log, td = self.batch_loss(data, idxs)
            total.backward()
            logs.append(log); tds.append(td)
        reg_vel = torch.sum(torch.square(self.poses_T[1:] - self.poses_T[:-1]))           # :560
        loss_t = self.coefs['reg_velocity'] * reg_vel
        reg_fv = 0
        if self.verts_filtered is not None and self.poses_T_filtered is not None:
            gv = self.full_sequence_verts()
            reg_fv = torch.sum(torch.square((gv[1:] - gv[:-1]) - (self.verts_filtered[1:] - self.verts_filtered[:-1])))
            loss_t = loss_t + self.coefs['reg_verts_filter'] * reg_fv                    # :571-574
        loss_t.backward()
        out = {k: float(np.mean([l[k] for l in logs])) for k in logs[0]}                 # Q4 :588-591
        out['reg_vel'] = float(reg_vel)
        out['reg_filter_verts'] = float(reg_fv) if not isinstance(reg_fv, int) else 0.0
        return out, tds

    def refresh_filters(self, min_cutoff1=0.01, beta1=0.02, min_cutoff2=0.001, beta2=0.5):
        """``optimizer.py:383-392``."""
        with torch.no_grad():
            self.poses_T_filtered = _t(rm.one_euro_filter_sequence(self.poses_T.detach().numpy(), min_cutoff1, beta1))
            gv = self.full_sequence_verts().detach().numpy()
            self.verts_filtered = _t(rm.one_euro_filter_sequence(gv, min_cutoff2, beta2))

    def update_scene_pointcloud(self, scene_depth, scene_mask):
        """``optimizer.py:605-616``: inverse-project the pixel centres, keep mask > 0.5."""
        H, W = self.H, self.W
        gx, gy = np.meshgrid(np.linspace(0.5, W - 0.5, W), np.linspace(0.5, H - 0.5, H), indexing='xy')
        uvd = np.stack([gx, gy, scene_depth], axis=-1).astype(np.float32).reshape(1, -1, 3)
        pcd = rm.camera_inverse_projection(_t(uvd), _t(self.cam_K)[None])[0]
        keep = _t(np.asarray(scene_mask).reshape(-1).astype(np.float32)) > 0.5
        self.scene_pcd = pcd[keep][None, None]
        self.scene_depth = scene_depth

    def set_scene_pcd(self, pcd):
        self.scene_pcd = _t(np.asarray(pcd, np.float32))[None, None]
        self.scene_depth = True

    def fit(self, data, batches, num_iter=250, update_filters_every=25, update_scene=True, grad_hook=None):
        """``optimizer.py:324-602`` with RMSprop(lr .01, alpha .5, momentum .9) + ExpLR .99."""
        opt = torch.optim.RMSprop(self.leaves(), lr=0.01, alpha=0.5, momentum=0.9)
        sch = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.99)
        logs = []
        for cycle in range(num_iter):
            if cycle >= 30 and cycle % update_filters_every == 0:
                self.refresh_filters()
            log, tds = self.cycle_grads(data, batches)
            if grad_hook is not None:
                grad_hook(cycle, self)
            if cycle >= 30 and update_scene:
                depths = np.concatenate([(1.0 / td).numpy() for td in tds], axis=0)
                order = np.concatenate(batches)
                img, dep, msk = scene_ref.aggregate_scene_median(depths, data['images'][order],
                                                                 data['backmasks'][order] / 1.0)
                sd = scene_ref.postprocess_depthmap(dep, msk, use_bilateral_filter=True)
                self.update_scene_pointcloud(sd, msk)
                self._ma = (img, msk)
            opt.step(); sch.step()
            logs.append(log)
        return logs

    def variables(self):
        with torch.no_grad():
            min_z = rm.softplus(self.zmin_lin)
            max_z = min_z + 1.0 + rm.softplus(self.zmax_lin)
            return {'scale_factor': torch.pow(1.1, self.xscale).numpy(), 'poses_T': self.poses_T.numpy().copy(),
                    'poses_smpl': self.poses_smpl.numpy().copy(), 'betas_smpl': self.betas.numpy().copy(),
                    'valid_smpl': self.valid_smpl.numpy().copy(), 'min_z': min_z.numpy(), 'max_z': max_z.numpy()}
