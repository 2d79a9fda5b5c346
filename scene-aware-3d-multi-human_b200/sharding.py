"""Frame sharding across ranks and the small per-cycle exchanges (host-side logic, backend-agnostic).

The reference is single-process (``mhmocap/predict.py:267-271``); sharding is this build's addition
(SURVEY.md section 8e).  Frames are split in CONTIGUOUS ranges whose inner edges are multiples of the batch
size B, because the foot-sliding term pairs adjacent entries of one batch only (``optimizer.py:512-518``) and
the per-batch priors (``:526, 531-542``) count whole batches.  Per cycle the ranks exchange
  * one halo frame with each neighbour (theta, T of the boundary frame) for the velocity and filtered-vertex
    terms (``optimizer.py:560-575``), and
  * an all-reduce(sum) of the gradients of the shared leaves (betas, xscale) plus the loss block.
Everything here works on any ``torch.distributed`` backend (NCCL on the GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def frame_range(T_total, B, rank, world):
    """Contiguous [t0, t1) of ``rank``: whole batches, as even as possible, the remainder batch goes last."""
    nb = (T_total + B - 1) // B                       # batches of the whole sequence (the last may be short)
    base, extra = divmod(nb, world)
    b0 = rank * base + min(rank, extra)
    b1 = b0 + base + (1 if rank < extra else 0)
    return min(b0 * B, T_total), min(b1 * B, T_total)


def neighbours(rank, world, ranges):
    """(prev, next) ranks owning the frames adjacent to this rank's range, skipping empty shards; None at the ends."""
    prev = nxt = None
    for r in range(rank - 1, -1, -1):
        if ranges[r][1] > ranges[r][0]:
            prev = r
            break
    for r in range(rank + 1, world):
        if ranges[r][1] > ranges[r][0]:
            nxt = r
            break
    return prev, nxt


def _peer(group, r):
    """``prev`` / ``next`` are ranks WITHIN ``group`` (they index ``ranges``); point-to-point calls address GLOBAL ranks."""
    if r is None or group is None:
        return r
    return dist.get_global_rank(group, r)


def exchange_halo(send, recv, prev, nxt, group=None):
    """send / recv: (2, n) tensors -- row 0 = this rank's FIRST frame, row 1 = its LAST frame; after the call
    recv[0] = prev rank's last frame, recv[1] = next rank's first frame.  Returns (has_prev, has_next)."""
    ops = []
    gp, gn = _peer(group, prev), _peer(group, nxt)
    if prev is not None:
        ops.append(dist.P2POp(dist.isend, send[0], gp, group=group))
        ops.append(dist.P2POp(dist.irecv, recv[0], gp, group=group))
    if nxt is not None:
        ops.append(dist.P2POp(dist.isend, send[1], gn, group=group))
        ops.append(dist.P2POp(dist.irecv, recv[1], gn, group=group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return prev is not None, nxt is not None


def allreduce_shared(shared, group=None):
    """Sum the shared-leaf gradients + loss block over the ranks (every rank then applies the same update)."""
    dist.all_reduce(shared, op=dist.ReduceOp.SUM, group=group)
    return shared


def pass_carry(carry_out, carry_in, prev, nxt, group=None):
    """Sequential One-Euro hand-over: receive the filter state from ``prev`` (blocking), to be called BEFORE this
    rank's scan; ``send_carry`` is called after it."""
    if prev is not None:
        dist.recv(carry_in, src=_peer(group, prev), group=group)


def send_carry(carry_out, nxt, group=None):
    if nxt is not None:
        dist.send(carry_out, dst=_peer(group, nxt), group=group)


def log_from_loss_block(L, n_batches_total):
    """The reference's ``optim_log`` entry of one cycle (``optimizer.py:546-554, 588-593``) from the 16-float loss
    block of GLOBAL sums: the seven per-batch terms are batch MEANS (Q4), the two temporal terms are plain sums."""
    nb = float(n_batches_total)
    return {
        'loss_pose24j': float(L[0]) / nb, 'loss_depth': float(L[1]) / nb, 'loss_silhouette': float(L[2]) / nb,
        'reg_ref_poses': float(L[3]) / nb, 'reg_scale': float(L[4]) / nb, 'reg_contact': float(L[5]) / nb,
        'reg_foot_sliding': float(L[6]) / nb, 'reg_vel': float(L[7]), 'reg_filter_verts': float(L[8]),
    }
