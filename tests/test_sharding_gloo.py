"""The per-cycle exchanges of the frame-sharded fit on a 2-rank gloo group (CPU tensors): halo frames, shared-gradient
all-reduce, One-Euro carry hand-over.  On the GPUs the same functions run over NCCL on the library's device buffers."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, T, B, out):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    pkg = ge.load_package()
    sh = sys.modules[pkg.__name__ + '.sharding']
    dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)
    try:
        ranges = [sh.frame_range(T, B, r, world) for r in range(world)]
        t0, t1 = ranges[rank]
        prev, nxt = sh.neighbours(rank, world, ranges)
        # halo: row 0 = my first frame, row 1 = my last frame (payload = the global frame index)
        n = 6
        send = torch.stack([torch.full((n,), float(t0)), torch.full((n,), float(t1 - 1))])
        recv = torch.full((2, n), -1.0)
        hp, hn = sh.exchange_halo(send, recv, prev, nxt)
        # shared block: every rank contributes its frame count
        shared = torch.tensor([float(t1 - t0), 1.0])
        sh.allreduce_shared(shared)
        # sequential carry: running sum of the frame indices, handed rank to rank
        carry_in = torch.zeros(1)
        sh.pass_carry(None, carry_in, prev, nxt)
        carry_out = carry_in + float(sum(range(t0, t1)))
        sh.send_carry(carry_out, nxt)
        out.put((rank, hp, hn, recv.numpy().copy(), shared.numpy().copy(), float(carry_out), (t0, t1)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('T,B', [(16, 4), (7, 2)])
def test_two_rank_exchanges(T, B):
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, T, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r = q.get(timeout=120)
        res[r[0]] = r
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (t00, t01), (t10, t11) = res[0][6], res[1][6]
    assert t00 == 0 and t01 == t10 and t11 == T and t01 % B == 0
    assert (res[0][1], res[0][2]) == (False, True) and (res[1][1], res[1][2]) == (True, False)
    assert np.all(res[0][3][1] == t10) and np.all(res[0][3][0] == -1)        # rank 0 got rank 1's FIRST frame as its "next" halo
    assert np.all(res[1][3][0] == t01 - 1) and np.all(res[1][3][1] == -1)    # rank 1 got rank 0's LAST frame as its "prev" halo
    for r in (0, 1):
        assert np.array_equal(res[r][4], [T, 2])
    assert res[1][5] == sum(range(T))


def _subgroup_worker(rank, world, port, out):
    """World of 3 processes; the job runs on the sub-group of GLOBAL ranks {0, 2}: group-local ranks 0 and 1."""
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    pkg = ge.load_package()
    sh = sys.modules[pkg.__name__ + '.sharding']
    dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)
    try:
        members = [0, 2]
        grp = dist.new_group(members)
        if rank in members:
            gr, gw = dist.get_rank(grp), dist.get_world_size(grp)
            T, B = 8, 2
            ranges = [sh.frame_range(T, B, r, gw) for r in range(gw)]
            t0, t1 = ranges[gr]
            prev, nxt = sh.neighbours(gr, gw, ranges)                      # GROUP-local neighbours
            send = torch.stack([torch.full((4,), float(t0)), torch.full((4,), float(t1 - 1))])
            recv = torch.full((2, 4), -1.0)
            sh.exchange_halo(send, recv, prev, nxt, grp)
            carry_in = torch.zeros(1)
            sh.pass_carry(None, carry_in, prev, nxt, grp)
            carry_out = carry_in + float(t1 - t0)
            sh.send_carry(carry_out, nxt, grp)
            shared = torch.tensor([1.0])
            sh.allreduce_shared(shared, grp)
            out.put((rank, gr, recv.numpy().copy(), float(carry_out), float(shared), (t0, t1)))
        else:
            out.put((rank, -1, None, 0.0, 0.0, (0, 0)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_exchanges_on_a_subgroup_address_global_ranks():
    """prev / next are ranks WITHIN the process group; the point-to-point calls need GLOBAL ranks: with the group
    {0, 2} of a 3-process world, group rank 1 is global rank 2."""
    world = 3
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_subgroup_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r = q.get(timeout=120)
        res[r[0]] = r
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == 0 and res[2][1] == 1 and res[1][1] == -1
    (a0, a1), (b0, b1) = res[0][5], res[2][5]
    assert a0 == 0 and a1 == b0 == 4 and b1 == 8
    assert np.all(res[0][2][1] == b0) and np.all(res[0][2][0] == -1)          # global rank 0 <- first frame of global rank 2
    assert np.all(res[2][2][0] == a1 - 1) and np.all(res[2][2][1] == -1)
    assert res[2][3] == 8.0 and res[0][4] == 2.0 and res[2][4] == 2.0
