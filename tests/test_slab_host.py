"""Host side of the multi-GPU path on CPU: the torch.distributed transport (gloo, world_size 2 and 3) and a CPU model of
the slab algorithm -- own rows + halo rows (ghost arithmetic across the periodic wrap) + local y/z ghosts -- whose
neighbour pairs must be exactly those of the single-domain run (oracle brute force, src/check_neighbourlist.f90:149-173)."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import parity
from ndspmhd_b200 import abi, setups, slab
from oracle import oracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, fn, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def run_ranks(world, fn):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


# ---- transport ----------------------------------------------------------------------------------------------------------
def _transport(rank, world):
    comm = slab.SlabComm(device="cpu")
    out = {}
    out["max"] = comm.allreduce([float(rank), -float(rank)], slab.OP_MAX)
    out["min"] = comm.allreduce([float(rank)], slab.OP_MIN)
    out["sum"] = comm.allreduce([1.0, float(rank)], slab.OP_SUM)
    # ring exchange through the C callbacks with raw host pointers, different sizes per side
    nl, nr = 3 + rank, 5 + 2 * rank
    sl = np.full(nl, 10 * rank + 1, np.uint8)
    sr = np.full(nr, 10 * rank + 2, np.uint8)
    cb_all, cb_cnt, cb_sr = comm.callbacks()
    sb = (C.c_longlong * 2)(nl, nr)
    rb = (C.c_longlong * 2)()
    assert cb_cnt(None, sb, rb) == 0
    rl, rr = np.zeros(rb[0], np.uint8), np.zeros(rb[1], np.uint8)
    sp = (C.c_void_p * 2)(sl.ctypes.data, sr.ctypes.data)
    rp = (C.c_void_p * 2)(rl.ctypes.data, rr.ctypes.data)
    assert cb_sr(None, sp, sb, rp, rb, None) == 0
    out["recv_left"], out["recv_right"] = rl.tolist(), rr.tolist()
    v = (C.c_double * 2)(rank + 1.0, 2.0)
    assert cb_all(None, v, 2, slab.OP_SUM) == 0
    out["cb_sum"] = [v[0], v[1]]
    return out


@pytest.mark.parametrize("world", [2, 3])
def test_transport_callbacks_over_gloo(world):
    res = run_ranks(world, _transport)
    for rank, out in enumerate(res):
        assert out["max"] == [world - 1.0, 0.0] and out["min"] == [0.0]
        assert out["sum"] == [float(world), sum(range(world))]
        left, right = (rank - 1) % world, (rank + 1) % world
        # from the left neighbour I get what it sent to ITS right (value 10*left+2, length 5+2*left), and vice versa
        assert out["recv_left"] == [10 * left + 2] * (5 + 2 * left)
        assert out["recv_right"] == [10 * right + 1] * (3 + right)
        assert out["cb_sum"] == [sum(r + 1.0 for r in range(world)), 2.0 * world]


# ---- partition ----------------------------------------------------------------------------------------------------------
def test_slab_edges_balance_and_ownership():
    rng = np.random.default_rng(3)
    x = rng.random(10001) - 0.5
    for nr in (1, 2, 3, 8):
        e = slab.slab_edges(x, nr, -0.5, 0.5)
        own = slab.owner_of(x, e)
        counts = np.bincount(own, minlength=nr)
        assert counts.sum() == x.size and counts.max() - counts.min() <= 1
        assert np.all(np.diff(e) > 0) and e[0] == -0.5 and e[-1] == 0.5
        assert np.all((x >= e[own]) & (x < e[own + 1]))


# ---- CPU model of the slab algorithm ---------------------------------------------------------------------------------------
def _slab_pairs(rank, world):
    """Each rank: own rows + halos from both neighbours (numpy restatement of k_halo_flags/k_halo_pack1, moved with the
    gloo transport) + local periodic ghosts in y,z by the oracle's set_ghost_particles; then brute-force pairs with an own
    row on at least one side, reported in GLOBAL particle ids (ghost/halo -> parent id)."""
    kw = dict(ndim=3, nx=12, zfrac=0.5, perturb_amp=0.3, evolved=True)
    o, pg = setups.orszag_tang(**kw)
    hmax = float(np.max(pg.hh[: pg.npart]))
    o, pl, info = setups.orszag_tang(slab=(rank, world), **kw)
    n = pl.npart
    lo, hi = info["edges"][rank], info["edges"][rank + 1]
    reach = 2.0 * hmax * (1.0 + 1e-10)
    comm = slab.SlabComm(device="cpu")
    hmax_g = comm.allreduce([float(np.max(pl.hh[:n]))], slab.OP_MAX)[0]
    assert hmax_g == hmax
    left, right = slab.halo_select_numpy(pl.x[:n, 0], lo, hi, reach, True, True)

    def pack(rows, shift):
        x = pl.x[rows].copy()
        if shift == "lo":      # crossing the xmin face: appears beyond xmax on the receiver
            x[:, 0] = slab.wrap_shift(x[:, 0], o.xmin[0], o.xmax[0])
        elif shift == "hi":
            x[:, 0] = slab.wrap_shift(x[:, 0], o.xmax[0], o.xmin[0])
        rec = np.concatenate([x, pl.hh[rows, None], info["rows"][rows, None].astype(np.float64)], axis=1)
        return torch.from_numpy(np.ascontiguousarray(rec).view(np.uint8).reshape(-1))

    sl = pack(left, "lo" if rank == 0 else None)
    sr = pack(right, "hi" if rank == world - 1 else None)
    nl, nr = comm.exchange_counts(sl.numel(), sr.numel())
    rl, rr = torch.zeros(nl, dtype=torch.uint8), torch.zeros(nr, dtype=torch.uint8)
    comm.sendrecv_tensors(sl, sr, rl, rr)
    halo = np.concatenate([rl.numpy().view(np.float64).reshape(-1, 5), rr.numpy().view(np.float64).reshape(-1, 5)], axis=0)
    nh = halo.shape[0]
    # rows [0,n) own, [n,n+nh) halo; then y/z ghosts from all of them, x treated as non-periodic locally
    q = abi.Particles(3, n + nh, 4 * (n + nh) + 64)
    q.x[:n], q.hh[:n] = pl.x[:n], pl.hh[:n]
    q.x[n:n + nh], q.hh[n:n + nh] = halo[:, :3], halo[:, 3]
    q.pmass[: n + nh] = pl.pmass[0]
    gid = np.concatenate([info["rows"], halo[:, 4].astype(np.int64)])
    o2 = abi.NdOptions.from_buffer_copy(o)
    o2.ibound[0] = 0
    o2.device_ghosts = 1
    # set_ghost_particles takes hhmax = max h of its rows; the halo rows carry the neighbours' h, and the global maximum
    # is what the library uses -- give the model the same by planting it on one halo/own row's reach only through options
    s, _ = oracle.derivs(o2, q, phases=oracle.NDO_GHOSTS)
    nt = q.ntotal
    assert s["hhmax"] <= hmax
    par = np.arange(nt)
    par[n + nh:nt] = q.ireal[n + nh:nt] - 1
    q.hh[n + nh:nt] = q.hh[par[n + nh:nt]]
    bi, bj = oracle.bruteforce_pairs(q, 4.0)              # pairs with i or j among rows [0, npart) = own + halo
    keep = (bi <= n) | (bj <= n)                          # at least one OWN row
    gi, gj = gid[par[bi[keep] - 1]], gid[par[bj[keep] - 1]]
    return parity.pair_set(gi + 1, gj + 1).tolist(), float(s["hhmax"]), hmax


@pytest.mark.parametrize("world", [2, 3])
def test_slab_halo_model_reproduces_single_domain_neighbour_pairs(world):
    kw = dict(ndim=3, nx=12, zfrac=0.5, perturb_amp=0.3, evolved=True)
    o, p = setups.orszag_tang(**kw)
    o.device_ghosts = 1
    s, _ = oracle.derivs(o, p, phases=oracle.NDO_GHOSTS)
    n, nt = p.npart, p.ntotal
    p.hh[n:nt] = p.hh[p.ireal[n:nt] - 1]
    bi, bj = oracle.bruteforce_pairs(p, 4.0)
    par = np.arange(nt)
    par[n:nt] = p.ireal[n:nt] - 1
    want = parity.pair_set(par[bi - 1] + 1, par[bj - 1] + 1)
    res = run_ranks(world, _slab_pairs)
    got = np.unique(np.concatenate([np.array(r[0], dtype=np.int64) for r in res]))
    # a slab's local ghost reach uses its own max h (<= global); every rank reports it for the record
    assert np.array_equal(got, want)


# ---- row migration between slabs (ndspmhd_b200_step on slab contexts) ---------------------------------------------------------
def _migrate(rank, world):
    """A CPU model of migrate_rows: rows drift and wrap, the plan (slab.migration_plan_numpy = the device kernels' rule) says who leaves
    where and which tail rows fill the holes, payloads travel over the gloo transport; afterwards every id must be owned once, by the
    rank whose slab holds it, with its payload intact."""
    rng = np.random.default_rng(7)
    nglobal = 4000
    xg = rng.uniform(-0.5, 0.5, nglobal)
    edges = slab.slab_edges(xg, world, -0.5, 0.5)
    rows = np.nonzero(slab.owner_of(xg, edges) == rank)[0]
    x, ids = xg[rows].copy(), rows.astype(np.int64)
    payload = np.stack([x, ids * 3.25], axis=1)
    comm = slab.SlabComm(device="cpu")
    width = float(np.min(np.diff(edges)))
    moved = 0
    for step in range(4):
        drift = np.random.default_rng(100 + step).uniform(-0.4, 0.4, nglobal)[ids] * width     # same draw for an id on every rank
        x = x + drift
        x = np.where(x > 0.5, -0.5 + x - 0.5, np.where(x < -0.5, 0.5 - (-0.5 - x), x))           # boundaryND.f90:65-93
        payload[:, 0] = x
        li, ri, moves, m = slab.migration_plan_numpy(x, edges, rank, periodic=True)
        sl = np.ascontiguousarray(np.concatenate([payload[li], ids[li, None].astype(np.float64)], axis=1)).view(np.uint8).reshape(-1)
        sr = np.ascontiguousarray(np.concatenate([payload[ri], ids[ri, None].astype(np.float64)], axis=1)).view(np.uint8).reshape(-1)
        nl, nr_ = comm.exchange_counts(sl.size, sr.size)
        rl, rr = np.zeros(nl, np.uint8), np.zeros(nr_, np.uint8)
        comm.sendrecv_tensors(torch.from_numpy(sl) if sl.size else None, torch.from_numpy(sr) if sr.size else None,
                              torch.from_numpy(rl) if nl else None, torch.from_numpy(rr) if nr_ else None)
        payload[moves[:, 1]] = payload[moves[:, 0]]
        ids[moves[:, 1]] = ids[moves[:, 0]]
        arr = np.concatenate([rl.view(np.float64).reshape(-1, 3), rr.view(np.float64).reshape(-1, 3)], axis=0)
        payload = np.concatenate([payload[:m], arr[:, :2]], axis=0)
        ids = np.concatenate([ids[:m], arr[:, 2].astype(np.int64)])
        x = payload[:, 0].copy()
        moved += li.size + ri.size
    inside = (x >= edges[rank]) & ((x < edges[rank + 1]) | ((rank == world - 1) & (x == edges[rank + 1])))
    return {"ids": ids.tolist(), "all_inside": bool(inside.all()), "payload_ok": bool(np.array_equal(payload[:, 1], ids * 3.25)), "moved": int(moved)}


@pytest.mark.parametrize("world", [2, 3])
def test_row_migration_plan_over_gloo(world):
    res = run_ranks(world, _migrate)
    allids = sorted(i for out in res for i in out["ids"])
    assert allids == list(range(4000))                      # every particle owned exactly once
    assert all(out["all_inside"] and out["payload_ok"] for out in res)
    assert sum(out["moved"] for out in res) > 500           # the case does move rows


def test_weak_scaling_setup_makes_one_field_period_per_rank():
    """setups.orszag_tang(weak=True): every rank makes only its own x-period of the box; together they are a valid periodic particle set --
    disjoint slabs with faces on the period boundaries, equal masses and h, the fields of period k equal to those of period 0."""
    world = 3
    parts = [setups.orszag_tang(ndim=3, nx=12, zfrac=0.25, perturb_amp=0.2, evolved=True, slab=(r, world), weak=True) for r in range(world)]
    o0, p0, i0 = parts[0]
    assert i0["nglobal"] == world * p0.npart and o0.xmax[0] - o0.xmin[0] == float(world)
    single_o, single_p = setups.orszag_tang(ndim=3, nx=12, zfrac=0.25, perturb_amp=0.2, evolved=True)
    assert np.isclose(p0.pmass[0], single_p.pmass[0], rtol=1e-14) and p0.hh[0] == single_p.hh[0]     # same resolution as the one-period box
    for r, (o, p, info) in enumerate(parts):
        x = p.x[: p.npart, 0]
        assert np.all(x >= info["edges"][r]) and np.all(x < info["edges"][r + 1])
        assert np.array_equal(info["edges"], o0.xmin[0] + np.arange(world + 1))
        # the lattice of period r is that of period 0 shifted by r (the perturbation differs: its seed depends on the rank)
        assert np.allclose(np.sort(np.round((x - r + 0.5) * 12 - 0.5)), np.sort(np.round((parts[0][1].x[: p0.npart, 0] + 0.5) * 12 - 0.5)))
        # fields have period 1 in x: v_y = sin(2 pi (x - xmin))
        assert np.allclose(p.vel[: p.npart, 1], np.sin(2.0 * np.pi * (x - o.xmin[0])), atol=1e-12)
