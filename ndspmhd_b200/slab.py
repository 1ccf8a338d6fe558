"""Multi-GPU host side: x-slab decomposition, one rank (process) per GPU.

The reference is serial (docs/about.rst:12); its only notion of "image particles" is the periodic ghost of
src/ghostND_mhd.f90:166-346.  The library (csrc/nd_host.cuh: halo_exchange_inputs / halo_exchange_density / migrate_rows) applies the
same rule at slab faces and owns selection, packing, row layout and -- when ndspmhd_b200_step runs on a slab context -- the migration
of rows whose x left the slab.  Two transports carry its traffic:

  * native (attach_nccl): the library calls NCCL itself on its compute stream (csrc/nd_nccl.cuh); this module only broadcasts the
    128-byte communicator id.  What bench.py --gpus N uses.
  * callbacks (SlabComm + attach): the three `nd_comm` callbacks of include/ndspmhd_b200.h over torch.distributed --

        allreduce(v[n], op)            small host-side all-reduces (hhmax, unconverged count, stressmax, vsigmax, dt's)
        sendrecv_counts(send[2])       byte-count handshake with the two x-neighbours (periodic ring)
        sendrecv(sendbuf[2], recvbuf[2])   the halo / migration payloads, device buffers, NCCL send/recv over NVLink

    on CUDA tensors with the NCCL backend and on CPU tensors with gloo (tests/test_slab_host.py runs the latter with world_size 2
    and 3): what an MPI host would supply.

slab_edges / owner_of / take_rows partition a global particle set; halo_select_numpy, wrap_shift and migration_plan_numpy restate the
device kernels' rules in numpy for the CPU tests.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi
from .abi import NdOptions, Particles

OP_MAX, OP_MIN, OP_SUM = 0, 1, 2

_ALLREDUCE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_int, C.c_int)
_COUNTS = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong))
_SENDRECV = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_longlong), C.POINTER(C.c_void_p),
                        C.POINTER(C.c_longlong), C.c_void_p)


class NdComm(C.Structure):
    """ctypes mirror of `nd_comm` (include/ndspmhd_b200.h)."""

    _fields_ = [("user", C.c_void_p), ("rank", C.c_int), ("nranks", C.c_int), ("slab_lo", C.c_double), ("slab_hi", C.c_double),
                ("nglobal", C.c_longlong), ("allreduce", _ALLREDUCE), ("sendrecv_counts", _COUNTS), ("sendrecv", _SENDRECV)]


# ---------------------------------------------------------------------------------------------------------------------
# partition
# ---------------------------------------------------------------------------------------------------------------------
def slab_edges(x0: np.ndarray, nranks: int, xmin: float, xmax: float) -> np.ndarray:
    """Slab faces along x with equal particle counts: faces sit half-way between the two particles either side of each
    quantile, so no particle lies on a face.  Returns nranks+1 edges with edges[0] = xmin, edges[-1] = xmax."""
    n = x0.shape[0]
    edges = np.empty(nranks + 1)
    edges[0], edges[-1] = xmin, xmax
    if nranks > 1:
        xs = np.sort(x0)
        for r in range(1, nranks):
            k = (n * r) // nranks
            edges[r] = 0.5 * (xs[k - 1] + xs[k]) if 0 < k < n else xmin + (xmax - xmin) * r / nranks
            if not (xs[k - 1] < edges[r] <= xs[k]) and 0 < k < n:   # duplicate coordinates at the quantile: fall back to xs[k]
                edges[r] = xs[k]
    return edges


def owner_of(x0: np.ndarray, edges: np.ndarray) -> np.ndarray:
    """Rank owning each particle: edges[r] <= x < edges[r+1] (the last slab is closed at xmax)."""
    r = np.searchsorted(edges, x0, side="right") - 1
    return np.clip(r, 0, len(edges) - 2).astype(np.int32)


def take_rows(p: Particles, rows: np.ndarray, extra: int = 0) -> Particles:
    """A rank's own particles, in global index order, with room for `extra` more rows."""
    n = int(rows.shape[0])
    q = Particles(p.ndim, n, n + extra)
    for k, v in p.arrays.items():
        q.arrays[k][:n] = v[rows]
    q.ntotal = n
    return q


def halo_select_numpy(x0: np.ndarray, lo: float, hi: float, reach: float, left_on: bool, right_on: bool):
    """numpy restatement of k_halo_flags: rows sent to the -x and +x neighbours."""
    left = np.nonzero(left_on & (x0 < lo + reach))[0]
    right = np.nonzero(right_on & (x0 > hi - reach))[0]
    return left, right


def wrap_shift(x: np.ndarray, xbound: float, xperbound: float) -> np.ndarray:
    """src/ghostND_mhd.f90:212-225: xnew = xperbound + (x - xbound), in that order of operations."""
    return xperbound + (x - xbound)


def migration_plan_numpy(x0: np.ndarray, edges: np.ndarray, rank: int, periodic: bool):
    """numpy restatement of k_migrate_flags + the hole filling of migrate_rows (csrc/nd_host.cuh): which own rows leave to the left and
    right neighbour after the integrator moved them (x already wrapped into the box, src/boundaryND.f90:65-93), and the moves
    (src row -> dst row) that fill the holes below m = nown - nleave from the staying rows of the tail, so that rows [0,m) stay and
    arrivals are appended at m.  Ownership: edges[r] <= x < edges[r+1], the last slab closed at xmax.  A row that is in neither
    adjacent slab raises (it moved farther than one slab in a step)."""
    nr = len(edges) - 1
    left, right = (rank - 1) % nr, (rank + 1) % nr

    def inside(x, r):
        return (x >= edges[r]) & ((x < edges[r + 1]) | ((r == nr - 1) & (x == edges[r + 1])))

    away = ~inside(x0, rank)
    to_right = away & inside(x0, right) & (periodic or rank < nr - 1)
    to_left = away & ~to_right & inside(x0, left) & (periodic or rank > 0)
    if left == right:                       # ring of two: one peer, everything goes out on the right side
        to_right, to_left = to_right | to_left, np.zeros_like(to_left)
    if np.any(away & ~to_right & ~to_left):
        raise ValueError("a row left its slab by more than the adjacent slab")
    holes = np.nonzero(away)[0]
    m = x0.shape[0] - holes.shape[0]
    tail_stay = m + np.nonzero(~away[m:])[0]
    moves = np.stack([tail_stay, holes[: tail_stay.shape[0]]], axis=1) if tail_stay.shape[0] else np.zeros((0, 2), np.int64)
    return np.nonzero(to_left)[0], np.nonzero(to_right)[0], moves, m


# ---------------------------------------------------------------------------------------------------------------------
# transport
# ---------------------------------------------------------------------------------------------------------------------
class _DevPtr:
    """Wraps a raw device pointer so torch.as_tensor can view it (CUDA array interface v2)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class SlabComm:
    """The three `nd_comm` callbacks over torch.distributed.  device='cuda' (NCCL) or 'cpu' (gloo)."""

    def __init__(self, group=None, device: str = "cuda", stream_ptr: int = 0):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.nranks = dist.get_world_size(group)
        self.device = device
        self.left = (self.rank - 1) % self.nranks
        self.right = (self.rank + 1) % self.nranks
        self.stream_ptr = stream_ptr
        self.n_allreduce = self.n_sendrecv = 0
        self.bytes_sent = 0

    # ---- python-level API (also what the gloo tests exercise) ----
    def _scratch(self):
        """Reused device / pinned-host staging for the small collectives (allocation and tensor construction per call cost more
        than the collective itself)."""
        if getattr(self, "_dev", None) is None:
            torch = self.torch
            pin = self.device != "cpu"
            self._host = torch.zeros(32, dtype=torch.float64, pin_memory=pin)
            self._dev = torch.zeros(32, dtype=torch.float64, device=self.device) if pin else self._host
            self._ihost = torch.zeros(2 * self.nranks, dtype=torch.int64, pin_memory=pin)
            self._idev = torch.zeros(2 * self.nranks, dtype=torch.int64, device=self.device) if pin else self._ihost
            self._imine = torch.zeros(2, dtype=torch.int64, device=self.device)
            self._rops = {OP_MAX: self.dist.ReduceOp.MAX, OP_MIN: self.dist.ReduceOp.MIN, OP_SUM: self.dist.ReduceOp.SUM}
        return self._host, self._dev

    def allreduce(self, vals, op: int):
        torch, dist = self.torch, self.dist
        host, dev = self._scratch()
        n = len(vals)
        for i in range(n):
            host[i] = vals[i]
        if dev is not host:
            dev[:n].copy_(host[:n], non_blocking=True)
        dist.all_reduce(dev[:n], op=self._rops[op], group=self.group)
        if dev is not host:
            host[:n].copy_(dev[:n], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        self.n_allreduce += 1
        return host[:n].tolist()

    def exchange_counts(self, send_left: int, send_right: int):
        """Returns (bytes arriving from the left neighbour, bytes arriving from the right neighbour)."""
        torch, dist = self.torch, self.dist
        self._scratch()
        self._ihost[0], self._ihost[1] = send_left, send_right
        self._imine.copy_(self._ihost[:2], non_blocking=True)
        dist.all_gather_into_tensor(self._idev, self._imine, group=self.group) if self.device != "cpu" else \
            dist.all_gather(list(self._idev.view(self.nranks, 2).unbind(0)), self._imine, group=self.group)
        if self._idev is not self._ihost:
            self._ihost.copy_(self._idev, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        allc = self._ihost.view(self.nranks, 2)
        # what my left neighbour sends to ITS right is for me, and vice versa
        return int(allc[self.left][1]), int(allc[self.right][0])

    def sendrecv_tensors(self, send_left, send_right, recv_left, recv_right):
        """uint8 tensors (or None for empty).  Matching order for a 2-rank ring, where both neighbours are the same peer:
        every rank posts [send->right, send->left, recv<-left, recv<-right]; sends and receives between one pair of ranks
        are matched in posting order, so the peer's first send (to ITS right = my left side) meets my first recv."""
        dist = self.dist
        ops = []
        if send_right is not None and send_right.numel():
            ops.append(dist.P2POp(dist.isend, send_right, self._peer(self.right), self.group))
        if send_left is not None and send_left.numel():
            ops.append(dist.P2POp(dist.isend, send_left, self._peer(self.left), self.group))
        if recv_left is not None and recv_left.numel():
            ops.append(dist.P2POp(dist.irecv, recv_left, self._peer(self.left), self.group))
        if recv_right is not None and recv_right.numel():
            ops.append(dist.P2POp(dist.irecv, recv_right, self._peer(self.right), self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        self.n_sendrecv += 1
        self.bytes_sent += (send_left.numel() if send_left is not None else 0) + (send_right.numel() if send_right is not None else 0)

    def _peer(self, group_rank: int) -> int:
        return self.dist.get_global_rank(self.group, group_rank) if self.group is not None else group_rank

    # ---- C callbacks ----
    def _view(self, ptr: int, nbytes: int):
        torch = self.torch
        if nbytes == 0 or not ptr:
            return None
        if self.device == "cpu":
            buf = (C.c_ubyte * nbytes).from_address(ptr)
            return torch.from_numpy(np.frombuffer(buf, dtype=np.uint8))
        return torch.as_tensor(_DevPtr(ptr, nbytes), device=self.device)

    def callbacks(self):
        torch = self.torch

        def _allreduce(user, v, n, op):
            try:
                out = self.allreduce([v[i] for i in range(n)], op)
                for i in range(n):
                    v[i] = out[i]
                return 0
            except Exception as e:  # pragma: no cover - reported through the C error path
                print("ndspmhd_b200.slab allreduce failed:", e, flush=True)
                return 1

        def _counts(user, sb, rb):
            try:
                rb[0], rb[1] = self.exchange_counts(sb[0], sb[1])
                return 0
            except Exception as e:  # pragma: no cover
                print("ndspmhd_b200.slab sendrecv_counts failed:", e, flush=True)
                return 1

        def _sendrecv(user, sbuf, sb, rbuf, rb, stream):
            try:
                views = [self._view(sbuf[0], sb[0]), self._view(sbuf[1], sb[1]), self._view(rbuf[0], rb[0]), self._view(rbuf[1], rb[1])]
                if self.device == "cpu":
                    self.sendrecv_tensors(*views)
                else:
                    # order the NCCL transfers after the pack kernels and before the unpack kernels on the library's stream
                    with torch.cuda.stream(torch.cuda.ExternalStream(int(stream))):
                        self.sendrecv_tensors(*views)
                return 0
            except Exception as e:  # pragma: no cover
                print("ndspmhd_b200.slab sendrecv failed:", e, flush=True)
                return 1

        self._cb = (_ALLREDUCE(_allreduce), _COUNTS(_counts), _SENDRECV(_sendrecv))   # keep alive
        return self._cb


def attach(hot, comm: SlabComm, slab_lo: float, slab_hi: float, nglobal: int) -> None:
    """Registers the transport with a Hotpath context (ndspmhd_b200_set_comm)."""
    cb = comm.callbacks()
    nc = NdComm()
    nc.user = None
    nc.rank, nc.nranks = comm.rank, comm.nranks
    nc.slab_lo, nc.slab_hi = slab_lo, slab_hi
    nc.nglobal = nglobal
    nc.allreduce, nc.sendrecv_counts, nc.sendrecv = cb
    hot._comm_struct = nc
    hot._comm = comm
    L = hot.L
    L.ndspmhd_b200_set_comm.argtypes = [C.c_void_p, C.POINTER(NdComm)]
    hot._chk(L.ndspmhd_b200_set_comm(hot.ctx, C.byref(nc)))


def attach_nccl(hot, rank: int, nranks: int, slab_lo: float, slab_hi: float, nglobal: int, group=None) -> None:
    """The native transport (ndspmhd_b200_set_comm_nccl): the library runs its collectives itself over NCCL on its own stream; the host
    only carries the 128-byte communicator id from rank 0 to the others (here through torch.distributed's object broadcast)."""
    import torch.distributed as dist

    L = hot.L
    buf = C.create_string_buffer(128)
    if rank == 0:
        hot._chk(L.ndspmhd_b200_nccl_unique_id(buf))
    obj = [buf.raw]
    dist.broadcast_object_list(obj, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    hot._chk(L.ndspmhd_b200_set_comm_nccl(hot.ctx, obj[0], rank, nranks, slab_lo, slab_hi, nglobal))
    hot._comm = "nccl"


def comm_stats(hot):
    """(all-reduces issued, halo payload bytes sent by this rank) since the context was created -- either transport."""
    a, b = C.c_longlong(), C.c_longlong()
    hot.L.ndspmhd_b200_comm_stats(hot.ctx, C.byref(a), C.byref(b))
    return a.value, b.value


def row_counts(hot):
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    hot.L.ndspmhd_b200_row_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    hot.L.ndspmhd_b200_row_counts(hot.ctx, C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value
