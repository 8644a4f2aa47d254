"""Row-partitioned extraction across the GPUs of one box (one process per GPU).

The reference's only parallel mechanism is PETSc's MPIAIJ row-block decomposition (``mpirun --np N``:
reference common.py:673-677 sets local row sizes, la_utils.py:116-125 updates ghosts, MatMatMult
fetches off-process rows internally).  The same layout is used here:

  rank r owns foreground rows [f_r, f_{r+1}) of A_f and M, and background rows [b_r, b_{r+1}) of
  A_b, b_b, u_b, with PETSC_DECIDE-style contiguous blocks (:func:`row_partition`).

PtAP is "owner computes" by OUTPUT row: a rank gathers (once per sparsity pattern)
  1. its block of M^T                 — triplets of M routed to the owner of their column,
  2. the A_f rows that block touches  — J = columns of the block (local rows + ghost rows),
  3. the M rows those A_f rows touch  — K = columns of A_f[J, :],
renumbers them compactly and runs the single-GPU two-phase kernel on the local triple product
C_r = R A P (``iife_rap_symbolic`` / ``iife_rap_numeric``); no partial results cross ranks.  On every
numeric call only the VALUES of the ghost rows are re-exchanged through the stored gather plan.
The solve uses the local block of A_b with columns renumbered [owned | ghost], a halo plan executed
by libiife (pack kernel + grouped ncclSend/ncclRecv) and NCCL allreduces for the dot products.

Setup-time routing below is torch tensor plumbing + ``torch.distributed`` point-to-point (it runs on
CPU tensors with gloo in the tests); the per-call hot work is in the CUDA library.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import check, lib


# --------------------------------------------------------------------------------------------------
# partition + collectives
# --------------------------------------------------------------------------------------------------
def row_partition(n: int, nranks: int) -> np.ndarray:
    """PETSC_DECIDE split (reference la_utils.py:87-90): n // P rows each, the first n % P ranks one more."""
    base, rem = divmod(int(n), int(nranks))
    sizes = np.full(nranks, base, dtype=np.int64)
    sizes[:rem] += 1
    off = np.zeros(nranks + 1, dtype=np.int64)
    np.cumsum(sizes, out=off[1:])
    return off


def _world():
    return (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)


def alltoallv(send, group=None, recv_counts=None, recv_out=None):
    """send[q] = 1-D tensor for rank q (all same dtype/device).  Returns the list received from each rank.
    Counts travel with all_gather (skipped when the caller knows ``recv_counts`` from a stored plan: no host
    synchronisation then), payloads with batched isend/irecv (works with NCCL and gloo).  ``recv_out``: preallocated
    tensors to receive into (the own block is copied into recv_out[rank])."""
    rank, world = _world()
    if world == 1:
        if recv_out is not None:
            recv_out[0].copy_(send[0])
            return recv_out
        return [send[0]]
    dev, dtype = send[0].device, send[0].dtype
    if recv_counts is None:
        counts = torch.tensor([int(t.numel()) for t in send], dtype=torch.int64, device=dev)
        gathered = [torch.empty_like(counts) for _ in range(world)]
        dist.all_gather(gathered, counts, group=group)
        recv_counts = [int(gathered[q][rank].item()) for q in range(world)]
    if recv_out is not None:
        recv = recv_out
        recv[rank].copy_(send[rank])
    else:
        recv = [torch.empty(int(recv_counts[q]), dtype=dtype, device=dev) for q in range(world)]
        recv[rank] = send[rank]
    ops = []
    for q in range(world):
        if q == rank:
            continue
        if send[q].numel() > 0:
            ops.append(dist.P2POp(dist.isend, send[q].contiguous(), q, group=group))
        if recv[q].numel() > 0:
            ops.append(dist.P2POp(dist.irecv, recv[q], q, group=group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return recv


def _split_by_owner(values_sorted_by_owner, counts):
    out, pos = [], 0
    for c in counts:
        out.append(values_sorted_by_owner[pos:pos + c])
        pos += c
    return out


def _owner_of(ids, part_t):
    """rank owning each global id under the partition offsets part_t (tensor of P+1 entries)."""
    return torch.bucketize(ids, part_t[1:], right=True)


def _seg_positions(starts, lens):
    """concatenation of ranges [starts[k], starts[k]+lens[k])"""
    total = int(lens.sum().item())
    if total == 0:
        return torch.empty(0, dtype=torch.int64, device=starts.device)
    offs = torch.cumsum(lens, 0) - lens
    return torch.arange(total, device=starts.device) - torch.repeat_interleave(offs, lens) + torch.repeat_interleave(starts, lens)


class RowFetchPlan:
    """How to (re)fetch rows ``wanted`` (sorted global ids) of a row-partitioned CSR matrix."""

    def __init__(self):
        self.gather_pos = None   # per peer: positions in MY val array to send
        self.req_local = None    # per peer: MY local row indices that peer asked for
        self.rowptr = None       # fetched pattern
        self.colind = None


def fetch_rows(rowptr, colind, val, row_start, part_t, wanted):
    """Fetch rows ``wanted`` (sorted, global ids) of the matrix whose local block is (rowptr, colind, val)
    starting at global row ``row_start``.  Returns (plan, rowptr_w, colind_w, val_w)."""
    rank, world = _world()
    own = _owner_of(wanted, part_t)
    counts = torch.bincount(own, minlength=world).tolist()
    req_out = _split_by_owner(wanted, counts)           # already grouped: wanted is sorted
    req_in = alltoallv(req_out)                          # rows other ranks want from me
    plan = RowFetchPlan()
    plan.req_local, plan.gather_pos = [], []
    send_len, send_col, send_val = [], [], []
    for q in range(world):
        idx = (req_in[q] - row_start).to(torch.int64)
        starts = rowptr[idx].to(torch.int64)
        lens = (rowptr[idx + 1] - rowptr[idx]).to(torch.int64)
        pos = _seg_positions(starts, lens)
        plan.req_local.append(idx)
        plan.gather_pos.append(pos)
        send_len.append(lens)
        send_col.append(colind[pos])
        send_val.append(val[pos])
    lens_parts = alltoallv(send_len)
    lens_w = torch.cat(lens_parts)
    # element counts of the blocks this rank receives: static for the plan (value refreshes skip the count exchange)
    plan.recv_entry_counts = [int(t.sum().item()) for t in lens_parts]
    plan.recv_row_counts = [int(t.numel()) for t in lens_parts]
    col_w = torch.cat(alltoallv(send_col, recv_counts=plan.recv_entry_counts))
    val_w = torch.cat(alltoallv(send_val, recv_counts=plan.recv_entry_counts))
    rowptr_w = torch.zeros(wanted.numel() + 1, dtype=torch.int64, device=wanted.device)
    torch.cumsum(lens_w, 0, out=rowptr_w[1:])
    plan.rowptr, plan.colind = rowptr_w, col_w
    return plan, rowptr_w, col_w, val_w


def refresh_values(plan: RowFetchPlan, val, out=None):
    """values of the fetched rows for new local values ``val`` (same pattern).  The block a rank "sends
    to itself" is usually ALL of its rows in order: then it is passed through without a gather."""
    rank, _ = _world()
    if getattr(plan, "self_identity", None) is None:
        p = plan.gather_pos[rank]
        plan.self_identity = bool(p.numel() == val.numel() and (p.numel() == 0 or (int(p[0]) == 0 and int(p[-1]) == p.numel() - 1
                                                                                   and bool((p[1:] > p[:-1]).all()))))
    send = [val if (q == rank and plan.self_identity) else val[p] for q, p in enumerate(plan.gather_pos)]
    if out is not None:  # straight into the operand's value array: no concatenation, no second copy
        offs = [0]
        for cnt in plan.recv_entry_counts:
            offs.append(offs[-1] + int(cnt))
        alltoallv(send, recv_counts=plan.recv_entry_counts, recv_out=[out[offs[q]:offs[q + 1]] for q in range(len(offs) - 1)])
        return out
    return torch.cat(alltoallv(send, recv_counts=plan.recv_entry_counts))


def fetch_entries(plan: RowFetchPlan, vec_local):
    """entries of a row-partitioned VECTOR at the rows of ``plan`` (b_f at J)."""
    rank, _ = _world()
    if getattr(plan, "self_rows_identity", None) is None:  # the rows a rank "requests from itself" are usually all of its rows in order
        p = plan.req_local[rank]
        plan.self_rows_identity = bool(p.numel() == vec_local.numel() and (p.numel() == 0 or (
            int(p[0]) == 0 and int(p[-1]) == p.numel() - 1 and bool((p[1:] > p[:-1]).all()))))
    send = [vec_local if (q == rank and plan.self_rows_identity) else vec_local[idx] for q, idx in enumerate(plan.req_local)]
    return torch.cat(alltoallv(send, recv_counts=plan.recv_row_counts))


# --------------------------------------------------------------------------------------------------
# distributed PtAP setup (pattern level) — pure tensor code, testable on CPU with gloo
# --------------------------------------------------------------------------------------------------
class LocalTriple:
    """The operands of one rank's local triple product C_r = R A P, compactly renumbered."""

    def __init__(self):
        self.R = self.A = self.P = None      # (n_rows, n_cols, rowptr, colind, val) with torch tensors
        self.J = self.K = None               # global foreground ids of the A rows / A cols kept
        self.planA = self.planM = None
        self.fg_part = self.bg_part = None
        self.rank = 0


def _csr_from_triplets_sorted_by_col(rows_local, n_rows, cols, vals):
    """CSR from triplets already sorted by column inside equal rows after a STABLE sort by row."""
    order = torch.argsort(rows_local, stable=True)
    rp = torch.zeros(n_rows + 1, dtype=torch.int64, device=rows_local.device)
    torch.cumsum(torch.bincount(rows_local, minlength=n_rows), 0, out=rp[1:])
    return rp, cols[order], vals[order], order


def setup_local_triple(n_f, n_b, M_loc, A_loc, group=None) -> LocalTriple:
    """M_loc, A_loc: this rank's row blocks as (rowptr, colind, val) torch tensors with GLOBAL column ids
    (rows [f_r, f_{r+1}) under row_partition(n_f, P))."""
    rank, world = _world()
    dev = M_loc[0].device
    fg_part = torch.as_tensor(row_partition(n_f, world), device=dev)
    bg_part = torch.as_tensor(row_partition(n_b, world), device=dev)
    f0 = int(fg_part[rank].item())
    b0, b1 = int(bg_part[rank].item()), int(bg_part[rank + 1].item())
    m_rp, m_ci, m_v = (t.to(torch.int64) if t.dtype != torch.float64 else t for t in M_loc)
    a_rp, a_ci, a_v = (t.to(torch.int64) if t.dtype != torch.float64 else t for t in A_loc)
    n_loc = m_rp.numel() - 1
    # 1. my block of M^T: route every entry of my M rows to the owner of its column
    lens = m_rp[1:] - m_rp[:-1]
    rows_g = torch.repeat_interleave(torch.arange(n_loc, device=dev) + f0, lens)
    own = _owner_of(m_ci, bg_part)
    order = torch.argsort(own, stable=True)
    counts = torch.bincount(own, minlength=world).tolist()
    r_j = torch.cat(alltoallv(_split_by_owner(rows_g[order], counts)))
    r_i = torch.cat(alltoallv(_split_by_owner(m_ci[order], counts)))
    r_w = torch.cat(alltoallv(_split_by_owner(m_v[order], counts)))
    # concatenation over source ranks is ascending in j (row blocks are ascending, rows inside a block too)
    R_rp, R_cj, R_v, mt_order = _csr_from_triplets_sorted_by_col(r_i - b0, b1 - b0, r_j, r_w)
    # 2. A_f rows touched: J
    mask = torch.zeros(n_f, dtype=torch.bool, device=dev)
    mask[R_cj] = True
    J = torch.nonzero(mask).flatten()
    planA, AJ_rp, AJ_ci, AJ_v = fetch_rows(a_rp, a_ci, a_v, f0, fg_part, J)
    # 3. M rows touched by those: K
    mask.zero_()
    mask[AJ_ci] = True
    K = torch.nonzero(mask).flatten()
    del mask
    planM, MK_rp, MK_ci, MK_v = fetch_rows(m_rp, m_ci, m_v, f0, fg_part, K)
    T = LocalTriple()
    T.rank, T.fg_part, T.bg_part, T.J, T.K, T.planA, T.planM = rank, fg_part, bg_part, J, K, planA, planM
    T.R = (b1 - b0, J.numel(), R_rp, torch.searchsorted(J, R_cj), R_v)
    T.A = (J.numel(), K.numel(), AJ_rp, torch.searchsorted(K, AJ_ci), AJ_v)
    T.P = (K.numel(), n_b, MK_rp, MK_ci, MK_v)
    # the M^T block's values in terms of an exchange (for new M values): keep the routing
    T.mt_route = (order, counts, mt_order)
    return T


def localize_operator(n_b, rowptr, colind, part_t, rank):
    """Local numbering [owned | ghost] of the columns of a row block with GLOBAL column ids.  Returns
    (local colind, ghost global ids sorted, halo description dict)."""
    rk, world = _world()
    b0, b1 = int(part_t[rank].item()), int(part_t[rank + 1].item())
    owned = (colind >= b0) & (colind < b1)
    mask = torch.zeros(n_b, dtype=torch.bool, device=colind.device)
    mask[colind[~owned]] = True
    G = torch.nonzero(mask).flatten()
    del mask
    local = torch.where(owned, colind - b0, (b1 - b0) + torch.searchsorted(G, colind))
    own = _owner_of(G, part_t)
    recv_counts = torch.bincount(own, minlength=world)
    want_in = alltoallv(_split_by_owner(G, recv_counts.tolist()))   # ids other ranks need from me
    send_counts = [int(t.numel()) for t in want_in]
    send_idx = torch.cat([(t - b0) for t in want_in]) if want_in else torch.empty(0, dtype=torch.int64)
    halo = {"n_owned": b1 - b0, "n_ghost": int(G.numel()), "send_counts": send_counts,
            "send_idx": send_idx.to(torch.int32), "recv_counts": recv_counts.tolist()}
    return local, G, halo


# --------------------------------------------------------------------------------------------------
# device side
# --------------------------------------------------------------------------------------------------
def init_comm():
    """Create libiife's NCCL communicator: rank 0 draws the unique id, torch.distributed broadcasts it."""
    rank, world = _world()
    if world == 1:
        check(lib.iife_comm_init(0, 1, None))
        return
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        raw = (ctypes.c_ubyte * 128)()
        check(lib.iife_comm_unique_id(raw))
        buf = torch.tensor(list(raw), dtype=torch.uint8)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = buf.to(dev)
    dist.broadcast(t, 0)
    raw = (ctypes.c_ubyte * 128)(*t.cpu().tolist())
    check(lib.iife_comm_init(rank, world, raw))


def _mat_from_tensors(core, shape, rp, ci, v, unsorted_ok=False):
    h = ctypes.c_void_p(0)
    rp32, ci32 = rp.to(torch.int32).contiguous(), ci.to(torch.int32).contiguous()
    vv = v.contiguous()
    fn = lib.iife_mat_create_csr_ex
    check(fn(int(shape[0]), int(shape[1]), ctypes.c_void_p(rp32.data_ptr()), ctypes.c_void_p(ci32.data_ptr()),
             ctypes.c_void_p(vv.data_ptr()), 4, core.MEM_DEVICE, 1 if unsorted_ok else 0, ctypes.byref(h)))
    lib.iife_sync()
    return core.DeviceMat(h.value)


class _RawCudaArray:
    """Minimal __cuda_array_interface__ carrier: lets torch wrap memory the library owns without copying."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def _device_view_f64(ptr, n):
    if not ptr or n <= 0:
        return None
    try:
        return torch.as_tensor(_RawCudaArray(ptr, n), device=torch.device("cuda", torch.cuda.current_device()))
    except Exception:  # fall back to the copying path
        return None


class DistExtraction:
    """Row-partitioned A_b = M^T A_f M, b_b = M^T b_f and Jacobi-CG on the GPUs of one box."""

    def __init__(self, n_f, n_b, M_loc, A_loc):
        from . import core

        self.core = core
        self.n_f, self.n_b = int(n_f), int(n_b)
        self.rank, self.world = _world()
        T = setup_local_triple(n_f, n_b, M_loc, A_loc)
        self.T = T
        self.R = _mat_from_tensors(core, T.R[:2], *T.R[2:])
        self.A = _mat_from_tensors(core, T.A[:2], *T.A[2:])
        self.P = _mat_from_tensors(core, T.P[:2], *T.P[2:])
        h = ctypes.c_void_p(0)
        check(lib.iife_rap_symbolic(self.R.handle, self.A.handle, self.P.handle, ctypes.byref(h)))
        self.plan = h
        self.C = None          # local block of A_b, GLOBAL column ids
        self.C_op = None       # same values, local [owned | ghost] column ids (KSP operator)
        self.halo = None
        self.n_owned = T.R[0]
        # torch view of the value array of the local A operand (None if the view cannot be made: values are copied then)
        self._A_val_view = _device_view_f64(self.A.device_ptrs()[2], self.A.nnz) if os.environ.get("IIFE_DIST_DIRECT", "1") != "0" else None
        # the pattern of this rank's rows of A_f as handed over (numeric_csr checks a fresh matrix against it)
        self._A_pattern = (A_loc[0].to(torch.int32, copy=True), A_loc[1].to(torch.int32, copy=True))

    def numeric_csr(self, rowptr, colind, values):
        """A freshly assembled local block of A_f (device CSR arrays: the reference builds a new matrix per
        ``assemble``, common.py:432-435): the pattern must be the one this object was set up with (compared on the
        device, one pass over the index arrays); then :meth:`numeric`.  Raises ValueError on a different pattern."""
        rp0, ci0 = self._A_pattern
        if (rowptr.numel() != rp0.numel() or colind.numel() != ci0.numel()
                or not bool(torch.equal(rowptr.to(torch.int32), rp0)) or not bool(torch.equal(colind.to(torch.int32), ci0))):
            raise ValueError("the pattern of A_f changed: set up a new DistExtraction")
        return self.numeric(values)

    def numeric(self, A_val_local):
        """New foreground values (same pattern): exchange ghost-row values, run the numeric phase."""
        if self._A_val_view is not None:  # the fetched values land in the operand's own array (same entry order)
            refresh_values(self.T.planA, A_val_local, out=self._A_val_view)
            check(lib.iife_mat_touch(self.A.handle))
        else:
            self.A.update_values(refresh_values(self.T.planA, A_val_local))
        h = ctypes.c_void_p(self.C.handle.value if self.C is not None else 0)
        check(lib.iife_rap_numeric(self.plan, self.R.handle, self.A.handle, self.P.handle, ctypes.byref(h)))
        if self.C is None:
            self.C = self.core.DeviceMat(h.value)
            self._build_operator()
        else:
            # refresh the KSP operator's values: same entry order as C
            _, _, cv = self.C.device_ptrs()
            check(lib.iife_mat_update_values(self.C_op.handle, ctypes.c_void_p(cv), self.core.MEM_DEVICE))
        return self.C

    def _build_operator(self):
        core = self.core
        n_rows, _, nnz = self.C.info()
        dev = torch.device("cuda", torch.cuda.current_device())
        rp = torch.empty(n_rows + 1, dtype=torch.int32, device=dev)
        ci = torch.empty(nnz, dtype=torch.int32, device=dev)
        v = torch.empty(nnz, dtype=torch.float64, device=dev)
        check(lib.iife_mat_get_csr(self.C.handle, ctypes.c_void_p(rp.data_ptr()), ctypes.c_void_p(ci.data_ptr()),
                                   ctypes.c_void_p(v.data_ptr()), 4, core.MEM_DEVICE))
        lib.iife_sync()
        local, G, halo = localize_operator(self.n_b, rp.to(torch.int64), ci.to(torch.int64), self.T.bg_part, self.rank)
        self.ghost_ids = G
        self.C_op = _mat_from_tensors(core, (n_rows, n_rows + int(G.numel())), rp, local, v, unsorted_ok=True)
        sc = np.asarray(halo["send_counts"], dtype=np.int64)
        rc = np.asarray(halo["recv_counts"], dtype=np.int64)
        si = halo["send_idx"].cpu().numpy().astype(np.int32)
        h = ctypes.c_void_p(0)
        check(lib.iife_halo_create(int(halo["n_owned"]), int(halo["n_ghost"]), sc.ctypes.data_as(ctypes.c_void_p),
                                   si.ctypes.data_as(ctypes.c_void_p), rc.ctypes.data_as(ctypes.c_void_p), ctypes.byref(h)))
        self.halo = h
        self.p2p = False
        if self.world > 1 and self.world <= 16 and os.environ.get("IIFE_P2P", "1") != "0":
            self._attach_p2p(halo)

    def _attach_p2p(self, halo):
        """NVLink peer-memory path: all-gather the IPC handles of every rank's solver vector + mailbox and
        the offsets at which each rank's send block lands in its neighbours' ghost sections."""
        dev = torch.device("cuda", torch.cuda.current_device())
        raw = (ctypes.c_ubyte * 128)()
        check(lib.iife_halo_p2p_export(self.halo, raw))
        mine = torch.tensor(list(raw), dtype=torch.uint8, device=dev)
        allh = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allh, mine)
        # start of the block received from each source inside MY vector: n_owned + recv offset
        rc = torch.tensor(halo["recv_counts"], dtype=torch.int64, device=dev)
        starts = int(halo["n_owned"]) + torch.cumsum(rc, 0) - rc
        all_starts = [torch.empty_like(starts) for _ in range(self.world)]
        dist.all_gather(all_starts, starts)
        dst = np.array([int(all_starts[q][self.rank].item()) for q in range(self.world)], dtype=np.int64)
        blob = torch.cat(allh).cpu().numpy().tobytes()
        buf = (ctypes.c_ubyte * len(blob)).from_buffer_copy(blob)
        check(lib.iife_halo_p2p_attach(self.halo, buf, dst.ctypes.data_as(ctypes.c_void_p)))
        dist.barrier()
        self.p2p = True

    def rhs(self, b_f_local):
        """b_b (owned block) = M^T b_f: the M^T block times the gathered entries of b_f."""
        bJ = fetch_entries(self.T.planA, b_f_local)
        return self.R.spmv(bJ)

    def _halo_from(self, halo):
        sc = np.asarray(halo["send_counts"], dtype=np.int64)
        rc = np.asarray(halo["recv_counts"], dtype=np.int64)
        si = halo["send_idx"].cpu().numpy().astype(np.int32)
        h = ctypes.c_void_p(0)
        check(lib.iife_halo_create(int(halo["n_owned"]), int(halo["n_ghost"]), sc.ctypes.data_as(ctypes.c_void_p),
                                   si.ctypes.data_as(ctypes.c_void_p), rc.ctypes.data_as(ctypes.c_void_p), ctypes.byref(h)))
        return h

    def transfer_to_foreground(self, u_b_owned, M_loc):
        """u_f (owned foreground rows) = M u_b with ghost entries of u_b fetched from their owners
        (reference common.py:123-140 under MPI: Mat.mult + ghost update).  ``M_loc`` = this rank's rows of M
        (rowptr, colind, val) with GLOBAL background column ids."""
        core = self.core
        dev = u_b_owned.device
        if getattr(self, "_M_op", None) is None:
            rp, ci, v = M_loc
            # M's columns live in the BACKGROUND partition, its rows in the foreground one: the "owned" column
            # block of this rank is its background block
            local, G, halo = localize_operator(self.n_b, rp.to(torch.int64), ci.to(torch.int64), self.T.bg_part, self.rank)
            self._M_op = _mat_from_tensors(core, (rp.numel() - 1, self.n_owned + int(G.numel())), rp, local, v, unsorted_ok=True)
            self._M_halo = self._halo_from(halo)
            self._M_xext = torch.zeros(self.n_owned + int(G.numel()), dtype=torch.float64, device=dev)
        self._M_xext[: self.n_owned].copy_(u_b_owned)
        u_f = torch.empty(self._M_op.shape[0], dtype=torch.float64, device=dev)
        check(lib.iife_spmv_dist(self._M_op.handle, self._M_halo, ctypes.c_void_p(self._M_xext.data_ptr()),
                                 ctypes.c_void_p(u_f.data_ptr())))
        return u_f

    def solve(self, b_owned, x_owned, rtol=1e-8, atol=1e-9, max_it=1000000, method="cg", restart=300):
        core = self.core
        res = _lib.KspResult()
        ksp_type = core.KSP_CG if method == "cg" else core.KSP_FGMRES
        if os.environ.get("IIFE_DBG_FIXED_ITS"):  # timing experiments: exactly this many iterations
            rtol, atol, max_it = 1e-300, 1e-300, int(os.environ["IIFE_DBG_FIXED_ITS"])
        check(lib.iife_ksp_solve_dist(self.C_op.handle, self.halo, ksp_type, core.PC_JACOBI, rtol, atol, 1e4,
                                      int(max_it), int(restart), ctypes.c_void_p(b_owned.data_ptr()),
                                      ctypes.c_void_p(x_owned.data_ptr()), ctypes.byref(res), None, 0))
        if self.p2p:
            err = ctypes.c_int(0)
            check(lib.iife_halo_p2p_error(self.halo, ctypes.byref(err)))
            if err.value:
                raise RuntimeError("peer-memory exchange timed out (a rank did not arrive)")
        return core.KSPInfo(int(res.iterations), int(res.reason), float(res.rnorm), float(res.rnorm0), np.zeros(0))
