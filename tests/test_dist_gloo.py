"""CPU tests (gloo, world_size 2 and 3) of the row-partitioned path's host logic: the routing that
gathers each rank's block of M^T, the A_f ghost rows and the M ghost rows (iife_b200.dist), the value
refresh plan, and the [owned | ghost] renumbering + halo plan of the solver.  Local arithmetic in these
tests is done by the oracle (test infrastructure); on the GPU box the same plans drive libiife.so."""
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
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_cells, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "interpolation-based-immersed-fea_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from iife_b200 import dist as idist
        from oracle import oracle as O
        from oracle.synthetic_cube import assemble_cube

        if isinstance(n_cells, tuple):  # ("s2", N_f, degree): unfitted case — rotated background, empty rows of A_b,
            from iife_b200 import synthetic  # background blocks that do not follow the foreground slabs

            gsy = synthetic.unfitted_operators(n_cells[1], n_cells[2])
            A = O.CSR(gsy["n_f"], gsy["n_f"], *gsy["A"])
            M = O.CSR(gsy["n_f"], gsy["n_b"], *gsy["M"])
            b = gsy["b_f"]
        else:
            A, M, b = assemble_cube(n_cells)
        rng = np.random.default_rng(0)
        # break the symmetry/regularity a little: drop the support of two background functions
        n_f, n_b = A.n_rows, M.n_cols
        fpart = idist.row_partition(n_f, world)
        bpart = idist.row_partition(n_b, world)
        f0, f1 = int(fpart[rank]), int(fpart[rank + 1])
        b0, b1 = int(bpart[rank]), int(bpart[rank + 1])

        def block(C, r0, r1):
            rp = C.rowptr[r0:r1 + 1] - C.rowptr[r0]
            sl = slice(C.rowptr[r0], C.rowptr[r1])
            return (torch.from_numpy(rp.copy()), torch.from_numpy(C.colind[sl].astype(np.int64)), torch.from_numpy(C.val[sl].copy()))

        T = idist.setup_local_triple(n_f, n_b, block(M, f0, f1), block(A, f0, f1))

        def csr(t):
            return O.CSR(int(t[0]), int(t[1]), t[2].numpy(), t[3].numpy().astype(np.int32), t[4].numpy())

        R, Al, Pl = csr(T.R), csr(T.A), csr(T.P)
        # the M^T block equals the corresponding rows of the global transpose (global column ids via J)
        MT = O.transpose(M)
        Jn = T.J.numpy()
        for i in range(b0, b1):
            g = slice(MT.rowptr[i], MT.rowptr[i + 1])
            l = slice(R.rowptr[i - b0], R.rowptr[i - b0 + 1])
            assert np.array_equal(Jn[R.colind[l]], MT.colind[g]) and np.array_equal(R.val[l], MT.val[g])
        # local triple product == my rows of the global product, bit for bit (same summation order)
        C = O.AT_R_A(M, A)
        Cl = O.matmult(O.matmult(R, Al), Pl)
        g = slice(C.rowptr[b0], C.rowptr[b1])
        assert np.array_equal(Cl.rowptr, C.rowptr[b0:b1 + 1] - C.rowptr[b0])
        assert np.array_equal(Cl.colind, C.colind[g])
        assert np.array_equal(Cl.val, C.val[g])
        # value refresh: new A values travel through the stored plan
        A2v = A.val * (1.0 + 0.1 * rng.standard_normal(A.nnz))
        v2 = idist.refresh_values(T.planA, torch.from_numpy(A2v[A.rowptr[f0]:A.rowptr[f1]].copy())).numpy()
        A2 = O.CSR(A.n_rows, A.n_cols, A.rowptr, A.colind, A2v)
        C2 = O.AT_R_A(M, A2)
        Cl2 = O.matmult(O.matmult(R, O.CSR(Al.n_rows, Al.n_cols, Al.rowptr, Al.colind, v2)), Pl)
        assert np.array_equal(Cl2.val, C2.val[g])
        # the same refresh received straight into a preallocated array (what DistExtraction.numeric does on the device)
        out = torch.full((v2.size,), np.nan, dtype=torch.float64)
        idist.refresh_values(T.planA, torch.from_numpy(A2v[A.rowptr[f0]:A.rowptr[f1]].copy()), out=out)
        assert np.array_equal(out.numpy(), v2)
        # b_b = M^T b_f through the gathered entries of b_f
        bJ = idist.fetch_entries(T.planA, torch.from_numpy(b[f0:f1].copy())).numpy()
        assert np.array_equal(bJ, b[Jn])
        bb = O.spmv(R, bJ)
        assert np.allclose(bb, O.AT_x(M, b)[b0:b1], rtol=1e-14, atol=0)
        # solver operator: [owned | ghost] renumbering + halo plan, exercised with gloo
        local, G, halo = idist.localize_operator(n_b, torch.from_numpy(Cl.rowptr), torch.from_numpy(Cl.colind.astype(np.int64)),
                                                 T.bg_part, rank)
        n_own = b1 - b0
        assert halo["n_owned"] == n_own and halo["n_ghost"] == G.numel() and sum(halo["recv_counts"]) == G.numel()
        x_glob = rng.standard_normal(n_b)          # same seed on every rank -> same vector
        x_own = torch.from_numpy(x_glob[b0:b1].copy())
        send = []
        pos = 0
        for q in range(world):
            c = halo["send_counts"][q]
            send.append(x_own[halo["send_idx"][pos:pos + c].to(torch.int64)])
            pos += c
        ghost = torch.cat(idist.alltoallv(send)).numpy()
        assert np.array_equal(ghost, x_glob[G.numpy()])
        x_ext = np.concatenate([x_glob[b0:b1], ghost])
        Cop = O.CSR(n_own, n_own + G.numel(), Cl.rowptr, local.numpy().astype(np.int32), Cl.val)
        y = O.spmv(Cop, x_ext)
        assert np.allclose(y, O.spmv(C, x_glob)[b0:b1], rtol=1e-13, atol=1e-300)
        # local numbering keeps the diagonal on (i, i)
        d = O.jacobi_inverse(Cop)
        assert np.array_equal(d, O.jacobi_inverse(C)[b0:b1])
        with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_cells", [(2, 3), (3, 4), (2, ("s2", 8, 1)), (3, ("s2", 6, 2))])
def test_row_partitioned_setup(tmp_path, world, n_cells):
    from oracle import oracle as O

    O.build()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_cells, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_row_partition_matches_petsc_decide():
    from iife_b200 import dist as idist

    assert idist.row_partition(10, 4).tolist() == [0, 3, 6, 8, 10]
    assert idist.row_partition(8, 8).tolist() == list(range(9))
    assert idist.row_partition(3, 5).tolist() == [0, 1, 2, 3, 3, 3]
