"""The dense-root algorithm of csrc/dense_root.cuh + islam_b200/dist.py, restated in NumPy and run on CPU ranks over gloo
(world 1 / 2 / 4): the pieces the CUDA kernels implement — Cholesky of a diagonal block TOGETHER with its inverse by
factoring [A; I] in narrow column steps (k_root_potrf), the panel below as a product with that inverse (k_root_trsm), the
two-panel block step with its narrow in-block update, the K = block trailing update on the tile columns a rank OWNS
(k_root_syrk, 1-D block-column-cyclic on the absolute tile grid), the partial-root all-reduce with the diagonal clamped
after the sum, the broadcast of every factored block from its owner, and the right-looking back-substitution with the
stored inverse (k_root_back) — must reproduce a dense solve.  Tile sizes are scaled down (T = 16 instead of 128) so that
ragged last blocks, several tile columns per rank and ranks without any column all occur at n ~ 50-120."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

T, NB, IB = 16, 8, 4          # tile column = ownership granule / block step, panel width, inner step of the diagonal block


def potrf_with_inverse(R, k0, nbk):
    """k_root_potrf: L into the lower triangle of the block, E = L^-T (strictly upper part) into its upper triangle."""
    M = np.zeros((2 * NB, NB))
    M[:nbk, :nbk] = np.tril(R[k0:k0 + nbk, k0:k0 + nbk])
    M[NB:NB + nbk, :nbk] = np.eye(nbk)
    for j0 in range(0, nbk, IB):
        w = min(IB, nbk - j0)
        D = M[j0:j0 + w, j0:j0 + w]
        Ljj = np.linalg.cholesky(np.tril(D) + np.tril(D, -1).T)
        W = np.linalg.inv(Ljj)
        rows = list(range(j0 + w, nbk)) + [NB + r for r in range(j0 + w)]          # A rows below + the E rows that are non-zero
        M[rows, j0:j0 + w] = M[rows, j0:j0 + w] @ W.T
        M[j0:j0 + w, j0:j0 + w] = Ljj
        for j in range(j0 + w, nbk):                                               # trailing columns
            for r in rows:
                if r < NB and r < j:
                    continue                                                       # upper triangle of A
                M[r, j] -= M[r, j0:j0 + w] @ M[j, j0:j0 + w]
    L, E = np.tril(M[:nbk, :nbk]), M[NB:NB + nbk, :nbk]
    assert np.allclose(np.tril(E, -1), 0.0) and np.allclose(E, np.linalg.inv(L).T, atol=1e-9 * np.abs(E).max())
    R[k0:k0 + nbk, k0:k0 + nbk] = L + np.triu(E, 1)


def e_full(R, k0, nbk):
    """The upper-triangular L^-T of a factored block: stored strict upper part + reciprocal diagonal of L."""
    B = R[k0:k0 + nbk, k0:k0 + nbk]
    return np.triu(B, 1) + np.diag(1.0 / np.diag(B))


def trsm(R, k0, nbk, n):
    R[k0 + nbk:n + 1, k0:k0 + nbk] = R[k0 + nbk:n + 1, k0:k0 + nbk] @ e_full(R, k0, nbk)      # X = A L^-T = A E


def syrk(R, k0, nk, base, c_hi, n, G, rank):
    """k_root_syrk: columns [base, min(c_hi, n)) of the tile columns this rank owns, rows >= column, down to the rhs row n."""
    P = R[:, k0:k0 + nk]
    for c in range(base, min(c_hi, n)):
        if (c // T) % G != rank:
            continue
        R[c:n + 1, c] -= P[c:n + 1] @ P[c]


def factor_distributed(R, n, G, rank, bcast):
    for k0 in range(0, n, T):
        owner = (k0 // T) % G
        if owner == rank:                                   # islam_pvgo_root_panel
            nb1 = min(NB, n - k0)
            potrf_with_inverse(R, k0, nb1); trsm(R, k0, nb1, n)
            if k0 + NB < n:
                syrk(R, k0, NB, k0 + NB, k0 + T, n, G, rank)
                nb2 = min(NB, n - k0 - NB)
                potrf_with_inverse(R, k0 + NB, nb2); trsm(R, k0 + NB, nb2, n)
        nk = min(T, n - k0)
        R[:, k0:k0 + nk] = bcast(R[:, k0:k0 + nk], owner)    # the factored block, whole columns (the inverse rides along)
        syrk(R, k0, nk, k0 + nk, n, n, G, rank)              # islam_pvgo_root_update


def back_substitute(R, n):
    """k_root_back: right-looking, x_b = E t_b with the stored inverse, then the columns left of the block."""
    t = R[n, :n].copy()
    x = np.zeros(n)
    for k0 in reversed(range(0, n, NB)):
        nbk = min(NB, n - k0)
        x[k0:k0 + nbk] = e_full(R, k0, nbk) @ t[k0:k0 + nbk]
        t[:k0] -= R[k0:k0 + nbk, :k0].T @ x[k0:k0 + nbk]
    return x


def _problem(n, seed):
    rng = np.random.default_rng(seed)
    B = rng.standard_normal((n, n))
    A = B @ B.T / n + np.diag(rng.random(n) * 1e-6)         # some diagonals fall under the clamp below
    return A, rng.standard_normal(n)


LM_MIN, SCALE = 1e-4, 1.0 + 1e-2


def _reference(A, b):
    Ad = A.copy()
    d = np.clip(np.diag(A), LM_MIN, 1e32) * SCALE
    Ad[np.arange(len(b)), np.arange(len(b))] = d
    return np.linalg.solve(Ad, b)


def _worker(rank, world, port, n, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    A, b = _problem(n, 5)
    # this rank's SHARE of the root (as the factors it owns would give it): a random split of A and b that sums to them
    rng = np.random.default_rng(100 + rank)
    shares = np.random.default_rng(7).dirichlet(np.ones(world), size=(n, n))
    shares = (shares + shares.transpose(1, 0, 2)) / 2
    R = np.zeros((n + 1, n))
    R[:n] = np.tril(A * shares[:, :, rank])
    R[n] = b / world
    diag = np.diag(R[:n]).copy()                              # the original diagonal travels apart: its clamp is not linear
    R[np.arange(n), np.arange(n)] = 0.0

    def allreduce(a):
        t = torch.from_numpy(np.ascontiguousarray(a))
        dist.all_reduce(t)
        return t.numpy()

    def bcast(a, src):
        t = torch.from_numpy(np.ascontiguousarray(a))
        dist.broadcast(t, src)
        return t.numpy()
    R = allreduce(R)
    diag = allreduce(diag)
    R[np.arange(n), np.arange(n)] += np.clip(diag, LM_MIN, 1e32) * SCALE            # k_root_diag
    factor_distributed(R, n, world, rank, bcast)
    x = back_substitute(R, n)                                  # replicated: every rank received every block
    xs = [torch.zeros(n, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(xs, torch.from_numpy(x))
    if rank == 0:
        np.savez(out, x=np.stack([t.numpy() for t in xs]), L=np.tril(R[:n]))
    dist.destroy_process_group()


@pytest.mark.parametrize('world,n', [(1, 53), (2, 53), (2, 64), (4, 37), (4, 121)])
def test_distributed_dense_root_scheme_reproduces_dense_solve(world, n, tmp_path):
    out = str(tmp_path / 'root.npz')
    mp.spawn(_worker, args=(world, 29700 + 7 * world + n % 17, n, out), nprocs=world, join=True)
    r = np.load(out)
    A, b = _problem(n, 5)
    xref = _reference(A, b)
    for x in r['x']:                                           # identical on every rank, equal to the dense solve
        assert np.array_equal(x, r['x'][0])
        assert np.abs(x - xref).max() <= 1e-8 * np.abs(xref).max()
    Ad = A.copy()
    Ad[np.arange(n), np.arange(n)] = np.clip(np.diag(A), LM_MIN, 1e32) * SCALE
    assert np.abs(r['L'] - np.linalg.cholesky(Ad)).max() <= 1e-9 * np.abs(r['L']).max()
