"""Seeded synthetic inputs for the orbital-optimisation path.

No molecule can be built here (pyscf / qiskit-nature are unavailable), so benchmarks and parity tests
run on synthetic tensors with the structure of the real ones:

* two-electron integrals: chemist-order (pq|rs) = sum_L B[pq,L] B[rs,L] with B symmetric in (p,q)
  (exactly 8-fold symmetric, positive semi-definite like a real ERI tensor), stored the way the
  reference stores them, g[p,q,r,s] = -1/2 (ps|qr)  (base_opt_orb_solver.py:90: minus the "++--"
  coefficient tensor), generated slab by slab so that no host copy of the M^4 tensor is needed;
* RDMs: ensemble-N-representable mixtures of Slater determinants in the reference's convention
  Gamma[p,q,r,s] = <a+_p a+_q a_s a_r> (base_opt_orb_solver.py:386), spin-blocked alpha|beta;
* U: Q factor of a seeded Gaussian matrix.
"""
from __future__ import annotations

import math

import numpy as np
import torch

SEED_ERI, SEED_H, SEED_RDM, SEED_U = 1234, 1235, 1236, 1237


def _factor(M: int, rank: int, seed: int) -> torch.Tensor:
    """B[p,q,L] = B[q,p,L] ~ N(0,1) exp(-(p+q)/M) / sqrt(rank), float64 on the CPU."""
    gen = torch.Generator().manual_seed(seed)
    B = torch.randn(M, M, rank, generator=gen, dtype=torch.float64)
    B = 0.5 * (B + B.transpose(0, 1))
    idx = torch.arange(M, dtype=torch.float64)
    decay = torch.exp(-(idx[:, None] + idx[None, :]) / M)
    return B * decay[:, :, None] / math.sqrt(rank)


def eri_spatial_shard(M: int, t0: int, mloc: int, seed: int = SEED_ERI, rank: int = 32,
                      device="cpu", scale: float = 1.0) -> torch.Tensor:
    """Rows [t0, t0+mloc) of the spatial tensor g[p,q,r,s] = -1/2 (ps|qr) (shape mloc x M x M x M)."""
    B = _factor(M, rank, seed).to(device)
    out = torch.empty(mloc, M, M, M, dtype=torch.float64, device=device)
    Bqr = B.reshape(M * M, rank)                       # (q r) x L
    for i in range(mloc):
        # g[p,q,r,s] = -1/2 sum_L B[p,s,L] B[q,r,L]
        torch.matmul(Bqr, B[t0 + i].transpose(0, 1), out=out[i].view(M * M, M))
    out.mul_(-0.5 * scale)
    return out


def eri_spatial_pair_packed(M: int, t0: int = 0, mloc: int = None, seed: int = SEED_ERI,
                            rank: int = 32, device="cpu", scale: float = 1.0) -> torch.Tensor:
    """The same tensor in pair-packed storage: [count, M, M] with the slabs g[t, q, :, :] of
    `distributed.pair_slab_list(M, t0, mloc)`; never materialises the dense shard."""
    from .distributed import pair_selected
    mloc = M - t0 if mloc is None else mloc
    B = _factor(M, rank, seed).to(device)
    rows = [[q for q in range(M) if pair_selected(t, q)] for t in range(t0, t0 + mloc)]
    out = torch.empty(sum(len(r) for r in rows), M, M, dtype=torch.float64, device=device)
    pos = 0
    for i, qs in enumerate(rows):
        # slab (t,q)[r,s] = -1/2 sum_L B[q,r,L] B[t,s,L]
        Bq = B[torch.tensor(qs, device=B.device)].reshape(len(qs) * M, rank)
        torch.matmul(Bq, B[t0 + i].transpose(0, 1), out=out[pos:pos + len(qs)].view(len(qs) * M, M))
        pos += len(qs)
    out.mul_(-0.5 * scale)
    return out


def eri_spatial(M: int, seed: int = SEED_ERI, rank: int = 32, device="cpu",
                scale: float = 1.0) -> torch.Tensor:
    return eri_spatial_shard(M, 0, M, seed, rank, device, scale)


def h_spatial(M: int, seed: int = SEED_H, device="cpu") -> torch.Tensor:
    gen = torch.Generator().manual_seed(seed)
    h = torch.randn(M, M, generator=gen, dtype=torch.float64)
    return (0.5 * (h + h.T)).to(device)


def random_partial_unitary(M: int, N: int, seed: int = SEED_U, device="cpu") -> torch.Tensor:
    gen = torch.Generator().manual_seed(seed)
    A = torch.randn(M, N, generator=gen, dtype=torch.float64)
    Q, R = torch.linalg.qr(A)
    Q = Q * torch.sign(torch.diagonal(R))[None, :]
    return Q.contiguous().to(device)


def rdms_spin(N: int, seed: int = SEED_RDM, n_dets: int = 8):
    """(D [2N,2N], Gamma [2N]^4): convex mixture of `n_dets` Slater determinants with
    n_alpha = n_beta = max(1, N//2) electrons; alpha spin-orbitals first, then beta."""
    rng = np.random.RandomState(seed)
    Q = 2 * N
    nocc = max(1, N // 2)
    w = rng.rand(n_dets)
    w /= w.sum()
    D = np.zeros((Q, Q))
    G = np.zeros((Q, Q, Q, Q))
    for m in range(n_dets):
        gam = np.zeros((Q, Q))
        for s in range(2):
            C, _ = np.linalg.qr(rng.randn(N, nocc))
            gam[s * N:(s + 1) * N, s * N:(s + 1) * N] = C @ C.T
        D += w[m] * gam
        G += w[m] * (np.einsum("pr,qs->pqrs", gam, gam) - np.einsum("ps,qr->pqrs", gam, gam))
    return torch.from_numpy(D), torch.from_numpy(G)


def rdms_spatial(N: int, seed: int = SEED_RDM, n_dets: int = 8, pattern: str = "abba"):
    """Spin-summed (D [N,N], Gamma [N]^4) matching integrals stored with spin pattern `pattern`
    ('abba': blocks (s,t,t,s) as in the reference; 'abab': blocks (s,t,s,t))."""
    D, G = rdms_spin(N, seed, n_dets)
    Ds = D[:N, :N] + D[N:, N:]
    Gs = torch.zeros(N, N, N, N, dtype=torch.float64)
    for s in range(2):
        for t in range(2):
            blk = (s, t, t, s) if pattern == "abba" else (s, t, s, t)
            sl = tuple(slice(b * N, (b + 1) * N) for b in blk)
            Gs += G[sl]
    return Ds.contiguous(), Gs.contiguous()


def spin_orbital_integrals(h: torch.Tensor, g: torch.Tensor, pattern: str = "abba"):
    """Embed spatial (h [M,M], g [M]^4) into the reference's spin-orbital layout
    (h [2M,2M] = diag(h,h); g [2M]^4 non-zero on the spin blocks named by `pattern`)."""
    M = h.shape[0]
    P = 2 * M
    hs = torch.zeros(P, P, dtype=torch.float64)
    hs[:M, :M] = h
    hs[M:, M:] = h
    gs = torch.zeros(P, P, P, P, dtype=torch.float64)
    for s in range(2):
        for t in range(2):
            blk = (s, t, t, s) if pattern == "abba" else (s, t, s, t)
            sl = tuple(slice(b * M, (b + 1) * M) for b in blk)
            gs[sl] = g
    return hs, gs
