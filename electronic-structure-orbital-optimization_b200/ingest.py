"""Spin-orbital -> spatial reduction of the reference's input tensors (host-side plumbing, torch).

The reference evaluates (base_opt_orb_solver.py:549-563)
    E = sum h_pq W_pi W_qj D_ij + sum g_pqrs W_pi W_qj W_rk W_sl Gamma_ijkl,   W = block_diag(U, U)
on spin-orbital tensors of extent P = 2M / Q = 2N (alpha block first, then beta).  Because W is
block diagonal, E splits into a sum over spin blocks (s1,s2,s3,s4) of g with the matching block of
Gamma.  For restricted integrals every non-zero block of g is the same spatial tensor g~ and
h = diag(h~, h~); then
    E = sum h~ U U D~ + sum g~ U U U U Gamma~,
    D~ = sum_s D[s,s],  Gamma~ = sum over the non-zero blocks of g of the same block of Gamma
which is 16x less data and 32x fewer flops.  The block pattern comes from qiskit-nature (not
available here), so it is *detected and verified* on the incoming tensor, never assumed.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass
from typing import List, Sequence, Tuple

import torch

Block = Tuple[int, int, int, int]


@dataclass
class SpinStructure:
    M: int
    blocks: List[Block]      # spin blocks of g that are non-zero (all equal to g_spatial)


def _blk(t: torch.Tensor, n: int, spins: Sequence[int]) -> torch.Tensor:
    return t[tuple(slice(s * n, (s + 1) * n) for s in spins)]


def reduce_integrals(h: torch.Tensor, g: torch.Tensor, rtol: float = 1e-12):
    """(h [2M,2M], g [2M]^4) -> (h~ [M,M], g~ [M]^4 contiguous, SpinStructure).

    Raises NotImplementedError when the tensors are not of the restricted block form."""
    if h.dim() != 2 or g.dim() != 4 or h.shape[0] != h.shape[1] or len(set(g.shape)) != 1 \
            or g.shape[0] != h.shape[0]:
        raise ValueError(f"expected h [P,P] and g [P,P,P,P], got {tuple(h.shape)} {tuple(g.shape)}")
    if h.dtype != torch.float64 or g.dtype != torch.float64:
        raise TypeError("integrals must be float64 (complex / lower precision is not supported)")
    P = h.shape[0]
    if P % 2:
        raise ValueError("spin-orbital tensors must have even extent")
    M = P // 2
    scale_h = float(h.abs().max()) or 1.0
    if float(_blk(h, M, (0, 1)).abs().max()) > rtol * scale_h or \
            float(_blk(h, M, (1, 0)).abs().max()) > rtol * scale_h:
        raise NotImplementedError("one-body integrals couple alpha and beta orbitals")
    h_sp = _blk(h, M, (0, 0))
    if float((_blk(h, M, (1, 1)) - h_sp).abs().max()) > rtol * scale_h:
        raise NotImplementedError("unrestricted one-body integrals (h_aa != h_bb) are not supported")
    scale_g = float(g.abs().max()) or 1.0
    blocks: List[Block] = []
    g_sp = None
    for spins in itertools.product((0, 1), repeat=4):
        b = _blk(g, M, spins)
        if float(b.abs().max()) <= rtol * scale_g:
            continue
        if g_sp is None:
            g_sp = b
        elif float((b - g_sp).abs().max()) > rtol * scale_g:
            raise NotImplementedError(
                "unrestricted two-body integrals (spin blocks differ) are not supported")
        blocks.append(tuple(spins))
    if g_sp is None:
        g_sp = _blk(g, M, (0, 0, 0, 0))
    return h_sp.contiguous(), g_sp.contiguous(), SpinStructure(M=M, blocks=blocks)


def block_mask(structure: SpinStructure) -> int:
    """Bit b = 8*s0 + 4*s1 + 2*s2 + s3 for every non-zero spin block (C-ABI convention)."""
    m = 0
    for s0, s1, s2, s3 in structure.blocks:
        m |= 1 << (8 * s0 + 4 * s1 + 2 * s2 + s3)
    return m


def reduce_integrals_device(h: torch.Tensor, g: torch.Tensor, rtol: float = 1e-12):
    """Same contract as reduce_integrals for CUDA tensors, with the P^4 scan, the block
    comparison and the extraction done by the library's kernels (oo_ingest_spin_g)."""
    import ctypes as C
    from . import _lib
    if not g.is_cuda:
        raise ValueError("reduce_integrals_device needs CUDA tensors")
    if h.dtype != torch.float64 or g.dtype != torch.float64:
        raise TypeError("integrals must be float64 (complex / lower precision is not supported)")
    P = h.shape[0]
    if h.dim() != 2 or g.dim() != 4 or tuple(g.shape) != (P,) * 4 or h.shape[1] != P or P % 2:
        raise ValueError(f"expected h [P,P] and g [P,P,P,P], got {tuple(h.shape)} {tuple(g.shape)}")
    M = P // 2
    scale_h = float(h.abs().max()) or 1.0
    if float(_blk(h, M, (0, 1)).abs().max()) > rtol * scale_h or \
            float(_blk(h, M, (1, 0)).abs().max()) > rtol * scale_h:
        raise NotImplementedError("one-body integrals couple alpha and beta orbitals")
    h_sp = _blk(h, M, (0, 0)).contiguous()
    if float((_blk(h, M, (1, 1)) - h_sp).abs().max()) > rtol * scale_h:
        raise NotImplementedError("unrestricted one-body integrals (h_aa != h_bb) are not supported")
    g = g.contiguous()
    g_sp = torch.empty(M, M, M, M, dtype=torch.float64, device=g.device)
    mask = C.c_uint(0)
    stats = (C.c_double * 2)()
    torch.cuda.current_stream(g.device).synchronize()
    lib = _lib.load()
    rc = lib.oo_ingest_spin_g(g.device.index, C.c_void_p(g.data_ptr()), M, float(rtol),
                              C.c_void_p(g_sp.data_ptr()), C.byref(mask), stats)
    if rc == -5:
        raise NotImplementedError(lib.oo_last_error().decode())
    _lib.check(rc)
    blocks = [tuple((b >> s) & 1 for s in (3, 2, 1, 0)) for b in range(16) if (mask.value >> b) & 1]
    return h_sp, g_sp, SpinStructure(M=M, blocks=blocks)


def reduce_rdms(oneRDM, twoRDM, structure: SpinStructure, weights=None):
    """Spin-sum (and state-average with `weights`) the reference's RDM arguments.

    oneRDM / twoRDM are tensors (ground state) or lists of tensors (excited states,
    opt_orb_eigensolver.py:149-169: E is linear in the RDMs, so sum_n w_n E(U; D_n, G_n) =
    E(U; sum_n w_n D_n, sum_n w_n G_n))."""
    ones = list(oneRDM) if isinstance(oneRDM, (list, tuple)) else [oneRDM]
    twos = list(twoRDM) if isinstance(twoRDM, (list, tuple)) else [twoRDM]
    if len(ones) != len(twos):
        raise ValueError("oneRDM and twoRDM lists differ in length")
    if weights is None:
        weights = [1.0] * len(ones)
    if len(weights) != len(ones):
        raise ValueError("number of weights does not match the number of states")
    D_sp = G_sp = None
    for w, D, G in zip(weights, ones, twos):
        if D.is_complex() or G.is_complex():
            raise NotImplementedError(
                "complex RDMs (base_opt_orb_solver.py:565-580) are not supported")
        if D.dtype != torch.float64 or G.dtype != torch.float64:
            raise TypeError("RDMs must be float64")
        Q = D.shape[0]
        if Q % 2 or tuple(G.shape) != (Q, Q, Q, Q) or tuple(D.shape) != (Q, Q):
            raise ValueError(f"bad RDM shapes {tuple(D.shape)} {tuple(G.shape)}")
        N = Q // 2
        d = _blk(D, N, (0, 0)) + _blk(D, N, (1, 1))
        gsum = torch.zeros(N, N, N, N, dtype=torch.float64, device=G.device)
        for spins in structure.blocks:
            gsum = gsum + _blk(G, N, spins)
        D_sp = float(w) * d if D_sp is None else D_sp + float(w) * d
        G_sp = float(w) * gsum if G_sp is None else G_sp + float(w) * gsum
    return D_sp.contiguous(), G_sp.contiguous()
