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


def _check_h(h: torch.Tensor, M: int, rtol: float) -> torch.Tensor:
    scale_h = float(h.abs().max()) or 1.0
    if float(_blk(h, M, (0, 1)).abs().max()) > rtol * scale_h or \
            float(_blk(h, M, (1, 0)).abs().max()) > rtol * scale_h:
        raise NotImplementedError("one-body integrals couple alpha and beta orbitals")
    h_sp = _blk(h, M, (0, 0)).contiguous()
    if float((_blk(h, M, (1, 1)) - h_sp).abs().max()) > rtol * scale_h:
        raise NotImplementedError("unrestricted one-body integrals (h_aa != h_bb) are not supported")
    return h_sp


def _check_shapes(h: torch.Tensor, g: torch.Tensor) -> int:
    if h.dtype != torch.float64 or g.dtype != torch.float64:
        raise TypeError("integrals must be float64 (complex / lower precision is not supported)")
    P = h.shape[0]
    if h.dim() != 2 or g.dim() != 4 or tuple(g.shape) != (P,) * 4 or h.shape[1] != P or P % 2:
        raise ValueError(f"expected h [P,P] and g [P,P,P,P], got {tuple(h.shape)} {tuple(g.shape)}")
    return P // 2


def reduce_integrals_device(h: torch.Tensor, g: torch.Tensor, rtol: float = 1e-12,
                            t0: int = 0, mloc: int = None, pad_even: bool = False):
    """Same contract as reduce_integrals for CUDA tensors, with the P^4 scan, the block
    comparison and the extraction done by the library's kernels (oo_ingest_spin_g_rows).

    t0 / mloc select the rows of the first index one rank of a multi-GPU run keeps (the M^4
    spatial tensor is never materialised); pad_even writes an odd M straight into the zero-padded
    even extent the engine needs (no second copy).  Returns (h~ [M,M], g~ rows, SpinStructure)."""
    import ctypes as C
    from . import _lib
    if not g.is_cuda:
        raise ValueError("reduce_integrals_device needs CUDA tensors")
    M = _check_shapes(h, g)
    mloc = M - t0 if mloc is None else mloc
    h_sp = _check_h(h, M, rtol)
    g = g.contiguous()
    Mp = M + (M % 2) if pad_even else M
    rows = mloc + (1 if (Mp != M and t0 + mloc == M) else 0)
    alloc = torch.zeros if Mp != M else torch.empty
    g_sp = alloc(rows, Mp, Mp, Mp, dtype=torch.float64, device=g.device)
    mask = C.c_uint(0)
    stats = (C.c_double * 2)()
    torch.cuda.current_stream(g.device).synchronize()
    lib = _lib.load()
    rc = lib.oo_ingest_spin_g_rows(g.device.index, C.c_void_p(g.data_ptr()), M, float(rtol),
                                   int(t0), int(mloc), int(Mp), C.c_void_p(g_sp.data_ptr()),
                                   C.byref(mask), stats)
    if rc == -5:
        raise NotImplementedError(lib.oo_last_error().decode())
    _lib.check(rc)
    blocks = [tuple((b >> s) & 1 for s in (3, 2, 1, 0)) for b in range(16) if (mask.value >> b) & 1]
    return h_sp, g_sp, SpinStructure(M=M, blocks=blocks)


def reduce_integrals_rows_host(h: torch.Tensor, g: torch.Tensor, t0: int, mloc: int,
                               rtol: float = 1e-12):
    """Host tensors, one shard: (h~ [M,M], g~[t0:t0+mloc] contiguous, SpinStructure) reading only
    the shard's rows of every spin block (the ranks of a multi-GPU run together read the tensor
    once).  The non-zero blocks are identified on the shard's rows; the caller must make sure
    that all ranks agree (PartialUnitaryProjectionOptimizer all-reduces the block mask)."""
    M = _check_shapes(h, g)
    h_sp = _check_h(h, M, rtol)
    rows = slice(t0, t0 + mloc)
    maxabs = {}
    for spins in itertools.product((0, 1), repeat=4):
        maxabs[spins] = float(_blk(g, M, spins)[rows].abs().max())
    scale_g = max(maxabs.values()) or 1.0
    blocks = [sp for sp, v in maxabs.items() if v > rtol * scale_g]
    ref = blocks[0] if blocks else (0, 0, 0, 0)
    g_rows = _blk(g, M, ref)[rows].contiguous()
    for sp in blocks[1:]:
        if float((_blk(g, M, sp)[rows] - g_rows).abs().max()) > rtol * scale_g:
            raise NotImplementedError(
                "unrestricted two-body integrals (spin blocks differ) are not supported")
    return h_sp, g_rows, SpinStructure(M=M, blocks=[tuple(b) for b in blocks])


def v4_asymmetry_rows(g: torch.Tensor, M: int, block: Block, t0: int, mloc: int):
    """(max |g - g o pi| over the V4 permutations, max |g|) of rows [t0, t0+mloc) of one spin
    block of the FULL spin-orbital tensor, wherever it lives (views only, mloc*M^3 temporaries):
    what a rank of a sharded run can verify about its shard."""
    blk = _blk(g, M, block)
    rows = slice(t0, t0 + mloc)
    own = blk[rows]
    asym = 0.0
    for perm in ((1, 0, 3, 2), (2, 3, 0, 1), (3, 2, 1, 0)):
        asym = max(asym, float((own - blk.permute(*perm)[rows]).abs().max()))
    return asym, float(own.abs().max())


class SpatialIntegrals:
    """Extension input format: integrals that are already spatial, and possibly already sharded.

    The reference passes h [2M,2M] and g [2M]^4 spin-orbital tensors
    (base_opt_orb_solver.py:89-90); at M=256 that tensor would take 550 GB, at M=400 3.3 TB, so
    BASELINE.json's configs 4 and 5 cannot be expressed in it.  An instance of this class can be
    handed to PartialUnitaryProjectionOptimizer.compute_optimal_rotation as `two_body_integrals`
    (with `one_body_integrals` = the spatial h [M,M]):

        h        [M,M] spatial one-body integrals
        g        this rank's rows [t0, t0+mloc) of the spatial tensor g[p,q,r,s] (the reference's
                 alpha-beta-beta-alpha block), dense [mloc,M,M,M] or pair-packed [count,M,M]
                 (distributed.pair_slab_list order) when packed=True
        blocks   which spin blocks of the reference's tensor are non-zero ('abba': (s,t,t,s) as
                 qiskit-nature produces them, 'abab': (s,t,s,t)); decides which blocks of the
                 spin-orbital 2-RDM enter the spin sum
        v4_symmetric  the caller's assertion g[pqrs]=g[qpsr]=g[rspq] (true for real orbitals); a
                 shard cannot be verified locally
    """

    def __init__(self, g: torch.Tensor, M: int, t0: int = 0, mloc: int = None,
                 packed: bool = False, pattern: str = "abba", v4_symmetric: bool = True,
                 g_pair_transposed: torch.Tensor = None):
        if pattern not in ("abba", "abab"):
            raise ValueError("pattern must be 'abba' or 'abab'")
        self.g, self.M, self.t0 = g, int(M), int(t0)
        self.mloc = self.M - self.t0 if mloc is None else int(mloc)
        self.packed, self.pattern, self.v4_symmetric = bool(packed), pattern, bool(v4_symmetric)
        self.g_pair_transposed = g_pair_transposed
        if packed and not v4_symmetric:
            raise ValueError("pair-packed storage needs a V4-symmetric tensor")
        if not v4_symmetric and g_pair_transposed is None and self.mloc != self.M:
            raise ValueError("a sharded non-symmetric tensor needs g_pair_transposed")

    @property
    def structure(self) -> SpinStructure:
        blocks = [(s, t, t, s) if self.pattern == "abba" else (s, t, s, t)
                  for s in (0, 1) for t in (0, 1)]
        return SpinStructure(M=self.M, blocks=blocks)


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
