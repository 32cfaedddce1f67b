"""OrbitalEngine — thin Python owner of one liboo_b200 context (one GPU, one shard of g).

PyTorch is used only for device memory and streams; every numerical step runs in the CUDA library.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _dev_f64(t: torch.Tensor, device: torch.device) -> torch.Tensor:
    if t.dtype != torch.float64:
        raise TypeError(f"float64 tensor required, got {t.dtype}")
    return t.to(device).contiguous()


class OrbitalEngine:
    """Energy / gradient / retraction engine for spatial integrals (h [M,M], g [mloc,M,M,M]).

    `t0, mloc` select the shard of g's first index held by this GPU (default: everything)."""

    def __init__(self, M: int, N: int, device="cuda:0", t0: int = 0, mloc: Optional[int] = None):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("OrbitalEngine runs on CUDA devices only (there is no CPU fallback)")
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device is available; the B200 path has no CPU fallback")
        self.M_user, self.N = int(M), int(N)
        self.M = self.M_user + (self.M_user % 2)      # TMA needs 16-byte row strides: pad odd M
        self.t0 = int(t0)
        self.mloc_user = self.M_user - self.t0 if mloc is None else int(mloc)
        self.mloc = self.mloc_user
        if self.M != self.M_user and self.t0 + self.mloc_user == self.M_user:
            self.mloc += 1                             # the padded (zero) orbital joins the last shard
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        self._ctx = C.c_void_p()
        _lib.check(self.lib.oo_create(idx, self.M, self.N, self.t0, self.mloc, C.byref(self._ctx)))
        self._keep = {}
        self._out = torch.zeros(self.M * self.N + 1, dtype=torch.float64, device=self.device)
        self.world = 1

    # -- lifetime -----------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self.lib.oo_destroy(self._ctx)
            self._ctx = C.c_void_p()
        self._keep = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- inputs -------------------------------------------------------------------------------
    def _inputs_ready(self) -> None:
        """Tensors handed to the library may still be in flight on torch's current stream."""
        torch.cuda.current_stream(self.device).synchronize()

    def _pad_h(self, h):
        if self.M == self.M_user:
            return h
        out = torch.zeros(self.M, self.M, dtype=torch.float64, device=h.device)
        out[:self.M_user, :self.M_user] = h
        return out

    def _pad_g(self, g):
        if self.M == self.M_user or tuple(g.shape) == (self.mloc,) + (self.M,) * 3:
            return g
        out = torch.zeros(self.mloc, self.M, self.M, self.M, dtype=torch.float64, device=g.device)
        out[:self.mloc_user, :self.M_user, :self.M_user, :self.M_user] = g
        return out

    def _pad_u(self, U):
        if self.M == self.M_user:
            return U
        out = torch.zeros(self.M, self.N, dtype=torch.float64, device=U.device)
        out[:self.M_user] = U
        return out

    def check_v4_symmetry(self, g_full: torch.Tensor) -> Tuple[float, float]:
        """(max |g - g∘pi| over the three V4 permutations, max |g|) of a full device tensor."""
        g_full = _dev_f64(g_full, self.device)
        out = (C.c_double * 2)()
        self._inputs_ready()
        _lib.check(self.lib.oo_check_v4_symmetry(self.device.index, _ptr(g_full), g_full.shape[0],
                                                 out))
        return float(out[0]), float(out[1])

    def set_integrals(self, h: torch.Tensor, g: torch.Tensor, assume_v4_symmetric: bool = False,
                      sym_rtol: float = 1e-11, allow_generic: bool = True,
                      g_pair_transposed: Optional[torch.Tensor] = None) -> None:
        """h [M,M]; g [mloc,M,M,M] (this shard's rows).  The tensors are used in place (no copy)
        when already on the device, contiguous and M is even.

        `g_pair_transposed` [mloc,M,M,M] = g.permute(2,3,0,1)[t0:t0+mloc] of the FULL tensor selects
        the generic (no permutational symmetry) path explicitly; a sharded engine needs it from
        the caller because the rows of the transposed tensor live in other shards."""
        if tuple(h.shape) != (self.M_user, self.M_user):
            raise ValueError(f"h must be [{self.M_user},{self.M_user}], got {tuple(h.shape)}")
        padded = (self.mloc,) + (self.M,) * 3       # odd M: already zero-padded by the ingest
        if tuple(g.shape) not in ((self.mloc_user,) + (self.M_user,) * 3, padded):
            raise ValueError(f"g must be [{self.mloc_user},{self.M_user},{self.M_user},"
                             f"{self.M_user}], got {tuple(g.shape)}")
        h = self._pad_h(_dev_f64(h, self.device))
        g = self._pad_g(_dev_f64(g, self.device))
        self.generic = False
        if g_pair_transposed is not None:
            if tuple(g_pair_transposed.shape) != (self.mloc_user,) + (self.M_user,) * 3:
                raise ValueError("g_pair_transposed must have the shape of the g shard")
            g_pt = self._pad_g(_dev_f64(g_pair_transposed, self.device))
            self._keep["h"], self._keep["g"], self._keep["g_pt"] = h, g, g_pt
            self._inputs_ready()
            _lib.check(self.lib.oo_set_integrals_generic(self._ctx, _ptr(h), _ptr(g), _ptr(g_pt)))
            self.generic = True
            return
        if not assume_v4_symmetric:
            if self.mloc != self.M:
                raise ValueError("a sharded g cannot be verified locally; verify the full tensor "
                                 "before sharding and pass assume_v4_symmetric=True")
            asym, gmax = self.check_v4_symmetry(g)
            if asym > sym_rtol * max(gmax, 1e-300):
                if not allow_generic:
                    raise NotImplementedError(
                        f"two-body integrals are not V4-symmetric (max asymmetry {asym:.3e}, "
                        f"max |g| {gmax:.3e}); the one-pass gradient needs g[pqrs]=g[qpsr]=g[rspq]")
                # generic path: second dense pass over the pair-transposed tensor (2x memory)
                g_pt = g.permute(2, 3, 0, 1).contiguous()
                self._keep["h"], self._keep["g"], self._keep["g_pt"] = h, g, g_pt
                self._inputs_ready()
                _lib.check(self.lib.oo_set_integrals_generic(self._ctx, _ptr(h), _ptr(g),
                                                             _ptr(g_pt)))
                self.generic = True
                return
        self._keep["h"], self._keep["g"] = h, g
        self._keep.pop("g_pt", None)
        self._inputs_ready()
        _lib.check(self.lib.oo_set_integrals(self._ctx, _ptr(h), _ptr(g), _lib.OO_G_V4_SYMMETRIC))

    def set_integrals_packed(self, h: torch.Tensor, g_packed: torch.Tensor) -> None:
        """Pair-packed storage of a V4-symmetric tensor: g_packed [count, M, M] holds the slabs
        g[t, q, :, :] listed by `distributed.pair_slab_list(M, t0, mloc)` (one of every pair
        {(t,q),(q,t)}), half the memory of the dense shard; kernel traffic is unchanged.  The
        symmetry is asserted by the caller (it cannot be verified from half of the tensor)."""
        if self.M != self.M_user:
            raise ValueError("pair-packed storage needs an even M (pad the integrals first)")
        count = self.streamed_slabs_pair()
        if tuple(h.shape) != (self.M, self.M):
            raise ValueError(f"h must be [{self.M},{self.M}], got {tuple(h.shape)}")
        if tuple(g_packed.shape) != (count, self.M, self.M):
            raise ValueError(f"g_packed must be [{count},{self.M},{self.M}], got "
                             f"{tuple(g_packed.shape)}")
        h, g = _dev_f64(h, self.device), _dev_f64(g_packed, self.device)
        self.generic = False
        self._keep["h"], self._keep["g"] = h, g
        self._keep.pop("g_pt", None)
        self._inputs_ready()
        _lib.check(self.lib.oo_set_integrals(self._ctx, _ptr(h), _ptr(g),
                                             _lib.OO_G_V4_SYMMETRIC | _lib.OO_G_PAIR_PACKED))

    def streamed_slabs_pair(self) -> int:
        """Number of slabs of this shard in pair-packed storage (= streamed in pair mode)."""
        n = int(self.lib.oo_pair_slab_list(self.M, self.t0, self.mloc, None, 0))
        if n < 0:
            _lib.check(n)
        return n

    def pack_pair_slabs(self, g: torch.Tensor) -> torch.Tensor:
        """Gather the pair-packed shard [count, M, M] from a dense shard [mloc, M, M, M] on the
        device (oo_pack_pair_slabs)."""
        if self.M != self.M_user:
            raise ValueError("pair-packed storage needs an even M (pad the integrals first)")
        if tuple(g.shape) != (self.mloc, self.M, self.M, self.M):
            raise ValueError(f"g must be [{self.mloc},{self.M},{self.M},{self.M}]")
        g = _dev_f64(g, self.device)
        out = torch.empty(self.streamed_slabs_pair(), self.M, self.M, dtype=torch.float64,
                          device=self.device)
        self._inputs_ready()
        _lib.check(self.lib.oo_pack_pair_slabs(self._ctx, _ptr(g), _ptr(out)))
        _lib.check(self.lib.oo_synchronize(self._ctx))
        return out

    def set_rdms(self, D: torch.Tensor, G: torch.Tensor) -> None:
        """Spatial spin-summed (and state-weighted) D [N,N], Gamma [N,N,N,N]."""
        N = self.N
        if tuple(D.shape) != (N, N) or tuple(G.shape) != (N, N, N, N):
            raise ValueError(f"RDM shapes {tuple(D.shape)} {tuple(G.shape)} do not match N={N}")
        D, G = _dev_f64(D, self.device), _dev_f64(G, self.device)
        self._inputs_ready()
        _lib.check(self.lib.oo_set_rdms(self._ctx, _ptr(D), _ptr(G)))
        _lib.check(self.lib.oo_synchronize(self._ctx))   # D, G may be freed by the caller now

    def set_rdms_spin(self, oneRDMs, twoRDMs, weights, block_mask: int) -> None:
        """Spin-orbital RDMs (lists of [2N,2N] / [2N]^4 tensors) -> spatial, weighted, symmetrised
        copies inside the library, all on the device (k_rdm_spin_sum + k_prepare_gamma)."""
        Q = 2 * self.N
        ones = [_dev_f64(d, self.device) for d in oneRDMs]
        twos = [_dev_f64(g, self.device) for g in twoRDMs]
        for d, g in zip(ones, twos):
            if tuple(d.shape) != (Q, Q) or tuple(g.shape) != (Q,) * 4:
                raise ValueError(f"bad spin-orbital RDM shapes {tuple(d.shape)} {tuple(g.shape)}")
        if len(ones) != len(twos) or len(weights) != len(ones):
            raise ValueError("number of states / weights mismatch")
        if len(ones) > 8:
            raise NotImplementedError("more than 8 states per call")
        n = len(ones)
        Dp = (C.c_void_p * n)(*[d.data_ptr() for d in ones])
        Gp = (C.c_void_p * n)(*[g.data_ptr() for g in twos])
        w = (C.c_double * n)(*[float(x) for x in weights])
        self._inputs_ready()
        _lib.check(self.lib.oo_set_rdms_spin(self._ctx, Dp, Gp, w, n, int(block_mask)))
        _lib.check(self.lib.oo_synchronize(self._ctx))

    def set_pair_symmetry(self, enable: bool) -> None:
        """Stream one slab of every pair {(t,q),(q,t)} (default) or every slab of the shard."""
        _lib.check(self.lib.oo_set_pair_symmetry(self._ctx, 1 if enable else 0))
        self._out.zero_()
        self._inputs_ready()

    def streamed_slabs(self) -> int:
        """M x M slabs one evaluation reads from HBM on this GPU."""
        n = int(self.lib.oo_streamed_slabs(self._ctx))
        if n < 0:
            _lib.check(n)
        return n

    # -- multi-GPU -----------------------------------------------------------------------------
    def attach_comm(self, unique_id: bytes, rank: int, world: int) -> None:
        buf = (C.c_char * 128).from_buffer_copy(unique_id)
        _lib.check(self.lib.oo_comm_init(self._ctx, buf, rank, world))
        self.world = world

    def peer_export(self) -> bytes:
        """64-byte CUDA IPC handle of this GPU's exchange buffer (fused all-reduce)."""
        buf = (C.c_char * 64)()
        _lib.check(self.lib.oo_peer_export(self._ctx, buf))
        return bytes(buf)

    def peer_attach(self, handles: bytes, rank: int, world: int) -> None:
        buf = (C.c_char * (64 * world)).from_buffer_copy(handles)
        _lib.check(self.lib.oo_peer_attach(self._ctx, buf, rank, world))
        self.world = world

    def set_peer_timeout(self, seconds: float) -> None:
        """How long the fused all-reduce waits for a peer before it poisons the result (NaN) and
        the next synchronous call raises (default 20 s)."""
        _lib.check(self.lib.oo_set_peer_timeout_ms(self._ctx, float(seconds) * 1e3))

    def peer_status(self) -> int:
        rc = int(self.lib.oo_peer_status(self._ctx))
        if rc < 0:
            _lib.check(rc)
        return rc

    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = (C.c_char * 128)()
        _lib.check(_lib.load().oo_nccl_unique_id(buf))
        return bytes(buf)

    # -- evaluation ----------------------------------------------------------------------------
    def energy_grad(self, U: torch.Tensor, allreduce: bool = True):
        """(E 0-dim tensor, dE/dU [M,N]) on the device, asynchronous w.r.t. the host."""
        U = self._pad_u(_dev_f64(U, self.device))
        self._inputs_ready()
        fn = self.lib.oo_energy_grad_allreduce if (allreduce and self.world > 1) \
            else self.lib.oo_energy_grad
        _lib.check(fn(self._ctx, _ptr(U), _ptr(self._out)))
        _lib.check(self.lib.oo_synchronize(self._ctx))
        MN = self.M * self.N
        grad = self._out[:MN].view(self.M, self.N)[:self.M_user].clone()
        return self._out[MN].clone(), grad

    def _host_u(self, U) -> np.ndarray:
        U = np.ascontiguousarray(U, dtype=np.float64)
        if self.M == self.M_user:
            if U.shape != (self.M, self.N):
                raise ValueError(f"U must be [{self.M_user},{self.N}], got {U.shape}")
            return U
        Uh = np.zeros((self.M, self.N), dtype=np.float64)
        Uh[:self.M_user] = U
        return Uh

    def energy_grad_host(self, U: np.ndarray):
        """Host-buffer entry point: (E float, dE/dU ndarray); H2D/D2H inside the call."""
        Uh = self._host_u(U)
        E = C.c_double()
        grad = np.empty((self.M, self.N), dtype=np.float64)
        _lib.check(self.lib.oo_energy_grad_host(self._ctx, Uh.ctypes.data_as(C.c_void_p),
                                                C.byref(E), grad.ctypes.data_as(C.c_void_p)))
        return float(E.value), grad[:self.M_user]

    def submit_host(self, U: np.ndarray, slot: int) -> None:
        """First half of a pipelined host-buffer evaluation (oo_eval_submit): returns at once."""
        Uh = self._host_u(U)
        _lib.check(self.lib.oo_eval_submit(self._ctx, Uh.ctypes.data_as(C.c_void_p), int(slot)))

    def wait_host(self, slot: int, want_grad: bool = True):
        """Second half (oo_eval_wait): (E float, dE/dU ndarray or None) of the slot."""
        E = C.c_double()
        grad = np.empty((self.M, self.N), dtype=np.float64) if want_grad else None
        gp = grad.ctypes.data_as(C.c_void_p) if want_grad else C.c_void_p(0)
        _lib.check(self.lib.oo_eval_wait(self._ctx, int(slot), C.byref(E), gp))
        return float(E.value), (grad[:self.M_user] if want_grad else None)

    def energies_host(self, Us) -> np.ndarray:
        """Energies of a sequence of partial unitaries through the pipelined host-buffer path
        (two slots in flight): what the finite-difference gradient of pupo.py:105-127 needs."""
        out = np.empty(len(Us), dtype=np.float64)
        pending = []
        for i, U in enumerate(Us):
            if len(pending) == 2:
                j, s = pending.pop(0)
                out[j] = self.wait_host(s, want_grad=False)[0]
            self.submit_host(U, i & 1)
            pending.append((i, i & 1))
        for j, s in pending:
            out[j] = self.wait_host(s, want_grad=False)[0]
        return out

    def transform(self, U: torch.Tensor):
        """Rotated integrals (h' [N,N], g' [N,N,N,N]) on the device (shard-partial when sharded)."""
        U = self._pad_u(_dev_f64(U, self.device))
        N = self.N
        h_rot = torch.zeros(N, N, dtype=torch.float64, device=self.device)
        g_rot = torch.zeros(N, N, N, N, dtype=torch.float64, device=self.device)
        self._inputs_ready()
        _lib.check(self.lib.oo_transform(self._ctx, _ptr(U), _ptr(h_rot), _ptr(g_rot)))
        if self.world > 1:
            _lib.check(self.lib.oo_allreduce(self._ctx, _ptr(h_rot), N * N))
            _lib.check(self.lib.oo_allreduce(self._ctx, _ptr(g_rot), N ** 4))
        _lib.check(self.lib.oo_synchronize(self._ctx))
        return h_rot, g_rot

    def orth(self, V: torch.Tensor) -> torch.Tensor:
        V = self._pad_u(_dev_f64(V, self.device))
        out = torch.empty_like(V)
        self._inputs_ready()
        _lib.check(self.lib.oo_orth(self._ctx, _ptr(V), _ptr(out)))
        _lib.check(self.lib.oo_synchronize(self._ctx))
        return out[:self.M_user]

    def bb_update(self, iteration, U_cur, U_prev, G_cur, G_prev, stepsize: float):
        """compute_updated_partial_unitary: returns (U_next [M,N] device tensor, new step size)."""
        U_cur = self._pad_u(_dev_f64(U_cur, self.device))
        G_cur = self._pad_u(_dev_f64(G_cur, self.device))
        U_prev = None if U_prev is None else self._pad_u(_dev_f64(U_prev, self.device))
        G_prev = None if G_prev is None else self._pad_u(_dev_f64(G_prev, self.device))
        alpha = torch.tensor([float(stepsize)], dtype=torch.float64, device=self.device)
        out = torch.empty_like(U_cur)
        self._inputs_ready()
        _lib.check(self.lib.oo_bb_update(self._ctx, int(iteration), _ptr(U_cur), _ptr(U_prev),
                                         _ptr(G_cur), _ptr(G_prev), _ptr(alpha), _ptr(out)))
        _lib.check(self.lib.oo_synchronize(self._ctx))
        return out[:self.M_user], float(alpha.item())

    def optimize(self, U0, bb0: float, tol: float, maxiter: int, decay: float = 0.8,
                 callback=None):
        """The whole inner loop on the device.  Returns dict(U ndarray, energy, n_iter, E_hist,
        stepsize).  `callback(iteration, energy)` is delivered live (at most 4 iterations late,
        after U has already advanced on the device) with the reference's arguments, on the
        calling thread; an exception raised by it stops the device loop and is re-raised here.
        With several GPUs a slow callback stalls this rank's enqueueing and therefore its peers
        (they wait inside the fused all-reduce, see set_peer_timeout)."""
        raised = []

        def trampoline(it, e, _user):
            # ctypes would print and swallow an exception raised here; keep it, ask the device
            # loop to stop, and re-raise once oo_optimize has returned (the reference's callback
            # exceptions leave compute_optimal_rotation, pupo.py:193-194)
            if raised:
                return
            try:
                callback(int(it), float(e))
            except BaseException as exc:          # noqa: BLE001  (KeyboardInterrupt included)
                raised.append(exc)
                self.lib.oo_request_stop(self._ctx)

        cb_c = _lib.CALLBACK_T(trampoline) if callback else C.cast(None, _lib.CALLBACK_T)
        _lib.check(self.lib.oo_set_callback(self._ctx, cb_c, None))
        Uh = np.zeros((self.M, self.N), dtype=np.float64)
        Uh[:self.M_user] = np.asarray(U0, dtype=np.float64)
        cap = max(int(maxiter), 0) + 8
        hist = np.zeros(cap, dtype=np.float64)
        n_iter, E, bb = C.c_int(), C.c_double(), C.c_double()
        rc = self.lib.oo_optimize(self._ctx, Uh.ctypes.data_as(C.c_void_p), float(bb0),
                                  float(tol), int(maxiter), float(decay),
                                  hist.ctypes.data_as(C.c_void_p), cap, C.byref(n_iter),
                                  C.byref(E), C.byref(bb))
        self.lib.oo_set_callback(self._ctx, C.cast(None, _lib.CALLBACK_T), None)
        if raised:
            raise raised[0]
        _lib.check(rc)
        ns, jac = C.c_int(), C.c_int()
        _lib.check(self.lib.oo_retraction_stats(self._ctx, C.byref(ns), C.byref(jac)))
        return {"U": Uh[:self.M_user].copy(), "energy": float(E.value), "n_iter": int(n_iter.value),
                "E_hist": hist, "stepsize": float(bb.value),
                "newton_schulz_iterations": int(ns.value), "jacobi_fallbacks": int(jac.value)}

    # -- measurement ---------------------------------------------------------------------------
    def use_stream(self, stream: Optional["torch.cuda.Stream"]) -> None:
        """Run all library work on a torch stream (so torch.cuda.Event timing sees it); None
        returns to a library-owned stream."""
        self._keep["stream"] = stream
        handle = C.c_void_p(0 if stream is None else stream.cuda_stream)
        _lib.check(self.lib.oo_set_stream(self._ctx, handle))

    def set_timing(self, on: bool) -> None:
        _lib.check(self.lib.oo_set_timing(self._ctx, 1 if on else 0))

    def last_timing(self):
        ms = (C.c_float * 5)()
        _lib.check(self.lib.oo_last_timing(self._ctx, ms))
        return [float(x) for x in ms]

    def launch_count(self) -> int:
        return int(self.lib.oo_launch_count(self._ctx))

    def enqueue_energy_grad(self, U_dev: torch.Tensor, allreduce: bool = True) -> None:
        """Asynchronous evaluation into the engine's output buffer (bench inner loop)."""
        fn = self.lib.oo_energy_grad_allreduce if (allreduce and self.world > 1) \
            else self.lib.oo_energy_grad
        _lib.check(fn(self._ctx, _ptr(U_dev), _ptr(self._out)))

    def synchronize(self) -> None:
        _lib.check(self.lib.oo_synchronize(self._ctx))


def measure_peaks(device_index: int = 0, nbytes: int = 8 << 30):
    """(DMMA TFLOP/s, DFMA TFLOP/s, streaming-read GB/s) measured on the device."""
    out = (C.c_double * 3)()
    _lib.check(_lib.load().oo_measure_peaks(device_index, nbytes, out))
    return float(out[0]), float(out[1]), float(out[2])
