// Input preparation kernels: permutation-symmetry verification of the ERI tensor and the
// symmetrised / padded / permuted copies of the 2-RDM (k_prepare_gamma2 for the fused evaluation,
// k_prepare_gamma for the tiles path), spin-orbital ingest, pair-packing.
#pragma once
#include "oo_common.cuh"

namespace oo {

// V4 symmetry verification, tiled so that both sides of every comparison are read coalesced
// (the first version gathered three permuted elements per element: 274 ms at M=256).
//   MODE 0:  g[p,q,r,s] == g[q,p,s,r]   pairs (p<=q), tiles over (r,s): mirror tile transposed
//   MODE 1:  g[p,q,r,s] == g[r,s,p,q]   pairs (p<=r), tiles over (q,s): mirror tile transposed
// The third V4 element (03)(12) is the product of these two, so its deviation is bounded by the
// sum of theirs.  out[0] = max deviation, out[1] = max |g| (bit patterns of non-negative doubles:
// atomicMax on unsigned long long is order preserving).  grid (M(M+1)/2, T, T), block (32, 8).
template <int MODE>
__global__ void __launch_bounds__(256) k_v4_symmetry_tiles(const double* __restrict__ g, int M,
                                                            unsigned long long* out) {
  __shared__ double tile[32][33];
  // decode the pair index (a <= b)
  int a = 0, rem = blockIdx.x;
  {
    // rows of the upper triangle have M, M-1, ... entries
    const double Md = (double)M;
    a = (int)floor((2.0 * Md + 1.0 - sqrt((2.0 * Md + 1.0) * (2.0 * Md + 1.0) - 8.0 * rem)) * 0.5);
    a = max(0, min(a, M - 1));
    while (a > 0 && (long)a * M - (long)a * (a - 1) / 2 > rem) --a;
    while ((long)(a + 1) * M - (long)(a + 1) * a / 2 <= rem) ++a;
    rem -= (int)((long)a * M - (long)a * (a - 1) / 2);
  }
  const int b = a + rem;
  const int u0 = blockIdx.y * 32, v0 = blockIdx.z * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const size_t M2 = (size_t)M * M, M3 = M2 * M;
  double asym = 0.0, amax = 0.0;
  // mirror tile: element (v0+ty.., u0+tx)
  for (int i = ty; i < 32; i += 8) {
    const int vv = v0 + i, uu = u0 + tx;
    double x = 0.0;
    if (vv < M && uu < M) {
      // MODE 0: g[b][a][vv][uu]      MODE 1: g[b][vv][a][uu]
      x = MODE == 0 ? g[b * M3 + a * M2 + (size_t)vv * M + uu]
                    : g[b * M3 + (size_t)vv * M2 + (size_t)a * M + uu];
    }
    tile[i][tx] = x;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int uu = u0 + i, vv = v0 + tx;
    if (uu < M && vv < M) {
      // MODE 0: g[a][b][uu][vv]      MODE 1: g[a][uu][b][vv]
      const double x = MODE == 0 ? g[a * M3 + b * M2 + (size_t)uu * M + vv]
                                 : g[a * M3 + (size_t)uu * M2 + (size_t)b * M + vv];
      amax = fmax(amax, fabs(x));
      asym = fmax(asym, fabs(x - tile[tx][i]));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    asym = fmax(asym, __shfl_xor_sync(0xffffffffu, asym, o));
    amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  }
  if (tx == 0) {
    if (asym != 0.0) atomicMax(out + 0, (unsigned long long)__double_as_longlong(asym));
    if (amax != 0.0) atomicMax(out + 1, (unsigned long long)__double_as_longlong(amax));
  }
}

// Logical Gp[a][j][l*Np+k] = (1/4)(G[a,j,k,l] + G[j,a,l,k] + G[k,l,a,j] + G[l,k,j,a])  (V4 average),
// zero in the padding (j, k or l >= N).  grid N*Np, block Np*Np threads (looped).
// symmetrise = 0 selects one of the four slot layouts of the generic (no symmetry) gradient:
//   slot 0: Gp[a][j][l*Np+k] = G[a,j,k,l]      slot 1: Gp[a][i][l*Np+k] = G[i,a,k,l]
//   slot 2: Gp[a][l][j*Np+i] = G[i,j,a,l]      slot 3: Gp[a][k][j*Np+i] = G[i,j,k,a]
// (middle index = the plane index of T3, tile index = (row, col) of the K1 tile as stored).
__global__ void k_prepare_gamma(const double* __restrict__ G, double* __restrict__ Gp, int N, int Np,
                                int symmetrise, int slot = 0) {
  const int a = blockIdx.x / Np, j = blockIdx.x % Np;
  const int Np2 = Np * Np;
  const size_t N2 = (size_t)N * N, N3 = N2 * N;
  for (int e = threadIdx.x; e < Np2; e += blockDim.x) {
    const int l = e / Np, k = e - l * Np;
    double v = 0.0;
    if (j < N && k < N && l < N) {
      if (!symmetrise && slot == 1) v = G[j * N3 + a * N2 + (size_t)k * N + l];
      else if (!symmetrise && slot == 2) v = G[k * N3 + l * N2 + (size_t)a * N + j];
      else if (!symmetrise && slot == 3) v = G[k * N3 + l * N2 + (size_t)j * N + a];
      else v = G[a * N3 + j * N2 + (size_t)k * N + l];
      if (symmetrise) {
        v += G[j * N3 + a * N2 + (size_t)l * N + k];
        v += G[k * N3 + l * N2 + (size_t)a * N + j];
        v += G[l * N3 + k * N2 + (size_t)j * N + a];
        v *= 0.25;
      }
    }
    // storage: a fastest, Gp[(j*Np2 + e)*Np + a] (rows a >= N stay zero from the allocation):
    // the tail kernel reads all a of one (j,e) with unit stride; strides of Np^3 doubles between
    // the a-planes made every CTA hit the same L2 slices at the same time (measured 10x slower)
    Gp[((size_t)j * Np2 + e) * Np + a] = v;
  }
}

// 2-RDM in the layout of the fused evaluation (k_prepare_q): G2[c][a][e], c = the index that is
// contracted with U, a = the gradient column, e = e1*Np + e0 the position inside a K1 tile
// (tile[e1*Np + e0] = Y[e0][e1], Y = U^T g_slab U).  Zero in the padding.  `kind` selects what is
// summed over which orbital (see oo_k1.cuh / DESIGN.md section 1):
//   0  V4 own     G2[j][a][l*Np+k] = Gs[a,j,k,l]   Gs = V4 average of G   (row t from slab (t,q), c = q side)
//   1  V4 mirror  G2[j][a][k*Np+l] = Gs[a,j,k,l]   (row q from slab (t,q): transposed tile, c = t side)
//   2  slot 0     G2[j][a][l*Np+k] = G[a,j,k,l]    generic pass over g:    row t, c = q
//   3  slot 1     G2[i][a][l*Np+k] = G[i,a,k,l]    generic pass over g:    row q, c = t
//   4  slot 2     G2[l][a][j*Np+i] = G[i,j,a,l]    generic pass over g_pt: row r, c = s
//   5  slot 3     G2[k][a][j*Np+i] = G[i,j,k,a]    generic pass over g_pt: row s, c = r
// grid Np*Np (c, a), block 256 over e.
__global__ void k_prepare_gamma2(const double* __restrict__ G, double* __restrict__ G2, int N,
                                 int Np, int kind) {
  const int c = blockIdx.x / Np, a = blockIdx.x % Np;
  const int Np2 = Np * Np;
  const size_t N2 = (size_t)N * N, N3 = N2 * N;
  auto at = [&](int i0, int i1, int i2, int i3) {
    return G[i0 * N3 + i1 * N2 + (size_t)i2 * N + i3];
  };
  for (int e = threadIdx.x; e < Np2; e += blockDim.x) {
    const int e1 = e / Np, e0 = e - e1 * Np;
    double v = 0.0;
    if (c < N && a < N && e0 < N && e1 < N) {
      if (kind <= 1) {
        const int k = kind == 0 ? e0 : e1, l = kind == 0 ? e1 : e0, j = c;
        v = 0.25 * (at(a, j, k, l) + at(j, a, l, k) + at(k, l, a, j) + at(l, k, j, a));
      } else if (kind == 2) {
        v = at(a, c, e0, e1);
      } else if (kind == 3) {
        v = at(c, a, e0, e1);
      } else if (kind == 4) {
        v = at(e0, e1, a, c);
      } else {
        v = at(e0, e1, c, a);
      }
    }
    G2[((size_t)c * Np + a) * Np2 + e] = v;
  }
}

// Pair-packed storage: packed[i] = dense slab coord[i] (M*M doubles each, 16-byte aligned because M
// is even).  grid (nslab, chunks), 256 threads, 16-byte copies.
__global__ void __launch_bounds__(256) k_pack_slabs(const double* __restrict__ dense,
                                                    double* __restrict__ packed,
                                                    const int* __restrict__ coord, int M) {
  const size_t n2 = (size_t)M * M / 2;
  const double2* src = reinterpret_cast<const double2*>(dense + (size_t)coord[blockIdx.x] * M * M);
  double2* dst = reinterpret_cast<double2*>(packed + (size_t)blockIdx.x * M * M);
  for (size_t i = (size_t)blockIdx.y * blockDim.x + threadIdx.x; i < n2;
       i += (size_t)gridDim.y * blockDim.x)
    dst[i] = src[i];
}

}  // namespace oo

namespace oo {

// ---------------------------------------------------------------------------------------------
// Spin-orbital -> spatial ingest of the reference's tensors (extent P = 2M / Q = 2N, alpha block
// first): base_opt_orb_solver.py:549 makes W = block_diag(U,U), so the energy only sees g and
// Gamma through matching spin blocks.  Block id b = 8*s0 + 4*s1 + 2*s2 + s3.
// ---------------------------------------------------------------------------------------------

// stats[b] = max |g[block b]| (bit pattern, atomicMax);  grid-stride over the P^4 elements.
__global__ void k_spin_block_maxabs(const double* __restrict__ g, int M, unsigned long long* stats) {
  const size_t P = 2 * (size_t)M, total = P * P * P * P;
  __shared__ unsigned long long s_max[16];
  if (threadIdx.x < 16) s_max[threadIdx.x] = 0ull;
  __syncthreads();
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int s = (int)(idx % P), r = (int)((idx / P) % P), q = (int)((idx / (P * P)) % P),
              p = (int)(idx / (P * P * P));
    const int b = 8 * (p >= M) + 4 * (q >= M) + 2 * (r >= M) + (s >= M);
    const double v = fabs(g[idx]);
    if (v != 0.0) atomicMax(&s_max[b], (unsigned long long)__double_as_longlong(v));
  }
  __syncthreads();
  if (threadIdx.x < 16 && s_max[threadIdx.x]) atomicMax(stats + threadIdx.x, s_max[threadIdx.x]);
}

// out[tl][q][r][s] = g[block ref](t0+tl, q, r, s) for tl < mloc, written with extent Mp >= M in
// the last three indices (Mp = M+1 pads an odd M for TMA; the padding is left untouched, i.e.
// zero from the caller's allocation), and diff[b] = max |g[block b] - g[block ref]| over the same
// rows for the blocks in mask.
__global__ void k_spin_block_extract(const double* __restrict__ g, int M, int ref, unsigned mask,
                                     int t0, int mloc, int Mp, double* __restrict__ out,
                                     unsigned long long* diff) {
  const size_t P = 2 * (size_t)M, total = (size_t)mloc * M * M * M;
  const size_t P3 = P * P * P, P2 = P * P;
  double dmax[16];
#pragma unroll
  for (int b = 0; b < 16; ++b) dmax[b] = 0.0;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int s = (int)(idx % M), r = (int)((idx / M) % M), q = (int)((idx / ((size_t)M * M)) % M),
              tl = (int)(idx / ((size_t)M * M * M));
    const int p = t0 + tl;
    auto at = [&](int b) {
      return g[(size_t)(p + M * ((b >> 3) & 1)) * P3 + (size_t)(q + M * ((b >> 2) & 1)) * P2 +
               (size_t)(r + M * ((b >> 1) & 1)) * P + (size_t)(s + M * (b & 1))];
    };
    const double v = at(ref);
    out[(((size_t)tl * Mp + q) * Mp + r) * Mp + s] = v;
#pragma unroll
    for (int b = 0; b < 16; ++b)
      if ((mask >> b) & 1u) dmax[b] = fmax(dmax[b], fabs(at(b) - v));
  }
#pragma unroll
  for (int b = 0; b < 16; ++b) {
    double v = dmax[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0 && v != 0.0)
      atomicMax(diff + b, (unsigned long long)__double_as_longlong(v));
  }
}

// Spatial, state-weighted RDMs from spin-orbital ones:
//   D~[i][j]       = sum_n w_n (D_n[i,j] + D_n[N+i,N+j])
//   G~[i][j][k][l] = sum_n w_n sum_{b in mask} G_n[block b](i,j,k,l)
struct RdmSpinParams {
  const double* D[8];
  const double* G[8];
  double w[8];
  int nstates, N;
  unsigned mask;
};
__global__ void k_rdm_spin_sum(const RdmSpinParams p, double* __restrict__ Dout,
                               double* __restrict__ Gout) {
  const int N = p.N;
  const size_t Q = 2 * (size_t)N, Q2 = Q * Q, Q3 = Q2 * Q;
  const size_t N4 = (size_t)N * N * N * N;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < N4;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int l = (int)(idx % N), k = (int)((idx / N) % N), j = (int)((idx / ((size_t)N * N)) % N),
              i = (int)(idx / ((size_t)N * N * N));
    double acc = 0.0;
    for (int n = 0; n < p.nstates; ++n) {
      double s = 0.0;
      for (int b = 0; b < 16; ++b)
        if ((p.mask >> b) & 1u)
          s += p.G[n][(size_t)(i + N * ((b >> 3) & 1)) * Q3 + (size_t)(j + N * ((b >> 2) & 1)) * Q2 +
                      (size_t)(k + N * ((b >> 1) & 1)) * Q + (size_t)(l + N * (b & 1))];
      acc = fma(p.w[n], s, acc);
    }
    Gout[idx] = acc;
    if (idx < (size_t)N * N) {
      const int a = (int)(idx / N), c = (int)(idx % N);
      double d = 0.0;
      for (int n = 0; n < p.nstates; ++n)
        d = fma(p.w[n], p.D[n][(size_t)a * Q + c] + p.D[n][(size_t)(a + N) * Q + (c + N)], d);
      Dout[idx] = d;
    }
  }
}

}  // namespace oo
