// oo_probe — stand-alone driver for liboo_b200.so (no Python): measures the device peaks, checks
// one evaluation against a plain host restatement on small problems, and times the evaluation on
// large synthetic problems.  Usage:
//   oo_probe peaks
//   oo_probe check  M N [seed]
//   oo_probe time   M N [mloc] [reps] [dense]
//   oo_probe opt    M N
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../../include/oo_b200.h"

#define CK(x)                                                                      \
  do {                                                                             \
    int _r = (x);                                                                  \
    if (_r != 0) {                                                                 \
      fprintf(stderr, "FAIL %s -> %d: %s\n", #x, _r, oo_last_error());             \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)
#define CUK(x)                                                                     \
  do {                                                                             \
    cudaError_t _e = (x);                                                          \
    if (_e != cudaSuccess) {                                                       \
      fprintf(stderr, "CUDA FAIL %s: %s\n", #x, cudaGetErrorString(_e));           \
      exit(3);                                                                     \
    }                                                                              \
  } while (0)

using vec = std::vector<double>;

static void random_orthonormal(vec& U, int M, int N, std::mt19937_64& rng) {
  std::normal_distribution<double> nd(0.0, 1.0);
  U.assign((size_t)M * N, 0.0);
  for (auto& x : U) x = nd(rng);
  // modified Gram-Schmidt on columns
  for (int j = 0; j < N; ++j) {
    for (int i = 0; i < j; ++i) {
      double d = 0;
      for (int t = 0; t < M; ++t) d += U[(size_t)t * N + i] * U[(size_t)t * N + j];
      for (int t = 0; t < M; ++t) U[(size_t)t * N + j] -= d * U[(size_t)t * N + i];
    }
    double n = 0;
    for (int t = 0; t < M; ++t) n += U[(size_t)t * N + j] * U[(size_t)t * N + j];
    n = 1.0 / std::sqrt(n);
    for (int t = 0; t < M; ++t) U[(size_t)t * N + j] *= n;
  }
}

// A[t][a] = sum_{qrs,jkl} g[t,q,r,s] U[q,j] U[r,k] U[s,l] G[a,j,k,l]  (host, staged transform)
static void host_A0(const vec& g, const vec& G, const vec& U, int M, int N, vec& A) {
  const size_t M2 = (size_t)M * M, M3 = M2 * M;
  const size_t N2 = (size_t)N * N, N3 = N2 * N;
  vec T1(M3 * N), T2(M2 * N2), T3((size_t)M * N3);
  for (size_t pqr = 0; pqr < M3; ++pqr)
    for (int l = 0; l < N; ++l) {
      double s = 0;
      for (int x = 0; x < M; ++x) s += g[pqr * M + x] * U[(size_t)x * N + l];
      T1[pqr * N + l] = s;
    }
  for (size_t pq = 0; pq < M2; ++pq)
    for (int k = 0; k < N; ++k)
      for (int l = 0; l < N; ++l) {
        double s = 0;
        for (int r = 0; r < M; ++r) s += T1[(pq * M + r) * N + l] * U[(size_t)r * N + k];
        T2[(pq * N + k) * N + l] = s;
      }
  for (int p = 0; p < M; ++p)
    for (int j = 0; j < N; ++j)
      for (size_t kl = 0; kl < N2; ++kl) {
        double s = 0;
        for (int q = 0; q < M; ++q) s += T2[((size_t)p * M + q) * N2 + kl] * U[(size_t)q * N + j];
        T3[((size_t)p * N + j) * N2 + kl] = s;
      }
  A.assign((size_t)M * N, 0.0);
  for (int p = 0; p < M; ++p)
    for (int a = 0; a < N; ++a) {
      double s = 0;
      for (size_t jkl = 0; jkl < N3; ++jkl) s += T3[(size_t)p * N3 + jkl] * G[(size_t)a * N3 + jkl];
      A[(size_t)p * N + a] = s;
    }
}

// swap index slot 0 with slot m of a rank-4 tensor of extent n
static vec swap0(const vec& x, int n, int m) {
  if (m == 0) return x;
  vec y(x.size());
  size_t st[4] = {(size_t)n * n * n, (size_t)n * n, (size_t)n, 1};
  for (int a = 0; a < n; ++a)
    for (int b = 0; b < n; ++b)
      for (int c = 0; c < n; ++c)
        for (int d = 0; d < n; ++d) {
          int idx[4] = {a, b, c, d};
          std::swap(idx[0], idx[m]);
          y[idx[0] * st[0] + idx[1] * st[1] + idx[2] * st[2] + idx[3] * st[3]] =
              x[a * st[0] + b * st[1] + c * st[2] + d * st[3]];
        }
  return y;
}

struct Problem {
  int M, N;
  vec h, g, D, G, U;
};

static Problem make_problem(int M, int N, unsigned seed) {
  Problem P;
  P.M = M;
  P.N = N;
  std::mt19937_64 rng(seed);
  std::normal_distribution<double> nd(0.0, 1.0);
  const size_t M2 = (size_t)M * M, M3 = M2 * M, M4 = M3 * M;
  P.h.resize(M2);
  for (int p = 0; p < M; ++p)
    for (int q = 0; q <= p; ++q) P.h[(size_t)p * M + q] = P.h[(size_t)q * M + p] = nd(rng);
  vec X(M4);
  for (auto& x : X) x = nd(rng) * 0.1;
  P.g.resize(M4);
  for (int p = 0; p < M; ++p)
    for (int q = 0; q < M; ++q)
      for (int r = 0; r < M; ++r)
        for (int s = 0; s < M; ++s)
          P.g[p * M3 + q * M2 + (size_t)r * M + s] =
              X[p * M3 + q * M2 + (size_t)r * M + s] + X[q * M3 + p * M2 + (size_t)s * M + r] +
              X[r * M3 + s * M2 + (size_t)p * M + q] + X[s * M3 + r * M2 + (size_t)q * M + p];
  P.D.resize((size_t)N * N);
  for (auto& x : P.D) x = nd(rng);
  P.G.resize((size_t)N * N * N * N);
  for (auto& x : P.G) x = nd(rng);
  random_orthonormal(P.U, M, N, rng);
  return P;
}

static void host_energy_grad(const Problem& P, double& E, vec& grad) {
  const int M = P.M, N = P.N;
  vec A;
  grad.assign((size_t)M * N, 0.0);
  E = 0;
  for (int m = 0; m < 4; ++m) {
    host_A0(swap0(P.g, M, m), swap0(P.G, N, m), P.U, M, N, A);
    for (size_t i = 0; i < grad.size(); ++i) grad[i] += A[i];
    if (m == 0)
      for (size_t i = 0; i < A.size(); ++i) E += P.U[i] * A[i];
  }
  // one body: E1 = sum h_pq U_pi U_qj D_ij ; G1 = h U D^T + h^T U D
  vec UD((size_t)M * N), UDt((size_t)M * N);
  for (int q = 0; q < M; ++q)
    for (int a = 0; a < N; ++a) {
      double s0 = 0, s1 = 0;
      for (int j = 0; j < N; ++j) {
        s0 += P.U[(size_t)q * N + j] * P.D[(size_t)j * N + a];
        s1 += P.U[(size_t)q * N + j] * P.D[(size_t)a * N + j];
      }
      UD[(size_t)q * N + a] = s0;
      UDt[(size_t)q * N + a] = s1;
    }
  for (int t = 0; t < M; ++t)
    for (int a = 0; a < N; ++a) {
      double b1 = 0, b2 = 0;
      for (int q = 0; q < M; ++q) {
        b1 += P.h[(size_t)t * M + q] * UDt[(size_t)q * N + a];
        b2 += P.h[(size_t)q * M + t] * UD[(size_t)q * N + a];
      }
      grad[(size_t)t * N + a] += b1 + b2;
      E += P.U[(size_t)t * N + a] * b1;
    }
}

static double* to_dev(const vec& v) {
  double* d;
  CUK(cudaMalloc((void**)&d, v.size() * sizeof(double)));
  CUK(cudaMemcpy(d, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice));
  return d;
}

static int cmd_peaks() {
  double out[3];
  CK(oo_measure_peaks(0, (size_t)8 << 30, out));
  printf("{\"dmma_tflops\": %.3f, \"dfma_tflops\": %.3f, \"stream_read_gbs\": %.1f}\n", out[0],
         out[1], out[2]);
  return 0;
}

static int cmd_check(int M, int N, unsigned seed) {
  Problem P = make_problem(M, N, seed);
  double Eh;
  vec gh;
  host_energy_grad(P, Eh, gh);
  double* dh = to_dev(P.h);
  double* dg = to_dev(P.g);
  double* dD = to_dev(P.D);
  double* dG = to_dev(P.G);
  double sym[2];
  CK(oo_check_v4_symmetry(0, dg, M, sym));
  oo_ctx* ctx;
  CK(oo_create(0, M, N, 0, M, &ctx));
  CK(oo_set_integrals(ctx, dh, dg, OO_G_V4_SYMMETRIC));
  CK(oo_set_rdms(ctx, dD, dG));
  double Ed;
  vec gd((size_t)M * N);
  CK(oo_energy_grad_host(ctx, P.U.data(), &Ed, gd.data()));
  double num = 0, den = 0;
  for (size_t i = 0; i < gd.size(); ++i) {
    num += (gd[i] - gh[i]) * (gd[i] - gh[i]);
    den += gh[i] * gh[i];
  }
  const double rel = std::sqrt(num / den);
  // orth check: U = orth(V) with V random -> U^T U = I and U (V^T V)^(1/2) = V is implied
  vec V((size_t)M * N), Uo((size_t)M * N);
  std::mt19937_64 rng(seed + 7);
  std::normal_distribution<double> nd(0.0, 1.0);
  for (auto& x : V) x = nd(rng);
  double* dV = to_dev(V);
  double* dUo = to_dev(V);
  CK(oo_orth(ctx, dV, dUo));
  CK(oo_synchronize(ctx));
  CUK(cudaMemcpy(Uo.data(), dUo, Uo.size() * sizeof(double), cudaMemcpyDeviceToHost));
  double orth_err = 0, polar_err = 0;
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      double s = 0, w = 0;
      for (int t = 0; t < M; ++t) {
        s += Uo[(size_t)t * N + i] * Uo[(size_t)t * N + j];
        w += Uo[(size_t)t * N + i] * V[(size_t)t * N + j];  // U^T V must be symmetric (polar)
      }
      orth_err = std::max(orth_err, std::fabs(s - (i == j ? 1.0 : 0.0)));
      double wt = 0;
      for (int t = 0; t < M; ++t) wt += Uo[(size_t)t * N + j] * V[(size_t)t * N + i];
      polar_err = std::max(polar_err, std::fabs(w - wt));
    }
  printf("{\"cmd\": \"check\", \"M\": %d, \"N\": %d, \"v4_asym\": %.3e, \"E_host\": %.15e, "
         "\"E_dev\": %.15e, \"dE\": %.3e, \"grad_rel_err\": %.3e, \"orth_err\": %.3e, "
         "\"polar_sym_err\": %.3e}\n",
         M, N, sym[0], Eh, Ed, std::fabs(Eh - Ed), rel, orth_err, polar_err);
  const bool ok = std::fabs(Eh - Ed) <= 1e-10 * std::max(1.0, std::fabs(Eh)) && rel <= 1e-9 &&
                  orth_err < 1e-12 && polar_err < 1e-10;
  CK(oo_destroy(ctx));
  cudaFree(dh); cudaFree(dg); cudaFree(dD); cudaFree(dG); cudaFree(dV); cudaFree(dUo);
  if (!ok) {
    printf("CHECK FAILED\n");
    return 1;
  }
  return 0;
}

__global__ void fill_hash(double* p, size_t n, unsigned long long seed) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    unsigned long long z = (i + seed) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    p[i] = ((double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5) * 1e-2;
  }
}

static int cmd_time(int M, int N, int mloc, int reps, bool dense) {
  const size_t M3 = (size_t)M * M * M;
  const size_t gcount = (size_t)mloc * M3;
  printf("# allocating g shard: %.2f GB\n", gcount * 8.0 / 1e9);
  double* dg;
  CUK(cudaMalloc((void**)&dg, gcount * sizeof(double)));
  fill_hash<<<148 * 8, 256>>>(dg, gcount, 1234);
  CUK(cudaDeviceSynchronize());
  std::mt19937_64 rng(5);
  std::normal_distribution<double> nd(0.0, 1.0);
  vec h((size_t)M * M), D((size_t)N * N), G((size_t)N * N * N * N), U;
  for (auto& x : h) x = nd(rng);
  for (auto& x : D) x = nd(rng);
  for (auto& x : G) x = nd(rng);
  random_orthonormal(U, M, N, rng);
  double *dh = to_dev(h), *dD = to_dev(D), *dG = to_dev(G), *dU = to_dev(U);
  oo_ctx* ctx;
  CK(oo_create(0, M, N, 0, mloc, &ctx));
  CK(oo_set_integrals(ctx, dh, dg, OO_G_V4_SYMMETRIC));
  CK(oo_set_rdms(ctx, dD, dG));
  CK(oo_set_pair_symmetry(ctx, dense ? 0 : 1));
  const int nslab = oo_streamed_slabs(ctx);
  CK(oo_set_timing(ctx, 1));
  float ms[5], best[5] = {1e30f, 1e30f, 1e30f, 1e30f, 1e30f}, sum[5] = {0, 0, 0, 0, 0};
  for (int w = 0; w < 3; ++w) {
    CK(oo_energy_grad(ctx, dU, nullptr));
    CK(oo_synchronize(ctx));
  }
  for (int r = 0; r < reps; ++r) {
    CK(oo_energy_grad(ctx, dU, nullptr));
    CK(oo_last_timing(ctx, ms));
    for (int i = 0; i < 5; ++i) {
      best[i] = std::min(best[i], ms[i]);
      sum[i] += ms[i];
    }
  }
  const double bytes = (double)nslab * M * M * 8.0;
  const double fl1 = (double)nslab * (2.0 * M * M * N + 2.0 * M * N * N);  // K1 flops
  const double Npd = 8.0 * ((N + 7) / 8);
  const double flp = (double)nslab * (2.0 * M * M * Npd + 2.0 * M * Npd * Npd);
  printf("{\"cmd\": \"time\", \"mode\": \"%s\", \"slabs\": %d, \"M\": %d, \"N\": %d, \"mloc\": %d, \"reps\": %d, "
         "\"k1_ms_avg\": %.4f, \"k1_ms_min\": %.4f, \"prep_ms\": %.4f, \"tail_ms\": %.4f, "
         "\"unused\": %.4f, \"eval_ms_avg\": %.4f, \"eval_ms_min\": %.4f, "
         "\"k1_gbs\": %.1f, \"k1_tflops_alg\": %.2f, \"k1_tflops_padded\": %.2f, "
         "\"evals_per_s\": %.2f}\n",
         dense ? "dense" : "pair-symmetric", nslab, M, N, mloc, reps, sum[0] / reps, best[0], sum[1] / reps, sum[2] / reps, sum[3] / reps,
         sum[4] / reps, best[4], bytes / (sum[0] / reps * 1e-3) / 1e9,
         fl1 / (sum[0] / reps * 1e-3) / 1e12, flp / (sum[0] / reps * 1e-3) / 1e12,
         1e3 / (sum[4] / reps));
  CK(oo_destroy(ctx));
  cudaFree(dg); cudaFree(dh); cudaFree(dD); cudaFree(dG); cudaFree(dU);
  return 0;
}

static int cmd_opt(int M, int N) {
  Problem P = make_problem(M, N, 11);
  // make the problem better conditioned for a descent run: scale the two-body part down
  for (auto& x : P.g) x *= 0.05;
  // symmetric PSD-like D and V4-symmetric Gamma are not required for the mechanics
  double* dh = to_dev(P.h);
  double* dg = to_dev(P.g);
  double* dD = to_dev(P.D);
  double* dG = to_dev(P.G);
  oo_ctx* ctx;
  CK(oo_create(0, M, N, 0, M, &ctx));
  CK(oo_set_integrals(ctx, dh, dg, OO_G_V4_SYMMETRIC));
  CK(oo_set_rdms(ctx, dD, dG));
  vec U = P.U, hist(4096, 0.0);
  int niter = 0;
  double Ef = 0, bb = 0;
  auto t0 = std::chrono::steady_clock::now();
  CK(oo_optimize(ctx, U.data(), 1e-3, 1e-9, 2000, 0.8, hist.data(), (int)hist.size(), &niter, &Ef,
                 &bb));
  auto t1 = std::chrono::steady_clock::now();
  const double secs = std::chrono::duration<double>(t1 - t0).count();
  printf("{\"cmd\": \"opt\", \"M\": %d, \"N\": %d, \"n_iter\": %d, \"E0\": %.12f, \"E_final\": %.12f, "
         "\"bb\": %.4e, \"seconds\": %.4f, \"iters_per_s\": %.1f, \"launches\": %lld}\n",
         M, N, niter, hist[0], Ef, bb, secs, niter / secs, oo_launch_count(ctx));
  CK(oo_destroy(ctx));
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: oo_probe peaks | check M N [seed] | time M N [mloc] [reps] | opt M N\n");
    return 64;
  }
  const std::string cmd = argv[1];
  if (cmd == "peaks") return cmd_peaks();
  if (argc < 4) return 64;
  const int M = atoi(argv[2]), N = atoi(argv[3]);
  if (cmd == "check") return cmd_check(M, N, argc > 4 ? (unsigned)atoi(argv[4]) : 1u);
  if (cmd == "time")
    return cmd_time(M, N, argc > 4 ? atoi(argv[4]) : M, argc > 5 ? atoi(argv[5]) : 10,
                    argc > 6 && std::string(argv[6]) == "dense");
  if (cmd == "opt") return cmd_opt(M, N);
  return 64;
}
