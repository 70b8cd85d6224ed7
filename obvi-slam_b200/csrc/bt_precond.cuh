// Block-tridiagonal preconditioner for the reduced camera system, factored by block cyclic reduction.
//
// Keyframe poses are grouped into super-blocks of 16 consecutive variable poses (96 x 96 scalars).  Points are
// tracked over a few consecutive keyframes, so almost all of the Schur complement lies in the block-tridiagonal
// part T of that partition (object couplings between far-apart keyframes are what is left out).  T is factored
// ONCE per LM iteration by cyclic reduction -- log2(#super-blocks) levels of independent dense 96 x 96 operations
// (Gauss-Jordan inverses + small GEMMs) that fill the whole GPU -- and applied inside the persistent PCG kernel.
// With M = T the PCG converges in a handful of iterations where block-Jacobi needs > 1000 on a 2000-keyframe chain.
#pragma once

#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ba_kernels.cuh"

namespace obvi {

constexpr int kSbPoses = 16;          // poses per super-block
constexpr int kB = 6 * kSbPoses;      // 96
constexpr int kBB = kB * kB;

// Scatter the scaled + damped reduced matrix (full BSR, scalar rows contiguous) into the diagonal super-blocks D
// and the sub-diagonal couplings C0[I] = T[I+1][I].  One warp per pose block row.  D / C0 must be zeroed before;
// padding rows of the last super-block get a unit diagonal.
__global__ void bt_assemble_kernel(int nf, int nsb, const uint32_t* __restrict__ sf_ptr, const uint32_t* __restrict__ sf_col,
                                   const double* __restrict__ Sf, double* __restrict__ D, double* __restrict__ C0) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= nsb * kSbPoses) return;
  const int I = i / kSbPoses, li = i % kSbPoses;
  if (i >= nf) {  // padding pose: identity
    if (lane < 6) D[(size_t)I * kBB + (size_t)(li * 6 + lane) * kB + li * 6 + lane] = 1.0;
    return;
  }
  const uint32_t p0 = sf_ptr[i], nb = sf_ptr[i + 1] - p0;
  const double* row = Sf + (size_t)p0 * 36;
  const uint32_t total = nb * 36;
  for (uint32_t t = lane; t < total; t += 32) {
    const uint32_t a = t / (nb * 6), rem = t - a * nb * 6, k = rem / 6, c = rem - 6 * k;
    const int j = (int)sf_col[p0 + k];
    const int Jb = j / kSbPoses, lj = j % kSbPoses;
    if (Jb == I) D[(size_t)I * kBB + (size_t)(li * 6 + a) * kB + lj * 6 + c] = row[t];
    else if (Jb == I - 1) C0[(size_t)(I - 1) * kBB + (size_t)(li * 6 + a) * kB + lj * 6 + c] = row[t];
  }
}

// Gauss-Jordan inverse (no pivoting: the blocks are SPD) of 96 x 96 matrices.  One CTA of 1024 threads per matrix;
// thread (ty, tx) keeps its 3 x 3 elements (rows ty + 32 m, columns tx + 32 n) in REGISTERS for the whole
// elimination.  Per pivot only the pivot row and column travel through shared memory (double-buffered), so there
// is a single block barrier per pivot.
constexpr int kInvThreads = 1024;
__global__ void __launch_bounds__(kInvThreads) bt_invert_kernel(const int* __restrict__ idx, const double* __restrict__ D,
                                                                 double* __restrict__ Dinv, double* __restrict__ scalars) {
  __shared__ double prow[2][kB], pcol[2][kB];
  const int blk = idx[blockIdx.x];
  const double* src = D + (size_t)blk * kBB;
  double* dst = Dinv + (size_t)blk * kBB;
  const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
  double v[3][3];
#pragma unroll
  for (int m = 0; m < 3; m++)
#pragma unroll
    for (int n = 0; n < 3; n++) v[m][n] = src[(size_t)(ty + 32 * m) * kB + tx + 32 * n];
  bool bad = false;
  // publish pivot row / column 0
  if (ty == 0) {
#pragma unroll
    for (int n = 0; n < 3; n++) prow[0][tx + 32 * n] = v[0][n];
  }
  if (tx == 0) {
#pragma unroll
    for (int m = 0; m < 3; m++) pcol[0][ty + 32 * m] = v[m][0];
  }
  __syncthreads();
  for (int p = 0; p < kB; p++) {
    const int buf = p & 1;
    const double piv = prow[buf][p];
    if (!(piv > 0.0)) bad = true;
    const double d = 1.0 / piv;
    double f[3], g[3];
#pragma unroll
    for (int m = 0; m < 3; m++) { f[m] = pcol[buf][ty + 32 * m]; g[m] = prow[buf][tx + 32 * m] * d; }
#pragma unroll
    for (int m = 0; m < 3; m++)
#pragma unroll
      for (int n = 0; n < 3; n++) {
        const int i = ty + 32 * m, j = tx + 32 * n;
        v[m][n] = (i == p) ? ((j == p) ? d : g[n]) : ((j == p) ? -f[m] * d : v[m][n] - f[m] * g[n]);
      }
    // publish the next pivot row / column from the updated registers
    const int q = p + 1;
    if (q < kB) {
      const int nb = q & 1;
      const int qb = q >> 5;  // which of the thread's 3 rows / columns (selected without dynamic register indexing)
      if ((q & 31) == ty) {
#pragma unroll
        for (int n = 0; n < 3; n++) prow[nb][tx + 32 * n] = qb == 0 ? v[0][n] : (qb == 1 ? v[1][n] : v[2][n]);
      }
      if ((q & 31) == tx) {
#pragma unroll
        for (int m = 0; m < 3; m++) pcol[nb][ty + 32 * m] = qb == 0 ? v[m][0] : (qb == 1 ? v[m][1] : v[m][2]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int m = 0; m < 3; m++)
#pragma unroll
    for (int n = 0; n < 3; n++) dst[(size_t)(ty + 32 * m) * kB + tx + 32 * n] = v[m][n];
  if (bad && threadIdx.x == 0) atomicAdd(&scalars[SC_BT_FAIL], 1.0);
}

// C = beta C + alpha1 op(A1) op(B1) [+ alpha2 op(A2) op(B2)], all 96 x 96 row-major.  One CTA per task.
struct GemmTask {
  double* C;
  const double *A1, *B1, *A2, *B2;
  double alpha1, alpha2, beta;
  int tA1, tB1, tA2, tB2;
};
__global__ void __launch_bounds__(256) bt_gemm_kernel(const GemmTask* __restrict__ tasks) {
  extern __shared__ double sm[];
  double* As = sm;             // As[m][k]
  double* Bs = sm + kBB;       // Bs[k][n]
  const GemmTask T = tasks[blockIdx.x];
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  double acc[6][6];
#pragma unroll
  for (int m = 0; m < 6; m++)
#pragma unroll
    for (int n = 0; n < 6; n++) acc[m][n] = 0.0;
  for (int term = 0; term < 2; term++) {
    const double* A = term ? T.A2 : T.A1;
    const double* Bm = term ? T.B2 : T.B1;
    if (!A) break;
    const int tA = term ? T.tA2 : T.tA1, tB = term ? T.tB2 : T.tB1;
    const double alpha = term ? T.alpha2 : T.alpha1;
    __syncthreads();
    for (int t = threadIdx.x; t < kBB; t += 256) {
      const int r = t / kB, c = t - r * kB;
      const double a = A[t], b = Bm[t];
      if (tA) As[c * kB + r] = a; else As[t] = a;
      if (tB) Bs[c * kB + r] = b; else Bs[t] = b;
    }
    __syncthreads();
    double part[6][6];
#pragma unroll
    for (int m = 0; m < 6; m++)
#pragma unroll
      for (int n = 0; n < 6; n++) part[m][n] = 0.0;
    for (int k = 0; k < kB; k++) {
      double a[6], b[6];
#pragma unroll
      for (int m = 0; m < 6; m++) a[m] = As[(ty + 16 * m) * kB + k];
#pragma unroll
      for (int n = 0; n < 6; n++) b[n] = Bs[k * kB + tx + 16 * n];
#pragma unroll
      for (int m = 0; m < 6; m++)
#pragma unroll
        for (int n = 0; n < 6; n++) part[m][n] += a[m] * b[n];
    }
#pragma unroll
    for (int m = 0; m < 6; m++)
#pragma unroll
      for (int n = 0; n < 6; n++) acc[m][n] += alpha * part[m][n];
  }
#pragma unroll
  for (int m = 0; m < 6; m++)
#pragma unroll
    for (int n = 0; n < 6; n++) {
      double* c = T.C + (size_t)(ty + 16 * m) * kB + tx + 16 * n;
      *c = (T.beta != 0.0 ? T.beta * *c : 0.0) + acc[m][n];
    }
}

// ---------------------------------------------------------------------------------------------------------
// PCG on S~ y = b~ with the block-tridiagonal preconditioner.  Persistent cooperative kernel, 16 warps per CTA.
// Vectors are padded to nsb * 96 entries (zeros beyond 6 nf).
struct BtApply {
  int nsb, nlev;
  const double *Dinv, *GaT, *GcT;
  double *w, *z;   // reduced right-hand side (scratch) and the preconditioned vector
};

__device__ __forceinline__ void bt_forward_block(const BtApply& P, int a, int i, int j, int lane, int wib) {
  // w_a -= GaT[i] w_i + GcT[j] w_j   (i: eliminated right neighbour of a, j: eliminated left neighbour; -1 = none)
  double acc[6] = {0, 0, 0, 0, 0, 0};
  const int r0 = 6 * wib;
  if (i >= 0) {
    const double* M = P.GaT + (size_t)i * kBB;
    const double* v = P.w + (size_t)i * kB;
    for (int c = lane; c < kB; c += 32) {
      const double vc = v[c];
#pragma unroll
      for (int m = 0; m < 6; m++) acc[m] += M[(size_t)(r0 + m) * kB + c] * vc;
    }
  }
  if (j >= 0) {
    const double* M = P.GcT + (size_t)j * kBB;
    const double* v = P.w + (size_t)j * kB;
    for (int c = lane; c < kB; c += 32) {
      const double vc = v[c];
#pragma unroll
      for (int m = 0; m < 6; m++) acc[m] += M[(size_t)(r0 + m) * kB + c] * vc;
    }
  }
#pragma unroll
  for (int m = 0; m < 6; m++) acc[m] = warp_sum(acc[m]);
  if (lane < 6) {
    const double v = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : lane == 3 ? acc[3] : lane == 4 ? acc[4] : acc[5];
    P.w[(size_t)a * kB + r0 + lane] -= v;
  }
}
// z_i = Dinv_i w_i - GaT[i]^T z_a - GcT[i]^T z_c.  All 16 warps of the CTA take part: warp w handles the 6 rows
// k = 6w .. 6w+5 of the three matrices for all 96 outputs (lane owns outputs lane, lane+32, lane+64; loads are
// coalesced over the outputs and all independent), the 16 partial sums are combined through shared memory.
// Also accumulates r . z over the block's rows into *rz (when r != nullptr).
__device__ __forceinline__ void bt_backward_block(const BtApply& P, int i, int a, int c, double (*part)[kB],
                                                  const double* r, int n, double* rz) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const double* Di = P.Dinv + (size_t)i * kBB;
  const double* wi = P.w + (size_t)i * kB;
  double acc[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int kk = 0; kk < 6; kk++) {
    const int k = 6 * w + kk;
    const double wk = wi[k];
#pragma unroll
    for (int m = 0; m < 3; m++) acc[m] += Di[(size_t)k * kB + lane + 32 * m] * wk;
  }
  if (a >= 0) {
    const double* M = P.GaT + (size_t)i * kBB;
    const double* za = P.z + (size_t)a * kB;
#pragma unroll
    for (int kk = 0; kk < 6; kk++) {
      const int k = 6 * w + kk;
      const double zk = za[k];
#pragma unroll
      for (int m = 0; m < 3; m++) acc[m] -= M[(size_t)k * kB + lane + 32 * m] * zk;
    }
  }
  if (c >= 0) {
    const double* M = P.GcT + (size_t)i * kBB;
    const double* zc = P.z + (size_t)c * kB;
#pragma unroll
    for (int kk = 0; kk < 6; kk++) {
      const int k = 6 * w + kk;
      const double zk = zc[k];
#pragma unroll
      for (int m = 0; m < 3; m++) acc[m] -= M[(size_t)k * kB + lane + 32 * m] * zk;
    }
  }
  __syncthreads();  // `part` may still be read by the previous call
#pragma unroll
  for (int m = 0; m < 3; m++) part[w][lane + 32 * m] = acc[m];
  __syncthreads();
  if (threadIdx.x < kB) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < kPcgThreads / 32; q++) s += part[q][threadIdx.x];
    const int row = i * kB + threadIdx.x;
    P.z[row] = s;
    if (r != nullptr) {
      double d = row < n ? r[row] * s : 0.0;
      d = warp_sum(d);
      if (lane == 0 && d != 0.0) atomicAdd(rz, d);
    }
  }
}

// z = T^-1 w (w is overwritten).  Called by every thread of the grid.  Levels that still have more than kNarrow
// active super-blocks are spread over the grid (one grid barrier each); the narrow top of the reduction tree is
// done by CTA 0 alone with block barriers, which removes most of the grid-wide barriers from the PCG iteration.
constexpr int kNarrow = 4;
__device__ void bt_apply(cg::grid_group& grid, const BtApply& P, double (*part)[kB], const double* r, int n, double* rz) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int lsplit = 0;  // first level handled by CTA 0 alone
  while (lsplit < P.nlev && (P.nsb + (1 << lsplit) - 1) / (1 << lsplit) > kNarrow) lsplit++;
  for (int l = 0; l < lsplit; l++) {
    const int s = 1 << l, nl = (P.nsb + s - 1) / s;
    const int nsurv = (nl + 1) / 2;
    for (int t = blockIdx.x; t < nsurv; t += gridDim.x) {
      const int k = 2 * t;
      bt_forward_block(P, k * s, (k + 1 < nl) ? (k + 1) * s : -1, (k >= 1) ? (k - 1) * s : -1, lane, wib);
    }
    grid.sync();
  }
  if (blockIdx.x == 0) {
    for (int l = lsplit; l < P.nlev; l++) {
      const int s = 1 << l, nl = (P.nsb + s - 1) / s;
      for (int k = 0; k < nl; k += 2) bt_forward_block(P, k * s, (k + 1 < nl) ? (k + 1) * s : -1, (k >= 1) ? (k - 1) * s : -1, lane, wib);
      __threadfence_block();
      __syncthreads();
    }
    bt_backward_block(P, 0, -1, -1, part, r, n, rz);
    __threadfence_block();
    __syncthreads();
    for (int l = P.nlev - 1; l >= lsplit; l--) {
      const int s = 1 << l, nl = (P.nsb + s - 1) / s;
      for (int k = 1; k < nl; k += 2) {
        bt_backward_block(P, k * s, (k - 1) * s, (k + 1 < nl) ? (k + 1) * s : -1, part, r, n, rz);
      }
      __threadfence_block();
      __syncthreads();
    }
  }
  grid.sync();
  for (int l = lsplit - 1; l >= 0; l--) {
    const int s = 1 << l, nl = (P.nsb + s - 1) / s;
    const int nel = nl / 2;
    for (int t = blockIdx.x; t < nel; t += gridDim.x) {
      const int k = 2 * t + 1;
      bt_backward_block(P, k * s, (k - 1) * s, (k + 1 < nl) ? (k + 1) * s : -1, part, r, n, rz);
    }
    grid.sync();
  }
}

__global__ void __launch_bounds__(kPcgThreads) pcg_bt_kernel(int nf, const uint32_t* __restrict__ sf_ptr,
                                                             const uint32_t* __restrict__ sf_col,
                                                             const double* __restrict__ Sf, const double* __restrict__ rhs,
                                                             BtApply P, double* __restrict__ y, double* r, double* p, double* q,
                                                             double* acc /*4x4*/, int max_iter, double tol,
                                                             double* __restrict__ scalars) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double red[3][kPcgThreads / 32];
  __shared__ double part[kPcgThreads / 32][kB];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int gw = (blockIdx.x * kPcgThreads + threadIdx.x) >> 5;
  const int GW = (gridDim.x * kPcgThreads) >> 5;
  const int gt = blockIdx.x * kPcgThreads + threadIdx.x, GT = gridDim.x * kPcgThreads;
  const int n = 6 * nf, npad = P.nsb * kB;
  double* z = P.z;
  double* w = P.w;

  auto block_acc = [&](double a0, double a1, double a2, double* dst) {
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
    if (lane == 0) { red[0][wib] = a0; red[1][wib] = a1; red[2][wib] = a2; }
    __syncthreads();
    if (threadIdx.x < 3) {
      double s = 0.0;
      for (int i = 0; i < kPcgThreads / 32; i++) s += red[threadIdx.x][i];
      if (s != 0.0) atomicAdd(&dst[threadIdx.x], s);
    }
    __syncthreads();
  };

  if (blockIdx.x == 0 && threadIdx.x < 16) acc[threadIdx.x] = 0.0;
  {
    double bb = 0.0;
    for (int i = gt; i < npad; i += GT) {
      const double rv = i < n ? rhs[i] : 0.0;
      if (i < n) { y[i] = 0.0; r[i] = rv; }
      w[i] = rv;
      bb += rv * rv;
    }
    grid.sync();
    block_acc(0.0, bb, 0.0, acc);
  }
  grid.sync();
  bt_apply(grid, P, part, r, n, &acc[0]);   // acc[0] += r . z (fused into the back-substitution)
  for (int i = gt; i < n; i += GT) p[i] = z[i];
  grid.sync();
  double rho = ((volatile double*)acc)[0];
  const double bb = ((volatile double*)acc)[1];
  int it = 0, brk = 0;
  double rr = bb;
  if (((volatile double*)scalars)[SC_BT_FAIL] != 0.0) brk = 2;  // factorisation failed: the host falls back to block-Jacobi
  if (bb > 0.0 && brk == 0) {
    while (it < max_iter) {
      double* A = acc + 4 * ((it + 1) & 3);
      double* Zc = acc + 4 * ((it + 3) & 3);
      if (blockIdx.x == 0 && threadIdx.x < 4) Zc[threadIdx.x] = 0.0;
      double pq = 0.0;
      for (int i = gw; i < nf; i += GW) {
        const uint32_t p0 = sf_ptr[i], nb = sf_ptr[i + 1] - p0;
        const uint32_t len = nb * 6;
        const double* row = Sf + (size_t)p0 * 36;
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0;
        for (uint32_t e0 = lane; e0 < len; e0 += 128) {
          // 4 entries per lane per trip, every load issued before the first use
          uint32_t ee[4]; double xv[4], sv[4][6];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            ee[u] = e0 + 32 * u;
            const bool in = ee[u] < len;
            const uint32_t e = in ? ee[u] : 0u;
            const uint32_t k = e / 6, c = e - 6 * k;
            xv[u] = in ? p[6 * sf_col[p0 + k] + c] : 0.0;
#pragma unroll
            for (int a = 0; a < 6; a++) sv[u][a] = in ? row[a * len + e] : 0.0;
          }
#pragma unroll
          for (int u = 0; u < 4; u++) {
            a0 += sv[u][0] * xv[u]; a1 += sv[u][1] * xv[u]; a2 += sv[u][2] * xv[u];
            a3 += sv[u][3] * xv[u]; a4 += sv[u][4] * xv[u]; a5 += sv[u][5] * xv[u];
          }
        }
        a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3); a4 = warp_sum(a4); a5 = warp_sum(a5);
        if (lane < 6) {
          const double qv = lane == 0 ? a0 : lane == 1 ? a1 : lane == 2 ? a2 : lane == 3 ? a3 : lane == 4 ? a4 : a5;
          q[6 * i + lane] = qv;
          pq += qv * p[6 * i + lane];
        }
      }
      block_acc(pq, 0.0, 0.0, A);
      grid.sync();
      const double pqs = ((volatile double*)A)[0];
      if (!(pqs > 0.0)) { brk = 1; break; }
      const double alpha = rho / pqs;
      double r2 = 0.0;
      for (int i = gt; i < n; i += GT) {
        y[i] += alpha * p[i];
        const double rv = r[i] - alpha * q[i];
        r[i] = rv; w[i] = rv;
        r2 += rv * rv;
      }
      block_acc(0.0, 0.0, r2, A);
      grid.sync();
      rr = ((volatile double*)A)[2];
      it++;
      if (rr <= tol * tol * bb) break;
      bt_apply(grid, P, part, r, n, &A[1]);   // A[1] += r . z
      const double rho_new = ((volatile double*)A)[1];
      const double beta = rho_new / rho;
      rho = rho_new;
      for (int i = gt; i < n; i += GT) p[i] = z[i] + beta * p[i];
      grid.sync();
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    scalars[SC_PCG_IT] = (double)it;
    scalars[SC_PCG_RES] = bb > 0.0 ? sqrt(rr / bb) : 0.0;
    scalars[SC_PCG_BB] = bb;
    scalars[SC_PCG_BREAK] = (double)brk;
  }
}

}  // namespace obvi
