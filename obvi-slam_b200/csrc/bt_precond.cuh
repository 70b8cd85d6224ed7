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
// and the sub-diagonal couplings C0[I] = T[I+1][I].  One CTA per pose block row.  D / C0 must be zeroed before;
// padding rows of the last super-block get a unit diagonal.
__global__ void bt_assemble_kernel(int nf, int nsb, const uint32_t* __restrict__ sf_ptr, const uint32_t* __restrict__ sf_col,
                                   const double* __restrict__ Sf, double* __restrict__ D, double* __restrict__ C0) {
  const int i = blockIdx.x;        // one CTA per pose block row
  const int lane = threadIdx.x;
  if (i >= nsb * kSbPoses) return;
  const int I = i / kSbPoses, li = i % kSbPoses;
  if (i >= nf) {  // padding pose: identity
    if (lane < 6) D[(size_t)I * kBB + (size_t)(li * 6 + lane) * kB + li * 6 + lane] = 1.0;
    return;
  }
  const uint32_t p0 = sf_ptr[i], nb = sf_ptr[i + 1] - p0;
  const double* row = Sf + (size_t)p0 * 36;
  const uint32_t total = nb * 36;
  for (uint32_t t = lane; t < total; t += blockDim.x) {
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

// ---- second-generation factorisation kernels (default; the two above are kept selectable with OBVI_BT=v1) ----
//
// bt_invert8_kernel: BLOCKED Gauss-Jordan (block size 8, no pivoting) of a 96 x 96 SPD matrix, one CTA of 1024 threads,
// thread (ty, tx) keeps its 3 x 3 elements in registers.  Per block pivot: (1) the 8 pivot rows and 8 pivot columns go
// to shared memory, (2) warp 0 inverts the 8 x 8 pivot block with shuffles (two elements per lane), (3) everybody scales
// the row panel by the inverse, (4) one rank-8 update of the registers.  12 block pivots x 3 barriers instead of 96
// scalar pivots each with its own barrier and per-element select logic.
__global__ void __launch_bounds__(kInvThreads) bt_invert8_kernel(const int* __restrict__ idx, const double* __restrict__ D,
                                                                  double* __restrict__ Dinv, double* __restrict__ scalars) {
  constexpr int b = 8;
  __shared__ double rowp[b][kB];    // raw pivot rows A[P, :]
  __shared__ double rowq[b][kB];    // scaled pivot rows  Pinv A[P, :]  (columns in P: Pinv itself)
  __shared__ double colp[2][kB][b + 1];  // raw pivot columns A[:, P] (double-buffered: read in step 4, rewritten in the next step 1)
  __shared__ double pinv[b][b];
  __shared__ int s_bad;
  const int blk = idx[blockIdx.x];
  const double* src = D + (size_t)blk * kBB;
  double* dst = Dinv + (size_t)blk * kBB;
  const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
  double v[3][3];
#pragma unroll
  for (int m = 0; m < 3; m++)
#pragma unroll
    for (int n = 0; n < 3; n++) v[m][n] = src[(size_t)(ty + 32 * m) * kB + tx + 32 * n];
  if (threadIdx.x == 0) s_bad = 0;
  for (int pb = 0; pb < kB; pb += b) {
    const int pm = pb >> 5, po = pb & 31;   // the pivot rows are the thread rows ty + 32 pm with ty in [po, po + 8)
    // (1) publish the panels
    if (ty >= po && ty < po + b) {
#pragma unroll
      for (int n = 0; n < 3; n++) rowp[ty - po][tx + 32 * n] = pm == 0 ? v[0][n] : (pm == 1 ? v[1][n] : v[2][n]);
    }
    if (tx >= po && tx < po + b) {
#pragma unroll
      for (int m = 0; m < 3; m++) colp[(pb >> 3) & 1][ty + 32 * m][tx - po] = pm == 0 ? v[m][0] : (pm == 1 ? v[m][1] : v[m][2]);
    }
    __syncthreads();
    // (2) warp 0: Gauss-Jordan on the 8 x 8 pivot block, lane (r = lane / 4, c0 = 2 (lane % 4)) holds [r][c0], [r][c0 + 1]
    if (ty == 0) {
      const int r = tx >> 2, c0 = 2 * (tx & 3);
      double e0 = rowp[r][pb + c0], e1 = rowp[r][pb + c0 + 1];
      bool bad = false;
#pragma unroll
      for (int p = 0; p < b; p++) {
        const int src_p = (p << 2) | (p >> 1);          // lane holding [p][p]
        const double pv0 = __shfl_sync(0xffffffffu, e0, src_p), pv1 = __shfl_sync(0xffffffffu, e1, src_p);
        const double piv = (p & 1) ? pv1 : pv0;
        if (!(piv > 0.0)) bad = true;
        const double d = __drcp_rn(piv);
        // pivot row entries for my two columns, pivot column entry for my row
        const int src_r = (p << 2) | (tx & 3);
        const double g0 = __shfl_sync(0xffffffffu, e0, src_r) * d, g1 = __shfl_sync(0xffffffffu, e1, src_r) * d;
        const int src_c = (r << 2) | (p >> 1);
        const double f0 = __shfl_sync(0xffffffffu, e0, src_c), f1 = __shfl_sync(0xffffffffu, e1, src_c);
        const double f = (p & 1) ? f1 : f0;
        if (r == p) {
          e0 = (c0 == p) ? d : g0;
          e1 = (c0 + 1 == p) ? d : g1;
        } else {
          e0 = (c0 == p) ? -f * d : e0 - f * g0;
          e1 = (c0 + 1 == p) ? -f * d : e1 - f * g1;
        }
      }
      pinv[r][c0] = e0; pinv[r][c0 + 1] = e1;
      if (bad) s_bad = 1;
    }
    __syncthreads();
    // (3) scaled row panel: rowq[k][j] = sum_m pinv[k][m] rowp[m][j]  (j outside P), pinv[k][j - pb] inside
    if (threadIdx.x < b * kB) {
      const int k = threadIdx.x / kB, j = threadIdx.x - k * kB;
      double s;
      if (j >= pb && j < pb + b) s = pinv[k][j - pb];
      else {
        s = 0.0;
#pragma unroll
        for (int m = 0; m < b; m++) s += pinv[k][m] * rowp[m][j];
      }
      rowq[k][j] = s;
    }
    __syncthreads();
    // (4) rank-8 update of the registers
    double s[3][3];
#pragma unroll
    for (int m = 0; m < 3; m++)
#pragma unroll
      for (int n = 0; n < 3; n++) s[m][n] = 0.0;
#pragma unroll
    for (int k = 0; k < b; k++) {
      double c[3], q[3];
#pragma unroll
      for (int m = 0; m < 3; m++) { c[m] = colp[(pb >> 3) & 1][ty + 32 * m][k]; q[m] = rowq[k][tx + 32 * m]; }
#pragma unroll
      for (int m = 0; m < 3; m++)
#pragma unroll
        for (int n = 0; n < 3; n++) s[m][n] += c[m] * q[n];
    }
    // i in P: the scaled pivot row;  j in P (i outside): -A[i, P] Pinv = -s;  otherwise A[i][j] - A[i, P] Pinv A[P, j]
#pragma unroll
    for (int m = 0; m < 3; m++) {
      const int i = ty + 32 * m;
      const bool irow = i >= pb && i < pb + b;
#pragma unroll
      for (int n = 0; n < 3; n++) {
        const int j = tx + 32 * n;
        const bool jcol = j >= pb && j < pb + b;
        v[m][n] = irow ? rowq[irow ? i - pb : 0][j] : (jcol ? -s[m][n] : v[m][n] - s[m][n]);
      }
    }
    // hazards: rowp is rewritten after barrier (3), rowq / pinv after the next barriers (1)-(2); colp is double-buffered
  }
#pragma unroll
  for (int m = 0; m < 3; m++)
#pragma unroll
    for (int n = 0; n < 3; n++) dst[(size_t)(ty + 32 * m) * kB + tx + 32 * n] = v[m][n];
  __syncthreads();
  if (threadIdx.x == 0 && s_bad) atomicAdd(&scalars[SC_BT_FAIL], 1.0);
}

// bt_gemm_mma_kernel: same task list as bt_gemm_kernel, on the fp64 tensor-core path (mma.sync m8n8k4, SASS DMMA).
// The two 96 x 96 operands of a term are staged in shared memory by TMA, one 768-byte bulk copy per matrix row into a
// row stride of 100 doubles: with that stride both the plain and the transposed fragment reads (8 x 4 / 4 x 8 doubles per
// warp) are bank-conflict free, so a transposed operand needs no transposing copy.  16 warps in a 4 x 4 grid, each owns a
// 24 x 24 output tile (3 x 3 fragments): 6 LDS + 9 DMMA per k-step of 4.
constexpr int kGemmLd = 100;
constexpr int kGemmSmem = 2 * kB * kGemmLd * 8;
__global__ void __launch_bounds__(512) bt_gemm_mma_kernel(const GemmTask* __restrict__ tasks) {
  extern __shared__ __align__(16) double gsm[];
  __shared__ uint64_t bar;
  double* As = gsm;
  double* Bs = gsm + kB * kGemmLd;
  const GemmTask T = tasks[blockIdx.x];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int m0 = 24 * (w >> 2), n0 = 24 * (w & 3);
  const int fr = lane >> 2, fk = lane & 3;
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  double2 acc[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) acc[i][j] = make_double2(0.0, 0.0);
  for (int term = 0; term < 2; term++) {
    const double* A = term ? T.A2 : T.A1;
    const double* Bm = term ? T.B2 : T.B1;
    if (!A) break;
    const int tA = term ? T.tA2 : T.tA1, tB = term ? T.tB2 : T.tB1;
    const double alpha = term ? T.alpha2 : T.alpha1;
    if (term) { __syncthreads(); fence_proxy_async_smem(); }   // everybody is done reading the previous operands
    if (threadIdx.x == 0) mbar_expect_tx(&bar, 2 * kBB * 8);
    __syncthreads();
    if (threadIdx.x < kB) tma_load_1d(As + threadIdx.x * kGemmLd, A + (size_t)threadIdx.x * kB, kB * 8, &bar);
    else if (threadIdx.x < 2 * kB) tma_load_1d(Bs + (threadIdx.x - kB) * kGemmLd, Bm + (size_t)(threadIdx.x - kB) * kB, kB * 8, &bar);
    mbar_wait(&bar, (uint32_t)term);
    // element strides in shared memory: op(A)[m][k] = As[m sAm + k sAk],  op(B)[k][n] = Bs[k sBk + n sBn]
    const int sAm = tA ? 1 : kGemmLd, sAk = tA ? kGemmLd : 1, sBk = tB ? 1 : kGemmLd, sBn = tB ? kGemmLd : 1;
    const double* ap = As + (m0 + fr) * sAm + fk * sAk;
    const double* bp = Bs + fk * sBk + (n0 + fr) * sBn;
    double2 part[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) part[i][j] = make_double2(0.0, 0.0);
#pragma unroll 4
    for (int k0 = 0; k0 < kB; k0 += 4) {
      double a[3], b[3];
#pragma unroll
      for (int i = 0; i < 3; i++) a[i] = ap[8 * i * sAm + k0 * sAk];
#pragma unroll
      for (int j = 0; j < 3; j++) b[j] = bp[k0 * sBk + 8 * j * sBn];
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) dmma_m8n8k4(part[i][j].x, part[i][j].y, a[i], b[j]);
    }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) { acc[i][j].x += alpha * part[i][j].x; acc[i][j].y += alpha * part[i][j].y; }
  }
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      double2* c = reinterpret_cast<double2*>(T.C + (size_t)(m0 + 8 * i + fr) * kB + n0 + 8 * j + 2 * fk);
      double2 o = acc[i][j];
      if (T.beta != 0.0) { const double2 old = *c; o.x += T.beta * old.x; o.y += T.beta * old.y; }
      *c = o;
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

// ---------------------------------------------------------------------------------------------------------
// pcg_bt_resident_kernel: the same PCG, organised around the B200's shared memory.  One CTA per super-block, all
// co-resident (cooperative launch, nsb <= #SMs).  CTA i keeps the three factor matrices of ITS super-block
// (Dinv_i, GaT_i, GcT_i: 3 x 96 x 96 doubles = 221 KB, row stride 97 -> conflict-free in both orientations) in shared
// memory for the whole solve: 148 SMs x 227 KB hold the entire 27 MB factorisation on chip, so applying the
// preconditioner touches only 96-double vectors in L2.  The cyclic-reduction sweeps run as a DATAFLOW instead of one grid
// barrier per level: an eliminated block pushes its two updates to its neighbours with red.add and bumps their arrival
// counters; a block starts as soon as its own counter reaches the expected value (forward) or its two neighbours have
// published z (backward).  Only the CG scalars (p.q, |r|^2, r.z) and the p exchange are grid-wide synchronisations.
constexpr int kLdR = 97;
constexpr int kResidentSmem = (3 * kB * kLdR + 5 * kB + 4 * kB) * 8;
struct ResidentSync {
  double* red_val;          // [4][4] rotating reduction slots
  unsigned int* red_cnt;    // [4]
  double* u;                // [nsb][96] forward contributions pushed by neighbours (zero between applies)
  unsigned int* cnt_w;      // [nsb] arrivals into u
  unsigned int* ready_z;    // [nsb] epoch of the published z block
};
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_inc(unsigned int* p) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ void st_release_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(kPcgThreads, 1) pcg_bt_resident_kernel(int nf, const uint32_t* __restrict__ sf_ptr,
                                                                          const uint32_t* __restrict__ sf_col,
                                                                          const double* __restrict__ Sf, const double* __restrict__ rhs,
                                                                          BtApply P, ResidentSync Y, double* __restrict__ y, double* r,
                                                                          double* p, int max_iter, double tol,
                                                                          double* __restrict__ scalars) {
  extern __shared__ __align__(16) double rsm[];
  double* sD = rsm;                       // Dinv_i
  double* sA = rsm + kB * kLdR;           // GaT_i
  double* sC = rsm + 2 * kB * kLdR;       // GcT_i
  double* part = rsm + 3 * kB * kLdR;     // [5][96] partial sums
  double* sw = part + 5 * kB;             // w_i
  double* sz = sw + kB;                   // z_i
  double* sv1 = sz + kB;                  // z_a  (or scratch)
  double* sv2 = sv1 + kB;                 // z_c
  __shared__ double red[kPcgThreads / 32];
  __shared__ double s_bcast;
  const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  const int I = blockIdx.x, nsb = P.nsb, nCTA = gridDim.x;
  const int n = 6 * nf;
  const int row0 = I * kB;
  const int my = row0 + tid;              // this thread's vector entry (tid < 96)
  const bool own = tid < kB && my < n;

  // ---- tree position of this block
  int lev = 0;                            // elimination level (root: nlev)
  if (I == 0) lev = P.nlev; else while (((I >> lev) & 1) == 0) lev++;
  const int sstep = 1 << lev;
  const int na = I == 0 ? -1 : I - sstep;
  int nc = -1;
  if (I != 0) { const int k = I >> lev, nl = (nsb + sstep - 1) >> lev; if (k + 1 < nl) nc = I + sstep; }
  int expected = 0;                       // contributions pushed into u_I before I is eliminated
  for (int l = 0; l < lev; l++) {
    const int s = 1 << l, k = I >> l, nl = (nsb + s - 1) >> l;
    if (k >= 1) expected++;
    if (k + 1 < nl) expected++;
  }

  // ---- factors -> shared memory (once per solve)
  for (int t = tid; t < kBB; t += kPcgThreads) {
    const int rr_ = t / kB, cc = t - rr_ * kB;
    sD[rr_ * kLdR + cc] = P.Dinv[(size_t)I * kBB + t];
    sA[rr_ * kLdR + cc] = P.GaT[(size_t)I * kBB + t];
    sC[rr_ * kLdR + cc] = P.GcT[(size_t)I * kBB + t];
  }

  unsigned int seq = 0;                   // grid-wide reduction / barrier sequence number (same in every CTA)
  // sum `v` (meaningful in thread 0 of each CTA after the block reduction) over the grid; everybody gets the total
  auto grid_sum = [&](double v) -> double {
    v = warp_sum(v);
    if (lane == 0) red[wib] = v;
    __syncthreads();
    const unsigned int slot = seq & 3u;
    if (tid == 0) {
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < kPcgThreads / 32; i++) s += red[i];
      if (I == 0) Y.red_val[4 * ((seq + 2u) & 3u)] = 0.0;   // nobody touches slot seq + 2 before CTA 0 arrives at seq + 1
      if (s != 0.0) atomicAdd(&Y.red_val[4 * slot], s);
      __threadfence();
      red_release_inc(&Y.red_cnt[slot]);
      const unsigned int target = (unsigned int)nCTA * (seq / 4u + 1u);
      while (ld_acquire_u32(&Y.red_cnt[slot]) < target) {}
      s_bcast = __ldcg(&Y.red_val[4 * slot]);
    }
    __syncthreads();
    seq++;
    return s_bcast;
  };

  unsigned int epoch = 0;
  // z_I = (T^-1 r)_I for every block, dataflow over the elimination tree.  Returns this CTA's share of r . z.
  auto apply = [&]() -> double {
    epoch++;
    // forward: wait for the neighbours' pushes, w_I = r_I - u_I
    if (tid == 0 && expected) { const unsigned int target = (unsigned int)expected * epoch; while (ld_acquire_u32(&Y.cnt_w[I]) < target) {} }
    __syncthreads();
    if (tid < kB) {
      const double rv = my < n ? r[my] : 0.0;
      double uv = 0.0;
      if (expected) { uv = __ldcg(&Y.u[(size_t)I * kB + tid]); Y.u[(size_t)I * kB + tid] = 0.0; }
      sw[tid] = rv - uv;
    }
    __syncthreads();
    if (I != 0) {
      // push GaT_I w_I to the left survivor, GcT_I w_I to the right one: thread (row, part) sums 20 columns
      if (tid < 5 * kB) {
        const int row = tid % kB, pt = tid / kB;
        const int c0 = pt * 20, c1 = min(c0 + 20, kB);
        double a = 0.0, c = 0.0;
        for (int cc = c0; cc < c1; cc++) { const double wv = sw[cc]; a += sA[row * kLdR + cc] * wv; c += sC[row * kLdR + cc] * wv; }
        // 5 partial sums per row go straight to the neighbours' accumulators
        if (na >= 0 && a != 0.0) atomicAdd(&Y.u[(size_t)na * kB + row], a);
        if (nc >= 0 && c != 0.0) atomicAdd(&Y.u[(size_t)nc * kB + row], c);
      }
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        if (na >= 0) red_release_inc(&Y.cnt_w[na]);
        if (nc >= 0) red_release_inc(&Y.cnt_w[nc]);
      }
      // backward: wait for z of the two survivors
      if (tid == 0) {
        if (na >= 0) while (ld_acquire_u32(&Y.ready_z[na]) < epoch) {}
        if (nc >= 0) while (ld_acquire_u32(&Y.ready_z[nc]) < epoch) {}
      }
      __syncthreads();
      if (tid < kB) {
        sv1[tid] = na >= 0 ? __ldcg(&P.z[(size_t)na * kB + tid]) : 0.0;
        sv2[tid] = nc >= 0 ? __ldcg(&P.z[(size_t)nc * kB + tid]) : 0.0;
      }
      __syncthreads();
    }
    // z_I[c] = sum_r Dinv[r][c] w[r] - GaT[r][c] za[r] - GcT[r][c] zc[r]   (thread (c, part) sums 20 rows)
    if (tid < 5 * kB) {
      const int col = tid % kB, pt = tid / kB;
      const int r0 = pt * 20, r1 = min(r0 + 20, kB);
      double a = 0.0;
      if (I != 0) {
        for (int rr_ = r0; rr_ < r1; rr_++) a += sD[rr_ * kLdR + col] * sw[rr_] - sA[rr_ * kLdR + col] * sv1[rr_] - sC[rr_ * kLdR + col] * sv2[rr_];
      } else {
        for (int rr_ = r0; rr_ < r1; rr_++) a += sD[rr_ * kLdR + col] * sw[rr_];
      }
      part[pt * kB + col] = a;
    }
    __syncthreads();
    double rz = 0.0;
    if (tid < kB) {
      const double zv = part[tid] + part[kB + tid] + part[2 * kB + tid] + part[3 * kB + tid] + part[4 * kB + tid];
      sz[tid] = zv;
      P.z[(size_t)row0 + tid] = zv;
      if (my < n) rz = r[my] * zv;
    }
    __syncthreads();
    if (tid == 0) { __threadfence(); st_release_u32(&Y.ready_z[I], epoch); }
    return rz;
  };

  // ---- initial residual
  double bbp = 0.0;
  if (tid < kB) {
    const double rv = my < n ? rhs[my] : 0.0;
    if (my < n) { y[my] = 0.0; r[my] = rv; }
    bbp = rv * rv;
  }
  __syncthreads();
  const double bb = grid_sum(bbp);
  int it = 0, brk = 0;
  double rr = bb;
  if (((volatile double*)scalars)[SC_BT_FAIL] != 0.0) brk = 2;  // factorisation failed: the host falls back to block-Jacobi
  if (bb > 0.0 && brk == 0) {
    double rho = grid_sum(apply());
    if (own) p[my] = sz[tid];
    (void)grid_sum(0.0);                  // p exchange
    while (it < max_iter) {
      // q_I = (S p)_I: one warp per pose row of this super-block
      double pq = 0.0, qv = 0.0;          // lanes 0..5 of warp w hold q for pose 16 I + w
      {
        const int i = I * kSbPoses + wib;
        if (i < nf) {
          const uint32_t p0 = sf_ptr[i], nb = sf_ptr[i + 1] - p0;
          const uint32_t len = nb * 6;
          const double* row = Sf + (size_t)p0 * 36;
          double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0;
          for (uint32_t e0 = lane; e0 < len; e0 += 128) {
            uint32_t ee[4]; double xv[4], sv[4][6];
#pragma unroll
            for (int u = 0; u < 4; u++) {
              ee[u] = e0 + 32 * u;
              const bool in = ee[u] < len;
              const uint32_t e = in ? ee[u] : 0u;
              const uint32_t k = e / 6, c = e - 6 * k;
              xv[u] = in ? __ldcg(&p[6 * sf_col[p0 + k] + c]) : 0.0;   // other CTAs' entries: L2, never a stale L1 line
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
            qv = lane == 0 ? a0 : lane == 1 ? a1 : lane == 2 ? a2 : lane == 3 ? a3 : lane == 4 ? a4 : a5;
            pq = qv * p[6 * i + lane];
          }
        }
      }
      // q travels through shared memory to the owner threads (tid < 96)
      if (lane < 6) sv1[6 * wib + lane] = qv;
      const double pqs = grid_sum(pq);    // (contains the block barriers that publish sv1)
      if (!(pqs > 0.0)) { brk = 1; break; }
      const double alpha = rho / pqs;
      double r2 = 0.0;
      if (own) {
        y[my] += alpha * p[my];
        const double rv = r[my] - alpha * sv1[tid];
        r[my] = rv;
        r2 = rv * rv;
      }
      __syncthreads();
      rr = grid_sum(r2);
      it++;
      if (rr <= tol * tol * bb) break;
      const double rho_new = grid_sum(apply());
      const double beta = rho_new / rho;
      rho = rho_new;
      if (own) p[my] = sz[tid] + beta * p[my];
      (void)grid_sum(0.0);                // p exchange
    }
  }
  if (I == 0 && tid == 0) {
    scalars[SC_PCG_IT] = (double)it;
    scalars[SC_PCG_RES] = bb > 0.0 ? sqrt(rr / bb) : 0.0;
    scalars[SC_PCG_BB] = bb;
    scalars[SC_PCG_BREAK] = (double)brk;
  }
}

// ------------------------------------------------------------------------------------------------------------------------
// The same resident PCG with FLAG-IN-DATA hand-offs.  Everything one CTA hands to another -- the forward contributions of
// the elimination tree, the published z blocks, the search direction p read by the neighbours' S p products, the partial
// sums of the grid-wide reductions -- travels as 16-byte words {lo32, tag, hi32, tag}: the consumer polls the DATA until
// both tags carry the expected value.  A hand-off is then one store and one (polled) load through L2; the version above
// pays atomics into an accumulator, a __threadfence, a release on a separate counter, an acquire spin on it and only then
// the load of the data -- about 3 us a hop against 1.2 us, on a critical path of 2 log2(nsb) hops per preconditioner
// application plus four reductions per iteration.  Tags increase monotonically (tag0 advances with every launch), so no
// buffer is cleared between uses; each 8-byte half validates itself, so the scheme does not depend on a 16-byte store
// being single-copy atomic.  The reductions add the per-CTA partials in a fixed order: the solve is run-to-run
// deterministic for a given S.  r, p and y live in the registers of the owning threads.
struct LLSync {
  uint4* red;               // [4][nsb] rotating reduction slots
  uint4* u;                 // [nsb][kLLSlots][96] forward contributions: slot 2 l + side (side 0: source on the left)
  uint4* z;                 // [nsb][96] published z blocks
  uint4* p;                 // [6 nf] search direction
  unsigned int tag0;        // first tag of this launch minus one
};
constexpr int kLLSlots = 16;  // 2 * (levels <= 8)
__device__ __forceinline__ void ll_store(uint4* a, double v, unsigned int tag) {
  const unsigned int lo = (unsigned int)__double2loint(v), hi = (unsigned int)__double2hiint(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(a), "r"(lo), "r"(tag), "r"(hi), "r"(tag) : "memory");
}
__device__ __forceinline__ uint4 ll_peek(const uint4* a) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(a) : "memory");
  return v;
}
__device__ __forceinline__ double ll_load(const uint4* a, unsigned int tag) {
  uint4 v;
  do { v = ll_peek(a); } while (v.y != tag || v.w != tag);
  return __hiloint2double((int)v.z, (int)v.x);
}

__global__ void __launch_bounds__(kPcgThreads, 1) pcg_bt_ll_kernel(int nf, const uint32_t* __restrict__ sf_ptr,
                                                                    const uint32_t* __restrict__ sf_col,
                                                                    const double* __restrict__ Sf, const double* __restrict__ rhs,
                                                                    BtApply P, LLSync Y, double* __restrict__ y, int max_iter,
                                                                    double tol, double* __restrict__ scalars) {
  extern __shared__ __align__(16) double rsm[];
  double* sD = rsm;                       // Dinv_i
  double* sA = rsm + kB * kLdR;           // GaT_i
  double* sC = rsm + 2 * kB * kLdR;       // GcT_i
  double* part = rsm + 3 * kB * kLdR;     // [5][96] partial sums
  double* sw = part + 5 * kB;             // w_i
  double* sz = sw + kB;                   // z_i
  double* sv1 = sz + kB;                  // z_a  (or q on its way to the owner threads)
  double* sv2 = sv1 + kB;                 // z_c
  __shared__ double red[kPcgThreads / 32], red2[kPcgThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  const int I = blockIdx.x, nsb = P.nsb, nCTA = gridDim.x;
  const int n = 6 * nf;
  const int row0 = I * kB;
  const int my = row0 + tid;              // this thread's vector entry (tid < 96)
  const bool own = tid < kB && my < n;

  // ---- tree position of this block
  int lev = 0;                            // elimination level (root: nlev)
  if (I == 0) lev = P.nlev; else while (((I >> lev) & 1) == 0) lev++;
  const int sstep = 1 << lev;
  const int na = I == 0 ? -1 : I - sstep;
  int nc = -1;
  if (I != 0) { const int k = I >> lev, nl = (nsb + sstep - 1) >> lev; if (k + 1 < nl) nc = I + sstep; }
  unsigned int umask = 0;                 // slots of u_I that receive a contribution before I is eliminated
  for (int l = 0; l < lev; l++) {
    const int s = 1 << l, k = I >> l, nl = (nsb + s - 1) >> l;
    if (k >= 1) umask |= 1u << (2 * l);
    if (k + 1 < nl) umask |= 1u << (2 * l + 1);
  }

  // ---- factors -> shared memory (once per solve)
  for (int t = tid; t < kBB; t += kPcgThreads) {
    const int rr_ = t / kB, cc = t - rr_ * kB;
    sD[rr_ * kLdR + cc] = P.Dinv[(size_t)I * kBB + t];
    sA[rr_ * kLdR + cc] = P.GaT[(size_t)I * kBB + t];
    sC[rr_ * kLdR + cc] = P.GcT[(size_t)I * kBB + t];
  }

  unsigned int seq = 0;                   // grid-wide reduction sequence number (same in every CTA)
  // sum `v` over the grid; everybody gets the total (partials added in CTA order)
  auto grid_sum = [&](double v) -> double {
    v = warp_sum(v);
    if (lane == 0) red[wib] = v;
    __syncthreads();
    const unsigned int tag = Y.tag0 + seq + 1u;
    uint4* slot = Y.red + (size_t)(seq & 3u) * nsb;
    if (tid == 0) {
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < kPcgThreads / 32; i++) s += red[i];
      ll_store(slot + I, s, tag);
    }
    double t = tid < nCTA ? ll_load(slot + tid, tag) : 0.0;
    t = warp_sum(t);
    if (lane == 0) red2[wib] = t;
    __syncthreads();
    double tot = 0.0;
    for (int i = 0; i < (nCTA + 31) / 32; i++) tot += red2[i];
    seq++;
    return tot;
  };

  unsigned int epoch = 0;
  // z_I = (T^-1 r)_I for every block, dataflow over the elimination tree; `rv` is the owner thread's residual entry.
  // Returns this thread's share of r . z.
  auto apply = [&](double rv) -> double {
    epoch++;
    const unsigned int tag = Y.tag0 + epoch;
    // forward: gather the neighbours' pushes, w_I = r_I - u_I
    if (umask) {
      if (tid < 5 * kB) {
        const int row = tid % kB, grp = tid / kB;
        const uint4* base = Y.u + ((size_t)I * kLLSlots) * kB + row;
        uint4 raw[4]; bool need[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int sl = grp + 5 * q;
          need[q] = sl < kLLSlots && ((umask >> sl) & 1u);
          if (need[q]) raw[q] = ll_peek(base + (size_t)sl * kB);
        }
        double a = 0.0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
          if (!need[q]) continue;
          while (raw[q].y != tag || raw[q].w != tag) raw[q] = ll_peek(base + (size_t)(grp + 5 * q) * kB);
          a += __hiloint2double((int)raw[q].z, (int)raw[q].x);
        }
        part[grp * kB + row] = a;
      }
      __syncthreads();
    }
    if (tid < kB) {
      double uv = 0.0;
      if (umask) uv = part[tid] + part[kB + tid] + part[2 * kB + tid] + part[3 * kB + tid] + part[4 * kB + tid];
      sw[tid] = (my < n ? rv : 0.0) - uv;
    }
    __syncthreads();
    if (I != 0) {
      // push GaT_I w_I to the left survivor, GcT_I w_I to the right one: thread (row, quarter) sums 24 columns, the four
      // quarters of a row sit in adjacent lanes
      if (tid < 4 * kB) {
        const int row = tid >> 2, pt = tid & 3;
        const int c0 = pt * 24;
        double a = 0.0, c = 0.0;
#pragma unroll 8
        for (int cc = c0; cc < c0 + 24; cc++) { const double wv = sw[cc]; a += sA[row * kLdR + cc] * wv; c += sC[row * kLdR + cc] * wv; }
        a += __shfl_xor_sync(0xffffffffu, a, 1); a += __shfl_xor_sync(0xffffffffu, a, 2);
        c += __shfl_xor_sync(0xffffffffu, c, 1); c += __shfl_xor_sync(0xffffffffu, c, 2);
        // I is the RIGHT neighbour of na (side 1) and the LEFT neighbour of nc (side 0)
        if (pt == 0 && na >= 0) ll_store(Y.u + ((size_t)na * kLLSlots + 2 * lev + 1) * kB + row, a, tag);
        if (pt == 1 && nc >= 0) ll_store(Y.u + ((size_t)nc * kLLSlots + 2 * lev) * kB + row, c, tag);
      }
      // backward: z of the two survivors
      if (tid < kB) sv1[tid] = na >= 0 ? ll_load(Y.z + (size_t)na * kB + tid, tag) : 0.0;
      else if (tid < 2 * kB) sv2[tid - kB] = nc >= 0 ? ll_load(Y.z + (size_t)nc * kB + (tid - kB), tag) : 0.0;
      __syncthreads();
    }
    // z_I[c] = sum_r Dinv[r][c] w[r] - GaT[r][c] za[r] - GcT[r][c] zc[r]   (thread (c, part) sums 20 rows)
    if (tid < 5 * kB) {
      const int col = tid % kB, pt = tid / kB;
      const int r0 = pt * 20, r1 = min(r0 + 20, kB);
      double a = 0.0;
      if (I != 0) {
        for (int rr_ = r0; rr_ < r1; rr_++) a += sD[rr_ * kLdR + col] * sw[rr_] - sA[rr_ * kLdR + col] * sv1[rr_] - sC[rr_ * kLdR + col] * sv2[rr_];
      } else {
        for (int rr_ = r0; rr_ < r1; rr_++) a += sD[rr_ * kLdR + col] * sw[rr_];
      }
      part[pt * kB + col] = a;
    }
    __syncthreads();
    double rz = 0.0;
    if (tid < kB) {
      const double zv = part[tid] + part[kB + tid] + part[2 * kB + tid] + part[3 * kB + tid] + part[4 * kB + tid];
      sz[tid] = zv;
      ll_store(Y.z + (size_t)row0 + tid, zv, tag);
      if (my < n) rz = rv * zv;
    }
    return rz;
  };

  // ---- initial residual
  double rown = 0.0, pown = 0.0, yown = 0.0;
  if (tid < kB && my < n) rown = rhs[my];
  __syncthreads();
  const double bb = grid_sum(rown * rown);
  int it = 0, brk = 0;
  double rr = bb;
  if (((volatile double*)scalars)[SC_BT_FAIL] != 0.0) brk = 2;  // factorisation failed: the host falls back to block-Jacobi
  if (bb > 0.0 && brk == 0) {
    double rho = grid_sum(apply(rown));
    if (own) { pown = sz[tid]; ll_store(Y.p + my, pown, Y.tag0 + 1u); }
    while (it < max_iter) {
      const unsigned int ptag = Y.tag0 + (unsigned int)it + 1u;
      // q_I = (S p)_I: one warp per pose row of this super-block
      double qv = 0.0;                    // lanes 0..5 of warp w hold q for pose 16 I + w
      {
        const int i = I * kSbPoses + wib;
        if (i < nf) {
          const uint32_t p0 = sf_ptr[i], nb = sf_ptr[i + 1] - p0;
          const uint32_t len = nb * 6;
          const double* row = Sf + (size_t)p0 * 36;
          double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0;
          for (uint32_t e0 = lane; e0 < len; e0 += 128) {
            uint32_t ee[4]; double xv[4], sv[4][6]; uint4 raw[4]; const uint4* xa[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
              ee[u] = e0 + 32 * u;
              const bool in = ee[u] < len;
              const uint32_t e = in ? ee[u] : 0u;
              const uint32_t k = e / 6, c = e - 6 * k;
              xa[u] = Y.p + (6 * (size_t)sf_col[p0 + k] + c);
              raw[u] = ll_peek(xa[u]);
#pragma unroll
              for (int a = 0; a < 6; a++) sv[u][a] = in ? row[a * len + e] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
              while (raw[u].y != ptag || raw[u].w != ptag) raw[u] = ll_peek(xa[u]);
              xv[u] = ee[u] < len ? __hiloint2double((int)raw[u].z, (int)raw[u].x) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
              a0 += sv[u][0] * xv[u]; a1 += sv[u][1] * xv[u]; a2 += sv[u][2] * xv[u];
              a3 += sv[u][3] * xv[u]; a4 += sv[u][4] * xv[u]; a5 += sv[u][5] * xv[u];
            }
          }
          a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3); a4 = warp_sum(a4); a5 = warp_sum(a5);
          if (lane < 6) qv = lane == 0 ? a0 : lane == 1 ? a1 : lane == 2 ? a2 : lane == 3 ? a3 : lane == 4 ? a4 : a5;
        }
      }
      // q travels through shared memory to the owner threads (tid < 96)
      if (lane < 6) sv1[6 * wib + lane] = qv;
      __syncthreads();
      const double qown = own ? sv1[tid] : 0.0;
      const double pqs = grid_sum(qown * pown);
      if (!(pqs > 0.0)) { brk = 1; break; }
      const double alpha = rho / pqs;
      if (own) { yown += alpha * pown; rown -= alpha * qown; }
      rr = grid_sum(own ? rown * rown : 0.0);
      it++;
      if (rr <= tol * tol * bb) break;
      const double rho_new = grid_sum(apply(rown));
      const double beta = rho_new / rho;
      rho = rho_new;
      if (own) { pown = sz[tid] + beta * pown; ll_store(Y.p + my, pown, Y.tag0 + (unsigned int)it + 1u); }
    }
  }
  if (own) y[my] = yown;
  if (I == 0 && tid == 0) {
    scalars[SC_PCG_IT] = (double)it;
    scalars[SC_PCG_RES] = bb > 0.0 ? sqrt(rr / bb) : 0.0;
    scalars[SC_PCG_BB] = bb;
    scalars[SC_PCG_BREAK] = (double)brk;
  }
}

}  // namespace obvi
