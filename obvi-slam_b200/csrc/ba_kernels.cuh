// CUDA kernels (sm_100a) of the bundle-adjustment hot path: residual + Jacobian evaluation, J^T J block
// accumulation, Schur elimination of points / objects, the reduced-camera-system PCG, back-substitution
// and the candidate-cost evaluation.  Everything is float64 (the reference is all-double).
//
// Data layout:
//   * reprojection observations are sorted pose-major; the Jacobian kernel writes one 128-byte chunk per
//     observation  [Jr 2x3 | Jl 2x3 | r 2 | pad]  (row-major, loss-corrected; the translation block of the pose
//     Jacobian is -Jl and is not stored), so a keyframe's Jacobian tile is one contiguous range;
//   * bbox observations are sorted object-major with 448-byte chunks [Jp 4x6 | Jo 4x7 | r 4];
//   * every e-block (point or object) has a CSR list of its observations (position, f index, merged pose slot)
//     and the precomputed index of the reduced-matrix block of every slot pair;
//   * the reduced camera system is accumulated into an upper block-CSR (6x6 blocks) in "pose-unscaled" form,
//     then `finish_kernel` applies Jacobi scaling + LM damping and mirrors it into a full symmetric BSR whose
//     scalar rows are contiguous for the PCG SpMV.
#pragma once

#include <cooperative_groups.h>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "factors.cuh"
#include "problem.hpp"

namespace obvi {
namespace cg = cooperative_groups;

// scalar slots in the device `scalars` array
enum : int {
  // 0-2 are produced by linearize (), 3-6 by take_step () / candidate_cost (): the two groups are reduced
  // across ranks separately in a sharded run
  SC_COST = 0, SC_FIXED = 1, SC_XNORM2 = 2, SC_CAND = 3, SC_CAND_FIXED = 4, SC_MODEL = 5, SC_STEP2 = 6,
  SC_GMAX = 7 /* u64 bits of a non-negative double */, SC_FAIL = 8, SC_PCG_IT = 9, SC_PCG_RES = 10, SC_PCG_BB = 11,
  SC_PCG_BREAK = 12, SC_BT_FAIL = 13, SC_COUNT = 16
};

struct LMParams { double radius, min_diag, max_diag; int compute_scale; double inv_radius; };

// ---- sharded runs: scalars that ride along with the collectives -------------------------------------------------
constexpr int kStage0Sums = 3;   // SC_COST, SC_FIXED, SC_XNORM2
constexpr int kPcgStatus = 5;    // SC_PCG_IT .. SC_BT_FAIL, appended to y for the broadcast of rank 0's solution
// tail of the all-reduced buffer: [cost, fixed, |x|^2 | (gradient max, failures) per rank].  `local` != 0: scalars 0-2 are this
// rank's partial sums (fresh linearisation); otherwise they already hold the global values and only rank 0 contributes them.
__global__ void pack_stage0_kernel(const double* __restrict__ scalars, double* __restrict__ tail, int rank, int world, int local) {
  const int i = threadIdx.x;
  if (i < kStage0Sums) tail[i] = (local || rank == 0) ? scalars[SC_COST + i] : 0.0;
  if (i < 2) tail[kStage0Sums + 2 * rank + i] = scalars[SC_GMAX + i];   // the other ranks' slots stay zero
}
__global__ void unpack_stage0_kernel(double* __restrict__ scalars, const double* __restrict__ tail, int world) {
  const int i = threadIdx.x;
  if (i < kStage0Sums) scalars[SC_COST + i] = tail[i];
  if (i == 0) { double g = 0.0, f = 0.0; for (int r = 0; r < world; r++) { g = fmax(g, tail[kStage0Sums + 2 * r]); f += tail[kStage0Sums + 2 * r + 1]; } scalars[SC_GMAX] = g; scalars[SC_FAIL] = f; }
}
__global__ void set_scalar_kernel(double* p, double v) { *p = v; }

// ------------------------------------------------------------------------------------------ reductions
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// sum over the block; result valid in every thread.  `red` needs >= 33 doubles of shared memory.
template <int T>
__device__ __forceinline__ double block_sum_all(double v, double* red) {
  v = warp_sum(v);
  if (T <= 32) return v;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  double s = 0;
#pragma unroll
  for (int i = 0; i < T / 32; i++) s += red[i];
  return s;
}
// Fire-and-forget fp64 reduction (SASS REDG).  In the e-block kernels nvcc turns a result-less atomicAdd into ATOMG (an atomic
// WITH a response: every one of them occupies a scoreboard slot until L2 answers), so the reduction is spelled out there.
__device__ __forceinline__ void red_add(double* p, double v) { asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v)); }
__device__ __forceinline__ void atomic_max_nonneg(double* slot, double v) {
  atomicMax(reinterpret_cast<unsigned long long*>(slot), (unsigned long long)__double_as_longlong(v));
}

// ------------------------------------------------------------------------------------------ pose / camera table
// One thread per (pose, camera): R_cw, t_cw and the rotation-derivative matrices shared by every
// observation of that pose (the reference recomputes them per observation on 9-wide Jets).
__global__ void pose_cam_kernel(const double* __restrict__ poses, int K, const Camera* __restrict__ cams, int C, int jac,
                                PoseCam* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K * C) return;
  const int k = i / C, c = i % C;
  double p[6];
#pragma unroll
  for (int a = 0; a < 6; a++) p[a] = poses[6 * k + a];
  PoseCam pc;
  make_pose_cam(p, cams[c].Rinv, cams[c].tinv, jac != 0, &pc);
  if (jac) {
    out[i] = pc;
  } else {
#pragma unroll
    for (int a = 0; a < 9; a++) out[i].Rcw[a] = pc.Rcw[a];
#pragma unroll
    for (int a = 0; a < 3; a++) out[i].tcw[a] = pc.tcw[a];
  }
}

// ------------------------------------------------------------------------------------------ reprojection: residual + Jacobian
// THE Jacobian-evaluation kernel.  One thread per observation.  Reads the 32-byte record (coalesced 2 x 16 B),
// the pose/camera entry (uniform across most of a warp: observations are pose-major), the point (gather through
// L2), writes the 128-byte chunk [Jr | Jl | r] with the Huber corrector applied, and reduces the cost.
constexpr int kJacThreads = 256;
// Reprojection chunk, 128 bytes per observation: [Jr 2x3 | Jl 2x3 | r 2 | pad 2], loss-corrected.  The translation block
// of the pose Jacobian is EXACTLY -Jl (d X_cam / d t = -R_cw = -d X_cam / d X, reproj_residual_jacobian), so only the
// rotation block Jr is stored and the full 2x6 pose Jacobian [-Jl | Jr] is rebuilt in registers by the consumers:
// 20 % less HBM traffic for the Jacobian kernel and for every kernel that re-reads the chunks, bit-identical values.
constexpr int kChunk = 16;     // doubles per reprojection observation
constexpr int kChunkJl = 6;    // offset of Jl
constexpr int kChunkR = 12;    // offset of r
__host__ __device__ __forceinline__ void decode_chunk(const double* ch, double* jp /*2x6*/, double* jl /*2x3*/, double* r /*2*/) {
#pragma unroll
  for (int a = 0; a < 6; a++) jl[a] = ch[kChunkJl + a];
#pragma unroll
  for (int k = 0; k < 2; k++)
#pragma unroll
    for (int a = 0; a < 3; a++) { jp[6 * k + a] = -jl[3 * k + a]; jp[6 * k + 3 + a] = ch[3 * k + a]; }
  r[0] = ch[kChunkR]; r[1] = ch[kChunkR + 1];
}
// 16-byte loads (device): the chunk is 16-byte aligned
__device__ __forceinline__ void load_chunk(const double* chunk, double* jp, double* jl, double& r0, double& r1) {
  const double2* c = reinterpret_cast<const double2*>(chunk);
  double raw[14];
#pragma unroll
  for (int a = 0; a < 7; a++) { const double2 v = c[a]; raw[2 * a] = v.x; raw[2 * a + 1] = v.y; }
  double rr[2];
  decode_chunk(raw, jp, jl, rr);
  r0 = rr[0]; r1 = rr[1];
}
__device__ __forceinline__ void store_chunk(double* chunk, const double* Jp, const double* Jl, const double* r, double sc, bool masked) {
  double2* out = reinterpret_cast<double2*>(chunk);
  if (masked) {   // removed in place (obvi_factor_remove without a structure rebuild): an all-zero block
#pragma unroll
    for (int a = 0; a < 8; a++) out[a] = make_double2(0.0, 0.0);
    return;
  }
  out[0] = make_double2(sc * Jp[3], sc * Jp[4]);
  out[1] = make_double2(sc * Jp[5], sc * Jp[9]);
  out[2] = make_double2(sc * Jp[10], sc * Jp[11]);
#pragma unroll
  for (int a = 0; a < 3; a++) out[3 + a] = make_double2(sc * Jl[2 * a], sc * Jl[2 * a + 1]);
  out[6] = make_double2(sc * r[0], sc * r[1]);
  out[7] = make_double2(0.0, 0.0);
}
constexpr uint32_t kObsMasked = 4u;  // flag bit 2 of ObsRec / BBoxRec: residual block removed in place (two-phase outlier exclusion)

// Pose-side sums of the reprojection blocks, kept apart from the reduced system (which is rebuilt for every trust-region
// radius while these only change with the linearisation point): per variable pose f the 6x6 block H_pp = sum Jp^T Jp
// (36, row-major), g_p = sum Jp^T r (6) and a spare 6 -> kPoseAcc doubles.  Filled by the Jacobian kernel, added into the
// reduced system by pose_acc_add_kernel.
constexpr int kPoseAcc = 48;

__global__ void __launch_bounds__(kJacThreads) reproj_jac_kernel(const ObsRec* __restrict__ obs, int64_t n,
                                                                  const PoseCam* __restrict__ pcam, int C,
                                                                  const CalibClass* __restrict__ cls,
                                                                  const double* __restrict__ points, int apply_loss,
                                                                  double* __restrict__ J, double* __restrict__ scalars) {
  __shared__ double red[33];
  const int64_t i = (int64_t)blockIdx.x * kJacThreads + threadIdx.x;
  double cost = 0.0, fixed = 0.0;
  if (i < n) {
    const double2 uv = reinterpret_cast<const double2*>(obs)[2 * i];
    const uint4 id = reinterpret_cast<const uint4*>(obs)[2 * i + 1];  // pose, point, chunk position, flags
    const CalibClass cc = cls[id.w >> 16];
    const PoseCam& pc = pcam[(size_t)id.x * C + cc.cam];
    const double X[3] = {points[3 * (size_t)id.y], points[3 * (size_t)id.y + 1], points[3 * (size_t)id.y + 2]};
    double r[2], Jp[12], Jl[6];
    reproj_residual_jacobian(pc, X, uv.x, uv.y, cc.mx, cc.my, r, Jp, Jl);
    const double s = r[0] * r[0] + r[1] * r[1];
    double sc = 1.0, c = 0.5 * s;
    if (apply_loss && cc.huber > 0.0) c = huber(cc.huber, s, &sc);
    const bool masked = (id.w & kObsMasked) != 0u;
    if (!masked) { if ((id.w & 3u) == 3u) fixed = c; else cost = c; }
    store_chunk(J + (size_t)id.z * kChunk, Jp, Jl, r, sc, masked);
  }
  cost = block_sum_all<kJacThreads>(cost, red);
  fixed = block_sum_all<kJacThreads>(fixed, red);
  if (threadIdx.x == 0) {
    if (cost != 0.0) atomicAdd(&scalars[SC_COST], cost);
    if (fixed != 0.0) atomicAdd(&scalars[SC_FIXED], fixed);
  }
}


// ------------------------------------------------------------------------------------------ TMA helpers (sm_90+ PTX)
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.b32 %0, 1, 0, p;\n"
      "}\n" : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
  return ok != 0u;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
// Warp-level wait: ONE lane spins on the barrier, the others join it at a warp barrier (which also carries the acquired view
// of the copied data to them): 31 fewer try_wait instructions per spin.
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity) {
  if ((threadIdx.x & 31) == 0) { while (!mbar_try_wait(bar, parity)) {} }
  __syncwarp();
}
// 1-D bulk copy global -> shared (TMA, completes on the mbarrier); size multiple of 16, 16-byte aligned
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
// 1-D bulk copy shared -> global (TMA store)
__device__ __forceinline__ void tma_store_1d(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_addr(src_smem)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// fp64 tensor-core product D (8x8) += A (8x4) B (4x8): lane holds A[lane/4][lane%4], B[lane%4][lane/4], D[lane/4][2 (lane%4) + {0,1}]
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// A row's chunk in the swizzled shared-memory tile image: 16-byte piece a of row r lives at piece a ^ (r & 7), so the 256
// threads of a CTA -- each writing the eight pieces of its own 128-byte row -- never collide on a bank.
__device__ __forceinline__ void store_chunk_swz(double* tile, int row, const double* Jp, const double* Jl, const double* r, double sc, bool masked) {
  double2* base = reinterpret_cast<double2*>(tile) + (size_t)row * 8;
  const int x = row & 7;
  double2 v[8];
  if (masked) {
#pragma unroll
    for (int a = 0; a < 8; a++) v[a] = make_double2(0.0, 0.0);
  } else {
    v[0] = make_double2(sc * Jp[3], sc * Jp[4]);
    v[1] = make_double2(sc * Jp[5], sc * Jp[9]);
    v[2] = make_double2(sc * Jp[10], sc * Jp[11]);
#pragma unroll
    for (int a = 0; a < 3; a++) v[3 + a] = make_double2(sc * Jl[2 * a], sc * Jl[2 * a + 1]);
    v[6] = make_double2(sc * r[0], sc * r[1]);
    v[7] = make_double2(0.0, 0.0);
  }
#pragma unroll
  for (int a = 0; a < 8; a++) base[a ^ x] = v[a];
}

// ------------------------------------------------------------------------------------------ reprojection Jacobians, fused
// THE Jacobian-evaluation kernel.  One thread per observation, observations in POSE-major order, chunks stored POINT-major.
//   * the pose/camera entries a 256-observation tile needs are a short contiguous range of the table: one 1-D TMA bulk copy
//     stages them in shared memory;
//   * every thread writes its 128-byte chunk into a swizzled shared-memory image of the tile (16-byte piece a of row r at
//     a ^ (r & 7): conflict free), then each WARP scatters its own 32 rows to their point-major positions -- eight lanes per
//     row, i.e. every store instruction writes four complete 128-byte lines;
//   * the pose-side normal-equation sums H_pp = sum Jp^T Jp, g_p = sum Jp^T r of the rows (which no later kernel could form
//     without gathering, now that the chunks are point-major) are taken from the same image with the fp64 tensor cores: per
//     pair of rows one m8n8k4 product [Jp | r]^T [Jp | r], A and B fragments being the same register.  Warps that span a
//     keyframe boundary run one masked pass per keyframe; every (warp, keyframe) product leaves as 42 fire-and-forget
//     reductions (combining the warps of a tile in shared memory first cost more instructions than it saved).
// Only the cost sum needs a block barrier: scatter and products read what the same warp wrote.
constexpr int kJacMaxPc = 6;                       // staged pose/camera entries per tile
constexpr int kJacMaxCls = 16;                     // calibration classes staged in shared memory
constexpr int kJacWarps = kJacThreads / 32;
constexpr int kJacTileBytes = kJacThreads * kChunk * 8;
constexpr int kJacSmemBytes = kJacTileBytes + kJacMaxPc * (int)sizeof(PoseCam) + 16 + 128;

#ifndef OBVI_JAC_MINB
#define OBVI_JAC_MINB 4
#endif
__global__ void __launch_bounds__(kJacThreads, OBVI_JAC_MINB) reproj_jac_fused_kernel(const ObsRec* __restrict__ obs, int64_t n,
                                                                           const PoseCam* __restrict__ pcam, int C,
                                                                           const CalibClass* __restrict__ cls, int ncls,
                                                                           const double* __restrict__ points, int apply_loss,
                                                                           const uint4* __restrict__ tile_pc,
                                                                           const int32_t* __restrict__ f_of_pose,
                                                                           double* __restrict__ J, double* __restrict__ pose_acc,
                                                                           double* __restrict__ scalars) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ CalibClass cls_s[kJacMaxCls];
  __shared__ double wcost[kJacWarps], wfixed[kJacWarps];
  unsigned char* smem = smem_raw + ((128u - (smem_addr(smem_raw) & 127u)) & 127u);
  double* tile = reinterpret_cast<double*>(smem);
  PoseCam* pc_s = reinterpret_cast<PoseCam*>(smem + kJacTileBytes);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kJacTileBytes + kJacMaxPc * sizeof(PoseCam));
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t i0 = (int64_t)blockIdx.x * kJacThreads;
  const int nt = (int)min((int64_t)kJacThreads, n - i0);
  const uint4 tp = tile_pc[blockIdx.x];  // first pose/camera entry of the tile, number of entries staged, first keyframe
  const int64_t i = i0 + threadIdx.x;
  const bool active = (int)threadIdx.x < nt;
  double2 uv = make_double2(0.0, 0.0);
  uint4 id = make_uint4(0, 0, 0, 3u);
  if (active) {  // issue the record loads first: they head the longest dependency chain (record -> point)
    uv = reinterpret_cast<const double2*>(obs)[2 * i];
    id = reinterpret_cast<const uint4*>(obs)[2 * i + 1];
  }
  if (threadIdx.x == 0) mbar_init(bar, 1);
  if ((int)threadIdx.x < ncls * 4) reinterpret_cast<double*>(cls_s)[threadIdx.x] = reinterpret_cast<const double*>(cls)[threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, tp.y * (uint32_t)sizeof(PoseCam));
    tma_load_1d(pc_s, pcam + tp.x, tp.y * (uint32_t)sizeof(PoseCam), bar);
  }
  double cost = 0.0, fixed = 0.0;
  double X[3] = {0.0, 0.0, 0.0};
  if (active) {
    X[0] = points[3 * (size_t)id.y]; X[1] = points[3 * (size_t)id.y + 1]; X[2] = points[3 * (size_t)id.y + 2];
  }
  const CalibClass cc = cls_s[(id.w >> 16) & (kJacMaxCls - 1)];
  const uint32_t pose0 = tp.z;
  // accumulation slot of this row: keyframe index inside the tile, -1: no pose-side sums (inactive row, constant pose, masked)
  int slot = -1;
  mbar_wait(bar, 0);
  {
    double r[2], Jp[12], Jl[6];
    double sc = 1.0;
    bool masked = true;
    if (active) {
      const uint32_t pci = id.x * (uint32_t)C + ((id.w >> 8) & 0xffu);
      const uint32_t rel = pci - tp.x;
      // two explicit paths so that the staged entries are read with shared-memory loads (a `cond ? smem : global`
      // reference would compile to generic loads, tracked by the long scoreboard)
      if (rel < tp.y) reproj_residual_jacobian(pc_s[rel], X, uv.x, uv.y, cc.mx, cc.my, r, Jp, Jl);
      else reproj_residual_jacobian(pcam[pci], X, uv.x, uv.y, cc.mx, cc.my, r, Jp, Jl);
      const double s = r[0] * r[0] + r[1] * r[1];
      double c = 0.5 * s;
      if (apply_loss && cc.huber > 0.0) c = huber(cc.huber, s, &sc);
      masked = (id.w & kObsMasked) != 0u;
      if (!masked) { if ((id.w & 3u) == 3u) fixed = c; else cost = c; }
      if (!(id.w & 1u) && !masked) slot = (int)(id.x - pose0);
    }
    if (active) store_chunk_swz(tile, (int)threadIdx.x, Jp, Jl, r, sc, masked);
    else {
      double2* base = reinterpret_cast<double2*>(tile) + (size_t)threadIdx.x * 8;
#pragma unroll
      for (int a = 0; a < 8; a++) base[a] = make_double2(0.0, 0.0);
    }
  }
  __syncwarp();
  // ---- scatter the warp's 32 rows: eight lanes per row, four complete lines per store instruction
  {
    const int sub = lane >> 3, piece = lane & 7;
#pragma unroll
    for (int it = 0; it < 8; it++) {
      const int rl = 4 * it + sub;                                   // row inside the warp
      const uint32_t d = __shfl_sync(0xffffffffu, id.z, rl);
      const int row = 32 * w + rl;
      const double2 v = reinterpret_cast<const double2*>(tile)[(size_t)row * 8 + (piece ^ (row & 7))];
      if (row < nt) reinterpret_cast<double2*>(J)[(size_t)d * 8 + piece] = v;
    }
  }
  // ---- pose-side sums of the warp's rows on the fp64 tensor cores; one set of reductions per (warp, keyframe)
  {
    const int a = lane >> 2, k = lane & 3;                           // fragment element [a][k]: parameter column a, residual row k
    const int rr = k & 1;                                            // residual row inside the observation
    // chunk [Jr 2x3 | Jl 2x3 | r 2 | pad]: columns 0-2 of Jp are -Jl, 3-5 are Jr, column 6 carries r (-> g_p), column 7 is zero
    const int idx = a < 3 ? 6 + 3 * rr + a : (a < 6 ? 3 * rr + (a - 3) : 12 + rr);
    const double sgn = a < 3 ? -1.0 : (a < 7 ? 1.0 : 0.0);
    const int pc = idx >> 1, wi = idx & 1;
    const int smin = __reduce_min_sync(0xffffffffu, slot < 0 ? 0x7fffffff : slot);
    const int smax = __reduce_max_sync(0xffffffffu, slot);
    for (int s = smin; s <= smax; s++) {
      const uint32_t m = __ballot_sync(0xffffffffu, slot == s);
      if (m == 0u) continue;
      double d0 = 0.0, d1 = 0.0;
#pragma unroll
      for (int j = 0; j < 16; j++) {
        const int rl = 2 * j + (k >> 1);
        const int row = 32 * w + rl;
        double v = tile[(size_t)row * kChunk + 2 * (pc ^ (row & 7)) + wi];
        v = ((m >> rl) & 1u) ? sgn * v : 0.0;
        dmma_m8n8k4(d0, d1, v, v);
      }
      const int f = f_of_pose[pose0 + (uint32_t)s];
      double* acc = pose_acc + (size_t)f * kPoseAcc;
      const int c0 = 2 * k;
      if (a < 6) {
        if (c0 < 6) { red_add(&acc[6 * a + c0], d0); red_add(&acc[6 * a + c0 + 1], d1); }
        else red_add(&acc[36 + a], d0);                              // column 6: g_p
      }
    }
  }
  cost = warp_sum(cost); fixed = warp_sum(fixed);
  if (lane == 0) { wcost[w] = cost; wfixed[w] = fixed; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double cs = 0.0, fs = 0.0;
#pragma unroll
    for (int ww = 0; ww < kJacWarps; ww++) { cs += wcost[ww]; fs += wfixed[ww]; }
    if (cs != 0.0) atomicAdd(&scalars[SC_COST], cs);
    if (fs != 0.0) atomicAdd(&scalars[SC_FIXED], fs);
  }
}

// Residual-only evaluation at the candidate point (cost only; nothing is stored).  kCostPer consecutive tiles of 256 records
// per CTA: the record loads of all of them are issued before the first gather, and the two block reductions + atomics are
// paid once per kCostPer * 256 observations.
constexpr int kCostPer = 4;
__global__ void __launch_bounds__(kJacThreads) reproj_cost_kernel(const ObsRec* __restrict__ obs, int64_t n,
                                                                   const PoseCam* __restrict__ pcam, int C,
                                                                   const CalibClass* __restrict__ cls,
                                                                   const double* __restrict__ points,
                                                                   double* __restrict__ scalars) {
  __shared__ double red[33];
  const int64_t i0 = (int64_t)blockIdx.x * (kJacThreads * kCostPer) + threadIdx.x;
  double cost = 0.0, fixed = 0.0;
  double2 uv[kCostPer]; uint4 id[kCostPer];
#pragma unroll
  for (int k = 0; k < kCostPer; k++) {
    const int64_t i = i0 + (int64_t)k * kJacThreads;
    if (i < n) { uv[k] = reinterpret_cast<const double2*>(obs)[2 * i]; id[k] = reinterpret_cast<const uint4*>(obs)[2 * i + 1]; }
  }
#pragma unroll
  for (int k = 0; k < kCostPer; k++) {
    const int64_t i = i0 + (int64_t)k * kJacThreads;
    if (i >= n) continue;
    const CalibClass cc = cls[id[k].w >> 16];
    const PoseCam& pc = pcam[(size_t)id[k].x * C + cc.cam];
    const double X[3] = {points[3 * (size_t)id[k].y], points[3 * (size_t)id[k].y + 1], points[3 * (size_t)id[k].y + 2]};
    double r[2];
    reproj_residual(pc, X, uv[k].x, uv[k].y, cc.mx, cc.my, r);
    const double s = r[0] * r[0] + r[1] * r[1];
    double sc, c = 0.5 * s;
    if (cc.huber > 0.0) c = huber(cc.huber, s, &sc);
    if (id[k].w & kObsMasked) c = 0.0;
    if ((id[k].w & 3u) == 3u) fixed += c; else cost += c;
  }
  cost = block_sum_all<kJacThreads>(cost, red);
  fixed = block_sum_all<kJacThreads>(fixed, red);
  if (threadIdx.x == 0) {
    if (cost != 0.0) atomicAdd(&scalars[SC_CAND], cost);
    if (fixed != 0.0) atomicAdd(&scalars[SC_CAND_FIXED], fixed);
  }
}

// Residual export (Problem::Evaluate): the two residuals of every chunk, packed -- 16 instead of 128 bytes per observation
// cross PCIe.
__global__ void extract_residuals_kernel(const double* __restrict__ J, int64_t n, double2* __restrict__ out) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) out[q] = *reinterpret_cast<const double2*>(J + q * kChunk + kChunkR);
}

// ------------------------------------------------------------------------------------------ pose-side J^T J accumulation
// Fallback beside reproj_jac_kernel (the kernel without TMA staging / fused sums): one CTA per keyframe walks the keyframe's
// pose-major records, gathers each chunk from its point-major position and accumulates H_pp, g_p with warp-shuffle
// reductions; one set of reductions per CTA.
constexpr int kPoseAccThreads = 256;
__global__ void __launch_bounds__(kPoseAccThreads) pose_accum_gather_kernel(const ObsRec* __restrict__ obs, const double* __restrict__ J,
                                                                             const uint32_t* __restrict__ pose_ptr,
                                                                             const int32_t* __restrict__ f_of_pose,
                                                                             double* __restrict__ pose_acc) {
  const int k = blockIdx.x;
  const int f = f_of_pose[k];
  if (f < 0) return;
  const uint32_t b0 = pose_ptr[k], b1 = pose_ptr[k + 1];
  if (b0 == b1) return;
  double acc[27];
#pragma unroll
  for (int a = 0; a < 27; a++) acc[a] = 0.0;
  for (uint32_t i = b0 + threadIdx.x; i < b1; i += kPoseAccThreads) {
    double jp[12], jl_[6];
    double2 r;
    load_chunk(J + (size_t)obs[i].dst * kChunk, jp, jl_, r.x, r.y);
    int t = 0;
#pragma unroll
    for (int a = 0; a < 6; a++) {
#pragma unroll
      for (int b = a; b < 6; b++) acc[t++] += jp[a] * jp[b] + jp[6 + a] * jp[6 + b];
    }
#pragma unroll
    for (int a = 0; a < 6; a++) acc[21 + a] += jp[a] * r.x + jp[6 + a] * r.y;
  }
  __shared__ double red[27][kPoseAccThreads / 32];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
  for (int a = 0; a < 27; a++) {
    const double v = warp_sum(acc[a]);
    if (l == 0) red[a][w] = v;
  }
  __syncthreads();
  if (threadIdx.x < 27) {
    double s = 0;
#pragma unroll
    for (int i = 0; i < kPoseAccThreads / 32; i++) s += red[threadIdx.x][i];
    double* A = pose_acc + (size_t)f * kPoseAcc;
    if (threadIdx.x < 21) {
      int a = 0, t = threadIdx.x;
      while (t >= 6 - a) { t -= 6 - a; a++; }
      const int b = a + t;
      red_add(&A[a * 6 + b], s);
      if (a != b) red_add(&A[b * 6 + a], s);
    } else {
      red_add(&A[36 + (threadIdx.x - 21)], s);
    }
  }
}
// reduced system += pose-side sums of the reprojection blocks (one thread per entry)
__global__ void pose_acc_add_kernel(int nf, const double* __restrict__ pose_acc, const uint32_t* __restrict__ su_ptr,
                                    double* __restrict__ S_upper, double* __restrict__ gp, double* __restrict__ hpp_diag) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nf * 42) return;
  const int f = t / 42, e = t - 42 * f;
  const double v = pose_acc[(size_t)f * kPoseAcc + e];
  if (v == 0.0) return;
  // reductions, not plain stores: the side-stream kernels (objects, priors, rel-pose) add into the same entries concurrently
  if (e < 36) {
    red_add(&S_upper[(size_t)su_ptr[f] * 36 + e], v);   // the diagonal block of row f is the first block of its upper row
    if (e % 7 == 0) red_add(&hpp_diag[6 * f + e / 7], v);
  } else {
    red_add(&gp[6 * f + (e - 36)], v);
  }
}

// The generic e-block kernels below index a chunk as [Jp KRx6 | Je KRxNE | r KR].  Object chunks (bounding boxes) are stored that
// way; point chunks are stored compact (see kChunk) and expanded into a register image of the same layout.
template <int NE, int KR>
__device__ __forceinline__ const double* chunk_full(const double* J, size_t pos, double* tmp /* KR (6 + NE + 1) */) {
  if constexpr (NE == 3 && KR == 2) {
    double rr[2];
    decode_chunk(J + pos * kChunk, tmp, tmp + 12, rr);
    tmp[18] = rr[0]; tmp[19] = rr[1];
    return tmp;
  } else {
    return J + pos * (size_t)(KR * (6 + NE + 1));
  }
}

// ------------------------------------------------------------------------------------------ e-block elimination
struct EArgs {
  const uint32_t* ptr; const uint32_t* pos; const int32_t* f; const uint16_t* slot;
  const uint32_t* pair_ptr; const uint16_t* nslots; const uint32_t* pair_blk;
  const uint8_t* cst;
  const double* J;          // chunks [Jp KRx6 | Je KRxNE | r KR]
  double* escale;           // NE per e-block (Jacobi scaling, written when compute_scale)
  double* einv;             // NE*NE per e-block: S_e (S_e H S_e + D^2)^-1 S_e
  double* eg;               // NE per e-block: E^T r
  const double* prior_H;    // NE*NE per e-block or null (unary factors on the e-block)
  const double* prior_g;    // NE per e-block or null
  double* overflow;         // staging for e-blocks with more than MAXS slots
  const uint32_t* overflow_off;  // per e-block offset (in doubles) into overflow, valid when nslots > MAXS
  const uint32_t* elist;         // optional: process e-block elist[blockIdx.x] instead of blockIdx.x
  int ne;
  int batch_begin = 0;      // streaming point kernels: first batch (4 points) that holds a regular point of this rank
};

// One CTA (T threads) per e-block.  Phase A: H_ee, g_e (shuffle reductions), damping, inverse.
// Phase B: per merged pose slot W = sum Jp^T Je and Z = W Hinv, staged in shared memory.
// Phase C: S_ab -= Z_a W_b^T for every slot pair, spread over the threads by block row.
// PREP_ONLY: stop after phase A (einv / eg / escale written; an all-zero einv on a failed inverse) -- the first half of the
// split object path, whose phases B and C run in obj_schur_kernel.
template <int NE, int KR, int T, int MAXS, bool POSE_SIDE, bool PREP_ONLY = false>
__global__ void __launch_bounds__(T) schur_eblock_kernel(EArgs A, LMParams lm, const uint32_t* __restrict__ su_ptr,
                                                          double* __restrict__ S_upper, double* __restrict__ gp,
                                                          double* __restrict__ hpp_diag, double* __restrict__ b_schur,
                                                          double* __restrict__ scalars) {
  constexpr int CH = KR * (6 + NE + 1);
  constexpr int NH = NE * (NE + 1) / 2;
  constexpr int SL = 12 * NE;  // doubles per slot in the staging buffer: Z (6xNE) then W (6xNE)
  __shared__ double red[33];
  __shared__ double stage_s[PREP_ONLY ? 1 : MAXS * SL];
  const int e = A.elist ? (int)A.elist[blockIdx.x] : (int)blockIdx.x;
  if (A.cst[e]) return;
  const uint32_t b0 = A.ptr[e], b1 = A.ptr[e + 1];
  if (b0 == b1 && A.prior_H == nullptr) return;
  // ---- phase A
  double H[NH], g[NE];
#pragma unroll
  for (int a = 0; a < NH; a++) H[a] = 0.0;
#pragma unroll
  for (int a = 0; a < NE; a++) g[a] = 0.0;
  for (uint32_t q = b0 + threadIdx.x; q < b1; q += T) {
    double chunk_tmp_1[CH];
    const double* ch = chunk_full<NE, KR>(A.J, (size_t)A.pos[q], chunk_tmp_1);
    const double* Je = ch + KR * 6;
    const double* r = ch + KR * (6 + NE);
#pragma unroll
    for (int k = 0; k < KR; k++) {
      double je[NE];
#pragma unroll
      for (int a = 0; a < NE; a++) je[a] = Je[k * NE + a];
      const double rk = r[k];
      int t = 0;
#pragma unroll
      for (int a = 0; a < NE; a++) {
        g[a] += je[a] * rk;
#pragma unroll
        for (int b = a; b < NE; b++) H[t++] += je[a] * je[b];
      }
    }
  }
#pragma unroll
  for (int a = 0; a < NH; a++) H[a] = block_sum_all<T>(H[a], red);
#pragma unroll
  for (int a = 0; a < NE; a++) g[a] = block_sum_all<T>(g[a], red);
  if (A.prior_H) {
    const double* ph = A.prior_H + (size_t)e * NE * NE;
    int t = 0;
#pragma unroll
    for (int a = 0; a < NE; a++) {
      g[a] += A.prior_g[(size_t)e * NE + a];
#pragma unroll
      for (int b = a; b < NE; b++) H[t++] += ph[a * NE + b];
    }
  }
  double s[NE], Hs[NE * NE], hinv[NE * NE];
  {
    int t = 0;
    double gm = 0.0;
#pragma unroll
    for (int a = 0; a < NE; a++) {
      const double haa = H[t];
      t += NE - a;
      s[a] = lm.compute_scale ? 1.0 / (1.0 + sqrt(haa)) : A.escale[(size_t)e * NE + a];
      gm = fmax(gm, fabs(g[a]));
    }
    if (threadIdx.x == 0) {
      if (lm.compute_scale) {
#pragma unroll
        for (int a = 0; a < NE; a++) A.escale[(size_t)e * NE + a] = s[a];
      }
      atomic_max_nonneg(&scalars[SC_GMAX], gm);
    }
    t = 0;
#pragma unroll
    for (int a = 0; a < NE; a++) {
#pragma unroll
      for (int b = a; b < NE; b++) {
        const double v = s[a] * H[t++] * s[b];
        Hs[a * NE + b] = v;
        Hs[b * NE + a] = v;
      }
    }
#pragma unroll
    for (int a = 0; a < NE; a++) Hs[a * NE + a] += fmin(fmax(Hs[a * NE + a], lm.min_diag), lm.max_diag) / lm.radius;
  }
  if (!spd_inverse<NE>(Hs, hinv)) {
    if (threadIdx.x == 0) {
      atomicAdd(&scalars[SC_FAIL], 1.0);
      if (PREP_ONLY) {
#pragma unroll
        for (int a = 0; a < NE * NE; a++) A.einv[(size_t)e * NE * NE + a] = 0.0;
#pragma unroll
        for (int a = 0; a < NE; a++) A.eg[(size_t)e * NE + a] = g[a];
      }
    }
    return;
  }
#pragma unroll
  for (int a = 0; a < NE; a++) {
#pragma unroll
    for (int b = 0; b < NE; b++) hinv[a * NE + b] *= s[a] * s[b];
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int a = 0; a < NE * NE; a++) A.einv[(size_t)e * NE * NE + a] = hinv[a];
#pragma unroll
    for (int a = 0; a < NE; a++) A.eg[(size_t)e * NE + a] = g[a];
  }
  if (PREP_ONLY) return;
  // ---- phase B
  const int ns = A.nslots[e];
  if (ns == 0) return;
  double* stage = (ns <= MAXS) ? stage_s : (A.overflow + A.overflow_off[e]);
  for (uint32_t q = b0 + threadIdx.x; q < b1; q += T) {
    const uint16_t sl = A.slot[q];
    if (sl == 0xFFFF) continue;
    if (q > b0 && A.slot[q - 1] == sl) continue;  // not the head of its run
    double Wm[6 * NE];
#pragma unroll
    for (int a = 0; a < 6 * NE; a++) Wm[a] = 0.0;
    double hp[21], gq[6];
    if (POSE_SIDE) {
#pragma unroll
      for (int a = 0; a < 21; a++) hp[a] = 0.0;
#pragma unroll
      for (int a = 0; a < 6; a++) gq[a] = 0.0;
    }
    for (uint32_t q2 = q; q2 < b1 && A.slot[q2] == sl; q2++) {
      double chunk_tmp_2[CH];
      const double* ch = chunk_full<NE, KR>(A.J, (size_t)A.pos[q2], chunk_tmp_2);
#pragma unroll
      for (int k = 0; k < KR; k++) {
        double jp[6], je[NE];
#pragma unroll
        for (int a = 0; a < 6; a++) jp[a] = ch[k * 6 + a];
#pragma unroll
        for (int a = 0; a < NE; a++) je[a] = ch[KR * 6 + k * NE + a];
#pragma unroll
        for (int a = 0; a < 6; a++) {
#pragma unroll
          for (int c = 0; c < NE; c++) Wm[a * NE + c] += jp[a] * je[c];
        }
        if (POSE_SIDE) {
          const double rk = ch[KR * (6 + NE) + k];
          int t = 0;
#pragma unroll
          for (int a = 0; a < 6; a++) {
            gq[a] += jp[a] * rk;
#pragma unroll
            for (int b = a; b < 6; b++) hp[t++] += jp[a] * jp[b];
          }
        }
      }
    }
    const int fi = A.f[q];
    double* st = stage + (size_t)sl * SL;
#pragma unroll
    for (int a = 0; a < 6; a++) {
      double zg = 0.0;
#pragma unroll
      for (int c = 0; c < NE; c++) {
        double z = 0.0;
#pragma unroll
        for (int d = 0; d < NE; d++) z += Wm[a * NE + d] * hinv[d * NE + c];
        st[a * NE + c] = z;
        st[6 * NE + a * NE + c] = Wm[a * NE + c];
        zg += z * g[c];
      }
      red_add(&b_schur[6 * fi + a], -zg);
    }
    if (POSE_SIDE) {
      double* Sd = S_upper + (size_t)su_ptr[fi] * 36;
      int t = 0;
#pragma unroll
      for (int a = 0; a < 6; a++) {
        red_add(&gp[6 * fi + a], gq[a]);
#pragma unroll
        for (int b = a; b < 6; b++) {
          const double v = hp[t++];
          red_add(&Sd[a * 6 + b], v);
          if (a != b) red_add(&Sd[b * 6 + a], v);
          else red_add(&hpp_diag[6 * fi + a], v);
        }
      }
    }
  }
  if (ns > MAXS) __threadfence_block();
  __syncthreads();
  // ---- phase C
  const int items = ns * (ns + 1) / 2 * 6;
  const uint32_t* pb = A.pair_blk + A.pair_ptr[e];
  for (int it = threadIdx.x; it < items; it += T) {
    const int pr = it / 6, row = it - 6 * pr;
    int a = 0, t = pr;
    while (t >= ns - a) { t -= ns - a; a++; }
    const int b = a + t;
    const double* Za = stage + (size_t)a * SL + row * NE;
    const double* Wb = stage + (size_t)b * SL + 6 * NE;
    double z[NE];
#pragma unroll
    for (int d = 0; d < NE; d++) z[d] = Za[d];
    double* Sb = S_upper + (size_t)pb[pr] * 36 + row * 6;
#pragma unroll
    for (int c = 0; c < 6; c++) {
      double v = 0.0;
#pragma unroll
      for (int d = 0; d < NE; d++) v += z[d] * Wb[c * NE + d];
      red_add(&Sb[c], -v);
    }
  }
}

// Second half of the split object path (objects: NE = 7, bbox chunks [Jp 4x6 | Je 4x7 | r 4]); H_e^-1 and g_e come from
// schur_eblock_kernel<7, 4, 32, ., true, PREP_ONLY>.  Phase B is spread over (pose slot, row a of the 6x7 block W) instead of
// one thread per slot, so a thread keeps 7 + 6 accumulators instead of 42 + 21 + the 49-entry inverse (which lives in shared
// memory here): 64 registers instead of 254, i.e. the CTAs of this latency-bound kernel no longer take the whole register
// file away from the pose-accumulation kernel they run beside.  Sums are formed in the same order as in the one-kernel
// version.  Phase C (S_ab -= Z_a W_b^T over slot pairs): the pair index is inverted with a square root instead of a loop.
constexpr int kObjMaxSlots = 56;   // pose slots staged on chip: 56 x 84 doubles + the 1596-entry pair table = 43 KB of static shared memory
constexpr int kObjThreads = 256, kObjMinBlocks = 4;   // 8 warps per object, <= 64 registers: every object CTA of C3 is resident at once
// `stage` is passed by the caller either as the shared-memory array itself (so that, after inlining, the accesses compile to
// LDS / STS) or as the global overflow area of an object with more than MAXS pose slots.
template <int T>
__device__ __forceinline__ void obj_schur_body(const EArgs& A, int e, uint32_t b0, uint32_t b1, int ns, double* stage, const uint32_t* pb,
                                               const double* hinv_s, const double* g_s, const uint32_t* __restrict__ su_ptr,
                                               double* __restrict__ S_upper, double* __restrict__ gp,
                                               double* __restrict__ hpp_diag, double* __restrict__ b_schur, bool fence) {
  constexpr int NE = 7, KR = 4, CH = KR * (6 + NE + 1), SL = 12 * NE;
  // ---- phase B
  const uint32_t nitems = (b1 - b0) * 6;
  for (uint32_t it = threadIdx.x; it < nitems; it += T) {
    const uint32_t q = b0 + it / 6;
    const int a = (int)(it % 6);
    const uint16_t sl = A.slot[q];
    if (sl == 0xFFFF) continue;
    if (q > b0 && A.slot[q - 1] == sl) continue;  // not the head of its run
    double Wm[NE], hp[6], gq = 0.0;
#pragma unroll
    for (int c = 0; c < NE; c++) Wm[c] = 0.0;
#pragma unroll
    for (int b = 0; b < 6; b++) hp[b] = 0.0;
    for (uint32_t q2 = q; q2 < b1 && A.slot[q2] == sl; q2++) {
      const double* ch = A.J + (size_t)A.pos[q2] * CH;
#pragma unroll
      for (int k = 0; k < KR; k++) {
        const double jpa = ch[k * 6 + a];
#pragma unroll
        for (int c = 0; c < NE; c++) Wm[c] += jpa * ch[KR * 6 + k * NE + c];
        gq += jpa * ch[KR * (6 + NE) + k];
#pragma unroll
        for (int b = 0; b < 6; b++) hp[b] += jpa * ch[k * 6 + b];
      }
    }
    const int fi = A.f[q];
    double* st = stage + (size_t)sl * SL;
    double zg = 0.0;
#pragma unroll
    for (int c = 0; c < NE; c++) {
      double z = 0.0;
#pragma unroll
      for (int d = 0; d < NE; d++) z += Wm[d] * hinv_s[d * NE + c];
      st[a * NE + c] = z;
      st[6 * NE + a * NE + c] = Wm[c];
      zg += z * g_s[c];
    }
    red_add(&b_schur[6 * fi + a], -zg);
    red_add(&gp[6 * fi + a], gq);
    double* Sd = S_upper + (size_t)su_ptr[fi] * 36;
#pragma unroll
    for (int b = 0; b < 6; b++) {
      if (b < a) continue;
      red_add(&Sd[a * 6 + b], hp[b]);
      if (a != b) red_add(&Sd[b * 6 + a], hp[b]);
      else red_add(&hpp_diag[6 * fi + a], hp[b]);
    }
  }
  if (fence) __threadfence_block();
  asm volatile("cp.async.wait_all;" ::: "memory");   // the pair -> S-block table staged by the caller (shared-memory path)
  __syncthreads();
  // ---- phase C: pair p = (a, b >= a) in row-major order of the upper triangle, six threads (rows of the 6x6 block) per pair
  const int items = ns * (ns + 1) / 2 * 6;
  const double tn = 2.0 * ns + 1.0;
  for (int it = threadIdx.x; it < items; it += T) {
    const int pr = it / 6, row = it - 6 * pr;
    // first pair of row a has index a (2 ns - a + 1) / 2: invert with a square root, then correct the rounding
    int a = (int)((tn - sqrt(tn * tn - 8.0 * pr)) * 0.5);
    a = max(0, min(a, ns - 1));
    while (a > 0 && a * (2 * ns - a + 1) / 2 > pr) a--;
    while ((a + 1) * (2 * ns - a) / 2 <= pr) a++;
    const int b = a + pr - a * (2 * ns - a + 1) / 2;
    const double* Za = stage + (size_t)a * SL + row * NE;
    const double* Wb = stage + (size_t)b * SL + 6 * NE;
    double z[NE];
#pragma unroll
    for (int d = 0; d < NE; d++) z[d] = Za[d];
    double* Sb = S_upper + (size_t)pb[pr] * 36 + row * 6;
#pragma unroll
    for (int c = 0; c < 6; c++) {
      double v = 0.0;
#pragma unroll
      for (int d = 0; d < NE; d++) v += z[d] * Wb[c * NE + d];
      red_add(&Sb[c], -v);
    }
  }
}

template <int T, int MAXS, int MINB>
__global__ void __launch_bounds__(T, MINB) obj_schur_kernel(EArgs A, const uint32_t* __restrict__ su_ptr, double* __restrict__ S_upper,
                                                             double* __restrict__ gp, double* __restrict__ hpp_diag,
                                                             double* __restrict__ b_schur) {
  constexpr int NE = 7, SL = 12 * NE;
  __shared__ double hinv_s[NE * NE];
  __shared__ double g_s[NE];
  __shared__ double stage_s[MAXS * SL];
  __shared__ uint32_t pb_s[MAXS * (MAXS + 1) / 2];
  const int e = (int)blockIdx.x;
  if (A.cst[e]) return;
  const uint32_t b0 = A.ptr[e], b1 = A.ptr[e + 1];
  const int ns = A.nslots[e];
  if (b0 == b1 || ns == 0) return;
  const uint32_t* pb = A.pair_blk + A.pair_ptr[e];
  if (ns <= MAXS) {
    // the S-block index of every slot pair is needed only in phase C: fetch the table asynchronously now (LDGSTS), so that
    // its DRAM latency is hidden behind phase B instead of being paid by every phase-C iteration
    const int npairs = ns * (ns + 1) / 2;
    for (int i = threadIdx.x; i < npairs; i += T)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(&pb_s[i])), "l"(pb + i) : "memory");
  }
  if (threadIdx.x < NE * NE) hinv_s[threadIdx.x] = A.einv[(size_t)e * NE * NE + threadIdx.x];
  if (threadIdx.x < NE) g_s[threadIdx.x] = A.eg[(size_t)e * NE + threadIdx.x];
  __syncthreads();
  if (ns <= MAXS) obj_schur_body<T>(A, e, b0, b1, ns, stage_s, pb_s, hinv_s, g_s, su_ptr, S_upper, gp, hpp_diag, b_schur, false);
  else obj_schur_body<T>(A, e, b0, b1, ns, A.overflow + A.overflow_off[e], pb, hinv_s, g_s, su_ptr, S_upper, gp, hpp_diag, b_schur, true);
}



// ------------------------------------------------------------------------------------------ points: row-owner elimination
// Two kernels eliminate the points:
//   point_prep_kernel  - per point H_ll, g_l, damping, the 3x3 inverse Hinv and its Cholesky factor F (Hinv = F F^T); per
//                        (point, keyframe) slot the record [U = (sum Jp^T Jl) F (6x3) | w = F^T g_l (3) | pad] written to `WZ`
//                        (24 doubles = 192 bytes per slot).  With U the Schur update is SYMMETRIC, S_ab -= U_a U_b^T, and the
//                        right-hand side b_a -= U_a w, so one 6x3 matrix per slot serves as both operands.
//   schur_rows_kernel  - a warp per work item = (rows 2m and 2m + 1 of the reduced matrix, 5 consecutive column offsets,
//                        <= 128 points with a slot in either row): per point six slot records are loaded and feed up to ten
//                        DMMAs (A = U_row, B = U_col^T, one per block pair); the 2 x 5 6x6 accumulators live in REGISTERS
//                        for the whole item and are flushed once.
// The Jacobian chunks are point-major, so a point's chunks are one contiguous range: point_prep and the back-substitution
// stream them through shared memory with per-warp double-buffered TMA bulk copies (WarpChunkPipe below) -- every global-load
// latency (list pointers, group records, chunks) is taken one batch ahead of its use.
constexpr int kWZ = 24;
constexpr int kWZw = 18;                  // offset of w in the slot record
constexpr uint32_t kNoSlot = 0xFFFFFFFFu;
struct RowGroupD { uint32_t pos0; int32_t f; uint32_t gs, cnt; };   // chunks pos0 .. pos0 + cnt - 1, pose f index (-1 constant), slot
struct RowItemD { uint32_t row, dlo, off, cnt; };

// Two shapes of the per-warp pipeline fill the shared memory of an SM (one CTA per SM either way):
//   <8 warps, 2 stages>  : every warp double-buffers (its next batch lands while it computes the current one);
//   <16 warps, 1 stage>  : twice the warps, each waits for its own copy -- the other 15 cover the wait.
constexpr int kStageChunks = 96;          // chunks per stage: 12 KB; a batch = 4 consecutive points (~80 chunks on average)
constexpr int pipe_smem(int warps, int stages) { return warps * stages * kStageChunks * kChunk * 8 + 128; }
static_assert(pipe_smem(8, 2) <= 227 * 1024 && pipe_smem(16, 1) <= 227 * 1024, "stages must fit the shared memory of one SM");

// Tables of one batch (points 4 b .. 4 b + 3), spread over the lanes: lane j < 5 holds chunk-list pointer and group pointer of
// point 4 b + j, lane j < 4 the `regular` flag.
struct PipeInfo { uint32_t cptr, gptr, reg; };
__device__ __forceinline__ PipeInfo pipe_load_info(const uint32_t* __restrict__ ptr, const uint32_t* __restrict__ grp_ptr,
                                                   const uint8_t* __restrict__ regular, int b, int ne, int lane) {
  // Every lane loads (clamped indices, the values of lanes >= 5 are never used): a predicated load into registers that were
  // first set to a default makes the compiler re-initialise them AFTER the load on some paths, and that write waits for the
  // load -- the prefetch then costs a full global-memory latency per batch (measured: 16 % of point_prep's stall samples).
  PipeInfo r;
  const int ec = min(4 * b + lane, ne);
  r.cptr = ptr[ec]; r.gptr = grp_ptr[ec]; r.reg = regular[min(ec, ne - 1)];
  return r;
}
// Stage layout of a batch: the chunk ranges of its regular points back to back, in point order; a point whose range does
// not fit into what is left of the stage is not staged (its lanes read global memory).  Same arithmetic on both sides.
struct PipeLayout { uint32_t p[5]; uint32_t off[4]; bool staged[4]; };
struct PipeMine { uint32_t p, off; bool staged; };   // the entries of the lane's own point (selected without dynamic indexing)
__device__ __forceinline__ PipeMine pipe_mine(const PipeLayout& L, int pt) {
  PipeMine m;
  m.p = pt == 0 ? L.p[0] : (pt == 1 ? L.p[1] : (pt == 2 ? L.p[2] : L.p[3]));
  m.off = pt == 0 ? L.off[0] : (pt == 1 ? L.off[1] : (pt == 2 ? L.off[2] : L.off[3]));
  m.staged = pt == 0 ? L.staged[0] : (pt == 1 ? L.staged[1] : (pt == 2 ? L.staged[2] : L.staged[3]));
  return m;
}
__device__ __forceinline__ PipeLayout pipe_layout(const PipeInfo& I) {
  PipeLayout L;
#pragma unroll
  for (int j = 0; j < 5; j++) L.p[j] = __shfl_sync(0xffffffffu, I.cptr, j);
  uint32_t off = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const uint32_t rg = __shfl_sync(0xffffffffu, I.reg, j);
    const uint32_t cnt = L.p[j + 1] - L.p[j];
    L.staged[j] = rg != 0u && cnt != 0u && off + cnt <= (uint32_t)kStageChunks;
    L.off[j] = off;
    if (L.staged[j]) off += cnt;
  }
  return L;
}
__device__ __forceinline__ void pipe_issue(const PipeLayout& L, const double* __restrict__ J, double* stage, uint64_t* bar, int lane) {
  if (lane == 0) {
    uint32_t bytes = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) if (L.staged[j]) bytes += (L.p[j + 1] - L.p[j]) * (uint32_t)(kChunk * 8);
    mbar_expect_tx(bar, bytes);
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (L.staged[j]) tma_load_1d(stage + (size_t)L.off[j] * kChunk, J + (size_t)L.p[j] * kChunk, (L.p[j + 1] - L.p[j]) * (uint32_t)(kChunk * 8), bar);
  }
}
// 3x3 lower Cholesky factor F of an SPD matrix given row-major (F F^T = A); tiny negative pivots from rounding clamp to 0
__device__ __forceinline__ void chol3(const double* A, double* F /* f00 f10 f11 f20 f21 f22 */) {
  F[0] = sqrt(fmax(A[0], 0.0));
  const double i0 = F[0] > 0.0 ? 1.0 / F[0] : 0.0;
  F[1] = A[3] * i0; F[3] = A[6] * i0;
  F[2] = sqrt(fmax(A[4] - F[1] * F[1], 0.0));
  const double i1 = F[2] > 0.0 ? 1.0 / F[2] : 0.0;
  F[4] = (A[7] - F[3] * F[1]) * i1;
  F[5] = sqrt(fmax(A[8] - F[3] * F[3] - F[4] * F[4], 0.0));
}

template <int WARPS, int STAGES>
__global__ void __launch_bounds__(32 * WARPS, 1) point_prep_kernel(EArgs A, const uint32_t* __restrict__ grp_ptr,
                                                                         const uint4* __restrict__ grp, const uint8_t* __restrict__ regular,
                                                                         LMParams lm, double* __restrict__ WZ, double* __restrict__ scalars) {
  extern __shared__ __align__(128) unsigned char pipe_raw[];
  __shared__ uint64_t bars[WARPS][2];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int pt = lane >> 3, sub = lane & 7;
  unsigned char* base = pipe_raw + ((128u - (smem_addr(pipe_raw) & 127u)) & 127u);
  double* stage0 = reinterpret_cast<double*>(base) + (size_t)(STAGES * w) * kStageChunks * kChunk;
  double* stage1 = STAGES == 2 ? stage0 + (size_t)kStageChunks * kChunk : stage0;
  if (lane == 0) { mbar_init(&bars[w][0], 1); mbar_init(&bars[w][1], 1); }
  __syncwarp();
  const int nb = (A.ne + 3) / 4;          // (A.ne stops at the last regular point of this rank, batch_begin starts at its first one)
  const int W = gridDim.x * WARPS;
  int b = A.batch_begin + blockIdx.x * WARPS + w;
  double gmax = 0.0;
  int nfail = 0;
  if (b >= nb) return;
  // prologue: batch b staged, tables of batch b + W requested
  PipeInfo Icur = pipe_load_info(A.ptr, grp_ptr, regular, b, A.ne, lane);
  PipeLayout Lcur = pipe_layout(Icur);
  pipe_issue(Lcur, A.J, stage0, &bars[w][0], lane);
  PipeInfo Inext = pipe_load_info(A.ptr, grp_ptr, regular, min(b + W, nb), A.ne, lane);
  // group records of this lane for the current batch (groups sub and sub + 8 of its point), fetched one batch ahead
  // (unconditional loads at clamped indices, see pipe_load_info; records past the point's range are dropped where used)
  const uint32_t ngrp = grp_ptr[A.ne];
  if (ngrp == 0u) return;                 // no regular point
  auto load_groups = [&](const PipeInfo& I, uint4& G0, uint4& G1) {
    const uint32_t g0 = __shfl_sync(0xffffffffu, I.gptr, pt);
    G0 = grp[min(g0 + sub, ngrp - 1u)];
    G1 = grp[min(g0 + sub + 8u, ngrp - 1u)];
  };
  uint4 G0, G1, G0n, G1n;
  load_groups(Icur, G0, G1);
  G0n = G0; G1n = G1;
  for (int it = 0; b < nb; b += W, it++) {
    const int cur = STAGES == 2 ? (it & 1) : 0;
    const bool more = b + W < nb;
    PipeLayout Lnext = Lcur;
    if (more) Lnext = pipe_layout(Inext);
    if (STAGES == 2 && more) {   // stage the next batch (its tables arrived during the previous iteration) and request the one after
      fence_proxy_async_smem();
      __syncwarp();
      pipe_issue(Lnext, A.J, cur ? stage0 : stage1, cur ? &bars[w][0] : &bars[w][1], lane);
    }
    if (more) load_groups(Inext, G0n, G1n);
    PipeInfo Iafter = pipe_load_info(A.ptr, grp_ptr, regular, min(b + 2 * W, nb), A.ne, lane);
    const int e = 4 * b + pt;
    const uint32_t reg_pt = __shfl_sync(0xffffffffu, Icur.reg, pt);   // (outside the && below: every lane must take part in the shuffle)
    const bool act = e < A.ne && reg_pt != 0u;
    // Inactive lanes (points past the end, points left to the generic kernels) run the SAME code on zero groups and skip the
    // stores, so that the warp stays converged through the shuffles and warp barriers of the iteration.
    const uint32_t g0 = __shfl_sync(0xffffffffu, Icur.gptr, pt);
    const uint32_t g1s = __shfl_sync(0xffffffffu, Icur.gptr, pt + 1);
    const uint32_t g1 = act ? g1s : g0;
    if (!(g0 + sub < g1)) G0 = make_uint4(0u, 0u, kNoSlot, 0u);
    if (!(g0 + sub + 8u < g1)) G1 = make_uint4(0u, 0u, kNoSlot, 0u);
    // per-point data that does not depend on the chunks: requested before the wait
    double s[3] = {1.0, 1.0, 1.0};
    if (act && !lm.compute_scale) { s[0] = A.escale[(size_t)e * 3]; s[1] = A.escale[(size_t)e * 3 + 1]; s[2] = A.escale[(size_t)e * 3 + 2]; }
    const PipeMine M = pipe_mine(Lcur, pt);
    // chunk `pos` of a staged point sits at cbase + pos * kChunk (derived from the shared-memory pointer only, so that the
    // staged path compiles to shared-memory loads)
    const double* cbase = (cur ? stage1 : stage0) + ((ptrdiff_t)M.off - (ptrdiff_t)M.p) * kChunk;
    // The choice between the shared-memory path and the global one is made PER WARP (uniform), so that the staged path keeps
    // shared-memory addressing (LDS) without lane-divergent copies of the sweeps.
    const bool staged = __all_sync(0xffffffffu, M.staged || !act);
    const double* gbase = M.staged ? cbase : A.J;          // mixed warp (a point did not fit into the stage): generic addressing
    mbar_wait_warp(cur ? &bars[w][1] : &bars[w][0], (uint32_t)(STAGES == 2 ? ((it >> 1) & 1) : (it & 1)));
    double H[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
    auto sweep_hg = [&](const double* cb, const uint4 G) {
      for (uint32_t k = 0; k < G.w; k++) {
        double jp[12], jl[6];
        double2 rv;
        load_chunk(cb + (size_t)(G.x + k) * kChunk, jp, jl, rv.x, rv.y);
        int t = 0;
#pragma unroll
        for (int a = 0; a < 3; a++) {
          g[a] += jl[a] * rv.x + jl[3 + a] * rv.y;
#pragma unroll
          for (int c = a; c < 3; c++) H[t++] += jl[a] * jl[c] + jl[3 + a] * jl[3 + c];
        }
      }
    };
    auto all_groups = [&](auto&& fn) {   // fn (record) for every group of this lane: the two prefetched ones, then the rare rest
      if (G0.w) fn(G0);
      if (G1.w) fn(G1);
      for (uint32_t gi = g0 + sub + 16; gi < g1; gi += 8) fn(grp[gi]);
    };
    if (staged) all_groups([&](const uint4 G) { sweep_hg(cbase, G); });
    else all_groups([&](const uint4 G) { sweep_hg(gbase, G); });
    __syncwarp();
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) {
#pragma unroll
      for (int a = 0; a < 6; a++) H[a] += __shfl_xor_sync(0xffffffffu, H[a], d);
#pragma unroll
      for (int a = 0; a < 3; a++) g[a] += __shfl_xor_sync(0xffffffffu, g[a], d);
    }
    if (act && A.prior_H) {
      const double* ph = A.prior_H + (size_t)e * 9;
      H[0] += ph[0]; H[1] += ph[1]; H[2] += ph[2]; H[3] += ph[4]; H[4] += ph[5]; H[5] += ph[8];
#pragma unroll
      for (int a = 0; a < 3; a++) g[a] += A.prior_g[(size_t)e * 3 + a];
    }
    const double hd[3] = {H[0], H[3], H[5]};
#pragma unroll
    for (int a = 0; a < 3; a++) {
      if (lm.compute_scale) s[a] = 1.0 / (1.0 + sqrt(hd[a]));
      gmax = fmax(gmax, fabs(g[a]));
    }
    double Hs[9], hinv[9], F[6];
    Hs[0] = s[0] * H[0] * s[0]; Hs[1] = Hs[3] = s[0] * H[1] * s[1]; Hs[2] = Hs[6] = s[0] * H[2] * s[2];
    Hs[4] = s[1] * H[3] * s[1]; Hs[5] = Hs[7] = s[1] * H[4] * s[2]; Hs[8] = s[2] * H[5] * s[2];
#pragma unroll
    for (int a = 0; a < 3; a++) Hs[4 * a] += fmin(fmax(Hs[4 * a], lm.min_diag), lm.max_diag) * lm.inv_radius;
    const bool ok = spd_inverse3_cofactor(Hs, hinv);
    if (!ok) {
#pragma unroll
      for (int a = 0; a < 9; a++) hinv[a] = 0.0;
    } else {
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
        for (int c = 0; c < 3; c++) hinv[3 * a + c] *= s[a] * s[c];
    }
    if (act && sub == 0) {
      if (!ok) nfail++;
      if (lm.compute_scale && ok) { A.escale[(size_t)e * 3] = s[0]; A.escale[(size_t)e * 3 + 1] = s[1]; A.escale[(size_t)e * 3 + 2] = s[2]; }
#pragma unroll
      for (int a = 0; a < 9; a++) A.einv[(size_t)e * 9 + a] = hinv[a];
#pragma unroll
      for (int a = 0; a < 3; a++) A.eg[(size_t)e * 3 + a] = g[a];
    }
    chol3(hinv, F);
    const double wv[3] = {F[0] * g[0] + F[1] * g[1] + F[3] * g[2], F[2] * g[1] + F[4] * g[2], F[5] * g[2]};   // F^T g
    auto emit = [&](const double* cb, const uint4 G) {
      double Wm[18];
#pragma unroll
      for (int a = 0; a < 18; a++) Wm[a] = 0.0;
      for (uint32_t k = 0; k < G.w; k++) {
        double jp[12], jl[6];
        double2 rv;
        load_chunk(cb + (size_t)(G.x + k) * kChunk, jp, jl, rv.x, rv.y);
#pragma unroll
        for (int a = 0; a < 6; a++)
#pragma unroll
          for (int c = 0; c < 3; c++) Wm[3 * a + c] += jp[a] * jl[c] + jp[6 + a] * jl[3 + c];
      }
      double U[18];
#pragma unroll
      for (int a = 0; a < 6; a++) {
        U[3 * a] = Wm[3 * a] * F[0] + Wm[3 * a + 1] * F[1] + Wm[3 * a + 2] * F[3];
        U[3 * a + 1] = Wm[3 * a + 1] * F[2] + Wm[3 * a + 2] * F[4];
        U[3 * a + 2] = Wm[3 * a + 2] * F[5];
      }
      if (G.z != kNoSlot) {
        double2* rec = reinterpret_cast<double2*>(WZ + (size_t)G.z * kWZ);
#pragma unroll
        for (int a = 0; a < 9; a++) rec[a] = make_double2(U[2 * a], U[2 * a + 1]);
        rec[9] = make_double2(wv[0], wv[1]);
        rec[10] = make_double2(wv[2], 0.0);
      }
    };
    if (staged) all_groups([&](const uint4 G) { emit(cbase, G); });
    else all_groups([&](const uint4 G) { emit(gbase, G); });
    __syncwarp();
    if (STAGES == 1 && more) {     // single stage: the next batch is requested once this one has been consumed
      fence_proxy_async_smem();
      pipe_issue(Lnext, A.J, stage0, &bars[w][0], lane);
    }
    Icur = Inext; Inext = Iafter; Lcur = Lnext; G0 = G0n; G1 = G1n;
  }
  __syncwarp();
  gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, 16));
  gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, 8));
  gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, 4));
  gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, 2));
  gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, 1));
  nfail = __reduce_add_sync(0xffffffffu, nfail);
  if (lane == 0) {
    if (gmax > 0.0) atomic_max_nonneg(&scalars[SC_GMAX], gmax);
    if (nfail) atomicAdd(&scalars[SC_FAIL], (double)nfail);
  }
}

constexpr int kRowWarps = 4;
constexpr int kRowRec = 8;                      // records per ring stage: 6 column records + 2 row records
// A warp per work item.  Rows fa = 2 it.x and fb = fa + 1; accA[j] = sum U_fa U_(fa + dlo + j)^T, accB[j] = sum U_fb U_(fb + dlo + j)^T
// over the item's points.  The operands are STAGED by the copy engine: per entry one 1-D bulk copy brings the point's (up
// to) six consecutive slot records into a per-warp ring in shared memory (plus the two row records for ranges past the
// first), and the lanes read their fragment element from the ring -- no global-load latency is left on the DMMA chain.
// The ring has two halves of H entries, one mbarrier each; a half is refilled in ONE step by H lanes in parallel -- every
// lane already holds the descriptor of "its" entry (entry k lives in lane k % 32), computes that entry's addresses and
// issues its copies -- so the serial per-entry cost on the consuming side is a shuffle, eight LDS and the products.
// The products of an entry are selected by ONE warp-uniform switch on the number of records (a C++ `if` or a predicate
// around mma.sync costs four register moves and a WARPSYNC per product: 207 instructions per entry of at most ten
// products, measured); a missing row multiplies by a zero A operand instead.
#ifndef OBVI_ROW_MINB
#define OBVI_ROW_MINB 6
#endif
template <int H>
__global__ void __launch_bounds__(32 * kRowWarps, OBVI_ROW_MINB) schur_rows_kernel(const uint4* __restrict__ items, int n_items,
                                                                     const uint32_t* __restrict__ ent, const double* __restrict__ WZ,
                                                                     const uint32_t* __restrict__ rowblk, int row_span,
                                                                     double* __restrict__ S_upper, double* __restrict__ b_schur) {
  static_assert(H == 2 || H == 4 || H == 8, "a half of the ring is refilled by H lanes of one 32-entry chunk");
  extern __shared__ __align__(128) unsigned char rows_smem[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * kRowWarps + w;
  if (item >= n_items) return;
  constexpr int kStage = kRowRec * kWZ;          // doubles per entry
  double* ring = reinterpret_cast<double*>(rows_smem) + (size_t)w * 2 * H * kStage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(rows_smem + (size_t)kRowWarps * 2 * H * kStage * 8) + w * 2;
  if (lane == 0) { mbar_init(&bars[0], H); mbar_init(&bars[1], H); }
  // records of a stage that an entry does not copy keep what an earlier entry left there (finite, times a zero A operand);
  // before the first entry that must not be a NaN pattern
  for (int q = lane; q < 2 * H * kStage; q += 32) ring[q] = 0.0;
  fence_proxy_async_smem();
  __syncwarp();
  const uint4 it = items[item];
  const int frow = lane >> 2, fk = lane & 3;
  const bool fvalid = frow < 6 && fk < 3;
  double2 accA[5], accB[5];
#pragma unroll
  for (int j = 0; j < 5; j++) { accA[j] = make_double2(0.0, 0.0); accB[j] = make_double2(0.0, 0.0); }
  const uint32_t* ep = ent + it.z;
  const uint32_t cnt = it.w;
  const bool first_range = it.y == 0;
  // Fragment element of this lane inside a slot record: U[frow][fk] for the 6x3 block; fragment row 6 reads w[fk] (column 6
  // of the B operand of the DIAGONAL pair: C[.][6] = U_a w, the right-hand side contribution); everything else is clamped to
  // a finite element and multiplied by a zero of the A operand (k = 3 column) or lands in rows / columns 6-7 that are never
  // flushed.  Slots are dense, so with g the slot of row fa, column j of row fa is record g + dlo + j and column j of row fb is
  // record g + dlo + j + 1: six records serve both rows.
  const int offb = frow < 6 ? 3 * frow + (fk < 3 ? fk : 2) : (frow == 6 ? kWZw + (fk < 3 ? fk : 2) : 0);
  // entry descriptors: entries c0 + lane (e_lo) and c0 + 32 + lane (e_hi)
  uint32_t e_lo = lane < cnt ? ep[lane] : 0u, e_hi = 32u + lane < cnt ? ep[32 + lane] : 0u, c0 = 0u;
  auto refill = [&](uint32_t i0, int half) {                 // entries i0 .. i0 + H - 1 (i0 a multiple of H) -> stages of `half`
    const uint32_t q = ((uint32_t)lane - i0) & 31u;
    if (q < (uint32_t)H) {
      const uint32_t k = i0 + q;
      if (k < cnt) {
        const uint32_t e = (k - c0) < 32u ? e_lo : e_hi;
        const size_t g1 = e & 0x3ffffffu;                   // g + 1 (g = -1 when the very first point starts on an odd row)
        const uint32_t nl = (e >> 26) & 7u, va = (e >> 29) & 1u, vb = (e >> 30) & 1u;
        const uint32_t j0 = 1u - va;
        double* st = ring + (size_t)(half * H + q) * kStage;
        const uint32_t bytes_b = (nl - j0) * (uint32_t)(kWZ * 8);
        const uint32_t na = first_range ? 0u : va + vb;
        mbar_expect_tx(&bars[half], bytes_b + na * (uint32_t)(kWZ * 8));
        tma_load_1d(st + j0 * kWZ, WZ + (g1 - 1 + it.y + j0) * kWZ, bytes_b, &bars[half]);
        if (na) tma_load_1d(st + (6 + j0) * kWZ, WZ + (g1 - 1 + j0) * kWZ, na * (uint32_t)(kWZ * 8), &bars[half]);
      } else {
        mbar_arrive(&bars[half]);
      }
    }
  };
  refill(0u, 0);
  if ((uint32_t)H < cnt) refill((uint32_t)H, 1);
  const uint32_t maskA = fvalid ? (1u << 29) : 0u, maskB = fvalid ? (1u << 30) : 0u;   // row flag of the entry AND a real fragment element
  int half = 0; uint32_t ph = 0u;
  for (uint32_t i0 = 0; i0 < cnt; i0 += H) {
    if (i0 - c0 == 32u) { c0 += 32u; e_lo = e_hi; e_hi = c0 + 32u + lane < cnt ? ep[c0 + 32u + lane] : 0u; }
    mbar_wait_warp(&bars[half], ph);
    const uint32_t qn = min((uint32_t)H, cnt - i0);
#pragma unroll 1
    for (uint32_t q = 0; q < qn; q++) {
      const uint32_t k = i0 + q;
      const uint32_t e = __shfl_sync(0xffffffffu, (k - c0) < 32u ? e_lo : e_hi, (int)(k & 31u));
      const double* st = ring + (size_t)(half * H + q) * kStage + offb;
      double b[6];
#pragma unroll
      for (int j = 0; j < 6; j++) b[j] = st[j * kWZ];
      double aA = first_range ? b[0] : st[6 * kWZ], aB = first_range ? b[1] : st[7 * kWZ];
      aA = (e & maskA) ? aA : 0.0;
      aB = (e & maskB) ? aB : 0.0;
      switch ((e >> 26) & 7u) {
        case 6: dmma_m8n8k4(accB[4].x, accB[4].y, aB, b[5]);
        case 5: dmma_m8n8k4(accA[4].x, accA[4].y, aA, b[4]); dmma_m8n8k4(accB[3].x, accB[3].y, aB, b[4]);
        case 4: dmma_m8n8k4(accA[3].x, accA[3].y, aA, b[3]); dmma_m8n8k4(accB[2].x, accB[2].y, aB, b[3]);
        case 3: dmma_m8n8k4(accA[2].x, accA[2].y, aA, b[2]); dmma_m8n8k4(accB[1].x, accB[1].y, aB, b[2]);
        case 2: dmma_m8n8k4(accA[1].x, accA[1].y, aA, b[1]); dmma_m8n8k4(accB[0].x, accB[0].y, aB, b[1]);
        default: dmma_m8n8k4(accA[0].x, accA[0].y, aA, b[0]);
      }
    }
    // The half is free once the products above have ISSUED (their operands are in registers), which program order
    // guarantees for the whole warp: the refill needs no barrier of its own.
    __syncwarp();
    if (i0 + 2 * H < cnt) refill(i0 + 2 * H, half);
    half ^= 1; if (half == 0) ph ^= 1u;
  }
  const uint32_t fa = 2u * it.x;
  if (fvalid) {
    const uint32_t* rb = rowblk + (size_t)fa * row_span + it.y;      // rowblk holds 2 ceil(nf / 2) rows: fb's row always exists
    const int coff = 6 * frow + 2 * fk;
#pragma unroll
    for (int j = 0; j < 5; j++) {
      if ((int)it.y + j >= row_span) break;
      const uint32_t blkA = rb[j], blkB = rb[row_span + j];
      if (blkA != 0xFFFFFFFFu) {
        double* C = S_upper + (size_t)blkA * 36 + coff;
        if (accA[j].x != 0.0) atomicAdd(C, -accA[j].x);
        if (accA[j].y != 0.0) atomicAdd(C + 1, -accA[j].y);
      }
      if (blkB != 0xFFFFFFFFu) {
        double* C = S_upper + (size_t)blkB * 36 + coff;
        if (accB[j].x != 0.0) atomicAdd(C, -accB[j].x);
        if (accB[j].y != 0.0) atomicAdd(C + 1, -accB[j].y);
      }
    }
  }
  // column 6 of the diagonal products: lanes with fk == 3 hold C[frow][6]
  if (first_range && fk == 3 && frow < 6) {
    if (accA[0].x != 0.0) atomicAdd(&b_schur[6 * fa + frow], -accA[0].x);
    if (accB[0].x != 0.0) atomicAdd(&b_schur[6 * (fa + 1) + frow], -accB[0].x);
  }
}
constexpr size_t rows_smem_bytes(int H) { return (size_t)kRowWarps * (2 * H * kRowRec * kWZ * 8 + 16); }

// Back-substitution for one e-block + its share of the model cost change and of the candidate point:
//   delta_e = -Hinv (g_e + sum_obs Je^T (Jp delta_p)),  model += sum m (r + m/2), m = Jp delta_p + Je delta_e
template <int NE, int KR, int T>
__global__ void __launch_bounds__(T) backsub_eblock_kernel(EArgs A, const double* __restrict__ dpose,
                                                            const double* __restrict__ x, double* __restrict__ x_cand,
                                                            double* __restrict__ delta_e, double* __restrict__ scalars) {
  constexpr int CH = KR * (6 + NE + 1);
  __shared__ double red[33];
  const int e = blockIdx.x;
  const uint32_t b0 = A.ptr[e], b1 = A.ptr[e + 1];
  const bool cst = A.cst[e] != 0;
  double de[NE];
#pragma unroll
  for (int a = 0; a < NE; a++) de[a] = 0.0;
  if (!cst) {
    if (b0 == b1 && A.prior_H == nullptr) return;
    double t[NE];
#pragma unroll
    for (int a = 0; a < NE; a++) t[a] = 0.0;
    for (uint32_t q = b0 + threadIdx.x; q < b1; q += T) {
      const int fi = A.f[q];
      if (fi < 0) continue;
      double chunk_tmp_3[CH];
      const double* ch = chunk_full<NE, KR>(A.J, (size_t)A.pos[q], chunk_tmp_3);
      double dp[6];
#pragma unroll
      for (int a = 0; a < 6; a++) dp[a] = dpose[6 * fi + a];
#pragma unroll
      for (int k = 0; k < KR; k++) {
        double jd = 0.0;
#pragma unroll
        for (int a = 0; a < 6; a++) jd += ch[k * 6 + a] * dp[a];
#pragma unroll
        for (int c = 0; c < NE; c++) t[c] += ch[KR * 6 + k * NE + c] * jd;
      }
    }
#pragma unroll
    for (int a = 0; a < NE; a++) t[a] = block_sum_all<T>(t[a], red) + A.eg[(size_t)e * NE + a];
    const double* hinv = A.einv + (size_t)e * NE * NE;
#pragma unroll
    for (int a = 0; a < NE; a++) {
      double v = 0.0;
#pragma unroll
      for (int b = 0; b < NE; b++) v += hinv[a * NE + b] * t[b];
      de[a] = -v;
    }
    if (threadIdx.x == 0) {
      double s2 = 0.0;
#pragma unroll
      for (int a = 0; a < NE; a++) {
        delta_e[(size_t)e * NE + a] = de[a];
        x_cand[(size_t)e * NE + a] = x[(size_t)e * NE + a] + de[a];
        s2 += de[a] * de[a];
      }
      atomicAdd(&scalars[SC_STEP2], s2);
    }
  }
  double mc = 0.0;
  for (uint32_t q = b0 + threadIdx.x; q < b1; q += T) {
    const int fi = A.f[q];
    if (cst && fi < 0) continue;
    double chunk_tmp_4[CH];
    const double* ch = chunk_full<NE, KR>(A.J, (size_t)A.pos[q], chunk_tmp_4);
    double dp[6];
#pragma unroll
    for (int a = 0; a < 6; a++) dp[a] = fi >= 0 ? dpose[6 * fi + a] : 0.0;
#pragma unroll
    for (int k = 0; k < KR; k++) {
      double m = 0.0;
#pragma unroll
      for (int a = 0; a < 6; a++) m += ch[k * 6 + a] * dp[a];
#pragma unroll
      for (int c = 0; c < NE; c++) m += ch[KR * 6 + k * NE + c] * de[c];
      mc += m * (ch[KR * (6 + NE) + k] + 0.5 * m);
    }
  }
  mc = block_sum_all<T>(mc, red);
  if (threadIdx.x == 0 && mc != 0.0) atomicAdd(&scalars[SC_MODEL], mc);
}

// Back-substitution for points, one warp per point, lane = observation, the 160-byte chunk held in registers
// between the two passes (10 x 16-byte loads per observation).
__global__ void __launch_bounds__(128) backsub_points_kernel(EArgs A, int n_e, const double* __restrict__ dpose,
                                                             const double* __restrict__ x, double* __restrict__ x_cand,
                                                             double* __restrict__ delta_e, double* __restrict__ scalars) {
  __shared__ double s_mc[4], s_s2[4];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int ei = blockIdx.x * 4 + wib;
  const int e = ei < n_e ? (A.elist ? (int)A.elist[ei] : ei) : 0;
  double mc = 0.0, s2 = 0.0;
  if (ei < n_e) {
    const uint32_t b0 = A.ptr[e], b1 = A.ptr[e + 1];
    const bool cst = A.cst[e] != 0;
    if (b1 > b0 || (!cst && A.prior_H != nullptr)) {
      double t[3] = {0.0, 0.0, 0.0};
      // first 32 observations stay in registers; longer tracks loop (second pass re-reads)
      double jp[12], jl[6], r[2], dp[6];
      int fi = -1;
      const bool have = b0 + lane < b1;
      if (have) {
        const uint32_t q = b0 + lane;
        fi = A.f[q];
        load_chunk(A.J + (size_t)A.pos[q] * kChunk, jp, jl, r[0], r[1]);
#pragma unroll
        for (int a = 0; a < 6; a++) dp[a] = fi >= 0 ? dpose[6 * fi + a] : 0.0;
      }
      auto accum_t = [&](const double* jpq, const double* jlq, const double* dpq) {
        const double jd0 = jpq[0] * dpq[0] + jpq[1] * dpq[1] + jpq[2] * dpq[2] + jpq[3] * dpq[3] + jpq[4] * dpq[4] + jpq[5] * dpq[5];
        const double jd1 = jpq[6] * dpq[0] + jpq[7] * dpq[1] + jpq[8] * dpq[2] + jpq[9] * dpq[3] + jpq[10] * dpq[4] + jpq[11] * dpq[5];
#pragma unroll
        for (int c = 0; c < 3; c++) t[c] += jlq[c] * jd0 + jlq[3 + c] * jd1;
      };
      if (!cst) {
        if (have && fi >= 0) accum_t(jp, jl, dp);
        for (uint32_t q = b0 + 32 + lane; q < b1; q += 32) {
          const int f2 = A.f[q];
          if (f2 < 0) continue;
          double jp2[12], jl2[6], r2[2], d2[6];
          load_chunk(A.J + (size_t)A.pos[q] * kChunk, jp2, jl2, r2[0], r2[1]);
#pragma unroll
          for (int a = 0; a < 6; a++) d2[a] = dpose[6 * f2 + a];
          accum_t(jp2, jl2, d2);
        }
      }
      double de[3] = {0.0, 0.0, 0.0};
      if (!cst) {
#pragma unroll
        for (int a = 0; a < 3; a++) t[a] = warp_sum(t[a]) + A.eg[(size_t)e * 3 + a];
        const double* hinv = A.einv + (size_t)e * 9;
#pragma unroll
        for (int a = 0; a < 3; a++) de[a] = -(hinv[3 * a] * t[0] + hinv[3 * a + 1] * t[1] + hinv[3 * a + 2] * t[2]);
        if (lane == 0) {
#pragma unroll
          for (int a = 0; a < 3; a++) {
            delta_e[(size_t)e * 3 + a] = de[a];
            x_cand[(size_t)e * 3 + a] = x[(size_t)e * 3 + a] + de[a];
            s2 += de[a] * de[a];
          }
        }
      }
      auto accum_m = [&](const double* jpq, const double* jlq, const double* rq, const double* dpq) {
#pragma unroll
        for (int k = 0; k < 2; k++) {
          double m = jlq[3 * k] * de[0] + jlq[3 * k + 1] * de[1] + jlq[3 * k + 2] * de[2];
#pragma unroll
          for (int a = 0; a < 6; a++) m += jpq[6 * k + a] * dpq[a];
          mc += m * (rq[k] + 0.5 * m);
        }
      };
      if (have && !(cst && fi < 0)) accum_m(jp, jl, r, dp);
      for (uint32_t q = b0 + 32 + lane; q < b1; q += 32) {
        const int f2 = A.f[q];
        if (cst && f2 < 0) continue;
        double jp2[12], jl2[6], r2[2], d2[6];
        load_chunk(A.J + (size_t)A.pos[q] * kChunk, jp2, jl2, r2[0], r2[1]);
#pragma unroll
        for (int a = 0; a < 6; a++) d2[a] = f2 >= 0 ? dpose[6 * f2 + a] : 0.0;
        accum_m(jp2, jl2, r2, d2);
      }
    }
  }
  mc = warp_sum(mc);
  if (lane == 0) { s_mc[wib] = mc; s_s2[wib] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double m = s_mc[0] + s_mc[1] + s_mc[2] + s_mc[3], s = s_s2[0] + s_s2[1] + s_s2[2] + s_s2[3];
    if (m != 0.0) atomicAdd(&scalars[SC_MODEL], m);
    if (s != 0.0) atomicAdd(&scalars[SC_STEP2], s);
  }
}

// Single-pass back-substitution for the points the row-owner path handles (8 lanes per point, a lane per pose group; the
// chunks streamed through shared memory like in point_prep_kernel).
// With a = Jp delta_p per observation, the point's share of the model cost change expands to
//   sum m (r + m/2),  m = a + Jl delta_e   =   sum a.(r + a/2)  +  delta_e.(g_l + t)  +  delta_e^T H_ll delta_e / 2
// with t = sum Jl^T a, g_l = sum Jl^T r, H_ll = sum Jl^T Jl, so one sweep over the Jacobian chunks is enough (the generic
// kernel above sweeps twice: once for t, once for m after delta_e is known).
template <int WARPS, int STAGES>
__global__ void __launch_bounds__(32 * WARPS, 1) backsub_rows_kernel(EArgs A, const uint32_t* __restrict__ grp_ptr, const uint4* __restrict__ grp,
                                                                           const uint8_t* __restrict__ regular,
                                                                           const double* __restrict__ dpose, const double* __restrict__ x,
                                                                           double* __restrict__ x_cand, double* __restrict__ delta_e,
                                                                           double* __restrict__ scalars) {
  extern __shared__ __align__(128) unsigned char pipe_raw[];
  __shared__ uint64_t bars[WARPS][2];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int pt = lane >> 3, sub = lane & 7;
  unsigned char* base = pipe_raw + ((128u - (smem_addr(pipe_raw) & 127u)) & 127u);
  double* stage0 = reinterpret_cast<double*>(base) + (size_t)(STAGES * w) * kStageChunks * kChunk;
  double* stage1 = STAGES == 2 ? stage0 + (size_t)kStageChunks * kChunk : stage0;
  if (lane == 0) { mbar_init(&bars[w][0], 1); mbar_init(&bars[w][1], 1); }
  __syncwarp();
  const int nb = (A.ne + 3) / 4;          // (A.ne stops at the last regular point of this rank, batch_begin starts at its first one)
  const int W = gridDim.x * WARPS;
  int b = A.batch_begin + blockIdx.x * WARPS + w;
  if (b >= nb) return;
  double mc_acc = 0.0, s2_acc = 0.0;
  PipeInfo Icur = pipe_load_info(A.ptr, grp_ptr, regular, b, A.ne, lane);
  PipeLayout Lcur = pipe_layout(Icur);
  pipe_issue(Lcur, A.J, stage0, &bars[w][0], lane);
  PipeInfo Inext = pipe_load_info(A.ptr, grp_ptr, regular, min(b + W, nb), A.ne, lane);
  const uint32_t ngrp = grp_ptr[A.ne];
  if (ngrp == 0u) return;                 // no regular point
  auto load_groups = [&](const PipeInfo& I, uint4& G0, uint4& G1) {
    const uint32_t g0 = __shfl_sync(0xffffffffu, I.gptr, pt);
    G0 = grp[min(g0 + sub, ngrp - 1u)];
    G1 = grp[min(g0 + sub + 8u, ngrp - 1u)];
  };
  uint4 G0, G1, G0n, G1n;
  load_groups(Icur, G0, G1);
  G0n = G0; G1n = G1;
  for (int it = 0; b < nb; b += W, it++) {
    const int cur = STAGES == 2 ? (it & 1) : 0;
    const bool more = b + W < nb;
    PipeLayout Lnext = Lcur;
    if (more) Lnext = pipe_layout(Inext);
    if (STAGES == 2 && more) {
      fence_proxy_async_smem();
      __syncwarp();
      pipe_issue(Lnext, A.J, cur ? stage0 : stage1, cur ? &bars[w][0] : &bars[w][1], lane);
    }
    if (more) load_groups(Inext, G0n, G1n);
    PipeInfo Iafter = pipe_load_info(A.ptr, grp_ptr, regular, min(b + 2 * W, nb), A.ne, lane);
    const int e = 4 * b + pt;
    const uint32_t reg_pt = __shfl_sync(0xffffffffu, Icur.reg, pt);   // (outside the && below: every lane must take part in the shuffle)
    const bool act = e < A.ne && reg_pt != 0u;
    const uint32_t g0 = __shfl_sync(0xffffffffu, Icur.gptr, pt), g1 = __shfl_sync(0xffffffffu, Icur.gptr, pt + 1);
    if (!(act && g0 + sub < g1)) G0 = make_uint4(0u, 0xFFFFFFFFu, kNoSlot, 0u);
    if (!(act && g0 + sub + 8u < g1)) G1 = make_uint4(0u, 0xFFFFFFFFu, kNoSlot, 0u);
    // pose steps of the two prefetched groups + the point's inverse / gradient: requested before the wait
    double dp0[6], dp1[6];
#pragma unroll
    for (int a = 0; a < 6; a++) {
      dp0[a] = (act && G0.w && (int)G0.y >= 0) ? dpose[6 * (size_t)G0.y + a] : 0.0;
      dp1[a] = (act && G1.w && (int)G1.y >= 0) ? dpose[6 * (size_t)G1.y + a] : 0.0;
    }
    double hinv[9], eg[3];
    if (act && sub == 0) {
#pragma unroll
      for (int a = 0; a < 9; a++) hinv[a] = A.einv[(size_t)e * 9 + a];
#pragma unroll
      for (int a = 0; a < 3; a++) eg[a] = A.eg[(size_t)e * 3 + a];
    }
    const PipeMine M = pipe_mine(Lcur, pt);
    const double* cbase = (cur ? stage1 : stage0) + ((ptrdiff_t)M.off - (ptrdiff_t)M.p) * kChunk;
    mbar_wait_warp(cur ? &bars[w][1] : &bars[w][0], (uint32_t)(STAGES == 2 ? ((it >> 1) & 1) : (it & 1)));
    double H[6] = {0, 0, 0, 0, 0, 0}, gl[3] = {0, 0, 0}, t[3] = {0, 0, 0}, sa = 0.0;
    auto sweep = [&](const double* cb, const uint4 G, const double* dp) {
      for (uint32_t k = 0; k < G.w; k++) {
        double jp[12], jl[6];
        double2 rv;
        load_chunk(cb + (size_t)(G.x + k) * kChunk, jp, jl, rv.x, rv.y);
        const double a0 = jp[0] * dp[0] + jp[1] * dp[1] + jp[2] * dp[2] + jp[3] * dp[3] + jp[4] * dp[4] + jp[5] * dp[5];
        const double a1 = jp[6] * dp[0] + jp[7] * dp[1] + jp[8] * dp[2] + jp[9] * dp[3] + jp[10] * dp[4] + jp[11] * dp[5];
        sa += a0 * (rv.x + 0.5 * a0) + a1 * (rv.y + 0.5 * a1);
        int q = 0;
#pragma unroll
        for (int a = 0; a < 3; a++) {
          t[a] += jl[a] * a0 + jl[3 + a] * a1;
          gl[a] += jl[a] * rv.x + jl[3 + a] * rv.y;
#pragma unroll
          for (int c = a; c < 3; c++) H[q++] += jl[a] * jl[c] + jl[3 + a] * jl[3 + c];
        }
      }
    };
    auto rest = [&](const double* cb) {   // groups beyond the two prefetched ones (tracks over more than 16 keyframes)
      for (uint32_t gi = g0 + sub + 16; gi < g1; gi += 8) {
        const uint4 G = grp[gi];
        double dp[6];
#pragma unroll
        for (int a = 0; a < 6; a++) dp[a] = (int)G.y >= 0 ? dpose[6 * (size_t)G.y + a] : 0.0;
        sweep(cb, G, dp);
      }
    };
    const bool staged = __all_sync(0xffffffffu, M.staged || !act);   // per-warp choice, see point_prep_kernel
    const double* gbase = M.staged ? cbase : A.J;
    if (staged) { if (act) { if (G0.w) sweep(cbase, G0, dp0); if (G1.w) sweep(cbase, G1, dp1); rest(cbase); } }
    else { if (act) { if (G0.w) sweep(gbase, G0, dp0); if (G1.w) sweep(gbase, G1, dp1); rest(gbase); } }
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) {
#pragma unroll
      for (int a = 0; a < 6; a++) H[a] += __shfl_xor_sync(0xffffffffu, H[a], d);
#pragma unroll
      for (int a = 0; a < 3; a++) { gl[a] += __shfl_xor_sync(0xffffffffu, gl[a], d); t[a] += __shfl_xor_sync(0xffffffffu, t[a], d); }
      sa += __shfl_xor_sync(0xffffffffu, sa, d);
    }
    if (act && sub == 0) {
      double u[3], de[3];
#pragma unroll
      for (int a = 0; a < 3; a++) u[a] = t[a] + eg[a];
#pragma unroll
      for (int a = 0; a < 3; a++) {
        de[a] = -(hinv[3 * a] * u[0] + hinv[3 * a + 1] * u[1] + hinv[3 * a + 2] * u[2]);
        delta_e[(size_t)e * 3 + a] = de[a];
        x_cand[(size_t)e * 3 + a] = x[(size_t)e * 3 + a] + de[a];
        s2_acc += de[a] * de[a];
      }
      const double hd0 = H[0] * de[0] + H[1] * de[1] + H[2] * de[2];
      const double hd1 = H[1] * de[0] + H[3] * de[1] + H[4] * de[2];
      const double hd2 = H[2] * de[0] + H[4] * de[1] + H[5] * de[2];
      mc_acc += sa + de[0] * (gl[0] + t[0] + 0.5 * hd0) + de[1] * (gl[1] + t[1] + 0.5 * hd1) + de[2] * (gl[2] + t[2] + 0.5 * hd2);
    }
    __syncwarp();
    if (STAGES == 1 && more) {     // single stage: the next batch is requested once this one has been consumed
      fence_proxy_async_smem();
      pipe_issue(Lnext, A.J, stage0, &bars[w][0], lane);
    }
    Icur = Inext; Inext = Iafter; Lcur = Lnext; G0 = G0n; G1 = G1n;
  }
  mc_acc = warp_sum(mc_acc); s2_acc = warp_sum(s2_acc);
  if (lane == 0) {
    if (mc_acc != 0.0) atomicAdd(&scalars[SC_MODEL], mc_acc);
    if (s2_acc != 0.0) atomicAdd(&scalars[SC_STEP2], s2_acc);
  }
}

// ------------------------------------------------------------------------------------------ bbox observations
constexpr int kBBoxChunk = 56;
// mode 0: residual + Jacobians into the chunk, cost; mode 1: candidate cost only
__global__ void bbox_kernel(const BBoxRec* __restrict__ rec, int64_t n, const PoseCam* __restrict__ pcam, int C,
                            const double* __restrict__ objs, int mode, int apply_loss, double* __restrict__ J,
                            double* __restrict__ scalars) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double cost = 0.0, fixed = 0.0;
  if (i < n) {
    const BBoxRec& R = rec[i];
    const PoseCam& pc = pcam[(size_t)R.pose * C + R.cam];
    double ell[7];
#pragma unroll
    for (int a = 0; a < 7; a++) ell[a] = objs[7 * (size_t)R.obj + a];
    double r[4];
    double* ch = J + (size_t)i * kBBoxChunk;
    if (mode == 0) bbox_residual_jacobian(pc, ell, R.A4, R.brect, R.invalid_err, r, ch + 24, ch);
    else bbox_residual_jacobian(pc, ell, R.A4, R.brect, R.invalid_err, r, nullptr, nullptr);
    const double s = r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3];
    double sc = 1.0, c = 0.5 * s;
    if (apply_loss && R.huber > 0.0) c = huber(R.huber, s, &sc);
    if (R.flags & kObsMasked) {   // removed in place: an all-zero block, no cost
      if (mode == 0) for (int a = 0; a < kBBoxChunk; a++) ch[a] = 0.0;
    } else {
      if (mode == 0) {
        for (int a = 0; a < 52; a++) ch[a] *= sc;
        for (int a = 0; a < 4; a++) ch[52 + a] = sc * r[a];
      }
      if ((R.flags & 3u) == 3u) fixed = c; else cost = c;
    }
  }
  // one atomic per warp instead of one per observation (they all hit the same two scalars)
  cost = warp_sum(cost); fixed = warp_sum(fixed);
  if ((threadIdx.x & 31) == 0) {
    if (cost != 0.0) atomicAdd(&scalars[mode == 0 ? SC_COST : SC_CAND], cost);
    if (fixed != 0.0) atomicAdd(&scalars[mode == 0 ? SC_FIXED : SC_CAND_FIXED], fixed);
  }
}

// ------------------------------------------------------------------------------------------ unary factors
// r = A (x[off:off+k] - mean).  mode 0: evaluate at x (store r, scale; cost); 1: accumulate J^T J / J^T r into the
// block's prior arrays (e-blocks) or the reduced system (poses); 2: model cost change; 3: candidate cost.
struct UnaryOut { double r[7]; double sc; };
__global__ void unary_kernel(const UnaryRec* __restrict__ rec, int64_t n, int mode, int apply_loss,
                             const double* __restrict__ poses, const double* __restrict__ points,
                             const double* __restrict__ objs, UnaryOut* __restrict__ out,
                             const int32_t* __restrict__ f_of_pose, const uint32_t* __restrict__ su_ptr,
                             double* __restrict__ S_upper, double* __restrict__ gp, double* __restrict__ hpp_diag,
                             double* __restrict__ prior_H_pt, double* __restrict__ prior_g_pt,
                             double* __restrict__ prior_H_obj, double* __restrict__ prior_g_obj,
                             const double* __restrict__ dpose, const double* __restrict__ dpoint,
                             const double* __restrict__ dobj, double* __restrict__ scalars) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const UnaryRec& R = rec[i];
  const int k = R.k, off = R.off;
  const int bs = R.kind == 0 ? 6 : (R.kind == 1 ? 3 : 7);
  if (mode == 0 || mode == 3) {
    const double* x = (R.kind == 0 ? poses + 6 * (size_t)R.idx : (R.kind == 1 ? points + 3 * (size_t)R.idx : objs + 7 * (size_t)R.idx));
    double r[7], s = 0.0;
    for (int a = 0; a < k; a++) {
      double v = 0.0;
      for (int c = 0; c < k; c++) v += R.A[a * k + c] * (x[off + c] - R.mean[c]);
      r[a] = v;
      s += v * v;
    }
    double sc = 1.0, c = 0.5 * s;
    if (apply_loss && R.huber > 0.0) c = huber(R.huber, s, &sc);
    if (mode == 0) {
      for (int a = 0; a < k; a++) out[i].r[a] = sc * r[a];
      out[i].sc = sc;
      atomicAdd(&scalars[R.flags ? SC_FIXED : SC_COST], c);
    } else {
      atomicAdd(&scalars[R.flags ? SC_CAND_FIXED : SC_CAND], c);
    }
    return;
  }
  if (R.flags) return;  // constant block: no columns
  const double sc = out[i].sc;
  if (mode == 1) {
    // J = sc * A on columns off..off+k-1
    double* Hd; double* gd; int ld;
    if (R.kind == 0) { const int f = f_of_pose[R.idx]; Hd = S_upper + (size_t)su_ptr[f] * 36; gd = gp + 6 * f; ld = 6; }
    else if (R.kind == 1) { Hd = prior_H_pt + 9 * (size_t)R.idx; gd = prior_g_pt + 3 * (size_t)R.idx; ld = 3; }
    else { Hd = prior_H_obj + 49 * (size_t)R.idx; gd = prior_g_obj + 7 * (size_t)R.idx; ld = 7; }
    for (int a = 0; a < k; a++) {
      double gv = 0.0;
      for (int m = 0; m < k; m++) gv += sc * R.A[m * k + a] * out[i].r[m];
      atomicAdd(&gd[off + a], gv);
      for (int c = 0; c < k; c++) {
        double hv = 0.0;
        for (int m = 0; m < k; m++) hv += R.A[m * k + a] * R.A[m * k + c];
        hv *= sc * sc;
        atomicAdd(&Hd[(off + a) * ld + off + c], hv);
        if (R.kind == 0 && a == c) atomicAdd(&hpp_diag[6 * f_of_pose[R.idx] + off + a], hv);
      }
    }
    (void)bs;
    return;
  }
  // mode 2
  const double* d = (R.kind == 0 ? dpose + 6 * (size_t)f_of_pose[R.idx] : (R.kind == 1 ? dpoint + 3 * (size_t)R.idx : dobj + 7 * (size_t)R.idx));
  double mc = 0.0;
  for (int a = 0; a < k; a++) {
    double m = 0.0;
    for (int c = 0; c < k; c++) m += sc * R.A[a * k + c] * d[off + c];
    mc += m * (out[i].r[a] + 0.5 * m);
  }
  atomicAdd(&scalars[SC_MODEL], mc);
}

// ------------------------------------------------------------------------------------------ relative-pose factors
struct RelOut { double r[6]; double J1[36]; double J2[36]; };
// mode 0 evaluate (+cost), 1 accumulate into the reduced system, 2 model cost change, 3 candidate cost
__global__ void relpose_kernel(const RelRec* __restrict__ rec, int64_t n, int mode, int apply_loss,
                               const double* __restrict__ poses, RelOut* __restrict__ out,
                               double* __restrict__ S_upper, double* __restrict__ gp, double* __restrict__ hpp_diag,
                               const double* __restrict__ dpose, double* __restrict__ scalars) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const RelRec& R = rec[i];
  const bool fixed = R.f1 < 0 && R.f2 < 0;
  if (mode == 0 || mode == 3) {
    double p1[6], p2[6];
    for (int a = 0; a < 6; a++) { p1[a] = poses[6 * (size_t)R.p1 + a]; p2[a] = poses[6 * (size_t)R.p2 + a]; }
    RelOut o;
    relpose_residual_jacobian(p1, p2, R.tm, R.Rm_inv, R.A6, o.r, mode == 0 ? o.J1 : nullptr, mode == 0 ? o.J2 : nullptr);
    double s = 0.0;
    for (int a = 0; a < 6; a++) s += o.r[a] * o.r[a];
    double sc = 1.0, c = 0.5 * s;
    if (apply_loss && R.huber > 0.0) c = huber(R.huber, s, &sc);
    if (mode == 0) {
      for (int a = 0; a < 6; a++) out[i].r[a] = sc * o.r[a];
      for (int a = 0; a < 36; a++) { out[i].J1[a] = sc * o.J1[a]; out[i].J2[a] = sc * o.J2[a]; }
      atomicAdd(&scalars[fixed ? SC_FIXED : SC_COST], c);
    } else {
      atomicAdd(&scalars[fixed ? SC_CAND_FIXED : SC_CAND], c);
    }
    return;
  }
  if (fixed) return;
  const RelOut& o = out[i];
  if (mode == 1) {
    const int fs[2] = {R.f1, R.f2};
    const double* Js[2] = {o.J1, o.J2};
    const int blk[2] = {R.blk11, R.blk22};
    for (int u = 0; u < 2; u++) {
      if (fs[u] < 0) continue;
      double* Sd = S_upper + (size_t)blk[u] * 36;
      for (int a = 0; a < 6; a++) {
        double gv = 0.0;
        for (int k = 0; k < 6; k++) gv += Js[u][k * 6 + a] * o.r[k];
        atomicAdd(&gp[6 * fs[u] + a], gv);
        for (int b = 0; b < 6; b++) {
          double hv = 0.0;
          for (int k = 0; k < 6; k++) hv += Js[u][k * 6 + a] * Js[u][k * 6 + b];
          atomicAdd(&Sd[a * 6 + b], hv);
          if (a == b) atomicAdd(&hpp_diag[6 * fs[u] + a], hv);
        }
      }
    }
    if (R.blk12 >= 0) {
      // upper block (min f, max f): rows from the smaller-f pose
      const double* Ja = R.swap12 ? o.J2 : o.J1;
      const double* Jb = R.swap12 ? o.J1 : o.J2;
      double* Sd = S_upper + (size_t)R.blk12 * 36;
      for (int a = 0; a < 6; a++)
        for (int b = 0; b < 6; b++) {
          double hv = 0.0;
          for (int k = 0; k < 6; k++) hv += Ja[k * 6 + a] * Jb[k * 6 + b];
          atomicAdd(&Sd[a * 6 + b], hv);
        }
    }
    return;
  }
  double mc = 0.0;
  for (int k = 0; k < 6; k++) {
    double m = 0.0;
    if (R.f1 >= 0) for (int a = 0; a < 6; a++) m += o.J1[k * 6 + a] * dpose[6 * R.f1 + a];
    if (R.f2 >= 0) for (int a = 0; a < 6; a++) m += o.J2[k * 6 + a] * dpose[6 * R.f2 + a];
    mc += m * (o.r[k] + 0.5 * m);
  }
  atomicAdd(&scalars[SC_MODEL], mc);
}

// ------------------------------------------------------------------------------------------ reduced system: finish
// Jacobi scaling of the pose columns from the first Jacobian (Ceres: 1 / (1 + column norm)).
__global__ void pose_scale_kernel(const double* __restrict__ hpp_diag, int n, double* __restrict__ pscale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) pscale[i] = 1.0 / (1.0 + sqrt(hpp_diag[i]));
}

// One warp per block row: S~ = S_p S_raw S_p + D_p^2 mirrored into the full BSR (scalar rows contiguous),
// b~ = S_p (g_p + b_schur), block-Jacobi preconditioner = inverse of the diagonal block, |g_p|_inf.
__global__ void finish_kernel(int nf, const uint32_t* __restrict__ sf_ptr, const uint32_t* __restrict__ sf_col,
                              const uint32_t* __restrict__ sf_src, const double* __restrict__ S_upper,
                              const double* __restrict__ pscale, const double* __restrict__ hpp_diag,
                              const double* __restrict__ gp, const double* __restrict__ b_schur, LMParams lm,
                              double* __restrict__ Sf, double* __restrict__ rhs, double* __restrict__ scalars) {
  // one CTA per block row (a warp per row left most of the GPU idle: 2000 rows of ~1700 elements each)
  const int i = blockIdx.x;
  const int lane = threadIdx.x;
  if (i >= nf) return;
  const uint32_t p0 = sf_ptr[i], nb = sf_ptr[i + 1] - p0;
  double* row = Sf + (size_t)p0 * 36;
  const uint32_t total = nb * 36;
  for (uint32_t t = threadIdx.x; t < total; t += blockDim.x) {
    // destination layout: [a][k][c] -> a * (nb*6) + k*6 + c
    const uint32_t a = t / (nb * 6), rem = t - a * nb * 6, k = rem / 6, c = rem - 6 * k;
    const uint32_t src = sf_src[p0 + k];
    const uint32_t j = sf_col[p0 + k];
    const double* sb = S_upper + (size_t)(src & 0x7fffffffu) * 36;
    double v = (src >> 31) ? sb[c * 6 + a] : sb[a * 6 + c];
    v *= pscale[6 * i + a] * pscale[6 * j + c];
    if (j == (uint32_t)i && a == c) {
      const double sd = pscale[6 * i + a] * pscale[6 * i + a] * hpp_diag[6 * i + a];
      v += fmin(fmax(sd, lm.min_diag), lm.max_diag) / lm.radius;
    }
    row[t] = v;
  }
  if (lane < 32) {   // first warp: rhs + one gradient-max atomic per row
    double ga = 0.0;
    if (lane < 6) {
      const double g = gp[6 * i + lane];
      rhs[6 * i + lane] = pscale[6 * i + lane] * (g + b_schur[6 * i + lane]);
      ga = fabs(g);
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) ga = fmax(ga, __shfl_xor_sync(0xffffffffu, ga, o));
    if (lane == 0 && ga > 0.0) atomic_max_nonneg(&scalars[SC_GMAX], ga);
  }
}

// Inverse of the diagonal 6x6 blocks of the finished matrix (block-Jacobi preconditioner): only needed when the
// block-tridiagonal factorisation is switched off or hit a non-positive pivot, so it is not part of finish_kernel.
__global__ void minv_kernel(int nf, const uint32_t* __restrict__ sf_ptr, const uint32_t* __restrict__ sf_col,
                            const double* __restrict__ Sf, double* __restrict__ Minv, double* __restrict__ scalars) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nf) return;
  const uint32_t p0 = sf_ptr[i], nb = sf_ptr[i + 1] - p0;
  const double* row = Sf + (size_t)p0 * 36;
  uint32_t kd = 0;
  for (uint32_t k = 0; k < nb; k++) if (sf_col[p0 + k] == (uint32_t)i) { kd = k; break; }
  double D[36], inv[36];
  for (int a = 0; a < 6; a++)
    for (int c = 0; c < 6; c++) D[a * 6 + c] = row[a * nb * 6 + kd * 6 + c];
  if (!spd_inverse<6>(D, inv)) {
    atomicAdd(&scalars[SC_FAIL], 1.0);
    for (int a = 0; a < 36; a++) inv[a] = (a % 7 == 0) ? 1.0 / D[a] : 0.0;
  }
  for (int a = 0; a < 36; a++) Minv[(size_t)i * 36 + a] = inv[a];
}

// ------------------------------------------------------------------------------------------ reduced system: PCG
// Persistent cooperative kernel: the whole preconditioned conjugate-gradient solve of S~ y = b~ runs in ONE launch
// (three grid barriers per iteration, no host round trips).  One warp owns a block row: it multiplies the row's
// 6 scalar rows (contiguous, coalesced) against p gathered from L2 and then updates its 6 entries of x, r, z, p.
constexpr int kPcgThreads = 512;
__global__ void __launch_bounds__(kPcgThreads) pcg_kernel(int nf, const uint32_t* __restrict__ sf_ptr,
                                                          const uint32_t* __restrict__ sf_col,
                                                          const double* __restrict__ Sf, const double* __restrict__ rhs,
                                                          const double* __restrict__ Minv, double* __restrict__ y,
                                                          double* r, double* z, double* p, double* q, double* acc /*4x4*/,
                                                          int max_iter, double tol, double* __restrict__ scalars) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double red[3][kPcgThreads / 32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int gw = (blockIdx.x * kPcgThreads + threadIdx.x) >> 5;
  const int GW = (gridDim.x * kPcgThreads) >> 5;
  volatile double* vacc = acc;

  auto block_acc = [&](double a0, double a1, double a2, double* dst) {
    if (lane == 0) { red[0][wib] = a0; red[1][wib] = a1; red[2][wib] = a2; }
    __syncthreads();
    if (threadIdx.x < 3) {
      double s = 0.0;
      for (int i = 0; i < kPcgThreads / 32; i++) s += red[threadIdx.x][i];
      if (s != 0.0) atomicAdd(&dst[threadIdx.x], s);
    }
    __syncthreads();
  };

  // init: y = 0, r = b, z = Minv r, p = z; acc[0] = {rz, bb}
  if (blockIdx.x == 0 && threadIdx.x < 16) acc[threadIdx.x] = 0.0;
  grid.sync();
  {
    double rz = 0.0, bb = 0.0;
    for (int i = gw; i < nf; i += GW) {
      const double rv = lane < 6 ? rhs[6 * i + lane] : 0.0;
      double zv = 0.0;
      for (int c = 0; c < 6; c++) {
        const double rc = __shfl_sync(0xffffffffu, rv, c);
        if (lane < 6) zv += Minv[(size_t)i * 36 + lane * 6 + c] * rc;
      }
      if (lane < 6) { y[6 * i + lane] = 0.0; r[6 * i + lane] = rv; z[6 * i + lane] = zv; p[6 * i + lane] = zv; rz += rv * zv; bb += rv * rv; }
    }
    rz = warp_sum(rz); bb = warp_sum(bb);
    block_acc(rz, bb, 0.0, acc);
  }
  grid.sync();
  double rho = vacc[0];
  const double bb = vacc[1];
  int it = 0;
  double rr = bb;
  int brk = 0;
  if (bb > 0.0) {
    for (it = 0; it < max_iter;) {
      double* A = acc + 4 * ((it + 1) & 3);      // this iteration's accumulators
      double* Z = acc + 4 * ((it + 3) & 3);      // cleared for iteration it+2
      if (blockIdx.x == 0 && threadIdx.x < 4) Z[threadIdx.x] = 0.0;
      // q = S p, pq
      double pq = 0.0;
      for (int i = gw; i < nf; i += GW) {
        const uint32_t p0 = sf_ptr[i], nb = sf_ptr[i + 1] - p0;
        const uint32_t len = nb * 6;
        const double* row = Sf + (size_t)p0 * 36;
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0;
        for (uint32_t e = lane; e < len; e += 32) {
          const uint32_t k = e / 6, c = e - 6 * k;
          const double xv = p[6 * sf_col[p0 + k] + c];
          a0 += row[e] * xv; a1 += row[len + e] * xv; a2 += row[2 * len + e] * xv;
          a3 += row[3 * len + e] * xv; a4 += row[4 * len + e] * xv; a5 += row[5 * len + e] * xv;
        }
        a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3); a4 = warp_sum(a4); a5 = warp_sum(a5);
        if (lane < 6) {
          const double qv = lane == 0 ? a0 : lane == 1 ? a1 : lane == 2 ? a2 : lane == 3 ? a3 : lane == 4 ? a4 : a5;
          q[6 * i + lane] = qv;
          pq += qv * p[6 * i + lane];
        }
      }
      pq = warp_sum(pq);
      block_acc(pq, 0.0, 0.0, A);
      grid.sync();
      const double pqs = ((volatile double*)A)[0];
      if (!(pqs > 0.0)) { brk = 1; break; }
      const double alpha = rho / pqs;
      double rz = 0.0, r2 = 0.0;
      for (int i = gw; i < nf; i += GW) {
        double rv = 0.0;
        if (lane < 6) {
          y[6 * i + lane] += alpha * p[6 * i + lane];
          rv = r[6 * i + lane] - alpha * q[6 * i + lane];
          r[6 * i + lane] = rv;
        }
        double zv = 0.0;
        for (int c = 0; c < 6; c++) {
          const double rc = __shfl_sync(0xffffffffu, rv, c);
          if (lane < 6) zv += Minv[(size_t)i * 36 + lane * 6 + c] * rc;
        }
        if (lane < 6) { z[6 * i + lane] = zv; rz += rv * zv; r2 += rv * rv; }
      }
      rz = warp_sum(rz); r2 = warp_sum(r2);
      block_acc(0.0, rz, r2, A);
      grid.sync();
      const double rho_new = ((volatile double*)A)[1];
      rr = ((volatile double*)A)[2];
      it++;
      if (rr <= tol * tol * bb) break;
      const double beta = rho_new / rho;
      rho = rho_new;
      for (int i = gw; i < nf; i += GW)
        if (lane < 6) p[6 * i + lane] = z[6 * i + lane] + beta * p[6 * i + lane];
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

// ------------------------------------------------------------------------------------------ step / norms
// delta_p = -s_p * y (undo the Jacobi scaling), candidate pose, |delta|^2
__global__ void pose_step_kernel(int nf, const int32_t* __restrict__ pose_of_f, const double* __restrict__ pscale,
                                 const double* __restrict__ y, const double* __restrict__ poses,
                                 double* __restrict__ poses_cand, double* __restrict__ dpose, int contribute,
                                 double* __restrict__ scalars) {
  __shared__ double red[33];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double s2 = 0.0;
  if (i < 6 * nf) {
    const int f = i / 6, a = i - 6 * f;
    const double d = -pscale[i] * y[i];
    dpose[i] = d;
    const size_t k = (size_t)pose_of_f[f];
    poses_cand[6 * k + a] = poses[6 * k + a] + d;
    s2 = d * d;
  }
  s2 = block_sum_all<256>(s2, red);
  if (threadIdx.x == 0 && s2 != 0.0 && contribute) atomicAdd(&scalars[SC_STEP2], s2);
}

// Per-block squared norm of the residual stored in a Jacobian chunk, written at the block's rank in order of
// addition (two-phase outlier rejection, offline_problem_runner.h:689-749: sum over the block's residual entries,
// accumulated left to right without fused multiply-add so that the keys equal the host arithmetic bit for bit).
__global__ void block_sqnorm_kernel(const double* __restrict__ chunks, int64_t n, int chunk, int roff, int k,
                                    const uint32_t* __restrict__ rank, double* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* r = chunks + (size_t)i * chunk + roff;
  double e = 0.0;
  for (int a = 0; a < k; a++) e = __dadd_rn(e, __dmul_rn(r[a], r[a]));
  // blocks removed in place have no rank among the live blocks: they are parked behind the live ones with key -1
  const uint32_t rk = rank[i];
  if (rk & 0x80000000u) { keys[rk & 0x7fffffffu] = -1.0; vals[rk & 0x7fffffffu] = 0xffffffffu; }
  else { keys[rk] = e; vals[rk] = rk; }
}
// set a flag bit in 32-bit words addressed as base[idx[i] * stride_words + word]
__global__ void or_flag_kernel(uint32_t* __restrict__ base, const uint32_t* __restrict__ idx, int64_t n, int stride_words, int word, uint32_t bit) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) base[(size_t)idx[i] * stride_words + word] |= bit;
}
__global__ void set_bytes_kernel(uint8_t* __restrict__ base, const uint32_t* __restrict__ idx, int64_t n, uint8_t v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) base[idx[i]] = v;
}
// flag the last element of every run of equal keys in the (stably) sorted sequence
__global__ void run_end_flags_kernel(const double* __restrict__ keys, int64_t n, uint8_t* __restrict__ flags) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = (i == n - 1) || (keys[i] != keys[i + 1]);
}

// ------------------------------------------------------------------------------------------ marginal covariances of objects
// (J^T J)^-1 restricted to ellipsoid blocks, from the same elimination as the solve (no damping):
//   Sigma_ab = [a == b] H_a^-1 + Z_a^T S^-1 Z_b,   Z_o = E_po H_o^-1  (one 6x7 block per pose slot of the object),
// where S is the reduced camera matrix (reference: ceres::Covariance on the LTM-extraction problem,
// src/refactoring/long_term_map/long_term_object_map_extraction.cpp:362-440).
// One thread per (object, pose slot): W = sum Jp^T Jo over the slot's bbox observations, Z = W Hinv.
__global__ void obj_z_kernel(int64_t n_slots, const uint32_t* __restrict__ slot_obj, const uint32_t* __restrict__ slot_d0,
                             const uint32_t* __restrict__ slot_cnt, const uint32_t* __restrict__ pos, const double* __restrict__ Jb,
                             const double* __restrict__ einv, double* __restrict__ Zo) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_slots) return;
  double W[42];
  for (int a = 0; a < 42; a++) W[a] = 0.0;
  for (uint32_t k = 0; k < slot_cnt[g]; k++) {
    const double* ch = Jb + (size_t)pos[slot_d0[g] + k] * 56;   // [Jp 4x6 | Jo 4x7 | r 4]
    for (int q = 0; q < 4; q++)
      for (int a = 0; a < 6; a++)
        for (int c = 0; c < 7; c++) W[a * 7 + c] += ch[q * 6 + a] * ch[24 + q * 7 + c];
  }
  const double* hinv = einv + (size_t)slot_obj[g] * 49;
  for (int a = 0; a < 6; a++)
    for (int c = 0; c < 7; c++) {
      double z = 0.0;
      for (int d = 0; d < 7; d++) z += W[a * 7 + d] * hinv[d * 7 + c];
      Zo[(size_t)g * 42 + a * 7 + c] = z;
    }
}
// rhs (scaled system) = P Z_o[:, col] scattered to the object's pose rows; rhs must be zero before
__global__ void cov_rhs_kernel(uint32_t s0, uint32_t ns, int col, const int32_t* __restrict__ slot_f, const double* __restrict__ Zo,
                               const double* __restrict__ pscale, double* __restrict__ rhs) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ns * 6) return;
  const uint32_t s = t / 6, a = t - 6 * s;
  const int f = slot_f[s0 + s];
  rhs[6 * f + a] = pscale[6 * f + a] * Zo[(size_t)(s0 + s) * 42 + a * 7 + col];
}
// out[pair][r][c] = [a == b] Hinv_a[r][c] + sum_slots sum_x Z_a[slot][x][r] pscale[6 f + x] Y[c][6 f + x]
struct CovPair { uint32_t s0, ns, obj_a, same; uint64_t out; };
__global__ void cov_dot_kernel(const CovPair* __restrict__ pairs, int n_pairs, const int32_t* __restrict__ slot_f,
                               const double* __restrict__ Zo, const double* __restrict__ pscale, const double* __restrict__ Y,
                               size_t ystride, const double* __restrict__ einv, double* __restrict__ out) {
  const int pi = blockIdx.x;
  if (pi >= n_pairs) return;
  const CovPair P = pairs[pi];
  const int r = threadIdx.x / 7, c = threadIdx.x - 7 * r;
  if (r >= 7) return;
  double acc = P.same ? einv[(size_t)P.obj_a * 49 + r * 7 + c] : 0.0;
  for (uint32_t s = 0; s < P.ns; s++) {
    const int f = slot_f[P.s0 + s];
    for (int x = 0; x < 6; x++) acc += Zo[(size_t)(P.s0 + s) * 42 + x * 7 + r] * pscale[6 * f + x] * Y[(size_t)c * ystride + 6 * f + x];
  }
  out[P.out + r * 7 + c] = acc;
}

// |x|^2 over the variable blocks: `mask` (per block, nonzero = skip), block size bs
__global__ void xnorm_kernel(const double* __restrict__ x, const uint8_t* __restrict__ skip, int64_t nblocks, int bs,
                             double* __restrict__ scalars) {
  __shared__ double red[33];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double s = 0.0;
  if (i < nblocks * bs && !skip[i / bs]) s = x[i] * x[i];
  s = block_sum_all<256>(s, red);
  if (threadIdx.x == 0 && s != 0.0) atomicAdd(&scalars[SC_XNORM2], s);
}

}  // namespace obvi
