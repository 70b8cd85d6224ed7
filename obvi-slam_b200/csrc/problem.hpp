// Host-side problem container: parameter blocks keyed by host pointer, factor records, and the
// structure build that turns them into the flat, sorted device layout (DESIGN.md "Data layout").
// Mirrors what ceres::Problem + the reference's residual_creator hold on the host
// (include/refactoring/optimization/residual_creator.h:20-436,
//  include/refactoring/optimization/object_pose_graph_optimizer.h:126-632).
#pragma once

#include <stdint.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <unordered_map>
#include <vector>

#include "../../include/obvi_ba.h"
#include "host_math.hpp"

namespace obvi {

struct ParamBlock {
  double* host;
  int32_t size;
  uint8_t constant;
  uint8_t alive;
};

struct ParamArray { double* base; int32_t size; int64_t count; int64_t first_id; };

struct Camera { double intr[4]; double Rinv[9]; double tinv[3]; };

struct ReprojFactor { int32_t pose, point, cam; uint8_t alive; double px, py, sigma, huber; };
struct BBoxFactor { int32_t obj, pose, cam; uint8_t alive; double brect[4]; double A4[16]; double invalid_err, huber; };
// unary factor on one block: r = A (x[off:off+k] - mean), k rows (shape 3, ltm 7, param prior 1)
struct UnaryFactor { int32_t block; int32_t type; int32_t k; int32_t off; uint8_t alive; double A[49]; double mean[7]; double huber; };
struct RelPoseFactor { int32_t p1, p2; uint8_t alive; double tm[3]; double Rm_inv[9]; double A6[36]; double huber; };

inline obvi_factor_id make_id(int type, uint64_t index) { return ((uint64_t)type << 56) | index; }
inline int id_type(obvi_factor_id id) { return (int)(id >> 56); }
inline uint64_t id_index(obvi_factor_id id) { return id & ((1ull << 56) - 1); }

// ---- device-facing records -------------------------------------------------------------------------
struct ObsRec {       // 32 B, one per reprojection observation, pose-major order
  double ur, vr;      // rectified feature
  uint32_t pose;      // pose index
  uint32_t point;     // internal point index
  uint32_t dst;       // position of this observation's Jacobian chunk: POINT-major (point, pose, camera) order
  uint32_t flags;     // bit0: pose constant, bit1: point constant, bit2: masked, bits 8-15: camera index (valid when C <= 256),
                      // bits 16-31: calibration class (camera, sigma, huber)
};
inline uint32_t obs_cls(const ObsRec& o) { return o.flags >> 16; }
struct CalibClass { double mx, my, huber; int32_t cam; int32_t pad; };  // 32 B
struct BBoxRec {      // one per bbox observation, object-major order
  double brect[4];
  double A4[16];
  double invalid_err, huber;
  uint32_t obj, pose, cam, flags;  // bit0: pose constant, bit1: object constant
};
struct UnaryRec { double A[49]; double mean[7]; double huber; int32_t kind /*0 pose 1 point 2 obj*/, idx, k, off; uint32_t flags, pad; };
struct RelRec { double tm[3]; double Rm_inv[9]; double A6[36]; double huber; int32_t p1, p2, f1, f2; int32_t blk11, blk12, blk22, swap12; };

constexpr int kRowSpan = 20;        // widest pose span (in f indices) of a point handled by the row-owner kernels
constexpr int kRowItemEnts = 128;   // entries per warp work item

// Flat structure produced by the build (host copies; the solver uploads them).
struct Structure {
  int K = 0, P = 0, O = 0, C = 0;   // poses, points, objects (referenced by live factors), cameras
  int nf = 0;                        // variable poses (f-blocks)
  int64_t n_obs = 0, n_bbox = 0, n_unary = 0, n_rel = 0;
  int64_t num_residual_blocks_reduced = 0, num_residuals_reduced = 0;
  int num_param_blocks_reduced = 0, num_params_reduced = 0;
  std::vector<int32_t> pose_block, point_block, obj_block;  // internal index -> param block id
  std::vector<int32_t> f_of_pose;                           // pose index -> f index or -1
  std::vector<uint8_t> point_const, obj_const;
  // reprojection observations, pose-major
  std::vector<ObsRec> obs;
  std::vector<uint32_t> obs_user;          // internal obs -> index in Problem::reproj
  std::vector<uint32_t> pose_ptr;          // K+1: obs segment of each pose
  std::vector<CalibClass> classes;
  // e-block lists (points then objects share the same layout)
  struct EList {
    std::vector<uint32_t> ptr;       // ne+1
    std::vector<uint32_t> pos;       // obs position (index into the J array) per list entry
    std::vector<int32_t> f;          // f index of the entry's pose (-1 constant)
    std::vector<uint16_t> slot;      // merged pose slot within the e-block (0xFFFF: constant pose)
    std::vector<uint32_t> pair_ptr;  // ne+1: offset into pair_blk (ns (ns+1)/2 entries per e-block)
    std::vector<uint16_t> nslots;    // ne
    std::vector<uint32_t> pair_blk;  // S_upper block index of slot pair (a<=b), row-major upper
    std::vector<int32_t> slot_f;     // per e-block slot -> f index (offset by slot_ptr)
    std::vector<uint32_t> slot_ptr;  // ne+1
    int max_slots = 0;
  } pts, objs;
  // Row-owner point elimination (point_prep_kernel + schur_rows_kernel).  A point is "regular" when its variable poses
  // span fewer than kRowSpan consecutive f indices; every other point goes to the generic per-e-block kernel.
  //   groups : per point, one record per run of observations taken from the same pose (stereo pair = one group)
  //   entries: per (PAIR of reduced-matrix rows (2m, 2m + 1), range of five column offsets): the points with a slot in
  //            either row (layout at row_pair_entries in build_structure)
  //   items  : <= kRowItemEnts consecutive entries of one (row pair, range) list = the work of one warp
  struct RowGroup { uint32_t pos0, pos1, gs, cnt; };   // chunks pos0 .. pos0 + cnt - 1 of the point-major Jacobian array; pos1 = f index of the pose (int32, -1 constant)
  struct RowItem { uint32_t row, dlo, off, cnt; };     // row = pair index m
  struct PointRows {
    std::vector<uint32_t> grp_ptr;      // P + 1
    std::vector<RowGroup> grp;
    std::vector<int32_t> grp_f;         // f index of the group's pose (-1 constant)
    std::vector<uint8_t> regular;       // P
    std::vector<uint32_t> ent;          // (slot of row 2m) + 1 | records << 26 | row flags << 29
    std::vector<RowItem> items;
    std::vector<uint32_t> rowblk;       // 2 ceil(nf / 2) * kRowSpan: S_upper block of (a, a + d) or 0xFFFFFFFF
    std::vector<uint32_t> fallback;
    int64_t n_slots = 0;
  } prow;
  std::vector<BBoxRec> bbox;
  std::vector<uint32_t> bbox_user;
  std::vector<UnaryRec> unary;
  std::vector<uint32_t> unary_user;        // index into Problem::unary
  std::vector<RelRec> rel;
  std::vector<uint32_t> rel_user;
  // reduced camera system
  int64_t n_upper = 0;                      // blocks in the upper triangle (incl. diagonal)
  std::vector<uint32_t> su_ptr, su_col;     // upper BSR (row i: cols >= i)
  std::vector<uint32_t> sf_ptr, sf_col, sf_src;  // full symmetric BSR; sf_src = upper block index | (transposed << 31)
};

struct Problem {
  int device = 0;
  std::string error;
  std::vector<ParamBlock> blocks;
  std::vector<ParamArray> arrays;            // sorted by base
  std::unordered_map<const double*, int32_t> single;
  std::vector<Camera> cams;
  std::vector<ReprojFactor> reproj;
  std::vector<BBoxFactor> bbox;
  std::vector<UnaryFactor> unary;
  std::vector<RelPoseFactor> rel;
  std::vector<obvi_factor_id> order;         // addition order (dead ids skipped lazily)
  int64_t n_live = 0;
  bool dirty = true;                         // structure must be rebuilt
  uint64_t const_epoch = 0;

  int32_t find_block(const double* ptr) const {
    if (!arrays.empty()) {
      auto it = std::upper_bound(arrays.begin(), arrays.end(), ptr, [](const double* p, const ParamArray& a) { return p < a.base; });
      if (it != arrays.begin()) {
        const ParamArray& a = *(it - 1);
        const ptrdiff_t d = ptr - a.base;
        if (d >= 0 && d < a.count * a.size && d % a.size == 0) return (int32_t)(a.first_id + d / a.size);
      }
    }
    auto it = single.find(ptr);
    return it == single.end() ? -1 : it->second;
  }
  int32_t add_block(double* ptr, int size) {
    int32_t id = find_block(ptr);
    if (id >= 0) {
      if (!blocks[id].alive) {
        // a removed block: the address may have been recycled for a block of another size (Ceres accepts that sequence)
        if (blocks[id].size == size) { blocks[id].alive = 1; blocks[id].constant = 0; dirty = true; return id; }
        if (single.count(ptr)) {   // replace the dead entry; a dead element of a registered ARRAY keeps its slot (and its size)
          single.erase(ptr);
          id = -1;
        } else {
          return -2;
        }
      } else {
        if (blocks[id].size != size) return -2;
        return id;
      }
    }
    id = (int32_t)blocks.size();
    blocks.push_back({ptr, size, 0, 1});
    single.emplace(ptr, id);
    dirty = true;
    return id;
  }
  // true when a singly registered block (alive or not) starts inside [base, base + count * size)
  bool array_overlaps_single(const double* base, int size, int64_t count) const {
    const double* end = base + count * size;
    if ((int64_t)single.size() < count) { for (const auto& kv : single) if (kv.first >= base && kv.first < end) return true; return false; }
    for (int64_t i = 0; i < count * size; i++) if (single.count(base + i)) return true;
    return false;
  }
  bool array_overlaps_array(const double* base, int size, int64_t count) const {
    const double* end = base + count * size;
    for (const ParamArray& a : arrays) if (base < a.base + a.count * a.size && a.base < end) return true;
    return false;
  }
  int add_array(double* base, int size, int64_t count) {
    ParamArray a{base, size, count, (int64_t)blocks.size()};
    for (int64_t i = 0; i < count; i++) blocks.push_back({base + i * size, size, 0, 1});
    arrays.insert(std::upper_bound(arrays.begin(), arrays.end(), a, [](const ParamArray& x, const ParamArray& y) { return x.base < y.base; }), a);
    dirty = true;
    return 0;
  }
};

// reduced-program bookkeeping (Summary::num_*_reduced): over the whole graph, not just this rank
inline void count_reduced(const Problem& pb, Structure& S) {
  const int nb = (int)pb.blocks.size();
  S.num_residual_blocks_reduced = 0; S.num_residuals_reduced = 0; S.num_param_blocks_reduced = 0; S.num_params_reduced = 0;
  std::vector<uint8_t> touched(nb, 0);
  auto var = [&](int b) { return !pb.blocks[b].constant; };
  for (const auto& f : pb.reproj) if (f.alive && (var(f.pose) || var(f.point))) { S.num_residual_blocks_reduced++; S.num_residuals_reduced += 2; touched[f.pose] = touched[f.point] = 1; }
  for (const auto& f : pb.bbox) if (f.alive && (var(f.pose) || var(f.obj))) { S.num_residual_blocks_reduced++; S.num_residuals_reduced += 4; touched[f.pose] = touched[f.obj] = 1; }
  for (const auto& f : pb.unary) if (f.alive && var(f.block)) { S.num_residual_blocks_reduced++; S.num_residuals_reduced += f.k; touched[f.block] = 1; }
  for (const auto& f : pb.rel) if (f.alive && (var(f.p1) || var(f.p2))) { S.num_residual_blocks_reduced++; S.num_residuals_reduced += 6; touched[f.p1] = touched[f.p2] = 1; }
  for (int b = 0; b < nb; b++) if (touched[b] && var(b)) { S.num_param_blocks_reduced++; S.num_params_reduced += pb.blocks[b].size; }
}

// ---- structure build ---------------------------------------------------------------------------
// rank/world: e-blocks (points, objects) are dealt to ranks in contiguous ranges of the internal
// (first-observing-keyframe) order, balanced by observation count; rank 0 also owns the pose-only factors.
#define OBVI_TMARK(name) do { if (tm_on) { auto t1 = std::chrono::steady_clock::now(); fprintf(stderr, "[build] before %s: +%.1f ms\n", name, std::chrono::duration<double, std::milli>(t1 - tm_t0).count()); tm_t0 = t1; } } while (0)
inline bool build_structure(const Problem& pb, Structure& S, int rank, int world, std::string& err) {
  const bool tm_on = getenv("OBVI_BUILD_TIMING") != nullptr;
  auto tm_t0 = std::chrono::steady_clock::now();
  S = Structure();
  const int nb = (int)pb.blocks.size();
  std::vector<int32_t> pose_of_block(nb, -1), point_of_block(nb, -1), obj_of_block(nb, -1);
  std::vector<uint8_t> used(nb, 0);
  const int64_t n_reproj = (int64_t)pb.reproj.size();
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < n_reproj; n++) { const ReprojFactor& f = pb.reproj[n]; if (f.alive) { used[f.pose] = 1; used[f.point] = 1; } }   // every writer stores 1
  for (const auto& f : pb.bbox) if (f.alive) { used[f.obj] = 1; used[f.pose] = 1; }
  for (const auto& f : pb.unary) if (f.alive) used[f.block] = 1;
  for (const auto& f : pb.rel) if (f.alive) { used[f.p1] = 1; used[f.p2] = 1; }
  for (int b = 0; b < nb; b++) {
    if (!used[b]) continue;
    if (!pb.blocks[b].alive) { err = "a live factor references a removed parameter block"; return false; }
    if (pb.blocks[b].size == 6) { pose_of_block[b] = S.K++; S.pose_block.push_back(b); }
  }
  S.C = (int)pb.cams.size();
  S.f_of_pose.assign(S.K, -1);
  for (int k = 0; k < S.K; k++) if (!pb.blocks[S.pose_block[k]].constant) S.f_of_pose[k] = S.nf++;

  // points: internal order = (first observing pose, block id)
  {
    std::vector<int32_t> first(nb, INT32_MAX);
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < n_reproj; n++) {
      const ReprojFactor& f = pb.reproj[n];
      if (!f.alive) continue;
      const int32_t k = pose_of_block[f.pose];
      int32_t seen = __atomic_load_n(&first[f.point], __ATOMIC_RELAXED);
      while (k < seen && !__atomic_compare_exchange_n(&first[f.point], &seen, k, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    }
    // stable counting sort of the used 3-blocks by first observing pose (bucket K: never observed)
    std::vector<int32_t> ids;
    {
      std::vector<uint32_t> start(S.K + 2, 0);
      auto bucket = [&](int b) { return first[b] == INT32_MAX ? S.K : first[b]; };
      for (int b = 0; b < nb; b++) if (used[b] && pb.blocks[b].size == 3) start[bucket(b) + 1]++;
      for (int k = 0; k <= S.K; k++) start[k + 1] += start[k];
      ids.resize(start[S.K + 1]);
      for (int b = 0; b < nb; b++) if (used[b] && pb.blocks[b].size == 3) ids[start[bucket(b)]++] = b;
    }
    S.P = (int)ids.size();
    S.point_block = ids;
    S.point_const.resize(S.P);
    for (int i = 0; i < S.P; i++) { point_of_block[ids[i]] = i; S.point_const[i] = pb.blocks[ids[i]].constant; }
  }
  for (int b = 0; b < nb; b++) if (used[b] && pb.blocks[b].size == 7) { obj_of_block[b] = S.O++; S.obj_block.push_back(b); }
  S.obj_const.resize(S.O);
  for (int i = 0; i < S.O; i++) S.obj_const[i] = pb.blocks[S.obj_block[i]].constant;

  // ownership ranges (multi-GPU): by cumulative observation count
  std::vector<uint32_t> pt_cnt(S.P, 0), ob_cnt(S.O, 0);
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < n_reproj; n++) { const ReprojFactor& f = pb.reproj[n]; if (f.alive) __atomic_fetch_add(&pt_cnt[point_of_block[f.point]], 1u, __ATOMIC_RELAXED); }
  for (const auto& f : pb.bbox) if (f.alive) ob_cnt[obj_of_block[f.obj]]++;
  auto owner_range = [&](const std::vector<uint32_t>& cnt, int& lo, int& hi) {
    const int n = (int)cnt.size();
    if (world <= 1) { lo = 0; hi = n; return; }
    uint64_t total = 0; for (uint32_t c : cnt) total += c + 1;
    uint64_t acc = 0; lo = n; hi = n; bool lo_set = false;
    for (int i = 0; i < n; i++) {
      const int r = (int)std::min<uint64_t>(world - 1, acc * world / std::max<uint64_t>(total, 1));
      if (r == rank && !lo_set) { lo = i; lo_set = true; }
      if (r > rank) { hi = i; break; }
      acc += cnt[i] + 1;
    }
    if (!lo_set) lo = hi;
  };
  int p_lo, p_hi, o_lo, o_hi;
  owner_range(pt_cnt, p_lo, p_hi);
  owner_range(ob_cnt, o_lo, o_hi);
  const bool owns_pose_factors = (rank == 0);

  // calibration classes
  auto class_of = [&](int cam, double sigma, double huber) {
    const double mx = pb.cams[cam].intr[0] / sigma, my = pb.cams[cam].intr[1] / sigma;
    for (size_t i = 0; i < S.classes.size(); i++)
      if (S.classes[i].cam == cam && S.classes[i].mx == mx && S.classes[i].my == my && S.classes[i].huber == huber) return (uint32_t)i;
    S.classes.push_back({mx, my, huber, cam, 0});
    return (uint32_t)(S.classes.size() - 1);
  };

  OBVI_TMARK("0");
  // ---- reprojection observations: stable counting sort by pose, then (camera, point) inside each pose.  The factors are
  // split into contiguous chunks, one per thread; chunk t's share of a pose segment follows chunk t - 1's, so the scatter is
  // the one a serial pass would produce.  Calibration classes are numbered in order of first appearance (per-chunk lists
  // merged in chunk order give exactly that order).
  {
    const int64_t nfac = (int64_t)pb.reproj.size();
    int nt = 1;
#ifdef _OPENMP
    nt = std::max(1, omp_get_max_threads());
#endif
    nt = (int)std::min<int64_t>(nt, std::max<int64_t>(1, nfac / 4096));
    struct ClsKey { int cam; double sigma, huber; };
    std::vector<std::vector<uint32_t>> hist(nt, std::vector<uint32_t>(S.K, 0));
    std::vector<std::vector<ClsKey>> local_cls(nt);
    auto chunk = [&](int t, int64_t& b, int64_t& e) { b = nfac * t / nt; e = nfac * (t + 1) / nt; };
    auto mine = [&](const ReprojFactor& f, int& pt) {
      if (!f.alive) return false;
      pt = point_of_block[f.point];
      return pt >= p_lo && pt < p_hi;
    };
#pragma omp parallel for num_threads(nt) schedule(static, 1)
    for (int t = 0; t < nt; t++) {
      int64_t b, e; chunk(t, b, e);
      std::vector<uint32_t>& h = hist[t];
      std::vector<ClsKey>& lc = local_cls[t];
      int last_cam = -1; double last_sigma = 0, last_huber = 0;
      for (int64_t n = b; n < e; n++) {
        const ReprojFactor& f = pb.reproj[n];
        int pt;
        if (!mine(f, pt)) continue;
        h[pose_of_block[f.pose]]++;
        if (f.cam != last_cam || f.sigma != last_sigma || f.huber != last_huber) {
          last_cam = f.cam; last_sigma = f.sigma; last_huber = f.huber;
          bool seen = false;
          for (const ClsKey& c : lc) if (c.cam == f.cam && c.sigma == f.sigma && c.huber == f.huber) { seen = true; break; }
          if (!seen) lc.push_back({f.cam, f.sigma, f.huber});
        }
      }
    }
    for (int t = 0; t < nt; t++) for (const ClsKey& c : local_cls[t]) class_of(c.cam, c.sigma, c.huber);
    if (S.classes.size() >= 65536) { err = "more than 65535 (camera, sigma, huber) calibration classes"; return false; }
    // segment starts, and the start of every chunk's share inside each segment
    S.pose_ptr.assign(S.K + 1, 0);
    for (int k = 0; k < S.K; k++) {
      uint32_t acc = S.pose_ptr[k];
      for (int t = 0; t < nt; t++) { const uint32_t c = hist[t][k]; hist[t][k] = acc; acc += c; }
      S.pose_ptr[k + 1] = acc;
    }
    S.n_obs = S.pose_ptr[S.K];
    // scatter (key, factor index); the key orders a pose segment by (camera, point)
    std::vector<uint64_t> key(S.n_obs);
    std::vector<uint32_t> fac(S.n_obs);
#pragma omp parallel for num_threads(nt) schedule(static, 1)
    for (int t = 0; t < nt; t++) {
      int64_t b, e; chunk(t, b, e);
      std::vector<uint32_t>& cur = hist[t];
      for (int64_t n = b; n < e; n++) {
        const ReprojFactor& f = pb.reproj[n];
        int pt;
        if (!mine(f, pt)) continue;
        const uint32_t q = cur[pose_of_block[f.pose]]++;
        key[q] = ((uint64_t)(uint32_t)f.cam << 32) | (uint32_t)pt;
        fac[q] = (uint32_t)n;
      }
    }
    // sort inside each pose segment (ties keep the order of addition), then write the records in their final place
    S.obs.resize(S.n_obs); S.obs_user.resize(S.n_obs);
#pragma omp parallel
    {
      std::vector<std::pair<uint64_t, uint32_t>> kv;
      uint32_t last_cls = 0; int last_cam = -1; double last_sigma = 0, last_huber = 0;
#pragma omp for schedule(dynamic, 1)     // (a 50-keyframe window has 50 segments of ~2000 records: chunks of 16 kept 4 threads busy)
      for (int k = 0; k < S.K; k++) {
        const uint32_t b = S.pose_ptr[k], e = S.pose_ptr[k + 1];
        kv.resize(e - b);
        for (uint32_t i = 0; i < e - b; i++) kv[i] = {key[b + i], i};
        std::sort(kv.begin(), kv.end());
        for (uint32_t i = 0; i < e - b; i++) {
          const uint32_t n = fac[b + kv[i].second];
          const ReprojFactor& f = pb.reproj[n];
          if (f.cam != last_cam || f.sigma != last_sigma || f.huber != last_huber) {
            const double mx = pb.cams[f.cam].intr[0] / f.sigma, my = pb.cams[f.cam].intr[1] / f.sigma;
            for (size_t c = 0; c < S.classes.size(); c++)
              if (S.classes[c].cam == f.cam && S.classes[c].mx == mx && S.classes[c].my == my && S.classes[c].huber == f.huber) { last_cls = (uint32_t)c; break; }
            last_cam = f.cam; last_sigma = f.sigma; last_huber = f.huber;
          }
          const Camera& c = pb.cams[f.cam];
          const int pt = (int)(uint32_t)kv[i].first;
          ObsRec& o = S.obs[b + i];
          o.ur = (f.px - c.intr[2]) / c.intr[0]; o.vr = (f.py - c.intr[3]) / c.intr[1];
          o.pose = k; o.point = pt; o.dst = 0;   // dst: filled once the point lists exist
          o.flags = (S.f_of_pose[k] < 0 ? 1u : 0u) | (S.point_const[pt] ? 2u : 0u) | (((uint32_t)f.cam & 0xffu) << 8) | (last_cls << 16);
          S.obs_user[b + i] = n;
        }
      }
    }
  }
  OBVI_TMARK("1");
  // ---- bbox observations, object-major (object, pose, camera)
  {
    std::vector<uint32_t> ids;
    for (size_t n = 0; n < pb.bbox.size(); n++) {
      const BBoxFactor& f = pb.bbox[n];
      if (!f.alive) continue;
      const int o = obj_of_block[f.obj];
      if (o < o_lo || o >= o_hi) continue;
      ids.push_back((uint32_t)n);
    }
    std::stable_sort(ids.begin(), ids.end(), [&](uint32_t x, uint32_t y) {
      const BBoxFactor &a = pb.bbox[x], &b = pb.bbox[y];
      const int oa = obj_of_block[a.obj], ob = obj_of_block[b.obj];
      if (oa != ob) return oa < ob;
      const int pa = pose_of_block[a.pose], pb2 = pose_of_block[b.pose];
      if (pa != pb2) return pa < pb2;
      return a.cam < b.cam; });
    S.n_bbox = (int64_t)ids.size();
    S.bbox.resize(ids.size()); S.bbox_user = ids;
    for (size_t i = 0; i < ids.size(); i++) {
      const BBoxFactor& f = pb.bbox[ids[i]];
      BBoxRec& r = S.bbox[i];
      std::memcpy(r.brect, f.brect, sizeof(r.brect)); std::memcpy(r.A4, f.A4, sizeof(r.A4));
      r.invalid_err = f.invalid_err; r.huber = f.huber;
      r.obj = obj_of_block[f.obj]; r.pose = pose_of_block[f.pose]; r.cam = f.cam;
      r.flags = (S.f_of_pose[r.pose] < 0 ? 1u : 0u) | (S.obj_const[r.obj] ? 2u : 0u);
    }
  }

  // ---- unary factors (shape / LTM / parameter priors)
  for (size_t n = 0; n < pb.unary.size(); n++) {
    const UnaryFactor& f = pb.unary[n];
    if (!f.alive) continue;
    UnaryRec r; std::memset(&r, 0, sizeof(r));
    std::memcpy(r.A, f.A, sizeof(r.A)); std::memcpy(r.mean, f.mean, sizeof(r.mean));
    r.huber = f.huber; r.k = f.k; r.off = f.off;
    const int sz = pb.blocks[f.block].size;
    if (sz == 6) { r.kind = 0; r.idx = pose_of_block[f.block]; r.flags = S.f_of_pose[r.idx] < 0 ? 1u : 0u; if (!owns_pose_factors) continue; }
    else if (sz == 3) { r.kind = 1; r.idx = point_of_block[f.block]; r.flags = S.point_const[r.idx] ? 1u : 0u; if (r.idx < p_lo || r.idx >= p_hi) continue; }
    else { r.kind = 2; r.idx = obj_of_block[f.block]; r.flags = S.obj_const[r.idx] ? 1u : 0u; if (r.idx < o_lo || r.idx >= o_hi) continue; }
    S.unary.push_back(r); S.unary_user.push_back((uint32_t)n);
  }
  S.n_unary = (int64_t)S.unary.size();

  OBVI_TMARK("2");
  // ---- e-block lists + reduced-system structure (bitmap over f x f)
  const int nf = S.nf;
  const int W = (nf + 63) / 64;
  std::vector<uint64_t> bits((size_t)nf * W, 0);
  auto setbit = [&](int i, int j) { bits[(size_t)i * W + (j >> 6)] |= 1ull << (j & 63); };
  for (int i = 0; i < nf; i++) setbit(i, i);

  // regular(e): the e-block is eliminated by the row-owner kernels, which never read its slot-pair table (12.5 M entries at C3)
  auto build_elist = [&](Structure::EList& L, int ne, auto entry_e, auto entry_pose, int64_t n_entries, auto regular) {
    L.ptr.assign(ne + 1, 0);
    L.pos.resize(n_entries); L.f.resize(n_entries); L.slot.resize(n_entries);
    int nt = 1;
#ifdef _OPENMP
    nt = std::max(1, omp_get_max_threads());
#endif
    nt = (int)std::min<int64_t>(nt, std::max<int64_t>(1, n_entries / 65536));
    if (nt > 1) {
      // stable counting sort by e-block with one histogram per thread (contiguous chunks of the entries, chunk t's share
      // of an e-block's range follows chunk t - 1's: the order a serial pass produces)
      std::vector<std::vector<uint32_t>> hist(nt, std::vector<uint32_t>(ne, 0));
      auto chunk = [&](int t, int64_t& b, int64_t& e) { b = n_entries * t / nt; e = n_entries * (t + 1) / nt; };
#pragma omp parallel for num_threads(nt) schedule(static, 1)
      for (int t = 0; t < nt; t++) { int64_t b, e; chunk(t, b, e); std::vector<uint32_t>& h = hist[t]; for (int64_t q = b; q < e; q++) h[entry_e(q)]++; }
      for (int e = 0; e < ne; e++) {
        uint32_t acc = L.ptr[e];
        for (int t = 0; t < nt; t++) { const uint32_t c = hist[t][e]; hist[t][e] = acc; acc += c; }
        L.ptr[e + 1] = acc;
      }
#pragma omp parallel for num_threads(nt) schedule(static, 1)
      for (int t = 0; t < nt; t++) {
        int64_t b, e; chunk(t, b, e); std::vector<uint32_t>& cur = hist[t];
        for (int64_t q = b; q < e; q++) { const uint32_t d = cur[entry_e(q)]++; L.pos[d] = (uint32_t)q; L.f[d] = S.f_of_pose[entry_pose(q)]; }
      }
    } else {
      for (int64_t q = 0; q < n_entries; q++) L.ptr[entry_e(q) + 1]++;
      for (int e = 0; e < ne; e++) L.ptr[e + 1] += L.ptr[e];
      std::vector<uint32_t> cur(L.ptr.begin(), L.ptr.end() - 1);
      for (int64_t q = 0; q < n_entries; q++) { const uint32_t d = cur[entry_e(q)]++; L.pos[d] = (uint32_t)q; L.f[d] = S.f_of_pose[entry_pose(q)]; }
    }
    L.nslots.assign(ne, 0); L.pair_ptr.assign(ne + 1, 0); L.slot_ptr.assign(ne + 1, 0);
    int max_slots = L.max_slots;
#pragma omp parallel for schedule(static, 1024) reduction(max : max_slots)
    for (int e = 0; e < ne; e++) {
      int ns = 0, lastf = -1;
      for (uint32_t d = L.ptr[e]; d < L.ptr[e + 1]; d++) {
        if (L.f[d] < 0) { L.slot[d] = 0xFFFF; continue; }
        if (L.f[d] != lastf) { ns++; lastf = L.f[d]; }
        L.slot[d] = (uint16_t)(ns - 1);
      }
      L.nslots[e] = (uint16_t)ns;
      max_slots = std::max(max_slots, ns);
    }
    L.max_slots = max_slots;
    for (int e = 0; e < ne; e++) L.slot_ptr[e + 1] = L.slot_ptr[e] + L.nslots[e];
    L.slot_f.resize(L.slot_ptr[ne]);
#pragma omp parallel for schedule(static, 1024)
    for (int e = 0; e < ne; e++) {
      uint32_t w = L.slot_ptr[e]; int lastf = -1;
      for (uint32_t d = L.ptr[e]; d < L.ptr[e + 1]; d++) if (L.f[d] >= 0 && L.f[d] != lastf) { L.slot_f[w++] = L.f[d]; lastf = L.f[d]; }
    }
    for (int e = 0; e < ne; e++) {
      const int ns = L.nslots[e];
      L.pair_ptr[e + 1] = L.pair_ptr[e] + (regular(e) ? 0u : (uint32_t)(ns * (ns + 1) / 2));
    }
  };
  // a point goes to the row-owner kernels when it has observations, is variable and its variable poses span < kRowSpan f indices
  auto point_regular = [&](int e) {
    if (S.pts.ptr[e] == S.pts.ptr[e + 1] || S.point_const[e]) return false;
    const int ns = S.pts.nslots[e];
    const int32_t* sf = &S.pts.slot_f[S.pts.slot_ptr[e]];
    return !(ns > 0 && sf[ns - 1] - sf[0] >= kRowSpan);
  };
  // list entries of one e-block are visited in obs order = (pose, camera): equal poses are adjacent
  build_elist(S.pts, S.P, [&](int64_t q) { return (int)S.obs[q].point; }, [&](int64_t q) { return (int)S.obs[q].pose; }, S.n_obs, point_regular);
  build_elist(S.objs, S.O, [&](int64_t q) { return (int)S.bbox[q].obj; }, [&](int64_t q) { return (int)S.bbox[q].pose; }, S.n_bbox, [](int) { return false; });
  // The reprojection Jacobian chunks are stored POINT-major: entry d of the point lists (ordered by point, then pose, then
  // camera) is chunk d, so every point's chunks are one contiguous range.  pts.pos keeps the pose-major record index of entry d
  // on the host (masking, exports); the record carries its chunk position.
#pragma omp parallel for schedule(static)
  for (int64_t d = 0; d < S.n_obs; d++) S.obs[S.pts.pos[d]].dst = (uint32_t)d;
  for (Structure::EList* L : {&S.pts, &S.objs}) {
    const int ne = (int)L->nslots.size();
    const std::vector<uint8_t>& cst = (L == &S.pts) ? S.point_const : S.obj_const;
#pragma omp parallel for schedule(static, 1024)
    for (int e = 0; e < ne; e++) {
      if (cst[e]) continue;
      const int32_t* sf = &L->slot_f[L->slot_ptr[e]]; const int ns = L->nslots[e];
      for (int a = 0; a < ns; a++) for (int b = a; b < ns; b++) {
        uint64_t* word = &bits[(size_t)sf[a] * W + (sf[b] >> 6)];
        const uint64_t m = 1ull << (sf[b] & 63);
        if (!(__atomic_load_n(word, __ATOMIC_RELAXED) & m)) __atomic_fetch_or(word, m, __ATOMIC_RELAXED);
      }
    }
  }
  OBVI_TMARK("3");
  // relative-pose factors
  for (size_t n = 0; n < pb.rel.size(); n++) {
    const RelPoseFactor& f = pb.rel[n];
    if (!f.alive || !owns_pose_factors) continue;
    RelRec r; std::memset(&r, 0, sizeof(r));
    std::memcpy(r.tm, f.tm, sizeof(r.tm)); std::memcpy(r.Rm_inv, f.Rm_inv, sizeof(r.Rm_inv)); std::memcpy(r.A6, f.A6, sizeof(r.A6));
    r.huber = f.huber; r.p1 = pose_of_block[f.p1]; r.p2 = pose_of_block[f.p2];
    r.f1 = S.f_of_pose[r.p1]; r.f2 = S.f_of_pose[r.p2];
    if (r.f1 >= 0 && r.f2 >= 0) setbit(std::min(r.f1, r.f2), std::max(r.f1, r.f2));
    S.rel.push_back(r); S.rel_user.push_back((uint32_t)n);
  }
  S.n_rel = (int64_t)S.rel.size();
  // In a multi-rank run every rank must hold the SAME S structure (it is all-reduced): the structure is
  // the union over all ranks' e-blocks, so ranks > 1 rebuild the bitmap from the full graph.
  if (world > 1) {
    // pairs induced by points / objects owned by other ranks, and by the pose-only factors
    std::vector<std::vector<int32_t>> fl_pt(S.P), fl_ob(S.O);
    for (const auto& f : pb.reproj) if (f.alive) { const int fi = S.f_of_pose[pose_of_block[f.pose]]; const int pt = point_of_block[f.point]; if (fi >= 0 && !S.point_const[pt]) fl_pt[pt].push_back(fi); }
    for (const auto& f : pb.bbox) if (f.alive) { const int fi = S.f_of_pose[pose_of_block[f.pose]]; const int ob = obj_of_block[f.obj]; if (fi >= 0 && !S.obj_const[ob]) fl_ob[ob].push_back(fi); }
    for (auto* fl : {&fl_pt, &fl_ob}) for (auto& v : *fl) {
      std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end());
      for (size_t a = 0; a < v.size(); a++) for (size_t b = a; b < v.size(); b++) setbit(v[a], v[b]);
    }
    for (const auto& f : pb.rel) if (f.alive) { const int f1 = S.f_of_pose[pose_of_block[f.p1]], f2 = S.f_of_pose[pose_of_block[f.p2]]; if (f1 >= 0 && f2 >= 0) setbit(std::min(f1, f2), std::max(f1, f2)); }
  }
  OBVI_TMARK("4");
  // upper BSR + O(1) rank lookup
  std::vector<uint32_t> wpre((size_t)nf * W + 1, 0);  // blocks before word w of row i (global)
  S.su_ptr.assign(nf + 1, 0);
  {
    uint32_t acc = 0;
    for (int i = 0; i < nf; i++) {
      S.su_ptr[i] = acc;
      for (int w = 0; w < W; w++) { wpre[(size_t)i * W + w] = acc; acc += (uint32_t)__builtin_popcountll(bits[(size_t)i * W + w]); }
    }
    S.su_ptr[nf] = acc; S.n_upper = acc;
  }
  S.su_col.resize(S.n_upper);
  for (int i = 0; i < nf; i++) {
    uint32_t w0 = S.su_ptr[i];
    for (int w = 0; w < W; w++) { uint64_t m = bits[(size_t)i * W + w]; while (m) { S.su_col[w0++] = (uint32_t)(w * 64 + __builtin_ctzll(m)); m &= m - 1; } }
  }
  auto blk_of = [&](int i, int j) -> uint32_t {  // i <= j, bit must be set
    const uint64_t word = bits[(size_t)i * W + (j >> 6)];
    return wpre[(size_t)i * W + (j >> 6)] + (uint32_t)__builtin_popcountll(word & ((1ull << (j & 63)) - 1));
  };
  for (Structure::EList* L : {&S.pts, &S.objs}) {
    const int ne = (int)L->nslots.size();
    L->pair_blk.resize(L->pair_ptr[ne]);
    const std::vector<uint8_t>& cst = (L == &S.pts) ? S.point_const : S.obj_const;
#pragma omp parallel for schedule(static, 1024)
    for (int e = 0; e < ne; e++) {
      if (L->pair_ptr[e] == L->pair_ptr[e + 1]) continue;       // no table: a regular point (or no slots)
      const int32_t* sf = &L->slot_f[L->slot_ptr[e]]; const int ns = L->nslots[e];
      uint32_t w = L->pair_ptr[e];
      for (int a = 0; a < ns; a++) for (int b = a; b < ns; b++) L->pair_blk[w++] = cst[e] ? 0u : blk_of(sf[a], sf[b]);
    }
  }
  OBVI_TMARK("6");
  // ---- row-owner structure.  Slots are DENSE per point: one record for every f index between the point's first and
  // last variable pose (gap poses keep an all-zero record), so the column at offset d of slot gs is simply slot gs + d.
  {
    Structure::PointRows& R = S.prow;
    R.regular.assign(S.P, 0); R.grp_ptr.assign(S.P + 1, 0);
    constexpr int kRanges = (kRowSpan + 4) / 5;
    const int npair = (nf + 1) / 2;
    std::vector<uint32_t> cnt((size_t)npair * kRanges + 1, 0);
    // Entries of one point: for every PAIR of reduced-matrix rows (fa, fb) = (2m, 2m + 1) the point has a slot in, and every
    // range r of five column offsets, one entry = (g + 1) | nl << 26 | va << 29 | vb << 30 with g the (dense) slot of row fa,
    // nl <= 6 the number of the point's records from g + 5r on, va / vb whether the point has the row at all.  Row fa takes
    // the products with records g + 5r + k, k < min(nl, 5), row fb those with k = 1 .. nl - 1: six operand records for up to
    // ten 6x6 products.  Gap poses inside a track own no row.
    auto row_pair_entries = [](const int32_t* sf, int ns, uint32_t d0, auto&& emit) {
      if (ns == 0) return;
      const int f0 = sf[0], f1 = sf[ns - 1];
      uint32_t present = 0;
      for (int a = 0; a < ns; a++) present |= 1u << (sf[a] - f0);
      for (int m = f0 >> 1; 2 * m <= f1; m++) {
        const int fa = 2 * m, fb = fa + 1;
        const bool va = fa >= f0 && ((present >> (fa - f0)) & 1u), vb = fb <= f1 && ((present >> (fb - f0)) & 1u);
        for (int r = 0; r < kRanges; r++) {
          const int nA = va ? std::max(0, std::min(5, f1 - fa - 5 * r + 1)) : 0, nB = vb ? std::max(0, std::min(5, f1 - fb - 5 * r + 1)) : 0;
          if (!nA && !nB) break;
          const int nl = std::max(nA, nB ? nB + 1 : 0);      // records g + 5r + k, k < nl, are this point's
          emit(m, r, (uint32_t)((int64_t)d0 + (fa - f0) + 1) | ((uint32_t)nl << 26) | ((uint32_t)(nA > 0) << 29) | ((uint32_t)(nB > 0) << 30));
        }
      }
    };
    std::vector<uint32_t> dptr(S.P + 1, 0);
    std::vector<uint8_t> pt_has_prior(S.P, 0);
    for (const UnaryRec& u : S.unary) if (u.kind == 1) pt_has_prior[u.idx] = 1;
    // pass A (parallel): classify every point, count its dense slots and its groups
    std::vector<uint8_t> kind(S.P, 0);          // 0: nothing to do, 1: generic kernels (fallback), 2: row-owner kernels
    std::vector<uint32_t> span_of(S.P, 0), ngs_of(S.P, 0);
#pragma omp parallel for schedule(static, 1024)
    for (int e = 0; e < S.P; e++) {
      // constant points with observations (they still carry model-cost terms) and prior-only points keep the generic kernels
      if (S.pts.ptr[e] == S.pts.ptr[e + 1]) { if (!S.point_const[e] && pt_has_prior[e]) kind[e] = 1; continue; }
      if (!point_regular(e)) { kind[e] = 1; continue; }     // constant, or a track over >= kRowSpan f indices
      const int32_t* sf = &S.pts.slot_f[S.pts.slot_ptr[e]]; const int ns = S.pts.nslots[e];
      kind[e] = 2;
      span_of[e] = ns ? (uint32_t)(sf[ns - 1] - sf[0] + 1) : 0u;
      uint32_t ngs = 0;
      for (uint32_t d = S.pts.ptr[e]; d < S.pts.ptr[e + 1];) {
        uint32_t d2 = d + 1;
        while (d2 < S.pts.ptr[e + 1] && S.pts.slot[d] != 0xFFFF && S.pts.slot[d2] == S.pts.slot[d]) d2++;
        ngs++; d = d2;
      }
      ngs_of[e] = ngs;
    }
    for (int e = 0; e < S.P; e++) {
      R.grp_ptr[e + 1] = R.grp_ptr[e] + ngs_of[e]; dptr[e + 1] = dptr[e] + span_of[e];
      if (kind[e] == 1) R.fallback.push_back((uint32_t)e);
      R.regular[e] = kind[e] == 2;
    }
    R.grp.resize(R.grp_ptr[S.P]); R.grp_f.resize(R.grp_ptr[S.P]);
    // pass B (parallel): one record per run of entries from the same pose (constant poses: one group per entry, no slot)
#pragma omp parallel for schedule(static, 1024)
    for (int e = 0; e < S.P; e++) {
      if (kind[e] != 2) continue;
      const int32_t* sf = &S.pts.slot_f[S.pts.slot_ptr[e]];
      uint32_t w = R.grp_ptr[e];
      for (uint32_t d = S.pts.ptr[e]; d < S.pts.ptr[e + 1];) {
        uint32_t d2 = d + 1;
        while (d2 < S.pts.ptr[e + 1] && S.pts.slot[d] != 0xFFFF && S.pts.slot[d2] == S.pts.slot[d]) d2++;
        Structure::RowGroup G;
        G.cnt = d2 - d; G.pos0 = d; G.pos1 = (uint32_t)S.pts.f[d];   // chunks d .. d + cnt - 1 (point-major positions); pos1 = f index of the pose
        G.gs = S.pts.slot[d] == 0xFFFF ? 0xFFFFFFFFu : dptr[e] + (uint32_t)(S.pts.f[d] - sf[0]);
        R.grp[w] = G; R.grp_f[w] = S.pts.f[d]; w++;
        d = d2;
      }
    }
    R.n_slots = dptr[S.P];
    if (R.n_slots >= (1ll << 26) - 1) { err = "too many point slots for the row-owner elimination"; return false; }
    // entries: counting sort by (row pair, range), points in order inside a list; one histogram per thread over a
    // contiguous chunk of the points (chunk t's share of a list follows chunk t - 1's: the serial order)
    {
      int nt = 1;
#ifdef _OPENMP
      nt = std::max(1, omp_get_max_threads());
#endif
      nt = std::min(nt, std::max(1, S.P / 4096));
      const size_t nl = (size_t)npair * kRanges;
      std::vector<std::vector<uint32_t>> hist(nt, std::vector<uint32_t>(nl, 0));
      auto chunk = [&](int t, int& b, int& e) { b = (int)((int64_t)S.P * t / nt); e = (int)((int64_t)S.P * (t + 1) / nt); };
#pragma omp parallel for num_threads(nt) schedule(static, 1)
      for (int t = 0; t < nt; t++) {
        int b, e2; chunk(t, b, e2); std::vector<uint32_t>& h = hist[t];
        for (int e = b; e < e2; e++) {
          if (kind[e] != 2) continue;
          row_pair_entries(&S.pts.slot_f[S.pts.slot_ptr[e]], S.pts.nslots[e], 0u, [&](int m, int r, uint32_t) { h[(size_t)m * kRanges + r]++; });
        }
      }
      for (size_t l = 0; l < nl; l++) {
        uint32_t acc = cnt[l];
        for (int t = 0; t < nt; t++) { const uint32_t c = hist[t][l]; hist[t][l] = acc; acc += c; }
        cnt[l + 1] = acc;
      }
      R.ent.resize(cnt.back());
#pragma omp parallel for num_threads(nt) schedule(static, 1)
      for (int t = 0; t < nt; t++) {
        int b, e2; chunk(t, b, e2); std::vector<uint32_t>& cur = hist[t];
        for (int e = b; e < e2; e++) {
          if (kind[e] != 2) continue;
          row_pair_entries(&S.pts.slot_f[S.pts.slot_ptr[e]], S.pts.nslots[e], dptr[e], [&](int m, int r, uint32_t ent) { R.ent[cur[(size_t)m * kRanges + r]++] = ent; });
        }
      }
    }
    // A warp walks its item's entries one after the other, so a small problem gets SHORT items (the longest item is the
    // kernel's run time: 54 us for the 128-entry items of a 50-keyframe window) and a large one the full kRowItemEnts, which
    // amortise the flush of the 2 x 5 accumulators: aim at ~4096 items, 16 <= entries per item <= kRowItemEnts.
    uint32_t item_ents = 16;
    while (item_ents < (uint32_t)kRowItemEnts && (uint64_t)item_ents * 4096u < R.ent.size()) item_ents *= 2;
    for (int m = 0; m < npair; m++) for (int r = 0; r < kRanges; r++) {
      const uint32_t b0 = cnt[(size_t)m * kRanges + r], b1 = cnt[(size_t)m * kRanges + r + 1];
      for (uint32_t o = b0; o < b1; o += item_ents) R.items.push_back({(uint32_t)m, (uint32_t)(5 * r), o, std::min<uint32_t>(item_ents, b1 - o)});
    }
    R.rowblk.assign((size_t)(2 * npair) * kRowSpan, 0xFFFFFFFFu);   // (an odd nf leaves one all-empty row behind the last pair)
    for (int a = 0; a < nf; a++) for (int d = 0; d < kRowSpan && a + d < nf; d++)
      if ((bits[(size_t)a * W + ((a + d) >> 6)] >> ((a + d) & 63)) & 1) R.rowblk[(size_t)a * kRowSpan + d] = blk_of(a, a + d);
  }
  for (RelRec& r : S.rel) {
    r.blk11 = r.f1 >= 0 ? (int32_t)blk_of(r.f1, r.f1) : -1;
    r.blk22 = r.f2 >= 0 ? (int32_t)blk_of(r.f2, r.f2) : -1;
    r.blk12 = -1; r.swap12 = 0;
    if (r.f1 >= 0 && r.f2 >= 0 && r.f1 != r.f2) { r.swap12 = r.f1 > r.f2; r.blk12 = (int32_t)blk_of(std::min(r.f1, r.f2), std::max(r.f1, r.f2)); }
  }
  OBVI_TMARK("7");
  // full symmetric BSR
  {
    std::vector<uint32_t> cnt(nf + 1, 0);
    for (int i = 0; i < nf; i++) for (uint32_t q = S.su_ptr[i]; q < S.su_ptr[i + 1]; q++) { cnt[i + 1]++; if ((int)S.su_col[q] != i) cnt[S.su_col[q] + 1]++; }
    for (int i = 0; i < nf; i++) cnt[i + 1] += cnt[i];
    S.sf_ptr = cnt; S.sf_col.resize(cnt[nf]); S.sf_src.resize(cnt[nf]);
    std::vector<uint32_t> cur(cnt.begin(), cnt.end() - 1);
    // lower part first (cols < i, ascending), produced by scanning rows in order
    for (int i = 0; i < nf; i++) for (uint32_t q = S.su_ptr[i]; q < S.su_ptr[i + 1]; q++) {
      const int j = (int)S.su_col[q];
      if (j == i) continue;
      S.sf_col[cur[j]] = (uint32_t)i; S.sf_src[cur[j]] = q | 0x80000000u; cur[j]++;
    }
    for (int i = 0; i < nf; i++) for (uint32_t q = S.su_ptr[i]; q < S.su_ptr[i + 1]; q++) { S.sf_col[cur[i]] = S.su_col[q]; S.sf_src[cur[i]] = q; cur[i]++; }
  }
  OBVI_TMARK("8");
  count_reduced(pb, S);
  return true;
}

}  // namespace obvi
