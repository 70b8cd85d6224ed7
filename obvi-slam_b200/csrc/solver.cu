// LM driver + C ABI (include/obvi_ba.h).  Host C++ orchestrates; all arithmetic is in ba_kernels.cuh.
// Replaces ceres::Solve(options, problem, &summary) as configured by
// ObjectPoseGraphOptimizer::solveOptimization (include/refactoring/optimization/object_pose_graph_optimizer.h:634-707)
// with Ceres' TrustRegionMinimizer / LevenbergMarquardtStrategy semantics (SURVEY.md Appendix B -- external
// knowledge of upstream Ceres): Jacobi scaling from the first Jacobian, clamped LM diagonal, exact-enough Schur
// solve, rho-based radius update, non-monotonic acceptance, the three tolerance tests, min-cost iterate returned.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <condition_variable>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include <cub/cub.cuh>

#include "ba_kernels.cuh"
#include "bt_precond.cuh"

namespace obvi {

static thread_local std::string g_create_error;

#define CUDA_OK(expr)                                                                                          \
  do {                                                                                                         \
    cudaError_t e__ = (expr);                                                                                  \
    if (e__ != cudaSuccess) {                                                                                  \
      char buf__[512];                                                                                         \
      snprintf(buf__, sizeof(buf__), "CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__, __LINE__,  \
               cudaGetErrorString(e__));                                                                       \
      throw std::runtime_error(buf__);                                                                         \
    }                                                                                                          \
  } while (0)

// ---- NCCL through dlopen: the library loads without NCCL; only obvi_comm_* need it -------------------------
struct Nccl {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load(std::string& err) {
    if (h) return true;
    h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { err = std::string("cannot load NCCL: ") + dlerror(); return false; }
    GetUniqueId = (decltype(GetUniqueId))dlsym(h, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(h, "ncclCommInitRank");
    AllReduce = (decltype(AllReduce))dlsym(h, "ncclAllReduce");
    Broadcast = (decltype(Broadcast))dlsym(h, "ncclBroadcast");
    CommDestroy = (decltype(CommDestroy))dlsym(h, "ncclCommDestroy");
    GetErrorString = (decltype(GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !AllReduce || !Broadcast || !CommDestroy) { err = "NCCL symbols missing"; return false; }
    return true;
  }
};
static Nccl g_nccl;

// ---- communicators ------------------------------------------------------------------------------------------
// The sharded solve needs three collectives on device buffers of doubles: all-reduce (sum), all-reduce (max), broadcast
// from rank 0.  NcclComm is the multi-process path (one process per GPU, NVLink / NVSwitch).  LocalComm joins several
// problem handles of ONE process (each driven by its own host thread, on the same or on peer-accessible devices): the
// ranks rendezvous on a host barrier and reduce each other's buffers with a plain kernel.  It exists so that the sharded
// code path -- structure partitioning, every sharded kernel, the merged write-back -- can be run and checked on a box
// with a single GPU, where NCCL refuses two ranks on one device.
struct Comm {
  virtual ~Comm() {}
  virtual void allreduce_sum(double* p, size_t n, cudaStream_t s, int rank) = 0;
  virtual void allreduce_max(double* p, size_t n, cudaStream_t s, int rank) = 0;
  virtual void broadcast0(double* p, size_t n, cudaStream_t s, int rank) = 0;
};
struct NcclComm : Comm {
  ncclComm_t comm = nullptr;
  ~NcclComm() override { if (comm && g_nccl.CommDestroy) g_nccl.CommDestroy(comm); }
  static void ck(ncclResult_t r, const char* what) { if (r != ncclSuccess) throw std::runtime_error(std::string(what) + ": " + g_nccl.GetErrorString(r)); }
  void allreduce_sum(double* p, size_t n, cudaStream_t s, int) override { ck(g_nccl.AllReduce(p, p, n, ncclDouble, ncclSum, comm, s), "ncclAllReduce"); }
  void allreduce_max(double* p, size_t n, cudaStream_t s, int) override { ck(g_nccl.AllReduce(p, p, n, ncclDouble, ncclMax, comm, s), "ncclAllReduce"); }
  void broadcast0(double* p, size_t n, cudaStream_t s, int) override { ck(g_nccl.Broadcast(p, p, n, ncclDouble, 0, comm, s), "ncclBroadcast"); }
};
__global__ void local_reduce_kernel(double* __restrict__ out, const double* const* __restrict__ in, int world, size_t n, int is_max) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v = in[0][i];
  for (int r = 1; r < world; r++) v = is_max ? fmax(v, in[r][i]) : v + in[r][i];   // fixed rank order: every rank gets the same bits
  out[i] = v;
}
struct LocalComm : Comm {
  int world;
  std::mutex m;
  std::condition_variable cv;
  int arrived = 0;
  uint64_t gen = 0;
  std::vector<double*> ptr;
  explicit LocalComm(int w) : world(w), ptr(w, nullptr) {}
  void barrier() {
    std::unique_lock<std::mutex> lk(m);
    const uint64_t g = gen;
    if (++arrived == world) { arrived = 0; gen++; cv.notify_all(); }
    else cv.wait(lk, [&] { return gen != g; });
  }
  void reduce(double* p, size_t n, cudaStream_t s, int rank, int is_max) {
    if (n == 0) return;
    CUDA_OK(cudaStreamSynchronize(s));          // p is final on this rank
    ptr[rank] = p;
    barrier();                                  // ... and on every other rank
    double* tmp = nullptr; const double** d_in = nullptr;
    CUDA_OK(cudaMalloc((void**)&tmp, n * sizeof(double)));
    CUDA_OK(cudaMalloc((void**)&d_in, world * sizeof(double*)));
    CUDA_OK(cudaMemcpyAsync(d_in, ptr.data(), world * sizeof(double*), cudaMemcpyHostToDevice, s));
    local_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(tmp, d_in, world, n, is_max);
    CUDA_OK(cudaStreamSynchronize(s));
    barrier();                                  // everybody has read everybody's input
    CUDA_OK(cudaMemcpyAsync(p, tmp, n * sizeof(double), cudaMemcpyDeviceToDevice, s));
    CUDA_OK(cudaStreamSynchronize(s));
    cudaFree(tmp); cudaFree(d_in);
  }
  void allreduce_sum(double* p, size_t n, cudaStream_t s, int rank) override { reduce(p, n, s, rank, 0); }
  void allreduce_max(double* p, size_t n, cudaStream_t s, int rank) override { reduce(p, n, s, rank, 1); }
  void broadcast0(double* p, size_t n, cudaStream_t s, int rank) override {
    if (n == 0) return;
    CUDA_OK(cudaStreamSynchronize(s));
    ptr[rank] = p;
    barrier();
    if (rank != 0) { CUDA_OK(cudaMemcpyAsync(p, ptr[0], n * sizeof(double), cudaMemcpyDefault, s)); CUDA_OK(cudaStreamSynchronize(s)); }
    barrier();
  }
};

// ---- device-memory cache --------------------------------------------------------------------------------
// The reference's schedule builds one problem per window (4069 solves at BASELINE config 4): ~100 device buffers each,
// and cudaMalloc / cudaFree cost 0.1 - 1 ms apiece (cudaFree also synchronises the device).  Freed blocks are kept per
// device and handed to the next request of about that size (<= 25 % larger).  A block may come back while work that used it
// is still in flight on another problem's stream, so the first reuse after any release synchronises the device once.
// OBVI_NO_CACHE=1 turns it off; OBVI_CACHE_MB caps the cached bytes (default 8192).
struct DevCache {
  std::mutex mu;
  std::multimap<size_t, void*> blocks;
  size_t cached = 0, limit = (size_t)8192 << 20;
  bool dirty = false, off = false;
  DevCache() {
    if (const char* e = getenv("OBVI_NO_CACHE")) off = std::string(e) == "1";
    if (const char* e = getenv("OBVI_CACHE_MB")) limit = (size_t)std::max(0, atoi(e)) << 20;
  }
  static DevCache& get() {
    static DevCache c[64];
    int dev = 0;
    cudaGetDevice(&dev);
    return c[dev & 63];
  }
  void* take(size_t bytes, size_t& got) {
    if (off) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    auto it = blocks.lower_bound(bytes);
    if (it == blocks.end() || it->first > bytes + bytes / 4 + 4096) return nullptr;
    void* p = it->second; got = it->first; cached -= got; blocks.erase(it);
    if (dirty) { cudaDeviceSynchronize(); dirty = false; }
    return p;
  }
  void release_all() {
    std::lock_guard<std::mutex> lk(mu);
    if (dirty) { cudaDeviceSynchronize(); dirty = false; }
    for (auto& b : blocks) cudaFree(b.second);
    blocks.clear(); cached = 0;
  }
  void give(void* p, size_t bytes) {
    {
      std::lock_guard<std::mutex> lk(mu);
      if (!off && cached + bytes <= limit) {
        // a recycled block reads as zeros, like the fresh pages cudaMalloc hands out (ordered before the reuse by the
        // device synchronisation in take ())
        cudaMemsetAsync(p, 0, bytes, 0);
        blocks.emplace(bytes, p); cached += bytes; dirty = true;
        return;
      }
    }
    cudaFree(p);
  }
};
// pinned host blocks of one size (the per-problem scalar read-back buffer): cudaMallocHost costs ~1 ms
struct PinnedPool {
  std::mutex mu;
  std::vector<void*> blocks;
  size_t bytes = 0;
  static PinnedPool& get() { static PinnedPool p; return p; }
  void* take(size_t n) {
    {
      std::lock_guard<std::mutex> lk(mu);
      if (bytes == n && !blocks.empty()) { void* p = blocks.back(); blocks.pop_back(); return p; }
    }
    void* p = nullptr;
    CUDA_OK(cudaMallocHost(&p, n));
    return p;
  }
  void give(void* p, size_t n) {
    std::lock_guard<std::mutex> lk(mu);
    if (bytes == 0) bytes = n;
    if (bytes == n && blocks.size() < 64) blocks.push_back(p); else cudaFreeHost(p);
  }
};

// ---- small RAII device buffer ---------------------------------------------------------------------------
template <class T>
struct DBuf {
  T* p = nullptr;
  size_t n = 0;          // elements requested
  size_t cap = 0;        // bytes of the block behind p
  void alloc(size_t count) {
    if (p && count * sizeof(T) <= cap) { n = std::max(n, count); return; }
    free();
    if (count == 0) count = 1;
    const size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
    void* q = DevCache::get().take(bytes, cap);
    if (!q) {
      if (cudaMalloc(&q, bytes) != cudaSuccess) {       // out of memory with blocks parked in the cache: release them and retry
        (void)cudaGetLastError();
        DevCache::get().release_all();
        CUDA_OK(cudaMalloc(&q, bytes));
      }
      cap = bytes;
    }
    p = (T*)q;
    n = count;
  }
  void upload(const std::vector<T>& v, cudaStream_t s) {
    alloc(v.size());
    if (!v.empty()) CUDA_OK(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void zero(cudaStream_t s) { if (p) CUDA_OK(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
  void free() { if (p) DevCache::get().give(p, cap); p = nullptr; n = 0; cap = 0; }
  ~DBuf() { free(); }
};

struct EListDev {
  DBuf<uint32_t> ptr, pos, pair_ptr, pair_blk, overflow_off;
  DBuf<int32_t> f;
  DBuf<uint16_t> slot, nslots;
  DBuf<uint8_t> cst;
  DBuf<double> escale, einv, eg, prior_H, prior_g, overflow, delta;
  int ne = 0;
  bool has_prior = false;
};

// In-situ phase timer (OBVI_PROFILE=1): CUDA events recorded on the main stream around named phases, resolved when the
// solve ends.  Timing in place sees warm caches and real overlap, unlike a serialised ncu launch list.
struct PhaseProfiler {
  bool on = false;
  std::vector<cudaEvent_t> pool;
  struct Span { int tag; size_t e0, e1; };
  std::vector<Span> spans;
  std::vector<const char*> names;
  size_t used = 0;
  size_t ev(cudaStream_t s) {
    if (used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
    cudaEventRecord(pool[used], s);
    return used++;
  }
  int tag(const char* n) { for (size_t i = 0; i < names.size(); i++) if (names[i] == n) return (int)i; names.push_back(n); return (int)names.size() - 1; }
  size_t begin(cudaStream_t s) { return on ? ev(s) : 0; }
  void end(const char* n, size_t e0, cudaStream_t s) { if (on) spans.push_back({tag(n), e0, ev(s)}); }
  void report(int steps) {
    if (!on) return;
    std::vector<double> tot(names.size(), 0.0); std::vector<int> cnt(names.size(), 0);
    for (const Span& sp : spans) { float ms = 0; cudaEventElapsedTime(&ms, pool[sp.e0], pool[sp.e1]); tot[sp.tag] += ms; cnt[sp.tag]++; }
    double all = 0; for (double t : tot) all += t;
    fprintf(stderr, "[obvi profile] %d LM steps, %.3f ms in timed phases (%.3f ms / step)\n", steps, all, all / std::max(steps, 1));
    for (size_t i = 0; i < names.size(); i++) fprintf(stderr, "[obvi profile]   %-22s n=%4d  mean %8.1f us  total %8.3f ms  %5.1f%%\n", names[i], cnt[i], 1e3 * tot[i] / std::max(cnt[i], 1), tot[i], 100.0 * tot[i] / all);
    spans.clear(); used = 0;
  }
  ~PhaseProfiler() { for (cudaEvent_t e : pool) cudaEventDestroy(e); }
};

struct Solver {
  Problem pb;
  PhaseProfiler prof;
  Structure st;
  cudaStream_t stream = nullptr;
  cudaStream_t s3 = nullptr;       // second side stream (candidate cost: bbox on s2, priors + rel-pose on s3)
  cudaEvent_t ev_join3 = nullptr;
  cudaStream_t s2 = nullptr;       // side stream: the small kernels (objects, priors, rel-pose) overlap the big point kernels
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev[8] = {};
  bool uploaded = false;
  int num_sms = 0;
  int pcg_blocks_per_sm = 0;
  // communicator
  std::shared_ptr<Comm> comm;
  int rank = 0, world = 1;
  int debug_skip = 0;            // OBVI_DEBUG_SKIP bit mask: 1 point_prep, 2 schur_rows, 4 backsub_rows (hang bisection only; results are garbage)
  int debug_bt_fail_rank = -1;   // OBVI_DEBUG_BT_FAIL_RANK: this rank reports a failed factorisation (tests of the fallback path)
  // device structure
  DBuf<Camera> cams;
  DBuf<CalibClass> classes;
  DBuf<ObsRec> obs;
  DBuf<uint32_t> pose_ptr, su_ptr, sf_ptr, sf_col, sf_src;
  DBuf<int32_t> f_of_pose, pose_of_f;
  DBuf<uint8_t> pose_skip, point_skip, obj_skip;
  DBuf<uint4> jac_tile;  // per 256-observation tile: first pose/camera entry, entries staged by TMA, first keyframe
  int jac_mode = 1;  // 1: TMA-staged tile kernel (default), 0: plain loads / stores (OBVI_JAC=plain)
  DBuf<double> pose_acc;    // per variable pose: H_pp (36) | g_p (6) | spare (6) of the reprojection blocks, filled by the Jacobian kernel
  DBuf<BBoxRec> bbox;
  DBuf<UnaryRec> unary;
  DBuf<RelRec> rel;
  EListDev pts, objs;
  // row-owner point elimination (point_prep_kernel + schur_rows_kernel)
  DBuf<uint32_t> pr_grp_ptr, pr_rowblk, pr_fallback;
  DBuf<Structure::RowGroup> pr_grp;
  DBuf<uint32_t> pr_ent;
  DBuf<Structure::RowItem> pr_items;
  DBuf<uint8_t> pr_regular;
  int pr_batch_begin = 0, pr_batch_end = 0;
  DBuf<double> WZ;
  int n_row_items = 0, n_row_fallback = 0;
  // in-place removal of reprojection / bbox blocks (two-phase outlier exclusion without a structure rebuild)
  std::vector<uint32_t> inv_rp, inv_bb;     // user factor index -> internal position (0xFFFFFFFF: not on this rank)
  std::vector<uint32_t> pend_rp, pend_bb;   // internal positions whose device flag still has to be set
  int64_t n_masked_rp = 0, n_masked_bb = 0;
  bool counts_stale = false;
  int64_t structure_builds = 0;
  // state
  DBuf<double> poses[3], points[3], objects[3];  // cur, cand, best
  int cur = 0;
  DBuf<PoseCam> pcam, pcam_cand;
  DBuf<double> J, Jb;
  DBuf<UnaryOut> unary_out;
  DBuf<RelOut> rel_out;
  DBuf<double> redbuf;  // [S_upper | gp | b_schur | hpp_diag]
  double *S_upper = nullptr, *gp = nullptr, *b_schur = nullptr, *hpp_diag = nullptr, *red_tail = nullptr;
  bool stage0_local = false;   // sharded run: scalars 0-2 hold this rank's partial sums (set by linearize (), cleared by the reduction)
  DBuf<double> pscale, Sf, rhs, Minv, y, cg_r, cg_z, cg_p, cg_q, cg_acc, dpose, scalars;
  // block-tridiagonal preconditioner (bt_precond.cuh)
  int nsb = 0, nlev = 0, pcg_bt_blocks_per_sm = 0;
  DBuf<double> bt_D, bt_Dinv, bt_GaT, bt_GcT, bt_C, bt_w, bt_z;
  DBuf<GemmTask> bt_tasks;
  DBuf<int> bt_idx;
  struct BtLevel { int inv_off, inv_n, g_off, g_n, u_off, u_n; };
  std::vector<BtLevel> bt_levels;
  int bt_last_inv_off = 0;
  bool use_bt = true;
  bool pcg_resident = true;  // OBVI_PCG=grid: the grid-barrier PCG kernel instead of the shared-memory-resident dataflow one
  bool resident_ok = false;
  DBuf<double> rs_dbl;        // [16 reduction slots | nsb * 96 forward accumulators]
  DBuf<uint4> ll_buf;         // flag-in-data hand-off words of pcg_bt_ll_kernel: [4 nsb reductions | nsb x 16 x 96 pushes | nsb x 96 z | nsb x 96 p]
  unsigned int ll_tag = 0;    // last tag handed out (tags only grow, the buffer is cleared when they would wrap)
  bool pcg_ll = true;         // OBVI_PCG=flags: the resident kernel with counter / flag hand-offs
  DBuf<unsigned int> rs_u32;  // [4 reduction counters | nsb arrival counters | nsb z epochs]
  bool defer_sync = true;        // OBVI_DEFER_SYNC=0: two host synchronisations per accepted LM iteration instead of one
  bool obj_split = true;      // OBVI_OBJ_SPLIT=0: one-kernel object elimination (254 registers; kept for A/B runs and tests)
  bool bt_v1 = false;   // OBVI_BT=v1: first-generation factorisation kernels (scalar-pivot Gauss-Jordan, FMA GEMM)
  // The factorisation is reused across LM iterations while it still preconditions well: it is redone when the
  // trust-region radius moved by more than 2x since it was computed or the last PCG needed more than
  // kRefactorPcgIters iterations.  (A stale preconditioner changes the PCG iteration count, never its answer.)
  double bt_radius = -1.0;
  int last_pcg_iters = 0;
  int bt_fresh_iters = 0;        // PCG iterations of the first solve after the current factorisation (0: not seen yet)
  bool bt_fresh_pending = false;
  int kRefactorPcgIters = 3;   // OBVI_REFACTOR_ITERS: refactor once the PCG needs this many iterations more than with a fresh factorisation
  double kRefactorRatio = 2.0; // OBVI_REFACTOR_RATIO
  int64_t bt_factorizations = 0;
  double* h_scalars = nullptr;  // pinned
  std::vector<double> h_poses, h_points, h_objects;
  int64_t launches = 0;

  ~Solver() {
    if (h_scalars) PinnedPool::get().give(h_scalars, SC_COUNT * sizeof(double));
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    if (s2) cudaStreamDestroy(s2);
    if (s3) cudaStreamDestroy(s3);
    if (ev_join3) cudaEventDestroy(ev_join3);
    if (stream) cudaStreamDestroy(stream);
  }

  void init_device() {
    CUDA_OK(cudaSetDevice(pb.device));
    {   // (cudaGetDeviceProperties takes about a millisecond: asked once per device, not once per problem)
      static std::mutex mu;
      static std::map<int, std::pair<int, int>> seen;     // device -> (compute capability major, SM count)
      std::lock_guard<std::mutex> lk(mu);
      auto it = seen.find(pb.device);
      if (it == seen.end()) {
        cudaDeviceProp prop;
        CUDA_OK(cudaGetDeviceProperties(&prop, pb.device));
        if (prop.major < 10) throw std::runtime_error(std::string("obvi_ba is built for sm_100a (B200); found ") + prop.name);
        it = seen.emplace(pb.device, std::make_pair(prop.major, prop.multiProcessorCount)).first;
      }
      num_sms = it->second.second;
    }
    CUDA_OK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    // the side streams carry the small latency-bound kernels (objects, priors, rel-pose): highest priority, so that their
    // few CTAs are placed as soon as a slot frees up instead of queueing behind the grid of a big point kernel
    int prio_lo = 0, prio_hi = 0;
    CUDA_OK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    CUDA_OK(cudaStreamCreateWithPriority(&s2, cudaStreamNonBlocking, prio_hi));
    CUDA_OK(cudaStreamCreateWithPriority(&s3, cudaStreamNonBlocking, prio_hi));
    CUDA_OK(cudaEventCreateWithFlags(&ev_join3, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    for (auto& e : ev) CUDA_OK(cudaEventCreate(&e));
    h_scalars = (double*)PinnedPool::get().take(SC_COUNT * sizeof(double));
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pcg_blocks_per_sm, pcg_kernel, kPcgThreads, 0));
    if (pcg_blocks_per_sm < 1) throw std::runtime_error("pcg_kernel cannot be made resident");
    CUDA_OK(cudaFuncSetAttribute(bt_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kBB * 8));
    CUDA_OK(cudaFuncSetAttribute(bt_gemm_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pcg_bt_blocks_per_sm, pcg_bt_kernel, kPcgThreads, 0));
    if (pcg_bt_blocks_per_sm < 1) throw std::runtime_error("pcg_bt_kernel cannot be made resident");
    if (const char* e = getenv("OBVI_PRECOND")) use_bt = std::string(e) != "jacobi";
    if (const char* e = getenv("OBVI_BT")) bt_v1 = std::string(e) == "v1";
    if (const char* e = getenv("OBVI_REFACTOR_ITERS")) kRefactorPcgIters = std::max(1, atoi(e));
    if (const char* e = getenv("OBVI_REFACTOR_RATIO")) kRefactorRatio = std::max(1.0, atof(e));
    if (const char* e = getenv("OBVI_DEFER_SYNC")) defer_sync = std::string(e) != "0";
    if (const char* e = getenv("OBVI_OBJ_SPLIT")) obj_split = std::string(e) != "0";
    if (const char* e = getenv("OBVI_PCG")) { pcg_resident = std::string(e) != "grid"; pcg_ll = std::string(e) != "flags"; }
    CUDA_OK(cudaFuncSetAttribute(pcg_bt_ll_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kResidentSmem));
    CUDA_OK(cudaFuncSetAttribute(pcg_bt_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kResidentSmem));
    if (const char* e = getenv("OBVI_PROFILE")) prof.on = std::string(e) == "1";
    if (const char* e = getenv("OBVI_DEBUG_BT_FAIL_RANK")) debug_bt_fail_rank = atoi(e);
    if (const char* e = getenv("OBVI_DEBUG_SKIP")) debug_skip = atoi(e);
    if (const char* e = getenv("OBVI_JAC")) jac_mode = std::string(e) == "plain" ? 0 : 1;
    CUDA_OK(cudaFuncSetAttribute(reproj_jac_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kJacSmemBytes));
    CUDA_OK(cudaFuncSetAttribute(point_prep_kernel<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, pipe_smem(8, 2)));
    CUDA_OK(cudaFuncSetAttribute(backsub_rows_kernel<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, pipe_smem(8, 2)));
    CUDA_OK(cudaFuncSetAttribute(point_prep_kernel<16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, pipe_smem(16, 1)));
    CUDA_OK(cudaFuncSetAttribute(backsub_rows_kernel<16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, pipe_smem(16, 1)));
    if (const char* e = getenv("OBVI_PIPE")) pipe_warps = std::string(e) == "16" ? 16 : 8;
    if (const char* e = getenv("OBVI_ROW_STAGES")) row_stages = atoi(e);
    CUDA_OK(cudaFuncSetAttribute(schur_rows_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, rows_smem_bytes(8)));
    CUDA_OK(cudaFuncSetAttribute(schur_rows_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, rows_smem_bytes(4)));
  }

  void upload_elist(EListDev& D, const Structure::EList& L, const std::vector<uint8_t>& cst, int NE, int maxs) {
    D.ne = (int)cst.size();
    D.ptr.upload(L.ptr, stream); D.pos.upload(L.pos, stream); D.f.upload(L.f, stream); D.slot.upload(L.slot, stream);
    D.pair_ptr.upload(L.pair_ptr, stream); D.nslots.upload(L.nslots, stream); D.pair_blk.upload(L.pair_blk, stream);
    D.cst.upload(cst, stream);
    D.escale.alloc((size_t)D.ne * NE); D.einv.alloc((size_t)D.ne * NE * NE); D.eg.alloc((size_t)D.ne * NE);
    D.delta.alloc((size_t)D.ne * NE); D.delta.zero(stream);
    std::vector<uint32_t> off(D.ne, 0);
    size_t tot = 0;
    for (int e = 0; e < D.ne; e++) if (L.nslots[e] > maxs) { off[e] = (uint32_t)tot; tot += (size_t)L.nslots[e] * 12 * NE; }
    D.overflow_off.upload(off, stream); D.overflow.alloc(tot);
  }

  void upload_structure() {
    const Structure& S = st;
    cams.upload(pb.cams, stream); classes.upload(S.classes, stream); obs.upload(S.obs, stream);
    pose_ptr.upload(S.pose_ptr, stream); f_of_pose.upload(S.f_of_pose, stream);
    std::vector<int32_t> pof(S.nf);
    for (int k = 0; k < S.K; k++) if (S.f_of_pose[k] >= 0) pof[S.f_of_pose[k]] = k;
    pose_of_f.upload(pof, stream);
    su_ptr.upload(S.su_ptr, stream); sf_ptr.upload(S.sf_ptr, stream); sf_col.upload(S.sf_col, stream); sf_src.upload(S.sf_src, stream);
    bbox.upload(S.bbox, stream); unary.upload(S.unary, stream); rel.upload(S.rel, stream);
    {
      const int64_t ntiles = (S.n_obs + kJacThreads - 1) / kJacThreads;
      std::vector<uint4> tp(ntiles);
      for (int64_t t = 0; t < ntiles; t++) {
        const ObsRec& a = S.obs[t * kJacThreads];
        const ObsRec& b = S.obs[std::min<int64_t>(S.n_obs, (t + 1) * kJacThreads) - 1];
        const uint32_t lo = a.pose * (uint32_t)S.C + (uint32_t)S.classes[obs_cls(a)].cam, hi = b.pose * (uint32_t)S.C + (uint32_t)S.classes[obs_cls(b)].cam;
        tp[t] = make_uint4(lo, std::min<uint32_t>(hi - lo + 1, kJacMaxPc), a.pose, 0u);
      }
      jac_tile.upload(tp, stream);
    }
    upload_elist(pts, S.pts, S.point_const, 3, 16);
    {   // device view: entry d of the point lists IS chunk d (host pts.pos keeps the pose-major record index)
      std::vector<uint32_t> ident(S.n_obs);
      for (int64_t d = 0; d < S.n_obs; d++) ident[d] = (uint32_t)d;
      pts.pos.upload(ident, stream);
    }
    upload_elist(objs, S.objs, S.obj_const, 7, kObjMaxSlots);   // overflow areas for every object the split kernel cannot stage on chip
    {
      const Structure::PointRows& R = S.prow;
      n_row_items = (int)R.items.size(); n_row_fallback = (int)R.fallback.size();
      pr_grp_ptr.upload(R.grp_ptr, stream); pr_grp.upload(R.grp, stream); pr_regular.upload(R.regular, stream);
      {   // batches of four points between the first and the last regular point of this rank
        int first = -1, last = -1;
        for (int e = 0; e < (int)R.regular.size(); e++) if (R.regular[e]) { if (first < 0) first = e; last = e; }
        pr_batch_begin = first < 0 ? 0 : first / 4; pr_batch_end = first < 0 ? 0 : last / 4 + 1;
      }
      pr_ent.upload(R.ent, stream); pr_items.upload(R.items, stream); pr_rowblk.upload(R.rowblk, stream); pr_fallback.upload(R.fallback, stream);
      WZ.alloc((size_t)std::max<int64_t>(R.n_slots, 1) * kWZ); WZ.zero(stream);   // gap slots stay zero
    }
    inv_rp.assign(pb.reproj.size(), 0xFFFFFFFFu); inv_bb.assign(pb.bbox.size(), 0xFFFFFFFFu);
    for (int64_t q = 0; q < S.n_obs; q++) inv_rp[S.obs_user[q]] = (uint32_t)q;
    for (int64_t q = 0; q < S.n_bbox; q++) inv_bb[S.bbox_user[q]] = (uint32_t)q;
    pend_rp.clear(); pend_bb.clear(); n_masked_rp = n_masked_bb = 0; counts_stale = false;
    pts.has_prior = objs.has_prior = false;
    for (const UnaryRec& u : S.unary) { if (u.kind == 1) pts.has_prior = true; if (u.kind == 2) objs.has_prior = true; }
    if (pts.has_prior) { pts.prior_H.alloc((size_t)S.P * 9); pts.prior_g.alloc((size_t)S.P * 3); }
    if (objs.has_prior) { objs.prior_H.alloc((size_t)S.O * 49); objs.prior_g.alloc((size_t)S.O * 7); }
    // x-norm masks: a block counts iff it is variable and (multi-GPU) this rank contributes it
    std::vector<uint8_t> ps(S.K), pt(S.P), ob(S.O);
    for (int k = 0; k < S.K; k++) ps[k] = (S.f_of_pose[k] < 0) || rank != 0;
    for (int i = 0; i < S.P; i++) pt[i] = S.point_const[i] || S.pts.ptr[i] == S.pts.ptr[i + 1];
    for (int i = 0; i < S.O; i++) ob[i] = S.obj_const[i] || !owns_object(i);
    pose_skip.upload(ps, stream); point_skip.upload(pt, stream); obj_skip.upload(ob, stream);
    for (int b = 0; b < 3; b++) { poses[b].alloc((size_t)S.K * 6); points[b].alloc((size_t)S.P * 3); objects[b].alloc((size_t)S.O * 7); }
    pcam.alloc((size_t)S.K * std::max(S.C, 1)); pcam_cand.alloc((size_t)S.K * std::max(S.C, 1));
    J.alloc((size_t)S.n_obs * kChunk); Jb.alloc((size_t)S.n_bbox * kBBoxChunk);
    pose_acc.alloc((size_t)std::max(S.nf, 1) * kPoseAcc);
    unary_out.alloc(S.n_unary); rel_out.alloc(S.n_rel);
    const size_t nf6 = (size_t)S.nf * 6;
    // a sharded run all-reduces this buffer once per build; its tail carries the linearisation's scalars along
    // (cost, fixed cost, |x|^2 as sums; one (gradient max, failure count) slot pair per rank): see pack_stage0_kernel
    redbuf.alloc((size_t)S.n_upper * 36 + 3 * nf6 + kStage0Sums + 2 * (size_t)world);
    S_upper = redbuf.p; gp = S_upper + (size_t)S.n_upper * 36; b_schur = gp + nf6; hpp_diag = b_schur + nf6; red_tail = hpp_diag + nf6;
    pscale.alloc(nf6); Sf.alloc((size_t)S.sf_col.size() * 36); rhs.alloc(nf6); Minv.alloc((size_t)S.nf * 36);
    y.alloc(nf6 + kPcgStatus); cg_r.alloc(nf6); cg_z.alloc(nf6); cg_p.alloc(nf6); cg_q.alloc(nf6); cg_acc.alloc(16); dpose.alloc(nf6);
    scalars.alloc(SC_COUNT);
    setup_bt();
    uploaded = true;
  }
  // level tables + GEMM task lists of the cyclic-reduction factorisation (static per structure)
  void setup_bt() {
    const Structure& S = st;
    nsb = (S.nf + kSbPoses - 1) / kSbPoses;
    nlev = 0;
    while ((1 << nlev) < nsb) nlev++;
    bt_levels.clear();
    if (nsb == 0) return;
    {
      int per_sm = 0;
      CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pcg_bt_resident_kernel, kPcgThreads, kResidentSmem));
      resident_ok = pcg_resident && per_sm >= 1 && nsb <= num_sms * per_sm;
      rs_dbl.alloc(16 + (size_t)nsb * kB); rs_u32.alloc(4 + 2 * (size_t)nsb);
      static_assert(kLLSlots >= 2 * 8, "levels of the elimination tree");
      if (nlev > kLLSlots / 2) resident_ok = false;
      ll_buf.alloc(4 * (size_t)nsb + (size_t)nsb * kLLSlots * kB + 2 * (size_t)nsb * kB); ll_buf.zero(stream); ll_tag = 0;
    }
    bt_D.alloc((size_t)nsb * kBB); bt_Dinv.alloc((size_t)nsb * kBB); bt_GaT.alloc((size_t)nsb * kBB); bt_GcT.alloc((size_t)nsb * kBB);
    bt_w.alloc((size_t)nsb * kB); bt_z.alloc((size_t)nsb * kB);
    // couplings: level l holds n_l - 1 blocks
    std::vector<size_t> coff(nlev + 1, 0);
    size_t ctot = 0;
    for (int l = 0; l <= nlev; l++) { const int s = 1 << l, nl = (nsb + s - 1) / s; coff[l] = ctot; ctot += (size_t)std::max(nl - 1, 0); }
    bt_C.alloc(std::max<size_t>(ctot, 1) * kBB);
    std::vector<GemmTask> tasks;
    std::vector<int> idx;
    auto Cl = [&](int l, int k) { return bt_C.p + (coff[l] + (size_t)k) * kBB; };
    for (int l = 0; l < nlev; l++) {
      const int s = 1 << l, nl = (nsb + s - 1) / s;
      BtLevel L;
      L.inv_off = (int)idx.size();
      for (int k = 1; k < nl; k += 2) idx.push_back(k * s);
      L.inv_n = (int)idx.size() - L.inv_off;
      L.g_off = (int)tasks.size();
      for (int k = 1; k < nl; k += 2) {
        const size_t i = (size_t)k * s;
        GemmTask t{}; t.C = bt_GaT.p + i * kBB; t.A1 = Cl(l, k - 1); t.tA1 = 1; t.B1 = bt_Dinv.p + i * kBB; t.alpha1 = 1.0; t.beta = 0.0;
        tasks.push_back(t);
        if (k + 1 < nl) { GemmTask u{}; u.C = bt_GcT.p + i * kBB; u.A1 = Cl(l, k); u.B1 = bt_Dinv.p + i * kBB; u.alpha1 = 1.0; u.beta = 0.0; tasks.push_back(u); }
      }
      L.g_n = (int)tasks.size() - L.g_off;
      L.u_off = (int)tasks.size();
      for (int k = 0; k < nl; k += 2) {
        const size_t a = (size_t)k * s;
        GemmTask t{}; t.C = bt_D.p + a * kBB; t.beta = 1.0;
        int nt = 0;
        if (k + 1 < nl) { t.A1 = bt_GaT.p + (size_t)(k + 1) * s * kBB; t.B1 = Cl(l, k); t.alpha1 = -1.0; nt = 1; }
        if (k >= 1) {
          const double* A = bt_GcT.p + (size_t)(k - 1) * s * kBB; const double* Bm = Cl(l, k - 1);
          if (nt == 0) { t.A1 = A; t.B1 = Bm; t.tB1 = 1; t.alpha1 = -1.0; } else { t.A2 = A; t.B2 = Bm; t.tB2 = 1; t.alpha2 = -1.0; }
          nt++;
        }
        if (nt) tasks.push_back(t);
        if (k + 2 < nl) { GemmTask c{}; c.C = Cl(l + 1, k / 2); c.A1 = bt_GcT.p + (size_t)(k + 1) * s * kBB; c.B1 = Cl(l, k); c.alpha1 = -1.0; c.beta = 0.0; tasks.push_back(c); }
      }
      L.u_n = (int)tasks.size() - L.u_off;
      bt_levels.push_back(L);
    }
    bt_last_inv_off = (int)idx.size();
    idx.push_back(0);
    bt_tasks.upload(tasks, stream); bt_idx.upload(idx, stream);
  }
  void factor_bt() {
    const Structure& S = st;
    if (!use_bt || nsb == 0) return;
    bt_D.zero(stream);
    if (nsb > 1) CUDA_OK(cudaMemsetAsync(bt_C.p, 0, (size_t)(nsb - 1) * kBB * 8, stream));
    bt_assemble_kernel<<<nsb * kSbPoses, 128, 0, stream>>>(S.nf, nsb, sf_ptr.p, sf_col.p, Sf.p, bt_D.p, bt_C.p);
    launches++;
    for (const BtLevel& L : bt_levels) {
      if (bt_v1) {
        if (L.inv_n) { bt_invert_kernel<<<L.inv_n, kInvThreads, 0, stream>>>(bt_idx.p + L.inv_off, bt_D.p, bt_Dinv.p, scalars.p); launches++; }
        if (L.g_n) { bt_gemm_kernel<<<L.g_n, 256, 2 * kBB * 8, stream>>>(bt_tasks.p + L.g_off); launches++; }
        if (L.u_n) { bt_gemm_kernel<<<L.u_n, 256, 2 * kBB * 8, stream>>>(bt_tasks.p + L.u_off); launches++; }
      } else {
        if (L.inv_n) { bt_invert8_kernel<<<L.inv_n, kInvThreads, 0, stream>>>(bt_idx.p + L.inv_off, bt_D.p, bt_Dinv.p, scalars.p); launches++; }
        if (L.g_n) { bt_gemm_mma_kernel<<<L.g_n, 512, kGemmSmem, stream>>>(bt_tasks.p + L.g_off); launches++; }
        if (L.u_n) { bt_gemm_mma_kernel<<<L.u_n, 512, kGemmSmem, stream>>>(bt_tasks.p + L.u_off); launches++; }
      }
    }
    if (bt_v1) bt_invert_kernel<<<1, kInvThreads, 0, stream>>>(bt_idx.p + bt_last_inv_off, bt_D.p, bt_Dinv.p, scalars.p);
    else bt_invert8_kernel<<<1, kInvThreads, 0, stream>>>(bt_idx.p + bt_last_inv_off, bt_D.p, bt_Dinv.p, scalars.p);
    launches++;
    if (debug_bt_fail_rank == rank) { set_scalar_kernel<<<1, 1, 0, stream>>>(scalars.p + SC_BT_FAIL, 1.0); launches++; }
  }
  bool owns_object(int o) const {
    // an object is "owned" when this rank holds its observations or (no observations) rank 0
    if (world <= 1) return true;
    if (st.objs.ptr[o] != st.objs.ptr[o + 1]) return true;
    for (const UnaryRec& u : st.unary) if (u.kind == 2 && u.idx == o) return true;
    return false;
  }

  // ---- parameter gather / scatter --------------------------------------------------------------------
  void gather_params() {
    const Structure& S = st;
    h_poses.resize((size_t)S.K * 6); h_points.resize((size_t)S.P * 3); h_objects.resize((size_t)S.O * 7);
    for (int k = 0; k < S.K; k++) std::memcpy(&h_poses[6 * (size_t)k], pb.blocks[S.pose_block[k]].host, 48);
#pragma omp parallel for schedule(static) if (S.P > 32768)
    for (int i = 0; i < S.P; i++) std::memcpy(&h_points[3 * (size_t)i], pb.blocks[S.point_block[i]].host, 24);
    for (int i = 0; i < S.O; i++) std::memcpy(&h_objects[7 * (size_t)i], pb.blocks[S.obj_block[i]].host, 56);
    if (S.K) CUDA_OK(cudaMemcpyAsync(poses[0].p, h_poses.data(), h_poses.size() * 8, cudaMemcpyHostToDevice, stream));
    if (S.P) CUDA_OK(cudaMemcpyAsync(points[0].p, h_points.data(), h_points.size() * 8, cudaMemcpyHostToDevice, stream));
    if (S.O) CUDA_OK(cudaMemcpyAsync(objects[0].p, h_objects.data(), h_objects.size() * 8, cudaMemcpyHostToDevice, stream));
    for (int b = 1; b < 3; b++) {  // candidate / best buffers start as copies (constant blocks are never rewritten)
      if (S.K) CUDA_OK(cudaMemcpyAsync(poses[b].p, poses[0].p, h_poses.size() * 8, cudaMemcpyDeviceToDevice, stream));
      if (S.P) CUDA_OK(cudaMemcpyAsync(points[b].p, points[0].p, h_points.size() * 8, cudaMemcpyDeviceToDevice, stream));
      if (S.O) CUDA_OK(cudaMemcpyAsync(objects[b].p, objects[0].p, h_objects.size() * 8, cudaMemcpyDeviceToDevice, stream));
    }
    cur = 0;
  }
  void scatter_params(int buf) {
    const Structure& S = st;
    if (world > 1) {
      // every rank owns a slice of the e-blocks: zero what it does not own and sum across ranks
      // (poses are replicated: rank 0 contributes them)
      merge_sharded(poses[buf].p, pose_skip.p, S.K, 6);
      merge_sharded(points[buf].p, point_skip.p, S.P, 3);
      merge_sharded(objects[buf].p, obj_skip.p, S.O, 7);
    }
    if (S.K) CUDA_OK(cudaMemcpyAsync(h_poses.data(), poses[buf].p, h_poses.size() * 8, cudaMemcpyDeviceToHost, stream));
    if (S.P) CUDA_OK(cudaMemcpyAsync(h_points.data(), points[buf].p, h_points.size() * 8, cudaMemcpyDeviceToHost, stream));
    if (S.O) CUDA_OK(cudaMemcpyAsync(h_objects.data(), objects[buf].p, h_objects.size() * 8, cudaMemcpyDeviceToHost, stream));
    CUDA_OK(cudaStreamSynchronize(stream));
    for (int k = 0; k < S.K; k++) if (S.f_of_pose[k] >= 0) std::memcpy(pb.blocks[S.pose_block[k]].host, &h_poses[6 * (size_t)k], 48);
#pragma omp parallel for schedule(static) if (S.P > 32768)
    for (int i = 0; i < S.P; i++) if (!S.point_const[i]) std::memcpy(pb.blocks[S.point_block[i]].host, &h_points[3 * (size_t)i], 24);
    for (int i = 0; i < S.O; i++) if (!S.obj_const[i]) std::memcpy(pb.blocks[S.obj_block[i]].host, &h_objects[7 * (size_t)i], 56);
  }
  void merge_sharded(double* x, const uint8_t* skip, int nblocks, int bs);

  // ---- kernel sequences -------------------------------------------------------------------------------
  static int nblk(int64_t n, int t) { return (int)((n + t - 1) / t); }
  // persistent streaming point kernels: one CTA of kPipeWarps warps per SM, a warp per batch of 4 points
  int pipe_grid() const { return std::max(1, std::min(num_sms, nblk(pr_batch_end - pr_batch_begin, pipe_warps))); }
  EArgs pipe_args() { EArgs a = eargs(pts, J.p); a.batch_begin = pr_batch_begin; a.ne = std::min(a.ne, 4 * pr_batch_end); return a; }
  int row_stages = 2;    // entries per half of the operand ring of schur_rows_kernel (OBVI_ROW_STAGES = 2 / 4 / 8)
  int pipe_warps = 8;    // OBVI_PIPE=16: sixteen single-stage warps per SM instead of eight double-buffered ones
  EArgs eargs(EListDev& D, const double* Jp) {
    EArgs a;
    a.ptr = D.ptr.p; a.pos = D.pos.p; a.f = D.f.p; a.slot = D.slot.p; a.pair_ptr = D.pair_ptr.p; a.nslots = D.nslots.p;
    a.pair_blk = D.pair_blk.p; a.cst = D.cst.p; a.J = Jp; a.escale = D.escale.p; a.einv = D.einv.p; a.eg = D.eg.p;
    a.prior_H = D.has_prior ? D.prior_H.p : nullptr; a.prior_g = D.has_prior ? D.prior_g.p : nullptr;
    a.overflow = D.overflow.p; a.overflow_off = D.overflow_off.p; a.ne = D.ne; a.elist = nullptr;
    return a;
  }
  void fork() { CUDA_OK(cudaEventRecord(ev_fork, stream)); CUDA_OK(cudaStreamWaitEvent(s2, ev_fork, 0)); }
  void join() { CUDA_OK(cudaEventRecord(ev_join, s2)); CUDA_OK(cudaStreamWaitEvent(stream, ev_join, 0)); }
  void zero_scalars(int first, int count) { CUDA_OK(cudaMemsetAsync(scalars.p + first, 0, count * sizeof(double), stream)); }

  // residuals + Jacobians at the current point
  // defer_side: the caller runs build_reduced () next, which takes the side streams back (the loop's case).  The side kernels
  // (bounding boxes, priors, rel-pose, |x|) are placed BEHIND the Jacobian kernel -- beside it they took SMs from the
  // bandwidth-bound kernel (230 us in situ against 184 alone) -- and then run under the point elimination, whose persistent
  // CTAs leave most of an SM's threads and registers free.
  bool side_pending = false;
  void linearize(int apply_loss, bool defer_side = false) {
    const Structure& S = st;
    const size_t pt0 = prof.begin(stream);
    stage0_local = world > 1;
    zero_scalars(SC_COST, 3);
    if (S.K * S.C > 0) { pose_cam_kernel<<<nblk((int64_t)S.K * S.C, 128), 128, 0, stream>>>(poses[cur].p, S.K, cams.p, S.C, 1, pcam.p); launches++; }
    prof.end("lin: pose_cam", pt0, stream);
    const size_t pt1 = prof.begin(stream);
    if (S.n_obs) launch_jacobian(apply_loss, points[cur].p);
    prof.end("lin: jacobian", pt1, stream);
    fork();
    CUDA_OK(cudaStreamWaitEvent(s3, ev_fork, 0));
    if (S.n_bbox) { bbox_kernel<<<nblk(S.n_bbox, 64), 64, 0, s2>>>(bbox.p, S.n_bbox, pcam.p, S.C, objects[cur].p, 0, apply_loss, Jb.p, scalars.p); launches++; }
    if (S.n_unary) { launch_unary(0, apply_loss, cur, s3); }
    if (S.n_rel) { relpose_kernel<<<nblk(S.n_rel, 64), 64, 0, s3>>>(rel.p, S.n_rel, 0, apply_loss, poses[cur].p, rel_out.p, S_upper, gp, hpp_diag, dpose.p, scalars.p); launches++; }
    if (S.K) { xnorm_kernel<<<nblk((int64_t)S.K * 6, 256), 256, 0, s3>>>(poses[cur].p, pose_skip.p, S.K, 6, scalars.p); launches++; }
    if (S.P) { xnorm_kernel<<<nblk((int64_t)S.P * 3, 256), 256, 0, s3>>>(points[cur].p, point_skip.p, S.P, 3, scalars.p); launches++; }
    if (S.O) { xnorm_kernel<<<nblk((int64_t)S.O * 7, 256), 256, 0, s3>>>(objects[cur].p, obj_skip.p, S.O, 7, scalars.p); launches++; }
    const size_t pt2 = prof.begin(stream);
    CUDA_OK(cudaEventRecord(ev_join3, s3));
    // unary factors on points are re-run on the MAIN stream at the start of build_reduced (): no deferral then
    if (defer_side && !(S.n_unary && pts.has_prior)) { side_pending = true; return; }
    CUDA_OK(cudaStreamWaitEvent(stream, ev_join3, 0));
    join();
    prof.end("lin: side join", pt2, stream);
  }
  // The reprojection Jacobian-evaluation kernel: chunks to their point-major positions + the pose-side sums into pose_acc
  // (zeroed here).  Fused TMA-staged kernel by default; OBVI_JAC=plain (or more than 16 calibration classes / 256 cameras)
  // selects the plain kernel followed by the gathering pose accumulation.
  void launch_jacobian(int apply_loss, const double* pts_dev) {
    const Structure& S = st;
    const bool fused_ok = (int)S.classes.size() <= kJacMaxCls && S.C <= 256;
    const int ntiles = nblk(S.n_obs, kJacThreads);
    pose_acc.zero(stream);
    if (jac_mode >= 1 && fused_ok) {
      reproj_jac_fused_kernel<<<ntiles, kJacThreads, kJacSmemBytes, stream>>>(obs.p, S.n_obs, pcam.p, S.C, classes.p, (int)S.classes.size(), pts_dev, apply_loss,
                                                                              jac_tile.p, f_of_pose.p, J.p, pose_acc.p, scalars.p);
      launches++;
    } else {
      reproj_jac_kernel<<<ntiles, kJacThreads, 0, stream>>>(obs.p, S.n_obs, pcam.p, S.C, classes.p, pts_dev, apply_loss, J.p, scalars.p);
      if (S.nf) pose_accum_gather_kernel<<<S.K, kPoseAccThreads, 0, stream>>>(obs.p, J.p, pose_ptr.p, f_of_pose.p, pose_acc.p);
      launches += 2;
    }
  }
  void launch_unary(int mode, int apply_loss, int buf, cudaStream_t strm) {
    const Structure& S = st;
    unary_kernel<<<nblk(S.n_unary, 64), 64, 0, strm>>>(unary.p, S.n_unary, mode, apply_loss, poses[buf].p, points[buf].p, objects[buf].p,
                                                         unary_out.p, f_of_pose.p, su_ptr.p, S_upper, gp, hpp_diag, pts.prior_H.p,
                                                         pts.prior_g.p, objs.prior_H.p, objs.prior_g.p, dpose.p, pts.delta.p,
                                                         objs.delta.p, scalars.p);
    launches++;
  }
  // Schur complement + rhs for the given radius (J fixed)
  void build_reduced(LMParams& lm) {
    const Structure& S = st;
    lm.inv_radius = 1.0 / lm.radius;
    size_t pt0 = prof.begin(stream);
    redbuf.zero(stream);
    zero_scalars(SC_GMAX, 2);  // gmax + fail
    zero_scalars(SC_BT_FAIL, 1);
    if (pts.has_prior) { pts.prior_H.zero(stream); pts.prior_g.zero(stream); }
    if (objs.has_prior) { objs.prior_H.zero(stream); objs.prior_g.zero(stream); }
    // unary factors on points feed the point elimination: keep them on the main stream in that case
    if (S.n_unary && pts.has_prior) launch_unary(1, 1, cur, stream);
    fork();
    if (side_pending) CUDA_OK(cudaStreamWaitEvent(s2, ev_join3, 0));   // s2 reads what linearize ()'s kernels on s3 wrote (unary_out, rel_out)
    // side stream: priors, rel-pose and the object elimination (all accumulate with atomics); enqueued first, high priority
    if (S.n_unary && !pts.has_prior) launch_unary(1, 1, cur, s2);
    if (S.n_rel) { relpose_kernel<<<nblk(S.n_rel, 64), 64, 0, s2>>>(rel.p, S.n_rel, 1, 1, poses[cur].p, rel_out.p, S_upper, gp, hpp_diag, dpose.p, scalars.p); launches++; }
    if (S.O && obj_split) {
      // split object elimination: a warp per object for H_e^-1 / g_e, then the low-register slot / pair kernel
      schur_eblock_kernel<7, 4, 32, 64, true, true><<<S.O, 32, 0, s2>>>(eargs(objs, Jb.p), lm, su_ptr.p, S_upper, gp, hpp_diag, b_schur, scalars.p);
      obj_schur_kernel<kObjThreads, kObjMaxSlots, kObjMinBlocks><<<S.O, kObjThreads, 0, s2>>>(eargs(objs, Jb.p), su_ptr.p, S_upper, gp, hpp_diag, b_schur);
      launches += 2;
    } else if (S.O) { schur_eblock_kernel<7, 4, 128, 64, true><<<S.O, 128, 0, s2>>>(eargs(objs, Jb.p), lm, su_ptr.p, S_upper, gp, hpp_diag, b_schur, scalars.p); launches++; }
    prof.end("zero", pt0, stream); pt0 = prof.begin(stream);
    if (S.n_obs && S.nf) {   // pose-side sums of the reprojection blocks (formed by the Jacobian kernel, radius-independent)
      pose_acc_add_kernel<<<nblk((int64_t)S.nf * 42, 256), 256, 0, stream>>>(S.nf, pose_acc.p, su_ptr.p, S_upper, gp, hpp_diag);
      launches++;
    }
    prof.end("pose_accum", pt0, stream); pt0 = prof.begin(stream);
    {
      if (S.P && !(debug_skip & 1)) {
        if (pipe_warps == 16) point_prep_kernel<16, 1><<<pipe_grid(), 32 * 16, pipe_smem(16, 1), stream>>>(pipe_args(), pr_grp_ptr.p, reinterpret_cast<const uint4*>(pr_grp.p), pr_regular.p, lm, WZ.p, scalars.p);
        else point_prep_kernel<8, 2><<<pipe_grid(), 32 * 8, pipe_smem(8, 2), stream>>>(pipe_args(), pr_grp_ptr.p, reinterpret_cast<const uint4*>(pr_grp.p), pr_regular.p, lm, WZ.p, scalars.p);
        launches++;
      }
      prof.end("point_prep", pt0, stream); pt0 = prof.begin(stream);
      if (n_row_items && !(debug_skip & 2)) {
        const uint4* items = reinterpret_cast<const uint4*>(pr_items.p);
        const int g = nblk(n_row_items, kRowWarps), t = 32 * kRowWarps;
        // measured on C3 (in situ): 477 / 495 / 686 us for H = 2 / 4 / 8 -- resident warps count for more than ring depth
        if (row_stages == 4) schur_rows_kernel<4><<<g, t, rows_smem_bytes(4), stream>>>(items, n_row_items, pr_ent.p, WZ.p, pr_rowblk.p, kRowSpan, S_upper, b_schur);
        else if (row_stages == 8) schur_rows_kernel<8><<<g, t, rows_smem_bytes(8), stream>>>(items, n_row_items, pr_ent.p, WZ.p, pr_rowblk.p, kRowSpan, S_upper, b_schur);
        else schur_rows_kernel<2><<<g, t, rows_smem_bytes(2), stream>>>(items, n_row_items, pr_ent.p, WZ.p, pr_rowblk.p, kRowSpan, S_upper, b_schur);
        launches++;
      }
      if (n_row_fallback) {
        EArgs a = eargs(pts, J.p); a.elist = pr_fallback.p;
        schur_eblock_kernel<3, 2, 32, 16, false><<<n_row_fallback, 32, 0, stream>>>(a, lm, su_ptr.p, S_upper, gp, hpp_diag, b_schur, scalars.p);
        launches++;
      }
    }
    prof.end("schur_points", pt0, stream); pt0 = prof.begin(stream);
    if (side_pending) { CUDA_OK(cudaStreamWaitEvent(stream, ev_join3, 0)); side_pending = false; }   // s3 of a deferred linearize ()
    join();
    prof.end("join(objects,rel)", pt0, stream); pt0 = prof.begin(stream);
    if (world > 1) {
      // ONE collective per build: the partial reduced system and, in its tail, the scalars of this linearisation
      pack_stage0_kernel<<<1, 32, 0, stream>>>(scalars.p, red_tail, rank, world, stage0_local ? 1 : 0);
      allreduce_sum(redbuf.p, (size_t)S.n_upper * 36 + 3 * (size_t)S.nf * 6 + kStage0Sums + 2 * (size_t)world);
      unpack_stage0_kernel<<<1, 32, 0, stream>>>(scalars.p, red_tail, world);
      launches += 2;
      stage0_local = false;
      prof.end("allreduce S", pt0, stream); pt0 = prof.begin(stream);
    }
    if (S.nf) {
      if (lm.compute_scale) { pose_scale_kernel<<<nblk((int64_t)S.nf * 6, 256), 256, 0, stream>>>(hpp_diag, S.nf * 6, pscale.p); launches++; }
      finish_kernel<<<S.nf, 128, 0, stream>>>(S.nf, sf_ptr.p, sf_col.p, sf_src.p, S_upper, pscale.p, hpp_diag, gp, b_schur, lm, Sf.p, rhs.p, scalars.p);
      launches++;
      if (!use_bt) { minv_kernel<<<nblk(S.nf, 64), 64, 0, stream>>>(S.nf, sf_ptr.p, sf_col.p, Sf.p, Minv.p, scalars.p); launches++; }
      prof.end("finish", pt0, stream); pt0 = prof.begin(stream);
      // redo the factorisation when the radius moved by more than kRefactorRatio since it was computed, or when the last PCG
      // needed clearly more iterations than the first one after that factorisation did (bt_fresh_iters: what a FRESH
      // preconditioner achieves on this problem -- an absolute threshold refactors every iteration on graphs whose fresh count
      // already sits at it)
      const bool worn = bt_fresh_iters > 0 ? last_pcg_iters >= bt_fresh_iters + kRefactorPcgIters : last_pcg_iters > 3 * kRefactorPcgIters;
      const bool stale = bt_radius <= 0.0 || lm.radius > kRefactorRatio * bt_radius || lm.radius * kRefactorRatio < bt_radius || worn;
      if (stale) { factor_bt(); bt_radius = lm.radius; bt_factorizations++; bt_fresh_iters = 0; bt_fresh_pending = true; prof.end("factor_bt", pt0, stream); }
    }
  }
  void solve_reduced(const obvi_solver_options& o, bool force_jacobi = false) {
    const Structure& S = st;
    if (!S.nf) return;
    int nf = S.nf, max_iter = o.pcg_max_iterations;
    double tol = o.pcg_relative_tolerance;
    if (use_bt && !force_jacobi && resident_ok && pcg_ll) {
      BtApply P; P.nsb = nsb; P.nlev = nlev; P.Dinv = bt_Dinv.p; P.GaT = bt_GaT.p; P.GcT = bt_GcT.p; P.w = bt_w.p; P.z = bt_z.p;
      // tags of one launch: reductions use tag0 + 1 .. tag0 + 3 max_iter + 3, applications and p fewer
      const unsigned int span = 4u * (unsigned int)std::max(max_iter, 1) + 16u;
      if (ll_tag > 0xFFFFFFFFu - 2u * span) { ll_buf.zero(stream); ll_tag = 0; }
      LLSync Y; Y.red = ll_buf.p; Y.u = Y.red + 4 * (size_t)nsb; Y.z = Y.u + (size_t)nsb * kLLSlots * kB; Y.p = Y.z + (size_t)nsb * kB; Y.tag0 = ll_tag;
      ll_tag += span;
      const uint32_t *a1 = sf_ptr.p, *a2 = sf_col.p;
      const double *a3 = Sf.p, *a4 = rhs.p;
      double *a6 = y.p, *a14 = scalars.p;
      void* args[] = {&nf, &a1, &a2, &a3, &a4, &P, &Y, &a6, &max_iter, &tol, &a14};
      CUDA_OK(cudaLaunchCooperativeKernel((void*)pcg_bt_ll_kernel, dim3(nsb), dim3(kPcgThreads), args, kResidentSmem, stream));
      launches++;
      share_solution();
      return;
    }
    if (use_bt && !force_jacobi && resident_ok) {
      BtApply P; P.nsb = nsb; P.nlev = nlev; P.Dinv = bt_Dinv.p; P.GaT = bt_GaT.p; P.GcT = bt_GcT.p; P.w = bt_w.p; P.z = bt_z.p;
      ResidentSync Y; Y.red_val = rs_dbl.p; Y.u = rs_dbl.p + 16; Y.red_cnt = rs_u32.p; Y.cnt_w = rs_u32.p + 4; Y.ready_z = rs_u32.p + 4 + nsb;
      rs_dbl.zero(stream); rs_u32.zero(stream);
      const uint32_t *a1 = sf_ptr.p, *a2 = sf_col.p;
      const double *a3 = Sf.p, *a4 = rhs.p;
      double *a6 = y.p, *a7 = cg_r.p, *a9 = cg_p.p, *a14 = scalars.p;
      void* args[] = {&nf, &a1, &a2, &a3, &a4, &P, &Y, &a6, &a7, &a9, &max_iter, &tol, &a14};
      CUDA_OK(cudaLaunchCooperativeKernel((void*)pcg_bt_resident_kernel, dim3(nsb), dim3(kPcgThreads), args, kResidentSmem, stream));
      launches++;
      share_solution();
      return;
    }
    if (use_bt && !force_jacobi) {
      BtApply P; P.nsb = nsb; P.nlev = nlev; P.Dinv = bt_Dinv.p; P.GaT = bt_GaT.p; P.GcT = bt_GcT.p; P.w = bt_w.p; P.z = bt_z.p;
      int grid = std::min(num_sms * pcg_bt_blocks_per_sm, std::max(1, nblk(nf, kPcgThreads / 32)));
      const uint32_t *a1 = sf_ptr.p, *a2 = sf_col.p;
      const double *a3 = Sf.p, *a4 = rhs.p;
      double *a6 = y.p, *a7 = cg_r.p, *a9 = cg_p.p, *a10 = cg_q.p, *a11 = cg_acc.p, *a14 = scalars.p;
      void* args[] = {&nf, &a1, &a2, &a3, &a4, &P, &a6, &a7, &a9, &a10, &a11, &max_iter, &tol, &a14};
      CUDA_OK(cudaLaunchCooperativeKernel((void*)pcg_bt_kernel, dim3(grid), dim3(kPcgThreads), args, 0, stream));
      launches++;
      share_solution();
      return;
    }
    int grid = std::min(num_sms * pcg_blocks_per_sm, std::max(1, nblk(nf, kPcgThreads / 32)));
    const uint32_t *a1 = sf_ptr.p, *a2 = sf_col.p;
    const double *a3 = Sf.p, *a4 = rhs.p, *a5 = Minv.p;
    double *a6 = y.p, *a7 = cg_r.p, *a8 = cg_z.p, *a9 = cg_p.p, *a10 = cg_q.p, *a11 = cg_acc.p, *a14 = scalars.p;
    void* args[] = {&nf, &a1, &a2, &a3, &a4, &a5, &a6, &a7, &a8, &a9, &a10, &a11, &max_iter, &tol, &a14};
    CUDA_OK(cudaLaunchCooperativeKernel((void*)pcg_kernel, dim3(grid), dim3(kPcgThreads), args, 0, stream));
    launches++;
    share_solution();
  }
  // Sharded run: every rank solved the (identical) reduced system; rank 0's solution AND its solver status travel in one
  // broadcast, so that every host decision that depends on them -- fallback to block-Jacobi, invalid step, refactorisation
  // -- is the same on every rank (the PCG sums with atomics: iteration counts may differ from rank to rank).
  void share_solution() {
    if (world <= 1) return;
    const size_t nf6 = (size_t)st.nf * 6;
    CUDA_OK(cudaMemcpyAsync(y.p + nf6, scalars.p + SC_PCG_IT, kPcgStatus * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    broadcast0(y.p, nf6 + kPcgStatus);
    CUDA_OK(cudaMemcpyAsync(scalars.p + SC_PCG_IT, y.p + nf6, kPcgStatus * sizeof(double), cudaMemcpyDeviceToDevice, stream));
  }
  // step, model cost change, candidate point
  void take_step() {
    const Structure& S = st;
    const int cand = 1 - cur;
    zero_scalars(SC_MODEL, 2);
    if (S.nf) { pose_step_kernel<<<nblk((int64_t)S.nf * 6, 256), 256, 0, stream>>>(S.nf, pose_of_f.p, pscale.p, y.p, poses[cur].p, poses[cand].p, dpose.p, rank == 0, scalars.p); launches++; }
    fork();
    if (S.P) {
      if (!(debug_skip & 4)) {
        if (pipe_warps == 16) backsub_rows_kernel<16, 1><<<pipe_grid(), 32 * 16, pipe_smem(16, 1), stream>>>(pipe_args(), pr_grp_ptr.p, reinterpret_cast<const uint4*>(pr_grp.p), pr_regular.p, dpose.p, points[cur].p, points[cand].p, pts.delta.p, scalars.p);
        else backsub_rows_kernel<8, 2><<<pipe_grid(), 32 * 8, pipe_smem(8, 2), stream>>>(pipe_args(), pr_grp_ptr.p, reinterpret_cast<const uint4*>(pr_grp.p), pr_regular.p, dpose.p, points[cur].p, points[cand].p, pts.delta.p, scalars.p);
      }
      launches++;
      if (n_row_fallback) {   // points outside the row-owner path: generic kernel on the fallback list
        EArgs a = eargs(pts, J.p); a.elist = pr_fallback.p;
        backsub_points_kernel<<<nblk(n_row_fallback, 4), 128, 0, stream>>>(a, n_row_fallback, dpose.p, points[cur].p, points[cand].p, pts.delta.p, scalars.p);
        launches++;
      }
    }
    if (S.O) { backsub_eblock_kernel<7, 4, 128><<<S.O, 128, 0, s2>>>(eargs(objs, Jb.p), dpose.p, objects[cur].p, objects[cand].p, objs.delta.p, scalars.p); launches++; }
    if (S.n_rel) { relpose_kernel<<<nblk(S.n_rel, 64), 64, 0, s2>>>(rel.p, S.n_rel, 2, 1, poses[cur].p, rel_out.p, S_upper, gp, hpp_diag, dpose.p, scalars.p); launches++; }
    join();
    if (S.n_unary) launch_unary(2, 1, cur, stream);
  }
  void candidate_cost() {
    const Structure& S = st;
    const int cand = 1 - cur;
    zero_scalars(SC_CAND, 2);
    if (S.K * S.C > 0) { pose_cam_kernel<<<nblk((int64_t)S.K * S.C, 128), 128, 0, stream>>>(poses[cand].p, S.K, cams.p, S.C, 0, pcam_cand.p); launches++; }
    fork();
    if (S.n_obs) { reproj_cost_kernel<<<nblk(S.n_obs, kJacThreads * kCostPer), kJacThreads, 0, stream>>>(obs.p, S.n_obs, pcam_cand.p, S.C, classes.p, points[cand].p, scalars.p); launches++; }
    CUDA_OK(cudaStreamWaitEvent(s3, ev_fork, 0));
    if (S.n_bbox) { bbox_kernel<<<nblk(S.n_bbox, 64), 64, 0, s2>>>(bbox.p, S.n_bbox, pcam_cand.p, S.C, objects[cand].p, 1, 1, Jb.p, scalars.p); launches++; }
    if (S.n_unary) launch_unary(3, 1, cand, s3);
    if (S.n_rel) { relpose_kernel<<<nblk(S.n_rel, 64), 64, 0, s3>>>(rel.p, S.n_rel, 3, 1, poses[cand].p, rel_out.p, S_upper, gp, hpp_diag, dpose.p, scalars.p); launches++; }
    CUDA_OK(cudaEventRecord(ev_join3, s3)); CUDA_OK(cudaStreamWaitEvent(stream, ev_join3, 0));
    join();
  }
  // stage 0: after linearize () + build_reduced (); stage 1: after take_step () + candidate_cost ()
  void reduce_scalars(int stage) {
    if (world <= 1) return;
    // replicated pose contributions (|x|^2, |delta|^2) are added by rank 0 only
    if (stage == 0) {
      // normally folded into the all-reduce of the reduced system (build_reduced); only an evaluation without a build
      // (obvi_evaluate) still has rank-local values here
      if (stage0_local) { allreduce_sum(scalars.p + SC_COST, 3); allreduce_max(scalars.p + SC_GMAX, 2); stage0_local = false; }
    } else {
      allreduce_sum(scalars.p + SC_CAND, 4);   // candidate cost, its fixed part, model cost change, |delta|^2
    }
  }
  void fetch_scalars(int stage, bool reduce = true) {
    if (reduce) reduce_scalars(stage);
    CUDA_OK(cudaMemcpyAsync(h_scalars, scalars.p, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CUDA_OK(cudaStreamSynchronize(stream));
    CUDA_OK(cudaGetLastError());
  }
  void allreduce_sum(double* p, size_t n) { comm->allreduce_sum(p, n, stream, rank); }
  void allreduce_max(double* p, size_t n) { comm->allreduce_max(p, n, stream, rank); }
  void broadcast0(double* p, size_t n) { comm->broadcast0(p, n, stream, rank); }

  // Remove a reprojection / bbox block without rebuilding the structure: its record is flagged, the Jacobian kernels then
  // write an all-zero block and no cost, so every downstream kernel sees a factor that contributes nothing
  // (reference: offline_problem_runner.h:752-801 + the Problem edit in object_pose_graph_optimizer.h:991-1155).
  bool mask_factor(int type, uint64_t i) {
    if (!uploaded || pb.dirty) return false;
    const bool rp = type == OBVI_FACTOR_REPROJECTION;
    std::vector<uint32_t>& inv = rp ? inv_rp : inv_bb;
    if (i >= inv.size()) return false;
    const uint32_t q = inv[i];
    if (q != 0xFFFFFFFFu) {
      if (rp) { st.obs[q].flags |= kObsMasked; pend_rp.push_back(q); n_masked_rp++; }
      else { st.bbox[q].flags |= kObsMasked; pend_bb.push_back(q); n_masked_bb++; }
    }
    counts_stale = true;
    return true;
  }
  void flush_masks() {
    Structure& S = st;
    auto push = [&](const std::vector<uint32_t>& idx, auto kernel_call) {
      if (idx.empty()) return;
      DBuf<uint32_t> d; d.upload(idx, stream);
      kernel_call(d.p, (int64_t)idx.size());
      CUDA_OK(cudaStreamSynchronize(stream));   // `d` is freed on return
    };
    // e-blocks that lost every observation leave the reduced program (Ceres drops unused parameter blocks): no |x| share
    std::vector<uint32_t> dead_pts, dead_objs;
    auto all_masked_pt = [&](uint32_t e) { for (uint32_t d = S.pts.ptr[e]; d < S.pts.ptr[e + 1]; d++) if (!(S.obs[S.pts.pos[d]].flags & kObsMasked)) return false; return true; };
    auto all_masked_ob = [&](uint32_t e) { for (uint32_t d = S.objs.ptr[e]; d < S.objs.ptr[e + 1]; d++) if (!(S.bbox[S.objs.pos[d]].flags & kObsMasked)) return false; return true; };
    for (uint32_t q : pend_rp) { const uint32_t e = S.obs[q].point; if (!S.point_const[e] && all_masked_pt(e)) dead_pts.push_back(e); }
    for (uint32_t q : pend_bb) {
      const uint32_t e = S.bbox[q].obj;
      bool has_unary = false;
      for (const UnaryRec& u : S.unary) if (u.kind == 2 && (uint32_t)u.idx == e) { has_unary = true; break; }
      if (!S.obj_const[e] && !has_unary && all_masked_ob(e)) dead_objs.push_back(e);
    }
    push(pend_rp, [&](const uint32_t* d, int64_t n) { or_flag_kernel<<<nblk(n, 256), 256, 0, stream>>>(reinterpret_cast<uint32_t*>(obs.p), d, n, (int)(sizeof(ObsRec) / 4), 7, kObsMasked); });
    push(pend_bb, [&](const uint32_t* d, int64_t n) { or_flag_kernel<<<nblk(n, 256), 256, 0, stream>>>(reinterpret_cast<uint32_t*>(bbox.p), d, n, (int)(sizeof(BBoxRec) / 4), (int)(offsetof(BBoxRec, flags) / 4), kObsMasked); });
    push(dead_pts, [&](const uint32_t* d, int64_t n) { set_bytes_kernel<<<nblk(n, 256), 256, 0, stream>>>(point_skip.p, d, n, 1); });
    push(dead_objs, [&](const uint32_t* d, int64_t n) { set_bytes_kernel<<<nblk(n, 256), 256, 0, stream>>>(obj_skip.p, d, n, 1); });
    pend_rp.clear(); pend_bb.clear();
    if (counts_stale) { count_reduced(pb, st); counts_stale = false; }
  }

  void ensure_structure(double* preprocess_seconds) {
    if (!stream) throw std::runtime_error("host-only problem handle (cuda_device = -1): no CUDA device attached, and this backend has no CPU fallback");
    const auto t0 = std::chrono::steady_clock::now();
    if (pb.dirty || !uploaded) {
      std::string err;
      if (!build_structure(pb, st, rank, world, err)) throw std::runtime_error(err);
      upload_structure();
      structure_builds++;
      pb.dirty = false;
    } else if (counts_stale || !pend_rp.empty() || !pend_bb.empty()) {
      flush_masks();
    }
    if (preprocess_seconds) *preprocess_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }

  int solve(const obvi_solver_options& o, obvi_summary* sum, obvi_iteration_summary* its, int cap);
  int object_covariances(int64_t n_pairs, double* const* obj_a, double* const* obj_b, double* out, std::string& err);
};

__global__ void mask_blocks_kernel(double* x, const uint8_t* skip, int64_t n, int bs) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n * bs && skip[i / bs]) x[i] = 0.0;
}
void Solver::merge_sharded(double* x, const uint8_t* skip, int nblocks, int bs) {
  if (!nblocks) return;
  // constant blocks are skipped by every rank: they are never written back, so zeros are harmless
  mask_blocks_kernel<<<nblk((int64_t)nblocks * bs, 256), 256, 0, stream>>>(x, skip, nblocks, bs);
  allreduce_sum(x, (size_t)nblocks * bs);
}

// Marginal covariance blocks of ellipsoids (ceres::Covariance::Compute + GetCovarianceBlock on the LTM-extraction problem,
// long_term_object_map_extraction.cpp:362-440).  See obj_z_kernel for the algebra.
int Solver::object_covariances(int64_t n_pairs, double* const* obj_a, double* const* obj_b, double* out, std::string& err) {
  if (world > 1) { err = "obvi_object_covariances is single-rank only"; return OBVI_ERR_INVALID_ARGUMENT; }
  ensure_structure(nullptr);
  const Structure& S = st;
  std::unordered_map<int32_t, int> obj_of_block;
  for (int o = 0; o < S.O; o++) obj_of_block[S.obj_block[o]] = o;
  std::vector<int> ia(n_pairs), ib(n_pairs);
  for (int64_t i = 0; i < n_pairs; i++) {
    const int32_t ba = pb.find_block(obj_a[i]), bb = pb.find_block(obj_b[i]);
    auto fa = obj_of_block.find(ba), fb = obj_of_block.find(bb);
    if (ba < 0 || bb < 0 || fa == obj_of_block.end() || fb == obj_of_block.end()) { err = "covariance requested for a block that is not an ellipsoid of this problem"; return OBVI_ERR_NOT_FOUND; }
    ia[i] = fa->second; ib[i] = fb->second;
  }
  gather_params();
  LMParams lm;
  lm.radius = std::numeric_limits<double>::infinity(); lm.inv_radius = 0.0; lm.min_diag = 1e-6; lm.max_diag = 1e32; lm.compute_scale = 1;
  bt_radius = -1.0; last_pcg_iters = 0; bt_fresh_iters = 0; bt_fresh_pending = false;
  linearize(1);
  build_reduced(lm);
  fetch_scalars(0);
  if (h_scalars[SC_FAIL] != 0.0) { err = "covariance: a point / object block of J^T J is singular (rank-deficient Jacobian)"; return OBVI_ERR_NUMERIC; }
  // per (object, slot) tables
  const int64_t n_os = S.objs.slot_ptr[S.O];
  std::vector<uint32_t> so(n_os), sd0(n_os), scnt(n_os, 0);
  for (int o = 0; o < S.O; o++) {
    for (uint32_t d = S.objs.ptr[o]; d < S.objs.ptr[o + 1]; d++) {
      if (S.objs.slot[d] == 0xFFFF) continue;
      const uint32_t g = S.objs.slot_ptr[o] + S.objs.slot[d];
      if (scnt[g] == 0) { sd0[g] = d; so[g] = (uint32_t)o; }
      scnt[g]++;
    }
  }
  DBuf<uint32_t> d_so, d_sd0, d_scnt;
  DBuf<int32_t> d_sf;
  DBuf<double> Zo, Y, d_out;
  d_so.upload(so, stream); d_sd0.upload(sd0, stream); d_scnt.upload(scnt, stream); d_sf.upload(S.objs.slot_f, stream);
  Zo.alloc((size_t)std::max<int64_t>(n_os, 1) * 42);
  if (n_os) { obj_z_kernel<<<nblk(n_os, 64), 64, 0, stream>>>(n_os, d_so.p, d_sd0.p, d_scnt.p, objs.pos.p, Jb.p, objs.einv.p, Zo.p); launches++; }
  const size_t nf6 = (size_t)S.nf * 6;
  Y.alloc(7 * std::max<size_t>(nf6, 1));
  d_out.alloc((size_t)n_pairs * 49); d_out.zero(stream);
  obvi_solver_options so_opts; obvi_solver_options_init(&so_opts);
  // group the requested pairs by their second object: 7 reduced solves per distinct one
  std::map<int, std::vector<int64_t>> by_b;
  for (int64_t i = 0; i < n_pairs; i++) if (!S.obj_const[ia[i]] && !S.obj_const[ib[i]]) by_b[ib[i]].push_back(i);
  int rc = OBVI_OK;
  for (auto& kv : by_b) {
    const int o = kv.first;
    const uint32_t s0 = S.objs.slot_ptr[o], ns = S.objs.slot_ptr[o + 1] - s0;
    Y.zero(stream);
    if (S.nf && ns) {
      for (int c = 0; c < 7 && rc == OBVI_OK; c++) {
        rhs.zero(stream);
        cov_rhs_kernel<<<nblk(ns * 6, 64), 64, 0, stream>>>(s0, ns, c, d_sf.p, Zo.p, pscale.p, rhs.p); launches++;
        solve_reduced(so_opts);
        CUDA_OK(cudaMemcpyAsync(Y.p + (size_t)c * nf6, y.p, nf6 * 8, cudaMemcpyDeviceToDevice, stream));
        CUDA_OK(cudaMemcpyAsync(h_scalars, scalars.p, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CUDA_OK(cudaStreamSynchronize(stream));
        // an all-zero column (bb == 0) is solved trivially; otherwise PCG must have converged on an SPD system
        if (h_scalars[SC_PCG_BREAK] != 0.0 || !(h_scalars[SC_PCG_RES] <= 1e-6)) { err = "covariance: the reduced camera system is not positive definite (rank-deficient Jacobian, e.g. no gauge fix)"; rc = OBVI_ERR_NUMERIC; }
      }
    }
    if (rc != OBVI_OK) break;
    std::vector<CovPair> cp;
    for (int64_t i : kv.second) {
      const int a = ia[i];
      CovPair P; P.s0 = S.objs.slot_ptr[a]; P.ns = S.nf ? S.objs.slot_ptr[a + 1] - P.s0 : 0u; P.obj_a = (uint32_t)a; P.same = a == o; P.out = (uint64_t)i * 49;
      cp.push_back(P);
    }
    DBuf<CovPair> d_cp; d_cp.upload(cp, stream);
    cov_dot_kernel<<<(int)cp.size(), 64, 0, stream>>>(d_cp.p, (int)cp.size(), d_sf.p, Zo.p, pscale.p, Y.p, std::max<size_t>(nf6, 1), objs.einv.p, d_out.p); launches++;
    CUDA_OK(cudaStreamSynchronize(stream));
  }
  if (rc == OBVI_OK && n_pairs) CUDA_OK(cudaMemcpy(out, d_out.p, (size_t)n_pairs * 49 * 8, cudaMemcpyDeviceToHost));
  return rc;
}

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int Solver::solve(const obvi_solver_options& o, obvi_summary* sum, obvi_iteration_summary* its, int cap) {
  const double t_start = now_s();
  std::memset(sum, 0, sizeof(*sum));
  launches = 0;
  double prep = 0;
  ensure_structure(&prep);
  const Structure& S = st;
  sum->preprocessor_time_in_seconds = prep;
  sum->num_parameter_blocks_reduced = S.num_param_blocks_reduced;
  sum->num_parameters_reduced = S.num_params_reduced;
  sum->num_residual_blocks_reduced = (int32_t)S.num_residual_blocks_reduced;
  sum->num_residuals_reduced = (int32_t)S.num_residuals_reduced;
  gather_params();
  bt_radius = -1.0; last_pcg_iters = 0; bt_fresh_iters = 0; bt_fresh_pending = false;

  int n_log = 0;
  int user_stop = 0;   // first non-zero return of the iteration callback: 1 abort, 2 terminate successfully
  auto push = [&](int iter, bool valid, bool ok, int lin_it, double cost, double cc, double gmax, double sn, double rd, double radius) {
    obvi_iteration_summary s;
    s.iteration = iter; s.step_is_valid = valid; s.step_is_successful = ok; s.linear_solver_iterations = lin_it;
    s.cost = cost; s.cost_change = cc; s.gradient_max_norm = gmax; s.step_norm = sn; s.relative_decrease = rd; s.trust_region_radius = radius;
    if (its && n_log < cap) its[n_log] = s;
    n_log++;
    if (o.iteration_callback && !user_stop) {
      // `cur` holds the iterate this summary describes (an accepted step is logged once its linearisation has been read back)
      if (o.update_state_every_iteration && world == 1 && (ok || iter == 0)) scatter_params(cur);
      const int r = o.iteration_callback(o.iteration_callback_user, &s);
      if (r == 1 || r == 2) user_stop = r;
    }
  };
  auto user_termination = [&]() { return user_stop == 2 ? OBVI_USER_SUCCESS : OBVI_USER_FAILURE; };
  float ms = 0;
  double t_jac = 0, t_lin = 0, t_res = 0;
  LMParams lm;
  lm.radius = o.initial_trust_region_radius; lm.min_diag = o.min_lm_diagonal; lm.max_diag = o.max_lm_diagonal; lm.compute_scale = 1;
  lm.inv_radius = 1.0 / lm.radius;

  CUDA_OK(cudaEventRecord(ev[6], stream));
  // ---- iteration 0
  CUDA_OK(cudaEventRecord(ev[0], stream));
  linearize(1, true);
  CUDA_OK(cudaEventRecord(ev[1], stream));
  build_reduced(lm);
  CUDA_OK(cudaEventRecord(ev[2], stream));
  fetch_scalars(0);
  lm.compute_scale = 0;
  CUDA_OK(cudaEventElapsedTime(&ms, ev[0], ev[1])); t_jac += ms * 1e-3;
  CUDA_OK(cudaEventElapsedTime(&ms, ev[1], ev[2])); t_lin += ms * 1e-3;
  double x_cost = h_scalars[SC_COST];
  const double fixed_cost = h_scalars[SC_FIXED];
  double gmax = h_scalars[SC_GMAX];
  double x_norm = std::sqrt(h_scalars[SC_XNORM2]);
  bool build_failed = h_scalars[SC_FAIL] != 0.0;
  sum->initial_cost = x_cost + fixed_cost; sum->fixed_cost = fixed_cost;
  push(0, false, false, 0, x_cost + fixed_cost, 0, gmax, 0, 0, lm.radius);
  double minimum_cost = x_cost;
  int best = cur;  // buffer index holding the minimum-cost iterate (buffer 2 once it diverges from `cur`)
  int termination = OBVI_NO_CONVERGENCE;
  const int max_nonmono = o.use_nonmonotonic_steps ? o.max_consecutive_nonmonotonic_steps : 0;
  double ev_min = x_cost, ev_cur = x_cost, ev_ref = x_cost, ev_cand = x_cost, acc_ref = 0, acc_cand = 0;
  int n_nonmono = 0, n_invalid = 0, iter = 0, lm_steps = 0, n_ok = 0, n_bad = 0;
  double decrease = 2.0;
  bool need_build = false;
  int64_t pcg_total = 0;
  // an accepted step whose new cost / gradient norm / |x| have not been read back yet (see defer_sync)
  struct { int iter, pcg_it; double cost_change, step_norm, rho, radius; } pend = {0, 0, 0, 0, 0, 0};
  bool pending0 = false;
  // needs h_scalars from a fetch made after that step's build_reduced (); returns true when the gradient tolerance is met
  auto finish_pending = [&]() {
    CUDA_OK(cudaEventElapsedTime(&ms, ev[0], ev[1])); t_jac += ms * 1e-3;
    CUDA_OK(cudaEventElapsedTime(&ms, ev[1], ev[5])); t_lin += ms * 1e-3;
    x_cost = h_scalars[SC_COST];
    gmax = h_scalars[SC_GMAX];
    x_norm = std::sqrt(h_scalars[SC_XNORM2]);
    build_failed = h_scalars[SC_FAIL] != 0.0;
    if (x_cost < minimum_cost) { minimum_cost = x_cost; best = cur; }
    push(pend.iter, true, true, pend.pcg_it, x_cost + fixed_cost, pend.cost_change, gmax, pend.step_norm, pend.rho, pend.radius);
    pending0 = false;
    return gmax <= o.gradient_tolerance || user_stop != 0;
  };
  const bool finite0 = std::isfinite(x_cost);
  if (!finite0) termination = OBVI_FAILURE;
  else if (S.num_params_reduced == 0 || gmax <= o.gradient_tolerance) termination = OBVI_CONVERGENCE;
  else while (true) {
    if (user_stop) { termination = user_termination(); break; }
    if (iter >= o.max_num_iterations) { termination = OBVI_NO_CONVERGENCE; break; }
    iter++; lm_steps++;
    CUDA_OK(cudaEventRecord(ev[2], stream));
    if (need_build) { build_reduced(lm); need_build = false; }
    size_t pt0 = prof.begin(stream);
    solve_reduced(o);
    prof.end("pcg", pt0, stream); pt0 = prof.begin(stream);
    take_step();
    prof.end("step+backsub", pt0, stream); pt0 = prof.begin(stream);
    CUDA_OK(cudaEventRecord(ev[3], stream));
    candidate_cost();
    prof.end("candidate_cost", pt0, stream);
    CUDA_OK(cudaEventRecord(ev[4], stream));
    fetch_scalars(1);
    CUDA_OK(cudaEventElapsedTime(&ms, ev[2], ev[3])); t_lin += ms * 1e-3;
    CUDA_OK(cudaEventElapsedTime(&ms, ev[3], ev[4])); t_res += ms * 1e-3;
    if (pending0 && finish_pending()) {
      // the previous (accepted) iterate already met the gradient tolerance (or the callback stopped the solve there): the
      // step enqueued speculatively is dropped
      iter--; lm_steps--;
      termination = user_stop ? user_termination() : OBVI_CONVERGENCE;
      break;
    }
    if (h_scalars[SC_PCG_BREAK] == 2.0) {
      // the block-tridiagonal factorisation hit a non-positive pivot: redo this step with block-Jacobi PCG
      minv_kernel<<<nblk(S.nf, 64), 64, 0, stream>>>(S.nf, sf_ptr.p, sf_col.p, Sf.p, Minv.p, scalars.p); launches++;
      solve_reduced(o, true);
      take_step();
      candidate_cost();
      fetch_scalars(1);
    }
    const int pcg_it = (int)h_scalars[SC_PCG_IT];
    pcg_total += pcg_it;
    last_pcg_iters = pcg_it;
    if (bt_fresh_pending) { bt_fresh_iters = pcg_it; bt_fresh_pending = false; }
    const double model_change = -h_scalars[SC_MODEL];
    bool valid = !build_failed && h_scalars[SC_FAIL] == 0.0 && h_scalars[SC_PCG_BREAK] == 0.0 && std::isfinite(model_change) &&
                 std::isfinite(h_scalars[SC_STEP2]) && model_change > 0.0;
    build_failed = false;
    if (!valid) {
      if (++n_invalid >= o.max_num_consecutive_invalid_steps) { termination = OBVI_FAILURE; break; }
      lm.radius /= decrease; decrease *= 2.0;  // LevenbergMarquardtStrategy::StepIsInvalid () = StepRejected (0): same rule as a rejected step
      need_build = true;
      n_bad++;
      push(iter, false, false, pcg_it, x_cost + fixed_cost, 0, gmax, 0, 0, lm.radius);
      continue;
    }
    n_invalid = 0;
    double cand_cost = h_scalars[SC_CAND];
    if (!std::isfinite(cand_cost)) cand_cost = std::numeric_limits<double>::max();
    const double step_norm = std::sqrt(h_scalars[SC_STEP2]);
    if (step_norm <= o.parameter_tolerance * (x_norm + o.parameter_tolerance)) { termination = OBVI_CONVERGENCE; break; }
    const double cost_change = x_cost - cand_cost;
    if (std::fabs(cost_change) <= o.function_tolerance * x_cost) { termination = OBVI_CONVERGENCE; break; }
    double rho;
    if (cand_cost >= std::numeric_limits<double>::max()) rho = std::numeric_limits<double>::lowest();
    else rho = std::max((ev_cur - cand_cost) / model_change, (ev_ref - cand_cost) / (acc_ref + model_change));
    if (rho > o.min_relative_decrease) {
      // keep the minimum-cost iterate alive in buffer 2 before the buffers rotate
      if (best == cur) { /* current is the best so far: nothing to save yet */ }
      const int old = cur;
      cur = 1 - cur;
      if (best == old) {
        // `old` becomes the candidate buffer of the next step and will be overwritten: save it
        const size_t np = (size_t)S.K * 6 * 8, nq = (size_t)S.P * 3 * 8, no = (size_t)S.O * 7 * 8;
        // on the second side stream: the main stream was synchronised by fetch_scalars (1), so `old` is final; linearize ()
        // below joins s3 back into the main stream before anything can overwrite `old`
        if (np) CUDA_OK(cudaMemcpyAsync(poses[2].p, poses[old].p, np, cudaMemcpyDeviceToDevice, s3));
        if (nq) CUDA_OK(cudaMemcpyAsync(points[2].p, points[old].p, nq, cudaMemcpyDeviceToDevice, s3));
        if (no) CUDA_OK(cudaMemcpyAsync(objects[2].p, objects[old].p, no, cudaMemcpyDeviceToDevice, s3));
        best = 2;
      }
      lm.radius = std::min(o.max_trust_region_radius, lm.radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rho - 1.0, 3)));
      decrease = 2.0;
      CUDA_OK(cudaEventRecord(ev[0], stream));
      linearize(1, true);
      CUDA_OK(cudaEventRecord(ev[1], stream));
      build_reduced(lm);
      CUDA_OK(cudaEventRecord(ev[5], stream));
      ev_cur = cand_cost; acc_cand += model_change; acc_ref += model_change;
      if (ev_cur < ev_min) { ev_min = ev_cur; n_nonmono = 0; ev_cand = ev_cur; acc_cand = 0; }
      else { n_nonmono++; if (ev_cur > ev_cand) { ev_cand = ev_cur; acc_cand = 0; } }
      if (n_nonmono == max_nonmono) { ev_ref = ev_cand; acc_ref = acc_cand; }
      n_ok++;
      pend = {iter, pcg_it, cost_change, step_norm, rho, lm.radius};
      pending0 = true;
      if (defer_sync) {
        // no host synchronisation here: the cost / gradient norm of the new point ride along with the next step's
        // scalars, and the next step is enqueued right away (it is dropped if this point turns out to be converged)
        reduce_scalars(0);
      } else {
        fetch_scalars(0);
        if (finish_pending()) { termination = user_stop ? user_termination() : OBVI_CONVERGENCE; break; }
      }
    } else {
      lm.radius /= decrease; decrease *= 2.0;
      need_build = true;
      n_bad++;
      push(iter, true, false, pcg_it, cand_cost + fixed_cost, cost_change, gmax, step_norm, rho, lm.radius);
    }
    if (lm.radius <= o.min_trust_region_radius) { termination = OBVI_CONVERGENCE; break; }
  }
  if (pending0) {   // the loop ended (iteration limit, minimum radius, failure) with an accepted step still unread
    fetch_scalars(0, false);   // its cross-rank reduction was enqueued when the step was accepted
    if (finish_pending() && termination == OBVI_NO_CONVERGENCE) termination = user_stop ? user_termination() : OBVI_CONVERGENCE;
  }
  if (user_stop && termination != OBVI_FAILURE) termination = user_termination();   // a stop requested on the very last summary
  CUDA_OK(cudaEventRecord(ev[7], stream));
  CUDA_OK(cudaStreamSynchronize(stream));
  CUDA_OK(cudaEventElapsedTime(&ms, ev[6], ev[7]));
  sum->minimizer_device_time_in_seconds = ms * 1e-3;
  prof.report(lm_steps);
  // write the minimum-cost iterate back to the caller's blocks (on FAILURE the initial values stay)
  sum->is_solution_usable = termination == OBVI_CONVERGENCE || termination == OBVI_NO_CONVERGENCE || termination == OBVI_USER_SUCCESS;
  if (sum->is_solution_usable) scatter_params(best);
  sum->termination_type = termination;
  sum->num_iterations = n_log; sum->num_lm_steps = lm_steps; sum->num_successful_steps = n_ok; sum->num_unsuccessful_steps = n_bad;
  sum->final_cost = minimum_cost + fixed_cost;
  sum->linear_solver_time_in_seconds = t_lin; sum->jacobian_evaluation_time_in_seconds = t_jac; sum->residual_evaluation_time_in_seconds = t_res;
  sum->pcg_iterations_total = pcg_total; sum->kernel_launches = launches;
  sum->total_time_in_seconds = now_s() - t_start;
  return OBVI_OK;
}

}  // namespace obvi

// =========================================================================================== C ABI
using namespace obvi;

struct obvi_problem { Solver s; };

#define API_BEGIN try {
#define API_END(p)                                                                   \
  }                                                                                  \
  catch (const std::exception& e) {                                                  \
    if (p) (p)->s.pb.error = e.what(); else g_create_error = e.what();               \
    return OBVI_ERR_CUDA;                                                            \
  }

static int fail(obvi_problem* p, int code, const char* msg) { p->s.pb.error = msg; return code; }

extern "C" {

const char* obvi_version(void) { return "obvi_ba 0.1 (sm_100a, fp64)"; }

int obvi_problem_create(int device, obvi_problem** out) {
  if (!out) return OBVI_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  obvi_problem* p = nullptr;
  try {
    if (device == -1) {  // host-only handle: problem assembly / structure inspection, every compute call fails
      p = new obvi_problem();
      p->s.pb.device = -1;
      *out = p;
      return OBVI_OK;
    }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) { g_create_error = std::string("no usable CUDA device: ") + cudaGetErrorString(e) + " (this backend has no CPU fallback)"; return OBVI_ERR_CUDA; }
    if (device < 0 || device >= n) { g_create_error = "cuda_device out of range"; return OBVI_ERR_INVALID_ARGUMENT; }
    p = new obvi_problem();
    p->s.pb.device = device;
    p->s.init_device();
    *out = p;
    return OBVI_OK;
  } catch (const std::exception& e) {
    g_create_error = e.what();
    delete p;
    return OBVI_ERR_CUDA;
  }
}
void obvi_problem_destroy(obvi_problem* p) { if (p) { if (p->s.pb.device >= 0) cudaSetDevice(p->s.pb.device); delete p; } }
const char* obvi_last_error(const obvi_problem* p) { return p ? p->s.pb.error.c_str() : g_create_error.c_str(); }

int obvi_param_add(obvi_problem* p, double* host, int size) {
  if (!p || !host || (size != 3 && size != 6 && size != 7)) return p ? fail(p, OBVI_ERR_INVALID_ARGUMENT, "parameter block size must be 3, 6 or 7") : OBVI_ERR_INVALID_ARGUMENT;
  const int id = p->s.pb.add_block(host, size);
  if (id == -2) return fail(p, OBVI_ERR_INVALID_ARGUMENT, "parameter block already registered with another size");
  return OBVI_OK;
}
int obvi_param_add_array(obvi_problem* p, double* base, int size, int64_t count) {
  if (!p || !base || count < 0 || (size != 3 && size != 6 && size != 7)) return OBVI_ERR_INVALID_ARGUMENT;
  if (count == 0) return OBVI_OK;
  // every address of the range is checked, not just its ends: a block registered singly in the MIDDLE of the range would
  // otherwise keep its old id in the factors that reference it while later look-ups resolve to the array's id
  if (p->s.pb.array_overlaps_array(base, size, count) || p->s.pb.array_overlaps_single(base, size, count))
    return fail(p, OBVI_ERR_INVALID_ARGUMENT, "array overlaps registered blocks");
  p->s.pb.add_array(base, size, count);
  return OBVI_OK;
}
int obvi_param_remove(obvi_problem* p, double* host) {
  if (!p) return OBVI_ERR_INVALID_ARGUMENT;
  const int id = p->s.pb.find_block(host);
  if (id < 0 || !p->s.pb.blocks[id].alive) return fail(p, OBVI_ERR_NOT_FOUND, "unknown parameter block");
  // Ceres removes the residual blocks that depend on the parameter block as well
  Problem& pb = p->s.pb;
  for (auto& f : pb.reproj) if (f.alive && (f.pose == id || f.point == id)) { f.alive = 0; pb.n_live--; }
  for (auto& f : pb.bbox) if (f.alive && (f.pose == id || f.obj == id)) { f.alive = 0; pb.n_live--; }
  for (auto& f : pb.unary) if (f.alive && f.block == id) { f.alive = 0; pb.n_live--; }
  for (auto& f : pb.rel) if (f.alive && (f.p1 == id || f.p2 == id)) { f.alive = 0; pb.n_live--; }
  pb.blocks[id].alive = 0;
  pb.dirty = true;
  return OBVI_OK;
}
int obvi_param_set_constant(obvi_problem* p, double* host, int c) {
  if (!p) return OBVI_ERR_INVALID_ARGUMENT;
  const int id = p->s.pb.find_block(host);
  if (id < 0 || !p->s.pb.blocks[id].alive) return fail(p, OBVI_ERR_NOT_FOUND, "unknown parameter block");
  if (p->s.pb.blocks[id].constant != (uint8_t)(c != 0)) { p->s.pb.blocks[id].constant = c != 0; p->s.pb.dirty = true; }
  return OBVI_OK;
}
int obvi_param_is_constant(const obvi_problem* p, const double* host, int* out) {
  if (!p || !out) return OBVI_ERR_INVALID_ARGUMENT;
  const int id = p->s.pb.find_block(host);
  if (id < 0 || !p->s.pb.blocks[id].alive) return OBVI_ERR_NOT_FOUND;
  *out = p->s.pb.blocks[id].constant;
  return OBVI_OK;
}

int obvi_camera_add(obvi_problem* p, const double intr[4], const double R[9], const double t[3], int* cam_id) {
  if (!p || !intr || !R || !t || !cam_id) return OBVI_ERR_INVALID_ARGUMENT;
  Camera c;
  std::memcpy(c.intr, intr, sizeof(c.intr));
  invert_extrinsics(R, t, c.Rinv, c.tinv);
  for (size_t i = 0; i < p->s.pb.cams.size(); i++)
    if (std::memcmp(&p->s.pb.cams[i], &c, sizeof(Camera)) == 0) { *cam_id = (int)i; return OBVI_OK; }
  p->s.pb.cams.push_back(c);
  *cam_id = (int)p->s.pb.cams.size() - 1;
  p->s.pb.dirty = true;
  return OBVI_OK;
}

static int add_id(Problem& pb, int type, size_t index, obvi_factor_id* id) {
  const obvi_factor_id v = make_id(type, index);
  pb.order.push_back(v);
  pb.n_live++;
  pb.dirty = true;
  if (id) *id = v;
  return OBVI_OK;
}

int obvi_factor_add_reproj(obvi_problem* p, double* pose, double* point, int cam, const double px[2], double sigma, double huber, obvi_factor_id* id) {
  if (!p || !pose || !point || !px) return OBVI_ERR_INVALID_ARGUMENT;
  Problem& pb = p->s.pb;
  if (cam < 0 || cam >= (int)pb.cams.size()) return fail(p, OBVI_ERR_INVALID_ARGUMENT, "unknown camera id");
  if (!(sigma > 0)) return fail(p, OBVI_ERR_INVALID_ARGUMENT, "reprojection_error_std_dev must be positive");
  const int a = pb.add_block(pose, 6), b = pb.add_block(point, 3);
  if (a < 0 || b < 0) return fail(p, OBVI_ERR_INVALID_ARGUMENT, "parameter block size mismatch");
  pb.reproj.push_back({a, b, cam, 1, px[0], px[1], sigma, huber});
  return add_id(pb, OBVI_FACTOR_REPROJECTION, pb.reproj.size() - 1, id);
}
int obvi_factor_add_reproj_batch(obvi_problem* p, int64_t n, double* const* poses, double* const* points, const int32_t* cams,
                                 const double* px, const double* sig, double huber, obvi_factor_id* ids) {
  if (!p || n < 0 || (n && (!poses || !points || !cams || !px || !sig))) return OBVI_ERR_INVALID_ARGUMENT;
  Problem& pb = p->s.pb;
  pb.reproj.reserve(pb.reproj.size() + n); pb.order.reserve(pb.order.size() + n);
  for (int64_t i = 0; i < n; i++) {
    const int rc = obvi_factor_add_reproj(p, poses[i], points[i], cams[i], px + 2 * i, sig[i], huber, ids ? ids + i : nullptr);
    if (rc != OBVI_OK) return rc;
  }
  return OBVI_OK;
}
int obvi_factor_add_bbox(obvi_problem* p, double* ell, double* pose, int cam, const double corners[4], const double cov[16],
                         double invalid_err, double huber, obvi_factor_id* id) {
  if (!p || !ell || !pose || !corners || !cov) return OBVI_ERR_INVALID_ARGUMENT;
  Problem& pb = p->s.pb;
  if (cam < 0 || cam >= (int)pb.cams.size()) return fail(p, OBVI_ERR_INVALID_ARGUMENT, "unknown camera id");
  const int a = pb.add_block(ell, 7), b = pb.add_block(pose, 6);
  if (a < 0 || b < 0) return fail(p, OBVI_ERR_INVALID_ARGUMENT, "parameter block size mismatch");
  BBoxFactor f;
  f.obj = a; f.pose = b; f.cam = cam; f.alive = 1; f.invalid_err = invalid_err; f.huber = huber;
  const Camera& c = pb.cams[cam];
  double sq[16];
  if (!sqrt_information(cov, 4, sq)) return fail(p, OBVI_ERR_NUMERIC, "bounding-box information matrix has NaN");
  const double sc[4] = {c.intr[0], c.intr[0], c.intr[1], c.intr[1]};
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) f.A4[4 * i + j] = sq[4 * i + j] * sc[j];
  f.brect[0] = (corners[0] - c.intr[2]) / c.intr[0]; f.brect[1] = (corners[1] - c.intr[2]) / c.intr[0];
  f.brect[2] = (corners[2] - c.intr[3]) / c.intr[1]; f.brect[3] = (corners[3] - c.intr[3]) / c.intr[1];
  pb.bbox.push_back(f);
  return add_id(pb, OBVI_FACTOR_BBOX, pb.bbox.size() - 1, id);
}
int obvi_factor_add_bbox_batch(obvi_problem* p, int64_t n, double* const* ells, double* const* poses, const int32_t* cams,
                               const double* corners, const double* covs, double invalid_err, double huber, obvi_factor_id* ids) {
  if (!p || n < 0 || (n && (!ells || !poses || !cams || !corners || !covs))) return OBVI_ERR_INVALID_ARGUMENT;
  for (int64_t i = 0; i < n; i++) {
    const int rc = obvi_factor_add_bbox(p, ells[i], poses[i], cams[i], corners + 4 * i, covs + 16 * i, invalid_err, huber, ids ? ids + i : nullptr);
    if (rc != OBVI_OK) return rc;
  }
  return OBVI_OK;
}
static int add_unary(obvi_problem* p, double* block, int bs, int type, int k, int off, const double* A, const double* mean, double huber, obvi_factor_id* id) {
  Problem& pb = p->s.pb;
  const int a = pb.add_block(block, bs);
  if (a < 0) return fail(p, OBVI_ERR_INVALID_ARGUMENT, "parameter block size mismatch");
  UnaryFactor f; std::memset(&f, 0, sizeof(f));
  f.block = a; f.type = type; f.k = k; f.off = off; f.alive = 1; f.huber = huber;
  std::memcpy(f.A, A, sizeof(double) * k * k); std::memcpy(f.mean, mean, sizeof(double) * k);
  pb.unary.push_back(f);
  return add_id(pb, type, pb.unary.size() - 1, id);
}
int obvi_factor_add_shape_prior(obvi_problem* p, double* ell, const double mean[3], const double cov[9], double huber, obvi_factor_id* id) {
  if (!p || !ell || !mean || !cov) return OBVI_ERR_INVALID_ARGUMENT;
  double A[9];
  if (!sqrt_information(cov, 3, A)) return fail(p, OBVI_ERR_NUMERIC, "shape-prior information matrix has NaN");
  return add_unary(p, ell, 7, OBVI_FACTOR_SHAPE_PRIOR, 3, 4, A, mean, huber, id);
}
int obvi_factor_add_ltm_prior(obvi_problem* p, double* ell, const double mean[7], const double cov[49], double huber, obvi_factor_id* id) {
  if (!p || !ell || !mean || !cov) return OBVI_ERR_INVALID_ARGUMENT;
  double A[49];
  if (!sqrt_information(cov, 7, A)) return fail(p, OBVI_ERR_NUMERIC, "LTM-prior information matrix has NaN");
  return add_unary(p, ell, 7, OBVI_FACTOR_LTM_PRIOR, 7, 0, A, mean, huber, id);
}
int obvi_factor_add_param_prior(obvi_problem* p, double* block, int idx, double mean, double sd, double huber, obvi_factor_id* id) {
  if (!p || !block) return OBVI_ERR_INVALID_ARGUMENT;
  const int b = p->s.pb.find_block(block);
  if (b < 0) return fail(p, OBVI_ERR_NOT_FOUND, "parameter prior on an unknown block (add the block first)");
  const int bs = p->s.pb.blocks[b].size;
  if (idx < 0 || idx >= bs || !(sd > 0)) return fail(p, OBVI_ERR_INVALID_ARGUMENT, "bad parameter prior");
  const double A = 1.0 / sd;
  return add_unary(p, block, bs, OBVI_FACTOR_PARAM_PRIOR, 1, idx, &A, &mean, huber, id);
}
int obvi_factor_add_rel_pose(obvi_problem* p, double* p1, double* p2, const double tm[3], const double Rm[9], const double cov[36], double huber, obvi_factor_id* id) {
  if (!p || !p1 || !p2 || !tm || !Rm || !cov) return OBVI_ERR_INVALID_ARGUMENT;
  Problem& pb = p->s.pb;
  const int a = pb.add_block(p1, 6), b = pb.add_block(p2, 6);
  if (a < 0 || b < 0) return fail(p, OBVI_ERR_INVALID_ARGUMENT, "parameter block size mismatch");
  RelPoseFactor f;
  f.p1 = a; f.p2 = b; f.alive = 1; f.huber = huber;
  std::memcpy(f.tm, tm, sizeof(f.tm));
  inverse3(Rm, f.Rm_inv);
  if (!sqrt_information(cov, 6, f.A6)) return fail(p, OBVI_ERR_NUMERIC, "Relative pose factor had NaN information matrix");
  pb.rel.push_back(f);
  return add_id(pb, OBVI_FACTOR_REL_POSE, pb.rel.size() - 1, id);
}
int obvi_factor_remove(obvi_problem* p, obvi_factor_id id) {
  if (!p) return OBVI_ERR_INVALID_ARGUMENT;
  Problem& pb = p->s.pb;
  const int t = id_type(id); const uint64_t i = id_index(id);
  uint8_t* alive = nullptr;
  if (t == OBVI_FACTOR_REPROJECTION && i < pb.reproj.size()) alive = &pb.reproj[i].alive;
  else if (t == OBVI_FACTOR_BBOX && i < pb.bbox.size()) alive = &pb.bbox[i].alive;
  else if ((t == OBVI_FACTOR_SHAPE_PRIOR || t == OBVI_FACTOR_LTM_PRIOR || t == OBVI_FACTOR_PARAM_PRIOR) && i < pb.unary.size() && pb.unary[i].type == t) alive = &pb.unary[i].alive;
  else if (t == OBVI_FACTOR_REL_POSE && i < pb.rel.size()) alive = &pb.rel[i].alive;
  if (!alive || !*alive) return fail(p, OBVI_ERR_NOT_FOUND, "unknown residual block id");
  *alive = 0; pb.n_live--;
  // reprojection / bbox blocks of an uploaded structure are removed in place; everything else forces a rebuild
  if (!((t == OBVI_FACTOR_REPROJECTION || t == OBVI_FACTOR_BBOX) && p->s.mask_factor(t, i))) pb.dirty = true;
  return OBVI_OK;
}
int obvi_factor_remove_batch(obvi_problem* p, const obvi_factor_id* ids, int64_t n) {
  if (!p || (n > 0 && !ids) || n < 0) return OBVI_ERR_INVALID_ARGUMENT;
  for (int64_t k = 0; k < n; k++) {
    const int rc = obvi_factor_remove(p, ids[k]);
    if (rc != OBVI_OK) return rc;             // blocks before k stay removed, like a loop of single calls
  }
  return OBVI_OK;
}
int64_t obvi_num_factors(const obvi_problem* p) { return p ? p->s.pb.n_live : 0; }
int64_t obvi_num_structure_builds(const obvi_problem* p) { return p ? p->s.structure_builds : 0; }

static bool id_alive(const Problem& pb, obvi_factor_id id, int* size) {
  const int t = id_type(id); const uint64_t i = id_index(id);
  switch (t) {
    case OBVI_FACTOR_REPROJECTION: *size = 2; return pb.reproj[i].alive;
    case OBVI_FACTOR_BBOX: *size = 4; return pb.bbox[i].alive;
    case OBVI_FACTOR_REL_POSE: *size = 6; return pb.rel[i].alive;
    default: *size = pb.unary[i].k; return pb.unary[i].alive;
  }
}
int obvi_residual_blocks(const obvi_problem* p, obvi_factor_id* ids, int32_t* types, int32_t* sizes, int64_t cap, int64_t* n) {
  if (!p || !n) return OBVI_ERR_INVALID_ARGUMENT;
  int64_t c = 0;
  for (obvi_factor_id id : p->s.pb.order) {
    int sz;
    if (!id_alive(p->s.pb, id, &sz)) continue;
    if (c < cap) { if (ids) ids[c] = id; if (types) types[c] = id_type(id); if (sizes) sizes[c] = sz; }
    c++;
  }
  *n = c;
  return OBVI_OK;
}

void obvi_solver_options_init(obvi_solver_options* o) {
  if (!o) return;
  o->max_num_iterations = 50; o->use_nonmonotonic_steps = 0;
  o->function_tolerance = 1e-6; o->gradient_tolerance = 1e-10; o->parameter_tolerance = 1e-8;
  o->initial_trust_region_radius = 1e4; o->max_trust_region_radius = 1e16;
  o->min_trust_region_radius = 1e-32; o->min_relative_decrease = 1e-3; o->min_lm_diagonal = 1e-6; o->max_lm_diagonal = 1e32;
  o->max_consecutive_nonmonotonic_steps = 5; o->max_num_consecutive_invalid_steps = 5;
  o->pcg_max_iterations = 2000; o->pcg_relative_tolerance = 1e-12;
  o->iteration_callback = nullptr; o->iteration_callback_user = nullptr; o->update_state_every_iteration = 0;
}

int obvi_solve(obvi_problem* p, const obvi_solver_options* o, obvi_summary* sum, obvi_iteration_summary* its, int32_t cap) {
  if (!p || !o || !sum) return OBVI_ERR_INVALID_ARGUMENT;
  API_BEGIN
  if (p->s.pb.device < 0) return fail(p, OBVI_ERR_CUDA, "host-only problem handle: no CUDA device attached (no CPU fallback exists)");
  CUDA_OK(cudaSetDevice(p->s.pb.device));
  return p->s.solve(*o, sum, its, cap);
  API_END(p)
}

// evaluate every type at the current host values; results land in the Solver's device buffers
static void evaluate_all(Solver& s, int apply_loss) {
  s.ensure_structure(nullptr);
  s.gather_params();
  s.linearize(apply_loss);
  s.fetch_scalars(0);
}

int obvi_evaluate_factor_type(obvi_problem* p, int type, int apply_loss, double* r, double* J0, double* J1) {
  if (!p) return OBVI_ERR_INVALID_ARGUMENT;
  API_BEGIN
  Solver& s = p->s;
  if (s.pb.device < 0) return fail(p, OBVI_ERR_CUDA, "host-only problem handle: no CUDA device attached (no CPU fallback exists)");
  CUDA_OK(cudaSetDevice(s.pb.device));
  if (s.world > 1) return fail(p, OBVI_ERR_INVALID_ARGUMENT, "obvi_evaluate_factor_type is single-rank only");
  evaluate_all(s, apply_loss);
  const Structure& S = s.st;
  if (type == OBVI_FACTOR_REPROJECTION) {
    std::vector<double> h((size_t)S.n_obs * kChunk);
    if (S.n_obs) CUDA_OK(cudaMemcpy(h.data(), s.J.p, h.size() * 8, cudaMemcpyDeviceToHost));
    // user order = position among live reprojection factors in order of addition
    std::vector<int64_t> live_rank(s.pb.reproj.size(), -1);
    int64_t c = 0;
    for (size_t i = 0; i < s.pb.reproj.size(); i++) if (s.pb.reproj[i].alive) live_rank[i] = c++;
    for (int64_t q = 0; q < S.n_obs; q++) {
      const int64_t u = live_rank[S.obs_user[q]];
      if (u < 0) continue;   // removed in place
      double jp[12], jl[6], rr[2];
      decode_chunk(&h[(size_t)S.obs[q].dst * kChunk], jp, jl, rr);
      if (J0) std::memcpy(J0 + 12 * u, jp, 96);
      if (J1) std::memcpy(J1 + 6 * u, jl, 48);
      if (r) std::memcpy(r + 2 * u, rr, 16);
    }
  } else if (type == OBVI_FACTOR_BBOX) {
    std::vector<double> h((size_t)S.n_bbox * kBBoxChunk);
    if (S.n_bbox) CUDA_OK(cudaMemcpy(h.data(), s.Jb.p, h.size() * 8, cudaMemcpyDeviceToHost));
    std::vector<int64_t> live_rank(s.pb.bbox.size(), -1);
    int64_t c = 0;
    for (size_t i = 0; i < s.pb.bbox.size(); i++) if (s.pb.bbox[i].alive) live_rank[i] = c++;
    for (int64_t q = 0; q < S.n_bbox; q++) {
      const int64_t u = live_rank[S.bbox_user[q]];
      if (u < 0) continue;   // removed in place
      const double* ch = &h[(size_t)q * kBBoxChunk];
      if (J1) std::memcpy(J1 + 24 * u, ch, 192);       // pose
      if (J0) std::memcpy(J0 + 28 * u, ch + 24, 224);  // ellipsoid
      if (r) std::memcpy(r + 4 * u, ch + 52, 32);
    }
  } else if (type == OBVI_FACTOR_REL_POSE) {
    std::vector<RelOut> h(S.n_rel);
    if (S.n_rel) CUDA_OK(cudaMemcpy(h.data(), s.rel_out.p, h.size() * sizeof(RelOut), cudaMemcpyDeviceToHost));
    std::vector<int64_t> live_rank(s.pb.rel.size(), -1);
    int64_t c = 0;
    for (size_t i = 0; i < s.pb.rel.size(); i++) if (s.pb.rel[i].alive) live_rank[i] = c++;
    for (int64_t q = 0; q < S.n_rel; q++) {
      const int64_t u = live_rank[S.rel_user[q]];
      if (r) std::memcpy(r + 6 * u, h[q].r, 48);
      if (J0) std::memcpy(J0 + 36 * u, h[q].J1, 288);
      if (J1) std::memcpy(J1 + 36 * u, h[q].J2, 288);
    }
  } else if (type == OBVI_FACTOR_SHAPE_PRIOR || type == OBVI_FACTOR_LTM_PRIOR || type == OBVI_FACTOR_PARAM_PRIOR) {
    std::vector<UnaryOut> h(S.n_unary);
    if (S.n_unary) CUDA_OK(cudaMemcpy(h.data(), s.unary_out.p, h.size() * sizeof(UnaryOut), cudaMemcpyDeviceToHost));
    std::vector<int64_t> live_rank(s.pb.unary.size(), -1);
    int64_t c = 0;
    for (size_t i = 0; i < s.pb.unary.size(); i++) if (s.pb.unary[i].alive && s.pb.unary[i].type == type) live_rank[i] = c++;
    for (int64_t q = 0; q < S.n_unary; q++) {
      const UnaryFactor& f = s.pb.unary[S.unary_user[q]];
      if (f.type != type) continue;
      const int64_t u = live_rank[S.unary_user[q]];
      const int k = f.k, bs = s.pb.blocks[f.block].size;
      const int ld = type == OBVI_FACTOR_PARAM_PRIOR ? 7 : bs;
      if (r) std::memcpy(r + (size_t)k * u, h[q].r, 8 * k);
      if (J0) {
        double* Jd = J0 + (size_t)k * ld * u;
        for (int a = 0; a < k * ld; a++) Jd[a] = 0.0;
        for (int a = 0; a < k; a++) for (int c2 = 0; c2 < k; c2++) Jd[a * ld + f.off + c2] = h[q].sc * f.A[a * k + c2];
      }
    }
  } else {
    return fail(p, OBVI_ERR_INVALID_ARGUMENT, "unknown factor type");
  }
  return OBVI_OK;
  API_END(p)
}

int obvi_evaluate(obvi_problem* p, int apply_loss, double* cost, double* residuals, int64_t cap, int64_t* nres) {
  if (!p) return OBVI_ERR_INVALID_ARGUMENT;
  API_BEGIN
  Solver& s = p->s;
  if (s.pb.device < 0) return fail(p, OBVI_ERR_CUDA, "host-only problem handle: no CUDA device attached (no CPU fallback exists)");
  CUDA_OK(cudaSetDevice(s.pb.device));
  evaluate_all(s, apply_loss);
  if (cost) *cost = s.h_scalars[SC_COST] + s.h_scalars[SC_FIXED];
  if (!residuals && !nres) return OBVI_OK;
  if (s.world > 1) return fail(p, OBVI_ERR_INVALID_ARGUMENT, "residual export is single-rank only");
  const Structure& S = s.st;
  // per-type residuals in internal order -> per factor index
  std::vector<double> hJ((size_t)S.n_obs * 2), hB((size_t)S.n_bbox * kBBoxChunk);     // hJ: the packed residuals, by chunk position
  std::vector<RelOut> hR(S.n_rel); std::vector<UnaryOut> hU(S.n_unary);
  if (S.n_obs) {
    DBuf<double> packed; packed.alloc((size_t)S.n_obs * 2);
    extract_residuals_kernel<<<Solver::nblk(S.n_obs, 256), 256, 0, s.stream>>>(s.J.p, S.n_obs, reinterpret_cast<double2*>(packed.p));
    CUDA_OK(cudaMemcpyAsync(hJ.data(), packed.p, hJ.size() * 8, cudaMemcpyDeviceToHost, s.stream));
    CUDA_OK(cudaStreamSynchronize(s.stream));
  }
  if (S.n_bbox) CUDA_OK(cudaMemcpy(hB.data(), s.Jb.p, hB.size() * 8, cudaMemcpyDeviceToHost));
  if (S.n_rel) CUDA_OK(cudaMemcpy(hR.data(), s.rel_out.p, hR.size() * sizeof(RelOut), cudaMemcpyDeviceToHost));
  if (S.n_unary) CUDA_OK(cudaMemcpy(hU.data(), s.unary_out.p, hU.size() * sizeof(UnaryOut), cudaMemcpyDeviceToHost));
  std::vector<uint32_t> inv_rp(s.pb.reproj.size(), 0), inv_bb(s.pb.bbox.size(), 0), inv_un(s.pb.unary.size(), 0), inv_rl(s.pb.rel.size(), 0);
  for (int64_t q = 0; q < S.n_obs; q++) inv_rp[S.obs_user[q]] = (uint32_t)q;
  for (int64_t q = 0; q < S.n_bbox; q++) inv_bb[S.bbox_user[q]] = (uint32_t)q;
  for (int64_t q = 0; q < S.n_unary; q++) inv_un[S.unary_user[q]] = (uint32_t)q;
  for (int64_t q = 0; q < S.n_rel; q++) inv_rl[S.rel_user[q]] = (uint32_t)q;
  int64_t w = 0;
  for (obvi_factor_id id : s.pb.order) {
    int sz;
    if (!id_alive(s.pb, id, &sz)) continue;
    const uint64_t i = id_index(id);
    const double* src;
    switch (id_type(id)) {
      case OBVI_FACTOR_REPROJECTION: src = &hJ[(size_t)S.obs[inv_rp[i]].dst * 2]; break;
      case OBVI_FACTOR_BBOX: src = &hB[(size_t)inv_bb[i] * kBBoxChunk + 52]; break;
      case OBVI_FACTOR_REL_POSE: src = hR[inv_rl[i]].r; break;
      default: src = hU[inv_un[i]].r; break;
    }
    if (residuals && w + sz <= cap) std::memcpy(residuals + w, src, 8 * sz);
    w += sz;
  }
  if (nres) *nres = w;
  return OBVI_OK;
  API_END(p)
}

// Problem::Evaluate with gradient / Jacobian output (long_term_object_map_extraction.cpp:251-252,591-598): the Jacobian of the
// listed residual blocks with respect to the listed parameter blocks as a CRS matrix.  Evaluated on the device, assembled here.
int obvi_evaluate_jacobian(obvi_problem* p, int apply_loss, const obvi_factor_id* ids, int64_t n_ids, double* const* blocks,
                           int64_t n_blocks, int64_t* num_rows, int64_t* num_cols, int64_t* nnz, int32_t* crs_rows,
                           int32_t* crs_cols, double* crs_values, double* gradient) {
  if (!p || !num_rows || !num_cols || !nnz) return OBVI_ERR_INVALID_ARGUMENT;
  API_BEGIN
  Solver& s = p->s;
  if (s.pb.device < 0) return fail(p, OBVI_ERR_CUDA, "host-only problem handle: no CUDA device attached (no CPU fallback exists)");
  CUDA_OK(cudaSetDevice(s.pb.device));
  if (s.world > 1) return fail(p, OBVI_ERR_INVALID_ARGUMENT, "Jacobian export is single-rank only");
  evaluate_all(s, apply_loss);
  const Structure& S = s.st;
  const Problem& pb = s.pb;
  // column offset of every variable parameter block (-1: constant / not listed)
  std::vector<int64_t> col_of(pb.blocks.size(), -1);
  int64_t ncol = 0;
  if (blocks) {
    for (int64_t i = 0; i < n_blocks; i++) {
      const int32_t b = pb.find_block(blocks[i]);
      if (b < 0) return fail(p, OBVI_ERR_NOT_FOUND, "Evaluate: unknown parameter block");
      if (pb.blocks[b].constant || col_of[b] >= 0) continue;
      col_of[b] = ncol; ncol += pb.blocks[b].size;
    }
  } else {   // every variable block of the built structure: poses, points, objects in internal order
    for (int32_t b : S.pose_block) if (!pb.blocks[b].constant) { col_of[b] = ncol; ncol += 6; }
    for (int32_t b : S.point_block) if (!pb.blocks[b].constant) { col_of[b] = ncol; ncol += 3; }
    for (int32_t b : S.obj_block) if (!pb.blocks[b].constant) { col_of[b] = ncol; ncol += 7; }
  }
  std::vector<double> hJ((size_t)S.n_obs * kChunk), hB((size_t)S.n_bbox * kBBoxChunk);
  std::vector<RelOut> hR(S.n_rel); std::vector<UnaryOut> hU(S.n_unary);
  if (S.n_obs) CUDA_OK(cudaMemcpy(hJ.data(), s.J.p, hJ.size() * 8, cudaMemcpyDeviceToHost));
  if (S.n_bbox) CUDA_OK(cudaMemcpy(hB.data(), s.Jb.p, hB.size() * 8, cudaMemcpyDeviceToHost));
  if (S.n_rel) CUDA_OK(cudaMemcpy(hR.data(), s.rel_out.p, hR.size() * sizeof(RelOut), cudaMemcpyDeviceToHost));
  if (S.n_unary) CUDA_OK(cudaMemcpy(hU.data(), s.unary_out.p, hU.size() * sizeof(UnaryOut), cudaMemcpyDeviceToHost));
  std::vector<uint32_t> inv_un(pb.unary.size(), 0), inv_rl(pb.rel.size(), 0);
  for (int64_t q = 0; q < S.n_unary; q++) inv_un[S.unary_user[q]] = (uint32_t)q;
  for (int64_t q = 0; q < S.n_rel; q++) inv_rl[S.rel_user[q]] = (uint32_t)q;
  std::vector<obvi_factor_id> all;
  if (!ids) { for (obvi_factor_id id : pb.order) { int sz; if (id_alive(pb, id, &sz)) all.push_back(id); } ids = all.data(); n_ids = (int64_t)all.size(); }
  const bool fill = crs_rows && crs_cols && crs_values;
  if (gradient) for (int64_t c = 0; c < ncol; c++) gradient[c] = 0.0;
  int64_t row = 0, nz = 0;
  if (fill) crs_rows[0] = 0;
  for (int64_t n = 0; n < n_ids; n++) {
    int sz;
    if (!id_alive(pb, ids[n], &sz)) return fail(p, OBVI_ERR_NOT_FOUND, "Evaluate: unknown residual block id");
    const uint64_t i = id_index(ids[n]);
    // up to two parameter blocks per residual block: (block id, width, pointer to the k x width row-major Jacobian, leading dim)
    struct Part { int32_t b; int w; const double* J; int ld; int off; double scale; };
    Part parts[2]; int np = 0;
    const double* r = nullptr;
    double unaryJ[49], rp_jp[12], rp_jl[6], rp_r[2];
    switch (id_type(ids[n])) {
      case OBVI_FACTOR_REPROJECTION: {
        decode_chunk(&hJ[(size_t)S.obs[s.inv_rp[i]].dst * kChunk], rp_jp, rp_jl, rp_r);
        parts[np++] = {pb.reproj[i].pose, 6, rp_jp, 6, 0, 1.0}; parts[np++] = {pb.reproj[i].point, 3, rp_jl, 3, 0, 1.0}; r = rp_r; break; }
      case OBVI_FACTOR_BBOX: {
        const double* ch = &hB[(size_t)s.inv_bb[i] * kBBoxChunk];
        parts[np++] = {pb.bbox[i].obj, 7, ch + 24, 7, 0, 1.0}; parts[np++] = {pb.bbox[i].pose, 6, ch, 6, 0, 1.0}; r = ch + 52; break; }
      case OBVI_FACTOR_REL_POSE: {
        const RelOut& o = hR[inv_rl[i]];
        parts[np++] = {pb.rel[i].p1, 6, o.J1, 6, 0, 1.0}; parts[np++] = {pb.rel[i].p2, 6, o.J2, 6, 0, 1.0}; r = o.r; break; }
      default: {
        const UnaryFactor& f = pb.unary[i];
        const UnaryOut& o = hU[inv_un[i]];
        for (int a = 0; a < f.k * f.k; a++) unaryJ[a] = o.sc * f.A[a];
        parts[np++] = {f.block, f.k, unaryJ, f.k, f.off, 1.0}; r = o.r; break; }
    }
    // CRS rows list their columns in ascending order
    if (np == 2 && col_of[parts[0].b] > col_of[parts[1].b]) std::swap(parts[0], parts[1]);
    for (int a = 0; a < sz; a++) {
      for (int q = 0; q < np; q++) {
        const int64_t c0 = col_of[parts[q].b];
        if (c0 < 0) continue;
        if (np == 2 && q == 1 && parts[1].b == parts[0].b) continue;   // same block twice: not produced by the reference's factors
        for (int c = 0; c < parts[q].w; c++) {
          const double v = parts[q].J[a * parts[q].ld + c];
          if (fill) { crs_cols[nz] = (int32_t)(c0 + parts[q].off + c); crs_values[nz] = v; }
          if (gradient) gradient[c0 + parts[q].off + c] += v * r[a];
          nz++;
        }
      }
      row++;
      if (fill) crs_rows[row] = (int32_t)nz;
    }
  }
  *num_rows = row; *num_cols = ncol; *nnz = nz;
  return OBVI_OK;
  API_END(p)
}

// Two-phase outlier rejection support (offline_problem_runner.h:752-801): per block sum r^2 from the raw
// residuals, inserted into std::map<double, id, std::greater<double>> (equal keys overwrite), then the first
// (size_t)(map.size() * fraction) entries.
int obvi_topk_outliers(obvi_problem* p, int type, double fraction, obvi_factor_id* ids, int64_t cap, int64_t* n) {
  if (!p || !n) return OBVI_ERR_INVALID_ARGUMENT;
  API_BEGIN
  Solver& s = p->s;
  if (s.pb.device < 0) return fail(p, OBVI_ERR_CUDA, "host-only problem handle: no CUDA device attached (no CPU fallback exists)");
  CUDA_OK(cudaSetDevice(s.pb.device));
  if ((type == OBVI_FACTOR_REPROJECTION || type == OBVI_FACTOR_BBOX) && s.world == 1) {
    // ---- device path: raw residuals -> per-block keys -> stable radix sort (descending) -> last of every run of equal
    //      keys (std::map semantics: equal keys overwrite, the last inserted block survives) -> first floor(n_u * fraction)
    evaluate_all(s, 0);
    const Structure& S = s.st;
    const bool rp = type == OBVI_FACTOR_REPROJECTION;
    const int64_t nb = rp ? S.n_obs : S.n_bbox;
    if (nb == 0) { *n = 0; return OBVI_OK; }
    // rank of every internal block among the live blocks of its type, in order of addition (= std::map insertion order)
    std::vector<uint32_t> rank(nb);
    std::vector<obvi_factor_id> id_of_rank(nb);
    {
      const size_t total = rp ? s.pb.reproj.size() : s.pb.bbox.size();
      std::vector<int64_t> live_rank(total, -1);
      int64_t c = 0;
      for (size_t i = 0; i < total; i++) if (rp ? s.pb.reproj[i].alive : s.pb.bbox[i].alive) { live_rank[i] = c; id_of_rank[c] = make_id(type, i); c++; }
      // blocks removed in place: parked behind the live ones (positions c, c + 1, ...) with the top bit set
      int64_t dead = c;
      for (int64_t q = 0; q < nb; q++) {
        const int64_t lr = live_rank[rp ? S.obs_user[q] : S.bbox_user[q]];
        // indexed by CHUNK position (reprojection chunks are point-major; bbox chunks are in record order)
        rank[rp ? S.obs[q].dst : (uint32_t)q] = lr >= 0 ? (uint32_t)lr : (0x80000000u | (uint32_t)(dead++));
      }
    }
    const int64_t n_dead = rp ? s.n_masked_rp : s.n_masked_bb;
    DBuf<uint32_t> d_rank, d_vals, d_vals_sorted, d_sel;
    DBuf<double> d_keys, d_keys_sorted;
    DBuf<uint8_t> d_flags, d_tmp;
    DBuf<int64_t> d_count;
    d_rank.upload(rank, s.stream);
    d_vals.alloc(nb); d_vals_sorted.alloc(nb); d_sel.alloc(nb); d_keys.alloc(nb); d_keys_sorted.alloc(nb); d_flags.alloc(nb); d_count.alloc(1);
    if (rp) block_sqnorm_kernel<<<Solver::nblk(nb, 256), 256, 0, s.stream>>>(s.J.p, nb, kChunk, kChunkR, 2, d_rank.p, d_keys.p, d_vals.p);
    else block_sqnorm_kernel<<<Solver::nblk(nb, 256), 256, 0, s.stream>>>(s.Jb.p, nb, kBBoxChunk, 52, 4, d_rank.p, d_keys.p, d_vals.p);
    size_t tmp1 = 0, tmp2 = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp1, d_keys.p, d_keys_sorted.p, d_vals.p, d_vals_sorted.p, (int)nb, 0, 64, s.stream);
    cub::DeviceSelect::Flagged(nullptr, tmp2, d_vals_sorted.p, d_flags.p, d_sel.p, d_count.p, (int)nb, s.stream);
    d_tmp.alloc(std::max(tmp1, tmp2));
    cub::DeviceRadixSort::SortPairsDescending(d_tmp.p, tmp1, d_keys.p, d_keys_sorted.p, d_vals.p, d_vals_sorted.p, (int)nb, 0, 64, s.stream);
    run_end_flags_kernel<<<Solver::nblk(nb, 256), 256, 0, s.stream>>>(d_keys_sorted.p, nb, d_flags.p);
    cub::DeviceSelect::Flagged(d_tmp.p, tmp2, d_vals_sorted.p, d_flags.p, d_sel.p, d_count.p, (int)nb, s.stream);
    int64_t n_unique = 0;
    CUDA_OK(cudaMemcpyAsync(&n_unique, d_count.p, sizeof(int64_t), cudaMemcpyDeviceToHost, s.stream));
    CUDA_OK(cudaStreamSynchronize(s.stream));
    if (n_dead > 0) n_unique -= 1;   // the run of -1 keys at the very end of the descending order
    const size_t k = (size_t)(n_unique * fraction);
    std::vector<uint32_t> sel(k);
    if (k) CUDA_OK(cudaMemcpy(sel.data(), d_sel.p, k * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    for (size_t c = 0; c < k && (int64_t)c < cap; c++) if (ids) ids[c] = id_of_rank[sel[c]];
    *n = (int64_t)k;
    return OBVI_OK;
  }
  // ---- other factor types (a handful of blocks): host ranking with the reference's std::map
  int64_t nres = 0;
  int rc = obvi_evaluate(p, 0, nullptr, nullptr, 0, &nres);
  if (rc != OBVI_OK) return rc;
  std::vector<double> res(nres);
  rc = obvi_evaluate(p, 0, nullptr, res.data(), nres, &nres);
  if (rc != OBVI_OK) return rc;
  std::map<double, obvi_factor_id, std::greater<double>> by_err;
  int64_t w = 0;
  for (obvi_factor_id id : s.pb.order) {
    int sz;
    if (!id_alive(s.pb, id, &sz)) continue;
    if (id_type(id) == type) {
      double e = 0;
      for (int a = 0; a < sz; a++) e += res[w + a] * res[w + a];
      by_err[e] = id;
    }
    w += sz;
  }
  const size_t k = (size_t)(by_err.size() * fraction);
  int64_t c = 0;
  for (auto it = by_err.begin(); it != by_err.end() && (size_t)c < k; ++it, ++c) if (ids && c < cap) ids[c] = it->second;
  *n = (int64_t)k;
  return OBVI_OK;
  API_END(p)
}

int obvi_object_covariances(obvi_problem* p, int64_t n_pairs, double* const* obj_a, double* const* obj_b, double* out) {
  if (!p || n_pairs < 0 || (n_pairs && (!obj_a || !obj_b || !out))) return OBVI_ERR_INVALID_ARGUMENT;
  API_BEGIN
  Solver& s = p->s;
  if (s.pb.device < 0) return fail(p, OBVI_ERR_CUDA, "host-only problem handle: no CUDA device attached (no CPU fallback exists)");
  CUDA_OK(cudaSetDevice(s.pb.device));
  std::string err;
  const int rc = s.object_covariances(n_pairs, obj_a, obj_b, out, err);
  if (rc != OBVI_OK) return fail(p, rc, err.c_str());
  return OBVI_OK;
  API_END(p)
}

int obvi_profile_jacobian(obvi_problem* p, int reps, double* sec, int64_t* bytes, int64_t* nobs) {
  if (!p || reps < 1 || !sec) return OBVI_ERR_INVALID_ARGUMENT;
  API_BEGIN
  Solver& s = p->s;
  if (s.pb.device < 0) return fail(p, OBVI_ERR_CUDA, "host-only problem handle: no CUDA device attached (no CPU fallback exists)");
  CUDA_OK(cudaSetDevice(s.pb.device));
  s.ensure_structure(nullptr);
  s.gather_params();
  const Structure& S = s.st;
  if (!S.n_obs) return fail(p, OBVI_ERR_INVALID_ARGUMENT, "no reprojection observations");
  pose_cam_kernel<<<Solver::nblk((int64_t)S.K * S.C, 128), 128, 0, s.stream>>>(s.poses[0].p, S.K, s.cams.p, S.C, 1, s.pcam.p);
  auto launch = [&]() { s.launch_jacobian(1, s.points[0].p); };
  for (int i = 0; i < 3; i++) launch();
  CUDA_OK(cudaEventRecord(s.ev[0], s.stream));
  for (int i = 0; i < reps; i++) launch();
  CUDA_OK(cudaEventRecord(s.ev[1], s.stream));
  CUDA_OK(cudaStreamSynchronize(s.stream));
  CUDA_OK(cudaGetLastError());
  float ms = 0;
  CUDA_OK(cudaEventElapsedTime(&ms, s.ev[0], s.ev[1]));
  *sec = ms * 1e-3 / reps;
  // unique parameter blocks read by one launch
  int64_t npts = 0;
  for (int i = 0; i < S.P; i++) npts += S.pts.ptr[i + 1] > S.pts.ptr[i];
  if (bytes) *bytes = S.n_obs * 192 + (int64_t)S.K * 48 + npts * 24;
  if (nobs) *nobs = S.n_obs;
  return OBVI_OK;
  API_END(p)
}

// Host-only inspection of the structure build for a given (rank, world): how the graph is sharded.
// stats: [0] observations on this rank, [1] bbox observations, [2] unary factors, [3] rel-pose factors,
// [4] variable poses, [5] upper blocks of the reduced matrix, [6] points with observations on this rank,
// [7] objects with observations on this rank, [8] point batches, [9] sum of sf_col (structure checksum),
// [10] reduced parameters, [11] reduced residual blocks.
int obvi_debug_partition(obvi_problem* p, int rank, int world, int64_t* stats) {
  if (!p || !stats || world < 1 || rank < 0 || rank >= world) return OBVI_ERR_INVALID_ARGUMENT;
  try {
    Structure S;
    std::string err;
    if (!build_structure(p->s.pb, S, rank, world, err)) return fail(p, OBVI_ERR_INVALID_ARGUMENT, err.c_str());
    int64_t np = 0, no = 0, ck = 0;
    for (int i = 0; i < S.P; i++) np += S.pts.ptr[i + 1] > S.pts.ptr[i];
    for (int i = 0; i < S.O; i++) no += S.objs.ptr[i + 1] > S.objs.ptr[i];
    for (uint32_t c : S.sf_col) ck += c;
    const int64_t v[12] = {S.n_obs, S.n_bbox, S.n_unary, S.n_rel, S.nf, S.n_upper, np, no, (int64_t)S.prow.items.size(), ck,
                           S.num_params_reduced, S.num_residual_blocks_reduced};
    std::memcpy(stats, v, sizeof(v));
    return OBVI_OK;
  } catch (const std::exception& e) {
    p->s.pb.error = e.what();
    return OBVI_ERR_INVALID_ARGUMENT;
  }
}

int obvi_debug_row_products(obvi_problem* p, int rank, int world, int64_t* stats) {
  if (!p || !stats || world < 1 || rank < 0 || rank >= world) return OBVI_ERR_INVALID_ARGUMENT;
  try {
    Structure S;
    std::string err;
    if (!build_structure(p->s.pb, S, rank, world, err)) return fail(p, OBVI_ERR_INVALID_ARGUMENT, err.c_str());
    const Structure::PointRows& R = S.prow;
    int64_t products = 0, in_items = 0, regular = 0, max_item = 0;
    for (uint32_t e : R.ent) {
      const int nl = (int)((e >> 26) & 7u), va = (int)((e >> 29) & 1u), vb = (int)((e >> 30) & 1u);
      products += (va ? std::min(nl, 5) : 0) + (vb ? nl - 1 : 0);      // what schur_rows_kernel's switch accumulates with a non-zero A operand
    }
    for (const Structure::RowItem& it : R.items) { in_items += it.cnt; max_item = std::max<int64_t>(max_item, it.cnt); }
    for (uint8_t r : R.regular) regular += r;
    const int64_t v[8] = {(int64_t)R.ent.size(), (int64_t)R.items.size(), products, R.n_slots, regular, (int64_t)R.fallback.size(), in_items, max_item};
    std::memcpy(stats, v, sizeof(v));
    return OBVI_OK;
  } catch (const std::exception& e) {
    p->s.pb.error = e.what();
    return OBVI_ERR_INVALID_ARGUMENT;
  }
}

int obvi_debug_structure_hash(obvi_problem* p, int rank, int world, uint64_t* hash) {
  if (!p || !hash || world < 1 || rank < 0 || rank >= world) return OBVI_ERR_INVALID_ARGUMENT;
  try {
    Structure S;
    std::string err;
    if (!build_structure(p->s.pb, S, rank, world, err)) return fail(p, OBVI_ERR_INVALID_ARGUMENT, err.c_str());
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const void* data, size_t bytes) {
      const unsigned char* c = static_cast<const unsigned char*>(data);
      for (size_t i = 0; i < bytes; i++) { h ^= c[i]; h *= 1099511628211ull; }
      h ^= bytes; h *= 1099511628211ull;
    };
    auto vec = [&](const auto& v) { if (!v.empty()) mix(v.data(), v.size() * sizeof(v[0])); else mix(nullptr, 0); };
    const int64_t scal[] = {S.K, S.P, S.O, S.C, S.nf, S.n_obs, S.n_bbox, S.n_unary, S.n_rel, S.num_residual_blocks_reduced, S.num_residuals_reduced,
                            S.num_param_blocks_reduced, S.num_params_reduced, S.n_upper, S.pts.max_slots, S.objs.max_slots, S.prow.n_slots};
    mix(scal, sizeof(scal));
    vec(S.pose_block); vec(S.point_block); vec(S.obj_block); vec(S.f_of_pose); vec(S.point_const); vec(S.obj_const);
    vec(S.obs); vec(S.obs_user); vec(S.pose_ptr);
    for (const CalibClass& c : S.classes) { mix(&c.mx, 8); mix(&c.my, 8); mix(&c.huber, 8); mix(&c.cam, 4); }
    for (const Structure::EList* L : {&S.pts, &S.objs}) {
      vec(L->ptr); vec(L->pos); vec(L->f); vec(L->slot); vec(L->pair_ptr); vec(L->nslots); vec(L->pair_blk); vec(L->slot_f); vec(L->slot_ptr);
    }
    vec(S.prow.grp_ptr); vec(S.prow.grp); vec(S.prow.grp_f); vec(S.prow.regular); vec(S.prow.ent); vec(S.prow.items); vec(S.prow.rowblk); vec(S.prow.fallback);
    for (const BBoxRec& r : S.bbox) { mix(r.brect, sizeof(r.brect)); mix(r.A4, sizeof(r.A4)); mix(&r.invalid_err, 8); mix(&r.huber, 8); mix(&r.obj, 16); }
    vec(S.bbox_user);
    for (const UnaryRec& r : S.unary) { mix(r.A, sizeof(r.A)); mix(r.mean, sizeof(r.mean)); mix(&r.huber, 8); mix(&r.kind, 16); mix(&r.flags, 4); }
    vec(S.unary_user);
    for (const RelRec& r : S.rel) { mix(r.tm, sizeof(r.tm)); mix(r.Rm_inv, sizeof(r.Rm_inv)); mix(r.A6, sizeof(r.A6)); mix(&r.huber, 8); mix(&r.p1, 32); }
    vec(S.rel_user);
    vec(S.su_ptr); vec(S.su_col); vec(S.sf_ptr); vec(S.sf_col); vec(S.sf_src);
    *hash = h;
    return OBVI_OK;
  } catch (const std::exception& e) {
    p->s.pb.error = e.what();
    return OBVI_ERR_INVALID_ARGUMENT;
  }
}

int obvi_comm_unique_id(void* out) {
  if (!out) return OBVI_ERR_INVALID_ARGUMENT;
  std::string err;
  if (!g_nccl.load(err)) { g_create_error = err; return OBVI_ERR_COMM; }
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) { g_create_error = "ncclGetUniqueId failed"; return OBVI_ERR_COMM; }
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
  std::memcpy(out, &id, 128);
  return OBVI_OK;
}
int obvi_comm_init(obvi_problem* p, const void* uid, int rank, int world) {
  if (!p || !uid || world < 1 || rank < 0 || rank >= world) return OBVI_ERR_INVALID_ARGUMENT;
  API_BEGIN
  Solver& s = p->s;
  if (s.pb.device < 0) return fail(p, OBVI_ERR_CUDA, "host-only problem handle: no CUDA device attached (no CPU fallback exists)");
  CUDA_OK(cudaSetDevice(s.pb.device));
  std::string err;
  if (!g_nccl.load(err)) return fail(p, OBVI_ERR_COMM, err.c_str());
  ncclUniqueId id;
  std::memcpy(&id, uid, 128);
  auto nc = std::make_shared<NcclComm>();
  ncclResult_t r = g_nccl.CommInitRank(&nc->comm, world, id, rank);
  if (r != ncclSuccess) return fail(p, OBVI_ERR_COMM, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "ncclCommInitRank failed");
  s.comm = nc; s.rank = rank; s.world = world; s.pb.dirty = true;
  return OBVI_OK;
  API_END(p)
}

// Share the communicator of `src` (set up with obvi_comm_init / obvi_comm_init_local) with another problem handle of the same
// process and device: a schedule solves hundreds of problems one after the other, a communicator is created once.
int obvi_comm_attach(obvi_problem* p, const obvi_problem* src) {
  if (!p || !src || !src->s.comm) return OBVI_ERR_INVALID_ARGUMENT;
  if (p->s.pb.device != src->s.pb.device) return fail(p, OBVI_ERR_INVALID_ARGUMENT, "obvi_comm_attach: the two problems live on different devices");
  p->s.comm = src->s.comm; p->s.rank = src->s.rank; p->s.world = src->s.world; p->s.pb.dirty = true;
  return OBVI_OK;
}

// Join `world` problem handles of THIS process into one sharded solve (handle i becomes rank i).  Every handle must then
// be driven by its own host thread, all of them making the same sequence of collective calls (obvi_solve, obvi_evaluate).
int obvi_comm_init_local(obvi_problem** handles, int world) {
  if (!handles || world < 1) return OBVI_ERR_INVALID_ARGUMENT;
  for (int r = 0; r < world; r++) if (!handles[r] || handles[r]->s.pb.device < 0) return OBVI_ERR_INVALID_ARGUMENT;
  auto lc = std::make_shared<LocalComm>(world);
  for (int r = 0; r < world; r++) { Solver& s = handles[r]->s; s.comm = lc; s.rank = r; s.world = world; s.pb.dirty = true; }
  return OBVI_OK;
}

}  // extern "C"
