// mvr_util.cu -- error reporting shared by the C-ABI entry points.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <omp.h>
#include <pthread.h>
#include <vector>

#include "mvr_common.cuh"

namespace mvr {

thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// Launch-time errors only (bad configuration, missing kernel image): never synchronises.
int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return 0;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return (int)e;
}

// ---- launch accounting / per-kernel device timing (bench.py's roofline leg) ---------------------
static const int kMaxProfEvents = 8192;
static std::atomic<long long> g_launches{0};
static std::mutex g_prof_mu;
static char g_prof_name[64] = "";
static std::vector<cudaEvent_t> g_prof_events;   // start/stop pairs, created lazily and reused
static int g_prof_used = 0;                        // number of events recorded (2 per launch)

void prof_begin(const char* name, cudaStream_t st) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (!g_prof_name[0]) return;
  if (name[0] == '(') ++name;      // MVR_LAUNCH((kernel<a, b>), ...)
  if (strncmp(name, "mvr::", 5) == 0) name += 5;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (strncmp(name, g_prof_name, strlen(g_prof_name)) != 0 || g_prof_used + 2 > kMaxProfEvents) return;   // prefix: template arguments ignored
  while ((int)g_prof_events.size() < g_prof_used + 2) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    g_prof_events.push_back(e);
  }
  cudaEventRecord(g_prof_events[g_prof_used], st);
}
void prof_end(const char* name, cudaStream_t st) {
  if (!g_prof_name[0]) return;
  if (name[0] == '(') ++name;
  if (strncmp(name, "mvr::", 5) == 0) name += 5;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (strncmp(name, g_prof_name, strlen(g_prof_name)) != 0 || g_prof_used + 2 > (int)g_prof_events.size()) return;
  cudaEventRecord(g_prof_events[g_prof_used + 1], st);
  g_prof_used += 2;
}

bool pdl_enabled(cudaStream_t st) {
  static const bool on = [] { const char* e = getenv("MVR_PDL"); return !(e && e[0] == '0'); }();      // on unless MVR_PDL=0
  if (!on) return false;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); return false; }
  return cs == cudaStreamCaptureStatusNone;
}

}  // namespace mvr

extern "C" long long mvr_launch_count(void) { return mvr::g_launches.load(); }

extern "C" int mvr_profile_enable(const char* kernel_name) {
  std::lock_guard<std::mutex> lk(mvr::g_prof_mu);
  mvr::g_prof_used = 0;
  if (!kernel_name) { mvr::g_prof_name[0] = 0; return 0; }
  strncpy(mvr::g_prof_name, kernel_name, sizeof(mvr::g_prof_name) - 1);
  mvr::g_prof_name[sizeof(mvr::g_prof_name) - 1] = 0;
  return 0;
}

extern "C" int mvr_profile_collect(double* total_ms, int* n_launches) {
  std::lock_guard<std::mutex> lk(mvr::g_prof_mu);
  double tot = 0.0;
  int n = 0;
  for (int i = 0; i + 1 < mvr::g_prof_used; i += 2) {
    cudaError_t e = cudaEventSynchronize(mvr::g_prof_events[i + 1]);
    float ms = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, mvr::g_prof_events[i], mvr::g_prof_events[i + 1]);
    if (e != cudaSuccess) { mvr::set_error("mvr_profile_collect: %s", cudaGetErrorString(e)); return (int)e; }
    tot += ms; ++n;
  }
  if (total_ms) *total_ms = tot;
  if (n_launches) *n_launches = n;
  mvr::g_prof_used = 0;
  mvr::g_prof_name[0] = 0;
  return 0;
}

extern "C" int mvr_abi_version(void) { return MVR_ABI_VERSION; }
extern "C" const char* mvr_last_error_string(void) { return mvr::g_err; }

// ---- host-side staging helper (collate -> device-resident packed geometry) -------------
// Gathers n host arrays into one (pinned) destination with all host cores; optionally narrows int64 -> int32
// on the way (faces), halving the bytes that cross PCIe.  Plain host code: no CUDA calls.
static int g_host_threads = 0;   // 0 = OpenMP default
extern "C" int mvr_host_set_threads(int n) { g_host_threads = n > 0 ? n : 0; return 0; }

extern "C" int mvr_host_gather(const void* const* srcs, const int64_t* counts, int n, void* dst, int elem_bytes,
                               int narrow_i64_to_i32) {
  if (n < 0 || (n > 0 && (!srcs || !counts || !dst))) { mvr::set_error("mvr_host_gather: null pointer"); return -1; }
  if (narrow_i64_to_i32 && elem_bytes != 8) { mvr::set_error("mvr_host_gather: narrowing needs 8-byte elements"); return -2; }
  std::vector<int64_t> off((size_t)n + 1, 0);
  for (int i = 0; i < n; ++i) {
    if (counts[i] < 0) { mvr::set_error("mvr_host_gather: negative count"); return -3; }
    off[i + 1] = off[i] + counts[i];
  }
  const int64_t total = off[n];
  const int64_t kChunk = 1 << 15;               // elements per work item
  std::vector<int64_t> work;                    // (array, start) pairs flattened
  for (int i = 0; i < n; ++i)
    for (int64_t s = 0; s < counts[i]; s += kChunk) { work.push_back(i); work.push_back(s); }
  const int64_t nwork = (int64_t)work.size() / 2;
  const int nthreads = g_host_threads > 0 ? g_host_threads : omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (int64_t wkk = 0; wkk < nwork; ++wkk) {
    const int i = (int)work[2 * wkk];
    const int64_t s = work[2 * wkk + 1];
    const int64_t cnt = std::min<int64_t>(kChunk, counts[i] - s);
    if (narrow_i64_to_i32) {
      const int64_t* src = (const int64_t*)srcs[i] + s;
      int32_t* d = (int32_t*)dst + off[i] + s;
      for (int64_t e = 0; e < cnt; ++e) d[e] = (int32_t)src[e];
    } else {
      memcpy((char*)dst + (off[i] + s) * elem_bytes, (const char*)srcs[i] + s * elem_bytes, (size_t)(cnt * elem_bytes));
    }
  }
  (void)total;
  return 0;
}

namespace {
// One staging job cut into (array, start) work items of kChunk elements; the items are independent, so any set of threads may run them.
struct StagePlan {
  const void* const* vert_srcs; const int64_t* vert_counts; const void* const* face_srcs; const int64_t* face_counts;
  int n, face_elem_bytes, face_out_bytes; float* pinned_verts; void* pinned_faces;
  std::vector<int64_t> voff, foff, vwork, fwork;
  int64_t nv = 0, nf = 0;
  static constexpr int64_t kChunk = 1 << 14;
  void vert_item(int64_t w) const {
    const int i = (int)vwork[2 * w];
    const int64_t s = vwork[2 * w + 1], cnt = std::min<int64_t>(kChunk, vert_counts[i] - s);
    memcpy(pinned_verts + voff[i] + s, (const float*)vert_srcs[i] + s, (size_t)cnt * sizeof(float));
  }
  void face_item(int64_t w) const {
    const int i = (int)fwork[2 * w];
    const int64_t s = fwork[2 * w + 1], cnt = std::min<int64_t>(kChunk, face_counts[i] - s);
    if (face_out_bytes == 2) {      // saturate to [0, 65535]: mvr_mesh_prepare then clamps to the mesh's vertex count, as it does for int32
      uint16_t* d16 = (uint16_t*)pinned_faces + foff[i] + s;
      if (face_elem_bytes == 8) {
        const int64_t* src = (const int64_t*)face_srcs[i] + s;
        for (int64_t e = 0; e < cnt; ++e) { const int64_t x = src[e]; d16[e] = (uint16_t)(x < 0 ? 0 : (x > 65535 ? 65535 : x)); }
      } else {
        const int32_t* src = (const int32_t*)face_srcs[i] + s;
        for (int64_t e = 0; e < cnt; ++e) { const int32_t x = src[e]; d16[e] = (uint16_t)(x < 0 ? 0 : (x > 65535 ? 65535 : x)); }
      }
      return;
    }
    int32_t* d = (int32_t*)pinned_faces + foff[i] + s;
    if (face_elem_bytes == 8) {
      const int64_t* src = (const int64_t*)face_srcs[i] + s;
      for (int64_t e = 0; e < cnt; ++e) d[e] = (int32_t)src[e];
    } else {
      memcpy(d, (const int32_t*)face_srcs[i] + s, (size_t)cnt * sizeof(int32_t));
    }
  }
};

// Helper threads of the asynchronous staging (mvr_host_stage_meshes*_begin): a private pool, independent of OpenMP -- an OpenMP team
// started from the staging thread would fight the caller's own (spinning) team for the cores.  The coordinator (the staging thread)
// publishes a plan and works on it too; helpers sleep on a condition variable between jobs.
struct GatherPool {
  std::mutex mu; std::condition_variable cv_work, cv_idle;
  const StagePlan* plan = nullptr; uint64_t epoch = 0; int busy = 0, nthreads = 0;
  std::atomic<int64_t> next_v{0}, next_f{0}, done_v{0}, done_f{0};
  void drain(const StagePlan& p) {
    for (int64_t w; (w = next_v.fetch_add(1, std::memory_order_relaxed)) < p.nv;) { p.vert_item(w); done_v.fetch_add(1, std::memory_order_release); }
    for (int64_t w; (w = next_f.fetch_add(1, std::memory_order_relaxed)) < p.nf;) { p.face_item(w); done_f.fetch_add(1, std::memory_order_release); }
  }
  void helper(uint64_t seen) {      // seen: the epoch at creation -- a new helper waits for the NEXT job, it never checks out of one it did not join
    for (;;) {
      const StagePlan* p;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_work.wait(lk, [&] { return epoch != seen; });
        seen = epoch; p = plan;
      }
      if (p) drain(*p);
      {
        std::lock_guard<std::mutex> lk(mu);
        if (--busy == 0) cv_idle.notify_all();
      }
    }
  }
  void ensure(int want) {      // (coordinator only)
    std::lock_guard<std::mutex> lk(mu);
    while (nthreads < want) { std::thread(&GatherPool::helper, this, epoch).detach(); ++nthreads; }
  }
};
// Leaked on purpose, like the staging thread.  After fork() the child has none of the threads (and possibly a locked mutex inside the
// old object): stage_atfork_child() swaps in a fresh pool and a fresh worker, so that the child starts its own threads on first use.
GatherPool* g_gather_pool = nullptr;
GatherPool* gather_pool() { return g_gather_pool; }      // created together with the worker (worker())
}  // namespace

// Stage a batch of meshes for the device in ONE parallel region (renderer.py:67-68 + Meshes(...) packing): gather the
// per-mesh vertex arrays into pinned_verts, enqueue their H2D copy from one thread while the others already gather /
// narrow the faces into pinned_faces, then enqueue the faces' copy.  HOST pointers except dev_*; faces are int64
// (face_elem_bytes == 8, narrowed on the way) or int32.  One fork/join and the vertex copy overlapped with the
// face gather: the step's time-to-first-kernel is bounded by this call.  pooled: run the items on the private helper pool (the
// caller is the staging thread) instead of an OpenMP team.
static int stage_meshes_impl(const void* const* vert_srcs, const int64_t* vert_counts,
                             const void* const* face_srcs, const int64_t* face_counts, int n,
                             int face_elem_bytes, float* pinned_verts, void* pinned_faces_any,
                             float* dev_verts, void* dev_faces, void* stream, int face_out_bytes = 4,
                             int32_t* pinned_offs = nullptr, int32_t* dev_offs = nullptr, bool pooled = false) {
  if (n < 0 || (n > 0 && (!vert_srcs || !vert_counts || !face_srcs || !face_counts || !pinned_verts || !pinned_faces_any))) {
    mvr::set_error("mvr_host_stage_meshes: null pointer"); return -1;
  }
  if (face_elem_bytes != 4 && face_elem_bytes != 8) { mvr::set_error("mvr_host_stage_meshes: faces must be int32 or int64"); return -2; }
  StagePlan pl;
  pl.vert_srcs = vert_srcs; pl.vert_counts = vert_counts; pl.face_srcs = face_srcs; pl.face_counts = face_counts;
  pl.n = n; pl.face_elem_bytes = face_elem_bytes; pl.face_out_bytes = face_out_bytes; pl.pinned_verts = pinned_verts; pl.pinned_faces = pinned_faces_any;
  std::vector<int64_t>& voff = pl.voff; std::vector<int64_t>& foff = pl.foff;
  voff.assign((size_t)n + 1, 0); foff.assign((size_t)n + 1, 0);
  for (int i = 0; i < n; ++i) {
    if (vert_counts[i] < 0 || face_counts[i] < 0) { mvr::set_error("mvr_host_stage_meshes: negative count"); return -3; }
    voff[i + 1] = voff[i] + vert_counts[i];      // counts are in ELEMENTS (floats / indices)
    foff[i + 1] = foff[i] + face_counts[i];
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (pinned_offs) {      // vertex | face offsets of the packed batch (n + 1 each, in vertices / faces), first on the wire: 2n + 2 words
    for (int i = 0; i <= n; ++i) { pinned_offs[i] = (int32_t)(voff[i] / 3); pinned_offs[n + 1 + i] = (int32_t)(foff[i] / 3); }
    if (dev_offs) {
      const cudaError_t e = cudaMemcpyAsync(dev_offs, pinned_offs, (size_t)(2 * n + 2) * sizeof(int32_t), cudaMemcpyHostToDevice, st);
      if (e != cudaSuccess) { mvr::set_error("mvr_host_stage_meshes: cudaMemcpyAsync(offsets): %s", cudaGetErrorString(e)); return (int)e; }
    }
  }
  const int64_t kChunk = StagePlan::kChunk;
  for (int i = 0; i < n; ++i) for (int64_t s = 0; s < vert_counts[i]; s += kChunk) { pl.vwork.push_back(i); pl.vwork.push_back(s); }
  for (int i = 0; i < n; ++i) for (int64_t s = 0; s < face_counts[i]; s += kChunk) { pl.fwork.push_back(i); pl.fwork.push_back(s); }
  const int64_t nv = pl.nv = (int64_t)pl.vwork.size() / 2, nf = pl.nf = (int64_t)pl.fwork.size() / 2;
  int nthreads = g_host_threads > 0 ? g_host_threads : (pooled ? (int)std::thread::hardware_concurrency() : omp_get_max_threads());
  if (nthreads > 8) nthreads = 8;                 // memory-bound: more threads only add wake-up latency
  if (nthreads < 1) nthreads = 1;
  cudaError_t err_v = cudaSuccess;
  if (pooled) {
    GatherPool* g = gather_pool();
    const int helpers = (int)std::min<int64_t>(nthreads - 1, std::max<int64_t>(nv + nf - 1, 0));
    g->ensure(helpers);
    {
      std::lock_guard<std::mutex> lk(g->mu);
      g->next_v = 0; g->next_f = 0; g->done_v = 0; g->done_f = 0;
      g->plan = &pl; g->busy = g->nthreads; ++g->epoch;      // every helper checks in (and out) once per job, late ones find no items
    }
    if (g->nthreads > 0) g->cv_work.notify_all();
    for (int64_t w; (w = g->next_v.fetch_add(1, std::memory_order_relaxed)) < nv;) { pl.vert_item(w); g->done_v.fetch_add(1, std::memory_order_release); }
    while (g->done_v.load(std::memory_order_acquire) < nv) std::this_thread::yield();      // every vertex is staged
    if (dev_verts && voff[n] > 0) err_v = cudaMemcpyAsync(dev_verts, pinned_verts, (size_t)voff[n] * sizeof(float), cudaMemcpyHostToDevice, st);
    for (int64_t w; (w = g->next_f.fetch_add(1, std::memory_order_relaxed)) < nf;) { pl.face_item(w); g->done_f.fetch_add(1, std::memory_order_release); }
    while (g->done_f.load(std::memory_order_acquire) < nf) std::this_thread::yield();
    {      // the plan lives on this stack frame: wait until every helper has let go of it
      std::unique_lock<std::mutex> lk(g->mu);
      g->cv_idle.wait(lk, [g] { return g->busy == 0; });
      g->plan = nullptr;
    }
  } else {
#pragma omp parallel num_threads(nthreads)
    {
#pragma omp for schedule(dynamic, 1)
      for (int64_t w = 0; w < nv; ++w) pl.vert_item(w);      // implicit barrier: every vertex is staged
#pragma omp single nowait
      {
        if (dev_verts && voff[n] > 0) err_v = cudaMemcpyAsync(dev_verts, pinned_verts, (size_t)voff[n] * sizeof(float), cudaMemcpyHostToDevice, st);
      }
#pragma omp for schedule(dynamic, 1)
      for (int64_t w = 0; w < nf; ++w) pl.face_item(w);
    }
  }
  if (err_v != cudaSuccess) { mvr::set_error("mvr_host_stage_meshes: cudaMemcpyAsync(verts): %s", cudaGetErrorString(err_v)); return (int)err_v; }
  if (dev_faces && foff[n] > 0) {
    const cudaError_t e = cudaMemcpyAsync(dev_faces, pinned_faces_any, (size_t)foff[n] * (size_t)face_out_bytes, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { mvr::set_error("mvr_host_stage_meshes: cudaMemcpyAsync(faces): %s", cudaGetErrorString(e)); return (int)e; }
  }
  return 0;
}

extern "C" int mvr_host_stage_meshes(const void* const* vert_srcs, const int64_t* vert_counts,
                                     const void* const* face_srcs, const int64_t* face_counts, int n,
                                     int face_elem_bytes, float* pinned_verts, int32_t* pinned_faces,
                                     float* dev_verts, int32_t* dev_faces, void* stream) {
  return stage_meshes_impl(vert_srcs, vert_counts, face_srcs, face_counts, n, face_elem_bytes, pinned_verts, pinned_faces,
                           dev_verts, dev_faces, stream);
}

// The same for a caller that also wants the batch's offset table written and copied, and the faces optionally narrowed to uint16 ids
// (face_out_bytes 2, saturating; for batches whose meshes all have at most 65536 vertices: MVR_FACES_U16 in mvr_mesh_prepare -- a third
// of the H2D bytes less than int32).  pinned_offs / dev_offs: 2n + 2 int32 = vertex offsets (n + 1) | face offsets (n + 1).
extern "C" int mvr_host_stage_meshes_packed(const void* const* vert_srcs, const int64_t* vert_counts,
                                            const void* const* face_srcs, const int64_t* face_counts, int n,
                                            int face_elem_bytes, int face_out_bytes, float* pinned_verts, void* pinned_faces,
                                            int32_t* pinned_offs, float* dev_verts, void* dev_faces, int32_t* dev_offs, void* stream) {
  if (face_out_bytes != 2 && face_out_bytes != 4) { mvr::set_error("mvr_host_stage_meshes_packed: face_out_bytes must be 2 or 4"); return -2; }
  for (int i = 0; i < n; ++i)
    if (vert_counts && face_counts && (vert_counts[i] % 3 || face_counts[i] % 3)) { mvr::set_error("mvr_host_stage_meshes_packed: counts must be multiples of 3"); return -3; }
  return stage_meshes_impl(vert_srcs, vert_counts, face_srcs, face_counts, n, face_elem_bytes, pinned_verts, pinned_faces,
                           dev_verts, dev_faces, stream, face_out_bytes, pinned_offs, dev_offs);
}

// ---- asynchronous variant: the staging runs on a persistent native worker thread while the caller (Python) prepares
// the rest of the step (cameras, output buffers); _end() joins it.  One job in flight at a time. -----------------------
namespace {
struct StageJob {
  const void* const* vert_srcs; const int64_t* vert_counts; const void* const* face_srcs; const int64_t* face_counts;
  int n, face_elem_bytes; float* pinned_verts; void* pinned_faces; float* dev_verts; void* dev_faces; void* stream;
  int device, status; bool pending, done; char err[512];
  int face_out_bytes; int32_t* pinned_offs; int32_t* dev_offs;
};
struct StageWorker {
  std::mutex mu; std::condition_variable cv_job, cv_done; StageJob job; bool started = false; int next_id = 1, cur_id = 0;
};
StageWorker* g_stage_worker = nullptr;
void stage_atfork_child() { g_stage_worker = new StageWorker(); g_gather_pool = new GatherPool(); }
StageWorker* worker() {      // leaked on purpose: outlives exit handlers
  static std::once_flag once;
  std::call_once(once, [] { g_stage_worker = new StageWorker(); g_gather_pool = new GatherPool(); pthread_atfork(nullptr, nullptr, stage_atfork_child); });
  return g_stage_worker;
}

void stage_worker_loop() {
  StageWorker* w = worker();
  for (;;) {
    std::unique_lock<std::mutex> lk(w->mu);
    w->cv_job.wait(lk, [w] { return w->job.pending; });
    StageJob j = w->job;
    lk.unlock();
    int rc = 0;
    if (j.dev_verts || j.dev_faces) {
      const cudaError_t e = cudaSetDevice(j.device);
      if (e != cudaSuccess) { mvr::set_error("mvr_host_stage_meshes_begin: cudaSetDevice(%d): %s", j.device, cudaGetErrorString(e)); rc = (int)e; }
    }
    if (rc == 0)
      rc = stage_meshes_impl(j.vert_srcs, j.vert_counts, j.face_srcs, j.face_counts, j.n, j.face_elem_bytes, j.pinned_verts,
                             j.pinned_faces, j.dev_verts, j.dev_faces, j.stream, j.face_out_bytes, j.pinned_offs, j.dev_offs, true);
    lk.lock();
    w->job.status = rc;
    strncpy(w->job.err, mvr::g_err, sizeof(w->job.err) - 1);
    w->job.err[sizeof(w->job.err) - 1] = 0;
    w->job.pending = false; w->job.done = true;
    w->cv_done.notify_all();
  }
}
}  // namespace

static int stage_begin_impl(const void* const* vert_srcs, const int64_t* vert_counts, const void* const* face_srcs,
                            const int64_t* face_counts, int n, int face_elem_bytes, int face_out_bytes, float* pinned_verts,
                            void* pinned_faces, int32_t* pinned_offs, float* dev_verts, void* dev_faces, int32_t* dev_offs,
                            int device, void* stream) {
  StageWorker* w = worker();
  std::unique_lock<std::mutex> lk(w->mu);
  if (w->job.pending || (w->cur_id != 0)) { mvr::set_error("mvr_host_stage_meshes_begin: a staging job is already in flight"); return -10; }
  if (!w->started) { std::thread(stage_worker_loop).detach(); w->started = true; }
  w->job = StageJob{vert_srcs, vert_counts, face_srcs, face_counts, n, face_elem_bytes, pinned_verts, pinned_faces, dev_verts,
                    dev_faces, stream, device, 0, true, false, {0}, face_out_bytes, pinned_offs, dev_offs};
  w->cur_id = w->next_id++;
  if (w->next_id <= 0) w->next_id = 1;
  w->cv_job.notify_one();
  return w->cur_id;
}

extern "C" int mvr_host_stage_meshes_begin(const void* const* vert_srcs, const int64_t* vert_counts,
                                           const void* const* face_srcs, const int64_t* face_counts, int n,
                                           int face_elem_bytes, float* pinned_verts, int32_t* pinned_faces,
                                           float* dev_verts, int32_t* dev_faces, int device, void* stream) {
  return stage_begin_impl(vert_srcs, vert_counts, face_srcs, face_counts, n, face_elem_bytes, 4, pinned_verts, pinned_faces, nullptr,
                          dev_verts, dev_faces, nullptr, device, stream);
}

// mvr_host_stage_meshes_packed on the worker thread (join with mvr_host_stage_meshes_end)
extern "C" int mvr_host_stage_meshes_packed_begin(const void* const* vert_srcs, const int64_t* vert_counts,
                                                  const void* const* face_srcs, const int64_t* face_counts, int n,
                                                  int face_elem_bytes, int face_out_bytes, float* pinned_verts, void* pinned_faces,
                                                  int32_t* pinned_offs, float* dev_verts, void* dev_faces, int32_t* dev_offs,
                                                  int device, void* stream) {
  if (face_out_bytes != 2 && face_out_bytes != 4) { mvr::set_error("mvr_host_stage_meshes_packed_begin: face_out_bytes must be 2 or 4"); return -2; }
  return stage_begin_impl(vert_srcs, vert_counts, face_srcs, face_counts, n, face_elem_bytes, face_out_bytes, pinned_verts, pinned_faces,
                          pinned_offs, dev_verts, dev_faces, dev_offs, device, stream);
}

extern "C" int mvr_host_stage_meshes_end(int job) {
  StageWorker* w = worker();
  std::unique_lock<std::mutex> lk(w->mu);
  if (job <= 0 || job != w->cur_id) { mvr::set_error("mvr_host_stage_meshes_end: unknown job %d", job); return -11; }
  w->cv_done.wait(lk, [w] { return w->job.done; });
  const int rc = w->job.status;
  if (rc != 0) mvr::set_error("%s", w->job.err);
  w->job.done = false;
  w->cur_id = 0;
  return rc;
}
