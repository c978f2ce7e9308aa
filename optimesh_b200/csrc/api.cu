// C-ABI of liboptimesh_b200.so (see include/optimesh_b200.h).
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include <sys/mman.h>
#include <unistd.h>

#include "common.cuh"
#include "geom.cuh"

static thread_local std::string g_err;

void om_set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}

int om_fetch_scalars(om_handle* h) {
  CUDA_TRY(cudaMemcpyAsync(h->hs, h->ds, sizeof(DevScalars), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return OM_OK;
}

// Evaluates the error bits of the LAST om_fetch_scalars (every caller fetches right before).
int om_check_dev_err(om_handle* h) {
  const int e = h->hs->err;
  if (!e) return OM_OK;
  CUDA_TRY(cudaMemsetAsync(&h->ds->err, 0, sizeof(int), h->stream));
  if (e & OM_DEV_INDEX) {
    om_set_error("cells refer to vertices outside [0, N)");
    return OM_ERR_INDEX;
  }
  if (e & OM_DEV_DEGENERATE) {
    om_set_error("Degenerate cells.");
    return OM_ERR_DEGENERATE;
  }
  om_set_error("inconsistent mesh topology (non-manifold edge or broken vertex star)");
  return OM_ERR_NONMANIFOLD;
}

namespace {

bool is_pinned_host(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    int cur = -1;
    cudaGetDevice(&cur);
    if (prev >= 0 && cur != prev) cudaSetDevice(prev);
  }
};

// ---- host <-> device transfers through pinned staging chunks.
// Pageable user memory is the slow side of the reference-facing API (640 MB of points and
// int64 cells at 10M vertices): a plain cudaMemcpy is bounded by one thread touching fresh
// pages.  Here PCIe moves pinned chunks while a few host threads copy the previous chunk
// to/from the user's array, converting int64 <-> int32 cell indices on the fly so that only
// 4 bytes per index cross the bus.
constexpr size_t STAGE_BYTES = 16u << 20;
// host threads that convert / copy between user memory and the pinned chunks
// (OM_STAGE_THREADS overrides).  Default: one per host core up to 16 -- on the 16-core B200
// host, optimize_points_cells at 9.95M vertices took 141 / 109 / 98 ms with 6 / 12 / 16 threads
// (the int64 <-> int32 cell conversion and the page faults of fresh result arrays scale).
int stage_threads() {
  static const int n = [] {
    const char* e = getenv("OM_STAGE_THREADS");
    const int hw = (int)std::thread::hardware_concurrency();
    int v = e ? atoi(e) : std::min(std::max(hw, 4), 16);
    // several processes on one host (one per GPU, torchrun sets LOCAL_WORLD_SIZE): they share
    // the cores, 16 staging threads each would only fight for them
    const char* lws = getenv("LOCAL_WORLD_SIZE");
    if (!e && lws && atoi(lws) > 1) v = std::max(2, std::max(hw, 4) / atoi(lws));
    return std::max(1, std::min(v, 64));
  }();
  return n;
}

struct Stage {
  void* pin[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  bool ok = false;
  Stage() {
    ok = cudaMallocHost(&pin[0], STAGE_BYTES) == cudaSuccess &&
         cudaMallocHost(&pin[1], STAGE_BYTES) == cudaSuccess &&
         cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming) == cudaSuccess;
  }
  ~Stage() {
    for (int i = 0; i < 2; i++) {
      if (pin[i]) cudaFreeHost(pin[i]);
      if (ev[i]) cudaEventDestroy(ev[i]);
    }
  }
};

// A small persistent pool of host threads (spawning 16 threads per 16 MB chunk cost more
// than the copies they did: ~40 chunks x 16 threads x 30 us per call of the API).
class HostPool {
 public:
  static HostPool& get() {
    static HostPool* pool = new HostPool(stage_threads());  // never destroyed: threads may
    return *pool;                                           // outlive static destruction order
  }
  // runs f(begin, end) over [0, n) split into one piece per thread; returns when all are done
  template <typename F>
  void run(size_t n, F f) {
    if (n == 0) return;
    const int T = (int)workers_.size() + 1;
    if (n < (1u << 16) || T == 1) {
      f((size_t)0, n);
      return;
    }
    std::lock_guard<std::mutex> serial(serial_);  // one parallel region at a time
    const size_t per = (n + T - 1) / T;
    std::function<void(int)> job = [&](int t) {
      const size_t b = std::min(n, (size_t)t * per), e = std::min(n, b + per);
      if (b < e) f(b, e);
    };
    {
      std::lock_guard<std::mutex> lock(m_);
      job_ = &job;
      pending_ = (int)workers_.size();
      generation_++;
    }
    cv_.notify_all();
    job(T - 1);  // the caller takes the last piece
    std::unique_lock<std::mutex> lock(m_);
    done_.wait(lock, [&] { return pending_ == 0; });
    job_ = nullptr;
  }

 private:
  explicit HostPool(int threads) {
    for (int t = 0; t + 1 < threads; t++)
      workers_.emplace_back([this, t] {
        uint64_t seen = 0;
        while (true) {
          std::function<void(int)>* job;
          {
            std::unique_lock<std::mutex> lock(m_);
            cv_.wait(lock, [&] { return generation_ != seen; });
            seen = generation_;
            job = job_;
          }
          (*job)(t);
          {
            std::lock_guard<std::mutex> lock(m_);
            if (--pending_ == 0) done_.notify_one();
          }
        }
      });
    for (auto& w : workers_) w.detach();
  }
  std::vector<std::thread> workers_;
  std::mutex m_, serial_;
  std::condition_variable cv_, done_;
  std::function<void(int)>* job_ = nullptr;
  int pending_ = 0;
  uint64_t generation_ = 0;
};

template <typename F>
void parallel_chunks(size_t n, F f) {  // f(begin, end) on the pool's host threads
  HostPool::get().run(n, f);
}

// Staging buffers are kept for the lifetime of the process and handed from handle to handle
// (pinning and unpinning 2 x 16 MB costs several milliseconds per optimize_points_cells call
// otherwise).  Entries left in the cache at exit are deliberately not freed: the CUDA context
// may already be gone by then.
std::mutex g_stage_mutex;
std::vector<std::pair<int, Stage*>> g_stage_cache;  // (device, buffers)

Stage& stage_of(om_handle* h) {
  if (!h->stage) {
    std::lock_guard<std::mutex> lock(g_stage_mutex);
    for (size_t i = 0; i < g_stage_cache.size(); i++)
      if (g_stage_cache[i].first == h->device) {
        h->stage = g_stage_cache[i].second;
        g_stage_cache.erase(g_stage_cache.begin() + i);
        break;
      }
  }
  if (!h->stage) h->stage = new Stage();
  return *(Stage*)h->stage;
}

void stage_release(om_handle* h) {
  Stage* st = (Stage*)h->stage;
  h->stage = nullptr;
  if (!st) return;
  if (!st->ok) {
    delete st;
    return;
  }
  std::lock_guard<std::mutex> lock(g_stage_mutex);
  if (g_stage_cache.size() < 4)
    g_stage_cache.emplace_back(h->device, st);
  else
    delete st;
}

// device (elements of DEV_T) -> user memory (elements of HOST_T)

template <typename DEV_T, typename HOST_T>
int staged_d2h(om_handle* h, const DEV_T* src_dev, HOST_T* dst_host, size_t n) {
  Stage& st = stage_of(h);
  if (!st.ok) {
    om_set_error("pinned staging allocation failed");
    return OM_ERR_CUDA;
  }
  const size_t per = STAGE_BYTES / sizeof(DEV_T);
  const size_t nchunks = (n + per - 1) / per;
  auto issue = [&](size_t k) {
    const size_t b = k * per, cnt = std::min(per, n - b);
    cudaMemcpyAsync(st.pin[k & 1], src_dev + b, cnt * sizeof(DEV_T), cudaMemcpyDeviceToHost,
                    h->stream);
    cudaEventRecord(st.ev[k & 1], h->stream);
  };
  if (nchunks) issue(0);
  for (size_t k = 0; k < nchunks; k++) {
    CUDA_TRY(cudaEventSynchronize(st.ev[k & 1]));
    if (k + 1 < nchunks) issue(k + 1);  // PCIe fills the other buffer while we copy this one
    const size_t b = k * per, cnt = std::min(per, n - b);
    const DEV_T* pin = (const DEV_T*)st.pin[k & 1];
    HOST_T* out = dst_host + b;
    parallel_chunks(cnt, [=](size_t lo, size_t hi) {
      if (sizeof(DEV_T) == sizeof(HOST_T))
        memcpy((void*)(out + lo), (const void*)(pin + lo), (hi - lo) * sizeof(DEV_T));
      else
        for (size_t i = lo; i < hi; i++) out[i] = (HOST_T)pin[i];
    });
  }
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return OM_OK;
}

// user memory (HOST_T) -> device (DEV_T); *bad is set if a narrowed index does not fit
template <typename HOST_T, typename DEV_T>
int staged_h2d(om_handle* h, const HOST_T* src_host, DEV_T* dst_dev, size_t n, bool* bad) {
  Stage& st = stage_of(h);
  cudaStream_t stream = h->stream;
  if (!st.ok) {
    om_set_error("pinned staging allocation failed");
    return OM_ERR_CUDA;
  }
  const size_t per = STAGE_BYTES / sizeof(DEV_T);
  const size_t nchunks = (n + per - 1) / per;
  std::atomic<bool> overflow(false);
  for (size_t k = 0; k < nchunks; k++) {
    const size_t b = k * per, cnt = std::min(per, n - b);
    if (k >= 2) CUDA_TRY(cudaEventSynchronize(st.ev[k & 1]));  // buffer free again?
    DEV_T* pin = (DEV_T*)st.pin[k & 1];
    const HOST_T* in = src_host + b;
    parallel_chunks(cnt, [&, pin, in](size_t lo, size_t hi) {
      if (sizeof(DEV_T) == sizeof(HOST_T)) {
        memcpy((void*)(pin + lo), (const void*)(in + lo), (hi - lo) * sizeof(DEV_T));
      } else {
        // branch-free so that the host compiler vectorises it: a value outside [0, 2^31) has
        // a bit at or above position 31 (negative values have them all)
        unsigned long long acc = 0ull;
        const HOST_T* __restrict__ src = in;
        DEV_T* __restrict__ dst = pin;
        for (size_t i = lo; i < hi; i++) {
          const unsigned long long v = (unsigned long long)(long long)src[i];
          acc |= v;
          dst[i] = (DEV_T)v;
        }
        if (acc >> 31) overflow = true;
      }
    });
    cudaMemcpyAsync(dst_dev + b, pin, cnt * sizeof(DEV_T), cudaMemcpyHostToDevice, stream);
    cudaEventRecord(st.ev[k & 1], stream);
  }
  CUDA_TRY(cudaStreamSynchronize(stream));
  if (bad) *bad = overflow.load();
  return OM_OK;
}

#define OM_ENTER(h)                              \
  if (!(h)) {                                    \
    om_set_error("null handle");                 \
    return OM_ERR_ARG;                           \
  }                                              \
  DeviceGuard _guard((h)->device)

template <int D>
__global__ void k_export_points(const double* __restrict__ x, const int* __restrict__ perm, int N,
                                double* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  Vec<D> p = ld_point<D>(x, i);
  size_t dst = perm ? perm[i] : i;
#pragma unroll
  for (int k = 0; k < D; k++) out[dst * D + k] = p.v[k];
}

template <int D>
__global__ void k_import_points(const double* __restrict__ in, const int* __restrict__ perm, int N,
                                double* __restrict__ x) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  size_t src = perm ? perm[i] : i;
  Vec<D> p;
#pragma unroll
  for (int k = 0; k < D; k++) p.v[k] = in[src * D + k];
  st_point<D>(x, i, p);
}

template <typename T>
__global__ void k_export_cells(const int4* __restrict__ cells, const int* __restrict__ perm, int C,
                               T* __restrict__ out) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  int4 cl = cells[c];
  size_t row = (size_t)cl.w;
  out[3 * row] = (T)(perm ? perm[cl.x] : cl.x);
  out[3 * row + 1] = (T)(perm ? perm[cl.y] : cl.y);
  out[3 * row + 2] = (T)(perm ? perm[cl.z] : cl.z);
}

__global__ void k_export_flags(const uint8_t* __restrict__ f, const int* __restrict__ perm, int N,
                               uint8_t* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) out[perm ? perm[i] : i] = f[i];
}

template <int D>
__global__ void k_pack(const double* __restrict__ x, const int* __restrict__ inv,
                       const int* __restrict__ idx, int64_t n, double* __restrict__ buf) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int v = idx[i];
  if (inv) v = inv[v];
  Vec<D> p = ld_point<D>(x, v);
#pragma unroll
  for (int k = 0; k < D; k++) buf[i * D + k] = p.v[k];
}

template <int D>
__global__ void k_unpack(double* __restrict__ x, const int* __restrict__ inv,
                         const int* __restrict__ idx, int64_t n, const double* __restrict__ buf) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int v = idx[i];
  if (inv) v = inv[v];
  Vec<D> p;
#pragma unroll
  for (int k = 0; k < D; k++) p.v[k] = buf[i * D + k];
  st_point<D>(x, v, p);
}

__global__ void k_pin(uint8_t* __restrict__ f, const int* __restrict__ inv,
                      const int* __restrict__ idx, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int v = idx[i];
  f[inv ? inv[v] : v] = 1;
}

int create_common(om_handle** out, int device, void* stream, int64_t N, int dim, int64_t C,
                  const double* points, const void* cells, int itemsize, int flags,
                  bool inputs_on_device) {
  if (!out) {
    om_set_error("null output handle");
    return OM_ERR_ARG;
  }
  *out = nullptr;
  if (dim != 2 && dim != 3) {
    om_set_error("points must have 2 or 3 columns, got %d", dim);
    return OM_ERR_ARG;
  }
  if (itemsize != 4 && itemsize != 8) {
    om_set_error("cells itemsize must be 4 or 8, got %d", itemsize);
    return OM_ERR_ARG;
  }
  if (N < 0 || C < 0 || N >= (1ll << 29) || C >= (1ll << 29)) {
    om_set_error("mesh size out of range (N=%lld, C=%lld)", (long long)N, (long long)C);
    return OM_ERR_ARG;
  }
  if ((N > 0 && !points) || (C > 0 && !cells)) {
    om_set_error("null input array");
    return OM_ERR_ARG;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    om_set_error("no CUDA device available (%s); optimesh_b200 has no CPU fallback",
                 cudaGetErrorString(e));
    return OM_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) {
    om_set_error("device %d out of range (have %d)", device, ndev);
    return OM_ERR_ARG;
  }
  DeviceGuard guard(device);
  {
    // keep released blocks in the device's stream-ordered pool (see om_malloc)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      uint64_t keep = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  om_handle* h = new om_handle();
  h->device = device;
  h->N = N;
  h->C = C;
  h->D = dim;
  h->PD = dim == 2 ? 2 : 4;
  h->cells_itemsize = itemsize;
  if (getenv("OM_NO_RINGS")) h->use_rings = false;  // diagnostics: star walk only
  int rc = OM_OK;
  auto fail = [&](int code) {
    om_destroy(h);
    return code;
  };
  if (stream) {
    h->stream = (cudaStream_t)stream;
  } else {
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
      om_set_error("cudaStreamCreate failed");
      return fail(OM_ERR_CUDA);
    }
    h->own_stream = true;
  }
  const double* pdev = points;
  const void* cdev = cells;
  double* pup = nullptr;
  void* cup = nullptr;
  if (!inputs_on_device) {
    // cells always reach the device as int32 (narrowed by the host threads if needed)
    const size_t pb = sizeof(double) * (size_t)N * dim, cb = sizeof(int) * 3 * (size_t)C;
    if (om_malloc(h, &pup, std::max<size_t>(pb, 8)) != cudaSuccess ||
        om_malloc(h, &cup, std::max<size_t>(cb, 8)) != cudaSuccess) {
      om_set_error("device allocation failed");
      om_free(h, pup);
      return fail(OM_ERR_CUDA);
    }
    bool bad = false;
    rc = staged_h2d<double, double>(h, points, pup, (size_t)N * dim, nullptr);
    if (rc == OM_OK)
      rc = itemsize == 4
               ? staged_h2d<int, int>(h, (const int*)cells, (int*)cup, (size_t)3 * C, nullptr)
               : staged_h2d<long long, int>(h, (const long long*)cells, (int*)cup,
                                            (size_t)3 * C, &bad);
    if (rc == OM_OK && bad) {
      om_set_error("cells refer to vertices outside [0, N)");
      rc = OM_ERR_INDEX;
    }
    if (rc != OM_OK) {
      om_free(h, pup);
      om_free(h, cup);
      return fail(rc);
    }
    h->cells_itemsize = 4;
    pdev = pup;
    cdev = cup;
  }
  rc = om_setup_mesh(h, pdev, cdev, flags);
  cudaStreamSynchronize(h->stream);
  om_free(h, pup);
  om_free(h, cup);
  if (rc != OM_OK) return fail(rc);
  *out = h;
  return OM_OK;
}

}  // namespace

extern "C" {

const char* om_last_error(void) { return g_err.c_str(); }

int om_device_count(int* n) {
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (n) *n = (e == cudaSuccess) ? c : 0;
  return OM_OK;
}

int om_create(om_handle** h, int device, void* stream, int64_t N, int dim, int64_t C,
              const double* points_host, const void* cells_host, int cells_itemsize, int flags) {
  return create_common(h, device, stream, N, dim, C, points_host, cells_host, cells_itemsize,
                       flags, false);
}

int om_create_device(om_handle** h, int device, void* stream, int64_t N, int dim, int64_t C,
                     const double* points_dev, const void* cells_dev, int cells_itemsize,
                     int flags) {
  return create_common(h, device, stream, N, dim, C, points_dev, cells_dev, cells_itemsize, flags,
                       true);
}

int om_destroy(om_handle* h) {
  if (!h) return OM_OK;
  DeviceGuard guard(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  om_pl_destroy(h);
  om_shared_destroy(h);  // (a shared handle's mesh arrays are mappings, not allocations)
  om_free(h, h->x);
  om_free(h, h->xnew);
  om_free(h, h->cells);
  om_free(h, h->adj);
  om_free(h, h->adj_tmp);
  om_free(h, h->v2c);
  om_free(h, h->bflag);
  om_free(h, h->ring);
  om_free(h, h->ringc);
  om_free(h, h->dirty);
  om_free(h, h->dirty_epoch);
  om_free(h, h->diff2);
  om_free(h, h->vflags);
  om_free(h, h->valid_epoch);
  om_free(h, h->band);
  om_free(h, h->band_mark);
  om_free(h, h->perm);
  om_free(h, h->inv_perm);
  om_free(h, h->cand);
  om_free(h, h->work);
  om_free(h, h->work_epoch);
  om_free(h, h->cand_epoch);
  om_free(h, h->sarr);
  om_free(h, h->recs);
  om_free(h, h->flip_epoch);
  om_free(h, h->reloc);
  om_free(h, h->nbr_ptr);
  om_free(h, h->nbr_idx);
  om_free(h, h->nbr_w);
  om_free(h, h->pcg_buf);
  om_free(h, h->target_buf);
  om_free(h, h->ds);
  om_free(h, h->partials);
  if (h->hs) cudaFreeHost(h->hs);
  for (int i = 0; i < 4; i++)
    if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  stage_release(h);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return OM_OK;
}

int om_set_method(om_handle* h, int method, double omega) {
  OM_ENTER(h);
  if (method < OM_LLOYD || method > OM_CPT_QUASI_NEWTON) {
    om_set_error("unknown method id %d", method);
    return OM_ERR_ARG;
  }
  h->method = method;
  h->omega = omega;
  return OM_OK;
}

int om_set_limiter(om_handle* h, int on) {
  OM_ENTER(h);
  h->limiter = on ? 1 : 0;
  return OM_OK;
}

int om_set_odt_boundary_barycenters(om_handle* h, int on) {
  OM_ENTER(h);
  h->odt_bary = on ? 1 : 0;
  return OM_OK;
}

int om_set_surface(om_handle* h, int kind, double tol, const double* params, int max_sweeps) {
  OM_ENTER(h);
  if (kind != 0 && kind != 1) {
    om_set_error("unknown surface kind %d", kind);
    return OM_ERR_ARG;
  }
  if (kind == 1 && h->D != 3) {
    om_set_error("the sphere surface needs 3D points");
    return OM_ERR_ARG;
  }
  h->surf_kind = kind;
  h->surf_tol = tol;
  if (params) memcpy(h->surf_params, params, 4 * sizeof(double));
  if (max_sweeps > 0) h->surf_max_sweeps = max_sweeps;
  return OM_OK;
}

int om_set_solver(om_handle* h, double rtol, int max_iter) {
  OM_ENTER(h);
  h->solver_rtol = rtol;
  h->solver_max_iter = max_iter;
  return OM_OK;
}

int om_flip_until_delaunay(om_handle* h, double tol, int max_rounds, int64_t* n_flips,
                           int32_t* n_rounds, int32_t* cap_hit) {
  OM_ENTER(h);
  return om_flip_impl(h, tol, max_rounds, n_flips, n_rounds, cap_hit);
}

int om_update_points(om_handle* h, double tol, om_step_stats* out) {
  OM_ENTER(h);
  if (out) memset(out, 0, sizeof(*out));
  return om_update_points_impl(h, tol, out, false, nullptr);
}

int om_project(om_handle* h, int32_t* sweeps) {
  OM_ENTER(h);
  h->delaunay_clean = false;
  return om_project_impl(h, sweeps);
}

int om_step(om_handle* h, double tol, om_step_stats* out) {
  OM_ENTER(h);
  om_step_stats st;
  memset(&st, 0, sizeof(st));
  // the statistics of the point update ride on the first readback of the flip pass
  const bool defer = h->surf_kind == 0 && h->C > 0 && !om_is_solve_method(h->method);
  OM_TRY(om_update_points_impl(h, tol, &st, false, nullptr, defer));
  OM_TRY(om_project_impl(h, &st.surface_sweeps));
  OM_TRY(om_flip_impl(h, 0.0, 100, &st.n_flips, &st.n_flip_rounds, &st.flip_cap_hit));
  if (defer) om_step_stats_from_scalars(h, tol, &st);
  if (out) *out = st;
  return OM_OK;
}

int om_run(om_handle* h, double tol, int64_t max_num_steps, int64_t* steps_done,
           om_step_stats* last) {
  OM_ENTER(h);
  if (max_num_steps < 1) {
    om_set_error("max_num_steps must be >= 1");
    return OM_ERR_ARG;
  }
  // fixed-point methods on the whole mesh without a surface: the loop runs on the device
  // (loop.cu); OM_NO_PIPELINE=1 keeps the step-by-step order below (diagnostics)
  static const bool no_pipeline = getenv("OM_NO_PIPELINE") != nullptr;
  if (!no_pipeline && !om_is_solve_method(h->method) && h->surf_kind == 0 && h->own_hi < 0 &&
      h->N > 0 && h->C > 0)
    return om_run_pipelined(h, tol, max_num_steps, steps_done, last);
  om_step_stats st;
  memset(&st, 0, sizeof(st));
  int64_t nf = 0;
  int32_t nr = 0, cap = 0;
  if (!h->delaunay_clean) OM_TRY(om_flip_impl(h, 0.0, 100, &nf, &nr, &cap));
  int64_t k = 0;
  h->run_flips = h->run_rounds = h->run_limited = h->run_deferred = 0;
  while (true) {
    k++;
    OM_TRY(om_step(h, tol, &st));
    h->run_flips += st.n_flips;
    h->run_rounds += st.n_flip_rounds;
    h->run_limited += st.n_limited;
    if (st.is_final || k >= max_num_steps) break;
  }
  if (steps_done) *steps_done = k;
  if (last) *last = st;
  return OM_OK;
}

int om_random_walk(om_handle* h, int rounds, uint64_t seed, double amplitude, int64_t* n_flips) {
  OM_ENTER(h);
  if (rounds < 0 || !(amplitude > 0.0) || amplitude > 1.0) {
    om_set_error("om_random_walk: rounds >= 0 and 0 < amplitude <= 1 expected");
    return OM_ERR_ARG;
  }
  int64_t total = 0, nf = 0;
  int32_t nr = 0, cap = 0;
  OM_TRY(om_flip_impl(h, 0.0, 100, &nf, &nr, &cap));
  total += nf;
  for (int r = 0; r < rounds; r++) {
    OM_TRY(om_random_move_impl(h, seed, r, amplitude));
    OM_TRY(om_flip_impl(h, 0.0, 100, &nf, &nr, &cap));
    total += nf;
  }
  if (n_flips) *n_flips = total;
  return OM_OK;
}

int om_run_prepare(om_handle* h) {
  OM_ENTER(h);
  if (om_is_solve_method(h->method) || h->surf_kind != 0 || h->own_hi >= 0) return OM_OK;
  return om_pl_prepare(h);
}

int om_get_run_totals(om_handle* h, int64_t* n_flips, int64_t* n_flip_rounds, int64_t* n_limited,
                      int64_t* n_deferred) {
  OM_ENTER(h);
  if (n_deferred) *n_deferred = h->run_deferred;
  if (n_flips) *n_flips = h->run_flips;
  if (n_flip_rounds) *n_flip_rounds = h->run_rounds;
  if (n_limited) *n_limited = h->run_limited;
  return OM_OK;
}

int om_new_points(om_handle* h, double* out_host) {
  OM_ENTER(h);
  if (h->N == 0) return OM_OK;
  double *target = nullptr, *flat = nullptr;
  CUDA_TRY(om_malloc(h, &target, sizeof(double) * h->N * h->PD));
  CUDA_TRY(om_malloc(h, &flat, sizeof(double) * h->N * h->D));
  int rc = om_update_points_impl(h, 0.0, nullptr, true, target);
  if (rc == OM_OK) {
    const int B = 256, G = om_grid(h->N, B);
    if (h->D == 2)
      OM_LAUNCH(h, k_export_points<2>, G, B, target, h->perm, (int)h->N, flat);
    else
      OM_LAUNCH(h, k_export_points<3>, G, B, target, h->perm, (int)h->N, flat);
    cudaMemcpyAsync(out_host, flat, sizeof(double) * h->N * h->D, cudaMemcpyDeviceToHost,
                    h->stream);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) {
      om_set_error("copy of new points failed: %s", cudaGetErrorString(cudaGetLastError()));
      rc = OM_ERR_CUDA;
    }
  }
  om_free(h, target);
  om_free(h, flat);
  return rc;
}

int om_targets_device(om_handle* h, double** targets_dev) {
  OM_ENTER(h);
  if (!targets_dev) {
    om_set_error("om_targets_device: null output pointer");
    return OM_ERR_ARG;
  }
  *targets_dev = nullptr;
  if (h->N == 0) return OM_OK;
  if (!h->target_buf)
    CUDA_TRY(om_malloc(h, &h->target_buf, sizeof(double) * (h->N + OM_POINT_PAD) * h->PD));
  OM_TRY(om_update_points_impl(h, 0.0, nullptr, true, h->target_buf));
  *targets_dev = h->target_buf;
  return OM_OK;
}

int om_update_from_targets(om_handle* h, const double* targets_dev, double tol,
                           om_step_stats* out) {
  OM_ENTER(h);
  if (out) memset(out, 0, sizeof(*out));
  if (!targets_dev && h->N > 0) {
    om_set_error("om_update_from_targets: null targets");
    return OM_ERR_ARG;
  }
  return om_update_from_targets_impl(h, targets_dev, tol, out);
}

int om_solve_graph_laplacian(om_handle* h, double rtol, int max_iter, int32_t* iters,
                             double* rel_residual) {
  OM_ENTER(h);
  OM_TRY(om_pcg_impl(h, rtol, max_iter, iters, rel_residual, h->xnew));
  std::swap(h->x, h->xnew);
  h->delaunay_clean = false;
  return OM_OK;
}

int om_stats(om_handle* h, int64_t* angle_hist72, int64_t* q_hist40, double* summary8) {
  OM_ENTER(h);
  return om_stats_impl(h, angle_hist72, q_hist40, summary8);
}

int om_get_points(om_handle* h, double* out_host) {
  OM_ENTER(h);
  if (h->N == 0) return OM_OK;
  double* flat = nullptr;
  CUDA_TRY(om_malloc(h, &flat, sizeof(double) * h->N * h->D));
  const int B = 256, G = om_grid(h->N, B);
  if (h->D == 2)
    OM_LAUNCH(h, k_export_points<2>, G, B, h->x, h->perm, (int)h->N, flat);
  else
    OM_LAUNCH(h, k_export_points<3>, G, B, h->x, h->perm, (int)h->N, flat);
  int rc = OM_OK;
  if (is_pinned_host(out_host)) {
    // pinned destination (om_result_alloc): one DMA, no staging
    if (cudaMemcpyAsync(out_host, flat, sizeof(double) * h->N * h->D, cudaMemcpyDeviceToHost,
                        h->stream) != cudaSuccess ||
        cudaStreamSynchronize(h->stream) != cudaSuccess) {
      om_set_error("copy of points failed: %s", cudaGetErrorString(cudaGetLastError()));
      rc = OM_ERR_CUDA;
    }
  } else {
    rc = staged_d2h<double, double>(h, flat, out_host, (size_t)h->N * h->D);
  }
  om_free(h, flat);
  return rc;
}

int om_set_points(om_handle* h, const double* in_host) {
  OM_ENTER(h);
  if (h->N == 0) return OM_OK;
  h->delaunay_clean = false;
  double* flat = nullptr;
  CUDA_TRY(om_malloc(h, &flat, sizeof(double) * h->N * h->D));
  {
    int rc = staged_h2d<double, double>(h, in_host, flat, (size_t)h->N * h->D, nullptr);
    if (rc != OM_OK) {
      om_free(h, flat);
      return rc;
    }
  }
  const int B = 256, G = om_grid(h->N, B);
  if (h->D == 2)
    OM_LAUNCH(h, k_import_points<2>, G, B, flat, h->perm, (int)h->N, h->x);
  else
    OM_LAUNCH(h, k_import_points<3>, G, B, flat, h->perm, (int)h->N, h->x);
  cudaError_t e = cudaStreamSynchronize(h->stream);
  om_free(h, flat);
  CUDA_TRY(e);
  return OM_OK;
}

int om_get_cells(om_handle* h, void* out_host, int itemsize) {
  OM_ENTER(h);
  if (itemsize != 4 && itemsize != 8) {
    om_set_error("itemsize must be 4 or 8");
    return OM_ERR_ARG;
  }
  if (h->C == 0) return OM_OK;
  const size_t n = (size_t)3 * h->C;
  const int B = 256, G = om_grid(h->C, B);
  if (is_pinned_host(out_host)) {
    // pinned destination (om_result_alloc): widened on the device, one DMA, no staging
    void* wide = nullptr;
    CUDA_TRY(om_malloc(h, &wide, (size_t)itemsize * n));
    if (itemsize == 4)
      OM_LAUNCH(h, k_export_cells<int>, G, B, h->cells, h->perm, (int)h->C, (int*)wide);
    else
      OM_LAUNCH(h, k_export_cells<long long>, G, B, h->cells, h->perm, (int)h->C,
                (long long*)wide);
    int rc = OM_OK;
    if (cudaMemcpyAsync(out_host, wide, (size_t)itemsize * n, cudaMemcpyDeviceToHost,
                        h->stream) != cudaSuccess ||
        cudaStreamSynchronize(h->stream) != cudaSuccess) {
      om_set_error("copy of cells failed: %s", cudaGetErrorString(cudaGetLastError()));
      rc = OM_ERR_CUDA;
    }
    om_free(h, wide);
    return rc;
  }
  int* flat = nullptr;
  CUDA_TRY(om_malloc(h, &flat, sizeof(int) * n));
  OM_LAUNCH(h, k_export_cells<int>, G, B, h->cells, h->perm, (int)h->C, flat);
  // 4 bytes per index cross the bus; 64-bit output is widened by the host threads
  int rc = itemsize == 4 ? staged_d2h<int, int>(h, flat, (int*)out_host, n)
                         : staged_d2h<int, long long>(h, flat, (long long*)out_host, n);
  om_free(h, flat);
  return rc;
}

int om_get_boundary_flags(om_handle* h, uint8_t* out_host) {
  OM_ENTER(h);
  if (h->N == 0) return OM_OK;
  uint8_t* flat = nullptr;
  CUDA_TRY(om_malloc(h, &flat, h->N));
  OM_LAUNCH(h, k_export_flags, om_grid(h->N, 256), 256, h->bflag, h->perm, (int)h->N, flat);
  cudaMemcpyAsync(out_host, flat, h->N, cudaMemcpyDeviceToHost, h->stream);
  cudaError_t e = cudaStreamSynchronize(h->stream);
  om_free(h, flat);
  CUDA_TRY(e);
  return OM_OK;
}

int om_device_ptrs(om_handle* h, double** points, int32_t** cells4, int32_t** perm,
                   int32_t* point_stride) {
  OM_ENTER(h);
  h->delaunay_clean = false;  // the caller may write through the pointer
  if (points) *points = h->x;
  if (cells4) *cells4 = (int32_t*)h->cells;
  if (perm) *perm = h->perm;
  if (point_stride) *point_stride = h->PD;
  return OM_OK;
}

int om_pack_points(om_handle* h, const int32_t* idx_dev, int64_t n, double* buf_dev) {
  OM_ENTER(h);
  if (n == 0) return OM_OK;
  if (h->D == 2)
    OM_LAUNCH(h, k_pack<2>, om_grid(n, 256), 256, h->x, h->inv_perm, idx_dev, n, buf_dev);
  else
    OM_LAUNCH(h, k_pack<3>, om_grid(n, 256), 256, h->x, h->inv_perm, idx_dev, n, buf_dev);
  CUDA_TRY(cudaGetLastError());
  return OM_OK;
}

int om_unpack_points(om_handle* h, const int32_t* idx_dev, int64_t n, const double* buf_dev) {
  OM_ENTER(h);
  if (n == 0) return OM_OK;
  h->delaunay_clean = false;
  if (h->D == 2)
    OM_LAUNCH(h, k_unpack<2>, om_grid(n, 256), 256, h->x, h->inv_perm, idx_dev, n, buf_dev);
  else
    OM_LAUNCH(h, k_unpack<3>, om_grid(n, 256), 256, h->x, h->inv_perm, idx_dev, n, buf_dev);
  CUDA_TRY(cudaGetLastError());
  return OM_OK;
}

int om_pin_vertices(om_handle* h, const int32_t* idx_host, int64_t n) {
  OM_ENTER(h);
  if (n == 0) return OM_OK;
  if (n < 0 || !idx_host) {
    om_set_error("om_pin_vertices: bad arguments");
    return OM_ERR_ARG;
  }
  for (int64_t i = 0; i < n; i++)
    if (idx_host[i] < 0 || idx_host[i] >= h->N) {
      om_set_error("om_pin_vertices: vertex %d outside [0, %lld)", (int)idx_host[i],
                   (long long)h->N);
      return OM_ERR_INDEX;
    }
  int* d = nullptr;
  CUDA_TRY(om_malloc(h, &d, sizeof(int) * n));
  cudaMemcpyAsync(d, idx_host, sizeof(int) * n, cudaMemcpyHostToDevice, h->stream);
  OM_LAUNCH(h, k_pin, om_grid(n, 256), 256, h->bflag, h->inv_perm, d, n);
  cudaError_t e = cudaStreamSynchronize(h->stream);
  om_free(h, d);
  CUDA_TRY(e);
  h->nbr_valid = false;
  OM_TRY(om_rebuild_rings(h, true));  // pinned vertices have no ring row
  return OM_OK;
}

int om_flip_check_range(om_handle* h, double tol, int64_t cell_lo, int64_t cell_hi,
                        int64_t* n_records, void** records_dev) {
  OM_ENTER(h);
  if (cell_lo < 0 || cell_hi < cell_lo || cell_hi > h->C) {
    om_set_error("cell range [%lld, %lld) outside [0, %lld]", (long long)cell_lo,
                 (long long)cell_hi, (long long)h->C);
    return OM_ERR_ARG;
  }
  OM_TRY(om_flip_check_range_impl(h, tol, cell_lo, cell_hi, n_records));
  if (records_dev) *records_dev = (void*)h->recs;
  return OM_OK;
}

int om_flip_add_records(om_handle* h, const void* records_dev, int64_t n) {
  OM_ENTER(h);
  return om_flip_add_records_impl(h, records_dev, n);
}

int om_flip_finish(om_handle* h, double tol, int max_rounds, int64_t* n_flips, int32_t* n_rounds,
                   int32_t* cap_hit) {
  OM_ENTER(h);
  return om_flip_impl(h, tol, max_rounds, n_flips, n_rounds, cap_hit, true);
}

__global__ void k_lower_bound_cells(const int4* __restrict__ cells, int C, int key, int* out) {
  // first cell whose smallest vertex is >= key (cells are sorted by it after setup)
  int lo = 0, hi = C;
  while (lo < hi) {
    const int mid = lo + (hi - lo) / 2;
    const int4 c = cells[mid];
    if (min(c.x, min(c.y, c.z)) < key)
      lo = mid + 1;
    else
      hi = mid;
  }
  *out = lo;
}

int om_cell_range_of_vertices(om_handle* h, int64_t vertex_lo, int64_t vertex_hi, int64_t* cell_lo,
                              int64_t* cell_hi) {
  OM_ENTER(h);
  int* d = nullptr;
  CUDA_TRY(om_malloc(h, &d, 2 * sizeof(int)));
  k_lower_bound_cells<<<1, 1, 0, h->stream>>>(h->cells, (int)h->C, (int)vertex_lo, d);
  k_lower_bound_cells<<<1, 1, 0, h->stream>>>(h->cells, (int)h->C, (int)vertex_hi, d + 1);
  int r[2] = {0, 0};
  CUDA_TRY(cudaMemcpyAsync(r, d, sizeof(r), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  om_free(h, d);
  if (cell_lo) *cell_lo = r[0];
  if (cell_hi) *cell_hi = r[1];
  return OM_OK;
}

int om_flip_pass_begin(om_handle* h) {
  OM_ENTER(h);
  return om_flip_pass_begin_impl(h);
}

int om_flip_round_check(om_handle* h, double tol, int first, int64_t cell_lo, int64_t cell_hi,
                        int64_t* n_records, void** records_dev, int32_t* stale) {
  OM_ENTER(h);
  if (cell_lo < 0 || cell_hi < cell_lo || cell_hi > h->C) {
    om_set_error("cell range [%lld, %lld) outside [0, %lld]", (long long)cell_lo,
                 (long long)cell_hi, (long long)h->C);
    return OM_ERR_ARG;
  }
  // n_records == NULL: leave the counts on the device (fixed-capacity exchange follows)
  OM_TRY(om_flip_round_check_impl(h, tol, first, cell_lo, cell_hi, n_records, stale,
                                  n_records != nullptr));
  if (records_dev) *records_dev = (void*)h->recs;
  return OM_OK;
}

int om_flip_round_pack(om_handle* h, int32_t capacity, void* slot_dev) {
  OM_ENTER(h);
  if (capacity < 1 || !slot_dev) {
    om_set_error("om_flip_round_pack: bad capacity or buffer");
    return OM_ERR_ARG;
  }
  return om_flip_round_pack_impl(h, capacity, slot_dev);
}

int om_flip_round_apply_gathered(om_handle* h, const void* gathered_dev, int32_t n_ranks,
                                 int32_t capacity, int64_t* n_candidates, int64_t* n_flips_total,
                                 int32_t* abort_bits, int64_t* own_records) {
  OM_ENTER(h);
  return om_flip_round_apply_gathered_impl(h, gathered_dev, n_ranks, capacity, n_candidates,
                                           n_flips_total, abort_bits, own_records);
}

int om_flip_round_apply(om_handle* h, int64_t total_records, int64_t* n_candidates,
                        int64_t* n_flips_total) {
  OM_ENTER(h);
  return om_flip_round_apply_impl(h, total_records, n_candidates, n_flips_total);
}

int om_flip_pass_end(om_handle* h, int64_t* n_flips, int32_t* n_rounds) {
  OM_ENTER(h);
  return om_flip_pass_end_impl(h, n_flips, n_rounds);
}

int om_band_build(om_handle* h, int depth, int64_t* n, int32_t** idx_dev) {
  OM_ENTER(h);
  if (depth < 1 || depth > 200) {
    om_set_error("band depth %d out of range", depth);
    return OM_ERR_ARG;
  }
  OM_TRY(om_band_build_impl(h, depth, n));
  if (idx_dev) *idx_dev = (int32_t*)h->band;
  return OM_OK;
}

int om_band_pack(om_handle* h, const int32_t* idx_dev, int64_t n, double* buf_dev) {
  OM_ENTER(h);
  return om_band_pack_impl(h, (const int*)idx_dev, n, buf_dev);
}

int om_band_unpack(om_handle* h, const int32_t* idx_dev, int64_t n, const double* buf_dev) {
  OM_ENTER(h);
  return om_band_unpack_impl(h, (const int*)idx_dev, n, buf_dev);
}

int om_coords_invalidate(om_handle* h) {
  OM_ENTER(h);
  OM_TRY(om_band_alloc(h));
  h->valid_stamp++;
  h->all_valid = false;
  return OM_OK;
}

int om_set_deferred_commit(om_handle* h, int on) {
  OM_ENTER(h);
  h->defer_commit = on != 0;
  return OM_OK;
}

int om_commit_points(om_handle* h) {
  OM_ENTER(h);
  h->delaunay_clean = false;
  return om_commit_points_impl(h);
}

int om_coords_all_valid(om_handle* h) {
  OM_ENTER(h);
  h->all_valid = true;
  return OM_OK;
}

int om_set_owned_range(om_handle* h, int64_t lo, int64_t hi) {
  OM_ENTER(h);
  if (hi >= 0 && (lo < 0 || lo > hi || hi > h->N)) {
    om_set_error("owned range [%lld, %lld) outside [0, %lld]", (long long)lo, (long long)hi,
                 (long long)h->N);
    return OM_ERR_ARG;
  }
  const bool changed = (hi >= 0 ? lo : 0) != h->own_lo || hi != h->own_hi;
  h->own_lo = hi >= 0 ? lo : 0;
  h->own_hi = hi;
  // rows outside the previous range were not maintained: make all of them current again
  if (changed && h->rings_partial) OM_TRY(om_rebuild_rings(h, true));
  return OM_OK;
}

int om_points_device(om_handle* h, double** points, int64_t* n_alloc, int32_t* stride) {
  OM_ENTER(h);
  h->delaunay_clean = false;  // the caller may write through the pointer
  if (points) *points = h->x;
  if (n_alloc) *n_alloc = h->N + OM_POINT_PAD;
  if (stride) *stride = h->PD;
  return OM_OK;
}

int om_set_timing(om_handle* h, int on) {
  OM_ENTER(h);
  if (on && !h->ev[0])
    for (int i = 0; i < 4; i++) CUDA_TRY(cudaEventCreate(&h->ev[i]));
  h->timing = on != 0;
  h->t_step_ms = h->t_flip_ms = 0.0;
  h->n_step = h->n_flip = 0;
  for (double& t : h->t_phase_ms) t = 0.0;
  h->n_phase = 0;
  return OM_OK;
}

int om_get_timing(om_handle* h, double* step_kernel_ms, int64_t* step_kernel_launches,
                  double* flip_pass_ms, int64_t* flip_passes) {
  OM_ENTER(h);
  if (step_kernel_ms) *step_kernel_ms = h->t_step_ms;
  if (step_kernel_launches) *step_kernel_launches = h->n_step;
  if (flip_pass_ms) *flip_pass_ms = h->t_flip_ms;
  if (flip_passes) *flip_passes = h->n_flip;
  return OM_OK;
}

int om_get_phase_timing(om_handle* h, double* phase_ms5, int64_t* iterations) {
  OM_ENTER(h);
  if (phase_ms5)
    for (int i = 0; i < 5; i++) phase_ms5[i] = h->t_phase_ms[i];
  if (iterations) *iterations = h->n_phase;
  return OM_OK;
}

int om_release_cached_memory(int device) {
  cudaMemPool_t pool;
  CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, device));
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemPoolTrimTo(pool, 0));
  {  // the cached pinned staging buffers of that device too
    std::lock_guard<std::mutex> lock(g_stage_mutex);
    for (size_t i = 0; i < g_stage_cache.size();) {
      if (g_stage_cache[i].first == device) {
        delete g_stage_cache[i].second;
        g_stage_cache.erase(g_stage_cache.begin() + i);
      } else {
        i++;
      }
    }
  }
  return OM_OK;
}

// ---- result buffers in pinned host memory, kept for the life of the process.
// A result copied into a fresh pageable array costs a pinned staging hop, a host copy and
// one page fault per 4 KB; into a cached pinned block it is one DMA at PCIe speed.  The
// Python layer wraps a block as the numpy array it returns and hands it back when that
// array is garbage collected.
namespace {
std::mutex g_res_mutex;
std::vector<std::pair<void*, size_t>> g_res_free;  // blocks not in use
size_t g_res_total = 0;
const size_t RES_LIMIT = (size_t)4 << 30;  // pinned bytes the cache may hold
}  // namespace

int om_result_alloc(int64_t bytes, void** out) {
  if (!out) return OM_ERR_ARG;
  *out = nullptr;
  if (bytes <= 0) return OM_OK;
  const size_t want = ((size_t)bytes + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
  {
    std::lock_guard<std::mutex> lock(g_res_mutex);
    for (size_t i = 0; i < g_res_free.size(); i++)
      if (g_res_free[i].second == want) {
        *out = g_res_free[i].first;
        g_res_free.erase(g_res_free.begin() + i);
        return OM_OK;
      }
    if (g_res_total + want > RES_LIMIT) return OM_OK;  // caller falls back to pageable memory
    g_res_total += want;
  }
  void* p = nullptr;
  if (cudaHostAlloc(&p, want, cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    std::lock_guard<std::mutex> lock(g_res_mutex);
    g_res_total -= want;
    return OM_OK;
  }
  *out = p;
  return OM_OK;
}

int om_result_free(void* p, int64_t bytes) {
  if (!p) return OM_OK;
  const size_t want = ((size_t)bytes + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
  std::lock_guard<std::mutex> lock(g_res_mutex);
  g_res_free.emplace_back(p, want);
  return OM_OK;
}

int om_prefault_host(void* p, int64_t bytes) {
  // Fresh result arrays cost one page fault per 4 KB when they are first written (155k
  // faults for the 637 MB of a 10M-vertex result): ask for huge pages and touch every page
  // from all host threads, while the GPU is busy with the steps.
  if (!p || bytes <= 0) return OM_OK;
  const size_t page = (size_t)sysconf(_SC_PAGESIZE);
  char* b = (char*)p;
  char* e = b + bytes;
  char* ab = (char*)(((uintptr_t)b + page - 1) & ~(uintptr_t)(page - 1));
  char* ae = (char*)((uintptr_t)e & ~(uintptr_t)(page - 1));
#ifdef MADV_HUGEPAGE
  static const bool huge = getenv("OM_NO_HUGEPAGE") == nullptr;
  if (huge && ae > ab) madvise(ab, (size_t)(ae - ab), MADV_HUGEPAGE);
#endif
  const size_t npages = ((size_t)bytes + page - 1) / page;
  parallel_chunks(std::max<size_t>(npages, (size_t)1 << 16), [=](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi && i < npages; i++) {
      volatile char* q = b + i * page;
      if ((char*)q < e) *q = 0;
    }
  });
  return OM_OK;
}

int om_launch_count(om_handle* h, int64_t* n) {
  OM_ENTER(h);
  if (n) *n = h->launches;
  return OM_OK;
}

int om_synchronize(om_handle* h) {
  OM_ENTER(h);
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return OM_OK;
}

int om_stream(om_handle* h, void** stream) {
  OM_ENTER(h);
  if (stream) *stream = (void*)h->stream;
  return OM_OK;
}

}  // extern "C"
