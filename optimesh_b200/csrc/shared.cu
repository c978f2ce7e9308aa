// One mesh in ONE address space over the GPUs of a box (NVLink 5 / NVSwitch peer memory).
//
// the reference is a single process; this is the decomposition of the
// optimize() loop (/root/reference/README.md:131-132) over N GPUs, one process per GPU.
//
// Every mesh array (points, cells, twins, ring rows, flip scratch) is cut into N chunks by
// vertex / cell id; chunk r is physical memory of GPU r (cuMemCreate), exported as a POSIX file
// descriptor, imported by every other rank and mapped -- all N chunks back to back -- into one
// reserved virtual address range per array (CUDA virtual memory management).  So every rank
// sees the WHOLE mesh under the same global ids, 1/N of it resident locally, the rest one
// NVSwitch hop away (650 GB/s measured for peer reads), and the kernels of the single-GPU
// pipeline (loop.cu) run unchanged: each rank launches them on its own vertex range and its
// own work lists, and whatever they touch across a chunk boundary -- the ring of a vertex
// next to the cut, the cell across a flipped edge, the stamp that dedupes a work list -- is
// an ordinary load, store or atomic on peer memory.  Memory per rank scales with 1/N; no halo
// buffers, no pack/unpack, no NCCL on the data path.
//
// What the GPUs must agree on is ORDER.  Between the phases of a step where one rank reads
// what another wrote (candidate marks -> selection -> flip -> twin patch -> next check; new
// points -> next update) the ranks meet in k_sync: a one-thread kernel that adds this rank's
// counters into every rank's control block with system-scope atomics (an all-reduce through
// peer memory), bumps every rank's arrival counter and spins on its own until all have
// arrived -- about 9 us per meeting on two B200s, launch gap included.  The reduced values
// (candidates left, flips done, max |diff|^2, limited vertices) drive the same device-side
// control as on one GPU (flip-round WHILE node, step WHILE node, limiter variant IF nodes), so
// all ranks take the same branches and the whole loop is one CUDA graph per rank: the host
// launches it once.  Flips are executed by whichever rank enlisted the cell (the stamps in
// shared memory make that exactly one), i.e. an owner-computes rule decided by an atomic.
//
// Setup (this round): every rank still builds the complete mesh on its own GPU first
// (om_create), then keeps its chunks; a distributed setup (sample sort of the edge keys) is
// the missing piece for meshes that do not fit one GPU during setup.
#include <cuda.h>
#include <unistd.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

#include "common.cuh"

namespace {

// ---- driver entry points (resolved at run time: no link-time dependency on libcuda)
struct Drv {
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*,
                        unsigned long long) = nullptr;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemExport)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType,
                        unsigned long long) = nullptr;
  CUresult (*MemImport)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr,
                                unsigned long long) = nullptr;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle,
                     unsigned long long) = nullptr;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*MemGetGranularity)(size_t*, const CUmemAllocationProp*,
                                CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  bool ok = false;
};

template <typename F>
bool load_sym(const char* name, F* fn) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult st;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &st) != cudaSuccess || !p) {
    cudaGetLastError();
    return false;
  }
  *fn = (F)p;
  return true;
}

Drv& drv() {
  static Drv d = [] {
    Drv x;
    x.ok = load_sym("cuMemCreate", &x.MemCreate) && load_sym("cuMemRelease", &x.MemRelease) &&
           load_sym("cuMemExportToShareableHandle", &x.MemExport) &&
           load_sym("cuMemImportFromShareableHandle", &x.MemImport) &&
           load_sym("cuMemAddressReserve", &x.MemAddressReserve) &&
           load_sym("cuMemAddressFree", &x.MemAddressFree) && load_sym("cuMemMap", &x.MemMap) &&
           load_sym("cuMemUnmap", &x.MemUnmap) && load_sym("cuMemSetAccess", &x.MemSetAccess) &&
           load_sym("cuMemGetAllocationGranularity", &x.MemGetGranularity) &&
           load_sym("cuGetErrorString", &x.GetErrorString);
    return x;
  }();
  return d;
}

#define DRV_TRY(expr)                                                                  \
  do {                                                                                 \
    CUresult _r = (expr);                                                              \
    if (_r != CUDA_SUCCESS) {                                                          \
      const char* _s = "?";                                                            \
      drv().GetErrorString(_r, &_s);                                                   \
      om_set_error("driver error at %s:%d: %s (%s)", __FILE__, __LINE__, _s, #expr);   \
      return OM_ERR_CUDA;                                                              \
    }                                                                                  \
  } while (0)

constexpr int SH_SLOTS = 64;
struct ShSlot {
  unsigned long long sum[4];
  unsigned long long mx[2];
};
struct ShCtrl {
  unsigned long long arrive;
  unsigned long long pad[7];
  ShSlot slots[SH_SLOTS];
};

struct ShArray {
  void** target;        // the handle's pointer that is set to the mapped range
  const void* source;   // the same array of the complete single-GPU handle
  size_t elem;          // bytes per vertex / cell
  bool per_cell;
  CUdeviceptr base = 0;
  size_t chunk = 0;     // bytes per rank
  CUmemGenericAllocationHandle own = 0;
  std::vector<CUmemGenericAllocationHandle> peers;
  int fd = -1;
};

struct om_shared {
  int rank = 0, world = 1, device = 0;
  int64_t nvc = 0, ncc = 0;  // vertices / cells per chunk
  int vlo = 0, vhi = 0;
  std::vector<ShArray> arrays;
  ShArray ctrl;
  char* ctrl_base = nullptr;
  size_t ctrl_stride = 0;
  om_handle* full = nullptr;  // the complete handle the chunks are copied from (until mapped)
  bool mapped = false;
  // the loop's graph (one per parity of the start buffer)
  cudaGraph_t graph[2] = {nullptr, nullptr};
  cudaGraphExec_t exec[2] = {nullptr, nullptr};
  const double* graph_a[2] = {nullptr, nullptr};
  int g_method = -1, g_limiter = -1, g_odt = -1;
  double g_omega = 0.0;
  cudaStream_t capture_stream = nullptr;
};

struct ShView {
  char* ctrl_base;
  size_t ctrl_stride;
  int me, world;
  __device__ __forceinline__ ShCtrl* of(int r) const {
    return reinterpret_cast<ShCtrl*>(ctrl_base + ctrl_stride * r);
  }
};

// what a meeting reduces
enum { SYNC_BARRIER = 0, SYNC_ROUND = 1, SYNC_STATS = 2 };

// The ranks meet: all-reduce of this rank's counters through every rank's control block, then
// a barrier on the arrival counters.  One thread; the kernel boundary before it has completed
// this rank's earlier kernels, the system-scope fences order its peer writes before the
// arrival is visible.
__device__ void sh_iter_end(DevScalars* ds, cudaGraphConditionalHandle handle);

__global__ void k_sync(ShView sv, DevScalars* ds, int kind, cudaGraphConditionalHandle handle) {
  // one warp: lane r talks to rank r, so the peer round trips overlap instead of adding up
  const int lane = threadIdx.x;
  if (ds->halt == 3 && ds->sync_dead) {
    // a meeting already timed out: do not wait again, and let the loops of the graph end
    if (lane == 0 && kind != SYNC_BARRIER) cudaGraphSetConditional(handle, 0u);
    return;
  }
  const unsigned long long s = ds->sync_seq + 1ull;
  const int slot = (int)(s % SH_SLOTS);
  unsigned long long sum[4] = {0ull, 0ull, 0ull, 0ull}, mx[2] = {0ull, 0ull};
  if (kind == SYNC_ROUND) {
    sum[0] = (unsigned long long)ds->n_cand;
    sum[1] = (unsigned long long)ds->n_flips;
    mx[1] = (unsigned long long)ds->err;
  } else if (kind == SYNC_STATS) {
    sum[0] = ds->n_limited;
    sum[1] = (unsigned long long)ds->n_deferred;
    sum[2] = (unsigned long long)ds->n_dirty;
    mx[0] = ds->max_diff2_bits;
    mx[1] = (unsigned long long)ds->err;
  }
  __threadfence_system();
  for (int r = lane; r < sv.world; r += 32) {
    ShCtrl* c = sv.of(r);
    if (kind != SYNC_BARRIER) {
#pragma unroll
      for (int i = 0; i < 4; i++)
        if (sum[i]) atomicAdd_system(&c->slots[slot].sum[i], sum[i]);
#pragma unroll
      for (int i = 0; i < 2; i++)
        if (mx[i]) atomicMax_system(&c->slots[slot].mx[i], mx[i]);
      __threadfence_system();  // the values are in place before the arrival is
    }
    atomicAdd_system(&c->arrive, 1ull);
  }
  __syncwarp();
  if (lane != 0) return;
  ShCtrl* mine = sv.of(sv.me);
  const unsigned long long target = (unsigned long long)sv.world * s;
  volatile unsigned long long* arrive = &mine->arrive;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (*arrive < target) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 20000000000ull) {  // 20 s: a rank is gone; stop instead of hanging the GPU
      ds->halt = 3;
      ds->sync_dead = 1;
      ds->err |= OM_DEV_WALK;
      break;
    }
  }
  __threadfence_system();
  if (kind != SYNC_BARRIER) {
    volatile ShSlot* sl = &mine->slots[slot];
    ds->g_sum[0] = sl->sum[0];
    ds->g_sum[1] = sl->sum[1];
    ds->g_sum[2] = sl->sum[2];
    ds->g_sum[3] = sl->sum[3];
    ds->g_max[0] = sl->mx[0];
    ds->g_max[1] = sl->mx[1];
  }
  // the slot half a ring ahead was last used 32 meetings ago: everyone has read it
  ShSlot* ahead = &mine->slots[(slot + SH_SLOTS / 2) % SH_SLOTS];
  ahead->sum[0] = ahead->sum[1] = ahead->sum[2] = ahead->sum[3] = 0ull;
  ahead->mx[0] = ahead->mx[1] = 0ull;
  ds->sync_seq = s;
  ds->pl_launches += 1;
  if (kind == SYNC_ROUND) {
    // ends the check of a flip round with the GLOBAL counts: the same decision on every rank
    unsigned go = 0u;
    if (!(ds->halt & 1)) {
      if (ds->g_max[1]) ds->err |= (int)ds->g_max[1];
      // kernels since the last decision (meetings count themselves): pass begin, flag check /
      // flip1, flip2, round begin, check
      ds->pl_launches += ds->pl_round > 0 ? 4 : 2;
      go = om_flip_decide(ds, ds->g_sum[0], ds->g_sum[1]);
    }
    cudaGraphSetConditional(handle, go);
  } else if (kind == SYNC_STATS) {
    sh_iter_end(ds, handle);
  }
}

__global__ void k_sh_init(DevScalars* ds, long long max_steps, double tol2, int mode_exact,
                          long long n_free, int limiter_on, int max_rounds, int div_exact) {
  ds->halt = 0;
  ds->lim_div_exact = div_exact;
  ds->k = 0;
  ds->max_steps = max_steps;
  ds->tol2 = tol2;
  ds->mode_exact = mode_exact;
  ds->n_free = n_free;
  ds->limiter_on = limiter_on;
  ds->max_rounds = max_rounds;
  ds->cap_hit = 0;
  ds->pl_go = 0;
  ds->total_flips = ds->total_rounds = ds->total_limited = 0;
  ds->total_deferred = 0;
  ds->pl_launches = 0;
  ds->g_flips_prev = 0;
}

// ends an iteration with the GLOBAL statistics of the update
__device__ void sh_iter_end(DevScalars* ds, cudaGraphConditionalHandle handle) {
  if (ds->halt == 4) {
    // the flush iteration: the flip pass of the last points is done
    ds->n_flips = (int)ds->g_flips;
    ds->total_flips += (long long)ds->g_flips;
    ds->total_rounds += ds->n_rounds;
    ds->halt = (ds->err || ds->g_max[1]) ? 3 : 1;
  } else if (!ds->halt) {
    ds->k++;
    ds->total_flips += (long long)ds->g_flips;
    ds->total_rounds += ds->n_rounds;
    const long long limited = (long long)ds->g_sum[0];
    ds->total_limited += limited;
    ds->total_deferred += (long long)ds->g_sum[1];
    ds->max_diff2_bits = ds->g_max[0];
    ds->n_limited = (unsigned long long)limited;
    ds->mode_exact = om_limiter_mode(ds->limiter_on != 0, limited, ds->n_free, ds->lim_div_exact);
    double md;
    memcpy(&md, &ds->g_max[0], 8);
    if (ds->err || ds->g_max[1])
      ds->halt = 3;
    else if (md < ds->tol2 || ds->k >= ds->max_steps)
      ds->halt = 4;  // one more pass of the loop body: only its flip pass runs
  }
  // reset, variant selection, ring kernel, post, ring rows, fix-up, reduce
  ds->pl_launches += 7;
  cudaGraphSetConditional(handle, (ds->halt & 1) ? 0u : 1u);
}

__global__ void k_sh_mode(const DevScalars* ds, cudaGraphConditionalHandle lazy,
                          cudaGraphConditionalHandle exact) {
  const bool run = !(ds->halt & 1);
  cudaGraphSetConditional(lazy, run && ds->mode_exact != 1 ? 1u : 0u);
  cudaGraphSetConditional(exact, run && ds->mode_exact == 1 ? 1u : 0u);
}

om_shared* shared_of(om_handle* h) { return (om_shared*)h->sh; }

void add_array(om_shared* sh, void** target, const void* source, size_t elem, bool per_cell) {
  ShArray a;
  a.target = target;
  a.source = source;
  a.elem = elem;
  a.per_cell = per_cell;
  sh->arrays.push_back(a);
}

int create_chunk(om_shared* sh, ShArray& a) {
  CUmemAllocationProp prop;
  memset(&prop, 0, sizeof(prop));
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = sh->device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  DRV_TRY(drv().MemCreate(&a.own, a.chunk, &prop, 0));
  DRV_TRY(drv().MemExport(&a.fd, a.own, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  return OM_OK;
}

int map_array(om_shared* sh, ShArray& a, const int* fds_of_rank, int stride, int index) {
  a.peers.assign(sh->world, 0);
  DRV_TRY(drv().MemAddressReserve(&a.base, a.chunk * sh->world, 0, 0, 0));
  for (int r = 0; r < sh->world; r++) {
    CUmemGenericAllocationHandle hd = a.own;
    if (r != sh->rank) {
      const int fd = fds_of_rank[(size_t)r * stride + index];
      DRV_TRY(drv().MemImport(&a.peers[r], (void*)(uintptr_t)fd,
                              CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
      hd = a.peers[r];
    }
    DRV_TRY(drv().MemMap(a.base + a.chunk * r, a.chunk, 0, hd, 0));
  }
  CUmemAccessDesc acc;
  memset(&acc, 0, sizeof(acc));
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = sh->device;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  DRV_TRY(drv().MemSetAccess(a.base, a.chunk * sh->world, &acc, 1));
  return OM_OK;
}

void unmap_array(om_shared* sh, ShArray& a) {
  if (a.base) {
    drv().MemUnmap(a.base, a.chunk * sh->world);
    drv().MemAddressFree(a.base, a.chunk * sh->world);
    a.base = 0;
  }
  for (auto p : a.peers)
    if (p) drv().MemRelease(p);
  a.peers.clear();
  if (a.own) drv().MemRelease(a.own);
  a.own = 0;
}

#define CU_TRY2(expr)                                                                 \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      om_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__,      \
                   __LINE__, cudaGetErrorString(_e));                                 \
      return OM_ERR_CUDA;                                                             \
    }                                                                                 \
  } while (0)

ShView view_of(om_shared* sh) {
  ShView v;
  v.ctrl_base = sh->ctrl_base;
  v.ctrl_stride = sh->ctrl_stride;
  v.me = sh->rank;
  v.world = sh->world;
  return v;
}

int sync_point(om_handle* h, int kind, cudaGraphConditionalHandle handle = 0) {
  OM_LAUNCH(h, k_sync, 1, 32, view_of(shared_of(h)), h->ds, kind, handle);
  CUDA_TRY(cudaGetLastError());
  return OM_OK;
}

// One flip round over all GPUs: the flips of the last check, the twin patch and the check of
// the cells they touched, with the meetings the single-GPU order gets from kernel boundaries.
// Ends in the meeting that sums candidates and flips and decides about the next round.
int round_with_meetings(om_handle* h, const double* xin, cudaGraphConditionalHandle inner) {
  OM_TRY(om_pl_launch_flips_part(h, 1));
  OM_TRY(sync_point(h, SYNC_BARRIER));  // every flip is recorded
  OM_TRY(om_pl_launch_flips_part(h, 2));
  OM_TRY(sync_point(h, SYNC_BARRIER));  // every twin is patched
  OM_TRY(om_pl_launch_round_check(h, xin));
  OM_TRY(sync_point(h, SYNC_ROUND, inner));  // every mark of the check is written
  return OM_OK;
}

template <typename F>
int capture_into(om_handle* h, cudaStream_t cs, cudaGraph_t graph, const cudaGraphNode_t* deps,
                 size_t ndeps, cudaGraphNode_t* last, F&& body) {
  CU_TRY2(cudaStreamBeginCaptureToGraph(cs, graph, deps, nullptr, ndeps,
                                        cudaStreamCaptureModeRelaxed));
  cudaStream_t keep = h->stream;
  h->stream = cs;
  const int64_t launches = h->launches;
  int rc = body();
  h->stream = keep;
  h->launches = launches;
  cudaStreamCaptureStatus st;
  const cudaGraphNode_t* leaf = nullptr;
  size_t nleaf = 0;
  cudaError_t e = cudaStreamGetCaptureInfo_v2(cs, &st, nullptr, nullptr, &leaf, &nleaf);
  cudaGraphNode_t tail = (e == cudaSuccess && nleaf > 0) ? leaf[nleaf - 1] : nullptr;
  const bool single = nleaf == 1;
  cudaGraph_t out = nullptr;
  cudaError_t e2 = cudaStreamEndCapture(cs, &out);
  if (rc != OM_OK) return rc;
  CU_TRY2(e);
  CU_TRY2(e2);
  if (!single || !tail) {
    om_set_error("graph capture of the shared loop did not end in a single node");
    return OM_ERR_CUDA;
  }
  *last = tail;
  return OM_OK;
}

int add_cond(cudaGraph_t parent, cudaGraphConditionalHandle handle, cudaGraphConditionalNodeType t,
             const cudaGraphNode_t* deps, size_t ndeps, cudaGraphNode_t* node, cudaGraph_t* body) {
  cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
  np.type = cudaGraphNodeTypeConditional;
  np.conditional.handle = handle;
  np.conditional.type = t;
  np.conditional.size = 1;
  CU_TRY2(cudaGraphAddNode(node, parent, deps, ndeps, &np));
  *body = np.conditional.phGraph_out[0];
  return OM_OK;
}

int build_graph(om_handle* h, om_shared* sh, int which, double* A, double* B) {
  if (!sh->capture_stream)
    CU_TRY2(cudaStreamCreateWithFlags(&sh->capture_stream, cudaStreamNonBlocking));
  cudaStream_t cs = sh->capture_stream;
  CU_TRY2(cudaGraphCreate(&sh->graph[which], 0));
  cudaGraph_t g = sh->graph[which];
  cudaGraphConditionalHandle outer;
  CU_TRY2(cudaGraphConditionalHandleCreate(&outer, g, 1u, cudaGraphCondAssignDefault));
  // The ranks meet once before the loop: whatever they enqueued before (set-up, the flip pass
  // of the start mesh) is complete everywhere.  Inside the loop the meeting that ends an
  // iteration is also the one the next update waits for.
  cudaGraphNode_t first;
  OM_TRY(capture_into(h, cs, g, nullptr, 0, &first,
                      [&]() -> int { return sync_point(h, SYNC_BARRIER); }));
  cudaGraphNode_t outer_node;
  cudaGraph_t body;
  OM_TRY(add_cond(g, outer, cudaGraphCondTypeWhile, &first, 1, &outer_node, &body));
  cudaGraphNode_t last = nullptr;
  for (int half = 0; half < 2; half++) {
    const double* xin = half == 0 ? A : B;
    double* xout = half == 0 ? B : A;
    cudaGraphConditionalHandle inner, lazy, exact;
    CU_TRY2(cudaGraphConditionalHandleCreate(&inner, body, 0u, cudaGraphCondAssignDefault));
    CU_TRY2(cudaGraphConditionalHandleCreate(&lazy, body, 0u, cudaGraphCondAssignDefault));
    CU_TRY2(cudaGraphConditionalHandleCreate(&exact, body, 0u, cudaGraphCondAssignDefault));
    OM_TRY(capture_into(h, cs, body, last ? &last : nullptr, last ? 1 : 0, &last, [&]() -> int {
      OM_TRY(om_pl_launch_update_part(h, xin, xout, 0));
      OM_LAUNCH(h, k_sh_mode, 1, 1, (const DevScalars*)h->ds, lazy, exact);
      CUDA_TRY(cudaGetLastError());
      return (int)OM_OK;
    }));
    cudaGraphNode_t if_lazy, if_exact, unused;
    cudaGraph_t sub;
    OM_TRY(add_cond(body, lazy, cudaGraphCondTypeIf, &last, 1, &if_lazy, &sub));
    OM_TRY(capture_into(h, cs, sub, nullptr, 0, &unused,
                        [&] { return om_pl_launch_update_part(h, xin, xout, 1); }));
    OM_TRY(add_cond(body, exact, cudaGraphCondTypeIf, &if_lazy, 1, &if_exact, &sub));
    OM_TRY(capture_into(h, cs, sub, nullptr, 0, &unused,
                        [&] { return om_pl_launch_update_part(h, xin, xout, 2); }));
    last = if_exact;
    OM_TRY(capture_into(h, cs, body, &last, 1, &last, [&]() -> int {
      OM_TRY(om_pl_launch_update_part(h, xin, xout, 3));  // k_post
      OM_TRY(om_pl_launch_flags_check(h, xin));
      OM_TRY(sync_point(h, SYNC_ROUND, inner));  // every mark of the check is written
      return (int)OM_OK;
    }));
    cudaGraphNode_t inner_node;
    cudaGraph_t rounds;
    OM_TRY(add_cond(body, inner, cudaGraphCondTypeWhile, &last, 1, &inner_node, &rounds));
    OM_TRY(capture_into(h, cs, rounds, nullptr, 0, &unused,
                        [&]() -> int { return round_with_meetings(h, xin, inner); }));
    OM_TRY(capture_into(h, cs, body, &inner_node, 1, &last, [&]() -> int {
      OM_TRY(om_pl_launch_tail(h, xin, xout));  // ring rows + recomputation of touched vertices
      // the rank that flipped an edge recomputes its four vertices, whoever owns them: the
      // owners reduce their |diff|^2 only after everybody is done
      OM_TRY(sync_point(h, SYNC_BARRIER));
      OM_TRY(om_launch_reduce_stats(h));
      OM_TRY(sync_point(h, SYNC_STATS, outer));
      return (int)OM_OK;
    }));
  }
  CU_TRY2(cudaGraphInstantiate(&sh->exec[which], g, 0));
  sh->graph_a[which] = A;
  return OM_OK;
}

int graph_for(om_handle* h, om_shared* sh, double* A, double* B, int* which_out);

void free_graphs(om_shared* sh) {
  for (int i = 0; i < 2; i++) {
    if (sh->exec[i]) cudaGraphExecDestroy(sh->exec[i]);
    if (sh->graph[i]) cudaGraphDestroy(sh->graph[i]);
    sh->exec[i] = nullptr;
    sh->graph[i] = nullptr;
    sh->graph_a[i] = nullptr;
  }
}

// the graph that starts from the buffer A (built and cached on demand; settings checked)
int graph_for(om_handle* h, om_shared* sh, double* A, double* B, int* which_out) {
  if (sh->g_method != h->method || sh->g_omega != h->omega || sh->g_limiter != h->limiter ||
      sh->g_odt != h->odt_bary) {
    free_graphs(sh);
    sh->g_method = h->method;
    sh->g_omega = h->omega;
    sh->g_limiter = h->limiter;
    sh->g_odt = h->odt_bary;
  }
  int which = -1;
  for (int i = 0; i < 2; i++)
    if (sh->graph_a[i] == A) which = i;
  if (which < 0) {
    which = sh->graph_a[0] ? 1 : 0;
    if (sh->exec[which]) {
      cudaGraphExecDestroy(sh->exec[which]);
      cudaGraphDestroy(sh->graph[which]);
      sh->exec[which] = nullptr;
      sh->graph[which] = nullptr;
      sh->graph_a[which] = nullptr;
    }
    OM_TRY(build_graph(h, sh, which, A, B));
  }
  *which_out = which;
  return OM_OK;
}

__global__ void k_clear_flags(unsigned short* f, int lo, int hi) {
  const int v = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (v < hi) f[v] = 0;
}

}  // namespace

void om_shared_vertex_range(om_handle* h, int* vlo, int* vhi) {
  om_shared* sh = shared_of(h);
  if (!sh) return;
  *vlo = sh->vlo;
  *vhi = sh->vhi;
}

void om_shared_destroy(om_handle* h) {
  om_shared* sh = shared_of(h);
  if (!sh) return;
  free_graphs(sh);
  if (sh->capture_stream) cudaStreamDestroy(sh->capture_stream);
  for (auto& a : sh->arrays) {
    if (a.fd >= 0) close(a.fd);
    unmap_array(sh, a);
    if (a.target) *a.target = nullptr;  // not the handle's own allocation
  }
  if (sh->ctrl.fd >= 0) close(sh->ctrl.fd);
  unmap_array(sh, sh->ctrl);
  delete sh;
  h->sh = nullptr;
}

extern "C" {

// Step 1 of 2: a new handle whose mesh arrays will live in the shared address space of `world`
// GPUs.  Allocates this rank's chunk of every array and returns their file descriptors
// (n_fds of them, in a fixed order) for the caller to pass to the other ranks.
int om_shared_begin(om_handle* full, int rank, int world, om_handle** out, int32_t* fds,
                    int32_t* n_fds) {
  if (!full || !out || !fds || !n_fds || world < 1 || rank < 0 || rank >= world) {
    om_set_error("om_shared_begin: bad arguments");
    return OM_ERR_ARG;
  }
  if (!drv().ok) {
    om_set_error("the CUDA driver does not provide the virtual memory management entry points");
    return OM_ERR_CUDA;
  }
  if (full->sh || full->own_hi >= 0) {
    om_set_error("om_shared_begin needs an ordinary, complete handle");
    return OM_ERR_ARG;
  }
  cudaSetDevice(full->device);
  CUDA_TRY(cudaStreamSynchronize(full->stream));
  om_handle* h = new om_handle();
  h->device = full->device;
  h->N = full->N;
  h->C = full->C;
  h->D = full->D;
  h->PD = full->PD;
  h->cells_itemsize = full->cells_itemsize;
  h->method = full->method;
  h->omega = full->omega;
  h->limiter = full->limiter;
  h->odt_bary = full->odt_bary;
  h->limited_frac = full->limited_frac;
  h->use_rings = full->use_rings;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete h;
    om_set_error("cudaStreamCreate failed");
    return OM_ERR_CUDA;
  }
  h->own_stream = true;
  om_shared* sh = new om_shared();
  h->sh = sh;
  sh->rank = rank;
  sh->world = world;
  sh->device = full->device;
  sh->full = full;
  // chunks: a multiple of 2^21 vertices, so that a chunk of ANY per-vertex array (down to one
  // byte per vertex) is a whole number of 2 MB pages of the memory manager; two cells per vertex
  const int64_t unit = (int64_t)1 << 21;
  const int64_t per = (h->N + 16 + world - 1) / world;  // (+16: padding of the flag words)
  sh->nvc = std::max<int64_t>((per + unit - 1) / unit, 1) * unit;
  const int64_t perc = (h->C + world - 1) / world;
  sh->ncc = std::max<int64_t>(std::max<int64_t>((perc + unit - 1) / unit, 1) * unit, 2 * sh->nvc);
  sh->vlo = (int)std::min<int64_t>(h->N, sh->nvc * rank);
  sh->vhi = (int)std::min<int64_t>(h->N, sh->nvc * (rank + 1));
  const size_t pd = sizeof(double) * h->PD;
  add_array(sh, (void**)&h->x, full->x, pd, false);
  add_array(sh, (void**)&h->xnew, full->xnew, pd, false);
  add_array(sh, (void**)&h->v2c, full->v2c, 4, false);
  add_array(sh, (void**)&h->bflag, full->bflag, 1, false);
  add_array(sh, (void**)&h->ring, full->ring, 4 * OM_RING_W, false);
  add_array(sh, (void**)&h->ringc, full->ringc, 4 * OM_RING_W, false);
  add_array(sh, (void**)&h->diff2, full->diff2, 8, false);
  add_array(sh, (void**)&h->vflags, full->vflags, 2, false);
  add_array(sh, (void**)&h->dirty_epoch, full->dirty_epoch, 4, false);
  add_array(sh, (void**)&h->cells, full->cells, 16, true);
  add_array(sh, (void**)&h->adj, full->adj, 16, true);
  add_array(sh, (void**)&h->adj_tmp, full->adj_tmp, 16, true);
  add_array(sh, (void**)&h->cand_epoch, full->cand_epoch, 4, true);
  add_array(sh, (void**)&h->work_epoch, full->work_epoch, 4, true);
  add_array(sh, (void**)&h->flip_epoch, full->flip_epoch, 4, true);
  add_array(sh, (void**)&h->sarr, full->sarr, 32, true);
  add_array(sh, (void**)&h->reloc, full->reloc, 16, true);
  int rc = OM_OK;
  for (auto& a : sh->arrays) {
    a.chunk = (size_t)(a.per_cell ? sh->ncc : sh->nvc) * a.elem;
    rc = create_chunk(sh, a);
    if (rc != OM_OK) break;
  }
  if (rc == OM_OK) {
    sh->ctrl.chunk = (size_t)unit;  // one 2 MB page holds the control block
    sh->ctrl.elem = 1;
    rc = create_chunk(sh, sh->ctrl);
  }
  if (rc != OM_OK) {
    om_destroy(h);
    return rc;
  }
  int n = 0;
  for (auto& a : sh->arrays) fds[n++] = a.fd;
  fds[n++] = sh->ctrl.fd;
  *n_fds = n;
  *out = h;
  return OM_OK;
}

// Step 2 of 2: fds_all holds, rank by rank, the n_fds descriptors every rank got from
// om_shared_begin (as received in THIS process; the own row is ignored).  Maps all chunks,
// copies this rank's share of the complete mesh into its chunks and makes the handle usable.
// The caller synchronises the ranks afterwards (and may then destroy the complete handle).
int om_shared_map(om_handle* h, const int32_t* fds_all, int32_t n_fds) {
  om_shared* sh = h ? shared_of(h) : nullptr;
  if (!sh || sh->mapped || !fds_all || n_fds != (int)sh->arrays.size() + 1) {
    om_set_error("om_shared_map: bad arguments");
    return OM_ERR_ARG;
  }
  cudaSetDevice(h->device);
  om_handle* full = sh->full;
  int idx = 0;
  for (auto& a : sh->arrays) {
    OM_TRY(map_array(sh, a, fds_all, n_fds, idx++));
    *a.target = (void*)a.base;
  }
  OM_TRY(map_array(sh, sh->ctrl, fds_all, n_fds, idx));
  sh->ctrl_base = (char*)sh->ctrl.base;
  sh->ctrl_stride = sh->ctrl.chunk;
  CUDA_TRY(cudaMemsetAsync(sh->ctrl_base + sh->ctrl_stride * sh->rank, 0, sizeof(ShCtrl),
                           h->stream));
  // this rank's share of every array (the rest of a chunk stays untouched: ids >= N / C)
  for (auto& a : sh->arrays) {
    const int64_t per = a.per_cell ? sh->ncc : sh->nvc;
    int64_t total = a.per_cell ? h->C : h->N;
    if (a.target == (void**)&h->vflags) total = h->N + 16;
    const int64_t lo = std::min<int64_t>(total, per * sh->rank);
    const int64_t hi = std::min<int64_t>(total, per * (sh->rank + 1));
    if (hi > lo)
      CUDA_TRY(cudaMemcpyAsync((char*)a.base + (size_t)lo * a.elem,
                               (const char*)a.source + (size_t)lo * a.elem,
                               (size_t)(hi - lo) * a.elem, cudaMemcpyDeviceToDevice, h->stream));
  }
  // local arrays: numbering maps, work lists, scalars (continued from the complete handle:
  // the stamps in the copied arrays refer to its counters)
  const size_t N = (size_t)std::max<int64_t>(h->N, 1), C = (size_t)std::max<int64_t>(h->C, 1);
  if (full->perm) {
    CUDA_TRY(om_malloc(h, &h->perm, 4 * N));
    CUDA_TRY(om_malloc(h, &h->inv_perm, 4 * N));
    CUDA_TRY(cudaMemcpyAsync(h->perm, full->perm, 4 * N, cudaMemcpyDeviceToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->inv_perm, full->inv_perm, 4 * N, cudaMemcpyDeviceToDevice,
                             h->stream));
  }
  CUDA_TRY(om_malloc(h, &h->cand, 4 * C));
  CUDA_TRY(om_malloc(h, &h->work, 4 * C));
  CUDA_TRY(om_malloc(h, &h->dirty, 4 * N));
  CUDA_TRY(om_malloc(h, &h->ds, sizeof(DevScalars)));
  CUDA_TRY(cudaMallocHost(&h->hs, sizeof(DevScalars)));
  CUDA_TRY(om_malloc(h, &h->partials, sizeof(double) * 8 * 2048));
  CUDA_TRY(cudaMemcpyAsync(h->ds, full->ds, sizeof(DevScalars), cudaMemcpyDeviceToDevice,
                           h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  CUDA_TRY(cudaMemsetAsync(&h->ds->sync_seq, 0, sizeof(unsigned long long), h->stream));
  CUDA_TRY(cudaMemsetAsync(&h->ds->sync_dead, 0, sizeof(int), h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  sh->mapped = true;
  sh->full = nullptr;
  for (auto& a : sh->arrays) a.source = nullptr;
  return OM_OK;
}

int om_shared_info(om_handle* h, int64_t* vertex_lo, int64_t* vertex_hi, int64_t* chunk_vertices,
                   int64_t* resident_bytes) {
  om_shared* sh = h ? shared_of(h) : nullptr;
  if (!sh) {
    om_set_error("not a shared handle");
    return OM_ERR_ARG;
  }
  if (vertex_lo) *vertex_lo = sh->vlo;
  if (vertex_hi) *vertex_hi = sh->vhi;
  if (chunk_vertices) *chunk_vertices = sh->nvc;
  if (resident_bytes) {
    size_t b = sh->ctrl.chunk;
    for (auto& a : sh->arrays) b += a.chunk;
    b += (size_t)h->C * 8 + (size_t)h->N * 4 + (h->perm ? (size_t)h->N * 8 : 0);
    *resident_bytes = (int64_t)b;
  }
  return OM_OK;
}

// The optimize() loop on a shared handle: every rank calls it with the same arguments.  The
// mesh must be Delaunay for its points (it is, right after om_shared_map of a handle that ran
// om_flip_until_delaunay, and after every om_shared_run).
int om_shared_run(om_handle* h, double tol, int64_t max_num_steps, int64_t* steps_done,
                  om_step_stats* last) {
  om_shared* sh = h ? shared_of(h) : nullptr;
  if (!sh || !sh->mapped) {
    om_set_error("om_shared_run: not a mapped shared handle");
    return OM_ERR_ARG;
  }
  if (max_num_steps < 1 || om_is_solve_method(h->method) || h->surf_kind != 0) {
    om_set_error("om_shared_run: fixed-point methods without a surface, max_num_steps >= 1");
    return OM_ERR_ARG;
  }
  cudaSetDevice(h->device);
  int which = -1;
  OM_TRY(graph_for(h, sh, h->x, h->xnew, &which));
  const int mode_exact = om_limiter_mode(h->limiter != 0, (long long)(h->limited_frac * 1.0e6), 1000000,
                                         om_lim_div());
  OM_LAUNCH(h, k_sh_init, 1, 1, h->ds, (long long)max_num_steps, tol * tol, mode_exact,
            (long long)h->N, h->limiter, 100, om_lim_div());
  CUDA_TRY(cudaGetLastError());
  double* A = h->x;
  double* B = h->xnew;
  CUDA_TRY(cudaGraphLaunch(sh->exec[which], h->stream));
  OM_TRY(om_fetch_scalars(h));
  h->launches += h->hs->pl_launches;
  const int64_t k = h->hs->k;
  h->x = (k & 1) ? B : A;
  h->xnew = (k & 1) ? A : B;
  h->run_flips = h->hs->total_flips;
  h->run_rounds = h->hs->total_rounds;
  h->run_limited = h->hs->total_limited;
  h->run_deferred = h->hs->total_deferred;
  if (h->hs->sync_dead) {
    om_set_error("a rank did not reach a meeting point within 20 s (shared loop abandoned)");
    return OM_ERR_CUDA;
  }
  OM_TRY(om_check_dev_err(h));
  om_step_stats st;
  memset(&st, 0, sizeof(st));
  om_step_stats_from_scalars(h, tol, &st);
  st.n_flips = h->hs->n_flips;
  st.n_flip_rounds = h->hs->n_rounds;
  st.flip_cap_hit = h->hs->cap_hit | (h->hs->not_delaunay ? 2 : 0);
  if (steps_done) *steps_done = k;
  if (last) *last = st;
  return OM_OK;
}

// builds the graphs of both buffer parities now (keeps the build out of a timed region)
int om_shared_prepare(om_handle* h) {
  om_shared* sh = h ? shared_of(h) : nullptr;
  if (!sh || !sh->mapped) {
    om_set_error("om_shared_prepare: not a mapped shared handle");
    return OM_ERR_ARG;
  }
  cudaSetDevice(h->device);
  int which = 0;
  OM_TRY(graph_for(h, sh, h->x, h->xnew, &which));
  OM_TRY(graph_for(h, sh, h->xnew, h->x, &which));
  return OM_OK;
}

// Times this rank's update kernel (the ring kernel with the fused check, lazy limiter) on its
// own vertex range with CUDA events, `reps` launches into the spare point buffer; the mesh is
// left as it was.  For the roofline line of a multi-GPU run (no collective, no meeting).
int om_shared_time_update(om_handle* h, int reps, double* ms_per_launch) {
  om_shared* sh = h ? shared_of(h) : nullptr;
  if (!sh || !sh->mapped || reps < 1 || !ms_per_launch) {
    om_set_error("om_shared_time_update: bad arguments");
    return OM_ERR_ARG;
  }
  cudaSetDevice(h->device);
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  OM_LAUNCH(h, k_sh_init, 1, 1, h->ds, 1ll, 0.0, 0, (long long)h->N, h->limiter, 100, om_lim_div());
  OM_TRY(om_pl_launch_update_part(h, h->x, h->xnew, 1));  // warm-up
  CUDA_TRY(cudaEventRecord(e0, h->stream));
  for (int i = 0; i < reps; i++) OM_TRY(om_pl_launch_update_part(h, h->x, h->xnew, 1));
  CUDA_TRY(cudaEventRecord(e1, h->stream));
  if (sh->vhi > sh->vlo)
    OM_LAUNCH(h, k_clear_flags, om_grid(sh->vhi - sh->vlo, 256), 256, h->vflags, sh->vlo, sh->vhi);
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms_per_launch = ms / reps;
  return OM_OK;
}

}  // extern "C"
