// The optimize() loop without the host in it.
//
// Replaces the body of optimesh.optimize(mesh, method, tol, max_num_steps)
// (/root/reference/README.md:131-132; loop semantics SURVEY.md A.5):
//     flip; repeat { update; flip } until final
// The reference order needs the geometry of the NEW points to find the edges to flip, i.e. a
// second pass over the whole mesh after every update.  Here the two passes are one: the update
// kernel of step k+1 evaluates the Delaunay indicator of every edge on the way (chain.cuh), on
// the points of step k and the topology of step k-1, so the loop runs as
//     flip_0;  repeat {  A: tentative x_{k+1} = U(x_k, M_{k-1})  +  flagged spokes of M_{k-1} at x_k
//                        B: flip pass from those flags              ->  M_k (Delaunay at x_k)
//                        C: vertices whose star changed: x_{k+1} = U(x_k, M_k), statistics  }
//     final flip pass (full check) for the last points.
// A vertex whose star did not change gets the same bits from A as from the reference order, so
// the trajectory is the reference's, step for step, flip for flip.
//
// Nothing in an iteration needs the host: list lengths, round stamps, the limiter variant of
// the next update, convergence and the flip-round loop all live on the device.  The whole loop
// is ONE CUDA graph: an outer WHILE node (two iterations per trip, the point buffers ping-pong)
// whose body holds an inner WHILE node per iteration for the flip rounds; the condition values
// are set by one-thread kernels (cudaGraphSetConditional).  The host launches the graph once
// and reads the scalars back once.
//
// OM_NO_GRAPH=1 (or om_set_timing: per-kernel events) runs the same kernels from the stream
// with one scalar readback per flip round and per step.
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

#include "common.cuh"

namespace {

__global__ void k_pl_init(DevScalars* ds, long long max_steps, double tol2, int mode_exact,
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
}

// ends an iteration: one more update is applied; decides what the host would decide
__global__ void k_pl_iter_end(DevScalars* ds, cudaGraphConditionalHandle handle, int use_handle) {
  if (!ds->halt) {
    ds->k++;
    ds->total_flips += ds->n_flips;
    ds->total_rounds += ds->n_rounds;
    ds->total_limited += (long long)ds->n_limited;
    ds->total_deferred += ds->n_deferred;
    ds->mode_exact = om_limiter_mode(ds->limiter_on != 0, (long long)ds->n_limited, ds->n_free,
                                     ds->lim_div_exact);
    double md;
    memcpy(&md, &ds->max_diff2_bits, 8);
    if (ds->err)
      ds->halt = 3;
    else if (md < ds->tol2 || ds->k >= ds->max_steps)
      ds->halt = 1;
  }
  if (use_handle) cudaGraphSetConditional(handle, ds->halt ? 0u : 1u);
}

// selects the limiter variant of the ring kernel for this iteration (two IF nodes)
__global__ void k_pl_mode(const DevScalars* ds, cudaGraphConditionalHandle lazy,
                          cudaGraphConditionalHandle exact) {
  const bool run = !(ds->halt & 1);
  cudaGraphSetConditional(lazy, run && ds->mode_exact != 1 ? 1u : 0u);
  cudaGraphSetConditional(exact, run && ds->mode_exact == 1 ? 1u : 0u);
}

struct PlGraph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  const double* buf_a = nullptr;  // the buffer that holds x_0 when the graph is launched
  int method = -1, limiter = -1, odt_bary = -1, use_rings = -1;
  double omega = 0.0;
};

struct PlCache {
  std::vector<PlGraph> graphs;
  cudaStream_t capture_stream = nullptr;
};

#define CU_TRY(expr)                                                                  \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      om_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__,      \
                   __LINE__, cudaGetErrorString(_e));                                 \
      return OM_ERR_CUDA;                                                             \
    }                                                                                 \
  } while (0)

// the kernels of one iteration after the ring kernel, up to the decision about the first flip
// round (a pass whose check flags nothing launches no flip kernel at all)
int enqueue_head_rest(om_handle* h, const double* xin, double* xout, unsigned long long inner,
                      int use_handle) {
  OM_TRY(om_pl_launch_update_part(h, xin, xout, 3));
  OM_TRY(om_pl_launch_flags_check(h, xin));
  OM_TRY(om_pl_launch_round_end(h, inner, use_handle));
  return OM_OK;
}

int enqueue_round(om_handle* h, const double* xin, unsigned long long inner, int use_handle) {
  OM_TRY(om_pl_launch_round(h, xin));
  OM_TRY(om_pl_launch_round_end(h, inner, use_handle));
  return OM_OK;
}

int enqueue_tail(om_handle* h, const double* xin, double* xout, unsigned long long outer,
                 int use_handle) {
  OM_TRY(om_pl_launch_tail(h, xin, xout));
  OM_LAUNCH(h, k_pl_iter_end, 1, 1, h->ds, (cudaGraphConditionalHandle)outer, use_handle);
  CUDA_TRY(cudaGetLastError());
  return OM_OK;
}

// Captures what `body` enqueues on the capture stream into `graph`, after the nodes `deps`;
// returns the last node of the captured chain.
template <typename F>
int capture_into(om_handle* h, cudaStream_t cs, cudaGraph_t graph, const cudaGraphNode_t* deps,
                 size_t ndeps, cudaGraphNode_t* last, F&& body) {
  CU_TRY(cudaStreamBeginCaptureToGraph(cs, graph, deps, nullptr, ndeps,
                                       cudaStreamCaptureModeRelaxed));
  cudaStream_t keep = h->stream;
  h->stream = cs;
  const int64_t launches = h->launches;
  int rc = body();
  h->stream = keep;
  h->launches = launches;  // captured, not launched
  cudaStreamCaptureStatus st;
  const cudaGraphNode_t* leaf = nullptr;
  size_t nleaf = 0;
  cudaError_t e = cudaStreamGetCaptureInfo_v2(cs, &st, nullptr, nullptr, &leaf, &nleaf);
  cudaGraphNode_t tail = (e == cudaSuccess && nleaf > 0) ? leaf[nleaf - 1] : nullptr;
  const bool single = nleaf == 1;
  cudaGraph_t out = nullptr;
  cudaError_t e2 = cudaStreamEndCapture(cs, &out);
  if (rc != OM_OK) return rc;
  CU_TRY(e);
  CU_TRY(e2);
  if (!single || !tail) {
    om_set_error("graph capture of the smoothing loop did not end in a single node");
    return OM_ERR_CUDA;
  }
  *last = tail;
  return OM_OK;
}

int add_while(cudaGraph_t parent, const cudaGraphNode_t* deps, size_t ndeps,
              cudaGraphConditionalHandle* handle, unsigned default_value, cudaGraphNode_t* node,
              cudaGraph_t* body) {
  CU_TRY(cudaGraphConditionalHandleCreate(handle, parent, default_value,
                                          cudaGraphCondAssignDefault));
  cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
  np.type = cudaGraphNodeTypeConditional;
  np.conditional.handle = *handle;
  np.conditional.type = cudaGraphCondTypeWhile;
  np.conditional.size = 1;
  CU_TRY(cudaGraphAddNode(node, parent, deps, ndeps, &np));
  *body = np.conditional.phGraph_out[0];
  return OM_OK;
}

// an IF node after `dep` whose body is what `body` enqueues
template <typename F>
int add_if(om_handle* h, cudaStream_t cs, cudaGraph_t parent, cudaGraphConditionalHandle handle,
           cudaGraphNode_t dep, cudaGraphNode_t* node, F&& body) {
  cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
  np.type = cudaGraphNodeTypeConditional;
  np.conditional.handle = handle;
  np.conditional.type = cudaGraphCondTypeIf;
  np.conditional.size = 1;
  CU_TRY(cudaGraphAddNode(node, parent, &dep, 1, &np));
  cudaGraphNode_t unused;
  return capture_into(h, cs, np.conditional.phGraph_out[0], nullptr, 0, &unused, body);
}

int build_graph(om_handle* h, PlCache* cache, PlGraph* g) {
  if (!cache->capture_stream)
    CU_TRY(cudaStreamCreateWithFlags(&cache->capture_stream, cudaStreamNonBlocking));
  cudaStream_t cs = cache->capture_stream;
  double* A = h->x;
  double* B = h->xnew;
  CU_TRY(cudaGraphCreate(&g->graph, 0));
  cudaGraphConditionalHandle outer;
  cudaGraphNode_t outer_node;
  cudaGraph_t body;
  OM_TRY(add_while(g->graph, nullptr, 0, &outer, 1u, &outer_node, &body));
  cudaGraphNode_t last = nullptr;
  for (int half = 0; half < 2; half++) {
    const double* xin = half == 0 ? A : B;
    double* xout = half == 0 ? B : A;
    // the handles have to exist before the kernels that set them are captured
    cudaGraphConditionalHandle inner, lazy, exact;
    CU_TRY(cudaGraphConditionalHandleCreate(&inner, body, 0u, cudaGraphCondAssignDefault));
    CU_TRY(cudaGraphConditionalHandleCreate(&lazy, body, 0u, cudaGraphCondAssignDefault));
    CU_TRY(cudaGraphConditionalHandleCreate(&exact, body, 0u, cudaGraphCondAssignDefault));
    OM_TRY(capture_into(h, cs, body, last ? &last : nullptr, last ? 1 : 0, &last, [&]() -> int {
      OM_TRY(om_pl_launch_update_part(h, xin, xout, 0));
      OM_LAUNCH(h, k_pl_mode, 1, 1, (const DevScalars*)h->ds, lazy, exact);
      CUDA_TRY(cudaGetLastError());
      return (int)OM_OK;
    }));
    // only the limiter variant the device selected runs (an empty 78k-block launch is 40 us)
    cudaGraphNode_t if_lazy, if_exact;
    OM_TRY(add_if(h, cs, body, lazy, last, &if_lazy,
                  [&] { return om_pl_launch_update_part(h, xin, xout, 1); }));
    OM_TRY(add_if(h, cs, body, exact, if_lazy, &if_exact,
                  [&] { return om_pl_launch_update_part(h, xin, xout, 2); }));
    last = if_exact;
    OM_TRY(capture_into(h, cs, body, &last, 1, &last, [&] {
      return enqueue_head_rest(h, xin, xout, (unsigned long long)inner, 1);
    }));
    cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
    np.type = cudaGraphNodeTypeConditional;
    np.conditional.handle = inner;
    np.conditional.type = cudaGraphCondTypeWhile;
    np.conditional.size = 1;
    cudaGraphNode_t inner_node;
    CU_TRY(cudaGraphAddNode(&inner_node, body, &last, 1, &np));
    cudaGraph_t rounds = np.conditional.phGraph_out[0];
    cudaGraphNode_t unused;
    OM_TRY(capture_into(h, cs, rounds, nullptr, 0, &unused, [&] {
      return enqueue_round(h, xin, (unsigned long long)inner, 1);
    }));
    OM_TRY(capture_into(h, cs, body, &inner_node, 1, &last, [&] {
      return enqueue_tail(h, xin, xout, (unsigned long long)outer, 1);
    }));
  }
  CU_TRY(cudaGraphInstantiate(&g->exec, g->graph, 0));
  g->buf_a = A;
  g->method = h->method;
  g->limiter = h->limiter;
  g->odt_bary = h->odt_bary;
  g->use_rings = h->use_rings ? 1 : 0;
  g->omega = h->omega;
  return OM_OK;
}

void free_graph(PlGraph& g) {
  if (g.exec) cudaGraphExecDestroy(g.exec);
  if (g.graph) cudaGraphDestroy(g.graph);
  g.exec = nullptr;
  g.graph = nullptr;
}

// the same loop driven from the stream: one scalar readback per flip round and per step.
// Timed (om_set_timing): CUDA events between the phases of every iteration.
int run_stream(om_handle* h, double* A, double* B, int mode) {
  const bool timed = h->timing;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  if (timed)
    for (auto& e : ev) CUDA_TRY(cudaEventCreate(&e));
  auto mark = [&](int i) {
    if (timed) cudaEventRecord(ev[i], h->stream);
  };
  int rc = OM_OK;
  for (int64_t it = 0; rc == OM_OK; it++) {
    const double* xin = (it & 1) ? B : A;
    double* xout = (it & 1) ? A : B;
    mark(0);
    if ((rc = om_pl_launch_update_part(h, xin, xout, 0)) != OM_OK) break;
    if (timed) cudaEventRecord(h->ev[0], h->stream);
    if ((rc = om_pl_launch_update_part(h, xin, xout, mode == 1 ? 2 : 1)) != OM_OK) break;
    if (timed) cudaEventRecord(h->ev[1], h->stream);
    mark(1);
    if ((rc = om_pl_launch_update_part(h, xin, xout, 3)) != OM_OK) break;
    mark(2);
    if ((rc = om_pl_launch_flags_check(h, xin)) != OM_OK) break;
    if ((rc = om_pl_launch_round_end(h, 0ull, 0)) != OM_OK) break;
    mark(3);
    while (rc == OM_OK) {
      if ((rc = om_fetch_scalars(h)) != OM_OK || !h->hs->pl_go) break;
      rc = enqueue_round(h, xin, 0ull, 0);
    }
    if (rc != OM_OK) break;
    mark(4);
    if ((rc = enqueue_tail(h, xin, xout, 0ull, 0)) != OM_OK) break;
    mark(5);
    if ((rc = om_fetch_scalars(h)) != OM_OK) break;
    if (timed) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]) == cudaSuccess) {
        h->t_step_ms += ms;
        h->n_step++;
      }
      for (int i = 0; i < 5; i++)
        if (cudaEventElapsedTime(&ms, ev[i], ev[i + 1]) == cudaSuccess) h->t_phase_ms[i] += ms;
      h->n_phase++;
    }
    if (h->hs->halt) break;
    mode = h->hs->mode_exact;
  }
  if (timed)
    for (auto& e : ev) cudaEventDestroy(e);
  return rc;
}

}  // namespace

void om_pl_destroy(om_handle* h) {
  PlCache* c = (PlCache*)h->pl;
  if (!c) return;
  for (auto& g : c->graphs) free_graph(g);
  if (c->capture_stream) cudaStreamDestroy(c->capture_stream);
  delete c;
  h->pl = nullptr;
}

namespace {
int get_graph(om_handle* h, PlGraph** out) {
  PlCache* cache = (PlCache*)h->pl;
  if (!cache) h->pl = cache = new PlCache();
  for (auto& c : cache->graphs)
    if (c.buf_a == h->x && c.method == h->method && c.limiter == h->limiter &&
        c.odt_bary == h->odt_bary && c.use_rings == (h->use_rings ? 1 : 0) &&
        c.omega == h->omega) {
      *out = &c;
      return OM_OK;
    }
  if (cache->graphs.size() >= 4) {
    for (auto& c : cache->graphs) free_graph(c);
    cache->graphs.clear();
  }
  cache->graphs.emplace_back();
  PlGraph* g = &cache->graphs.back();
  const int rc = build_graph(h, cache, g);
  if (rc != OM_OK) {
    free_graph(*g);
    cache->graphs.pop_back();
    return rc;
  }
  *out = g;
  return OM_OK;
}
}  // namespace

// builds (and caches) the graph om_run would launch now, without running it
int om_pl_prepare(om_handle* h) {
  static const bool no_graph = getenv("OM_NO_GRAPH") != nullptr;
  if (no_graph || h->N == 0 || h->C == 0) return OM_OK;
  PlGraph* g = nullptr;
  return get_graph(h, &g);
}

// The whole loop for the fixed-point methods on one GPU without a surface.
int om_run_pipelined(om_handle* h, double tol, int64_t max_num_steps, int64_t* steps_done,
                     om_step_stats* last) {
  om_step_stats st;
  memset(&st, 0, sizeof(st));
  int64_t nf = 0;
  int32_t nr = 0, cap = 0;
  if (!h->delaunay_clean) OM_TRY(om_flip_impl(h, 0.0, 100, &nf, &nr, &cap));
  h->delaunay_clean = false;  // the points are about to move
  const int mode_exact = om_limiter_mode(h->limiter != 0, (long long)(h->limited_frac * 1.0e6), 1000000,
                                         om_lim_div());
  OM_LAUNCH(h, k_pl_init, 1, 1, h->ds, (long long)max_num_steps, tol * tol, mode_exact,
            (long long)h->N, h->limiter, 100, om_lim_div());
  CUDA_TRY(cudaGetLastError());
  double* A = h->x;
  double* B = h->xnew;
  static const bool no_graph = getenv("OM_NO_GRAPH") != nullptr;
  if (no_graph || h->timing) {
    OM_TRY(run_stream(h, A, B, mode_exact));
  } else {
    PlGraph* g = nullptr;
    OM_TRY(get_graph(h, &g));
    CUDA_TRY(cudaGraphLaunch(g->exec, h->stream));
    OM_TRY(om_fetch_scalars(h));
  }
  h->launches += h->hs->pl_launches;
  const int64_t k = h->hs->k;
  h->run_flips = h->hs->total_flips;
  h->run_rounds = h->hs->total_rounds;
  h->run_limited = h->hs->total_limited;
  h->run_deferred = h->hs->total_deferred;
  // x_k is in A after an even number of updates
  h->x = (k & 1) ? B : A;
  h->xnew = (k & 1) ? A : B;
  h->nbr_valid = false;
  OM_TRY(om_check_dev_err(h));
  om_step_stats_from_scalars(h, tol, &st);
  const int32_t cap_before = h->hs->cap_hit | (h->hs->not_delaunay ? 2 : 0);
  // the flip pass of the last step: nothing has looked at the last points yet
  OM_TRY(om_flip_impl(h, 0.0, 100, &st.n_flips, &st.n_flip_rounds, &st.flip_cap_hit));
  st.flip_cap_hit |= cap_before;
  // totals: the flips of step j are found by iteration j+1, those of the last step just now
  h->run_flips += st.n_flips;
  h->run_rounds += st.n_flip_rounds;
  if (steps_done) *steps_done = k;
  if (last) *last = st;
  return OM_OK;
}
