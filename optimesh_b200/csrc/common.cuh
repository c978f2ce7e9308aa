// Shared declarations for the optimesh_b200 device library (sm_100a).
#pragma once
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/optimesh_b200.h"

#define OM_NONE_CELL 0x7fffffff
#define OM_RING_W 8

// ---------------------------------------------------------------- error plumbing
void om_set_error(const char* fmt, ...);

#define CUDA_TRY(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      om_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__,    \
                   __LINE__, cudaGetErrorString(_e));                               \
      return OM_ERR_CUDA;                                                           \
    }                                                                               \
  } while (0)

#define OM_TRY(expr)            \
  do {                          \
    int _r = (expr);            \
    if (_r != OM_OK) return _r; \
  } while (0)

// error bits raised by kernels
enum { OM_DEV_DEGENERATE = 1, OM_DEV_NONMANIFOLD = 2, OM_DEV_INDEX = 4, OM_DEV_WALK = 8 };

// a flagged edge found by the sharded Delaunay check: both half-edges and s
struct FlipRec {
  int he;
  int twin;
  double s;
};

// device-side scalars, mirrored into pinned host memory after each phase
struct DevScalars {
  unsigned long long max_diff2_bits;  // non-negative double compared as integer
  unsigned long long max_f_bits;      // surface: max |f|
  unsigned long long n_limited;
  int n_flagged;   // non-Delaunay half-edges found in the last check
  int n_flips;     // flips applied in the last round
  int n_cand;      // candidate list length
  int n_work;      // work list length
  int n_rec;       // flagged-edge records written by the sharded check
  int n_dirty;     // vertices whose ring row must be rebuilt after the flip pass
  int n_over;      // vertices without a ring row met by the step kernel
  int n_rounds;    // rounds of the current flip pass that flipped at least one edge
  int stale;       // sharded check met a coordinate this rank does not hold
  int abort;       // gathered round rejected: 1 = some rank stale, 2 = some slot overflowed
  int max_count;   // largest record count among the gathered slots
  int epoch;       // stamp of the current flip round (dedupe of the cell lists)
  int flips_prev;  // n_flips at the last round boundary
  int dirty_pass;  // stamp of the current flip pass (dedupe of the dirty-vertex list)
  int err;         // OM_DEV_* bits
  int not_delaunay;  // a pass ended with flagged edges but no mutual pair (exact ties)
  double dot[4];   // PCG dot products
  // ---- pipelined loop (loop.cu): everything the host would decide between two steps
  int halt;        // 0 running, 1 final step reached, 3 device error (odd: stopped);
                   // 4 flush: only the flip pass of the last points is still to run (shared.cu)
  int mode_exact;  // limiter variant of the next point update (om_limiter_mode)
  int pl_go;       // another flip round follows (set by k_pl_round_end)
  int cap_hit;     // a flip pass ran out of rounds with flagged edges left
  int max_rounds;  // rounds a flip pass may take
  long long k;     // point updates applied so far
  long long max_steps;
  double tol2;
  long long total_flips, total_rounds, total_limited;  // over the steps of this run
  long long n_free;  // vertices the limiter statistics refer to
  long long pl_launches;  // kernels run by the pipelined loop (they are launched by the graph)
  int limiter_on;
  int n_deferred;  // vertices the ring kernel left to k_post in the last update
  long long total_deferred;
  // ---- shared address space over several GPUs (shared.cu)
  unsigned long long sync_seq;      // meetings of the ranks so far
  unsigned long long g_sum[4], g_max[2];  // what the last meeting reduced over the ranks
  unsigned long long g_flips_prev, g_flips;  // flips of the pass over all ranks
  int sync_dead;                    // a meeting timed out
  int pl_round;                     // flip rounds executed in the current pass (loop.cu)
  int lim_div_exact;                // threshold of om_limiter_mode (set by k_pl_init)
};

// Limiter variant of the ring kernel, chosen from the share of vertices the PREVIOUS update
// limited.
//   0 lazy: a division-free bound proves "not limited"; the vertices that fail it (about 1.3 x
//     the limited ones) are handed to k_post, which evaluates them exactly from their ring rows;
//   1 exact everywhere (+45 % for every warp): pays above 1/8 limited, e.g. in the first steps
//     on a fresh mesh, where most vertices are limited.
// Either variant gives a vertex the same bits.  OM_LIM_EXACT_DIV overrides the 8 (measurements).
__host__ __device__ inline int om_limiter_mode(bool limiter_on, long long limited,
                                               long long n_free, int div_exact = 8) {
  return (limiter_on && (long long)div_exact * limited > n_free) ? 1 : 0;
}
inline int om_lim_div() {
  static const int de = getenv("OM_LIM_EXACT_DIV") ? atoi(getenv("OM_LIM_EXACT_DIV")) : 8;
  return de;
}

// After the check of a flip round: do its flips run?  `cand` = edges the check flagged,
// `flips` = flips of the pass so far (both over all GPUs when there are several).  A round
// that flagged edges but flipped none (exact ties of s around a cycle of cells) ends the pass
// with not_delaunay set; the round cap ends it with cap_hit when edges are still flagged.
__device__ __forceinline__ unsigned om_flip_decide(DevScalars* ds, unsigned long long cand,
                                                   unsigned long long flips) {
  const bool progress = flips > ds->g_flips_prev;
  if (progress) {
    ds->n_rounds++;
    ds->g_flips_prev = flips;
  }
  ds->g_flips = flips;
  unsigned go = 0u;
  if (ds->pl_round > 0 && !progress) {
    ds->not_delaunay = 1;
  } else if (cand > 0) {
    if (ds->n_rounds >= ds->max_rounds) {
      ds->cap_hit = 1;
    } else {
      go = 1u;
      ds->pl_round++;
    }
  }
  ds->pl_go = (int)go;
  return go;
}

struct om_handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int64_t N = 0, C = 0;
  int D = 2, PD = 2;
  int cells_itemsize = 8;
  // mesh state (internal numbering)
  double* x = nullptr;       // N*PD current points
  double* xnew = nullptr;    // N*PD next points (ping-pong)
  int4* cells = nullptr;     // C: vertex ids, .w = caller's cell row
  int4* adj = nullptr;       // C: twin half-edge (4*cell + slot) per local edge, -1 = boundary
  int* v2c = nullptr;        // N: one incident cell (OM_NONE_CELL for orphans)
  uint8_t* bflag = nullptr;  // N: 1 = pinned (boundary / ghost)
  int* ring = nullptr;       // N x OM_RING_W: one-ring vertex ids of free interior vertices
  int* ringc = nullptr;      // N x OM_RING_W: the cells between them (cell q = (v, n_q, n_q+1))
  int* dirty = nullptr;      // N: vertices touched by flips (ring rows to rebuild)
  int* dirty_epoch = nullptr;// N: dedupe stamps for `dirty`
  double* diff2 = nullptr;   // N: |diff|^2 of the last point update, sign bit = limited
  unsigned short* vflags = nullptr;  // N (+pad): per-vertex flag words between step kernels
  bool use_rings = true;
  int* perm = nullptr;       // internal -> caller vertex id (nullptr: identity)
  int* inv_perm = nullptr;   // caller -> internal
  // flip scratch
  int* cand = nullptr;       // C: cells with a flagged edge in the current round
  int* work = nullptr;       // C: cells to re-check in the next round
  int* work_epoch = nullptr; // C: dedupe stamps for the work list
  int* cand_epoch = nullptr; // C: dedupe stamps for the candidate list
  double* sarr = nullptr;    // 4C: s of flagged half-edges (+inf when not flagged)
  FlipRec* recs = nullptr;   // C: records of the sharded check (allocated on first use)
  int* flip_epoch = nullptr; // C
  int* reloc = nullptr;      // 4C
  int4* adj_tmp = nullptr;   // C
  // PCG scratch (allocated on first use)
  int* nbr_ptr = nullptr;    // N+1
  int* nbr_idx = nullptr;    // nnz
  float* nbr_w = nullptr;    // nnz edge multiplicities
  int64_t nnz = 0;
  double* pcg_buf = nullptr; // 4 vectors of N*PD
  double* target_buf = nullptr;  // N*PD: targets handed to the caller (om_targets_device)
  bool nbr_valid = false;
  // scalars
  DevScalars* ds = nullptr;  // device
  DevScalars* hs = nullptr;  // pinned host
  double* partials = nullptr;  // per-block partial sums (stats)
  // settings
  int method = OM_LLOYD;
  double omega = 1.0;
  int limiter = 1;
  int surf_kind = 0;
  double surf_tol = 1e-10;
  double surf_params[4] = {0, 0, 0, 1};
  int surf_max_sweeps = 100;
  double solver_rtol = 1e-13;
  int solver_max_iter = 100000;
  int64_t launches = 0;
  // owned vertex range (internal numbering) when the step is sharded across handles
  int64_t own_lo = 0, own_hi = -1;  // hi < 0: whole mesh
  // partitioned coordinates (dist.py): which foreign vertices hold current coordinates
  int* valid_epoch = nullptr;  // N: stamp of the last band exchange that refreshed the vertex
  int valid_stamp = 1;
  bool all_valid = true;       // right after a full all-gather (or single GPU)
  bool rings_partial = false;  // ring rows outside the owned range may be out of date
  bool defer_commit = false;   // om_update_points leaves the own range in xnew (om_commit_points)
  int* band = nullptr;         // N: own vertices other ranks may need (internal ids)
  uint8_t* band_mark = nullptr;  // N: hop distance to a foreign vertex (0: far)
  int64_t pass_work_bound = 0; // upper bound of the work list length in a round-wise pass
  int odt_bary = 1;            // ODT: barycenters for cells with a boundary edge
  // round-wise (partitioned) flip pass: only own vertices are enlisted for the ring rebuild and
  // only cells of the own range for the next check (every rank applies all flips, but each
  // examines and rebuilds only its part).  Full ranges outside such a pass.
  int flt_vlo = 0, flt_vhi = 0x7fffffff, flt_clo = 0, flt_chi = 0x7fffffff;
  int flip_spec = 0;           // flip rounds to enqueue before the first readback (flip.cu)
  double limited_frac = 1.0;  // share of vertices limited in the previous step
  // optional event timing (om_set_timing)
  void* stage = nullptr;  // pinned staging buffers of the host transfers (api.cu)
  bool timing = false;
  bool ev_pending = false;  // ev[0..1] recorded, elapsed time not read yet
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  void* pl = nullptr;  // cached graphs of the pipelined loop (loop.cu)
  void* sh = nullptr;  // shared address space over several GPUs (shared.cu); the mesh arrays
                       // of such a handle are not its own allocations
  // the last whole-mesh flip pass ended without flagged edges and no coordinate has changed
  // since: the flip pass that opens the loop has nothing to do
  bool delaunay_clean = false;
  int64_t run_flips = 0, run_rounds = 0, run_limited = 0, run_deferred = 0;  // last om_run
  double t_step_ms = 0.0, t_flip_ms = 0.0;
  // phases of the pipelined loop when it is timed (loop.cu, run_stream): reset+ring kernel,
  // k_post, flag check, flip rounds, ring rows + recomputation + statistics
  double t_phase_ms[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  int64_t n_phase = 0;
  int64_t n_step = 0, n_flip = 0;
};

// (A rank of a shared mesh may own no vertex at all -- a mesh of fewer than world x 2^21
// vertices: its range-sized grids are empty, but it must enqueue the same kernels as the others,
// the loop is a captured graph and the ranks meet in it.  Every kernel returns at once on an
// empty range.)
#define OM_LAUNCH(h, kernel, grid, block, ...)                         \
  do {                                                                 \
    int _g = (int)(grid);                                              \
    if (_g < 1 && (h)->sh) _g = 1;                                     \
    if (_g > 0) {                                                      \
      kernel<<<_g, (block), 0, (h)->stream>>>(__VA_ARGS__);            \
      (h)->launches++;                                                 \
    }                                                                  \
  } while (0)

// the same with dynamic shared memory
#define OM_LAUNCH_SMEM(h, kernel, grid, block, smem, ...)                  \
  do {                                                                    \
    int _g = (int)(grid);                                                 \
    if (_g < 1 && (h)->sh) _g = 1;                                        \
    if (_g > 0) {                                                         \
      kernel<<<_g, (block), (smem), (h)->stream>>>(__VA_ARGS__);          \
      (h)->launches++;                                                    \
    }                                                                     \
  } while (0)

static inline int om_grid(int64_t n, int block) { return (int)((n + block - 1) / block); }

// Device memory comes from the device's stream-ordered pool (kept across handles: the
// release threshold is raised at the first om_create): allocation and release are ordered on
// the handle's stream, cost microseconds once the pool is warm and never synchronise the
// device the way cudaMalloc/cudaFree do.
template <typename T>
static inline cudaError_t om_malloc(om_handle* h, T** p, size_t bytes) {
  return cudaMallocAsync((void**)p, bytes ? bytes : 1, h->stream);
}
static inline cudaError_t om_free(om_handle* h, void* p) {
  return p ? cudaFreeAsync(p, h->stream) : cudaSuccess;
}

// fetch DevScalars to host (synchronises the stream)
int om_fetch_scalars(om_handle* h);
int om_check_dev_err(om_handle* h);

// setup.cu
int om_setup_mesh(om_handle* h, const double* points_dev, const void* cells_dev, int flags);
// flip.cu
int om_flip_impl(om_handle* h, double tol, int max_rounds, int64_t* n_flips, int32_t* n_rounds,
                 int32_t* cap_hit, bool first_round_given = false);
int om_flip_check_range_impl(om_handle* h, double tol, int64_t clo, int64_t chi,
                             int64_t* n_records);
int om_flip_add_records_impl(om_handle* h, const void* recs, int64_t n);
int om_flip_pass_begin_impl(om_handle* h);
int om_flip_round_check_impl(om_handle* h, double tol, int first, int64_t clo, int64_t chi,
                             int64_t* n_records, int32_t* stale, bool fetch = true);
int om_flip_round_pack_impl(om_handle* h, int cap, void* slot_dev);
int om_flip_round_apply_gathered_impl(om_handle* h, const void* gathered, int P, int cap,
                                      int64_t* n_cand, int64_t* n_flips_total, int32_t* abort_bits,
                                      int64_t* own_records);
int om_flip_round_apply_impl(om_handle* h, int64_t total_records, int64_t* n_cand,
                             int64_t* n_flips_total);
int om_flip_pass_end_impl(om_handle* h, int64_t* n_flips, int32_t* n_rounds);
int om_band_build_impl(om_handle* h, int depth, int64_t* n);
int om_band_alloc(om_handle* h);
int om_band_pack_impl(om_handle* h, const int* idx_dev, int64_t n, double* buf_dev);
int om_band_unpack_impl(om_handle* h, const int* idx_dev, int64_t n, const double* buf_dev);
// step.cu
int om_update_points_impl(om_handle* h, double tol, om_step_stats* out, bool target_only,
                          double* target_out, bool defer_fetch = false);
void om_step_stats_from_scalars(om_handle* h, double tol, om_step_stats* out);
int om_project_impl(om_handle* h, int32_t* sweeps);
int om_random_move_impl(om_handle* h, uint64_t seed, int round, double amplitude);
int om_commit_points_impl(om_handle* h);
int om_rebuild_rings(om_handle* h, bool all, bool device = false);
int om_launch_point_update(om_handle* h, double* out, bool check);
int om_launch_reduce_stats(om_handle* h);
int om_update_from_targets_impl(om_handle* h, const double* targets_dev, double tol,
                                om_step_stats* out);
int om_launch_fixup(om_handle* h, double* out);
// pipelined loop (loop.cu): one iteration = update (with the fused Delaunay check) from xin
// into xout, flip pass on xin, recomputation of the vertices whose star changed
int om_pl_launch_update(om_handle* h, const double* xin, double* xout, bool timed);
int om_pl_launch_update_part(om_handle* h, const double* xin, double* xout, int what);
int om_pl_launch_tail(om_handle* h, const double* xin, double* xout);
int om_pl_launch_flags_check(om_handle* h, const double* xin);
int om_pl_launch_flips(om_handle* h);
int om_pl_launch_flips_part(om_handle* h, int which);  // 0 select, 1 flip, 2 twin patch
int om_pl_launch_round_check(om_handle* h, const double* xin);
int om_pl_launch_round(om_handle* h, const double* xin);
int om_pl_launch_round_end(om_handle* h, unsigned long long handle, int use_handle);
int om_run_pipelined(om_handle* h, double tol, int64_t max_num_steps, int64_t* steps_done,
                     om_step_stats* last);
void om_pl_destroy(om_handle* h);
// shared.cu: the vertex range this rank updates when the mesh lives in the shared address space
// of several GPUs (unchanged for an ordinary handle)
void om_shared_vertex_range(om_handle* h, int* vlo, int* vhi);
void om_shared_destroy(om_handle* h);
int om_pl_prepare(om_handle* h);
// pcg.cu
int om_pcg_impl(om_handle* h, double rtol, int max_iter, int32_t* iters, double* relres,
                double* out /* N*PD, may alias h->xnew */);
int om_quasi_newton_impl(om_handle* h, double rtol, int max_iter, int32_t* iters, double* relres,
                         double* out);
// methods whose target comes from a linear solve instead of the fused step kernel
inline bool om_is_solve_method(int m) {
  return m == OM_CPT_LINEAR_SOLVE || m == OM_CPT_QUASI_NEWTON;
}
// stats.cu
int om_stats_impl(om_handle* h, int64_t* angle_hist72, int64_t* q_hist40, double* summary8);
