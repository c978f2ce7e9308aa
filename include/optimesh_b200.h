/* optimesh_b200 -- C-ABI of the B200-native smoothing step.
 *
 * Drop-in boundary for ONE hot path of meshpro/optimesh: the per-step smoothing update
 * (relaxed Lloyd, CVT block-diagonal, CPT fixed-point / linear-solve, ODT fixed-point)
 * plus the flip-until-Delaunay pass that follows every step.  The reference is pure
 * Python (no FFI of its own); the entry points below are what a ctypes binding under
 * the reference's Python API would call.  Reference interface replaced by each entry
 * point is cited as /root/reference/README.md:line (the mounted tree holds only the
 * README; the arithmetic follows SURVEY.md Appendix A).
 *
 * Conventions: every function returns 0 on success, non-zero on error; the message is
 * available from om_last_error() (thread-local).  No C++ exception crosses the
 * boundary.  The caller owns all host buffers; the library owns all device memory
 * inside the handle.  One handle = one mesh on one GPU, one CUDA stream; a handle must
 * not be used from two threads at once.  There is no CPU fallback: without a CUDA
 * device om_create fails.
 */
#ifndef OPTIMESH_B200_H
#define OPTIMESH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct om_handle om_handle;

/* vertices of slack at the end of the device point array (see om_points_device) */
#define OM_POINT_PAD 1024

/* method ids (names: README.md:80, :90, :104) */
enum {
  OM_LLOYD = 0,              /* --method lloyd                 README.md:80  */
  OM_CVT_BLOCK_DIAGONAL = 1, /* --method cvt-block-diagonal    README.md:80  */
  OM_CPT_FIXED_POINT = 2,    /* --method cpt-fixed-point       README.md:90  */
  OM_ODT_FIXED_POINT = 3,    /* --method odt-fixed-point       README.md:104 */
  OM_CPT_LINEAR_SOLVE = 4,   /* --method cpt-linear-solve      README.md:90  */
  OM_ODT_DP_FP = 5,          /* --method odt-dp-fp             README.md:104 */
  OM_CPT_QUASI_NEWTON = 6    /* --method cpt-quasi-newton      README.md:90  */
};

/* error codes */
enum {
  OM_OK = 0,
  OM_ERR_CUDA = 1,
  OM_ERR_ARG = 2,
  OM_ERR_DEGENERATE = 3,   /* zero-area cell (upstream: "Degenerate cells.") */
  OM_ERR_NONMANIFOLD = 4,  /* an edge with more than two adjacent cells, or bad topology */
  OM_ERR_INDEX = 5,        /* cell refers to a vertex outside [0, N) */
  OM_ERR_NOT_CONVERGED = 6
};

/* creation flags */
enum {
  OM_RENUMBER = 1 /* spatially (Morton) renumber vertices and cells internally; the
                     numbering seen through the API is always the caller's */
};

typedef struct om_step_stats {
  double max_diff2;      /* max_i |omega (new_i - x_i)|^2, before the step limiter   */
  int64_t n_limited;     /* vertices whose step was shortened by the limiter          */
  int64_t n_flips;       /* edge flips in the flip-until-Delaunay pass after the step */
  int32_t n_flip_rounds; /* rounds that flipped at least one edge                     */
  int32_t flip_cap_hit;  /* 1 if max rounds were exhausted with violations left       */
  int32_t is_final;      /* max_diff2 < tol^2 (the caller adds "k >= max_num_steps")  */
  int32_t solver_iters;  /* PCG iterations (cpt-linear-solve), else 0                 */
  int32_t surface_sweeps;/* Newton sweeps of the implicit-surface projection          */
  int32_t reserved;
} om_step_stats;

const char* om_last_error(void);

/* number of visible CUDA devices */
int om_device_count(int* n);

/* Builds the device mesh from host arrays -- replaces meshplex.MeshTri(points, cells)
 * (README.md:131) under optimize_points_cells (README.md:124-126).
 *   points_host: N x dim float64, C-contiguous (dim 2 or 3)
 *   cells_host : C x 3 integers of cells_itemsize bytes (4 or 8)
 *   stream     : a cudaStream_t to run on, or NULL for a private (non-blocking) stream;
 *                pass cudaStreamLegacy ((void*)1) to share the legacy default stream
 * Does: upload, optional renumbering, half-edge twin table, boundary flags.      */
int om_create(om_handle** h, int device, void* stream, int64_t N, int dim, int64_t C,
              const double* points_host, const void* cells_host, int cells_itemsize,
              int flags);
/* Same, inputs already resident in device memory (same layouts). */
int om_create_device(om_handle** h, int device, void* stream, int64_t N, int dim, int64_t C,
                     const double* points_dev, const void* cells_dev, int cells_itemsize,
                     int flags);
int om_destroy(om_handle* h);

/* method + omega relaxation: optimize(..., method, omega=...) README.md:80, :124-126 */
int om_set_method(om_handle* h, int method, double omega);
/* step limiter of the driver loop on/off (default on) */
int om_set_limiter(om_handle* h, int on);
/* ODT methods (odt-fixed-point, odt-dp-fp; README.md:104-113): cells with a boundary edge
 * contribute their barycenter instead of their circumcenter, which can lie outside the
 * domain there (default on; 0 = circumcenters everywhere, SURVEY.md A.8 as written) */
int om_set_odt_boundary_barycenters(om_handle* h, int on);
/* implicit surface, README.md:146-176.  kind 0: none; kind 1: sphere
 * f(x) = R^2 - |x - c|^2, params = {cx, cy, cz, R} (the README's Sphere is {0,0,0,1}) */
int om_set_surface(om_handle* h, int kind, double tol, const double* params, int max_sweeps);
/* PCG controls for cpt-linear-solve */
int om_set_solver(om_handle* h, double rtol, int max_iter);

/* mesh.flip_until_delaunay() of the driver loop (meshplex) */
int om_flip_until_delaunay(om_handle* h, double tol, int max_rounds, int64_t* n_flips,
                           int32_t* n_rounds, int32_t* cap_hit);

/* One iteration of the optimize() loop (README.md:131-132): new points, pin boundary,
 * omega, limiter, surface projection, then flip-until-Delaunay. */
int om_step(om_handle* h, double tol, om_step_stats* out);
/* Only the point update (no surface projection, no flips): for callers that project
 * on the host (generic implicit_surface objects) and for single-step parity tests. */
int om_update_points(om_handle* h, double tol, om_step_stats* out);
/* Projection alone (built-in surfaces). */
int om_project(om_handle* h, int32_t* sweeps);
/* The whole loop of optimize(mesh, method, tol, max_num_steps) on the device. */
int om_run(om_handle* h, double tol, int64_t max_num_steps, int64_t* steps_done,
           om_step_stats* last);
/* For the fixed-point methods on one GPU without a surface om_run launches ONE CUDA graph
 * that holds the whole loop (update with the fused Delaunay check, flip rounds on a device
 * condition, recomputation of the vertices whose star changed); om_run_prepare builds and
 * caches that graph for the current method / omega / limiter settings without running it, so
 * that a caller can keep the build out of a timed region.  om_get_run_totals: flips, flip
 * rounds, limited vertices and vertices the ring kernel left to its list-driven companion
 * (no ring row, or the lazy limiter bound failed), summed over the steps of the last om_run. */
int om_run_prepare(om_handle* h);
int om_get_run_totals(om_handle* h, int64_t* n_flips, int64_t* n_flip_rounds,
                      int64_t* n_limited, int64_t* n_deferred);

/* Synthetic workloads ("a randomly generated disk mesh", README.md:70-74): `rounds` times,
 * every free vertex moves by a random vector of length <= amplitude/2 x its smallest
 * incident inradius (no cell can invert), followed by flip-until-Delaunay.  Turns any valid
 * flat mesh into a random Delaunay mesh of the same domain without leaving the device; the
 * random numbers depend on (seed, round, caller vertex id) only.  2D meshes. */
int om_random_walk(om_handle* h, int rounds, uint64_t seed, double amplitude, int64_t* n_flips);

/* optimesh.get_new_points(mesh, method) (README.md:141): un-relaxed, un-limited target
 * positions, N x dim, caller numbering. */
int om_new_points(om_handle* h, double* out_host);

/* Device side of the reference's per-step hooks (README.md:146-149 boundary_step, :157-162
 * generic implicit surfaces): the caller's code works on device memory between the phases of a
 * step, nothing goes through the host.
 *   om_targets_device      un-relaxed targets of the current method (what om_new_points
 *                          returns) in a device buffer owned by the handle: internal numbering
 *                          and layout, like om_device_ptrs' points (N x point_stride doubles);
 *                          the caller may overwrite entries (boundary_step moves the targets
 *                          of the boundary vertices);
 *   om_update_from_targets x <- x + limiter(omega (target - x)) for EVERY vertex, boundary
 *                          vertices included (the loop's semantics when boundary_step is
 *                          given), statistics like om_update_points.  The flip pass and the
 *                          projection are the caller's next calls. */
int om_targets_device(om_handle* h, double** targets_dev);
int om_update_from_targets(om_handle* h, const double* targets_dev, double tol,
                           om_step_stats* out);

/* cpt-linear-solve alone: solve the Dirichlet graph Laplacian for all coordinates and
 * overwrite the interior points with the solution. */
int om_solve_graph_laplacian(om_handle* h, double rtol, int max_iter, int32_t* iters,
                             double* rel_residual);

/* print_stats (README.md:55-60): 72 angle bins of 2.5 deg, 40 quality bins of 0.025,
 * summary8 = {angle min, max, avg, std, q min, avg, max, std}. */
int om_stats(om_handle* h, int64_t* angle_hist72, int64_t* q_hist40, double* summary8);

/* mesh.points / mesh.cells (README.md:133), caller numbering and cell row order */
int om_get_points(om_handle* h, double* out_host);
int om_set_points(om_handle* h, const double* in_host);
int om_get_cells(om_handle* h, void* out_host, int itemsize);
int om_get_boundary_flags(om_handle* h, uint8_t* out_host);

/* Raw device pointers (internal numbering/layout), for zero-copy interop:
 *   points: N x pd float64 (pd = 2 for dim 2, 4 for dim 3), cells: C x int4 (.w = caller
 *   cell row), perm: internal -> caller vertex id (NULL if identity). */
int om_device_ptrs(om_handle* h, double** points, int32_t** cells4, int32_t** perm,
                   int32_t* point_stride);

/* Halo support for meshes partitioned across GPUs (one process per GPU; the exchange
 * itself is done by the caller with NCCL):
 *   om_pack_points  : buf[i] = x[idx[i]] for i < n   (device buffers, caller ids)
 *   om_unpack_points: x[idx[i]] = buf[i]
 *   idx_dev are caller vertex ids resident on the device. */
int om_pack_points(om_handle* h, const int32_t* idx_dev, int64_t n, double* buf_dev);
int om_unpack_points(om_handle* h, const int32_t* idx_dev, int64_t n, const double* buf_dev);
/* Vertices listed here are treated as pinned (ghost vertices of a partition). */
int om_pin_vertices(om_handle* h, const int32_t* idx_host, int64_t n);

/* Sharding the point update across several handles that hold the SAME mesh (one process
 * per GPU; identical inputs give identical internal numbering): the handle only updates
 * the internal vertex range [lo, hi) in om_update_points / om_step; the caller then makes
 * every range visible everywhere (NCCL all-gather straight on the point array:
 * om_points_device gives its address, layout N x stride doubles + OM_POINT_PAD vertices of
 * slack so equal-sized chunks may overrun N).  hi < 0 restores the whole mesh. */
int om_set_owned_range(om_handle* h, int64_t lo, int64_t hi);
/* Sharded first round of flip-until-Delaunay (the only round that scans every cell):
 *   om_flip_check_range  examines the internal cell range [cell_lo, cell_hi) and leaves the
 *                        flagged edges as 16-byte records {int32 half_edge, int32 twin,
 *                        double s} in a device buffer owned by the handle;
 *   om_flip_add_records  applies records (this rank's or gathered from other ranks; device
 *                        memory) -- call once per segment;
 *   om_flip_finish       runs the remaining rounds (work lists only) like
 *                        om_flip_until_delaunay.
 * Every rank must add the records of ALL ranks, so that topology stays identical. */
int om_flip_check_range(om_handle* h, double tol, int64_t cell_lo, int64_t cell_hi,
                        int64_t* n_records, void** records_dev);
int om_flip_add_records(om_handle* h, const void* records_dev, int64_t n);
int om_flip_finish(om_handle* h, double tol, int max_rounds, int64_t* n_flips,
                   int32_t* n_rounds, int32_t* cap_hit);
int om_points_device(om_handle* h, double** points, int64_t* n_alloc, int32_t* stride);

/* Partitioned coordinates (one process per GPU, identical topology on every rank, each rank
 * only keeps CURRENT coordinates for its own vertex range plus a band around it):
 *   om_coords_invalidate  foreign coordinates are stale from now on (call after the update)
 *   om_band_build         internal ids of the own vertices within `depth` edges of a foreign
 *                         vertex (pinned vertices excluded: they never move); device list
 *   om_band_pack/unpack   gather own band coordinates into a buffer (stride = point stride) /
 *                         scatter a received band (internal ids) and mark it current
 *   om_coords_all_valid   every coordinate is current again (after a full all-gather)
 * and the flip pass round by round, so that flagged-edge records can be exchanged in every
 * round (only the check reads coordinates; select/flip/patch are integer work every rank
 * repeats on the identical topology):
 *   om_flip_pass_begin
 *   om_flip_round_check   first != 0: cells [cell_lo, cell_hi); else the work-list cells in that
 *                         range.  *stale = 1: a coordinate outside own range + band would have
 *                         been read -- refresh (full all-gather) and call again
 *   om_flip_add_records   the records of ALL ranks (declared above)
 *   om_flip_round_apply   select + flip + patch; *n_candidates is the same on every rank,
 *                         0 ends the pass
 *   om_flip_pass_end      ring rows of touched vertices; totals of the pass              */
/* Cells whose smallest vertex lies in the internal vertex range [vertex_lo, vertex_hi): a
 * contiguous cell range right after om_create (cells are sorted by their smallest vertex);
 * call it before any flip.  Gives each rank the cells that sit on its own vertices. */
int om_cell_range_of_vertices(om_handle* h, int64_t vertex_lo, int64_t vertex_hi,
                              int64_t* cell_lo, int64_t* cell_hi);
/* With deferred commit om_update_points leaves the updated own range aside and reports in
 * om_step_stats.reserved whether it met a stale coordinate; the caller either commits (every
 * rank clean) or refreshes the coordinates and repeats the update. */
int om_set_deferred_commit(om_handle* h, int on);
int om_commit_points(om_handle* h);
int om_coords_invalidate(om_handle* h);
int om_coords_all_valid(om_handle* h);
int om_band_build(om_handle* h, int depth, int64_t* n, int32_t** idx_dev);
int om_band_pack(om_handle* h, const int32_t* idx_dev, int64_t n, double* buf_dev);
int om_band_unpack(om_handle* h, const int32_t* idx_dev, int64_t n, const double* buf_dev);
int om_flip_pass_begin(om_handle* h);
int om_flip_round_check(om_handle* h, double tol, int first, int64_t cell_lo, int64_t cell_hi,
                        int64_t* n_records, void** records_dev, int32_t* stale);
int om_flip_round_apply(om_handle* h, int64_t total_records, int64_t* n_candidates,
                        int64_t* n_flips_total);
/* Same round without a host readback between check and flips: call om_flip_round_check with
 * n_records == NULL, then om_flip_round_pack writes this rank's slot of (1 + capacity) 16-byte
 * records (slot[0] = {count, stale}), the caller all-gathers the slots, and
 * om_flip_round_apply_gathered applies all of them and flips.  *abort_bits != 0 (1: a rank was
 * stale, 2: a slot overflowed) means nothing was applied: repeat the round with
 * om_flip_round_check(n_records != NULL) / om_flip_add_records / om_flip_round_apply. */
int om_flip_round_pack(om_handle* h, int32_t capacity, void* slot_dev);
int om_flip_round_apply_gathered(om_handle* h, const void* gathered_dev, int32_t n_ranks,
                                 int32_t capacity, int64_t* n_candidates, int64_t* n_flips_total,
                                 int32_t* abort_bits, int64_t* max_records_of_any_rank);
int om_flip_pass_end(om_handle* h, int64_t* n_flips, int32_t* n_rounds);

/* ONE mesh in ONE address space over the GPUs of a box (one process per GPU; NVLink / NVSwitch
 * peer memory through CUDA virtual memory management): every mesh array is cut into
 * `world` chunks by vertex / cell id, chunk r is physical memory of rank r's GPU, and all
 * chunks are mapped back to back into one virtual range on every rank -- memory per rank
 * scales with 1/world, the kernels use the same global ids everywhere, what they touch across
 * a chunk boundary is an ordinary access to peer memory.
 *   om_shared_begin  from a complete handle (om_create + om_flip_until_delaunay, identical on
 *                    every rank): allocates this rank's chunks, returns their POSIX file
 *                    descriptors (at most 32) for the caller to pass to the other ranks
 *                    (SCM_RIGHTS over a Unix socket);
 *   om_shared_map    fds_all[r * n_fds + i] = descriptor i of rank r as received here: maps
 *                    everything, copies this rank's share of the mesh.  Synchronise the ranks
 *                    afterwards; the complete handle may then be destroyed;
 *   om_shared_run    the optimize() loop, called by every rank with the same arguments: one
 *                    CUDA graph per rank, the ranks meet on the device (barrier + all-reduce
 *                    through peer memory, ~9 us); results through om_get_points / om_get_cells
 *                    on any rank (after a host barrier);
 *   om_shared_info   this rank's vertex range, the chunk size and its resident bytes.
 * om_destroy of a shared handle: only after every rank has stopped using the mesh. */
int om_shared_begin(om_handle* complete, int rank, int world, om_handle** out, int32_t* fds,
                    int32_t* n_fds);
int om_shared_map(om_handle* h, const int32_t* fds_all, int32_t n_fds);
int om_shared_run(om_handle* h, double tol, int64_t max_num_steps, int64_t* steps_done,
                  om_step_stats* last);
int om_shared_info(om_handle* h, int64_t* vertex_lo, int64_t* vertex_hi, int64_t* chunk_vertices,
                   int64_t* resident_bytes);
/* om_shared_prepare builds the loop's CUDA graphs now (keeps the build out of a timed region);
 * om_shared_time_update times this rank's update kernel on its own vertex range with CUDA
 * events (`reps` launches into the spare buffer; the mesh is left as it was). */
int om_shared_prepare(om_handle* h);
int om_shared_time_update(om_handle* h, int reps, double* ms_per_launch);

/* Per-kernel timing with CUDA events on the handle's stream (bench.py's roofline line):
 * when on, the fused step kernel (K1) and each flip-until-Delaunay pass are bracketed by
 * events; om_get_timing returns the accumulated device times and counts. */
int om_set_timing(om_handle* h, int on);
int om_get_timing(om_handle* h, double* step_kernel_ms, int64_t* step_kernel_launches,
                  double* flip_pass_ms, int64_t* flip_passes);
/* While timing is on, om_run's loop (driven from the stream) also times its phases with CUDA
 * events: phase_ms5 = {reset + ring kernel, vertices it left (k_post), flag-driven Delaunay
 * check, flip rounds, ring rows + recomputation of touched vertices + statistics}, summed over
 * `iterations` loop iterations. */
int om_get_phase_timing(om_handle* h, double* phase_ms5, int64_t* iterations);

/* Device memory of destroyed handles stays in the device's stream-ordered pool so that the
 * next om_create is cheap; this returns it to the driver. */
int om_release_cached_memory(int device);

/* Result buffers in pinned host memory, cached for the life of the process: om_get_points /
 * om_get_cells into such a block are one DMA at PCIe speed (cells are widened on the device)
 * instead of a staged copy into fresh pageable pages.  *out == NULL: the cache is full (4 GB)
 * or pinning failed -- use ordinary memory.  om_result_free hands a block back to the cache. */
int om_result_alloc(int64_t bytes, void** out);
int om_result_free(void* p, int64_t bytes);

/* Touches every page of a host buffer the caller is about to receive results in (huge pages
 * requested), from the library's host threads; meant to run while the device is busy. */
int om_prefault_host(void* p, int64_t bytes);

/* kernels launched by this handle so far (bench.py's gpu_launches) */
int om_launch_count(om_handle* h, int64_t* n);
int om_synchronize(om_handle* h);
/* stream the handle launches on (cudaStream_t) */
int om_stream(om_handle* h, void** stream);

#ifdef __cplusplus
}
#endif
#endif
