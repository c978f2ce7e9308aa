// K5: Jacobi-preconditioned CG for the Dirichlet graph Laplacian of cpt-linear-solve.
//
// Replaces optimesh.cpt linear_solve's COO->CSR assembly + scipy.sparse.linalg.spsolve
// (/root/reference/README.md:90, :92-95).  The matrix is never formed:
// for an interior vertex i the row is  2 (deg_i x_i - sum_{j ~ i} x_j) = 0  (every edge at
// an interior vertex is interior, hence counted by two cells); boundary rows are the
// identity.  Eliminating the boundary values gives an SPD system on the interior unknowns
// which is solved for all coordinates at once (one alpha/beta per coordinate).
// Dot products are reduced in a fixed order (per-block partials, then the last block sums
// them by index) so the iteration is bitwise reproducible.
//
// The same iteration, with edge weights, solves the approximate-Hessian system of
// cpt-quasi-newton (README.md:90, :97-98; Chen-Holst): with d = 2 and |w_i| the area of the
// star of i,
//   H_ii = 2/(d+1) |w_i|,   H_ij = -2/(d+1)^2 (|t_a| + |t_b|)   (t_a, t_b: the cells on edge ij),
//   rhs_i = -dE_i = -2/(d+1) sum_{t in star(i)} |t| (x_i - b_t),   boundary rows: identity, 0.
// Every cell of a star meets two edges at i, so the off-diagonal row sum is 2/3 of the
// diagonal: H is strictly diagonally dominant, SPD, with a condition number <= 5 whatever the
// mesh size -- a few dozen iterations.
#include <cub/cub.cuh>
#include <vector>

#include "common.cuh"
#include "geom.cuh"

namespace {

constexpr int PB = 256;
constexpr int MAXG = 1184;  // 148 SMs x 8 resident blocks

struct PcgScal {
  double rz[3];   // r.z of the current iterate, per coordinate
  double pq[3];   // p.Ap
  double upd[6];  // after k_pcg_update: new r.z in [0,D), r.r in [D,2D)
  double ini[6];  // after k_pcg_init: r.z in [0,D), b.b in [D,2D)
  unsigned int ticket;
};

// ---- neighbour lists by walking vertex stars (interior vertices only)
template <bool FILL>
__global__ void __launch_bounds__(256)
    k_ring(const int4* __restrict__ cells, const int* __restrict__ adj, const int* __restrict__ v2c,
           const uint8_t* __restrict__ bflag, int N, int* __restrict__ cnt_or_ptr,
           int* __restrict__ idx, DevScalars* ds) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= N) return;
  int c0 = v2c[v];
  if (c0 == OM_NONE_CELL || bflag[v]) {
    if (!FILL) cnt_or_ptr[v] = 0;
    return;
  }
  int n = 0;
  int base = FILL ? cnt_or_ptr[v] : 0;
  int cur = c0;
  int4 cl = __ldg(cells + cur);
  int j = slot_of(cl, v);
  if (j < 0) {
    atomicOr(&ds->err, OM_DEV_WALK);
    if (!FILL) cnt_or_ptr[v] = 0;
    return;
  }
  int kexit = (j + 1) % 3;
  while (true) {
    // leaving through the edge opposite slot kexit: the ring vertex on that edge is the one
    // that is neither v nor the vertex at slot kexit
    int jn = slot_of(cl, v);
    int other = cell_get(cl, 3 - jn - kexit);
    if (FILL) idx[base + n] = other;
    n++;
    int t = __ldg(adj + 4 * (size_t)cur + kexit);
    if (t < 0 || n > 4096) {
      atomicOr(&ds->err, OM_DEV_WALK);
      break;
    }
    int cn = t >> 2, kn = t & 3;
    if (cn == c0) break;
    cl = __ldg(cells + cn);
    int j2 = slot_of(cl, v);
    if (j2 < 0 || j2 == kn) {
      atomicOr(&ds->err, OM_DEV_WALK);
      break;
    }
    cur = cn;
    kexit = 3 - j2 - kn;
  }
  if (!FILL) cnt_or_ptr[v] = n;
}

__global__ void k_set_last(int* ptr, const int* cnt, int N) {
  // ptr holds the exclusive scan of cnt; close the CSR
  ptr[N] = ptr[N - 1] + cnt[N - 1];
}

// block reduction of ND values per thread into partials[blk*ND + i]; the last block to
// arrive sums the partials in index order into out[0..ND)
template <int ND>
__device__ __forceinline__ void block_reduce_store(double (&vals)[ND], double* partials,
                                                   double* out, unsigned int* ticket) {
  __shared__ double red[ND][PB / 32];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < ND; i++) {
    double v = vals[i];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[i][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < ND) {
    double v = 0.0;
    for (int w = 0; w < PB / 32; w++) v += red[threadIdx.x][w];
    partials[(size_t)blockIdx.x * ND + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    if (threadIdx.x < ND) {
      double v = 0.0;
      for (unsigned int b = 0; b < gridDim.x; b++)
        v += ((volatile double*)partials)[(size_t)b * ND + threadIdx.x];
      out[threadIdx.x] = v;
    }
    if (threadIdx.x == 0) *ticket = 0u;
  }
}

// (A y)_i = deg_i y_i - sum_j y_j for interior i; y is zero on fixed vertices
template <int D>
__device__ __forceinline__ Vec<D> apply_row(const double* __restrict__ y, const int* __restrict__ ptr,
                                            const int* __restrict__ idx, int v, const Vec<D>& yi) {
  const int b = ptr[v], e = ptr[v + 1];
  Vec<D> s;
#pragma unroll
  for (int k = 0; k < D; k++) s.v[k] = (double)(e - b) * yi.v[k];
  for (int q = b; q < e; q++) {
    Vec<D> yj = ld_point<D>(y, __ldg(idx + q));
#pragma unroll
    for (int k = 0; k < D; k++) s.v[k] -= yj.v[k];
  }
  return s;
}

// weighted row: (A y)_i = diag_i y_i - sum_j w_ij y_j
template <int D>
__device__ __forceinline__ Vec<D> apply_row_w(const double* __restrict__ y,
                                              const int* __restrict__ ptr,
                                              const int* __restrict__ idx,
                                              const double* __restrict__ w, double diag, int v,
                                              const Vec<D>& yi) {
  const int b = ptr[v], e = ptr[v + 1];
  Vec<D> s;
#pragma unroll
  for (int k = 0; k < D; k++) s.v[k] = diag * yi.v[k];
  for (int q = b; q < e; q++) {
    const Vec<D> yj = ld_point<D>(y, __ldg(idx + q));
    const double wq = __ldg(w + q);
#pragma unroll
    for (int k = 0; k < D; k++) s.v[k] -= wq * yj.v[k];
  }
  return s;
}

// area of the triangle (P0, P1, P2); 0 and an error flag if degenerate
template <int D>
__device__ __forceinline__ double tri_area(const Vec<D>& P0, const Vec<D>& P1, const Vec<D>& P2,
                                           int& err) {
  const Vec<D> a = vsub<D>(P1, P0), b = vsub<D>(P2, P0);
  const double aa = vdot<D>(a, a), bb = vdot<D>(b, b), ab = vdot<D>(a, b);
  const double v4 = fma(aa, bb, -ab * ab);  // 4 A^2
  if (!(v4 > 0.0)) {
    err |= OM_DEV_DEGENERATE;
    return 0.0;
  }
  return 0.5 * sqrt(v4);
}

// cpt-quasi-newton: weights, diagonal and right-hand side from the neighbour rows.  The row
// of an interior vertex lists its ring in walk order (a closed cycle), so cell p of the star
// is (v, ring[p], ring[p+1]) and the edge to ring[p] is shared by cells p-1 and p.  Every
// quantity is written as a sum over that cycle: no dependence on where the row starts.
// Also starts the iteration from delta = 0: r = rhs, z = r / diag, p = z, rz, b.b.
template <int D>
__global__ void __launch_bounds__(PB)
    k_qn_setup(const double* __restrict__ x, const int* __restrict__ ptr,
               const int* __restrict__ idx, int N, double* __restrict__ w,
               double* __restrict__ diag, double* __restrict__ r, double* __restrict__ p,
               double* partials, PcgScal* sc, DevScalars* ds) {
  constexpr double C_DIAG = 2.0 / 3.0, C_OFF = 2.0 / 9.0;  // 2/(d+1), 2/(d+1)^2 with d = 2
  double vals[2 * D];
#pragma unroll
  for (int k = 0; k < 2 * D; k++) vals[k] = 0.0;
  int err = 0;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < N; v += gridDim.x * blockDim.x) {
    Vec<D> rv, pv;
#pragma unroll
    for (int k = 0; k < D; k++) rv.v[k] = pv.v[k] = 0.0;
    const int b = ptr[v], e = ptr[v + 1];
    double dg = 0.0;
    if (e > b) {
      const Vec<D> xv = ld_point<D>(x, v);
      // cell p = (v, ring[p], ring[p+1]); A_prev is the area of cell p-1
      Vec<D> Pp = ld_point<D>(x, __ldg(idx + e - 1));  // ring[-1]
      Vec<D> Pc = ld_point<D>(x, __ldg(idx + b));      // ring[0]
      double A_prev = tri_area<D>(xv, Pp, Pc, err);
      double area_sum = 0.0;
      Vec<D> g;  // sum_p A_p (2 x_v - ring[p] - ring[p+1]) = 3 sum_p A_p (x_v - b_p)
#pragma unroll
      for (int k = 0; k < D; k++) g.v[k] = 0.0;
      for (int q = b; q < e; q++) {
        const Vec<D> Pn = ld_point<D>(x, __ldg(idx + (q + 1 < e ? q + 1 : b)));  // ring[p+1]
        const double A = tri_area<D>(xv, Pc, Pn, err);                            // cell p
        w[q] = C_OFF * (A_prev + A);  // edge (v, ring[p]): cells p-1 and p
        area_sum += A;
#pragma unroll
        for (int k = 0; k < D; k++) g.v[k] += A * (2.0 * xv.v[k] - Pc.v[k] - Pn.v[k]);
        A_prev = A;
        Pc = Pn;
      }
      dg = C_DIAG * area_sum;
      const double invd = dg > 0.0 ? 1.0 / dg : 0.0;
#pragma unroll
      for (int k = 0; k < D; k++) {
        rv.v[k] = -C_DIAG * g.v[k] * (1.0 / 3.0);
        pv.v[k] = rv.v[k] * invd;
        vals[k] += rv.v[k] * pv.v[k];
        vals[D + k] += rv.v[k] * rv.v[k];
      }
    }
    diag[v] = dg;
    st_point<D>(r, v, rv);
    st_point<D>(p, v, pv);
  }
  if (err) atomicOr(&ds->err, err);
  block_reduce_store<2 * D>(vals, partials, sc->ini, &sc->ticket);
}

template <int D>
__global__ void __launch_bounds__(PB)
    k_qn_spmv(const double* __restrict__ p, const int* __restrict__ ptr,
              const int* __restrict__ idx, const double* __restrict__ w,
              const double* __restrict__ diag, int N, double* __restrict__ q, double* partials,
              PcgScal* sc) {
  double vals[D];
#pragma unroll
  for (int k = 0; k < D; k++) vals[k] = 0.0;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < N; v += gridDim.x * blockDim.x) {
    if (ptr[v + 1] > ptr[v]) {
      const Vec<D> pi = ld_point<D>(p, v);
      const Vec<D> qi = apply_row_w<D>(p, ptr, idx, w, diag[v], v, pi);
      st_point<D>(q, v, qi);
#pragma unroll
      for (int k = 0; k < D; k++) vals[k] += pi.v[k] * qi.v[k];
    }
  }
  block_reduce_store<D>(vals, partials, sc->pq, &sc->ticket);
}

// r = b - A x_I with b_i = sum_{j fixed} x_j, i.e. r_i = sum_all_j x_j - deg_i x_i;
// z = r / deg; p = z; rz = r.z; bb = b.b
template <int D>
__global__ void __launch_bounds__(PB)
    k_pcg_init(const double* __restrict__ x, const int* __restrict__ ptr, const int* __restrict__ idx,
               const uint8_t* __restrict__ bflag, int N, double* __restrict__ r,
               double* __restrict__ p, double* partials, PcgScal* sc) {
  double vals[2 * D];
#pragma unroll
  for (int k = 0; k < 2 * D; k++) vals[k] = 0.0;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < N; v += gridDim.x * blockDim.x) {
    Vec<D> rv, pv;
#pragma unroll
    for (int k = 0; k < D; k++) rv.v[k] = pv.v[k] = 0.0;
    const int b = ptr[v], e = ptr[v + 1];
    if (e > b) {
      Vec<D> xi = ld_point<D>(x, v);
      Vec<D> bi;
#pragma unroll
      for (int k = 0; k < D; k++) {
        rv.v[k] = -(double)(e - b) * xi.v[k];
        bi.v[k] = 0.0;
      }
      for (int q = b; q < e; q++) {
        const int j = __ldg(idx + q);
        Vec<D> xj = ld_point<D>(x, j);
        const bool fixed = bflag[j] != 0;
#pragma unroll
        for (int k = 0; k < D; k++) {
          rv.v[k] += xj.v[k];
          if (fixed) bi.v[k] += xj.v[k];
        }
      }
      const double invd = 1.0 / (double)(e - b);
#pragma unroll
      for (int k = 0; k < D; k++) {
        pv.v[k] = rv.v[k] * invd;
        vals[k] += rv.v[k] * pv.v[k];
        vals[D + k] += bi.v[k] * bi.v[k];
      }
    }
    st_point<D>(r, v, rv);
    st_point<D>(p, v, pv);
  }
  block_reduce_store<2 * D>(vals, partials, sc->ini, &sc->ticket);
}

template <int D>
__global__ void __launch_bounds__(PB)
    k_pcg_spmv(const double* __restrict__ p, const int* __restrict__ ptr, const int* __restrict__ idx,
               int N, double* __restrict__ q, double* partials, PcgScal* sc) {
  double vals[D];
#pragma unroll
  for (int k = 0; k < D; k++) vals[k] = 0.0;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < N; v += gridDim.x * blockDim.x) {
    if (ptr[v + 1] > ptr[v]) {
      Vec<D> pi = ld_point<D>(p, v);
      Vec<D> qi = apply_row<D>(p, ptr, idx, v, pi);
      st_point<D>(q, v, qi);
#pragma unroll
      for (int k = 0; k < D; k++) vals[k] += pi.v[k] * qi.v[k];
    }
  }
  block_reduce_store<D>(vals, partials, sc->pq, &sc->ticket);
}

template <int D>
__global__ void __launch_bounds__(PB)
    k_pcg_update(double* x, double* r, const double* __restrict__ p,
                 const double* __restrict__ q, const int* __restrict__ ptr, int N, double* partials,
                 PcgScal* sc, const double* __restrict__ diag) {  // diag == nullptr: the degree
  double alpha[D];
#pragma unroll
  for (int k = 0; k < D; k++) alpha[k] = (sc->pq[k] != 0.0) ? sc->rz[k] / sc->pq[k] : 0.0;
  double vals[2 * D];
#pragma unroll
  for (int k = 0; k < 2 * D; k++) vals[k] = 0.0;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < N; v += gridDim.x * blockDim.x) {
    const int deg = ptr[v + 1] - ptr[v];
    if (deg > 0) {
      Vec<D> xi = ld_point_rw<D>(x, v), ri = ld_point_rw<D>(r, v), pi = ld_point<D>(p, v),
             qi = ld_point<D>(q, v);
      const double invd = 1.0 / (diag ? diag[v] : (double)deg);
#pragma unroll
      for (int k = 0; k < D; k++) {
        xi.v[k] += alpha[k] * pi.v[k];
        ri.v[k] -= alpha[k] * qi.v[k];
        vals[k] += ri.v[k] * ri.v[k] * invd;  // r.z
        vals[D + k] += ri.v[k] * ri.v[k];     // r.r
      }
      st_point<D>(x, v, xi);
      st_point<D>(r, v, ri);
    }
  }
  block_reduce_store<2 * D>(vals, partials, sc->upd, &sc->ticket);
}

// p = z + beta p with z = r / deg; then rz <- rz_new
template <int D>
__global__ void __launch_bounds__(PB)
    k_pcg_dir(double* p, const double* __restrict__ r, const int* __restrict__ ptr, int N,
              PcgScal* sc, const double* __restrict__ diag) {
  double beta[D];
#pragma unroll
  for (int k = 0; k < D; k++) beta[k] = (sc->rz[k] != 0.0) ? sc->upd[k] / sc->rz[k] : 0.0;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < N; v += gridDim.x * blockDim.x) {
    const int deg = ptr[v + 1] - ptr[v];
    if (deg > 0) {
      Vec<D> ri = ld_point<D>(r, v), pi = ld_point_rw<D>(p, v);
      const double invd = 1.0 / (diag ? diag[v] : (double)deg);
#pragma unroll
      for (int k = 0; k < D; k++) pi.v[k] = ri.v[k] * invd + beta[k] * pi.v[k];
      st_point<D>(p, v, pi);
    }
  }
  // rz is read by every block above, so the shift happens in the next kernel (k_shift)
}

template <int D>
__global__ void k_shift(PcgScal* sc, bool from_init) {
#pragma unroll
  for (int k = 0; k < D; k++) sc->rz[k] = from_init ? sc->ini[k] : sc->upd[k];
}

int build_neighbours(om_handle* h) {
  if (h->nbr_valid) return OM_OK;
  // the neighbour rows are indexed with 32-bit offsets (about 6 entries per vertex)
  if (h->N > (int64_t)268000000) {
    om_set_error("mesh too large for the solve methods (%lld vertices; limit 268 M)",
                 (long long)h->N);
    return OM_ERR_ARG;
  }
  const int N = (int)h->N;
  const int G = om_grid(N, 256);
  if (!h->nbr_ptr) CUDA_TRY(om_malloc(h, &h->nbr_ptr, sizeof(int) * (N + 1)));
  int* cnt = nullptr;
  CUDA_TRY(om_malloc(h, &cnt, sizeof(int) * N));
  OM_LAUNCH(h, (k_ring<false>), G, 256, h->cells, (const int*)h->adj, h->v2c, h->bflag, N, cnt,
            (int*)nullptr, h->ds);
  size_t bytes = 0;
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, bytes, cnt, h->nbr_ptr, N, h->stream));
  void* tmp = nullptr;
  CUDA_TRY(om_malloc(h, &tmp, bytes ? bytes : 1));
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp, bytes, cnt, h->nbr_ptr, N, h->stream));
  OM_LAUNCH(h, k_set_last, 1, 1, h->nbr_ptr, cnt, N);
  int nnz = 0;
  CUDA_TRY(cudaMemcpyAsync(&nnz, h->nbr_ptr + N, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  om_free(h, tmp);
  om_free(h, cnt);
  if (h->nbr_idx) om_free(h, h->nbr_idx);
  h->nbr_idx = nullptr;
  CUDA_TRY(om_malloc(h, &h->nbr_idx, sizeof(int) * std::max(nnz, 1)));
  h->nnz = nnz;
  OM_LAUNCH(h, (k_ring<true>), G, 256, h->cells, (const int*)h->adj, h->v2c, h->bflag, N,
            h->nbr_ptr, h->nbr_idx, h->ds);
  OM_TRY(om_fetch_scalars(h));
  OM_TRY(om_check_dev_err(h));
  h->nbr_valid = true;
  return OM_OK;
}

template <int D>
int pcg(om_handle* h, double rtol, int max_iter, int32_t* iters, double* relres, double* out) {
  const int N = (int)h->N;
  const size_t vec = sizeof(double) * (size_t)N * h->PD;
  OM_TRY(build_neighbours(h));
  if (!h->pcg_buf) CUDA_TRY(om_malloc(h, &h->pcg_buf, 3 * vec + sizeof(double) * 8 * MAXG + 256));
  double* r = h->pcg_buf;
  double* p = r + (size_t)N * h->PD;
  double* q = p + (size_t)N * h->PD;
  double* partials = q + (size_t)N * h->PD;
  PcgScal* sc = nullptr;
  CUDA_TRY(om_malloc(h, &sc, sizeof(PcgScal)));
  CUDA_TRY(cudaMemsetAsync(sc, 0, sizeof(PcgScal), h->stream));
  // the iterate lives in `out`; fixed vertices keep their coordinates
  if (out != h->x) CUDA_TRY(cudaMemcpyAsync(out, h->x, vec, cudaMemcpyDeviceToDevice, h->stream));
  const int G = std::min(om_grid(N, PB), MAXG);
  OM_LAUNCH(h, k_pcg_init<D>, G, PB, out, h->nbr_ptr, h->nbr_idx, h->bflag, N, r, p, partials, sc);
  OM_LAUNCH(h, k_shift<D>, 1, 1, sc, true);
  PcgScal hs;
  CUDA_TRY(cudaMemcpyAsync(&hs, sc, sizeof(PcgScal), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  double bb[3] = {0, 0, 0};
  for (int k = 0; k < D; k++) bb[k] = hs.ini[D + k];
  double rz0 = 0.0;
  for (int k = 0; k < D; k++) rz0 += hs.ini[k];
  if (bb[0] == 0.0 && bb[1] == 0.0 && bb[D - 1] == 0.0 && rz0 == 0.0) {
    // nothing to solve (no free vertex, or the points already satisfy the equations)
    om_free(h, sc);
    if (iters) *iters = 0;
    if (relres) *relres = 0.0;
    return OM_OK;
  }
  if (bb[0] == 0.0 && bb[1] == 0.0 && bb[D - 1] == 0.0) {
    // no free vertex has a fixed neighbour (closed surface, or nothing pinned): the graph
    // Laplacian is singular and "the solution" would be a collapsed mesh
    om_free(h, sc);
    om_set_error("cpt-linear-solve needs a boundary: no free vertex is next to a fixed one, "
                 "the Dirichlet graph Laplacian is singular");
    return OM_ERR_ARG;
  }
  int it = 0;
  double worst = INFINITY, best = INFINITY;
  int since_best = 0;
  bool stagnated = false;
  const int check = 25;
  // scale-free threshold: |r| <= rtol * max(|b|, tiny)
  while (it < max_iter) {
    for (int s = 0; s < check && it < max_iter; s++, it++) {
      OM_LAUNCH(h, k_pcg_spmv<D>, G, PB, p, h->nbr_ptr, h->nbr_idx, N, q, partials, sc);
      OM_LAUNCH(h, k_pcg_update<D>, G, PB, out, r, p, q, h->nbr_ptr, N, partials, sc,
                (const double*)nullptr);
      OM_LAUNCH(h, k_pcg_dir<D>, G, PB, p, r, h->nbr_ptr, N, sc, (const double*)nullptr);
      OM_LAUNCH(h, k_shift<D>, 1, 1, sc, false);
    }
    CUDA_TRY(cudaMemcpyAsync(&hs, sc, sizeof(PcgScal), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    worst = 0.0;
    for (int k = 0; k < D; k++) {
      double rr = hs.upd[D + k];
      double denom = bb[k] > 0.0 ? bb[k] : 1.0;
      worst = std::max(worst, sqrt(rr / denom));
    }
    if (!(worst > rtol)) break;
    // stagnation: the residual has not dropped by 1 % over the last 40 checks -- fp64 cannot
    // do better on this system; more iterations would only burn time
    if (worst < best * 0.99) {
      best = worst;
      since_best = 0;
    } else if (++since_best >= 40) {
      stagnated = true;
      break;
    }
  }
  om_free(h, sc);
  if (iters) *iters = it;
  if (relres) *relres = worst;
  CUDA_TRY(cudaGetLastError());
  // a residual that stagnates below 1e-9 is the fp64 floor of this system, not a failure
  // (the default rtol of 1e-13 is out of reach on meshes of millions of vertices)
  if (worst > rtol && !(stagnated && worst <= 1.0e-9)) {
    om_set_error("cpt-linear-solve: PCG stopped after %d iterations at relative residual %.3e "
                 "(requested %.3e); raise max_iter or rtol (om_set_solver)", it, worst, rtol);
    return OM_ERR_NOT_CONVERGED;
  }
  return OM_OK;
}

// cpt-quasi-newton: out = x + delta with H delta = -dE (see the header of this file)
template <int D>
int quasi_newton(om_handle* h, double rtol, int max_iter, int32_t* iters, double* relres,
                 double* out) {
  const int N = (int)h->N;
  const size_t vec = sizeof(double) * (size_t)N * h->PD;
  OM_TRY(build_neighbours(h));
  if (!h->pcg_buf) CUDA_TRY(om_malloc(h, &h->pcg_buf, 3 * vec + sizeof(double) * 8 * MAXG + 256));
  double* r = h->pcg_buf;
  double* p = r + (size_t)N * h->PD;
  double* q = p + (size_t)N * h->PD;
  double* partials = q + (size_t)N * h->PD;
  double *w = nullptr, *diag = nullptr;
  CUDA_TRY(om_malloc(h, &w, sizeof(double) * (size_t)std::max<int64_t>(h->nnz, 1)));
  CUDA_TRY(om_malloc(h, &diag, sizeof(double) * N));
  PcgScal* sc = nullptr;
  CUDA_TRY(om_malloc(h, &sc, sizeof(PcgScal)));
  CUDA_TRY(cudaMemsetAsync(sc, 0, sizeof(PcgScal), h->stream));
  // the iterate x + delta lives in `out`; fixed vertices keep their coordinates
  if (out != h->x) CUDA_TRY(cudaMemcpyAsync(out, h->x, vec, cudaMemcpyDeviceToDevice, h->stream));
  const int G = std::min(om_grid(N, PB), MAXG);
  OM_LAUNCH(h, k_qn_setup<D>, G, PB, h->x, h->nbr_ptr, h->nbr_idx, N, w, diag, r, p, partials, sc,
            h->ds);
  OM_LAUNCH(h, k_shift<D>, 1, 1, sc, true);
  PcgScal hs;
  CUDA_TRY(cudaMemcpyAsync(&hs, sc, sizeof(PcgScal), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  int rc = om_fetch_scalars(h);
  if (rc == OM_OK) rc = om_check_dev_err(h);  // degenerate cells
  double bb[3] = {0, 0, 0};
  for (int k = 0; k < D; k++) bb[k] = hs.ini[D + k];
  int it = 0;
  double worst = 0.0;
  bool nonzero = false;
  for (int k = 0; k < D; k++) nonzero = nonzero || bb[k] > 0.0;
  const int check = 10;
  while (rc == OM_OK && nonzero && it < max_iter) {
    for (int s2 = 0; s2 < check && it < max_iter; s2++, it++) {
      OM_LAUNCH(h, k_qn_spmv<D>, G, PB, p, h->nbr_ptr, h->nbr_idx, w, diag, N, q, partials, sc);
      OM_LAUNCH(h, k_pcg_update<D>, G, PB, out, r, p, q, h->nbr_ptr, N, partials, sc,
                (const double*)diag);
      OM_LAUNCH(h, k_pcg_dir<D>, G, PB, p, r, h->nbr_ptr, N, sc, (const double*)diag);
      OM_LAUNCH(h, k_shift<D>, 1, 1, sc, false);
    }
    CUDA_TRY(cudaMemcpyAsync(&hs, sc, sizeof(PcgScal), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    worst = 0.0;
    for (int k = 0; k < D; k++) {
      const double denom = bb[k] > 0.0 ? bb[k] : 1.0;
      worst = std::max(worst, sqrt(hs.upd[D + k] / denom));
    }
    if (!(worst > rtol)) break;
  }
  om_free(h, sc);
  om_free(h, w);
  om_free(h, diag);
  if (iters) *iters = it;
  if (relres) *relres = worst;
  if (rc != OM_OK) return rc;
  CUDA_TRY(cudaGetLastError());
  if (worst > rtol) {
    om_set_error("cpt-quasi-newton: PCG stopped after %d iterations at relative residual %.3e "
                 "(requested %.3e); raise max_iter or rtol (om_set_solver)", it, worst, rtol);
    return OM_ERR_NOT_CONVERGED;
  }
  return OM_OK;
}

}  // namespace

int om_quasi_newton_impl(om_handle* h, double rtol, int max_iter, int32_t* iters, double* relres,
                         double* out) {
  if (h->N == 0) return OM_OK;
  if (h->D == 2) return quasi_newton<2>(h, rtol, max_iter, iters, relres, out);
  return quasi_newton<3>(h, rtol, max_iter, iters, relres, out);
}

int om_pcg_impl(om_handle* h, double rtol, int max_iter, int32_t* iters, double* relres,
                double* out) {
  if (h->N == 0) return OM_OK;
  if (h->D == 2) return pcg<2>(h, rtol, max_iter, iters, relres, out);
  return pcg<3>(h, rtol, max_iter, iters, relres, out);
}
