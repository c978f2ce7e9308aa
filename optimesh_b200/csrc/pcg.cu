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
#include <cstdlib>
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

// ------------------------------------------------------------------ aggregation multigrid
// Jacobi-PCG needs O(sqrt(N)) iterations on the graph Laplacian (3,925 at 5 M vertices).  The
// preconditioner below is a V(2,2)-cycle of plain (unsmoothed) aggregation multigrid:
//   * vertices are Morton-sorted (setup.cu), so the aggregates are simply 4 consecutive ids,
//     i >> 2 -- spatially compact, no matching pass, restriction is a 4-lane shuffle;
//   * coarse operators are Galerkin products P^T A P with piecewise-constant P: the weight
//     between two aggregates is the number of fine edges between them (integers: exact in
//     fp64 whatever the summation order), built by one radix sort + reduce-by-key per level;
//   * damped Jacobi smoothing (omega 0.9), coarse corrections over-weighted by 1.6 (the usual
//     remedy for the poor approximation of piecewise-constant interpolation), the coarsest
//     level (<= 256 unknowns) solved by 64 Jacobi sweeps in one block.
// Every piece is a fixed symmetric linear operator, so plain PCG applies; every sum has a
// fixed order, so the iteration stays bitwise reproducible.  Measured on the oracle side
// (scipy prototype, square meshes): 63 / 99 iterations at 40 k / 250 k vertices with a V(1,1)
// cycle, 73 at 1 M with this V(2,2) cycle, against 578 / 1,420 / ~2,800 for Jacobi.
constexpr int MG_COARSEST = 256;
constexpr double MG_OMEGA = 0.9, MG_SCALE = 1.6;
constexpr int MG_COARSE_SWEEPS = 64;

struct MgLevel {
  int n = 0, nnz = 0;
  const int* ptr = nullptr;
  const int* col = nullptr;
  const double* val = nullptr;  // nullptr: every weight is 1 (level 0 is matrix free)
  double* diag = nullptr;       // 0 marks a row that is not an unknown
  double *r = nullptr, *x = nullptr, *y = nullptr;  // n * PD each
  std::vector<void*> owned;
};

__global__ void k_mg_diag0(const int* __restrict__ ptr, int n, double* __restrict__ diag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) diag[i] = (double)(ptr[i + 1] - ptr[i]);
}

// one key per matrix entry: (coarse row, coarse column) of an entry that couples two
// aggregates, all ones for the others (they sort to the end)
__global__ void k_mg_emit(int n, const int* __restrict__ ptr, const int* __restrict__ col,
                          const double* __restrict__ val, const double* __restrict__ diag,
                          unsigned long long* __restrict__ keys, double* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool free_i = diag[i] > 0.0;
  for (int q = ptr[i]; q < ptr[i + 1]; q++) {
    const int j = col[q];
    const bool keep = free_i && (j >> 2) != (i >> 2) && diag[j] > 0.0;
    keys[q] = keep ? (((unsigned long long)(i >> 2) << 32) | (unsigned)(j >> 2)) : ~0ull;
    vals[q] = val ? val[q] : 1.0;
  }
}

// diagonal of P^T A P: the diagonals of the aggregate minus the weights inside it
__global__ void k_mg_coarse_diag(int n, int nc, const int* __restrict__ ptr,
                                 const int* __restrict__ col, const double* __restrict__ val,
                                 const double* __restrict__ diag, double* __restrict__ diag_c) {
  const int I = blockIdx.x * blockDim.x + threadIdx.x;
  if (I >= nc) return;
  double s = 0.0;
  for (int i = 4 * I; i < min(4 * I + 4, n); i++) {
    if (!(diag[i] > 0.0)) continue;
    s += diag[i];
    for (int q = ptr[i]; q < ptr[i + 1]; q++) {
      const int j = col[q];
      if ((j >> 2) == I && diag[j] > 0.0) s -= val ? val[q] : 1.0;
    }
  }
  diag_c[I] = s > 0.0 ? s : 0.0;
}

__global__ void k_mg_count(const unsigned long long* __restrict__ ukeys,
                           const int* __restrict__ num, int* __restrict__ nnz_out) {
  const int m = *num;
  *nnz_out = (m > 0 && ukeys[m - 1] == ~0ull) ? m - 1 : m;
}

// CSR of the coarse level from the sorted unique keys
__global__ void k_mg_rows(int nc, int nnz, const unsigned long long* __restrict__ ukeys,
                          int* __restrict__ ptr_c, int* __restrict__ col_c) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nnz) col_c[t] = (int)(unsigned)(ukeys[t] & 0xffffffffull);
  if (t <= nc) {
    const unsigned long long want = (unsigned long long)t << 32;
    int lo = 0, hi = nnz;  // first entry with key >= want
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (ukeys[mid] < want) lo = mid + 1; else hi = mid;
    }
    ptr_c[t] = lo;
  }
}

template <int D>
__device__ __forceinline__ Vec<D> mg_offdiag(const int* __restrict__ ptr,
                                             const int* __restrict__ col,
                                             const double* __restrict__ val,
                                             const double* __restrict__ x, int i) {
  Vec<D> s;
#pragma unroll
  for (int k = 0; k < D; k++) s.v[k] = 0.0;
  for (int q = ptr[i]; q < ptr[i + 1]; q++) {
    const Vec<D> xj = ld_point<D>(x, __ldg(col + q));
    const double w = val ? __ldg(val + q) : 1.0;
#pragma unroll
    for (int k = 0; k < D; k++) s.v[k] = fma(w, xj.v[k], s.v[k]);
  }
  return s;
}

// first sweep from a zero guess: x = omega r / diag
template <int D>
__global__ void __launch_bounds__(PB)
    k_mg_smooth0(int n, const double* __restrict__ diag, const double* __restrict__ r,
                 double* __restrict__ x) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double d = diag[i];
    if (d > 0.0) {
      const Vec<D> ri = ld_point<D>(r, i);
      Vec<D> xi;
      const double f = MG_OMEGA / d;
#pragma unroll
      for (int k = 0; k < D; k++) xi.v[k] = f * ri.v[k];
      st_point<D>(x, i, xi);
    }
  }
}

// damped Jacobi sweep: xout = xin + omega (r - A xin) / diag
template <int D>
__global__ void __launch_bounds__(PB)
    k_mg_smooth(int n, const int* __restrict__ ptr, const int* __restrict__ col,
                const double* __restrict__ val, const double* __restrict__ diag,
                const double* __restrict__ r, const double* __restrict__ xin,
                double* __restrict__ xout) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double d = diag[i];
    if (d > 0.0) {
      const Vec<D> ri = ld_point<D>(r, i), xi = ld_point<D>(xin, i);
      const Vec<D> s = mg_offdiag<D>(ptr, col, val, xin, i);
      Vec<D> xo;
      const double f = MG_OMEGA / d;
#pragma unroll
      for (int k = 0; k < D; k++) xo.v[k] = fma(f, ri.v[k] - d * xi.v[k] + s.v[k], xi.v[k]);
      st_point<D>(xout, i, xo);
    }
  }
}

// residual of the fine level summed over each aggregate (4 consecutive rows = 4 lanes)
template <int D>
__global__ void __launch_bounds__(PB)
    k_mg_restrict(int n, const int* __restrict__ ptr, const int* __restrict__ col,
                  const double* __restrict__ val, const double* __restrict__ diag,
                  const double* __restrict__ r, const double* __restrict__ x,
                  double* __restrict__ rc) {
  // whole warps stay in the loop (the shuffles below name all 32 lanes)
  const int n32 = (n + 31) & ~31;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n32; i += gridDim.x * blockDim.x) {
    Vec<D> res;
#pragma unroll
    for (int k = 0; k < D; k++) res.v[k] = 0.0;
    if (i < n) {
      const double d = diag[i];
      if (d > 0.0) {
        const Vec<D> ri = ld_point<D>(r, i), xi = ld_point<D>(x, i);
        const Vec<D> s = mg_offdiag<D>(ptr, col, val, x, i);
#pragma unroll
        for (int k = 0; k < D; k++) res.v[k] = ri.v[k] - d * xi.v[k] + s.v[k];
      }
    }
    // (blockDim and the grid stride are multiples of 32: the four lanes of an aggregate are
    // always in the loop together)
#pragma unroll
    for (int k = 0; k < D; k++) {
      res.v[k] += __shfl_xor_sync(0xffffffffu, res.v[k], 1);
      res.v[k] += __shfl_xor_sync(0xffffffffu, res.v[k], 2);
    }
    if ((i & 3) == 0 && i < n) st_point<D>(rc, i >> 2, res);
  }
}

// x += scale * (coarse correction of the aggregate)
template <int D>
__global__ void __launch_bounds__(PB)
    k_mg_prolong(int n, const double* __restrict__ diag, double* __restrict__ x,
                 const double* __restrict__ xc) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (diag[i] > 0.0) {
      Vec<D> xi = ld_point_rw<D>(x, i);
      const Vec<D> c = ld_point<D>(xc, i >> 2);
#pragma unroll
      for (int k = 0; k < D; k++) xi.v[k] = fma(MG_SCALE, c.v[k], xi.v[k]);
      st_point<D>(x, i, xi);
    }
  }
}

// coarsest level: MG_COARSE_SWEEPS Jacobi sweeps from zero in one block (n <= MG_COARSEST)
template <int D>
__global__ void __launch_bounds__(MG_COARSEST)
    k_mg_coarsest(int n, const int* __restrict__ ptr, const int* __restrict__ col,
                  const double* __restrict__ val, const double* __restrict__ diag,
                  const double* __restrict__ r, double* __restrict__ x) {
  __shared__ double buf[2][MG_COARSEST][D];
  const int i = threadIdx.x;
  const double d = i < n ? diag[i] : 0.0;
  Vec<D> ri;
#pragma unroll
  for (int k = 0; k < D; k++) ri.v[k] = 0.0;
  if (d > 0.0) ri = ld_point<D>(r, i);
  const double f = d > 0.0 ? MG_OMEGA / d : 0.0;
#pragma unroll
  for (int k = 0; k < D; k++) buf[0][i][k] = f * ri.v[k];
  __syncthreads();
  int cur = 0;
  for (int s = 1; s < MG_COARSE_SWEEPS; s++) {
    double xo[D];
#pragma unroll
    for (int k = 0; k < D; k++) xo[k] = 0.0;
    if (d > 0.0) {
      double acc[D];
#pragma unroll
      for (int k = 0; k < D; k++) acc[k] = ri.v[k] - d * buf[cur][i][k];
      for (int q = ptr[i]; q < ptr[i + 1]; q++) {
        const int j = col[q];
        const double w = val ? val[q] : 1.0;
#pragma unroll
        for (int k = 0; k < D; k++) acc[k] = fma(w, buf[cur][j][k], acc[k]);
      }
#pragma unroll
      for (int k = 0; k < D; k++) xo[k] = fma(f, acc[k], buf[cur][i][k]);
    }
#pragma unroll
    for (int k = 0; k < D; k++) buf[cur ^ 1][i][k] = xo[k];
    __syncthreads();
    cur ^= 1;
  }
  if (i < n) {
    Vec<D> xi;
#pragma unroll
    for (int k = 0; k < D; k++) xi.v[k] = buf[cur][i][k];
    st_point<D>(x, i, xi);
  }
}

// x += alpha p, r -= alpha q, r.r (the preconditioned product r.z follows the V-cycle)
template <int D>
__global__ void __launch_bounds__(PB)
    k_mg_update(double* x, double* r, const double* __restrict__ p, const double* __restrict__ q,
                const double* __restrict__ diag, int N, double* partials, PcgScal* sc) {
  double alpha[D];
#pragma unroll
  for (int k = 0; k < D; k++) alpha[k] = (sc->pq[k] != 0.0) ? sc->rz[k] / sc->pq[k] : 0.0;
  double vals[D];
#pragma unroll
  for (int k = 0; k < D; k++) vals[k] = 0.0;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < N; v += gridDim.x * blockDim.x) {
    if (diag[v] > 0.0) {
      Vec<D> xi = ld_point_rw<D>(x, v), ri = ld_point_rw<D>(r, v);
      const Vec<D> pi = ld_point<D>(p, v), qi = ld_point<D>(q, v);
#pragma unroll
      for (int k = 0; k < D; k++) {
        xi.v[k] += alpha[k] * pi.v[k];
        ri.v[k] -= alpha[k] * qi.v[k];
        vals[k] += ri.v[k] * ri.v[k];
      }
      st_point<D>(x, v, xi);
      st_point<D>(r, v, ri);
    }
  }
  block_reduce_store<D>(vals, partials, sc->upd + D, &sc->ticket);
}

template <int D>
__global__ void __launch_bounds__(PB)
    k_mg_dot(const double* __restrict__ r, const double* __restrict__ z,
             const double* __restrict__ diag, int N, double* partials, PcgScal* sc) {
  double vals[D];
#pragma unroll
  for (int k = 0; k < D; k++) vals[k] = 0.0;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < N; v += gridDim.x * blockDim.x) {
    if (diag[v] > 0.0) {
      const Vec<D> ri = ld_point<D>(r, v), zi = ld_point<D>(z, v);
#pragma unroll
      for (int k = 0; k < D; k++) vals[k] += ri.v[k] * zi.v[k];
    }
  }
  block_reduce_store<D>(vals, partials, sc->upd, &sc->ticket);
}

// p = z + beta p (beta = 0 while rz is still 0: the first direction)
template <int D>
__global__ void __launch_bounds__(PB)
    k_mg_dir(double* p, const double* __restrict__ z, const double* __restrict__ diag, int N,
             PcgScal* sc) {
  double beta[D];
#pragma unroll
  for (int k = 0; k < D; k++) beta[k] = (sc->rz[k] != 0.0) ? sc->upd[k] / sc->rz[k] : 0.0;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < N; v += gridDim.x * blockDim.x) {
    if (diag[v] > 0.0) {
      const Vec<D> zi = ld_point<D>(z, v);
      Vec<D> pi = ld_point_rw<D>(p, v);
#pragma unroll
      for (int k = 0; k < D; k++) pi.v[k] = zi.v[k] + beta[k] * pi.v[k];
      st_point<D>(p, v, pi);
    }
  }
}

__global__ void k_mg_zero_rz(PcgScal* sc) {
  for (int k = 0; k < 3; k++) sc->rz[k] = 0.0;
}

void mg_free(om_handle* h, std::vector<MgLevel>& lv) {
  for (auto& l : lv)
    for (void* p : l.owned) om_free(h, p);
  lv.clear();
}

// builds the hierarchy under level 0 (= the neighbour rows of the handle)
int mg_build(om_handle* h, std::vector<MgLevel>& lv) {
  const int N = (int)h->N, PD = h->PD;
  auto alloc = [&](MgLevel& l, auto** p, size_t bytes) -> cudaError_t {
    cudaError_t e = om_malloc(h, p, std::max<size_t>(bytes, 16));
    if (e == cudaSuccess) l.owned.push_back((void*)*p);
    return e;
  };
  lv.emplace_back();
  {
    MgLevel& l0 = lv.back();
    l0.n = N;
    l0.nnz = (int)h->nnz;
    l0.ptr = h->nbr_ptr;
    l0.col = h->nbr_idx;
    CUDA_TRY(alloc(l0, &l0.diag, sizeof(double) * N));
    CUDA_TRY(alloc(l0, &l0.x, sizeof(double) * (size_t)N * PD));
    CUDA_TRY(alloc(l0, &l0.y, sizeof(double) * (size_t)N * PD));
    CUDA_TRY(cudaMemsetAsync(l0.x, 0, sizeof(double) * (size_t)N * PD, h->stream));
    CUDA_TRY(cudaMemsetAsync(l0.y, 0, sizeof(double) * (size_t)N * PD, h->stream));
    OM_LAUNCH(h, k_mg_diag0, om_grid(N, 256), 256, l0.ptr, N, l0.diag);
  }
  // scratch of the Galerkin products, sized for level 0 (the largest)
  const size_t m0 = (size_t)std::max(lv[0].nnz, 1);
  unsigned long long *keys = nullptr, *skeys = nullptr, *ukeys = nullptr;
  double *vals = nullptr, *svals = nullptr;
  int* dnum = nullptr;
  CUDA_TRY(om_malloc(h, &keys, 8 * m0));
  CUDA_TRY(om_malloc(h, &skeys, 8 * m0));
  CUDA_TRY(om_malloc(h, &ukeys, 8 * m0));
  CUDA_TRY(om_malloc(h, &vals, 8 * m0));
  CUDA_TRY(om_malloc(h, &svals, 8 * m0));
  CUDA_TRY(om_malloc(h, &dnum, 2 * sizeof(int)));
  size_t b1 = 0, b2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, b1, keys, skeys, vals, svals, (int)m0, 0, 64,
                                  h->stream);
  cub::DeviceReduce::ReduceByKey(nullptr, b2, skeys, ukeys, svals, vals, dnum, cub::Sum(),
                                 (int)m0, h->stream);
  void* tmp = nullptr;
  CUDA_TRY(om_malloc(h, &tmp, std::max(b1, b2) + 16));
  size_t tmp_bytes = std::max(b1, b2) + 16;
  int rc = OM_OK;
  while (lv.back().n > MG_COARSEST && lv.size() < 24) {
    const MgLevel f = lv.back();  // (copy: the vector grows below)
    const int nc = (f.n + 3) >> 2;
    MgLevel c;
    c.n = nc;
    auto fail = [&](cudaError_t e) {
      om_set_error("CUDA error %s in the multigrid setup: %s", cudaGetErrorName(e),
                   cudaGetErrorString(e));
      rc = OM_ERR_CUDA;
      for (void* p : c.owned) om_free(h, p);  // the level under construction is not in lv yet
      c.owned.clear();
    };
    cudaError_t e;
    if ((e = alloc(c, &c.diag, sizeof(double) * nc)) != cudaSuccess) { fail(e); break; }
    OM_LAUNCH(h, k_mg_coarse_diag, om_grid(nc, 256), 256, f.n, nc, f.ptr, f.col, f.val, f.diag,
              c.diag);
    int nnz_c = 0;
    if (f.nnz > 0) {
      OM_LAUNCH(h, k_mg_emit, om_grid(f.n, 256), 256, f.n, f.ptr, f.col, f.val, f.diag, keys,
                vals);
      // keys: coarse row in the high word (< 2^30), column in the low word; all ones last
      size_t tb = tmp_bytes;
      if ((e = cub::DeviceRadixSort::SortPairs(tmp, tb, keys, skeys, vals, svals, f.nnz, 0, 64,
                                               h->stream)) != cudaSuccess) { fail(e); break; }
      tb = tmp_bytes;
      if ((e = cub::DeviceReduce::ReduceByKey(tmp, tb, skeys, ukeys, svals, vals, dnum,
                                              cub::Sum(), f.nnz, h->stream)) != cudaSuccess) {
        fail(e);
        break;
      }
      OM_LAUNCH(h, k_mg_count, 1, 1, ukeys, dnum, dnum + 1);
      if ((e = cudaMemcpyAsync(&nnz_c, dnum + 1, sizeof(int), cudaMemcpyDeviceToHost,
                               h->stream)) != cudaSuccess ||
          (e = cudaStreamSynchronize(h->stream)) != cudaSuccess) { fail(e); break; }
    }
    c.nnz = nnz_c;
    int *ptr_c = nullptr, *col_c = nullptr;
    double* val_c = nullptr;
    if ((e = alloc(c, &ptr_c, sizeof(int) * (nc + 1))) != cudaSuccess ||
        (e = alloc(c, &col_c, sizeof(int) * std::max(nnz_c, 1))) != cudaSuccess ||
        (e = alloc(c, &val_c, sizeof(double) * std::max(nnz_c, 1))) != cudaSuccess ||
        (e = alloc(c, &c.r, sizeof(double) * (size_t)nc * PD)) != cudaSuccess ||
        (e = alloc(c, &c.x, sizeof(double) * (size_t)nc * PD)) != cudaSuccess ||
        (e = alloc(c, &c.y, sizeof(double) * (size_t)nc * PD)) != cudaSuccess) { fail(e); break; }
    OM_LAUNCH(h, k_mg_rows, om_grid(std::max(nnz_c, nc + 1), 256), 256, nc, nnz_c, ukeys, ptr_c,
              col_c);
    if (nnz_c > 0)
      cudaMemcpyAsync(val_c, vals, sizeof(double) * nnz_c, cudaMemcpyDeviceToDevice, h->stream);
    cudaMemsetAsync(c.r, 0, sizeof(double) * (size_t)nc * PD, h->stream);
    cudaMemsetAsync(c.x, 0, sizeof(double) * (size_t)nc * PD, h->stream);
    cudaMemsetAsync(c.y, 0, sizeof(double) * (size_t)nc * PD, h->stream);
    c.ptr = ptr_c;
    c.col = col_c;
    c.val = val_c;
    lv.push_back(c);
  }
  om_free(h, tmp);
  om_free(h, dnum);
  om_free(h, svals);
  om_free(h, vals);
  om_free(h, ukeys);
  om_free(h, skeys);
  om_free(h, keys);
  if (rc != OM_OK) return rc;
  CUDA_TRY(cudaGetLastError());
  return OM_OK;
}

// z = M r: one V(2,2) cycle; the result is lv[0].y
template <int D>
int mg_vcycle(om_handle* h, std::vector<MgLevel>& lv) {
  const int L = (int)lv.size();
  auto grid = [](int n) { return std::min(om_grid(n, PB), MAXG); };
  for (int l = 0; l < L; l++) {
    MgLevel& a = lv[l];
    if (l == L - 1 && a.n <= MG_COARSEST) {
      OM_LAUNCH(h, k_mg_coarsest<D>, 1, MG_COARSEST, a.n, a.ptr, a.col, a.val, a.diag, a.r, a.y);
      break;
    }
    OM_LAUNCH(h, k_mg_smooth0<D>, grid(a.n), PB, a.n, a.diag, a.r, a.x);
    OM_LAUNCH(h, k_mg_smooth<D>, grid(a.n), PB, a.n, a.ptr, a.col, a.val, a.diag, a.r, a.x, a.y);
    if (l == L - 1) break;  // (hierarchy cut short: the last level is only smoothed)
    OM_LAUNCH(h, k_mg_restrict<D>, grid((a.n + 31) & ~31), PB, a.n, a.ptr, a.col, a.val, a.diag,
              a.r, a.y, lv[l + 1].r);
  }
  for (int l = L - 2; l >= 0; l--) {
    MgLevel& a = lv[l];
    OM_LAUNCH(h, k_mg_prolong<D>, grid(a.n), PB, a.n, a.diag, a.y, lv[l + 1].y);
    OM_LAUNCH(h, k_mg_smooth<D>, grid(a.n), PB, a.n, a.ptr, a.col, a.val, a.diag, a.r, a.y, a.x);
    OM_LAUNCH(h, k_mg_smooth<D>, grid(a.n), PB, a.n, a.ptr, a.col, a.val, a.diag, a.r, a.x, a.y);
  }
  CUDA_TRY(cudaGetLastError());
  return OM_OK;
}

// cpt-linear-solve with the multigrid preconditioner
template <int D>
int pcg_mg(om_handle* h, double rtol, int max_iter, int32_t* iters, double* relres, double* out) {
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
  if (out != h->x) CUDA_TRY(cudaMemcpyAsync(out, h->x, vec, cudaMemcpyDeviceToDevice, h->stream));
  std::vector<MgLevel> lv;
  int rc = mg_build(h, lv);
  if (rc != OM_OK) {
    mg_free(h, lv);
    om_free(h, sc);
    return rc;
  }
  lv[0].r = r;
  const double* diag = lv[0].diag;
  const int G = std::min(om_grid(N, PB), MAXG);
  auto finish = [&](int code) {
    mg_free(h, lv);
    om_free(h, sc);
    return code;
  };
  OM_LAUNCH(h, k_pcg_init<D>, G, PB, out, h->nbr_ptr, h->nbr_idx, h->bflag, N, r, p, partials, sc);
  PcgScal hs;
  if (cudaMemcpyAsync(&hs, sc, sizeof(PcgScal), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
      cudaStreamSynchronize(h->stream) != cudaSuccess) {
    om_set_error("CUDA error in the multigrid solve (initial residual)");
    return finish(OM_ERR_CUDA);
  }
  double bb[3] = {0, 0, 0};
  for (int k = 0; k < D; k++) bb[k] = hs.ini[D + k];
  double rz0 = 0.0;
  for (int k = 0; k < D; k++) rz0 += hs.ini[k];
  const bool no_b = bb[0] == 0.0 && bb[1] == 0.0 && bb[D - 1] == 0.0;
  if (no_b && rz0 == 0.0) {
    if (iters) *iters = 0;
    if (relres) *relres = 0.0;
    return finish(OM_OK);
  }
  if (no_b) {
    om_set_error("cpt-linear-solve needs a boundary: no free vertex is next to a fixed one, "
                 "the Dirichlet graph Laplacian is singular");
    return finish(OM_ERR_ARG);
  }
  // first direction: p = z = M r, rz = r.z
  OM_LAUNCH(h, k_mg_zero_rz, 1, 1, sc);
  if ((rc = mg_vcycle<D>(h, lv)) != OM_OK) return finish(rc);
  OM_LAUNCH(h, k_mg_dot<D>, G, PB, r, lv[0].y, diag, N, partials, sc);
  OM_LAUNCH(h, k_mg_dir<D>, G, PB, p, lv[0].y, diag, N, sc);
  OM_LAUNCH(h, k_shift<D>, 1, 1, sc, false);
  int it = 0;
  double worst = INFINITY, best = INFINITY;
  int since_best = 0;
  bool stagnated = false;
  const int check = 5;
  while (it < max_iter) {
    for (int s = 0; s < check && it < max_iter; s++, it++) {
      OM_LAUNCH(h, k_pcg_spmv<D>, G, PB, p, h->nbr_ptr, h->nbr_idx, N, q, partials, sc);
      OM_LAUNCH(h, k_mg_update<D>, G, PB, out, r, p, q, diag, N, partials, sc);
      if ((rc = mg_vcycle<D>(h, lv)) != OM_OK) return finish(rc);
      OM_LAUNCH(h, k_mg_dot<D>, G, PB, r, lv[0].y, diag, N, partials, sc);
      OM_LAUNCH(h, k_mg_dir<D>, G, PB, p, lv[0].y, diag, N, sc);
      OM_LAUNCH(h, k_shift<D>, 1, 1, sc, false);
    }
    if (cudaMemcpyAsync(&hs, sc, sizeof(PcgScal), cudaMemcpyDeviceToHost, h->stream) !=
            cudaSuccess ||
        cudaStreamSynchronize(h->stream) != cudaSuccess) {
      om_set_error("CUDA error in the multigrid solve");
      return finish(OM_ERR_CUDA);
    }
    worst = 0.0;
    for (int k = 0; k < D; k++) {
      const double denom = bb[k] > 0.0 ? bb[k] : 1.0;
      worst = std::max(worst, sqrt(hs.upd[D + k] / denom));
    }
    if (!(worst > rtol)) break;
    if (worst < best * 0.99) {
      best = worst;
      since_best = 0;
    } else if (++since_best >= 20) {
      stagnated = true;
      break;
    }
  }
  if (iters) *iters = it;
  if (relres) *relres = worst;
  if (cudaGetLastError() != cudaSuccess) {
    om_set_error("CUDA error in the multigrid solve (launch)");
    return finish(OM_ERR_CUDA);
  }
  if (worst > rtol && !(stagnated && worst <= 1.0e-9)) {
    om_set_error("cpt-linear-solve: PCG stopped after %d iterations at relative residual %.3e "
                 "(requested %.3e); raise max_iter or rtol (om_set_solver)", it, worst, rtol);
    return finish(OM_ERR_NOT_CONVERGED);
  }
  return finish(OM_OK);
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
  // small systems: the hierarchy costs more than Jacobi-PCG needs; OM_PCG_JACOBI=1 keeps the
  // plain iteration for comparison
  static const bool jacobi_only = getenv("OM_PCG_JACOBI") != nullptr;
  if (jacobi_only || h->N < 20000) {
    if (h->D == 2) return pcg<2>(h, rtol, max_iter, iters, relres, out);
    return pcg<3>(h, rtol, max_iter, iters, relres, out);
  }
  if (h->D == 2) return pcg_mg<2>(h, rtol, max_iter, iters, relres, out);
  return pcg_mg<3>(h, rtol, max_iter, iters, relres, out);
}
