// K1: the fused smoothing step.  One thread per vertex walks the vertex star through the
// half-edge twin table (a deterministic vertex-centric gather: no atomics on coordinates),
// recomputes the fp64 geometry of every incident cell, applies the method formula, pins the
// boundary, relaxes with omega, limits the step to half the smallest incident inradius and
// writes the new position -- one pass over the mesh per step.
//
// Replaces, per step: get_new_points of the five methods
// (/root/reference/README.md:80, :90, :104, :141), the numpy scatter-adds
// (np.bincount / np.add.at / np.minimum.at) and the body of the optimize() loop
// (README.md:131-132).  Arithmetic: SURVEY.md Appendix A.2-A.5, A.8, A.9.
//
// fp64 budget (the kernel is fp64-pipe/latency bound, not HBM bound): one rsqrt per cell
// visit and no division -- 1/(4A) = rsqrt(16 A^2), the circumcentre weights use
// sum_k ee_k ed_k = -8 A^2, thirds are applied once per vertex.  The inradius (3 sqrt +
// 1 div per cell) is only needed where the limiter bites, so the LAZY variant first tests
// the division-free bound r_in^2 >= 4 A^2 / (3 sum ee) and re-walks the star exactly only
// for vertices that fail it; both variants produce identical bits.
#include <cstring>
#include <utility>

#include "common.cuh"
#include "geom.cuh"

namespace {

template <int D>
struct Acc {
  double w;                   // control volume (Lloyd/CVT) or summed cell area (CPT/ODT)
  Vec<D> num;                 // 3 x weighted offsets from the vertex (thirds applied at the end)
  double H[D * (D + 1) / 2];  // CVT block: sum -0.5 ce_k e_k e_k^T (upper triangle)
  double rmin;                // EXACT: smallest incident inradius
  double lb_num, lb_den;      // LAZY: cell minimising A^2 / sum(ee) (kept as a fraction)
};

// exact inradius of one cell (A.3): 2A / (l0 + l1 + l2)
template <int D>
__device__ __forceinline__ double inradius(const CellGeo<D>& g) {
  const double A = sqrt(g.vol2);
  return 2.0 * A / (sqrt(g.ee0) + sqrt(g.ee1) + sqrt(g.ee2));
}

// Contribution of one incident cell to vertex P0 (P1, P2 follow in slot order).
template <int D, int METHOD, bool EXACT>
__device__ __forceinline__ void accumulate_cell(const Vec<D>& P0, const Vec<D>& P1,
                                                const Vec<D>& P2, Acc<D>& a, int& err) {
  const CellGeo<D> g = cell_geo<D>(P0, P1, P2);
  if (!(g.vol2 > 0.0)) {
    err |= OM_DEV_DEGENERATE;
    return;
  }
  if (EXACT) {
    a.rmin = fmin(a.rmin, inradius<D>(g));
  } else {
    const double S = g.ee0 + g.ee1 + g.ee2;
    if (g.vol2 * a.lb_den < a.lb_num * S) {
      a.lb_num = g.vol2;
      a.lb_den = S;
    }
  }
  const double r = rsqrt(g.vol2);  // 1/A
  if (METHOD == OM_CPT_FIXED_POINT) {
    // 3 (barycenter - P0) = e2 - e1
    const double A = g.vol2 * r;
    a.w += A;
#pragma unroll
    for (int k = 0; k < D; k++) a.num.v[k] += A * (g.e2.v[k] - g.e1.v[k]);
    return;
  }
  // circumcenter - P0 = alpha1 (P1 - P0) + alpha2 (P2 - P0) = alpha1 e2 - alpha2 e1 with
  // alpha_k = ee_k ed_k / sum_j ee_j ed_j and sum_j ee_j ed_j = -8 A^2
  const double inva = -0.125 * (r * r);
  const double al1 = g.ee1 * g.ed1 * inva, al2 = g.ee2 * g.ed2 * inva;
  if (METHOD == OM_ODT_FIXED_POINT) {
    const double A = g.vol2 * r;
    a.w += A;
    const double s2 = 3.0 * A * al1, s1 = 3.0 * A * al2;
#pragma unroll
    for (int k = 0; k < D; k++) a.num.v[k] += s2 * g.e2.v[k] - s1 * g.e1.v[k];
    return;
  }
  // Lloyd / CVT block-diagonal (A.4, A.9)
  const double inv4A = 0.25 * r;
  const double ce0 = -g.ed0 * inv4A, ce1 = -g.ed1 * inv4A, ce2 = -g.ed2 * inv4A;
  if (ce0 < -0.5 || ce1 < -0.5 || ce2 < -0.5) return;  // cell masked (an angle > 135 deg)
  const double part1 = 0.25 * g.ee1 * ce1, part2 = 0.25 * g.ee2 * ce2;
  const double pw = part1 + part2;
  a.w += pw;
  // 3 x sub-triangle centroids relative to P0: part_k ((m_k - P0) + (cc - P0)) with
  // m_1 - P0 = -e1/2, m_2 - P0 = e2/2, cc - P0 = al1 e2 - al2 e1
  const double s2 = pw * al1 + 0.5 * part2, s1 = pw * al2 + 0.5 * part1;
#pragma unroll
  for (int k = 0; k < D; k++) a.num.v[k] += s2 * g.e2.v[k] - s1 * g.e1.v[k];
  if (METHOD == OM_CVT_BLOCK_DIAGONAL) {
    const double h1 = -0.5 * ce1, h2 = -0.5 * ce2;
    int q = 0;
#pragma unroll
    for (int i = 0; i < D; i++)
#pragma unroll
      for (int j = i; j < D; j++) {
        a.H[q] += h1 * g.e1.v[i] * g.e1.v[j] + h2 * g.e2.v[i] * g.e2.v[j];
        q++;
      }
  }
}

template <int D>
__device__ __forceinline__ bool solve_sym(const double* H, double diag, const Vec<D>& rhs,
                                          Vec<D>& out);
template <>
__device__ __forceinline__ bool solve_sym<2>(const double* H, double diag, const Vec<2>& rhs,
                                             Vec<2>& out) {
  const double a = H[0] + diag, b = H[1], d = H[2] + diag;
  const double det = a * d - b * b;
  if (det == 0.0) return false;
  const double inv = 1.0 / det;
  out.v[0] = (d * rhs.v[0] - b * rhs.v[1]) * inv;
  out.v[1] = (a * rhs.v[1] - b * rhs.v[0]) * inv;
  return true;
}
template <>
__device__ __forceinline__ bool solve_sym<3>(const double* H, double diag, const Vec<3>& rhs,
                                             Vec<3>& out) {
  const double a = H[0] + diag, b = H[1], c = H[2], d = H[3] + diag, e = H[4], f = H[5] + diag;
  const double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
  const double det = a * c00 + b * c01 + c * c02;
  if (det == 0.0) return false;
  const double inv = 1.0 / det;
  const double c11 = a * f - c * c, c12 = b * c - a * e, c22 = a * d - b * b;
  out.v[0] = (c00 * rhs.v[0] + c01 * rhs.v[1] + c02 * rhs.v[2]) * inv;
  out.v[1] = (c01 * rhs.v[0] + c11 * rhs.v[1] + c12 * rhs.v[2]) * inv;
  out.v[2] = (c02 * rhs.v[0] + c12 * rhs.v[1] + c22 * rhs.v[2]) * inv;
  return true;
}

struct StepParams {
  const double* x;
  double* xout;
  const int4* cells;
  const int* adj;  // flat view of the int4 twin table: adj[4*c + k]
  const int* v2c;
  const uint8_t* bflag;
  int N;
  int lo, hi;  // vertices [lo, hi) are processed
  double omega;
  int limiter;
  DevScalars* ds;
};

// Block-level reduction of the step statistics (256 threads): one conditional atomic per
// block on the shared scalars (max and integer add are order independent).
__device__ __forceinline__ void reduce_step_stats(double diff2, int limited, DevScalars* ds) {
  __shared__ double s_d[8];
  __shared__ int s_l[8];
  for (int o = 16; o > 0; o >>= 1) {
    diff2 = fmax(diff2, __shfl_xor_sync(0xffffffffu, diff2, o));
    limited += __shfl_xor_sync(0xffffffffu, limited, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s_d[threadIdx.x >> 5] = diff2;
    s_l[threadIdx.x >> 5] = limited;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double d = s_d[0];
    int l = s_l[0];
#pragma unroll
    for (int w = 1; w < 8; w++) {
      d = fmax(d, s_d[w]);
      l += s_l[w];
    }
    if (d > 0.0) {
      const unsigned long long bits = (unsigned long long)__double_as_longlong(d);
      if (bits > *(volatile unsigned long long*)&ds->max_diff2_bits)
        atomicMax(&ds->max_diff2_bits, bits);
    }
    if (l) atomicAdd(&ds->n_limited, (unsigned long long)l);
  }
}

constexpr int MAX_RING = 4096;

// Visits every cell around vertex v, starting at cell c0 (where v sits in slot j) and
// leaving through local edge (j+1)%3; an open fan (boundary vertex) is completed from the
// start cell in the other direction.  f(P1, P2) receives the other two vertices of each
// cell in slot order.  The visiting order depends only on the mesh, never on scheduling.
template <int D, typename F>
__device__ __forceinline__ void walk_star(const StepParams& p, int v, int c0, const int4& cell0,
                                          int j, int& err, F&& f) {
  f(ld_point<D>(p.x, cell_get(cell0, (j + 1) % 3)), ld_point<D>(p.x, cell_get(cell0, (j + 2) % 3)));
  bool closed = false;
  for (int dir = 0; dir < 2 && !closed; dir++) {
    int cur = c0;
    int kexit = (j + 1 + dir) % 3;
    int hops = 0;
    while (true) {
      const int t = __ldg(p.adj + 4 * (size_t)cur + kexit);
      if (t < 0) break;  // boundary edge: open fan
      const int cn = t >> 2, kn = t & 3;
      if (cn == c0) {
        closed = true;
        break;
      }
      const int4 cl = __ldg(p.cells + cn);
      const int jn = slot_of(cl, v);
      if (jn < 0 || jn == kn || ++hops > MAX_RING) {
        err |= OM_DEV_WALK;
        closed = true;
        break;
      }
      f(ld_point<D>(p.x, cell_get(cl, (jn + 1) % 3)), ld_point<D>(p.x, cell_get(cl, (jn + 2) % 3)));
      cur = cn;
      kexit = 3 - jn - kn;
    }
  }
}

// MODE 0: step, exact inradius in the main walk (most vertices limited: early steps)
// MODE 1: step, lazy limiter (bound first, exact re-walk only where needed)
// MODE 2: write the un-relaxed, un-limited target (get_new_points)
template <int D, int METHOD, int MODE>
__global__ void __launch_bounds__(256, (D == 2 ? 4 : 3)) k_step(StepParams p) {
  constexpr bool TARGET = MODE == 2;
  constexpr bool EXACT = MODE == 0;
  const int v = p.lo + blockIdx.x * blockDim.x + threadIdx.x;
  double diff2 = 0.0;
  int limited = 0;
  int err = 0;
  if (v < p.hi) {
    const Vec<D> P0 = ld_point<D>(p.x, v);
    Vec<D> out = P0;
    const int c0 = p.v2c[v];
    const bool pinned = p.bflag[v] != 0;
    // the Lloyd target of a boundary vertex is its real control-volume centroid; every
    // other consumer pins boundary vertices
    const bool walk = (c0 != OM_NONE_CELL) && (!pinned || (TARGET && METHOD == OM_LLOYD));
    if (walk) {
      Acc<D> acc;
      acc.w = 0.0;
      acc.rmin = INFINITY;
      acc.lb_num = INFINITY;
      acc.lb_den = 1.0;
#pragma unroll
      for (int k = 0; k < D; k++) acc.num.v[k] = 0.0;
#pragma unroll
      for (int k = 0; k < D * (D + 1) / 2; k++) acc.H[k] = 0.0;
      const int4 cell = __ldg(p.cells + c0);
      const int j = slot_of(cell, v);
      if (j < 0) {
        err |= OM_DEV_WALK;
      } else {
        walk_star<D>(p, v, c0, cell, j, err, [&](const Vec<D>& P1, const Vec<D>& P2) {
          accumulate_cell<D, METHOD, EXACT>(P0, P1, P2, acc, err);
        });
        // method formula -> offset of the target from the vertex.  The reference divides
        // by the control volume whatever its sign; only 0/0 (every adjacent cell masked)
        // leaves the vertex where it is.
        Vec<D> d;
        bool ok = acc.w != 0.0;
        if (METHOD == OM_CVT_BLOCK_DIAGONAL) {
          Vec<D> rhs;  // -2 cv (x - c) = 2/3 num
#pragma unroll
          for (int k = 0; k < D; k++) rhs.v[k] = (2.0 / 3.0) * acc.num.v[k];
          ok = ok && solve_sym<D>(acc.H, 2.0 * acc.w, rhs, d);
        } else if (ok) {
          const double inv = 1.0 / (3.0 * acc.w);
#pragma unroll
          for (int k = 0; k < D; k++) d.v[k] = acc.num.v[k] * inv;
        }
        if (ok && !(pinned && !TARGET)) {
          if (!TARGET) {
#pragma unroll
            for (int k = 0; k < D; k++) d.v[k] *= p.omega;
            diff2 = vdot<D>(d, d);
            if (p.limiter) {
              // limited iff |d| > r/2 with r the smallest incident inradius.  LAZY: every
              // inradius satisfies r^2 >= 4 A^2 / (3 sum ee), so 3 |d|^2 sum_ee <= A^2 for
              // the minimising cell proves "not limited" without a sqrt or a division.
              bool check = EXACT || !(3.0 * diff2 * acc.lb_den * (1.0 + 1e-12) <= acc.lb_num);
              if (check) {
                if (!EXACT) {
                  acc.rmin = INFINITY;
                  walk_star<D>(p, v, c0, cell, j, err, [&](const Vec<D>& P1, const Vec<D>& P2) {
                    const CellGeo<D> g = cell_geo<D>(P0, P1, P2);
                    if (g.vol2 > 0.0) acc.rmin = fmin(acc.rmin, inradius<D>(g));
                  });
                }
                const double len = sqrt(diff2);
                const double maxs = 0.5 * acc.rmin;
                if (len > maxs) {
                  const double s = maxs / len;
#pragma unroll
                  for (int k = 0; k < D; k++) d.v[k] *= s;
                  limited = 1;
                }
              }
            }
          }
#pragma unroll
          for (int k = 0; k < D; k++) out.v[k] = P0.v[k] + d.v[k];
        }
      }
    }
    st_point<D>(p.xout, v, out);
  }
  if (!TARGET) reduce_step_stats(diff2, limited, p.ds);
  if (err) atomicOr(&p.ds->err, err);
}

template <int D, int MODE>
int launch_step(om_handle* h, const StepParams& p) {
  const int B = 256;
  const int G = om_grid(p.hi - p.lo, B);
  if (G == 0) return OM_OK;
  switch (h->method) {
    case OM_LLOYD:
      OM_LAUNCH(h, (k_step<D, OM_LLOYD, MODE>), G, B, p);
      break;
    case OM_CVT_BLOCK_DIAGONAL:
      OM_LAUNCH(h, (k_step<D, OM_CVT_BLOCK_DIAGONAL, MODE>), G, B, p);
      break;
    case OM_CPT_FIXED_POINT:
      OM_LAUNCH(h, (k_step<D, OM_CPT_FIXED_POINT, MODE>), G, B, p);
      break;
    case OM_ODT_FIXED_POINT:
      OM_LAUNCH(h, (k_step<D, OM_ODT_FIXED_POINT, MODE>), G, B, p);
      break;
    default:
      om_set_error("method %d has no fixed-point kernel", h->method);
      return OM_ERR_ARG;
  }
  CUDA_TRY(cudaGetLastError());
  return OM_OK;
}

template <int D>
int launch_step_mode(om_handle* h, const StepParams& p, int mode) {
  if (mode == 0) return launch_step<D, 0>(h, p);
  if (mode == 1) return launch_step<D, 1>(h, p);
  return launch_step<D, 2>(h, p);
}

// x <- x + omega (target - x), limited: the driver-loop tail for methods whose target
// comes from a solve (cpt-linear-solve).
template <int D>
__global__ void __launch_bounds__(256) k_relax_from_target(StepParams p, const double* target) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  double diff2 = 0.0;
  int limited = 0, err = 0;
  if (v < p.N) {
    const Vec<D> P0 = ld_point<D>(p.x, v);
    Vec<D> out = P0;
    const int c0 = p.v2c[v];
    if (c0 != OM_NONE_CELL && !p.bflag[v]) {
      const Vec<D> T = ld_point<D>(target, v);
      Vec<D> d;
#pragma unroll
      for (int k = 0; k < D; k++) d.v[k] = p.omega * (T.v[k] - P0.v[k]);
      diff2 = vdot<D>(d, d);
      if (p.limiter) {
        const int4 cell = __ldg(p.cells + c0);
        const int j = slot_of(cell, v);
        double rmin = INFINITY;
        if (j < 0) {
          err |= OM_DEV_WALK;
        } else {
          walk_star<D>(p, v, c0, cell, j, err, [&](const Vec<D>& P1, const Vec<D>& P2) {
            const CellGeo<D> g = cell_geo<D>(P0, P1, P2);
            if (g.vol2 > 0.0)
              rmin = fmin(rmin, inradius<D>(g));
            else
              err |= OM_DEV_DEGENERATE;
          });
        }
        const double len = sqrt(diff2), maxs = 0.5 * rmin;
        if (len > maxs) {
          const double s = maxs / len;
#pragma unroll
          for (int k = 0; k < D; k++) d.v[k] *= s;
          limited = 1;
        }
      }
#pragma unroll
      for (int k = 0; k < D; k++) out.v[k] = P0.v[k] + d.v[k];
    }
    st_point<D>(p.xout, v, out);
  }
  reduce_step_stats(diff2, limited, p.ds);
  if (err) atomicOr(&p.ds->err, err);
}

// ---- implicit surface: sphere f(x) = R^2 - |x - c|^2, grad = -2 (x - c)
// (README.md:157-162 protocol; A.5: sweep all points while max |f| > tol)
__global__ void k_sphere_eval(const double* __restrict__ x, int N, double cx, double cy, double cz,
                              double R2, DevScalars* ds) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  double af = 0.0;
  if (v < N) {
    Vec<3> P = ld_point<3>(x, v);
    double dx = P.v[0] - cx, dy = P.v[1] - cy, dz = P.v[2] - cz;
    af = fabs(R2 - (dx * dx + dy * dy + dz * dz));
  }
  for (int o = 16; o > 0; o >>= 1) af = fmax(af, __shfl_xor_sync(0xffffffffu, af, o));
  if ((threadIdx.x & 31) == 0 && af > 0.0) atomic_max_nonneg(&ds->max_f_bits, af);
}

__global__ void k_sphere_sweep(double* x, int N, double cx, double cy, double cz, double R2,
                               DevScalars* ds) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  double af = 0.0;
  if (v < N) {
    Vec<3> P = ld_point_rw<3>(x, v);
    double dx = P.v[0] - cx, dy = P.v[1] - cy, dz = P.v[2] - cz;
    double r2 = dx * dx + dy * dy + dz * dz;
    double f = R2 - r2;
    // x -= grad f / |grad|^2 with grad = -2 d, |grad|^2 = 4 r2
    double s = f / (4.0 * r2);
    double gx = -2.0 * dx, gy = -2.0 * dy, gz = -2.0 * dz;
    P.v[0] -= gx * s;
    P.v[1] -= gy * s;
    P.v[2] -= gz * s;
    st_point<3>(x, v, P);
    dx = P.v[0] - cx;
    dy = P.v[1] - cy;
    dz = P.v[2] - cz;
    af = fabs(R2 - (dx * dx + dy * dy + dz * dz));
  }
  for (int o = 16; o > 0; o >>= 1) af = fmax(af, __shfl_xor_sync(0xffffffffu, af, o));
  if ((threadIdx.x & 31) == 0 && af > 0.0) atomic_max_nonneg(&ds->max_f_bits, af);
}

__global__ void k_reset_step_scalars(DevScalars* ds) {
  ds->max_diff2_bits = 0ull;
  ds->n_limited = 0ull;
  ds->max_f_bits = 0ull;
}
__global__ void k_reset_f(DevScalars* ds) { ds->max_f_bits = 0ull; }

StepParams make_params(om_handle* h, double* out) {
  StepParams p;
  p.x = h->x;
  p.xout = out;
  p.cells = h->cells;
  p.adj = (const int*)h->adj;
  p.v2c = h->v2c;
  p.bflag = h->bflag;
  p.N = (int)h->N;
  p.lo = 0;
  p.hi = (int)h->N;
  p.omega = h->omega;
  p.limiter = h->limiter;
  p.ds = h->ds;
  return p;
}

double bits_to_double(unsigned long long b) {
  double d;
  memcpy(&d, &b, 8);
  return d;
}

}  // namespace

int om_update_points_impl(om_handle* h, double tol, om_step_stats* out, bool target_only,
                          double* target_out) {
  if (h->N == 0) return OM_OK;
  OM_LAUNCH(h, k_reset_step_scalars, 1, 1, h->ds);
  int32_t iters = 0;
  if (h->method == OM_CPT_LINEAR_SOLVE) {
    double relres = 0.0;
    double* sol = target_only ? target_out : h->xnew;
    OM_TRY(om_pcg_impl(h, h->solver_rtol, h->solver_max_iter, &iters, &relres, sol));
    if (!target_only) {
      // relax + limit from the solved target; result must not alias the target
      double* tmp = nullptr;
      CUDA_TRY(cudaMalloc(&tmp, sizeof(double) * h->N * h->PD));
      StepParams p = make_params(h, tmp);
      const int B = 256, G = om_grid(h->N, B);
      if (h->D == 2)
        OM_LAUNCH(h, k_relax_from_target<2>, G, B, p, sol);
      else
        OM_LAUNCH(h, k_relax_from_target<3>, G, B, p, sol);
      CUDA_TRY(cudaGetLastError());
      CUDA_TRY(cudaMemcpyAsync(h->xnew, tmp, sizeof(double) * h->N * h->PD,
                               cudaMemcpyDeviceToDevice, h->stream));
      CUDA_TRY(cudaStreamSynchronize(h->stream));
      cudaFree(tmp);
    }
  } else {
    StepParams p = make_params(h, target_only ? target_out : h->xnew);
    const bool ranged = !target_only && h->own_hi >= 0;
    if (ranged) {
      p.lo = (int)h->own_lo;
      p.hi = (int)h->own_hi;
    }
    // the lazy limiter pays off once few vertices are limited (the previous step tells)
    const int mode = target_only ? 2 : ((h->limiter && h->limited_frac > 0.25) ? 0 : 1);
    if (h->timing) cudaEventRecord(h->ev[0], h->stream);
    if (h->D == 2)
      OM_TRY(launch_step_mode<2>(h, p, mode));
    else
      OM_TRY(launch_step_mode<3>(h, p, mode));
    if (h->timing) cudaEventRecord(h->ev[1], h->stream);
  }
  OM_TRY(om_fetch_scalars(h));
  if (h->timing && h->method != OM_CPT_LINEAR_SOLVE) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]) == cudaSuccess) {
      h->t_step_ms += ms;
      h->n_step++;
    }
  }
  OM_TRY(om_check_dev_err(h));
  if (!target_only) {
    if (h->own_hi >= 0 && h->method != OM_CPT_LINEAR_SOLVE) {
      // sharded step: only [lo, hi) was written; fold it back, the rest of x stays
      const size_t off = (size_t)h->own_lo * h->PD, cnt = (size_t)(h->own_hi - h->own_lo) * h->PD;
      if (cnt)
        CUDA_TRY(cudaMemcpyAsync(h->x + off, h->xnew + off, sizeof(double) * cnt,
                                 cudaMemcpyDeviceToDevice, h->stream));
      h->limited_frac =
          (double)h->hs->n_limited / (double)std::max<int64_t>(h->own_hi - h->own_lo, 1);
    } else {
      std::swap(h->x, h->xnew);
      h->limited_frac = (double)h->hs->n_limited / (double)h->N;
    }
  }
  if (out) {
    out->max_diff2 = bits_to_double(h->hs->max_diff2_bits);
    out->n_limited = (int64_t)h->hs->n_limited;
    out->is_final = out->max_diff2 < tol * tol ? 1 : 0;
    out->solver_iters = iters;
  }
  return OM_OK;
}

int om_project_impl(om_handle* h, int32_t* sweeps) {
  if (sweeps) *sweeps = 0;
  if (h->surf_kind == 0 || h->N == 0) return OM_OK;
  if (h->surf_kind != 1 || h->D != 3) {
    om_set_error("built-in surface kind %d needs dim 3 (kind 1 = sphere)", h->surf_kind);
    return OM_ERR_ARG;
  }
  const int B = 256, G = om_grid(h->N, B);
  const double cx = h->surf_params[0], cy = h->surf_params[1], cz = h->surf_params[2];
  const double R2 = h->surf_params[3] * h->surf_params[3];
  OM_LAUNCH(h, k_reset_f, 1, 1, h->ds);
  OM_LAUNCH(h, k_sphere_eval, G, B, h->x, (int)h->N, cx, cy, cz, R2, h->ds);
  int n = 0;
  while (true) {
    OM_TRY(om_fetch_scalars(h));
    double maxf = bits_to_double(h->hs->max_f_bits);
    if (!(maxf > h->surf_tol) || n >= h->surf_max_sweeps) break;
    OM_LAUNCH(h, k_reset_f, 1, 1, h->ds);
    OM_LAUNCH(h, k_sphere_sweep, G, B, h->x, (int)h->N, cx, cy, cz, R2, h->ds);
    n++;
  }
  if (sweeps) *sweeps = n;
  return OM_OK;
}
