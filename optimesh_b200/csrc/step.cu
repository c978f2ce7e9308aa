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
#include <cstring>
#include <utility>

#include "common.cuh"
#include "geom.cuh"

namespace {

template <int D>
struct Acc {
  double w;                     // control volume (Lloyd/CVT) or summed cell area (CPT/ODT)
  Vec<D> num;                   // weighted offsets from the vertex
  double H[D * (D + 1) / 2];    // CVT block: sum -0.5 ce_k e_k e_k^T (upper triangle)
  double rmin;                  // smallest incident inradius
};

// Contribution of one incident cell to vertex P0 (P1, P2 follow in slot order).
template <int D, int METHOD>
__device__ __forceinline__ void accumulate_cell(const Vec<D>& P0, const Vec<D>& P1,
                                                const Vec<D>& P2, Acc<D>& a, int& err) {
  CellGeo<D> g = cell_geo<D>(P0, P1, P2);
  if (!(g.vol2 > 0.0)) {
    err |= OM_DEV_DEGENERATE;
    return;
  }
  const double A = sqrt(g.vol2);
  // inradius (A.3): 2A / (l0 + l1 + l2)
  const double per = sqrt(g.ee0) + sqrt(g.ee1) + sqrt(g.ee2);
  a.rmin = fmin(a.rmin, 2.0 * A / per);
  if (METHOD == OM_CPT_FIXED_POINT) {
    // barycenter - P0 = (e2 - e1) / 3
    a.w += A;
#pragma unroll
    for (int k = 0; k < D; k++) a.num.v[k] += A * ((g.e2.v[k] - g.e1.v[k]) / 3.0);
    return;
  }
  // circumcenter - P0 = alpha1 (P1 - P0) + alpha2 (P2 - P0) = alpha1 e2 - alpha2 e1
  const double asum = g.ee0 * g.ed0 + g.ee1 * g.ed1 + g.ee2 * g.ed2;
  const double inva = 1.0 / asum;
  const double al1 = g.ee1 * g.ed1 * inva, al2 = g.ee2 * g.ed2 * inva;
  Vec<D> cc;
#pragma unroll
  for (int k = 0; k < D; k++) cc.v[k] = al1 * g.e2.v[k] - al2 * g.e1.v[k];
  if (METHOD == OM_ODT_FIXED_POINT) {
    a.w += A;
#pragma unroll
    for (int k = 0; k < D; k++) a.num.v[k] += A * cc.v[k];
    return;
  }
  // Lloyd / CVT block-diagonal (A.4, A.9)
  const double inv4A = 0.25 / A;
  const double ce0 = -g.ed0 * inv4A, ce1 = -g.ed1 * inv4A, ce2 = -g.ed2 * inv4A;
  if (ce0 < -0.5 || ce1 < -0.5 || ce2 < -0.5) return;  // cell masked (an angle > 135 deg)
  const double part1 = 0.25 * g.ee1 * ce1, part2 = 0.25 * g.ee2 * ce2;
  a.w += part1 + part2;
  // sub-triangle centroids relative to P0: ((m_k - P0) + (cc - P0)) / 3,
  // m_1 - P0 = -e1/2, m_2 - P0 = e2/2
#pragma unroll
  for (int k = 0; k < D; k++) {
    a.num.v[k] += (part1 * (cc.v[k] - 0.5 * g.e1.v[k]) + part2 * (cc.v[k] + 0.5 * g.e2.v[k])) / 3.0;
  }
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
  out.v[0] = (d * rhs.v[0] - b * rhs.v[1]) / det;
  out.v[1] = (a * rhs.v[1] - b * rhs.v[0]) / det;
  return true;
}
template <>
__device__ __forceinline__ bool solve_sym<3>(const double* H, double diag, const Vec<3>& rhs,
                                             Vec<3>& out) {
  const double a = H[0] + diag, b = H[1], c = H[2], d = H[3] + diag, e = H[4], f = H[5] + diag;
  const double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
  const double det = a * c00 + b * c01 + c * c02;
  if (det == 0.0) return false;
  const double c11 = a * f - c * c, c12 = b * c - a * e, c22 = a * d - b * b;
  out.v[0] = (c00 * rhs.v[0] + c01 * rhs.v[1] + c02 * rhs.v[2]) / det;
  out.v[1] = (c01 * rhs.v[0] + c11 * rhs.v[1] + c12 * rhs.v[2]) / det;
  out.v[2] = (c02 * rhs.v[0] + c12 * rhs.v[1] + c22 * rhs.v[2]) / det;
  return true;
}

struct StepParams {
  const double* x;
  double* xout;
  const int4* cells;
  const int* adj;  // flat view of int4 adjacency: adj[4*c + k]
  const int* v2c;
  const uint8_t* bflag;
  int N;
  double omega;
  int limiter;
  DevScalars* ds;
};

constexpr int MAX_RING = 4096;

// TARGET: write the un-relaxed, un-limited target (get_new_points) instead of the step.
template <int D, int METHOD, bool TARGET>
__global__ void __launch_bounds__(256) k_step(StepParams p) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  double diff2 = 0.0;
  int limited = 0;
  int err = 0;
  if (v < p.N) {
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
#pragma unroll
      for (int k = 0; k < D; k++) acc.num.v[k] = 0.0;
#pragma unroll
      for (int k = 0; k < D * (D + 1) / 2; k++) acc.H[k] = 0.0;

      int4 cell = __ldg(p.cells + c0);
      int j = slot_of(cell, v);
      if (j < 0) {
        err |= OM_DEV_WALK;
      } else {
        // direction A leaves the start cell through local edge (j+1)%3, direction B
        // (open fans only) through (j+2)%3.
        int cur = c0;
        int kexit = (j + 1) % 3;
        bool closed = false;
        {
          Vec<D> P1 = ld_point<D>(p.x, cell_get(cell, (j + 1) % 3));
          Vec<D> P2 = ld_point<D>(p.x, cell_get(cell, (j + 2) % 3));
          accumulate_cell<D, METHOD>(P0, P1, P2, acc, err);
        }
        for (int dir = 0; dir < 2 && !closed; dir++) {
          if (dir == 1) {
            cur = c0;
            kexit = (j + 2) % 3;
          }
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
            Vec<D> P1 = ld_point<D>(p.x, cell_get(cl, (jn + 1) % 3));
            Vec<D> P2 = ld_point<D>(p.x, cell_get(cl, (jn + 2) % 3));
            accumulate_cell<D, METHOD>(P0, P1, P2, acc, err);
            cur = cn;
            kexit = 3 - jn - kn;
          }
        }
        // method formula -> offset of the target from the vertex
        Vec<D> d;
        // the reference divides by the control volume whatever its sign; only 0/0 (every
        // adjacent cell masked) leaves the vertex where it is
        bool ok = acc.w != 0.0;
        if (METHOD == OM_CVT_BLOCK_DIAGONAL) {
          Vec<D> rhs;
#pragma unroll
          for (int k = 0; k < D; k++) rhs.v[k] = 2.0 * acc.num.v[k];
          ok = ok && solve_sym<D>(acc.H, 2.0 * acc.w, rhs, d);
        } else if (ok) {
          const double inv = 1.0 / acc.w;
#pragma unroll
          for (int k = 0; k < D; k++) d.v[k] = acc.num.v[k] * inv;
        }
        if (ok && !(pinned && !TARGET)) {
          if (TARGET) {
#pragma unroll
            for (int k = 0; k < D; k++) out.v[k] = P0.v[k] + d.v[k];
          } else {
#pragma unroll
            for (int k = 0; k < D; k++) d.v[k] *= p.omega;
            diff2 = vdot<D>(d, d);
            if (p.limiter) {
              const double len = sqrt(diff2);
              const double maxs = 0.5 * acc.rmin;
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
        }
      }
    }
    st_point<D>(p.xout, v, out);
  }
  if (!TARGET) {
    // warp-level reduction, one atomic per warp (max / integer add: order independent)
    for (int o = 16; o > 0; o >>= 1) {
      diff2 = fmax(diff2, __shfl_xor_sync(0xffffffffu, diff2, o));
      limited += __shfl_xor_sync(0xffffffffu, limited, o);
    }
    if ((threadIdx.x & 31) == 0) {
      if (diff2 > 0.0) atomic_max_nonneg(&p.ds->max_diff2_bits, diff2);
      if (limited) atomicAdd(&p.ds->n_limited, (unsigned long long)limited);
    }
  }
  if (err) atomicOr(&p.ds->err, err);
}

template <int D, bool TARGET>
int launch_step(om_handle* h, const StepParams& p) {
  const int B = 256;
  const int G = om_grid(h->N, B);
  switch (h->method) {
    case OM_LLOYD:
      OM_LAUNCH(h, (k_step<D, OM_LLOYD, TARGET>), G, B, p);
      break;
    case OM_CVT_BLOCK_DIAGONAL:
      OM_LAUNCH(h, (k_step<D, OM_CVT_BLOCK_DIAGONAL, TARGET>), G, B, p);
      break;
    case OM_CPT_FIXED_POINT:
      OM_LAUNCH(h, (k_step<D, OM_CPT_FIXED_POINT, TARGET>), G, B, p);
      break;
    case OM_ODT_FIXED_POINT:
      OM_LAUNCH(h, (k_step<D, OM_ODT_FIXED_POINT, TARGET>), G, B, p);
      break;
    default:
      om_set_error("method %d has no fixed-point kernel", h->method);
      return OM_ERR_ARG;
  }
  CUDA_TRY(cudaGetLastError());
  return OM_OK;
}

// x <- x + omega (xsol - x), limited: the driver-loop tail for methods whose target comes
// from a solve (cpt-linear-solve).  The limiter needs the smallest incident inradius, so
// it reuses the star walk with the CPT accumulator.
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
      Vec<D> T = ld_point<D>(target, v);
      Vec<D> d;
#pragma unroll
      for (int k = 0; k < D; k++) d.v[k] = p.omega * (T.v[k] - P0.v[k]);
      diff2 = vdot<D>(d, d);
      if (p.limiter) {
        Acc<D> acc;
        acc.w = 0.0;
        acc.rmin = INFINITY;
#pragma unroll
        for (int k = 0; k < D; k++) acc.num.v[k] = 0.0;
        int4 cell = __ldg(p.cells + c0);
        int j = slot_of(cell, v);
        int cur = c0, kexit = (j + 1) % 3, hops = 0;
        if (j < 0) err |= OM_DEV_WALK;
        while (j >= 0) {
          const int4 cl = __ldg(p.cells + cur);
          const int jn = slot_of(cl, v);
          Vec<D> P1 = ld_point<D>(p.x, cell_get(cl, (jn + 1) % 3));
          Vec<D> P2 = ld_point<D>(p.x, cell_get(cl, (jn + 2) % 3));
          accumulate_cell<D, OM_CPT_FIXED_POINT>(P0, P1, P2, acc, err);
          const int t = __ldg(p.adj + 4 * (size_t)cur + kexit);
          if (t < 0 || (t >> 2) == c0 || ++hops > MAX_RING) break;  // interior: closed ring
          const int cn = t >> 2, kn = t & 3;
          const int4 cl2 = __ldg(p.cells + cn);
          const int j2 = slot_of(cl2, v);
          if (j2 < 0 || j2 == kn) {
            err |= OM_DEV_WALK;
            break;
          }
          cur = cn;
          kexit = 3 - j2 - kn;
        }
        const double len = sqrt(diff2), maxs = 0.5 * acc.rmin;
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
  for (int o = 16; o > 0; o >>= 1) {
    diff2 = fmax(diff2, __shfl_xor_sync(0xffffffffu, diff2, o));
    limited += __shfl_xor_sync(0xffffffffu, limited, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (diff2 > 0.0) atomic_max_nonneg(&p.ds->max_diff2_bits, diff2);
    if (limited) atomicAdd(&p.ds->n_limited, (unsigned long long)limited);
  }
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
    // x -= grad f / |grad|^2 with grad = -2 d  ->  x += d f / (2 r2)
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
    if (h->timing) cudaEventRecord(h->ev[0], h->stream);
    if (h->D == 2) {
      if (target_only)
        OM_TRY((launch_step<2, true>(h, p)));
      else
        OM_TRY((launch_step<2, false>(h, p)));
    } else {
      if (target_only)
        OM_TRY((launch_step<3, true>(h, p)));
      else
        OM_TRY((launch_step<3, false>(h, p)));
    }
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
  if (!target_only) std::swap(h->x, h->xnew);
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
