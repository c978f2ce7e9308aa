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
// Instruction budget (ncu: the kernel is bound by instruction issue -- 70 % of the issue slots
// busy, fp64 pipe 47 % -- not by HBM): one rsqrt per cell
// visit and no division -- 1/(4A) = rsqrt(16 A^2), the circumcentre weights use
// sum_k ee_k ed_k = -8 A^2, thirds are applied once per vertex.  The inradius (3 sqrt +
// 1 div per cell) is only needed where the limiter bites, so the LAZY variant first tests
// the division-free bound r_in^2 >= 4 A^2 / (3 sum ee) and re-walks the star exactly only
// for vertices that fail it; both variants produce identical bits.
#include <cstring>
#include <utility>

#include "common.cuh"
#include "geom.cuh"

// tuning knobs of the ring-row step kernel
#ifndef OM_K1_UNROLL
#define OM_K1_UNROLL 2
#endif
#ifndef OM_K1_MINB
#define OM_K1_MINB 8
#endif

namespace {

constexpr int K1_UNROLL = OM_K1_UNROLL;

// Per-vertex accumulators, kept in SCALED units so that constant factors are applied once
// per vertex instead of once per cell visit (q = 1/(4A), t_k = ed_k q = -ce_k,
// w_k = ee_k t_k = -4 part_k):
//   Lloyd/CVT:  w   = sum (w_1 + w_2)               = -4 cv
//               num = sum s2' e2 - s1' e1           = -12 cv (centroid - x)
//               H   = sum t_1 e1 e1^T + t_2 e2 e2^T = 2 x (CVT Hessian block without 2 cv I)
//   CPT/ODT:    w = sum A, num = 3 sum A (r_c - x)
template <int D>
struct Acc {
  double w;
  Vec<D> num;
  double H[D * (D + 1) / 2];  // upper triangle
  double rmin;                // EXACT: smallest incident inradius
  double lb_num, lb_den;      // LAZY: cell minimising A^2 / (sum(ee)/2) (kept as a fraction)
};

// 1/sqrt(x) for positive, normal x (the caller has checked x > 0): hardware seed
// (rsqrt.approx.f64, relative error < 2^-22) + ONE third-order step
//   e = 1 - x y^2,  y <- y (1 + e/2 + 3 e^2/8)        (remainder 5 e^3/16 < 2^-67)
// -- five fp64 instructions instead of the seven of two Newton steps (the kernel is bound by
// instruction issue: every instruction saved counts, whatever its pipe).
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x * y, y, 1.0);
  const double p = fma(0.375, e, 0.5);
  return fma(y * e, p, y);
}

// t > 0.5 on the integer pipe (exact: compares the bit pattern with that of 0.5)
__device__ __forceinline__ bool gt_half(double t) {
  // as one 64-bit integer compare: t > 0.5  <=>  bits(t) > bits(0.5) for every non-NaN t
  // (negative t have the sign bit set, i.e. are negative as signed integers)
  return __double_as_longlong(t) > 0x3fe0000000000000ll;
}

// a < b for non-negative doubles (or +inf) on the integer pipe: their order is the order of
// their bit patterns
__device__ __forceinline__ bool lt_nonneg(double a, double b) {
  return __double_as_longlong(a) < __double_as_longlong(b);
}

// exact inradius of one cell (A.3): 2A / (l0 + l1 + l2)
template <int D>
__device__ __forceinline__ double inradius(const CellGeo<D>& g) {
  const double A = sqrt(g.vol2);
  return 2.0 * A / (sqrt(g.ee0) + sqrt(g.ee1) + sqrt(g.ee2));
}

// Contribution of one incident cell to vertex P0 (P1, P2 follow in slot order).
// Only the two edges at the vertex are formed (e1 = P0 - P2, e2 = P1 - P0); since
// e0 + e1 + e2 = 0 every other product follows from ee1, ee2 and ed0 = e1.e2:
//   ed1 = -(ed0 + ee2), ed2 = -(ed0 + ee1), ee0 = ee1 + ee2 + 2 ed0, A^2 = (ee1 ee2 - ed0^2)/4.
// bary: ODT methods only -- the cell has a boundary edge and contributes its barycenter
// instead of its circumcenter (which may lie outside the domain there).
template <int D, int METHOD, bool EXACT>
__device__ __forceinline__ void accumulate_cell(const Vec<D>& P0, const Vec<D>& P1,
                                                const Vec<D>& P2, Acc<D>& a, int& err,
                                                bool bary) {
  const Vec<D> e1 = vsub<D>(P0, P2), e2 = vsub<D>(P1, P0);
  const double ee1 = vdot<D>(e1, e1), ee2 = vdot<D>(e2, e2), ed0 = vdot<D>(e1, e2);
  const double vol2 = 0.25 * fma(ee1, ee2, -ed0 * ed0);
  if (!(vol2 > 0.0)) {
    err |= OM_DEV_DEGENERATE;
    return;
  }
  const double ed1 = -(ed0 + ee2), ed2 = -(ed0 + ee1);
  if (EXACT) {
    const double A = sqrt(vol2);
    a.rmin = fmin(a.rmin, 2.0 * A / (sqrt(-(ed1 + ed2)) + sqrt(ee1) + sqrt(ee2)));  // ee0
  } else {
    const double Sh = ee2 - ed2;  // (ee0 + ee1 + ee2) / 2 = ee1 + ee2 + ed0
    if (lt_nonneg(vol2 * a.lb_den, a.lb_num * Sh)) {
      a.lb_num = vol2;
      a.lb_den = Sh;
    }
  }
  const double r = fast_rsqrt(vol2);  // 1/A
  if (METHOD == OM_CPT_FIXED_POINT) {
    // 3 (barycenter - P0) = e2 - e1
    const double A = vol2 * r;
    a.w += A;
#pragma unroll
    for (int k = 0; k < D; k++) a.num.v[k] += A * (e2.v[k] - e1.v[k]);
    return;
  }
  if (METHOD == OM_ODT_FIXED_POINT || METHOD == OM_ODT_DP_FP) {
    // volume averaged: weight A = vol2 r; count averaged (density preserving): weight 1
    // 3 (circumcenter - P0) = 3 (al1 e2 - al2 e1), al_k = ee_k ed_k / (-8 A^2)
    const bool dp = METHOD == OM_ODT_DP_FP;
    const double A = vol2 * r;
    a.w += dp ? 1.0 : A;
    const double f = -0.375 * r * (dp ? r : 1.0);
    const double wb = dp ? 1.0 : A;  // barycenter: 3 (b - P0) = e2 - e1
    const double s2 = bary ? wb : ee1 * ed1 * f, s1 = bary ? wb : ee2 * ed2 * f;
#pragma unroll
    for (int k = 0; k < D; k++) a.num.v[k] = fma(s2, e2.v[k], fma(-s1, e1.v[k], a.num.v[k]));
    return;
  }
  // Lloyd / CVT block-diagonal (A.4, A.9), scaled as described at Acc
  const double q = 0.25 * r;
  const double t0 = ed0 * q, t1 = ed1 * q, t2 = ed2 * q;  // -ce_k
  if (gt_half(t0) | gt_half(t1) | gt_half(t2)) return;  // cell masked (an angle > 135 deg)
  const double w1 = ee1 * t1, w2 = ee2 * t2;  // -4 part_k
  const double ws = w1 + w2;
  a.w += ws;
  // al_k = ee_k ed_k (-1/(8A^2)) = -2 q w_k;  -12 x [part1 (cc - e1/2) + part2 (cc + e2/2)]
  //   = e2 (u w1 + w2/2 ... ) with u = -2 q ws:  s2' = u w1 - 0.5 w2 ... (signs folded below)
  const double u = -2.0 * q * ws;
  const double s2 = fma(u, w1, 0.5 * w2), s1 = fma(u, w2, 0.5 * w1);
  // accumulated by fma chains: two fp64 instructions per component instead of three
#pragma unroll
  for (int k = 0; k < D; k++) a.num.v[k] = fma(s2, e2.v[k], fma(-s1, e1.v[k], a.num.v[k]));
  if (METHOD == OM_CVT_BLOCK_DIAGONAL) {
    Vec<D> a1, a2;
#pragma unroll
    for (int k = 0; k < D; k++) {
      a1.v[k] = t1 * e1.v[k];
      a2.v[k] = t2 * e2.v[k];
    }
    int qi = 0;
#pragma unroll
    for (int i = 0; i < D; i++)
#pragma unroll
      for (int j = i; j < D; j++) {
        a.H[qi] = fma(a1.v[i], e1.v[j], fma(a2.v[i], e2.v[j], a.H[qi]));
        qi++;
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
  const int* ring;  // N x OM_RING_W ring rows, or nullptr
  int* over;        // vertices the main launch leaves to the list-driven (SRC 2) launch
  // partitioned coordinates: foreign vertices are only current if pinned or stamped by the
  // last band exchange (nullptr: everything is current)
  const int* valid_epoch;
  int valid_stamp;
  __device__ __forceinline__ bool valid(int u) const {
    return valid_epoch == nullptr || (u >= lo && u < hi) || bflag[u] != 0 ||
           valid_epoch[u] == valid_stamp;
  }
  int N;
  int lo, hi;  // vertices [lo, hi) are processed
  double omega;
  int limiter;
  int odt_bary;  // ODT: cells with a boundary edge contribute their barycenter
  DevScalars* ds;
};

// Block-level reduction of the step statistics (up to 256 threads): one conditional atomic per
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
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) {
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
// BC: f also receives "this cell has a boundary edge" (read from the cell's twin row).
template <int D, bool BC = false, typename F>
__device__ __forceinline__ void walk_star(const StepParams& p, int v, int c0, const int4& cell0,
                                          int j, int& err, F&& f) {
  auto has_boundary_edge = [&](int c) {
    if (!BC) return false;
    const int4 ta = __ldg(reinterpret_cast<const int4*>(p.adj) + c);
    return ta.x < 0 || ta.y < 0 || ta.z < 0;
  };
  if (!(p.valid(cell_get(cell0, (j + 1) % 3)) && p.valid(cell_get(cell0, (j + 2) % 3))))
    p.ds->stale = 1;
  f(ld_point<D>(p.x, cell_get(cell0, (j + 1) % 3)), ld_point<D>(p.x, cell_get(cell0, (j + 2) % 3)),
    has_boundary_edge(c0));
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
      if (!(p.valid(cell_get(cl, (jn + 1) % 3)) && p.valid(cell_get(cl, (jn + 2) % 3))))
        p.ds->stale = 1;
      f(ld_point<D>(p.x, cell_get(cl, (jn + 1) % 3)), ld_point<D>(p.x, cell_get(cl, (jn + 2) % 3)),
        has_boundary_edge(cn));
      cur = cn;
      kexit = 3 - jn - kn;
    }
  }
}

// ---- ring rows: the one-ring of a vertex as (up to) OM_RING_W neighbour vertex ids
// Entry q holds n_q | (f_q << 30); cell q of the star is (v, n_q, n_{q+1 mod k}) and f_q says
// whether its slot order is (v, n_{q+1}, n_q) instead of (v, n_q, n_{q+1}).  Unused entries are
// -1; entry 0 == -2 marks a vertex the kernel must walk instead (pinned/boundary vertex, open
// fan, or more than OM_RING_W cells).  The order is the walk order, so the sums are
// bit-identical to the walk; the rows only remove the dependent adj -> cell -> point chains.
// threads per block of the step kernels (3D: smaller, its ring staging is twice as wide)
// 128-thread blocks, 8 per SM: the same 1024 resident threads as 256 x 4, but a block retires as
// soon as its four warps are done (measured: 0.461 ms against 0.470 ms; 64 x 16: 0.464 ms)
#ifndef OM_K1_BLOCK
#define OM_K1_BLOCK 128
#endif
template <int D>
__host__ __device__ constexpr int step_block() {
  return D == 2 ? OM_K1_BLOCK : 128;
}
constexpr int RING_FLAG = 1 << 30;   // cell q has slot order (v, n_{q+1}, n_q)
constexpr int RING_BCELL = 1 << 29;  // cell q has a boundary edge (ODT uses its barycenter)
constexpr int RING_MASK = RING_BCELL - 1;

template <bool LIST>
__global__ void __launch_bounds__(256)
    k_build_rings(const int4* __restrict__ cells, const int* __restrict__ adj,
                  const int* __restrict__ v2c, const uint8_t* __restrict__ bflag, int n,
                  const int* __restrict__ list, int* __restrict__ ring, int lo, int hi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int v = LIST ? list[i] : i;
  if (v < lo || v >= hi) return;  // partitioned run: only the own range's rows are read
  int e[OM_RING_W];
#pragma unroll
  for (int q = 0; q < OM_RING_W; q++) e[q] = -1;
  bool ok = false;
  const int c0 = v2c[v];
  if (c0 != OM_NONE_CELL && !bflag[v]) {
    const int4 cell0 = __ldg(cells + c0);
    const int j = slot_of(cell0, v);
    if (j >= 0) {
      int k = 0;  // cells closed so far
      int last = cell_get(cell0, (j + 2) % 3);
      e[0] = cell_get(cell0, (j + 1) % 3);  // cell 0 = (v, n0, n1), slot order kept
      int cur = c0, kexit = (j + 1) % 3;
      ok = true;
      while (true) {
        // twin row of cell k (`cur`): the exit edge, and whether the cell has a boundary edge
        const int4 ta = __ldg(reinterpret_cast<const int4*>(adj) + cur);
        if (ta.x < 0 || ta.y < 0 || ta.z < 0) {
#pragma unroll
          for (int q = 0; q < OM_RING_W; q++)
            if (q == k) e[q] |= RING_BCELL;
        }
        k++;
        const int t = cell_get(ta, kexit);
        if (t < 0) {  // open fan
          ok = false;
          break;
        }
        const int cn = t >> 2, kn = t & 3;
        if (cn == c0) break;  // closed: `last` is n0 again
        if (k >= OM_RING_W) {
          ok = false;
          break;
        }
        const int4 cl = __ldg(cells + cn);
        const int jn = slot_of(cl, v);
        if (jn < 0 || jn == kn) {
          ok = false;
          break;
        }
        const int p1 = cell_get(cl, (jn + 1) % 3), p2 = cell_get(cl, (jn + 2) % 3);
        // cell k = (v, last, new): slot order (v, last, new) iff p1 == last
        const int flag = (p1 == last) ? 0 : RING_FLAG;
        const int nxt = (p1 == last) ? p2 : p1;
#pragma unroll
        for (int q = 1; q < OM_RING_W; q++)
          if (q == k) e[q] = last | flag;
        last = nxt;
        cur = cn;
        kexit = 3 - jn - kn;
      }
    }
  }
  if (!ok) e[0] = -2;
  int4* out = reinterpret_cast<int4*>(ring + (size_t)OM_RING_W * v);
  out[0] = make_int4(e[0], e[1], e[2], e[3]);
  out[1] = make_int4(e[4], e[5], e[6], e[7]);
}

// MODE 0: step, exact inradius in the main pass (most vertices limited: early steps)
// MODE 1: step, lazy limiter (bound first, exact second pass only where needed)
// MODE 2: write the un-relaxed, un-limited target (get_new_points)
// SRC 0: vertices [lo, hi), one-ring from the ring rows; vertices without a row that must
//        move are appended to p.over and left to a SRC 2 launch
// SRC 1: vertices [lo, hi), star walk
// SRC 2: vertices p.list[0 .. n_list), star walk
template <int D, int METHOD, int MODE, int SRC>
__global__ void __launch_bounds__(step_block<D>(), (D == 2 ? OM_K1_MINB : 3))
    k_step(StepParams p) {
  constexpr bool TARGET = MODE == 2;
  constexpr bool EXACT = MODE == 0;
  // SRC 2: the list and its length live on the device (written by the SRC 0/1 launch that
  // precedes this one on the stream): block-stride loop, uniform trip count per block
  const int n_list = SRC == 2 ? p.ds->n_over : 0;
  for (int base = blockIdx.x * blockDim.x; SRC != 2 ? base == (int)(blockIdx.x * blockDim.x)
                                                      : base < n_list;
       base += gridDim.x * blockDim.x) {
  const int i = base + threadIdx.x;
  const bool active = SRC == 2 ? (i < n_list) : (p.lo + i < p.hi);
  double diff2 = 0.0;
  int limited = 0;
  int err = 0;
  bool deferred = false;  // left to the list-driven exact launch
  int vdef = 0;
  if (active) {
    const int v = SRC == 2 ? p.over[i] : p.lo + i;
    vdef = v;
    const Vec<D> P0 = ld_point<D>(p.x, v);
    Vec<D> out = P0;
    const int c0 = p.v2c[v];
    const bool pinned = p.bflag[v] != 0;
    // the Lloyd target of a boundary vertex is its real control-volume centroid; every
    // other consumer pins boundary vertices
    bool walk = (c0 != OM_NONE_CELL) && (!pinned || (TARGET && METHOD == OM_LLOYD));

    // SRC 0: the ring vertices are staged in shared memory by cp.async (one 16-byte copy
    // per vertex in 2D, two in 3D; all in flight at once, no registers held), slot
    // [q][thread] so that a warp's accesses are conflict free.  Every thread only reads the
    // slots it filled itself: no block barrier is needed.
    __shared__ double2 ring_sm[SRC == 0 ? OM_RING_W * (D == 2 ? 1 : 2) * step_block<D>() : 1];
    int nring = 0;         // cells (= ring vertices) in the row
    unsigned rflags = 0u;  // bit q: cell q has slot order (v, n_{q+1}, n_q)
    unsigned bcells = 0u;  // bit q: cell q has a boundary edge (ODT methods only)
    constexpr bool ODT = METHOD == OM_ODT_FIXED_POINT || METHOD == OM_ODT_DP_FP;
    int4 cell = make_int4(0, 0, 0, 0);
    int j = 0;
    if (walk) {
      if (SRC == 0) {
        const int4* rp = reinterpret_cast<const int4*>(p.ring + (size_t)OM_RING_W * v);
        const int4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
        const int e[OM_RING_W] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
        if (e[0] == -2) {
          // no row (more than OM_RING_W cells, or an open fan): the walk kernel does it
          deferred = true;
          walk = false;
        } else {
          constexpr int PER = (D == 2) ? 1 : 2;  // 16-byte pieces per vertex
#pragma unroll
          for (int q = 0; q < OM_RING_W; q++)
            if (e[q] >= 0) {
              if (!p.valid(e[q] & RING_MASK)) p.ds->stale = 1;  // caller refreshes and repeats
              const double2* src =
                  reinterpret_cast<const double2*>(p.x) + (size_t)PER * (e[q] & RING_MASK);
#pragma unroll
              for (int h2 = 0; h2 < PER; h2++) {
                const unsigned dst = (unsigned)__cvta_generic_to_shared(
                    &ring_sm[(q * PER + h2) * step_block<D>() + threadIdx.x]);
                #ifdef OM_K1_CPASYNC_CG
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst),
                             "l"(src + h2));
#else
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst),
                             "l"(src + h2));
#endif
              }
              nring = q + 1;
              rflags |= (e[q] & RING_FLAG) ? (1u << q) : 0u;
              if (ODT) bcells |= (e[q] & RING_BCELL) ? (1u << q) : 0u;
            }
          asm volatile("cp.async.commit_group;");
          asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
      } else {
        cell = __ldg(p.cells + c0);
        j = slot_of(cell, v);
        if (j < 0) {
          err |= OM_DEV_WALK;
          walk = false;
        }
      }
    }
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
      // visits the cells of the star in walk order; f(P1, P2) gets the other two vertices
      // of each cell in slot order
      auto for_each_cell = [&](auto&& f) {
        if (SRC == 0) {
          constexpr int PER = (D == 2) ? 1 : 2;
          auto ld_ring = [&](int q) {
            Vec<D> r;
            const double2 a = ring_sm[(q * PER) * step_block<D>() + threadIdx.x];
            r.v[0] = a.x;
            r.v[1] = a.y;
            if (D == 3) r.v[D - 1] = ring_sm[(q * PER + PER - 1) * step_block<D>() + threadIdx.x].x;
            return r;
          };
          Vec<D> A = ld_ring(0);
#pragma unroll K1_UNROLL
          for (int q = 0; q < nring; q++) {
            const Vec<D> B = ld_ring(q + 1 < nring ? q + 1 : 0);
            // select the operands (no divergent branch around the cell arithmetic)
            const bool sw = (rflags >> q) & 1u;
            Vec<D> X1, X2;
#pragma unroll
            for (int k = 0; k < D; k++) {
              X1.v[k] = sw ? B.v[k] : A.v[k];
              X2.v[k] = sw ? A.v[k] : B.v[k];
            }
            f(X1, X2, ODT && ((bcells >> q) & 1u));
            A = B;
          }
        } else {
          walk_star<D, ODT>(p, v, c0, cell, j, err, f);
        }
      };

      const bool odt_bary = ODT && p.odt_bary != 0;
      for_each_cell([&](const Vec<D>& P1, const Vec<D>& P2, bool bcell) {
        accumulate_cell<D, METHOD, EXACT>(P0, P1, P2, acc, err, odt_bary && bcell);
      });
      // method formula -> offset of the target from the vertex.  The reference divides by
      // the control volume whatever its sign; only 0/0 (every adjacent cell masked) leaves
      // the vertex where it is.
      Vec<D> d;
      bool ok = acc.w != 0.0;
      if (METHOD == OM_CVT_BLOCK_DIAGONAL) {
        // (2 cv I + Hess) d = -2 cv (x - c) in the scaled accumulators: (w I - H) d = num / 3
        Vec<D> rhs;
        double Hn[D * (D + 1) / 2];
#pragma unroll
        for (int k = 0; k < D; k++) rhs.v[k] = acc.num.v[k] * (1.0 / 3.0);
#pragma unroll
        for (int k = 0; k < D * (D + 1) / 2; k++) Hn[k] = -acc.H[k];
        ok = ok && solve_sym<D>(Hn, acc.w, rhs, d);
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
            // inradius satisfies r^2 >= 4 A^2 / (3 sum ee), so 3 |d|^2 sum_ee <= A^2 (lb_den holds sum_ee / 2) for
            // the minimising cell proves "not limited" without a sqrt or a division.
            const bool check =
                EXACT || !(6.0 * diff2 * acc.lb_den * (1.0 + 1e-12) <= acc.lb_num);
            if (check && !EXACT) {
              // rare (a few % of the vertices) and expensive: running it here would keep
              // whole warps busy for one lane.  The vertex goes to the list-driven exact
              // launch instead, where all lanes have work.
              deferred = true;
              diff2 = 0.0;
            } else if (check) {
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
    if (!deferred) st_point<D>(p.xout, v, out);
  }
  if (!TARGET) reduce_step_stats(diff2, limited, p.ds);
  if (!TARGET && SRC != 2) {
    const int vals[1] = {vdef};
    const bool preds[1] = {deferred};
    block_append<1>(&p.ds->n_over, p.over, vals, preds);
  }
  if (err) atomicOr(&p.ds->err, err);
  }  // block-stride loop (one trip unless SRC == 2)
}

template <int D, int MODE, int SRC>
int launch_step(om_handle* h, const StepParams& p) {
  const int B = step_block<D>();
  const int G = SRC == 2 ? 148 * 4 : om_grid(p.hi - p.lo, B);
  if (G == 0) return OM_OK;
  switch (h->method) {
    case OM_LLOYD:
      OM_LAUNCH(h, (k_step<D, OM_LLOYD, MODE, SRC>), G, B, p);
      break;
    case OM_CVT_BLOCK_DIAGONAL:
      OM_LAUNCH(h, (k_step<D, OM_CVT_BLOCK_DIAGONAL, MODE, SRC>), G, B, p);
      break;
    case OM_CPT_FIXED_POINT:
      OM_LAUNCH(h, (k_step<D, OM_CPT_FIXED_POINT, MODE, SRC>), G, B, p);
      break;
    case OM_ODT_FIXED_POINT:
      OM_LAUNCH(h, (k_step<D, OM_ODT_FIXED_POINT, MODE, SRC>), G, B, p);
      break;
    case OM_ODT_DP_FP:
      OM_LAUNCH(h, (k_step<D, OM_ODT_DP_FP, MODE, SRC>), G, B, p);
      break;
    default:
      om_set_error("method %d has no fixed-point kernel", h->method);
      return OM_ERR_ARG;
  }
  CUDA_TRY(cudaGetLastError());
  return OM_OK;
}

// src: 0 ring rows over [lo,hi), 1 walk over [lo,hi), 2 walk over the overflow list
template <int D>
int launch_step_mode(om_handle* h, const StepParams& p, int mode, int src) {
  if (mode == 2) return launch_step<D, 2, 1>(h, p);
  if (mode == 0) {
    if (src == 0) return launch_step<D, 0, 0>(h, p);
    if (src == 1) return launch_step<D, 0, 1>(h, p);
    return launch_step<D, 0, 2>(h, p);
  }
  if (src == 0) return launch_step<D, 1, 0>(h, p);
  if (src == 1) return launch_step<D, 1, 1>(h, p);
  return launch_step<D, 0, 2>(h, p);  // the deferred list is always done exactly
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
          walk_star<D>(p, v, c0, cell, j, err, [&](const Vec<D>& P1, const Vec<D>& P2, bool) {
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
  ds->n_over = 0;
  ds->stale = 0;
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
  p.ring = h->ring;
  p.over = h->over;
  p.valid_epoch = (h->own_hi >= 0 && h->valid_epoch && !h->all_valid) ? h->valid_epoch : nullptr;
  p.valid_stamp = h->valid_stamp;
  p.N = (int)h->N;
  p.lo = 0;
  p.hi = (int)h->N;
  p.omega = h->omega;
  p.limiter = h->limiter;
  p.odt_bary = h->odt_bary;
  p.ds = h->ds;
  return p;
}

double bits_to_double(unsigned long long b) {
  double d;
  memcpy(&d, &b, 8);
  return d;
}

}  // namespace

// Fills the step statistics from the scalars of the last readback.
void om_step_stats_from_scalars(om_handle* h, double tol, om_step_stats* out) {
  if (h->timing && !om_is_solve_method(h->method) && h->ev_pending) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]) == cudaSuccess) {
      h->t_step_ms += ms;
      h->n_step++;
    }
    h->ev_pending = false;
  }
  const int64_t nown = h->own_hi >= 0 ? std::max<int64_t>(h->own_hi - h->own_lo, 1)
                                      : std::max<int64_t>(h->N, 1);
  h->limited_frac = (double)h->hs->n_limited / (double)nown;
  if (out) {
    out->max_diff2 = bits_to_double(h->hs->max_diff2_bits);
    out->n_limited = (int64_t)h->hs->n_limited;
    out->is_final = out->max_diff2 < tol * tol ? 1 : 0;
  }
}

int om_update_points_impl(om_handle* h, double tol, om_step_stats* out, bool target_only,
                          double* target_out, bool defer_fetch) {
  if (h->N == 0) return OM_OK;
  OM_LAUNCH(h, k_reset_step_scalars, 1, 1, h->ds);
  int32_t iters = 0;
  if (om_is_solve_method(h->method)) {
    double relres = 0.0;
    double* sol = target_only ? target_out : h->xnew;
    if (h->method == OM_CPT_LINEAR_SOLVE)
      OM_TRY(om_pcg_impl(h, h->solver_rtol, h->solver_max_iter, &iters, &relres, sol));
    else
      OM_TRY(om_quasi_newton_impl(h, h->solver_rtol, h->solver_max_iter, &iters, &relres, sol));
    if (!target_only) {
      // relax + limit from the solved target; result must not alias the target
      double* tmp = nullptr;
      CUDA_TRY(om_malloc(h, &tmp, sizeof(double) * h->N * h->PD));
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
      om_free(h, tmp);
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
    const int src = (mode != 2 && h->ring && h->use_rings) ? 0 : 1;
    if (h->timing) cudaEventRecord(h->ev[0], h->stream);
    if (h->D == 2)
      OM_TRY(launch_step_mode<2>(h, p, mode, src));
    else
      OM_TRY(launch_step_mode<3>(h, p, mode, src));
    if (h->timing) {
      cudaEventRecord(h->ev[1], h->stream);
      h->ev_pending = true;
    }
    if (mode != 2) {
      // vertices the main launch deferred (no ring row, or the lazy limiter bound failed):
      // second, list-driven launch with the exact limiter; it reads the list length on the
      // device, so no readback is needed in between
      if (h->D == 2)
        OM_TRY(launch_step_mode<2>(h, p, mode, 2));
      else
        OM_TRY(launch_step_mode<3>(h, p, mode, 2));
    }
  }
  if (defer_fetch && !target_only && !om_is_solve_method(h->method)) {
    // om_step reads the statistics back together with the first flip-round readback
    if (h->own_hi >= 0) {
      const size_t off = (size_t)h->own_lo * h->PD, cnt = (size_t)(h->own_hi - h->own_lo) * h->PD;
      if (cnt)
        CUDA_TRY(cudaMemcpyAsync(h->x + off, h->xnew + off, sizeof(double) * cnt,
                                 cudaMemcpyDeviceToDevice, h->stream));
    } else {
      std::swap(h->x, h->xnew);
    }
    return OM_OK;
  }
  OM_TRY(om_fetch_scalars(h));
  OM_TRY(om_check_dev_err(h));
  if (!target_only && h->defer_commit && h->own_hi >= 0 && !om_is_solve_method(h->method)) {
    // partitioned run: the caller commits once every rank reports "no stale coordinate"
    om_step_stats_from_scalars(h, tol, out);
    if (out) {
      out->solver_iters = iters;
      out->reserved = h->hs->stale;
    }
    return OM_OK;
  }
  if (!target_only) {
    if (h->own_hi >= 0 && !om_is_solve_method(h->method)) {
      // sharded step: only [lo, hi) was written; fold it back, the rest of x stays
      const size_t off = (size_t)h->own_lo * h->PD, cnt = (size_t)(h->own_hi - h->own_lo) * h->PD;
      if (cnt)
        CUDA_TRY(cudaMemcpyAsync(h->x + off, h->xnew + off, sizeof(double) * cnt,
                                 cudaMemcpyDeviceToDevice, h->stream));
    } else {
      std::swap(h->x, h->xnew);
    }
  }
  if (!target_only) om_step_stats_from_scalars(h, tol, out);
  if (out) out->solver_iters = iters;
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

// (Re)builds ring rows: all vertices, or the vertices touched by flips since the last call.
int om_rebuild_rings(om_handle* h, bool all) {
  if (!h->ring || h->N == 0) return OM_OK;
  const int B = 256;
  if (all) {
    OM_LAUNCH(h, (k_build_rings<false>), om_grid(h->N, B), B, h->cells, (const int*)h->adj, h->v2c,
              h->bflag, (int)h->N, (const int*)nullptr, h->ring, 0, (int)h->N);
    h->rings_partial = false;
  } else {
    const int n = h->hs->n_dirty;  // fetched by the flip pass
    // with an owned range only its rows are kept current (nothing else reads ring rows);
    // om_set_owned_range(whole mesh) rebuilds all of them
    const bool ranged = h->own_hi >= 0;
    if (ranged) h->rings_partial = true;
    if (n > 0)
      OM_LAUNCH(h, (k_build_rings<true>), om_grid(n, B), B, h->cells, (const int*)h->adj, h->v2c,
                h->bflag, n, h->dirty, h->ring, ranged ? (int)h->own_lo : 0,
                ranged ? (int)h->own_hi : (int)h->N);
  }
  CUDA_TRY(cudaGetLastError());
  return OM_OK;
}

// ---- band of a vertex range: own vertices within `depth` edges of a foreign vertex
// (partitioned coordinates, dist.py).  Neighbours come from the ring row, or from a star walk
// for vertices without one (boundary fans, valence > OM_RING_W).
namespace {

template <typename F>
__device__ __forceinline__ void for_each_neighbour(const int4* __restrict__ cells,
                                                   const int* __restrict__ adj,
                                                   const int* __restrict__ v2c,
                                                   const int* __restrict__ ring, int v, F&& f) {
  const int4* rp = reinterpret_cast<const int4*>(ring + (size_t)OM_RING_W * v);
  const int4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
  const int e[OM_RING_W] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
  if (e[0] != -2) {
#pragma unroll
    for (int q = 0; q < OM_RING_W; q++)
      if (e[q] >= 0) f(e[q] & RING_MASK);
    return;
  }
  const int c0 = v2c[v];
  if (c0 == OM_NONE_CELL) return;
  const int4 cell0 = __ldg(cells + c0);
  const int j = slot_of(cell0, v);
  if (j < 0) return;
  f(cell_get(cell0, (j + 1) % 3));
  f(cell_get(cell0, (j + 2) % 3));
  bool closed = false;
  for (int dir = 0; dir < 2 && !closed; dir++) {
    int cur = c0, kexit = (j + 1 + dir) % 3, hops = 0;
    while (true) {
      const int t = __ldg(adj + 4 * (size_t)cur + kexit);
      if (t < 0) break;
      const int cn = t >> 2, kn = t & 3;
      if (cn == c0) {
        closed = true;
        break;
      }
      const int4 cl = __ldg(cells + cn);
      const int jn = slot_of(cl, v);
      if (jn < 0 || jn == kn || ++hops > MAX_RING) {
        closed = true;
        break;
      }
      f(cell_get(cl, (jn + 1) % 3));
      f(cell_get(cl, (jn + 2) % 3));
      cur = cn;
      kexit = 3 - jn - kn;
    }
  }
}

__global__ void __launch_bounds__(256)
    k_band_mark(const int4* __restrict__ cells, const int* __restrict__ adj,
                const int* __restrict__ v2c, const int* __restrict__ ring, int lo, int hi, int d,
                uint8_t* __restrict__ mark) {
  const int v = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= hi || mark[v] != 0) return;
  bool hit = false;
  for_each_neighbour(cells, adj, v2c, ring, v, [&](int u) {
    if (d == 1)
      hit |= (u < lo || u >= hi);
    else if (u >= lo && u < hi) {
      const int m = mark[u];
      hit |= (m >= 1 && m < d);
    }
  });
  if (hit) mark[v] = (uint8_t)d;
}

__global__ void __launch_bounds__(256)
    k_band_collect(const uint8_t* __restrict__ mark, const uint8_t* __restrict__ bflag, int lo,
                   int hi, int* __restrict__ band, DevScalars* ds) {
  const int v = lo + blockIdx.x * blockDim.x + threadIdx.x;
  const int vals[1] = {v};
  const bool preds[1] = {v < hi && mark[v] != 0 && bflag[v] == 0};  // pinned ones never move
  block_append<1>(&ds->n_over, band, vals, preds);
}

template <int PD>
__global__ void k_band_pack(const double* __restrict__ x, const int* __restrict__ idx, int n,
                            double* __restrict__ buf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double2* src = reinterpret_cast<const double2*>(x) + (size_t)(PD / 2) * idx[i];
  double2* dst = reinterpret_cast<double2*>(buf) + (size_t)(PD / 2) * i;
#pragma unroll
  for (int k = 0; k < PD / 2; k++) dst[k] = src[k];
}

template <int PD>
__global__ void k_band_unpack(double* __restrict__ x, const int* __restrict__ idx, int n,
                              const double* __restrict__ buf, int* __restrict__ valid_epoch,
                              int stamp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int v = idx[i];
  const double2* src = reinterpret_cast<const double2*>(buf) + (size_t)(PD / 2) * i;
  double2* dst = reinterpret_cast<double2*>(x) + (size_t)(PD / 2) * v;
#pragma unroll
  for (int k = 0; k < PD / 2; k++) dst[k] = src[k];
  valid_epoch[v] = stamp;
}

}  // namespace

int om_band_alloc(om_handle* h) {
  if (h->valid_epoch) return OM_OK;
  const size_t N = (size_t)std::max<int64_t>(h->N, 1);
  CUDA_TRY(om_malloc(h, &h->valid_epoch, sizeof(int) * N));
  CUDA_TRY(om_malloc(h, &h->band, sizeof(int) * N));
  CUDA_TRY(om_malloc(h, &h->band_mark, N));
  CUDA_TRY(cudaMemsetAsync(h->valid_epoch, 0, sizeof(int) * N, h->stream));
  return OM_OK;
}

int om_band_build_impl(om_handle* h, int depth, int64_t* n) {
  if (n) *n = 0;
  if (h->own_hi < 0 || h->N == 0) return OM_OK;
  OM_TRY(om_band_alloc(h));
  const int lo = (int)h->own_lo, hi = (int)h->own_hi;
  const int B = 256, G = om_grid(hi - lo, B);
  CUDA_TRY(cudaMemsetAsync(h->band_mark, 0, (size_t)h->N, h->stream));
  CUDA_TRY(cudaMemsetAsync(&h->ds->n_over, 0, sizeof(int), h->stream));
  for (int d = 1; d <= depth; d++)
    OM_LAUNCH(h, k_band_mark, G, B, h->cells, (const int*)h->adj, h->v2c, h->ring, lo, hi, d,
              h->band_mark);
  OM_LAUNCH(h, k_band_collect, G, B, h->band_mark, h->bflag, lo, hi, h->band, h->ds);
  OM_TRY(om_fetch_scalars(h));
  OM_TRY(om_check_dev_err(h));
  if (n) *n = h->hs->n_over;
  return OM_OK;
}

int om_band_pack_impl(om_handle* h, const int* idx_dev, int64_t n, double* buf_dev) {
  if (n <= 0) return OM_OK;
  if (h->PD == 2)
    OM_LAUNCH(h, k_band_pack<2>, om_grid(n, 256), 256, h->x, idx_dev, (int)n, buf_dev);
  else
    OM_LAUNCH(h, k_band_pack<4>, om_grid(n, 256), 256, h->x, idx_dev, (int)n, buf_dev);
  CUDA_TRY(cudaGetLastError());
  return OM_OK;
}

int om_band_unpack_impl(om_handle* h, const int* idx_dev, int64_t n, const double* buf_dev) {
  if (n <= 0) return OM_OK;
  OM_TRY(om_band_alloc(h));
  if (h->PD == 2)
    OM_LAUNCH(h, k_band_unpack<2>, om_grid(n, 256), 256, h->x, idx_dev, (int)n, buf_dev,
              h->valid_epoch, h->valid_stamp);
  else
    OM_LAUNCH(h, k_band_unpack<4>, om_grid(n, 256), 256, h->x, idx_dev, (int)n, buf_dev,
              h->valid_epoch, h->valid_stamp);
  CUDA_TRY(cudaGetLastError());
  return OM_OK;
}

// partitioned run: fold the freshly computed own range into the point array
int om_commit_points_impl(om_handle* h) {
  if (h->own_hi < 0) return OM_OK;
  const size_t off = (size_t)h->own_lo * h->PD, cnt = (size_t)(h->own_hi - h->own_lo) * h->PD;
  if (cnt)
    CUDA_TRY(cudaMemcpyAsync(h->x + off, h->xnew + off, sizeof(double) * cnt,
                             cudaMemcpyDeviceToDevice, h->stream));
  return OM_OK;
}
