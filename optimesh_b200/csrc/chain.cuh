// The star of a vertex evaluated as a chain of SPOKES (shared by every step kernel).
//
// With P0 the vertex and n_0 .. n_{k-1} its one-ring in walk order, spoke q is
// d_q = x[n_q] - P0 and cell q = (P0, n_q, n_{q+1}) lies between the spokes q and q+1.  In the
// cell's local terms (A.1 with the vertex in slot 0) e2 = d_q, e1 = -d_{q+1}, so with
//   L_q = d_q.d_q,  c_q = d_q.d_{q+1}:   ee2 = L_q, ee1 = L_{q+1}, ed0 = -c_q,
//   ed1 = c_q - L_q (angle at n_q),  ed2 = c_q - L_{q+1} (angle at n_{q+1}),
//   V4 = L_q L_{q+1} - c_q^2 = (2A)^2.
// Every per-vertex sum of SURVEY.md A.4 / A.8 / A.9 is a combination of the two spokes of each
// cell, hence regroups into ONE coefficient per spoke, fed by the two cells next to it:
//   spoke q:  cH_q = t2(cell q) + t1(cell q-1),  cN_q = s2(cell q) + s1(cell q-1)
//   W += L_q cH_q,   H += cH_q d_q d_q^T,   NUM += cN_q d_q.
// cH_q is (minus twice) the Delaunay indicator s of the edge (v, n_q) (A.7), so the fused
// check of the flip pass costs two more instructions per spoke.
//
// Cost per cell visit (2D, CVT block-diagonal, lazy limiter): 31 fp64 instructions -- one new
// spoke (2 sub, 2 for L, 2 for c), V4 (2), one rsqrt (MUFU seed + 5), ed1/ed2 (2), two t (2),
// w1 w2 (2), ws uu (2), s1 s2 (2), spoke coefficients (2), W (1), H (3: its trace is W, so the
// last diagonal entry is never summed), NUM (2) -- against 46 in the cell-by-cell form (step.cu
// of round 1), and no operand selects: the formulas are symmetric in the orientation of a
// cell, only the walk direction matters.  Everything that is a comparison (degenerate cell,
// masked cell, limiter bound, Delaunay pre-check) runs on the integer pipe: the fp64 pipe is
// the one that bounds the kernel.
// Scaling: rs = 1/sqrt(V4) = 1/(2A) is used as it comes; t'' = ed rs = 2 t (t = -ce, A.2),
// w'' = 2 w, s'' = 4 s.  Constant factors are undone once per vertex (finish()).
//
// Every vertex must get the same bits whichever kernel evaluates it (ring rows with the lazy
// or the exact limiter, with or without the fused check, or a star walk): the file is compiled
// with -fmad=false and every fused multiply-add is written out, so that no instantiation is
// contracted differently from another.
//
// tests/chain_model.py restates this file in Python; tests/test_chain_model.py checks that
// restatement against the oracle on the CPU.
#pragma once
#include "common.cuh"
#include "geom.cuh"

// method id of a chain that only tracks the step limiter (solve methods)
#define OM_CHAIN_LIMITER_ONLY (-1)

// 1/sqrt(x) for positive, normal x: hardware seed (rsqrt.approx.f64, relative error < 2^-22)
// + ONE third-order step  e = 1 - x y^2,  y <- y (1 + e/2 + 3 e^2/8)   (remainder < 2^-67)
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x * y, y, 1.0);
  const double p = fma(0.375, e, 0.5);
  return fma(y * e, p, y);
}

// 1/x for normal x: hardware seed (rcp.approx.f64, < 2^-22) + two Newton steps (< 2^-80)
__device__ __forceinline__ double fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}

// x is a positive, normal, finite double (integer pipe: the fp64 pipe is the busy one)
__device__ __forceinline__ bool pos_normal(double x) {
  return (unsigned)(__double2hiint(x) - 0x00100000) < 0x7fe00000u;
}

template <int D>
__device__ __forceinline__ bool solve_sym(const double* H, double diag, const Vec<D>& rhs,
                                          Vec<D>& out);
template <>
__device__ __forceinline__ bool solve_sym<2>(const double* H, double diag, const Vec<2>& rhs,
                                             Vec<2>& out) {
  // H[2] is not summed: trace(H) = diag, so diag - H[2] = H[0]
  const double a = diag - H[0], b = -H[1], d = H[0];
  const double det = fma(a, d, -b * b);
  if (det == 0.0) return false;
  const double inv = fast_rcp(det);
  out.v[0] = fma(d, rhs.v[0], -b * rhs.v[1]) * inv;
  out.v[1] = fma(a, rhs.v[1], -b * rhs.v[0]) * inv;
  return true;
}
template <>
__device__ __forceinline__ bool solve_sym<3>(const double* H, double diag, const Vec<3>& rhs,
                                             Vec<3>& out) {
  // H[5] is not summed: trace(H) = diag, so diag - H[5] = H[0] + H[3]
  const double a = diag - H[0], b = -H[1], c = -H[2], d = diag - H[3], e = -H[4],
               f = H[0] + H[3];
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

// EXACT: track the smallest incident inradius (as the fraction 2A / perimeter);
// otherwise only a division-free lower bound of it (smallest V4, largest half sum of the
// squared edges, both as high words).  CHECK: collect the spokes that may violate the
// Delaunay criterion (bit q of `flags`).
template <int D, int METHOD, bool EXACT, bool CHECK>
struct Chain {
  static constexpr bool LLOYD_LIKE = METHOD == OM_LLOYD || METHOD == OM_CVT_BLOCK_DIAGONAL;
  static constexpr bool CVT = METHOD == OM_CVT_BLOCK_DIAGONAL;
  static constexpr bool CPT = METHOD == OM_CPT_FIXED_POINT;
  static constexpr bool ODT = METHOD == OM_ODT_FIXED_POINT || METHOD == OM_ODT_DP_FP;
  static constexpr bool DP = METHOD == OM_ODT_DP_FP;
  static constexpr bool NONE = METHOD == OM_CHAIN_LIMITER_ONLY;
  static constexpr bool NEED_T = LLOYD_LIKE || ODT || CHECK;
  static constexpr int NH = D * (D + 1) / 2;

  Vec<D> P0;
  // accumulators (scaled, see the header)
  double W;
  Vec<D> NUM;
  double H[NH];
  // limiter
  double rn, rd;          // EXACT: 2A and perimeter of the cell with the smallest inradius
  int minq_hi;            // lazy bound: min over the cells of hi(V4) - hi(max L)
  int min_v4_hi;          // smallest high word of V4 (degenerate cell: below that of 2^-1022)
  // current spoke (the second spoke of the last cell) and what that cell leaves for it
  Vec<D> dq;
  double Lq, lenq;
  double t1p, s1p;
  // what cell 0 leaves for spoke 0 (closed fans finish it last)
  double t2_0, s2_0;
  unsigned flags;
  int q;      // cells done
  int err;

  __device__ __forceinline__ void init(const Vec<D>& p0) {
    P0 = p0;
    W = 0.0;
#pragma unroll
    for (int k = 0; k < D; k++) NUM.v[k] = 0.0;
#pragma unroll
    for (int k = 0; k < NH; k++) H[k] = 0.0;
    rn = INFINITY;
    rd = 1.0;
    minq_hi = 0x7fffffff;
    min_v4_hi = 0x7fffffff;
    t1p = s1p = t2_0 = s2_0 = 0.0;
    lenq = 0.0;
    flags = 0u;
    q = 0;
    err = 0;
  }

  __device__ __forceinline__ void spoke(const Vec<D>& P, Vec<D>& d, double& L, double& len) {
    d = vsub<D>(P, P0);
    L = vdot<D>(d, d);
    if (EXACT) len = L * fast_rsqrt(L);
  }

  // first spoke of the chain
  __device__ __forceinline__ void start(const Vec<D>& P) { spoke(P, dq, Lq, lenq); }

  // adds what the two cells next to spoke `d` leave for it
  __device__ __forceinline__ void finish_spoke(const Vec<D>& d, double L, double t2, double s2,
                                               double t1, double s1, int bit, bool interior) {
    if (!NONE) {
      const double cN = s2 + s1;
#pragma unroll
      for (int k = 0; k < D; k++) NUM.v[k] = fma(cN, d.v[k], NUM.v[k]);
    }
    if (LLOYD_LIKE || CHECK) {
      const double cH = t2 + t1;
      if (LLOYD_LIKE) W = fma(L, cH, W);
      if (CVT) {
        // trace(H) = sum cH |d|^2 = W: the last diagonal entry follows in target_offset()
        int qi = 0;
#pragma unroll
        for (int i = 0; i < D - 1; i++) {
          const double a = cH * d.v[i];
#pragma unroll
          for (int j = i; j < D; j++) {
            H[qi] = fma(a, d.v[j], H[qi]);
            qi++;
          }
        }
      }
      if (CHECK && interior) {
        // s = ce + ce' < 0  <=>  cH > 0.  Everything above -2^-20 is kept: as an unsigned
        // integer the high word of such a value (positive, +-0, or negative and tiny) is at
        // most that of -2^-20 -- ONE compare.  The rounding error of the sum is below
        // 2^-49 max|T|, far inside the margin for |T| < 2^29; cells beyond that have named
        // their spokes in cell().  The flip pass decides on the exact s.
        if ((unsigned)__double2hiint(cH) <= 0xBEB00000u) flags |= 1u << bit;
      }
    }
  }

  // The cell between the current spoke and the spoke dn.  bary: ODT methods, the cell has a
  // boundary edge and contributes its barycenter.  Leaves what the cell gives to its two
  // spokes: (t2, s2) to the current one, (t1, s1) to dn.
  __device__ __forceinline__ void cell(const Vec<D>& dn, double Ln, double lenn, bool bary,
                                       double& t1, double& t2, double& s1, double& s2,
                                       int bit_cur, int bit_next) {
    unsigned mflags = 0u;  // bit 0: the current spoke, bit 1: dn (set on the rare paths only)
    const double c = vdot<D>(dq, dn);
    const double cc = c * c;
    const double V4 = fma(Lq, Ln, -cc);
    // a degenerate cell (V4 negative, zero or subnormal) raises the error and the step is
    // abandoned by the host: no need to keep its garbage out of the sums.  The smallest high
    // word is tracked (one instruction) and looked at once per vertex, together with
    // non-finite input, which shows up in |d|^2 (finite()).
    min_v4_hi = min(min_v4_hi, __double2hiint(V4));
    const double rs = fast_rsqrt(V4);  // 1 / (2A)
    if (EXACT) {
      // inradius 2A / (l0 + l1 + l2), compared as fractions (no division per cell)
      const double A2 = V4 * rs;
      const double ee0 = fma(-2.0, c, Lq + Ln);
      const double per = fma(ee0, fast_rsqrt(ee0), lenq + lenn);
      if (__double_as_longlong(A2 * rd) < __double_as_longlong(rn * per)) {
        rn = A2;
        rd = per;
      }
    } else {
      // r_in^2 = V4 / (l0 + l1 + l2)^2 >= V4 / (3 sum ee) and ee0 <= 2 (ee1 + ee2), so
      // r_in^2 >= V4 / (18 max(L_q, L_{q+1})) for this cell.  The smallest such quotient over
      // the star is tracked as a DIFFERENCE OF HIGH WORDS (a fixed-point log2 with a known
      // error, undone in proves_unlimited()): three integer instructions per cell, nothing on
      // the fp64 pipe.
      minq_hi = min(minq_hi, __double2hiint(V4) - max(__double2hiint(Lq), __double2hiint(Ln)));
    }
    if (NEED_T) {
      const double T1 = (c - Lq) * rs, T2 = (c - Ln) * rs;  // 2 t: angles at n_q, n_{q+1}
      // Fused Delaunay check, part 1: cH = t2 + t1 of a spoke is compared with an ABSOLUTE
      // threshold (finish_spoke), which covers the rounding of the sum as long as |T| < 2^29.
      // A cell with a larger |T| (an angle below 4e-9 rad: T = -cot is hugely negative, its
      // high word the largest as an unsigned integer) names both its spokes itself.
      bool rare = false;
      if (CHECK)
        rare = max((unsigned)__double2hiint(T1), (unsigned)__double2hiint(T2)) >= 0xC1C00000u;
      if (LLOYD_LIKE) {
        // cell masked (an angle > 135 deg): some t > 1/2, i.e. some T > 1.  T0 = -c rs > 1
        // <=> c < 0 and c^2 > V4.  Masked cells are rare, the three exact tests are not cheap
        // (64-bit integer compares: bit patterns of positive doubles order like integers), so
        // they sit behind a pre-test on the high words that is false for almost every cell:
        // 5 instructions instead of 17 on the issue-bound path.  A masked cell gives nothing
        // to its spokes: with t1 = t2 = 0 every product below is 0.
        rare |= max(__double2hiint(T1), __double2hiint(T2)) >= 0x3ff00000;
        rare |= (__double2hiint(c) < 0) & (__double2hiint(cc) >= __double2hiint(V4));
        t1 = T1;
        t2 = T2;
        if (rare) {
          const long long one = 0x3ff0000000000000ll;
          const bool m0 =
              (__double2hiint(c) < 0) & (__double_as_longlong(cc) > __double_as_longlong(V4));
          const bool m1 = __double_as_longlong(T1) > one, m2 = __double_as_longlong(T2) > one;
          // Fused Delaunay check, part 2: an edge can only violate the criterion if one of
          // its two opposite angles is obtuse.  A masked cell hides its t from the sums below,
          // so it names the spoke opposite its > 135 deg angle itself (the angle at n_{q+1}
          // faces the current spoke, the one at n_q faces dn); its two small angles (< 45 deg
          // together) cannot make their edges violate it unless the cell across does the same.
          mflags = (m2 ? 1u : 0u) | (m1 ? 2u : 0u);
          if (CHECK &&
              max((unsigned)__double2hiint(T1), (unsigned)__double2hiint(T2)) >= 0xC1C00000u)
            mflags = 3u;
          if (m0 | m1 | m2) t1 = t2 = 0.0;
          if (CHECK) flags |= ((mflags & 1u) << bit_cur) | ((mflags >> 1) << bit_next);
        }
        const double w1 = Ln * t1, w2 = Lq * t2;
        const double uu = rs * (w1 + w2);
        s2 = fma(-uu, w1, w2);
        s1 = fma(-uu, w2, w1);
      } else if (ODT) {
        const double A2 = V4 * rs;
        W += DP ? 1.0 : A2;
        if (bary) {
          s1 = s2 = DP ? 1.0 : A2;
        } else {
          const double f = DP ? -1.5 * rs : -1.5;
          s2 = f * (Ln * T1);
          s1 = f * (Lq * T2);
        }
        t1 = T1;
        t2 = T2;
        if (rare) flags |= (1u << bit_cur) | (1u << bit_next);
      } else if (CPT) {
        const double A2 = V4 * rs;
        W += A2;
        s1 = s2 = A2;
        t1 = T1;
        t2 = T2;
        if (rare) flags |= (1u << bit_cur) | (1u << bit_next);
      } else {
        s1 = s2 = 0.0;
        t1 = T1;
        t2 = T2;
        if (rare) flags |= (1u << bit_cur) | (1u << bit_next);
      }
    } else if (CPT) {
      const double A2 = V4 * rs;
      W += A2;
      s1 = s2 = A2;
      t1 = t2 = 0.0;
    } else {
      t1 = t2 = s1 = s2 = 0.0;
    }
  }

  // second spoke of the chain: the first cell.  first_spoke_interior: closed fan (spoke 0 is
  // finished by close()); otherwise spoke 0 is a boundary edge with this cell only.
  __device__ __forceinline__ void first(const Vec<D>& P, bool bary, bool first_spoke_interior) {
    Vec<D> dn;
    double Ln, lenn = 0.0, t1, t2, s1, s2;
    spoke(P, dn, Ln, lenn);
    cell(dn, Ln, lenn, bary, t1, t2, s1, s2, 0, 1);
    t2_0 = t2;
    s2_0 = s2;
    if (!first_spoke_interior) finish_spoke(dq, Lq, t2, s2, 0.0, 0.0, 0, false);
    t1p = t1;
    s1p = s1;
    dq = dn;
    Lq = Ln;
    lenq = lenn;
    q = 1;
  }

  // next spoke of the chain: processes the cell between the current spoke and P
  __device__ __forceinline__ void next(const Vec<D>& P, bool bary) {
    Vec<D> dn;
    double Ln, lenn = 0.0, t1, t2, s1, s2;
    spoke(P, dn, Ln, lenn);
    cell(dn, Ln, lenn, bary, t1, t2, s1, s2, q, q + 1);
    finish_spoke(dq, Lq, t2, s2, t1p, s1p, q, true);
    t1p = t1;
    s1p = s1;
    dq = dn;
    Lq = Ln;
    lenq = lenn;
    q++;
  }

  // closed fan: the cell between the last spoke and the first one (P = first ring vertex)
  __device__ __forceinline__ void close(const Vec<D>& P, bool bary) {
    Vec<D> dn;
    double Ln, lenn = 0.0, t1, t2, s1, s2;
    spoke(P, dn, Ln, lenn);
    cell(dn, Ln, lenn, bary, t1, t2, s1, s2, q, 0);
    finish_spoke(dq, Lq, t2, s2, t1p, s1p, q, true);
    finish_spoke(dn, Ln, t2_0, s2_0, t1, s1, 0, true);
    q++;
  }

  // open fan: the last spoke is a boundary edge with the last cell only
  __device__ __forceinline__ void end_open() { finish_spoke(dq, Lq, 0.0, 0.0, t1p, s1p, q, false); }

  // offset of the un-relaxed target from the vertex; false: 0/0 (every cell masked) or a
  // singular block -- the vertex stays where it is
  __device__ __forceinline__ bool target_offset(Vec<D>& d) const {
    if (NONE) return false;
    if (W == 0.0) return false;
    if (CVT) {
      // (2 cv I + Hess) d = -2 cv (x - c)  <=>  (W I - H) d = NUM / 6 in the scaled sums
      Vec<D> rhs;
#pragma unroll
      for (int k = 0; k < D; k++) rhs.v[k] = NUM.v[k] * (1.0 / 6.0);
      return solve_sym<D>(H, W, rhs, d);
    }
    const double inv = fast_rcp((LLOYD_LIKE ? 6.0 : 3.0) * W);
#pragma unroll
    for (int k = 0; k < D; k++) d.v[k] = NUM.v[k] * inv;
    return true;
  }

  // |d|^2 of the relaxed update must be a finite number (non-finite coordinates, overflow)
  __device__ __forceinline__ void finite(double diff2) {
    if ((__double2hiint(diff2) & 0x7ff00000) == 0x7ff00000) err |= OM_DEV_DEGENERATE;
  }
  // OM_DEV_* bits of the star (read once per vertex, after the last cell)
  __device__ __forceinline__ int error() const {
    return err | (min_v4_hi < 0x00100000 ? (int)OM_DEV_DEGENERATE : 0);
  }

  // lazy limiter: true if |d|^2 = diff2 provably stays below (r_in / 2)^2 for every cell
  __device__ __forceinline__ bool proves_unlimited(double diff2) const {
    // With h(x) the high word of x > 0, (h(a) - h(b) - 1) / 2^20 is a lower bound of
    // log2(a / b) up to the error of the piecewise-linear log2 the exponent/mantissa layout
    // amounts to: (1 + f) / 2^f lies in [1, 1.0615], so a / b >= D / 1.0615^2 with D the
    // double whose high word is h(a) - h(b) - 1 + 0x3ff00000.  Hence r_in^2 >= D / (18 * 1.127)
    // for every cell and |d|^2 <= r_in^2 / 4 follows from 81.2 |d|^2 <= D.
    if (minq_hi < -0x3fe00000 || minq_hi > 0x3fe00000) return false;  // out of the double range
    const double Dq = __hiloint2double(minq_hi - 1 + 0x3ff00000, 0);
    return 81.2 * diff2 <= Dq;
  }

  // exact limiter: scales d if it is longer than half the smallest incident inradius
  __device__ __forceinline__ bool limit(Vec<D>& d, double diff2) const {
    // |d| > rn / (2 rd)  <=>  4 diff2 rd^2 > rn^2
    const double lhs = 4.0 * diff2 * rd * rd, rhs = rn * rn;
    if (!(lhs > rhs)) return false;
    const double s = 0.5 * rn / (rd * sqrt(diff2));
#pragma unroll
    for (int k = 0; k < D; k++) d.v[k] *= s;
    return true;
  }
};
