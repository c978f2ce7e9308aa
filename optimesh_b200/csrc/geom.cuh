// Per-cell fp64 geometry shared by the step, flip and statistics kernels.
// Arithmetic: SURVEY.md Appendix A.1-A.3 (meshplex MeshTri quantities; reference call
// site /root/reference/README.md:131).
#pragma once
#include "common.cuh"

template <int D>
struct Vec {
  double v[D];
};

template <int D>
__device__ __forceinline__ Vec<D> ld_point(const double* __restrict__ x, int i);

template <>
__device__ __forceinline__ Vec<2> ld_point<2>(const double* __restrict__ x, int i) {
  double2 t = __ldg(reinterpret_cast<const double2*>(x) + i);
  Vec<2> r;
  r.v[0] = t.x;
  r.v[1] = t.y;
  return r;
}
template <>
__device__ __forceinline__ Vec<3> ld_point<3>(const double* __restrict__ x, int i) {
  const double2* p = reinterpret_cast<const double2*>(x) + 2 * (size_t)i;
  double2 a = __ldg(p), b = __ldg(p + 1);
  Vec<3> r;
  r.v[0] = a.x;
  r.v[1] = a.y;
  r.v[2] = b.x;
  return r;
}

// coherent loads, for arrays that the same kernel also writes
template <int D>
__device__ __forceinline__ Vec<D> ld_point_rw(const double* x, int i) {
  Vec<D> r;
  if (D == 2) {
    double2 t = reinterpret_cast<const double2*>(x)[i];
    r.v[0] = t.x;
    r.v[1] = t.y;
  } else {
    const double2* p = reinterpret_cast<const double2*>(x) + 2 * (size_t)i;
    double2 a = p[0], b = p[1];
    r.v[0] = a.x;
    r.v[1] = a.y;
    r.v[D - 1] = b.x;
  }
  return r;
}

template <int D>
__device__ __forceinline__ void st_point(double* __restrict__ x, int i, const Vec<D>& p);
template <>
__device__ __forceinline__ void st_point<2>(double* __restrict__ x, int i, const Vec<2>& p) {
  reinterpret_cast<double2*>(x)[i] = make_double2(p.v[0], p.v[1]);
}
template <>
__device__ __forceinline__ void st_point<3>(double* __restrict__ x, int i, const Vec<3>& p) {
  double2* q = reinterpret_cast<double2*>(x) + 2 * (size_t)i;
  q[0] = make_double2(p.v[0], p.v[1]);
  q[1] = make_double2(p.v[2], 0.0);
}

template <int D>
__device__ __forceinline__ Vec<D> vsub(const Vec<D>& a, const Vec<D>& b) {
  Vec<D> r;
#pragma unroll
  for (int k = 0; k < D; k++) r.v[k] = a.v[k] - b.v[k];
  return r;
}
template <int D>
__device__ __forceinline__ double vdot(const Vec<D>& a, const Vec<D>& b) {
  // explicit fma chain: the same bits in every kernel, whatever the compiler would contract
  double s = a.v[0] * b.v[0];
#pragma unroll
  for (int k = 1; k < D; k++) s = fma(a.v[k], b.v[k], s);
  return s;
}

__device__ __forceinline__ int cell_get(const int4& c, int k) {
  return k == 0 ? c.x : (k == 1 ? c.y : c.z);
}
__device__ __forceinline__ int slot_of(const int4& c, int v) {
  return c.x == v ? 0 : (c.y == v ? 1 : (c.z == v ? 2 : -1));
}

// Edge vectors and dot products of one cell with vertices (P0,P1,P2):
// e_k runs from vertex (k+1)%3 to (k+2)%3; ee_k = e_k.e_k; ed_k = e_{k+1}.e_{k+2}.
template <int D>
struct CellGeo {
  Vec<D> e0, e1, e2;
  double ee0, ee1, ee2, ed0, ed1, ed2;
  double vol2;  // squared area
};

template <int D>
__device__ __forceinline__ CellGeo<D> cell_geo(const Vec<D>& P0, const Vec<D>& P1,
                                                const Vec<D>& P2) {
  CellGeo<D> g;
  g.e0 = vsub<D>(P2, P1);
  g.e1 = vsub<D>(P0, P2);
  g.e2 = vsub<D>(P1, P0);
  g.ee0 = vdot<D>(g.e0, g.e0);
  g.ee1 = vdot<D>(g.e1, g.e1);
  g.ee2 = vdot<D>(g.e2, g.e2);
  g.ed0 = vdot<D>(g.e1, g.e2);
  g.ed1 = vdot<D>(g.e2, g.e0);
  g.ed2 = vdot<D>(g.e0, g.e1);
  g.vol2 = 0.25 * (g.ed2 * g.ed0 + g.ed0 * g.ed1 + g.ed1 * g.ed2);
  return g;
}

// ordered-integer encoding of non-negative doubles for atomicMax
__device__ __forceinline__ void atomic_max_nonneg(unsigned long long* addr, double v) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  // a stale read can only under-estimate the maximum: skipping is safe when bits <= it
  if (bits > *(volatile unsigned long long*)addr) atomicMax(addr, bits);
}

// Block-aggregated list append: every thread of the (256-thread) block calls it with up to K
// (value, predicate) pairs; ONE returning atomic per block reserves the space.  The shared
// counters live on a single address each, and a returning atomic per warp made the flip
// kernels wait on the L2 atomic unit instead of on memory (ncu: > 50 % of the stall samples
// sat on the shuffle that broadcasts the atomic's result).
template <int K>
__device__ __forceinline__ void block_append(int* counter, int* __restrict__ list,
                                             const int (&vals)[K], const bool (&preds)[K]) {
  __shared__ int s_warp[8];
  __shared__ int s_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int cnt = 0;
#pragma unroll
  for (int q = 0; q < K; q++) cnt += preds[q] ? 1 : 0;
  int incl = cnt;  // inclusive scan over the warp
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    const int nw = (int)(blockDim.x >> 5);
    for (int w = 0; w < nw; w++) {
      const int t = s_warp[w];
      s_warp[w] = tot;
      tot += t;
    }
    s_base = tot ? atomicAdd(counter, tot) : 0;
  }
  __syncthreads();
  int pos = s_base + s_warp[warp] + incl - cnt;
#pragma unroll
  for (int q = 0; q < K; q++)
    if (preds[q]) list[pos++] = vals[q];
  __syncthreads();  // s_warp / s_base may be reused by the next call
}

