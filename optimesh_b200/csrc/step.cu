// K1: the fused smoothing step.
//
// Replaces, per step: get_new_points of the fixed-point methods
// (/root/reference/README.md:80, :90, :104, :141), the numpy scatter-adds
// (np.bincount / np.add.at / np.minimum.at) and the body of the optimize() loop
// (README.md:131-132).  Arithmetic: SURVEY.md Appendix A.2-A.5, A.8, A.9, regrouped per spoke
// (chain.cuh).
//
//   k_step_ring  one thread per vertex of [lo, hi): the one-ring comes from the vertex's ring
//                row (8 neighbour ids in walk order), the ring coordinates are staged in
//                shared memory by cp.async (all in flight at once, slot [q][thread]: conflict
//                free, every thread reads only what it staged -- no barrier anywhere in the
//                kernel), the star is evaluated as a chain of spokes, then method formula,
//                pin, omega, limiter, write.  Per vertex it leaves |diff|^2 (sign bit = "was
//                limited") in `diff2` and, where something is left to do, a flag word:
//                bit 8 "update me in k_post" (no ring row: more than 8 cells; or the lazy
//                limiter bound failed), bits 0-7 "spoke q may violate the Delaunay criterion",
//                bit 9 "check all my spokes" (fused check of the flip pass, CHECK variants).
//   k_post       scans the flag words, compacts the flagged vertices block by block and
//                updates them by a star walk over the twin table with the exact limiter.
//   k_walk_list  the same walk for an explicit vertex list (vertices whose star was changed
//                by flips: the pipelined loop recomputes them, flip.cu).
//   k_reduce_stats  max |diff|^2 and the number of limited vertices from `diff2`.
// No kernel on this path uses an atomic on coordinates or a floating-point atomic: every sum
// runs in walk order, which is a function of the mesh only (bitwise reproducible).
#include <cstdlib>
#include <cstring>
#include <utility>

#include "chain.cuh"
#include "common.cuh"
#include "geom.cuh"

// tuning knobs of the ring-row step kernel
#ifndef OM_K1_BLOCK
#define OM_K1_BLOCK 128
#endif
#ifndef OM_K1_MINB
#define OM_K1_MINB 8
#endif
#ifndef OM_K1_UNROLL
#define OM_K1_UNROLL 1
#endif
#ifndef OM_K1_MINB_EXACT
#define OM_K1_MINB_EXACT 6
#endif

namespace {

constexpr int MAX_RING = 4096;
constexpr int K1_UNROLL = OM_K1_UNROLL;

// ---- ring rows: the one-ring of a free interior vertex as OM_RING_W neighbour vertex ids in
// walk order.  Entry q < k holds n_q | (b_q << 29) where b_q says that cell q = (v, n_q,
// n_{q+1}) has a boundary edge (ODT uses its barycenter); bit 30 of the entries 0..2 holds
// the three bits of k - 1 (3 <= k <= 8 cells); unused entries hold v itself, so that the
// staging copy of all eight entries is unconditional.  Entry 0 < 0 marks a vertex without a
// row: RING_WALK (free vertex with more than OM_RING_W cells: k_post walks its star),
// RING_FIXED (pinned or boundary vertex with cells), RING_ORPHAN (no cell).
constexpr int RING_BCELL = 1 << 29;
constexpr int RING_KBIT = 1 << 30;
constexpr int RING_MASK = RING_BCELL - 1;
constexpr int RING_WALK = -2, RING_FIXED = -3, RING_ORPHAN = -4;

// flag word of a vertex (om_handle::vflags)
constexpr unsigned VF_SPOKES = 0xffu, VF_DEFER = 0x100u, VF_CHECKALL = 0x200u;

template <int D>
__host__ __device__ constexpr int step_block() {
  return OM_K1_BLOCK;
}

template <bool LIST>
__global__ void __launch_bounds__(256)
    k_build_rings(const int4* __restrict__ cells, const int* __restrict__ adj,
                  const int* __restrict__ v2c, const uint8_t* __restrict__ bflag, int n,
                  const int* __restrict__ list, const int* __restrict__ n_dev,
                  int* __restrict__ ring, int* __restrict__ ringc, int lo, int hi,
                  const int* __restrict__ halt) {
  if (halt && (*halt & 1)) return;  // (a flush iteration still rebuilds rows)
  if (n_dev) n = *n_dev;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int v = LIST ? list[i] : i;
    if (v < lo || v >= hi) continue;  // partitioned run: only the own range's rows are read
    int e[OM_RING_W], ec[OM_RING_W];  // neighbour ids and the cells between them
#pragma unroll
    for (int q = 0; q < OM_RING_W; q++) {
      e[q] = v;
      ec[q] = -1;
    }
    int marker = RING_ORPHAN;
    const int c0 = v2c[v];
    if (c0 != OM_NONE_CELL) {
      marker = RING_FIXED;
      if (!bflag[v]) {
        marker = RING_WALK;
        const int4 cell0 = __ldg(cells + c0);
        const int j = slot_of(cell0, v);
        if (j >= 0) {
          int k = 0;  // cells closed so far
          int last = cell_get(cell0, (j + 2) % 3);
          e[0] = cell_get(cell0, (j + 1) % 3);  // cell 0 = (v, n0, n1)
          int cur = c0, kexit = (j + 1) % 3;
          bool ok = true;
          while (true) {
            // twin row of cell k (`cur`): the exit edge, and whether it has a boundary edge
            const int4 ta = __ldg(reinterpret_cast<const int4*>(adj) + cur);
            const bool bc = ta.x < 0 || ta.y < 0 || ta.z < 0;
#pragma unroll
            for (int q = 0; q < OM_RING_W; q++)
              if (q == k) {
                if (bc) e[q] |= RING_BCELL;
                ec[q] = cur;  // cell q = (v, n_q, n_{q+1})
              }
            k++;
            const int t = cell_get(ta, kexit);
            if (t < 0) {  // open fan (cannot happen for a vertex that is not on the boundary)
              ok = false;
              break;
            }
            const int cn = t >> 2, kn = t & 3;
            if (cn == c0) break;  // closed: the vertex opposite the entry edge is n0 again
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
#pragma unroll
            for (int q = 1; q < OM_RING_W; q++)
              if (q == k) e[q] = last;
            last = cell_get(cl, kn);  // the vertex opposite the edge we came through
            cur = cn;
            kexit = 3 - jn - kn;
          }
          if (ok && k >= 3) {
            marker = 0;
            const int km = k - 1;
            e[0] |= (km & 1) ? RING_KBIT : 0;
            e[1] |= (km & 2) ? RING_KBIT : 0;
            e[2] |= (km & 4) ? RING_KBIT : 0;
          }
        }
      }
    }
    if (marker) e[0] = marker;
    int4* out = reinterpret_cast<int4*>(ring + (size_t)OM_RING_W * v);
    out[0] = make_int4(e[0], e[1], e[2], e[3]);
    out[1] = make_int4(e[4], e[5], e[6], e[7]);
    if (ringc && !marker) {
      // only read for vertices with a row and a flagged spoke (fused Delaunay check, flip.cu)
      int4* outc = reinterpret_cast<int4*>(ringc + (size_t)OM_RING_W * v);
      outc[0] = make_int4(ec[0], ec[1], ec[2], ec[3]);
      outc[1] = make_int4(ec[4], ec[5], ec[6], ec[7]);
    }
  }
}

struct StepParams {
  const double* x;
  double* xout;
  const int4* cells;
  const int* adj;  // flat view of the int4 twin table: adj[4*c + k]
  const int* v2c;
  const uint8_t* bflag;
  const int* ring;          // N x OM_RING_W ring rows
  double* diff2;            // N: |omega (target - x)|^2, sign bit = limited
  unsigned short* vflags;   // N: flag words (zero except between the kernels of one step)
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
  int odt_bary;    // ODT: cells with a boundary edge contribute their barycenter
  int force_walk;  // diagnostics (OM_NO_RINGS): every free vertex goes through k_post
  int prefetch_ahead;  // vertices between a block and the one that runs a wave later
  int gate;  // pipelined loop: return at once if the loop has halted / the other limiter mode is on
  DevScalars* ds;
};

// ------------------------------------------------------------------ ring-row kernel
// EXACT: exact inradius in this pass (most vertices limited: early steps); otherwise the
// lazy bound, and the few vertices that fail it are left to k_post.  CHECK: also collect the
// spokes that may violate the Delaunay criterion.  PART: coordinates are partitioned, every
// foreign ring vertex is validated.
template <int D, int METHOD, bool EXACT, bool CHECK, bool PART>
__global__ void __launch_bounds__(OM_K1_BLOCK, (D == 2 ? (EXACT ? OM_K1_MINB_EXACT : OM_K1_MINB) : 4))
    k_step_ring(StepParams p) {
  constexpr int BLOCK = OM_K1_BLOCK;
  constexpr int PER = (D == 2) ? 1 : 2;  // 16-byte pieces per vertex
  constexpr bool ODT = METHOD == OM_ODT_FIXED_POINT || METHOD == OM_ODT_DP_FP;
  __shared__ double2 ring_sm[OM_RING_W * PER * BLOCK];
  if (p.gate && ((p.ds->halt & 1) || (p.ds->mode_exact == 1) != EXACT)) return;
  const int v = p.lo + (int)(blockIdx.x * BLOCK + threadIdx.x);
  if (v >= p.hi) return;
  const int4* rp = reinterpret_cast<const int4*>(p.ring + (size_t)OM_RING_W * v);
  const int4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
  const Vec<D> P0 = ld_point<D>(p.x, v);
#ifndef OM_K1_NO_PREFETCH
  {
    // The first thing a warp does is wait for its ring row to arrive from DRAM (15 % of the
    // stall samples).  Ask L2 for the row and the point of the vertex that the block one
    // wave later will start with: same DRAM traffic, but that block finds them in L2.
    const int vp = v + p.prefetch_ahead;
    if (vp < p.hi) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p.ring + (size_t)OM_RING_W * vp));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p.x + (size_t)(D == 2 ? 2 : 4) * vp));
    }
  }
#endif
  if (r0.x < 0 || p.force_walk) {
    if (r0.x == RING_WALK || (p.force_walk && r0.x >= 0)) {
      // free vertex without a row: k_post updates it and checks its spokes
      p.vflags[v] = (unsigned short)(VF_DEFER | (CHECK ? VF_CHECKALL : 0u));
    } else {
      st_point<D>(p.xout, v, P0);
      p.diff2[v] = 0.0;
      if (CHECK && r0.x == RING_FIXED) p.vflags[v] = (unsigned short)VF_CHECKALL;
    }
    return;
  }
  const int e[OM_RING_W] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
  const int k = 1 + ((e[0] >> 30) & 1) + ((e[1] >> 29) & 2) + ((e[2] >> 28) & 4);
  unsigned bcells = 0u;
#pragma unroll
  for (int q = 0; q < OM_RING_W; q++) {
    const unsigned u = (unsigned)e[q] & (unsigned)RING_MASK;
    if (PART && q < k && !p.valid((int)u)) p.ds->stale = 1;  // caller refreshes and repeats
    if (ODT) bcells |= ((unsigned)(e[q] >> 29) & 1u) << q;
    const double2* src = reinterpret_cast<const double2*>(
        reinterpret_cast<const char*>(p.x) + (unsigned long long)u * (16ull * PER));
#pragma unroll
    for (int h2 = 0; h2 < PER; h2++) {
      const unsigned dst =
          (unsigned)__cvta_generic_to_shared(&ring_sm[(q * PER + h2) * BLOCK + threadIdx.x]);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + h2));
    }
  }
  asm volatile("cp.async.commit_group;");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  auto ld_ring = [&](int q) {
    Vec<D> r;
    const double2 a = ring_sm[(q * PER) * BLOCK + threadIdx.x];
    r.v[0] = a.x;
    r.v[1] = a.y;
    if (D == 3) r.v[D - 1] = ring_sm[(q * PER + PER - 1) * BLOCK + threadIdx.x].x;
    return r;
  };
  const bool odt_bary = ODT && p.odt_bary != 0;
  Chain<D, METHOD, EXACT, CHECK> ch;
  ch.init(P0);
  ch.start(ld_ring(0));
  ch.first(ld_ring(1), odt_bary && (bcells & 1u), true);
#pragma unroll K1_UNROLL
  for (int j = 2; j < k; j++) ch.next(ld_ring(j), odt_bary && ((bcells >> (j - 1)) & 1u));
  ch.close(ld_ring(0), odt_bary && ((bcells >> (k - 1)) & 1u));

  Vec<D> d;
  Vec<D> out = P0;
  double diff2 = 0.0;
  bool deferred = false, limited = false;
  if (ch.target_offset(d)) {
#pragma unroll
    for (int i = 0; i < D; i++) d.v[i] *= p.omega;
    diff2 = vdot<D>(d, d);
    ch.finite(diff2);
    if (p.limiter) {
      if (EXACT) {
        limited = ch.limit(d, diff2);
      } else if (!ch.proves_unlimited(diff2)) {
        // The division-free bound cannot rule the limiter out (about 1.3 x the limited
        // vertices): k_post compacts these vertices and evaluates the smallest incident
        // inradius exactly from their ring rows, with full warps.  (Evaluating it here, by a
        // second pass over the staged ring, was measured slower even when almost no lane
        // needs it: 0.293 vs 0.273 ms, the extra code costs registers on the main path.)
        deferred = true;
      }
    }
#pragma unroll
    for (int i = 0; i < D; i++) out.v[i] = P0.v[i] + d.v[i];
  }
  if (!deferred) {
    st_point<D>(p.xout, v, out);
    p.diff2[v] = limited ? -diff2 : diff2;
  }
  const unsigned f = (CHECK ? (ch.flags & VF_SPOKES) : 0u) | (deferred ? VF_DEFER : 0u);
  if (f) p.vflags[v] = (unsigned short)f;
  if (const int e2 = ch.error()) atomicOr(&p.ds->err, e2);
}

// ------------------------------------------------------------------ star walk
// Updates one vertex by walking its star through the twin table, in the order the ring row
// would list it (so a vertex gets the same bits whichever kernel updates it): closed fans start
// at v2c[v]; open fans (boundary vertices: Lloyd targets only) are first rewound to one end.
// TARGET: un-relaxed, un-limited target (get_new_points); otherwise the step with the exact
// limiter.
template <int D, int METHOD, bool TARGET>
__device__ __forceinline__ void walk_vertex(const StepParams& p, int v, int& err) {
  constexpr bool ODT = METHOD == OM_ODT_FIXED_POINT || METHOD == OM_ODT_DP_FP;
  const Vec<D> P0 = ld_point<D>(p.x, v);
  Vec<D> out = P0;
  double diff2 = 0.0;
  bool limited = false;
  const int c0 = p.v2c[v];
  const bool pinned = p.bflag[v] != 0;
  // the Lloyd target of a boundary vertex is its real control-volume centroid; every other
  // consumer pins boundary vertices
  const bool move = (c0 != OM_NONE_CELL) && (!pinned || (TARGET && METHOD == OM_LLOYD));
  if (move) {
    int cur = c0;
    int4 cl = __ldg(p.cells + cur);
    int jc = slot_of(cl, v);
    bool open = false, bad = jc < 0;
    if (!bad && pinned) {
      // rewind against the walk direction until a boundary edge (open) or back at c0 (closed)
      int kexit = (jc + 2) % 3, hops = 0;
      while (true) {
        const int t = __ldg(p.adj + 4 * (size_t)cur + kexit);
        if (t < 0) {
          open = true;
          break;
        }
        const int cn = t >> 2, kn = t & 3;
        if (cn == c0) {
          cur = c0;
          cl = __ldg(p.cells + cur);
          jc = slot_of(cl, v);
          break;
        }
        const int4 cln = __ldg(p.cells + cn);
        const int jn = slot_of(cln, v);
        if (jn < 0 || jn == kn || ++hops > MAX_RING) {
          bad = true;
          break;
        }
        cur = cn;
        cl = cln;
        jc = jn;
        kexit = 3 - jn - kn;
      }
      // open: the boundary edge of `cur` is opposite slot kexit; walk away from it
      if (open) jc = jc | (kexit << 2);
    }
    if (bad) {
      err |= OM_DEV_WALK;
    } else {
      // first and second spoke of the start cell, and the edge we leave through
      int sf, ss;
      if (open) {
        const int kb = jc >> 2;
        jc &= 3;
        ss = kb;             // the vertex opposite the boundary edge comes second
        sf = 3 - jc - kb;    // the other end of the boundary edge comes first
      } else {
        sf = (jc + 1) % 3;
        ss = (jc + 2) % 3;
      }
      const int first = cell_get(cl, sf);
      int newid = cell_get(cl, ss);
      if (!(p.valid(first) && p.valid(newid))) p.ds->stale = 1;
      const Vec<D> Pfirst = ld_point<D>(p.x, first);
      const bool odt_bary = ODT && p.odt_bary != 0;
      Chain<D, METHOD, !TARGET, false> ch;
      ch.init(P0);
      ch.start(Pfirst);
      int4 ta = __ldg(reinterpret_cast<const int4*>(p.adj) + cur);
      ch.first(ld_point<D>(p.x, newid), odt_bary && (ta.x < 0 || ta.y < 0 || ta.z < 0), !open);
      int kexit = sf;  // leave through the edge (v, second spoke): opposite the first spoke
      int hops = 0;
      while (true) {
        const int t = cell_get(ta, kexit);
        if (t < 0) {
          if (open)
            ch.end_open();
          else
            err |= OM_DEV_WALK;
          break;
        }
        const int cn = t >> 2, kn = t & 3;
        const int4 cln = __ldg(p.cells + cn);
        const int jn = slot_of(cln, v);
        if (jn < 0 || jn == kn || ++hops > MAX_RING) {
          err |= OM_DEV_WALK;
          break;
        }
        newid = cell_get(cln, kn);
        ta = __ldg(reinterpret_cast<const int4*>(p.adj) + cn);
        const bool bary = odt_bary && (ta.x < 0 || ta.y < 0 || ta.z < 0);
        if (!open && newid == first) {
          ch.close(Pfirst, bary);
          break;
        }
        if (!p.valid(newid)) p.ds->stale = 1;
        ch.next(ld_point<D>(p.x, newid), bary);
        kexit = 3 - jn - kn;
      }
      err |= ch.error();
      Vec<D> d;
      if (ch.target_offset(d) && !(pinned && !TARGET)) {
        if (!TARGET) {
#pragma unroll
          for (int i = 0; i < D; i++) d.v[i] *= p.omega;
          diff2 = vdot<D>(d, d);
          ch.finite(diff2);
          err |= ch.error();
          if (p.limiter) limited = ch.limit(d, diff2);
        }
#pragma unroll
        for (int i = 0; i < D; i++) out.v[i] = P0.v[i] + d.v[i];
      }
    }
  }
  st_point<D>(p.xout, v, out);
  if (!TARGET) p.diff2[v] = limited ? -diff2 : diff2;
}

// One vertex from its ring row, outside the main kernel (vertices the lazy limiter left over,
// vertices whose star was changed by flips): the ring coordinates are gathered straight into
// registers, the chain is the main kernel's with the exact limiter -- the same bits as the
// main kernel and as the star walk.  false: the vertex has no row (the caller walks its star).
template <int D, int METHOD>
__device__ __forceinline__ bool ring_vertex(const StepParams& p, int v, int& err) {
  constexpr bool ODT = METHOD == OM_ODT_FIXED_POINT || METHOD == OM_ODT_DP_FP;
  const int4* rp = reinterpret_cast<const int4*>(p.ring + (size_t)OM_RING_W * v);
  const int4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
  if (r0.x < 0 || p.force_walk) return false;
  const int e[OM_RING_W] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
  const int k = 1 + ((e[0] >> 30) & 1) + ((e[1] >> 29) & 2) + ((e[2] >> 28) & 4);
  const Vec<D> P0 = ld_point<D>(p.x, v);
  Vec<D> R[OM_RING_W];
  unsigned bcells = 0u;
#pragma unroll
  for (int q = 0; q < OM_RING_W; q++) {
    const int u = e[q] & RING_MASK;  // unused entries hold v itself: always a valid address
    if (q < k && !p.valid(u)) p.ds->stale = 1;
    if (ODT) bcells |= ((unsigned)(e[q] >> 29) & 1u) << q;
    R[q] = ld_point<D>(p.x, u);
  }
  const bool odt_bary = ODT && p.odt_bary != 0;
  Chain<D, METHOD, true, false> ch;
  ch.init(P0);
  ch.start(R[0]);
  ch.first(R[1], odt_bary && (bcells & 1u), true);
#pragma unroll
  for (int j = 2; j < OM_RING_W; j++)
    if (j < k) ch.next(R[j], odt_bary && ((bcells >> (j - 1)) & 1u));
  ch.close(R[0], odt_bary && ((bcells >> (k - 1)) & 1u));
  err |= ch.error();
  Vec<D> d;
  Vec<D> out = P0;
  double diff2 = 0.0;
  bool limited = false;
  if (ch.target_offset(d)) {
#pragma unroll
    for (int i = 0; i < D; i++) d.v[i] *= p.omega;
    diff2 = vdot<D>(d, d);
    ch.finite(diff2);
    err |= ch.error();
    if (p.limiter) limited = ch.limit(d, diff2);
#pragma unroll
    for (int i = 0; i < D; i++) out.v[i] = P0.v[i] + d.v[i];
  }
  st_point<D>(p.xout, v, out);
  p.diff2[v] = limited ? -diff2 : diff2;
  return true;
}

template <int D, int METHOD>
__device__ __forceinline__ void update_vertex(const StepParams& p, int v, int& err) {
  if (!ring_vertex<D, METHOD>(p, v, err)) walk_vertex<D, METHOD, false>(p, v, err);
}

// get_new_points: every vertex by its star walk
template <int D, int METHOD>
__global__ void __launch_bounds__(128) k_target(StepParams p) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= p.N) return;
  int err = 0;
  walk_vertex<D, METHOD, true>(p, v, err);
  if (err) atomicOr(&p.ds->err, err);
}

// Vertices left by k_step_ring (flag bit VF_DEFER).  Every WARP scans runs of 256 flag words
// (one 16-byte load per lane), compacts the flagged vertices of its run into shared memory
// with ballots and a shuffle scan -- no block barrier, no global counter, no atomics; the
// order is the vertex order -- and updates them with its lanes.
constexpr int POST_BLOCK = 256;
constexpr int POST_PER = 8;                  // flag words per lane (one 16-byte load)
constexpr int POST_RUN = 32 * POST_PER;      // vertices per warp and trip
template <int D, int METHOD>
__global__ void __launch_bounds__(POST_BLOCK) k_post(StepParams p) {
  if (p.gate && (p.ds->halt & 1)) return;
  __shared__ int s_v[POST_BLOCK / 32][POST_RUN];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gwarp = blockIdx.x * (POST_BLOCK / 32) + warp, nwarps = gridDim.x * (POST_BLOCK / 32);
  const int base0 = p.lo & ~(POST_PER - 1);  // 16-byte aligned start of the scan
  const int nruns = (p.hi - base0 + POST_RUN - 1) / POST_RUN;
  int err = 0;
  for (int run = gwarp; run < nruns; run += nwarps) {
    const int vb = base0 + run * POST_RUN + lane * POST_PER;
    unsigned hits = 0u;  // bit i: vertex vb + i is to be updated here
    if (vb < p.hi) {
      // (the flag array is padded to a multiple of POST_PER words beyond N)
      uint4 raw = *reinterpret_cast<const uint4*>(p.vflags + vb);
      unsigned r[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
      for (int i = 0; i < POST_PER; i++) {
        const unsigned bit = VF_DEFER << (16 * (i & 1));
        const int v = vb + i;
        if ((r[i >> 1] & bit) && v >= p.lo && v < p.hi) {
          hits |= 1u << i;
          r[i >> 1] &= ~bit;  // the other bits belong to the flip pass
        }
      }
      if (hits) {
        raw = make_uint4(r[0], r[1], r[2], r[3]);
        *reinterpret_cast<uint4*>(p.vflags + vb) = raw;
      }
    }
    if (!__any_sync(0xffffffffu, hits != 0u)) continue;
    const int cnt = __popc(hits);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (lane == 0) atomicAdd(&p.ds->n_deferred, total);
    int pos = incl - cnt;
    while (hits) {
      const int i = __ffs(hits) - 1;
      hits &= hits - 1;
      s_v[warp][pos++] = vb + i;
    }
    __syncwarp();
    for (int i = lane; i < total; i += 32) update_vertex<D, METHOD>(p, s_v[warp][i], err);
    __syncwarp();  // s_v is reused by the next run
  }
  if (err) atomicOr(&p.ds->err, err);
}

// the same walk for an explicit list whose length lives on the device
template <int D, int METHOD>
__global__ void __launch_bounds__(128)
    k_walk_list(StepParams p, const int* __restrict__ list, const int* __restrict__ n_dev) {
  if (p.gate && p.ds->halt) return;
  const int n = *n_dev;
  int err = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int v = list[i];
    if (v >= p.lo && v < p.hi) update_vertex<D, METHOD>(p, v, err);
  }
  if (err) atomicOr(&p.ds->err, err);
}

// max |diff|^2 and number of limited vertices (sign bit) over [lo, hi)
__global__ void __launch_bounds__(256)
    k_reduce_stats(const double* __restrict__ diff2, int lo, int hi, DevScalars* ds, int gate) {
  if (gate && ds->halt) return;
  unsigned long long mx = 0ull;
  int lim = 0;
  for (int v = lo + blockIdx.x * blockDim.x + threadIdx.x; v < hi; v += gridDim.x * blockDim.x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(__ldg(diff2 + v));
    lim += (int)(b >> 63);
    mx = max(mx, b & 0x7fffffffffffffffull);  // non-negative doubles order like integers
  }
  __shared__ unsigned long long s_m[8];
  __shared__ int s_l[8];
  for (int o = 16; o > 0; o >>= 1) {
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    lim += __shfl_xor_sync(0xffffffffu, lim, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s_m[threadIdx.x >> 5] = mx;
    s_l[threadIdx.x >> 5] = lim;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) {
      mx = max(mx, s_m[w]);
      lim += s_l[w];
    }
    if (mx) atomicMax(&ds->max_diff2_bits, mx);
    if (lim) atomicAdd(&ds->n_limited, (unsigned long long)lim);
  }
}

template <int D, int METHOD, bool CHECK>
int launch_ring(om_handle* h, const StepParams& p, bool exact, bool part) {
  const int B = OM_K1_BLOCK, G = om_grid(std::max(p.hi - p.lo, 0), B);
  if (G == 0 && !h->sh) return OM_OK;  // (a rank of a shared mesh enqueues it anyway: OM_LAUNCH)
  if (exact) {
    if (part)
      OM_LAUNCH(h, (k_step_ring<D, METHOD, true, CHECK, true>), G, B, p);
    else
      OM_LAUNCH(h, (k_step_ring<D, METHOD, true, CHECK, false>), G, B, p);
  } else {
    if (part)
      OM_LAUNCH(h, (k_step_ring<D, METHOD, false, CHECK, true>), G, B, p);
    else
      OM_LAUNCH(h, (k_step_ring<D, METHOD, false, CHECK, false>), G, B, p);
  }
  return OM_OK;
}

// what: 0 ring kernel, 1 ring kernel with the fused Delaunay check, 2 k_post, 3 k_target,
// 4 k_walk_list over h->dirty
template <int D, int METHOD>
int launch_method(om_handle* h, const StepParams& p, int what, bool exact, bool part) {
  switch (what) {
    case 0:
      return launch_ring<D, METHOD, false>(h, p, exact, part);
    case 1:
      return launch_ring<D, METHOD, true>(h, p, exact, part);
    case 2: {
      const int blocks = om_grid(p.hi - (p.lo & ~(POST_PER - 1)), POST_BLOCK * POST_PER);
      OM_LAUNCH(h, (k_post<D, METHOD>), std::min(blocks, 148 * 8), POST_BLOCK, p);
      return OM_OK;
    }
    case 3:
      OM_LAUNCH(h, (k_target<D, METHOD>), om_grid(p.N, 128), 128, p);
      return OM_OK;
    default:
      OM_LAUNCH(h, (k_walk_list<D, METHOD>), 148 * 8, 128, p, (const int*)h->dirty,
                (const int*)&h->ds->n_dirty);
      return OM_OK;
  }
}

template <int D>
int launch_dim(om_handle* h, const StepParams& p, int what, bool exact, bool part) {
  switch (h->method) {
    case OM_LLOYD:
      return launch_method<D, OM_LLOYD>(h, p, what, exact, part);
    case OM_CVT_BLOCK_DIAGONAL:
      return launch_method<D, OM_CVT_BLOCK_DIAGONAL>(h, p, what, exact, part);
    case OM_CPT_FIXED_POINT:
      return launch_method<D, OM_CPT_FIXED_POINT>(h, p, what, exact, part);
    case OM_ODT_FIXED_POINT:
      return launch_method<D, OM_ODT_FIXED_POINT>(h, p, what, exact, part);
    case OM_ODT_DP_FP:
      return launch_method<D, OM_ODT_DP_FP>(h, p, what, exact, part);
    default:
      om_set_error("method %d has no fixed-point kernel", h->method);
      return OM_ERR_ARG;
  }
}

int launch_step(om_handle* h, const StepParams& p, int what, bool exact = true) {
  const bool part = p.valid_epoch != nullptr;
  const int rc = h->D == 2 ? launch_dim<2>(h, p, what, exact, part)
                           : launch_dim<3>(h, p, what, exact, part);
  if (rc != OM_OK) return rc;
  CUDA_TRY(cudaGetLastError());
  return OM_OK;
}

// Smallest incident inradius of a free vertex with a closed fan, by a star walk (as the
// fraction rn / rd = 2A / perimeter of the minimising cell).
template <int D>
__device__ __forceinline__ void star_limiter(const StepParams& p, int v, int c0, const Vec<D>& P0,
                                             Chain<D, OM_CHAIN_LIMITER_ONLY, true, false>& ch,
                                             int& err) {
  ch.init(P0);
  const int4 cl = __ldg(p.cells + c0);
  const int j = slot_of(cl, v);
  if (j < 0) {
    err |= OM_DEV_WALK;
    return;
  }
  const int first = cell_get(cl, (j + 1) % 3);
  const Vec<D> Pfirst = ld_point<D>(p.x, first);
  ch.start(Pfirst);
  ch.first(ld_point<D>(p.x, cell_get(cl, (j + 2) % 3)), false, true);
  int cur = c0, kexit = (j + 1) % 3, hops = 0;
  while (true) {
    const int t = __ldg(p.adj + 4 * (size_t)cur + kexit);
    if (t < 0) {
      err |= OM_DEV_WALK;
      break;
    }
    const int cn = t >> 2, kn = t & 3;
    const int4 cln = __ldg(p.cells + cn);
    const int jn = slot_of(cln, v);
    if (jn < 0 || jn == kn || ++hops > MAX_RING) {
      err |= OM_DEV_WALK;
      break;
    }
    const int newid = cell_get(cln, kn);
    if (newid == first) {
      ch.close(Pfirst, false);
      break;
    }
    ch.next(ld_point<D>(p.x, newid), false);
    cur = cn;
    kexit = 3 - jn - kn;
  }
  err |= ch.error();
}

// x <- x + omega (target - x), limited: the driver-loop tail for methods whose target
// comes from a solve (cpt-linear-solve, cpt-quasi-newton).
template <int D>
__global__ void __launch_bounds__(128) k_relax_from_target(StepParams p, const double* target) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= p.N) return;
  int err = 0;
  const Vec<D> P0 = ld_point<D>(p.x, v);
  Vec<D> out = P0;
  double diff2 = 0.0;
  bool limited = false;
  const int c0 = p.v2c[v];
  if (c0 != OM_NONE_CELL && !p.bflag[v]) {
    const Vec<D> T = ld_point<D>(target, v);
    Vec<D> d;
#pragma unroll
    for (int k = 0; k < D; k++) d.v[k] = p.omega * (T.v[k] - P0.v[k]);
    diff2 = vdot<D>(d, d);
    if (p.limiter) {
      Chain<D, OM_CHAIN_LIMITER_ONLY, true, false> ch;
      star_limiter<D>(p, v, c0, P0, ch, err);
      if (!err) limited = ch.limit(d, diff2);
    }
#pragma unroll
    for (int k = 0; k < D; k++) out.v[k] = P0.v[k] + d.v[k];
  }
  st_point<D>(p.xout, v, out);
  p.diff2[v] = limited ? -diff2 : diff2;
  if (err) atomicOr(&p.ds->err, err);
}

// ---- synthetic workloads (om_random_walk): every free vertex moves by a random vector of
// length <= amplitude/2 x its smallest incident inradius (no cell can invert: every height
// of a triangle is at least twice its inradius).  The random numbers are a hash of (seed,
// round, caller vertex id): the result does not depend on the internal numbering.
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

__global__ void __launch_bounds__(128)
    k_random_move(StepParams p, const int* __restrict__ perm, unsigned long long seed, int round,
                  double amplitude) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= p.N) return;
  int err = 0;
  const Vec<2> P0 = ld_point<2>(p.x, v);
  Vec<2> out = P0;
  const int c0 = p.v2c[v];
  if (c0 != OM_NONE_CELL && !p.bflag[v]) {
    Chain<2, OM_CHAIN_LIMITER_ONLY, true, false> ch;
    star_limiter<2>(p, v, c0, P0, ch, err);
    if (!err) {
      const unsigned long long id = (unsigned long long)(perm ? perm[v] : v);
      const unsigned long long h0 =
          splitmix64(seed ^ splitmix64(id + ((unsigned long long)round << 40)));
      const unsigned long long h1 = splitmix64(h0);
      const double u1 = (double)(h0 >> 11) * (1.0 / 9007199254740992.0);
      const double u2 = (double)(h1 >> 11) * (1.0 / 9007199254740992.0);
      // uniform in the disk of radius amplitude/2 x r_min
      const double r = 0.5 * amplitude * (ch.rn / ch.rd) * sqrt(u1);
      double sn, cs;
      sincospi(2.0 * u2, &sn, &cs);
      out.v[0] = P0.v[0] + r * cs;
      out.v[1] = P0.v[1] + r * sn;
    }
  }
  st_point<2>(p.xout, v, out);
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

__global__ void k_reset_step_scalars(DevScalars* ds, int gate) {
  if (gate && ds->halt) return;
  ds->n_deferred = 0;
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
  p.diff2 = h->diff2;
  p.vflags = h->vflags;
  p.valid_epoch = (h->own_hi >= 0 && h->valid_epoch && !h->all_valid) ? h->valid_epoch : nullptr;
  p.valid_stamp = h->valid_stamp;
  p.N = (int)h->N;
  p.lo = 0;
  p.hi = (int)h->N;
  p.omega = h->omega;
  p.limiter = h->limiter;
  p.odt_bary = h->odt_bary;
  p.force_walk = h->use_rings ? 0 : 1;
  {
    static const int waves = getenv("OM_K1_PREFETCH_BLOCKS") ? atoi(getenv("OM_K1_PREFETCH_BLOCKS"))
                                                             : 148 * OM_K1_MINB;
    p.prefetch_ahead = waves * OM_K1_BLOCK;
  }
  p.gate = 0;
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

// The point update of the fixed-point methods on [lo, hi) into `out`: ring-row kernel, the
// vertices it left behind, statistics.  check: also collect the suspicious spokes (fused
// Delaunay check of the pipelined loop).
int om_launch_point_update(om_handle* h, double* out, bool check) {
  StepParams p = make_params(h, out);
  if (h->own_hi >= 0) {
    p.lo = (int)h->own_lo;
    p.hi = (int)h->own_hi;
  }
  // the lazy limiter pays off once few vertices are limited (the previous step tells; see
  // k_pl_iter_end in loop.cu for the break-even)
  const bool exact = om_limiter_mode(h->limiter != 0, (long long)(h->limited_frac * 1.0e6),
                                     1000000, om_lim_div()) == 1;
  if (h->timing) cudaEventRecord(h->ev[0], h->stream);
  OM_TRY(launch_step(h, p, check ? 1 : 0, exact));
  if (h->timing) {
    cudaEventRecord(h->ev[1], h->stream);
    h->ev_pending = true;
  }
  OM_TRY(launch_step(h, p, 2));
  return OM_OK;
}

int om_launch_reduce_stats(om_handle* h) {
  int lo = h->own_hi >= 0 ? (int)h->own_lo : 0;
  int hi = h->own_hi >= 0 ? (int)h->own_hi : (int)h->N;
  om_shared_vertex_range(h, &lo, &hi);
  if (hi > lo)
    OM_LAUNCH(h, k_reduce_stats, std::min(om_grid(hi - lo, 256 * 8), 148 * 8), 256, h->diff2, lo, hi,
              h->ds, h->sh ? 1 : 0);
  CUDA_TRY(cudaGetLastError());
  return OM_OK;
}

// recomputes the vertices of h->dirty (stars changed by flips) from h->x into `out`
int om_launch_fixup(om_handle* h, double* out) {
  StepParams p = make_params(h, out);
  return launch_step(h, p, 4);
}

// ---- launchers of the pipelined loop (loop.cu).  Every kernel is gated: it returns at once
// when the loop has halted, and of the two limiter variants of the ring kernel only the one
// the device selected (ds->mode_exact) does the work -- the host enqueues both.
int om_pl_launch_update(om_handle* h, const double* xin, double* xout, bool timed) {
  StepParams p = make_params(h, xout);
  p.x = xin;
  p.gate = 1;
  OM_LAUNCH(h, k_reset_step_scalars, 1, 1, h->ds, 1);
  if (timed) cudaEventRecord(h->ev[0], h->stream);
  OM_TRY(launch_step(h, p, 1, false));
  OM_TRY(launch_step(h, p, 1, true));
  if (timed) cudaEventRecord(h->ev[1], h->stream);
  OM_TRY(launch_step(h, p, 2));
  return OM_OK;
}

// the same in three pieces, for a caller that runs only the variant the device selected
// (conditional graph nodes): what = 0 reset, 1 lazy ring kernel, 2 exact ring kernel, 3 k_post
int om_pl_launch_update_part(om_handle* h, const double* xin, double* xout, int what) {
  StepParams p = make_params(h, xout);
  p.x = xin;
  p.gate = 1;
  om_shared_vertex_range(h, &p.lo, &p.hi);  // shared address space: this rank's vertices
  if (what == 0) {
    OM_LAUNCH(h, k_reset_step_scalars, 1, 1, h->ds, 1);
    CUDA_TRY(cudaGetLastError());
    return OM_OK;
  }
  if (what == 3) return launch_step(h, p, 2);
  return launch_step(h, p, 1, what == 2);
}

__global__ void k_pl_count(DevScalars* ds, int n) {
  if (!ds->halt) ds->pl_launches += n;
}

// after the flip pass: ring rows of the vertices whose star changed, their update recomputed
// from xin on the new topology, statistics of the whole update
int om_pl_launch_tail(om_handle* h, const double* xin, double* xout) {
  OM_LAUNCH(h, (k_build_rings<true>), 148 * 4, 256, h->cells, (const int*)h->adj, h->v2c, h->bflag,
            0, h->dirty, (const int*)&h->ds->n_dirty, h->ring, h->ringc, 0, (int)h->N,
            (const int*)&h->ds->halt);
  StepParams p = make_params(h, xout);
  p.x = xin;
  p.gate = 1;
  OM_TRY(launch_step(h, p, 4));
  if (h->sh) return OM_OK;  // shared address space: the caller synchronises the GPUs first
  if (h->N > 0)
    OM_LAUNCH(h, k_reduce_stats, std::min(om_grid(h->N, 256 * 8), 148 * 8), 256, h->diff2, 0,
              (int)h->N, h->ds, 1);
  // reset, 2 ring kernels, post, pass begin, flag check, select, flip1, flip2, round end,
  // rings, fix-up, reduce, this one, iteration end
  OM_LAUNCH(h, k_pl_count, 1, 1, h->ds, 9);
  CUDA_TRY(cudaGetLastError());
  return OM_OK;
}

int om_update_points_impl(om_handle* h, double tol, om_step_stats* out, bool target_only,
                          double* target_out, bool defer_fetch) {
  if (h->N == 0) return OM_OK;
  if (!target_only) h->delaunay_clean = false;
  OM_LAUNCH(h, k_reset_step_scalars, 1, 1, h->ds, 0);
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
      const int B = 128, G = om_grid(h->N, B);
      if (h->D == 2)
        OM_LAUNCH(h, k_relax_from_target<2>, G, B, p, sol);
      else
        OM_LAUNCH(h, k_relax_from_target<3>, G, B, p, sol);
      CUDA_TRY(cudaGetLastError());
      CUDA_TRY(cudaMemcpyAsync(h->xnew, tmp, sizeof(double) * h->N * h->PD,
                               cudaMemcpyDeviceToDevice, h->stream));
      om_free(h, tmp);  // stream ordered: released after the copy
      OM_TRY(om_launch_reduce_stats(h));
    }
  } else if (target_only) {
    StepParams p = make_params(h, target_out);
    OM_TRY(launch_step(h, p, 3));
  } else {
    OM_TRY(om_launch_point_update(h, h->xnew, false));
    OM_TRY(om_launch_reduce_stats(h));
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

// ---- the update from caller-supplied targets (device side of the boundary_step hook)
namespace {
// per vertex: smallest inradius over its cells, as the bits of a non-negative double
template <int D>
__global__ void __launch_bounds__(256)
    k_cell_inradius_min(const double* __restrict__ x, const int4* __restrict__ cells, int C,
                        unsigned long long* __restrict__ minr) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int4 cl = __ldg(cells + c);
  const Vec<D> P0 = ld_point<D>(x, cl.x), P1 = ld_point<D>(x, cl.y), P2 = ld_point<D>(x, cl.z);
  const Vec<D> e0 = vsub<D>(P2, P1), e1 = vsub<D>(P0, P2), e2 = vsub<D>(P1, P0);
  const double l0 = sqrt(vdot<D>(e0, e0)), l1 = sqrt(vdot<D>(e1, e1)), l2 = sqrt(vdot<D>(e2, e2));
  const double d12 = vdot<D>(e1, e2);
  // |e1 x e2|^2 = |e1|^2 |e2|^2 - (e1.e2)^2 in any embedding dimension
  const double area = 0.5 * sqrt(fmax(l1 * l1 * l2 * l2 - d12 * d12, 0.0));
  const double r = 2.0 * area / (l0 + l1 + l2);
  const unsigned long long b = (unsigned long long)__double_as_longlong(r);
  atomicMin(minr + cl.x, b);  // (integer min of the bit patterns: order independent)
  atomicMin(minr + cl.y, b);
  atomicMin(minr + cl.z, b);
}

// x <- x + omega (target - x), limited to half the smallest incident inradius, for EVERY
// vertex with a target (boundary vertices included: their targets come from the caller)
template <int D>
__global__ void __launch_bounds__(256)
    k_relax_all(StepParams p, const double* __restrict__ target,
                const unsigned long long* __restrict__ minr) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= p.N) return;
  const Vec<D> P0 = ld_point<D>(p.x, v), T = ld_point<D>(target, v);
  Vec<D> d;
#pragma unroll
  for (int k = 0; k < D; k++) d.v[k] = p.omega * (T.v[k] - P0.v[k]);
  const double diff2 = vdot<D>(d, d);
  bool limited = false;
  if (p.limiter) {
    const double limit = 0.5 * __longlong_as_double((long long)minr[v]);
    const double length = sqrt(diff2);
    if (length > limit) {
      const double f = limit / length;
#pragma unroll
      for (int k = 0; k < D; k++) d.v[k] *= f;
      limited = true;
    }
  }
  Vec<D> out;
#pragma unroll
  for (int k = 0; k < D; k++) out.v[k] = P0.v[k] + d.v[k];
  st_point<D>(p.xout, v, out);
  p.diff2[v] = limited ? -diff2 : diff2;
}
}  // namespace

int om_update_from_targets_impl(om_handle* h, const double* targets_dev, double tol,
                                om_step_stats* out) {
  if (h->N == 0) return OM_OK;
  h->delaunay_clean = false;
  OM_LAUNCH(h, k_reset_step_scalars, 1, 1, h->ds, 0);
  unsigned long long* minr = nullptr;
  CUDA_TRY(om_malloc(h, &minr, sizeof(unsigned long long) * h->N));
  // +inf for vertices without a cell (0x7ff0... as four identical 16-bit halves is not
  // possible: fill with the byte 0x7f -> 0x7f7f..., a huge finite double)
  CUDA_TRY(cudaMemsetAsync(minr, 0x7f, sizeof(unsigned long long) * h->N, h->stream));
  StepParams p = make_params(h, h->xnew);
  if (h->C > 0) {
    if (h->D == 2)
      OM_LAUNCH(h, k_cell_inradius_min<2>, om_grid(h->C, 256), 256, h->x, h->cells, (int)h->C,
                minr);
    else
      OM_LAUNCH(h, k_cell_inradius_min<3>, om_grid(h->C, 256), 256, h->x, h->cells, (int)h->C,
                minr);
  }
  if (h->D == 2)
    OM_LAUNCH(h, k_relax_all<2>, om_grid(h->N, 256), 256, p, targets_dev, minr);
  else
    OM_LAUNCH(h, k_relax_all<3>, om_grid(h->N, 256), 256, p, targets_dev, minr);
  CUDA_TRY(cudaGetLastError());
  om_free(h, minr);
  OM_LAUNCH(h, k_reduce_stats, std::min(om_grid(h->N, 256 * 8), 148 * 8), 256, h->diff2, 0,
            (int)h->N, h->ds, 0);
  OM_TRY(om_fetch_scalars(h));
  OM_TRY(om_check_dev_err(h));
  std::swap(h->x, h->xnew);
  om_step_stats_from_scalars(h, tol, out);
  return OM_OK;
}

int om_random_move_impl(om_handle* h, uint64_t seed, int round, double amplitude) {
  if (h->N == 0) return OM_OK;
  if (h->D != 2) {
    om_set_error("om_random_walk is for flat (2D) meshes");
    return OM_ERR_ARG;
  }
  StepParams p = make_params(h, h->xnew);
  OM_LAUNCH(h, k_random_move, om_grid(h->N, 128), 128, p, (const int*)h->perm,
            (unsigned long long)seed, round, amplitude);
  CUDA_TRY(cudaGetLastError());
  std::swap(h->x, h->xnew);
  h->delaunay_clean = false;
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

// (Re)builds ring rows: all vertices, or the vertices touched by flips since the last call
// (device == true: the list length is read on the device, no host copy of it is needed).
int om_rebuild_rings(om_handle* h, bool all, bool device) {
  if (!h->ring || h->N == 0) return OM_OK;
  const int B = 256;
  if (all) {
    OM_LAUNCH(h, (k_build_rings<false>), om_grid(h->N, B), B, h->cells, (const int*)h->adj, h->v2c,
              h->bflag, (int)h->N, (const int*)nullptr, (const int*)nullptr, h->ring, h->ringc, 0,
              (int)h->N, (const int*)nullptr);
    h->rings_partial = false;
  } else {
    // with an owned range only its rows are kept current (nothing else reads ring rows);
    // om_set_owned_range(whole mesh) rebuilds all of them
    const bool ranged = h->own_hi >= 0;
    if (ranged) h->rings_partial = true;
    const int lo = ranged ? (int)h->own_lo : 0, hi = ranged ? (int)h->own_hi : (int)h->N;
    if (device) {
      OM_LAUNCH(h, (k_build_rings<true>), 148 * 4, B, h->cells, (const int*)h->adj, h->v2c, h->bflag,
                0, h->dirty, (const int*)&h->ds->n_dirty, h->ring, h->ringc, lo, hi,
                (const int*)nullptr);
    } else {
      const int n = h->hs->n_dirty;  // fetched by the flip pass
      if (n > 0)
        OM_LAUNCH(h, (k_build_rings<true>), std::min(om_grid(n, B), 148 * 8), B, h->cells,
                  (const int*)h->adj, h->v2c, h->bflag, n, h->dirty, (const int*)nullptr, h->ring,
                  h->ringc, lo, hi, (const int*)nullptr);
    }
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
  if (e[0] >= 0) {
    const int k = 1 + ((e[0] >> 30) & 1) + ((e[1] >> 29) & 2) + ((e[2] >> 28) & 4);
#pragma unroll
    for (int q = 0; q < OM_RING_W; q++)
      if (q < k) f(e[q] & RING_MASK);
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
  h->delaunay_clean = false;
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
