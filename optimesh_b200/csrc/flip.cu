// K2 + K3: GPU flip-until-Delaunay.
//
// Replaces meshplex's MeshTri.flip_until_delaunay() called by the optimize() loop after
// every step (/root/reference/README.md:131-132).  Per round:
//   k_suspect one thread per cell: an edge can only violate the Delaunay criterion
//             s = ce_k(c) + ce_k'(c') < -tol (ce_k = -ed_k / (4A)) if one of its two opposite
//             angles is obtuse, and a triangle has at most one obtuse angle -- so each cell
//             examines at most one edge, and only then touches its neighbour (whose term is
//             recomputed from the opposite vertex; no per-cell intermediate array).
//   k_flip1   candidates only: every flagged cell keeps its most negative flagged edge (ties:
//             lowest local index); an edge kept by BOTH adjacent cells is flipped (an
//             independent set: every cell takes part in at most one flip); rewrites the two
//             cells and records where the four outer half-edges move.
//   k_flip2   candidates only: patches the twin table from the relocation records
//             (race-free when neighbouring cells flip in the same round) and builds the
//             work list of the next round: flipped cells, their outer neighbours and the
//             candidates that lost -- the only cells whose edges can still be flagged.
// Round 1 checks every cell, later rounds only the work list, until nothing is flagged.
// (a0,k0) of a flip is the half-edge with the smaller id 3*row+k in the caller's cell
// numbering, so the cell array is identical to the oracle's, row for row.  All dot
// products are explicit fma chains: both cells of an edge see bit-identical s.
#include <algorithm>

#include <cstdlib>

#include "common.cuh"
#include "geom.cuh"

namespace {

template <int D>
__device__ __forceinline__ double fdot(const Vec<D>& a, const Vec<D>& b) {
  double s = __dmul_rn(a.v[0], b.v[0]);
#pragma unroll
  for (int k = 1; k < D; k++) s = __fma_rn(a.v[k], b.v[k], s);
  return s;
}

// ed_k = e_{k+1} . e_{k+2} of a cell with vertices Q[0..2] (slot order)
template <int D>
__device__ __forceinline__ void cell_ed(const Vec<D> (&Q)[3], double (&ed)[3]) {
  const Vec<D> e0 = vsub<D>(Q[2], Q[1]), e1 = vsub<D>(Q[0], Q[2]), e2 = vsub<D>(Q[1], Q[0]);
  ed[0] = fdot<D>(e1, e2);
  ed[1] = fdot<D>(e2, e0);
  ed[2] = fdot<D>(e0, e1);
}
__device__ __forceinline__ double vol2_of(const double (&ed)[3]) {
  // 0.25 (ed2 ed0 + ed0 ed1 + ed1 ed2)
  double s = __dmul_rn(ed[2], ed[0]);
  s = __fma_rn(ed[0], ed[1], s);
  s = __fma_rn(ed[1], ed[2], s);
  return 0.25 * s;
}
__device__ __forceinline__ double sel3(const double (&a)[3], int k) {
  return k == 0 ? a[0] : (k == 1 ? a[1] : a[2]);
}
__device__ __forceinline__ int sel3i(const int (&a)[3], int k) {
  return k == 0 ? a[0] : (k == 1 ? a[1] : a[2]);
}

// s = ce_k(c) + ce_k'(c') (A.7) from the two cells' own ed and squared areas: the one place
// where s is evaluated, so that every kernel sees the same bits for the same edge
__device__ __forceinline__ double delaunay_s(double edk, double vol2, double ednk, double vol2n) {
  const double inv4A = 0.25 / sqrt(vol2), inv4An = 0.25 / sqrt(vol2n);
  return __dadd_rn(__dmul_rn(-edk, inv4A), __dmul_rn(-ednk, inv4An));
}

// One thread per cell (all cells, or the work list).  A triangle has at most one obtuse
// angle, and an edge can only violate the Delaunay criterion if one of its two opposite
// angles is obtuse (ed > 0): so every cell examines at most ONE edge -- the one opposite
// its own obtuse angle -- and the cheap path touches nothing but the cell's own vertices.
// A flagged edge writes s into the slots of both half-edges and enlists both cells.
// MODE 0: cells [off, off+n); MODE 1: cells list[0..n); MODE 2: cells [off, off+n), flagged
// edges are appended to `recs` instead of being applied (sharded check: the records of all
// ranks are exchanged and applied by k_apply_records).
// Which coordinates a rank may trust when the coordinates are partitioned (dist.py): its own
// vertex range, pinned vertices (they never move) and the vertices refreshed by the last
// band exchange.  valid_epoch == nullptr: everything is valid (single GPU, or right after a
// full all-gather).
struct ShardInfo {
  int vlo, vhi;   // own vertex range
  int clo, chi;   // cell range this rank examines (MODE 3 filters the work list with it)
  const int* valid_epoch;
  int valid_stamp;
  const uint8_t* bflag;
  __device__ __forceinline__ bool valid(int v) const {
    return valid_epoch == nullptr || (v >= vlo && v < vhi) || bflag[v] != 0 ||
           valid_epoch[v] == valid_stamp;
  }
};

// MODE 3: like MODE 2 (records) but over the work list, restricted to cells in [clo, chi).
template <int D, int MODE>
__global__ void __launch_bounds__(256, (D == 2 ? 5 : 4))  // 2D: <= 51 registers, 5 blocks per SM
    k_suspect(const double* __restrict__ x, const int4* __restrict__ cells,
              const int* __restrict__ adj, int off, int n, const int* __restrict__ list,
              double tol, double* __restrict__ sarr, int* __restrict__ cand,
              int* __restrict__ cand_epoch, FlipRec* __restrict__ recs,
              DevScalars* ds, const int* __restrict__ n_dev, ShardInfo sh) {
  if (n_dev) n = *n_dev;  // length known only on the device (chained rounds)
  const int epoch = ds->epoch;
  // list modes: block-stride loop (chained rounds run on a fixed grid whatever the list length
  // is); range modes: exactly one trip per block (the loop form cost the range check 0.1 ms)
  constexpr bool LOOP = MODE == 1 || MODE == 3;
  for (int base = blockIdx.x * blockDim.x;
       LOOP ? base < n : base == (int)(blockIdx.x * blockDim.x);
       base += gridDim.x * blockDim.x) {
  const int i = base + threadIdx.x;
  bool flag = false;
  int c = -1, cn = -1, he = -1, tt = -1;
  double sval = 0.0;
  do {  // no early exit: the list appends below are block-collective
    if (i >= n) break;
    c = (MODE == 1 || MODE == 3) ? list[i] : off + i;
    if (MODE == 3 && (c < sh.clo || c >= sh.chi)) break;  // another rank examines it
    const int4 cl = cells[c];
    if ((MODE == 2 || MODE == 3) && !(sh.valid(cl.x) && sh.valid(cl.y) && sh.valid(cl.z))) {
      ds->stale = 1;  // a coordinate this rank does not hold: the caller refreshes and repeats
      break;
    }
    // fetched up front (coalesced) so the slow path does not wait for it after the geometry;
    // loading the twin only for suspects was measured slower (0.577 vs 0.551 ms per pass)
    const int4 tw = __ldg(reinterpret_cast<const int4*>(adj) + c);
    Vec<D> P[3] = {ld_point<D>(x, cl.x), ld_point<D>(x, cl.y), ld_point<D>(x, cl.z)};
    double ed[3];
    cell_ed<D>(P, ed);
    const double vol2 = vol2_of(ed);
    if (!(vol2 > 0.0)) {
      atomicOr(&ds->err, OM_DEV_DEGENERATE);
      break;
    }
    const int k = ed[0] > 0.0 ? 0 : (ed[1] > 0.0 ? 1 : (ed[2] > 0.0 ? 2 : -1));
    if (k < 0) break;  // all angles <= 90 deg: nothing to examine from this side
    const int t = k == 0 ? tw.x : (k == 1 ? tw.y : tw.z);
    if (t < 0) break;  // boundary edge
    cn = t >> 2;
    const int kn = t & 3;
    const int4 cln = __ldg(cells + cn);
    const int nid[3] = {cln.x, cln.y, cln.z};
    if ((MODE == 2 || MODE == 3) && !sh.valid(sel3i(nid, kn))) {
      ds->stale = 1;
      break;
    }
    // neighbour's vertices in ITS slot order: slot kn is the opposite vertex, the other
    // two are shared with this cell (slots (k+1)%3 and (k+2)%3 here)
    const Vec<D> O = ld_point<D>(x, sel3i(nid, kn));
    const int vid[3] = {cl.x, cl.y, cl.z};
    const int ia = sel3i(vid, (k + 1) % 3);
    const Vec<D> Pa = k == 0 ? P[1] : (k == 1 ? P[2] : P[0]);
    const Vec<D> Pb = k == 0 ? P[2] : (k == 1 ? P[0] : P[1]);
    Vec<D> Q[3];
#pragma unroll
    for (int s = 0; s < 3; s++) Q[s] = (s == kn) ? O : (nid[s] == ia ? Pa : Pb);
    double edn[3];
    cell_ed<D>(Q, edn);
    const double vol2n = vol2_of(edn);
    if (!(vol2n > 0.0)) break;  // reported by the neighbour itself
    // Cheap rejection before any sqrt/division.  With ea = ed_k > 0 (obtuse here) and
    // eb = ed_k' < 0 (acute there): s < 0  <=>  ea / A > -eb / A'  <=>  ea^2 A'^2 > eb^2 A^2.
    // Edges within 1e-9 of equality still go through the exact evaluation below, so every
    // flag decision is taken on the exact s.
    {
      const double ea = sel3(ed, k), eb = sel3(edn, kn);
      if (tol >= 0.0 && eb < 0.0 && !(ea * ea * vol2n > eb * eb * vol2 * (1.0 - 1e-9))) break;
    }
    const double s = delaunay_s(sel3(ed, k), vol2, sel3(edn, kn), vol2n);
    if (s < -tol) {
      he = 4 * c + k;
      tt = t;
      sval = s;
      flag = true;
      if (MODE != 2 && MODE != 3) {
        sarr[he] = s;
        sarr[t] = s;
      }
    }
  } while (false);
  if (MODE == 2 || MODE == 3) {
    // records: reserve one slot per flagged thread (block-aggregated), then write
    __shared__ int r_warp[8];
    __shared__ int r_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned m = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) r_warp[warp] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); w++) {
        const int t2 = r_warp[w];
        r_warp[w] = tot;
        tot += t2;
      }
      r_base = tot ? atomicAdd(&ds->n_rec, tot) : 0;
    }
    __syncthreads();
    if (flag) {
      FlipRec r;
      r.he = he;
      r.twin = tt;
      r.s = sval;
      recs[r_base + r_warp[warp] + __popc(m & ((1u << lane) - 1u))] = r;
    }
    if (!LOOP) return;
    __syncthreads();  // r_warp / r_base are reused by the next trip
    continue;
  }
  // enlist both cells once (stamp dedupes; one atomic per block on the shared counter)
  const int vals[2] = {c, cn};
  const bool preds[2] = {flag && atomicExch(&cand_epoch[c], epoch) != epoch,
                         flag && atomicExch(&cand_epoch[cn], epoch) != epoch};
  block_append<2>(&ds->n_cand, cand, vals, preds);
  }
}

// flagged-edge records (own or received from other ranks) -> s slots + candidate list
__global__ void __launch_bounds__(256)
    k_apply_records(const FlipRec* __restrict__ recs, int n, double* __restrict__ sarr,
                    int* __restrict__ cand, int* __restrict__ cand_epoch, DevScalars* ds) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int epoch = ds->epoch;
  int c = -1, cn = -1;
  if (i < n) {
    const FlipRec r = recs[i];
    sarr[r.he] = r.s;
    sarr[r.twin] = r.s;
    c = r.he >> 2;
    cn = r.twin >> 2;
  }
  const int vals[2] = {c, cn};
  const bool preds[2] = {c >= 0 && atomicExch(&cand_epoch[c], epoch) != epoch,
                         cn >= 0 && atomicExch(&cand_epoch[cn], epoch) != epoch};
  block_append<2>(&ds->n_cand, cand, vals, preds);
}

// ---- fixed-capacity record exchange (no host readback between check and flips).
// Every rank publishes one slot of 1 + cap records: slot[0] = {count, stale flag}, then its
// records.  After the all-gather k_round_scan looks at the P headers: a stale rank or a count
// above cap sets ds->abort, which turns the rest of the round (apply, select, flips) into
// no-ops; the host then repeats the round the slow way.
__global__ void k_round_pack(const FlipRec* __restrict__ recs, const DevScalars* ds, int cap,
                             FlipRec* __restrict__ slot) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = ds->n_rec;
  if (i == 0) {
    FlipRec hd;
    hd.he = n;
    hd.twin = ds->stale;
    hd.s = 0.0;
    slot[0] = hd;
  }
  if (i < n && i < cap) slot[1 + i] = recs[i];
}

__global__ void k_round_scan(const FlipRec* __restrict__ gathered, int P, int cap,
                             DevScalars* ds) {
  int ab = 0, mx = 0;
  for (int r = 0; r < P; r++) {
    const FlipRec hd = gathered[(size_t)r * (cap + 1)];
    if (hd.twin) ab |= 1;
    if (hd.he > cap) ab |= 2;
    mx = max(mx, hd.he);
  }
  ds->abort = ab;
  ds->max_count = mx;  // the same on every rank: lets them agree on the next capacity
}

__global__ void __launch_bounds__(256)
    k_apply_gathered(const FlipRec* __restrict__ gathered, int P, int cap,
                     double* __restrict__ sarr, int* __restrict__ cand,
                     int* __restrict__ cand_epoch, DevScalars* ds) {
  if (ds->abort) return;
  const int epoch = ds->epoch;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  int c = -1, cn = -1;
  if (i < (long long)P * cap) {
    const int r = (int)(i / cap), j = (int)(i % cap);
    const FlipRec* slot = gathered + (size_t)r * (cap + 1);
    if (j < slot[0].he) {
      const FlipRec rec = slot[1 + j];
      sarr[rec.he] = rec.s;
      sarr[rec.twin] = rec.s;
      c = rec.he >> 2;
      cn = rec.twin >> 2;
    }
  }
  const int vals[2] = {c, cn};
  const bool preds[2] = {c >= 0 && atomicExch(&cand_epoch[c], epoch) != epoch,
                         cn >= 0 && atomicExch(&cand_epoch[cn], epoch) != epoch};
  block_append<2>(&ds->n_cand, cand, vals, preds);
}

// the choice of a cell: its most negative flagged edge (ties: lowest local index), -1 if none.
// Evaluated where it is needed (k_flip1, for the candidate and for the cell across its edge)
// instead of in a kernel of its own: one launch -- and, with several GPUs, one meeting --
// less per round.  The s slots are cleared by k_flip2, when nobody reads them any more.
__device__ __forceinline__ int best_edge(const double* __restrict__ sarr, int c) {
  const double2* p = reinterpret_cast<const double2*>(sarr + 4 * (size_t)c);
  const double2 s01 = p[0], s2 = p[1];
  const double sv[3] = {s01.x, s01.y, s2.x};
  int b = -1;
  double sb = 0.0;
#pragma unroll
  for (int k = 0; k < 3; k++)
    if (sv[k] < INFINITY && (b < 0 || sv[k] < sb)) {
      b = k;
      sb = sv[k];
    }
  return b;
}

__global__ void __launch_bounds__(256)
    k_flip1(int4* __restrict__ cells, const int4* __restrict__ adj, const double* __restrict__ sarr,
            const int* __restrict__ cand, int n, int* __restrict__ flip_epoch,
            int* __restrict__ reloc, int4* __restrict__ adj_tmp, int* __restrict__ v2c,
            int* __restrict__ dirty, int* __restrict__ dirty_epoch,
            DevScalars* ds, const int* __restrict__ n_dev, int vlo, int vhi) {
  if (ds->abort) return;
  if (n_dev) n = *n_dev;
  const int epoch = ds->epoch, dirty_pass = ds->dirty_pass;
  // the work list was consumed by the check of this round
  if (blockIdx.x == 0 && threadIdx.x == 0) ds->n_work = 0;
  for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
  const int i = base + threadIdx.x;
  int nf = 0;
  int dv[4] = {0, 0, 0, 0};  // the four vertices of the flip this thread applied
  if (i < n) {
    const int a0 = cand[i];
    const int k0 = best_edge(sarr, a0);  // a candidate has at least one flagged edge
    const int4 adjA = adj[a0];
    const int t = k0 >= 0 ? cell_get(adjA, k0) : -1;
    const int a1 = t >> 2, k1 = t & 3;
    if (t >= 0 && best_edge(sarr, a1) == k1) {
      const int4 A = cells[a0];
      const int4 Bc = cells[a1];
      // the half-edge with the smaller caller-numbering id owns the flip
      const long long hA = 3ll * A.w + k0, hB = 3ll * Bc.w + k1;
      if (hA < hB) {
        const int4 adjB = adj[a1];
        const int v0 = cell_get(A, k0), v2 = cell_get(A, (k0 + 1) % 3),
                  v3 = cell_get(A, (k0 + 2) % 3), v1 = cell_get(Bc, k1);
        const int s2 = slot_of(Bc, v2), s3 = slot_of(Bc, v3);
        if (s2 < 0 || s3 < 0) {
          atomicOr(&ds->err, OM_DEV_NONMANIFOLD);
        } else {
          const int tA1 = cell_get(adjA, (k0 + 1) % 3);  // across (v3,v0)
          const int tA2 = cell_get(adjA, (k0 + 2) % 3);  // across (v0,v2)
          const int tB2 = cell_get(adjB, s2);            // across (v1,v3)
          const int tB3 = cell_get(adjB, s3);            // across (v1,v2)
          cells[a0] = make_int4(v0, v1, v2, A.w);
          cells[a1] = make_int4(v0, v1, v3, Bc.w);
          // new outer edges, still naming the OLD twins; slot 2 is the shared new edge
          adj_tmp[a0] = make_int4(tB3, tA2, 4 * a1 + 2, 0);
          adj_tmp[a1] = make_int4(tB2, tA1, 4 * a0 + 2, 0);
          reloc[4 * a0 + (k0 + 1) % 3] = 4 * a1 + 1;
          reloc[4 * a0 + (k0 + 2) % 3] = 4 * a0 + 1;
          reloc[4 * a1 + s2] = 4 * a1 + 0;
          reloc[4 * a1 + s3] = 4 * a0 + 0;
          flip_epoch[a0] = epoch;
          flip_epoch[a1] = epoch;
          // v2 lost a1, v3 lost a0 (at most one flip per round can own v2c[v])
          if (v2c[v2] == a1) v2c[v2] = a0;
          if (v2c[v3] == a0) v2c[v3] = a1;
          // the four stars changed: their ring rows are rebuilt after the pass
          dv[0] = v0;
          dv[1] = v1;
          dv[2] = v2;
          dv[3] = v3;
          nf = 1;
        }
      }
    }
  }
  if (dirty) {
    bool dp[4];
#pragma unroll
    for (int q = 0; q < 4; q++)
      dp[q] = nf && dv[q] >= vlo && dv[q] < vhi &&
              atomicExch(&dirty_epoch[dv[q]], dirty_pass) != dirty_pass;
    block_append<4>(&ds->n_dirty, dirty, dv, dp);
  }
  for (int o = 16; o > 0; o >>= 1) nf += __shfl_xor_sync(0xffffffffu, nf, o);
  if ((threadIdx.x & 31) == 0 && nf) atomicAdd(&ds->n_flips, nf);
  }
}

__global__ void __launch_bounds__(256)
    k_flip2(int* __restrict__ adj, const int4* __restrict__ adj_tmp,
            const int* __restrict__ flip_epoch, const int* __restrict__ reloc,
            const int* __restrict__ cand, int n, int* __restrict__ work_epoch,
            int* __restrict__ work, double* __restrict__ sarr, DevScalars* ds,
            const int* __restrict__ n_dev, int clo, int chi) {
  if (ds->abort) return;
  if (n_dev) n = *n_dev;
  const int epoch = ds->epoch;
  for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
  const int i = base + threadIdx.x;
  // cells to enlist for the next round: self (flipped or lost), the two outer neighbours
  // and the flip partner
  int add[4] = {-1, -1, -1, -1};
  if (i < n) {
    const int c = cand[i];
    // k_flip1 is done with the s slots; non-candidates must read "no flagged edge"
    double2* p = reinterpret_cast<double2*>(sarr + 4 * (size_t)c);
    p[0] = make_double2(INFINITY, INFINITY);
    p[1] = make_double2(INFINITY, INFINITY);
    add[0] = c;
    if (flip_epoch[c] == epoch) {
      const int4 t = adj_tmp[c];
      const int told[2] = {t.x, t.y};
#pragma unroll
      for (int s = 0; s < 2; s++) {
        const int to = told[s];
        int tn = to;
        if (to >= 0) {
          if (flip_epoch[to >> 2] == epoch)
            tn = reloc[to];  // the neighbour flipped too: it patches its own side
          else
            adj[to] = 4 * c + s;
          add[1 + s] = tn >> 2;
        }
        adj[4 * c + s] = tn;
      }
      adj[4 * c + 2] = t.z;
      add[3] = t.z >> 2;
    }
  }
  bool preds[4];
#pragma unroll
  for (int q = 0; q < 4; q++)
    preds[q] = add[q] >= clo && add[q] < chi &&
               atomicExch(&work_epoch[add[q]], epoch) != epoch;
  block_append<4>(&ds->n_work, work, add, preds);
  }
}

// Starts a check round: new stamp for the cell lists; new_pass also starts a flip pass (clears
// the per-pass counters, new stamp for the dirty-vertex list).
__global__ void k_reset_flip_scalars(DevScalars* ds, int new_pass) {
  ds->epoch++;
  ds->n_flagged = 0;
  ds->stale = 0;
  ds->abort = 0;
  ds->n_cand = 0;
  ds->n_rec = 0;
  if (new_pass) {
    ds->dirty_pass++;
    ds->n_dirty = 0;
    ds->n_flips = 0;
    ds->n_work = 0;
    ds->n_rounds = 0;
    ds->flips_prev = 0;
    ds->not_delaunay = 0;
  } else if (ds->n_flips > ds->flips_prev) {
    // the round that just ended flipped something (the reference counts only those)
    ds->n_rounds++;
    ds->flips_prev = ds->n_flips;
  }
}

// ---- fused check of the pipelined loop (loop.cu).  The step kernel evaluates, for every
// spoke of every vertex, twice the Delaunay indicator of that edge on the way (chain.cuh) and
// leaves "spoke q may violate the criterion" bits in the vertex's flag word (bit 9: check all
// my spokes -- vertices the ring kernel does not evaluate itself).  This kernel scans the flag
// words warp by warp (runs of 256 words, compacted with a shuffle scan: no block barrier),
// names the half-edge of each flagged spoke from the vertex's ring rows (neighbour ids + the
// cells between them: no star walk) and takes the decision on the exact s, like k_suspect.
// Flagged cells go to the candidate list through a small per-warp list in shared memory: one
// atomic on the global counter per warp and run.
constexpr unsigned VF_SPOKES = 0xffu, VF_DEFER = 0x100u, VF_CHECKALL = 0x200u;
// FLG_PER flag words per lane and trip (a multiple of 8: 16-byte loads).  The kernel is a chain
// of dependent gathers with a handful of flagged vertices per 256 words.  Measured: 16 words per
// lane (half as many trips, same number of resident warps) is SLOWER, 0.113 vs 0.081 ms per
// step -- the slowest lane of a trip (a star walk) holds its warp, smaller trips balance better.
#ifndef OM_FLG_PER
#define OM_FLG_PER 8
#endif
constexpr int FLG_BLOCK = 256, FLG_PER = OM_FLG_PER, FLG_RUN = 32 * FLG_PER,
              FLG_OUT = 24 * FLG_PER, FLG_HE = 32 * FLG_PER;
static_assert(FLG_PER % 8 == 0, "flag words are read 8 at a time");
constexpr size_t FLG_SMEM =
    (FLG_BLOCK / 32) * (sizeof(int) * (FLG_RUN + FLG_HE + FLG_OUT) + sizeof(unsigned short) * FLG_RUN);
constexpr int RING_IDMASK = (1 << 29) - 1;

template <int D>
__device__ __forceinline__ void check_half_edge(const double* __restrict__ x,
                                                const int4* __restrict__ cells,
                                                const int* __restrict__ adj, int c, int k,
                                                const int4& cl, double tol,
                                                double* __restrict__ sarr,
                                                int* __restrict__ cand_epoch, int epoch,
                                                int* s_out, int* s_nout, int* __restrict__ cand,
                                                DevScalars* ds) {
  const int t = __ldg(adj + 4 * (size_t)c + k);
  if (t < 0) return;  // boundary edge
  const int cn = t >> 2, kn = t & 3;
  const int4 cln = __ldg(cells + cn);
  const Vec<D> P[3] = {ld_point<D>(x, cl.x), ld_point<D>(x, cl.y), ld_point<D>(x, cl.z)};
  const Vec<D> Q[3] = {ld_point<D>(x, cln.x), ld_point<D>(x, cln.y), ld_point<D>(x, cln.z)};
  double ed[3], edn[3];
  cell_ed<D>(P, ed);
  cell_ed<D>(Q, edn);
  const double vol2 = vol2_of(ed), vol2n = vol2_of(edn);
  if (!(vol2 > 0.0) || !(vol2n > 0.0)) {
    atomicOr(&ds->err, OM_DEV_DEGENERATE);
    return;
  }
  const double s = delaunay_s(sel3(ed, k), vol2, sel3(edn, kn), vol2n);
  if (!(s < -tol)) return;
  sarr[4 * (size_t)c + k] = s;
  sarr[t] = s;
  const int ids[2] = {c, cn};
#pragma unroll
  for (int i = 0; i < 2; i++)
    if (atomicExch(&cand_epoch[ids[i]], epoch) != epoch) {
      const int pos = atomicAdd(s_nout, 1);
      if (pos < FLG_OUT)
        s_out[pos] = ids[i];
      else
        cand[atomicAdd(&ds->n_cand, 1)] = ids[i];  // overflow of the warp's list (rare)
    }
}

// appends a half-edge to the warp's work list (checked at once if the list is full)
template <int D>
__device__ __forceinline__ void push_half_edge(int he, int* s_he, int* s_nhe,
                                               const double* __restrict__ x,
                                               const int4* __restrict__ cells,
                                               const int* __restrict__ adj, double tol,
                                               double* __restrict__ sarr,
                                               int* __restrict__ cand_epoch, int epoch, int* s_out,
                                               int* s_nout, int* __restrict__ cand,
                                               DevScalars* ds) {
  const int pos = atomicAdd(s_nhe, 1);
  if (pos < FLG_HE) {
    s_he[pos] = he;
  } else {
    const int4 cl = __ldg(cells + (he >> 2));
    check_half_edge<D>(x, cells, adj, he >> 2, he & 3, cl, tol, sarr, cand_epoch, epoch, s_out,
                       s_nout, cand, ds);
  }
}

#ifndef OM_FLG_MINB
#define OM_FLG_MINB 4
#endif
template <int D>
__global__ void __launch_bounds__(FLG_BLOCK, OM_FLG_MINB)
    k_suspect_flags(const double* __restrict__ x, const int4* __restrict__ cells,
                    const int* __restrict__ adj, const int* __restrict__ v2c,
                    const int* __restrict__ ring, const int* __restrict__ ringc,
                    unsigned short* __restrict__ vflags, int N, double tol,
                    double* __restrict__ sarr, int* __restrict__ cand,
                    int* __restrict__ cand_epoch, DevScalars* ds, int vlo, int vhi) {
  if (ds->halt & 1) return;
  constexpr int NW = FLG_BLOCK / 32;
  // dynamic shared memory (more than the 48 KB a static allocation may have): FLG_SMEM bytes
  extern __shared__ __align__(16) unsigned char flg_smem[];
  int(*s_v)[FLG_RUN] = reinterpret_cast<int(*)[FLG_RUN]>(flg_smem);
  int(*s_he)[FLG_HE] = reinterpret_cast<int(*)[FLG_HE]>(s_v + NW);
  int(*s_out)[FLG_OUT] = reinterpret_cast<int(*)[FLG_OUT]>(s_he + NW);
  unsigned short(*s_f)[FLG_RUN] = reinterpret_cast<unsigned short(*)[FLG_RUN]>(s_out + NW);
  __shared__ int s_nout[NW], s_nhe[NW];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gwarp = blockIdx.x * NW + warp, nwarps = gridDim.x * NW;
  const int epoch = ds->epoch;
  // the flag words of the vertices [vlo, vhi) (vlo a multiple of FLG_PER)
  const int nruns = (vhi - vlo + FLG_RUN - 1) / FLG_RUN;
  N = vhi;
  for (int run = gwarp; run < nruns; run += nwarps) {
    const int vb = vlo + run * FLG_RUN + lane * FLG_PER;
    unsigned short w[FLG_PER];
    int cnt = 0;
    if (vb < N) {
      // (the flag array is padded to a multiple of FLG_PER words beyond N)
#pragma unroll
      for (int g = 0; g < FLG_PER / 8; g++) {
        const uint4 raw = *reinterpret_cast<const uint4*>(vflags + vb + 8 * g);
        const unsigned r[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int i = 0; i < 8; i++) {
          unsigned short f =
              (unsigned short)((r[i >> 1] >> (16 * (i & 1))) & (VF_SPOKES | VF_CHECKALL));
          if (vb + 8 * g + i >= N) f = 0;
          w[8 * g + i] = f;
          cnt += f ? 1 : 0;
        }
        if (raw.x | raw.y | raw.z | raw.w)  // every bit has been consumed by now
          *reinterpret_cast<uint4*>(vflags + vb + 8 * g) = make_uint4(0u, 0u, 0u, 0u);
      }
    } else {
#pragma unroll
      for (int i = 0; i < FLG_PER; i++) w[i] = 0;
    }
    if (!__any_sync(0xffffffffu, cnt != 0)) continue;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int pos = incl - cnt;
#pragma unroll
    for (int i = 0; i < FLG_PER; i++)
      if (w[i]) {
        s_v[warp][pos] = vb + i;
        s_f[warp][pos] = w[i];
        pos++;
      }
    if (lane == 0) {
      s_nout[warp] = 0;
      s_nhe[warp] = 0;
    }
    __syncwarp();
    // phase 1, one lane per flagged vertex: name the half-edges of its flagged spokes
    for (int it = lane; it < total; it += 32) {
      const int v = s_v[warp][it];
      const unsigned f = s_f[warp][it];
      if (!(f & VF_CHECKALL)) {
        // the vertex has ring rows: spoke q is the edge (v, n_q) of cell q = (v, n_q, n_{q+1})
        const int4* rp = reinterpret_cast<const int4*>(ring + (size_t)OM_RING_W * v);
        const int4* cp = reinterpret_cast<const int4*>(ringc + (size_t)OM_RING_W * v);
        const int4 a0 = __ldg(rp), a1 = __ldg(rp + 1), b0 = __ldg(cp), b1 = __ldg(cp + 1);
        const int e[OM_RING_W] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const int ec[OM_RING_W] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        const int k = 1 + ((e[0] >> 30) & 1) + ((e[1] >> 29) & 2) + ((e[2] >> 28) & 4);
#pragma unroll
        for (int q = 0; q < OM_RING_W; q++)
          if (q < k && ((f >> q) & 1u)) {
            int nn = e[0];  // n_{q+1}, cyclic
#pragma unroll
            for (int r2 = 1; r2 < OM_RING_W; r2++)
              if (r2 == q + 1 && r2 < k) nn = e[r2];
            nn &= RING_IDMASK;
            const int c = ec[q];
            if (c < 0) {
              atomicOr(&ds->err, OM_DEV_WALK);
              continue;
            }
            const int4 cl = __ldg(cells + c);
            const int ks = slot_of(cl, nn);
            if (ks < 0 || slot_of(cl, v) < 0)
              atomicOr(&ds->err, OM_DEV_WALK);  // the rows do not match the cells
            else
              push_half_edge<D>(4 * c + ks, s_he[warp], &s_nhe[warp], x, cells, adj, tol, sarr,
                                cand_epoch, epoch, s_out[warp], &s_nout[warp], cand, ds);
          }
        continue;
      }
      // no row (pinned vertex, or more than OM_RING_W cells): every spoke, by a star walk
      const int c0 = v2c[v];
      if (c0 == OM_NONE_CELL) continue;
      int4 cl = __ldg(cells + c0);
      int j = slot_of(cl, v);
      if (j < 0) {
        atomicOr(&ds->err, OM_DEV_WALK);
        continue;
      }
      // the edge of the start cell that the walk does NOT leave through
      push_half_edge<D>(4 * c0 + (j + 2) % 3, s_he[warp], &s_nhe[warp], x, cells, adj, tol, sarr,
                        cand_epoch, epoch, s_out[warp], &s_nout[warp], cand, ds);
      int cur = c0, kexit = (j + 1) % 3, hops = 0;
      bool open = false;
      while (true) {
        const int t = __ldg(adj + 4 * (size_t)cur + kexit);
        if (t < 0) {
          open = true;
          break;
        }
        const int cn = t >> 2, kn = t & 3;
        if (cn == c0) break;  // closed: back at the first edge
        push_half_edge<D>(4 * cur + kexit, s_he[warp], &s_nhe[warp], x, cells, adj, tol, sarr,
                          cand_epoch, epoch, s_out[warp], &s_nout[warp], cand, ds);
        cl = __ldg(cells + cn);
        const int jn = slot_of(cl, v);
        if (jn < 0 || jn == kn || ++hops > 4096) {
          atomicOr(&ds->err, OM_DEV_WALK);
          break;
        }
        cur = cn;
        kexit = 3 - jn - kn;
      }
      if (open) {
        // open fan (pinned boundary vertex): the spokes on the other side of the start cell
        cur = c0;
        kexit = (j + 2) % 3;  // pushed above
        hops = 0;
        while (true) {
          const int t = __ldg(adj + 4 * (size_t)cur + kexit);
          if (t < 0) break;
          const int cn = t >> 2, kn = t & 3;
          cl = __ldg(cells + cn);
          const int jn = slot_of(cl, v);
          if (jn < 0 || jn == kn || ++hops > 4096) {
            atomicOr(&ds->err, OM_DEV_WALK);
            break;
          }
          cur = cn;
          kexit = 3 - jn - kn;
          push_half_edge<D>(4 * cur + kexit, s_he[warp], &s_nhe[warp], x, cells, adj, tol, sarr,
                            cand_epoch, epoch, s_out[warp], &s_nout[warp], cand, ds);
        }
      }
    }
    __syncwarp();
    // phase 2, one lane per half-edge: the decision on the exact s
    const int nhe = min(s_nhe[warp], FLG_HE);
    for (int i = lane; i < nhe; i += 32) {
      const int he = s_he[warp][i];
      const int4 cl = __ldg(cells + (he >> 2));
      check_half_edge<D>(x, cells, adj, he >> 2, he & 3, cl, tol, sarr, cand_epoch, epoch,
                         s_out[warp], &s_nout[warp], cand, ds);
    }
    __syncwarp();
    const int nout = min(s_nout[warp], FLG_OUT);
    int base = 0;
    if (lane == 0 && nout) base = atomicAdd(&ds->n_cand, nout);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (int i = lane; i < nout; i += 32) cand[base + i] = s_out[warp][i];
    __syncwarp();  // the shared lists are reused by the next run
  }
}

// ---- control of the pipelined loop: what the host decides between the rounds of a pass
// starts a flip pass (and its first check round)
__global__ void k_pl_pass_begin(DevScalars* ds) {
  if (ds->halt & 1) return;
  ds->epoch++;
  ds->dirty_pass++;
  ds->n_dirty = 0;
  ds->n_flagged = 0;
  ds->stale = 0;
  ds->abort = 0;
  ds->n_cand = 0;
  ds->n_rec = 0;
  ds->n_flips = 0;
  ds->n_work = 0;
  ds->n_rounds = 0;
  ds->flips_prev = 0;
  ds->not_delaunay = 0;
  ds->pl_round = 0;
  ds->g_flips_prev = 0;
}

// starts the check of a further round
__global__ void k_pl_round_begin(DevScalars* ds) {
  if (ds->halt & 1) return;
  ds->epoch++;
  ds->n_cand = 0;
}

// Decides, after a check, whether its flips run.  The check before knows the candidates of the
// round to come, n_flips what the round before achieved: a round whose check flags nothing is
// never launched (with several GPUs that is three meetings saved per pass).
__global__ void k_pl_round_end(DevScalars* ds, cudaGraphConditionalHandle handle, int use_handle) {
  unsigned go = 0u;
  if (!(ds->halt & 1)) {
    // first decision of a pass: pass begin, flag check, this; later: flip1, flip2, round begin,
    // check, this
    ds->pl_launches += ds->pl_round > 0 ? 5 : 3;
    go = om_flip_decide(ds, (unsigned long long)ds->n_cand, (unsigned long long)ds->n_flips);
  }
  if (use_handle) cudaGraphSetConditional(handle, go);
}

// Host loop of one flip pass.  Every round is ONE chain of launches -- check, select, flip,
// twin patch -- whose kernels read their list lengths on the device and run block-stride
// loops on a bounded grid, so no round needs the host.  The pass therefore enqueues as many
// rounds as the previous pass needed (h->flip_spec) behind the full check and reads the
// scalars back ONCE; only if that guess was short does it continue with one readback per
// round.  Rounds enqueued past convergence are empty launches (a few microseconds each).
// flip_spec == 0 (the previous pass flagged nothing) reads back right after the check.
template <int D>
int flip_rounds(om_handle* h, double tol, int max_rounds, int64_t* n_flips, int32_t* n_rounds,
                int32_t* cap_hit, bool first_round_given, bool keep_pass = false) {
  const int C = (int)h->C;
  const int B = 256;
  const int GMAX = 148 * 8;  // resident blocks of the list kernels on one B200
  auto grid_for = [&](long long bound) {
    return std::min(om_grid(std::min<long long>(bound, C), B), GMAX);
  };
  auto launch_flips = [&](int n_host, const int* n_dev, long long bound) {
    const int G = grid_for(bound);
    OM_LAUNCH(h, k_flip1, G, B, h->cells, h->adj, h->sarr, h->cand, n_host,
              h->flip_epoch, h->reloc, h->adj_tmp, h->v2c, h->dirty, h->dirty_epoch,
              h->ds, n_dev, h->flt_vlo, h->flt_vhi);
    OM_LAUNCH(h, k_flip2, G, B, (int*)h->adj, h->adj_tmp, h->flip_epoch, h->reloc, h->cand,
              n_host, h->work_epoch, h->work, h->sarr, h->ds, n_dev, h->flt_clo, h->flt_chi);
    h->nbr_valid = false;
  };
  int rounds = 0, cap = 0;
  int spec = first_round_given ? 0 : h->flip_spec;
  // ---- round 0: full check (or the records of the sharded check), then its flips
  if (!first_round_given) {
    OM_LAUNCH(h, k_reset_flip_scalars, 1, 1, h->ds, keep_pass ? 0 : 1);
    OM_LAUNCH(h, (k_suspect<D, 0>), om_grid(C, B), B, h->x, h->cells, (const int*)h->adj, 0, C,
              (const int*)nullptr, tol, h->sarr, h->cand, h->cand_epoch,
              (FlipRec*)nullptr, h->ds, (const int*)nullptr, ShardInfo{0, 0, 0, 0, nullptr, 0, nullptr});
  }
  long long wb = 0;  // bound on the length of the next work list
  bool go = true;
  if (spec > 0 && max_rounds > 0) {
    launch_flips(0, &h->ds->n_cand, C);
    wb = C;
  } else {
    OM_TRY(om_fetch_scalars(h));
    OM_TRY(om_check_dev_err(h));
    const int n0 = h->hs->n_cand;
    if (n0 == 0) {
      go = false;
    } else if (max_rounds <= 0) {
      cap = 1;
      go = false;
    } else {
      launch_flips(n0, nullptr, n0);
      wb = 4ll * n0;
    }
    spec = 0;
  }
  // ---- rounds 1, 2, ...: `chain` rounds per readback
  int flips_seen = 0;
  for (int r = 1; go;) {
    bool capped = false;
    const int chain = std::max(spec, 1);
    spec = 0;
    for (int q = 0; q < chain && !capped; q++, r++) {
      const long long cb = std::min<long long>(2 * std::min<long long>(wb, C), C);  // candidates
      OM_LAUNCH(h, k_reset_flip_scalars, 1, 1, h->ds, 0);
      OM_LAUNCH(h, (k_suspect<D, 1>), grid_for(wb), B, h->x, h->cells, (const int*)h->adj, 0, 0,
                h->work, tol, h->sarr, h->cand, h->cand_epoch, (FlipRec*)nullptr,
                h->ds, (const int*)&h->ds->n_work, ShardInfo{0, 0, 0, 0, nullptr, 0, nullptr});
      if (r < max_rounds) {
        launch_flips(0, &h->ds->n_cand, cb);
        wb = std::min<long long>(4 * cb, C);
      } else {
        capped = true;  // this round only checks
      }
    }
    OM_TRY(om_fetch_scalars(h));
    OM_TRY(om_check_dev_err(h));
    rounds = h->hs->n_rounds;  // rounds that flipped at least one edge
    const int n_cand = h->hs->n_cand;
    if (n_cand == 0) break;  // nothing flagged any more (the chained flips were no-ops)
    if (capped) {
      cap = 1;
      break;
    }
    if (h->hs->n_flips == flips_seen) {
      // flagged but no mutual pair (exact ties of s around a cycle of cells): the mesh is
      // left as it is and the caller is told (the Python layer warns)
      cap = 2;
      break;
    }
    flips_seen = h->hs->n_flips;
    wb = 4ll * n_cand;
  }
  if (!first_round_given) h->flip_spec = std::min(h->hs->n_rounds, 8);
  if (n_flips) *n_flips = h->hs->n_flips;
  if (n_rounds) *n_rounds = rounds;
  if (cap_hit) *cap_hit = cap;
  return OM_OK;
}

}  // namespace

static ShardInfo shard_info(om_handle* h, int64_t clo, int64_t chi) {
  ShardInfo sh;
  sh.vlo = (int)h->own_lo;
  sh.vhi = (int)(h->own_hi >= 0 ? h->own_hi : h->N);
  sh.clo = (int)clo;
  sh.chi = (int)chi;
  const bool partitioned = h->own_hi >= 0 && h->valid_epoch && !h->all_valid;
  sh.valid_epoch = partitioned ? h->valid_epoch : nullptr;
  sh.valid_stamp = h->valid_stamp;
  sh.bflag = h->bflag;
  return sh;
}

// ---- round-wise pass for partitioned coordinates (om_flip_pass_begin / om_flip_round_check /
//      om_flip_add_records / om_flip_round_apply / om_flip_pass_end)
int om_flip_pass_begin_impl(om_handle* h) {
  if (!h->recs) CUDA_TRY(om_malloc(h, &h->recs, sizeof(FlipRec) * std::max<int64_t>(h->C, 1)));
  OM_LAUNCH(h, k_reset_flip_scalars, 1, 1, h->ds, 1);
  h->pass_work_bound = 0;
  return OM_OK;
}

int om_flip_round_check_impl(om_handle* h, double tol, int first, int64_t clo, int64_t chi,
                             int64_t* n_records, int32_t* stale, bool fetch) {
  const int B = 256;
  // also right for a repeated call after a coordinate refresh (clears records and `stale`)
  OM_LAUNCH(h, k_reset_flip_scalars, 1, 1, h->ds, 0);
  const ShardInfo sh = shard_info(h, clo, chi);
  // every rank applies all flips of the round, but enlists only what it will look at itself
  static const bool no_filter = getenv("OM_NO_LIST_FILTER") != nullptr;  // diagnostics
  if (!no_filter) {
    h->flt_vlo = sh.vlo;
    h->flt_vhi = sh.vhi;
    h->flt_clo = sh.clo;
    h->flt_chi = sh.chi;
  }
  if (first) {
    const int n = (int)(chi - clo);
    if (h->D == 2)
      OM_LAUNCH(h, (k_suspect<2, 2>), om_grid(n, B), B, h->x, h->cells, (const int*)h->adj,
                (int)clo, n, (const int*)nullptr, tol, h->sarr, h->cand, h->cand_epoch,
                h->recs, h->ds, (const int*)nullptr, sh);
    else
      OM_LAUNCH(h, (k_suspect<3, 2>), om_grid(n, B), B, h->x, h->cells, (const int*)h->adj,
                (int)clo, n, (const int*)nullptr, tol, h->sarr, h->cand, h->cand_epoch,
                h->recs, h->ds, (const int*)nullptr, sh);
  } else if (h->pass_work_bound > 0) {
    const int nb = (int)std::min<int64_t>(h->pass_work_bound, h->C);
    if (h->D == 2)
      OM_LAUNCH(h, (k_suspect<2, 3>), om_grid(nb, B), B, h->x, h->cells, (const int*)h->adj, 0, 0,
                h->work, tol, h->sarr, h->cand, h->cand_epoch, h->recs, h->ds,
                (const int*)&h->ds->n_work, sh);
    else
      OM_LAUNCH(h, (k_suspect<3, 3>), om_grid(nb, B), B, h->x, h->cells, (const int*)h->adj, 0, 0,
                h->work, tol, h->sarr, h->cand, h->cand_epoch, h->recs, h->ds,
                (const int*)&h->ds->n_work, sh);
  }
  if (!fetch) return OM_OK;  // the counts stay on the device (om_flip_round_pack)
  OM_TRY(om_fetch_scalars(h));
  OM_TRY(om_check_dev_err(h));
  if (n_records) *n_records = h->hs->n_rec;
  if (stale) *stale = h->hs->stale;
  return OM_OK;
}

int om_flip_round_pack_impl(om_handle* h, int cap, void* slot_dev) {
  OM_LAUNCH(h, k_round_pack, om_grid(std::max(cap, 1), 256), 256, h->recs, h->ds, cap,
            (FlipRec*)slot_dev);
  CUDA_TRY(cudaGetLastError());
  return OM_OK;
}

int om_flip_round_apply_gathered_impl(om_handle* h, const void* gathered, int P, int cap,
                                      int64_t* n_cand, int64_t* n_flips_total, int32_t* abort_bits,
                                      int64_t* own_records) {
  const int B = 256;
  const FlipRec* g = (const FlipRec*)gathered;
  OM_LAUNCH(h, k_round_scan, 1, 1, g, P, cap, h->ds);
  OM_LAUNCH(h, k_apply_gathered, om_grid((int64_t)P * cap, B), B, g, P, cap, h->sarr, h->cand,
            h->cand_epoch, h->ds);
  const int bound = (int)std::min<int64_t>(2ll * P * cap, h->C);
  const int* nd = &h->ds->n_cand;
  OM_LAUNCH(h, k_flip1, om_grid(bound, B), B, h->cells, h->adj, h->sarr, h->cand, 0,
            h->flip_epoch, h->reloc, h->adj_tmp, h->v2c, h->dirty, h->dirty_epoch,
            h->ds, nd, h->flt_vlo, h->flt_vhi);
  OM_LAUNCH(h, k_flip2, om_grid(bound, B), B, (int*)h->adj, h->adj_tmp, h->flip_epoch, h->reloc,
            h->cand, 0, h->work_epoch, h->work, h->sarr, h->ds, nd, h->flt_clo, h->flt_chi);
  h->nbr_valid = false;
  OM_TRY(om_fetch_scalars(h));
  OM_TRY(om_check_dev_err(h));
  if (abort_bits) *abort_bits = h->hs->abort;
  if (own_records) *own_records = h->hs->max_count;
  if (!h->hs->abort) h->pass_work_bound = h->hs->n_work;
  if (n_cand) *n_cand = h->hs->n_cand;
  if (n_flips_total) *n_flips_total = h->hs->n_flips;
  return OM_OK;
}

// records of ALL ranks have been added: select, flip, patch twins, next work list
int om_flip_round_apply_impl(om_handle* h, int64_t total_records, int64_t* n_cand,
                             int64_t* n_flips_total) {
  const int B = 256;
  const int bound = (int)std::min<int64_t>(2 * total_records, h->C);
  if (bound > 0) {
    const int* nd = &h->ds->n_cand;
    OM_LAUNCH(h, k_flip1, om_grid(bound, B), B, h->cells, h->adj, h->sarr, h->cand, 0,
              h->flip_epoch, h->reloc, h->adj_tmp, h->v2c, h->dirty, h->dirty_epoch,
              h->ds, nd, h->flt_vlo, h->flt_vhi);
    OM_LAUNCH(h, k_flip2, om_grid(bound, B), B, (int*)h->adj, h->adj_tmp, h->flip_epoch, h->reloc,
              h->cand, 0, h->work_epoch, h->work, h->sarr, h->ds, nd, h->flt_clo, h->flt_chi);
    h->nbr_valid = false;
  }
  OM_TRY(om_fetch_scalars(h));
  OM_TRY(om_check_dev_err(h));
  h->pass_work_bound = h->hs->n_work;
  if (n_cand) *n_cand = h->hs->n_cand;
  if (n_flips_total) *n_flips_total = h->hs->n_flips;
  return OM_OK;
}

int om_flip_pass_end_impl(om_handle* h, int64_t* n_flips, int32_t* n_rounds) {
  // the last readback of the pass (a check round that found nothing) is current
  if (n_flips) *n_flips = h->hs->n_flips;
  if (n_rounds) *n_rounds = h->hs->n_rounds;
  if (h->hs->n_flips > 0) OM_TRY(om_rebuild_rings(h, false));
  return OM_OK;
}

// ---- sharded first round (om_flip_check_range / om_flip_add_records / om_flip_finish)
int om_flip_check_range_impl(om_handle* h, double tol, int64_t clo, int64_t chi,
                             int64_t* n_records) {
  const int B = 256;
  if (!h->recs) CUDA_TRY(om_malloc(h, &h->recs, sizeof(FlipRec) * std::max<int64_t>(h->C, 1)));
  OM_LAUNCH(h, k_reset_flip_scalars, 1, 1, h->ds, 1);  // a new flip pass starts here
  const int n = (int)(chi - clo);
  if (n > 0) {
    if (h->D == 2)
      OM_LAUNCH(h, (k_suspect<2, 2>), om_grid(n, B), B, h->x, h->cells, (const int*)h->adj,
                (int)clo, n, (const int*)nullptr, tol, h->sarr, h->cand, h->cand_epoch,
                h->recs, h->ds, (const int*)nullptr, shard_info(h, clo, chi));
    else
      OM_LAUNCH(h, (k_suspect<3, 2>), om_grid(n, B), B, h->x, h->cells, (const int*)h->adj,
                (int)clo, n, (const int*)nullptr, tol, h->sarr, h->cand, h->cand_epoch,
                h->recs, h->ds, (const int*)nullptr, shard_info(h, clo, chi));
  }
  OM_TRY(om_fetch_scalars(h));
  OM_TRY(om_check_dev_err(h));
  if (n_records) *n_records = h->hs->n_rec;
  return OM_OK;
}

int om_flip_add_records_impl(om_handle* h, const void* recs, int64_t n) {
  if (n <= 0) return OM_OK;
  OM_LAUNCH(h, k_apply_records, om_grid(n, 256), 256, (const FlipRec*)recs, (int)n, h->sarr,
            h->cand, h->cand_epoch, h->ds);
  CUDA_TRY(cudaGetLastError());
  return OM_OK;
}

int om_flip_impl(om_handle* h, double tol, int max_rounds, int64_t* n_flips, int32_t* n_rounds,
                 int32_t* cap_hit, bool first_round_given) {
  int64_t local_flips = 0;
  if (!n_flips) n_flips = &local_flips;
  if (n_flips) *n_flips = 0;
  if (n_rounds) *n_rounds = 0;
  if (cap_hit) *cap_hit = 0;
  if (h->C == 0) return OM_OK;
  // whole-mesh pass: every list covers the whole mesh
  h->flt_vlo = h->flt_clo = 0;
  h->flt_vhi = h->flt_chi = 0x7fffffff;
  if (h->timing) cudaEventRecord(h->ev[2], h->stream);
  int rc = (h->D == 2)
               ? flip_rounds<2>(h, tol, max_rounds, n_flips, n_rounds, cap_hit, first_round_given)
               : flip_rounds<3>(h, tol, max_rounds, n_flips, n_rounds, cap_hit, first_round_given);
  // ring rows of the vertices whose stars changed (hs->n_dirty is current: the last
  // readback of the pass came after the last flip)
  if (rc == OM_OK && n_flips && *n_flips > 0) rc = om_rebuild_rings(h, false);
  h->delaunay_clean = rc == OM_OK && (!cap_hit || *cap_hit == 0) && tol == 0.0 && h->own_hi < 0;
  if (h->timing && rc == OM_OK) {
    cudaEventRecord(h->ev[3], h->stream);
    cudaEventSynchronize(h->ev[3]);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]) == cudaSuccess) {
      h->t_flip_ms += ms;
      h->n_flip++;
    }
  }
  return rc;
}

// ---- launchers of the pipelined loop (loop.cu); everything reads its list lengths and
// stamps on the device and returns at once when the loop has halted
int om_pl_launch_flags_check(om_handle* h, const double* xin) {
  OM_LAUNCH(h, k_pl_pass_begin, 1, 1, h->ds);
  int vlo = 0, vhi = (int)h->N;
  om_shared_vertex_range(h, &vlo, &vhi);
  const int blocks = om_grid(vhi - vlo, FLG_BLOCK * FLG_PER);
  const int G = std::min(blocks, 148 * 8);
  static bool attr_set_dev[64] = {};  // (the attribute is per device)
  bool& attr_set = attr_set_dev[h->device & 63];
  if (!attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(k_suspect_flags<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)FLG_SMEM));
    CUDA_TRY(cudaFuncSetAttribute(k_suspect_flags<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)FLG_SMEM));
    attr_set = true;
  }
  if (h->D == 2)
    OM_LAUNCH_SMEM(h, k_suspect_flags<2>, G, FLG_BLOCK, FLG_SMEM, xin, h->cells,
                   (const int*)h->adj, h->v2c, (const int*)h->ring, (const int*)h->ringc,
                   h->vflags, (int)h->N, 0.0, h->sarr, h->cand, h->cand_epoch, h->ds, vlo, vhi);
  else
    OM_LAUNCH_SMEM(h, k_suspect_flags<3>, G, FLG_BLOCK, FLG_SMEM, xin, h->cells,
                   (const int*)h->adj, h->v2c, (const int*)h->ring, (const int*)h->ringc,
                   h->vflags, (int)h->N, 0.0, h->sarr, h->cand, h->cand_epoch, h->ds, vlo, vhi);
  CUDA_TRY(cudaGetLastError());
  return OM_OK;
}

// flip + twin patch on the candidate list (the flip kernels return at
// once on an empty list; ds->abort is 0 in the pipelined loop)
int om_pl_launch_flips(om_handle* h) {
  const int B = 256, G = 148 * 8;
  const int* nd = &h->ds->n_cand;
  OM_LAUNCH(h, k_flip1, G, B, h->cells, h->adj, h->sarr, h->cand, 0, h->flip_epoch, h->reloc,
            h->adj_tmp, h->v2c, h->dirty, h->dirty_epoch, h->ds, nd, 0, 0x7fffffff);
  OM_LAUNCH(h, k_flip2, G, B, (int*)h->adj, h->adj_tmp, h->flip_epoch, h->reloc, h->cand, 0,
            h->work_epoch, h->work, h->sarr, h->ds, nd, 0, 0x7fffffff);
  CUDA_TRY(cudaGetLastError());
  h->nbr_valid = false;
  return OM_OK;
}

int om_pl_launch_flips_part(om_handle* h, int which) {
  const int B = 256, G = 148 * 8;
  const int* nd = &h->ds->n_cand;
  if (which == 1)
    OM_LAUNCH(h, k_flip1, G, B, h->cells, h->adj, h->sarr, h->cand, 0, h->flip_epoch, h->reloc,
              h->adj_tmp, h->v2c, h->dirty, h->dirty_epoch, h->ds, nd, 0, 0x7fffffff);
  else
    OM_LAUNCH(h, k_flip2, G, B, (int*)h->adj, h->adj_tmp, h->flip_epoch, h->reloc, h->cand, 0,
              h->work_epoch, h->work, h->sarr, h->ds, nd, 0, 0x7fffffff);
  CUDA_TRY(cudaGetLastError());
  h->nbr_valid = false;
  return OM_OK;
}

// the check of a further round alone (its flips follow through om_pl_launch_flips_part)
int om_pl_launch_round_check(om_handle* h, const double* xin) {
  const int B = 256, G = 148 * 8;
  OM_LAUNCH(h, k_pl_round_begin, 1, 1, h->ds);
  const ShardInfo none{0, 0, 0, 0, nullptr, 0, nullptr};
  if (h->D == 2)
    OM_LAUNCH(h, (k_suspect<2, 1>), G, B, xin, h->cells, (const int*)h->adj, 0, 0, h->work, 0.0,
              h->sarr, h->cand, h->cand_epoch, (FlipRec*)nullptr, h->ds,
              (const int*)&h->ds->n_work, none);
  else
    OM_LAUNCH(h, (k_suspect<3, 1>), G, B, xin, h->cells, (const int*)h->adj, 0, 0, h->work, 0.0,
              h->sarr, h->cand, h->cand_epoch, (FlipRec*)nullptr, h->ds,
              (const int*)&h->ds->n_work, none);
  CUDA_TRY(cudaGetLastError());
  return OM_OK;
}

// a further round: the flips of the last check, then the exact check of the cells they touched
int om_pl_launch_round(om_handle* h, const double* xin) {
  OM_TRY(om_pl_launch_flips(h));
  return om_pl_launch_round_check(h, xin);
}

int om_pl_launch_round_end(om_handle* h, unsigned long long handle, int use_handle) {
  OM_LAUNCH(h, k_pl_round_end, 1, 1, h->ds, (cudaGraphConditionalHandle)handle, use_handle);
  CUDA_TRY(cudaGetLastError());
  return OM_OK;
}
