// K2 + K3: GPU flip-until-Delaunay.
//
// Replaces meshplex's MeshTri.flip_until_delaunay() called by the optimize() loop after
// every step (/root/reference/README.md:131-132).  Per round:
//   k_ce      one thread per cell: covolume/edge ratios ce_k = -ed_k / (4A) of its 3 edges
//   k_select  one thread per cell: s = ce(own) + ce(twin) per interior edge; flag s < -tol;
//             keep the most negative flagged edge of the cell (ties: lowest local index)
//   k_flip1   an edge kept by BOTH adjacent cells is flipped (an independent set: every
//             cell takes part in at most one flip); rewrites the two cells, records where
//             the four outer half-edges move
//   k_flip2   patches the twin table using the relocation records (race-free when two
//             neighbouring cells flip in the same round)
// iterated until no edge is flagged.  (a0,k0) of a flip is the half-edge with the smaller
// id 3*row+k in the caller's cell numbering, so the cell array is identical to the
// oracle's, row for row.
#include "common.cuh"
#include "geom.cuh"

namespace {

template <int D>
__global__ void __launch_bounds__(256)
    k_ce(const double* __restrict__ x, const int4* __restrict__ cells, int C,
         double* __restrict__ ce, DevScalars* ds) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  int4 cl = cells[c];
  Vec<D> P0 = ld_point<D>(x, cl.x), P1 = ld_point<D>(x, cl.y), P2 = ld_point<D>(x, cl.z);
  CellGeo<D> g = cell_geo<D>(P0, P1, P2);
  if (!(g.vol2 > 0.0)) {
    atomicOr(&ds->err, OM_DEV_DEGENERATE);
    return;
  }
  const double inv4A = 0.25 / sqrt(g.vol2);
  double2* o = reinterpret_cast<double2*>(ce + 4 * (size_t)c);
  o[0] = make_double2(-g.ed0 * inv4A, -g.ed1 * inv4A);
  o[1] = make_double2(-g.ed2 * inv4A, 0.0);
}

__global__ void __launch_bounds__(256)
    k_select(const int4* __restrict__ adj, const double* __restrict__ ce, int C, double tol,
             int8_t* __restrict__ best, DevScalars* ds) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  int nflag = 0;
  if (c < C) {
    int4 a = adj[c];
    const double2* o = reinterpret_cast<const double2*>(ce + 4 * (size_t)c);
    double2 c01 = o[0], c2 = o[1];
    double own[3] = {c01.x, c01.y, c2.x};
    int tw[3] = {a.x, a.y, a.z};
    int b = -1;
    double sb = 0.0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      if (tw[k] >= 0) {
        double s = own[k] + __ldg(ce + tw[k]);
        if (s < -tol) {
          nflag++;
          if (b < 0 || s < sb) {
            b = k;
            sb = s;
          }
        }
      }
    }
    best[c] = (int8_t)b;
  }
  for (int o = 16; o > 0; o >>= 1) nflag += __shfl_xor_sync(0xffffffffu, nflag, o);
  if ((threadIdx.x & 31) == 0 && nflag) atomicAdd(&ds->n_flagged, nflag);
}

__global__ void __launch_bounds__(256)
    k_flip1(int4* __restrict__ cells, const int4* __restrict__ adj, const int8_t* __restrict__ best,
            int C, int epoch, int* __restrict__ flip_epoch, int* __restrict__ reloc,
            int4* __restrict__ adj_tmp, int* __restrict__ v2c, DevScalars* ds) {
  int a0 = blockIdx.x * blockDim.x + threadIdx.x;
  int nf = 0;
  if (a0 < C) {
    const int k0 = best[a0];
    if (k0 >= 0) {
      const int4 adjA = adj[a0];
      const int t = cell_get(adjA, k0);
      const int a1 = t >> 2, k1 = t & 3;
      if (best[a1] == k1) {
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
            nf = 1;
          }
        }
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) nf += __shfl_xor_sync(0xffffffffu, nf, o);
  if ((threadIdx.x & 31) == 0 && nf) atomicAdd(&ds->n_flips, nf);
}

__global__ void __launch_bounds__(256)
    k_flip2(int* __restrict__ adj, const int4* __restrict__ adj_tmp,
            const int* __restrict__ flip_epoch, const int* __restrict__ reloc, int C, int epoch) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (flip_epoch[c] != epoch) return;
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
    }
    adj[4 * c + s] = tn;
  }
  adj[4 * c + 2] = t.z;
}

__global__ void k_reset_flip_scalars(DevScalars* ds) {
  ds->n_flagged = 0;
  ds->n_flips = 0;
}

template <int D>
int flip_rounds(om_handle* h, double tol, int max_rounds, int64_t* n_flips, int32_t* n_rounds,
                int32_t* cap_hit) {
  const int C = (int)h->C;
  const int B = 256, G = om_grid(C, B);
  int64_t total = 0;
  int rounds = 0;
  int cap = 0;
  for (int r = 0;; r++) {
    OM_LAUNCH(h, k_reset_flip_scalars, 1, 1, h->ds);
    OM_LAUNCH(h, k_ce<D>, G, B, h->x, h->cells, C, h->ce, h->ds);
    OM_LAUNCH(h, k_select, G, B, h->adj, h->ce, C, tol, h->best, h->ds);
    OM_TRY(om_fetch_scalars(h));
    OM_TRY(om_check_dev_err(h));
    if (h->hs->n_flagged == 0) break;
    if (r >= max_rounds) {
      cap = 1;
      break;
    }
    h->epoch++;
    OM_LAUNCH(h, k_flip1, G, B, h->cells, h->adj, h->best, C, h->epoch, h->flip_epoch, h->reloc,
              h->adj_tmp, h->v2c, h->ds);
    OM_LAUNCH(h, k_flip2, G, B, (int*)h->adj, h->adj_tmp, h->flip_epoch, h->reloc, C, h->epoch);
    OM_TRY(om_fetch_scalars(h));
    OM_TRY(om_check_dev_err(h));
    total += h->hs->n_flips;
    rounds++;
    h->nbr_valid = false;
  }
  if (n_flips) *n_flips = total;
  if (n_rounds) *n_rounds = rounds;
  if (cap_hit) *cap_hit = cap;
  return OM_OK;
}

}  // namespace

int om_flip_impl(om_handle* h, double tol, int max_rounds, int64_t* n_flips, int32_t* n_rounds,
                 int32_t* cap_hit) {
  if (n_flips) *n_flips = 0;
  if (n_rounds) *n_rounds = 0;
  if (cap_hit) *cap_hit = 0;
  if (h->C == 0) return OM_OK;
  if (h->timing) cudaEventRecord(h->ev[2], h->stream);
  int rc = (h->D == 2) ? flip_rounds<2>(h, tol, max_rounds, n_flips, n_rounds, cap_hit)
                       : flip_rounds<3>(h, tol, max_rounds, n_flips, n_rounds, cap_hit);
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
