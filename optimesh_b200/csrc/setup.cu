// Mesh setup on the device: upload, spatial (Morton) renumbering of vertices and cells,
// half-edge twin table, boundary flags, vertex -> incident-cell table.
//
// Replaces what meshplex.MeshTri(points, cells) + create_edges() do on the host
// (/root/reference/README.md:131): np.unique over sorted vertex pairs
// becomes one radix sort of 3C edge keys.  The sorts use CUB (ships with the toolkit);
// they run once per mesh, not per step.
#include <cub/cub.cuh>
#include <algorithm>
#include <cstring>

#include "common.cuh"
#include "geom.cuh"

namespace {

__device__ __forceinline__ unsigned long long enc_double(double d) {
  unsigned long long b = (unsigned long long)__double_as_longlong(d);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ inline double dec_double(unsigned long long e) {
  unsigned long long b = (e >> 63) ? (e & 0x7fffffffffffffffull) : ~e;
  double d;
  memcpy(&d, &b, 8);
  return d;
}

// bbox[0..2] = encoded min per dim, bbox[3..5] = encoded max per dim
template <int D>
__global__ void k_bbox(const double* __restrict__ raw, int64_t N, unsigned long long* bbox) {
  double lo[D], hi[D];
#pragma unroll
  for (int k = 0; k < D; k++) {
    lo[k] = INFINITY;
    hi[k] = -INFINITY;
  }
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N;
       i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int k = 0; k < D; k++) {
      double v = raw[i * D + k];
      lo[k] = fmin(lo[k], v);
      hi[k] = fmax(hi[k], v);
    }
  }
#pragma unroll
  for (int k = 0; k < D; k++) {
    for (int o = 16; o > 0; o >>= 1) {
      lo[k] = fmin(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
      hi[k] = fmax(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&bbox[k], enc_double(lo[k]));
      atomicMax(&bbox[3 + k], enc_double(hi[k]));
    }
  }
}

__device__ __forceinline__ unsigned long long spread2(unsigned long long x) {
  // 32 bits -> every other bit of 64
  x &= 0xffffffffull;
  x = (x | (x << 16)) & 0x0000ffff0000ffffull;
  x = (x | (x << 8)) & 0x00ff00ff00ff00ffull;
  x = (x | (x << 4)) & 0x0f0f0f0f0f0f0f0full;
  x = (x | (x << 2)) & 0x3333333333333333ull;
  x = (x | (x << 1)) & 0x5555555555555555ull;
  return x;
}
__device__ __forceinline__ unsigned long long spread3(unsigned long long x) {
  // 21 bits -> every third bit of 63
  x &= 0x1fffffull;
  x = (x | (x << 32)) & 0x1f00000000ffffull;
  x = (x | (x << 16)) & 0x1f0000ff0000ffull;
  x = (x | (x << 8)) & 0x100f00f00f00f00full;
  x = (x | (x << 4)) & 0x10c30c30c30c30c3ull;
  x = (x | (x << 2)) & 0x1249249249249249ull;
  return x;
}

template <int D>
__global__ void k_morton(const double* __restrict__ raw, int64_t N,
                         const unsigned long long* __restrict__ bbox,
                         unsigned long long* __restrict__ keys, int* __restrict__ vals) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= N) return;
  unsigned long long q[D];
  const double scale_max = (D == 2) ? 4294967296.0 : 2097152.0;
#pragma unroll
  for (int k = 0; k < D; k++) {
    double lo = dec_double(bbox[k]), hi = dec_double(bbox[3 + k]);
    double w = hi - lo;
    double t = (w > 0.0) ? (raw[i * D + k] - lo) / w : 0.0;
    double s = t * scale_max;
    s = fmin(fmax(s, 0.0), scale_max - 1.0);
    q[k] = (unsigned long long)s;
  }
  unsigned long long key;
  if (D == 2)
    key = spread2(q[0]) | (spread2(q[1]) << 1);
  else
    key = spread3(q[0]) | (spread3(q[1]) << 1) | (spread3(q[D - 1]) << 2);
  keys[i] = key;
  vals[i] = (int)i;
}

__global__ void k_invert_perm(const int* __restrict__ perm, int64_t N, int* __restrict__ inv) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < N) inv[perm[i]] = (int)i;
}

template <int D>
__global__ void k_gather_points(const double* __restrict__ raw, const int* __restrict__ perm,
                                int64_t N, double* __restrict__ x) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= N) return;
  int64_t src = perm ? perm[i] : i;
  Vec<D> p;
#pragma unroll
  for (int k = 0; k < D; k++) p.v[k] = raw[src * D + k];
  st_point<D>(x, (int)i, p);
}

template <typename T>
__global__ void k_relabel_cells(const T* __restrict__ raw, int64_t C, int64_t N,
                                const int* __restrict__ inv, int* __restrict__ tmp3,
                                unsigned int* __restrict__ keys, int* __restrict__ vals,
                                int* __restrict__ err) {
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= C) return;
  long long a = (long long)raw[3 * c], b = (long long)raw[3 * c + 1], d = (long long)raw[3 * c + 2];
  if (a < 0 || b < 0 || d < 0 || a >= N || b >= N || d >= N) {
    atomicOr(err, OM_DEV_INDEX);
    a = b = d = 0;
  }
  int ia = inv ? inv[a] : (int)a, ib = inv ? inv[b] : (int)b, id = inv ? inv[d] : (int)d;
  tmp3[3 * c] = ia;
  tmp3[3 * c + 1] = ib;
  tmp3[3 * c + 2] = id;
  if (keys) {
    keys[c] = (unsigned int)min(ia, min(ib, id));
    vals[c] = (int)c;
  }
}

__global__ void k_build_cells4(const int* __restrict__ tmp3, const int* __restrict__ cperm,
                               int64_t C, int4* __restrict__ cells) {
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= C) return;
  int64_t src = cperm ? cperm[c] : c;
  cells[c] = make_int4(tmp3[3 * src], tmp3[3 * src + 1], tmp3[3 * src + 2], (int)src);
}

__global__ void k_edge_keys(const int4* __restrict__ cells, int64_t C, int bits,
                            unsigned long long* __restrict__ keys, int* __restrict__ vals) {
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= C) return;
  int4 cl = cells[c];
  int v[3] = {cl.x, cl.y, cl.z};
#pragma unroll
  for (int k = 0; k < 3; k++) {
    int u = v[(k + 1) % 3], w = v[(k + 2) % 3];
    unsigned long long lo = (unsigned long long)min(u, w), hi = (unsigned long long)max(u, w);
    keys[3 * c + k] = (lo << bits) | hi;
    vals[3 * c + k] = (int)(4 * c + k);
  }
}

__global__ void k_pair_twins(const unsigned long long* __restrict__ keys,
                             const int* __restrict__ vals, int64_t M, int bits,
                             int* __restrict__ adj, uint8_t* __restrict__ bflag,
                             int* __restrict__ err) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M) return;
  unsigned long long k = keys[i];
  bool prev_same = i > 0 && keys[i - 1] == k;
  if (prev_same) return;
  bool next_same = i + 1 < M && keys[i + 1] == k;
  if (next_same) {
    if (i + 2 < M && keys[i + 2] == k) atomicOr(err, OM_DEV_NONMANIFOLD);
    int a = vals[i], b = vals[i + 1];
    adj[a] = b;
    adj[b] = a;
  } else {
    adj[vals[i]] = -1;
    unsigned long long mask = (1ull << bits) - 1ull;
    bflag[(int)(k >> bits)] = 1;
    bflag[(int)(k & mask)] = 1;
  }
}

__global__ void k_fill_double(double* p, int64_t n, double v) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

__global__ void k_fill_int(int* p, int64_t n, int v) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

__global__ void k_v2c(const int4* __restrict__ cells, int64_t C, int* __restrict__ v2c) {
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= C) return;
  int4 cl = cells[c];
  atomicMin(&v2c[cl.x], (int)c);
  atomicMin(&v2c[cl.y], (int)c);
  atomicMin(&v2c[cl.z], (int)c);
}

template <typename K, typename V>
int sort_pairs(om_handle* h, const K* kin, K* kout, const V* vin, V* vout, int64_t n, int end_bit) {
  size_t bytes = 0;
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, n, 0, end_bit,
                                           h->stream));
  void* tmp = nullptr;
  CUDA_TRY(om_malloc(h, &tmp, bytes ? bytes : 1));
  cudaError_t e =
      cub::DeviceRadixSort::SortPairs(tmp, bytes, kin, kout, vin, vout, n, 0, end_bit, h->stream);
  om_free(h, tmp);  // stream ordered: released once the sort has run
  CUDA_TRY(e);
  return OM_OK;
}

int bits_for(int64_t n) {
  int b = 1;
  while ((1ll << b) < n) b++;
  return b;
}

template <int D>
int setup_points(om_handle* h, const double* raw, bool renumber) {
  const int64_t N = h->N;
  const int B = 256;
  if (renumber && N > 1) {
    unsigned long long* bbox = nullptr;
    CUDA_TRY(om_malloc(h, &bbox, 6 * sizeof(unsigned long long)));
    unsigned long long init[6];
    for (int k = 0; k < 3; k++) {
      init[k] = ~0ull;
      init[3 + k] = 0ull;
    }
    CUDA_TRY(cudaMemcpyAsync(bbox, init, sizeof(init), cudaMemcpyHostToDevice, h->stream));
    int g = (int)std::min<int64_t>(om_grid(N, B), 148 * 8);
    OM_LAUNCH(h, k_bbox<D>, g, B, raw, N, bbox);
    unsigned long long *keys = nullptr, *keys2 = nullptr;
    int *vals = nullptr;
    CUDA_TRY(om_malloc(h, &keys, N * 8));
    CUDA_TRY(om_malloc(h, &keys2, N * 8));
    CUDA_TRY(om_malloc(h, &vals, N * 4));
    CUDA_TRY(om_malloc(h, &h->perm, N * 4));
    CUDA_TRY(om_malloc(h, &h->inv_perm, N * 4));
    OM_LAUNCH(h, k_morton<D>, om_grid(N, B), B, raw, N, bbox, keys, vals);
    OM_TRY(sort_pairs(h, keys, keys2, vals, h->perm, N, D == 2 ? 64 : 63));
    OM_LAUNCH(h, k_invert_perm, om_grid(N, B), B, h->perm, N, h->inv_perm);
    om_free(h, keys);
    om_free(h, keys2);
    om_free(h, vals);
    om_free(h, bbox);
  }
  OM_LAUNCH(h, k_gather_points<D>, om_grid(N, B), B, raw, h->perm, N, h->x);
  CUDA_TRY(cudaMemsetAsync(h->xnew, 0, sizeof(double) * (N + OM_POINT_PAD) * h->PD, h->stream));
  return OM_OK;
}

}  // namespace

int om_setup_mesh(om_handle* h, const double* points_dev, const void* cells_dev, int flags) {
  const int64_t N = h->N, C = h->C;
  const int B = 256;
  const bool renumber = (flags & OM_RENUMBER) != 0;
  // OM_POINT_PAD extra vertices: lets a caller all-gather equal-sized chunks in place
  CUDA_TRY(om_malloc(h, &h->x, sizeof(double) * (N + OM_POINT_PAD) * h->PD));
  CUDA_TRY(om_malloc(h, &h->xnew, sizeof(double) * (N + OM_POINT_PAD) * h->PD));
  CUDA_TRY(cudaMemsetAsync(h->x, 0, sizeof(double) * (N + OM_POINT_PAD) * h->PD, h->stream));
  CUDA_TRY(om_malloc(h, &h->cells, sizeof(int4) * std::max<int64_t>(C, 1)));
  CUDA_TRY(om_malloc(h, &h->adj, sizeof(int4) * std::max<int64_t>(C, 1)));
  CUDA_TRY(om_malloc(h, &h->adj_tmp, sizeof(int4) * std::max<int64_t>(C, 1)));
  CUDA_TRY(om_malloc(h, &h->v2c, sizeof(int) * N));
  CUDA_TRY(om_malloc(h, &h->bflag, N));
  CUDA_TRY(om_malloc(h, &h->cand, sizeof(int) * std::max<int64_t>(C, 1)));
  CUDA_TRY(om_malloc(h, &h->work, sizeof(int) * std::max<int64_t>(C, 1)));
  CUDA_TRY(om_malloc(h, &h->work_epoch, sizeof(int) * std::max<int64_t>(C, 1)));
  CUDA_TRY(om_malloc(h, &h->cand_epoch, sizeof(int) * std::max<int64_t>(C, 1)));
  CUDA_TRY(om_malloc(h, &h->sarr, sizeof(double) * 4 * std::max<int64_t>(C, 1)));
  CUDA_TRY(om_malloc(h, &h->flip_epoch, sizeof(int) * std::max<int64_t>(C, 1)));
  CUDA_TRY(om_malloc(h, &h->reloc, sizeof(int) * 4 * std::max<int64_t>(C, 1)));
  CUDA_TRY(om_malloc(h, &h->ds, sizeof(DevScalars)));
  CUDA_TRY(cudaMallocHost(&h->hs, sizeof(DevScalars)));
  CUDA_TRY(om_malloc(h, &h->partials, sizeof(double) * 8 * 2048));
  CUDA_TRY(cudaMemsetAsync(h->ds, 0, sizeof(DevScalars), h->stream));
  CUDA_TRY(cudaMemsetAsync(h->flip_epoch, 0, sizeof(int) * std::max<int64_t>(C, 1), h->stream));
  CUDA_TRY(cudaMemsetAsync(h->work_epoch, 0, sizeof(int) * std::max<int64_t>(C, 1), h->stream));
  CUDA_TRY(cudaMemsetAsync(h->cand_epoch, 0, sizeof(int) * std::max<int64_t>(C, 1), h->stream));
  OM_LAUNCH(h, k_fill_double, om_grid(4 * std::max<int64_t>(C, 1), B), B, h->sarr,
            4 * std::max<int64_t>(C, 1), (double)INFINITY);
  CUDA_TRY(cudaMemsetAsync(h->bflag, 0, N, h->stream));
  CUDA_TRY(cudaMemsetAsync(h->adj, 0xff, sizeof(int4) * std::max<int64_t>(C, 1), h->stream));

  if (h->D == 2)
    OM_TRY(setup_points<2>(h, points_dev, renumber));
  else
    OM_TRY(setup_points<3>(h, points_dev, renumber));

  // cells: relabel, sort by smallest vertex, pack to int4
  int* tmp3 = nullptr;
  unsigned int *ckeys = nullptr, *ckeys2 = nullptr;
  int *cvals = nullptr, *cperm = nullptr;
  CUDA_TRY(om_malloc(h, &tmp3, sizeof(int) * 3 * std::max<int64_t>(C, 1)));
  const bool sort_cells = renumber && C > 1;
  if (sort_cells) {
    CUDA_TRY(om_malloc(h, &ckeys, 4 * C));
    CUDA_TRY(om_malloc(h, &ckeys2, 4 * C));
    CUDA_TRY(om_malloc(h, &cvals, 4 * C));
    CUDA_TRY(om_malloc(h, &cperm, 4 * C));
  }
  if (C > 0) {
    if (h->cells_itemsize == 4)
      OM_LAUNCH(h, k_relabel_cells<int>, om_grid(C, B), B, (const int*)cells_dev, C, N,
                h->inv_perm, tmp3, ckeys, cvals, &h->ds->err);
    else
      OM_LAUNCH(h, k_relabel_cells<long long>, om_grid(C, B), B, (const long long*)cells_dev, C, N,
                h->inv_perm, tmp3, ckeys, cvals, &h->ds->err);
    if (sort_cells) OM_TRY(sort_pairs(h, ckeys, ckeys2, cvals, cperm, C, bits_for(N)));
    OM_LAUNCH(h, k_build_cells4, om_grid(C, B), B, tmp3, cperm, C, h->cells);
  }
  om_free(h, tmp3);  // (stream-ordered frees: no synchronisation needed)
  om_free(h, ckeys);
  om_free(h, ckeys2);
  om_free(h, cvals);
  om_free(h, cperm);
  OM_TRY(om_fetch_scalars(h));
  OM_TRY(om_check_dev_err(h));

  // half-edge twins
  if (C > 0) {
    const int64_t M = 3 * C;
    const int bits = bits_for(N);
    unsigned long long *ek = nullptr, *ek2 = nullptr;
    int *ev = nullptr, *ev2 = nullptr;
    CUDA_TRY(om_malloc(h, &ek, 8 * M));
    CUDA_TRY(om_malloc(h, &ek2, 8 * M));
    CUDA_TRY(om_malloc(h, &ev, 4 * M));
    CUDA_TRY(om_malloc(h, &ev2, 4 * M));
    OM_LAUNCH(h, k_edge_keys, om_grid(C, B), B, h->cells, C, bits, ek, ev);
    OM_TRY(sort_pairs(h, ek, ek2, ev, ev2, M, 2 * bits));
    OM_LAUNCH(h, k_pair_twins, om_grid(M, B), B, ek2, ev2, M, bits, (int*)h->adj, h->bflag,
              &h->ds->err);
    om_free(h, ek);
    om_free(h, ek2);
    om_free(h, ev);
    om_free(h, ev2);
  }
  OM_LAUNCH(h, k_fill_int, om_grid(N, B), B, h->v2c, N, OM_NONE_CELL);
  if (C > 0) OM_LAUNCH(h, k_v2c, om_grid(C, B), B, h->cells, C, h->v2c);
  if (N > 0) {
    CUDA_TRY(om_malloc(h, &h->ring, sizeof(int) * OM_RING_W * N));
    CUDA_TRY(om_malloc(h, &h->ringc, sizeof(int) * OM_RING_W * N));
    CUDA_TRY(om_malloc(h, &h->dirty, sizeof(int) * N));
    CUDA_TRY(om_malloc(h, &h->dirty_epoch, sizeof(int) * N));
    CUDA_TRY(om_malloc(h, &h->diff2, sizeof(double) * N));
    CUDA_TRY(om_malloc(h, &h->vflags, sizeof(unsigned short) * (N + 16)));
    CUDA_TRY(cudaMemsetAsync(h->diff2, 0, sizeof(double) * N, h->stream));
    CUDA_TRY(cudaMemsetAsync(h->vflags, 0, sizeof(unsigned short) * (N + 16), h->stream));
    CUDA_TRY(cudaMemsetAsync(h->dirty_epoch, 0, sizeof(int) * N, h->stream));
    OM_TRY(om_rebuild_rings(h, true));
  }
  OM_TRY(om_fetch_scalars(h));
  OM_TRY(om_check_dev_err(h));
  return OM_OK;
}
