// K6: mesh statistics -- angle histogram (72 bins of 2.5 deg) and quality histogram
// (q = 2 r_in / r_circ, 40 bins of 0.025) plus min/avg/max/std.
// Replaces optimesh.helpers.print_stats (/root/reference/README.md:55-60, :203-204;
// SURVEY.md A.3, A.11).  Bin edges are those of numpy.histogram over
// linspace(0, 180, 73) and linspace(0, 1, 41): left-closed bins, last bin closed.
#include <vector>

#include "common.cuh"
#include "geom.cuh"

namespace {

constexpr int NB_A = 72, NB_Q = 40;
constexpr int MAX_BLOCKS = 2048;

__device__ __forceinline__ int bin_of(double x, double step, int nb, double last_edge) {
  // numpy: edges[k] = k * step (edges[nb] = last_edge); bin k holds edges[k] <= x < edges[k+1]
  if (!(x >= 0.0) || x > last_edge) return -1;
  int b = (int)(x / step);
  if (b > nb - 1) b = nb - 1;
  while (b > 0 && x < b * step) b--;
  while (b < nb - 1 && x >= (b + 1) * step) b++;
  return b;
}

template <int D>
__global__ void __launch_bounds__(256)
    k_stats(const double* __restrict__ x, const int4* __restrict__ cells, int C,
            unsigned long long* __restrict__ hist, double* __restrict__ partials, DevScalars* ds) {
  __shared__ unsigned int sh[NB_A + NB_Q];
  __shared__ double red[8][8];
  for (int i = threadIdx.x; i < NB_A + NB_Q; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  // partial: angle min, max, sum, sumsq; q min, max, sum, sumsq
  double amin = INFINITY, amax = -INFINITY, asum = 0, asq = 0;
  double qmin = INFINITY, qmax = -INFINITY, qsum = 0, qsq = 0;
  const double qstep = 1.0 / 40.0;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
    int4 cl = cells[c];
    Vec<D> P0 = ld_point<D>(x, cl.x), P1 = ld_point<D>(x, cl.y), P2 = ld_point<D>(x, cl.z);
    CellGeo<D> g = cell_geo<D>(P0, P1, P2);
    if (!(g.vol2 > 0.0)) {
      atomicOr(&ds->err, OM_DEV_DEGENERATE);
      continue;
    }
    const double l0 = sqrt(g.ee0), l1 = sqrt(g.ee1), l2 = sqrt(g.ee2);
    const double cosv[3] = {-g.ed0 / (l1 * l2), -g.ed1 / (l2 * l0), -g.ed2 / (l0 * l1)};
#pragma unroll
    for (int k = 0; k < 3; k++) {
      double cv = fmin(1.0, fmax(-1.0, cosv[k]));
      double ang = acos(cv) / 3.141592653589793 * 180.0;
      int b = bin_of(ang, 2.5, NB_A, 180.0);
      if (b >= 0) atomicAdd(&sh[b], 1u);
      amin = fmin(amin, ang);
      amax = fmax(amax, ang);
      asum += ang;
      asq += ang * ang;
    }
    const double q = (-l0 + l1 + l2) * (l0 - l1 + l2) * (l0 + l1 - l2) / (l0 * l1 * l2);
    int b = bin_of(q, qstep, NB_Q, 1.0);
    if (b >= 0) atomicAdd(&sh[NB_A + b], 1u);
    qmin = fmin(qmin, q);
    qmax = fmax(qmax, q);
    qsum += q;
    qsq += q * q;
  }
  double vals[8] = {amin, amax, asum, asq, qmin, qmax, qsum, qsq};
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    double v = vals[i];
    for (int o = 16; o > 0; o >>= 1) {
      double w = __shfl_xor_sync(0xffffffffu, v, o);
      if (i == 0 || i == 4)
        v = fmin(v, w);
      else if (i == 1 || i == 5)
        v = fmax(v, w);
      else
        v = v + w;
    }
    if (lane == 0) red[i][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    const int i = threadIdx.x;
    double v = red[i][0];
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) {
      double u = red[i][w];
      if (i == 0 || i == 4)
        v = fmin(v, u);
      else if (i == 1 || i == 5)
        v = fmax(v, u);
      else
        v = v + u;
    }
    partials[8 * blockIdx.x + i] = v;
  }
  for (int i = threadIdx.x; i < NB_A + NB_Q; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}

}  // namespace

int om_stats_impl(om_handle* h, int64_t* angle_hist72, int64_t* q_hist40, double* summary8) {
  const int C = (int)h->C;
  const int B = 256;
  int G = std::min(om_grid(std::max<int64_t>(C, 1), B), MAX_BLOCKS);
  unsigned long long* hist = nullptr;
  CUDA_TRY(om_malloc(h, &hist, sizeof(unsigned long long) * (NB_A + NB_Q)));
  CUDA_TRY(cudaMemsetAsync(hist, 0, sizeof(unsigned long long) * (NB_A + NB_Q), h->stream));
  if (h->D == 2)
    OM_LAUNCH(h, k_stats<2>, G, B, h->x, h->cells, C, hist, h->partials, h->ds);
  else
    OM_LAUNCH(h, k_stats<3>, G, B, h->x, h->cells, C, hist, h->partials, h->ds);
  CUDA_TRY(cudaGetLastError());
  std::vector<unsigned long long> hh(NB_A + NB_Q);
  std::vector<double> part(8 * (size_t)G);
  CUDA_TRY(cudaMemcpyAsync(hh.data(), hist, sizeof(unsigned long long) * hh.size(),
                           cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaMemcpyAsync(part.data(), h->partials, sizeof(double) * part.size(),
                           cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  om_free(h, hist);
  OM_TRY(om_fetch_scalars(h));
  OM_TRY(om_check_dev_err(h));
  for (int i = 0; i < NB_A; i++) angle_hist72[i] = (int64_t)hh[i];
  for (int i = 0; i < NB_Q; i++) q_hist40[i] = (int64_t)hh[NB_A + i];
  double amin = INFINITY, amax = -INFINITY, asum = 0, asq = 0, qmin = INFINITY, qmax = -INFINITY,
         qsum = 0, qsq = 0;
  for (int b = 0; b < G; b++) {  // fixed order: deterministic
    const double* p = &part[8 * (size_t)b];
    amin = std::min(amin, p[0]);
    amax = std::max(amax, p[1]);
    asum += p[2];
    asq += p[3];
    qmin = std::min(qmin, p[4]);
    qmax = std::max(qmax, p[5]);
    qsum += p[6];
    qsq += p[7];
  }
  const double na = 3.0 * C, nq = (double)C;
  const double aavg = C ? asum / na : 0.0, qavg = C ? qsum / nq : 0.0;
  summary8[0] = amin;
  summary8[1] = amax;
  summary8[2] = aavg;
  summary8[3] = C ? sqrt(std::max(0.0, asq / na - aavg * aavg)) : 0.0;
  summary8[4] = qmin;
  summary8[5] = qavg;
  summary8[6] = qmax;
  summary8[7] = C ? sqrt(std::max(0.0, qsq / nq - qavg * qavg)) : 0.0;
  return OM_OK;
}
