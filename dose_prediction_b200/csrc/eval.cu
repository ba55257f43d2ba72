// On-device evaluation of a predicted dose volume: what LinkedNet.test_step / Pyfer.test_step do
// on the CPU with numpy after a D2H copy (train_light_pyfer.py:210-216, train_light_linked_model.py:171-176,
// DosePrediction/Evaluate/evaluate_openKBP.py:17-81,149-222):
//   - post-processing: pred[mask < 1 or pred < 0] = 0, x70 Gy;
//   - dose score: mean |pred - gt| inside the possible-dose mask (get_3D_Dose_dif, :42-48);
//   - IVS at the 101 isodose levels linspace(0, 70, 101) (IVS, :17-39) from three cumulative histograms;
//   - DVH metrics per structure (get_DVH_metrics, :51-81): D1 / D95 / D99 / mean for targets, D_0.1cc / mean for
//     OARs.  np.percentile (linear interpolation) needs exact order statistics of the ROI doses: a 3-level
//     radix select (11 + 11 + 10 bits of the order-preserving uint32 image of the float) run for all
//     (structure, prediction|ground truth, rank) targets at once: three passes over the volume, each followed by
//     a tiny scan kernel that narrows every target's key prefix.  No sort, no host round trip.
#include "common.cuh"
#include "dose_b200.h"

namespace dp {

constexpr int kMaxStruct = 16;
constexpr int kSlots = 4;                       // quantiles per (structure, source): D1, D95, D99 | D_0.1cc
constexpr int kTargetsPer = kSlots * 2;         // lower and upper order statistic of every quantile

__device__ __forceinline__ uint32_t order_key(float f) {      // monotone float -> uint32
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
  const uint32_t u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
  return __uint_as_float(u);
}

__global__ void __launch_bounds__(256) dose_postprocess_kernel(const float* pred, const float* mask, long long n, float scale,
                                                               float* out) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float p = pred[i];
  out[i] = (mask[i] < 1.f || p < 0.f) ? 0.f : p * scale;
}

// index of the highest isodose level <= v (levels ascending), or -1
__device__ __forceinline__ int level_index(float v, const double* levels, int n_levels) {
  const double x = static_cast<double>(v);
  if (!(x >= levels[0])) return -1;
  const double step = (levels[n_levels - 1] - levels[0]) / (n_levels - 1);
  int i = static_cast<int>((x - levels[0]) / step);
  i = max(0, min(n_levels - 1, i));
  while (i + 1 < n_levels && levels[i + 1] <= x) ++i;
  while (i > 0 && levels[i] > x) --i;
  return i;
}

// acc[0] += sum |pred-gt| over mask>0, acc[1] += count; hist[0|1|2][n_levels+1]: pred, gt, min(pred, gt) level index (+1)
__global__ void __launch_bounds__(256) dose_stats_kernel(const float* pred, const float* gt, const float* mask, long long n,
                                                         const double* levels, int n_levels, double* acc,
                                                         unsigned long long* hist) {
  extern __shared__ unsigned int sh[];          // [3][n_levels + 1]
  const int nb = n_levels + 1;
  for (int i = threadIdx.x; i < 3 * nb; i += blockDim.x) sh[i] = 0u;
  __syncthreads();
  float a = 0.f, c = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float p = pred[i], g = gt[i];
    if (mask[i] > 0.f) { a += fabsf(p - g); c += 1.f; }
    const int ip = level_index(p, levels, n_levels), ig = level_index(g, levels, n_levels);
    atomicAdd(&sh[ip + 1], 1u);
    atomicAdd(&sh[nb + ig + 1], 1u);
    atomicAdd(&sh[2 * nb + min(ip, ig) + 1], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * nb; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], static_cast<unsigned long long>(sh[i]));
  __shared__ float red[2][8];
  const float sa = warp_sum(a), sc = warp_sum(c);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sa; red[1][threadIdx.x >> 5] = sc; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += static_cast<double>(red[threadIdx.x][w]);
    atomicAdd(&acc[threadIdx.x], t);
  }
}
// ivs[l] = 2 #(pred>=L and gt>=L) / (#(pred>=L) + #(gt>=L));  dose_dif = acc[0]/acc[1]
__global__ void dose_stats_finalize_kernel(const unsigned long long* hist, int n_levels, const double* acc, float* ivs,
                                           float* dose_dif) {
  const int nb = n_levels + 1;
  unsigned long long cp = 0, cg = 0, cb = 0;
  for (int l = n_levels - 1; l >= 0; --l) {
    cp += hist[l + 1]; cg += hist[nb + l + 1]; cb += hist[2 * nb + l + 1];
    ivs[l] = static_cast<float>(2.0 * static_cast<double>(cb) / static_cast<double>(cp + cg));
  }
  *dose_dif = static_cast<float>(acc[0] / acc[1]);
}

// ------------------------------------------------------------------ DVH: multi-target radix select
struct DvhState {
  unsigned int roi_n[kMaxStruct];
  double roi_sum[kMaxStruct][2];
  unsigned int prefix[kMaxStruct][2][kTargetsPer];
  unsigned int rank[kMaxStruct][2][kTargetsPer];
  double frac[kMaxStruct][kSlots];
  int n_slots[kMaxStruct];
  unsigned int blocks_done;
};

// pass 0: ROI sizes, ROI dose sums, histogram of the top 11 key bits; pass 1 / 2: next 11 / last 10 bits per target
__global__ void __launch_bounds__(256) dvh_pass_kernel(const float* pred, const float* gt, const float* masks, int n_struct,
                                                       long long vox, int pass, DvhState* st, unsigned int* hist) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < vox;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float d[2] = {pred[i], gt[i]};
    const uint32_t key[2] = {order_key(d[0]), order_key(d[1])};
    for (int s = 0; s < n_struct; ++s) {
      if (!(masks[static_cast<size_t>(s) * vox + i] > 0.f)) continue;
      if (pass == 0) {
        atomicAdd(&st->roi_n[s], 1u);
        atomicAdd(&st->roi_sum[s][0], static_cast<double>(d[0]));
        atomicAdd(&st->roi_sum[s][1], static_cast<double>(d[1]));
        atomicAdd(&hist[(s * 2 + 0) * 2048 + (key[0] >> 21)], 1u);
        atomicAdd(&hist[(s * 2 + 1) * 2048 + (key[1] >> 21)], 1u);
      } else {
        const int nt = st->n_slots[s] * 2;
        for (int w = 0; w < 2; ++w)
          for (int t = 0; t < nt; ++t) {
            const uint32_t pre = st->prefix[s][w][t];
            if (pass == 1) {
              if ((key[w] >> 21) == pre) atomicAdd(&hist[((s * 2 + w) * kTargetsPer + t) * 2048 + ((key[w] >> 10) & 2047u)], 1u);
            } else {
              if ((key[w] >> 10) == pre) atomicAdd(&hist[((s * 2 + w) * kTargetsPer + t) * 1024 + (key[w] & 1023u)], 1u);
            }
          }
      }
    }
  }
}

// after pass 0: quantile positions -> ranks (np.percentile 'linear': pos = q/100 (n-1)), first digit of every target.
// after pass 1 / 2: next digit.  After pass 2 the prefix is the full key; metrics are written:
//   out[s][which][0..3] = the quantile values (targets: D1, D95, D99; OAR: D_0.1cc), out[s][which][4] = mean
__global__ void dvh_scan_kernel(int n_struct, int pass, const int* is_target, float voxels_in_tenth_of_cc, DvhState* st,
                                const unsigned int* hist, float* out, float* dvh_dif) {
  const int s = blockIdx.x, w = threadIdx.x >> 3, t = threadIdx.x & 7;       // 16 threads: (which, target)
  if (s < n_struct && threadIdx.x < 16) {
    const unsigned int n = st->roi_n[s];
    if (pass == 0) {
      const int ns = is_target[s] ? 3 : 1;
      if (threadIdx.x == 0) st->n_slots[s] = n ? ns : 0;
      const int slot = t >> 1;
      if (n && slot < ns) {
        double q;
        if (is_target[s]) q = slot == 0 ? 99.0 : (slot == 1 ? 5.0 : 1.0);
        else q = 100.0 - static_cast<double>(voxels_in_tenth_of_cc) / n * 100.0;
        q = fmin(fmax(q, 0.0), 100.0);
        const double pos = q / 100.0 * (static_cast<double>(n) - 1.0);
        const double lo = floor(pos);
        unsigned int r = static_cast<unsigned int>(lo) + (t & 1);
        if (r > n - 1) r = n - 1;
        if (w == 0 && (t & 1) == 0) st->frac[s][slot] = pos - lo;
        const unsigned int* h = hist + (s * 2 + w) * 2048;
        unsigned int cum = 0, b = 0;
        for (; b < 2048; ++b) { if (cum + h[b] > r) break; cum += h[b]; }
        st->prefix[s][w][t] = b;
        st->rank[s][w][t] = r - cum;
      }
    } else if (t < st->n_slots[s] * 2) {
      const int bins = pass == 1 ? 2048 : 1024;
      const unsigned int* h = hist + ((s * 2 + w) * kTargetsPer + t) * bins;
      const unsigned int r = st->rank[s][w][t];
      unsigned int cum = 0, b = 0;
      for (; b < static_cast<unsigned int>(bins); ++b) { if (cum + h[b] > r) break; cum += h[b]; }
      st->prefix[s][w][t] = (st->prefix[s][w][t] << (pass == 1 ? 11 : 10)) | b;
      st->rank[s][w][t] = r - cum;
    }
  }
  if (pass != 2) return;
  __syncthreads();
  if (s < n_struct && threadIdx.x < 2) {
    const int which = threadIdx.x;
    const int ns = st->n_slots[s];
    float* o = out + (s * 2 + which) * 5;
    for (int j = 0; j < 4; ++j) o[j] = 0.f;
    for (int slot = 0; slot < ns; ++slot) {
      const float a = key_to_float(st->prefix[s][which][2 * slot]), b = key_to_float(st->prefix[s][which][2 * slot + 1]);
      const double g = st->frac[s][slot];
      const double diff = static_cast<double>(b) - a;
      o[slot] = static_cast<float>(g >= 0.5 ? b - diff * (1.0 - g) : a + diff * g);      // numpy _lerp
    }
    o[4] = st->roi_n[s] ? static_cast<float>(st->roi_sum[s][which] / st->roi_n[s]) : 0.f;
  }
  // mean |gt - pred| over all metrics of all delineated structures (evaluate_openKBP.py:206-222): last block to finish
  __shared__ bool last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    last = atomicAdd(&st->blocks_done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double sum = 0.0;
    int cnt = 0;
    for (int q = 0; q < n_struct; ++q) {
      const int ns = st->n_slots[q];
      if (!ns) continue;
      volatile const float* op = out + (q * 2 + 0) * 5;
      volatile const float* og = out + (q * 2 + 1) * 5;
      for (int j = 0; j < ns; ++j) { sum += fabs(static_cast<double>(og[j]) - op[j]); ++cnt; }
      sum += fabs(static_cast<double>(og[4]) - op[4]);
      ++cnt;
    }
    *dvh_dif = cnt ? static_cast<float>(sum / cnt) : 0.f;
  }
}

// ------------------------------------------------------------------ seg validation: Dice per (volume, class)
// monai 0.7.0 DiceMetric(include_background=False) on post_pred = one-hot(argmax) vs the label map
// (OARSegmentation/train_light_transeg.py:199-216, config.py:69-70): counts[n][c] = {|pred==c & label==c|, |label==c|, |pred==c|}
__global__ void __launch_bounds__(256) dice_counts_kernel(const float* logits, const float* label, int C, long long vox,
                                                          unsigned long long* counts) {
  __shared__ unsigned int sh[16][3];
  const int n = blockIdx.y;
  if (threadIdx.x < 48) (&sh[0][0])[threadIdx.x] = 0u;
  __syncthreads();
  for (long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; v < vox;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float* z = logits + static_cast<size_t>(n) * C * vox + v;
    int best = 0;
    float bv = z[0];
    for (int c = 1; c < C; ++c) {
      const float x = z[static_cast<size_t>(c) * vox];
      if (x > bv) { bv = x; best = c; }                    // first maximum wins (torch.argmax)
    }
    const int lab = static_cast<int>(label[static_cast<size_t>(n) * vox + v]);
    atomicAdd(&sh[best][2], 1u);
    if (lab >= 0 && lab < C) {
      atomicAdd(&sh[lab][1], 1u);
      if (lab == best) atomicAdd(&sh[lab][0], 1u);
    }
  }
  __syncthreads();
  if (threadIdx.x < C * 3) {
    const unsigned int val = (&sh[0][0])[threadIdx.x];
    if (val) atomicAdd(&counts[static_cast<size_t>(n) * 48 + threadIdx.x], static_cast<unsigned long long>(val));
  }
}
// dice[n][c] = 2 I / (G + P) (NaN when the class is absent from the label, like monai); mean over the non-NaN foreground entries
// monai 0.7.0 do_metric_reduction(f, "mean"): NaN entries (class absent from the label) are ignored; first the mean over
// the classes of each sample, then the mean over the samples that have at least one class; 0 when there is none.
__global__ void dice_finalize_kernel(const unsigned long long* counts, int N, int C, float* dice, float* mean_dice) {
  double batch_sum = 0.0;
  int batch_cnt = 0;
  for (int n = 0; n < N; ++n) {
    double sum = 0.0;
    int cnt = 0;
    for (int c = 0; c < C; ++c) {
      const unsigned long long* q = counts + static_cast<size_t>(n) * 48 + c * 3;
      float d = nanf("");
      if (q[1] > 0) d = static_cast<float>(2.0 * static_cast<double>(q[0]) / static_cast<double>(q[1] + q[2]));
      dice[n * C + c] = d;
      if (c >= 1 && q[1] > 0) { sum += d; ++cnt; }
    }
    if (cnt) { batch_sum += sum / cnt; ++batch_cnt; }
  }
  *mean_dice = batch_cnt ? static_cast<float>(batch_sum / batch_cnt) : 0.f;
}

// ------------------------------------------------------------------ 95th-percentile Hausdorff distance (seg val / test)
// monai 0.7.0 HausdorffDistanceMetric(include_background=False, percentile=95) on one-hot(argmax) vs one-hot(label)
// (OARSegmentation/train_light_transeg.py:158-166,199-216): per class, the surface voxels of both masks
// (mask ^ binary_erosion(mask), 6-neighbourhood, outside = background), the Euclidean distance of every surface voxel of one
// mask to the nearest surface voxel of the other (monai: distance_transform_edt), np.percentile(.., 95) in both directions,
// the larger of the two.  Surfaces are a few 10^4 voxels, so the exact nearest-surface search is done by brute force over
// compacted coordinate lists; squared distances are integers, so a histogram over d^2 gives exact order statistics.
__global__ void __launch_bounds__(256)
hd_edges_kernel(const float* __restrict__ logits, const float* __restrict__ label, int C, int D, int H, int W, int cap,
                int* coords /* [2][C][cap] packed */, int* counts /* [2][C] */) {
  const long long vox = static_cast<long long>(D) * H * W;
  const long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (v >= vox) return;
  const int w = static_cast<int>(v % W), h = static_cast<int>((v / W) % H), d = static_cast<int>(v / (static_cast<long long>(W) * H));
  auto cls_pred = [&](long long u) {
    int best = 0;
    float bv = logits[u];
    for (int c = 1; c < C; ++c) { const float x = logits[c * vox + u]; if (x > bv) { bv = x; best = c; } }
    return best;
  };
  auto cls_gt = [&](long long u) { return static_cast<int>(label[u]); };
  for (int which = 0; which < 2; ++which) {
    const int c = which == 0 ? cls_pred(v) : cls_gt(v);
    if (c < 1 || c >= C) continue;
    bool edge = (d == 0 || d == D - 1 || h == 0 || h == H - 1 || w == 0 || w == W - 1);
    const long long nb[6] = {v - static_cast<long long>(W) * H, v + static_cast<long long>(W) * H, v - W, v + W, v - 1, v + 1};
    for (int q = 0; q < 6 && !edge; ++q) edge = (which == 0 ? cls_pred(nb[q]) : cls_gt(nb[q])) != c;
    if (edge) {
      const int slot = atomicAdd(&counts[which * C + c], 1);
      if (slot < cap) coords[(static_cast<size_t>(which) * C + c) * cap + slot] = (d << 20) | (h << 10) | w;
    }
  }
}
// grid: (source-point blocks, class, direction); direction 0 = pred surface -> gt surface, 1 = gt -> pred
__global__ void __launch_bounds__(256)
hd_dist_kernel(const int* __restrict__ coords, const int* __restrict__ counts, int C, int cap, int nbins, int* hist /* [2][C][nbins] */) {
  const int c = blockIdx.y, dir = blockIdx.z;
  if (c < 1) return;
  const int na = min(counts[dir * C + c], cap), nb = min(counts[(1 - dir) * C + c], cap);
  const int* A = coords + (static_cast<size_t>(dir) * C + c) * cap;
  const int* B = coords + (static_cast<size_t>(1 - dir) * C + c) * cap;
  __shared__ int tile[1024];
  for (int base = blockIdx.x * blockDim.x; base < na; base += gridDim.x * blockDim.x) {
    const int i = base + threadIdx.x;
    const int p = i < na ? A[i] : 0;
    const int pd = p >> 20, ph = (p >> 10) & 1023, pw = p & 1023;
    int best = 0x7fffffff;
    for (int t0 = 0; t0 < nb; t0 += 1024) {
      __syncthreads();
      for (int j = threadIdx.x; j < 1024 && t0 + j < nb; j += blockDim.x) tile[j] = B[t0 + j];
      __syncthreads();
      const int lim = min(1024, nb - t0);
      for (int j = 0; j < lim; ++j) {
        const int q = tile[j];
        const int dd = pd - (q >> 20), dh = ph - ((q >> 10) & 1023), dw = pw - (q & 1023);
        best = min(best, dd * dd + dh * dh + dw * dw);
      }
    }
    if (i < na && nb > 0) atomicAdd(&hist[(static_cast<size_t>(dir) * C + c) * nbins + min(best, nbins - 1)], 1);
  }
}
// one thread per (class, direction): np.percentile(distances, pct) with linear interpolation from the d^2 histogram
__global__ void hd_percentile_kernel(const int* __restrict__ hist, const int* __restrict__ counts, int C, int cap, int nbins, float pct,
                                     float* hd /* [C] */) {
  const int c = threadIdx.x;
  if (c >= C) return;
  if (c < 1) { hd[c] = nanf(""); return; }
  float out[2];
  for (int dir = 0; dir < 2; ++dir) {
    const int na = min(counts[dir * C + c], cap), nb = min(counts[(1 - dir) * C + c], cap);
    if (na == 0) { out[dir] = nanf(""); continue; }                // no surface on the source side: np.nan
    if (nb == 0) { out[dir] = INFINITY; continue; }                 // nothing to measure to: inf
    const double q = static_cast<double>(na - 1) * pct / 100.0;
    const long long k0 = static_cast<long long>(floor(q)), k1 = min(static_cast<long long>(na - 1), k0 + 1);
    const double frac = q - static_cast<double>(k0);
    const int* hcd = hist + (static_cast<size_t>(dir) * C + c) * nbins;
    long long seen = 0;
    double v0 = 0.0, v1 = 0.0;
    bool got0 = false;
    for (int b = 0; b < nbins; ++b) {
      const long long nxt = seen + hcd[b];
      if (!got0 && k0 < nxt) { v0 = sqrt(static_cast<double>(b)); got0 = true; }
      if (k1 < nxt) { v1 = sqrt(static_cast<double>(b)); break; }
      seen = nxt;
    }
    out[dir] = static_cast<float>(v0 + (v1 - v0) * frac);
  }
  // max(distance_1, distance_2) with numpy semantics: nan propagates
  hd[c] = (isnan(out[0]) || isnan(out[1])) ? nanf("") : fmaxf(out[0], out[1]);
}

static inline unsigned eblk(long long n, int threads, unsigned cap) {
  const long long b = (n + threads - 1) / threads;
  return static_cast<unsigned>(b < cap ? b : cap);
}

}  // namespace dp

using namespace dp;

extern "C" int dp_dose_postprocess(const float* pred, const float* mask, long long n, float scale, float* out,
                                   cudaStream_t stream) {
  dose_postprocess_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(pred, mask, n, scale, out);
  return check_cuda(cudaGetLastError(), "dose_postprocess");
}

extern "C" int dp_dose_stats(const float* pred, const float* gt, const float* mask, long long n, const double* levels,
                             int n_levels, double* acc, unsigned long long* hist, float* ivs, float* dose_dif,
                             cudaStream_t stream) {
  DP_REQUIRE(n_levels >= 2 && n_levels <= 1024, "dose_stats: 2..1024 isodose levels");
  DP_CHECK(cudaMemsetAsync(acc, 0, 2 * sizeof(double), stream));
  DP_CHECK(cudaMemsetAsync(hist, 0, 3 * static_cast<size_t>(n_levels + 1) * sizeof(unsigned long long), stream));
  const size_t smem = 3 * static_cast<size_t>(n_levels + 1) * sizeof(unsigned int);
  dose_stats_kernel<<<eblk(n, 256, 8 * 148), 256, smem, stream>>>(pred, gt, mask, n, levels, n_levels, acc, hist);
  DP_CHECK(cudaGetLastError());
  dose_stats_finalize_kernel<<<1, 1, 0, stream>>>(hist, n_levels, acc, ivs, dose_dif);
  return check_cuda(cudaGetLastError(), "dose_stats");
}

extern "C" long long dp_dvh_workspace_bytes(void) {
  return static_cast<long long>(sizeof(DvhState)) + static_cast<long long>(kMaxStruct) * 2 * kTargetsPer * 2048 * sizeof(unsigned int);
}

extern "C" int dp_dvh_metrics(const float* pred, const float* gt, const float* masks, int n_struct, const int* is_target,
                              long long vox, float voxels_in_tenth_of_cc, void* workspace, float* out, float* dvh_dif,
                              cudaStream_t stream) {
  DP_REQUIRE(n_struct >= 1 && n_struct <= kMaxStruct, "dvh_metrics: 1..%d structures", kMaxStruct);
  DvhState* st = static_cast<DvhState*>(workspace);
  unsigned int* hist = reinterpret_cast<unsigned int*>(static_cast<char*>(workspace) + sizeof(DvhState));
  const size_t hist_bytes = static_cast<size_t>(kMaxStruct) * 2 * kTargetsPer * 2048 * sizeof(unsigned int);
  DP_CHECK(cudaMemsetAsync(st, 0, sizeof(DvhState), stream));
  const unsigned blocks = eblk(vox, 256, 8 * 148);
  for (int pass = 0; pass < 3; ++pass) {
    DP_CHECK(cudaMemsetAsync(hist, 0, hist_bytes, stream));
    dvh_pass_kernel<<<blocks, 256, 0, stream>>>(pred, gt, masks, n_struct, vox, pass, st, hist);
    DP_CHECK(cudaGetLastError());
    dvh_scan_kernel<<<n_struct, 32, 0, stream>>>(n_struct, pass, is_target, voxels_in_tenth_of_cc, st, hist, out, dvh_dif);
    DP_CHECK(cudaGetLastError());
  }
  return 0;
}

extern "C" int dp_dice_metric(const float* logits, const float* label, int N, int C, long long vox, unsigned long long* counts,
                              float* dice, float* mean_dice, cudaStream_t stream) {
  DP_REQUIRE(C >= 2 && C <= 16, "dice_metric: 2..16 classes");
  DP_CHECK(cudaMemsetAsync(counts, 0, static_cast<size_t>(N) * 48 * sizeof(unsigned long long), stream));
  dim3 grid(eblk(vox, 256, 4 * 148), static_cast<unsigned>(N));
  dice_counts_kernel<<<grid, 256, 0, stream>>>(logits, label, C, vox, counts);
  DP_CHECK(cudaGetLastError());
  dice_finalize_kernel<<<1, 1, 0, stream>>>(counts, N, C, dice, mean_dice);
  return check_cuda(cudaGetLastError(), "dice_metric");
}

extern "C" long long dp_hd95_workspace_bytes(int C, int D, int H, int W) {
  const long long vox = static_cast<long long>(D) * H * W;
  const long long cap = vox / 4 + 1024;
  const long long nbins = static_cast<long long>(D - 1) * (D - 1) + static_cast<long long>(H - 1) * (H - 1) +
                          static_cast<long long>(W - 1) * (W - 1) + 2;
  return (2LL * C * cap + 2LL * C + 2LL * C * nbins) * 4;
}

extern "C" int dp_hd95(const float* logits, const float* label, int C, int D, int H, int W, float percentile, void* workspace,
                       float* hd, cudaStream_t stream) {
  DP_REQUIRE(C >= 2 && C <= 16 && D <= 1023 && H <= 1023 && W <= 1023, "hd95: 2..16 classes, dims <= 1023");
  DP_REQUIRE(percentile >= 0.f && percentile <= 100.f, "hd95: percentile should be within [0, 100], got %f", percentile);
  const long long vox = static_cast<long long>(D) * H * W;
  const int cap = static_cast<int>(vox / 4 + 1024);
  const int nbins = (D - 1) * (D - 1) + (H - 1) * (H - 1) + (W - 1) * (W - 1) + 2;
  int* coords = static_cast<int*>(workspace);
  int* counts = coords + 2LL * C * cap;
  int* hist = counts + 2 * C;
  DP_CHECK(cudaMemsetAsync(counts, 0, (2LL * C + 2LL * C * nbins) * sizeof(int), stream));
  hd_edges_kernel<<<static_cast<unsigned>((vox + 255) / 256), 256, 0, stream>>>(logits, label, C, D, H, W, cap, coords, counts);
  DP_CHECK(cudaGetLastError());
  hd_dist_kernel<<<dim3(64, C, 2), 256, 0, stream>>>(coords, counts, C, cap, nbins, hist);
  DP_CHECK(cudaGetLastError());
  hd_percentile_kernel<<<1, 32, 0, stream>>>(hist, counts, C, cap, nbins, percentile, hd);
  return check_cuda(cudaGetLastError(), "hd95");
}
