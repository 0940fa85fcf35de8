// Stage 5 (three_sigma / interloper_removal, fof.py:396-516) and stage 6 (virial mass estimator,
// analysis/mass_estimator.py:36-112,148-184; mass_analysis.py:8-44) over label-sorted members
// (kernels K12, K13 of DESIGN.md).
#include <cub/cub.cuh>

#include <random>

#include "clean.cuh"

namespace fofb {

static unsigned char* cub_tmp(fofb_handle* h, size_t bytes) { return h->cub_tmp.alloc(bytes + 256); }

// owner cluster of member t: last k with off[k] <= t
__device__ __forceinline__ int owner_of(const long long* off, int K, long long t) {
  int lo = 0, hi = K;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (off[mid] <= t) lo = mid; else hi = mid;
  }
  return lo;
}

// ---- numpy's pairwise summation (add.reduce on a contiguous float64 vector), serial -------------
// Same blocking as numpy/core/src/umath/loops_utils.h.src pairwise_sum: < 8 plain loop, <= 128 eight
// accumulators, else split in halves rounded to a multiple of 8.
template <class F>
__device__ __forceinline__ double np_pairwise_leaf(F get, long long lo, long long n) {  // n <= 128
  if (n < 8) {
    double res = 0.0;
    for (long long i = 0; i < n; ++i) res = __dadd_rn(res, get(lo + i));
    return res;
  }
  double r[8];
  for (int j = 0; j < 8; ++j) r[j] = get(lo + j);
  long long i;
  for (i = 8; i < n - (n % 8); i += 8)
    for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], get(lo + i + j));
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __dadd_rn(res, get(lo + i));
  return res;
}

// numpy's recursion on halves (n2 = n/2 rounded down to a multiple of 8) as an explicit post-order stack; `leaf`
// evaluates a block of <= 128 elements.
template <class L>
__device__ double np_pairwise_tree(L leaf, long long lo, long long n) {
  if (n <= 128) return leaf(lo, n);
  long long s_lo[40], s_n[40];
  double s_left[40];
  int s_state[40];
  int sp = 0;
  s_lo[0] = lo; s_n[0] = n; s_state[0] = 0;
  double ret = 0.0;
  while (sp >= 0) {
    const long long cl = s_lo[sp], cn = s_n[sp];
    if (cn <= 128) {
      ret = leaf(cl, cn);
      --sp;
      continue;
    }
    long long n2 = cn / 2;
    n2 -= n2 % 8;
    if (s_state[sp] == 0) {
      s_state[sp] = 1;
      ++sp;
      s_lo[sp] = cl; s_n[sp] = n2; s_state[sp] = 0;
    } else if (s_state[sp] == 1) {
      s_left[sp] = ret;
      s_state[sp] = 2;
      ++sp;
      s_lo[sp] = cl + n2; s_n[sp] = cn - n2; s_state[sp] = 0;
    } else {
      ret = __dadd_rn(s_left[sp], ret);
      --sp;
    }
  }
  return ret;
}

template <class F>
__device__ double np_pairwise_sum(F get, long long lo, long long n) {
  return np_pairwise_tree([&](long long l, long long c) -> double { return np_pairwise_leaf(get, l, c); }, lo, n);
}

// The same sum by one warp, bit for bit: numpy's eight strided accumulators of a block are one lane each, combined as
// numpy combines them; the recursion above 128 elements is walked by every lane alike (its shape depends on n only).
// All 32 lanes must call it with the same arguments; every lane gets the result.
template <class F>
__device__ __forceinline__ double warp_np_leaf(F get, long long lo, long long n, int lane) {  // n <= 128
  if (n < 8) {
    double res = 0.0;
    for (long long i = 0; i < n; ++i) res = __dadd_rn(res, get(lo + i));
    return res;
  }
  const long long body = n - (n % 8);
  double r = 0.0;
  if (lane < 8) {
    r = get(lo + lane);
    for (long long i = 8 + lane; i < body; i += 8) r = __dadd_rn(r, get(lo + i));
  }
  const double r0 = __shfl_sync(0xffffffffu, r, 0), r1 = __shfl_sync(0xffffffffu, r, 1), r2 = __shfl_sync(0xffffffffu, r, 2),
               r3 = __shfl_sync(0xffffffffu, r, 3), r4 = __shfl_sync(0xffffffffu, r, 4), r5 = __shfl_sync(0xffffffffu, r, 5),
               r6 = __shfl_sync(0xffffffffu, r, 6), r7 = __shfl_sync(0xffffffffu, r, 7);
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r0, r1), __dadd_rn(r2, r3)), __dadd_rn(__dadd_rn(r4, r5), __dadd_rn(r6, r7)));
  for (long long i = body; i < n; ++i) res = __dadd_rn(res, get(lo + i));
  return res;
}

template <class F>
__device__ double warp_np_pairwise_sum(F get, long long n, int lane) {
  return np_pairwise_tree([&](long long l, long long c) -> double { return warp_np_leaf(get, l, c, lane); }, 0, n);
}

// The same sum by a whole CTA, bit for bit: thread 0 lists the leaves of the recursion, the threads evaluate them
// (each leaf is serial, as in numpy), thread 0 adds them up along the same tree.  Falls back to the serial walk when
// the leaf table is too small.  All threads of the block must call it; the result is returned to every thread.
constexpr int kPwLeaves = 2048;
struct PairwiseScratch {
  int lo[kPwLeaves];
  short cnt[kPwLeaves];
  double sum[kPwLeaves];
  int nleaves;
  double result;
};

template <class F>
__device__ double cta_pairwise_sum(F get, long long n, PairwiseScratch& sc) {
  const int tid = threadIdx.x;
  __syncthreads();
  if (tid == 0) {
    int nl = 0;
    np_pairwise_tree([&](long long l, long long c) -> double {
      if (nl < kPwLeaves) { sc.lo[nl] = (int)l; sc.cnt[nl] = (short)c; }
      ++nl;
      return 0.0;
    }, 0, n);
    sc.nleaves = nl;
  }
  __syncthreads();
  const int nl = sc.nleaves;
  if (nl > kPwLeaves) {
    if (tid == 0) sc.result = np_pairwise_sum(get, 0, n);
  } else {
    for (int t = tid; t < nl; t += blockDim.x) sc.sum[t] = np_pairwise_leaf(get, sc.lo[t], sc.cnt[t]);
    __syncthreads();
    if (tid == 0) {
      int next = 0;
      sc.result = np_pairwise_tree([&](long long, long long) -> double { return sc.sum[next++]; }, 0, n);
    }
  }
  __syncthreads();
  return sc.result;
}

// ---- K12a: per member velocity and projected radius (fof.py:418-431) ---------------------------------
__global__ void k_member_vr(long long total, int K, const long long* off, View M, const double* cen, Cosmo cosmo,
                            double* vel, double* rad, double* dA_out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int k = owner_of(off, K, t);
  const double cra = cen[3 * k], cdec = cen[3 * k + 1], cz = cen[3 * k + 2];
  vel[t] = redshift_to_velocity(M.at(t, 2), cz);
  const double sep = vincenty_rad(__dmul_rn(cra, kDeg2Rad), __dmul_rn(cdec, kDeg2Rad), __dmul_rn(M.at(t, 0), kDeg2Rad),
                                  __dmul_rn(M.at(t, 1), kDeg2Rad));
  const double dA = cosmo_angular_diameter_distance(cosmo, cz);
  rad[t] = __dmul_rn(__dmul_rn(__dmul_rn(sep, kRad2Deg), dA), kDeg2Rad);
  if (t == off[k]) dA_out[k] = dA;
}

// ---- K12b: radial bins (np.linspace + np.digitize, fof.py:436-437), stable partition by bin -----------
__global__ void k_bin_members(int K, const long long* off, const double* rad, int nbins, int* perm, int* bin_start) {
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (k >= K) return;
  const long long s0 = off[k];
  const int n = (int)(off[k + 1] - s0);
  int* bs = bin_start + (long long)k * (nbins + 2);
  if (n == 0) {
    for (int b = lane; b < nbins + 2; b += 32) bs[b] = 0;
    return;
  }
  double mn = INFINITY, mx = -INFINITY;
  for (int t = lane; t < n; t += 32) {
    const double r = rad[s0 + t];
    mn = fmin(mn, r);
    mx = fmax(mx, r);
  }
  for (int o = 16; o; o >>= 1) {
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  // np.linspace(start, stop, num): step = (stop-start)/(num-1); y = arange(num)*step + start; y[-1] = stop
  const int div = nbins - 1;
  const double delta = __dsub_rn(mx, mn);
  const double step = div > 0 ? __ddiv_rn(delta, (double)div) : 0.0;
  auto edge = [&](int i) -> double {
    if (i == nbins - 1 && nbins > 1) return mx;
    if (div <= 0) return mn;
    if (step == 0.0) return __dadd_rn(__dmul_rn(__ddiv_rn((double)i, (double)div), delta), mn);
    return __dadd_rn(__dmul_rn((double)i, step), mn);
  };
  // digitize (bins increasing, right=False): index = #edges <= x
  int pos = 0;
  for (int b = 1; b <= nbins; ++b) {
    if (lane == 0) bs[b] = pos;
    for (int base = 0; base < n; base += 32) {
      const int t = base + lane;
      bool in = false;
      if (t < n) {
        const double x = rad[s0 + t];
        int cnt = 0;
        for (int i = 0; i < nbins; ++i) cnt += edge(i) <= x ? 1 : 0;
        in = cnt == b;
      }
      const unsigned m = __ballot_sync(0xffffffffu, in);
      if (in) perm[s0 + pos + __popc(m & ((1u << lane) - 1u))] = t;
      pos += __popc(m);
    }
  }
  if (lane == 0) {
    bs[0] = 0;
    bs[nbins + 1] = pos;
  }
}

// ---- K12c: iterative 3-sigma clipping of one radial bin, one thread per (cluster, bin) (fof.py:439-492)
// Bins longer than kBigBin are left to the batched path (k_big_*); clusters of at most n_smem members belong to
// k_clip_bins_smem.
constexpr int kBigBin = 256;

__global__ void k_clip_bins(int K, int nbins, const long long* off, const double* vel, const double* rad, int* perm,
                            const int* bin_start, double* dev, int* kept, int n_smem) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)K * nbins) return;
  const int k = (int)(g / nbins), b = (int)(g % nbins) + 1;
  const long long s0 = off[k];
  if (off[k + 1] - s0 <= n_smem) return;
  const int* bs = bin_start + (long long)k * (nbins + 2);
  const int lo = bs[b], hi = (b == nbins) ? bs[nbins + 1] : bs[b + 1];
  if (hi - lo > kBigBin) return;
  int* p = perm + s0 + lo;   // this bin's members (positions inside the cluster), reordered in place
  double* d = dev + s0 + lo;
  int len = hi - lo;
  bool changed = true;
  while (changed && len > 0) {
    changed = false;
    if (len > 2) {
      auto getv = [&](long long i) -> double { return vel[s0 + p[i]]; };
      const double mean = __ddiv_rn(np_pairwise_sum(getv, 0, len), (double)len);
      for (int i = 0; i < len; ++i) d[i] = fabs(__dsub_rn(vel[s0 + p[i]], mean));
      // stable insertion sort by deviation (np.argsort; ties pinned to the stable order)
      for (int i = 1; i < len; ++i) {
        const double di = d[i];
        const int pi = p[i];
        int j = i - 1;
        while (j >= 0 && d[j] > di) {
          d[j + 1] = d[j];
          p[j + 1] = p[j];
          --j;
        }
        d[j + 1] = di;
        p[j + 1] = pi;
      }
      if (rad[s0 + p[len - 1]] != 0.0) {  // the largest deviation is not the cluster centre itself
        const int m = len - 1;
        if (m > 1) {
          const double bin_mean = __ddiv_rn(np_pairwise_sum(getv, 0, m), (double)m);
          double acc = 0.0;  // builtin sum(): sequential
          for (int i = 0; i < m; ++i) {
            const double x = __dsub_rn(vel[s0 + p[i]], bin_mean);
            acc = __dadd_rn(acc, __dmul_rn(x, x));
          }
          const double disp = __ddiv_rn(acc, (double)(m - 1));
          if (fabs(__dsub_rn(vel[s0 + p[len - 1]], bin_mean)) > __dmul_rn(3.0, sqrt(disp))) {
            --len;
            changed = true;
          }
        }
      }
    }
  }
  kept[g] = len;
}

// numpy's mean of v[p[0 .. len)] by one warp, bit for bit (np.mean = pairwise add.reduce / len): below 8 elements the
// plain loop, up to 128 the eight strided accumulators -- one lane each -- combined as numpy combines them, then the
// tail; longer slices (rare: bins are cut at kBigBin) run numpy's recursion on every lane.  All lanes get the result.
__device__ __forceinline__ double warp_np_mean(const double* v, const int* p, int len, int lane) {
  double res;
  if (len < 8) {
    res = 0.0;
    for (int i = 0; i < len; ++i) res = __dadd_rn(res, v[p[i]]);
  } else if (len <= 128) {
    const int body = len - (len % 8);
    double r = 0.0;
    if (lane < 8) {
      r = v[p[lane]];
      for (int i = 8 + lane; i < body; i += 8) r = __dadd_rn(r, v[p[i]]);
    }
    const double r0 = __shfl_sync(0xffffffffu, r, 0), r1 = __shfl_sync(0xffffffffu, r, 1), r2 = __shfl_sync(0xffffffffu, r, 2),
                 r3 = __shfl_sync(0xffffffffu, r, 3), r4 = __shfl_sync(0xffffffffu, r, 4), r5 = __shfl_sync(0xffffffffu, r, 5),
                 r6 = __shfl_sync(0xffffffffu, r, 6), r7 = __shfl_sync(0xffffffffu, r, 7);
    res = __dadd_rn(__dadd_rn(__dadd_rn(r0, r1), __dadd_rn(r2, r3)), __dadd_rn(__dadd_rn(r4, r5), __dadd_rn(r6, r7)));
    for (int i = body; i < len; ++i) res = __dadd_rn(res, v[p[i]]);
  } else {
    auto getv = [&](long long i) -> double { return v[p[i]]; };
    res = np_pairwise_sum(getv, 0, len);
  }
  return __ddiv_rn(res, (double)len);
}

// The same clipping with the cluster's velocities, bin permutation and scratch in shared memory: one warp per cluster
// walks the radial bins one after the other and works on each together -- the mean (numpy's summation order, see
// warp_np_mean), the deviations, a stable rank-by-counting sort (the permutation np.argsort's stable order gives; an
// insertion sort by one thread is quadratic in the bin and leaves 31 lanes idle), the dispersion of the rest.
// Dynamic shared memory: n_max * 33 bytes (launched only when that fits).
__global__ void __launch_bounds__(32) k_clip_bins_smem(int K, int nbins, const long long* off, const double* vel, const double* rad,
                                                       int* perm, const int* bin_start, int* kept, int n_smem) {
  extern __shared__ double smem_d[];
  const int k = blockIdx.x;
  if (k >= K) return;
  const long long s0 = off[k];
  const int n = (int)(off[k + 1] - s0);
  if (n > n_smem) return;
  const int lane = threadIdx.x;
  double* sv = smem_d;                                        // velocity by position in the cluster
  double* sd = sv + n;                                        // deviation, by slot
  double* sd2 = sd + n;                                       // sort / squares scratch, by slot
  int* sp = reinterpret_cast<int*>(sd2 + n);                  // permutation (slot -> position)
  int* sp2 = sp + n;
  unsigned char* sz = reinterpret_cast<unsigned char*>(sp2 + n);  // rad == 0 (the cluster centre itself)
  for (int i = lane; i < n; i += 32) {
    sv[i] = vel[s0 + i];
    sz[i] = rad[s0 + i] == 0.0 ? 1 : 0;
    sp[i] = perm[s0 + i];
  }
  __syncwarp();
  const int* bs = bin_start + (long long)k * (nbins + 2);
  for (int b = 1; b <= nbins; ++b) {
    const int lo = bs[b], hi = (b == nbins) ? bs[nbins + 1] : bs[b + 1];
    if (hi - lo > kBigBin) continue;
    int* p = sp + lo;
    int* p2 = sp2 + lo;
    double* d = sd + lo;
    double* d2 = sd2 + lo;
    int len = hi - lo;
    bool changed = true;
    while (changed && len > 0) {
      changed = false;
      if (len > 2) {
        const double mean = warp_np_mean(sv, p, len, lane);
        for (int i = lane; i < len; i += 32) d[i] = fabs(__dsub_rn(sv[p[i]], mean));
        __syncwarp();
        for (int i = lane; i < len; i += 32) {  // stable ascending order by deviation
          const double di = d[i];
          int rank = 0;
          for (int j = 0; j < len; ++j) {
            const double dj = d[j];
            rank += (dj < di || (dj == di && j < i)) ? 1 : 0;
          }
          p2[rank] = p[i];
        }
        __syncwarp();
        for (int i = lane; i < len; i += 32) p[i] = p2[i];
        __syncwarp();
        const int worst = p[len - 1];
        if (!sz[worst]) {  // the largest deviation is not the cluster centre itself
          const int m = len - 1;
          if (m > 1) {
            const double bin_mean = warp_np_mean(sv, p, m, lane);
            for (int i = lane; i < m; i += 32) {
              const double x = __dsub_rn(sv[p[i]], bin_mean);
              d2[i] = __dmul_rn(x, x);
            }
            __syncwarp();
            double acc = 0.0;  // builtin sum(): sequential
            if (lane == 0)
              for (int i = 0; i < m; ++i) acc = __dadd_rn(acc, d2[i]);
            acc = __shfl_sync(0xffffffffu, acc, 0);
            const double disp = __ddiv_rn(acc, (double)(m - 1));
            if (fabs(__dsub_rn(sv[worst], bin_mean)) > __dmul_rn(3.0, sqrt(disp))) {
              --len;
              changed = true;
            }
          }
        }
        __syncwarp();
      }
    }
    if (lane == 0) kept[(long long)k * nbins + b - 1] = len;
  }
  __syncwarp();
  for (int i = lane; i < n; i += 32) perm[s0 + i] = sp[i];
}

// ---- K12d: long bins, all of them together, one clipping iteration per round of launches -------------------------
// A bin of 10^4..10^5 members cannot be clipped by one thread (the insertion sort alone is quadratic).  The batched
// path keeps the arithmetic of fof.py:439-492 bit for bit -- numpy's pairwise means over the members in their current
// order, the stable argsort by deviation -- but spreads it: a CTA per bin for the exact means, one segmented stable
// sort over all long bins, a CTA per bin for the 3-sigma decision.  Only the dispersion, a sequential builtin sum()
// in the reference, is reduced in parallel; it decides a comparison, and when the two sides are within 1e-9 of each
// other thread 0 repeats it in the reference's order.
__global__ void k_big_list(int K, int nbins, const long long* off, const int* bin_start, int* count, int* big_begin, int* big_len,
                           int* big_slot) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)K * nbins) return;
  const int k = (int)(g / nbins), b = (int)(g % nbins) + 1;
  const int* bs = bin_start + (long long)k * (nbins + 2);
  const int lo = bs[b], hi = (b == nbins) ? bs[nbins + 1] : bs[b + 1];
  if (hi - lo <= kBigBin) return;
  const int i = atomicAdd(count, 1);
  big_begin[i] = (int)(off[k] + lo);   // position of the bin inside perm / dev (total < 2^31)
  big_len[i] = hi - lo;
  big_slot[i] = (int)g;                // where its kept count goes
}

// exact mean of the bin in its current order, deviations, and the segment [begin, end) for the sort
__global__ void __launch_bounds__(256) k_big_mean_dev(int nb, const int* big_begin, const int* big_len, const int* big_active,
                                                      const long long* cl_off_of_pos, const double* vel, const int* perm,
                                                      double* dev, int* seg_begin, int* seg_end) {
  __shared__ PairwiseScratch sc;
  const int i = blockIdx.x;
  if (i >= nb) return;
  const int begin = big_begin[i], len = big_len[i];
  const bool act = big_active[i] && len > 2;
  if (threadIdx.x == 0) {
    seg_begin[i] = begin;
    seg_end[i] = act ? begin + len : begin;
  }
  if (!act) return;
  const long long s0 = cl_off_of_pos[i];
  const int* p = perm + begin;
  auto getv = [&](long long t) -> double { return vel[s0 + p[t]]; };
  const double mean = __ddiv_rn(cta_pairwise_sum(getv, len, sc), (double)len);
  for (int t = threadIdx.x; t < len; t += blockDim.x) dev[begin + t] = fabs(__dsub_rn(vel[s0 + p[t]], mean));
}

__global__ void __launch_bounds__(256) k_big_decide(int nb, const int* big_begin, int* big_len, int* big_active,
                                                    const long long* cl_off_of_pos, const double* vel, const double* rad,
                                                    int* perm, const int* perm_sorted, int* any_changed) {
  __shared__ PairwiseScratch sc;
  typedef cub::BlockReduce<double, 256> Reduce;
  __shared__ typename Reduce::TempStorage rtmp;
  __shared__ double s_acc;
  const int i = blockIdx.x;
  if (i >= nb) return;
  const int begin = big_begin[i], len = big_len[i];
  if (!big_active[i]) return;
  if (len <= 2) {
    if (threadIdx.x == 0) big_active[i] = 0;
    return;
  }
  const long long s0 = cl_off_of_pos[i];
  int* p = perm + begin;
  for (int t = threadIdx.x; t < len; t += blockDim.x) p[t] = perm_sorted[begin + t];
  __syncthreads();
  const int last = p[len - 1];
  bool removed = false;
  if (rad[s0 + last] != 0.0) {  // the largest deviation is not the cluster centre itself
    const int m = len - 1;
    if (m > 1) {
      auto getv = [&](long long t) -> double { return vel[s0 + p[t]]; };
      const double bin_mean = __ddiv_rn(cta_pairwise_sum(getv, m, sc), (double)m);
      double part = 0.0;
      for (int t = threadIdx.x; t < m; t += blockDim.x) {
        const double x = vel[s0 + p[t]] - bin_mean;
        part += x * x;
      }
      double acc = Reduce(rtmp).Sum(part);
      if (threadIdx.x == 0) {
        const double lhs = fabs(__dsub_rn(vel[s0 + last], bin_mean));
        double rhs = __dmul_rn(3.0, sqrt(__ddiv_rn(acc, (double)(m - 1))));
        if (fabs(lhs - rhs) <= 1e-9 * rhs) {  // too close to call: builtin sum(), sequential, as the reference
          acc = 0.0;
          for (int t = 0; t < m; ++t) {
            const double x = __dsub_rn(vel[s0 + p[t]], bin_mean);
            acc = __dadd_rn(acc, __dmul_rn(x, x));
          }
          rhs = __dmul_rn(3.0, sqrt(__ddiv_rn(acc, (double)(m - 1))));
        }
        s_acc = lhs > rhs ? 1.0 : 0.0;
      }
      __syncthreads();
      removed = s_acc != 0.0;
    }
  }
  if (threadIdx.x == 0) {
    if (removed) {
      big_len[i] = len - 1;
      *any_changed = 1;
    } else {
      big_active[i] = 0;
    }
  }
}

__global__ void k_fill_i32(int n, int* out, int v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = v;
}

__global__ void k_big_finish(int nb, const int* big_len, const int* big_slot, int* kept) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nb) kept[big_slot[i]] = big_len[i];
}

__global__ void k_big_cluster_off(int nb, int K, const long long* off, const int* big_slot, int nbins, long long* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nb) out[i] = off[big_slot[i] / nbins];
}

__global__ void k_max_cluster_len(int K, const long long* off, int* out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < K) atomicMax(out, (int)(off[k + 1] - off[k]));
}

__global__ void k_kept_per_cluster(int K, int nbins, const int* kept, long long* len) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  long long s = 0;
  for (int b = 0; b < nbins; ++b) s += kept[(long long)k * nbins + b];
  len[k] = s;
}

__global__ void k_emit_cleaned(int K, int nbins, const long long* off, const int* perm, const int* bin_start,
                               const int* kept, const long long* out_off, int* out_index) {
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (k >= K) return;
  const long long s0 = off[k];
  const int* bs = bin_start + (long long)k * (nbins + 2);
  long long dst = out_off[k];
  for (int b = 1; b <= nbins; ++b) {
    const int lo = bs[b];
    const int len = kept[(long long)k * nbins + b - 1];
    for (int t = lane; t < len; t += 32) out_index[dst + t] = perm[s0 + lo + t];
    dst += len;
  }
}

// Long bins (more than kBigBin members): see K12d above.
void CleanState::clip_big_bins(fofb_handle* h, int K, int nbins, const long long* off, long long total, const double* vel,
                               const double* rad, double* dev, int* perm, const int* bstart, int* kept) {
  cudaStream_t st = h->stream;
  FOFB_REQUIRE(total < (1ll << 31), "three_sigma: more than 2^31 members");
  const size_t cap = (size_t)(total / kBigBin) + 1;  // at most this many long bins
  int* cnt = d_big_count.alloc(2);
  int* bb = d_big_begin.alloc(cap);
  int* bl = d_big_len.alloc(cap);
  int* bsl = d_big_slot.alloc(cap);
  FOFB_CUDA(cudaMemsetAsync(cnt, 0, 2 * sizeof(int), st));
  k_big_list<<<grid_for((int64_t)K * nbins, 256), 256, 0, st>>>(K, nbins, off, bstart, cnt, bb, bl, bsl);
  FOFB_LAUNCH_CHECK(h);
  int nb = 0;
  FOFB_CUDA(cudaMemcpyAsync(&nb, cnt, sizeof(int), cudaMemcpyDeviceToHost, st));
  FOFB_CUDA(cudaStreamSynchronize(st));
  if (nb == 0) return;
  int* act = d_big_active.alloc(nb);
  int* sb = d_big_seg.alloc((size_t)nb * 2);
  int* se = sb + nb;
  long long* coff = d_big_off.alloc(nb);
  double* dev_s = d_big_dev.alloc((size_t)total);
  int* perm_s = d_big_perm.alloc((size_t)total);
  int* changed = cnt + 1;
  k_fill_i32<<<grid_for(nb, 256), 256, 0, st>>>(nb, act, 1);
  FOFB_LAUNCH_CHECK(h);
  k_big_cluster_off<<<grid_for(nb, 256), 256, 0, st>>>(nb, K, off, bsl, nbins, coff);
  FOFB_LAUNCH_CHECK(h);
  size_t bytes = 0;
  FOFB_CUDA(cub::DeviceSegmentedSort::StableSortPairs(nullptr, bytes, dev, dev_s, perm, perm_s, total, nb, sb, se, st));
  unsigned char* tmp = cub_tmp(h, bytes);
  for (int iter = 0;; ++iter) {
    FOFB_CUDA(cudaMemsetAsync(changed, 0, sizeof(int), st));
    FOFB_PROFILED(h, "k_big_mean_dev", k_big_mean_dev<<<nb, 256, 0, st>>>(nb, bb, bl, act, coff, vel, perm, dev, sb, se));
    FOFB_CUDA(cub::DeviceSegmentedSort::StableSortPairs(tmp, bytes, dev, dev_s, perm, perm_s, total, nb, sb, se, st));
    count_launch(h, 4);
    FOFB_PROFILED(h, "k_big_decide", k_big_decide<<<nb, 256, 0, st>>>(nb, bb, bl, act, coff, vel, rad, perm, perm_s, changed));
    int any = 0;
    FOFB_CUDA(cudaMemcpyAsync(&any, changed, sizeof(int), cudaMemcpyDeviceToHost, st));
    FOFB_CUDA(cudaStreamSynchronize(st));
    if (!any) break;
    FOFB_REQUIRE(iter < (1 << 24), "three_sigma: clipping did not terminate");
  }
  k_big_finish<<<grid_for(nb, 256), 256, 0, st>>>(nb, bl, bsl, kept);
  FOFB_LAUNCH_CHECK(h);
}

long long CleanState::three_sigma_core(fofb_handle* h, const fofb_params& p, int K, const long long* off, long long total,
                                       const View& M, const double* cen, int nbins) {
  cudaStream_t st = h->stream;
  FOFB_REQUIRE(K >= 0 && nbins >= 1 && nbins <= 4096, "three_sigma: bad K or n_bins");
  if (K == 0) return 0;
  FOFB_REQUIRE(total == M.rows, "three_sigma: offsets do not cover the member matrix");
  FOFB_REQUIRE(M.cols >= 3, "three_sigma: members need ra, dec, z columns");
  const int B = 256;
  const Cosmo cosmo = device_cosmo(h, p);
  const size_t tt = (size_t)std::max<long long>(total, 1);
  double* vel = d_vel.alloc(tt);
  double* rad = d_rad.alloc(tt);
  double* dev = d_dev.alloc(tt);
  int* perm = d_perm.alloc(tt);
  int* bstart = d_bin_start.alloc((size_t)K * (nbins + 2));
  int* kept = d_kept.alloc((size_t)K * nbins);
  if (total > 0) {
    FOFB_PROFILED(h, "k_member_vr", k_member_vr<<<grid_for(total, B), B, 0, st>>>(total, K, off, M, cen, cosmo, vel, rad, d_dA.alloc(K)));
  }
  FOFB_PROFILED(h, "k_bin_members", k_bin_members<<<grid_for((int64_t)K * 32, 128), 128, 0, st>>>(K, off, rad, nbins, perm, bstart));
  // largest cluster decides whether the shared-memory variant fits
  int* dmax = d_status.alloc(2) + 1;
  FOFB_CUDA(cudaMemsetAsync(dmax, 0, sizeof(int), st));
  k_max_cluster_len<<<grid_for(K, B), B, 0, st>>>(K, off, dmax);
  FOFB_LAUNCH_CHECK(h);
  int nmax = 0;
  FOFB_CUDA(cudaMemcpyAsync(&nmax, dmax, sizeof(int), cudaMemcpyDeviceToHost, st));
  FOFB_CUDA(cudaStreamSynchronize(st));
  const int smem_cap = 200 * 1024;
  const int n_smem = (smem_cap - 64) / 33;  // clusters up to this many members are clipped in shared memory
  const size_t smem = (size_t)std::min(nmax, n_smem) * 33 + 64;
  if (smem > 48 * 1024 && smem > clip_smem_set) {
    FOFB_CUDA(cudaFuncSetAttribute(k_clip_bins_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_cap));
    clip_smem_set = smem_cap;
  }
  FOFB_PROFILED(h, "k_clip_bins", k_clip_bins_smem<<<K, 32, smem, st>>>(K, nbins, off, vel, rad, perm, bstart, kept, n_smem));
  if (nmax > n_smem) {
    FOFB_PROFILED(h, "k_clip_bins_global", k_clip_bins<<<grid_for((int64_t)K * nbins, 64), 64, 0, st>>>(K, nbins, off, vel, rad, perm, bstart, dev, kept, n_smem));
  }
  if (nmax > kBigBin) clip_big_bins(h, K, nbins, off, total, vel, rad, dev, perm, bstart, kept);
  long long* len = d_len.alloc((size_t)K + 1);
  long long* ooff = d_out_off.alloc((size_t)K + 1);
  k_kept_per_cluster<<<grid_for(K, B), B, 0, st>>>(K, nbins, kept, len);
  FOFB_LAUNCH_CHECK(h);
  size_t bytes = 0;
  FOFB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, len, ooff, K, st));
  FOFB_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp(h, bytes), bytes, len, ooff, K, st));
  count_launch(h, 2);
  int* oidx = d_out_index.alloc(tt);
  k_emit_cleaned<<<grid_for((int64_t)K * 32, 128), 128, 0, st>>>(K, nbins, off, perm, bstart, kept, ooff, oidx);
  FOFB_LAUNCH_CHECK(h);
  long long last_off = 0, last_len = 0;
  FOFB_CUDA(cudaMemcpyAsync(&last_off, ooff + (K - 1), sizeof(long long), cudaMemcpyDeviceToHost, st));
  FOFB_CUDA(cudaMemcpyAsync(&last_len, len + (K - 1), sizeof(long long), cudaMemcpyDeviceToHost, st));
  FOFB_CUDA(cudaStreamSynchronize(st));
  const long long kept_total = last_off + last_len;
  FOFB_CUDA(cudaMemcpyAsync(ooff + K, &kept_total, sizeof(long long), cudaMemcpyHostToDevice, st));
  return kept_total;
}

void CleanState::three_sigma(fofb_handle* h, const fofb_params& p, int K, const long long* off_host, const View& M,
                             const double* centres_host, int nbins, long long* out_off_host, int* out_index_host) {
  cudaStream_t st = h->stream;
  out_off_host[0] = 0;
  if (K == 0) return;
  long long* off = d_off.alloc((size_t)K + 1);
  double* cen = d_cen.alloc((size_t)K * 3);
  FOFB_CUDA(cudaMemcpyAsync(off, off_host, sizeof(long long) * ((size_t)K + 1), cudaMemcpyHostToDevice, st));
  FOFB_CUDA(cudaMemcpyAsync(cen, centres_host, sizeof(double) * (size_t)K * 3, cudaMemcpyHostToDevice, st));
  const long long kept_total = three_sigma_core(h, p, K, off, off_host[K], M, cen, nbins);
  FOFB_CUDA(cudaMemcpyAsync(out_off_host, d_out_off.p, sizeof(long long) * ((size_t)K + 1), cudaMemcpyDeviceToHost, st));
  if (kept_total > 0)
    FOFB_CUDA(cudaMemcpyAsync(out_index_host, d_out_index.p, sizeof(int) * kept_total, cudaMemcpyDeviceToHost, st));
  FOFB_CUDA(cudaStreamSynchronize(st));
}

// ---- stage 6 ----------------------------------------------------------------------------------------------
// numpy legacy randint(0, n) stream after np.random.seed(1): MT19937 words, masked, rejected if > n-1
// (astropy bootstrap under NumpyRNGContext(1), mass_estimator.py:178-184).
static void mt_stream(std::vector<unsigned>& out, size_t count) {
  std::mt19937 gen(1u);
  out.resize(count);
  for (size_t i = 0; i < count; ++i) out[i] = (unsigned)gen();
}

void bootstrap_indices_host(long long n, long long count, int* out) {
  std::mt19937 gen(1u);
  unsigned mask = (unsigned)(n - 1);
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
  for (long long i = 0; i < count; ++i) {
    unsigned v;
    do { v = (unsigned)gen() & mask; } while (v > (unsigned)(n - 1));
    out[i] = (int)v;
  }
}

// per-member sin/cos of longitude and latitude for the O(n^2) pair sum
__global__ void k_member_trig(long long total, View M, double4* trig) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  double sl, cl, sp, cp;
  sincos(__dmul_rn(M.at(t, 0), kDeg2Rad), &sl, &cl);
  sincos(__dmul_rn(M.at(t, 1), kDeg2Rad), &sp, &cp);
  trig[t] = make_double4(sl, cl, sp, cp);
}

constexpr int kVirialThreads = 128;
constexpr int kVirialCache = 512;  // clusters up to this size keep z, z errors and v^2 in shared memory
constexpr int kVirialTrig = 256;   // and up to this size their members' sines and cosines too  // clusters up to this size keep their members' sines and cosines in shared memory
constexpr int kBootNum = 100;

// One CTA per cluster.
// 1 / separation [rad] of two members from Vincenty's numerator^2 (s = num1^2 + num2^2) and denominator
// (den = cos(sep)), the two quantities astropy feeds to arctan2 (App. B.2).  Members of one cluster are arc-minutes
// apart, so t = sqrt(s) / den is tiny: with eps = 1 - den, 1/den = 1 + eps + eps^2 + ...; atan t = t (1 - a),
// a = t^2/3 - t^4/5 + t^6/7; 1/(1 - a) = 1 + a + a^2 + a^3.  For t < 0.02 every truncated term is below 1e-15
// relative, and the pair costs one reciprocal square root instead of a square root, two divisions and an atan2.
// Returns a negative number when the shortcut does not apply (wide pair, or closer than 1e-6 rad where the
// angle-difference identities lose digits): the caller then evaluates the formula itself.
__device__ __forceinline__ double inv_small_separation(double s, double den) {
  if (!(den > 0.0 && s < 4.0e-4 * den * den && s >= 1.0e-12)) return -1.0;
  const double eps = 1.0 - den;
  const double inv_den = 1.0 + eps * (1.0 + eps * (1.0 + eps * (1.0 + eps)));
  const double t2 = s * inv_den * inv_den;
  const double a = t2 * (1.0 / 3.0 - t2 * (1.0 / 5.0 - t2 * (1.0 / 7.0)));
  return den * rsqrt(s) * (1.0 + a * (1.0 + a * (1.0 + a)));
}

// ---- pair sum of a giant cluster, spread over the grid ----------------------------------------------------------
// k_virial_mass gives a cluster one CTA; its O(n^2) separations are the whole cost once n reaches 10^4..10^5.  For
// clusters above kBigCluster members the sum over i < j of 1 / sep_ij [rad] is taken by one CTA per 64 rows i, then
// added up per cluster in tile order (deterministic).
constexpr int kBigCluster = 4096;
constexpr int kPairRows = 64;

__global__ void k_big_clusters(int K, const long long* off, int* count, int* list) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  if (off[k + 1] - off[k] > kBigCluster) list[atomicAdd(count, 1)] = k;
}

__global__ void __launch_bounds__(256) k_pair_sum_big(const int* list, const long long* off, View M, const double4* trig,
                                                      int tiles_max, double* partial) {
  const int k = list[blockIdx.y];
  const long long s0 = off[k];
  const int n = (int)(off[k + 1] - s0);
  const int i0 = blockIdx.x * kPairRows;
  double* dst = partial + (long long)blockIdx.y * tiles_max + blockIdx.x;
  if (i0 >= n - 1) {
    if (threadIdx.x == 0) *dst = 0.0;
    return;
  }
  const int i1 = min(i0 + kPairRows, n - 1);
  __shared__ double4 rows[kPairRows];
  if (threadIdx.x < i1 - i0) rows[threadIdx.x] = trig[s0 + i0 + threadIdx.x];
  __syncthreads();
  double psum = 0.0;
  for (int j = i0 + 1 + threadIdx.x; j < n; j += 256) {
    const double4 b = trig[s0 + j];
    const int top = min(i1, j);  // rows i < j
    for (int i = i0; i < top; ++i) {
      const double4 a = rows[i - i0];
      const double sdlon = b.x * a.y - b.y * a.x, cdlon = b.y * a.y + b.x * a.x;
      const double num1 = b.w * sdlon;
      const double num2 = a.w * b.z - a.z * b.w * cdlon;
      const double den = a.z * b.z + a.w * b.w * cdlon;
      const double ss = num1 * num1 + num2 * num2;
      double inv = inv_small_separation(ss, den);
      if (inv < 0.0) {
        double sep = atan2(sqrt(ss), den);
        if (sep < 1e-6)
          sep = vincenty_rad(__dmul_rn(M.at(s0 + i, 0), kDeg2Rad), __dmul_rn(M.at(s0 + i, 1), kDeg2Rad),
                             __dmul_rn(M.at(s0 + j, 0), kDeg2Rad), __dmul_rn(M.at(s0 + j, 1), kDeg2Rad));
        inv = 1.0 / sep;
      }
      psum += inv;
    }
  }
  typedef cub::BlockReduce<double, 256> Reduce;
  __shared__ typename Reduce::TempStorage rtmp;
  const double t = Reduce(rtmp).Sum(psum);
  if (threadIdx.x == 0) *dst = t;
}

__global__ void k_pair_sum_finish(int nbig, const int* list, int tiles_max, const double* partial, double* big_psum) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbig) return;
  double acc = 0.0;
  for (int t = 0; t < tiles_max; ++t) acc += partial[(long long)b * tiles_max + t];
  big_psum[list[b]] = acc;
}

// ---- bootstrap index table --------------------------------------------------------------------------------------
// astropy's bootstrap under NumpyRNGContext(1) (mass_estimator.py:178-184) draws, for a cluster of n members, 100
// resamples of n - 1 indices with np.random.randint(0, n): 32-bit words of the seed-1 MT19937 stream, masked to the
// next power of two and rejected above n - 1.  The index sequence depends on n alone, so it is built once per distinct
// cluster size (one CTA per size, the masked-rejection filter as a counting pass, a block scan and a writing pass) and
// kept for the life of the handle: entry r * (n - 1) + j of the slice at boot_table_offset(n).
constexpr int kBootTabThreads = 256;
__host__ __device__ inline long long boot_table_offset(int n) { return (long long)kBootNum * (n - 1) * (n - 2) / 2; }

__global__ void k_sizes_present(int K, const long long* off, int max_n, int* present) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const long long n = off[k + 1] - off[k];
  if (n >= 2 && n <= max_n) present[n] = 1;
}

__global__ void __launch_bounds__(kBootTabThreads) k_boot_table(const int* todo, const unsigned* mt, long long mt_len,
                                                                unsigned short* table, int* status) {
  const int n = todo[blockIdx.x];
  const int tid = threadIdx.x;
  unsigned mask = (unsigned)(n - 1);
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
  const long long want = (long long)kBootNum * (n - 1);
  typedef cub::BlockScan<long long, kBootTabThreads> Scan;
  __shared__ typename Scan::TempStorage stmp;
  long long L = (long long)((double)want * ((double)mask + 1.0) / (double)n * 1.05) + 2048;
  if (L > mt_len) L = mt_len;
  long long lo_w, hi_w, base, total;
  for (;;) {
    const long long span = (L + kBootTabThreads - 1) / kBootTabThreads;
    lo_w = (long long)tid * span;
    hi_w = lo_w + span < L ? lo_w + span : L;
    long long cnt = 0;
    for (long long w = lo_w; w < hi_w; ++w) cnt += (mt[w] & mask) <= (unsigned)(n - 1) ? 1 : 0;
    __syncthreads();
    Scan(stmp).ExclusiveSum(cnt, base, total);
    if (total >= want || L >= mt_len) break;
    L = L + L / 4 + 4096;
    if (L > mt_len) L = mt_len;
  }
  if (total < want) {
    if (tid == 0) atomicAdd(status, 1);
    return;
  }
  unsigned short* dst = table + boot_table_offset(n);
  long long ord = base;
  for (long long w = lo_w; w < hi_w && ord < want; ++w) {
    const unsigned val = mt[w] & mask;
    if (val > (unsigned)(n - 1)) continue;
    dst[ord++] = (unsigned short)val;
  }
}

__global__ void __launch_bounds__(kVirialThreads) k_virial_mass(int K, const long long* off, View M, const double4* trig,
                                                                Cosmo cosmo, const unsigned* mt, long long mt_len,
                                                                double mass_unit, double* out, int* status,
                                                                const double* big_psum, const unsigned short* boot_table) {
  const int k = blockIdx.x;
  if (k >= K) return;
  const long long s0 = off[k];
  const int n = (int)(off[k + 1] - s0);
  const int tid = threadIdx.x;
  __shared__ double boot[kBootNum];
  __shared__ double s_zbar, s_vdisp, s_zavg_err;
  __shared__ long long s_taken;
  __shared__ int s_fail;
  double* o = out + (long long)k * 8;
  if (n < 2) {
    if (tid == 0) {
      for (int i = 0; i < 8; ++i) o[i] = NAN;
      o[7] = n;
    }
    return;
  }
  // The cluster's redshifts, redshift errors and v^2 live in shared memory (clusters of up to 1024 members; longer
  // ones read the member matrix): thread 0's numpy-ordered sums below are chains of dependent loads, which cost a
  // shared-memory access each instead of a cache round trip.
  __shared__ double sv2[kVirialCache], s_z[kVirialCache], s_e[kVirialCache];
  __shared__ double4 s_trig[kVirialTrig];
  const bool cached = n <= kVirialCache;
  const bool trig_cached = n <= kVirialTrig;
  if (cached)
    for (int i = tid; i < n; i += kVirialThreads) {
      const double z = M.at(s0 + i, 2);
      s_z[i] = z;
      s_e[i] = fmax(fabs(__dsub_rn(M.at(s0 + i, 3), z)), fabs(__dsub_rn(M.at(s0 + i, 4), z)));
    }
  if (trig_cached)
    for (int i = tid; i < n; i += kVirialThreads) s_trig[i] = trig[s0 + i];
  for (int i = tid; i < kBootNum; i += kVirialThreads) boot[i] = 0.0;
  __syncthreads();
  auto getz = [&](long long i) -> double { return cached ? s_z[i] : M.at(s0 + i, 2); };
  // The numpy-ordered sums: a chain of dependent adds each, which used to leave 127 threads at a barrier behind thread
  // 0.  Now warp 0 runs the pairwise sums with numpy's eight accumulators on eight lanes, every thread computes the
  // v^2 terms (one division each) in between, and the sequential builtin sum() of velocity_error runs on warp 1
  // beside the second pairwise sum.  Same operations in the same order: the results are bit-identical.
  const int lane0 = tid & 31, warp0 = tid >> 5;
  if (warp0 == 0) {
    const double zsum = warp_np_pairwise_sum(getz, n, lane0);
    if (lane0 == 0) s_zbar = __ddiv_rn(zsum, (double)n);  // np.mean
  }
  __syncthreads();
  if (cached)
    for (int i = tid; i < n; i += kVirialThreads) {
      const double v = redshift_to_velocity(s_z[i], s_zbar);
      sv2[i] = __dmul_rn(v, v);
    }
  __syncthreads();
  if (warp0 == 0) {
    auto getv2 = [&](long long i) -> double {
      if (cached) return sv2[i];
      const double v = redshift_to_velocity(getz(i), s_zbar);
      return __dmul_rn(v, v);
    };
    const double vsum = warp_np_pairwise_sum(getv2, n, lane0);
    if (lane0 == 0) {
      s_vdisp = __ddiv_rn(vsum, (double)(n - 1));  // np.sum(v**2)/(n-1)
      s_taken = 0;
      s_fail = 0;
    }
  } else if (tid == 32) {
    // velocity_error: z_average_err = sqrt(sum(zerr**2)) / n   (builtin sum: sequential)
    double acc = 0.0;
    for (int i = 0; i < n; ++i) {
      double e;
      if (cached) {
        e = s_e[i];
      } else {
        const double z = M.at(s0 + i, 2);
        e = fmax(fabs(__dsub_rn(M.at(s0 + i, 3), z)), fabs(__dsub_rn(M.at(s0 + i, 4), z)));
      }
      acc = __dadd_rn(acc, __dmul_rn(e, e));
    }
    s_zavg_err = __ddiv_rn(sqrt(acc), (double)n);
  }
  __syncthreads();
  const double zbar = s_zbar;
  // harmonic term: (1/n) * (sum(1/err^2))^-1   (order-free reduction; tolerance 1e-9)
  double part = 0.0, mag = 0.0;
  for (int i = tid; i < n; i += kVirialThreads) {
    const double z = getz(i);
    const double v = redshift_to_velocity(z, zbar);
    const double e = cached ? s_e[i] : fmax(fabs(M.at(s0 + i, 3) - z), fabs(M.at(s0 + i, 4) - z));
    const double zd = sqrt(s_zavg_err * s_zavg_err + e * e);
    const double a = s_zavg_err / zbar, b = zd / (z - zbar);
    const double err = fabs(v) * sqrt(a * a + b * b);
    part += 1.0 / (err * err);
    mag += pow(10.0, -0.4 * M.at(s0 + i, 5));
  }
  typedef cub::BlockReduce<double, kVirialThreads> Reduce;
  __shared__ typename Reduce::TempStorage rtmp;
  const double inv_err_sum = Reduce(rtmp).Sum(part);
  __syncthreads();
  const double mag_sum = Reduce(rtmp).Sum(mag);
  __syncthreads();
  // pair sum over i < j of 1 / (sep_deg * d_A * pi/180 * h)
  const double dA = cosmo_angular_diameter_distance(cosmo, zbar);
  double psum = 0.0;
  if (n <= kBigCluster) {
    // pairs (i, j), i < j, in row-major order of the upper triangle, strided over the block.  Vincenty's
    // formula with the angle-difference identities on the precomputed sines and cosines; pairs closer than
    // 1e-6 rad (where the identities lose digits) are recomputed from the angles themselves.
    int i = 0;
    long long j = 1 + tid;
    while (i < n - 1) {
      while (j >= n && i < n - 1) {
        j = j - n + (i + 2);
        ++i;
      }
      if (i >= n - 1) break;
      const double4 a = trig_cached ? s_trig[i] : trig[s0 + i], b = trig_cached ? s_trig[j] : trig[s0 + j];
      const double sdlon = b.x * a.y - b.y * a.x, cdlon = b.y * a.y + b.x * a.x;
      const double num1 = b.w * sdlon;
      const double num2 = a.w * b.z - a.z * b.w * cdlon;
      const double den = a.z * b.z + a.w * b.w * cdlon;
      const double ss = num1 * num1 + num2 * num2;
      double inv = inv_small_separation(ss, den);
      if (inv < 0.0) {
        double sep = atan2(sqrt(ss), den);
        if (sep < 1e-6)
          sep = vincenty_rad(__dmul_rn(M.at(s0 + i, 0), kDeg2Rad), __dmul_rn(M.at(s0 + i, 1), kDeg2Rad),
                             __dmul_rn(M.at(s0 + j, 0), kDeg2Rad), __dmul_rn(M.at(s0 + j, 1), kDeg2Rad));
        inv = 1.0 / sep;
      }
      psum += inv;
      j += kVirialThreads;
    }
  }
  // sum over pairs of 1 / sep [rad]  ->  sum of 1 / (sep_deg * d_A * pi/180 * h)
  double total_sep = Reduce(rtmp).Sum(psum) / (kRad2Deg * dA * kDeg2Rad * cosmo.h);
  if (n > kBigCluster) total_sep = big_psum[k] / (kRad2Deg * dA * kDeg2Rad * cosmo.h);  // summed by k_pair_sum_big
  __syncthreads();
  // bootstrap: 100 resamples of n-1 indices from the masked MT19937 stream.  Every thread owns a contiguous span of
  // the raw stream: one counting pass, one block scan for the ordinal of its first accepted word, then a second
  // pass that accumulates v^2 per resample and touches shared memory only when the resample changes.
  const long long want = (long long)kBootNum * (n - 1);
  typedef cub::BlockScan<long long, kVirialThreads> Scan;
  __shared__ typename Scan::TempStorage stmp;
  if (cached && boot_table) {
    // the resample indices of this cluster size come from the table: a warp per resample, fixed summation order
    const unsigned short* idx = boot_table + boot_table_offset(n);
    const int lane = tid & 31, warp = tid >> 5;
    for (int r = warp; r < kBootNum; r += kVirialThreads / 32) {
      const unsigned short* row = idx + (long long)r * (n - 1);
      double acc = 0.0;
      for (int j = lane; j < n - 1; j += 32) acc += sv2[row[j]];
      for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) boot[r] = acc;
    }
    if (tid == 0) s_taken = want;
  } else {
  unsigned mask = (unsigned)(n - 1);
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
  long long L = (long long)((double)want * ((double)mask + 1.0) / (double)n * 1.05) + 2048;
  if (L > mt_len) L = mt_len;
  long long lo_w, hi_w, base, total;
  for (;;) {
    const long long span = (L + kVirialThreads - 1) / kVirialThreads;
    lo_w = (long long)tid * span;
    hi_w = lo_w + span < L ? lo_w + span : L;
    long long cnt = 0;
    for (long long w = lo_w; w < hi_w; ++w) cnt += (mt[w] & mask) <= (unsigned)(n - 1) ? 1 : 0;
    __syncthreads();
    Scan(stmp).ExclusiveSum(cnt, base, total);
    if (total >= want || L >= mt_len) break;
    L = L + L / 4 + 4096;
    if (L > mt_len) L = mt_len;
  }
  if (tid == 0) s_taken = total;
  {
    long long ord = base;
    int r = (int)(ord / (n - 1));
    int rem = (int)(ord - (long long)r * (n - 1));
    double acc = 0.0;
    for (long long w = lo_w; w < hi_w && ord < want; ++w) {
      const unsigned val = mt[w] & mask;
      if (val > (unsigned)(n - 1)) continue;
      double v2;
      if (cached) {
        v2 = sv2[val];
      } else {
        const double v = redshift_to_velocity(M.at(s0 + val, 2), zbar);
        v2 = v * v;
      }
      acc += v2;
      ++ord;
      if (++rem == n - 1) {
        atomicAdd(&boot[r], acc);
        acc = 0.0;
        rem = 0;
        ++r;
      }
    }
    if (acc != 0.0 && r < kBootNum) atomicAdd(&boot[r], acc);
  }
  }
  __syncthreads();
  // statistic: sum(x^2) / (len(x) - 1) with len(x) = n - 1; np.std over the 100 values.  The divisions and the
  // deviations are a thread each; the two sums stay sequential on thread 0, in the order they always had.
  __shared__ double s_bmean;
  for (int r = tid; r < kBootNum; r += kVirialThreads) boot[r] = boot[r] / (double)(n - 2);
  __syncthreads();
  if (tid == 0) {
    double mean = 0.0;
    for (int r = 0; r < kBootNum; ++r) mean += boot[r];
    s_bmean = mean / kBootNum;
  }
  __syncthreads();
  for (int r = tid; r < kBootNum; r += kVirialThreads) boot[r] = boot[r] - s_bmean;
  __syncthreads();
  if (tid == 0) {
    if (s_taken < want) s_fail = 1;
    double var = 0.0;
    for (int r = 0; r < kBootNum; ++r) var = fma(boot[r], boot[r], var);
    const double bootstrap_disp = sqrt(var / kBootNum);
    const double harmonic = (1.0 / n) * (1.0 / inv_err_sum);
    const double combined = sqrt(harmonic * harmonic + bootstrap_disp * bootstrap_disp);
    const double radius = (double)n * (double)(n - 1) / total_sep;
    const double mass = 1.5 * kPi * (s_vdisp * radius) * mass_unit;
    o[0] = mass;
    o[1] = mass * (combined / s_vdisp);
    o[2] = s_vdisp;
    o[3] = combined;
    o[4] = radius;
    o[5] = -2.5 * log10(mag_sum);
    o[6] = zbar;
    o[7] = (double)n;
    if (s_fail) atomicAdd(status, 1);
  }
}

// ---- projected mass estimator (analysis/mass_estimator.py:115-145), one CTA per cluster -----------------------
__global__ void __launch_bounds__(kVirialThreads) k_projected_mass(int K, const long long* off, View M, Cosmo cosmo,
                                                                   double mass_unit, double* out) {
  const int k = blockIdx.x;
  if (k >= K) return;
  const long long s0 = off[k];
  const int n = (int)(off[k + 1] - s0);
  const int tid = threadIdx.x;
  typedef cub::BlockReduce<double, kVirialThreads> Reduce;
  __shared__ typename Reduce::TempStorage rtmp;
  __shared__ double s_zbar, s_ra, s_dec, s_zavg_err;
  double* o = out + (long long)k * 2;
  if (n < 2) {
    if (tid == 0) { o[0] = NAN; o[1] = NAN; }
    return;
  }
  if (tid == 0) {
    auto getz = [&](long long i) -> double { return M.at(s0 + i, 2); };
    auto getra = [&](long long i) -> double { return M.at(s0 + i, 0); };
    auto getdec = [&](long long i) -> double { return M.at(s0 + i, 1); };
    s_zbar = __ddiv_rn(np_pairwise_sum(getz, 0, n), (double)n);
    s_ra = __ddiv_rn(np_pairwise_sum(getra, 0, n), (double)n);    // centroid = plain means of ra and dec
    s_dec = __ddiv_rn(np_pairwise_sum(getdec, 0, n), (double)n);
    double acc = 0.0;
    for (int i = 0; i < n; ++i) {
      const double z = M.at(s0 + i, 2);
      const double e = fmax(fabs(__dsub_rn(M.at(s0 + i, 3), z)), fabs(__dsub_rn(M.at(s0 + i, 4), z)));
      acc = __dadd_rn(acc, __dmul_rn(e, e));
    }
    s_zavg_err = __ddiv_rn(sqrt(acc), (double)n);
  }
  __syncthreads();
  const double zbar = s_zbar;
  const double dA = cosmo_angular_diameter_distance(cosmo, zbar);
  const double clon = __dmul_rn(s_ra, kDeg2Rad), clat = __dmul_rn(s_dec, kDeg2Rad);
  double sp = 0.0, se = 0.0;
  for (int i = tid; i < n; i += kVirialThreads) {
    const double z = M.at(s0 + i, 2);
    const double v = redshift_to_velocity(z, zbar);
    const double sep = vincenty_rad(clon, clat, __dmul_rn(M.at(s0 + i, 0), kDeg2Rad), __dmul_rn(M.at(s0 + i, 1), kDeg2Rad));
    const double actual = sep * kRad2Deg * dA * kDeg2Rad * cosmo.h;
    sp += actual * v * v;
    const double e = fmax(fabs(M.at(s0 + i, 3) - z), fabs(M.at(s0 + i, 4) - z));
    const double zd = sqrt(s_zavg_err * s_zavg_err + e * e);
    const double a = s_zavg_err / zbar, b = zd / (z - zbar);
    const double err = fabs(v) * sqrt(a * a + b * b);
    const double t = 2.0 * v * err;
    se += t * t;
  }
  const double sumproduct = Reduce(rtmp).Sum(sp);
  __syncthreads();
  const double errsum = Reduce(rtmp).Sum(se);
  if (tid == 0) {
    o[0] = (32.0 / kPi) * (1.0 / ((double)n - 1.5)) * sumproduct * mass_unit;
    o[1] = sqrt(errsum);
  }
}

void CleanState::projected(fofb_handle* h, const fofb_params& p, int K, const long long* off_host, const View& M,
                           double* out_host) {
  cudaStream_t st = h->stream;
  if (K == 0) return;
  FOFB_REQUIRE(off_host[K] == M.rows, "projected: offsets do not cover the member matrix");
  FOFB_REQUIRE(M.cols >= 5, "projected: members need columns ra, dec, z, z_lower, z_upper");
  const Cosmo cosmo = device_cosmo(h, p);
  long long* off = d_off.alloc((size_t)K + 1);
  FOFB_CUDA(cudaMemcpyAsync(off, off_host, sizeof(long long) * ((size_t)K + 1), cudaMemcpyHostToDevice, st));
  double* out = d_props.alloc((size_t)K * 8);
  const double pc_m = 149597870700.0 * 648000.0 / kPi;
  const double msun_kg = 1.3271244e20 / 6.6743e-11;
  const double mass_unit = 1.0e6 * (1.0e6 * pc_m) / 6.6743e-11 / msun_kg;
  FOFB_PROFILED(h, "k_projected_mass", k_projected_mass<<<K, kVirialThreads, 0, st>>>(K, off, M, cosmo, mass_unit, out));
  FOFB_CUDA(cudaMemcpyAsync(out_host, out, sizeof(double) * (size_t)K * 2, cudaMemcpyDeviceToHost, st));
  FOFB_CUDA(cudaStreamSynchronize(st));
}

void CleanState::virial_core(fofb_handle* h, const fofb_params& p, int K, const long long* off, const View& M, int nmax,
                             double* out) {
  cudaStream_t st = h->stream;
  if (K == 0) return;
  FOFB_REQUIRE(M.cols >= 6, "virial: members need columns ra, dec, z, z_lower, z_upper, RMag");
  const Cosmo cosmo = device_cosmo(h, p);
  // masked rejection keeps >= half of the words on average; 4x + slack bounds the need generously
  const size_t need = (size_t)kBootNum * (size_t)std::max(nmax, 2) * 4 + 65536;
  if (need > mt_len) {
    std::vector<unsigned> words;
    mt_stream(words, need);
    unsigned* d = d_mt.alloc(need);
    FOFB_CUDA(cudaMemcpyAsync(d, words.data(), sizeof(unsigned) * need, cudaMemcpyHostToDevice, st));
    FOFB_CUDA(cudaStreamSynchronize(st));
    mt_len = need;
  }
  int* status = d_status.alloc(2);
  FOFB_CUDA(cudaMemsetAsync(status, 0, sizeof(int), st));
  // (km/s)^2 Mpc / G in solar masses: CODATA-2018 G, IAU-2015 GM_sun and au
  const double pc_m = 149597870700.0 * 648000.0 / kPi;
  const double msun_kg = 1.3271244e20 / 6.6743e-11;
  const double mass_unit = 1.0e6 * (1.0e6 * pc_m) / 6.6743e-11 / msun_kg;
  double4* trig = d_trig.alloc((size_t)std::max<long long>(M.rows, 1));
  if (M.rows > 0) {
    k_member_trig<<<grid_for(M.rows, 256), 256, 0, st>>>(M.rows, M, trig);
    FOFB_LAUNCH_CHECK(h);
  }
  double* big_psum = d_big_psum.alloc((size_t)K);
  if (nmax > kBigCluster) {
    int* cnt = d_big_count.alloc(2);
    int* list = d_big_slot.alloc((size_t)K);
    FOFB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int), st));
    k_big_clusters<<<grid_for(K, 256), 256, 0, st>>>(K, off, cnt, list);
    FOFB_LAUNCH_CHECK(h);
    int nbig = 0;
    FOFB_CUDA(cudaMemcpyAsync(&nbig, cnt, sizeof(int), cudaMemcpyDeviceToHost, st));
    FOFB_CUDA(cudaStreamSynchronize(st));
    if (nbig > 0) {
      FOFB_REQUIRE(nbig <= 65535, "virial: more than 65535 clusters above 4096 members");
      const int tiles = (nmax + kPairRows - 1) / kPairRows;
      double* partial = d_big_dev.alloc((size_t)nbig * tiles);
      FOFB_PROFILED(h, "k_pair_sum_big", k_pair_sum_big<<<dim3(tiles, nbig), 256, 0, st>>>(list, off, M, trig, tiles, partial));
      k_pair_sum_finish<<<grid_for(nbig, 64), 64, 0, st>>>(nbig, list, tiles, partial, big_psum);
      FOFB_LAUNCH_CHECK(h);
    }
  }
  // bootstrap index table: add the slices of the cluster sizes not seen before (usually none after the first call)
  {
    if (boot_built.empty()) boot_built.assign(kVirialCache + 1, 0);
    int* present = d_present.alloc(kVirialCache + 2);
    FOFB_CUDA(cudaMemsetAsync(present, 0, sizeof(int) * (kVirialCache + 2), st));
    k_sizes_present<<<grid_for(K, 256), 256, 0, st>>>(K, off, kVirialCache, present);
    FOFB_LAUNCH_CHECK(h);
    std::vector<int> hp(kVirialCache + 2);
    FOFB_CUDA(cudaMemcpyAsync(hp.data(), present, sizeof(int) * (kVirialCache + 2), cudaMemcpyDeviceToHost, st));
    FOFB_CUDA(cudaStreamSynchronize(st));
    std::vector<int> todo;
    for (int n = 2; n <= kVirialCache; ++n)
      if (hp[n] && !boot_built[n]) todo.push_back(n);
    if (!todo.empty()) {
      if (!d_boot_table.p) d_boot_table.alloc((size_t)boot_table_offset(kVirialCache + 1) + 64);
      int* dtodo = d_present.p;  // reuse: the presence flags have been read
      FOFB_CUDA(cudaMemcpyAsync(dtodo, todo.data(), sizeof(int) * todo.size(), cudaMemcpyHostToDevice, st));
      FOFB_PROFILED(h, "k_boot_table", k_boot_table<<<(unsigned)todo.size(), kBootTabThreads, 0, st>>>(dtodo, d_mt.p, (long long)mt_len, d_boot_table.p, status));
      FOFB_CUDA(cudaStreamSynchronize(st));  // todo[] lives on the host stack of this call
      for (int n : todo) boot_built[n] = 1;
    }
  }
  FOFB_PROFILED(h, "k_virial_mass", k_virial_mass<<<K, kVirialThreads, 0, st>>>(K, off, M, trig, cosmo, d_mt.p, (long long)mt_len, mass_unit, out, status, big_psum, d_boot_table.p));
  int bad = 0;
  FOFB_CUDA(cudaMemcpyAsync(&bad, status, sizeof(int), cudaMemcpyDeviceToHost, st));
  FOFB_CUDA(cudaStreamSynchronize(st));
  if (bad) throw Error(FOFB_ERR_INTERNAL, "bootstrap random stream exhausted");
}

void CleanState::virial(fofb_handle* h, const fofb_params& p, int K, const long long* off_host, const View& M,
                        double* out_host) {
  cudaStream_t st = h->stream;
  if (K == 0) return;
  const long long total = off_host[K];
  FOFB_REQUIRE(total == M.rows, "virial: offsets do not cover the member matrix");
  int nmax = 0;
  for (int k = 0; k < K; ++k) nmax = std::max<long long>(nmax, off_host[k + 1] - off_host[k]);
  long long* off = d_off.alloc((size_t)K + 1);
  FOFB_CUDA(cudaMemcpyAsync(off, off_host, sizeof(long long) * ((size_t)K + 1), cudaMemcpyHostToDevice, st));
  double* out = d_props.alloc((size_t)K * 8);
  virial_core(h, p, K, off, M, nmax, out);
  FOFB_CUDA(cudaMemcpyAsync(out_host, out, sizeof(double) * (size_t)K * 8, cudaMemcpyDeviceToHost, st));
  FOFB_CUDA(cudaStreamSynchronize(st));
}

}  // namespace fofb
