// Stages 2-4 of FoF() (fof.py:258-392): ordered overlap contests between candidate clusters,
// the BCG recentring check, own-member N(0.5) recount and the 300-point overdensity
// (kernels K10, K11 of DESIGN.md).
#include <cub/cub.cuh>

#include <cstring>
#include <ctime>

#include "fof_stage.cuh"
#include "mt_jump_tables.h"
#include "tile.cuh"

namespace fofb {

static unsigned char* cub_tmp(fofb_handle* h, size_t bytes) { return h->cub_tmp.alloc(bytes + 256); }


__global__ void k_scan_total_ov(const long long* len, long long* off, int n) { off[n] = off[n - 1] + len[n - 1]; }

static long long scan_offsets(fofb_handle* h, const long long* len, long long* off, int n) {
  if (n == 0) return 0;
  size_t bytes = 0;
  FOFB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, len, off, n, h->stream));
  FOFB_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp(h, bytes), bytes, len, off, n, h->stream));
  k_scan_total_ov<<<1, 1, 0, h->stream>>>(len, off, n);  // off[n] = total, on the device
  count_launch(h, 3);
  return d2h_scalar(h, off + n);
}

struct OverlapScratch {
  DBuf<double> kz, kz_sorted, ids, ids_sorted, idk, idk_sorted;
  DBuf<int> kidx, kidx_sorted, first_by_id, idv, idv_sorted, seg_a, seg_b;
  DBuf<int> wlo, whi, ecount, ev_a, ev_c, ev_loser, state, flag;
  DBuf<long long> eoff, elen, loff, soff;
  DBuf<unsigned long long> ekey, ekey_sorted, lkey, lkey_sorted;
  DBuf<double4> bbox;
  DBuf<PackedPoint> od_pay_a, od_pay_b;
  DBuf<unsigned char> keep;
  DBuf<int> sel, selcount, attr_iota;
  DBuf<double> attr;
};

__global__ void k_iota(int n, int* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = i;
}

__global__ void k_cluster_z(int K, const int* centre, const double* c_z, const double* c_id, double* kz, int* kidx,
                            double* idk, int* idv) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  kz[k] = c_z[centre[k]];
  kidx[k] = k;
  idk[k] = c_id[centre[k]] + 0.0;
  idv[k] = k;
}

// candidates[gal_id_space == id][0]: the first cluster carrying each gal_id (fof.py:288)
__global__ void k_first_by_id(int K, const double* idk_sorted, const int* idv_sorted, int* first_by_id) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= K) return;
  int s = t;
  const double v = idk_sorted[t];
  // runs are short; walk back to the start of the run of equal ids (stable sort => ascending k)
  while (s > 0 && idk_sorted[s - 1] == v) --s;
  first_by_id[idv_sorted[t]] = (v == v) ? idv_sorted[s] : -1;  // NaN id matches nothing
}

// Per cluster a (one warp): velocity window over the candidate-cluster list, its bounding box, and the
// overlapping centres inside theta_vir(a.z) in GriSPy's return order (fof.py:270-283).
template <int MODE>
__global__ void k_overlap_events(int K, const int* centre, const double* kz_sorted, const int* kidx_sorted,
                                 const double* c_ra, const double* c_dec, const double* c_cos, const double* c_z,
                                 const double* c_thvir, double vmax, int grid_cells, int* ecount, const long long* eoff,
                                 unsigned long long* ekey) {
  const int a = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (a >= K) return;
  const int ca = centre[a];
  const double ra = c_ra[ca], dec = c_dec[ca], z = c_z[ca], r = c_thvir[ca];
  int lo, hi;
  velocity_window(kz_sorted, K, z, vmax, &lo, &hi);
  double4 bb = make_double4(INFINITY, -INFINITY, INFINITY, -INFINITY);
  for (int t = lo + lane; t < hi; t += 32) {
    const int cc = centre[kidx_sorted[t]];
    bb = bbox_merge(bb, make_double4(c_ra[cc], c_ra[cc], c_dec[cc], c_dec[cc]));
  }
  for (int o = 16; o; o >>= 1) {
    double4 other;
    other.x = __shfl_xor_sync(0xffffffffu, bb.x, o); other.y = __shfl_xor_sync(0xffffffffu, bb.y, o);
    other.z = __shfl_xor_sync(0xffffffffu, bb.z, o); other.w = __shfl_xor_sync(0xffffffffu, bb.w, o);
    bb = bbox_merge(bb, other);
  }
  const BBox b{bb.x, bb.y, bb.z, bb.w};
  const Box box = grid_box(b, grid_cells, ra, dec, r);
  BubbleTest bt;
  bt.init(ra, dec, c_cos[ca], r);
  GridAxis gx, gy;
  gx.init(b.ra_min, b.ra_max, grid_cells);
  gy.init(b.dec_min, b.dec_max, grid_cells);
  int cnt = 0;
  for (int base = lo; base < hi; base += 32) {
    const int t = base + lane;
    bool hit = false;
    int kc = 0;
    double pra = 0, pdec = 0;
    if (t < hi) {
      kc = kidx_sorted[t];
      const int cc = centre[kc];
      pra = c_ra[cc]; pdec = c_dec[cc];
      hit = box.contains(pra, pdec) && bt.inside(pra, pdec, c_cos[cc]);
    }
    const unsigned mask = __ballot_sync(0xffffffffu, hit);
    if (MODE == 1 && hit) {
      const unsigned long long cellkey = (unsigned)(gy.digit_clamped(pdec) * grid_cells + gx.digit_clamped(pra));
      ekey[eoff[a] + cnt + __popc(mask & ((1u << lane) - 1u))] = (cellkey << 32) | (unsigned)kc;
    }
    cnt += __popc(mask);
  }
  if (MODE == 0 && lane == 0) ecount[a] = cnt;
}

__global__ void k_to_ll(int n, const int* in, long long* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

// member IDs (column 6) per candidate cluster, to be segment-sorted for the subset test
__global__ void k_member_ids(long long total, const int* rows, View G, double* ids) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) ids[i] = G.at(rows[i], 6);
}

// the same with the cluster index beside every id (one warp per cluster); ext_ids: ids given directly
__global__ void k_member_ids_seg(int K, const long long* off, const int* rows, View G, const double* ext_ids, double* ids, int* seg) {
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (k >= K) return;
  for (long long t = off[k] + lane; t < off[k + 1]; t += 32) {
    ids[t] = ext_ids ? ext_ids[t] : G.at(rows[t], 6);
    seg[t] = k;
  }
}

// Per event (one warp): partner lookup, the "c is not a subset of center" test (helper.py:138-146 via
// fof.py:297-300) and the static loser of the contest (cluster.py:65-82).
__global__ void k_event_static(int K, long long E, const long long* eoff, const unsigned long long* ekey_sorted,
                               const int* first_by_id, const int* centre, const double* c_id, const double* c_N,
                               const double* c_mag, const long long* moff, const double* ids_sorted, int* ev_a,
                               int* ev_c, int* ev_loser, int* state) {
  const long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (e >= E) return;
  // owner a of event e: last a with eoff[a] <= e
  int lo = 0, hi = K;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (eoff[mid] <= e) lo = mid; else hi = mid;
  }
  const int a = lo;
  const int kc = (int)(unsigned)(ekey_sorted[e] & 0xffffffffull);
  const int c = first_by_id[kc];
  const int ca = centre[a];
  bool eligible = false;
  int loser = -1;
  if (c >= 0 && !(c_id[ca] == c_id[centre[c]])) {
    // any member id of c missing from a's (sorted) ids?
    const long long a0 = moff[a], a1 = moff[a + 1], s0 = moff[c], s1 = moff[c + 1];
    bool missing = false;
    for (long long t = s0 + lane; t < s1 && !missing; t += 32) {
      const double v = ids_sorted[t];
      long long x = a0, y = a1;
      while (x < y) {
        const long long mid = (x + y) >> 1;
        if (ids_sorted[mid] < v) x = mid + 1; else y = mid;
      }
      if (!(x < a1 && ids_sorted[x] == v)) missing = true;
    }
    eligible = __any_sync(0xffffffffu, missing);
    const int cc = centre[c];
    const double Nc = c_N[cc], Na = c_N[ca];
    if (Nc > Na) loser = a;
    else if (Nc < Na) loser = c;
    else if (fabs(c_mag[cc]) > fabs(c_mag[ca])) loser = a;
    else loser = c;
  }
  if (lane == 0) {
    ev_a[e] = a;
    ev_c[e] = c;
    ev_loser[e] = eligible ? loser : -1;
    state[e] = eligible ? 0 : 2;  // 0 undecided, 1 fired, 2 not fired
  }
}

__global__ void k_loser_keys(long long E, const int* ev_loser, int K, unsigned long long* key) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int l = ev_loser[e];
  key[e] = ((unsigned long long)(unsigned)(l < 0 ? K : l) << 32) | (unsigned)e;
}

__global__ void k_loser_starts(long long E, const unsigned long long* key, int K, long long* start) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k > E) return;
  const int cur = k < E ? (int)(key[k] >> 32) : K + 1;
  const int prev = k > 0 ? (int)(key[k - 1] >> 32) : -1;
  for (int c = prev + 1; c <= cur && c <= K + 1; ++c) start[c] = k;
}

// One relaxation round of the ordered contests: an event fires iff it is eligible and neither of its
// clusters lost an earlier fired event (fof.py:294-310 with the in-place mutation of Q6).
__global__ void k_event_round(long long E, const int* ev_a, const int* ev_c, const long long* lstart,
                              const unsigned long long* lkey_sorted, volatile int* state, int* undecided) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E || state[e] != 0) return;
  bool dead = false, pending = false;
  const int xs[2] = {ev_a[e], ev_c[e]};
  for (int s = 0; s < 2 && !dead; ++s) {
    const int x = xs[s];
    for (long long t = lstart[x]; t < lstart[x + 1]; ++t) {
      const long long e2 = (long long)(unsigned)(lkey_sorted[t] & 0xffffffffull);
      if (e2 >= e) break;
      const int st = state[e2];
      if (st == 1) { dead = true; break; }
      if (st == 0) pending = true;
    }
  }
  if (dead) state[e] = 2;
  else if (!pending) state[e] = 1;
  else atomicAdd(undecided, 1);
}

__global__ void k_cluster_keep(int K, const long long* lstart, const unsigned long long* lkey_sorted, const int* state,
                               int* keep) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= K) return;
  int alive = 1;
  for (long long t = lstart[x]; t < lstart[x + 1]; ++t)
    if (state[(long long)(unsigned)(lkey_sorted[t] & 0xffffffffull)] == 1) { alive = 0; break; }
  keep[x] = alive;
}

__global__ void k_sublist_sizes(int K2, const int* sel, const long long* off, long long* len) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < K2) len[k] = off[sel[k] + 1] - off[sel[k]];
}

__global__ void k_sublist_gather(int K2, const int* sel, const int* centre, const long long* off, const int* rows,
                                 const long long* off2, int* centre2, int* rows2) {
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (k >= K2) return;
  const int s = sel[k];
  const long long src = off[s], dst = off2[k], len = off[s + 1] - off[s];
  for (long long t = lane; t < len; t += 32) rows2[dst + t] = rows[src + t];
  if (lane == 0) centre2[k] = centre[s];
}

static void sublist(fofb_handle* h, const ClusterList& in, const int* keep_flags, ClusterList& out, OverlapScratch& sc) {
  cudaStream_t st = h->stream;
  out.K = 0;
  out.total = 0;
  if (in.K == 0) return;
  const int B = 256;
  int* iota = sc.flag.alloc(in.K);
  k_iota<<<grid_for(in.K, B), B, 0, st>>>(in.K, iota);
  FOFB_LAUNCH_CHECK(h);
  int* sel = sc.sel.alloc(in.K);
  int* dcount = sc.selcount.alloc(1);
  size_t bytes = 0;
  FOFB_CUDA(cub::DeviceSelect::Flagged(nullptr, bytes, iota, keep_flags, sel, dcount, in.K, st));
  FOFB_CUDA(cub::DeviceSelect::Flagged(cub_tmp(h, bytes), bytes, iota, keep_flags, sel, dcount, in.K, st));
  count_launch(h, 2);
  const int K2 = d2h_scalar(h, dcount);
  out.K = K2;
  if (K2 == 0) return;
  long long* len = sc.elen.alloc((size_t)K2 + 1);
  long long* off2 = out.off.alloc((size_t)K2 + 1);
  k_sublist_sizes<<<grid_for(K2, B), B, 0, st>>>(K2, sel, in.off.p, len);
  FOFB_LAUNCH_CHECK(h);
  out.total = scan_offsets(h, len, off2, K2);
  k_sublist_gather<<<grid_for((int64_t)K2 * 32, 128), 128, 0, st>>>(K2, sel, in.centre.p, in.off.p, in.rows.p, off2,
                                                                    out.centre.alloc(K2),
                                                                    out.rows.alloc((size_t)std::max<long long>(out.total, 1)));
  FOFB_LAUNCH_CHECK(h);
}

// ---- stage 3: would the BCG replacement of fof.py:346-367 ever fire? -----------------------------------------------
__global__ void k_bcg_check(int K, const int* centre, const long long* off, const int* rows, View G,
                            const double* ra_row, const double* dec_row, const double* cos_row, const double* c_ra,
                            const double* c_dec, const double* c_cos, const double* c_thvir, const double* c_mag,
                            const double* c_id, double frac, int grid_cells, int* fired) {
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (k >= K) return;
  const int c = centre[k];
  const long long s0 = off[k], s1 = off[k + 1];
  double4 bb = make_double4(INFINITY, -INFINITY, INFINITY, -INFINITY);
  for (long long t = s0 + lane; t < s1; t += 32) {
    const int row = rows[t];
    bb = bbox_merge(bb, make_double4(ra_row[row], ra_row[row], dec_row[row], dec_row[row]));
  }
  for (int o = 16; o; o >>= 1) {
    double4 other;
    other.x = __shfl_xor_sync(0xffffffffu, bb.x, o); other.y = __shfl_xor_sync(0xffffffffu, bb.y, o);
    other.z = __shfl_xor_sync(0xffffffffu, bb.z, o); other.w = __shfl_xor_sync(0xffffffffu, bb.w, o);
    bb = bbox_merge(bb, other);
  }
  const BBox b{bb.x, bb.y, bb.z, bb.w};
  const double r = frac * c_thvir[c];
  const Box box = grid_box(b, grid_cells, c_ra[c], c_dec[c], r);
  BubbleTest bt;
  bt.init(c_ra[c], c_dec[c], c_cos[c], r);
  const double mag = fabs(c_mag[c]);
  bool fire = false;
  // conservative: ANY in-radius member with |col 3| > |bcg_absMag| and col 4 != gal_id could become the
  // "brightest" pick, whatever the mask against existing centres removes
  for (long long t = s0 + lane; t < s1; t += 32) {
    const int row = rows[t];
    if (box.contains(ra_row[row], dec_row[row]) && bt.inside(ra_row[row], dec_row[row], cos_row[row]) &&
        fabs(G.at(row, 3)) > mag && G.at(row, 4) != c_id[c])
      fire = true;
  }
  if (__any_sync(0xffffffffu, fire) && lane == 0) atomicAdd(fired, 1);
}

void fof_bcg_check(fofb_handle* h, FofState& S);

static OverlapScratch& scratch(FofState& S) {
  if (!S.overlap_scratch) {
    S.overlap_scratch = new OverlapScratch();
    S.overlap_scratch_free = [](void* p) { delete static_cast<OverlapScratch*>(p); };
  }
  return *static_cast<OverlapScratch*>(S.overlap_scratch);
}

struct CentreAttrs {  // per-centre scalars stage 2 reads, indexed by the cluster list's centre index
  const double *ra, *dec, *cos, *z, *thvir, *N, *mag, *id;
};

static CentreAttrs attrs_of(const FofState& S) {
  return CentreAttrs{S.c_ra.p, S.c_dec.p, S.c_cos.p, S.c_z.p, S.c_thvir.p, S.c_N.p, S.c_mag.p, S.c_id.p};
}

// centre rows (8 columns) -> the scalars of CentreAttrs (fof.py:277-279; cluster.py:29-38)
__global__ void k_attrs_from_rows(int K, const double* rows, Cosmo cosmo, double virial_radius, double* ra, double* dec,
                                  double* cs, double* z, double* thvir, double* N, double* mag, double* id, int* iota) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const double* r = rows + 8LL * k;
  ra[k] = r[0];
  dec[k] = r[1];
  cs[k] = cos(__dmul_rn(r[1], kDeg2Rad));
  z[k] = r[2];
  thvir[k] = cosmo_linear_to_angular_dist(cosmo, virial_radius, r[2]);
  mag[k] = r[5];
  id[k] = r[6];
  N[k] = r[7];
  iota[k] = k;
}

// Stage 2, first half, on an arbitrary cluster list (centre index per cluster, CSR offsets, member ids): the ordered
// list of contests.  Fills sc.ev_a / ev_c / ev_loser (cluster a, its partner c, the static loser or -1 when the contest
// cannot fire), ordered by a and, within a, in GriSPy's return order; returns their number.
// `ext_ids` == nullptr: ids are column 6 of S.G at `rows`.
static long long overlap_events(fofb_handle* h, FofState& S, const CentreAttrs& A, int K, const int* centre, const long long* off,
                                long long total_members, const int* rows, const double* ext_ids) {
  cudaStream_t st = h->stream;
  OverlapScratch& sc = scratch(S);
  const int B = 256;
  if (K == 0) return 0;
  const fofb_params& p = S.params;
  size_t bytes = 0;
  // clusters sorted by centre redshift -> velocity windows over the candidate list
  k_cluster_z<<<grid_for(K, B), B, 0, st>>>(K, centre, A.z, A.id, sc.kz.alloc(K), sc.kidx.alloc(K),
                                            sc.idk.alloc(K), sc.idv.alloc(K));
  FOFB_LAUNCH_CHECK(h);
  sc.kz_sorted.alloc(K); sc.kidx_sorted.alloc(K); sc.idk_sorted.alloc(K); sc.idv_sorted.alloc(K);
  FOFB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, sc.kz.p, sc.kz_sorted.p, sc.kidx.p, sc.kidx_sorted.p, K, 0, 64, st));
  FOFB_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp(h, bytes), bytes, sc.kz.p, sc.kz_sorted.p, sc.kidx.p, sc.kidx_sorted.p, K, 0, 64, st));
  FOFB_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp(h, bytes), bytes, sc.idk.p, sc.idk_sorted.p, sc.idv.p, sc.idv_sorted.p, K, 0, 64, st));
  count_launch(h, 16);
  k_first_by_id<<<grid_for(K, B), B, 0, st>>>(K, sc.idk_sorted.p, sc.idv_sorted.p, sc.first_by_id.alloc(K));
  FOFB_LAUNCH_CHECK(h);
  // events
  const int wblocks = grid_for((int64_t)K * 32, 128);
  int* ecount = sc.ecount.alloc(K);
  k_overlap_events<0><<<wblocks, 128, 0, st>>>(K, centre, sc.kz_sorted.p, sc.kidx_sorted.p, A.ra, A.dec,
                                               A.cos, A.z, A.thvir, p.max_velocity, p.grid_cells, ecount, nullptr, nullptr);
  FOFB_LAUNCH_CHECK(h);
  long long* elen = sc.elen.alloc((size_t)K + 1);
  long long* eoff = sc.eoff.alloc((size_t)K + 1);
  k_to_ll<<<grid_for(K, B), B, 0, st>>>(K, ecount, elen);
  FOFB_LAUNCH_CHECK(h);
  const long long E = scan_offsets(h, elen, eoff, K);
  FOFB_REQUIRE(E < (1ll << 31), "too many overlap events");
  if (E == 0) return 0;
  unsigned long long* ekey = sc.ekey.alloc((size_t)E);
  unsigned long long* ekey_s = sc.ekey_sorted.alloc((size_t)E);
  k_overlap_events<1><<<wblocks, 128, 0, st>>>(K, centre, sc.kz_sorted.p, sc.kidx_sorted.p, A.ra, A.dec,
                                               A.cos, A.z, A.thvir, p.max_velocity, p.grid_cells, ecount, eoff, ekey);
  FOFB_LAUNCH_CHECK(h);
  FOFB_CUDA(cub::DeviceSegmentedSort::SortKeys(nullptr, bytes, ekey, ekey_s, E, K, eoff, eoff + 1, st));
  FOFB_CUDA(cub::DeviceSegmentedSort::SortKeys(cub_tmp(h, bytes), bytes, ekey, ekey_s, E, K, eoff, eoff + 1, st));
  count_launch(h, 4);
  // sorted member ids per candidate cluster: two stable device-wide radix sorts (by id, then by cluster) -- millions of
  // ids in segments of a hundred are several times faster that way than through a segmented sort
  const long long total = total_members;
  FOFB_REQUIRE(total < (1ll << 31), "too many stage-1 members for the overlap contests");
  double* ids_a = sc.ids.alloc((size_t)total);
  double* ids_s = sc.ids_sorted.alloc((size_t)total);
  int* seg_a = sc.seg_a.alloc((size_t)total);
  int* seg_b = sc.seg_b.alloc((size_t)total);
  k_member_ids_seg<<<grid_for((int64_t)K * 32, 128), 128, 0, st>>>(K, off, rows, S.G, ext_ids, ids_a, seg_a);
  FOFB_LAUNCH_CHECK(h);
  FOFB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, ids_a, ids_s, seg_a, seg_b, (int)total, 0, 64, st));
  FOFB_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp(h, bytes), bytes, ids_a, ids_s, seg_a, seg_b, (int)total, 0, 64, st));
  int seg_bits = 1;
  while ((1 << seg_bits) < K) ++seg_bits;
  FOFB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, seg_b, seg_a, ids_s, ids_a, (int)total, 0, seg_bits, st));
  FOFB_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp(h, bytes), bytes, seg_b, seg_a, ids_s, ids_a, (int)total, 0, seg_bits, st));
  count_launch(h, 12);
  ids_s = ids_a;  // sorted by (cluster, id): the same CSR offsets
  int* ev_a = sc.ev_a.alloc((size_t)E);
  int* ev_c = sc.ev_c.alloc((size_t)E);
  int* ev_l = sc.ev_loser.alloc((size_t)E);
  int* state = sc.state.alloc((size_t)E + K + 8);
  FOFB_PROFILED(h, "k_event_static", k_event_static<<<grid_for(E * 32, 128), 128, 0, st>>>(K, E, eoff, ekey_s, sc.first_by_id.p, centre, A.id,
                                                        A.N, A.mag, off, ids_s, ev_a, ev_c, ev_l, state));
  return E;
}

__global__ void k_state_from_loser(long long E, const int* ev_loser, int* state) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < E) state[e] = ev_loser[e] >= 0 ? 0 : 2;
}

// Stage 2, second half: which clusters survive the ordered contests (ev_* in contest order, device arrays of E entries
// with cluster indices in [0, K)).  Returns K keep flags (device; non-zero = survives).
static int* overlap_resolve(fofb_handle* h, FofState& S, int K, long long E, const int* ev_a, const int* ev_c, const int* ev_l) {
  cudaStream_t st = h->stream;
  OverlapScratch& sc = scratch(S);
  const int B = 256;
  size_t bytes = 0;
  // event states [0, E) followed by the per-cluster keep flags
  int* state = sc.state.alloc((size_t)std::max<long long>(E, 1) + K + 8);
  int* keep_flags = state + std::max<long long>(E, 1);
  if (E > 0) {
    k_state_from_loser<<<grid_for(E, B), B, 0, st>>>(E, ev_l, state);
    FOFB_LAUNCH_CHECK(h);
    // per-cluster lists of the events it would lose, in event order
    unsigned long long* lkey = sc.lkey.alloc((size_t)E);
    unsigned long long* lkey_s = sc.lkey_sorted.alloc((size_t)E);
    k_loser_keys<<<grid_for(E, B), B, 0, st>>>(E, ev_l, K, lkey);
    FOFB_LAUNCH_CHECK(h);
    FOFB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, bytes, lkey, lkey_s, (int)E, 0, 64, st));
    FOFB_CUDA(cub::DeviceRadixSort::SortKeys(cub_tmp(h, bytes), bytes, lkey, lkey_s, (int)E, 0, 64, st));
    count_launch(h, 8);
    long long* lstart = sc.loff.alloc((size_t)K + 3);
    k_loser_starts<<<grid_for(E + 1, B), B, 0, st>>>(E, lkey_s, K, lstart);
    FOFB_LAUNCH_CHECK(h);
    int* und = sc.selcount.alloc(1);
    for (int it = 0;; ++it) {
      FOFB_CUDA(cudaMemsetAsync(und, 0, sizeof(int), st));
      k_event_round<<<grid_for(E, B), B, 0, st>>>(E, ev_a, ev_c, lstart, lkey_s, state, und);
      FOFB_LAUNCH_CHECK(h);
      if (d2h_scalar(h, und) == 0) break;
      if (it > 100000) throw Error(FOFB_ERR_INTERNAL, "overlap contests did not converge");
    }
    k_cluster_keep<<<grid_for(K, B), B, 0, st>>>(K, lstart, lkey_s, state, keep_flags);
    FOFB_LAUNCH_CHECK(h);
  } else {
    FOFB_CUDA(cudaMemsetAsync(keep_flags, 1, sizeof(int) * K, st));  // any non-zero value keeps
  }
  return keep_flags;
}

static int* overlap_keep_flags(fofb_handle* h, FofState& S, const CentreAttrs& A, int K, const int* centre, const long long* off,
                               long long total_members, const int* rows, const double* ext_ids) {
  if (K == 0) return nullptr;
  OverlapScratch& sc = scratch(S);
  const long long E = overlap_events(h, S, A, K, centre, off, total_members, rows, ext_ids);
  return overlap_resolve(h, S, K, E, sc.ev_a.p, sc.ev_c.p, sc.ev_loser.p);
}

void fof_overlap_and_bcg(fofb_handle* h, FofState& S) {
  cudaStream_t st = h->stream;
  OverlapScratch& sc = scratch(S);
  const fofb_params& p = S.params;
  const int K = S.candidates.K;
  S.merged.K = S.bcg.K = 0;
  S.merged.total = S.bcg.total = 0;
  if (K == 0) return;
  int* keep_flags = overlap_keep_flags(h, S, attrs_of(S), K, S.candidates.centre.p, S.candidates.off.p, S.candidates.total, S.candidates.rows.p, nullptr);
  sublist(h, S.candidates, keep_flags, S.merged, sc);
  fof_bcg_check(h, S);
}

// Stage 2 for a cluster list assembled from several processes (device arrays); keep_out: K int32 flags (device).
void fof_overlap_external(fofb_handle* h, FofState& S, int K, const int* centre, const long long* off, long long total,
                          const double* ids, int* keep_out) {
  if (K == 0) return;
  int* keep = overlap_keep_flags(h, S, attrs_of(S), K, centre, off, total, nullptr, ids);
  FOFB_CUDA(cudaMemcpyAsync(keep_out, keep, sizeof(int) * K, cudaMemcpyDeviceToDevice, h->stream));
  FOFB_CUDA(cudaStreamSynchronize(h->stream));
}

// Stage 2 for a cluster list whose centres come as rows (K x 8, device) rather than as indices into this state's
// centre table: used when every process only knows the centres of its own redshift neighbourhood.
void fof_overlap_external_rows(fofb_handle* h, FofState& S, int K, const double* centre_rows, const long long* off, long long total,
                               const double* ids, int* keep_out) {
  if (K == 0) return;
  cudaStream_t st = h->stream;
  OverlapScratch& sc = scratch(S);
  double* buf = sc.attr.alloc((size_t)K * 8);
  int* iota = sc.attr_iota.alloc(K);
  const Cosmo cosmo = device_cosmo(h, S.params);
  k_attrs_from_rows<<<grid_for(K, 256), 256, 0, st>>>(K, centre_rows, cosmo, S.params.virial_radius, buf, buf + K, buf + 2LL * K,
                                                      buf + 3LL * K, buf + 4LL * K, buf + 5LL * K, buf + 6LL * K, buf + 7LL * K, iota);
  FOFB_LAUNCH_CHECK(h);
  const CentreAttrs A{buf, buf + K, buf + 2LL * K, buf + 3LL * K, buf + 4LL * K, buf + 5LL * K, buf + 6LL * K, buf + 7LL * K};
  int* keep = overlap_keep_flags(h, S, A, K, iota, off, total, nullptr, ids);
  FOFB_CUDA(cudaMemcpyAsync(keep_out, keep, sizeof(int) * K, cudaMemcpyDeviceToDevice, st));
  FOFB_CUDA(cudaStreamSynchronize(st));
}

// Distributed stage 2, step 1: the contests of a cluster list that holds this process's clusters and the neighbours'
// boundary clusters (centres as rows, K x 8, in global priority order).  Returns the number of contests; the arrays
// stay in the scratch until fof_overlap_events_fetch copies them out.
long long fof_overlap_events_rows(fofb_handle* h, FofState& S, int K, const double* centre_rows, const long long* off, long long total,
                                  const double* ids) {
  if (K == 0) return 0;
  cudaStream_t st = h->stream;
  OverlapScratch& sc = scratch(S);
  double* buf = sc.attr.alloc((size_t)K * 8);
  int* iota = sc.attr_iota.alloc(K);
  const Cosmo cosmo = device_cosmo(h, S.params);
  k_attrs_from_rows<<<grid_for(K, 256), 256, 0, st>>>(K, centre_rows, cosmo, S.params.virial_radius, buf, buf + K, buf + 2LL * K,
                                                      buf + 3LL * K, buf + 4LL * K, buf + 5LL * K, buf + 6LL * K, buf + 7LL * K, iota);
  FOFB_LAUNCH_CHECK(h);
  const CentreAttrs A{buf, buf + K, buf + 2LL * K, buf + 3LL * K, buf + 4LL * K, buf + 5LL * K, buf + 6LL * K, buf + 7LL * K};
  return overlap_events(h, S, A, K, iota, off, total, nullptr, ids);
}

void fof_overlap_events_fetch(fofb_handle* h, FofState& S, long long E, int* ev_a, int* ev_c, int* ev_loser) {
  if (E <= 0) return;
  OverlapScratch& sc = scratch(S);
  cudaStream_t st = h->stream;
  FOFB_CUDA(cudaMemcpyAsync(ev_a, sc.ev_a.p, sizeof(int) * E, cudaMemcpyDeviceToDevice, st));
  FOFB_CUDA(cudaMemcpyAsync(ev_c, sc.ev_c.p, sizeof(int) * E, cudaMemcpyDeviceToDevice, st));
  FOFB_CUDA(cudaMemcpyAsync(ev_loser, sc.ev_loser.p, sizeof(int) * E, cudaMemcpyDeviceToDevice, st));
}

// Distributed stage 2, step 2: the all-process contest list (global contest order, cluster indices into the all-process
// cluster list) resolved on every process alike.
void fof_overlap_resolve_external(fofb_handle* h, FofState& S, int K, long long E, const int* ev_a, const int* ev_c, const int* ev_loser,
                                  int* keep_out) {
  if (K == 0) return;
  int* keep = overlap_resolve(h, S, K, E, ev_a, ev_c, ev_loser);
  FOFB_CUDA(cudaMemcpyAsync(keep_out, keep, sizeof(int) * K, cudaMemcpyDeviceToDevice, h->stream));
}

// merged := the local candidates whose flag is set (flags on the device, one per local candidate)
void fof_set_merged(fofb_handle* h, FofState& S, const int* keep_local) {
  S.merged.K = S.bcg.K = 0;
  S.merged.total = S.bcg.total = 0;
  if (S.candidates.K == 0) return;
  sublist(h, S.candidates, keep_local, S.merged, scratch(S));
  fof_bcg_check(h, S);
}

void fof_bcg_check(fofb_handle* h, FofState& S) {
  cudaStream_t st = h->stream;
  OverlapScratch& sc = scratch(S);
  const fofb_params& p = S.params;
  // stage 3: BCG recentring (fof.py:318-371).  Under the pipeline's column layout the replacement
  // test compares |z_lower| with |RMag| and never fires; the device path verifies that.
  if (S.merged.K > 0) {
    int* fired = sc.selcount.alloc(1);
    FOFB_CUDA(cudaMemsetAsync(fired, 0, sizeof(int), st));
    k_bcg_check<<<grid_for((int64_t)S.merged.K * 32, 128), 128, 0, st>>>(
        S.merged.K, S.merged.centre.p, S.merged.off.p, S.merged.rows.p, S.G, S.cat_().ra_row.p, S.cat_().dec_row.p, S.cos_row.p,
        S.c_ra.p, S.c_dec.p, S.c_cos.p, S.c_thvir.p, S.c_mag.p, S.c_id.p, p.bcg_fraction, p.grid_cells, fired);
    FOFB_LAUNCH_CHECK(h);
    if (d2h_scalar(h, fired) != 0)
      throw Error(FOFB_ERR_UNSUPPORTED,
                  "BCG recentring (fof.py:346-367) would replace a centre: |column 3| exceeds |column 5| for a member "
                  "near a cluster centre. The device path implements the pipeline column layout, where this never happens.");
  }
  // bcg_clusters == merged clusters (same order)
  S.bcg.K = S.merged.K;
  S.bcg.total = S.merged.total;
}

// ---- stage 4 ----------------------------------------------------------------------------------------------------------
// N := number of OWN members within theta_0.5 of the centre (fof.py:377-379; helper.py:181-188)
__global__ void k_own_count(int K, const int* centre, const long long* off, const int* rows, const double* ra_row,
                            const double* dec_row, const double* cos_row, const double* c_ra, const double* c_dec,
                            const double* c_cos, const double* c_th05, int grid_cells, int* N_out) {
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (k >= K) return;
  const int c = centre[k];
  const long long s0 = off[k], s1 = off[k + 1];
  double4 bb = make_double4(INFINITY, -INFINITY, INFINITY, -INFINITY);
  for (long long t = s0 + lane; t < s1; t += 32) {
    const int row = rows[t];
    bb = bbox_merge(bb, make_double4(ra_row[row], ra_row[row], dec_row[row], dec_row[row]));
  }
  for (int o = 16; o; o >>= 1) {
    double4 other;
    other.x = __shfl_xor_sync(0xffffffffu, bb.x, o); other.y = __shfl_xor_sync(0xffffffffu, bb.y, o);
    other.z = __shfl_xor_sync(0xffffffffu, bb.z, o); other.w = __shfl_xor_sync(0xffffffffu, bb.w, o);
    bb = bbox_merge(bb, other);
  }
  const BBox b{bb.x, bb.y, bb.z, bb.w};
  const double ra = c_ra[c], dec = c_dec[c], r = c_th05[c];
  const bool oof = (ra - r > b.ra_max + 1e-6) || (ra + r < b.ra_min - 1e-6) || (dec - r > b.dec_max + 1e-6) ||
                   (dec + r < b.dec_min - 1e-6);
  const Box box = grid_box(b, grid_cells, ra, dec, r);
  BubbleTest bt;
  bt.init(ra, dec, c_cos[c], r);
  int cnt = 0;
  if (!oof)
    for (long long t = s0 + lane; t < s1; t += 32) {
      const int row = rows[t];
      if (box.contains(ra_row[row], dec_row[row]) && bt.inside(ra_row[row], dec_row[row], cos_row[row])) ++cnt;
    }
  for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) N_out[k] = cnt;
}

// per cluster: window bbox and the "all 300 points out of field" flag (helper.py:207-213, App. B.1), packed with the
// radius and the window for the point queries.  whi == wlo: nothing to count.
struct OdCluster {
  double4 bbox;
  double r;
  double inv_wx, inv_wy;  // grid_cells / (bins[-1] - bins[0]) of the window's GriSPy grid, per axis
  int wlo, whi;
  float T_lo, T_hi;       // fp32 decision thresholds of the tiled count (tile.cuh), the same for the cluster's 300 points
  float pad[2];
};

__global__ void k_od_prep(int K, int npts, const int* centre, CatalogView cat, const int* c_wlo, const int* c_whi,
                          const double* c_th05, const double* rra, const double* rdec, int grid_cells, double cell_deg,
                          OdCluster* oc, unsigned long long* sums) {
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (k >= K) return;
  const int c = centre[k];
  const int lo = c_wlo[c], hi = c_whi[c];
  BBox b{0, 0, 0, 0};
  if (lane == 0 && hi > lo) b = range_bbox(cat, lo, hi);
  b.ra_min = __shfl_sync(0xffffffffu, b.ra_min, 0); b.ra_max = __shfl_sync(0xffffffffu, b.ra_max, 0);
  b.dec_min = __shfl_sync(0xffffffffu, b.dec_min, 0); b.dec_max = __shfl_sync(0xffffffffu, b.dec_max, 0);
  const double r = c_th05[c];
  bool all = true;
  for (int t = lane; t < npts; t += 32) {
    const double ra = rra[(long long)k * npts + t], dec = rdec[(long long)k * npts + t];
    const bool oof = (ra - r > b.ra_max + 1e-6) || (ra + r < b.ra_min - 1e-6) || (dec - r > b.dec_max + 1e-6) ||
                     (dec + r < b.dec_min - 1e-6);
    if (!oof) all = false;
  }
  all = __all_sync(0xffffffffu, all);
  if (lane == 0) {
    OdCluster o;
    o.bbox = make_double4(b.ra_min, b.ra_max, b.dec_min, b.dec_max);
    o.r = r;
    o.inv_wx = (double)grid_cells / ((b.ra_max + 1.0e-6) - (b.ra_min - 1.0e-6));
    o.inv_wy = (double)grid_cells / ((b.dec_max + 1.0e-6) - (b.dec_min - 1.0e-6));
    tile_thresholds(r, cell_deg, &o.T_lo, &o.T_hi);
    o.pad[0] = o.pad[1] = 0.f;
    sums[2 * k] = sums[2 * k + 1] = 0ull;
    o.wlo = lo;
    o.whi = (all || hi <= lo) ? lo : hi;
    oc[k] = o;
  }
}

// cell of every query point, so that the queries can be processed in cell order (the 300 points of one cluster are
// scattered over the whole footprint), and what the count needs of the point as a 16-byte record that travels through
// the sort with the key: offsets from the cell corner and cos(dec) in fp32, and the point's index (zr field).
__global__ void k_od_keys(long long total, CellListView cl, const double* rra, const double* rdec, unsigned* key, PackedPoint* val) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= total) return;
  const double ra = rra[q], dec = rdec[q];
  const int cx = cl.cx_of(ra), cy = cl.cy_of(dec);
  key[q] = (unsigned)(cy * cl.ncx + cx);
  PackedPoint p;
  p.fx = (float)(ra - (cl.ra0 + (double)cx * cl.s_ra));
  p.fy = (float)(dec - (cl.dec0 + (double)cy * cl.s_dec));
  p.fcos = (float)cos(__dmul_rn(dec, kDeg2Rad));
  p.zr = (int)q;
  val[q] = p;
}

// first query (in cell order) of every strip: qstart[strip] = lower bound of the strip's first cell key, qstart[nstrips] = total
__global__ void k_od_strip_starts(int nstrips, int nsx, int ncx, long long total, const unsigned* __restrict__ qkey,
                                  int* __restrict__ qstart) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s > nstrips) return;
  long long lo = 0, hi = total;
  if (s < nstrips) {
    const unsigned key = (unsigned)((s / nsx) * ncx + (s % nsx) * kCtW);
    while (lo < hi) {
      const long long mid = (lo + hi) >> 1;
      if (__ldg(qkey + mid) < key) lo = mid + 1; else hi = mid;
    }
  } else {
    lo = total;
  }
  qstart[s] = (int)lo;
}

// Galaxies of the centre's velocity window within theta_0.5 of each random point (helper.py:207-216).  One CTA per
// strip of cells, as in k_count (counts.cu): the strip's neighbourhood is staged in shared memory once with a bulk
// asynchronous copy and serves every query point that fell into the strip -- the points of many different clusters,
// each with its own window.  A point arrives as the 16-byte record the sort carried along (sequential reads); its
// full-precision coordinates are fetched only by the few points that need them (a pair inside the fp32 band, a GriSPy
// box that may bind).  The counts are not stored: D needs their sum and sum of squares per cluster (helper.py:216-221),
// exact integers whatever the order, so each point adds its two terms to its cluster's accumulators.
__global__ void __launch_bounds__(kCtThreads, kCtResident) k_od_count(CellListView cl, int nsx, int npts, const int* __restrict__ qstart,
                                                            const unsigned* __restrict__ qkey, const PackedPoint* __restrict__ pay,
                                                            const OdCluster* __restrict__ oc, const double* __restrict__ rra,
                                                            const double* __restrict__ rdec, int grid_cells,
                                                            unsigned long long* __restrict__ sums) {
  extern __shared__ __align__(128) unsigned char ct_smem[];
  __shared__ int s_cs[3][kCtW + 3];
  __shared__ int s_delta[3];
  __shared__ int s_fits;
  __shared__ __align__(8) unsigned long long s_bar;
  Tile t;
  t.cs = s_cs;
  t.delta = s_delta;
  const int y = blockIdx.x / nsx, x0 = (blockIdx.x % nsx) * kCtW;
  // a row's strips are consecutive key ranges except the last, which ends with the row: the next row's first strip
  // starts exactly there, so [qstart[b], qstart[b + 1]) is this strip's slice of the cell-sorted queries
  const int q0 = qstart[blockIdx.x], q1 = qstart[blockIdx.x + 1];
  if (q1 <= q0) return;
  tile_stage(cl, y, x0, t, reinterpret_cast<PackedPoint*>(ct_smem),
             reinterpret_cast<unsigned short*>(ct_smem + (size_t)kCtMaxPts * sizeof(PackedPoint)), &s_bar, &s_fits);
  const int xb = t.xa + t.ncs - 2;
  const float sxf = (float)cl.s_ra, syf = (float)cl.s_dec;
  for (int ti = q0 + (int)threadIdx.x; ti < q1; ti += kCtThreads) {
    const PackedPoint P = pay[ti];
    const int qid = P.zr, k = qid / npts;
    const OdCluster o = oc[k];
    if (o.whi <= o.wlo) continue;
    const int cx = (int)qkey[ti] - y * cl.ncx;
    const double r = o.r;
    // the point from its record: exact to ~1e-8 deg, enough for everything that carries a margin
    const double ra_a = cl.ra0 + (double)cx * cl.s_ra + (double)P.fx, dec_a = cl.dec0 + (double)y * cl.s_dec + (double)P.fy;
    const double reach = ra_reach(r, dec_a);
    const float reach_f = (float)reach, rdec_f = (float)(r * 1.00001);
    int cnt = 0;
    // neighbour cells the bubble can reach (1 % margin in reach, 1e-5 in dec: the record's rounding is 1e-6 of that)
    const bool wide = reach >= cl.s_ra || r * 1.00001 >= cl.s_dec || P.fx < -sxf || P.fx > 2.f * sxf || P.fy < -syf || P.fy > 2.f * syf;
    // GriSPy's cell box: cannot bind for almost every point -- decided on the approximate point with margins a
    // thousand times its rounding; otherwise the exact routine on the exact point
    GridAxis gx, gy;
    gx.init(o.bbox.x, o.bbox.y, grid_cells);
    gy.init(o.bbox.z, o.bbox.w, grid_cells);
    const double dd = fmin(89.0, fabs(dec_a) + r);
    const double reach_box = 1.01 * r / cos(dd * kDeg2Rad) + 1e-12;
    bool boxed = false;
    Box bx{-INFINITY, INFINITY, -INFINITY, INFINITY};
    if (!(grid_axis_unbound_fast(gx, o.inv_wx, ra_a, r, reach_box, 1.0e-4) &&
          grid_axis_unbound_fast(gy, o.inv_wy, dec_a, r, r * (1.0 + 1e-9) + 1e-12, 1.0e-4))) {
      bx = grid_box(BBox{o.bbox.x, o.bbox.y, o.bbox.z, o.bbox.w}, grid_cells, rra[qid], rdec[qid], r);
      boxed = bx.ra_lo != -INFINITY || bx.ra_hi != INFINITY || bx.dec_lo != -INFINITY || bx.dec_hi != INFINITY;
    }
    auto exact = [&](double pra, double pdec, double pcos) -> bool {
      const double ra = rra[qid], dec = rdec[qid];
      BubbleTest bt;
      bt.init(ra, dec, cos(__dmul_rn(dec, kDeg2Rad)), r);
      return bt.inside(pra, pdec, pcos);
    };
    const int cx0 = (P.fx - reach_f < 0.f && cx > 0) ? cx - 1 : cx, cx1 = (P.fx + reach_f >= sxf && cx < cl.ncx - 1) ? cx + 1 : cx;
    const int cy0 = (P.fy - rdec_f < 0.f && y > 0) ? y - 1 : y, cy1 = (P.fy + rdec_f >= syf && y < cl.ncy - 1) ? y + 1 : y;
    if (!wide && cx0 >= t.xa && cx1 <= xb) {
      TileQuery q;
      q.cx = cx;
      q.cy = y;
      q.fx = P.fx; q.fy = P.fy; q.fcos = P.fcos;
      q.wlo = o.wlo;
      q.whi = o.whi;
      q.T_lo = o.T_lo; q.T_hi = o.T_hi;
      cnt = tile_count(cl, t, y, q, cy0 - (y - 1), cy1 - (y - 1), cx0, cx1, exact, boxed, bx);
    } else {  // a bubble wider than one cell, or a point far outside the cell list's footprint: walk the cell list
      const double ra = rra[qid], dec = rdec[qid];
      const Box bxe = grid_box(BBox{o.bbox.x, o.bbox.y, o.bbox.z, o.bbox.w}, grid_cells, ra, dec, r);
      BubbleTest bt;
      bt.init(ra, dec, cos(__dmul_rn(dec, kDeg2Rad)), r);
      const double rch = ra_reach(r, dec);
      const int ex0 = cl.cx_of(ra - rch), ex1 = cl.cx_of(ra + rch);
      const int ey0 = cl.cy_of(dec - r * 1.000001), ey1 = cl.cy_of(dec + r * 1.000001);
      for (int yy = ey0; yy <= ey1; ++yy)
        for (int xx = ex0; xx <= ex1; ++xx) {
          const int s = cl.cell_start[yy * cl.ncx + xx], e = cl.cell_start[yy * cl.ncx + xx + 1];
          const int a = lower_bound_zr(cl.zrank, s, e, o.wlo);
          const int b = lower_bound_zr(cl.zrank, a, e, o.whi);
          for (int j = a; j < b; ++j) {
            const double pra = cl.ra[j], pdec = cl.dec[j];
            if (bxe.contains(pra, pdec) && bt.inside(pra, pdec, cl.cosdec[j])) ++cnt;
          }
        }
    }
    if (cnt) {
      atomicAdd(sums + 2 * k, (unsigned long long)cnt);
      atomicAdd(sums + 2 * k + 1, (unsigned long long)cnt * (unsigned long long)cnt);
    }
  }
}

// D = (N - mean) / rms over the 300 counts (helper.py:216-221), and the final selection (fof.py:385-390)
__global__ void k_od_finish(int K, int npts, const unsigned long long* sums, const int* N_own, const long long* off,
                            double overdensity, int min_count, int richness, double* D_out, int* keep) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const long long s1 = (long long)sums[2 * k], s2 = (long long)sums[2 * k + 1];
  {
    const double mean = __ddiv_rn((double)s1, (double)npts);
    const double rms = sqrt(__ddiv_rn((double)s2, (double)npts));
    const double D = __ddiv_rn(__dsub_rn((double)N_own[k], mean), rms);
    D_out[k] = D;
    keep[k] = (N_own[k] >= min_count && (off[k + 1] - off[k]) >= richness && D >= overdensity) ? 1 : 0;
  }
}

__global__ void k_gather_nd(int K2, const int* sel, const int* N_in, const double* D_in, int* N_out, double* D_out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K2) return;
  N_out[k] = N_in[sel[k]];
  D_out[k] = D_in[sel[k]];
}

// helper.py:197-202: np.random.uniform(low, high) == low + (high - low) * next_double, separate roundings
__global__ void k_scale_points(long long total, int npts, const double* u, double ra_lo, double ra_range, double dec_lo,
                               double dec_range, double* ra, double* dec) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= total) return;
  const long long k = q / npts, i = q % npts;
  ra[q] = __dadd_rn(ra_lo, __dmul_rn(ra_range, u[(2 * k) * npts + i]));
  dec[q] = __dadd_rn(dec_lo, __dmul_rn(dec_range, u[(2 * k + 1) * npts + i]));
}

// ---- numpy's legacy MT19937 on the device ------------------------------------------------------------------
// center_overdensity (helper.py:196-202) consumes np.random's global MT19937 stream: 2 words per double,
// (a >> 5, b >> 6) -> (a * 2^26 + b) / 2^53.  k_mt19937_chunks writes the tempered words (one CTA per chunk of the
// stream, after a polynomial jump to the chunk's start); a second kernel turns them into points.
__device__ __forceinline__ unsigned mt_twist(unsigned u, unsigned v) {
  const unsigned y = (u & 0x80000000u) | (v & 0x7fffffffu);
  return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

__device__ __forceinline__ unsigned mt_temper(unsigned y) {
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

// One regeneration: raw block `nx` from raw block `c` (three dependent phases of 227 / 227 / 170 words).
__device__ __forceinline__ void mt_next_block(const unsigned* c, unsigned* nx, int tid) {
  if (tid < 227) nx[tid] = c[tid + 397] ^ mt_twist(c[tid], c[tid + 1]);
  __syncthreads();
  if (tid < 227) nx[227 + tid] = nx[tid] ^ mt_twist(c[227 + tid], c[228 + tid]);
  __syncthreads();
  if (tid < 170) {
    const int k = 454 + tid;
    nx[k] = nx[k - 227] ^ mt_twist(c[k], k == 623 ? nx[0] : c[k + 1]);
  }
  __syncthreads();
}

constexpr int kMtThreads = 640;
constexpr int kMtSeqBlocks = 33;  // 19937 + 623 raw words fit in 33 blocks
constexpr size_t kMtSmemBytes = sizeof(unsigned) * (kMtSeqBlocks * 624 + 624);

// Generates `nblocks` new state blocks after `key_in` (624 raw words) and writes their tempered words to
// out[0 .. 624*nblocks); when first == 1 the unread tail key_in[pos .. 624) of the current block is tempered into
// head[0 .. 624-pos) first.  The raw state after the last block goes to key_out (distinct from key_in).
//
// The stream is cut into chunks of kMtChunkBlocks blocks, one CTA each.  CTA c first jumps its copy of the state
// c * kMtChunkBlocks blocks ahead: the transition is linear over GF(2), so the block J words ahead is g(A) x with
// g = t^J mod (minimal polynomial), evaluated as a correlation -- word j of the jumped block is the XOR over the set
// coefficients i of g of raw word x[i + j] of the sequence starting at the current block (33 blocks of it, in
// shared memory).  kMtJump[k] holds g for 2^k chunks (scripts/make_mt_jump.py, checked against numpy); the binary
// digits of c are applied one after the other.  The jumped block may differ from the true one in the low 31 bits of
// word 0, which never reach an output or a later block.
// A launch may cover only the chunks [chunk0, chunk0 + gridDim.x) of nchunks (slab runs: every rank generates its share of
// the stream and the shares are all-gathered); the head words are written by the first CTA of every launch.
__global__ void __launch_bounds__(kMtThreads) k_mt19937_chunks(const unsigned* key_in, unsigned* key_out, int pos, int first,
                                                               unsigned* head, long long nblocks, unsigned* out, long long chunk0,
                                                               long long nchunks) {
  extern __shared__ unsigned mt_smem[];
  unsigned* seq = mt_smem;
  unsigned* poly = mt_smem + kMtSeqBlocks * 624;
  const int tid = threadIdx.x;
  const long long c = blockIdx.x + chunk0;
  if (tid < 624) seq[tid] = key_in[tid];
  __syncthreads();
  if (blockIdx.x == 0 && first)
    for (int k = pos + tid; k < 624; k += kMtThreads) head[k - pos] = mt_temper(seq[k]);
  for (int k = 0; (c >> k) != 0; ++k) {
    if (!((c >> k) & 1)) continue;
    for (int b = 0; b + 1 < kMtSeqBlocks; ++b) mt_next_block(seq + b * 624, seq + (b + 1) * 624, tid);
    if (tid < 624) poly[tid] = kMtJump[k][tid];
    __syncthreads();
    unsigned acc = 0;
    if (tid < 624) {
      for (int w = 0; w < 624; ++w) {
        unsigned bits = poly[w];
        const unsigned* base = seq + w * 32 + tid;
        while (bits) {
          acc ^= base[__ffs(bits) - 1];
          bits &= bits - 1;
        }
      }
    }
    __syncthreads();
    if (tid < 624) seq[tid] = acc;
    __syncthreads();
  }
  const long long b0 = c * kMtChunkBlocks;
  const long long b1 = c == nchunks - 1 ? nblocks : b0 + kMtChunkBlocks;
  int cur = 0;
  for (long long b = b0; b < b1; ++b) {
    const unsigned* cb = seq + cur * 624;
    unsigned* nx = seq + (cur ^ 1) * 624;
    mt_next_block(cb, nx, tid);
    if (tid < 624) out[b * 624 + tid] = mt_temper(nx[tid]);
    cur ^= 1;
  }
  if (c == nchunks - 1 && tid < 624) key_out[tid] = seq[cur * 624 + tid];
}

// key_in = mt_key[0 .. 624) (or the previous final state), key_out = mt_key[1248 .. 1872)
static void launch_mt19937(fofb_handle* h, cudaStream_t st, const unsigned* key_in, unsigned* key_out, int pos, int first,
                           unsigned* head, long long nblocks, unsigned* out, int part = 0, int parts = 1) {
  // per device, so set on every launch (a handle may live on any device of the process)
  FOFB_CUDA(cudaFuncSetAttribute(k_mt19937_chunks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMtSmemBytes));
  long long nchunks = (nblocks + kMtChunkBlocks - 1) / kMtChunkBlocks;
  nchunks = std::min<long long>(std::max<long long>(nchunks, 1), 1ll << kMtJumpTables);
  const long long per = nchunks / parts;  // parts > 1: nblocks is a multiple of parts * kMtChunkBlocks (mt_prefetch)
  k_mt19937_chunks<<<(unsigned)per, kMtThreads, kMtSmemBytes, st>>>(key_in, key_out, pos, first, head, nblocks, out, part * per, nchunks);
  count_launch(h);
  FOFB_CUDA(cudaGetLastError());
}

// raw state block from its tempered words (the tempering transform is a bijection)
__global__ void k_mt_untemper(const unsigned* words, unsigned* key) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= 624) return;
  unsigned y = words[k];
  y ^= y >> 18;
  y ^= (y << 15) & 0xefc60000u;
  y ^= ((y << 7) & 0x9d2c5680u) ^ ((y << 14) & 0x94284000u) ^ ((y << 21) & 0x14200000u) ^ ((y << 28) & 0x10000000u);
  y ^= (y >> 11) ^ (y >> 22);
  key[k] = y;
}

__global__ void k_points_from_words(long long total, int npts, const unsigned* w, long long head, const int* gidx,
                                    double ra_lo, double ra_range, double dec_lo, double dec_range, double* ra, double* dec) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= total) return;
  // cluster q / npts consumes block k of the stream: its own index, or its index in the global (all-process) list
  const long long k = gidx ? gidx[q / npts] : q / npts, i = q % npts;
  const long long ia = ((2 * k) * npts + i) * 2, id = ((2 * k + 1) * npts + i) * 2;
  auto word = [&](long long t) -> unsigned { return t < head ? w[t] : w[624 + (t - head)]; };
  const double ua = ((double)(word(ia) >> 5) * 67108864.0 + (double)(word(ia + 1) >> 6)) / 9007199254740992.0;
  const double ud = ((double)(word(id) >> 5) * 67108864.0 + (double)(word(id + 1) >> 6)) / 9007199254740992.0;
  ra[q] = __dadd_rn(ra_lo, __dmul_rn(ra_range, ua));
  dec[q] = __dadd_rn(dec_lo, __dmul_rn(dec_range, ud));
}

// The next n_words 32-bit outputs of numpy's legacy generator in state (key, pos), and the state after them
// (fofb_mt19937_words: the device generator on its own, for checks against np.random).
void mt19937_words(fofb_handle* h, unsigned* key_host, int* pos_host, long long n_words, unsigned* out, int out_mem) {
  FOFB_REQUIRE(key_host && pos_host && *pos_host >= 0 && *pos_host <= 624 && n_words >= 0, "bad MT19937 state");
  if (n_words == 0) return;
  cudaStream_t st = h->stream;
  const int pos = *pos_host;
  const long long head = 624 - pos;
  const long long nblocks = n_words > head ? (n_words - head + 623) / 624 : 0;
  DBuf<unsigned> key, words;
  key.alloc(624 * 3);
  words.alloc((size_t)(624 + std::max<long long>(nblocks, 1) * 624));
  FOFB_CUDA(cudaMemcpyAsync(key.p, key_host, sizeof(unsigned) * 624, cudaMemcpyHostToDevice, st));
  // tempered words are laid out contiguously: the unread tail of the current block, then the new blocks
  unsigned* base = words.p + pos;  // so that word t of the request is base[t]
  launch_mt19937(h, st, key.p, key.p + 624, pos, 1, base, nblocks, base + head);
  FOFB_CUDA(cudaMemcpyAsync(out, base, sizeof(unsigned) * (size_t)n_words, out_mem == 1 ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
  if (n_words <= head) {
    *pos_host = pos + (int)n_words;
  } else {
    const long long rem = n_words - head;
    FOFB_CUDA(cudaMemcpyAsync(key_host, key.p + 624, sizeof(unsigned) * 624, cudaMemcpyDeviceToHost, st));
    *pos_host = (int)(rem - (nblocks - 1) * 624);
  }
  FOFB_CUDA(cudaStreamSynchronize(st));
}

// Start generating the stream on a side stream while stages 0-3 run: enough words for k_spec clusters.
void FofState::mt_prefetch(fofb_handle* h, const unsigned* key_host, int pos, long long k_spec, int part, int parts) {
  FOFB_REQUIRE(key_host && pos >= 0 && pos <= 624, "bad MT19937 state");
  FOFB_REQUIRE(parts >= 1 && part >= 0 && part < parts, "bad MT19937 share");
  if (!rng_stream) {
    FOFB_CUDA(cudaStreamCreateWithFlags(&rng_stream, cudaStreamNonBlocking));
    FOFB_CUDA(cudaEventCreateWithFlags(&rng_done, cudaEventDisableTiming));
  }
  mt_pos0 = pos;
  mt_head = 624 - pos;
  const long long W = std::max<long long>(k_spec, 1) * params_n_random_hint * 4;
  mt_blocks = std::max<long long>((W - mt_head + 623) / 624, 1);
  if (parts > 1) {  // equal shares of whole chunks; too long a stream for the jump tables: everybody generates everything
    const long long unit = (long long)parts * kMtChunkBlocks;
    const long long padded = (mt_blocks + unit - 1) / unit * unit;
    if (padded / kMtChunkBlocks <= (1ll << kMtJumpTables)) mt_blocks = padded; else { part = 0; parts = 1; }
  }
  mt_part = part;
  mt_parts = parts;
  unsigned* key = mt_key.alloc(624 * 4);  // [0] input state, [1] untempered output, [2] state after the stream, [3] scratch
  unsigned* words = mt_words.alloc((size_t)(624 + mt_blocks * 624));
  std::memcpy(mt_key_host, key_host, sizeof(unsigned) * 624);
  FOFB_CUDA(cudaMemcpyAsync(key, mt_key_host, sizeof(unsigned) * 624, cudaMemcpyHostToDevice, rng_stream));
  launch_mt19937(h, rng_stream, key, key + 1248, pos, 1, words, mt_blocks, words + 624, part, parts);
  FOFB_CUDA(cudaEventRecord(rng_done, rng_stream));
  mt_prefetched = true;
}

// word t of the consumed stream lives at: t < head ? words[t] : words[624 + (t - head)]
void FofState::finish_mt(fofb_handle* h, unsigned* key_host, int* pos_host) {
  if (!ran) throw Error(FOFB_ERR_STATE, "fofb_fof_finish called before fofb_fof_run");
  const int K = merged.K;
  if (K == 0 && K_glob_merged <= 0) {
    mt_prefetched = false;
    finish(h, nullptr, nullptr, 1);
    return;
  }
  FOFB_REQUIRE(key_host && pos_host && *pos_host >= 0 && *pos_host <= 624, "bad MT19937 state");
  cudaStream_t st = h->stream;
  const int npts = params.n_random;
  const long long total = (long long)K * npts;
  const long long K_stream = K_glob_merged >= 0 ? K_glob_merged : K;  // distributed: the stream covers every process's clusters
  const long long W = K_stream * npts * 4;  // 2 coordinates x 2 words
  if (!mt_prefetched || mt_pos0 != *pos_host || std::memcmp(mt_key_host, key_host, sizeof(unsigned) * 624) != 0) {
    params_n_random_hint = npts;
    mt_prefetch(h, key_host, *pos_host, K_stream);
  }
  const bool trace = getenv("FOFB_TRACE") != nullptr;
  timespec t_a, t_b;
  if (trace) { cudaStreamSynchronize(st); clock_gettime(CLOCK_MONOTONIC, &t_a); }
  FOFB_CUDA(cudaStreamWaitEvent(st, rng_done, 0));
  if (trace) {
    cudaStreamSynchronize(st);
    clock_gettime(CLOCK_MONOTONIC, &t_b);
    fprintf(stderr, "[trace]   waited %.2f ms for the MT19937 stream (%lld blocks prefetched)\n",
            (t_b.tv_sec - t_a.tv_sec) * 1e3 + (t_b.tv_nsec - t_a.tv_nsec) * 1e-6, mt_blocks);
  }
  const long long need_blocks = W > mt_head ? (W - mt_head + 623) / 624 : 0;
  if (need_blocks > mt_blocks && mt_parts > 1) {
    // a shared stream that fell short: the state after its last block lives on another process; generate the stream
    // here in full (rare: the speculation is the previous run's cluster count plus a quarter)
    mt_prefetch(h, key_host, *pos_host, K_stream, 0, 1);
    FOFB_CUDA(cudaStreamWaitEvent(st, rng_done, 0));
  }
  if (need_blocks > mt_blocks) {  // the speculation fell short: continue from the last generated state
    const long long extra = need_blocks - mt_blocks;
    DBuf<unsigned> bigger;
    bigger.alloc((size_t)(624 + need_blocks * 624));
    FOFB_CUDA(cudaMemcpyAsync(bigger.p, mt_words.p, sizeof(unsigned) * (size_t)(624 + mt_blocks * 624), cudaMemcpyDeviceToDevice, st));
    FOFB_CUDA(cudaStreamSynchronize(st));
    std::swap(bigger.p, mt_words.p);
    std::swap(bigger.cap, mt_words.cap);
    launch_mt19937(h, st, mt_key.p + 1248, mt_key.p + 1872, 624, 0, nullptr, extra, mt_words.p + 624 + mt_blocks * 624);
    FOFB_CUDA(cudaMemcpyAsync(mt_key.p + 1248, mt_key.p + 1872, sizeof(unsigned) * 624, cudaMemcpyDeviceToDevice, st));
    mt_blocks = need_blocks;
  }
  double* pts = unit_pts.alloc((size_t)total * 2);
  // the box helper.py:197-202 draws from: min/max of galaxy_data (all processes' FoF() rows when distributed)
  const Catalog& cat = cat_();
  const BBox pb = have_global_bbox ? global_bbox : cat.gbox;
  const double ra_range = pb.ra_max - pb.ra_min, dec_range = pb.dec_max - pb.dec_min;
  const double ra_lo = pb.ra_min, dec_lo = pb.dec_min;
  if (total > 0) {
    k_points_from_words<<<grid_for(total, 256), 256, 0, st>>>(total, npts, mt_words.p, mt_head,
                                                              K_glob_merged >= 0 ? merged_gidx.p : nullptr, ra_lo, ra_range,
                                                              dec_lo, dec_range, pts, pts + total);
    FOFB_LAUNCH_CHECK(h);
  }
  // the generator state after exactly W words
  if (W <= mt_head) {
    *pos_host += (int)W;
  } else {
    const long long rem = W - mt_head;
    const long long last = (rem - 1) / 624;  // index of the last block touched
    unsigned* key_out = mt_key.p + 624;
    k_mt_untemper<<<3, 256, 0, st>>>(mt_words.p + 624 + last * 624, key_out);
    FOFB_LAUNCH_CHECK(h);
    FOFB_CUDA(cudaMemcpyAsync(key_host, key_out, sizeof(unsigned) * 624, cudaMemcpyDeviceToHost, st));
    *pos_host = (int)(rem - last * 624);
  }
  mt_prefetched = false;
  if (K > 0) finish(h, pts, pts + total, 1);
  else finish(h, nullptr, nullptr, 1);
  FOFB_CUDA(cudaStreamSynchronize(st));
}

void FofState::finish_unit(fofb_handle* h, const double* u, int mem) {
  if (!ran) throw Error(FOFB_ERR_STATE, "fofb_fof_finish called before fofb_fof_run");
  const int K = merged.K;
  if (K == 0) {
    finish(h, nullptr, nullptr, 1);
    return;
  }
  FOFB_REQUIRE(u != nullptr, "uniform samples are null");
  cudaStream_t st = h->stream;
  const int npts = params.n_random;
  const long long total = (long long)K * npts;
  const double* du = u;
  if (mem == 0) {
    double* a = h->upload2.alloc((size_t)total * 2);
    FOFB_CUDA(cudaMemcpyAsync(a, u, sizeof(double) * total * 2, cudaMemcpyHostToDevice, st));
    du = a;
  }
  double* pts = unit_pts.alloc((size_t)total * 2);
  const Catalog& cat = cat_();
  // volatile-free host arithmetic: high - low is one rounding, as in numpy
  const double ra_range = cat.gbox.ra_max - cat.gbox.ra_min, dec_range = cat.gbox.dec_max - cat.gbox.dec_min;
  k_scale_points<<<grid_for(total, 256), 256, 0, st>>>(total, npts, du, cat.gbox.ra_min, ra_range, cat.gbox.dec_min,
                                                       dec_range, pts, pts + total);
  FOFB_LAUNCH_CHECK(h);
  finish(h, pts, pts + total, 1);
}

void FofState::finish(fofb_handle* h, const double* rand_ra, const double* rand_dec, int mem) {
  cudaStream_t st = h->stream;
  OverlapScratch& sc = scratch(*this);
  const fofb_params& p = params;
  if (!ran) throw Error(FOFB_ERR_STATE, "fofb_fof_finish called before fofb_fof_run");
  final_.K = 0;
  final_.total = 0;
  finished = true;
  const int K = merged.K;
  if (K == 0) return;
  const int npts = p.n_random;
  FOFB_REQUIRE(npts > 0, "n_random must be positive");
  FOFB_REQUIRE(rand_ra && rand_dec, "random points are null");
  const long long total = (long long)K * npts;
  const double *rra = rand_ra, *rdec = rand_dec;
  if (mem == 0) {
    double* a = h->upload2.alloc((size_t)total * 2);
    FOFB_CUDA(cudaMemcpyAsync(a, rand_ra, sizeof(double) * total, cudaMemcpyHostToDevice, st));
    FOFB_CUDA(cudaMemcpyAsync(a + total, rand_dec, sizeof(double) * total, cudaMemcpyHostToDevice, st));
    rra = a;
    rdec = a + total;
  }
  const int B = 256;
  const Catalog& cat = cat_();
  const int wblocks = grid_for((int64_t)K * 32, 128);
  int* N_own = merged.N.alloc(K);
  double* D = merged.D.alloc(K);
  FOFB_PROFILED(h, "k_own_count", k_own_count<<<wblocks, 128, 0, st>>>(K, merged.centre.p, merged.off.p, merged.rows.p, cat.ra_row.p, cat.dec_row.p, cos_row.p,
                                       c_ra.p, c_dec.p, c_cos.p, c_th05.p, p.grid_cells, N_own));
  const CellListView scl = smallp->view();
  OdCluster* oc = reinterpret_cast<OdCluster*>(sc.bbox.alloc((size_t)K * 3 + 4));  // 80 bytes per cluster
  unsigned long long* sums = sc.lkey.alloc((size_t)K * 2);
  k_od_prep<<<wblocks, 128, 0, st>>>(K, npts, merged.centre.p, cat.view(), c_wlo.p, c_whi.p, c_th05.p, rra, rdec, p.grid_cells,
                                     fmax(scl.s_ra, scl.s_dec), oc, sums);
  FOFB_LAUNCH_CHECK(h);
  FOFB_REQUIRE(total < (1ll << 31), "too many overdensity sample points");
  unsigned* qk = reinterpret_cast<unsigned*>(sc.ev_c.alloc((size_t)total));
  unsigned* qk2 = reinterpret_cast<unsigned*>(sc.ev_loser.alloc((size_t)total));
  PackedPoint* qv = sc.od_pay_a.alloc((size_t)total);
  PackedPoint* qv2 = sc.od_pay_b.alloc((size_t)total);
  k_od_keys<<<grid_for(total, 256), 256, 0, st>>>(total, scl, rra, rdec, qk, qv);
  FOFB_LAUNCH_CHECK(h);
  {
    int bits = 1;
    while ((1ll << bits) < (long long)scl.ncx * scl.ncy) ++bits;
    size_t bytes = 0;
    FOFB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, qk, qk2, qv, qv2, (int)total, 0, bits, st));
    FOFB_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp(h, bytes), bytes, qk, qk2, qv, qv2, (int)total, 0, bits, st));
    count_launch(h, 4);
  }
  {
    static bool attr_set[64] = {};
    if (!attr_set[h->device & 63]) {
      FOFB_CUDA(cudaFuncSetAttribute(k_od_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCtSmemBytes));
      attr_set[h->device & 63] = true;
    }
    const int nsx = (scl.ncx + kCtW - 1) / kCtW;
    const int nstrips = nsx * scl.ncy;
    int* qstart = sc.wlo.alloc((size_t)nstrips + 1);
    k_od_strip_starts<<<grid_for(nstrips + 1, 256), 256, 0, st>>>(nstrips, nsx, scl.ncx, total, qk2, qstart);
    FOFB_LAUNCH_CHECK(h);
    FOFB_PROFILED(h, "k_od_count", k_od_count<<<nstrips, kCtThreads, kCtSmemBytes, st>>>(scl, nsx, npts, qstart, qk2, qv2, oc, rra, rdec,
                                                                                          p.grid_cells, sums));
  }
  int* keep = sc.state.alloc(K);
  k_od_finish<<<grid_for(K, 128), 128, 0, st>>>(K, npts, sums, N_own, merged.off.p, p.overdensity, p.min_count, p.richness, D, keep);
  FOFB_LAUNCH_CHECK(h);
  sublist(h, merged, keep, final_, sc);
  if (final_.K > 0) {
    k_gather_nd<<<grid_for(final_.K, B), B, 0, st>>>(final_.K, sc.sel.p, N_own, D, final_.N.alloc(final_.K), final_.D.alloc(final_.K));
    FOFB_LAUNCH_CHECK(h);
  }
}

}  // namespace fofb
