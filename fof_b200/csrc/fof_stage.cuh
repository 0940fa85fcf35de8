// State of the FoF() pipeline (fof.py:98-392) kept in the handle between run / finish / fetch.
#pragma once
#include "catalog_host.cuh"

namespace fofb {

// A list of clusters: member rows of cluster k are rows[off[k] .. off[k+1]) in reference order.
struct ClusterList {
  int K = 0;
  int64_t total = 0;
  DBuf<int> centre;       // index into the candidate-centre matrix; >= m: galaxy row (centre - m) of a BCG-created cluster
  DBuf<long long> off;    // K + 1
  DBuf<int> rows;         // galaxy row indices
  DBuf<int> N;            // Cluster.N
  DBuf<double> D;         // Cluster.D
};

struct FofState {
  bool ran = false, finished = false;
  int n = 0, m = 0;
  View G, Cn;
  DBuf<double> g_store, c_store;  // fofb_fof_run: host matrices are staged here, not in the handle's shared buffers
  Catalog cat_own;
  CellList big;        // sized for the stage-1 virial search radius
  CellList small_own;  // sized for theta_0.5 (overdensity counts)
  // The whole-job pipeline has already sorted these very rows and built the theta_0.5 cell list for stage 0: when
  // no row was cut between stage 0 and FoF() it lends them (adopt) instead of having them built a second time.
  Catalog* catp = &cat_own;
  CellList* smallp = &small_own;
  void adopt(Catalog* c, CellList* s) {
    catp = c ? c : &cat_own;
    smallp = (c && s) ? s : &small_own;
  }
  Catalog& cat_() { return *catp; }
  const Catalog& cat_() const { return *catp; }
  fofb_params params{};

  // per-centre (index = priority)
  DBuf<double> c_ra, c_dec, c_cos, c_z, c_R1, c_T1, c_thvir, c_th05, c_mag, c_N, c_id;
  DBuf<int> c_wlo, c_whi;
  DBuf<unsigned char> alive, decided;
  DBuf<long long> t_off;     // per centre: its slice of t_rows
  DBuf<int> t_rows, t_count; // the galaxy rows whose membership switches a centre off (they carry its column-4 value)
  int duplicated_keys = 0;   // rows beyond the first that share a kill target (0 for unique column-4 values)
  DBuf<int> kill_target;   // per galaxy row: first centre whose column 4 equals the row's column 4, or -1
  DBuf<double> cos_row;

  // centre cell list (blocking test)
  DBuf<int> cc_start, cc_idx;
  DBuf<unsigned> cc_key_a, cc_key_b;
  DBuf<int> cc_val_a;
  double cc_ra0 = 0, cc_dec0 = 0, cc_sra = 1, cc_sdec = 1;
  int cc_ncx = 1, cc_ncy = 1;
  double R1max = 0;

  // round scratch
  DBuf<int> active_a, active_b, ready, flags, counts, nv, nF, nM, cursor, pass, resume_row, resume_cell, resume_k, seg_of_a, seg_of_b;
  DBuf<long long> voff, hoff, hcap, moff;
  DBuf<unsigned long long> keys1, keys1b, keys2, keys2b, htab;
  DBuf<int> v2row, mrows, nm_list, first_batch;
  DBuf<double> cc_z, v2ra, v2dec, v2cos;
  DBuf<int> owner, nm_sorted, hop2_qcount, hop2_qlist, hop2_nnm;
  DBuf<unsigned char> ismember;
  DBuf<double> LL;
  DBuf<double4> bbox1, bbox2, cc_a, cc_b;
  DBuf<long long> scal;    // device scalars
  DBuf<double> sort_keys_d, sort_keys_d2;
  DBuf<int> sort_vals, sort_vals2;

  // candidate clusters accumulated over rounds (unordered), then ordered lists per stage
  DBuf<int> pool_rows, pool_centre, pool_len;
  DBuf<long long> pool_off;
  int64_t pool_used = 0;
  int pool_K = 0;
  ClusterList candidates, merged, bcg, final_;
  int rounds = 0;
  int64_t evaluated = 0;

  void run(fofb_handle* h, const fofb_params& p, const View& G, const View& Cn);
  // the same in steps (the distributed driver merges alive/decided across processes between evaluate and advance)
  void prepare(fofb_handle* h, const fofb_params& p, const View& G, const View& Cn, const unsigned char* owned);
  void restart_rounds(fofb_handle* h, double linking_length_factor, int richness, double overdensity);
  void round_evaluate(fofb_handle* h);
  void round_advance(fofb_handle* h);
  void finalize(fofb_handle* h);
  void* overlap_scratch = nullptr;          // OverlapScratch (overlap.cu)
  void (*overlap_scratch_free)(void*) = nullptr;
  ~FofState() {
    if (overlap_scratch && overlap_scratch_free) overlap_scratch_free(overlap_scratch);
    if (rng_done) cudaEventDestroy(rng_done);
    if (rng_stream) cudaStreamDestroy(rng_stream);
    for (int q = 0; q < 2; ++q) {
      if (hop2_done[q]) cudaEventDestroy(hop2_done[q]);
      if (hop2_side[q]) cudaStreamDestroy(hop2_side[q]);
    }
    if (hop2_go) cudaEventDestroy(hop2_go);
  }
  bool distributed = false, act_in_a = true;
  int na = 0;
  void finish(fofb_handle* h, const double* rand_ra, const double* rand_dec, int mem);
  // u: K x 2 x n_random samples of [0, 1) (per cluster the RA block, then the Dec block)
  void finish_unit(fofb_handle* h, const double* u, int mem);
  DBuf<double> unit_pts;
  // draw the points on the device from numpy's legacy MT19937 state (624-word key + position), advanced in place
  void finish_mt(fofb_handle* h, unsigned* key_host, int* pos_host);
  DBuf<unsigned> mt_key, mt_words;
  void mt_prefetch(fofb_handle* h, const unsigned* key_host, int pos, long long k_spec, int part = 0, int parts = 1);
  int mt_part = 0, mt_parts = 1;  // this process generated chunk share `part` of `parts` (slab runs all-gather the shares)
  cudaStream_t rng_stream = nullptr;
  cudaStream_t hop2_side[2] = {nullptr, nullptr};   // the larger hop-2 queues run beside the first (fof_stage.cu)
  cudaEvent_t hop2_done[2] = {nullptr, nullptr}, hop2_go = nullptr;
  cudaEvent_t rng_done = nullptr;
  bool mt_prefetched = false;
  int mt_pos0 = 0, params_n_random_hint = 300;
  long long mt_head = 0, mt_blocks = 0;
  unsigned mt_key_host[624];
  int last_K = 0;
  // distributed runs: index of each local merged cluster in the all-process list, and that list's length
  DBuf<int> merged_gidx;
  long long K_glob_merged = -1;
  bool have_global_bbox = false;
  BBox global_bbox{};
};

}  // namespace fofb
