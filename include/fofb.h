/* fofb.h -- C ABI of the B200-native FoF cluster-finding hot path (libfofb200.so).
 *
 * The reference (kencx/FoF) is pure Python and has no FFI: its boundary for this path is the set
 * of Python call signatures used by /root/reference/FoF/pipeline.py.  Each entry point below names
 * the reference interface it stands behind; the Python facade in fof_b200/ keeps those signatures
 * and calls these functions through ctypes (see INTEGRATION.md for the binding).
 *
 * Conventions
 *   - every function returns 0 on success, a negative fofb_status on failure; nothing throws across
 *     the ABI; fofb_last_error(h) returns the message of the last failure on that handle;
 *   - plain pointers and sizes only; buffers are caller-owned and borrowed for the call;
 *   - a fofb_matrix describes a strided float64 matrix in HOST (mem = 0) or DEVICE (mem = 1) memory;
 *     the whole underlying block [data, data + extent) must be readable;
 *   - variable-size results stay in the handle after a *_run call; the caller reads the sizes,
 *     allocates, and copies them out with the matching *_fetch call;
 *   - a handle is bound to one CUDA device and is not thread-safe; calls are synchronous.
 * Column layout of galaxy / centre rows (pipeline.py:51-53 + fof.py:61):
 *   0 ra[deg] 1 dec[deg] 2 z 3 z_lower 4 z_upper 5 RMag 6 ID 7 N(0.5)
 */
#ifndef FOFB_H
#define FOFB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FOFB_ABI_VERSION 1

typedef enum fofb_status {
  FOFB_OK = 0,
  FOFB_ERR_INVALID = -1,     /* bad argument / shape / struct_size                      */
  FOFB_ERR_CUDA = -2,        /* CUDA runtime error (message holds cudaGetErrorString)    */
  FOFB_ERR_NO_DEVICE = -3,   /* no usable CUDA device: there is NO CPU fallback          */
  FOFB_ERR_STATE = -4,       /* call order violated (e.g. fetch before run)              */
  FOFB_ERR_UNSUPPORTED = -5, /* degenerate input the device path refuses (documented)    */
  FOFB_ERR_INTERNAL = -6
} fofb_status;

typedef struct fofb_handle fofb_handle;

/* params.py:16-27 plus the magic numbers inlined in fof.py / helper.py.  */
typedef struct fofb_params {
  size_t struct_size;            /* = sizeof(fofb_params); versioning                           */
  double H0, Om0, Ode0;          /* params.py:16   LambdaCDM(70, 0.3, 0.7), Tcmb0 = 0           */
  double max_velocity;           /* params.py:23   km/s                                          */
  double linking_length_factor;  /* params.py:24   b (FoF() default 0.1, fof.py:104)             */
  double virial_radius;          /* params.py:25   Mpc/h                                         */
  double overdensity;            /* params.py:27   D                                             */
  double count_radius;           /* fof.py:73,378; helper.py:212   0.5 Mpc/h                     */
  double bcg_fraction;           /* fof.py:334     0.25                                          */
  double survey_area;            /* fof.py:192     1.7 (used as a radius in degrees)             */
  int32_t richness;              /* params.py:26                                                 */
  int32_t min_window;            /* fof.py:71      8                                             */
  int32_t min_count;             /* fof.py:77,386  8                                             */
  int32_t min_virial;            /* fof.py:185     12                                            */
  int32_t n_random;              /* helper.py:196  300                                           */
  int32_t n_bins;                /* fof.py:506     10                                            */
  int32_t grid_cells;            /* GriSPy N_cells 64                                            */
  int32_t reserved;
} fofb_params;

typedef struct fofb_matrix {
  const double* data;  /* element (i, c) at data[i * row_stride + c * col_stride]               */
  int64_t rows;
  int32_t cols;
  int32_t mem;         /* 0 = host, 1 = device                                                  */
  int64_t row_stride;  /* in elements                                                           */
  int64_t col_stride;  /* in elements                                                           */
} fofb_matrix;

/* ---- lifetime ---------------------------------------------------------------------------- */
int fofb_abi_version(void);
const char* fofb_build_info(void);
int fofb_params_default(fofb_params* p);             /* params.py defaults                       */
int fofb_create(int device, fofb_handle** out);      /* fails with FOFB_ERR_NO_DEVICE w/o a GPU  */
void fofb_destroy(fofb_handle* h);
const char* fofb_last_error(const fofb_handle* h);
int fofb_synchronize(fofb_handle* h);
/* Run the handle's work on the caller's CUDA stream (a cudaStream_t, e.g. torch.cuda.current_stream().cuda_stream)
 * instead of the handle's own: kernels of the caller and NCCL collectives enqueued on that stream are then ordered
 * with the library's without host synchronisation.  NULL returns to the handle's own stream.  The stream is borrowed. */
int fofb_set_stream(fofb_handle* h, void* cuda_stream);
/* CUDA-event time [ms] of the kernels launched by the last compute call, and their count.    */
int fofb_last_timing(const fofb_handle* h, double* device_ms, int64_t* kernel_launches);

/* Per-kernel CUDA-event timing of the library's own kernels (recorded on the launching stream).
 * on = 1 enable, 2 enable and reset, 0 disable.  fofb_profile_read writes "name total_ms launches" lines. */
int fofb_profile_enable(fofb_handle* h, int32_t on);
int fofb_profile_read(fofb_handle* h, char* buf, int64_t capacity);

/* ---- host scalar helpers: analysis/methods.py (no GPU needed) ----------------------------------
 * fofb_linear_to_angular_dist : analysis/methods.py:12-32   h^-1 Mpc (proper) at z -> degrees
 * fofb_mean_separation        : analysis/methods.py:35-65   Mpc
 * fofb_redshift_to_velocity   : analysis/methods.py:68-87   km/s
 * fofb_comoving_distance / fofb_angular_diameter_distance : astropy LambdaCDM (Mpc)
 */
int fofb_comoving_distance(const fofb_params* p, int64_t n, const double* z, double* out);
int fofb_angular_diameter_distance(const fofb_params* p, int64_t n, const double* z, double* out);
int fofb_linear_to_angular_dist(const fofb_params* p, double distance_mpc_h, int64_t n, const double* z, double* out);
int fofb_mean_separation(const fofb_params* p, double n_objects, double z, double* out);
int fofb_redshift_to_velocity(int64_t n, const double* z, double center_z, double* out);

/* ---- stage 0: candidate_center_search (fof.py:24-95) ---------------------------------------------
 * N(0.5) for every row of `galaxies` (>= 3 columns: ra, dec, z), written to N_out[rows] (host or
 * device pointer per N_mem) in input row order; the number of rows with N >= min_count is returned
 * in *n_candidates.  fofb_candidate_order then writes the row indices of the candidate centres
 * sorted by N descending, ties by descending row index (oracle pin D1 for fof.py:79-81).
 */
int fofb_number_counts(fofb_handle* h, const fofb_params* p, const fofb_matrix* galaxies,
                       int32_t* N_out, int32_t N_mem, int64_t* n_candidates);
int fofb_candidate_order(fofb_handle* h, int32_t* rows_out, int32_t rows_mem, int64_t capacity);

/* ---- stages 1-4: FoF() (fof.py:98-392) ---------------------------------------------------------------
 * fofb_fof_run    : stage 1 (per-centre virial bubble, hop 1 / hop 2, richness gate, greedy tracker,
 *                   fof.py:145-246), stage 2 (ordered overlap contests, fof.py:258-314; cluster.py:65-82)
 *                   and stage 3 (BCG check, fof.py:318-371).  `galaxies` (n x C) and `centres` (m x C),
 *                   C >= 8, in the column layout above; row i of `centres` has priority i.  On return
 *                   *n_clusters is the number of clusters that reach stage 4 and bbox[4] =
 *                   (ra_min, ra_max, dec_min, dec_max) of `galaxies` (the bounds helper.py:197-202 draws
 *                   its uniform points from).
 * fofb_fof_finish : stage 4 (fof.py:374-392): N := own members within theta_0.5, D from n_random points
 *                   per cluster.  rand_ra / rand_dec hold n_clusters * n_random values, cluster-major, drawn
 *                   by the CALLER from numpy's legacy global RNG exactly as helper.py:197-202 does (RA block
 *                   then Dec block per cluster), so a seeded reference run and this path see the same points.
 * fofb_fof_sizes / fofb_fof_fetch : read a cluster list kept in the handle.  which = 0 candidates after
 *                   stage 1, 1 after overlap removal, 2 after the BCG stage, 3 final.  offsets[K+1] (int64),
 *                   members[total] (int32 galaxy row indices in the reference's member order), centre[K]
 *                   (int32 row of `centres`), N[K] (float64 Cluster.N), D[K] (float64 Cluster.D).
 *                   Any output pointer may be NULL.
 */
int fofb_fof_run(fofb_handle* h, const fofb_params* p, const fofb_matrix* galaxies, const fofb_matrix* centres,
                 int64_t* n_clusters, double* bbox);
int fofb_fof_finish(fofb_handle* h, const double* rand_ra, const double* rand_dec, int32_t mem, int64_t* n_final);
/* Same as fofb_fof_finish, but the points are drawn ON THE DEVICE from numpy's legacy MT19937 generator: mt_key[624]
 * and *mt_pos are the 'key' and 'pos' entries of np.random.get_state(); they are advanced in place by exactly the
 * 4 * n_clusters * n_random words the reference's np.random.uniform calls would consume (helper.py:197-202), so
 * np.random.set_state() with the returned values leaves the caller's generator where the reference would. */
int fofb_fof_finish_mt(fofb_handle* h, uint32_t* mt_key, int32_t* mt_pos, int64_t* n_final);
int fofb_fof_sizes(fofb_handle* h, int32_t which, int64_t* n_clusters, int64_t* n_members);
int fofb_fof_fetch(fofb_handle* h, int32_t which, int64_t* offsets, int32_t* members, int32_t* centre, double* N,
                   double* D);
/* diagnostics of the last fofb_fof_run: priority rounds and centres actually evaluated */
int fofb_fof_stats(fofb_handle* h, int64_t* rounds, int64_t* evaluated);

/* ---- stage 5: three_sigma / interloper_removal (fof.py:396-516) -----------------------------------------
 * K clusters; cluster k owns rows offsets[k] .. offsets[k+1] of `members` (>= 3 columns: ra, dec, z) and has
 * centre (ra, dec, z) = centres[3k .. 3k+2] (host).  For each cluster the members are binned in n_bins radial
 * bins and clipped iteratively at 3 sigma in velocity.  out_offsets[K+1] and out_index[<= total] (host)
 * receive, per cluster, the positions (0-based inside the cluster) of the surviving members in the
 * reference's output order (bins concatenated, each bin sorted by its last deviation sort).
 * The richness filter of fof.py:514 is the caller's one-liner.
 *
 * ---- stage 6: virial_mass_estimator + find_masses (mass_estimator.py:36-112,148-184; mass_analysis.py:8-44) ---
 * out[8k .. 8k+7] (host) = cluster_mass [Msun/h], mass_err, vel_disp [km^2/s^2], vel_disp_err,
 * projected_radius [Mpc/h], total_absmag, mean redshift, member count.  `members` needs the first six columns.
 *
 * fofb_bootstrap_indices (host): the first `count` values of np.random.randint(0, n) after np.random.seed(1),
 * i.e. the index stream astropy.stats.bootstrap consumes under NumpyRNGContext(1).
 */
int fofb_three_sigma(fofb_handle* h, const fofb_params* p, int64_t K, const int64_t* offsets, const fofb_matrix* members,
                     const double* centres, int64_t* out_offsets, int32_t* out_index);
int fofb_virial(fofb_handle* h, const fofb_params* p, int64_t K, const int64_t* offsets, const fofb_matrix* members,
                double* out);
/* projected_mass_estimator (mass_estimator.py:115-145): out[2k] = mass [Msun/h], out[2k+1] = mass_err [km^2/s^2] */
int fofb_projected_mass(fofb_handle* h, const fofb_params* p, int64_t K, const int64_t* offsets, const fofb_matrix* members,
                        double* out);
int fofb_bootstrap_indices(int64_t n, int64_t count, int32_t* out);

/* The device generator behind fofb_*_finish_mt on its own: the next n_words 32-bit outputs of numpy's legacy MT19937
 * (np.random.get_state(): key624, pos) into out (host, or device when out_mem = 1); key624 / pos are advanced exactly as
 * np.random would be.  The stream is produced in parallel chunks reached by polynomial jump-ahead. */
int fofb_mt19937_words(fofb_handle* h, uint32_t* key624, int32_t* pos, int64_t n_words, uint32_t* out, int32_t out_mem);

/* ---- whole job: pipeline.py:62-64,80-99,113,173 with every intermediate on the device ---------------------
 * fofb_pipeline_begin : candidate_center_search on `galaxies` (n x 7), the z cuts of pipeline.py:80-86
 *                       (galaxies zmin <= z <= zmax_gal, centres zmin <= z <= zmax_cen), FoF() stages 1-3.
 *                       *n_clusters = clusters reaching the overdensity stage.
 * fofb_pipeline_finish: `u` holds n_clusters * 2 * n_random samples of [0, 1) in numpy stream order (per
 *                       cluster the RA block then the Dec block, i.e. np.random.random_sample((K, 2, n)));
 *                       the points are low + (high - low) * u as in helper.py:197-202.  Then stage 4,
 *                       interloper_removal (richness filter of fof.py:514 included) and find_masses("virial").
 * fofb_pipeline_fetch : offsets[K+1], member_rows[total] and centre_row[K] index rows of the INPUT matrix;
 *                       N[K], D[K], props[K x 8] as in fofb_virial.  NULL pointers are skipped.
 */
int fofb_pipeline_begin(fofb_handle* h, const fofb_params* p, const fofb_matrix* galaxies, double zmin, double zmax_gal,
                        double zmax_cen, int64_t* n_clusters);
int fofb_pipeline_finish(fofb_handle* h, const double* u, int32_t mem, int64_t* n_final, int64_t* n_members);
int fofb_pipeline_finish_mt(fofb_handle* h, uint32_t* mt_key, int32_t* mt_pos, int64_t* n_final, int64_t* n_members);
/* begin + finish_mt in one call; the MT19937 stream is generated on a side stream while stages 0-3 run */
int fofb_pipeline_run(fofb_handle* h, const fofb_params* p, const fofb_matrix* galaxies, double zmin, double zmax_gal,
                      double zmax_cen, uint32_t* mt_key, int32_t* mt_pos, int64_t* n_final, int64_t* n_members);
/* parameter sweep on the catalogue of the last fofb_pipeline_run / begin: stage 0 and everything that depends only on
 * max_velocity, the radii and the cosmology is reused; the rounds, overlap removal, selection, interloper removal and
 * masses run again with the new linking_length_factor / richness / overdensity (the thesis tuned exactly these). */
int fofb_pipeline_rerun(fofb_handle* h, double linking_length_factor, int32_t richness, double overdensity, uint32_t* mt_key,
                        int32_t* mt_pos, int64_t* n_final, int64_t* n_members);
int fofb_pipeline_fetch(fofb_handle* h, int64_t* offsets, int32_t* member_rows, int32_t* centre_row, double* N, double* D,
                        double* props);
/* N(0.5) of every INPUT row of the last pipeline run (column 7 of main_arr, fof.py:61-74): what the reference's
 * Cluster.galaxies rows carry in their last column.  N_row[n_input_rows], host (mem 0) or device (mem 1). */
int fofb_pipeline_fetch_counts(fofb_handle* h, int32_t* N_row, int32_t mem);
/* sizes seen by the last pipeline run: galaxies after the z cut, candidate centres, stage-1 candidate clusters,
 * total stage-1 members, priority rounds, centres evaluated, clusters entering stage 4 */
int fofb_pipeline_stats(fofb_handle* h, int64_t* stats7);

/* ---- one catalogue over several GPUs: redshift slabs with window-wide halos --------------
 * One process per GPU holds the rows of a redshift slab plus halos; these calls are the device-side steps, the
 * driver (fof_b200/distributed.py) runs the collectives between them:
 *   fofb_dist_begin          stage 0 on the local rows; zcut3 = (zmin, zmax_gal, zmax_cen) of pipeline.py:80-86,
 *                            z_valid2 = redshift range whose velocity windows are complete on this process,
 *                            z_own2 = [lo, hi) of the centres this process evaluates
 *   fofb_dist_local_inputs   local candidate centres (8 columns, local priority order), their source rows, the
 *                            source row of every FoF() row, and the bounding box of the local FoF() rows
 *   [all-gather centres; global priority order = N desc, global row desc]
 *   fofb_dist_prepare        centres[n x 8] in global priority order, owned[n] flags; returns the bounding box of the
 *                            local FoF() rows [all-reduce min/max -> the box helper.py:197-202 draws points from]
 *   loop: fofb_dist_round_evaluate; fofb_dist_state(0); [all-reduce MIN alive, MAX decided]; fofb_dist_state(1);
 *         fofb_dist_round_advance  -> until the global number of active centres is 0
 *   fofb_dist_finalize, fofb_dist_local_clusters   [all-gather; order by centre index]
 *   fofb_dist_overlap        stage 2 on the all-process list (replicated)
 *   fofb_dist_set_merged     then fofb_pipeline_finish_mt / fofb_pipeline_fetch as in the single-GPU path; the
 *                            MT19937 stream covers every process's clusters, each process uses its own blocks;
 *                            centre_row of fofb_pipeline_fetch is then an index into the global centre table.
 */
int fofb_dist_begin(fofb_handle* h, const fofb_params* p, const fofb_matrix* galaxies, const double* zcut3,
                    const double* z_valid2, const double* z_own2, int64_t* n_fof_rows, int64_t* n_local_centres);
int fofb_dist_prefetch_mt(fofb_handle* h, const uint32_t* mt_key, int32_t mt_pos, int64_t k_spec);
/* the same for share `part` of `parts` of the stream (whole chunks of it): every process generates its share, the caller
 * all-gathers the shares in place -- words_ptr[blocks_offset + r * share_words ..) is share r (share_words = 0: the stream
 * was too long to share and has been generated in full) -- after fofb_dist_mt_wait has put the generator's side stream
 * in front of the handle's stream */
int fofb_dist_prefetch_mt_part(fofb_handle* h, const uint32_t* mt_key, int32_t mt_pos, int64_t k_spec, int32_t part, int32_t parts,
                               uint64_t* words_ptr, int64_t* blocks_offset, int64_t* share_words);
int fofb_dist_mt_wait(fofb_handle* h);
int fofb_dist_local_inputs(fofb_handle* h, double* centre_rows, int32_t* centre_src_row, int32_t* fof_src_row, double* bbox4);
int fofb_dist_prepare(fofb_handle* h, const double* centres, int64_t n_centres, const uint8_t* owned, double* local_bbox4);
int fofb_dist_round_evaluate(fofb_handle* h);
int fofb_dist_state(fofb_handle* h, uint8_t* alive, uint8_t* decided, int32_t direction);
int fofb_dist_round_advance(fofb_handle* h, int64_t* n_active);
/* Stage 2 distributed.  A contest needs its two clusters within one velocity window of each other, so a process can list
 * the contests of its own clusters from its own clusters plus the neighbours' boundary clusters:
 *   fofb_dist_overlap_events       : cluster list (centre rows K x 8 in global priority order, CSR offsets, member ids;
 *                                    device arrays) -> number of contests, kept in the handle
 *   fofb_dist_overlap_events_fetch : ev_a / ev_c / ev_loser [n_events] (device): cluster a, its partner c, the static
 *                                    loser (-1: the contest cannot fire), ordered by a, then in GriSPy's return order
 *   [all-gather the contests of the owned clusters; order them by a's global priority; indices -> all-process list]
 *   fofb_dist_overlap_resolve      : the ordered relaxation (fof.py:294-310 with the in-place mutation of Q6) on the
 *                                    all-process contest list, replicated; keep[K] (device, non-zero = survives) */
int fofb_dist_overlap_events(fofb_handle* h, int64_t K, const double* centre_rows, const int64_t* offsets, int64_t total,
                             const double* ids, int64_t* n_events);
int fofb_dist_overlap_events_fetch(fofb_handle* h, int64_t n_events, int32_t* ev_a, int32_t* ev_c, int32_t* ev_loser);
int fofb_dist_overlap_resolve(fofb_handle* h, int64_t K, int64_t n_events, const int32_t* ev_a, const int32_t* ev_c,
                              const int32_t* ev_loser, int32_t* keep);
/* The same loop with ONE collective per round.  The tracker state of a centre is 0 (alive, undecided), 1 (switched off)
 * or 2 (evaluated); it only ever rises and 2 wins, so an element-wise MAX over processes merges it.
 *   fofb_dist_state_pack   : buf[0 .. n_buf) := 0, buf[shared_pos[t]] := state of local centre local_pos[t]
 *                            (t < n_shared), buf[n_buf - 1] := 1 if this process entered the round with active centres
 *   [all-reduce MAX over buf]
 *   fofb_dist_state_unpack : the merged states back into the local table
 *   fofb_dist_round_advance2: as fofb_dist_round_advance; *any_active = buf[n_buf - 1] (somebody still had work)
 * All pointers are device pointers; local_pos / shared_pos are int64. */
int fofb_dist_state_pack(fofb_handle* h, int64_t n_shared, const int64_t* local_pos, const int64_t* shared_pos, uint8_t* buf,
                         int64_t n_buf);
int fofb_dist_state_unpack(fofb_handle* h, int64_t n_shared, const int64_t* local_pos, const int64_t* shared_pos,
                           const uint8_t* buf);
int fofb_dist_round_advance2(fofb_handle* h, const uint8_t* buf, int64_t n_buf, int64_t* n_active, int32_t* any_active);
int fofb_dist_finalize(fofb_handle* h, int64_t* n_clusters, int64_t* n_members);
int fofb_dist_local_clusters(fofb_handle* h, int32_t* centre, int64_t* offsets, double* ids);
int fofb_dist_overlap(fofb_handle* h, int64_t K, const int32_t* centre, const int64_t* offsets, const double* ids, int32_t* keep);
/* device-array variants: clusters' centres as rows instead of indices into a replicated centre table */
int fofb_dist_local_clusters_dev(fofb_handle* h, int32_t* centre, int64_t* sizes, double* ids);
int fofb_dist_overlap_rows(fofb_handle* h, int64_t K, const double* centre_rows, const int64_t* offsets, int64_t total,
                           const double* ids, int32_t* keep);
int fofb_dist_set_merged(fofb_handle* h, const int32_t* keep_local, const int32_t* global_index, int64_t n_global_merged,
                         const double* global_bbox4);

/* ---- helper-level bubble counts: GriSPy(points).bubble_neighbors(centres, r) as counts --------------------------
 * find_number_count (helper.py:149-188) and the 300-point counts of center_overdensity (helper.py:207-216).  The grid
 * is built on the rows of `points` (ra, dec[, z]) that lie in the velocity window of `window_z` when use_window != 0
 * (p->max_velocity), else on all rows.  centres[2q], radius_deg[q], counts[q] are host arrays. */
int fofb_bubble_count(fofb_handle* h, const fofb_params* p, const fofb_matrix* points, int64_t n_centres, const double* centres,
                      const double* radius_deg, int32_t use_window, double window_z, int32_t* counts);

/* ---- catalogue matching: compare_clusters (catalog_matching.py:21-111) -------------------------------------------
 * sample: (S, >=3) rows [ra, dec, z] of found clusters; catalogue: (C, >=3) rows [ra, dec, z] of a published survey.
 * Per catalogue row i: match_row[i] = -2 when its z lies outside [min z_s - 0.1, max z_s + 0.1] (:66-68, such rows are
 * dropped by the reference before indexing); otherwise the nearest sample row on the sky among those with
 * |z_s - z_i| <= z_cut (:78-80, :89-91; ties: smallest squared chord, then smallest row -- pinned order D5) when its
 * Vincenty separation is within the angle of :85-88 for `matching_distance_mpc_h`, else -1.  match_sep_deg[i] = that
 * separation (nan when no sample row passes the cut).  Outputs are host (out_mem = 0) or device (1) arrays of C. */
int fofb_match_catalog(fofb_handle* h, const fofb_params* p, const fofb_matrix* sample, const fofb_matrix* catalogue,
                       double matching_distance_mpc_h, double z_cut, int32_t* match_row, double* match_sep_deg, int32_t out_mem,
                       int64_t* n_matched);

/* ---- transitive friends-of-friends (connected components) ------------------------------------------------------
 * NOT a reference code path (kencx/FoF never computes connected components): the lock-free
 * union-find mode named in BASELINE.json.  Galaxies i, j are friends iff each lies in the other's velocity window
 * (analysis/methods.py:68-87 with p->max_velocity) and their haversine separation is within the angle of the
 * transverse linking length [h^-1 Mpc proper] at both redshifts (analysis/methods.py:12-32).  labels[i] = smallest
 * row index of the component of row i.  n_links (optional): number of friend pairs; passing NULL lets the kernel skip
 * pairs whose ends are already connected (much faster on percolating blobs).  Oracle: oracle/transitive_oracle.py
 * (scipy connected components).
 */
int fofb_transitive_labels(fofb_handle* h, const fofb_params* p, const fofb_matrix* galaxies, double linking_length_mpc_h,
                           int32_t* labels, int32_t labels_mem, int64_t* n_groups, int64_t* n_links);

#ifdef __cplusplus
}
#endif
#endif /* FOFB_H */
