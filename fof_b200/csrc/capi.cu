// extern "C" surface of libfofb200 (include/fofb.h).  Exceptions stop here.
#include <cstring>
#include <new>

#include "catalog_host.cuh"
#include "stage0.cuh"
#include "fof_stage.cuh"
#include "clean.cuh"
#include "pipeline.cuh"
#include "transitive.cuh"
#include "bubble.cuh"
#include "match.cuh"

namespace fofb {
void fof_overlap_external(fofb_handle* h, FofState& S, int K, const int* centre, const long long* off, long long total,
                          const double* ids, int* keep_out);
void fof_overlap_external_rows(fofb_handle* h, FofState& S, int K, const double* centre_rows, const long long* off, long long total,
                               const double* ids, int* keep_out);
void fof_set_merged(fofb_handle* h, FofState& S, const int* keep_local);
long long fof_overlap_events_rows(fofb_handle* h, FofState& S, int K, const double* centre_rows, const long long* off, long long total,
                                  const double* ids);
void fof_overlap_events_fetch(fofb_handle* h, FofState& S, long long E, int* ev_a, int* ev_c, int* ev_loser);
void fof_overlap_resolve_external(fofb_handle* h, FofState& S, int K, long long E, const int* ev_a, const int* ev_c, const int* ev_loser,
                                  int* keep_out);
void mt19937_words(fofb_handle* h, unsigned* key_host, int* pos_host, long long n_words, unsigned* out, int out_mem);
void dist_state_pack(fofb_handle* h, FofState& S, long long n_shared, const long long* lpos, const long long* bpos, unsigned char* buf,
                     long long n_buf);
void dist_state_unpack(fofb_handle* h, FofState& S, long long n_shared, const long long* lpos, const long long* bpos,
                       const unsigned char* buf);
}

namespace fofb {

struct Stage0Holder {
  Catalog cat;
  CellList cells;
  Stage0 s0;
};

Cosmo host_cosmo_with(const fofb_params& p, std::vector<double>& table) {
  table.resize(kCosmoTableSize);
  cosmo_build_table(p.Om0, p.Ode0, table.data());
  Cosmo c{};
  c.H0 = p.H0; c.Om0 = p.Om0; c.Ode0 = p.Ode0; c.Ok0 = 1.0 - p.Om0 - p.Ode0;
  c.h = p.H0 / 100.0; c.DH = kCkms / p.H0;
  c.zmax = (double)(kCosmoTableSize - 1) / kCosmoPanelsPerUnit;
  c.table = table.data();
  return c;
}

Cosmo device_cosmo(fofb_handle* h, const fofb_params& p) {
  if (h->cosmo_key[0] != p.Om0 || h->cosmo_key[1] != p.Ode0 || !h->cosmo_table.p) {
    std::vector<double> t;
    host_cosmo_with(p, t);
    double* d = h->cosmo_table.alloc(kCosmoTableSize);
    FOFB_CUDA(cudaMemcpyAsync(d, t.data(), sizeof(double) * kCosmoTableSize, cudaMemcpyHostToDevice, h->stream));
    FOFB_CUDA(cudaStreamSynchronize(h->stream));
    h->cosmo_key[0] = p.Om0;
    h->cosmo_key[1] = p.Ode0;
  }
  std::vector<double> unused;
  Cosmo c{};
  c.H0 = p.H0; c.Om0 = p.Om0; c.Ode0 = p.Ode0; c.Ok0 = 1.0 - p.Om0 - p.Ode0;
  c.h = p.H0 / 100.0; c.DH = kCkms / p.H0;
  c.zmax = (double)(kCosmoTableSize - 1) / kCosmoPanelsPerUnit;
  c.table = h->cosmo_table.p;
  return c;
}

View to_device(fofb_handle* h, const fofb_matrix& m, DBuf<double>& staging) {
  FOFB_REQUIRE(m.data != nullptr && m.rows >= 0 && m.cols > 0, "matrix: null data or bad shape");
  FOFB_REQUIRE(m.row_stride >= 0 && m.col_stride >= 0, "matrix: negative strides are not supported");
  if (h->upload_pending) {  // a split upload that an error left unfinished: its side-stream copies target the staging buffers
    FOFB_CUDA(cudaStreamWaitEvent(h->stream, h->upload_done, 0));
    h->upload_pending = false;
  }
  View v;
  v.rows = m.rows; v.cols = m.cols; v.rs = m.row_stride; v.cs = m.col_stride;
  if (m.mem == 1) {
    v.data = m.data;
    return v;
  }
  FOFB_REQUIRE(m.mem == 0, "matrix: mem must be 0 (host) or 1 (device)");
  const size_t extent = m.rows > 0 ? (size_t)((m.rows - 1) * m.row_stride + (m.cols - 1) * m.col_stride + 1) : 0;
  double* d = staging.alloc(extent ? extent : 1);
  if (extent) FOFB_CUDA(cudaMemcpyAsync(d, m.data, extent * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  v.data = d;
  return v;
}

// Column-major host tables (a pandas block: every column contiguous) for the whole-job path: the three columns stage 0
// reads go first on the compute stream; the others follow on a side stream while stage 0 runs.  The caller makes the
// compute stream wait for h->upload_done before anything touches columns >= first_cols (Pipeline::stage0_and_inputs).
View to_device_split(fofb_handle* h, const fofb_matrix& m, DBuf<double>& staging, int first_cols) {
  const bool column_major = m.mem == 0 && m.row_stride == 1 && m.col_stride >= m.rows && m.cols > first_cols && m.rows > 0;
  h->upload_pending = false;
  if (!column_major) return to_device(h, m, staging);
  FOFB_REQUIRE(m.data != nullptr, "matrix: null data");
  if (!h->copy_stream) {
    FOFB_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    FOFB_CUDA(cudaEventCreateWithFlags(&h->upload_done, cudaEventDisableTiming));
    FOFB_CUDA(cudaEventCreateWithFlags(&h->upload_go, cudaEventDisableTiming));
  }
  const size_t n = (size_t)m.rows;
  double* d = staging.alloc(n * m.cols);
  // the staging buffer may still be read by work enqueued earlier on the compute stream
  FOFB_CUDA(cudaEventRecord(h->upload_go, h->stream));
  FOFB_CUDA(cudaStreamWaitEvent(h->copy_stream, h->upload_go, 0));
  for (int c = 0; c < m.cols; ++c)
    FOFB_CUDA(cudaMemcpyAsync(d + (size_t)c * n, m.data + (size_t)c * m.col_stride, n * sizeof(double), cudaMemcpyHostToDevice,
                              c < first_cols ? h->stream : h->copy_stream));
  FOFB_CUDA(cudaEventRecord(h->upload_done, h->copy_stream));
  h->upload_pending = true;
  View v;
  v.data = d; v.rows = m.rows; v.cols = m.cols; v.rs = 1; v.cs = (int64_t)n;
  return v;
}

static void check_params(const fofb_params* p) {
  FOFB_REQUIRE(p != nullptr, "params is null");
  FOFB_REQUIRE(p->struct_size == sizeof(fofb_params), "params.struct_size does not match this library (ABI mismatch)");
  FOFB_REQUIRE(p->H0 > 0 && p->Om0 > 0 && p->max_velocity >= 0, "params: H0, Om0 must be positive, max_velocity >= 0");
  FOFB_REQUIRE(p->grid_cells >= 1 && p->grid_cells <= 32768, "params: grid_cells out of range");
  FOFB_REQUIRE(p->count_radius > 0 && p->virial_radius > 0, "params: radii must be positive");
}

struct Timed {  // CUDA-event bracket around one compute call
  fofb_handle* h;
  explicit Timed(fofb_handle* h_) : h(h_) {
    h->launches = 0;
    cudaEventRecord(h->ev0, h->stream);
  }
  void stop() {
    cudaEventRecord(h->ev1, h->stream);
    cudaEventSynchronize(h->ev1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    h->last_ms = ms;
    h->prof.resolve();
  }
};

}  // namespace fofb

using namespace fofb;

// An entry point that fails returns while copies it enqueued may still be reading the caller's host buffers (pinned
// memory: cudaMemcpyAsync is truly asynchronous) -- drain the handle's streams before handing the error back, so that
// the caller may free or reuse its arrays as soon as any call has returned, failed or not.
static void drain_after_error(fofb_handle* h) {
  if (!h) return;
  cudaStreamSynchronize(h->stream);
  if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
}

#define FOFB_TRY(h) try {
#define FOFB_CATCH(h)                                              \
  }                                                                \
  catch (const fofb::Error& e) {                                   \
    if (h) (h)->last_error = e.what();                             \
    drain_after_error(h);                                          \
    return e.code;                                                 \
  }                                                                \
  catch (const std::bad_alloc&) {                                  \
    if (h) (h)->last_error = "host allocation failed";             \
    drain_after_error(h);                                          \
    return FOFB_ERR_INTERNAL;                                      \
  }                                                                \
  catch (const std::exception& e) {                                \
    if (h) (h)->last_error = std::string("internal: ") + e.what(); \
    drain_after_error(h);                                          \
    return FOFB_ERR_INTERNAL;                                      \
  }

static thread_local std::string g_create_error;

extern "C" {

int fofb_abi_version(void) { return FOFB_ABI_VERSION; }

const char* fofb_build_info(void) { return "libfofb200 sm_100a " __DATE__ " " __TIME__; }

int fofb_params_default(fofb_params* p) {
  if (!p) return FOFB_ERR_INVALID;
  std::memset(p, 0, sizeof(*p));
  p->struct_size = sizeof(fofb_params);
  p->H0 = 70.0; p->Om0 = 0.3; p->Ode0 = 0.7;
  p->max_velocity = 2000.0;
  p->linking_length_factor = 0.2;
  p->virial_radius = 1.5;
  p->overdensity = 2.0;
  p->count_radius = 0.5;
  p->bcg_fraction = 0.25;
  p->survey_area = 1.7;
  p->richness = 12;
  p->min_window = 8; p->min_count = 8; p->min_virial = 12;
  p->n_random = 300; p->n_bins = 10; p->grid_cells = 64;
  return FOFB_OK;
}

int fofb_create(int device, fofb_handle** out) {
  if (!out) return FOFB_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0 || device < 0 || device >= count) {
    g_create_error = std::string("no usable CUDA device (") + (e != cudaSuccess ? cudaGetErrorString(e) : "index out of range") +
                     "); libfofb200 has no CPU fallback";
    return FOFB_ERR_NO_DEVICE;
  }
  fofb_handle* h = new (std::nothrow) fofb_handle();
  if (!h) return FOFB_ERR_INTERNAL;
  FOFB_TRY(h)
  h->device = device;
  FOFB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  FOFB_CUDA(cudaGetDeviceProperties(&prop, device));
  h->sm_count = prop.multiProcessorCount;
  FOFB_CUDA(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  h->stream = h->own_stream;
  FOFB_CUDA(cudaEventCreate(&h->ev0));
  FOFB_CUDA(cudaEventCreate(&h->ev1));
  *out = h;
  return FOFB_OK;
  }
  catch (const fofb::Error& err) {
    g_create_error = err.what();
    delete h;
    return err.code;
  }
}

void fofb_destroy(fofb_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  delete reinterpret_cast<Stage0Holder*>(h->catalog);
  h->catalog = nullptr;
  delete h->fof;
  h->fof = nullptr;
  delete h->clean;
  h->clean = nullptr;
  delete h->pipe;
  h->pipe = nullptr;
  delete h->trans;
  h->trans = nullptr;
  delete h->bubble;
  h->bubble = nullptr;
  delete h->matcher;
  h->matcher = nullptr;
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->upload_done) cudaEventDestroy(h->upload_done);
  if (h->upload_go) cudaEventDestroy(h->upload_go);
  delete h;
}

const char* fofb_last_error(const fofb_handle* h) { return h ? h->last_error.c_str() : g_create_error.c_str(); }

int fofb_synchronize(fofb_handle* h) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  FOFB_CUDA(cudaSetDevice(h->device));
  FOFB_CUDA(cudaStreamSynchronize(h->stream));
  return FOFB_OK;
  FOFB_CATCH(h)
}

int fofb_set_stream(fofb_handle* h, void* cuda_stream) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  FOFB_CUDA(cudaSetDevice(h->device));
  FOFB_CUDA(cudaStreamSynchronize(h->stream));
  h->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->own_stream;
  return FOFB_OK;
  FOFB_CATCH(h)
}

int fofb_last_timing(const fofb_handle* h, double* device_ms, int64_t* kernel_launches) {
  if (!h) return FOFB_ERR_INVALID;
  if (device_ms) *device_ms = h->last_ms;
  if (kernel_launches) *kernel_launches = h->launches;
  return FOFB_OK;
}

int fofb_profile_enable(fofb_handle* h, int32_t on) {
  if (!h) return FOFB_ERR_INVALID;
  h->prof.on = on != 0;
  if (on == 2) h->prof.acc.clear();  // 2 = enable and reset the accumulators
  return FOFB_OK;
}

int fofb_profile_read(fofb_handle* h, char* buf, int64_t capacity) {
  if (!h || !buf || capacity <= 0) return FOFB_ERR_INVALID;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  h->prof.resolve();  // events of entry points that do not end with a timer (the distributed steps)
  std::string out;
  for (auto& kv : h->prof.acc) out += kv.first + " " + std::to_string(kv.second.first) + " " + std::to_string(kv.second.second) + "\n";
  if ((int64_t)out.size() + 1 > capacity) return FOFB_ERR_INVALID;
  std::memcpy(buf, out.c_str(), out.size() + 1);
  return FOFB_OK;
}

// ---- host scalar helpers ------------------------------------------------------------------------
int fofb_comoving_distance(const fofb_params* p, int64_t n, const double* z, double* out) {
  fofb_handle* h = nullptr;
  FOFB_TRY(h)
  check_params(p);
  std::vector<double> t;
  const Cosmo c = host_cosmo_with(*p, t);
  for (int64_t i = 0; i < n; ++i) out[i] = cosmo_comoving_distance(c, z[i]);
  return FOFB_OK;
  FOFB_CATCH(h)
}

int fofb_angular_diameter_distance(const fofb_params* p, int64_t n, const double* z, double* out) {
  fofb_handle* h = nullptr;
  FOFB_TRY(h)
  check_params(p);
  std::vector<double> t;
  const Cosmo c = host_cosmo_with(*p, t);
  for (int64_t i = 0; i < n; ++i) out[i] = cosmo_angular_diameter_distance(c, z[i]);
  return FOFB_OK;
  FOFB_CATCH(h)
}

int fofb_linear_to_angular_dist(const fofb_params* p, double distance_mpc_h, int64_t n, const double* z, double* out) {
  fofb_handle* h = nullptr;
  FOFB_TRY(h)
  check_params(p);
  std::vector<double> t;
  const Cosmo c = host_cosmo_with(*p, t);
  for (int64_t i = 0; i < n; ++i) out[i] = cosmo_linear_to_angular_dist(c, distance_mpc_h, z[i]);
  return FOFB_OK;
  FOFB_CATCH(h)
}

int fofb_mean_separation(const fofb_params* p, double n_objects, double z, double* out) {
  fofb_handle* h = nullptr;
  FOFB_TRY(h)
  check_params(p);
  std::vector<double> t;
  const Cosmo c = host_cosmo_with(*p, t);
  *out = cosmo_mean_separation(c, n_objects, z, p->max_velocity, p->survey_area);
  return FOFB_OK;
  FOFB_CATCH(h)
}

int fofb_redshift_to_velocity(int64_t n, const double* z, double center_z, double* out) {
  if (!z || !out) return FOFB_ERR_INVALID;
  for (int64_t i = 0; i < n; ++i) out[i] = fofb::redshift_to_velocity(z[i], center_z);
  return FOFB_OK;
}

// ---- stage 0 ------------------------------------------------------------------------------------------
int fofb_number_counts(fofb_handle* h, const fofb_params* p, const fofb_matrix* galaxies, int32_t* N_out,
                       int32_t N_mem, int64_t* n_candidates) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  check_params(p);
  FOFB_REQUIRE(galaxies && N_out, "galaxies / N_out is null");
  FOFB_CUDA(cudaSetDevice(h->device));
  if (!h->catalog) h->catalog = reinterpret_cast<Catalog*>(new Stage0Holder());
  Stage0Holder* s = reinterpret_cast<Stage0Holder*>(h->catalog);
  s->s0.valid = false;
  Timed timer(h);
  const View v = to_device(h, *galaxies, h->upload);
  s->cat.build(h, v);
  s->s0.run(h, *p, s->cat, s->cells);
  FOFB_CUDA(cudaMemcpyAsync(N_out, s->s0.N_row.p, sizeof(int32_t) * s->cat.n,
                            N_mem == 1 ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
  timer.stop();
  if (n_candidates) *n_candidates = s->s0.n_candidates;
  return FOFB_OK;
  FOFB_CATCH(h)
}

int fofb_candidate_order(fofb_handle* h, int32_t* rows_out, int32_t rows_mem, int64_t capacity) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  Stage0Holder* s = reinterpret_cast<Stage0Holder*>(h->catalog);
  if (!s || !s->s0.valid) throw fofb::Error(FOFB_ERR_STATE, "fofb_candidate_order called before fofb_number_counts");
  FOFB_REQUIRE(rows_out && capacity >= s->s0.n_candidates, "rows_out too small");
  FOFB_CUDA(cudaSetDevice(h->device));
  if (s->s0.n_candidates > 0)
    FOFB_CUDA(cudaMemcpyAsync(rows_out, s->s0.cand_rows.p, sizeof(int32_t) * s->s0.n_candidates,
                              rows_mem == 1 ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
  FOFB_CUDA(cudaStreamSynchronize(h->stream));
  return FOFB_OK;
  FOFB_CATCH(h)
}

// ---- stages 1-4 ---------------------------------------------------------------------------------------
int fofb_fof_run(fofb_handle* h, const fofb_params* p, const fofb_matrix* galaxies, const fofb_matrix* centres,
                 int64_t* n_clusters, double* bbox) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  check_params(p);
  FOFB_REQUIRE(galaxies && centres, "galaxies / centres is null");
  FOFB_CUDA(cudaSetDevice(h->device));
  if (!h->fof) h->fof = new FofState();
  h->fof->ran = h->fof->finished = false;
  h->fof->adopt(nullptr, nullptr);
  Timed timer(h);
  // the state owns its copies: other entry points on this handle reuse h->upload between run / finish / fetch
  const View G = to_device(h, *galaxies, h->fof->g_store);
  const View Cn = to_device(h, *centres, h->fof->c_store);
  FOFB_REQUIRE(G.rows > 0, "galaxy_data is empty");
  h->fof->run(h, *p, G, Cn);
  timer.stop();
  if (n_clusters) *n_clusters = h->fof->merged.K;
  if (bbox) {
    bbox[0] = h->fof->cat_().gbox.ra_min; bbox[1] = h->fof->cat_().gbox.ra_max;
    bbox[2] = h->fof->cat_().gbox.dec_min; bbox[3] = h->fof->cat_().gbox.dec_max;
  }
  return FOFB_OK;
  FOFB_CATCH(h)
}

int fofb_fof_finish(fofb_handle* h, const double* rand_ra, const double* rand_dec, int32_t mem, int64_t* n_final) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  if (!h->fof || !h->fof->ran) throw fofb::Error(FOFB_ERR_STATE, "fofb_fof_finish called before fofb_fof_run");
  FOFB_CUDA(cudaSetDevice(h->device));
  const double prev_ms = h->last_ms;
  const int64_t prev_launches = h->launches;
  Timed timer(h);
  h->fof->finish(h, rand_ra, rand_dec, mem);
  timer.stop();
  h->last_ms += prev_ms;  // run + finish are one FoF() call
  h->launches += prev_launches;
  if (n_final) *n_final = h->fof->final_.K;
  return FOFB_OK;
  FOFB_CATCH(h)
}

int fofb_fof_finish_mt(fofb_handle* h, uint32_t* mt_key, int32_t* mt_pos, int64_t* n_final) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  if (!h->fof || !h->fof->ran) throw fofb::Error(FOFB_ERR_STATE, "fofb_fof_finish_mt called before fofb_fof_run");
  FOFB_CUDA(cudaSetDevice(h->device));
  const double prev_ms = h->last_ms;
  const int64_t prev_launches = h->launches;
  Timed timer(h);
  h->fof->finish_mt(h, mt_key, mt_pos);
  timer.stop();
  h->last_ms += prev_ms;
  h->launches += prev_launches;
  if (n_final) *n_final = h->fof->final_.K;
  return FOFB_OK;
  FOFB_CATCH(h)
}

static const ClusterList* pick_list(fofb_handle* h, int32_t which) {
  if (!h->fof || !h->fof->ran) throw fofb::Error(FOFB_ERR_STATE, "no FoF result: call fofb_fof_run first");
  switch (which) {
    case 0: return &h->fof->candidates;
    case 1: return &h->fof->merged;
    case 2: return &h->fof->merged;  // the BCG stage never replaces a centre on the device path
    case 3:
      if (!h->fof->finished) throw fofb::Error(FOFB_ERR_STATE, "final list requested before fofb_fof_finish");
      return &h->fof->final_;
    default: throw fofb::Error(FOFB_ERR_INVALID, "which must be 0..3");
  }
}

int fofb_fof_sizes(fofb_handle* h, int32_t which, int64_t* n_clusters, int64_t* n_members) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  const ClusterList* L = pick_list(h, which);
  if (n_clusters) *n_clusters = L->K;
  if (n_members) *n_members = L->K ? L->total : 0;
  return FOFB_OK;
  FOFB_CATCH(h)
}

int fofb_fof_fetch(fofb_handle* h, int32_t which, int64_t* offsets, int32_t* members, int32_t* centre, double* N,
                   double* D) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  const ClusterList* L = pick_list(h, which);
  FOFB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  const int K = L->K;
  if (K == 0) {
    if (offsets) offsets[0] = 0;
    return FOFB_OK;
  }
  static_assert(sizeof(long long) == sizeof(int64_t), "offset width");
  if (offsets) FOFB_CUDA(cudaMemcpyAsync(offsets, L->off.p, sizeof(int64_t) * (K + 1), cudaMemcpyDeviceToHost, st));
  if (members && L->total) FOFB_CUDA(cudaMemcpyAsync(members, L->rows.p, sizeof(int32_t) * L->total, cudaMemcpyDeviceToHost, st));
  std::vector<int32_t> cidx(K);
  FOFB_CUDA(cudaMemcpyAsync(cidx.data(), L->centre.p, sizeof(int32_t) * K, cudaMemcpyDeviceToHost, st));
  std::vector<int32_t> Nint;
  if (which == 3) {
    if (N) {
      Nint.resize(K);
      FOFB_CUDA(cudaMemcpyAsync(Nint.data(), L->N.p, sizeof(int32_t) * K, cudaMemcpyDeviceToHost, st));
    }
    if (D) FOFB_CUDA(cudaMemcpyAsync(D, L->D.p, sizeof(double) * K, cudaMemcpyDeviceToHost, st));
  }
  FOFB_CUDA(cudaStreamSynchronize(st));
  if (centre) std::memcpy(centre, cidx.data(), sizeof(int32_t) * K);
  if (which == 3) {
    if (N) for (int k = 0; k < K; ++k) N[k] = (double)Nint[k];
  } else {
    if (N) {  // before stage 4 Cluster.N is still the centre's N(0.5) column (cluster.py:38)
      std::vector<double> cN(h->fof->m);
      FOFB_CUDA(cudaMemcpy(cN.data(), h->fof->c_N.p, sizeof(double) * h->fof->m, cudaMemcpyDeviceToHost));
      for (int k = 0; k < K; ++k) N[k] = cN[cidx[k]];
    }
    if (D) for (int k = 0; k < K; ++k) D[k] = 0.0;
  }
  return FOFB_OK;
  FOFB_CATCH(h)
}

int fofb_fof_stats(fofb_handle* h, int64_t* rounds, int64_t* evaluated) {
  if (!h || !h->fof) return FOFB_ERR_STATE;
  if (rounds) *rounds = h->fof->rounds;
  if (evaluated) *evaluated = h->fof->evaluated;
  return FOFB_OK;
}

// ---- stages 5-6 ---------------------------------------------------------------------------------------
int fofb_three_sigma(fofb_handle* h, const fofb_params* p, int64_t K, const int64_t* offsets, const fofb_matrix* members,
                     const double* centres, int64_t* out_offsets, int32_t* out_index) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  check_params(p);
  FOFB_REQUIRE(offsets && members && centres && out_offsets && out_index, "three_sigma: null argument");
  FOFB_REQUIRE(K >= 0 && K < (1ll << 31), "three_sigma: bad K");
  FOFB_CUDA(cudaSetDevice(h->device));
  if (!h->clean) h->clean = new CleanState();
  Timed timer(h);
  const View M = to_device(h, *members, h->upload);
  h->clean->three_sigma(h, *p, (int)K, reinterpret_cast<const long long*>(offsets), M, centres, p->n_bins,
                        reinterpret_cast<long long*>(out_offsets), out_index);
  timer.stop();
  return FOFB_OK;
  FOFB_CATCH(h)
}

int fofb_virial(fofb_handle* h, const fofb_params* p, int64_t K, const int64_t* offsets, const fofb_matrix* members,
                double* out) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  check_params(p);
  FOFB_REQUIRE(offsets && members && out, "virial: null argument");
  FOFB_REQUIRE(K >= 0 && K < (1ll << 31), "virial: bad K");
  FOFB_CUDA(cudaSetDevice(h->device));
  if (!h->clean) h->clean = new CleanState();
  Timed timer(h);
  const View M = to_device(h, *members, h->upload);
  h->clean->virial(h, *p, (int)K, reinterpret_cast<const long long*>(offsets), M, out);
  timer.stop();
  return FOFB_OK;
  FOFB_CATCH(h)
}

int fofb_projected_mass(fofb_handle* h, const fofb_params* p, int64_t K, const int64_t* offsets, const fofb_matrix* members,
                        double* out) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  check_params(p);
  FOFB_REQUIRE(offsets && members && out, "projected: null argument");
  FOFB_REQUIRE(K >= 0 && K < (1ll << 31), "projected: bad K");
  FOFB_CUDA(cudaSetDevice(h->device));
  if (!h->clean) h->clean = new CleanState();
  Timed timer(h);
  const View M = to_device(h, *members, h->upload);
  h->clean->projected(h, *p, (int)K, reinterpret_cast<const long long*>(offsets), M, out);
  timer.stop();
  return FOFB_OK;
  FOFB_CATCH(h)
}

int fofb_bootstrap_indices(int64_t n, int64_t count, int32_t* out) {
  if (n < 1 || count < 0 || !out) return FOFB_ERR_INVALID;
  if (n == 1) {
    for (int64_t i = 0; i < count; ++i) out[i] = 0;
    return FOFB_OK;
  }
  fofb::bootstrap_indices_host(n, count, out);
  return FOFB_OK;
}

// ---- whole job -------------------------------------------------------------------------------------------
int fofb_pipeline_begin(fofb_handle* h, const fofb_params* p, const fofb_matrix* galaxies, double zmin, double zmax_gal,
                        double zmax_cen, int64_t* n_clusters) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  check_params(p);
  FOFB_REQUIRE(galaxies, "galaxies is null");
  FOFB_CUDA(cudaSetDevice(h->device));
  if (!h->pipe) h->pipe = new Pipeline();
  Timed timer(h);
  const View raw = to_device(h, *galaxies, h->upload);
  h->pipe->begin(h, *p, raw, zmin, zmax_gal, zmax_cen);
  timer.stop();
  if (n_clusters) *n_clusters = h->pipe->fof.merged.K;
  return FOFB_OK;
  FOFB_CATCH(h)
}

int fofb_pipeline_finish(fofb_handle* h, const double* u, int32_t mem, int64_t* n_final, int64_t* n_members) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  if (!h->pipe || !h->pipe->began) throw fofb::Error(FOFB_ERR_STATE, "fofb_pipeline_finish called before fofb_pipeline_begin");
  FOFB_CUDA(cudaSetDevice(h->device));
  const double prev_ms = h->last_ms;
  const int64_t prev_launches = h->launches;
  Timed timer(h);
  h->pipe->finish(h, u, mem);
  timer.stop();
  h->last_ms += prev_ms;
  h->launches += prev_launches;
  if (n_final) *n_final = h->pipe->K_out;
  if (n_members) *n_members = h->pipe->total_out;
  return FOFB_OK;
  FOFB_CATCH(h)
}

int fofb_pipeline_finish_mt(fofb_handle* h, uint32_t* mt_key, int32_t* mt_pos, int64_t* n_final, int64_t* n_members) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  if (!h->pipe || !h->pipe->began) throw fofb::Error(FOFB_ERR_STATE, "fofb_pipeline_finish_mt called before fofb_pipeline_begin");
  FOFB_REQUIRE(mt_key && mt_pos, "MT19937 state is null");
  FOFB_CUDA(cudaSetDevice(h->device));
  const double prev_ms = h->last_ms;
  const int64_t prev_launches = h->launches;
  Timed timer(h);
  h->pipe->finish(h, nullptr, 0, mt_key, mt_pos);
  timer.stop();
  h->last_ms += prev_ms;
  h->launches += prev_launches;
  if (n_final) *n_final = h->pipe->K_out;
  if (n_members) *n_members = h->pipe->total_out;
  return FOFB_OK;
  FOFB_CATCH(h)
}

int fofb_pipeline_run(fofb_handle* h, const fofb_params* p, const fofb_matrix* galaxies, double zmin, double zmax_gal,
                      double zmax_cen, uint32_t* mt_key, int32_t* mt_pos, int64_t* n_final, int64_t* n_members) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  check_params(p);
  FOFB_REQUIRE(galaxies && mt_key && mt_pos, "galaxies / MT19937 state is null");
  FOFB_CUDA(cudaSetDevice(h->device));
  if (!h->pipe) h->pipe = new Pipeline();
  Timed timer(h);
  const View raw = to_device_split(h, *galaxies, h->upload, 3);
  h->pipe->begin(h, *p, raw, zmin, zmax_gal, zmax_cen, mt_key, *mt_pos);
  h->pipe->finish(h, nullptr, 0, mt_key, mt_pos);
  timer.stop();
  if (n_final) *n_final = h->pipe->K_out;
  if (n_members) *n_members = h->pipe->total_out;
  return FOFB_OK;
  FOFB_CATCH(h)
}

int fofb_pipeline_rerun(fofb_handle* h, double linking_length_factor, int32_t richness, double overdensity, uint32_t* mt_key,
                        int32_t* mt_pos, int64_t* n_final, int64_t* n_members) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  if (!h->pipe) throw fofb::Error(FOFB_ERR_STATE, "fofb_pipeline_rerun called before fofb_pipeline_run");
  FOFB_REQUIRE(mt_key && mt_pos, "MT19937 state is null");
  FOFB_CUDA(cudaSetDevice(h->device));
  Timed timer(h);
  h->pipe->rerun(h, linking_length_factor, richness, overdensity, mt_key, *mt_pos);
  h->pipe->finish(h, nullptr, 0, mt_key, mt_pos);
  timer.stop();
  if (n_final) *n_final = h->pipe->K_out;
  if (n_members) *n_members = h->pipe->total_out;
  return FOFB_OK;
  FOFB_CATCH(h)
}

int fofb_pipeline_fetch(fofb_handle* h, int64_t* offsets, int32_t* member_rows, int32_t* centre_row, double* N, double* D,
                        double* props) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  if (!h->pipe || !h->pipe->finished) throw fofb::Error(FOFB_ERR_STATE, "fofb_pipeline_fetch called before fofb_pipeline_finish");
  FOFB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  Pipeline& P = *h->pipe;
  const int K = P.K_out;
  if (K == 0) {
    if (offsets) offsets[0] = 0;
    return FOFB_OK;
  }
  if (offsets) FOFB_CUDA(cudaMemcpyAsync(offsets, P.out_off.p, sizeof(int64_t) * (K + 1), cudaMemcpyDefault, st));
  if (member_rows) FOFB_CUDA(cudaMemcpyAsync(member_rows, P.out_rows.p, sizeof(int32_t) * P.total_out, cudaMemcpyDefault, st));
  if (centre_row) FOFB_CUDA(cudaMemcpyAsync(centre_row, P.out_centre.p, sizeof(int32_t) * K, cudaMemcpyDefault, st));
  if (N) FOFB_CUDA(cudaMemcpyAsync(N, P.out_N.p, sizeof(double) * K, cudaMemcpyDefault, st));
  if (D) FOFB_CUDA(cudaMemcpyAsync(D, P.out_D.p, sizeof(double) * K, cudaMemcpyDefault, st));
  if (props) FOFB_CUDA(cudaMemcpyAsync(props, P.props.p, sizeof(double) * K * 8, cudaMemcpyDefault, st));
  FOFB_CUDA(cudaStreamSynchronize(st));
  return FOFB_OK;
  FOFB_CATCH(h)
}

int fofb_pipeline_fetch_counts(fofb_handle* h, int32_t* N_row, int32_t mem) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  if (!h->pipe || !h->pipe->s0.valid) throw fofb::Error(FOFB_ERR_STATE, "fofb_pipeline_fetch_counts called before a pipeline run");
  FOFB_REQUIRE(N_row, "N_row is null");
  FOFB_CUDA(cudaSetDevice(h->device));
  FOFB_CUDA(cudaMemcpyAsync(N_row, h->pipe->s0.N_row.p, sizeof(int32_t) * h->pipe->n_in,
                            mem == 1 ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
  FOFB_CUDA(cudaStreamSynchronize(h->stream));
  return FOFB_OK;
  FOFB_CATCH(h)
}

int fofb_pipeline_stats(fofb_handle* h, int64_t* s) {
  if (!h || !h->pipe || !s) return FOFB_ERR_STATE;
  Pipeline& P = *h->pipe;
  s[0] = P.n_g; s[1] = P.m_c; s[2] = P.fof.candidates.K; s[3] = P.fof.candidates.total;
  s[4] = P.fof.rounds; s[5] = P.fof.evaluated; s[6] = P.fof.merged.K;
  return FOFB_OK;
}

// ---- redshift-slab decomposition: the steps a multi-process driver interleaves with its collectives ------------
static Pipeline& dist_pipe(fofb_handle* h) {
  if (!h->pipe) throw fofb::Error(FOFB_ERR_STATE, "no distributed run in progress: call fofb_dist_begin first");
  return *h->pipe;
}

int fofb_dist_begin(fofb_handle* h, const fofb_params* p, const fofb_matrix* galaxies, const double* zcut3,
                    const double* z_valid2, const double* z_own2, int64_t* n_fof_rows, int64_t* n_local_centres) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  check_params(p);
  FOFB_REQUIRE(galaxies && zcut3 && z_valid2 && z_own2, "fofb_dist_begin: null argument");
  FOFB_CUDA(cudaSetDevice(h->device));
  if (!h->pipe) h->pipe = new Pipeline();
  Timed timer(h);
  const View raw = to_device(h, *galaxies, h->upload);
  h->pipe->begin_dist(h, *p, raw, zcut3[0], zcut3[1], zcut3[2], z_valid2[0], z_valid2[1], z_own2[0], z_own2[1]);
  timer.stop();
  if (n_fof_rows) *n_fof_rows = h->pipe->n_g;
  if (n_local_centres) *n_local_centres = h->pipe->m_c;
  return FOFB_OK;
  FOFB_CATCH(h)
}

/* start generating the MT19937 stream for about k_spec clusters on the side stream (overlaps the rounds) */
int fofb_dist_prefetch_mt(fofb_handle* h, const uint32_t* mt_key, int32_t mt_pos, int64_t k_spec) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  Pipeline& P = dist_pipe(h);
  FOFB_CUDA(cudaSetDevice(h->device));
  P.fof.params_n_random_hint = P.params.n_random;
  P.fof.mt_prefetch(h, mt_key, mt_pos, k_spec);
  return FOFB_OK;
  FOFB_CATCH(h)
}

/* Slab runs: generate only share `part` of `parts` of the stream (whole chunks), to be all-gathered by the caller.
 * Returns the device address of the word buffer and where the shares sit in it: words[blocks_offset + r * share_words ..)
 * is share r.  fofb_dist_mt_wait makes the handle's stream wait for the generator's side stream. */
int fofb_dist_prefetch_mt_part(fofb_handle* h, const uint32_t* mt_key, int32_t mt_pos, int64_t k_spec, int32_t part, int32_t parts,
                               uint64_t* words_ptr, int64_t* blocks_offset, int64_t* share_words) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  Pipeline& P = dist_pipe(h);
  FOFB_REQUIRE(words_ptr && blocks_offset && share_words, "fofb_dist_prefetch_mt_part: null output");
  FOFB_CUDA(cudaSetDevice(h->device));
  P.fof.params_n_random_hint = P.params.n_random;
  P.fof.mt_prefetch(h, mt_key, mt_pos, k_spec, part, parts);
  *words_ptr = reinterpret_cast<uint64_t>(P.fof.mt_words.p);
  *blocks_offset = 624;
  *share_words = P.fof.mt_parts > 1 ? P.fof.mt_blocks / P.fof.mt_parts * 624 : 0;  // 0: generated in full here
  return FOFB_OK;
  FOFB_CATCH(h)
}

int fofb_dist_mt_wait(fofb_handle* h) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  Pipeline& P = dist_pipe(h);
  FOFB_CUDA(cudaSetDevice(h->device));
  if (P.fof.rng_done && P.fof.mt_prefetched) FOFB_CUDA(cudaStreamWaitEvent(h->stream, P.fof.rng_done, 0));
  return FOFB_OK;
  FOFB_CATCH(h)
}

int fofb_dist_local_inputs(fofb_handle* h, double* centre_rows, int32_t* centre_src_row, int32_t* fof_src_row, double* bbox4) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  Pipeline& P = dist_pipe(h);
  FOFB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  if (centre_rows && P.m_c) FOFB_CUDA(cudaMemcpyAsync(centre_rows, P.c_rows.p, sizeof(double) * 8 * P.m_c, cudaMemcpyDefault, st));
  if (centre_src_row && P.m_c) FOFB_CUDA(cudaMemcpyAsync(centre_src_row, P.c_src.p, sizeof(int32_t) * P.m_c, cudaMemcpyDefault, st));
  if (fof_src_row && P.n_g) FOFB_CUDA(cudaMemcpyAsync(fof_src_row, P.g_src.p, sizeof(int32_t) * P.n_g, cudaMemcpyDefault, st));
  FOFB_CUDA(cudaStreamSynchronize(st));
  (void)bbox4;
  return FOFB_OK;
  FOFB_CATCH(h)
}

int fofb_dist_prepare(fofb_handle* h, const double* centres, int64_t n_centres, const uint8_t* owned, double* local_bbox4) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  Pipeline& P = dist_pipe(h);
  FOFB_REQUIRE(centres && owned && n_centres >= 0 && n_centres < (1ll << 31), "fofb_dist_prepare: bad argument");
  FOFB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  Timed timer(h);
  const size_t m = (size_t)n_centres;
  double* c = P.cn_glob.alloc(std::max<size_t>(m, 1) * 8);
  unsigned char* o = P.owned.alloc(std::max<size_t>(m, 1));
  FOFB_CUDA(cudaMemcpyAsync(c, centres, sizeof(double) * 8 * m, cudaMemcpyDefault, st));
  FOFB_CUDA(cudaMemcpyAsync(o, owned, m, cudaMemcpyDefault, st));
  View Cv;
  Cv.data = c; Cv.rows = n_centres; Cv.cols = 8; Cv.rs = 8; Cv.cs = 1;
  P.fof.prepare(h, P.params, P.Gv, Cv, o);
  timer.stop();
  if (local_bbox4) {  // bounding box of this process's FoF() rows; the driver reduces it over processes
    local_bbox4[0] = P.fof.cat_().gbox.ra_min; local_bbox4[1] = P.fof.cat_().gbox.ra_max;
    local_bbox4[2] = P.fof.cat_().gbox.dec_min; local_bbox4[3] = P.fof.cat_().gbox.dec_max;
  }
  return FOFB_OK;
  FOFB_CATCH(h)
}

int fofb_dist_round_evaluate(fofb_handle* h) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  Pipeline& P = dist_pipe(h);
  FOFB_CUDA(cudaSetDevice(h->device));
  P.fof.round_evaluate(h);
  return FOFB_OK;
  FOFB_CATCH(h)
}

/* direction 0: library -> caller buffers (device pointers, n_centres bytes each); 1: caller -> library */
int fofb_dist_state(fofb_handle* h, uint8_t* alive, uint8_t* decided, int32_t direction) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  Pipeline& P = dist_pipe(h);
  FOFB_REQUIRE(alive && decided, "fofb_dist_state: null buffer");
  FOFB_CUDA(cudaSetDevice(h->device));
  const size_t m = (size_t)P.fof.m;
  if (m) {
    if (direction == 0) {
      FOFB_CUDA(cudaMemcpyAsync(alive, P.fof.alive.p, m, cudaMemcpyDeviceToDevice, h->stream));
      FOFB_CUDA(cudaMemcpyAsync(decided, P.fof.decided.p, m, cudaMemcpyDeviceToDevice, h->stream));
    } else {
      FOFB_CUDA(cudaMemcpyAsync(P.fof.alive.p, alive, m, cudaMemcpyDeviceToDevice, h->stream));
      FOFB_CUDA(cudaMemcpyAsync(P.fof.decided.p, decided, m, cudaMemcpyDeviceToDevice, h->stream));
    }
  }
  FOFB_CUDA(cudaStreamSynchronize(h->stream));
  return FOFB_OK;
  FOFB_CATCH(h)
}

int fofb_dist_round_advance(fofb_handle* h, int64_t* n_active) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  Pipeline& P = dist_pipe(h);
  FOFB_CUDA(cudaSetDevice(h->device));
  P.fof.round_advance(h);
  if (n_active) *n_active = P.fof.na;
  return FOFB_OK;
  FOFB_CATCH(h)
}

int fofb_dist_state_pack(fofb_handle* h, int64_t n_shared, const int64_t* local_pos, const int64_t* shared_pos, uint8_t* buf,
                         int64_t n_buf) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  Pipeline& P = dist_pipe(h);
  FOFB_REQUIRE(buf && n_buf >= 1 && n_shared >= 0 && (n_shared == 0 || (local_pos && shared_pos)), "fofb_dist_state_pack: bad argument");
  FOFB_CUDA(cudaSetDevice(h->device));
  fofb::dist_state_pack(h, P.fof, n_shared, reinterpret_cast<const long long*>(local_pos), reinterpret_cast<const long long*>(shared_pos),
                        buf, n_buf);
  return FOFB_OK;
  FOFB_CATCH(h)
}

int fofb_dist_state_unpack(fofb_handle* h, int64_t n_shared, const int64_t* local_pos, const int64_t* shared_pos,
                           const uint8_t* buf) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  Pipeline& P = dist_pipe(h);
  FOFB_REQUIRE(buf && n_shared >= 0 && (n_shared == 0 || (local_pos && shared_pos)), "fofb_dist_state_unpack: bad argument");
  FOFB_CUDA(cudaSetDevice(h->device));
  fofb::dist_state_unpack(h, P.fof, n_shared, reinterpret_cast<const long long*>(local_pos), reinterpret_cast<const long long*>(shared_pos), buf);
  return FOFB_OK;
  FOFB_CATCH(h)
}

int fofb_dist_round_advance2(fofb_handle* h, const uint8_t* buf, int64_t n_buf, int64_t* n_active, int32_t* any_active) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  Pipeline& P = dist_pipe(h);
  FOFB_REQUIRE(buf && n_buf >= 1 && any_active, "fofb_dist_round_advance2: bad argument");
  FOFB_CUDA(cudaSetDevice(h->device));
  // its own pinned word: d2h_scalar (common.cuh) stages through word 0 of the same buffer while the rounds advance
  uint8_t* flag = reinterpret_cast<uint8_t*>(h->pinned_i64.alloc(16) + 8);
  FOFB_CUDA(cudaMemcpyAsync(flag, buf + (n_buf - 1), 1, cudaMemcpyDeviceToHost, h->stream));
  P.fof.round_advance(h);  // synchronises the stream when there is anything to advance
  FOFB_CUDA(cudaStreamSynchronize(h->stream));
  *any_active = *flag ? 1 : 0;
  if (n_active) *n_active = P.fof.na;
  return FOFB_OK;
  FOFB_CATCH(h)
}

int fofb_dist_finalize(fofb_handle* h, int64_t* n_clusters, int64_t* n_members) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  Pipeline& P = dist_pipe(h);
  FOFB_CUDA(cudaSetDevice(h->device));
  P.fof.finalize(h);
  if (n_clusters) *n_clusters = P.fof.candidates.K;
  if (n_members) *n_members = P.fof.candidates.K ? P.fof.candidates.total : 0;
  return FOFB_OK;
  FOFB_CATCH(h)
}

/* local stage-1 clusters: centre[K] = index into the all-process centre table, offsets[K+1], ids[total] = column 6 */
int fofb_dist_local_clusters(fofb_handle* h, int32_t* centre, int64_t* offsets, double* ids) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  Pipeline& P = dist_pipe(h);
  FOFB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  const ClusterList& L = P.fof.candidates;
  if (L.K == 0) {
    if (offsets) offsets[0] = 0;
    return FOFB_OK;
  }
  if (centre) FOFB_CUDA(cudaMemcpyAsync(centre, L.centre.p, sizeof(int32_t) * L.K, cudaMemcpyDeviceToHost, st));
  if (offsets) FOFB_CUDA(cudaMemcpyAsync(offsets, L.off.p, sizeof(int64_t) * (L.K + 1), cudaMemcpyDeviceToHost, st));
  if (ids) {
    std::vector<int32_t> rows(L.total);
    FOFB_CUDA(cudaMemcpyAsync(rows.data(), L.rows.p, sizeof(int32_t) * L.total, cudaMemcpyDeviceToHost, st));
    std::vector<double> col6(P.n_g);
    FOFB_CUDA(cudaMemcpy2DAsync(col6.data(), sizeof(double), P.g_rows.p + 6, sizeof(double) * 8, sizeof(double), P.n_g,
                                cudaMemcpyDeviceToHost, st));
    FOFB_CUDA(cudaStreamSynchronize(st));
    for (long long t = 0; t < L.total; ++t) ids[t] = col6[rows[t]];
  }
  FOFB_CUDA(cudaStreamSynchronize(st));
  return FOFB_OK;
  FOFB_CATCH(h)
}

/* stage 2 on the all-process cluster list (host arrays, ordered by centre index); keep[K] out */
int fofb_dist_overlap(fofb_handle* h, int64_t K, const int32_t* centre, const int64_t* offsets, const double* ids, int32_t* keep) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  Pipeline& P = dist_pipe(h);
  FOFB_REQUIRE(K >= 0 && K < (1ll << 31), "fofb_dist_overlap: bad K");
  if (K == 0) return FOFB_OK;
  FOFB_REQUIRE(centre && offsets && ids && keep, "fofb_dist_overlap: null argument");
  FOFB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  const long long total = offsets[K];
  int* c = P.ext_centre.alloc((size_t)K);
  long long* o = P.ext_off.alloc((size_t)K + 1);
  double* d = P.ext_ids.alloc((size_t)std::max<long long>(total, 1));
  int* kd = P.keep_dev.alloc((size_t)K);
  FOFB_CUDA(cudaMemcpyAsync(c, centre, sizeof(int32_t) * K, cudaMemcpyHostToDevice, st));
  FOFB_CUDA(cudaMemcpyAsync(o, offsets, sizeof(int64_t) * (K + 1), cudaMemcpyHostToDevice, st));
  FOFB_CUDA(cudaMemcpyAsync(d, ids, sizeof(double) * total, cudaMemcpyHostToDevice, st));
  fof_overlap_external(h, P.fof, (int)K, c, o, total, d, kd);
  FOFB_CUDA(cudaMemcpyAsync(keep, kd, sizeof(int32_t) * K, cudaMemcpyDeviceToHost, st));
  FOFB_CUDA(cudaStreamSynchronize(st));
  return FOFB_OK;
  FOFB_CATCH(h)
}

/* stage 2 on the all-process cluster list given as DEVICE arrays ordered by priority: centre_rows[K x 8], offsets[K+1],
 * ids[total]; keep[K] (device, int32) out */
int fofb_dist_overlap_rows(fofb_handle* h, int64_t K, const double* centre_rows, const int64_t* offsets, int64_t total,
                           const double* ids, int32_t* keep) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  Pipeline& P = dist_pipe(h);
  FOFB_REQUIRE(K >= 0 && K < (1ll << 31), "fofb_dist_overlap_rows: bad K");
  if (K == 0) return FOFB_OK;
  FOFB_REQUIRE(centre_rows && offsets && ids && keep, "fofb_dist_overlap_rows: null argument");
  FOFB_CUDA(cudaSetDevice(h->device));
  fof_overlap_external_rows(h, P.fof, (int)K, centre_rows, reinterpret_cast<const long long*>(offsets), total, ids, keep);
  return FOFB_OK;
  FOFB_CATCH(h)
}

int fofb_dist_overlap_events(fofb_handle* h, int64_t K, const double* centre_rows, const int64_t* offsets, int64_t total,
                             const double* ids, int64_t* n_events) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  Pipeline& P = dist_pipe(h);
  FOFB_REQUIRE(K >= 0 && K < (1ll << 31) && n_events, "fofb_dist_overlap_events: bad argument");
  *n_events = 0;
  if (K == 0) return FOFB_OK;
  FOFB_REQUIRE(centre_rows && offsets && (ids || total == 0), "fofb_dist_overlap_events: null argument");
  FOFB_CUDA(cudaSetDevice(h->device));
  *n_events = fof_overlap_events_rows(h, P.fof, (int)K, centre_rows, reinterpret_cast<const long long*>(offsets), total, ids);
  return FOFB_OK;
  FOFB_CATCH(h)
}

int fofb_dist_overlap_events_fetch(fofb_handle* h, int64_t n_events, int32_t* ev_a, int32_t* ev_c, int32_t* ev_loser) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  Pipeline& P = dist_pipe(h);
  FOFB_REQUIRE(n_events >= 0 && (n_events == 0 || (ev_a && ev_c && ev_loser)), "fofb_dist_overlap_events_fetch: bad argument");
  FOFB_CUDA(cudaSetDevice(h->device));
  fof_overlap_events_fetch(h, P.fof, n_events, ev_a, ev_c, ev_loser);
  return FOFB_OK;
  FOFB_CATCH(h)
}

int fofb_dist_overlap_resolve(fofb_handle* h, int64_t K, int64_t n_events, const int32_t* ev_a, const int32_t* ev_c,
                              const int32_t* ev_loser, int32_t* keep) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  Pipeline& P = dist_pipe(h);
  FOFB_REQUIRE(K >= 0 && K < (1ll << 31) && n_events >= 0 && n_events < (1ll << 31), "fofb_dist_overlap_resolve: bad size");
  if (K == 0) return FOFB_OK;
  FOFB_REQUIRE(keep && (n_events == 0 || (ev_a && ev_c && ev_loser)), "fofb_dist_overlap_resolve: null argument");
  FOFB_CUDA(cudaSetDevice(h->device));
  fof_overlap_resolve_external(h, P.fof, (int)K, n_events, ev_a, ev_c, ev_loser, keep);
  return FOFB_OK;
  FOFB_CATCH(h)
}

/* member ids (column 6) and sizes of the local stage-1 clusters as DEVICE arrays: centre[K] (int32 local centre index),
 * sizes[K] (int64), ids[total] (float64) */
int fofb_dist_local_clusters_dev(fofb_handle* h, int32_t* centre, int64_t* sizes, double* ids) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  Pipeline& P = dist_pipe(h);
  FOFB_CUDA(cudaSetDevice(h->device));
  P.export_local_clusters(h, centre, reinterpret_cast<long long*>(sizes), ids);
  return FOFB_OK;
  FOFB_CATCH(h)
}

/* keep_local[K_local]; global_index[number kept] = position of each kept local cluster in the all-process merged list */
int fofb_dist_set_merged(fofb_handle* h, const int32_t* keep_local, const int32_t* global_index, int64_t n_global_merged,
                         const double* global_bbox4) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  Pipeline& P = dist_pipe(h);
  FOFB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  const int K = P.fof.candidates.K;
  int kept = 0;
  if (K > 0) {
    FOFB_REQUIRE(keep_local, "fofb_dist_set_merged: null flags");
    int* kd = P.keep_dev.alloc((size_t)K);
    FOFB_CUDA(cudaMemcpyAsync(kd, keep_local, sizeof(int32_t) * K, cudaMemcpyDefault, st));
    fof_set_merged(h, P.fof, kd);
    kept = P.fof.merged.K;
  } else {
    P.fof.merged.K = 0;
    P.fof.merged.total = 0;
  }
  if (kept > 0) {
    FOFB_REQUIRE(global_index, "fofb_dist_set_merged: null indices");
    FOFB_CUDA(cudaMemcpyAsync(P.fof.merged_gidx.alloc(kept), global_index, sizeof(int32_t) * kept, cudaMemcpyDefault, st));
  }
  FOFB_CUDA(cudaStreamSynchronize(st));
  P.fof.K_glob_merged = n_global_merged;
  FOFB_REQUIRE(global_bbox4, "fofb_dist_set_merged: null bounding box");
  P.fof.have_global_bbox = true;
  P.fof.global_bbox = BBox{global_bbox4[0], global_bbox4[1], global_bbox4[2], global_bbox4[3]};
  P.fof.ran = true;
  P.began = true;
  return FOFB_OK;
  FOFB_CATCH(h)
}

// ---- helper-level bubble counts ------------------------------------------------------------------------------
int fofb_bubble_count(fofb_handle* h, const fofb_params* p, const fofb_matrix* points, int64_t n_centres, const double* centres,
                      const double* radius_deg, int32_t use_window, double window_z, int32_t* counts) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  check_params(p);
  FOFB_REQUIRE(points && centres && radius_deg && counts && n_centres >= 0 && n_centres < (1ll << 26), "bubble count: bad argument");
  FOFB_CUDA(cudaSetDevice(h->device));
  if (!h->bubble) h->bubble = new Bubble();
  Timed timer(h);
  View v;
  if (points->rows > 0) {
    v = to_device(h, *points, h->upload);
  } else {
    v.rows = 0; v.cols = points->cols;
  }
  h->bubble->count(h, *p, v, (int)n_centres, centres, radius_deg, use_window, window_z, counts);
  timer.stop();
  return FOFB_OK;
  FOFB_CATCH(h)
}

// ---- numpy's legacy MT19937 stream on the device (the generator behind fofb_*_finish_mt) ------------------------
int fofb_mt19937_words(fofb_handle* h, uint32_t* key624, int32_t* pos, int64_t n_words, uint32_t* out, int32_t out_mem) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  FOFB_REQUIRE(out || n_words == 0, "mt19937 words: null output");
  FOFB_CUDA(cudaSetDevice(h->device));
  Timed timer(h);
  int p = pos ? *pos : -1;
  mt19937_words(h, key624, &p, n_words, out, out_mem);
  timer.stop();
  *pos = p;
  return FOFB_OK;
  FOFB_CATCH(h)
}

// ---- catalogue matching ---------------------------------------------------------------------------------------
int fofb_match_catalog(fofb_handle* h, const fofb_params* p, const fofb_matrix* sample, const fofb_matrix* catalogue,
                       double matching_distance_mpc_h, double z_cut, int32_t* match_row, double* match_sep_deg, int32_t out_mem,
                       int64_t* n_matched) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  check_params(p);
  FOFB_REQUIRE(sample && catalogue && match_row && match_sep_deg, "catalogue matching: null argument");
  FOFB_REQUIRE(sample->rows < (1ll << 31) && catalogue->rows < (1ll << 26), "catalogue matching: too many rows");
  FOFB_REQUIRE(matching_distance_mpc_h >= 0 && z_cut >= 0, "catalogue matching: negative distance or redshift cut");
  FOFB_CUDA(cudaSetDevice(h->device));
  if (!h->matcher) h->matcher = new Matcher();
  Timed timer(h);
  View vs, vc;
  if (sample->rows > 0) vs = to_device(h, *sample, h->upload); else { vs.rows = 0; vs.cols = sample->cols; }
  if (catalogue->rows > 0) vc = to_device(h, *catalogue, h->upload2); else { vc.rows = 0; vc.cols = catalogue->cols; }
  h->matcher->run(h, *p, vs, vc, matching_distance_mpc_h, z_cut);
  if (catalogue->rows > 0) {
    const cudaMemcpyKind kind = out_mem == 1 ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    FOFB_CUDA(cudaMemcpyAsync(match_row, h->matcher->match_row.p, sizeof(int32_t) * catalogue->rows, kind, h->stream));
    FOFB_CUDA(cudaMemcpyAsync(match_sep_deg, h->matcher->match_sep.p, sizeof(double) * catalogue->rows, kind, h->stream));
  }
  timer.stop();
  if (n_matched) *n_matched = h->matcher->n_matched;
  return FOFB_OK;
  FOFB_CATCH(h)
}

// ---- transitive mode (not a reference code path) ------------------------------------------------------------
int fofb_transitive_labels(fofb_handle* h, const fofb_params* p, const fofb_matrix* galaxies, double linking_length_mpc_h,
                           int32_t* labels, int32_t labels_mem, int64_t* n_groups, int64_t* n_links) {
  if (!h) return FOFB_ERR_INVALID;
  FOFB_TRY(h)
  check_params(p);
  FOFB_REQUIRE(galaxies && labels, "galaxies / labels is null");
  FOFB_CUDA(cudaSetDevice(h->device));
  if (!h->trans) h->trans = new Transitive();
  Timed timer(h);
  const View v = to_device(h, *galaxies, h->upload);
  h->trans->run(h, *p, v, linking_length_mpc_h, n_links != nullptr);
  FOFB_CUDA(cudaMemcpyAsync(labels, h->trans->label.p, sizeof(int32_t) * h->trans->n_rows,
                            labels_mem == 1 ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
  timer.stop();
  if (n_groups) *n_groups = h->trans->n_groups;
  if (n_links) *n_links = h->trans->n_links;
  return FOFB_OK;
  FOFB_CATCH(h)
}

}  // extern "C"
