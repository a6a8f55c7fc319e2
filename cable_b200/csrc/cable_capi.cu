// cable_capi.cu -- C ABI (include/cable_b200.h) over the fused cbm() kernel.
//
// A handle owns: one device arena holding every registry field as a (mp,n1,n2)
// column-major SoA block, a ring of forcing slots filled asynchronously on a side
// stream, the compute stream, and the host bindings supplied by the caller (the
// Fortran shim passes C_LOC of each derived-type member once).  There is no CPU
// path in this library: every entry point that computes requires the CUDA device.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <unordered_map>
#include <cstdint>
#include <algorithm>
#include <dlfcn.h>
#include "cbm_kernel.cuh"
#include "cbm_driver.cuh"

using namespace cbl;

// cable_fast.cu: kernel A compiled with CBL_FASTDIV=1 in its own namespace (own copy of the constant-memory config)
int cblf_launch_A(const void *devptrs, size_t devptrs_bytes, const void *cfg, size_t cfg_bytes, int mp, int i0, int i1, float dels,
                  int first, unsigned long long *warn, int *redo, int big, int lvl, int max_l1, cudaStream_t st);
void cblf_debug_dump();

// occupancy targets of the three kernel variants (tuned on B200); build-time so that only the
// variants that ship are compiled
#ifndef CBL_MINB_A
#define CBL_MINB_A 1
#endif
#ifndef CBL_MINB_B
#define CBL_MINB_B 2            // with CBL_BLOCK_B = 384: 24 warps per SM at 85 registers
#endif
#ifndef CBL_MINB_FUSED
#define CBL_MINB_FUSED 8
#endif
// threads per block of kernel A / B (A's phase barriers make its block the unit that shares instruction fetches)
#ifndef CBL_BLOCK_B
#define CBL_BLOCK_B 384         // r02 sweep (profiles/r02_kernelB_geometry_probe.txt): 384 x 2 beats 128 x 4/5/6/8, 256 x 2/3, 512 x 1 by 1-2 % of the step
#endif
// kernel A's geometry for ranges that do not fill the chip with 768-thread blocks (a shard of a strong-scaling run, a
// pipeline chunk, the remainder chain): such a launch is bound by the latency of one warp's dependent chain, not by
// issue slots, so it gets small blocks and a high register cap (no spill traffic on the chain)
#ifndef CBL_SMALL_BLOCK
#define CBL_SMALL_BLOCK 128
#endif
#ifndef CBL_SMALL_MINB
#define CBL_SMALL_MINB 3
#endif

namespace {

thread_local std::string g_err;
int fail(int code, const std::string &msg) { g_err = msg; return code; }
#define CUDA_TRY(expr)                                                                          \
  do { cudaError_t e_ = (expr);                                                                 \
       if (e_ != cudaSuccess) return fail(CABLE_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); } while (0)

// ---- registry table from the X-macro ----------------------------------------------------------
constexpr unsigned FORCING = CABLE_ROLE_FORCING, PARAM = CABLE_ROLE_PARAM, STATE = CABLE_ROLE_STATE, DIAG = CABLE_ROLE_DIAG;
template <typename T> struct dt_of;
template <> struct dt_of<float>  { static constexpr int v = CABLE_DT_F32; };
template <> struct dt_of<double> { static constexpr int v = CABLE_DT_F64; };
template <> struct dt_of<int>    { static constexpr int v = CABLE_DT_I32; };

const cable_field_info g_fields[] = {
#define CABLE_FA(T, m, ct, n1, n2, role, flags) {#T "_" #m, dt_of<ct>::v, n1, n2, role, (unsigned)(flags)},
#include "../../include/cable_b200_fields.def"
};
static_assert(sizeof(g_fields) / sizeof(g_fields[0]) == NFIELDS, "registry size");

// parameter-table slot / class of every field (-1 / 0: not tabulated), from the same predicate the kernels use
const int g_tbl_slot[] = {
#define CABLE_FA(T, m, ct, n1, n2, role, flags) (CBL_TBL_ON(T, m, ct, role, flags) ? (int)TBL_##T##_##m : -1),
#include "../../include/cable_b200_fields.def"
};
const int g_tbl_drow[] = {      // row of DevPtrs::tbl_d for the REAL(r_2) soil members, else -1
#define CABLE_FA(T, m, ct, n1, n2, role, flags) (CBL_TBLD_ON(T, m, ct, role, flags) ? CBL_TBLD_ROW(m) : -1),
#include "../../include/cable_b200_fields.def"
};
const int g_tbl_class[] = {
#define CABLE_FA(T, m, ct, n1, n2, role, flags) CBL_CLASS_##T,
#include "../../include/cable_b200_fields.def"
};

size_t elem_size(int dt) { return dt == CABLE_DT_F64 ? 8 : 4; }
size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

bool carbon_tables(int mvtype, float *rw, float *tfcl, float *tvclst) {   // cable_carbon.F90:94-150
  static const float rw13[] = {16.f, 8.7f, 12.5f, 16.f, 18.f, 7.5f, 6.1f, .84f, 10.4f, 15.1f, 9.f, 5.8f, 0.001f};
  static const float tf13[] = {0.248f, 0.345f, 0.31f, 0.42f, 0.38f, 0.35f, 0.997f, 0.95f, 2.4f, 0.73f, 2.4f, 0.55f, 0.9500f};
  static const float tv13[] = {283.f, 278.f, 278.f, 235.f, 268.f, 278.0f, 278.0f, 278.0f, 278.0f, 235.f, 278.f, 278.f, 268.f};
  static const float rw15[] = {16.f, 16.f, 18.f, 8.7f, 10.4f, 6.1f, 6.1f, 6.1f, 5.8f, 5.8f, 0.001f, 9.0f, 0.001f, 0.001f, 0.001f};
  static const float tf15[] = {0.42f, 0.248f, 0.38f, 0.345f, 2.4f, 0.997f, 0.997f, 0.997f, 0.55f, 0.55f, 0.9500f, 2.4f, 0.9500f, 0.9500f, 0.9500f};
  static const float tv15[] = {235.f, 283.f, 268.f, 278.f, 278.0f, 278.0f, 278.0f, 278.0f, 278.f, 278.f, 278.0f, 278.f, 278.f, 278.f, 268.f};
  static const float rw17[] = {16.f, 16.f, 18.f, 8.7f, 12.5f, 15.1f, 10.4f, 7.5f, 6.1f, 6.1f, 0.001f, 5.8f, 0.001f, 5.8f, 0.001f, 9.0f, 0.001f};
  static const float tf17[] = {0.42f, 0.248f, 0.38f, 0.345f, 0.31f, 0.73f, 2.4f, 0.35f, 0.997f, 0.997f, 0.9500f, 0.55f, 0.9500f, 0.55f, 0.9500f, 2.4f, 0.9500f};
  static const float tv17[] = {235.f, 283.f, 268.f, 278.f, 278.f, 235.f, 278.0f, 278.0f, 278.0f, 278.0f, 278.0f, 278.f, 278.f, 278.f, 268.f, 278.f, 278.f};
  const float *a, *b, *c; int n;
  switch (mvtype) {
    case 13: a = rw13; b = tf13; c = tv13; n = 13; break;
    case 15: a = rw15; b = tf15; c = tv15; n = 15; break;
    case 16: a = rw17; b = tf17; c = tv17; n = 16; break;   // IGBP without water bodies = first 16 of the 17 table
    case 17: a = rw17; b = tf17; c = tv17; n = 17; break;
    default: return false;
  }
  for (int i = 0; i < 17; i++) { rw[i] = i < n ? a[i] : 0.f; tfcl[i] = i < n ? b[i] : 0.f; tvclst[i] = i < n ? c[i] : 0.f; }
  return true;
}

}  // namespace

struct cable_casa_state;            // CASA-CNP daily step (casa_capi.inc)
namespace {
void casa_free(cable_handle *h);
int casa_icycle(const cable_handle *h);                   // icycle of sumcflux: 0 without CASA; -1 when cfg.icycle > 0 but no CASA state
void casa_sumcflux_ptrs(cable_handle *h, PostIn &p);
}

struct cable_handle {
  int mp = 0, device = 0;
  cable_casa_state *casa = nullptr;
  cable_cfg cfg{};
  DevCfg dcfg{};
  char *arena = nullptr; size_t arena_bytes = 0;
  size_t off[NFIELDS]{};             // arena offset of each field (forcing: offset inside a slot)
  size_t bytes[NFIELDS]{};
  size_t forcing_slot_bytes = 0, forcing_base = 0;
  size_t off_tvair_in = 0, off_oldcansto_in = 0;   // per-slot copies of the two inputs that are not FORCING rows
  bool dirty[NFIELDS]{};             // cable_b200_mark_dirty: host-side writes to resident fields, uploaded by the next cbm()
  bool any_dirty = false;
  float *d_tbl = nullptr; double *d_tbl_d = nullptr;   // per-PFT / per-soil-type parameter tables (cbm_types.cuh)
  int tbl_classes = 0, tbl_enable = 1;
  bool veg_tables_off = false;       // cable_b200_casa_feedback made veg%vcmax / ejmax per-tile on the device
  // preferred shared-memory share of the unified L1 array, per cent.  -1 = the driver's default.  The kernels hold ~3.5 KB of
  // static shared memory (parameter tables): a preference BELOW what they need (0 = all to L1, as in round 1 when they had
  // none) costs 45 % of the step on B200 (1.47 vs 1.00 ms, profiles/r02_carveout_sweep.txt); 8-15 % and the default tie.
  int carveout = -1;
  bool out_mask_on = false;          // cable_b200_set_output_mask: what cable_b200_cbm() mirrors to the host each step
  bool out_mask[NFIELDS]{};
  int nslots = 2;
  void *host[NFIELDS]{};
  bool host_pinned[NFIELDS]{};
  cudaStream_t s_compute = nullptr, s_compute2 = nullptr, s_copy = nullptr, s_d2h = nullptr;
  // The resident step as a pipeline (cable_b200_step, pipe_chunk > 0): the shard is cut into chunks of pipe_chunk tiles, chunk j
  // runs its kernels A -> B on chain stream j % S, and NOTHING joins the chains at the end of a step: a chunk of step k+1
  // depends only on the same chunk of step k (same stream), so kernel B of one chunk and the uneven tail of every launch
  // overlap kernel A of the next chunk / the next step.  Every other entry point reaches the compute stream through
  // main_stream(), which joins the chains first.
  int pipe_chunk = 0;
  int big_min = 0, big_min_pipe = 0; // ranges of at least this many tiles run kernel A as CBL_BLOCK_A-thread blocks, one per SM (launch_range)
  std::vector<cudaStream_t> s_chain;
  std::vector<cudaEvent_t> ev_chain_done, ev_chain_slot;     // [S], [nslots * S]: chain c has finished its share of the latest step / of the latest step that read forcing slot s
  bool chains_pending = false;
  bool prof_open = false; std::vector<int> prof_steps;       // profile intervals: event pair k covers prof_steps[k] steps
  cudaEvent_t ev_join_c2 = nullptr;
  int chunk_tiles = 0;
  std::vector<cudaEvent_t> ev_chunk_in, ev_chunk_done;   // pipelined cable_b200_cbm()
  int nchunks = 1;
  // CUDA graphs of the pipelined drop-in step, keyed by (slot, dels, bound host pointers)
  std::unordered_map<uint64_t, cudaGraphExec_t> graphs;
  cudaEvent_t ev_fork = nullptr, ev_join_copy = nullptr, ev_join_d2h = nullptr;
  int use_graph = 1, trace = 0;
  std::vector<cudaEvent_t> tr;     // CABLE_B200_TRACE=1: timing events of one pipelined step (debug aid)
  long long graph_launches = 0, graph_h2d = 0, graph_d2h = 0;
  std::vector<cudaEvent_t> ev_forcing_ready, ev_slot_free;
  std::vector<char> slot_has_data;
  unsigned long long *d_warn = nullptr;
  long long soil_snow_calls = 0;       // the reference's  INTEGER, SAVE :: ktau  (cbl_soilsnow_main.F90:60)
  int last_slot = 0;                   // forcing slot of the most recent step
  int fastdiv = 1;                     // kernel A's CBL_FASTDIV build first (CABLE_B200_FASTDIV=0: ordinary build only)
  int *d_redo = nullptr;               // per-block redo flags written by the fast build
  int xsw = 0;                         // any of litter / l_rev_corr / l_new_roughness_soil / soil_thermal_fix set
  int block = 128, split = 1, minb_a = CBL_MINB_A, minb_b = CBL_MINB_B, sms = 148, max_l1 = 1, step_chains = 2;
  // driver stages (cbm_driver.cuh); allocated by cable_b200_driver_init
  struct Driver {
    bool on = false;
    int nland = 0;
    int *d_cstart = nullptr, *d_cend = nullptr, *d_tile_land = nullptr;
    float *d_patchfrac = nullptr, *d_latitude = nullptr;
    std::vector<float *> d_met_land;            // one per forcing slot: [CABLE_MET_NROWS][nland]
    char *arr_block = nullptr;                  // backing store of `arr`
    DriverArrays arr{};
    std::vector<OutRow> rows; OutRow *d_rows = nullptr;
    double *d_agg = nullptr; int agg_counter = 0;
    float *d_out[2] = {nullptr, nullptr}; cudaEvent_t ev_out_free[2] = {nullptr, nullptr}; int out_buf = 0;
    cudaEvent_t ev_reduced = nullptr;
    // pipelined step: the output reduction runs per chunk on the chain streams (reduce_rows)
    std::vector<int> h_cstart;                   // host copy of landpt%cstart: land points that start in a chunk
    std::vector<int> h_cend;
    std::vector<cudaEvent_t> ev_red;             // [chunks] chunk j's share of the reduction is done
    float *d_partial = nullptr;                  // [2 staging buffers][chunks][rows]: running sums of land points that cross a chunk edge
  } drv;
  // multi-GPU gather of the output block (cable_b200_comm_init / cable_b200_output_gather_async)
  struct Comm {
    void *comm = nullptr;            // ncclComm_t
    int rank = 0, nranks = 1;
    float *d_recv = nullptr; size_t recv_floats = 0;    // root: the other ranks' blocks, rank after rank
  } comm;
  // measurement
  cable_counters ctr{};
  bool profile = false;
  std::vector<cudaEvent_t> prof_ev; size_t prof_n = 0;
};

namespace {
// The compute stream for everything except the pipelined resident step: joins the step's chain streams first, so whatever
// is enqueued next sees every chunk of every enqueued step (and, in profile mode, closes the open timing interval).
cudaStream_t main_stream(cable_handle *h) {
  if (h->chains_pending) {
    for (size_t c = 0; c < h->s_chain.size(); c++) {
      cudaEventRecord(h->ev_chain_done[c], h->s_chain[c]);
      cudaStreamWaitEvent(h->s_compute, h->ev_chain_done[c], 0);
    }
    h->chains_pending = false;
    if (h->prof_open) { cudaEventRecord(h->prof_ev[h->prof_n - 1], h->s_compute); h->prof_open = false; }
  }
  return h->s_compute;
}
// `st` must not overwrite forcing slot `slot` before its readers are done: the last step that read it (every chain of the
// pipelined step) and the post-step kernels
void wait_slot_free(cable_handle *h, cudaStream_t st, int slot) {
  cudaStreamWaitEvent(st, h->ev_slot_free[slot], 0);
  const size_t S = h->s_chain.size();
  for (size_t c = 0; c < S; c++) cudaStreamWaitEvent(st, h->ev_chain_slot[(size_t)slot * S + c], 0);
}
// Per-tile work that follows a step (post-step statements, CASA accumulation / biogeochem) rides the step pipeline: chunk j's
// share goes to chunk j's chain stream, behind that chunk's kernels, so it needs no join.  Without a pipeline: one launch
// over [0, mp) on the compute stream.  fn(i0, i1, stream) enqueues the work for tiles [i0, i1).
bool piped(const cable_handle *h) { return !h->s_chain.empty() && h->pipe_chunk > 0 && h->mp > h->pipe_chunk; }
template <class F>
int per_chunk(cable_handle *h, F fn) {
  if (!piped(h)) return fn(0, h->mp, main_stream(h));
  const int S = (int)h->s_chain.size();
  // the chains see what the compute stream has been given since they last forked from it (free when nothing was)
  CUDA_TRY(cudaEventRecord(h->ev_fork, h->s_compute));
  for (int c = 0; c < S; c++) CUDA_TRY(cudaStreamWaitEvent(h->s_chain[c], h->ev_fork, 0));
  int j = 0;
  for (int i0 = 0; i0 < h->mp; i0 += h->pipe_chunk, j++) {
    const int i1 = (i0 + h->pipe_chunk < h->mp) ? i0 + h->pipe_chunk : h->mp;
    int rc = fn(i0, i1, h->s_chain[j % S]); if (rc) return rc;
  }
  h->chains_pending = true;
  return CABLE_OK;
}
// forcing slot `slot` has a new last reader on every chain
int mark_slot_read_by_chains(cable_handle *h, int slot) {
  const size_t S = h->s_chain.size();
  for (size_t c = 0; c < S; c++) CUDA_TRY(cudaEventRecord(h->ev_chain_slot[(size_t)slot * S + c], h->s_chain[c]));
  return CABLE_OK;
}
// a timing interval (event pair) on the compute stream covering `steps` steps
int prof_begin(cable_handle *h) {
  if (h->prof_n + 2 > h->prof_ev.size()) {
    size_t old = h->prof_ev.size(); h->prof_ev.resize(old + 512);
    for (size_t k = old; k < h->prof_ev.size(); k++) cudaEventCreate(&h->prof_ev[k]);
  }
  h->prof_steps.resize(h->prof_ev.size() / 2, 1);
  h->prof_steps[h->prof_n / 2] = 0;
  h->prof_n += 2;
  return (int)h->prof_n - 2;
}
}  // namespace

namespace {

void driver_free(cable_handle *h);      // driver stages, end of this file

// what the caller sets before every CALL cbm: the FORCING rows, met%tvair/tvrad when the caller's values are to be
// honoured (met_tv_is_tk = 0), and canopy%oldcansto when the caller keeps its own `oldcansto = cansto` statement
// (caller_duties = 0, cable_serial.F90:573)
bool is_forcing_input(const cable_handle *h, int id) {
  const cable_field_info &f = g_fields[id];
  if (id == FID_met_tvair || id == FID_met_tvrad) return !h->cfg.met_tv_is_tk;
  if (id == FID_canopy_oldcansto) return !h->cfg.caller_duties;
  return f.role == FORCING && !(f.flags & CABLE_FLAG_HOSTONLY);
}

void *dev_ptr(const cable_handle *h, int id, int slot) {
  const cable_field_info &f = g_fields[id];
  if (f.flags & CABLE_FLAG_HOSTONLY) return nullptr;
  if (f.role == FORCING) return h->arena + h->forcing_base + (size_t)slot * h->forcing_slot_bytes + h->off[id];
  return h->arena + h->off[id];
}

// where a per-step input lands on the device: its forcing-slot block, or the slot's side copy for the two inputs
// whose registry row is a resident (slot-independent) field
void *dev_in_ptr(const cable_handle *h, int id, int slot) {
  char *slot_base = h->arena + h->forcing_base + (size_t)slot * h->forcing_slot_bytes;
  if (id == FID_met_tvair) return slot_base + h->off_tvair_in;
  if (id == FID_canopy_oldcansto) return slot_base + h->off_oldcansto_in;
  return dev_ptr(h, id, slot);
}

DevPtrs make_ptrs(const cable_handle *h, int slot) {
  DevPtrs d;
  int id = 0;
#define CABLE_FA(T, m, ct, n1, n2, role, flags) d.T##_##m = (ct *)dev_ptr(h, id, slot); id++;
#include "../../include/cable_b200_fields.def"
  d.tbl = h->d_tbl; d.tbl_d = h->d_tbl_d; d.tbl_classes = h->tbl_classes;
  d.met_tvair_in = (const float *)dev_in_ptr(h, FID_met_tvair, slot);
  d.canopy_oldcansto_in = (const float *)dev_in_ptr(h, FID_canopy_oldcansto, slot);
  return d;
}

// copy tiles [i0, i1) of one field: every component is a contiguous run, the components are mp elements apart
int copy_field_range(cable_handle *h, int id, int slot, bool to_device, cudaStream_t s, int i0, int i1) {
  if (!h->host[id] || i1 <= i0) return CABLE_OK;
  char *dp = (char *)(to_device && slot >= 0 ? dev_in_ptr(h, id, slot) : dev_ptr(h, id, slot < 0 ? 0 : slot));
  if (!dp) return CABLE_OK;
  const cable_field_info &f = g_fields[id];
  const size_t es = elem_size(f.dtype), pitch = (size_t)h->mp * es, width = (size_t)(i1 - i0) * es, rows = (size_t)f.n1 * f.n2;
  char *hp = (char *)h->host[id] + (size_t)i0 * es;
  dp += (size_t)i0 * es;
  if (to_device) { CUDA_TRY(cudaMemcpy2DAsync(dp, pitch, hp, pitch, width, rows, cudaMemcpyHostToDevice, s)); h->ctr.h2d_bytes += (long long)(width * rows); }
  else { CUDA_TRY(cudaMemcpy2DAsync(hp, pitch, dp, pitch, width, rows, cudaMemcpyDeviceToHost, s)); h->ctr.d2h_bytes += (long long)(width * rows); }
  return CABLE_OK;
}

int copy_field(cable_handle *h, int id, int slot, bool to_device, cudaStream_t s) {
  if (!h->host[id]) return CABLE_OK;
  void *dp = to_device && slot >= 0 ? dev_in_ptr(h, id, slot) : dev_ptr(h, id, slot < 0 ? 0 : slot);
  if (!dp) return CABLE_OK;
  if (to_device) { CUDA_TRY(cudaMemcpyAsync(dp, h->host[id], h->bytes[id], cudaMemcpyHostToDevice, s)); h->ctr.h2d_bytes += (long long)h->bytes[id]; }
  else { CUDA_TRY(cudaMemcpyAsync(h->host[id], dp, h->bytes[id], cudaMemcpyDeviceToHost, s)); h->ctr.d2h_bytes += (long long)h->bytes[id]; }
  return CABLE_OK;
}

// launch the step kernels for tiles [i0, i1) on the compute stream
int launch_range(cable_handle *h, const DevPtrs &d_in, float dels, int first, int i0, int i1, cudaStream_t st, bool in_pipeline = false) {
  if (i1 <= i0) return CABLE_OK;
  DevPtrs d = d_in;
  // kernel A (surface + canopy) then kernel B (soil/snow/carbon) on the same stream, or the fused variant.
  // CBL_MINB_x = resident blocks per SM the compiler must allow (register cap 65536 / (BLOCK*MINB)).
#define CBL_LAUNCH_X(PH, BL, MB, LV, XS) {                                                                                     \
    {                                                                                                                          \
      static bool once_[64] = {};   /* per instantiation and device */                                                        \
      const int dv_ = h->device & 63;                                                                                          \
      if (!once_[dv_]) {                                                                                                       \
        /* shared-memory share of the unified array (handle::carveout; the kernels hold 3.5 KB of static tables) */            \
        if (h->max_l1) cudaFuncSetAttribute(cbm_kernel<PH, BL, MB, LV, XS>, cudaFuncAttributePreferredSharedMemoryCarveout, h->carveout); \
        cudaGetLastError();                                                                                                    \
        once_[dv_] = true;                                                                                                     \
      }                                                                                                                        \
    }                                                                                                                          \
    cbm_kernel<PH, BL, MB, LV, XS><<<(i1 - i0 + (BL) - 1) / (BL), BL, 0, st>>>(d, h->dcfg, h->mp, i0, i1, dels, first, h->d_warn, redo_); }
#define CBL_LAUNCH(PH, BL, MB, LV) CBL_LAUNCH_X(PH, BL, MB, LV, 0)
#define CBL_DISPATCH(PH, BL, MB)                                                     \
  switch (h->cfg.output_level) { case 0: CBL_LAUNCH(PH, BL, MB, 0); break; case 1: CBL_LAUNCH(PH, BL, MB, 1); break; default: CBL_LAUNCH(PH, BL, MB, 2); break; }
  // the instantiations that carry litter / l_rev_corr / l_new_roughness_soil / soil_thermal_fix (output levels 1 and 2
  // only: level 0 runs as level 1; kernel A always as 256-thread blocks)
#define CBL_DISPATCH_X(PH, BL, MB)                                                   \
  switch (h->cfg.output_level) { case 2: CBL_LAUNCH_X(PH, BL, MB, 2, 1); break; default: CBL_LAUNCH_X(PH, BL, MB, 1, 1); break; }
  int *redo_ = nullptr;
  if (h->xsw) {
    CBL_DISPATCH_X(1, 256, 3);
    CUDA_TRY(cudaGetLastError());
    CBL_DISPATCH_X(2, CBL_BLOCK_B, CBL_MINB_B);
    h->ctr.kernel_launches++;
  } else if (h->split) {
    // kernel A: one 768-thread block per SM when the range fills the chip that way; small ranges (a shard of an
    // 8-GPU run, a pipeline chunk) take 256-thread blocks, three per SM, so that every SM still gets work
    // kernel A runs first as its CBL_FASTDIV build (cable_fast.cu: IEEE divisions / square roots without the slow-path
    // scaffolding); the ordinary build that follows only computes the blocks that build flagged (normally none)
    const bool big = i1 - i0 >= (in_pipeline ? h->big_min_pipe : h->big_min);
    if (h->fastdiv) {
      const int bl = big ? CBL_BLOCK_A : CBL_SMALL_BLOCK, nblk = (i1 - i0 + bl - 1) / bl;
      if (i0 % 256) return fail(CABLE_E_ARG, "launch_range: range start must be a multiple of 256 (redo-flag slices)");
      redo_ = h->d_redo + (size_t)(i0 / 64);               // one entry per block (>= 64 threads): disjoint slices for ranges launched concurrently
      const int rc = cblf_launch_A(&d, sizeof(d), &h->dcfg, sizeof(h->dcfg), h->mp, i0, i1, dels, first, h->d_warn, redo_, big ? 1 : 0, h->cfg.output_level, h->max_l1 ? h->carveout : -2, st);
      if (rc) return fail(CABLE_E_CUDA, std::string("fast kernel A launch: ") + cudaGetErrorString((cudaError_t)rc));
      if (getenv("CABLE_B200_FASTDIV_DEBUG")) {             // debugging aid: how many blocks the fast build handed back
        std::vector<int> fl(nblk);
        cudaStreamSynchronize(st);
        cudaMemcpy(fl.data(), redo_, nblk * sizeof(int), cudaMemcpyDeviceToHost);
        int n = 0; for (int v : fl) n += v != 0;
        fprintf(stderr, "[cable_b200] fastdiv: %d of %d blocks flagged (tiles %d..%d)\n", n, nblk, i0, i1);
        cblf_debug_dump();
      }
      h->ctr.kernel_launches++;
    }
    if (big) { CBL_DISPATCH(1, CBL_BLOCK_A, CBL_MINB_A); }
    else { CBL_DISPATCH(1, CBL_SMALL_BLOCK, CBL_SMALL_MINB); }
    redo_ = nullptr;
    CUDA_TRY(cudaGetLastError());
    CBL_DISPATCH(2, CBL_BLOCK_B, CBL_MINB_B);
    h->ctr.kernel_launches++;
  } else {
    CBL_DISPATCH(3, 128, CBL_MINB_FUSED);
  }
#undef CBL_DISPATCH_X
#undef CBL_LAUNCH_X
#undef CBL_DISPATCH
#undef CBL_LAUNCH
  CUDA_TRY(cudaGetLastError());
  h->ctr.kernel_launches++;
  return CABLE_OK;
}

bool wanted_output(const cable_handle *h, int id) {
  const cable_field_info &f = g_fields[id];
  if (f.flags & CABLE_FLAG_HOSTONLY) return false;
  const int lvl = h->cfg.output_level;
  if (h->out_mask_on) return h->out_mask[id];
  if (lvl < 1) return false;
  return (f.role == STATE) || (f.role == DIAG && (lvl >= 2 || (f.flags & CABLE_FLAG_STAR)));
}

// PARAM fields flagged OPTIN are inputs of one non-default switch: they must be bound only when it is set
bool optin_param_needed(const cable_handle *h, int id) {
  if (id == FID_veg_clitt) return h->cfg.litter != 0;
  if (id == FID_climate_qtemp_max_last_year) return h->cfg.call_climate != 0;
  if (id == FID_soil_cnsd_vec || id == FID_soil_sand_vec || id == FID_soil_watr) return h->cfg.soil_thermal_fix != 0;
  return true;
}

// soil%*_vec must be the spreads the default configuration builds (cable_parameters.F90:1685-1691);
// the device promotes the per-tile scalars instead of reading them.
int check_spreads(cable_handle *h) {
  const int mp = h->mp;
  struct { int vec, scal; } pairs[] = {{FID_soil_swilt_vec, FID_soil_swilt}, {FID_soil_sfc_vec, FID_soil_sfc}, {FID_soil_ssat_vec, FID_soil_ssat}};
  for (auto &p : pairs) {
    const double *v = (const double *)h->host[p.vec]; const float *s = (const float *)h->host[p.scal];
    if (!v || !s) continue;
    for (int k = 0; k < CABLE_MS; k++)
      for (int i = 0; i < mp; i++)
        if (v[(size_t)i + (size_t)mp * k] != (double)s[i])
          return fail(CABLE_E_PARAM, std::string(g_fields[p.vec].name) + " is not SPREAD(" + g_fields[p.scal].name + "): per-layer soil parameters are not supported");
  }
  if (const double *v = (const double *)h->host[FID_soil_zse_vec])
    for (int k = 0; k < CABLE_MS; k++)
      for (int i = 0; i < mp; i++)
        if (v[(size_t)i + (size_t)mp * k] != (double)h->cfg.zse[k]) return fail(CABLE_E_PARAM, "soil_zse_vec is not SPREAD(soil%zse)");
  if (const double *v = (const double *)h->host[FID_canopy_fes_cor])
    for (int i = 0; i < mp; i++) if (v[i] != 0.0) return fail(CABLE_E_PARAM, "canopy_fes_cor must be 0 offline (cable_serial.F90:440)");
  return CABLE_OK;
}

// Per-PFT / per-soil-type tables: a class (veg%* keyed by veg%iveg, soil%* keyed by soil%isoilm) is served from tables when
// every tabulated member of it is a pure function of the key over all tiles of this handle (bit-for-bit), which is how the
// offline driver fills them from pft_params.nml / cable_soilparm.nml.  Otherwise the class stays per-tile.
int build_param_tables(cable_handle *h) {
  h->tbl_classes = 0;
  if (!h->tbl_enable) return CABLE_OK;
  const int mp = h->mp;
  std::vector<float> tbl((size_t)TBL_COUNT * CBL_TBL_KEYS, 0.f);
  std::vector<double> tbld((size_t)2 * CBL_TBL_KEYS, 0.0);
  for (int cls = 1; cls <= 2; cls++) {
    const int *key = (const int *)h->host[cls == 1 ? FID_veg_iveg : FID_soil_isoilm];
    if (!key) continue;
    bool ok = true;
    std::vector<char> seen(CBL_TBL_KEYS);
    for (int i = 0; i < mp && ok; i++) ok = key[i] >= 0 && key[i] < CBL_TBL_KEYS;
    for (int id = 0; id < NFIELDS && ok; id++) {
      if (g_tbl_class[id] != cls || (g_tbl_slot[id] < 0 && g_tbl_drow[id] < 0)) continue;
      if (!h->host[id]) { ok = false; break; }
      const cable_field_info &f = g_fields[id];
      for (int k = 0; k < f.n1 * f.n2 && ok; k++) {
        std::fill(seen.begin(), seen.end(), 0);
        if (g_tbl_slot[id] >= 0) {
          const uint32_t *v = (const uint32_t *)h->host[id] + (size_t)k * mp;
          uint32_t *row = (uint32_t *)&tbl[(size_t)(g_tbl_slot[id] + k) * CBL_TBL_KEYS];
          for (int i = 0; i < mp; i++) {
            const int q = key[i];
            if (!seen[q]) { seen[q] = 1; row[q] = v[i]; }
            else if (row[q] != v[i]) { ok = false; break; }
          }
        } else {
          const uint64_t *v = (const uint64_t *)h->host[id];
          uint64_t *row = (uint64_t *)&tbld[(size_t)g_tbl_drow[id] * CBL_TBL_KEYS];
          for (int i = 0; i < mp; i++) {
            const int q = key[i];
            if (!seen[q]) { seen[q] = 1; row[q] = v[i]; }
            else if (row[q] != v[i]) { ok = false; break; }
          }
        }
      }
    }
    if (ok && !(cls == 1 && h->veg_tables_off)) h->tbl_classes |= cls;
  }
  if (!h->d_tbl) {
    CUDA_TRY(cudaMalloc(&h->d_tbl, tbl.size() * sizeof(float)));
    CUDA_TRY(cudaMalloc(&h->d_tbl_d, tbld.size() * sizeof(double)));
  }
  CUDA_TRY(cudaMemcpyAsync(h->d_tbl, tbl.data(), tbl.size() * sizeof(float), cudaMemcpyHostToDevice, main_stream(h)));
  CUDA_TRY(cudaMemcpyAsync(h->d_tbl_d, tbld.data(), tbld.size() * sizeof(double), cudaMemcpyHostToDevice, main_stream(h)));
  CUDA_TRY(cudaStreamSynchronize(main_stream(h)));
  return CABLE_OK;
}

// host-side writes to resident fields announced with cable_b200_mark_dirty go up before the next step reads them
int flush_dirty(cable_handle *h) {
  if (!h->any_dirty) return CABLE_OK;
  bool param = false;
  for (int id = 0; id < NFIELDS; id++) {
    if (!h->dirty[id]) continue;
    int rc = copy_field(h, id, -1, true, main_stream(h)); if (rc) return rc;
    h->dirty[id] = false;
    param = param || g_fields[id].role == PARAM;
  }
  h->any_dirty = false;
  if (param) { int rc = build_param_tables(h); if (rc) return rc; }      // a changed parameter may leave (or re-enter) its table
  return CABLE_OK;
}

}  // namespace

extern "C" {

int cable_b200_abi_version(void) { return CABLE_B200_ABI_VERSION; }
const char *cable_b200_last_error(void) { return g_err.c_str(); }
int cable_b200_nfields(void) { return NFIELDS; }

int cable_b200_field_id(const char *name) {
  if (!name) return CABLE_E_ARG;
  for (int i = 0; i < NFIELDS; i++) if (!strcmp(g_fields[i].name, name)) return i;
  return fail(CABLE_E_ARG, std::string("unknown field ") + name);
}

int cable_b200_field_info(int id, cable_field_info *out) {
  if (id < 0 || id >= NFIELDS || !out) return fail(CABLE_E_ARG, "bad field id");
  *out = g_fields[id];
  return CABLE_OK;
}

void cable_b200_default_cfg(cable_cfg *c) {
  if (!c) return;
  memset(c, 0, sizeof(*c));
  c->struct_bytes = (int)sizeof(*c);
  c->gs_switch = CABLE_GS_LEUNING;            // src/offline/cable.nml:62
  c->fwsoil_switch = CABLE_FWSOIL_STANDARD;   // cable.nml:58
  c->ssnow_potev = CABLE_POTEV_HDM;           // cable.nml:71
  c->diag_soil_resp_on = 1;                   // cable.nml:63
  c->icycle = 0;                              // cable.nml:37
  c->mvtype = 17;
  c->snmin = 1.0f;                            // cable_runtime_opts_mod.F90:9
  c->max_glacier_snowd = 1100.0f; c->snow_ccnsw = 2.0f; c->max_ssdn = 750.0f;   // cable_common.F90:217-222
  c->max_sconds = 2.51f; c->frozen_limit = 0.85f;
  c->wiltParam = 0.5f; c->satuParam = 0.8f;   // cable.nml:55-56 (only read when redistrb is set)
  const float zse[CABLE_MS] = {.022f, .058f, .154f, .409f, 1.085f, 2.872f};      // cable_parameters.F90:1241
  for (int k = 0; k < CABLE_MS; k++) c->zse[k] = zse[k];
  c->zshh[0] = 0.5f * zse[0];                                                    // cable_parameters.F90:1828-1831
  c->zshh[CABLE_MS] = 0.5f * zse[CABLE_MS - 1];
  for (int k = 1; k < CABLE_MS; k++) c->zshh[k] = 0.5f * (zse[k - 1] + zse[k]);
  c->ratecp[0] = 1.0f; c->ratecp[1] = 0.03f; c->ratecp[2] = 0.14f;               // pft_params.nml ratecp1-3
  c->ratecs[0] = 2.0f; c->ratecs[1] = 0.5f;                                      // pft_params.nml ratecs1-2
  c->met_tv_is_tk = 1; c->caller_duties = 1; c->output_level = 1; c->n_forcing_slots = 2;
  c->threads_per_block = 0;
}

int cable_b200_create(int mp, const cable_cfg *cfg, int device, cable_handle **out) {
  if (!out) return fail(CABLE_E_ARG, "out is null");
  *out = nullptr;
  if (mp <= 0 || !cfg) return fail(CABLE_E_ARG, "mp must be > 0 and cfg non-null");
  if (cfg->struct_bytes != (int)sizeof(cable_cfg)) return fail(CABLE_E_ARG, "cable_cfg size mismatch (ABI)");
  // switch combinations the device path does not implement
  if (cfg->or_evap || cfg->gw_model || cfg->soil_struc_sli || cfg->runtime_um)
    return fail(CABLE_E_UNSUPPORTED, "unsupported switch: or_evap/gw_model/soil_struc=sli/cable_runtime%um must be off");
  if (cfg->gs_switch != CABLE_GS_LEUNING && cfg->gs_switch != CABLE_GS_MEDLYN)
    return fail(CABLE_E_UNSUPPORTED, "gs_model_switch failed.");                 // cbl_dryLeaf.F90:436
  if (cfg->fwsoil_switch < 0 || cfg->fwsoil_switch > CABLE_FWSOIL_LAI_KTAUL)
    return fail(CABLE_E_UNSUPPORTED, "fwsoil_switch failed.");                   // cbl_dryLeaf.F90:179
  if (cfg->ssnow_potev != CABLE_POTEV_HDM && cfg->ssnow_potev != CABLE_POTEV_PM) return fail(CABLE_E_UNSUPPORTED, "ssnow_potev");
  if (cfg->output_level < 0 || cfg->output_level > 2 || cfg->n_forcing_slots < 1) return fail(CABLE_E_ARG, "output_level/n_forcing_slots");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(CABLE_E_NODEVICE, "no CUDA device visible: cable_b200 has no CPU fallback");
  if (device < 0) { const char *lr = getenv("LOCAL_RANK"); device = lr ? atoi(lr) % ndev : 0; }
  if (device >= ndev) return fail(CABLE_E_ARG, "device index out of range");
  CUDA_TRY(cudaSetDevice(device));

  cable_handle *h = new cable_handle();
  h->mp = mp; h->device = device; h->cfg = *cfg; h->nslots = cfg->n_forcing_slots;
  h->block = 128;
  // tuning knobs (DESIGN.md 'Kernel'): split step into kernels A/B, min resident blocks per SM of each
  if (const char *e = getenv("CABLE_B200_SPLIT")) h->split = atoi(e);
  if (const char *e = getenv("CABLE_B200_MAXL1")) h->max_l1 = atoi(e);
  if (const char *e = getenv("CABLE_B200_STEP_CHAINS")) h->step_chains = atoi(e);
  if (const char *e = getenv("CABLE_B200_FASTDIV")) h->fastdiv = atoi(e);
  if (const char *e = getenv("CABLE_B200_TABLES")) h->tbl_enable = atoi(e);
  if (const char *e = getenv("CABLE_B200_CARVEOUT")) h->carveout = atoi(e);
  // device-side config + host-evaluated constants
  DevCfg &d = h->dcfg;
  d.gs_switch = cfg->gs_switch; d.fwsoil_switch = cfg->fwsoil_switch; d.ssnow_potev = cfg->ssnow_potev;
  d.diag_soil_resp_on = cfg->diag_soil_resp_on; d.l_new_runoff_speed = cfg->l_new_runoff_speed;
  d.l_new_reduce_soilevp = cfg->l_new_reduce_soilevp; d.icycle = cfg->icycle; d.mvtype = cfg->mvtype;
  d.litter = cfg->litter != 0; d.l_rev_corr = cfg->l_rev_corr != 0; d.soil_thermal_fix = cfg->soil_thermal_fix != 0;
  d.l_new_roughness_soil = cfg->l_new_roughness_soil != 0;
  d.redistrb = cfg->redistrb != 0; d.call_climate = cfg->call_climate != 0;
  d.wiltParam = cfg->wiltParam; d.satuParam = cfg->satuParam;
  h->xsw = d.litter || d.l_rev_corr || d.soil_thermal_fix || d.l_new_roughness_soil || d.redistrb || d.call_climate;
  d.met_tv_is_tk = cfg->met_tv_is_tk; d.caller_duties = cfg->caller_duties; d.output_level = cfg->output_level;
  d.snmin = cfg->snmin; d.max_glacier_snowd = cfg->max_glacier_snowd; d.snow_ccnsw = cfg->snow_ccnsw;
  d.max_ssdn = cfg->max_ssdn; d.max_sconds = cfg->max_sconds; d.frozen_limit = cfg->frozen_limit;
  float zsetot = 0.f;
  for (int k = 0; k < CABLE_MS; k++) { d.zse[k] = cfg->zse[k]; zsetot = zsetot + cfg->zse[k]; }
  d.zsetot = zsetot;
  for (int k = 0; k <= CABLE_MS; k++) d.zshh[k] = cfg->zshh[k];
  for (int k = 0; k < CABLE_NCP; k++) d.ratecp[k] = cfg->ratecp[k];
  for (int k = 0; k < CABLE_NCS; k++) d.ratecs[k] = cfg->ratecs[k];
  const float pi180 = 3.1415927f / 180.0f, ang[3] = {15.0f, 45.0f, 75.0f};
  // correctly rounded (fp64 evaluation, one rounding): what the reference compiler's constant folding yields
  for (int b = 0; b < 3; b++) d.cos3[b] = (float)cos((double)(pi180 * ang[b]));
  d.log60 = (float)log(60.0); d.log250 = (float)log(250.0);
  d.prandt_third = (float)pow((double)0.71f, (double)(1.0f / 3.0f)); d.log_cccw = (float)log(2.0);
  if (cfg->icycle == 0 && !carbon_tables(cfg->mvtype, d.rw, d.tfcl, d.tvclst)) {
    delete h;
    return fail(CABLE_E_UNSUPPORTED, "Error! Dimension not compatible with CASA or CSIRO or IGBP types! (mvtype)");   // cable_carbon.F90:142
  }
  // arena layout: [PARAM][STATE][DIAG STAR][DIAG other][forcing slot 0..n-1]
  size_t cur = 0;
  auto place = [&](unsigned role, int star) {
    for (int id = 0; id < NFIELDS; id++) {
      const cable_field_info &f = g_fields[id];
      h->bytes[id] = (size_t)mp * f.n1 * f.n2 * elem_size(f.dtype);
      if (f.role != role || (f.flags & CABLE_FLAG_HOSTONLY)) continue;
      if (role == DIAG && star >= 0 && (int)((f.flags & CABLE_FLAG_STAR) != 0) != star) continue;
      h->off[id] = cur; cur = align_up(cur + h->bytes[id], 256);
    }
  };
  place(PARAM, -1); place(STATE, -1); place(DIAG, 1); place(DIAG, 0);
  h->forcing_base = cur;
  size_t fcur = 0;
  for (int id = 0; id < NFIELDS; id++) {
    const cable_field_info &f = g_fields[id];
    if (f.role != FORCING || (f.flags & CABLE_FLAG_HOSTONLY)) continue;
    h->off[id] = fcur; fcur = align_up(fcur + h->bytes[id], 256);
  }
  h->off_tvair_in = fcur; fcur = align_up(fcur + (size_t)mp * sizeof(float), 256);
  h->off_oldcansto_in = fcur; fcur = align_up(fcur + (size_t)mp * sizeof(float), 256);
  h->forcing_slot_bytes = fcur;
  h->arena_bytes = cur + fcur * (size_t)h->nslots;
  cudaError_t e = cudaMalloc(&h->arena, h->arena_bytes);
  if (e != cudaSuccess) { delete h; return fail(CABLE_E_CUDA, std::string("cudaMalloc arena: ") + cudaGetErrorString(e)); }
  cudaMemset(h->arena, 0, h->arena_bytes);
  cudaMalloc(&h->d_redo, ((size_t)mp / 64 + 2) * sizeof(int));
  cudaMemset(h->d_redo, 0, ((size_t)mp / 64 + 2) * sizeof(int));
  cudaMalloc(&h->d_warn, 2 * sizeof(unsigned long long));       // [0] dryLeaf soft warnings, [1] blocks recomputed after a fast-path miss
  cudaMemset(h->d_warn, 0, 2 * sizeof(unsigned long long));
  {
    int lo = 0, hi = 0;                                   // numerically lower = higher priority
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    cudaStreamCreateWithPriority(&h->s_compute, cudaStreamNonBlocking, hi);
    cudaStreamCreateWithPriority(&h->s_compute2, cudaStreamNonBlocking, lo);
  }
  cudaStreamCreateWithFlags(&h->s_copy, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&h->ev_join_c2, cudaEventDisableTiming);
  // drop-in call pipelining: tiles are cut into chunks so that forcing H2D, the kernels and the D2H of results of
  // different chunks overlap (PCIe is full duplex).  The call is bound by the D2H of the mirrored fields, one strided copy
  // per field component and chunk: TWO equal chunks keep the copies large (280 copies of ~350 KB reach 31 GB/s, 140 of
  // ~700 KB 34 GB/s, profiles/r02_dropin_probe.txt) and still hide the second chunk's kernels; shards of less than one
  // round of kernel A are not cut.  Chunks alternate between two compute streams.
  int edge = CBL_BLOCK_A; while (edge % 256) edge += CBL_BLOCK_A;
  {
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device); h->sms = sms;
    const int wave = sms * CBL_MINB_A * CBL_BLOCK_A;
    h->nchunks = mp > wave ? 2 : 1;
    if (const char *e = getenv("CABLE_B200_CHUNK_WAVES")) { const int w = atoi(e) > 0 ? atoi(e) : 1; h->nchunks = (mp + wave * w - 1) / (wave * w); }
    h->chunk_tiles = ((mp + h->nchunks - 1) / h->nchunks + edge - 1) / edge * edge;
  }
  if (const char *e = getenv("CABLE_B200_CHUNKS")) {        // explicit override: equal chunks
    h->nchunks = atoi(e) > 0 ? atoi(e) : 1;
    h->chunk_tiles = ((mp + h->nchunks - 1) / h->nchunks + edge - 1) / edge * edge;
  }
  if (h->nchunks > 64) { h->nchunks = 64; h->chunk_tiles = ((mp + 63) / 64 + edge - 1) / edge * edge; }
  // chunk edges are multiples of lcm(kernel A's block, 256): two concurrently launched ranges never share a 256-tile
  // redo-flag entry (launch_range checks it)
  if (const char *e = getenv("CABLE_B200_GRAPH")) h->use_graph = atoi(e);
  if (const char *e = getenv("CABLE_B200_TRACE")) { h->trace = atoi(e); if (h->trace) h->use_graph = 0; }
  cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->ev_join_copy, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->ev_join_d2h, cudaEventDisableTiming);
  h->ev_chunk_in.resize(h->nchunks); h->ev_chunk_done.resize(h->nchunks);
  for (int c = 0; c < h->nchunks; c++) {
    cudaEventCreateWithFlags(&h->ev_chunk_in[c], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_chunk_done[c], cudaEventDisableTiming);
  }
  h->ev_forcing_ready.resize(h->nslots); h->ev_slot_free.resize(h->nslots); h->slot_has_data.assign(h->nslots, 0);
  {
    // pipelined resident step: chunks of whole kernel-A rounds (one 640-thread block per SM) on S chain streams
    // kernel A's geometry: up to one wave of the small blocks (sms x CBL_SMALL_MINB x CBL_SMALL_BLOCK tiles) the small, high-register
    // geometry has the shorter dependent chain; well beyond it one round of big blocks (even partly filled) is faster
    // (profiles/r02_mid_probe.txt: a launch on its own ties at ~1.4 waves -- 77 500 tiles 0.43 ms either way, 57 000 tiles 0.365
    // small / 0.395 big; as the remainder chunk of the pipelined step, next to big blocks of other chunks, 60 280 tiles cost
    // 0.575 ms/step small and 0.510 big)
    h->big_min_pipe = h->sms * CBL_SMALL_MINB * CBL_SMALL_BLOCK + 1;
    h->big_min = 80000;
    if (const char *e = getenv("CABLE_B200_BIG_MIN")) h->big_min = h->big_min_pipe = atoi(e);
    int S = 4; h->pipe_chunk = h->sms * CBL_BLOCK_A;       // 4 chains: 0.945 ms/step against 1.005 unpipelined at 310 000 tiles (profiles/r02_pipe_probe.txt)
    if (const char *e = getenv("CABLE_B200_PIPE_STREAMS")) S = atoi(e);
    if (const char *e = getenv("CABLE_B200_PIPE_CHUNK")) h->pipe_chunk = atoi(e);
    { int l = CBL_BLOCK_A; while (l % 256) l += CBL_BLOCK_A;     // chunk edges: multiples of lcm(kernel A's block, the 256-tile redo-flag slice)
      h->pipe_chunk = h->pipe_chunk / l * l; }
    if (S < 1 || h->pipe_chunk <= 0 || !h->split || h->xsw) { S = 0; h->pipe_chunk = 0; }
    h->s_chain.resize(S); h->ev_chain_done.resize(S); h->ev_chain_slot.resize((size_t)S * h->nslots);
    int lo_ = 0, hi_ = 0; cudaDeviceGetStreamPriorityRange(&lo_, &hi_);
    for (int c = 0; c < S; c++) {
      cudaStreamCreateWithPriority(&h->s_chain[c], cudaStreamNonBlocking, hi_);
      cudaEventCreateWithFlags(&h->ev_chain_done[c], cudaEventDisableTiming);
    }
    for (auto &ev : h->ev_chain_slot) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  }
  for (int s = 0; s < h->nslots; s++) {
    cudaEventCreateWithFlags(&h->ev_forcing_ready[s], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_slot_free[s], cudaEventDisableTiming);
  }
  CUDA_TRY(cudaDeviceSynchronize());
  *out = h;
  return CABLE_OK;
}

int cable_b200_destroy(cable_handle *h) {
  if (!h) return CABLE_OK;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  cable_b200_comm_destroy(h);          // while the streams still exist
  casa_free(h);
  for (int id = 0; id < NFIELDS; id++) if (h->host_pinned[id]) cudaHostUnregister(h->host[id]);
  for (auto ev : h->ev_forcing_ready) cudaEventDestroy(ev);
  for (auto ev : h->ev_slot_free) cudaEventDestroy(ev);
  for (auto st : h->s_chain) cudaStreamDestroy(st);
  for (auto ev : h->ev_chain_done) cudaEventDestroy(ev);
  for (auto ev : h->ev_chain_slot) cudaEventDestroy(ev);
  for (auto ev : h->prof_ev) cudaEventDestroy(ev);
  if (h->s_compute) cudaStreamDestroy(h->s_compute);
  if (h->s_copy) cudaStreamDestroy(h->s_copy);
  if (h->s_d2h) cudaStreamDestroy(h->s_d2h);
  if (h->s_compute2) cudaStreamDestroy(h->s_compute2);
  if (h->ev_join_c2) cudaEventDestroy(h->ev_join_c2);
  for (auto &g : h->graphs) cudaGraphExecDestroy(g.second);
  if (h->ev_fork) { cudaEventDestroy(h->ev_fork); cudaEventDestroy(h->ev_join_copy); cudaEventDestroy(h->ev_join_d2h); }
  for (auto ev : h->ev_chunk_in) cudaEventDestroy(ev);
  for (auto ev : h->ev_chunk_done) cudaEventDestroy(ev);
  if (h->drv.on) driver_free(h);
  if (h->d_warn) cudaFree(h->d_warn);
  if (h->d_redo) cudaFree(h->d_redo);
  if (h->d_tbl) cudaFree(h->d_tbl);
  if (h->d_tbl_d) cudaFree(h->d_tbl_d);
  if (h->arena) cudaFree(h->arena);
  delete h;
  return CABLE_OK;
}

int cable_b200_bind_field(cable_handle *h, int id, void *host) {
  if (!h || id < 0 || id >= NFIELDS) return fail(CABLE_E_ARG, "bind_field: bad handle or id");
  cudaSetDevice(h->device);
  if (h->host_pinned[id]) { cudaHostUnregister(h->host[id]); h->host_pinned[id] = false; }
  h->host[id] = host;
  // pin per-step traffic in place so H2D/D2H are true async DMA; failure only costs speed
  const cable_field_info &f = g_fields[id];
  const bool per_step = (f.role == FORCING) || (f.role == STATE) || (f.role == DIAG && (f.flags & CABLE_FLAG_STAR));
  if (host && per_step && !(f.flags & CABLE_FLAG_HOSTONLY)) {
    cudaPointerAttributes attr{};
    const bool already = (cudaPointerGetAttributes(&attr, host) == cudaSuccess) && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (!already) {       // caller-pinned buffers (cudaHostAlloc / registered elsewhere) are used as they are
      if (cudaHostRegister(host, h->bytes[id], cudaHostRegisterDefault) == cudaSuccess) h->host_pinned[id] = true;
      else cudaGetLastError();
    }
  }
  return CABLE_OK;
}

int cable_b200_upload(cable_handle *h, unsigned role_mask) {
  if (!h) return fail(CABLE_E_ARG, "null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  if (role_mask & PARAM) { int rc = check_spreads(h); if (rc) return rc; }
  for (int id = 0; id < NFIELDS; id++) {
    const cable_field_info &f = g_fields[id];
    if (!(f.role & role_mask) || (f.flags & CABLE_FLAG_HOSTONLY)) continue;
    if (f.role == FORCING) continue;                       // forcing goes through set_forcing_async
    if (f.role == PARAM && (f.flags & CABLE_FLAG_OPTIN) && !optin_param_needed(h, id)) {
      if (h->host[id]) { int rc = copy_field(h, id, -1, true, main_stream(h)); if (rc) return rc; }
      continue;
    }
    if ((f.role & (PARAM | STATE)) && !h->host[id])
      return fail(CABLE_E_UNBOUND, std::string("field not bound: ") + f.name);
    int rc = copy_field(h, id, -1, true, main_stream(h)); if (rc) return rc;
  }
  if ((role_mask & STATE) && h->cfg.l_new_roughness_soil) {     // canopy%us feeds the next ruff_resist (cable_roughness.F90:197)
    if (!h->host[FID_canopy_us]) return fail(CABLE_E_UNBOUND, "field not bound: canopy_us (l_new_roughness_soil)");
    int rc = copy_field(h, FID_canopy_us, -1, true, main_stream(h)); if (rc) return rc;
  }
  CUDA_TRY(cudaStreamSynchronize(main_stream(h)));
  if (role_mask & PARAM) { int rc = build_param_tables(h); if (rc) return rc; }
  return CABLE_OK;
}

int cable_b200_param_table_classes(cable_handle *h) {
  if (!h) return fail(CABLE_E_ARG, "null handle");
  return h->tbl_classes;
}

int cable_b200_download(cable_handle *h, unsigned role_mask, unsigned flag_mask) {
  if (!h) return fail(CABLE_E_ARG, "null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  for (int id = 0; id < NFIELDS; id++) {
    const cable_field_info &f = g_fields[id];
    if (!(f.role & role_mask) || (f.flags & CABLE_FLAG_HOSTONLY) || f.role == FORCING) continue;
    if (f.role == DIAG && flag_mask && !(f.flags & flag_mask)) continue;
    int rc = copy_field(h, id, 0, false, main_stream(h)); if (rc) return rc;
  }
  CUDA_TRY(cudaStreamSynchronize(main_stream(h)));
  return CABLE_OK;
}

int cable_b200_mark_dirty(cable_handle *h, int id) {
  if (!h || id < 0 || id >= NFIELDS) return fail(CABLE_E_ARG, "mark_dirty: bad handle or id");
  const cable_field_info &f = g_fields[id];
  if ((f.flags & CABLE_FLAG_HOSTONLY) || f.role == FORCING || f.role == DIAG)
    return fail(CABLE_E_ARG, std::string("mark_dirty: not a resident PARAM/STATE field: ") + f.name);
  if (!h->host[id]) return fail(CABLE_E_UNBOUND, std::string("field not bound: ") + f.name);
  h->dirty[id] = true; h->any_dirty = true;
  return CABLE_OK;
}

int cable_b200_set_output_mask(cable_handle *h, const int *field_ids, int n) {
  if (!h || n < 0 || (n > 0 && !field_ids)) return fail(CABLE_E_ARG, "set_output_mask: bad argument");
  CUDA_TRY(cudaSetDevice(h->device));
  bool m[NFIELDS] = {};
  for (int k = 0; k < n; k++) {
    const int id = field_ids[k];
    if (id < 0 || id >= NFIELDS) return fail(CABLE_E_ARG, "set_output_mask: unknown field id");
    const cable_field_info &f = g_fields[id];
    if ((f.flags & CABLE_FLAG_HOSTONLY) || f.role == FORCING || f.role == PARAM)
      return fail(CABLE_E_ARG, std::string("set_output_mask: not an output of cbm: ") + f.name);
    if (f.role == DIAG && !(f.flags & CABLE_FLAG_STAR) && h->cfg.output_level < 2)
      return fail(CABLE_E_ARG, std::string("set_output_mask: ") + f.name + " is only written at output_level 2");
    if (f.role == DIAG && h->cfg.output_level < 1)
      return fail(CABLE_E_ARG, std::string("set_output_mask: ") + f.name + " needs output_level >= 1");
    if (!h->host[id]) return fail(CABLE_E_UNBOUND, std::string("field not bound: ") + f.name);
    m[id] = true;
  }
  h->out_mask_on = n > 0;
  memcpy(h->out_mask, m, sizeof(m));
  // the captured drop-in pipelines copy the previous selection
  CUDA_TRY(cudaStreamSynchronize(main_stream(h)));
  for (auto &g : h->graphs) cudaGraphExecDestroy(g.second);
  h->graphs.clear();
  return CABLE_OK;
}

int cable_b200_set_forcing_async(cable_handle *h, int slot) {
  if (!h || slot < 0 || slot >= h->nslots) return fail(CABLE_E_ARG, "set_forcing_async: bad slot");
  CUDA_TRY(cudaSetDevice(h->device));
  // do not overwrite a slot a running step still reads
  wait_slot_free(h, h->s_copy, slot);
  for (int id = 0; id < NFIELDS; id++) {
    if (!is_forcing_input(h, id)) continue;
    if (!h->host[id]) return fail(CABLE_E_UNBOUND, std::string("forcing field not bound: ") + g_fields[id].name);
    int rc = copy_field(h, id, slot, true, h->s_copy); if (rc) return rc;
  }
  CUDA_TRY(cudaEventRecord(h->ev_forcing_ready[slot], h->s_copy));
  h->slot_has_data[slot] = 1;
  return CABLE_OK;
}

int cable_b200_step(cable_handle *h, int ktau, float dels, int slot) {
  (void)ktau;   // cbm only uses ktau for a diagnostic print (ktau_gl); the soil_snow SAVE counter is ours
  if (!h || slot < 0 || slot >= h->nslots) return fail(CABLE_E_ARG, "step: bad slot");
  if (!(dels > 0.f)) return fail(CABLE_E_ARG, "step: dels must be > 0");
  CUDA_TRY(cudaSetDevice(h->device));
  { int rc = flush_dirty(h); if (rc) return rc; }
  const DevPtrs d = make_ptrs(h, slot);
  const int first = (h->soil_snow_calls == 0) ? 1 : 0;
  const int S = (int)h->s_chain.size();
  if (S > 0 && h->pipe_chunk > 0 && h->mp > h->pipe_chunk) {
    // ---- pipelined: chunk j on chain stream j % S, no join at the end of the step (see cable_handle::pipe_chunk).
    // The chains first see whatever the compute stream has been given since they last forked from it (uploads, the
    // post-step kernels of the previous step, ...): with nothing in between, that event has already fired.
    cudaStream_t sm = h->s_compute;                       // not main_stream(): the step must not join its own pipeline
    if (h->profile && !h->prof_open) {
      const int k = prof_begin(h);
      CUDA_TRY(cudaEventRecord(h->prof_ev[k], h->chains_pending ? h->s_chain[0] : sm));
      h->prof_open = true;
    }
    if (h->prof_open) h->prof_steps[h->prof_n / 2 - 1]++;
    CUDA_TRY(cudaEventRecord(h->ev_fork, sm));
    for (int c = 0; c < S; c++) {
      CUDA_TRY(cudaStreamWaitEvent(h->s_chain[c], h->ev_fork, 0));
      if (h->slot_has_data[slot]) CUDA_TRY(cudaStreamWaitEvent(h->s_chain[c], h->ev_forcing_ready[slot], 0));
    }
    int j = 0;
    for (int i0 = 0; i0 < h->mp; i0 += h->pipe_chunk, j++) {
      const int i1 = (i0 + h->pipe_chunk < h->mp) ? i0 + h->pipe_chunk : h->mp;
      int rc = launch_range(h, d, dels, first, i0, i1, h->s_chain[j % S], true); if (rc) return rc;
    }
    for (int c = 0; c < S; c++) CUDA_TRY(cudaEventRecord(h->ev_chain_slot[(size_t)slot * S + c], h->s_chain[c]));
    h->chains_pending = true;
  } else {
    cudaStream_t sm = main_stream(h);
    if (h->slot_has_data[slot]) CUDA_TRY(cudaStreamWaitEvent(sm, h->ev_forcing_ready[slot], 0));
    int k = -1;
    if (h->profile) { k = prof_begin(h); h->prof_steps[k / 2] = 1; CUDA_TRY(cudaEventRecord(h->prof_ev[k], sm)); }
    // Kernel A runs one block per SM, so a shard is walked in whole "rounds" of sms*BLOCK_A tiles and the last round
    // leaves SMs idle.  The tiles of the full rounds and the remainder therefore go out as two chains A->B on two
    // streams (the first at higher priority): kernel B of the full rounds starts while the remainder's kernel A still
    // occupies only part of the chip, and fills the idle SMs.
    const int round_tiles = h->sms * CBL_BLOCK_A;
    const int head = h->step_chains > 1 ? (h->mp / round_tiles) * round_tiles : 0;
    if (h->split && head > 0 && head < h->mp) {
      CUDA_TRY(cudaEventRecord(h->ev_fork, sm));
      CUDA_TRY(cudaStreamWaitEvent(h->s_compute2, h->ev_fork, 0));
      { int rc = launch_range(h, d, dels, first, 0, head, sm); if (rc) return rc; }
      { int rc = launch_range(h, d, dels, first, head, h->mp, h->s_compute2); if (rc) return rc; }
      CUDA_TRY(cudaEventRecord(h->ev_join_c2, h->s_compute2));
      CUDA_TRY(cudaStreamWaitEvent(sm, h->ev_join_c2, 0));
    } else {
      int rc = launch_range(h, d, dels, first, 0, h->mp, sm); if (rc) return rc;
    }
    if (k >= 0) CUDA_TRY(cudaEventRecord(h->prof_ev[k + 1], sm));
    CUDA_TRY(cudaEventRecord(h->ev_slot_free[slot], sm));
  }
  h->soil_snow_calls++;
  h->ctr.steps++;
  h->last_slot = slot;
  return CABLE_OK;
}

namespace {

// Enqueue one pipelined step over tile chunks on the three streams:
//   [H2D forcing c] (s_copy) -> [kernel A c, kernel B c] (s_compute) -> [D2H results c] (s_d2h)
// so the PCIe transfers of one chunk hide behind the arithmetic of its neighbours (PCIe is full duplex).
// Chunk edges are multiples of 128.  Used directly, or under stream capture to build a CUDA graph.
int enqueue_pipelined_step(cable_handle *h, const DevPtrs &d, float dels, int first, int slot) {
  const int nch = h->nchunks, per = h->chunk_tiles;
  const bool tr = h->trace && (h->trace++ == 4);      // the 4th pipelined call after start-up
  if (tr) { h->tr.resize(1 + 3 * nch); for (auto &e : h->tr) cudaEventCreate(&e); cudaEventRecord(h->tr[0], h->s_copy); }
  for (int c = 0; c < nch; c++) {
    const int i0 = c * per, i1 = (i0 + per < h->mp) ? i0 + per : h->mp;
    if (i0 >= i1) break;
    for (int id = 0; id < NFIELDS; id++)
      if (is_forcing_input(h, id)) { int rc = copy_field_range(h, id, slot, true, h->s_copy, i0, i1); if (rc) return rc; }
    CUDA_TRY(cudaEventRecord(h->ev_chunk_in[c], h->s_copy));
    if (tr) cudaEventRecord(h->tr[1 + 3 * c], h->s_copy);
  }
  for (int c = 0; c < nch; c++) {
    const int i0 = c * per, i1 = (i0 + per < h->mp) ? i0 + per : h->mp;
    if (i0 >= i1) break;
    cudaStream_t st = (c & 1) ? h->s_compute2 : main_stream(h);
    CUDA_TRY(cudaStreamWaitEvent(st, h->ev_chunk_in[c], 0));
    { int rc = launch_range(h, d, dels, first, i0, i1, st); if (rc) return rc; }
    CUDA_TRY(cudaEventRecord(h->ev_chunk_done[c], st));
    if (tr) cudaEventRecord(h->tr[2 + 3 * c], st);
    if (h->cfg.output_level >= 1) {
      CUDA_TRY(cudaStreamWaitEvent(h->s_d2h, h->ev_chunk_done[c], 0));
      for (int id = 0; id < NFIELDS; id++)
        if (wanted_output(h, id)) { int rc = copy_field_range(h, id, 0, false, h->s_d2h, i0, i1); if (rc) return rc; }
    }
    if (tr) cudaEventRecord(h->tr[3 + 3 * c], h->s_d2h);
  }
  if (tr) {
    cudaDeviceSynchronize();
    for (int c = 0; c < nch; c++) {
      float a = 0, b = 0, d2 = 0;
      cudaEventElapsedTime(&a, h->tr[0], h->tr[1 + 3 * c]); cudaEventElapsedTime(&b, h->tr[0], h->tr[2 + 3 * c]);
      cudaEventElapsedTime(&d2, h->tr[0], h->tr[3 + 3 * c]);
      fprintf(stderr, "[trace] chunk %d tiles [%d,%d): h2d done %.3f ms, kernels done %.3f ms, d2h done %.3f ms\n", c, c * per,
              (c * per + per < h->mp) ? c * per + per : h->mp, a, b, d2);
    }
    for (auto &e : h->tr) cudaEventDestroy(e);
    h->tr.clear();
  }
  return CABLE_OK;
}

uint64_t step_key(const cable_handle *h, int slot, float dels) {
  uint64_t k = 1469598103934665603ull;
  auto mix = [&k](uint64_t v) { k ^= v; k *= 1099511628211ull; };
  uint32_t db; memcpy(&db, &dels, 4);
  mix((uint64_t)slot); mix(db); mix((uint64_t)h->cfg.output_level); mix((uint64_t)h->nchunks); mix((uint64_t)h->split);
  for (int id = 0; id < NFIELDS; id++)
    if (is_forcing_input(h, id) || wanted_output(h, id)) mix((uint64_t)(uintptr_t)h->host[id]);
  return k;
}

}  // namespace

int cable_b200_cbm(cable_handle *h, int ktau, float dels) {
  (void)ktau;
  if (!h) return fail(CABLE_E_ARG, "null handle");
  if (!(dels > 0.f)) return fail(CABLE_E_ARG, "cbm: dels must be > 0");
  CUDA_TRY(cudaSetDevice(h->device));
  const int slot = (int)(h->ctr.steps % h->nslots);
  for (int id = 0; id < NFIELDS; id++)
    if (is_forcing_input(h, id) && !h->host[id]) return fail(CABLE_E_UNBOUND, std::string("forcing field not bound: ") + g_fields[id].name);
  const DevPtrs d = make_ptrs(h, slot);
  const int first = (h->soil_snow_calls == 0) ? 1 : 0;
  // a previous asynchronous user of this slot (set_forcing_async/step) must have drained
  CUDA_TRY(cudaStreamSynchronize(h->s_copy));
  { int rc = flush_dirty(h); if (rc) return rc; }
  CUDA_TRY(cudaStreamSynchronize(main_stream(h)));
  const long long launches0 = h->ctr.kernel_launches, h2d0 = h->ctr.h2d_bytes, d2h0 = h->ctr.d2h_bytes;
  if (h->use_graph && !first) {
    // The whole pipeline (hundreds of strided copies + kernels) is replayed as ONE graph launch: the per-call
    // enqueue cost of the copies would otherwise exceed the transfer time it is meant to hide.
    const uint64_t key = step_key(h, slot, dels);
    auto it = h->graphs.find(key);
    if (it == h->graphs.end()) {
      cudaGraph_t g = nullptr;
      CUDA_TRY(cudaStreamBeginCapture(main_stream(h), cudaStreamCaptureModeThreadLocal));
      int rc = CABLE_OK;
      cudaError_t e = cudaEventRecord(h->ev_fork, main_stream(h));
      if (e == cudaSuccess) e = cudaStreamWaitEvent(h->s_copy, h->ev_fork, 0);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(h->s_d2h, h->ev_fork, 0);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(h->s_compute2, h->ev_fork, 0);
      if (e == cudaSuccess) rc = enqueue_pipelined_step(h, d, dels, 0, slot);
      if (e == cudaSuccess && !rc) e = cudaEventRecord(h->ev_join_copy, h->s_copy);
      if (e == cudaSuccess && !rc) e = cudaEventRecord(h->ev_join_d2h, h->s_d2h);
      if (e == cudaSuccess && !rc) e = cudaEventRecord(h->ev_join_c2, h->s_compute2);
      if (e == cudaSuccess && !rc) e = cudaStreamWaitEvent(main_stream(h), h->ev_join_c2, 0);
      if (e == cudaSuccess && !rc) e = cudaStreamWaitEvent(main_stream(h), h->ev_join_copy, 0);
      if (e == cudaSuccess && !rc) e = cudaStreamWaitEvent(main_stream(h), h->ev_join_d2h, 0);
      cudaError_t e2 = cudaStreamEndCapture(main_stream(h), &g);
      if (rc) { if (g) cudaGraphDestroy(g); return rc; }
      if (e != cudaSuccess || e2 != cudaSuccess) { if (g) cudaGraphDestroy(g); return fail(CABLE_E_CUDA, std::string("graph capture: ") + cudaGetErrorString(e != cudaSuccess ? e : e2)); }
      cudaGraphExec_t ge = nullptr;
      e = cudaGraphInstantiate(&ge, g, 0);
      cudaGraphDestroy(g);
      if (e != cudaSuccess) return fail(CABLE_E_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
      if (h->graphs.size() >= 64) { for (auto &x : h->graphs) cudaGraphExecDestroy(x.second); h->graphs.clear(); }
      it = h->graphs.emplace(key, ge).first;
      // capture only recorded the work: remember what one replay moves / launches
      h->graph_launches = h->ctr.kernel_launches - launches0;
      h->graph_h2d = h->ctr.h2d_bytes - h2d0; h->graph_d2h = h->ctr.d2h_bytes - d2h0;
    }
    h->ctr.kernel_launches = launches0 + h->graph_launches;
    h->ctr.h2d_bytes = h2d0 + h->graph_h2d; h->ctr.d2h_bytes = d2h0 + h->graph_d2h;
    CUDA_TRY(cudaGraphLaunch(it->second, main_stream(h)));
  } else {
    int rc = enqueue_pipelined_step(h, d, dels, first, slot); if (rc) return rc;
  }
  h->slot_has_data[slot] = 0;
  h->soil_snow_calls++;
  h->ctr.steps++;
  h->last_slot = slot;
  CUDA_TRY(cudaStreamSynchronize(main_stream(h)));
  CUDA_TRY(cudaStreamSynchronize(h->s_compute2));
  CUDA_TRY(cudaStreamSynchronize(h->s_d2h));
  CUDA_TRY(cudaStreamSynchronize(h->s_copy));
  return CABLE_OK;
}

int cable_b200_sync(cable_handle *h) {
  if (!h) return fail(CABLE_E_ARG, "null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaStreamSynchronize(h->s_copy));
  CUDA_TRY(cudaStreamSynchronize(main_stream(h)));
  CUDA_TRY(cudaStreamSynchronize(h->s_d2h));
  return CABLE_OK;
}

void *cable_b200_device_ptr(cable_handle *h, int id, int slot) {
  if (!h || id < 0 || id >= NFIELDS || slot < 0 || slot >= h->nslots) return nullptr;
  return dev_ptr(h, id, slot);
}

void *cable_b200_compute_stream(cable_handle *h) { return h ? (void *)main_stream(h) : nullptr; }   // joins the step pipeline first

int cable_b200_profile(cable_handle *h, int enable) {
  if (!h) return fail(CABLE_E_ARG, "null handle");
  h->profile = enable != 0;
  return CABLE_OK;
}

int cable_b200_get_counters(cable_handle *h, cable_counters *out) {
  if (!h || !out) return fail(CABLE_E_ARG, "null");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaStreamSynchronize(main_stream(h)));
  for (size_t k = 0; k + 1 < h->prof_n; k += 2) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->prof_ev[k], h->prof_ev[k + 1]) == cudaSuccess) { h->ctr.kernel_ms += ms; h->ctr.kernel_ms_count += h->prof_steps[k / 2]; }
  }
  h->prof_n = 0;
  unsigned long long w = 0;
  CUDA_TRY(cudaMemcpy(&w, h->d_warn, sizeof(w), cudaMemcpyDeviceToHost));
  h->ctr.n_dryleaf_warn = (long long)w;
  CUDA_TRY(cudaMemcpy(&w, h->d_warn + 1, sizeof(w), cudaMemcpyDeviceToHost));
  h->ctr.n_fastdiv_redo_blocks = (long long)w;
  *out = h->ctr;
  return CABLE_OK;
}

int cable_b200_reset_counters(cable_handle *h) {
  if (!h) return fail(CABLE_E_ARG, "null");
  cudaSetDevice(h->device);
  cudaStreamSynchronize(main_stream(h));
  const long long steps = h->ctr.steps;       // steps also drives the forcing-slot rotation: keep it
  h->ctr = cable_counters{}; h->ctr.steps = steps;
  h->prof_n = 0;
  cudaMemset(h->d_warn, 0, 2 * sizeof(unsigned long long));
  return CABLE_OK;
}

int cable_b200_grid_reduce(cable_handle *h, int id, int comp, const float *d_patchfrac, const int *d_cstart,
                           const int *d_cend, int nland, float *d_out) {
  if (!h || id < 0 || id >= NFIELDS || nland <= 0) return fail(CABLE_E_ARG, "grid_reduce: bad argument");
  const cable_field_info &f = g_fields[id];
  if (f.dtype != CABLE_DT_F32 || comp < 0 || comp >= f.n1 * f.n2 || (f.flags & CABLE_FLAG_HOSTONLY) || f.role == FORCING)
    return fail(CABLE_E_ARG, "grid_reduce: field must be a resident f32 field");
  CUDA_TRY(cudaSetDevice(h->device));
  const float *x = (const float *)dev_ptr(h, id, 0) + (size_t)comp * h->mp;
  grid_reduce_kernel<<<(nland + 127) / 128, 128, 0, main_stream(h)>>>(x, d_patchfrac, d_cstart, d_cend, nland, d_out);
  CUDA_TRY(cudaGetLastError());
  h->ctr.kernel_launches++;
  return CABLE_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// Driver stages either side of cbm() (include/cable_b200.h, second half; kernels in cbm_driver.cuh)
// ---------------------------------------------------------------------------------------------------------------
namespace {

struct DrvName { const char *name; size_t off; bool f64; };
#define DRV(n) {#n, offsetof(DriverArrays, n), false}
const DrvName g_drv_names[] = {
  {"canopy_tscrn_max_daily", offsetof(DriverArrays, tscrn_max_daily), false},
  {"canopy_tscrn_min_daily", offsetof(DriverArrays, tscrn_min_daily), false},
  {"sum_flux_sumpn", offsetof(DriverArrays, sumpn), false}, {"sum_flux_sumrp", offsetof(DriverArrays, sumrp), false},
  {"sum_flux_sumrpw", offsetof(DriverArrays, sumrpw), false}, {"sum_flux_sumrpr", offsetof(DriverArrays, sumrpr), false},
  {"sum_flux_sumrs", offsetof(DriverArrays, sumrs), false}, {"sum_flux_sumrd", offsetof(DriverArrays, sumrd), false},
  {"sum_flux_dsumpn", offsetof(DriverArrays, dsumpn), false}, {"sum_flux_dsumrp", offsetof(DriverArrays, dsumrp), false},
  {"sum_flux_dsumrd", offsetof(DriverArrays, dsumrd), false},
  {"bal_owb", offsetof(DriverArrays, owb), true},
  {"bal_wbal", offsetof(DriverArrays, wbal), false}, {"bal_wbal_tot", offsetof(DriverArrays, wbal_tot), false},
  {"bal_precip_tot", offsetof(DriverArrays, precip_tot), false}, {"bal_rnoff_tot", offsetof(DriverArrays, rnoff_tot), false},
  {"bal_evap_tot", offsetof(DriverArrays, evap_tot), false},
  {"bal_Radbal", offsetof(DriverArrays, radbal), false}, {"bal_EbalSoil", offsetof(DriverArrays, ebalsoil), false},
  {"bal_Ebalveg", offsetof(DriverArrays, ebalveg), false}, {"bal_ebal", offsetof(DriverArrays, ebal), false},
  {"bal_ebal_tot", offsetof(DriverArrays, ebal_tot), false}, {"bal_Radbalsum", offsetof(DriverArrays, radbalsum), false},
};
#undef DRV
constexpr int N_DRV = (int)(sizeof(g_drv_names) / sizeof(g_drv_names[0]));

void *drv_ptr(const cable_handle *h, int k) { return *(void *const *)((const char *)&h->drv.arr + g_drv_names[k].off); }

void driver_free(cable_handle *h) {
  auto &v = h->drv;
  cudaFree(v.d_cstart); cudaFree(v.d_cend); cudaFree(v.d_tile_land); cudaFree(v.d_patchfrac); cudaFree(v.d_latitude);
  for (auto p : v.d_met_land) cudaFree(p);
  cudaFree(v.arr_block); cudaFree(v.d_rows); cudaFree(v.d_agg); cudaFree(v.d_out[0]); cudaFree(v.d_out[1]);
  for (int b = 0; b < 2; b++) if (v.ev_out_free[b]) cudaEventDestroy(v.ev_out_free[b]);
  if (v.ev_reduced) cudaEventDestroy(v.ev_reduced);
  for (auto ev : v.ev_red) cudaEventDestroy(ev);
  cudaFree(v.d_partial);
  v = cable_handle::Driver{};
}

}  // namespace

extern "C" {

int cable_b200_driver_field_id(const char *name) {
  if (!name) return -1;
  for (int k = 0; k < N_DRV; k++) if (!strcmp(name, g_drv_names[k].name)) return k;
  return -1;
}

int cable_b200_driver_init(cable_handle *h, int nland, const int *cstart, const int *cend, const float *patchfrac,
                           const float *latitude) {
  if (!h || nland <= 0 || !cstart || !cend || !patchfrac || !latitude) return fail(CABLE_E_ARG, "driver_init: null/empty argument");
  CUDA_TRY(cudaSetDevice(h->device));
  // the land points must tile [0, mp) in order: landpt(l)%cstart..%cend (cable_input.F90:158-160)
  std::vector<int> tile_land(h->mp, -1);
  int expect = 0;
  for (int l = 0; l < nland; l++) {
    if (cstart[l] != expect || cend[l] < cstart[l] || cend[l] >= h->mp) return fail(CABLE_E_ARG, "driver_init: cstart/cend must partition [0, mp) in order");
    for (int i = cstart[l]; i <= cend[l]; i++) tile_land[i] = l;
    expect = cend[l] + 1;
  }
  if (expect != h->mp) return fail(CABLE_E_ARG, "driver_init: land points do not cover all mp tiles");
  if (h->drv.on) driver_free(h);
  auto &v = h->drv;
  v.nland = nland;
  const size_t mp = (size_t)h->mp;
  CUDA_TRY(cudaMalloc(&v.d_cstart, nland * sizeof(int))); CUDA_TRY(cudaMalloc(&v.d_cend, nland * sizeof(int)));
  CUDA_TRY(cudaMalloc(&v.d_tile_land, mp * sizeof(int)));
  CUDA_TRY(cudaMalloc(&v.d_patchfrac, mp * sizeof(float))); CUDA_TRY(cudaMalloc(&v.d_latitude, mp * sizeof(float)));
  CUDA_TRY(cudaMemcpy(v.d_cstart, cstart, nland * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(v.d_cend, cend, nland * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(v.d_tile_land, tile_land.data(), mp * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(v.d_patchfrac, patchfrac, mp * sizeof(float), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(v.d_latitude, latitude, mp * sizeof(float), cudaMemcpyHostToDevice));
  v.d_met_land.assign(h->nslots, nullptr);
  for (int s = 0; s < h->nslots; s++) CUDA_TRY(cudaMalloc(&v.d_met_land[s], (size_t)CABLE_MET_NROWS * nland * sizeof(float)));
  // driver-owned per-tile arrays: one block, 8 bytes per element so the double array fits anywhere
  CUDA_TRY(cudaMalloc(&v.arr_block, (size_t)N_DRV * mp * 8));
  CUDA_TRY(cudaMemset(v.arr_block, 0, (size_t)N_DRV * mp * 8));
  for (int k = 0; k < N_DRV; k++) *(void **)((char *)&v.arr + g_drv_names[k].off) = v.arr_block + (size_t)k * mp * 8;
  // aggregator initial values: max -> -huge, min -> +huge (aggregator.F90:1015-1119)
  {
    std::vector<float> hi(mp, 3.402823466e38f), lo(mp, -3.402823466e38f);
    CUDA_TRY(cudaMemcpy(v.arr.tscrn_min_daily, hi.data(), mp * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(v.arr.tscrn_max_daily, lo.data(), mp * sizeof(float), cudaMemcpyHostToDevice));
  }
  for (int b = 0; b < 2; b++) CUDA_TRY(cudaEventCreateWithFlags(&v.ev_out_free[b], cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&v.ev_reduced, cudaEventDisableTiming));
  v.h_cstart.assign(cstart, cstart + nland); v.h_cend.assign(cend, cend + nland);
  if (piped(h)) {
    v.ev_red.resize((h->mp + h->pipe_chunk - 1) / h->pipe_chunk);
    for (auto &ev : v.ev_red) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  }
  v.on = true;
  return CABLE_OK;
}

int cable_b200_set_met_async(cable_handle *h, int slot, const float *met_land, const cable_met_convert *cv) {
  if (!h || !h->drv.on) return fail(CABLE_E_ARG, "set_met_async: call cable_b200_driver_init first");
  if (slot < 0 || slot >= h->nslots || !met_land || !cv) return fail(CABLE_E_ARG, "set_met_async: bad argument");
  CUDA_TRY(cudaSetDevice(h->device));
  auto &v = h->drv;
  wait_slot_free(h, h->s_copy, slot);     // a running step may still read the slot
  const size_t bytes = (size_t)CABLE_MET_NROWS * v.nland * sizeof(float);
  CUDA_TRY(cudaMemcpyAsync(v.d_met_land[slot], met_land, bytes, cudaMemcpyHostToDevice, h->s_copy));
  h->ctr.h2d_bytes += (long long)bytes;
  MetConvert c{cv->tair_offset, cv->psurf_scale, cv->rainf_scale, cv->co2_scale, cv->snowf_from_tair,
               (float)sin((double)(23.45f * (3.1415927f / 180.0f)))};
  MetOut o{(float *)dev_ptr(h, FID_met_fsd, slot), (float *)dev_ptr(h, FID_met_tk, slot), (float *)dev_ptr(h, FID_met_pmb, slot),
           (float *)dev_ptr(h, FID_met_qv, slot), (float *)dev_ptr(h, FID_met_ua, slot), (float *)dev_ptr(h, FID_met_precip, slot),
           (float *)dev_ptr(h, FID_met_precip_sn, slot), (float *)dev_ptr(h, FID_met_fld, slot), (float *)dev_ptr(h, FID_met_ca, slot),
           (float *)dev_ptr(h, FID_met_coszen, slot), (float *)dev_ptr(h, FID_met_doy, slot)};
  met_expand_kernel<<<(h->mp + 255) / 256, 256, 0, h->s_copy>>>(v.d_met_land[slot], v.nland, v.d_tile_land, v.d_latitude, c, o, h->mp);
  CUDA_TRY(cudaGetLastError());
  h->ctr.kernel_launches++;
  CUDA_TRY(cudaEventRecord(h->ev_forcing_ready[slot], h->s_copy));
  h->slot_has_data[slot] = 1;
  return CABLE_OK;
}

int cable_b200_upload_lai(cable_handle *h) {
  if (!h) return fail(CABLE_E_ARG, "null handle");
  if (!h->host[FID_veg_vlai]) return fail(CABLE_E_UNBOUND, "forcing field not bound: veg_vlai");
  CUDA_TRY(cudaSetDevice(h->device));
  for (int s = 0; s < h->nslots; s++) {
    wait_slot_free(h, h->s_copy, s);
    int rc = copy_field(h, FID_veg_vlai, s, true, h->s_copy); if (rc) return rc;
  }
  CUDA_TRY(cudaStreamSynchronize(h->s_copy));
  return CABLE_OK;
}

int cable_b200_post_step(cable_handle *h, int ktau, int kstart, float dels, int do_mass_bal, int do_energy_bal) {
  if (!h || !h->drv.on) return fail(CABLE_E_ARG, "post_step: call cable_b200_driver_init first");
  const int icycle = casa_icycle(h);
  if (icycle < 0) return fail(CABLE_E_ARG, "post_step: icycle > 0 needs the CASA-CNP state (cable_b200_casa_init) for sumcflux");
  if (h->ctr.steps <= 0) return fail(CABLE_E_ARG, "post_step: no step has run");
  CUDA_TRY(cudaSetDevice(h->device));
  // met%* live in the forcing slot of the step just enqueued; dev_ptr ignores the slot for resident fields
  auto F = [&](int id) { return (float *)dev_ptr(h, id, h->last_slot); };
  PostIn p{};
  p.smelt = F(FID_ssnow_smelt); p.rnof1 = F(FID_ssnow_rnof1); p.rnof2 = F(FID_ssnow_rnof2); p.runoff = F(FID_ssnow_runoff);
  p.tscrn = F(FID_canopy_tscrn); p.fpn = F(FID_canopy_fpn); p.frday = F(FID_canopy_frday); p.frp = F(FID_canopy_frp);
  p.fnpp = F(FID_canopy_fnpp); p.fgpp = F(FID_canopy_fgpp); p.fra = F(FID_canopy_fra);
  p.icycle = icycle;
  if (icycle > 0) casa_sumcflux_ptrs(h, p);
  p.frpw = F(FID_canopy_frpw); p.frpr = F(FID_canopy_frpr); p.frs = F(FID_canopy_frs); p.fnee = F(FID_canopy_fnee);
  p.precip = F(FID_met_precip); p.delwc = F(FID_canopy_delwc); p.snowd = F(FID_ssnow_snowd); p.osnowd = F(FID_ssnow_osnowd);
  p.fevw = F(FID_canopy_fevw); p.fev = F(FID_canopy_fev); p.cls = F(FID_ssnow_cls); p.rlam = F(FID_air_rlam);
  p.fsd = F(FID_met_fsd); p.fld = F(FID_met_fld); p.albedo = F(FID_rad_albedo); p.transd = F(FID_rad_transd);
  p.otss = F(FID_ssnow_otss); p.tv = F(FID_canopy_tv); p.fnv = F(FID_canopy_fnv); p.fns = F(FID_canopy_fns);
  p.fhs = F(FID_canopy_fhs); p.ga = F(FID_canopy_ga); p.fhv = F(FID_canopy_fhv); p.fh = F(FID_canopy_fh);
  p.qcan = F(FID_rad_qcan); p.qssabs = F(FID_rad_qssabs); p.flws = F(FID_rad_flws);
  p.wbtot = (const double *)dev_ptr(h, FID_ssnow_wbtot, 0); p.fevc = (const double *)dev_ptr(h, FID_canopy_fevc, 0);
  p.fes = (const double *)dev_ptr(h, FID_canopy_fes, 0);
  const DriverArrays arr = h->drv.arr;
  const int mp = h->mp;
  { int rc = per_chunk(h, [&](int i0, int i1, cudaStream_t st) {
      post_step_kernel<<<(i1 - i0 + 255) / 256, 256, 0, st>>>(p, arr, mp, i0, i1, ktau, kstart, dels, do_mass_bal, do_energy_bal);
      CUDA_TRY(cudaGetLastError());
      h->ctr.kernel_launches++;
      return (int)CABLE_OK; });
    if (rc) return rc; }
  // this kernel is the slot's last reader (met%precip/fsd/fld): a prefetch into the slot must wait for it, not only for the step
  if (piped(h)) return mark_slot_read_by_chains(h, h->last_slot);
  CUDA_TRY(cudaEventRecord(h->ev_slot_free[h->last_slot], main_stream(h)));
  return CABLE_OK;
}

int cable_b200_output_plan(cable_handle *h, int nrows, const int *field_id, const int *comp, const int *method,
                           const float *scale, const float *div, const float *offset) {
  if (!h || !h->drv.on) return fail(CABLE_E_ARG, "output_plan: call cable_b200_driver_init first");
  if (nrows <= 0 || !field_id || !comp || !method) return fail(CABLE_E_ARG, "output_plan: bad argument");
  CUDA_TRY(cudaSetDevice(h->device));
  auto &v = h->drv;
  std::vector<OutRow> rows(nrows);
  for (int r = 0; r < nrows; r++) {
    OutRow &o = rows[r];
    if (method[r] < CABLE_AGG_POINT || method[r] > CABLE_AGG_MAX) return fail(CABLE_E_ARG, "output_plan: unknown aggregation method");
    o.method = method[r]; o.scale = scale ? scale[r] : 1.0f; o.div = div ? div[r] : 1.0f; o.offset = offset ? offset[r] : 0.0f;
    if (field_id[r] >= 0) {
      if (field_id[r] >= NFIELDS) return fail(CABLE_E_ARG, "output_plan: unknown field");
      const cable_field_info &f = g_fields[field_id[r]];
      if ((f.flags & CABLE_FLAG_HOSTONLY) || f.role == FORCING || comp[r] < 0 || comp[r] >= f.n1 * f.n2)
        return fail(CABLE_E_ARG, std::string("output_plan: not a resident field component: ") + f.name);
      if (f.role == DIAG && !(f.flags & CABLE_FLAG_STAR) && h->cfg.output_level < 2)
        return fail(CABLE_E_ARG, std::string("output_plan: ") + f.name + " is only written at output_level 2");
      o.dtype = f.dtype;
      o.src = (const char *)dev_ptr(h, field_id[r], 0) + (size_t)comp[r] * h->mp * elem_size(f.dtype);
    } else {
      const int k = -field_id[r] - 1;
      if (k >= N_DRV || comp[r] != 0) return fail(CABLE_E_ARG, "output_plan: unknown driver array");
      o.dtype = g_drv_names[k].f64 ? CABLE_DT_F64 : CABLE_DT_F32;
      o.src = drv_ptr(h, k);
    }
  }
  CUDA_TRY(cudaStreamSynchronize(main_stream(h))); CUDA_TRY(cudaStreamSynchronize(h->s_d2h));
  cudaFree(v.d_rows); cudaFree(v.d_agg); cudaFree(v.d_out[0]); cudaFree(v.d_out[1]);
  v.d_rows = nullptr; v.d_agg = nullptr; v.d_out[0] = v.d_out[1] = nullptr;
  CUDA_TRY(cudaMalloc(&v.d_rows, nrows * sizeof(OutRow)));
  CUDA_TRY(cudaMemcpy(v.d_rows, rows.data(), nrows * sizeof(OutRow), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMalloc(&v.d_agg, (size_t)nrows * h->mp * sizeof(double)));
  for (int b = 0; b < 2; b++) CUDA_TRY(cudaMalloc(&v.d_out[b], (size_t)nrows * v.nland * sizeof(float)));
  cudaFree(v.d_partial); v.d_partial = nullptr;
  if (!v.ev_red.empty()) CUDA_TRY(cudaMalloc(&v.d_partial, (size_t)2 * v.ev_red.size() * nrows * sizeof(float)));
  v.rows = rows; v.agg_counter = 0; v.out_buf = 0;
  // 'point' rows have no reset value (aggregator.F90 never resets them) and the accumulate pass reads every row before
  // it overwrites: give them a defined first value (compute-sanitizer initcheck, tools/gpu_sanitize.sh)
  CUDA_TRY(cudaMemsetAsync(v.d_agg, 0, (size_t)nrows * h->mp * sizeof(double), main_stream(h)));
  aggregate_reset_kernel<<<(h->mp + 255) / 256, 256, 0, main_stream(h)>>>(v.d_rows, nrows, v.d_agg, h->mp);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(main_stream(h)));
  return CABLE_OK;
}

int cable_b200_output_accumulate(cable_handle *h) {
  if (!h || !h->drv.on || h->drv.rows.empty()) return fail(CABLE_E_ARG, "output_accumulate: no output plan");
  CUDA_TRY(cudaSetDevice(h->device));
  auto &v = h->drv;
  aggregate_kernel<<<(h->mp + 255) / 256, 256, 0, main_stream(h)>>>(v.d_rows, (int)v.rows.size(), v.d_agg, h->mp, v.agg_counter);
  CUDA_TRY(cudaGetLastError());
  v.agg_counter++;
  h->ctr.kernel_launches++;
  return CABLE_OK;
}

}  // extern "C"

namespace {

// end of an output interval on this rank: reduce every row patch -> grid cell into staging buffer `b` (compute stream), reset
// the aggregators, and make the D2H stream wait for it.  The caller then moves the block (D2H, or NCCL + D2H).
int reduce_rows(cable_handle *h, int &b_out) {
  auto &v = h->drv;
  const int nrows = (int)v.rows.size(), b = v.out_buf;
  if (piped(h) && v.agg_counter == 0 && !v.ev_red.empty()) {
    // One sample per interval on a pipelined step: chunk j's share of the reduction runs on chunk j's chain, behind that
    // chunk's kernels -- no join.  Land points are reduced where they START; one whose tiles cross into chunk j+1 hands its
    // running sum forward (output_reduce_kernel), so the only coupling is chunk j+1 waiting for chunk j's (tiny) launch.
    const int S = (int)h->s_chain.size(), nch = (int)v.ev_red.size();
    CUDA_TRY(cudaEventRecord(h->ev_fork, h->s_compute));
    for (int c = 0; c < S; c++) CUDA_TRY(cudaStreamWaitEvent(h->s_chain[c], h->ev_fork, 0));
    float *partial = v.d_partial + (size_t)b * nch * nrows;
    for (int j = 0; j < nch; j++) {
      cudaStream_t st = h->s_chain[j % S];
      const int i0 = j * h->pipe_chunk, i1 = (i0 + h->pipe_chunk < h->mp) ? i0 + h->pipe_chunk : h->mp;
      int l0 = (int)(std::lower_bound(v.h_cstart.begin(), v.h_cstart.end(), i0) - v.h_cstart.begin());
      const int l1 = (int)(std::lower_bound(v.h_cstart.begin(), v.h_cstart.end(), i1) - v.h_cstart.begin());
      const bool carried_in = l0 > 0 && v.h_cend[l0 - 1] >= i0;        // the land point before l0 started in chunk j-1 and ends here
      if (carried_in) { l0--; CUDA_TRY(cudaStreamWaitEvent(st, v.ev_red[j - 1], 0)); }
      CUDA_TRY(cudaStreamWaitEvent(st, v.ev_out_free[b], 0));          // the D2H that last used this staging buffer (and its partial sums) must have drained
      if (l1 > l0) {
        const dim3 grid((l1 - l0 + 127) / 128, nrows);
        output_reduce_kernel<<<grid, 128, 0, st>>>(v.d_rows, nrows, v.d_agg, 0, v.d_patchfrac, v.d_cstart, v.d_cend, v.nland, h->mp,
                                                   v.d_out[b], l0, l1, i0, i1, j > 0 ? partial + (size_t)(j - 1) * nrows : nullptr,
                                                   partial + (size_t)j * nrows);
        CUDA_TRY(cudaGetLastError());
        h->ctr.kernel_launches++;
      }
      CUDA_TRY(cudaEventRecord(v.ev_red[j], st));
      CUDA_TRY(cudaStreamWaitEvent(h->s_d2h, v.ev_red[j], 0));
    }
    h->chains_pending = true;
    b_out = b;
    v.out_buf ^= 1;
    return CABLE_OK;
  }
  // the D2H that last used this staging buffer must have drained
  CUDA_TRY(cudaStreamWaitEvent(main_stream(h), v.ev_out_free[b], 0));
  const dim3 grid((v.nland + 127) / 128, nrows);
  output_reduce_kernel<<<grid, 128, 0, main_stream(h)>>>(v.d_rows, nrows, v.d_agg, v.agg_counter > 0 ? 1 : 0, v.d_patchfrac,
                                                       v.d_cstart, v.d_cend, v.nland, h->mp, v.d_out[b], 0, v.nland, 0, h->mp, nullptr, nullptr);
  CUDA_TRY(cudaGetLastError());
  h->ctr.kernel_launches++;
  if (v.agg_counter > 0) {
    aggregate_reset_kernel<<<(h->mp + 255) / 256, 256, 0, main_stream(h)>>>(v.d_rows, nrows, v.d_agg, h->mp);
    CUDA_TRY(cudaGetLastError());
    h->ctr.kernel_launches++;
    v.agg_counter = 0;
  }
  CUDA_TRY(cudaEventRecord(v.ev_reduced, main_stream(h)));
  CUDA_TRY(cudaStreamWaitEvent(h->s_d2h, v.ev_reduced, 0));
  b_out = b;
  v.out_buf ^= 1;
  return CABLE_OK;
}

// ---- NCCL, bound at run time (dlopen): the library has no link-time dependency on it, a single-GPU caller never loads it,
// and a process that already carries a libnccl.so.2 (e.g. torch's) shares that copy -------------------------------------
struct ncclUniqueId_ { char internal[128]; };
typedef void *ncclComm_t_;
struct Nccl {
  void *lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId_ *) = nullptr;
  int (*CommInitRank)(ncclComm_t_ *, int, ncclUniqueId_, int) = nullptr;
  int (*CommDestroy)(ncclComm_t_) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void *, size_t, int, int, ncclComm_t_, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, ncclComm_t_, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
} g_nccl;
constexpr int NCCL_FLOAT32 = 7;       // ncclFloat32 (nccl.h: ncclInt8 = 0 ... ncclFloat16 = 6, ncclFloat32 = 7)

int nccl_load() {
  if (g_nccl.lib) return CABLE_OK;
  void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return fail(CABLE_E_UNSUPPORTED, std::string("NCCL not found: ") + dlerror());
#define NCCL_SYM(field, name) *(void **)(&g_nccl.field) = dlsym(lib, name); if (!g_nccl.field) return fail(CABLE_E_UNSUPPORTED, "NCCL symbol missing: " name)
  NCCL_SYM(GetUniqueId, "ncclGetUniqueId"); NCCL_SYM(CommInitRank, "ncclCommInitRank"); NCCL_SYM(CommDestroy, "ncclCommDestroy");
  NCCL_SYM(GroupStart, "ncclGroupStart"); NCCL_SYM(GroupEnd, "ncclGroupEnd"); NCCL_SYM(Send, "ncclSend"); NCCL_SYM(Recv, "ncclRecv");
  NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef NCCL_SYM
  g_nccl.lib = lib;
  return CABLE_OK;
}
#define NCCL_TRY(expr) do { int r_ = (expr); if (r_ != 0) return fail(CABLE_E_CUDA, std::string(#expr) + ": " + g_nccl.GetErrorString(r_)); } while (0)

}  // namespace

extern "C" {

int cable_b200_output_fetch_async(cable_handle *h, float *host_out) {
  if (!h || !h->drv.on || h->drv.rows.empty() || !host_out) return fail(CABLE_E_ARG, "output_fetch_async: no output plan / null buffer");
  CUDA_TRY(cudaSetDevice(h->device));
  auto &v = h->drv;
  int b = 0;
  { int rc = reduce_rows(h, b); if (rc) return rc; }
  const size_t bytes = (size_t)v.rows.size() * v.nland * sizeof(float);
  CUDA_TRY(cudaMemcpyAsync(host_out, v.d_out[b], bytes, cudaMemcpyDeviceToHost, h->s_d2h));
  CUDA_TRY(cudaEventRecord(v.ev_out_free[b], h->s_d2h));
  h->ctr.d2h_bytes += (long long)bytes;
  return CABLE_OK;
}

int cable_b200_comm_unique_id(void *id128) {
  if (!id128) return fail(CABLE_E_ARG, "comm_unique_id: null buffer");
  { int rc = nccl_load(); if (rc) return rc; }
  NCCL_TRY(g_nccl.GetUniqueId((ncclUniqueId_ *)id128));
  return CABLE_OK;
}

int cable_b200_comm_init(cable_handle *h, const void *id128, int rank, int nranks) {
  if (!h || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(CABLE_E_ARG, "comm_init: bad argument");
  { int rc = nccl_load(); if (rc) return rc; }
  CUDA_TRY(cudaSetDevice(h->device));
  cable_b200_comm_destroy(h);
  ncclUniqueId_ id; memcpy(&id, id128, sizeof(id));
  NCCL_TRY(g_nccl.CommInitRank(&h->comm.comm, nranks, id, rank));
  h->comm.rank = rank; h->comm.nranks = nranks;
  return CABLE_OK;
}

int cable_b200_comm_destroy(cable_handle *h) {
  if (!h) return CABLE_OK;
  if (h->comm.comm) { cudaSetDevice(h->device); cudaStreamSynchronize(h->s_d2h); g_nccl.CommDestroy(h->comm.comm); h->comm.comm = nullptr; }
  if (h->comm.d_recv) { cudaFree(h->comm.d_recv); h->comm.d_recv = nullptr; h->comm.recv_floats = 0; }
  h->comm.nranks = 1; h->comm.rank = 0;
  return CABLE_OK;
}

int cable_b200_output_gather_async(cable_handle *h, int root, float *host_out_root, const int *nland_of_rank) {
  if (!h || !h->drv.on || h->drv.rows.empty() || !nland_of_rank) return fail(CABLE_E_ARG, "output_gather_async: no output plan / null argument");
  auto &v = h->drv; auto &c = h->comm;
  if (c.nranks > 1 && !c.comm) return fail(CABLE_E_ARG, "output_gather_async: call cable_b200_comm_init first");
  if (root < 0 || root >= c.nranks || nland_of_rank[c.rank] != v.nland) return fail(CABLE_E_ARG, "output_gather_async: root / nland_of_rank do not match this rank");
  if (c.rank == root && !host_out_root) return fail(CABLE_E_ARG, "output_gather_async: the root needs a host buffer");
  CUDA_TRY(cudaSetDevice(h->device));
  const int nrows = (int)v.rows.size();
  long long total = 0, others = 0;
  for (int r = 0; r < c.nranks; r++) { if (nland_of_rank[r] < 0) return fail(CABLE_E_ARG, "output_gather_async: negative count"); total += nland_of_rank[r]; if (r != root) others += nland_of_rank[r]; }
  int b = 0;
  { int rc = reduce_rows(h, b); if (rc) return rc; }
  if (c.rank == root) {
    const size_t need = (size_t)others * nrows;
    if (need > c.recv_floats) {
      CUDA_TRY(cudaStreamSynchronize(h->s_d2h));
      if (c.d_recv) cudaFree(c.d_recv);
      CUDA_TRY(cudaMalloc(&c.d_recv, need * sizeof(float)));
      c.recv_floats = need;
    }
    // uneven blocks (land-point counts differ by <= 1 under master_decomp): one grouped receive per peer
    if (c.nranks > 1) {
      NCCL_TRY(g_nccl.GroupStart());
      size_t off = 0;
      for (int r = 0; r < c.nranks; r++) {
        if (r == root) continue;
        const size_t n = (size_t)nland_of_rank[r] * nrows;
        if (n) NCCL_TRY(g_nccl.Recv(c.d_recv + off, n, NCCL_FLOAT32, r, c.comm, h->s_d2h));
        off += n;
      }
      NCCL_TRY(g_nccl.GroupEnd());
    }
    // every block lands in its columns of the [nrows][total] host array (rank order = land-point order)
    size_t off = 0; long long col = 0;
    for (int r = 0; r < c.nranks; r++) {
      const float *src = (r == root) ? v.d_out[b] : c.d_recv + off;
      const size_t w = (size_t)nland_of_rank[r] * sizeof(float);
      if (w) CUDA_TRY(cudaMemcpy2DAsync(host_out_root + col, (size_t)total * sizeof(float), src, w, w, nrows, cudaMemcpyDeviceToHost, h->s_d2h));
      if (r != root) off += (size_t)nland_of_rank[r] * nrows;
      col += nland_of_rank[r];
    }
    h->ctr.d2h_bytes += (long long)total * nrows * (long long)sizeof(float);
  } else {
    const size_t n = (size_t)v.nland * nrows;
    NCCL_TRY(g_nccl.GroupStart());
    if (n) NCCL_TRY(g_nccl.Send(v.d_out[b], n, NCCL_FLOAT32, root, c.comm, h->s_d2h));
    NCCL_TRY(g_nccl.GroupEnd());
  }
  CUDA_TRY(cudaEventRecord(v.ev_out_free[b], h->s_d2h));
  return CABLE_OK;
}

int cable_b200_output_wait(cable_handle *h) {
  if (!h) return fail(CABLE_E_ARG, "null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaStreamSynchronize(h->s_d2h));
  return CABLE_OK;
}

int cable_b200_driver_download(cable_handle *h, const char *name, void *host) {
  if (!h || !h->drv.on || !host) return fail(CABLE_E_ARG, "driver_download: bad argument");
  const int k = cable_b200_driver_field_id(name);
  if (k < 0) return fail(CABLE_E_ARG, std::string("driver_download: unknown driver array ") + (name ? name : "(null)"));
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaStreamSynchronize(main_stream(h)));
  CUDA_TRY(cudaMemcpy(host, drv_ptr(h, k), (size_t)h->mp * (g_drv_names[k].f64 ? 8 : 4), cudaMemcpyDeviceToHost));
  return CABLE_OK;
}

}  // extern "C"

#include "casa_capi.inc"
