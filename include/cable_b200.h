/* cable_b200.h -- C ABI of the B200-native cbm() land-surface step.
 *
 * This is the drop-in boundary for ONE reference interface:
 *
 *   SUBROUTINE cbm( ktau, dels, air, bgc, canopy, met, bal, rad, rough, soil,
 *                   ssnow, sum_flux, veg, climate, xk, c1, rhoch )
 *   -- reference: src/offline/cbl_model_driver_offline.F90:38-40
 *      callers:   src/offline/cable_serial.F90:594,
 *                 src/offline/cable_mpiworker.F90:503
 *
 * The reference passes eleven derived types of POINTER arrays dimensioned
 * (mp[,k[,b]]).  Those are not C-interoperable, so the Fortran shim
 * (fortran/cable_cbm_b200.F90) binds every member array
 * once with cable_b200_bind_field(handle, id, C_LOC(array)) and then calls
 * cable_b200_cbm(handle, ktau, dels) each timestep.  Field ids, dtypes and
 * extents come from include/cable_b200_fields.def (mirror of
 * src/offline/cable_define_types.F90:79-717).
 *
 * All entry points return 0 on success, a negative CABLE_E_* code otherwise;
 * cable_b200_last_error() gives the text.  The reference has no status
 * returns -- it STOPs (cbl_dryLeaf.F90:179,436, cable_carbon.F90:148) -- so
 * the shim turns a non-zero status into  CALL cable_abort(...).
 *
 * No torch types, no C++ types: plain pointers, ints and floats only.
 * There is no CPU fallback: without a CUDA device create() fails.
 */
#ifndef CABLE_B200_H
#define CABLE_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define CABLE_B200_ABI_VERSION 3

/* dims fixed by the reference (cable_define_types.F90:62-71) */
#define CABLE_MS    6   /* soil layers            */
#define CABLE_MSN   3   /* snow layers            */
#define CABLE_MF    2   /* sunlit / shaded leaves */
#define CABLE_NRB   3   /* radiation bands        */
#define CABLE_SWB   2   /* shortwave bands        */
#define CABLE_NITER 4   /* stability iterations   */
#define CABLE_NCP   3   /* plant carbon pools     */
#define CABLE_NCS   2   /* soil carbon pools      */

/* field roles / flags (bit masks), see cable_b200_fields.def */
#define CABLE_ROLE_FORCING 1u
#define CABLE_ROLE_PARAM   2u
#define CABLE_ROLE_STATE   4u
#define CABLE_ROLE_DIAG    8u
#define CABLE_ROLE_ALL     15u
#define CABLE_FLAG_STAR     1u
#define CABLE_FLAG_COND     2u
#define CABLE_FLAG_HOSTONLY 4u
#define CABLE_FLAG_OPTIN    8u
#define CABLE_FLAG_XCH     16u   /* device routing: kernel A -> kernel B exchange   */
#define CABLE_FLAG_PHB     32u   /* device routing: DIAG finalised by kernel B      */
#define CABLE_FLAG_STA     64u   /* device routing: STATE modified by kernel A      */

#define CABLE_DT_F32 0
#define CABLE_DT_F64 1
#define CABLE_DT_I32 2

/* status codes */
#define CABLE_OK              0
#define CABLE_E_ARG          -1   /* bad argument / unknown field          */
#define CABLE_E_UNSUPPORTED  -2   /* switch combination not on the device  */
#define CABLE_E_CUDA         -3   /* CUDA runtime error                    */
#define CABLE_E_UNBOUND      -4   /* a required field has no host binding  */
#define CABLE_E_PARAM        -5   /* parameter consistency check failed    */
#define CABLE_E_NODEVICE     -6   /* no CUDA device: there is no fallback  */

/* enumerations for cable_user switches (src/util/cable_runtime_opts_mod.F90:14-125) */
#define CABLE_GS_LEUNING 0        /* cable_user%gs_switch = 'leuning' */
#define CABLE_GS_MEDLYN  1        /*                      = 'medlyn'  */
#define CABLE_FWSOIL_STANDARD   0 /* fwsoil_switch = 'standard'                 */
#define CABLE_FWSOIL_NONLINEAR  1 /*               = 'non-linear extrapolation' */
#define CABLE_FWSOIL_LAI_KTAUL  2 /*               = 'Lai and Ktaul 2000'       */
#define CABLE_FWSOIL_HAVERD2013 3 /* unsupported (needs SLI)                    */
#define CABLE_POTEV_HDM 0         /* ssnow_potev = 'HDM' or ''  */
#define CABLE_POTEV_PM  1         /*             = 'P-M'        */

/* Module-level state the reference cbm() reads implicitly:
 * cable_user%*, cable_runtime%*, icycle, snow constants, soil%zse ...      */
typedef struct cable_cfg {
  int   struct_bytes;          /* = sizeof(cable_cfg), ABI check             */
  /* cable_user switches */
  int   gs_switch;             /* CABLE_GS_*                                 */
  int   fwsoil_switch;         /* CABLE_FWSOIL_*                             */
  int   ssnow_potev;           /* CABLE_POTEV_*                              */
  int   diag_soil_resp_on;     /* 0 <=> DIAG_SOIL_RESP=='off' (cable_carbon.F90:256) */
  int   l_new_runoff_speed;    /* cbl_smoisturev.F90:112                     */
  int   l_new_reduce_soilevp;  /* cbl_latent_heat.F90:205                    */
  int   litter, or_evap, gw_model, l_rev_corr, soil_thermal_fix,
        l_new_roughness_soil, call_climate, redistrb, soil_struc_sli;
                               /* litter, l_rev_corr, soil_thermal_fix, l_new_roughness_soil, call_climate, redistrb:
                                  supported (second kernel instantiation); or_evap, gw_model, soil_struc_sli must be
                                  0: CABLE_E_UNSUPPORTED otherwise (dead or SLI-only in this reference snapshot) */
  /* cable_runtime, casadimension */
  int   runtime_um;            /* must be 0 (offline path)                   */
  int   icycle;                /* 0: simple carbon inside cbm (cbm:214)      */
  int   mvtype;                /* 13,15,16,17 (cable_carbon.F90:94)          */
  /* snow / soil tunables (src/util/cable_common.F90:217-222, runtime_opts:9) */
  float snmin;
  float max_glacier_snowd;
  float snow_ccnsw;
  float max_ssdn;
  float max_sconds;
  float frozen_limit;
  float wiltParam, satuParam;  /* hydraulic_redistribution limits (cable_runtime_opts_mod.F90:6-7; cable.nml:55-56) */
  /* non-per-tile members of the derived types */
  float zse[CABLE_MS];         /* soil%zse                                   */
  float zshh[CABLE_MS + 1];    /* soil%zshh                                  */
  float ratecp[CABLE_NCP];     /* bgc%ratecp                                 */
  float ratecs[CABLE_NCS];     /* bgc%ratecs                                 */
  /* boundary behaviour */
  int   met_tv_is_tk;          /* 1: met%tvair=met%tvrad=met%tk on entry, as every
                                  offline caller sets them (cable_input.F90:2679-2680,
                                  cable_mpiworker.F90:487-488); 0: upload both   */
  int   caller_duties;         /* 1: device does canopy%oldcansto=canopy%cansto
                                  before the step (cable_serial.F90:573)        */
  int   output_level;          /* 0: state only; 1: + STAR diagnostics;
                                  2: every DIAG field (parity / debugging)      */
  int   n_forcing_slots;       /* device forcing ring, >= 1 (2 = double buffer) */
  int   threads_per_block;     /* 0 = library default                           */
} cable_cfg;

typedef struct cable_field_info {
  const char *name;            /* "<type>_<member>", e.g. "ssnow_tgg"        */
  int   dtype;                 /* CABLE_DT_*                                 */
  int   n1, n2;                /* trailing extents: array is (mp,n1,n2)      */
  unsigned role;               /* CABLE_ROLE_*                               */
  unsigned flags;              /* CABLE_FLAG_*                               */
} cable_field_info;

typedef struct cable_counters {
  long long steps;             /* cbm steps run                              */
  long long kernel_launches;   /* our kernels launched                       */
  long long h2d_bytes;         /* total host->device bytes                   */
  long long d2h_bytes;         /* total device->host bytes                   */
  double    kernel_ms;         /* summed event time of profiled launches     */
  long long kernel_ms_count;   /* launches in kernel_ms                      */
  long long n_dryleaf_warn;    /* tiles that hit the 'oldevapfbl not right'
                                  soft failure (cbl_dryLeaf.F90:630)         */
  long long n_fastdiv_redo_blocks; /* kernel-A blocks recomputed by the ordinary
                                  build after the fast IEEE divide / sqrt paths
                                  met an operand outside their window        */
} cable_counters;

typedef struct cable_handle cable_handle;

int          cable_b200_abi_version(void);
const char  *cable_b200_last_error(void);

/* registry */
int          cable_b200_nfields(void);
int          cable_b200_field_id(const char *name);           /* <0 if unknown */
int          cable_b200_field_info(int id, cable_field_info *out);

/* configuration: defaults = shipped src/offline/cable.nml (leuning, standard,
 * HDM, icycle 0, snmin 1) with zse of cable_parameters.F90:1241             */
void         cable_b200_default_cfg(cable_cfg *cfg);

/* life cycle.  device < 0 => use LOCAL_RANK env var, else device 0.         */
int          cable_b200_create(int mp, const cable_cfg *cfg, int device,
                               cable_handle **out);
int          cable_b200_destroy(cable_handle *h);

/* bind a caller-owned column-major host array to a field (replaces passing
 * the derived type).  The pointer must stay valid until destroy/rebind.     */
int          cable_b200_bind_field(cable_handle *h, int field_id, void *host);

/* bulk transfers of every bound field whose role is in role_mask.
 * download() also filters on flags: a DIAG field is copied when
 * (flags & flag_mask) != 0 or flag_mask == 0.                                */
int          cable_b200_upload(cable_handle *h, unsigned role_mask);
int          cable_b200_download(cable_handle *h, unsigned role_mask,
                                 unsigned flag_mask);

/* Which parameter classes the kernels serve from per-type tables staged in shared memory instead of the per-tile
 * arrays: bit 0 = veg%* (every member a pure function of veg%iveg over this handle's tiles, as init_veg_from_vegin
 * fills them, cable_parameters.F90:3277), bit 1 = soil%* by soil%isoilm.  Decided at cable_b200_upload(PARAM) and again
 * after cable_b200_mark_dirty of a parameter; a class that fails the check is read per tile as the interface says.  */
int          cable_b200_param_table_classes(cable_handle *h);

/* A host-side write to a resident field (PARAM or STATE) between two steps -- a restart read, a parameter the
 * driver changes mid-run, casa feedback into veg%vcmax ... -- is announced per field; the next cable_b200_cbm() /
 * cable_b200_step() uploads the bound array before it runs.  (The reference cbm reads the host arrays every call;
 * the device copy cannot see a write it is not told about.)                                                       */
int          cable_b200_mark_dirty(cable_handle *h, int field_id);

/* SURVEY.md 8b sync_outputs(handle, mask): restrict what cable_b200_cbm() mirrors back to the bound host arrays
 * every step to these fields (STATE or DIAG; non-STAR DIAG fields need output_level 2).  n = 0 restores the
 * default selection of cfg.output_level.  Prognostic state left out of the mask stays current on the device and
 * comes back with cable_b200_download(h, CABLE_ROLE_STATE, 0) whenever the caller wants it (restart, end of run). */
int          cable_b200_set_output_mask(cable_handle *h, const int *field_ids, int n);

/* forcing: pack the bound FORCING arrays into pinned memory and start an
 * asynchronous H2D copy into ring slot `slot` on the side stream.            */
int          cable_b200_set_forcing_async(cable_handle *h, int slot);

/* one timestep for all mp tiles from forcing slot `slot`; asynchronous (waits for that slot's upload event).
 * Shards larger than one round of kernel A run as a PIPELINE: the tiles are cut into chunks, each chunk's kernels go to one of
 * four internal chain streams, and consecutive steps are not joined -- a chunk of step k+1 only waits for the same chunk of
 * step k.  The per-tile calls that follow a step (cable_b200_post_step, cable_b200_bgcdriver, cable_b200_casa_feedback, the
 * output reduction of cable_b200_output_fetch_async / _gather_async) ride the same chains.  Every other entry point
 * (download, sync, cbm, upload, mark_dirty, device-side access through cable_b200_compute_stream ...) first joins the chains,
 * so the host never observes a partially stepped shard.  CABLE_B200_PIPE_STREAMS=0 in the environment switches the
 * pipeline off (one launch chain per step, joined).                                                                       */
int          cable_b200_step(cable_handle *h, int ktau, float dels, int slot);

/* The drop-in call: exactly what CALL cbm(ktau, dels, ...) does as seen from
 * the host -- upload forcing from the bound arrays, run the step, bring back
 * the outputs selected by cfg.output_level into the bound arrays, and
 * synchronise.                                                              */
int          cable_b200_cbm(cable_handle *h, int ktau, float dels);

int          cable_b200_sync(cable_handle *h);

/* device-side access (for device-resident drivers and for torch/NCCL plumbing
 * via the pointer; no torch types cross this boundary)                       */
void        *cable_b200_device_ptr(cable_handle *h, int field_id, int slot);
void        *cable_b200_compute_stream(cable_handle *h);      /* cudaStream_t; joins the step pipeline first: work enqueued on it sees every step issued so far */

/* measurement */
int          cable_b200_profile(cable_handle *h, int enable); /* event-time each launch */
int          cable_b200_get_counters(cable_handle *h, cable_counters *out);
int          cable_b200_reset_counters(cable_handle *h);

/* patch -> grid-cell area-weighted reduction of one field on the device
 * (reference: src/util/cable_grid_reductions.F90:49-75): out[l] =
 * sum_{i in cstart[l]..cend[l]} x[i,comp]*patchfrac[i].  Device pointers in,
 * device pointer out; used before the per-output-interval NCCL gather.       */
int          cable_b200_grid_reduce(cable_handle *h, int field_id, int comp,
                                    const float *d_patchfrac,
                                    const int *d_cstart, const int *d_cend,
                                    int nland, float *d_out);

/* ---------------------------------------------------------------------------
 * Driver stages either side of cbm(), kept on the device.
 * They replace, for a caller that wants them, the per-tile host loops of the
 * offline drivers; cbm()'s own interface above is unchanged.
 *
 *   before the step  get_met_data's tile expansion + unit conversion + sinbet
 *                    (src/offline/cable_input.F90:1880-1883, 2139-2213,
 *                    2666-2680; src/science/radiation/cbl_sinbet.F90:12-28)
 *   after the step   cable_serial.F90:602-608 (runoff*dels, daily tscrn max/min),
 *                    sumcflux (src/science/casa-cnp/casa_sumcflux.F90:76-102),
 *                    mass_balance / energy_balance (src/offline/cable_checks.F90:
 *                    472-618), the output module's time aggregators
 *                    (src/util/aggregator.F90:585-1172) and patch -> grid-cell
 *                    reduction (src/util/cable_grid_reductions.F90:49-75).
 * ------------------------------------------------------------------------- */

/* rows of the per-land-point forcing block: float met_land[CABLE_MET_NROWS][nland],
 * in the units of the met file (ALMA): SWdown W/m2, Tair K or degC, Qair kg/kg,
 * PSurf Pa|hPa|kPa, Wind m/s, Rainf and Snowf kg/m2/s or mm/h, LWdown W/m2,
 * CO2air ppm, then local hour of day and day of year of the land point
 * (met%hod, met%doy -- the calendar itself stays on the host).               */
#define CABLE_MET_SWDOWN 0
#define CABLE_MET_TAIR   1
#define CABLE_MET_QAIR   2
#define CABLE_MET_PSURF  3
#define CABLE_MET_WIND   4
#define CABLE_MET_RAINF  5
#define CABLE_MET_SNOWF  6
#define CABLE_MET_LWDOWN 7
#define CABLE_MET_CO2    8
#define CABLE_MET_HOD    9
#define CABLE_MET_DOY    10
#define CABLE_MET_NROWS  11

typedef struct cable_met_convert {   /* cable_input.F90:1053-1209: convert%*     */
  float tair_offset;                 /* 0 (K) or 273.16 (degC)                   */
  float psurf_scale;                 /* 0.01 (Pa), 1 (hPa), 10 (kPa)             */
  float rainf_scale;                 /* dels (kg/m2/s) or dels/3600 (mm/h)       */
  float co2_scale;                   /* 1e-6: ppm -> mol/mol                     */
  int   snowf_from_tair;             /* 1: file has no (or an all-zero) Snowf:
                                        precip_sn = precip where tk <= tfrz      */
} cable_met_convert;

/* aggregation methods of the output module (aggregator.F90) */
#define CABLE_AGG_POINT 0
#define CABLE_AGG_MEAN  1
#define CABLE_AGG_SUM   2
#define CABLE_AGG_MIN   3
#define CABLE_AGG_MAX   4

/* Decomposition of this shard: tiles of land point l are cstart[l]..cend[l]
 * (0-based, inclusive; landpt(l)%cstart-1 / %cend-1), patchfrac = patch(:)%frac,
 * latitude = rad%latitude (per tile).  Allocates the driver-owned arrays
 * (bal%*, sum_flux%*, canopy%tscrn_{max,min}_daily).                          */
int          cable_b200_driver_init(cable_handle *h, int nland, const int *cstart,
                                    const int *cend, const float *patchfrac,
                                    const float *latitude);

/* Asynchronous H2D of one time slice of per-land-point forcing (pinned host
 * memory recommended) into ring slot `slot`, expanded to the tiles' met%* and
 * coszen on the device; replaces set_forcing_async for that slot.  veg%vlai
 * is not part of the slice: cable_b200_upload_lai copies the bound veg%vlai
 * to every slot (monthly, cable_input.F90:2377).                             */
int          cable_b200_set_met_async(cable_handle *h, int slot, const float *met_land,
                                      const cable_met_convert *cv);
int          cable_b200_upload_lai(cable_handle *h);

/* The statements between CALL cbm and the output module, for the step just
 * enqueued (stream ordered after it).  ktau/kstart as in cable_serial;
 * do_mass_bal / do_energy_bal = check%mass_bal / check%energy_bal.            */
int          cable_b200_post_step(cable_handle *h, int ktau, int kstart, float dels,
                                  int do_mass_bal, int do_energy_bal);

/* Output plan: nrows rows, each one component of a field sampled as
 * scale*x/div + offset and aggregated in time with `method`.  field_id >= 0 is
 * a registry field; field_id = -(1+k) is driver array k (cable_b200_driver_field_id). */
int          cable_b200_output_plan(cable_handle *h, int nrows, const int *field_id,
                                    const int *comp, const int *method,
                                    const float *scale, const float *div,
                                    const float *offset);
int          cable_b200_driver_field_id(const char *name);   /* k >= 0, or < 0 */
/* sample the plan's sources once (aggregator%accumulate of every row) */
int          cable_b200_output_accumulate(cable_handle *h);
/* end of an output interval: reduce every row patch -> grid cell on the device,
 * start the D2H of the [nrows][nland] float block into host_out (async, own
 * stream, double buffered) and reset the aggregators.  With exactly one sample
 * per interval (output%averaging='all') call it WITHOUT output_accumulate: the
 * rows are sampled and reduced in one pass.                                    */
int          cable_b200_output_fetch_async(cable_handle *h, float *host_out);
int          cable_b200_output_wait(cable_handle *h);
/* ---------------------------------------------------------------------------
 * Multi-GPU: land points shard across ranks as contiguous blocks (master_decomp, src/offline/cable_mpimaster.F90:
 * 1428-1463), one process and one handle per GPU, nothing is exchanged inside a step.  Once per output interval the
 * grid-cell output block is gathered to one rank over NCCL -- this replaces the reference workers' per-step MPI_Send of
 * ~225 fields (src/offline/cable_mpiworker.F90:552) and the master's matching receives (cable_mpimaster.F90:8066-8072).
 * NCCL is bound at run time (dlopen of libnccl.so.2): single-GPU callers never load it.
 *
 *   rank 0:   cable_b200_comm_unique_id(id)      then the driver's own MPI_Bcast(id, 128, MPI_BYTE, 0, comm)
 *   all:      cable_b200_comm_init(h, id, rank, nranks)
 *   all:      cable_b200_output_gather_async(h, root, host_out, nland_of_rank)   instead of output_fetch_async
 *   root:     cable_b200_output_wait(h);  host_out is [nrows][sum(nland_of_rank)], blocks in rank order
 * ------------------------------------------------------------------------- */
#define CABLE_B200_COMM_ID_BYTES 128
int          cable_b200_comm_unique_id(void *id128);
int          cable_b200_comm_init(cable_handle *h, const void *id128, int rank, int nranks);
int          cable_b200_comm_destroy(cable_handle *h);
/* end of an output interval on every rank: reduce the plan's rows patch -> grid cell on the device, then send the
 * [nrows][nland_of_rank[rank]] block to `root`, which receives every block and copies them (asynchronously, D2H stream)
 * into host_out_root.  nland_of_rank[r] = land points of rank r; host_out_root may be NULL on the other ranks.  With a
 * single rank (no comm_init) it is output_fetch_async.                                                             */
int          cable_b200_output_gather_async(cable_handle *h, int root, float *host_out_root, const int *nland_of_rank);

/* copy a driver array (float, or double for "bal_owb") to the host; synchronous */
int          cable_b200_driver_download(cable_handle *h, const char *name, void *host);

/* ---------------------------------------------------------------------------
 * CASA-CNP daily biogeochemistry (BASELINE config 5) on the same resident tile layout.
 * Replaces, for a caller that runs with icycle > 0,
 *     CALL bgcdriver(ktau, kstart, kend, dels, met, ssnow, canopy, veg, soil, climate, casabiome, casapool, casaflux,
 *                    casamet, casabal, phen, pop, spinConv, spinup, ktauday, idoy, loy, dump_read, dump_write, LALLOC)
 * (src/science/casa-cnp/bgcdriver.F90:7; call site src/offline/cable_serial.F90:621-629): every step it accumulates the day's
 * casamet%tairk / tsoil / moist and casaflux%meangpp / meanrleaf from the device-resident met%tk, ssnow%tgg, ssnow%wb,
 * canopy%fpn, canopy%frday; at the end of a day it runs biogeochem (biogeochem_casa.F90:7) for every tile.
 * Members of casa_biome / casa_pool / casa_flux / casa_met / casa_balance / phen_variable are bound by name
 * "<type>_<member>" (registry: include/cable_b200_casa_fields.def, generated from the reference's own type definitions);
 * casabiome / phen%TKshed rows are per vegetation type (mvtype), casabiome%xkplab/xkpsorb/xkpocc per soil order (12).
 * "soil_silt" / "soil_clay" (REAL (mp)) are bound through the same call.  Unsupported switches are rejected at init:
 * CALL_POP, LALLOC = 2, cable_user%SRF, PHENOLOGY_SWITCH = 'climate', l_landuse.                                      */
typedef struct cable_casa_cfg {
  int struct_bytes;
  int icycle;                  /* 1: C, 2: C+N, 3: C+N+P (casadimension)                        */
  int lalloc;                  /* LALLOC of bgcdriver: 0 fixed, 1 dynamic, 3 LA:SA               */
  int call_climate;            /* cable_user%call_climate (casa_rplant acclimation branch)      */
  int l_limit_labile;          /* cable_user%l_limit_labile                                      */
  int mvtype;                  /* rows of the casabiome tables                                   */
  int call_pop, srf, phenology_climate, l_landuse;   /* must be 0                                */
} cable_casa_cfg;
void         cable_b200_casa_default_cfg(cable_casa_cfg *cfg);
int          cable_b200_casa_nfields(void);
int          cable_b200_casa_field_id(const char *name);
int          cable_b200_casa_field_info(int id, cable_field_info *out, int *key);   /* key: 0 tile, 1 veg type, 2 soil order */
int          cable_b200_casa_init(cable_handle *h, const cable_casa_cfg *cfg);
int          cable_b200_casa_bind(cable_handle *h, const char *name, void *host);
int          cable_b200_casa_upload(cable_handle *h);      /* every bound array H2D                  */
int          cable_b200_casa_download(cable_handle *h);    /* every bound per-tile array D2H          */
int          cable_b200_bgcdriver(cable_handle *h, int ktau, int kstart, int kend, float dels, int ktauday, int idoy, int loy);
int          cable_b200_casa_biogeochem(cable_handle *h, int idoy);   /* biogeochem alone (bgcdriver's dump_read branch) */
/* IF (l_vcmaxFeedbk) CALL casa_feedback(ktau, veg, casabiome, casapool, casamet)  (src/science/casa-cnp/casa_feedback.F90:37);
 * IF (l_laiFeedbk .AND. icycle > 0) veg%vlai = casamet%glai   (src/offline/cable_serial.F90:587-590) -- on the device, for the
 * step that will read forcing slot `slot` (after that slot's upload, before cable_b200_step).  vcmax_walker2014: cable_user%vcmax
 * = 'Walker2014' instead of 'standard'. */
int          cable_b200_casa_feedback(cable_handle *h, int slot, int l_vcmaxfeedbk, int l_laifeedbk, int vcmax_walker2014);

#ifdef __cplusplus
}
#endif
#endif /* CABLE_B200_H */
