// host_mirror.hpp -- C++ host-side mirror of the reference interface of the hot path.
//
// The reference is compiled Fortran; no Fortran compiler exists in this image, so the host side above the C ABI
// is mirrored in C++ with the same names and argument list as
//     SUBROUTINE cbm(ktau, dels, air, bgc, canopy, met, bal, rad, rough, soil, ssnow, sum_flux, veg, climate,
//                    xk, c1, rhoch)          (src/offline/cbl_model_driver_offline.F90:38-40)
// and the same division of labour as fortran/cable_cbm_b200.F90: first call = create + bind + upload,
// every call = cable_b200_cbm().  A non-zero status throws (the reference STOPs).
#pragma once
#include <stdexcept>
#include <string>
#include "../../include/cable_b200.h"

namespace cable_cbm_module {

#include "host_mirror_types.inc"

struct sum_flux_type {};   // untouched by cbm (cable_define_types.F90:688-702)

class cbm_device {
 public:
  explicit cbm_device(int mp, const cable_cfg *cfg = nullptr) : mp_(mp) {
    if (cfg) cfg_ = *cfg; else cable_b200_default_cfg(&cfg_);
  }
  ~cbm_device() { if (h_) cable_b200_destroy(h_); }
  cbm_device(const cbm_device &) = delete;
  cbm_device &operator=(const cbm_device &) = delete;

  // same argument list as the reference cbm()
  void cbm(int ktau, float dels, air_type &air, bgc_pool_type &bgc, canopy_type &canopy, met_type &met, balances_type &bal,
           radiation_type &rad, roughness_type &rough, soil_parameter_type &soil, soil_snow_type &ssnow, sum_flux_type &,
           veg_parameter_type &veg, const climate_type &climate, float *xk, float *c1, float *rhoch) {
    if (!h_) {
      check(cable_b200_create(mp_, &cfg_, -1, &h_));
      cbm_scratch_type scr; scr.xk = xk; scr.c1 = c1; scr.rhoch = rhoch;
#define CABLE_HM_BIND(name, ptr) if (ptr) check(cable_b200_bind_field(h_, cable_b200_field_id(name), (void *)(ptr)))
      CABLE_HOST_MIRROR_BIND_ALL(CABLE_HM_BIND)
#undef CABLE_HM_BIND
      check(cable_b200_upload(h_, CABLE_ROLE_PARAM));
      check(cable_b200_upload(h_, CABLE_ROLE_STATE));
    }
    check(cable_b200_cbm(h_, ktau, dels));
  }
  cable_handle *handle() { return h_; }

 private:
  void check(int rc) { if (rc) throw std::runtime_error(std::string("cable_b200: ") + cable_b200_last_error()); }
  int mp_;
  cable_cfg cfg_{};
  cable_handle *h_ = nullptr;
};

}  // namespace cable_cbm_module

// ---- CASA-CNP: C++ mirror of MODULE bgcdriver_mod (src/science/casa-cnp/bgcdriver.F90:1-184), same division of labour as
// fortran/cable_bgcdriver_b200.F90: first call = casa_init + bind + upload on the handle cbm created, every call =
// cable_b200_bgcdriver(), end of a model day = cable_b200_casa_download().
namespace bgcdriver_mod {

using cable_cbm_module::met_type; using cable_cbm_module::soil_snow_type; using cable_cbm_module::canopy_type;
using cable_cbm_module::veg_parameter_type; using cable_cbm_module::soil_parameter_type; using cable_cbm_module::climate_type;

#include "casa_host_mirror_types.inc"

struct POP_TYPE {};        // CALL_POP = .FALSE.: never touched

// the module-scope inputs bgcdriver reads (casadimension::icycle, cable_user%*, mvtype) and soil%silt / soil%clay, which
// are members of soil_parameter_type outside cbm's registry
struct casa_globals {
  int icycle = 1, mvtype = 17;
  bool call_climate = false, l_limit_labile = false, call_pop = false, srf = false, phenology_climate = false, l_landuse = false;
  float *soil_silt = nullptr, *soil_clay = nullptr;   // (mp)
};

class bgc_device {
 public:
  bgc_device(cable_cbm_module::cbm_device &cbm, const casa_globals &g) : cbm_(cbm), g_(g) {}

  // same argument list as the reference bgcdriver()
  void bgcdriver(int ktau, int kstart, int kend, float dels, met_type &, soil_snow_type &, canopy_type &, veg_parameter_type &,
                 soil_parameter_type &, const climate_type &, casa_biome &casabiome, casa_pool &casapool, casa_flux &casaflux,
                 casa_met &casamet, casa_balance &casabal, phen_variable &phen, POP_TYPE &, bool /*spinConv*/, bool /*spinup*/,
                 int ktauday, int idoy, int loy, bool dump_read, bool dump_write, int LALLOC) {
    if (dump_read || dump_write) throw std::runtime_error("cable_b200: casa met dump files are host-side I/O");
    cable_handle *h = cbm_.handle();
    if (!h) throw std::runtime_error("cable_b200: bgcdriver called before the first cbm");
    if (!ready_) {
      cable_casa_cfg c; cable_b200_casa_default_cfg(&c);
      c.icycle = g_.icycle; c.lalloc = LALLOC; c.mvtype = g_.mvtype; c.call_climate = g_.call_climate; c.l_limit_labile = g_.l_limit_labile;
      c.call_pop = g_.call_pop; c.srf = g_.srf; c.phenology_climate = g_.phenology_climate; c.l_landuse = g_.l_landuse;
      check(cable_b200_casa_init(h, &c));
#define CABLE_CM_BIND(name, ptr) if (ptr) check(cable_b200_casa_bind(h, name, (void *)(ptr)))
      CABLE_CASA_MIRROR_BIND_ALL(CABLE_CM_BIND)
#undef CABLE_CM_BIND
      if (g_.soil_silt) check(cable_b200_casa_bind(h, "soil_silt", g_.soil_silt));
      if (g_.soil_clay) check(cable_b200_casa_bind(h, "soil_clay", g_.soil_clay));
      check(cable_b200_casa_upload(h));
      ready_ = true;
    }
    check(cable_b200_bgcdriver(h, ktau, kstart, kend, dels, ktauday, idoy, loy));
    if ((ktau - kstart + 1) % ktauday == 0) check(cable_b200_casa_download(h));   // end of day: the host's daily casa output reads the pools
  }

 private:
  void check(int rc) { if (rc) throw std::runtime_error(std::string("cable_b200: ") + cable_b200_last_error()); }
  cable_cbm_module::cbm_device &cbm_;
  casa_globals g_;
  bool ready_ = false;
};

}  // namespace bgcdriver_mod
