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
