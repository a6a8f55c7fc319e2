!> Drop-in replacement of MODULE cable_cbm_module (src/offline/cbl_model_driver_offline.F90:29-233).
!!
!! Same module name, same PUBLIC procedure, same argument list as the reference `cbm`, so
!! `serialdrv` (src/offline/cable_serial.F90:594) and `mpidrv_worker`
!! (src/offline/cable_mpiworker.F90:503) call it unchanged.  The physics runs in libcable_b200.so
!! (include/cable_b200.h); this file only
!!   1. on the first call creates a handle, binds every member array of the derived types with
!!      C_LOC (the arrays are (mp[,k[,b]]) column-major = the library's SoA layout), uploads
!!      parameters and prognostic state;
!!   2. on every call runs cable_b200_cbm(handle, ktau, dels): forcing (+ the caller's canopy%oldcansto) H2D, one
!!      fused step, state + driver-visible diagnostics D2H (or the cable_b200_set_output_mask selection), synchronise;
!!   3. turns a non-zero status into the reference's own abort (src/offline/cable_abort.F90).
!!
!! NOTE: this image has no Fortran compiler (gfortran/flang/nvfortran/ifx all absent), so this shim
!! is source only; tests/ drive the identical C ABI sequence from C++ (cable_b200/csrc/host_mirror.hpp)
!! and Python (cable_b200/cbm.py).  Build it with the reference's CMake by replacing
!! src/offline/cbl_model_driver_offline.F90 with this file and linking -lcable_b200 -lcudart.
MODULE cable_cbm_module

  USE, INTRINSIC :: ISO_C_BINDING
  IMPLICIT NONE
  PRIVATE
  PUBLIC cbm
  PUBLIC b200_device_handle          ! the device handle, for the CASA-CNP shim (fortran/cable_bgcdriver_b200.F90)

  !> struct cable_cfg of include/cable_b200.h (field order and types must match)
  TYPE, BIND(C) :: cable_cfg
     INTEGER(C_INT) :: struct_bytes
     INTEGER(C_INT) :: gs_switch, fwsoil_switch, ssnow_potev, diag_soil_resp_on
     INTEGER(C_INT) :: l_new_runoff_speed, l_new_reduce_soilevp
     INTEGER(C_INT) :: litter, or_evap, gw_model, l_rev_corr, soil_thermal_fix
     INTEGER(C_INT) :: l_new_roughness_soil, call_climate, redistrb, soil_struc_sli
     INTEGER(C_INT) :: runtime_um, icycle, mvtype
     REAL(C_FLOAT)  :: snmin, max_glacier_snowd, snow_ccnsw, max_ssdn, max_sconds, frozen_limit
     REAL(C_FLOAT)  :: wiltParam, satuParam
     REAL(C_FLOAT)  :: zse(6), zshh(7), ratecp(3), ratecs(2)
     INTEGER(C_INT) :: met_tv_is_tk, caller_duties, output_level, n_forcing_slots, threads_per_block
  END TYPE cable_cfg

  INTERFACE
     SUBROUTINE cable_b200_default_cfg(cfg) BIND(C, NAME="cable_b200_default_cfg")
       IMPORT :: cable_cfg
       TYPE(cable_cfg), INTENT(OUT) :: cfg
     END SUBROUTINE
     INTEGER(C_INT) FUNCTION cable_b200_create(mp, cfg, device, handle) BIND(C, NAME="cable_b200_create")
       IMPORT :: C_INT, C_PTR, cable_cfg
       INTEGER(C_INT), VALUE :: mp, device
       TYPE(cable_cfg), INTENT(IN) :: cfg
       TYPE(C_PTR), INTENT(OUT) :: handle
     END FUNCTION
     INTEGER(C_INT) FUNCTION cable_b200_field_id(name) BIND(C, NAME="cable_b200_field_id")
       IMPORT :: C_INT, C_CHAR
       CHARACTER(KIND=C_CHAR), INTENT(IN) :: name(*)
     END FUNCTION
     INTEGER(C_INT) FUNCTION cable_b200_bind_field(handle, id, host) BIND(C, NAME="cable_b200_bind_field")
       IMPORT :: C_INT, C_PTR
       TYPE(C_PTR), VALUE :: handle, host
       INTEGER(C_INT), VALUE :: id
     END FUNCTION
     INTEGER(C_INT) FUNCTION cable_b200_upload(handle, role_mask) BIND(C, NAME="cable_b200_upload")
       IMPORT :: C_INT, C_PTR
       TYPE(C_PTR), VALUE :: handle
       INTEGER(C_INT), VALUE :: role_mask
     END FUNCTION
     INTEGER(C_INT) FUNCTION cable_b200_cbm(handle, ktau, dels) BIND(C, NAME="cable_b200_cbm")
       IMPORT :: C_INT, C_PTR, C_FLOAT
       TYPE(C_PTR), VALUE :: handle
       INTEGER(C_INT), VALUE :: ktau
       REAL(C_FLOAT), VALUE :: dels
     END FUNCTION
     INTEGER(C_INT) FUNCTION cable_b200_mark_dirty(handle, id) BIND(C, NAME="cable_b200_mark_dirty")
       IMPORT :: C_INT, C_PTR
       TYPE(C_PTR), VALUE :: handle
       INTEGER(C_INT), VALUE :: id
     END FUNCTION
     INTEGER(C_INT) FUNCTION cable_b200_set_output_mask(handle, field_ids, n) BIND(C, NAME="cable_b200_set_output_mask")
       IMPORT :: C_INT, C_PTR
       TYPE(C_PTR), VALUE :: handle
       INTEGER(C_INT), INTENT(IN) :: field_ids(*)
       INTEGER(C_INT), VALUE :: n
     END FUNCTION
     FUNCTION cable_b200_last_error() BIND(C, NAME="cable_b200_last_error")
       IMPORT :: C_PTR
       TYPE(C_PTR) :: cable_b200_last_error
     END FUNCTION
  END INTERFACE

  TYPE(C_PTR), SAVE :: handle = C_NULL_PTR

CONTAINS

  FUNCTION b200_device_handle() RESULT(h)
    TYPE(C_PTR) :: h
    h = handle
  END FUNCTION b200_device_handle

  SUBROUTINE cbm( ktau, dels, air, bgc, canopy, met, bal, rad, rough, soil,                     &
                  ssnow, sum_flux, veg, climate, xk, c1, rhoch )
    USE cable_def_types_mod
    USE cable_common_module, ONLY : cable_user, cable_runtime, redistrb, snow_ccnsw, max_ssdn,   &
                                    max_sconds, frozen_limit, max_glacier_snowd, snmin, wiltParam, satuParam
    USE casadimension,       ONLY : icycle
    TYPE (air_type),            INTENT(INOUT), TARGET :: air
    TYPE (bgc_pool_type),       INTENT(INOUT), TARGET :: bgc
    TYPE (canopy_type),         INTENT(INOUT), TARGET :: canopy
    TYPE (met_type),            INTENT(INOUT), TARGET :: met
    TYPE (balances_type),       INTENT(INOUT), TARGET :: bal
    TYPE (radiation_type),      INTENT(INOUT), TARGET :: rad
    TYPE (roughness_type),      INTENT(INOUT), TARGET :: rough
    TYPE (soil_snow_type),      INTENT(INOUT), TARGET :: ssnow
    TYPE (sum_flux_type),       INTENT(INOUT)         :: sum_flux
    TYPE (climate_type),        INTENT(IN)            :: climate
    TYPE (soil_parameter_type), INTENT(INOUT), TARGET :: soil
    TYPE (veg_parameter_type),  INTENT(INOUT), TARGET :: veg
    REAL,    INTENT(IN) :: dels
    INTEGER, INTENT(IN) :: ktau
    REAL, TARGET :: c1(mp,nrb), rhoch(mp,nrb), xk(mp,nrb)
    TYPE(cable_cfg) :: cfg
    INTEGER(C_INT)  :: rc

    IF (.NOT. C_ASSOCIATED(handle)) THEN
       CALL cable_b200_default_cfg(cfg)
       cfg%gs_switch      = MERGE(1, 0, cable_user%gs_switch == 'medlyn')
       SELECT CASE (TRIM(cable_user%fwsoil_switch))
       CASE ('standard');                 cfg%fwsoil_switch = 0
       CASE ('non-linear extrapolation'); cfg%fwsoil_switch = 1
       CASE ('Lai and Ktaul 2000');       cfg%fwsoil_switch = 2
       CASE DEFAULT;                      cfg%fwsoil_switch = 3     ! rejected by create(): 'fwsoil_switch failed.'
       END SELECT
       cfg%ssnow_potev       = MERGE(1, 0, cable_user%ssnow_potev == 'P-M')
       cfg%diag_soil_resp_on = MERGE(0, 1, cable_user%diag_soil_resp == 'off' .OR. cable_user%diag_soil_resp == 'OFF')
       cfg%l_new_runoff_speed   = MERGE(1, 0, cable_user%l_new_runoff_speed)
       cfg%l_new_reduce_soilevp = MERGE(1, 0, cable_user%l_new_reduce_soilevp)
       cfg%litter = MERGE(1, 0, cable_user%litter);            cfg%or_evap = MERGE(1, 0, cable_user%or_evap)
       cfg%gw_model = MERGE(1, 0, cable_user%gw_model);        cfg%l_rev_corr = MERGE(1, 0, cable_user%l_rev_corr)
       cfg%soil_thermal_fix = MERGE(1, 0, cable_user%soil_thermal_fix)
       cfg%l_new_roughness_soil = MERGE(1, 0, cable_user%l_new_roughness_soil)
       cfg%call_climate = MERGE(1, 0, cable_user%call_climate); cfg%redistrb = MERGE(1, 0, redistrb)
       cfg%runtime_um = MERGE(1, 0, cable_runtime%um)
       cfg%icycle = icycle;  cfg%mvtype = mvtype
       cfg%snmin = snmin;  cfg%max_glacier_snowd = max_glacier_snowd;  cfg%snow_ccnsw = snow_ccnsw
       cfg%max_ssdn = max_ssdn;  cfg%max_sconds = max_sconds;  cfg%frozen_limit = frozen_limit
       cfg%wiltParam = wiltParam;  cfg%satuParam = satuParam       ! cable_runtime_opts_mod.F90:6-7
       cfg%zse = soil%zse;  cfg%zshh = soil%zshh;  cfg%ratecp = bgc%ratecp;  cfg%ratecs = bgc%ratecs
       ! The Fortran driver keeps doing  canopy%oldcansto = canopy%cansto  itself (cable_serial.F90:573): with
       ! caller_duties = 0 the library treats canopy%oldcansto as a per-step INPUT and uploads the bound array with the
       ! forcing on every call (4 bytes per tile), so cable_canopy.F90:169 sees exactly what the reference would.
       cfg%caller_duties = 0
       rc = cable_b200_create(INT(mp, C_INT), cfg, -1_C_INT, handle);  CALL check(rc)
       ! one bind per registry row (include/cable_b200_fields.def), generated by tools/gen_fortran_binds.py
#include "cable_b200_binds.inc"
       rc = cable_b200_upload(handle, 2_C_INT);  CALL check(rc)     ! CABLE_ROLE_PARAM
       rc = cable_b200_upload(handle, 4_C_INT);  CALL check(rc)     ! CABLE_ROLE_STATE
    END IF
    ! xk, c1, rhoch are explicit-shape dummies: a compiler may hand cbm a contiguous temporary whose address changes
    ! from call to call, so (unlike the POINTER components of the derived types) they are re-bound on every call.
    ! They are scratch of init_radiation, written back only at output_level = 2; re-binding is three table writes.
    CALL bind('scr_xk', C_LOC(xk));  CALL bind('scr_c1', C_LOC(c1));  CALL bind('scr_rhoch', C_LOC(rhoch))
    ! Any OTHER host-side write to a resident array between two calls (restart read, casa feedback into veg%vcmax, a
    ! parameter perturbation ...) must be announced:  rc = cable_b200_mark_dirty(handle, cable_b200_field_id(name)).
    rc = cable_b200_cbm(handle, INT(ktau, C_INT), REAL(dels, C_FLOAT));  CALL check(rc)

  CONTAINS
    SUBROUTINE bind(name, ptr)
      CHARACTER(LEN=*), INTENT(IN) :: name
      TYPE(C_PTR), INTENT(IN) :: ptr
      rc = cable_b200_bind_field(handle, cable_b200_field_id(name // C_NULL_CHAR), ptr);  CALL check(rc)
    END SUBROUTINE
    SUBROUTINE check(status)
      USE cable_abort_module, ONLY : cable_abort
      INTEGER(C_INT), INTENT(IN) :: status
      IF (status /= 0) CALL cable_abort('cable_b200: device cbm failed (see cable_b200_last_error)', __FILE__, __LINE__)
    END SUBROUTINE
  END SUBROUTINE cbm

END MODULE cable_cbm_module
