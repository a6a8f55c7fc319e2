!> Drop-in replacement of MODULE bgcdriver_mod (src/science/casa-cnp/bgcdriver.F90:1-184): the CASA-CNP daily step on the
!! device that already holds the cbm state (SURVEY.md 8f rank 3, BASELINE config 5).
!!
!! Same module name, same PUBLIC procedure, same argument list as the reference `bgcdriver`, so `serialdrv`
!! (src/offline/cable_serial.F90:621-629) and `mpidrv_worker` call it unchanged, right after CALL cbm.  This file only
!!   1. on the first call initialises the CASA state of the handle that fortran/cable_cbm_b200.F90 created, binds every member
!!      array of casabiome / casapool / casaflux / casamet / casabal / phen with C_LOC (column-major (mp[,k[,b]]) = the library's
!!      layout; biome tables are (mvtype, ...), the soil-order tables (mso)) plus soil%silt / soil%clay, and uploads them;
!!   2. on every call runs cable_b200_bgcdriver(handle, ktau, kstart, kend, dels, ktauday, idoy, loy): the day's accumulation of
!!      casamet / casaflux from the device-resident met%tk, ssnow%tgg, ssnow%wb, canopy%fpn, canopy%frday (no host traffic) and, at
!!      the end of a model day, biogeochem for every tile;
!!   3. at the end of a day brings the per-tile CASA arrays back (cable_b200_casa_download) -- the host's casa output and
!!      restart code reads them once a day; sumcflux's icycle > 0 branch runs on the device inside cable_b200_post_step, or on
!!      the host from the downloaded casaflux as before.
!! What the device does not carry is refused loudly at initialisation (CABLE_E_UNSUPPORTED -> cable_abort): CALL_POP, LALLOC = 2,
!! cable_user%SRF, PHENOLOGY_SWITCH = 'climate', l_landuse; dump_read / dump_write (casa met dump files) stay with the host code.
!! Source only, like cable_cbm_b200.F90 (no Fortran compiler in this image); tests/test_abi.py checks it with f2py's parser.
MODULE bgcdriver_mod

  USE, INTRINSIC :: ISO_C_BINDING
  IMPLICIT NONE
  PRIVATE
  PUBLIC bgcdriver

  !> struct cable_casa_cfg of include/cable_b200.h (field order and types must match)
  TYPE, BIND(C) :: cable_casa_cfg
     INTEGER(C_INT) :: struct_bytes
     INTEGER(C_INT) :: icycle, lalloc, call_climate, l_limit_labile, mvtype
     INTEGER(C_INT) :: call_pop, srf, phenology_climate, l_landuse
  END TYPE cable_casa_cfg

  INTERFACE
     SUBROUTINE cable_b200_casa_default_cfg(cfg) BIND(C, NAME="cable_b200_casa_default_cfg")
       IMPORT :: cable_casa_cfg
       TYPE(cable_casa_cfg), INTENT(OUT) :: cfg
     END SUBROUTINE
     INTEGER(C_INT) FUNCTION cable_b200_casa_init(handle, cfg) BIND(C, NAME="cable_b200_casa_init")
       IMPORT :: C_INT, C_PTR, cable_casa_cfg
       TYPE(C_PTR), VALUE :: handle
       TYPE(cable_casa_cfg), INTENT(IN) :: cfg
     END FUNCTION
     INTEGER(C_INT) FUNCTION cable_b200_casa_bind(handle, name, host) BIND(C, NAME="cable_b200_casa_bind")
       IMPORT :: C_INT, C_PTR, C_CHAR
       TYPE(C_PTR), VALUE :: handle, host
       CHARACTER(KIND=C_CHAR), INTENT(IN) :: name(*)
     END FUNCTION
     INTEGER(C_INT) FUNCTION cable_b200_casa_upload(handle) BIND(C, NAME="cable_b200_casa_upload")
       IMPORT :: C_INT, C_PTR
       TYPE(C_PTR), VALUE :: handle
     END FUNCTION
     INTEGER(C_INT) FUNCTION cable_b200_casa_download(handle) BIND(C, NAME="cable_b200_casa_download")
       IMPORT :: C_INT, C_PTR
       TYPE(C_PTR), VALUE :: handle
     END FUNCTION
     INTEGER(C_INT) FUNCTION cable_b200_bgcdriver(handle, ktau, kstart, kend, dels, ktauday, idoy, loy) &
          BIND(C, NAME="cable_b200_bgcdriver")
       IMPORT :: C_INT, C_PTR, C_FLOAT
       TYPE(C_PTR), VALUE :: handle
       INTEGER(C_INT), VALUE :: ktau, kstart, kend, ktauday, idoy, loy
       REAL(C_FLOAT), VALUE :: dels
     END FUNCTION
  END INTERFACE

  LOGICAL, SAVE :: casa_ready = .FALSE.

CONTAINS

  SUBROUTINE bgcdriver(ktau, kstart, kend, dels, met, ssnow, canopy, veg, soil,                    &
                       climate, casabiome, casapool, casaflux, casamet, casabal, phen,             &
                       pop, spinConv, spinup, ktauday, idoy, loy, dump_read,                       &
                       dump_write, LALLOC)
    USE cable_def_types_mod
    USE cable_common_module, ONLY : cable_user, l_landuse
    USE casadimension,       ONLY : icycle
    USE casavariable
    USE phenvariable
    USE POP_TYPES,           ONLY : POP_TYPE
    USE cable_cbm_module,    ONLY : b200_device_handle
    INTEGER, INTENT(IN) :: ktau, kstart, kend
    INTEGER, INTENT(IN) :: idoy, loy
    INTEGER, INTENT(IN) :: ktauday
    LOGICAL, INTENT(IN) :: spinConv, spinup
    LOGICAL, INTENT(IN) :: dump_read, dump_write
    INTEGER, INTENT(IN) :: LALLOC
    REAL,    INTENT(IN) :: dels
    TYPE (met_type),            INTENT(INOUT)         :: met
    TYPE (soil_snow_type),      INTENT(INOUT)         :: ssnow
    TYPE (canopy_type),         INTENT(INOUT)         :: canopy
    TYPE (veg_parameter_type),  INTENT(INOUT)         :: veg
    TYPE (soil_parameter_type), INTENT(INOUT), TARGET :: soil
    TYPE (casa_biome),          INTENT(INOUT), TARGET :: casabiome
    TYPE (casa_pool),           INTENT(INOUT), TARGET :: casapool
    TYPE (casa_flux),           INTENT(INOUT), TARGET :: casaflux
    TYPE (casa_met),            INTENT(INOUT), TARGET :: casamet
    TYPE (casa_balance),        INTENT(INOUT), TARGET :: casabal
    TYPE (phen_variable),       INTENT(INOUT), TARGET :: phen
    TYPE (POP_TYPE),            INTENT(INOUT)         :: pop
    TYPE (climate_type),        INTENT(IN)            :: climate
    TYPE(cable_casa_cfg) :: cfg
    TYPE(C_PTR)          :: handle
    INTEGER(C_INT)       :: rc

    IF (dump_read .OR. dump_write) CALL fail('casa met dump files (dump_read / dump_write) are host-side I/O')
    handle = b200_device_handle()
    IF (.NOT. C_ASSOCIATED(handle)) CALL fail('bgcdriver called before the first cbm')
    IF (.NOT. casa_ready) THEN
       CALL cable_b200_casa_default_cfg(cfg)
       cfg%icycle = icycle;  cfg%lalloc = LALLOC;  cfg%mvtype = mvtype
       cfg%call_climate   = MERGE(1, 0, cable_user%call_climate)
       cfg%l_limit_labile = MERGE(1, 0, cable_user%l_limit_labile)
       cfg%call_pop = MERGE(1, 0, cable_user%CALL_POP);  cfg%srf = MERGE(1, 0, cable_user%SRF)
       cfg%phenology_climate = MERGE(1, 0, TRIM(cable_user%PHENOLOGY_SWITCH) == 'climate')
       cfg%l_landuse = MERGE(1, 0, l_landuse)
       rc = cable_b200_casa_init(handle, cfg);  CALL check(rc)
       ! one bind per registry row (include/cable_b200_casa_fields.def), generated by tools/gen_casa_fortran_binds.py
#include "cable_b200_casa_binds.inc"
       CALL cbind('soil_silt', C_LOC(soil%silt));  CALL cbind('soil_clay', C_LOC(soil%clay))
       rc = cable_b200_casa_upload(handle);  CALL check(rc)
       casa_ready = .TRUE.
    END IF
    rc = cable_b200_bgcdriver(handle, INT(ktau, C_INT), INT(kstart, C_INT), INT(kend, C_INT), REAL(dels, C_FLOAT),   &
                              INT(ktauday, C_INT), INT(idoy, C_INT), INT(loy, C_INT));  CALL check(rc)
    ! end of a model day: biogeochem has run, the host's daily casa output / restart code reads the pools and fluxes
    IF (MOD(ktau - kstart + 1, ktauday) == 0) THEN
       rc = cable_b200_casa_download(handle);  CALL check(rc)
    END IF

  CONTAINS
    SUBROUTINE cbind(name, ptr)
      CHARACTER(LEN=*), INTENT(IN) :: name
      TYPE(C_PTR), INTENT(IN) :: ptr
      rc = cable_b200_casa_bind(handle, name // C_NULL_CHAR, ptr);  CALL check(rc)
    END SUBROUTINE
    SUBROUTINE check(status)
      INTEGER(C_INT), INTENT(IN) :: status
      IF (status /= 0) CALL fail('device bgcdriver failed (see cable_b200_last_error)')
    END SUBROUTINE
    SUBROUTINE fail(msg)
      USE cable_abort_module, ONLY : cable_abort
      CHARACTER(LEN=*), INTENT(IN) :: msg
      CALL cable_abort('cable_b200: ' // msg, __FILE__, __LINE__)
    END SUBROUTINE
  END SUBROUTINE bgcdriver

END MODULE bgcdriver_mod
