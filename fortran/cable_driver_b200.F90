!> ISO_C_BINDING interfaces of the DRIVER stages either side of cbm() that libcable_b200.so keeps on the device
!! (include/cable_b200.h, second half; SURVEY.md 8f ranks 1 and 2), and how serialdrv would call them.
!!
!! They replace, for a driver that opts in, the per-tile host loops around CALL cbm:
!!   get_met_data's tile expansion + sinbet     src/offline/cable_input.F90:1880-1883, 2139-2213, 2666-2680
!!   ssnow%runoff*dels ..., tscrn daily extremes src/offline/cable_serial.F90:602-608
!!   sumcflux                                    src/science/casa-cnp/casa_sumcflux.F90:76-102
!!   mass_balance / energy_balance               src/offline/cable_checks.F90:472-618
!!   aggregators + grid_cell_average             src/util/aggregator.F90, src/util/cable_grid_reductions.F90:49-75
!! Source only: this image has no Fortran compiler; tests/test_gpu_driver.py drives the same C ABI sequence.
MODULE cable_driver_b200

  USE, INTRINSIC :: ISO_C_BINDING
  IMPLICIT NONE
  PUBLIC

  INTEGER(C_INT), PARAMETER :: CABLE_MET_NROWS = 11     ! SWdown Tair Qair PSurf Wind Rainf Snowf LWdown CO2air hod doy
  INTEGER(C_INT), PARAMETER :: CABLE_AGG_POINT = 0, CABLE_AGG_MEAN = 1, CABLE_AGG_SUM = 2, CABLE_AGG_MIN = 3, CABLE_AGG_MAX = 4

  TYPE, BIND(C) :: cable_met_convert                     ! convert%* of cable_input.F90:1053-1209
     REAL(C_FLOAT)  :: tair_offset, psurf_scale, rainf_scale, co2_scale
     INTEGER(C_INT) :: snowf_from_tair
  END TYPE cable_met_convert

  INTERFACE
     INTEGER(C_INT) FUNCTION cable_b200_driver_init(handle, nland, cstart, cend, patchfrac, latitude) &
          BIND(C, NAME="cable_b200_driver_init")
       IMPORT :: C_INT, C_PTR, C_FLOAT
       TYPE(C_PTR), VALUE :: handle
       INTEGER(C_INT), VALUE :: nland
       INTEGER(C_INT), INTENT(IN) :: cstart(*), cend(*)          ! landpt(:)%cstart-1, landpt(:)%cend-1
       REAL(C_FLOAT), INTENT(IN) :: patchfrac(*), latitude(*)    ! patch(:)%frac, rad%latitude
     END FUNCTION
     INTEGER(C_INT) FUNCTION cable_b200_set_met_async(handle, slot, met_land, cv) BIND(C, NAME="cable_b200_set_met_async")
       IMPORT :: C_INT, C_PTR, C_FLOAT, cable_met_convert
       TYPE(C_PTR), VALUE :: handle
       INTEGER(C_INT), VALUE :: slot
       REAL(C_FLOAT), INTENT(IN) :: met_land(*)                  ! (mland, CABLE_MET_NROWS), column-major
       TYPE(cable_met_convert), INTENT(IN) :: cv
     END FUNCTION
     INTEGER(C_INT) FUNCTION cable_b200_upload_lai(handle) BIND(C, NAME="cable_b200_upload_lai")
       IMPORT :: C_INT, C_PTR
       TYPE(C_PTR), VALUE :: handle
     END FUNCTION
     INTEGER(C_INT) FUNCTION cable_b200_step(handle, ktau, dels, slot) BIND(C, NAME="cable_b200_step")
       IMPORT :: C_INT, C_PTR, C_FLOAT
       TYPE(C_PTR), VALUE :: handle
       INTEGER(C_INT), VALUE :: ktau, slot
       REAL(C_FLOAT), VALUE :: dels
     END FUNCTION
     INTEGER(C_INT) FUNCTION cable_b200_post_step(handle, ktau, kstart, dels, do_mass_bal, do_energy_bal) &
          BIND(C, NAME="cable_b200_post_step")
       IMPORT :: C_INT, C_PTR, C_FLOAT
       TYPE(C_PTR), VALUE :: handle
       INTEGER(C_INT), VALUE :: ktau, kstart, do_mass_bal, do_energy_bal
       REAL(C_FLOAT), VALUE :: dels
     END FUNCTION
     INTEGER(C_INT) FUNCTION cable_b200_output_plan(handle, nrows, field_id, comp, method, scale, div, offset) &
          BIND(C, NAME="cable_b200_output_plan")
       IMPORT :: C_INT, C_PTR, C_FLOAT
       TYPE(C_PTR), VALUE :: handle
       INTEGER(C_INT), VALUE :: nrows
       INTEGER(C_INT), INTENT(IN) :: field_id(*), comp(*), method(*)
       REAL(C_FLOAT), INTENT(IN) :: scale(*), div(*), offset(*)
     END FUNCTION
     INTEGER(C_INT) FUNCTION cable_b200_output_accumulate(handle) BIND(C, NAME="cable_b200_output_accumulate")
       IMPORT :: C_INT, C_PTR
       TYPE(C_PTR), VALUE :: handle
     END FUNCTION
     INTEGER(C_INT) FUNCTION cable_b200_output_fetch_async(handle, host_out) BIND(C, NAME="cable_b200_output_fetch_async")
       IMPORT :: C_INT, C_PTR, C_FLOAT
       TYPE(C_PTR), VALUE :: handle
       REAL(C_FLOAT), INTENT(OUT) :: host_out(*)                 ! (mland, nrows), column-major
     END FUNCTION
     INTEGER(C_INT) FUNCTION cable_b200_output_wait(handle) BIND(C, NAME="cable_b200_output_wait")
       IMPORT :: C_INT, C_PTR
       TYPE(C_PTR), VALUE :: handle
     END FUNCTION
     ! CASA-CNP on the device-resident loop (icycle > 0): bgcdriver and casa_feedback / l_laiFeedbk (fortran/cable_bgcdriver_b200.F90
     ! holds the initialisation: cable_b200_casa_init / _bind / _upload)
     INTEGER(C_INT) FUNCTION cable_b200_bgcdriver(handle, ktau, kstart, kend, dels, ktauday, idoy, loy) &
          BIND(C, NAME="cable_b200_bgcdriver")
       IMPORT :: C_INT, C_PTR, C_FLOAT
       TYPE(C_PTR), VALUE :: handle
       INTEGER(C_INT), VALUE :: ktau, kstart, kend, ktauday, idoy, loy
       REAL(C_FLOAT), VALUE :: dels
     END FUNCTION
     INTEGER(C_INT) FUNCTION cable_b200_casa_feedback(handle, slot, l_vcmaxfeedbk, l_laifeedbk, vcmax_walker2014) &
          BIND(C, NAME="cable_b200_casa_feedback")
       IMPORT :: C_INT, C_PTR
       TYPE(C_PTR), VALUE :: handle
       INTEGER(C_INT), VALUE :: slot, l_vcmaxfeedbk, l_laifeedbk, vcmax_walker2014
     END FUNCTION
  END INTERFACE

CONTAINS

  !> The time loop of serialdrv (src/offline/cable_serial.F90:540-746) with every per-tile stage on the device.
  !! `met_slice(:, :)` is the (mland, 11) block get_met_data has just read for step ktau (its own NF90_GET_VAR
  !! buffers, before the landpt loops); `out(:, :, 2)` receives the grid-cell rows of the output plan.
  SUBROUTINE serial_time_step(handle, ktau, kstart, dels, met_slice, cv, out, do_bal)
    TYPE(C_PTR), INTENT(IN) :: handle
    INTEGER, INTENT(IN) :: ktau, kstart
    REAL, INTENT(IN) :: dels
    REAL(C_FLOAT), INTENT(IN) :: met_slice(:, :)
    TYPE(cable_met_convert), INTENT(IN) :: cv
    REAL(C_FLOAT), INTENT(INOUT) :: out(:, :, :)
    LOGICAL, INTENT(IN) :: do_bal
    INTEGER(C_INT) :: rc, slot
    slot = MOD(ktau, 2)
    rc = cable_b200_set_met_async(handle, slot, met_slice, cv)                 ! get_met_data: tiles, units, snow, coszen
    IF (rc == 0) rc = cable_b200_step(handle, INT(ktau, C_INT), dels, slot)    ! CALL cbm
    IF (rc == 0) rc = cable_b200_post_step(handle, INT(ktau, C_INT), INT(kstart, C_INT), dels, &
                                           MERGE(1_C_INT, 0_C_INT, do_bal), MERGE(1_C_INT, 0_C_INT, do_bal))
    IF (rc == 0) rc = cable_b200_output_wait(handle)                           ! rows of step ktau-1 are in out(:,:,1+MOD(ktau-1,2))
    ! ... cable_output_write of step ktau-1 goes here (netCDF, unchanged) ...
    IF (rc == 0) rc = cable_b200_output_fetch_async(handle, out(:, :, 1 + MOD(ktau, 2)))
    IF (rc /= 0) ERROR STOP 999                                                ! cable_abort convention
  END SUBROUTINE serial_time_step

  !> The same step with CASA-CNP (icycle > 0; cable_serial.F90:584-715): casa_feedback / l_laiFeedbk before cbm, bgcdriver after
  !! it, then sumcflux (its icycle > 0 branch reads casaflux on the device) inside cable_b200_post_step.  Nothing here joins the
  !! step pipeline: every call is per tile (or per land point) and rides the chunk chains.
  SUBROUTINE serial_time_step_casa(handle, ktau, kstart, kend, dels, ktauday, idoy, loy, l_vcmaxFeedbk, l_laiFeedbk, &
                                   met_slice, cv, out, do_bal)
    TYPE(C_PTR), INTENT(IN) :: handle
    INTEGER, INTENT(IN) :: ktau, kstart, kend, ktauday, idoy, loy
    LOGICAL, INTENT(IN) :: l_vcmaxFeedbk, l_laiFeedbk
    REAL, INTENT(IN) :: dels
    REAL(C_FLOAT), INTENT(IN) :: met_slice(:, :)
    TYPE(cable_met_convert), INTENT(IN) :: cv
    REAL(C_FLOAT), INTENT(INOUT) :: out(:, :, :)
    LOGICAL, INTENT(IN) :: do_bal
    INTEGER(C_INT) :: rc, slot
    slot = MOD(ktau, 2)
    rc = cable_b200_set_met_async(handle, slot, met_slice, cv)
    IF (rc == 0) rc = cable_b200_casa_feedback(handle, slot, MERGE(1_C_INT, 0_C_INT, l_vcmaxFeedbk), &
                                               MERGE(1_C_INT, 0_C_INT, l_laiFeedbk), 0_C_INT)       ! cable_serial.F90:587-590
    IF (rc == 0) rc = cable_b200_step(handle, INT(ktau, C_INT), dels, slot)                          ! CALL cbm           :594
    IF (rc == 0) rc = cable_b200_bgcdriver(handle, INT(ktau, C_INT), INT(kstart, C_INT), INT(kend, C_INT), dels, &
                                           INT(ktauday, C_INT), INT(idoy, C_INT), INT(loy, C_INT))   ! CALL bgcdriver     :621
    IF (rc == 0) rc = cable_b200_post_step(handle, INT(ktau, C_INT), INT(kstart, C_INT), dels, &     ! CALL sumcflux      :713
                                           MERGE(1_C_INT, 0_C_INT, do_bal), MERGE(1_C_INT, 0_C_INT, do_bal))
    IF (rc == 0) rc = cable_b200_output_wait(handle)
    IF (rc == 0) rc = cable_b200_output_fetch_async(handle, out(:, :, 1 + MOD(ktau, 2)))
    IF (rc /= 0) ERROR STOP 999
  END SUBROUTINE serial_time_step_casa

END MODULE cable_driver_b200
