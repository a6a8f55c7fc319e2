// C++ test double of the Fortran caller: fills column-major buffers the way serialdrv would, calls
// cable_cbm_module::cbm_device::cbm(...) with the reference argument list for a few steps and prints a checksum.
// Reads its inputs from a flat binary dump written by tests/test_gpu_host_mirror.py:
//   header: int mp, int nsteps, then for every registry field (registry order) ncomp*mp elements,
//   then nsteps forcing sets (FORCING non-OPTIN fields, registry order).
// With two more arguments (casa_in.bin casa_out.bin) it also plays serialdrv's CASA-CNP part: after every cbm it calls
// bgcdriver_mod::bgc_device::bgcdriver(...) with the reference argument list (cable_serial.F90:621-629).  casa_in.bin: ints icycle,
// LALLOC, mvtype, ktauday, doy0, then every CASA registry row (registry order; per-tile, per-type or per-soil-order extents),
// then soil%silt and soil%clay (mp floats each); casa_out.bin receives every CASA row back.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../cable_b200/csrc/host_mirror.hpp"

using namespace cable_cbm_module;

int main(int argc, char **argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
  FILE *fi = fopen(argv[1], "rb");
  if (!fi) return 2;
  int mp = 0, nsteps = 0;
  if (fread(&mp, 4, 1, fi) != 1 || fread(&nsteps, 4, 1, fi) != 1) return 2;
  const int nf = cable_b200_nfields();
  std::vector<std::vector<char>> buf(nf);
  std::vector<cable_field_info> info(nf);
  for (int id = 0; id < nf; id++) {
    cable_b200_field_info(id, &info[id]);
    size_t bytes = (size_t)mp * info[id].n1 * info[id].n2 * (info[id].dtype == CABLE_DT_F64 ? 8 : 4);
    buf[id].resize(bytes);
    if (fread(buf[id].data(), 1, bytes, fi) != bytes) return 2;
  }
  air_type air; bgc_pool_type bgc; canopy_type canopy; met_type met; balances_type bal; radiation_type rad;
  roughness_type rough; soil_parameter_type soil; soil_snow_type ssnow; veg_parameter_type veg; cbm_scratch_type scr;
  sum_flux_type sum_flux; climate_type climate;
  // point every derived-type member at its buffer (what ALLOCATE does in alloc_cbm_var, cable_define_types.F90:728)
#define SETP(name, member) { int id = cable_b200_field_id(name); member = (decltype(member))buf[id].data(); }
  CABLE_HOST_MIRROR_BIND_ALL(SETP)
#undef SETP
  cable_cfg cfg; cable_b200_default_cfg(&cfg); cfg.output_level = 1;
  // ---- optional CASA-CNP part
  using namespace bgcdriver_mod;
  const bool with_casa = argc >= 5;
  int chead[5] = {1, 0, 17, 8, 1};
  const int ncf = cable_b200_casa_nfields();
  std::vector<std::vector<char>> cbuf(ncf);
  std::vector<float> silt(mp), clay(mp);
  casa_biome casabiome; casa_pool casapool; casa_flux casaflux; casa_met casamet; casa_balance casabal; phen_variable phen; POP_TYPE pop;
  if (with_casa) {
    FILE *fc = fopen(argv[3], "rb");
    if (!fc || fread(chead, 4, 5, fc) != 5) return 2;
    for (int id = 0; id < ncf; id++) {
      cable_field_info ci; int key = 0;
      cable_b200_casa_field_info(id, &ci, &key);
      const size_t lead = key == 0 ? (size_t)mp : key == 1 ? (size_t)chead[2] : 12;
      cbuf[id].resize(lead * ci.n1 * ci.n2 * (ci.dtype == CABLE_DT_F64 ? 8 : 4));
      if (fread(cbuf[id].data(), 1, cbuf[id].size(), fc) != cbuf[id].size()) return 2;
    }
    if (fread(silt.data(), 4, mp, fc) != (size_t)mp || fread(clay.data(), 4, mp, fc) != (size_t)mp) return 2;
    fclose(fc);
#define SETC(name, member) { int id = cable_b200_casa_field_id(name); member = (decltype(member))cbuf[id].data(); }
    CABLE_CASA_MIRROR_BIND_ALL(SETC)
#undef SETC
    cfg.icycle = chead[0];                                  // cbm leaves its simple carbon model out (cbm:214)
  }
  try {
    cbm_device dev(mp, &cfg);
    casa_globals cg; cg.icycle = chead[0]; cg.mvtype = chead[2]; cg.soil_silt = silt.data(); cg.soil_clay = clay.data();
    bgc_device bgc_dev(dev, cg);
    for (int k = 0; k < nsteps; k++) {
      for (int id = 0; id < nf; id++)
        if (info[id].role == CABLE_ROLE_FORCING && !(info[id].flags & CABLE_FLAG_OPTIN))
          if (fread(buf[id].data(), 1, buf[id].size(), fi) != buf[id].size()) return 2;
      dev.cbm(k + 1, 10800.0f, air, bgc, canopy, met, bal, rad, rough, soil, ssnow, sum_flux, veg, climate, scr.xk, scr.c1, scr.rhoch);
      if (with_casa)
        bgc_dev.bgcdriver(k + 1, 1, 1 << 30, 10800.0f, met, ssnow, canopy, veg, soil, climate, casabiome, casapool, casaflux, casamet, casabal,
                          phen, pop, false, false, chead[3], chead[4] + k / chead[3], 365, false, false, chead[1]);
    }
  } catch (const std::exception &e) { fprintf(stderr, "%s\n", e.what()); return 1; }
  fclose(fi);
  FILE *fo = fopen(argv[2], "wb");
  for (int id = 0; id < nf; id++) fwrite(buf[id].data(), 1, buf[id].size(), fo);
  fclose(fo);
  if (with_casa) {
    FILE *fc = fopen(argv[4], "wb");
    for (int id = 0; id < ncf; id++) fwrite(cbuf[id].data(), 1, cbuf[id].size(), fc);
    fclose(fc);
  }
  double s = 0; for (int i = 0; i < mp; i++) s += canopy.fe[i];
  printf("host_mirror ok: mp=%d steps=%d sum(fe)=%.6f\n", mp, nsteps, s);
  return 0;
}
