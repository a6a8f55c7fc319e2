// C++ test double of the Fortran caller: fills column-major buffers the way serialdrv would, calls
// cable_cbm_module::cbm_device::cbm(...) with the reference argument list for a few steps and prints a checksum.
// Reads its inputs from a flat binary dump written by tests/test_gpu_host_mirror.py:
//   header: int mp, int nsteps, then for every registry field (registry order) ncomp*mp elements,
//   then nsteps forcing sets (FORCING non-OPTIN fields, registry order).
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
  try {
    cbm_device dev(mp, &cfg);
    for (int k = 0; k < nsteps; k++) {
      for (int id = 0; id < nf; id++)
        if (info[id].role == CABLE_ROLE_FORCING && !(info[id].flags & CABLE_FLAG_OPTIN))
          if (fread(buf[id].data(), 1, buf[id].size(), fi) != buf[id].size()) return 2;
      dev.cbm(k + 1, 10800.0f, air, bgc, canopy, met, bal, rad, rough, soil, ssnow, sum_flux, veg, climate, scr.xk, scr.c1, scr.rhoch);
    }
  } catch (const std::exception &e) { fprintf(stderr, "%s\n", e.what()); return 1; }
  fclose(fi);
  FILE *fo = fopen(argv[2], "wb");
  for (int id = 0; id < nf; id++) fwrite(buf[id].data(), 1, buf[id].size(), fo);
  fclose(fo);
  double s = 0; for (int i = 0; i < mp; i++) s += canopy.fe[i];
  printf("host_mirror ok: mp=%d steps=%d sum(fe)=%.6f\n", mp, nsteps, s);
  return 0;
}
