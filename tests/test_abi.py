"""CPU: the C-ABI library loads, exports every symbol include/cable_b200.h declares, and agrees with the
registry file.  No compute calls (there is no GPU here, and no CPU fallback to call)."""
import ctypes as C
import os
import re

import pytest

from cable_b200 import casa, lib
from cable_b200.registry import FIELDS, DTYPE
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "cable_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(cable_b200_\w+)\s*\(", hdr)))
    assert len(declared) >= 20
    L = lib.load()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/cable_b200.h but not exported"
    assert sorted(lib.EXPORTS + casa.EXPORTS) == declared
    assert L.cable_b200_abi_version() == 3


def test_registry_matches_def_file():
    L = lib.load()
    assert L.cable_b200_nfields() == len(FIELDS)
    codes = {np.float32: 0, np.float64: 1, np.int32: 2}
    for f in FIELDS:
        info = lib.FieldInfo()
        assert L.cable_b200_field_info(f.id, C.byref(info)) == 0
        assert info.name.decode() == f.name
        assert (info.dtype, info.n1, info.n2, info.role, info.flags) == (codes[f.dtype], f.n1, f.n2, f.role, f.flags)
        assert L.cable_b200_field_id(f.name.encode()) == f.id
    assert L.cable_b200_field_id(b"no_such_field") < 0


def test_registry_covers_reference_interface():
    names = {f.name for f in FIELDS}
    # the reference's own per-step forcing message (cable_mpiworker.F90:3449-3609) ...
    for n in ("met_fsd", "met_tk", "met_pmb", "met_qv", "met_ua", "met_precip", "met_precip_sn", "met_fld", "met_ca",
              "met_coszen", "veg_vlai", "met_doy"):
        assert n in names
    # ... and its restart set = prognostic state
    for n in ("ssnow_tgg", "ssnow_wb", "ssnow_wbice", "ssnow_gammzz", "ssnow_tss", "ssnow_ssdnn", "ssnow_ssdn", "ssnow_snowd",
              "ssnow_smass", "ssnow_sdepth", "ssnow_tggsn", "ssnow_snage", "ssnow_rtsoil", "ssnow_isflag", "canopy_cansto",
              "bgc_cplant", "bgc_csoil"):
        assert n in names


def test_default_cfg_is_shipped_namelist():
    cfg = lib.default_cfg()
    assert cfg.struct_bytes == C.sizeof(lib.CableCfg)
    assert (cfg.gs_switch, cfg.fwsoil_switch, cfg.ssnow_potev, cfg.icycle) == (0, 0, 0, 0)   # cable.nml:58-71,37
    assert cfg.snmin == 1.0 and abs(cfg.frozen_limit - 0.85) < 1e-7 and cfg.max_glacier_snowd == 1100.0
    assert [round(z, 3) for z in cfg.zse] == [0.022, 0.058, 0.154, 0.409, 1.085, 2.872]      # cable_parameters.F90:1241
    assert abs(cfg.zshh[0] - 0.011) < 1e-7 and abs(cfg.zshh[6] - 1.436) < 1e-6               # :1828-1831


def test_create_error_behaviour_without_compute():
    L = lib.load()
    h = C.c_void_p()
    cfg = lib.default_cfg()
    # unsupported switches are rejected before any device work (reference: STOP 'fwsoil_switch failed.')
    for sw in ("or_evap", "gw_model", "soil_struc_sli", "runtime_um"):
        cfg = lib.default_cfg()
        setattr(cfg, sw, 1)
        assert L.cable_b200_create(10, C.byref(cfg), 0, C.byref(h)) == -2, sw
        assert b"unsupported" in L.cable_b200_last_error()
    cfg = lib.default_cfg()
    cfg.fwsoil_switch = 3                      # Haverd2013
    assert L.cable_b200_create(10, C.byref(cfg), 0, C.byref(h)) == -2
    cfg = lib.default_cfg()
    cfg.struct_bytes = 4
    assert L.cable_b200_create(10, C.byref(cfg), 0, C.byref(h)) == -1
    cfg = lib.default_cfg()
    assert L.cable_b200_create(0, C.byref(cfg), 0, C.byref(h)) == -1
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        # no device: loud failure, never a CPU fallback
        assert L.cable_b200_create(10, C.byref(cfg), 0, C.byref(h)) == -6
        assert b"no CPU fallback" in L.cable_b200_last_error()
        assert not h.value


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "cable_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "oracle/" not in txt, fn
    for fn in os.listdir(os.path.join(ROOT, "tools")):                      # developer tooling: checkers live in tests/checks/
        txt = open(os.path.join(ROOT, "tools", fn), errors="ignore").read()
        assert "pyoracle" not in txt and "liboracle" not in txt and "import oracle" not in txt and "from oracle" not in txt, fn


def _shim_blocks(path):
    """Parse a Fortran source with f2py's front end (no compiler needed) -> list of (kind, name, args)."""
    import contextlib
    import io
    import numpy.f2py.crackfortran as cf
    out = []

    def walk(blocks):
        for b in blocks:
            if b.get("block") in ("module", "subroutine", "function"):
                out.append((b["block"], b.get("name", "").lower(), [a.lower() for a in b.get("args", [])]))
            walk(b.get("body", []))

    with contextlib.redirect_stdout(io.StringIO()):
        cf.verbose, cf.quiet = 0, 1
        walk(cf.crackfortran([path]))
    return out


def _c_prototypes():
    """name -> number of parameters, from include/cable_b200.h."""
    import re
    txt = open(os.path.join(ROOT, "include", "cable_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", " ", txt, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(cable_b200_\w+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return protos


def test_fortran_shims_parse_and_match_the_c_abi_and_the_reference_signature():
    """No Fortran compiler in this image, so the ISO_C_BINDING shims are checked with f2py's parser: both files parse, the
    drop-in module has the reference's name and `cbm` its 17 dummy arguments in the reference's order
    (cbl_model_driver_offline.F90:38-40), and every BIND(C) interface names an exported symbol of the library with the
    parameter count of its prototype in include/cable_b200.h."""
    protos = _c_prototypes()
    assert set(lib.EXPORTS) <= set(protos), set(lib.EXPORTS) - set(protos)
    cbm_shim = _shim_blocks(os.path.join(ROOT, "fortran", "cable_cbm_b200.F90"))
    drv_shim = _shim_blocks(os.path.join(ROOT, "fortran", "cable_driver_b200.F90"))
    assert ("module", "cable_cbm_module", []) in cbm_shim
    cbm = [b for b in cbm_shim if b[0] == "subroutine" and b[1] == "cbm"]
    assert len(cbm) == 1 and cbm[0][2] == ["ktau", "dels", "air", "bgc", "canopy", "met", "bal", "rad", "rough", "soil", "ssnow",
                                           "sum_flux", "veg", "climate", "xk", "c1", "rhoch"]
    bound = [b for b in cbm_shim + drv_shim if b[1].startswith("cable_b200_") and b[0] in ("function", "subroutine")]
    assert len(bound) >= 15
    for kind, name, args in bound:
        assert name in lib.EXPORTS + casa.EXPORTS, name
        assert len(args) == protos[name], (name, args, protos[name])
    assert any(b[0] == "subroutine" and b[1] == "serial_time_step_casa" for b in drv_shim)


def _shim_types(path):
    """BIND(C) derived types of a shim -> {type name: [(component, 'integer'|'real', n elements)]} in declaration order."""
    import contextlib
    import io
    import numpy.f2py.crackfortran as cf
    out = {}

    def walk(blocks):
        for b in blocks:
            if b.get("block") == "type":
                comps = []
                for v in b["sortvars"]:
                    d = b["vars"][v]
                    n = 1
                    for ext in d.get("dimension") or []:
                        n *= int(ext)
                    comps.append((v.lower(), d["typespec"], d.get("kindselector", {}).get("kind", "").lower(), n))
                out[b["name"].lower()] = comps
            walk(b.get("body", []))

    with contextlib.redirect_stdout(io.StringIO()):
        cf.verbose, cf.quiet = 0, 1
        walk(cf.crackfortran([path]))
    return out


def test_fortran_bind_c_types_mirror_the_c_structs():
    """TYPE, BIND(C) :: cable_cfg / cable_met_convert of the shims against the ctypes mirrors of include/cable_b200.h (which
    test_default_cfg_is_shipped_namelist / the library's struct_bytes check tie to the header): same components, same
    order, same C kinds, same array extents -- a reordered component would be a silent ABI break."""
    import ctypes as C
    kinds = {C.c_int: ("integer", "c_int"), C.c_float: ("real", "c_float"), C.c_double: ("real", "c_double")}

    def mirror(struct):
        out = []
        for name, ct in struct._fields_:
            n = 1
            while hasattr(ct, "_length_"):
                n *= ct._length_; ct = ct._type_
            out.append((name.lower(),) + kinds[ct] + (n,))
        return out

    t = _shim_types(os.path.join(ROOT, "fortran", "cable_cbm_b200.F90"))
    assert t["cable_cfg"] == mirror(lib.CableCfg)
    t = _shim_types(os.path.join(ROOT, "fortran", "cable_driver_b200.F90"))
    assert t["cable_met_convert"] == mirror(lib.MetConvert)


def test_fortran_bind_list_is_generated_from_the_registry_and_names_real_members():
    """fortran/cable_b200_binds.inc is what tools/gen_fortran_binds.py renders from the registry (one bind per row), and --
    where the reference tree is present (this container, not the GPU box) -- every `type%member` it takes C_LOC of is a
    component of that derived type in the reference (cable_define_types.F90, cable_climate_type_mod.F90)."""
    import re
    import subprocess
    import sys
    assert subprocess.call([sys.executable, os.path.join(ROOT, "tools", "gen_fortran_binds.py"), "--check"]) == 0
    inc = open(os.path.join(ROOT, "fortran", "cable_b200_binds.inc")).read()
    binds = re.findall(r"CALL bind\('(\w+)', C_LOC\(([\w%]+)\)\)", inc)
    assert len(binds) == len(FIELDS) and [b[0] for b in binds] == [f.name for f in FIELDS]
    shim = open(os.path.join(ROOT, "fortran", "cable_cbm_b200.F90")).read()
    assert '#include "cable_b200_binds.inc"' in shim
    ref = "/root/reference/src"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not present")
    src = open(os.path.join(ref, "offline", "cable_define_types.F90")).read() + \
        open(os.path.join(ref, "util", "cable_climate_type_mod.F90")).read()
    src = re.sub(r"!.*", "", src).lower()
    type_of = {"met": "met_type", "air": "air_type", "veg": "veg_parameter_type", "soil": "soil_parameter_type",
               "ssnow": "soil_snow_type", "canopy": "canopy_type", "rad": "radiation_type", "rough": "roughness_type",
               "bal": "balances_type", "bgc": "bgc_pool_type", "climate": "climate_type"}
    members = {}
    for var, tname in type_of.items():
        m = re.search(r"\n\s*type\s+" + tname + r"\b(.*?)\n\s*end\s*type", src, flags=re.S)
        assert m, tname
        members[var] = set(re.findall(r"[a-z_]\w*", m.group(1)))
    for name, target in binds:
        if "%" in target:
            var, member = target.split("%")
            assert member.lower() in members[var], (name, target)
        else:
            assert target in ("xk", "c1", "rhoch"), target


def test_generated_host_mirror_types_are_up_to_date():
    """cable_b200/csrc/host_mirror_types.inc and casa_host_mirror_types.inc (the C++ mirrors of the reference's derived types)
    are what tools/gen_host_mirror.py renders from the two registries today."""
    import subprocess
    import sys
    paths = [os.path.join(ROOT, "cable_b200", "csrc", n) for n in ("host_mirror_types.inc", "casa_host_mirror_types.inc")]
    before = [(open(p).read(), os.stat(p)) for p in paths]
    try:
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_host_mirror.py")], stdout=subprocess.DEVNULL)
        for p, (txt, _) in zip(paths, before):
            assert open(p).read() == txt, p
    finally:
        for p, (txt, st) in zip(paths, before):
            open(p, "w").write(txt)
            os.utime(p, ns=(st.st_atime_ns, st.st_mtime_ns))            # keep make from rebuilding the library


def test_casa_registry_matches_def_file_and_reference_types():
    """include/cable_b200_casa_fields.def (generated from the reference's casa_variable.F90 / casa_phenology.F90 by
    tests/golden/gen_casa_registry.py) is what the library was compiled with; where /root/reference is present the member
    names are checked against the type definitions in the source text."""
    from cable_b200 import casa
    L = casa._bind_lib()
    assert L.cable_b200_casa_nfields() == len(casa.FIELDS) > 200
    codes = {np.float32: 0, np.float64: 1, np.int32: 2}
    for f in casa.FIELDS:
        info = lib.FieldInfo(); key = C.c_int(-1)
        assert L.cable_b200_casa_field_info(f.id, C.byref(info), C.byref(key)) == 0
        assert info.name.decode() == f.name and (info.dtype, info.n1, info.n2, key.value) == (codes[f.dtype], f.n1, f.n2, f.key)
        assert L.cable_b200_casa_field_id(f.name.encode()) == f.id
    cfg = casa.default_cfg()
    assert cfg.struct_bytes == C.sizeof(casa.CasaCfg) and (cfg.icycle, cfg.lalloc, cfg.mvtype) == (1, 0, 17)
    src = "/root/reference/src/science/casa-cnp/casa_variable.F90"
    if os.path.exists(src):
        txt = open(src).read().lower() + open("/root/reference/src/science/casa-cnp/casa_phenology.F90").read().lower()
        for f in casa.FIELDS:
            assert re.search(r"\b" + re.escape(f.member) + r"\b", txt), f.name


def test_bgcdriver_shim_matches_the_reference_signature_and_the_c_abi():
    """fortran/cable_bgcdriver_b200.F90 (CASA-CNP drop-in): module and procedure carry the reference's names, `bgcdriver` its 25
    dummy arguments in the reference's order (bgcdriver.F90:7-10, parsed from the reference source where it is present), every
    BIND(C) interface names an exported symbol with its prototype's parameter count, TYPE cable_casa_cfg mirrors the C struct,
    and the bind list is what tools/gen_casa_fortran_binds.py renders from the CASA registry (one bind per row)."""
    import ctypes as C
    import subprocess
    import sys
    from cable_b200 import casa
    path = os.path.join(ROOT, "fortran", "cable_bgcdriver_b200.F90")
    shim = _shim_blocks(path)
    assert ("module", "bgcdriver_mod", []) in shim
    drv = [b for b in shim if b[0] == "subroutine" and b[1] == "bgcdriver"]
    want = ["ktau", "kstart", "kend", "dels", "met", "ssnow", "canopy", "veg", "soil", "climate", "casabiome", "casapool", "casaflux",
            "casamet", "casabal", "phen", "pop", "spinconv", "spinup", "ktauday", "idoy", "loy", "dump_read", "dump_write", "lalloc"]
    assert len(drv) == 1 and drv[0][2] == want
    ref = "/root/reference/src/science/casa-cnp/bgcdriver.F90"
    if os.path.exists(ref):
        rb = [b for b in _shim_blocks(ref) if b[0] == "subroutine" and b[1] == "bgcdriver"]
        assert rb and rb[0][2] == want
    protos = _c_prototypes()
    bound = [b for b in shim if b[1].startswith("cable_b200_") and b[0] in ("function", "subroutine")]
    assert {b[1] for b in bound} == {"cable_b200_casa_default_cfg", "cable_b200_casa_init", "cable_b200_casa_bind",
                                     "cable_b200_casa_upload", "cable_b200_casa_download", "cable_b200_bgcdriver"}
    for kind, name, args in bound:
        assert name in casa.EXPORTS and len(args) == protos[name], (name, args, protos[name])
    t = _shim_types(path)
    assert t["cable_casa_cfg"] == [(n.lower(), "integer", "c_int", 1) for n, _ in casa.CasaCfg._fields_]
    assert subprocess.call([sys.executable, os.path.join(ROOT, "tools", "gen_casa_fortran_binds.py"), "--check"]) == 0
    inc = open(os.path.join(ROOT, "fortran", "cable_b200_casa_binds.inc")).read()
    binds = re.findall(r"CALL cbind\('(\w+)', C_LOC\(([\w%]+)\)\)", inc)
    assert [b[0] for b in binds] == [f.name for f in casa.FIELDS] and all(b[1] == f"{f.type}%{f.member}" for b, f in zip(binds, casa.FIELDS))
    assert '#include "cable_b200_casa_binds.inc"' in open(path).read()
    assert "b200_device_handle" in open(os.path.join(ROOT, "fortran", "cable_cbm_b200.F90")).read()
