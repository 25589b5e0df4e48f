"""The C-ABI library loads and exports every symbol include/gdm/cuda/gdm_c_api.h declares (CPU only)."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "gdm", "cuda", "gdm_c_api.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(gdm_[a-z0-9_]+)\s*\(", src))
    names -= {n for n in names if n.endswith("_fn")}
    return names


def test_every_declared_symbol_is_exported(lib):
    out = subprocess.check_output(["nm", "-D", "--defined-only",
                                   os.path.join(ROOT, "dealii-galerkin-difference-methods_b200", "libgdm_b200.so")], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    declared = _declared()
    assert len(declared) > 50
    missing = declared - exported
    assert not missing, missing


def test_binding_table_covers_header(lib):
    import gdm_b200
    declared = _declared()
    assert declared == set(gdm_b200.capi.SIGNATURES), declared ^ set(gdm_b200.capi.SIGNATURES)
    assert lib.gdm_api_version() == 1


def test_compute_fails_loudly_without_gpu(lib):
    """No CPU fallback: a description-only context refuses vectors and operators."""
    import ctypes as C
    import pytest
    import gdm_b200
    from gdm_b200 import capi
    ctx = C.c_void_p()
    assert lib.gdm_context_create(-1, None, C.byref(ctx)) == 0
    d = capi.SystemDesc()
    d.dim, d.fe_degree, d.n_components = 2, 3, 1
    d.n_subdivisions[0] = d.n_subdivisions[1] = 8
    d.hi[0] = d.hi[1] = 1.0
    d.rank, d.n_ranks = 0, 1
    sys_h = C.c_void_p()
    assert lib.gdm_system_create(ctx, C.byref(d), C.byref(sys_h)) == 0
    v = C.c_void_p()
    assert lib.gdm_vector_create(sys_h, C.byref(v)) == capi.ERR_CUDA
    od = capi.OperatorDesc()
    od.kind = capi.OP_MASS
    op = C.c_void_p()
    assert lib.gdm_operator_create(sys_h, None, C.byref(od), C.byref(op)) == capi.ERR_CUDA
    assert b"no CPU fallback" in lib.gdm_last_error()
    lib.gdm_system_destroy(sys_h)
    lib.gdm_context_destroy(ctx)


def test_cpp_drivers_compile_against_include_gdm(lib, tmp_path):
    """The C++ mirror of the reference interface (include/gdm/*.h) and the drivers written against it compile and link
    (no run without a GPU): examples/{poisson_01_gdm, mass_01_gdm, advection_01_gdm, cut_poisson_01_gdm, wave_app}.cc."""
    pkg = os.path.join(ROOT, "dealii-galerkin-difference-methods_b200")
    for name in ("poisson_01_gdm", "mass_01_gdm", "advection_01_gdm", "cut_poisson_01_gdm", "wave_app"):
        exe = str(tmp_path / name)
        subprocess.check_call(["g++", "-O0", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                               "-o", exe, os.path.join(ROOT, "examples", name + ".cc"), "-L" + pkg, "-lgdm_b200",
                               "-Wl,-rpath," + pkg])
        assert os.path.exists(exe)
    # every member of the cut-cell set-up mirror, in every dimension (the drivers use only some of them)
    tu = tmp_path / "instantiate.cc"
    tu.write_text('#include <gdm/system.h>\n#include <gdm/matrix_creator.h>\n#include <gdm/vector_tools.h>\n'
                  'template class GDM::CutCellSetup<1>;\ntemplate class GDM::CutCellSetup<2>;\n'
                  'template class GDM::CutCellSetup<3>;\nint main() { return 0; }\n')
    subprocess.check_call(["g++", "-O0", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           "-o", str(tmp_path / "instantiate"), str(tu), "-L" + pkg, "-lgdm_b200", "-Wl,-rpath," + pkg])


def test_cpp_cut_setup_mirror_on_the_host(lib, tmp_path, golden_dir):
    """GDM::CutCellSetup is host-only: the first line of applications/wave/tests/wave_1.output (interpolation error of
    J0(3 pi r) over the inside part, 2D, level set of degree 3) from a C++ program over include/gdm, no GPU involved."""
    pkg = os.path.join(ROOT, "dealii-galerkin-difference-methods_b200")
    tu = tmp_path / "cut_host.cc"
    tu.write_text(r"""
#include <gdm/system.h>
#include <gdm/matrix_creator.h>
#include <gdm/vector_tools.h>
#include <cmath>
#include <cstdio>
using namespace dealii;
struct Sphere : Function<2> { double value(const Point<2> &p, const unsigned int = 0) const override { return std::sqrt(p[0] * p[0] + p[1] * p[1]) - 1.0; } };
struct Exact : Function<2> { double value(const Point<2> &p, const unsigned int = 0) const override { return std::cyl_bessel_j(0.0, 3.0 * M_PI * std::sqrt(p[0] * p[0] + p[1] * p[1])); } };
int main()
{
  GDM::CutCellSetup<2>::Parameters prm;
  prm.kind_mass = true; prm.gp_h_power = 3; prm.ghost_parameter = 0.25 * std::sqrt(3.0); prm.rhs_value = 0.0; prm.level_set_degree = 3;
  GDM::CutCellSetup<2> cut(3, 40, -1.21, 1.21, Sphere(), prm);
  std::vector<double> u(41 * 41);
  Exact exact;
  for (int j = 0; j < 41; ++j)
    for (int i = 0; i < 41; ++i)
      { Point<2> x; x[0] = -1.21 + i * (2.42 / 40); x[1] = -1.21 + j * (2.42 / 40); u[i + 41 * j] = exact.value(x); }
  const auto e = cut.error_norms_inside(u, exact);
  printf("%5d %8.5f %14.8e %14.8e %14.8e\n", 0, 0.0, e[0], e[1], e[2]);
}
""")
    exe = str(tmp_path / "cut_host")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-o", exe, str(tu),
                           "-L" + pkg, "-lgdm_b200", "-Wl,-rpath," + pkg])
    out = subprocess.check_output([exe], text=True).split()
    gold = open(os.path.join(golden_dir, "app_wave_wave_1.output")).readline().split()
    assert out[:2] == gold[:2]
    for a, b in zip(out[2:], gold[2:]):
        assert abs(float(a) - float(b)) <= 6e-9 * float(b), (out, gold)
