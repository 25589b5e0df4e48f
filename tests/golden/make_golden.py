#!/usr/bin/env python
"""Collect the reference's golden vectors into tests/golden/ (run in the build container).

`/root/reference` does not exist on the GPU box, so everything the tests need is
copied here once and committed:
  * the `.output` files the reference's ctest harness diffs stdout against
    (tests/, the applications' tests and the prototypes named in BASELINE.json);
  * `fe_coefficients.json`: the numeric basis tables of `include/gdm/fe.h:62-320`
    parsed into plain numbers (data, not source), used to pin the oracle's closed
    form Lagrange basis.
Usage: python tests/golden/make_golden.py [/root/reference]
"""
import json
import os
import re
import shutil
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

OUTPUTS = [
    "tests/poly_01.output",
    "tests/fe_02_gdm.output",
    "tests/poisson_01_gdm.output",
    "tests/poisson_02_gdm.mpirun=1.output",
    "tests/poisson_02_gdm.mpirun=3.output",
    "tests/mass_01_gdm.output",
    "tests/mass_02_gdm.output",
    "tests/elasticity_01_gdm.output",
    "prototypes/advection_01_gdm.output",
    "prototypes/advection_02_gdm.output",
    "prototypes/cut_poisson_01_gdm.output",
    "applications/wave/tests/wave_0.output",
    "applications/wave/tests/wave_1.output",
    "applications/wave/tests/heat_0.output",
    "applications/wave/tests/heat_1.output",
    "applications/wave/tests/step85_0.output",
    "applications/wave/tests/wave_composite_0.output",
    "applications/wave/tests/heat_composite_0.output",
    "applications/advection/tests/test_01.output",
]


def parse_fe_tables(path):
    """fe.h `all_coefficients`: [degree/2][variant][basis][coef high->low] as floats."""
    src = open(path).read()
    start = src.index("all_coefficients =")
    end = src.index("// clang-format on", start)
    body = src[start:end]
    body = body[body.index("{"):body.rindex("}") + 1]
    # a/b rationals -> numbers, double braces -> brackets
    body = re.sub(r"//[^\n]*", "", body)
    body = re.sub(r"(-?\d+\.\d+)\s*/\s*(\d+\.\d+)", lambda m: repr(float(m.group(1)) / float(m.group(2))), body)
    body = body.replace("{{", "[").replace("}}", "]")
    body = re.sub(r",\s*\]", "]", body)
    tables = json.loads(body)
    return tables


def main():
    for rel in OUTPUTS:
        src = os.path.join(REF, rel)
        dst = os.path.join(HERE, os.path.basename(rel))
        if rel.startswith("prototypes/"):
            dst = os.path.join(HERE, "prototypes_" + os.path.basename(rel))
        if rel.startswith("applications/"):
            dst = os.path.join(HERE, "app_" + rel.split("/")[1] + "_" + os.path.basename(rel))
        shutil.copyfile(src, dst)
        print("copied", rel)
    tables = parse_fe_tables(os.path.join(REF, "include/gdm/fe.h"))
    out = {str(2 * i + 1): t for i, t in enumerate(tables)}
    with open(os.path.join(HERE, "fe_coefficients.json"), "w") as f:
        json.dump(out, f)
    print("fe tables: degrees", list(out), "variants", [len(t) for t in tables])


if __name__ == "__main__":
    main()
