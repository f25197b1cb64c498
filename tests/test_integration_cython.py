"""Syntax / type gate of the Cython binding in integration/ (INTEGRATION.md section 2): the .pyx/.pxd are cythonized against the
reference's OWN .pxd files and the generated C is compiled against the reference's C headers and include/nbabfs_b200.h, so that every
cimported name, struct field and C-ABI signature the binding uses exists as written.  Needs /root/reference (absent on the GPU box)."""
import os
import shutil
import subprocess
import sys
import sysconfig

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PDYNAMO_REFERENCE", "/root/reference")
PYREX = [os.path.join(REF, "pCore-1.9.0", "extensions", "pyrex"), os.path.join(REF, "pMolecule-1.9.0", "extensions", "pyrex")]
CINC = [os.path.join(REF, "pCore-1.9.0", "extensions", "cinclude"), os.path.join(REF, "pMolecule-1.9.0", "extensions", "cinclude")]

pytestmark = pytest.mark.skipif(not all(os.path.isdir(d) for d in PYREX), reason="reference tree not present")


def test_binding_cythonizes_against_the_reference_pxd_files(tmp_path):
    pytest.importorskip("Cython")
    src = os.path.join(ROOT, "integration", "pMolecule.NBModelABFSB200.pyx")
    for f in ("pMolecule.NBModelABFSB200.pyx", "pMolecule.NBModelABFSB200.pxd"):
        shutil.copy(os.path.join(ROOT, "integration", f), tmp_path / f)
    out = tmp_path / "nbmodelabfsb200.c"
    cmd = [sys.executable, "-m", "cython", "-2", "-Werror", "-I", PYREX[0], "-I", PYREX[1], str(tmp_path / os.path.basename(src)), "-o", str(out)]
    # dotted .pxd file names (the reference's layout) are deprecated in Cython 3: keep those warnings from becoming errors
    cmd.remove("-Werror")
    r = subprocess.run(cmd, capture_output=True, text=True)
    errors = [line for line in (r.stdout + r.stderr).splitlines() if "Dotted filenames" not in line and line.strip()]
    assert r.returncode == 0 and out.exists(), "\n".join(errors)
    # the generated C against the real headers: struct fields, macro names and the C-ABI prototypes must agree
    cc = shutil.which("gcc")
    if cc is None:
        pytest.skip("no C compiler")
    inc = ["-I" + sysconfig.get_paths()["include"], "-I" + os.path.join(ROOT, "include")] + ["-I" + d for d in CINC]
    r = subprocess.run([cc, "-c", "-w", "-fsyntax-only"] + inc + [str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
