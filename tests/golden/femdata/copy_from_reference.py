"""Copies the two FEM (ANSYS) comparison datasets the reference ships with its docs into tests/golden/femdata/
(byte-identical fixtures; SURVEY 8c: the only reference-held data that pins rotated cells and multi-magnet
collections):

  /root/reference/src/magpylib_material_response/package_data/datasets/FEMdata_test_softmag.csv
      docs/examples/soft_magnets.md:47-69 -- hard cuboid (chi 0.5) + soft cube (chi 3999) rotated 45 deg about y;
      columns Bx_k / Bz_k [T] on the lines x = -4..6 mm (1001 points), y = 0, z = -1 / -3 / -5 mm (k = 0, 1, 2)
  /root/reference/src/magpylib_material_response/package_data/datasets/FEMdata_test_cuboids.csv
      docs/examples/cuboids_demagnetization.md -- three cuboids (chi 0.3 / 1.0 / 0.5, two of them rotated);
      columns Bx / By / Bz [T] on the line x = -4..4 mm (301 points), y = 0, z = -1 mm

Run in the build container (the reference tree does not exist on the GPU box):  python tests/golden/femdata/copy_from_reference.py
"""
import pathlib
import shutil

SRC = pathlib.Path("/root/reference/src/magpylib_material_response/package_data/datasets")
DST = pathlib.Path(__file__).resolve().parent
for name in ("FEMdata_test_softmag.csv", "FEMdata_test_cuboids.csv"):
    shutil.copyfile(SRC / name, DST / name)
    print("copied", name, (DST / name).stat().st_size, "bytes")
