"""Stand-in for the reference's unchanged Fortran host (InitMesh / InitInterpolation / InitEquation / TimeDisc / IO).

NOT the deliverable: in a drop-in deployment the reference's own Fortran host (src/mesh, src/interpolation, src/equations,
src/timedisc, src/io_hdf5) builds these tables and calls the C ABI of include/dgx.h through include/dgx_mod.f90. nvfortran
is absent from this image, so these Python modules restate that init code (same tables, pinned bit-exactly / to 100 eps
against the reference's unit-test dumps in tests/test_host_goldens.py) only to feed the CUDA library and the CPU oracle with
identical inputs in tests and in bench.py.
"""
