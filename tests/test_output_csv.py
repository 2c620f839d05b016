"""Analysis CSV files in the reference's format (output/output.f90:762-815): parsing the reference's own CSV files and writing
them again must reproduce them byte for byte."""
import os

import numpy as np
import pytest

from galaexi_b200.host_standin import output

REF = "/root/reference"


def test_format_e23_matches_fortran_e_editing():
    assert output.format_e23(0.0) == "0.00000000000000E+00000"
    assert output.format_e23(0.46875000000001e-3) == "0.46875000000001E-00003"
    assert output.format_e23(-10.671740566753) == "-.10671740566753E+00002"
    assert output.format_e23(1.0) == "0.10000000000000E+00001"
    assert output.format_e23(0.999999999999996) == "0.10000000000000E+00001"      # rounding carries into the exponent
    assert output.format_e23(1.5664490941789e-21) == "0.15664490941789E-00020"
    assert output.format_e23(-0.0) == "-.00000000000000E+00000"
    assert all(len(output.format_e23(v)) == 23 for v in (1e-300, -1e300, 3.14, -2.5e-7))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
@pytest.mark.parametrize("rel", ["regressioncheck/checks/tgv/split/TGV_Re1600_Split_TGVAnalysis_Reference.csv",
                                 "regressioncheck/checks/tgv/oInt/TGV_Re1600_OInt_TGVAnalysis_Reference.csv"])
def test_reference_csv_files_are_reproduced_byte_for_byte(tmp_path, rel):
    src = os.path.join(REF, rel)
    lines = open(src).read().splitlines()
    names = lines[0].split(",")[1:]
    conv = lambda s: float(s.replace("E+0", "E+").replace("E-0", "E-"))
    data = np.array([[conv(x) for x in ln.split(",")] for ln in lines[1:]])
    fn, last = output.init_output_to_file(str(tmp_path / "out"), names)
    assert last is None
    output.output_to_file(fn, data[:, 0], data[:, 1:])
    assert open(fn).read().splitlines() == lines
    # a restart at the time of record k resumes the file: that record and everything behind it is cut off and re-written by
    # the restarted run (output.f90:690-729: search, BACKSPACE, ENDFILE); no duplicate time rows
    k = len(data) // 2
    fn2, last = output.init_output_to_file(str(tmp_path / "out"), names, RestartTime=float(data[k, 0]))
    assert fn2 == fn and np.allclose(last, data[k]) and open(fn).read().splitlines() == lines[:1 + k]
    output.output_to_file(fn, data[k:, 0], data[k:, 1:])
    assert open(fn).read().splitlines() == lines
    # a restart time behind the last record: nothing to cut, records are appended
    output.init_output_to_file(str(tmp_path / "out"), names, RestartTime=float(data[-1, 0]) + 1.0)
    assert open(fn).read().splitlines() == lines
    # a fresh run (RestartTime = 0) keeps only the header of an existing file; RestartTime < 0 writes a new file
    output.init_output_to_file(str(tmp_path / "out"), names)
    assert open(fn).read().splitlines() == lines[:1]
    output.output_to_file(fn, data[:3, 0], data[:3, 1:])
    output.init_output_to_file(str(tmp_path / "out"), names, RestartTime=-1.0)
    assert open(fn).read().splitlines() == lines[:1]
    with pytest.raises(RuntimeError, match="cannot open"):
        output.output_to_file(str(tmp_path / "missing"), [0.0], [[1.0]])


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_tgv_column_names_are_the_reference_header():
    from galaexi_b200.host_standin import analyze as an
    hdr = open(os.path.join(REF, "regressioncheck/checks/tgv/split/TGV_Re1600_Split_TGVAnalysis_Reference.csv")).readline().strip().split(",")
    assert hdr == ["Time"] + list(an.TGV_COLUMNS)
