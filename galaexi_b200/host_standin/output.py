"""ASCII analysis files in the reference's format (output/output.f90: InitOutputToFile :690-760, OutputToFile :762-815): the
*_TGVAnalysis.csv, *_BodyForces_<BC>.csv, *_WallVel_<BC>.csv ... files reggie compares (compare_data_file in analyze.ini).

CSV variant (ASCIIOutputFormat = CSV): a header line ``Time,<name>,<name>...`` and one record per sample written with
``(E23.14E5, n(",",1X,E23.14E5))`` -- Fortran E editing with a 14-digit fraction in [0.1,1) and a five-digit exponent:
``0.46875000000001E-00003``, negative values drop the leading zero (``-.10671740566753E+00002``) to fit the width of 23.
"""
from __future__ import annotations

import math
import os


def format_e23(x: float) -> str:
    """Fortran ``E23.14E5`` of a double."""
    x = float(x)
    if math.isnan(x):
        return "NaN".rjust(23)
    if math.isinf(x):
        return ("Infinity" if x > 0 else "-Infinity").rjust(23)
    if x == 0.0:
        digits, e = "0" * 14, 0
    else:
        m, ex = ("%.13E" % abs(x)).split("E")            # d.ddddddddddddd, correctly rounded to 14 significant digits
        digits, e = m.replace(".", ""), int(ex) + 1      # -> 0.dddddddddddddd x 10^(e)
    neg = x < 0.0 or (x == 0.0 and math.copysign(1.0, x) < 0.0)
    body = "." + digits + "E" + ("-" if e < 0 else "+") + "%05d" % abs(e)
    return ("-" + body) if neg else ("0" + body)


def init_output_to_file(path: str, var_names, RestartTime: float = 0.0):
    """InitOutputToFile (output/output.f90:637-760): create ``path``.csv with the header, or -- when the file exists and the run
    is a restart (RestartTime >= 0) -- resume it: the records from the first one with Time >= RestartTime on are cut off, so
    that the restarted run does not leave duplicate or overlapping time rows. Like the reference, the search starts behind
    the first record (its header loop reads two lines); a file without a record at or after RestartTime is appended to; a file
    too short to hold a record is rewritten. Returns (file name, lastLine): lastLine = the values of the record the file was
    cut at (the reference's optional argument), or None."""
    fn = path if path.endswith(".csv") else path + ".csv"
    exists = os.path.exists(fn) and RestartTime >= 0.0      # RestartTime = 0: a fresh run keeps the header only
    last = None
    if exists:
        with open(fn, "r") as f:
            lines = f.readlines()
        if len(lines) < 2:
            exists = False                      # "file is broken, rewrite"
        else:
            keep, t, found = 2, 0.0, True       # header + first record are behind the read position
            while t < RestartTime:
                if keep >= len(lines):
                    found = False               # end of file: "failed. Appending data to end of file."
                    break
                try:
                    t = float(lines[keep].split(",")[0])
                except ValueError:
                    found = False
                    break
                keep += 1
            if found:
                cut = keep - 1                  # BACKSPACE + ENDFILE: the record just read goes, with everything behind it
                try:
                    last = [float(v) for v in lines[cut].split(",")]
                except ValueError:
                    last = None
                with open(fn, "w") as f:
                    f.writelines(lines[:cut])
    if not exists:
        with open(fn, "w") as f:
            f.write("Time," + ",".join(var_names) + "\n")
    return fn, last


def output_to_file(path: str, times, rows):
    """OutputToFile(FileName,time,nVar,output): append one record per sample."""
    fn = path if path.endswith(".csv") else path + ".csv"
    if not os.path.exists(fn):
        raise RuntimeError("ERROR: cannot open " + fn)
    with open(fn, "a") as f:
        for t, r in zip(times, rows):
            f.write(format_e23(t) + "".join(", " + format_e23(v) for v in r) + "\n")
