"""State files in the reference layout and restart from them.

Follows /root/reference/src:
  io_hdf5/hdf5_output.f90:84-211   WriteState: DG_Solution(nVar,0:NOut,0:NOut,0:NOut,nGlobalElems) + ElemData, optional
                                   output degree NOut (J*U and J projected separately, U = (JU)/J)
  io_hdf5/hdf5_output.f90:743-849  GenerateFileSkeleton: root attributes VarNames, Program, File_Type, Project_Name,
                                   File_Version, N, Dimension, Time, MeshFile, NextFile, NodeType, NComputation
  io_hdf5/hdf5_output.f90:856-871  MarkWriteSuccessfull: the TIME attribute is added last
  io_hdf5/hdf5_output.f90:877-931  FlushFiles: delete the chain of state files that would be rewritten
  globals/globals.f90:455-482      TIMESTAMP: Project_State_0000001.000000000.h5
  restart/restart.f90:60-135       InitRestart: N_Restart / NodeType_Restart / RestartTime, InterpolateSolution
  restart/restart.f90:304-560      Restart: read U, or interpolate from (N_Restart, NodeType_Restart); for N_Restart > N the
                                   projection is conservative (J on 3*NGeo points -> N_Restart, JU projected, * sJ)
The files are written by h5write (no libhdf5 here) with the structures libhdf5 emits, so the untouched posti tools and
the reference's own Restart read them; the arrays stay in the reference's memory layout, which is the layout
dgx_get_state / dgx_set_state use: C order [elem][k][j][i][var] == Fortran (var,i,j,k,elem).

Several ranks (one per GPU): the data set is contiguous with the element index slowest and the partition is a set of
contiguous element ranges (mesh_readin.f90:766-778), so every rank's slab is one byte range of the file. Rank 0 writes
the skeleton, every rank writes its own range (the role of the collective MPI-IO write in GatheredWriteArray), rank 0
marks the file complete. ``barrier`` is torch.distributed.barrier or any callable.
"""
from __future__ import annotations

import datetime
import os

import numpy as np

from . import basis as bs
from . import h5lite
from . import h5write
from .metrics import change_basis_volume

STR_VAR_NAMES = ("Density", "MomentumX", "MomentumY", "MomentumZ", "EnergyStagnationDensity")   # equation_vars.f90 StrVarNames
FILE_VERSION = 0.1
_DATA_RESERVE = 256   # room in the metadata block for the TIME attribute added by mark_write_successful


def timestamp(name: str, t: float) -> str:
    """TIMESTAMP(Filename,Time): F17.9 with the leading blanks replaced by zeros."""
    return name + "_" + ("%17.9f" % t).replace(" ", "0")


def state_file_name(project: str, t: float, file_type: str = "State") -> str:
    return timestamp(project + "_" + file_type, t) + ".h5"


def make_userblock(ini_text: str) -> bytes:
    """Userblock of a state file: the ini file between the reference's markers (output.f90:160-180). The compressed build
    information the reference appends ({[( COMPRESSED )]}) describes the Fortran build and is left out."""
    return ("{[( START USERBLOCK )]}\n{[( INIFILE )]}\n" + ini_text.rstrip("\n") + "\n{[( END USERBLOCK )]}\n").encode()


def project_to_nout(U: np.ndarray, sJ: np.ndarray, N: int, NOut: int, node_type: str) -> np.ndarray:
    """hdf5_output.f90:131-149: U on NOut from the separately projected J*U and J."""
    V = bs.get_vandermonde(N, node_type, NOut, node_type, modal=True)
    J = 1.0 / sJ
    JU = change_basis_volume(V, U * J[..., None])
    JOut = change_basis_volume(V, J[..., None])
    return JU / JOut


def _skeleton(project, file_type, mesh_file, N, NOut, node_type, t, t_next, nGlobalElems, nVar, var_names, elem_names,
              userblock) -> h5write.H5Writer:
    w = h5write.H5Writer(userblock)
    w.set_attr("VarNames", [h5write.fortran_str(s) for s in var_names])
    w.set_attr("Program", "Flexi")
    w.set_attr("File_Type", file_type)
    w.set_attr("Project_Name", project)
    w.set_attr("File_Version", FILE_VERSION)
    w.set_attr("N", N)
    w.set_attr("Dimension", 3)
    w.set_attr("Time", float(t))
    w.set_attr("MeshFile", mesh_file)
    w.set_attr("NextFile", [h5write.fortran_str(state_file_name(project, t_next, file_type))])
    w.set_attr("NodeType", [h5write.fortran_str(node_type)])
    w.set_attr("NComputation", N)
    w.reserve_dataset("DG_Solution", (nGlobalElems, NOut + 1, NOut + 1, NOut + 1, nVar), "<f8")
    if elem_names:
        w.set_attr("VarNamesAdd", [h5write.fortran_str(s) for s in elem_names])
        w.reserve_dataset("ElemData", (nGlobalElems, len(elem_names)), "<f8")
    return w


def write_state(U: np.ndarray, N: int, node_type: str, project: str, mesh_file: str, t: float, t_next: float, *,
                out_dir: str = ".", sJ: np.ndarray | None = None, NOut: int | None = None, elem_data: dict | None = None,
                ini_text: str = "", is_error_file: bool = False, offsetElem: int = 0, nGlobalElems: int | None = None,
                rank: int = 0, barrier=None, now: datetime.datetime | None = None) -> str:
    """WriteState(MeshFileName,OutputTime,FutureTime,isErrorFile). ``U`` is this rank's state [e,k,j,i,var] (what
    DGSolver.get_state returns), ``elem_data`` maps a name to a per-element array or a scalar (the ElementOut list:
    e.g. myRank, dt). Returns the file name. Called by all ranks when ``barrier`` is given."""
    nE = U.shape[0]
    nG = nE if nGlobalElems is None else nGlobalElems
    NOut = N if NOut is None else NOut
    if NOut != N:
        if sJ is None:
            raise ValueError("NOut != N needs sJ (the projection is done on J*U)")
        U = project_to_nout(U, sJ, N, NOut, node_type)
    ftype = "ERROR_State" if is_error_file else "State"
    path = os.path.join(out_dir, state_file_name(project, t, ftype))
    names = list(elem_data) if elem_data else []
    w = _skeleton(project, "State", mesh_file, N, NOut, node_type, t, t_next, nG, U.shape[-1], STR_VAR_NAMES[:U.shape[-1]],
                  names, make_userblock(ini_text) if ini_text else b"")
    off, data_start = w.layout(reserve=_DATA_RESERVE)
    if rank == 0:
        os.makedirs(out_dir, exist_ok=True)
        w.write(path, data_start=data_start)
    if barrier is not None:
        barrier()
    with open(path, "r+b") as f:
        f.seek(off["DG_Solution"] + offsetElem * (NOut + 1) ** 3 * U.shape[-1] * 8)
        np.ascontiguousarray(U, dtype="<f8").tofile(f)
        if names:
            ed = np.empty((nE, len(names)))
            for i, k in enumerate(names):
                ed[:, i] = np.asarray(elem_data[k], dtype=np.float64)
            f.seek(off["ElemData"] + offsetElem * len(names) * 8)
            ed.tofile(f)
    if barrier is not None:
        barrier()
    if rank == 0:
        mark_write_successful(w, path, data_start, now)
    return path


def write_baseflow(SpBaseFlow: np.ndarray, N: int, node_type: str, project: str, mesh_file: str, t: float, t_next: float, *,
                   out_dir: str = ".", offsetElem: int = 0, nGlobalElems: int | None = None, rank: int = 0, barrier=None,
                   now: datetime.datetime | None = None) -> str:
    """WriteBaseflow(MeshFileName,OutputTime,FutureTime) (hdf5_output.f90:527-603): the sponge base flow [e,k,j,i,var] as
    Project_BaseFlow_<time>.h5 -- a file of File_Type 'BaseFlow' with the data set DG_Solution on the computation's degree, no
    userblock, no ElemData. InitSponge reads it back on a restart (sponge.f90:174-190)."""
    nE = SpBaseFlow.shape[0]
    nG = nE if nGlobalElems is None else nGlobalElems
    path = os.path.join(out_dir, state_file_name(project, t, "BaseFlow"))
    w = _skeleton(project, "BaseFlow", mesh_file, N, N, node_type, t, t_next, nG, SpBaseFlow.shape[-1], STR_VAR_NAMES[:SpBaseFlow.shape[-1]],
                  [], b"")
    off, data_start = w.layout(reserve=_DATA_RESERVE)
    if rank == 0:
        os.makedirs(out_dir, exist_ok=True)
        w.write(path, data_start=data_start)
    if barrier is not None:
        barrier()
    with open(path, "r+b") as f:
        f.seek(off["DG_Solution"] + offsetElem * (N + 1) ** 3 * SpBaseFlow.shape[-1] * 8)
        np.ascontiguousarray(SpBaseFlow, dtype="<f8").tofile(f)
    if barrier is not None:
        barrier()
    if rank == 0:
        mark_write_successful(w, path, data_start, now)
    return path


def read_baseflow(path: str, N: int, node_type: str, *, offsetElem: int = 0, nElems: int | None = None,
                  nGlobalElems: int | None = None, nVar: int = 5) -> np.ndarray:
    """ReadBaseFlow(FileName) (sponge.f90:468-518): DG_Solution of a BaseFlow / State / TimeAvg file, interpolated (plain change of
    basis, no Jacobian) when its degree or node type differ."""
    info = read_state_attrs(path)
    if nGlobalElems is not None and info["nGlobalElems"] != nGlobalElems:
        raise RuntimeError(f"Baseflow file does not match solution. Elements,nVar {info['nGlobalElems']} {info['nVar']}")
    nE = info["nGlobalElems"] - offsetElem if nElems is None else nElems
    with h5lite.H5File(path) as f:
        U = np.array(f.dataset_rows("DG_Solution", offsetElem, nE)[..., :nVar])
    if info["N"] == N and info["NodeType"] == node_type:
        return np.ascontiguousarray(U)
    return change_basis_volume(bs.get_vandermonde(info["N"], info["NodeType"], N, node_type, modal=True), U)


def mark_write_successful(w: h5write.H5Writer, path: str, data_start: int, now: datetime.datetime | None = None):
    """MarkWriteSuccessfull: DATE_AND_TIME values as the TIME attribute, written when all data is in the file."""
    d = now or datetime.datetime.now().astimezone()
    tz = int(d.utcoffset().total_seconds() // 60) if d.utcoffset() is not None else 0
    w.set_attr("TIME", np.array([d.year, d.month, d.day, tz, d.hour, d.minute, d.second, d.microsecond // 1000], dtype=np.int32))
    w.write(path, data_start=data_start, metadata_only=True)


def read_state_attrs(path: str) -> dict:
    """What InitRestart reads: N_Restart, NodeType_Restart, RestartTime, nVar_Restart, nElems_Restart."""
    with h5lite.H5File(path) as f:
        a = {k: np.array(v) for k, v in f.attrs().items()}
        shape = f.dataset_shape("DG_Solution")
    # N_Restart is the degree of the DATA (hdf5_input.f90:386 GetDataProps: N_HDF5 = Dims(2)-1); the root attribute 'N' is always
    # the computation degree PP_N, also when the state was written on NOut /= N
    return dict(N=int(shape[1]) - 1, NComputation=int(a["N"][0]), NodeType=a["NodeType"][0].decode().strip(), Time=float(a["Time"][0]),
                MeshFile=a["MeshFile"][0].decode().strip(), Project_Name=a["Project_Name"][0].decode().strip(),
                NextFile=a["NextFile"][0].decode().strip() if "NextFile" in a else "",
                complete="TIME" in a, nVar=shape[-1], nGlobalElems=shape[0], shape=shape)


def restart(path: str, N: int, node_type: str, *, sJ: np.ndarray | None = None, detJac_Ref: np.ndarray | None = None,
            NGeo: int = 1, offsetElem: int = 0, nElems: int | None = None, nGlobalElems: int | None = None,
            ResetTime: bool = False, nVar: int = 5) -> tuple[np.ndarray, float]:
    """InitRestart + Restart for RestartMode 1 (a state file with the conservative variables): returns this rank's
    U [e,k,j,i,var] on (N, node_type) and RestartTime."""
    info = read_state_attrs(path)
    NR, ntR = info["N"], info["NodeType"]
    shape = info["shape"]
    if nGlobalElems is not None and shape[0] != nGlobalElems or shape[2] != NR + 1 or shape[3] != NR + 1:
        raise RuntimeError("Dimensions of restart file do not match!")
    if info["nVar"] < nVar:
        raise RuntimeError("Provided file for restart has not all conservative/primitive variables available!")
    nE = shape[0] - offsetElem if nElems is None else nElems
    with h5lite.H5File(path) as f:
        U_local = np.array(f.dataset_rows("DG_Solution", offsetElem, nE)[..., :nVar])
    t = 0.0 if ResetTime else info["Time"]
    interpolate = NR != N or ntR != node_type
    if not interpolate:
        return np.ascontiguousarray(U_local), t
    if min(NR, N) < NGeo:
        print(f"WARNING: The geometry is or was underresolved and will potentially change on restart! (N={N}, N_Restart={NR}, NGeo={NGeo})")
    V = bs.get_vandermonde(NR, ntR, N, node_type, modal=True)
    if NR > N:
        if sJ is None or detJac_Ref is None:
            raise ValueError("restart from a higher degree needs sJ and detJac_Ref (conservative projection)")
        V3 = bs.get_vandermonde(3 * NGeo, node_type, NR, ntR, modal=True)
        JNR = change_basis_volume(V3, detJac_Ref[..., None])
        U = change_basis_volume(V, U_local * JNR)
        U *= sJ[..., None]          # ApplyJacobianCons(U,toPhysical=.TRUE.)
        return U, t
    return change_basis_volume(V, U_local), t


def flush_files(project: str, flush_time: float = 0.0, out_dir: str = ".") -> list[str]:
    """FlushFiles(FlushTime): delete the state file at FlushTime's successors along the NextFile chain, and the file at
    FlushTime itself (a restart re-writes it as its first output). Returns the deleted paths."""
    deleted = []
    name = state_file_name(project, flush_time)
    seen = set()
    while name and name not in seen:
        seen.add(name)
        p = os.path.join(out_dir, name)
        if not os.path.exists(p):
            break
        try:
            nxt = read_state_attrs(p)["NextFile"]
        except Exception:
            nxt = ""
        os.remove(p)
        deleted.append(p)
        name = nxt
    return deleted
