"""Multi-rank driver for the CPU oracle: the reference's own halo flow (mpi/mpi.f90:277-387 with the call sites of
dg/dg.f90:335-401 and dg/lifting/lifting_br1.t90:80-165) executed over torch.distributed (gloo on CPU).

SendID=1: send MINE / receive YOUR (master -> slave);  SendID=2: send YOUR / receive MINE (slave -> master).
Test infrastructure only."""
import numpy as np
import torch
import torch.distributed as dist


def exchange_reference(mesh, arr, send_id):
    """arr: numpy [nSides, ...] face array; in place, like StartReceive/StartSend/FinishExchangeMPIData."""
    if not mesh.nNbProcs:
        return
    flat = arr.reshape(arr.shape[0], -1)
    reqs, recvs = [], []
    for ib, nb in enumerate(mesh.NbProc):
        m = (int(mesh.offsetMPISides_MINE[ib]), int(mesh.offsetMPISides_MINE[ib + 1]))
        y = (int(mesh.offsetMPISides_YOUR[ib]), int(mesh.offsetMPISides_YOUR[ib + 1]))
        snd, rcv = (m, y) if send_id == 1 else (y, m)
        if rcv[1] > rcv[0]:
            buf = torch.empty((rcv[1] - rcv[0], flat.shape[1]), dtype=torch.float64)
            reqs.append(dist.irecv(buf, src=int(nb)))
            recvs.append((rcv, buf))
        if snd[1] > snd[0]:
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(flat[snd[0]:snd[1]])), dst=int(nb)))
    for r in reqs:
        r.wait()
    for rcv, buf in recvs:
        flat[rcv[0]:rcv[1]] = buf.numpy()


def execute_plan(plan, am, as_):
    """Executes the library's halo message plan (galaexi_b200.dg.halo_plan) on numpy face arrays [nSides, ...]:
    messages between one pair of ranks are matched in issue order, as NCCL does inside a group."""
    fm, fs = am.reshape(am.shape[0], -1), as_.reshape(as_.shape[0], -1)
    reqs, recvs = [], []
    for peer, is_send, slave, s0, ns in plan:
        a = fs if slave else fm
        if is_send:
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(a[s0:s0 + ns])), dst=peer))
        else:
            buf = torch.empty((ns, a.shape[1]), dtype=torch.float64)
            reqs.append(dist.irecv(buf, src=peer))
            recvs.append((a, s0, ns, buf))
    for r in reqs:
        r.wait()
    for a, s0, ns, buf in recvs:
        a[s0:s0 + ns] = buf.numpy()


class MultiRankOracle:
    """DGTimeDerivative_weakForm / TimeStepByLSERKW2 / CalcTimeStep of one rank, halos over torch.distributed."""

    def __init__(self, case):
        from oracle.oracle import Oracle
        self.c = case
        self.o = Oracle(case)

    def set_state(self, U):
        self.o.set_state(U)

    def time_derivative(self, t=0.0):
        o, m, par = self.o, self.c.mesh, self.c.parabolic
        o.rhs_phase(0)
        exchange_reference(m, o.array("U_slave"), 2)
        o.rhs_phase(1)
        if par:
            exchange_reference(m, o.array("gradUz_slave"), 1)   # lifting flux lives in this buffer
        o.rhs_phase(2)
        if par:
            for nm in ("gradUx_slave", "gradUy_slave", "gradUz_slave"):
                exchange_reference(m, o.array(nm), 2)
        o.rhs_phase(3)
        exchange_reference(m, o.array("Flux_slave"), 1)
        o.rhs_phase(4)
        return o.array("Ut")

    def calc_timestep(self):
        dt = self.o.calc_timestep()[0]
        t = torch.tensor([dt], dtype=torch.float64)
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)   # calctimestep.f90:181
        return float(t.item())

    def rk_step(self, t, dt):
        td = self.c.timedisc
        for st in range(td.nRKStages):
            self.time_derivative(t if st == 0 else t + td.RKc[st] * dt)
            self.o.rk_update(0.0 if st == 0 else -1.0 * td.RKA[st], td.RKb[st] * dt)

    def close(self):
        self.o.close()
