// C++ host side above the C ABI (include/dgx.h): the reference's procedures of the hot path under their own names.
//
// GALAEXI's host is Fortran: module procedures over module-global arrays. A Fortran compiler is not part of
// this build environment, so the compiled-language mirror of that interface is this header-only C++ class; the Fortran
// interface module a maintainer would use is include/dgx_mod.f90.
//
//   reference procedure (file:line, relative to /root/reference/src)            here
//   InitDG + the H2D copies of the Init* routines     dg/dg.f90:67-171          DG::DG(const dgx_config&)
//   FinalizeDG                                         dg/dg.f90:464-495         DG::~DG / FinalizeDG()
//   d_U = U / U = d_U                                  timedisc/timedisc.f90:108 SetState / GetState
//   DGTimeDerivative_weakForm(t)                       dg/dg.f90:255-425         DGTimeDerivative_weakForm(t)
//   TimeStepByLSERKW2(t)                               timedisc/timestep.f90:49  TimeStepByLSERKW2(t, dt)
//   CalcTimeStep(errType)                              .../calctimestep.f90:98   CalcTimeStep(errType)
//   TimeDisc: the time loop incl. UpdateTimeStep       timedisc/timedisc.f90:36-203, timedisc_func.f90:246-300   TimeDisc(t0, tEnd)
//   AnalyzeTestcase (Taylor-Green vortex)              testcase/taylorgreenvortex/testcase.f90:283-515            AnalyzeTestcase(...)
//   CalcForcing / TestcaseSource (channel)             testcase/channel/testcase.f90:241-296                     CalcForcing, SetChannelForcing
//   CalcBodyForces(BodyForce,Fp,Fv)                    equations/navierstokes/calcbodyforces.f90:41-110          CalcBodyForces(...)
//   Abort(__STAMP__, msg)                              globals/globals.f90:175-221 dgx::Abort (exception carrying dgx_last_error)
#pragma once
#include <functional>
#include <vector>
#include <stdexcept>
#include <string>

#include "dgx.h"

namespace dgx {

struct Abort : std::runtime_error {
    using std::runtime_error::runtime_error;
};

class DG {
  public:
    explicit DG(const dgx_config& cfg) {
        if (dgx_create(&h_, &cfg)) {
            std::string msg = h_ ? dgx_last_error(h_) : "dgx_create failed";
            if (h_) dgx_destroy(h_);
            h_ = nullptr;
            throw Abort(msg);
        }
    }
    DG(const DG&) = delete;
    DG& operator=(const DG&) = delete;
    ~DG() { FinalizeDG(); }
    void FinalizeDG() {
        if (h_) dgx_destroy(h_);
        h_ = nullptr;
    }

    void SetState(const double* U) { ck(dgx_set_state(h_, U)); }
    void GetState(double* U) { ck(dgx_get_state(h_, U)); }
    void GetUt(double* Ut) { ck(dgx_get_ut(h_, Ut)); }
    void GetGradients(double* gx, double* gy, double* gz) { ck(dgx_get_gradients(h_, gx, gy, gz)); }
    void SetKeepGradients(bool on) { ck(dgx_set_keep_gradients(h_, on ? 1 : 0)); }

    void DGTimeDerivative_weakForm(double t) { ck(dgx_time_derivative(h_, t)); }
    void TimeStepByLSERKW2(double t, double dt) { ck(dgx_rk_step(h_, t, dt)); }
    // timestep.f90:129 -- the same entry point: the library applies the 3-register update when created with RKdelta/RKg1-3
    void TimeStepByLSERKK3(double t, double dt) { ck(dgx_rk_step(h_, t, dt)); }
    // one iteration of the stage loop (timestep.f90:86-105 / :160-190) for hosts that keep per-stage hooks
    void RKStage(int iStage, double tStage, double dt) { ck(dgx_rk_stage(h_, iStage, tStage, dt)); }
    // sponge/pruettdamping.f90:69-92 and the base flow the reference writes with WriteBaseflow
    void TempFilterTimeDeriv(double dt, double tempFilterWidth) { ck(dgx_temp_filter_time_deriv(h_, dt, tempFilterWidth)); }
    void GetBaseFlow(double* SpBaseFlow) { ck(dgx_get_baseflow(h_, SpBaseFlow)); }
    double CalcTimeStep(int& errType) {
        double dt = 0.0;
        ck(dgx_calc_timestep(h_, &dt, &errType));
        return dt;
    }

    // The time loop of TimeDisc with UpdateTimeStep: dt = min(CalcTimeStep, tEnd - t); a remainder below dt/100 is taken with
    // the last step; an inadmissible state aborts like timedisc_func.f90:262-265. `each` (optional) runs after every step
    // (AnalyzeTimeStep hook). Returns the number of time steps.
    long TimeDisc(double t0, double tEnd, const std::function<void(long, double)>& each = nullptr, long maxIter = -1) {
        double t = t0;
        long iter = 0;
        while (maxIter < 0 || iter < maxIter) {
            int errType = 0;
            double dt = CalcTimeStep(errType);
            if (errType) throw Abort("Error: (1) density, (2) convective / (3) viscous timestep is NaN. Type/time: " + std::to_string(errType) + " " + std::to_string(t));
            const double dtEnd = tEnd - t;
            bool finalize = false;
            if (dt >= dtEnd) { dt = dtEnd; finalize = true; }
            else if (dtEnd - dt < dt / 100.0) { dt = dtEnd; finalize = true; }
            TimeStepByLSERKW2(t, dt);
            t += dt;
            iter++;
            if (each) each(iter, finalize ? tEnd : t);
            if (finalize) break;
        }
        return iter;
    }

    void AnalyzeTestcase(int NAnalyze, const double* Vdm_GaussN_NAnalyze, const double* wAnalyze, double Vol, double rho0, double out15[15]) {
        ck(dgx_analyze_tgv(h_, NAnalyze, Vdm_GaussN_NAnalyze, wAnalyze, Vol, rho0, out15));
    }
    double CalcForcing(const double* wGP, double Vol) {
        double b = 0.0;
        ck(dgx_calc_bulk_velocity(h_, wGP, Vol, &b));
        return b;
    }
    // Fp, Fv, BodyForce: (3,nBCs) column-major; BC(1:nBCSides) the mesh array of 1-based boundary-condition indices
    void CalcBodyForces(const double* wGP, const int* BC, int nBCs, double* BodyForce, double* Fp, double* Fv) {
        ck(dgx_calc_body_forces(h_, wGP, BC, nBCs, Fp, Fv));
        for (int i = 0; i < 3 * nBCs; i++) BodyForce[i] = Fv[i] + Fp[i];
    }
    void CalcWallVelocity(const double* wGP, const int* BC, int nBCs, const double* Surf, double* maxV, double* minV, double* meanV) {
        ck(dgx_calc_wall_velocity(h_, wGP, BC, nBCs, Surf, maxV, minV, meanV));
    }
    void SetChannelForcing(double dpdx, double BulkVel, bool on = true) { ck(dgx_set_channel_forcing(h_, on ? 1 : 0, dpdx, BulkVel)); }

    // The steps between two analyze points without the host in the loop (timedisc.f90:176-200 on the device): adaptive dt kept in
    // device memory, CalcTimeStep inside stage 1's lifting kernel, optional CUDA-graph replay; returns the dt of every step
    // (t advances by their sum, added in step order). Bit-identical to CalcTimeStep + TimeStepByLSERKW2 per step.
    std::vector<double> RunSteps(int nSteps, double t, bool graph = true, bool calcForcingEveryStep = false) {
        float ms = 0.f;
        long long launches = 0;
        ck(dgx_run_steps(h_, nSteps, t, 0.0, 1 | (calcForcingEveryStep ? 2 : 0) | 4 | (graph ? 8 : 0), &ms, &launches));
        std::vector<double> dts((size_t)(nSteps > 0 ? nSteps : 0));
        int count = 0;
        ck(dgx_get_dt_history(h_, nSteps, dts.data(), &count));
        dts.resize((size_t)(count < nSteps ? count : nSteps));
        return dts;
    }
    // face arrays of the last RHS in the reference layout (parity checks): which = 0 U_master ... 8 gradUz_slave (dgx.h)
    void GetFaceArray(int which, double* out) { ck(dgx_get_face_array(h_, which, out)); }

    long long LaunchCount() const { return dgx_launch_count(h_); }
    dgx_handle* handle() { return h_; }

  private:
    void ck(int rc) {
        if (rc) throw Abort(dgx_last_error(h_));
    }
    dgx_handle* h_ = nullptr;
};

}  // namespace dgx
