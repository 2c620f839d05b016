/*
 * dgx.h -- C ABI of the B200-native DGSEM right-hand side + low-storage Runge-Kutta stage.
 *
 * Drop-in boundary for GALAEXI's hot path. The reference has no FFI layer: the boundary is a set of
 * Fortran module procedures operating on module-global device arrays. Each entry point
 * below replaces one of them; the Fortran ISO_C_BINDING interface a maintainer would add is in
 * INTEGRATION.md. All array arguments are HOST pointers in the reference's own memory layout
 * (Fortran column-major, variable index fastest), caller-owned, copied during the call; the library
 * owns all device memory. Every function returns 0 on success, non-zero on error
 * (dgx_last_error gives the message; the Fortran shim maps it to CALL Abort(__STAMP__,...)).
 * Calls are collective across ranks and not thread-safe per handle.
 *
 * Reference interfaces replaced (paths relative to /root/reference/src):
 *   dgx_create            InitDG dg/dg.f90:67-171, InitLifting dg/lifting/lifting.f90:114,
 *                         InitTimeDisc timedisc/timedisc_func.f90:94 (RK tables), InitMPIvars mpi/mpi.f90:156,
 *                         + the H2D copies of mesh/metrics (mesh/mesh.f90:244-245,339-348,403-406,446-451)
 *   dgx_destroy           FinalizeDG dg/dg.f90:464-495, FinalizeLifting lifting.f90:212
 *   dgx_set_state         "d_U = U" timedisc/timedisc.f90:108, dg/dg.f90:168
 *   dgx_get_state         "U = d_U" at analyze steps, timedisc/timedisc_func.f90:347-350
 *   dgx_get_ut            Ut = d_Ut (diagnostics)
 *   dgx_get_gradients     gradUx/y/z = d_gradUx/y/z, testcase/taylorgreenvortex/testcase.f90:361-364
 *   dgx_time_derivative   DGTimeDerivative_weakForm(t) dg/dg.f90:255-425
 *   dgx_rk_stage          one iteration of the stage loop of TimeStepByLSERKW2 timedisc/timestep.f90:86-105
 *   dgx_rk_step           TimeStepByLSERKW2(t) timedisc/timestep.f90:49-120
 *   dgx_calc_timestep     CalcTimeStep(errType) equations/navierstokes/calctimestep.f90:98-186
 *   halo exchange         StartReceive/StartSend/FinishExchangeMPIData mpi/mpi.f90:277,348,494 are internal
 *                         (NCCL send/recv on a dedicated stream); the caller only supplies the neighbour tables.
 */
#ifndef DGX_H
#define DGX_H

#ifdef __cplusplus
extern "C" {
#endif

#define DGX_NVAR      5 /* PP_nVar */
#define DGX_NVARPRIM  6 /* PP_nVarPrim */
#define DGX_NVARLIFT  4 /* PP_nVarLifting with PP_OPTLIFT=1: (u,v,w,T), equations/navierstokes/idealgas/eos.h:145-153 */

typedef struct dgx_handle dgx_handle;

typedef struct dgx_config {
    /* discretisation (compile-time options of the reference become run-time here) */
    int N;              /* polynomial degree PP_N */
    int nodeType;       /* PP_NodeType: 1 Gauss, 2 Gauss-Lobatto */
    int splitDG;        /* SPLIT_DG: -1 off (weak form), 0 SD, 1 MO, 2 DU, 3 KG, 4 PI (src/CMakeLists.txt:73-92) */
    int riemann;        /* RIEMANN: 0 LF, 1 Roe, 2 RoeL2, 3 RoeEntropyFix, 4 HLL, 5 HLLC, 6 HLLE, 7 HLLEM (4-7: non-split
                         * only, src/CMakeLists.txt:97-130); 9 FluxAverage (split only, riemann.f90:1239) */
    int parabolic;      /* PARABOLIC: 0 Euler, 1 Navier-Stokes with BR1 (or BR2, see `lifting`) gradients */
    int viscLaw;        /* PP_VISC: 0 constant, 1 Sutherland */
    /* mesh sizes and side ranges (1-based inclusive, mesh/mesh.f90:259-283) */
    int nElems, nSides, nBCSides;
    int firstInnerSide, lastInnerSide;
    int firstMPISide_MINE, lastMPISide_MINE, firstMPISide_YOUR, lastMPISide_YOUR;
    /* equation of state: EOS_Vars(1:8) = kappa,R,Pr,mu0,Ts,Tref,ExpoSuth,cSuth (eos.h:44-62) */
    double EOS_Vars[8];
    int nRefState;
    const double *RefStatePrim;   /* (6,nRefState), after InitBC (getboundaryflux.f90:194-212: type 27 direction vector) */
    const int *BCSides;           /* (2,nBCSides): BC_TYPE, BC_STATE (getboundaryflux.f90:244-252); types 2,3,4,9,91,23,24,25,27 */
    /* operators (0:N,0:N) / (0:N), dg/dg.f90:181-242 */
    const double *D_T, *D_Hat_T, *DVolSurf, *L_Minus, *L_Plus, *L_HatMinus, *L_HatPlus;
    /* connectivity: ElemToSide(3,6,nElems), S2V2/S2V2_inv(2,0:N,0:N,0:4,1:6) */
    const int *ElemToSide, *S2V2, *S2V2_inv;
    /* geometry: Metrics_{f,g,h}Tilde(3,0:N,0:N,0:N,nElems), sJ(0:N,0:N,0:N,nElems),
     * NormVec/TangVec1/TangVec2(3,0:N,0:N,nSides), SurfElem(0:N,0:N,nSides) */
    const double *Metrics_fTilde, *Metrics_gTilde, *Metrics_hTilde, *sJ;
    const double *NormVec, *TangVec1, *TangVec2, *SurfElem;
    /* time integration: Williamson 2N tables with RKA[0]=RKc[0]=0 (timedisc_vars.f90:122-470) and the
     * already scaled CFL/DFL numbers (timedisc_func.f90:480-550) */
    int nRKStages;
    const double *RKA, *RKb, *RKc;
    double CFLScale, DFLScale;
    /* domain decomposition (mpi/mpi_vars.f90, mesh/prepare_mesh.f90:196-320); nRanks==1: all unused */
    int myRank, nRanks, nNbProcs;
    const int *NbProc;               /* (nNbProcs) neighbour ranks, ascending */
    const int *nMPISides_MINE_Proc;  /* (nNbProcs) */
    const int *nMPISides_YOUR_Proc;  /* (nNbProcs) */
    const int *offsetMPISides_MINE;  /* (0:nNbProcs) */
    const int *offsetMPISides_YOUR;  /* (0:nNbProcs) */
    const char *ncclUniqueId;        /* 128 bytes from ncclGetUniqueId on rank 0 (broadcast by the host), or NULL */
    int device;                      /* CUDA device ordinal for this rank */
    /* lifting variant (PP_Lifting, src/CMakeLists.txt:161-183): 0/1 BR1 (the only one GALAEXI builds), 2 BR2 (host FLEXI
     * code, dg/lifting/lifting_br2.t90:43-311) with the penalties etaBR2 / etaBR2_wall (lifting.f90:86-91) */
    int lifting;
    double etaBR2, etaBR2_wall;
    /* non-conforming interfaces (host FLEXI code src/mortar/; GALAEXI aborts on them, mesh/mesh.f90:140-143):
     * side ranges mesh.f90:271-283, MortarType(2,nSides), MortarInfo(2,4,nMortarSides) (mesh.f90:318-322,
     * prepare_mesh.f90:746-775), 1-D operators (0:N,0:N) as stored by mortar.f90:111-256 (transposed). nMortarSides==0:
     * all unused. */
    int nMortarSides, firstMortarInnerSide, lastMortarInnerSide, firstMortarMPISide, lastMortarMPISide;
    const int *MortarType, *MortarInfo;
    const double *M_0_1, *M_0_2, *M_1_0, *M_2_0;
    /* modal filter of step 1 of the RHS (dg/dg.f90:331, filter/filter.f90:272-306): FilterMat(0:N,0:N) as built by
     * InitFilter (filter.f90:95-212, FilterType cutoff / modal), or NULL for FilterType 0. Applied in place to U at the
     * start of every dgx_time_derivative / RK stage, like the reference. */
    const double *FilterMat;
    /* source term of the manufactured solutions (dg/dg.f90:418 CalcSource, idealgas/exactfunc.f90:946-1113): IniExactFunc
     * (4: oblique sine wave of the convergence tests; 0 or any other value: no source), AdvVel(3), and the node
     * coordinates Elem_xGP(3,0:N,0:N,0:N,nElems) (only read when a source is active) */
    int IniExactFunc;
    double AdvVel[3];
    const double *Elem_xGP;
    /* non-default lifting forms (ini keys doWeakLifting, doConservativeLifting; lifting.f90:81-85, 139-141): weak form
     * (ignored by BR2, which is always strong) and conservative volume form (implied by the weak form) */
    int doWeakLifting, doConservativeLifting;
    /* sponge zone (sponge/sponge.f90; step 13 of the RHS, dg.f90:419): SpongeMat(0:N,0:N,0:N,nElems) = damping sigma / sJ as
     * CalcSpongeRamp leaves it (:449-454), expanded to all elements (zero outside the SpongeMap), and the initial base flow
     * SpBaseFlow(PP_nVar,0:N,0:N,0:N,nElems) (InitSponge :203-243). NULL: no sponge. */
    const double *SpongeMat, *SpBaseFlow;
    /* three-register low-storage Runge-Kutta, TimeDiscType LSERKK3 (timedisc_vars.f90:140-141, 464-773: ketchesonrk4-20,
     * ketchesonrk4-18; stage update of TimeStepByLSERKK3, timestep.f90:129-200): RKdelta, RKg1, RKg2, RKg3 (1:nRKStages);
     * RKb, RKc as above (RKc(1) = 0), RKA unused. All four NULL: Williamson 2N (TimeStepByLSERKW2). */
    const double *RKdelta, *RKg1, *RKg2, *RKg3;
    /* overintegration of JU_t, step 14 of the RHS (dg/overintegration.f90:92-165 InitOverintegration, :179-340; host FLEXI code --
     * GALAEXI's GPU build stops at :108-114): OverintegrationType 0 none, 1 cut-off (OverintegrationMat(0:N,0:N) applied to JU_t,
     * then the Jacobian), 2 conservative cut-off (Vdm_N_NUnder(0:NUnder,0:N), Vdm_NUnder_N(0:N,0:NUnder),
     * sJNUnder(0:NUnder,0:NUnder,0:NUnder,nElems)). The CFL scaling with NEff (timedisc_func.f90:171-173) is the host's. */
    int OverintegrationType, NUnder;
    const double *OverintegrationMat, *Vdm_N_NUnder, *Vdm_NUnder_N, *sJNUnder;
} dgx_config;

int dgx_create(dgx_handle **h, const dgx_config *cfg);
void dgx_destroy(dgx_handle *h);
const char *dgx_last_error(const dgx_handle *h);

/* U / Ut in the reference layout U(PP_nVar,0:N,0:N,0:N,nElems); gradients (DGX_NVARLIFT,0:N,0:N,0:N,nElems).
 * host pointers; pinned memory makes the copies asynchronous-capable but is not required */
int dgx_set_state(dgx_handle *h, const double *U);
int dgx_get_state(dgx_handle *h, double *U);
int dgx_get_ut(dgx_handle *h, double *Ut);
int dgx_get_gradients(dgx_handle *h, double *gradUx, double *gradUy, double *gradUz);
/* The volume gradients d_gradUx/y/z are read by analysis only (testcase.f90:361-364); the viscous volume integral is formed
 * inside the lifting kernel and the 12 gradient components per node are not written in every stage. on != 0 (default): the
 * LAST stage of every dgx_rk_step stores them, so that after a step they hold what the reference's device arrays hold (the
 * gradients of the last stage's RHS); on == 0: RK stages never store them (steps that no analysis follows).
 * dgx_time_derivative always stores them. dgx_get_gradients / dgx_analyze_tgv fail if the last RHS did not store. */
int dgx_set_keep_gradients(dgx_handle *h, int on);

int dgx_time_derivative(dgx_handle *h, double t);
int dgx_rk_stage(dgx_handle *h, int iStage /* 1-based */, double t, double dt);
int dgx_rk_step(dgx_handle *h, double t, double dt);
/* dt = min(convective, viscous) over all ranks; errType != 0 if the state is not finite */
int dgx_calc_timestep(dgx_handle *h, double *dt, int *errType);

/* AnalyzeTestcase of testcase/taylorgreenvortex/testcase.f90:283-515, on the device (the reference downloads U and the three
 * gradient arrays and integrates on the host, :361-364): the nTGVvars = 15 columns of the *_TGVAnalysis.csv file
 * (Dissipation Rate Incompressible, ... , ED_D; order of testcase.f90:497-498), integrated on the Gauss-Lobatto analysis
 * nodes of degree NAnalyze from U and the lifted gradients of the most recent dgx_time_derivative / RK stage -- i.e. the
 * data the reference analyses when called after TimeStep (timedisc_func.f90:351).
 *   Vdm_GaussN_NAnalyze(0:NAnalyze,0:N), wAnalyze(0:NAnalyze): analyze.f90:236-272 (wGPVolAnalyze is the tensor product of
 *   wAnalyze); Vol: global volume (analyze.f90:160-189); rho0: testcase reference density. Sums are reduced over all ranks.
 * Requires PARABOLIC. */
int dgx_analyze_tgv(dgx_handle *h, int NAnalyze, const double *Vdm_GaussN_NAnalyze, const double *wAnalyze, double Vol,
                    double rho0, double *out15);

/* Channel testcase (testcase/channel/testcase.f90; absent from GALAEXI's GPU path, dg.f90:419 is commented out):
 * dgx_calc_bulk_velocity = CalcForcing (:241-271): BulkVel = 1/Vol sum u wGPVol / sJ over the solution nodes of all ranks,
 * wGP(0:N) the 1-D weights of the solution nodes, Vol the global volume.
 * dgx_set_channel_forcing = the parameters of TestcaseSource (:277-296): with on != 0 every following RHS adds
 * Ut(MOM1) -= dpdx / sJ, Ut(ENER) -= dpdx / sJ * BulkVel before the Jacobian is applied. */
int dgx_calc_bulk_velocity(dgx_handle *h, const double *wGP, double Vol, double *BulkVel);
int dgx_set_channel_forcing(dgx_handle *h, int on, double dpdx, double BulkVel);

/* CalcBodyForces (equations/navierstokes/calcbodyforces.f90:41-110, called from AnalyzeEquation, analyze_equation.f90:205),
 * on the device from the face states and lifted gradient traces of the most recent dgx_time_derivative (the reference calls
 * DGTimeDerivative_weakForm before analysing, timedisc_func.f90:348): for every boundary condition iBC that is a wall (BC type
 * 3, 4 or 9, analyze_equation.f90:113-121) Fp(:,iBC) = sum p n wGPSurf SurfElem and Fv(:,iBC) = -sum tau n wGPSurf SurfElem
 * over its sides, summed over all ranks (every rank gets the result); zero for other boundary conditions.
 *   wGP(0:N): 1-D weights of the solution nodes; BC(1:nBCSides): the mesh array BC (1-based index of the boundary condition
 *   of every boundary side, mesh_vars.f90); Fp, Fv: (3,nBCs) column-major. BodyForce = Fp + Fv is left to the caller. */
int dgx_calc_body_forces(dgx_handle *h, const double *wGP, const int *BC, int nBCs, double *Fp, double *Fv);

/* CalcWallVelocity (equations/navierstokes/analyze_equation.f90:435-499), same data and arguments: per wall boundary condition
 * the maximum, minimum and surface mean (sum |v| wGPSurf SurfElem / Surf(iBC)) of the velocity magnitude on its sides, reduced
 * over all ranks. Surf(1:nBCs): the global surface of every boundary condition (analyze.f90:200-230). Boundary conditions
 * without wall sides return the reference's initial values maxV = -1e14, minV = 1e14, meanV = 0. */
int dgx_calc_wall_velocity(dgx_handle *h, const double *wGP, const int *BC, int nBCs, const double *Surf, double *maxV,
                           double *minV, double *meanV);

/* TempFilterTimeDeriv (sponge/pruettdamping.f90:69-92), called once per time step after TimeStep (timedisc_func.f90:357) when
 * SpongeBaseFlow = pruett: SpBaseFlow += (U - SpBaseFlow) dt / tempFilterWidth. dgx_get_baseflow downloads it (the reference
 * writes it to the *_BaseFlow_* file). */
int dgx_temp_filter_time_deriv(dgx_handle *h, double dt, double tempFilterWidth);
int dgx_get_baseflow(dgx_handle *h, double *SpBaseFlow);

/* Face arrays of the last RHS evaluation in the reference's layout (nVar,0:N,0:N,nSides): which = 0 U_master, 1 U_slave,
 * 2 Flux_master (5 variables; dg_vars.f90:62-75), 3..5 gradUx/y/z_master, 6..8 gradUx/y/z_slave (the 4 lifted variables u, v, w,
 * T; lifting_vars.f90:43-60). For parity checks of orientation and sign conventions; not on the hot path. */
int dgx_get_face_array(dgx_handle *h, int which, double *out);
int dgx_sync(dgx_handle *h);
/* nSteps RK time steps with the state resident in HBM: the body of the reference's time loop between two analyze points
 * (timedisc/timedisc.f90:176-200: CalcForcing, CalcTimeStep, TimeStep). flags:
 *   1  adaptive dt: CalcTimeStep before every step (otherwise the fixed dt given);
 *   2  CalcForcing before every step (channel testcase, testcase.f90:241-271) with the weights / volume of the last
 *      dgx_calc_bulk_velocity call; the bulk velocity feeds TestcaseSource;
 *   4  device-paced: dt (and the bulk velocity) never leave the device -- stage 1 of every step evaluates CalcTimeStep inside
 *      its lifting kernel, the min over the ranks is an NCCL all-reduce on the communication stream and the stage kernels form
 *      b_dt = RKb * dt themselves; no host synchronisation inside the call. Needs flag 1, a time-independent source and a
 *      2-register scheme. An inadmissible state (calctimestep.f90:134-146) is reported when the call returns. The step sizes
 *      are read back with dgx_get_dt_history (t advances by their sum, added in step order like the reference's t = t + dt);
 *   8  with 4: pairs of steps are replayed as one CUDA graph (dgx_step_graph_active tells whether capture succeeded).
 * Results are bit-identical for every combination of 4 and 8. ms: device time of the call (CUDA events on the launching
 * stream); launches: kernels launched (graph replays count the kernels they contain). */
int dgx_run_steps(dgx_handle *h, int nSteps, double t, double dt, int flags, float *ms, long long *launches);
/* the dt of every step of the last device-paced dgx_run_steps: min(cap, count) values into dts, *count = number of steps */
int dgx_get_dt_history(dgx_handle *h, int cap, double *dts, int *count);
int dgx_step_graph_active(const dgx_handle *h);
/* per-kernel CUDA-event timing of one RK stage: names[i] -> ms[i], *count entries (<= cap) */
int dgx_profile_stage(dgx_handle *h, double t, double dt, int cap, const char **names, float *ms, int *count);
int dgx_nccl_unique_id(char *out128);
/* The face-halo message plan the library executes with NCCL for one exchange phase (no GPU needed): writes up to
 * cap records {peer, isSend, slaveArray, firstSide (0-based), nSides} in issue order, returns the number of messages.
 * Replaces the loop bodies of StartReceiveMPIData/StartSendMPIData (mpi/mpi.f90:277-387). */
int dgx_halo_plan(int nNbProcs, const int *NbProc, const int *nMPISides_MINE_Proc, const int *nMPISides_YOUR_Proc,
                  const int *offsetMPISides_MINE, const int *offsetMPISides_YOUR, int cap, int *out);
long long dgx_launch_count(const dgx_handle *h);
/* sizeof(dgx_config) as compiled: lets a foreign-language binding verify its struct mirror */
unsigned long dgx_sizeof_config(void);

#ifdef __cplusplus
}
#endif
#endif /* DGX_H */
