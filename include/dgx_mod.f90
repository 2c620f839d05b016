!==================================================================================================================================
! ISO_C_BINDING interface to libdgx.so (include/dgx.h): the module a GALAEXI maintainer adds under src/dg/ to route the hot path
! (DGTimeDerivative_weakForm + TimeStepByLSERKW2 + CalcTimeStep) through the B200-native library.
! Shipped as source: it cannot be compiled in the build container (no Fortran compiler); field order and kinds mirror
! `struct dgx_config` one to one (tests/test_abi.py checks the C side against the ctypes mirror with dgx_sizeof_config()).
!==================================================================================================================================
MODULE MOD_DGX
USE ISO_C_BINDING
IMPLICIT NONE
PRIVATE

TYPE, BIND(C), PUBLIC :: dgx_config
  INTEGER(C_INT) :: N, nodeType, splitDG, riemann, parabolic, viscLaw
  INTEGER(C_INT) :: nElems, nSides, nBCSides
  INTEGER(C_INT) :: firstInnerSide, lastInnerSide
  INTEGER(C_INT) :: firstMPISide_MINE, lastMPISide_MINE, firstMPISide_YOUR, lastMPISide_YOUR
  REAL(C_DOUBLE) :: EOS_Vars(8)
  INTEGER(C_INT) :: nRefState
  TYPE(C_PTR)    :: RefStatePrim, BCSides
  TYPE(C_PTR)    :: D_T, D_Hat_T, DVolSurf, L_Minus, L_Plus, L_HatMinus, L_HatPlus
  TYPE(C_PTR)    :: ElemToSide, S2V2, S2V2_inv
  TYPE(C_PTR)    :: Metrics_fTilde, Metrics_gTilde, Metrics_hTilde, sJ
  TYPE(C_PTR)    :: NormVec, TangVec1, TangVec2, SurfElem
  INTEGER(C_INT) :: nRKStages
  TYPE(C_PTR)    :: RKA, RKb, RKc
  REAL(C_DOUBLE) :: CFLScale, DFLScale
  INTEGER(C_INT) :: myRank, nRanks, nNbProcs
  TYPE(C_PTR)    :: NbProc, nMPISides_MINE_Proc, nMPISides_YOUR_Proc, offsetMPISides_MINE, offsetMPISides_YOUR
  TYPE(C_PTR)    :: ncclUniqueId
  INTEGER(C_INT) :: device
  ! lifting variant (PP_Lifting: 1 BR1, 2 BR2) and BR2 penalties (lifting.f90:86-91)
  INTEGER(C_INT) :: lifting
  REAL(C_DOUBLE) :: etaBR2, etaBR2_wall
  ! non-conforming interfaces (src/mortar, mesh.f90:271-283,318-322); nMortarSides=0: unused
  INTEGER(C_INT) :: nMortarSides, firstMortarInnerSide, lastMortarInnerSide, firstMortarMPISide, lastMortarMPISide
  TYPE(C_PTR)    :: MortarType, MortarInfo
  TYPE(C_PTR)    :: M_0_1, M_0_2, M_1_0, M_2_0
  TYPE(C_PTR)    :: FilterMat   ! filter.f90:203, C_NULL_PTR if FilterType=0
  INTEGER(C_INT) :: IniExactFunc ! source term selection (exactfunc.f90:665-926)
  REAL(C_DOUBLE) :: AdvVel(3)
  TYPE(C_PTR)    :: Elem_xGP
  INTEGER(C_INT) :: doWeakLifting, doConservativeLifting   ! lifting.f90:139-141
  TYPE(C_PTR)    :: SpongeMat, SpBaseFlow                  ! sponge.f90 (SpongeMat expanded to all elements), C_NULL_PTR: no sponge
  TYPE(C_PTR)    :: RKdelta, RKg1, RKg2, RKg3              ! TimeDiscType LSERKK3 (timedisc_vars.f90:464-773), C_NULL_PTR: LSERKW2
  INTEGER(C_INT) :: OverintegrationType, NUnder            ! overintegration_vars.f90:30-36
  TYPE(C_PTR)    :: OverintegrationMat, Vdm_N_NUnder, Vdm_NUnder_N, sJNUnder   ! overintegration_vars.f90:38-51
END TYPE dgx_config

TYPE(C_PTR), PUBLIC, SAVE :: dgx = C_NULL_PTR   !< the library handle (one per rank)

INTERFACE
  INTEGER(C_INT) FUNCTION dgx_create(h,cfg) BIND(C,NAME='dgx_create')
    IMPORT; TYPE(C_PTR),INTENT(OUT) :: h; TYPE(dgx_config),INTENT(IN) :: cfg
  END FUNCTION
  SUBROUTINE dgx_destroy(h) BIND(C,NAME='dgx_destroy')
    IMPORT; TYPE(C_PTR),VALUE :: h
  END SUBROUTINE
  TYPE(C_PTR) FUNCTION dgx_last_error(h) BIND(C,NAME='dgx_last_error')
    IMPORT; TYPE(C_PTR),VALUE :: h
  END FUNCTION
  INTEGER(C_INT) FUNCTION dgx_set_state(h,U) BIND(C,NAME='dgx_set_state')
    IMPORT; TYPE(C_PTR),VALUE :: h; REAL(C_DOUBLE),INTENT(IN) :: U(*)
  END FUNCTION
  INTEGER(C_INT) FUNCTION dgx_get_state(h,U) BIND(C,NAME='dgx_get_state')
    IMPORT; TYPE(C_PTR),VALUE :: h; REAL(C_DOUBLE),INTENT(OUT) :: U(*)
  END FUNCTION
  INTEGER(C_INT) FUNCTION dgx_get_ut(h,Ut) BIND(C,NAME='dgx_get_ut')
    IMPORT; TYPE(C_PTR),VALUE :: h; REAL(C_DOUBLE),INTENT(OUT) :: Ut(*)
  END FUNCTION
  INTEGER(C_INT) FUNCTION dgx_get_gradients(h,gx,gy,gz) BIND(C,NAME='dgx_get_gradients')
    IMPORT; TYPE(C_PTR),VALUE :: h; REAL(C_DOUBLE),INTENT(OUT) :: gx(*),gy(*),gz(*)
  END FUNCTION
  INTEGER(C_INT) FUNCTION dgx_set_keep_gradients(h,on) BIND(C,NAME='dgx_set_keep_gradients')
    IMPORT; TYPE(C_PTR),VALUE :: h; INTEGER(C_INT),VALUE :: on
  END FUNCTION
  INTEGER(C_INT) FUNCTION dgx_time_derivative(h,t) BIND(C,NAME='dgx_time_derivative')
    IMPORT; TYPE(C_PTR),VALUE :: h; REAL(C_DOUBLE),VALUE :: t
  END FUNCTION
  INTEGER(C_INT) FUNCTION dgx_rk_stage(h,iStage,t,dt) BIND(C,NAME='dgx_rk_stage')
    IMPORT; TYPE(C_PTR),VALUE :: h; INTEGER(C_INT),VALUE :: iStage; REAL(C_DOUBLE),VALUE :: t,dt
  END FUNCTION
  INTEGER(C_INT) FUNCTION dgx_rk_step(h,t,dt) BIND(C,NAME='dgx_rk_step')
    IMPORT; TYPE(C_PTR),VALUE :: h; REAL(C_DOUBLE),VALUE :: t,dt
  END FUNCTION
  INTEGER(C_INT) FUNCTION dgx_calc_timestep(h,dt,errType) BIND(C,NAME='dgx_calc_timestep')
    IMPORT; TYPE(C_PTR),VALUE :: h; REAL(C_DOUBLE),INTENT(OUT) :: dt; INTEGER(C_INT),INTENT(OUT) :: errType
  END FUNCTION
  INTEGER(C_INT) FUNCTION dgx_nccl_unique_id(id) BIND(C,NAME='dgx_nccl_unique_id')
    IMPORT; CHARACTER(KIND=C_CHAR),INTENT(OUT) :: id(128)
  END FUNCTION
  !> AnalyzeTestcase of the Taylor-Green vortex (testcase/taylorgreenvortex/testcase.f90:283-515) on the device
  INTEGER(C_INT) FUNCTION dgx_analyze_tgv(h,NAnalyze,Vdm_GaussN_NAnalyze,wAnalyze,Vol,rho0,out15) BIND(C,NAME='dgx_analyze_tgv')
    IMPORT; TYPE(C_PTR),VALUE :: h; INTEGER(C_INT),VALUE :: NAnalyze
    REAL(C_DOUBLE),INTENT(IN) :: Vdm_GaussN_NAnalyze(*),wAnalyze(*); REAL(C_DOUBLE),VALUE :: Vol,rho0
    REAL(C_DOUBLE),INTENT(OUT) :: out15(15)
  END FUNCTION
  !> CalcForcing / TestcaseSource of the channel testcase (testcase/channel/testcase.f90:241-296)
  INTEGER(C_INT) FUNCTION dgx_calc_bulk_velocity(h,wGP,Vol,BulkVel) BIND(C,NAME='dgx_calc_bulk_velocity')
    IMPORT; TYPE(C_PTR),VALUE :: h; REAL(C_DOUBLE),INTENT(IN) :: wGP(*); REAL(C_DOUBLE),VALUE :: Vol; REAL(C_DOUBLE),INTENT(OUT) :: BulkVel
  END FUNCTION
  !> CalcBodyForces (equations/navierstokes/calcbodyforces.f90:41-110): Fp, Fv(3,nBCs) of the wall boundary conditions
  INTEGER(C_INT) FUNCTION dgx_calc_body_forces(h,wGP,BC,nBCs,Fp,Fv) BIND(C,NAME='dgx_calc_body_forces')
    IMPORT; TYPE(C_PTR),VALUE :: h; REAL(C_DOUBLE),INTENT(IN) :: wGP(*); INTEGER(C_INT),INTENT(IN) :: BC(*); INTEGER(C_INT),VALUE :: nBCs
    REAL(C_DOUBLE),INTENT(OUT) :: Fp(3,*),Fv(3,*)
  END FUNCTION
  !> CalcWallVelocity (equations/navierstokes/analyze_equation.f90:435-499)
  INTEGER(C_INT) FUNCTION dgx_calc_wall_velocity(h,wGP,BC,nBCs,Surf,maxV,minV,meanV) BIND(C,NAME='dgx_calc_wall_velocity')
    IMPORT; TYPE(C_PTR),VALUE :: h; REAL(C_DOUBLE),INTENT(IN) :: wGP(*),Surf(*); INTEGER(C_INT),INTENT(IN) :: BC(*); INTEGER(C_INT),VALUE :: nBCs
    REAL(C_DOUBLE),INTENT(OUT) :: maxV(*),minV(*),meanV(*)
  END FUNCTION
  !> TempFilterTimeDeriv (sponge/pruettdamping.f90:69-92)
  INTEGER(C_INT) FUNCTION dgx_temp_filter_time_deriv(h,dt,tempFilterWidth) BIND(C,NAME='dgx_temp_filter_time_deriv')
    IMPORT; TYPE(C_PTR),VALUE :: h; REAL(C_DOUBLE),VALUE :: dt,tempFilterWidth
  END FUNCTION
  INTEGER(C_INT) FUNCTION dgx_get_baseflow(h,SpBaseFlow) BIND(C,NAME='dgx_get_baseflow')
    IMPORT; TYPE(C_PTR),VALUE :: h; REAL(C_DOUBLE),INTENT(OUT) :: SpBaseFlow(*)
  END FUNCTION
  INTEGER(C_INT) FUNCTION dgx_set_channel_forcing(h,on,dpdx,BulkVel) BIND(C,NAME='dgx_set_channel_forcing')
    IMPORT; TYPE(C_PTR),VALUE :: h; INTEGER(C_INT),VALUE :: on; REAL(C_DOUBLE),VALUE :: dpdx,BulkVel
  END FUNCTION
  !> the time loop between two analyze points on the device (timedisc.f90:176-200); flags: 1 adaptive dt, 2 CalcForcing per step,
  !> 4 device-paced (no host synchronisation), 8 CUDA-graph replay
  INTEGER(C_INT) FUNCTION dgx_run_steps(h,nSteps,t,dt,flags,ms,launches) BIND(C,NAME='dgx_run_steps')
    IMPORT; TYPE(C_PTR),VALUE :: h; INTEGER(C_INT),VALUE :: nSteps,flags; REAL(C_DOUBLE),VALUE :: t,dt
    REAL(C_FLOAT),INTENT(OUT) :: ms; INTEGER(C_LONG_LONG),INTENT(OUT) :: launches
  END FUNCTION
  INTEGER(C_INT) FUNCTION dgx_get_dt_history(h,cap,dts,count) BIND(C,NAME='dgx_get_dt_history')
    IMPORT; TYPE(C_PTR),VALUE :: h; INTEGER(C_INT),VALUE :: cap; REAL(C_DOUBLE),INTENT(OUT) :: dts(*); INTEGER(C_INT),INTENT(OUT) :: count
  END FUNCTION
  INTEGER(C_INT) FUNCTION dgx_step_graph_active(h) BIND(C,NAME='dgx_step_graph_active')
    IMPORT; TYPE(C_PTR),VALUE :: h
  END FUNCTION
  INTEGER(C_INT) FUNCTION dgx_get_face_array(h,which,arr) BIND(C,NAME='dgx_get_face_array')
    IMPORT; TYPE(C_PTR),VALUE :: h; INTEGER(C_INT),VALUE :: which; REAL(C_DOUBLE),INTENT(OUT) :: arr(*)
  END FUNCTION
  INTEGER(C_INT) FUNCTION dgx_sync(h) BIND(C,NAME='dgx_sync')
    IMPORT; TYPE(C_PTR),VALUE :: h
  END FUNCTION
END INTERFACE

PUBLIC :: dgx_run_steps,dgx_get_dt_history,dgx_step_graph_active,dgx_sync,dgx_get_face_array
PUBLIC :: dgx_create,dgx_destroy,dgx_last_error,dgx_set_state,dgx_get_state,dgx_get_ut,dgx_get_gradients
PUBLIC :: dgx_set_keep_gradients,dgx_time_derivative,dgx_rk_stage,dgx_rk_step,dgx_calc_timestep,dgx_nccl_unique_id
PUBLIC :: dgx_analyze_tgv,dgx_calc_bulk_velocity,dgx_set_channel_forcing,dgx_temp_filter_time_deriv,dgx_get_baseflow
PUBLIC :: DGX_Check

CONTAINS

!> non-zero return code -> CALL Abort(__STAMP__,message), the reference's error path (globals/globals.f90:175-221)
SUBROUTINE DGX_Check(rc,stamp_file,stamp_line)
USE MOD_Globals, ONLY: Abort
INTEGER(C_INT),INTENT(IN)   :: rc
CHARACTER(LEN=*),INTENT(IN) :: stamp_file
INTEGER,INTENT(IN)          :: stamp_line
CHARACTER(KIND=C_CHAR),POINTER :: msg(:)
CHARACTER(LEN=512)          :: text
INTEGER                     :: i
IF(rc.EQ.0) RETURN
CALL C_F_POINTER(dgx_last_error(dgx),msg,[512])
text=''
DO i=1,512
  IF(msg(i).EQ.C_NULL_CHAR) EXIT
  text(i:i)=msg(i)
END DO
CALL Abort(stamp_file,stamp_line,'',TRIM(text))
END SUBROUTINE DGX_Check

END MODULE MOD_DGX
