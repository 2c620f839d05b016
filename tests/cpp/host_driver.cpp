// Compiled-language host of the parity tests: reads a case dumped by tests/test_cpp_host.py (the arrays the Fortran host
// passes at init time + an initial state + oracle results), drives the library through include/dgx.hpp with the reference's
// call sequence (timedisc.f90:108-203) and prints the deviations as one JSON line.
//   usage: host_driver <case.bin> [--link-only]
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "dgx.hpp"

namespace {
struct Blob {
    std::vector<char> bytes;
    size_t count = 0;
};
// file format: repeated records  [int32 nameLen][name][int32 kind 0=int32,1=float64][int64 count][payload]
std::map<std::string, Blob> read_case(const char* path) {
    std::map<std::string, Blob> m;
    FILE* f = fopen(path, "rb");
    if (!f) throw dgx::Abort(std::string("cannot open ") + path);
    for (;;) {
        int32_t nl;
        if (fread(&nl, 4, 1, f) != 1) break;
        std::string name(nl, ' ');
        int32_t kind;
        int64_t cnt;
        if (fread(&name[0], 1, nl, f) != (size_t)nl || fread(&kind, 4, 1, f) != 1 || fread(&cnt, 8, 1, f) != 1) throw dgx::Abort("truncated case file");
        Blob b;
        b.count = (size_t)cnt;
        b.bytes.resize((size_t)cnt * (kind ? 8 : 4));
        if (cnt && fread(b.bytes.data(), 1, b.bytes.size(), f) != b.bytes.size()) throw dgx::Abort("truncated case file");
        m[name] = std::move(b);
    }
    fclose(f);
    return m;
}
double rel_l2(const std::vector<double>& a, const double* b) {
    double num = 0, den = 0;
    for (size_t i = 0; i < a.size(); i++) { num += (a[i] - b[i]) * (a[i] - b[i]); den += b[i] * b[i]; }
    return std::sqrt(num) / std::fmax(std::sqrt(den), 1e-300);
}
}  // namespace

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: host_driver <case.bin>\n"); return 2; }
    if (argc > 2 && !strcmp(argv[2], "--link-only")) { printf("{\"linked\": true, \"sizeof_config\": %lu}\n", dgx_sizeof_config()); return 0; }
    try {
        auto m = read_case(argv[1]);
        auto I = [&](const char* k) { return reinterpret_cast<const int*>(m.at(k).bytes.data()); };
        auto D = [&](const char* k) { return reinterpret_cast<const double*>(m.at(k).bytes.data()); };
        const int* sc = I("scalars_int");
        const double* sd = D("scalars_real");
        dgx_config c;
        memset(&c, 0, sizeof c);
        c.N = sc[0]; c.nodeType = sc[1]; c.splitDG = sc[2]; c.riemann = sc[3]; c.parabolic = sc[4]; c.viscLaw = sc[5];
        c.nElems = sc[6]; c.nSides = sc[7]; c.nBCSides = sc[8]; c.firstInnerSide = sc[9]; c.lastInnerSide = sc[10];
        c.firstMPISide_MINE = sc[11]; c.lastMPISide_MINE = sc[12]; c.firstMPISide_YOUR = sc[13]; c.lastMPISide_YOUR = sc[14];
        c.nRefState = sc[15]; c.nRKStages = sc[16]; c.lifting = sc[17];
        c.nMortarSides = sc[18]; c.firstMortarInnerSide = sc[19]; c.lastMortarInnerSide = sc[20]; c.firstMortarMPISide = sc[21]; c.lastMortarMPISide = sc[22];
        for (int k = 0; k < 8; k++) c.EOS_Vars[k] = sd[k];
        c.CFLScale = sd[8]; c.DFLScale = sd[9]; c.etaBR2 = sd[10]; c.etaBR2_wall = sd[11];
        c.RefStatePrim = D("RefStatePrim"); c.BCSides = I("BCSides");
        c.D_T = D("D_T"); c.D_Hat_T = D("D_Hat_T"); c.DVolSurf = D("DVolSurf"); c.L_Minus = D("L_Minus"); c.L_Plus = D("L_Plus");
        c.L_HatMinus = D("L_HatMinus"); c.L_HatPlus = D("L_HatPlus");
        c.ElemToSide = I("ElemToSide"); c.S2V2 = I("S2V2"); c.S2V2_inv = I("S2V2_inv");
        c.Metrics_fTilde = D("Metrics_fTilde"); c.Metrics_gTilde = D("Metrics_gTilde"); c.Metrics_hTilde = D("Metrics_hTilde"); c.sJ = D("sJ");
        c.NormVec = D("NormVec"); c.TangVec1 = D("TangVec1"); c.TangVec2 = D("TangVec2"); c.SurfElem = D("SurfElem");
        c.RKA = D("RKA"); c.RKb = D("RKb"); c.RKc = D("RKc");
        c.myRank = 0; c.nRanks = 1; c.nNbProcs = 0; c.device = 0;
        c.MortarType = I("MortarType"); c.MortarInfo = I("MortarInfo");
        c.M_0_1 = D("M_0_1"); c.M_0_2 = D("M_0_2"); c.M_1_0 = D("M_1_0"); c.M_2_0 = D("M_2_0");

        const size_t nU = m.at("U0").count;
        dgx::DG dg(c);
        dg.SetState(D("U0"));                      // timedisc.f90:108  d_U = U
        dg.DGTimeDerivative_weakForm(0.0);         // timedisc.f90:126
        std::vector<double> Ut(nU), U(nU);
        dg.GetUt(Ut.data());
        const double eUt = rel_l2(Ut, D("Ut_ref"));
        int errType = 0;
        const double dt0 = dg.CalcTimeStep(errType);
        const double tEnd = sd[12];
        long nsteps = dg.TimeDisc(0.0, tEnd);
        dg.GetState(U.data());
        const double eU = rel_l2(U, D("U_ref"));
        // error path: a null state pointer is reported, not crashed on
        bool aborted = false;
        try {
            dgx_config bad = c;
            bad.N = 42;
            dgx::DG nope(bad);
        } catch (const dgx::Abort& a) { aborted = std::string(a.what()).find("polynomial degree") != std::string::npos; }
        printf("{\"ut_rel_l2\": %.3e, \"u_rel_l2\": %.3e, \"dt0\": %.17g, \"dt0_ref\": %.17g, \"errType\": %d, \"nsteps\": %ld, \"nsteps_ref\": %d, \"abort_ok\": %s, \"launches\": %lld}\n",
               eUt, eU, dt0, sd[13], errType, nsteps, sc[23], aborted ? "true" : "false", dg.LaunchCount());
    } catch (const dgx::Abort& a) {
        printf("{\"abort\": \"%s\"}\n", a.what());
        return 1;
    }
    return 0;
}
