// grb_builtin_data.h -- built-in copies of the numerical inputs the path needs when no GRANIITTI tree is
// given (graniitti_base == NULL): the TUNE0 model parameters (values of reference
// modeldata/TUNE0/GENERAL.json, v1.051) and the subset of the PDG-2018 mass/width table the BASELINE
// configurations touch (values of reference modeldata/mass_width_2018.mcd).  When graniitti_base is given
// the host layer reads the reference's own files instead and these are unused; tests/test_host_parser.py
// checks that both routes flatten to the same grb_process_desc.
#pragma once

namespace grbhost {

// Same block/key names as GENERAL.json so that one loader serves both.
static const char *const kBuiltinTune0 = R"JSON({
"PARAM_SOFT": {"DELTA_P": 0.098, "ALPHA_P": 0.065, "gN_P": 8.20, "gN_O": 0.10, "g3P": 0.15, "gamma": 0.45,
               "fc1": 0.34, "fc2": 5.0, "fc3": 10.0, "ODDERON_ON": false},
"PARAM_SPIN": {"SPINGEN": true, "SPINDEC": true, "FRAME": "CS", "JMAX": 5},
"PARAM_FLAT": {"b": 6},
"PARAM_REGGE": {"a0": [1.088, 0.55, 0.55], "ap": [0.20, 0.93, 0.93], "sgn": [1, -1, 1], "s0": 1.0,
                "reggeize": false, "omega": 0.5, "offshellFF": 2,
                "b_EXP": 0.944, "a_OREAR": 0.638, "b_OREAR": 1.886, "b_POW": 0.500},
"PARAM_PI":    {"c": [35.00184, 81.63672, 10.86264], "n": [true, false, false]},
"PARAM_K":     {"c": [30.35376, 44.31084, 23.38164], "n": [true, false, false]},
"PARAM_P":     {"c": [27.0, 0, 0], "n": [true, false, false]},
"PARAM_RHO":   {"c": [25.0, 0, 0], "n": [true, false, false]},
"PARAM_PHI":   {"c": [18.0, 0, 0], "n": [true, false, false]},
"PARAM_F0500": {"c": [35.0, 0, 0], "n": [true, false, false]},
"PARAM_X":     {"c": [15.0, 0, 0], "n": [true, false, false]},
"PARAM_TENSORPOM": {"delta_P": 0.0808, "delta_O": 0.0808, "delta_1R": -0.5475, "delta_2R": -0.5475,
                    "ap_P": 0.25, "ap_O": 0.25, "ap_1R": 0.9, "ap_2R": 0.9, "eta_O": 1.0,
                    "gPNN": 1.87, "gPpipi": 1.76, "gPKK": 1.54, "gPrhorho": [0.70, 6.20],
                    "gPphiphi": [0.49, 4.27], "M_O": 1.00, "M_1R": 1.41},
"PARAM_STRUCTURE": {"F2": "CKMT", "EM": "KELLY", "QED_alpha": "ZERO"}
})JSON";

// Resonance parameter sets (values of reference modeldata/TUNE0/RES_<name>.json; same keys).
static const struct {
  const char *name, *json;
} kBuiltinRes[] = {
    {"f0_500", R"J({"PARAM_RES":{"PDG":9000221,"M":0.5,"W":0.4,"J":0,"P":1,"BW":2,"g_Tensor":[0.1,0.0],"random_rho":0,"g_A":0.05,"g_phi":-3.14159,"g_FF":1.0,"rho_real":[[0]],"rho_imag":[[0]]}})J"},
    {"f0_980", R"J({"PARAM_RES":{"PDG":9010221,"M":0.98,"W":0.08,"J":0,"P":1,"BW":2,"g_Tensor":[0.4,2.0],"random_rho":0,"g_A":0.065,"g_phi":-3.14159,"g_FF":1.0,"rho_real":[[0]],"rho_imag":[[0]]}})J"},
    {"f0_1500", R"J({"PARAM_RES":{"PDG":9030221,"M":1.506,"W":0.112,"J":0,"P":1,"BW":2,"g_Tensor":[0.1,0.0],"random_rho":0,"g_A":0.03,"g_phi":-3.14159,"g_FF":1.0,"rho_real":[[0]],"rho_imag":[[0]]}})J"},
    {"f0_1710", R"J({"PARAM_RES":{"PDG":10331,"M":1.704,"W":0.123,"J":0,"P":1,"BW":2,"g_Tensor":[0.45,2.6],"random_rho":0,"g_A":0.03,"g_phi":-3.14159,"g_FF":1.0,"rho_real":[[0]],"rho_imag":[[0]]}})J"},
    {"rho_770", R"J({"PARAM_RES":{"PDG":113,"M":0.77526,"W":0.1491,"J":1,"P":-1,"BW":2,"g_Tensor":[0.7,6.2],"random_rho":0,"g_A":0.19,"g_phi":-2.36,"g_FF":1.0,"rho_real":[[0.5,0,0],[0,0,0],[0,0,0.5]],"rho_imag":[[0,0,0],[0,0,0],[0,0,0]]}})J"},
    {"phi_1020", R"J({"PARAM_RES":{"PDG":333,"M":1.019461,"W":0.004249,"J":1,"P":-1,"BW":2,"g_Tensor":[0.49,4.27],"random_rho":0,"g_A":0.165,"g_phi":-2.36,"g_FF":1.0,"rho_real":[[0.5,0,0],[0,0,0],[0,0,0.5]],"rho_imag":[[0,0,0],[0,0,0],[0,0,0]]}})J"},
    {"f2_1270", R"J({"PARAM_RES":{"PDG":225,"M":1.2755,"W":0.1867,"J":2,"P":1,"BW":2,"g_Tensor":[2.0,0,0,0,0,0,0],"random_rho":0,"g_A":0.7,"g_phi":-1.57,"g_FF":1.0,"rho_real":[[0.5,0,0,0,0],[0,0,0,0,0],[0,0,0,0,0],[0,0,0,0,0],[0,0,0,0,0.5]],"rho_imag":[[0,0,0,0,0],[0,0,0,0,0],[0,0,0,0,0],[0,0,0,0,0],[0,0,0,0,0]]}})J"},
    {"f2_1525", R"J({"PARAM_RES":{"PDG":335,"M":1.525,"W":0.073,"J":2,"P":1,"BW":2,"g_Tensor":[0.0,0.5,0,0,0,0,0],"random_rho":0,"g_A":0.19,"g_phi":-1.57,"g_FF":1.0,"rho_real":[[0.5,0,0,0,0],[0,0,0,0,0],[0,0,0,0,0],[0,0,0,0,0],[0,0,0,0,0.5]],"rho_imag":[[0,0,0,0,0],[0,0,0,0,0],[0,0,0,0,0],[0,0,0,0,0],[0,0,0,0,0]]}})J"},
    {"f2_1950", R"J({"PARAM_RES":{"PDG":9050225,"M":1.936,"W":0.464,"J":2,"P":1,"BW":2,"g_Tensor":[1.0,0.0,0,0,0,0,0],"random_rho":0,"g_A":0.1,"g_phi":-1.57,"g_FF":1.0,"rho_real":[[0.5,0,0,0,0],[0,0,0,0,0],[0,0,0,0,0],[0,0,0,0,0],[0,0,0,0,0.5]],"rho_imag":[[0,0,0,0,0],[0,0,0,0,0],[0,0,0,0,0],[0,0,0,0,0],[0,0,0,0,0]]}})J"},
};

// Decay table subset (values of reference modeldata/TUNE0/BRANCHING.json; same layout): the resonances above
// into pseudoscalar pairs.
static const char *const kBuiltinBranching = R"JSON({
"0":{"0":{"PDG":[-211,211],"BR":1.0,"alpha_ls":[[0,0,1.0,0.0]],"P_conservation":true},"1":{"PDG":[-321,321],"BR":1.0,"alpha_ls":[[0,0,1.0,0.0]],"P_conservation":true}},
"9000221":{"0":{"PDG":[-211,211],"BR":1.0,"alpha_ls":[[0,0,1.0,0.0]],"P_conservation":true},"1":{"PDG":[-321,321],"BR":0.01,"alpha_ls":[[0,0,1.0,0.0]],"P_conservation":true}},
"113":{"0":{"PDG":[-211,211],"BR":1.0,"alpha_ls":[[1,0,1.0,0.0]],"P_conservation":true},"1":{"PDG":[321,-321],"BR":0.0,"alpha_ls":[[1,0,1.0,0.0]],"P_conservation":true}},
"9010221":{"0":{"PDG":[211,-211],"BR":0.5,"alpha_ls":[[0,0,1.0,0.0]],"P_conservation":true},"1":{"PDG":[321,-321],"BR":0.345,"alpha_ls":[[0,0,1.0,0.0]],"P_conservation":true}},
"333":{"0":{"PDG":[321,-321],"BR":0.489,"alpha_ls":[[1,0,1.0,0.0]],"P_conservation":true},"1":{"PDG":[130,310],"BR":0.342,"alpha_ls":[[1,0,1.0,0.0]],"P_conservation":true},"2":{"PDG":[211,-211],"BR":7.3e-05,"alpha_ls":[[1,0,1.0,0.0]],"P_conservation":true}},
"225":{"0":{"PDG":[211,-211],"BR":0.562,"alpha_ls":[[2,0,1.0,0.0]],"P_conservation":true},"1":{"PDG":[321,-321],"BR":0.0227,"alpha_ls":[[2,0,1.0,0.0]],"P_conservation":true}},
"9030221":{"0":{"PDG":[211,-211],"BR":0.2299,"alpha_ls":[[0,0,1.0,0.0]],"P_conservation":true},"1":{"PDG":[321,-321],"BR":0.0425,"alpha_ls":[[0,0,1.0,0.0]],"P_conservation":true}},
"335":{"0":{"PDG":[321,-321],"BR":0.4435,"alpha_ls":[[2,0,1.0,0.0]],"P_conservation":true},"1":{"PDG":[211,-211],"BR":0.00546,"alpha_ls":[[2,0,1.0,0.0]],"P_conservation":true}},
"10331":{"0":{"PDG":[321,-321],"BR":0.18,"alpha_ls":[[0,0,1.0,0.0]],"P_conservation":true},"1":{"PDG":[211,-211],"BR":0.0552,"alpha_ls":[[0,0,1.0,0.0]],"P_conservation":true}},
"9050225":{"0":{"PDG":[211,-211],"BR":1.0,"alpha_ls":[[2,0,1.0,0.0]],"P_conservation":true}}
})JSON";

struct BuiltinParticle {
  int         pdg;
  double      mass, width;
  const char *name;     // PDG table name without charge
  const char *charges;  // charge states as in the .mcd file ("+", "0", "0,+", "-")
};

// PDG ids with two charge states ("0,+") list the neutral id first and the charged id in pdg2.
static const struct {
  int         pdg, pdg2;
  double      mass, width;
  const char *name, *charges;
} kBuiltinPDG[] = {
    {21, 0, 0.0, 0.0, "g", "0"},
    {22, 0, 0.0, 0.0, "gamma", "0"},
    {11, 0, 5.109989461E-04, 0.0, "e", "-"},
    {13, 0, 1.056583745E-01, 2.9959847E-19, "mu", "-"},
    {15, 0, 1.77686E+00, 2.267E-12, "tau", "-"},
    {211, 0, 1.3957061E-01, 2.5284E-17, "pi", "+"},
    {111, 0, 1.349770E-01, 7.73E-09, "pi", "0"},
    {221, 0, 5.47862E-01, 1.31E-06, "eta", "0"},
    {9000221, 0, 4.75E-01, 5.50E-01, "f(0)(500)", "0"},
    {113, 213, 7.7526E-01, 1.491E-01, "rho(770)", "0,+"},
    {223, 0, 7.8265E-01, 8.49E-03, "omega(782)", "0"},
    {331, 0, 9.5778E-01, 1.96E-04, "eta'(958)", "0"},
    {9010221, 0, 9.90E-01, 6.E-02, "f(0)(980)", "0"},
    {333, 0, 1.019461E+00, 4.249E-03, "phi(1020)", "0"},
    {225, 0, 1.2755E+00, 1.867E-01, "f(2)(1270)", "0"},
    {9030221, 0, 1.504E+00, 1.09E-01, "f(0)(1500)", "0"},
    {335, 0, 1.525E+00, 7.3E-02, "f(2)'(1525)", "0"},
    {10331, 0, 1.723E+00, 1.39E-01, "f(0)(1710)", "0"},
    {321, 0, 4.93677E-01, 5.317E-17, "K", "+"},
    {311, 0, 4.97611E-01, 0.0, "K", "0"},
    {310, 0, 4.97611E-01, 7.3508E-15, "K(S)", "0"},
    {130, 0, 4.97611E-01, 1.287E-17, "K(L)", "0"},
    {443, 0, 3.096900E+00, 9.29E-05, "J/psi(1S)", "0"},
    {2212, 0, 9.38272081E-01, 0.0, "p", "+"},
    {2112, 0, 9.39565413E-01, 7.478E-28, "n", "0"},
    {3122, 0, 1.115683E+00, 2.501E-15, "Lambda", "0"},
};

}  // namespace grbhost
