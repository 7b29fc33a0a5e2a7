// Rule-based greenhouse climate controller on the device.
//
// Follows RuleBasedController.predict (gl_gym/environments/baseline.py:68-227, proportional_control :226-227, settings
// gl_gym/configs/agents/rule_based.yml) for one env: u[6] from the state before the step, the weather row of the
// current timestep (all 10 columns) and the env clock -- what experiments/evaluate_baseline.py:21-23 does on the host
// between two step_raw_control calls.  Fused into the step kernels' prologue (control mode 2) so a rule-based rollout
// needs no host round trip.  Evaluated once per env-step: plain IEEE division and the accurate branch-free glg_exp_acc.
#pragma once
#include "glg_math.h"

#define GLG_NCTRL 29
enum GlgCtrl {  // order of configs/agents/rule_based.yml
    CT_LAMPS_ON, CT_LAMPS_OFF, CT_DAY_START, CT_DAY_STOP, CT_OFF_SUN, CT_RAD_LIMIT, CT_TSP_DAY, CT_TSP_NIGHT, CT_HEAT_CORR,
    CT_HEAT_DEAD, CT_CO2_DAY, CT_VENT_HEAT_PB, CT_RH_MAX, CT_MECH_PB, CT_VENT_RH_PB, CT_T_VENT_OFF, CT_VENT_COLD_PB,
    CT_TH_SP_DAY, CT_TH_SP_NIGHT, CT_TH_PB, CT_TH_DEAD, CT_TH_RH, CT_TH_RH_PB, CT_LAMP_EXTRA_HEAT, CT_BL_EXTRA_RH, CT_RHMAX2,
    CT_T_HEAT_BAND, CT_CO2_BAND, CT_USE_BL
};

// sigmoid proportional band, baseline.py:226-227
GLG_HD double glg_prop_ctrl(double pv, double sp, double pb, double minv, double maxv) {
    const double ln100 = 4.605170185988092;  // np.log(100)
    const double e = glg_exp_acc(-2.0 / pb * ln100 * (pv - sp - pb / 2.0));
    return minv + (maxv - minv) * (1.0 / (1.0 + e));
}
// 1 inside the open interval (lo, hi); an interval with lo > hi wraps around (baseline.py:76-77, :85-86)
GLG_HD double glg_window01(double lo, double hi, double v) {
    const bool inside = (lo < v) && (v < hi), wrap = (lo < v) || (v < hi);
    return ((lo <= hi) ? inside : wrap) ? 1.0 : 0.0;
}
template <class S>
GLG_HD void glg_rule_control(const S &s, const double *x, const double *d, double hod, double doy, double *u) {
    const double tAir = x[2];
    const double lamps_on = s[CT_LAMPS_ON], lamps_off = s[CT_LAMPS_OFF];
    const double lamp_doy = glg_window01(s[CT_DAY_START], s[CT_DAY_STOP], doy);
    const double below_limit = d[7] < s[CT_RAD_LIMIT] ? 1.0 : 0.0;
    const double lamp_no_cons = (d[0] < s[CT_OFF_SUN] ? 1.0 : 0.0) * below_limit * glg_window01(lamps_on, lamps_off, hod) * lamp_doy;
    const double sw_on = fmax(0.0, fmin(1.0, hod - lamps_on + 1));
    const double sw_off = fmax(0.0, fmin(1.0, lamps_off - hod + 1));
    const double both = lamps_on == lamps_off ? 0.0 : (lamps_on < lamps_off ? fmin(sw_on, sw_off) : fmax(sw_on, sw_off));
    const double is_day = fmax(both * below_limit * lamp_doy, d[8]);
    const double heat_sp = is_day * s[CT_TSP_DAY] + (1 - is_day) * s[CT_TSP_NIGHT] + s[CT_HEAT_CORR] * lamp_no_cons;
    const double heat_max = heat_sp + s[CT_HEAT_DEAD];
    const double co2_ppm = 1e6 * 8.3144598 * (tAir + 273.15) * (1e-6 * x[0]) / (101325 * 44.01e-3);  // utils.py:352-361
    const double rh_in = 100 * x[15] / (610.78 * glg_exp_acc(17.2694 * tAir / (tAir + 238.3)));         // utils.py:363-364 (unclipped)
    const double vent_heat = glg_prop_ctrl(tAir, heat_max, s[CT_VENT_HEAT_PB], 0, 1);
    const double vent_rh = glg_prop_ctrl(rh_in, s[CT_RH_MAX] + 0 * s[CT_MECH_PB], s[CT_VENT_RH_PB], 0, 1);
    const double vent_cold = glg_prop_ctrl(tAir, heat_sp - s[CT_T_VENT_OFF], s[CT_VENT_COLD_PB], 1, 0);
    const double th_sp = d[8] * s[CT_TH_SP_DAY] + (1 - d[8]) * s[CT_TH_SP_NIGHT];
    const double th_cold = glg_prop_ctrl(d[1], th_sp, s[CT_TH_PB], 0, 1);
    const double th_heat = glg_prop_ctrl(tAir, heat_sp + s[CT_TH_DEAD], -s[CT_TH_PB], 1, 0);
    const double th_rh = fmax(glg_prop_ctrl(rh_in, s[CT_RHMAX2] + s[CT_TH_RH], s[CT_TH_RH_PB], 1, 0), 1 - vent_cold);
    const double lamp_on = lamp_no_cons * glg_prop_ctrl(tAir, heat_max + s[CT_LAMP_EXTRA_HEAT], -0.5, 0, 1) * (d[9] + (1 - d[9])) *
                           fmax(glg_prop_ctrl(rh_in, s[CT_RHMAX2] + s[CT_BL_EXTRA_RH], -0.5, 0, 1), 1 - vent_cold);
    u[0] = glg_prop_ctrl(tAir, heat_sp, s[CT_T_HEAT_BAND], 0, 1);
    u[1] = glg_prop_ctrl(co2_ppm, is_day * s[CT_CO2_DAY], s[CT_CO2_BAND], 0, 1);
    u[2] = fmin(th_cold, fmax(th_heat, th_rh));
    u[3] = fmin(vent_cold, fmax(vent_heat, vent_rh));
    u[4] = lamp_on;
    u[5] = s[CT_USE_BL] * (1 - d[9]) * lamp_on;
}
