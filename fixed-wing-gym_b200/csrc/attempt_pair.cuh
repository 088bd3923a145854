// fw_attempt_pair_kernel — the fp64 dopri5 attempt kernel with TWO WARPS PER 32 AIRCRAFT (SURVEY §7 option (ii)).
//
// Why: one thread per aircraft needs ~250 registers (dynamics.cuh: y, stage state, f, the RHS temporaries), so only
// 8 warps fit an SM = 2 per scheduler, and the FP64 pipe idles on dependent-issue stalls (ncu, profiles/r2a_*: issue
// slots 52 % busy, pipe 49 %, "wait" stall 1.0 per issue).  Here a block of 64 threads serves 32 aircraft: lane i of
// warp T and lane i of warp L work on the SAME aircraft, each with its own instruction stream (no divergence) and half
// of the state, so a thread needs <= 128 registers and 16 warps fit an SM — the same 256 aircraft per SM (the K stages
// fill the shared memory either way) with twice the warps to hide latency, and a single aircraft's pass is shared by
// two warps (the tail of the kernel is a chain of dependent passes).
//
//   warp T ("lateral")      owns quaternion e0..e3, p, r, throttle + its rate, the position sums.  Per stage: airspeed
//                           (Va, beta, sin/cos of alpha and beta), side force / roll / yaw moment coefficients,
//                           d/dt of its components.
//   warp L ("longitudinal") owns q, u, v, w, both elevons + their rates.  Per stage: alpha, the stall blending sigma,
//                           lift / drag / pitch moment, the force rotation, gravity, propulsion, d/dt of its components.
//
// Per stage two block barriers: (1) after the stage state is formed each warp publishes the conditioned values the
// other one needs (12 doubles per aircraft); (2) T hands Va, 1/Va, sin/cos(alpha, beta), the side force and the
// beta-drag term to L.  The first exchange lives in the K slot the stage is about to fill (slot s is dead until the end
// of stage s; stage 6 uses slot 1: a_61 = e_1 = 0), so the block needs only 2.9 KB on top of the 24 KB of K stages.
// At the end of the pass the two partial error norms and failure masks are exchanged once; every control decision
// (accept / reject, step-size factor, finish, fail) is then computed by BOTH warps from identical inputs with the same
// instructions, so they never disagree.  scipy's control flow (rk.py:111-176) is unchanged; results differ from the
// one-thread kernel only by the summation order of the error norm (<= 1 ulp).
//
// FwSpecShipped / FwSpecGeneric only (per-env model parameters, FwSpecRand, and fp32 keep fw_attempt_kernel).
#pragma once
#include "dynamics.cuh"

#define FW_PAIR_THREADS 64
#define FW_PAIR_MIN_BLOCKS 8
// Rows of the stage's scratch K slot.  Values sit in rows OWNED BY THE READER (fw_pair_kc): the writer of a K stage only
// ever overwrites its own rows, i.e. exchange values it has itself already consumed, so no barrier is needed between
// reading the exchange and storing the stage's derivatives.
enum { PX_E0 = 5, PX_E1 = 7, PX_E2 = 8, PX_E3 = 9, PX_P = 10, PX_R = 11, PX_TH = 13,      // T -> L, in L's rows
       PX_Q = 0, PX_U = 1, PX_V = 2, PX_W = 3, PX_AIL = 4 };                                // L -> T, in T's rows
enum { PH_VA = 0, PH_INVVA, PH_SA, PH_CA, PH_SB, PH_CB, PH_FY, PH_CDB, PH_N };
// shared memory of one block (doubles first): K stages, hand-over rows, error partials, failure words, adoption queue
#define FW_PAIR_SM_K 0
#define FW_PAIR_SM_H (6 * FW_N_KC * 32)
#define FW_PAIR_SM_E (FW_PAIR_SM_H + PH_N * 32)
#define FW_PAIR_SM_DOUBLES (FW_PAIR_SM_E + 2 * 32)
#define FW_PAIR_SMEM_BYTES (FW_PAIR_SM_DOUBLES * 8 + 2 * 32 * 4 + 36 * 4)

// K components (dynamics.cuh: kc 0..6 = e0..e3, p, q, r; 7..9 = u, v, w; 10..12 = elevon_l, elevon_r, throttle;
// 13..15 = their rates) owned by each role; local index i -> kc
__device__ __forceinline__ constexpr int fw_pair_kc(int role, int i) {
  return role == 0 ? (i < 5 ? i : (i == 5 ? 6 : (i == 6 ? 12 : 15)))          // T: e0 e1 e2 e3 p r th th_dot
                   : (i == 0 ? 5 : (i < 6 ? 6 + i : 7 + i));                  // L: q u v w el er el_dot er_dot
}

__device__ __forceinline__ void fw_pair_bar() { asm volatile("bar.sync 1, 64;" ::: "memory"); }
// hand-over barrier: the producer (warp T) arrives and runs on, the consumer (warp L) waits.  The rows it protects are
// rewritten only after the next full barrier (1), which L reaches after it has read them.
__device__ __forceinline__ void fw_pair_arrive() { asm volatile("bar.arrive 2, 64;" ::: "memory"); }
__device__ __forceinline__ void fw_pair_wait() { asm volatile("bar.sync 2, 64;" ::: "memory"); }

struct FwPairCtl {
  double t, h_abs;
  int status, rejected, attempts, accepted, fail;
};

#define FW_PK(slot, kc) Kl[((slot) * FW_N_KC + (kc)) * 32]
#define FW_PG(k) P.gammas[k]

// rate clip of actuator i (ControlVariable.apply_conditions)
template <class Spec>
__device__ __forceinline__ double fw_pair_rate(const fw_sim_t& P, int i, double x) {
  if ((Spec::clip >> (FW_R_AD0 + i)) & 1u) {
    const double m = P.act_has_dot_max[i] ? P.act_dot_max[i] : (double)CUDART_INF;
    x = x < -m ? -m : (x > m ? m : x);
  }
  return x;
}

// stage state of the 8 own components: y + h sum_j a_sj K_j (one fma per term, as fw_ivp_attempt)
template <int ROLE>
__device__ __forceinline__ void fw_pair_stage_state(const double* Kl, int s, double h, const double (&y)[8], double (&ys)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) ys[i] = y[i];
  for (int j = 0; j < s; ++j) {
    const double ha = h * c_dpA[s][j];
#pragma unroll
    for (int i = 0; i < 8; ++i) ys[i] = fma(ha, FW_PK(j, fw_pair_kc(ROLE, i)), ys[i]);
  }
}

// error-norm contribution of the 8 own components after stage 6 (rk.py:150-152)
// (f_new is parked in K slot 1, whose stage is dead after stage 5: fewer live registers across the pass)
template <int ROLE>
__device__ __forceinline__ double fw_pair_err(const fw_sim_t& P, const double* Kl, double h, const double (&y)[8],
                                              const double (&ys)[8]) {
  double se = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int kc = fw_pair_kc(ROLE, i);
    double e = c_dpE[6] * FW_PK(1, kc);
#pragma unroll
    for (int j = 0; j < 6; ++j)
      if (j != 1) e += c_dpE[j] * FW_PK(j, kc);
    e *= h;
    const double ay = fabs(y[i]), an = fabs(ys[i]);
    const double q = e * fw_rcp(P.atol + (ay > an ? ay : an) * P.rtol);
    se += q * q;
  }
  return se;
}

// ---- warp T: the six stages of one attempt --------------------------------------------------------------------------
// y: e0 e1 e2 e3 p r th th_dot; aux: gust u v w, angular gust p, r, throttle command
template <class Spec>
__device__ __forceinline__ void fw_pair_stages_T(const fw_sim_t& P, double* Kl, double* Hl, double h, const double (&y)[8],
                                                 const double (&ypos)[3], const double (&k0pos)[3], const double (&aux)[6],
                                                 const double (&wind)[3], double (&ys)[8], double (&yspos)[3],
                                                 double (&fpos)[3], double& se, uint32_t& fmask, int& fstage) {
  const FwPar<double, FW_PAR_CONST> PP{P, nullptr, 0, nullptr};
  double accB[3], accE[3], f[8];
#pragma unroll
  for (int j = 0; j < 3; ++j) { accB[j] = c_dpA[6][0] * k0pos[j]; accE[j] = c_dpE[0] * k0pos[j]; }
#pragma unroll 1
  for (int s = 1; s <= 6; ++s) {
    fw_pair_stage_state<0>(Kl, s, h, y, ys);
    uint32_t fm = 0u;
    const double e0 = ys[0], e1 = ys[1], e2 = ys[2], e3 = ys[3];
    const double p = fw_cond_r<double, Spec, FW_R_P>(P, P.var[FW_SV_OMEGA_P], ys[4], fm);
    const double r = fw_cond_r<double, Spec, FW_R_R>(P, P.var[FW_SV_OMEGA_R], ys[5], fm);
    const double th = fw_cond_r<double, Spec, FW_R_TH>(P, P.var[FW_SV_THROTTLE], ys[6], fm);
    const double thd = fw_pair_rate<Spec>(P, 2, ys[7]);
    double* X = Kl + (s < 6 ? s : 1) * FW_N_KC * 32;
    X[PX_E0 * 32] = e0; X[PX_E1 * 32] = e1; X[PX_E2 * 32] = e2; X[PX_E3 * 32] = e3;
    X[PX_P * 32] = p; X[PX_R * 32] = r; X[PX_TH * 32] = th;
    fw_pair_bar();   // (1)
    const double q = X[PX_Q * 32], u = X[PX_U * 32], v = X[PX_V * 32], w = X[PX_W * 32], ail = X[PX_AIL * 32];
    // ---- airspeed factors (fw_rhs, same expressions) ----
    double ur = u, vr = v, wr = w;
    if (Spec::generic && P.wind_enabled) {
      const double wn = wind[0], we = wind[1], wd = wind[2];
      ur -= (-1.0 + 2 * (e0 * e0 + e1 * e1)) * wn + 2 * (e1 * e2 + e3 * e0) * we + 2 * (e1 * e3 - e2 * e0) * wd;
      vr -= 2 * (e1 * e2 - e3 * e0) * wn + (-1.0 + 2 * (e0 * e0 + e2 * e2)) * we + 2 * (e2 * e3 + e1 * e0) * wd;
      wr -= 2 * (e1 * e3 + e2 * e0) * wn + 2 * (e2 * e3 - e1 * e0) * we + (-1.0 + 2 * (e0 * e0 + e3 * e3)) * wd;
    }
    ur -= aux[0]; vr -= aux[1]; wr -= aux[2];
    const double pa = p - aux[3], ra = r - aux[4];
    const double hxz2 = ur * ur + wr * wr;
    double Va_raw, invVa_raw, hxz, ih;
    fwm_sqrt_rsqrt(hxz2 + vr * vr, &Va_raw, &invVa_raw);
    fwm_sqrt_rsqrt(hxz2, &hxz, &ih);
    double beta = fwm_atan2(vr, hxz);
    const double Va = fw_cond_r<double, Spec, FW_R_VA>(P, P.var[FW_SV_VA], Va_raw, fm);
    double invVa = invVa_raw;
    if (Va != Va_raw) invVa = fwm_rcp(Va);
    beta = fw_cond_r<double, Spec, FW_R_BETA>(P, P.var[FW_SV_BETA], beta, fm);
    const bool nz = hxz > 0.0;
    double sb, cb;
    if (!Spec::generic || (P.var[FW_SV_ALPHA].flags | P.var[FW_SV_BETA].flags) == 0u) {
      sb = vr * invVa_raw;
      cb = hxz * invVa_raw;
    } else {
      sincos(beta, &sb, &cb);
    }
    const double pre = 0.5 * PP.rho() * Va * Va * PP.S_wing();
    const double b2Va = PP.b() * (0.5 * invVa);
    const double rud = 0.0;
    const double CDb = PP.C_D_beta1() * beta + PP.C_D_beta2() * beta * beta;
    const double fy = pre * (PP.C_Y_0() + PP.C_Y_beta() * beta + PP.C_Y_p() * b2Va * pa + PP.C_Y_r() * b2Va * ra +
                             PP.C_Y_delta_a() * ail + PP.C_Y_delta_r() * rud);
    Hl[PH_VA * 32] = Va; Hl[PH_INVVA * 32] = invVa;
    Hl[PH_SA * 32] = nz ? wr * ih : 0.0; Hl[PH_CA * 32] = nz ? ur * ih : 1.0;
    Hl[PH_SB * 32] = sb; Hl[PH_CB * 32] = cb; Hl[PH_FY * 32] = fy; Hl[PH_CDB * 32] = CDb;
    fw_pair_arrive();   // (2)
    const double ll = pre * PP.b() * (PP.C_l_0() + PP.C_l_beta() * beta + PP.C_l_p() * b2Va * pa + PP.C_l_r() * b2Va * ra +
                                      PP.C_l_delta_a() * ail + PP.C_l_delta_r() * rud);
    const double nn = pre * PP.b() * (PP.C_n_0() + PP.C_n_beta() * beta + PP.C_n_p() * b2Va * pa + PP.C_n_r() * b2Va * ra +
                                      PP.C_n_delta_a() * ail + PP.C_n_delta_r() * rud);
    const double kot = PP.k_Omega() * th;
    const double tprop = -PP.k_T_P() * kot * kot;
    const double tl = ll + tprop, tn = nn;
    f[0] = 0.5 * (-p * e1 - q * e2 - r * e3);
    f[1] = 0.5 * (p * e0 + r * e2 - q * e3);
    f[2] = 0.5 * (q * e0 - r * e1 + p * e3);
    f[3] = 0.5 * (r * e0 + q * e1 - p * e2);
    f[4] = FW_PG(1) * p * q - FW_PG(2) * q * r + FW_PG(3) * tl + FW_PG(4) * tn;
    f[5] = FW_PG(7) * p * q - FW_PG(1) * q * r + FW_PG(4) * tl + FW_PG(8) * tn;
    fpos[0] = (e1 * e1 + e0 * e0 - e2 * e2 - e3 * e3) * u + 2 * (e1 * e2 - e3 * e0) * v + 2 * (e1 * e3 + e2 * e0) * w;
    fpos[1] = 2 * (e1 * e2 + e3 * e0) * u + (e2 * e2 + e0 * e0 - e1 * e1 - e3 * e3) * v + 2 * (e2 * e3 - e1 * e0) * w;
    fpos[2] = 2 * (e1 * e3 - e2 * e0) * u + 2 * (e2 * e3 + e1 * e0) * v + (e3 * e3 + e0 * e0 - e1 * e1 - e2 * e2) * w;
    f[6] = th * P.act_coef[2][0] + aux[5] * P.act_coef[2][2] + thd * P.act_coef[2][1];
    f[7] = th * P.act_coef[2][3] + aux[5] * P.act_coef[2][5] + thd * P.act_coef[2][4];
    if (fm && !fstage) { fstage = s; fmask = fm; }
    {
      double* Ks = Kl + (s < 6 ? s : 1) * FW_N_KC * 32;
#pragma unroll
      for (int i = 0; i < 8; ++i) Ks[fw_pair_kc(0, i) * 32] = f[i];
    }
    if (s < 6) {
      const double b = c_dpA[6][s], e = c_dpE[s];
#pragma unroll
      for (int j = 0; j < 3; ++j) { accB[j] += b * fpos[j]; accE[j] += e * fpos[j]; }
    }
  }
  se = fw_pair_err<0>(P, Kl, h, y, ys);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    yspos[j] = ypos[j] + h * accB[j];   // y_new of the position components: the B row is complete after stage 5
    const double e = (accE[j] + c_dpE[6] * fpos[j]) * h;
    const double ay = fabs(ypos[j]), an = fabs(yspos[j]);
    const double q = e * fw_rcp(P.atol + (ay > an ? ay : an) * P.rtol);
    se += q * q;
  }
}

// ---- warp L -----------------------------------------------------------------------------------------------------------
// y: q u v w el er el_dot er_dot; aux: gust u, w, angular gust q, elevon commands l, r
template <class Spec>
__device__ __forceinline__ void fw_pair_stages_L(const fw_sim_t& P, double* Kl, const double* Hl, double h, const double (&y)[8],
                                                 const double (&aux)[6], const double (&wind)[3], double (&ys)[8],
                                                 double& se, uint32_t& fmask, int& fstage) {
  const FwPar<double, FW_PAR_CONST> PP{P, nullptr, 0, nullptr};
  double f[8];
#pragma unroll 1
  for (int s = 1; s <= 6; ++s) {
    fw_pair_stage_state<1>(Kl, s, h, y, ys);
    uint32_t fm = 0u;
    const double q = fw_cond_r<double, Spec, FW_R_Q>(P, P.var[FW_SV_OMEGA_Q], ys[0], fm);
    const double u = fw_cond_r<double, Spec, FW_R_U>(P, P.var[FW_SV_VEL_U], ys[1], fm);
    const double v = fw_cond_r<double, Spec, FW_R_V>(P, P.var[FW_SV_VEL_V], ys[2], fm);
    const double w = fw_cond_r<double, Spec, FW_R_W>(P, P.var[FW_SV_VEL_W], ys[3], fm);
    const double el = fw_cond_r<double, Spec, FW_R_EL>(P, P.var[FW_SV_ELEVON_L], ys[4], fm);
    const double er = fw_cond_r<double, Spec, FW_R_ER>(P, P.var[FW_SV_ELEVON_R], ys[5], fm);
    const double eld = fw_pair_rate<Spec>(P, 0, ys[6]), erd = fw_pair_rate<Spec>(P, 1, ys[7]);
    const double ail = fw_cond_r<double, Spec, FW_R_AIL>(P, P.var[FW_SV_AILERON], (-er + el) * 0.5, fm);
    const double elev = fw_cond_r<double, Spec, FW_R_ELEV>(P, P.var[FW_SV_ELEVATOR], (er + el) * 0.5, fm);
    double* X = Kl + (s < 6 ? s : 1) * FW_N_KC * 32;
    X[PX_Q * 32] = q; X[PX_U * 32] = u; X[PX_V * 32] = v; X[PX_W * 32] = w; X[PX_AIL * 32] = ail;
    fw_pair_bar();   // (1)
    const double e0 = X[PX_E0 * 32], e1 = X[PX_E1 * 32], e2 = X[PX_E2 * 32], e3 = X[PX_E3 * 32];
    const double p = X[PX_P * 32], r = X[PX_R * 32], th = X[PX_TH * 32];
    double ur = u, wr = w;
    if (Spec::generic && P.wind_enabled) {
      const double wn = wind[0], we = wind[1], wd = wind[2];
      ur -= (-1.0 + 2 * (e0 * e0 + e1 * e1)) * wn + 2 * (e1 * e2 + e3 * e0) * we + 2 * (e1 * e3 - e2 * e0) * wd;
      wr -= 2 * (e1 * e3 + e2 * e0) * wn + 2 * (e2 * e3 - e1 * e0) * we + (-1.0 + 2 * (e0 * e0 + e3 * e3)) * wd;
    }
    ur -= aux[0]; wr -= aux[1];
    const double qa = q - aux[2];
    double alpha = fwm_atan2(wr, ur);
    alpha = fw_cond_r<double, Spec, FW_R_ALPHA>(P, P.var[FW_SV_ALPHA], alpha, fm);
    const double CLlin = PP.C_L_0() + PP.C_L_alpha() * alpha;
    const double x1 = fwm_exp(-PP.M() * (alpha - PP.a_0()));
    const double C = PP.exp_2Ma0();
    const double sigma = fwm_div(fma(x1, x1, x1) + C, (1 + x1) * (x1 + C));
    const double sgn = alpha > 0 ? 1.0 : (alpha < 0 ? -1.0 : 0.0);
    const double mg = PP.mass() * PP.g();
    const double fgx = mg * (2 * (e1 * e3 - e2 * e0));
    const double fgy = mg * (2 * (e2 * e3 + e1 * e0));
    const double fgz = mg * (e3 * e3 + e0 * e0 - e1 * e1 - e2 * e2);
    double* Ks = Kl + (s < 6 ? s : 1) * FW_N_KC * 32;   // (its rows 10, 11, 13, 14 held exchange values read above)
    Ks[fw_pair_kc(1, 4) * 32] = el * P.act_coef[0][0] + aux[3] * P.act_coef[0][2] + eld * P.act_coef[0][1];
    Ks[fw_pair_kc(1, 5) * 32] = er * P.act_coef[1][0] + aux[4] * P.act_coef[1][2] + erd * P.act_coef[1][1];
    Ks[fw_pair_kc(1, 6) * 32] = el * P.act_coef[0][3] + aux[3] * P.act_coef[0][5] + eld * P.act_coef[0][4];
    Ks[fw_pair_kc(1, 7) * 32] = er * P.act_coef[1][3] + aux[4] * P.act_coef[1][5] + erd * P.act_coef[1][4];
    fw_pair_wait();   // (2)
    const double Va = Hl[PH_VA * 32], invVa = Hl[PH_INVVA * 32];
    double sa = Hl[PH_SA * 32], ca = Hl[PH_CA * 32];
    const double sb = Hl[PH_SB * 32], cb = Hl[PH_CB * 32], fy = Hl[PH_FY * 32], CDb = Hl[PH_CDB * 32];
    if (Spec::generic && (P.var[FW_SV_ALPHA].flags | P.var[FW_SV_BETA].flags) != 0u) sincos(alpha, &sa, &ca);
    const double pre = 0.5 * PP.rho() * Va * Va * PP.S_wing();
    const double inv2Va = 0.5 * invVa;
    const double c2Va = PP.c() * inv2Va, b2Va = PP.b() * inv2Va;
    const double CL = (1 - sigma) * CLlin + sigma * (2 * sgn * sa * sa * ca);
    const double lift = pre * (CL + PP.C_L_q() * c2Va * qa + PP.C_L_delta_e() * elev);
    double CDa;
    if (!Spec::generic || P.drag_model == 0)
      CDa = PP.C_D_p() + (1 - sigma) * CLlin * CLlin * PP.inv_pi_e_ar() + sigma * (2 * sgn * sa * sa * sa);
    else
      CDa = PP.C_D_0() + PP.C_D_alpha1() * alpha + PP.C_D_alpha2() * alpha * alpha;
    const double drag = pre * (CDa + CDb + PP.C_D_q() * c2Va * qa + PP.C_D_delta_e() * elev * elev);
    const double Cm = (1 - sigma) * (PP.C_m_0() + PP.C_m_alpha() * alpha) + sigma * (PP.C_m_fp() * sgn * sa * sa);
    const double mm = pre * PP.c() * (Cm + PP.C_m_q() * b2Va * qa + PP.C_m_delta_e() * elev);
    const double fax = ca * cb * (-drag) + ca * sb * fy + sa * lift;
    const double fay = sb * drag + cb * fy;
    const double faz = sa * cb * (-drag) + sa * sb * fy - ca * lift;
    const double Vd = Va + th * (PP.k_motor() - Va);
    const double fprop = 0.5 * PP.rho() * PP.S_prop() * PP.C_prop() * Vd * (Vd - Va);
    const double fx = fprop + fgx + fax, fyy = fgy + fay, fz = fgz + faz;
    const double im = PP.inv_mass();
    f[0] = FW_PG(5) * p * r - FW_PG(6) * (p * p - r * r) + mm * P.inv_Jy;
    f[1] = r * v - q * w + fx * im;
    f[2] = p * w - r * u + fyy * im;
    f[3] = q * u - p * v + fz * im;
    if (fm && !fstage) { fstage = s; fmask = fm; }
#pragma unroll
    for (int i = 0; i < 4; ++i) Ks[fw_pair_kc(1, i) * 32] = f[i];
  }
  se = fw_pair_err<1>(P, Kl, h, y, ys);
}

template <class Spec>
__global__ void __launch_bounds__(FW_PAIR_THREADS, FW_PAIR_MIN_BLOCKS)
fw_attempt_pair_kernel(const __grid_constant__ fw_sim_t P, const FwDynArgs a) {
  FW_TL_BEGIN(1);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* smd = reinterpret_cast<double*>(smem_raw);
  uint32_t* smFM = reinterpret_cast<uint32_t*>(smd + FW_PAIR_SM_DOUBLES);   // [2][32]
  int32_t* smQ = reinterpret_cast<int32_t*>(smFM + 64);                      // [32] adopted env, [32..33] queue flags
  const int lane = threadIdx.x & 31;
  // Which warp plays T: a block's two warps sit on neighbouring schedulers (warp slot mod 4), so a fixed assignment
  // would give two schedulers of every SM only T code and the other two only L code; alternate per pair of blocks.
  if (threadIdx.x == 0) {
    unsigned wid;
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
    smQ[34] = a.pair_mix ? (int)((wid >> 2) & 1u) : 0;
  }
  __syncthreads();
  const int role = (int)(threadIdx.x >> 5) ^ smQ[34];
  double* Kl = smd + FW_PAIR_SM_K + lane;
  double* Hl = smd + FW_PAIR_SM_H + lane;
  double* El = smd + FW_PAIR_SM_E + lane;
  if (a.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  const int n_long = a.q[Q_LONG_COUNT];
  asm volatile("griddepcontrol.launch_dependents;" :: "r"(n_long) : "memory");
  const unsigned full = 0xffffffffu;
  const int n_nat = (int)a.n;
  FwPairCtl S;
  S.status = FW_STATUS_FINISHED; S.fail = 0; S.rejected = 0; S.attempts = 0; S.accepted = 0; S.t = 0; S.h_abs = 0;
  double y[8], ypos[3], k0pos[3], aux[6], wind[3];
#pragma unroll
  for (int j = 0; j < 8; ++j) y[j] = 0;
#pragma unroll
  for (int j = 0; j < 3; ++j) { ypos[j] = 0; k0pos[j] = 0; wind[j] = 0; }
#pragma unroll
  for (int j = 0; j < 6; ++j) aux[j] = 0;
  int64_t env = -1;
  bool long_left = n_long > 0, nat_left = true;
  unsigned long long passes = 0, lane_attempts = 0;
  for (;;) {
    // ---- adoption: warp T draws from the queues, both warps load their half of the aircraft ----
    const unsigned idle = __ballot_sync(full, S.status != FW_STATUS_RUNNING);
    if (idle && (long_left || nat_left)) {   // block-uniform: both warps hold identical control state
      if (role == 0) {
        const int want = __popc(idle);
        const int rank = __popc(idle & ((1u << lane) - 1u));
        int got_long = 0, got_nat = 0, base_long = 0, base_nat = 0;
        bool ll_ = long_left, nl_ = nat_left;
        if (ll_) {
          if (lane == 0) base_long = atomicAdd(a.q + Q_LONG_CURSOR, want);
          base_long = __shfl_sync(full, base_long, 0);
          got_long = min(want, max(0, n_long - base_long));
          if (base_long + want >= n_long) ll_ = false;
        }
        if (got_long < want && nl_) {
          const int need = want - got_long;
          if (lane == 0) base_nat = atomicAdd(a.q + Q_NAT_CURSOR, need);
          base_nat = __shfl_sync(full, base_nat, 0);
          got_nat = min(need, max(0, n_nat - base_nat));
          if (base_nat + need >= n_nat) nl_ = false;
        }
        int enc = -1;   // env | from_long << 30
        if (S.status != FW_STATUS_RUNNING) {
          if (rank < got_long) enc = a.long_list[base_long + rank] | (1 << 30);
          else if (rank - got_long < got_nat) {
            int e = base_nat + (rank - got_long);
            if (a.order) e = a.order[e];
            enc = e;
          }
        }
        smQ[lane] = enc;
        if (lane == 0) { smQ[32] = ll_ ? 1 : 0; smQ[33] = nl_ ? 1 : 0; }
      }
      __syncthreads();
      const int enc = smQ[lane];
      long_left = smQ[32] != 0;
      nat_left = smQ[33] != 0;
      if (S.status != FW_STATUS_RUNNING) {
        env = -1;
        if (enc >= 0) {
          const bool from_long = (enc >> 30) & 1;
          const int64_t e = enc & ((1 << 30) - 1);
          const double* cd = a.cd + e;
          const double h0s = cd[CY_H * a.stride];
          const int failv = a.ci[CI_FAIL * a.stride + e];
          const bool skip = failv != 0 || (!from_long && h0s < 0.0);
          if (!skip) {
            env = e;
            FwEnvCtx c{a.d, a.i, a.stride, e};
            const bool turb = P.turbulence != 0;
            if (role == 0) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                y[i] = c.D(fw_kc_to_ode(fw_pair_kc(0, i)));
                FW_PK(0, fw_pair_kc(0, i)) = cd[(CY_K0 + fw_pair_kc(0, i)) * a.stride];
              }
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                ypos[j] = c.D(D_POS + j);
                k0pos[j] = cd[(CY_KP + j) * a.stride];
                aux[j] = turb ? c.D(D_GUST + j) : 0.0;
              }
              aux[3] = turb ? c.D(D_GUST + 3) : 0.0;
              aux[4] = turb ? c.D(D_GUST + 5) : 0.0;
              aux[5] = cd[(CY_CMD + 2) * a.stride];
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                y[i] = c.D(fw_kc_to_ode(fw_pair_kc(1, i)));
                FW_PK(0, fw_pair_kc(1, i)) = cd[(CY_K0 + fw_pair_kc(1, i)) * a.stride];
              }
              aux[0] = turb ? c.D(D_GUST + 0) : 0.0;
              aux[1] = turb ? c.D(D_GUST + 2) : 0.0;
              aux[2] = turb ? c.D(D_GUST + 4) : 0.0;
              aux[3] = cd[(CY_CMD + 0) * a.stride];
              aux[4] = cd[(CY_CMD + 1) * a.stride];
            }
            if (Spec::generic) {
#pragma unroll
              for (int j = 0; j < 3; ++j) wind[j] = P.wind_enabled ? c.D(D_WIND + j) : 0.0;
            }
            S.t = 0; S.h_abs = fabs(h0s); S.rejected = 0; S.attempts = 0; S.accepted = 0; S.fail = 0;
            S.status = FW_STATUS_RUNNING;
          }
        }
      }
      __syncthreads();   // smQ is free again
    }
    const unsigned running = __ballot_sync(full, S.status == FW_STATUS_RUNNING);
    if (!running) {
      if (!long_left && !nat_left) break;
      continue;
    }
    if (role == 0) { ++passes; if (S.status == FW_STATUS_RUNNING) ++lane_attempts; }
    // ---- one dopri5 step attempt (rk.py:111-176); lanes without an aircraft run along (the barriers need them) ----
    const bool was_running = S.status == FW_STATUS_RUNNING;
    bool act = was_running;
    const double tb = P.dt;
    const double min_step = 10 * fabs(FwMath<double>::next_up(S.t) - S.t);
    if (act && !S.rejected && S.h_abs < min_step) S.h_abs = min_step;
    if (act) {   // NaN step sizes and the attempt cap end the step as a NUMERIC failure (dynamics.cuh, fw_ivp_attempt)
      const int cap = P.max_attempts > 0 ? P.max_attempts : FW_MAX_ATTEMPTS;
      if (!(S.h_abs >= min_step) || S.attempts >= cap) {
        if (S.h_abs < min_step) S.status = FW_STATUS_TOO_SMALL;
        else { S.fail = FW_TERM_NUMERIC; S.status = FW_STATUS_FINISHED; }
        act = false;
      }
    }
    double t_new = S.t + S.h_abs;
    if (t_new - tb > 0) t_new = tb;
    const double h = act ? t_new - S.t : 0.0;
    if (act) { S.h_abs = fabs(h); ++S.attempts; }
    double ys[8], yspos[3], fpos[3], se = 0;
    uint32_t fmask = 0u;
    int fstage = 0;
    if (role == 0) fw_pair_stages_T<Spec>(P, Kl, Hl, h, y, ypos, k0pos, aux, wind, ys, yspos, fpos, se, fmask, fstage);
    else fw_pair_stages_L<Spec>(P, Kl, Hl, h, y, aux, wind, ys, se, fmask, fstage);
    El[role * 32] = se;
    smFM[role * 32 + lane] = ((uint32_t)fstage << 24) | fmask;
    fw_pair_bar();   // (3)
    // ---- control: identical instructions on identical inputs in both warps ----
    const double se_all = El[0] + El[32];
    const uint32_t w0 = smFM[lane], w1 = smFM[32 + lane];
    const uint32_t st0 = w0 >> 24, st1 = w1 >> 24;
    uint32_t fm_all = 0u;
    if (st0 | st1) {   // the earliest failing stage names the exception; within a stage the first variable in PyFly's order
      const uint32_t first = st0 == 0u ? st1 : (st1 == 0u ? st0 : min(st0, st1));
      fm_all = (st0 == first ? (w0 & 0xffffffu) : 0u) | (st1 == first ? (w1 & 0xffffffu) : 0u);
    }
    if (act) {
      if (fm_all) {
        S.fail = fw_fail_code<double>(fm_all);
        S.status = FW_STATUS_FINISHED;
      } else {
        const double err = fwm_sqrt(se_all) * (1.0 / 4.358898943540674);
        const double pw = 0.9 * fwm_pow(err, -0.2);
        if (err < 1.0) {
          double factor = pw < 10.0 ? pw : 10.0;
          if (S.rejected) factor = factor < 1.0 ? factor : 1.0;
          S.h_abs *= factor;
          S.t = t_new;
#pragma unroll
          for (int i = 0; i < 8; ++i) y[i] = ys[i];
          if (role == 0) {   // FSAL: f_new (parked in slot 1) becomes K0
#pragma unroll
            for (int i = 0; i < 8; ++i) FW_PK(0, fw_pair_kc(0, i)) = FW_PK(1, fw_pair_kc(0, i));
#pragma unroll
            for (int j = 0; j < 3; ++j) { ypos[j] = yspos[j]; k0pos[j] = fpos[j]; }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) FW_PK(0, fw_pair_kc(1, i)) = FW_PK(1, fw_pair_kc(1, i));
          }
          ++S.accepted;
          S.rejected = 0;
          if (S.t - tb >= 0) S.status = FW_STATUS_FINISHED;
        } else {
          S.h_abs *= pw > 0.2 ? pw : 0.2;   // NaN error norms land here too (Python's max(0.2, nan) == 0.2)
          S.rejected = 1;
        }
      }
    }
    // ---- park finished aircraft: both halves of the raw final state, then ONE release by warp T ----
    const bool park = was_running && S.status != FW_STATUS_RUNNING;
    if (__ballot_sync(full, park)) {   // block-uniform
      if (park) {
        double* cd = a.cd + env;
        if (role == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) cd[(CY_RES + fw_kc_to_ode(fw_pair_kc(0, i))) * a.stride] = y[i];
#pragma unroll
          for (int j = 0; j < 3; ++j) cd[(CY_RES + 7 + j) * a.stride] = ypos[j];
          int32_t* ci = a.ci + env;
          ci[CI_FAIL * a.stride] = S.fail;
          ci[CI_ATTEMPTS * a.stride] = S.attempts;
          ci[CI_ACCEPTED * a.stride] = S.accepted;
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) cd[(CY_RES + fw_kc_to_ode(fw_pair_kc(1, i))) * a.stride] = y[i];
        }
      }
      __syncthreads();   // warp L's stores happen-before the release below (fence cumulativity)
      if (park && role == 0) {
        __threadfence();
        atomicAdd(FW_CHUNK_DONE(a.q, env), 1);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lane_attempts += __shfl_xor_sync(full, lane_attempts, o);
  if (role == 0 && lane == 0 && passes) {
    atomicAdd(a.ctr + CTR_WARP_MAX, passes);
    atomicAdd(a.ctr + CTR_WARP_STEPS, lane_attempts);
  }
  FW_TL_END(1);
}
#undef FW_PK
#undef FW_PG
