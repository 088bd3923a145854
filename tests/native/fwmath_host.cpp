// Host twin of csrc/fwmath.cuh for the CPU accuracy test (tests/test_fwmath_cpu.py): same source, g++-compiled.
#include "../../fixed-wing-gym_b200/csrc/fwmath.cuh"

extern "C" {
void t_atan2(const double* y, const double* x, double* o, long n) { for (long i = 0; i < n; ++i) o[i] = fwm_atan2(y[i], x[i]); }
void t_exp(const double* x, double* o, long n) { for (long i = 0; i < n; ++i) o[i] = fwm_exp(x[i]); }
void t_log(const double* x, double* o, long n) { for (long i = 0; i < n; ++i) o[i] = fwm_log(x[i]); }
void t_pow(const double* x, double p, double* o, long n) { for (long i = 0; i < n; ++i) o[i] = fwm_pow(x[i], p); }
void t_sqrt(const double* x, double* o, long n) { for (long i = 0; i < n; ++i) o[i] = fwm_sqrt(x[i]); }
void t_rsqrt(const double* x, double* o, long n) { for (long i = 0; i < n; ++i) { double s; fwm_sqrt_rsqrt(x[i], &s, &o[i]); } }
void t_rcp(const double* x, double* o, long n) { for (long i = 0; i < n; ++i) o[i] = fwm_rcp(x[i]); }
void t_div(const double* a, const double* b, double* o, long n) { for (long i = 0; i < n; ++i) o[i] = fwm_div(a[i], b[i]); }
void t_sinpi(const double* x, double* o, long n) { for (long i = 0; i < n; ++i) { double c; fwm_sincospi(x[i], &o[i], &c); } }
void t_cospi(const double* x, double* o, long n) { for (long i = 0; i < n; ++i) { double s; fwm_sincospi(x[i], &s, &o[i]); } }
}
