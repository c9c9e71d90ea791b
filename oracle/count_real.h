/* count_real.h -- TEST / MEASUREMENT INFRASTRUCTURE (never part of the product).
 *
 * With -DORACLE_COUNT, oracle/mjstep.c is compiled as C++ (g++ -x c++ -fpermissive) with `real` replaced by this
 * class: a double whose arithmetic operators tally floating-point operations per stage of mjx.step.  Counting rule
 * (SURVEY.md section 8d): add, subtract, multiply, divide, sqrt, sin, cos, pow = 1 each (so a multiply-add = 2);
 * comparisons, min / max selections, abs, negation, copies and integer work = 0.  tools/count_flops.py drives it and
 * writes profiles/r2_flop_count.json, which bench.py reads as F_A, the algorithmic work per env-step. */
#pragma once
#include <cmath>
#include <type_traits>

#define ORACLE_NSTAGE 8
extern thread_local long long g_cnt[ORACLE_NSTAGE];
extern thread_local int g_stage;
#define OSTAGE(k) (g_stage = (k))
#define OTICK() (++g_cnt[g_stage])

struct real {
  double v;
  real() = default;
  real(double x) : v(x) {}
  operator double() const { return v; }
};
template <class T> using IfNum = typename std::enable_if<std::is_arithmetic<T>::value, int>::type;

#define ORACLE_BINOP(op)                                                                                        \
  inline real operator op(real a, real b) { OTICK(); return real(a.v op b.v); }                                 \
  template <class T, IfNum<T> = 0> inline real operator op(real a, T b) { OTICK(); return real(a.v op (double)b); } \
  template <class T, IfNum<T> = 0> inline real operator op(T a, real b) { OTICK(); return real((double)a op b.v); }
ORACLE_BINOP(+)
ORACLE_BINOP(-)
ORACLE_BINOP(*)
ORACLE_BINOP(/)
#undef ORACLE_BINOP
#define ORACLE_ASSIGNOP(op, bop)                                                                          \
  inline real& operator op(real& a, real b) { OTICK(); a.v = a.v bop b.v; return a; }                     \
  template <class T, IfNum<T> = 0> inline real& operator op(real& a, T b) { OTICK(); a.v = a.v bop (double)b; return a; }
ORACLE_ASSIGNOP(+=, +)
ORACLE_ASSIGNOP(-=, -)
ORACLE_ASSIGNOP(*=, *)
ORACLE_ASSIGNOP(/=, /)
#undef ORACLE_ASSIGNOP
inline real operator-(real a) { return real(-a.v); }
#define ORACLE_CMP(op)                                                                                    \
  inline bool operator op(real a, real b) { return a.v op b.v; }                                          \
  template <class T, IfNum<T> = 0> inline bool operator op(real a, T b) { return a.v op (double)b; }      \
  template <class T, IfNum<T> = 0> inline bool operator op(T a, real b) { return (double)a op b.v; }
ORACLE_CMP(<)
ORACLE_CMP(>)
ORACLE_CMP(<=)
ORACLE_CMP(>=)
ORACLE_CMP(==)
ORACLE_CMP(!=)
#undef ORACLE_CMP
inline real cnt_sqrt(real a) { OTICK(); return real(std::sqrt(a.v)); }
inline real cnt_sin(real a) { OTICK(); return real(std::sin(a.v)); }
inline real cnt_cos(real a) { OTICK(); return real(std::cos(a.v)); }
inline real cnt_pow(real a, real b) { OTICK(); return real(std::pow(a.v, b.v)); }
inline real cnt_fabs(real a) { return real(std::fabs(a.v)); }
#define RSQRT cnt_sqrt
#define RFABS cnt_fabs
#define RSIN cnt_sin
#define RCOS cnt_cos
#define RPOW cnt_pow
