/* Shared dispatcher (host + device) used by the math tests: fn id -> s2m_math.h function. */
#pragma once
#include "s2m_math.h"
enum { FN_SIN, FN_COS, FN_TAN, FN_ASIN, FN_ACOS, FN_ATAN, FN_EXP, FN_EXP2, FN_LOG, FN_LOG2,
       FN_SINH, FN_COSH, FN_TANH, FN_SQRT, FN_ABS, FN_FLOOR, FN_FRACT, FN_SIGN, FN_ROUND, FN_ASINH, FN_ACOSH, FN_ATANH, FN_COUNT1,
       FN_ATAN2 = 100, FN_POW, FN_MIN, FN_MAX, FN_DIV, FN_FMOD, FN_MOD, FN_STEP };
S2M_HD float s2m_dispatch1(int fn, float x) {
  switch (fn) {
    case FN_SIN: return s2m_sin(x);   case FN_COS: return s2m_cos(x);   case FN_TAN: return s2m_tan(x);
    case FN_ASIN: return s2m_asin(x); case FN_ACOS: return s2m_acos(x); case FN_ATAN: return s2m_atan(x);
    case FN_EXP: return s2m_exp(x);   case FN_EXP2: return s2m_exp2(x); case FN_LOG: return s2m_log(x);
    case FN_LOG2: return s2m_log2(x); case FN_SINH: return s2m_sinh(x); case FN_COSH: return s2m_cosh(x);
    case FN_TANH: return s2m_tanh(x); case FN_SQRT: return s2m_sqrt(x); case FN_ABS: return s2m_abs(x);
    case FN_FLOOR: return s2m_floor(x); case FN_FRACT: return s2m_fract(x); case FN_SIGN: return s2m_sign(x);
    case FN_ROUND: return s2m_round(x);
    case FN_ASINH: return s2m_asinh(x); case FN_ACOSH: return s2m_acosh(x); case FN_ATANH: return s2m_atanh(x);
  }
  return 0.0f;
}
S2M_HD float s2m_dispatch2(int fn, float x, float y) {
  switch (fn) {
    case FN_ATAN2: return s2m_atan2(x, y); case FN_POW: return s2m_pow(x, y);
    case FN_MIN: return s2m_min(x, y);     case FN_MAX: return s2m_max(x, y);
    case FN_DIV: return x / y;             case FN_FMOD: return s2m_fmod_trunc(x, y);
    case FN_MOD: return s2m_mod_floor(x, y); case FN_STEP: return s2m_step(x, y);
  }
  return 0.0f;
}
