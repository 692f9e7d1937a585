// builder.cpp -- type rules, implicit conversions and constant evaluation (naga's `Typifier` +
// `ConstantEvaluator` roles).  WGSL abstract-float/abstract-int expressions are evaluated in
// f64 / i64 and only rounded when they meet a concrete type, as naga does -- this decides the
// f32 value of e.g. `const C = 1.0; ... C*0.45` in /root/reference/examples/martin_cube.sdf3d.
#include "builder.h"

#include <cmath>
#include <cstring>

#include "../s2m_math.h"

namespace s2m_frontend {

std::string Type::str() const {
  if (k == Void) return "void";
  if (k == Struct) return sdef->name;
  if (k == Array) return "array<" + adef->elem.str() + ", " + std::to_string(adef->len) + ">";
  const char* s = sk == Sk::Bool ? "bool" : sk == Sk::I32 ? "i32" : sk == Sk::U32 ? "u32" : sk == Sk::F32 ? "f32"
                  : sk == Sk::AInt ? "abstract-int" : "abstract-float";
  if (k == Scalar) return s;
  if (k == Matrix) return "mat" + std::to_string(n) + "x" + std::to_string(n) + "<f32>";
  return "vec" + std::to_string(n) + "<" + s + ">";
}

void Builder::error(const std::string& msg) const {
  throw FrontendError(4 /*S2M_ERR_VALIDATION*/, "validation error at line " + std::to_string(cur_line) + ": " + msg);
}
void Builder::unsupported(const std::string& msg) const {
  throw FrontendError(11 /*S2M_ERR_UNSUPPORTED*/, "unsupported at line " + std::to_string(cur_line) + ": " + msg);
}

ExprP Builder::mk(Expr::K k, Type ty) {
  ExprP e = std::make_shared<Expr>();
  e->k = k; e->ty = ty; e->line = cur_line;
  return e;
}
ExprP Builder::lit_float(double v, Sk sk) {
  ExprP e = mk(Expr::Lit, Type::scalar(sk));
  e->lit.ty = e->ty;
  e->lit.f[0] = (sk == Sk::F32) ? (double)(float)v : v;
  return e;
}
ExprP Builder::lit_int(int64_t v, Sk sk) {
  ExprP e = mk(Expr::Lit, Type::scalar(sk));
  e->lit.ty = e->ty;
  e->lit.i[0] = v;
  return e;
}
ExprP Builder::lit_bool(bool v) { return lit_int(v ? 1 : 0, Sk::Bool); }
ExprP Builder::lit_from(const ConstVal& cv) {
  auto one = [&](int c) -> ExprP {
    Sk sk = cv.ty.sk;
    if (sk == Sk::F32 || sk == Sk::AFloat) return lit_float(cv.f[c], sk);
    return lit_int(cv.i[c], sk);
  };
  if (cv.ty.is_scalar()) return one(0);
  ExprP e = mk(Expr::Construct, cv.ty);
  for (int c = 0; c < cv.ty.n; ++c) e->args.push_back(one(c));
  return e;
}
ExprP Builder::var_ref(Var* v) {
  ExprP e = mk(Expr::VarRef, v->ty);
  e->var = v;
  return e;
}

bool Builder::is_lvalue(const Expr& e) {
  if (e.k == Expr::VarRef) return !e.var->immutable && !e.var->is_ptr;
  if (e.k == Expr::Deref) return true;
  if (e.k == Expr::Swizzle || e.k == Expr::Member || e.k == Expr::Index) return is_lvalue(*e.args[0]);
  return false;
}

// ------------------------------------------------------------------ constant evaluation
namespace {
double as_f(const ConstVal& v, int c) {
  if (v.ty.is_float()) return v.f[v.ty.is_scalar() ? 0 : c];
  return (double)v.i[v.ty.is_scalar() ? 0 : c];
}
int64_t as_i(const ConstVal& v, int c) { return v.i[v.ty.is_scalar() ? 0 : c]; }
double round_to(Sk sk, double x) { return sk == Sk::F32 ? (double)(float)x : x; }
int64_t wrap_to(Sk sk, int64_t x) {
  if (sk == Sk::I32) return (int64_t)(int32_t)(uint32_t)x;
  if (sk == Sk::U32) return (int64_t)(uint32_t)x;
  return x;
}
ConstVal convert_cv(const ConstVal& v, Sk sk) {
  ConstVal o;
  o.ty = v.ty.with_sk(sk);
  for (int c = 0; c < v.ty.n; ++c) {
    if (sk == Sk::F32 || sk == Sk::AFloat) o.f[c] = round_to(sk, v.ty.is_float() ? v.f[c] : (double)v.i[c]);
    else if (sk == Sk::Bool) o.i[c] = v.ty.is_float() ? (v.f[c] != 0.0) : (v.i[c] != 0);
    else {
      if (v.ty.is_float()) {
        double d = v.f[c];
        float fl = (float)d;
        o.i[c] = sk == Sk::U32 ? (int64_t)s2m_f2uint(fl) : (sk == Sk::I32 ? (int64_t)s2m_f2int(fl) : (int64_t)d);
      } else o.i[c] = wrap_to(sk, v.i[c]);
    }
  }
  return o;
}
}  // namespace

bool Builder::is_const_expr(const Expr& e) const {
  ConstVal cv;
  if (const_eval(e, &cv)) return true;
  if (e.k == Expr::VarRef) return e.var->storage == Var::ModuleConst;
  if (e.k == Expr::Construct || e.k == Expr::Member || e.k == Expr::Index || e.k == Expr::Swizzle) {
    for (const ExprP& a : e.args) if (!is_const_expr(*a)) return false;
    return true;
  }
  return false;
}

bool Builder::const_eval(const Expr& e, ConstVal* out) const {
  if (e.ty.is_matrix() || e.ty.is_aggregate() || e.k == Expr::Member || e.k == Expr::Index) return false;
  for (const ExprP& a : e.args) if (a && (a->ty.is_matrix() || a->ty.is_aggregate())) return false;
  switch (e.k) {
    case Expr::Lit: *out = e.lit; return true;
    case Expr::VarRef:
      // a mutable module-scope variable has a known INITIAL value, which is not a constant
      if (e.var->has_const && (e.var->storage == Var::ModuleConst || e.var->immutable)) { *out = e.var->cval; return true; }
      return false;
    case Expr::Convert: {
      ConstVal a;
      if (!const_eval(*e.args[0], &a)) return false;
      *out = convert_cv(a, e.ty.sk);
      return true;
    }
    case Expr::Unary: {
      ConstVal a;
      if (!const_eval(*e.args[0], &a)) return false;
      out->ty = e.ty;
      for (int c = 0; c < e.ty.n; ++c) {
        if (e.op == Op::Neg) { if (e.ty.is_float()) out->f[c] = -a.f[c]; else out->i[c] = wrap_to(e.ty.sk, -a.i[c]); }
        else if (e.op == Op::Not) out->i[c] = !a.i[c];
        else if (e.op == Op::BitNot) out->i[c] = wrap_to(e.ty.sk, ~a.i[c]);
        else return false;
      }
      return true;
    }
    case Expr::Binary: {
      ConstVal a, b;
      if (!const_eval(*e.args[0], &a) || !const_eval(*e.args[1], &b)) return false;
      out->ty = e.ty;
      const Sk osk = (a.ty.is_float() || b.ty.is_float()) ? (a.ty.sk == Sk::F32 || b.ty.sk == Sk::F32 ? Sk::F32 : Sk::AFloat)
                                                         : (a.ty.sk == Sk::AInt ? b.ty.sk : a.ty.sk);
      const int n = std::max(a.ty.n, b.ty.n);
      for (int c = 0; c < n; ++c) {
        const bool cmp = e.op == Op::Lt || e.op == Op::Le || e.op == Op::Gt || e.op == Op::Ge || e.op == Op::Eq || e.op == Op::Ne;
        if (osk == Sk::F32 || osk == Sk::AFloat) {
          const double x = as_f(a, c), y = as_f(b, c);
          double r = 0;
          switch (e.op) {
            case Op::Add: r = round_to(osk, x + y); break;
            case Op::Sub: r = round_to(osk, x - y); break;
            case Op::Mul: r = osk == Sk::F32 ? (double)((float)x * (float)y) : x * y; break;
            case Op::Div: r = osk == Sk::F32 ? (double)((float)x / (float)y) : x / y; break;
            case Op::Rem: r = osk == Sk::F32 ? (double)s2m_fmod_trunc((float)x, (float)y) : x - y * std::trunc(x / y); break;
            case Op::Lt: out->i[c] = x < y; break;
            case Op::Le: out->i[c] = x <= y; break;
            case Op::Gt: out->i[c] = x > y; break;
            case Op::Ge: out->i[c] = x >= y; break;
            case Op::Eq: out->i[c] = x == y; break;
            case Op::Ne: out->i[c] = x != y; break;
            default: return false;
          }
          if (!cmp) out->f[c] = r;
        } else {
          const int64_t x = as_i(a, c), y = as_i(b, c);
          int64_t r = 0;
          switch (e.op) {
            case Op::Add: r = x + y; break;
            case Op::Sub: r = x - y; break;
            case Op::Mul: r = x * y; break;
            case Op::Div: if (y == 0) return false; r = x / y; break;
            case Op::Rem: if (y == 0) return false; r = x % y; break;
            case Op::BitAnd: case Op::And: r = x & y; break;
            case Op::BitOr: case Op::Or: r = x | y; break;
            case Op::BitXor: r = x ^ y; break;
            case Op::Shl: r = x << (y & 31); break;
            case Op::Shr: r = osk == Sk::U32 ? (int64_t)((uint32_t)x >> (y & 31)) : (x >> (y & 31)); break;
            case Op::Lt: r = x < y; break;
            case Op::Le: r = x <= y; break;
            case Op::Gt: r = x > y; break;
            case Op::Ge: r = x >= y; break;
            case Op::Eq: r = x == y; break;
            case Op::Ne: r = x != y; break;
            default: return false;
          }
          out->i[c] = cmp ? r : wrap_to(e.ty.sk, r);
        }
      }
      return true;
    }
    case Expr::Construct: {
      out->ty = e.ty;
      if (e.args.empty()) return true;  // zero value
      int c = 0;
      for (const ExprP& a : e.args) {
        ConstVal v;
        if (!const_eval(*a, &v)) return false;
        v = convert_cv(v, e.ty.sk);
        if (e.args.size() == 1 && v.ty.is_scalar()) {  // splat
          for (int k = 0; k < e.ty.n; ++k) { out->f[k] = v.f[0]; out->i[k] = v.i[0]; }
          return true;
        }
        for (int k = 0; k < v.ty.n && c < e.ty.n; ++k, ++c) { out->f[c] = v.f[k]; out->i[c] = v.i[k]; }
      }
      return true;
    }
    case Expr::Swizzle: {
      ConstVal a;
      if (!const_eval(*e.args[0], &a)) return false;
      out->ty = e.ty;
      for (int c = 0; c < e.nswz; ++c) { out->f[c] = a.f[e.swz[c]]; out->i[c] = a.i[e.swz[c]]; }
      return true;
    }
    case Expr::Ternary: {
      ConstVal c, t, f;
      if (!const_eval(*e.args[0], &c) || !const_eval(*e.args[1], &t) || !const_eval(*e.args[2], &f)) return false;
      *out = c.i[0] ? t : f;
      return true;
    }
    case Expr::Call: {
      std::vector<ConstVal> a(e.args.size());
      for (size_t i = 0; i < e.args.size(); ++i)
        if (!const_eval(*e.args[i], &a[i])) return false;
      if (!e.ty.is_float() && !(e.callee == "dot" || e.callee == "length")) return false;
      const bool f32 = e.ty.sk == Sk::F32;
      out->ty = e.ty;
      auto R = [&](double x) { return f32 ? (double)(float)x : x; };
      if (e.callee == "dot" || e.callee == "length" || e.callee == "distance") {
        double acc = 0;
        const int n = a[0].ty.n;
        for (int c = 0; c < n; ++c) {
          double x = as_f(a[0], c), y = e.callee == "dot" ? as_f(a[1], c) : (e.callee == "distance" ? as_f(a[1], c) : 0.0);
          double t = e.callee == "dot" ? R(x * y) : (e.callee == "distance" ? R(R(x - y) * R(x - y)) : R(x * x));
          acc = c == 0 ? t : R(acc + t);
        }
        out->f[0] = e.callee == "dot" ? acc : R(std::sqrt(acc));
        return true;
      }
      for (int c = 0; c < e.ty.n; ++c) {
        const double x = as_f(a[0], c);
        const double y = a.size() > 1 ? as_f(a[1], c) : 0.0, z = a.size() > 2 ? as_f(a[2], c) : 0.0;
        const float xf = (float)x, yf = (float)y, zf = (float)z;
        double r;
        const std::string& n = e.callee;
        if (n == "abs") r = std::fabs(x);
        else if (n == "min") r = f32 ? s2m_min(xf, yf) : std::fmin(x, y);
        else if (n == "max") r = f32 ? s2m_max(xf, yf) : std::fmax(x, y);
        else if (n == "clamp") r = f32 ? s2m_clamp(xf, yf, zf) : std::fmin(std::fmax(x, y), z);
        else if (n == "saturate") r = std::fmin(std::fmax(x, 0.0), 1.0);
        else if (n == "floor") r = std::floor(x);
        else if (n == "ceil") r = std::ceil(x);
        else if (n == "trunc") r = std::trunc(x);
        else if (n == "round") r = std::nearbyint(x);
        else if (n == "fract") r = R(x - std::floor(x));
        else if (n == "sign") r = x > 0 ? 1.0 : (x < 0 ? -1.0 : 0.0);
        else if (n == "sqrt") r = f32 ? sqrtf(xf) : std::sqrt(x);
        else if (n == "inversesqrt") r = f32 ? s2m_inversesqrt(xf) : 1.0 / std::sqrt(x);
        else if (n == "step") r = x <= y ? 1.0 : 0.0;
        else if (n == "mix") r = f32 ? s2m_mix(xf, yf, zf) : x * (1.0 - z) + y * z;
        else if (n == "sin") r = f32 ? s2m_sin(xf) : std::sin(x);
        else if (n == "cos") r = f32 ? s2m_cos(xf) : std::cos(x);
        else if (n == "tan") r = f32 ? s2m_tan(xf) : std::tan(x);
        else if (n == "asin") r = f32 ? s2m_asin(xf) : std::asin(x);
        else if (n == "acos") r = f32 ? s2m_acos(xf) : std::acos(x);
        else if (n == "atan") r = f32 ? s2m_atan(xf) : std::atan(x);
        else if (n == "atan2") r = f32 ? s2m_atan2(xf, yf) : std::atan2(x, y);
        else if (n == "exp") r = f32 ? s2m_exp(xf) : std::exp(x);
        else if (n == "exp2") r = f32 ? s2m_exp2(xf) : std::exp2(x);
        else if (n == "log") r = f32 ? s2m_log(xf) : std::log(x);
        else if (n == "log2") r = f32 ? s2m_log2(xf) : std::log2(x);
        else if (n == "pow") r = f32 ? s2m_pow(xf, yf) : std::pow(x, y);
        else if (n == "radians") r = f32 ? s2m_radians(xf) : x * 0.017453292519943295;
        else if (n == "degrees") r = f32 ? s2m_degrees(xf) : x * 57.29577951308232;
        else return false;
        out->f[c] = R(r);
      }
      return true;
    }
    default: return false;
  }
}

// ------------------------------------------------------------------ conversions
ExprP Builder::convert_sk(ExprP e, Sk sk) {
  if (e->ty.sk == sk) return e;
  ConstVal cv;
  if ((e->ty.is_abstract() || e->k == Expr::Lit) && const_eval(*e, &cv)) return lit_from(convert_cv(cv, sk));
  if (e->ty.is_abstract()) error("abstract-typed expression is not a constant expression");
  ExprP c = mk(Expr::Convert, e->ty.with_sk(sk));
  c->args.push_back(e);
  return c;
}

ExprP Builder::concretize(ExprP e) {
  if (!e->ty.is_abstract()) return e;
  return convert_sk(e, e->ty.sk == Sk::AFloat ? Sk::F32 : Sk::I32);
}

ExprP Builder::coerce(ExprP e, Type target, const char* what) {
  if (e->ty == target) return e;
  const bool same_shape = e->ty.k == target.k && e->ty.n == target.n;
  if (same_shape && !e->ty.is_void()) {
    if (e->ty.sk == Sk::AInt && (target.sk == Sk::I32 || target.sk == Sk::U32 || target.sk == Sk::F32 || target.sk == Sk::AFloat)) return convert_sk(e, target.sk);
    if (e->ty.sk == Sk::AFloat && (target.sk == Sk::F32)) return convert_sk(e, target.sk);
    if (lang == Lang::Glsl) {
      if ((e->ty.sk == Sk::I32 || e->ty.sk == Sk::U32) && target.sk == Sk::F32) return convert_sk(e, Sk::F32);
      if (e->ty.sk == Sk::I32 && target.sk == Sk::U32) return convert_sk(e, Sk::U32);
    }
  }
  error(std::string("type mismatch in ") + what + ": expected " + target.str() + ", found " + e->ty.str());
}

// ------------------------------------------------------------------ operators
namespace {
bool is_cmp(Op op) { return op == Op::Lt || op == Op::Le || op == Op::Gt || op == Op::Ge || op == Op::Eq || op == Op::Ne; }
}

ExprP Builder::unary(Op op, ExprP a) {
  if (op == Op::Neg) {
    if (!(a->ty.is_float() || a->ty.is_int())) error("unary - needs a numeric operand, found " + a->ty.str());
    if (a->k == Expr::Lit) {  // fold the sign into the literal (keeps -0.0)
      ConstVal cv = a->lit;
      if (a->ty.is_float()) cv.f[0] = -cv.f[0]; else cv.i[0] = wrap_to(a->ty.sk, -cv.i[0]);
      return lit_from(cv);
    }
  } else if (op == Op::Not) {
    if (!a->ty.is_bool()) error("! needs a bool operand, found " + a->ty.str());
  } else if (op == Op::BitNot) {
    if (!a->ty.is_int()) error("~ needs an integer operand, found " + a->ty.str());
  }
  ExprP e = mk(Expr::Unary, a->ty);
  e->op = op;
  e->args.push_back(a);
  return e;
}

ExprP Builder::binary(Op op, ExprP a, ExprP b) {
  if (a->ty.is_void() || b->ty.is_void()) error("void operand");
  if (a->ty.is_aggregate() || b->ty.is_aggregate()) error("operator not defined for " + (a->ty.is_aggregate() ? a->ty.str() : b->ty.str()));
  if (op == Op::And || op == Op::Or) {
    if (!a->ty.is_bool() || !b->ty.is_bool() || !a->ty.is_scalar() || !b->ty.is_scalar())
      error("logical operator needs bool operands, found " + a->ty.str() + " and " + b->ty.str());
    ExprP e = mk(Expr::Binary, Type::scalar(Sk::Bool));
    e->op = op; e->args = {a, b};
    return e;
  }
  if ((op == Op::BitAnd || op == Op::BitOr || op == Op::Eq || op == Op::Ne) && a->ty.is_bool() && b->ty.is_bool() && a->ty.n == b->ty.n) {
    ExprP e = mk(Expr::Binary, (op == Op::Eq || op == Op::Ne) ? Type::vec(Sk::Bool, a->ty.n) : a->ty);
    e->op = op; e->args = {a, b};
    return e;
  }
  if (a->ty.is_bool() || b->ty.is_bool()) error("arithmetic on bool operands");
  if (op == Op::Shl || op == Op::Shr) {
    // e1 << e2: the result has e1's type; e2 is an unsigned (GLSL: any integer) scalar or a vector of e1's width
    if (!a->ty.is_int() || !b->ty.is_int() || a->ty.is_matrix() || b->ty.is_matrix()) error("shift needs integer operands, found " + a->ty.str() + " and " + b->ty.str());
    if (b->ty.is_vector() && b->ty.n != a->ty.n) error("shift count width differs from the shifted value");
    ConstVal ca, cb;
    if (a->ty.is_abstract() && b->ty.is_abstract() && const_eval(*a, &ca) && const_eval(*b, &cb)) {
      ConstVal r = ca;
      for (int c = 0; c < ca.ty.n; ++c) {
        const int64_t sh = cb.i[cb.ty.n > 1 ? c : 0];
        if (sh < 0 || sh > 62) error("shift count out of range");
        r.i[c] = op == Op::Shl ? (ca.i[c] << sh) : (ca.i[c] >> sh);
      }
      return lit_from(r);
    }
    a = concretize(a);
    if (b->ty.is_abstract()) b = coerce(b, b->ty.with_sk(Sk::U32), "shift count");
    if (lang == Lang::Wgsl && b->ty.sk != Sk::U32) error("shift count must be u32, found " + b->ty.str());
    ExprP e = mk(Expr::Binary, a->ty);
    e->op = op; e->args = {a, b};
    return e;
  }
  if (a->ty.is_matrix() || b->ty.is_matrix()) {
    // linear algebra: mat*vec, vec*mat, mat*mat, mat*scalar, scalar*mat, mat/scalar, mat+-mat
    if (a->ty.is_abstract()) a = concretize(a);
    if (b->ty.is_abstract()) b = concretize(b);
    if (lang == Lang::Glsl) {
      if (a->ty.is_int() && !a->ty.is_matrix()) a = convert_sk(a, Sk::F32);
      if (b->ty.is_int() && !b->ty.is_matrix()) b = convert_sk(b, Sk::F32);
    }
    if (a->ty.sk != Sk::F32 || b->ty.sk != Sk::F32) error("matrix arithmetic needs f32 operands, found " + a->ty.str() + " and " + b->ty.str());
    Type rt;
    const bool am = a->ty.is_matrix(), bm = b->ty.is_matrix();
    if (op == Op::Mul) {
      if (am && bm) { if (a->ty.n != b->ty.n) error("matrix size mismatch"); rt = a->ty; }
      else if (am && b->ty.is_vector()) { if (a->ty.n != b->ty.n) error("matrix * vector size mismatch"); rt = b->ty; }
      else if (bm && a->ty.is_vector()) { if (a->ty.n != b->ty.n) error("vector * matrix size mismatch"); rt = a->ty; }
      else if (am && b->ty.is_scalar()) rt = a->ty;
      else if (bm && a->ty.is_scalar()) rt = b->ty;
      else error("bad operands for *: " + a->ty.str() + " and " + b->ty.str());
    } else if (op == Op::Add || op == Op::Sub) {
      if (!(am && bm && a->ty.n == b->ty.n)) error("matrix +/- needs two matrices of the same size");
      rt = a->ty;
    } else if (op == Op::Div) {
      if (!(am && b->ty.is_scalar())) error("matrix / needs a scalar divisor");
      rt = a->ty;
    } else error("operator not defined for matrices");
    ExprP e = mk(Expr::Binary, rt);
    e->op = op; e->args = {a, b};
    return e;
  }
  // shapes
  int n = 1;
  if (a->ty.is_vector() && b->ty.is_vector()) {
    if (a->ty.n != b->ty.n) error("vector size mismatch: " + a->ty.str() + " vs " + b->ty.str());
    n = a->ty.n;
  } else if (a->ty.is_vector()) n = a->ty.n;
  else if (b->ty.is_vector()) n = b->ty.n;
  // scalar kinds
  Sk sk;
  const bool aa = a->ty.is_abstract(), ba = b->ty.is_abstract();
  if (aa && ba) sk = (a->ty.sk == Sk::AFloat || b->ty.sk == Sk::AFloat) ? Sk::AFloat : Sk::AInt;
  else if (aa) { sk = b->ty.sk; a = coerce(a, a->ty.with_sk(sk), "binary operand"); }
  else if (ba) { sk = a->ty.sk; b = coerce(b, b->ty.with_sk(sk), "binary operand"); }
  else if (a->ty.sk == b->ty.sk) sk = a->ty.sk;
  else if (lang == Lang::Glsl) {
    if (a->ty.sk == Sk::F32 || b->ty.sk == Sk::F32) sk = Sk::F32; else sk = Sk::U32;
    a = coerce(a, a->ty.with_sk(sk), "binary operand");
    b = coerce(b, b->ty.with_sk(sk), "binary operand");
  } else error("operands of different types: " + a->ty.str() + " and " + b->ty.str());
  const bool bitop = op == Op::BitAnd || op == Op::BitOr || op == Op::BitXor || op == Op::Shl || op == Op::Shr;
  if (bitop && !(sk == Sk::I32 || sk == Sk::U32 || sk == Sk::AInt)) error("bit operation on non-integer operands");
  ExprP e = mk(Expr::Binary, is_cmp(op) ? Type::vec(Sk::Bool, n) : Type::vec(sk, n));
  e->op = op; e->args = {a, b};
  if (aa && ba && is_cmp(op)) {
    // a comparison of two abstract values has a concrete (bool) result, so nothing downstream would
    // concretize its operands: it is a constant expression, evaluated in abstract precision here
    ConstVal cv;
    if (const_eval(*e, &cv)) return lit_from(cv);
    e->args = {concretize(a), concretize(b)};
  }
  return e;
}

ExprP Builder::ternary(ExprP c, ExprP t, ExprP f) {
  if (!c->ty.is_bool() || !c->ty.is_scalar()) error("?: condition must be a scalar bool");
  if (t->ty.is_void() || f->ty.is_void()) error("void operand of ?:");
  if (t->ty != f->ty) {
    if (t->ty.is_abstract() || (lang == Lang::Glsl && f->ty.sk == Sk::F32 && t->ty.sk != Sk::F32)) t = coerce(t, f->ty, "?: operand");
    else f = coerce(f, t->ty, "?: operand");
  }
  ExprP e = mk(Expr::Ternary, t->ty);
  e->args = {c, t, f};
  return e;
}

ExprP Builder::member(ExprP base, const std::string& name) {
  if (!base->ty.is_struct()) return swizzle(base, name);
  const int f = base->ty.sdef->field(name);
  if (f < 0) error("struct " + base->ty.sdef->name + " has no member '" + name + "'");
  ExprP e = mk(Expr::Member, base->ty.sdef->field_types[(size_t)f]);
  e->args.push_back(base);
  e->nswz = 1;
  e->swz[0] = f;
  return e;
}

ExprP Builder::index(ExprP base, ExprP idx) {
  if (!(base->ty.is_vector() || base->ty.is_matrix() || base->ty.is_array())) error("cannot index a value of type " + base->ty.str());
  idx = concretize(idx);
  if (!idx->ty.is_scalar() || !(idx->ty.sk == Sk::I32 || idx->ty.sk == Sk::U32)) error("index must be an integer scalar, found " + idx->ty.str());
  const int len = base->ty.is_array() ? base->ty.adef->len : base->ty.n;
  ConstVal cv;
  if (const_eval(*idx, &cv)) {
    if (cv.i[0] < 0 || cv.i[0] >= len) error("index " + std::to_string(cv.i[0]) + " out of range for " + base->ty.str());
    if (base->ty.is_matrix()) return matrix_column(base, (int)cv.i[0]);
    if (base->ty.is_vector()) return swizzle(base, std::string(1, "xyzw"[cv.i[0]]));
  }
  const Type rt = base->ty.is_array() ? base->ty.adef->elem : base->ty.is_matrix() ? Type::vec(Sk::F32, base->ty.n) : Type::scalar(base->ty.sk);
  ExprP e = mk(Expr::Index, rt);
  e->args = {base, idx};
  return e;
}

ExprP Builder::bitcast(Sk target, ExprP e) {
  e = concretize(e);
  if (!(e->ty.is_scalar() || e->ty.is_vector()) || e->ty.is_bool()) error("bitcast of " + e->ty.str());
  if (target != Sk::F32 && target != Sk::I32 && target != Sk::U32) error("bitcast to a non-numeric type");
  ExprP c = mk(Expr::Call, e->ty.with_sk(target));
  c->callee = target == Sk::F32 ? "bits_f" : target == Sk::I32 ? "bits_i" : "bits_u";
  c->args.push_back(e);
  return c;
}

ExprP Builder::array_length(ExprP base) {
  if (!base->ty.is_array()) error(".length() of a non-array");
  return lit_int(base->ty.adef->len, Sk::I32);
}

ExprP Builder::swizzle(ExprP base, const std::string& comps) {
  if (base->ty.is_void()) error("swizzle of void");
  if (base->ty.is_aggregate()) error(base->ty.str() + " has no member '" + comps + "'");
  if (base->ty.is_matrix()) error("matrices have no named members; use m[i]");
  if (comps.empty() || comps.size() > 4) error("bad swizzle ." + comps);
  static const char* sets[] = {"xyzw", "rgba", "stpq"};
  int idx[4];
  const char* used = nullptr;
  for (size_t i = 0; i < comps.size(); ++i) {
    idx[i] = -1;
    for (const char* s : sets) {
      const char* p = strchr(s, comps[i]);
      if (p && *p) { if (used && used != s) error("mixed swizzle sets in ." + comps); used = s; idx[i] = (int)(p - s); }
    }
    if (idx[i] < 0) error("unknown member ." + comps);
    if (idx[i] >= base->ty.n) error("swizzle ." + comps + " out of range for " + base->ty.str());
  }
  if (base->ty.is_scalar()) {
    if (lang != Lang::Glsl) error("swizzle of a scalar");
    if (comps.size() == 1) return base;
    ExprP e = mk(Expr::Construct, Type::vec(base->ty.sk, (int)comps.size()));
    e->args.push_back(base);
    return e;
  }
  ExprP e = mk(Expr::Swizzle, Type::vec(base->ty.sk, (int)comps.size()));
  e->args.push_back(base);
  e->nswz = (int)comps.size();
  for (int i = 0; i < e->nswz; ++i) e->swz[i] = idx[i];
  return e;
}

ExprP Builder::matrix_column(ExprP base, int col) {
  if (!base->ty.is_matrix()) error("column access on a non-matrix");
  if (col < 0 || col >= base->ty.n) error("matrix column index out of range");
  ExprP e = mk(Expr::Swizzle, Type::vec(Sk::F32, base->ty.n));
  e->args.push_back(base);
  e->nswz = 1;
  e->swz[0] = col;
  return e;
}

ExprP Builder::construct(Type target, bool infer_sk, std::vector<ExprP> args) {
  for (const ExprP& a : args) if (a->ty.is_void()) error("void constructor argument");
  if (target.is_struct()) {
    const StructDef& d = *target.sdef;
    if (!args.empty()) {  // S() is the zero value
      if (args.size() != d.field_types.size()) error("constructor of " + d.name + " takes " + std::to_string(d.field_types.size()) + " arguments");
      for (size_t i = 0; i < args.size(); ++i) args[i] = coerce(args[i], d.field_types[i], "struct constructor");
    }
    ExprP e = mk(Expr::Construct, target);
    e->args = args;
    return e;
  }
  if (target.is_array()) {
    if (!args.empty()) {
      if ((int)args.size() != target.adef->len) error("constructor of " + target.str() + " takes " + std::to_string(target.adef->len) + " elements");
      for (ExprP& a : args) a = coerce(a, target.adef->elem, "array constructor");
    }
    ExprP e = mk(Expr::Construct, target);
    e->args = args;
    return e;
  }
  for (const ExprP& a : args) if (a->ty.is_aggregate()) error("cannot construct " + target.str() + " from " + a->ty.str());
  if (target.is_matrix()) {
    const int n = target.n;
    for (ExprP& a : args) {
      if (a->ty.is_matrix()) {
        if (args.size() != 1) error("a matrix constructor takes one matrix or scalars and vectors");
        if (a->ty.n == n) return a;
        // matN(matM): the upper-left block of the argument, the rest of the identity (GLSL 5.4.2)
        const int m = a->ty.n;
        static const char* const kFirst[] = {"", "x", "xy", "xyz", "xyzw"};
        std::vector<ExprP> cols;
        for (int c = 0; c < n; ++c) {
          std::vector<ExprP> parts;
          if (c < m) {
            ExprP col = index(a, lit_int(c, Sk::I32));
            parts.push_back(m >= n ? swizzle(col, kFirst[n]) : col);
          }
          for (int r = (c < m ? std::min(m, n) : 0); r < n; ++r) parts.push_back(lit_float(r == c ? 1.0 : 0.0, Sk::F32));
          cols.push_back(construct(Type::vec(Sk::F32, n), false, parts));
        }
        return construct(target, false, cols);
      }
      if (a->ty.is_abstract() || (lang == Lang::Glsl && a->ty.is_int())) a = convert_sk(a, Sk::F32);
      if (a->ty.sk != Sk::F32) error("matrix constructor needs f32 components, found " + a->ty.str());
    }
    int total = 0;
    for (const ExprP& a : args) total += a->ty.n;
    const bool diag = args.size() == 1 && args[0]->ty.is_scalar();
    if (!diag && !args.empty() && total != n * n) error("mat" + std::to_string(n) + " constructor has " + std::to_string(total) + " components");
    if (diag && lang != Lang::Glsl) error("WGSL has no diagonal matrix constructor");
    ExprP e = mk(Expr::Construct, target);
    e->args = args;
    return e;
  }
  if (target.is_scalar()) {
    if (args.empty()) return target.is_float() ? lit_float(0, target.sk) : lit_int(0, target.sk);
    if (args.size() != 1) error("scalar constructor takes one argument");
    ExprP a = args[0];
    if (a->ty.is_vector()) { if (lang != Lang::Glsl) error("scalar constructor from a vector"); a = swizzle(a, "x"); }
    return convert_sk(a, target.sk);
  }
  // vector
  Sk sk = target.sk;
  if (infer_sk) {
    if (args.empty()) error("cannot infer the component type of an empty vector constructor");
    bool any_concrete = false, any_afloat = false;
    for (const ExprP& a : args) {
      if (!a->ty.is_abstract()) { if (!any_concrete) sk = a->ty.sk; any_concrete = true; }
      else if (a->ty.sk == Sk::AFloat) any_afloat = true;
    }
    if (!any_concrete) sk = any_afloat ? Sk::AFloat : Sk::AInt;
  }
  int total = 0;
  for (const ExprP& a : args) total += a->ty.n;
  const int n = target.n;
  if (!args.empty()) {
    const bool splat = args.size() == 1 && args[0]->ty.is_scalar();
    const bool trunc = lang == Lang::Glsl && args.size() == 1 && args[0]->ty.is_vector() && total > n;
    if (!splat && !trunc && total != n)
      error("vec" + std::to_string(n) + " constructor has " + std::to_string(total) + " components");
    if (trunc) args[0] = swizzle(args[0], std::string("xyzw").substr(0, (size_t)n));
  }
  for (ExprP& a : args) {
    if (a->ty.sk == sk) continue;
    const bool conv_ctor = args.size() == 1 && a->ty.is_vector();  // vec3<f32>(ivec3): conversion constructor
    if (a->ty.is_abstract() || lang == Lang::Glsl || conv_ctor) {
      if (a->ty.is_abstract() && (sk == Sk::AFloat || sk == Sk::AInt)) continue;  // stays abstract, folded later
      if (a->ty.sk == Sk::AFloat && (sk == Sk::I32 || sk == Sk::U32)) error("abstract-float cannot convert to an integer vector component");
      a = convert_sk(a, sk);
    } else error("vector constructor component of type " + a->ty.str() + " in " + Type::vec(sk, n).str());
  }
  ExprP e = mk(Expr::Construct, Type::vec(sk, n));
  e->args = args;
  return e;
}

ExprP Builder::addr_of(ExprP a) {
  if (!is_lvalue(*a)) error("& needs a variable");
  ExprP e = mk(Expr::AddrOf, a->ty);
  e->args.push_back(a);
  return e;
}
ExprP Builder::deref(ExprP a) {
  if (!(a->k == Expr::VarRef && a->var->is_ptr)) error("* needs a pointer parameter");
  ExprP e = mk(Expr::Deref, a->ty);
  e->args.push_back(a);
  return e;
}

// ------------------------------------------------------------------ builtins
namespace {
struct BuiltinInfo { const char* name; const char* canon; int arity; char kind; int langs; };
// kind: 'm' component-wise float map; 'n' numeric map (float or int: abs min max clamp sign);
// 's' -> scalar result (length, distance, dot); 'x' cross; 'v' vector-only map (normalize);
// 'r' reflect; 'b' any/all; 'S' select;  langs: 1 wgsl, 2 glsl, 3 both
const BuiltinInfo kBuiltins[] = {
    {"abs", "abs", 1, 'n', 3},       {"sign", "sign", 1, 'm', 3},     {"floor", "floor", 1, 'm', 3},   {"ceil", "ceil", 1, 'm', 3},
    {"fract", "fract", 1, 'm', 3},   {"sqrt", "sqrt", 1, 'm', 3},     {"inverseSqrt", "inversesqrt", 1, 'm', 1},
    {"inversesqrt", "inversesqrt", 1, 'm', 2},                         {"sin", "sin", 1, 'm', 3},       {"cos", "cos", 1, 'm', 3},
    {"tan", "tan", 1, 'm', 3},       {"asin", "asin", 1, 'm', 3},     {"acos", "acos", 1, 'm', 3},     {"atan", "atan", 1, 'm', 3},
    {"sinh", "sinh", 1, 'm', 3},     {"cosh", "cosh", 1, 'm', 3},     {"tanh", "tanh", 1, 'm', 3},
    {"asinh", "asinh", 1, 'm', 3},   {"acosh", "acosh", 1, 'm', 3},   {"atanh", "atanh", 1, 'm', 3},     {"exp", "exp", 1, 'm', 3},
    {"exp2", "exp2", 1, 'm', 3},     {"log", "log", 1, 'm', 3},       {"log2", "log2", 1, 'm', 3},     {"radians", "radians", 1, 'm', 3},
    {"degrees", "degrees", 1, 'm', 3}, {"round", "round", 1, 'm', 3}, {"roundEven", "round", 1, 'm', 2}, {"trunc", "trunc", 1, 'm', 3},
    {"saturate", "saturate", 1, 'm', 1}, {"normalize", "normalize", 1, 'v', 3}, {"length", "length", 1, 's', 3},
    {"min", "min", 2, 'n', 3},       {"max", "max", 2, 'n', 3},       {"pow", "pow", 2, 'm', 3},       {"atan2", "atan2", 2, 'm', 1},
    {"atan", "atan2", 2, 'm', 2},    {"step", "step", 2, 'm', 3},     {"mod", "mod", 2, 'm', 2},       {"distance", "distance", 2, 's', 3},
    {"dot", "dot", 2, 's', 3},       {"cross", "cross", 2, 'x', 3},   {"reflect", "reflect", 2, 'r', 3},
    {"clamp", "clamp", 3, 'n', 3},   {"mix", "mix", 3, 'm', 3},       {"smoothstep", "smoothstep", 3, 'm', 3}, {"fma", "fma", 3, 'm', 3},
    {"select", "select", 3, 'S', 3}, {"any", "any", 1, 'b', 3},       {"all", "all", 1, 'b', 3},
    {"transpose", "transpose", 1, 'T', 3}, {"determinant", "determinant", 1, 'D', 3},
    {"inverse", "inverse", 1, 'T', 3},  // GLSL; accepted in WGSL text too (naga writes GLSL's inverse() as a helper function)
    // integer-only component-wise maps; the result has the argument's type (GLSL's bitCount / findMSB / findLSB return int: parse_glsl converts)
    {"countOneBits", "countOneBits", 1, 'i', 1}, {"bitCount", "countOneBits", 1, 'i', 2}, {"reverseBits", "reverseBits", 1, 'i', 1},
    {"bitfieldReverse", "reverseBits", 1, 'i', 2}, {"countLeadingZeros", "countLeadingZeros", 1, 'i', 1},
    {"countTrailingZeros", "countTrailingZeros", 1, 'i', 1}, {"firstLeadingBit", "firstLeadingBit", 1, 'i', 1},
    {"findMSB", "firstLeadingBit", 1, 'i', 2}, {"firstTrailingBit", "firstTrailingBit", 1, 'i', 1}, {"findLSB", "firstTrailingBit", 1, 'i', 2},
    {"extractBits", "extractBits", 3, 'E', 1}, {"bitfieldExtract", "extractBits", 3, 'E', 2},
    {"insertBits", "insertBits", 4, 'E', 1}, {"bitfieldInsert", "insertBits", 4, 'E', 2},
    {"refract", "refract", 3, 'R', 3}, {"faceForward", "faceforward", 3, 'F', 1}, {"faceforward", "faceforward", 3, 'F', 2},
};
}  // namespace

bool is_builtin_name(const std::string& name, Lang lang) {
  const int bit = lang == Lang::Wgsl ? 1 : 2;
  for (const auto& b : kBuiltins) if ((b.langs & bit) && name == b.name) return true;
  return false;
}

ExprP Builder::call_builtin(const std::string& name, std::vector<ExprP> args) {
  const int bit = lang == Lang::Wgsl ? 1 : 2;
  const BuiltinInfo* bi = nullptr;
  bool name_known = false;
  for (const auto& b : kBuiltins)
    if ((b.langs & bit) && name == b.name) { name_known = true; if (b.arity == (int)args.size()) { bi = &b; break; } }
  if (!bi) {
    if (name_known) error("wrong number of arguments for " + name + "()");
    return nullptr;
  }
  for (const ExprP& a : args) if (a->ty.is_void()) error("void argument to " + name + "()");
  auto call = [&](Type ty) { ExprP e = mk(Expr::Call, ty); e->callee = bi->canon; e->args = args; return e; };
  if (bi->kind == 'T' || bi->kind == 'D') {
    if (!args[0]->ty.is_matrix()) error(name + "() needs a matrix");
    return call(bi->kind == 'T' ? args[0]->ty : Type::scalar(Sk::F32));
  }
  for (const ExprP& a : args) if (a->ty.is_matrix()) error("matrix argument to " + name + "()");
  if (bi->kind == 'b') {
    if (!args[0]->ty.is_bool()) error(name + "() needs a bool vector");
    return call(Type::scalar(Sk::Bool));
  }
  if (bi->kind == 'S') {  // select(f, t, cond)
    if (!args[2]->ty.is_bool()) error("select() condition must be bool");
    if (args[0]->ty != args[1]->ty) {
      if (args[0]->ty.is_abstract()) args[0] = coerce(args[0], args[1]->ty, "select()"); else args[1] = coerce(args[1], args[0]->ty, "select()");
    }
    args[0] = concretize(args[0]); args[1] = concretize(args[1]);
    if (args[2]->ty.is_vector() && args[2]->ty.n != args[0]->ty.n) error("select() condition width mismatch");
    return call(args[0]->ty);
  }
  if (bi->kind == 'E') {  // extractBits(e, offset, count) / insertBits(e, newbits, offset, count): offset and count are scalars, u32 in the IR
    const size_t nv = args.size() - 2;
    for (size_t k = 0; k < nv; ++k) args[k] = concretize(args[k]);
    if (!args[0]->ty.is_int() || !(args[0]->ty.is_scalar() || args[0]->ty.is_vector())) error(name + "() needs an integer scalar or vector");
    if (nv == 2) { if (args[1]->ty.is_abstract()) args[1] = coerce(args[1], args[0]->ty, "insertBits()"); if (args[1]->ty != args[0]->ty) error(name + "(): value and inserted bits differ in type"); }
    for (size_t k = nv; k < args.size(); ++k) {
      if (!args[k]->ty.is_int() || !args[k]->ty.is_scalar()) error(name + "(): offset and count must be integer scalars");
      if (args[k]->ty.is_abstract()) args[k] = coerce(args[k], Type::scalar(Sk::U32), "bit-field argument");
      else if (args[k]->ty.sk != Sk::U32) { if (lang == Lang::Wgsl) error(name + "(): offset and count must be u32"); args[k] = convert_sk(args[k], Sk::U32); }
    }
    ExprP e = call(args[0]->ty);
    e->callee = std::string("i_") + bi->canon;
    return e;
  }
  if (bi->kind == 'i') {
    args[0] = concretize(args[0]);
    if (!args[0]->ty.is_int() || !(args[0]->ty.is_scalar() || args[0]->ty.is_vector())) error(name + "() needs an integer scalar or vector");
    ExprP e = call(args[0]->ty);
    e->callee = std::string("i_") + bi->canon;
    return e;
  }
  // width: all vector arguments must agree; scalars broadcast
  int n = 1;
  for (const ExprP& a : args)
    if (a->ty.is_vector()) { if (n != 1 && n != a->ty.n) error("vector size mismatch in " + name + "()"); n = a->ty.n; }
  bool all_abstract = true, all_int = true;
  for (const ExprP& a : args) {
    if (a->ty.is_bool()) error("bool argument to " + name + "()");
    if (!a->ty.is_abstract()) all_abstract = false;
    if (!a->ty.is_int()) all_int = false;
  }
  if ((bi->kind == 'n' || std::string(bi->canon) == "sign") && all_int && !all_abstract) {  // integer abs / min / max / clamp / sign
    Sk sk = Sk::I32;
    for (const ExprP& a : args) if (!a->ty.is_abstract()) sk = a->ty.sk;
    for (const ExprP& a : args) if (!a->ty.is_abstract() && a->ty.sk != sk) error("integer arguments of different types to " + name + "()");
    if (std::string(bi->canon) == "sign" && sk == Sk::U32) error("sign() of an unsigned value");
    for (ExprP& a : args) a = coerce(a, a->ty.with_sk(sk), "integer builtin argument");
    ExprP e = call(Type::vec(sk, n));
    e->callee = std::string("i_") + bi->canon;
    return e;
  }
  if ((bi->kind == 'n' || std::string(bi->canon) == "sign") && all_int && all_abstract) {
    // abs(-5), sign(-2147483648), min(1, 2), clamp(7, 0, 3): abstract-int in, abstract-int out (they used to turn into
    // abstract-float, which then did not combine with an i32)
    std::vector<ConstVal> cv(args.size());
    bool folded = true;
    for (size_t k = 0; k < args.size(); ++k) folded = folded && const_eval(*args[k], &cv[k]);
    if (folded) {
      ConstVal r;
      r.ty = Type::vec(Sk::AInt, n);
      const std::string fn = bi->canon;
      for (int c = 0; c < n; ++c) {
        auto at = [&](size_t k) { return cv[k].i[cv[k].ty.n > 1 ? c : 0]; };
        const int64_t x = at(0);
        if (fn == "abs") r.i[c] = x < 0 ? -x : x;
        else if (fn == "sign") r.i[c] = x > 0 ? 1 : (x < 0 ? -1 : 0);
        else if (fn == "min") r.i[c] = std::min(x, at(1));
        else if (fn == "max") r.i[c] = std::max(x, at(1));
        else r.i[c] = std::min(std::max(x, at(1)), at(2));   // clamp
      }
      return lit_from(r);
    }
  }
  Sk sk = Sk::F32;
  if (all_abstract) {
    // keep abstract if the call folds; otherwise fall through to f32
    for (ExprP& a : args) if (a->ty.sk == Sk::AInt) a = convert_sk(a, Sk::AFloat);
    ExprP probe = call(Type::vec(Sk::AFloat, bi->kind == 's' ? 1 : n));
    ConstVal cv;
    if (const_eval(*probe, &cv)) return lit_from(cv);
  }
  for (ExprP& a : args) {
    if (a->ty.sk == Sk::F32) continue;
    if (a->ty.is_abstract() || (lang == Lang::Glsl && a->ty.is_int())) a = convert_sk(a, Sk::F32);
    else error("argument of type " + a->ty.str() + " to " + name + "(); expected floating point");
  }
  switch (bi->kind) {
    case 's':
      if (bi->arity == 2 && args[0]->ty.n != args[1]->ty.n) error(name + "() operands differ in size");
      return call(Type::scalar(sk));
    case 'x':
      if (args[0]->ty.n != 3 || args[1]->ty.n != 3) error("cross() needs vec3 operands");
      return call(Type::vec(sk, 3));
    case 'v':
      if (n == 1) error("normalize() needs a vector");
      return call(Type::vec(sk, n));
    case 'r':
      if (n == 1 || args[0]->ty.n != args[1]->ty.n) error("reflect() needs two vectors of the same size");
      return call(Type::vec(sk, n));
    case 'R':
      if (n == 1 || args[0]->ty.n != n || args[1]->ty.n != n || !args[2]->ty.is_scalar()) error("refract() needs two vectors of the same size and a scalar");
      return call(Type::vec(sk, n));
    case 'F':
      if (n == 1 || args[0]->ty.n != n || args[1]->ty.n != n || args[2]->ty.n != n) error(name + "() needs three vectors of the same size");
      return call(Type::vec(sk, n));
    default:
      return call(Type::vec(sk, n));
  }
}

ExprP Builder::call_user(Function* fn, std::vector<ExprP> args) {
  if (args.size() != fn->params.size())
    error("function " + fn->name + " takes " + std::to_string(fn->params.size()) + " arguments, " + std::to_string(args.size()) + " given");
  for (size_t i = 0; i < args.size(); ++i) {
    Var* p = fn->params[i];
    if (p->is_ptr) {
      if (args[i]->k != Expr::AddrOf && !(args[i]->k == Expr::VarRef && args[i]->var->is_ptr)) error("argument " + std::to_string(i + 1) + " of " + fn->name + " must be a pointer (&var)");
      if (args[i]->ty != p->ty) error("pointer argument type mismatch for " + fn->name);
    } else if (p->by_ref) {
      if (!is_lvalue(*args[i])) error("argument " + std::to_string(i + 1) + " of " + fn->name + " must be a variable (out/inout)");
      if (args[i]->ty != p->ty) error("out/inout argument type mismatch for " + fn->name);
    } else {
      args[i] = coerce(args[i], p->ty, ("argument of " + fn->name).c_str());
    }
  }
  ExprP e = mk(Expr::UserCall, fn->ret);
  e->fn = fn;
  e->args = args;
  return e;
}

}  // namespace s2m_frontend
