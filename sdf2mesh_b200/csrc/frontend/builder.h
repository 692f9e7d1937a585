// builder.h -- typed expression construction (type rules, implicit conversions, abstract-numeric
// constant evaluation) shared by both parsers.
#pragma once
#include <string>
#include <vector>

#include "ir.h"

namespace s2m_frontend {

enum class Lang { Wgsl, Glsl };

class Builder {
 public:
  explicit Builder(Lang l) : lang(l) {}
  Lang lang;
  int cur_line = 0;

  [[noreturn]] void error(const std::string& msg) const;                 // S2M_ERR_VALIDATION with line
  [[noreturn]] void unsupported(const std::string& msg) const;           // S2M_ERR_UNSUPPORTED

  ExprP lit_float(double v, Sk sk);
  ExprP lit_int(int64_t v, Sk sk);
  ExprP lit_bool(bool v);
  ExprP lit_from(const ConstVal& cv);  // scalar Lit or Construct of Lits
  ExprP var_ref(Var* v);
  ExprP unary(Op op, ExprP a);
  ExprP binary(Op op, ExprP a, ExprP b);
  ExprP ternary(ExprP c, ExprP t, ExprP f);
  ExprP swizzle(ExprP base, const std::string& comps);
  ExprP matrix_column(ExprP base, int col);  // m[i]
  ExprP member(ExprP base, const std::string& name);   // struct field or vector swizzle
  ExprP index(ExprP base, ExprP idx);                  // a[i], v[i], m[i] (constant or dynamic index)
  ExprP array_length(ExprP base);                      // GLSL a.length()
  ExprP bitcast(Sk target, ExprP e);                   // WGSL bitcast<T>(e), GLSL floatBitsToInt(e) ...: same width, other scalar kind
  Module* module = nullptr;                            // for interning array types
  ExprP construct(Type target, bool infer_sk, std::vector<ExprP> args);  // vecN(...) / scalar casts
  ExprP call_builtin(const std::string& name, std::vector<ExprP> args);  // returns null if not a builtin
  ExprP call_user(Function* fn, std::vector<ExprP> args);
  ExprP addr_of(ExprP a);
  ExprP deref(ExprP a);

  // Convert e to `target` where the language allows it implicitly (abstract -> concrete; GLSL
  // int -> float); errors otherwise.  `what` names the context for the message.
  ExprP coerce(ExprP e, Type target, const char* what);
  // Make e concrete: abstract float -> f32, abstract int -> i32 (WGSL default concretisation).
  ExprP concretize(ExprP e);
  bool const_eval(const Expr& e, ConstVal* out) const;
  bool is_const_expr(const Expr& e) const;  // const_eval-able, or an aggregate built from such values
  static bool is_lvalue(const Expr& e);

 private:
  ExprP mk(Expr::K k, Type ty);
  ExprP convert_sk(ExprP e, Sk sk);  // explicit scalar-kind conversion node (or folded literal)
};

bool is_builtin_name(const std::string& name, Lang lang);

}  // namespace s2m_frontend
