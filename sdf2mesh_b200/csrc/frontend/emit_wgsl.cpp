// emit_wgsl.cpp -- IR -> WGSL text.  Plays the role of naga::back::wgsl::Writer in
// convert_glsl_to_wgsl (/root/reference/src/shadertoy.rs:169-194): the GLSL front-end's module is
// written out as WGSL, which is what the reference then munges (remove `fn main_1(`, `fn main(`,
// `@fragment`), extends with sdf3d_normal + the sdf3d wrapper (shader.rs:84-98) and compiles.  It is
// also what --debug-wgsl shows for --glsl inputs.  The text is naga-shaped (typed literals, params
// copied into locals, out-params as ptr<function,T>, `fn main_1()` + `@fragment fn main()`), not
// byte-identical to naga's output.
#include <cmath>
#include <cstdio>
#include <map>
#include <set>
#include <sstream>

#include "parse.h"

namespace s2m_frontend {

namespace {

class WgslWriter {
 public:
  explicit WgslWriter(const Module& m) : m_(m) {}

  std::string run() {
    for (const auto& v : m_.vars) used_.insert(v->name);
    for (const auto& f : m_.functions) used_.insert(f->name);
    for (const auto& sd : m_.structs) {
      out_ << "struct " << sd->name << " {\n";
      for (size_t i = 0; i < sd->field_names.size(); ++i) out_ << "    " << sd->field_names[i] << ": " << type_name(sd->field_types[i]) << ",\n";
      out_ << "}\n\n";
    }
    for (const Var* g : m_.globals) {
      if (g->storage == Var::ModuleConst && g->has_const) {
        out_ << "const " << g->name << ": " << type_name(g->ty) << " = " << const_lit(g->cval) << ";\n";
      } else if (g->storage == Var::ModuleConst && (g->ty.is_aggregate() || g->ty.is_matrix()) && m_.global_init.count(g)) {
        out_ << "const " << g->name << ": " << type_name(g->ty) << " = " << expr(*m_.global_init.at(g)) << ";\n";
      } else {
        out_ << "var<private> " << g->name << ": " << type_name(g->ty);
        auto it = m_.global_init.find(g);
        if (it != m_.global_init.end()) out_ << " = " << expr(*it->second);
        out_ << ";\n";
      }
    }
    if (!m_.globals.empty()) out_ << "\n";
    for (const auto& f : m_.functions) {
      if (f->is_entry || !f->body) continue;
      function(*f);
      out_ << "\n";
    }
    for (const auto& d : m_.dropped) {
      std::string why = d.second;
      for (char& c : why) if (c == '\n' || c == '\r') c = ' ';
      out_ << "// left out: fn " << d.first << " -- " << why << "\n";
    }
    if (!m_.dropped.empty()) out_ << "\n";
    // naga's shape for an empty GLSL `void main() {}`
    out_ << "fn main_1() {\n    return;\n}\n\n@fragment \nfn main() {\n    main_1();\n    return;\n}\n";
    return out_.str();
  }

 private:
  const Module& m_;
  std::ostringstream out_;
  std::set<std::string> used_;
  std::map<const Var*, std::string> rename_;  // params copied into a mutable local
  int tmp_ = 0;

  static std::string type_name(const Type& t) {
    const char* s = t.sk == Sk::F32 ? "f32" : t.sk == Sk::I32 ? "i32" : t.sk == Sk::U32 ? "u32" : "bool";
    if (t.is_void()) return "void";
    if (t.is_struct()) return t.sdef->name;
    if (t.is_array()) return "array<" + type_name(t.adef->elem) + ", " + std::to_string(t.adef->len) + ">";
    if (t.is_matrix()) return "mat" + std::to_string(t.n) + "x" + std::to_string(t.n) + "<f32>";
    if (t.is_scalar()) return s;
    return "vec" + std::to_string(t.n) + "<" + s + ">";
  }
  std::string fresh(const std::string& base) {
    for (int i = 1;; ++i) {
      std::string n = base + "_" + std::to_string(i);
      if (!used_.count(n)) { used_.insert(n); return n; }
    }
  }
  static std::string float_lit(double v) {
    const float f = (float)v;
    if (f != f) return "(0f / 0f)";
    if (std::isinf(f)) return f > 0 ? "(1f / 0f)" : "(-1f / 0f)";
    char buf[64];
    snprintf(buf, sizeof buf, "%.9g", (double)f);
    return std::string(buf) + "f";
  }
  static std::string scalar_lit(const ConstVal& cv, int c) {
    switch (cv.ty.sk) {
      case Sk::F32: return float_lit(cv.f[c]);
      case Sk::I32: return cv.i[c] == INT32_MIN ? "(-2147483647i - 1i)" : std::to_string((int32_t)cv.i[c]) + "i";
      case Sk::U32: return std::to_string((uint32_t)cv.i[c]) + "u";
      case Sk::Bool: return cv.i[c] ? "true" : "false";
      default: throw FrontendError(4, "internal: abstract literal in the WGSL writer");
    }
  }
  static std::string const_lit(const ConstVal& cv) {
    if (cv.ty.is_aggregate() || cv.ty.is_matrix()) return type_name(cv.ty) + "()";  // zero value
    if (cv.ty.is_scalar()) return scalar_lit(cv, 0);
    std::string s = type_name(cv.ty) + "(";
    for (int c = 0; c < cv.ty.n; ++c) s += (c ? ", " : "") + scalar_lit(cv, c);
    return s + ")";
  }
  std::string var_name(const Var* v) {
    auto it = rename_.find(v);
    if (it != rename_.end()) return it->second;
    if (v->by_ref && v->storage == Var::Param) return "(*" + v->name + ")";
    return v->name;
  }
  static std::string wgsl_builtin(const std::string& canon) {
    if (canon == "inversesqrt") return "inverseSqrt";
    if (canon == "faceforward") return "faceForward";
    return canon;
  }

  std::string splat_to(const Expr& a, int n) {
    std::string s = expr(a);
    if (n > 1 && a.ty.is_scalar()) return type_name(Type::vec(a.ty.sk, n)) + "(" + s + ")";
    return s;
  }

  std::string expr(const Expr& e) {
    switch (e.k) {
      case Expr::Lit: return scalar_lit(e.lit, 0);
      case Expr::VarRef: return var_name(e.var);
      case Expr::Unary: {
        const char* o = e.op == Op::Neg ? "-" : e.op == Op::Not ? "!" : "~";
        return std::string(o) + "(" + expr(*e.args[0]) + ")";
      }
      case Expr::Binary: {
        static const std::map<Op, const char*> ops = {
            {Op::Add, "+"}, {Op::Sub, "-"}, {Op::Mul, "*"}, {Op::Div, "/"}, {Op::Rem, "%"}, {Op::And, "&&"}, {Op::Or, "||"},
            {Op::BitAnd, "&"}, {Op::BitOr, "|"}, {Op::BitXor, "^"}, {Op::Shl, "<<"}, {Op::Shr, ">>"}, {Op::Lt, "<"},
            {Op::Le, "<="}, {Op::Gt, ">"}, {Op::Ge, ">="}, {Op::Eq, "=="}, {Op::Ne, "!="}};
        std::string rhs = expr(*e.args[1]);
        if ((e.op == Op::Shl || e.op == Op::Shr) && e.args[1]->ty.sk != Sk::U32)  // WGSL shift counts are u32
          rhs = type_name(e.args[1]->ty.with_sk(Sk::U32)) + "(" + rhs + ")";
        return "(" + expr(*e.args[0]) + " " + ops.at(e.op) + " " + rhs + ")";
      }
      case Expr::Ternary:  // arms with calls or effects were lowered to if / else by the GLSL parser; what is left is pure
        return "select(" + expr(*e.args[2]) + ", " + expr(*e.args[1]) + ", " + expr(*e.args[0]) + ")";
      case Expr::Call: {
        std::string name = e.callee;
        if (name.compare(0, 5, "bits_") == 0) return "bitcast<" + type_name(e.ty) + ">(" + expr(*e.args[0]) + ")";
        if (name.compare(0, 2, "i_") == 0) name = name.substr(2);
        int n = 1;
        for (const ExprP& a : e.args) n = std::max(n, a->ty.n);
        if (name == "mod") {  // GLSL mod(x, y) = x - y * floor(x / y)
          const std::string x = splat_to(*e.args[0], n), y = splat_to(*e.args[1], n);
          return "(" + x + " - " + y + " * floor(" + x + " / " + y + "))";
        }
        std::string s = wgsl_builtin(name) + "(";
        for (size_t i = 0; i < e.args.size(); ++i) {
          const bool keep_scalar = (name == "mix" && i == 2) || (name == "refract" && i == 2) || name == "dot" || name == "length" || name == "distance" || name == "select" || name == "any" || name == "all" ||
                                   (name == "extractBits" && i >= 1) || (name == "insertBits" && i >= 2);   // offset, count: u32 scalars
          s += (i ? ", " : "") + (keep_scalar ? expr(*e.args[i]) : splat_to(*e.args[i], n));
        }
        return s + ")";
      }
      case Expr::UserCall: {
        std::string s = e.fn->name + "(";
        for (size_t i = 0; i < e.args.size(); ++i) {
          s += i ? ", " : "";
          if (e.fn->params[i]->by_ref) {
            const Expr& a = *e.args[i];
            if (a.k == Expr::VarRef && a.var->by_ref && a.var->storage == Var::Param && !rename_.count(a.var)) s += a.var->name;  // forward the pointer
            else s += "&" + expr(a);
          } else s += expr(*e.args[i]);
        }
        return s + ")";
      }
      case Expr::Construct: {
        if (e.args.empty()) return type_name(e.ty) + "()";
        if (e.ty.is_matrix() && e.args.size() == 1 && e.args[0]->ty.is_scalar()) {  // GLSL diagonal constructor
          const std::string d = expr(*e.args[0]);
          std::string m = type_name(e.ty) + "(";
          for (int c = 0; c < e.ty.n; ++c)
            for (int r = 0; r < e.ty.n; ++r) m += std::string(c || r ? ", " : "") + (c == r ? d : "0f");
          return m + ")";
        }
        std::string s = type_name(e.ty) + "(";
        for (size_t i = 0; i < e.args.size(); ++i) s += (i ? ", " : "") + expr(*e.args[i]);
        return s + ")";
      }
      case Expr::Swizzle: {
        if (e.args[0]->ty.is_matrix()) return expr(*e.args[0]) + "[" + std::to_string(e.swz[0]) + "]";
        std::string s = expr(*e.args[0]) + ".";
        for (int i = 0; i < e.nswz; ++i) s += "xyzw"[e.swz[i]];
        return s;
      }
      case Expr::Member: return expr(*e.args[0]) + "." + e.args[0]->ty.sdef->field_names[(size_t)e.swz[0]];
      case Expr::Index: return expr(*e.args[0]) + "[" + expr(*e.args[1]) + "]";
      case Expr::Convert: return type_name(e.ty) + "(" + expr(*e.args[0]) + ")";
      case Expr::AddrOf: return "&" + expr(*e.args[0]);
      case Expr::Deref: return "(*" + expr(*e.args[0]) + ")";
    }
    return "";
  }

  void indent(int d) { for (int i = 0; i < d; ++i) out_ << "    "; }

  void function(const Function& f) {
    rename_.clear();
    out_ << "fn " << f.name << "(";
    for (size_t i = 0; i < f.params.size(); ++i) {
      const Var* p = f.params[i];
      out_ << (i ? ", " : "") << p->name << ": ";
      if (p->by_ref) out_ << "ptr<function, " << type_name(p->ty) << ">";
      else out_ << type_name(p->ty);
    }
    out_ << ")";
    if (!f.ret.is_void()) out_ << " -> " << type_name(f.ret);
    out_ << " {\n";
    // WGSL parameters are immutable: copy the ones the GLSL body assigns to (naga does this for all)
    for (const Var* p : f.params) {
      if (p->by_ref || !p->written) continue;
      const std::string local = fresh(p->name);
      indent(1);
      out_ << "var " << local << ": " << type_name(p->ty) << " = " << p->name << ";\n";
      rename_[p] = local;
    }
    for (const StmtP& s : f.body->body) stmt(*s, 1);
    out_ << "}\n";
  }

  void block(const Stmt& s, int d) {
    out_ << "{\n";
    for (const StmtP& c : s.body) stmt(*c, d + 1);
    indent(d);
    out_ << "}";
  }

  // assignment without indentation / terminator; multi-component swizzle stores are expanded
  void assign(const Stmt& s, int d, bool in_header) {
    const Expr& lhs = *s.a;
    if (lhs.k == Expr::Swizzle && lhs.nswz > 1) {
      if (in_header) throw FrontendError(11, "unsupported: swizzle assignment in a for header");
      const Expr& base = *lhs.args[0];
      bool identity = lhs.nswz == base.ty.n;
      for (int i = 0; i < lhs.nswz; ++i) identity = identity && lhs.swz[i] == i;
      if (identity) { out_ << expr(base) << " = " << expr(*s.b) << ";"; return; }
      const std::string t = fresh("_e");
      out_ << "{\n";
      indent(d + 1);
      out_ << "let " << t << " = " << expr(*s.b) << ";\n";
      for (int i = 0; i < lhs.nswz; ++i) {
        indent(d + 1);
        out_ << expr(base) << "." << "xyzw"[lhs.swz[i]] << " = " << t << "." << "xyzw"[i] << ";\n";
      }
      indent(d);
      out_ << "}";
      return;
    }
    out_ << expr(lhs) << " = " << expr(*s.b);
    if (!in_header) out_ << ";";
  }

  void header_stmt(const Stmt& s, int d) {
    if (s.k == Stmt::Assign) assign(s, d, true);
    else if (s.k == Stmt::CallStmt) out_ << expr(*s.a);
    else if (s.k == Stmt::VarDecl) {
      out_ << "var " << s.var->name << ": " << type_name(s.var->ty);
      if (s.a) out_ << " = " << expr(*s.a);
    } else throw FrontendError(11, "unsupported statement in a for header");
  }

  void stmt(const Stmt& s, int d) {
    indent(d);
    switch (s.k) {
      case Stmt::Block: block(s, d); out_ << "\n"; break;
      case Stmt::VarDecl:
        out_ << (s.var->immutable && s.a ? "let " : "var ") << s.var->name << ": " << type_name(s.var->ty);
        if (s.a) out_ << " = " << expr(*s.a);
        out_ << ";\n";
        break;
      case Stmt::Assign: assign(s, d, false); out_ << "\n"; break;
      case Stmt::CallStmt: {
        // f(p.xz, a) with an out/inout parameter: WGSL cannot take the address of a swizzle, so the
        // components travel through a temporary (what naga does for GLSL out-parameters)
        const Expr& c = *s.a;
        std::vector<std::pair<const Expr*, std::string>> temps;
        if (c.k == Expr::UserCall) {
          for (size_t i = 0; i < c.args.size(); ++i) {
            const Expr* a = c.args[i].get();
            if (c.fn->params[i]->by_ref && a->k == Expr::Swizzle && a->nswz > 1 && !a->args[0]->ty.is_matrix()) {
              const std::string t = fresh("swz");
              temps.emplace_back(a, t);
              out_ << "var " << t << ": " << type_name(a->ty) << " = " << expr(*a) << ";\n";
              indent(d);
            }
          }
        }
        if (temps.empty()) { out_ << expr(c) << ";\n"; break; }
        out_ << c.fn->name << "(";
        for (size_t i = 0; i < c.args.size(); ++i) {
          out_ << (i ? ", " : "");
          const std::string* t = nullptr;
          for (const auto& tp : temps) if (tp.first == c.args[i].get()) t = &tp.second;
          if (t) out_ << "&" << *t;
          else if (c.fn->params[i]->by_ref) {
            const Expr& a = *c.args[i];
            if (a.k == Expr::VarRef && a.var->by_ref && a.var->storage == Var::Param && !rename_.count(a.var)) out_ << a.var->name;
            else out_ << "&" << expr(a);
          } else out_ << expr(*c.args[i]);
        }
        out_ << ");\n";
        for (const auto& tp : temps)
          for (int k = 0; k < tp.first->nswz; ++k) {
            indent(d);
            out_ << expr(*tp.first->args[0]) << "." << "xyzw"[tp.first->swz[k]] << " = " << tp.second << "." << "xyzw"[k] << ";\n";
          }
        break;
      }
      case Stmt::Return:
        if (s.a) out_ << "return " << expr(*s.a) << ";\n"; else out_ << "return;\n";
        break;
      case Stmt::Break: out_ << "break;\n"; break;
      case Stmt::Continue: out_ << "continue;\n"; break;
      case Stmt::Discard: out_ << "discard;\n"; break;
      case Stmt::If: emit_if(s, d); out_ << "\n"; break;
      case Stmt::For:
        out_ << "for (";
        if (s.init) header_stmt(*s.init, d);
        out_ << "; ";
        if (s.a) out_ << expr(*s.a);
        out_ << "; ";
        if (s.cont) header_stmt(*s.cont, d);
        out_ << ") ";
        block(*s.body[0], d);
        out_ << "\n";
        break;
      case Stmt::While:
        out_ << "while " << expr(*s.a) << " ";
        block(*s.body[0], d);
        out_ << "\n";
        break;
      case Stmt::DoWhile:
        out_ << "loop {\n";
        for (const StmtP& c : s.body[0]->body) stmt(*c, d + 1);
        indent(d + 1);
        out_ << "continuing {\n";
        indent(d + 2);
        out_ << "break if !(" << expr(*s.a) << ");\n";
        indent(d + 1);
        out_ << "}\n";
        indent(d);
        out_ << "}\n";
        break;
      case Stmt::Switch: {
        out_ << "switch " << expr(*s.a) << " {\n";
        const bool uns = s.a->ty.sk == Sk::U32;
        bool has_default = false;
        for (const StmtP& c : s.body) {
          indent(d + 1);
          out_ << (c->case_values.empty() ? "default" : "case ");
          for (size_t i = 0; i < c->case_values.size(); ++i) out_ << (i ? ", " : "") << (uns ? std::to_string((uint32_t)c->case_values[i]) + "u" : std::to_string((int32_t)c->case_values[i]) + "i");
          if (c->is_default && !c->case_values.empty()) out_ << ", default";
          has_default = has_default || c->is_default;
          out_ << ": ";
          block(*c->body[0], d + 1);
          out_ << "\n";
        }
        if (!has_default) { indent(d + 1); out_ << "default: {\n"; indent(d + 1); out_ << "}\n"; }
        indent(d);
        out_ << "}\n";
        break;
      }
      case Stmt::Case: throw FrontendError(4, "internal: case outside of a switch");
      case Stmt::Loop:
        out_ << "loop {\n";
        for (const StmtP& c : s.body[0]->body) stmt(*c, d + 1);
        if (s.cont || s.break_if) {
          indent(d + 1);
          out_ << "continuing {\n";
          if (s.cont) for (const StmtP& c : s.cont->body) stmt(*c, d + 2);
          if (s.break_if) { indent(d + 2); out_ << "break if " << expr(*s.break_if) << ";\n"; }
          indent(d + 1);
          out_ << "}\n";
        }
        indent(d);
        out_ << "}\n";
        break;
    }
  }

  void emit_if(const Stmt& s, int d) {
    out_ << "if " << expr(*s.a) << " ";
    block(*s.then_s, d);
    if (s.else_s) {
      out_ << " else ";
      if (s.else_s->k == Stmt::If) emit_if(*s.else_s, d);
      else block(*s.else_s, d);
    }
  }
};

}  // namespace

std::string emit_wgsl(const Module& m) { return WgslWriter(m).run(); }

// A GLSL program may name things `f32`, `fn`, `target`, ...: legal there, not in WGSL.  naga's namer
// (which writes the text the reference dumps with --debug-wgsl and compiles,
// /root/reference/src/shadertoy.rs:169-194) appends `_` to such names; so does this pass, over every
// name the writer prints: functions, variables, parameters, struct types and their fields.
void make_names_wgsl_safe(Module& m) {
  static const std::set<std::string> reserved = {
      // keywords
      "alias", "break", "case", "const", "const_assert", "continue", "continuing", "default", "diagnostic", "discard", "else",
      "enable", "false", "fn", "for", "if", "let", "loop", "override", "requires", "return", "struct", "switch", "true", "var", "while",
      // predeclared types and type generators
      "bool", "f16", "f32", "i32", "u32", "vec2", "vec3", "vec4", "vec2f", "vec3f", "vec4f", "vec2i", "vec3i", "vec4i", "vec2u", "vec3u",
      "vec4u", "vec2h", "vec3h", "vec4h", "mat2x2", "mat2x3", "mat2x4", "mat3x2", "mat3x3", "mat3x4", "mat4x2", "mat4x3", "mat4x4",
      "mat2x2f", "mat2x3f", "mat2x4f", "mat3x2f", "mat3x3f", "mat3x4f", "mat4x2f", "mat4x3f", "mat4x4f", "mat2x2h", "mat3x3h", "mat4x4h",
      "array", "atomic", "ptr", "sampler", "sampler_comparison", "texture_1d", "texture_2d", "texture_2d_array", "texture_3d",
      "texture_cube", "texture_cube_array", "texture_multisampled_2d", "texture_storage_1d", "texture_storage_2d",
      "texture_storage_2d_array", "texture_storage_3d", "texture_depth_2d", "texture_depth_2d_array", "texture_depth_cube",
      "texture_depth_cube_array", "texture_depth_multisampled_2d", "bitcast",
      // reserved words (WGSL specification, section "Reserved Words")
      "NULL", "Self", "abstract", "active", "alignas", "alignof", "as", "asm", "asm_fragment", "async", "attribute", "auto", "await",
      "become", "binding_array", "cast", "catch", "class", "co_await", "co_return", "co_yield", "coherent", "column_major", "common",
      "compile", "compile_fragment", "concept", "const_cast", "consteval", "constexpr", "constinit", "crate", "debugger", "decltype",
      "delete", "demote", "demote_to_helper", "do", "dynamic_cast", "enum", "explicit", "export", "extends", "extern", "external",
      "fallthrough", "filter", "final", "finally", "friend", "from", "fxgroup", "get", "goto", "groupshared", "highp", "impl",
      "implements", "import", "inline", "instanceof", "interface", "layout", "lowp", "macro", "macro_rules", "match", "mediump", "meta",
      "mod", "module", "move", "mut", "mutable", "namespace", "new", "nil", "noexcept", "noinline", "nointerpolation", "noperspective",
      "null", "nullptr", "of", "operator", "package", "packoffset", "partition", "pass", "patch", "pixelfragment", "precise",
      "precision", "premerge", "priv", "protected", "pub", "public", "readonly", "ref", "regardless", "register", "reinterpret_cast",
      "require", "resource", "restrict", "self", "set", "shared", "sizeof", "smooth", "snorm", "static", "static_assert", "static_cast",
      "std", "subroutine", "super", "target", "template", "this", "thread_local", "throw", "trait", "try", "type", "typedef", "typeid",
      "typename", "typeof", "union", "unless", "unorm", "unsafe", "unsized", "use", "using", "varying", "virtual", "volatile", "wgsl",
      "where", "with", "writeonly", "yield"};
  std::set<std::string> taken;
  for (const auto& v : m.vars) taken.insert(v->name);
  for (const auto& f : m.functions) taken.insert(f->name);
  for (const auto& sd : m.structs) taken.insert(sd->name);
  std::map<std::string, std::string> renamed;  // one new name per old name, so that equal names stay equal
  auto safe = [&](std::string& name) {
    if (!reserved.count(name)) return;
    auto it = renamed.find(name);
    if (it == renamed.end()) {
      std::string n = name + "_";
      while (taken.count(n)) n += "_";
      taken.insert(n);
      it = renamed.emplace(name, n).first;
    }
    name = it->second;
  };
  for (auto& v : m.vars) safe(v->name);
  for (auto& f : m.functions)
    if (!f->is_entry && !f->builtin_lib) safe(f->name);
  for (auto& sd : m.structs) {
    safe(sd->name);
    std::set<std::string> fields(sd->field_names.begin(), sd->field_names.end());
    for (std::string& fld : sd->field_names)
      if (reserved.count(fld)) {
        std::string n = fld + "_";
        while (fields.count(n)) n += "_";
        fields.insert(n);
        fld = n;
      }
  }
}

}  // namespace s2m_frontend
