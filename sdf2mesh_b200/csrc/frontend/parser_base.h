// parser_base.h -- token cursor, scopes and shared statement helpers for the two parsers.
#pragma once
#include <map>
#include <set>
#include <string>
#include <vector>

#include "builder.h"
#include "lexer.h"

namespace s2m_frontend {

class ParserBase {
 public:
  ParserBase(Lang lang, Module* m) : b(lang), mod(m) { b.module = m; }

 protected:
  Builder b;
  Module* mod;
  std::vector<Token> toks;
  size_t pos = 0;
  std::vector<std::map<std::string, Var*>> scopes;   // innermost last; scopes[0] = module scope
  std::map<std::string, Function*> functions;
  std::map<std::string, const StructDef*> structs;
  Function* cur_fn = nullptr;
  int loop_depth = 0;
  int switch_depth = 0;

  const Token& peek(size_t o = 0) const { return toks[std::min(pos + o, toks.size() - 1)]; }
  const Token& advance() { const Token& t = toks[pos]; if (pos + 1 < toks.size()) ++pos; b.cur_line = t.line; return t; }
  bool is_punct(const char* p, size_t o = 0) const { return peek(o).k == Token::Punct && peek(o).text == p; }
  bool is_ident(const char* p, size_t o = 0) const { return peek(o).k == Token::Ident && peek(o).text == p; }
  bool accept(const char* p) { if (is_punct(p)) { advance(); return true; } return false; }
  bool accept_ident(const char* p) { if (is_ident(p)) { advance(); return true; } return false; }
  [[noreturn]] void perr(const std::string& msg) const {
    const Token& t = peek();
    throw FrontendError(3, "parse error at " + std::to_string(t.line) + ":" + std::to_string(t.col) + ": " + msg +
                               (t.k == Token::End ? " (at end of input)" : " (at '" + t.text + "')"));
  }
  void expect(const char* p) { if (!accept(p)) perr(std::string("expected '") + p + "'"); }
  std::string expect_ident(const char* what) {
    if (peek().k != Token::Ident) perr(std::string("expected ") + what);
    return advance().text;
  }
  Var* lookup(const std::string& name) const {
    for (size_t i = scopes.size(); i-- > 0;) {
      auto it = scopes[i].find(name);
      if (it != scopes[i].end()) return it->second;
    }
    return nullptr;
  }
  Var* declare(const std::string& name, Type ty, Var::Storage st) {
    if (scopes.back().count(name)) b.error("redefinition of '" + name + "'");
    Var* v = mod->new_var();
    v->name = name; v->ty = ty; v->storage = st;
    scopes.back()[name] = v;
    return v;
  }
  void push_scope() { scopes.emplace_back(); }
  void pop_scope() { scopes.pop_back(); }
  // skip a balanced {...} block starting at '{'
  void skip_braces() {
    expect("{");
    int depth = 1;
    while (depth > 0) {
      if (peek().k == Token::End) perr("unterminated block");
      if (is_punct("{")) ++depth;
      if (is_punct("}")) --depth;
      advance();
    }
  }
  static void mark_written(const Expr& lhs) {
    const Expr* e = &lhs;
    while (e->k == Expr::Swizzle || e->k == Expr::Deref || e->k == Expr::AddrOf || e->k == Expr::Member || e->k == Expr::Index) e = e->args[0].get();
    if (e->k == Expr::VarRef) e->var->written = true;
  }
  StmtP mk_stmt(Stmt::K k) { StmtP s = std::make_shared<Stmt>(); s->k = k; s->line = b.cur_line; return s; }
  StmtP make_assign(ExprP lhs, ExprP rhs) {
    if (!Builder::is_lvalue(*lhs)) b.error("left-hand side is not assignable");
    if (lhs->k == Expr::Swizzle) {
      for (int i = 0; i < lhs->nswz; ++i)
        for (int j = i + 1; j < lhs->nswz; ++j)
          if (lhs->swz[i] == lhs->swz[j]) b.error("repeated component in swizzle assignment");
      if (b.lang == Lang::Wgsl && lhs->nswz > 1) b.error("WGSL cannot assign to a multi-component swizzle");
    }
    rhs = b.coerce(rhs, lhs->ty, "assignment");
    mark_written(*lhs);
    StmtP s = mk_stmt(Stmt::Assign);
    s->a = lhs; s->b = rhs;
    return s;
  }
  StructDef* declare_struct(const std::string& name) {
    if (structs.count(name)) b.error("redefinition of struct '" + name + "'");
    mod->structs.emplace_back(new StructDef());
    mod->structs.back()->name = name;
    structs[name] = mod->structs.back().get();
    return mod->structs.back().get();
  }
  void add_field(StructDef* d, const std::string& fname, const Type& t) {
    if (t.is_void()) b.error("struct member of type void");
    if (d->field(fname) >= 0) b.error("duplicate member '" + fname + "' in struct " + d->name);
    d->field_names.push_back(fname);
    d->field_types.push_back(t);
  }
  // array length from a constant expression
  int array_len(ExprP e) {
    ConstVal cv;
    e = b.concretize(e);
    if (!e->ty.is_scalar() || !e->ty.is_int() || !b.const_eval(*e, &cv)) b.error("array length must be a constant integer expression");
    if (cv.i[0] < 1 || cv.i[0] > 65536) b.error("array length " + std::to_string(cv.i[0]) + " out of range");
    return (int)cv.i[0];
  }
  // selector of a switch: concrete i32 / u32 scalar
  ExprP switch_selector(ExprP e) {
    e = b.concretize(e);
    if (!e->ty.is_scalar() || !(e->ty.sk == Sk::I32 || e->ty.sk == Sk::U32)) b.error("switch selector must be an integer scalar, found " + e->ty.str());
    return e;
  }
  int64_t case_value(ExprP e, const Type& sel) {
    e = b.coerce(e, sel, "case selector");
    ConstVal cv;
    if (!b.const_eval(*e, &cv)) b.error("case selector is not a constant expression");
    return cv.i[0];
  }
  void check_cases(const Stmt& sw) {
    std::set<int64_t> seen;
    int defaults = 0;
    for (const StmtP& c : sw.body) {
      if (c->is_default) ++defaults;
      for (int64_t v : c->case_values) if (!seen.insert(v).second) b.error("duplicate case value " + std::to_string(v));
    }
    if (defaults > 1) b.error("switch has more than one default");
  }
  static Op compound_op(const std::string& p) {
    switch (p[0]) {
      case '+': return Op::Add; case '-': return Op::Sub; case '*': return Op::Mul; case '/': return Op::Div;
      case '%': return Op::Rem; case '&': return Op::BitAnd; case '|': return Op::BitOr; case '^': return Op::BitXor;
      case '<': return Op::Shl; default: return Op::Shr;
    }
  }
};

}  // namespace s2m_frontend
